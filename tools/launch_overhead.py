"""Is the per-level launch path on the critical path?  adder128_p15 at batch 1 (127 levels x 3-4 launches): host time to
ENQUEUE one fbs_run (returns before the GPU finishes) vs GPU time of the run (CUDA events)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from tfhe_fbs_map_b200 import levelize, params
from tfhe_fbs_map_b200.backend import B200Backend, RunStats
from tfhe_fbs_map_b200.formats import read_lbf_file

be = B200Backend(params.DEFAULT_SET, device=0, seed=5)
env = read_lbf_file("tests/golden/lbf/adder128_p15.lbf")
prog = levelize(env, 15, preserve_inputs=True)
cp = be.load(prog)
for B in (1, 16):
    bits = np.random.default_rng(0).integers(0, 2, (prog.n_inputs, B)).astype(np.uint8)
    wires = torch.empty(be.wires_bytes(cp, B) // 8, dtype=torch.int64, device="cuda")
    d_in = torch.from_numpy(bits).cuda()
    sp = torch.cuda.current_stream().cuda_stream
    be.encrypt_inputs(cp, d_in.data_ptr(), B, wires.data_ptr(), stream=sp)
    be.run(cp, B, wires.data_ptr(), stream=sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    be.run(cp, B, wires.data_ptr(), stream=sp)
    t_enq = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps(dict(workload="adder128_p15", batch=B, levels=prog.n_levels, host_enqueue_ms=round(t_enq * 1e3, 3), gpu_ms=round(e0.elapsed_time(e1), 3))), flush=True)
