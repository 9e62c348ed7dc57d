"""Per-phase clock split of k_blind_rotate2 (profiling build only):
    nvcc ... -DFBS_PHASE_CLK=1 -o build_exp/libfbs_phase.so tfhe_fbs_map_b200/csrc/api.cu
    FBS_B200_LIB=$PWD/build_exp/libfbs_phase.so python tools/phase_clock.py [param set] [count]
Warp 0 of every CTA accumulates SM clocks per phase of a blind-rotation step; prints the shares and clocks per step."""
import ctypes, json, sys
import numpy as np
sys.path.insert(0, ".")
from tfhe_fbs_map_b200 import backend, params
name = (sys.argv[1] if len(sys.argv) > 1 else "") or params.DEFAULT_SET
count = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
be = backend.B200Backend(name, device=0, seed=1)
lib = backend.load_library()
lib.fbs_debug_phase_clk.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
buf = (ctypes.c_ulonglong * 8)()
p = 17; rng = np.random.default_rng(0)
msgs = rng.integers(0, p, count).astype(np.uint8)
tables = rng.integers(0, 2, (count, 2 * p)).astype(np.uint8)
lens = np.full(count, p, np.uint8)
be.pbs_batch(p, msgs, tables, lens)
lib.fbs_debug_phase_clk(buf)                       # reset after the warm-up launch
out = be.pbs_batch(p, msgs, tables, lens)
lib.fbs_debug_phase_clk(buf)
v = [int(x) for x in buf][:5]
ps = params.get(name)
steps = (ps.n + ps.bsk_unroll - 1) // ps.bsk_unroll
ctas = (count + 1) // 2
tot = sum(v[:4])
names = ["accumulate+decompose", "forward NTT", "exchange+point-wise (incl. key waits)", "inverse NTT", "  of which waiting for key blocks"]
print(json.dumps(dict(set=name, count=count, ok=int((out == tables[np.arange(count), msgs]).sum()), clocks_per_step=round(tot / ctas / steps, 1),
                      phases={n: dict(clocks_per_step=round(x / ctas / steps, 1), share=round(x / tot, 4)) for n, x in zip(names, v)},
                      ms_blind_rotate=be.last_stats["ms_blind_rotate"])))
