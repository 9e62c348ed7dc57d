"""BASELINE.json configs[4]: PBS micro-benchmark sweep over p in {3..17} and batch sizes, one process per GPU.

    python tools/pbs_sweep.py [param set] [batches]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pbs_sweep.py ...

Every bootstrap is checked: decrypt(PBS(enc(m), table)) == table[m]; failures are counted (expected 0).  Under torchrun every
rank sweeps its own GPU with its own messages/tables (`batch` = bootstraps PER GPU, keys replicated by the seed, no collective on
the data path); rank 0 prints one JSON line per (p, batch) with the aggregate PBS/s (sum over GPUs of count / max time) and the
per-GPU spread."""
import json, os, sys
import numpy as np
sys.path.insert(0, ".")
from tfhe_fbs_map_b200.backend import B200Backend
from tfhe_fbs_map_b200 import params

name = (sys.argv[1] if len(sys.argv) > 1 else "") or params.DEFAULT_SET
batches = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 64, 148, 296, 1184, 4736, 16384, 65536]
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ps = params.get(name)
be = B200Backend(name, device=local, seed=5)
peak = be.measure_int_peak()
rng = np.random.default_rng(1000 * rank)
for p in (3, 5, 7, 9, 11, 13, 15, 17):
    for count in batches:
        # half the tables use the negacyclic extension (length 2p, f(x+p) = 1 - f(x)), messages uniform over the table
        low = rng.integers(0, 2, (count, p)).astype(np.uint8)
        tables = np.concatenate([low, 1 - low], axis=1)
        lens = np.where(np.arange(count) % 2 == 0, p, 2 * p).astype(np.uint8)
        msgs = (rng.integers(0, 1 << 30, count) % lens).astype(np.uint8)
        best = None
        if dist is not None:
            dist.barrier()
        for rep in range(2 if count >= 4736 else 3):
            out = be.pbs_batch(p, msgs, tables, lens)
            st = be.last_stats
            ms = st["ms_lincomb"] + st["ms_keyswitch"] + st["ms_blind_rotate"]
            best = ms if best is None else min(best, ms)
        fails = int((out != tables[np.arange(count), msgs]).sum())
        mine = dict(ms=best, failures=fails, ks=st["ms_keyswitch"], br=st["ms_blind_rotate"])
        allr = [mine]
        if dist is not None:
            allr = [None] * world
            dist.all_gather_object(allr, mine)
        if rank == 0:
            tmax = max(r["ms"] for r in allr)
            rate = world * count / (tmax * 1e-3)
            print(json.dumps(dict(param_set=name, p=p, n_gpus=world, batch_per_gpu=count, pbs_per_s=round(rate, 1), pbs_per_s_per_gpu=round(rate / world, 1),
                                  ms_max=round(tmax, 3), ms_min=round(min(r["ms"] for r in allr), 3), failures=sum(r["failures"] for r in allr),
                                  p_fail_model=ps.p_fail(p, 1.0), int_roofline_frac_blind_rotate=round(count / (max(r["br"] for r in allr) * 1e-3) * 4 * ps.modmul_per_pbs() / peak, 4),
                                  ms_keyswitch=round(st["ms_keyswitch"], 3), ms_blind_rotate=round(st["ms_blind_rotate"], 3))), flush=True)
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
