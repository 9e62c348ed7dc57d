"""One launch configuration of the blind rotation, for ncu: python tools/probe_one.py <param set> <count> <cluster mode> [reps]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from tfhe_fbs_map_b200.backend import B200Backend
from tfhe_fbs_map_b200 import params
name, count, mode = sys.argv[1] or params.DEFAULT_SET, int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
be = B200Backend(name, device=0, seed=5)
be.set_cluster(mode)
p = 17
rng = np.random.default_rng(0)
low = rng.integers(0, 2, (count, p)).astype(np.uint8)
tables = np.concatenate([low, 1 - low], axis=1)
msgs = rng.integers(0, 2 * p, count).astype(np.uint8)
for _ in range(reps):
    out = be.pbs_batch(p, msgs, tables, np.full(count, 2 * p, np.uint8))
    print(be.last_stats["ms_blind_rotate"], int((out != tables[np.arange(count), msgs]).sum()), flush=True)
