"""Latency of narrow launches: ms of the blind rotation for `count` independent bootstraps, with the cluster-split kernel
forced off / to 2 / 4 / 8 CTAs per bootstrap / auto.  One JSON line per (count, mode)."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
from tfhe_fbs_map_b200.backend import B200Backend
from tfhe_fbs_map_b200 import params

name = (sys.argv[1] if len(sys.argv) > 1 else "") or params.DEFAULT_SET
counts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8, 18, 37, 64, 74, 148]
be = B200Backend(name, device=0, seed=5)
p = 17
rng = np.random.default_rng(0)
for count in counts:
    low = rng.integers(0, 2, (count, p)).astype(np.uint8)
    tables = np.concatenate([low, 1 - low], axis=1)
    lens = np.full(count, 2 * p, np.uint8)
    msgs = rng.integers(0, 2 * p, count).astype(np.uint8)
    for mode in (1, 2, 4, 8, 12, 14, 18, 0):
        if mode > 1 and count * (mode % 10) > 4 * 148:
            continue
        be.set_cluster(mode)
        best, fails = None, 0
        for rep in range(4):
            out = be.pbs_batch(p, msgs, tables, lens)
            st = be.last_stats
            best = st["ms_blind_rotate"] if best is None else min(best, st["ms_blind_rotate"])
            fails += int((out != tables[np.arange(count), msgs]).sum())
        print(json.dumps(dict(param_set=name, batch=count, cluster=mode, ms_blind_rotate=round(best, 4), ms_keyswitch=round(st["ms_keyswitch"], 4),
                              ms_lincomb=round(st["ms_lincomb"], 4), pbs_per_s=round(count / (best * 1e-3), 1), failures=fails)), flush=True)
be.set_cluster(0)
