"""BASELINE.json configs[4]: PBS micro-benchmark sweep over p in {3..17} and batch sizes, one GPU per process.
Every bootstrap is checked: decrypt(PBS(enc(m), table)) == table[m]; failures are counted (expected 0)."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
from tfhe_fbs_map_b200.backend import B200Backend
from tfhe_fbs_map_b200 import params

name = (sys.argv[1] if len(sys.argv) > 1 else "") or params.DEFAULT_SET
batches = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 64, 148, 296, 1184, 4736, 16384, 65536]
ps = params.get(name)
be = B200Backend(name, device=0, seed=5)
peak = be.measure_int_peak()
rng = np.random.default_rng(0)
for p in (3, 5, 7, 9, 11, 13, 15, 17):
    for count in batches:
        # half the tables use the negacyclic extension (length 2p, f(x+p) = 1 - f(x)), messages uniform over the table
        low = rng.integers(0, 2, (count, p)).astype(np.uint8)
        tables = np.concatenate([low, 1 - low], axis=1)
        lens = np.where(np.arange(count) % 2 == 0, p, 2 * p).astype(np.uint8)
        msgs = (rng.integers(0, 1 << 30, count) % lens).astype(np.uint8)
        best = None
        for rep in range(2 if count >= 4736 else 3):
            out = be.pbs_batch(p, msgs, tables, lens)
            st = be.last_stats
            ms = st["ms_lincomb"] + st["ms_keyswitch"] + st["ms_blind_rotate"]
            best = ms if best is None else min(best, ms)
        fails = int((out != tables[np.arange(count), msgs]).sum())
        rate = count / (best * 1e-3)
        print(json.dumps(dict(param_set=name, p=p, batch=count, pbs_per_s=round(rate, 1), ms=round(best, 3), failures=fails,
                              p_fail_model=ps.p_fail(p, 1.0), int_roofline_frac=round(rate * ps.mul32_per_pbs() / peak, 4),
                              ms_keyswitch=round(st["ms_keyswitch"], 3), ms_blind_rotate=round(st["ms_blind_rotate"], 3))), flush=True)
