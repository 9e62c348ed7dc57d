"""Quick PBS throughput probe (set A) -- prints per-phase device times."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from tfhe_fbs_map_b200.backend import B200Backend
from tfhe_fbs_map_b200 import params
name = (sys.argv[1] if len(sys.argv) > 1 else "") or params.DEFAULT_SET
counts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [148, 296, 1184]
t = time.time(); be = B200Backend(name, device=0, seed=1); print("keygen s", round(time.time() - t, 2), be.info())
p = 17; rng = np.random.default_rng(0)
for count in counts:
    msgs = rng.integers(0, p, count).astype(np.uint8)
    tables = rng.integers(0, 2, (count, 2 * p)).astype(np.uint8)
    lens = np.full(count, p, np.uint8)
    for rep in range(2):
        out = be.pbs_batch(p, msgs, tables, lens)
    ok = int((out == tables[np.arange(count), msgs]).sum())
    st = be.last_stats
    print(json.dumps(dict(set=name, count=count, ok=ok, pbs_per_s=round(count / (st["ms_blind_rotate"] + st["ms_keyswitch"] + st["ms_lincomb"]) * 1e3, 1), **{k: round(v, 3) if isinstance(v, float) else v for k, v in st.items()})))
