"""Map the AES-128 circuit (tfhe_fbs_map_b200/circuits.py) with --fbs_size 11 --mapper search and store the .lbf fixture
(gzip) used by the AES workload of bench.py and tests/test_gpu_aes.py.  Takes a few minutes (host Python mapper)."""
import gzip, io, json, logging, sys, time
sys.path.insert(0, ".")
import numpy as np
from tfhe_fbs_map_b200.circuits import aes128
from tfhe_fbs_map_b200.mapper import MapToFBSHeur
from tfhe_fbs_map_b200 import levelize
logging.disable(logging.CRITICAL)
p = int(sys.argv[1]) if len(sys.argv) > 1 else 11
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 10
env = aes128(rounds=rounds)
t = time.time()
lut = MapToFBSHeur("search", fbs_size=p, max_fbs_size=2 * p, max_truth_table_size=16).map(env)
lut.remove_dangling_nodes()
dt = time.time() - t
st = lut.stats()
prog = levelize(lut, p)
print(json.dumps(dict(p=p, rounds=rounds, map_time_s=round(dt, 1), stats=st, levels=prog.n_levels, slots=prog.n_slots,
                      width_max=max(prog.level_widths), width_median=int(np.median(prog.level_widths)))))
s = io.StringIO(); lut.write_lbf(os=s)
name = f"tests/golden/lbf/aes128_r{rounds}_p{p}.lbf.gz"
with gzip.open(name, "wt") as f:
    f.write(s.getvalue())
print("wrote", name, len(s.getvalue()), "bytes uncompressed")
