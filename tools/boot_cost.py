"""Measure boot_cost on the GPU: microseconds per bootstrap (saturating batch) and single-bootstrap latency for every shipped
128-bit-secure parameter set, plus the failure-probability grid over the reference's (fbs_size, sq_norm2) range.
Writes gpurun_out/boot_cost_b200.json (copy to tfhe_fbs_map_b200/boot_cost_b200.json and profiles/)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from tfhe_fbs_map_b200.backend import B200Backend
from tfhe_fbs_map_b200 import params

out = dict(device="NVIDIA B200", unit="us per bootstrap (key switch + blind rotation + sample extraction, CUDA events)", sets={}, p_fail={})
rng = np.random.default_rng(0)
for name in ("A3", "A2", "A", "C", "S"):
    ps = params.get(name)
    be = B200Backend(name, device=0, seed=5)
    row = {}
    for p in ((15,) if name != "A3" else (3, 9, 15, 17)):
        if name == "S" and p > 8:
            p = 7
        for count in (1, 4736):
            low = rng.integers(0, 2, (count, p)).astype(np.uint8)
            tables = np.concatenate([low, 1 - low], axis=1)
            msgs = rng.integers(0, 2 * p, count).astype(np.uint8)
            best, fails = None, 0
            for rep in range(3):
                o = be.pbs_batch(p, msgs, tables, np.full(count, 2 * p, np.uint8))
                st = be.last_stats
                ms = st["ms_lincomb"] + st["ms_keyswitch"] + st["ms_blind_rotate"]
                best = ms if best is None else min(best, ms)
                fails += int((o != tables[np.arange(count), msgs]).sum())
            row[f"p{p}_batch{count}_ms"] = round(best, 4)
            row[f"p{p}_batch{count}_failures"] = fails
            if count == 1:
                row["latency_ms"] = round(best, 4)
            else:
                row["us_per_pbs"] = round(best * 1e3 / count, 3)
                row["pbs_per_s"] = round(count / (best * 1e-3), 1)
    row.update(n=ps.n, k=ps.k, N=ps.N, bsk_l=ps.bsk_l, bsk_beta=ps.bsk_beta, ks_l=ps.ks_l, ks_beta=ps.ks_beta, bsk_unroll=ps.bsk_unroll,
               modmul_per_pbs=ps.modmul_per_pbs(), bsk_mb=round(ps.bsk_bytes / 1e6, 1), ksk_mb=round(ps.ksk_bytes / 1e6, 1))
    out["sets"][name] = row
    out["p_fail"][name] = {str(p): {str(n2): ps.p_fail(p, n2) for n2 in (1, 2, 10, 35, 70, 125, 238, 281)} for p in (3, 5, 7, 9, 11, 13, 15, 17)}
    be.close()
    print(name, row, flush=True)
with open("gpurun_out/boot_cost_b200.json", "w") as f:
    json.dump(out, f, indent=1)
