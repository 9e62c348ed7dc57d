"""Measured bootstrap cost and parameter search: the replacement for the reference's cost hook.

The reference's only notion of execution time is ``total_cost = nb_bootstrap * boot_cost(fbs_size, norm2_linprod)``
(reference experiments/analyse_results.py:10) where ``boot_cost`` is the *complexity model* value a patched
``concrete-optimizer`` prints for ``--precision=p --sq-norm2=norm2`` (reference experiments/add_exec_estimates.py:9-16;
output columns ``k, N, n, br_l, br_b, ks_l, ks_b, cost, p_error``, reference experiments/concrete.patch:159-175).

Here ``boot_cost`` is MEASURED B200 time: ``boot_cost_b200.json`` (next to this file) holds, per shipped parameter set, the
microseconds per bootstrap at a saturating batch and the latency of one bootstrap, written by ``tools/boot_cost.py`` on the
GPU.  ``search`` explores the decomposition parameters (blind-rotation levels / base, key-switch levels / base, key bits per
step) around the shipped 128-bit-secure anchors (n, k, N, noise) under the reference's patched noise bound, and prices every
candidate in time -- measured when a kernel for its shape is compiled, otherwise scaled by its multiply count from the nearest
measured set (flagged ``measured: false``).

    python -m tfhe_fbs_map_b200.cost --precision=15 --sq-norm2=70        # prints the optimizer's line; cost in nanoseconds

so ``add_exec_estimates.py --opt tools/optimizer_b200`` works unchanged (it parses the second-to-last field).
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import asdict

from . import params as _params
from .params import ParamSet

_HERE = os.path.dirname(os.path.abspath(__file__))
TABLE_PATH = os.path.join(_HERE, "boot_cost_b200.json")
TARGET_P_ERROR = math.erfc(4.0 / math.sqrt(2.0))          # concrete's default 4 sigma (reference concrete.patch:101-102)

# (logN, k, bsk_l, unroll) shapes a blind-rotation kernel is compiled for (csrc/api.cu: g_br_variants), full-size only
COMPILED = {(11, 1, 1, 1), (11, 1, 1, 2), (11, 1, 1, 3), (11, 1, 2, 1), (10, 2, 1, 1)}
ANCHORS = ("A", "C", "S")                                  # (n, k, N, sigma_lwe, sigma_glwe): the security-relevant part


def load_table(path: str = TABLE_PATH) -> dict:
    with open(path) as f:
        return json.load(f)


def _shape(ps: ParamSet):
    return (int(math.log2(ps.N)), ps.k, ps.bsk_l, ps.bsk_unroll if ps.bsk_unroll in (2, 3) else 1)


def time_us(ps: ParamSet, table: dict | None = None) -> tuple[float, bool]:
    """Microseconds per bootstrap on one B200 at a saturating batch: (value, measured?)."""
    table = table or load_table()
    sets = table["sets"]
    if ps.name in sets:
        return sets[ps.name]["us_per_pbs"], True
    # same kernel shape as a measured set: time scales with the number of blind-rotation steps (n) and key-switch rows
    best = None
    for nm, row in sets.items():
        ref = _params.get(nm)
        scale = ps.modmul_per_pbs() / ref.modmul_per_pbs()
        same = _shape(ref) == _shape(ps)
        cand = (0 if same else 1, abs(math.log(scale)), row["us_per_pbs"] * scale, same)
        if best is None or cand < best:
            best = cand
    return best[2], False


def candidates(anchor: ParamSet):
    """Decomposition choices around a security anchor (n, k, N and the noise levels stay fixed)."""
    d0 = asdict(anchor)
    for l in (1, 2, 3):
        for beta in range(4, 25):
            if l * beta > 46 or l * beta < 12:
                continue
            for ks_l in range(2, 9):
                for ks_beta in range(1, 7):
                    if ks_l * ks_beta < 8 or ks_l * ks_beta > 30:
                        continue
                    for m in ((1, 2, 3) if l == 1 else (1,)):
                        d = dict(d0)
                        d.update(name=f"{anchor.name}/l{l}b{beta}k{ks_l}x{ks_beta}m{m}", bsk_l=l, bsk_beta=beta, ks_l=ks_l, ks_beta=ks_beta, bsk_unroll=m)
                        yield ParamSet(**d)


def search(p: int, norm2: float, target: float = TARGET_P_ERROR, compiled_only: bool = True, table: dict | None = None, mv_norm2: float = 1.0) -> dict:
    """Cheapest parameter choice (in B200 time) whose failure probability meets ``target`` for message space Z_p and lincomb
    squared norm ``norm2``.  Shipped sets are priced by measurement; other candidates by scaling (``measured`` False)."""
    table = table or load_table()
    best = None
    pool = [_params.get(nm) for nm in table["sets"]]
    for a in ANCHORS:
        pool.extend(candidates(_params.get(a)))
    for ps in pool:
        if compiled_only and _shape(ps) not in COMPILED:
            continue
        pf = ps.p_fail(p, norm2, mv_norm2)
        if not (pf <= target):
            continue
        us, measured = time_us(ps, table)
        key = (us, not measured)
        if best is None or key < best[0]:
            best = (key, ps, pf, us, measured)
    if best is None:
        raise ValueError(f"no parameter choice reaches p_error <= {target:.1e} for p={p}, sq_norm2={norm2}")
    _, ps, pf, us, measured = best
    return dict(param_set=ps.name, k=ps.k, N=ps.N, n=ps.n, br_l=ps.bsk_l, br_b=ps.bsk_beta, ks_l=ps.ks_l, ks_b=ps.ks_beta, bsk_unroll=ps.bsk_unroll,
                cost=int(round(us * 1000)), cost_unit="ns per bootstrap on one B200 (saturating batch)", p_error=pf, measured=measured,
                modmul=ps.modmul_per_pbs(), bsk_bytes=ps.bsk_bytes, ksk_bytes=ps.ksk_bytes)


def boot_cost(p: int, norm2: float, table: dict | None = None) -> int:
    """Drop-in for ``get_boot_cost`` (reference add_exec_estimates.py:9-16): nanoseconds per bootstrap, cheapest shipped set."""
    table = table or load_table()
    best = None
    for nm, row in table["sets"].items():
        ps = _params.get(nm)
        if ps.p_fail(p, norm2) <= TARGET_P_ERROR and (best is None or row["us_per_pbs"] < best):
            best = row["us_per_pbs"]
    if best is None:
        raise ValueError(f"no shipped parameter set reaches p_error <= {TARGET_P_ERROR:.1e} for p={p}, sq_norm2={norm2}")
    return int(round(best * 1000))


def optimizer_line(p: int, norm2: float) -> str:
    """One line in the patched optimizer's column order: k, N, n, br_l, br_b, ks_l, ks_b, cost, p_error."""
    r = search(p, norm2)
    return f"{r['k']}, {r['N']}, {r['n']}, {r['br_l']}, {r['br_b']}, {r['ks_l']}, {r['ks_b']}, {r['cost']}, {r['p_error']:.3e}"


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="B200 replacement of `optimizer --precision P --sq-norm2 N`")
    ap.add_argument("--precision", type=int, required=True)
    ap.add_argument("--sq-norm2", type=float, required=True)
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args(argv)
    if a.json:
        print(json.dumps(search(a.precision, a.sq_norm2)))
    else:
        print(optimizer_line(a.precision, a.sq_norm2))


if __name__ == "__main__":
    main()
