"""Map the 64x64 array multiplier (tfhe_fbs_map_b200/circuits.py: synthetic stand-in for EPFL multiplier.blif, BASELINE
configs[3]) with the REFERENCE mapper (--fbs_size 17 --mapper search) in this container and freeze the result as
tests/golden/lbf/mult64_p17.lbf.gz (+ its entry in tests/golden/lbf/index.json: stats and the sha256 of the reference's own
LutExecEnv.eval outputs under the CLI self-check protocol, map_circuit.py:137-140,174-180).

TEST / BENCH INFRASTRUCTURE ONLY (needs /root/reference; takes minutes):   python oracle/gen_mult64.py [nbits] [p]
"""
import gzip
import io
import json
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as gg           # noqa: E402  (puts the reference on sys.path)

from tfhe_fbs_map_b200 import circuits   # noqa: E402


def main():
    nbits = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 17
    logging.disable(logging.CRITICAL)
    env = circuits.array_multiplier(nbits)
    s = io.StringIO()
    env.to_blif(fs=s, model_name=f"mult{nbits}")
    e, out2, _ = gg.map_with_reference(s.getvalue(), p, "search")
    fn = f"mult{nbits}_p{p}.lbf.gz"
    with gzip.open(os.path.join(gg.GOLD, "lbf", fn), "wt") as f:
        f.write(e["lbf"])
    idx_path = os.path.join(gg.GOLD, "lbf", "index.json")
    index = [x for x in json.load(open(idx_path)) if x["file"] != fn]
    index.append(dict(file=fn, circuit=f"mult{nbits}", p=p, stats=e["stats"], out_sha256=e["out_sha256"],
                      input_names=e["input_names"], map_time=e["map_time"]))
    with open(idx_path, "w") as f:
        json.dump(index, f, indent=1)
    print(fn, e["stats"], round(e["map_time"], 1), "s")


if __name__ == "__main__":
    main()
