"""``map_circuit.py`` CLI with the reference's flags and outputs, executing the self-check on the B200.

Mirror of reference ``fbs_mapper/map_circuit.py:92-188``::

    python -m tfhe_fbs_map_b200.map_circuit FILE [--type blif|bristol] [--fbs_size N] [--mapper basic|naive|search]
           [--strict_fbs_size] [--output F.fbs] [--output_lbf F.lbf] [--max_tt_size N] [-v ...]

Same protocol: parse, draw 1000 random input vectors with seed 42 (map_circuit.py:137-139), evaluate the source
circuit BEFORE mapping (:140), map, drop dangling nodes, print the statistics dict as the LAST stdout line (:155-159,
parsed by experiments/build_csv.py:24-25), check that the mapped circuit reproduces the source outputs (:174-180), write
``.fbs`` / ``.lbf`` (:182-188).  Added flags (defaults keep the reference's behaviour and output line):

    --exec clear|b200|none   how the self-check evaluates the circuits: ``clear`` = cleartext CUDA kernels (default),
                             ``b200`` = the mapped circuit runs ENCRYPTED (TFHE) on the GPU and is decrypted,
                             ``none`` = map and write files only (no GPU needed)
    --batch N                number of random vectors (default 1000, the reference's value)
    --param-set NAME         TFHE parameter set for --exec b200 (default: the library default, A3)
"""
import argparse
import json
import logging
import sys
import time
import traceback

import os

import numpy as np

from .bit_env import BitExecEnv
from .formats import parse_blif_file, parse_bristol
from . import mapper as map_to_fbs


def build_parser():
    parser = argparse.ArgumentParser(description="Map logic gates to Functional Boostrapping (FBS)")
    parser.add_argument("filename", help="input circuit")
    parser.add_argument("--type", choices=["blif", "bristol"], default="blif", help="format")
    parser.add_argument("--fbs_size", default=3, type=int, help="FBS size")
    parser.add_argument("--mapper", choices=["basic", "naive", "search"], default="search", help="mapping strategy")
    parser.add_argument("--strict_fbs_size", action="store_true", help="do not use anti-cyclic ring property")
    parser.add_argument("--output", help="output mapped circuit file")
    parser.add_argument("--output_lbf", help="output mapped circuit file in LBS format")
    parser.add_argument("--max_tt_size", default=16, type=int, help="maximal truth table size (log2) before bootstrapping")
    parser.add_argument("--verbose", "-v", action="count", default=0)
    # additions
    parser.add_argument("--exec", dest="exec_mode", choices=["clear", "b200", "none"], default="clear",
                        help="self-check backend: cleartext CUDA kernel (falls back to 'none' with a warning when no GPU / library is "
                             "available, so the reference invocation `map_circuit FILE --fbs_size N` works on any host), encrypted TFHE "
                             "on B200, or none")
    parser.add_argument("--batch", type=int, default=1000, help="random input vectors for the self-check")
    parser.add_argument("--param-set", default=None, help="TFHE parameter set for --exec b200 (default: the library default, A3)")
    parser.add_argument("--gpus", type=int, default=1,
                        help="--exec b200: GPUs to use (this process drives GPU 0..gpus-1 in turn for --shard instances; node sharding "
                             "needs one process per GPU: launch with torchrun, see bench.py --shard nodes)")
    parser.add_argument("--shard", choices=["instances", "nodes"], default="instances",
                        help="--exec b200 with several GPUs: split the batch of input vectors (instances, no collective) or each level's "
                             "bootstraps (nodes, under torchrun)")
    return parser


EXTRA_KEYS = ("exec_mode", "batch", "param_set", "gpus", "shard")


def main(argv=None):
    args = build_parser().parse_args(argv)
    levels = [logging.CRITICAL, logging.ERROR, logging.WARNING, logging.INFO, logging.DEBUG]
    logging.basicConfig(level=levels[min(args.verbose, len(levels) - 1)])

    fbs_size = args.fbs_size
    max_fbs_size = fbs_size if args.strict_fbs_size else 2 * fbs_size
    args.max_fbs_size = max_fbs_size

    if args.mapper == "basic":
        mapper = map_to_fbs.MapToFBSBasic()
    else:
        mapper = map_to_fbs.MapToFBSHeur(fbs_size=fbs_size, max_fbs_size=max_fbs_size,
                                         max_truth_table_size=args.max_tt_size, cone_merger=args.mapper)
    if args.type == "blif":
        bit_env = parse_blif_file(args.filename)
    else:
        with open(args.filename) as f:
            bit_env = parse_bristol(f.read())

    np.random.seed(42)
    input_vals = {inp.name: np.random.randint(0, 2, (args.batch)) for inp in bit_env.inputs}
    backend = None
    output_values1 = None
    from . import params as _params
    if args.param_set is None:
        args.param_set = _params.DEFAULT_SET
    backends = []
    if args.exec_mode != "none":
        from . import backend as _be
        try:
            n_gpus = max(1, args.gpus) if args.exec_mode == "b200" and args.shard == "instances" else 1
            backends = [_be.B200Backend(args.param_set, device=d, seed=0xC11 if n_gpus > 1 else None, keygen=(args.exec_mode == "b200"))
                        for d in range(n_gpus)]
            backend = backends[0]
        except (RuntimeError, OSError) as ex:
            if args.exec_mode == "b200":
                raise                                                  # the encrypted executor has no CPU fallback
            # --exec clear is only the self-check of the mapping: the reference CLI has no GPU dependency at all
            print(f"warning: GPU self-check unavailable ({ex}); continuing with --exec none", file=sys.stderr)
            args.exec_mode = "none"
    if args.exec_mode != "none":
        output_values1 = bit_env.eval(input_vals, backend=backend)      # before mapping: the mapper mutates gate tables

    start = time.time()
    try:
        lut_env = mapper.map(bit_env)
    except Exception:
        logging.critical(traceback.format_exc())
        sys.exit()
    lut_env.remove_dangling_nodes()
    duration = time.time() - start

    stats = lut_env.stats()
    stats.update({k: v for k, v in args.__dict__.items() if k not in EXTRA_KEYS})
    stats["time"] = duration
    if args.exec_mode != "b200":
        print(stats)

    if args.exec_mode != "none":
        t0 = time.time()
        if args.exec_mode == "b200":
            p = max(fbs_size, 2)
            if args.shard == "nodes" and int(os.environ.get("WORLD_SIZE", "1")) <= 1 and args.gpus > 1:
                print("warning: --shard nodes needs one process per GPU (torchrun); evaluating on one GPU", file=sys.stderr)
            if len(backends) > 1:
                # instance sharding from one process: contiguous slices of the batch, one per GPU, keys replicated by the seed
                from .dist import instance_shard
                from concurrent.futures import ThreadPoolExecutor
                names = [i.name for i in bit_env.inputs]

                def part(r):
                    off, cnt = instance_shard(args.batch, len(backends), r)
                    return lut_env.eval({nm: input_vals[nm][off:off + cnt] for nm in names}, fbs_size=p, backend=backends[r]) if cnt else None
                with ThreadPoolExecutor(len(backends)) as pool:
                    parts = [x for x in pool.map(part, range(len(backends))) if x is not None]
                output_values2 = {k: (np.concatenate([np.atleast_1d(x[k]) for x in parts]) if isinstance(parts[0][k], np.ndarray) else parts[0][k])
                                  for k in parts[0]}
                st = dict(n_pbs=sum((b.last_stats or {}).get("n_pbs", 0) for b in backends),
                          ms_total=max((b.last_stats or {}).get("ms_total", 0.0) for b in backends))
            else:
                output_values2 = lut_env.eval(input_vals, fbs_size=p, backend=backend)
                st = dict(backend.last_stats or {})
            wall = time.time() - t0
            n_pbs = st.get("n_pbs", 0)
            info = dict(exec="b200", param_set=args.param_set, gpus=len(backends), shard=args.shard, batch=args.batch, n_pbs=n_pbs, wall_s=wall,
                        pbs_per_s=n_pbs / max(st.get("ms_total", 0.0) * 1e-3, 1e-9), evals_per_s=args.batch / max(wall, 1e-9),
                        p_fail_per_pbs=backend.params.p_fail(p, stats["norm2_linprod"]))
            print(json.dumps(info))
        else:
            output_values2 = lut_env.eval_clear(input_vals, backend=backend)
        assert output_values1.keys() == output_values2.keys()
        for k in output_values1.keys():
            equal = np.all(output_values1[k] == output_values2[k])
            if not equal:
                print(f"output {k} do not match {output_values1[k]} {output_values2[k]}")
            assert equal
        if args.exec_mode == "b200":
            print(stats)

    if args.output is not None:
        with open(args.output, "w") as file:
            lut_env.print(show_outputs=True, os=file)
    if args.output_lbf is not None:
        with open(args.output_lbf, "w") as file:
            lut_env.write_lbf(os=file)
    return stats


if __name__ == "__main__":
    main()
