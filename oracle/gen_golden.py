"""Generate golden fixtures by RUNNING THE REFERENCE (ssmiler/tfhe_fbs_map at /root/reference) in this container.

    python oracle/gen_golden.py            # writes tests/golden/*

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so everything the tests, smoke() and
bench.py need from the reference is frozen here as small text/JSON fixtures:

* tests/golden/blif/<name>.blif      -- source circuits written by the reference's own generator
                                        (experiments/generate_benchmarks.py:39-447, BitExecEnv.to_blif)
* tests/golden/ref_mapped.json       -- for (circuit, p): the reference mapper's stats (fbs_exec_env.py:245-276), its
                                        .lbf and .fbs text (fbs_exec_env.py:158-206), and the outputs of the reference's
                                        own BitExecEnv.eval / LutExecEnv.eval under the CLI self-check protocol
                                        (map_circuit.py:137-140,174-180: seed 42, 1000 vectors)
* tests/golden/lbf/<name>_p<p>.lbf   -- larger circuits pre-mapped by the reference mapper (the synthetic stand-ins
                                        for EPFL adder / multiplier, see tfhe_fbs_map_b200/circuits.py), with expected
                                        output hashes in tests/golden/lbf/index.json
* tests/golden/demos.json            -- outputs of the reference modules' __main__ demos
"""
import hashlib
import io
import json
import os
import sys
import time

import numpy as np

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "fbs_mapper"))
sys.path.insert(0, os.path.join(REF, "experiments"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bit_exec_env as ref_bit          # noqa: E402  (reference, bare import as map_to_fbs.py:1-2 expects)
import fbs_exec_env as ref_lut          # noqa: E402
import map_to_fbs as ref_map            # noqa: E402
import generate_benchmarks as ref_gen   # noqa: E402  (reference generator; imports fbs_mapper.bit_exec_env)

from tfhe_fbs_map_b200.formats import parse_blif            # noqa: E402
from tfhe_fbs_map_b200 import circuits                      # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def selfcheck_inputs(bit_env, batch=1000):
    np.random.seed(42)                                                     # map_circuit.py:137
    return {inp.name: np.random.randint(0, 2, (batch)) for inp in bit_env.inputs}   # :138-139


def out_hash(outputs):
    h = hashlib.sha256()
    for name, arr in outputs.items():
        h.update(str(name).encode())
        h.update(np.asarray(arr).astype(np.uint8).tobytes())
    return h.hexdigest()


def pack(outputs):
    return {str(k): np.packbits(np.asarray(v).astype(np.uint8)).tobytes().hex() for k, v in outputs.items()}


def int_tables(lut_env):
    """numpy>=2 prints np.int64(1) inside tables; the pinned numpy 1.26 prints 1 (requirements.txt:12)."""
    for instr in lut_env.instructions:
        if type(instr).__name__ == "Bootstrap":
            instr.table = [int(t) for t in instr.table]


def map_with_reference(blif_text, p, mapper="search", strict=False, max_tt=16):
    bit_env = parse_blif(blif_text, env_cls=ref_bit.BitExecEnv)
    inputs = selfcheck_inputs(bit_env)
    out1 = bit_env.eval(inputs)                                            # before mapping (map_circuit.py:140, quirk Q1)
    if mapper == "basic":
        m = ref_map.MapToFBSBasic()
    else:
        m = ref_map.MapToFBSHeur(fbs_size=p, max_fbs_size=p if strict else 2 * p, max_truth_table_size=max_tt, cone_merger=mapper)
    t0 = time.time()
    lut_env = m.map(bit_env)
    lut_env.remove_dangling_nodes()
    dt = time.time() - t0
    int_tables(lut_env)
    out2 = lut_env.eval(inputs)
    assert out1.keys() == out2.keys()
    for k in out1:
        assert np.all(out1[k] == out2[k]), f"reference self-check failed on {k}"
    lbf, fbs = io.StringIO(), io.StringIO()
    lut_env.write_lbf(os=lbf)
    lut_env.print(os=fbs, show_outputs=True)
    return dict(stats=lut_env.stats(), lbf=lbf.getvalue(), fbs=fbs.getvalue(), out_sha256=out_hash(out2),
                input_names=[i.name for i in bit_env.inputs], map_time=dt), out2, bit_env


def main():
    os.makedirs(os.path.join(GOLD, "blif"), exist_ok=True)
    os.makedirs(os.path.join(GOLD, "lbf"), exist_ok=True)
    import logging
    logging.disable(logging.CRITICAL)          # the mapper logs "Cone with sparse mvt" at CRITICAL (map_to_fbs.py:202-203)

    # ---- 1. reference-generated source circuits
    gens = {"half_adder": ref_gen.half_adder_bench, "full_adder": ref_gen.full_adder_bench, "aoi21": ref_gen.aoi21_bench,
            "oai21": ref_gen.oai21_bench, "simon_iter": ref_gen.simon_iter, "_2_input_gates": ref_gen._2_input_gates,
            "ascon_lut": ref_gen.ascon_lut, "aes_sbox": ref_gen.aes_sbox,
            "trivium_iter_v1": ref_gen.TriviumIter.trivium_iter_v1, "kreyvium_iter_v1": ref_gen.KreyviumIter.kreyvium_iter_v1}
    blifs = {}
    for name, gen in gens.items():
        env = ref_gen.BitExecEnv()
        ref_gen.Bit.set_env(env)
        gen()
        env.remove_dangling_nodes()
        s = io.StringIO()
        env.to_blif(fs=s, model_name=name)
        blifs[name] = s.getvalue()
        with open(os.path.join(GOLD, "blif", f"{name}.blif"), "w") as f:
            f.write(blifs[name])

    # ---- 2. reference mapper + reference evaluators on them
    entries = []
    for name in gens:
        for p in (11, 15, 17):
            for mapper in (("search", "naive", "basic") if name in ("half_adder", "full_adder", "aes_sbox") and p == 15 else ("search",)):
                e, out2, _ = map_with_reference(blifs[name], p, mapper)
                e.update(circuit=name, p=p, mapper=mapper, outputs=pack(out2))
                entries.append(e)
                print(name, p, mapper, e["stats"])
    e, out2, _ = map_with_reference(blifs["aes_sbox"], 15, "search", strict=True)
    e.update(circuit="aes_sbox", p=15, mapper="search", strict=True, outputs=pack(out2))
    entries.append(e)
    with open(os.path.join(GOLD, "ref_mapped.json"), "w") as f:
        json.dump(entries, f, indent=0)

    # ---- 3. larger pre-mapped circuits (synthetic stand-ins for the EPFL files)
    index = []
    big = [("adder8", circuits.ripple_carry_adder(8), 15), ("adder32", circuits.ripple_carry_adder(32), 15),
           ("adder128", circuits.ripple_carry_adder(128), 15), ("adder128", circuits.ripple_carry_adder(128), 5),
           ("mult8", circuits.array_multiplier(8), 17), ("mult16", circuits.array_multiplier(16), 17)]
    for name, env, p in big:
        s = io.StringIO()
        env.to_blif(fs=s, model_name=name)
        e, out2, _ = map_with_reference(s.getvalue(), p, "search")
        fn = f"{name}_p{p}.lbf"
        with open(os.path.join(GOLD, "lbf", fn), "w") as f:
            f.write(e["lbf"])
        index.append(dict(file=fn, circuit=name, p=p, stats=e["stats"], out_sha256=e["out_sha256"],
                          input_names=e["input_names"], map_time=e["map_time"]))
        print(fn, e["stats"], round(e["map_time"], 1), "s")
    for name in ("aes_sbox", "ascon_lut", "trivium_iter_v1"):
        for p in (11, 15, 17):
            ent = next(x for x in entries if x["circuit"] == name and x["p"] == p and x["mapper"] == "search" and not x.get("strict"))
            fn = f"{name}_p{p}.lbf"
            with open(os.path.join(GOLD, "lbf", fn), "w") as f:
                f.write(ent["lbf"])
            index.append(dict(file=fn, circuit=name, p=p, stats=ent["stats"], out_sha256=ent["out_sha256"],
                              input_names=ent["input_names"], map_time=ent["map_time"]))
    with open(os.path.join(GOLD, "lbf", "index.json"), "w") as f:
        json.dump(index, f, indent=1)

    # ---- 4. __main__ demos of the reference modules
    demos = {}
    env = ref_lut.LutExecEnv()                                  # fbs_exec_env.py:279-301
    a, b, c = env.input("a"), env.input("b"), env.const(1)
    d = env.linear([1, 2], [a, b]); e_ = env.linear([1, 1], [c, d]); f_ = env.bootstrap(e_, [1, 0, 1, 1, 0])
    g = env.linear([2, 1], [a, f_]); h = env.bootstrap(g, [1, 1, 0, 2]); env.bootstrap(h, [1, 0, 1])
    env.output("f", f_); env.output("g", g); env.output("h", h)
    s = io.StringIO(); env.print(os=s)
    r = env.eval({"a": [1, 0], "b": [1, 0], "c": [1, 0]})
    demos["fbs_exec_env_main"] = dict(program=s.getvalue(), outputs={k: [int(x) for x in v] for k, v in r.items()})
    benv = ref_bit.BitExecEnv()                                 # map_to_fbs.py:550-596
    a, b, c = benv.input("a"), benv.input("b"), benv.input("c")
    d = benv.op_and(a, b); e_ = benv.op_xor(c, d); f_ = benv.op_lut([e_, d], [0, 1, 0, 0])
    benv.output("d", d); benv.output("e", e_); benv.output("f", f_)
    iv = {"a": [0, 0, 1, 1], "b": [0, 1, 0, 1], "c": [0, 0, 1, 1]}
    r0 = benv.eval(iv)
    demos["map_to_fbs_main"] = dict(inputs=iv, bit_outputs={k: [int(x) for x in v] for k, v in r0.items()})
    with open(os.path.join(GOLD, "demos.json"), "w") as f:
        json.dump(demos, f, indent=1)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
