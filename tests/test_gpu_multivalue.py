"""Multi-value bootstrap: all tables that share a lincomb (reference fbs_exec_env.py:93-100)
are evaluated by ONE blind rotation + one sparse product / sample extraction per table.  Ciphertext-level parity against the
oracle's restatement (oracle/tfhe_ref.c: ref_pbs_multi), decrypted parity against the cleartext interpreter."""
import numpy as np
import pytest

from conftest import load_ref_mapped, selfcheck_inputs, unpack_outputs
from oracle import cleartext
from oracle.tfhe_ref import RefTFHE
from tfhe_fbs_map_b200 import levelize, params
from tfhe_fbs_map_b200.formats import read_lbf

pytestmark = pytest.mark.gpu


def table_cases(p, rng):
    low = [int(x) for x in rng.integers(0, 2, p)]
    return [(low, 1), (low + [1 - x for x in low], 1), ([0] + low[1:] + [0], 0), ([1] + low[1:] + [1], 2), (low[:max(1, p - 2)], 1)]


@pytest.mark.parametrize("name,cluster", [("toy3", 0), ("toy1", 0), ("toy2", 0), ("toy3u", 0), ("toy3v", 0), ("toy5v", 1), ("toy5v", 4), ("toy5u", 2)])
def test_multi_value_bit_exact_against_oracle(name, cluster):
    """Every blind-rotation kernel family (classic, key-unrolled, cluster-split) in multi-value mode: accumulator and all T
    extracted ciphertexts equal the oracle's; all of them decrypt to table[m]."""
    from tfhe_fbs_map_b200.backend import B200Backend
    be = B200Backend(name, device=0, seed=808)
    ref = RefTFHE(params.get(name), seed=808)
    try:
        be.set_cluster(cluster)
        for p in (5, 7):
            cases = table_cases(p, np.random.default_rng(p))
            T = len(cases)
            msgs = np.arange(2 * p, dtype=np.int32)
            count = len(msgs)
            tabs = np.zeros((count, T, 2 * p), np.uint8); lens = np.zeros((count, T), np.uint8); modes = np.zeros((count, T), np.int32)
            for i, (t, md) in enumerate(cases):
                tabs[:, i, :len(t)] = t; lens[:, i] = len(t); modes[:, i] = md
            cts = ref.encrypt(p, msgs, np.arange(count), 3)
            out, acc = be.debug_pbs_multi(p, cts, tabs, lens, modes)
            for m in range(count):
                routs, racc = ref.pbs_multi(p, cts[m], tabs[m], lens[m], modes[m])
                assert np.array_equal(acc[m], racc), f"{name} p={p} m={m}: accumulator"
                assert np.array_equal(out[m], routs), f"{name} p={p} m={m}: extracted ciphertexts"
                dec = be.debug_decrypt(p, out[m])
                for i, (t, md) in enumerate(cases):
                    if m < len(t):
                        assert dec[i] == t[m], f"{name} p={p} m={m} table {i}"
    finally:
        be.close()


def test_two_input_gates_use_two_rotations_at_full_size():
    """`_2_input_gates` mapped at fbs_size 15 by the reference mapper has 10 bootstraps on 2 lincombs: multi-value evaluation
    runs 2 blind rotations per instance instead of 10 and decrypts to the reference's outputs; so does a circuit without any
    sharing (groups of one)."""
    from tfhe_fbs_map_b200.backend import B200Backend
    be = B200Backend(params.DEFAULT_SET, device=0, seed=2024)
    try:
        for circuit, p in (("_2_input_gates", 15), ("aes_sbox", 11)):
            e = next(x for x in load_ref_mapped() if x["circuit"] == circuit and x["p"] == p and x["mapper"] == "search" and not x.get("strict"))
            env = read_lbf(e["lbf"])
            B = 48
            inputs = {k: v[:B] for k, v in selfcheck_inputs(e["input_names"]).items()}
            want = unpack_outputs(e, batch=B)
            prog = levelize(env, p, multi_value=True)
            got = env.eval(inputs, fbs_size=p, backend=be, multi_value=True)
            rot_mv = be.last_stats["n_pbs"]
            for k in got:
                assert np.array_equal(np.asarray(got[k]), want[str(k)]), (circuit, k)
            got1 = env.eval(inputs, fbs_size=p, backend=be)
            rot_1 = be.last_stats["n_pbs"]
            for k in got1:
                assert np.array_equal(np.asarray(got1[k]), want[str(k)]), (circuit, k)
            assert rot_mv == prog.n_groups * B and rot_1 == prog.n_boots * B
            if circuit == "_2_input_gates":
                assert (prog.n_boots, prog.n_groups) == (10, 2)
    finally:
        be.close()


def test_multi_value_program_equals_oracle_program():
    """Whole program, toy size: GPU multi-value evaluation == the oracle's multi-value evaluation == cleartext."""
    from tfhe_fbs_map_b200.backend import B200Backend
    be = B200Backend("toy3v", device=0, seed=55)
    ref = RefTFHE(params.get("toy3v"), seed=55)
    try:
        e = next(x for x in load_ref_mapped() if x["circuit"] == "_2_input_gates" and x["p"] == 11 and x["mapper"] == "search")
        env = read_lbf(e["lbf"])
        prog = levelize(env, 11, multi_value=True)
        cp = be.load(prog)
        B = 8
        inputs = selfcheck_inputs(e["input_names"])
        bits = np.array([inputs[nm][:B] for nm in prog.input_names], dtype=np.uint8)
        got = be.eval_bits(cp, bits, enc_seed=12)
        assert np.array_equal(got, ref.eval_prog(prog, bits, enc_seed=12, multi_value=True))
        want = cleartext.lut_eval(env, {nm: bits[i] for i, nm in enumerate(prog.input_names)})
        for nm in prog.output_names:
            assert np.array_equal(got[prog.out_index[nm]], np.asarray(want[nm])), nm
    finally:
        be.close()
