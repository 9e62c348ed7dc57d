"""The CPU TFHE oracle (oracle/tfhe_ref.c) checked against its own ground truth: schoolbook products, exhaustive
table look-ups in all three negacyclic modes (reference map_to_fbs.py:81-98), and the cleartext interpreter
(reference fbs_exec_env.py:208-229) on reference-mapped circuits.  Toy parameter sets keep this in seconds."""
import numpy as np
import pytest

from conftest import load_ref_mapped, selfcheck_inputs, unpack_outputs
from oracle import cleartext
from oracle.tfhe_ref import RefTFHE, lib, GOLDILOCKS_P
from tfhe_fbs_map_b200 import levelize, params
from tfhe_fbs_map_b200.formats import read_lbf

P = GOLDILOCKS_P


@pytest.fixture(scope="module", params=["toy1", "toy2", "toy3", "toy2u", "toy3u", "toy7u", "toy3v"])   # *u: key-unrolled blind rotation
def ref(request):
    return RefTFHE(params.get(request.param), seed=2024)


def test_field_and_prng_primitives():
    L = lib()
    rng = np.random.default_rng(1)
    for _ in range(2000):
        a, b = int(rng.integers(0, P, dtype=np.uint64)), int(rng.integers(0, P, dtype=np.uint64))
        assert L.ref_mulmod(a, b) == (a * b) % P
    assert L.ref_mulmod(P - 1, P - 1) == 1
    assert L.ref_delta(17) == (P + 17) // 34
    assert L.ref_gadget(23, 0) == (P + (1 << 22)) >> 23 and P == 0x3FFE8001 * 0x3FFF4001
    # noise: zero mean, std == scale within 5 %
    sc = 1 << 30
    xs = np.array([L.ref_noise(7, 8, i, sc) for i in range(20000)], dtype=np.float64)
    xs = np.where(xs > P / 2, xs - P, xs)
    assert abs(xs.mean()) < 0.05 * sc and abs(xs.std() / sc - 1) < 0.05


def test_decomposition_reconstructs():
    L = lib()
    rng = np.random.default_rng(2)
    for beta, l in ((23, 1), (15, 2), (8, 3), (3, 5), (4, 4)):
        g = [L.ref_gadget(beta, j) for j in range(l)]
        d = np.zeros(l, np.int32)
        for _ in range(300):
            x = int(rng.integers(0, P, dtype=np.uint64))
            L.ref_decompose(x, beta, l, d.ctypes.data)
            assert np.all(d >= -(1 << (beta - 1))) and np.all(d < (1 << (beta - 1)))
            rec = sum(int(dj) * gj for dj, gj in zip(d, g)) % P
            err = min((rec - x) % P, (x - rec) % P)
            assert err <= (P >> (beta * l)) // 2 + (l << beta) + 2 ** 32


def test_ntt_product_equals_schoolbook(ref):
    rng = np.random.default_rng(3)
    N = ref.ps.N
    a = rng.integers(0, P, N, dtype=np.uint64)
    b = rng.integers(0, P, N, dtype=np.uint64)
    assert np.array_equal(ref.polymul_ntt(a, b), ref.polymul_schoolbook(a, b))
    assert np.array_equal(ref.ntt(ref.ntt(a), inverse=True), a)


def test_encrypt_decrypt_round_trip(ref):
    for p in (2, 5, 17):
        msgs = np.arange(2 * p, dtype=np.int32)
        cts = ref.encrypt(p, msgs, np.arange(2 * p) + 1000, 5)
        assert np.array_equal(ref.decrypt(p, cts), msgs)


@pytest.mark.parametrize("p", [2, 3, 4, 7])
def test_pbs_every_index_every_mode(ref, p):
    rng = np.random.default_rng(p)
    low = [int(x) for x in rng.integers(0, 2, p)]
    tables = [(low, 1), (low + [1 - x for x in low], 1), ([0] + low[1:] + [0], 0), ([1] + low[1:] + [1], 2),
              ([0, 2, 1][:min(3, p)], 1)]
    for tab, mode in tables:
        cts = ref.encrypt(p, np.arange(len(tab), dtype=np.int32), np.arange(len(tab)), 9)
        for m in range(len(tab)):
            out, ks, ms, acc = ref.pbs(p, cts[m], tab, mode)
            assert ref.decrypt(p, out[None, :])[0] == tab[m], (tab, mode, m)


@pytest.mark.parametrize("circuit,p", [("half_adder", 15), ("full_adder", 11), ("aoi21", 15), ("_2_input_gates", 15), ("ascon_lut", 11)])
def test_encrypted_program_equals_cleartext(circuit, p):
    ref = RefTFHE(params.get("toy3"), seed=5)     # N=512 leaves slot width 512/15 = 34 at toy noise
    e = next(x for x in load_ref_mapped() if x["circuit"] == circuit and x["p"] == p and x["mapper"] == "search")
    env = read_lbf(e["lbf"])
    prog = levelize(env, p)
    B = 12
    inputs = selfcheck_inputs(e["input_names"])            # the golden outputs are for the 1000-vector protocol
    bits = np.array([inputs[nm][:B] for nm in prog.input_names], dtype=np.uint8)
    got = ref.eval_prog(prog, bits, enc_seed=3)
    want = unpack_outputs(e, batch=B)
    for nm in prog.output_names:
        assert np.array_equal(got[prog.out_index[nm]], want[str(nm)]), nm


def test_unrolled_parameter_model():
    """Key-unrolled twins: three GGSW per key pair, same shape otherwise; the failure probability moves by a hair only
    (key-switch and modulus-switch noise dominate), the canonical multiply count drops."""
    a, a2, a3 = params.get("A"), params.get("A2"), params.get("A3")
    assert a2.n_ggsw == 3 * a.n // 2 and a2.bsk_bytes * 2 == 3 * a.bsk_bytes
    assert a3.n_ggsw == 7 * ((a.n + 2) // 3)                       # 742 key bits -> 248 triples (two zero bits of padding)
    assert a3.modmul_per_pbs() < a2.modmul_per_pbs() < 0.75 * a.modmul_per_pbs()
    for p, norm2 in ((15, 70), (17, 202), (11, 155)):
        for u in (a2, a3):
            assert a.p_fail(p, norm2) <= u.p_fail(p, norm2) < 1.5 * a.p_fail(p, norm2)
    assert params.estimate(15, 70)["param_set"] == params.DEFAULT_SET == "A3"


def test_oracle_multi_value_bootstrap_decrypts_to_every_table():
    """oracle/tfhe_ref.c: ref_pbs_multi (one blind rotation of the base polynomial, sparse product per table) decrypts to
    table[m] for every table mode, and a whole program evaluated with multi-value groups equals the cleartext interpreter."""
    from conftest import load_ref_mapped, selfcheck_inputs
    from oracle import cleartext
    from tfhe_fbs_map_b200 import levelize
    from tfhe_fbs_map_b200.formats import read_lbf
    ps = params.get("toy3")
    ref = RefTFHE(ps, seed=3)
    for p in (3, 7, 8):
        rng = np.random.default_rng(p)
        low = [int(x) for x in rng.integers(0, 2, p)]
        cases = [(low, 1), (low + [1 - x for x in low], 1), ([0] + low[1:] + [0], 0), ([1] + low[1:] + [1], 2), (low[:max(1, p - 2)], 1)]
        tabs = np.zeros((len(cases), 2 * p), np.uint8)
        for i, (t, _) in enumerate(cases):
            tabs[i, :len(t)] = t
        for m in range(2 * p):
            ct = ref.encrypt(p, np.array([m], np.int32), np.array([m]), 9)[0]
            outs, _ = ref.pbs_multi(p, ct, tabs, np.array([len(t) for t, _ in cases], np.uint8), np.array([md for _, md in cases], np.int32))
            dec = ref.decrypt(p, outs)
            for i, (t, _) in enumerate(cases):
                if m < len(t):
                    assert dec[i] == t[m], (p, m, i)
    e = next(x for x in load_ref_mapped() if x["circuit"] == "_2_input_gates" and x["p"] == 11 and x["mapper"] == "search")
    env = read_lbf(e["lbf"])
    prog = levelize(env, 11, multi_value=True)
    assert (prog.n_boots, prog.n_groups, prog.n_rotations) == (10, 2, 2)
    inputs = selfcheck_inputs(e["input_names"])
    bits = np.array([inputs[nm][:4] for nm in prog.input_names], dtype=np.uint8)
    got = ref.eval_prog(prog, bits, enc_seed=1, multi_value=True)
    want = cleartext.lut_eval(env, {nm: bits[i] for i, nm in enumerate(prog.input_names)})
    for nm in prog.output_names:
        assert np.array_equal(got[prog.out_index[nm]], np.asarray(want[nm])), nm
