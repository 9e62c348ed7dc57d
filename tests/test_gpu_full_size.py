"""Full-size (parameter set A, 128-bit-secure shape) checks on the B200: decrypted results against the reference's
cleartext semantics (golden outputs of the reference's own eval), plus size-independent properties."""
import numpy as np
import pytest

from conftest import load_lbf_index, out_hash, read_golden_lbf, selfcheck_inputs
from oracle import cleartext
from oracle.tfhe_ref import RefTFHE
from tfhe_fbs_map_b200 import levelize, params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["A", "A2", "A3"])     # A2 / A3: same shape, two / three key bits per blind-rotation step
def be(request):
    from tfhe_fbs_map_b200.backend import B200Backend
    b = B200Backend(request.param, device=0, seed=20241018)
    yield b
    b.close()


@pytest.mark.parametrize("p", [3, 5, 7, 9, 11, 13, 15, 17])
def test_pbs_sweep_exhaustive_tables(be, p):
    """BASELINE config 5 correctness leg: every index of random tables in all three modes, 0 failures expected
    (p_fail(A, p=17, norm2=1) ~ 1e-11 per PBS)."""
    rng = np.random.default_rng(p)
    rows, msgs, lens, modes, want = [], [], [], [], []
    for rep in range(6):
        low = [int(x) for x in rng.integers(0, 2, p)]
        for tab, mode in ((low, 1), (low + [1 - x for x in low], 1), ([0] + low[1:] + [0], 0), ([1] + low[1:] + [1], 2)):
            for m in range(len(tab)):
                row = np.zeros(2 * p, np.uint8); row[:len(tab)] = tab
                rows.append(row); msgs.append(m); lens.append(len(tab)); modes.append(mode); want.append(tab[m])
    got = be.pbs_batch(p, np.array(msgs), np.array(rows), np.array(lens), np.array(modes))
    assert np.array_equal(got, np.array(want, np.uint8)), f"{int((got != np.array(want)).sum())} PBS failures of {len(want)}"


def test_one_pbs_bit_exact_against_cpu_oracle(be):
    ref = RefTFHE(be.params, seed=20241018)
    p = 17
    tab = [0, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 0, 1, 0, 1, 1, 0]
    tab2 = tab + [1 - x for x in tab]
    cts = ref.encrypt(p, np.array([3, 20], np.int32), np.array([1, 2]), 5)
    tables = np.array([tab2, tab2], np.uint8)
    out, ks, ms, acc = be.debug_pbs(p, cts, tables, np.array([34, 34], np.uint8), np.array([1, 1], np.int32))
    for i in range(2):
        ro, rks, rms, racc = ref.pbs(p, cts[i], tab2, 1)
        assert np.array_equal(ks[i], rks) and np.array_equal(ms[i], rms)
        assert np.array_equal(acc[i], racc) and np.array_equal(out[i], ro)


@pytest.mark.parametrize("fn", ["adder8_p15.lbf", "aes_sbox_p11.lbf", "aes_sbox_p15.lbf", "ascon_lut_p17.lbf", "mult8_p17.lbf"])
def test_circuits_under_real_parameters(be, fn):
    item = next(x for x in load_lbf_index() if x["file"] == fn)
    env = read_golden_lbf(fn)
    B = 64
    inputs = {k: v[:B] for k, v in selfcheck_inputs(item["input_names"]).items()}
    want = cleartext.lut_eval(env, inputs)
    got = env.eval(inputs, fbs_size=item["p"], backend=be)
    for k in want:
        assert np.array_equal(got[k], np.asarray(want[k])), k


def test_adder128_round_trip_property(be):
    """Size-independent property at BASELINE configs[1]'s circuit: a + b computed under encryption equals integer
    addition; adding 0 is the identity; the carry chain is exercised by all-ones + 1."""
    env = read_golden_lbf("adder128_p15.lbf")
    rng = np.random.default_rng(5)
    A = [int.from_bytes(rng.bytes(16), "little") for _ in range(13)] + [2 ** 128 - 1, 0, 2 ** 127]
    Bv = [int.from_bytes(rng.bytes(16), "little") for _ in range(13)] + [1, 0, 2 ** 127]
    iv = {f"a{i}": [(x >> i) & 1 for x in A] for i in range(128)}
    iv.update({f"b{i}": [(x >> i) & 1 for x in Bv] for i in range(128)})
    got = env.eval(iv, fbs_size=15, backend=be)
    sums = [sum(int(got[f"f{i}"][j]) << i for i in range(129)) for j in range(len(A))]
    assert sums == [x + y for x, y in zip(A, Bv)]


def test_blind_rotation_is_deterministic_under_load(be):
    """Race detector of last resort (compute-sanitizer is not available on the GPU pool): the same 700 bootstraps -- two full
    waves of paired CTAs plus a partial wave on the one-bootstrap kernel -- run three times must give bit-identical
    ciphertexts at every tap; a missing barrier in the transposes or the key ring shows up as a difference."""
    p, count = 11, 700
    rng = np.random.default_rng(9)
    msgs = rng.integers(0, 2 * p, count).astype(np.int32)
    low = rng.integers(0, 2, (count, p)).astype(np.uint8)
    tables = np.concatenate([low, 1 - low], axis=1)
    cts = be.debug_encrypt(p, msgs, np.arange(count, dtype=np.uint64) + 17, enc_seed=4)
    runs = [be.debug_pbs(p, cts, tables, np.full(count, 2 * p, np.uint8), np.ones(count, np.int32)) for _ in range(3)]
    for r in runs[1:]:
        for a, b in zip(runs[0], r):
            assert np.array_equal(a, b)
    assert np.array_equal(be.debug_decrypt(p, runs[0][0]), tables[np.arange(count), msgs])


def test_640_pbs_bit_exact_against_cpu_oracle(be):
    """Default set A3 at full size, batched: 640 bootstraps = two full waves of the two-bootstraps-per-CTA kernel
    (k_blind_rotate2<11,1,2,2,3>) plus a 48-job tail on the one-bootstrap-per-CTA instantiation (<11,1,1,1,3>); every
    accumulator and every extracted ciphertext must equal the C oracle's bit for bit."""
    if be.params.name != "A3":
        pytest.skip("batched full-size ciphertext parity runs on the default set only (CPU oracle time)")
    ref = RefTFHE(be.params, seed=20241018)
    p, count = 15, 640
    rng = np.random.default_rng(21)
    msgs = rng.integers(0, 2 * p, count).astype(np.int32)
    low = rng.integers(0, 2, (count, p)).astype(np.uint8)
    tables = np.concatenate([low, 1 - low], axis=1)
    cts = be.debug_encrypt(p, msgs, np.arange(count, dtype=np.uint64) + 5, enc_seed=8)
    assert np.array_equal(cts, ref.encrypt(p, msgs, np.arange(count, dtype=np.uint64) + 5, 8))
    lens, modes = np.full(count, 2 * p, np.uint8), np.ones(count, np.int32)
    be.set_cluster(1)                                  # this test is about the two one-CTA instantiations
    try:
        out, ks, ms, acc = be.debug_pbs(p, cts, tables, lens, modes)
    finally:
        be.set_cluster(0)
    rout, racc = ref.pbs_batch(p, cts, tables, lens, modes)
    bad = [i for i in range(count) if not (np.array_equal(acc[i], racc[i]) and np.array_equal(out[i], rout[i]))]
    assert not bad, f"{len(bad)} of {count} bootstraps differ from the CPU oracle, first {bad[:5]}"
    assert np.array_equal(be.debug_decrypt(p, out), tables[np.arange(count), msgs])
