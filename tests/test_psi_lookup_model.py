"""CPU check of the index arithmetic behind the three-instruction psi-table look-up of the blind-rotation kernels
(tfhe_fbs_map_b200/csrc/kernels.cuh, k_blind_rotate2 `factor`).

The kernels store psi^x - 1 at table slot psw(x) = x ^ ((x >> 5) & 15) and read, for spectrum element e of thread position tau and
monomial exponent E, the byte offset (PK * (1 + (brev3(e) << 12))) & 0x7FF8 with PK = 8 * psw(x0) | (E & 7), x0 = E * odd0 mod 4096.
This must be the slot of the evaluation point's exponent E * (2 * brev11(8 tau + e) + 1) mod 4096 for every (tau, e, E)."""
import numpy as np

LOGN, N = 11, 2048


def brev(v, bits):
    r = 0
    for i in range(bits):
        r |= ((v >> i) & 1) << (bits - 1 - i)
    return r


def psw(x):
    return x ^ ((x >> 5) & 15)


def test_psw_is_a_bijection_that_keeps_the_element_bits():
    xs = np.arange(2 * N)
    assert sorted(psw(xs).tolist()) == xs.tolist()
    assert np.array_equal(psw(xs) >> 9, xs >> 9)            # bits 9..11 (the element field) are untouched by the fold
    assert np.array_equal(psw(xs) >> 4, xs >> 4)            # only the low nibble moves


def test_packed_word_multiply_reaches_every_elements_slot():
    br3 = [brev(e, 3) for e in range(8)]
    rng = np.random.default_rng(7)
    for tau in list(range(0, 256, 17)) + [255]:
        odd0 = 2 * brev(tau, LOGN - 3) + 1
        for E in rng.integers(0, 2 * N, 64).tolist() + [0, 1, 7, 8, 2 * N - 1, 3 * N]:
            x0 = (E * odd0) % (2 * N)
            pk = (8 * psw(x0)) | (E & 7)
            for e in range(8):
                off = ((pk * (1 + (br3[e] << (LOGN + 1)))) & 0xFFFFFFFF) & ((2 * N - 1) << 3)
                point = (E * (2 * brev(8 * tau + e, LOGN) + 1)) % (2 * N)      # exponent of psi at NTT output 8 tau + e
                assert off == 8 * psw(point), (tau, E, e)


def test_half_warp_conflict_model_matches_the_documented_numbers():
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("psi_hash_model", os.path.join(os.path.dirname(__file__), "..", "tools", "psi_hash_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert abs(m.cost([1, 2, 4, 8, 0, 0, 0]) - 2.3359375) < 1e-9        # the fold the kernels use
    assert abs(m.cost([2, 4, 8, 1, 2, 4, 8]) - 2.37890625) < 1e-9       # round-1 fold
