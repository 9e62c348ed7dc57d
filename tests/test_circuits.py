"""Synthetic source circuits (tfhe_fbs_map_b200/circuits.py): arithmetic identities, AES S-box table from first
principles, FIPS-197 known answer.  Cleartext oracle only (CPU)."""
import gzip
import os

import numpy as np
import pytest

from conftest import GOLD
from oracle import cleartext
from tfhe_fbs_map_b200 import circuits, levelize
from tfhe_fbs_map_b200.formats import read_lbf


def _gmul(a, b):
    r = 0
    for _ in range(8):
        if b & 1:
            r ^= a
        hi = a & 0x80
        a = (a << 1) & 0xFF
        if hi:
            a ^= 0x1B
        b >>= 1
    return r


def _sbox(a):
    x = 0
    if a:
        x = 1
        for _ in range(254):
            x = _gmul(x, a)
    y = 0
    for i in range(8):
        bit = ((x >> i) ^ (x >> ((i + 4) % 8)) ^ (x >> ((i + 5) % 8)) ^ (x >> ((i + 6) % 8)) ^ (x >> ((i + 7) % 8)) ^ (0x63 >> i)) & 1
        y |= bit << i
    return y


def aes_inputs(keys, pts):
    inp = {}
    for i in range(128):
        inp[f"k{i}"] = [(k[i // 8] >> (7 - i % 8)) & 1 for k in keys]
        inp[f"p{i}"] = [(p[i // 8] >> (7 - i % 8)) & 1 for p in pts]
    return inp


def aes_outputs(out, count):
    return [bytes(sum(int(np.asarray(out[f"c{8 * i + j}"])[b]) << (7 - j) for j in range(8)) for i in range(16)) for b in range(count)]


def test_sbox_matches_gf256_inverse_plus_affine():
    env = circuits.aes_sbox_full()
    vals = np.arange(256)
    out = cleartext.bit_eval(env, {f"x{i}": (vals >> (7 - i)) & 1 for i in range(8)})
    got = sum(np.asarray(out[f"s{i}"]).astype(int) << (7 - i) for i in range(8))
    assert got.tolist() == [_sbox(int(v)) for v in vals]
    st = env.stats()
    assert st["nb_and"] == 32                      # Boyar-Peralta: 32 AND gates


def test_aes128_fips197_known_answer():
    env = circuits.aes128()
    st = env.stats()
    assert st["nb_and"] == 6400 and st["nb_inp"] == 256 and st["nb_out"] == 128      # same AND count as Bristol aes_128
    key, pt = bytes(range(16)), bytes.fromhex("00112233445566778899aabbccddeeff")
    out = cleartext.bit_eval(env, aes_inputs([key, bytes(16)], [pt, bytes(16)]))
    cts = aes_outputs(out, 2)
    assert cts[0].hex() == "69c4e0d86a7b0430d8cdb78070b4c55a"               # FIPS-197 appendix C.1
    assert cts[1].hex() == "66e94bd4ef8a2c3b884cfa59ca342b2e"               # AES-128(0^128, 0^128)


@pytest.mark.parametrize("n", [1, 5, 16])
def test_adder_and_multiplier(n):
    rng = np.random.default_rng(n)
    B = 64
    A, Bv = rng.integers(0, 2 ** n, B), rng.integers(0, 2 ** n, B)
    inp = {f"a{i}": (A >> i) & 1 for i in range(n)}
    inp.update({f"b{i}": (Bv >> i) & 1 for i in range(n)})
    out = cleartext.bit_eval(circuits.ripple_carry_adder(n), inp)
    assert np.all(sum(np.asarray(out[f"f{i}"]).astype(np.int64) << i for i in range(n + 1)) == A + Bv)
    if n > 1:
        out = cleartext.bit_eval(circuits.array_multiplier(n), inp)
        assert np.all(sum(np.asarray(out[f"f{i}"]).astype(np.int64) << i for i in range(len(out))) == A * Bv)


@pytest.mark.parametrize("fn,rounds", [("aes128_r1_p11.lbf.gz", 1), ("aes128_r10_p11.lbf.gz", 10)])
def test_mapped_aes_fixture_equals_source_circuit(fn, rounds):
    path = os.path.join(GOLD, "lbf", fn)
    if not os.path.exists(path):
        pytest.skip(f"{fn} not generated (tools/map_aes128.py)")
    lut = read_lbf(gzip.open(path, "rt").read())
    rng = np.random.default_rng(1)
    keys = [bytes(range(16))] + [rng.bytes(16) for _ in range(7)]
    pts = [bytes.fromhex("00112233445566778899aabbccddeeff")] + [rng.bytes(16) for _ in range(7)]
    inp = aes_inputs(keys, pts)
    want = cleartext.bit_eval(circuits.aes128(rounds=rounds), inp)
    got = cleartext.lut_eval(lut, inp)
    for k in want:
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), k
    if rounds == 10:
        assert aes_outputs(got, 1)[0].hex() == "69c4e0d86a7b0430d8cdb78070b4c55a"
    prog = levelize(lut, 11)
    assert prog.n_boots == lut.stats()["nb_bootstrap"]
