"""Cluster-split blind rotation (k_blind_rotate_cl: one bootstrap over a thread-block cluster of 2 / 4 / 8 CTAs, exchanges
through distributed shared memory) against the CPU oracle at every ciphertext tap, and against the one-CTA kernels at full
size.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

from oracle.tfhe_ref import RefTFHE
from tfhe_fbs_map_b200 import params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["toy5v", "toy5u"])          # N = 2048, n = 10, three / two key bits per step
@pytest.mark.parametrize("mode", [2, 4, 8, 12, 14, 18])          # 1x: prime-split twin (two threads per ring element)
def test_cluster_split_bit_exact_against_cpu_oracle(name, mode):
    from tfhe_fbs_map_b200.backend import B200Backend
    be = B200Backend(name, device=0, seed=4242)
    ref = RefTFHE(params.get(name), seed=4242)
    try:
        be.set_cluster(mode)
        for p in (5, 7):
            rng = np.random.default_rng(10 * p + mode)
            low = [int(x) for x in rng.integers(0, 2, p)]
            cases = [(low, 1), (low + [1 - x for x in low], 1), ([0] + low[1:] + [0], 0), ([1] + low[1:] + [1], 2)]
            msgs, rows, lens, modes = [], [], [], []
            for tab, md in cases:
                for m in range(len(tab)):
                    row = np.zeros(2 * p, np.uint8); row[:len(tab)] = tab
                    msgs.append(m); rows.append(row); lens.append(len(tab)); modes.append(md)
            cts = ref.encrypt(p, np.array(msgs, np.int32), np.arange(len(msgs)), 3)
            out, ks, ms, acc = be.debug_pbs(p, cts, np.array(rows), np.array(lens, np.uint8), np.array(modes, np.int32))
            for i in range(len(msgs)):
                ro, rks, rms, racc = ref.pbs(p, cts[i], rows[i][:lens[i]], modes[i])
                assert np.array_equal(ks[i], rks) and np.array_equal(ms[i], rms), f"{name} C={mode} p={p} #{i}: key switch"
                assert np.array_equal(acc[i], racc), f"{name} C={mode} p={p} #{i}: accumulator"
                assert np.array_equal(out[i], ro), f"{name} C={mode} p={p} #{i}: extracted ciphertext"
            want = [rows[i][msgs[i]] for i in range(len(msgs))]
            assert be.debug_decrypt(p, out).tolist() == [int(x) for x in want]
    finally:
        be.close()


@pytest.mark.parametrize("pset", ["A3", "A2"])
def test_cluster_split_matches_one_cta_kernels_full_size(pset):
    """Full size (n = 742): auto mode (cluster of 4 for 20 jobs), and every forced cluster size, give the ciphertexts of
    the one-CTA kernels bit for bit; three repeats are identical (no race in the exchanges or the key ring)."""
    from tfhe_fbs_map_b200.backend import B200Backend
    be = B200Backend(pset, device=0, seed=99)
    try:
        p, count = 17, 20
        rng = np.random.default_rng(3)
        msgs = rng.integers(0, 2 * p, count).astype(np.int32)
        low = rng.integers(0, 2, (count, p)).astype(np.uint8)
        tables = np.concatenate([low, 1 - low], axis=1)
        cts = be.debug_encrypt(p, msgs, np.arange(count, dtype=np.uint64), enc_seed=6)
        lens, modes = np.full(count, 2 * p, np.uint8), np.ones(count, np.int32)
        be.set_cluster(1)
        base = be.debug_pbs(p, cts, tables, lens, modes)
        assert np.array_equal(be.debug_decrypt(p, base[0]), tables[np.arange(count), msgs])
        for mode in (0, 2, 4, 8, 12, 14, 18, 14, 0):
            be.set_cluster(mode)
            got = be.debug_pbs(p, cts, tables, lens, modes)
            for a, b, what in zip(base, got, ("out", "ks", "ms", "acc")):
                assert np.array_equal(a, b), f"{pset} cluster mode {mode}: {what} differs from the one-CTA kernel"
        # a full wave of paired CTAs plus a 30-job tail: in auto mode the tail runs cluster-split (4 CTAs per bootstrap)
        count = 2 * be.info()["sm_count"] + 30
        msgs = rng.integers(0, 2 * p, count).astype(np.int32)
        low = rng.integers(0, 2, (count, p)).astype(np.uint8)
        tables = np.concatenate([low, 1 - low], axis=1)
        cts = be.debug_encrypt(p, msgs, np.arange(count, dtype=np.uint64), enc_seed=7)
        lens, modes = np.full(count, 2 * p, np.uint8), np.ones(count, np.int32)
        be.set_cluster(1)
        base = be.debug_pbs(p, cts, tables, lens, modes)
        be.set_cluster(0)
        got = be.debug_pbs(p, cts, tables, lens, modes)
        for a, b, what in zip(base, got, ("out", "ks", "ms", "acc")):
            assert np.array_equal(a, b), f"{pset} cluster-split tail: {what} differs"
        assert np.array_equal(be.debug_decrypt(p, got[0]), tables[np.arange(count), msgs])
    finally:
        be.close()
