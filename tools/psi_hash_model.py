"""Exact bank-conflict model of the psi-power table look-ups of the blind-rotation kernels.

Thread position tau of a polynomial holds the NTT outputs 8*tau + e (e < 8), i.e. the evaluations at psi^(odd0 + (brev3(e) << 9)),
odd0 = 2*brev8(tau) + 1 (N = 2048).  For a monomial X^E the table index is x = E * (odd0 + (brev3(e) << 9)) mod 4096.  The 16 lanes of
a half-warp (one 64-bit shared-memory transaction) differ in tau bits 0..3 = odd0 bits 8..5, so their indices are
x0 + 32*E*j (j < 16): only index bits 5..11 vary, as y_j = (y0 + E*j) mod 128 with y = x >> 5.  The table is stored at
idx' = x ^ fold(x), fold = XOR of one 4-bit column per set index bit above bit 3; the bank pair of an 8-byte entry is idx' mod 16.

cost(cols) = mean over all (y0, E mod 128) of the largest number of DISTINCT addresses that fall into one bank pair
           = wavefronts per half-warp request (1.0 = conflict-free).

Results (printed by this script):
  round-1 fold  nibble1 ^ nibble2                      2.379
  fold of index bits 5..8 only (any bijection)         2.336   <- used: keeps the element bits 9..11 an additive offset field
  best GF(2)-linear fold found (hill climbing)         ~1.93   (needs bits 9..11 in the fold)
usage: python tools/psi_hash_model.py [--search]"""
import random
import sys

import numpy as np

J = np.arange(16)
Y0, E = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")
YS = np.sort(((Y0[..., None] + E[..., None] * J) % 128).reshape(-1, 16), axis=1)      # [16384, 16]


def cost(cols):
    """cols[k] = 4-bit column XORed into the low index nibble when index bit 5 + k is set (k < 7)."""
    tab = np.zeros(128, dtype=np.int64)
    for k in range(7):
        tab ^= ((np.arange(128) >> k) & 1) * cols[k]
    key = np.sort(tab[YS] * 128 + YS, axis=1)
    first = np.ones(key.shape, dtype=bool)
    first[:, 1:] = key[:, 1:] != key[:, :-1]
    onehot = ((key >> 7)[:, :, None] == np.arange(16)[None, None, :]) & first[:, :, None]
    return float(onehot.sum(1).max(1).mean())


def main():
    print("round-1 fold (both upper nibbles)      ", cost([2, 4, 8, 1, 2, 4, 8]))
    print("bits 5..8 only: (x >> 5) & 15  [used]  ", cost([1, 2, 4, 8, 0, 0, 0]))
    print("bits 5..8 only, another bijection      ", cost([1, 4, 2, 11, 0, 0, 0]))
    print("a good unrestricted linear fold        ", cost([0, 2, 7, 3, 15, 0, 0]))
    if "--search" in sys.argv:
        random.seed(1)
        best = (9.0, None)
        for restart in range(6):
            cols = [random.randrange(16) for _ in range(7)]
            c = cost(cols)
            improved = True
            while improved:
                improved = False
                for k in range(7):
                    for v in range(16):
                        if v == cols[k]:
                            continue
                        old = cols[k]
                        cols[k] = v
                        c2 = cost(cols)
                        if c2 < c - 1e-9:
                            c, improved = c2, True
                        else:
                            cols[k] = old
            if c < best[0]:
                best = (c, list(cols))
                print("restart", restart, best, flush=True)


if __name__ == "__main__":
    main()
