"""Search a GF(2)-linear XOR swizzle of the low 4 index bits that makes every NTT transpose layout
bank-conflict free for 64-bit shared-memory accesses (a half-warp = 16 lanes must hit 16 distinct
8-byte bank pairs).  Layout lb: idx(tau,e) = ((tau>>lb)<<(lb+3)) | (e<<lb) | (tau & ((1<<lb)-1)).
Output: per LOGN the 4-bit column vectors for index bits 4..7 (bits 0..3 map to themselves)."""
import itertools, sys

def layouts(logn):
    npass = (logn + 2) // 3
    fwd = [max(logn - 1 - 3 * p - 2, 0) for p in range(npass)]
    inv = [min(3 * p, logn - 3) for p in range(npass)]
    return sorted(set(fwd + inv))

def idx(tau, e, lb):
    return ((tau >> lb) << (lb + 3)) | (e << lb) | (tau & ((1 << lb) - 1))

def sw(i, cols):
    x = i
    for b, c in enumerate(cols):
        if (i >> (4 + b)) & 1:
            x ^= c
    return x

def ok(logn, cols):
    T = (1 << logn) // 8
    for lb in layouts(logn):
        for e in range(8):
            for h in range(0, T, 16):
                s = {sw(idx(t, e, lb), cols) & 15 for t in range(h, min(h + 16, T))}
                if len(s) != min(16, T - h):
                    return False
    return True

for logn in range(8, 13):
    found = None
    for cols in itertools.product(range(16), repeat=4):
        if ok(logn, cols):
            found = cols; break
    print(logn, layouts(logn), found)
