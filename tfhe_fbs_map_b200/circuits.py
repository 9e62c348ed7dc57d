"""Synthetic source circuits for the benchmark configurations whose input files are not available offline.

BASELINE.json names EPFL ``adder.blif`` / ``multiplier.blif`` and Bristol ``aes_128.txt``; the reference fetches
them with git/wget (reference experiments/gen_makefile_epfl.bash:7, gen_makefile_bristol.bash:4) and there is no
network here.  These builders produce circuits of the same function and gate vocabulary (AND / inverter graphs
expressed as 2-input LUTs, as in an EPFL AIG-derived BLIF) through the ``BitExecEnv`` builder API, so a real
file can be substituted by passing its path to the CLI instead.  ``env_cls`` may be the reference's BitExecEnv.
"""
from __future__ import annotations

from .bit_env import BitExecEnv

# 2-input AND with optional input inversions, as truth tables indexed by 2*a+b (first input is the MSB)
_AND = {(0, 0): [0, 0, 0, 1], (1, 0): [0, 1, 0, 0], (0, 1): [0, 0, 1, 0], (1, 1): [1, 0, 0, 0]}


class _Aig:
    """Tiny AIG helper: literals are (node, inverted)."""

    def __init__(self, env):
        self.env = env

    def land(self, x, y):
        (a, ia), (b, ib) = x, y
        return (self.env.op_lut([a, b], list(_AND[(ia, ib)])), 0)

    def lnot(self, x):
        return (x[0], 1 - x[1])

    def lor(self, x, y):
        return self.lnot(self.land(self.lnot(x), self.lnot(y)))

    def lxor(self, x, y):
        # x^y = ~(~(x & ~y) & ~(~x & y))
        return self.lor(self.land(x, self.lnot(y)), self.land(self.lnot(x), y))

    def materialise(self, x):
        node, inv = x
        return self.env.op_lut([node], [1, 0]) if inv else node


def ripple_carry_adder(nbits=128, env_cls=BitExecEnv):
    """nbits + nbits -> nbits+1 adder as an AND-inverter graph (2*nbits inputs, nbits+1 outputs)."""
    env = env_cls()
    g = _Aig(env)
    a = [(env.input(f"a{i}"), 0) for i in range(nbits)]
    b = [(env.input(f"b{i}"), 0) for i in range(nbits)]
    carry = None
    for i in range(nbits):
        axb = g.lxor(a[i], b[i])
        if carry is None:
            s, carry = axb, g.land(a[i], b[i])
        else:
            s = g.lxor(axb, carry)
            carry = g.lor(g.land(a[i], b[i]), g.land(axb, carry))
        env.output(f"f{i}", g.materialise(s))
    env.output(f"f{nbits}", g.materialise(carry))
    return env


def array_multiplier(nbits=16, env_cls=BitExecEnv):
    """nbits x nbits -> 2*nbits array multiplier (AND partial products, ripple rows of AIG full adders)."""
    env = env_cls()
    g = _Aig(env)
    a = [(env.input(f"a{i}"), 0) for i in range(nbits)]
    b = [(env.input(f"b{i}"), 0) for i in range(nbits)]

    def full_add(x, y, c):
        xy = g.lxor(x, y)
        return g.lxor(xy, c), g.lor(g.land(x, y), g.land(xy, c))

    def half_add(x, y):
        return g.lxor(x, y), g.land(x, y)

    row = [g.land(a[j], b[0]) for j in range(nbits)]          # weights 0..n-1
    outs = [row[0]]
    acc = row[1:]                                             # weights 1..n-1 relative to next row's 0
    top = None
    for i in range(1, nbits):
        pp = [g.land(a[j], b[i]) for j in range(nbits)]
        new, carry = [], None
        for j in range(nbits):
            x = pp[j]
            y = acc[j] if j < len(acc) else top
            if y is None:
                s, c = (x, None) if carry is None else half_add(x, carry)
            elif carry is None:
                s, c = half_add(x, y)
            else:
                s, c = full_add(x, y, carry)
            new.append(s)
            carry = c
        outs.append(new[0])
        acc, top = new[1:], carry
    final = acc + ([top] if top is not None else [])
    outs.extend(final)
    for i, o in enumerate(outs):
        env.output(f"f{i}", g.materialise(o))
    return env


# ---------------------------------------------------------------------------------------------------------------------
# AES-128 (BASELINE.json configs[2]: Bristol aes_128.txt is not available offline).
# S-box: the public Boyar-Peralta 113-gate circuit (32 AND, 81 XOR/XNOR; "A small depth-16 circuit for the AES S-box",
# 2012), whose non-linear middle part is also the reference's `aes_sbox` benchmark (experiments/generate_benchmarks.py:87).
# Bit 0 of a byte is the MOST significant bit, as in that paper.  tests/test_circuits.py checks the S-box against the
# table derived from GF(2^8) inversion + affine map, and the full cipher against the FIPS-197 appendix C.1 vector.
# ---------------------------------------------------------------------------------------------------------------------
def _sbox_bits(env, x):
    """x: list of 8 nodes (x[0] = MSB) -> list of 8 nodes (MSB first)."""
    X, A = env.op_xor, env.op_and

    def XN(a, b):
        return env.op_not(env.op_xor(a, b))
    x0, x1, x2, x3, x4, x5, x6, x7 = x
    # top linear layer
    y14 = X(x3, x5); y13 = X(x0, x6); y9 = X(x0, x3); y8 = X(x0, x5); t0 = X(x1, x2); y1 = X(t0, x7); y4 = X(y1, x3)
    y12 = X(y13, y14); y2 = X(y1, x0); y5 = X(y1, x6); y3 = X(y5, y8); t1 = X(x4, y12); y15 = X(t1, x5); y20 = X(t1, x1)
    y6 = X(y15, x7); y10 = X(y15, t0); y11 = X(y20, y9); y7 = X(x7, y11); y17 = X(y10, y11); y19 = X(y10, y8)
    y16 = X(t0, y11); y21 = X(y13, y16); y18 = X(x0, y16)
    # shared non-linear part
    t2 = A(y12, y15); t3 = A(y3, y6); t4 = X(t3, t2); t5 = A(y4, x7); t6 = X(t5, t2); t7 = A(y13, y16); t8 = A(y5, y1)
    t9 = X(t8, t7); t10 = A(y2, y7); t11 = X(t10, t7); t12 = A(y9, y11); t13 = A(y14, y17); t14 = X(t13, t12)
    t15 = A(y8, y10); t16 = X(t15, t12); t17 = X(t4, t14); t18 = X(t6, t16); t19 = X(t9, t14); t20 = X(t11, t16)
    t21 = X(t17, y20); t22 = X(t18, y19); t23 = X(t19, y21); t24 = X(t20, y18); t25 = X(t21, t22); t26 = A(t21, t23)
    t27 = X(t24, t26); t28 = A(t25, t27); t29 = X(t28, t22); t30 = X(t23, t24); t31 = X(t22, t26); t32 = A(t31, t30)
    t33 = X(t32, t24); t34 = X(t23, t33); t35 = X(t27, t33); t36 = A(t24, t35); t37 = X(t36, t34); t38 = X(t27, t36)
    t39 = A(t29, t38); t40 = X(t25, t39); t41 = X(t40, t37); t42 = X(t29, t33); t43 = X(t29, t40); t44 = X(t33, t37)
    t45 = X(t42, t41)
    z0 = A(t44, y15); z1 = A(t37, y6); z2 = A(t33, x7); z3 = A(t43, y16); z4 = A(t40, y1); z5 = A(t29, y7)
    z6 = A(t42, y11); z7 = A(t45, y17); z8 = A(t41, y10); z9 = A(t44, y12); z10 = A(t37, y3); z11 = A(t33, y4)
    z12 = A(t43, y13); z13 = A(t40, y5); z14 = A(t29, y2); z15 = A(t42, y9); z16 = A(t45, y14); z17 = A(t41, y8)
    # bottom linear layer
    t46 = X(z15, z16); t47 = X(z10, z11); t48 = X(z5, z13); t49 = X(z9, z10); t50 = X(z2, z12); t51 = X(z2, z5)
    t52 = X(z7, z8); t53 = X(z0, z3); t54 = X(z6, z7); t55 = X(z16, z17); t56 = X(z12, t48); t57 = X(t50, t53)
    t58 = X(z4, t46); t59 = X(z3, t54); t60 = X(t46, t57); t61 = X(z14, t57); t62 = X(t52, t58); t63 = X(t49, t58)
    t64 = X(z4, t59); t65 = X(t61, t62); t66 = X(z1, t63); s0 = X(t59, t63); s6 = XN(t56, t62); s7 = XN(t48, t60)
    t67 = X(t64, t65); s3 = X(t53, t66); s4 = X(t51, t66); s5 = X(t47, t65); s1 = XN(t64, s3); s2 = XN(t55, t67)
    return [s0, s1, s2, s3, s4, s5, s6, s7]


def aes_sbox_full(env_cls=BitExecEnv):
    """Stand-alone 8 -> 8 S-box circuit (inputs x0..x7, outputs s0..s7, MSB first)."""
    env = env_cls()
    x = [env.input(f"x{i}") for i in range(8)]
    for i, s in enumerate(_sbox_bits(env, x)):
        env.output(f"s{i}", s)
    return env


def _xtime(env, b):
    """Multiply a byte (MSB-first node list) by x in GF(2^8) mod x^8+x^4+x^3+x+1."""
    m = b[0]
    out = b[1:] + [m]                 # shift left, bit0 (LSB) = carry
    # reduction polynomial 0x1B: bits 4, 3, 1, 0 (LSB numbering) -> MSB-first positions 3, 4, 6, 7
    out[3] = env.op_xor(out[3], m)
    out[4] = env.op_xor(out[4], m)
    out[6] = env.op_xor(out[6], m)
    return out


def _xor_bytes(env, a, b):
    return [env.op_xor(p, q) for p, q in zip(a, b)]


def aes128(env_cls=BitExecEnv, rounds=10):
    """AES-128 encryption: inputs k0..k127 (key) then p0..p127 (plaintext), outputs c0..c127; bit i of the 128-bit block
    is bit (7 - i%8) of byte i//8 (i.e. MSB-first within each byte, bytes in FIPS-197 order)."""
    env = env_cls()
    key = [env.input(f"k{i}") for i in range(128)]
    pt = [env.input(f"p{i}") for i in range(128)]
    kb = [key[8 * i:8 * i + 8] for i in range(16)]          # round-key bytes
    st = [_xor_bytes(env, pt[8 * i:8 * i + 8], kb[i]) for i in range(16)]
    rcon = [0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36]
    for rnd in range(1, rounds + 1):
        # key schedule: w[i] words are kb[4i..4i+3]
        tmp = [kb[13], kb[14], kb[15], kb[12]]                # RotWord of the last word
        tmp = [_sbox_bits(env, t) for t in tmp]
        rc = rcon[rnd - 1]
        tmp[0] = [env.op_not(b) if (rc >> (7 - i)) & 1 else b for i, b in enumerate(tmp[0])]
        nk = []
        for w in range(4):
            for j in range(4):
                src = tmp[j] if w == 0 else nk[4 * (w - 1) + j]
                nk.append(_xor_bytes(env, kb[4 * w + j], src))
        kb = nk
        # SubBytes + ShiftRows (state byte index = 4*col + row)
        sb = [_sbox_bits(env, b) for b in st]
        sr = [sb[4 * ((c + r) % 4) + r] for c in range(4) for r in range(4)]
        if rnd < rounds or rounds < 10:
            mc = []
            for c in range(4):
                a = sr[4 * c:4 * c + 4]
                xa = [_xtime(env, b) for b in a]
                for r in range(4):
                    # 2*a[r] + 3*a[r+1] + a[r+2] + a[r+3]
                    t = _xor_bytes(env, xa[r], xa[(r + 1) % 4])
                    t = _xor_bytes(env, t, a[(r + 1) % 4])
                    t = _xor_bytes(env, t, a[(r + 2) % 4])
                    t = _xor_bytes(env, t, a[(r + 3) % 4])
                    mc.append(t)
            sr = mc
        st = [_xor_bytes(env, sr[i], kb[i]) for i in range(16)]
    for i in range(16):
        for j in range(8):
            env.output(f"c{8 * i + j}", st[i][j])
    return env
