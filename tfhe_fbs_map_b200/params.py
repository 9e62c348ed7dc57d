"""TFHE parameter sets and the noise / failure-probability estimator.

The reference never runs TFHE; it asks a patched ``concrete-optimizer`` for a cost per
``(fbs_size, norm2_linprod)`` (reference experiments/add_exec_estimates.py:9-16).  The patch changes the
noise bound to the *absolute* number of message values p (reference experiments/concrete.patch:21-27,
``2^(log q - 2) / p``: one negacyclic padding bit, decision half-interval q/(4p)) and scales the input
noise by ``sqrt(sq_norm2)`` (concrete.patch:133-134).  ``estimate()`` below applies exactly that bound to the
parameter sets this executor ships, so every benchmark line can state its PBS failure probability.

Ciphertext modulus is q = p1*p2, two 30-bit NTT primes; all standard deviations are
relative to the torus (i.e. in units of Q).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, asdict

FBS_P1, FBS_P2 = 0x3FFE8001, 0x3FFF4001
FBS_Q = FBS_P1 * FBS_P2             # ciphertext modulus: product of two 30-bit NTT primes (60 bits)
GOLDILOCKS_P = FBS_Q   # old name kept for the tests' imports


@dataclass(frozen=True)
class ParamSet:
    name: str
    n: int            # small LWE dimension
    k: int            # GLWE dimension
    N: int            # polynomial size
    bsk_l: int        # blind-rotate decomposition levels
    bsk_beta: int     # log2 of blind-rotate base
    ks_l: int         # key-switch levels
    ks_beta: int      # log2 of key-switch base
    lwe_sigma: float  # std of small-LWE (KSK) noise, torus units
    glwe_sigma: float # std of GLWE (BSK, fresh input) noise, torus units
    secure: bool = True
    bsk_unroll: int = 1  # m = 2 or 3 key bits per blind-rotation step (2^m - 1 GGSW per key group, ceil(n/m) steps; needs bsk_l = 1)

    @property
    def big_dim(self) -> int:
        return self.k * self.N

    @property
    def ct_words(self) -> int:
        return self.k * self.N + 1

    @property
    def lwe_noise_scale(self) -> int:
        return max(0, int(round(self.lwe_sigma * FBS_Q)))

    @property
    def glwe_noise_scale(self) -> int:
        return max(0, int(round(self.glwe_sigma * FBS_Q)))

    @property
    def n_ggsw(self) -> int:
        m = self.bsk_unroll if self.bsk_unroll in (2, 3) else 1      # key bits per step; n padded with zero bits to a multiple
        return self.n if m == 1 else ((1 << m) - 1) * ((self.n + m - 1) // m)

    @property
    def bsk_bytes(self) -> int:
        return self.n_ggsw * (self.k + 1) ** 2 * self.bsk_l * self.N * 8

    @property
    def ksk_bytes(self) -> int:
        return self.k * self.N * self.ks_l * (self.n + 1) * 8

    # ---- algorithmic work per PBS -------------------------------------------
    def modmul_per_pbs(self) -> int:
        """64-bit modular multiplies of one blind rotation, canonical radix-2 count."""
        k1, l, N = self.k + 1, self.bsk_l, self.N
        ntt = (k1 * l + k1) * (N // 2) * int(math.log2(N))
        if self.bsk_unroll in (2, 3): # per key group: one transform set, 2^m - 1 factor x key products + 1 digit x bundle product per key word
            m = self.bsk_unroll
            return ((self.n + m - 1) // m) * (ntt + (1 << m) * k1 * k1 * l * N)
        return self.n * (ntt + k1 * k1 * l * N)

    def mul32_per_pbs(self) -> int:
        """32x32->64 multiplies: 4 per modmul + 2 per key-switch MAC."""
        return 4 * self.modmul_per_pbs() + 2 * self.k * self.N * self.ks_l * (self.n + 1)

    # ---- noise model (binary keys, torus-normalised) -----------------------------
    def variances(self, norm2: float = 1.0) -> dict:
        B = 2.0 ** self.bsk_beta
        Bk = 2.0 ** self.ks_beta
        kN = self.k * self.N
        t_key = self.bsk_l * (self.k + 1) * self.N * (B * B + 2) / 12.0 * self.glwe_sigma ** 2     # GGSW noise through the digits
        t_round = (1 + kN / 2.0) / (24.0 * B ** (2 * self.bsk_l))                                 # rounding error times a key bit (E m^2 = 1/2)
        if self.bsk_unroll in (2, 3):
            # one external product per key group with the bundle sum_c (X^{e_c} - 1) GGSW_c: 2^m - 1 key-noise terms scaled by
            # |X^e - 1|^2 = 2, and a plaintext X^e - 1 (norm^2 2, present with probability 1 - 2^-m) instead of a bit
            # (t_round carries E m^2 = 1/2 of a bit: the factor is 2 * 2 * (1 - 2^-m))
            m = self.bsk_unroll
            v_br = (self.n / float(m)) * (2.0 * ((1 << m) - 1) * t_key + 4.0 * (1.0 - 2.0 ** -m) * t_round)
        else:
            v_br = self.n * (t_key + t_round)
        v_ks = kN * (self.ks_l * (Bk * Bk + 2) / 12.0 * self.lwe_sigma ** 2 + 1.0 / (24.0 * Bk ** (2 * self.ks_l)))
        v_ms = (1.0 / 12.0 + self.n / 24.0) / (2.0 * self.N) ** 2
        return dict(v_br=v_br, v_ks=v_ks, v_ms=v_ms, v_in=norm2 * v_br + v_ks + v_ms)

    def p_fail(self, p: int, norm2: float = 1.0, mv_norm2: float = 1.0) -> float:
        """Per-PBS failure probability for message space Z_p with the patched bound q/(4p).  ``mv_norm2``: squared norm of the
        table polynomial e_f of a multi-value bootstrap (multiplies the blind-rotation noise; at most p + 3, 1 = ordinary PBS)."""
        v = self.variances(norm2)
        v_in = norm2 * mv_norm2 * v["v_br"] + v["v_ks"] + v["v_ms"]
        return math.erfc((1.0 / (4.0 * p)) / math.sqrt(2.0 * v_in))

    def as_dict(self) -> dict:
        return asdict(self)


class CParams(ctypes.Structure):
    """Mirror of ``fbs_params`` in include/fbs_b200.h."""
    _fields_ = [("n", ctypes.c_int32), ("k", ctypes.c_int32), ("N", ctypes.c_int32),
                ("bsk_l", ctypes.c_int32), ("bsk_beta", ctypes.c_int32),
                ("ks_l", ctypes.c_int32), ("ks_beta", ctypes.c_int32), ("bsk_unroll", ctypes.c_int32),
                ("lwe_noise", ctypes.c_uint64), ("glwe_noise", ctypes.c_uint64)]


def to_c(ps: ParamSet) -> CParams:
    return CParams(ps.n, ps.k, ps.N, ps.bsk_l, ps.bsk_beta, ps.ks_l, ps.ks_beta, ps.bsk_unroll,
                   ps.lwe_noise_scale, ps.glwe_noise_scale)


# 128-bit-secure anchor: the public tfhe-rs "message 2 / carry 2, KS->PBS" shape (n=742, k=1, N=2048,
# PBS base 2^23 x 1 level, KS base 2^3 x 5 levels, sigma_lwe=2^-17.1, sigma_glwe=2^-51.6).  Re-validate with a
# lattice estimator before claiming security; this repository is a benchmark harness.
SET_A = ParamSet("A", n=742, k=1, N=2048, bsk_l=1, bsk_beta=23, ks_l=5, ks_beta=3,
                 lwe_sigma=2.0 ** -17.1, glwe_sigma=2.0 ** -51.6)
# conservative two-level variant
SET_C = ParamSet("C", n=800, k=1, N=2048, bsk_l=2, bsk_beta=15, ks_l=6, ks_beta=3,
                 lwe_sigma=2.0 ** -18.63, glwe_sigma=2.0 ** -51.6)
# small-p variant (k=2, N=1024): ~25% fewer multiplies, larger mod-switch noise -> only for p <= 8
SET_S = ParamSet("S", n=700, k=2, N=1024, bsk_l=1, bsk_beta=23, ks_l=5, ks_beta=3,
                 lwe_sigma=2.0 ** -15.99, glwe_sigma=2.0 ** -51.6)

# INSECURE toy sets: exist so the CPU oracle finishes in milliseconds and every kernel template
# (k, l, N variants) gets exercised bit-exactly.  Never use outside tests.
TOY_1 = ParamSet("toy1", n=16, k=1, N=256, bsk_l=2, bsk_beta=12, ks_l=4, ks_beta=4,
                 lwe_sigma=2.0 ** -30, glwe_sigma=2.0 ** -50, secure=False)
TOY_2 = ParamSet("toy2", n=20, k=2, N=256, bsk_l=1, bsk_beta=22, ks_l=3, ks_beta=5,
                 lwe_sigma=2.0 ** -30, glwe_sigma=2.0 ** -52, secure=False)
TOY_3 = ParamSet("toy3", n=24, k=1, N=512, bsk_l=1, bsk_beta=23, ks_l=5, ks_beta=3,
                 lwe_sigma=2.0 ** -30, glwe_sigma=2.0 ** -52, secure=False)
TOY_4 = ParamSet("toy4", n=12, k=1, N=1024, bsk_l=3, bsk_beta=8, ks_l=2, ks_beta=8,
                 lwe_sigma=2.0 ** -30, glwe_sigma=2.0 ** -50, secure=False)
TOY_5 = ParamSet("toy5", n=10, k=1, N=2048, bsk_l=1, bsk_beta=23, ks_l=5, ks_beta=3,
                 lwe_sigma=2.0 ** -30, glwe_sigma=2.0 ** -52, secure=False)

TOY_6 = ParamSet("toy6", n=8, k=1, N=2048, bsk_l=2, bsk_beta=15, ks_l=6, ks_beta=3,
                 lwe_sigma=2.0 ** -30, glwe_sigma=2.0 ** -52, secure=False)


def _unrolled(ps: ParamSet, name: str, m: int = 2) -> ParamSet:
    d = asdict(ps); d.update(name=name, bsk_unroll=m)
    return ParamSet(**d)


# key-unrolled twins (two key bits per blind-rotation step, 1.5x the bootstrapping key): same security and shape
SET_A2 = _unrolled(SET_A, "A2")
TOY_2U, TOY_3U, TOY_5U = _unrolled(TOY_2, "toy2u"), _unrolled(TOY_3, "toy3u"), _unrolled(TOY_5, "toy5u")
_d7 = asdict(TOY_3); _d7.update(name="toy7u", n=15, bsk_unroll=2)      # odd n: the last key pair is padded with a zero bit
TOY_7U = ParamSet(**_d7)
# three key bits per step (seven GGSW per key triple)
SET_A3 = _unrolled(SET_A, "A3", 3)
TOY_3V, TOY_5V = _unrolled(TOY_3, "toy3v", 3), _unrolled(TOY_5, "toy5v", 3)

PARAM_SETS = {ps.name: ps for ps in (SET_A, SET_A2, SET_C, SET_S, TOY_1, TOY_2, TOY_3, TOY_4, TOY_5, TOY_6, TOY_2U, TOY_3U, TOY_5U, TOY_7U, SET_A3, TOY_3V, TOY_5V)}
DEFAULT_SET = "A3"


def get(name: str | ParamSet) -> ParamSet:
    if isinstance(name, ParamSet):
        return name
    return PARAM_SETS[name]


def estimate(p: int, norm2: float, sets=("S", "A3", "C")) -> dict:
    """Replacement for ``optimizer --precision=p --sq-norm2=norm2`` (reference add_exec_estimates.py:14):
    first shipped set whose failure probability meets concrete's default target 4 sigma ~ 6.3e-5
    (reference concrete.patch:101-102); returns its shape, p_fail and the algorithmic cost in modmuls."""
    target = math.erfc(4.0 / math.sqrt(2.0))
    for nm in sets:
        ps = PARAM_SETS[nm]
        pf = ps.p_fail(p, norm2)
        if pf <= target:
            return dict(param_set=nm, k=ps.k, N=ps.N, n=ps.n, br_l=ps.bsk_l, br_b=ps.bsk_beta, ks_l=ps.ks_l,
                        ks_b=ps.ks_beta, cost=ps.modmul_per_pbs(), p_error=pf)
    raise ValueError(f"no shipped parameter set reaches p_error <= {target:.1e} for p={p}, norm2={norm2}")
