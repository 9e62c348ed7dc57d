// common.cuh -- deterministic PRNG, gadget decomposition, modulus switch, encoding.
#pragma once
#include "fq.cuh"

// ---- counter-based PRNG: rnd64(seed, domain, index) ----------------------------------------------
FQ_HD u64 fbs_mix64(u64 z)
{
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
FQ_HD u64 fbs_rnd64(u64 seed, u64 dom, u64 idx)
{
    u64 h = fbs_mix64(seed ^ (dom * 0xD1B54A32D192ED03ULL));
    return fbs_mix64(h + (idx + 1) * 0x9E3779B97F4A7C15ULL);
}
FQ_HD u64 fbs_rnd_uniform(u64 seed, u64 dom, u64 idx)
{
    // 60 random bits; the 0.02 % of draws that land in [q, 2^60) are replaced by a second, scaled draw
    u64 u = fbs_rnd64(seed, dom, idx) >> 4;
    if (u >= FQ_Q) u = fq_mulhi(fbs_rnd64(seed, dom + 64, idx), FQ_Q);
    return u;
}
// Irwin-Hall(12) noise with standard deviation `scale` (units of 1/q): integer-only so that host, device
// and oracle produce identical samples.
FQ_HD u64 fbs_rnd_noise(u64 seed, u64 dom, u64 idx, u64 scale)
{
    u64 S = 0;
#pragma unroll
    for (int t = 0; t < 6; t++) {
        u64 r = fbs_rnd64(seed, dom, idx * 6 + t);
        S += (r & 0xFFFFFFFFULL) + (r >> 32);
    }
    i64 c = (i64)S - (i64)(6ULL * 0xFFFFFFFFULL);
    // e = floor((c*scale + 2^31) / 2^32) with a 128-bit signed product
#if defined(__CUDA_ARCH__)
    i64 hi = __mul64hi(c, (i64)scale);
    u64 lo = (u64)c * scale;
#else
    __int128 pr = (__int128)c * (__int128)scale;
    i64 hi = (i64)(pr >> 64);
    u64 lo = (u64)pr;
#endif
    u64 lo2 = lo + 0x80000000ULL;
    hi += (lo2 < lo) ? 1 : 0;
    i64 e = (i64)(((u64)hi << 32) | (lo2 >> 32));
    return fq_from_i64(e);
}
enum { DOM_SLWE = 1, DOM_SGLWE = 2, DOM_BSK_MASK = 3, DOM_BSK_NOISE = 4, DOM_KSK_MASK = 5, DOM_KSK_NOISE = 6,
       DOM_ENC_MASK = 7, DOM_ENC_NOISE = 8 };

// ---- encoding --------------------------------------------------------------------------------------
FQ_HD u64 fbs_delta(int p) { return (FQ_Q + (u64)p) / (2ULL * (u64)p); }   // round(Q / 2p)

static inline u64 fbs_gadget_host(int beta, int j)                          // round(Q / B^(j+1))
{
    unsigned __int128 B = (unsigned __int128)1 << (beta * (j + 1));
    return (u64)(((unsigned __int128)FQ_Q + B / 2) / B);
}

// y = round(x * 2^bits / q) mod 2^bits for x in [0, q)  (the rounding used by gadget decomposition and modulus switch).
// Write x = r1 + p1 * t (r1 = x mod p1, t = floor(x / p1) in [0, p2)): x/q = t/p2 + r1/q.  Up to 24 bits
//     y = (t * K63 + 8 * r1 + 2^(s-1)) >> s,   s = 63 - bits,  K63 = floor(2^63 / p2),  8 ~ 2^63 / q
// which needs only the CRT digit t, no 60-bit reconstruction, inside the blind-rotation loop; the truncation errors
// are below 2^-9 of a rounding step, so the rounding is unbiased (a biased rounding accumulates over the n*N
// decompositions of a blind rotation).  Beyond 24 bits the full integer is scaled by a 64-bit reciprocal.
#define FBS_K63 0x20006000AULL        /* floor(2^63 / p2) */
FQ_HD u64 fbs_round_top_t(u32 t, u32 r1, int bits)
{
    const int s = 63 - bits;
    return (((u64)t * FBS_K63 + ((u64)r1 << 3) + (1ULL << (s - 1))) >> s) & ((1ULL << bits) - 1);
}
FQ_HD u64 fbs_round_top(u64 x, int bits)
{
    if (bits <= 24) { const u64 t = x / FQ_P1; return fbs_round_top_t((u32)t, (u32)(x - t * FQ_P1), bits); }
    const u64 s = fq_mulhi(x, FQ_RQ);                     // x * 2^59 / q
    return ((s + (1ULL << (58 - bits))) >> (59 - bits)) & ((1ULL << bits) - 1);
}
FQ_HD u32 fbs_modswitch(u64 x, int log2_2N) { return (u32)fbs_round_top(x, log2_2N); }

// balanced base-2^beta digits of y (bl = beta*l bits), d[0] = most significant level; digits in [-B/2, B/2)
template <int L>
FQ_HD void fbs_balanced_digits(u64 y, int beta, int (&d)[L])
{
    const u64 Bm = (1ULL << beta) - 1, half = 1ULL << (beta - 1);
#pragma unroll
    for (int j = L - 1; j >= 0; j--) {
        u64 dig = y & Bm;
        y >>= beta;
        int c = dig >= half;
        d[j] = (int)dig - (c << beta);
        y += c;
    }
}

// One-level case (L = 1, beta <= 24) of the two functions above in six instructions: the digit is y sign-extended from
// beta bits.  With W = t*K63 + 8*r1 + 2^(s-1), s = 63 - beta, y occupies bits [s, 63) of W, i.e. bits [31-beta, 31) of
// W's high word; K63 = 2^33 + 0x6000A, so W_hi = hi32(t*0x6000A + 8*r1 + 2^(s-1)) + 2t.  `rc` = 2^(s-1) = 2^(62-beta).
FQ_HD int fbs_digit1_t(u32 t, u32 r1, int beta, u64 rc)
{
    // beta <= 24: rc >= 2^38 has no low word, 8*r1 < 2^33 splits into (r1 << 3, r1 >> 29)
    const u64 v = r32_madwide2(t, (u32)(FBS_K63 & 0xFFFFFFFFu), r1 << 3, (r1 >> 29) + (u32)(rc >> 32));
    const u32 whi = (u32)(v >> 32) + 2u * t;
    return (int)(whi << 1) >> (32 - beta);
}

// decode phase -> message in Z_2p
FQ_HD int fbs_decode(u64 phase, int p)
{
    u64 delta = fbs_delta(p), half = delta >> 1;
    if (phase >= FQ_Q - half) return 0;
    return (int)(((phase + half) / delta) % (2ULL * (u64)p));
}
