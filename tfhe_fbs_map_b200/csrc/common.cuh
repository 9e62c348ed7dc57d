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
    u64 u = fbs_rnd64(seed, dom, idx) >> 2;          // 62 bits; Q = 2^62 - 2^16 + 1, bias 2^-46
    return u >= FQ_Q ? u - FQ_Q : u;
}
// Irwin-Hall(12) noise with standard deviation `scale` (units of 1/Q): integer-only so that host, device
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

// closest multiple of Q/2^bits as the bits-wide integer y = round(x * 2^bits / 2^62) (wraps: y == 2^bits -> 0).
// Q differs from 2^62 by 2^-46 relative, far below the rounding step for every bits <= 48.
FQ_HD u64 fbs_round_top(u64 x, int bits)
{
    const u64 t = x + (1ULL << (61 - bits));
    return (t >> (62 - bits)) & ((1ULL << bits) - 1);
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

// decode phase -> message in Z_2p
FQ_HD int fbs_decode(u64 phase, int p)
{
    u64 delta = fbs_delta(p), half = delta >> 1;
    if (phase >= FQ_Q - half) return 0;
    return (int)(((phase + half) / delta) % (2ULL * (u64)p));
}
