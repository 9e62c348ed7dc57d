// fq.cuh -- arithmetic in Z_Q, Q = 2^62 - 2^16 + 1 (prime, Q-1 = 2^16 * (2^46-1): negacyclic NTTs up to N = 2^15),
// for host and device.  Canonical representatives live in [0, Q); "lazy" values in [0, 2Q) or [0, 4Q) (4Q < 2^64)
// are used inside the NTT (Harvey butterflies with Shoup-precomputed twiddles) and around the Montgomery reduction of
// the bootstrapping-key products.  Everything that leaves a kernel is canonical, which is what makes the CUDA path
// comparable bit for bit with oracle/tfhe_ref.c.
//
// Why not Goldilocks (2^64 - 2^32 + 1)?  It has no headroom in a 64-bit word: every add/sub needs a carry test and a
// select, which lands on the ALU pipe.  The first version of this kernel was 80 % ALU-pipe bound with the IMAD pipe
// at 22 % (profiles/r1_v1_goldilocks_*).  With two spare bits the butterflies need no canonicalisation and their
// work moves to IMAD.WIDE / IMAD on the otherwise idle FMA pipe.
#pragma once
#include <stdint.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;

#define FQ_Q 0x3FFFFFFFFFFF0001ULL        /* 2^62 - 2^16 + 1 */
#define FQ_2Q 0x7FFFFFFFFFFE0002ULL
#define FQ_QINV_NEG 0x3FFEFFFEFFFEFFFFULL /* -Q^{-1} mod 2^64 */
#define FQ_R 0x3FFFCULL                   /* 2^64 mod Q */

#if defined(__CUDACC__)
#define FQ_HD __host__ __device__ __forceinline__
#define FQ_HDM __host__ __device__ __forceinline__   /* for class members */
#else
#define FQ_HD static inline
#define FQ_HDM inline
#endif

FQ_HD u64 fq_csub(u64 x, u64 m) { return x >= m ? x - m : x; }          // conditional subtract
// lazy fold used by the forward butterflies: subtract 2Q iff bit 63 is set (a single test on the high word)
FQ_HD u64 fq_lazy_fold(u64 x) { return ((i64)x < 0) ? x - FQ_2Q : x; }
FQ_HD u64 fq_add(u64 a, u64 b) { return fq_csub(a + b, FQ_Q); }          // canonical in, canonical out
FQ_HD u64 fq_sub(u64 a, u64 b) { return a >= b ? a - b : a + FQ_Q - b; }
FQ_HD u64 fq_neg(u64 a) { return a ? FQ_Q - a : 0; }
FQ_HD u64 fq_from_i64(i64 v) { return v >= 0 ? (u64)v : FQ_Q - (u64)(-v); }  // |v| < Q

FQ_HD void fq_mul_wide(u64 a, u64 b, u64 &lo, u64 &hi)
{
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    lo = (u64)x; hi = (u64)(x >> 64);
#endif
}
FQ_HD u64 fq_mulhi(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}
// x = hi*2^64 + lo < 2^124  ->  x mod Q, canonical.  2^62 = 2^16 - 1 (mod Q):
//   x = a*2^62 + b,  a = a1*2^46 + a0   =>   x = b + a0*2^16 + a1*(2^16-1) - a   (mod Q)
FQ_HD u64 fq_reduce128(u64 lo, u64 hi)
{
    const u64 b = lo & 0x3FFFFFFFFFFFFFFFULL;
    const u64 a = (hi << 2) | (lo >> 62);
    const u64 a0 = a & 0x3FFFFFFFFFFFULL, a1 = a >> 46;
    const u64 pos = b + (a0 << 16) + (a1 << 16);        // < 2^63 + 2^32
    const u64 neg = a + a1;                             // pos - neg > -Q
    u64 r = pos - neg;
    if (pos < neg) r += FQ_Q;
    r = fq_csub(r, FQ_2Q);
    return fq_csub(r, FQ_Q);
}
FQ_HD u64 fq_mul(u64 a, u64 b)                         // canonical in, canonical out (not used in hot loops)
{
    u64 lo, hi;
    fq_mul_wide(a, b, lo, hi);
    return fq_reduce128(lo, hi);
}
// Shoup / Harvey lazy multiplication by a constant w with ws = floor(w * 2^64 / Q):
// returns w*y mod Q + {0, Q}, i.e. a value in [0, 2Q), for ANY 64-bit y.
FQ_HD u64 fq_mul_shoup(u64 y, u64 w, u64 ws) { return w * y + fq_mulhi(ws, y) * (0ULL - FQ_Q); }   // all IMAD, no subtract
// Montgomery reduction: (hi:lo) * 2^-64 mod Q, lazy: result < hi + Q + 1
FQ_HD u64 fq_redc(u64 lo, u64 hi)
{
    const u64 m = lo * FQ_QINV_NEG;
    return hi + fq_mulhi(m, FQ_Q) + (lo != 0 ? 1 : 0);
}

static inline u64 fq_pow_host(u64 b, u64 e)
{
    u64 r = 1;
    while (e) { if (e & 1) r = fq_mul(r, b); b = fq_mul(b, b); e >>= 1; }
    return r;
}
static inline u64 fq_shoup_host(u64 w) { return (u64)((((unsigned __int128)w) << 64) / FQ_Q); }
