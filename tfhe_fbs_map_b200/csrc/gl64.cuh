// gl64.cuh -- arithmetic in Z_P, P = 2^64 - 2^32 + 1 (Goldilocks), for host and device.
// All public functions take and return CANONICAL representatives in [0, P): that is what makes the CUDA
// path comparable bit for bit with oracle/tfhe_ref.c at every ciphertext tap.
#pragma once
#include <stdint.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL   /* 2^64 mod P */

#if defined(__CUDACC__)
#define GL_HD __host__ __device__ __forceinline__
#define GL_HDM __host__ __device__ __forceinline__   /* for class members */
#else
#define GL_HD static inline
#define GL_HDM inline
#endif

GL_HD u64 gl_add(u64 a, u64 b)
{
    u64 s = a + b;
    // a + b < 2P: subtract P when the 64-bit add wrapped or the sum is not canonical
    return (s < a || s >= GL_P) ? s - GL_P : s;
}
GL_HD u64 gl_sub(u64 a, u64 b)
{
    u64 d = a - b;
    return (a < b) ? d + GL_P : d;
}
GL_HD u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }

// x = hi*2^64 + lo  ->  x mod P.   2^64 = 2^32-1 and 2^96 = -1 (mod P):
// x = lo - (hi >> 32) + (hi & 0xffffffff) * (2^32 - 1)
GL_HD u64 gl_reduce128(u64 lo, u64 hi)
{
    u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= GL_EPS;            // wrapped by 2^64 = EPS (mod P)
    u64 t1 = hl * GL_EPS;                 // < 2^64
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS;              // cannot wrap twice
    return r >= GL_P ? r - GL_P : r;
}
GL_HD void gl_mul_wide(u64 a, u64 b, u64 &lo, u64 &hi)
{
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    lo = (u64)x; hi = (u64)(x >> 64);
#endif
}
GL_HD u64 gl_mul(u64 a, u64 b)
{
    u64 lo, hi;
    gl_mul_wide(a, b, lo, hi);
    return gl_reduce128(lo, hi);
}
// signed small integer -> field element
GL_HD u64 gl_from_i64(i64 v) { return v >= 0 ? (u64)v : GL_P - (u64)(-v); }

static inline u64 gl_pow_host(u64 b, u64 e)
{
    u64 r = 1;
    while (e) { if (e & 1) r = gl_mul(r, b); b = gl_mul(b, b); e >>= 1; }
    return r;
}
