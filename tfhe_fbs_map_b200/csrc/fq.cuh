// fq.cuh -- arithmetic modulo the ciphertext modulus q = p1 * p2, for host and device.
//
//   p1 = 0x3FFE8001 = 1073643521,  p2 = 0x3FFF4001 = 1073692673   (primes, both = 1 mod 2^14, both < 2^30)
//   q  = p1 * p2 = 0x0FFF70019FFDC001  (60 bits)
//
// Ciphertexts in HBM are plain integers in [0, q) (what oracle/tfhe_ref.c computes with).  Because q is a product of
// two NTT-friendly 30-bit primes, the blind-rotation kernel works in the residue number system (RNS): every ring
// element is a pair (x mod p1, x mod p2) packed in one 64-bit word, negacyclic NTTs run independently per prime
// with 32-bit Harvey lazy butterflies and Shoup twiddles (3 IMAD + 4 ALU instructions per prime and butterfly), and
// CRT is only needed where a non-linear function of the integer is taken (gadget decomposition, sample extraction).
// Z_q[X]/(X^N+1) is isomorphic to the product of the two prime rings, so results are exactly those of integer
// arithmetic mod q: the CUDA path stays comparable bit for bit with the oracle.
//
// History (profiles/): v1 Goldilocks 2^64-2^32+1 was 80 % ALU-pipe bound (carry tests); v2/v3 a 62-bit prime with
// 64-bit Harvey butterflies was 79 % FMA-heavy-pipe bound (six IMAD.WIDE + four IMAD per butterfly).
#pragma once
#include <stdint.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;

#define FQ_P1 0x3FFE8001u
#define FQ_P2 0x3FFF4001u
#define FQ_Q 0x0FFF70019FFDC001ULL
#define FQ_QBITS 60
#define FQ_QN 0xFFF70019FFDC0010ULL       /* q << 4: normalised divisor for 128-by-64 reduction */
#define FQ_QNV 0x000900370129060FULL      /* floor((2^128-1) / (q<<4)) - 2^64  (Moeller-Granlund reciprocal) */
#define FQ_RQ 0x8004801B80948307ULL       /* floor(2^(64+59) / q): reciprocal for round(x * 2^bits / q) */
#define FQ_P1_INVNEG 0xFFFE7FFFu           /* -p1^-1 mod 2^32 */
#define FQ_P2_INVNEG 0xAFFF3FFFu           /* -p2^-1 mod 2^32 */
#define FQ_P1INV_P2 357919402u            /* p1^-1 mod p2 */
#define FQ_P1INV_P2_S 1431743146u         /* floor(p1inv * 2^32 / p2) */

#if defined(__CUDACC__)
#define FQ_HD __host__ __device__ __forceinline__
#define FQ_HDM __host__ __device__ __forceinline__   /* for class members */
#else
#define FQ_HD static inline
#define FQ_HDM inline
#endif

// ---------------------------------------------------------------------------------------------- integers mod q
FQ_HD u64 fq_csub(u64 x, u64 m) { return x >= m ? x - m : x; }
FQ_HD u64 fq_add(u64 a, u64 b) { return fq_csub(a + b, FQ_Q); }          // canonical in, canonical out
FQ_HD u64 fq_sub(u64 a, u64 b) { return a >= b ? a - b : a + FQ_Q - b; }
FQ_HD u64 fq_neg(u64 a) { return a ? FQ_Q - a : 0; }
FQ_HD u64 fq_from_i64(i64 v) { return v >= 0 ? (u64)v : FQ_Q - (u64)(-v); }  // |v| < q

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
// (hi:lo) mod q for hi < q, canonical.  Division of a two-word number by the normalised divisor q<<4 with a
// precomputed reciprocal (Moeller & Granlund, "Improved division by invariant integers", algorithm 4).
FQ_HD u64 fq_reduce128(u64 lo, u64 hi)
{
    const u64 u1 = (hi << 4) | (lo >> 60), u0 = lo << 4;
    u64 q0, q1;
    fq_mul_wide(FQ_QNV, u1, q0, q1);
    q0 += u0;
    q1 += u1 + (q0 < u0 ? 1 : 0) + 1;
    u64 r = u0 - q1 * FQ_QN;
    if (r > q0) r += FQ_QN;
    if (r >= FQ_QN) r -= FQ_QN;
    return r >> 4;
}
FQ_HD u64 fq_mul(u64 a, u64 b)                         // canonical in, canonical out (not used in hot loops)
{
    u64 lo, hi;
    fq_mul_wide(a, b, lo, hi);
    return fq_reduce128(lo, hi);
}

// ---------------------------------------------------------------------------------------------- one 30-bit prime
// Lazy ranges: forward butterflies take and return values in [0, 4p), inverse butterflies [0, 2p) (4p < 2^32).
FQ_HD u32 r32_fold(u32 x, u32 p2x)                     // x in [0,4p) -> [0,2p): min(x, x - 2p) as unsigned
{
    const u32 y = x - p2x;
    return y < x ? y : x;
}
FQ_HD u32 r32_mulhi(u32 a, u32 b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (u32)(((u64)a * b) >> 32);
#endif
}
// Shoup multiplication by the constant w (ws = floor(w * 2^32 / p)): w*y mod p + {0, p}, for ANY 32-bit y
FQ_HD u32 r32_mul_shoup(u32 y, u32 w, u32 ws, u32 p) { return w * y - r32_mulhi(ws, y) * p; }
// 32 x 32 -> 64 multiply(-add).  On the device these are spelled as mul.wide / mad.wide: left to the compiler the operands
// get widened to 64 bits first and every product drags an extra add of a zero high word along.
FQ_HD u64 r32_mulwide(u32 a, u32 b)
{
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
#else
    return (u64)a * b;
#endif
}
FQ_HD u64 r32_madwide(u32 a, u32 b, u64 c)
{
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
#else
    return (u64)a * b + c;
#endif
}
FQ_HD u64 r32_madwide2(u32 a, u32 b, u32 clo, u32 chi)     // a*b + (chi:clo)
{
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("{\n\t.reg .b64 c;\n\tmov.b64 c, {%3, %4};\n\tmad.wide.u32 %0, %1, %2, c;\n\t}" : "=l"(r) : "r"(a), "r"(b), "r"(clo), "r"(chi));
    return r;
#else
    return (u64)a * b + (((u64)chi << 32) | clo);
#endif
}
// Montgomery reduction of acc < 2^63: a representative of acc * 2^-32 mod p in (0, acc/2^32 + p].  Borrow-free form:
// with m = lo(acc) * p^-1 mod 2^32 the low words of acc and m*p agree, so (acc - m*p) / 2^32 = hi(acc) - mulhi(m, p).
// (`z` is the kernels' run-time zero that keeps the subtraction a 3-input IADD3 on the ALU pipe, see ntt.cuh)
FQ_HD u32 r32_redc(u64 acc, u32 p, u32 pinv_neg, u32 z = 0)
{
    const u32 m = (u32)acc * (0u - pinv_neg);
    return (u32)(acc >> 32) - r32_mulhi(m, p) + z + p;
}
FQ_HD u32 r32_csub(u32 x, u32 p)                       // x >= p ? x - p : x, written as an unsigned min (one VIADDMNMX)
{
    const u32 y = x - p;
    return y < x ? y : x;
}

// residue pair <-> packed word
struct rns2 { u32 a, b; };
FQ_HD u64 rns_pack(rns2 v) { return (u64)v.a | ((u64)v.b << 32); }
FQ_HD rns2 rns_unpack(u64 w) { rns2 v; v.a = (u32)w; v.b = (u32)(w >> 32); return v; }
FQ_HD rns2 rns_split(u64 w)                             // rns_unpack with the halves as opaque 32-bit registers
{
    rns2 v;
#if defined(__CUDA_ARCH__)
    asm("mov.b64 {%0, %1}, %2;" : "=r"(v.a), "=r"(v.b) : "l"(w));
#else
    v.a = (u32)w; v.b = (u32)(w >> 32);
#endif
    return v;
}
FQ_HD rns2 rns_from_int(u64 x) { rns2 v; v.a = (u32)(x % FQ_P1); v.b = (u32)(x % FQ_P2); return v; }   // x < 2^64
FQ_HD rns2 rns_from_small(int d)                      // |d| < 2^29
{
    rns2 v;
    v.a = d >= 0 ? (u32)d : FQ_P1 - (u32)(-d);
    v.b = d >= 0 ? (u32)d : FQ_P2 - (u32)(-d);
    return v;
}
// CRT: canonical residues -> t = floor(x / p1) in [0, p2)  (x = a + p1 * t)
FQ_HD u32 rns_crt_hi(rns2 v)
{
    const u32 diff = v.b + FQ_P2 - v.a;                // a < p1 < p2, so diff in (0, 2 p2)
    return r32_csub(r32_mul_shoup(diff, FQ_P1INV_P2, FQ_P1INV_P2_S, FQ_P2), FQ_P2);
}
FQ_HD u64 rns_to_int(rns2 v) { return (u64)v.a + (u64)FQ_P1 * rns_crt_hi(v); }
FQ_HD rns2 rns_neg(rns2 v) { rns2 r; r.a = v.a ? FQ_P1 - v.a : 0; r.b = v.b ? FQ_P2 - v.b : 0; return r; }
FQ_HD rns2 rns_sub(rns2 x, rns2 y)                    // canonical
{
    rns2 r;
    r.a = x.a >= y.a ? x.a - y.a : x.a + FQ_P1 - y.a;
    r.b = x.b >= y.b ? x.b - y.b : x.b + FQ_P2 - y.b;
    return r;
}
FQ_HD rns2 rns_add(rns2 x, rns2 y) { rns2 r; r.a = r32_csub(x.a + y.a, FQ_P1); r.b = r32_csub(x.b + y.b, FQ_P2); return r; }

// ---------------------------------------------------------------------------------------------- host helpers
static inline u64 pow_mod_host(u64 b, u64 e, u64 m)
{
    unsigned __int128 r = 1, x = b % m;
    while (e) { if (e & 1) r = r * x % m; x = x * x % m; e >>= 1; }
    return (u64)r;
}
static inline u32 shoup32_host(u32 w, u32 p) { return (u32)(((u64)w << 32) / p); }
