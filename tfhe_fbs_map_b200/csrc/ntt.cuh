// ntt.cuh -- negacyclic NTT over Z_q[X]/(X^N+1), q = p1*p2, in RNS: one polynomial held by N/8 threads, 8 coefficients
// (= 8 residue pairs) each.
//
// Structure: radix-8 register passes (3 butterfly stages on the 8 values a thread
// holds) separated by shared-memory transposes.  Forward = Cooley-Tukey, natural order in, bit-reversed
// out, negacyclic twist merged into the twiddles psi_rev[m+i] = psi^bitrev(m+i); inverse = Gentleman-
// Sande, bit-reversed in, natural out, WITHOUT the 1/N scale (it is folded into the bootstrapping key).
// Butterflies are Harvey's lazy ones with Shoup twiddles (w, ws = floor(w*2^32/p)), independently per prime:
//   forward: inputs and outputs in [0,4p);   inverse: inputs and outputs in [0,2p).
//
// A pass is described by `lb`, the lowest of its three in-thread index bits:
//     idx(tau, e) = ((tau >> lb) << (lb+3)) | (e << lb) | (tau & ((1<<lb)-1)),   e = 0..7
// Everything here is __host__ __device__ so tests/host_emul.cu can run the exact index / twiddle logic on
// the CPU (this container has no GPU) by looping over tau and treating a transpose as an array copy.
#pragma once
#include "fq.cuh"

template <int LOGN>
struct NttPlan {
    static constexpr int N = 1 << LOGN;
    static constexpr int T = N / 8;                 // threads per polynomial
    static constexpr int NPASS = (LOGN + 2) / 3;
    // forward pass p covers index bits hi = LOGN-1-3p down to max(hi-2, 0)
    FQ_HDM static constexpr int fwd_hi(int p) { return LOGN - 1 - 3 * p; }
    FQ_HDM static constexpr int fwd_lb(int p) { return fwd_hi(p) - 2 > 0 ? fwd_hi(p) - 2 : 0; }
    // inverse pass p holds index bits lb..lb+2 in registers and processes those not done yet (bits >= inv_lo(p)).  The layouts
    // mirror the forward ones (0, 3, .., LOGN-6, LOGN-3): the short pass comes second to last, so that the three highest
    // index bits stay the WARP bits until the last transpose and only that one crosses warps.
    FQ_HDM static constexpr int inv_lb(int p) { return p == NPASS - 1 ? LOGN - 3 : (3 * p < LOGN - 6 ? 3 * p : (LOGN - 6 > 0 ? LOGN - 6 : 0)); }
    FQ_HDM static constexpr int inv_lo(int p) { return p == 0 ? 0 : inv_lb(p - 1) + 3; }
    // bank-conflict-free XOR swizzle of the low 4 index bits (tools/swizzle_search.py):
    // 4-bit columns added for index bits 4, 5, 6
    FQ_HDM static constexpr int sw_c0() { return (LOGN % 3 == 2) ? 1 : (LOGN % 3 == 0) ? 1 : 2; }
    FQ_HDM static constexpr int sw_c1() { return (LOGN % 3 == 2) ? 4 : (LOGN % 3 == 0) ? 2 : 4; }
    FQ_HDM static constexpr int sw_c2() { return (LOGN % 3 == 2) ? 10 : (LOGN % 3 == 0) ? 12 : 9; }
    FQ_HDM static constexpr int swz(int i)
    {
        return i ^ (((i >> 4) & 1) * sw_c0()) ^ (((i >> 5) & 1) * sw_c1()) ^ (((i >> 6) & 1) * sw_c2());
    }
    FQ_HDM static constexpr int idx(int tau, int e, int lb)
    {
        return ((tau >> lb) << (lb + 3)) | (e << lb) | (tau & ((1 << lb) - 1));
    }
    // swz and idx are GF(2)-linear in (tau, e): swz(idx(tau, e, lb)) = swz(idx(tau, 0, lb)) ^ swz(idx(0, e, lb)).  The kernels keep
    // the thread part as a BYTE offset in a register (one per layout) and XOR the compile-time element part into it.
    // thread that holds logical index i in layout lb, and whether a transpose lb0 -> lb1 only moves words between threads of
    // the same warp (then __syncwarp() orders it, no block-level barrier needed)
    FQ_HDM static constexpr int owner(int i, int lb) { return ((i >> (lb + 3)) << lb) | (i & ((1 << lb) - 1)); }
    FQ_HDM static constexpr bool intra_warp(int lb0, int lb1)
    {
        for (int i = 0; i < N; i++) if ((owner(i, lb0) >> 5) != (owner(i, lb1) >> 5)) return false;
        return true;
    }
    // the same with 2^wlog thread positions per warp (prime-split kernels: a warp = 16 positions x 2 primes)
    FQ_HDM static constexpr bool intra_warp_w(int lb0, int lb1, int wlog)
    {
        for (int i = 0; i < N; i++) if ((owner(i, lb0) >> wlog) != (owner(i, lb1) >> wlog)) return false;
        return true;
    }
    FQ_HDM static constexpr u32 tau_boff(int tau, int lb) { return 8u * (u32)swz(idx(tau, 0, lb)); }
    FQ_HDM static constexpr u32 elem_boff(int e, int lb) { return 8u * (u32)swz(idx(0, e, lb)); }
};
static_assert(NttPlan<11>::swz(NttPlan<11>::idx(173, 5, 2)) * 8 == (NttPlan<11>::tau_boff(173, 2) ^ NttPlan<11>::elem_boff(5, 2)), "linear addressing");
static_assert(NttPlan<10>::swz(NttPlan<10>::idx(97, 3, 4)) * 8 == (NttPlan<10>::tau_boff(97, 4) ^ NttPlan<10>::elem_boff(3, 4)), "linear addressing");

// ---- one forward pass on registers -------------------------------------------------------------------
struct alignas(16) fq_tw { u32 w1, ws1, w2, ws2; };      // twiddle + Shoup companion for both primes: one 128-bit load
template <bool TWS = false>                              // TWS: the table lives in shared memory
FQ_HD fq_tw fq_tw_load(const fq_tw *p)
{
#if defined(__CUDA_ARCH__)
    if constexpr (TWS) {
        uint4 v;
        asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((u32)__cvta_generic_to_shared(p)));
        return fq_tw{v.x, v.y, v.z, v.w};
    } else {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
        return fq_tw{v.x, v.y, v.z, v.w};
    }
#else
    return *p;
#endif
}

// ptxas balances 2-input integer adds between IADD3 (ALU pipe) and IMAD.IADD (FMA-heavy pipe) assuming equal pipe
// load; in these kernels the FMA-heavy pipe is the busier one (IMAD, IMAD.HI).  Adding a zero that only exists at run
// time (a kernel argument) makes the add 3-input, which only IADD3 can do, and pins it to the ALU pipe.
struct NttZero { u32 z; };
// NB = bootstraps per thread: the NB butterflies at one position share the twiddle load and the index arithmetic
template <int LOGN, int PASS, int NB, bool TWS = false>
FQ_HD void ntt_fwd_pass_n(rns2 (&x)[NB][8], int tau, const fq_tw *__restrict__ psi_rev, u32 z = 0)
{
    using P = NttPlan<LOGN>;
    constexpr int hi = P::fwd_hi(PASS), lb = P::fwd_lb(PASS);
    const int th = tau >> lb;
#pragma unroll
    for (int q = 2; q >= 0; q--) {
        if (lb + q > hi) continue;                       // bit already handled by an earlier pass
        const int bit = lb + q, m = 1 << (LOGN - 1 - bit);
#pragma unroll
        for (int e0 = 0; e0 < 8; e0++) {
            if (e0 & (1 << q)) continue;
            const int e1 = e0 | (1 << q);
            const fq_tw w = fq_tw_load<TWS>(psi_rev + m + ((th << (2 - q)) | (e0 >> (q + 1))));
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const u32 ua = r32_fold(x[b][e0].a, 2 * FQ_P1), va = r32_mul_shoup(x[b][e1].a, w.w1, w.ws1, FQ_P1);
                const u32 ub = r32_fold(x[b][e0].b, 2 * FQ_P2), vb = r32_mul_shoup(x[b][e1].b, w.w2, w.ws2, FQ_P2);
                x[b][e0].a = ua + va + z; x[b][e1].a = ua - va + z + 2 * FQ_P1;
                x[b][e0].b = ub + vb + z; x[b][e1].b = ub - vb + z + 2 * FQ_P2;
            }
        }
    }
}
template <int LOGN, int PASS>
FQ_HD void ntt_fwd_pass(rns2 (&x)[8], int tau, const fq_tw *__restrict__ psi_rev, u32 z = 0)
{
    ntt_fwd_pass_n<LOGN, PASS, 1>(*reinterpret_cast<rns2(*)[1][8]>(&x), tau, psi_rev, z);
}
// ---- one inverse pass on registers ---------------------------------------------------------------------
// MIRROR: psi^-bitrev(m+i) = -psi^bitrev(2m-1-i), so the inverse can read the FORWARD table mirrored within each block
// [m, 2m) and swap the operands of its subtraction; the blind-rotate kernel does, which halves the twiddle footprint in L1.
template <int LOGN, int PASS, int NB, bool MIRROR = false, bool TWS = false>
FQ_HD void ntt_inv_pass_n(rns2 (&x)[NB][8], int tau, const fq_tw *__restrict__ psi_inv_rev, u32 z = 0)
{
    using P = NttPlan<LOGN>;
    constexpr int lo = P::inv_lo(PASS), lb = P::inv_lb(PASS);
    const int th = tau >> lb;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        if (lb + q < lo) continue;                       // bit already handled by an earlier pass
        const int bit = lb + q, m = 1 << (LOGN - 1 - bit);
#pragma unroll
        for (int e0 = 0; e0 < 8; e0++) {
            if (e0 & (1 << q)) continue;
            const int e1 = e0 | (1 << q);
            const int ti = (th << (2 - q)) | (e0 >> (q + 1));
            const fq_tw w = fq_tw_load<TWS>(psi_inv_rev + (MIRROR ? 2 * m - 1 - ti : m + ti));
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const u32 ua = x[b][e0].a, va = x[b][e1].a, ub = x[b][e0].b, vb = x[b][e1].b;
                x[b][e0].a = r32_fold(ua + va + z, 2 * FQ_P1);
                x[b][e1].a = r32_mul_shoup(MIRROR ? va - ua + 2 * FQ_P1 : ua - va + 2 * FQ_P1, w.w1, w.ws1, FQ_P1);
                x[b][e0].b = r32_fold(ub + vb + z, 2 * FQ_P2);
                x[b][e1].b = r32_mul_shoup(MIRROR ? vb - ub + 2 * FQ_P2 : ub - vb + 2 * FQ_P2, w.w2, w.ws2, FQ_P2);
            }
        }
    }
}
template <int LOGN, int PASS>
FQ_HD void ntt_inv_pass(rns2 (&x)[8], int tau, const fq_tw *__restrict__ psi_inv_rev, u32 z = 0)
{
    ntt_inv_pass_n<LOGN, PASS, 1>(*reinterpret_cast<rns2(*)[1][8]>(&x), tau, psi_inv_rev, z);
}

// ---- cluster-split transform: one polynomial spread over C = 2^LOGC CTAs (k_blind_rotate_cl) ----------
// Index j = h*Ns + r, Ns = N/C.  The first LOGC forward stages (index bits LOGN-1 .. LOGN-LOGC) couple the C sub-blocks h at
// equal r; after them sub-block h is an independent transform of size Ns whose twiddles are the entries
// psi_rev[(C + h)*m + blk] of the full table (m, blk: the sub-transform's own block size / block index).  So CTA h runs the
// ordinary NttPlan<LOGN-LOGC> passes on a LOCAL table  TWF[m + x] = psi_rev[(C + h)*m + x]  (x < m; m = 1, 2, .., Ns/2),  and
// the mirrored inverse passes on  TWI[m + x] = psi_rev[(2C - 1 - h)*m + x]  (reads TWI[2m - 1 - ti] = psi_rev[(2C - h)*m - 1 - ti]
// = the mirror image of full-table index C*m + h*m + ti).  In the coefficient domain a thread owns the C sub-block values of
// R = 8/C residues r (register e = h*R + ri), so the cross stages are register butterflies with compile-time twiddle
// indices psi_rev[m + (h >> (s+1))], m = 2^(LOGC-1-s); between the cross stages and the local passes the C CTAs exchange
// registers through distributed shared memory (thread tau of CTA c <-> thread tau of CTA h: a C x C block transpose).
FQ_HD int ntt_local_src(int i, int hpre)              // full-table index of local-table entry i (1 <= i < Ns): hpre*m + x
{
    int m = 1;
    while (2 * m <= i) m *= 2;
    return hpre * m + (i - m);
}
template <int LOGC>
FQ_HD void ntt_cross_fwd(rns2 (&x)[8], const fq_tw *__restrict__ cw /* psi_rev[0 .. C) */, u32 z = 0)
{
    constexpr int C = 1 << LOGC, R = 8 / C;
#pragma unroll
    for (int s = LOGC - 1; s >= 0; s--) {
        const int m = 1 << (LOGC - 1 - s);
#pragma unroll
        for (int h0 = 0; h0 < C; h0++) {
            if (h0 & (1 << s)) continue;
            const int h1 = h0 | (1 << s);
            const fq_tw w = cw[m + (h0 >> (s + 1))];
#pragma unroll
            for (int ri = 0; ri < R; ri++) {
                const int e0 = h0 * R + ri, e1 = h1 * R + ri;
                const u32 ua = r32_fold(x[e0].a, 2 * FQ_P1), va = r32_mul_shoup(x[e1].a, w.w1, w.ws1, FQ_P1);
                const u32 ub = r32_fold(x[e0].b, 2 * FQ_P2), vb = r32_mul_shoup(x[e1].b, w.w2, w.ws2, FQ_P2);
                x[e0].a = ua + va + z; x[e1].a = ua - va + z + 2 * FQ_P1;
                x[e0].b = ub + vb + z; x[e1].b = ub - vb + z + 2 * FQ_P2;
            }
        }
    }
}
template <int LOGC>
FQ_HD void ntt_cross_inv(rns2 (&x)[8], const fq_tw *__restrict__ cw /* psi_rev[0 .. C), read mirrored */, u32 z = 0)
{
    constexpr int C = 1 << LOGC, R = 8 / C;
#pragma unroll
    for (int s = 0; s < LOGC; s++) {
        const int m = 1 << (LOGC - 1 - s);
#pragma unroll
        for (int h0 = 0; h0 < C; h0++) {
            if (h0 & (1 << s)) continue;
            const int h1 = h0 | (1 << s);
            const fq_tw w = cw[2 * m - 1 - (h0 >> (s + 1))];
#pragma unroll
            for (int ri = 0; ri < R; ri++) {
                const int e0 = h0 * R + ri, e1 = h1 * R + ri;
                const u32 ua = x[e0].a, va = x[e1].a, ub = x[e0].b, vb = x[e1].b;
                x[e0].a = r32_fold(ua + va + z, 2 * FQ_P1);
                x[e1].a = r32_mul_shoup(va - ua + 2 * FQ_P1, w.w1, w.ws1, FQ_P1);
                x[e0].b = r32_fold(ub + vb + z, 2 * FQ_P2);
                x[e1].b = r32_mul_shoup(vb - ub + 2 * FQ_P2, w.w2, w.ws2, FQ_P2);
            }
        }
    }
}

// ---- prime-split passes: a thread transforms the residues of ONE prime (l = 0: p1, l = 1: p2) of its 8 positions ----------
// Same index / twiddle logic as ntt_fwd_pass_n / ntt_inv_pass_n; the two residues of a position live in two threads (adjacent
// lanes), which doubles the threads per bootstrap and halves every thread's butterfly and point-wise work (k_blind_rotate_cs).
template <bool TWS = false>
FQ_HD void fq_tw_load1(const fq_tw *p, int l, u32 &w, u32 &ws)
{
#if defined(__CUDA_ARCH__)
    if constexpr (TWS) {
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w), "=r"(ws) : "r"((u32)__cvta_generic_to_shared(p) + 8u * (u32)l));
    } else {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p) + l);
        w = v.x; ws = v.y;
    }
#else
    w = l ? p->w2 : p->w1; ws = l ? p->ws2 : p->ws1;
#endif
}
template <int LOGN, int PASS, bool TWS = false>
FQ_HD void ntt_fwd_pass_1p(u32 (&x)[8], int tau, const fq_tw *__restrict__ psi_rev, int l, u32 p, u32 z = 0)
{
    using P = NttPlan<LOGN>;
    constexpr int hi = P::fwd_hi(PASS), lb = P::fwd_lb(PASS);
    const int th = tau >> lb;
    const u32 p2 = 2 * p;
#pragma unroll
    for (int q = 2; q >= 0; q--) {
        if (lb + q > hi) continue;
        const int bit = lb + q, m = 1 << (LOGN - 1 - bit);
#pragma unroll
        for (int e0 = 0; e0 < 8; e0++) {
            if (e0 & (1 << q)) continue;
            const int e1 = e0 | (1 << q);
            u32 w, ws;
            fq_tw_load1<TWS>(psi_rev + m + ((th << (2 - q)) | (e0 >> (q + 1))), l, w, ws);
            const u32 u = r32_fold(x[e0], p2), v = r32_mul_shoup(x[e1], w, ws, p);
            x[e0] = u + v + z; x[e1] = u - v + z + p2;
        }
    }
}
template <int LOGN, int PASS, bool TWS = false>
FQ_HD void ntt_inv_pass_1p(u32 (&x)[8], int tau, const fq_tw *__restrict__ psi_rev /* forward table, read mirrored */, int l, u32 p, u32 z = 0)
{
    using P = NttPlan<LOGN>;
    constexpr int lo = P::inv_lo(PASS), lb = P::inv_lb(PASS);
    const int th = tau >> lb;
    const u32 p2 = 2 * p;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        if (lb + q < lo) continue;
        const int bit = lb + q, m = 1 << (LOGN - 1 - bit);
#pragma unroll
        for (int e0 = 0; e0 < 8; e0++) {
            if (e0 & (1 << q)) continue;
            const int e1 = e0 | (1 << q);
            const int ti = (th << (2 - q)) | (e0 >> (q + 1));
            u32 w, ws;
            fq_tw_load1<TWS>(psi_rev + (2 * m - 1 - ti), l, w, ws);
            const u32 u = x[e0], v = x[e1];
            x[e0] = r32_fold(u + v + z, p2);
            x[e1] = r32_mul_shoup(v - u + p2, w, ws, p);
        }
    }
}
template <int LOGC>
FQ_HD void ntt_cross_fwd_1p(u32 (&x)[8], const u32 (&cw)[1 << LOGC], const u32 (&cws)[1 << LOGC], u32 p, u32 z = 0)
{
    constexpr int C = 1 << LOGC, R = 8 / C;
    const u32 p2 = 2 * p;
#pragma unroll
    for (int s = LOGC - 1; s >= 0; s--) {
        const int m = 1 << (LOGC - 1 - s);
#pragma unroll
        for (int h0 = 0; h0 < C; h0++) {
            if (h0 & (1 << s)) continue;
            const int h1 = h0 | (1 << s), ti = m + (h0 >> (s + 1));
#pragma unroll
            for (int ri = 0; ri < R; ri++) {
                const int e0 = h0 * R + ri, e1 = h1 * R + ri;
                const u32 u = r32_fold(x[e0], p2), v = r32_mul_shoup(x[e1], cw[ti], cws[ti], p);
                x[e0] = u + v + z; x[e1] = u - v + z + p2;
            }
        }
    }
}
template <int LOGC>
FQ_HD void ntt_cross_inv_1p(u32 (&x)[8], const u32 (&cw)[1 << LOGC], const u32 (&cws)[1 << LOGC], u32 p, u32 z = 0)
{
    constexpr int C = 1 << LOGC, R = 8 / C;
    const u32 p2 = 2 * p;
#pragma unroll
    for (int s = 0; s < LOGC; s++) {
        const int m = 1 << (LOGC - 1 - s);
#pragma unroll
        for (int h0 = 0; h0 < C; h0++) {
            if (h0 & (1 << s)) continue;
            const int h1 = h0 | (1 << s), ti = 2 * m - 1 - (h0 >> (s + 1));
#pragma unroll
            for (int ri = 0; ri < R; ri++) {
                const int e0 = h0 * R + ri, e1 = h1 * R + ri;
                const u32 u = x[e0], v = x[e1];
                x[e0] = r32_fold(u + v + z, p2);
                x[e1] = r32_mul_shoup(v - u + p2, cw[ti], cws[ti], p);
            }
        }
    }
}

#if defined(__CUDACC__)
// ---- device drivers: transposes through two alternating swizzled buffers -------------------------------
// `sync` is a callable that synchronises the T threads working on this polynomial.
template <int LOGN, int PASS, class Sync>
__device__ __forceinline__ void ntt_fwd_from(rns2 (&x)[8], int tau, u64 *bufA, u64 *bufB, const fq_tw *psi_rev, Sync sync)
{
    using P = NttPlan<LOGN>;
    ntt_fwd_pass<LOGN, PASS>(x, tau, psi_rev);
    if constexpr (PASS + 1 < P::NPASS) {
        u64 *buf = (PASS & 1) ? bufB : bufA;
        constexpr int lb0 = P::fwd_lb(PASS), lb1 = P::fwd_lb(PASS + 1);
#pragma unroll
        for (int e = 0; e < 8; e++) buf[P::swz(P::idx(tau, e, lb0))] = rns_pack(x[e]);
        sync();
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = rns_unpack(buf[P::swz(P::idx(tau, e, lb1))]);
        ntt_fwd_from<LOGN, PASS + 1>(x, tau, bufA, bufB, psi_rev, sync);
    }
}
// forward NTT: x enters in layout fwd_lb(0) (idx = tau + e*T), leaves in layout fwd_lb(NPASS-1) = 0 (idx = 8*tau+e),
// values at bit-reversed positions.
template <int LOGN, class Sync>
__device__ __forceinline__ void ntt_forward(rns2 (&x)[8], int tau, u64 *bufA, u64 *bufB, const fq_tw *psi_rev, Sync sync)
{
    ntt_fwd_from<LOGN, 0>(x, tau, bufA, bufB, psi_rev, sync);
}
// ---- single-buffer variants: one scratch polynomial per thread group and bootstrap, two barriers per transpose ------
// (used by the blind-rotate kernel, where shared memory rather than barrier count limits residency).  A thread may carry
// NB bootstraps: their scratch polynomials are `stride` words apart and are transposed under the same barriers.
// ONE barrier per transpose suffices: a word's address depends only on its logical index, so the words a thread writes
// for transpose t+1 are exactly the words it read itself in transpose t (nobody else reads them in between).  Only the
// first write needs the caller's guarantee (`buf_free`) that the previous users of the scratch are done.  Transposes that
// stay inside a warp (NttPlan::intra_warp: 2 of the 3 forward and 1 of the 3 inverse ones at N = 2048) only need __syncwarp().
// Addressing: `bo[lb]` = NttPlan::tau_boff(tau, lb) (+ any higher-order offset that selects the polynomial) for every
// layout lb, `buf` a byte pointer, `stride` the byte distance between the scratch polynomials of the NB bootstraps.
template <int LOGN, int PASS, int NB, class Sync, bool TWS = false>
__device__ __forceinline__ void ntt_fwd1_from(rns2 (&x)[NB][8], int tau, unsigned char *buf, size_t stride, const u32 (&bo)[LOGN], const fq_tw *psi_rev, Sync sync, bool buf_free, u32 z = 0)
{
    using P = NttPlan<LOGN>;
    ntt_fwd_pass_n<LOGN, PASS, NB, TWS>(x, tau, psi_rev, z);
    if constexpr (PASS + 1 < P::NPASS) {
        constexpr int lb0 = P::fwd_lb(PASS), lb1 = P::fwd_lb(PASS + 1);
        if (!buf_free) sync();                          // earlier readers of buf (other threads) are done
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 o = bo[lb0] ^ P::elem_boff(e, lb0);
#pragma unroll
            for (int b = 0; b < NB; b++) *(u64 *)(buf + b * stride + o) = rns_pack(x[b][e]);
        }
        if constexpr (P::intra_warp(lb0, lb1)) __syncwarp(); else sync();
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 o = bo[lb1] ^ P::elem_boff(e, lb1);
#pragma unroll
            for (int b = 0; b < NB; b++) x[b][e] = rns_unpack(*(const u64 *)(buf + b * stride + o));
        }
        ntt_fwd1_from<LOGN, PASS + 1, NB, Sync, TWS>(x, tau, buf, stride, bo, psi_rev, sync, true, z);
    }
}
// inverse: `after_pass0` runs between pass 0 and the first write to buf
// `psi_rev` is the FORWARD table (read mirrored, see ntt_inv_pass_n)
template <int LOGN, int PASS, int NB, class Sync0, class Sync, bool TWS = false>
__device__ __forceinline__ void ntt_inv1_from(rns2 (&x)[NB][8], int tau, unsigned char *buf, size_t stride, const u32 (&bo)[LOGN], const fq_tw *psi_rev, Sync0 after_pass0, Sync sync, u32 z = 0)
{
    using P = NttPlan<LOGN>;
    ntt_inv_pass_n<LOGN, PASS, NB, true, TWS>(x, tau, psi_rev, z);
    if constexpr (PASS == 0) after_pass0();             // later passes write the words they read themselves: no barrier
    if constexpr (PASS + 1 < P::NPASS) {
        constexpr int lb0 = P::inv_lb(PASS), lb1 = P::inv_lb(PASS + 1);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 o = bo[lb0] ^ P::elem_boff(e, lb0);
#pragma unroll
            for (int b = 0; b < NB; b++) *(u64 *)(buf + b * stride + o) = rns_pack(x[b][e]);
        }
        if constexpr (P::intra_warp(lb0, lb1)) __syncwarp(); else sync();
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 o = bo[lb1] ^ P::elem_boff(e, lb1);
#pragma unroll
            for (int b = 0; b < NB; b++) x[b][e] = rns_unpack(*(const u64 *)(buf + b * stride + o));
        }
        ntt_inv1_from<LOGN, PASS + 1, NB, Sync, Sync, TWS>(x, tau, buf, stride, bo, psi_rev, sync, sync, z);
    }
}

// ---- prime-split drivers: as ntt_fwd1_from / ntt_inv1_from with one bootstrap per thread and 32-bit accesses to the packed scratch
// (`buf` already points at this thread's half of the words: + 4*l); WLOG = log2 of the thread positions per warp (4: a warp
// holds 16 positions x 2 primes)
template <int LOGN, int PASS, int WLOG, class Sync, bool TWS = true>
__device__ __forceinline__ void ntt_fwd1p_from(u32 (&x)[8], int tau, unsigned char *buf, const u32 (&bo)[LOGN], const fq_tw *tw, int l, u32 p, Sync sync, bool buf_free, u32 z = 0)
{
    using P = NttPlan<LOGN>;
    ntt_fwd_pass_1p<LOGN, PASS, TWS>(x, tau, tw, l, p, z);
    if constexpr (PASS + 1 < P::NPASS) {
        constexpr int lb0 = P::fwd_lb(PASS), lb1 = P::fwd_lb(PASS + 1);
        if (!buf_free) sync();
#pragma unroll
        for (int e = 0; e < 8; e++) *(u32 *)(buf + (bo[lb0] ^ P::elem_boff(e, lb0))) = x[e];
        if constexpr (P::intra_warp_w(lb0, lb1, WLOG)) __syncwarp(); else sync();
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = *(const u32 *)(buf + (bo[lb1] ^ P::elem_boff(e, lb1)));
        ntt_fwd1p_from<LOGN, PASS + 1, WLOG, Sync, TWS>(x, tau, buf, bo, tw, l, p, sync, true, z);
    }
}
template <int LOGN, int PASS, int WLOG, class Sync0, class Sync, bool TWS = true>
__device__ __forceinline__ void ntt_inv1p_from(u32 (&x)[8], int tau, unsigned char *buf, const u32 (&bo)[LOGN], const fq_tw *tw, int l, u32 p, Sync0 after_pass0, Sync sync, u32 z = 0)
{
    using P = NttPlan<LOGN>;
    ntt_inv_pass_1p<LOGN, PASS, TWS>(x, tau, tw, l, p, z);
    if constexpr (PASS == 0) after_pass0();
    if constexpr (PASS + 1 < P::NPASS) {
        constexpr int lb0 = P::inv_lb(PASS), lb1 = P::inv_lb(PASS + 1);
#pragma unroll
        for (int e = 0; e < 8; e++) *(u32 *)(buf + (bo[lb0] ^ P::elem_boff(e, lb0))) = x[e];
        if constexpr (P::intra_warp_w(lb0, lb1, WLOG)) __syncwarp(); else sync();
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = *(const u32 *)(buf + (bo[lb1] ^ P::elem_boff(e, lb1)));
        ntt_inv1p_from<LOGN, PASS + 1, WLOG, Sync, Sync, TWS>(x, tau, buf, bo, tw, l, p, sync, sync, z);
    }
}

// inverse NTT passes PASS.. ; the caller supplies x in layout inv_lb(PASS)
template <int LOGN, int PASS, class Sync0, class Sync>
__device__ __forceinline__ void ntt_inv_from(rns2 (&x)[8], int tau, u64 *bufA, u64 *bufB, const fq_tw *psi_inv_rev, Sync0 sync_first, Sync sync)
{
    using P = NttPlan<LOGN>;
    ntt_inv_pass<LOGN, PASS>(x, tau, psi_inv_rev);
    if constexpr (PASS + 1 < P::NPASS) {
        u64 *buf = (PASS & 1) ? bufB : bufA;
        constexpr int lb0 = P::inv_lb(PASS), lb1 = P::inv_lb(PASS + 1);
#pragma unroll
        for (int e = 0; e < 8; e++) buf[P::swz(P::idx(tau, e, lb0))] = rns_pack(x[e]);
        if constexpr (PASS == 0) sync_first(); else sync();
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = rns_unpack(buf[P::swz(P::idx(tau, e, lb1))]);
        ntt_inv_from<LOGN, PASS + 1>(x, tau, bufA, bufB, psi_inv_rev, sync, sync);
    } else if constexpr (PASS == 0) {
        sync_first();
    }
}
#endif
