// kernels.cuh -- sm_100a kernels of the encrypted FBS executor.
//   K1a k_lincomb_decomp : LWE linear combination (weighted sum mod P) + key-switch gadget decomposition
//   K1b k_keyswitch_mma  : digit x KSK-byte GEMM on the tensor cores (mma.sync u8), byte recombination, modulus switch
//   K2  k_blind_rotate   : test polynomial, n CMUX steps (decompose, NTT, BSK MAC, inverse NTT), accumulator in
//                          shared memory, BSK rows streamed with cp.async.bulk (TMA) + mbarrier; K3 sample
//                          extraction (+ table-mode offset) as epilogue
//   K4  encrypt / decrypt / keygen, K5 stand-alone NTT (tests + BSK preprocessing), cleartext evaluator
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "ntt.cuh"

typedef uint8_t u8;
typedef uint16_t u16;

// ------------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u64 *bar, u32 parity)
{
    u32 ok;
#ifdef FBS_WAIT_HINT_NS
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"((u32)FBS_WAIT_HINT_NS) : "memory");
#else
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) { while (!mbar_try_wait(bar, parity)) { } }
// 1-D bulk tensor-memory-accelerator copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// block-wide sum in Z_Q (addition mod Q is associative, so the tree order does not matter)
template <int THREADS>
__device__ __forceinline__ u64 block_sum_gl(u64 v, u64 *sh /* [THREADS/32] */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 w = __shfl_xor_sync(0xffffffffu, v, o);
        v = fq_add(v, w);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    u64 r = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) r = fq_add(r, sh[w]);
    return r;
}

// ------------------------------------------------------------------------------------------------------
// K4: key generation, all seeded and deterministic
// ------------------------------------------------------------------------------------------------------
__global__ void k_gen_bits(u8 *out, int count, u64 seed, u64 dom)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (u8)(fbs_rnd64(seed, dom, (u64)i) & 1);
}
// KSK row r = i*lk + j : LWE_{s_lwe}( s_big[i] * g_j ), body last
__global__ void __launch_bounds__(256) k_gen_ksk(u64 *ksk, int n, int lk, u64 seed, const u8 *__restrict__ s_lwe,
                                                const u8 *__restrict__ s_big, u64 noise_scale, const u64 *__restrict__ gadgets)
{
    __shared__ u64 sh[8];
    const size_t r = blockIdx.x;
    u64 *row = ksk + r * (size_t)(n + 1);
    u64 part = 0;
    for (int q = threadIdx.x; q < n; q += 256) {
        u64 a = fbs_rnd_uniform(seed, DOM_KSK_MASK, r * (u64)n + q);
        row[q] = a;
        if (s_lwe[q]) part = fq_add(part, a);
    }
    u64 body = block_sum_gl<256>(part, sh);
    if (threadIdx.x == 0) {
        body = fq_add(body, fbs_rnd_noise(seed, DOM_KSK_NOISE, r, noise_scale));
        if (s_big[r / lk]) body = fq_add(body, gadgets[r % lk]);
        row[n] = body;
    }
}
__global__ void k_ksk_colsum(const u64 *__restrict__ ksk, int R, int cols, u64 *colsum)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    u64 s = 0;
    for (int r = 0; r < R; r++) s = fq_add(s, ksk[(size_t)r * cols + c]);
    colsum[c] = s;
}
// BSK in coefficient domain: masks uniform, body = noise (the key product is added by k_bsk_body)
__global__ void k_gen_bsk_fill(u64 *bsk, int k, int N, u64 seed, u64 noise_scale, size_t total /* n*rows*(k+1)*N */)
{
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    int q = (int)(g % N);
    size_t poly = g / N;
    int v = (int)(poly % (k + 1));
    size_t ir = poly / (k + 1);
    bsk[g] = (v < k) ? fbs_rnd_uniform(seed, DOM_BSK_MASK, (ir * k + v) * (u64)N + q)
                     : fbs_rnd_noise(seed, DOM_BSK_NOISE, ir * (u64)N + q, noise_scale);
}
// body += sum_v A_v * S_v (negacyclic, S binary), then add s_lwe[i]*g_j to coefficient 0 of poly u
// GGSW i encrypts s_lwe[i] (classic) or, key-unrolled by m = 2 or 3, the indicator bit of subset c = i % (2^m - 1) of key
// group t = i / (2^m - 1): prod_{j in mask} s_j * prod_{j not in mask} (1 - s_j) with zero padding beyond n
// (oracle/tfhe_ref.c: ggsw_bit, mask_of).
__host__ __device__ __forceinline__ int fbs_unroll_mask(int m, int c) { return m == 2 ? (c == 0 ? 3 : c) : c + 1; }
__device__ __forceinline__ int fbs_ggsw_bit(const u8 *__restrict__ s_lwe, int i, int unroll, int n)
{
    if (unroll != 2 && unroll != 3) return s_lwe[i];
    const int ns = (1 << unroll) - 1, t = i / ns, mask = fbs_unroll_mask(unroll, i % ns);
    int bit = 1;
    for (int j = 0; j < unroll; j++) {
        const int kb = (unroll * t + j < n) ? s_lwe[unroll * t + j] : 0;
        bit &= ((mask >> j) & 1) ? kb : (kb ^ 1);
    }
    return bit;
}
__global__ void __launch_bounds__(256) k_bsk_body(u64 *bsk, int k, int N, int l, const u8 *__restrict__ s_lwe,
                                                 const u8 *__restrict__ s_big, const u64 *__restrict__ gadgets, int unroll, int n)
{
    extern __shared__ u64 sh_a[];                 // [N] mask poly, then [N] key bits as bytes
    u8 *sh_s = (u8 *)(sh_a + N);
    const int ir = blockIdx.x, rows = (k + 1) * l;
    const int i = ir / rows, r = ir % rows, u = r / l, j = r % l;
    u64 *row = bsk + (size_t)ir * (k + 1) * N;
    for (int v = 0; v < k; v++) {
        __syncthreads();
        for (int q = threadIdx.x; q < N; q += 256) { sh_a[q] = row[(size_t)v * N + q]; sh_s[q] = s_big[v * N + q]; }
        __syncthreads();
        for (int c = threadIdx.x; c < N; c += 256) {
            u64 acc = row[(size_t)k * N + c];
            for (int jj = 0; jj < N; jj++) {
                if (!sh_s[jj]) continue;
                int idx = c - jj;
                acc = (idx >= 0) ? fq_add(acc, sh_a[idx]) : fq_sub(acc, sh_a[idx + N]);
            }
            row[(size_t)k * N + c] = acc;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && fbs_ggsw_bit(s_lwe, i, unroll, n)) row[(size_t)u * N] = fq_add(row[(size_t)u * N], gadgets[j]);
}

// ------------------------------------------------------------------------------------------------------
// K5: stand-alone NTT.  mode 0: forward (natural in, bit-reversed out, array order as oracle/tfhe_ref.c)
//                       mode 1: inverse incl. 1/N (for tests)
//                       mode 2: BSK preprocessing: forward, residues times 2^32/N (Montgomery form with the inverse
//                               transform's 1/N folded in), packed (mod p1 | mod p2 << 32), stored SWIZZLED for the
//                               blind-rotate kernel
//                       mode 3: key-unrolled BSK preprocessing: as mode 2 with the caller's scale 2^64/N (Montgomery form
//                               twice: the key goes through two reductions), stored as [key group][element e][tau >> 5][c][u][v][tau & 31]:
//                               the words k_blind_rotate2 needs for one element are one contiguous slice (one bulk copy), and so are
//                               the words of any contiguous range of 32-thread blocks (k_blind_rotate_cl: one CTA of a cluster)
// ------------------------------------------------------------------------------------------------------
template <int LOGN>
__global__ void __launch_bounds__(NttPlan<LOGN>::T) k_ntt(const u64 *__restrict__ in, u64 *__restrict__ out, int mode,
                                                         const fq_tw *__restrict__ psi_rev, const fq_tw *__restrict__ psi_inv_rev, u32 scale1, u32 scale2, int g1 /* mode 3: (k+1) | GGSW-per-group << 8 */)
{
    using P = NttPlan<LOGN>;
    __shared__ u64 bufA[P::N];
    __shared__ u64 bufB[P::N];
    const int tau = threadIdx.x;
    const u64 *src = in + (size_t)blockIdx.x * P::N;
    u64 *dst = out + (size_t)blockIdx.x * P::N;
    rns2 x[8];
    auto sync = [] { __syncthreads(); };
    if (mode == 1) {                                   // inverse, scale = 1/N per prime; integers in, integers out
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = rns_from_int(src[P::idx(tau, e, P::inv_lb(0))]);
        ntt_inv_from<LOGN, 0>(x, tau, bufA, bufB, psi_inv_rev, sync, sync);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            rns2 v;
            v.a = (u32)((u64)r32_csub(x[e].a, FQ_P1) * scale1 % FQ_P1);
            v.b = (u32)((u64)r32_csub(x[e].b, FQ_P2) * scale2 % FQ_P2);
            dst[P::idx(tau, e, P::inv_lb(P::NPASS - 1))] = rns_to_int(v);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = rns_from_int(src[P::idx(tau, e, P::fwd_lb(0))]);
        ntt_forward<LOGN>(x, tau, bufA, bufB, psi_rev, sync);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int id = P::idx(tau, e, P::fwd_lb(P::NPASS - 1));
            rns2 v;                                                      // lazy [0,4p) -> canonical
            v.a = r32_csub(r32_fold(x[e].a, 2 * FQ_P1), FQ_P1);
            v.b = r32_csub(r32_fold(x[e].b, 2 * FQ_P2), FQ_P2);
            if (mode == 0) {
                dst[id] = rns_to_int(v);                                 // integer spectrum, as the oracle computes it
            } else {                                                     // mode 2: residues * 2^32/N, packed, swizzled
                v.a = (u32)((u64)v.a * scale1 % FQ_P1);
                v.b = (u32)((u64)v.b * scale2 % FQ_P2);
                if (mode == 3) {                                          // polynomial (ggsw = 3t + c, u, v); id = 8*tau + e
                    const int G1 = g1 & 0xFF, ns = g1 >> 8;              // k+1 and GGSW per key group (3 or 7), packed by the host
                    const int pv = blockIdx.x % (G1 * G1), ggsw = blockIdx.x / (G1 * G1), t = ggsw / ns, c = ggsw % ns;
                    // [key group t][element e][block of 32 thread positions tau >> 5][c][pv][tau & 31]: everything a CTA -- or one CTA of a
                    // cluster that owns a contiguous range of tau blocks -- needs for element e is ONE contiguous run (one bulk copy)
                    out[((((size_t)t * 8 + e) * (P::T / 32) + (tau >> 5)) * ns * G1 * G1 + (size_t)c * G1 * G1 + pv) * 32 + (tau & 31)] = rns_pack(v);
                }
                else dst[P::swz(id)] = rns_pack(v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K4: encrypt / decrypt
// ------------------------------------------------------------------------------------------------------
struct EncArgs {
    const u8 *msgs;          // program inputs: [n_inputs][B] ; explicit: [count]
    const int32_t *msgs32;   // explicit int32 messages (debug), used when != nullptr
    const u64 *ct_ids;       // explicit ids or nullptr
    const int32_t *in_slot;  // program inputs -> slot, or nullptr (explicit: ciphertext e goes to out + e*CT)
    u64 *out;
    const u8 *s_big;
    long long B, inst_offset, B_total;
    int D, p;
    u64 enc_seed, noise_scale;
};
__global__ void __launch_bounds__(256) k_encrypt(EncArgs a)
{
    __shared__ u64 sh[8];
    const long long e = blockIdx.x;
    long long m;
    u64 id;
    u64 *ct;
    if (a.in_slot) {
        const long long i = e / a.B, b = e % a.B;
        m = a.msgs[e];
        id = (u64)i * (u64)a.B_total + (u64)(a.inst_offset + b);
        ct = a.out + ((size_t)a.in_slot[i] * a.B + b) * (size_t)(a.D + 1);
    } else {
        m = a.msgs32 ? a.msgs32[e] : a.msgs[e];
        id = a.ct_ids ? a.ct_ids[e] : (u64)e;
        ct = a.out + (size_t)e * (a.D + 1);
    }
    u64 part = 0;
    for (int q = threadIdx.x; q < a.D; q += 256) {
        u64 x = fbs_rnd_uniform(a.enc_seed, DOM_ENC_MASK, id * (u64)a.D + q);
        ct[q] = x;
        if (a.s_big[q]) part = fq_add(part, x);
    }
    u64 body = block_sum_gl<256>(part, sh);
    if (threadIdx.x == 0) {
        body = fq_add(body, fbs_rnd_noise(a.enc_seed, DOM_ENC_NOISE, id, a.noise_scale));
        const int p2 = 2 * a.p;
        int mm = (int)(((m % p2) + p2) % p2);
        ct[a.D] = fq_add(body, fq_mul((u64)mm, fbs_delta(a.p)));
    }
}
// outputs are lincombs of wires: phase(sum c_o ct_o) + const*Delta, decoded to Z_2p
struct OutArgs {
    const u64 *wires; const u8 *s_big;
    const int32_t *out_ptr, *out_slot, *out_coef, *out_const;   // nullptr out_ptr: plain decrypt of cts [count]
    u8 *out8; int32_t *out32;
    long long B; int D, p;
};
__global__ void __launch_bounds__(256) k_decrypt(OutArgs a)
{
    __shared__ u64 sh[8];
    const long long e = blockIdx.x;
    const size_t CT = (size_t)a.D + 1;
    u64 part = 0, body = 0;
    if (a.out_ptr) {
        const long long q = e / a.B, b = e % a.B;
        const int o0 = a.out_ptr[q], o1 = a.out_ptr[q + 1];
        for (int o = o0; o < o1; o++) {
            const u64 *ct = a.wires + ((size_t)a.out_slot[o] * a.B + b) * CT;
            const u64 cf = fq_from_i64(a.out_coef[o]);
            u64 acc = 0;
            for (int w = threadIdx.x; w < a.D; w += 256) if (a.s_big[w]) acc = fq_add(acc, ct[w]);
            part = fq_add(part, fq_mul(acc, cf));
            body = fq_add(body, fq_mul(ct[a.D], cf));
        }
        body = fq_add(body, fq_mul(fq_from_i64(a.out_const[q]), fbs_delta(a.p)));
    } else {
        const u64 *ct = a.wires + (size_t)e * CT;
        for (int w = threadIdx.x; w < a.D; w += 256) if (a.s_big[w]) part = fq_add(part, ct[w]);
        body = ct[a.D];
    }
    u64 mask = block_sum_gl<256>(part, sh);
    if (threadIdx.x == 0) {
        int m = fbs_decode(fq_sub(body, mask), a.p);
        if (a.out8) a.out8[e] = (u8)m;
        if (a.out32) a.out32[e] = m;
    }
}

// ------------------------------------------------------------------------------------------------------
// K1a: linear combination + key-switch decomposition.  One CTA = 16 (lincomb, instance) pairs x all kN+1 words.
// Digits go out row-major, digits[m][r = i*lk + j] (uint8, offset form d + B/2 in [0, B)): the A operand of the
// tensor-core key switch.
// ------------------------------------------------------------------------------------------------------
struct LCArgs {
    const u64 *wires;
    const int32_t *lc_ptr, *lc_slot, *lc_coef, *lc_const;
    u8 *digits; u64 *body;
    long long B, M;          // M = (#lincombs in range) * B
    int lc_begin, D, p, ks_beta;
};
template <int LK>
__global__ void __launch_bounds__(256) k_lincomb_decomp(LCArgs a)
{
    __shared__ int s_lc[16];
    __shared__ long long s_inst[16];
    const long long tile = blockIdx.x;
    if (threadIdx.x < 16) {
        long long m = tile * 16 + threadIdx.x;
        s_lc[threadIdx.x] = (m < a.M) ? a.lc_begin + (int)(m / a.B) : -1;
        s_inst[threadIdx.x] = (m < a.M) ? (m % a.B) : 0;
    }
    __syncthreads();
    const size_t CT = (size_t)a.D + 1;
    const size_t R = (size_t)a.D * LK;
    const int halfB = 1 << (a.ks_beta - 1);
    // gridDim.y column chunks: a level of a narrow circuit has only a few tiles of 16 ciphertexts, too few to fill the SMs
    const int chunk = ((a.D + 1 + (int)gridDim.y - 1) / (int)gridDim.y + 255) / 256 * 256, i_end = min(a.D + 1, ((int)blockIdx.y + 1) * chunk);
    for (int i = blockIdx.y * chunk + threadIdx.x; i < i_end; i += 256) {
#pragma unroll
        for (int mm = 0; mm < 16; mm++) {
            const int lc = s_lc[mm];
            if (lc < 0) continue;
            const long long inst = s_inst[mm];
            // sum_o c_o x_o with small integer coefficients: accumulate the exact 128-bit integer (|c| < 2^31, x < 2^60, <= 64 operands:
            // < 2^97), a negative coefficient multiplies q - x; ONE reduction mod q per word instead of one modular multiply per operand
            u64 lo = 0, hi = 0;
            for (int o = a.lc_ptr[lc]; o < a.lc_ptr[lc + 1]; o++) {
                u64 x = __ldg(a.wires + ((size_t)a.lc_slot[o] * a.B + inst) * CT + i);
                const int cf = a.lc_coef[o];
                if (cf < 0) x = fq_neg(x);
                const u64 m = (u64)(cf < 0 ? -(long long)cf : (long long)cf);
                const u64 pl = x * m, ph = __umul64hi(x, m);
                lo += pl; hi += ph + (lo < pl ? 1 : 0);
            }
            u64 acc = fq_reduce128(lo, hi);
            if (i == a.D) {
                acc = fq_add(acc, fq_mul(fq_from_i64(a.lc_const[lc]), fbs_delta(a.p)));
                a.body[tile * 16 + mm] = acc;
            } else {
                int d[LK];
                fbs_balanced_digits<LK>(fbs_round_top(acc, a.ks_beta * LK), a.ks_beta, d);
                u8 *dst = a.digits + (size_t)(tile * 16 + mm) * R + (size_t)i * LK;
#pragma unroll
                for (int j = 0; j < LK; j++) dst[j] = (u8)(d[j] + halfB);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K1b: key switch on the tensor cores.  out[m][c] = [c==n]*body[m] - sum_r d[m][r] * KSK[r][c]  (mod q), then
// modulus switch to 2N.  The KSK entry is split into its 8 bytes: C[m][(c,b)] = sum_r du[m][r] * byte_b(KSK[r][c]) is an
// exact u8 x u8 -> s32 GEMM (7 * 255 * kN*l_ks < 2^31), run as mma.sync.m16n8k32; the 8 byte sums of a column are
// recombined as sum_b C_b * 2^(8b) mod q in the epilogue and the digit offset B/2 is removed with (B/2)*colsum[c].
// This is the one GEMM-shaped step of the path: 8x more MACs than the integer-pipe version
// (2 IMAD.WIDE per MAC), but on a unit that is otherwise idle -- measured 18.3 -> 1.3 ms per 4736 ciphertexts (profiles/README.md).
// KbT[(c*8+b)][r] is the byte-transposed key (k-contiguous "col" operand), CTA tile 64 ciphertexts x 8 columns x 128 rows,
// 4-stage cp.async pipeline, 144-byte padded rows (bank-conflict-free 32-bit fragment loads).
// ------------------------------------------------------------------------------------------------------
struct KSArgs {
    const u8 *digits; const u64 *body; const u8 *kbt; const u64 *colsum;
    u16 *ms; u64 *tap_ks;
    long long M; int R, n, ks_beta, log2_2N;
};
#define KS_BM 64
#define KS_BN 64
#define KS_BK 128
#define KS_LD 144
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void mma_u8(int (&c)[4], u32 a0, u32 a1, u32 a2, u32 a3, u32 b0, u32 b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
#ifndef KS_STAGES
#define KS_STAGES 4      /* cp.async pipeline depth: with 2 a narrow level's key switch is one L2 round trip per 128 key rows (80 in a row) */
#endif
#define KS_SMEM_BYTES ((size_t)KS_STAGES * (KS_BM + KS_BN) * KS_LD)
__global__ void __launch_bounds__(256) k_keyswitch_mma(KSArgs a)
{
    extern __shared__ __align__(16) unsigned char ks_smem[];
    u8 (*sA)[KS_BM * KS_LD] = (u8 (*)[KS_BM * KS_LD])ks_smem;
    u8 (*sB)[KS_BN * KS_LD] = (u8 (*)[KS_BN * KS_LD])(ks_smem + (size_t)KS_STAGES * KS_BM * KS_LD);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const long long m0 = (long long)blockIdx.x * KS_BM;
    const int n0 = blockIdx.y * KS_BN;
    const size_t R = (size_t)a.R;
    const u8 *gA = a.digits + (size_t)m0 * R, *gB = a.kbt + (size_t)n0 * R;
    auto load_stage = [&](int st, int k0) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int id = tid + 256 * h, row = id >> 3, c16 = (id & 7) * 16;
            cp_async16(&sA[st][row * KS_LD + c16], gA + (size_t)row * R + k0 + c16);
            cp_async16(&sB[st][row * KS_LD + c16], gB + (size_t)row * R + k0 + c16);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0; }
    const int nk = a.R / KS_BK;
    // KS_STAGES - 1 stages in flight; every iteration commits exactly one group (an empty one past the end), so that
    // wait_group KS_STAGES - 2 always means "stage ks has landed"
    for (int st = 0; st < KS_STAGES - 1; st++) {
        if (st < nk) load_stage(st, st * KS_BK); else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int ks = 0; ks < nk; ks++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(KS_STAGES - 2) : "memory");
        __syncthreads();                                 // stage ks visible to all; everybody is done with stage ks - 1 (refilled next)
        if (ks + KS_STAGES - 1 < nk) load_stage((ks + KS_STAGES - 1) % KS_STAGES, (ks + KS_STAGES - 1) * KS_BK);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const u8 *A = sA[ks % KS_STAGES] + (wm * 16 + g) * KS_LD + t * 4;
        const u8 *Bp = sB[ks % KS_STAGES] + (wn * 32 + g) * KS_LD + t * 4;
#pragma unroll
        for (int kk = 0; kk < KS_BK / 32; kk++) {
            const u32 a0 = *(const u32 *)(A + kk * 32), a1 = *(const u32 *)(A + 8 * KS_LD + kk * 32);
            const u32 a2 = *(const u32 *)(A + kk * 32 + 16), a3 = *(const u32 *)(A + 8 * KS_LD + kk * 32 + 16);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const u32 b0 = *(const u32 *)(Bp + j * 8 * KS_LD + kk * 32), b1 = *(const u32 *)(Bp + j * 8 * KS_LD + kk * 32 + 16);
                mma_u8(acc[j], a0, a1, a2, a3, b0, b1);
            }
        }
    }
    // ---- epilogue: recombine the 8 byte sums of each column, remove the digit offset, add the body, modulus switch
    const int cols = a.n + 1;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int c = (n0 >> 3) + wn * 4 + j;
        const bool col_ok = c < cols;
        const u64 corr = col_ok ? fq_mul(a.colsum[c], (u64)(1u << (a.ks_beta - 1))) : 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const u64 v = (u64)(u32)acc[j][2 * h] + ((u64)(u32)acc[j][2 * h + 1] << 8);      // bytes 2t, 2t+1
            const int sh = 16 * t;
            u64 x = fq_reduce128(v << sh, sh ? (v >> (64 - sh)) : 0);
            x = fq_add(x, __shfl_xor_sync(0xffffffffu, x, 1));
            x = fq_add(x, __shfl_xor_sync(0xffffffffu, x, 2));
            const long long m = m0 + wm * 16 + g + 8 * h;
            if (t == 0 && col_ok && m < a.M) {
                u64 val = fq_sub(corr, x);
                if (c == a.n) val = fq_add(val, a.body[m]);
                a.ms[(size_t)m * cols + c] = (u16)fbs_modswitch(val, a.log2_2N);
                if (a.tap_ks) a.tap_ks[(size_t)m * cols + c] = val;
            }
        }
    }
}
// byte-transpose the key-switching key for the tensor-core path: kbt[(c*8+b)][r] = byte b of ksk[r][c]
__global__ void k_ksk_bytes_t(const u64 *__restrict__ ksk, u8 *__restrict__ kbt, int R, int cols, int cols_pad)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= R) return;
    const u64 x = (c < cols) ? ksk[(size_t)r * cols + c] : 0;
#pragma unroll
    for (int b = 0; b < 8; b++) kbt[((size_t)c * 8 + b) * R + r] = (u8)(x >> (8 * b));
}

// ------------------------------------------------------------------------------------------------------
// K2 + K3: blind rotation + sample extraction.  One CTA runs PB bootstraps in lock-step (they share the BSK row in
// shared memory); each bootstrap has (K+1) thread groups of N/8 threads, group g owns accumulator polynomial g: it
// decomposes / forward-transforms input polynomial g and produces / inverse-transforms output polynomial g.
// Shared memory (u64 words):
//   per bootstrap: ACC[(K+1)][N] natural order | S[(K+1)][N] transpose scratch (swizzled) | DH[(K+1)L][N] digit spectra
//                  (aliases S when L == 1)
//   per CTA      : BS[(K+1)L][(K+1)][N] current BSK row (TMA target, when BSK_SMEM)
// PB = 2 doubles the resident warps per SM (8 per scheduler at 64 registers per thread): the first version of this
// kernel ran 4 warps per scheduler and was issue-latency bound (profiles/r1_v2_*).
// ------------------------------------------------------------------------------------------------------
struct BRArgs {
    const u16 *ms;                      // [(lincomb - lc_begin) * B + inst][n+1]
    const u64 *bsk;
    const fq_tw *psi_rev, *psi_inv_rev;
    const u64 *psi_pow;                 // [2N] packed (psi^x - 1) mod (p1 | p2 << 32): evaluations of X^x - 1 (key-unrolled kernel)
    const int32_t *bs_lc, *bs_slot, *bs_tab_ptr, *bs_mode;
    const u8 *bs_tab;
    u64 *wires; u64 *tap_acc;
    long long B, jobs, job_begin;       // this launch covers jobs [job_begin, jobs) of the level
    int node_begin, lc_begin, n, p, beta;
    u32 zero;                           // always 0; unknown to ptxas, see ntt.cuh (pins adds to the ALU pipe)
    int n_peers;                        // node-sharded multi-GPU: replicas of the wire buffer on the other GPUs
    u64 *peer_wires[8];                 // (peer-mapped pointers, same layout): the sample-extract epilogue stores to all
    // multi-value bootstrap: jobs are (group, instance); the rotation starts from the table-independent base
    // polynomial and leaves the accumulator in tap_acc for k_multi_extract (no sample extraction here)
    const int32_t *grp_first;           // nullptr: ordinary bootstrap, node = node_begin + job / B
};
// H = Delta / 2 mod q (q is odd): coefficient of the base test polynomial and unit of the table-mode offset of a multi-value bootstrap
__device__ __forceinline__ u64 fbs_half_delta(int p) { return fq_mul(fbs_delta(p), (FQ_Q + 1) / 2); }
// TP = bootstraps carried by each thread (1, or PB: every thread works on all PB bootstraps of the CTA, so twiddle loads,
// BSK reads, index arithmetic and barriers are shared between them and the instruction-level parallelism doubles).
template <int LOGN, int K, int L, bool BSK_SMEM, int PB, int TP>
struct BRCfg {
    static_assert(TP == 1 || TP == PB, "a thread carries one bootstrap or all of the CTA's");
    static constexpr int N = 1 << LOGN, G = K + 1, T = N / 8, PT = G * T, THREADS = (PB / TP) * PT;
    static constexpr size_t acc_w = (size_t)G * N, s_w = (size_t)G * N;
    static constexpr size_t dh_w = (L == 1) ? 0 : (size_t)G * L * N;
    static constexpr size_t per_pbs_w = acc_w + s_w + dh_w;
    static constexpr size_t bs_w = BSK_SMEM ? (size_t)G * L * G * N : 0;
    static constexpr size_t row_w = (size_t)G * L * G * N;       // BSK words per CMUX step
    // Position of key polynomial (row r = rg*L + jj, column v) of a GGSW row in SHARED memory: [(rg - v) mod G][jj][v].  With this
    // rotation the keys thread group v multiplies its own spectra with sit at a compile-time distance from the word the
    // thread addresses in its own scratch polynomial (the TMA copy permutes; the layout in HBM stays [r][v]).
    __host__ __device__ static constexpr int key_slot(int r, int v) { return ((((r / L) - v + G) % G) * L + (r % L)) * G + v; }
    __host__ __device__ static constexpr size_t ms_stride(int n) { return (((size_t)(n + 1) * 2 + 15) / 16) * 16; }
    __host__ __device__ static constexpr size_t smem_bytes(int n)
    {
        // keep set A at 2 bootstraps/CTA under 195 KiB so the SM stays in the 196 KiB carve-out and L1 keeps ~60 KB for twiddles
        return 8 * (PB * per_pbs_w + bs_w) + 16 /* mbarrier */ + PB * ((((size_t)(n + 1) * 2 + 15) / 16) * 16);
    }
};

template <int LOGN, int K, int L, bool BSK_SMEM, int PB, int TP>
__global__ void __launch_bounds__((PB / TP) * (K + 1) * (1 << LOGN) / 8, 1) k_blind_rotate(BRArgs a)
{
    using C = BRCfg<LOGN, K, L, BSK_SMEM, PB, TP>;
    using P = NttPlan<LOGN>;
    constexpr int N = C::N, G = C::G, T = C::T;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, pb0 = (TP == 1) ? tid / C::PT : 0, ptid = tid % C::PT, g = ptid / T, tau = ptid % T;
    u64 *base = (u64 *)smem_raw + (size_t)pb0 * C::per_pbs_w;   // bootstrap q of this thread lives at base + q*per_pbs_w
    u64 *ACC = base;
    u64 *S = ACC + C::acc_w;
    u64 *DH = (L == 1) ? S : S + C::s_w;
    u64 *BS = (u64 *)smem_raw + (size_t)PB * C::per_pbs_w;
    u64 *mbar = BS + C::bs_w;
    const size_t ms_stride = C::ms_stride(a.n);
    u16 *s_ms = (u16 *)((unsigned char *)(mbar + 2) + (size_t)pb0 * ms_stride);
    constexpr size_t PW = C::per_pbs_w;
    const int n = a.n, p = a.p;

    // a bootstrap slot beyond the job list recomputes the last job and skips the store
    bool live[TP]; int node[TP], tabL[TP], tab0[TP], mode[TP]; long long inst[TP], job[TP];
#pragma unroll
    for (int q = 0; q < TP; q++) {
        job[q] = a.job_begin + (long long)blockIdx.x * PB + pb0 + q;
        live[q] = job[q] < a.jobs;
        if (!live[q]) job[q] = a.jobs - 1;
        node[q] = a.node_begin + (int)(job[q] / a.B);
        if (a.grp_first) node[q] = a.grp_first[node[q]];           // multi-value: first bootstrap of the group (they share the lincomb)
        inst[q] = job[q] % a.B;
        const u16 *ms = a.ms + ((size_t)(a.bs_lc[node[q]] - a.lc_begin) * a.B + inst[q]) * (size_t)(n + 1);
        tab0[q] = a.bs_tab_ptr[node[q]]; tabL[q] = a.bs_tab_ptr[node[q] + 1] - tab0[q]; mode[q] = a.bs_mode[node[q]];
        u16 *dst = (u16 *)((unsigned char *)s_ms + (size_t)q * ms_stride);
        for (int i = ptid; i <= n; i += C::PT) dst[i] = ms[i];
    }
    if (BSK_SMEM && tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (BSK_SMEM && tid == 0) {
        fence_proxy_async();
        mbar_expect_tx(mbar, (u32)(C::row_w * 8));
#pragma unroll 1
        for (int q = 0; q < G * L * G; q++) tma_load_1d(BS + (size_t)C::key_slot(q / G, q % G) * N, a.bsk + (size_t)q * N, N * 8, mbar);
    }
    // ---- accumulator init: ACC = (0, .., 0, X^{-b~} * TV) ; TV[j] = F(round(j*p/N)), F(x) = tv[x]*Delta - s*Delta/2
    // The accumulator lives in shared memory as packed residue pairs (mod p1 | mod p2 << 32), canonical.
    u64 *acc = ACC + (size_t)g * N;
#pragma unroll
    for (int q = 0; q < TP; q++) {
        const u64 delta = fbs_delta(p), off = fq_mul((u64)mode[q], delta >> 1);
        const int bt = ((const u16 *)((const unsigned char *)s_ms + (size_t)q * ms_stride))[n];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int j = tau + e * T;
            u64 val = 0;
            if (g == K) {
                int src = j + bt;                       // (X^{-bt} TV)[j] = +-TV[(j + bt) mod N]
                bool neg = false;
                if (src >= 2 * N) src -= 2 * N;
                if (src >= N) { src -= N; neg = true; }
                int x = (int)((2LL * src * p + N) / (2LL * N));
                if (x >= p) { x -= p; neg = neg != (a.grp_first == nullptr); }     // half slot: the table's step function is negated there, the multi-value base polynomial is not
                const u64 tvx = (x < tabL[q]) ? (u64)__ldg(a.bs_tab + tab0[q] + x) : 0;
                const u64 F = a.grp_first ? fbs_half_delta(p) : fq_sub(fq_mul(tvx, delta), off);
                val = neg ? fq_neg(F) : F;
            }
            acc[q * PW + j] = rns_pack(rns_from_int(val));
        }
    }
    __syncthreads();

    const int bar_g = 1 + pb0 * G + g, bar_p = 1 + PB * G + pb0;    // named barriers: per group, per bootstrap
    auto gsync = [bar_g] { bar_sync_named(bar_g, T); };
    auto psync = [bar_p] { if (TP == 1) bar_sync_named(bar_p, C::PT); else __syncthreads(); };
    const int beta = a.beta, bits = beta * L;
    constexpr bool one_digit = (L == 1);                           // fbs_digit1_t applies (beta <= 24 checked by fbs_ctx_create)
    const u64 rc = 1ULL << (62 - (one_digit ? beta : 24));
    static_assert(P::idx(1, 1, P::inv_lb(P::NPASS - 1)) == 1 + T && P::fwd_lb(0) == P::inv_lb(P::NPASS - 1),
                  "the inverse transform leaves coefficient tau + e*T in register e, the layout the rotation reads");
    // Shared-memory addressing (ntt.cuh): byte offset of a word = (thread part, one register per layout, including the
    // group's polynomial select g << SH) XOR (compile-time element part).
    constexpr int SH = LOGN + 3;                                   // log2 of a polynomial's size in bytes
    constexpr size_t PWB = PW * 8;
    unsigned char *Sb = (unsigned char *)S, *DHb = (unsigned char *)DH, *BSb = (unsigned char *)BS;
    u32 bo[LOGN];
#pragma unroll
    for (int lb = 0; lb < LOGN; lb++) bo[lb] = P::tau_boff(tau, lb) | ((u32)g << SH);
    // this thread's own accumulator coefficients j = tau + e*T stay mirrored in registers between the steps
    rns2 av[TP][8];
#pragma unroll
    for (int q = 0; q < TP; q++)
#pragma unroll
        for (int e = 0; e < 8; e++) av[q][e] = rns_unpack(acc[q * PW + tau + e * T]);

    for (int i = 0; i < n; i++) {
        // ---- rotate, subtract, decompose: digits of (X^{ai} ACC_g - ACC_g) in the first forward layout.
        // Only the high mixed-radix digit t = floor(x / p1) of the difference is needed for the rounding (<= 24 bits).
        rns2 dg[L][TP][8];
#pragma unroll
        for (int q = 0; q < TP; q++) {
            const int ai = ((const u16 *)((const unsigned char *)s_ms + (size_t)q * ms_stride))[i];
            const unsigned char *ac = (const unsigned char *)(acc + q * PW);
            const int base8 = (tau - ai) * 8;               // source of coefficient j = tau + e*T is (tau - ai + e*T) mod 2N
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int src8 = base8 + e * T * 8;
                const rns2 rot = rns_unpack(*(const u64 *)(ac + (src8 & ((N - 1) * 8))));
                const bool neg = (src8 & (N * 8)) != 0;     // X^N = -1
                // +-rot - acc, made canonical with unsigned minima: r in [0,p], r - acc + p in [1,2p]
                const u32 ra = neg ? FQ_P1 - rot.a : rot.a, rb = neg ? FQ_P2 - rot.b : rot.b;
                rns2 diff;
                diff.a = r32_csub(r32_csub(ra - av[q][e].a + a.zero + FQ_P1, FQ_P1), FQ_P1);
                diff.b = r32_csub(rb - av[q][e].b + a.zero + FQ_P2, FQ_P2);        // in [0,p2]: enough for the CRT digit
                const u32 t = rns_crt_hi(diff);
                if constexpr (one_digit) {
                    const u32 d = (u32)fbs_digit1_t(t, diff.a, beta, rc);
                    dg[0][q][e].a = d + FQ_P1;              // lazy residues in (0, 2p) of a digit in [-B/2, B/2)
                    dg[0][q][e].b = d + FQ_P2;
                } else {
                    const u64 y = (bits <= 24) ? fbs_round_top_t(t, diff.a, bits) : fbs_round_top((u64)diff.a + (u64)FQ_P1 * t, bits);
                    int d[L];
                    fbs_balanced_digits<L>(y, beta, d);
#pragma unroll
                    for (int jj = 0; jj < L; jj++) dg[jj][q][e] = rns_from_small(d[jj]);
                }
            }
        }
        // ---- forward NTTs through the group's scratch polynomial; spectra to DH (swizzled, layout lb = 0).  With L == 1
        // DH aliases S and the spectrum goes to the words this thread read in the last transpose: no barrier in between.
#pragma unroll
        for (int jj = 0; jj < L; jj++) {
            ntt_fwd1_from<LOGN, 0, TP>(dg[jj], tau, Sb, PWB, bo, a.psi_rev, gsync, jj == 0, a.zero);
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const u32 o = bo[0] ^ P::elem_boff(e, 0);
#pragma unroll
                for (int q = 0; q < TP; q++) {
                    if constexpr (L == 1) *(u64 *)(Sb + q * PWB + o) = rns_pack(dg[jj][q][e]);
                    else *(u64 *)(DHb + q * PWB + ((size_t)(g * L + jj) << SH) + (o & (N * 8 - 1))) = rns_pack(dg[jj][q][e]);
                }
            }
        }
        psync();
        // ---- pointwise: out_g = sum_r DH[r] * BSK[r][g], per prime.  DH is lazy (< 4p), the key is canonical and in
        // Montgomery form: a PAIR of 64-bit products (< 8p^2 < 2^63) is reduced by one REDC to < 3p, then folded to < 2p.
        // The key word is read once and used for every bootstrap the thread carries.  Term order: the thread's own L
        // spectra first (still in registers), then the other groups' from shared memory.
        if (BSK_SMEM) mbar_wait(mbar, (u32)(i & 1));
        rns2 x[TP][8];
        {
            const unsigned char *browb = (const unsigned char *)(a.bsk + (size_t)i * C::row_w);     // !BSK_SMEM: row in L2
            auto key = [&](int tt, u32 o) -> rns2 {          // key of term tt = og*L + jj for this thread's column g
                const int og = tt / L, jj = tt % L;
                if constexpr (BSK_SMEM) return rns_unpack(*(const u64 *)(BSb + (((size_t)(og * L + jj) * G) << SH) + o));
                int gg = g + og; if (gg >= G) gg -= G;
                return rns_unpack(*(const u64 *)(browb + (((size_t)(gg * L + jj) * G + g) << SH) + (o & (N * 8 - 1))));
            };
            auto digit = [&](int tt, int q, int e, u32 o) -> rns2 {
                const int og = tt / L, jj = tt % L;
                if (og == 0) return dg[jj < L ? jj : 0][q][e];
                int gg = g + og; if (gg >= G) gg -= G;
                if constexpr (L == 1) {
                    const u32 xg = (G == 2) ? (1u << SH) : ((u32)(g ^ gg) << SH);     // o selects polynomial g: flip to gg
                    return rns_unpack(*(const u64 *)(Sb + q * PWB + (o ^ xg)));
                }
                return rns_unpack(*(const u64 *)(DHb + q * PWB + ((size_t)(gg * L + jj) << SH) + (o & (N * 8 - 1))));
            };
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const u32 o = bo[0] ^ P::elem_boff(e, 0);
#pragma unroll
                for (int t0 = 0; t0 < G * L; t0 += 2) {
                    const rns2 k0 = key(t0, o);
                    rns2 k1; k1.a = 0; k1.b = 0;
                    if (t0 + 1 < G * L) k1 = key(t0 + 1, o);
#pragma unroll
                    for (int q = 0; q < TP; q++) {
                        const rns2 d0 = digit(t0, q, e, o);
                        u64 pa = r32_mulwide(d0.a, k0.a), pb2 = r32_mulwide(d0.b, k0.b);
                        if (t0 + 1 < G * L) {
                            const rns2 d1 = digit(t0 + 1, q, e, o);
                            pa = r32_madwide(d1.a, k1.a, pa);
                            pb2 = r32_madwide(d1.b, k1.b, pb2);
                        }
                        const u32 ta = r32_fold(r32_redc(pa, FQ_P1, FQ_P1_INVNEG, a.zero), 2 * FQ_P1);
                        const u32 tb = r32_fold(r32_redc(pb2, FQ_P2, FQ_P2_INVNEG, a.zero), 2 * FQ_P2);
                        x[q][e].a = (t0 == 0) ? ta : r32_fold(x[q][e].a + ta, 2 * FQ_P1);
                        x[q][e].b = (t0 == 0) ? tb : r32_fold(x[q][e].b + tb, 2 * FQ_P2);
                    }
                }
            }
        }
        // ---- inverse NTT.  After its first register pass a CTA-wide barrier guarantees that nobody reads DH / BS of
        // this step any more: the scratch may be overwritten and the next GGSW row is prefetched by TMA.
        auto after_pass0 = [&] {
            __syncthreads();
            if (BSK_SMEM && tid == 0 && i + 1 < n) {
                fence_proxy_async();
                mbar_expect_tx(mbar, (u32)(C::row_w * 8));
                const u64 *src = a.bsk + (size_t)(i + 1) * C::row_w;
#pragma unroll 1
                for (int q = 0; q < G * L * G; q++) tma_load_1d(BS + (size_t)C::key_slot(q / G, q % G) * N, src + (size_t)q * N, N * 8, mbar);
            }
        };
        ntt_inv1_from<LOGN, 0, TP>(x, tau, Sb, PWB, bo, a.psi_rev, after_pass0, gsync, a.zero);
        // ---- accumulate: ACC_g += out_g (inverse output < 2p), canonical; coefficient tau + e*T sits in register e
#pragma unroll
        for (int e = 0; e < 8; e++) {
#pragma unroll
            for (int q = 0; q < TP; q++) {
                av[q][e].a = r32_csub(r32_fold(av[q][e].a + x[q][e].a + a.zero, 2 * FQ_P1), FQ_P1);
                av[q][e].b = r32_csub(r32_fold(av[q][e].b + x[q][e].b + a.zero, 2 * FQ_P2), FQ_P2);
                acc[q * PW + tau + e * T] = rns_pack(av[q][e]);
            }
        }
        gsync();                                         // ACC_g is only read by group g (next step's rotation)
    }
    // ---- K3: sample extraction of coefficient 0 (+ table-mode offset s*Delta/2 on the body); CRT back to integers.
    // Fused exchange for node-sharded levels: besides the local wire buffer the extracted ciphertext is stored straight
    // into every peer GPU's replica (NVLink peer stores), so no separate all-gather pass over the level's outputs is
    // needed -- the host only barriers between levels (tfhe_fbs_map_b200/dist.py).
    psync();
#pragma unroll
    for (int q = 0; q < TP; q++) {
        if (!live[q]) continue;
        const u64 *A = ACC + q * PW;
        const size_t CT = (size_t)K * N + 1;
        const size_t off = ((size_t)a.bs_slot[node[q]] * a.B + inst[q]) * CT;
        u64 *out = a.wires + off;
        for (int w = ptid; w < K * N && !a.grp_first; w += C::PT) {
            const int u = w / N, j = w % N;
            const rns2 v = rns_unpack(A[(size_t)u * N + (j == 0 ? 0 : N - j)]);
            const u64 val = rns_to_int(j == 0 ? v : rns_neg(v));
            out[w] = val;
            for (int pr = 0; pr < a.n_peers; pr++) a.peer_wires[pr][off + w] = val;
        }
        if (ptid == 0 && !a.grp_first) {
            const u64 val = fq_add(rns_to_int(rns_unpack(A[(size_t)K * N])), fq_mul((u64)mode[q], fbs_delta(p) >> 1));
            out[(size_t)K * N] = val;
            for (int pr = 0; pr < a.n_peers; pr++) a.peer_wires[pr][off + (size_t)K * N] = val;
        }
        if (a.tap_acc) {
            u64 *t = a.tap_acc + (size_t)job[q] * G * N;
            for (int w = ptid; w < G * N; w += C::PT) t[w] = rns_to_int(rns_unpack(A[w]));
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K2u: blind rotation with M = 2 or 3 key bits per step (key unrolling, params bsk_unroll = M; L = 1).
//   For key group t and every non-empty subset T of its M bits: e_T = sum_{i in T} a_i and the indicator bit
//   m_T = prod_{i in T} s_i * prod_{i not in T} (1 - s_i).  Then X^(sum_i a_i s_i) - 1 = sum_T (X^{e_T} - 1) m_T, so with GGSW
//   ciphertexts G_T of the 2^M - 1 indicator bits:   ACC <- ACC + Dec(ACC) [x] ( sum_T (X^{e_T} - 1) G_T ).
//   (M = 2:  X^(a1 s1 + a2 s2) - 1 = (X^(a1+a2) - 1) s1 s2 + (X^a1 - 1) s1 (1 - s2) + (X^a2 - 1) (1 - s1) s2.)
// One decomposition and one forward / inverse transform pair serve M key bits; the monomial factors are applied to the
// KEY in the NTT domain, where X^e is the point-wise multiplication by psi^(e * (2 brev(i) + 1)) (table psi_pow), so the
// accumulator is decomposed as it is: no rotated reads, ACC lives in registers for the whole blind rotation and shared
// memory only holds the transpose scratch (32 KB per bootstrap at N = 2048, k = 1).  Per coefficient and prime the
// point-wise part is
//   bundle_u = REDC( sum_T f_T * key_T[u][g] ),  out_g = REDC( sum_u D_u * bundle_u )       (key in Montgomery form twice)
// i.e. 2 (2^M - 1) + 2 products and 3 reductions for M key bits, against a whole saved step of transforms per extra bit.
// The key ((2^M - 1) / M x the classic one) is streamed by TMA in per-element slices through a shared-memory ring
// (BR2Cfg, full/empty mbarriers); a key word is read from shared memory once per thread and used for every bootstrap the
// thread carries.  The table psi_pow (psi^x - 1) sits in shared memory too, index nibble XOR-folded against bank conflicts.
// ------------------------------------------------------------------------------------------------------
template <int LOGN, int K, int PB, int TP, int M = 2>
struct BR2Cfg {
    static_assert(M == 2 || M == 3, "two or three key bits per step");
    static constexpr int NC = (1 << M) - 1;                       // GGSW ciphertexts (monomial factors) per key group
    static_assert(TP == 1 || TP == PB, "a thread carries one bootstrap or all of the CTA's");
    static constexpr int N = 1 << LOGN, G = K + 1, T = N / 8, PT = G * T, THREADS = (PB / TP) * PT;
    static constexpr size_t s_w = (size_t)G * N;                  // transpose scratch = digit spectra = final accumulator
    static constexpr size_t psi_w = 2 * (size_t)N;                // psi^x table
#ifndef FBS_TW_SMEM
#define FBS_TW_SMEM 1
#endif
#ifndef FBS_TW_SMEM_PB1
#define FBS_TW_SMEM_PB1 1      /* measured: 2.968 -> 2.937 ms per wave of 148 one-bootstrap CTAs */
#endif
    // NTT twiddle table in shared memory (16 B per entry): at M = 2 always; at M = 3 the two-bootstrap CTA has no room for it, the
    // one-bootstrap tail variant has (FBS_TW_SMEM_PB1)
    static constexpr bool TWS = FBS_TW_SMEM != 0 && (M == 2 || (FBS_TW_SMEM_PB1 != 0 && PB == 1));
    static constexpr size_t tw_w = TWS ? 2 * (size_t)N : 0;
    // Key slice = every key word the CTA needs for ONE of the 8 elements a thread holds: [tau >> 5][c < NC][u < G][v < G][tau & 31]; a
    // step consumes 8 slices in element order.  HBM layout [key group][element][tau >> 5][c][u][v][tau & 31]: one bulk copy per slice.
    static constexpr size_t slice_w = NC * (size_t)G * G * T;
    static constexpr size_t fixed_b = 8 * (PB * s_w + psi_w + tw_w) + PB * 2048 + 512;   // ms rows budgeted for n < 1024; barriers
    static constexpr int R_fit = (int)((227 * 1024 - fixed_b) / (8 * slice_w));
#ifndef FBS_PSI_STATIC
#define FBS_PSI_STATIC 1    /* psi^x table in static shared memory (constant address folded into the look-up) */
#endif
#ifndef FBS_LATE_XSYNC
#define FBS_LATE_XSYNC 1   /* wait for the partner warps' digit spectra after element 0's key products instead of before the point-wise loop */
#endif
#ifndef FBS_LAST_REFILLS
#define FBS_LAST_REFILLS 0  /* 1: a block's ring entry is refilled by whichever of its warps releases it last (no waiting issuer) */
#endif
#ifndef FBS_EARLY_RELEASE
#define FBS_EARLY_RELEASE 1 /* release a key-ring entry (and wait for the next) right after its last load instead of at the end of the element */
#endif
#ifndef FBS_BLOCK_RING
#define FBS_BLOCK_RING 1    /* one key ring per block of 32 thread positions (0: one ring of whole slices, refilled by thread 0) */
#endif
#ifndef FBS_RING_MAX
#define FBS_RING_MAX 3      /* measured at set A2: 2 slots 45.8 k, 3 slots 46.3 k, 4 45.8 k, 5 45.1 k PBS/s -- deeper rings only take L1 from the twiddles */
#endif
    static constexpr int R = R_fit > FBS_RING_MAX ? FBS_RING_MAX : R_fit;                   // ring slots (prefetch distance)
    static_assert(R >= 2, "key ring does not fit shared memory");
    __host__ __device__ static constexpr size_t ms_stride(int n) { return (((size_t)(n + 1) * 2 + 15) / 16) * 16; }
    static constexpr int KBMAX = T / 32;                          // rings when FBS_BLOCK_RING (barrier space is reserved either way)
    static constexpr size_t static_b = FBS_PSI_STATIC ? 8 * psi_w : 0;   // part of smem_bytes() that is static shared memory
    __host__ __device__ static constexpr size_t smem_bytes(int n) { return 8 * (PB * s_w + psi_w + tw_w + R * slice_w + 2 * R * KBMAX) + PB * ms_stride(n); }
};
#ifdef FBS_PHASE_CLK   /* profiling builds only: SM clocks warp 0 of every CTA spends per phase of a step (tools/phase_clock.py) */
__device__ unsigned long long g_phase_clk[8];
#define PHASE_MARK(i) do { if (tid == 0) { const long long now_ = clock64(); ph_[i] += now_ - last_; last_ = now_; } } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif
template <int LOGN, int K, int PB, int TP, int M>
__global__ void __launch_bounds__((PB / TP) * (K + 1) * (1 << LOGN) / 8, 1) k_blind_rotate2(BRArgs a)
{
    using C = BR2Cfg<LOGN, K, PB, TP, M>;
    constexpr int NC = C::NC;
    using P = NttPlan<LOGN>;
    constexpr int N = C::N, G = C::G, T = C::T, R = C::R;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, pb0 = (TP == 1) ? tid / C::PT : 0, ptid = tid % C::PT, g = ptid / T, tau = ptid % T;
    constexpr size_t PW = C::s_w, PWB = PW * 8;
    u64 *S = (u64 *)smem_raw + (size_t)pb0 * PW;                 // bootstrap q of this thread: S + q*PW
#if FBS_PSI_STATIC
    // psi^x table in STATIC shared memory: its address is a link-time constant, so a look-up is LDS [offset + constant] and needs
    // no add of the table base (one IMAD.IADD per look-up less, 98 per warp and step)
    __shared__ __align__(16) u64 PSI[2 * N];
    u64 *TW = (u64 *)smem_raw + (size_t)PB * PW;                 // twiddles (when C::TWS)
#else
    u64 *PSI = (u64 *)smem_raw + (size_t)PB * PW;
    u64 *TW = PSI + C::psi_w;                                    // twiddles (when C::TWS)
#endif
    u64 *RING = TW + C::tw_w;
    // Key ring.  FBS_BLOCK_RING: the 32 thread positions of block kb = tau >> 5 only ever read block kb of a slice (7 KB at M = 3),
    // so every block has its OWN ring entries, full/empty mbarriers and issuing lane: the (PB / TP) * G warps that share a block
    // refill it as soon as THEY are through with it, in 7 KB copies, instead of the CTA waiting for its slowest warp and for a
    // 56 KB copy.  Otherwise one ring of whole slices, refilled by thread 0 when all warps have released a slot.
    constexpr int KB = FBS_BLOCK_RING ? T / 32 : 1;               // independent rings
    constexpr size_t KW = C::slice_w / KB;                        // words per ring entry
    constexpr u32 KBYTES = (u32)(KW * 8);
    const int kb = FBS_BLOCK_RING ? tau >> 5 : 0;
    u64 *ring = RING + (size_t)kb * R * KW;
    u64 *full = RING + (size_t)R * C::slice_w + (size_t)kb * 2 * R, *empty = full + R;   // mbarriers: entry landed / entry released by its warps
    const u64 *ksrc = a.bsk + (size_t)kb * KW;
    const bool issuer = FBS_BLOCK_RING ? (pb0 == 0 && g == 0 && (tau & 31) == 0) : tid == 0;
    const size_t ms_stride = C::ms_stride(a.n);
    u16 *s_ms = (u16 *)((unsigned char *)(RING + (size_t)R * C::slice_w + (size_t)C::KBMAX * 2 * R) + (size_t)pb0 * ms_stride);
    const int n = a.n, p = a.p;
    const int n_pairs = (n + M - 1) / M, n_slices = 8 * n_pairs;  // key groups; n is padded with zero key bits (a_i = 0) to a multiple of M

    bool live[TP]; int node[TP], tabL[TP], tab0[TP], mode[TP]; long long inst[TP], job[TP];
#pragma unroll
    for (int q = 0; q < TP; q++) {
        job[q] = a.job_begin + (long long)blockIdx.x * PB + pb0 + q;
        live[q] = job[q] < a.jobs;
        if (!live[q]) job[q] = a.jobs - 1;
        node[q] = a.node_begin + (int)(job[q] / a.B);
        if (a.grp_first) node[q] = a.grp_first[node[q]];           // multi-value: first bootstrap of the group (they share the lincomb)
        inst[q] = job[q] % a.B;
        const u16 *ms = a.ms + ((size_t)(a.bs_lc[node[q]] - a.lc_begin) * a.B + inst[q]) * (size_t)(n + 1);
        tab0[q] = a.bs_tab_ptr[node[q]]; tabL[q] = a.bs_tab_ptr[node[q] + 1] - tab0[q]; mode[q] = a.bs_mode[node[q]];
        u16 *dst = (u16 *)((unsigned char *)s_ms + (size_t)q * ms_stride);
        for (int i = ptid; i <= n; i += C::PT) dst[i] = ms[i];
    }
    // psi^x table, low index nibble XOR-folded with higher index bits: the lanes of a half-warp look up exponents that differ by
    // multiples of 32 E (bit-reversed evaluation points), which would all fall into one bank pair otherwise.  At N = 2048 the fold
    // takes index bits 5..8 only (any bijection of the four lane-dependent bits below the element bits 9..11 gives the same
    // conflict count as folding the top nibble in too -- tools/psi_hash_model.py: 2.34 against 2.38 wavefronts per half-warp
    // request), so the three bits that differ between a thread's 8 elements stay a plain ADDITIVE field of the byte offset.
    constexpr bool fast_psi = (LOGN - 2 >= 9) && (LOGN + 1 <= 12);
    auto psw = [](u32 x) { return fast_psi ? x ^ ((x >> 5) & 15u) : x ^ (((x >> 4) ^ (x >> 8)) & 15u); };
    for (int i = tid; i < 2 * N; i += C::THREADS) PSI[psw((u32)i)] = a.psi_pow[i];
    if (C::TWS) for (int i = tid; i < 2 * N; i += C::THREADS) TW[i] = ((const u64 *)a.psi_rev)[i];
    const fq_tw *twp = C::TWS ? (const fq_tw *)TW : a.psi_rev;
    if (issuer) {
        for (int r = 0; r < R; r++) { mbar_init(full + r, 1); mbar_init(empty + r, FBS_BLOCK_RING ? (PB / TP) * G : C::THREADS / 32); }
    }
    __syncthreads();
    if (issuer) {                                               // prologue: the first R slices
        fence_proxy_async();
        for (int sl = 0; sl < R && sl < n_slices; sl++) {
            mbar_expect_tx(full + sl, KBYTES);
            tma_load_1d(ring + (size_t)sl * KW, ksrc + (size_t)sl * C::slice_w, KBYTES, full + sl);
        }
    }
    // ---- accumulator init in registers: ACC = (0, .., 0, X^{-b~} * TV); this thread holds coefficients j = tau + e*T of polynomial g
    rns2 av[TP][8];
#pragma unroll
    for (int q = 0; q < TP; q++) {
        const u64 delta = fbs_delta(p), off = fq_mul((u64)mode[q], delta >> 1);
        const int bt = ((const u16 *)((const unsigned char *)s_ms + (size_t)q * ms_stride))[n];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int j = tau + e * T;
            u64 val = 0;
            if (g == K) {
                int src = j + bt;
                bool neg = false;
                if (src >= 2 * N) src -= 2 * N;
                if (src >= N) { src -= N; neg = true; }
                int x = (int)((2LL * src * p + N) / (2LL * N));
                if (x >= p) { x -= p; neg = neg != (a.grp_first == nullptr); }     // half slot: the table's step function is negated there, the multi-value base polynomial is not
                const u64 tvx = (x < tabL[q]) ? (u64)__ldg(a.bs_tab + tab0[q] + x) : 0;
                const u64 F = a.grp_first ? fbs_half_delta(p) : fq_sub(fq_mul(tvx, delta), off);
                val = neg ? fq_neg(F) : F;
            }
            av[q][e] = rns_from_int(val);
        }
    }
    const int bar_g = 1 + pb0 * G + g, bar_p = 1 + PB * G + pb0;
    auto gsync = [bar_g] { bar_sync_named(bar_g, T); };
    auto psync = [bar_p] { if (TP == 1) bar_sync_named(bar_p, C::PT); else __syncthreads(); };
    // Digit spectra are exchanged between the G threads that hold the SAME spectrum positions (same tau, one per group):
    // a named barrier over those G warps replaces a CTA-wide one, so the groups only couple warp by warp.
    constexpr int BAR_X0 = 1 + PB * G + PB;                       // after the group (bar_g) and bootstrap (bar_p) barrier ids
    // 16 hardware barriers: when (bootstraps) x (warps per group) pairs do not fit, XG neighbouring warp positions share one
    constexpr int XW = (PB / TP) * (T / 32), XG = BAR_X0 + XW <= 16 ? 1 : BAR_X0 + XW / 2 <= 16 ? 2 : 4;
    static_assert(BAR_X0 + XW / XG <= 16 && (T / 32) % XG == 0, "named barriers");
    const int bar_x = BAR_X0 + pb0 * (T / 32 / XG) + (tau >> 5) / XG;
    auto xsync = [bar_x] { bar_sync_named(bar_x, G * 32 * XG); };
    const int beta = a.beta;
    const u64 rc = 1ULL << (62 - beta);
    constexpr int SH = LOGN + 3;
    unsigned char *Sb = (unsigned char *)S;
    u32 bo[LOGN];
#pragma unroll
    for (int lb = 0; lb < LOGN; lb++) bo[lb] = P::tau_boff(tau, lb) | ((u32)g << SH);
    // NTT output position 8*tau + e holds the evaluation at psi^(2 brev(8 tau + e) + 1) = psi^(odd0 + (brev3(e) << (LOGN-2)))
    const u32 odd0 = 2u * (__brev((u32)tau) >> (32 - (LOGN - 3))) + 1u;
    // key words of this thread inside a slice: (((tau >> 5)*NC + c)*G*G + u*G + g)*32 + (tau & 31), u = (g + og) mod G
    u32 koff[G];
#pragma unroll
    for (int og = 0; og < G; og++) { int gg = g + og; if (gg >= G) gg -= G; koff[og] = (u32)((((FBS_BLOCK_RING ? 0 : tau >> 5) * NC * G * G + gg * G + g) * 32 + (tau & 31)) * 8); }
    int slot = 0; u32 par = 0;                                    // ring position of the next slice to consume

#ifdef FBS_PHASE_CLK
    long long ph_[5] = {0, 0, 0, 0, 0}, last_ = clock64();
#endif
    for (int t = 0; t < n_pairs; t++) {
        // ---- decompose ACC_g itself: one balanced digit per coefficient (L = 1), lazy residues in (0, 2p)
        rns2 dg[1][TP][8];
#pragma unroll
        for (int q = 0; q < TP; q++)
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const u32 tt = rns_crt_hi(av[q][e]);
                const u32 d = (u32)fbs_digit1_t(tt, av[q][e].a, beta, rc);
                dg[0][q][e].a = d + FQ_P1;
                dg[0][q][e].b = d + FQ_P2;
            }
        PHASE_MARK(0);                                       // decomposition
        ntt_fwd1_from<LOGN, 0, TP, decltype(gsync), C::TWS>(dg[0], tau, Sb, PWB, bo, twp, gsync, true, a.zero);
        PHASE_MARK(1);                                       // forward transform
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 o = bo[0] ^ P::elem_boff(e, 0);
#pragma unroll
            for (int q = 0; q < TP; q++) {
                dg[0][q][e].a = r32_fold(dg[0][q][e].a, 2 * FQ_P1);     // < 2p: keeps the sums below within 64 bits / the folds in range
                dg[0][q][e].b = r32_fold(dg[0][q][e].b, 2 * FQ_P2);
                *(u64 *)(Sb + q * PWB + o) = rns_pack(dg[0][q][e]);
            }
        }
        rns2 x[TP][8];
        // slice consumed by this warp (its key words are in registers); when every warp is through, thread 0 refills the
        // slot with the slice R ahead, so the copy overlaps the other elements' arithmetic or the transforms
        auto release_slot = [&](int e) {
            __syncwarp();
#if FBS_LAST_REFILLS && FBS_BLOCK_RING
            // The warp whose arrival completes the entry's "empty" phase refills it -- known from the state the arrive returns (pending
            // count 1 before the arrive = this was the last one): exactly one lane issues the copy and nobody ever waits for a partner.
            if ((tid & 31) == 0) {
                u64 st_; u32 pend_;
                asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st_) : "r"(smem_u32(empty + slot)) : "memory");
                asm("mbarrier.pending_count.b64 %0, %1;" : "=r"(pend_) : "l"(st_));
                const int nxt = 8 * t + e + R;
                if (pend_ == 1u && nxt < n_slices) {
                    asm volatile("fence.acq_rel.cta;" ::: "memory");        // the other warps' reads of the entry (released by their arrives)
                    fence_proxy_async();
                    mbar_expect_tx(full + slot, KBYTES);
                    tma_load_1d(ring + (size_t)slot * KW, ksrc + (size_t)nxt * C::slice_w, KBYTES, full + slot);
                }
            }
#else
            if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + slot)) : "memory");
            if (issuer) {
                const int nxt = 8 * t + e + R;
                if (nxt < n_slices) {
                    mbar_wait(empty + slot, par);
                    fence_proxy_async();
                    mbar_expect_tx(full + slot, KBYTES);
                    tma_load_1d(ring + (size_t)slot * KW, ksrc + (size_t)nxt * C::slice_w, KBYTES, full + slot);
                }
            }
#endif
            if (++slot == R) { slot = 0; par ^= 1; }
        };
        static_assert(M == 2 || G == 2, "the M = 3 sums are sized for k = 1");
        // Exponents of the NC monomial factors, E_c = sum of the a_i in subset c.  Element e evaluates X^E at
        // psi^(E*odd0 + (brev3(e) << (LOGN-2)) * E): the low LOGN-2 bits of the table index are the same for the 8 elements,
        // only the top three move (by brev3(e) * E mod 8).  One packed word per (bootstrap, factor).  Fast path (N = 2048): the
        // table's fold does not touch those three bits, so with PK = byte offset of element 0 | (E & 7) -- E mod 8 parked in the
        // three low bits that are zero in an 8-byte offset -- element e's offset is (PK * (1 + (brev3(e) << (LOGN + 1)))) & mask:
        // the multiply adds brev3(e) * (E & 7) into the field at bits LOGN+1.. (overflow and the parked bits are masked off).
        // One IMAD (or shift-add), one LOP3 and the load per look-up.  Generic path: x0 | (E & 7) << 20.
        u32 PK[TP][NC];
#pragma unroll
        for (int q = 0; q < TP; q++) {
            const u16 *msq = (const u16 *)((const unsigned char *)s_ms + (size_t)q * ms_stride);
            u32 ai[M];
#pragma unroll
            for (int i = 0; i < M; i++) ai[i] = (M * t + i < n) ? msq[M * t + i] : 0u;   // msq[n] is the body, not a mask element
#pragma unroll
            for (int c = 0; c < NC; c++) {
                u32 E = 0;
#pragma unroll
                for (int i = 0; i < M; i++) if ((fbs_unroll_mask(M, c) >> i) & 1) E += ai[i];
                const u32 x0 = (E * odd0) & (2 * N - 1);
                if constexpr (fast_psi) PK[q][c] = (8u * psw(x0)) | (E & 7u);
                else PK[q][c] = x0 | ((E & 7u) << 20);
            }
        }
#if !FBS_LATE_XSYNC
        xsync();                                             // the partner warps' spectra are in shared memory
#endif
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 o = bo[0] ^ P::elem_boff(e, 0);
            constexpr int BR3[8] = {0, 4, 2, 6, 1, 5, 3, 7};
#ifdef FBS_PHASE_CLK
            const long long w0_ = clock64();
#endif
            if (!FBS_EARLY_RELEASE || e == 0) mbar_wait(full + slot, par);      // (elements 1..7: waited for below, one element ahead)
#ifdef FBS_PHASE_CLK
            if (tid == 0) ph_[4] += clock64() - w0_;         // part of phase 2 spent waiting for the key block
#endif
            const unsigned char *ks = (const unsigned char *)(ring + (size_t)slot * KW);
            auto factor = [&](int q, int c) -> rns2 {        // X^{E_c} - 1 at this thread's element e
                const u32 pk = PK[q][c];
                if constexpr (fast_psi) {
#ifdef FBS_PSI_FAKE   /* timing experiment only (wrong results): the same instructions, but conflict-free addresses */
                    const u32 off = ((pk * (1u + ((u32)BR3[e] << (LOGN + 1)))) & (7u << (LOGN + 1))) | ((u32)(tid & 31) << 3);
#else
                    const u32 off = (pk * (1u + ((u32)BR3[e] << (LOGN + 1)))) & ((2u * N - 1u) << 3);
#endif
                    return rns_split(*(const u64 *)((const unsigned char *)PSI + off));
                } else {
                    const u32 xi = ((pk & 0xFFFFu) + (((u32)BR3[e] * (pk >> 20)) << (LOGN - 2))) & (2 * N - 1);
                    return rns_split(PSI[psw(xi)]);
                }
            };
            // bundle_u = REDC(sum_c f_c * key_c[u][g]); factors outermost so that one key word (shared by the bootstraps
            // the thread carries) and one factor are live at a time next to the 64-bit accumulators
            u64 pa[TP][G], pb2[TP][G];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                rns2 kk[G];
#pragma unroll
                for (int og = 0; og < G; og++) kk[og] = rns_split(*(const u64 *)(ks + koff[og] + (size_t)c * G * G * 32 * 8));
#pragma unroll
                for (int q = 0; q < TP; q++) {
                    const rns2 f = factor(q, c);
#pragma unroll
                    for (int og = 0; og < G; og++) {
                        if (c == 0) { pa[q][og] = r32_mulwide(f.a, kk[og].a); pb2[q][og] = r32_mulwide(f.b, kk[og].b); }
                        else { pa[q][og] = r32_madwide(f.a, kk[og].a, pa[q][og]); pb2[q][og] = r32_madwide(f.b, kk[og].b, pb2[q][og]); }
                    }                                                                    // < NC p^2 <= 7 p^2 < 2^63
                }
            }
#if FBS_LATE_XSYNC
            if (e == 0) xsync();                             // the partner warps' spectra are only needed from here on: element 0's key
#endif                                                       // products do not wait for the partner
            rns2 dsp[TP][G];                                 // digit spectra of this element (< 2p): own group's and the partner groups'
#pragma unroll
            for (int q = 0; q < TP; q++)
#pragma unroll
                for (int og = 0; og < G; og++) {
                    int gg = g + og; if (gg >= G) gg -= G;
                    const u32 xg = (og == 0) ? 0u : (G == 2) ? (1u << SH) : ((u32)(g ^ gg) << SH);
                    dsp[q][og] = rns_split(*(const u64 *)(Sb + q * PWB + (o ^ xg)));
                }
#if FBS_EARLY_RELEASE
            // The key words of this element are in registers: release its ring entry and wait for the next element's entry NOW, so
            // that the reductions below (no key loads) overlap the first loads of the next element instead of sitting between two
            // scheduling barriers.  (mbarrier.arrive has release semantics: the loads above are performed before it is observed.)
            release_slot(e);
            if (e < 7) mbar_wait(full + slot, par);
#endif
#pragma unroll
            for (int q = 0; q < TP; q++) {
                u64 oa = 0, ob = 0;
#pragma unroll
                for (int og = 0; og < G; og++) {
                    const u32 ba = r32_redc(pa[q][og], FQ_P1, FQ_P1_INVNEG), bb = r32_redc(pb2[q][og], FQ_P2, FQ_P2_INVNEG);   // < (NC/4 + 1) p + 1
                    const rns2 d = dsp[q][og];
                    if (og == 0) { oa = r32_mulwide(d.a, ba); ob = r32_mulwide(d.b, bb); }
                    else { oa = r32_madwide(d.a, ba, oa); ob = r32_madwide(d.b, bb, ob); }        // < G * 2 (NC/4 + 1) p^2 < 2^64 (M = 2: G <= 3, M = 3: G = 2)
                }
                x[q][e].a = r32_fold(r32_redc(oa, FQ_P1, FQ_P1_INVNEG), 2 * FQ_P1);               // < 4 p before the fold
                x[q][e].b = r32_fold(r32_redc(ob, FQ_P2, FQ_P2_INVNEG), 2 * FQ_P2);
            }
#if !FBS_EARLY_RELEASE
            release_slot(e);
#endif
        }
        PHASE_MARK(2);                                           // spectra exchange + point-wise products
        auto after_pass0 = [&] { xsync(); };                     // the partner warps have read this step's digit spectra
        ntt_inv1_from<LOGN, 0, TP, decltype(after_pass0), decltype(gsync), C::TWS>(x, tau, Sb, PWB, bo, twp, after_pass0, gsync, a.zero);
        PHASE_MARK(3);                                           // inverse transform
#pragma unroll
        for (int e = 0; e < 8; e++)
#pragma unroll
            for (int q = 0; q < TP; q++) {
                av[q][e].a = r32_csub(r32_fold(av[q][e].a + x[q][e].a + a.zero, 2 * FQ_P1), FQ_P1);
                av[q][e].b = r32_csub(r32_fold(av[q][e].b + x[q][e].b + a.zero, 2 * FQ_P2), FQ_P2);
            }
    }
#ifdef FBS_PHASE_CLK
    if (tid == 0) for (int i = 0; i < 5; i++) atomicAdd(&g_phase_clk[i], (unsigned long long)ph_[i]);
#endif
    // ---- accumulator to shared memory (natural order), then K3 as in k_blind_rotate
    gsync();                                             // the group's last transposed reads are done
#pragma unroll
    for (int q = 0; q < TP; q++)
#pragma unroll
        for (int e = 0; e < 8; e++) S[q * PW + (size_t)g * N + tau + e * T] = rns_pack(av[q][e]);
    psync();
#pragma unroll
    for (int q = 0; q < TP; q++) {
        if (!live[q]) continue;
        const u64 *A = S + q * PW;
        const size_t CT = (size_t)K * N + 1;
        const size_t off = ((size_t)a.bs_slot[node[q]] * a.B + inst[q]) * CT;
        u64 *out = a.wires + off;
        for (int w = ptid; w < K * N && !a.grp_first; w += C::PT) {
            const int u = w / N, j = w % N;
            const rns2 v = rns_unpack(A[(size_t)u * N + (j == 0 ? 0 : N - j)]);
            const u64 val = rns_to_int(j == 0 ? v : rns_neg(v));
            out[w] = val;
            for (int pr = 0; pr < a.n_peers; pr++) a.peer_wires[pr][off + w] = val;
        }
        if (ptid == 0 && !a.grp_first) {
            const u64 val = fq_add(rns_to_int(rns_unpack(A[(size_t)K * N])), fq_mul((u64)mode[q], fbs_delta(p) >> 1));
            out[(size_t)K * N] = val;
            for (int pr = 0; pr < a.n_peers; pr++) a.peer_wires[pr][off + (size_t)K * N] = val;
        }
        if (a.tap_acc) {
            u64 *t2 = a.tap_acc + (size_t)job[q] * G * N;
            for (int w = ptid; w < G * N; w += C::PT) t2[w] = rns_to_int(rns_unpack(A[w]));
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K2c: low-latency blind rotation -- ONE bootstrap split across a thread-block cluster of C = 2^LOGC CTAs.
// For launches with fewer jobs than SMs (narrow circuit levels, node-sharded levels, small batches) the one-CTA kernels leave
// most of the chip idle and a bootstrap takes n/M strictly sequential steps of one SM's latency-bound work.  Here the
// ring elements are split over C SMs (ntt.cuh "cluster-split transform"): CTA h owns, in the coefficient domain, the residues
// r in [h*Ns/C, (h+1)*Ns/C) of every sub-block (accumulate / decompose / cross butterflies are register work), and in the
// spectral domain the positions [h*Ns, (h+1)*Ns) (local transforms of size Ns = N/C, point-wise products with 1/C of the key).
// Per step two all-to-all exchanges through DISTRIBUTED SHARED MEMORY (st.async into the destination CTA's inbox,
// completion counted in bytes on that CTA's mbarrier): no cluster-wide barrier inside the loop, the data dependence is the
// flow control (a CTA cannot send exchange k+1 before it has received everybody's exchange k, and everybody sent that only
// after consuming what it received before).  The key is streamed by TMA as in k_blind_rotate2 (same HBM layout, blocked by 32
// thread positions so that a CTA's share of a slice is one contiguous bulk copy), ring depth up to 8 slices = one whole step ahead.
// Arithmetic, rounding and term order are those of k_blind_rotate2<M>: results are bit-identical.
// ------------------------------------------------------------------------------------------------------
template <int LOGN, int K, int M, int LOGC, int OCC = 1>         // OCC: CTAs per SM the shared-memory budget (key ring depth) is sized for
struct BRCCfg {
    static_assert(M == 2 || M == 3, "two or three key bits per step");
    static_assert(LOGC >= 1 && LOGC <= 3, "cluster of 2, 4 or 8 CTAs");
    static constexpr int NC = (1 << M) - 1;
    static constexpr int N = 1 << LOGN, G = K + 1, C = 1 << LOGC, LOGNS = LOGN - LOGC, Ns = N >> LOGC, Ts = Ns / 8, R = 8 / C, T = N / 8;
    static_assert(Ts >= 32 && Ts % 32 == 0, "a polynomial's threads in a CTA are whole warps");
    static constexpr int THREADS = G * Ts;
    static constexpr int RUNS = NC * G * G;                       // key words per thread position and slice
    static constexpr size_t s_w = (size_t)G * Ns;                 // transpose scratch / digit spectra
    static constexpr size_t inbox_w = (size_t)G * 8 * Ts;         // one exchange direction
    static constexpr size_t psi_w = 2 * (size_t)N;
    static constexpr size_t tw_w = 2 * (size_t)Ns;                // one local twiddle table (16 B entries)
    static constexpr size_t slice_w = (size_t)RUNS * Ts;
    static constexpr size_t fixed_b = 8 * (s_w + 2 * inbox_w + psi_w + 2 * tw_w) + 2048 + 256;
    static constexpr int R_fit = (int)(((227 * 1024) / OCC - 1024 * (OCC > 1) - fixed_b) / (8 * slice_w));       // 1 KB per CTA is reserved by the system
    static constexpr int RING = R_fit > 8 ? 8 : R_fit;
    static_assert(RING >= 2, "key ring does not fit shared memory");
    __host__ __device__ static constexpr size_t ms_stride(int n) { return (((size_t)(n + 1) * 2 + 15) / 16) * 16; }
    __host__ __device__ static constexpr size_t smem_bytes(int n) { return 8 * (s_w + 2 * inbox_w + psi_w + 2 * tw_w + RING * slice_w + 2 * RING + 2) + ms_stride(n); }
    static constexpr size_t static_b = FBS_PSI_STATIC ? 8 * psi_w : 0;    // part of smem_bytes() that is static shared memory
};
__device__ __forceinline__ u32 cluster_ctarank() { u32 r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ u32 mapa_shared(u32 local_addr, u32 cta) { u32 r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 8-byte store into another CTA's shared memory, completion (8 bytes) counted on that CTA's mbarrier
__device__ __forceinline__ void st_async_u64(u32 remote_addr, u64 v, u32 remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(v), "r"(remote_bar) : "memory");
}
// Wait for an exchange inbox: the data arrives in THIS CTA's shared memory with the mbarrier's complete_tx, so the ordinary
// (CTA-scope) acquire of try_wait covers it -- a cluster-scope acquire makes ptxas add CCTL.IVALL (an L1 invalidate that
// shared memory does not need; measured 4.5 % of the stall samples, profiles/r2_cl4_v1_*).  The spin is bounded: a cluster
// whose exchange never completes (a lost store would be a bug, not a load condition) traps after FBS_XCHG_TIMEOUT_NS instead of
// hanging the GPU.
#ifndef FBS_XCHG_TIMEOUT_NS
#define FBS_XCHG_TIMEOUT_NS 4000000000ULL
#endif
__device__ __forceinline__ void mbar_wait_cluster(u64 *bar, u32 parity)
{
    u64 t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        u64 t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t0 == 0) t0 = t1;
        else if (t1 - t0 > FBS_XCHG_TIMEOUT_NS) __trap();
    }
}
template <int LOGN, int K, int M, int LOGC>
__global__ void __launch_bounds__((K + 1) * (1 << (LOGN - LOGC)) / 8, 1) k_blind_rotate_cl(BRArgs a)
{
    using Cf = BRCCfg<LOGN, K, M, LOGC>;
    using P = NttPlan<Cf::LOGNS>;
    constexpr int NC = Cf::NC, N = Cf::N, G = Cf::G, C = Cf::C, Ns = Cf::Ns, Ts = Cf::Ts, R = Cf::R, T = Cf::T, RING = Cf::RING, LOGNS = Cf::LOGNS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, g = tid / Ts, tau = tid % Ts, lane = tid & 31, warp = tid >> 5;
    const int h = (int)cluster_ctarank();                         // this CTA's sub-block / residue range
    const int tau_g = h * Ts + tau;                               // spectrum positions 8*tau_g + e of the full transform
    u64 *S = (u64 *)smem_raw;
    u64 *INF = S + Cf::s_w, *INI = INF + Cf::inbox_w;             // inboxes: forward / inverse exchange, [g][e][tau]
#if FBS_PSI_STATIC
    __shared__ __align__(16) u64 PSI[2 * N];                      // static: constant address folded into the look-ups (see k_blind_rotate2)
    u64 *TWF = INI + Cf::inbox_w, *TWI = TWF + Cf::tw_w;          // local twiddle tables (fq_tw entries)
#else
    u64 *PSI = INI + Cf::inbox_w;
    u64 *TWF = PSI + Cf::psi_w, *TWI = TWF + Cf::tw_w;            // local twiddle tables (fq_tw entries)
#endif
    u64 *RNG = TWI + Cf::tw_w;
    u64 *full = RNG + (size_t)RING * Cf::slice_w, *empty = full + RING, *xbar = empty + RING;   // xbar[0]: forward inbox full, xbar[1]: inverse
    u16 *s_ms = (u16 *)(xbar + 2);
    const int n = a.n, p = a.p;
    const int n_pairs = (n + M - 1) / M, n_slices = 8 * n_pairs;

    const long long job = a.job_begin + (long long)(blockIdx.x >> LOGC);      // grid = exactly C CTAs per job
    const int node = a.grp_first ? a.grp_first[a.node_begin + (int)(job / a.B)] : a.node_begin + (int)(job / a.B);
    const long long inst = job % a.B;
    {
        const u16 *ms = a.ms + ((size_t)(a.bs_lc[node] - a.lc_begin) * a.B + inst) * (size_t)(n + 1);
        for (int i = tid; i <= n; i += Cf::THREADS) s_ms[i] = ms[i];
    }
    const int tab0 = a.bs_tab_ptr[node], tabL = a.bs_tab_ptr[node + 1] - tab0, mode = a.bs_mode[node];
    constexpr bool fast_psi = (LOGN - 2 >= 9) && (LOGN + 1 <= 12);                    // as in k_blind_rotate2
    auto psw = [](u32 x) { return fast_psi ? x ^ ((x >> 5) & 15u) : x ^ (((x >> 4) ^ (x >> 8)) & 15u); };
    for (int i = tid; i < 2 * N; i += Cf::THREADS) PSI[psw((u32)i)] = a.psi_pow[i];
    for (int i = 1 + tid; i < Ns; i += Cf::THREADS) {
        ((uint4 *)TWF)[i] = __ldg((const uint4 *)a.psi_rev + ntt_local_src(i, C + h));
        ((uint4 *)TWI)[i] = __ldg((const uint4 *)a.psi_rev + ntt_local_src(i, 2 * C - 1 - h));
    }
    fq_tw cw[C];                                                  // cross-stage twiddles psi_rev[1 .. C)
#pragma unroll
    for (int i = 1; i < C; i++) cw[i] = fq_tw_load<false>(a.psi_rev + i);
    cw[0] = cw[1];
    if (tid == 0) {
        for (int r = 0; r < RING; r++) { mbar_init(full + r, 1); mbar_init(empty + r, Cf::THREADS / 32); }
        mbar_init(xbar, 1); mbar_init(xbar + 1, 1);
    }
    __syncthreads();
    cluster_sync_all();                                           // every CTA's barriers exist before anybody stores remotely
    // key slice `sl` of this CTA: the Ts/32 blocks of 32 thread positions it owns are contiguous in HBM: one bulk copy of slice_w words
    auto issue_slice = [&](int sl, int slot) {                    // called by thread 0
        fence_proxy_async();
        mbar_expect_tx(full + slot, (u32)(Cf::slice_w * 8));
        tma_load_1d(RNG + (size_t)slot * Cf::slice_w, a.bsk + ((size_t)sl * C + h) * Cf::slice_w, (u32)(Cf::slice_w * 8), full + slot);
    };
    if (tid == 0) for (int sl = 0; sl < RING && sl < n_slices; sl++) issue_slice(sl, sl);
    // remote addresses of the inboxes / exchange barriers of every CTA of the cluster
    u32 r_inf[C], r_ini[C], r_xf[C], r_xi[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        r_inf[c] = mapa_shared(smem_u32(INF), (u32)c); r_ini[c] = mapa_shared(smem_u32(INI), (u32)c);
        r_xf[c] = mapa_shared(smem_u32(xbar), (u32)c); r_xi[c] = mapa_shared(smem_u32(xbar + 1), (u32)c);
    }
    const u32 send_off = (u32)((((g * 8 + h * R) * Ts) + tau) * 8);    // my R values land in rows h*R + ri of the receiver's inbox
    const u32 inbox_bytes = (u32)(Cf::inbox_w * 8);
    // ---- accumulator init in registers: register e = hh*R + ri holds coefficient j = hh*Ns + h*(Ns/C) + ri*Ts + tau of polynomial g
    rns2 av[8];
    {
        const u64 delta = fbs_delta(p), off = fq_mul((u64)mode, delta >> 1);
        const int bt = s_ms[n];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int j = (e / R) * Ns + h * (Ns / C) + (e % R) * Ts + tau;
            u64 val = 0;
            if (g == K) {
                int src = j + bt;
                bool neg = false;
                if (src >= 2 * N) src -= 2 * N;
                if (src >= N) { src -= N; neg = true; }
                int x = (int)((2LL * src * p + N) / (2LL * N));
                if (x >= p) { x -= p; neg = neg != (a.grp_first == nullptr); }     // half slot: the table's step function is negated there, the multi-value base polynomial is not
                const u64 tvx = (x < tabL) ? (u64)__ldg(a.bs_tab + tab0 + x) : 0;
                const u64 F = a.grp_first ? fbs_half_delta(p) : fq_sub(fq_mul(tvx, delta), off);
                val = neg ? fq_neg(F) : F;
            }
            av[e] = rns_from_int(val);
        }
    }
    const int bar_g = 1 + g;
    auto gsync = [bar_g] { bar_sync_named(bar_g, Ts); };
    const int bar_x = 1 + G + (tau >> 5);
    auto xsync = [bar_x] { bar_sync_named(bar_x, G * 32); };
    const int beta = a.beta;
    const u64 rc = 1ULL << (62 - beta);
    constexpr int SH = LOGNS + 3;
    constexpr size_t PWB = Cf::s_w * 8;
    unsigned char *Sb = (unsigned char *)S;
    u32 bo[LOGNS];
#pragma unroll
    for (int lb = 0; lb < LOGNS; lb++) bo[lb] = P::tau_boff(tau, lb) | ((u32)g << SH);
    static_assert(P::idx(1, 1, P::inv_lb(P::NPASS - 1)) == 1 + Ts && P::idx(1, 1, P::fwd_lb(0)) == 1 + Ts, "local transforms start / end with register e at local index tau + e*Ts");
    const u32 odd0 = 2u * (__brev((u32)tau_g) >> (32 - (LOGN - 3))) + 1u;
    u32 koff[G];
#pragma unroll
    for (int og = 0; og < G; og++) { int gg = g + og; if (gg >= G) gg -= G; koff[og] = (u32)((((tau >> 5) * NC * G * G + gg * G + g) * 32 + (tau & 31)) * 8); }
    int slot = 0; u32 par = 0;
    const fq_tw *twf = (const fq_tw *)TWF, *twi = (const fq_tw *)TWI;

    for (int t = 0; t < n_pairs; t++) {
        const u32 xpar = (u32)(t & 1);
        // ---- decompose, cross butterflies, forward exchange
        rns2 dg[1][8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 tt = rns_crt_hi(av[e]);
            const u32 d = (u32)fbs_digit1_t(tt, av[e].a, beta, rc);
            dg[0][e].a = d + FQ_P1;
            dg[0][e].b = d + FQ_P2;
        }
        ntt_cross_fwd<LOGC>(dg[0], cw, a.zero);
        if (tid == 0) mbar_expect_tx(xbar, inbox_bytes);
#pragma unroll
        for (int e = 0; e < 8; e++) st_async_u64(r_inf[e / R] + send_off + (u32)((e % R) * Ts * 8), rns_pack(dg[0][e]), r_xf[e / R]);
        mbar_wait_cluster(xbar, xpar);
#pragma unroll
        for (int e = 0; e < 8; e++) dg[0][e] = rns_unpack(INF[(size_t)(g * 8 + e) * Ts + tau]);
        // ---- local forward transform of size Ns; spectra (position 8*tau_g + e) to S for the partner groups
        ntt_fwd1_from<LOGNS, 0, 1, decltype(gsync), true>(dg, tau, Sb, PWB, bo, twf, gsync, true, a.zero);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 o = bo[0] ^ P::elem_boff(e, 0);
            dg[0][e].a = r32_fold(dg[0][e].a, 2 * FQ_P1);
            dg[0][e].b = r32_fold(dg[0][e].b, 2 * FQ_P2);
            *(u64 *)(Sb + o) = rns_pack(dg[0][e]);
        }
        rns2 x[1][8];
        auto release_slot = [&](int e) {
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + slot)) : "memory");
            if (tid == 0) {
                const int nxt = 8 * t + e + RING;
                if (nxt < n_slices) { mbar_wait(empty + slot, par); issue_slice(nxt, slot); }
            }
            if (++slot == RING) { slot = 0; par ^= 1; }
        };
        u32 PK[NC];
        {
            u32 ai[M];
#pragma unroll
            for (int i = 0; i < M; i++) ai[i] = (M * t + i < n) ? s_ms[M * t + i] : 0u;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                u32 E = 0;
#pragma unroll
                for (int i = 0; i < M; i++) if ((fbs_unroll_mask(M, c) >> i) & 1) E += ai[i];
                const u32 x0 = (E * odd0) & (2 * N - 1);
                if constexpr (fast_psi) PK[c] = (8u * psw(x0)) | (E & 7u);                  // see k_blind_rotate2
                else PK[c] = x0 | ((E & 7u) << 20);
            }
        }
#if !FBS_LATE_XSYNC
        xsync();
#endif
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 o = bo[0] ^ P::elem_boff(e, 0);
            constexpr int BR3[8] = {0, 4, 2, 6, 1, 5, 3, 7};
            if (!FBS_EARLY_RELEASE || e == 0) mbar_wait(full + slot, par);      // elements 1..7: waited for one element ahead (see k_blind_rotate2)
            const unsigned char *ks = (const unsigned char *)(RNG + (size_t)slot * Cf::slice_w);
            auto factor = [&](int c) -> rns2 {
                const u32 pk = PK[c];
                if constexpr (fast_psi) {
                    const u32 off = (pk * (1u + ((u32)BR3[e] << (LOGN + 1)))) & ((2u * N - 1u) << 3);
                    return rns_split(*(const u64 *)((const unsigned char *)PSI + off));
                } else {
                    const u32 xi = ((pk & 0xFFFFu) + (((u32)BR3[e] * (pk >> 20)) << (LOGN - 2))) & (2 * N - 1);
                    return rns_split(PSI[psw(xi)]);
                }
            };
            u64 pa[G], pb2[G];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const rns2 f = factor(c);
#pragma unroll
                for (int og = 0; og < G; og++) {
                    const rns2 kk = rns_split(*(const u64 *)(ks + koff[og] + (size_t)c * G * G * 32 * 8));
                    if (c == 0) { pa[og] = r32_mulwide(f.a, kk.a); pb2[og] = r32_mulwide(f.b, kk.b); }
                    else { pa[og] = r32_madwide(f.a, kk.a, pa[og]); pb2[og] = r32_madwide(f.b, kk.b, pb2[og]); }
                }
            }
#if FBS_LATE_XSYNC
            if (e == 0) xsync();                             // partner spectra are needed from here on (see k_blind_rotate2)
#endif
#if FBS_EARLY_RELEASE
            release_slot(e);                                 // key words are in registers: free the entry, wait for the next one now
            if (e < 7) mbar_wait(full + slot, par);
#endif
            u64 oa = 0, ob = 0;
#pragma unroll
            for (int og = 0; og < G; og++) {
                const u32 ba = r32_redc(pa[og], FQ_P1, FQ_P1_INVNEG), bb = r32_redc(pb2[og], FQ_P2, FQ_P2_INVNEG);
                int gg = g + og; if (gg >= G) gg -= G;
                const u32 xg = (og == 0) ? 0u : (G == 2) ? (1u << SH) : ((u32)(g ^ gg) << SH);
                const rns2 d = rns_split(*(const u64 *)(Sb + (o ^ xg)));
                if (og == 0) { oa = r32_mulwide(d.a, ba); ob = r32_mulwide(d.b, bb); }
                else { oa = r32_madwide(d.a, ba, oa); ob = r32_madwide(d.b, bb, ob); }
            }
            x[0][e].a = r32_fold(r32_redc(oa, FQ_P1, FQ_P1_INVNEG), 2 * FQ_P1);
            x[0][e].b = r32_fold(r32_redc(ob, FQ_P2, FQ_P2_INVNEG), 2 * FQ_P2);
#if !FBS_EARLY_RELEASE
            release_slot(e);
#endif
        }
        // ---- local inverse transform, inverse exchange, cross butterflies, accumulate
        auto after_pass0 = [&] { xsync(); };
        ntt_inv1_from<LOGNS, 0, 1, decltype(after_pass0), decltype(gsync), true>(x, tau, Sb, PWB, bo, twi, after_pass0, gsync, a.zero);
        if (tid == 0) mbar_expect_tx(xbar + 1, inbox_bytes);
#pragma unroll
        for (int e = 0; e < 8; e++) st_async_u64(r_ini[e / R] + send_off + (u32)((e % R) * Ts * 8), rns_pack(x[0][e]), r_xi[e / R]);
        mbar_wait_cluster(xbar + 1, xpar);
#pragma unroll
        for (int e = 0; e < 8; e++) x[0][e] = rns_unpack(INI[(size_t)(g * 8 + e) * Ts + tau]);
        ntt_cross_inv<LOGC>(x[0], cw, a.zero);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            av[e].a = r32_csub(r32_fold(av[e].a + x[0][e].a + a.zero, 2 * FQ_P1), FQ_P1);
            av[e].b = r32_csub(r32_fold(av[e].b + x[0][e].b + a.zero, 2 * FQ_P2), FQ_P2);
        }
    }
    // ---- K3: sample extraction straight from the registers (every CTA writes the coefficients it owns), fused peer stores
    {
        const size_t CT = (size_t)K * N + 1;
        const size_t off = ((size_t)a.bs_slot[node] * a.B + inst) * CT;
        u64 *out = a.wires + off;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int j = (e / R) * Ns + h * (Ns / C) + (e % R) * Ts + tau;
            if (a.grp_first) {                                       // multi-value: k_multi_extract finishes from tap_acc
            } else if (g < K) {                                      // mask: out[g*N + jj] = (jj == 0 ? a_0 : -a_{N-jj}), jj = (N - j) mod N
                const int jj = (j == 0) ? 0 : N - j;
                const u64 val = rns_to_int(j == 0 ? av[e] : rns_neg(av[e]));
                out[(size_t)g * N + jj] = val;
                for (int pr = 0; pr < a.n_peers; pr++) a.peer_wires[pr][off + (size_t)g * N + jj] = val;
            } else if (j == 0) {
                const u64 val = fq_add(rns_to_int(av[e]), fq_mul((u64)mode, fbs_delta(p) >> 1));
                out[(size_t)K * N] = val;
                for (int pr = 0; pr < a.n_peers; pr++) a.peer_wires[pr][off + (size_t)K * N] = val;
            }
            if (a.tap_acc) a.tap_acc[(size_t)job * G * N + (size_t)g * N + j] = rns_to_int(av[e]);
        }
    }
    cluster_sync_all();                                           // nobody leaves while a neighbour may still address its shared memory
}

// ------------------------------------------------------------------------------------------------------
// K2s: cluster-split AND prime-split blind rotation.  Same decomposition of one bootstrap over C CTAs as
// k_blind_rotate_cl, but the two residues (mod p1, mod p2) of every ring element live in two THREADS (adjacent lanes) instead
// of one: 2x the threads (two warps per scheduler at C = 4), every thread does the butterflies and point-wise products of one
// prime only.  The transforms, point-wise products and exchanges are independent per prime; only the gadget digit (needs the CRT
// of both residues: one shuffle with the partner lane) and the sample extraction couple them.  Shared-memory data stays in
// the packed layout (mod p1 | mod p2 << 32): a thread touches its 32-bit half, a warp = 16 positions x 2 primes touches the same
// 128 bytes a half-warp of the packed kernels does (conflict-free with the same swizzle).  Bit-identical results.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_async_u32(u32 remote_addr, u32 v, u32 remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(v), "r"(remote_bar) : "memory");
}
template <int LOGN, int K, int M, int LOGC, int OCC = 1>
__global__ void __launch_bounds__(2 * (K + 1) * (1 << (LOGN - LOGC)) / 8, OCC) k_blind_rotate_cs(BRArgs a)
{
    using Cf = BRCCfg<LOGN, K, M, LOGC, OCC>;
    using P = NttPlan<Cf::LOGNS>;
    constexpr int NC = Cf::NC, N = Cf::N, G = Cf::G, C = Cf::C, Ns = Cf::Ns, Ts = Cf::Ts, R = Cf::R, T = Cf::T, RING = Cf::RING, LOGNS = Cf::LOGNS;
    constexpr int THREADS = 2 * Cf::THREADS;
    static_assert(Ts >= 16 && Ts % 16 == 0, "a warp holds 16 thread positions x 2 primes");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, l = tid & 1, g = (tid >> 1) / Ts, tau = (tid >> 1) % Ts, lane = tid & 31;
    const u32 pr = l ? FQ_P2 : FQ_P1, pinv = l ? FQ_P2_INVNEG : FQ_P1_INVNEG, pr2 = 2 * pr;
    const int h = (int)cluster_ctarank();
    const int tau_g = h * Ts + tau;
    u64 *S = (u64 *)smem_raw;
    u64 *INF = S + Cf::s_w, *INI = INF + Cf::inbox_w;
#if FBS_PSI_STATIC
    __shared__ __align__(16) u64 PSI[2 * N];
    u64 *TWF = INI + Cf::inbox_w, *TWI = TWF + Cf::tw_w;
#else
    u64 *PSI = INI + Cf::inbox_w;
    u64 *TWF = PSI + Cf::psi_w, *TWI = TWF + Cf::tw_w;
#endif
    u64 *RNG = TWI + Cf::tw_w;
    u64 *full = RNG + (size_t)RING * Cf::slice_w, *empty = full + RING, *xbar = empty + RING;
    u16 *s_ms = (u16 *)(xbar + 2);
    const int n = a.n, p = a.p;
    const int n_pairs = (n + M - 1) / M, n_slices = 8 * n_pairs;

    const long long job = a.job_begin + (long long)(blockIdx.x >> LOGC);
    const int node = a.grp_first ? a.grp_first[a.node_begin + (int)(job / a.B)] : a.node_begin + (int)(job / a.B);
    const long long inst = job % a.B;
    {
        const u16 *ms = a.ms + ((size_t)(a.bs_lc[node] - a.lc_begin) * a.B + inst) * (size_t)(n + 1);
        for (int i = tid; i <= n; i += THREADS) s_ms[i] = ms[i];
    }
    const int tab0 = a.bs_tab_ptr[node], tabL = a.bs_tab_ptr[node + 1] - tab0, mode = a.bs_mode[node];
    constexpr bool fast_psi = (LOGN - 2 >= 9) && (LOGN + 1 <= 12);                    // as in k_blind_rotate2
    auto psw = [](u32 x) { return fast_psi ? x ^ ((x >> 5) & 15u) : x ^ (((x >> 4) ^ (x >> 8)) & 15u); };
    for (int i = tid; i < 2 * N; i += THREADS) PSI[psw((u32)i)] = a.psi_pow[i];
    for (int i = 1 + tid; i < Ns; i += THREADS) {
        ((uint4 *)TWF)[i] = __ldg((const uint4 *)a.psi_rev + ntt_local_src(i, C + h));
        ((uint4 *)TWI)[i] = __ldg((const uint4 *)a.psi_rev + ntt_local_src(i, 2 * C - 1 - h));
    }
    u32 cw[C], cws[C];                                            // cross-stage twiddles psi_rev[1 .. C) of this thread's prime
#pragma unroll
    for (int i = 1; i < C; i++) fq_tw_load1<false>(a.psi_rev + i, l, cw[i], cws[i]);
    cw[0] = cw[1]; cws[0] = cws[1];
    if (tid == 0) {
        for (int r = 0; r < RING; r++) { mbar_init(full + r, 1); mbar_init(empty + r, THREADS / 32); }
        mbar_init(xbar, 1); mbar_init(xbar + 1, 1);
    }
    __syncthreads();
    cluster_sync_all();
    auto issue_slice = [&](int sl, int slot) {                    // called by thread 0
        fence_proxy_async();
        mbar_expect_tx(full + slot, (u32)(Cf::slice_w * 8));
        tma_load_1d(RNG + (size_t)slot * Cf::slice_w, a.bsk + ((size_t)sl * C + h) * Cf::slice_w, (u32)(Cf::slice_w * 8), full + slot);
    };
    if (tid == 0) for (int sl = 0; sl < RING && sl < n_slices; sl++) issue_slice(sl, sl);
    u32 r_inf[C], r_ini[C], r_xf[C], r_xi[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        r_inf[c] = mapa_shared(smem_u32(INF), (u32)c); r_ini[c] = mapa_shared(smem_u32(INI), (u32)c);
        r_xf[c] = mapa_shared(smem_u32(xbar), (u32)c); r_xi[c] = mapa_shared(smem_u32(xbar + 1), (u32)c);
    }
    const u32 half = 4u * (u32)l;                                 // this thread's half of every packed word
    const u32 send_off = (u32)((((g * 8 + h * R) * Ts) + tau) * 8) + half;
    const u32 inbox_bytes = (u32)(Cf::inbox_w * 8);
    // ---- accumulator init: register e = hh*R + ri holds coefficient j = hh*Ns + h*(Ns/C) + ri*Ts + tau of polynomial g, residue of prime l
    u32 av[8];
    {
        const u64 delta = fbs_delta(p), off = fq_mul((u64)mode, delta >> 1);
        const int bt = s_ms[n];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int j = (e / R) * Ns + h * (Ns / C) + (e % R) * Ts + tau;
            u64 val = 0;
            if (g == K) {
                int src = j + bt;
                bool neg = false;
                if (src >= 2 * N) src -= 2 * N;
                if (src >= N) { src -= N; neg = true; }
                int x = (int)((2LL * src * p + N) / (2LL * N));
                if (x >= p) { x -= p; neg = neg != (a.grp_first == nullptr); }
                const u64 tvx = (x < tabL) ? (u64)__ldg(a.bs_tab + tab0 + x) : 0;
                const u64 F = a.grp_first ? fbs_half_delta(p) : fq_sub(fq_mul(tvx, delta), off);
                val = neg ? fq_neg(F) : F;
            }
            av[e] = (u32)(val % pr);
        }
    }
    const int bar_g = 1 + g;
    auto gsync = [bar_g] { bar_sync_named(bar_g, 2 * Ts); };
    const int bar_x = 1 + G + (tau >> 4);
    auto xsync = [bar_x] { bar_sync_named(bar_x, G * 32); };
    static_assert(1 + G + Ts / 16 <= 16, "named barriers");
    const int beta = a.beta;
    const u64 rc = 1ULL << (62 - beta);
    constexpr int SH = LOGNS + 3;
    unsigned char *Sb = (unsigned char *)S + half;
    u32 bo[LOGNS];
#pragma unroll
    for (int lb = 0; lb < LOGNS; lb++) bo[lb] = P::tau_boff(tau, lb) | ((u32)g << SH);
    static_assert(P::idx(1, 1, P::inv_lb(P::NPASS - 1)) == 1 + Ts && P::idx(1, 1, P::fwd_lb(0)) == 1 + Ts, "local transforms start / end with register e at local index tau + e*Ts");
    const u32 odd0 = 2u * (__brev((u32)tau_g) >> (32 - (LOGN - 3))) + 1u;
    u32 koff[G];
#pragma unroll
    for (int og = 0; og < G; og++) { int gg = g + og; if (gg >= G) gg -= G; koff[og] = (u32)((((tau >> 5) * NC * G * G + gg * G + g) * 32 + (tau & 31)) * 8) + half; }
    int slot = 0; u32 par = 0;
    const fq_tw *twf = (const fq_tw *)TWF, *twi = (const fq_tw *)TWI;
    const unsigned char *INFb = (const unsigned char *)INF + half, *INIb = (const unsigned char *)INI + half, *PSIb = (const unsigned char *)PSI + half;

    for (int t = 0; t < n_pairs; t++) {
        const u32 xpar = (u32)(t & 1);
        // ---- decompose (the digit needs both residues: one shuffle with the partner lane), cross butterflies, forward exchange
        u32 dg[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 other = __shfl_xor_sync(0xffffffffu, av[e], 1);
            rns2 v; v.a = l ? other : av[e]; v.b = l ? av[e] : other;
            const u32 d = (u32)fbs_digit1_t(rns_crt_hi(v), v.a, beta, rc);
            dg[e] = d + pr;
        }
        ntt_cross_fwd_1p<LOGC>(dg, cw, cws, pr, a.zero);
        if (tid == 0) mbar_expect_tx(xbar, inbox_bytes);
#pragma unroll
        for (int e = 0; e < 8; e++) st_async_u32(r_inf[e / R] + send_off + (u32)((e % R) * Ts * 8), dg[e], r_xf[e / R]);
        mbar_wait_cluster(xbar, xpar);
#pragma unroll
        for (int e = 0; e < 8; e++) dg[e] = *(const u32 *)(INFb + ((size_t)(g * 8 + e) * Ts + tau) * 8);
        ntt_fwd1p_from<LOGNS, 0, 4, decltype(gsync), true>(dg, tau, Sb, bo, twf, l, pr, gsync, true, a.zero);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            dg[e] = r32_fold(dg[e], pr2);
            *(u32 *)(Sb + (bo[0] ^ P::elem_boff(e, 0))) = dg[e];
        }
        u32 x[8];
        auto release_slot = [&](int e) {
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + slot)) : "memory");
            if (tid == 0) {
                const int nxt = 8 * t + e + RING;
                if (nxt < n_slices) { mbar_wait(empty + slot, par); issue_slice(nxt, slot); }
            }
            if (++slot == RING) { slot = 0; par ^= 1; }
        };
        u32 PK[NC];
        {
            u32 ai[M];
#pragma unroll
            for (int i = 0; i < M; i++) ai[i] = (M * t + i < n) ? s_ms[M * t + i] : 0u;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                u32 E = 0;
#pragma unroll
                for (int i = 0; i < M; i++) if ((fbs_unroll_mask(M, c) >> i) & 1) E += ai[i];
                const u32 x0 = (E * odd0) & (2 * N - 1);
                if constexpr (fast_psi) PK[c] = (8u * psw(x0)) | (E & 7u);                  // see k_blind_rotate2
                else PK[c] = x0 | ((E & 7u) << 20);
            }
        }
#if !FBS_LATE_XSYNC
        xsync();
#endif
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 o = bo[0] ^ P::elem_boff(e, 0);
            constexpr int BR3[8] = {0, 4, 2, 6, 1, 5, 3, 7};
            if (!FBS_EARLY_RELEASE || e == 0) mbar_wait(full + slot, par);      // elements 1..7: waited for one element ahead (see k_blind_rotate2)
            const unsigned char *ks = (const unsigned char *)(RNG + (size_t)slot * Cf::slice_w);
            auto factor = [&](int c) -> u32 {
                const u32 pk = PK[c];
                if constexpr (fast_psi) {
                    const u32 off = (pk * (1u + ((u32)BR3[e] << (LOGN + 1)))) & ((2u * N - 1u) << 3);
                    return *(const u32 *)(PSIb + off);
                } else {
                    const u32 xi = ((pk & 0xFFFFu) + (((u32)BR3[e] * (pk >> 20)) << (LOGN - 2))) & (2 * N - 1);
                    return *(const u32 *)(PSIb + 8u * psw(xi));
                }
            };
            u64 pa[G];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const u32 f = factor(c);
#pragma unroll
                for (int og = 0; og < G; og++) {
                    const u32 kk = *(const u32 *)(ks + koff[og] + (size_t)c * G * G * 32 * 8);
                    pa[og] = (c == 0) ? r32_mulwide(f, kk) : r32_madwide(f, kk, pa[og]);
                }
            }
#if FBS_LATE_XSYNC
            if (e == 0) xsync();                             // partner spectra are needed from here on (see k_blind_rotate2)
#endif
#if FBS_EARLY_RELEASE
            release_slot(e);                                 // key words are in registers: free the entry, wait for the next one now
            if (e < 7) mbar_wait(full + slot, par);
#endif
            u64 oa = 0;
#pragma unroll
            for (int og = 0; og < G; og++) {
                const u32 ba = r32_redc(pa[og], pr, pinv);
                int gg = g + og; if (gg >= G) gg -= G;
                const u32 xg = (og == 0) ? 0u : (G == 2) ? (1u << SH) : ((u32)(g ^ gg) << SH);
                const u32 d = *(const u32 *)(Sb + (o ^ xg));
                oa = (og == 0) ? r32_mulwide(d, ba) : r32_madwide(d, ba, oa);
            }
            x[e] = r32_fold(r32_redc(oa, pr, pinv), pr2);
#if !FBS_EARLY_RELEASE
            release_slot(e);
#endif
        }
        auto after_pass0 = [&] { xsync(); };
        ntt_inv1p_from<LOGNS, 0, 4, decltype(after_pass0), decltype(gsync), true>(x, tau, Sb, bo, twi, l, pr, after_pass0, gsync, a.zero);
        if (tid == 0) mbar_expect_tx(xbar + 1, inbox_bytes);
#pragma unroll
        for (int e = 0; e < 8; e++) st_async_u32(r_ini[e / R] + send_off + (u32)((e % R) * Ts * 8), x[e], r_xi[e / R]);
        mbar_wait_cluster(xbar + 1, xpar);
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = *(const u32 *)(INIb + ((size_t)(g * 8 + e) * Ts + tau) * 8);
        ntt_cross_inv_1p<LOGC>(x, cw, cws, pr, a.zero);
#pragma unroll
        for (int e = 0; e < 8; e++) av[e] = r32_csub(r32_fold(av[e] + x[e] + a.zero, pr2), pr);
    }
    // ---- K3: sample extraction from the registers; the two lanes of a position share the 8 words (even e: lane of p1, odd e: lane of p2)
    {
        const size_t CT = (size_t)K * N + 1;
        const size_t off = ((size_t)a.bs_slot[node] * a.B + inst) * CT;
        u64 *out = a.wires + off;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 other = __shfl_xor_sync(0xffffffffu, av[e], 1);
            if ((e & 1) != l) continue;
            rns2 v; v.a = l ? other : av[e]; v.b = l ? av[e] : other;
            const int j = (e / R) * Ns + h * (Ns / C) + (e % R) * Ts + tau;
            if (a.grp_first) {
            } else if (g < K) {
                const int jj = (j == 0) ? 0 : N - j;
                const u64 val = rns_to_int(j == 0 ? v : rns_neg(v));
                out[(size_t)g * N + jj] = val;
                for (int prr = 0; prr < a.n_peers; prr++) a.peer_wires[prr][off + (size_t)g * N + jj] = val;
            } else if (j == 0) {
                const u64 val = fq_add(rns_to_int(v), fq_mul((u64)mode, fbs_delta(p) >> 1));
                out[(size_t)K * N] = val;
                for (int prr = 0; prr < a.n_peers; prr++) a.peer_wires[prr][off + (size_t)K * N] = val;
            }
            if (a.tap_acc) a.tap_acc[(size_t)job * G * N + (size_t)g * N + j] = rns_to_int(v);
        }
    }
    cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------------
// Multi-value bootstrap, finishing step (oracle/tfhe_ref.c: ref_multi_extract).  The blind rotation of a
// (group, instance) job left ACC = GLWE(X^-mu * TV0), TV0 = H (1 + X + .. + X^(N-1)), as integers mod q in acc[job][(K+1)][N].
// For every table f of the group: out_f = SampleExtract(ACC * e_f) + s*H, where e_f is the sparse polynomial with
// Delta e_f = (1 - X) TV_f: at most p non-zero coefficients in {-2..2} at the slot boundaries j_x = ceil(N (2x - 1) / 2p).
// One CTA per job; the sums are at most p terms per output word.
// ------------------------------------------------------------------------------------------------------
struct MVArgs {
    const u64 *acc; const int32_t *grp_first, *bs_slot, *bs_tab_ptr, *bs_mode; const u8 *bs_tab;
    u64 *wires; long long B; int grp_begin, N, K, p;
    int n_peers; u64 *peer_wires[8];
};
__global__ void __launch_bounds__(256) k_multi_extract(MVArgs a)
{
    __shared__ int s_pos[128], s_coef[128], s_cnt;
    const long long job = blockIdx.x;
    const int gi = a.grp_begin + (int)(job / a.B), N = a.N, K = a.K, p = a.p;
    const long long inst = job % a.B;
    const u64 *acc = a.acc + (size_t)job * (K + 1) * N;
    const size_t CT = (size_t)K * N + 1;
    for (int q = a.grp_first[gi]; q < a.grp_first[gi + 1]; q++) {
        const int t0 = a.bs_tab_ptr[q], L = a.bs_tab_ptr[q + 1] - t0, s = a.bs_mode[q];
        __syncthreads();
        if (threadIdx.x == 0) {
            int cnt = 0;
            for (int x = 1; x <= p; x++) {
                const int tx = (x < p && x < L) ? a.bs_tab[t0 + x] : 0, tp = (x - 1 < L) ? a.bs_tab[t0 + x - 1] : 0;
                const int e = (x < p) ? tx - tp : -((int)a.bs_tab[t0] + tp - s);
                if (!e) continue;
                s_pos[cnt] = (int)(((long long)N * (2 * x - 1) + 2 * p - 1) / (2 * p)); s_coef[cnt] = e; cnt++;
            }
            s_cnt = cnt;
        }
        __syncthreads();
        const int cnt = s_cnt;
        const size_t off = ((size_t)a.bs_slot[q] * a.B + inst) * CT;
        for (int w = threadIdx.x; w <= K * N; w += 256) {
            const int u = w / N, jj = w % N;                       // w == K*N: the body = coefficient 0 of polynomial K
            const int i = (jj == 0) ? 0 : N - jj;                  // extracted mask word jj is -(ACC_u e_f)[N - jj] (jj > 0)
            u64 sum = 0;
            for (int t = 0; t < cnt; t++) {
                const int idx = i - s_pos[t];
                const u64 x = idx >= 0 ? acc[(size_t)u * N + idx] : fq_neg(acc[(size_t)u * N + idx + N]);
                const int cf = s_coef[t];
                const u64 term = (cf == 2 || cf == -2) ? fq_add(x, x) : x;
                sum = cf > 0 ? fq_add(sum, term) : fq_sub(sum, term);
            }
            u64 val;
            if (w == K * N) val = fq_add(sum, fq_mul((u64)s, fbs_half_delta(p)));
            else val = (jj == 0) ? sum : fq_neg(sum);
            a.wires[off + w] = val;
            for (int pr = 0; pr < a.n_peers; pr++) a.peer_wires[pr][off + w] = val;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Device-side level hand-off for node-sharded multi-GPU runs.  Every rank's wire replica ends in a
// flag page: flags[r] = number of levels rank r has completed.  After a level's blind rotation (whose epilogue stored the
// outputs into every peer replica) k_level_signal publishes the new epoch into every peer's page with a system-scope
// release; before the next level's lincomb k_level_wait spins with system-scope acquires until every rank has published
// it.  The spin is bounded (FBS_SYNC_TIMEOUT_NS): on timeout it raises *err and lets the stream continue, so a lost peer
// shows up as an error code instead of a hung GPU.
// ------------------------------------------------------------------------------------------------------
#define FBS_FLAG_PAGE 4096
#ifndef FBS_SYNC_TIMEOUT_NS
#define FBS_SYNC_TIMEOUT_NS 20000000000ULL
#endif
struct LevelPeers { int n; u64 *flags[8]; };
__global__ void k_level_signal(LevelPeers lp, int rank, u64 epoch)
{
    __threadfence_system();                       // the preceding kernels' peer stores are ordered before the flags
    if ((int)threadIdx.x < lp.n)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(lp.flags[threadIdx.x] + rank), "l"(epoch) : "memory");
}
__global__ void k_level_wait(const u64 *flags, int world, int rank, u64 epoch, int *err)
{
    const int r = threadIdx.x;
    if (r < world && r != rank) {
        u64 t0, t1, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
            if (v >= epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > FBS_SYNC_TIMEOUT_NS) { atomicExch(err, 1 + r); break; }
            __nanosleep(200);
        }
    }
    __syncwarp();
    __threadfence_system();
}

// ------------------------------------------------------------------------------------------------------
// cleartext evaluator: the reference's hot loop (fbs_exec_env.py:215-220) as one thread per instance
// wire values live in vals[slot][B] (uint8) so that neighbouring threads touch neighbouring bytes
// ------------------------------------------------------------------------------------------------------
struct ClearArgs {
    const int32_t *lc_level_ptr, *bs_level_ptr, *lc_ptr, *lc_slot, *lc_coef, *lc_const, *bs_lc, *bs_slot, *bs_tab_ptr, *in_slot;
    const u8 *bs_tab;
    const int32_t *out_ptr, *out_slot, *out_coef, *out_const;
    const u8 *in; u8 *vals; u8 *out; int32_t *lcv;     // lcv[n_lincombs_max_per_level][B] scratch
    int *err;
    long long B; int n_inputs, n_levels, n_outputs, lc_stride;
};
__global__ void __launch_bounds__(256) k_clear_eval(ClearArgs a)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    for (int i = 0; i < a.n_inputs; i++) a.vals[(size_t)a.in_slot[i] * a.B + b] = a.in[(size_t)i * a.B + b];
    for (int lv = 0; lv < a.n_levels; lv++) {
        const int l0 = a.lc_level_ptr[lv], l1 = a.lc_level_ptr[lv + 1];
        for (int q = l0; q < l1; q++) {
            int acc = a.lc_const[q];
            for (int o = a.lc_ptr[q]; o < a.lc_ptr[q + 1]; o++) acc += a.lc_coef[o] * (int)a.vals[(size_t)a.lc_slot[o] * a.B + b];
            a.lcv[(size_t)(q - l0) * a.B + b] = acc;
        }
        for (int q = a.bs_level_ptr[lv]; q < a.bs_level_ptr[lv + 1]; q++) {
            const int idx = a.lcv[(size_t)(a.bs_lc[q] - l0) * a.B + b];
            const int t0 = a.bs_tab_ptr[q], len = a.bs_tab_ptr[q + 1] - t0;
            u8 v = 0;
            if (idx < 0 || idx >= len) atomicExch(a.err, 1 + q); else v = a.bs_tab[t0 + idx];
            a.vals[(size_t)a.bs_slot[q] * a.B + b] = v;
        }
    }
    for (int q = 0; q < a.n_outputs; q++) {
        int acc = a.out_const[q];
        for (int o = a.out_ptr[q]; o < a.out_ptr[q + 1]; o++) acc += a.out_coef[o] * (int)a.vals[(size_t)a.out_slot[o] * a.B + b];
        a.out[(size_t)q * a.B + b] = (u8)acc;
    }
}
