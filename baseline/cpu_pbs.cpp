// cpu_pbs.cpp -- TUNED CPU arm of the benchmark (bench.py --impl reference, cpu_baseline): the same encrypted evaluation the
// GPU executor runs (levelised FBS program; per bootstrap: LWE lincomb -> key switch -> modulus switch -> key-unrolled
// blind rotation -> sample extract), written for host cores.
//
// WHY THIS FILE EXISTS.  The reference (ssmiler/tfhe_fbs_map) has no encrypted executor, and zama-ai/concrete's CPU PBS
// cannot be built offline (no Rust toolchain / sources).  oracle/tfhe_ref.c is the parity CHECKER: deliberately slow and
// obvious (every modular multiply is a 128-bit `%`), so a GPU/oracle ratio flatters the GPU.  This file is the honest CPU
// number to put next to the GPU's: residue number system over the two 30-bit NTT primes, Harvey lazy butterflies with
// Shoup twiddles (auto-vectorised by -O3 -march=native), Montgomery point-wise products with lazy 64-bit accumulation,
// the key-switch as 32-bit multiply-adds on the byte-free residues, three key bits per blind-rotation step, OpenMP over
// the independent instances.  It is NOT the product (the product has no CPU path) and NOT the checker.
//
// It computes EXACTLY the function of the specification: it shares the scalar helpers of
// tfhe_fbs_map_b200/csrc/{fq,common}.cuh (host builds of the same inline functions: rounding, digits, CRT, PRNG), so its
// ciphertexts are bit-identical to the oracle's and the CUDA path's (tests/test_cpu_arm.py) -- the speed-up it shows over
// the oracle is implementation, not a different algorithm.  Keys are taken from the oracle's seeded key generation
// (coefficient domain) and pre-processed here.
//
// Build: g++ -O3 -march=native -fopenmp -shared -fPIC (baseline/Makefile).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../tfhe_fbs_map_b200/csrc/fq.cuh"
#include "../tfhe_fbs_map_b200/csrc/common.cuh"

typedef uint8_t u8;
typedef uint16_t u16;

struct cpu_params { int32_t n, k, N, bsk_l, bsk_beta, ks_l, ks_beta, bsk_unroll; uint64_t lwe_noise, glwe_noise; };
struct cpu_prog_desc {                 // same flat descriptor as include/fbs_b200.h
    int32_t p, n_inputs, n_lincombs, n_boots, n_levels, n_slots, n_outputs, contiguous_levels;
    const int32_t *lc_level_ptr, *bs_level_ptr;
    const int32_t *lc_ptr, *lc_slot, *lc_coef, *lc_const;
    const int32_t *bs_lc, *bs_slot, *bs_tab_ptr; const u8 *bs_tab; const int32_t *bs_mode;
    const int32_t *in_slot;
    const int32_t *out_ptr, *out_slot, *out_coef, *out_const;
};

static const u32 PRM[2] = {FQ_P1, FQ_P2};
static const u32 PINVNEG[2] = {FQ_P1_INVNEG, FQ_P2_INVNEG};

struct cpu_ctx {
    cpu_params P; int logN, M, NC, n_groups;
    std::vector<u8> s_big;
    // per prime: twiddles psi^brev(i) (+ Shoup), inverse ones, psi^x - 1 (x < 2N)
    std::vector<u32> w[2], ws[2], wi[2], wis[2], psipow[2];
    std::vector<u32> odd;                      // 2 brev(i) + 1: evaluation exponent of spectrum position i
    // bootstrapping key, NTT domain, Montgomery form twice x 1/N: [group][c][u][v][N] per prime
    std::vector<u32> bsk[2];
    // key-switching key residues [r][n+1] per prime, column sums (integers mod q)
    std::vector<u32> ksk[2];
    std::vector<u64> colsum;
};

static u32 bitrev32(u32 x, int bits) { u32 r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }
static inline u32 mulmod32(u32 a, u32 b, u32 p) { return (u32)((u64)a * b % p); }

// ---- negacyclic NTT per prime: forward Cooley-Tukey (natural in, bit-reversed out, values in [0,4p)),
//      inverse Gentleman-Sande (bit-reversed in, natural out, [0,2p)), no 1/N (folded into the key) -----------------------
static inline void bf_fwd(u32 &x, u32 &y, u32 W, u32 Ws, u32 p, u32 p2)
{
    u32 u = x; const u32 uf = u - p2; u = uf < u ? uf : u;
    const u32 q = (u32)(((u64)Ws * y) >> 32), v = W * y - q * p;
    x = u + v; y = u - v + p2;
}
static inline void bf_inv(u32 &x, u32 &y, u32 W, u32 Ws, u32 p, u32 p2)
{
    const u32 u = x, v = y;
    u32 s = u + v; const u32 sf = s - p2; s = sf < s ? sf : s;
    const u32 d = u - v + p2, q = (u32)(((u64)Ws * d) >> 32);
    x = s; y = W * d - q * p;
}
static void ntt_fwd(u32 *__restrict__ a, const u32 *__restrict__ w, const u32 *__restrict__ ws, u32 p, int N)
{
    const u32 p2 = 2 * p;
    int t = N, m = 1;
    for (; t > 8; m <<= 1) {                                    // stages with >= 8 contiguous butterflies per twiddle: vectorised
        t >>= 1;
        for (int i = 0; i < m; i++) {
            const u32 W = w[m + i], Ws = ws[m + i];
            u32 *__restrict__ x = a + 2 * i * t, *__restrict__ y = x + t;
            for (int j = 0; j < t; j++) {
                u32 u = x[j]; const u32 uf = u - p2; u = uf < u ? uf : u;
                const u32 yv = y[j], q = (u32)(((u64)Ws * yv) >> 32), v = W * yv - q * p;
                x[j] = u + v; y[j] = u - v + p2;
            }
        }
    }
    // last three stages (t = 4, 2, 1) fused: one radix-8 butterfly network per block of 8 contiguous values, in registers
    const int m4 = N / 8, m2 = N / 4, m1 = N / 2;
    for (int b = 0; b < N / 8; b++) {
        u32 *__restrict__ x = a + 8 * b;
        u32 v0 = x[0], v1 = x[1], v2 = x[2], v3 = x[3], v4 = x[4], v5 = x[5], v6 = x[6], v7 = x[7];
        { const u32 W = w[m4 + b], Ws = ws[m4 + b]; bf_fwd(v0, v4, W, Ws, p, p2); bf_fwd(v1, v5, W, Ws, p, p2); bf_fwd(v2, v6, W, Ws, p, p2); bf_fwd(v3, v7, W, Ws, p, p2); }
        { const u32 W = w[m2 + 2 * b], Ws = ws[m2 + 2 * b]; bf_fwd(v0, v2, W, Ws, p, p2); bf_fwd(v1, v3, W, Ws, p, p2); }
        { const u32 W = w[m2 + 2 * b + 1], Ws = ws[m2 + 2 * b + 1]; bf_fwd(v4, v6, W, Ws, p, p2); bf_fwd(v5, v7, W, Ws, p, p2); }
        bf_fwd(v0, v1, w[m1 + 4 * b], ws[m1 + 4 * b], p, p2); bf_fwd(v2, v3, w[m1 + 4 * b + 1], ws[m1 + 4 * b + 1], p, p2);
        bf_fwd(v4, v5, w[m1 + 4 * b + 2], ws[m1 + 4 * b + 2], p, p2); bf_fwd(v6, v7, w[m1 + 4 * b + 3], ws[m1 + 4 * b + 3], p, p2);
        x[0] = v0; x[1] = v1; x[2] = v2; x[3] = v3; x[4] = v4; x[5] = v5; x[6] = v6; x[7] = v7;
    }
}
static void ntt_inv(u32 *__restrict__ a, const u32 *__restrict__ wi, const u32 *__restrict__ wis, u32 p, int N)
{
    const u32 p2 = 2 * p;
    const int m4 = N / 8, m2 = N / 4, m1 = N / 2;
    for (int b = 0; b < N / 8; b++) {                           // first three stages (t = 1, 2, 4) fused
        u32 *__restrict__ x = a + 8 * b;
        u32 v0 = x[0], v1 = x[1], v2 = x[2], v3 = x[3], v4 = x[4], v5 = x[5], v6 = x[6], v7 = x[7];
        bf_inv(v0, v1, wi[m1 + 4 * b], wis[m1 + 4 * b], p, p2); bf_inv(v2, v3, wi[m1 + 4 * b + 1], wis[m1 + 4 * b + 1], p, p2);
        bf_inv(v4, v5, wi[m1 + 4 * b + 2], wis[m1 + 4 * b + 2], p, p2); bf_inv(v6, v7, wi[m1 + 4 * b + 3], wis[m1 + 4 * b + 3], p, p2);
        { const u32 W = wi[m2 + 2 * b], Ws = wis[m2 + 2 * b]; bf_inv(v0, v2, W, Ws, p, p2); bf_inv(v1, v3, W, Ws, p, p2); }
        { const u32 W = wi[m2 + 2 * b + 1], Ws = wis[m2 + 2 * b + 1]; bf_inv(v4, v6, W, Ws, p, p2); bf_inv(v5, v7, W, Ws, p, p2); }
        { const u32 W = wi[m4 + b], Ws = wis[m4 + b]; bf_inv(v0, v4, W, Ws, p, p2); bf_inv(v1, v5, W, Ws, p, p2); bf_inv(v2, v6, W, Ws, p, p2); bf_inv(v3, v7, W, Ws, p, p2); }
        x[0] = v0; x[1] = v1; x[2] = v2; x[3] = v3; x[4] = v4; x[5] = v5; x[6] = v6; x[7] = v7;
    }
    int t = 8;
    for (int m = N >> 4; m >= 1; m >>= 1) {
        for (int i = 0; i < m; i++) {
            const u32 W = wi[m + i], Ws = wis[m + i];
            u32 *__restrict__ x = a + 2 * i * t, *__restrict__ y = x + t;
            for (int j = 0; j < t; j++) {
                const u32 u = x[j], v = y[j];
                u32 s = u + v; const u32 sf = s - p2; s = sf < s ? sf : s;
                const u32 d = u - v + p2, q = (u32)(((u64)Ws * d) >> 32);
                x[j] = s; y[j] = W * d - q * p;
            }
        }
        t <<= 1;
    }
}
static inline u32 redc64(u64 acc, u32 p, u32 pinv_neg) { return r32_redc(acc, p, pinv_neg); }      // acc * 2^-32 mod p, in (0, acc/2^32 + p]
static inline int unroll_mask(int m, int c) { return m == 2 ? (c == 0 ? 3 : c) : c + 1; }           // key-bit subset of GGSW c of a group

extern "C" void cpu_ctx_destroy(cpu_ctx *c) { delete c; }

// s_big [kN] bits; ksk [kN*ks_l][n+1] integers mod q; bsk_coef [n_ggsw][(k+1)l][k+1][N] integers mod q (oracle layout)
extern "C" cpu_ctx *cpu_ctx_create(const cpu_params *Pp, const u8 *s_big, const u64 *ksk, const u64 *bsk_coef)
{
    const cpu_params &P = *Pp;
    if (P.k != 1 || P.bsk_l != 1 || !(P.bsk_unroll == 2 || P.bsk_unroll == 3) || P.ks_l > 8) return nullptr;   // the shape this arm is tuned for
    cpu_ctx *c = new cpu_ctx();
    c->P = P; c->M = P.bsk_unroll; c->NC = (1 << c->M) - 1; c->n_groups = (P.n + c->M - 1) / c->M;
    const int N = P.N; int logN = 0; while ((1 << logN) < N) logN++; c->logN = logN;
    c->s_big.assign(s_big, s_big + (size_t)P.k * N);
    u32 mont2_ninv[2];
    for (int l = 0; l < 2; l++) {
        const u32 p = PRM[l];
        const u64 psi = pow_mod_host(3, (p - 1) / (2ULL * N), p), psi_inv = pow_mod_host(psi, p - 2, p);
        c->w[l].resize(N); c->ws[l].resize(N); c->wi[l].resize(N); c->wis[l].resize(N); c->psipow[l].resize(2 * N);
        for (int i = 0; i < N; i++) {
            const u32 r = bitrev32((u32)i, logN);
            c->w[l][i] = (u32)pow_mod_host(psi, r, p); c->ws[l][i] = shoup32_host(c->w[l][i], p);
            c->wi[l][i] = (u32)pow_mod_host(psi_inv, r, p); c->wis[l][i] = shoup32_host(c->wi[l][i], p);
        }
        u64 x = 1;
        for (int e = 0; e < 2 * N; e++) { c->psipow[l][e] = (u32)(x - 1); x = x * psi % p; }      // psi^e - 1 (psi^e >= 1)
        const u32 ninv = (u32)pow_mod_host((u64)N, p - 2, p), r32 = (u32)((1ULL << 32) % p);
        mont2_ninv[l] = mulmod32(mulmod32(r32, r32, p), ninv, p);                                 // 2^64 / N
    }
    c->odd.resize(N);
    for (int i = 0; i < N; i++) c->odd[i] = 2 * bitrev32((u32)i, logN) + 1;
    // ---- bootstrapping key: [ggsw = group*NC + cc][row u][col v][N] -> NTT, x 2^64/N, per prime
    const size_t n_ggsw = (size_t)c->NC * c->n_groups, polys = n_ggsw * 4;
    for (int l = 0; l < 2; l++) c->bsk[l].resize(polys * N);
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < (long long)polys; q++) {
        std::vector<u32> tmp(N);
        for (int l = 0; l < 2; l++) {
            const u32 p = PRM[l];
            for (int j = 0; j < N; j++) tmp[j] = (u32)(bsk_coef[(size_t)q * N + j] % p);
            ntt_fwd(tmp.data(), c->w[l].data(), c->ws[l].data(), p, N);
            u32 *dst = c->bsk[l].data() + (size_t)q * N;
            for (int j = 0; j < N; j++) dst[j] = mulmod32(tmp[j] % p, mont2_ninv[l], p);
        }
    }
    // ---- key-switching key residues + column sums
    const size_t R = (size_t)P.k * N * P.ks_l, cols = (size_t)P.n + 1;
    for (int l = 0; l < 2; l++) c->ksk[l].resize(R * cols);
    c->colsum.assign(cols, 0);
    for (size_t r = 0; r < R; r++)
        for (size_t cc = 0; cc < cols; cc++) {
            const u64 x = ksk[r * cols + cc];
            c->ksk[0][r * cols + cc] = (u32)(x % FQ_P1); c->ksk[1][r * cols + cc] = (u32)(x % FQ_P2);
            c->colsum[cc] = fq_add(c->colsum[cc], x);
        }
    return c;
}

// ---- one bootstrap: big LWE `in` (after the lincomb) -> big LWE `out` -----------------------------------------------------
struct Scratch {
    std::vector<u32> acc[2][2], dig[2][2], outp[2][2];      // [poly][prime][N]
    std::vector<u64> ksacc[2];                              // [prime][n+1]
    std::vector<u16> ms;
    std::vector<u64> bun[4];                                // bundle accumulators [u][v][N]
    std::vector<u32> fac;                                   // one factor over all spectrum positions
    Scratch(int N, int n) { for (int g = 0; g < 2; g++) for (int l = 0; l < 2; l++) { acc[g][l].resize(N); dig[g][l].resize(N); outp[g][l].resize(N); }
                            ksacc[0].resize(n + 1); ksacc[1].resize(n + 1); ms.resize(n + 1); for (auto &b : bun) b.resize(N); fac.resize(N); }
};
template <int LK>
static void keyswitch_modswitch(const cpu_ctx *c, const u64 *in, Scratch &S)
{
    const cpu_params &P = c->P; const int D = P.k * P.N, n = P.n, cols = n + 1, halfB = 1 << (P.ks_beta - 1);
    u64 *__restrict__ a0 = S.ksacc[0].data(), *__restrict__ a1 = S.ksacc[1].data();
    memset(a0, 0, sizeof(u64) * cols); memset(a1, 0, sizeof(u64) * cols);
    for (int i = 0; i < D; i++) {
        int d[LK];
        fbs_balanced_digits<LK>(fbs_round_top(in[i], P.ks_beta * LK), P.ks_beta, d);
        for (int j = 0; j < LK; j++) {
            const u32 du = (u32)(d[j] + halfB);                                    // offset form, [0, B)
            if (!du) continue;
            const u32 *__restrict__ k0 = c->ksk[0].data() + ((size_t)i * LK + j) * cols, *__restrict__ k1 = c->ksk[1].data() + ((size_t)i * LK + j) * cols;
            for (int cc = 0; cc < cols; cc++) { a0[cc] += (u64)du * k0[cc]; a1[cc] += (u64)du * k1[cc]; }     // < 2^47: no reduction needed
        }
    }
    for (int cc = 0; cc < cols; cc++) {
        rns2 v; v.a = (u32)(a0[cc] % FQ_P1); v.b = (u32)(a1[cc] % FQ_P2);
        u64 val = fq_sub(fq_mul(c->colsum[cc], (u64)halfB), rns_to_int(v));        // - sum_r d_r ksk[r][cc]
        if (cc == n) val = fq_add(val, in[D]);
        S.ms[cc] = (u16)fbs_modswitch(val, c->logN + 1);
    }
}
static void blind_rotate_extract(const cpu_ctx *c, int p, const u8 *table, int L, int mode, Scratch &S, u64 *out)
{
    const cpu_params &P = c->P; const int N = P.N, n = P.n, M = c->M, NC = c->NC, beta = P.bsk_beta;
    const u64 delta = fbs_delta(p), off = fq_mul((u64)mode, delta >> 1), rc = 1ULL << (62 - beta);
    // accumulator (0, X^{-b} TV), canonical residues
    const int bt = S.ms[n];
    for (int j = 0; j < N; j++) {
        int src = j + bt; bool neg = false;
        if (src >= 2 * N) src -= 2 * N;
        if (src >= N) { src -= N; neg = true; }
        int x = (int)((2LL * src * p + N) / (2LL * N));
        if (x >= p) { x -= p; neg = !neg; }
        const u64 F = fq_sub(fq_mul((x < L) ? (u64)table[x] : 0, delta), off), val = neg ? fq_neg(F) : F;
        S.acc[0][0][j] = S.acc[0][1][j] = 0;
        S.acc[1][0][j] = (u32)(val % FQ_P1); S.acc[1][1][j] = (u32)(val % FQ_P2);
    }
    for (int t = 0; t < c->n_groups; t++) {
        // decompose the accumulator itself: one balanced digit per coefficient
        for (int g = 0; g < 2; g++) {
            const u32 *__restrict__ a0 = S.acc[g][0].data(), *__restrict__ a1 = S.acc[g][1].data();
            u32 *__restrict__ d0 = S.dig[g][0].data(), *__restrict__ d1 = S.dig[g][1].data();
            for (int j = 0; j < N; j++) {
                rns2 v; v.a = a0[j]; v.b = a1[j];
                const u32 d = (u32)fbs_digit1_t(rns_crt_hi(v), v.a, beta, rc);
                d0[j] = d + FQ_P1; d1[j] = d + FQ_P2;
            }
            for (int l = 0; l < 2; l++) ntt_fwd(S.dig[g][l].data(), c->w[l].data(), c->ws[l].data(), PRM[l], N);
        }
        // exponents of the NC monomial factors of this key group
        u32 E[7];
        for (int cc = 0; cc < NC; cc++) {
            const int mask = unroll_mask(M, cc);
            u32 e = 0;
            for (int i = 0; i < M; i++) if ((mask >> i) & 1) e += (M * t + i < n) ? S.ms[M * t + i] : 0u;
            E[cc] = e;
        }
        for (int l = 0; l < 2; l++) {
            const u32 pr = PRM[l];
            const u32 *__restrict__ psp = c->psipow[l].data();
            const u32 *__restrict__ key = c->bsk[l].data() + (size_t)t * NC * 4 * N;
            const u32 *__restrict__ od = c->odd.data();
            u32 *__restrict__ D0 = S.dig[0][l].data(), *__restrict__ D1 = S.dig[1][l].data();
            u32 *__restrict__ o0 = S.outp[0][l].data(), *__restrict__ o1 = S.outp[1][l].data();
            u64 *__restrict__ b00 = S.bun[0].data(), *__restrict__ b01 = S.bun[1].data(), *__restrict__ b10 = S.bun[2].data(), *__restrict__ b11 = S.bun[3].data();
            u32 *__restrict__ F = S.fac.data();
            const u32 p2 = 2 * pr, msk = 2 * N - 1;
            // bundle[u][v] = sum_c f_c key_c[u][v] (< 7 p^2 < 2^63): factor gather, then four straight multiply-add streams per factor
            for (int cc = 0; cc < NC; cc++) {
                const u32 e = E[cc];
                for (int i = 0; i < N; i++) F[i] = psp[(e * od[i]) & msk];
                const u32 *__restrict__ k0 = key + (size_t)cc * 4 * N, *__restrict__ k1 = k0 + N, *__restrict__ k2 = k0 + 2 * N, *__restrict__ k3 = k0 + 3 * N;
                if (cc == 0) for (int i = 0; i < N; i++) { const u64 f = F[i]; b00[i] = f * k0[i]; b01[i] = f * k1[i]; b10[i] = f * k2[i]; b11[i] = f * k3[i]; }
                else for (int i = 0; i < N; i++) { const u64 f = F[i]; b00[i] += f * k0[i]; b01[i] += f * k1[i]; b10[i] += f * k2[i]; b11[i] += f * k3[i]; }
            }
            for (int i = 0; i < N; i++) {
                const u32 r00 = redc64(b00[i], pr, PINVNEG[l]), r01 = redc64(b01[i], pr, PINVNEG[l]), r10 = redc64(b10[i], pr, PINVNEG[l]), r11 = redc64(b11[i], pr, PINVNEG[l]);
                u32 x0 = D0[i]; { const u32 y = x0 - p2; x0 = y < x0 ? y : x0; }
                u32 x1 = D1[i]; { const u32 y = x1 - p2; x1 = y < x1 ? y : x1; }
                u32 v0 = redc64((u64)x0 * r00 + (u64)x1 * r10, pr, PINVNEG[l]), v1 = redc64((u64)x0 * r01 + (u64)x1 * r11, pr, PINVNEG[l]);
                { const u32 y = v0 - p2; v0 = y < v0 ? y : v0; } { const u32 y = v1 - p2; v1 = y < v1 ? y : v1; }
                o0[i] = v0; o1[i] = v1;
            }
            ntt_inv(o0, c->wi[l].data(), c->wis[l].data(), pr, N);
            ntt_inv(o1, c->wi[l].data(), c->wis[l].data(), pr, N);
            for (int g = 0; g < 2; g++) {
                u32 *__restrict__ a = S.acc[g][l].data(); const u32 *__restrict__ o = S.outp[g][l].data();
                for (int j = 0; j < N; j++) {
                    u32 s = a[j] + o[j]; { const u32 y = s - p2; s = y < s ? y : s; } { const u32 y = s - pr; s = y < s ? y : s; }
                    a[j] = s;
                }
            }
        }
    }
    // sample extraction of coefficient 0
    for (int j = 0; j < N; j++) {
        rns2 v; v.a = S.acc[0][0][j == 0 ? 0 : N - j]; v.b = S.acc[0][1][j == 0 ? 0 : N - j];
        out[j] = rns_to_int(j == 0 ? v : rns_neg(v));
    }
    rns2 b; b.a = S.acc[1][0][0]; b.b = S.acc[1][1][0];
    out[N] = fq_add(rns_to_int(b), off);
}
static void pbs_one(const cpu_ctx *c, int p, const u64 *in, const u8 *table, int L, int mode, Scratch &S, u64 *out)
{
    switch (c->P.ks_l) {
    case 1: keyswitch_modswitch<1>(c, in, S); break; case 2: keyswitch_modswitch<2>(c, in, S); break; case 3: keyswitch_modswitch<3>(c, in, S); break;
    case 4: keyswitch_modswitch<4>(c, in, S); break; case 5: keyswitch_modswitch<5>(c, in, S); break; case 6: keyswitch_modswitch<6>(c, in, S); break;
    case 7: keyswitch_modswitch<7>(c, in, S); break; default: keyswitch_modswitch<8>(c, in, S); break;
    }
    blind_rotate_extract(c, p, table, L, mode, S, out);
}
// parity hook: `count` independent bootstraps of given ciphertexts (tests/test_cpu_arm.py compares with the oracle bit for bit)
extern "C" void cpu_pbs_batch(const cpu_ctx *c, int p, const u64 *in, const u8 *tables, int tab_stride, const u8 *tlen, const int32_t *modes,
                              int64_t count, u64 *out, int threads)
{
    const size_t CT = (size_t)c->P.k * c->P.N + 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
    {
        Scratch S(c->P.N, c->P.n);
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < count; i++)
            pbs_one(c, p, in + (size_t)i * CT, tables + (size_t)i * tab_stride, tlen[i], modes ? modes[i] : 1, S, out + (size_t)i * CT);
    }
}

// ---- whole program on B instances: encrypt inputs, all levels, decrypt outputs (what fbs_eval_bits does on the GPU) ----------
// max_levels > 0: evaluate only the first max_levels levels and skip the outputs (a bounded sample of a long program: the per-
// bootstrap cost does not depend on the level); returns the number of bootstraps executed per instance.
extern "C" int cpu_eval_prog(const cpu_ctx *c, const cpu_prog_desc *g, const u8 *in, int64_t B, int64_t inst_offset, int64_t B_total,
                             u64 enc_seed, u8 *out, int threads, int max_levels)
{
    const int n_levels = (max_levels > 0 && max_levels < g->n_levels) ? max_levels : g->n_levels;
    const bool truncated = n_levels < g->n_levels;
    const cpu_params &P = c->P; const int D = P.k * P.N, p = g->p; const size_t CT = (size_t)D + 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
    {
        Scratch S(P.N, P.n);
        std::vector<u64> w((size_t)g->n_slots * CT), lc(CT);
        std::vector<u16> msl;
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; b++) {
            for (int i = 0; i < g->n_inputs; i++) {                            // fresh encryptions (same PRNG as every other implementation)
                u64 *ct = w.data() + (size_t)g->in_slot[i] * CT;
                const u64 id = (u64)i * (u64)B_total + (u64)(inst_offset + b);
                u64 body = 0;
                for (int q = 0; q < D; q++) { const u64 x = fbs_rnd_uniform(enc_seed, DOM_ENC_MASK, id * (u64)D + q); ct[q] = x; if (c->s_big[q]) body = fq_add(body, x); }
                body = fq_add(body, fbs_rnd_noise(enc_seed, DOM_ENC_NOISE, id, P.glwe_noise));
                ct[D] = fq_add(body, fq_mul((u64)(in[(size_t)i * B + b] % (2 * p)), fbs_delta(p)));
            }
            for (int lv = 0; lv < n_levels; lv++) {
                // all lincombs of the level first (a bootstrap may recycle the slot of an operand), one key switch per lincomb
                const int l0 = g->lc_level_ptr[lv], l1 = g->lc_level_ptr[lv + 1];
                if ((size_t)(l1 - l0) * (P.n + 1) > msl.size()) msl.resize((size_t)(l1 - l0) * (P.n + 1));
                for (int l = l0; l < l1; l++) {
                    for (size_t i = 0; i < CT; i++) lc[i] = 0;
                    for (int o = g->lc_ptr[l]; o < g->lc_ptr[l + 1]; o++) {
                        const u64 *x = w.data() + (size_t)g->lc_slot[o] * CT; const u64 cf = fq_from_i64(g->lc_coef[o]);
                        for (size_t i = 0; i < CT; i++) lc[i] = fq_add(lc[i], fq_mul(x[i], cf));
                    }
                    lc[D] = fq_add(lc[D], fq_mul(fq_from_i64(g->lc_const[l]), fbs_delta(p)));
                    switch (P.ks_l) {
                    case 1: keyswitch_modswitch<1>(c, lc.data(), S); break; case 2: keyswitch_modswitch<2>(c, lc.data(), S); break;
                    case 3: keyswitch_modswitch<3>(c, lc.data(), S); break; case 4: keyswitch_modswitch<4>(c, lc.data(), S); break;
                    case 5: keyswitch_modswitch<5>(c, lc.data(), S); break; case 6: keyswitch_modswitch<6>(c, lc.data(), S); break;
                    case 7: keyswitch_modswitch<7>(c, lc.data(), S); break; default: keyswitch_modswitch<8>(c, lc.data(), S); break;
                    }
                    memcpy(msl.data() + (size_t)(l - l0) * (P.n + 1), S.ms.data(), sizeof(u16) * (P.n + 1));
                }
                for (int q = g->bs_level_ptr[lv]; q < g->bs_level_ptr[lv + 1]; q++) {
                    memcpy(S.ms.data(), msl.data() + (size_t)(g->bs_lc[q] - l0) * (P.n + 1), sizeof(u16) * (P.n + 1));
                    blind_rotate_extract(c, p, g->bs_tab + g->bs_tab_ptr[q], g->bs_tab_ptr[q + 1] - g->bs_tab_ptr[q], g->bs_mode[q], S,
                                         w.data() + (size_t)g->bs_slot[q] * CT);
                }
            }
            for (int q = 0; q < (truncated ? 0 : g->n_outputs); q++) {           // outputs are lincombs of wires: decrypt phase, decode
                u64 mask = 0, body = 0;
                for (int o = g->out_ptr[q]; o < g->out_ptr[q + 1]; o++) {
                    const u64 *ct = w.data() + (size_t)g->out_slot[o] * CT; const u64 cf = fq_from_i64(g->out_coef[o]);
                    u64 acc = 0;
                    for (int i = 0; i < D; i++) if (c->s_big[i]) acc = fq_add(acc, ct[i]);
                    mask = fq_add(mask, fq_mul(acc, cf)); body = fq_add(body, fq_mul(ct[D], cf));
                }
                body = fq_add(body, fq_mul(fq_from_i64(g->out_const[q]), fbs_delta(p)));
                out[(size_t)q * B + b] = (u8)fbs_decode(fq_sub(body, mask), p);
            }
        }
    }
    return g->bs_level_ptr[n_levels];
}
extern "C" int cpu_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
