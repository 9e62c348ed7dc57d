/*
 * oracle/tfhe_ref.c -- CPU restatement of the encrypted FBS pipeline.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (tfhe_fbs_map_b200/) never imports, links or executes it.
 *
 * What it restates
 * ----------------
 * The reference (ssmiler/tfhe_fbs_map) contains NO encrypted execution (SURVEY.md section 8(a) row
 * a10): its only executable semantics for a mapped circuit is the cleartext interpreter
 * fbs_mapper/fbs_exec_env.py:208-229 (restated in oracle/cleartext.py).  The TFHE arithmetic the
 * reference *assumes* lives in zama-ai/concrete @ nightly-2024.04.17 (README.md:17-18), which is
 * neither vendored nor called by the reference for execution.  PARITY UNPINNED at ciphertext level:
 * no reference test or golden vector pins any ciphertext.  What IS pinned is the decrypted result,
 * which must equal LutExecEnv.eval (fbs_exec_env.py:208-229) on the same inputs.
 *
 * This file therefore restates the published CGGI "key-switch -> programmable bootstrap" atomic
 * pattern (the pattern experiments/concrete.patch:62-74 edits) with the message encoding of
 * experiments/concrete.patch:21-27 (absolute number of message values p, one negacyclic padding
 * "bit": Delta = q/(2p), decision half-interval q/(4p)) and the table modes of
 * fbs_mapper/map_to_fbs.py:81-98, over q = p1*p2 (two 30-bit NTT primes, 60 bits) as ciphertext modulus.  It is deliberately written with plain loops and unsigned __int128 so that it shares no
 * code (and no bugs) with the CUDA product; both follow the written spec in DESIGN.md section 3, so
 * with identical seeds they must agree BIT FOR BIT at every ciphertext tap.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef uint8_t u8;
typedef unsigned __int128 u128;
typedef __int128 i128;

#define RP1 1073643521ULL            /* p1 = 0x3FFE8001 */
#define RP2 1073692673ULL            /* p2 = 0x3FFF4001 */
#define GLP (RP1 * RP2)              /* ciphertext modulus q = p1*p2 = 0x0FFF70019FFDC001 */

/* ---------------- field arithmetic (slow-and-obvious on purpose) ---------------- */
static inline u64 f_add(u64 a, u64 b) { u128 s = (u128)a + b; if (s >= GLP) s -= GLP; return (u64)s; }
static inline u64 f_sub(u64 a, u64 b) { return a >= b ? a - b : (u64)((u128)a + GLP - b); }
static inline u64 f_neg(u64 a) { return a ? GLP - a : 0; }
static inline u64 f_mul(u64 a, u64 b) { return (u64)(((u128)a * b) % GLP); }   /* slow and obvious on purpose */
static u64 pmod(u64 b, u64 e, u64 m) { u128 r = 1, x = b % m; while (e) { if (e & 1) r = r * x % m; x = x * x % m; e >>= 1; } return (u64)r; }
static u64 crt(u64 a, u64 b)       /* the x in [0,q) with x = a mod p1, x = b mod p2 */
{
    u64 p1inv = pmod(RP1, RP2 - 2, RP2);
    u64 t = (u64)((u128)((b + RP2 - a % RP2) % RP2) * p1inv % RP2);
    return a + RP1 * t;
}
static u64 f_pow(u64 b, u64 e) { u64 r = 1; while (e) { if (e & 1) r = f_mul(r, b); b = f_mul(b, b); e >>= 1; } return r; }
static inline u64 f_from_i64(i64 v) { return v >= 0 ? (u64)v % GLP : GLP - ((u64)(-v) % GLP); }

/* ---------------- deterministic counter PRNG ---------------- */
static inline u64 mix64(u64 z)
{
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31; return z;
}
static inline u64 rnd64(u64 seed, u64 dom, u64 idx)
{
    u64 h = mix64(seed ^ (dom * 0xD1B54A32D192ED03ULL));
    return mix64(h + (idx + 1) * 0x9E3779B97F4A7C15ULL);
}
static inline u64 rnd_uniform(u64 seed, u64 dom, u64 idx)
{
    u64 u = rnd64(seed, dom, idx) >> 4;                       /* 60 bits */
    if (u >= GLP) u = (u64)(((u128)rnd64(seed, dom + 64, idx) * GLP) >> 64);   /* rare second draw, scaled */
    return u;
}
/* Irwin-Hall(12) over 32-bit uniforms, std = scale (in units of 1/Q of the torus) */
static inline u64 rnd_noise(u64 seed, u64 dom, u64 idx, u64 scale)
{
    u64 S = 0;
    for (int t = 0; t < 6; t++) { u64 r = rnd64(seed, dom, idx * 6 + t); S += (r & 0xFFFFFFFFULL) + (r >> 32); }
    i64 c = (i64)S - (i64)(6ULL * 0xFFFFFFFFULL);
    i128 prod = (i128)c * (i128)scale;
    i128 e = (prod + ((i128)1 << 31)) >> 32;   /* arithmetic shift = floor */
    return f_from_i64((i64)e);
}
enum { DOM_SLWE = 1, DOM_SGLWE = 2, DOM_BSK_MASK = 3, DOM_BSK_NOISE = 4, DOM_KSK_MASK = 5, DOM_KSK_NOISE = 6,
       DOM_ENC_MASK = 7, DOM_ENC_NOISE = 8 };

/* ---------------- parameters / context ---------------- */
typedef struct {
    int32_t n, k, N, bsk_l, bsk_beta, ks_l, ks_beta, bsk_unroll;   /* bsk_unroll: 0/1 classic, 2 / 3 = that many key bits per step */
    u64 lwe_noise, glwe_noise;   /* round(sigma * P) */
} ref_params;

typedef struct {
    ref_params P; u64 seed; int logN;
    u8 *s_lwe;      /* [n] */
    u8 *s_big;      /* [k*N]  (s_big[u*N+j] = S_u[j]) */
    u64 *ksk;       /* [k*N*ks_l][n+1] */
    u64 *bsk_coef;  /* [n_ggsw][(k+1)*l][k+1][N] coefficient domain (n_ggsw = n, or 3n/2 when unrolled) */
    u64 *bsk_ntt;   /* same, oracle-order NTT domain */
    u64 *psi_rev, *psi_inv_rev; u64 ninv;
} ref_ctx;

static int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }
static u32 bitrev(u32 x, int bits) { u32 r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

/* gadget element g_j = round(Q / B^(j+1)), j = 0..l-1 */
static u64 gadget(int beta, int j) { u128 B = (u128)1 << (beta * (j + 1)); return (u64)(((u128)GLP + B / 2) / B); }

/* y = round(x * 2^bits / q) mod 2^bits.  With x = r1 + p1*t: up to 24 bits
 * y = (t*K63 + 8*r1 + 2^(s-1)) >> s, s = 63 - bits, K63 = floor(2^63 / p2); beyond, through floor(2^123 / q). */
static u64 round_top(u64 x, int bits)
{
    u64 mask = (1ULL << bits) - 1;
    if (bits <= 24) {
        u64 t = x / RP1, r1 = x % RP1, K63 = (1ULL << 63) / RP2;
        int s = 63 - bits;
        return ((t * K63 + (r1 << 3) + (1ULL << (s - 1))) >> s) & mask;
    }
    u64 RQ = (u64)((((u128)1) << 123) / GLP);
    u64 sc = (u64)(((u128)x * RQ) >> 64);
    return ((sc + (1ULL << (58 - bits))) >> (59 - bits)) & mask;
}
/* decomposition: closest multiple of q/B^l, balanced digits in [-B/2, B/2), d[0] = most significant level */
static void decompose(u64 x, int beta, int l, int32_t *d)
{
    u64 y = round_top(x, beta * l);
    u64 Bm = (1ULL << beta) - 1, half = 1ULL << (beta - 1);
    for (int j = l - 1; j >= 0; j--) {
        u64 dig = y & Bm; y >>= beta;
        if (dig >= half) { d[j] = (int32_t)((i64)dig - (i64)(1LL << beta)); y += 1; } else d[j] = (int32_t)dig;
    }
}
static inline u32 modswitch(u64 x, int log2N /* log2(2N) */) { return (u32)round_top(x, log2N); }
static inline u64 delta_of(int p) { return (GLP + (u64)p) / (2ULL * (u64)p); }

/* ---------------- negacyclic NTT (Longa-Naehrig layout, natural in -> bit-reversed out) ---------------- */
static void ntt_fwd(const ref_ctx *c, u64 *a)
{
    int N = c->P.N; int t = N;
    for (int m = 1; m < N; m <<= 1) {
        t >>= 1;
        for (int i = 0; i < m; i++) {
            u64 S = c->psi_rev[m + i];
            for (int j = 2 * i * t; j < 2 * i * t + t; j++) {
                u64 U = a[j], V = f_mul(a[j + t], S);
                a[j] = f_add(U, V); a[j + t] = f_sub(U, V);
            }
        }
    }
}
static void ntt_inv(const ref_ctx *c, u64 *a)
{
    int N = c->P.N; int t = 1;
    for (int m = N >> 1; m >= 1; m >>= 1) {
        for (int i = 0; i < m; i++) {
            u64 S = c->psi_inv_rev[m + i];
            for (int j = 2 * i * t; j < 2 * i * t + t; j++) {
                u64 U = a[j], V = a[j + t];
                a[j] = f_add(U, V); a[j + t] = f_mul(f_sub(U, V), S);
            }
        }
        t <<= 1;
    }
    for (int j = 0; j < N; j++) a[j] = f_mul(a[j], c->ninv);
}

void ref_polymul_schoolbook(int N, const u64 *a, const u64 *b, u64 *out)
{
    for (int i = 0; i < N; i++) out[i] = 0;
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) {
        u64 pr = f_mul(a[i], b[j]); int kx = i + j;
        if (kx < N) out[kx] = f_add(out[kx], pr); else out[kx - N] = f_sub(out[kx - N], pr);
    }
}
void ref_polymul_ntt(const ref_ctx *c, const u64 *a, const u64 *b, u64 *out)
{
    int N = c->P.N; u64 *x = malloc(8 * N), *y = malloc(8 * N);
    memcpy(x, a, 8 * N); memcpy(y, b, 8 * N); ntt_fwd(c, x); ntt_fwd(c, y);
    for (int i = 0; i < N; i++) out[i] = f_mul(x[i], y[i]);
    ntt_inv(c, out); free(x); free(y);
}
void ref_ntt_fwd(const ref_ctx *c, u64 *a) { ntt_fwd(c, a); }
void ref_ntt_inv(const ref_ctx *c, u64 *a) { ntt_inv(c, a); }
u64 ref_mulmod(u64 a, u64 b) { return f_mul(a, b); }
u64 ref_rnd64(u64 s, u64 d, u64 i) { return rnd64(s, d, i); }
u64 ref_noise(u64 s, u64 d, u64 i, u64 scale) { return rnd_noise(s, d, i, scale); }
void ref_decompose(u64 x, int beta, int l, int32_t *d) { decompose(x, beta, l, d); }
u32 ref_modswitch_word(u64 x, int log2N) { return modswitch(x, log2N); }
u64 ref_delta(int p) { return delta_of(p); }
u64 ref_gadget(int beta, int j) { return gadget(beta, j); }

/* ---------------- context ---------------- */
ref_ctx *ref_ctx_create(const ref_params *P, u64 seed)
{
    ref_ctx *c = calloc(1, sizeof(ref_ctx));
    c->P = *P; c->seed = seed; c->logN = ilog2(P->N);
    int N = P->N;
    /* primitive 2N-th root of unity mod q = CRT of roots mod each prime; 3 is a quadratic non-residue mod p1 and p2,
     * so 3^((p-1)/2N) has order exactly 2N.  q is composite: inverses come from CRT, not from Fermat. */
    u64 psi1 = pmod(3, (RP1 - 1) / (2ULL * N), RP1), psi2 = pmod(3, (RP2 - 1) / (2ULL * N), RP2);
    u64 psi = crt(psi1, psi2), psi_inv = crt(pmod(psi1, RP1 - 2, RP1), pmod(psi2, RP2 - 2, RP2));
    c->psi_rev = malloc(8 * N); c->psi_inv_rev = malloc(8 * N);
    for (int i = 0; i < N; i++) {
        u32 r = bitrev((u32)i, c->logN);
        c->psi_rev[i] = f_pow(psi, r); c->psi_inv_rev[i] = f_pow(psi_inv, r);
    }
    c->ninv = crt(pmod((u64)N, RP1 - 2, RP1), pmod((u64)N, RP2 - 2, RP2));
    return c;
}
void ref_ctx_destroy(ref_ctx *c)
{
    if (!c) return;
    free(c->s_lwe); free(c->s_big); free(c->ksk); free(c->bsk_coef); free(c->bsk_ntt); free(c->psi_rev); free(c->psi_inv_rev); free(c);
}

/* Key unrolling (m = bsk_unroll key bits per blind-rotation step, m = 2 or 3):
 *   X^(sum_i a_i s_i) - 1 = sum over the non-empty subsets T of the m key bits of (X^(sum_{i in T} a_i) - 1) * [key bits == T],
 * so group t of the key gets 2^m - 1 GGSW ciphertexts, of the indicator bits prod_{i in T} s_i * prod_{i not in T} (1 - s_i)
 * (GGSW index (2^m - 1) t + c, subset mask_of(m, c); bit i of the mask stands for key bit m t + i).  For m = 2:
 *   X^(a1 s1 + a2 s2) - 1 = (X^(a1+a2) - 1) s1 s2 + (X^a1 - 1) s1 (1 - s2) + (X^a2 - 1) (1 - s1) s2.
 * n is padded with zero key bits (and a_i = 0) to a multiple of m.  Classic (m <= 1): GGSW i encrypts s_i. */
static int unroll_m(const ref_params *P) { return P->bsk_unroll == 2 || P->bsk_unroll == 3 ? P->bsk_unroll : 1; }
static int n_groups(const ref_params *P) { int m = unroll_m(P); return (P->n + m - 1) / m; }
static int n_sub(const ref_params *P) { return (1 << unroll_m(P)) - 1; }
static int mask_of(int m, int c) { static const int m2[3] = {3, 1, 2}; return m == 2 ? m2[c] : c + 1; }
static int n_ggsw(const ref_params *P) { return unroll_m(P) == 1 ? P->n : n_sub(P) * n_groups(P); }
static int key_bit(const ref_ctx *c, int i) { return i < c->P.n ? c->s_lwe[i] : 0; }
static int ggsw_bit(const ref_ctx *c, int g)
{
    const int m = unroll_m(&c->P);
    if (m == 1) return c->s_lwe[g];
    const int ns = n_sub(&c->P), t = g / ns, mask = mask_of(m, g % ns);
    int bit = 1;
    for (int i = 0; i < m; i++) bit &= ((mask >> i) & 1) ? key_bit(c, m * t + i) : !key_bit(c, m * t + i);
    return bit;
}
void ref_keygen(ref_ctx *c)
{
    const ref_params *P = &c->P; int n = P->n, k = P->k, N = P->N, l = P->bsk_l, lk = P->ks_l;
    c->s_lwe = malloc(n); c->s_big = malloc((size_t)k * N);
    for (int i = 0; i < n; i++) c->s_lwe[i] = (u8)(rnd64(c->seed, DOM_SLWE, i) & 1);
    for (int i = 0; i < k * N; i++) c->s_big[i] = (u8)(rnd64(c->seed, DOM_SGLWE, i) & 1);
    /* KSK: row r = i*lk + j encrypts s_big[i] * g_j under s_lwe */
    size_t R = (size_t)k * N * lk;
    c->ksk = malloc(R * (n + 1) * 8);
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < R; r++) {
        size_t i = r / lk; int j = (int)(r % lk);
        u64 *row = c->ksk + r * (n + 1); u64 body = 0;
        for (int q = 0; q < n; q++) {
            u64 a = rnd_uniform(c->seed, DOM_KSK_MASK, r * n + q); row[q] = a;
            if (c->s_lwe[q]) body = f_add(body, a);
        }
        body = f_add(body, rnd_noise(c->seed, DOM_KSK_NOISE, r, P->lwe_noise));
        if (c->s_big[i]) body = f_add(body, gadget(P->ks_beta, j));
        row[n] = body;
    }
    /* BSK: GGSW(s_lwe[i]); row r = u*l + j (u <= k input poly, j level); polys v = 0..k (v = k is the body) */
    int rows = (k + 1) * l; const int ng = n_ggsw(P); size_t polys = (size_t)ng * rows * (k + 1);
    c->bsk_coef = malloc(polys * N * 8); c->bsk_ntt = malloc(polys * N * 8);
#pragma omp parallel for schedule(dynamic, 4)
    for (int ir = 0; ir < ng * rows; ir++) {
        int i = ir / rows, r = ir % rows, u = r / l, j = r % l;
        u64 *row = c->bsk_coef + (size_t)ir * (k + 1) * N;
        u64 *body = row + (size_t)k * N; u64 *sp = malloc(8 * N), *tmp = malloc(8 * N);
        for (int q = 0; q < N; q++) body[q] = rnd_noise(c->seed, DOM_BSK_NOISE, (u64)ir * N + q, P->glwe_noise);
        for (int v = 0; v < k; v++) {
            u64 *A = row + (size_t)v * N;
            for (int q = 0; q < N; q++) { A[q] = rnd_uniform(c->seed, DOM_BSK_MASK, ((u64)ir * k + v) * N + q); sp[q] = c->s_big[v * N + q]; }
            ref_polymul_ntt(c, A, sp, tmp);
            for (int q = 0; q < N; q++) body[q] = f_add(body[q], tmp[q]);
        }
        if (ggsw_bit(c, i)) { u64 g = gadget(P->bsk_beta, j); u64 *tgt = row + (size_t)u * N; tgt[0] = f_add(tgt[0], g); }
        free(sp); free(tmp);
        u64 *rown = c->bsk_ntt + (size_t)ir * (k + 1) * N;
        memcpy(rown, row, (size_t)(k + 1) * N * 8);
        for (int v = 0; v <= k; v++) ntt_fwd(c, rown + (size_t)v * N);
    }
}
void ref_get_keys(const ref_ctx *c, u8 *s_lwe, u8 *s_big, u64 *ksk, u64 *bsk_coef)
{
    const ref_params *P = &c->P;
    if (s_lwe) memcpy(s_lwe, c->s_lwe, P->n);
    if (s_big) memcpy(s_big, c->s_big, (size_t)P->k * P->N);
    if (ksk) memcpy(ksk, c->ksk, (size_t)P->k * P->N * P->ks_l * (P->n + 1) * 8);
    if (bsk_coef) memcpy(bsk_coef, c->bsk_coef, (size_t)n_ggsw(P) * (P->k + 1) * P->bsk_l * (P->k + 1) * P->N * 8);
}

/* ---------------- LWE ops on "big" ciphertexts (dimension kN, body last) ---------------- */
void ref_encrypt(const ref_ctx *c, int p, const int32_t *msg, const u64 *ct_id, int64_t count, u64 enc_seed, u64 *out)
{
    int D = c->P.k * c->P.N; u64 delta = delta_of(p);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < count; e++) {
        u64 *ct = out + (size_t)e * (D + 1); u64 body = 0;
        for (int q = 0; q < D; q++) {
            u64 a = rnd_uniform(enc_seed, DOM_ENC_MASK, ct_id[e] * (u64)D + q); ct[q] = a;
            if (c->s_big[q]) body = f_add(body, a);
        }
        body = f_add(body, rnd_noise(enc_seed, DOM_ENC_NOISE, ct_id[e], c->P.glwe_noise));
        int m = ((msg[e] % (2 * p)) + 2 * p) % (2 * p);
        ct[D] = f_add(body, f_mul((u64)m, delta));
    }
}
u64 ref_phase(const ref_ctx *c, const u64 *ct)
{
    int D = c->P.k * c->P.N; u64 acc = 0;
    for (int q = 0; q < D; q++) if (c->s_big[q]) acc = f_add(acc, ct[q]);
    return f_sub(ct[D], acc);
}
static int32_t decode(u64 phase, int p)
{
    u64 delta = delta_of(p), half = delta >> 1;
    if (phase >= GLP - half) return 0;
    return (int32_t)(((phase + half) / delta) % (2ULL * p));
}
void ref_decrypt(const ref_ctx *c, int p, const u64 *cts, int64_t count, int32_t *out)
{
    int D = c->P.k * c->P.N;
    for (int64_t e = 0; e < count; e++) out[e] = decode(ref_phase(c, cts + (size_t)e * (D + 1)), p);
}
/* out = sum coef_i * op_i + cst*Delta (body only).  SURVEY Appendix A.5 / fbs_exec_env.py:215-217 */
void ref_lincomb(const ref_ctx *c, int p, int nops, const u64 *const *ops, const int32_t *coefs, int32_t cst, u64 *out)
{
    int D = c->P.k * c->P.N;
    for (int q = 0; q <= D; q++) {
        u64 acc = 0;
        for (int o = 0; o < nops; o++) acc = f_add(acc, f_mul(ops[o][q], f_from_i64(coefs[o])));
        out[q] = acc;
    }
    out[D] = f_add(out[D], f_mul(f_from_i64(cst), delta_of(p)));
}
/* key switch kN -> n: out = (0,..,0,b) - sum_{i,j} d_ij * KSK[i*lk+j] */
void ref_keyswitch(const ref_ctx *c, const u64 *in, u64 *out)
{
    const ref_params *P = &c->P; int D = P->k * P->N, n = P->n, lk = P->ks_l;
    for (int q = 0; q < n; q++) out[q] = 0;
    out[n] = in[D];
    int32_t d[64];
    for (int i = 0; i < D; i++) {
        decompose(in[i], P->ks_beta, lk, d);
        for (int j = 0; j < lk; j++) {
            if (!d[j]) continue;
            const u64 *row = c->ksk + ((size_t)i * lk + j) * (n + 1); u64 dj = f_from_i64(d[j]);
            for (int q = 0; q <= n; q++) out[q] = f_sub(out[q], f_mul(dj, row[q]));
        }
    }
}
void ref_modswitch(const ref_ctx *c, const u64 *in, uint16_t *out)
{
    int lg = c->logN + 1;
    for (int q = 0; q <= c->P.n; q++) out[q] = (uint16_t)modswitch(in[q], lg);
}
/* test polynomial (table modes of map_to_fbs.py:81-98) */
void ref_test_poly(const ref_ctx *c, int p, const u8 *table, int L, int s, u64 *tv)
{
    int N = c->P.N; u64 delta = delta_of(p); u64 off = f_mul((u64)s, delta >> 1);
    for (int j = 0; j < N; j++) {
        int x = (int)((2LL * j * p + N) / (2LL * N)); int neg = 0;
        if (x >= p) { x -= p; neg = 1; }
        u64 tvx = (x < L) ? (u64)table[x] : 0;
        u64 F = f_sub(f_mul(tvx, delta), off);
        tv[j] = neg ? f_neg(F) : F;
    }
}
/* (X^a * poly)[j], a in [0, 2N) */
static inline u64 rot_coef(const u64 *poly, int N, int j, int a)
{
    int idx = j - a; while (idx < 0) idx += 2 * N;
    return idx < N ? poly[idx] : f_neg(poly[idx - N]);
}
/* dig[r] = NTT of digit polynomial r of the gadget decomposition of `src` (k+1 polynomials) */
static void decompose_ntt(const ref_ctx *c, const u64 *src, int rotate_by /* < 0: decompose src itself; else X^a src - src */, u64 *dig)
{
    const ref_params *P = &c->P; int k = P->k, N = P->N, l = P->bsk_l;
    int32_t d[16];
    for (int u = 0; u <= k; u++) {
        const u64 *pu = src + (size_t)u * N;
        for (int j = 0; j < N; j++) {
            u64 x = rotate_by < 0 ? pu[j] : f_sub(rot_coef(pu, N, j, rotate_by), pu[j]);
            decompose(x, P->bsk_beta, l, d);
            for (int jj = 0; jj < l; jj++) dig[((size_t)(u * l + jj)) * N + j] = f_from_i64(d[jj]);
        }
    }
    for (int r = 0; r < (k + 1) * l; r++) ntt_fwd(c, dig + (size_t)r * N);
}
/* outp[v] = sum_r dig[r] * GGSW_g[r][v]  (external product with the already decomposed input), coefficient domain */
static void ext_product(const ref_ctx *c, const u64 *dig, int g, u64 *outp)
{
    const ref_params *P = &c->P; int k = P->k, N = P->N, rows = (k + 1) * P->bsk_l;
    for (int v = 0; v <= k; v++) {
        u64 *o = outp + (size_t)v * N;
        for (int j = 0; j < N; j++) o[j] = 0;
        for (int r = 0; r < rows; r++) {
            const u64 *b = c->bsk_ntt + (((size_t)g * rows + r) * (k + 1) + v) * N; const u64 *dd = dig + (size_t)r * N;
            for (int j = 0; j < N; j++) o[j] = f_add(o[j], f_mul(dd[j], b[j]));
        }
        ntt_inv(c, o);
    }
}
void ref_blind_rotate(const ref_ctx *c, const uint16_t *ms, const u64 *tv, u64 *acc /* [(k+1)][N] */)
{
    const ref_params *P = &c->P; int n = P->n, k = P->k, N = P->N, l = P->bsk_l, rows = (k + 1) * l;
    memset(acc, 0, (size_t)(k + 1) * N * 8);
    int bt = ms[n];
    for (int j = 0; j < N; j++) acc[(size_t)k * N + j] = rot_coef(tv, N, j, (2 * N - bt) % (2 * N));
    u64 *dig = malloc((size_t)rows * N * 8), *outp = malloc((size_t)(k + 1) * N * 8);
    if (unroll_m(P) > 1) {
        /* ACC <- ACC + sum_c (X^{e_c} - 1) * (Dec(ACC) [x] GGSW_{ns t + c}),  e_c = sum of the a_i in subset c: one decomposition per key group */
        const int m = unroll_m(P), ns = n_sub(P);
        u64 *delta = malloc((size_t)(k + 1) * N * 8);
        for (int t = 0; t < n_groups(P); t++) {
            decompose_ntt(c, acc, -1, dig);
            memset(delta, 0, (size_t)(k + 1) * N * 8);
            for (int cc = 0; cc < ns; cc++) {
                int e = 0;
                for (int i = 0; i < m; i++) if (((mask_of(m, cc) >> i) & 1) && m * t + i < n) e += ms[m * t + i];
                e %= 2 * N;
                ext_product(c, dig, ns * t + cc, outp);
                for (int v = 0; v <= k; v++) for (int j = 0; j < N; j++) {
                    const u64 *o = outp + (size_t)v * N;
                    delta[(size_t)v * N + j] = f_add(delta[(size_t)v * N + j], f_sub(rot_coef(o, N, j, e), o[j]));
                }
            }
            for (size_t w = 0; w < (size_t)(k + 1) * N; w++) acc[w] = f_add(acc[w], delta[w]);
        }
        free(delta);
    } else {
        for (int i = 0; i < n; i++) {
            decompose_ntt(c, acc, ms[i], dig);
            ext_product(c, dig, i, outp);
            for (size_t w = 0; w < (size_t)(k + 1) * N; w++) acc[w] = f_add(acc[w], outp[w]);
        }
    }
    free(dig); free(outp);
}
void ref_sample_extract(const ref_ctx *c, const u64 *acc, u64 body_offset, u64 *out)
{
    int k = c->P.k, N = c->P.N;
    for (int u = 0; u < k; u++) for (int j = 0; j < N; j++)
        out[(size_t)u * N + j] = (j == 0) ? acc[(size_t)u * N] : f_neg(acc[(size_t)u * N + N - j]);
    out[(size_t)k * N] = f_add(acc[(size_t)k * N], body_offset);
}
/* full bootstrap of one big-LWE ciphertext; optional taps for stage-wise parity */
void ref_pbs(const ref_ctx *c, int p, const u64 *in, const u8 *table, int L, int s, u64 *out,
             u64 *tap_ks /* [n+1] or NULL */, uint16_t *tap_ms /* [n+1] or NULL */, u64 *tap_acc /* [(k+1)N] or NULL */)
{
    const ref_params *P = &c->P; int n = P->n, k = P->k, N = P->N;
    u64 *ks = malloc((size_t)(n + 1) * 8); uint16_t *ms = malloc((size_t)(n + 1) * 2);
    u64 *tv = malloc((size_t)N * 8), *acc = malloc((size_t)(k + 1) * N * 8);
    ref_keyswitch(c, in, ks); ref_modswitch(c, ks, ms);
    ref_test_poly(c, p, table, L, s, tv);
    ref_blind_rotate(c, ms, tv, acc);
    ref_sample_extract(c, acc, f_mul((u64)s, delta_of(p) >> 1), out);
    if (tap_ks) memcpy(tap_ks, ks, (size_t)(n + 1) * 8);
    if (tap_ms) memcpy(tap_ms, ms, (size_t)(n + 1) * 2);
    if (tap_acc) memcpy(tap_acc, acc, (size_t)(k + 1) * N * 8);
    free(ks); free(ms); free(tv); free(acc);
}

/* ---------------- multi-value bootstrap (several tables on ONE lincomb, fbs_exec_env.py:93-100) ---------
 * One blind rotation serves every table f applied to the same input: rotate the table-independent base polynomial
 *     TV0 = H * (1 + X + .. + X^(N-1)),   H = Delta * 2^-1 mod q      (so that (1 - X) * TV0 = 2H = Delta),
 * then multiply the accumulator by the sparse small polynomial e_f with  Delta * e_f = (1 - X) * TV_f,  where
 * TV_f[j] = F'(slot(j)), F'(x) = (2 tv[x] - s) * H, is the (negacyclic) step function of the table: e_f is non-zero only at
 * the p slot boundaries j_x = ceil(N (2x - 1) / 2p), x = 1..p, with e_f[j_x] = tv[x] - tv[x-1] (x < p) and
 * e_f[j_p] = -(tv[0] + tv[p-1] - s) (the wrap into the negated half).  ACC * e_f = GLWE(X^-mu * TV_f); sample-extract
 * coefficient 0 and add s * H: the result encrypts tv[mu] * Delta exactly as the one-table bootstrap does (2 H = Delta mod q), with
 * the blind-rotation noise multiplied by |e_f|^2 <= p + 3 (Carpov, Izabachene, Mollimard, "New techniques for multi-value
 * input homomorphic evaluation and applications", CT-RSA 2019). */
static u64 half_delta(int p) { return f_mul(delta_of(p), (GLP + 1) / 2); }          /* H = Delta / 2 mod q (q is odd) */
void ref_base_test_poly(const ref_ctx *c, int p, u64 *tv) { for (int j = 0; j < c->P.N; j++) tv[j] = half_delta(p); }
/* e_f as (position, coefficient) pairs; returns their number (<= p) */
static int table_steps(int N, int p, const u8 *table, int L, int s, int *pos, int *coef)
{
    int cnt = 0;
    for (int x = 1; x <= p; x++) {
        int tx = (x < p) ? (x < L ? table[x] : 0) : 0, tp = (x - 1 < L) ? table[x - 1] : 0;
        int e = (x < p) ? tx - tp : -((L > 0 ? table[0] : 0) + tp - s);
        if (!e) continue;
        pos[cnt] = (int)(((long long)N * (2 * x - 1) + 2 * p - 1) / (2 * p)); coef[cnt] = e; cnt++;
    }
    return cnt;
}
/* out = sample_extract( acc * e_f ) + s*H on the body */
void ref_multi_extract(const ref_ctx *c, int p, const u64 *acc, const u8 *table, int L, int s, u64 *out)
{
    int k = c->P.k, N = c->P.N; int pos[128], coef[128];
    int cnt = table_steps(N, p, table, L, s, pos, coef);
    u64 *prod = calloc((size_t)(k + 1) * N, 8);
    for (int v = 0; v <= k; v++) for (int j = 0; j < N; j++) {
        u64 a = 0;
        for (int t = 0; t < cnt; t++) {                                   /* (acc * X^pos)[j] = +-acc[(j - pos) mod N] */
            int idx = j - pos[t]; u64 x = idx >= 0 ? acc[(size_t)v * N + idx] : f_neg(acc[(size_t)v * N + idx + N]);
            u64 cf = f_from_i64(coef[t]);
            a = f_add(a, f_mul(x, cf));
        }
        prod[(size_t)v * N + j] = a;
    }
    ref_sample_extract(c, prod, f_mul((u64)s, half_delta(p)), out);
    free(prod);
}
/* one rotation, T tables: tables [T][tab_stride], outs [T][kN+1]; taps as ref_pbs (tap_acc = accumulator BEFORE the e_f products) */
void ref_pbs_multi(const ref_ctx *c, int p, const u64 *in, const u8 *tables, int tab_stride, const u8 *tlen, const int32_t *modes, int T,
                   u64 *outs, u64 *tap_acc)
{
    const ref_params *P = &c->P; int n = P->n, k = P->k, N = P->N; size_t CT = (size_t)k * N + 1;
    u64 *ks = malloc((size_t)(n + 1) * 8); uint16_t *ms = malloc((size_t)(n + 1) * 2);
    u64 *tv = malloc((size_t)N * 8), *acc = malloc((size_t)(k + 1) * N * 8);
    ref_keyswitch(c, in, ks); ref_modswitch(c, ks, ms);
    ref_base_test_poly(c, p, tv);
    ref_blind_rotate(c, ms, tv, acc);
    for (int t = 0; t < T; t++) ref_multi_extract(c, p, acc, tables + (size_t)t * tab_stride, tlen[t], modes ? modes[t] : 1, outs + (size_t)t * CT);
    if (tap_acc) memcpy(tap_acc, acc, (size_t)(k + 1) * N * 8);
    free(ks); free(ms); free(tv); free(acc);
}

/* `count` independent bootstraps (one ref_pbs each, OpenMP over the batch): the checker of the full-size batched parity
 * tests.  tables [count][tab_stride], out [count][kN+1], tap_acc [count][(k+1)N] or NULL. */
void ref_pbs_batch(const ref_ctx *c, int p, const u64 *in, const u8 *tables, int tab_stride, const u8 *tlen, const int32_t *modes,
                   int64_t count, u64 *out, u64 *tap_acc, int threads)
{
    size_t CT = (size_t)c->P.k * c->P.N + 1, AW = (size_t)(c->P.k + 1) * c->P.N;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < count; i++)
        ref_pbs(c, p, in + (size_t)i * CT, tables + (size_t)i * tab_stride, tlen[i], modes ? modes[i] : 1, out + (size_t)i * CT,
                NULL, NULL, tap_acc ? tap_acc + (size_t)i * AW : NULL);
}

/* ---------------- levelised program (same flat descriptor as include/fbs_b200.h) ---------------- */
typedef struct {
    int32_t p, n_inputs, n_lincombs, n_boots, n_levels, n_slots, n_outputs, reserved;   /* reserved: contiguous_levels of the product's descriptor (unused here) */
    const int32_t *lc_level_ptr, *bs_level_ptr;
    const int32_t *lc_ptr, *lc_slot, *lc_coef, *lc_const;
    const int32_t *bs_lc, *bs_slot, *bs_tab_ptr; const u8 *bs_tab; const int32_t *bs_mode;
    const int32_t *in_slot;
    const int32_t *out_ptr, *out_slot, *out_coef, *out_const;
} ref_prog_desc;

/* Encrypted evaluation of B instances; in [n_inputs][B] bits, out [n_outputs][B] (values mod 2p).
 * Follows the cleartext interpreter's order of evaluation, fbs_exec_env.py:208-229, one level at a time. */
int ref_eval_prog_mv(const ref_ctx *c, const ref_prog_desc *g, const u8 *in, int64_t B, int64_t inst_offset, int64_t B_total,
                     u64 enc_seed, u8 *out, int threads, int multi_value);
int ref_eval_prog(const ref_ctx *c, const ref_prog_desc *g, const u8 *in, int64_t B, int64_t inst_offset, int64_t B_total,
                  u64 enc_seed, u8 *out, int threads)
{
    return ref_eval_prog_mv(c, g, in, B, inst_offset, B_total, enc_seed, out, threads, 0);
}
int ref_eval_prog_mv(const ref_ctx *c, const ref_prog_desc *g, const u8 *in, int64_t B, int64_t inst_offset, int64_t B_total,
                     u64 enc_seed, u8 *out, int threads, int multi_value)
{
    int D = c->P.k * c->P.N; size_t CT = (size_t)D + 1; int p = g->p;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; b++) {
        u64 *w = malloc((size_t)g->n_slots * CT * 8);
        u64 *lc = malloc((size_t)(g->n_lincombs ? g->n_lincombs : 1) * CT * 8);
        for (int i = 0; i < g->n_inputs; i++) {
            int32_t m = in[(size_t)i * B + b]; u64 id = (u64)i * (u64)B_total + (u64)(inst_offset + b);
            ref_encrypt(c, p, &m, &id, 1, enc_seed, w + (size_t)g->in_slot[i] * CT);
        }
        for (int lv = 0; lv < g->n_levels; lv++) {
            for (int q = g->lc_level_ptr[lv]; q < g->lc_level_ptr[lv + 1]; q++) {
                int nops = g->lc_ptr[q + 1] - g->lc_ptr[q]; const u64 *ops[64];
                for (int o = 0; o < nops; o++) ops[o] = w + (size_t)g->lc_slot[g->lc_ptr[q] + o] * CT;
                ref_lincomb(c, p, nops, ops, g->lc_coef + g->lc_ptr[q], g->lc_const[q], lc + (size_t)q * CT);
            }
            for (int q = g->bs_level_ptr[lv]; q < g->bs_level_ptr[lv + 1]; ) {
                int L = g->bs_tab_ptr[q + 1] - g->bs_tab_ptr[q];
                if (!multi_value) {
                    ref_pbs(c, p, lc + (size_t)g->bs_lc[q] * CT, g->bs_tab + g->bs_tab_ptr[q], L, g->bs_mode[q],
                            w + (size_t)g->bs_slot[q] * CT, NULL, NULL, NULL);
                    q++;
                    continue;
                }
                /* multi-value: the maximal run of bootstraps on the same lincomb (they are sorted by lincomb) shares one rotation */
                int q1 = q; while (q1 < g->bs_level_ptr[lv + 1] && g->bs_lc[q1] == g->bs_lc[q] && q1 - q < 64) q1++;   /* at most 64 tables per call (buffers below); a longer run continues with another rotation of the same input: same accumulator, same results */
                int T = q1 - q; u8 tabs[64 * 64]; u8 lens[64]; int32_t md[64]; u64 *outs = malloc((size_t)T * CT * 8);
                for (int t = 0; t < T; t++) {
                    lens[t] = (u8)(g->bs_tab_ptr[q + t + 1] - g->bs_tab_ptr[q + t]); md[t] = g->bs_mode[q + t];
                    memcpy(tabs + 64 * t, g->bs_tab + g->bs_tab_ptr[q + t], lens[t]);
                }
                ref_pbs_multi(c, p, lc + (size_t)g->bs_lc[q] * CT, tabs, 64, lens, md, T, outs, NULL);
                for (int t = 0; t < T; t++) memcpy(w + (size_t)g->bs_slot[q + t] * CT, outs + (size_t)t * CT, CT * 8);
                free(outs);
                q = q1;
            }
        }
        u64 *o = malloc(CT * 8);
        for (int q = 0; q < g->n_outputs; q++) {
            int nops = g->out_ptr[q + 1] - g->out_ptr[q]; const u64 *ops[64];
            for (int t = 0; t < nops; t++) ops[t] = w + (size_t)g->out_slot[g->out_ptr[q] + t] * CT;
            ref_lincomb(c, p, nops, ops, g->out_coef + g->out_ptr[q], g->out_const[q], o);
            int32_t m; ref_decrypt(c, p, o, 1, &m);
            out[(size_t)q * B + b] = (u8)m;
        }
        free(o); free(w); free(lc);
    }
    return bad;
}
int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
