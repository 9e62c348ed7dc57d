// Host emulation of the device NTT pass/layout/twiddle logic (tfhe_fbs_map_b200/csrc/ntt.cuh).
// Loops over tau play the threads, array copies play the shared-memory transposes.  Compared against the
// textbook in-place negacyclic NTT loops and against a schoolbook product.  Exit code 0 = all good.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../tfhe_fbs_map_b200/csrc/ntt.cuh"
#include "../../tfhe_fbs_map_b200/csrc/common.cuh"

static u32 bitrev(u32 x, int bits) { u32 r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

template <int LOGN> struct Tables {
    static constexpr int N = 1 << LOGN;
    std::vector<fq_tw> psi_rev, psi_inv_rev; u64 ninv;
    Tables() : psi_rev(N), psi_inv_rev(N) {
        u64 psi = fq_pow_host(7, (FQ_Q - 1) / (2ULL * N)), psi_inv = fq_pow_host(psi, FQ_Q - 2);
        for (int i = 0; i < N; i++) {
            u32 r = bitrev(i, LOGN);
            u64 w = fq_pow_host(psi, r), wi = fq_pow_host(psi_inv, r);
            psi_rev[i] = fq_tw{w, fq_shoup_host(w)}; psi_inv_rev[i] = fq_tw{wi, fq_shoup_host(wi)};
        }
        ninv = fq_pow_host(N, FQ_Q - 2);
    }
};
static u64 canon(u64 x) { return fq_csub(fq_csub(fq_csub(x, FQ_2Q), FQ_2Q), FQ_Q); }
template <int LOGN> void ref_fwd(const Tables<LOGN> &t, std::vector<u64> &a) {
    int N = 1 << LOGN, tt = N;
    for (int m = 1; m < N; m <<= 1) { tt >>= 1; for (int i = 0; i < m; i++) { u64 S = t.psi_rev[m + i].w;
        for (int j = 2 * i * tt; j < 2 * i * tt + tt; j++) { u64 U = a[j], V = fq_mul(a[j + tt], S); a[j] = fq_add(U, V); a[j + tt] = fq_sub(U, V); } } }
}
template <int LOGN> void ref_inv(const Tables<LOGN> &t, std::vector<u64> &a) {
    int N = 1 << LOGN, tt = 1;
    for (int m = N >> 1; m >= 1; m >>= 1) { for (int i = 0; i < m; i++) { u64 S = t.psi_inv_rev[m + i].w;
        for (int j = 2 * i * tt; j < 2 * i * tt + tt; j++) { u64 U = a[j], V = a[j + tt]; a[j] = fq_add(U, V); a[j + tt] = fq_mul(fq_sub(U, V), S); } } tt <<= 1; }
}
template <int LOGN, int PASS> void emu_fwd(const Tables<LOGN> &t, std::vector<u64> &arr) {
    using P = NttPlan<LOGN>;
    if constexpr (PASS < P::NPASS) {
        std::vector<u64> sm(P::N);
        for (int tau = 0; tau < P::T; tau++) {
            u64 x[8];
            for (int e = 0; e < 8; e++) x[e] = arr[P::idx(tau, e, P::fwd_lb(PASS))];
            ntt_fwd_pass<LOGN, PASS>(x, tau, t.psi_rev.data());
            for (int e = 0; e < 8; e++) sm[P::swz(P::idx(tau, e, P::fwd_lb(PASS)))] = x[e];   // swizzled store
        }
        for (int i = 0; i < P::N; i++) arr[i] = sm[P::swz(i)];                               // swizzled load
        emu_fwd<LOGN, PASS + 1>(t, arr);
    }
}
template <int LOGN, int PASS> void emu_inv(const Tables<LOGN> &t, std::vector<u64> &arr) {
    using P = NttPlan<LOGN>;
    if constexpr (PASS < P::NPASS) {
        std::vector<u64> sm(P::N);
        for (int tau = 0; tau < P::T; tau++) {
            u64 x[8];
            for (int e = 0; e < 8; e++) x[e] = arr[P::idx(tau, e, P::inv_lb(PASS))];
            ntt_inv_pass<LOGN, PASS>(x, tau, t.psi_inv_rev.data());
            for (int e = 0; e < 8; e++) sm[P::swz(P::idx(tau, e, P::inv_lb(PASS)))] = x[e];
        }
        for (int i = 0; i < P::N; i++) arr[i] = sm[P::swz(i)];
        emu_inv<LOGN, PASS + 1>(t, arr);
    }
}
template <int LOGN> int run() {
    using P = NttPlan<LOGN>; Tables<LOGN> t; int N = P::N, bad = 0;
    // swizzle is a permutation
    { std::vector<int> seen(N, 0); for (int i = 0; i < N; i++) seen[P::swz(i)]++; for (int i = 0; i < N; i++) if (seen[i] != 1) bad++; }
    // every pass layout is a permutation of [0,N)
    for (int p = 0; p < P::NPASS; p++) for (int lb : {P::fwd_lb(p), P::inv_lb(p)}) {
        std::vector<int> seen(N, 0); for (int tau = 0; tau < P::T; tau++) for (int e = 0; e < 8; e++) seen[P::idx(tau, e, lb)]++;
        for (int i = 0; i < N; i++) if (seen[i] != 1) bad++;
    }
    std::vector<u64> a(N), b(N);
    for (int i = 0; i < N; i++) { a[i] = fbs_rnd_uniform(42 + LOGN, 99, i); b[i] = fbs_rnd_uniform(43 + LOGN, 98, i); }
    std::vector<u64> r = a, e = a;
    ref_fwd<LOGN>(t, r); emu_fwd<LOGN, 0>(t, e);
    for (int i = 0; i < N; i++) { e[i] = canon(e[i]); if (r[i] != e[i]) bad++; }
    std::vector<u64> r2 = r, e2 = e;
    ref_inv<LOGN>(t, r2); emu_inv<LOGN, 0>(t, e2);
    for (int i = 0; i < N; i++) { if (e2[i] >= FQ_2Q) bad++; e2[i] = canon(e2[i]); if (r2[i] != e2[i]) bad++; if (fq_mul(e2[i], t.ninv) != a[i]) bad++; }
    // lazy inputs: the forward transform must accept anything in [0, 4Q)
    { std::vector<u64> lz = a; for (int i = 0; i < N; i++) lz[i] += (i % 4) * FQ_Q; emu_fwd<LOGN, 0>(t, lz);
      for (int i = 0; i < N; i++) if (canon(lz[i]) != r[i]) bad++; }
    // negacyclic product through the emulated transforms vs schoolbook (only for small N: O(N^2))
    if (N <= 512) {
        std::vector<u64> fa = a, fb = b, prod(N), sb(N, 0);
        emu_fwd<LOGN, 0>(t, fa); emu_fwd<LOGN, 0>(t, fb);
        // Montgomery path as in the blind-rotate kernel: fb -> (fb * 2^64 / N) canonical, product reduced lazily by REDC
        for (int i = 0; i < N; i++) {
            u64 bm = fq_mul(canon(fb[i]), fq_mul(FQ_R, t.ninv)), lo, hi;
            fq_mul_wide(fa[i], bm, lo, hi);
            prod[i] = fq_csub(fq_redc(lo, hi), FQ_2Q);
            if (prod[i] >= FQ_2Q) bad++;
        }
        emu_inv<LOGN, 0>(t, prod);
        for (int i = 0; i < N; i++) prod[i] = canon(prod[i]);
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { u64 pr = fq_mul(a[i], b[j]); int k = i + j;
            if (k < N) sb[k] = fq_add(sb[k], pr); else sb[k - N] = fq_sub(sb[k - N], pr); }
        for (int i = 0; i < N; i++) if (sb[i] != prod[i]) bad++;
    }
    printf("LOGN=%d bad=%d\n", LOGN, bad);
    return bad;
}
int main() {
    int bad = 0;
    // field sanity: reduce128 against __int128
    for (int i = 0; i < 200000; i++) {
        u64 a = fbs_rnd_uniform(1, 2, i), b = fbs_rnd_uniform(3, 4, i);
        if (i < 64) { u64 edge[8] = {0, 1, FQ_Q - 1, FQ_Q - 2, 0xFFFFFFFFULL, 0x100000000ULL, 0x3FFFFFFFFFFF0000ULL, 2}; a = edge[i % 8]; b = edge[i / 8]; }
        u64 want = (u64)(((unsigned __int128)a * b) % FQ_Q);
        if (fq_mul(a, b) != want) bad++;
        if (fq_add(a, b) != (u64)(((unsigned __int128)a + b) % FQ_Q)) bad++;
        if (fq_sub(a, b) != (u64)(((unsigned __int128)a + FQ_Q - b) % FQ_Q)) bad++;
        // Shoup: any 64-bit y, result in [0, 2Q) and congruent
        u64 y = fbs_rnd64(5, 6, i), sh = fq_mul_shoup(y, a, fq_shoup_host(a));
        if (sh >= FQ_2Q || sh % FQ_Q != (u64)(((unsigned __int128)a * (y % FQ_Q)) % FQ_Q)) bad++;
        // Montgomery reduction of a lazy product
        u64 lz = a + (i % 4) * FQ_Q, lo, hi; fq_mul_wide(lz, b, lo, hi);
        u64 rd = fq_redc(lo, hi);
        if (rd >= 3 * FQ_Q || ((unsigned __int128)(rd % FQ_Q) * FQ_R) % FQ_Q != want) bad++;
        // wide accumulators as in the key switch (hi < 2^22)
        u64 lo2 = fbs_rnd64(7, 8, i), hi2 = fbs_rnd64(9, 10, i) >> 42;
        if (fq_reduce128(lo2, hi2) != (u64)((((unsigned __int128)hi2 << 64) | lo2) % FQ_Q)) bad++;
    }
    printf("field bad=%d\n", bad);
    bad += run<8>(); bad += run<9>(); bad += run<10>(); bad += run<11>(); bad += run<12>();
    return bad ? 1 : 0;
}
