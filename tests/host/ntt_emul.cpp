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
    std::vector<u64> psi_rev, psi_inv_rev; u64 ninv;
    Tables() : psi_rev(N), psi_inv_rev(N) {
        u64 psi = gl_pow_host(7, (GL_P - 1) / (2ULL * N)), psi_inv = gl_pow_host(psi, GL_P - 2);
        for (int i = 0; i < N; i++) { u32 r = bitrev(i, LOGN); psi_rev[i] = gl_pow_host(psi, r); psi_inv_rev[i] = gl_pow_host(psi_inv, r); }
        ninv = gl_pow_host(N, GL_P - 2);
    }
};
template <int LOGN> void ref_fwd(const Tables<LOGN> &t, std::vector<u64> &a) {
    int N = 1 << LOGN, tt = N;
    for (int m = 1; m < N; m <<= 1) { tt >>= 1; for (int i = 0; i < m; i++) { u64 S = t.psi_rev[m + i];
        for (int j = 2 * i * tt; j < 2 * i * tt + tt; j++) { u64 U = a[j], V = gl_mul(a[j + tt], S); a[j] = gl_add(U, V); a[j + tt] = gl_sub(U, V); } } }
}
template <int LOGN> void ref_inv(const Tables<LOGN> &t, std::vector<u64> &a) {
    int N = 1 << LOGN, tt = 1;
    for (int m = N >> 1; m >= 1; m >>= 1) { for (int i = 0; i < m; i++) { u64 S = t.psi_inv_rev[m + i];
        for (int j = 2 * i * tt; j < 2 * i * tt + tt; j++) { u64 U = a[j], V = a[j + tt]; a[j] = gl_add(U, V); a[j + tt] = gl_mul(gl_sub(U, V), S); } } tt <<= 1; }
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
    for (int i = 0; i < N; i++) if (r[i] != e[i]) bad++;
    std::vector<u64> r2 = r, e2 = e;
    ref_inv<LOGN>(t, r2); emu_inv<LOGN, 0>(t, e2);
    for (int i = 0; i < N; i++) { if (r2[i] != e2[i]) bad++; if (gl_mul(e2[i], t.ninv) != a[i]) bad++; }
    // negacyclic product through the emulated transforms vs schoolbook (only for small N: O(N^2))
    if (N <= 512) {
        std::vector<u64> fa = a, fb = b, prod(N), sb(N, 0);
        emu_fwd<LOGN, 0>(t, fa); emu_fwd<LOGN, 0>(t, fb);
        for (int i = 0; i < N; i++) prod[i] = gl_mul(gl_mul(fa[i], fb[i]), t.ninv);
        emu_inv<LOGN, 0>(t, prod);
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { u64 pr = gl_mul(a[i], b[j]); int k = i + j;
            if (k < N) sb[k] = gl_add(sb[k], pr); else sb[k - N] = gl_sub(sb[k - N], pr); }
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
        if (i < 8) { u64 edge[8] = {0, 1, GL_P - 1, GL_P - 2, 0xFFFFFFFFULL, 0x100000000ULL, 0xFFFFFFFF00000000ULL, 2}; a = edge[i]; b = edge[(i * 3) % 8]; }
        u64 want = (u64)(((unsigned __int128)a * b) % GL_P);
        if (gl_mul(a, b) != want) bad++;
        if (gl_add(a, b) != (u64)(((unsigned __int128)a + b) % GL_P)) bad++;
        if (gl_sub(a, b) != (u64)(((unsigned __int128)a + GL_P - b) % GL_P)) bad++;
    }
    printf("field bad=%d\n", bad);
    bad += run<8>(); bad += run<9>(); bad += run<10>(); bad += run<11>(); bad += run<12>();
    return bad ? 1 : 0;
}
