// Host emulation of the device arithmetic and NTT pass/layout/twiddle logic (tfhe_fbs_map_b200/csrc/fq.cuh, ntt.cuh).
// Loops over tau play the threads, array copies play the shared-memory transposes.  Compared against textbook
// per-prime negacyclic NTT loops and against a schoolbook product over the integers mod q.  Exit code 0 = all good.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../tfhe_fbs_map_b200/csrc/ntt.cuh"
#include "../../tfhe_fbs_map_b200/csrc/common.cuh"

typedef unsigned __int128 u128;
static u32 bitrev(u32 x, int bits) { u32 r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }
static const u32 PR[2] = {FQ_P1, FQ_P2};

template <int LOGN> struct Tables {
    static constexpr int N = 1 << LOGN;
    std::vector<fq_tw> psi_rev, psi_inv_rev; u32 ninv[2];
    std::vector<u32> w[2], wi[2];
    Tables() : psi_rev(N), psi_inv_rev(N) {
        for (int l = 0; l < 2; l++) {
            const u32 p = PR[l];
            const u64 psi = pow_mod_host(3, (p - 1) / (2ULL * N), p), psi_inv = pow_mod_host(psi, p - 2, p);
            w[l].resize(N); wi[l].resize(N);
            for (int i = 0; i < N; i++) { u32 r = bitrev(i, LOGN); w[l][i] = (u32)pow_mod_host(psi, r, p); wi[l][i] = (u32)pow_mod_host(psi_inv, r, p); }
            ninv[l] = (u32)pow_mod_host(N, p - 2, p);
        }
        for (int i = 0; i < N; i++) {
            psi_rev[i] = fq_tw{w[0][i], shoup32_host(w[0][i], FQ_P1), w[1][i], shoup32_host(w[1][i], FQ_P2)};
            psi_inv_rev[i] = fq_tw{wi[0][i], shoup32_host(wi[0][i], FQ_P1), wi[1][i], shoup32_host(wi[1][i], FQ_P2)};
        }
    }
};
static u32 mulp(u32 a, u32 b, u32 p) { return (u32)((u64)a * b % p); }
template <int LOGN> void ref_fwd(const Tables<LOGN> &t, std::vector<u32> &a, int l) {
    int N = 1 << LOGN, tt = N; u32 p = PR[l];
    for (int m = 1; m < N; m <<= 1) { tt >>= 1; for (int i = 0; i < m; i++) { u32 S = t.w[l][m + i];
        for (int j = 2 * i * tt; j < 2 * i * tt + tt; j++) { u32 U = a[j], V = mulp(a[j + tt], S, p); a[j] = (U + V) % p; a[j + tt] = (U + p - V) % p; } } }
}
template <int LOGN> void ref_inv(const Tables<LOGN> &t, std::vector<u32> &a, int l) {
    int N = 1 << LOGN, tt = 1; u32 p = PR[l];
    for (int m = N >> 1; m >= 1; m >>= 1) { for (int i = 0; i < m; i++) { u32 S = t.wi[l][m + i];
        for (int j = 2 * i * tt; j < 2 * i * tt + tt; j++) { u32 U = a[j], V = a[j + tt]; a[j] = (U + V) % p; a[j + tt] = mulp((U + p - V) % p, S, p); } } tt <<= 1; }
}
template <int LOGN, int PASS> void emu_fwd(const Tables<LOGN> &t, std::vector<u64> &arr) {
    using P = NttPlan<LOGN>;
    if constexpr (PASS < P::NPASS) {
        std::vector<u64> sm(P::N);
        for (int tau = 0; tau < P::T; tau++) {
            rns2 x[8];
            for (int e = 0; e < 8; e++) x[e] = rns_unpack(arr[P::idx(tau, e, P::fwd_lb(PASS))]);
            ntt_fwd_pass<LOGN, PASS>(x, tau, t.psi_rev.data());
            for (int e = 0; e < 8; e++) sm[P::swz(P::idx(tau, e, P::fwd_lb(PASS)))] = rns_pack(x[e]);   // swizzled store
        }
        for (int i = 0; i < P::N; i++) arr[i] = sm[P::swz(i)];                                         // swizzled load
        emu_fwd<LOGN, PASS + 1>(t, arr);
    }
}
template <int LOGN, int PASS> void emu_inv(const Tables<LOGN> &t, std::vector<u64> &arr) {
    using P = NttPlan<LOGN>;
    if constexpr (PASS < P::NPASS) {
        std::vector<u64> sm(P::N);
        for (int tau = 0; tau < P::T; tau++) {
            rns2 x[8];
            for (int e = 0; e < 8; e++) x[e] = rns_unpack(arr[P::idx(tau, e, P::inv_lb(PASS))]);
            ntt_inv_pass<LOGN, PASS>(x, tau, t.psi_inv_rev.data());
            for (int e = 0; e < 8; e++) sm[P::swz(P::idx(tau, e, P::inv_lb(PASS)))] = rns_pack(x[e]);
        }
        for (int i = 0; i < P::N; i++) arr[i] = sm[P::swz(i)];
        emu_inv<LOGN, PASS + 1>(t, arr);
    }
}
static rns2 canon4(rns2 v) { v.a %= FQ_P1; v.b %= FQ_P2; return v; }
template <int LOGN> int run() {
    using P = NttPlan<LOGN>; Tables<LOGN> t; int N = P::N, bad = 0;
    { std::vector<int> seen(N, 0); for (int i = 0; i < N; i++) seen[P::swz(i)]++; for (int i = 0; i < N; i++) if (seen[i] != 1) bad++; }
    for (int p = 0; p < P::NPASS; p++) for (int lb : {P::fwd_lb(p), P::inv_lb(p)}) {
        std::vector<int> seen(N, 0); for (int tau = 0; tau < P::T; tau++) for (int e = 0; e < 8; e++) seen[P::idx(tau, e, lb)]++;
        for (int i = 0; i < N; i++) if (seen[i] != 1) bad++;
    }
    std::vector<u64> a(N), b(N);                         // integers mod q
    for (int i = 0; i < N; i++) { a[i] = fbs_rnd_uniform(42 + LOGN, 99, i); b[i] = fbs_rnd_uniform(43 + LOGN, 98, i); }
    std::vector<u64> e(N);
    for (int i = 0; i < N; i++) { rns2 v = rns_from_int(a[i]); v.a += (i % 4) * FQ_P1; v.b += ((i + 1) % 4) * FQ_P2; e[i] = rns_pack(v); }  // lazy inputs < 4p
    emu_fwd<LOGN, 0>(t, e);
    for (int l = 0; l < 2; l++) {
        std::vector<u32> r(N); for (int i = 0; i < N; i++) r[i] = (u32)(a[i] % PR[l]);
        ref_fwd<LOGN>(t, r, l);
        for (int i = 0; i < N; i++) { rns2 v = rns_unpack(e[i]); u32 g = l ? v.b : v.a; if (g >= 4ULL * PR[l] || g % PR[l] != r[i]) bad++; }
        std::vector<u32> r2 = r; ref_inv<LOGN>(t, r2, l);
        for (int i = 0; i < N; i++) if (mulp(r2[i], t.ninv[l], PR[l]) != a[i] % PR[l]) bad++;
    }
    // inverse on canonical-ish (<2p) inputs
    std::vector<u64> e2(N);
    for (int i = 0; i < N; i++) { rns2 v = canon4(rns_unpack(e[i])); v.a += (i & 1) * FQ_P1; v.b += ((i >> 1) & 1) * FQ_P2; e2[i] = rns_pack(v); }
    emu_inv<LOGN, 0>(t, e2);
    for (int i = 0; i < N; i++) { rns2 v = rns_unpack(e2[i]); if (v.a >= 2 * FQ_P1 || v.b >= 2 * FQ_P2) bad++;
        if (mulp(v.a % FQ_P1, t.ninv[0], FQ_P1) != a[i] % FQ_P1 || mulp(v.b % FQ_P2, t.ninv[1], FQ_P2) != a[i] % FQ_P2) bad++; }
    // negacyclic product as in the blind-rotate kernel: key in Montgomery form with 1/N folded, REDC of a lazy product,
    // inverse transform, CRT back to the integer -- against the schoolbook product mod q (small N only: O(N^2))
    if (N <= 512) {
        std::vector<u64> fa(N), fb(N), prod(N), sb(N, 0);
        for (int i = 0; i < N; i++) { fa[i] = rns_pack(rns_from_int(a[i])); fb[i] = rns_pack(rns_from_int(b[i])); }
        emu_fwd<LOGN, 0>(t, fa); emu_fwd<LOGN, 0>(t, fb);
        const u32 m1 = mulp((u32)((1ULL << 32) % FQ_P1), t.ninv[0], FQ_P1), m2 = mulp((u32)((1ULL << 32) % FQ_P2), t.ninv[1], FQ_P2);
        const u32 pin1 = 0xFFFE7FFFu, pin2 = 0xAFFF3FFFu;
        for (int i = 0; i < N; i++) {
            rns2 x = rns_unpack(fa[i]), y = canon4(rns_unpack(fb[i]));
            u32 k1 = mulp(y.a, m1, FQ_P1), k2 = mulp(y.b, m2, FQ_P2);
            rns2 o; o.a = r32_fold(r32_redc((u64)x.a * k1, FQ_P1, pin1), 2 * FQ_P1); o.b = r32_fold(r32_redc((u64)x.b * k2, FQ_P2, pin2), 2 * FQ_P2);
            if (o.a >= 2 * FQ_P1 || o.b >= 2 * FQ_P2) bad++;
            prod[i] = rns_pack(o);
        }
        emu_inv<LOGN, 0>(t, prod);
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { u64 pr = fq_mul(a[i], b[j]); int k = i + j;
            if (k < N) sb[k] = fq_add(sb[k], pr); else sb[k - N] = fq_sub(sb[k - N], pr); }
        for (int i = 0; i < N; i++) { rns2 v = canon4(rns_unpack(prod[i])); if (rns_to_int(v) != sb[i]) bad++; }
    }
    // the inverse pass reading the FORWARD table mirrored (blind-rotate kernels) gives the same residues
    for (int tau = 0; tau < P::T; tau++) {
        rns2 x1[1][8], x2[1][8];
        for (int el = 0; el < 8; el++) { rns2 v = canon4(rns_unpack(e2[(tau * 8 + el) % N])); v.a += FQ_P1 * (el & 1); x1[0][el] = x2[0][el] = v; }
        ntt_inv_pass_n<LOGN, 1, 1, false>(x1, tau, t.psi_inv_rev.data());
        ntt_inv_pass_n<LOGN, 1, 1, true>(x2, tau, t.psi_rev.data());
        for (int el = 0; el < 8; el++) if (x1[0][el].a % FQ_P1 != x2[0][el].a % FQ_P1 || x1[0][el].b % FQ_P2 != x2[0][el].b % FQ_P2 || x2[0][el].a >= 2 * FQ_P1 || x2[0][el].b >= 2 * FQ_P2) bad++;
    }
    // GF(2)-linear addressing used by the kernels
    for (int tau = 0; tau < P::T; tau++) for (int el = 0; el < 8; el++) for (int p = 0; p < P::NPASS; p++) for (int lb : {P::fwd_lb(p), P::inv_lb(p)})
        if (8u * (u32)P::swz(P::idx(tau, el, lb)) != (P::tau_boff(tau, lb) ^ P::elem_boff(el, lb))) bad++;
    printf("LOGN=%d bad=%d\n", LOGN, bad);
    return bad;
}
// ---- cluster-split transform (ntt.cuh: ntt_cross_fwd / ntt_cross_inv + local tables), as k_blind_rotate_cl runs it:
// CTA loops play the cluster, the inbox arrays play the distributed-shared-memory exchange
template <int LOGNS, int PASS, bool INV> void emu_local(const std::vector<fq_tw> &tw, std::vector<u64> &arr) {
    using P = NttPlan<LOGNS>;
    if constexpr (PASS < P::NPASS) {
        std::vector<u64> sm(P::N);
        const int lb = INV ? P::inv_lb(PASS) : P::fwd_lb(PASS);
        for (int tau = 0; tau < P::T; tau++) {
            rns2 x[1][8];
            for (int e = 0; e < 8; e++) x[0][e] = rns_unpack(arr[P::idx(tau, e, lb)]);
            if (INV) ntt_inv_pass_n<LOGNS, PASS, 1, true>(x, tau, tw.data()); else ntt_fwd_pass_n<LOGNS, PASS, 1>(x, tau, tw.data());
            for (int e = 0; e < 8; e++) sm[P::swz(P::idx(tau, e, lb))] = rns_pack(x[0][e]);
        }
        for (int i = 0; i < P::N; i++) arr[i] = sm[P::swz(i)];
        emu_local<LOGNS, PASS + 1, INV>(tw, arr);
    }
}
template <int LOGN, int LOGC> int run_cluster() {
    constexpr int N = 1 << LOGN, C = 1 << LOGC, LOGNS = LOGN - LOGC, Ns = N / C, Ts = Ns / 8, R = 8 / C;
    Tables<LOGN> t; int bad = 0;
    std::vector<u64> a(N), lazy(N);
    for (int i = 0; i < N; i++) { a[i] = fbs_rnd_uniform(52 + LOGN, 9 + LOGC, i); rns2 v = rns_from_int(a[i]); v.a += (i % 2) * FQ_P1; v.b += ((i + 1) % 2) * FQ_P2; lazy[i] = rns_pack(v); }  // digits: < 2p
    std::vector<u64> full = lazy;
    emu_fwd<LOGN, 0>(t, full);                                           // the one-CTA transform is the reference here
    // forward: cross stages on registers, exchange, local passes with the local table
    std::vector<std::vector<u64>> inbox(C, std::vector<u64>(8 * Ts));
    for (int c = 0; c < C; c++) for (int tau = 0; tau < Ts; tau++) {
        rns2 x[8];
        for (int h = 0; h < C; h++) for (int ri = 0; ri < R; ri++) x[h * R + ri] = rns_unpack(lazy[h * Ns + c * (Ns / C) + ri * Ts + tau]);
        ntt_cross_fwd<LOGC>(x, t.psi_rev.data());
        for (int h = 0; h < C; h++) for (int ri = 0; ri < R; ri++) inbox[h][(c * R + ri) * Ts + tau] = rns_pack(x[h * R + ri]);
    }
    std::vector<std::vector<u64>> spec(C);
    for (int h = 0; h < C; h++) {
        std::vector<fq_tw> twf(Ns);
        for (int i = 1; i < Ns; i++) twf[i] = t.psi_rev[ntt_local_src(i, C + h)];
        std::vector<u64> arr(Ns);
        for (int tau = 0; tau < Ts; tau++) for (int e = 0; e < 8; e++) arr[tau + e * Ts] = inbox[h][e * Ts + tau];
        emu_local<LOGNS, 0, false>(twf, arr);
        for (int i = 0; i < Ns; i++) { rns2 v = rns_unpack(arr[i]), w = rns_unpack(full[h * Ns + i]);
            if (v.a >= 4ULL * FQ_P1 || v.b >= 4ULL * FQ_P2 || v.a % FQ_P1 != w.a % FQ_P1 || v.b % FQ_P2 != w.b % FQ_P2) bad++; }
        spec[h] = arr;
    }
    // inverse: local mirrored passes, exchange, cross stages; against the textbook inverse (unscaled) of the canonical spectrum
    std::vector<u64> sp2(N);
    for (int i = 0; i < N; i++) { rns2 v = canon4(rns_unpack(full[i])); v.a += (i & 1) * FQ_P1; v.b += ((i >> 1) & 1) * FQ_P2; sp2[i] = rns_pack(v); }
    std::vector<std::vector<u64>> inbox2(C, std::vector<u64>(8 * Ts));
    for (int h = 0; h < C; h++) {
        std::vector<fq_tw> twi(Ns);
        for (int i = 1; i < Ns; i++) twi[i] = t.psi_rev[ntt_local_src(i, 2 * C - 1 - h)];
        std::vector<u64> arr(sp2.begin() + h * Ns, sp2.begin() + (h + 1) * Ns);
        emu_local<LOGNS, 0, true>(twi, arr);
        for (int tau = 0; tau < Ts; tau++) for (int e = 0; e < 8; e++) inbox2[e / R][(h * R + e % R) * Ts + tau] = arr[tau + e * Ts];
    }
    std::vector<u64> back(N);
    for (int c = 0; c < C; c++) for (int tau = 0; tau < Ts; tau++) {
        rns2 x[8];
        for (int e = 0; e < 8; e++) x[e] = rns_unpack(inbox2[c][e * Ts + tau]);
        ntt_cross_inv<LOGC>(x, t.psi_rev.data());
        for (int h = 0; h < C; h++) for (int ri = 0; ri < R; ri++) back[h * Ns + c * (Ns / C) + ri * Ts + tau] = rns_pack(x[h * R + ri]);
    }
    for (int i = 0; i < N; i++) { rns2 v = rns_unpack(back[i]); if (v.a >= 2 * FQ_P1 || v.b >= 2 * FQ_P2) bad++;
        if (mulp(v.a % FQ_P1, t.ninv[0], FQ_P1) != a[i] % FQ_P1 || mulp(v.b % FQ_P2, t.ninv[1], FQ_P2) != a[i] % FQ_P2) bad++; }
    printf("cluster LOGN=%d C=%d bad=%d\n", LOGN, C, bad);
    return bad;
}
int main() {
    int bad = 0;
    if ((u128)FQ_P1 * FQ_P2 != FQ_Q) bad++;
    if (((u64)FQ_P1 * FQ_P1INV_P2) % FQ_P2 != 1) bad++;
    if ((u32)(FQ_P1 * 0xFFFE7FFFu) != 0xFFFFFFFFu || (u32)(FQ_P2 * 0xAFFF3FFFu) != 0xFFFFFFFFu) bad++;
    if (FBS_K63 != (1ULL << 63) / FQ_P2) bad++;
    if (FQ_RQ != (u64)((((u128)1) << 123) / FQ_Q)) bad++;
    if (FQ_QN != (FQ_Q << 4) || FQ_QNV != (u64)((~(u128)0) / FQ_QN - (((u128)1) << 64))) bad++;
    // the rounding must be unbiased: a bias of a fraction of a step accumulates over the n*N decompositions of a PBS
    for (int bits : {12, 15, 23}) {
        double sum = 0; const int M = 200000;
        for (int i = 0; i < M; i++) {
            u64 x = fbs_rnd_uniform(77, 78, i), y = fbs_round_top(x, bits);
            double exact = (double)x / (double)FQ_Q * (double)(1ULL << bits), d = (double)y - exact;
            if (d > (double)(1ULL << (bits - 1))) d -= (double)(1ULL << bits);       // wrap-around at the top
            if (d < -(double)(1ULL << (bits - 1))) d += (double)(1ULL << bits);
            sum += d;
        }
        if (sum / M > 0.005 || sum / M < -0.005) { printf("rounding bias %d bits: %f\n", bits, sum / M); bad++; }
    }
    for (int i = 0; i < 300000; i++) {
        u64 a = fbs_rnd_uniform(1, 2, i), b = fbs_rnd_uniform(3, 4, i);
        if (i < 64) { u64 edge[8] = {0, 1, FQ_Q - 1, FQ_Q - 2, 0xFFFFFFFFULL, 0x100000000ULL, FQ_P1, FQ_P2}; a = edge[i % 8]; b = edge[i / 8]; }
        u64 want = (u64)(((u128)a * b) % FQ_Q);
        if (fq_mul(a, b) != want) bad++;
        if (fq_add(a, b) != (u64)(((u128)a + b) % FQ_Q)) bad++;
        if (fq_sub(a, b) != (u64)(((u128)a + FQ_Q - b) % FQ_Q)) bad++;
        // wide accumulators as in the key switch (hi < 2^22) and arbitrary hi < q
        u64 lo2 = fbs_rnd64(7, 8, i), hi2 = fbs_rnd64(9, 10, i) >> 42;
        if (fq_reduce128(lo2, hi2) != (u64)((((u128)hi2 << 64) | lo2) % FQ_Q)) bad++;
        u64 hi3 = fbs_rnd_uniform(11, 12, i);
        if (fq_reduce128(lo2, hi3) != (u64)((((u128)hi3 << 64) | lo2) % FQ_Q)) bad++;
        // RNS round trip and CRT
        rns2 r = rns_from_int(a);
        if (rns_to_int(r) != a) bad++;
        if (rns_crt_hi(r) != (u32)(a / FQ_P1)) bad++;
        rns2 s = rns_sub(rns_from_int(a), rns_from_int(b));
        if (rns_to_int(s) != fq_sub(a, b)) bad++;
        // 32-bit Shoup: any y, result in [0,2p) and congruent
        u32 y = (u32)fbs_rnd64(5, 6, i), w = (u32)(a % FQ_P2), sh = r32_mul_shoup(y, w, shoup32_host(w, FQ_P2), FQ_P2);
        if (sh >= 2 * FQ_P2 || sh % FQ_P2 != (u32)((u64)w * (y % FQ_P2) % FQ_P2)) bad++;
        // one-level digit (fbs_digit1_t, the blind-rotation kernels' fast path) == rounding + balanced digit of the spec,
        // r32_csub == conditional subtraction, borrow-free REDC == acc * 2^-32 mod p within (0, hi + p]
        for (int beta : {8, 12, 22, 23, 24}) {
            const u32 t = rns_crt_hi(r);
            int dref[1]; fbs_balanced_digits<1>(fbs_round_top_t(t, r.a, beta), beta, dref);
            if (fbs_digit1_t(t, r.a, beta, 1ULL << (62 - beta)) != dref[0]) bad++;
        }
        { const u32 xx = (u32)fbs_rnd64(13, 14, i); if (r32_csub(xx, FQ_P1) != (xx >= FQ_P1 ? xx - FQ_P1 : xx)) bad++; }
        { const u64 acc = fbs_rnd64(15, 16, i) >> 1; const u32 rr = r32_redc(acc, FQ_P2, 0xAFFF3FFFu);
          if (rr % FQ_P2 != (u32)((u128)(acc % FQ_P2) * pow_mod_host((1ULL << 32) % FQ_P2, FQ_P2 - 2, FQ_P2) % FQ_P2) || rr > (u32)(acc >> 32) + FQ_P2) bad++; }
        // rounding: both definitions stay within one unit of the exact quotient
        for (int bits : {12, 15, 23, 24, 30}) {
            u64 yq = fbs_round_top(a, bits);
            u128 exact2 = ((u128)a << (bits + 1)) / FQ_Q;              // floor(2 * a * 2^bits / q)
            u64 nearest = (u64)((exact2 + 1) >> 1) & ((1ULL << bits) - 1);
            u64 d = (yq - nearest) & ((1ULL << bits) - 1);
            if (!(d == 0 || d == 1 || d == ((1ULL << bits) - 1))) bad++;
        }
    }
    printf("field bad=%d\n", bad);
    bad += run<8>(); bad += run<9>(); bad += run<10>(); bad += run<11>(); bad += run<12>();
    bad += run_cluster<11, 1>(); bad += run_cluster<11, 2>(); bad += run_cluster<11, 3>(); bad += run_cluster<10, 1>(); bad += run_cluster<10, 2>(); bad += run_cluster<9, 1>();
    return bad ? 1 : 0;
}
