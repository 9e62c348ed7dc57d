// api.cu -- C ABI of the B200 encrypted executor (include/fbs_b200.h): contexts, key generation, level
// scheduler, host-buffer evaluation, parity taps.  No CPU fallback anywhere: every compute entry point
// launches the CUDA kernels in kernels.cuh or fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/fbs_b200.h"
#include "kernels.cuh"

// ------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(FBS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)
#define CKR(expr) do { int r_ = (expr); if (r_ != FBS_OK) return r_; } while (0)

extern "C" const char *fbs_last_error(void) { return g_err.c_str(); }
extern "C" int fbs_abi_version(void) { return 1; }

// ------------------------------------------------------------------------------------------------------
// blind-rotate kernel variants
// ------------------------------------------------------------------------------------------------------
typedef cudaError_t (*br_launch_fn)(const BRArgs &, long long jobs, size_t smem, cudaStream_t);
struct BRVariant { int logN, k, l; bool bsk_smem; int pb, tp, threads; size_t (*smem)(int n); br_launch_fn launch; cudaError_t (*prepare)(size_t smem); int unr; };

template <int LOGN, int K, int L, bool SM, int PB, int TP>
static cudaError_t br_launch(const BRArgs &a, long long jobs, size_t smem, cudaStream_t st)
{
    const long long grid = (jobs - a.job_begin + PB - 1) / PB;
    k_blind_rotate<LOGN, K, L, SM, PB, TP><<<(unsigned)grid, BRCfg<LOGN, K, L, SM, PB, TP>::THREADS, smem, st>>>(a);
    return cudaGetLastError();
}
template <int LOGN, int K, int L, bool SM, int PB, int TP>
static cudaError_t br_prepare(size_t smem)
{
    return cudaFuncSetAttribute(k_blind_rotate<LOGN, K, L, SM, PB, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <int LOGN, int K, int L, bool SM, int PB, int TP> static size_t br_smem(int n) { return BRCfg<LOGN, K, L, SM, PB, TP>::smem_bytes(n); }
#define BRV(LOGN, K, L, SM, PB, TP) { LOGN, K, L, SM, PB, TP, BRCfg<LOGN, K, L, SM, PB, TP>::THREADS, br_smem<LOGN, K, L, SM, PB, TP>, br_launch<LOGN, K, L, SM, PB, TP>, br_prepare<LOGN, K, L, SM, PB, TP>, 1 }
// key-unrolled kernels (bsk_unroll = M = 2 or 3 key bits per step, one decomposition level)
template <int LOGN, int K, int PB, int TP, int M>
static cudaError_t br2_launch(const BRArgs &a, long long jobs, size_t smem, cudaStream_t st)
{
    const long long grid = (jobs - a.job_begin + PB - 1) / PB;
    k_blind_rotate2<LOGN, K, PB, TP, M><<<(unsigned)grid, BR2Cfg<LOGN, K, PB, TP, M>::THREADS, smem - BR2Cfg<LOGN, K, PB, TP, M>::static_b, st>>>(a);   // smem = static + dynamic
    return cudaGetLastError();
}
template <int LOGN, int K, int PB, int TP, int M>
static cudaError_t br2_prepare(size_t smem)
{
    return cudaFuncSetAttribute(k_blind_rotate2<LOGN, K, PB, TP, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem - BR2Cfg<LOGN, K, PB, TP, M>::static_b));
}
template <int LOGN, int K, int PB, int TP, int M> static size_t br2_smem(int n) { return BR2Cfg<LOGN, K, PB, TP, M>::smem_bytes(n); }
#define BRV2(LOGN, K, PB, TP, M) { LOGN, K, 1, false, PB, TP, BR2Cfg<LOGN, K, PB, TP, M>::THREADS, br2_smem<LOGN, K, PB, TP, M>, br2_launch<LOGN, K, PB, TP, M>, br2_prepare<LOGN, K, PB, TP, M>, M }
#ifndef FBS_SETA_TP
#define FBS_SETA_TP 2     /* bootstraps per thread in the set-A kernel: 2 = every thread carries both bootstraps of its CTA */
#endif
#ifndef FBS_A3_TP
#define FBS_A3_TP 2       /* 1 = 1024 threads, one bootstrap per thread, 64 registers (experiment) */
#endif
static const BRVariant g_br_variants[] = {
    BRV(11, 1, 1, true, 2, FBS_SETA_TP),   // set A: two bootstraps per CTA share the TMA-streamed BSK row (192 KB shared memory)
    BRV(11, 1, 1, true, 1, 1),             // set A, one bootstrap per CTA: used when a launch has no more jobs than SMs
    BRV(11, 1, 2, false, 1, 1),            // set C (row does not fit shared memory next to the accumulator: BSK read from L2)
    BRV(10, 2, 1, true, 1, 1),             // set S
    BRV(8, 1, 2, true, 2, 2), BRV(8, 2, 1, true, 2, 1), BRV(9, 1, 1, true, 2, 2), BRV(10, 1, 3, true, 1, 1),   // toy sets (tests)
    BRV2(11, 1, 2, 2, 2), BRV2(11, 1, 1, 1, 2),    // set A2 (two key bits per step)
    BRV2(9, 1, 2, 2, 2), BRV2(8, 2, 2, 1, 2),      // toy3u / toy7u, toy2u
    BRV2(11, 1, 2, FBS_A3_TP, 3), BRV2(11, 1, 1, 1, 3),    // set A3 (three key bits per step)
    BRV2(9, 1, 2, 2, 3),                           // toy3v
};

// cluster-split low-latency kernels (one bootstrap over C = 2^LOGC CTAs), for launches that would leave SMs idle; PS = prime-split
// variant (two threads per ring element, one per RNS prime: k_blind_rotate_cs); OCC = CTAs per SM its shared memory / registers are
// sized for (2: twice the co-resident clusters, shallower key ring, 128 registers)
template <int LOGN, int K, int M, int LOGC, bool PS, int OCC> struct BrcKernel {
    using Cf = BRCCfg<LOGN, K, M, LOGC, OCC>;
    static auto fn() { if constexpr (PS) return k_blind_rotate_cs<LOGN, K, M, LOGC, OCC>; else return k_blind_rotate_cl<LOGN, K, M, LOGC>; }
    static void config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *at, unsigned clusters, size_t smem, cudaStream_t st)
    {
        cfg = cudaLaunchConfig_t{};
        cfg.gridDim = dim3(clusters << LOGC); cfg.blockDim = dim3((PS ? 2 : 1) * Cf::THREADS); cfg.dynamicSmemBytes = smem - Cf::static_b; cfg.stream = st;      // smem = static + dynamic
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1 << LOGC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
    }
    static cudaError_t launch(const BRArgs &a, long long jobs, size_t smem, cudaStream_t st)
    {
        cudaLaunchConfig_t cfg; cudaLaunchAttribute at[1];
        config(cfg, at, (unsigned)(jobs - a.job_begin), smem, st);
        return cudaLaunchKernelEx(&cfg, fn(), a);
    }
    static cudaError_t prepare(size_t smem) { return cudaFuncSetAttribute(fn(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem - Cf::static_b)); }
    static size_t smem(int n) { return Cf::smem_bytes(n); }
    // how many clusters of this kernel the device can hold at once (the hardware may strand SMs: 33 clusters of 4 on a 148-SM B200)
    static cudaError_t max_clusters(size_t smem, int *out)
    {
        cudaLaunchConfig_t cfg; cudaLaunchAttribute at[1];
        config(cfg, at, 1, smem, nullptr);
        return cudaOccupancyMaxActiveClusters(out, fn(), &cfg);
    }
};
struct BRCVariant { int logN, k, unr, logC; bool ps; int occ; size_t (*smem)(int n); br_launch_fn launch; cudaError_t (*prepare)(size_t smem); cudaError_t (*max_clusters)(size_t smem, int *out); };
#define BRCV(LOGN, K, M, LOGC, PS, OCC) { LOGN, K, M, LOGC, PS, OCC, BrcKernel<LOGN, K, M, LOGC, PS, OCC>::smem, BrcKernel<LOGN, K, M, LOGC, PS, OCC>::launch, BrcKernel<LOGN, K, M, LOGC, PS, OCC>::prepare, BrcKernel<LOGN, K, M, LOGC, PS, OCC>::max_clusters }
static const BRCVariant g_brc_variants[] = {
    BRCV(11, 1, 3, 1, false, 1), BRCV(11, 1, 3, 2, false, 1), BRCV(11, 1, 3, 3, false, 1),      // sets A3 / toy5v: clusters of 2, 4, 8
    BRCV(11, 1, 2, 1, false, 1), BRCV(11, 1, 2, 2, false, 1), BRCV(11, 1, 2, 3, false, 1),      // sets A2 / toy5u
    BRCV(11, 1, 3, 1, true, 1), BRCV(11, 1, 3, 2, true, 1), BRCV(11, 1, 3, 3, true, 1),         // prime-split twins
    BRCV(11, 1, 2, 1, true, 1), BRCV(11, 1, 2, 2, true, 1), BRCV(11, 1, 2, 3, true, 1),
};

typedef void (*ntt_launch_fn)(const u64 *, u64 *, int, const fq_tw *, const fq_tw *, u32, u32, long long, cudaStream_t, int);
template <int LOGN>
static void ntt_launch(const u64 *in, u64 *out, int mode, const fq_tw *pr, const fq_tw *pir, u32 s1, u32 s2, long long count, cudaStream_t st, int g1)
{
    k_ntt<LOGN><<<(unsigned)count, NttPlan<LOGN>::T, 0, st>>>(in, out, mode, pr, pir, s1, s2, g1);
}
static ntt_launch_fn ntt_for(int logN)
{
    switch (logN) { case 8: return ntt_launch<8>; case 9: return ntt_launch<9>; case 10: return ntt_launch<10>; case 11: return ntt_launch<11>; }
    return nullptr;
}

// ------------------------------------------------------------------------------------------------------
struct fbs_ctx {
    fbs_params P; int device = 0; u64 seed = 0; int logN = 0, sm_count = 0;
    bool have_keys = false;
    const BRVariant *br = nullptr; size_t br_smem = 0;        // widest variant (most bootstraps per CTA)
    const BRVariant *br1 = nullptr; size_t br1_smem = 0;      // one bootstrap per CTA, for launches with <= sm_count jobs
    // [kind][log2 C], kind 0 = packed, 1 = prime-split, 2 = prime-split sized for two CTAs per SM: one bootstrap per cluster of C CTAs ...
    const BRCVariant *brc[3][4] = {}; size_t brc_smem[3][4] = {};
    int brc_max[3][4] = {};                                   // ... for launches of at most this many jobs (co-resident clusters)
    int cluster_mode = 0;                                     // 0 auto, 1 never, 2 / 4 / 8 force that cluster size, 12 / 14 / 18 prime-split, 24 prime-split 2 CTAs/SM
    u8 *d_s_lwe = nullptr, *d_s_big = nullptr;
    u64 *d_ksk = nullptr, *d_colsum = nullptr, *d_bsk = nullptr, *d_bsk_coef = nullptr;
    u8 *d_kbt = nullptr;                                       // byte-transposed KSK for the tensor-core key switch
    // node-sharded multi-GPU (fbs_set_peers): peer replicas of ONE registered local wire buffer; the epilogue only stores to
    // them when a level runs on exactly that buffer.  Each replica ends in a flag page (fbs_wires_alloc): flags[r] = last level
    // epoch rank r has completed (written by rank r with a system-scope release after its peer stores).
    int n_peers = 0, peer_rank = 0; u64 *peers[8] = {};
    u64 *peer_local = nullptr; size_t peer_bytes = 0;
    u64 *flags_local = nullptr, *peer_flags[8] = {};
    u64 epoch = 0; int *d_sync_err = nullptr;
    fq_tw *d_psi_rev = nullptr, *d_psi_inv_rev = nullptr;
    u64 *d_psi_pow = nullptr;                                  // psi^x - 1, x < 2N, packed residues (key-unrolled kernel)
    int unroll = 1, n_ggsw = 0; u32 mont2_ninv[2] = {0, 0};    // 2^64/N per prime
    u64 *d_gad_bsk = nullptr, *d_gad_ks = nullptr;
    u32 ninv[2] = {0, 0}, mont_ninv[2] = {0, 0};     // 1/N and 2^32/N per prime
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    std::vector<cudaEvent_t> ev_pool;          // 4 per level, for per-phase timing inside fbs_run / fbs_eval_bits
    // grow-only level scratch
    u8 *d_digits = nullptr; size_t cap_digits = 0;
    u64 *d_body = nullptr; size_t cap_body = 0;
    u16 *d_ms = nullptr; size_t cap_ms = 0;
    // staging for host-buffer calls
    u8 *d_io = nullptr; size_t cap_io = 0;
    u64 *d_wires = nullptr; size_t cap_wires = 0;
    u64 *d_mvacc = nullptr; size_t cap_mvacc = 0;              // multi-value bootstrap: accumulators between rotation and finishing
    u8 *d_clear_io = nullptr; size_t cap_clear_io = 0;         // fbs_clear_eval scratch
    int32_t *d_clear_lcv = nullptr; size_t cap_clear_lcv = 0;
};
struct fbs_prog {
    fbs_ctx *ctx = nullptr;
    int32_t p = 0, n_inputs = 0, n_lincombs = 0, n_boots = 0, n_levels = 0, n_slots = 0, n_outputs = 0;
    bool contiguous_levels = false;                       // slots are never recycled: node sub-ranges of a level may run alone
    std::vector<int32_t> lc_level_ptr, bs_level_ptr, bs_lc;
    int32_t n_groups = 0; std::vector<int32_t> grp_level_ptr, grp_first;     // multi-value bootstrap (n_groups > 0)
    int32_t *d_grp_first = nullptr;
    int max_lc_per_level = 0;
    int32_t *d_i32 = nullptr; u8 *d_tab = nullptr;        // one arena for all int32 arrays
    int32_t *d_lc_level_ptr, *d_bs_level_ptr, *d_lc_ptr, *d_lc_slot, *d_lc_coef, *d_lc_const, *d_bs_lc, *d_bs_slot,
        *d_bs_tab_ptr, *d_bs_mode, *d_in_slot, *d_out_ptr, *d_out_slot, *d_out_coef, *d_out_const;
};

template <class T> static int dev_alloc(T **p, size_t count)
{
    *p = nullptr;
    if (count == 0) count = 1;
    CK(cudaMalloc((void **)p, count * sizeof(T)));
    return FBS_OK;
}
template <class T> static int grow(T **p, size_t *cap, size_t need)
{
    if (need <= *cap) return FBS_OK;
    if (*p) CK(cudaFree(*p));
    *p = nullptr; *cap = 0;
    size_t want = need + need / 8;
    CK(cudaMalloc((void **)p, want * sizeof(T)));
    *cap = want;
    return FBS_OK;
}
// Host -> device copy that has LANDED when it returns.  A plain cudaMemcpy from pageable memory may return once the data sits in
// the staging buffer, before the DMA to the device completes; kernels on this library's NON-BLOCKING streams are not ordered
// after the legacy default stream, so they could read stale memory (seen: the last ciphertexts of a parity-tap batch).
static cudaError_t h2d_sync(void *dst, const void *src, size_t bytes)
{
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(0);
}
static u32 bitrev32(u32 x, int bits) { u32 r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

// ------------------------------------------------------------------------------------------------------
extern "C" int fbs_ctx_destroy(fbs_ctx *c);
static int ctx_create_impl(const fbs_params *params, int device, uint64_t seed, fbs_ctx **out, fbs_ctx **partial);
extern "C" int fbs_ctx_create(const fbs_params *params, int device, uint64_t seed, fbs_ctx **out)
{
    fbs_ctx *partial = nullptr;
    const int rc = ctx_create_impl(params, device, seed, out, &partial);
    if (rc != FBS_OK && partial) { const std::string keep = g_err; fbs_ctx_destroy(partial); g_err = keep; }   // no leak on failure
    return rc;
}
static int ctx_create_impl(const fbs_params *params, int device, uint64_t seed, fbs_ctx **out, fbs_ctx **partial)
{
    if (!params || !out) return fail(FBS_ERR_ARG, "fbs_ctx_create: null argument");
    const fbs_params &P = *params;
    int logN = 0; while ((1 << logN) < P.N) logN++;
    if ((1 << logN) != P.N) return fail(FBS_ERR_ARG, "N must be a power of two");
    if (P.ks_beta < 1 || P.ks_beta > 8 || P.ks_l < 1 || P.ks_l > 8 || P.ks_beta * P.ks_l > 40)
        return fail(FBS_ERR_ARG, "unsupported key-switch decomposition (need 1<=ks_beta<=8, 1<=ks_l<=8)");
    if (P.bsk_beta * P.bsk_l > 48 || P.bsk_beta < 2 || P.bsk_beta > 28) return fail(FBS_ERR_ARG, "unsupported blind-rotate decomposition");
    if (P.bsk_l == 1 && P.bsk_beta > 24) return fail(FBS_ERR_ARG, "one-level blind-rotate decomposition needs bsk_beta <= 24");
    if (P.n < 1 || P.n > 4095) return fail(FBS_ERR_ARG, "n out of range");
    const BRVariant *br = nullptr, *br1 = nullptr;
    const int unroll = (P.bsk_unroll == 2 || P.bsk_unroll == 3) ? P.bsk_unroll : 1;
    if (P.bsk_unroll < 0 || P.bsk_unroll > 3) return fail(FBS_ERR_ARG, "bsk_unroll must be 0, 1, 2 or 3");
    if (unroll > 1 && P.bsk_l != 1) return fail(FBS_ERR_ARG, "bsk_unroll > 1 needs bsk_l = 1");
    for (const BRVariant &v : g_br_variants) if (v.logN == logN && v.k == P.k && v.l == P.bsk_l && v.unr == unroll) {
        if (!br || v.pb > br->pb) br = &v;
        if (v.pb == 1) br1 = &v;
    }
    if (!br) return fail(FBS_ERR_ARG, "no blind-rotate kernel compiled for (N=" + std::to_string(P.N) + ", k=" + std::to_string(P.k) + ", l=" + std::to_string(P.bsk_l) + ")");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(FBS_ERR_ARG, "no such CUDA device");
    CK(cudaSetDevice(device));
    fbs_ctx *c = new fbs_ctx();
    *partial = c;
    c->P = P; c->device = device; c->seed = seed; c->logN = logN; c->br = br; c->br1 = br1 ? br1 : br;
    c->unroll = unroll; c->n_ggsw = unroll > 1 ? ((1 << unroll) - 1) * ((P.n + unroll - 1) / unroll) : P.n;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->br_smem = br->smem(P.n);
    if (c->br_smem > (size_t)prop.sharedMemPerBlockOptin)
        return fail(FBS_ERR_ARG, "blind-rotate kernel needs " + std::to_string(c->br_smem) + " B shared memory, device offers " + std::to_string(prop.sharedMemPerBlockOptin));
    // the attribute belongs to the FUNCTION, not to this context: contexts with different n share a kernel instantiation,
    // so it is raised to the device limit once instead of to this context's need (a later, smaller context would lower it)
    CK(br->prepare((size_t)prop.sharedMemPerBlockOptin));
    CK(cudaFuncSetAttribute(k_keyswitch_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KS_SMEM_BYTES));
    c->br1_smem = c->br1->smem(P.n);
    if (c->br1 != c->br) CK(c->br1->prepare((size_t)prop.sharedMemPerBlockOptin));
    for (const BRCVariant &v : g_brc_variants) if (v.logN == logN && v.k == P.k && v.unr == unroll && P.bsk_l == 1) {
        const size_t sm = v.smem(P.n);
        if (sm > (size_t)prop.sharedMemPerBlockOptin) continue;
        CK(v.prepare((size_t)prop.sharedMemPerBlockOptin));
        int mc = 0;
        if (v.max_clusters(sm, &mc) != cudaSuccess || mc < 1) { cudaGetLastError(); continue; }
        const int kind = v.ps ? v.occ : 0;
        c->brc[kind][v.logC] = &v; c->brc_smem[kind][v.logC] = sm; c->brc_max[kind][v.logC] = mc;
    }
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto &e : c->ev) CK(cudaEventCreate(&e));
    // twiddles psi^bitrev(i): 7 generates Z_P^*
    const int N = P.N;
    std::vector<fq_tw> pr(N), pir(N);
    const u32 primes[2] = {FQ_P1, FQ_P2};
    std::vector<u32> w[2], wi[2];
    for (int l = 0; l < 2; l++) {          // 3 is a quadratic non-residue mod both primes: 3^((p-1)/2N) has order exactly 2N
        const u32 p = primes[l];
        const u64 psi = pow_mod_host(3, (p - 1) / (2ULL * N), p), psi_inv = pow_mod_host(psi, p - 2, p);
        w[l].resize(N); wi[l].resize(N);
        for (int i = 0; i < N; i++) { u32 r = bitrev32((u32)i, logN); w[l][i] = (u32)pow_mod_host(psi, r, p); wi[l][i] = (u32)pow_mod_host(psi_inv, r, p); }
        c->ninv[l] = (u32)pow_mod_host((u64)N, p - 2, p);
        c->mont_ninv[l] = (u32)((u64)((1ULL << 32) % p) * c->ninv[l] % p);
        c->mont2_ninv[l] = (u32)((u64)((1ULL << 32) % p) * c->mont_ninv[l] % p);
    }
    {
        std::vector<u64> pp(2 * (size_t)N);
        const u64 ps1 = pow_mod_host(3, (FQ_P1 - 1) / (2ULL * N), FQ_P1), ps2 = pow_mod_host(3, (FQ_P2 - 1) / (2ULL * N), FQ_P2);
        u64 x1 = 1, x2 = 1;
        for (int x = 0; x < 2 * N; x++) { pp[x] = (x1 - 1) | ((x2 - 1) << 32); x1 = x1 * ps1 % FQ_P1; x2 = x2 * ps2 % FQ_P2; }   // psi^x - 1 (psi^x >= 1)
        CKR(dev_alloc(&c->d_psi_pow, 2 * (size_t)N));
        CK(h2d_sync(c->d_psi_pow, pp.data(), 16 * (size_t)N));
    }
    for (int i = 0; i < N; i++) {
        pr[i] = fq_tw{w[0][i], shoup32_host(w[0][i], FQ_P1), w[1][i], shoup32_host(w[1][i], FQ_P2)};
        pir[i] = fq_tw{wi[0][i], shoup32_host(wi[0][i], FQ_P1), wi[1][i], shoup32_host(wi[1][i], FQ_P2)};
    }
    CKR(dev_alloc(&c->d_psi_rev, N)); CKR(dev_alloc(&c->d_psi_inv_rev, N));
    CK(h2d_sync(c->d_psi_rev, pr.data(), sizeof(fq_tw) * N));
    CK(h2d_sync(c->d_psi_inv_rev, pir.data(), sizeof(fq_tw) * N));
    std::vector<u64> gb(8, 0), gk(8, 0);
    for (int j = 0; j < P.bsk_l; j++) gb[j] = fbs_gadget_host(P.bsk_beta, j);
    for (int j = 0; j < P.ks_l; j++) gk[j] = fbs_gadget_host(P.ks_beta, j);
    CKR(dev_alloc(&c->d_gad_bsk, 8)); CKR(dev_alloc(&c->d_gad_ks, 8));
    CK(h2d_sync(c->d_gad_bsk, gb.data(), 64));
    CK(h2d_sync(c->d_gad_ks, gk.data(), 64));
    *out = c;
    *partial = nullptr;
    return FBS_OK;
}

extern "C" int fbs_keygen(fbs_ctx *c)
{
    if (!c) return fail(FBS_ERR_ARG, "fbs_keygen: null ctx");
    CK(cudaSetDevice(c->device));
    const fbs_params &P = c->P; const int n = P.n, k = P.k, N = P.N, l = P.bsk_l, lk = P.ks_l, D = k * N;
    cudaStream_t st = c->stream;
    if (!c->d_s_lwe) { CKR(dev_alloc(&c->d_s_lwe, n)); CKR(dev_alloc(&c->d_s_big, D)); }
    k_gen_bits<<<(n + 255) / 256, 256, 0, st>>>(c->d_s_lwe, n, c->seed, DOM_SLWE);
    k_gen_bits<<<(D + 255) / 256, 256, 0, st>>>(c->d_s_big, D, c->seed, DOM_SGLWE);
    const size_t R = (size_t)D * lk;
    if (!c->d_ksk) { CKR(dev_alloc(&c->d_ksk, R * (n + 1))); CKR(dev_alloc(&c->d_colsum, n + 1)); }
    k_gen_ksk<<<(unsigned)R, 256, 0, st>>>(c->d_ksk, n, lk, c->seed, c->d_s_lwe, c->d_s_big, P.lwe_noise, c->d_gad_ks);
    k_ksk_colsum<<<(n + 1 + 127) / 128, 128, 0, st>>>(c->d_ksk, (int)R, n + 1, c->d_colsum);
    const int cols_pad = (n + 1 + 7) / 8 * 8;
    if (!c->d_kbt) CKR(dev_alloc(&c->d_kbt, (size_t)cols_pad * 8 * R));
    k_ksk_bytes_t<<<dim3((unsigned)((R + 255) / 256), (unsigned)cols_pad), 256, 0, st>>>(c->d_ksk, c->d_kbt, (int)R, n + 1, cols_pad);
    const int rows = (k + 1) * l;
    const size_t total = (size_t)c->n_ggsw * rows * (k + 1) * N;
    if (!c->d_bsk) { CKR(dev_alloc(&c->d_bsk, total)); CKR(dev_alloc(&c->d_bsk_coef, total)); }
    k_gen_bsk_fill<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(c->d_bsk_coef, k, N, c->seed, P.glwe_noise, total);
    k_bsk_body<<<c->n_ggsw * rows, 256, (size_t)N * 9, st>>>(c->d_bsk_coef, k, N, l, c->d_s_lwe, c->d_s_big, c->d_gad_bsk, c->unroll, n);
    ntt_launch_fn nf = ntt_for(c->logN);
    if (!nf) return fail(FBS_ERR_ARG, "no NTT kernel for this N");
    if (c->unroll > 1) nf(c->d_bsk_coef, c->d_bsk, 3, c->d_psi_rev, c->d_psi_inv_rev, c->mont2_ninv[0], c->mont2_ninv[1], (long long)(total / N), st, (k + 1) | (((1 << c->unroll) - 1) << 8));
    else nf(c->d_bsk_coef, c->d_bsk, 2, c->d_psi_rev, c->d_psi_inv_rev, c->mont_ninv[0], c->mont_ninv[1], (long long)(total / N), st, k + 1);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    // the coefficient-domain copy is only a parity tap: keep it for toy sizes, drop it for real key sizes
    if (total * 8 > (64u << 20)) { CK(cudaFree(c->d_bsk_coef)); c->d_bsk_coef = nullptr; }
    c->have_keys = true;
    return FBS_OK;
}

extern "C" int fbs_ctx_destroy(fbs_ctx *c)
{
    if (!c) return FBS_OK;
    cudaSetDevice(c->device);
    void *ptrs[] = {c->d_kbt, c->d_s_lwe, c->d_s_big, c->d_ksk, c->d_colsum, c->d_bsk, c->d_bsk_coef, c->d_psi_rev, c->d_psi_inv_rev, c->d_psi_pow,
                    c->d_gad_bsk, c->d_gad_ks, c->d_digits, c->d_body, c->d_ms, c->d_io, c->d_wires, c->d_clear_io, c->d_clear_lcv, c->d_sync_err, c->d_mvacc};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->ev_pool) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return FBS_OK;
}
extern "C" int fbs_ctx_set_cluster(fbs_ctx *c, int32_t mode)
{
    const bool ok = mode == 0 || mode == 1 || mode == 2 || mode == 4 || mode == 8 || mode == 12 || mode == 14 || mode == 18 || mode == 24;
    if (!c || !ok) return fail(FBS_ERR_ARG, "fbs_ctx_set_cluster: mode must be 0 (auto), 1 (off), 2, 4, 8, or 12, 14, 18 (prime-split), 24 (prime-split, 2 CTAs per SM)");
    if (mode > 1) {
        const int sz = mode % 10, lc = sz == 2 ? 1 : sz == 4 ? 2 : 3;
        if (!c->brc[mode / 10][lc]) return fail(FBS_ERR_ARG, "fbs_ctx_set_cluster: no cluster kernel of that kind for this parameter set");
    }
    c->cluster_mode = mode;
    return FBS_OK;
}
extern "C" int fbs_ctx_info(const fbs_ctx *c, int32_t *sm_count, int64_t *bsk_bytes, int64_t *ksk_bytes, int32_t *br_smem_bytes)
{
    if (!c) return fail(FBS_ERR_ARG, "null ctx");
    const fbs_params &P = c->P;
    if (sm_count) *sm_count = c->sm_count;
    if (bsk_bytes) *bsk_bytes = (int64_t)c->n_ggsw * (P.k + 1) * P.bsk_l * (P.k + 1) * P.N * 8;
    if (ksk_bytes) *ksk_bytes = (int64_t)P.k * P.N * P.ks_l * (P.n + 1) * 8;
    if (br_smem_bytes) *br_smem_bytes = (int32_t)c->br_smem;
    return FBS_OK;
}

// ------------------------------------------------------------------------------------------------------
// program
// ------------------------------------------------------------------------------------------------------
// CSR pointer array: starts at 0, monotone, `count + 1` entries; returns the final value or -1
static int csr_total(const int32_t *ptr, int count)
{
    if (count == 0) return 0;
    if (!ptr || ptr[0] != 0) return -1;
    for (int i = 0; i < count; i++) if (ptr[i + 1] < ptr[i]) return -1;
    return ptr[count];
}
static int prog_load_impl(fbs_ctx *c, const fbs_prog_desc *d, fbs_prog *g);
extern "C" int fbs_prog_load(fbs_ctx *c, const fbs_prog_desc *d, fbs_prog **out)
{
    if (!c || !d || !out) return fail(FBS_ERR_ARG, "fbs_prog_load: null argument");
    *out = nullptr;
    fbs_prog *g = new fbs_prog();
    const int rc = prog_load_impl(c, d, g);
    if (rc != FBS_OK) { const std::string keep = g_err; fbs_prog_free(g); g_err = keep; return rc; }   // no leak on any error path
    *out = g;
    return FBS_OK;
}
static int prog_load_impl(fbs_ctx *c, const fbs_prog_desc *d, fbs_prog *g)
{
    g->ctx = c;
    if (d->p < 2 || d->p > 128) return fail(FBS_ERR_ARG, "p out of range");
    if (d->n_inputs < 0 || d->n_lincombs < 0 || d->n_boots < 0 || d->n_levels < 0 || d->n_slots < 0 || d->n_outputs < 0)
        return fail(FBS_ERR_ARG, "fbs_prog_load: negative count");
    if (!d->lc_level_ptr || !d->bs_level_ptr) return fail(FBS_ERR_ARG, "fbs_prog_load: null level pointers");
    // every CSR / level pointer array: monotone from 0, final value = the count it partitions
    const int nnz = csr_total(d->lc_ptr, d->n_lincombs), onz = csr_total(d->out_ptr, d->n_outputs), tabn = csr_total(d->bs_tab_ptr, d->n_boots);
    if (nnz < 0 || onz < 0 || tabn < 0) return fail(FBS_ERR_ARG, "fbs_prog_load: lc_ptr / out_ptr / bs_tab_ptr must start at 0 and be monotone");
    if (csr_total(d->lc_level_ptr, d->n_levels) != d->n_lincombs || csr_total(d->bs_level_ptr, d->n_levels) != d->n_boots)
        return fail(FBS_ERR_ARG, "fbs_prog_load: level pointers must be monotone from 0 and end at n_lincombs / n_boots");
    if ((nnz && (!d->lc_slot || !d->lc_coef)) || (d->n_lincombs && !d->lc_const) || (onz && (!d->out_slot || !d->out_coef)) || (d->n_outputs && !d->out_const) ||
        (d->n_boots && (!d->bs_lc || !d->bs_slot || !d->bs_mode || !d->bs_tab)) || (d->n_inputs && !d->in_slot))
        return fail(FBS_ERR_ARG, "fbs_prog_load: null array");
    for (int i = 0; i < nnz; i++) if (d->lc_slot[i] < 0 || d->lc_slot[i] >= d->n_slots) return fail(FBS_ERR_ARG, "lincomb operand slot out of range");
    for (int i = 0; i < onz; i++) if (d->out_slot[i] < 0 || d->out_slot[i] >= d->n_slots) return fail(FBS_ERR_ARG, "output operand slot out of range");
    for (int i = 0; i < d->n_inputs; i++) if (d->in_slot[i] < 0 || d->in_slot[i] >= d->n_slots) return fail(FBS_ERR_ARG, "input slot out of range");
    for (int q = 0; q < d->n_boots; q++) {
        const int L = d->bs_tab_ptr[q + 1] - d->bs_tab_ptr[q];
        if (L < 1 || L > 2 * d->p || L > 64) return fail(FBS_ERR_ARG, "bootstrap table longer than 2p (or 64) entries");
        if (d->bs_slot[q] < 0 || d->bs_slot[q] >= d->n_slots) return fail(FBS_ERR_ARG, "bootstrap slot out of range");
        if (d->bs_lc[q] < 0 || d->bs_lc[q] >= d->n_lincombs) return fail(FBS_ERR_ARG, "bootstrap lincomb index out of range");
        if (d->bs_mode[q] < 0 || d->bs_mode[q] >= 2 * d->p) return fail(FBS_ERR_ARG, "bootstrap table mode out of range");
        for (int t = d->bs_tab_ptr[q]; t < d->bs_tab_ptr[q + 1]; t++) if (d->bs_tab[t] >= 2 * d->p) return fail(FBS_ERR_ARG, "bootstrap table entry >= 2p");
    }
    CK(cudaSetDevice(c->device));
    g->p = d->p; g->n_inputs = d->n_inputs; g->n_lincombs = d->n_lincombs; g->n_boots = d->n_boots;
    g->n_levels = d->n_levels; g->n_slots = d->n_slots; g->n_outputs = d->n_outputs; g->contiguous_levels = d->contiguous_levels != 0;
    g->lc_level_ptr.assign(d->lc_level_ptr, d->lc_level_ptr + d->n_levels + 1);
    g->bs_level_ptr.assign(d->bs_level_ptr, d->bs_level_ptr + d->n_levels + 1);
    g->bs_lc.assign(d->bs_lc, d->bs_lc + d->n_boots);
    for (int lv = 0; lv < d->n_levels; lv++) {
        g->max_lc_per_level = std::max(g->max_lc_per_level, g->lc_level_ptr[lv + 1] - g->lc_level_ptr[lv]);
        for (int q = g->bs_level_ptr[lv]; q < g->bs_level_ptr[lv + 1]; q++) {
            if (g->bs_lc[q] < g->lc_level_ptr[lv] || g->bs_lc[q] >= g->lc_level_ptr[lv + 1]) return fail(FBS_ERR_ARG, "bootstrap uses a lincomb of another level");
            if (q > g->bs_level_ptr[lv] && g->bs_lc[q] < g->bs_lc[q - 1]) return fail(FBS_ERR_ARG, "bootstraps of a level must be sorted by lincomb index");
        }
    }
    struct Piece { const int32_t *src; size_t n; int32_t **dst; };
    Piece pieces[] = {
        {d->lc_level_ptr, (size_t)d->n_levels + 1, &g->d_lc_level_ptr}, {d->bs_level_ptr, (size_t)d->n_levels + 1, &g->d_bs_level_ptr},
        {d->lc_ptr, (size_t)d->n_lincombs + 1, &g->d_lc_ptr}, {d->lc_slot, (size_t)nnz, &g->d_lc_slot}, {d->lc_coef, (size_t)nnz, &g->d_lc_coef},
        {d->lc_const, (size_t)d->n_lincombs, &g->d_lc_const}, {d->bs_lc, (size_t)d->n_boots, &g->d_bs_lc}, {d->bs_slot, (size_t)d->n_boots, &g->d_bs_slot},
        {d->bs_tab_ptr, (size_t)d->n_boots + 1, &g->d_bs_tab_ptr}, {d->bs_mode, (size_t)d->n_boots, &g->d_bs_mode}, {d->in_slot, (size_t)d->n_inputs, &g->d_in_slot},
        {d->out_ptr, (size_t)d->n_outputs + 1, &g->d_out_ptr}, {d->out_slot, (size_t)onz, &g->d_out_slot}, {d->out_coef, (size_t)onz, &g->d_out_coef},
        {d->out_const, (size_t)d->n_outputs, &g->d_out_const}};
    size_t tot = 0;
    for (auto &pc : pieces) tot += pc.n + 4;
    std::vector<int32_t> host(tot, 0);
    CKR(dev_alloc(&g->d_i32, tot));
    size_t off = 0;
    for (auto &pc : pieces) {
        if (pc.n && pc.src) memcpy(host.data() + off, pc.src, pc.n * 4);
        *pc.dst = g->d_i32 + off;
        off += (pc.n + 3) & ~(size_t)3;
    }
    CK(h2d_sync(g->d_i32, host.data(), tot * 4));
    CKR(dev_alloc(&g->d_tab, (size_t)tabn + 16));
    if (tabn) CK(h2d_sync(g->d_tab, d->bs_tab, tabn));
    if (d->n_groups > 0) {
        // multi-value: groups partition every level's bootstraps into runs that share one lincomb
        if (!d->grp_level_ptr || !d->grp_first || csr_total(d->grp_level_ptr, d->n_levels) != d->n_groups || csr_total(d->grp_first, d->n_groups) != d->n_boots)
            return fail(FBS_ERR_ARG, "fbs_prog_load: group pointers must be monotone from 0 and end at n_groups / n_boots");
        g->n_groups = d->n_groups;
        g->grp_level_ptr.assign(d->grp_level_ptr, d->grp_level_ptr + d->n_levels + 1);
        g->grp_first.assign(d->grp_first, d->grp_first + d->n_groups + 1);
        for (int lv = 0; lv < d->n_levels; lv++) {
            if (g->grp_first[g->grp_level_ptr[lv]] != g->bs_level_ptr[lv] && g->grp_level_ptr[lv] < g->grp_level_ptr[lv + 1]) return fail(FBS_ERR_ARG, "fbs_prog_load: groups do not start at the level's first bootstrap");
            if (g->grp_level_ptr[lv] == g->grp_level_ptr[lv + 1] && g->bs_level_ptr[lv] != g->bs_level_ptr[lv + 1]) return fail(FBS_ERR_ARG, "fbs_prog_load: level with bootstraps but no group");
            for (int gi = g->grp_level_ptr[lv]; gi < g->grp_level_ptr[lv + 1]; gi++) {
                if (g->grp_first[gi + 1] <= g->grp_first[gi] || g->grp_first[gi + 1] > g->bs_level_ptr[lv + 1]) return fail(FBS_ERR_ARG, "fbs_prog_load: empty group or group across levels");
                for (int q = g->grp_first[gi]; q < g->grp_first[gi + 1]; q++) if (g->bs_lc[q] != g->bs_lc[g->grp_first[gi]]) return fail(FBS_ERR_ARG, "fbs_prog_load: bootstraps of a group must share their lincomb");
            }
        }
        CKR(dev_alloc(&g->d_grp_first, (size_t)d->n_groups + 1));
        CK(h2d_sync(g->d_grp_first, g->grp_first.data(), ((size_t)d->n_groups + 1) * 4));
    }
    return FBS_OK;
}
extern "C" int fbs_prog_free(fbs_prog *g)
{
    if (!g) return FBS_OK;
    cudaSetDevice(g->ctx->device);
    if (g->d_i32) cudaFree(g->d_i32);
    if (g->d_tab) cudaFree(g->d_tab);
    if (g->d_grp_first) cudaFree(g->d_grp_first);
    delete g;
    return FBS_OK;
}

// ------------------------------------------------------------------------------------------------------
// level scheduler
// ------------------------------------------------------------------------------------------------------
static inline size_t ct_words(const fbs_ctx *c) { return (size_t)c->P.k * c->P.N + 1; }

extern "C" int fbs_wires_bytes(const fbs_ctx *c, const fbs_prog *g, int64_t B, size_t *bytes)
{
    if (!c || !g || !bytes || B < 1) return fail(FBS_ERR_ARG, "fbs_wires_bytes: bad argument");
    *bytes = (size_t)g->n_slots * (size_t)B * ct_words(c) * 8;
    return FBS_OK;
}

extern "C" int fbs_encrypt_inputs(fbs_ctx *c, fbs_prog *g, const uint8_t *in_dev, int64_t B, int64_t inst_offset, int64_t B_total,
                                  uint64_t enc_seed, uint64_t *wires_dev, void *stream)
{
    if (!c || !g || !in_dev || !wires_dev || B < 1) return fail(FBS_ERR_ARG, "fbs_encrypt_inputs: bad argument");
    if (!c->have_keys) return fail(FBS_ERR_STATE, "fbs_encrypt_inputs before fbs_keygen");
    CK(cudaSetDevice(c->device));
    if (g->n_inputs == 0) return FBS_OK;
    EncArgs a{};
    a.msgs = in_dev; a.in_slot = g->d_in_slot; a.out = wires_dev; a.s_big = c->d_s_big;
    a.B = B; a.inst_offset = inst_offset; a.B_total = B_total; a.D = c->P.k * c->P.N; a.p = g->p;
    a.enc_seed = enc_seed; a.noise_scale = c->P.glwe_noise;
    k_encrypt<<<(unsigned)((long long)g->n_inputs * B), 256, 0, (cudaStream_t)stream>>>(a);
    CK(cudaGetLastError());
    return FBS_OK;
}

template <int LK> static void launch_lc(const LCArgs &a, long long tiles, cudaStream_t st, int sm_count)
{
    // column chunks so that even a level with few ciphertexts puts about two CTAs on every SM
    int ny = (int)std::min<long long>(std::max<long long>(1, (2 * sm_count + tiles - 1) / tiles), (a.D + 1 + 255) / 256);
    k_lincomb_decomp<LK><<<dim3((unsigned)tiles, (unsigned)ny), 256, 0, st>>>(a);
}

static int run_level_impl(fbs_ctx *c, fbs_prog *g, int level, int nb, int ne, int64_t B, u64 *wires, cudaStream_t st,
                          fbs_run_stats *stats, u64 *tap_ks, u64 *tap_acc, bool timed, cudaEvent_t *evq = nullptr)
{
    cudaEvent_t *E = evq ? evq : c->ev;
    const bool rec = timed || evq;
    const int b0 = g->bs_level_ptr[level], b1 = g->bs_level_ptr[level + 1];
    const bool multi = g->n_groups > 0;
    // the unit of a node range: bootstraps -- or, for a multi-value program, GROUPS (all tables of a group share one rotation)
    const int grp0 = multi ? g->grp_level_ptr[level] : 0, units = multi ? g->grp_level_ptr[level + 1] - grp0 : b1 - b0;
    const bool whole = (nb < 0 && ne < 0) || (nb == 0 && ne == units);
    if (nb < 0 && ne < 0) { nb = 0; ne = units; }
    if (nb < 0 || ne > units || nb > ne) return fail(FBS_ERR_ARG, "fbs_run_level: node range outside the level");
    if (!whole && !g->contiguous_levels)
        return fail(FBS_ERR_ARG, "fbs_run_level: a node sub-range needs a program levelised with contiguous_levels (slots are recycled otherwise)");
    // Node-sharded level on the registered, peer-mapped wire buffer: wait (on the device) until every rank has finished the
    // previous level, and publish this rank's completion after the blind rotation's peer stores -- no host sync between levels.
    const bool fused = c->n_peers > 0 && wires == c->peer_local, handoff = fused && c->peer_rank >= 0;
    if (fused && (size_t)g->n_slots * (size_t)B * ct_words(c) * 8 > c->peer_bytes) return fail(FBS_ERR_ARG, "fbs_run_level: program does not fit the registered peer wire buffer");
    if (handoff) {
        if (c->epoch > 0) { k_level_wait<<<1, 32, 0, st>>>(c->flags_local, c->n_peers + 1, c->peer_rank, c->epoch, c->d_sync_err); CK(cudaGetLastError()); }
    }
    struct Signal {            // runs on every exit path of a fused level, including an empty node range
        fbs_ctx *c; cudaStream_t st; bool on;
        ~Signal() { if (on) { LevelPeers lp{}; lp.n = c->n_peers; for (int i = 0; i < c->n_peers; i++) lp.flags[i] = c->peer_flags[i];
                              c->epoch++; k_level_signal<<<1, 32, 0, st>>>(lp, c->peer_rank, c->epoch); } }
    } signal{c, st, handoff};
    if (nb == ne) return FBS_OK;
    const fbs_params &P = c->P;
    const int D = P.k * P.N, n = P.n;
    const int node0 = multi ? g->grp_first[grp0 + nb] : b0 + nb, node1 = multi ? g->grp_first[grp0 + ne] : b0 + ne;
    const int lc0 = g->bs_lc[node0], lc1 = g->bs_lc[node1 - 1] + 1;     // bootstraps are sorted by lincomb
    const long long M = (long long)(lc1 - lc0) * B, tiles = (M + 15) / 16, mtiles = (M + KS_BM - 1) / KS_BM;
    const size_t R = (size_t)D * P.ks_l;
    if (R % KS_BK) return fail(FBS_ERR_ARG, "k*N*ks_l must be a multiple of 128");
    CKR(grow(&c->d_digits, &c->cap_digits, (size_t)mtiles * KS_BM * R));
    CKR(grow(&c->d_body, &c->cap_body, (size_t)mtiles * KS_BM));
    CKR(grow(&c->d_ms, &c->cap_ms, (size_t)mtiles * KS_BM * (n + 1)));
    if (rec) CK(cudaEventRecord(E[0], st));
    LCArgs la{};
    la.wires = wires; la.lc_ptr = g->d_lc_ptr; la.lc_slot = g->d_lc_slot; la.lc_coef = g->d_lc_coef; la.lc_const = g->d_lc_const;
    la.digits = c->d_digits; la.body = c->d_body; la.B = B; la.M = M; la.lc_begin = lc0; la.D = D; la.p = g->p; la.ks_beta = P.ks_beta;
    switch (P.ks_l) {
    case 1: launch_lc<1>(la, tiles, st, c->sm_count); break; case 2: launch_lc<2>(la, tiles, st, c->sm_count); break;
    case 3: launch_lc<3>(la, tiles, st, c->sm_count); break; case 4: launch_lc<4>(la, tiles, st, c->sm_count); break;
    case 5: launch_lc<5>(la, tiles, st, c->sm_count); break; case 6: launch_lc<6>(la, tiles, st, c->sm_count); break;
    case 7: launch_lc<7>(la, tiles, st, c->sm_count); break; default: launch_lc<8>(la, tiles, st, c->sm_count); break;
    }
    CK(cudaGetLastError());
    if (rec) CK(cudaEventRecord(E[1], st));
    KSArgs ka{};
    ka.digits = c->d_digits; ka.body = c->d_body; ka.kbt = c->d_kbt; ka.colsum = c->d_colsum; ka.ms = c->d_ms; ka.tap_ks = tap_ks;
    ka.M = M; ka.R = (int)R; ka.n = n; ka.ks_beta = P.ks_beta; ka.log2_2N = c->logN + 1;
    const unsigned ngrid = (unsigned)(((n + 1 + 7) / 8 * 8 * 8 + KS_BN - 1) / KS_BN);
    k_keyswitch_mma<<<dim3((unsigned)mtiles, ngrid), 256, KS_SMEM_BYTES, st>>>(ka);
    CK(cudaGetLastError());
    if (rec) CK(cudaEventRecord(E[2], st));
    BRArgs ba{};
    ba.ms = c->d_ms; ba.bsk = c->d_bsk; ba.psi_rev = c->d_psi_rev; ba.psi_inv_rev = c->d_psi_inv_rev; ba.psi_pow = c->d_psi_pow;
    ba.bs_lc = g->d_bs_lc; ba.bs_slot = g->d_bs_slot; ba.bs_tab_ptr = g->d_bs_tab_ptr; ba.bs_mode = g->d_bs_mode; ba.bs_tab = g->d_tab;
    // multi-value: one job per (group, instance); the kernels index groups and leave the accumulators in d_mvacc for k_multi_extract
    const long long jobs = multi ? (long long)(ne - nb) * B : (long long)(node1 - node0) * B;
    if (multi && !tap_acc) { CKR(grow(&c->d_mvacc, &c->cap_mvacc, (size_t)jobs * (P.k + 1) * P.N)); tap_acc = c->d_mvacc; }
    ba.wires = wires; ba.tap_acc = tap_acc; ba.B = B; ba.jobs = jobs; ba.node_begin = multi ? grp0 + nb : node0;
    ba.grp_first = multi ? g->d_grp_first : nullptr;
    ba.n_peers = fused ? c->n_peers : 0;      // peers are bound to the registered buffer only (never to c->d_wires or a tap buffer)
    for (int pr = 0; pr < ba.n_peers; pr++) ba.peer_wires[pr] = c->peers[pr];
    ba.lc_begin = lc0; ba.n = n; ba.p = g->p; ba.beta = P.bsk_beta;
    const long long wave = (long long)c->sm_count * c->br->pb, tail = jobs % wave;
    int n_br_launches = 1;
    // Fewer jobs than SMs: split each bootstrap over a cluster of C CTAs, cutting the latency of the launch instead of idling SMs.
    // Measured, set A3, per bootstrap (profiles/r2_latency_A3_1gpu_v3.jsonl): one CTA 3.0 ms; packed clusters C = 2 / 4 / 8: 1.87 /
    // 1.34 / 1.38 ms; prime-split clusters C = 2 / 4 / 8: 2.02 / 1.20 / 0.93 ms (1.14 ms with 33 of them, two CTAs per SM).
    // Preference: the fastest kind whose clusters are all co-resident.  Results are bit-identical to the one-CTA kernels.
    auto pick_cluster = [&](long long nj) -> int {              // returns log2 C + 8 * kind (0 packed, 1 prime-split, 2 prime-split x2), 0 = none
        if (c->cluster_mode == 0) {
            static const int pref[][2] = {{1, 3}, {1, 2}, {0, 2}, {0, 3}, {0, 1}, {1, 1}};     // {kind, log2 C}; kind 2 (two CTAs per SM) measured slower than packed C = 2
            for (auto &pc : pref) if (c->brc[pc[0]][pc[1]] && nj <= c->brc_max[pc[0]][pc[1]]) return pc[1] + 8 * pc[0];
            return 0;
        }
        if (c->cluster_mode > 1) {
            const int kind = c->cluster_mode / 10, sz = c->cluster_mode % 10;
            for (int lc = 1; lc <= 3; lc++) if ((1 << lc) == sz && c->brc[kind][lc]) return lc + 8 * kind;
        }
        return 0;
    };
    auto launch_cluster = [&](int code, const BRArgs &args) { const int kind = code >> 3, lc = code & 7; return c->brc[kind][lc]->launch(args, jobs, c->brc_smem[kind][lc], st); };
    const int logC = pick_cluster(jobs);
    if (logC) CK(launch_cluster(logC, ba));
    else if (c->br1 != c->br && tail > 0 && tail <= c->sm_count) {
        // Full waves run the widest variant (pb bootstraps per CTA, one CTA per SM).  A last partial wave of at most one job per
        // SM goes to a narrower kernel instead of half-idle paired CTAs: cluster-split if its clusters fit the chip, else one
        // bootstrap per CTA (it starts as soon as SMs drain from the first launch).
        if (jobs > tail) { BRArgs bw = ba; CK(c->br->launch(bw, jobs - tail, c->br_smem, st)); n_br_launches = 2; }
        ba.job_begin = jobs - tail;
        const int tlc = c->cluster_mode == 0 ? pick_cluster(tail) : 0;
        if (tlc) CK(launch_cluster(tlc, ba));
        else CK(c->br1->launch(ba, jobs, c->br1_smem, st));
    } else CK(c->br->launch(ba, jobs, c->br_smem, st));
    if (multi) {
        MVArgs ma{};
        ma.acc = tap_acc; ma.grp_first = g->d_grp_first; ma.bs_slot = g->d_bs_slot; ma.bs_tab_ptr = g->d_bs_tab_ptr; ma.bs_mode = g->d_bs_mode; ma.bs_tab = g->d_tab;
        ma.wires = wires; ma.B = B; ma.grp_begin = grp0 + nb; ma.N = P.N; ma.K = P.k; ma.p = g->p;
        ma.n_peers = ba.n_peers; for (int pr = 0; pr < ba.n_peers; pr++) ma.peer_wires[pr] = ba.peer_wires[pr];
        k_multi_extract<<<(unsigned)jobs, 256, 0, st>>>(ma);
        CK(cudaGetLastError());
        n_br_launches++;
    }
    if (rec) CK(cudaEventRecord(E[3], st));
    if (stats) { stats->n_pbs += jobs; stats->n_launches += 2 + n_br_launches; }
    if (timed && stats) {
        CK(cudaEventSynchronize(c->ev[3]));
        float t;
        CK(cudaEventElapsedTime(&t, c->ev[0], c->ev[1])); stats->ms_lincomb += t;
        CK(cudaEventElapsedTime(&t, c->ev[1], c->ev[2])); stats->ms_keyswitch += t;
        CK(cudaEventElapsedTime(&t, c->ev[2], c->ev[3])); stats->ms_blind_rotate += t;
    }
    return FBS_OK;
}

extern "C" int fbs_run_level(fbs_ctx *c, fbs_prog *g, int32_t level, int32_t node_begin, int32_t node_end, int64_t B,
                             uint64_t *wires_dev, void *stream, fbs_run_stats *stats)
{
    if (!c || !g || !wires_dev || B < 1) return fail(FBS_ERR_ARG, "fbs_run_level: bad argument");
    if (!c->have_keys) return fail(FBS_ERR_STATE, "fbs_run_level before fbs_keygen");
    if (level < 0 || level >= g->n_levels) return fail(FBS_ERR_ARG, "fbs_run_level: no such level");
    CK(cudaSetDevice(c->device));
    return run_level_impl(c, g, level, node_begin, node_end, B, wires_dev, (cudaStream_t)stream, stats, nullptr, nullptr, stats != nullptr);
}

static int ensure_pool(fbs_ctx *c, size_t need)
{
    while (c->ev_pool.size() < need) { cudaEvent_t e; CK(cudaEventCreate(&e)); c->ev_pool.push_back(e); }
    return FBS_OK;
}
static int collect_phase_times(fbs_ctx *c, int n_levels, fbs_run_stats *stats)
{
    for (int lv = 0; lv < n_levels; lv++) {
        cudaEvent_t *E = &c->ev_pool[4 * (size_t)lv];
        float t;
        if (cudaEventQuery(E[3]) != cudaSuccess) CK(cudaEventSynchronize(E[3]));
        CK(cudaEventElapsedTime(&t, E[0], E[1])); stats->ms_lincomb += t;
        CK(cudaEventElapsedTime(&t, E[1], E[2])); stats->ms_keyswitch += t;
        CK(cudaEventElapsedTime(&t, E[2], E[3])); stats->ms_blind_rotate += t;
    }
    return FBS_OK;
}

extern "C" int fbs_run(fbs_ctx *c, fbs_prog *g, int64_t B, uint64_t *wires_dev, void *stream, fbs_run_stats *stats)
{
    if (!c || !g || !wires_dev || B < 1) return fail(FBS_ERR_ARG, "fbs_run: bad argument");
    if (!c->have_keys) return fail(FBS_ERR_STATE, "fbs_run before fbs_keygen");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (stats) { CKR(ensure_pool(c, 4 * (size_t)g->n_levels)); CK(cudaEventRecord(c->ev[4], st)); }
    for (int lv = 0; lv < g->n_levels; lv++)
        CKR(run_level_impl(c, g, lv, -1, -1, B, wires_dev, st, stats, nullptr, nullptr, false, stats ? &c->ev_pool[4 * (size_t)lv] : nullptr));
    if (stats) {
        CK(cudaEventRecord(c->ev[5], st));
        CK(cudaEventSynchronize(c->ev[5]));
        float t; CK(cudaEventElapsedTime(&t, c->ev[4], c->ev[5])); stats->ms_total += t;
        CKR(collect_phase_times(c, g->n_levels, stats));
    }
    return FBS_OK;
}

extern "C" int fbs_decrypt_outputs(fbs_ctx *c, fbs_prog *g, int64_t B, const uint64_t *wires_dev, uint8_t *out_dev, void *stream)
{
    if (!c || !g || !wires_dev || !out_dev || B < 1) return fail(FBS_ERR_ARG, "fbs_decrypt_outputs: bad argument");
    if (!c->have_keys) return fail(FBS_ERR_STATE, "fbs_decrypt_outputs before fbs_keygen");
    CK(cudaSetDevice(c->device));
    if (g->n_outputs == 0) return FBS_OK;
    OutArgs a{};
    a.wires = wires_dev; a.s_big = c->d_s_big; a.out_ptr = g->d_out_ptr; a.out_slot = g->d_out_slot; a.out_coef = g->d_out_coef;
    a.out_const = g->d_out_const; a.out8 = out_dev; a.B = B; a.D = c->P.k * c->P.N; a.p = g->p;
    k_decrypt<<<(unsigned)((long long)g->n_outputs * B), 256, 0, (cudaStream_t)stream>>>(a);
    CK(cudaGetLastError());
    return FBS_OK;
}

// ------------------------------------------------------------------------------------------------------
// host-buffer evaluation: the drop-in for LutExecEnv.eval
// ------------------------------------------------------------------------------------------------------
extern "C" int fbs_eval_bits(fbs_ctx *c, fbs_prog *g, const uint8_t *in, int64_t B, int64_t inst_offset, int64_t B_total,
                             uint64_t enc_seed, size_t max_wire_bytes, uint8_t *out, fbs_run_stats *stats)
{
    if (!c || !g || (!in && g->n_inputs) || (!out && g->n_outputs) || B < 1) return fail(FBS_ERR_ARG, "fbs_eval_bits: bad argument");
    if (!c->have_keys) return fail(FBS_ERR_STATE, "fbs_eval_bits before fbs_keygen");
    CK(cudaSetDevice(c->device));
    if (B_total < B) B_total = B;
    const fbs_params &P = c->P;
    const size_t CT = ct_words(c) * 8;
    const size_t R = (size_t)P.k * P.N * P.ks_l;
    int max_grp = 0;                                            // multi-value: one accumulator (32 KB at N = 2048) per group and instance
    for (int lv = 0; lv < g->n_levels && g->n_groups > 0; lv++) max_grp = std::max(max_grp, g->grp_level_ptr[lv + 1] - g->grp_level_ptr[lv]);
    const size_t per_inst = (size_t)g->n_slots * CT + (size_t)std::max(1, g->max_lc_per_level) * (R + 8 + 2 * (size_t)(P.n + 1)) +
                            (size_t)max_grp * (P.k + 1) * P.N * 8;
    if (max_wire_bytes == 0) {
        size_t fr = 0, tot = 0;
        CK(cudaMemGetInfo(&fr, &tot));
        max_wire_bytes = std::min<size_t>((size_t)48 << 30, (fr + c->cap_wires * 8 + c->cap_digits) / 2);
    }
    int64_t Bc = (int64_t)std::max<size_t>(1, max_wire_bytes / per_inst);
    if (Bc >= 16) Bc &= ~(int64_t)15;
    Bc = std::min<int64_t>(Bc, B);
    cudaStream_t st = c->stream;
    CKR(grow(&c->d_wires, &c->cap_wires, (size_t)g->n_slots * Bc * ct_words(c)));
    CKR(grow(&c->d_io, &c->cap_io, (size_t)(g->n_inputs + g->n_outputs + 1) * Bc));
    u8 *d_in = c->d_io, *d_out = c->d_io + (size_t)g->n_inputs * Bc;
    if (stats) CK(cudaEventRecord(c->ev[6], st));
    for (int64_t off = 0; off < B; off += Bc) {
        const int64_t bc = std::min<int64_t>(Bc, B - off);
        if (g->n_inputs) {
            CK(cudaMemcpy2DAsync(d_in, (size_t)bc, in + off, (size_t)B, (size_t)bc, (size_t)g->n_inputs, cudaMemcpyHostToDevice, st));
            if (stats) CK(cudaEventRecord(c->ev[0], st));
            CKR(fbs_encrypt_inputs(c, g, d_in, bc, inst_offset + off, B_total, enc_seed, c->d_wires, st));
            if (stats) { CK(cudaEventRecord(c->ev[1], st)); stats->n_launches += 1; }
        }
        if (stats) CKR(ensure_pool(c, 4 * (size_t)g->n_levels));
        for (int lv = 0; lv < g->n_levels; lv++)
            CKR(run_level_impl(c, g, lv, -1, -1, bc, c->d_wires, st, stats, nullptr, nullptr, false, stats ? &c->ev_pool[4 * (size_t)lv] : nullptr));
        if (g->n_outputs) {
            if (stats) CK(cudaEventRecord(c->ev[2], st));
            CKR(fbs_decrypt_outputs(c, g, bc, c->d_wires, d_out, st));
            if (stats) { CK(cudaEventRecord(c->ev[3], st)); stats->n_launches += 1; }
            CK(cudaMemcpy2DAsync(out + off, (size_t)B, d_out, (size_t)bc, (size_t)bc, (size_t)g->n_outputs, cudaMemcpyDeviceToHost, st));
        }
        if (stats) {
            CKR(collect_phase_times(c, g->n_levels, stats));
            float t;
            if (g->n_inputs) { CK(cudaEventSynchronize(c->ev[1])); CK(cudaEventElapsedTime(&t, c->ev[0], c->ev[1])); stats->ms_encrypt += t; }
            if (g->n_outputs) { CK(cudaEventSynchronize(c->ev[3])); CK(cudaEventElapsedTime(&t, c->ev[2], c->ev[3])); stats->ms_decrypt += t; }
        }
    }
    if (stats) {
        CK(cudaEventRecord(c->ev[7], st));
        CK(cudaEventSynchronize(c->ev[7]));
        float t; CK(cudaEventElapsedTime(&t, c->ev[6], c->ev[7])); stats->ms_total += t;
    }
    CK(cudaStreamSynchronize(st));
    return FBS_OK;
}

// ------------------------------------------------------------------------------------------------------
// cleartext evaluation on the GPU
// ------------------------------------------------------------------------------------------------------
extern "C" int fbs_clear_eval(fbs_ctx *c, fbs_prog *g, const uint8_t *in, int64_t B, uint8_t *out, fbs_run_stats *stats)
{
    if (!c || !g || (!in && g->n_inputs) || (!out && g->n_outputs) || B < 1) return fail(FBS_ERR_ARG, "fbs_clear_eval: bad argument");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t nio = (size_t)(g->n_inputs + g->n_outputs + g->n_slots + 1) * B;
    // grow-only scratch of the context (no allocation per call): io bytes, lincomb values, error word
    CKR(grow(&c->d_clear_io, &c->cap_clear_io, nio)); CKR(grow(&c->d_clear_lcv, &c->cap_clear_lcv, (size_t)std::max(1, g->max_lc_per_level) * B + 1));
    u8 *d_io = c->d_clear_io; int32_t *d_lcv = c->d_clear_lcv + 1; int *d_err = c->d_clear_lcv;
    CK(cudaMemsetAsync(d_err, 0, 4, st));
    u8 *d_in = d_io, *d_out = d_in + (size_t)g->n_inputs * B, *d_vals = d_out + (size_t)g->n_outputs * B;
    if (g->n_inputs) CK(cudaMemcpyAsync(d_in, in, (size_t)g->n_inputs * B, cudaMemcpyHostToDevice, st));
    if (stats) CK(cudaEventRecord(c->ev[6], st));
    ClearArgs a{};
    a.lc_level_ptr = g->d_lc_level_ptr; a.bs_level_ptr = g->d_bs_level_ptr; a.lc_ptr = g->d_lc_ptr; a.lc_slot = g->d_lc_slot; a.lc_coef = g->d_lc_coef;
    a.lc_const = g->d_lc_const; a.bs_lc = g->d_bs_lc; a.bs_slot = g->d_bs_slot; a.bs_tab_ptr = g->d_bs_tab_ptr; a.in_slot = g->d_in_slot; a.bs_tab = g->d_tab;
    a.out_ptr = g->d_out_ptr; a.out_slot = g->d_out_slot; a.out_coef = g->d_out_coef; a.out_const = g->d_out_const;
    a.in = d_in; a.vals = d_vals; a.out = d_out; a.lcv = d_lcv; a.err = d_err; a.B = B; a.n_inputs = g->n_inputs; a.n_levels = g->n_levels; a.n_outputs = g->n_outputs;
    k_clear_eval<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(a);
    CK(cudaGetLastError());
    if (stats) CK(cudaEventRecord(c->ev[7], st));
    int err = 0;
    if (g->n_outputs) CK(cudaMemcpyAsync(out, d_out, (size_t)g->n_outputs * B, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (stats) { float t; CK(cudaEventElapsedTime(&t, c->ev[6], c->ev[7])); stats->ms_total += t; stats->n_launches += 1; stats->n_pbs += (int64_t)g->n_boots * B; }
    if (err) return fail(FBS_ERR_ARG, "cleartext evaluation: table index out of range at bootstrap #" + std::to_string(err - 1));
    return FBS_OK;
}

// ------------------------------------------------------------------------------------------------------
// PBS batch + parity taps, built on a one-level program: count inputs, identity lincombs, one table each
// ------------------------------------------------------------------------------------------------------
static int make_pbs_prog(fbs_ctx *c, int p, const uint8_t *tables, int tab_stride, const uint8_t *tlen, const int32_t *modes, int64_t count, fbs_prog **out)
{
    const int n = (int)count;
    std::vector<int32_t> lvl = {0, n}, lc_ptr(n + 1), lc_slot(n), lc_coef(n, 1), lc_const(n, 0), bs_lc(n), bs_slot(n), tab_ptr(n + 1), mode(n), in_slot(n),
                         out_ptr(n + 1), out_slot(n), out_coef(n, 1), out_const(n, 0);
    std::vector<u8> tab;
    tab_ptr[0] = 0;
    for (int i = 0; i < n; i++) {
        lc_ptr[i] = i; lc_slot[i] = i; bs_lc[i] = i; bs_slot[i] = n + i; in_slot[i] = i; out_ptr[i] = i; out_slot[i] = n + i;
        mode[i] = modes ? modes[i] : 1;
        for (int t = 0; t < tlen[i]; t++) tab.push_back(tables[(size_t)i * tab_stride + t]);
        tab_ptr[i + 1] = (int)tab.size();
    }
    lc_ptr[n] = n; out_ptr[n] = n;
    fbs_prog_desc d{};
    d.p = p; d.n_inputs = n; d.n_lincombs = n; d.n_boots = n; d.n_levels = 1; d.n_slots = 2 * n; d.n_outputs = n; d.contiguous_levels = 1;
    d.lc_level_ptr = lvl.data(); d.bs_level_ptr = lvl.data(); d.lc_ptr = lc_ptr.data(); d.lc_slot = lc_slot.data(); d.lc_coef = lc_coef.data();
    d.lc_const = lc_const.data(); d.bs_lc = bs_lc.data(); d.bs_slot = bs_slot.data(); d.bs_tab_ptr = tab_ptr.data(); d.bs_tab = tab.data();
    d.bs_mode = mode.data(); d.in_slot = in_slot.data(); d.out_ptr = out_ptr.data(); d.out_slot = out_slot.data(); d.out_coef = out_coef.data(); d.out_const = out_const.data();
    return fbs_prog_load(c, &d, out);
}

extern "C" int fbs_pbs_batch(fbs_ctx *c, int32_t p, const uint8_t *msgs, const uint8_t *tables, const uint8_t *tlen, const int32_t *modes,
                             int64_t count, uint64_t enc_seed, uint8_t *out, fbs_run_stats *stats)
{
    if (!c || !msgs || !tables || !tlen || !out || count < 1 || count > (1 << 24)) return fail(FBS_ERR_ARG, "fbs_pbs_batch: bad argument");
    if (!c->have_keys) return fail(FBS_ERR_STATE, "fbs_pbs_batch before fbs_keygen");
    CK(cudaSetDevice(c->device));
    fbs_prog *g = nullptr;
    CKR(make_pbs_prog(c, p, tables, 2 * p, tlen, modes, count, &g));
    cudaStream_t st = c->stream;
    int rc = FBS_OK;
    do {
        if ((rc = grow(&c->d_wires, &c->cap_wires, (size_t)g->n_slots * ct_words(c))) != FBS_OK) break;
        if ((rc = grow(&c->d_io, &c->cap_io, (size_t)2 * count + 16)) != FBS_OK) break;
        u8 *d_in = c->d_io, *d_out = c->d_io + count;
        if (cudaMemcpyAsync(d_in, msgs, count, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = fail(FBS_ERR_CUDA, "pbs_batch H2D"); break; }
        // B = 1 with one input per PBS: input i, instance 0 -> ciphertext id i
        if ((rc = fbs_encrypt_inputs(c, g, d_in, 1, 0, 1, enc_seed, c->d_wires, st)) != FBS_OK) break;
        if ((rc = run_level_impl(c, g, 0, -1, -1, 1, c->d_wires, st, stats, nullptr, nullptr, stats != nullptr)) != FBS_OK) break;
        if ((rc = fbs_decrypt_outputs(c, g, 1, c->d_wires, d_out, st)) != FBS_OK) break;
        if (cudaMemcpyAsync(out, d_out, count, cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = fail(FBS_ERR_CUDA, "pbs_batch D2H"); break; }
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = fail(FBS_ERR_CUDA, std::string("pbs_batch sync: ") + cudaGetErrorString(cudaGetLastError())); break; }
        if (stats) { stats->n_launches += 2; stats->ms_total += stats->ms_lincomb + stats->ms_keyswitch + stats->ms_blind_rotate; }
    } while (0);
    fbs_prog_free(g);
    return rc;
}

extern "C" int fbs_debug_get_keys(fbs_ctx *c, uint8_t *s_lwe, uint8_t *s_big, uint64_t *ksk, uint64_t *bsk_coef)
{
    if (!c || !c->have_keys) return fail(FBS_ERR_STATE, "fbs_debug_get_keys before fbs_keygen");
    CK(cudaSetDevice(c->device));
    const fbs_params &P = c->P;
    if (s_lwe) CK(cudaMemcpy(s_lwe, c->d_s_lwe, P.n, cudaMemcpyDeviceToHost));
    if (s_big) CK(cudaMemcpy(s_big, c->d_s_big, (size_t)P.k * P.N, cudaMemcpyDeviceToHost));
    if (ksk) CK(cudaMemcpy(ksk, c->d_ksk, (size_t)P.k * P.N * P.ks_l * (P.n + 1) * 8, cudaMemcpyDeviceToHost));
    if (bsk_coef) {
        if (!c->d_bsk_coef) return fail(FBS_ERR_STATE, "coefficient-domain BSK is only kept for key sizes <= 64 MiB");
        CK(cudaMemcpy(bsk_coef, c->d_bsk_coef, (size_t)c->n_ggsw * (P.k + 1) * P.bsk_l * (P.k + 1) * P.N * 8, cudaMemcpyDeviceToHost));
    }
    return FBS_OK;
}
extern "C" int fbs_debug_ntt(fbs_ctx *c, uint64_t *polys, int64_t count, int32_t inverse)
{
    if (!c || !polys || count < 1) return fail(FBS_ERR_ARG, "fbs_debug_ntt: bad argument");
    CK(cudaSetDevice(c->device));
    ntt_launch_fn nf = ntt_for(c->logN);
    if (!nf) return fail(FBS_ERR_ARG, "no NTT kernel for this N");
    u64 *d_a = nullptr, *d_b = nullptr; const size_t words = (size_t)count * c->P.N;
    CKR(dev_alloc(&d_a, words)); CKR(dev_alloc(&d_b, words));
    CK(h2d_sync(d_a, polys, words * 8));
    nf(d_a, d_b, inverse ? 1 : 0, c->d_psi_rev, c->d_psi_inv_rev, c->ninv[0], c->ninv[1], count, c->stream, c->P.k + 1);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(polys, d_b, words * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_a); cudaFree(d_b);
    return FBS_OK;
}
extern "C" int fbs_debug_encrypt(fbs_ctx *c, int32_t p, const int32_t *msgs, const uint64_t *ct_ids, int64_t count, uint64_t enc_seed, uint64_t *out_cts)
{
    if (!c || !msgs || !out_cts || count < 1) return fail(FBS_ERR_ARG, "fbs_debug_encrypt: bad argument");
    if (!c->have_keys) return fail(FBS_ERR_STATE, "fbs_debug_encrypt before fbs_keygen");
    CK(cudaSetDevice(c->device));
    int32_t *d_m = nullptr; u64 *d_id = nullptr, *d_ct = nullptr; const size_t CT = ct_words(c);
    CKR(dev_alloc(&d_m, count)); CKR(dev_alloc(&d_id, count)); CKR(dev_alloc(&d_ct, (size_t)count * CT));
    CK(h2d_sync(d_m, msgs, count * 4));
    if (ct_ids) CK(h2d_sync(d_id, ct_ids, count * 8));
    EncArgs a{};
    a.msgs32 = d_m; a.ct_ids = ct_ids ? d_id : nullptr; a.out = d_ct; a.s_big = c->d_s_big; a.D = c->P.k * c->P.N; a.p = p;
    a.enc_seed = enc_seed; a.noise_scale = c->P.glwe_noise; a.B = 1; a.B_total = 1;
    k_encrypt<<<(unsigned)count, 256, 0, c->stream>>>(a);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out_cts, d_ct, (size_t)count * CT * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_m); cudaFree(d_id); cudaFree(d_ct);
    return FBS_OK;
}
extern "C" int fbs_debug_decrypt(fbs_ctx *c, int32_t p, const uint64_t *cts, int64_t count, int32_t *out)
{
    if (!c || !cts || !out || count < 1) return fail(FBS_ERR_ARG, "fbs_debug_decrypt: bad argument");
    if (!c->have_keys) return fail(FBS_ERR_STATE, "fbs_debug_decrypt before fbs_keygen");
    CK(cudaSetDevice(c->device));
    u64 *d_ct = nullptr; int32_t *d_o = nullptr; const size_t CT = ct_words(c);
    CKR(dev_alloc(&d_ct, (size_t)count * CT)); CKR(dev_alloc(&d_o, count));
    CK(h2d_sync(d_ct, cts, (size_t)count * CT * 8));
    OutArgs a{};
    a.wires = d_ct; a.s_big = c->d_s_big; a.out32 = d_o; a.B = 1; a.D = c->P.k * c->P.N; a.p = p;
    k_decrypt<<<(unsigned)count, 256, 0, c->stream>>>(a);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out, d_o, count * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_ct); cudaFree(d_o);
    return FBS_OK;
}
extern "C" int fbs_debug_pbs(fbs_ctx *c, int32_t p, const uint64_t *in_cts, const uint8_t *tables, const uint8_t *tlen, const int32_t *modes,
                             int64_t count, uint64_t *out_cts, uint64_t *tap_ks, uint16_t *tap_ms, uint64_t *tap_acc)
{
    if (!c || !in_cts || !tables || !tlen || !out_cts || count < 1 || count > 65536) return fail(FBS_ERR_ARG, "fbs_debug_pbs: bad argument");
    if (!c->have_keys) return fail(FBS_ERR_STATE, "fbs_debug_pbs before fbs_keygen");
    CK(cudaSetDevice(c->device));
    fbs_prog *g = nullptr;
    CKR(make_pbs_prog(c, p, tables, 2 * p, tlen, modes, count, &g));
    const size_t CT = ct_words(c); const fbs_params &P = c->P; cudaStream_t st = c->stream;
    u64 *d_w = nullptr, *d_ks = nullptr, *d_acc = nullptr;
    int rc = FBS_OK;
    do {
        if ((rc = dev_alloc(&d_w, (size_t)2 * count * CT)) != FBS_OK) break;
        if ((rc = dev_alloc(&d_ks, (size_t)(count + 16) * (P.n + 1))) != FBS_OK) break;
        if ((rc = dev_alloc(&d_acc, (size_t)count * (P.k + 1) * P.N)) != FBS_OK) break;
        if (h2d_sync(d_w, in_cts, (size_t)count * CT * 8) != cudaSuccess) { rc = fail(FBS_ERR_CUDA, std::string("debug_pbs H2D: ") + cudaGetErrorString(cudaGetLastError())); break; }
        if ((rc = run_level_impl(c, g, 0, -1, -1, 1, d_w, st, nullptr, d_ks, d_acc, false)) != FBS_OK) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = fail(FBS_ERR_CUDA, std::string("debug_pbs: ") + cudaGetErrorString(cudaGetLastError())); break; }
        cudaError_t e = cudaMemcpy(out_cts, d_w + (size_t)count * CT, (size_t)count * CT * 8, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && tap_ks) e = cudaMemcpy(tap_ks, d_ks, (size_t)count * (P.n + 1) * 8, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && tap_ms) e = cudaMemcpy(tap_ms, c->d_ms, (size_t)count * (P.n + 1) * 2, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && tap_acc) e = cudaMemcpy(tap_acc, d_acc, (size_t)count * (P.k + 1) * P.N * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(FBS_ERR_CUDA, std::string("debug_pbs copies: ") + cudaGetErrorString(e));
    } while (0);
    if (d_w) cudaFree(d_w); if (d_ks) cudaFree(d_ks); if (d_acc) cudaFree(d_acc);
    fbs_prog_free(g);
    return rc;
}

// multi-value tap: a one-level program with `count` inputs, identity lincombs, T tables per lincomb, groups = the lincombs
extern "C" int fbs_debug_pbs_multi(fbs_ctx *c, int32_t p, const uint64_t *in_cts, const uint8_t *tables, const uint8_t *tlen, const int32_t *modes,
                                   int64_t count, int32_t T, uint64_t *out_cts, uint64_t *tap_acc)
{
    if (!c || !in_cts || !tables || !tlen || !out_cts || count < 1 || count > 4096 || T < 1 || T > 64) return fail(FBS_ERR_ARG, "fbs_debug_pbs_multi: bad argument");
    if (!c->have_keys) return fail(FBS_ERR_STATE, "fbs_debug_pbs_multi before fbs_keygen");
    CK(cudaSetDevice(c->device));
    const int n = (int)count, nb = n * T;
    std::vector<int32_t> lvl = {0, n}, blvl = {0, nb}, lc_ptr(n + 1), lc_slot(n), lc_coef(n, 1), lc_const(n, 0), bs_lc(nb), bs_slot(nb), tab_ptr(nb + 1), mode(nb), in_slot(n),
                         out_ptr(nb + 1), out_slot(nb), out_coef(nb, 1), out_const(nb, 0), glvl = {0, n}, gfirst(n + 1);
    std::vector<u8> tab;
    tab_ptr[0] = 0;
    for (int i = 0; i < n; i++) { lc_ptr[i] = i; lc_slot[i] = i; in_slot[i] = i; gfirst[i] = i * T; }
    lc_ptr[n] = n; gfirst[n] = nb;
    for (int q = 0; q < nb; q++) {
        bs_lc[q] = q / T; bs_slot[q] = n + q; out_ptr[q] = q; out_slot[q] = n + q; mode[q] = modes ? modes[q] : 1;
        for (int t = 0; t < tlen[q]; t++) tab.push_back(tables[(size_t)q * 2 * p + t]);
        tab_ptr[q + 1] = (int)tab.size();
    }
    out_ptr[nb] = nb;
    fbs_prog_desc d{};
    d.p = p; d.n_inputs = n; d.n_lincombs = n; d.n_boots = nb; d.n_levels = 1; d.n_slots = n + nb; d.n_outputs = nb; d.contiguous_levels = 1;
    d.lc_level_ptr = lvl.data(); d.bs_level_ptr = blvl.data(); d.lc_ptr = lc_ptr.data(); d.lc_slot = lc_slot.data(); d.lc_coef = lc_coef.data();
    d.lc_const = lc_const.data(); d.bs_lc = bs_lc.data(); d.bs_slot = bs_slot.data(); d.bs_tab_ptr = tab_ptr.data(); d.bs_tab = tab.data();
    d.bs_mode = mode.data(); d.in_slot = in_slot.data(); d.out_ptr = out_ptr.data(); d.out_slot = out_slot.data(); d.out_coef = out_coef.data(); d.out_const = out_const.data();
    d.n_groups = n; d.grp_level_ptr = glvl.data(); d.grp_first = gfirst.data();
    fbs_prog *g = nullptr;
    CKR(fbs_prog_load(c, &d, &g));
    const size_t CT = ct_words(c); const fbs_params &P = c->P; cudaStream_t st = c->stream;
    u64 *d_w = nullptr, *d_acc = nullptr;
    int rc = FBS_OK;
    do {
        if ((rc = dev_alloc(&d_w, (size_t)(n + nb) * CT)) != FBS_OK) break;
        if ((rc = dev_alloc(&d_acc, (size_t)count * (P.k + 1) * P.N)) != FBS_OK) break;
        if (h2d_sync(d_w, in_cts, (size_t)count * CT * 8) != cudaSuccess) { rc = fail(FBS_ERR_CUDA, "debug_pbs_multi H2D"); break; }
        if ((rc = run_level_impl(c, g, 0, -1, -1, 1, d_w, st, nullptr, nullptr, d_acc, false)) != FBS_OK) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = fail(FBS_ERR_CUDA, std::string("debug_pbs_multi: ") + cudaGetErrorString(cudaGetLastError())); break; }
        cudaError_t e = cudaMemcpy(out_cts, d_w + (size_t)count * CT, (size_t)nb * CT * 8, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && tap_acc) e = cudaMemcpy(tap_acc, d_acc, (size_t)count * (P.k + 1) * P.N * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(FBS_ERR_CUDA, std::string("debug_pbs_multi copies: ") + cudaGetErrorString(e));
    } while (0);
    if (d_w) cudaFree(d_w); if (d_acc) cudaFree(d_acc);
    fbs_prog_free(g);
    return rc;
}

// ------------------------------------------------------------------------------------------------------
// integer-multiply roofline probe: independent mad.wide.u32 chains, reports 32x32->64 multiplies per second
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_imad_peak(u64 *sink, int iters, u32 a0, u32 b0)
{
    u64 acc[8];
    u32 a = a0 + threadIdx.x, b = b0 + blockIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a), "r"(b));
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    if (s == 0x1234567ULL) sink[0] = s;
}
extern "C" int fbs_measure_int_peak(fbs_ctx *c, double *mul32_per_s)
{
    if (!c || !mul32_per_s) return fail(FBS_ERR_ARG, "fbs_measure_int_peak: bad argument");
    CK(cudaSetDevice(c->device));
    u64 *d = nullptr; CKR(dev_alloc(&d, 1));
    const int iters = 1 << 14, blocks = c->sm_count * 8;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(c->ev[0], c->stream));
        k_imad_peak<<<blocks, 256, 0, c->stream>>>(d, iters, 12345u, 6789u);
        CK(cudaEventRecord(c->ev[1], c->stream));
        CK(cudaEventSynchronize(c->ev[1]));
        float ms; CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
        double rate = (double)blocks * 256 * 8 * iters / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaFree(d);
    *mul32_per_s = best;
    return FBS_OK;
}

// ------------------------------------------------------------------------------------------------------
// peer-mapped wire buffers for the fused sample-extract + exchange of node-sharded levels
// ------------------------------------------------------------------------------------------------------
static inline size_t flag_offset(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
extern "C" int fbs_wires_alloc(fbs_ctx *c, size_t bytes, uint64_t **out)
{
    if (!c || !out || bytes == 0) return fail(FBS_ERR_ARG, "fbs_wires_alloc: bad argument");
    CK(cudaSetDevice(c->device));
    // a plain cudaMalloc allocation (exportable with CUDA IPC), followed by one flag page for the device-side level hand-off
    CK(cudaMalloc((void **)out, flag_offset(bytes) + FBS_FLAG_PAGE));
    CK(cudaMemset((char *)*out + flag_offset(bytes), 0, FBS_FLAG_PAGE));
    return FBS_OK;
}
extern "C" int fbs_wires_free(fbs_ctx *c, uint64_t *p)
{
    if (!c) return fail(FBS_ERR_ARG, "fbs_wires_free: null ctx");
    CK(cudaSetDevice(c->device));
    if (p) CK(cudaFree(p));
    return FBS_OK;
}
extern "C" int fbs_ipc_export(fbs_ctx *c, const uint64_t *dev_ptr, unsigned char handle[64])
{
    if (!c || !dev_ptr || !handle) return fail(FBS_ERR_ARG, "fbs_ipc_export: bad argument");
    CK(cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, (void *)dev_ptr));
    memcpy(handle, &h, 64);
    return FBS_OK;
}
extern "C" int fbs_ipc_import(fbs_ctx *c, const unsigned char handle[64], uint64_t **out)
{
    if (!c || !handle || !out) return fail(FBS_ERR_ARG, "fbs_ipc_import: bad argument");
    CK(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle((void **)out, h, cudaIpcMemLazyEnablePeerAccess));
    return FBS_OK;
}
extern "C" int fbs_ipc_close(fbs_ctx *c, uint64_t *peer_ptr)
{
    if (!c) return fail(FBS_ERR_ARG, "fbs_ipc_close: null ctx");
    CK(cudaSetDevice(c->device));
    if (peer_ptr) CK(cudaIpcCloseMemHandle(peer_ptr));
    return FBS_OK;
}
extern "C" int fbs_set_peers(fbs_ctx *c, uint64_t *wires_local, size_t wires_bytes, uint64_t *const *peer_wires, int32_t n_peers, int32_t rank)
{
    if (!c || n_peers < 0 || n_peers > 8 || (n_peers && (!peer_wires || !wires_local || wires_bytes == 0)) || rank > n_peers)
        return fail(FBS_ERR_ARG, "fbs_set_peers: at most 8 peers, rank in [0, n_peers] (or < 0: peer stores only, caller orders the levels)");
    CK(cudaSetDevice(c->device));
    CK(cudaDeviceSynchronize());                    // nothing in flight still uses the old binding
    c->n_peers = n_peers; c->peer_rank = rank; c->epoch = 0;
    c->peer_local = n_peers ? wires_local : nullptr; c->peer_bytes = n_peers ? wires_bytes : 0;
    c->flags_local = n_peers ? (u64 *)((char *)wires_local + flag_offset(wires_bytes)) : nullptr;
    for (int i = 0; i < n_peers; i++) {
        if (!peer_wires[i]) return fail(FBS_ERR_ARG, "fbs_set_peers: null peer pointer");
        c->peers[i] = peer_wires[i];
        c->peer_flags[i] = (u64 *)((char *)peer_wires[i] + flag_offset(wires_bytes));
    }
    if (n_peers) {
        CK(cudaMemset(c->flags_local, 0, FBS_FLAG_PAGE));   // the caller barriers across ranks before the first level
        if (!c->d_sync_err) CKR(dev_alloc(&c->d_sync_err, 1));
        CK(cudaMemset(c->d_sync_err, 0, sizeof(int)));
    }
    return FBS_OK;
}
extern "C" int fbs_level_sync(fbs_ctx *c, uint64_t *wires_local, void *stream)
{
    if (!c || !wires_local) return fail(FBS_ERR_ARG, "fbs_level_sync: bad argument");
    if (c->n_peers == 0 || wires_local != c->peer_local || c->peer_rank < 0) return fail(FBS_ERR_STATE, "fbs_level_sync: buffer is not registered with fbs_set_peers for the device-side hand-off");
    CK(cudaSetDevice(c->device));
    if (c->epoch > 0) { k_level_wait<<<1, 32, 0, (cudaStream_t)stream>>>(c->flags_local, c->n_peers + 1, c->peer_rank, c->epoch, c->d_sync_err); CK(cudaGetLastError()); }
    return FBS_OK;
}
extern "C" int fbs_sync_status(fbs_ctx *c, int32_t *timed_out)
{
    if (!c || !timed_out) return fail(FBS_ERR_ARG, "fbs_sync_status: bad argument");
    *timed_out = 0;
    if (!c->d_sync_err) return FBS_OK;
    CK(cudaSetDevice(c->device));
    int v = 0;
    CK(cudaMemcpy(&v, c->d_sync_err, sizeof(int), cudaMemcpyDeviceToHost));
    *timed_out = v;
    return FBS_OK;
}

#ifdef FBS_PHASE_CLK
// profiling builds only (not part of include/fbs_b200.h): read and reset the per-phase clock sums of k_blind_rotate2
extern "C" int fbs_debug_phase_clk(unsigned long long *out8)
{
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out8, g_phase_clk, sizeof(unsigned long long) * 8));
    unsigned long long z[8] = {0};
    CK(cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z)));
    return FBS_OK;
}
#endif
