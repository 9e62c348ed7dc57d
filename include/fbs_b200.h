/*
 * fbs_b200.h -- C ABI of the B200 encrypted executor for tfhe_fbs_map circuits.
 *
 * The reference (ssmiler/tfhe_fbs_map) has no FFI: its operator API for this path is the Python method
 * LutExecEnv.eval (reference fbs_mapper/fbs_exec_env.py:208-229) and BitExecEnv.eval
 * (reference fbs_mapper/bit_exec_env.py:173-194), called from the CLI at reference
 * fbs_mapper/map_circuit.py:140,174.  The entry points below are what a ctypes binding of that path
 * binds (INTEGRATION.md shows the stub); each cites the reference interface it stands in for.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function returns FBS_OK (0) or a
 * negative error code and never throws/aborts; fbs_last_error() returns the message of the calling
 * thread's last failure.  "dev" pointers are CUDA device pointers on the context's device; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  One context per host thread.
 *
 * SECURITY CAVEAT: benchmark harness.  Keys, masks and noise come from a counter-based splitmix64 construction keyed by
 * the seeds passed in (shared with the test oracle): NOT a cryptographically secure generator.  The caller
 * must never reuse an (enc_seed, instance id) pair: ciphertexts that share it share mask and noise.
 */
#ifndef FBS_B200_H
#define FBS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FBS_OK 0
#define FBS_ERR_ARG (-1)      /* invalid argument / unsupported parameter shape */
#define FBS_ERR_CUDA (-2)     /* CUDA runtime failure (message has the CUDA error string) */
#define FBS_ERR_STATE (-3)    /* call order violated (e.g. eval before keygen) */
#define FBS_ERR_NOMEM (-4)

/* TFHE parameter set.  Ciphertext modulus is q = p1*p2 = 0x3FFE8001 * 0x3FFF4001 (60 bits). */
typedef struct fbs_params {
    int32_t n;          /* small LWE dimension                         */
    int32_t k;          /* GLWE dimension                              */
    int32_t N;          /* polynomial size (power of two, 256..2048)   */
    int32_t bsk_l;      /* blind-rotate decomposition levels           */
    int32_t bsk_beta;   /* log2 blind-rotate base                      */
    int32_t ks_l;       /* key-switch levels                           */
    int32_t ks_beta;    /* log2 key-switch base (<= 8)                 */
    int32_t bsk_unroll; /* key bits per blind-rotation step: 0/1 classic; 2 or 3: 2^m - 1 GGSW per key group, bsk_l = 1 */
    uint64_t lwe_noise; /* round(sigma_lwe  * Q)                       */
    uint64_t glwe_noise;/* round(sigma_glwe * Q)                       */
} fbs_params;

/*
 * Levelised FBS program: flat arrays produced from LutExecEnv.instructions (reference
 * fbs_exec_env.py:63-70) by tfhe_fbs_map_b200/levelize.py.  A "wire" is an Input or Bootstrap node
 * (reference fbs_exec_env.py:30-35,51-61) and lives in a ciphertext slot; LinearProd nodes (reference
 * fbs_exec_env.py:37-49) are never stored, they appear as CSR rows over wire slots.  Level lv holds the
 * lincombs [lc_level_ptr[lv], lc_level_ptr[lv+1]) and the bootstraps [bs_level_ptr[lv], bs_level_ptr[lv+1]);
 * a lincomb of level lv only reads slots written by inputs or by bootstraps of levels < lv.
 */
typedef struct fbs_prog_desc {
    int32_t p;            /* plaintext modulus = --fbs_size (reference map_circuit.py:98)           */
    int32_t n_inputs, n_lincombs, n_boots, n_levels, n_slots, n_outputs;
    int32_t contiguous_levels;    /* != 0: slots are never recycled (levelize reuse_slots=False / shard_pad): fbs_run_level accepts node
                                     sub-ranges; without it a sub-range would corrupt recycled slots and is refused */
    const int32_t *lc_level_ptr;  /* [n_levels+1]                                                 */
    const int32_t *bs_level_ptr;  /* [n_levels+1]                                                 */
    const int32_t *lc_ptr;        /* [n_lincombs+1] CSR row pointers                              */
    const int32_t *lc_slot;       /* [nnz] operand wire slot                                      */
    const int32_t *lc_coef;       /* [nnz] integer coefficient (reference fbs_exec_env.py:215-217)*/
    const int32_t *lc_const;      /* [n_lincombs] const_coef                                      */
    const int32_t *bs_lc;         /* [n_boots] lincomb feeding this bootstrap                     */
    const int32_t *bs_slot;       /* [n_boots] output wire slot                                   */
    const int32_t *bs_tab_ptr;    /* [n_boots+1] into bs_tab                                      */
    const uint8_t *bs_tab;        /* table entries (reference fbs_exec_env.py:218-220)            */
    const int32_t *bs_mode;       /* [n_boots] s = tv[x]+tv[x+p] (map_to_fbs.py:81-98), 1 if len<=p */
    const int32_t *in_slot;       /* [n_inputs]                                                   */
    const int32_t *out_ptr;       /* [n_outputs+1] outputs are lincombs too (1-x, pass-through, const) */
    const int32_t *out_slot, *out_coef, *out_const;
    /* Multi-value bootstrap: n_groups > 0 switches it on.  A group = the bootstraps of one level that share a
     * lincomb (reference fbs_exec_env.py:93-100 de-duplicates LinearProds, which makes the sharing visible); they are contiguous
     * because a level's bootstraps are sorted by lincomb.  One blind rotation per (group, instance) of a table-independent base
     * polynomial, then one sparse polynomial product + sample extraction per table.  grp_first[g] = first bootstrap of group g
     * (n_groups + 1 entries), grp_level_ptr[lv] = first group of level lv.  NULL / 0: one rotation per bootstrap. */
    const int32_t *grp_level_ptr, *grp_first;
    int32_t n_groups, reserved2;
} fbs_prog_desc;

typedef struct fbs_run_stats {
    int64_t n_pbs;         /* blind rotations executed (bootstrap nodes -- or multi-value groups -- x instances) */
    int64_t n_launches;    /* CUDA kernels launched by this call               */
    float ms_total;        /* device time of the call (CUDA events)            */
    float ms_encrypt, ms_lincomb, ms_keyswitch, ms_blind_rotate, ms_decrypt;
    float reserved[2];
} fbs_run_stats;

typedef struct fbs_ctx fbs_ctx;    /* device, secret keys, BSK (NTT domain), KSK, scratch */
typedef struct fbs_prog fbs_prog;  /* device-resident levelised program tied to one ctx    */

const char *fbs_last_error(void);
int fbs_abi_version(void);

/* ---- context / keys ------------------------------------------------------------------------------ */
int fbs_ctx_create(const fbs_params *params, int device, uint64_t seed, fbs_ctx **out);
int fbs_keygen(fbs_ctx *ctx);                 /* seeded, deterministic: identical on every rank/device */
int fbs_ctx_destroy(fbs_ctx *ctx);
/* Blind-rotation kernel choice for launches with fewer jobs than SMs: 0 = auto (split each bootstrap over a thread-block
 * cluster of up to 8 CTAs when that still fits one wave), 1 = never, 2 / 4 / 8 = always that cluster size (tests, sweeps),
 * 12 / 14 / 18 = that size with two threads per ring element (one per RNS prime).  All choices produce bit-identical ciphertexts. */
int fbs_ctx_set_cluster(fbs_ctx *ctx, int32_t mode);
int fbs_ctx_info(const fbs_ctx *ctx, int32_t *sm_count, int64_t *bsk_bytes, int64_t *ksk_bytes, int32_t *br_smem_bytes);

/* ---- program ------------------------------------------------------------------------------------- */
int fbs_prog_load(fbs_ctx *ctx, const fbs_prog_desc *desc, fbs_prog **out);   /* host arrays are copied */
int fbs_prog_free(fbs_prog *prog);

/* ---- one-call evaluation with HOST buffers: the drop-in for LutExecEnv.eval ------------------------
 * in  : [n_inputs][B] bits, row-major (one row per Input in instruction order, reference fbs_exec_env.py:213-214)
 * out : [n_outputs][B] values mod 2p (one row per entry of LutExecEnv.outputs, reference fbs_exec_env.py:225-229)
 * inst_offset/B_total number the instances globally so that sharded ranks encrypt with distinct randomness.
 * Instances are processed in chunks that fit `max_wire_bytes` of device memory (0 = default budget). */
int fbs_eval_bits(fbs_ctx *ctx, fbs_prog *prog, const uint8_t *in, int64_t B, int64_t inst_offset, int64_t B_total,
                  uint64_t enc_seed, size_t max_wire_bytes, uint8_t *out, fbs_run_stats *stats);

/* ---- split form with device-resident ciphertexts (timing, multi-GPU node sharding) ------------------
 * wires_dev: [n_slots][B][k*N+1] uint64, caller-allocated (fbs_wires_bytes). */
int fbs_wires_bytes(const fbs_ctx *ctx, const fbs_prog *prog, int64_t B, size_t *bytes);
int fbs_encrypt_inputs(fbs_ctx *ctx, fbs_prog *prog, const uint8_t *in_dev, int64_t B, int64_t inst_offset,
                       int64_t B_total, uint64_t enc_seed, uint64_t *wires_dev, void *stream);
/* runs bootstraps [node_begin, node_end) of level `level` (indices relative to the level; -1,-1 = all; for a multi-value
 * program, n_groups > 0, the indices count GROUPS: all tables of a group are produced by the rank that rotates it).  A proper
 * sub-range needs a program with contiguous_levels != 0 (recycled slots would be corrupted otherwise: refused). */
int fbs_run_level(fbs_ctx *ctx, fbs_prog *prog, int32_t level, int32_t node_begin, int32_t node_end, int64_t B,
                  uint64_t *wires_dev, void *stream, fbs_run_stats *stats);
int fbs_run(fbs_ctx *ctx, fbs_prog *prog, int64_t B, uint64_t *wires_dev, void *stream, fbs_run_stats *stats);
int fbs_decrypt_outputs(fbs_ctx *ctx, fbs_prog *prog, int64_t B, const uint64_t *wires_dev, uint8_t *out_dev, void *stream);

/* ---- fused exchange for node-sharded levels (BASELINE.json configs[3]) ---------------------------------
 * Every rank keeps a replica of the wire buffer (reference semantics: the level loop of fbs_exec_env.py:208-229 with the
 * bootstraps of a level split across GPUs).  Buffers come from fbs_wires_alloc (payload + one flag page) and travel between
 * processes as CUDA IPC handles (64 opaque bytes).  fbs_set_peers BINDS the peer replicas to one local buffer: when
 * fbs_run_level runs on exactly that buffer, (1) the sample-extract epilogue of the blind-rotation kernel stores each
 * output ciphertext into the local buffer AND into every peer replica (NVLink peer stores; replaces the per-level
 * all-gather), and (2) levels are ordered ON THE DEVICE: after a level every rank publishes an epoch into the peers' flag
 * pages (system-scope release) and the next level's first kernel waits for all of them (bounded spin) -- no host
 * synchronisation between levels.  Every rank must call fbs_run_level for every level, in the same order (empty node
 * ranges included), and barrier once on the host after fbs_set_peers / after (re-)encrypting inputs.  Any other wires
 * pointer (fbs_eval_bits, fbs_pbs_batch, parity taps) never touches the peers.  rank = this process's index among the
 * n_peers + 1 participants (rank < 0: peer stores only, the caller orders the levels itself, e.g. with a host barrier);
 * n_peers = 0 unbinds.  fbs_sync_status reports a peer that never arrived (0 = none, else 1 + rank). */
int fbs_wires_alloc(fbs_ctx *ctx, size_t bytes, uint64_t **out);
int fbs_wires_free(fbs_ctx *ctx, uint64_t *wires_dev);
int fbs_ipc_export(fbs_ctx *ctx, const uint64_t *wires_dev, unsigned char handle[64]);
int fbs_ipc_import(fbs_ctx *ctx, const unsigned char handle[64], uint64_t **peer_wires);
int fbs_ipc_close(fbs_ctx *ctx, uint64_t *peer_wires);
int fbs_set_peers(fbs_ctx *ctx, uint64_t *wires_local, size_t wires_bytes, uint64_t *const *peer_wires, int32_t n_peers, int32_t rank);
/* enqueue a wait until every rank has completed all levels issued so far (end of a run, before decrypting) */
int fbs_level_sync(fbs_ctx *ctx, uint64_t *wires_local, void *stream);
int fbs_sync_status(fbs_ctx *ctx, int32_t *timed_out);

/* ---- PBS micro-benchmark (BASELINE.json configs[4]) -------------------------------------------------
 * count independent bootstraps: encrypt msgs[i] in Z_2p, bootstrap with tables[i*2p .. i*2p+tlen[i]) (modes[i] = table
 * mode s, NULL = 1), decrypt into out[i].  Host buffers; stats->ms_lincomb / ms_keyswitch / ms_blind_rotate cover the
 * bootstrap proper (device time, CUDA events), encryption / decryption / copies are outside them. */
int fbs_pbs_batch(fbs_ctx *ctx, int32_t p, const uint8_t *msgs, const uint8_t *tables, const uint8_t *tlen,
                  const int32_t *modes, int64_t count, uint64_t enc_seed, uint8_t *out, fbs_run_stats *stats);

/* ---- cleartext evaluation of the same program on the GPU ---------------------------------------------
 * The literal counterpart of the reference's hot loop (fbs_exec_env.py:218-220 / bit_exec_env.py:183-185):
 * integer lincomb + table look-up per (node, instance).  Needs no keys. */
int fbs_clear_eval(fbs_ctx *ctx, fbs_prog *prog, const uint8_t *in, int64_t B, uint8_t *out, fbs_run_stats *stats);

/* ---- integer roofline probe: sustained mad.wide.u32 (IMAD.WIDE) rate of the device, 32x32->64 multiplies/s.
 * MEASURED_PEAKS.json holds only HBM and bf16 peaks; BASELINE.json asks for the integer-multiply roofline. */
int fbs_measure_int_peak(fbs_ctx *ctx, double *mul32_per_s);

/* ---- parity taps (used by tests/ only; product code never calls them) ------------------------------- */
int fbs_debug_get_keys(fbs_ctx *ctx, uint8_t *s_lwe, uint8_t *s_big, uint64_t *ksk, uint64_t *bsk_coef);
int fbs_debug_ntt(fbs_ctx *ctx, uint64_t *polys_host, int64_t count, int32_t inverse);
/* one PBS per input ciphertext with every intermediate: ks [count][n+1] u64, ms [count][n+1] u16, acc [count][(k+1)N] */
int fbs_debug_pbs(fbs_ctx *ctx, int32_t p, const uint64_t *in_cts, const uint8_t *tables, const uint8_t *tlen,
                  const int32_t *modes, int64_t count, uint64_t *out_cts, uint64_t *tap_ks, uint16_t *tap_ms, uint64_t *tap_acc);
/* multi-value tap: count inputs, T tables each (tables [count*T][2p]); one rotation per input; out [count*T][kN+1], acc [count][(k+1)N] */
int fbs_debug_pbs_multi(fbs_ctx *ctx, int32_t p, const uint64_t *in_cts, const uint8_t *tables, const uint8_t *tlen,
                        const int32_t *modes, int64_t count, int32_t T, uint64_t *out_cts, uint64_t *tap_acc);
int fbs_debug_encrypt(fbs_ctx *ctx, int32_t p, const int32_t *msgs, const uint64_t *ct_ids, int64_t count, uint64_t enc_seed, uint64_t *out_cts);
int fbs_debug_decrypt(fbs_ctx *ctx, int32_t p, const uint64_t *cts, int64_t count, int32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* FBS_B200_H */
