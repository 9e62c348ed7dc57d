"""ctypes wrapper of oracle/libtfhe_ref.so (CPU restatement of the encrypted pipeline).
TEST INFRASTRUCTURE ONLY -- see the header of oracle/tfhe_ref.c.  Never imported by tfhe_fbs_map_b200/."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libtfhe_ref.so")


def build(force=False):
    src = os.path.join(_HERE, "tfhe_ref.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        r = subprocess.run(["make", "-C", _HERE, "-B", "libtfhe_ref.so"], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB


class RefParams(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("k", ctypes.c_int32), ("N", ctypes.c_int32), ("bsk_l", ctypes.c_int32),
                ("bsk_beta", ctypes.c_int32), ("ks_l", ctypes.c_int32), ("ks_beta", ctypes.c_int32),
                ("bsk_unroll", ctypes.c_int32), ("lwe_noise", ctypes.c_uint64), ("glwe_noise", ctypes.c_uint64)]


class RefProgDesc(ctypes.Structure):
    _P32 = ctypes.POINTER(ctypes.c_int32)
    _P8 = ctypes.POINTER(ctypes.c_uint8)
    _fields_ = [("p", ctypes.c_int32), ("n_inputs", ctypes.c_int32), ("n_lincombs", ctypes.c_int32),
                ("n_boots", ctypes.c_int32), ("n_levels", ctypes.c_int32), ("n_slots", ctypes.c_int32),
                ("n_outputs", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("lc_level_ptr", _P32), ("bs_level_ptr", _P32),
                ("lc_ptr", _P32), ("lc_slot", _P32), ("lc_coef", _P32), ("lc_const", _P32),
                ("bs_lc", _P32), ("bs_slot", _P32), ("bs_tab_ptr", _P32), ("bs_tab", _P8), ("bs_mode", _P32),
                ("in_slot", _P32),
                ("out_ptr", _P32), ("out_slot", _P32), ("out_coef", _P32), ("out_const", _P32)]


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        vp, u64, i32, i64 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int64
        L.ref_ctx_create.restype = vp
        L.ref_ctx_create.argtypes = [ctypes.POINTER(RefParams), u64]
        L.ref_ctx_destroy.argtypes = [vp]
        L.ref_keygen.argtypes = [vp]
        L.ref_get_keys.argtypes = [vp, vp, vp, vp, vp]
        L.ref_encrypt.argtypes = [vp, i32, vp, vp, i64, u64, vp]
        L.ref_decrypt.argtypes = [vp, i32, vp, i64, vp]
        L.ref_phase.restype = u64
        L.ref_phase.argtypes = [vp, vp]
        L.ref_keyswitch.argtypes = [vp, vp, vp]
        L.ref_modswitch.argtypes = [vp, vp, vp]
        L.ref_test_poly.argtypes = [vp, i32, vp, i32, i32, vp]
        L.ref_blind_rotate.argtypes = [vp, vp, vp, vp]
        L.ref_sample_extract.argtypes = [vp, vp, u64, vp]
        L.ref_pbs.argtypes = [vp, i32, vp, vp, i32, i32, vp, vp, vp, vp]
        L.ref_pbs_batch.argtypes = [vp, i32, vp, vp, i32, vp, vp, i64, vp, vp, i32]
        L.ref_pbs_multi.argtypes = [vp, i32, vp, vp, i32, vp, vp, i32, vp, vp]
        L.ref_eval_prog_mv.argtypes = [vp, ctypes.POINTER(RefProgDesc), vp, i64, i64, i64, u64, vp, i32, i32]
        L.ref_ntt_fwd.argtypes = [vp, vp]
        L.ref_ntt_inv.argtypes = [vp, vp]
        L.ref_polymul_schoolbook.argtypes = [i32, vp, vp, vp]
        L.ref_polymul_ntt.argtypes = [vp, vp, vp, vp]
        L.ref_mulmod.restype = u64
        L.ref_mulmod.argtypes = [u64, u64]
        L.ref_rnd64.restype = u64
        L.ref_rnd64.argtypes = [u64, u64, u64]
        L.ref_noise.restype = u64
        L.ref_noise.argtypes = [u64, u64, u64, u64]
        L.ref_decompose.argtypes = [u64, i32, i32, vp]
        L.ref_modswitch_word.restype = ctypes.c_uint32
        L.ref_modswitch_word.argtypes = [u64, i32]
        L.ref_delta.restype = u64
        L.ref_delta.argtypes = [i32]
        L.ref_gadget.restype = u64
        L.ref_gadget.argtypes = [i32, i32]
        L.ref_eval_prog.argtypes = [vp, ctypes.POINTER(RefProgDesc), vp, i64, i64, i64, u64, vp, i32]
        L.ref_eval_prog.restype = i32
        L.ref_max_threads.restype = i32
        _lib = L
    return _lib


FBS_P1, FBS_P2 = 0x3FFE8001, 0x3FFF4001
FBS_Q = FBS_P1 * FBS_P2             # ciphertext modulus: product of two 30-bit NTT primes (60 bits)
GOLDILOCKS_P = FBS_Q   # old name kept for the tests' imports


class RefTFHE:
    """CPU oracle context.  ``ps`` needs attributes n,k,N,bsk_l,bsk_beta,ks_l,ks_beta,lwe_noise_scale,glwe_noise_scale."""

    def __init__(self, ps, seed, keygen=True):
        self.L = lib()
        self.ps = ps
        self.cp = RefParams(ps.n, ps.k, ps.N, ps.bsk_l, ps.bsk_beta, ps.ks_l, ps.ks_beta, getattr(ps, "bsk_unroll", 1),
                            ps.lwe_noise_scale, ps.glwe_noise_scale)
        self.ctx = self.L.ref_ctx_create(ctypes.byref(self.cp), seed)
        self.ct_words = ps.k * ps.N + 1
        if keygen:
            self.L.ref_keygen(self.ctx)

    def __del__(self):
        try:
            self.L.ref_ctx_destroy(self.ctx)
        except Exception:
            pass

    def keys(self):
        ps = self.ps
        s_lwe = np.zeros(ps.n, np.uint8)
        s_big = np.zeros(ps.k * ps.N, np.uint8)
        ksk = np.zeros((ps.k * ps.N * ps.ks_l, ps.n + 1), np.uint64)
        n_ggsw = ps.n_ggsw if hasattr(ps, "n_ggsw") else ps.n
        bsk = np.zeros((n_ggsw, (ps.k + 1) * ps.bsk_l, ps.k + 1, ps.N), np.uint64)
        self.L.ref_get_keys(self.ctx, _p(s_lwe), _p(s_big), _p(ksk), _p(bsk))
        return s_lwe, s_big, ksk, bsk

    def encrypt(self, p, msgs, ct_ids, enc_seed):
        msgs = np.ascontiguousarray(msgs, np.int32)
        ids = np.ascontiguousarray(ct_ids, np.uint64)
        out = np.zeros((len(msgs), self.ct_words), np.uint64)
        self.L.ref_encrypt(self.ctx, p, _p(msgs), _p(ids), len(msgs), enc_seed, _p(out))
        return out

    def decrypt(self, p, cts):
        cts = np.ascontiguousarray(cts, np.uint64)
        out = np.zeros(cts.shape[0], np.int32)
        self.L.ref_decrypt(self.ctx, p, _p(cts), cts.shape[0], _p(out))
        return out

    def phase(self, ct):
        ct = np.ascontiguousarray(ct, np.uint64)
        return int(self.L.ref_phase(self.ctx, _p(ct)))

    def ntt(self, poly, inverse=False):
        a = np.ascontiguousarray(poly, np.uint64).copy()
        (self.L.ref_ntt_inv if inverse else self.L.ref_ntt_fwd)(self.ctx, _p(a))
        return a

    def polymul_schoolbook(self, a, b):
        a = np.ascontiguousarray(a, np.uint64)
        b = np.ascontiguousarray(b, np.uint64)
        out = np.zeros_like(a)
        self.L.ref_polymul_schoolbook(len(a), _p(a), _p(b), _p(out))
        return out

    def polymul_ntt(self, a, b):
        a = np.ascontiguousarray(a, np.uint64)
        b = np.ascontiguousarray(b, np.uint64)
        out = np.zeros_like(a)
        self.L.ref_polymul_ntt(self.ctx, _p(a), _p(b), _p(out))
        return out

    def pbs(self, p, ct, table, mode):
        ps = self.ps
        ct = np.ascontiguousarray(ct, np.uint64)
        tab = np.ascontiguousarray(table, np.uint8)
        out = np.zeros(self.ct_words, np.uint64)
        ks = np.zeros(ps.n + 1, np.uint64)
        ms = np.zeros(ps.n + 1, np.uint16)
        acc = np.zeros((ps.k + 1, ps.N), np.uint64)
        self.L.ref_pbs(self.ctx, p, _p(ct), _p(tab), len(tab), mode, _p(out), _p(ks), _p(ms), _p(acc))
        return out, ks, ms, acc

    def pbs_batch(self, p, cts, tables, tlens, modes=None, want_acc=True, threads=0):
        """count independent bootstraps on all host threads: (out [count][kN+1], acc [count][k+1][N] or None)."""
        ps = self.ps
        cts = np.ascontiguousarray(cts, np.uint64)
        tables = np.ascontiguousarray(tables, np.uint8)
        tlens = np.ascontiguousarray(tlens, np.uint8)
        modes_a = None if modes is None else np.ascontiguousarray(modes, np.int32)
        count = cts.shape[0]
        out = np.zeros((count, self.ct_words), np.uint64)
        acc = np.zeros((count, ps.k + 1, ps.N), np.uint64) if want_acc else None
        self.L.ref_pbs_batch(self.ctx, p, _p(cts), _p(tables), tables.shape[1], _p(tlens), None if modes_a is None else _p(modes_a),
                             count, _p(out), None if acc is None else _p(acc), threads)
        return out, acc

    def pbs_multi(self, p, ct, tables, tlens, modes=None):
        """ONE blind rotation, several tables on the same input (multi-value bootstrap): (outs [T][kN+1], acc [k+1][N])."""
        ps = self.ps
        ct = np.ascontiguousarray(ct, np.uint64)
        tables = np.ascontiguousarray(tables, np.uint8)
        tlens = np.ascontiguousarray(tlens, np.uint8)
        modes_a = None if modes is None else np.ascontiguousarray(modes, np.int32)
        T = tables.shape[0]
        outs = np.zeros((T, self.ct_words), np.uint64)
        acc = np.zeros((ps.k + 1, ps.N), np.uint64)
        self.L.ref_pbs_multi(self.ctx, p, _p(ct), _p(tables), tables.shape[1], _p(tlens), None if modes_a is None else _p(modes_a), T, _p(outs), _p(acc))
        return outs, acc

    def eval_prog(self, program, in_bits, inst_offset=0, total=None, enc_seed=0, threads=0, multi_value=False):
        """program: tfhe_fbs_map_b200.levelize.Program (only its flat arrays are used)."""
        a = program.arrays
        d = RefProgDesc()
        d.p, d.n_inputs, d.n_lincombs, d.n_boots = program.p, program.n_inputs, program.n_lincombs, program.n_boots
        d.n_levels, d.n_slots, d.n_outputs = program.n_levels, program.n_slots, len(program.output_names)
        for name, _ in RefProgDesc._fields_[8:]:
            arr = a[name]
            ct = ctypes.c_uint8 if arr.dtype == np.uint8 else ctypes.c_int32
            setattr(d, name, arr.ctypes.data_as(ctypes.POINTER(ct)))
        in_bits = np.ascontiguousarray(in_bits, np.uint8)
        B = in_bits.shape[1]
        out = np.zeros((len(program.output_names), B), np.uint8)
        self.L.ref_eval_prog_mv(self.ctx, ctypes.byref(d), _p(in_bits), B, inst_offset, total or B, enc_seed, _p(out), threads, 1 if multi_value else 0)
        return out
