"""ctypes binding of the C ABI in include/fbs_b200.h and the ``B200Backend`` object (device + keys).

This is the only place the Python host code touches the GPU library.  There is deliberately no CPU path:
if ``libfbs_b200.so`` is missing or no CUDA device is present, construction raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

from . import params as _params
from .levelize import CProgDesc, Program, levelize

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FBS_B200_LIB") or os.path.join(_HERE, "libfbs_b200.so")   # env override: kernel-variant experiments

EXPORTS = [
    "fbs_last_error", "fbs_abi_version", "fbs_ctx_create", "fbs_keygen", "fbs_ctx_destroy", "fbs_ctx_set_cluster", "fbs_ctx_info",
    "fbs_prog_load", "fbs_prog_free", "fbs_eval_bits", "fbs_wires_bytes", "fbs_encrypt_inputs", "fbs_run_level",
    "fbs_run", "fbs_decrypt_outputs", "fbs_pbs_batch", "fbs_clear_eval", "fbs_debug_get_keys", "fbs_debug_ntt",
    "fbs_debug_pbs", "fbs_debug_pbs_multi", "fbs_debug_encrypt", "fbs_debug_decrypt", "fbs_measure_int_peak",
    "fbs_wires_alloc", "fbs_wires_free", "fbs_ipc_export", "fbs_ipc_import", "fbs_ipc_close", "fbs_set_peers", "fbs_level_sync", "fbs_sync_status",
]


class RunStats(ctypes.Structure):
    """Mirror of ``fbs_run_stats``."""
    _fields_ = [("n_pbs", ctypes.c_int64), ("n_launches", ctypes.c_int64), ("ms_total", ctypes.c_float),
                ("ms_encrypt", ctypes.c_float), ("ms_lincomb", ctypes.c_float), ("ms_keyswitch", ctypes.c_float),
                ("ms_blind_rotate", ctypes.c_float), ("ms_decrypt", ctypes.c_float), ("reserved", ctypes.c_float * 2)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_lib = None
_lib_lock = threading.Lock()


def load_library(path: str | None = None):
    """dlopen the CUDA library; raises if it has not been built (``python -m tfhe_fbs_map_b200.build``)."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise RuntimeError(f"{p} not found: build it with `python -m tfhe_fbs_map_b200.build` "
                               "(there is no CPU fallback for the encrypted executor)")
        lib = ctypes.CDLL(p)
        lib.fbs_last_error.restype = ctypes.c_char_p
        vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
        lib.fbs_ctx_create.argtypes = [ctypes.POINTER(_params.CParams), ctypes.c_int, u64, ctypes.POINTER(vp)]
        lib.fbs_keygen.argtypes = [vp]
        lib.fbs_ctx_destroy.argtypes = [vp]
        lib.fbs_ctx_set_cluster.argtypes = [vp, i32]
        lib.fbs_ctx_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i32)]
        lib.fbs_prog_load.argtypes = [vp, ctypes.POINTER(CProgDesc), ctypes.POINTER(vp)]
        lib.fbs_prog_free.argtypes = [vp]
        lib.fbs_eval_bits.argtypes = [vp, vp, vp, i64, i64, i64, u64, ctypes.c_size_t, vp, ctypes.POINTER(RunStats)]
        lib.fbs_wires_bytes.argtypes = [vp, vp, i64, ctypes.POINTER(ctypes.c_size_t)]
        lib.fbs_encrypt_inputs.argtypes = [vp, vp, vp, i64, i64, i64, u64, vp, vp]
        lib.fbs_run_level.argtypes = [vp, vp, i32, i32, i32, i64, vp, vp, ctypes.POINTER(RunStats)]
        lib.fbs_run.argtypes = [vp, vp, i64, vp, vp, ctypes.POINTER(RunStats)]
        lib.fbs_decrypt_outputs.argtypes = [vp, vp, i64, vp, vp, vp]
        lib.fbs_pbs_batch.argtypes = [vp, i32, vp, vp, vp, vp, i64, u64, vp, ctypes.POINTER(RunStats)]
        lib.fbs_clear_eval.argtypes = [vp, vp, vp, i64, vp, ctypes.POINTER(RunStats)]
        lib.fbs_debug_get_keys.argtypes = [vp, vp, vp, vp, vp]
        lib.fbs_debug_ntt.argtypes = [vp, vp, i64, i32]
        lib.fbs_debug_pbs.argtypes = [vp, i32, vp, vp, vp, vp, i64, vp, vp, vp, vp]
        lib.fbs_debug_pbs_multi.argtypes = [vp, i32, vp, vp, vp, vp, i64, i32, vp, vp]
        lib.fbs_debug_encrypt.argtypes = [vp, i32, vp, vp, i64, u64, vp]
        lib.fbs_debug_decrypt.argtypes = [vp, i32, vp, i64, vp]
        lib.fbs_measure_int_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
        lib.fbs_wires_alloc.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(vp)]
        lib.fbs_wires_free.argtypes = [vp, vp]
        lib.fbs_ipc_export.argtypes = [vp, vp, ctypes.c_char_p]
        lib.fbs_ipc_import.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]
        lib.fbs_ipc_close.argtypes = [vp, vp]
        lib.fbs_set_peers.argtypes = [vp, vp, ctypes.c_size_t, ctypes.POINTER(vp), i32, i32]
        lib.fbs_level_sync.argtypes = [vp, vp, vp]
        lib.fbs_sync_status.argtypes = [vp, ctypes.POINTER(i32)]
        for name in EXPORTS:
            if name != "fbs_last_error":
                getattr(lib, name).restype = ctypes.c_int
        if path is None:
            _lib = lib
        return lib


class FbsError(RuntimeError):
    pass


def _ptr(arr):
    return None if arr is None else ctypes.c_void_p(arr.ctypes.data)


class CompiledProgram:
    """A levelised program resident on one backend's device."""

    def __init__(self, backend, program: Program):
        self.backend = backend
        self.program = program
        self.handle = ctypes.c_void_p()
        desc = program.c_desc()
        backend._check(backend.lib.fbs_prog_load(backend.ctx, ctypes.byref(desc), ctypes.byref(self.handle)))

    # convenience passthroughs
    def __getattr__(self, item):
        return getattr(self.program, item)

    def close(self):
        if self.handle:
            self.backend.lib.fbs_prog_free(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class B200Backend:
    """Device context + TFHE keys.  ``B200Backend("A", device=0, seed=1)`` generates keys on the GPU.

    ``seed=None`` (default) draws the key seed from ``os.urandom``; pass an explicit seed for reproducible keys (tests,
    benchmarks, and multi-GPU runs where every rank must derive the SAME keys).  Randomness is a counter-based
    splitmix64 construction shared with the oracle -- a benchmark harness, not a CSPRNG.
    Every encrypting call uses a fresh encryption seed (base seed mixed with a per-backend call counter), so no
    (seed, ciphertext id) pair -- i.e. no mask/noise pair -- is ever reused across calls; ``enc_seed=`` overrides it
    for bit-exact comparisons against the oracle."""

    def __init__(self, param_set="A", device: int | None = None, seed: int | None = None, keygen: bool = True, lib_path: str | None = None):
        self.lib = load_library(lib_path)
        self.params = _params.get(param_set)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        self.seed = seed
        self.ctx = ctypes.c_void_p()
        cp = _params.to_c(self.params)
        self._check(self.lib.fbs_ctx_create(ctypes.byref(cp), device, ctypes.c_uint64(seed), ctypes.byref(self.ctx)))
        self.have_keys = False
        if keygen:
            self.keygen()
        self.enc_seed = (seed ^ 0xE2C0DE) & (2 ** 64 - 1)     # base; each encrypting call derives its own (see _fresh_enc_seed)
        self._enc_calls = 0
        self.last_enc_seed = None
        self.last_stats = None

    def _fresh_enc_seed(self, explicit=None) -> int:
        """Encryption seed of the next encrypting call: explicit, or base seed advanced by the call counter (splitmix64
        increment), so ciphertext ids restarting at 0 never meet the same seed twice on one backend."""
        if explicit is None:
            self._enc_calls += 1
            explicit = (self.enc_seed + self._enc_calls * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        self.last_enc_seed = int(explicit) & (2 ** 64 - 1)
        return self.last_enc_seed

    def _check(self, rc):
        if rc != 0:
            raise FbsError(f"fbs error {rc}: {self.lib.fbs_last_error().decode()}")

    def keygen(self):
        self._check(self.lib.fbs_keygen(self.ctx))
        self.have_keys = True

    def set_cluster(self, mode: int):
        """0 auto, 1 off, 2 / 4 / 8: force the cluster-split blind rotation of that size, 12 / 14 / 18: its prime-split twin
        (bit-identical results)."""
        self._check(self.lib.fbs_ctx_set_cluster(self.ctx, mode))

    def info(self):
        sm, bsk, ksk, sme = ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
        self._check(self.lib.fbs_ctx_info(self.ctx, ctypes.byref(sm), ctypes.byref(bsk), ctypes.byref(ksk), ctypes.byref(sme)))
        return dict(sm_count=sm.value, bsk_bytes=bsk.value, ksk_bytes=ksk.value, br_smem_bytes=sme.value)

    # ------------------------------------------------------------------ peer-mapped wire buffers (node sharding)
    def wires_alloc(self, nbytes: int) -> int:
        p = ctypes.c_void_p()
        self._check(self.lib.fbs_wires_alloc(self.ctx, nbytes, ctypes.byref(p)))
        return p.value

    def wires_free(self, ptr: int):
        self._check(self.lib.fbs_wires_free(self.ctx, ctypes.c_void_p(ptr)))

    def ipc_export(self, ptr: int) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self._check(self.lib.fbs_ipc_export(self.ctx, ctypes.c_void_p(ptr), buf))
        return buf.raw

    def ipc_import(self, handle: bytes) -> int:
        p = ctypes.c_void_p()
        self._check(self.lib.fbs_ipc_import(self.ctx, handle, ctypes.byref(p)))
        return p.value

    def ipc_close(self, ptr: int):
        self._check(self.lib.fbs_ipc_close(self.ctx, ctypes.c_void_p(ptr)))

    def set_peers(self, local_ptr, nbytes, ptrs, rank):
        """Bind the peer replicas (other ranks' IPC-mapped buffers) to the local wire buffer ``local_ptr``; [] unbinds."""
        arr = (ctypes.c_void_p * max(1, len(ptrs)))(*ptrs)
        self._check(self.lib.fbs_set_peers(self.ctx, ctypes.c_void_p(local_ptr or 0), nbytes, arr, len(ptrs), rank))

    def run_level_sync(self, local_ptr, stream=0):
        self._check(self.lib.fbs_level_sync(self.ctx, ctypes.c_void_p(local_ptr), ctypes.c_void_p(stream)))

    def sync_status(self) -> int:
        """0, or 1 + rank of a peer whose level flag never arrived (device-side level hand-off timed out)."""
        v = ctypes.c_int32()
        self._check(self.lib.fbs_sync_status(self.ctx, ctypes.byref(v)))
        return v.value

    def measure_int_peak(self) -> float:
        """Sustained IMAD.WIDE rate of this device in 32x32->64 multiplies per second."""
        v = ctypes.c_double()
        self._check(self.lib.fbs_measure_int_peak(self.ctx, ctypes.byref(v)))
        return v.value

    def close(self):
        if self.ctx:
            self.lib.fbs_ctx_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ programs
    def compile(self, env, fbs_size=None, clear=False, shard_pad=1, reuse_slots=True, multi_value=False) -> CompiledProgram:
        key = (id(self), fbs_size, clear, shard_pad, reuse_slots, multi_value, len(env.instructions), len(env.outputs))
        cached = getattr(env, "_compiled", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        prog = levelize(env, fbs_size, reuse_slots=reuse_slots, shard_pad=shard_pad, clear=clear, multi_value=multi_value)
        cp = CompiledProgram(self, prog)
        try:
            env._compiled = (key, cp)
        except Exception:
            pass
        return cp

    def load(self, program: Program) -> CompiledProgram:
        return CompiledProgram(self, program)

    # ------------------------------------------------------------------ evaluation with host buffers
    def eval_bits(self, cprog: CompiledProgram, in_bits: np.ndarray, inst_offset=0, total=None, max_wire_bytes=0,
                  out: np.ndarray | None = None, in_ptr=None, out_ptr=None, B=None, enc_seed=None) -> np.ndarray:
        """in_bits: uint8 [n_inputs][B] -> uint8 [n_outputs][B] (values mod 2p).  Encrypt, run every level as
        one batched bootstrap, decrypt.  ``in_ptr``/``out_ptr`` allow pinned host buffers (bench.py)."""
        if not self.have_keys:
            raise FbsError("keys not generated")
        if in_ptr is None:
            in_bits = np.ascontiguousarray(in_bits, dtype=np.uint8)
            B = in_bits.shape[1] if in_bits.ndim == 2 else (B or 1)
            in_ptr = in_bits.ctypes.data
        n_out = len(cprog.program.output_names)
        if out_ptr is None:
            if out is None:
                out = np.zeros((n_out, B), dtype=np.uint8)
            out_ptr = out.ctypes.data
        st = RunStats()
        self._check(self.lib.fbs_eval_bits(self.ctx, cprog.handle, ctypes.c_void_p(in_ptr), B, inst_offset, total or B,
                                           ctypes.c_uint64(self._fresh_enc_seed(enc_seed)), max_wire_bytes, ctypes.c_void_p(out_ptr), ctypes.byref(st)))
        self.last_stats = st.as_dict()
        return out

    def eval_clear(self, cprog: CompiledProgram, in_bits: np.ndarray) -> np.ndarray:
        in_bits = np.ascontiguousarray(in_bits, dtype=np.uint8)
        B = in_bits.shape[1]
        out = np.zeros((len(cprog.program.output_names), B), dtype=np.uint8)
        st = RunStats()
        self._check(self.lib.fbs_clear_eval(self.ctx, cprog.handle, _ptr(in_bits), B, _ptr(out), ctypes.byref(st)))
        self.last_stats = st.as_dict()
        return out

    def pbs_batch(self, p, msgs, tables, tlens, modes=None, enc_seed=None):
        """BASELINE config 5: independent bootstraps.  tables: uint8 [count][2p]."""
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
        tables = np.ascontiguousarray(tables, dtype=np.uint8)
        tlens = np.ascontiguousarray(tlens, dtype=np.uint8)
        count = len(msgs)
        assert tables.shape == (count, 2 * p)
        modes_a = None if modes is None else np.ascontiguousarray(modes, dtype=np.int32)
        out = np.zeros(count, dtype=np.uint8)
        st = RunStats()
        self._check(self.lib.fbs_pbs_batch(self.ctx, p, _ptr(msgs), _ptr(tables), _ptr(tlens), _ptr(modes_a), count,
                                           ctypes.c_uint64(self._fresh_enc_seed(enc_seed)), _ptr(out), ctypes.byref(st)))
        self.last_stats = st.as_dict()
        return out

    # ------------------------------------------------------------------ device-resident split form
    def wires_bytes(self, cprog, B):
        n = ctypes.c_size_t()
        self._check(self.lib.fbs_wires_bytes(self.ctx, cprog.handle, B, ctypes.byref(n)))
        return n.value

    def encrypt_inputs(self, cprog, in_dev_ptr, B, wires_ptr, stream=0, inst_offset=0, total=None, enc_seed=None):
        self._check(self.lib.fbs_encrypt_inputs(self.ctx, cprog.handle, ctypes.c_void_p(in_dev_ptr), B, inst_offset, total or B,
                                                ctypes.c_uint64(self._fresh_enc_seed(enc_seed)), ctypes.c_void_p(wires_ptr), ctypes.c_void_p(stream)))

    def run_level(self, cprog, level, B, wires_ptr, node_begin=-1, node_end=-1, stream=0, stats=None):
        self._check(self.lib.fbs_run_level(self.ctx, cprog.handle, level, node_begin, node_end, B, ctypes.c_void_p(wires_ptr),
                                           ctypes.c_void_p(stream), None if stats is None else ctypes.byref(stats)))

    def run(self, cprog, B, wires_ptr, stream=0, stats=None):
        self._check(self.lib.fbs_run(self.ctx, cprog.handle, B, ctypes.c_void_p(wires_ptr), ctypes.c_void_p(stream),
                                     None if stats is None else ctypes.byref(stats)))

    def decrypt_outputs(self, cprog, B, wires_ptr, out_dev_ptr, stream=0):
        self._check(self.lib.fbs_decrypt_outputs(self.ctx, cprog.handle, B, ctypes.c_void_p(wires_ptr), ctypes.c_void_p(out_dev_ptr),
                                                 ctypes.c_void_p(stream)))

    # ------------------------------------------------------------------ parity taps (tests only)
    def debug_keys(self, want_bsk=True):
        P = self.params
        s_lwe = np.zeros(P.n, np.uint8)
        s_big = np.zeros(P.k * P.N, np.uint8)
        ksk = np.zeros((P.k * P.N * P.ks_l, P.n + 1), np.uint64)
        bsk = np.zeros((P.n_ggsw, (P.k + 1) * P.bsk_l, P.k + 1, P.N), np.uint64) if want_bsk else None
        self._check(self.lib.fbs_debug_get_keys(self.ctx, _ptr(s_lwe), _ptr(s_big), _ptr(ksk), _ptr(bsk)))
        return s_lwe, s_big, ksk, bsk

    def debug_ntt(self, polys, inverse=False):
        a = np.ascontiguousarray(polys, dtype=np.uint64).copy()
        self._check(self.lib.fbs_debug_ntt(self.ctx, _ptr(a), a.size // self.params.N, 1 if inverse else 0))
        return a

    def debug_encrypt(self, p, msgs, ct_ids=None, enc_seed=None):
        msgs = np.ascontiguousarray(msgs, dtype=np.int32)
        ids = None if ct_ids is None else np.ascontiguousarray(ct_ids, dtype=np.uint64)
        out = np.zeros((len(msgs), self.params.ct_words), np.uint64)
        self._check(self.lib.fbs_debug_encrypt(self.ctx, p, _ptr(msgs), _ptr(ids), len(msgs),
                                               ctypes.c_uint64(self.enc_seed if enc_seed is None else enc_seed), _ptr(out)))
        return out

    def debug_decrypt(self, p, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint64)
        out = np.zeros(cts.shape[0], np.int32)
        self._check(self.lib.fbs_debug_decrypt(self.ctx, p, _ptr(cts), cts.shape[0], _ptr(out)))
        return out

    def debug_pbs(self, p, in_cts, tables, tlens, modes=None):
        P = self.params
        in_cts = np.ascontiguousarray(in_cts, dtype=np.uint64)
        count = in_cts.shape[0]
        tables = np.ascontiguousarray(tables, dtype=np.uint8)
        tlens = np.ascontiguousarray(tlens, dtype=np.uint8)
        modes_a = None if modes is None else np.ascontiguousarray(modes, dtype=np.int32)
        out = np.zeros((count, P.ct_words), np.uint64)
        ks = np.zeros((count, P.n + 1), np.uint64)
        ms = np.zeros((count, P.n + 1), np.uint16)
        acc = np.zeros((count, P.k + 1, P.N), np.uint64)
        self._check(self.lib.fbs_debug_pbs(self.ctx, p, _ptr(in_cts), _ptr(tables), _ptr(tlens), _ptr(modes_a), count,
                                           _ptr(out), _ptr(ks), _ptr(ms), _ptr(acc)))
        return out, ks, ms, acc


    def debug_pbs_multi(self, p, in_cts, tables, tlens, modes=None):
        """Multi-value bootstrap tap: ``count`` input ciphertexts, T tables EACH (tables [count][T][2p], tlens / modes [count][T]),
        one blind rotation per input: (out [count][T][kN+1], acc [count][k+1][N] = accumulator before the per-table products)."""
        P = self.params
        in_cts = np.ascontiguousarray(in_cts, dtype=np.uint64)
        tables = np.ascontiguousarray(tables, dtype=np.uint8)
        count, T = tables.shape[0], tables.shape[1]
        assert tables.shape[2] == 2 * p and in_cts.shape[0] == count
        tlens = np.ascontiguousarray(tlens, dtype=np.uint8).reshape(count, T)
        modes_a = None if modes is None else np.ascontiguousarray(modes, dtype=np.int32).reshape(count, T)
        out = np.zeros((count, T, P.ct_words), np.uint64)
        acc = np.zeros((count, P.k + 1, P.N), np.uint64)
        self._check(self.lib.fbs_debug_pbs_multi(self.ctx, p, _ptr(in_cts), _ptr(tables), _ptr(tlens), _ptr(modes_a), count, T, _ptr(out), _ptr(acc)))
        return out, acc


_default = {}


def default_backend(need_keys: bool = True, param_set: str | None = None) -> B200Backend:
    """Process-wide backend used by ``LutExecEnv.eval`` when none is passed (keys are generated lazily)."""
    name = param_set or os.environ.get("FBS_PARAM_SET", _params.DEFAULT_SET)
    be = _default.get(name)
    if be is None:
        be = B200Backend(name, keygen=need_keys)
        _default[name] = be
    elif need_keys and not be.have_keys:
        be.keygen()
    return be
