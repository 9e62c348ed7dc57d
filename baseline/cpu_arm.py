"""ctypes wrapper of baseline/libcpu_pbs.so: the tuned CPU implementation of the encrypted evaluation that bench.py times
next to the GPU (`--impl reference`, `cpu_baseline`).  Built on the box it runs on (-march=native).  Keys come from the
oracle's seeded key generation (oracle/tfhe_ref.py) -- key generation is not part of any timed region."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libcpu_pbs.so")


def _host_tag():
    """Identity of this host's CPU: the library is compiled with -march=native, so a copy built elsewhere (it travels with the
    repository snapshot to the GPU box) must be rebuilt when the CPU differs."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        model = next((ln.split(":", 1)[1].strip() for ln in txt.splitlines() if ln.startswith("model name")), "")
        flags = next((ln.split(":", 1)[1].strip() for ln in txt.splitlines() if ln.startswith("flags")), "")
        return model + " " + hashlib.sha1(flags.encode()).hexdigest()[:12]
    except OSError:
        return "unknown"


def build(force=False):
    src = os.path.join(_HERE, "cpu_pbs.cpp")
    tag_file = LIB + ".host"
    tag = _host_tag()
    try:
        built_for = open(tag_file).read()
    except OSError:
        built_for = None
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src) or built_for != tag:
        env = dict(os.environ)
        env.pop("CC", None); env.pop("CXX", None)
        r = subprocess.run(["make", "-C", _HERE, "-B", "libcpu_pbs.so"], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("CPU arm build failed:\n" + r.stdout + r.stderr)
        with open(tag_file, "w") as f:
            f.write(tag)
    return LIB


class CpuParams(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("k", ctypes.c_int32), ("N", ctypes.c_int32), ("bsk_l", ctypes.c_int32),
                ("bsk_beta", ctypes.c_int32), ("ks_l", ctypes.c_int32), ("ks_beta", ctypes.c_int32),
                ("bsk_unroll", ctypes.c_int32), ("lwe_noise", ctypes.c_uint64), ("glwe_noise", ctypes.c_uint64)]


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
        L.cpu_ctx_create.argtypes = [ctypes.POINTER(CpuParams), vp, vp, vp]
        L.cpu_ctx_create.restype = vp
        L.cpu_ctx_destroy.argtypes = [vp]
        L.cpu_pbs_batch.argtypes = [vp, i32, vp, vp, i32, vp, vp, i64, vp, i32]
        L.cpu_eval_prog.argtypes = [vp, vp, vp, i64, i64, i64, u64, vp, i32, i32]
        L.cpu_eval_prog.restype = ctypes.c_int
        L.cpu_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


class CpuTFHE:
    """Tuned CPU context for a key-unrolled one-level parameter set (A2 / A3 and their toy twins), keys from ``ref``
    (an ``oracle.tfhe_ref.RefTFHE`` with the same parameters and seed)."""

    def __init__(self, ps, ref):
        self.L, self.ps = lib(), ps
        _, s_big, ksk, bsk = ref.keys()
        cp = CpuParams(ps.n, ps.k, ps.N, ps.bsk_l, ps.bsk_beta, ps.ks_l, ps.ks_beta, ps.bsk_unroll, ps.lwe_noise_scale, ps.glwe_noise_scale)
        self.ctx = self.L.cpu_ctx_create(ctypes.byref(cp), _p(s_big), _p(ksk), _p(bsk))
        if not self.ctx:
            raise ValueError("the CPU arm handles k = 1, one decomposition level, bsk_unroll 2 or 3")
        self.ct_words = ps.k * ps.N + 1

    def __del__(self):
        try:
            self.L.cpu_ctx_destroy(self.ctx)
        except Exception:
            pass

    def pbs_batch(self, p, cts, tables, tlens, modes=None, threads=0):
        cts = np.ascontiguousarray(cts, np.uint64)
        tables = np.ascontiguousarray(tables, np.uint8)
        tlens = np.ascontiguousarray(tlens, np.uint8)
        modes_a = None if modes is None else np.ascontiguousarray(modes, np.int32)
        out = np.zeros((cts.shape[0], self.ct_words), np.uint64)
        self.L.cpu_pbs_batch(self.ctx, p, _p(cts), _p(tables), tables.shape[1], _p(tlens), _p(modes_a), cts.shape[0], _p(out), threads)
        return out

    def eval_prog(self, program, in_bits, inst_offset=0, total=None, enc_seed=0, threads=0, max_levels=0):
        """Encrypt, evaluate every level (or only the first ``max_levels``: then no outputs), decrypt.  ``self.last_pbs`` =
        bootstraps executed per instance."""
        desc = program.c_desc()                  # identical field layout to cpu_prog_desc
        in_bits = np.ascontiguousarray(in_bits, np.uint8)
        B = in_bits.shape[1]
        out = np.zeros((len(program.output_names), B), np.uint8)
        self.last_pbs = self.L.cpu_eval_prog(self.ctx, ctypes.byref(desc), _p(in_bits), B, inst_offset, total or B, enc_seed, _p(out), threads, max_levels)
        return out
