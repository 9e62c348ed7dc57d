"""The C-ABI library: builds for sm_100a, loads, and exports every symbol include/fbs_b200.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from tfhe_fbs_map_b200 import backend, build

HEADER = os.path.join(ROOT, "include", "fbs_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fbs_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    return build.build()


def test_header_and_binding_agree():
    assert set(declared_functions()) == set(backend.EXPORTS)


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for fn in declared_functions():
        assert hasattr(lib, fn), fn
    lib.fbs_abi_version.restype = ctypes.c_int
    assert lib.fbs_abi_version() == 1


def test_library_is_sm100a_with_tma(libpath):
    out = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_Z14k_blind_rotateILi11ELi1ELi1ELb1ELi2ELi2EEv6BRArgs", libpath],
                          capture_output=True, text=True).stdout
    assert "UBLKCP" in sass          # cp.async.bulk = TMA bulk copy streams the bootstrapping key
    assert "IMAD.WIDE" in sass       # the arithmetic runs on the integer pipes
    # the default (key-unrolled) kernel: TMA key ring with mbarriers, no global load inside the step loop's transforms
    sass2 = subprocess.run(["cuobjdump", "-sass", "-fun", "_Z15k_blind_rotate2ILi11ELi1ELi2ELi2ELi2EEv6BRArgs", libpath],
                           capture_output=True, text=True).stdout
    assert "UBLKCP" in sass2 and "SYNCS.ARRIVE" in sass2 and "IMAD.HI.U32" in sass2
    # the DEFAULT instantiation (set A3: three key bits per step, two bootstraps per CTA) and its one-bootstrap tail variant
    for fun in ("_Z15k_blind_rotate2ILi11ELi1ELi2ELi2ELi3EEv6BRArgs", "_Z15k_blind_rotate2ILi11ELi1ELi1ELi1ELi3EEv6BRArgs"):
        s3 = subprocess.run(["cuobjdump", "-sass", "-fun", fun, libpath], capture_output=True, text=True).stdout
        assert "UBLKCP" in s3 and "SYNCS.ARRIVE" in s3 and "IMAD.WIDE.U32" in s3 and "IMAD.HI.U32" in s3, fun
    # the cluster-split low-latency kernel: remote shared-memory stores with mbarrier completion (st.async), cluster barrier, TMA key ring
    s4 = subprocess.run(["cuobjdump", "-sass", "-fun", "_Z17k_blind_rotate_clILi11ELi1ELi3ELi2EEv6BRArgs", libpath], capture_output=True, text=True).stdout
    assert "UBLKCP" in s4 and "UCGABAR" in s4 and ("STAS" in s4 or "ST.ASYNC" in s4 or "STS.ASYNC" in s4), "cluster kernel: no async remote stores in SASS"


def test_parameter_validation_needs_no_gpu(libpath):
    """Shape checks of fbs_ctx_create come before any CUDA call: they are testable without a device."""
    from tfhe_fbs_map_b200 import params
    lib = backend.load_library(libpath)
    out = ctypes.c_void_p()
    for bad, msg in ((dict(bsk_unroll=4), b"bsk_unroll"),
                     (dict(bsk_unroll=2, bsk_l=2, bsk_beta=12), b"bsk_l = 1"), (dict(bsk_unroll=3, bsk_l=2, bsk_beta=12), b"bsk_l = 1"), (dict(N=1000), b"power of two"),
                     (dict(bsk_l=1, bsk_beta=26), b"bsk_beta")):
        d = params.get("A").as_dict(); d.update(bad); d["name"] = "bad"
        cp = params.to_c(params.ParamSet(**d))
        assert lib.fbs_ctx_create(ctypes.byref(cp), 0, 1, ctypes.byref(out)) == -1, bad
        assert msg in lib.fbs_last_error(), (bad, lib.fbs_last_error())


def test_errors_are_codes_not_crashes(libpath):
    lib = backend.load_library(libpath)
    rc = lib.fbs_ctx_create(None, 0, 0, None)
    assert rc == -1 and b"null" in lib.fbs_last_error()
    assert lib.fbs_keygen(None) == -1
