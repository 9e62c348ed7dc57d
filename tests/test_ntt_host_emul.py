"""Runs the device NTT index/twiddle/swizzle logic (csrc/ntt.cuh) on the CPU by emulating threads with loops."""
import os
import subprocess

from conftest import ROOT


def test_ntt_pass_logic_on_host(tmp_path):
    exe = tmp_path / "ntt_emul"
    src = os.path.join(ROOT, "tests", "host", "ntt_emul.cpp")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-o", str(exe), src], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
