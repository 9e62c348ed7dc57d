"""Build libfbs_b200.so in-tree for sm_100a:  python -m tfhe_fbs_map_b200.build [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "api.cu")
OUT = os.path.join(HERE, "libfbs_b200.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("api.cu", "kernels.cuh", "ntt.cuh", "common.cuh", "fq.cuh")] + \
       [os.path.join(HERE, "..", "include", "fbs_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    env = dict(os.environ)
    # $CC in this image points at a gcc without some spec files; nvcc finds the system g++ itself
    env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
