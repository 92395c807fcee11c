"""Builds ip_mcmc_b200/libipmcmc.so (hand-written sm_100a CUDA + the C ABI of include/ipmcmc.h).

nvcc cross-compiles without a GPU.  The library is built IN-TREE so it travels to the GPU box
with the repo snapshot; it is git-ignored.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libipmcmc.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "philox.cuh", "burgers.cuh", "burgers_kernels.cuh", "burgers_team.cuh", "lorenz.cuh",
           "lorenz_kernels.cuh", "sampler.cuh", os.path.join("..", "..", "include", "ipmcmc.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
              "-std=c++17", "--extended-lambda", "--split-compile=0", "-shared", "-Xcompiler", "-fPIC"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources. Returns the path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
