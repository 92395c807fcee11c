#!/usr/bin/env python
"""Build Lorenz-kernel variants of libipmcmc.so for A/B measurements on the GPU box: only engine.cu is recompiled
(35 s per variant), the Burgers objects of the product build are reused.
    python tools/build_lorenz_variants.py base: gsum1:-DIPMCMC_LORENZ_GSUM2=0
writes gpurun_variants/libipmcmc_<tag>.so (git-ignored; travels with the snapshot); select one with IPMCMC_LIB=..."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ip_mcmc_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_variants")


def one(spec):
    tag, _, flags = spec.partition(":")
    flags = [f for f in flags.split(",") if f]
    obj = os.path.join(OUT, "engine_%s.o" % tag)
    lib = os.path.join(OUT, "libipmcmc_%s.so" % tag)
    r = subprocess.run(["nvcc"] + B.CFLAGS + flags + ["-c", "-o", obj, "engine.cu"], cwd=B.CSRC, capture_output=True, text=True)
    if r.returncode:
        return tag, 1, r.stderr[-2000:]
    others = [o for o in glob.glob(os.path.join(B.OBJ, "burgers_cpl*.o"))]
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, obj] + others,
                       capture_output=True, text=True)
    return tag, r.returncode, r.stderr[-2000:]


if __name__ == "__main__":
    B.build()
    os.makedirs(OUT, exist_ok=True)
    with ThreadPoolExecutor(4) as ex:
        for tag, rc, err in ex.map(one, sys.argv[1:]):
            print(tag, "ok" if rc == 0 else "FAILED\n" + err)
