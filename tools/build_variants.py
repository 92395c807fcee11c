#!/usr/bin/env python
"""Build kernel variants of libipmcmc.so for A/B measurements on the GPU box:
    python tools/build_variants.py base: nopos:-DIPMCMC_POSPATH=0 plain:-DIPMCMC_PIPELINED=0
writes gpurun_variants/libipmcmc_<tag>.so (git-ignored; travels with the snapshot).  Select one with
IPMCMC_LIB=gpurun_variants/libipmcmc_<tag>.so."""
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
    lib = os.path.join(OUT, "libipmcmc_%s.so" % tag)
    try:
        B.compile_units(lib, extra_flags=flags, obj_dir=os.path.join(OUT, "obj_" + tag))
    except RuntimeError as e:
        return tag, 1, str(e)[-2000:]
    return tag, 0, ""


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    with ThreadPoolExecutor(2) as ex:
        for tag, rc, err in ex.map(one, sys.argv[1:]):
            print(tag, "ok" if rc == 0 else "FAILED\n" + err)
