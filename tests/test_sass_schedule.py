"""Build-time guard (no GPU needed): ptxas' schedule of the Burgers time-step loop is sensitive to
unrelated edits (register pressure at the call site of the solver decides whether the second SSPRK2
stage is interleaved or serialised: ~390 vs ~570 static stall cycles per time step, a 15-20 %
throughput difference measured on B200).  The bench kernels must keep the good schedule."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _time_step_loops(sass_path, kernel_regex):
    """(instructions, fp64 instructions, static stall sum) of every time-step loop of the kernel."""
    from sass_loops import functions
    from sass_loop import parse
    import tempfile
    found = []
    for name, body in functions(sass_path):
        if not re.search(kernel_regex, name):
            continue
        with tempfile.NamedTemporaryFile("w", suffix=".sass", delete=False) as f:
            f.write(body)
        ins = parse(f.name)
        os.unlink(f.name)
        by_addr = {x["addr"]: k for k, x in enumerate(ins)}
        for k, x in enumerate(ins):
            m = re.search(r"BRA\S*\s+(?:\w+,\s*)*(0x[0-9a-f]+)", x["text"])
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt < x["addr"] and tgt in by_addr:
                loop = ins[by_addr[tgt]:k + 1]
                f64 = sum(1 for y in loop if re.search(r"\b(DADD|DMUL|DFMA|DSETP)\b", y["text"]))
                if len(loop) < 400 and f64 >= 100:
                    found.append((len(loop), f64, sum(max(1, y["stall"]) for y in loop)))
    return found


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_bench_kernel_time_step_loop_schedule(tmp_path):
    from ip_mcmc_b200 import build
    lib = build.build()
    sass = tmp_path / "lib.sass"
    # the 1024 x 256 bench shape: CPL = 8, FUSED numerics, dynamic scheduler, 128-register build
    fun = "_ZN6ipmcmc26burgers_chain_queue_kernelILi8ELi1ELb0ELi2EEEvNS_10BurgersDevENS_10SamplerDevENS_11ChainBufDevExxi"
    with open(sass, "w") as f:
        subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], stdout=f, stderr=subprocess.DEVNULL, check=False)
    loops = _time_step_loops(str(sass), "burgers_chain_queue_kernelILi8ELi1ELb0ELi2E")
    # four rotated time-step loops per solve (burgers.cuh): the monotone-state ones (time_loop_mono: max|u| from the
    # two end cells; what a Riemann initial condition runs) for positive and for sign-changing states, and the
    # general ones (time_loop_pipelined, nested in the loop that repairs a wrong high-word guess) that KL modes
    # and blow-up solves fall back to.  Pin the hot ones: the leanest loop of each fp64 class.
    inner = {}
    for n_ins, n_f64, stalls in loops:
        key = "pos" if n_f64 < 125 else "gen"
        if key not in inner or n_ins < inner[key][0]:
            inner[key] = (n_ins, n_f64, stalls)
    assert set(inner) == {"pos", "gen"}, "expected the positive-state and the general time-step loop, found %r" % (loops,)
    pos, gen = inner["pos"], inner["gen"]
    assert pos[0] <= 160 and gen[0] <= 210, "instructions per 256-cell time step grew: %r" % (loops,)
    assert pos[1] <= 116 and gen[1] <= 134, "fp64 instructions per 256-cell time step grew: %r" % (loops,)
    assert pos[2] <= 300 and gen[2] <= 350, (
        "ptxas serialised a time-step loop (static stall sums %d / %d, limits 300 / 350): unrelated edits move "
        "its register allocation; see tools/sass_loops.py and DESIGN.md section 4.1" % (pos[2], gen[2]))
