#!/usr/bin/env python
"""Find the loops (backward branches) of every function in a cuobjdump -sass dump and report,
per loop that contains fp64 work: instruction count, fp64 count, sum of static stall counts.
usage: python tools/sass_loops.py dump.sass [min_instructions]"""
import re
import sys

sys.path.insert(0, __import__("os").path.dirname(__file__))
from sass_loop import parse  # noqa: E402


def functions(path):
    txt = open(path).read()
    parts = re.split(r"\n\s*Function : (\S+)\n", txt)
    for i in range(1, len(parts), 2):
        yield parts[i], parts[i + 1]


def main(path, min_ins=60):
    import tempfile, os
    for name, body in functions(path):
        with tempfile.NamedTemporaryFile("w", suffix=".sass", delete=False) as f:
            f.write(body)
        ins = parse(f.name)
        os.unlink(f.name)
        by_addr = {x["addr"]: k for k, x in enumerate(ins)}
        print("== %s: %d instructions" % (name[:90], len(ins)))
        for k, x in enumerate(ins):
            m = re.search(r"BRA\S*\s+(?:\w+,\s*)*(0x[0-9a-f]+)", x["text"])
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt < x["addr"] and tgt in by_addr:
                body_ins = ins[by_addr[tgt]:k + 1]
                if len(body_ins) < min_ins:
                    continue
                f64 = sum(1 for y in body_ins if re.search(r"\b(DADD|DMUL|DFMA|DSETP)\b", y["text"]))
                st = sum(max(1, y["stall"]) for y in body_ins)
                big = sum(1 for y in body_ins if y["stall"] >= 5)
                print("   loop %05x..%05x: %4d instr, %3d fp64, stall sum %4d, %2d instr with stall>=5"
                      % (tgt, x["addr"], len(body_ins), f64, st, big))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
