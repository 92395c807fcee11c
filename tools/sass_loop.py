#!/usr/bin/env python
"""Decode a SASS address range of a cuobjdump -sass dump with the scheduling control fields
(stall count, yield, write/read barrier, wait mask) of each 128-bit sm_100 instruction.
usage: python tools/sass_loop.py dump.sass 0x2a50 0x39e0"""
import re
import sys
import collections


def parse(path):
    out = []
    lines = open(path).read().split("\n")
    i = 0
    pat = re.compile(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
    pat2 = re.compile(r"/\* (0x[0-9a-f]{16}) \*/")
    while i < len(lines):
        m = pat.search(lines[i])
        if m and i + 1 < len(lines):
            m2 = pat2.search(lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ctrl = hi >> 41
                out.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=ctrl & 0xf, yld=(ctrl >> 4) & 1,
                                wb=(ctrl >> 5) & 7, rb=(ctrl >> 8) & 7, wait=(ctrl >> 11) & 0x3f))
                i += 2
                continue
        i += 1
    return out


if __name__ == "__main__":
    ins = parse(sys.argv[1])
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
    sel = [x for x in ins if lo <= x["addr"] <= hi]
    cyc = 0
    mix = collections.Counter()
    for x in sel:
        op = x["text"].split()
        o = op[1] if op[0].startswith("@") else op[0]
        mix[o.split(".")[0]] += 1
        print("%05x  c%5d  st%2d %s wb%d rb%d wait%02x   %s" % (x["addr"], cyc, x["stall"], "Y" if x["yld"] else " ",
                                                            x["wb"], x["rb"], x["wait"], x["text"]))
        cyc += max(1, x["stall"])
    print("# %d instructions, sum of stall counts %d" % (len(sel), cyc))
    print("# mix:", dict(mix.most_common()))
