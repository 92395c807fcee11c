#!/usr/bin/env python
"""Dump the hottest basic block (by executions) of an ncu source-page export with per-instruction
samples and the dominant stall reason.  usage: python tools/ncu_hot.py <file.src.csv.gz> [first] [last]"""
import csv, gzip, sys
rows = list(csv.reader(gzip.open(sys.argv[1], 'rt')))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r; data = []
    elif hdr and r and r[0] != "Kernel Name": data.append(r)
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
st = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
last = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
tot = sum(int(r[isamp]) for r in data)
cum = 0
for k in range(first, min(last, len(data))):
    r = data[k]
    s = int(r[isamp]); cum += s
    top = sorted(((int(r[i]), hdr[i][6:]) for i in st), reverse=True)[:2]
    print("%5d %9s %6d %5.2f%%  %-14s %-14s %s" % (k, r[iex], s, 100.0 * s / tot, "%s:%d" % (top[0][1], top[0][0]), "%s:%d" % (top[1][1], top[1][0]), r[isrc][:90]))
print("samples in range: %d of %d (%.1f%%)" % (cum, tot, 100.0 * cum / tot))
