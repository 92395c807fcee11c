#!/usr/bin/env python
"""Group the SASS of an ncu source-page export (tools/ncu_export.sh) into runs of instructions with
the same execution count and print, per run: #instructions, executions, stall samples, avg active
threads, the fp64 share and the first/last opcode.  Shows where a latency-bound kernel spends time.
usage: python tools/ncu_regions.py gpurun_out/<name>.src.csv.gz [min_samples]"""
import csv
import gzip
import sys


def main(path, min_samples=0):
    rows = list(csv.reader(gzip.open(path, "rt")))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], rows=[])
            blocks.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    b = blocks[-1]
    h = b["hdr"]
    ia, isrc, isamp, ithr = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples"), h.index("Avg. Threads Executed")
    tot_s = sum(int(r[isamp]) for r in b["rows"])
    tot_i = sum(int(r[ia]) for r in b["rows"])
    print("kernel:", b["name"][:100])
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    runs = []
    for k, r in enumerate(b["rows"]):
        n = int(r[ia])
        if runs and abs(runs[-1]["n"] - n) <= 0.002 * max(n, 1):
            runs[-1]["rows"].append(r)
        else:
            runs.append(dict(n=n, rows=[r], first=k))
    print("%6s %5s %12s %8s %6s %5s %5s  %s" % ("idx", "ninst", "exec", "samples", "%smp", "thr", "fp64", "first .. last"))
    for run in runs:
        s = sum(int(r[isamp]) for r in run["rows"])
        if s < min_samples:
            continue
        ops = [(r[isrc].split()[1] if r[isrc].split()[0].startswith("@") else r[isrc].split()[0]) for r in run["rows"]]
        f64 = sum(1 for o in ops if o.split(".")[0] in ("DADD", "DMUL", "DFMA", "DSETP"))
        thr = sum(float(r[ithr]) for r in run["rows"]) / len(run["rows"])
        print("%6d %5d %12d %8d %5.1f%% %5.1f %5d  %s .. %s" % (run["first"], len(ops), run["n"], s, 100.0 * s / tot_s, thr, f64,
                                                           ops[0], ops[-1]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
