#!/usr/bin/env python
"""Summarise ncu exports (tools/ncu_export.sh): key raw metrics, dynamic instruction mix and
stall reasons.  usage: python tools/ncu_summary.py gpurun_out/<name>   (without extension)"""
import collections
import csv
import gzip
import sys

KEEP = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__cycles_elapsed.avg', 'sm__cycles_active.avg', 'smsp__cycles_active.avg', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio']


def main(base):
    rows = list(csv.reader(open(base + ".raw.csv")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("metric,unit," + ",".join("launch%d" % i for i in range(len(data))))
    for k in ['Kernel Name'] + KEEP:
        if k in hdr:
            i = hdr.index(k)
            print(",".join([k, units[i]] + ['"%s"' % r[i] if k == 'Kernel Name' else r[i] for r in data]))
    try:
        src = list(csv.reader(gzip.open(base + ".src.csv.gz", "rt")))
    except FileNotFoundError:
        return
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], rows=[])
            blocks.append(cur)
        elif r and r[0] == "Address":
            cur['hdr'] = r
        elif cur is not None and r:
            cur['rows'].append(r)
    b = blocks[-1]
    h = b['hdr']
    ia, isrc, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    tot, stalls = collections.Counter(), collections.Counter()
    total = 0
    for r in b['rows']:
        op = r[isrc].split()
        if not op:
            continue
        o = op[1] if op[0].startswith('@') else op[0]
        o = o.split('.')[0]
        n = int(r[ia])
        total += n
        tot[o] += n
        for i in stall_cols:
            stalls[h[i]] += int(r[i])
    print("# dynamic warp-instruction mix (last profiled launch), total", total)
    for o, n in tot.most_common(14):
        print("#   %-10s %14d %5.1f%%" % (o, n, 100.0 * n / total))
    ts = sum(stalls.values())
    print("# warp stall samples")
    for s, n in stalls.most_common(8):
        print("#   %-26s %5.1f%%" % (s, 100.0 * n / ts))


if __name__ == "__main__":
    main(sys.argv[1])
