#!/usr/bin/env python
"""profiles/traffic.json <- dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full captures exported by
tools/ncu_export.sh:  python tools/update_traffic.py <workload> gpurun_out/<name>.raw.csv "<what was captured>" """
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
wl, raw, what = sys.argv[1:4]
rows = list(csv.reader(open(raw)))
h, u, d = rows[0], rows[1], rows[2]
rd = float(d[h.index("dram__bytes_read.sum")]) * UNIT[u[h.index("dram__bytes_read.sum")]]
wr = float(d[h.index("dram__bytes_write.sum")]) * UNIT[u[h.index("dram__bytes_write.sum")]]
p = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(p))
t[wl] = dict(dram_bytes_read=rd, dram_bytes_write=wr, traffic=rd + wr, source=what)
json.dump(t, open(p, "w"), indent=1)
print(wl, t[wl])
