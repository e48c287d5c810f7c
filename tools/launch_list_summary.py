#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <cmd>` launch list.

    python tools/launch_list_summary.py gpurun_out/launches.csv "<header line>" > profiles/rNN_launches_summary.txt
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ci = {n: i for i, n in enumerate(rows[hdr])}
agg, tot = collections.OrderedDict(), 0.0
for r in rows[hdr + 1:]:
    if not r or not r[0].isdigit() or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ci["Metric Value"]].replace(",", ""))
    us = v / 1e3 if r[ci["Metric Unit"]] in ("ns", "nsecond") else v
    a = agg.setdefault(r[ci["Kernel Name"]], [0, 0.0])
    a[0] += 1
    a[1] += us
    tot += us
for line in sys.argv[2:]:
    print(line)
vml = 0.0
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={n:4d} avg={us / n:8.2f}us share={us / tot:.3f}")
    if "vml::" in k:
        vml += us
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches; vml:: kernels {vml / tot:.3f} of the time")
