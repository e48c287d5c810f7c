#!/usr/bin/env python
"""Top stall locations of one kernel from `ncu -i <rep> --page source --csv` (capture made with --import-source on).
    python tools/ncu_source_hotspots.py gpurun_out/x_source.csv [N] > profiles/rNN_<kernel>_hotspots.txt
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
idx = {n: i for i, n in enumerate(h)}
data = rows[hdr + 1:]
S, SRC, EX = idx["# Samples"], idx["Source"], idx["Instructions Executed"]
tot = sum(int(r[S]) for r in data)
print(f"# {rows[0][1][:110]}")
print(f"# {len(data)} SASS instructions, {sum(int(r[EX]) for r in data)} warp-level instructions executed, {tot} stall samples")
print(f"# {'idx':>5s} {'samples':>8s} {'share':>6s} {'executed':>10s}  instruction")
for i, r in sorted(enumerate(data), key=lambda t: -int(t[1][S]))[:n_top]:
    print(f"  {i:5d} {int(r[S]):8d} {int(r[S]) / max(tot, 1):6.3f} {int(r[EX]):10d}  {r[SRC].strip()[:100]}")
