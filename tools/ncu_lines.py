#!/usr/bin/env python
"""Per-CUDA-source-line totals from `ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:K`:
stall samples, warp instructions executed and the dominant stall reasons of every line (inlined headers included).
    python tools/ncu_lines.py cu_src2.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        samp_i = r.index("# Samples")
        inst_i = r.index("Instructions Executed")
        stall_is = [(i, n) for i, n in enumerate(r) if n.startswith("stall_") and "Not Issued" not in n]
    elif hdr and r and r[0].strip().isdigit():
        def f(x):
            try:
                return float(x.replace(",", ""))
            except Exception:
                return 0.0
        st = sorted(((f(r[i]), n[6:]) for i, n in stall_is if i < len(r)), reverse=True)[:3]
        out.append((f(r[samp_i]), f(r[inst_i]), fname, int(r[0]), r[1].strip()[:90], " ".join(f"{n}:{int(v)}" for v, n in st if v > 0)))
tot_s, tot_i = sum(o[0] for o in out), sum(o[1] for o in out)
# stall reasons summed over the whole kernel
agg = {}
hdr2 = None
for r in rows:
    if r and r[0] == "Line No":
        hdr2 = [(i, n[6:]) for i, n in enumerate(r) if n.startswith("stall_") and "Not Issued" not in n]
    elif hdr2 and r and r[0].strip().isdigit():
        for i, n in hdr2:
            try:
                agg[n] = agg.get(n, 0.0) + float(r[i].replace(",", ""))
            except Exception:
                pass
tt = sum(agg.values()) or 1.0
print("# stall reasons: " + "  ".join(f"{n} {v / tt:.2f}" for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print(f"# total samples {int(tot_s)}, warp instructions {int(tot_i)}")
print("# by samples")
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0] / tot_s:6.3f} {int(o[0]):6d} smp {int(o[1]):9d} inst  {o[2]}:{o[3]:<4d} {o[4]}   [{o[5]}]")
print("# by instructions")
for o in sorted(out, key=lambda o: -o[1])[:top]:
    print(f"{o[1] / tot_i:6.3f} {int(o[1]):9d} inst {int(o[0]):6d} smp  {o[2]}:{o[3]:<4d} {o[4]}")
