#!/usr/bin/env python
"""Per-launch table from `ncu -i <rep> --page raw --csv` (the --set full capture): duration, DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum), DRAM / tensor-pipe utilisation, occupancy, registers.

    python tools/ncu_summary.py gpurun_out/prof_raw.csv [--json out.json] > profiles/rNN_ncu_full_summary.txt
"""
import csv
import json
import sys


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return float("nan")


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * scale.get(unit, 1)


def to_us(v, unit):
    scale = {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}
    return v * scale.get(unit, 1)


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}

    def pick(sub):
        for n, i in col.items():
            if n.endswith(sub):
                return i
        return None
    want = {
        "dur": pick("gpu__time_duration.sum"),
        "rd": pick("dram__bytes_read.sum"),
        "wr": pick("dram__bytes_write.sum"),
        "dram_pct": pick("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_pct": pick("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "tensor_pct": pick("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
                      or pick("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
        "occ": pick("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "l2hit": pick("lts__t_sector_hit_rate.pct"),
        "regs": pick("launch__registers_per_thread"),
        "smem": pick("launch__shared_mem_per_block_dynamic"),
    }
    out = []
    print(f"{'kernel':46s} {'grid':>6s} {'dur_us':>8s} {'rd_MB':>8s} {'wr_MB':>8s} {'GB/s':>7s} {'dram%':>6s} {'sm%':>6s} {'tens%':>6s} {'occ%':>6s} {'L2hit%':>6s} {'regs':>5s}")
    for r in rows[hdr + 2:]:
        if not r or not r[0].strip().isdigit():
            continue
        g = lambda k: (num(r[want[k]]) if want[k] is not None else float("nan"))
        u = lambda k: (units[want[k]] if want[k] is not None else "")
        dur = to_us(g("dur"), u("dur"))
        rd, wr = to_bytes(g("rd"), u("rd")), to_bytes(g("wr"), u("wr"))
        name = r[col["Kernel Name"]]
        grid = r[col["Grid Size"]]
        ent = {"kernel": name, "grid": grid, "dur_us": round(dur, 2), "dram_read_bytes": rd, "dram_write_bytes": wr,
               "traffic_bytes": rd + wr, "dram_gbs": round((rd + wr) / dur / 1e3, 1) if dur else None, "dram_pct": g("dram_pct"),
               "sm_pct": g("sm_pct"), "tensor_pct": g("tensor_pct"), "occupancy_pct": g("occ"), "l2_hit_pct": g("l2hit"),
               "regs": g("regs")}
        out.append(ent)
        short = name.replace("vml::", "")[:46]
        print(f"{short:46s} {grid:>6s} {dur:8.2f} {rd / 1e6:8.2f} {wr / 1e6:8.2f} {ent['dram_gbs'] or 0:7.1f} {g('dram_pct'):6.1f} {g('sm_pct'):6.1f} "
              f"{g('tensor_pct'):6.1f} {g('occ'):6.1f} {g('l2hit'):6.1f} {g('regs'):5.0f}")
    if "--json" in sys.argv:
        # per-kernel mean traffic per launch -> bench.py's roofline.traffic
        agg = {}
        for e in out:
            # "void vml::gemm_umma_kernel<256, vml::EpiMomentOutPre>(CUtensorMap_st, ...)" -> "gemm_umma_kernel<256, EpiMomentOutPre>"
            base = e["kernel"].replace("void ", "").replace("vml::", "").replace("(int)", "")
            depth, cut = 0, len(base)
            for i, ch in enumerate(base):
                depth += ch == "<"
                depth -= ch == ">"
                if ch == "(" and depth == 0:
                    cut = i
                    break
            base = base[:cut].strip()
            a = agg.setdefault(base, {"launches": 0, "traffic_bytes": 0.0, "dur_us": 0.0})
            a["launches"] += 1
            a["traffic_bytes"] += e["traffic_bytes"]
            a["dur_us"] += e["dur_us"]
        for a in agg.values():
            a["traffic_bytes_per_launch"] = a["traffic_bytes"] / a["launches"]
            a["dur_us_per_launch"] = a["dur_us"] / a["launches"]
        doc = {"kernels": agg}
        if "--meta" in sys.argv:                      # e.g. --meta config=charadessta,queries_in_pass=192
            for kv in sys.argv[sys.argv.index("--meta") + 1].split(","):
                k, v = kv.split("=")
                doc[k] = int(v) if v.isdigit() else v
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
