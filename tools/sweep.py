#!/usr/bin/env python
"""BASELINE.json configs[4]: span-map scaling sweep, kernel-level roofline report (GPU box only).

    python tools/sweep.py [--out gpurun_out/sweep.json] [--quick]

For map side N in {16,32,64,128}, batch B in {32..1024}, window ratio r = T/N in {4 (regular), 2
(ActivityNet-style irregular)}: times, with CUDA events on the launching stream and L2 flushed
before every launch, the fused span-pool/fusion kernel (HBM roofline, SURVEY.md 8d bytes) and the
moment-unit GEMM = the "map-conv stack" (tensor roofline, 4*V*D^2 flops), full-length videos
(V = N(N+1)/2 valid cells per query).  Peaks: MEASURED_PEAKS.json.

``--stages`` additionally runs the whole drop-in forward at every point (random-init SMIN of that map size, d0 = 64) and times
the two stages that dominate it: the content unit (one persistent tcgen05 kernel per layer; HBM roofline on the bytes of
DESIGN.md section 4) and the boundary unit (gate + rows + stream kernels), by re-issuing the recorded launches of the first
SMI layer.  For ncu counters per point run the same command under
``ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum,sm__inst_executed.avg.per_cycle_elapsed
-k regex:'content_unit_kernel|boundary' --clock-control none`` with ``--quick`` (tools/ncu_sweep_summary.py turns the CSV into a table).
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vml_b200  # noqa
from vml_b200 import lib as L_
from vml_b200.lib import Dims, call, ptr, stream_ptr
from vml_b200.smin import Workspace, make_cells

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


CLEAN = False   # --clean-flush: after the 256 MiB memset, read a second 256 MiB buffer so that L2 ends up full of CLEAN lines


def cold_time(fn, flush, reps=8):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        if CLEAN:
            _flush_read.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3   # us


def stage_times(T, N, B, n, flush, hbm):
    """Content unit and boundary unit of the first SMI layer inside a full forward of a random-init model of this map size."""
    from vml_b200 import synth
    from vml_b200.configs import SminConfig
    from vml_b200.smin import SMIN
    D, C, layers = 512, 4, 3
    cfg = SminConfig("sweep", T, N, C, D, 128, layers, 64, 13, 256)
    torch.manual_seed(7)
    model = SMIN(*cfg.ctor_args(), device=torch.device("cuda"), precision="bf16").cuda().eval()
    parts = [synth.make_batch(cfg, min(64, B - lo), 50 + lo, full_length=True) for lo in range(0, B, 64)]
    b = {k: torch.cat([p[k] for p in parts]).cuda() for k in synth.MODEL_INPUT_KEYS}
    model(*[b[k] for k in synth.MODEL_INPUT_KEYS], overlap=False)           # buffers of this shape exist
    torch.cuda.synchronize()
    rec, marks = [], []
    L_.set_recorder(rec)
    model(*[b[k] for k in synth.MODEL_INPUT_KEYS], mark=lambda name: marks.append((name, len(rec))))
    L_.set_recorder(None)
    torch.cuda.synchronize()
    out, lo, seen = {}, 0, set()
    for name, hi in marks:
        calls, lo = rec[lo:hi], hi
        if name not in ("content_unit", "boundary_unit") or name in seen or not calls:
            continue
        seen.add(name)                                                          # first SMI layer only
        us = cold_time(lambda: [L_.call(fn, *a[:-1], stream_ptr()) for fn, a in calls], flush)
        if name == "content_unit":
            bytes_ = 2 * (2 * n * C * D + 2 * n * D)          # first layer: fc in, cu out, fbar in, mean_c cu out (bf16)
            out.update(content_unit_us=round(us, 1), content_unit_GBs=round(bytes_ / us / 1e3, 1),
                       content_unit_frac=round(bytes_ / us / 1e3 / hbm, 4))
        else:
            bytes_ = 2 * n * D + 4 * 3 * B * N * D
            out.update(boundary_unit_us=round(us, 1), boundary_unit_GBs=round(bytes_ / us / 1e3, 1),
                       boundary_unit_frac=round(bytes_ / us / 1e3 / hbm, 4))
    del model, b
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--stages", action="store_true", help="also time the content unit and the boundary unit of a full forward per point")
    ap.add_argument("--clean-flush", action="store_true",
                    help="memset flush followed by a 256 MiB read: the timed kernel then does not pay the write-back of the flush's dirty lines")
    args = ap.parse_args()
    L_.load()
    dev = torch.device("cuda")
    hbm, tf, src = peaks()
    flush = torch.empty(256 * 2**20, device=dev, dtype=torch.uint8)
    global CLEAN, _flush_read
    CLEAN = args.clean_flush
    _flush_read = torch.zeros(64 * 2**20, device=dev, dtype=torch.float32)
    D, C = 512, 4
    rows = []
    Ns = (16, 64) if args.quick else (16, 32, 64, 128)
    Bs = (64, 1024) if args.quick else (32, 64, 128, 256, 512, 1024)
    for N in Ns:
        for r in (4, 2):
            T = r * N
            for B in Bs:
                V = N * (N + 1) // 2
                n = B * V
                if n * C * D * 2 > 40e9:          # keep fc under 40 GB
                    continue
                dims = Dims(T, N, C, D, 128, 3, 1024, 13, 256)
                ws = Workspace(dev)
                cells = make_cells(ws, B, N)
                mmask = torch.ones(N, N, dtype=torch.bool, device=dev).triu().unsqueeze(0).expand(B, N, N).contiguous().view(torch.uint8)
                st = stream_ptr()
                call("vml_build_cells", ptr(mmask), B, N, cells, st)
                fv = (torch.randn(B * T, D, device=dev) * 0.5).to(torch.bfloat16)
                fs = torch.randn(B, D, device=dev)
                fc = torch.empty(n, C, D, device=dev, dtype=torch.bfloat16)
                fm = torch.empty(n, D, device=dev, dtype=torch.bfloat16)
                fb = torch.empty(B, N, D, device=dev)
                us = cold_time(lambda: call("vml_span_pool_fuse", ptr(fv), ptr(fs), cells, ptr(fc), ptr(fm), ptr(fb), B, dims, L_.BF16, st), flush)
                bytes_ = 2 * (B * T * D + n * C * D + n * D) + 4 * (B * D + B * N * D)
                gbs = bytes_ / us / 1e3
                row = {"N": N, "r": r, "T": T, "B": B, "cells": n, "span_pool_us": round(us, 1), "span_pool_GBs": round(gbs, 1),
                       "span_pool_frac": round(gbs / hbm, 4)}
                del fc
                # moment-unit GEMM: mu = [bu_i*bu_j | mean_c cu] . [Wfb|Wfc]^T + b + fm
                op = (torch.randn(n, 2 * D, device=dev) * 0.1).to(torch.bfloat16)
                W = (torch.randn(D, 2 * D, device=dev) * 0.03).to(torch.bfloat16)
                bias = torch.randn(D, device=dev)
                mu = torch.empty(n, D, device=dev, dtype=torch.bfloat16)
                us2 = cold_time(lambda: call("vml_moment_out", ptr(op), ptr(W), ptr(bias), ptr(fm), cells, ptr(mu), dims, L_.BF16, st), flush)
                tfl = 4.0 * n * D * D / us2 / 1e6
                row.update(moment_gemm_us=round(us2, 1), moment_gemm_TFs=round(tfl, 1), moment_gemm_frac=round(tfl / tf, 4))
                del op, mu, fm, fv
                torch.cuda.empty_cache()
                if args.stages and n * C * D * 2 * 2.5 < 60e9:
                    row.update(stage_times(T, N, B, n, flush, hbm))
                rows.append(row)
                print(json.dumps(row), flush=True)
                torch.cuda.empty_cache()
    out = {"peaks": {"hbm_gbs": hbm, "bf16_tflops": tf, "source": src}, "timing": "CUDA events, L2 flushed (256 MiB memset" + (" + 256 MiB read: clean lines" if CLEAN else "") + ") before every launch, mean of 8",
           "rows": rows}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
