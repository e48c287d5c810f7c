#!/usr/bin/env python
"""Time individual C-ABI entry points with CUDA events (GPU box only).
    python tools/time_stage.py lstm|gemm ...
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vml_b200  # noqa
from vml_b200 import lib as L_
from vml_b200.lib import call, ptr, stream_ptr


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def lstm(B=64, Nq=13, H=256):
    gin = torch.randn(B * Nq, 8 * H, device="cuda") * 0.1
    whh = torch.randn(2, H, 4 * H, device="cuda") * 0.05
    qlen = torch.randint(3, Nq + 1, (B,), device="cuda", dtype=torch.int32)
    y = torch.empty(B, Nq, 2 * H, device="cuda")
    fs = torch.empty(B, 2 * H, device="cuda")
    st = stream_ptr()
    for full in (False, True):
        if full:
            qlen.fill_(Nq)
        us = timeit(lambda: call("vml_lstm_layer", ptr(gin), ptr(whh), ptr(qlen), ptr(y), None, ptr(fs), None, B, Nq, H, st))
        print(f"lstm_layer B={B} Nq={Nq} H={H} full={full}: {us:.1f} us  ({us / Nq:.2f} us/step)")


def lstm_tc(B=64, Nq=13, H=256):
    from vml_b200.smin import pack_lstm_fragments
    gin = torch.randn(B * Nq, 8 * H, device="cuda") * 0.1
    frag = pack_lstm_fragments(torch.randn(4 * H, H, device="cuda") * 0.05, torch.randn(4 * H, H, device="cuda") * 0.05)
    qlen = torch.full((B,), Nq, device="cuda", dtype=torch.int32)
    y = torch.empty(B, Nq, 2 * H, device="cuda")
    y16 = torch.empty(B, Nq, 2 * H, device="cuda", dtype=torch.bfloat16)
    fs = torch.empty(B, 2 * H, device="cuda")
    st = stream_ptr()
    us = timeit(lambda: call("vml_lstm_layer_tc", ptr(gin), ptr(frag), ptr(qlen), ptr(y), ptr(y16), ptr(fs), None, B, Nq, H, st))
    print(f"lstm_layer_tc B={B} Nq={Nq}: {us:.1f} us  ({us / Nq:.2f} us/step)")
    return us


def gemm(M, N, K, prec="bf16", out32=0):
    p = L_.PREC[prec]
    dt = torch.bfloat16 if p == L_.BF16 else torch.float32
    a = torch.randn(M, K, device="cuda").to(dt)
    w = torch.randn(N, K, device="cuda").to(dt)
    b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if (out32 or p == L_.FP32) else torch.bfloat16)
    st = stream_ptr()
    us = timeit(lambda: call("vml_linear", ptr(a), ptr(w), ptr(b), ptr(out), M, N, K, N, None, 1, p, out32, st))
    print(f"linear {prec} M={M} N={N} K={K}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.1f} TFLOP/s")


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "lstm":
        lstm()
        lstm(B=8)
        lstm(B=64, Nq=20)
    elif what == "lstm_tc":
        a = lstm_tc(Nq=13)
        b = lstm_tc(Nq=26)
        print(f"per-step slope {(b - a) / 13:.2f} us, fixed {a - 13 * (b - a) / 13:.1f} us")
        lstm_tc(B=16, Nq=13)
        lstm_tc(B=256, Nq=13)
    elif what == "gemm":
        for shp in [(21504, 128, 512), (21504, 512, 128), (5376, 512, 1024), (4096, 512, 1024), (832, 2048, 512), (896, 2816, 512),
                    (65536, 512, 1024), (262144, 128, 512)]:
            gemm(*shp)
