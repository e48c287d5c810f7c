#!/usr/bin/env python
"""Development probe: vml_clip_projection / vml_linear GEMMs of the Charades pass, hot and cold L2, tile-width knob."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vml_b200  # noqa
from vml_b200 import lib as L_
from vml_b200.lib import Dims, call, ptr, stream_ptr

dev = torch.device("cuda")
flush = torch.empty(256 * 2**20, device=dev, dtype=torch.uint8)


def timeit(fn, cold, reps=20):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


B, T, D, d0 = 256, 64, 512, 1024
dims = Dims(T, 16, 4, D, 128, 3, d0, 13, 256)
v = (torch.randn(B * T, d0, device=dev) * 0.5).to(torch.bfloat16)
W = (torch.randn(D, d0, device=dev) * 0.03).to(torch.bfloat16)
bias, pe = torch.randn(D, device=dev), torch.randn(T, D, device=dev)
vmask = torch.ones(B * T, device=dev, dtype=torch.uint8)
fv = torch.empty(B * T, D, device=dev, dtype=torch.bfloat16)
st = stream_ptr()
for bn in ("", "128", "256"):
    if bn:
        os.environ["VML_GEMM_BN"] = bn
    else:
        os.environ.pop("VML_GEMM_BN", None)
    for cold in (False, True):
        us = timeit(lambda: call("vml_clip_projection", ptr(v), ptr(W), ptr(bias), ptr(pe), ptr(vmask), ptr(fv), B, dims, d0, L_.BF16, st), cold)
        print(f"clip_projection M={B*T} N={D} K={d0} BN={bn or 'auto'} cold={cold}: {us:.1f} us  {2.0*B*T*D*d0/us/1e6:.0f} TF/s")
# query-side linears: gin0 [B*Nq, 304] x [2048, 304], gin1 [B*Nq, 512] x [2048, 512], qproj [B*Nq+B, 512] x [2432, 512]
for M, N, K in ((3328, 2048, 304), (3328, 2048, 512), (3584, 2432, 512)):
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    o = torch.empty(M, N, device=dev)
    for bn in ("", "256"):
        if bn:
            os.environ["VML_GEMM_BN"] = bn
        else:
            os.environ.pop("VML_GEMM_BN", None)
        for cold in (False, True):
            us = timeit(lambda: call("vml_linear", ptr(a), ptr(w), ptr(b), ptr(o), M, N, K, N, None, 1, L_.BF16, 1, st), cold)
            print(f"linear M={M} N={N} K={K} BN={bn or 'auto'} cold={cold}: {us:.1f} us  {2.0*M*N*K/us/1e6:.0f} TF/s")
