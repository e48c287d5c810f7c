#!/usr/bin/env python
"""Development probe: gemm_tf32_kernel in its four operand layouts, with the MN-major descriptor knobs swept."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vml_b200  # noqa
from vml_b200.training import _gemm

torch.manual_seed(0)
M, N, K = 256, 256, 2048
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
ref = (A.double() @ B.double().t()).float()
At, Bt = A.t().contiguous(), B.t().contiguous()


def run(a_mn, b_mn):
    a = (At.data_ptr(), 1, M, 0) if a_mn else (A.data_ptr(), K, 1, 0)
    b = (Bt.data_ptr(), 1, N, 0) if b_mn else (B.data_ptr(), K, 1, 0)
    C = torch.zeros(M, N, device="cuda")
    try:
        _gemm(*a, *b, C.data_ptr(), N, 1, 0, M, N, K)
        torch.cuda.synchronize()
    except Exception as e:
        return f"EXC {e}"
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    return f"rel_err={err:.3e} C00={C[0,0].item():.3f} ref00={ref[0,0].item():.3f} nz={int((C!=0).sum())}"


print("KK", run(False, False))
print("default B_MN", run(False, True)); print("default A_MN", run(True, False)); print("default both", run(True, True))
for swz, lt, lbo, sbo, ks in itertools.product((4, 3), (1, 2), (4096,), (512, 1024, 256), (1024, 512)):
    os.environ.update(VML_TF_SWZ=str(swz), VML_TF_LTYPE=str(lt), VML_TF_LBO=str(lbo), VML_TF_SBO=str(sbo), VML_TF_KSTEP=str(ks))
    r = run(False, True)
    tag = " <==" if "rel_err" in r and float(r.split("rel_err=")[1].split()[0]) < 5e-3 else ""
    print(f"B_MN swz={swz} ltype={lt} lbo={lbo} sbo={sbo} kstep={ks}: {r}{tag}")
