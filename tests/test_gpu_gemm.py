"""vml_linear: CUDA-core fp32 GEMM and tcgen05/TMEM bf16 GEMM against float64 matmul."""
import pytest
import torch

from vml_b200 import lib as L_
from vml_b200.lib import call, ptr, stream_ptr

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


@pytest.mark.parametrize("M,N,K", [(64, 64, 16), (130, 68, 300), (1000, 512, 128), (777, 2048, 300), (33, 1920, 512)])
def test_linear_fp32_simt(M, N, K):
    a, w, b = _rand((M, K), 1), _rand((N, K), 2) / K ** 0.5, _rand((N,), 3)
    out = torch.empty(M, N, device="cuda")
    ad, wd, bd = a.cuda(), w.cuda(), b.cuda()
    call("vml_linear", ptr(ad), ptr(wd), ptr(bd), ptr(out), M, N, K, N, None, 1, L_.FP32, 1, stream_ptr())
    ref = (a.double() @ w.double().t() + b.double())
    assert (out.cpu().double() - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 128), (300, 128, 512), (1000, 512, 128), (4096, 512, 1024),
                                   (257, 64, 40), (100, 32, 32), (20000, 512, 1024), (5000, 128, 512)])
@pytest.mark.parametrize("out_fp32", [1, 0])
def test_linear_bf16_umma(M, N, K, out_fp32):
    a = _rand((M, K), 4).to(torch.bfloat16)
    w = (_rand((N, K), 5) / K ** 0.5).to(torch.bfloat16)
    b = _rand((N,), 6)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if out_fp32 else torch.bfloat16)
    ad, wd, bd = a.cuda(), w.cuda(), b.cuda()
    call("vml_linear", ptr(ad), ptr(wd), ptr(bd), ptr(out), M, N, K, N, None, 1, L_.BF16, out_fp32, stream_ptr())
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t() + b.double()
    err = (out.cpu().double() - ref).abs().max().item()
    tol = (2e-5 if out_fp32 else 1e-2) * max(1.0, ref.abs().max().item())
    assert err < tol, err


def test_linear_bf16_dynamic_row_count():
    """Live row count read from device memory (n_cells * C): rows beyond it must stay untouched."""
    M, N, K, live, scale = 2048, 128, 512, 301, 4
    a = _rand((M, K), 7).to(torch.bfloat16)
    w = (_rand((N, K), 8) / K ** 0.5).to(torch.bfloat16)
    out = torch.full((M, N), -7.0, device="cuda", dtype=torch.bfloat16)
    n_dev = torch.tensor([live], device="cuda", dtype=torch.int32)
    ad, wd = a.cuda(), w.cuda()
    call("vml_linear", ptr(ad), ptr(wd), None, ptr(out), M, N, K, N, ptr(n_dev), scale, L_.BF16, 0, stream_ptr())
    ref = a.double() @ w.double().t()
    rows = live * scale
    assert (out[:rows].cpu().double() - ref[:rows]).abs().max().item() < 1e-2 * ref.abs().max().item()
    assert torch.all(out[rows:] == -7.0)
    out32 = torch.full((M, N), -7.0, device="cuda")
    af, wf = a.float().cuda(), w.float().cuda()
    call("vml_linear", ptr(af), ptr(wf), None, ptr(out32), M, N, K, N, ptr(n_dev), scale, L_.FP32, 1, stream_ptr())
    assert (out32[:rows].cpu().double() - ref[:rows]).abs().max().item() < 1e-4
    assert torch.all(out32[rows:] == -7.0)


@pytest.mark.parametrize("variant", ["pair", "multicast"])
@pytest.mark.parametrize("name,B", [("charadessta", 200), ("tacos", 70), ("activitynet", 10)])
def test_moment_gemm_cluster_variants_are_bit_identical(name, B, variant, monkeypatch):
    """Two opt-in variants of the moment GEMM: an SM pair (VML_GEMM_PAIR=1: tcgen05.mma.cta_group::2, 256 x 256 tile, each
    CTA loading its own A rows and half of the weight tile, the leader issuing one MMA for both) and 2-CTA clusters with
    the weight tile loaded as two TMA-multicast halves (VML_GEMM_CLUSTER=1).  Both walk the same tiles with the same K
    order as the default single-CTA kernel: the whole forward is bit-identical, including an odd number of row tiles
    and the ragged last tile."""
    from oracle import CONFIGS, init_params
    from vml_b200 import synth
    from gpu_util import model_for
    cfg = CONFIGS[name]
    model = model_for(cfg, "bf16", init_params(cfg, 43))
    b = {k: v.cuda() for k, v in synth.make_batch(cfg, B, 31).items()}
    want = [t.clone() for t in model(*[b[k] for k in synth.MODEL_INPUT_KEYS])]
    monkeypatch.setenv("VML_GEMM_PAIR" if variant == "pair" else "VML_GEMM_CLUSTER", "1")
    got = model(*[b[k] for k in synth.MODEL_INPUT_KEYS])
    for a, c in zip(got, want):
        assert torch.equal(a, c)
    assert float(got[0].abs().sum()) > 0
