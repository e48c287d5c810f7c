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


def _tf32_ref(A, B):
    """fp64 product of the fp32 operands (what every variant approximates)."""
    return (A.double() @ B.double().t()).float()


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(1000, 384, 520), (260, 132, 4100), (128, 128, 8192)])
def test_gemm_tf32_all_operand_layouts(a_mn, b_mn, M, N, K):
    """gemm_tf32_kernel (tcgen05 kind::tf32, operands straight from fp32 tensors by TMA): K-major and MN-major operands in all
    four combinations, ragged M / N / K, plain store, accumulate and split-K (few output tiles).  Tolerance: TF32 inputs
    (10-bit mantissa, rounded) with fp32 accumulation -> 2e-3 of the result's scale."""
    from vml_b200 import lib
    from vml_b200.training import _gemm
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    ref = _tf32_ref(A, B)
    At, Bt = A.t().contiguous(), B.t().contiguous()                 # [K, M], [K, N]: the MN-major storage
    a_args = (At.data_ptr(), 1, M, 0) if a_mn else (A.data_ptr(), K, 1, 0)
    b_args = (Bt.data_ptr(), 1, N, 0) if b_mn else (B.data_ptr(), K, 1, 0)
    ldc = N + 4
    C = torch.full((M, ldc), 7.0, device="cuda")
    _gemm(*a_args, *b_args, C.data_ptr(), ldc, 1, 0, M, N, K)
    assert "gemm_tf32_kernel" in lib.kernel_names(), "the product did not take the tensor-core path"
    scale = ref.abs().max().item()
    assert (C[:, :N] - ref).abs().max().item() <= 2e-3 * scale
    assert bool((C[:, N:] == 7.0).all()), "columns beyond N were touched"
    # accumulate on top of existing values, with alpha
    _gemm(*a_args, *b_args, C.data_ptr(), ldc, 1, 0, M, N, K, alpha=0.5, acc=1)
    assert (C[:, :N] - 1.5 * ref).abs().max().item() <= 3e-3 * scale


def test_gemm_tf32_device_limited_rows_and_contraction():
    """Live row counts read from device memory: M (dX-type product) and K (dW-type product, both operands MN-major); rows
    beyond the live count hold NaN and must influence nothing."""
    from vml_b200.training import _gemm
    torch.manual_seed(3)
    cap, live, D1, D2 = 6000, 4321, 256, 384
    n_dev = torch.tensor([live], device="cuda", dtype=torch.int32)
    X = torch.randn(cap, D1, device="cuda")
    Y = torch.randn(cap, D2, device="cuda")
    X[live:] = float("nan")
    Y[live:] = float("nan")
    W = torch.randn(D2, D1, device="cuda")                        # Y ~ X.W^T
    # dW[D2, D1] = sum_m Y[m, :]^T X[m, :]   (both MN-major, K = live rows)
    dW = torch.zeros(D2, D1, device="cuda")
    _gemm(Y.data_ptr(), 1, D2, 0, X.data_ptr(), 1, D1, 0, dW.data_ptr(), D1, 1, 0, D2, D1, cap, acc=1, splits=2, k_dev=n_dev.data_ptr())
    ref = (Y[:live].double().t() @ X[:live].double()).float()
    assert torch.isfinite(dW).all()
    assert (dW - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    # dX[cap, D1] = Y . W   (A K-major with device-limited M, B MN-major)
    dX = torch.zeros(cap, D1, device="cuda")
    _gemm(Y.data_ptr(), D2, 1, 0, W.data_ptr(), 1, D1, 0, dX.data_ptr(), D1, 1, 0, cap, D1, D2, m_dev=n_dev.data_ptr())
    ref = (Y[:live].double() @ W.double()).float()
    assert (dX[:live] - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    assert bool((dX[live:] == 0).all()), "rows beyond the live count were written"
