"""Training path: the hand-written backward (csrc/backward.cu + vml_gemm_strided) against autograd of the
CPU oracle, parameter by parameter, and one Adam step against torch.optim.Adam."""
import os

import pytest
import torch

from oracle import CONFIGS, init_params, smin_forward as oracle_forward
from oracle import metrics_oracle as mo
from vml_b200 import synth
from vml_b200.losses import loss_fn

pytestmark = pytest.mark.gpu

from gpu_util import model_for  # noqa: E402


def _oracle_grads(cfg, params, batch):
    p = {k: v.clone().double().requires_grad_(True) for k, v in params.items()}
    b = {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()}
    out = oracle_forward(p, cfg, *[b[k] for k in synth.MODEL_INPUT_KEYS])
    loss = mo.loss_fn(out[0], b["ym"], b["sm"], b["moment_mask"], out[1], b["ys"], b["ss"], out[2], b["ye"], b["se"], out[3], b["ya"],
                      b["length_mask"])
    loss.backward()
    return loss.item(), {k: v.grad for k, v in p.items()}


def test_gemm_strided_matches_torch():
    from vml_b200.training import _gemm
    torch.manual_seed(0)
    Bt, M, N, K = 3, 70, 45, 130
    A = torch.randn(Bt, M, K, device="cuda")
    Bm = torch.randn(Bt, N, K, device="cuda")
    C = torch.zeros(Bt, M, N, device="cuda")
    _gemm(A.data_ptr(), K, 1, M * K, Bm.data_ptr(), K, 1, N * K, C.data_ptr(), N, 1, M * N, M, N, K, batch=Bt)
    ref = A @ Bm.transpose(1, 2)
    assert (C - ref).abs().max() < 1e-3
    # transposed operands + split-K accumulation:  C2[m][n] += sum_k A[0][k][m] * Bm[0][k][n]   (a dW-style product)
    A2, B2 = torch.randn(5000, 33, device="cuda"), torch.randn(5000, 20, device="cuda")
    C2 = torch.ones(33, 20, device="cuda")
    _gemm(A2.data_ptr(), 1, 33, 0, B2.data_ptr(), 1, 20, 0, C2.data_ptr(), 20, 1, 0, 33, 20, 5000, acc=1, splits=4)
    ref2 = 1.0 + A2.t() @ B2
    assert ((C2 - ref2).abs() / ref2.abs().clamp_min(1.0)).max() < 1e-4


@pytest.mark.parametrize("name,B,seed", [("tiny", 5, 21), ("tiny_r2", 4, 22), ("charadessta", 3, 23), ("activitynet", 2, 24),
                                         ("tacos", 3, 25)])          # tacos: the shape of BASELINE configs[3]
def test_backward_matches_oracle_autograd(name, B, seed):
    cfg = CONFIGS[name]
    params = init_params(cfg, 43)
    batch = synth.make_batch(cfg, B, seed)
    ref_loss, ref = _oracle_grads(cfg, params, batch)
    model = model_for(cfg, "fp32", params)
    model.train()
    d = {k: v.cuda() for k, v in batch.items()}
    out = model(*[d[k] for k in synth.MODEL_INPUT_KEYS])
    loss = loss_fn(out[0], d["ym"], d["sm"], d["moment_mask"], out[1], d["ys"], d["ss"], out[2], d["ye"], d["se"], out[3], d["ya"],
                   d["length_mask"])
    # the training path runs its dense products as TF32 on the tensor cores (forward and backward): loss within 2e-4
    # (1e-5 with VML_TRAIN_FP32=1, the CUDA-core validation arithmetic), gradients within 2e-3 of their scale
    assert abs(loss.item() - ref_loss) < (1e-5 if os.environ.get("VML_TRAIN_FP32") else 2e-4) * abs(ref_loss)
    loss.backward()
    bad = []
    for n, p in model.named_parameters():
        assert p.grad is not None, n
        r = ref[n].float()
        assert torch.isfinite(r).all(), ("oracle gradient", n)          # a NaN reference would make the comparison vacuous
        assert torch.isfinite(p.grad).all(), n
        err = (p.grad.cpu() - r).abs().max().item()
        scale = max(r.abs().max().item(), 1e-6)
        if not (err <= 2e-3 * scale + 1e-7):
            bad.append((n, err, scale))
    assert not bad, bad


def test_adam_step_matches_torch():
    from vml_b200.optim import FusedAdam
    torch.manual_seed(1)
    ps = [torch.randn(300, 7, device="cuda", requires_grad=True), torch.randn(11, device="cuda", requires_grad=True)]
    ref_ps = [p.detach().clone().requires_grad_(True) for p in ps]
    ours, ref = FusedAdam(ps, lr=1e-3), torch.optim.Adam(ref_ps, lr=1e-3)
    for step in range(3):
        for p, r in zip(ps, ref_ps):
            gr = torch.randn_like(p)
            p.grad, r.grad = gr.clone(), gr.clone()
        ours.step(); ref.step()
    for p, r in zip(ps, ref_ps):
        assert (p - r).abs().max() < 1e-6


def test_fused_adam_checkpoint_round_trips_with_torch_adam():
    """optimizer.state_dict() of the reference's checkpoints (main.py:270-274) moves between FusedAdam and torch.optim.Adam."""
    from vml_b200.optim import FusedAdam
    torch.manual_seed(2)
    shapes = [(40, 9), (13,), (3, 5, 2)]
    ps = [torch.randn(*sh, device="cuda", requires_grad=True) for sh in shapes]
    ref_ps = [p.detach().clone().requires_grad_(True) for p in ps]
    ours, ref = FusedAdam(ps, lr=3e-3), torch.optim.Adam(ref_ps, lr=3e-3)

    def step_both(a, a_ps, b, b_ps):
        for p, r in zip(a_ps, b_ps):
            gr = torch.randn_like(p)
            p.grad, r.grad = gr.clone(), gr.clone()
        a.step(); b.step()

    for _ in range(2):
        step_both(ours, ps, ref, ref_ps)
    # ours -> stock Adam
    ps2 = [p.detach().clone().requires_grad_(True) for p in ps]
    stock = torch.optim.Adam(ps2, lr=1.0)
    stock.load_state_dict(ours.state_dict())
    assert stock.param_groups[0]["lr"] == 3e-3
    # stock Adam -> ours
    ps3 = [r.detach().clone().requires_grad_(True) for r in ref_ps]
    mine = FusedAdam(ps3, lr=1.0)
    mine.load_state_dict(ref.state_dict())
    assert mine.t == 2 and mine.lr == 3e-3
    for _ in range(2):
        grads = [torch.randn_like(p) for p in ps]
        for group in (ps, ref_ps, ps2, ps3):
            for p, g in zip(group, grads):
                p.grad = g.clone()
        ours.step(); ref.step(); stock.step(); mine.step()
    for a, b, c, d in zip(ps, ref_ps, ps2, ps3):
        assert (a - b).abs().max() < 1e-6 and (c - b).abs().max() < 1e-6 and (d - b).abs().max() < 1e-6


def test_fused_adam_steps_refresh_packed_weights():
    """Regression (round-1 advisor finding): ``FusedAdam.step`` writes the parameters from a kernel, so SMIN's packed-weight
    cache must be invalidated explicitly.  Three ``train_step``s of SMIN + FusedAdam must track SMIN + torch.optim.Adam on the
    same batches (losses and parameters), and an eval forward must move after a step."""
    from vml_b200.optim import FusedAdam
    from vml_b200.trainer import train_step
    cfg = CONFIGS["tiny"]
    params = init_params(cfg, 43)
    ours, ref = model_for(cfg, "fp32", params), model_for(cfg, "fp32", params)
    ours.train(); ref.train()
    opt, ref_opt = FusedAdam(ours.parameters(), lr=1e-2), torch.optim.Adam(ref.parameters(), lr=1e-2)
    batches = [{k: v.cuda() for k, v in synth.make_batch(cfg, 4, 300 + i).items()} for i in range(2)]
    probe = batches[0]

    def eval_scores(m):
        m.eval()
        with torch.no_grad():
            out = m(*[probe[k] for k in synth.MODEL_INPUT_KEYS])[0].clone()
        m.train()
        return out

    before = eval_scores(ours)
    losses, ref_losses = [], []
    for step in range(3):
        d = batches[step % 2]                      # batch 0 is repeated at step 2: its loss must have moved
        losses.append(train_step(ours, opt, d).item())
        ref_opt.zero_grad()
        out = ref(*[d[k] for k in synth.MODEL_INPUT_KEYS])
        loss = loss_fn(out[0], d["ym"], d["sm"], d["moment_mask"], out[1], d["ys"], d["ss"], out[2], d["ye"], d["se"], out[3], d["ya"],
                       d["length_mask"])
        loss.backward()
        ref_opt.step()
        ref_losses.append(loss.item())
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 1e-5 * abs(b), (losses, ref_losses)
    assert abs(losses[2] - losses[0]) > 1e-4 * abs(losses[0]), "the loss on the repeated batch did not move: frozen weights"
    for (n, p), (_, r) in zip(ours.named_parameters(), ref.named_parameters()):
        # Adam divides by sqrt(v): rounding noise in small gradients is amplified towards lr; 0.5 % of lr bounds it
        assert (p - r).abs().max() <= 5e-3 * 1e-2, n
    after = eval_scores(ours)
    assert (after - before).abs().max() > 1e-4, "eval forward unchanged after three optimizer steps: stale packed weights"
    assert (after - eval_scores(ref)).abs().max() < 1e-4


def test_fused_adam_refuses_rebound_parameters():
    from vml_b200 import lib
    from vml_b200.optim import FusedAdam
    ps = [torch.randn(10, 3, device="cuda", requires_grad=True)]
    opt = FusedAdam(ps, lr=1e-3)
    ps[0].grad = torch.ones_like(ps[0])
    opt.step()
    ps[0].data = ps[0].data.clone()                # what model.to() / flatten_parameters would do
    with pytest.raises(lib.VmlError):
        opt.step()
