"""Scaled-IoU BCE loss on the GPU: drop-in for the reference ``main.loss_fn`` / ``bce_loss``
(main.py:89-116).  The reference constructs ``BCELoss(reduction=None)``, which raises on
first use; this implements the evident intent (``reduction='none'``), like the oracle.
One fused kernel computes all four terms; gradients w.r.t. the four score tensors are
produced in the same pass and exposed through autograd.
"""
from __future__ import annotations

import torch

from . import lib as L_
from .lib import call, ptr, stream_ptr


def _u8(t):
    return t.to(torch.uint8).contiguous()


def _launch(pm, ym, sm, moment_mask, ps, ys, ss, pe, ye, se, pa, ya, length_mask, want_grad):
    if not pm.is_cuda:
        raise L_.VmlError("vml_b200.losses runs on CUDA only; there is no CPU path")
    B, L = pm.shape[0], pm.shape[1]
    dev = pm.device
    out = torch.empty(5, device=dev, dtype=torch.float32)          # [loss, L_m, L_s, L_e, L_a]
    scratch = torch.empty(4 * B, device=dev, dtype=torch.float32)
    f = lambda t: t.detach().float().contiguous()
    pm_, ps_, pe_, pa_ = f(pm), f(ps), f(pe), f(pa)
    grads = [torch.empty_like(t) for t in (pm_, ps_, pe_, pa_)] if want_grad else [None] * 4
    # converted operands must stay referenced until the launch is enqueued
    ym_, mm_, ys_, ye_, ya_, lm_ = (_u8(t) for t in (ym, moment_mask, ys, ye, ya, length_mask))
    sm_, ss_, se_ = f(sm), f(ss), f(se)
    call("vml_scaled_iou_bce", ptr(pm_), ptr(ym_), ptr(sm_), ptr(mm_), ptr(ps_), ptr(ys_), ptr(ss_),
         ptr(pe_), ptr(ye_), ptr(se_), ptr(pa_), ptr(ya_), ptr(lm_), B, L,
         out.data_ptr(), out.data_ptr() + 4, ptr(scratch), *[ptr(g) for g in grads], stream_ptr())
    return out, grads


class _ScaledIouBce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pm, ps, pe, pa, ym, sm, moment_mask, ys, ss, ye, se, ya, length_mask):
        need = any(t.requires_grad for t in (pm, ps, pe, pa))
        out, grads = _launch(pm, ym, sm, moment_mask, ps, ys, ss, pe, ye, se, pa, ya, length_mask, need)
        if need:
            ctx.save_for_backward(*grads)
        parts = out[1:]                      # ONE view object: the one marked is the one returned
        ctx.mark_non_differentiable(parts)
        return out[0], parts

    @staticmethod
    def backward(ctx, g_loss, _g_parts):
        gpm, gps, gpe, gpa = ctx.saved_tensors
        return (gpm * g_loss, gps * g_loss, gpe * g_loss, gpa * g_loss) + (None,) * 9


def loss_fn(pm, ym, sm, moment_mask, ps, ys, ss, pe, ye, se, pa, ya, length_mask):
    """main.py:110-116:  L = L_m + L_s + L_e + 0.5 L_a  (scalar tensor on the device)."""
    loss, _ = _ScaledIouBce.apply(pm, ps, pe, pa, ym, sm, moment_mask, ys, ss, ye, se, ya, length_mask)
    return loss


def loss_terms(pm, ym, sm, moment_mask, ps, ys, ss, pe, ye, se, pa, ya, length_mask):
    """(loss, [L_m, L_s, L_e, L_a]) -- the four terms of main.py:111-114, for tests/logging."""
    return _ScaledIouBce.apply(pm, ps, pe, pa, ym, sm, moment_mask, ys, ss, ye, se, ya, length_mask)


class _SingleTerm(torch.autograd.Function):
    """One term of the loss (main.py:89-108) through the same fused kernel, differentiable w.r.t. its score tensor: the kernel
    returns d(L_m + L_s + L_e + 0.5 L_a) / d(score), so the auxiliary term's gradient is rescaled by 2."""

    @staticmethod
    def forward(ctx, p, which, args):
        need = p.requires_grad
        out, grads = _launch(*args, want_grad=need)
        ctx.which = which
        if need:
            ctx.save_for_backward(grads[which] * (2.0 if which == 3 else 1.0))
        return out[1 + which].clone()

    @staticmethod
    def backward(ctx, g):
        (gp,) = ctx.saved_tensors
        return gp * g, None, None


def bce_loss(p, y, s, mask):
    """main.py:89-108 for a single term (3-D map branch or 2-D boundary branch); differentiable w.r.t. ``p`` like the
    reference's (autograd-tested against the oracle)."""
    if p.dim() == 3:
        B, L = p.shape[0], p.shape[1]
        z = torch.zeros(B, L, device=p.device)
        zm = torch.ones(B, L, device=p.device, dtype=torch.uint8)
        half = torch.full((B, L), 0.5, device=p.device)
        return _SingleTerm.apply(p, 0, (p, y, s, mask, half, z, z, half, z, z, half, z, zm))
    B, L = p.shape
    zmap = torch.full((B, L, L), 0.5, device=p.device)
    z3 = torch.zeros(B, L, L, device=p.device)
    o3 = torch.ones(B, L, L, device=p.device, dtype=torch.uint8)
    half = torch.full((B, L), 0.5, device=p.device)
    z = torch.zeros(B, L, device=p.device)
    if s is None:
        return _SingleTerm.apply(p, 3, (zmap, z3, z3, o3, half, z, z, half, z, z, p, y, mask))
    return _SingleTerm.apply(p, 1, (zmap, z3, z3, o3, p, y, s, half, z, z, half, z, mask))
