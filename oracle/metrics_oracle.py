"""CPU restatement of the reference loss and R@n,IoU=m metric (test infrastructure).

* ``scaled_iou_bce`` / ``loss_fn``  follow main.py:89-116 with the documented
  one-token fix (``BCELoss(reduction=None)`` raises; evident intent ``'none'``).
* ``compute_ious``               follows utils.py:10-31 (no NMS, as the reference).
* ``nms_topk``                   is OUR definition of temporal NMS (the reference
  has none, utils.py:14) -> parity unpinned; bypassed when threshold >= 1.
"""
from __future__ import annotations

import numpy as np
import torch


def _bce_terms(p, y):
    """Elementwise BCE exactly as the reference computes it: ``nn.BCELoss(reduction='none')`` (main.py:94-96) is
    aten ``binary_cross_entropy`` -- logs clamped at -100 in the forward, and the ANALYTIC backward
    (p - y) / max(p (1 - p), 1e-12), which stays finite at the masked entries p == 0 (autograd through
    ``clamp(log(p))`` would give 0 * inf = NaN there, which the reference never produces)."""
    return torch.nn.functional.binary_cross_entropy(p, y, reduction="none")


def scaled_iou_bce(p, y, s, mask):
    """main.py:89-108.  Two weighted BCE layers: weight s*y on BCE(p, y) plus weight
    (1-s)(1-y) on BCE(1-p, 1-y); masked; per-sample masked mean; batch mean."""
    yf = y.to(p.dtype)
    mk = mask.to(p.dtype)
    if s is not None:
        loss = (s * y.long()) * _bce_terms(p, yf) + ((1 - s) * (1 - y.long())) * _bce_terms(1 - p, 1 - yf)
    else:
        loss = _bce_terms(p, yf)
    loss = loss * mk
    dims = (1, 2) if mask.dim() == 3 else (1,)
    per_sample = loss.sum(dim=dims) / mk.sum(dim=dims)
    return per_sample.mean()


def loss_fn(pm, ym, sm, moment_mask, ps, ys, ss, pe, ye, se, pa, ya, length_mask):
    """main.py:110-116:  L = L_m + L_s + L_e + 0.5 L_a."""
    return (scaled_iou_bce(pm, ym, sm, moment_mask) + scaled_iou_bce(ps, ys, ss, length_mask)
            + scaled_iou_bce(pe, ye, se, length_mask) + 0.5 * scaled_iou_bce(pa, ya, None, length_mask))


def proposal_scores(pm, ps, pe, moment_mask):
    """utils.py:17-21, same op order: ((pm*sqrt(ps_i))*sqrt(pe_j))*mask, flattened."""
    s = pm * torch.sqrt(ps.unsqueeze(2)) * torch.sqrt(pe.unsqueeze(1))
    s = s * moment_mask
    return s.reshape(s.shape[0], -1)


def topk_lowest_index(scores: torch.Tensor, k: int) -> torch.Tensor:
    """Top-k by score descending, ties broken by LOWEST flat index (the product's
    defined tie-break; torch.topk's CPU tie order is unspecified)."""
    s = scores.detach().cpu().numpy()
    n = s.shape[1]
    out = np.empty((s.shape[0], k), dtype=np.int64)
    for b in range(s.shape[0]):
        order = np.lexsort((np.arange(n), -s[b].astype(np.float64)))
        out[b] = order[:k]
    return torch.from_numpy(out)


def nms_topk(scores: torch.Tensor, L: int, k: int, thresh: float) -> torch.Tensor:
    """Greedy temporal NMS over grid proposals (i,j) = [i, j+1) in snippet units:
    repeatedly take the best surviving proposal (ties: lowest flat index) and suppress
    every proposal whose temporal IoU with it is > thresh.  thresh >= 1 suppresses
    nothing and equals ``topk_lowest_index``.  If fewer than k proposals survive, the
    remaining slots are -1."""
    s = scores.detach().cpu().numpy().astype(np.float64)
    B = s.shape[0]
    ii, jj = np.divmod(np.arange(L * L), L)
    out = np.full((B, k), -1, dtype=np.int64)
    for b in range(B):
        alive = np.ones(L * L, dtype=bool)
        for r in range(k):
            if not alive.any():
                break
            cand = np.where(alive)[0]
            best = cand[np.lexsort((cand, -s[b, cand]))[0]]
            out[b, r] = best
            alive[best] = False
            if thresh < 1.0:
                bi, bj = ii[best], jj[best]
                inter = np.maximum(0, np.minimum(jj, bj) + 1 - np.maximum(ii, bi))
                union = np.maximum(jj, bj) + 1 - np.minimum(ii, bi)
                alive &= ~((inter > thresh * union) & (union > 0))
    return torch.from_numpy(out)


def compute_ious(pm, ps, pe, moment_mask, sm, n=(1, 5), m=(0.1, 0.3, 0.5, 0.7), top_indices=None):
    """utils.py:10-31 -> dict 'R@{n}, IoU={m}' -> float COUNT (caller divides by
    num_samples, main.py:163,189,209).  ``top_indices`` lets a test feed the reference's
    own torch.topk order; default is the lowest-index tie-break."""
    scores = proposal_scores(pm, ps, pe, moment_mask)
    if top_indices is None:
        top_indices = topk_lowest_index(scores, max(n))
    top_ious = torch.gather(sm.reshape(sm.shape[0], -1), 1, top_indices)
    metrics = {}
    for n_ in n:
        for m_ in m:
            metrics[f"R@{n_}, IoU={m_}"] = float(((top_ious[:, :n_] > m_).sum(dim=1) > 0).sum().item())
    return metrics
