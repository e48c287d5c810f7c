"""R@n, IoU=m evaluation on the GPU: drop-in for the reference ``utils.compute_ious``
(utils.py:10-31) plus a device-resident accumulator that removes its 8 ``.item()`` syncs.

Top-k ties are broken by LOWEST flat index (torch.topk leaves tie order unspecified).
Temporal NMS is OFF by default, like the reference (utils.py:14 "NMS NOT IMPLEMENTED YET");
``nms_threshold < 1`` enables greedy NMS on the integer proposal grid (parity unpinned).
"""
from __future__ import annotations

import ctypes
from collections import defaultdict
from fractions import Fraction

import torch

from . import lib as L_
from .lib import call, ptr, stream_ptr

_NS = (1, 5)
_MS = (0.1, 0.3, 0.5, 0.7)


def _as_u8(t):
    return t.to(torch.uint8).contiguous()


def score_topk_recall(pm, ps, pe, moment_mask, sm, k: int = 5, nms_threshold: float = 1.0, counts=None, step_counts=None,
                      step_group: int = 0, n=None, m=None):
    """One launch: scores, top-k indices/scores/IoUs per sample, and the hit counters
    (int64 [len(n), len(m)] -- [2,4] for the reference defaults --, accumulated in place when ``counts`` is given;
    ``step_counts``, optional, is a second accumulator for per-step read-back).  Everything stays on the device.
    ``n`` / ``m``: the caller's own lists (utils.py:10); ``None`` = the reference defaults."""
    if not pm.is_cuda:
        raise L_.VmlError("vml_b200.evaluate runs on CUDA only; there is no CPU path")
    B, L = pm.shape[0], pm.shape[1]
    dev = pm.device
    default_nm = n is None and m is None
    ns = list(_NS) if n is None else [int(x) for x in n]
    ms = list(_MS) if m is None else [float(x) for x in m]
    if not default_nm:
        if not (1 <= len(ns) <= 8 and 1 <= len(ms) <= 8) or min(ns) < 1:
            raise ValueError("compute_ious: 1..8 values of n (each >= 1) and 1..8 values of m")
        k = max(ns)                      # the reference takes topk(max(n)) (utils.py:23)
    if k > min(32, L * L):
        raise ValueError("compute_ious: max(n) must be <= min(32, L*L)")
    if counts is None:
        counts = torch.zeros(len(ns), len(ms), device=dev, dtype=torch.int64)
    top_idx = torch.empty(B, k, device=dev, dtype=torch.int32)
    top_score = torch.empty(B, k, device=dev, dtype=torch.float32)
    top_iou = torch.empty(B, k, device=dev, dtype=torch.float32)
    fr = Fraction(nms_threshold).limit_denominator(1000) if nms_threshold < 1.0 else Fraction(1, 1)
    # keep every converted operand referenced until the launch is enqueued
    pm_, ps_, pe_, sm_ = (t.float().contiguous() for t in (pm, ps, pe, sm))
    mask_ = _as_u8(moment_mask)
    if default_nm:
        call("vml_score_topk_recall", ptr(pm_), ptr(ps_), ptr(pe_), ptr(mask_), ptr(sm_), B, L, k, fr.numerator, fr.denominator,
             ptr(top_idx), ptr(top_score), ptr(top_iou), ptr(counts), ptr(step_counts), step_group, stream_ptr())
    else:
        ns_c, ms_c = (ctypes.c_int32 * len(ns))(*ns), (ctypes.c_float * len(ms))(*ms)      # host arrays, read at enqueue time
        call("vml_score_topk_recall_nm", ptr(pm_), ptr(ps_), ptr(pe_), ptr(mask_), ptr(sm_), B, L, k, fr.numerator, fr.denominator,
             ptr(top_idx), ptr(top_score), ptr(top_iou), ptr(counts), ptr(step_counts), step_group,
             ctypes.cast(ns_c, ctypes.c_void_p), len(ns), ctypes.cast(ms_c, ctypes.c_void_p), len(ms), stream_ptr())
    return top_idx, top_score, top_iou, counts


def compute_ious(pm, ps, pe, moment_mask, sm, n=[1, 5], m=[0.1, 0.3, 0.5, 0.7], nms_threshold: float = 1.0):
    """Same signature and return value as the reference: dict 'R@{n}, IoU={m}' -> float count
    (the caller divides by num_samples, main.py:163,189,209).  One D2H copy of 8 counters."""
    n, m = list(n), list(m)
    _, _, _, counts = score_topk_recall(pm, ps, pe, moment_mask, sm, k=max(n), nms_threshold=nms_threshold, n=n, m=m)
    host = counts.cpu()
    metrics = defaultdict(lambda: 0.0)
    for a, n_ in enumerate(n):                     # key text = the caller's own values, like the f-string of utils.py:29
        for t, m_ in enumerate(m):
            metrics[f"R@{n_}, IoU={m_}"] += float(host[a, t].item())
    return metrics


class RecallAccumulator:
    """Device-side running sums over batches (replaces the per-batch dict adds of
    main.py:155-156,184-185,205-206); ``result()`` does the single D2H read."""

    def __init__(self, device, nms_threshold: float = 1.0):
        self.counts = torch.zeros(2, 4, device=device, dtype=torch.int64)
        self.num_samples = 0
        self.nms_threshold = nms_threshold

    def update(self, pm, ps, pe, moment_mask, sm):
        score_topk_recall(pm, ps, pe, moment_mask, sm, 5, self.nms_threshold, self.counts)
        self.num_samples += pm.shape[0]

    def result(self, normalize: bool = True):
        host = self.counts.cpu()
        den = float(self.num_samples) if normalize and self.num_samples else 1.0
        return {f"R@{n_}, IoU={m_}": float(host[a, t].item()) / den for a, n_ in enumerate(_NS) for t, m_ in enumerate(_MS)}
