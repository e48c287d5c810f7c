"""Seeded synthetic (video, query) batches in the reference's collate format.

Host-side logic.  Shapes, dtypes and distributions follow SURVEY.md section 8(d);
masks and labels restate the reference dataset's formulas so that loss and metric
see inputs of the same kind the reference would produce:

  * length_mask / moment_mask      dataset.py:142-149
  * sm  (IoU map, inter / hull)    dataset.py:95-110
  * ss / se (Gaussian penalties)   dataset.py:112-121
  * ya  (snippet inside GT)        dataset.py:123-127
  * ym / ys / ye = (. > 0.5)       dataset.py:151-158

Random draws use an exact integer generator (numpy PCG64 -> 24-bit integers) so the
same seed yields bit-identical batches on every machine.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch


def _unit_normalish(rng: np.random.Generator, shape) -> np.ndarray:
    """Irwin-Hall(4) scaled to unit variance, float32; no libm calls."""
    acc = np.zeros(shape, dtype=np.float64)
    for _ in range(4):
        acc += rng.integers(0, 1 << 24, size=shape, dtype=np.int64)
    return ((acc / float(1 << 24) - 2.0) * math.sqrt(3.0)).astype(np.float32)


def _uniform01(rng: np.random.Generator, n: int) -> np.ndarray:
    return rng.integers(0, 1 << 24, size=n, dtype=np.int64).astype(np.float64) / float(1 << 24)


def iou_map(L: int, gt_s: float, gt_e: float, duration: float) -> torch.Tensor:
    """sm[L,L]: proposal (i,j) = [i*dur/L, (j+1)*dur/L]; IoU = inter / hull, float32,
    same op order as dataset.py:95-110."""
    s_times = torch.arange(0, L).float() * duration / L
    e_times = torch.arange(1, L + 1).float() * duration / L
    ps = s_times.repeat_interleave(L)
    pe = e_times.repeat(L)
    gt = torch.tensor([gt_s, gt_e])
    zero = torch.tensor(0.0)
    inter = torch.max(zero, torch.min(pe, gt[1]) - torch.max(ps, gt[0]))
    union = torch.max(zero, torch.max(pe, gt[1]) - torch.min(ps, gt[0]))
    return (inter / union).reshape(L, L)


def boundary_penalties(L: int, tau_s: float, tau_e: float, duration: float):
    """ss, se[L] = exp(-(t - tau)^2 / (2 sigma^2)), sigma = (tau_e - tau_s)/5  (dataset.py:112-121)."""
    s_times = torch.arange(0, L).float() * duration / L
    e_times = torch.arange(1, L + 1).float() * duration / L
    sigma = (tau_e - tau_s) / 5.0
    return (torch.exp(-(s_times - tau_s) ** 2 / (2.0 * sigma ** 2)),
            torch.exp(-(e_times - tau_e) ** 2 / (2.0 * sigma ** 2)))


def snippet_label(L: int, tau_s: float, tau_e: float, duration: float) -> torch.Tensor:
    """ya[L]: snippet fully inside the ground-truth moment (dataset.py:123-127)."""
    s_times = torch.arange(0, L).float() * duration / L
    e_times = torch.arange(1, L + 1).float() * duration / L
    return torch.logical_and(s_times >= tau_s, e_times <= tau_e)


def masks_from_nfeats(nfeats: int, T: int, L: int):
    """video_mask[T,1] u8, length_mask[L] bool, moment_mask[L,L] bool (dataset.py:142-149)."""
    vm = np.zeros((T, 1), dtype=np.uint8)
    vm[:nfeats] = 1
    lm = np.zeros(L, dtype=bool)
    lm[: math.ceil(nfeats / (T / L))] = True
    mm = np.triu(np.logical_and.outer(lm, lm))
    return vm, lm, mm


def make_batch(cfg, B: int, seed: int = 0, full_length: bool = False,
               features: bool = True, nfeats_range=None) -> Dict[str, torch.Tensor]:
    """One synthetic batch (CPU tensors) keyed like the reference's ``collate_fn`` output
    (dataset.py:76-90,165-186).  ``cfg`` needs attributes T, L, d0, Nq.  With
    ``full_length`` every video has T clips (worst-case map occupancy)."""
    T, L, d0, Nq = cfg.T, cfg.L, cfg.d0, cfg.Nq
    rng = np.random.Generator(np.random.PCG64(seed))
    r = T // L
    nfeats = rng.integers(T // 2, T + 1, size=B)
    if nfeats_range is not None:                          # edge-case batches: very short (or fixed-range) videos
        nfeats = rng.integers(nfeats_range[0], nfeats_range[1] + 1, size=B)
    elif full_length:
        nfeats[:] = T
    else:
        nfeats[0] = T                                     # >= 1 full-length sample
        if B > 1 and r > 1:
            nfeats[1] = T - r - 1                         # >= 1 length that is not a multiple of T/L
    qlen = rng.integers(min(3, Nq), Nq + 1, size=B)

    out = {}
    if features:
        vf = _unit_normalish(rng, (B, T, d0))
        qf = _unit_normalish(rng, (B, Nq, 300))
    else:
        vf = np.zeros((B, T, d0), dtype=np.float32)
        qf = np.zeros((B, Nq, 300), dtype=np.float32)
    vmask = np.zeros((B, T, 1), dtype=np.uint8)
    lmask = np.zeros((B, L), dtype=bool)
    mmask = np.zeros((B, L, L), dtype=bool)
    qmask = np.zeros((B, Nq, 1), dtype=np.uint8)
    for b in range(B):
        vmask[b], lmask[b], mmask[b] = masks_from_nfeats(int(nfeats[b]), T, L)
        vf[b, int(nfeats[b]):] = 0.0
        qmask[b, : int(qlen[b])] = 1
        qf[b, int(qlen[b]):] = 0.0

    dur = 10.0 + 50.0 * _uniform01(rng, B)
    tau_s = 0.6 * dur * _uniform01(rng, B)
    tau_e = np.minimum(dur, tau_s + 2.0 + (0.4 * dur - 2.0) * _uniform01(rng, B))
    sm = torch.stack([iou_map(L, float(tau_s[b]), float(tau_e[b]), float(dur[b])) for b in range(B)])
    pen = [boundary_penalties(L, float(tau_s[b]), float(tau_e[b]), float(dur[b])) for b in range(B)]
    ss = torch.stack([x[0] for x in pen])
    se = torch.stack([x[1] for x in pen])
    ya = torch.stack([snippet_label(L, float(tau_s[b]), float(tau_e[b]), float(dur[b])) for b in range(B)])

    out["video_features"] = torch.from_numpy(vf)
    out["video_mask"] = torch.from_numpy(vmask)
    out["query_features"] = torch.from_numpy(qf)
    out["query_mask"] = torch.from_numpy(qmask)
    out["length_mask"] = torch.from_numpy(lmask)
    out["moment_mask"] = torch.from_numpy(mmask)
    out["sm"] = sm
    out["ym"] = sm > 0.5
    out["ss"], out["ys"] = ss, ss > 0.5
    out["se"], out["ye"] = se, se > 0.5
    out["ya"] = ya
    out["nfeats"] = torch.from_numpy(nfeats.astype(np.int64))
    out["duration"] = torch.from_numpy(dur)
    out["times"] = torch.from_numpy(np.stack([tau_s, tau_e], axis=1))
    return out


MODEL_INPUT_KEYS = ("video_features", "video_mask", "query_features", "query_mask", "length_mask", "moment_mask")
LOSS_LABEL_KEYS = ("sm", "ym", "ss", "ys", "se", "ye", "ya")
