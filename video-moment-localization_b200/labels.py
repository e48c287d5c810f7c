"""Labels and masks of a collated batch generated on the device (``vml_make_labels``) from the annotation
scalars, instead of being built per sample on the host and copied (dataset.py:95-127,139-158; main.py:118-133
copies 13 tensors per step, 10 of which this call produces in place on the GPU)."""
from __future__ import annotations

from typing import Dict

import torch

from . import lib as L_
from .lib import call, ptr, stream_ptr


def make_labels(times: torch.Tensor, duration: torch.Tensor, nfeats: torch.Tensor, T: int, L: int) -> Dict[str, torch.Tensor]:
    """``times`` [B,2] (ground-truth start/end, seconds), ``duration`` [B], ``nfeats`` [B] -- CUDA tensors (any
    float / int dtype; converted to float64 / int64, the precision of the reference's Python scalars).  Returns the
    reference's keys: sm, ym, ss, ys, se, ye, ya, length_mask, moment_mask, video_mask (dataset.py:165-186)."""
    if not times.is_cuda:
        raise L_.VmlError("make_labels runs on CUDA (sm_100a) only; there is no CPU path")
    L_.load()
    dev, B = times.device, times.shape[0]
    t64 = times.to(torch.float64).contiguous()
    d64 = duration.to(device=dev, dtype=torch.float64).contiguous()
    n64 = nfeats.to(device=dev, dtype=torch.int64).contiguous()
    u8 = lambda *s: torch.empty(*s, device=dev, dtype=torch.uint8)
    f32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    o = {"sm": f32(B, L, L), "ym": u8(B, L, L), "ss": f32(B, L), "ys": u8(B, L), "se": f32(B, L), "ye": u8(B, L), "ya": u8(B, L),
         "length_mask": u8(B, L), "moment_mask": u8(B, L, L), "video_mask": u8(B, T, 1)}
    call("vml_make_labels", ptr(t64), ptr(d64), ptr(n64), B, T, L, *[ptr(o[k]) for k in
         ("sm", "ym", "ss", "ys", "se", "ye", "ya", "length_mask", "moment_mask", "video_mask")], stream_ptr())
    for k in ("ym", "ys", "ye", "ya", "length_mask", "moment_mask"):
        o[k] = o[k].view(torch.bool)
    return o


def sample_clips(raw: torch.Tensor, offsets: torch.Tensor, T: int, start_pos=None, end_pos=None, spos=None) -> Dict[str, torch.Tensor]:
    """Fixed-length clip sampling of a whole batch on the device (``get_fixed_length_features``, dataset.py:40-74).

    ``raw`` [sum nfeats, d0] float32: every video's own clip features back to back; ``offsets`` [B+1] int64 row ranges.
    ``start_pos`` / ``end_pos`` [B]: normalised ground-truth positions (dataset.py:133-134); ``spos`` [B] int32: the random
    start offsets of the training split (dataset.py:44-49; None = 0 as for val / test).  Returns video_features [B,T,d0],
    video_mask [B,T,1] u8, nfeats [B] int64, start_index / end_index [B] int32 -- the values ``__getitem__`` collates."""
    if not raw.is_cuda:
        raise L_.VmlError("sample_clips runs on CUDA (sm_100a) only; there is no CPU path")
    L_.load()
    dev = raw.device
    B, d0 = offsets.numel() - 1, raw.shape[1]
    raw = raw.float().contiguous()
    off = offsets.to(device=dev, dtype=torch.int64).contiguous()
    f64 = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.float64).to(dev).contiguous()
    sp, ep = f64(start_pos), f64(end_pos)
    so = None if spos is None else torch.as_tensor(spos, dtype=torch.int32).to(dev).contiguous()
    out = {"video_features": torch.empty(B, T, d0, device=dev, dtype=torch.float32),
           "video_mask": torch.empty(B, T, 1, device=dev, dtype=torch.uint8),
           "nfeats": torch.empty(B, device=dev, dtype=torch.int64),
           "start_index": torch.empty(B, device=dev, dtype=torch.int32), "end_index": torch.empty(B, device=dev, dtype=torch.int32)}
    status = torch.zeros(1, device=dev, dtype=torch.int32)
    call("vml_sample_clips", ptr(raw), ptr(off), ptr(so), ptr(sp), ptr(ep), B, T, d0, ptr(out["video_features"]), ptr(out["video_mask"]),
         ptr(out["nfeats"]), ptr(out["start_index"]), ptr(out["end_index"]), ptr(status), stream_ptr())
    out["status"] = status                  # bit 1: a sample on which the reference would raise its AssertionError
    return out
