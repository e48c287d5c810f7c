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
