"""Helpers shared by the -m gpu parity tests (all compute through the C ABI)."""
import ctypes

import torch

from oracle import CONFIGS, init_params
from vml_b200 import lib as L_
from vml_b200 import synth
from vml_b200.lib import Cells, Dims, call, ptr, stream_ptr
from vml_b200.smin import SMIN, Workspace, pack_weights, smin_forward


def dims_of(cfg):
    return Dims(cfg.T, cfg.L, cfg.C, cfg.D, cfg.dl, cfg.layers, cfg.d0, cfg.Nq, cfg.H)


def to_dev(batch, dev="cuda"):
    return {k: v.to(dev) for k, v in batch.items()}


def model_for(cfg, precision, params=None):
    m = SMIN(*cfg.ctor_args(), device=torch.device("cuda"), precision=precision)
    m.load_state_dict(params if params is not None else init_params(cfg, 43), strict=True)
    return m.to("cuda").eval()


def unpack(packed, cells, B, L, inner, prec):
    dt = torch.bfloat16 if prec == L_.BF16 else torch.float32
    dense = torch.empty(B, L, L, inner, device=packed.device, dtype=dt)
    call("vml_unpack_cells", ptr(packed), ptr(dense), cells, B, L, inner, prec, stream_ptr())
    return dense.float()


def rel_err(a, b, floor=1e-6):
    a = a.double().cpu()
    b = b.double().cpu()
    return ((a - b).abs() / b.abs().clamp_min(floor)).max().item()


def scaled_err(a, b):
    """max |a-b| / max|b|  -- for intermediates whose entries pass through zero."""
    a = a.double().cpu()
    b = b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
