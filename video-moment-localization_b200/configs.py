"""Model configurations (the reference's config/*.yml:5-13) and deterministic random-init
parameters keyed like the reference ``SMIN.state_dict()``.  Host-side, no compute: shared by
the product (bench, tests) and by the CPU oracle."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class SminConfig:
    """Model hyper-parameters, one per reference YAML (config/*.yml:5-13)."""
    name: str
    T: int          # clips per video
    L: int          # side of the moment map
    C: int          # sub-clips per moment
    D: int          # feature width
    dl: int         # content/word interaction width
    layers: int     # number of SMI layers
    d0: int         # input clip-feature width
    Nq: int         # max query words
    H: int          # LSTM hidden size (D == 2H)

    def ctor_args(self):
        return (self.T, self.L, self.C, self.D, self.dl, self.layers, self.d0, self.Nq, self.H)


CONFIGS = {
    "charadessta": SminConfig("charadessta", 64, 16, 4, 512, 128, 3, 1024, 13, 256),
    "activitynet": SminConfig("activitynet", 128, 64, 4, 512, 128, 3, 500, 20, 256),
    "tacos":       SminConfig("tacos", 128, 32, 4, 512, 128, 3, 4096, 14, 256),
    # small shapes for fast CPU tests / compute-sanitizer (not a reference config)
    "tiny":        SminConfig("tiny", 32, 8, 4, 64, 32, 2, 40, 6, 32),
    # ActivityNet-style irregular windows (T/L = 2 < C) at small size
    "tiny_r2":     SminConfig("tiny_r2", 16, 8, 4, 64, 32, 2, 24, 5, 32),
}


# --------------------------------------------------------------------------
# deterministic parameters (exact integer RNG -> identical on every machine)
# --------------------------------------------------------------------------
def _uniform(rng: np.random.Generator, shape, bound: float) -> torch.Tensor:
    u = rng.integers(0, 1 << 24, size=shape, dtype=np.int64).astype(np.float64) / float(1 << 24)
    return torch.from_numpy(((2.0 * u - 1.0) * bound).astype(np.float32))


def _normalish(rng: np.random.Generator, shape) -> torch.Tensor:
    # Irwin-Hall(4) scaled to unit variance: exact integer draws, no libm.
    u = rng.integers(0, 1 << 24, size=(4,) + tuple(shape), dtype=np.int64).astype(np.float64) / float(1 << 24)
    return torch.from_numpy(((u.sum(0) - 2.0) * math.sqrt(3.0)).astype(np.float32))


def init_params(cfg: SminConfig, seed: int = 43) -> dict:
    """Random-init parameters with the same key names / shapes / scale as
    ``SMIN.state_dict()`` (models.py:21-23,46,134-135,204-205,236-240,285-286,329-332).
    Scales follow PyTorch's defaults (U(+-1/sqrt(fan_in)); N(0,1) embedding)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = {}
    D, dl, H = cfg.D, cfg.dl, cfg.H

    def lin(name, out_f, in_f, shape_w=None):
        b = 1.0 / math.sqrt(in_f)
        p[name + ".weight"] = _uniform(rng, shape_w or (out_f, in_f), b)
        p[name + ".bias"] = _uniform(rng, (out_f,), b)

    lin("backbone.videoencoder.ve", D, cfg.d0)
    p["backbone.videoencoder.pe.weight"] = _normalish(rng, (cfg.T, D))
    b = 1.0 / math.sqrt(H)
    for layer in range(2):
        for sfx in ("", "_reverse"):
            in_f = 300 if layer == 0 else 2 * H
            pre = "backbone.queryencoder.lstm."
            p[f"{pre}weight_ih_l{layer}{sfx}"] = _uniform(rng, (4 * H, in_f), b)
            p[f"{pre}weight_hh_l{layer}{sfx}"] = _uniform(rng, (4 * H, H), b)
            p[f"{pre}bias_ih_l{layer}{sfx}"] = _uniform(rng, (4 * H,), b)
            p[f"{pre}bias_hh_l{layer}{sfx}"] = _uniform(rng, (4 * H,), b)
    for k in range(cfg.layers):
        cu = f"smis.{k}.content_unit."
        lin(cu + "linear_c_hat", dl, D)
        lin(cu + "linear_w_hat", dl, D)
        lin(cu + "linear_s_hat", dl, D)
        lin(cu + "linear_c", D, dl)
        lin(cu + "attn_layer.W_q", dl, dl)
        lin(cu + "attn_layer.W_k", dl, dl)
        bu = f"smis.{k}.boundary_unit.attn_layer."
        lin(bu + "W_q", D, D)
        lin(bu + "W_k", D, D)
        mu = f"smis.{k}.moment_unit."
        lin(mu + "conv_layer_fb", D, D, (D, D, 1, 1))
        lin(mu + "conv_layer_fc", D, D, (D, D, 1, 1))
    lin("localization.conv_layer_pm", 1, D, (1, D, 1, 1))
    for nm in ("ps", "pe", "pa"):
        lin(f"localization.conv_layer_{nm}", 1, D, (1, D, 1))
    return p
