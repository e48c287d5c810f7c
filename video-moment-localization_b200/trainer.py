"""One training step of the drop-in SMIN (the body of the reference's ``train_epoch`` loop, main.py:136-150):

    optimizer.zero_grad(); out = model(...); loss = loss_fn(...); loss.backward(); optimizer.step()

with the data-parallel gradient exchange in between: every rank scores its own contiguous slice of the
(video, query) batch (``dist.shard_batch``); ``loss_fn`` is a mean over the LOCAL samples (main.py:106), so
the global-batch gradient is the mean over ranks of the local gradients when the slices are equal-sized.
``FusedAdam`` keeps all gradients in one flat buffer: the exchange is ONE all-reduce (NCCL over NVLink on
GPUs), and the 1/world factor is folded into the Adam kernel's ``grad_scale``.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist

from .losses import loss_fn
from .synth import MODEL_INPUT_KEYS


def allreduce_mean_(flat_grad: torch.Tensor, scale_here: bool = True) -> float:
    """Sum ``flat_grad`` over ranks in place.  Returns the factor that turns the sum into the mean over ranks
    (1/world); with ``scale_here`` it is applied in place and 1.0 is returned (CPU/gloo tests), otherwise the
    caller folds it into the optimizer step."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 1.0
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    f = 1.0 / dist.get_world_size()
    if scale_here:
        flat_grad.mul_(f)
        return 1.0
    return f


def train_step(model, optimizer, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
    """``batch``: this rank's device tensors keyed like ``collate_fn``'s output (dataset.py:165-186).
    ``optimizer``: ``optim.FusedAdam`` over ``model.parameters()``.  Returns the local loss (device scalar)."""
    optimizer.zero_grad()
    pm, ps, pe, pa = model(*[batch[k] for k in MODEL_INPUT_KEYS])
    loss = loss_fn(pm, batch["ym"], batch["sm"], batch["moment_mask"], ps, batch["ys"], batch["ss"], pe, batch["ye"], batch["se"],
                   pa, batch["ya"], batch["length_mask"])
    loss.backward()
    flat = optimizer.gather_grads()
    scale = allreduce_mean_(flat, scale_here=False)
    optimizer.step(grad_scale=scale, gathered=True)
    return loss.detach()
