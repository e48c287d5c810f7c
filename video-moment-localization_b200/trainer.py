"""One training step of the drop-in SMIN (the body of the reference's ``train_epoch`` loop, main.py:136-150):

    optimizer.zero_grad(); out = model(...); loss = loss_fn(...); loss.backward(); optimizer.step()

with the data-parallel gradient exchange in between: every rank scores its own contiguous slice of the
(video, query) batch (``dist.shard_batch``); ``loss_fn`` is a mean over the LOCAL samples (main.py:106), so
the global-batch gradient is  sum_r (B_r / B_global) * grad_r  -- the plain mean over ranks only when the slices
are equal-sized.  ``train_step`` therefore weights each rank's gradient by its share of the global batch
(ragged last batches, ``B % world != 0``), folded into the Adam kernel's ``grad_scale`` when the shares are equal.
``FusedAdam`` keeps all gradients in one flat buffer: the exchange is ONE all-reduce (NCCL over NVLink on GPUs).
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


def shard_weight(local_b: int, global_b: int | None, device) -> float | torch.Tensor:
    """This rank's weight B_local / B_global in the global-batch gradient, relative to the 1/world the caller
    applies afterwards (so equal shards give exactly 1.0).  With ``global_b`` unknown the sample counts are summed
    over ranks on the device (no host sync)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 1.0
    world = dist.get_world_size()
    if global_b is not None:
        return float(local_b) * world / float(global_b)
    cnt = torch.tensor([float(local_b)], device=device, dtype=torch.float64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    return (float(local_b) * world / cnt).reshape(())        # 0-dim: multiplies an fp32 buffer without promoting it


def train_step(model, optimizer, batch: Dict[str, torch.Tensor], global_batch: int | None = None) -> torch.Tensor:
    """``batch``: this rank's device tensors keyed like ``collate_fn``'s output (dataset.py:165-186).
    ``optimizer``: ``optim.FusedAdam`` over ``model.parameters()``.  ``global_batch``: number of samples over all
    ranks when the caller knows it (skips the count all-reduce).  Returns the local loss (device scalar)."""
    optimizer.zero_grad()
    pm, ps, pe, pa = model(*[batch[k] for k in MODEL_INPUT_KEYS])
    loss = loss_fn(pm, batch["ym"], batch["sm"], batch["moment_mask"], ps, batch["ys"], batch["ss"], pe, batch["ye"], batch["se"],
                   pa, batch["ya"], batch["length_mask"])
    loss.backward()
    flat = optimizer.gather_grads()
    w = shard_weight(batch["video_features"].shape[0], global_batch, flat.device)
    if not (isinstance(w, float) and w == 1.0):
        flat.mul_(w)                         # uneven shards only
    scale = allreduce_mean_(flat, scale_here=False)
    optimizer.step(grad_scale=scale, gathered=True)
    return loss.detach()
