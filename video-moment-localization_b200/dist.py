"""Multi-GPU plumbing: one process per GPU, (video, query) batches sharded by rank.

The hot path has no cross-sample op (no BatchNorm, no cross-sample
attention; batch-slice invariance is tested on the GPU), so ranks never exchange activations.
The only collectives are
  * evaluation: one all-reduce (sum) of the 8 R@n,IoU=m hit counters + the sample count --
    exactly the running sums the reference keeps in main.py:155-156,205-209 -- and, when rank 0
    must emit predictions, an all-gather of the per-sample top-k records ("score gather");
  * training (later round): the gradient all-reduce.
NCCL over NVLink/NVSwitch on GPUs; the same code runs on the gloo backend for CPU tests.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None, device: torch.device | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; initialises the default process
    group when WORLD_SIZE > 1 (nccl on CUDA, gloo otherwise; rendezvous on 127.0.0.1 by default)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n samples: the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """This rank's contiguous slice of every tensor of a collated batch (dataset.py:76-90 keys)."""
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def allreduce_recall(counts: torch.Tensor, num_samples: int) -> Tuple[torch.Tensor, int]:
    """Sum the [2,4] int64 hit counters and the sample count over ranks (no-op for world 1)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts, num_samples
    packed = torch.cat([counts.reshape(-1).to(torch.int64),
                        torch.tensor([num_samples], dtype=torch.int64, device=counts.device)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    return packed[:-1].reshape(counts.shape), int(packed[-1].item())


def gather_topk(top_idx: torch.Tensor, top_score: torch.Tensor, n_total: int):
    """All-gather per-sample top-k records (flat index int32, score f32) in rank order; shards may
    be uneven, so each rank pads to the largest shard.  Returns ([n_total,k], [n_total,k])."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return top_idx, top_score
    world = dist.get_world_size()
    k = top_idx.shape[1]
    biggest = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    pad = biggest - top_idx.shape[0]
    idx = torch.nn.functional.pad(top_idx, (0, 0, 0, pad), value=-1)
    sc = torch.nn.functional.pad(top_score, (0, 0, 0, pad))
    idx_all = [torch.empty_like(idx) for _ in range(world)]
    sc_all = [torch.empty_like(sc) for _ in range(world)]
    dist.all_gather(idx_all, idx)
    dist.all_gather(sc_all, sc)
    keep = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    return (torch.cat([t[:c] for t, c in zip(idx_all, keep)]).view(-1, k),
            torch.cat([t[:c] for t, c in zip(sc_all, keep)]).view(-1, k))


def recall_dict(counts: torch.Tensor, num_samples: int, normalize: bool = True) -> Dict[str, float]:
    """The reference's metric dict ('R@{n}, IoU={m}', utils.py:29), optionally / num_samples (main.py:209)."""
    host = counts.cpu()
    den = float(num_samples) if normalize and num_samples else 1.0
    return {f"R@{n_}, IoU={m_}": float(host[a, t].item()) / den
            for a, n_ in enumerate((1, 5)) for t, m_ in enumerate((0.1, 0.3, 0.5, 0.7))}
