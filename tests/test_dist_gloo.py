"""world_size-2 gloo tests (CPU) of the data-parallel host logic: contiguous batch sharding,
hit-counter all-reduce and top-k score gather reproduce the single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import CONFIGS, init_params, smin_forward
from oracle import metrics_oracle as mo
from vml_b200 import dist as vdist
from vml_b200 import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _counts(pm, ps, pe, mask, sm):
    m = mo.compute_ious(pm, ps, pe, mask, sm)
    return torch.tensor([[m[f"R@{n}, IoU={t}"] for t in (0.1, 0.3, 0.5, 0.7)] for n in (1, 5)], dtype=torch.int64)


def _worker(rank, world, port, n_total, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = vdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    cfg = CONFIGS["tiny"]
    params = init_params(cfg, 43)
    full = synth.make_batch(cfg, n_total, 77)
    mine = vdist.shard_batch(full, rank, world)
    with torch.no_grad():
        pm, ps, pe, pa = smin_forward(params, cfg, *[mine[k] for k in synth.MODEL_INPUT_KEYS])
    counts = _counts(pm, ps, pe, mine["moment_mask"], mine["sm"])
    total, n = vdist.allreduce_recall(counts, pm.shape[0])
    scores = mo.proposal_scores(pm, ps, pe, mine["moment_mask"])
    top = mo.topk_lowest_index(scores, 5).to(torch.int32)
    idx_all, sc_all = vdist.gather_topk(top, torch.gather(scores, 1, top.long()), n_total)
    if rank == 0:
        torch.save({"counts": total, "n": n, "idx": idx_all, "score": sc_all}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 8])
def test_sharded_eval_equals_single_process(tmp_path, n_total):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), n_total, out), nprocs=2, join=True)
    got = torch.load(out)
    cfg = CONFIGS["tiny"]
    params = init_params(cfg, 43)
    full = synth.make_batch(cfg, n_total, 77)
    with torch.no_grad():
        pm, ps, pe, pa = smin_forward(params, cfg, *[full[k] for k in synth.MODEL_INPUT_KEYS])
    want = _counts(pm, ps, pe, full["moment_mask"], full["sm"])
    assert got["n"] == n_total
    assert torch.equal(got["counts"], want)
    scores = mo.proposal_scores(pm, ps, pe, full["moment_mask"])
    top = mo.topk_lowest_index(scores, 5)
    assert torch.equal(got["idx"].long(), top)
    # samples are independent up to fp reassociation of batched matmuls (SURVEY invariant 4)
    assert torch.allclose(got["score"], torch.gather(scores, 1, top), rtol=1e-5, atol=0)
    assert vdist.recall_dict(got["counts"], got["n"])["R@5, IoU=0.1"] == want[1, 0].item() / n_total


def test_shard_range_partitions():
    for n in (0, 1, 5, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            spans = [vdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_collectives_are_noops():
    c = torch.arange(8, dtype=torch.int64).view(2, 4)
    total, n = vdist.allreduce_recall(c, 5)
    assert torch.equal(total, c) and n == 5


def _grad_worker(rank, world, port, n_total, out):
    """Data-parallel training math (trainer.allreduce_mean_): every rank back-propagates the loss of its own slice
    (autograd of the CPU oracle stands in for the CUDA backward), flattens the gradients in parameter order like
    FusedAdam.flat_grad, and one all-reduce yields the global-batch gradient."""
    from vml_b200.trainer import allreduce_mean_, shard_weight
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    vdist.init_from_env(backend="gloo")
    cfg = CONFIGS["tiny"]
    params = {k: v.clone().double().requires_grad_(True) for k, v in init_params(cfg, 43).items()}
    mine = {k: (v.double() if v.is_floating_point() else v)
            for k, v in vdist.shard_batch(synth.make_batch(cfg, n_total, 78), rank, world).items()}
    pm, ps, pe, pa = smin_forward(params, cfg, *[mine[k] for k in synth.MODEL_INPUT_KEYS])
    loss = mo.loss_fn(pm, mine["ym"], mine["sm"], mine["moment_mask"], ps, mine["ys"], mine["ss"], pe, mine["ye"], mine["se"], pa,
                      mine["ya"], mine["length_mask"])
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in params.values()])
    # B_local / B_global weighting (uneven shards): global size summed over ranks == global size given
    nb = mine["video_features"].shape[0]
    w = float(shard_weight(nb, None, flat.device))
    assert abs(w - shard_weight(nb, n_total, flat.device)) < 1e-12
    assert (w == 1.0) == (n_total % world == 0)
    flat.mul_(float(w))
    assert allreduce_mean_(flat) == 1.0
    if rank == 0:
        torch.save(flat, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [6, 7])             # 6: equal slices; 7: ragged last batch (4 + 3), weighted by B_r / B
def test_data_parallel_gradient_equals_global_batch_gradient(tmp_path, n_total):
    out = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(2, _free_port(), n_total, out), nprocs=2, join=True)
    got = torch.load(out)
    cfg = CONFIGS["tiny"]
    params = {k: v.clone().double().requires_grad_(True) for k, v in init_params(cfg, 43).items()}
    full = {k: (v.double() if v.is_floating_point() else v) for k, v in synth.make_batch(cfg, n_total, 78).items()}
    pm, ps, pe, pa = smin_forward(params, cfg, *[full[k] for k in synth.MODEL_INPUT_KEYS])
    mo.loss_fn(pm, full["ym"], full["sm"], full["moment_mask"], ps, full["ys"], full["ss"], pe, full["ye"], full["se"], pa,
               full["ya"], full["length_mask"]).backward()
    want = torch.cat([p.grad.reshape(-1) for p in params.values()])
    assert got.shape == want.shape
    assert torch.isfinite(want).all() and want.abs().max() > 0
    assert (got - want).abs().max() <= 1e-9 * want.abs().max()


def test_allreduce_mean_single_process_is_noop():
    from vml_b200.trainer import allreduce_mean_
    g = torch.arange(5, dtype=torch.float32)
    assert allreduce_mean_(g, scale_here=False) == 1.0 and torch.equal(g, torch.arange(5, dtype=torch.float32))
