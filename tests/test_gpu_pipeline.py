"""ScoringPipeline (ingest launch + CUDA-graph replay, several batches in flight) must give
exactly what the eager drop-in module + compute_ious give."""
import pytest
import torch

from oracle import CONFIGS
from vml_b200 import synth
from vml_b200.evaluate import RecallAccumulator
from vml_b200.pipeline import INPUT_KEYS, ScoringPipeline

pytestmark = pytest.mark.gpu

from gpu_util import model_for  # noqa: E402


@pytest.mark.parametrize("prec,slots,graph", [("bf16", 2, True), ("fp32", 2, True), ("bf16", 3, False), ("bf16", 1, True)])
def test_pipeline_matches_eager(prec, slots, graph):
    cfg = CONFIGS["charadessta"]
    model = model_for(cfg, prec)
    batches = [synth.make_batch(cfg, 8, 300 + i) for i in range(7)]
    dev = [{k: v.cuda() for k, v in b.items()} for b in batches]
    acc = RecallAccumulator(torch.device("cuda"))
    eager = []
    for b in dev:
        out = model(*[b[k] for k in synth.MODEL_INPUT_KEYS], overlap=False)
        acc.update(out[0], out[1], out[2], b["moment_mask"], b["sm"])
        eager.append([o.clone() for o in out])
    pipe = ScoringPipeline(model, slots=slots, use_graph=graph)
    for i, b in enumerate(dev):
        ev, slot = pipe.submit({k: b[k] for k in INPUT_KEYS})
        ev.synchronize()
        for a, e in zip(slot.outputs[0], eager[i]):
            assert torch.equal(a, e), (i, "pipeline output differs from the eager module")
    assert torch.equal(pipe.counts.cpu(), acc.counts.cpu())
    assert pipe.result() == acc.result()


def test_pipeline_from_pinned_host_and_readback():
    cfg = CONFIGS["charadessta"]
    model = model_for(cfg, "bf16")
    batches = [synth.make_batch(cfg, 8, 400 + i) for i in range(6)]
    pinned = [{k: b[k].pin_memory() for k in INPUT_KEYS} for b in batches]
    acc = RecallAccumulator(torch.device("cuda"))
    for b in batches:
        d = {k: v.cuda() for k, v in b.items()}
        out = model(*[d[k] for k in synth.MODEL_INPUT_KEYS])
        acc.update(out[0], out[1], out[2], d["moment_mask"], d["sm"])
    pipe = ScoringPipeline(model, slots=2)
    rb = [torch.zeros(2, 4, dtype=torch.int64).pin_memory() for _ in range(len(batches))]
    evs = [pipe.submit(p, from_host=True, readback=rb[i])[0] for i, p in enumerate(pinned)]
    for ev in evs:
        ev.synchronize()
    assert torch.equal(sum(rb), acc.counts.cpu())          # per-step hits add up to the total
    assert torch.equal(pipe.counts.cpu(), acc.counts.cpu())


def test_overlap_matches_serial_bitwise():
    """Two-stream overlap inside a step changes scheduling only."""
    cfg = CONFIGS["tacos"]
    model = model_for(cfg, "bf16")
    b = {k: v.cuda() for k, v in synth.make_batch(cfg, 6, 77).items()}
    a = model(*[b[k] for k in synth.MODEL_INPUT_KEYS], overlap=False)
    a = [t.clone() for t in a]
    for _ in range(3):
        o = model(*[b[k] for k in synth.MODEL_INPUT_KEYS], overlap=True)
        for x, y in zip(a, o):
            assert torch.equal(x, y)
