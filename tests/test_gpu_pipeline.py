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


@pytest.mark.parametrize("prec,slots,graph,coalesce", [("bf16", 2, True, 1), ("fp32", 2, True, 1), ("bf16", 3, False, 1),
                                                        ("bf16", 1, True, 1), ("bf16", 2, True, 3), ("fp32", 1, True, 2)])
def test_pipeline_matches_eager(prec, slots, graph, coalesce):
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
    pipe = ScoringPipeline(model, slots=slots, use_graph=graph, coalesce=coalesce)
    group = []
    for i, b in enumerate(dev):
        group.append((i, pipe.submit({k: b[k] for k in INPUT_KEYS})))
        if len(group) == coalesce or i == len(dev) - 1:
            pipe.flush()                                  # (only the ragged last group actually needs it)
            for j, t in group:
                t.synchronize()
                for a, e in zip(t.slot.outputs[0], eager[j]):
                    rows = slice(t.index * 8, t.index * 8 + 8)
                    assert torch.equal(a[rows], e), (j, "pipeline output differs from the eager module")
            group = []
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
    pipe = ScoringPipeline(model, slots=2, coalesce=2)
    rb = [torch.zeros(2, 4, dtype=torch.int64).pin_memory() for _ in range(len(batches))]
    tickets = [pipe.submit(p, from_host=True, readback=rb[i]) for i, p in enumerate(pinned)]
    pipe.flush()
    for t in tickets:
        t.synchronize()
    assert torch.equal(sum(rb), acc.counts.cpu())          # per-step hits add up to the total
    assert torch.equal(pipe.counts.cpu(), acc.counts.cpu())


def test_pipeline_single_blob_h2d():
    """pack_host_batch: one pinned blob per batch, one H2D copy; same counts as the eager path."""
    from vml_b200.pipeline import pack_host_batch
    cfg = CONFIGS["charadessta"]
    model = model_for(cfg, "bf16")
    batches = [synth.make_batch(cfg, 8, 500 + i) for i in range(5)]
    acc = RecallAccumulator(torch.device("cuda"))
    for b in batches:
        d = {k: v.cuda() for k, v in b.items()}
        out = model(*[d[k] for k in synth.MODEL_INPUT_KEYS])
        acc.update(out[0], out[1], out[2], d["moment_mask"], d["sm"])
    pipe = ScoringPipeline(model, slots=2, coalesce=2)
    for b in batches:
        pipe.submit(pack_host_batch(b), from_host=True)
    assert pipe.result() == acc.result()


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


def test_bf16_host_features_are_bit_identical_in_bf16_mode():
    """Half-width host features (vml_ingest_bf16): the fp32 -> bf16 rounding is round-to-nearest on either side of the
    H2D copy, so in bf16 precision every score is bit-identical to the fp32-host path -- through the module API and
    through the pipeline's blob path (generic first pass + recorded fast path), counts included."""
    from vml_b200.pipeline import pack_host_batch
    cfg = CONFIGS["charadessta"]
    model = model_for(cfg, "bf16")
    batches = [synth.make_batch(cfg, 8, 600 + i) for i in range(7)]
    d = {k: v.cuda() for k, v in batches[0].items()}
    want = model(*[d[k] for k in synth.MODEL_INPUT_KEYS])
    want = [t.clone() for t in want]
    d16 = dict(d, video_features=d["video_features"].bfloat16(), query_features=d["query_features"].bfloat16())
    got = model(*[d16[k] for k in synth.MODEL_INPUT_KEYS])
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    pipes = []
    for dt in (None, torch.bfloat16):
        pipe = ScoringPipeline(model, slots=2, coalesce=2)
        for b in batches:
            pipe.submit(pack_host_batch(b, feature_dtype=dt), from_host=True)
        pipes.append(pipe)
    r0, r1 = pipes[0].result(), pipes[1].result()          # result() flushes the partial last group and synchronises
    assert r0 == r1 and torch.equal(pipes[0].counts.cpu(), pipes[1].counts.cpu())
    blob16 = pack_host_batch(batches[0], feature_dtype=torch.bfloat16)["_blob"].numel()
    assert blob16 < 0.55 * pack_host_batch(batches[0])["_blob"].numel()


def test_fp32_mode_accepts_bf16_host_features():
    """fp32 validation mode with bf16 inputs == fp32 mode on the same values widened to float32."""
    cfg = CONFIGS["tiny"]
    model = model_for(cfg, "fp32")
    d = {k: v.cuda() for k, v in synth.make_batch(cfg, 5, 610).items()}
    v16, q16 = d["video_features"].bfloat16(), d["query_features"].bfloat16()
    a = model(v16, d["video_mask"], q16, d["query_mask"], d["length_mask"], d["moment_mask"])
    b = model(v16.float(), d["video_mask"], q16.float(), d["query_mask"], d["length_mask"], d["moment_mask"])
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.parametrize("from_host", [False, True])
def test_pipeline_scores_a_ragged_tail_batch(from_host):
    """len(split) % batch != 0 (the last batch of main.py:168-189): a batch of another size is scored eagerly into
    the same counters, before and after full groups, with its own read-back."""
    from vml_b200.pipeline import pack_host_batch
    cfg = CONFIGS["charadessta"]
    model = model_for(cfg, "bf16")
    sizes = [8, 8, 8, 5, 8, 8, 3]
    batches = [synth.make_batch(cfg, n, 700 + i) for i, n in enumerate(sizes)]
    acc = RecallAccumulator(torch.device("cuda"))
    per_batch = []
    for b in batches:
        d = {k: v.cuda() for k, v in b.items()}
        out = model(*[d[k] for k in synth.MODEL_INPUT_KEYS])
        one = RecallAccumulator(torch.device("cuda"))
        one.update(out[0], out[1], out[2], d["moment_mask"], d["sm"])
        acc.update(out[0], out[1], out[2], d["moment_mask"], d["sm"])
        per_batch.append(one.counts.cpu())
    pipe = ScoringPipeline(model, slots=2, coalesce=2)
    rb = [torch.zeros(2, 4, dtype=torch.int64).pin_memory() for _ in sizes]
    tickets = []
    for i, b in enumerate(batches):
        src = pack_host_batch(b) if from_host else {k: v.cuda() for k, v in b.items()}
        tickets.append(pipe.submit(src, from_host=from_host, readback=rb[i]))
    assert pipe.result() == acc.result() and pipe.num_samples == sum(sizes)
    for t in tickets:
        t.synchronize()
    for got, want in zip(rb, per_batch):
        assert torch.equal(got, want)


@pytest.mark.parametrize("from_host", [True, False])
def test_compact_batches_with_device_built_labels_equal_full_batches(from_host):
    """SURVEY 8(f)-3 wired into the pipeline: a batch that carries only features, the word mask and (times, duration, nfeats)
    -- masks and the IoU map built by vml_make_labels on the device -- scores bit-identically to the full 7-tensor batch, and
    moves fewer bytes over PCIe (incl. a ragged tail batch)."""
    from vml_b200.pipeline import COMPACT_KEYS, pack_host_batch
    cfg = CONFIGS["charadessta"]
    model = model_for(cfg, "bf16")
    batches = [synth.make_batch(cfg, 8, 700 + i) for i in range(5)] + [synth.make_batch(cfg, 3, 799)]
    ref = ScoringPipeline(model, slots=2, coalesce=2)
    for b in batches:
        ref.submit({k: b[k].cuda() for k in INPUT_KEYS})
    want = ref.result(normalize=False)
    pipe = ScoringPipeline(model, slots=2, coalesce=2)
    tickets = []
    for b in batches:
        if from_host:
            full, compact = pack_host_batch(b), pack_host_batch(b, compact=True)
            assert compact["_blob"].numel() < full["_blob"].numel()
            tickets.append(pipe.submit(compact, from_host=True))
        else:
            tickets.append(pipe.submit({k: b[k].cuda() for k in COMPACT_KEYS}))
    assert pipe.result(normalize=False) == want
    assert torch.equal(pipe.counts, ref.counts)
    # the first group's scores, bit for bit
    a = ScoringPipeline(model, slots=1, coalesce=1)
    b_ = ScoringPipeline(model, slots=1, coalesce=1)
    t1 = a.submit({k: batches[0][k].cuda() for k in INPUT_KEYS}); a.flush(); t1.synchronize()
    src = pack_host_batch(batches[0], compact=True) if from_host else {k: batches[0][k].cuda() for k in COMPACT_KEYS}
    t2 = b_.submit(src, from_host=from_host); b_.flush(); t2.synchronize()
    for x, y in zip(t1.slot.outputs[0], t2.slot.outputs[0]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("prec,feature_dtype", [("bf16", None), ("bf16", torch.bfloat16), ("fp32", None)])
def test_packed_batches_equal_padded_batches(prec, feature_dtype):
    """Clip features shipped WITHOUT their all-zero padding rows (dataset.py:69-73 pads to T): vml_ingest_packed re-creates the
    padding on the device.  Scores, hit counters and per-step read-backs are bit-identical to the padded compact batches;
    the blob shrinks to ~mean(nfeats) / T.  Includes full-length videos, a 1-clip video, a tail batch of another size and a
    switch from padded to packed blobs on the same staging ring."""
    from vml_b200.pipeline import pack_host_batch
    cfg = CONFIGS["charadessta"]
    model = model_for(cfg, prec)
    batches = [synth.make_batch(cfg, 8, 900 + i) for i in range(4)]
    batches.append(synth.make_batch(cfg, 8, 950, full_length=True))
    batches.append(synth.make_batch(cfg, 8, 951, nfeats_range=(1, 5)))
    batches += [synth.make_batch(cfg, 8, 960 + i) for i in range(2)]
    batches.append(synth.make_batch(cfg, 3, 999))
    ref = ScoringPipeline(model, slots=2, coalesce=2)
    rb_ref = [torch.zeros(2, 4, dtype=torch.int64).pin_memory() for _ in batches]
    for i, b in enumerate(batches):
        ref.submit(pack_host_batch(b, feature_dtype=feature_dtype, compact=True), from_host=True, readback=rb_ref[i])
    want = ref.result(normalize=False)
    pipe = ScoringPipeline(model, slots=2, coalesce=2)
    rb = [torch.zeros(2, 4, dtype=torch.int64).pin_memory() for _ in batches]
    padded_bytes = packed_bytes = 0
    for rounds in range(2):                   # second round: fast path (recorded ingest launches) + reused staging areas
        for i, b in enumerate(batches):
            pk = pack_host_batch(b, feature_dtype=feature_dtype, packed=True)
            padded_bytes += pack_host_batch(b, feature_dtype=feature_dtype, compact=True)["_blob"].numel()
            packed_bytes += pk["_blob"].numel()
            pipe.submit(pk, from_host=True, readback=rb[i])
    got = pipe.result(normalize=False)
    assert got == {k: 2 * v for k, v in want.items()}
    for x, y in zip(rb, rb_ref):
        assert torch.equal(x, y)
    assert packed_bytes < 0.9 * padded_bytes
    # scores of one batch, bit for bit (and a padded blob followed by a packed one on the same staging area)
    a = ScoringPipeline(model, slots=1, coalesce=1)
    t1 = a.submit(pack_host_batch(batches[0], feature_dtype=feature_dtype, compact=True), from_host=True); a.flush(); t1.synchronize()
    want_out = [x.clone() for x in t1.slot.outputs[0]]
    for _ in range(3):
        t2 = a.submit(pack_host_batch(batches[0], feature_dtype=feature_dtype, packed=True), from_host=True); a.flush(); t2.synchronize()
        for x, y in zip(want_out, t2.slot.outputs[0]):
            assert torch.equal(x, y)
    # a device-resident packed batch is re-padded and scored as usual
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in pack_host_batch(batches[0], packed=True).items() if k != "_blob"}
    b3 = ScoringPipeline(model, slots=1, coalesce=1)
    t3 = b3.submit(dev); b3.flush(); t3.synchronize()
    if feature_dtype is None:
        for x, y in zip(want_out, t3.slot.outputs[0]):
            assert torch.equal(x, y)
