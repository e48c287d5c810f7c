"""Parity at the shapes bench.py times (round-1 verdict: oracle comparisons used B = 2..5 only).

  * CUDA path vs the REFERENCE's own outputs (tests/golden/*_b64 / *_b16.npz, made by tools/make_golden.py from the
    unmodified reference; charadessta_b64 is the bench's first batch) and vs the oracle, both arithmetic modes, B = 64 / 16;
  * logits (not sigmoid outputs) and last-layer maps in bf16 mode, so the check is not flattened by the sigmoid;
  * ``ScoringPipeline(coalesce=4)`` -- the 256-query pass the bench runs -- against the ORACLE, not against the eager module;
  * R@n,IoU=m counts equal to the reference's in fp32 mode on batches whose counts are non-zero.
"""
import os

import numpy as np
import pytest
import torch

from oracle import CONFIGS, init_params, smin_forward as oracle_forward
from oracle import metrics_oracle as mo
from vml_b200 import lib as L_
from vml_b200 import synth
from vml_b200.evaluate import compute_ious, score_topk_recall
from vml_b200.pipeline import INPUT_KEYS, ScoringPipeline

pytestmark = pytest.mark.gpu

from gpu_util import dims_of, model_for, rel_err, scaled_err, unpack  # noqa: E402

FP32_TOL, BF16_TOL = 1e-5, 1e-2          # north_star
BF16_LOGIT_TOL = 5e-3                    # |logit - logit_ref|: 8x tighter than what 1e-2 relative on p ~ 0.5 allows (4e-2); measured 1e-3
GOLDEN = [("charadessta_b64", "charadessta"), ("tacos_b64", "tacos"), ("activitynet_b16", "activitynet"), ("activitynet_b4", "activitynet")]


def _logit(p):
    p = p.double()
    return torch.log(p / (1 - p))


@pytest.mark.parametrize("gname,cname", GOLDEN)
def test_bench_shape_scores_match_reference_outputs(gname, cname, golden_dir):
    g = np.load(os.path.join(golden_dir, f"{gname}.npz"))
    cfg = CONFIGS[cname]
    B, seed = int(g["B"]), int(g["seed"])
    params = init_params(cfg, 43)
    batch = synth.make_batch(cfg, B, seed)
    dev = {k: v.cuda() for k, v in batch.items()}
    ref = [torch.from_numpy(g[k]) for k in ("pm", "ps", "pe", "pa")]
    ref64 = [torch.from_numpy(g[k]) for k in ("pm64", "ps64", "pe64", "pa64")]
    want_counts = dict(zip(g["metric_keys"].tolist(), g["metric_vals"].tolist()))
    assert sum(want_counts.values()) > 0
    for prec, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL)):
        model = model_for(cfg, prec, params)
        out = model(*[dev[k] for k in synth.MODEL_INPUT_KEYS])
        for key, o, r, r64 in zip(("pm", "ps", "pe", "pa"), out, ref, ref64):
            assert rel_err(o, r) < tol, (prec, key, rel_err(o, r))
            assert torch.equal(o.cpu() == 0, r == 0), (prec, key)
            valid = r != 0
            lerr = (_logit(o.cpu()[valid]) - _logit(r64[valid])).abs().max().item()
            assert lerr < (BF16_LOGIT_TOL if prec == "bf16" else 2e-5), (prec, key, "logit", lerr)
        if prec == "fp32":       # end-to-end exactness is defined in fp32 mode
            top = score_topk_recall(out[0], out[1], out[2], dev["moment_mask"], dev["sm"])[0].cpu().long()
            same = (top == torch.from_numpy(g["ref_topk"])).all(1)
            # a top-6 gap below the fp32 noise floor may legitimately swap two neighbours; the fixtures have none
            assert bool(same.all()), f"{int((~same).sum())} samples differ from the reference's torch.topk order"
            assert dict(compute_ious(out[0], out[1], out[2], dev["moment_mask"], dev["sm"])) == want_counts
    # kernel-isolated: our metric kernel on the reference's own score tensors
    got = compute_ious(ref[0].cuda(), ref[1].cuda(), ref[2].cuda(), dev["moment_mask"], dev["sm"])
    assert dict(got) == want_counts


@pytest.mark.parametrize("name,B", [("charadessta", 64), ("tacos", 64), ("activitynet", 16)])
def test_bench_shape_last_layer_maps_match_oracle_bf16(name, B):
    """Last SMI layer's fm / fb (what the scoring heads read) at the bench's batch size, scaled error, both modes."""
    from vml_b200.smin import Workspace, pack_weights, smin_forward
    cfg = CONFIGS[name]
    params = init_params(cfg, 43)
    batch = synth.make_batch(cfg, B, 1000 + B)
    with torch.no_grad():
        _, inter = oracle_forward(params, cfg, *[batch[k] for k in synth.MODEL_INPUT_KEYS], return_intermediates=True)
    dims = dims_of(cfg)
    for prec, tol in (("bf16", 2e-2), ("fp32", 1e-5)):
        p = L_.PREC[prec]
        pk = pack_weights(params, dims, p, torch.device("cuda"))
        keep = {}
        smin_forward(pk, dims, p, Workspace(torch.device("cuda")), *[batch[k].cuda() for k in synth.MODEL_INPUT_KEYS], keep=keep)
        k = cfg.layers
        fm = unpack(keep[f"fm{k}"], keep["cells"], B, cfg.L, cfg.D, p)
        assert scaled_err(fm, inter[f"fm{k}"]) < tol, (prec, "fm", scaled_err(fm, inter[f"fm{k}"]))
        assert scaled_err(keep[f"fb{k}"], inter[f"fb{k}"]) < tol, (prec, "fb", scaled_err(keep[f"fb{k}"], inter[f"fb{k}"]))
        fc = unpack(keep[f"fc{k}"], keep["cells"], B, cfg.L, cfg.C * cfg.D, p).view(B, cfg.L, cfg.L, cfg.C, cfg.D)
        assert scaled_err(fc, inter[f"fc{k}"]) < tol, (prec, "cu", scaled_err(fc, inter[f"fc{k}"]))


@pytest.mark.parametrize("name,prec", [("charadessta", "bf16"), ("charadessta", "fp32"), ("tacos", "bf16")])
def test_coalesced_256_query_pass_matches_oracle(name, prec):
    """The bench's pass: 4 submitted batches of 64 scored by one graph replay.  Scores vs the oracle at the north_star
    tolerance; in fp32 mode the pipeline's accumulated R@n,IoU=m counters equal the oracle's sums."""
    cfg = CONFIGS[name]
    params = init_params(cfg, 43)
    model = model_for(cfg, prec, params)
    batches = [synth.make_batch(cfg, 64, 2100 + i) for i in range(4)]
    pipe = ScoringPipeline(model, slots=2, coalesce=4)
    tickets = [pipe.submit({k: b[k].cuda() for k in INPUT_KEYS}) for b in batches]
    pipe.flush()
    total = None
    tol = BF16_TOL if prec == "bf16" else FP32_TOL
    for b, t in zip(batches, tickets):
        t.synchronize()
        with torch.no_grad():
            ref = oracle_forward(params, cfg, *[b[k] for k in synth.MODEL_INPUT_KEYS])
        rows = slice(t.index * 64, t.index * 64 + 64)
        for key, o, r in zip(("pm", "ps", "pe", "pa"), t.slot.outputs[0], ref):
            assert rel_err(o[rows], r) < tol, (key, rel_err(o[rows], r))
            assert torch.equal(o[rows].cpu() == 0, r == 0), key
        m = mo.compute_ious(ref[0], ref[1], ref[2], b["moment_mask"], b["sm"])
        total = m if total is None else {k: total[k] + m[k] for k in m}
    if prec == "fp32":
        assert pipe.result(normalize=False) == total
    else:        # bf16 reorders near-ties: counts are close, not exact; exactness is asserted kernel-isolated
        got = pipe.result(normalize=False)
        assert all(abs(got[k] - total[k]) <= 0.05 * 256 + 1 for k in total), (got, total)
