"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Run on the B200:
    python -m pytest tests -m gpu -x -q
Tolerances are the ones north_star states: proposal scores within 1e-5 relative in the fp32
validation mode and 1e-2 relative in bf16; masked entries exactly 0; indices/counts exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import CONFIGS, init_params, smin_forward as oracle_forward
from oracle import metrics_oracle as mo
from vml_b200 import lib as L_
from vml_b200 import synth

pytestmark = pytest.mark.gpu

from gpu_util import dims_of, model_for, rel_err, scaled_err, to_dev, unpack  # noqa: E402

FP32_TOL = 1e-5   # north_star: fp32 validation mode
BF16_TOL = 1e-2   # north_star: bf16 mode


def _oracle(cfg, batch, params, inter=False):
    with torch.no_grad():
        return oracle_forward(params, cfg, *[batch[k] for k in synth.MODEL_INPUT_KEYS], return_intermediates=inter)


@pytest.mark.parametrize("name,B,seed", [("tiny", 5, 104), ("tiny_r2", 5, 105), ("charadessta", 4, 101),
                                         ("tacos", 3, 102), ("activitynet", 2, 103)])
def test_forward_fp32_matches_oracle_and_golden(name, B, seed, golden_dir):
    cfg = CONFIGS[name]
    params = init_params(cfg, 43)
    batch = synth.make_batch(cfg, B, seed)
    model = model_for(cfg, "fp32", params)
    out = model(*[batch[k].cuda() for k in synth.MODEL_INPUT_KEYS])
    ref = _oracle(cfg, batch, params)
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    for key, o, r in zip(("pm", "ps", "pe", "pa"), out, ref):
        assert o.dtype == torch.float32 and o.shape == r.shape
        assert rel_err(o, r) < FP32_TOL, (key, rel_err(o, r))
        assert rel_err(o, torch.from_numpy(g[key])) < FP32_TOL, ("golden", key)      # the reference's own output
        assert torch.equal(o.cpu() == 0, r == 0), f"{key}: masked entries must be exactly 0"


@pytest.mark.parametrize("name,B,seed", [("tiny", 5, 104), ("charadessta", 4, 101), ("tacos", 3, 102), ("activitynet", 2, 103)])
def test_forward_bf16_matches_oracle(name, B, seed):
    cfg = CONFIGS[name]
    params = init_params(cfg, 43)
    batch = synth.make_batch(cfg, B, seed)
    model = model_for(cfg, "bf16", params)
    out = model(*[batch[k].cuda() for k in synth.MODEL_INPUT_KEYS])
    ref = _oracle(cfg, batch, params)
    for key, o, r in zip(("pm", "ps", "pe", "pa"), out, ref):
        assert rel_err(o, r) < BF16_TOL, (key, rel_err(o, r))
        assert torch.equal(o.cpu() == 0, r == 0), key


@pytest.mark.parametrize("name,prec", [("tiny", "fp32"), ("tiny_r2", "fp32"), ("charadessta", "fp32"), ("tiny", "bf16"),
                                       ("charadessta", "bf16")])
def test_every_stage_matches_oracle(name, prec):
    """fv/fs/fw, pooled fc/fm/fb and each SMI layer's (cu, mu, bu), packed -> dense."""
    from vml_b200.smin import Workspace, pack_weights, smin_forward
    cfg = CONFIGS[name]
    params = init_params(cfg, 43)
    B = 4
    batch = synth.make_batch(cfg, B, 7)
    _, inter = _oracle(cfg, batch, params, inter=True)
    p = L_.PREC[prec]
    dims = dims_of(cfg)
    pk = pack_weights(params, dims, p, torch.device("cuda"))
    ws = Workspace(torch.device("cuda"))
    keep = {}
    smin_forward(pk, dims, p, ws, *[batch[k].cuda() for k in synth.MODEL_INPUT_KEYS], keep=keep)
    tol = 2e-6 if prec == "fp32" else 2e-2
    cells = keep["cells"]
    assert scaled_err(keep["fv"].float().view(B, cfg.T, cfg.D), inter["fv"]) < tol
    assert scaled_err(keep["fs"], inter["fs"]) < tol           # recurrence is fp32 in both modes; bf16 mode feeds it bf16 input projections
    assert scaled_err(keep["fw"], inter["fw"]) < tol
    for k in range(cfg.layers + 1):
        fc = unpack(keep[f"fc{k}"], cells, B, cfg.L, cfg.C * cfg.D, p).view(B, cfg.L, cfg.L, cfg.C, cfg.D)
        fm = unpack(keep[f"fm{k}"], cells, B, cfg.L, cfg.D, p)
        assert scaled_err(fc, inter[f"fc{k}"]) < tol * (1 + k), (k, "fc", scaled_err(fc, inter[f"fc{k}"]))
        assert scaled_err(fm, inter[f"fm{k}"]) < tol * (1 + k), (k, "fm", scaled_err(fm, inter[f"fm{k}"]))
        assert scaled_err(keep[f"fb{k}"], inter[f"fb{k}"]) < tol * (1 + k), (k, "fb")
        if prec == "fp32":   # invalid cells exactly zero (they are never stored; unpack zero-fills)
            assert torch.equal(fm.cpu() == 0, inter[f"fm{k}"] == 0)


@pytest.mark.parametrize("name,B,lo,hi", [("charadessta", 24, 1, 12), ("activitynet", 12, 2, 9), ("tacos", 16, 4, 20)])
def test_forward_bf16_short_videos(name, B, lo, hi):
    """Very short videos (1..few map rows): many samples share one 128-row tile of the fused
    content kernel (multi-group path), and samples with < 5 valid cells exist."""
    cfg = CONFIGS[name]
    params = init_params(cfg, 43)
    batch = synth.make_batch(cfg, B, 900 + B, nfeats_range=(lo, hi))
    ref = _oracle(cfg, batch, params)
    for prec, tol in (("bf16", BF16_TOL), ("fp32", FP32_TOL)):
        model = model_for(cfg, prec, params)
        out = model(*[batch[k].cuda() for k in synth.MODEL_INPUT_KEYS])
        for key, o, r in zip(("pm", "ps", "pe", "pa"), out, ref):
            assert rel_err(o, r) < tol, (prec, key, rel_err(o, r))
            assert torch.equal(o.cpu() == 0, r == 0), key


def test_batch_slice_invariance_fp32():
    """Samples are independent (invariant 4): the data-parallel split is exact."""
    cfg = CONFIGS["charadessta"]
    batch = synth.make_batch(cfg, 6, 11)
    model = model_for(cfg, "fp32")
    full = model(*[batch[k].cuda() for k in synth.MODEL_INPUT_KEYS])
    part = model(*[batch[k][2:5].cuda() for k in synth.MODEL_INPUT_KEYS])
    for a, b in zip(full, part):
        assert torch.equal(a[2:5], b)


def test_padding_invariance_fp32():
    """Garbage in clips >= nfeats and words >= len must not change the result (invariant 5)."""
    cfg = CONFIGS["charadessta"]
    batch = synth.make_batch(cfg, 4, 12)
    model = model_for(cfg, "fp32")
    clean = model(*[batch[k].cuda() for k in synth.MODEL_INPUT_KEYS])
    dirty = {k: v.clone() for k, v in batch.items()}
    vm = batch["video_mask"].bool().expand_as(dirty["video_features"])
    qm = batch["query_mask"].bool().expand_as(dirty["query_features"])
    dirty["video_features"][~vm] = 123.0
    dirty["query_features"][~qm] = -77.0
    out = model(*[dirty[k].cuda() for k in synth.MODEL_INPUT_KEYS])
    for a, b in zip(clean, out):
        assert torch.equal(a, b)


def test_same_seed_same_init_as_reference_contract():
    """Construction order mirrors the reference, state_dict keys/shapes match init_params (section 8b)."""
    cfg = CONFIGS["charadessta"]
    from vml_b200.smin import SMIN
    m = SMIN(*cfg.ctor_args())
    ref = init_params(cfg, 43)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(sd[k].shape == ref[k].shape for k in sd)


def test_cpu_input_raises():
    cfg = CONFIGS["tiny"]
    from vml_b200.smin import SMIN
    m = SMIN(*cfg.ctor_args())
    batch = synth.make_batch(cfg, 2, 1)
    with pytest.raises(L_.VmlError):
        m(*[batch[k] for k in synth.MODEL_INPUT_KEYS])


@pytest.mark.parametrize("variant", ["v1", "v2", "v3", "v4"])
@pytest.mark.parametrize("name,B,rng", [("charadessta", 48, None), ("charadessta", 24, (1, 12)), ("tacos", 9, None),
                                        ("tacos", 16, (4, 20)), ("activitynet", 5, None), ("activitynet", 12, (2, 9))])
def test_one_kernel_content_unit_against_split(name, B, rng, variant, monkeypatch):
    """vml_content_unit (fc tile resident in shared memory, one kernel per layer) against vml_content_in_attention +
    vml_content_out, over full tiles, ragged last tiles, multi-sample tiles and the skipped last-layer cu store.
    v1 (VML_CU_V1=1, the round-1 kernel) performs the same arithmetic in the same order: every layer's cu / fm / fb and
    the final scores are BIT-IDENTICAL.  v2 / v3 / v4 (v4 = default: streamed tile) add the residual on the tensor cores (fp32 accumulation
    inside the MMA instead of an FADD) and take mean_c from the rounded bf16 tile (v2: read back by the row warps, v3: by
    the tensor cores): cu within 1 bf16 ulp of the split path, the maps that follow within the drift that implies (all
    variants are separately held to the oracle at 1e-2 elsewhere)."""
    from vml_b200.smin import Workspace, pack_weights, smin_forward
    monkeypatch.delenv("VML_CU_V1", raising=False)
    monkeypatch.delenv("VML_CU_VARIANT", raising=False)
    if variant == "v1":
        monkeypatch.setenv("VML_CU_V1", "1")
    elif variant in ("v2", "v3"):
        monkeypatch.setenv("VML_CU_VARIANT", variant[1])
    cfg = CONFIGS[name]
    params = init_params(cfg, 43)
    batch = synth.make_batch(cfg, B, 1300 + B, **({"nfeats_range": rng} if rng else {}))
    dims = dims_of(cfg)
    pk = pack_weights(params, dims, L_.BF16, torch.device("cuda"))
    dev_in = [batch[k].cuda() for k in synth.MODEL_INPUT_KEYS]
    got, want = {}, {}
    out_f = smin_forward(pk, dims, L_.BF16, Workspace(torch.device("cuda")), *dev_in, keep=got)
    out_s = smin_forward(pk, dims, L_.BF16, Workspace(torch.device("cuda")), *dev_in, keep=want, split_content=True)
    n = int(batch["moment_mask"].sum())

    def same(a, b, what, layer):
        if variant == "v1":
            assert torch.equal(a, b), (layer, what)
            return
        a, b = a.float(), b.float()
        scale = b.abs().max().item()
        err = (a - b).abs().max().item()
        # bf16 ulp at the tensor's scale is 2^-8 * scale; first layer: one rounding flip; later layers inherit the drift
        assert err <= (2.0 ** -7) * scale * layer, (layer, what, err, scale)
        # (a large share of the elements may sit one rounding apart; what is bounded is how far)
        assert (a - b).abs().mean().item() <= (2.0 ** -10) * scale * layer, (layer, what, "mean drift")

    for k in range(1, cfg.layers + 1):
        same(got[f"fc{k}"][:n], want[f"fc{k}"][:n], "cu", k)
        same(got[f"fm{k}"][:n], want[f"fm{k}"][:n], "fm", k)
        same(got[f"fb{k}"], want[f"fb{k}"], "fb", k)
    # production path (no `keep`): last layer's cu store skipped, two-stream overlap
    model = model_for(cfg, "bf16", params)
    prod = model(*dev_in)
    prod_split = model(*dev_in, split_content=True)
    for a, b_, c in zip(prod, prod_split, out_s):
        assert torch.equal(b_, c)                          # the split path does not depend on how it is driven
        assert torch.equal(a, out_f[prod.index(a)] if False else a)
        if variant == "v1":
            assert torch.equal(a, b_)
        else:
            assert (a - b_).abs().max().item() < 2e-3      # sigmoid outputs: a few 1e-4 apart at most
    for a, f in zip(prod, out_f):
        assert torch.equal(a, f)                           # keep-mode (serial) == production (overlapped, skipped store)


@pytest.mark.parametrize("name,prec,B,kw", [("charadessta", "bf16", 9, {}), ("charadessta", "bf16", 6, {"nfeats_range": (1, 9)}),
                                            ("charadessta", "bf16", 5, {"full_length": True}), ("tacos", "bf16", 4, {}),
                                            ("activitynet", "bf16", 3, {}), ("activitynet", "bf16", 3, {"nfeats_range": (3, 40)}),
                                            ("charadessta", "fp32", 4, {}), ("tiny", "bf16", 5, {})])
def test_boundary_schedules_are_bit_identical(name, prec, B, kw, monkeypatch):
    """The schedule knobs of the boundary unit change WHEN things are loaded, never what is computed: rows staged by
    cp.async.bulk (default) vs through registers, D as a compile-time constant, the per-sample streaming kernel (default
    for L <= 16 in fast mode) vs the warp-per-row kernel.  Scores must agree bit for bit in every combination (the
    launchers read the knobs per call)."""
    from vml_b200.smin import Workspace, pack_weights, smin_forward
    cfg = CONFIGS[name]
    p, dims = L_.PREC[prec], dims_of(cfg)
    pk = pack_weights(init_params(cfg, 43), dims, p, torch.device("cuda"))
    batch = to_dev(synth.make_batch(cfg, B, 4242, **kw))
    outs = []
    for gate_bulk, gate_dt, stream_sample in (("1", "1", "1"), ("0", "1", "1"), ("1", "0", "0"), ("0", "0", "0"), ("1", "1", "0")):
        if True:
            monkeypatch.setenv("VML_GATE_BULK", gate_bulk)
            monkeypatch.setenv("VML_GATE_DT", gate_dt)              # D as a compile-time constant in the gate+rows kernel
            monkeypatch.setenv("VML_STREAM_SAMPLE", stream_sample)
            keep = {}
            out = smin_forward(pk, dims, p, Workspace(torch.device("cuda")), *[batch[k] for k in synth.MODEL_INPUT_KEYS], keep=keep)
            torch.cuda.synchronize()
            n_live = int(batch["moment_mask"].sum().item())          # rows past the live cell count are never written
            outs.append([o.clone() for o in out] + [keep[f"fb{cfg.layers}"].clone(), keep[f"fm{cfg.layers}"][:n_live].clone()])
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert torch.equal(x, y)


@pytest.mark.parametrize("name,B,kw", [("charadessta", 9, {}), ("charadessta", 64, {}), ("charadessta", 5, {"nfeats_range": (1, 9)}),
                                       ("charadessta", 3, {"full_length": True}), ("tacos", 6, {}), ("activitynet", 3, {}),
                                       ("tiny", 5, {})])
def test_ping_pong_content_unit_is_bit_identical(name, B, kw, monkeypatch):
    """content_unit_pp_kernel (two row-warp groups on alternate tiles, one Y accumulator, mean_c in 64-column halves, 8-box
    ring) computes exactly what content_unit_kernel<.., 4> computes: every layer's cu / fm / fb and the scores bit for bit,
    over single-tile, multi-tile, multi-sample-per-tile and ragged-last-tile maps."""
    from vml_b200.smin import Workspace, pack_weights, smin_forward
    cfg = CONFIGS[name]
    dims = dims_of(cfg)
    pk = pack_weights(init_params(cfg, 43), dims, L_.BF16, torch.device("cuda"))
    batch = to_dev(synth.make_batch(cfg, B, 5151, **kw))
    n = int(batch["moment_mask"].sum().item())
    res = {}
    for variant in ("4", "5"):
        monkeypatch.setenv("VML_CU_VARIANT", variant)
        keep = {}
        out = smin_forward(pk, dims, L_.BF16, Workspace(torch.device("cuda")), *[batch[k] for k in synth.MODEL_INPUT_KEYS], keep=keep)
        torch.cuda.synchronize()
        res[variant] = [o.clone() for o in out] + [keep[f"f{x}{k}"][:n if x != "b" else None].clone()
                                                   for k in range(1, cfg.layers + 1) for x in ("c", "m", "b")]
    for x, y in zip(res["4"], res["5"]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("name,prec,B,kw", [("charadessta", "bf16", 7, {}), ("charadessta", "fp32", 5, {}), ("tacos", "bf16", 3, {}),
                                            ("activitynet", "bf16", 2, {}), ("activitynet", "fp32", 2, {}),
                                            ("charadessta", "bf16", 6, {"nfeats_range": (1, 9)}), ("tiny", "fp32", 5, {}),
                                            ("tiny_r2", "fp32", 5, {})])
def test_span_pool_c4_kernel_matches_generic_kernel(name, prec, B, kw, monkeypatch):
    """span_pool_c4_kernel (C = 4, compile-time slice width, table-driven clip sizes, f32x2 math) against span_pool_kernel
    on the same inputs: pooled clips fc, moment features fm and boundary features fb bit for bit (same operations in the
    same order), ActivityNet's irregular windows and 1-clip videos included."""
    from vml_b200.smin import Workspace, pack_weights, smin_forward
    cfg = CONFIGS[name]
    p, dims = L_.PREC[prec], dims_of(cfg)
    pk = pack_weights(init_params(cfg, 43), dims, p, torch.device("cuda"))
    batch = to_dev(synth.make_batch(cfg, B, 6161, **kw))
    n = int(batch["moment_mask"].sum().item())
    got = {}
    for generic in (False, True):
        if generic:
            monkeypatch.setenv("VML_SPAN_GENERIC", "1")
        else:
            monkeypatch.delenv("VML_SPAN_GENERIC", raising=False)
        keep = {}
        smin_forward(pk, dims, p, Workspace(torch.device("cuda")), *[batch[k] for k in synth.MODEL_INPUT_KEYS], keep=keep)
        torch.cuda.synchronize()
        got[generic] = (keep["fc0"][:n].clone(), keep["fm0"][:n].clone(), keep["fb0"].clone())
    for x, y in zip(got[False], got[True]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("name,B,kw", [("charadessta", 9, {}), ("charadessta", 64, {}), ("tacos", 5, {}), ("activitynet", 3, {}),
                                       ("charadessta", 5, {"nfeats_range": (1, 9)}), ("tiny", 5, {})])
def test_transposing_gemm_epilogues_are_bit_identical(name, B, kw, monkeypatch):
    """EpiBiasT / EpiClipT / EpiMomentOutT (tcgen05 GEMM results re-laid through a shared-memory tile so that every warp
    store covers whole lines) perform the same additions in the same order as the register epilogues they replace
    (VML_EPI_DIRECT=1): clip projection, query states, every layer's fm and the scores agree bit for bit."""
    from vml_b200.smin import Workspace, pack_weights, smin_forward
    cfg = CONFIGS[name]
    dims = dims_of(cfg)
    pk = pack_weights(init_params(cfg, 43), dims, L_.BF16, torch.device("cuda"))
    batch = to_dev(synth.make_batch(cfg, B, 7171, **kw))
    n = int(batch["moment_mask"].sum().item())
    res = {}
    for direct in (False, True):
        if direct:
            monkeypatch.setenv("VML_EPI_DIRECT", "1")
        else:
            monkeypatch.delenv("VML_EPI_DIRECT", raising=False)
        keep = {}
        out = smin_forward(pk, dims, L_.BF16, Workspace(torch.device("cuda")), *[batch[k] for k in synth.MODEL_INPUT_KEYS], keep=keep)
        torch.cuda.synchronize()
        res[direct] = [o.clone() for o in out] + [keep["fv"].clone(), keep["fs"].clone(), keep["fw"].clone()] + \
                      [keep[f"fm{k}"][:n].clone() for k in range(1, cfg.layers + 1)]
    for x, y in zip(res[False], res[True]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("name,B,kw", [("charadessta", 9, {}), ("charadessta", 64, {}), ("tacos", 5, {}), ("activitynet", 3, {}),
                                       ("activitynet", 2, {"nfeats_range": (3, 40)}), ("tacos", 3, {"full_length": True}),
                                       ("charadessta", 5, {"nfeats_range": (1, 9)}), ("charadessta", 3, {"full_length": True}),
                                       ("tiny", 5, {})])
def test_pair_products_from_the_boundary_kernel_are_bit_identical(name, B, kw, monkeypatch):
    """vml_boundary_unit_pair (the per-sample streaming kernel also writes operand[n, 0:D] = bu_i * bu_j) against
    vml_boundary_unit + vml_moment_pair (VML_PAIR_SPLIT=1): every layer's fm / fb and the scores bit for bit."""
    from vml_b200.smin import Workspace, pack_weights, smin_forward
    cfg = CONFIGS[name]
    dims = dims_of(cfg)
    pk = pack_weights(init_params(cfg, 43), dims, L_.BF16, torch.device("cuda"))
    batch = to_dev(synth.make_batch(cfg, B, 8181, **kw))
    n = int(batch["moment_mask"].sum().item())
    res = {}
    for split in (False, True):
        if split:
            monkeypatch.setenv("VML_PAIR_SPLIT", "1")
        else:
            monkeypatch.delenv("VML_PAIR_SPLIT", raising=False)
        keep = {}
        out = smin_forward(pk, dims, L_.BF16, Workspace(torch.device("cuda")), *[batch[k] for k in synth.MODEL_INPUT_KEYS], keep=keep)
        torch.cuda.synchronize()
        res[split] = [o.clone() for o in out] + [keep[f"f{x}{k}"][:n if x == "m" else None].clone()
                                                 for k in range(1, cfg.layers + 1) for x in ("m", "b")]
    for x, y in zip(res[False], res[True]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("name,B,kw", [("charadessta", 9, {}), ("charadessta", 64, {}), ("tacos", 5, {}), ("activitynet", 3, {}),
                                       ("charadessta", 5, {"nfeats_range": (1, 9)}), ("charadessta", 3, {"full_length": True}),
                                       ("tiny", 5, {})])
def test_moment_gemm_with_generated_pair_operand_is_bit_identical(name, B, kw, monkeypatch):
    """vml_moment_out_gen (four generator warps write the bu_i * bu_j k-blocks of the A operand straight into the GEMM's
    pipeline stages) against the path that materialises the pair tensor (VML_MOMENT_GEN=0: boundary kernel / vml_moment_pair,
    then vml_moment_out): every layer's fm / fb and the scores bit for bit, over 128- and 256-wide tiles, ragged last tiles
    and tiles spanning several samples."""
    from vml_b200.smin import Workspace, pack_weights, smin_forward
    cfg = CONFIGS[name]
    dims = dims_of(cfg)
    pk = pack_weights(init_params(cfg, 43), dims, L_.BF16, torch.device("cuda"))
    batch = to_dev(synth.make_batch(cfg, B, 9191, **kw))
    n = int(batch["moment_mask"].sum().item())
    res = {}
    for gen in ("1", "0"):
        monkeypatch.setenv("VML_MOMENT_GEN", gen)
        keep = {}
        out = smin_forward(pk, dims, L_.BF16, Workspace(torch.device("cuda")), *[batch[k] for k in synth.MODEL_INPUT_KEYS], keep=keep)
        torch.cuda.synchronize()
        res[gen] = [o.clone() for o in out] + [keep[f"f{x}{k}"][:n if x == "m" else None].clone()
                                               for k in range(1, cfg.layers + 1) for x in ("m", "b")]
    for x, y in zip(res["1"], res["0"]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("name,prec,B,kw", [("tacos", "bf16", 5, {}), ("activitynet", "bf16", 3, {}), ("activitynet", "fp32", 2, {}),
                                            ("tacos", "fp32", 3, {}), ("activitynet", "bf16", 3, {"nfeats_range": (3, 40)})])
def test_per_sample_rows_kernel_matches_tiled_rows_kernel(name, prec, B, kw, monkeypatch):
    """boundary_rows_big_kernel (16 < L <= 64: the sample's rows resident in shared memory, full-depth score tiles per warp)
    against boundary_rows_mma_kernel (VML_ROWS_TILED=1: 16-row CTAs, split-K partials).  Same formulas; the score sums are
    associated differently, so the comparison is to rounding: 1e-5 of scale in the fp32 (3xTF32) mode, bf16-level in fast mode
    -- and both are held to the oracle by the parity tests above."""
    from vml_b200.smin import Workspace, pack_weights, smin_forward
    cfg = CONFIGS[name]
    p, dims = L_.PREC[prec], dims_of(cfg)
    pk = pack_weights(init_params(cfg, 43), dims, p, torch.device("cuda"))
    batch = to_dev(synth.make_batch(cfg, B, 1212, **kw))
    res = {}
    for tiled in (False, True):
        if tiled:
            monkeypatch.setenv("VML_ROWS_TILED", "1")
        else:
            monkeypatch.delenv("VML_ROWS_TILED", raising=False)
        keep = {}
        out = smin_forward(pk, dims, p, Workspace(torch.device("cuda")), *[batch[k] for k in synth.MODEL_INPUT_KEYS], keep=keep)
        torch.cuda.synchronize()
        res[tiled] = [o.clone() for o in out] + [keep[f"fb{k}"].clone() for k in range(1, cfg.layers + 1)]
    tol = 1e-5 if prec == "fp32" else 4e-3
    for x, y in zip(res[False], res[True]):
        assert scaled_err(x, y) < tol, scaled_err(x, y)
        assert torch.equal(x == 0, y == 0)
