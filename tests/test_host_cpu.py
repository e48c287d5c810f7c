"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares,
the drop-in module honours the reference's state_dict contract, host logic."""
import ctypes
import os

import pytest
import torch

from oracle import CONFIGS, init_params
from vml_b200 import lib as L_
from vml_b200 import synth


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(L_.LIB_PATH):
        from vml_b200 import build
        build.build()
    return ctypes.CDLL(L_.LIB_PATH)


def test_library_exports_every_declared_symbol(built_lib):
    names = L_.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/vml_b200.h but not exported"
    assert set(L_._SIGS) == set(names), "ctypes table and header disagree"


def test_library_loads_without_gpu_and_reports_version(built_lib):
    lib = L_.load()
    assert lib.vml_version() >= 1
    assert isinstance(lib.vml_last_error(), bytes)


@pytest.mark.parametrize("name", ["charadessta", "activitynet", "tacos", "tiny"])
def test_state_dict_contract(name):
    from vml_b200.smin import SMIN
    cfg = CONFIGS[name]
    m = SMIN(*cfg.ctor_args())
    ref = init_params(cfg, 43)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert len(sd) == 47 + 20 * (cfg.layers - 1)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    m.load_state_dict(ref, strict=True)
    assert sum(p.numel() for p in m.parameters()) == sum(v.numel() for v in ref.values())


def test_constructor_validates_like_the_reference_needs():
    from vml_b200.smin import SMIN
    with pytest.raises(ValueError):
        SMIN(64, 16, 4, 512, 128, 3, 1024, 13, 128)      # D != 2H breaks models.py:81
    with pytest.raises(ValueError):
        SMIN(60, 16, 4, 512, 128, 3, 1024, 13, 256)      # L must divide T


def test_no_cpu_fallback():
    from vml_b200.smin import SMIN
    from vml_b200.evaluate import compute_ious
    from vml_b200.losses import loss_fn
    cfg = CONFIGS["tiny"]
    b = synth.make_batch(cfg, 2, 1)
    m = SMIN(*cfg.ctor_args())
    with pytest.raises(L_.VmlError):
        m(*[b[k] for k in synth.MODEL_INPUT_KEYS])
    z = torch.zeros(2, cfg.L, cfg.L)
    with pytest.raises(L_.VmlError):
        compute_ious(z, z[:, 0], z[:, 0], b["moment_mask"], b["sm"])
    with pytest.raises(L_.VmlError):
        loss_fn(z, b["ym"], b["sm"], b["moment_mask"], z[:, 0], b["ys"], b["ss"], z[:, 0], b["ye"], b["se"], z[:, 0], b["ya"], b["length_mask"])


def test_synth_batch_properties():
    cfg = CONFIGS["charadessta"]
    b = synth.make_batch(cfg, 8, 5)
    r = cfg.T // cfg.L
    assert b["nfeats"][0] == cfg.T and b["nfeats"][1] % r != 0
    assert torch.equal(b["length_mask"].sum(1), torch.ceil(b["nfeats"].float() / r).long())
    assert torch.equal(b["moment_mask"], torch.triu(b["length_mask"][:, :, None] & b["length_mask"][:, None, :]))
    assert (b["video_features"][~b["video_mask"].bool().expand_as(b["video_features"])] == 0).all()
    assert (b["query_mask"].sum((1, 2)) >= 3).all()
    b2 = synth.make_batch(cfg, 8, 5)
    assert all(torch.equal(b[k], b2[k]) for k in b)
    assert (b["length_mask"].sum(1) >= 3).all()          # >= 5 valid cells: top-5 well defined


def test_dropin_modules_resolve_like_main_py_imports():
    """`from models import SMIN` / `from utils import compute_ious` (main.py:3,5) resolve to this
    implementation when the dropin directory precedes the reference on sys.path."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dropin = os.path.join(root, "video-moment-localization_b200", "dropin")
    code = ("from models import SMIN; from utils import compute_ious, get_tokens; import vml_b200.smin as s, vml_b200.evaluate as e;"
            "assert SMIN is s.SMIN and compute_ious is e.compute_ious; assert get_tokens('A man, walks.') == ['a','man','walks']; print('ok')")
    env = dict(os.environ, PYTHONPATH=dropin)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_pack_video_rows_round_trip():
    """Packed clip features (no padding rows) <-> the padded [B, T, d0] tensor of dataset.py:69-73."""
    import torch
    from vml_b200 import synth
    from vml_b200.configs import CONFIGS
    from vml_b200.pipeline import pack_query_rows, pack_video_rows, unpack_query_rows, unpack_video_rows
    cfg = CONFIGS["charadessta"]
    for kw in ({}, {"full_length": True}, {"nfeats_range": (1, 3)}):
        b = synth.make_batch(cfg, 6, 5, **kw)
        rows = pack_video_rows(b["video_features"], b["nfeats"])
        assert rows.shape == (int(b["nfeats"].clamp(max=cfg.T).sum()), cfg.d0)
        assert torch.equal(unpack_video_rows(rows, b["nfeats"], cfg.T), b["video_features"])
        qrows = pack_query_rows(b["query_features"], b["query_mask"])
        assert qrows.shape == (int(b["query_mask"].sum()), 300)
        assert torch.equal(unpack_query_rows(qrows, b["query_mask"], cfg.Nq), b["query_features"])


def test_packed_blob_layout(monkeypatch):
    """Layout contract between pack_host_batch(packed=True) and vml_ingest_packed: fixed-size tensors first, then the packed clip
    rows at a 256-byte boundary, then the packed word rows at the NEXT 256-byte boundary behind them (the kernel derives that
    address from nfeats); a staging area sized by _full_bytes holds any batch of the same shape."""
    import torch
    from vml_b200 import synth
    from vml_b200.configs import CONFIGS
    from vml_b200.pipeline import PACKED_KEYS, pack_host_batch
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)      # no CUDA here
    cfg = CONFIGS["charadessta"]
    for dt, es in ((None, 4), (torch.bfloat16, 2)):
        b = synth.make_batch(cfg, 6, 11)
        p = pack_host_batch(b, feature_dtype=dt, packed=True)
        blob = p["_blob"]
        base = blob.data_ptr()
        offs = {k: p[k].data_ptr() - base for k in PACKED_KEYS}
        assert all(o % 256 == 0 for o in offs.values())
        assert [k for k, _ in sorted(offs.items(), key=lambda kv: kv[1])] == list(PACKED_KEYS)
        rows = int(b["nfeats"].clamp(max=cfg.T).sum())
        words = int(b["query_mask"].sum())
        assert p["video_features"].shape == (rows, cfg.d0) and p["query_features"].shape == (words, 300)
        vbytes = rows * cfg.d0 * es
        assert offs["query_features"] == offs["video_features"] + (vbytes + 255) // 256 * 256
        assert blob.numel() == offs["query_features"] + (words * 300 * es + 255) // 256 * 256
        full = synth.make_batch(cfg, 6, 12, full_length=True)
        full["query_mask"][:] = 1
        assert pack_host_batch(full, feature_dtype=dt, packed=True)["_blob"].numel() == p["_full_bytes"]
        assert p["_rows_max"] == 6 * cfg.T and p["_q_shape"] == (6, cfg.Nq, 300)
