"""vml_make_labels (device-side dataset.py:95-127,139-158) against the batches of synth.make_batch, whose label
functions are pinned bit-exactly to the reference's dataset.py by tests/test_oracle_golden.py."""
import pytest
import torch

from oracle import CONFIGS
from vml_b200 import synth
from vml_b200.labels import make_labels

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,B,seed,rng", [("charadessta", 64, 3, None), ("tacos", 33, 4, None), ("activitynet", 16, 5, None),
                                             ("tiny", 7, 6, None), ("charadessta", 40, 7, (1, 12)), ("tiny_r2", 9, 8, (1, 16))])
def test_labels_match_reference_formulas(name, B, seed, rng):
    cfg = CONFIGS[name]
    b = synth.make_batch(cfg, B, seed, features=False, **({"nfeats_range": rng} if rng else {}))
    got = make_labels(b["times"].cuda(), b["duration"].cuda(), b["nfeats"].cuda(), cfg.T, cfg.L)
    for k in ("sm", "ym", "ya", "length_mask", "moment_mask", "video_mask"):         # IEEE +,-,*,/ only: bit-exact
        assert got[k].dtype == b[k].dtype and got[k].shape == b[k].shape, k
        assert torch.equal(got[k].cpu(), b[k]), k
    for k, yk in (("ss", "ys"), ("se", "ye")):                                      # exp: CUDA expf vs SLEEF, <= 2 ulp
        g, w = got[k].cpu(), b[k]
        assert torch.allclose(g, w, rtol=3e-7, atol=1e-37), k
        differ = got[yk].cpu() != b[yk]
        assert not differ.any() or bool(((w[differ] - 0.5).abs() < 1e-6).all()), yk   # a threshold flip needs |s - 0.5| ~ 1 ulp
