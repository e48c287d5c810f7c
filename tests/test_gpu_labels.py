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


def _reference_dataset(cfg, split):
    from baseline import loader as bl
    if not bl.available():
        pytest.skip("baseline/_ref not installed")
    ns = bl.load_reference()
    ds = ns.dataset.AbstractDataset.__new__(ns.dataset.AbstractDataset)
    ds.T, ds.L, ds.split = cfg.T, cfg.L, split
    return ds


@pytest.mark.parametrize("name,split", [("charadessta", "test"), ("charadessta", "train"), ("activitynet", "test"), ("tiny", "train")])
def test_sample_clips_matches_reference_get_fixed_length_features(name, split):
    """vml_sample_clips vs the reference's own ``get_fixed_length_features`` (dataset.py:40-74): sampled clip rows bit-exact,
    nfeats / start_index / end_index equal, for videos shorter than, equal to and (much) longer than T; the training
    split's random start offset is drawn like the reference does (np.random.randint) and passed in."""
    import numpy as np
    from vml_b200.labels import sample_clips
    cfg = CONFIGS[name]
    ds = _reference_dataset(cfg, split)
    T, d0 = cfg.T, 24
    rng = np.random.default_rng(5)
    lens = [1, 2, T - 1, T, T + 1, 2 * T, 2 * T + 1, 3 * T - 1, 5 * T + 3, int(1.5 * T), T // 2, 7 * T] + list(rng.integers(1, 6 * T, 20))
    raws, spos, sp_n, ep_n, want = [], [], [], [], []
    for i, n in enumerate(lens):
        n = int(n)
        feat = rng.standard_normal((n, d0)).astype(np.float32)
        a = float(rng.uniform(0.0, 0.6))
        b = float(min(1.0, a + rng.uniform(0.05, 0.4)))
        np.random.seed(100 + i)                       # the reference draws its start offset from the global numpy RNG
        out, nf, si, ei = ds.get_fixed_length_features(feat, a, b)
        np.random.seed(100 + i)
        if split == "train":                          # dataset.py:44-49
            stride = 1.0 if n <= T else n * 1.0 / T
            random_end = -0.5 + stride
            if random_end == np.floor(random_end):
                random_end = random_end - 1.0
            spos.append(int(np.random.randint(0, random_end + 1)))
        else:
            spos.append(0)
        raws.append(feat); sp_n.append(a); ep_n.append(b)
        want.append((torch.FloatTensor(out), nf, si, ei))
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64)
    got = sample_clips(torch.from_numpy(np.concatenate(raws)).cuda(), offsets.cuda(), T, sp_n, ep_n, spos)
    assert int(got["status"].item()) == 0
    for i, (out, nf, si, ei) in enumerate(want):
        assert int(got["nfeats"][i]) == nf, (i, lens[i])
        assert torch.equal(got["video_features"][i].cpu(), out), (i, lens[i])
        assert int(got["start_index"][i]) == si and int(got["end_index"][i]) == ei, (i, lens[i], si, ei)
        assert int(got["video_mask"][i].sum()) == nf and bool(got["video_mask"][i, :nf].all())
