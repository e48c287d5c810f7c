"""The reference's OWN loops (main.py:135-211: ``train_epoch``, ``eval_epoch``, ``test_model``) over the reference's own
``AbstractDataset.__getitem__`` / ``collate_fn`` batches (dataset.py:76-90,129-187; synthetic annotations), executed

  * with the reference's ``models.SMIN`` / ``utils.compute_ious`` on the CPU  (the baseline: unmodified code), and
  * with ``dropin/`` bound to the names ``models`` / ``utils`` that ``main.py:3,5`` imports  (the product, on the GPU),

must return the same losses and metric dicts.  The reference files come from ``baseline/_ref`` (see baseline/install.py).
"""
import os

import pytest
import torch
from torch.utils.data import DataLoader

from baseline import loader as bl
from oracle import CONFIGS, init_params, smin_forward as oracle_forward
from oracle import metrics_oracle as mo
from vml_b200 import synth

needs_ref = pytest.mark.skipif(not bl.available(), reason="baseline/_ref not installed (python -m baseline.install)")


def _loader(ns, cfg, n, seed, bs, split="test"):
    ds, base = bl.SyntheticAnnotations.make(ns, cfg, n, seed, split)
    return DataLoader(ds, batch_size=bs, shuffle=False, collate_fn=ds.collate_fn, num_workers=0), base


def _params_for(cfg, dev):
    return {"T": cfg.T, "L": cfg.L, "C": cfg.C, "d": cfg.D, "dl": cfg.dl, "num_smi_layers": cfg.layers, "input_video_dim": cfg.d0,
            "max_query_length": cfg.Nq, "lstm_hidden_size": cfg.H, "device": dev, "model": "SMIN", "optimizer": "Adam", "lr": 1e-3}


@needs_ref
def test_reference_dataset_pipeline_reproduces_synth_batches():
    """synth.make_batch restates dataset.py's mask / label formulas: the reference's own __getitem__ + collate_fn, fed the same
    annotations, must give the same 13 tensors bit for bit (pins the inputs of every loss / metric parity test)."""
    ns = bl.load_reference()
    for name, n in (("charadessta", 6), ("tiny_r2", 5), ("activitynet", 3)):
        cfg = CONFIGS[name]
        dl, base = _loader(ns, cfg, n, 7, n)
        b = next(iter(dl))
        for k in synth.MODEL_INPUT_KEYS + synth.LOSS_LABEL_KEYS:
            assert b[k].dtype == base[k].dtype and torch.equal(b[k], base[k]), (name, k)


@needs_ref
def test_reference_eval_epoch_on_cpu_equals_oracle():
    """main.eval_epoch / test_model with the reference model on CPU == the oracle's forward + loss + metric: pins the oracle
    through the reference's own loop (loss with the documented reduction fix)."""
    ns = bl.load_reference()
    cfg = CONFIGS["tiny"]
    params = init_params(cfg, 43)
    P = _params_for(cfg, torch.device("cpu"))
    model = ns.main.get_model(P)
    model.load_state_dict(params, strict=True)
    dl, base = _loader(ns, cfg, 10, 11, 4)
    with torch.no_grad():
        loss, metrics = ns.main.eval_epoch(model, dl, torch.device("cpu"), P)
        metrics_t = ns.main.test_model(model, dl, torch.device("cpu"), P)
    want_loss, want = 0.0, {}
    for lo in range(0, 10, 4):
        b = {k: v[lo:lo + 4] for k, v in base.items()}
        with torch.no_grad():
            pm, ps, pe, pa = oracle_forward(params, cfg, *[b[k] for k in synth.MODEL_INPUT_KEYS])
        want_loss += mo.loss_fn(pm, b["ym"], b["sm"], b["moment_mask"], ps, b["ys"], b["ss"], pe, b["ye"], b["se"], pa, b["ya"],
                                b["length_mask"]).item() * pm.shape[0]
        m = mo.compute_ious(pm, ps, pe, b["moment_mask"], b["sm"])
        want = {k: want.get(k, 0.0) + v for k, v in m.items()}
    assert abs(loss - want_loss / 10) < 1e-6 * abs(loss)
    assert metrics == {k: v / 10 for k, v in want.items()} and metrics_t == metrics


@needs_ref
def test_main_py_resolves_its_imports_to_the_dropin():
    ns = bl.load_reference(dropin=True)
    import vml_b200.evaluate as e
    import vml_b200.smin as s
    assert ns.main.SMIN is s.SMIN and ns.main.compute_ious is e.compute_ious
    assert ns.models.SMIN is not s.SMIN                       # the reference's own class is still there, separately


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("name,n,bs", [("tiny", 10, 4), ("charadessta", 12, 5)])
def test_reference_loops_run_unchanged_on_the_dropin(name, n, bs, monkeypatch):
    """eval_epoch, test_model and one train_epoch of the reference's main.py, bound to the drop-in, against the same loops
    bound to the reference's own modules on the CPU.  fp32 mode: metric dicts equal, losses within 1e-5."""
    monkeypatch.setenv("VML_PRECISION", "fp32")
    ref, ours = bl.load_reference(), bl.load_reference(dropin=True)
    cfg = CONFIGS[name]
    params = init_params(cfg, 43)
    cpu, dev = torch.device("cpu"), torch.device("cuda")
    Pr, Po = _params_for(cfg, cpu), _params_for(cfg, dev)
    m_ref = ref.main.get_model(Pr)
    m_ref.load_state_dict(params, strict=True)
    m_our = ours.main.get_model(Po).to(dev)                   # main.py:288-291
    m_our.load_state_dict(params, strict=True)
    assert type(m_our).__module__.startswith("vml_b200")
    dl_r, _ = _loader(ref, cfg, n, 13, bs)
    dl_o, _ = _loader(ours, cfg, n, 13, bs)
    with torch.no_grad():
        loss_r, met_r = ref.main.eval_epoch(m_ref, dl_r, cpu, Pr)
        test_r = ref.main.test_model(m_ref, dl_r, cpu, Pr)
    loss_o, met_o = ours.main.eval_epoch(m_our, dl_o, dev, Po)          # as the reference calls it: no no_grad
    test_o = ours.main.test_model(m_our, dl_o, dev, Po)
    assert met_o == met_r and test_o == test_r
    assert abs(loss_o - loss_r) < 1e-5 * abs(loss_r)
    # one training epoch: reference Adam on both sides (main.get_optimizer, main.py:77-87)
    opt_r, opt_o = ref.main.get_optimizer(m_ref, Pr), ours.main.get_optimizer(m_our, Po)
    tl_r, tm_r = ref.main.train_epoch(m_ref, opt_r, dl_r, cpu, Pr)
    tl_o, tm_o = ours.main.train_epoch(m_our, opt_o, dl_o, dev, Po)
    assert abs(tl_o - tl_r) < 2e-5 * abs(tl_r)
    assert set(tm_o) == set(tm_r)
    bad = [(k, tm_o[k], tm_r[k]) for k in tm_r if abs(tm_o[k] - tm_r[k]) > 1.01 / n]      # <= one near-tie flip after updates
    assert not bad, bad
    with torch.no_grad():
        loss_r2, met_r2 = ref.main.eval_epoch(m_ref, dl_r, cpu, Pr)
    loss_o2, met_o2 = ours.main.eval_epoch(m_our, dl_o, dev, Po)
    assert abs(loss_r2 - loss_r) > 1e-4 * abs(loss_r), "the reference's epoch did not move the loss"
    assert abs(loss_o2 - loss_r2) < 1e-3 * abs(loss_r2)                 # three Adam steps later the two models still agree
    sd_o, sd_r = m_our.state_dict(), m_ref.state_dict()
    assert list(sd_o) == list(sd_r)
    # Adam's first steps move every element by ~lr whatever the gradient's size, so an element whose gradient is rounding
    # noise may legitimately differ by a few lr; the bulk must agree far better than that
    steps = -(-n // bs)
    diff = torch.cat([(sd_o[k].cpu() - sd_r[k]).abs().reshape(-1) for k in sd_r])
    assert diff.mean().item() < 0.05 * 1e-3 * steps and diff.max().item() <= 2.5 * 1e-3 * steps, (diff.mean().item(), diff.max().item())
