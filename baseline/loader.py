"""Import the unmodified reference from ``baseline/_ref`` (or, in the build container, straight from /root/reference).

    ref = load_reference()             # ref.models, ref.utils, ref.dataset, ref.main  -- all the reference's own code
    ref = load_reference(dropin=True)  # ref.main is the reference's main.py, but its `from models import SMIN` /
                                       # `from utils import compute_ious` (main.py:3,5) resolved to the drop-in

Stubs: ``torchtext`` (a one-word GloVe stand-in; ``dataset.py:19-24`` builds the vocabulary in a
class body at import) and ``h5py`` are absent in this image.  ``torch.nn.BCELoss(reduction=None)`` (main.py:92-97) raises
in every torch release; it is read as ``reduction='none'`` by a wrapper bound into ``main``'s namespace only -- no file is
edited and the global ``torch.nn.BCELoss`` is left alone.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COPY = os.path.join(ROOT, "baseline", "_ref")
REF_SOURCE = os.environ.get("VML_REFERENCE_DIR", "/root/reference")
DROPIN = os.path.join(ROOT, "video-moment-localization_b200", "dropin")


def reference_dir() -> str | None:
    for d in (REF_COPY, REF_SOURCE):
        if os.path.exists(os.path.join(d, "models.py")):
            return d
    return None


def available() -> bool:
    return reference_dir() is not None


def _stub_missing_deps():
    import torch
    try:
        import torchtext  # noqa: F401
    except Exception:
        class FakeVocab:
            def __init__(self):
                self.itos, self.stoi, self.vectors, self.dim = ["the"], {"the": 0}, torch.zeros(1, 300), 300

        tt = types.ModuleType("torchtext")
        tt.vocab = types.ModuleType("torchtext.vocab")
        tt.vocab.pretrained_aliases = {"glove.6B.300d": FakeVocab}
        sys.modules.update({"torchtext": tt, "torchtext.vocab": tt.vocab})
    try:
        import h5py  # noqa: F401
    except Exception:
        sys.modules["h5py"] = types.ModuleType("h5py")


def _exec(path: str, name: str, bindings: dict):
    """Execute one source file as module ``name`` while the top-level names in ``bindings`` (``models``, ``utils``,
    ``dataset``: what the reference's import statements ask for) resolve to the given modules."""
    saved = {k: sys.modules.get(k) for k in bindings}
    sys.modules.update(bindings)
    old_flag, sys.dont_write_bytecode = sys.dont_write_bytecode, True
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = old_flag
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class _TorchNNProxy(types.ModuleType):
    """``torch.nn`` with ``BCELoss(reduction=None)`` read as ``'none'`` (the documented one-token fix)."""

    def __init__(self, real):
        super().__init__("torch.nn")
        self.__dict__["_real"] = real

    def __getattr__(self, k):
        return getattr(self.__dict__["_real"], k)

    def BCELoss(self, *a, **kw):
        if "reduction" in kw and kw["reduction"] is None:
            kw["reduction"] = "none"
        return self.__dict__["_real"].BCELoss(*a, **kw)


class _TorchProxy(types.ModuleType):
    def __init__(self, real):
        super().__init__("torch")
        self.__dict__["_real"] = real
        self.__dict__["nn"] = _TorchNNProxy(real.nn)

    def __getattr__(self, k):
        return getattr(self.__dict__["_real"], k)


def load_reference(dropin: bool = False, tag: str | None = None):
    d = reference_dir()
    if d is None:
        raise FileNotFoundError("the reference is not installed: run `python -m baseline.install` in the build container "
                                "(copies /root/reference/{models,utils,main,dataset}.py into git-ignored baseline/_ref/)")
    import torch
    _stub_missing_deps()
    tag = tag or ("dropin" if dropin else "ref")
    ns = types.SimpleNamespace(dir=d, dropin=dropin)
    ns.utils = _exec(os.path.join(d, "utils.py"), f"vml_{tag}_ref_utils", {})
    ns.models = _exec(os.path.join(d, "models.py"), f"vml_{tag}_ref_models", {})
    ns.dataset = _exec(os.path.join(d, "dataset.py"), f"vml_{tag}_ref_dataset", {"utils": ns.utils})
    if dropin:
        import vml_b200  # noqa: F401
        m = _exec(os.path.join(DROPIN, "models.py"), f"vml_{tag}_dropin_models", {})
        u = _exec(os.path.join(DROPIN, "utils.py"), f"vml_{tag}_dropin_utils", {})
        ns.main = _exec(os.path.join(d, "main.py"), f"vml_{tag}_ref_main", {"models": m, "utils": u, "dataset": ns.dataset})
        ns.dropin_models, ns.dropin_utils = m, u
    else:
        ns.main = _exec(os.path.join(d, "main.py"), f"vml_{tag}_ref_main", {"models": ns.models, "utils": ns.utils, "dataset": ns.dataset})
    ns.main.torch = _TorchProxy(torch)          # main's own global name `torch`: only main.bce_loss sees the fix
    return ns


def yaml_params(ns, name: str) -> dict:
    """The reference's ``config/<name>.yml`` as ``main.get_parameters`` would load it (main.py:20-26)."""
    import yaml
    with open(os.path.join(ns.dir, "config", f"{name}.yml")) as f:
        params = yaml.load(f, Loader=yaml.SafeLoader)
    params["experiment"], params["test"] = name, False
    return params


class SyntheticAnnotations:
    """Feeds the reference's own ``AbstractDataset.__getitem__`` / ``collate_fn`` (dataset.py:76-90,129-187) with
    seeded synthetic annotations + clip features, so every mask and label a batch carries is computed by reference
    code.  ``make(ns, cfg, n, seed)`` returns a ``ns.dataset.AbstractDataset`` subclass instance."""

    @staticmethod
    def make(ns, cfg, n: int, seed: int, split: str = "test"):
        import numpy as np
        import torch
        from vml_b200 import synth

        base = synth.make_batch(cfg, n, seed)

        class Synthetic(ns.dataset.AbstractDataset):
            def __init__(self):
                self.T, self.L, self.max_query_length, self.split = cfg.T, cfg.L, cfg.Nq, split
                self.annotations = []
                for b in range(n):
                    nf = int(base["nfeats"][b])
                    qlen = int(base["query_mask"][b].sum())
                    pad = self.vocab.stoi["<pad>"]
                    tok = torch.full((cfg.Nq,), pad, dtype=torch.long)
                    tok[:qlen] = 0
                    self.annotations.append({"video_id": b, "times": [float(x) for x in base["times"][b]],
                                             "duration": float(base["duration"][b]), "query_features": base["query_features"][b],
                                             "token_idx": tok, "nfeats": nf})

            def _load_video_features(self, vid):
                nf = self.annotations[vid]["nfeats"]
                return base["video_features"][vid, :nf].numpy().astype(np.float64)

        return Synthetic(), base
