"""B200-native (sm_100a) implementation of the SMIN proposal-scoring hot path.

Drop-in for the reference's ``models.SMIN`` / ``utils.compute_ious`` / ``main.loss_fn``
(see INTEGRATION.md).  All compute goes through the C-ABI CUDA library
``libvml_b200.so`` (include/vml_b200.h); there is no CPU fallback -- importing the
compute modules without the built library raises.

Submodules are imported lazily so that host-only helpers (``synth``, ``build``) stay
usable on a machine without the built library or a GPU.
"""
import importlib

__all__ = ["SMIN", "compute_ious", "loss_fn", "bce_loss", "synth", "build", "lib"]

_LAZY = {
    "SMIN": ("smin", "SMIN"),
    "compute_ious": ("evaluate", "compute_ious"),
    "loss_fn": ("losses", "loss_fn"),
    "bce_loss": ("losses", "bce_loss"),
}


def __getattr__(name):
    if name in _LAZY:
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    if name in ("synth", "build", "lib", "smin", "evaluate", "losses", "dist", "pack"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
