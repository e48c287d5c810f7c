"""Fused Adam over one flat fp32 buffer (drop-in for ``torch.optim.Adam(model.parameters(), lr)``, main.py:83).

Parameters are re-homed as views into a single contiguous buffer at construction, so that a step is ONE
``vml_adam_step`` launch (plus one gradient flatten) and a data-parallel gradient all-reduce is ONE NCCL call
on ``flat_grad``.  Defaults are torch.optim.Adam's (betas 0.9/0.999, eps 1e-8, no weight decay, no amsgrad)."""
from __future__ import annotations

import torch

from . import lib as L_
from .lib import call, ptr, stream_ptr


class FusedAdam:
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params = [p for p in params]
        if not self.params or not all(p.is_cuda and p.dtype == torch.float32 for p in self.params):
            raise L_.VmlError("FusedAdam needs fp32 CUDA parameters; there is no CPU path")
        L_.load()
        self.lr, self.betas, self.eps, self.t = lr, betas, eps, 0
        self.generation = 0                 # number of steps applied to the flat buffer (diagnostics / tests)
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        o = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[o:o + k].copy_(p.reshape(-1))
                p.data = self.flat[o:o + k].view(p.shape)            # the parameter now lives inside the flat buffer
                o += k
        self._offsets = self._param_offsets()
        self._grad_views = [self.flat_grad[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self._offsets)]

    def _param_offsets(self):
        out, o = [], 0
        for p in self.params:
            out.append(o)
            o += p.numel()
        return out

    def _check_homed(self):
        """A later ``model.to()`` / ``.float()`` / ``nn.LSTM.flatten_parameters`` rebinds ``p.data`` and would leave the
        optimizer updating a dead buffer: refuse to step instead."""
        base = self.flat.data_ptr()
        for p, o in zip(self.params, self._offsets):
            if p.data_ptr() != base + 4 * o:
                raise L_.VmlError("FusedAdam: a parameter no longer lives in the optimizer's flat buffer (it was re-bound by "
                                  ".to()/.float()/flatten_parameters after the optimizer was built); rebuild the optimizer")

    # -- checkpoint compatibility (main.py:270-274 saves {"epoch", "model", "optimizer": optimizer.state_dict()}) -----
    @property
    def param_groups(self):
        return [{"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.params)))}]

    def state_dict(self):
        """Same layout as ``torch.optim.Adam.state_dict()``: loads into a stock Adam over the same parameter list
        (and vice versa through ``load_state_dict``), so checkpoints move freely between the two."""
        state, o = {}, 0
        if self.t > 0:
            for i, p in enumerate(self.params):
                k = p.numel()
                state[i] = {"step": torch.tensor(float(self.t)), "exp_avg": self.m[o:o + k].view(p.shape).clone(),
                            "exp_avg_sq": self.v[o:o + k].view(p.shape).clone()}
                o += k
        return {"state": state, "param_groups": self.param_groups}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if sum(len(g["params"]) for g in groups) != len(self.params):
            raise ValueError("loaded state dict has a different number of parameters")
        g0 = groups[0]
        self.lr, self.betas, self.eps = float(g0["lr"]), tuple(g0["betas"]), float(g0["eps"])
        if any(g.get("weight_decay", 0) or g.get("amsgrad", False) for g in groups):
            raise L_.VmlError("FusedAdam implements plain Adam only (no weight decay / amsgrad)")
        order = [i for g in groups for i in g["params"]]
        self.m.zero_(); self.v.zero_()
        steps, o = set(), 0
        with torch.no_grad():
            for idx, p in zip(order, self.params):
                k = p.numel()
                st = sd["state"].get(idx)
                if st is not None:
                    self.m[o:o + k].copy_(st["exp_avg"].reshape(-1))
                    self.v[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.add(int(float(st["step"])))
                o += k
        if len(steps) > 1:
            raise L_.VmlError("FusedAdam keeps one step counter; the loaded state has several")
        self.t = steps.pop() if steps else 0

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            p.grad = None

    def gather_grads(self):
        """Copy the per-parameter .grad tensors into ``flat_grad`` (zeros where a parameter has no grad)."""
        have = [(v, p.grad) for v, p in zip(self._grad_views, self.params) if p.grad is not None]
        if len(have) < len(self.params):
            self.flat_grad.zero_()
        if have:            # one multi-tensor copy instead of one launch per parameter (107 on the 3-layer model)
            torch._foreach_copy_([v for v, _ in have], [g.reshape(v.shape) for v, g in have])
        return self.flat_grad

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, gathered: bool = False):
        self._check_homed()
        if not gathered:
            self.gather_grads()
        self.t += 1
        call("vml_adam_step", ptr(self.flat), ptr(self.flat_grad), ptr(self.m), ptr(self.v), self.flat.numel(), self.lr,
             self.betas[0], self.betas[1], self.eps, self.t, grad_scale, stream_ptr())
        # The kernel wrote the parameters behind autograd's back.  Each Parameter re-homed with ``p.data = view`` keeps its
        # OWN version counter (it does not share ``flat``'s), so bump every one explicitly: SMIN's packed-weight cache is
        # keyed on (data_ptr, _version) and must see the step.
        torch.autograd.graph.increment_version(self.params)
        self.generation += 1
