#!/usr/bin/env python
"""A few training steps at a given config / batch (for ncu launch lists of the training path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vml_b200  # noqa
from vml_b200 import synth
from vml_b200.configs import CONFIGS, init_params
from vml_b200.optim import FusedAdam
from vml_b200.smin import SMIN
from vml_b200.trainer import train_step

name = sys.argv[1] if len(sys.argv) > 1 else "tacos"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = CONFIGS[name]
m = SMIN(*cfg.ctor_args(), device=torch.device("cuda"), precision="fp32")
m.load_state_dict(init_params(cfg, 43))
m = m.cuda().train()
opt = FusedAdam(m.parameters(), lr=1e-4)
b = {k: v.cuda() for k, v in synth.make_batch(cfg, B, 5).items()}
for it in range(3):
    loss = train_step(m, opt, b)
torch.cuda.synchronize()
print("ok", name, B, float(loss))
