#!/usr/bin/env python
"""Minimal driver for profilers: N serial eager forwards of the drop-in module at one pass shape.
    ncu --set full --import-source on -k regex:content_unit_kernel -s 3 -c 1 -o gpurun_out/cu python tools/run_forward.py charadessta 256 3
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vml_b200  # noqa
from vml_b200 import synth
from vml_b200.configs import CONFIGS, init_params
from vml_b200.smin import SMIN

name = sys.argv[1] if len(sys.argv) > 1 else "charadessta"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
cfg = CONFIGS[name]
m = SMIN(*cfg.ctor_args(), device=torch.device("cuda"), precision=prec)
m.load_state_dict(init_params(cfg, 43))
m = m.cuda().eval()
parts = [synth.make_batch(cfg, 64, 1000 + i) for i in range((B + 63) // 64)]
b = {k: torch.cat([p[k] for p in parts])[:B].cuda() for k in synth.MODEL_INPUT_KEYS}
for it in range(iters):
    out = m(*[b[k] for k in synth.MODEL_INPUT_KEYS], overlap=False)
torch.cuda.synchronize()
print("ok", float(out[0].sum()))
