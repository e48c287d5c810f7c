#!/usr/bin/env python
"""One eager forward + eval at a given batch size (for compute-sanitizer / ncu spot checks on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vml_b200  # noqa
from vml_b200 import synth
from vml_b200.configs import CONFIGS, init_params
from vml_b200.evaluate import compute_ious
from vml_b200.smin import SMIN

name = sys.argv[1] if len(sys.argv) > 1 else "charadessta"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 192
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
cfg = CONFIGS[name]
m = SMIN(*cfg.ctor_args(), device=torch.device("cuda"), precision=prec)
m.load_state_dict(init_params(cfg, 43))
m = m.cuda().eval()
b = {k: v.cuda() for k, v in synth.make_batch(cfg, B, 5).items()}
for it in range(2):
    out = m(*[b[k] for k in synth.MODEL_INPUT_KEYS], overlap=(it == 1))
    r = compute_ious(out[0], out[1], out[2], b["moment_mask"], b["sm"])
torch.cuda.synchronize()
print("ok", name, B, prec, dict(r))
