#!/usr/bin/env python
"""Per-phase timeline of content_unit_kernel's CTA 0 (needs a library built with
VML_EXTRA_CFLAGS=-DVML_CU_TIMING; development aid, not part of the product)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vml_b200  # noqa
from vml_b200 import lib, synth
from vml_b200.configs import CONFIGS, init_params
from vml_b200.smin import SMIN

name = sys.argv[1] if len(sys.argv) > 1 else "charadessta"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 192
cfg = CONFIGS[name]
m = SMIN(*cfg.ctor_args(), device=torch.device("cuda"), precision="bf16")
m.load_state_dict(init_params(cfg, 43))
m = m.cuda().eval()
b = {k: v.cuda() for k, v in synth.make_batch(cfg, B, 5).items()}
for it in range(3):
    m(*[b[k] for k in synth.MODEL_INPUT_KEYS], overlap=False)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 192)()
L = lib.load()
L.vml_debug_cu_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert L.vml_debug_cu_timing(buf, 192) == 0
names = ["start", "staged", "chat_full", "chat_parked", "sfull", "softmax", "afull", "gate", "cc_ready",
         "y0_full", "y0_done", "y1_full", "y1_done", "y2_full", "y2_done", "y3_full", "y3_done", "tile_end"]
t0 = buf[0]
for it in range(4):
    row = [buf[it * 48 + s] for s in range(18)]
    print(f"tile iteration {it}: start at +{(row[0] - t0) / 1e3:.2f} us")
    for s in range(1, 18):
        print(f"   {names[s]:12s} +{(row[s] - row[s - 1]) / 1e3:7.2f} us   (t = {(row[s] - row[0]) / 1e3:7.2f})")
    for nb in range(4):
        x = [buf[it * 48 + 18 + 3 * nb + j] for j in range(3)] + [buf[it * 48 + 30 + nb]]
        print(f"   nb {nb}: loop2 done t={(buf[it * 48 + 38 + nb] - row[0]) / 1e3:7.2f}  fq issued t={(buf[it * 48 + 34 + nb] - row[0]) / 1e3:7.2f}")
        print(f"   nb {nb}: W2 load issued t={(x[3] - row[0]) / 1e3:7.2f}  mma: yempty ok t={(x[0] - row[0]) / 1e3:7.2f}  wfull ok t={(x[1] - row[0]) / 1e3:7.2f}  committed t={(x[2] - row[0]) / 1e3:7.2f}")
