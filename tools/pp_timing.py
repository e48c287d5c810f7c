#!/usr/bin/env python
"""Timeline of content_unit_pp_kernel's CTA 0 (first 8 tiles): both row groups, the MMA issuer, the store warp and the TMA
producer on one clock (%globaltimer).  Needs `python -m vml_b200.build --timing` and VML_LIB=.../libvml_b200_timing.so
VML_CU_VARIANT=5; development aid, not part of the product."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vml_b200  # noqa
from vml_b200 import lib, synth
from vml_b200.configs import CONFIGS, init_params
from vml_b200.smin import SMIN

name = sys.argv[1] if len(sys.argv) > 1 else "charadessta"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 640
cfg = CONFIGS[name]
m = SMIN(*cfg.ctor_args(), device=torch.device("cuda"), precision="bf16")
m.load_state_dict(init_params(cfg, 43))
m = m.cuda().eval()
parts = [synth.make_batch(cfg, 64, 1000 + i) for i in range((B + 63) // 64)]
b = {k: torch.cat([p[k] for p in parts])[:B].cuda() for k in synth.MODEL_INPUT_KEYS}
for it in range(3):
    m(*[b[k] for k in synth.MODEL_INPUT_KEYS], overlap=False)
torch.cuda.synchronize()
N = 8 * 64
buf = (ctypes.c_longlong * N)()
L = lib.load()
L.vml_debug_pp_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert L.vml_debug_pp_timing(buf, N) == 0
t0 = min(x for x in buf if x > 0)
us = lambda t, s: (buf[t * 64 + s] - t0) / 1e3 if buf[t * 64 + s] > 0 else float("nan")
for t in range(8):
    print(f"tile {t} (group {t & 1}):")
    print(f"  rows   start {us(t,0):7.2f} turn {us(t,1):7.2f} chat_full {us(t,2):7.2f} cc_ready {us(t,3):7.2f} tail_turn {us(t,4):7.2f} | "
          + " ".join(f"y{nb} {us(t,5+2*nb):7.2f}/{us(t,6+2*nb):7.2f}" for nb in range(4)))
    print(f"  mma    main {us(t,16):7.2f}..{us(t,17):7.2f} tail cc_ready {us(t,18):7.2f} | "
          + " ".join(f"b{nb} yempty {us(t,19+3*nb):7.2f} boxes {us(t,20+3*nb):7.2f} commit {us(t,21+3*nb):7.2f}" for nb in range(4)))
    print(f"  store  " + " ".join(f"b{nb} sready {us(t,32+2*nb):7.2f} released {us(t,33+2*nb):7.2f}" for nb in range(4)))
    print(f"  tma    main {us(t,40):7.2f}..{us(t,41):7.2f} tail {us(t,42):7.2f}..{us(t,43):7.2f}")
