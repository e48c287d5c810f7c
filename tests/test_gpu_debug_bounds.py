"""compute-sanitizer is not available on the GPU pool: the bounds-checking build (``python -m vml_b200.build --debug``,
-DVML_DEBUG_BOUNDS: device-side asserts on cell codes, live counts, ring slots, tensor-memory columns and shared-memory boxes)
runs the forward on awkward shapes in a subprocess -- and must TRAP on a deliberately corrupted cell list, which shows the
checks are live."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "video-moment-localization_b200", "libvml_b200_dbg.so")

SCRIPT = r'''
import sys, torch
sys.path.insert(0, %(root)r)
import vml_b200
from vml_b200 import lib, synth
from vml_b200.configs import CONFIGS, init_params
from vml_b200.smin import SMIN
assert lib.LIB_PATH.endswith("_dbg.so")
mode = sys.argv[1]
for name, B, rng, prec in (("charadessta", 24, (1, 12), "bf16"), ("activitynet", 5, None, "bf16"), ("tacos", 9, None, "bf16"),
                           ("tiny_r2", 5, None, "fp32")):
    cfg = CONFIGS[name]
    m = SMIN(*cfg.ctor_args(), device=torch.device("cuda"), precision=prec)
    m.load_state_dict(init_params(cfg, 43))
    m = m.cuda().eval()
    b = synth.make_batch(cfg, B, 31, **({"nfeats_range": rng} if rng else {}))
    out = m(*[b[k].cuda() for k in synth.MODEL_INPUT_KEYS])
    torch.cuda.synchronize()
    assert all(bool(torch.isfinite(o).all()) for o in out), name
    if mode == "corrupt" and name == "charadessta":
        # a cell code whose j lies outside the map: the span-pool kernel's assert must fire
        from vml_b200.lib import call, ptr, stream_ptr, Dims
        ws = m._ws[str(b["video_features"].cuda().device)]
        code = ws.buf["cell_code"]
        code[3] = (0 << 16) | (2 << 8) | 200
        cells = vml_b200.smin.make_cells(ws, B, cfg.L)
        dims = Dims(cfg.T, cfg.L, cfg.C, cfg.D, cfg.dl, cfg.layers, cfg.d0, cfg.Nq, cfg.H)
        fs_ptr = ws.buf["fwfs"].data_ptr() + B * cfg.Nq * cfg.D * 4          # sentence states follow the word states
        call("vml_span_pool_fuse", ptr(ws.buf["fv"]), fs_ptr, cells, ptr(ws.buf["fc_a"]), ptr(ws.buf["fm_a"]),
             ptr(ws.buf["fb_a"]), B, dims, lib.BF16, stream_ptr())
        torch.cuda.synchronize()
        print("NOT TRAPPED")
print("ok")
'''


def _run(mode):
    env = dict(os.environ, VML_LIB=DBG)
    return subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}, mode], capture_output=True, text=True, env=env, timeout=600)


@pytest.mark.skipif(not os.path.exists(DBG), reason="debug build missing: python -m vml_b200.build --debug")
def test_forward_passes_every_device_side_bounds_assert():
    r = _run("clean")
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-2000:], r.stderr[-2000:])
    assert "VML_DBG_ASSERT" not in r.stdout


@pytest.mark.skipif(not os.path.exists(DBG), reason="debug build missing: python -m vml_b200.build --debug")
def test_bounds_asserts_are_live():
    r = _run("corrupt")
    assert "NOT TRAPPED" not in r.stdout
    assert r.returncode != 0 and "VML_DBG_ASSERT failed" in (r.stdout + r.stderr), (r.stdout[-2000:], r.stderr[-2000:])
