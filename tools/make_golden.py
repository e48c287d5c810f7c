#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py

For every config in oracle.CONFIGS it
  1. builds the reference ``models.SMIN`` (imported from /root/reference, nothing
     copied), loads ``oracle.init_params(cfg, 43)`` into it via ``load_state_dict``
     (strict) -- which also pins the state_dict key/shape contract,
  2. runs it in fp32 and in fp64 on ``synth.make_batch(cfg, B, seed)``,
  3. runs the reference ``utils.compute_ious`` and the reference ``main.loss_fn``
     (with ``BCELoss(reduction=None)`` read as ``'none'``: the unpatched call raises
     ``ValueError``, SURVEY.md F4 -- that one-token fix is applied by a wrapper
     around ``torch.nn.BCELoss`` here, the reference file is not edited),
  4. recomputes the labels with the reference ``dataset.AbstractDataset`` methods
     (torchtext / h5py stubbed),
and stores only OUTPUTS (inputs and weights are regenerated from seeds) plus input
checksums.  For the two ``tiny*`` configs every intermediate of the reference is
stored too (forward hooks), for stage-level oracle tests.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VML_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

import vml_b200  # noqa: E402,F401
from vml_b200 import synth  # noqa: E402
from oracle import CONFIGS, init_params  # noqa: E402


def import_reference():
    """The unmodified reference through baseline.loader (torchtext / h5py stubs, ``reduction=None`` read as ``'none'`` inside
    main's own namespace only); straight from /root/reference when baseline/_ref has not been installed."""
    from baseline import loader as bl
    ns = bl.load_reference()
    return ns.models, ns.utils, ns.dataset, ns.main


# golden file -> (config, batch size, batch seed).  The *_b64 / *_b16 files are the shapes bench.py times (charadessta_b64 is
# the bench's own first batch, seed 1000); activitynet_b4 has non-zero R@n counts (the B=2 file's are all zero).
BATCHES = {"charadessta": ("charadessta", 4, 101), "tacos": ("tacos", 3, 102), "activitynet": ("activitynet", 2, 103),
           "tiny": ("tiny", 5, 104), "tiny_r2": ("tiny_r2", 5, 105),
           "activitynet_b4": ("activitynet", 4, 110), "charadessta_b64": ("charadessta", 64, 1000),
           "tacos_b64": ("tacos", 64, 1001), "activitynet_b16": ("activitynet", 16, 1002)}


def main():
    ref_models, ref_utils, ref_dataset, ref_main = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    only = sys.argv[1:]
    for name, (cfg_name, B, seed) in BATCHES.items():
        if only and name not in only:
            continue
        cfg = CONFIGS[cfg_name]
        params = init_params(cfg, 43)
        batch = synth.make_batch(cfg, B, seed)
        model = ref_models.SMIN(*cfg.ctor_args(), torch.device("cpu"))
        sd = model.state_dict()
        assert list(sd.keys()) == list(params.keys()), "state_dict key order/name contract changed"
        for k in sd:
            assert tuple(sd[k].shape) == tuple(params[k].shape), k
        model.load_state_dict(params, strict=True)
        model.eval()
        args = [batch[k] for k in synth.MODEL_INPUT_KEYS]

        store = {}
        inter = {}
        if name.startswith("tiny"):
            def hook(tag):
                def fn(_m, _i, o):
                    inter.setdefault(tag, []).append(o)
                return fn
            model.backbone.videoencoder.register_forward_hook(hook("fv"))
            model.backbone.queryencoder.register_forward_hook(hook("q"))
            model.pgm.register_forward_hook(hook("pgm"))
            for k, smi in enumerate(model.smis):
                smi.register_forward_hook(hook(f"smi{k}"))
        with torch.no_grad():
            pm, ps, pe, pa = model(*args)
        if inter:
            store["fv"] = inter["fv"][0].numpy()
            store["fs"], store["fw"] = (t.numpy() for t in inter["q"][0])
            store["fc0"], store["fm0"], store["fb0"] = (t.numpy() for t in inter["pgm"][0])
            for k in range(cfg.layers):
                cu, mu, bu = inter[f"smi{k}"][0]
                store[f"fc{k + 1}"], store[f"fm{k + 1}"], store[f"fb{k + 1}"] = cu.numpy(), mu.numpy(), bu.numpy()
            for h in list(model._forward_hooks.values()):
                pass
        store.update(pm=pm.numpy(), ps=ps.numpy(), pe=pe.numpy(), pa=pa.numpy())

        # fp64 adjudication run (Wc is a plain attribute -> convert by hand)
        import copy
        m64 = copy.deepcopy(model).double()
        m64._forward_hooks.clear()
        for mod in m64.modules():
            mod._forward_hooks.clear()
        m64.pgm.Wc = m64.pgm.Wc.double()
        with torch.no_grad():
            o64 = m64(batch["video_features"].double(), batch["video_mask"], batch["query_features"].double(),
                      batch["query_mask"], batch["length_mask"], batch["moment_mask"])
        store.update(pm64=o64[0].numpy(), ps64=o64[1].numpy(), pe64=o64[2].numpy(), pa64=o64[3].numpy())

        # metric: reference compute_ious + the reference's own top-k order
        metrics = ref_utils.compute_ious(pm, ps, pe, batch["moment_mask"], batch["sm"])
        score = (pm * torch.sqrt(ps.unsqueeze(2)) * torch.sqrt(pe.unsqueeze(1)) * batch["moment_mask"]).view(B, -1)
        store["ref_topk"] = score.topk(k=5, dim=1)[1].numpy()
        store["metric_keys"] = np.array(sorted(metrics.keys()))
        store["metric_vals"] = np.array([metrics[k] for k in sorted(metrics.keys())], dtype=np.float64)

        # loss: reference loss_fn with the one-token fix
        loss = ref_main.loss_fn(pm, batch["ym"], batch["sm"], batch["moment_mask"], ps, batch["ys"], batch["ss"],
                                pe, batch["ye"], batch["se"], pa, batch["ya"], batch["length_mask"])
        store["loss"] = np.array(loss.item(), dtype=np.float64)
        parts = [ref_main.bce_loss(pm, batch["ym"], batch["sm"], batch["moment_mask"]),
                 ref_main.bce_loss(ps, batch["ys"], batch["ss"], batch["length_mask"]),
                 ref_main.bce_loss(pe, batch["ye"], batch["se"], batch["length_mask"]),
                 ref_main.bce_loss(pa, batch["ya"], None, batch["length_mask"])]
        store["loss_parts"] = np.array([p.item() for p in parts], dtype=np.float64)

        # labels through the reference dataset methods
        ds = ref_dataset.AbstractDataset.__new__(ref_dataset.AbstractDataset)
        ds.T, ds.L = cfg.T, cfg.L
        sm_ref, ss_ref, se_ref, ya_ref = [], [], [], []
        for b in range(min(B, 8)):
            ts, te = (float(x) for x in batch["times"][b])
            du = float(batch["duration"][b])
            sm_ref.append(ds.get_iou(ts, te, du))
            s_s, s_e = ds.get_boundary_penalties(ts, te, du)
            ss_ref.append(s_s)
            se_ref.append(s_e)
            ya_ref.append(ds.get_snippet_label(ts, te, du))
        store["sm_ref"] = torch.stack(sm_ref).numpy()
        store["ss_ref"] = torch.stack(ss_ref).numpy()
        store["se_ref"] = torch.stack(se_ref).numpy()
        store["ya_ref"] = torch.stack(ya_ref).numpy()
        store["Wc_nnz"] = np.array(int((model.pgm.Wc != 0).sum()))
        if name.startswith("tiny"):
            store["Wc"] = model.pgm.Wc.numpy()

        # input checksums (detect generator drift)
        store["chk_video"] = np.array(batch["video_features"].double().sum().item())
        store["chk_query"] = np.array(batch["query_features"].double().sum().item())
        store["chk_params"] = np.array(sum(v.double().sum().item() for v in params.values()))
        store["n_params"] = np.array(sum(v.numel() for v in params.values()))
        store["B"], store["seed"] = np.array(B), np.array(seed)
        path = os.path.join(out_dir, f"{name}.npz")
        np.savez_compressed(path, **store)
        print(f"{name}: B={B} params={int(store['n_params'])} loss={float(store['loss']):.6f} "
              f"metrics={dict(zip(store['metric_keys'].tolist(), store['metric_vals'].tolist()))} "
              f"-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
