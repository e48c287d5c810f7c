"""Functional CPU restatement of the reference SMIN forward (test infrastructure).

Every function takes a flat ``params`` dict keyed exactly like the reference
``SMIN.state_dict()`` and plain tensors, and executes the
same dense math as the reference (all L*L cells, dense content matrix), so it is
both the parity checker and a cost-faithful CPU baseline.  Works in fp32 or
fp64 depending on the dtype of ``params`` / inputs.

Reference citations are to /root/reference (read-only in the build container).
"""
from __future__ import annotations

import math

import torch

import vml_b200  # noqa: F401  (registers the hyphenated package directory)


from vml_b200.configs import CONFIGS, SminConfig, init_params  # noqa: E402,F401  (host-side config, no compute)


# --------------------------------------------------------------------------
# a1  clip projection            models.py:25-36
# --------------------------------------------------------------------------
def clip_projection(p, video_features, video_mask):
    """fv = ve(v)*mask + pe[arange(T)]*mask   (models.py:27-34)."""
    W = p["backbone.videoencoder.ve.weight"]
    m = video_mask.to(W.dtype)                                # [B,T,1]
    x = torch.addmm(p["backbone.videoencoder.ve.bias"], video_features.reshape(-1, W.shape[1]), W.t())
    x = x.view(video_features.shape[0], video_features.shape[1], -1) * m
    pos = p["backbone.videoencoder.pe.weight"][: video_mask.shape[1]].unsqueeze(0) * m
    return x + pos


# --------------------------------------------------------------------------
# a2  query encoder (2-layer bi-LSTM over packed sequences)  models.py:48-64
# --------------------------------------------------------------------------
def _lstm_direction(x, lengths, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one layer with packed-sequence semantics: sample b only
    steps over t < lengths[b]; outputs at t >= lengths[b] are zero
    (pad_packed_sequence); the reverse direction starts at t = lengths[b]-1.
    Gate order i,f,g,o (PyTorch)."""
    B, N, _ = x.shape
    Hh = w_hh.shape[1]
    h = x.new_zeros(B, Hh)
    c = x.new_zeros(B, Hh)
    out = x.new_zeros(B, N, Hh)
    gin = x @ w_ih.t() + (b_ih + b_hh)                       # [B,N,4H]
    steps = range(N - 1, -1, -1) if reverse else range(N)
    for t in steps:
        act = (lengths > t).to(x.dtype).unsqueeze(1)         # [B,1]
        g = gin[:, t] + h @ w_hh.t()
        i, f, gg, o = g.split(Hh, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        c = act * c_new + (1 - act) * c
        h = act * h_new + (1 - act) * h
        out[:, t] = act * h_new
    return out


def query_encoder(p, query_features, query_mask):
    """fs[B,2H], fw[B,Nq,2H]  (models.py:48-64)."""
    lengths = query_mask.reshape(query_mask.shape[0], -1).sum(1).long()
    pre = "backbone.queryencoder.lstm."
    x = query_features
    for layer in range(2):
        outs = []
        for sfx, rev in (("", False), ("_reverse", True)):
            outs.append(_lstm_direction(
                x, lengths,
                p[f"{pre}weight_ih_l{layer}{sfx}"], p[f"{pre}weight_hh_l{layer}{sfx}"],
                p[f"{pre}bias_ih_l{layer}{sfx}"], p[f"{pre}bias_hh_l{layer}{sfx}"], rev))
        x = torch.cat(outs, dim=2)
    fw = x
    Hh = fw.shape[2] // 2
    last = (lengths - 1).view(-1, 1, 1).expand(-1, 1, Hh)
    fs = torch.cat([fw[:, :, :Hh].gather(1, last).squeeze(1), fw[:, 0, Hh:]], dim=1)
    return fs, fw


# --------------------------------------------------------------------------
# a3+a4  fusion + span pooling     models.py:81, 88-98, 115-126
# --------------------------------------------------------------------------
def content_matrix(T, L, C, dtype=torch.float32):
    """Dense Wc[L,L,C,T] of clip-averaging weights (models.py:88-98), built
    without the reference's triple Python loop."""
    r = T // L
    i = torch.arange(L).view(L, 1, 1, 1)
    j = torch.arange(L).view(1, L, 1, 1)
    c = torch.arange(C).view(1, 1, C, 1)
    t = torch.arange(T).view(1, 1, 1, T)
    nf = (j - i + 1) * r
    cs = torch.clamp(torch.div(nf, C, rounding_mode="floor"), min=1)
    start = i * r + c * cs
    on = (j >= i) & (c < torch.minimum(torch.tensor(C), nf)) & (t >= start) & (t < start + cs)
    # the reference builds Wc in float32 (1/clip_size rounded to fp32) and only then casts
    return (on.to(torch.float32) / cs.to(torch.float32)).to(dtype)


def span_pool(f, moment_mask, T, L, C):
    """fc[B,L,L,C,D], fm[B,L,L,D], fb[B,L,D] from fused clips f[B,T,D]
    (models.py:115-126).  fm always divides by C; fb is an unmasked average pool."""
    Wc = content_matrix(T, L, C, f.dtype)
    B, _, D = f.shape
    fc = (Wc.view(L * L * C, T) @ f).view(B, L, L, C, D) * moment_mask.view(B, L, L, 1, 1).to(f.dtype)
    fm = fc.mean(dim=3)
    r = T // L
    fb = f[:, : L * r].reshape(B, L, r, D).mean(dim=2)
    return fc, fm, fb


# --------------------------------------------------------------------------
# a5/a6  content unit              models.py:207-226, 242-276
# --------------------------------------------------------------------------
def _masked_softmax(scores, key_mask):
    """scores*mask -> masked_fill(mask==0,-1e9) -> softmax (models.py:146-150,216-220)."""
    scores = scores * key_mask
    scores = scores.masked_fill(key_mask == 0, -1e9)
    return torch.softmax(scores, dim=-1)


def content_unit(p, k, fc, fw, fs, fm, query_mask, moment_mask):
    pre = f"smis.{k}.content_unit."
    B, L, _, C, D = fc.shape
    dl = p[pre + "linear_c_hat.weight"].shape[0]
    m5 = moment_mask.view(B, L, L, 1, 1).to(fc.dtype)
    qm = query_mask.reshape(B, -1).to(fc.dtype)                                   # [B,Nq]

    c_hat = (fc @ p[pre + "linear_c_hat.weight"].t() + p[pre + "linear_c_hat.bias"]) * m5
    w_hat = (fw @ p[pre + "linear_w_hat.weight"].t() + p[pre + "linear_w_hat.bias"]) * qm.unsqueeze(-1)
    s_hat = fs @ p[pre + "linear_s_hat.weight"].t() + p[pre + "linear_s_hat.bias"]

    # content-word attention (models.py:207-226); value = un-projected w_hat
    q = c_hat @ p[pre + "attn_layer.W_q.weight"].t() + p[pre + "attn_layer.W_q.bias"]
    kk = w_hat @ p[pre + "attn_layer.W_k.weight"].t() + p[pre + "attn_layer.W_k.bias"]
    att = torch.einsum("blmcd,bnd->blmcn", q, kk) / math.sqrt(dl)
    att = _masked_softmax(att, qm.view(B, 1, 1, 1, -1))
    caq = torch.einsum("blmcn,bnd->blmcd", att, w_hat) * m5

    cq = c_hat * (caq + s_hat.view(B, 1, 1, 1, dl))
    a_c = torch.softmax(cq @ cq.transpose(3, 4) / math.sqrt(dl), dim=-1) * m5        # [B,L,L,C,C]
    cc_hat = a_c @ c_hat
    cc = (cc_hat @ p[pre + "linear_c.weight"].t() + p[pre + "linear_c.bias"]) * m5

    gate = torch.sigmoid(fm * fs.view(B, 1, 1, D)) * fm
    return cc + fc + gate.unsqueeze(3)


# --------------------------------------------------------------------------
# a7  boundary unit                models.py:137-154, 164-196
# --------------------------------------------------------------------------
def boundary_unit(p, k, fb, fw, fs, fm, query_mask, length_mask):
    pre = f"smis.{k}.boundary_unit.attn_layer."
    B, L, D = fb.shape
    lm = length_mask.to(fb.dtype)                                                  # [B,L]
    qm = query_mask.reshape(B, -1).to(fb.dtype)

    q = fb @ p[pre + "W_q.weight"].t() + p[pre + "W_q.bias"]
    kk = fw @ p[pre + "W_k.weight"].t() + p[pre + "W_k.bias"]
    att = _masked_softmax(q @ kk.transpose(1, 2) / math.sqrt(D), qm.unsqueeze(1))
    baq = (att @ fw) * lm.unsqueeze(-1)

    bq = fb * (baq + fs.unsqueeze(1))
    a_b = _masked_softmax(bq @ bq.transpose(1, 2) / math.sqrt(D), lm.unsqueeze(1)) * lm.unsqueeze(-1)
    bb = (a_b @ fb) * lm.unsqueeze(-1)

    gated = torch.sigmoid(fm * fs.view(B, 1, 1, D)) * fm                           # [B,L,L,D]
    bm = (a_b.unsqueeze(3) * gated).sum(dim=2)
    return bb + fb + bm


# --------------------------------------------------------------------------
# a8  moment unit (two 1x1 convs)  models.py:288-303
# --------------------------------------------------------------------------
def moment_unit(p, k, cu, fm, bu, moment_mask):
    pre = f"smis.{k}.moment_unit."
    B, L, _, C, D = cu.shape
    m4 = moment_mask.view(B, L, L, 1).to(cu.dtype)
    w_fb = p[pre + "conv_layer_fb.weight"].view(D, D)
    w_fc = p[pre + "conv_layer_fc.weight"].view(D, D)
    pair = bu.unsqueeze(2) * bu.unsqueeze(1)                                       # [B,L,L,D]
    conv_fb = (pair @ w_fb.t() + p[pre + "conv_layer_fb.bias"]) * m4
    conv_fc = (cu.mean(dim=3) @ w_fc.t() + p[pre + "conv_layer_fc.bias"]) * m4
    return conv_fb + conv_fc + fm


# --------------------------------------------------------------------------
# a9  localization                 models.py:335-344
# --------------------------------------------------------------------------
def localization(p, fm, fb, length_mask, moment_mask):
    D = fm.shape[-1]
    pre = "localization.conv_layer_"
    pm = torch.sigmoid(fm @ p[pre + "pm.weight"].view(D) + p[pre + "pm.bias"]) * moment_mask

    def head(nm):
        return torch.sigmoid(fb @ p[pre + nm + ".weight"].view(D) + p[pre + nm + ".bias"]) * length_mask

    return pm, head("ps"), head("pe"), head("pa")


# --------------------------------------------------------------------------
# whole forward                    models.py:367-377
# --------------------------------------------------------------------------
def smin_forward(p, cfg: SminConfig, video_features, video_mask, query_features, query_mask,
                 length_mask, moment_mask, return_intermediates=False):
    fv = clip_projection(p, video_features, video_mask)
    fs, fw = query_encoder(p, query_features, query_mask)
    f = fv * fs.unsqueeze(1)                                                       # models.py:81
    fc, fm, fb = span_pool(f, moment_mask, cfg.T, cfg.L, cfg.C)
    inter = {"fv": fv, "fs": fs, "fw": fw, "fc0": fc, "fm0": fm, "fb0": fb}
    for k in range(cfg.layers):
        cu = content_unit(p, k, fc, fw, fs, fm, query_mask, moment_mask)
        bu = boundary_unit(p, k, fb, fw, fs, fm, query_mask, length_mask)
        mu = moment_unit(p, k, cu, fm, bu, moment_mask)
        fc, fm, fb = cu, mu, bu
        inter[f"fc{k + 1}"], inter[f"fm{k + 1}"], inter[f"fb{k + 1}"] = fc, fm, fb
    out = localization(p, fm, fb, length_mask, moment_mask)
    return (out, inter) if return_intermediates else out
