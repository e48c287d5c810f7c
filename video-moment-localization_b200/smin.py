"""Drop-in ``SMIN`` module backed by libvml_b200.so.

Same constructor, ``forward`` signature, outputs and ``state_dict`` keys/shapes as the
reference ``models.SMIN`` (models.py:346-377), so the reference's
``main.py`` / ``dataset.py`` / ``config/*.yml`` run unchanged with

    from vml_b200.dropin.models import SMIN        # instead of `from models import SMIN`

All compute is hand-written CUDA reached through the C ABI; PyTorch only owns device
memory, streams and the parameter containers.  ``precision='bf16'`` (default) runs the
tcgen05 path, ``precision='fp32'`` the CUDA-core validation path (1e-5 parity).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import lib as L_
from .lib import Cells, Dims, call, ptr, stream_ptr


# ---------------------------------------------------------------------------------------
# parameter containers: same module tree / names / construction order as the reference,
# so state_dict keys match and torch.manual_seed(s) yields the reference's initial weights.
# None of these modules' forward() is ever called.
# ---------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container only; compute goes through libvml_b200")


class _VideoEncoder(_Holder):
    def __init__(self, T, d, d0):
        super().__init__()
        self.ve = nn.Linear(d0, d)
        self.pe = nn.Embedding(T, d)


class _QueryEncoder(_Holder):
    def __init__(self, H):
        super().__init__()
        self.lstm = nn.LSTM(input_size=300, hidden_size=H, num_layers=2, bidirectional=True, batch_first=True)


class _Backbone(_Holder):
    def __init__(self, T, d, d0, H):
        super().__init__()
        self.videoencoder = _VideoEncoder(T, d, d0)
        self.queryencoder = _QueryEncoder(H)


class _Attn(_Holder):
    def __init__(self, D):
        super().__init__()
        self.W_q = nn.Linear(D, D)
        self.W_k = nn.Linear(D, D)


class _ContentUnit(_Holder):
    def __init__(self, D, dl):
        super().__init__()
        self.linear_c_hat = nn.Linear(D, dl)
        self.linear_w_hat = nn.Linear(D, dl)
        self.linear_s_hat = nn.Linear(D, dl)
        self.linear_c = nn.Linear(dl, D)
        self.attn_layer = _Attn(dl)


class _BoundaryUnit(_Holder):
    def __init__(self, D):
        super().__init__()
        self.attn_layer = _Attn(D)


class _MomentUnit(_Holder):
    def __init__(self, D):
        super().__init__()
        self.conv_layer_fb = nn.Conv2d(D, D, 1)
        self.conv_layer_fc = nn.Conv2d(D, D, 1)


class _SMI(_Holder):
    def __init__(self, D, dl):
        super().__init__()
        self.content_unit = _ContentUnit(D, dl)
        self.boundary_unit = _BoundaryUnit(D)
        self.moment_unit = _MomentUnit(D)


class _Localization(_Holder):
    def __init__(self, D):
        super().__init__()
        self.conv_layer_pm = nn.Conv2d(D, 1, 1)
        self.conv_layer_ps = nn.Conv1d(D, 1, 1)
        self.conv_layer_pe = nn.Conv1d(D, 1, 1)
        self.conv_layer_pa = nn.Conv1d(D, 1, 1)


def _round_up(x, m):
    return (x + m - 1) // m * m


# ---------------------------------------------------------------------------------------
# weight packing (layout transforms only; done once per parameter version)
# ---------------------------------------------------------------------------------------
def pack_weights(sd: Dict[str, torch.Tensor], dims: Dims, prec: int, device) -> Dict[str, torch.Tensor]:
    """Re-lay the reference-named parameters for the kernels: concatenated LSTM input
    projections, transposed recurrent weights, concatenated query-side projections,
    [W_fb | W_fc] for the moment unit, bf16 copies (K padded to a TMA-legal stride)."""
    f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
    H, D, dl, layers = dims.H, dims.D, dims.dl, dims.layers
    pk: Dict[str, torch.Tensor] = {}
    ve = "backbone.videoencoder."
    pk["ve_b"] = f32(sd[ve + "ve.bias"])
    pk["pe"] = f32(sd[ve + "pe.weight"])
    w_ve = f32(sd[ve + "ve.weight"])
    if prec == L_.BF16:
        kp = _round_up(dims.d0, 8)
        w = torch.zeros(D, kp, device=device, dtype=torch.bfloat16)
        w[:, : dims.d0] = w_ve.to(torch.bfloat16)
        pk["ve_w"] = w
        pk["ve_kpad"] = kp
    else:
        pk["ve_w"] = w_ve
        pk["ve_kpad"] = dims.d0
    ls = "backbone.queryencoder.lstm."
    for layer in range(2):
        wih = torch.cat([f32(sd[f"{ls}weight_ih_l{layer}"]), f32(sd[f"{ls}weight_ih_l{layer}_reverse"])], 0)
        bias = torch.cat([f32(sd[f"{ls}bias_ih_l{layer}"]) + f32(sd[f"{ls}bias_hh_l{layer}"]),
                          f32(sd[f"{ls}bias_ih_l{layer}_reverse"]) + f32(sd[f"{ls}bias_hh_l{layer}_reverse"])], 0)
        whh_t = torch.stack([f32(sd[f"{ls}weight_hh_l{layer}"]).t().contiguous(),
                             f32(sd[f"{ls}weight_hh_l{layer}_reverse"]).t().contiguous()], 0)
        pk[f"lstm_wih{layer}"], pk[f"lstm_b{layer}"], pk[f"lstm_whht{layer}"] = wih.contiguous(), bias, whh_t.contiguous()
    qw, qb = [], []
    for k in range(layers):
        cu, bu, mu = f"smis.{k}.content_unit.", f"smis.{k}.boundary_unit.attn_layer.", f"smis.{k}.moment_unit."
        qw += [f32(sd[cu + "linear_w_hat.weight"]), f32(sd[bu + "W_k.weight"])]
        qb += [f32(sd[cu + "linear_w_hat.bias"]), f32(sd[bu + "W_k.bias"])]
        for nm, key in (("ck_w", "attn_layer.W_k.weight"), ("ck_b", "attn_layer.W_k.bias"), ("cq_w", "attn_layer.W_q.weight"),
                        ("cq_b", "attn_layer.W_q.bias"), ("cs_w", "linear_s_hat.weight"), ("cs_b", "linear_s_hat.bias"),
                        ("chat_b", "linear_c_hat.bias"), ("cout_b", "linear_c.bias")):
            pk[f"{nm}{k}"] = f32(sd[cu + key])
        chat_w, cout_w = f32(sd[cu + "linear_c_hat.weight"]), f32(sd[cu + "linear_c.weight"])
        mu_w = torch.cat([f32(sd[mu + "conv_layer_fb.weight"]).view(D, D), f32(sd[mu + "conv_layer_fc.weight"]).view(D, D)], 1)
        pk[f"mu_b{k}"] = f32(sd[mu + "conv_layer_fb.bias"]) + f32(sd[mu + "conv_layer_fc.bias"])
        pk[f"bq_w{k}"], pk[f"bq_b{k}"] = f32(sd[bu + "W_q.weight"]), f32(sd[bu + "W_q.bias"])
        if prec == L_.BF16:
            chat_w, cout_w, mu_w = (t.to(torch.bfloat16).contiguous() for t in (chat_w, cout_w, mu_w))
        pk[f"chat_w{k}"], pk[f"cout_w{k}"], pk[f"mu_w{k}"] = chat_w, cout_w, mu_w.contiguous()
    pk["qcat_w"], pk["qcat_b"] = torch.cat(qw, 0).contiguous(), torch.cat(qb, 0).contiguous()
    lo = "localization.conv_layer_"
    pk["loc_w"] = torch.stack([f32(sd[lo + n + ".weight"]).reshape(D) for n in ("pm", "ps", "pe", "pa")], 0).contiguous()
    pk["loc_b"] = torch.cat([f32(sd[lo + n + ".bias"]).reshape(1) for n in ("pm", "ps", "pe", "pa")], 0).contiguous()
    return pk


class Workspace:
    """Named device buffers, reused across calls (no allocation in steady state)."""

    def __init__(self, device):
        self.device = device
        self.buf: Dict[str, torch.Tensor] = {}

    def get(self, name, shape, dtype):
        t = self.buf.get(name)
        n = 1
        for s in shape:
            n *= int(s)
        if t is None or t.dtype != dtype or t.numel() < n:
            t = torch.empty(max(n, 1), device=self.device, dtype=dtype)
            self.buf[name] = t
        return t[:n].view(*shape) if n else t[:0]


def make_cells(ws: Workspace, B: int, L: int, capacity: Optional[int] = None) -> Cells:
    cap = capacity if capacity is not None else B * (L * (L + 1) // 2)
    code = ws.get("cell_code", (cap,), torch.int32)
    row_start = ws.get("cell_row_start", (B * L + 1,), torch.int32)
    meta = ws.get("cell_meta", (2,), torch.int32)
    return Cells(code.data_ptr(), row_start.data_ptr(), meta.data_ptr(), meta.data_ptr() + 4, cap)


def smin_forward(pk: Dict[str, torch.Tensor], dims: Dims, prec: int, ws: Workspace, video_features, video_mask,
                 query_features, query_mask, length_mask, moment_mask, keep: Optional[dict] = None, mark=None):
    """The whole hot path, stage by stage.  ``keep`` (tests only)
    receives references to intermediates; ``mark(name)`` (bench only) is called after each
    stage has been enqueued so the caller can record CUDA events on the launching stream."""
    mark = mark or (lambda name: None)
    dev = video_features.device
    B = video_features.shape[0]
    T, Lm, Cc, D, dl, layers, d0, Nq, H = (dims.T, dims.L, dims.C, dims.D, dims.dl, dims.layers, dims.d0, dims.Nq, dims.H)
    st = stream_ptr()
    act = torch.bfloat16 if prec == L_.BF16 else torch.float32
    f32 = torch.float32
    vf = video_features.contiguous()
    qf = query_features.contiguous()
    vmask = video_mask.reshape(B, T).to(torch.uint8).contiguous()
    qmask = query_mask.reshape(B, Nq).to(torch.uint8).contiguous()
    lmask = length_mask.to(torch.uint8).contiguous()
    mmask = moment_mask.to(torch.uint8).contiguous()

    # ---- a1 clip projection --------------------------------------------------------------
    fv = ws.get("fv", (B * T, D), act)
    if prec == L_.BF16:
        kp = pk["ve_kpad"]
        v16 = ws.get("v16", (B * T, kp), torch.bfloat16)
        call("vml_cast_pad_bf16", ptr(vf), ptr(v16), B * T, d0, kp, st)
        mark("clip_cast")
        call("vml_clip_projection", ptr(v16), ptr(pk["ve_w"]), ptr(pk["ve_b"]), ptr(pk["pe"]), ptr(vmask), ptr(fv), B, dims, kp, prec, st)
    else:
        call("vml_clip_projection", ptr(vf), ptr(pk["ve_w"]), ptr(pk["ve_b"]), ptr(pk["pe"]), ptr(vmask), ptr(fv), B, dims, d0, prec, st)
    mark("clip_projection")

    # ---- a2 query encoder ------------------------------------------------------------------
    qlen = ws.get("qlen", (B,), torch.int32)
    call("vml_query_lengths", ptr(qmask), ptr(qlen), B, Nq, st)
    gin = ws.get("gin", (B * Nq, 8 * H), f32)
    y0 = ws.get("lstm_y0", (B, Nq, 2 * H), f32)
    fw = ws.get("fw", (B, Nq, 2 * H), f32)
    fs = ws.get("fs", (B, 2 * H), f32)
    call("vml_linear", ptr(qf), ptr(pk["lstm_wih0"]), ptr(pk["lstm_b0"]), ptr(gin), B * Nq, 8 * H, 300, 8 * H, None, 1, L_.FP32, 1, st)
    call("vml_lstm_layer", ptr(gin), ptr(pk["lstm_whht0"]), ptr(qlen), ptr(y0), None, None, B, Nq, H, st)
    call("vml_linear", ptr(y0), ptr(pk["lstm_wih1"]), ptr(pk["lstm_b1"]), ptr(gin), B * Nq, 8 * H, 2 * H, 8 * H, None, 1, L_.FP32, 1, st)
    call("vml_lstm_layer", ptr(gin), ptr(pk["lstm_whht1"]), ptr(qlen), ptr(fw), None, ptr(fs), B, Nq, H, st)
    mark("query_lstm")

    # query-side projections of every SMI layer, hoisted (fw / fs do not change across layers)
    ncat = layers * (dl + D)
    wproj = ws.get("wproj", (B * Nq, ncat), f32)
    call("vml_linear", ptr(fw), ptr(pk["qcat_w"]), ptr(pk["qcat_b"]), ptr(wproj), B * Nq, ncat, D, ncat, None, 1, L_.FP32, 1, st)
    w_hat = ws.get("w_hat", (layers, B, Nq, dl), f32)
    ktil = ws.get("ktil", (layers, B, Nq, dl), f32)
    beta = ws.get("beta", (layers, B, Nq), f32)
    s_hat = ws.get("s_hat", (layers, B, dl), f32)
    for k in range(layers):
        call("vml_query_prep", ptr(wproj), ncat, k * (dl + D), ptr(fs), ptr(qmask), ptr(pk[f"ck_w{k}"]), ptr(pk[f"ck_b{k}"]),
             ptr(pk[f"cq_w{k}"]), ptr(pk[f"cq_b{k}"]), ptr(pk[f"cs_w{k}"]), ptr(pk[f"cs_b{k}"]), ptr(w_hat[k]), ptr(ktil[k]),
             ptr(beta[k]), ptr(s_hat[k]), B, dims, st)
    mark("query_prep")

    # ---- cells + a3/a4 span pooling ---------------------------------------------------------
    cells = make_cells(ws, B, Lm)
    cap = cells.capacity
    call("vml_build_cells", ptr(mmask), B, Lm, cells, st)
    mark("build_cells")
    fc = [ws.get("fc_a", (cap, Cc, D), act), ws.get("fc_b", (cap, Cc, D), act)]
    fm = [ws.get("fm_a", (cap, D), act), ws.get("fm_b", (cap, D), act)]
    fb = [ws.get("fb_a", (B, Lm, D), f32), ws.get("fb_b", (B, Lm, D), f32)]
    call("vml_span_pool_fuse", ptr(fv), ptr(fs), cells, ptr(fc[0]), ptr(fm[0]), ptr(fb[0]), B, dims, prec, st)
    mark("span_pool_fuse")
    if keep is not None:
        keep.update(fv=fv, fs=fs, fw=fw, cells=cells, fc0=fc[0].clone(), fm0=fm[0].clone(), fb0=fb[0].clone())

    c_hat = ws.get("c_hat", (cap * Cc, dl), act)
    cc_hat = ws.get("cc_hat", (cap * Cc, dl), act)
    qb = ws.get("bu_q", (B * Lm, D), f32)
    g_scr = ws.get("bu_g", (B, Lm, D), f32)
    mu_op = ws.get("mu_op", (cap, 2 * D), act)
    n_dev = cells.n_cells
    cur = 0
    for k in range(layers):
        nxt = cur ^ 1
        # a7 boundary unit
        call("vml_linear", ptr(fb[cur]), ptr(pk[f"bq_w{k}"]), ptr(pk[f"bq_b{k}"]), ptr(qb), B * Lm, D, D, D, None, 1, L_.FP32, 1, st)
        call("vml_boundary_unit", ptr(qb), ptr(wproj), ncat, k * (dl + D) + dl, ptr(fw), ptr(fs), ptr(fb[cur]), ptr(fm[cur]),
             ptr(qmask), ptr(lmask), cells, ptr(g_scr), ptr(fb[nxt]), B, dims, prec, st)
        mark("boundary_unit")
        # a5+a6 content unit
        call("vml_linear", ptr(fc[cur]), ptr(pk[f"chat_w{k}"]), ptr(pk[f"chat_b{k}"]), ptr(c_hat), cap * Cc, dl, D, dl, n_dev, Cc,
             prec, 0, st)
        mark("content_in_gemm")
        call("vml_content_attention", ptr(c_hat), ptr(ktil[k]), ptr(beta[k]), ptr(w_hat[k]), ptr(s_hat[k]), ptr(qmask), cells,
             ptr(cc_hat), B, dims, prec, st)
        mark("content_attention")
        call("vml_content_out", ptr(cc_hat), ptr(pk[f"cout_w{k}"]), ptr(pk[f"cout_b{k}"]), ptr(fc[cur]), ptr(fm[cur]), ptr(fs),
             cells, ptr(fc[nxt]), dims, prec, st)
        mark("content_out_gemm")
        # a8 moment unit
        call("vml_moment_operand", ptr(fc[nxt]), ptr(fb[nxt]), cells, ptr(mu_op), dims, prec, st)
        mark("moment_operand")
        call("vml_moment_out", ptr(mu_op), ptr(pk[f"mu_w{k}"]), ptr(pk[f"mu_b{k}"]), ptr(fm[cur]), cells, ptr(fm[nxt]), dims, prec, st)
        mark("moment_out_gemm")
        cur = nxt
        if keep is not None:
            keep[f"fc{k + 1}"], keep[f"fm{k + 1}"], keep[f"fb{k + 1}"] = fc[cur].clone(), fm[cur].clone(), fb[cur].clone()

    # ---- a9 localization ----------------------------------------------------------------------
    pm = torch.empty(B, Lm, Lm, device=dev, dtype=f32)
    ps = torch.empty(B, Lm, device=dev, dtype=f32)
    pe = torch.empty(B, Lm, device=dev, dtype=f32)
    pa = torch.empty(B, Lm, device=dev, dtype=f32)
    call("vml_localize", ptr(fm[cur]), ptr(fb[cur]), ptr(pk["loc_w"]), ptr(pk["loc_b"]), cells, ptr(lmask), ptr(pm), ptr(ps),
         ptr(pe), ptr(pa), B, dims, prec, st)
    mark("localize")
    return pm, ps, pe, pa


class SMIN(nn.Module):
    """B200-native SMIN.  Interface of the reference ``models.SMIN`` (models.py:348,367)."""

    def __init__(self, T, L, C, D, dl, num_smi_layers, input_video_dim, max_query_length, lstm_hidden_size,
                 device="cpu", precision: str = "bf16"):
        super().__init__()
        if D != 2 * lstm_hidden_size:
            raise ValueError("D must equal 2*lstm_hidden_size (models.py:81 multiplies fv[B,T,D] by fs[B,2H])")
        if T % L != 0:
            raise ValueError("L must divide T (models.py:93,113)")
        self.T, self.L, self.C, self.D, self.dl = T, L, C, D, dl
        self.num_smi_layers, self.input_video_dim = num_smi_layers, input_video_dim
        self.max_query_length, self.lstm_hidden_size, self.device = max_query_length, lstm_hidden_size, device
        self.precision = precision
        # construction order mirrors the reference so the RNG stream yields the same init
        self.backbone = _Backbone(T, D, input_video_dim, lstm_hidden_size)
        self.pgm = _Holder()
        self.smis = nn.ModuleList([_SMI(D, dl) for _ in range(num_smi_layers)])
        self.localization = _Localization(D)
        self._dims = Dims(T, L, C, D, dl, num_smi_layers, input_video_dim, max_query_length, lstm_hidden_size)
        self._packed = None
        self._packed_key = None
        self._ws: Dict[str, Workspace] = {}

    # -- packed weights, refreshed when any parameter changed (optimizer step / load_state_dict)
    def _weights(self, device, prec):
        key = (str(device), prec, tuple((p.data_ptr(), p._version) for p in self.parameters()))
        if self._packed is None or self._packed_key != key:
            self._packed = pack_weights(self.state_dict(), self._dims, prec, device)
            self._packed_key = key
        return self._packed

    def forward(self, video_features, video_mask, query_features, query_mask, length_mask, moment_mask, mark=None):
        if not video_features.is_cuda:
            raise L_.VmlError("vml_b200.SMIN runs on CUDA (sm_100a) only; there is no CPU path. "
                              "Move the module and its inputs to a B200 device.")
        L_.load()
        prec = L_.PREC[self.precision]
        dev = video_features.device
        if video_features.shape[0] >= 32768:
            raise ValueError("batch size must be < 32768")
        with torch.no_grad():
            pk = self._weights(dev, prec)
            ws = self._ws.setdefault(str(dev), Workspace(dev))
            return smin_forward(pk, self._dims, prec, ws, video_features.float(), video_mask, query_features.float(),
                                query_mask, length_mask, moment_mask, mark=mark)
