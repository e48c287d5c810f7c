"""Training path of the drop-in SMIN: forward with saved activations + hand-written backward.

``SMIN.forward`` routes here when autograd is recording and a parameter requires grad (the reference's
``train_epoch``: ``loss.backward(); optimizer.step()``, main.py:136-150).  The whole step runs in fp32
(CUDA-core contractions, 3xTF32 in the boundary unit): forward kernels are the validation-mode ones plus
two that save activations (the LSTM and the boundary gate); the backward is ``csrc/backward.cu`` plus
``vml_gemm_strided`` for every dense product.  PyTorch only provides the autograd hook
(``torch.autograd.Function``), device memory, and -- on small weight-space matrices only -- the chain rule
through the algebraic folding of the query-side weights (``smin.fold_query_weights``).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List

import torch

from . import lib as L_
from .lib import Cells, Dims, call, ptr, stream_ptr
from .smin import _round_up, fold_query_weights, query_layout

F32 = torch.float32


def _gemm(A, sam, sak, sab, B, sbn, sbk, sbb, C, scm, scn, scb, M, N, K, batch=1, alpha=1.0, acc=0, splits=1, m_dev=None,
          m_scale=1, k_dev=None, k_scale=1):
    """C[b][m][n] (=|+=) alpha * sum_k A[b][m][k] * B[b][n][k]; A/B/C are raw device pointers (ints)."""
    call("vml_gemm_strided", A, sam, sak, sab, B, sbn, sbk, sbb, C, scm, scn, scb, M, N, K, batch, alpha, acc, splits,
         m_dev, m_scale, k_dev, k_scale, stream_ptr())


def _splits(K: int) -> int:
    return max(1, min(64, K // 2048))


class Tape:
    """Everything the backward pass needs from one forward pass (device tensors + the cell list)."""
    pass


class ZeroArena:
    """Zero-initialised fp32 buffers of one backward pass carved out of ONE allocation with ONE memset (the pass needs
    ~60 of them; a torch.zeros each was 60 fill launches).  The size is learnt from the previous step; a request that
    does not fit falls back to torch.zeros and enlarges the next arena."""
    _want: Dict[str, int] = {}

    def __init__(self, key: str, device):
        self.key, self.device, self.used = key, device, 0
        n = ZeroArena._want.get(key, 0)
        self.buf = torch.zeros(n, device=device, dtype=F32) if n else None

    def zeros(self, *shape):
        n = 1
        for d in shape:
            n *= int(d)
        n_al = (n + 63) // 64 * 64                       # 256-byte aligned pieces
        lo = self.used
        self.used += n_al
        if self.buf is not None and lo + n <= self.buf.numel():
            return self.buf[lo:lo + n].view(*shape)
        return torch.zeros(*shape, device=self.device, dtype=F32)

    def close(self):
        ZeroArena._want[self.key] = max(ZeroArena._want.get(self.key, 0), self.used)


def train_forward(pk: Dict[str, torch.Tensor], dims: Dims, inp: dict) -> tuple:
    """fp32 forward that keeps its intermediates.  ``inp`` comes from ``smin_ingest(..., prec=FP32)``."""
    P = L_.FP32
    # the dense products of the forward run on the tensor cores as TF32 (fp32 tensors as they are, fp32 accumulation) like the
    # backward's (gemm_tf32.cu); VML_TRAIN_FP32=1 keeps every product on CUDA-core FFMA (bit-for-bit the validation path)
    PG = L_.FP32 if os.environ.get("VML_TRAIN_FP32") else L_.TF32
    B = inp["B"]
    T, Lm, Cc, D, dl, layers, d0, Nq, H = (dims.T, dims.L, dims.C, dims.D, dims.dl, dims.layers, dims.d0, dims.Nq, dims.H)
    dev = inp["qlen"].device
    st = stream_ptr()
    E = lambda *shape: torch.empty(*shape, device=dev, dtype=F32)
    tp = Tape()
    tp.inp, tp.B = inp, B
    vmask, qmask, lmask, mmask, qlen = inp["vmask"], inp["qmask"], inp["lmask"], inp["mmask"], inp["qlen"]
    # a1
    tp.fv = E(B * T, D)
    call("vml_clip_projection", ptr(inp["v"]), ptr(pk["ve_w"]), ptr(pk["ve_b"]), ptr(pk["pe"]), ptr(vmask), ptr(tp.fv), B, dims, d0, PG, st)
    # a2 (saving gate activations)
    gin = E(B * Nq, 8 * H)
    call("vml_linear", ptr(inp["q"]), ptr(pk["lstm_wih0"]), ptr(pk["lstm_b0"]), ptr(gin), B * Nq, 8 * H, 300, 8 * H, None, 1, PG, 1, st)
    tp.y0, tp.acts0 = E(B, Nq, 2 * H), torch.zeros(B, Nq, 2, 5, H, device=dev, dtype=F32)
    call("vml_lstm_train_fwd", ptr(gin), ptr(pk["lstm_whht0"]), ptr(qlen), ptr(tp.y0), None, ptr(tp.acts0), B, Nq, H, st)
    gin1 = E(B * Nq, 8 * H)
    call("vml_linear", ptr(tp.y0), ptr(pk["lstm_wih1"]), ptr(pk["lstm_b1"]), ptr(gin1), B * Nq, 8 * H, 2 * H, 8 * H, None, 1, PG, 1, st)
    tp.fwfs, tp.acts1 = E(B * Nq + B, 2 * H), torch.zeros(B, Nq, 2, 5, H, device=dev, dtype=F32)
    fw, fs = tp.fwfs[: B * Nq], tp.fwfs[B * Nq:]
    call("vml_lstm_train_fwd", ptr(gin1), ptr(pk["lstm_whht1"]), ptr(qlen), ptr(fw), ptr(fs), ptr(tp.acts1), B, Nq, H, st)
    lay = query_layout(dims)
    ld = lay["ld"]
    tp.qproj = E(B * Nq + B, ld)
    call("vml_linear", ptr(tp.fwfs), ptr(pk["qcat_w"]), ptr(pk["qcat_b"]), ptr(tp.qproj), B * Nq + B, ld, D, ld, None, 1, PG, 1, st)
    s_hat_base = tp.qproj.data_ptr() + (B * Nq * ld + lay["s0"]) * 4
    # cells + a3/a4
    cap = B * (Lm * (Lm + 1) // 2)
    tp.cell_code = torch.empty(cap, device=dev, dtype=torch.int32)
    tp.cell_rows = torch.empty(B * Lm + 1, device=dev, dtype=torch.int32)
    tp.cell_meta = torch.zeros(2, device=dev, dtype=torch.int32)
    cells = Cells(tp.cell_code.data_ptr(), tp.cell_rows.data_ptr(), tp.cell_meta.data_ptr(), tp.cell_meta.data_ptr() + 4, cap)
    tp.cells, tp.cap = cells, cap
    call("vml_build_cells", ptr(mmask), B, Lm, cells, st)
    tp.fc, tp.fm, tp.fb = [E(cap, Cc, D)], [E(cap, D)], [E(B, Lm, D)]
    call("vml_span_pool_fuse", ptr(tp.fv), ptr(fs), cells, ptr(tp.fc[0]), ptr(tp.fm[0]), ptr(tp.fb[0]), B, dims, P, st)
    tp.G, tp.Ab, tp.Pw, tp.U, tp.c_hat, tp.cc_hat, tp.op = [], [], [], [], [], [], []
    g_dummy = None
    for k in range(layers):
        o = k * lay["blk"]
        G, Ab, bu = E(B, Lm, D), E(B, Lm, Lm), E(B, Lm, D)
        Pw, U = torch.zeros(B, Lm, Nq, device=dev, dtype=F32), E(B, Lm, D)
        call("vml_boundary_unit", ptr(tp.qproj), ld, o + 2 * dl, o + 2 * dl + D + 1, ptr(fw), ptr(fs), ptr(tp.fb[k]), ptr(tp.fm[k]),
             ptr(qmask), ptr(lmask), cells, ptr(G), ptr(Ab), ptr(bu), None, None, ptr(Pw), ptr(U), B, dims, P, st)
        c_hat, cc_hat, cu = E(cap * Cc, dl), E(cap * Cc, dl), E(cap, Cc, D)
        call("vml_linear", ptr(tp.fc[k]), ptr(pk[f"chat_w{k}"]), ptr(pk[f"chat_b{k}"]), ptr(c_hat), cap * Cc, dl, D, dl,
             cells.n_cells, Cc, PG, 0, st)
        call("vml_content_attention", ptr(c_hat), ptr(tp.qproj), ld, o, o + dl, o + 2 * dl + D, s_hat_base + k * dl * 4, ld,
             ptr(qmask), cells, ptr(cc_hat), B, dims, P, st)
        call("vml_content_out", ptr(cc_hat), ptr(pk[f"cout_w{k}"]), ptr(pk[f"cout_b{k}"]), ptr(tp.fc[k]), ptr(tp.fm[k]), ptr(fs),
             None, None, cells, ptr(cu), dims, PG, st)
        op, mu = E(cap, 2 * D), E(cap, D)
        call("vml_moment_operand", ptr(cu), ptr(bu), cells, ptr(op), dims, P, st)
        call("vml_moment_out", ptr(op), ptr(pk[f"mu_w{k}"]), ptr(pk[f"mu_b{k}"]), ptr(tp.fm[k]), cells, ptr(mu), dims, PG, st)
        tp.G.append(G); tp.Ab.append(Ab); tp.Pw.append(Pw); tp.U.append(U)
        tp.c_hat.append(c_hat); tp.cc_hat.append(cc_hat); tp.op.append(op)
        tp.fc.append(cu); tp.fm.append(mu); tp.fb.append(bu)
    pm = torch.empty(B, Lm, Lm, device=dev, dtype=F32)
    ps, pe, pa = E(B, Lm), E(B, Lm), E(B, Lm)
    call("vml_localize", ptr(tp.fm[-1]), ptr(tp.fb[-1]), ptr(pk["loc_w"]), ptr(pk["loc_b"]), cells, ptr(lmask), ptr(pm), ptr(ps),
         ptr(pe), ptr(pa), B, dims, P, st)
    tp.out = (pm, ps, pe, pa)
    return (pm, ps, pe, pa), tp


def train_backward(pk: Dict[str, torch.Tensor], dims: Dims, tp: Tape, g_pm, g_ps, g_pe, g_pa) -> Dict[str, torch.Tensor]:
    """Gradients of the PACKED parameters (keys of ``pk`` that are trainable views) given d loss / d outputs."""
    B = tp.B
    T, Lm, Cc, D, dl, layers, d0, Nq, H = (dims.T, dims.L, dims.C, dims.D, dims.dl, dims.layers, dims.d0, dims.Nq, dims.H)
    inp, cells, cap = tp.inp, tp.cells, tp.cap
    dev = tp.fv.device
    st = stream_ptr()
    arena = ZeroArena(f"bwd:{B}:{cap}:{dev}", dev)
    Z = arena.zeros
    E = lambda *shape: torch.empty(*shape, device=dev, dtype=F32)
    vmask, qmask, lmask, qlen = inp["vmask"], inp["qmask"], inp["lmask"], inp["qlen"]
    lay = query_layout(dims)
    ld = lay["ld"]
    n_dev = cells.n_cells
    fw, fs = tp.fwfs[: B * Nq], tp.fwfs[B * Nq:]
    R = B * Nq + B
    g: Dict[str, torch.Tensor] = {}
    dq = Z(R, ld)                       # gradient of the folded query projection (same layout as qproj)
    dfw_direct, dfs_direct = Z(B * Nq, D), Z(B, D)
    inv_sqrt_d = 1.0 / math.sqrt(D)
    pm, ps, pe, pa = tp.out
    gc = lambda t: t.detach().to(F32).contiguous()
    g_pm, g_ps, g_pe, g_pa = gc(g_pm), gc(g_ps), gc(g_pe), gc(g_pa)

    # ---- a9 ------------------------------------------------------------------------------------------------
    g["loc_w"], g["loc_b"] = Z(4, D), Z(4)
    d_fm, d_fb = Z(cap, D), E(B, Lm, D)
    call("vml_localize_bwd", ptr(tp.fm[-1]), ptr(tp.fb[-1]), ptr(pk["loc_w"]), ptr(pm), ptr(ps), ptr(pe), ptr(pa), ptr(g_pm), ptr(g_ps),
         ptr(g_pe), ptr(g_pa), ptr(lmask), cells, ptr(d_fm), ptr(d_fb), ptr(g["loc_w"]), ptr(g["loc_b"]), B, dims, st)
    d_fc_next = None
    ones = torch.ones(max(Lm, Nq, 8), device=dev, dtype=F32)
    for k in reversed(range(layers)):
        o = k * lay["blk"]
        fc_k, fm_k, fb_k, bu = tp.fc[k], tp.fm[k], tp.fb[k], tp.fb[k + 1]
        # ---- a8: mu = op.Wcat^T + b + fm ------------------------------------------------------------------------
        d_op = Z(cap, 2 * D)
        _gemm(ptr(d_fm), D, 1, 0, ptr(pk[f"mu_w{k}"]), 1, 2 * D, 0, ptr(d_op), 2 * D, 1, 0, cap, 2 * D, D, m_dev=n_dev)
        g[f"mu_w{k}"], g[f"mu_b{k}"] = Z(D, 2 * D), Z(D)
        _gemm(ptr(d_fm), 1, D, 0, ptr(tp.op[k]), 1, 2 * D, 0, ptr(g[f"mu_w{k}"]), 2 * D, 1, 0, D, 2 * D, cap, acc=1, splits=_splits(cap),
              k_dev=n_dev)
        call("vml_colsum", ptr(d_fm), D, 0, ptr(g[f"mu_b{k}"]), 0, cap, D, 1, n_dev, 1, 1.0, st)
        call("vml_pair_bwd", ptr(d_op), 2 * D, ptr(bu), cells, ptr(d_fb), B, dims, st)            # d_bu (in d_fb) += pair terms
        dY, d_gbar = Z(cap * Cc, D), Z(cap, D)
        call("vml_cu_tail_bwd", ptr(d_fc_next), ptr(d_op), 2 * D, cells, ptr(dY), ptr(d_gbar), dims, st)
        # ---- a6 tail: cu = cc_hat.Wc^T + bc + fc + gbar -----------------------------------------------------------
        d_cc = Z(cap * Cc, dl)
        _gemm(ptr(dY), D, 1, 0, ptr(pk[f"cout_w{k}"]), 1, dl, 0, ptr(d_cc), dl, 1, 0, cap * Cc, dl, D, m_dev=n_dev, m_scale=Cc)
        g[f"cout_w{k}"], g[f"cout_b{k}"] = Z(D, dl), Z(D)
        _gemm(ptr(dY), 1, D, 0, ptr(tp.cc_hat[k]), 1, dl, 0, ptr(g[f"cout_w{k}"]), dl, 1, 0, D, dl, cap * Cc, acc=1,
              splits=_splits(cap * Cc), k_dev=n_dev, k_scale=Cc)
        call("vml_colsum", ptr(dY), D, 0, ptr(g[f"cout_b{k}"]), 0, cap * Cc, D, 1, n_dev, Cc, 1.0, st)
        # ---- a5/a6 attention block ----------------------------------------------------------------------------------
        d_chat = Z(cap * Cc, dl)
        s_hat_base = tp.qproj.data_ptr() + (B * Nq * ld + lay["s0"] + k * dl) * 4
        d_shat_base = dq.data_ptr() + (B * Nq * ld + lay["s0"] + k * dl) * 4
        call("vml_content_attn_bwd", ptr(tp.c_hat[k]), ptr(d_cc), ptr(tp.qproj), ld, o, o + dl, o + 2 * dl + D, s_hat_base, ld, ptr(qmask),
             cells, ptr(d_chat), ptr(dq), d_shat_base, B, dims, st)
        # d fc_k = dY + d_chat.W_chat ; parameter gradients of linear_c_hat
        _gemm(ptr(d_chat), dl, 1, 0, ptr(pk[f"chat_w{k}"]), 1, D, 0, ptr(dY), D, 1, 0, cap * Cc, D, dl, acc=1, m_dev=n_dev, m_scale=Cc)
        g[f"chat_w{k}"], g[f"chat_b{k}"] = Z(dl, D), Z(dl)
        _gemm(ptr(d_chat), 1, dl, 0, ptr(fc_k), 1, D, 0, ptr(g[f"chat_w{k}"]), D, 1, 0, dl, D, cap * Cc, acc=1, splits=_splits(cap * Cc),
              k_dev=n_dev, k_scale=Cc)
        call("vml_colsum", ptr(d_chat), dl, 0, ptr(g[f"chat_b{k}"]), 0, cap * Cc, dl, 1, n_dev, Cc, 1.0, st)
        # ---- a7 boundary unit: bu = A_b.fb*lm + fb + sum_j A_b gbar ----------------------------------------------------
        d_bu = d_fb                                              # [B,L,D] total gradient of this layer's bu
        dbuL = E(B, Lm, D)
        call("vml_mask_rows", ptr(d_bu), ptr(lmask), ptr(dbuL), B * Lm, D, 0, st)
        dAb = E(B, Lm, Lm)
        _gemm(ptr(dbuL), D, 1, Lm * D, ptr(fb_k), D, 1, Lm * D, ptr(dAb), Lm, 1, Lm * Lm, Lm, Lm, D, batch=B)
        d_fm_k = Z(cap, D)
        call("vml_gbar_bwd", ptr(fm_k), ptr(fs), ptr(tp.Ab[k]), ptr(d_bu), ptr(d_gbar), ptr(d_fm), cells, ptr(dAb), ptr(d_fm_k),
             ptr(dfs_direct), B, dims, st)
        d_fb_k = d_bu.clone()                                    # residual "+ f_b"
        _gemm(ptr(tp.Ab[k]), 1, Lm, Lm * Lm, ptr(dbuL), 1, D, Lm * D, ptr(d_fb_k), D, 1, Lm * D, Lm, D, Lm, batch=B, acc=1)
        dS = E(B, Lm, Lm)
        call("vml_softmax_bwd", ptr(tp.Ab[k]), ptr(dAb), ptr(lmask), ptr(dS), B, Lm, Lm, inv_sqrt_d, st)
        dG = E(B, Lm, D)
        _gemm(ptr(dS), Lm, 1, Lm * Lm, ptr(tp.G[k]), 1, D, Lm * D, ptr(dG), D, 1, Lm * D, Lm, D, Lm, batch=B)
        _gemm(ptr(dS), 1, Lm, Lm * Lm, ptr(tp.G[k]), 1, D, Lm * D, ptr(dG), D, 1, Lm * D, Lm, D, Lm, batch=B, acc=1)
        dAq, tmp = E(B, Lm, D), E(B, Lm, D)
        call("vml_gate_bwd", ptr(dG), ptr(fb_k), ptr(tp.U[k]), ptr(lmask), ptr(d_fb_k), ptr(dAq), ptr(tmp), B, dims, st)
        call("vml_colsum", ptr(tmp), D, Lm * D, ptr(dfs_direct), D, Lm, D, B, None, 1, 1.0, st)
        dP = E(B, Lm, Nq)
        _gemm(ptr(dAq), D, 1, Lm * D, ptr(fw), D, 1, Nq * D, ptr(dP), Nq, 1, Lm * Nq, Lm, Nq, D, batch=B)
        _gemm(ptr(tp.Pw[k]), 1, Nq, Lm * Nq, ptr(dAq), 1, D, Lm * D, ptr(dfw_direct), D, 1, Nq * D, Nq, D, Lm, batch=B, acc=1)
        dsc = E(B, Lm, Nq)
        call("vml_softmax_bwd", ptr(tp.Pw[k]), ptr(dP), ptr(qmask), ptr(dsc), B, Lm, Nq, inv_sqrt_d, st)
        kbt = tp.qproj.data_ptr() + (o + 2 * dl) * 4
        dkbt = dq.data_ptr() + (o + 2 * dl) * 4
        dbetab = dq.data_ptr() + (o + 2 * dl + D + 1) * 4
        _gemm(ptr(dsc), Nq, 1, Lm * Nq, kbt, 1, ld, Nq * ld, ptr(d_fb_k), D, 1, Lm * D, Lm, D, Nq, batch=B, acc=1)
        _gemm(ptr(dsc), 1, Nq, Lm * Nq, ptr(fb_k), 1, D, Lm * D, dkbt, ld, 1, Nq * ld, Nq, D, Lm, batch=B, acc=1)
        _gemm(ptr(dsc), 1, Nq, Lm * Nq, ptr(ones), 0, 1, 0, dbetab, ld, 1, Nq * ld, Nq, 1, Lm, batch=B, acc=1)
        d_fb, d_fm, d_fc_next = d_fb_k, d_fm_k, dY

    # ---- a3/a4 + a1 ---------------------------------------------------------------------------------------------
    d_fv = E(B * T, D)
    call("vml_span_pool_bwd", ptr(d_fc_next), ptr(d_fm), ptr(d_fb), ptr(tp.fv), ptr(fs), cells, ptr(d_fv), ptr(dfs_direct), B, dims, st)
    call("vml_mask_rows", ptr(d_fv), ptr(vmask), ptr(d_fv), B * T, D, 0, st)                       # dz = d_fv * video_mask
    g["ve_w"], g["ve_b"], g["pe"] = Z(D, d0), Z(D), Z(T, D)
    _gemm(ptr(d_fv), 1, D, 0, ptr(inp["v"]), 1, d0, 0, ptr(g["ve_w"]), d0, 1, 0, D, d0, B * T, acc=1, splits=_splits(B * T))
    call("vml_colsum", ptr(d_fv), D, 0, ptr(g["ve_b"]), 0, B * T, D, 1, None, 1, 1.0, st)
    call("vml_colsum", ptr(d_fv), T * D, D, ptr(g["pe"]), D, B, D, T, None, 1, 1.0, st)
    # ---- folded query projection -------------------------------------------------------------------------------------
    d_fwfs = E(R, D)
    _gemm(ptr(dq), ld, 1, 0, ptr(pk["qcat_w"]), 1, D, 0, ptr(d_fwfs), D, 1, 0, R, D, ld)
    d_fwfs[: B * Nq] += dfw_direct
    d_fwfs[B * Nq:] += dfs_direct
    g["qcat_w"], g["qcat_b"] = Z(ld, D), Z(ld)
    _gemm(ptr(dq), 1, ld, 0, ptr(tp.fwfs), 1, D, 0, ptr(g["qcat_w"]), D, 1, 0, ld, D, R, acc=1, splits=1)
    call("vml_colsum", ptr(dq), ld, 0, ptr(g["qcat_b"]), 0, R, ld, 1, None, 1, 1.0, st)
    # ---- a2: two bi-LSTM layers, BPTT -------------------------------------------------------------------------------------
    rows = B * Nq

    def lstm_layer_bwd(layer, dy, dfs, acts, x, xk, y_own):
        dgin, dgin_rec = E(rows, 8 * H), E(rows, 8 * H)
        call("vml_lstm_train_bwd", ptr(dy), ptr(dfs) if dfs is not None else None, ptr(pk[f"lstm_whh{layer}"]), ptr(acts), ptr(qlen),
             ptr(dgin), ptr(dgin_rec), B, Nq, H, st)
        g[f"lstm_wih{layer}"], g[f"lstm_b{layer}"], g[f"lstm_whh{layer}"] = Z(8 * H, xk), Z(8 * H), Z(2, 4 * H, H)
        _gemm(ptr(dgin), 1, 8 * H, 0, ptr(x), 1, xk, 0, ptr(g[f"lstm_wih{layer}"]), xk, 1, 0, 8 * H, xk, rows, acc=1, splits=1)
        call("vml_colsum", ptr(dgin), 8 * H, 0, ptr(g[f"lstm_b{layer}"]), 0, rows, 8 * H, 1, None, 1, 1.0, st)
        whh_g = g[f"lstm_whh{layer}"]
        # forward direction: h_prev of row r is y[r-1, :H];  reverse direction: h_prev of row r is y[r+1, H:]
        _gemm(dgin_rec.data_ptr() + 8 * H * 4, 1, 8 * H, 0, y_own.data_ptr(), 1, 2 * H, 0, whh_g.data_ptr(), H, 1, 0, 4 * H, H, rows - 1, acc=1)
        _gemm(dgin_rec.data_ptr() + 4 * H * 4, 1, 8 * H, 0, y_own.data_ptr() + (2 * H + H) * 4, 1, 2 * H, 0, whh_g.data_ptr() + 4 * H * H * 4,
              H, 1, 0, 4 * H, H, rows - 1, acc=1)
        return dgin

    dgin1 = lstm_layer_bwd(1, d_fwfs[: B * Nq], d_fwfs[B * Nq:], tp.acts1, tp.y0, 2 * H, tp.fwfs)
    d_y0 = E(rows, 2 * H)
    _gemm(ptr(dgin1), 8 * H, 1, 0, ptr(pk["lstm_wih1"]), 1, 2 * H, 0, ptr(d_y0), 2 * H, 1, 0, rows, 2 * H, 8 * H)
    lstm_layer_bwd(0, d_y0, None, tp.acts0, inp["q"], 300, tp.y0)
    arena.close()
    return g


def unpack_grads(g: Dict[str, torch.Tensor], params: Dict[str, torch.Tensor], dims: Dims) -> Dict[str, torch.Tensor]:
    """Packed-parameter gradients -> gradients keyed like the reference ``state_dict`` (models.py:21-23,46,...)."""
    D, dl, H, layers = dims.D, dims.dl, dims.H, dims.layers
    out: Dict[str, torch.Tensor] = {}
    ve = "backbone.videoencoder."
    out[ve + "ve.weight"], out[ve + "ve.bias"], out[ve + "pe.weight"] = g["ve_w"], g["ve_b"], g["pe"]
    ls = "backbone.queryencoder.lstm."
    for layer in range(2):
        wih, b, whh = g[f"lstm_wih{layer}"], g[f"lstm_b{layer}"], g[f"lstm_whh{layer}"]
        for di, suf in enumerate(("", "_reverse")):
            out[f"{ls}weight_ih_l{layer}{suf}"] = wih[di * 4 * H:(di + 1) * 4 * H]
            out[f"{ls}weight_hh_l{layer}{suf}"] = whh[di]
            out[f"{ls}bias_ih_l{layer}{suf}"] = b[di * 4 * H:(di + 1) * 4 * H]
            out[f"{ls}bias_hh_l{layer}{suf}"] = b[di * 4 * H:(di + 1) * 4 * H]
    for k in range(layers):
        cu, mu = f"smis.{k}.content_unit.", f"smis.{k}.moment_unit."
        out[cu + "linear_c_hat.weight"], out[cu + "linear_c_hat.bias"] = g[f"chat_w{k}"], g[f"chat_b{k}"]
        out[cu + "linear_c.weight"], out[cu + "linear_c.bias"] = g[f"cout_w{k}"], g[f"cout_b{k}"]
        out[mu + "conv_layer_fb.weight"] = g[f"mu_w{k}"][:, :D].reshape(D, D, 1, 1)
        out[mu + "conv_layer_fc.weight"] = g[f"mu_w{k}"][:, D:].reshape(D, D, 1, 1)
        out[mu + "conv_layer_fb.bias"], out[mu + "conv_layer_fc.bias"] = g[f"mu_b{k}"], g[f"mu_b{k}"]
    lo = "localization.conv_layer_"
    out[lo + "pm.weight"] = g["loc_w"][0].reshape(1, D, 1, 1)
    for i, n in enumerate(("ps", "pe", "pa")):
        out[lo + n + ".weight"] = g["loc_w"][i + 1].reshape(1, D, 1)
    for i, n in enumerate(("pm", "ps", "pe", "pa")):
        out[lo + n + ".bias"] = g["loc_b"][i:i + 1]
    # chain rule through the weight folding (weight-space matrices only; fp64)
    leaf = {k: v.detach().to(torch.float64).requires_grad_(True) for k, v in params.items()
            if k.startswith("smis.") and any(f in k for f in ("linear_w_hat", "attn_layer.W_k", "attn_layer.W_q", "linear_s_hat"))}
    with torch.enable_grad():
        qw, qb = fold_query_weights(leaf, dims)
        names = list(leaf.keys())
        grads = torch.autograd.grad([qw, qb], [leaf[n] for n in names],
                                    [g["qcat_w"].to(torch.float64), g["qcat_b"].to(torch.float64)], allow_unused=True)
    for n, gr in zip(names, grads):
        out[n] = (gr if gr is not None else torch.zeros_like(leaf[n])).to(F32)
    return out


class SminTrainFunction(torch.autograd.Function):
    """autograd hook: forward = ``train_forward``; backward = ``train_backward`` + ``unpack_grads``."""

    @staticmethod
    def forward(ctx, model, inp, names, *params):
        pk = model._weights(params[0].device, L_.FP32)
        out, tape = train_forward(pk, model._dims, inp)
        ctx.model, ctx.tape, ctx.names, ctx.pk = model, tape, names, pk
        ctx.params = {n: p for n, p in zip(names, params)}
        return out

    @staticmethod
    def backward(ctx, g_pm, g_ps, g_pe, g_pa):
        tape, dims = ctx.tape, ctx.model._dims
        zeros = lambda ref: torch.zeros_like(ref)
        pm, ps, pe, pa = tape.out
        g = train_backward(ctx.pk, dims, tape, g_pm if g_pm is not None else zeros(pm), g_ps if g_ps is not None else zeros(ps),
                           g_pe if g_pe is not None else zeros(pe), g_pa if g_pa is not None else zeros(pa))
        named = unpack_grads(g, ctx.params, dims)
        grads: List[torch.Tensor] = []
        for n in ctx.names:
            gr = named.get(n)
            grads.append(None if gr is None else gr.reshape(ctx.params[n].shape).contiguous())
        ctx.tape = None
        return (None, None, None, *grads)
