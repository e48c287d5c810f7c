"""ctypes binding of libvml_b200.so (C ABI declared in include/vml_b200.h).

There is no fallback: ``load()`` raises if the shared library has not been built
(``python -m vml_b200.build`` / ``__graft_entry__.build()``), and every wrapper raises
``VmlError`` when a launcher returns a negative status.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VML_LIB") or os.path.join(HERE, "libvml_b200.so")      # VML_LIB: e.g. the bounds-checking debug build
HEADER = os.path.join(os.path.dirname(HERE), "include", "vml_b200.h")

FP32, BF16, TF32 = 0, 1, 2          # TF32: fp32 tensors, dense products as tcgen05 kind::tf32 (training path only)
PREC = {"fp32": FP32, "bf16": BF16}


class VmlError(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("T", "L", "C", "D", "dl", "layers", "d0", "Nq", "H")]


class Cells(C.Structure):
    _fields_ = [("code", C.c_void_p), ("row_start", C.c_void_p), ("n_cells", C.c_void_p),
                ("status", C.c_void_p), ("capacity", C.c_int32)]


_P, _I, _I64 = C.c_void_p, C.c_int, C.c_int64

# name -> argtypes (return type is int unless listed in _RET)
_SIGS = {
    "vml_last_error": [],
    "vml_version": [],
    "vml_launch_count": [],
    "vml_kernel_names": [],
    "vml_build_cells": [_P, _I, _I, Cells, _P],
    "vml_unpack_cells": [_P, _P, Cells, _I, _I, _I, _I, _P],
    "vml_pack_cells": [_P, _P, Cells, _I, _I, _I, _I, _P],
    "vml_cast_pad_bf16": [_P, _P, _I64, _I, _I, _P],
    "vml_ingest": [_P] * 15 + [_I, Dims, _I, _I, _I, _P],
    "vml_ingest_bf16": [_P] * 15 + [_I, Dims, _I, _I, _I, _P],
    "vml_ingest_packed": [_P] * 16 + [_I, Dims, _I, _I, _I, _I, _P],
    "vml_gemm_strided": [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _I, _I, _I, _I, C.c_float, _I, _I,
                         _P, _I, _P, _I, _P],
    "vml_linear": [_P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _I, _I, _P],
    "vml_clip_projection": [_P, _P, _P, _P, _P, _P, _I, Dims, _I, _I, _P],
    "vml_lstm_layer": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "vml_lstm_layer_tc": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "vml_query_lengths": [_P, _P, _I, _I, _P],
    "vml_span_pool_fuse": [_P, _P, Cells, _P, _P, _P, _I, Dims, _I, _P],
    "vml_content_attention": [_P, _P, _I, _I, _I, _I, _P, _I, _P, Cells, _P, _I, Dims, _I, _P],
    "vml_content_in_attention": [_P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P, Cells, _P, _I, Dims, _P],
    "vml_copy_h2d_async": [_P, _P, _I64, _P],
    "vml_make_labels": [_P, _P, _P, _I, _I, _I] + [_P] * 10 + [_P],
    "vml_sample_clips": [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "vml_content_unit": [_P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P, Cells, _P, _P, _P, _P, _P, _I, Dims, _I, _P],
    "vml_content_unit_supported": [Dims],
    "vml_content_out": [_P, _P, _P, _P, _P, _P, _P, _P, Cells, _P, Dims, _I, _P],
    "vml_boundary_unit": [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, Cells, _P, _P, _P, _P, _P, _P, _P, _I, Dims, _I, _P],
    "vml_boundary_pair_fused": [Dims, _I],
    "vml_boundary_unit_pair": [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, Cells, _P, _P, _P, _P, _P, _P, _I, Dims, _I, _P],
    "vml_colsum": [_P, _I64, _I64, _P, _I64, _I, _I, _I, _P, _I, C.c_float, _P],
    "vml_localize_bwd": [_P] * 12 + [Cells, _P, _P, _P, _P, _I, Dims, _P],
    "vml_pair_bwd": [_P, _I, _P, Cells, _P, _I, Dims, _P],
    "vml_cu_tail_bwd": [_P, _P, _I, Cells, _P, _P, Dims, _P],
    "vml_content_attn_bwd": [_P, _P, _P, _I, _I, _I, _I, _P, _I, _P, Cells, _P, _P, _P, _I, Dims, _P],
    "vml_gbar_bwd": [_P, _P, _P, _P, _P, _P, Cells, _P, _P, _P, _I, Dims, _P],
    "vml_softmax_bwd": [_P, _P, _P, _P, _I, _I, _I, C.c_float, _P],
    "vml_gate_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, Dims, _P],
    "vml_mask_rows": [_P, _P, _P, _I64, _I, _I, _P],
    "vml_span_pool_bwd": [_P, _P, _P, _P, _P, Cells, _P, _P, _I, Dims, _P],
    "vml_adam_step": [_P, _P, _P, _P, _I64, C.c_float, C.c_float, C.c_float, C.c_float, _I, C.c_float, _P],
    "vml_lstm_train_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "vml_lstm_train_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "vml_moment_pair": [_P, Cells, _P, Dims, _I, _P],
    "vml_moment_operand": [_P, _P, Cells, _P, Dims, _I, _P],
    "vml_moment_out": [_P, _P, _P, _P, Cells, _P, Dims, _I, _P],
    "vml_moment_gen_supported": [Dims, _I],
    "vml_moment_out_gen": [_P, _P, _P, _P, Cells, _P, _P, _I, Dims, _I, _P],
    "vml_localize": [_P, _P, _P, _P, Cells, _P, _P, _P, _P, _P, _I, Dims, _I, _P],
    "vml_scaled_iou_bce": [_P] * 13 + [_I, _I] + [_P] * 7 + [_P],
    "vml_score_topk_recall": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P],
    "vml_score_topk_recall_nm": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _I, _P, _I, _P],
}
_RET = {"vml_last_error": C.c_char_p, "vml_kernel_names": C.c_char_p, "vml_launch_count": C.c_int64}

_lib = None


def declared_symbols():
    """Entry points declared in include/vml_b200.h (used by the CPU export test)."""
    with open(HEADER) as f:
        return sorted(set(re.findall(r"\b(vml_[a-z0-9_]+)\s*\(", f.read())))


def load():
    """Load the CUDA library; raises (loudly) when it is missing -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VmlError(f"{LIB_PATH} not found: build it with `python -m vml_b200.build` "
                       "(__graft_entry__.build()).  vml_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RET.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        raise VmlError(f"{what} failed ({rc}): {load().vml_last_error().decode()}")


def kernel_names():
    return [k for k in load().vml_kernel_names().decode().split("\n") if k]


def launch_count() -> int:
    return int(load().vml_launch_count())


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


# bench instrumentation: when a list is installed here, every launcher call is also appended to it
# as (name, args) so that a stage's launches can be re-issued (e.g. captured into a CUDA graph).
_recorder = None


def set_recorder(rec):
    global _recorder
    _recorder = rec


def call(name: str, *args):
    if _recorder is not None:
        _recorder.append((name, args))
    check(getattr(load(), name)(*args), name)
