"""Throughput-oriented scoring of a stream of batches: ``SMIN.forward`` + ``compute_ious`` for a
whole split (the loop of main.py:168-189), with the host taken out of the steady state.

* ``slots`` passes are in flight at once; each owns a CUDA stream, a workspace and -- after its first
  use -- a captured CUDA graph of everything behind ``vml_ingest`` (forward + R@n,IoU=m evaluation).
  Slots run concurrently, so the small kernels of one pass fill the SMs another pass leaves idle.
* ``coalesce`` submitted batches are scored by ONE pass: every ``submit`` is an eager ingest launch
  (the only kernel that reads caller memory) that fills its share of the slot's operand buffers; the
  last one of a group replays the graph.  Samples are independent (batch-slice invariance is a GPU
  test), so results are identical to scoring each batch alone -- the passes are just better filled.
* Pinned host batches are copied H2D on a dedicated copy stream into a ring of staging areas, so the
  copy engine runs ahead of the compute slots.
* Hit counters are accumulated on the device with atomics; each submitted batch can have its own
  hits read back asynchronously (``readback``).
"""
from __future__ import annotations

from collections import deque
from typing import Dict, List, Optional

import torch

from . import lib as L_
from .evaluate import score_topk_recall
from .smin import SMIN, Workspace, smin_core, smin_ingest

INPUT_KEYS = ("video_features", "video_mask", "query_features", "query_mask", "length_mask", "moment_mask", "sm")
# compact batches: the masks and the IoU map are generated on the device (vml_make_labels, dataset.py:95-110,139-149) from
# the annotation scalars -- only the features, the word mask and three scalars per sample cross PCIe
COMPACT_KEYS = ("video_features", "query_features", "query_mask", "times", "duration", "nfeats")
# packed batches: compact, and ``video_features`` holds only each sample's first ``nfeats`` rows back to back ([rows, d0]) --
# the all-zero rows get_fixed_length_features pads with (dataset.py:69-73) do not cross PCIe.  The variable-size tensor is
# the last one of the blob, so every other offset is the same for every batch.
# The word vectors are packed the same way (the first qlen = sum(query_mask) rows of every sample, right after the clip rows).
PACKED_KEYS = ("query_mask", "times", "duration", "nfeats", "video_features", "query_features")


def _is_compact(batch) -> bool:
    return "times" in batch and "sm" not in batch


def _is_packed(batch) -> bool:
    return bool(batch.get("_packed", False))


def _keys_of(batch):
    return PACKED_KEYS if _is_packed(batch) else COMPACT_KEYS if _is_compact(batch) else INPUT_KEYS


def pack_video_rows(video_features: torch.Tensor, nfeats: torch.Tensor) -> torch.Tensor:
    """[B, T, d0] -> [sum_b min(nfeats[b], T), d0]: the rows that are not padding, back to back."""
    B, T, _ = video_features.shape
    nf = nfeats.to(torch.int64).clamp(0, T).tolist()
    return torch.cat([video_features[b, : nf[b]] for b in range(B)], dim=0) if B else video_features.reshape(0, video_features.shape[-1])


def pack_query_rows(query_features: torch.Tensor, query_mask: torch.Tensor) -> torch.Tensor:
    """[B, Nq, 300] -> [sum_b qlen[b], 300], qlen = sum(query_mask): the rows models.py:50-54 reads (the sequence is packed to its length)."""
    B, Nq, _ = query_features.shape
    ql = query_mask.reshape(B, Nq).ne(0).sum(1).tolist()
    return torch.cat([query_features[b, : ql[b]] for b in range(B)], dim=0) if B else query_features.reshape(0, query_features.shape[-1])


def unpack_query_rows(rows: torch.Tensor, query_mask: torch.Tensor, Nq: int) -> torch.Tensor:
    """Inverse of ``pack_query_rows``; the rows past a query's length are zero (the reference never reads them)."""
    B = query_mask.shape[0]
    ql = query_mask.reshape(B, Nq).ne(0).sum(1).to(torch.int64)
    total = int(ql.sum().item())
    out = rows.new_zeros(B, Nq, rows.shape[1])
    b_idx = torch.repeat_interleave(torch.arange(B, device=rows.device), ql.to(rows.device))
    first = (torch.cumsum(ql, 0) - ql).to(rows.device)
    w_idx = torch.arange(total, device=rows.device) - first[b_idx]
    out[b_idx, w_idx] = rows[:total]
    return out


def unpack_video_rows(rows: torch.Tensor, nfeats: torch.Tensor, T: int) -> torch.Tensor:
    """Inverse of ``pack_video_rows`` (padding rows are zero, as dataset.py:69-73 makes them)."""
    nf = nfeats.to(torch.int64).clamp(0, T)
    B, total = nf.numel(), int(nf.sum().item())
    out = rows.new_zeros(B, T, rows.shape[1])
    b_idx = torch.repeat_interleave(torch.arange(B, device=rows.device), nf.to(rows.device))
    first = (torch.cumsum(nf, 0) - nf).to(rows.device)
    t_idx = torch.arange(total, device=rows.device) - first[b_idx]
    out[b_idx, t_idx] = rows[:total]
    return out


def pack_host_batch(batch: Dict[str, torch.Tensor], feature_dtype: Optional[torch.dtype] = None,
                    compact: bool = False, packed: bool = False) -> Dict[str, torch.Tensor]:
    """Re-lay one batch (dict with INPUT_KEYS) as views into ONE pinned host blob (key ``"_blob"``), so that
    ``ScoringPipeline.submit(..., from_host=True)`` moves it with a single H2D copy.  (A collate function
    can write straight into such a blob; the layout is the INPUT_KEYS order, each tensor 256-byte aligned.)
    ``feature_dtype=torch.bfloat16`` stores the clip features and word vectors as bf16 (half the bytes over PCIe;
    in bf16 precision the scores are bit-identical, the rounding just happens before the copy instead of after).
    ``compact=True`` keeps only COMPACT_KEYS (``times`` [B,2] / ``duration`` [B] as float64, ``nfeats`` [B] int64, as
    ``synth.make_batch`` and the dataset's annotations provide them): ``ScoringPipeline`` then builds the video / length /
    moment masks and the IoU map ``sm`` on the device, bit-identical to the host-built ones (GPU test).
    ``packed=True`` (implies compact): the clip features travel without their all-zero padding rows (``PACKED_KEYS``; the
    blob is as long as the batch's videos are) and ``vml_ingest_packed`` re-creates the padding on the device --
    bit-identical operands, ``mean(nfeats) / T`` of the bytes."""
    compact = compact or packed
    keys = PACKED_KEYS if packed else COMPACT_KEYS if compact else INPUT_KEYS
    B_, T_ = batch["video_features"].shape[0], batch["video_features"].shape[1]
    if compact:
        batch = {"video_features": batch["video_features"], "query_features": batch["query_features"],
                 "query_mask": batch["query_mask"], "times": batch["times"].to(torch.float64),
                 "duration": batch["duration"].to(torch.float64), "nfeats": batch["nfeats"].to(torch.int64)}
    if feature_dtype is not None:
        batch = dict(batch)
        for k in ("video_features", "query_features"):
            batch[k] = batch[k].to(feature_dtype)
    if packed:
        batch = dict(batch)
        d0_ = batch["video_features"].shape[-1]
        q_shape_ = tuple(batch["query_features"].shape)
        batch["video_features"] = pack_video_rows(batch["video_features"], batch["nfeats"])
        batch["query_features"] = pack_query_rows(batch["query_features"], batch["query_mask"])
    offs, total = {}, 0
    for k in keys:
        offs[k] = total
        total += (batch[k].numel() * batch[k].element_size() + 255) // 256 * 256
    blob = torch.empty(total, dtype=torch.uint8).pin_memory()
    out = {"_blob": blob}
    if packed:      # what a staging area must hold for ANY batch of this shape: every sample at full length
        out["_packed"] = True
        out["_rows_max"] = B_ * T_
        out["_q_shape"] = q_shape_
        es_ = batch["video_features"].element_size()
        out["_full_bytes"] = (offs["video_features"] + (B_ * T_ * d0_ * es_ + 255) // 256 * 256 +
                              (q_shape_[0] * q_shape_[1] * q_shape_[2] * es_ + 255) // 256 * 256)
    for k in keys:
        t = batch[k].contiguous()
        nbytes = t.numel() * t.element_size()
        view = blob[offs[k]: offs[k] + nbytes].view(t.dtype).view(t.shape)
        view.copy_(t)
        out[k] = view
    return out


def _blob_views(blob: torch.Tensor, like: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out, total = {}, 0
    for k in _keys_of(like):
        t = like[k]
        shape = tuple(t.shape)
        if _is_packed(like):      # the two packed tensors: the views span what a full-length batch would need; the word rows'
            # real position depends on the batch (right after its clip rows) and is derived by vml_ingest_packed itself
            shape = (like["_rows_max"], t.shape[1]) if k == "video_features" else like["_q_shape"] if k == "query_features" else shape
        nbytes = t.element_size()
        for n in shape:
            nbytes *= n
        out[k] = blob[total: total + nbytes].view(t.dtype).view(shape)
        total += (nbytes + 255) // 256 * 256
    return out


class Ticket:
    """Handle of one submitted batch; ``event`` is recorded once its pass has been enqueued."""

    def __init__(self, slot, index):
        self.slot, self.index, self.event = slot, index, None

    def synchronize(self):
        if self.event is None:
            raise RuntimeError("the batch's group has not been launched yet: call ScoringPipeline.flush()")
        self.event.synchronize()


class _Slot:
    def __init__(self, device, coalesce):
        self.stream = torch.cuda.Stream(device=device)
        self.ws = Workspace(device)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.outputs = None
        self.warm = 0
        self.fill = 0                      # batches ingested into the current group
        self.inp = None
        self.tickets: List[Ticket] = []
        self.readbacks: List[Optional[torch.Tensor]] = []
        self.step_counts = torch.zeros(coalesce, 2, 4, device=device, dtype=torch.int64)   # hits per batch of the group


class _Staging:
    """One H2D landing area: filled on the copy stream, consumed by a slot's ingest launch."""

    def __init__(self):
        self.buf: Optional[Dict[str, torch.Tensor]] = None
        self.blob: Optional[torch.Tensor] = None
        self.ready = torch.cuda.Event()
        self.free = torch.cuda.Event()         # re-recorded by every consumer
        self.used = False
        self.src_ptrs = None                   # device pointers of the 7 views, in vml_ingest argument order
        self.labels: Optional[Dict[str, torch.Tensor]] = None     # compact batches: device-built masks + sm of this area
        self.label_args = None
        self.packed = False                    # the area receives packed batches (video rows without padding, last in the blob)


# canonical dtypes of a collated batch (dataset.py:165-176); anything else takes the generic (converting) path
_CANON = {"video_features": torch.float32, "video_mask": torch.uint8, "query_features": torch.float32,
          "query_mask": torch.uint8, "length_mask": torch.bool, "moment_mask": torch.bool, "sm": torch.float32}
# vml_ingest takes (vf, qf, vmask, qmask, lmask, mmask, sm)
_INGEST_ORDER = ("video_features", "query_features", "video_mask", "query_mask", "length_mask", "moment_mask", "sm")


class _Plan:
    """The recorded ingest launch of one (slot, position in the group)."""

    def __init__(self, inp: dict, stream_ptr: int):
        a = inp["_ingest_args"]
        self.tail = tuple(a[inp["_ingest_nsrc"]:-1]) + (stream_ptr,)
        self.fn_name = inp["_ingest_fn"]
        self.tag = None                        # _canonical() of the batch the launch was recorded for
        self.fn = getattr(L_.load(), self.fn_name)
        self.inp = inp
        self.ev = torch.cuda.Event()


def _canonical(batch: Dict[str, torch.Tensor]):
    """None if the batch needs the generic (converting) path, else the ingest entry point its dtypes select."""
    f16 = batch["video_features"].dtype is torch.bfloat16
    for k, dt in _CANON.items():
        if k not in batch:                      # compact batch: the masks / sm are built on the device in canonical form
            continue
        t = batch[k]
        want = torch.bfloat16 if (f16 and k in ("video_features", "query_features")) else dt
        if t.dtype is not want or not t.is_contiguous():
            return None
    if _is_packed(batch):
        if batch["nfeats"].dtype is not torch.int64 or batch["video_features"].dim() != 2:
            return None
        return "vml_ingest_packed:bf16" if f16 else "vml_ingest_packed:f32"
    return "vml_ingest_bf16" if f16 else "vml_ingest"


class ScoringPipeline:
    def __init__(self, model: SMIN, slots: int = 3, coalesce: int = 1, use_graph: bool = True, nms_threshold: float = 1.0,
                 split_content: bool = False, timing_events: bool = False):
        L_.load()
        self.split_content = split_content
        self.timing_events = timing_events     # pass-completion events carry timestamps (bench: steady-state window)
        self.model = model
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise L_.VmlError("ScoringPipeline runs on CUDA (sm_100a) only; there is no CPU path")
        self.prec = L_.PREC[model.precision]
        self.dims = model._dims
        self.coalesce = max(1, coalesce)
        self.slots: List[_Slot] = [_Slot(self.device, self.coalesce) for _ in range(max(1, slots))]
        self.use_graph = use_graph
        self.nms_threshold = nms_threshold
        self.counts = torch.zeros(2, 4, device=self.device, dtype=torch.int64)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.staging = [_Staging() for _ in range((len(self.slots) + 1) * self.coalesce)]
        self._next_staging = 0
        self.num_samples = 0
        self._cur = 0
        self._batch = None
        self._pk = None
        # per-step fast path: the ingest launch of (slot, position) re-issued with new source pointers -- ~20 us of host
        # work per submitted batch instead of ~180 us through the generic module path (measured; the device step is ~210 us)
        self._plans: Dict[tuple, "_Plan"] = {}
        self._inflight = deque()               # (event, pinned host blob) of H2D copies that may still be running
        self._ingest_fn = getattr(L_.load(), "vml_ingest")
        self._h2d_fn = getattr(L_.load(), "vml_copy_h2d_async")
        self._labels_fn = getattr(L_.load(), "vml_make_labels")

    def _attach_label_buffers(self, stg: "_Staging", B: int):
        """Device outputs of ``vml_make_labels`` for a staging area that receives compact batches; afterwards ``stg.buf``
        offers all INPUT_KEYS, so everything downstream (ingest launch, fast path) is unchanged."""
        T, Lm, dev = self.dims.T, self.dims.L, self.device
        lab = {"sm": torch.empty(B, Lm, Lm, device=dev, dtype=torch.float32),
               "length_mask": torch.empty(B, Lm, device=dev, dtype=torch.uint8).view(torch.bool),
               "moment_mask": torch.empty(B, Lm, Lm, device=dev, dtype=torch.uint8).view(torch.bool),
               "video_mask": torch.empty(B, T, 1, device=dev, dtype=torch.uint8)}
        stg.labels = lab
        stg.buf.update(lab)
        p = lambda k: lab[k].data_ptr()
        # (times, duration, nfeats, B, T, L, sm, ym, ss, ys, se, ye, ya, length_mask, moment_mask, video_mask, stream)
        stg.label_args = (stg.buf["times"].data_ptr(), stg.buf["duration"].data_ptr(), stg.buf["nfeats"].data_ptr(), B, T, Lm,
                          p("sm"), None, None, None, None, None, None, p("length_mask"), p("moment_mask"), p("video_mask"))

    # -- one pass on a slot ------------------------------------------------------------------------------
    def _core_and_eval(self, slot: _Slot, pk, inp, group):
        out = smin_core(pk, self.dims, self.prec, slot.ws, inp, split_content=self.split_content)
        slot.step_counts.zero_()
        top = score_topk_recall(out[0], out[1], out[2], inp["mmask"], inp["sm"], 5, self.nms_threshold, self.counts,
                                step_counts=slot.step_counts, step_group=group)
        return out, top

    def _launch(self, slot: _Slot):
        B = self._batch
        with torch.no_grad(), torch.cuda.stream(slot.stream):
            pk = self._pk
            full = slot.fill == self.coalesce
            inp = dict(slot.inp)
            inp["B"] = slot.fill * B
            if self.use_graph and full and slot.graph is not None:
                slot.graph.replay()
            elif self.use_graph and full and slot.warm >= 1:
                # second full pass of the slot: every lazily-initialised piece has run once -> capture
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=slot.stream):
                    slot.outputs = self._core_and_eval(slot, pk, inp, B)
                slot.graph = g
                g.replay()
            else:
                slot.outputs = self._core_and_eval(slot, pk, inp, B)
                slot.warm += 1 if full else 0
            for i, rb in enumerate(slot.readbacks):
                if rb is not None:
                    rb.copy_(slot.step_counts[i], non_blocking=True)
            done = torch.cuda.Event(enable_timing=self.timing_events)
            done.record(slot.stream)
        for t in slot.tickets:
            t.event = done
        slot.fill, slot.tickets, slot.readbacks = 0, [], []

    def submit(self, batch: Dict[str, torch.Tensor], from_host: bool = False, readback: Optional[torch.Tensor] = None) -> Ticket:
        """Enqueue one batch (dict with INPUT_KEYS).  ``from_host``: the tensors are pinned host memory
        (H2D on the copy stream).  ``readback``: pinned int64 [2,4] host tensor that receives THIS
        batch's hit counters (async D2H after its pass).  Returns a Ticket; after
        ``ticket.synchronize()``, ``ticket.slot.outputs`` holds the pass's (pm, ps, pe, pa) and top-k
        records (rows ``ticket.index * B ...`` belong to this batch) until the slot is reused."""
        B = batch["query_mask"].shape[0]
        if _is_packed(batch) and not from_host:       # device-resident packed batch: nothing to save, re-pad it
            batch = {**{k: batch[k] for k in COMPACT_KEYS},
                     "video_features": unpack_video_rows(batch["video_features"], batch["nfeats"], self.dims.T),
                     "query_features": unpack_query_rows(batch["query_features"], batch["query_mask"], self.dims.Nq)}
        if _is_compact(batch) and not from_host:      # device-resident compact batch: build the masks + sm, then as usual
            from .labels import make_labels
            lab = make_labels(batch["times"], batch["duration"], batch["nfeats"], self.dims.T, self.dims.L)
            batch = {**{k: batch[k] for k in ("video_features", "query_features", "query_mask")},
                     **{k: lab[k] for k in ("video_mask", "length_mask", "moment_mask", "sm")}}
        if self._batch is None:
            self._batch = B
        if B != self._batch:
            return self._submit_ragged(batch, from_host, readback)
        slot = self.slots[self._cur]
        if slot.fill == 0 or self._pk is None:       # parameters are looked at once per pass, not per batch
            with torch.no_grad():
                pk = self.model._weights(self.device, self.prec)
                if pk is not self._pk:               # (re)packed on the caller's stream: publish to every slot stream
                    torch.cuda.synchronize(self.device)
                    self._pk = pk
                    self.invalidate()
        plan = self._plans.get((self._cur, slot.fill))
        fast = plan is not None and _canonical(batch) == plan.tag
        stg = None
        if from_host:
            stg = self.staging[self._next_staging]
            self._next_staging = (self._next_staging + 1) % len(self.staging)
            blob = batch.get("_blob")
            packed = _is_packed(batch)
            need = batch["_full_bytes"] if packed else (blob.numel() if blob is not None else 0)
            if packed and blob is None:
                raise L_.VmlError("packed batches travel as one pinned blob: make them with pack_host_batch(..., packed=True)")
            if stg.buf is None or stg.packed != packed or (blob is not None and (stg.blob is None or stg.blob.numel() != need)):
                if stg.buf is not None:                       # layout change mid-run (rare): let the old area drain first
                    torch.cuda.synchronize(self.device)
                stg.packed, stg.label_args, stg.labels = packed, None, None
                if blob is not None:                          # one device blob mirroring the host blob: one copy per batch
                    stg.blob = torch.empty(need, dtype=torch.uint8, device=self.device)
                    stg.buf = _blob_views(stg.blob, batch)
                else:
                    stg.buf = {k: torch.empty(batch[k].shape, dtype=batch[k].dtype, device=self.device) for k in _keys_of(batch)}
                if _is_compact(batch):
                    self._attach_label_buffers(stg, B)
                stg.src_ptrs = tuple(stg.buf[k].data_ptr() for k in _INGEST_ORDER + (("nfeats",) if packed else ()))
            if stg.used:
                self.copy_stream.wait_event(stg.free)         # the previous consumer's ingest has read it
            if blob is not None and stg.blob is not None:
                # raw cudaMemcpyAsync: PyTorch's pinned-memory allocator does not see it, so the blob is kept alive
                # (self._inflight) until its copy is known to have landed -- checked with event queries, never by blocking
                L_.check(self._h2d_fn(stg.blob.data_ptr(), blob.data_ptr(), blob.numel(), self.copy_stream.cuda_stream),
                         "vml_copy_h2d_async")
                stg.ready = torch.cuda.Event()
                stg.ready.record(self.copy_stream)
                self._inflight.append((stg.ready, blob))
                while self._inflight and self._inflight[0][0].query():
                    self._inflight.popleft()
            else:
                with torch.cuda.stream(self.copy_stream):
                    for k in _keys_of(batch):
                        stg.buf[k].copy_(batch[k], non_blocking=True)
            if blob is None or stg.blob is None:
                stg.ready.record(self.copy_stream)
            slot.stream.wait_event(stg.ready)
            if stg.label_args is not None:                    # compact batch: masks + sm built on the device, on the slot's stream
                L_.check(self._labels_fn(*stg.label_args, slot.stream.cuda_stream), "vml_make_labels")
            src = stg.buf
        else:
            src = batch
        if fast:
            if not from_host:
                plan.ev.record(torch.cuda.current_stream(self.device))
                slot.stream.wait_event(plan.ev)               # the caller's tensors are ready
                ptrs = tuple(batch[k].data_ptr() for k in _INGEST_ORDER)
                for k in _INGEST_ORDER:                       # read on the slot's stream: no reuse of the memory before that
                    batch[k].record_stream(slot.stream)
            else:
                ptrs = stg.src_ptrs
            L_.check(plan.fn(*ptrs, *plan.tail), plan.fn_name)
            slot.inp = plan.inp
        else:
            with torch.no_grad():
                if not from_host:
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(self.device))
                    slot.stream.wait_event(ev)               # the caller's tensors are ready
                with torch.cuda.stream(slot.stream):
                    slot.inp = smin_ingest(self.dims, self.prec, slot.ws, *[src[k] for k in INPUT_KEYS], static=True,
                                           b_off=slot.fill * B, b_total=self.coalesce * B,
                                           nfeats=src["nfeats"] if (from_host and stg.packed) else None,
                                           q_packed=bool(from_host and stg.packed))
                if not from_host:
                    for k in INPUT_KEYS:
                        src[k].record_stream(slot.stream)
            tag = _canonical(batch if from_host else src)
            if tag is not None:
                self._plans[(self._cur, slot.fill)] = plan = _Plan(slot.inp, slot.stream.cuda_stream)
                plan.tag = tag
        if from_host:
            stg.free.record(slot.stream)
            stg.used = True
        ticket = Ticket(slot, slot.fill)
        slot.tickets.append(ticket)
        slot.readbacks.append(readback)
        slot.fill += 1
        self.num_samples += B
        if slot.fill == self.coalesce:
            self._launch(slot)
            self._cur = (self._cur + 1) % len(self.slots)
        return ticket

    def _submit_ragged(self, batch, from_host, readback) -> Ticket:
        """A batch of another size (the ragged tail of a split, main.py:168-189 with len(dataset) % batch != 0): the
        pending group is launched and the batch is scored eagerly through the module on the caller's stream, its hits
        added to the same device counters.  Once per split, so no graph and no staging ring."""
        self.flush()
        caller = torch.cuda.current_stream(self.device)
        for s in self.slots:                      # the counters are shared: order after everything enqueued so far
            ev = torch.cuda.Event()
            ev.record(s.stream)
            caller.wait_event(ev)
        with torch.no_grad():
            dev = {k: (batch[k].to(self.device, non_blocking=True) if not batch[k].is_cuda else batch[k]) for k in _keys_of(batch)}
            if _is_packed(batch):
                dev["video_features"] = unpack_video_rows(dev["video_features"], dev["nfeats"], self.dims.T)
                dev["query_features"] = unpack_query_rows(dev["query_features"], dev["query_mask"], self.dims.Nq)
            if _is_compact(batch):
                from .labels import make_labels
                dev.update(make_labels(dev["times"], dev["duration"], dev["nfeats"], self.dims.T, self.dims.L))
            from .synth import MODEL_INPUT_KEYS
            out = self.model(*[dev[k] for k in MODEL_INPUT_KEYS], split_content=self.split_content)
            step = torch.zeros(1, 2, 4, device=self.device, dtype=torch.int64)
            top = score_topk_recall(out[0], out[1], out[2], dev["moment_mask"], dev["sm"], 5, self.nms_threshold, self.counts,
                                    step_counts=step, step_group=out[0].shape[0])
            if readback is not None:
                readback.copy_(step[0], non_blocking=True)
        done = torch.cuda.Event()
        done.record(caller)
        for s in self.slots:                      # later passes must not overtake the eager kernels' counter updates
            s.stream.wait_event(done)
        self.num_samples += out[0].shape[0]

        class _Eager:                             # what Ticket.slot exposes: the outputs of this batch alone
            outputs = (out, top)
        t = Ticket(_Eager, 0)
        t.event = done
        return t

    def flush(self):
        """Launch a partially filled group (end of the split)."""
        slot = self.slots[self._cur]
        if slot.fill:
            self._launch(slot)
            self._cur = (self._cur + 1) % len(self.slots)

    def invalidate(self):
        """Drop the captured graphs (call after the model's parameters changed)."""
        for s in self.slots:
            s.graph, s.warm = None, 0

    def wait_all(self, stream=None):
        """Make ``stream`` (default: the caller's current stream) wait for every enqueued pass."""
        self.flush()
        stream = stream or torch.cuda.current_stream(self.device)
        for s in self.slots:
            ev = torch.cuda.Event()
            ev.record(s.stream)
            stream.wait_event(ev)

    def synchronize(self):
        self.flush()
        for s in self.slots:
            s.stream.synchronize()
        self._inflight.clear()                 # every copy has landed: the host blobs may go

    def result(self, normalize: bool = True):
        """Recall table like main.py:163,189,209 (one D2H read)."""
        self.synchronize()
        host = self.counts.cpu()
        den = float(self.num_samples) if normalize and self.num_samples else 1.0
        return {f"R@{n_}, IoU={m_}": float(host[a, t].item()) / den
                for a, n_ in enumerate((1, 5)) for t, m_ in enumerate((0.1, 0.3, 0.5, 0.7))}
