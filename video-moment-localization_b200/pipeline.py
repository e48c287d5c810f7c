"""Throughput-oriented scoring of a stream of batches: ``SMIN.forward`` + ``compute_ious`` for a
whole split (the loop of main.py:168-189), with the host taken out of the steady state.

Each of ``slots`` in-flight batches owns a CUDA stream, a workspace and -- after its first use --
a captured CUDA graph of everything behind ``vml_ingest`` (forward + R@n,IoU=m evaluation).  A
step is then two host calls: the eager ingest launch (the only kernel that reads caller memory)
and one graph replay.  Slots run concurrently, so the small kernels of one batch fill the SMs
another batch leaves idle.  Hit counters are accumulated on the device with atomics and read
back once (or per step, asynchronously, if the caller wants them).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import lib as L_
from .evaluate import score_topk_recall
from .smin import SMIN, Workspace, smin_core, smin_ingest

INPUT_KEYS = ("video_features", "video_mask", "query_features", "query_mask", "length_mask", "moment_mask", "sm")


class _Slot:
    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.ws = Workspace(device)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.done = torch.cuda.Event()
        self.outputs = None
        self.warm = 0
        self.step_counts = torch.zeros(2, 4, device=device, dtype=torch.int64)   # hits of the slot's current step


class _Staging:
    """One H2D landing area: filled on the copy stream, consumed by a slot's ingest launch."""

    def __init__(self):
        self.buf: Optional[Dict[str, torch.Tensor]] = None
        self.ready = torch.cuda.Event()
        self.free: Optional[torch.cuda.Event] = None


class ScoringPipeline:
    def __init__(self, model: SMIN, slots: int = 2, use_graph: bool = True, nms_threshold: float = 1.0):
        L_.load()
        self.model = model
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise L_.VmlError("ScoringPipeline runs on CUDA (sm_100a) only; there is no CPU path")
        self.prec = L_.PREC[model.precision]
        self.dims = model._dims
        self.slots: List[_Slot] = [_Slot(self.device) for _ in range(max(1, slots))]
        self.use_graph = use_graph
        self.nms_threshold = nms_threshold
        self.counts = torch.zeros(2, 4, device=self.device, dtype=torch.int64)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.staging = [_Staging() for _ in range(len(self.slots) + 1)]
        self._next_staging = 0
        self.num_samples = 0
        self._next = 0
        self._batch = None
        self._pk = None

    # -- one step on a slot ---------------------------------------------------------------------------
    def _core_and_eval(self, slot: _Slot, pk, inp):
        out = smin_core(pk, self.dims, self.prec, slot.ws, inp)
        slot.step_counts.zero_()
        top = score_topk_recall(out[0], out[1], out[2], inp["mmask"], inp["sm"], 5, self.nms_threshold, self.counts,
                                step_counts=slot.step_counts)
        return out, top

    def submit(self, batch: Dict[str, torch.Tensor], from_host: bool = False, readback: Optional[torch.Tensor] = None):
        """Enqueue one batch (dict with INPUT_KEYS).  ``from_host``: the tensors are pinned host
        memory; they are copied H2D on the pipeline's copy stream into a ring of staging areas, so the
        copy engine runs ahead of the compute slots.  ``readback``: pinned int64 [2,4] host tensor that
        receives THIS step's hit counters (async D2H on the slot's stream).  Returns
        (event, slot): the event is recorded after the step; ``slot.outputs`` holds (pm, ps, pe, pa),
        (top_idx, top_score, top_iou, counts), valid until the slot is reused."""
        B = batch["video_features"].shape[0]
        if self._batch is None:
            self._batch = B
        slot = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        caller = torch.cuda.current_stream(self.device)
        with torch.no_grad():
            pk = self.model._weights(self.device, self.prec)
            if pk is not self._pk:                   # (re)packed on the caller's stream: publish to every slot stream
                torch.cuda.synchronize(self.device)
                self._pk = pk
                self.invalidate()
        with torch.no_grad(), torch.cuda.stream(slot.stream):
            if from_host:
                stg = self.staging[self._next_staging]
                self._next_staging = (self._next_staging + 1) % len(self.staging)
                if stg.buf is None or stg.buf["video_features"].shape[0] != B:
                    stg.buf = {k: torch.empty(batch[k].shape, dtype=batch[k].dtype, device=self.device) for k in INPUT_KEYS}
                with torch.cuda.stream(self.copy_stream):
                    if stg.free is not None:
                        self.copy_stream.wait_event(stg.free)     # the previous consumer's ingest has read it
                    for k in INPUT_KEYS:
                        stg.buf[k].copy_(batch[k], non_blocking=True)
                    stg.ready.record(self.copy_stream)
                slot.stream.wait_event(stg.ready)
                src = stg.buf
            else:
                ev = torch.cuda.Event()
                ev.record(caller)
                slot.stream.wait_event(ev)           # the caller's tensors are ready
                src = batch
            inp = smin_ingest(self.dims, self.prec, slot.ws, *[src[k] for k in INPUT_KEYS], static=True)
            if from_host:
                stg.free = torch.cuda.Event()
                stg.free.record(slot.stream)
            graphable = self.use_graph and B == self._batch
            if graphable and slot.graph is not None:
                slot.graph.replay()
            elif graphable and slot.warm >= 1:
                # second use of the slot: every lazily-initialised piece has run once -> capture
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=slot.stream):
                    slot.outputs = self._core_and_eval(slot, pk, inp)
                slot.graph = g
                g.replay()
            else:
                slot.outputs = self._core_and_eval(slot, pk, inp)
                slot.warm += 1
            if readback is not None:
                readback.copy_(slot.step_counts, non_blocking=True)
            done = torch.cuda.Event()
            done.record(slot.stream)
            slot.done = done
        self.num_samples += B
        return done, slot

    def wait_all(self, stream=None):
        """Make ``stream`` (default: the caller's current stream) wait for every enqueued step."""
        stream = stream or torch.cuda.current_stream(self.device)
        for s in self.slots:
            ev = torch.cuda.Event()
            ev.record(s.stream)
            stream.wait_event(ev)

    def invalidate(self):
        """Drop the captured graphs (call after the model's parameters changed)."""
        for s in self.slots:
            s.graph, s.warm = None, 0

    def synchronize(self):
        for s in self.slots:
            s.stream.synchronize()

    def result(self, normalize: bool = True):
        """Recall table like main.py:163,189,209 (one D2H read)."""
        self.synchronize()
        host = self.counts.cpu()
        den = float(self.num_samples) if normalize and self.num_samples else 1.0
        return {f"R@{n_}, IoU={m_}": float(host[a, t].item()) / den
                for a, n_ in enumerate((1, 5)) for t, m_ in enumerate((0.1, 0.3, 0.5, 0.7))}
