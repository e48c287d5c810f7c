"""CPU oracle for the SMIN proposal-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it may be
imported by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- as the checker or
as the timed CPU baseline -- and by nothing else.  The product path
(``video-moment-localization_b200``) never routes through it and raises when its
CUDA library is missing.

The oracle is a functional restatement (plain ``torch`` CPU ops over a flat
parameter dict) of the reference's ``models.py`` / ``utils.py`` /
``main.py`` loss, each function citing the reference file:line it follows.  It
is pinned against the reference itself: ``tools/make_golden.py`` imports the
unmodified reference from ``/root/reference`` in the build container, loads the
same deterministic parameters into it, runs it, and commits the reference's
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the oracle
against those vectors on every CPU run.

Parity status:
  * model forward / metric / dataset label formulas: pinned by live execution of
    the reference (golden vectors).
  * loss: pinned against the reference ``main.bce_loss`` with its documented
    one-token fix (``reduction=None`` -> ``'none'``; the unpatched function raises).
  * temporal NMS: the reference has no implementation (``utils.py:14``:
    "NMS NOT IMPLEMENTED YET") -> **parity unpinned**; the oracle's NMS is our
    own definition and is bypassed by default.
"""
from .smin_oracle import (  # noqa: F401
    SminConfig, CONFIGS, init_params, smin_forward, clip_projection, query_encoder,
    span_pool, content_unit, boundary_unit, moment_unit, localization, content_matrix,
)
from .metrics_oracle import compute_ious, scaled_iou_bce, loss_fn, topk_lowest_index, nms_topk  # noqa: F401
