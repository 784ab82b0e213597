"""Operator layer: each function is one C-ABI entry point on torch CUDA tensors.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in
``csrc/libtspn_b200.so``.  Every function enqueues on the current CUDA stream and returns
without synchronising.
"""
from __future__ import annotations

import os

from typing import Dict, Optional

import torch

from . import _lib
from ._lib import check, load, ptr, stream_ptr
from .batch import DeviceBatch

PREC = {"fp32": _lib.PREC_FP32_EXACT, "exact": _lib.PREC_FP32_EXACT, "fp32_exact": _lib.PREC_FP32_EXACT,
        "tensor": _lib.PREC_TENSOR, "bf16": _lib.PREC_TENSOR, "tf32": _lib.PREC_TENSOR}


# kernels launched by this process through the C ABI (bench.py reports it as gpu_launches)
_LAUNCHES = 0


def launch_count() -> int:
    return _LAUNCHES


def _count(n: int) -> None:
    global _LAUNCHES
    _LAUNCHES += n


def count_launches(n: int) -> None:
    """Account for kernels launched by a CUDA-graph replay (pipeline.GraphedStage)."""
    _count(n)


def _cuda(t: torch.Tensor, dtype=None) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("tspn_b200 ops need CUDA tensors (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError("expected %s, got %s" % (dtype, t.dtype))
    return t.contiguous()


def require_device() -> None:
    """Raise unless the current device can run the library (sm_100)."""
    if not torch.cuda.is_available():
        raise RuntimeError("tspn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    check(load().tspn_check_device(), "tspn_check_device")


# ---------------------------------------------------------------------------------------------
def enumerate_pairs(batch: DeviceBatch) -> torch.Tensor:
    """``[sum P, 2]`` int64 (s, o) — the h5 ``pairs`` table (vrdataset.py:208)."""
    out = torch.empty((batch.total_pairs, 2), dtype=torch.int64, device=batch.device)
    check(load().tspn_enumerate_pairs(ptr(batch.table), batch.num_videos, batch.total_pairs, ptr(out),
                                      stream_ptr()), "tspn_enumerate_pairs")
    _count(1)
    return out


def geo_window_offsets(batch: DeviceBatch, geo_off: Optional[torch.Tensor] = None,
                       total: Optional[torch.Tensor] = None):
    """Row offsets of the WINDOWED geometry layout (``tspn_geo_window_offsets``): ``geo_off`` int64 ``[sum P]`` (floats)
    and ``total`` int64 ``[1]``, from the tracklet spans of the batch as it is in HBM now; current stream."""
    if geo_off is None:
        geo_off = torch.zeros(max(batch.total_pairs, 1), dtype=torch.int64, device=batch.device)
        total = torch.zeros(1, dtype=torch.int64, device=batch.device)
    check(load().tspn_geo_window_offsets(ptr(batch.table), batch.num_videos, batch.total_pairs, ptr(batch.span),
                                         ptr(geo_off), ptr(total), stream_ptr()), "tspn_geo_window_offsets")
    _count(1)
    return geo_off, total


def pair_geometry_outputs(batch: DeviceBatch, write_geo: bool = True, windowed: bool = False) -> Dict[str, torch.Tensor]:
    """Caller-owned outputs and workspace of ``tspn_pair_geo_viou`` for this batch.  ``windowed``: the opt-in
    WINDOWED layout of the rows (``tspn_pair_geo_viou_windowed``: per pair only its overlap window's frames, 7
    channels) - ``out["geo_off"]`` are the batch's row offsets, which ``DeviceBatch.enable_windows`` keeps current
    across refills; the buffer holds 7/8 of the dense floats, the layout's upper bound."""
    dev, tot, p = batch.device, batch.totals, batch.total_pairs
    out = {}
    n_geo = int(tot[_lib.TOT_GEO_FLOATS])
    if windowed and write_geo:
        out["geo_off"], out["geo_total"] = batch.enable_windows()
        n_geo = n_geo // _lib.GEO_CHANNELS * (_lib.GEO_CHANNELS - 1)
    out["geo"] = torch.empty(n_geo, dtype=torch.float32, device=dev) if write_geo else None
    out["viou"] = torch.empty(p, dtype=torch.float32, device=dev)
    out["tiou"] = torch.empty(p, dtype=torch.float32, device=dev)
    out["overlap"] = torch.empty((p, 2), dtype=torch.int32, device=dev)
    ws_bytes = load().tspn_pair_geo_workspace_bytes(batch.total_tracklets, batch.total_pairs,
                                                    int(tot[_lib.TOT_MAX_CHUNKS]))
    out["workspace"] = torch.empty(ws_bytes // 8, dtype=torch.float64, device=dev)
    return out


def pair_geometry_phase(batch: DeviceBatch, out: Dict[str, torch.Tensor], phase: int, clipped: bool = False,
                        dense_ctas: Optional[bool] = None, persistent: Optional[bool] = None,
                        reserve_sms: int = 0) -> None:
    """One phase of ``tspn_pair_geo_viou`` on the current stream: ``_lib.GEO_PHASE_PRE`` (per-tracklet volumes),
    ``GEO_PHASE_MAIN`` (the pair kernel: geometry rows, per-chunk fixed-point sums, overlap windows),
    ``GEO_PHASE_POST`` (vIoU / tIoU), or 0 for all three.  PRE may run on another stream concurrently with MAIN
    (every sum has a single writer, nothing is zeroed); POST needs both.
    ``reserve_sms``: SM slots the persistent pair kernel leaves to concurrent streams (TSPN_GEO_RESERVE_SHIFT)."""
    if dense_ctas is None:
        dense_ctas = os.environ.get("TSPN_GEO_DENSE", "0") == "1"
    tot = batch.totals
    flags = (_lib.VIOU_CLIPPED if clipped else _lib.VIOU_FULL) | (_lib.GEO_DENSE_CTAS if dense_ctas else 0)
    if persistent is None:
        persistent = os.environ.get("TSPN_GEO_PERSISTENT", "1") == "1"
    if persistent and not dense_ctas:
        flags |= _lib.GEO_PERSISTENT | (max(0, min(int(reserve_sms), 255)) << _lib.GEO_RESERVE_SHIFT)
    if out.get("geo_off") is not None:
        check(load().tspn_pair_geo_viou_windowed(
            ptr(batch.table), batch.num_videos, int(tot[_lib.TOT_ITEMS]), int(tot[_lib.TOT_GEO_CHUNK]),
            int(tot[_lib.TOT_MAX_CHUNKS]), batch.total_tracklets, batch.total_pairs,
            int(tot[_lib.TOT_BOXES]), ptr(batch.boxes), ptr(batch.span), ptr(out["geo"]), ptr(out["geo_off"]),
            ptr(out["viou"]), ptr(out["tiou"]), ptr(out["overlap"]), flags | phase, ptr(out["workspace"]),
            stream_ptr()), "tspn_pair_geo_viou_windowed")
    else:
        check(load().tspn_pair_geo_viou(
            ptr(batch.table), batch.num_videos, int(tot[_lib.TOT_ITEMS]), int(tot[_lib.TOT_GEO_CHUNK]),
            int(tot[_lib.TOT_MAX_CHUNKS]), batch.total_tracklets, batch.total_pairs,
            int(tot[_lib.TOT_BOXES]), ptr(batch.boxes), ptr(batch.span), ptr(out.get("geo")), ptr(out["viou"]),
            ptr(out["tiou"]), ptr(out["overlap"]), flags | phase, ptr(out["workspace"]), stream_ptr()),
            "tspn_pair_geo_viou")
    _count(3 if phase == 0 else 1)      # volumes, pair kernel (+ its queue reset), per-pair finalize


def pair_geometry(batch: DeviceBatch, write_geo: bool = True, clipped: bool = False,
                  out: Optional[Dict[str, torch.Tensor]] = None,
                  dense_ctas: Optional[bool] = None, events=None,
                  persistent: Optional[bool] = None, windowed: bool = False) -> Dict[str, torch.Tensor]:
    """All-pairs per-frame geometry + vIoU/tIoU/overlap (trajectory.py:85-141, common.py:65-106).
    ``events``: a pair of CUDA events recorded immediately before and after the pair kernel itself (the
    volume pre-kernel and the per-pair finalize are issued as separate phases around them).
    ``dense_ctas`` selects the 1024-threads-per-SM shape of the kernel (2-stage ring, 64 registers) instead of
    the default ~512 threads per SM (bit-identical results, measured slower; A/B timing only - default from
    the environment variable TSPN_GEO_DENSE).  ``persistent``: the pair kernel as persistent CTAs pulling work
    items from a queue (default, TSPN_GEO_PERSISTENT) or one CTA per work item; bit-identical."""
    if out is None:
        out = pair_geometry_outputs(batch, write_geo, windowed=windowed)
    if events is None:
        pair_geometry_phase(batch, out, 0, clipped, dense_ctas, persistent)
    else:
        stream = torch.cuda.current_stream(batch.device)
        pair_geometry_phase(batch, out, _lib.GEO_PHASE_PRE, clipped, dense_ctas, persistent)
        events[0].record(stream)
        pair_geometry_phase(batch, out, _lib.GEO_PHASE_MAIN, clipped, dense_ctas, persistent)
        events[1].record(stream)
        pair_geometry_phase(batch, out, _lib.GEO_PHASE_POST, clipped, dense_ctas, persistent)
    return out


def cubic_iou(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """``[n, t, 4] x [m, t, 4] -> [n, m]`` (trajectory.py:127-141)."""
    b1, b2 = _cuda(b1, torch.float32), _cuda(b2, torch.float32)
    if b1.dim() != 3 or b2.dim() != 3 or b1.shape[1] != b2.shape[1] or b1.shape[2] != 4 or b2.shape[2] != 4:
        raise ValueError("cubic_iou expects [n, t, 4] and [m, t, 4]")
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    check(load().tspn_cubic_iou(ptr(b1), b1.shape[0], ptr(b2), b2.shape[0], b1.shape[1], ptr(out), stream_ptr()),
          "tspn_cubic_iou")
    _count(1)
    return out


def viou_pairs(pool: torch.Tensor, traj_off: torch.Tensor, traj_span: torch.Tensor, a: torch.Tensor,
               b: torch.Tensor, clipped: bool = False) -> torch.Tensor:
    """vIoU of explicit trajectory pairs (evaluation/common.py:65-106; association.py:35-48)."""
    pool = _cuda(pool, torch.float32)
    out = torch.empty(a.shape[0], dtype=torch.float32, device=pool.device)
    check(load().tspn_viou_pairs(ptr(pool), ptr(_cuda(traj_off, torch.int64)), ptr(_cuda(traj_span, torch.int32)),
                                 ptr(_cuda(a, torch.int32)), ptr(_cuda(b, torch.int32)), a.shape[0],
                                 _lib.VIOU_CLIPPED if clipped else _lib.VIOU_FULL, ptr(out), stream_ptr()),
          "tspn_viou_pairs")
    _count(1)
    return out


def viou_pairs_f64(pool: torch.Tensor, traj_off: torch.Tensor, traj_span: torch.Tensor, a: torch.Tensor,
                   b: torch.Tensor, clipped: bool = False, traj_len: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp64 vIoU of explicit trajectory pairs with per-trajectory volumes summed once (the evaluation
    loop of lib/evaluation/visual_relation_detection.py:8-36; association.py:35-48 when ``clipped``).
    ``traj_len``: boxes per trajectory when a list is longer than its duration (volumes run over the list)."""
    pool = _cuda(pool, torch.float32)
    n_traj = int(traj_span.shape[0])
    out = torch.empty(a.shape[0], dtype=torch.float64, device=pool.device)
    ws = torch.empty(load().tspn_viou_pairs_workspace_bytes(n_traj), dtype=torch.uint8, device=pool.device)
    check(load().tspn_viou_pairs_f64(ptr(pool), ptr(_cuda(traj_off, torch.int64)), ptr(_cuda(traj_span, torch.int32)),
                                     ptr(_cuda(traj_len, torch.int32)) if traj_len is not None else None,
                                     n_traj, ptr(_cuda(a, torch.int32)), ptr(_cuda(b, torch.int32)), a.shape[0],
                                     _lib.VIOU_CLIPPED if clipped else _lib.VIOU_FULL, ptr(out), ptr(ws),
                                     stream_ptr()),
          "tspn_viou_pairs_f64")
    _count(1 if clipped else 2)
    return out


def normalize_motion(motion: torch.Tensor) -> torch.Tensor:
    """L1-normalise the four 1000-wide BoW blocks of every tracklet (vrdataset.py:227-236); ``motion`` is
    fp32 ``[n, 4000]`` or the compact u8 counts of ``HostBatch(compact=True)``."""
    if motion.dtype == torch.uint8:
        motion = _cuda(motion, torch.uint8)
        out = torch.empty(motion.shape, dtype=torch.float32, device=motion.device)
        check(load().tspn_normalize_motion_u8(ptr(motion), motion.shape[0], ptr(out), stream_ptr()),
              "tspn_normalize_motion_u8")
        _count(1)
        return out
    motion = _cuda(motion, torch.float32)
    out = torch.empty_like(motion)
    check(load().tspn_normalize_motion(ptr(motion), motion.shape[0], ptr(out), stream_ptr()),
          "tspn_normalize_motion")
    _count(1)
    return out


def feature_dim(n_classes: int) -> int:
    return 2 * n_classes + 2 * _lib.MOTION_DIM + _lib.REL_DIM


def padded(n: int, mult: int) -> int:
    return (n + mult - 1) // mult * mult


def assemble_features(batch: DeviceBatch, motion_norm: torch.Tensor, geo: torch.Tensor, overlap: torch.Tensor,
                      rows: Optional[torch.Tensor] = None, want_fp32: bool = True, want_bf16: bool = False,
                      out_fp32: Optional[torch.Tensor] = None, out_bf16: Optional[torch.Tensor] = None):
    """Feature rows ``[n_rows, F]`` in the layout of vrdataset.py:219-243.

    Returns ``(feat_fp32 | None, feat_bf16 | None)``; both are views ``[:, :F]`` of row-padded
    buffers (stride multiple of 4 / 8 elements) so they can be fed to TMA.
    """
    c = batch.cls.shape[1]
    f = feature_dim(c)
    n_rows = batch.total_pairs if rows is None else int(rows.shape[0])
    dev = batch.device
    ld32, ld16 = padded(f, 4), padded(f, 8)
    if want_fp32 and out_fp32 is None:
        out_fp32 = torch.empty((n_rows, ld32), dtype=torch.float32, device=dev)
    if want_bf16 and out_bf16 is None:
        out_bf16 = torch.empty((n_rows, ld16), dtype=torch.bfloat16, device=dev)
    check(load().tspn_assemble_features(
        ptr(batch.table), batch.num_videos, batch.total_pairs, int(batch.totals[_lib.TOT_MAX_T]), ptr(batch.cls), c,
        ptr(motion_norm), ptr(geo),
        ptr(overlap), ptr(rows), n_rows, ptr(out_fp32), out_fp32.stride(0) if out_fp32 is not None else 0,
        ptr(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0, stream_ptr()), "tspn_assemble_features")
    _count(1)
    return (out_fp32[:, :f] if out_fp32 is not None else None,
            out_bf16[:, :f] if out_bf16 is not None else None)


PPN_KEYS = ("sub_emb.0.weight", "sub_emb.0.bias", "sub_emb.2.weight", "sub_emb.2.bias",
            "obj_emb.0.weight", "obj_emb.0.bias", "obj_emb.2.weight", "obj_emb.2.bias")


def _check_ppn_shapes(w, c: int) -> int:
    """The relationness kernels contract the two embeddings over the class dimension: they need
    ``RELPN.PPN.OUT_CHANNELS == RELPN.PPN.IN_CHANNELS == C`` (the reference's defaults, 35/35 and 80/80) and
    raise otherwise instead of reading the weights out of bounds.  Returns the hidden width."""
    h = int(w[0].shape[0])
    want = {0: (h, c), 1: (h,), 2: (c, h), 3: (c,), 4: (h, c), 5: (h,), 6: (c, h), 7: (c,)}
    for i, shape in want.items():
        if tuple(w[i].shape) != shape:
            raise ValueError("PPNHead weight %s has shape %s, expected %s: the CUDA relationness path needs "
                             "RELPN.PPN.IN_CHANNELS == RELPN.PPN.OUT_CHANNELS == number of classes (%d)"
                             % (PPN_KEYS[i], tuple(w[i].shape), shape, c))
    return h


def relationness_tc_supported(batch: DeviceBatch, hidden: int = 64) -> bool:
    return bool(load().tspn_relationness_tc_supported(int(batch.totals[_lib.TOT_MAX_N]), int(batch.cls.shape[1]),
                                                      int(hidden)))


def relationness(batch: DeviceBatch, weights, cls: Optional[torch.Tensor] = None,
                 precision: str = "fp32") -> torch.Tensor:
    """PPNHead scores of every video, flat ``[sum N*N]`` (ppn.py:92-112).  ``weights`` is the
    8-tuple of ``PPN_KEYS`` tensors.  ``precision``: ``"fp32"`` (exact order, bit-reproducible) or ``"tensor"``
    (tcgen05, tf32 operands, fp32 accumulate: within 1e-2 of the float64 definition)."""
    cls = batch.cls if cls is None else cls
    cls = _cuda(cls, torch.float32)
    w = [_cuda(t, torch.float32) for t in weights]
    c = int(cls.shape[1])
    h = _check_ppn_shapes(w, c)
    dev = cls.device
    scores = torch.empty(batch.total(_lib.TOT_SCORES), dtype=torch.float32, device=dev)
    ws = torch.empty(load().tspn_relationness_workspace_bytes(batch.total_tracklets, c, h) // 4,
                     dtype=torch.float32, device=dev)
    check(load().tspn_relationness(ptr(batch.table), batch.num_videos, batch.total_tracklets,
                                   int(batch.totals[_lib.TOT_MAX_N]), ptr(cls), c, h, *[ptr(t) for t in w],
                                   ptr(scores), PREC[precision], ptr(ws), stream_ptr()), "tspn_relationness")
    _count(2)
    return scores


def relationness_topk_supported(batch: DeviceBatch, precision: str = "fp32", hidden: int = 64) -> bool:
    if PREC[precision] == _lib.PREC_TENSOR:
        return relationness_tc_supported(batch, hidden)
    return bool(load().tspn_relationness_topk_supported(int(batch.totals[_lib.TOT_MAX_N]), int(batch.cls.shape[1])))


def relationness_topk(batch: DeviceBatch, weights, k: int, exclude_diagonal: bool = False, precision: str = "fp32"):
    """``relationness`` + ``topk_pairs`` in two launches (``tspn_relationness_topk``): the scores of a video are
    computed, written and ranked by one CTA.  Returns ``(scores, idx, val, row)``; in ``"fp32"`` precision
    bit-identical to the two calls."""
    cls = _cuda(batch.cls, torch.float32)
    w = [_cuda(t, torch.float32) for t in weights]
    c = int(cls.shape[1])
    h = _check_ppn_shapes(w, c)
    dev, v = cls.device, batch.num_videos
    scores = torch.empty(batch.total(_lib.TOT_SCORES), dtype=torch.float32, device=dev)
    ws = torch.empty(load().tspn_relationness_workspace_bytes(batch.total_tracklets, c, h) // 4,
                     dtype=torch.float32, device=dev)
    idx = torch.empty((v, k), dtype=torch.int64, device=dev)
    val = torch.empty((v, k), dtype=torch.float32, device=dev)
    row = torch.empty((v, k), dtype=torch.int64, device=dev)
    check(load().tspn_relationness_topk(
        ptr(batch.table), v, batch.total_tracklets, int(batch.totals[_lib.TOT_MAX_N]), ptr(cls), c, h,
        *[ptr(t) for t in w], ptr(scores), k,
        _lib.TOPK_EXCLUDE_DIAGONAL if exclude_diagonal else _lib.TOPK_KEEP_DIAGONAL, PREC[precision], ptr(idx),
        ptr(val), ptr(row), ptr(ws), stream_ptr()), "tspn_relationness_topk")
    _count(2 if PREC[precision] != _lib.PREC_TENSOR or int(batch.totals[_lib.TOT_MAX_N]) <= 128 else 3)
    return scores, idx, val, row


def topk_pairs(batch: DeviceBatch, scores: torch.Tensor, k: int, exclude_diagonal: bool = False):
    """Per video: flat indices ``s*N+o`` ``[V, K]`` (-1 beyond K_eff), scores, global pair rows."""
    dev = scores.device
    v = batch.num_videos
    idx = torch.empty((v, k), dtype=torch.int64, device=dev)
    val = torch.empty((v, k), dtype=torch.float32, device=dev)
    row = torch.empty((v, k), dtype=torch.int64, device=dev)
    check(load().tspn_topk_pairs(ptr(batch.table), v, ptr(scores), k,
                                 _lib.TOPK_EXCLUDE_DIAGONAL if exclude_diagonal else _lib.TOPK_KEEP_DIAGONAL,
                                 ptr(idx), ptr(val), ptr(row), stream_ptr()), "tspn_topk_pairs")
    _count(1)
    return idx, val, row


def pack_predicate_weights(weight: torch.Tensor) -> torch.Tensor:
    """One-time repack of ``rel_predictor.weight [R, F]`` for the tensor-core head."""
    weight = _cuda(weight, torch.float32)
    r, f = weight.shape
    nbytes = load().tspn_predicate_packed_bytes(r, f)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    check(load().tspn_pack_predicate_weights(ptr(weight), r, f, ptr(packed), stream_ptr()),
          "tspn_pack_predicate_weights")
    _count(1)
    return packed


def predicate_head(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, precision: str = "fp32",
                   packed: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``sigmoid(x W^T + b)`` (model.py:85-88).  ``x`` may be a row-padded view (stride >= F)."""
    if not x.is_cuda:
        raise RuntimeError("tspn_b200 ops need CUDA tensors (there is no CPU path)")
    prec = PREC[precision]
    if x.stride(-1) != 1:
        x = x.contiguous()
    m, f = x.shape
    r = weight.shape[0]
    is_bf16 = x.dtype == torch.bfloat16
    if not is_bf16 and x.dtype != torch.float32:
        raise TypeError("x must be float32 or bfloat16")
    y = torch.empty((m, r), dtype=torch.float32, device=x.device)
    ws = None
    if prec == _lib.PREC_TENSOR:
        if packed is None:
            packed = pack_predicate_weights(weight)
        nbytes = load().tspn_predicate_workspace_bytes(m, f, r, prec)
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
    check(load().tspn_predicate_head(ptr(x), int(is_bf16), x.stride(0) if m > 1 else max(f, x.stride(0)), m, f,
                                     ptr(_cuda(weight, torch.float32)), ptr(packed),
                                     ptr(_cuda(bias, torch.float32)), r, ptr(y), prec, ptr(ws), stream_ptr()),
          "tspn_predicate_head")
    _count(1 if prec == _lib.PREC_FP32_EXACT or ws is None or ws.numel() <= 16 else 2)
    return y


def tracklet_rows(batch: DeviceBatch) -> torch.Tensor:
    """``[n_tracklets, ld]`` bf16 rows ``[cls | L1-normalised motion | 0]`` for the decomposed predicate head."""
    c = int(batch.cls.shape[1])
    ld = (c + _lib.MOTION_DIM + 7) // 8 * 8
    n = batch.total_tracklets
    out = torch.empty((n, ld), dtype=torch.bfloat16, device=batch.device)
    motion = batch.motion
    check(load().tspn_tracklet_rows(ptr(_cuda(batch.cls, torch.float32)), c, ptr(motion),
                                    int(motion.dtype == torch.uint8), n, ptr(out), ld, stream_ptr()),
          "tspn_tracklet_rows")
    _count(1)
    return out


def predicate_head_affine(x: torch.Tensor, packed: torch.Tensor, n_outputs: int, bias: Optional[torch.Tensor] = None,
                          row_bias: Optional[torch.Tensor] = None, raw: bool = False,
                          k_dim: Optional[int] = None, background: bool = False) -> torch.Tensor:
    """``act(x Wp^T + bias + row_bias[row])`` on the tensor cores (``tspn_predicate_head_affine``); ``packed``
    from ``pack_predicate_weights`` of the ``[n_outputs, k_dim]`` weight slice; ``raw`` skips the sigmoid;
    ``background``: small-footprint launch for a side stream under the pair kernel (TSPN_AFFINE_BACKGROUND)."""
    x = _cuda(x)
    m = x.shape[0]
    f = int(k_dim if k_dim is not None else x.shape[1])
    is_bf16 = x.dtype == torch.bfloat16
    if not is_bf16 and x.dtype != torch.float32:
        raise TypeError("x must be float32 or bfloat16")
    y = torch.empty((m, n_outputs), dtype=torch.float32, device=x.device)
    nbytes = load().tspn_predicate_workspace_bytes(m, f, n_outputs, _lib.PREC_TENSOR)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
    check(load().tspn_predicate_head_affine(
        ptr(x), int(is_bf16), x.stride(0) if m > 1 else max(f, x.stride(0)), m, f, ptr(packed),
        ptr(_cuda(bias, torch.float32)) if bias is not None else None,
        ptr(_cuda(row_bias, torch.float32)) if row_bias is not None else None,
        row_bias.stride(0) if row_bias is not None else 0, n_outputs, ptr(y),
        (_lib.AFFINE_RAW if raw else 0) | (_lib.AFFINE_BACKGROUND if background else 0), ptr(ws),
        stream_ptr()), "tspn_predicate_head_affine")
    _count(1 if ws.numel() <= 16 else 2)
    return y


def assemble_relative(batch: DeviceBatch, geo: torch.Tensor, overlap: torch.Tensor, rows: Optional[torch.Tensor],
                      terms_subject: torch.Tensor, terms_object: torch.Tensor):
    """Per scored row the pooled relative block (bf16 ``[n_rows, 3000]``, [SPEC] s4) and the bias row
    ``A_s[subject] + A_o[object]`` gathered from the per-tracklet terms (``[n_tracklets, R]`` each)."""
    dev = batch.device
    n_rows = int(rows.shape[0]) if rows is not None else batch.total_pairs
    r = int(terms_subject.shape[1])
    rel = torch.empty((n_rows, _lib.REL_DIM), dtype=torch.bfloat16, device=dev)
    row_bias = torch.empty((n_rows, r), dtype=torch.float32, device=dev)
    check(load().tspn_assemble_relative(
        ptr(batch.table), batch.num_videos, batch.total_pairs, int(batch.totals[_lib.TOT_MAX_T]), ptr(geo), ptr(overlap),
        ptr(rows), n_rows, ptr(rel), rel.stride(0), ptr(_cuda(terms_subject, torch.float32)),
        ptr(_cuda(terms_object, torch.float32)), r, ptr(row_bias), stream_ptr()), "tspn_assemble_relative")
    _count(1)
    return rel, row_bias


def survivor_rows_supported(batch: DeviceBatch, n_anchors: int) -> bool:
    return bool(load().tspn_survivor_rows_supported(int(batch.totals[_lib.TOT_MAX_T]), int(n_anchors)))


def gather_pair_terms(batch: DeviceBatch, rows: Optional[torch.Tensor], terms_subject: torch.Tensor,
                      terms_object: torch.Tensor) -> torch.Tensor:
    """Bias rows of the decomposed predicate head: ``A_s[subject] + A_o[object]`` per scored row (zeros for
    padding rows)."""
    n_rows = int(rows.numel()) if rows is not None else batch.total_pairs
    r = int(terms_subject.shape[1])
    out = torch.empty((n_rows, r), dtype=torch.float32, device=batch.device)
    check(load().tspn_gather_pair_terms(ptr(batch.table), batch.num_videos,
                                        ptr(_cuda(rows, torch.int64)) if rows is not None else None, n_rows,
                                        ptr(_cuda(terms_subject, torch.float32)), ptr(_cuda(terms_object, torch.float32)),
                                        r, ptr(out), stream_ptr()), "tspn_gather_pair_terms")
    _count(1)
    return out


def survivor_rows(batch: DeviceBatch, rows: torch.Tensor, terms_subject: Optional[torch.Tensor] = None,
                  terms_object: Optional[torch.Tensor] = None, span_weights=None, sizes: Optional[torch.Tensor] = None,
                  stride: float = 0.0):
    """``tspn_survivor_rows``: for the surviving rows ``rows [V, K]`` (global pair rows, -1 = padding) the pooled
    relative block (bf16 ``[V*K, 3000]``), the bias rows ``A_s[s] + A_o[o]`` and - with ``span_weights =
    (conv_w, conv_b, pred_w, pred_b)`` - the decoded span proposals ``[V*K, L_max * A, 2]`` int32, all
    recomputed from the boxes (bit-identical to ``assemble_relative`` / ``span_proposals`` on stored rows).
    Without the terms no bias rows are produced (``gather_pair_terms`` builds them separately)."""
    dev = batch.device
    rows = _cuda(rows, torch.int64)
    k = int(rows.shape[1])
    n_rows = int(rows.numel())
    max_t = int(batch.totals[_lib.TOT_MAX_T])
    rel = torch.empty((n_rows, _lib.REL_DIM), dtype=torch.bfloat16, device=dev)
    row_bias, r = None, 1
    if terms_subject is not None:
        r = int(terms_subject.shape[1])
        row_bias = torch.empty((n_rows, r), dtype=torch.float32, device=dev)
    spans, a_n, ld_spans = None, 4, 0
    cw = cb = pw = pb = None
    if span_weights is not None:
        cw, cb, pw, pb = (_cuda(w, torch.float32) if w is not None else None for w in span_weights)
        a_n = pw.shape[0] // 2
        ld_spans = span_num_locations(max_t, stride) * a_n * 2
        spans = torch.empty((n_rows, ld_spans // 2, 2), dtype=torch.int32, device=dev)
    check(load().tspn_survivor_rows(
        ptr(batch.table), batch.num_videos, max_t, ptr(batch.boxes), ptr(batch.span), ptr(rows), n_rows, k,
        ptr(rel), rel.stride(0), ptr(_cuda(terms_subject, torch.float32)) if terms_subject is not None else None,
        ptr(_cuda(terms_object, torch.float32)) if terms_object is not None else None, r,
        ptr(row_bias), ptr(cw), ptr(cb), ptr(pw), ptr(pb), a_n, ptr(_cuda(sizes, torch.float32)) if sizes is not None
        else None, float(stride), ptr(spans), ld_spans, stream_ptr()), "tspn_survivor_rows")
    _count(1)
    return rel, row_bias, spans


def span_head(x: torch.Tensor, conv_w: torch.Tensor, conv_b: torch.Tensor, pred_w: torch.Tensor,
              pred_b: torch.Tensor, rows: Optional[torch.Tensor] = None, t: Optional[int] = None,
              precision: str = "fp32", row_base: int = 0) -> torch.Tensor:
    """DPNHead (dpn.py:69-73) on ``x [K, Cin, T]`` or, with ``rows``, on gathered rows of ``x``.

    ``x`` may be a ``[P, Cin, Tp]`` buffer whose rows are padded to ``Tp >= t``.
    """
    x = _cuda(x, torch.float32)
    if x.dim() != 3:
        raise ValueError("span_head expects [K, Cin, T]")
    cin, ld_t = x.shape[1], x.shape[2]
    t = ld_t if t is None else int(t)
    k = x.shape[0] if rows is None else int(rows.shape[0])
    a2 = pred_w.shape[0]
    prec = PREC[precision]
    if x.numel() == 0:        # a video without pairs: every requested row is padding
        return torch.zeros((k, a2, t), dtype=torch.float32, device=x.device)
    out = torch.empty((k, a2, t), dtype=torch.float32, device=x.device)
    ws = None
    if prec == _lib.PREC_TENSOR:
        nbytes = load().tspn_span_head_workspace_bytes(k, cin, t, a2, prec)
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
    check(load().tspn_span_head(ptr(x), ptr(rows), int(row_base), cin * ld_t, ld_t, k, cin, t,
                                ptr(_cuda(conv_w, torch.float32)), ptr(_cuda(conv_b, torch.float32)),
                                ptr(_cuda(pred_w.reshape(a2, cin), torch.float32)),
                                ptr(_cuda(pred_b, torch.float32)), a2, ptr(out), prec, ptr(ws), stream_ptr()),
          "tspn_span_head")
    _count(1)
    return out


def span_proposals(x: torch.Tensor, conv_w: torch.Tensor, conv_b: torch.Tensor, pred_w: torch.Tensor,
                   pred_b: torch.Tensor, sizes: torch.Tensor, stride: float, rows: Optional[torch.Tensor] = None,
                   t: Optional[int] = None, row_base: int = 0) -> torch.Tensor:
    """DPNHead (dpn.py:69-73) + anchors + decode ([SPEC] s5) in one kernel: int32 frame bounds
    ``[K, L*A, 2]``.  The head is evaluated only at the ``L`` anchor columns the decode reads, so the
    ``[K, 2A, T]`` regression tensor is never materialised; bit-identical to
    ``span_decode(span_head(x, ..., precision="fp32"), sizes, stride)``."""
    x = _cuda(x, torch.float32)
    if x.dim() != 3:
        raise ValueError("span_proposals expects [K, Cin, T]")
    cin, ld_t = x.shape[1], x.shape[2]
    t = ld_t if t is None else int(t)
    k = x.shape[0] if rows is None else int(rows.shape[0])
    a2 = pred_w.shape[0]
    n_loc = span_num_locations(t, stride)
    out = torch.empty((k, n_loc * (a2 // 2), 2), dtype=torch.int32, device=x.device)
    if x.numel() == 0:        # a video without pairs: every requested row is padding (regressions 0)
        x = torch.zeros((1, cin, ld_t), dtype=torch.float32, device=x.device)
        rows = torch.full((k,), -1, dtype=torch.int64, device=x.device)
    check(load().tspn_span_proposals(ptr(x), ptr(rows), int(row_base), cin * ld_t, ld_t, k, cin, t,
                                     ptr(_cuda(conv_w, torch.float32)), ptr(_cuda(conv_b, torch.float32)),
                                     ptr(_cuda(pred_w.reshape(a2, cin), torch.float32)),
                                     ptr(_cuda(pred_b, torch.float32)), a2 // 2,
                                     ptr(_cuda(sizes, torch.float32)), float(stride), ptr(out), stream_ptr()),
          "tspn_span_proposals")
    _count(1)
    return out


def span_num_locations(t: int, stride: float) -> int:
    return int(load().tspn_span_num_locations(int(t), float(stride)))


def span_decode(reg: torch.Tensor, sizes: torch.Tensor, stride: float) -> torch.Tensor:
    """Regressions ``[K, 2A, T]`` -> int32 frame bounds ``[K, L*A, 2]`` ([SPEC] s5)."""
    reg = _cuda(reg, torch.float32)
    k, a2, t = reg.shape
    a = a2 // 2
    n_loc = span_num_locations(t, stride)
    out = torch.empty((k, n_loc * a, 2), dtype=torch.int32, device=reg.device)
    check(load().tspn_span_decode(ptr(reg), k, a, t, ptr(_cuda(sizes, torch.float32)), float(stride), ptr(out),
                                  stream_ptr()), "tspn_span_decode")
    _count(1)
    return out


def span_select(cands: torch.Tensor, n_anchors: int, stride: float, n_keep: int = 64, nms_threshold: float = 0.5,
                batch: Optional[DeviceBatch] = None, rows: Optional[torch.Tensor] = None,
                windows: Optional[torch.Tensor] = None, int16: bool = True):
    """Temporal NMS + top-``n_keep`` of decoded span proposals ([SPEC] s8; what ``RelNMS``,
    lib/modeling/relpn/rel_nms.py:6-15, is meant to do).  ``cands [n_rows, M, 2]`` int32.

    With ``batch``: row r scores global pair row ``rows[r]`` (``None`` = r, negative = padding), has
    ``locations(T_video) * n_anchors`` candidates and - unless ``windows`` is given - the temporal overlap window
    of its two tracklets.  Without: every row has ``M`` candidates and the window ``windows[r]``.
    Returns ``(kept [n_rows, n_keep, 2] int16 | int32 zero padded, counts [n_rows] int32)``."""
    cands = _cuda(cands, torch.int32)
    n_rows, m = int(cands.shape[0]), int(cands.shape[1])
    dev = cands.device
    out = torch.empty((n_rows, n_keep, 2), dtype=torch.int16 if int16 else torch.int32, device=dev)
    counts = torch.empty(n_rows, dtype=torch.int32, device=dev)
    if windows is not None:
        windows = _cuda(windows, torch.int32)
    if rows is not None:
        rows = _cuda(rows, torch.int64)
    check(load().tspn_span_select(
        ptr(batch.table) if batch is not None else None, batch.num_videos if batch is not None else 0,
        int(batch.totals[_lib.TOT_MAX_T]) if batch is not None else 0,
        ptr(batch.span) if batch is not None else None, ptr(rows), n_rows, ptr(windows), ptr(cands), 2 * m, m,
        int(n_anchors), float(stride), int(n_keep), float(nms_threshold), _lib.SPANS_I16 if int16 else 0, ptr(out),
        ptr(counts), stream_ptr()), "tspn_span_select")
    _count(1)
    return out, counts


RECORD_FIELDS = ("score", "s_cls", "pred", "o_cls", "s_tid", "o_tid", "start", "end")


def postprocess(batch: DeviceBatch, logits: torch.Tensor, overlap: Optional[torch.Tensor], topk_per_pair: int = 20,
                topk_per_video: int = 200, rows: Optional[torch.Tensor] = None,
                row_video_off: Optional[torch.Tensor] = None, mirror_q4: bool = False):
    """Relation triplet records (predict.py:66-117): ``records [V, topk_per_video, 8]`` int32
    (field 0 is the fp32 score's bit pattern) and ``counts [V]`` int32.  ``overlap=None``: the windows are
    derived from the tracklet spans (no dependence on the pair kernel's outputs)."""
    logits = _cuda(logits, torch.float32)
    m, r = logits.shape
    v = batch.num_videos
    dev = logits.device
    records = torch.empty((v, topk_per_video, 8), dtype=torch.int32, device=dev)
    counts = torch.empty(v, dtype=torch.int32, device=dev)
    ws = torch.empty(load().tspn_postprocess_workspace_bytes(m, topk_per_pair), dtype=torch.uint8, device=dev)
    check(load().tspn_postprocess(ptr(batch.table), v, ptr(logits), ptr(rows), ptr(row_video_off), m, r,
                                  ptr(batch.cls), batch.cls.shape[1], ptr(overlap), ptr(batch.span), topk_per_pair,
                                  topk_per_video,
                                  1 if mirror_q4 else 0, ptr(records), ptr(counts), ptr(ws), stream_ptr()),
          "tspn_postprocess")
    _count(2)
    return records, counts


def record_scores(records: torch.Tensor) -> torch.Tensor:
    return records[..., 0].contiguous().view(torch.float32)
