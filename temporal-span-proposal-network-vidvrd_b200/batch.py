"""Batch layout of the pair stage: several videos per launch.

The reference processes one 30-frame segment per forward (lib/modeling/predict.py:42-57,
``TEST_BATCH_SIZE: 1``); per-video problems are far too small for a B200, so the CUDA
entry points take a *batch* of videos described by the video table of
``include/tspn_b200.h``.  ``HostBatch`` packs per-video arrays into pinned host buffers in
that layout (box rows padded to a multiple of 8 frames for the TMA tile), ``DeviceBatch`` is
its image in HBM.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (TOT_BOXES, TOT_GEO_FLOATS, TOT_ITEMS, TOT_MAX_N, TOT_MAX_T, TOT_PAIRS, TOT_SCORES,
                   TOT_TRACKLETS, VT_BOX_OFF, VT_COLS, VT_GEO_OFF, VT_N, VT_PAIR_OFF, VT_SCORE_OFF, VT_T,
                   VT_TB, VT_TP, VT_TRK_OFF)


def _np(x, dtype):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


_ARENA_ALIGN = 256


def _arena_layout(fields):
    """``{name: (offset, shape, dtype)}`` + total bytes for the fields that are present."""
    out, off = {}, 0
    for f in fields:
        if f is None:
            continue
        name, shape, dtype = f
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        out[name] = (off, tuple(int(x) for x in shape), dtype)
        off = (off + nbytes + _ARENA_ALIGN - 1) // _ARENA_ALIGN * _ARENA_ALIGN
    out["bytes"] = max(off, _ARENA_ALIGN)
    return out


def _arena_views(arena: torch.Tensor, layout):
    views = {}
    for name, spec in layout.items():
        if name == "bytes":
            continue
        off, shape, dtype = spec
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        views[name] = arena[off:off + nbytes].view(dtype).view(shape)
    return views


class HostBatch:
    """Packed, pinned host image of a batch of videos."""

    def __init__(self, boxes: Sequence, span: Sequence, cls: Optional[Sequence] = None,
                 motion: Optional[Sequence] = None, pin: bool = True, compact: bool = True):
        """``compact``: ship boxes as u16 pixel coordinates and motion histograms as u8 counts when every
        value is exactly representable (integer boxes in [0, 65535], integer counts in [0, 255]) - lossless,
        expanded on the device by ``tspn_unpack_boxes_u16`` / ``tspn_normalize_motion_u8``; otherwise (or
        with ``compact=False``) the fields travel as fp32.  The serving loop is PCIe-bound, and these two
        fields are 99 % of a batch's bytes."""
        boxes = [_np(b, np.float32) for b in boxes]
        span = [_np(s, np.int32).reshape(-1, 2) for s in span]
        assert len(boxes) == len(span)
        self.n = [int(b.shape[0]) for b in boxes]
        self.t = [int(b.shape[1]) if b.ndim == 3 else 1 for b in boxes]
        for b, s in zip(boxes, span):
            if b.ndim != 3 or b.shape[2] != 4 or s.shape[0] != b.shape[0]:
                raise ValueError("boxes must be [N, T, 4] with span [N, 2]; got %s / %s" % (b.shape, s.shape))
            if s.size and (s.min() < 0 or s[:, 1].max() > b.shape[1] or (s[:, 0] > s[:, 1]).any()):
                raise ValueError("span must satisfy 0 <= pstart <= pend <= T")
        self.table, self.totals = _lib.build_video_table(self.n, self.t)
        tot = self.totals
        use_pin = pin and torch.cuda.is_available()
        n_trk = int(tot[TOT_TRACKLETS])
        if cls is not None:
            cls = [_np(c, np.float32) for c in cls]
        if motion is not None:
            motion = [_np(m, np.float32) for m in motion]
        def exact_ints(arrays, hi):
            return all(a.size == 0 or (a.min() >= 0 and a.max() <= hi and np.array_equal(a, np.rint(a)))
                       for a in arrays)
        self.boxes_compact = bool(compact) and exact_ints(boxes, 65535)
        self.motion_compact = bool(compact) and motion is not None and exact_ints(motion, 255)
        # One pinned arena holds every field (256-byte aligned segments), so that a step's inputs cross
        # PCIe as ONE copy; the fields below are views of it.  DeviceBatch mirrors the layout in HBM.
        self.layout = _arena_layout([
            ("table", (len(self.n), VT_COLS), torch.int64),
            ("boxes", (int(tot[TOT_BOXES]), 4), torch.int16 if self.boxes_compact else torch.float32),
            ("span", (n_trk, 2), torch.int32),
            ("cls", (n_trk, int(cls[0].shape[1])), torch.float32) if cls is not None else None,
            ("motion", (n_trk, _lib.MOTION_DIM), torch.uint8 if self.motion_compact else torch.float32)
            if motion is not None else None,
        ])
        self.arena = torch.zeros(self.layout["bytes"], dtype=torch.uint8)
        if use_pin:
            self.arena = self.arena.pin_memory()
        views = _arena_views(self.arena, self.layout)
        self.boxes, self.span = views["boxes"], views["span"]     # boxes: fp32, or the u16 bit patterns (int16 view)
        bview = self.boxes.numpy().view(np.uint16) if self.boxes_compact else self.boxes.numpy()
        sview = self.span.numpy()
        for v, (b, s) in enumerate(zip(boxes, span)):
            row = self.table[v]
            n, t, tb = int(row[VT_N]), int(row[VT_T]), int(row[VT_TB])
            if n == 0:
                continue
            dst = bview[int(row[VT_BOX_OFF]):int(row[VT_BOX_OFF]) + n * tb].reshape(n, tb, 4)
            dst[:, :t] = b
            sview[int(row[VT_TRK_OFF]):int(row[VT_TRK_OFF]) + n] = s
        self.cls = views.get("cls")
        self.motion = views.get("motion")
        if cls is not None and n_trk:
            self.cls.numpy()[:] = np.concatenate(cls, axis=0)
        if motion is not None and n_trk:
            self.motion.numpy()[:] = np.concatenate(motion, axis=0)
        self.table_t = views["table"]
        self.table_t.numpy()[:] = np.ascontiguousarray(self.table).reshape(-1, VT_COLS)

    @classmethod
    def from_videos(cls, videos, pin: bool = True, compact: bool = True) -> "HostBatch":
        """From ``tspn_b200.synth.VideoTracklets`` (or anything with boxes/span/cls/motion)."""
        return cls([v.boxes for v in videos], [v.span for v in videos], [v.cls for v in videos],
                   [v.motion for v in videos], pin=pin, compact=compact)

    @property
    def num_videos(self) -> int:
        return len(self.n)

    def h2d_bytes(self) -> int:
        """Bytes one step copies host -> device (the whole arena, alignment padding included)."""
        return int(self.arena.numel())

    def to_device(self, device="cuda", non_blocking: bool = True) -> "DeviceBatch":
        return DeviceBatch(self, device, non_blocking)


class DeviceBatch:
    """HBM-resident batch: the pointers the C ABI consumes."""

    def __init__(self, host: HostBatch, device="cuda", non_blocking: bool = True):
        self.host = host
        self.n, self.t = host.n, host.t
        self.table_host, self.totals = host.table, host.totals
        dev = torch.device(device)
        self.device = dev
        self.layout = host.layout
        self.arena = host.arena.to(dev, non_blocking=non_blocking)        # one H2D copy
        views = _arena_views(self.arena, self.layout)
        self.table, self.span = views["table"], views["span"]
        self.cls, self.motion = views.get("cls"), views.get("motion")     # motion: fp32, or u8 counts (compact)
        if host.boxes_compact:
            # u16 coordinates in the arena, expanded right behind the copy (same stream) into this fp32 buffer -
            # the layout the kernels and the TMA tensor map read.  The expansion belongs to the upload, not to
            # the step: in the serving loop it runs on the H2D stream, off the kernels' critical path.
            self.boxes_u16 = views["boxes"]
            self.boxes = torch.empty(self.boxes_u16.shape, dtype=torch.float32, device=dev)
            self._unpack()
        else:
            self.boxes_u16 = None
            self.boxes = views["boxes"]

    def _unpack(self) -> None:
        if self.boxes_u16 is not None and self.boxes.shape[0] > 0:
            with torch.cuda.device(self.device):
                _lib.check(_lib.load().tspn_unpack_boxes_u16(self.boxes_u16.data_ptr(), self.boxes.shape[0],
                                                             self.boxes.data_ptr(),
                                                             torch.cuda.current_stream(self.device).cuda_stream),
                           "tspn_unpack_boxes_u16")

    def copy_from(self, host: HostBatch) -> "DeviceBatch":
        """Refill the device buffers from another host batch of the same layout (non-blocking, on the
        current stream): the steady-state H2D of a serving loop."""
        if host.n != self.n or host.t != self.t or host.layout != self.layout:
            raise ValueError("copy_from needs a host batch with the same per-video shapes and fields")
        self.host = host
        self.arena.copy_(host.arena, non_blocking=True)                   # one H2D copy
        self._unpack()
        return self

    # sizes -----------------------------------------------------------------------------
    @property
    def num_videos(self) -> int:
        return len(self.n)

    def total(self, col: int) -> int:
        return int(self.totals[col])

    @property
    def total_pairs(self) -> int:
        return self.total(TOT_PAIRS)

    @property
    def total_tracklets(self) -> int:
        return self.total(TOT_TRACKLETS)

    # per-video views -----------------------------------------------------------------------
    def pair_slice(self, v: int) -> slice:
        off = int(self.table_host[v, VT_PAIR_OFF])
        n = self.n[v]
        return slice(off, off + n * max(n - 1, 0))

    def tracklet_slice(self, v: int) -> slice:
        off = int(self.table_host[v, VT_TRK_OFF])
        return slice(off, off + self.n[v])

    def score_view(self, scores: torch.Tensor, v: int) -> torch.Tensor:
        off = int(self.table_host[v, VT_SCORE_OFF])
        n = self.n[v]
        return scores[off:off + n * n].view(n, n)

    def geo_rows(self, geo: torch.Tensor, v: int) -> torch.Tensor:
        """``[P, 8, Tp]`` view of video v's geometry including the row pad (Tp = T rounded up to 4)."""
        row = self.table_host[v]
        n, tp = int(row[VT_N]), int(row[VT_TP])
        p = n * max(n - 1, 0)
        off = int(row[VT_GEO_OFF])
        return geo[off:off + p * _lib.GEO_CHANNELS * tp].view(p, _lib.GEO_CHANNELS, tp)

    def geo_view(self, geo: torch.Tensor, v: int) -> torch.Tensor:
        """``[P, 8, T]`` view of video v's geometry (the row pad to Tp is sliced off)."""
        return self.geo_rows(geo, v)[:, :, :self.t[v]]
