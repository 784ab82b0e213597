"""Batch layout of the pair stage: several videos per launch.

The reference processes one 30-frame segment per forward (lib/modeling/predict.py:42-57,
``TEST_BATCH_SIZE: 1``); per-video problems are far too small for a B200, so the CUDA
entry points take a *batch* of videos described by the video table of
``include/tspn_b200.h``.  ``HostBatch`` packs per-video arrays into pinned host buffers in
that layout (box rows padded to a multiple of 8 frames for the TMA tile), ``DeviceBatch`` is
its image in HBM.

Real VidVRD / VidOR batches are ragged (every video has its own tracklet and frame count).  A
``Capacity`` fixes the *sizes of the buffers and grids* of a batch - videos, tracklets, pairs,
geometry floats, work items, boxes, scores, the pair kernel's chunk - while the table (whose sentinel
row carries the true totals to the device) describes what is actually in it.  Every batch packed for
the same capacity has the same arena layout and the same launch arguments, so ONE recorded CUDA graph
serves all of them (``pipeline.GraphedStage``, ``serving.PipelinedStage``); ``pack_batches`` splits a
list of ragged videos into such batches.
"""
from __future__ import annotations

import ctypes
import dataclasses
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (TOT_BOXES, TOT_COLS, TOT_GEO_CHUNK, TOT_GEO_FLOATS, TOT_ITEMS, TOT_MAX_CHUNKS, TOT_MAX_N, TOT_MAX_T,
                   TOT_PAIRS, TOT_SCORES, TOT_TRACKLETS, VT_BOX_OFF, VT_COLS, VT_GEO_OFF, VT_N, VT_PAIR_OFF,
                   VT_SCORE_OFF, VT_T, VT_TB, VT_TP, VT_TRK_OFF)


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


def _exact_ints(arrays, hi) -> bool:
    return all(a.size == 0 or (a.min() >= 0 and a.max() <= hi and np.array_equal(a, np.rint(a))) for a in arrays)


@dataclasses.dataclass(frozen=True)
class Capacity:
    """Upper bounds that fix a batch's buffer sizes, grid sizes and transport dtypes (one CUDA graph per
    capacity).  ``boxes_u16`` / ``motion_u8``: the compact host->device transport is a property of the
    capacity, not of the data - a batch whose values do not fit raises instead of silently changing layout."""
    videos: int
    tracklets: int
    pairs: int
    geo_floats: int
    items: int
    boxes: int
    scores: int
    max_n: int
    max_t: int
    geo_chunk: int
    max_chunks: int
    n_classes: int
    boxes_u16: bool = True
    motion_u8: bool = True
    has_motion: bool = True

    def totals(self) -> np.ndarray:
        """The launch totals (``TOT_*`` layout) of every batch of this capacity."""
        tot = np.zeros(TOT_COLS, dtype=np.int64)
        tot[TOT_TRACKLETS], tot[TOT_PAIRS], tot[TOT_GEO_FLOATS] = self.tracklets, self.pairs, self.geo_floats
        tot[TOT_ITEMS], tot[TOT_BOXES], tot[TOT_SCORES] = self.items, self.boxes, self.scores
        tot[TOT_MAX_N], tot[TOT_MAX_T], tot[TOT_GEO_CHUNK] = self.max_n, self.max_t, self.geo_chunk
        tot[TOT_MAX_CHUNKS] = self.max_chunks
        return tot

    @classmethod
    def for_shapes(cls, shapes: Sequence[Tuple[int, int]], n_classes: int, max_n: Optional[int] = None,
                   max_t: Optional[int] = None, videos: Optional[int] = None, boxes_u16: bool = True,
                   motion_u8: bool = True, has_motion: bool = True) -> "Capacity":
        """The tightest capacity that holds a batch of these ``(N, T)`` videos; ``max_n`` / ``max_t`` /
        ``videos`` may be raised to the bounds of a whole bucket (they select kernel variants and size
        per-video outputs, not memory per pair)."""
        n = [int(s[0]) for s in shapes]
        t = [int(s[1]) for s in shapes]
        mt = max([max_t or 1] + t)
        chunk = int(_lib.load().tspn_geo_chunk(mt))
        _, tot = _lib.build_video_table(n, t, geo_chunk=chunk)
        return cls(videos=max(videos or 0, len(n), 1), tracklets=int(tot[TOT_TRACKLETS]), pairs=int(tot[TOT_PAIRS]),
                   geo_floats=int(tot[TOT_GEO_FLOATS]), items=int(tot[TOT_ITEMS]), boxes=int(tot[TOT_BOXES]),
                   scores=int(tot[TOT_SCORES]), max_n=max([max_n or 0] + n), max_t=mt, geo_chunk=chunk,
                   max_chunks=(mt + chunk - 1) // chunk, n_classes=int(n_classes), boxes_u16=bool(boxes_u16),
                   motion_u8=bool(motion_u8), has_motion=bool(has_motion))

    def grown(self, **kw) -> "Capacity":
        return dataclasses.replace(self, **kw)

    def fits(self, tot: np.ndarray, n_videos: int) -> bool:
        return (n_videos <= self.videos and tot[TOT_TRACKLETS] <= self.tracklets and tot[TOT_PAIRS] <= self.pairs
                and tot[TOT_GEO_FLOATS] <= self.geo_floats and tot[TOT_ITEMS] <= self.items
                and tot[TOT_BOXES] <= self.boxes and tot[TOT_SCORES] <= self.scores and tot[TOT_MAX_N] <= self.max_n
                and tot[TOT_MAX_T] <= self.max_t)


class HostBatch:
    """Packed, pinned host image of a batch of videos."""

    def __init__(self, boxes: Sequence, span: Sequence, cls: Optional[Sequence] = None,
                 motion: Optional[Sequence] = None, pin: bool = True, compact: bool = True,
                 capacity: Optional[Capacity] = None, delta: Optional[bool] = None):
        """``compact`` (batches without a capacity): ship boxes as u16 pixel coordinates and motion
        histograms as u8 counts when every value is exactly representable (integer boxes in [0, 65535],
        integer counts in [0, 255]) - lossless, expanded on the device by ``tspn_unpack_boxes_u16`` /
        ``tspn_normalize_motion_u8``; otherwise (or with ``compact=False``) the fields travel as fp32.  The
        serving loop is PCIe-bound, and these two fields are 99 % of a batch's bytes.  Compact boxes are also
        SPAN-PACKED: only the frames ``[pstart, pend)`` of each tracklet travel (nothing on the path reads a box
        outside its tracklet's span), expanded into the dense zero-padded rows by ``tspn_unpack_boxes_spans``.
        ``delta`` (default on; environment ``TSPN_BOX_DELTA=0`` turns it off): span-packed tracklets whose boxes move
        by at most [-128, 127] pixels per frame travel as a first box + 4 x i8 differences per frame - half the
        bytes, decoded by an integer prefix sum on the device; any other tracklet stays raw u16 (decided per
        tracklet, so the arena layout - sized for raw - does not depend on the data).

        ``capacity``: pack for that capacity instead (serving: every batch of a capacity shares one arena
        layout and one captured graph).  The transport dtypes are then the capacity's; data that does not fit
        them - a fractional coordinate, a count above 255 - raises ``ValueError``."""
        boxes = [_np(b, np.float32) for b in boxes]
        span = [_np(s, np.int32).reshape(-1, 2) for s in span]
        assert len(boxes) == len(span)
        if delta is None:
            delta = os.environ.get("TSPN_BOX_DELTA", "1") != "0"
        n_act = [int(b.shape[0]) for b in boxes]
        t_act = [int(b.shape[1]) if b.ndim == 3 else 1 for b in boxes]
        for b, s in zip(boxes, span):
            if b.ndim != 3 or b.shape[2] != 4 or s.shape[0] != b.shape[0]:
                raise ValueError("boxes must be [N, T, 4] with span [N, 2]; got %s / %s" % (b.shape, s.shape))
            if s.size and (s.min() < 0 or s[:, 1].max() > b.shape[1] or (s[:, 0] > s[:, 1]).any()):
                raise ValueError("span must satisfy 0 <= pstart <= pend <= T")
        if cls is not None:
            cls = [_np(c, np.float32) for c in cls]
        if motion is not None:
            motion = [_np(m, np.float32) for m in motion]
        self.capacity = capacity
        self.num_real = len(n_act)
        if capacity is None:
            self.table, self.actual = _lib.build_video_table(n_act, t_act)
            self.totals = self.actual
            self.boxes_compact = bool(compact) and _exact_ints(boxes, 65535)
            self.motion_compact = bool(compact) and motion is not None and _exact_ints(motion, 255)
            n_cls = int(cls[0].shape[1]) if cls else 0
            has_cls, has_motion = cls is not None, motion is not None
            rows = len(n_act)
        else:
            cap = capacity
            self.table, self.actual = _lib.build_video_table(n_act, t_act, table_rows=cap.videos,
                                                             geo_chunk=cap.geo_chunk)
            if not cap.fits(self.actual, len(n_act)):
                raise ValueError("batch does not fit its capacity: totals %s (videos %d) > %s"
                                 % (self.actual.tolist(), len(n_act), cap))
            self.totals = cap.totals()
            self.boxes_compact, self.motion_compact = cap.boxes_u16, cap.motion_u8 and cap.has_motion
            if self.boxes_compact and not _exact_ints(boxes, 65535):
                raise ValueError("this capacity ships boxes as u16 pixel coordinates, but a box coordinate is "
                                 "fractional or outside [0, 65535]; use a capacity with boxes_u16=False")
            if self.motion_compact and motion is not None and not _exact_ints(motion, 255):
                raise ValueError("this capacity ships motion histograms as u8 counts, but a value is fractional "
                                 "or above 255; use a capacity with motion_u8=False")
            if cap.has_motion and motion is None or cls is None:
                raise ValueError("a capacity batch carries the cls%s fields" % (" and motion" if cap.has_motion else ""))
            if cls and int(cls[0].shape[1]) != cap.n_classes:
                raise ValueError("cls has %d classes, the capacity %d" % (cls[0].shape[1], cap.n_classes))
            n_cls, has_cls, has_motion, rows = cap.n_classes, True, cap.has_motion, cap.videos
        # per-video sizes over all table rows (padding rows are empty videos)
        self.n = n_act + [0] * (rows - len(n_act))
        self.t = t_act + [1] * (rows - len(t_act))
        tot = self.totals
        use_pin = pin and torch.cuda.is_available()
        n_trk = int(tot[TOT_TRACKLETS])
        # One pinned arena holds every field (256-byte aligned segments), so that a step's inputs cross
        # PCIe in one piece; the fields below are views of it.  DeviceBatch mirrors the layout in HBM.
        self.layout = _arena_layout([
            ("table", (rows + 1, VT_COLS), torch.int64),
            ("boxes", (int(tot[TOT_BOXES]), 4), torch.int16 if self.boxes_compact else torch.float32),
            ("box_off", (n_trk,), torch.int64) if self.boxes_compact else None,
            ("span", (n_trk, 2), torch.int32),
            ("cls", (n_trk, n_cls), torch.float32) if has_cls else None,
            ("motion", (n_trk, _lib.MOTION_DIM), torch.uint8 if self.motion_compact else torch.float32)
            if has_motion else None,
        ])
        self.arena = torch.zeros(self.layout["bytes"], dtype=torch.uint8)
        if use_pin:
            self.arena = self.arena.pin_memory()
        views = _arena_views(self.arena, self.layout)
        self.boxes, self.span = views["boxes"], views["span"]     # boxes: fp32 dense rows, or span-packed u16
        self.box_off = views.get("box_off")                       # first packed box of every tracklet (compact)
        bview = self.boxes.numpy().view(np.uint16) if self.boxes_compact else self.boxes.numpy()
        sview = self.span.numpy()
        packed = 0
        for v, (b, s) in enumerate(zip(boxes, span)):
            row = self.table[v]
            n, t, tb = int(row[VT_N]), int(row[VT_T]), int(row[VT_TB])
            if n == 0:
                continue
            trk0 = int(row[VT_TRK_OFF])
            sview[trk0:trk0 + n] = s
            if self.boxes_compact:
                packed += self._pack_span_boxes(b, s, trk0, packed, bview, delta)
            else:
                dst = bview[int(row[VT_BOX_OFF]):int(row[VT_BOX_OFF]) + n * tb].reshape(n, tb, 4)
                dst[:, :t] = b
        self.packed_boxes = packed                                # boxes that travel (compact transport)
        self.cls = views.get("cls")
        self.motion = views.get("motion")
        n_act_trk = int(self.actual[TOT_TRACKLETS])
        if cls is not None and n_act_trk:
            self.cls.numpy()[:n_act_trk] = np.concatenate(cls, axis=0)
        if motion is not None and n_act_trk and self.motion is not None:
            self.motion.numpy()[:n_act_trk] = np.concatenate(motion, axis=0)
        self.table_t = views["table"]
        self.table_t.numpy()[:] = np.ascontiguousarray(self.table).reshape(-1, VT_COLS)

    def _pack_span_boxes(self, b: np.ndarray, s: np.ndarray, trk0: int, slot0: int, bview: np.ndarray,
                         delta: bool) -> int:
        """Span-packed transport of one video's boxes into the u16 arena, from 8-byte slot ``slot0`` on; returns the
        slots used.  Tracklet-major; per tracklet either RAW (a slot of 4 x u16 per frame) or, when every coordinate
        moves by at most [-128, 127] pixels per frame, DELTA (the first frame's slot, then 4 x i8 per further frame,
        two frames per slot; ``_lib.PACKED_DELTA`` set in its ``box_off``) - see ``tspn_unpack_boxes_spans``.  The
        loops run in the library (``tspn_host_pack_boxes_spans``, csrc/host_pack.cu: CPU code): the numpy form of this
        method took 11 ms per VidOR video - 250 GPU steps per packed batch (kept as the reference in
        tests/test_cpu_ragged.py)."""
        n = int(b.shape[0])
        b = np.ascontiguousarray(b, dtype=np.float32)
        s = np.ascontiguousarray(s, dtype=np.int32)
        off = self.box_off.numpy()
        used = ctypes.c_int64(0)
        _lib.check(_lib.load().tspn_host_pack_boxes_spans(
            b.ctypes.data, n, int(b.shape[1]), s.ctypes.data, int(bool(delta)), bview.ctypes.data,
            int(bview.shape[0]), int(slot0), off[trk0:].ctypes.data, ctypes.addressof(used)),
            "tspn_host_pack_boxes_spans")
        return int(used.value)

    def unpacked_boxes(self) -> List[np.ndarray]:
        """Host-side decode of the box transport (what ``tspn_unpack_boxes_spans`` does on the device): the dense
        ``[N, T, 4]`` float32 boxes of every real video.  For tests of the packer; nothing on the product path uses it."""
        out = []
        sview = self.span.numpy()
        for v in range(self.num_real):
            row = self.table[v]
            n, t, tb, trk0 = int(row[VT_N]), int(row[VT_T]), int(row[VT_TB]), int(row[VT_TRK_OFF])
            dense = np.zeros((n, t, 4), dtype=np.float32)
            if not self.boxes_compact:
                b0 = int(row[VT_BOX_OFF])
                dense[:] = self.boxes.numpy()[b0:b0 + n * tb].reshape(n, tb, 4)[:, :t]
            else:
                u16 = self.boxes.numpy().view(np.uint16)
                for i in range(n):
                    ps, pe = int(sview[trk0 + i, 0]), int(sview[trk0 + i, 1])
                    po = int(self.box_off.numpy()[trk0 + i])
                    slot, length = po & ~_lib.PACKED_DELTA, pe - ps
                    if length <= 0:
                        continue
                    if not (po & _lib.PACKED_DELTA):
                        dense[i, ps:pe] = u16[slot:slot + length]
                    else:
                        dl = u16[slot + 1:slot + 1 + length // 2].view(np.int8).reshape(-1, 4)[:length - 1]
                        dense[i, ps:pe] = np.concatenate([u16[slot:slot + 1].astype(np.int64),
                                                          dl.astype(np.int64)]).cumsum(axis=0)
            out.append(dense)
        return out

    @classmethod
    def from_videos(cls, videos, pin: bool = True, compact: bool = True,
                    capacity: Optional[Capacity] = None, delta: Optional[bool] = None) -> "HostBatch":
        """From ``tspn_b200.synth.VideoTracklets`` (or anything with boxes/span/cls/motion)."""
        return cls([v.boxes for v in videos], [v.span for v in videos], [v.cls for v in videos],
                   [v.motion for v in videos], pin=pin, compact=compact, capacity=capacity, delta=delta)

    @property
    def num_videos(self) -> int:
        """Table rows (a capacity batch counts its empty padding videos; ``num_real`` are the caller's)."""
        return len(self.n)

    def used_segments(self) -> List[Tuple[int, int]]:
        """``(offset, bytes)`` of the part of every arena field the batch actually fills: what a step copies."""
        act = self.actual
        used = {"table": None, "boxes": self.packed_boxes if self.boxes_compact else int(act[TOT_BOXES]),
                "box_off": int(act[TOT_TRACKLETS]), "span": int(act[TOT_TRACKLETS]),
                "cls": int(act[TOT_TRACKLETS]), "motion": int(act[TOT_TRACKLETS])}
        segs = []
        for name, spec in self.layout.items():
            if name == "bytes":
                continue
            off, shape, dtype = spec
            rows = shape[0] if used[name] is None else min(used[name], shape[0])
            nbytes = rows * int(np.prod(shape[1:])) * torch.empty((), dtype=dtype).element_size()
            if nbytes:
                segs.append((off, nbytes))
        return segs

    def h2d_bytes(self) -> int:
        """Bytes one step copies host -> device (the part of every arena field the batch fills)."""
        return int(sum(b for _, b in self.used_segments()))

    def to_device(self, device="cuda", non_blocking: bool = True) -> "DeviceBatch":
        return DeviceBatch(self, device, non_blocking)


class DeviceBatch:
    """HBM-resident batch: the pointers the C ABI consumes."""

    def __init__(self, host: HostBatch, device="cuda", non_blocking: bool = True):
        dev = torch.device(device)
        self.device = dev
        self.layout = host.layout
        self.capacity = host.capacity
        self.totals = host.totals                    # launch totals: the capacity's, or the batch's own
        self._adopt(host)
        self.arena = host.arena.to(dev, non_blocking=non_blocking)        # one H2D copy
        views = _arena_views(self.arena, self.layout)
        self.table, self.span = views["table"], views["span"]
        self.cls, self.motion = views.get("cls"), views.get("motion")     # motion: fp32, or u8 counts (compact)
        self.box_off = views.get("box_off")
        self.geo_off = self.geo_total = None         # WINDOWED geometry layout: see enable_windows()
        if host.boxes_compact:
            # span-packed u16 coordinates in the arena, expanded right behind the copy (same stream) into this fp32 buffer -
            # the layout the kernels and the TMA tensor map read.  The expansion belongs to the upload, not to
            # the step: in the serving loop it runs on the H2D stream, off the kernels' critical path.
            self.boxes_u16 = views["boxes"]
            self.boxes = torch.zeros(self.boxes_u16.shape, dtype=torch.float32, device=dev)
            self._unpack()
        else:
            self.boxes_u16 = None
            self.boxes = views["boxes"]

    def _adopt(self, host: HostBatch) -> None:
        self.host = host
        self.n, self.t = host.n, host.t
        self.num_real = host.num_real
        self.table_host, self.actual = host.table, host.actual

    def _unpack(self) -> None:
        if self.boxes_u16 is not None and int(self.actual[TOT_TRACKLETS]) > 0:
            with torch.cuda.device(self.device):
                _lib.check(_lib.load().tspn_unpack_boxes_spans(
                    self.table.data_ptr(), len(self.n), int(self.totals[TOT_TRACKLETS]), self.span.data_ptr(),
                    self.box_off.data_ptr(), self.boxes_u16.data_ptr(), self.boxes.data_ptr(),
                    torch.cuda.current_stream(self.device).cuda_stream), "tspn_unpack_boxes_spans")

    def enable_windows(self):
        """Row offsets of the opt-in WINDOWED geometry layout (``tspn_geo_window_offsets``): ``geo_off`` int64
        ``[sum P]`` and ``geo_total`` int64 ``[1]`` (floats), computed from the spans on the current stream now and
        again behind every refill (``copy_from`` / ``copy_from_device``) - part of the upload, like the box expansion,
        not of the step.  The tensors keep their addresses (captured graphs read them)."""
        if self.geo_off is None:
            p = max(int(self.totals[TOT_PAIRS]), 1)
            self.geo_off = torch.zeros(p, dtype=torch.int64, device=self.device)
            self.geo_total = torch.zeros(1, dtype=torch.int64, device=self.device)
            self._windows()
        return self.geo_off, self.geo_total

    def _windows(self) -> None:
        if self.geo_off is not None:
            with torch.cuda.device(self.device):
                _lib.check(_lib.load().tspn_geo_window_offsets(
                    self.table.data_ptr(), len(self.n), int(self.totals[TOT_PAIRS]), self.span.data_ptr(),
                    self.geo_off.data_ptr(), self.geo_total.data_ptr(),
                    torch.cuda.current_stream(self.device).cuda_stream), "tspn_geo_window_offsets")

    def copy_from(self, host: HostBatch) -> "DeviceBatch":
        """Refill the device buffers from another host batch of the same layout (non-blocking, on the
        current stream): the steady-state H2D of a serving loop.  Without a capacity the batch must have the
        same per-video shapes; with one, any batch packed for that capacity."""
        if host.layout != self.layout or host.capacity != self.capacity or \
                (self.capacity is None and (host.n != self.n or host.t != self.t)):
            raise ValueError("copy_from needs a host batch of the same capacity (or, without one, the same "
                             "per-video shapes and fields)")
        self._adopt(host)
        for off, nbytes in host.used_segments():                          # only what the batch fills
            self.arena[off:off + nbytes].copy_(host.arena[off:off + nbytes], non_blocking=True)
        self._unpack()
        self._windows()
        return self

    def copy_from_device(self, other: "DeviceBatch") -> "DeviceBatch":
        """The same refill from a batch that is already resident in HBM (device-to-device, current stream): the
        fields the kernels read - table, spans, class logits, motion rows and the fp32 boxes (the u16 transport
        image of the boxes is not copied: it only feeds the expansion)."""
        if other.layout != self.layout or other.capacity != self.capacity:
            raise ValueError("copy_from_device needs a batch of the same capacity")
        self._adopt(other.host)
        skip = (self.layout["boxes"][0], self.layout["box_off"][0]) if self.boxes_u16 is not None else ()
        for off, nbytes in other.host.used_segments():
            if off not in skip:
                self.arena[off:off + nbytes].copy_(other.arena[off:off + nbytes], non_blocking=True)
        n_boxes = int(self.actual[TOT_BOXES])
        if self.boxes_u16 is not None and n_boxes:
            self.boxes[:n_boxes].copy_(other.boxes[:n_boxes], non_blocking=True)
        self._windows()
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

    @property
    def actual_pairs(self) -> int:
        return int(self.actual[TOT_PAIRS])

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


# ---------------------------------------------------------------------------------------------------
# ragged videos -> batches of a few capacities
# ---------------------------------------------------------------------------------------------------
T_CLASSES = (512, 1024, 2048)          # the pair kernel's chunk shapes (tspn_geo_chunk); longer videos: chunk 2048


def t_class(t: int) -> int:
    """The pair kernel's chunk for a video of ``t`` frames: videos are batched with others of their class,
    because a CTA stages whole chunks of box rows - a 300-frame video in a 2048-frame chunk would read 7x the
    bytes it uses."""
    for c in T_CLASSES:
        if t <= c:
            return c
    return T_CLASSES[-1]


def pack_batches(shapes: Sequence[Tuple[int, int]], geo_budget_bytes: int = 4 << 30, max_videos: int = 64,
                 merge_below_bytes: int = 512 << 20) -> List[Tuple[int, List[int]]]:
    """Split ragged videos ``(N_i, T_i)`` into batches: ``[(t_class, [video indices]), ...]``.

    Videos are grouped by chunk class and, inside a class, taken in descending cost order (first-fit into the
    open batch) until the batch's geometry output (32 B per pair-frame) would exceed ``geo_budget_bytes`` or it
    holds ``max_videos`` videos.  A video larger than the budget gets a batch of its own.  A class whose videos
    add up to less than ``merge_below_bytes`` of geometry is merged into the next larger class: a batch has a fixed
    cost (its side chain is latency-bound, ~0.5 ms however small the batch), which outweighs staging short videos in
    a longer chunk when there is little of them - the case of a rank's shard in a many-GPU run.  Deterministic."""
    def geo_bytes(i):
        n, t = shapes[i]
        return int(n) * max(int(n) - 1, 0) * ((int(t) + 3) // 4 * 4) * 32
    by_class = {}
    for i, (n, t) in enumerate(shapes):
        by_class.setdefault(t_class(int(t)), []).append(i)
    classes = sorted(by_class)
    for k, c in enumerate(classes[:-1]):
        if by_class[c] and sum(geo_bytes(i) for i in by_class[c]) < merge_below_bytes:
            by_class[classes[k + 1]].extend(by_class[c])
            by_class[c] = []
    out: List[Tuple[int, List[int]]] = []
    for c in classes:
        order = sorted(by_class[c], key=lambda i: (-geo_bytes(i), i))
        cur, cur_bytes = [], 0
        for i in order:
            b = geo_bytes(i)
            if cur and (cur_bytes + b > geo_budget_bytes or len(cur) >= max_videos):
                out.append((c, sorted(cur)))
                cur, cur_bytes = [], 0
            cur.append(i)
            cur_bytes += b
        if cur:
            out.append((c, sorted(cur)))
    return out


def bucket_capacities(shapes: Sequence[Tuple[int, int]], batches: Iterable[Tuple[int, List[int]]], n_classes: int,
                      **kw) -> dict:
    """One ``Capacity`` per chunk class that holds every batch of that class: ``{t_class: Capacity}``."""
    caps = {}
    for c, vids in batches:
        cap = Capacity.for_shapes([shapes[i] for i in vids], n_classes, max_t=c, **kw)   # chunk = the class's
        old = caps.get(c)
        if old is None:
            caps[c] = cap
        else:
            caps[c] = Capacity(**{f.name: (max(getattr(old, f.name), getattr(cap, f.name))
                                           if f.name not in ("n_classes", "boxes_u16", "motion_u8", "has_motion",
                                                             "geo_chunk")
                                           else getattr(old, f.name)) for f in dataclasses.fields(Capacity)})
    return caps
