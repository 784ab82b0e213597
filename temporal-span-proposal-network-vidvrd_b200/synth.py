"""Seeded synthetic tracklet inputs for the pair stage.

There is no dataset in the build container or on the GPU box, so every test and
bench line runs on inputs made here.  The generator is numpy-only and uses
``numpy.random.Generator(PCG64(seed))`` whose stream is stable across numpy
versions, so a (seed, shape) pair names the same tensors on every machine.

Shapes follow the reference's data contract:

* boxes are inclusive-pixel ``(x1, y1, x2, y2)`` rows as in the reference README
  (``/root/reference/README.md:55-60``), rounded to integers and stored as f32;
* a tracklet lives on ``[pstart, pend)`` (``lib/modeling/trajectory.py:21-22``),
  rows outside the span are zero;
* ``cls`` plays the role of ``track_cls_logits`` / classeme
  (``lib/dataset/vrdataset.py:61-83``);
* ``motion`` is the per-tracklet 4x1000 bag-of-words block that
  ``lib/dataset/vrdataset.py:219-243`` L1-normalises.
"""
from __future__ import annotations

import dataclasses
from typing import List, Sequence

import numpy as np

FRAME_W = 1920
FRAME_H = 1080
MOTION_DIM = 4000          # 4 BoW blocks of 1000 (vrdataset.py:227-236)
MOTION_BLOCK = 1000
REL_DIM = 3000             # relative position + size + motion (vrdataset.py:238-241)


@dataclasses.dataclass
class VideoTracklets:
    """One video's tracklets, dense over T frames."""
    boxes: np.ndarray    # [N, T, 4] f32, zero outside span
    span: np.ndarray     # [N, 2]   i32, [pstart, pend)
    cls: np.ndarray      # [N, C]   f32
    motion: np.ndarray   # [N, 4000] f32, non-negative counts
    seed: int = 0

    @property
    def n_tracklets(self) -> int:
        return int(self.boxes.shape[0])

    @property
    def n_frames(self) -> int:
        return int(self.boxes.shape[1])

    @property
    def n_pairs(self) -> int:
        n = self.n_tracklets
        return n * (n - 1)


def make_video(n_tracklets: int, n_frames: int, n_classes: int, seed: int = 0,
               full_span: bool = False, integer_boxes: bool = True,
               spread: float = 0.35) -> VideoTracklets:
    """Smooth random-walk tracklets in a 1920x1080 frame.

    ``spread`` is the fraction of the frame (around its centre) in which the walks
    start: 1.0 scatters the tracklets over the whole frame (almost no pair ever
    intersects), the default 0.35 gives the interacting-objects regime in which the
    intersection / vIoU arithmetic is actually exercised.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n, t = int(n_tracklets), int(n_frames)
    if full_span or t < 4:
        pstart = np.zeros(n, dtype=np.int64)
        pend = np.full(n, t, dtype=np.int64)
    else:
        pstart = rng.integers(0, t // 2 + 1, size=n)
        length = rng.integers(max(t // 4, 1), t + 1, size=n)
        pend = np.minimum(pstart + length, t)
    # centre random walk, size jitter walk
    cx0 = rng.uniform(0.5 * (1 - spread) * FRAME_W, 0.5 * (1 + spread) * FRAME_W, size=(n, 1))
    cy0 = rng.uniform(0.5 * (1 - spread) * FRAME_H, 0.5 * (1 + spread) * FRAME_H, size=(n, 1))
    cx = cx0 + np.cumsum(rng.normal(0.0, 4.0, size=(n, t)), axis=1)
    cy = cy0 + np.cumsum(rng.normal(0.0, 4.0, size=(n, t)), axis=1)
    w0 = rng.integers(16, 321, size=(n, 1)).astype(np.float64)
    h0 = rng.integers(16, 321, size=(n, 1)).astype(np.float64)
    w = np.clip(w0 + np.cumsum(rng.integers(-1, 2, size=(n, t)), axis=1), 8, 640)
    h = np.clip(h0 + np.cumsum(rng.integers(-1, 2, size=(n, t)), axis=1), 8, 640)
    x1 = np.clip(cx - 0.5 * w, 0, FRAME_W - 9)
    y1 = np.clip(cy - 0.5 * h, 0, FRAME_H - 9)
    x2 = np.minimum(x1 + w - 1, FRAME_W - 1)
    y2 = np.minimum(y1 + h - 1, FRAME_H - 1)
    boxes = np.stack([x1, y1, x2, y2], axis=-1)
    if integer_boxes:
        boxes = np.rint(boxes)
    # keep x2 >= x1, y2 >= y1 after rounding
    boxes[..., 2] = np.maximum(boxes[..., 2], boxes[..., 0])
    boxes[..., 3] = np.maximum(boxes[..., 3], boxes[..., 1])
    frame = np.arange(t)[None, :]
    alive = (frame >= pstart[:, None]) & (frame < pend[:, None])
    boxes = np.where(alive[..., None], boxes, 0.0).astype(np.float32)
    logits = rng.normal(0.0, 1.0, size=(n, n_classes)) * 3.0
    logits -= logits.max(axis=1, keepdims=True)
    e = np.exp(logits)
    cls = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
    motion = rng.poisson(0.05, size=(n, MOTION_DIM)).astype(np.float32)
    span = np.stack([pstart, pend], axis=1).astype(np.int32)
    return VideoTracklets(boxes=boxes, span=span, cls=cls, motion=motion, seed=seed)


# ---------------------------------------------------------------------------
# The five BASELINE.json configurations.
# ---------------------------------------------------------------------------
CONFIGS = {
    # name: (n_videos, N spec, T spec, C, R, K)
    "vidvrd_single": dict(videos=1, n=(20, 20), t=(300, 300), classes=35, predicates=132, topk=256),
    "vidvrd_test": dict(videos=200, n=(2, 40), t=(90, 1200), classes=35, predicates=132, topk=256),
    "vidor_single": dict(videos=1, n=(64, 64), t=(2000, 2000), classes=80, predicates=50, topk=256),
    "vidor_val": dict(videos=835, n=(2, 64), t=(300, 2000), classes=80, predicates=50, topk=256),
    "stress": dict(videos=1, n=(256, 256), t=(4096, 4096), classes=80, predicates=50, topk=1024),
}


def config_shapes(name: str, seed: int = 0, videos: int | None = None) -> List[tuple]:
    """(N, T) per video of a named configuration, drawn from one PCG64 stream."""
    spec = CONFIGS[name]
    v = spec["videos"] if videos is None else int(videos)
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    ns = rng.integers(spec["n"][0], spec["n"][1] + 1, size=v)
    ts = rng.integers(spec["t"][0], spec["t"][1] + 1, size=v)
    return [(int(a), int(b)) for a, b in zip(ns, ts)]


def make_config(name: str, seed: int = 0, videos: int | None = None,
                full_span: bool = False) -> List[VideoTracklets]:
    """All videos of a named configuration; video i uses seed ``seed + i``."""
    spec = CONFIGS[name]
    shapes = config_shapes(name, seed, videos)
    return [make_video(n, t, spec["classes"], seed=seed + i, full_span=full_span)
            for i, (n, t) in enumerate(shapes)]


def make_weights(n_classes: int, n_predicates: int, feature_dim: int, hidden: int = 64,
                 dpn_in: int = 8, n_anchors: int = 4, seed: int = 0,
                 ppn_gain: float = 4.0) -> dict:
    """Random-init weights under the reference's state_dict keys.

    The reference initialises the classifier and both DPN convolutions with
    N(0, 0.01) / zero bias (``lib/modeling/model.py:81-83``,
    ``lib/modeling/relpn/dpn.py:65-67``) and leaves ``PPNHead`` on torch's
    default ``Linear`` init (``lib/modeling/relpn/ppn.py:92-105``).  We draw the
    same distributions from PCG64 so the tensors are reproducible without torch's
    RNG; ``ppn_gain`` widens the PPN weights so that top-K margins are far above
    one ulp ('bit-exact top-K').
    """
    rng = np.random.Generator(np.random.PCG64(seed + 104729))
    c, r, f = int(n_classes), int(n_predicates), int(feature_dim)

    def lin(out_f, in_f, gain=1.0):
        bound = gain / np.sqrt(in_f)
        return (rng.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32),
                rng.uniform(-bound, bound, size=(out_f,)).astype(np.float32))

    sd = {}
    for br in ("sub_emb", "obj_emb"):
        w0, b0 = lin(hidden, c, ppn_gain)
        w2, b2 = lin(c, hidden, ppn_gain)
        p = "relpn.pair_proposal_network.ppn_head.%s." % br
        sd[p + "0.weight"], sd[p + "0.bias"] = w0, b0
        sd[p + "2.weight"], sd[p + "2.bias"] = w2, b2
    p = "relpn.duration_proposal_network.dpn_head."
    sd[p + "conv.weight"] = rng.normal(0, 0.01, size=(dpn_in, dpn_in, 3)).astype(np.float32)
    sd[p + "conv.bias"] = np.zeros(dpn_in, dtype=np.float32)
    sd[p + "duration_pred.weight"] = rng.normal(0, 0.01, size=(2 * n_anchors, dpn_in, 1)).astype(np.float32)
    sd[p + "duration_pred.bias"] = np.zeros(2 * n_anchors, dtype=np.float32)
    sd["classifier.rel_predictor.weight"] = rng.normal(0, 0.01, size=(r, f)).astype(np.float32)
    sd["classifier.rel_predictor.bias"] = np.zeros(r, dtype=np.float32)
    return sd


def feature_dim(n_classes: int) -> int:
    """2C classeme + 2x4000 motion BoW + 3000 relative block (vrdataset.py:219-243)."""
    return 2 * int(n_classes) + 2 * MOTION_DIM + REL_DIM
