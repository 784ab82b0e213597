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


# ---------------------------------------------------------------------------
# Relation-level synthetic data for the rows after the pair stage:
# evaluation (N3) and greedy relational association (N2).
# ---------------------------------------------------------------------------
def _walk_boxes(rng, length: int, integer: bool = True) -> np.ndarray:
    """[length, 4] inclusive-pixel boxes of one smooth random walk."""
    cx = rng.uniform(300, FRAME_W - 300) + np.cumsum(rng.normal(0.0, 3.0, size=length))
    cy = rng.uniform(200, FRAME_H - 200) + np.cumsum(rng.normal(0.0, 3.0, size=length))
    w = np.clip(rng.integers(40, 300) + np.cumsum(rng.integers(-1, 2, size=length)), 16, 600)
    h = np.clip(rng.integers(40, 300) + np.cumsum(rng.integers(-1, 2, size=length)), 16, 600)
    x1 = np.clip(cx - 0.5 * w, 0, FRAME_W - 17)
    y1 = np.clip(cy - 0.5 * h, 0, FRAME_H - 17)
    b = np.stack([x1, y1, np.minimum(x1 + w - 1, FRAME_W - 1), np.minimum(y1 + h - 1, FRAME_H - 1)], axis=-1)
    return np.rint(b) if integer else b


def _jitter(rng, boxes: np.ndarray, amount: float, integer: bool = True) -> np.ndarray:
    """A detector's view of a ground-truth trajectory: offset + per-frame noise proportional to size."""
    w = boxes[:, 2] - boxes[:, 0] + 1
    h = boxes[:, 3] - boxes[:, 1] + 1
    off = rng.normal(0.0, amount, size=4)
    noise = rng.normal(0.0, 0.25 * amount, size=boxes.shape)
    scale = np.stack([w, h, w, h], axis=-1)
    out = boxes + (off[None, :] + noise) * scale
    out[:, 2] = np.maximum(out[:, 2], out[:, 0] + 4)
    out[:, 3] = np.maximum(out[:, 3], out[:, 1] + 4)
    return np.rint(out) if integer else out


def _as_lists(boxes: np.ndarray, integer: bool):
    return [[int(c) for c in r] for r in boxes] if integer else [[float(c) for c in r] for r in boxes]


def make_relation_eval_case(seed: int = 0, n_videos: int = 6, max_gt: int = 12, max_pred: int = 60,
                            max_frames: int = 240, n_objects: int = 5, n_predicates: int = 4,
                            integer_boxes: bool = True):
    """``(groundtruth, prediction)`` dicts in the JSON layout lib/evaluation/README.md describes and
    ``evaluate`` (lib/evaluation/visual_relation_detection.py:64) consumes: per video a list of relations
    ``{triplet, duration [fstart, fend), sub_traj, obj_traj}`` (+ ``score`` for predictions).  Predictions
    are jittered / time-shifted copies of ground truth (so vIoU spreads around the 0.5 threshold),
    wrong-triplet copies and unrelated walks; one video has no ground truth and one no predictions."""
    rng = np.random.Generator(np.random.PCG64(seed + 15485863))
    gt, pred = {}, {}
    for v in range(n_videos):
        vid = "video_%03d" % v
        n_gt = 0 if v == n_videos - 1 else int(rng.integers(1, max_gt + 1))
        gts = []
        for _ in range(n_gt):
            length = int(rng.integers(8, max_frames + 1))
            fstart = int(rng.integers(0, max_frames))
            sub, obj = _walk_boxes(rng, length, integer_boxes), _walk_boxes(rng, length, integer_boxes)
            trip = ["obj%d" % rng.integers(n_objects), "pred%d" % rng.integers(n_predicates),
                    "obj%d" % rng.integers(n_objects)]
            gts.append({"triplet": trip, "duration": [fstart, fstart + length], "_sub": sub, "_obj": obj})
        preds = []
        n_pred = 0 if v == n_videos - 2 else int(rng.integers(1, max_pred + 1))
        for _ in range(n_pred):
            kind = rng.random()
            if gts and kind < 0.75:
                g = gts[int(rng.integers(len(gts)))]
                f0, f1 = g["duration"]
                glen = f1 - f0
                cut0 = int(rng.integers(0, max(glen // 3, 1)))          # trim / extend in time
                cut1 = int(rng.integers(0, max(glen // 3, 1)))
                lo, hi = cut0, max(glen - cut1, cut0 + 1)
                amount = float(rng.choice([0.01, 0.03, 0.06, 0.12]))
                sub = _jitter(rng, g["_sub"][lo:hi], amount, integer_boxes)
                obj = _jitter(rng, g["_obj"][lo:hi], amount, integer_boxes)
                ext = int(rng.integers(0, 6))                            # a few frames past the ground truth
                if ext:
                    sub = np.concatenate([sub, np.repeat(sub[-1:], ext, axis=0)])
                    obj = np.concatenate([obj, np.repeat(obj[-1:], ext, axis=0)])
                trip = list(g["triplet"])
                if kind > 0.65:
                    trip[1] = "pred%d" % rng.integers(n_predicates)
                dur = [f0 + lo, f0 + hi + ext]
            else:
                length = int(rng.integers(8, max_frames + 1))
                fstart = int(rng.integers(0, max_frames))
                sub, obj = _walk_boxes(rng, length, integer_boxes), _walk_boxes(rng, length, integer_boxes)
                trip = ["obj%d" % rng.integers(n_objects), "pred%d" % rng.integers(n_predicates),
                        "obj%d" % rng.integers(n_objects)]
                dur = [fstart, fstart + length]
            # a coarse score grid produces equal scores: the stable sort order matters
            preds.append({"triplet": trip, "score": float(np.round(rng.random(), 2)), "duration": dur,
                          "sub_traj": _as_lists(sub, integer_boxes), "obj_traj": _as_lists(obj, integer_boxes)})
        gt[vid] = [{"triplet": g["triplet"], "duration": g["duration"],
                    "sub_traj": _as_lists(g["_sub"], integer_boxes), "obj_traj": _as_lists(g["_obj"], integer_boxes)}
                   for g in gts]
        pred[vid] = preds
    return gt, pred


def make_association_case(seed: int = 0, n_segments: int = 8, n_objects: int = 5, n_classes: int = 35,
                          n_predicates: int = 132, preds_per_segment: int = 40, seg_len: int = 30,
                          seg_stride: int = 15, vid: str = "synthetic_video"):
    """Input of ``greedy_relational_association`` (lib/modeling/association.py:117-175) for one video:

    * ``short_term_relations``: list of ``((vid, fstart, fend), (pred_list, iou, trackid))`` with
      ``pred_list[i] = (score array, triplet array[3], (s_tid, o_tid) array)`` as predict.py:110-117 builds;
    * ``segment_trajs``: ``{(vid, fstart, fend): [trajectory kwargs]}`` - what
      ``object_trajectory_proposal`` (lib/modeling/trajectory.py:161-180) would load from disk.

    Objects persist over 30-frame segments with 15-frame overlap (lib/modeling/__init__.py:36-42); each
    segment re-detects them with jitter, so the overlap-clipped vIoU of consecutive detections spreads
    around the 0.5 merge threshold; tracklet order is shuffled per segment."""
    rng = np.random.Generator(np.random.PCG64(seed + 32452843))
    total = seg_stride * (n_segments - 1) + seg_len
    walks = [_walk_boxes(rng, total, integer=False) for _ in range(n_objects)]
    cats = rng.integers(0, n_classes, size=n_objects)
    rel_pool = [(int(s), int(rng.integers(n_predicates)), int(o)) for s in range(n_objects) for o in range(n_objects)
                if s != o]
    rel_pool = [rel_pool[i] for i in rng.permutation(len(rel_pool))[:max(4, len(rel_pool) // 2)]]
    short_term, segment_trajs = [], {}
    for k in range(n_segments):
        f0 = k * seg_stride
        f1 = f0 + seg_len
        order = rng.permutation(n_objects)
        present = [int(o) for o in order if rng.random() < 0.9]
        trajs = []
        for o in present:
            amount = float(rng.choice([0.0, 0.02, 0.05, 0.15]))
            rois = _jitter(rng, walks[o][f0:f1], amount, integer=False) if amount else walks[o][f0:f1]
            trajs.append(dict(pstart=f0, pend=f1, rois=[tuple(float(c) for c in r) for r in rois],
                              score=float(rng.random()), category=int(cats[o]), classeme=[], vsig=None,
                              gt_trackid=-1))
        slot = {o: i for i, o in enumerate(present)}
        plist = []
        for _ in range(preds_per_segment):
            s, p, o = rel_pool[int(rng.integers(len(rel_pool)))]
            if s not in slot or o not in slot:
                continue
            if rng.random() < 0.15:
                p = int(rng.integers(n_predicates))
            plist.append((np.array(np.float32(np.round(rng.random(), 2))),
                          np.array([int(cats[s]), p, int(cats[o])]), np.array([slot[s], slot[o]])))
        n = len(trajs)
        index = (vid, f0, f1)
        short_term.append((index, (plist, np.eye(n, dtype=np.float32), np.full(n, -1))))
        segment_trajs[index] = trajs
    order = rng.permutation(len(short_term))          # the association sorts segments by fstart itself
    return [short_term[i] for i in order], segment_trajs


def make_gt_from_relations(relations, seed: int = 0, keep_every: int = 3):
    """Ground truth derived from serialized video relations (the output of greedy association): every
    ``keep_every``-th relation, its box lists cut to the duration and shifted by 0-3 pixels.  Used to
    evaluate association output, whose box lists can be LONGER than their durations (shared trajectories
    extended by other relations' merges) - the case lib/evaluation/common.py:100-105 sums volumes over."""
    rng = np.random.Generator(np.random.PCG64(seed + 49979687))
    out = []
    for i, r in enumerate(relations):
        if i % keep_every:
            continue
        n = r["duration"][1] - r["duration"][0]
        shift = float(rng.integers(0, 4))
        out.append({"triplet": list(r["triplet"]), "duration": list(r["duration"]),
                    "sub_traj": [[float(c) + shift for c in b] for b in r["sub_traj"][:n]],
                    "obj_traj": [[float(c) - shift for c in b] for b in r["obj_traj"][:n]]})
    return out
