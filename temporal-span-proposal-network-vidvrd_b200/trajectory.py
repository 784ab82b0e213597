"""Trajectory geometry — mirror of lib/modeling/trajectory.py:12-158, lib/evaluation/common.py:65-106
and lib/modeling/association.py:35-48 on the CUDA path.

``cubic_iou`` / ``traj_iou`` / ``viou`` keep the reference signatures and return the reference's
types (numpy float32 matrix / python float); arrays may also be torch tensors (CPU or CUDA), in
which case a tensor on the same device comes back.  Sums are accumulated in fp64 on the GPU, so
results agree with the reference's float32 sequential accumulation to ~1e-6 relative and do not
depend on summation order.
"""
from __future__ import annotations

from collections import deque
from typing import Sequence

import numpy as np
import torch

from . import ops


class Trajectory:
    """Bounding-box trajectory on ``[pstart, pend)`` (trajectory.py:12-82) without the dlib type:
    ``rois`` holds ``(left, top, right, bottom)`` tuples."""

    def __init__(self, pstart, pend, rois, score, category, classeme, vsig=None, gt_trackid=-1):
        assert len(rois) == pend - pstart
        self.pstart, self.pend = pstart, pend
        self.rois = deque(tuple(float(c) for c in roi) for roi in rois)
        self.score, self.category, self.classeme = score, category, classeme
        self.vsig, self.gt_trackid = vsig, gt_trackid
        self._version = 0            # bumped by every in-place edit (association caches IoUs per version)

    def __lt__(self, other):
        return self.score < other.score

    def head(self):
        return self.rois[0]

    def tail(self):
        return self.rois[-1]

    def at(self, i):
        return self.rois[i]

    def roi_at(self, p):
        return self.rois[p - self.pstart]

    def bbox_at(self, p):
        l, t, r, b = self.rois[p - self.pstart]
        return (l, t, r - l + 1, b - t + 1)

    def length(self):
        return self.pend - self.pstart

    def predict(self, roi, reverse=False):
        if reverse:
            self.rois.appendleft(tuple(roi))
            self.pstart -= 1
        else:
            self.rois.append(tuple(roi))
            self.pend += 1
        self._version += 1
        return roi

    def serialize(self):
        return dict(pstart=int(self.pstart), pend=int(self.pend), rois=[tuple(r) for r in self.rois],
                    score=float(self.score), category=int(self.category),
                    classeme=[float(x) for x in self.classeme] if self.classeme is not None else None,
                    vsig=self.vsig, gt_trackid=self.gt_trackid)


def _dev():
    ops.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def cubic_iou(bboxes1, bboxes2):
    """``[n, t, 4] x [m, t, 4] -> [n, m]`` float32 (trajectory.py:127-141); float input required
    like the reference (quirk Q6)."""
    is_t = isinstance(bboxes1, torch.Tensor)
    a = bboxes1 if is_t else torch.from_numpy(np.ascontiguousarray(bboxes1))
    b = a if bboxes2 is bboxes1 else (bboxes2 if isinstance(bboxes2, torch.Tensor)
                                      else torch.from_numpy(np.ascontiguousarray(bboxes2)))
    if not (a.is_floating_point() and b.is_floating_point()):
        raise TypeError("cubic_iou needs floating-point boxes")
    dev = a.device if a.is_cuda else _dev()
    out = ops.cubic_iou(a.to(dev, torch.float32), b.to(dev, torch.float32))
    if is_t:
        return out if a.is_cuda else out.cpu()
    return out.cpu().numpy()


def traj_iou(trajs1: Sequence, trajs2: Sequence):
    """Pairwise trajectory IoU of equal-span trajectories (trajectory.py:144-158)."""
    def arr(trajs):
        return np.asarray([[tuple(r) if not hasattr(r, "left") else (r.left(), r.top(), r.right(), r.bottom())
                            for r in t.rois] for t in trajs], dtype=np.float32)
    b1 = arr(trajs1)
    b2 = b1 if trajs1 is trajs2 else arr(trajs2)
    return cubic_iou(b1, b2)


def viou_batch(trajs, durations, pairs, clipped: bool = False, f64: bool = False) -> np.ndarray:
    """vIoU of many trajectory pairs in one launch (the O(#pred x #gt) loop of
    lib/evaluation/visual_relation_detection.py:8-36).  ``trajs[j]`` is a list/array of boxes on
    ``durations[j] = (fstart, fend)``; ``pairs`` is ``[M, 2]`` indices into ``trajs``.  ``f64`` keeps the
    sums and the ratio in fp64 (``tspn_viou_pairs_f64``: volumes once per trajectory) — for integer boxes
    the values are then bit-identical to the reference's python arithmetic."""
    dev = _dev()
    lens = [len(t) for t in trajs]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pool = np.zeros((max(int(off[-1]), 1), 4), dtype=np.float32)
    for j, t in enumerate(trajs):
        if lens[j]:
            pool[off[j]:off[j + 1]] = np.asarray(t, dtype=np.float32).reshape(-1, 4)
    span = np.asarray(durations, dtype=np.int32).reshape(-1, 2)
    dur = span[:, 1] - span[:, 0]
    lens_a = np.asarray(lens, dtype=np.int32)
    longer = bool((lens_a != dur).any())
    if longer and (not f64 or clipped or (lens_a < dur).any()):
        # common.py:100-105 sums volumes over the whole box list, so a list LONGER than its duration is
        # meaningful (association emits such relations) - supported by the fp64 entry point only; a
        # shorter list makes the reference index out of range
        j = int(np.nonzero(lens_a != dur)[0][0])
        raise ValueError("trajectory %d: %d boxes for duration %s" % (j, lens[j], tuple(span[j])))
    pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    if pairs.shape[0] == 0:
        return np.zeros(0, dtype=np.float64 if f64 else np.float32)
    args = (torch.from_numpy(pool).to(dev), torch.from_numpy(off[:-1].copy()).to(dev),
            torch.from_numpy(span).to(dev), torch.from_numpy(pairs[:, 0].copy()).to(dev),
            torch.from_numpy(pairs[:, 1].copy()).to(dev))
    if f64:
        out = ops.viou_pairs_f64(*args, clipped=clipped,
                                 traj_len=torch.from_numpy(lens_a).to(dev) if longer else None)
    else:
        out = ops.viou_pairs(*args, clipped=clipped)
    return out.cpu().numpy()


def viou(traj_1, duration_1, traj_2, duration_2) -> float:
    """Voluminal IoU of two trajectories with durations (evaluation/common.py:65-106).  Goes through the fp64
    entry point: it accepts box lists longer than their duration (common.py:100-105 sums the volume over the whole
    list, and association emits such relations) and returns python's float arithmetic bit for bit on integer boxes."""
    return float(viou_batch([traj_1, traj_2], [duration_1, duration_2], [(0, 1)], f64=True)[0])


def _traj_iou(traj_1: Trajectory, traj_2: Trajectory) -> float:
    """Overlap-clipped trajectory IoU (association.py:35-48)."""
    if traj_1.pend <= traj_2.pstart or traj_2.pend <= traj_1.pstart:
        return 0
    return float(viou_batch([list(traj_1.rois), list(traj_2.rois)],
                            [(traj_1.pstart, traj_1.pend), (traj_2.pstart, traj_2.pend)], [(0, 1)],
                            clipped=True)[0])
