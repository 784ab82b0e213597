"""Oracle: pair enumeration and trajectory geometry.  TEST INFRASTRUCTURE ONLY.

Reference-faithful ports (same operation order and dtypes as the reference, so
they are compared bit-for-bit with reference outputs in tests/golden):

* ``intersect_ref`` / ``union_ref`` / ``cubic_iou_ref``  lib/modeling/trajectory.py:85-141
* ``viou_ref``                                             lib/evaluation/common.py:65-106
* ``traj_iou_clipped_ref``                                 lib/modeling/association.py:35-48

Float64 definitions used as the parity target of the CUDA kernels:

* ``cubic_iou_f64``, ``pair_geometry`` ([SPEC] s2/s3 of SURVEY.md section 8a)
"""
from __future__ import annotations

import numpy as np

GEO_CHANNELS = 8


# ---------------------------------------------------------------------------
# a1: pair enumeration
# ---------------------------------------------------------------------------
def enumerate_pairs(n: int) -> np.ndarray:
    """All ordered pairs (s, o), s != o, subject-major.

    The reference stores this table in the per-segment h5 file (``pairs``,
    lib/dataset/vrdataset.py:208) in ``itertools.permutations(range(N), 2)`` order
    (sample shown at lib/modeling/predict.py:133-140); row ``p`` of pair ``(s, o)``
    is ``s*(N-1) + o - [o > s]``.
    """
    n = int(n)
    if n < 2:
        return np.zeros((0, 2), dtype=np.int64)
    s = np.repeat(np.arange(n, dtype=np.int64), n - 1)
    k = np.tile(np.arange(n - 1, dtype=np.int64), n)
    o = k + (k >= s)
    return np.stack([s, o], axis=1)


def pair_row(s: int, o: int, n: int) -> int:
    return s * (n - 1) + o - (1 if o > s else 0)


# ---------------------------------------------------------------------------
# a3 / a4: cubic IoU, reference-faithful (float32 sequential accumulation)
# ---------------------------------------------------------------------------
def intersect_ref(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    """Sum over frames of the inclusive-pixel box intersection, [n1, n2] f32.

    ``b1``/``b2`` are frame-major ``[t, n, 4]`` as at trajectory.py:87.  The
    reference accumulates one frame at a time in float32 (trajectory.py:91-106):
    ``w = (min(x2) + 1) - max(x1)`` clipped at 0, same for ``h``, ``inters += w*h``.
    """
    assert b1.shape[0] == b2.shape[0]
    acc = np.zeros((b1.shape[1], b2.shape[1]), dtype=np.float32)
    one = np.float32(1)
    for f in range(b1.shape[0]):
        lo_x = np.maximum(b1[f, :, None, 0], b2[f, None, :, 0]).astype(np.float32)
        hi_x = np.minimum(b1[f, :, None, 2], b2[f, None, :, 2]).astype(np.float32)
        ww = np.maximum((hi_x + one) - lo_x, np.float32(0))
        lo_y = np.maximum(b1[f, :, None, 1], b2[f, None, :, 1]).astype(np.float32)
        hi_y = np.minimum(b1[f, :, None, 3], b2[f, None, :, 3]).astype(np.float32)
        hh = np.maximum((hi_y + one) - lo_y, np.float32(0))
        acc += ww * hh
    return acc


def _volume_ref(b: np.ndarray) -> np.ndarray:
    # trajectory.py:112-114 — (x2 - x1 + 1) * (y2 - y1 + 1) summed over frames with
    # np.sum in the input dtype.
    w = b[:, :, 2] - b[:, :, 0] + 1
    h = b[:, :, 3] - b[:, :, 1] + 1
    return np.sum(w * h, axis=0)


def union_ref(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    """Outer sum of the per-trajectory volumes (trajectory.py:110-124)."""
    v1 = _volume_ref(b1)
    v2 = v1 if b1 is b2 else _volume_ref(b2)
    return np.add.outer(v1, v2)


def cubic_iou_ref(bboxes1: np.ndarray, bboxes2: np.ndarray) -> np.ndarray:
    """``[n, t, 4] x [m, t, 4] -> [n, m]`` cubic IoU, trajectory.py:127-141.

    All trajectories share one span (V1).  Like the reference, integer inputs are
    not supported (in-place true-divide into the float32 intersection buffer).
    """
    same = bboxes1 is bboxes2
    f1 = np.transpose(bboxes1, (1, 0, 2))
    f2 = f1 if same else np.transpose(bboxes2, (1, 0, 2))
    inter = intersect_ref(f1, f2)
    uni = union_ref(f1, f2)
    if not np.issubdtype(uni.dtype, np.floating):
        raise TypeError("cubic_iou needs floating-point boxes (reference quirk Q6)")
    # trajectory.py:138-139: union -= inter (in union's dtype: f32, or f64 when called
    # through traj_iou), then inter = inter / union written back into the f32 buffer.
    uni = uni - inter
    return (inter / uni).astype(np.float32)


def cubic_iou_f64(bboxes1: np.ndarray, bboxes2: np.ndarray) -> np.ndarray:
    """Float64 cubic IoU (order-free sums); the CUDA matrix kernel's parity target."""
    a = np.asarray(bboxes1, dtype=np.float64)
    b = np.asarray(bboxes2, dtype=np.float64)
    n, t, _ = a.shape
    m = b.shape[0]
    inter = np.zeros((n, m))
    for f in range(t):
        iw = np.minimum(a[:, None, f, 2], b[None, :, f, 2]) - np.maximum(a[:, None, f, 0], b[None, :, f, 0]) + 1
        ih = np.minimum(a[:, None, f, 3], b[None, :, f, 3]) - np.maximum(a[:, None, f, 1], b[None, :, f, 1]) + 1
        inter += np.maximum(iw, 0) * np.maximum(ih, 0)
    va = ((a[..., 2] - a[..., 0] + 1) * (a[..., 3] - a[..., 1] + 1)).sum(axis=1)
    vb = ((b[..., 2] - b[..., 0] + 1) * (b[..., 3] - b[..., 1] + 1)).sum(axis=1)
    return inter / (va[:, None] + vb[None, :] - inter)


# ---------------------------------------------------------------------------
# a6: vIoU with different durations (V2), evaluation/common.py:65-106
# ---------------------------------------------------------------------------
def viou_ref(traj_1, duration_1, traj_2, duration_2) -> float:
    """Voluminal IoU of two box lists living on ``[fstart, fend)`` durations.

    Restates evaluation/common.py:65-106: the intersection runs over the temporal
    overlap window ``[max(s1,s2), min(e1,e2))`` (the head/tail offset cases at
    :71-90 all reduce to that window), each volume over the trajectory's *full*
    list, 0.0 for disjoint durations.  Python numbers, exact for integer boxes.
    """
    s1, e1 = int(duration_1[0]), int(duration_1[1])
    s2, e2 = int(duration_2[0]), int(duration_2[1])
    if s1 >= e2 or e1 <= s2:
        return 0.0
    lo, hi = max(s1, s2), min(e1, e2)
    inter = 0
    for f in range(lo, hi):
        r1 = traj_1[f - s1]
        r2 = traj_2[f - s2]
        iw = min(r1[2], r2[2]) - max(r1[0], r2[0]) + 1
        ih = min(r1[3], r2[3]) - max(r1[1], r2[1]) + 1
        inter += max(0, iw) * max(0, ih)
    v1 = 0
    for r in traj_1:
        v1 += (r[2] - r[0] + 1) * (r[3] - r[1] + 1)
    v2 = 0
    for r in traj_2:
        v2 += (r[2] - r[0] + 1) * (r[3] - r[1] + 1)
    return float(inter) / (v1 + v2 - inter)


# ---------------------------------------------------------------------------
# a7: overlap-clipped vIoU (V3), association.py:35-48
# ---------------------------------------------------------------------------
def traj_iou_clipped_ref(rois_1: np.ndarray, span_1, rois_2: np.ndarray, span_2) -> float:
    """Clip both trajectories to their temporal overlap, then cubic IoU (V1).

    association.py:35-48: the earlier-starting trajectory is cut to
    ``[t2.pstart, t1.pend)`` and the later one to its first ``t1.pend - t2.pstart``
    boxes — i.e. the reference assumes the earlier-starting trajectory also ends
    first.  We keep that window (``min`` of the ends is identical under the
    reference's assumption) and feed float64 boxes to the float32-accumulating
    ``cubic_iou`` exactly as ``traj_iou`` does (trajectory.py:150-156).
    """
    s1, e1 = int(span_1[0]), int(span_1[1])
    s2, e2 = int(span_2[0]), int(span_2[1])
    if e1 <= s2 or e2 <= s1:
        return 0.0
    if s1 > s2:
        rois_1, rois_2 = rois_2, rois_1
        s1, e1, s2, e2 = s2, e2, s1, e1
    a = np.asarray(rois_1, dtype=np.float64)[s2 - s1:e1 - s1]
    b = np.asarray(rois_2, dtype=np.float64)[0:e1 - s2]
    return float(cubic_iou_ref(a[None], b[None])[0, 0])


# ---------------------------------------------------------------------------
# [SPEC] s2 / s3: per-frame pair geometry and per-pair reductions (float64)
# ---------------------------------------------------------------------------
def pair_geometry(boxes: np.ndarray, span: np.ndarray, s_idx: np.ndarray | None = None,
                  o_idx: np.ndarray | None = None, clip_volumes: bool = False):
    """Per-frame geometry of ordered tracklet pairs — the definition the kernels follow.

    boxes ``[N, T, 4]`` inclusive pixels, span ``[N, 2]`` = ``[pstart, pend)``
    (trajectory.py:21-22).  Returns ``geo [P, 8, T]`` f64, ``viou [P]``, ``tiou [P]``,
    ``overlap [P, 2]`` i32.  With ``w = x2-x1+1``, ``h = y2-y1+1``, ``cx = (x1+x2)/2``,
    ``cy = (y1+y2)/2`` and the overlap window ``[a, b) = [max(ps,qs), min(pe,qe))``:

    ==  =====================================================================
    0   ``(cx_s - cx_o) / w_o``
    1   ``(cy_s - cy_o) / h_o``
    2   ``log(w_s / w_o)``
    3   ``log(h_s / h_o)``
    4   per-frame IoU (the integrand of trajectory.py:96-106 over the frame union)
    5   forward difference of channel 0 (0 on the last overlap frame)
    6   forward difference of channel 1
    7   overlap mask ``1[a <= t < b]`` (the window of evaluation/common.py:71-90)
    ==  =====================================================================

    Every channel is 0 outside the overlap window.  ``viou`` follows
    evaluation/common.py:65-106 (V2: volumes over each full span) or, with
    ``clip_volumes``, association.py:35-48 (V3: volumes over the overlap only);
    with all spans equal both equal trajectory.py:127-141 (V1).  ``tiou`` =
    ``|overlap| / (len_s + len_o - |overlap|)``; ``overlap`` is ``(a, b)`` or
    ``(0, 0)`` when empty.
    """
    b = np.asarray(boxes, dtype=np.float64)
    n, t, _ = b.shape
    span = np.asarray(span, dtype=np.int64)
    if s_idx is None:
        pr = enumerate_pairs(n)
        s_idx, o_idx = pr[:, 0], pr[:, 1]
    s_idx = np.asarray(s_idx, dtype=np.int64)
    o_idx = np.asarray(o_idx, dtype=np.int64)
    p = s_idx.shape[0]
    frame = np.arange(t)[None, :]
    w = b[..., 2] - b[..., 0] + 1
    h = b[..., 3] - b[..., 1] + 1
    cx = 0.5 * (b[..., 0] + b[..., 2])
    cy = 0.5 * (b[..., 1] + b[..., 3])
    alive = (frame >= span[:, :1]) & (frame < span[:, 1:2])
    vol = np.where(alive, w * h, 0.0)

    a = np.maximum(span[s_idx, 0], span[o_idx, 0])
    e = np.minimum(span[s_idx, 1], span[o_idx, 1])
    has = e > a
    mask = has[:, None] & (frame >= a[:, None]) & (frame < e[:, None])

    geo = np.zeros((p, GEO_CHANNELS, t), dtype=np.float64)
    bs, bo = b[s_idx], b[o_idx]
    ws, wo, hs, ho = w[s_idx], w[o_idx], h[s_idx], h[o_idx]
    c0 = (cx[s_idx] - cx[o_idx]) / wo
    c1 = (cy[s_idx] - cy[o_idx]) / ho
    c2 = np.log(ws / wo)
    c3 = np.log(hs / ho)
    iw = np.minimum(bs[..., 2], bo[..., 2]) - np.maximum(bs[..., 0], bo[..., 0]) + 1
    ih = np.minimum(bs[..., 3], bo[..., 3]) - np.maximum(bs[..., 1], bo[..., 1]) + 1
    inter = np.maximum(iw, 0) * np.maximum(ih, 0)
    c4 = inter / (ws * hs + wo * ho - inter)
    nxt = np.zeros_like(mask)
    nxt[:, :-1] = mask[:, 1:]
    d0 = np.zeros_like(c0)
    d1 = np.zeros_like(c1)
    d0[:, :-1] = c0[:, 1:] - c0[:, :-1]
    d1[:, :-1] = c1[:, 1:] - c1[:, :-1]
    fwd = mask & nxt
    geo[:, 0] = np.where(mask, c0, 0)
    geo[:, 1] = np.where(mask, c1, 0)
    geo[:, 2] = np.where(mask, c2, 0)
    geo[:, 3] = np.where(mask, c3, 0)
    geo[:, 4] = np.where(mask, c4, 0)
    geo[:, 5] = np.where(fwd, d0, 0)
    geo[:, 6] = np.where(fwd, d1, 0)
    geo[:, 7] = mask

    isum = np.where(mask, inter, 0).sum(axis=1)
    if clip_volumes:
        vs = np.where(mask, ws * hs, 0).sum(axis=1)
        vo = np.where(mask, wo * ho, 0).sum(axis=1)
    else:
        vs = vol[s_idx].sum(axis=1)
        vo = vol[o_idx].sum(axis=1)
    den = vs + vo - isum
    viou = np.where(has & (den > 0), isum / np.where(den > 0, den, 1), 0.0)
    ov = np.where(has, e - a, 0)
    ls = span[s_idx, 1] - span[s_idx, 0]
    lo = span[o_idx, 1] - span[o_idx, 0]
    tden = ls + lo - ov
    tiou = np.where(has & (tden > 0), ov / np.where(tden > 0, tden, 1), 0.0)
    overlap = np.stack([np.where(has, a, 0), np.where(has, e, 0)], axis=1).astype(np.int32)
    return geo, viou, tiou, overlap


def pair_geometry_chunked(boxes, span, clip_volumes=False, max_pairs: int = 256):
    """``pair_geometry`` over all ordered pairs, evaluated ``max_pairs`` at a time."""
    n = boxes.shape[0]
    pr = enumerate_pairs(n)
    outs = []
    for i in range(0, pr.shape[0], max_pairs):
        outs.append(pair_geometry(boxes, span, pr[i:i + max_pairs, 0], pr[i:i + max_pairs, 1],
                                  clip_volumes=clip_volumes))
    if not outs:
        t = boxes.shape[1]
        return (np.zeros((0, GEO_CHANNELS, t)), np.zeros(0), np.zeros(0), np.zeros((0, 2), np.int32))
    return tuple(np.concatenate([o[k] for o in outs], axis=0) for k in range(4))


def viou_pairs_ref(args):
    """``viou_ref`` (the port of evaluation/common.py:65-106, pure Python per frame) for a list of ordered pairs of
    one video: ``args = (boxes [N, T, 4], span [N, 2], pairs [M, 2])`` -> ``[M]`` float64.  A top-level function
    of numpy-only code, so that a process pool can map it (the reference calls ``viou`` once per trajectory pair)."""
    boxes, span, pairs = args
    trajs = [[[float(c) for c in row] for row in boxes[i, int(span[i, 0]):int(span[i, 1])]] for i in range(boxes.shape[0])]
    out = np.zeros(len(pairs), dtype=np.float64)
    for j, (s, o) in enumerate(pairs):
        s, o = int(s), int(o)
        out[j] = viou_ref(trajs[s], (int(span[s, 0]), int(span[s, 1])), trajs[o], (int(span[o, 0]), int(span[o, 1])))
    return out
