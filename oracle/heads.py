"""Oracle: relationness, top-K, predicate head, span head.  TEST INFRASTRUCTURE ONLY.

The reference's heads are a handful of stock torch ops, so the faithful port calls
the same ops on CPU tensors (``*_ref``); float64 numpy versions (``*_f64``) are the
tolerance targets of the bf16 tensor-core kernels.  State-dict keys are the
reference's (``relpn.pair_proposal_network.ppn_head.*``,
``relpn.duration_proposal_network.dpn_head.*``, ``classifier.rel_predictor.*``).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

PPN = "relpn.pair_proposal_network.ppn_head."
DPNK = "relpn.duration_proposal_network.dpn_head."
CLSK = "classifier.rel_predictor."
# log(1000/16): the usual clamp on exp() of a width regression so that exp never
# overflows; [SPEC] s5.
SPAN_DW_CLAMP = math.log(1000.0 / 16.0)


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


# ---------------------------------------------------------------------------
# a8: PPNHead.forward, lib/modeling/relpn/ppn.py:92-112
# ---------------------------------------------------------------------------
def ppn_head_ref(sub_logits, obj_logits, sd) -> torch.Tensor:
    """``sigmoid(S @ O.T)`` with ``S = W2s relu(W1s x + b1s) + b2s`` (ppn.py:107-112)."""
    x, y = _t(sub_logits).float(), _t(obj_logits).float()

    def emb(v, br):
        hid = F.relu(F.linear(v, _t(sd[PPN + br + ".0.weight"]), _t(sd[PPN + br + ".0.bias"])))
        return F.linear(hid, _t(sd[PPN + br + ".2.weight"]), _t(sd[PPN + br + ".2.bias"]))

    return torch.sigmoid(torch.mm(emb(x, "sub_emb"), emb(y, "obj_emb").t()))


def ppn_head_f64(sub_logits, obj_logits, sd) -> np.ndarray:
    x = np.asarray(sub_logits, dtype=np.float64)
    y = np.asarray(obj_logits, dtype=np.float64)

    def emb(v, br):
        hid = np.maximum(v @ np.asarray(sd[PPN + br + ".0.weight"], np.float64).T
                         + np.asarray(sd[PPN + br + ".0.bias"], np.float64), 0)
        return hid @ np.asarray(sd[PPN + br + ".2.weight"], np.float64).T \
            + np.asarray(sd[PPN + br + ".2.bias"], np.float64)

    z = emb(x, "sub_emb") @ emb(y, "obj_emb").T
    return 1.0 / (1.0 + np.exp(-z))


# ---------------------------------------------------------------------------
# a9 + [SPEC] s6: top-K of the flattened N x N relationness matrix
# ---------------------------------------------------------------------------
def topk_stable(scores, k: int) -> np.ndarray:
    """First ``min(k, N*N)`` flat indices ``s*N+o`` in descending score order.

    ppn.py:84-85 sorts the flattened matrix (diagonal included, quirk Q1) with an
    unstable ``torch.sort`` and keeps the first K (quirk Q2: ``K_eff = min(K, N^2)``).
    Ties are broken towards the lower flat index — ``torch.sort(stable=True)`` — which
    is the one order the unstable sort is allowed to return that is reproducible.
    """
    flat = np.asarray(scores).reshape(-1)
    order = np.argsort(-flat.astype(np.float64), kind="stable")
    return order[:min(int(k), flat.shape[0])].astype(np.int64)


def topk_margin(scores, k: int) -> float:
    """Gap between the K-th and (K+1)-th score (inf when everything is kept)."""
    flat = np.sort(np.asarray(scores, dtype=np.float64).reshape(-1))[::-1]
    if k >= flat.shape[0]:
        return float("inf")
    return float(flat[k - 1] - flat[k])


# ---------------------------------------------------------------------------
# a14: RelationPredictor.forward, lib/modeling/model.py:85-88
# ---------------------------------------------------------------------------
def relation_predictor_ref(feats, sd) -> torch.Tensor:
    return torch.sigmoid(F.linear(_t(feats).float(), _t(sd[CLSK + "weight"]), _t(sd[CLSK + "bias"])))


def relation_predictor_f64(feats, sd) -> np.ndarray:
    z = np.asarray(feats, np.float64) @ np.asarray(sd[CLSK + "weight"], np.float64).T \
        + np.asarray(sd[CLSK + "bias"], np.float64)
    return 1.0 / (1.0 + np.exp(-z))


# ---------------------------------------------------------------------------
# a11: DPNHead.forward, lib/modeling/relpn/dpn.py:69-73
# ---------------------------------------------------------------------------
def dpn_head_ref(feats, sd) -> torch.Tensor:
    """``Conv1d(k=3, pad=1) -> ReLU -> Conv1d(k=1)`` on ``[K, Cin, T]`` (dpn.py:69-73)."""
    x = _t(feats).float()
    hid = F.relu(F.conv1d(x, _t(sd[DPNK + "conv.weight"]), _t(sd[DPNK + "conv.bias"]), padding=1))
    return F.conv1d(hid, _t(sd[DPNK + "duration_pred.weight"]), _t(sd[DPNK + "duration_pred.bias"]))


def dpn_head_f64(feats, sd) -> np.ndarray:
    x = np.asarray(feats, np.float64)
    w = np.asarray(sd[DPNK + "conv.weight"], np.float64)          # [Co, Ci, 3]
    b = np.asarray(sd[DPNK + "conv.bias"], np.float64)
    k, ci, t = x.shape
    xp = np.zeros((k, ci, t + 2))
    xp[:, :, 1:-1] = x
    hid = b[None, :, None] + sum(np.einsum("oc,kct->kot", w[:, :, d], xp[:, :, d:d + t]) for d in range(3))
    hid = np.maximum(hid, 0)
    w2 = np.asarray(sd[DPNK + "duration_pred.weight"], np.float64)[:, :, 0]
    b2 = np.asarray(sd[DPNK + "duration_pred.bias"], np.float64)
    return np.einsum("oc,kct->kot", w2, hid) + b2[None, :, None]


# ---------------------------------------------------------------------------
# a12: anchors, restated from lib/modeling/relpn/anchor_generator.py:48-104
# ---------------------------------------------------------------------------
def base_anchors(sizes, stride: float) -> np.ndarray:
    """``[A, 2]`` windows ``[-w/2, +w/2]`` for ``w = stride * (size/stride)``.

    anchor_generator.py:66-104: the reference scales a ``[0, stride]`` reference
    window by ``sizes/stride`` around centre 0.  (Its own generator cannot run on
    numpy >= 1.24 because of ``np.float``, :72/:80 — hence a restatement.)
    """
    ratio = np.asarray(sizes, dtype=np.float64) / float(stride)
    ws = float(stride) * ratio
    return np.stack([0.0 - 0.5 * ws, 0.0 + 0.5 * ws], axis=1).astype(np.float32)


def grid_anchors(time_width: int, sizes, stride: float) -> np.ndarray:
    """``[L*A, 2]`` anchors, location-major: shifts ``arange(0, T+1, stride)`` (:48-59)."""
    shifts = torch.arange(0, time_width + 1, step=stride, dtype=torch.float32).numpy()
    base = base_anchors(sizes, stride)
    return (shifts.reshape(-1, 1, 1) + base.reshape(1, -1, 1)).reshape(-1, 2).astype(np.float32)


# ---------------------------------------------------------------------------
# [SPEC] s5: span decode to integer frame bounds (float64 definition)
# ---------------------------------------------------------------------------
def decode_spans_f64(reg: np.ndarray, sizes, stride: float) -> np.ndarray:
    """``reg [K, 2A, T]`` -> ``[K, L*A, 2]`` int32 ``[start, end)`` frame bounds.

    Location ``l`` has anchor centre ``a_c = l*stride`` and samples the regression at
    feature column ``min(floor(a_c), T-1)``; anchor ``a`` has width ``a_w = sizes[a]``
    and uses channels ``(2a, 2a+1) = (delta_c, delta_w)``:
    ``ctr = a_c + delta_c*a_w``, ``w = a_w*exp(min(delta_w, log(1000/16)))``,
    ``start = clamp(floor(ctr - w/2 + 1/2), 0, T-1)``,
    ``end = clamp(floor(ctr + w/2 + 1/2), start+1, T)``.
    """
    k, c2, t = reg.shape
    a_n = c2 // 2
    n_loc = int(math.ceil((t + 1) / float(stride)))   # len(arange(0, T+1, stride))
    out = np.zeros((k, n_loc * a_n, 2), dtype=np.int32)
    for l in range(n_loc):
        ac = float(np.float32(l) * np.float32(stride))
        col = min(int(math.floor(ac)), t - 1)
        for a in range(a_n):
            aw = float(np.float32(sizes[a]))
            dc = reg[:, 2 * a, col].astype(np.float64)
            dw = np.minimum(reg[:, 2 * a + 1, col].astype(np.float64), SPAN_DW_CLAMP)
            ctr = ac + dc * aw
            w = aw * np.exp(dw)
            st = np.clip(np.floor(ctr - 0.5 * w + 0.5), 0, t - 1)
            en = np.clip(np.floor(ctr + 0.5 * w + 0.5), st + 1, t)
            out[:, l * a_n + a, 0] = st.astype(np.int32)
            out[:, l * a_n + a, 1] = en.astype(np.int32)
    return out


# ---------------------------------------------------------------------------
# a12 + [SPEC] s8: span suppression and top-n.  lib/modeling/relpn/rel_nms.py:6-15 carries the
# parameters (nms_threshold 0.5, top_k_proposals = NUM_DURATION_PROPOSALS) but its forward is a
# stub, so THIS function is the definition ("parity unpinned by the reference").
# ---------------------------------------------------------------------------
def select_spans(cands, windows, n_keep: int = 64, nms_threshold: float = 0.5, n_cand=None):
    """Greedy temporal NMS + top-``n_keep`` per row, integers only.

    ``cands [R, M, 2]`` int (start, end) in the decode's order (location major, anchor minor), ``windows
    [R, 2]`` the pairs' temporal overlap windows (empty window: ``end <= start``), ``n_cand [R]`` the number
    of valid candidates of each row (default M).  Rank key of candidate i: ``q_i = floor(2^15 * inter / union)``
    of its span with the window, ties to the lower i.  Repeatedly keep the best live candidate and drop every
    live candidate j with ``inter(j, kept) * 1024 > round(1024 * thr) * union(j, kept)``.
    Returns ``(kept [R, n_keep, 2] int32 zero padded, counts [R] int32)``.
    """
    c = np.asarray(cands, dtype=np.int64)
    r_n, m = c.shape[0], c.shape[1]
    w = np.asarray(windows, dtype=np.int64).reshape(r_n, 2)
    wa = np.where(w[:, 1] > w[:, 0], w[:, 0], 0)[:, None]
    wb = np.where(w[:, 1] > w[:, 0], w[:, 1], 0)[:, None]
    s, e = c[:, :, 0], c[:, :, 1]
    idx = np.arange(m, dtype=np.int64)[None, :]
    nc = np.full(r_n, m, dtype=np.int64) if n_cand is None else np.asarray(n_cand, dtype=np.int64)
    valid = idx < nc[:, None]
    inter = np.clip(np.minimum(e, wb) - np.maximum(s, wa), 0, None)
    uni = (e - s) + (wb - wa) - inter
    q = np.where(uni > 0, (inter << 15) // np.maximum(uni, 1), 0)
    key = np.where(valid, (1 << 28) | (q << 12) | (4095 - idx), 0)
    thr = int(nms_threshold * 1024.0 + 0.5)
    out = np.zeros((r_n, n_keep, 2), dtype=np.int32)
    counts = np.zeros(r_n, dtype=np.int32)
    rows = np.arange(r_n)
    for k in range(n_keep):
        best = key.max(axis=1) if m else np.zeros(r_n, dtype=np.int64)
        alive = best > 0
        if not alive.any():
            break
        i = 4095 - (best & 4095)
        i = np.where(alive, i, 0)
        sw, ew = s[rows, i], e[rows, i]
        out[alive, k, 0] = sw[alive]
        out[alive, k, 1] = ew[alive]
        counts += alive
        it = np.minimum(e, ew[:, None]) - np.maximum(s, sw[:, None])
        un = (e - s) + (ew - sw)[:, None] - it
        sup = (it > 0) & (it * 1024 > thr * un)
        sup[rows, i] = True
        key[sup & alive[:, None]] = 0
    return out, counts


# ---------------------------------------------------------------------------
# N1: predict.py:66-117 top-K post-processing (per video)
# ---------------------------------------------------------------------------
def postprocess_ref(rel_logit, cls, pairs, topk_per_pair: int, topk_per_seg: int,
                    fix_q4: bool = True):
    """Top-``topk_per_pair`` predicates per pair, then top-``topk_per_seg`` per video.

    predict.py:70-81 sorts each pair's predicate scores, keeps the first 20, flattens
    and sorts again, keeps the first 200.  Ties are broken towards the lower index at
    both levels ([SPEC] s6).  Returns ``score [M]``, ``triplet [M, 3]`` =
    (subject class, predicate, object class) and ``pair_tid [M, 2]``.  The reference
    gathers the class logits from row ``(N-1)*tid`` of the *pair* table
    (predict.py:89-90), which for the object is the wrong row (quirk Q4);
    ``fix_q4=True`` takes ``argmax(cls[tid])`` for both, ``False`` mirrors the quirk.
    """
    z = np.asarray(rel_logit, dtype=np.float32)
    p, r = z.shape
    kp = min(topk_per_pair, r)
    order = np.argsort(-z.astype(np.float64), axis=1, kind="stable")[:, :kp]
    top = np.take_along_axis(z, order, axis=1)
    flat = top.reshape(-1)
    sel = np.argsort(-flat.astype(np.float64), kind="stable")[:min(topk_per_seg, flat.shape[0])]
    pi, ci = sel // kp, sel % kp
    pred = order[pi, ci]
    tid = np.asarray(pairs)[pi]
    cls = np.asarray(cls)
    n = cls.shape[0]
    if fix_q4:
        s_lab = cls[tid[:, 0]].argmax(axis=1)
        o_lab = cls[tid[:, 1]].argmax(axis=1)
    else:
        # rows of the pair-feature table: subject classeme of pair row (N-1)*tid
        pr = np.asarray(pairs)
        s_lab = cls[pr[(n - 1) * tid[:, 0], 0]].argmax(axis=1)
        o_lab = cls[pr[np.minimum((n - 1) * tid[:, 1], pr.shape[0] - 1), 1]].argmax(axis=1)
    trip = np.stack([s_lab, pred, o_lab], axis=1).astype(np.int64)
    return flat[sel], trip, tid.astype(np.int64)
