"""ctypes front-end of oracle/exact/tspn_exact.c.  TEST INFRASTRUCTURE ONLY.

Fixed-order fp32 arithmetic (see the C file's header) for the bit-exact checks of
relationness scores / top-K selection, predicate scores and span frame bounds.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np

_LIB = None
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtspn_exact.so")
        src = os.path.join(os.path.dirname(path), "tspn_exact.c")
        if not os.path.exists(path) or (os.path.exists(src) and os.path.getmtime(path) < os.path.getmtime(src)):
            from oracle.build import build
            build(force=True)
        L = ctypes.CDLL(path)
        L.tspn_exact_exp.restype = ctypes.c_float
        L.tspn_exact_exp.argtypes = [ctypes.c_float]
        L.tspn_exact_sigmoid.restype = ctypes.c_float
        L.tspn_exact_sigmoid.argtypes = [ctypes.c_float]
        L.tspn_exact_linear_t.restype = None
        L.tspn_exact_linear_t.argtypes = [_f32p, ctypes.c_int64, _f32p, _f32p, _f32p, ctypes.c_int64,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.tspn_exact_pair_scores.restype = None
        L.tspn_exact_pair_scores.argtypes = [_f32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int]
        L.tspn_exact_span_head.restype = None
        L.tspn_exact_span_head.argtypes = [_f32p] * 7 + [ctypes.c_int] * 4
        L.tspn_exact_span_decode.restype = None
        L.tspn_exact_span_decode.argtypes = [_f32p, _f32p, ctypes.c_float, _i32p] + [ctypes.c_int] * 4
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(_f32p)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


ACT = {None: 0, "relu": 1, "sigmoid": 2}


def exp(x: float) -> float:
    return float(lib().tspn_exact_exp(ctypes.c_float(x)))


def sigmoid(x: float) -> float:
    return float(lib().tspn_exact_sigmoid(ctypes.c_float(x)))


def linear(x, weight, bias, act=None) -> np.ndarray:
    """``act(x @ weight.T + bias)`` with the k-ascending fma chain.  weight is [Out, In]."""
    x = np.asarray(x, dtype=np.float32)
    if x.strides[-1] != 4:
        x = np.ascontiguousarray(x)
    m, n_in = x.shape
    ldx = x.strides[0] // 4
    wt = _c(np.asarray(weight, np.float32).T)
    n_out = wt.shape[1]
    b = _c(bias)
    y = np.empty((m, n_out), dtype=np.float32)
    xp = ctypes.cast(x.ctypes.data, _f32p)
    lib().tspn_exact_linear_t(xp, ldx, _p(wt), _p(b), _p(y), n_out, m, n_in, n_out, ACT[act])
    return y


PPN = "relpn.pair_proposal_network.ppn_head."
DPNK = "relpn.duration_proposal_network.dpn_head."
CLSK = "classifier.rel_predictor."


def ppn_embeddings(cls, sd):
    out = []
    for br in ("sub_emb", "obj_emb"):
        hid = linear(cls, sd[PPN + br + ".0.weight"], sd[PPN + br + ".0.bias"], "relu")
        out.append(linear(hid, sd[PPN + br + ".2.weight"], sd[PPN + br + ".2.bias"]))
    return out


def relationness(cls, sd) -> np.ndarray:
    """PPNHead scores ``[N, N]`` in the fixed order (ppn.py:107-112 semantics)."""
    s, o = ppn_embeddings(cls, sd)
    n, c = s.shape
    m = np.empty((n, n), dtype=np.float32)
    lib().tspn_exact_pair_scores(_p(_c(s)), _p(_c(o)), _p(m), n, c)
    return m


def topk(scores, k: int) -> np.ndarray:
    """Descending, ties to the lower flat index; first ``min(k, N^2)`` ([SPEC] s6)."""
    flat = np.asarray(scores, dtype=np.float32).reshape(-1)
    order = np.argsort(-flat.astype(np.float64), kind="stable")
    return order[:min(int(k), flat.shape[0])].astype(np.int64)


def predicate(feats, sd) -> np.ndarray:
    return linear(feats, sd[CLSK + "weight"], sd[CLSK + "bias"], "sigmoid")


def span_head(feats, sd) -> np.ndarray:
    x = _c(feats)
    k, cin, t = x.shape
    cw = _c(sd[DPNK + "conv.weight"])
    cb = _c(sd[DPNK + "conv.bias"])
    pw = _c(np.asarray(sd[DPNK + "duration_pred.weight"]).reshape(-1, cin))
    pb = _c(sd[DPNK + "duration_pred.bias"])
    a2 = pw.shape[0]
    out = np.empty((k, a2, t), dtype=np.float32)
    hid = np.empty((cin, t), dtype=np.float32)
    lib().tspn_exact_span_head(_p(x), _p(cw), _p(cb), _p(pw), _p(pb), _p(out), _p(hid), k, cin, t, a2)
    return out


def n_locations(t: int, stride: float) -> int:
    """Length of ``torch.arange(0, T+1, step=stride)`` (anchor_generator.py:50-52)."""
    return int(math.ceil((t + 1) / float(stride)))


def span_decode(reg, sizes, stride: float) -> np.ndarray:
    r = _c(reg)
    k, c2, t = r.shape
    a = c2 // 2
    sz = _c(sizes)
    n_loc = n_locations(t, stride)
    out = np.empty((k, n_loc * a, 2), dtype=np.int32)
    lib().tspn_exact_span_decode(_p(r), _p(sz), ctypes.c_float(stride),
                                 out.ctypes.data_as(_i32p), k, a, t, n_loc)
    return out
