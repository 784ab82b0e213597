"""Oracle: per-pair feature rows.  TEST INFRASTRUCTURE ONLY.

Layout of one row (lib/dataset/vrdataset.py:219-243, consumed at
lib/modeling/predict.py:66-67 and lib/modeling/model.py:31,57)::

    [0   : C      )  subject classeme
    [C   : 2C     )  object classeme
    [2C  : 2C+4000)  subject motion BoW, 4 blocks of 1000, each L1-normalised
    [+4000: +8000 )  object motion BoW, same
    [F-3000 : F   )  relative position (1000) + size (1000) + motion (1000)

The reference loads the last 3000 columns precomputed from h5 and has no code that
computes them; ``relative_block`` is the [SPEC] s4 definition (PARITY UNPINNED by
the reference).
"""
from __future__ import annotations

import numpy as np

MOTION_DIM = 4000
MOTION_BLOCK = 1000
REL_BINS = 500
REL_DIM = 6 * REL_BINS
# (channel pairs of oracle.geometry.pair_geometry) -> position, size, motion
REL_CHANNELS = (0, 1, 2, 3, 5, 6)


def l1_normalize_ref(x: np.ndarray, axis: int = -1) -> np.ndarray:
    """``normalize(x, axis, order=1)`` of lib/utils/miscellaneous.py:32-35.

    ``x / ||x||_1`` with zero norms replaced by 1 so empty histograms stay zero.
    """
    nrm = np.atleast_1d(np.linalg.norm(x, 1, axis))
    nrm[nrm == 0] = 1
    return x / np.expand_dims(nrm, axis)


def normalize_motion_ref(motion: np.ndarray) -> np.ndarray:
    """L1-normalise each of the four 1000-wide BoW blocks (vrdataset.py:227-236)."""
    out = np.array(motion, dtype=np.float32, copy=True)
    for k in range(MOTION_DIM // MOTION_BLOCK):
        sl = slice(k * MOTION_BLOCK, (k + 1) * MOTION_BLOCK)
        out[:, sl] = l1_normalize_ref(out[:, sl], axis=-1)
    return out


def adaptive_bins(length: int, bins: int = REL_BINS):
    """Start/end (exclusive) source index of each output bin of ``adaptive_avg_pool1d``.

    ``start = floor(i*L/bins)``, ``end = ceil((i+1)*L/bins)`` — torch's
    AdaptiveAvgPool definition, integer arithmetic.
    """
    i = np.arange(bins, dtype=np.int64)
    st = (i * length) // bins
    en = -((-(i + 1) * length) // bins)
    return st, en


def relative_block(geo: np.ndarray, overlap: np.ndarray, bins: int = REL_BINS) -> np.ndarray:
    """[SPEC] s4: adaptive average pooling of the geometry channels over the overlap window.

    ``geo [P, 8, T]``, ``overlap [P, 2]`` -> ``[P, 6*bins]`` laid out as
    ``[ch0 | ch1 | ch2 | ch3 | ch5 | ch6]`` (position, size, motion: the reference's
    3x1000 blocks, vrdataset.py:238-241).  Pairs with an empty window give zeros.
    """
    p = geo.shape[0]
    out = np.zeros((p, len(REL_CHANNELS) * bins), dtype=np.float64)
    for r in range(p):
        a, b = int(overlap[r, 0]), int(overlap[r, 1])
        length = b - a
        if length <= 0:
            continue
        st, en = adaptive_bins(length, bins)
        for j, ch in enumerate(REL_CHANNELS):
            cs = np.concatenate([[0.0], np.cumsum(geo[r, ch, a:b], dtype=np.float64)])
            out[r, j * bins:(j + 1) * bins] = (cs[en] - cs[st]) / (en - st)
    return out


def relative_block_direct(geo_row: np.ndarray, a: int, b: int, bins: int = REL_BINS) -> np.ndarray:
    """Same as ``relative_block`` for one pair, by direct per-bin summation (cross-check)."""
    out = np.zeros(len(REL_CHANNELS) * bins, dtype=np.float64)
    length = b - a
    if length <= 0:
        return out
    st, en = adaptive_bins(length, bins)
    for j, ch in enumerate(REL_CHANNELS):
        for i in range(bins):
            out[j * bins + i] = geo_row[ch, a + st[i]:a + en[i]].sum() / (en[i] - st[i])
    return out


def assemble_features(cls: np.ndarray, motion: np.ndarray, rel: np.ndarray,
                      pairs: np.ndarray) -> np.ndarray:
    """Build ``[P, F]`` rows in the layout of vrdataset.py:219-243 (see module docstring)."""
    mn = normalize_motion_ref(motion).astype(np.float64)
    c = np.asarray(cls, dtype=np.float64)
    s, o = pairs[:, 0], pairs[:, 1]
    return np.concatenate([c[s], c[o], mn[s], mn[o], rel], axis=1)
