"""On-disk formats either side of the pair stage.

The reference stores one relation-feature file per 30-frame segment,
``<features>/relation/<vid>/<vid>-<fstart>-<fend>-relation.h5`` with four datasets
(lib/dataset/vrdataset.py:190-217, lib/modeling/feature.py:118-145):

    trackid  int   [N + M]           -1 for the N tracklet proposals, >= 0 for the M ground-truth tracklets
    pairs    int   [(N+M)(N+M-1), 2] every ordered pair of the N + M tracklets
    feats    float [same rows, F]    [subject classeme | object classeme | 2 x 4000 motion BoW | 3000 relative]
    iou      float [N + M, N + M]    trajectory vIoU matrix (trajectory.py:144-158)

and ``VRDataset.__getitem__`` (vrdataset.py:61-83) turns it into the ``PairList`` the model consumes: rows whose two
tracklets are both proposals (vrdataset.py:140-145), the eight BoW blocks L1-normalised (vrdataset.py:219-243).

``h5py`` is not part of this image, so the segment container read here is ``.npz`` with the same four arrays under
the same names; ``convert_h5_segment`` writes it from an upstream file wherever ``h5py`` exists (the one step that
needs it), and ``load_segment`` accepts either.  The CUDA path does the BoW normalisation
(``tspn_normalize_motion``) - the rows are handed over raw, exactly as the file holds them.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from . import ops
from .list_pair import PairList

SEGMENT_KEYS = ("trackid", "pairs", "feats", "iou")


def segment_signature(vid: str, fstart: int, fend: int) -> str:
    """lib/modeling/__init__.py:5-9."""
    return "{}-{:04d}-{:04d}".format(vid, fstart, fend)


def convert_h5_segment(h5_path: str, npz_path: Optional[str] = None) -> str:
    """``*-relation.h5`` -> ``*-relation.npz`` (needs ``h5py``; run it where the upstream features live)."""
    try:
        import h5py
    except ImportError as e:
        raise RuntimeError("convert_h5_segment needs h5py (not installed here); run it next to the dataset, the "
                           ".npz it writes is what load_segment reads") from e
    npz_path = npz_path or os.path.splitext(h5_path)[0] + ".npz"
    with h5py.File(h5_path, "r") as fin:
        arrays = {k: fin[k][:] for k in SEGMENT_KEYS}
    np.savez(npz_path, **arrays)
    return npz_path


def load_segment(path: str):
    """``(pairs, feats, iou, trackid)`` of one segment file (``.npz``, or ``.h5`` when h5py is importable) - the
    tuple ``VRDataset._get_rel_feature`` returns (vrdataset.py:190-217)."""
    if path.endswith(".h5"):
        try:
            import h5py
        except ImportError as e:
            raise RuntimeError("reading %s needs h5py; convert it with formats.convert_h5_segment where h5py is "
                               "available" % path) from e
        with h5py.File(path, "r") as fin:
            d = {k: fin[k][:] for k in SEGMENT_KEYS}
    else:
        with np.load(path, allow_pickle=False) as fin:
            missing = [k for k in SEGMENT_KEYS if k not in fin]
            if missing:
                raise ValueError("%s lacks the datasets %s" % (path, missing))
            d = {k: fin[k] for k in SEGMENT_KEYS}
    pairs, feats, iou, trackid = d["pairs"], d["feats"], d["iou"], d["trackid"]
    n = int(trackid.shape[0])
    if pairs.ndim != 2 or pairs.shape[1] != 2 or pairs.shape[0] != feats.shape[0] or iou.shape != (n, n):
        raise ValueError("%s: inconsistent shapes pairs %s feats %s iou %s trackid %s"
                         % (path, pairs.shape, feats.shape, iou.shape, trackid.shape))
    return pairs, feats, iou, trackid


def segment_pair_list(pairs, feats, iou, trackid, track_cls_logits, normalize: bool = True,
                      device: Optional[str] = None) -> PairList:
    """The ``PairList`` of ``VRDataset.__getitem__`` (vrdataset.py:61-83) without the labels: proposal-proposal
    rows only (vrdataset.py:140-145), ``tracklet_pairs`` / ``track_cls_logits`` / ``num_tracklets`` / ``ious`` /
    ``track_ids`` fields.  ``normalize``: L1-normalise the eight 1000-wide BoW blocks (vrdataset.py:219-243) -
    on the GPU (``tspn_normalize_motion``) when ``device`` is a CUDA device, else with the same arithmetic in torch."""
    trackid = np.asarray(trackid)
    pairs = np.asarray(pairs, dtype=np.int64)
    keep = (trackid[pairs[:, 0]] < 0) & (trackid[pairs[:, 1]] < 0)
    rows = torch.as_tensor(np.asarray(feats)[keep], dtype=torch.float32)
    cls = torch.as_tensor(track_cls_logits, dtype=torch.float32)
    c = int(cls.shape[1])
    if rows.shape[1] != 2 * c + 11000:
        raise ValueError("feature rows have %d columns, expected 2*%d + 8000 + 3000" % (rows.shape[1], c))
    if normalize and rows.shape[0]:
        blocks = rows[:, 2 * c:2 * c + 8000]
        if device is not None and str(device).startswith("cuda"):
            flat = blocks.reshape(-1, 4000).to(device)            # rows of 4 blocks: the kernel's [n, 4000] layout
            rows[:, 2 * c:2 * c + 8000] = ops.normalize_motion(flat.contiguous()).reshape(-1, 8000).cpu()
        else:
            b = blocks.reshape(rows.shape[0], 8, 1000)
            s = b.abs().sum(dim=2, keepdim=True)
            s[s == 0] = 1                                          # lib/utils/miscellaneous.py:32-35
            rows[:, 2 * c:2 * c + 8000] = (b / s).reshape(rows.shape[0], 8000)
    pl = PairList(rows)
    pl.add_field("tracklet_pairs", torch.as_tensor(pairs[keep]))
    pl.add_field("track_cls_logits", cls)
    pl.add_field("num_tracklets", int((trackid < 0).sum()))
    pl.add_field("ious", torch.as_tensor(np.asarray(iou), dtype=torch.float32))
    pl.add_field("track_ids", torch.as_tensor(trackid))
    return pl
