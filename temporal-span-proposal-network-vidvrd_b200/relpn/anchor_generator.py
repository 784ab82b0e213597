"""Temporal anchors — restatement of lib/modeling/relpn/anchor_generator.py:31-104.

(The reference's generator uses ``np.float`` and cannot run on numpy >= 1.24.)  Anchors are a
host-side table of at most a few thousand floats; the decode that consumes them runs on the
GPU (``tspn_span_decode``), which regenerates centre/width from ``(l*stride, size)``.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn


def generate_anchors(stride=8, sizes=(4, 8, 16)):
    """``[A, 2]`` windows ``[-w/2, +w/2]`` (anchor_generator.py:66-104: a ``[0, stride]`` reference
    window scaled by ``sizes/stride`` around centre 0)."""
    ratio = np.asarray(sizes, dtype=np.float64) / float(stride)
    ws = float(stride) * ratio
    return torch.from_numpy(np.stack([0.0 - 0.5 * ws, 0.0 + 0.5 * ws], axis=1))


class AnchorGenerator(nn.Module):
    def __init__(self, sizes=(4, 8, 16), anchor_stride=8):
        super().__init__()
        self.stride = anchor_stride
        self.sizes = tuple(float(s) for s in sizes)
        self.register_buffer("cell_anchors", generate_anchors(anchor_stride, sizes).float())

    def num_anchors_per_location(self):
        return [len(self.cell_anchors)]

    def grid_anchors(self, time_width):
        shifts = torch.arange(0, time_width + 1, step=self.stride, dtype=torch.float32,
                              device=self.cell_anchors.device)
        return [(shifts.view(-1, 1, 1) + self.cell_anchors.view(1, -1, 1)).reshape(-1, 2)]

    def forward(self, rel_feats):
        return self.grid_anchors(rel_feats.shape[2])      # N x C x T (time dimension)


def make_anchor_generator(cfg):
    sizes = cfg.RELPN.DPN.ANCHOR_SIZES
    stride = cfg.RELPN.DPN.ANCHOR_STRIDE
    if not isinstance(sizes, (list, tuple)):
        sizes = (sizes,)
    return AnchorGenerator(sizes, stride)
