"""Duration Proposal Network — mirror of lib/modeling/relpn/dpn.py:9-81 on the CUDA path."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib, ops
from .rel_nms import RelNMS
from ._runtime import WeightCache, batch_from_pair_lists, compute_device, like_input


class DPNHead(nn.Module):
    """``Conv1d(C->C, k3, p1) -> ReLU -> Conv1d(C->2A, k1)`` on ``[K, C, T]`` (dpn.py:55-73)."""

    def __init__(self, in_channels, num_windows, precision="fp32"):
        super().__init__()
        self.conv = nn.Conv1d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)
        self.duration_pred = nn.Conv1d(in_channels, num_windows * 2, kernel_size=1, stride=1)
        for l in [self.conv, self.duration_pred]:
            torch.nn.init.normal_(l.weight, std=0.01)
            torch.nn.init.constant_(l.bias, 0)
        self.precision = precision
        self._cache = WeightCache()

    def device_weights(self, device):
        w = self._cache.get(self, device)
        return w["conv.weight"], w["conv.bias"], w["duration_pred.weight"], w["duration_pred.bias"]

    def forward(self, feats):
        if self.training:
            return self.duration_pred(F.relu(self.conv(feats)))
        dev = compute_device(feats)
        with torch.cuda.device(dev):            # the library launches on the current device's current stream
            x = feats.detach().to(dev, torch.float32)
            cw, cb, pw, pb = self.device_weights(dev)
            out = ops.span_head(x, cw, cb, pw, pb, precision=self.precision)
            return like_input(out, feats.is_cuda)


class DPN(nn.Module):
    """dpn.py:9-52.  The reference's ``forward`` raises NameError (quirk Q5) and its NMS is a stub;
    eval here is [SPEC] s5 + s8: DPNHead on the per-frame geometry of every pair of the PairList
    (``boxes``/``span`` fields), decoded to integer frame bounds, then ``RelNMS``: ``[P, NUM_DURATION_PROPOSALS,
    2]`` int16 (``[P, L*A, 2]`` int32, every decoded span, when ``NUM_DURATION_PROPOSALS`` is 0)."""

    def __init__(self, cfg, in_channels, num_windows):
        super().__init__()
        self.dpn_head = DPNHead(in_channels=in_channels, num_windows=num_windows)
        self.rel_nms = RelNMS(cfg)
        sizes = getattr(cfg.RELPN.DPN, "ANCHOR_SIZES", None)
        stride = getattr(cfg.RELPN.DPN, "ANCHOR_STRIDE", None)
        if not isinstance(sizes, (list, tuple)) or len(sizes) != num_windows:
            sizes, stride = [15.0 * (i + 1) for i in range(num_windows)], 7.5
        self.anchor_sizes = tuple(float(s) for s in sizes)
        self.anchor_stride = float(stride)
        self.rel_nms.n_anchors, self.rel_nms.anchor_stride = int(num_windows), self.anchor_stride

    def forward(self, pair_list, target_list=None):
        if self.training:
            raise NotImplementedError("DPN training is undefined in the reference (dpn.py:24-28 raises NameError)")
        return self._forward_test(pair_list)

    def _forward_test(self, pair_list):
        if not all(pl.has_field("boxes") for pl in pair_list):
            raise ValueError("DPN needs the dense tracklet fields 'boxes' and 'span' on every PairList")
        ref = pair_list[0].get_field("boxes")
        dev = compute_device(ref)
        with torch.cuda.device(dev):
            return self._forward_cuda(pair_list, ref, dev)

    def _forward_cuda(self, pair_list, ref, dev):
        batch = batch_from_pair_lists(pair_list, dev, need_motion=False)
        geom = ops.pair_geometry(batch, write_geo=True)
        cw, cb, pw, pb = self.dpn_head.device_weights(dev)
        if cw.shape[1] != _lib.GEO_CHANNELS:
            raise ValueError("DPN consumes the %d geometry channels: RELPN.DPN.IN_CHANNELS must be %d"
                             % (_lib.GEO_CHANNELS, _lib.GEO_CHANNELS))
        sizes = torch.tensor(self.anchor_sizes, dtype=torch.float32, device=dev)
        out = []
        for v in range(batch.num_videos):
            sp = ops.span_proposals(batch.geo_rows(geom["geo"], v), cw, cb, pw, pb, sizes, self.anchor_stride,
                                    t=batch.t[v])
            if sp.shape[0]:
                sp = self.rel_nms(None, sp, windows=geom["overlap"][batch.pair_slice(v)])
            out.append(like_input(sp, ref.is_cuda))
        return out, {}


def make_dpn(cfg):
    return DPN(cfg, in_channels=cfg.RELPN.DPN.IN_CHANNELS, num_windows=cfg.RELPN.DPN.NUM_ANCHORS_PER_LOCATION)
