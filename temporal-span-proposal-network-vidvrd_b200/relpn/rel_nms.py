"""RelNMS — span suppression behind lib/modeling/relpn/rel_nms.py:6-15.

The reference's module carries the parameters (``nms_threshold`` 0.5, ``top_k_proposals`` =
``RELPN.DPN.NUM_DURATION_PROPOSALS``) but its ``forward`` is a stub that returns ``None``.  Here the same
constructor and ``forward(relationness, duration_proposals)`` signature run [SPEC] s8 on the GPU
(``tspn_span_select``; definition: oracle/heads.py:select_spans): per pair, greedy temporal NMS over its decoded
spans ranked by their temporal IoU with the pair's overlap window, top ``top_k_proposals`` kept."""
import torch
import torch.nn as nn

from .. import ops


class RelNMS(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.fg_iou_threshold = 0.7
        self.bg_iou_threshold = 0.3
        self.nms_threshold = float(getattr(cfg.RELPN.DPN, "NMS_THRESHOLD", 0.5))
        self.top_k_proposals = cfg.RELPN.DPN.NUM_DURATION_PROPOSALS
        self.anchor = None
        self.n_anchors = int(cfg.RELPN.DPN.NUM_ANCHORS_PER_LOCATION)
        self.anchor_stride = 7.5           # set by DPN from the anchor configuration

    def forward(self, relationness, duration_proposals, windows=None):
        """``duration_proposals [P, L*A, 2]`` int32 decoded spans (location major, anchor minor), ``windows
        [P, 2]`` the pairs' temporal overlap windows.  ``relationness`` is accepted for signature parity: the
        rule ranks the spans of ONE pair against each other, where a per-pair score is a constant.
        Returns ``[P, top_k_proposals, 2]`` int16 (zero padded) - or the input unchanged when ``top_k_proposals``
        is 0 or no windows are given (nothing to rank by)."""
        if not self.top_k_proposals or windows is None:
            return duration_proposals
        dev = duration_proposals.device
        if not duration_proposals.is_cuda:
            ops.require_device()
            dev = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):            # the library launches on the current device's current stream
            kept, _ = ops.span_select(duration_proposals.to(dev, torch.int32), self.n_anchors, self.anchor_stride,
                                      int(self.top_k_proposals), self.nms_threshold,
                                      windows=windows.to(dev, torch.int32))
        return kept if duration_proposals.is_cuda else kept.cpu()
