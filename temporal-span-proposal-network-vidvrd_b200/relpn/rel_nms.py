"""RelNMS — the reference's is a stub (lib/modeling/relpn/rel_nms.py:14-15: ``forward`` evaluates
``relationness`` and returns None).  Kept for constructor parity; it carries the same thresholds and
defines no behaviour of its own, so decoded spans are returned unsuppressed."""
import torch.nn as nn


class RelNMS(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.fg_iou_threshold = 0.7
        self.bg_iou_threshold = 0.3
        self.nms_threshold = 0.5
        self.top_k_proposals = cfg.RELPN.DPN.NUM_DURATION_PROPOSALS
        self.anchor = None

    def forward(self, relationness, duration_proposals):
        return duration_proposals
