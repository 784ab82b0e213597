"""RelPN = PPN + DPN dispatcher — mirror of lib/modeling/relpn/relpn.py:9-59."""
import torch.nn as nn

from .dpn import make_dpn
from .ppn import make_ppn


class RelPN(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.use_ppn = cfg.RELPN.USE_PPN
        self.use_dpn = cfg.RELPN.USE_DPN
        self.pair_proposal_network = make_ppn(cfg)
        self.duration_proposal_network = make_dpn(cfg)

    def forward(self, pair_list, target_list=None):
        if self.training:
            return self._forward_train(pair_list, target_list)
        return self._forward_test(pair_list)

    def _forward_train(self, pair_list, target_list):
        losses, pair_proposals, duration_proposals = {}, None, None
        if self.use_ppn:
            pair_proposals, loss_ppn = self.pair_proposal_network(pair_list, target_list)
            losses.update(loss_ppn)
        if self.use_dpn:
            duration_proposals, loss_dpn = self.duration_proposal_network(pair_list, target_list)
            losses.update(loss_dpn)
        return pair_proposals, duration_proposals, losses

    def _forward_test(self, pair_list):
        pair_proposals, duration_proposals = None, None
        if self.use_ppn:
            pair_proposals, _ = self.pair_proposal_network(pair_list)
        if self.use_dpn:
            duration_proposals, _ = self.duration_proposal_network(pair_list)
        return pair_proposals, duration_proposals, {}


def make_relpn(cfg):
    return RelPN(cfg)
