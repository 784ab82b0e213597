"""Pair Proposal Network — mirror of lib/modeling/relpn/ppn.py:7-118 on the CUDA path."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..batch import HostBatch
from ._runtime import WeightCache, batch_from_pair_lists, compute_device, like_input


class PPNHead(nn.Module):
    """``sigmoid(sub_emb(x) @ obj_emb(y).T)`` (ppn.py:92-112).  Same parameter names, so a
    reference checkpoint loads unchanged; eval-mode forward runs ``tspn_relationness``."""

    def __init__(self, in_channels, hidden_channels, out_channels):
        super().__init__()
        self.sub_emb = nn.Sequential(nn.Linear(in_channels, hidden_channels), nn.ReLU(True),
                                     nn.Linear(hidden_channels, out_channels))
        self.obj_emb = nn.Sequential(nn.Linear(in_channels, hidden_channels), nn.ReLU(True),
                                     nn.Linear(hidden_channels, out_channels))
        self._cache = WeightCache()

    def device_weights(self, device):
        w = self._cache.get(self, device)
        return [w[k] for k in ops.PPN_KEYS]

    def forward(self, sub_logits, obj_logits):
        if self.training:       # training is outside the CUDA path
            return torch.sigmoid(torch.mm(self.sub_emb(sub_logits), self.obj_emb(obj_logits).t()))
        dev = compute_device(sub_logits, obj_logits)
        with torch.cuda.device(dev):            # the library launches on the current device's current stream
            return self._forward_cuda(sub_logits, obj_logits, dev)

    def _forward_cuda(self, sub_logits, obj_logits, dev):
        same = sub_logits is obj_logits or (sub_logits.data_ptr() == obj_logits.data_ptr()
                                             and sub_logits.shape == obj_logits.shape)
        ns, no = int(sub_logits.shape[0]), int(obj_logits.shape[0])
        x = sub_logits if same else torch.cat([sub_logits, obj_logits], dim=0)
        n = int(x.shape[0])
        host = HostBatch([torch.zeros((n, 1, 4))], [torch.tensor([[0, 1]] * n, dtype=torch.int32).reshape(n, 2)],
                         [x.detach().float().cpu() if not x.is_cuda else x.detach().float()])
        batch = host.to_device(dev)
        scores = ops.relationness(batch, self.device_weights(dev)).view(n, n)
        if not same:
            scores = scores[:ns, ns:ns + no].contiguous()
        return like_input(scores, sub_logits.is_cuda)


class PPN(nn.Module):
    """ppn.py:7-90.  Eval: per PairList the first ``min(K, N*N)`` flat indices ``s*N+o`` of the
    descending relationness scores, diagonal included (quirks Q1/Q2), ties to the lower index."""

    def __init__(self, cfg):
        super().__init__()
        self.num_pair_proposals = cfg.RELPN.PPN.NUM_PAIR_PROPOSALS
        self.ppn_head = PPNHead(in_channels=cfg.RELPN.PPN.IN_CHANNELS,
                                hidden_channels=cfg.RELPN.PPN.HIDDEN_CHANNELS,
                                out_channels=cfg.RELPN.PPN.OUT_CHANNELS)

    def forward(self, pair_list, target_list=None):
        if self.training:
            return self._forward_train(pair_list, target_list)
        return self._forward_test(pair_list)

    def _forward_test(self, pair_list):
        cls0 = pair_list[0].get_field("track_cls_logits")
        dev = compute_device(cls0)
        with torch.cuda.device(dev):
            return self._forward_cuda(pair_list, cls0, dev)

    def _forward_cuda(self, pair_list, cls0, dev):
        batch = batch_from_pair_lists(pair_list, dev, need_motion=False)
        scores = ops.relationness(batch, self.ppn_head.device_weights(dev))
        idx, _, _ = ops.topk_pairs(batch, scores, int(self.num_pair_proposals), exclude_diagonal=False)
        out = []
        for v, n in enumerate(batch.n):
            out.append(like_input(idx[v, :min(int(self.num_pair_proposals), n * n)], cls0.is_cuda))
        return out, {}

    # -- training: stock autograd, as in the reference (ppn.py:36-77); not part of the CUDA path --
    def _forward_train(self, pair_list, target_list):
        pair_proposals, loss = [], 0
        for plist, tlist in zip(pair_list, target_list):
            cl = plist.get_field("track_cls_logits")
            m = self.ppn_head(cl, cl)
            n = plist.get_field("num_tracklets")
            gt = torch.zeros(n, n, device=m.device)
            pos = tlist.target.sum(dim=1) > 0
            tp = plist.get_field("tracklet_pairs")[pos.to(plist.get_field("tracklet_pairs").device)]
            gt[tp[:, 0], tp[:, 1]] = 1
            loss = loss + F.binary_cross_entropy(m, gt)
            order = torch.sort(m.view(-1), descending=True, stable=True)[1]
            pair_proposals.append(order[:self.num_pair_proposals])
        return pair_proposals, {"loss_pair": loss}


def make_ppn(cfg):
    return PPN(cfg)
