"""BaseModel / RelOIPool / RelationPredictor — mirror of lib/modeling/model.py:7-88.

Same constructor, same ``forward(pair_list, target_list=None)`` signature, same ``state_dict``
keys (``relpn.pair_proposal_network.ppn_head.*``, ``relpn.duration_proposal_network.dpn_head.*``,
``classifier.rel_predictor.*``), so the reference's predict.py:57 call and its checkpoints work
unchanged.  Eval-mode arithmetic runs in csrc/libtspn_b200.so on the current CUDA device; CPU
inputs (predict.py keeps everything on the CPU) are copied in and the results copied back.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .pipeline import PairStage, StageConfig
from .relpn import make_relpn
from .relpn._runtime import WeightCache, batch_from_pair_lists, compute_device, like_input


class RelOIPool:
    """model.py:68-73: identity without duration proposals.  With them the reference indexes the
    *list* of feature tensors by the proposals, which cannot run; spans select frames, not feature
    rows, so the features pass through unchanged."""

    def __call__(self, feats, duration_proposals):
        return feats


class RelationPredictor(nn.Module):
    """``sigmoid(Linear(F -> R))`` (model.py:76-88), init N(0, 0.01) / zero bias."""

    def __init__(self, in_channels, out_channels, precision="fp32"):
        super().__init__()
        self.rel_predictor = nn.Linear(in_channels, out_channels)
        for l in [self.rel_predictor]:
            torch.nn.init.normal_(l.weight, std=0.01)
            torch.nn.init.constant_(l.bias, 0)
        self.precision = precision
        self._cache = WeightCache()
        self._packed = (None, None)

    def device_weights(self, device):
        w = self._cache.get(self, device)
        return w["rel_predictor.weight"], w["rel_predictor.bias"]

    def packed(self, device):
        wt, _ = self.device_weights(device)
        if self._packed[0] != (wt.data_ptr(), str(device)):
            self._packed = ((wt.data_ptr(), str(device)), ops.pack_predicate_weights(wt))
        return self._packed[1]

    def forward(self, reloi_feats):
        if self.training:
            return torch.sigmoid(self.rel_predictor(reloi_feats))
        dev = compute_device(reloi_feats)
        with torch.cuda.device(dev):            # the library launches on the current device's current stream
            return self._forward_cuda(reloi_feats, dev)

    def _forward_cuda(self, reloi_feats, dev):
        x = reloi_feats.detach().to(dev)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        wt, b = self.device_weights(dev)
        tensor = self.precision == "tensor"
        if tensor and x.dtype == torch.float32 and (x.stride(0) % 4 != 0 or x.data_ptr() % 16 != 0):
            ld = ops.padded(x.shape[1], 4)           # TMA needs 16-byte row strides
            buf = torch.zeros((x.shape[0], ld), dtype=torch.float32, device=dev)
            buf[:, :x.shape[1]] = x
            x = buf[:, :x.shape[1]]
        y = ops.predicate_head(x, wt, b, precision=self.precision, packed=self.packed(dev) if tensor else None)
        return like_input(y, reloi_feats.is_cuda)


class BaseModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.use_ppn = cfg.RELPN.USE_PPN
        self.use_dpn = cfg.RELPN.USE_DPN
        self.relpn = make_relpn(cfg)
        self.rel_of_interest_pool = RelOIPool()
        self.stage_config = StageConfig.from_cfg(cfg)
        self.classifier = RelationPredictor(in_channels=cfg.PREDICT.FEATURE_DIM,
                                            out_channels=cfg.PREDICT.PREDICATE_NUM,
                                            precision=self.stage_config.precision)
        self._cache = WeightCache()
        self._stage = None
        self._stage_key = None

    def forward(self, pair_list, target_list=None):
        if self.training:
            return self._forward_train(pair_list, target_list)
        return self._forward_test(pair_list)

    # -- training: stock autograd exactly as the reference (model.py:26-51); outside the CUDA path --
    def _forward_train(self, pair_list, target_list):
        loss_dict = {}
        feats = [plist.features for plist in pair_list]
        targets = [tlist.target for tlist in target_list]
        duration_proposals = None
        if self.use_ppn or self.use_dpn:
            _, duration_proposals, relpn_losses = self.relpn(pair_list, target_list)
            loss_dict.update(relpn_losses)
        reloi_feats = self.rel_of_interest_pool(feats, duration_proposals)
        loss_relation = 0
        for reloi_feat, target in zip(reloi_feats, targets):
            loss_relation = loss_relation + F.binary_cross_entropy(self.classifier(reloi_feat), target)
        loss_dict.update({"loss_rel": loss_relation})
        return loss_dict

    # -- inference: the pair stage on the GPU ------------------------------------------------------------
    def stage(self, device) -> PairStage:
        w = self._cache.get(self, device)
        key = (id(w), str(device))
        if self._stage is None or self._stage_key != key:
            st = PairStage(self.stage_config)
            st.load_weights(w, device)
            self._stage, self._stage_key = st, key
        return self._stage

    def _forward_test(self, pair_list):
        if len(pair_list) == 0:
            return (None, None, [])
        first = pair_list[0]
        ref = first.features if first.features is not None else first.get_field("track_cls_logits")
        dev = compute_device(ref)
        on_cuda = ref.is_cuda
        have_feats = all(pl.features is not None for pl in pair_list)
        have_trk = all(pl.has_tracklets() for pl in pair_list)
        cfg = self.stage_config
        if not have_feats and not have_trk:
            raise ValueError("each PairList needs either precomputed .features or the tracklet fields "
                             "boxes/span/track_cls_logits/motion")
        if cfg.use_dpn and not have_trk:
            raise ValueError("RELPN.USE_DPN needs the tracklet fields 'boxes' and 'span' on every PairList")
        with torch.cuda.device(dev):            # the library launches on the current device's current stream
            return self._forward_cuda(pair_list, dev, on_cuda, have_feats, cfg)

    def _forward_cuda(self, pair_list, dev, on_cuda, have_feats, cfg):
        stage = self.stage(dev)
        batch = batch_from_pair_lists(pair_list, dev, need_motion=not have_feats)
        feats = None
        if have_feats:
            fl = [pl.features.detach().to(dev, torch.float32) for pl in pair_list]
            ld = ops.padded(fl[0].shape[1], 4)
            feats = torch.zeros((sum(f.shape[0] for f in fl), ld), dtype=torch.float32, device=dev)
            off = 0
            for f in fl:
                feats[off:off + f.shape[0], :f.shape[1]] = f
                off += f.shape[0]
            feats = feats[:, :fl[0].shape[1]]
        res = stage.forward(batch, features=feats)
        v_n = batch.num_videos
        pair_proposals = [like_input(res.pair_proposals(v), on_cuda) for v in range(v_n)] if cfg.use_ppn else None
        duration_proposals = [like_input(s, on_cuda) for s in res.spans] if cfg.use_dpn else None
        rel_logits = [like_input(res.logits(v), on_cuda) for v in range(v_n)]
        self.last_result = res
        return pair_proposals, duration_proposals, rel_logits
