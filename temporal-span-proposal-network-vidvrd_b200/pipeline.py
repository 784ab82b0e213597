"""The batched pair stage: geometry + vIoU -> features -> relationness + top-K -> heads.

This is the engine behind ``BaseModel.forward`` (lib/modeling/model.py:53-65) and the thing
``bench.py`` times.  One call processes a whole batch of videos with a fixed, small number of
kernel launches on the current stream and no host synchronisation.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

import torch

from . import _lib, ops
from .batch import DeviceBatch, HostBatch

PPN_PREFIX = "relpn.pair_proposal_network.ppn_head."
DPN_PREFIX = "relpn.duration_proposal_network.dpn_head."
CLS_PREFIX = "classifier.rel_predictor."


@dataclasses.dataclass
class StageConfig:
    n_classes: int = 35
    n_predicates: int = 132
    hidden: int = 64
    topk: int = 256
    use_ppn: bool = True
    use_dpn: bool = True
    sparsify: bool = False
    precision: str = "fp32"            # "fp32" (exact order) | "tensor" (tcgen05)
    write_geo: bool = True
    anchor_sizes: tuple = (15.0, 30.0, 45.0, 60.0)
    anchor_stride: float = 7.5
    viou_clipped: bool = False
    topk_per_pair: int = 20            # predict.py:70
    topk_per_video: int = 200          # predict.py:76 (TOPK_PER_SEG)
    records: bool = True               # build the triplet records (row N1)
    mirror_q4: bool = False

    @classmethod
    def from_cfg(cls, cfg) -> "StageConfig":
        pr, rp = cfg.PREDICT, cfg.RELPN

        def opt(node, key, default):
            try:
                return node[key] if isinstance(node, dict) else getattr(node, key)
            except (KeyError, AttributeError):
                return default
        sizes = opt(rp.DPN, "ANCHOR_SIZES", (15.0, 30.0, 45.0, 60.0))
        stride = opt(rp.DPN, "ANCHOR_STRIDE", 7.5)
        n_anchor = int(rp.DPN.NUM_ANCHORS_PER_LOCATION)
        if not isinstance(sizes, (list, tuple)) or len(sizes) != n_anchor:
            # defaults.py:66-67 carries placeholder scalars (35 / 132): fall back to the only anchors the
            # reference ever instantiates (anchor_generator.py:118-120), stretched to A entries
            sizes = tuple(15.0 * (i + 1) for i in range(n_anchor))
            stride = 7.5
        return cls(n_classes=int(pr.OBJECT_NUM), n_predicates=int(pr.PREDICATE_NUM),
                   hidden=int(rp.PPN.HIDDEN_CHANNELS), topk=int(rp.PPN.NUM_PAIR_PROPOSALS),
                   use_ppn=bool(rp.USE_PPN), use_dpn=bool(rp.USE_DPN),
                   sparsify=bool(opt(pr, "SPARSIFY", False)), precision=str(opt(pr, "PRECISION", "fp32")),
                   topk_per_pair=int(pr.TOPK_PER_PAIR), topk_per_video=int(pr.TOPK_PER_SEG),
                   mirror_q4=not bool(opt(pr, "FIX_OBJECT_LABEL", True)),
                   anchor_sizes=tuple(float(s) for s in sizes), anchor_stride=float(stride))


@dataclasses.dataclass
class StageResult:
    batch: DeviceBatch
    geom: Dict[str, torch.Tensor]
    scores: Optional[torch.Tensor]          # [sum N*N]
    topk_idx: Optional[torch.Tensor]        # [V, K] int64, -1 padded
    topk_score: Optional[torch.Tensor]
    topk_row: Optional[torch.Tensor]        # [V, K] global pair rows, -1 = diagonal / padding
    features: Optional[torch.Tensor]        # [rows, F] fp32 view
    features_bf16: Optional[torch.Tensor]
    rel_logits: Optional[torch.Tensor]      # [rows, R]
    span_reg: Optional[List[torch.Tensor]]  # per video [K_v, 2A, T_v]
    spans: Optional[List[torch.Tensor]]     # per video [K_v, L*A, 2] int32
    k_eff: List[int]
    sparsify: bool
    records: Optional[torch.Tensor] = None      # [V, topk_per_video, 8] int32 (ops.RECORD_FIELDS)
    record_counts: Optional[torch.Tensor] = None

    # per-video views -------------------------------------------------------------------
    def pair_proposals(self, v: int) -> Optional[torch.Tensor]:
        return None if self.topk_idx is None else self.topk_idx[v, :self.k_eff[v]]

    def logits(self, v: int) -> torch.Tensor:
        if self.sparsify:
            k = self.topk_idx.shape[1]
            return self.rel_logits[v * k:v * k + self.k_eff[v]]
        return self.rel_logits[self.batch.pair_slice(v)]


class PairStage:
    def __init__(self, config: StageConfig):
        self.cfg = config
        self.w: Dict[str, torch.Tensor] = {}
        self.packed_cls: Optional[torch.Tensor] = None
        self.sizes_dev: Optional[torch.Tensor] = None
        self._row_off: Optional[torch.Tensor] = None
        self._row_off_k = None

    # ---- weights -------------------------------------------------------------------------
    def load_weights(self, state_dict, device="cuda") -> None:
        """Copy the reference-keyed ``state_dict`` (14 tensors) to the device
        and pre-pack the classifier for the tensor-core head."""
        ops.require_device()
        dev = torch.device(device)
        self.w = {}
        for k, v in state_dict.items():
            k = k[7:] if k.startswith("module.") else k      # lib/utils/serialize.py:13-20
            t = v if isinstance(v, torch.Tensor) else torch.as_tensor(v)
            self.w[k] = t.detach().to(dev, torch.float32).contiguous()
        self.packed_cls = None
        if self.cfg.precision == "tensor" and (CLS_PREFIX + "weight") in self.w:
            self.packed_cls = ops.pack_predicate_weights(self.w[CLS_PREFIX + "weight"])
        self.sizes_dev = torch.tensor(self.cfg.anchor_sizes, dtype=torch.float32, device=dev)

    def ppn_weights(self):
        return [self.w[PPN_PREFIX + k] for k in ops.PPN_KEYS]

    # ---- forward ----------------------------------------------------------------------------
    def k_effective(self, batch: DeviceBatch) -> List[int]:
        c = self.cfg
        if not c.use_ppn:
            return [n * max(n - 1, 0) for n in batch.n]
        return [min(c.topk, n * (n - 1) if c.sparsify else n * n) for n in batch.n]

    def forward(self, batch: DeviceBatch, features: Optional[torch.Tensor] = None,
                heads: bool = True, timers: Optional[dict] = None) -> StageResult:
        """``features``: optional precomputed ``[sum P, F]`` rows (reference mode: rows loaded from
        h5, lib/modeling/predict.py:42-57); when ``None`` they are constructed on the GPU."""
        c = self.cfg
        need_geo = features is None or c.use_dpn
        if timers is not None:      # CUDA events around the dominant kernel, on the launching stream
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        geom = ops.pair_geometry(batch, write_geo=need_geo and c.write_geo, clipped=c.viou_clipped)
        if timers is not None:
            ev1.record()
            timers["geo"] = (ev0, ev1)
        scores = idx = val = row = None
        if c.use_ppn:
            scores = ops.relationness(batch, self.ppn_weights())
            idx, val, row = ops.topk_pairs(batch, scores, c.topk, exclude_diagonal=c.sparsify)
        k_eff = self.k_effective(batch)
        sparsify = c.sparsify and c.use_ppn
        feats32 = feats16 = logits = None
        tensor = c.precision == "tensor"
        if features is None:
            if batch.motion is None or batch.cls is None:
                raise ValueError("feature construction needs the cls and motion tracklet fields")
            mn = ops.normalize_motion(batch.motion)
            rows = row.reshape(-1) if sparsify else None
            feats32, feats16 = ops.assemble_features(batch, mn, geom["geo"], geom["overlap"], rows,
                                                     want_fp32=not tensor, want_bf16=tensor)
        else:
            feats32 = features
            if sparsify:
                sel = row.reshape(-1)
                feats32 = torch.where((sel >= 0)[:, None], features[sel.clamp_min(0)], features.new_zeros(()))
        if heads:
            x = feats16 if feats16 is not None else feats32
            logits = ops.predicate_head(x, self.w[CLS_PREFIX + "weight"], self.w[CLS_PREFIX + "bias"],
                                        precision=c.precision, packed=self.packed_cls)
        span_reg = spans = None
        if heads and c.use_dpn:
            span_reg, spans = self._span_heads(batch, geom, row, k_eff)
        records = counts = None
        if heads and c.records:
            if sparsify:
                if self._row_off is None or self._row_off.shape[0] != batch.num_videos + 1 \
                        or self._row_off_k != c.topk or self._row_off.device != batch.device:
                    self._row_off = torch.arange(batch.num_videos + 1, dtype=torch.int64,
                                                 device=batch.device) * c.topk
                    self._row_off_k = c.topk
                records, counts = ops.postprocess(batch, logits, geom["overlap"], c.topk_per_pair, c.topk_per_video,
                                                  rows=row.reshape(-1), row_video_off=self._row_off,
                                                  mirror_q4=c.mirror_q4)
            else:
                records, counts = ops.postprocess(batch, logits, geom["overlap"], c.topk_per_pair, c.topk_per_video,
                                                  mirror_q4=c.mirror_q4)
        return StageResult(batch, geom, scores, idx, val, row, feats32, feats16, logits, span_reg, spans, k_eff,
                           sparsify, records, counts)

    def _span_heads(self, batch: DeviceBatch, geom, row, k_eff):
        """DPNHead + decode on the surviving pairs of every video (rows gathered inside the kernel)."""
        c = self.cfg
        cw, cb = self.w[DPN_PREFIX + "conv.weight"], self.w[DPN_PREFIX + "conv.bias"]
        pw, pb = self.w[DPN_PREFIX + "duration_pred.weight"], self.w[DPN_PREFIX + "duration_pred.bias"]
        if cw.shape[1] != _lib.GEO_CHANNELS:
            raise ValueError("the pair stage feeds the span head with the %d geometry channels; "
                             "RELPN.DPN.IN_CHANNELS must be %d (got %d)" % (_lib.GEO_CHANNELS, _lib.GEO_CHANNELS,
                                                                            cw.shape[1]))
        regs, spans = [], []
        th = batch.table_host
        same_t = len(set(batch.t)) == 1
        groups = [list(range(batch.num_videos))] if same_t else [[v] for v in range(batch.num_videos)]
        geo = geom["geo"]
        for vids in groups:
            v0 = vids[0]
            t, tp = batch.t[v0], int(th[v0, _lib.VT_TP])
            g0 = int(th[v0, _lib.VT_GEO_OFF])
            p0 = int(th[v0, _lib.VT_PAIR_OFF])
            p_cnt = sum(batch.n[v] * max(batch.n[v] - 1, 0) for v in vids)
            x = geo[g0:g0 + p_cnt * _lib.GEO_CHANNELS * tp].view(p_cnt, _lib.GEO_CHANNELS, tp)
            rsel = row[vids[0]:vids[-1] + 1].reshape(-1) if row is not None else None
            reg = ops.span_head(x, cw, cb, pw, pb, rows=rsel, t=t, row_base=p0,
                                precision=c.precision if cw.shape[1] >= 64 else "fp32")
            sp = ops.span_decode(reg, self.sizes_dev, c.anchor_stride)
            if row is not None:
                k = row.shape[1]
                for j, v in enumerate(vids):
                    regs.append(reg[j * k:j * k + k_eff[v]])
                    spans.append(sp[j * k:j * k + k_eff[v]])
            else:
                off = 0
                for v in vids:
                    pv = batch.n[v] * max(batch.n[v] - 1, 0)
                    regs.append(reg[off:off + pv])
                    spans.append(sp[off:off + pv])
                    off += pv
        return regs, spans


def stage_from_videos(videos, weights, config: StageConfig, device="cuda"):
    """Convenience: pack ``synth.VideoTracklets`` and build a ready stage."""
    host = HostBatch.from_videos(videos)
    stage = PairStage(config)
    stage.load_weights(weights, device)
    return stage, host
