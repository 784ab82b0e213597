"""The batched pair stage: geometry + vIoU -> features -> relationness + top-K -> heads.

This is the engine behind ``BaseModel.forward`` (lib/modeling/model.py:53-65) and the thing
``bench.py`` times.  One call processes a whole batch of videos with a fixed, small number of
kernel launches on the current stream and no host synchronisation.
"""
from __future__ import annotations

import dataclasses
import os
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
    relationness_precision: str = "fp32"   # PPNHead: "fp32" = exact order, hence a bit-exact top-K selection (also
                                           # when the heads run in tensor precision); "tensor" = tcgen05 (tf32 operands)
    write_geo: bool = True
    geo_layout: str = "dense"          # "dense" = [P, 8, Tp] rows (the parity layout) | "windowed" = per pair only the
                                       # frames of its overlap window, 7 channels (tspn_pair_geo_viou_windowed); the
                                       # windowed rows are an output only, so they need the survivor path (heads from
                                       # the boxes: precision "tensor" + sparsify)
    anchor_sizes: tuple = (15.0, 30.0, 45.0, 60.0)
    anchor_stride: float = 7.5
    viou_clipped: bool = False
    topk_per_pair: int = 20            # predict.py:70
    topk_per_video: int = 200          # predict.py:76 (TOPK_PER_SEG)
    records: bool = True               # build the triplet records (row N1)
    mirror_q4: bool = False
    keep_span_reg: bool = False        # also return DPNHead's raw [K, 2A, T] regressions (two-kernel path)
    materialize_features: bool = False  # tensor precision: also build the [rows, F] rows (see PairStage._decomposed)
    geo_reserve_sms: int = 8           # survivor path: SMs the persistent pair kernel leaves to the side branches
                                       # (measured on the bench workload: 0 -> 0.748 ms, 8 -> 0.735 ms, 16 -> 0.751 ms)
    num_span_proposals: int = 0        # RELPN.DPN.NUM_DURATION_PROPOSALS (defaults.py:62 = 64): temporal NMS + top-n of
                                       # the decoded spans ([SPEC] s8, tspn_span_select); 0 = all L*A decoded spans
    nms_threshold: float = 0.5         # rel_nms.py:10

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
                   relationness_precision=str(opt(pr, "RELATIONNESS_PRECISION", "fp32")),
                   topk_per_pair=int(pr.TOPK_PER_PAIR), topk_per_video=int(pr.TOPK_PER_SEG),
                   mirror_q4=not bool(opt(pr, "FIX_OBJECT_LABEL", False)),
                   anchor_sizes=tuple(float(s) for s in sizes), anchor_stride=float(stride),
                   num_span_proposals=int(opt(rp.DPN, "NUM_DURATION_PROPOSALS", 0)),
                   nms_threshold=float(opt(rp.DPN, "NMS_THRESHOLD", 0.5)))


class StageResult:
    """Outputs of one step.  The tensors are sized for the batch's launch totals (its capacity, when it has one);
    the per-video accessors and ``host_outputs`` slice them to what the batch actually holds, reading the sizes
    from the batch at call time - so the result of a replayed CUDA graph always describes the batch that was
    replayed last."""

    def __init__(self, stage: "PairStage", batch: DeviceBatch, geom: Dict[str, torch.Tensor], scores=None, topk_idx=None,
                 topk_score=None, topk_row=None, features=None, features_bf16=None, rel_logits=None, span_reg=None,
                 spans=None, sparsify: bool = False, records=None, record_counts=None, span_buffers=None,
                 span_sel=None, span_counts=None):
        self.stage, self.batch, self.geom = stage, batch, geom
        self.scores = scores                    # [sum N*N]
        self.topk_idx = topk_idx                # [V, K] int64, -1 padded
        self.topk_score = topk_score
        self.topk_row = topk_row                # [V, K] global pair rows, -1 = diagonal / padding
        self.features = features                # [rows, F] fp32 view
        self.features_bf16 = features_bf16
        self.rel_logits = rel_logits            # [rows, R]
        self.span_reg = span_reg                # per video [K_v, 2A, T_v] (StageConfig.keep_span_reg only)
        self._spans = spans                     # per video [K_v, L*A, 2] int32 views (stored-rows path)
        self.sparsify = sparsify
        self.records = records                  # [V, topk_per_video, 8] int32 (ops.RECORD_FIELDS)
        self.record_counts = record_counts
        self.span_buffers = span_buffers        # the contiguous tensors the decoded spans are views of
        self.span_sel = span_sel                # [rows, n_keep, 2] int16: spans kept by the NMS ([SPEC] s8)
        self.span_counts = span_counts          # [rows] int32

    @property
    def k_eff(self) -> List[int]:
        return self.stage.k_effective(self.batch)

    @property
    def spans(self) -> Optional[List[torch.Tensor]]:
        """Per video the decoded spans of its scored rows: ``[K_v, L_v * A, 2]`` int32 (all anchors), or with span
        selection ``[K_v, n_keep, 2]`` int16 (zero padded; ``span_count(v)`` gives the kept counts)."""
        b, c = self.batch, self.stage.cfg
        if self.span_sel is not None:
            return [self.span_sel[self._span_rows(v)] for v in range(b.num_real)]
        if self._spans is not None:
            return self._spans
        if not self.span_buffers:
            return None
        sp = self.span_buffers[0]               # survivor path: [V * K, L_max * A, 2]
        k, k_eff = self.topk_row.shape[1], self.k_eff
        a_n = sp.shape[1] // ops.span_num_locations(int(b.totals[_lib.TOT_MAX_T]), c.anchor_stride)
        return [sp[v * k:v * k + k_eff[v], :ops.span_num_locations(b.t[v], c.anchor_stride) * a_n]
                for v in range(b.num_real)]

    def span_count(self, v: int) -> Optional[torch.Tensor]:
        return None if self.span_counts is None else self.span_counts[self._span_rows(v)]

    def _span_rows(self, v: int) -> slice:
        """Rows of video v in the span outputs: the span head runs on the top-K proposals whenever there are any
        (also when the predicate head scores all P pairs, quirk Q3), else on every pair."""
        if self.topk_row is not None:
            k = self.topk_row.shape[1]
            return slice(v * k, v * k + self.k_eff[v])
        return self.batch.pair_slice(v)

    def _rows(self, v: int) -> slice:
        """Scored rows of video v in the row-indexed outputs (rel_logits, span_sel, ...)."""
        if self.sparsify:
            k = self.topk_idx.shape[1]
            return slice(v * k, v * k + self.k_eff[v])
        return self.batch.pair_slice(v)

    def host_outputs(self, full: bool = False) -> Dict[str, torch.Tensor]:
        """The tensors a caller reads back: what BaseModel.forward returns (proposals, spans, predicate
        scores), the per-pair reductions and the triplet records - sliced to the batch's true sizes (``full``:
        unsliced, i.e. sized for the batch's capacity - what a serving slot sizes its pinned buffers with)."""
        b = self.batch
        p, v = (b.total_pairs, b.num_videos) if full else (b.actual_pairs, b.num_real)
        out = {"viou": self.geom["viou"][:p], "tiou": self.geom["tiou"][:p], "overlap": self.geom["overlap"][:p]}
        rows = p
        if self.topk_idx is not None:
            out["topk_idx"], out["topk_score"] = self.topk_idx[:v], self.topk_score[:v]
            if self.sparsify:
                rows = v * self.topk_idx.shape[1]
        if self.rel_logits is not None:
            out["rel_logits"] = self.rel_logits[:rows]
        if self.span_sel is not None:
            srows = v * self.topk_row.shape[1] if self.topk_row is not None else p
            out["spans"], out["span_counts"] = self.span_sel[:srows], self.span_counts[:srows]
        else:
            for i, buf in enumerate(self.span_buffers or []):
                out["spans%d" % i] = buf[:rows] if len(self.span_buffers) == 1 else buf
        if self.records is not None:
            out["records"], out["record_counts"] = self.records[:v], self.record_counts[:v]
        return out

    def geo_window_rows(self, v: int) -> List[torch.Tensor]:
        """WINDOWED layout (``StageConfig.geo_layout='windowed'``): per pair of video v the ``[7, Lw]`` view of its rows
        - frames ``[a & ~3, (b + 3) & ~3)`` of channels 0..6, ``[a, b)`` = the pair's overlap window (``[7, 0]`` when
        the pair has none).  Synchronises (reads the offsets and windows back)."""
        if self.geom.get("geo_off") is None:
            raise ValueError("the step wrote the dense layout: use batch.geo_view(result.geom['geo'], v)")
        sl = self.batch.pair_slice(v)
        off = self.geom["geo_off"][sl].cpu().tolist()
        win = self.geom["overlap"][sl].cpu().tolist()
        ch = _lib.GEO_CHANNELS - 1
        rows = []
        for o, (a, b) in zip(off, win):
            lw = ((b + 3) & ~3) - (a & ~3) if b > a else 0
            rows.append(self.geom["geo"][o:o + ch * lw].view(ch, lw))
        return rows

    # per-video views -------------------------------------------------------------------
    def pair_proposals(self, v: int) -> Optional[torch.Tensor]:
        return None if self.topk_idx is None else self.topk_idx[v, :self.k_eff[v]]

    def logits(self, v: int) -> torch.Tensor:
        return self.rel_logits[self._rows(v)]


def side_priority() -> int:
    """Priority of the side branch's stream (-1 = high, the default).  The pair kernel keeps every SM full
    (one 512-thread CTA each, 134 KB of shared memory), so a side-branch CTA only starts when a pair CTA
    retires.  At equal priority the block scheduler hands the freed SM to the next pending pair CTA and the
    side branch (relationness, top-K, per-tracklet predicate terms - ~130 us of small kernels) is left for
    the pair kernel's last wave, i.e. back on the critical path; at high priority it claims the first SMs
    that free up and is done long before the pair kernel ends.  Captured kernel nodes keep the priority."""
    return int(os.environ.get("TSPN_SIDE_PRIORITY", "-1"))


class PairStage:
    def __init__(self, config: StageConfig):
        self.cfg = config
        self.w: Dict[str, torch.Tensor] = {}
        self.packed_cls: Optional[torch.Tensor] = None
        self.sizes_dev: Optional[torch.Tensor] = None
        self._row_off: Dict[tuple, torch.Tensor] = {}
        self._side: Dict[str, torch.cuda.Stream] = {}

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
        self.packed_cls = self.packed_sub = self.packed_obj = self.packed_rel = None
        if self.cfg.precision == "tensor" and (CLS_PREFIX + "weight") in self.w:
            w = self.w[CLS_PREFIX + "weight"]
            self.packed_cls = ops.pack_predicate_weights(w)
            c, md = self.cfg.n_classes, _lib.MOTION_DIM
            if w.shape[1] == 2 * c + 2 * md + _lib.REL_DIM:
                # decomposed head (include/tspn_b200.h, a14 decomposed): per-tracklet slices W_s, W_o over
                # [cls | motion] and the slice that multiplies the pooled relative block
                w_s = torch.cat([w[:, :c], w[:, 2 * c:2 * c + md]], dim=1)
                w_o = torch.cat([w[:, c:2 * c], w[:, 2 * c + md:2 * c + 2 * md]], dim=1)
                self.packed_sub = ops.pack_predicate_weights(w_s.contiguous())
                self.packed_obj = ops.pack_predicate_weights(w_o.contiguous())
                self.packed_rel = ops.pack_predicate_weights(w[:, 2 * c + 2 * md:].contiguous())
        self.sizes_dev = torch.tensor(self.cfg.anchor_sizes, dtype=torch.float32, device=dev)

    def ppn_weights(self):
        return [self.w[PPN_PREFIX + k] for k in ops.PPN_KEYS]

    # ---- forward ----------------------------------------------------------------------------
    def k_effective(self, batch: DeviceBatch) -> List[int]:
        c = self.cfg
        if not c.use_ppn:
            return [n * max(n - 1, 0) for n in batch.n]
        return [min(c.topk, n * (n - 1) if c.sparsify else n * n) for n in batch.n]

    # The step is three segments.  `side` depends only on the tracklet inputs (class logits, motion rows, boxes,
    # spans), so it runs on side streams underneath the HBM-bound all-pairs kernel of `geo`: relationness + top-K
    # and the per-tracklet predicate terms always; on the survivor path (tensor precision + sparsify, the bench
    # path) also the surviving pairs' relative features and span proposals recomputed from the boxes, the predicate
    # head and the records - then `tail` only finalises vIoU / tIoU and joins.  Otherwise `tail` builds the feature
    # rows from the stored geometry rows and runs the heads behind `geo`.  Eager `forward` forks and joins with
    # stream events; `capture` freezes the step into one CUDA graph (GraphedStage).
    def _side_stream(self, device, which: int = 0) -> torch.cuda.Stream:
        """Stream 0: relationness -> top-K -> surviving-pair heads (the long chain: high priority); stream 1: the
        per-tracklet predicate terms and the per-pair finalize (default priority)."""
        key = "%s/%d" % (device, which)
        if self._side.get(key) is None:
            prio = side_priority() if which == 0 else int(os.environ.get("TSPN_SIDE1_PRIORITY", "0"))
            self._side[key] = torch.cuda.Stream(device, priority=prio)
        return self._side[key]

    def _survivor_path(self, batch: DeviceBatch, features, heads: bool) -> bool:
        """Heads of the K survivors entirely on the side branch, from the boxes (csrc/survivors.cu): tensor
        precision + sparsify, features built on the GPU, and a batch the kernel supports."""
        c = self.cfg
        if not (heads and c.use_ppn and c.sparsify and self._decomposed(features)) or c.keep_span_reg:
            return False
        if os.environ.get("TSPN_SURVIVOR_PATH", "1") != "1":
            return False
        a_n = 4
        if c.use_dpn:
            cw, pw = self.w[DPN_PREFIX + "conv.weight"], self.w[DPN_PREFIX + "duration_pred.weight"]
            if cw.shape[1] != _lib.GEO_CHANNELS:
                return False                            # _span_heads raises the descriptive error
            a_n = pw.shape[0] // 2
        return batch.total_pairs > 0 and ops.survivor_rows_supported(batch, a_n)

    def _seg_side(self, batch: DeviceBatch, features: Optional[torch.Tensor], heads: bool = False):
        c = self.cfg
        scores = idx = val = row = mn = early = None
        if features is None and (batch.motion is None or batch.cls is None):
            raise ValueError("feature construction needs the cls and motion tracklet fields")
        # Two independent chains - relationness -> top-K, and the per-tracklet inputs of the predicate head - each
        # a string of small latency-bound kernels that co-reside with the pair kernel: they run side by side
        # on two streams and join before anything that needs both.
        cur = torch.cuda.current_stream(batch.device)
        second = self._side_stream(batch.device, 1)
        second.wait_stream(cur)                       # fork
        if c.use_ppn:
            rp = c.relationness_precision
            if rp != "fp32" and not (batch.cls is not None and ops.relationness_tc_supported(batch, c.hidden)):
                rp = "fp32"                           # shapes the tensor-core form does not cover: exact order
            if c.topk > 0 and batch.cls is not None and ops.relationness_topk_supported(batch, rp, c.hidden) \
                    and os.environ.get("TSPN_FUSED_TOPK", "1") == "1":
                scores, idx, val, row = ops.relationness_topk(batch, self.ppn_weights(), c.topk,
                                                              exclude_diagonal=c.sparsify, precision=rp)
            else:
                scores = ops.relationness(batch, self.ppn_weights(), precision=rp)
                idx, val, row = ops.topk_pairs(batch, scores, c.topk, exclude_diagonal=c.sparsify)
        survivors = self._survivor_path(batch, features, heads)
        rel16 = sp = row_bias = rows_done = None
        if survivors:
            # nothing below reads an output of the pair kernel: relative block and span proposals come from the
            # boxes, the record windows from the spans - the whole chain stays on this branch.  The recomputation
            # (the longest kernel of the branch) needs only the top-K rows, so it starts before the per-tracklet
            # terms of the other chain are done; this stream has the higher priority for the SM slots the pair
            # kernel leaves free.
            topk_done = torch.cuda.Event()
            topk_done.record(cur)
            sw = None
            if c.use_dpn:
                sw = (self.w[DPN_PREFIX + "conv.weight"], self.w[DPN_PREFIX + "conv.bias"],
                      self.w[DPN_PREFIX + "duration_pred.weight"], self.w[DPN_PREFIX + "duration_pred.bias"])
            rel16, _, sp = ops.survivor_rows(batch, row, span_weights=sw, sizes=self.sizes_dev, stride=c.anchor_stride)
            if sp is not None and c.num_span_proposals > 0:
                rows_done = torch.cuda.Event()
                rows_done.record(cur)
        terms_done = None
        sel = sel_counts = None
        with torch.cuda.stream(second):
            if features is None:
                if self._decomposed(features):
                    # per-tracklet terms of the decomposed head: A_s = [cls | motion_norm] W_s^T, A_o likewise
                    x_trk = ops.tracklet_rows(batch)
                    kd = c.n_classes + _lib.MOTION_DIM
                    mn = (ops.predicate_head_affine(x_trk, self.packed_sub, c.n_predicates, raw=True, k_dim=kd,
                                                    background=True),
                          ops.predicate_head_affine(x_trk, self.packed_obj, c.n_predicates, raw=True, k_dim=kd,
                                                    background=True))
                else:
                    mn = ops.normalize_motion(batch.motion)
            if survivors:
                second.wait_event(topk_done)          # the bias rows need the terms and the top-K rows only
                row_bias = ops.gather_pair_terms(batch, row.reshape(-1), mn[0], mn[1])
                terms_done = torch.cuda.Event()
                terms_done.record(second)
                if rows_done is not None:
                    # temporal NMS + top-n of the survivors' decoded spans ([SPEC] s8), beside the predicate head
                    # and the records on the other stream; joined by _seg_tail
                    second.wait_event(rows_done)
                    sel, sel_counts = ops.span_select(sp, self._n_anchors(), c.anchor_stride, c.num_span_proposals,
                                                      c.nms_threshold, batch=batch, rows=row.reshape(-1))
        if terms_done is not None:
            cur.wait_event(terms_done)                # join the terms; the span selection keeps running
        else:
            cur.wait_stream(second)                   # join
        if not torch.cuda.is_current_stream_capturing():
            for u in (mn if isinstance(mn, tuple) else (mn,)) + (row_bias,):
                if u is not None:
                    u.record_stream(cur)              # allocated on the second side stream, consumed on this one
            if row is not None:
                row.record_stream(second)
            if sp is not None:
                sp.record_stream(second)
        if survivors:
            logits = ops.predicate_head_affine(rel16, self.packed_rel, c.n_predicates, bias=self.w[CLS_PREFIX + "bias"],
                                               row_bias=row_bias, background=True)
            records = counts = None
            if c.records:
                records, counts = ops.postprocess(batch, logits, None, c.topk_per_pair, c.topk_per_video,
                                                  rows=row.reshape(-1), row_video_off=self._row_offsets(batch),
                                                  mirror_q4=c.mirror_q4)
            early = {"logits": logits, "records": records, "counts": counts, "spans": sp, "sel": sel,
                     "sel_counts": sel_counts}
        return scores, idx, val, row, mn, early

    def _n_anchors(self) -> int:
        return int(self.w[DPN_PREFIX + "duration_pred.weight"].shape[0]) // 2

    def _row_offsets(self, batch: DeviceBatch) -> torch.Tensor:
        """First scored row of every video in sparsify mode ([V + 1] int64: v * K); one tensor per (table rows, K,
        device), kept for the life of the stage: captured graphs hold its address, and different capacity buckets
        have different numbers of table rows."""
        key = (batch.num_videos, self.cfg.topk, str(batch.device))
        off = self._row_off.get(key)
        if off is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("row offsets must be created before capture (GraphedStage warms the stage up first)")
            off = torch.arange(batch.num_videos + 1, dtype=torch.int64, device=batch.device) * self.cfg.topk
            self._row_off[key] = off
        return off

    def _decomposed(self, features) -> bool:
        """Tensor precision with features built on the GPU: the classifier is evaluated as
        ``A_s[s] + A_o[o] + rel(pair) W_rel^T`` and the ``[rows, F]`` feature matrix is never materialised
        (``StageConfig.materialize_features`` forces the rows into existence: one GEMM over F instead)."""
        return (features is None and self.cfg.precision == "tensor" and self.packed_rel is not None
                and not self.cfg.materialize_features)

    # The geometry call is three phases (include/tspn_b200.h): PRE (per-tracklet volumes; zeroing of the
    # per-pair sums when a video spans several chunks), MAIN (the pair kernel) and POST (vIoU / tIoU).  Only
    # MAIN is on the critical path: on a single-chunk batch PRE runs on the side stream under MAIN, and POST
    # always runs on the side stream under the tail (feature rows and records read the overlap windows, which
    # MAIN writes; vIoU / tIoU are outputs only).
    def _geo_alloc(self, batch: DeviceBatch, features: Optional[torch.Tensor], heads: bool = True):
        c = self.cfg
        need_geo = features is None or c.use_dpn
        if c.geo_layout not in ("dense", "windowed"):
            raise ValueError("geo_layout must be 'dense' or 'windowed' (got %r)" % (c.geo_layout,))
        windowed = c.geo_layout == "windowed" and need_geo and c.write_geo
        if windowed and heads and not self._survivor_path(batch, features, heads):
            raise ValueError("geo_layout='windowed': the kernels that read stored geometry rows read the dense layout; "
                             "the windowed rows need the survivor path (precision='tensor', sparsify, PPN on)")
        return ops.pair_geometry_outputs(batch, write_geo=need_geo and c.write_geo, windowed=windowed)

    def _seg_pre(self, batch: DeviceBatch, geom) -> None:
        ops.pair_geometry_phase(batch, geom, _lib.GEO_PHASE_PRE, clipped=self.cfg.viou_clipped)

    def _seg_geo(self, batch: DeviceBatch, geom, events=None, with_pre: bool = False, reserve_sms: int = 0):
        stream = torch.cuda.current_stream(batch.device)
        if with_pre:
            self._seg_pre(batch, geom)
        if events is not None:
            events[0].record(stream)
        ops.pair_geometry_phase(batch, geom, _lib.GEO_PHASE_MAIN, clipped=self.cfg.viou_clipped,
                                reserve_sms=reserve_sms)
        if events is not None:
            events[1].record(stream)
        return geom

    def _seg_post(self, batch: DeviceBatch, geom) -> None:
        """vIoU / tIoU on the second side stream, ordered after the pair kernel only (call right after _seg_geo,
        before the caller's stream joins the heads' chain); _seg_tail(post_done=True) joins it."""
        main = torch.cuda.current_stream(batch.device)
        second = self._side_stream(batch.device, 1)
        second.wait_stream(main)
        with torch.cuda.stream(second):
            ops.pair_geometry_phase(batch, geom, _lib.GEO_PHASE_POST, clipped=self.cfg.viou_clipped)

    def _seg_tail(self, batch: DeviceBatch, features, heads, side, geom, post_done: bool = False) -> StageResult:
        c = self.cfg
        scores, idx, val, row, mn, early = side
        k_eff = self.k_effective(batch)[:batch.num_real]
        sparsify = c.sparsify and c.use_ppn
        if early is not None:
            # survivor path: the heads already ran on the side branch; only the per-pair finalize is left
            if not post_done:
                self._seg_post(batch, geom)
            torch.cuda.current_stream(batch.device).wait_stream(self._side_stream(batch.device, 1))
            sp = early["spans"]
            if not torch.cuda.is_current_stream_capturing():
                for u in (early["sel"], early["sel_counts"]):
                    if u is not None:
                        u.record_stream(torch.cuda.current_stream(batch.device))
            return StageResult(self, batch, geom, scores, idx, val, row, rel_logits=early["logits"], sparsify=sparsify,
                               records=early["records"], record_counts=early["counts"],
                               span_buffers=[sp] if sp is not None else None, span_sel=early["sel"],
                               span_counts=early["sel_counts"])
        feats32 = feats16 = logits = None
        tensor = c.precision == "tensor"
        # The span head (HBM-bound reads of the surviving geometry rows) runs on the side stream underneath
        # the feature-row -> predicate-head chain (issue- / latency-bound), joined before the records.
        span_reg = spans = span_bufs = sel = sel_counts = None
        main = torch.cuda.current_stream(batch.device)
        side_stream = self._side_stream(batch.device)
        fork_spans = heads and c.use_dpn
        side_stream.wait_stream(main)
        with torch.cuda.stream(side_stream):
            ops.pair_geometry_phase(batch, geom, _lib.GEO_PHASE_POST, clipped=c.viou_clipped)
            if fork_spans:
                span_reg, spans, span_bufs, sel, sel_counts = self._span_heads(batch, geom, row, k_eff)
        decomposed = self._decomposed(features)
        if decomposed:
            rows = row.reshape(-1) if sparsify else None
            rel16, row_bias = ops.assemble_relative(batch, geom["geo"], geom["overlap"], rows, mn[0], mn[1])
            if heads:
                logits = ops.predicate_head_affine(rel16, self.packed_rel, c.n_predicates,
                                                   bias=self.w[CLS_PREFIX + "bias"], row_bias=row_bias)
        elif features is None:
            rows = row.reshape(-1) if sparsify else None
            feats32, feats16 = ops.assemble_features(batch, mn, geom["geo"], geom["overlap"], rows,
                                                     want_fp32=not tensor, want_bf16=tensor)
        else:
            feats32 = features
            if sparsify:
                sel = row.reshape(-1)
                feats32 = torch.where((sel >= 0)[:, None], features[sel.clamp_min(0)], features.new_zeros(()))
        if heads and not decomposed:
            x = feats16 if feats16 is not None else feats32
            logits = ops.predicate_head(x, self.w[CLS_PREFIX + "weight"], self.w[CLS_PREFIX + "bias"],
                                        precision=c.precision, packed=self.packed_cls)
        records = counts = None
        if heads and c.records:
            if sparsify:
                records, counts = ops.postprocess(batch, logits, geom["overlap"], c.topk_per_pair, c.topk_per_video,
                                                  rows=row.reshape(-1), row_video_off=self._row_offsets(batch),
                                                  mirror_q4=c.mirror_q4)
            else:
                records, counts = ops.postprocess(batch, logits, geom["overlap"], c.topk_per_pair, c.topk_per_video,
                                                  mirror_q4=c.mirror_q4)
        main.wait_stream(side_stream)
        if fork_spans and not torch.cuda.is_current_stream_capturing():
            for tns in (span_bufs or []) + (span_reg or []) + [u for u in (sel, sel_counts) if u is not None]:
                tns.record_stream(main)          # allocated on the side stream, read on the caller's
        return StageResult(self, batch, geom, scores, idx, val, row, feats32, feats16, logits, span_reg, spans,
                           sparsify, records, counts, span_bufs, sel, sel_counts)

    def forward(self, batch: DeviceBatch, features: Optional[torch.Tensor] = None,
                heads: bool = True, timers: Optional[dict] = None) -> StageResult:
        """``features``: optional precomputed ``[sum P, F]`` rows (reference mode: rows loaded from
        h5, lib/modeling/predict.py:42-57); when ``None`` they are constructed on the GPU.
        ``timers``: receives ``{"geo": (start, end)}`` CUDA events around the geometry kernel."""
        main = torch.cuda.current_stream(batch.device)
        side_stream = self._side_stream(batch.device)
        geom = self._geo_alloc(batch, features, heads)
        pre_aside = True                              # PRE under MAIN: every sum has a single writer (see _seg_geo)
        side_stream.wait_stream(main)                 # fork: inputs are ready on the caller's stream
        with torch.cuda.stream(side_stream):
            if pre_aside:
                self._seg_pre(batch, geom)
            side = self._seg_side(batch, features, heads)
        for t in side:
            for u in (t if isinstance(t, tuple) else tuple(t.values()) if isinstance(t, dict) else (t,)):
                if u is not None:
                    u.record_stream(main)             # allocated on the side stream, consumed on main
        events = None
        if timers is not None:      # CUDA events around the dominant kernel, on the launching stream
            events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            timers["geo"] = events
        early = side[5] is not None
        self._seg_geo(batch, geom, events=events, with_pre=not pre_aside,
                      reserve_sms=self.cfg.geo_reserve_sms if early else 0)
        if early:
            self._seg_post(batch, geom)               # under the heads' chain, not behind it
        main.wait_stream(side_stream)                 # join
        return self._seg_tail(batch, features, heads, side, geom, post_done=early)

    def capture(self, batch: DeviceBatch, features: Optional[torch.Tensor] = None,
                heads: bool = True, single: bool = True) -> "GraphedStage":
        """Freeze the step for this batch's shapes into a CUDA graph (see GraphedStage)."""
        return GraphedStage(self, batch, features, heads, single=single)

    def _span_heads(self, batch: DeviceBatch, geom, row, k_eff):
        """DPNHead + decode on the scored pairs of every video (rows gathered inside the kernel), from the stored
        geometry rows; with ``num_span_proposals`` the decoded spans then go through the temporal NMS + top-n."""
        c = self.cfg
        cw, cb = self.w[DPN_PREFIX + "conv.weight"], self.w[DPN_PREFIX + "conv.bias"]
        pw, pb = self.w[DPN_PREFIX + "duration_pred.weight"], self.w[DPN_PREFIX + "duration_pred.bias"]
        if cw.shape[1] != _lib.GEO_CHANNELS:
            raise ValueError("the pair stage feeds the span head with the %d geometry channels; "
                             "RELPN.DPN.IN_CHANNELS must be %d (got %d)" % (_lib.GEO_CHANNELS, _lib.GEO_CHANNELS,
                                                                            cw.shape[1]))
        regs, spans, bufs, sels, sel_counts = [], [], [], [], []
        th = batch.table_host
        real = list(range(batch.num_real))
        same_t = len(set(batch.t[v] for v in real)) == 1
        groups = [real] if same_t else [[v] for v in real]
        geo = geom["geo"]
        for vids in groups:
            v0 = vids[0]
            t, tp = batch.t[v0], int(th[v0, _lib.VT_TP])
            g0 = int(th[v0, _lib.VT_GEO_OFF])
            p0 = int(th[v0, _lib.VT_PAIR_OFF])
            p_cnt = sum(batch.n[v] * max(batch.n[v] - 1, 0) for v in vids)
            x = geo[g0:g0 + p_cnt * _lib.GEO_CHANNELS * tp].view(p_cnt, _lib.GEO_CHANNELS, tp)
            rsel = row[vids[0]:vids[-1] + 1].reshape(-1) if row is not None else None
            if c.keep_span_reg:
                reg = ops.span_head(x, cw, cb, pw, pb, rows=rsel, t=t, row_base=p0,
                                    precision=c.precision if cw.shape[1] >= 64 else "fp32")
                sp = ops.span_decode(reg, self.sizes_dev, c.anchor_stride)
            else:       # fused: the head only at the anchor columns, regressions stay in registers
                reg = None
                sp = ops.span_proposals(x, cw, cb, pw, pb, self.sizes_dev, c.anchor_stride, rows=rsel, t=t,
                                        row_base=p0)
            bufs.append(sp)
            if c.num_span_proposals > 0 and sp.shape[0] > 0:
                grows = rsel if rsel is not None else torch.arange(p0, p0 + p_cnt, dtype=torch.int64,
                                                                   device=batch.device)
                so, sc = ops.span_select(sp, pw.shape[0] // 2, c.anchor_stride, c.num_span_proposals, c.nms_threshold,
                                         batch=batch, rows=grows, windows=geom["overlap"][grows.clamp_min(0)])
                sels.append(so)
                sel_counts.append(sc)
            elif c.num_span_proposals > 0:
                sels.append(torch.zeros((0, c.num_span_proposals, 2), dtype=torch.int16, device=batch.device))
                sel_counts.append(torch.zeros(0, dtype=torch.int32, device=batch.device))
            if row is not None:
                k = row.shape[1]
                for j, v in enumerate(vids):
                    if reg is not None:
                        regs.append(reg[j * k:j * k + k_eff[v]])
                    spans.append(sp[j * k:j * k + k_eff[v]])
            else:
                off = 0
                for v in vids:
                    pv = batch.n[v] * max(batch.n[v] - 1, 0)
                    if reg is not None:
                        regs.append(reg[off:off + pv])
                    spans.append(sp[off:off + pv])
                    off += pv
        sel = torch.cat(sels, dim=0) if sels else None
        cnt = torch.cat(sel_counts, dim=0) if sel_counts else None
        return (regs if c.keep_span_reg else None), spans, bufs, sel, cnt


class GraphedStage:
    """The step of one fixed-shape batch as ONE CUDA graph (``single=True``, default) or as three
    (side / geo / tail, joined with stream events on the host).

    Launch-bound host work (17 C-ABI calls, their output allocations and tensor-map encodes) is paid
    once at capture.  In the single graph the side branches (stream 0: relationness -> top-K -> surviving-pair
    rows -> predicate head -> records; stream 1: per-tracklet predicate terms, volumes, vIoU finalize) are forks
    inside the graph, and the two events that time the all-pairs kernel are *external* event-record nodes
    (``torch.cuda.Event(external=True)``): every replay re-records them, so ``timers["geo"]`` must be read
    (after a synchronize) before the next replay.  The geometry outputs are allocated once, before the
    capture; all other outputs live in the graph's private memory pool; every replay overwrites both:
    ``result`` always refers to the latest one.  Refill the inputs with ``batch.copy_from(host)`` (same per-video shapes) between
    replays.
    """

    def __init__(self, stage: PairStage, batch: DeviceBatch, features: Optional[torch.Tensor] = None,
                 heads: bool = True, single: bool = True):
        self.stage, self.batch = stage, batch
        dev = batch.device
        self.side_stream = torch.cuda.Stream(dev, priority=side_priority())   # own stream: replays of different
                                                                              # slots may overlap
        self._cap_stream = torch.cuda.Stream(dev)
        stage.forward(batch, features=features, heads=heads)      # warm-up: lazy init outside capture
        torch.cuda.synchronize(dev)
        self.single = bool(single)
        pool = torch.cuda.graph_pool_handle()
        n0 = ops.launch_count()
        # the pair kernel's own timing: external event-record nodes inside the graph
        self.ev_geo = (torch.cuda.Event(enable_timing=True, external=True),
                       torch.cuda.Event(enable_timing=True, external=True))
        # the geometry outputs are allocated outside the graphs: the PRE phase (side branch) and the pair kernel
        # (main branch) both write into them
        geom = stage._geo_alloc(batch, features, heads)
        pre_aside = True
        if self.single:
            self.graph = torch.cuda.CUDAGraph()
            fork = stage._side_stream(dev)
            with torch.cuda.graph(self.graph, pool=pool, stream=self._cap_stream):
                cap = torch.cuda.current_stream(dev)
                fork.wait_stream(cap)
                with torch.cuda.stream(fork):
                    if pre_aside:
                        stage._seg_pre(batch, geom)
                    side = stage._seg_side(batch, features, heads)
                early = side[5] is not None
                stage._seg_geo(batch, geom, events=self.ev_geo, with_pre=not pre_aside,
                               reserve_sms=stage.cfg.geo_reserve_sms if early else 0)
                if early:
                    stage._seg_post(batch, geom)
                cap.wait_stream(fork)
                self.result = stage._seg_tail(batch, features, heads, side, geom, post_done=early)
        else:
            self.g_side, self.g_geo, self.g_tail = (torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(),
                                                    torch.cuda.CUDAGraph())
            # g_side and g_geo replay CONCURRENTLY (different streams): they must not share a memory pool, or a
            # temporary freed at the end of one capture is handed to the other and both write it at replay
            with torch.cuda.graph(self.g_side, stream=self._cap_stream):
                if pre_aside:
                    stage._seg_pre(batch, geom)
                side = stage._seg_side(batch, features, heads)
                torch.cuda.current_stream(dev).wait_stream(stage._side_stream(dev, 1))   # the span selection's join
            with torch.cuda.graph(self.g_geo, stream=self._cap_stream):
                stage._seg_geo(batch, geom, events=self.ev_geo, with_pre=not pre_aside,
                               reserve_sms=stage.cfg.geo_reserve_sms if side[5] is not None else 0)
            with torch.cuda.graph(self.g_tail, stream=self._cap_stream):
                self.result = stage._seg_tail(batch, features, heads, side, geom)
        self.kernels_per_replay = ops.launch_count() - n0
        self.graph_launches = 1 if self.single else 3
        torch.cuda.synchronize(dev)

    def replay(self, timers: Optional[dict] = None) -> StageResult:
        main = torch.cuda.current_stream(self.batch.device)
        if self.single:
            self.graph.replay()
            if timers is not None:
                timers["geo"] = self.ev_geo          # valid until the next replay
            ops.count_launches(self.kernels_per_replay)
            return self.result
        self.side_stream.wait_stream(main)
        with torch.cuda.stream(self.side_stream):
            self.g_side.replay()
        self.g_geo.replay()
        if timers is not None:
            timers["geo"] = self.ev_geo          # external event nodes inside g_geo; valid until the next replay
        main.wait_stream(self.side_stream)
        self.g_tail.replay()
        ops.count_launches(self.kernels_per_replay)
        return self.result


def stage_from_videos(videos, weights, config: StageConfig, device="cuda"):
    """Convenience: pack ``synth.VideoTracklets`` and build a ready stage."""
    host = HostBatch.from_videos(videos)
    stage = PairStage(config)
    stage.load_weights(weights, device)
    return stage, host
