"""Video visual relation evaluation — mirror of lib/evaluation/visual_relation_detection.py and
lib/evaluation/common.py with the trajectory vIoU on the GPU.

The reference calls the pure-python ``viou`` (common.py:65-106) once per (prediction, ground truth)
pair of equal triplet, twice (subject and object trajectory), inside the greedy matching loop
(visual_relation_detection.py:8-36).  None of those values depends on the state of the matching, so
here every candidate pair of every video of the call is scored by ONE launch of
``tspn_viou_pairs_f64`` (per-trajectory volumes summed once, fp64 sums) and the matching loop then
walks the precomputed values in the reference's order: the threshold test ``ov >= viou_threshold``,
the strict arg-max ``ov > ov_max`` and the first-come tie rule are unchanged, so ``hit_scores`` and
everything derived from them (AP, recall@N, precision@N) are identical for integer boxes.

Signatures, argument meaning and return types follow the reference.  Relations are the dicts of the
prediction / ground-truth JSON (lib/evaluation/README.md): ``triplet``, ``score`` (predictions),
``duration`` = [fstart, fend), ``sub_traj`` / ``obj_traj`` = one [x1, y1, x2, y2] per frame.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .trajectory import viou, viou_batch  # noqa: F401  (viou re-exported like common.py)


# ---- common.py ---------------------------------------------------------------------------------
def voc_ap(rec, prec, use_07_metric=False):
    """VOC average precision of a precision/recall curve (common.py:4-37)."""
    rec, prec = np.asarray(rec), np.asarray(prec)
    if use_07_metric:
        total = 0.
        for thr in np.arange(0., 1.1, 0.1):
            sel = rec >= thr
            total = total + (np.max(prec[sel]) if np.sum(sel) != 0 else 0) / 11.
        return total
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]                 # precision envelope
    step = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[step + 1] - mrec[step]) * mpre[step + 1])


def iou(bbox_1, bbox_2):
    """IoU of two inclusive-pixel boxes (common.py:40-62)."""
    area_1 = (bbox_1[2] - bbox_1[0] + 1) * (bbox_1[3] - bbox_1[1] + 1)
    area_2 = (bbox_2[2] - bbox_2[0] + 1) * (bbox_2[3] - bbox_2[1] + 1)
    ow = max(0, min(bbox_1[2], bbox_2[2]) - max(bbox_1[0], bbox_2[0]) + 1)
    oh = max(0, min(bbox_1[3], bbox_2[3]) - max(bbox_1[1], bbox_2[1]) + 1)
    return ow * oh * 1.0 / (area_1 + area_2 - ow * oh)


# ---- batched candidate scoring --------------------------------------------------------------------
def _sorted_predictions(pred_relations):
    return sorted(pred_relations, key=lambda x: x['score'], reverse=True)      # stable, like the reference


def candidate_overlaps(jobs: Sequence[Tuple[list, list]]) -> List[Dict[int, List[Tuple[int, float]]]]:
    """``jobs[j] = (gt_relations, pred_relations_sorted)``.  Returns per job ``{pred_idx: [(gt_idx, ov)]}``
    in ground-truth order, ``ov = min(viou(subjects), viou(objects))`` for every pair of equal triplet —
    the quantity visual_relation_detection.py:17-21 computes, for all jobs in one GPU launch."""
    trajs, durs = [], []
    cand = []                    # (job, pred_idx, gt_idx)
    pa, pb = [], []

    def add(rel, key):
        trajs.append(rel[key])
        durs.append(tuple(rel['duration']))
        return len(trajs) - 1

    for j, (gts, preds) in enumerate(jobs):
        by_triplet = defaultdict(list)
        for gi, g in enumerate(gts):
            by_triplet[tuple(g['triplet'])].append(gi)
        gt_slot: Dict[int, Tuple[int, int]] = {}
        for pi, p in enumerate(preds):
            hits = by_triplet.get(tuple(p['triplet']))
            if not hits:
                continue
            ps, po = add(p, 'sub_traj'), add(p, 'obj_traj')
            for gi in hits:
                if gi not in gt_slot:
                    gt_slot[gi] = (add(gts[gi], 'sub_traj'), add(gts[gi], 'obj_traj'))
                cand.append((j, pi, gi))
                pa += [ps, po]
                pb += [gt_slot[gi][0], gt_slot[gi][1]]
    out: List[Dict[int, List[Tuple[int, float]]]] = [defaultdict(list) for _ in jobs]
    if not cand:
        return out
    v = viou_batch(trajs, durs, np.stack([pa, pb], axis=1), f64=True)
    ov = np.minimum(v[0::2], v[1::2])
    for (j, pi, gi), o in zip(cand, ov):
        out[j][pi].append((gi, float(o)))
    return out


def _detection_from_overlaps(gt_relations, pred_relations, overlaps, viou_threshold):
    gt_detected = np.zeros((len(gt_relations),), dtype=bool)
    hit_scores = np.ones((len(pred_relations))) * -np.inf
    for pred_idx, pred_relation in enumerate(pred_relations):
        ov_max, k_max = -float('Inf'), -1
        for gt_idx, ov in overlaps.get(pred_idx, ()):
            if not gt_detected[gt_idx] and ov >= viou_threshold and ov > ov_max:
                ov_max, k_max = ov, gt_idx
        if k_max >= 0:
            hit_scores[pred_idx] = pred_relation['score']
            gt_detected[k_max] = True
    return _curves(hit_scores, len(gt_relations))


def _curves(hit_scores, n_gt):
    tp = np.isfinite(hit_scores)
    cum_tp = np.cumsum(tp).astype(np.float32)
    cum_fp = np.cumsum(~tp).astype(np.float32)
    rec = cum_tp / np.maximum(n_gt, np.finfo(np.float32).eps)
    prec = cum_tp / np.maximum(cum_tp + cum_fp, np.finfo(np.float32).eps)
    return prec, rec, hit_scores


# ---- visual_relation_detection.py -----------------------------------------------------------------
def eval_detection_scores(gt_relations, pred_relations, viou_threshold):
    """Greedy matching of score-sorted predictions to ground truth (visual_relation_detection.py:8-36);
    returns ``(prec, rec, hit_scores)``."""
    preds = _sorted_predictions(pred_relations)
    overlaps = candidate_overlaps([(gt_relations, preds)])[0]
    return _detection_from_overlaps(gt_relations, preds, overlaps, viou_threshold)


def eval_tagging_scores(gt_relations, pred_relations):
    """Relation tagging: trajectories ignored (visual_relation_detection.py:39-61)."""
    preds = _sorted_predictions(pred_relations)
    gt_triplets = set(tuple(r['triplet']) for r in gt_relations)
    seen, hit_scores = {}, []
    for r in preds:
        triplet = tuple(r['triplet'])
        if triplet not in seen:
            seen[triplet] = len(hit_scores)
            hit_scores.append(r['score'] if triplet in gt_triplets else -np.inf)
    return _curves(np.asarray(hit_scores, dtype=np.float64), len(gt_triplets))


def evaluate(groundtruth, prediction, viou_threshold=0.5, det_nreturns=[50, 100, 1000], tag_nreturns=[1, 5, 10],
             verbose=True):
    """Detection mean AP, recall@N and tagging precision@N over all videos
    (visual_relation_detection.py:64-123); one vIoU launch for the whole call."""
    vids = [vid for vid, gts in groundtruth.items() if len(gts) != 0]
    if verbose:
        print('Computing average precision AP over {} videos...'.format(len(groundtruth)))
    jobs = [(groundtruth[vid], _sorted_predictions(prediction[vid])) for vid in vids]
    overlaps = candidate_overlaps(jobs)
    video_ap = dict()
    tot_scores, tot_tp, prec_at_n = defaultdict(list), defaultdict(list), defaultdict(list)
    tot_gt_relations = 0
    for vid, (gts, preds), ov in zip(vids, jobs, overlaps):
        tot_gt_relations += len(gts)
        det_prec, det_rec, det_scores = _detection_from_overlaps(gts, preds, ov, viou_threshold)
        video_ap[vid] = voc_ap(det_rec, det_prec)
        tp = np.isfinite(det_scores)
        for nre in det_nreturns:
            cut_off = min(nre, det_scores.size)
            tot_scores[nre].append(det_scores[:cut_off])
            tot_tp[nre].append(tp[:cut_off])
        tag_prec, _, _ = eval_tagging_scores(gts, preds)
        for nre in tag_nreturns:
            cut_off = min(nre, tag_prec.size)
            prec_at_n[nre].append(tag_prec[cut_off - 1] if cut_off > 0 else 0.)
    mean_ap = np.mean(list(video_ap.values()))
    rec_at_n = dict()
    for nre in det_nreturns:
        scores = np.concatenate(tot_scores[nre])
        tps = np.concatenate(tot_tp[nre])
        tps = tps[np.argsort(scores)[::-1]]
        cum_tp = np.cumsum(tps).astype(np.float32)
        rec_at_n[nre] = (cum_tp / np.maximum(tot_gt_relations, np.finfo(np.float32).eps))[-1]
    mprec_at_n = {nre: np.mean(prec_at_n[nre]) for nre in tag_nreturns}
    if verbose:
        print('detection mean AP (used in challenge): {}'.format(mean_ap))
        for nre in det_nreturns:
            print('detection recall@{}: {}'.format(nre, rec_at_n[nre]))
        for nre in tag_nreturns:
            print('tagging precision@{}: {}'.format(nre, mprec_at_n[nre]))
    return mean_ap, rec_at_n, mprec_at_n
