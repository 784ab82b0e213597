"""CPU oracle for the rows after the pair stage: relation evaluation (N3) and
greedy relational association (N2).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, one trajectory pair at a time exactly as the reference walks them,
``lib/evaluation/visual_relation_detection.py:8-123`` (+ ``voc_ap``, ``common.py:4-37``) and
``lib/modeling/association.py:16-175``.  Pinned by ``tests/golden/relations_outputs.npz`` - outputs of
the unmodified reference on the seeded cases of ``tspn_b200.synth`` (``make_golden_relations.py``);
``tests/test_oracle_golden.py`` checks this module against them.

The product (``tspn_b200.evaluation`` / ``tspn_b200.association``) computes the same quantities from
vIoU values batched on the GPU; this file deliberately keeps the reference's lazy, per-pair structure so
that the batching itself is what the parity tests exercise.
"""
from __future__ import annotations

import numpy as np

from .geometry import traj_iou_clipped_ref, viou_ref


# ---- evaluation -----------------------------------------------------------------------------------
def voc_ap(rec, prec):
    """Area under the precision envelope (common.py:19-36, the non-07 branch)."""
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = max(mpre[i - 1], mpre[i])
    idx = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[idx + 1] - mrec[idx]) * mpre[idx + 1])


def _pr_curves(hit_scores, n_gt):
    tp = np.isfinite(hit_scores)
    ctp = np.cumsum(tp).astype(np.float32)
    cfp = np.cumsum(~tp).astype(np.float32)
    eps = np.finfo(np.float32).eps
    return ctp / np.maximum(ctp + cfp, eps), ctp / np.maximum(n_gt, eps), hit_scores


def eval_detection_scores(gt_relations, pred_relations, viou_threshold):
    """visual_relation_detection.py:8-36: predictions in descending score order (stable); each claims the
    not-yet-detected ground truth of equal triplet with the largest ``min(vIoU_sub, vIoU_obj)`` that
    reaches the threshold (strict ``>`` keeps the first of equals)."""
    order = sorted(range(len(pred_relations)), key=lambda i: pred_relations[i]['score'], reverse=True)
    taken = [False] * len(gt_relations)
    hit = np.full(len(pred_relations), -np.inf)
    for rank, pi in enumerate(order):
        p = pred_relations[pi]
        best, best_k = -float('inf'), -1
        for k, g in enumerate(gt_relations):
            if taken[k] or tuple(p['triplet']) != tuple(g['triplet']):
                continue
            ov = min(viou_ref(p['sub_traj'], p['duration'], g['sub_traj'], g['duration']),
                     viou_ref(p['obj_traj'], p['duration'], g['obj_traj'], g['duration']))
            if ov >= viou_threshold and ov > best:
                best, best_k = ov, k
        if best_k >= 0:
            hit[rank] = p['score']
            taken[best_k] = True
    return _pr_curves(hit, len(gt_relations))


def eval_tagging_scores(gt_relations, pred_relations):
    """visual_relation_detection.py:39-61: first occurrence of each predicted triplet, in score order."""
    ordered = sorted(pred_relations, key=lambda x: x['score'], reverse=True)
    truth = {tuple(r['triplet']) for r in gt_relations}
    seen, hit = [], []
    for r in ordered:
        t = tuple(r['triplet'])
        if t not in seen:
            seen.append(t)
            hit.append(r['score'] if t in truth else -np.inf)
    return _pr_curves(np.asarray(hit, dtype=np.float64), len(truth))


def evaluate(groundtruth, prediction, viou_threshold=0.5, det_nreturns=(50, 100, 1000), tag_nreturns=(1, 5, 10)):
    """visual_relation_detection.py:64-123 -> ``(mean_ap, rec_at_n, mprec_at_n)``."""
    aps, n_gt = [], 0
    det_scores = {n: [] for n in det_nreturns}
    det_tp = {n: [] for n in det_nreturns}
    tag_prec = {n: [] for n in tag_nreturns}
    for vid, gts in groundtruth.items():
        if len(gts) == 0:
            continue
        n_gt += len(gts)
        prec, rec, hit = eval_detection_scores(gts, prediction[vid], viou_threshold)
        aps.append(voc_ap(rec, prec))
        for n in det_nreturns:
            det_scores[n].append(hit[:min(n, hit.size)])
            det_tp[n].append(np.isfinite(hit)[:min(n, hit.size)])
        tprec, _, _ = eval_tagging_scores(gts, prediction[vid])
        for n in tag_nreturns:
            cut = min(n, tprec.size)
            tag_prec[n].append(tprec[cut - 1] if cut > 0 else 0.)
    rec_at_n = {}
    for n in det_nreturns:
        s, t = np.concatenate(det_scores[n]), np.concatenate(det_tp[n])
        ctp = np.cumsum(t[np.argsort(s)[::-1]]).astype(np.float32)
        rec_at_n[n] = (ctp / np.maximum(n_gt, np.finfo(np.float32).eps))[-1]
    return np.mean(aps), rec_at_n, {n: np.mean(tag_prec[n]) for n in tag_nreturns}


# ---- association ----------------------------------------------------------------------------------
class Traj:
    """The three things association.py touches on a Trajectory: ``pstart``, ``pend`` and the box list.
    Instances are shared between relations exactly as the reference shares its Trajectory objects
    (``straj = trajs[s_tididx]``, association.py:147), so an in-place merge is seen by every holder."""

    def __init__(self, pstart, pend, rois):
        self.pstart, self.pend = int(pstart), int(pend)
        self.rois = [tuple(float(c) for c in r) for r in rois]


def merge_trajs(t1: Traj, t2: Traj) -> Traj:
    """association.py:16-32: average the boxes over ``t1``'s last ``t1.pend - t2.pstart`` frames with
    ``t2``'s first ones, then append the rest of ``t2`` (in place on ``t1``)."""
    ov = max(t1.pend - t2.pstart, 0)
    n1 = t1.pend - t1.pstart
    for i in range(ov):
        a, b = t1.rois[n1 - ov + i], t2.rois[i]
        t1.rois[n1 - ov + i] = tuple((x + y) / 2 for x, y in zip(a, b))
    for i in range(ov, t2.pend - t2.pstart):
        t1.rois.append(t2.rois[i])
        t1.pend += 1
    return t1


def clipped_iou(t1: Traj, t2: Traj) -> float:
    """association.py:35-48 (float32 sequential accumulation, as the reference's cubic_iou)."""
    if t1.pend <= t2.pstart or t2.pend <= t1.pstart:
        return 0
    return traj_iou_clipped_ref(np.asarray(t1.rois), (t1.pstart, t1.pend), np.asarray(t2.rois), (t2.pstart, t2.pend))


class Relation:
    def __init__(self, triplet, straj, otraj, conf=1):
        self.triplet, self.straj, self.otraj = tuple(int(x) for x in triplet), straj, otraj
        self.confs = [conf]
        self.fstart, self.fend = straj.pstart, straj.pend


def greedy_relational_association(short_term_relations, segment_trajs, max_traj_num_in_clip=100, iou_thr=0.5):
    """association.py:117-175 with ``segment_trajs[(vid, fstart, fend)]`` standing in for the on-disk
    trajectory proposals.  Returns one dict per video relation: ``triplet`` (class ids), ``score``,
    ``duration``, ``sub_traj``, ``obj_traj``.  Reference behaviours kept: a relation that starts in a
    later segment is created with confidence 1 instead of its score (association.py:169), ``fend``
    follows the OBJECT trajectory after a merge (:98), trajectory objects are shared and merged in place."""
    segs = sorted(short_term_relations, key=lambda x: int(x[0][1]))
    relations, last = [], []
    for i, (index, (pred_list, _iou, _trackid)) in enumerate(segs):
        vid, fstart, fend = index
        preds = sorted(pred_list, key=lambda x: x[0], reverse=True)[:max_traj_num_in_clip]
        trajs = [Traj(fstart, fend, t['rois']) for t in segment_trajs[(vid, fstart, fend)]]
        cur = []
        for score, triplet, tids in preds:
            straj, otraj = trajs[int(tids[0])], trajs[int(tids[1])]
            merged = False
            if i > 0:
                last.sort(key=lambda r: np.mean(r.confs), reverse=True)
                for r in last:
                    if tuple(int(x) for x in triplet) != r.triplet:
                        continue
                    if straj.pstart < r.fend and otraj.pstart < r.fend \
                            and clipped_iou(r.straj, straj) >= iou_thr and clipped_iou(r.otraj, otraj) >= iou_thr:
                        r.straj = merge_trajs(r.straj, straj)
                        r.otraj = merge_trajs(r.otraj, otraj)
                        r.confs.append(score)
                        r.fstart, r.fend = r.straj.pstart, r.otraj.pend
                        last.remove(r)
                        cur.append(r)
                        merged = True
                        break
            if not merged:
                r = Relation(triplet, straj, otraj, score if i == 0 else 1)
                relations.append(r)
                cur.append(r)
        last = cur
    return [dict(triplet=r.triplet, score=float(np.mean(r.confs)), duration=[int(r.fstart), int(r.fend)],
                 sub_traj=list(r.straj.rois), obj_traj=list(r.otraj.rois)) for r in relations]
