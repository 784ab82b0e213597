"""Generate tests/golden/relations_outputs.npz by running the UNMODIFIED reference's evaluation and
association code on seeded synthetic relations.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden_relations.py

Executed from /root/reference: ``evaluate`` / ``eval_detection_scores`` / ``eval_tagging_scores``
(lib/evaluation/visual_relation_detection.py:8-123) with ``viou`` / ``voc_ap`` (lib/evaluation/common.py)
and ``greedy_relational_association`` with ``VideoRelation`` / ``_merge_trajs`` / ``_traj_iou``
(lib/modeling/association.py:16-175).  Stubs: ``dlib`` (the ``drectangle`` accessor type only),
``IPython.embed``; ``object_trajectory_proposal`` (a JSON file reader, lib/modeling/trajectory.py:161-180)
is replaced by a lookup into the synthetic per-segment trajectories.  Only OUTPUTS are stored; the
inputs are regenerated from the seeds by ``tspn_b200.synth``.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _install_stubs, REF, ROOT  # noqa: E402,F401

EVAL_CASES = {"int": dict(seed=0, integer_boxes=True), "int2": dict(seed=5, integer_boxes=True, n_videos=4, max_pred=90),
              "frac": dict(seed=1, integer_boxes=False)}
ASSOC_CASES = {"a": dict(seed=0, n_segments=6, preds_per_segment=30), "b": dict(seed=3, n_segments=5, n_objects=4,
                                                                                preds_per_segment=24)}


class NameOnlyDataset:
    """``dataset`` argument of greedy_relational_association: only the two name lookups are used."""

    def get_object_name(self, cid):
        return "obj%d" % int(cid)

    def get_predicate_name(self, pid):
        return "pred%d" % int(pid)


def pack_relations(rels):
    """Serialized video relations -> flat arrays (triplet ids, score, duration, trajectories)."""
    trip = np.array([[int(r["triplet"][0][3:]), int(r["triplet"][1][4:]), int(r["triplet"][2][3:])] for r in rels],
                    dtype=np.int64).reshape(-1, 3)
    score = np.array([r["score"] for r in rels], dtype=np.float64)
    dur = np.array([r["duration"] for r in rels], dtype=np.int64).reshape(-1, 2)
    lens = np.array([[len(r["sub_traj"]), len(r["obj_traj"])] for r in rels], dtype=np.int64).reshape(-1, 2)
    boxes = [np.asarray(r[k], dtype=np.float64).reshape(-1, 4) for r in rels for k in ("sub_traj", "obj_traj")]
    boxes = np.concatenate(boxes, axis=0) if boxes else np.zeros((0, 4))
    return trip, score, dur, lens, boxes


def main():
    _install_stubs()
    from lib.evaluation.visual_relation_detection import evaluate, eval_detection_scores, eval_tagging_scores
    from lib.evaluation.common import viou
    import lib.modeling.association as assoc
    from lib.modeling.trajectory import Trajectory
    from tspn_b200 import synth

    out = {}
    for tag, kw in EVAL_CASES.items():
        gt, pred = synth.make_relation_eval_case(**kw)
        with contextlib.redirect_stdout(io.StringIO()):
            mean_ap, rec_at_n, mprec_at_n = evaluate(gt, pred)
        out[f"eval_{tag}_mean_ap"] = np.float64(mean_ap)
        out[f"eval_{tag}_rec_at_n"] = np.array([rec_at_n[k] for k in (50, 100, 1000)], dtype=np.float64)
        out[f"eval_{tag}_mprec_at_n"] = np.array([mprec_at_n[k] for k in (1, 5, 10)], dtype=np.float64)
        hits, precs, recs, tags, sizes = [], [], [], [], []
        for vid in gt:
            if not gt[vid]:
                continue
            p, r, h = eval_detection_scores(gt[vid], pred[vid], 0.5)
            tp, _, th = eval_tagging_scores(gt[vid], pred[vid])
            hits.append(h), precs.append(p), recs.append(r), tags.append(tp)
            sizes.append((len(h), len(tp)))
        out[f"eval_{tag}_sizes"] = np.array(sizes, dtype=np.int64)
        out[f"eval_{tag}_hit_scores"] = np.concatenate(hits)
        out[f"eval_{tag}_prec"] = np.concatenate(precs)
        out[f"eval_{tag}_rec"] = np.concatenate(recs)
        out[f"eval_{tag}_tag_prec"] = np.concatenate(tags)
        # a lower threshold changes which ground truth each prediction claims
        p, r, h = eval_detection_scores(gt["video_001"], pred["video_001"], 0.2)
        out[f"eval_{tag}_hit_scores_thr02_video1"] = h
        # the raw vIoU values of the first video's equal-triplet pairs, in (prediction, ground truth) order
        vals = []
        for pr in sorted(pred["video_001"], key=lambda x: x["score"], reverse=True):
            for g in gt["video_001"]:
                if tuple(pr["triplet"]) == tuple(g["triplet"]):
                    vals.append((viou(pr["sub_traj"], pr["duration"], g["sub_traj"], g["duration"]),
                                 viou(pr["obj_traj"], pr["duration"], g["obj_traj"], g["duration"])))
        out[f"eval_{tag}_viou_video1"] = np.array(vals, dtype=np.float64).reshape(-1, 2)

    for tag, kw in ASSOC_CASES.items():
        short_term, seg_trajs = synth.make_association_case(**kw)
        assoc.object_trajectory_proposal = lambda dataset, vid, fstart, fend: [
            Trajectory(**t) for t in seg_trajs[(vid, fstart, fend)]]
        rels = assoc.greedy_relational_association(NameOnlyDataset(), short_term, max_traj_num_in_clip=100)
        trip, score, dur, lens, boxes = pack_relations(rels)
        out[f"assoc_{tag}_triplet"], out[f"assoc_{tag}_score"] = trip, score
        out[f"assoc_{tag}_duration"], out[f"assoc_{tag}_lens"], out[f"assoc_{tag}_boxes"] = dur, lens, boxes
        # cap on predictions per clip (association.py:126-128)
        rels = assoc.greedy_relational_association(NameOnlyDataset(), short_term, max_traj_num_in_clip=7)
        trip, score, dur, lens, boxes = pack_relations(rels)
        out[f"assoc_{tag}_cap7_triplet"], out[f"assoc_{tag}_cap7_score"] = trip, score
        out[f"assoc_{tag}_cap7_duration"] = dur
        print(tag, "relations", len(score), "merged", int((dur[:, 1] - dur[:, 0] > 30).sum()))
        # evaluation of association output: box lists longer than their durations (common.py:100-105)
        rels = assoc.greedy_relational_association(NameOnlyDataset(), short_term, max_traj_num_in_clip=100)
        rels = [dict(r, sub_traj=[list(b) for b in r["sub_traj"]], obj_traj=[list(b) for b in r["obj_traj"]])
                for r in rels]
        gt = {"synthetic_video": synth.make_gt_from_relations(rels, seed=kw["seed"])}
        pred = {"synthetic_video": rels}
        with contextlib.redirect_stdout(io.StringIO()):
            mean_ap, rec_at_n, mprec_at_n = evaluate(gt, pred)
        p_, r_, h_ = eval_detection_scores(gt["synthetic_video"], rels, 0.5)
        out[f"assoc_{tag}_eval_mean_ap"] = np.float64(mean_ap)
        out[f"assoc_{tag}_eval_rec_at_n"] = np.array([rec_at_n[k] for k in (50, 100, 1000)], dtype=np.float64)
        out[f"assoc_{tag}_eval_hit_scores"] = h_
        longer = sum(len(r["sub_traj"]) != r["duration"][1] - r["duration"][0] for r in rels)
        print(tag, "eval of association output: mAP", mean_ap, "hits", int(np.isfinite(h_).sum()), "of", len(h_),
              "relations with list != duration:", longer)

    path = os.path.join(HERE, "relations_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
