"""Greedy relational association — mirror of lib/modeling/association.py with the trajectory IoUs of a
whole segment batched onto the GPU.

The reference merges the short-term relations of consecutive 30-frame segments (15 frames apart) into
video-level relations: for every prediction of segment i, in descending score order, it scans the
relations touched in segment i-1 (descending mean confidence) for one with the same triplet whose subject
AND object trajectories overlap the new ones with overlap-clipped vIoU >= 0.5 (``_traj_iou``,
association.py:35-48: two ``deepcopy``s + ``islice`` + numpy per call).  Those IoUs only depend on
trajectory contents, so here all candidate (relation, prediction) trajectory pairs of a segment are
scored by ONE ``tspn_viou_pairs_f64`` launch (flag ``TSPN_VIOU_CLIPPED``) before the greedy loop runs.

One subtlety is kept exactly: the reference shares Trajectory objects between relations
(``straj = trajs[s_tididx]``, association.py:147) and ``_merge_trajs`` edits them in place, so a merge can
change a trajectory another relation still holds.  Every ``Trajectory`` carries a version counter; a
precomputed IoU is used only while both versions are unchanged, otherwise that pair is rescored.

Reference behaviours mirrored (see ``oracle/relations.py``, pinned by the golden file): relations started
after the first segment get confidence 1 instead of their score (association.py:169); ``fend`` follows the
object trajectory after ``extend`` (:98).  Values: the reference accumulates in float32, the GPU in fp64 -
agreement ~1e-6 relative, identical merge decisions unless an IoU lies that close to the threshold.
"""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .trajectory import Trajectory, viou_batch


def get_segment_signature(vid, fstart, fend):
    """lib/modeling/__init__.py:5-9."""
    return '{}-{:04d}-{:04d}'.format(vid, fstart, fend)


def object_trajectory_proposal(dataset, vid, fstart, fend, gt=False, root='./vidvrd-baseline-output'):
    """Read one segment's trajectory proposals (lib/modeling/trajectory.py:161-180): the JSON list written
    next to the features, ``<root>/features/<name>/<vid>/<vsig>-<name>.json``; [] when absent."""
    name = 'traj_cls_gt' if gt else 'traj_cls'
    path = os.path.join(root, 'features', name, vid, '{}-{}.json'.format(get_segment_signature(vid, fstart, fend), name))
    if not os.path.exists(path):
        return []
    with open(path, 'r') as fin:
        return [Trajectory(**traj) for traj in json.load(fin)]


def _touch(traj: Trajectory) -> None:
    traj._version = getattr(traj, '_version', 0) + 1


def _merge_trajs(traj_1: Trajectory, traj_2: Trajectory) -> Trajectory:
    """association.py:16-32, in place on ``traj_1``."""
    if not (traj_1.pend > traj_2.pstart and traj_1.pstart < traj_2.pend):
        print('{}-{} {}-{}'.format(traj_1.pstart, traj_1.pend, traj_2.pstart, traj_2.pend))
    overlap_length = max(traj_1.pend - traj_2.pstart, 0)
    base = traj_1.length() - overlap_length
    for i in range(overlap_length):
        a, b = traj_1.rois[base + i], traj_2.rois[i]
        traj_1.rois[base + i] = ((a[0] + b[0]) / 2, (a[1] + b[1]) / 2, (a[2] + b[2]) / 2, (a[3] + b[3]) / 2)
    for i in range(overlap_length, traj_2.length()):
        traj_1.predict(traj_2.rois[i])
    _touch(traj_1)
    return traj_1


def _key(a: Trajectory, b: Trajectory) -> Tuple[int, int, int, int]:
    return (id(a), getattr(a, '_version', 0), id(b), getattr(b, '_version', 0))


def batched_traj_iou(pairs: List[Tuple[Trajectory, Trajectory]]) -> Dict[Tuple[int, int, int, int], float]:
    """Overlap-clipped vIoU (association.py:35-48) of many trajectory pairs in one launch; keyed by object
    identity + version of both trajectories."""
    pool: Dict[int, int] = {}
    trajs: List[Trajectory] = []
    idx, keys = [], []
    for a, b in pairs:
        k = _key(a, b)
        if a.pend <= b.pstart or b.pend <= a.pstart:
            continue                                         # no temporal overlap: 0 without a launch
        for t in (a, b):
            if id(t) not in pool:
                pool[id(t)] = len(trajs)
                trajs.append(t)
        idx.append((pool[id(a)], pool[id(b)]))
        keys.append(k)
    out: Dict[Tuple[int, int, int, int], float] = {}
    if idx:
        vals = viou_batch([np.asarray(t.rois, dtype=np.float32) for t in trajs], [(t.pstart, t.pend) for t in trajs],
                          np.asarray(idx, dtype=np.int32), clipped=True, f64=True)
        out = {k: float(v) for k, v in zip(keys, vals)}
    return out


def _traj_iou(traj_1: Trajectory, traj_2: Trajectory, cache: Optional[dict] = None):
    """association.py:35-48; ``cache`` holds batched values (``batched_traj_iou``)."""
    if traj_1.pend <= traj_2.pstart or traj_2.pend <= traj_1.pstart:
        return 0
    k = _key(traj_1, traj_2)
    if cache is not None and k in cache:
        return cache[k]
    val = batched_traj_iou([(traj_1, traj_2)])[k]
    if cache is not None:
        cache[k] = val
    return val


class VideoRelation(object):
    """Video-level relation instance (association.py:51-114)."""

    def __init__(self, vid, s_cid, pid, o_cid, straj, otraj, confs=1):
        self.vid, self.s_cid, self.pid, self.o_cid = vid, s_cid, pid, o_cid
        self.straj, self.otraj = straj, otraj
        self.confs_list = [confs]
        self.fstart, self.fend = straj.pstart, straj.pend

    def __repr__(self):
        return '<VideoRelation {}[{:04d}-{:04d}] {}-{}-{}>'.format(
            self.vid, self.fstart, self.fend, self.s_cid, self.pid, self.o_cid)

    def triplet(self):
        return (self.s_cid, self.pid, self.o_cid)

    def mean_confs(self):
        return np.mean(self.confs_list)

    def both_overlap(self, straj, otraj, iou_thr=0.5, cache=None):
        return bool(_traj_iou(self.straj, straj, cache) >= iou_thr and _traj_iou(self.otraj, otraj, cache) >= iou_thr)

    def extend(self, straj, otraj, confs):
        self.straj = _merge_trajs(self.straj, straj)
        self.otraj = _merge_trajs(self.otraj, otraj)
        self.confs_list.append(confs)
        self.fstart, self.fend = self.straj.pstart, self.otraj.pend

    def serialize(self, dataset):
        return {'triplet': [dataset.get_object_name(self.s_cid), dataset.get_predicate_name(self.pid),
                            dataset.get_object_name(self.o_cid)],
                'score': float(self.mean_confs()),
                'duration': [int(self.fstart), int(self.fend)],
                'sub_traj': self.straj.serialize()['rois'],
                'obj_traj': self.otraj.serialize()['rois']}


def greedy_relational_association(dataset, short_term_relations, max_traj_num_in_clip=100,
                                  trajectory_proposal: Optional[Callable] = None):
    """association.py:117-175.  ``short_term_relations``: ``[((vid, fstart, fend), (pred_list, iou,
    trackid))]`` as predict.py:110-117 builds them; ``trajectory_proposal(dataset, vid, fstart, fend)``
    supplies a segment's ``Trajectory`` list (default: the reference's on-disk JSON files).  Returns the
    serialized video relations."""
    load = trajectory_proposal or object_trajectory_proposal
    short_term_relations.sort(key=lambda x: int(x[0][1]))
    video_relation_list: List[VideoRelation] = []
    last_modify_rel_list: List[VideoRelation] = []
    for i, (index, prediction) in enumerate(short_term_relations):
        vid, fstart, fend = index
        pred_list, _iou, _trackid = prediction
        sorted_pred_list = sorted(pred_list, key=lambda x: x[0], reverse=True)[:max_traj_num_in_clip]
        trajs = load(dataset, vid, fstart, fend)
        for traj in trajs:
            traj.pstart, traj.pend = fstart, fend
            traj.vsig = get_segment_signature(vid, fstart, fend)
        cur_modify_rel_list: List[VideoRelation] = []
        # removal keeps a sorted list sorted and no remaining relation changes its confidence inside a
        # segment, so the reference's per-prediction stable sort (association.py:156) is done once
        last_modify_rel_list.sort(key=lambda r: r.mean_confs(), reverse=True)
        by_triplet: Dict[tuple, List[VideoRelation]] = {}
        for r in last_modify_rel_list:
            by_triplet.setdefault(tuple(int(x) for x in r.triplet()), []).append(r)
        # one launch for every IoU the greedy loop below can ask for
        wanted = []
        for pred in sorted_pred_list if i > 0 else []:
            straj, otraj = trajs[pred[2][0]], trajs[pred[2][1]]
            for r in by_triplet.get(tuple(int(x) for x in pred[1]), ()):
                if straj.pstart < r.fend and otraj.pstart < r.fend:
                    wanted += [(r.straj, straj), (r.otraj, otraj)]
        cache = batched_traj_iou(wanted)
        for pred in sorted_pred_list:
            conf_score = pred[0]
            s_cid, pid, o_cid = pred[1]
            straj, otraj = trajs[pred[2][0]], trajs[pred[2][1]]
            is_merged = False
            if i > 0:
                cands = by_triplet.get((int(s_cid), int(pid), int(o_cid)), [])
                for r in cands:
                    if (straj.pstart < r.fend and otraj.pstart < r.fend) and r.both_overlap(straj, otraj, cache=cache):
                        r.extend(straj, otraj, conf_score)
                        cands.remove(r)
                        last_modify_rel_list.remove(r)
                        cur_modify_rel_list.append(r)
                        is_merged = True
                        break
            if not is_merged:
                # association.py:139 passes the score in the first segment, :169 does not (confidence 1)
                r = VideoRelation(vid, s_cid, pid, o_cid, straj, otraj, confs=conf_score) if i == 0 \
                    else VideoRelation(vid, s_cid, pid, o_cid, straj, otraj)
                video_relation_list.append(r)
                cur_modify_rel_list.append(r)
        last_modify_rel_list = cur_modify_rel_list
    return [rel.serialize(dataset) for rel in video_relation_list]
