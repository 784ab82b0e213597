"""The glue after the pair stage — mirror of the tail of lib/modeling/predict.py and of the JSON writer
in base.py, fed by the on-device triplet records.

``predict()`` in the reference (lib/modeling/predict.py:14-120) turns each segment's ``rel_logits`` into
``short_term_relations[(vid, fstart, fend)] = (predictions, iou, trackid)`` with python loops over two
full sorts; the same selection is already done on the GPU by ``tspn_postprocess`` (``StageResult.records``:
``[V, 200, 8]`` int32, fields ``ops.RECORD_FIELDS``).  This module only reshapes those records into the
reference's containers, so that ``association.greedy_relational_association`` (base.py:98-105) and the
prediction JSON (base.py:107-113) consume them unchanged.
"""
from __future__ import annotations

import json
from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .ops import RECORD_FIELDS

F_SCORE, F_SCLS, F_PRED, F_OCLS, F_STID, F_OTID, F_START, F_END = range(len(RECORD_FIELDS))


def records_to_predictions(records, count: int) -> List[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """One video's ``[200, 8]`` int32 records -> ``[(score, triplet[3], pair_tid[2])]`` exactly as
    predict.py:110-113 builds them (numpy scalars/arrays; descending score, ties to the lower index)."""
    rec = records.cpu().numpy() if isinstance(records, torch.Tensor) else np.asarray(records)
    rec = np.ascontiguousarray(rec[:int(count)], dtype=np.int32)
    scores = rec[:, F_SCORE].copy().view(np.float32)
    return [(np.array(scores[i]), np.array([rec[i, F_SCLS], rec[i, F_PRED], rec[i, F_OCLS]], dtype=np.int64),
             np.array([rec[i, F_STID], rec[i, F_OTID]], dtype=np.int64)) for i in range(rec.shape[0])]


def short_term_relations(records, counts, indexes: Sequence[tuple], ious: Optional[Sequence] = None,
                         trackids: Optional[Sequence] = None) -> Dict[tuple, tuple]:
    """``{(vid, fstart, fend): (predictions, iou, trackid)}`` (predict.py:115-119) for a batch of segments:
    ``records [V, 200, 8]`` / ``counts [V]`` from ``StageResult`` (or from ``sharding.gather_records``),
    ``indexes[v]`` the segment key of batch entry v.  Segments without a relation are skipped, as
    predict.py:60-64 does."""
    rec = records.cpu().numpy() if isinstance(records, torch.Tensor) else np.asarray(records)
    cnt = counts.cpu().numpy() if isinstance(counts, torch.Tensor) else np.asarray(counts)
    out = {}
    for v, index in enumerate(indexes):
        if int(cnt[v]) <= 0:
            continue
        out[tuple(index)] = (records_to_predictions(rec[v], int(cnt[v])),
                             np.array(ious[v]) if ious is not None else None,
                             np.array(trackids[v]) if trackids is not None else None)
    return out


def group_by_video(short_term: Dict[tuple, tuple]) -> Dict[str, list]:
    """base.py:92-96: ``{vid: [(index, short_term_relation)]}``."""
    by_video = defaultdict(list)
    for index, st_rel in short_term.items():
        by_video[index[0]].append((index, st_rel))
    return by_video


def detect_video_relations(dataset, short_term: Dict[tuple, tuple], max_traj_num_in_clip: int = 100,
                           trajectory_proposal=None) -> Dict[str, list]:
    """base.py:98-105: greedy relational association of every video's short-term relations."""
    from . import association
    return {vid: association.greedy_relational_association(dataset, rels, max_traj_num_in_clip=max_traj_num_in_clip,
                                                           trajectory_proposal=trajectory_proposal)
            for vid, rels in group_by_video(short_term).items()}


def save_video_relations(video_relations: Dict[str, list], path: str) -> None:
    """The prediction file ``evaluate.py`` reads (base.py:107-113): ``{'version', 'results'}``."""
    with open(path, 'w') as fout:
        json.dump({'version': 'VERSION 1.0', 'results': video_relations}, fout)


def load_video_relations(path: str) -> Dict[str, list]:
    """``pred['results']`` of a prediction file (lib/evaluation/visual_relation_detection.py:141-145)."""
    with open(path, 'r') as fin:
        return json.load(fin)['results']
