"""Per-video sharding across GPUs and the one collective of the pair stage.

Videos are independent units (the reference treats every ``(vid, fstart, fend)`` as its own dataset
item, lib/dataset/vrdataset.py:56-83, and regroups per video at base.py:92-96), so the path shards
with no data-path collective: each rank runs the whole pair stage on its own videos.  The only
exchange is the all-gather of the fixed-size top-K triplet records at the end (32 bytes x
TOPK_PER_SEG per video), which replaces the pickle-based ``comm.all_gather`` of
lib/utils/comm.py:48-88 (two all-gathers plus padding) by a single ``all_gather_into_tensor``.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def video_cost(n_tracklets: int, n_frames: int) -> int:
    """Work of one video: ordered pairs x frames (the geometry kernel's bytes are 32*T per pair)."""
    return int(n_tracklets) * max(int(n_tracklets) - 1, 0) * int(n_frames)


def shard_videos(shapes: Sequence[Tuple[int, int]], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of videos to ranks.

    ``shapes[i] = (N_i, T_i)``.  Returns ``world_size`` lists of video indices (each sorted
    ascending); deterministic: ties go to the lower video index, then to the lower rank.  The
    reference's ``DistributedSampler`` (lib/dataset/samplers/distributed.py:49-58) is the naive
    equivalent: a contiguous split with wrap-around padding and no cost model.
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(shapes)), key=lambda i: (-video_cost(*shapes[i]), i))
    loads = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda q: (loads[q], q))
        out[r].append(i)
        loads[r] += video_cost(*shapes[i])
    return [sorted(v) for v in out]


def imbalance(shapes: Sequence[Tuple[int, int]], shards: List[List[int]]) -> float:
    """max rank load / mean rank load (1.0 = perfect)."""
    loads = [sum(video_cost(*shapes[i]) for i in s) for s in shards]
    mean = sum(loads) / max(len(loads), 1)
    return max(loads) / mean if mean > 0 else 1.0


def gather_records(records: torch.Tensor, counts: torch.Tensor, shards: List[List[int]], group=None):
    """All-gather the per-video triplet records of every rank and put them back in video order.

    ``records [V_local, M, 8]`` int32, ``counts [V_local]`` int32 for this rank's videos (in the order
    of ``shards[rank]``).  Ranks hold different numbers of videos, so each pads to ``max_v`` videos;
    one ``all_gather_into_tensor`` moves ``world * max_v * (M*8 + 1)`` int32.  Returns
    ``(records_all [V_total, M, 8], counts_all [V_total])`` indexed by global video id, identical on
    every rank.  Works on NCCL (CUDA tensors) and gloo (CPU tensors).
    """
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    v_local, m = int(records.shape[0]), int(records.shape[1])
    if v_local != len(shards[rank]):
        raise ValueError("rank %d holds %d videos but its shard lists %d" % (rank, v_local, len(shards[rank])))
    max_v = max(len(s) for s in shards)
    width = m * 8 + 1
    send = torch.zeros((max_v, width), dtype=torch.int32, device=records.device)
    if v_local:
        send[:v_local, :m * 8] = records.reshape(v_local, m * 8)
        send[:v_local, m * 8] = counts
    recv = torch.empty((world, max_v, width), dtype=torch.int32, device=records.device)
    dist.all_gather_into_tensor(recv.view(world * max_v, width), send, group=group)
    total = sum(len(s) for s in shards)
    out_r = torch.zeros((total, m, 8), dtype=torch.int32, device=records.device)
    out_c = torch.zeros(total, dtype=torch.int32, device=records.device)
    for r, vids in enumerate(shards):
        if vids:
            idx = torch.as_tensor(vids, dtype=torch.int64, device=records.device)
            out_r[idx] = recv[r, :len(vids), :m * 8].reshape(len(vids), m, 8)
            out_c[idx] = recv[r, :len(vids), m * 8]
    return out_r, out_c
