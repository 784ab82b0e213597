"""Multi-GPU host logic on CPU: LPT sharding and the record all-gather over gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tspn_b200 import sharding, synth


def test_lpt_sharding_is_balanced_and_deterministic():
    shapes = synth.config_shapes("vidor_val", seed=0)            # 835 VidOR-val-shaped videos
    assert len(shapes) == 835
    for world in (1, 2, 4, 8):
        sh = sharding.shard_videos(shapes, world)
        assert sorted(i for s in sh for i in s) == list(range(835))
        assert sharding.imbalance(shapes, sh) < 1.02               # < 2 % at 835 videos
        assert sh == sharding.shard_videos(shapes, world)
    # naive contiguous split (the reference's DistributedSampler) is much worse
    naive = [list(range(r, 835, 1))[r * 105:(r + 1) * 105] for r in range(8)]
    assert sharding.imbalance(shapes, sharding.shard_videos(shapes, 8)) <= sharding.imbalance(shapes, naive) + 1e-9
    assert sharding.shard_videos([(3, 10)], 4) == [[0], [], [], []]
    assert sharding.shard_videos([], 2) == [[], []]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shapes, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shards = sharding.shard_videos(shapes, world)
        mine = shards[rank]
        m = 5
        rec = torch.zeros((len(mine), m, 8), dtype=torch.int32)
        cnt = torch.zeros(len(mine), dtype=torch.int32)
        for j, vid in enumerate(mine):                          # content encodes the global video id
            rec[j] = vid * 1000 + torch.arange(m * 8, dtype=torch.int32).view(m, 8)
            cnt[j] = vid % (m + 1)
        all_r, all_c = sharding.gather_records(rec, cnt, shards)
        torch.save((all_r, all_c), os.path.join(result_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_videos", [7, 2, 1])
def test_gather_records_gloo_world2(tmp_path, n_videos):
    shapes = [(2 + (i * 5) % 9, 10 + 7 * i) for i in range(n_videos)]
    port = _free_port()
    mp.spawn(_worker, args=(2, port, shapes, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert torch.equal(r0[0], r1[0]) and torch.equal(r0[1], r1[1])      # identical on every rank
    m = 5
    for vid in range(n_videos):
        want = vid * 1000 + torch.arange(m * 8, dtype=torch.int32).view(m, 8)
        assert torch.equal(r0[0][vid], want)
        assert int(r0[1][vid]) == vid % (m + 1)
