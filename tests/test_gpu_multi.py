"""Two ranks over NCCL on real GPUs: the sharded pair stage's gathered records equal the single-GPU records.

Skipped on a box with fewer than two GPUs (the CPU/gloo counterpart is tests/test_sharding_gloo.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from tspn_b200 import sharding, synth
from tspn_b200.pipeline import PairStage, StageConfig
from tspn_b200.serving import host_batches_for

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
C, R = 80, 50
shapes = synth.config_shapes("vidor_val", 11, 24)
sd = synth.make_weights(C, R, synth.feature_dim(C), dpn_in=8, seed=0)
stage = PairStage(StageConfig(n_classes=C, n_predicates=R, topk=64, sparsify=True, precision="tensor",
                              num_span_proposals=16, anchor_sizes=(16.0, 64.0, 256.0, 1024.0), anchor_stride=16.0))
stage.load_weights(sd, "cuda")

def records_of(ids):
    vids = [synth.make_video(shapes[i][0], shapes[i][1], C, seed=i) for i in ids]
    hosts, batch_vids, _ = host_batches_for(vids, C, geo_budget_bytes=256 << 20, max_videos=8)
    rec = torch.zeros((len(ids), 200, 8), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(len(ids), dtype=torch.int32, device="cuda")
    for h, bv in zip(hosts, batch_vids):
        res = stage.forward(h.to_device("cuda"))
        idx = torch.as_tensor(bv, dtype=torch.int64, device="cuda")
        rec.index_copy_(0, idx, res.records[:len(bv)])
        cnt.index_copy_(0, idx, res.record_counts[:len(bv)])
    return rec, cnt

shards = sharding.shard_videos(shapes, world)
rec, cnt = records_of(shards[rank])
all_rec, all_cnt = sharding.gather_records(rec, cnt, shards)
torch.cuda.synchronize()
ok = True
if rank == 0:
    one_rec, one_cnt = records_of(list(range(len(shapes))))        # the same videos on one GPU, other batches
    ok = bool(torch.equal(one_rec, all_rec) and torch.equal(one_cnt, all_cnt) and int(all_cnt.sum()) > 0)
    print(json.dumps({"ok": ok, "records": int(all_cnt.sum()), "imbalance": sharding.imbalance(shapes, shards)}))
# every rank holds the same gathered tensor
chk = all_rec.to(torch.int64).sum().reshape(1)
both = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(both, chk)
assert all(int(b) == int(both[0]) for b in both)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_sharded_records_over_nccl_equal_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok"]
