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


PEER_WORKER = r'''
import json, os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from tspn_b200 import sharding, synth
from tspn_b200.batch import HostBatch
from tspn_b200.pipeline import PairStage, StageConfig
from tspn_b200.serving import PipelinedStage

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
ok, info = True, {}

# 1. the exchange alone: 40 steps on a ring of 4, random records, two producer streams used alternately (so steps may
#    execute out of order), every gathered step compared with NCCL's all-gather of the same records
peer = sharding.PeerRecords((6, 200, 8), dist.group.WORLD, dev, ring=4)
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
consumer = torch.cuda.Stream(dev)
g = torch.Generator(device="cuda").manual_seed(1000 + rank)
for s in range(40):
    rec = torch.randint(-2**31, 2**31 - 1, (6, 200, 8), dtype=torch.int32, device=dev, generator=g)
    want = torch.empty((world, 6, 200, 8), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(want, rec)
    torch.cuda.synchronize()
    st = streams[s %% 2]
    with torch.cuda.stream(st):
        peer.scatter(rec)
        done = torch.cuda.Event(); done.record(st)
    with torch.cuda.stream(consumer):
        consumer.wait_event(done)
        got = peer.gather().clone()
        peer.release()
    consumer.synchronize()
    ok = ok and bool(torch.equal(got, want))
peer.check()
info["exchange_steps"] = 40

# 2. the serving loop: the same batches through collective="nccl" and collective="peer"
C, R = 35, 132
sd = synth.make_weights(C, R, synth.feature_dim(C), dpn_in=8, seed=0)
stage = PairStage(StageConfig(n_classes=C, n_predicates=R, topk=64, sparsify=True, precision="tensor",
                              num_span_proposals=16))
stage.load_weights(sd, dev)
hosts = [HostBatch.from_videos([synth.make_video(10, 300, C, seed=1000 * rank + 10 * q + j) for j in range(3)])
         for q in range(7)]
outs = {}
for coll in ("nccl", "peer"):
    pipe = PipelinedStage(stage, hosts[0], device=dev, depth=3, group=dist.group.WORLD, collective=coll)
    outs[coll] = [{k: v.clone() for k, v in o.items()} for o in pipe.run(iter(hosts))]
    if coll == "peer":
        pipe.peer.check()
    torch.cuda.synchronize()
    dist.barrier()
for a, b in zip(outs["nccl"], outs["peer"]):
    ok = ok and a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    ok = ok and bool(torch.equal(b["records_all_ranks"][rank], b["records"]))
    ok = ok and int(b["record_counts"].sum()) > 0
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"ok": bool(flag.item()), **info}))
dist.destroy_process_group()
sys.exit(0 if bool(flag.item()) else 1)
'''


def test_peer_record_exchange_equals_nccl_all_gather(tmp_path):
    """csrc/peer_records.cu on two GPUs: stores into the peers' gather buffers + flags give, step for step, what
    all_gather_into_tensor gives - alone (ring reuse, out-of-order producer streams) and inside the serving loop."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "peer_worker.py"
    script.write_text(PEER_WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29733", str(script)],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok"]
