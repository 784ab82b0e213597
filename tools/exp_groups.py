#!/usr/bin/env python
"""Experiment: software-pipeline one step across G groups of videos inside one CUDA graph, so that the
latency-bound tail of group g runs underneath the HBM-bound geometry kernel of group g+1.
Usage (GPU box): python tools/exp_groups.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tspn_b200 import synth  # noqa: E402
from tspn_b200.batch import HostBatch  # noqa: E402
from tspn_b200.pipeline import PairStage, StageConfig  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
spec = synth.CONFIGS["vidor_single"]
c, r, k = spec["classes"], spec["predicates"], spec["topk"]
n, t = spec["n"][0], spec["t"][0]
cfg = StageConfig(n_classes=c, n_predicates=r, topk=k, use_ppn=True, use_dpn=True, sparsify=True, precision="tensor",
                  anchor_sizes=(16.0, 64.0, 256.0, 1024.0), anchor_stride=16.0)
sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0)
stage = PairStage(cfg)
stage.load_weights(sd, dev)
videos = [synth.make_video(n, t, c, seed=i) for i in range(16)]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def capture(groups, geo_streams, tail_priority):
    batches = [HostBatch.from_videos(videos[g::groups]).to_device(dev) for g in range(groups)]
    for b in batches:
        stage.forward(b)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream(dev)
    s_geo = [torch.cuda.Stream(dev) for _ in range(geo_streams)]
    s_side = torch.cuda.Stream(dev)
    s_tail = torch.cuda.Stream(dev, priority=tail_priority)
    results = []
    with torch.cuda.graph(graph, stream=cap):
        cur = torch.cuda.current_stream(dev)
        for s in s_geo + [s_side, s_tail]:
            s.wait_stream(cur)
        sides, ev_side = [], []
        with torch.cuda.stream(s_side):
            for b in batches:
                sides.append(stage._seg_side(b, None))
                e = torch.cuda.Event()
                e.record(s_side)
                ev_side.append(e)
        geoms, ev_geo = [], []
        for g, b in enumerate(batches):
            s = s_geo[g % geo_streams]
            with torch.cuda.stream(s):
                geoms.append(stage._seg_geo(b, stage._geo_alloc(b, None), with_pre=True))
                e = torch.cuda.Event()
                e.record(s)
                ev_geo.append(e)
        with torch.cuda.stream(s_tail):
            for g, b in enumerate(batches):
                s_tail.wait_event(ev_geo[g])
                s_tail.wait_event(ev_side[g])
                results.append(stage._seg_tail(b, None, True, sides[g], geoms[g]))
        for s in s_geo + [s_side, s_tail]:
            cur.wait_stream(s)
    torch.cuda.synchronize()
    return graph, results, batches


def time_graph(graph):
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms)), float(np.min(ms))


pairs = 16 * n * (n - 1)
ref = None
for groups, geo_streams, prio in [(1, 1, 0), (2, 1, -1), (2, 2, -1), (4, 1, -1), (4, 2, -1), (4, 2, 0), (8, 2, -1), (4, 4, -1)]:
    graph, results, batches = capture(groups, geo_streams, prio)
    mean, best = time_graph(graph)
    # the records of every video must not depend on the grouping
    rec = torch.cat([res.records for res in results]).cpu()
    order = np.concatenate([np.arange(16)[g::groups] for g in range(groups)])
    rec = rec[np.argsort(order)]
    if ref is None:
        ref = rec
    same = bool(torch.equal(rec, ref))
    print("groups %d geo_streams %d tail_prio %d: step %.4f ms (best %.4f)  %.2f M pairs/s  records identical: %s"
          % (groups, geo_streams, prio, mean, best, pairs / mean / 1e3, same), flush=True)
    del graph, results, batches
    torch.cuda.empty_cache()
