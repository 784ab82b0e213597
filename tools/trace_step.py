"""Timeline of one step of the bench workload: kernel start / duration / stream from CUPTI (torch.profiler),
relative to the first kernel of the step.  Unlike the ncu launch list (serialised, cold cache) this shows the
step as it really runs: which kernels overlap the pair kernel and what is left on the critical path.

    python tools/trace_step.py [--videos 16] [--steps 3] > gpurun_out/trace.txt
"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from tspn_b200 import synth  # noqa: E402
from tspn_b200.batch import HostBatch  # noqa: E402
from tspn_b200.pipeline import PairStage, StageConfig  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--no-dpn", action="store_true", help="without the span head")
    ap.add_argument("--span-proposals", type=int, default=64, help="spans kept per pair by the NMS (0 = none)")
    ap.add_argument("--workload", default="vidor_single", help="synth.CONFIGS key; ragged workloads trace their first "
                                                               "batches (capacity graphs, as the bench replays them)")
    ap.add_argument("--batches", type=int, default=3, help="ragged workloads: batches to trace")
    ap.add_argument("--relationness", default="fp32", choices=["fp32", "tensor"])
    ap.add_argument("--geo-layout", default="dense", choices=["dense", "windowed"])
    args = ap.parse_args()
    spec = synth.CONFIGS[args.workload]
    c, r, k = spec["classes"], spec["predicates"], spec["topk"]
    n, t = spec["n"][1], spec["t"][1]
    vidvrd = c == 35
    cfg = StageConfig(n_classes=c, n_predicates=r, topk=k, use_ppn=True, use_dpn=not args.no_dpn, sparsify=True,
                      precision="tensor", relationness_precision=args.relationness,
                      anchor_sizes=(15.0, 30.0, 45.0, 60.0) if vidvrd else (16.0, 64.0, 256.0, 1024.0),
                      anchor_stride=7.5 if vidvrd else 16.0, num_span_proposals=args.span_proposals,
                      geo_layout=args.geo_layout)
    stage = PairStage(cfg)
    stage.load_weights(synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0), "cuda")
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    if spec["videos"] == 1:
        count = args.videos if args.workload != "stress" else 1
        host = HostBatch.from_videos([synth.make_video(n, t, c, seed=i) for i in range(count)], compact=True)
        batch = host.to_device("cuda")
        graphed = None if args.eager else stage.capture(batch)

        def step():
            flush.zero_()
            if graphed is not None:
                graphed.replay()
            else:
                stage.forward(batch)
            torch.cuda.synchronize()
    else:
        from tspn_b200.serving import host_batches_for
        shapes = synth.config_shapes(args.workload, 0, 160)
        vids = [synth.make_video(a, b, c, seed=i) for i, (a, b) in enumerate(shapes)]
        hosts, _, caps = host_batches_for(vids, c)
        hosts = hosts[:args.batches]
        graphs = {}
        for h in hosts:
            if h.capacity not in graphs:
                b = h.to_device("cuda")
                graphs[h.capacity] = (b, None if args.eager else stage.capture(b))
        print("# %d batches: %s" % (len(hosts), [(h.num_real, int(h.actual[1]), h.capacity.geo_chunk) for h in hosts]))

        def step():
            for h in hosts:                       # one flush per batch: every batch prints as its own "step"
                flush.zero_()
                b, g = graphs[h.capacity]
                b.copy_from(h)
                if g is not None:
                    g.replay()
                else:
                    stage.forward(b)
                torch.cuda.synchronize()

    for _ in range(3):
        step()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            step()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    with open(path) as f:
        ev = [e for e in json.load(f)["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
    ev.sort(key=lambda e: e["ts"])
    # split into steps at the flush kernel
    steps, cur = [], []
    for e in ev:
        name = e["name"]
        if "FillFunctor" in name or "fill_kernel" in name.lower() or e.get("cat") == "gpu_memset" and e["dur"] > 20:
            if cur:
                steps.append(cur)
            cur = []
            continue
        cur.append(e)
    if cur:
        steps.append(cur)
    n_show = args.steps if spec["videos"] == 1 else args.steps * args.batches
    for si, evs in enumerate(steps[-n_show:]):
        if not evs:
            continue
        t0 = evs[0]["ts"]
        end = max(e["ts"] + e["dur"] for e in evs)
        print("== step %d: %.1f us from first kernel start to last kernel end, %d kernels" % (si, end - t0, len(evs)))
        for e in evs:
            print("  %8.1f +%7.1f  s%-3s %s" % (e["ts"] - t0, e["dur"], e.get("args", {}).get("stream", "?"),
                                              e["name"][:90]))


if __name__ == "__main__":
    main()
