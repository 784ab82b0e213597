"""What limits the host-facing loop when several ranks share one box?  Run under torchrun with N ranks:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_limiter.py

Every rank runs, on the bench workload's batch (16 VidOR-shaped videos), the pieces of the serving loop in isolation
and together, all ranks at once (barrier before each timed region), and rank 0 prints one JSON line per piece with
the per-rank mean / min rate:

    h2d      only the step's host->device copy (pinned arena -> HBM), `depth` copies in flight
    d2h      only the step's device->host copy (results -> pinned buffers)
    both     the two copies on their own streams, concurrently
    kernels  only the step's CUDA graph on resident inputs
    loop     the full serving loop (tspn_b200.serving.PipelinedStage), with and without the per-step NCCL all-gather
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tspn_b200 import affinity, synth  # noqa: E402
from tspn_b200.pipeline import PairStage, StageConfig  # noqa: E402
from tspn_b200.serving import PipelinedStage, host_batches_for  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--videos", type=int, default=16)
    ap.add_argument("--span-proposals", type=int, default=64)
    ap.add_argument("--no-affinity", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bound = None if args.no_affinity else affinity.bind_to_gpu(local, world)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    spec = synth.CONFIGS["vidor_single"]
    c, r, k = spec["classes"], spec["predicates"], spec["topk"]
    stage = PairStage(StageConfig(n_classes=c, n_predicates=r, topk=k, sparsify=True, precision="tensor",
                                  anchor_sizes=(16.0, 64.0, 256.0, 1024.0), anchor_stride=16.0,
                                  num_span_proposals=args.span_proposals))
    stage.load_weights(synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0), dev)
    vids = [synth.make_video(64, 2000, c, seed=1000 * rank + i) for i in range(args.videos)]
    hosts, _, _ = host_batches_for(vids, c)
    host = hosts[0]
    pairs = sum(v.n_pairs for v in vids)
    depth = 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def report(name, seconds, steps, nbytes=None, extra=None):
        t = torch.tensor([seconds], dtype=torch.float64, device=dev)
        if world > 1:
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            secs = [float(x.item()) for x in allt]
        else:
            secs = [seconds]
        if rank == 0:
            line = {"piece": name, "ranks": world, "steps": steps,
                    "ms_per_step_mean": 1e3 * sum(secs) / len(secs) / steps, "ms_per_step_max": 1e3 * max(secs) / steps,
                    "Mpairs_per_s_aggregate": world * pairs * steps / max(secs) / 1e6}
            if nbytes:
                line["GBps_per_rank_mean"] = nbytes * steps / (sum(secs) / len(secs)) / 1e9
                line["GBps_per_rank_min"] = nbytes * steps / max(secs) / 1e9
                line["bytes_per_step"] = nbytes
            line.update(extra or {})
            print(json.dumps(line), flush=True)

    pipe = PipelinedStage(stage, host, device=dev, depth=depth, group=None)
    slots = pipe.slots
    # one full step so that the slots' pinned result buffers exist
    for out in pipe.run(iter([host] * depth)):
        pass
    outs = [s.keep[1] for s in slots]
    d2h_bytes = sum(v.numel() * v.element_size() for v in outs[0].values())
    h2d_bytes = host.h2d_bytes()
    s_a, s_b = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def h2d_loop(n):
        with torch.cuda.stream(s_a):
            for i in range(n):
                slots[i % depth].batch.copy_from(host)

    def d2h_loop(n):
        with torch.cuda.stream(s_b):
            for i in range(n):
                s = slots[i % depth]
                for key, src in outs[i % depth].items():
                    s.pinned[key][:src.shape[0]].copy_(src, non_blocking=True)

    for name, fn, nbytes in (("h2d", lambda n: h2d_loop(n), h2d_bytes), ("d2h", lambda n: d2h_loop(n), d2h_bytes),
                             ("both", lambda n: (h2d_loop(n), d2h_loop(n)), h2d_bytes + d2h_bytes)):
        fn(5)
        barrier()
        t0 = time.perf_counter()
        fn(args.steps)
        torch.cuda.synchronize()
        report(name, time.perf_counter() - t0, args.steps, nbytes, {"cpu_affinity": bound} if name == "h2d" else None)

    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        slots[i % depth].graphed.replay()
    torch.cuda.synchronize()
    report("kernels", time.perf_counter() - t0, args.steps)

    def loop(p, n):
        cnt = 0
        for out in p.run(host for _ in range(n)):
            cnt += int(out["record_counts"].shape[0])
        return cnt

    loop(pipe, 5)
    barrier()
    t0 = time.perf_counter()
    loop(pipe, args.steps)
    report("loop (no collective)", time.perf_counter() - t0, args.steps, None,
           {"h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes})
    if world > 1:
        del pipe
        pipe2 = PipelinedStage(stage, host, device=dev, depth=depth, group=dist.group.WORLD)
        loop(pipe2, 5)
        barrier()
        t0 = time.perf_counter()
        loop(pipe2, args.steps)
        report("loop (all-gather of the records per step)", time.perf_counter() - t0, args.steps)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
