#!/usr/bin/env python
"""bench.py — pair-stage throughput (BASELINE.json metric: tracklet pairs scored / s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU reference arm (oracle port)

A *step* is one pass of the hot path (all-pairs geometry + vIoU || relationness + top-K -> relative
features, predicate and span heads of the K survivors -> triplet records) over one batch of synthetic
VidOR-shaped videos
(N=64 tracklets, T=2000 frames, 80 classes, 50 predicates: BASELINE.json configs[2]).
`value` counts P = N(N-1) ordered pairs per video with the inputs resident in HBM; `e2e` is the same
metric through the host-facing call with pinned host buffers, H2D and D2H inside the timed region.
Under torchrun each rank processes its own shard of videos (weak scaling, no data-path collective)
and the per-video top-K triplet records are all-gathered once at the end of the step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOAD = "vidor_single"          # N=64, T=2000, C=80, R=50, K=256
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--videos", type=int, default=16, help="videos per GPU per step")
    ap.add_argument("--precision", default="tensor", choices=["tensor", "fp32"])
    ap.add_argument("--no-sparsify", action="store_true", help="heads on all P pairs (reference quirk Q3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from the host instead of CUDA graphs")
    ap.add_argument("--depth", type=int, default=3, help="batches in flight in the end-to-end serving loop")
    ap.add_argument("--fp32-transport", action="store_true",
                    help="ship boxes / motion histograms as fp32 instead of the lossless compact u16 / u8 form")
    ap.add_argument("--three-graphs", action="store_true",
                    help="replay the step as three CUDA graphs joined on the host instead of one graph")
    ap.add_argument("--compute-streams", type=int, default=2, choices=[1, 2],
                    help="compute streams of the serving loop (2: the tail of step i overlaps the geometry of i+1)")
    return ap.parse_args()


def measured_traffic(videos):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this
    workload (profiles/r1_geo_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum); None when the
    capture does not describe the requested batch."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_geo_traffic.json")) as f:
            t = json.load(f)
        if int(t["videos_per_launch"]) != int(videos):
            return None
        return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:  # noqa: BLE001
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    NVML (pynvml) is polled every few milliseconds from a thread — the timed region of this bench is
    tens of milliseconds, too short for `nvidia-smi -lms`; nvidia-smi is the fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, cuda_index: int):
        self.cuda_index = cuda_index
        self.sm, self.mask, self.power = [], 0, []
        self.max_mhz = None
        self._stop = threading.Event()
        self.thread = None
        self.source = None
        self.period = float(os.environ.get("TSPN_BENCH_CLOCK_PERIOD", "0.005"))

    def _nvml_loop(self, nv, handle):
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                self.mask |= int(get_reasons(handle))
                self.power.append(nv.nvmlDeviceGetPowerUsage(handle) / 1000.0)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self._stop.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.cuda_index), "--query-gpu=" + q,
                                               "--format=csv,noheader,nounits"], text=True, timeout=5)
                r = [c.strip() for c in out.strip().split(",")]
                self.sm.append(float(r[0]))
                self.max_mhz = float(r[1])
                for b, val in zip(bits, r[2:6]):
                    if val.lower().startswith("active"):
                        self.mask |= b
            except Exception:  # noqa: BLE001
                break

    def start(self):
        if os.environ.get("TSPN_BENCH_SAMPLER", "") == "off":
            return
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.cuda_index).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            try:
                handle = nv.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:  # noqa: BLE001
                handle = nv.nvmlDeviceGetHandleByUUID(uuid)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
        except Exception:  # noqa: BLE001
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
        self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=6)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm), "source": self.source,
                "power_w_max": max(self.power) if self.power else None}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's CPU path) — the only place oracle/ is executed
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(video, sd, topk, sizes, stride, sample_pairs=256):
    """One video through the reference's CPU path; returns (seconds for the whole video, detail).

    Geometry / feature rows are computed for a bounded sample of pairs and scaled to P; everything
    else runs at full size.  The vIoU + per-frame geometry is the float64 oracle port ("oracle, not
    reference": the reference has no code for the per-frame channels); relationness, sort, classifier
    and span head are the reference's own torch CPU ops (oracle.heads.*_ref)."""
    from oracle import features as ofeat, geometry as ogeo, heads as oheads
    n, p = video.n_tracklets, video.n_pairs
    pr = ogeo.enumerate_pairs(n)
    rng = np.random.Generator(np.random.PCG64(0))
    sel = np.sort(rng.choice(p, size=min(sample_pairs, p), replace=False))
    det = {}
    # numpy releases the GIL: the sampled pairs are split over all host cores
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    chunks = [c for c in np.array_split(sel, cores) if len(c)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        parts = list(ex.map(lambda c: ogeo.pair_geometry(video.boxes, video.span, pr[c, 0], pr[c, 1]), chunks))
    det["geometry_viou"] = (time.perf_counter() - t0) * p / len(sel)
    geo, viou, tiou, ov = (np.concatenate([q[j] for q in parts], axis=0) for j in range(4))
    t0 = time.perf_counter()
    scores = oheads.ppn_head_ref(video.cls, video.cls, sd)
    order = torch.sort(scores.view(-1), descending=True)[1][:topk]
    det["relationness_topk"] = time.perf_counter() - t0
    k = int(order.shape[0])
    ksel = sel[:min(k, len(sel))]
    t0 = time.perf_counter()
    rel = ofeat.relative_block(geo[:len(ksel)], ov[:len(ksel)])
    feats = ofeat.assemble_features(video.cls, video.motion, rel, pr[ksel]).astype(np.float32)
    det["features_topk_rows"] = (time.perf_counter() - t0) * k / len(ksel)
    if len(ksel) < k:
        feats = np.concatenate([feats] * (k // len(ksel) + 1), axis=0)[:k]
    t0 = time.perf_counter()
    with torch.no_grad():
        oheads.relation_predictor_ref(feats, sd)
    det["predicate_head"] = time.perf_counter() - t0
    x = np.ascontiguousarray(np.concatenate([geo] * (k // geo.shape[0] + 1), axis=0)[:k], dtype=np.float32)
    t0 = time.perf_counter()
    with torch.no_grad():
        reg = oheads.dpn_head_ref(x, sd).numpy()
    oheads.decode_spans_f64(reg, sizes, stride)
    det["span_head_decode"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        logits = oheads.relation_predictor_ref(feats, sd).numpy()
    rows = np.resize(sel, k)
    oheads.postprocess_ref(logits, video.cls, pr[rows], 20, 200)
    det["postprocess"] = time.perf_counter() - t0
    return sum(det.values()), det


def run_reference(args, rank, world):
    """`--impl reference`: rank 0 alone times the CPU path; other ranks exit 0 without work."""
    if rank != 0:
        return
    from tspn_b200 import synth
    spec = synth.CONFIGS[WORKLOAD]
    c, r, k = spec["classes"], spec["predicates"], spec["topk"]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0)
    sizes, stride = (16.0, 64.0, 256.0, 1024.0), 16.0
    n, t = spec["n"][0], spec["t"][0]
    sample = 16 * cores
    times = []
    for i in range(args.warmup + args.steps):
        v = synth.make_video(n, t, c, seed=i)
        sec, det = cpu_reference_step(v, sd, k, sizes, stride, sample_pairs=sample)
        if i >= args.warmup:
            times.append(sec)
    pairs = n * (n - 1)
    ms = 1e3 * float(np.mean(times))
    value = pairs / (ms / 1e3)
    line = {
        "metric": "tracklet pairs scored/sec (N=64,T=2000)", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 (numpy + torch CPU)", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": "VidOR-shaped video N=64 T=2000 C=80 R=50 K=256, 1 video per step",
                   "sparsify": True, "host_cpu_count": cores, "torch_threads": torch.get_num_threads()},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": "1 video/step; geometry+feature rows on %d sampled pairs scaled to P=%d, "
                                   "relationness/top-K/classifier/span head at full size" % (sample, pairs),
                         "detail_s": det},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from tspn_b200 import _lib, ops, synth
    from tspn_b200.batch import HostBatch
    from tspn_b200.pipeline import PairStage, StageConfig
    from tspn_b200.serving import PipelinedStage

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ops.require_device()
    spec = synth.CONFIGS[WORKLOAD]
    c, r, k = spec["classes"], spec["predicates"], spec["topk"]
    n, t = spec["n"][0], spec["t"][0]
    sparsify = not args.no_sparsify
    sizes, stride = (16.0, 64.0, 256.0, 1024.0), 16.0
    cfg = StageConfig(n_classes=c, n_predicates=r, topk=k, use_ppn=True, use_dpn=True, sparsify=sparsify,
                      precision=args.precision, anchor_sizes=sizes, anchor_stride=stride)
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0)
    stage = PairStage(cfg)
    stage.load_weights(sd, dev)
    videos = [synth.make_video(n, t, c, seed=100000 * rank + i) for i in range(args.videos)]
    host = HostBatch.from_videos(videos, compact=not args.fp32_transport)
    pairs_per_step = sum(v.n_pairs for v in videos)

    # ---- resident-input steps ------------------------------------------------------------------
    # The serving loop below owns `depth` slots (device inputs + captured CUDA graphs + pinned result
    # buffers); the resident-input measurement replays slot 0's graphs on inputs already in HBM.
    group = dist.group.WORLD if world > 1 else None
    pipe = PipelinedStage(stage, host, device=dev, depth=args.depth, graphs=not args.eager, group=group,
                          compute_streams=args.compute_streams, single_graph=not args.three_graphs)
    slot0 = pipe.slots[0]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)    # > 126 MB L2
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(timers=None):
        if slot0.graphed is not None:
            return slot0.graphed.replay(timers=timers)
        return stage.forward(slot0.batch, timers=timers)

    geo_ms, step_ms = [], []
    sampler = ClockSampler(local_rank)
    sampler.start()                                     # samples through warm-up, timed steps and e2e
    launches0 = ops.launch_count()
    for i in range(args.warmup):
        one_step()
    barrier()
    launches_per_step = (ops.launch_count() - launches0) // max(args.warmup, 1)
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                                   # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        timers = {}
        e0.record()
        one_step(timers=timers)
        e1.record()
        # the single-graph replay times the geometry kernel with external event nodes that the next
        # replay re-records: read them now (the steps are separated by the untimed L2 flush anyway)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        geo_ms.append(timers["geo"][0].elapsed_time(timers["geo"][1]))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    # the pair kernel on its own (same inputs, nothing else on the device): what the co-resident side branch costs it
    alone_ms = []
    geom_alone = slot0.graphed.result.geom if slot0.graphed is not None else ops.pair_geometry_outputs(slot0.batch)
    for i in range(3 + min(args.steps, 10)):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.pair_geometry_phase(slot0.batch, geom_alone, _lib.GEO_PHASE_MAIN)
        e1.record()
        e1.synchronize()
        if i >= 3:
            alone_ms.append(e0.elapsed_time(e1))
    total_ms = float(sum(step_ms))
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = world * pairs_per_step / (ms_per_step / 1e3)

    # ---- end-to-end steps: pinned host in, pinned host out ------------------------------------
    # tspn_b200.serving.PipelinedStage (the host-facing call): every step copies its own inputs from
    # pinned host memory and its own results back to pinned host memory inside the timed region; H2D of
    # step i+1 and D2H of step i-1 overlap the kernels of step i (three streams, --depth batches in flight).
    def e2e_loop(steps):
        n_out = 0
        for out in pipe.run(host for _ in range(steps)):
            n_out += int(out["record_counts"][0] >= 0)      # touch the host result
        assert n_out == steps

    e2e_loop(max(args.warmup, 2))
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * pairs_per_step * args.steps / e2e_s
    d2h = pipe.d2h_bytes()
    clocks = sampler.stop()

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (all-pairs geometry + vIoU) ------------------------------
    tp, tb = (t + 3) // 4 * 4, (t + 7) // 8 * 8
    p1 = n * (n - 1)
    alg_bytes = args.videos * (32 * tp * p1 + 16 * p1 + 16 * n * tb + 8 * n)      # DESIGN.md, SURVEY 8d
    geo_avg_ms = float(np.mean(geo_ms))
    achieved = alg_bytes / (geo_avg_ms / 1e3) / 1e9
    peak, peak_src = measured_peak()
    line = {
        "metric": "tracklet pairs scored/sec (N=64,T=2000)", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 geometry/features, f64 volume sums, %s heads" % ("bf16 tcgen05" if args.precision == "tensor"
                                                                         else "f32 exact-order"),
        "data": "synthetic",
        "config": {"workload": "VidOR-shaped videos N=64 T=2000 C=80 R=50 K=256 (BASELINE.json configs[2]), "
                               "%d videos per GPU per step" % args.videos,
                   "videos_per_gpu": args.videos, "pairs_per_step_per_gpu": pairs_per_step, "sparsify": sparsify,
                   "precision": args.precision, "sharding": "per video, no data-path collective",
                   "l2": "256 MiB buffer zeroed between timed iterations (untimed); each step also writes "
                         "%.1f GB of outputs" % (alg_bytes / 1e9),
                   "wall_s_timed_region": t_wall,
                   "launch": "eager C-ABI calls" if args.eager else
                             ("3 CUDA-graph launches per step (side: relationness+top-K+motion norm || geo; tail)"
                              if args.three_graphs else
                              "1 CUDA-graph launch per step; three branches inside the graph: the persistent all-pairs "
                              "kernel | relationness -> top-K -> surviving-pair rows (relative block + span proposals "
                              "recomputed from the boxes) -> predicate head -> records | per-tracklet predicate terms, "
                              "volumes, vIoU finalize - the side branches co-reside with the all-pairs kernel"),
                   "e2e_pipeline": "tspn_b200.serving.PipelinedStage, depth %d: one H2D copy of the pinned input "
                                   "arena per step; H2D(i+1..) and D2H(i-1) overlap the kernels of step i; %d compute "
                                   "stream(s)" % (args.depth, args.compute_streams),
                   "h2d_transport": "u16 boxes + u8 motion counts (lossless for these inputs), expanded on the device"
                                    if host.boxes_compact and host.motion_compact else "fp32",
                   "multi_gpu_collective": "all_gather of [V,200,8] int32 triplet records per step (e2e loop)"},
        "roofline": {"bound": "hbm", "kernel": "pair_geo_kernel (CUDA events immediately around this launch)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(args.videos), "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                     "avg_launch_ms": geo_avg_ms, "share_of_step": geo_avg_ms / (float(np.mean(step_ms))),
                     "note": "achieved / frac are measured inside the timed steps, where the side branches' kernels "
                             "co-reside with this kernel on every SM; `alone` is the same launch with an idle device",
                     "alone": {"avg_launch_ms": float(np.mean(alone_ms)),
                               "achieved": alg_bytes / (float(np.mean(alone_ms)) / 1e3) / 1e9,
                               "frac": alg_bytes / (float(np.mean(alone_ms)) / 1e3) / 1e9 / peak}},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": host.h2d_bytes(),
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches_per_step * args.steps),
        "launches_per_step": int(launches_per_step),
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sec, det = cpu_reference_step(videos[0], sd, k, sizes, stride, sample_pairs=16 * cores)
        line["cpu_baseline"] = {"value": p1 / sec, "unit": "pairs/s", "cores": cores, "kind": "port",
                                "sample": "1 video (P=%d); geometry+feature rows on %d sampled pairs (split over "
                                          "%d threads) scaled to P, heads at full size; torch threads=%d"
                                          % (p1, 16 * cores, cores, torch.get_num_threads()),
                                "detail_s": det}
    emit(line)


_JSON_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the run, on the process's original stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Libraries write to fd 1 behind python's back (NCCL prints its version banner there under torchrun):
    # keep the real stdout for the JSON line only and send everything else to stderr.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
