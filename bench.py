#!/usr/bin/env python
"""bench.py — pair-stage throughput (BASELINE.json metric: tracklet pairs scored / s).

    python bench.py --gpus N --steps K --warmup W [--workload NAME]      # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W [--workload NAME]   # CPU reference arm

A *step* is one pass of the hot path (all-pairs geometry + vIoU || relationness + top-K -> relative
features, predicate and span heads of the K survivors, span NMS -> triplet records) over the workload's
synthetic videos.  Workloads = BASELINE.json configs:

    vidvrd_single  configs[0]  N=20,  T=300,   35 classes, 132 predicates (64 such videos per GPU per step)
    vidvrd_test    configs[1]  200 ragged videos N<=40, T<=1200                       (sharded per video)
    vidor_single   configs[2]  N=64,  T=2000,  80 classes, 50 predicates (16 videos per GPU per step)  [default]
    vidor_val      configs[3]  835 ragged videos N<=64, T<=2000, LPT-sharded per video over the ranks,
                               NCCL all-gather of the top-K triplet records inside the timed region
    stress         configs[4]  N=256, T=4096, K=1024 (1 video per GPU per step)
    baseline_yaml  the reference's SHIPPING mode (configs/baseline.yaml: USE_PPN / USE_DPN off, precomputed [P, F]
                   feature rows as loaded from the h5 files): one Linear(11070 -> 132) + sigmoid over all P pairs of
                   32 configs[0]-shaped videos per step (--precision fp32 = exact order, tensor = tcgen05 tf32)

Ragged videos are packed into batches per chunk class of the pair kernel; every batch of a class replays the
ONE CUDA graph captured for that class's capacity (tspn_b200.batch / serving).  `value` counts P = N(N-1)
ordered pairs per video with the inputs resident in HBM; `e2e` is the same metric through the host-facing
call with pinned host buffers, H2D and D2H inside the timed region.
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

FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback

# name -> (synth.CONFIGS key, BASELINE.json configs index, scaling, default videos per GPU (weak), anchors)
VIDVRD_ANCHORS = ((15.0, 30.0, 45.0, 60.0), 7.5)          # anchor_generator.py:118-120
VIDOR_ANCHORS = ((16.0, 64.0, 256.0, 1024.0), 16.0)
WORKLOADS = {
    "vidvrd_single": ("vidvrd_single", 0, "weak", 64, VIDVRD_ANCHORS),
    "vidvrd_test": ("vidvrd_test", 1, "strong", None, VIDVRD_ANCHORS),
    "vidor_single": ("vidor_single", 2, "weak", 16, VIDOR_ANCHORS),
    "vidor_val": ("vidor_val", 3, "strong", None, VIDOR_ANCHORS),
    "stress": ("stress", 4, "weak", 1, VIDOR_ANCHORS),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vidor_single", choices=sorted(WORKLOADS) + ["baseline_yaml"])
    ap.add_argument("--videos", type=int, default=None,
                    help="videos per GPU per step (weak workloads) / total videos (sharded workloads)")
    ap.add_argument("--precision", default="tensor", choices=["tensor", "fp32"])
    ap.add_argument("--no-sparsify", action="store_true", help="heads on all P pairs (reference quirk Q3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from the host instead of CUDA graphs")
    ap.add_argument("--depth", type=int, default=3, help="batches in flight in the end-to-end serving loop")
    ap.add_argument("--span-proposals", type=int, default=64,
                    help="RELPN.DPN.NUM_DURATION_PROPOSALS: spans kept per pair by the temporal NMS (0 = all decoded)")
    ap.add_argument("--geo-budget-gb", type=float, default=4.0, help="geometry output per batch of ragged videos")
    ap.add_argument("--three-graphs", action="store_true",
                    help="replay the step as three CUDA graphs joined on the host instead of one graph")
    ap.add_argument("--compute-streams", type=int, default=2, choices=[1, 2],
                    help="compute streams of the serving loop (2: the tail of step i overlaps the geometry of i+1)")
    ap.add_argument("--no-affinity", action="store_true", help="do not bind the rank to its GPU's CPU cores")
    ap.add_argument("--reserve-sms", type=int, default=None,
                    help="SMs the persistent all-pairs kernel leaves to the side branches (default: StageConfig's)")
    ap.add_argument("--collective", default="nccl", choices=["nccl", "peer"],
                    help="the per-step exchange of the triplet records in the e2e loop (N > 1, weak scaling): one NCCL "
                         "all-gather, or plain stores into every peer's gather buffer over NVLink (csrc/peer_records.cu)")
    ap.add_argument("--geo-layout", default="dense", choices=["dense", "windowed"],
                    help="layout of the per-frame geometry rows: dense [P, 8, Tp] (the parity layout) or windowed "
                         "(per pair only its overlap window's frames, 7 channels: tspn_pair_geo_viou_windowed)")
    ap.add_argument("--serial-batches", action="store_true",
                    help="`value`: run the batches of a step one after the other on one stream (A/B; default: the "
                         "serving loop's two compute streams alternate, so a batch's side chain overlaps the next "
                         "batch's all-pairs kernel)")
    ap.add_argument("--max-videos", type=int, default=64, help="videos per batch at most")
    ap.add_argument("--no-layout-extra", action="store_true",
                    help="skip the extra run of the same workload with the windowed layout (`windowed_layout` in the line)")
    ap.add_argument("--relationness", default="fp32", choices=["fp32", "tensor"],
                    help="PPNHead arithmetic: fp32 exact order (bit-exact top-K) or tcgen05 (tf32 operands)")
    return ap.parse_args()


def workload_config(args, world: int) -> dict:
    """The `config` object both arms print (same workload, same keys)."""
    from tspn_b200 import synth
    key, idx, scaling, per_gpu, (sizes, stride) = WORKLOADS[args.workload]
    spec = synth.CONFIGS[key]
    if scaling == "weak":
        n_vid = args.videos or per_gpu
        what = "%d video(s) of N=%d T=%d per GPU per step" % (n_vid, spec["n"][1], spec["t"][1])
    else:
        n_vid = args.videos or spec["videos"]
        what = "%d ragged videos N in [%d, %d], T in [%d, %d], LPT-sharded per video over the GPUs" % (
            n_vid, spec["n"][0], spec["n"][1], spec["t"][0], spec["t"][1])
    return {"workload": "%s = BASELINE.json configs[%d]: %s; C=%d R=%d K=%d" % (
                args.workload, idx, what, spec["classes"], spec["predicates"], spec["topk"]),
            "name": args.workload, "videos": n_vid, "classes": spec["classes"], "predicates": spec["predicates"],
            "topk": spec["topk"], "sparsify": not args.no_sparsify, "span_proposals": args.span_proposals,
            "anchor_sizes": list(sizes), "anchor_stride": stride}


def measured_traffic(workload, videos, layout="dense"):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this
    workload (profiles/*_geo_traffic*.json: dram__bytes_read.sum + dram__bytes_write.sum); None when no
    capture describes the requested batch."""
    for name in ("r2_geo_traffic_windowed.json", "r2_geo_traffic.json", "r1_geo_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            if t.get("workload", "vidor_single") != workload or int(t["videos_per_launch"]) != int(videos) or \
                    t.get("geo_layout", "dense") != layout:
                continue
            return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
        except Exception:  # noqa: BLE001
            continue
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
CPU_PAIR_FRAME_BUDGET = 8_500_000        # pair-frames of geometry per step (one VidOR-shaped video is 8.06 M)


def cpu_reference_video(video, sd, topk, sizes, stride, n_keep, pool, cores, pair_sample=None):
    """One video through the CPU path at FULL size (every ordered pair, every frame, no extrapolation) unless
    ``pair_sample`` bounds the geometry to that many pairs (stress: 65 280 pairs x 4096 frames), in which case the
    two geometry stages are scaled by P / sample and the caller says so.  Returns (seconds, per-stage detail).

    vIoU is the reference's own algorithm - `cubic_iou` (trajectory.py:127-141, numpy, one thread) for the
    same-span matrix and the per-pair pure-Python `viou` (evaluation/common.py:65-106) for every ordered pair,
    mapped over all host cores; relationness, sort, classifier and span head are the reference's torch CPU ops
    (oracle.heads.*_ref, all torch threads).  The per-frame geometry channels, the pooled relative block, span
    decode and span NMS have no reference code: the numpy oracle is timed ("oracle, not reference")."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import features as ofeat, geometry as ogeo, heads as oheads
    n, p = video.n_tracklets, video.n_pairs
    pr = ogeo.enumerate_pairs(n)
    det = {}
    sel = np.arange(p)
    scale = 1.0
    if pair_sample is not None and pair_sample < p:
        sel = np.sort(np.random.Generator(np.random.PCG64(0)).choice(p, size=pair_sample, replace=False))
        scale = p / float(pair_sample)
    t0 = time.perf_counter()
    ogeo.cubic_iou_ref(video.boxes, video.boxes)
    det["cubic_iou_matrix"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    chunks = [c for c in np.array_split(sel, 4 * cores) if len(c)]
    list(pool.map(ogeo.viou_pairs_ref, [(video.boxes, video.span, pr[c]) for c in chunks]))
    det["viou_all_pairs"] = (time.perf_counter() - t0) * scale
    # numpy releases the GIL: the pairs are split over all host cores
    t0 = time.perf_counter()
    gchunks = [c for c in np.array_split(sel, max(cores, len(sel) // 64)) if len(c)]
    with ThreadPoolExecutor(max_workers=cores) as ex:
        parts = list(ex.map(lambda c: tuple(x.astype(np.float32) if x.dtype == np.float64 else x for x in
                                            ogeo.pair_geometry(video.boxes, video.span, pr[c, 0], pr[c, 1])), gchunks))
    det["geometry_channels_oracle"] = (time.perf_counter() - t0) * scale
    geo, viou, tiou, ov = (np.concatenate([q[j] for q in parts], axis=0) for j in range(4))
    t0 = time.perf_counter()
    scores = oheads.ppn_head_ref(video.cls, video.cls, sd)
    sc = scores.clone()
    sc.fill_diagonal_(-1.0)                                 # sparsify: survivors are real pairs
    order = torch.sort(sc.view(-1), descending=True)[1][:min(topk, p)].numpy()
    det["relationness_topk"] = time.perf_counter() - t0
    s_i, o_i = order // n, order % n
    rows = s_i * (n - 1) + o_i - (o_i > s_i)
    pos = np.searchsorted(sel, rows)
    pos = np.where((pos < len(sel)) & (sel[np.minimum(pos, len(sel) - 1)] == rows), pos, pos % len(sel))
    t0 = time.perf_counter()
    rel = ofeat.relative_block(geo[pos].astype(np.float64), ov[pos])
    feats = ofeat.assemble_features(video.cls, video.motion, rel, pr[rows]).astype(np.float32)
    det["feature_rows_of_survivors"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        logits = oheads.relation_predictor_ref(feats, sd).numpy()
    det["predicate_head"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        reg = oheads.dpn_head_ref(np.ascontiguousarray(geo[pos], dtype=np.float32), sd).numpy()
    spans = oheads.decode_spans_f64(reg, sizes, stride)
    if n_keep:
        oheads.select_spans(spans, ov[pos], n_keep, 0.5)
    det["span_head_decode_nms"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    oheads.postprocess_ref(logits, video.cls, pr[rows], 20, 200)
    det["postprocess"] = time.perf_counter() - t0
    return sum(det.values()), det


def reference_sample(args, cfgd):
    """The bounded sample of the workload one CPU step processes: (videos, pair_sample or None, description)."""
    from tspn_b200 import synth
    key, _, scaling, _, _ = WORKLOADS[args.workload]
    spec = synth.CONFIGS[key]
    c = spec["classes"]
    if scaling == "weak":
        n, t = spec["n"][1], spec["t"][1]
        vids, budget = [], CPU_PAIR_FRAME_BUDGET
        pf = n * (n - 1) * t
        if pf > budget:          # stress: one video, geometry on a bounded sample of its pairs
            sample = max(256, int(budget // t))
            return [synth.make_video(n, t, c, seed=0)], sample, (
                "1 video of the step (N=%d T=%d, P=%d); vIoU + per-frame geometry on %d sampled pairs scaled to P, "
                "relationness / top-K / heads / records at full size" % (n, t, n * (n - 1), sample))
        count = max(1, min(cfgd["videos"], int(budget // max(pf, 1))))
        vids = [synth.make_video(n, t, c, seed=i) for i in range(count)]
        return vids, None, "%d of the step's %d videos per GPU, every pair and frame (no extrapolation)" % (
            count, cfgd["videos"])
    shapes = synth.config_shapes(key, 0, cfgd["videos"])
    vids, used = [], 0
    for i, (n, t) in enumerate(shapes):
        if used >= CPU_PAIR_FRAME_BUDGET:
            break
        vids.append(synth.make_video(n, t, c, seed=i))
        used += n * (n - 1) * t
    return vids, None, "the first %d of the %d videos (%.1f M pair-frames), every pair and frame" % (
        len(vids), len(shapes), used / 1e6)


def run_reference(args, rank, world):
    """`--impl reference`: rank 0 alone times the CPU path; other ranks exit 0 without work."""
    if rank != 0:
        return
    if args.workload == "baseline_yaml":
        return run_reference_baseline_yaml(args)
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    from tspn_b200 import synth
    cfgd = workload_config(args, world)
    c, r, k = cfgd["classes"], cfgd["predicates"], cfgd["topk"]
    sizes, stride = tuple(cfgd["anchor_sizes"]), cfgd["anchor_stride"]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0)
    vids, pair_sample, what = reference_sample(args, cfgd)
    pairs = sum(v.n_pairs for v in vids)
    times, det_sum = [], {}
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as pool:
        list(pool.map(abs, range(cores)))                      # workers up before the clock starts
        for i in range(args.warmup + args.steps):
            sec, det_sum = 0.0, {}
            for v in vids:
                s, det = cpu_reference_video(v, sd, k, sizes, stride, args.span_proposals, pool, cores, pair_sample)
                sec += s
                for kk, vv in det.items():
                    det_sum[kk] = det_sum.get(kk, 0.0) + vv
            if i >= args.warmup:
                times.append(sec)
    ms = 1e3 * float(np.mean(times))
    value = pairs / (ms / 1e3)
    line = {
        "metric": "tracklet pairs scored/sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": WORKLOADS[args.workload][2], "vs_baseline": None, "dtype": "f32/f64 (numpy + torch CPU)",
        "data": "synthetic", "impl": "reference", "config": cfgd,
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": what + "; vIoU = the reference's per-pair python viou over %d processes + cubic_iou, "
                                          "heads = the reference's torch CPU ops with %d threads, per-frame geometry = numpy "
                                          "oracle over %d threads" % (cores, torch.get_num_threads(), cores),
                         "pairs_per_step": pairs, "detail_s": det_sum},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_reference_baseline_yaml(args):
    """The reference's shipping forward on the CPU: RelationPredictor over every pair of one video's [P, F] rows
    (model.py:53-65, 85-88) + the post-processing of predict.py:66-117, all torch threads."""
    from oracle import geometry as ogeo, heads as oheads
    from tspn_b200 import synth
    c, r, n, v_n = 35, 132, 20, args.videos or 32
    f = synth.feature_dim(c)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_weights(c, r, f, dpn_in=8, seed=0)
    rng = np.random.Generator(np.random.PCG64(0))
    p = n * (n - 1)
    vids = [synth.make_video(n, 8, c, seed=i) for i in range(min(v_n, 8))]
    rows = []
    for _ in vids:
        x = rng.random((p, f), dtype=np.float32)
        x[rng.random((p, f), dtype=np.float32) >= 0.1] = 0.0
        rows.append(x)
    pr = ogeo.enumerate_pairs(n)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for v, x in zip(vids, rows):
            with torch.no_grad():
                logits = oheads.relation_predictor_ref(x, sd).numpy()
            oheads.postprocess_ref(logits, v.cls, pr, 20, 200)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = len(vids) * p / (ms / 1e3)
    emit({"metric": "tracklet pairs scored/sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32 (torch CPU)", "data": "synthetic", "impl": "reference",
          "config": {"workload": "baseline_yaml = configs/baseline.yaml as shipped (USE_PPN / USE_DPN off, precomputed "
                                 "[P, F] rows): %d videos of N=20 (P=380), F=%d, R=%d per GPU per step" % (v_n, f, r),
                     "name": "baseline_yaml", "videos": v_n},
          "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                           "sample": "%d of the step's %d videos, every row; torch CPU Linear + sigmoid with %d threads, "
                                     "python post-processing" % (len(vids), v_n, torch.get_num_threads())},
          "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def alg_bytes_of(shapes) -> int:
    """Algorithmic bytes of the all-pairs kernel for these (N, T) videos."""
    tot = 0
    for n, t in shapes:
        tp, tb, p = (t + 3) // 4 * 4, (t + 7) // 8 * 8, n * max(n - 1, 0)
        tot += 32 * tp * p + 16 * p + 16 * n * tb + 8 * n
    return tot


def alg_bytes_windowed(videos) -> int:
    """The same for the WINDOWED layout (tspn_pair_geo_viou_windowed): per pair 7 channels x the frames of its overlap
    window rounded out to multiples of 4, plus the row offset it reads (8 B)."""
    tot = 0
    for v in videos:
        n, t = v.n_tracklets, v.n_frames
        tb, p = (t + 7) // 8 * 8, n * max(n - 1, 0)
        a = np.maximum(v.span[:, None, 0], v.span[None, :, 0]).astype(np.int64)
        b = np.minimum(v.span[:, None, 1], v.span[None, :, 1]).astype(np.int64)
        lw = np.where(b > a, ((b + 3) & ~3) - (a & ~3), 0)
        np.fill_diagonal(lw, 0)
        tot += 28 * int(lw.sum()) + 24 * p + 16 * n * tb + 8 * n
    return tot


def run_baseline_yaml(args, rank, world, local_rank):
    """configs/baseline.yaml as shipped: both proposal nets off, the [P, F] rows precomputed (vrdataset.py:190-217),
    BaseModel = RelationPredictor over every pair (model.py:53-65) + the records of predict.py:66-117."""
    from tspn_b200 import ops, synth
    from tspn_b200.batch import HostBatch
    from tspn_b200.pipeline import CLS_PREFIX, PairStage, StageConfig
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ops.require_device()
    c, r, n, v_n = 35, 132, 20, args.videos or 32
    f = synth.feature_dim(c)
    prec = args.precision if "--precision" in sys.argv else "fp32"
    cfg = StageConfig(n_classes=c, n_predicates=r, use_ppn=False, use_dpn=False, sparsify=False, precision=prec)
    sd = synth.make_weights(c, r, f, dpn_in=8, seed=0)
    stage = PairStage(cfg)
    stage.load_weights(sd, dev)
    vids = [synth.make_video(n, 8, c, seed=1000 * rank + i) for i in range(v_n)]
    host = HostBatch([np.zeros((n, 1, 4), np.float32)] * v_n, [np.tile(np.array([[0, 1]], np.int32), (n, 1))] * v_n,
                     [v.cls for v in vids], None)
    batch = host.to_device(dev)
    p_tot = v_n * n * (n - 1)
    rng = np.random.Generator(np.random.PCG64(rank))
    ld = ops.padded(f, 4)
    feats_host = torch.zeros((p_tot, ld), dtype=torch.float32).pin_memory()
    x = rng.random((p_tot, f), dtype=np.float32)
    x[rng.random((p_tot, f), dtype=np.float32) >= 0.1] = 0.0               # sparse rows, like the h5 feats
    feats_host[:, :f] = torch.from_numpy(x)
    feats_dev = feats_host.to(dev)
    graphed = stage.capture(batch, features=feats_dev[:, :f])
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        graphed.replay()
    torch.cuda.synchronize()
    step_ms = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graphed.replay()
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    # the dominant kernel on its own stream position: the classifier over the resident rows
    wt, bias = stage.w[CLS_PREFIX + "weight"], stage.w[CLS_PREFIX + "bias"]
    head_ms = []
    for i in range(3 + args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.predicate_head(feats_dev[:, :f], wt, bias, precision=prec, packed=stage.packed_cls)
        e1.record()
        e1.synchronize()
        if i >= 3:
            head_ms.append(e0.elapsed_time(e1))
    # e2e: the rows cross PCIe every step (they are what the reference loads from disk), logits + records come back
    res = graphed.result
    out_host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in res.host_outputs().items()}
    def e2e_step():
        feats_dev.copy_(feats_host, non_blocking=True)
        batch.copy_from(host)
        r2 = graphed.replay()
        for k, v in r2.host_outputs().items():
            out_host[k].copy_(v, non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(2):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    if rank != 0:
        return
    ms = float(np.mean(step_ms))
    peak, peak_src = measured_peak()
    nbytes = p_tot * f * 4 + r * f * 4 + p_tot * r * 4
    hm = float(np.mean(head_ms))
    line = {
        "metric": "tracklet pairs scored/sec", "value": world * p_tot / (ms / 1e3), "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 exact-order" if prec == "fp32" else "tf32 tcgen05 on fp32 rows", "data": "synthetic",
        "config": {"workload": "baseline_yaml = configs/baseline.yaml as shipped (USE_PPN / USE_DPN off, precomputed "
                               "[P, F] rows): %d videos of N=20 (P=380), F=%d, R=%d per GPU per step" % (v_n, f, r),
                   "name": "baseline_yaml", "videos": v_n, "precision": prec,
                   "l2": "256 MiB buffer zeroed between timed iterations; the rows are %.2f GB" % (p_tot * f * 4 / 1e9)},
        "roofline": {"bound": "hbm", "kernel": "predicate_%s_kernel (timed on its own, same inputs, L2 flushed)"
                                               % ("exact" if prec == "fp32" else "tc"),
                     "achieved": nbytes / hm / 1e6, "peak": peak, "unit": "GB/s", "frac": nbytes / hm / 1e6 / peak,
                     "traffic": None, "peak_source": peak_src, "avg_launch_ms": hm, "share_of_step": hm / ms,
                     "tflops": 2.0 * p_tot * f * r / hm / 1e9,
                     "note": "fp32 exact order is a CUDA-core GEMM with a fixed k-ascending fma chain (bit-exact "
                             "scores): compute-bound far below the HBM roofline by construction; --precision tensor is "
                             "the HBM-bound form"},
        "e2e": {"value": world * p_tot * args.steps / e2e_s, "unit": "pairs/s",
                "h2d_bytes_per_step": int(feats_host.numel() * 4 + host.h2d_bytes()),
                "d2h_bytes_per_step": int(sum(t.numel() * t.element_size() for t in out_host.values()))},
        "gpu_launches": int(graphed.kernels_per_replay * args.steps), "launches_per_step": int(graphed.kernels_per_replay),
        "clocks": clocks,
    }
    emit(line)


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from tspn_b200 import _lib, affinity, ops, sharding, synth
    from tspn_b200.pipeline import PairStage, StageConfig
    from tspn_b200.serving import PipelinedStage, bucket_key, host_batches_for

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ops.require_device()
    bound = None if args.no_affinity else affinity.bind_to_gpu(local_rank, world)
    cfgd = workload_config(args, world)
    key, _, scaling, _, _ = WORKLOADS[args.workload]
    c, r, k = cfgd["classes"], cfgd["predicates"], cfgd["topk"]
    sizes, stride = tuple(cfgd["anchor_sizes"]), cfgd["anchor_stride"]
    sparsify = not args.no_sparsify
    cfg = StageConfig(n_classes=c, n_predicates=r, topk=k, use_ppn=True, use_dpn=True, sparsify=sparsify,
                      precision=args.precision, anchor_sizes=sizes, anchor_stride=stride,
                      num_span_proposals=args.span_proposals, relationness_precision=args.relationness,
                      geo_layout=args.geo_layout)
    if args.reserve_sms is not None:
        cfg.geo_reserve_sms = args.reserve_sms
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0)
    stage = PairStage(cfg)
    stage.load_weights(sd, dev)

    # ---- this rank's videos ------------------------------------------------------------------------
    spec = synth.CONFIGS[key]
    shards = None
    if scaling == "weak":
        videos = [synth.make_video(spec["n"][1], spec["t"][1], c, seed=100000 * rank + i) for i in range(cfgd["videos"])]
        all_shapes = [(v.n_tracklets, v.n_frames) for v in videos] * world
        my_ids = list(range(len(videos)))
        imbalance = 1.0
    else:
        all_shapes = synth.config_shapes(key, 0, cfgd["videos"])
        shards = sharding.shard_videos(all_shapes, world)
        imbalance = sharding.imbalance(all_shapes, shards)
        my_ids = shards[rank]
        videos = [synth.make_video(all_shapes[i][0], all_shapes[i][1], c, seed=i) for i in my_ids]
    pairs_global = sum(n * max(n - 1, 0) for n, _ in all_shapes)
    hosts, batch_vids, caps = host_batches_for(videos, c, geo_budget_bytes=int(args.geo_budget_gb * (1 << 30)),
                                               max_videos=args.max_videos)
    residents = [h.to_device(dev) for h in hosts]            # the step's inputs, resident in HBM
    templates, seen = [], set()
    for h in hosts:
        if bucket_key(h) not in seen:
            seen.add(bucket_key(h))
            templates.append(h)

    # The serving loop owns, per capacity bucket, `depth` slots (device inputs + captured CUDA graph + pinned result
    # buffers); the resident-input measurement replays the same graphs on inputs already in HBM (device-to-device
    # refill of the slot's input arena).
    group = dist.group.WORLD if (world > 1 and scaling == "weak") else None
    pipe = PipelinedStage(stage, templates, device=dev, depth=args.depth, graphs=not args.eager, group=group,
                          compute_streams=args.compute_streams, single_graph=not args.three_graphs,
                          collective=args.collective)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)    # > 126 MB L2
    n_local = len(videos)
    tpv = cfg.topk_per_video
    rec_local = torch.zeros((max(n_local, 1), tpv, 8), dtype=torch.int32, device=dev)
    cnt_local = torch.zeros(max(n_local, 1), dtype=torch.int32, device=dev)
    batch_idx = [torch.as_tensor(v, dtype=torch.int64, device=dev) for v in batch_vids]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    geo_acc = {"ms": 0.0, "launches": 0}
    pending = {}                                        # slot -> its pair-kernel events not read yet

    def read_geo(slot):
        ev = pending.pop(slot, None)
        if ev is not None:
            ev[1].synchronize()
            geo_acc["ms"] += ev[0].elapsed_time(ev[1])
            geo_acc["launches"] += 1

    def gather_step():
        if shards is not None and world > 1:
            return sharding.gather_records(rec_local[:n_local], cnt_local[:n_local], shards)
        return rec_local, cnt_local

    # Resident refill: a bucket's slots are shared by every batch of its capacity, so a batch's inputs (already in
    # HBM) are copied device-to-device into the slot it will replay on - on a copy stream, up to `depth` batches
    # ahead, so the refill of batch j+1 runs under the kernels of batch j.  A workload of one batch per step keeps
    # its inputs in the slot: nothing is copied.
    s_copy = torch.cuda.Stream(dev)
    holds = {}                                          # slot -> the resident batch its inputs currently are

    # A step of several batches alternates the serving loop's two compute streams, exactly as PipelinedStage.submit
    # does: batches are independent (different slots, different buffers), so the latency-bound side chain of batch j
    # (survivor rows -> heads -> records) runs underneath the all-pairs kernel of batch j + 1 instead of in front of
    # it.  The step's clock (events on the caller's stream) starts before the fork and stops after the join.
    lanes = pipe.compute if (len(hosts) > 1 and len(pipe.compute) > 1 and not args.serial_batches) else None

    def one_step():
        main = torch.cuda.current_stream(dev)
        used, plan = {}, []
        for j, (resident, host) in enumerate(zip(residents, hosts)):
            bucket = pipe._bucket(host)
            i = used.get(id(bucket), 0)
            used[id(bucket)] = i + 1
            plan.append((j, resident, bucket.slots[i % len(bucket.slots)]))
        if lanes is not None:
            for st in lanes:
                st.wait_stream(main)                    # fork
        for j, resident, slot in plan:
            st = lanes[j % len(lanes)] if lanes is not None else main
            with torch.cuda.stream(st):
                if holds.get(slot) is not resident:
                    with torch.cuda.stream(s_copy):
                        s_copy.wait_event(slot.kernels_done)    # the slot's previous replay has read its inputs
                        slot.batch.copy_from_device(resident)
                        slot.h2d_done.record(s_copy)
                    holds[slot] = resident
                    st.wait_event(slot.h2d_done)
                else:
                    slot.batch._adopt(resident.host)
                st.wait_event(slot.kernels_done)        # the slot's previous replay (the other lane's, possibly)
                read_geo(slot)                          # its events are re-recorded by the next replay
                timers = {}
                res = slot.graphed.replay(timers=timers) if slot.graphed is not None else \
                    stage.forward(slot.batch, timers=timers)
                pending[slot] = timers["geo"]
                if shards is not None:
                    nr = len(batch_vids[j])
                    rec_local.index_copy_(0, batch_idx[j], res.records[:nr])
                    cnt_local.index_copy_(0, batch_idx[j], res.record_counts[:nr])
                slot.kernels_done.record(st)
        if lanes is not None:
            for st in lanes:
                main.wait_stream(st)                    # join
        return gather_step()

    def drain_geo():
        for slot in list(pending):
            read_geo(slot)

    step_ms = []
    sampler = ClockSampler(local_rank)
    sampler.start()                                     # samples through warm-up, timed steps and e2e
    launches0 = ops.launch_count()
    for i in range(args.warmup):
        one_step()
    barrier()
    drain_geo()
    launches_per_step = (ops.launch_count() - launches0) // max(args.warmup, 1)
    geo_acc["ms"], geo_acc["launches"] = 0.0, 0
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                                   # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step()
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        drain_geo()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    geo_ms_total, geo_launches = geo_acc["ms"], geo_acc["launches"]
    # the pair kernel on its own (largest batch, nothing else on the device): what the co-resident side branch costs it
    big = max(range(len(hosts)), key=lambda j: int(hosts[j].actual[_lib.TOT_GEO_FLOATS]))
    bslot = pipe._bucket(hosts[big]).slots[0]
    bslot.batch.copy_from_device(residents[big])
    geom_alone = bslot.graphed.result.geom if bslot.graphed is not None else \
        ops.pair_geometry_outputs(bslot.batch, windowed=args.geo_layout == "windowed")
    alone_ms = []
    for i in range(3 + min(args.steps, 10)):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.pair_geometry_phase(bslot.batch, geom_alone, _lib.GEO_PHASE_MAIN)
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
    value = pairs_global / (ms_per_step / 1e3)

    # ---- end-to-end steps: pinned host in, pinned host out ------------------------------------
    # tspn_b200.serving.PipelinedStage (the host-facing call): every batch of every step is copied from pinned
    # host memory and its results are copied back to pinned host memory inside the timed region; H2D of batch
    # i+1 and D2H of batch i-1 overlap the kernels of batch i (three streams, --depth batches in flight).  The
    # loop runs warm-up + timed steps back to back (a serving loop does not drain between requests); the clock
    # starts when the last warm-up step's results are on the host and stops when the last timed step's are.
    rec_host = torch.empty((len(all_shapes) if shards is not None else 1, tpv, 8), dtype=torch.int32).pin_memory()
    s_gather = torch.cuda.Stream(dev)
    nb = len(hosts)
    ring_n = args.depth + 2                             # steps whose records can be in flight at once
    rings = [(torch.zeros_like(rec_local), torch.zeros_like(cnt_local)) for _ in range(ring_n)] \
        if shards is not None else None

    def e2e_steps(n_warm, n_timed):
        step_events, gather_done = {}, {}

        def post(q, res):                               # on the batch's compute stream, right behind its kernels
            if shards is None:
                return
            s, j = divmod(q, nb)
            old = gather_done.pop(s - ring_n, None)     # the ring slot's previous step has left the device
            if old is not None:
                torch.cuda.current_stream(dev).wait_event(old)
            rl, cl = rings[s % ring_n]
            nr = len(batch_vids[j])
            rl.index_copy_(0, batch_idx[j], res.records[:nr])
            cl.index_copy_(0, batch_idx[j], res.record_counts[:nr])
            ev = torch.cuda.Event()
            ev.record()
            step_events.setdefault(s, []).append(ev)

        t_start, touched, q_out = None, 0, 0
        total = (n_warm + n_timed) * nb
        for out in pipe.run((hosts[q % nb] for q in range(total)), post=post):
            touched += int(out["record_counts"].shape[0])          # touch the host result
            q_out += 1
            if q_out % nb:
                continue
            s = q_out // nb - 1                         # every batch of step s is back on the host
            if shards is not None:                      # the step's one collective + the gathered records' D2H
                with torch.cuda.stream(s_gather):
                    for ev in step_events.pop(s):
                        s_gather.wait_event(ev)
                    rl, cl = rings[s % ring_n]
                    ra = sharding.gather_records(rl[:n_local], cl[:n_local], shards)[0] if world > 1 else rl[:n_local]
                    rec_host[:ra.shape[0]].copy_(ra, non_blocking=True)
                    gd = torch.cuda.Event()
                    gd.record()
                    gather_done[s] = gd
            if s == n_warm - 1:                         # warm-up over: the clock starts with the pipeline full
                s_gather.synchronize()
                t_start = time.perf_counter()
        s_gather.synchronize()
        return time.perf_counter() - t_start, touched

    e2e_s, touched = e2e_steps(max(args.warmup, 2), args.steps)
    assert touched > 0
    if world > 1:
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = pairs_global * args.steps / e2e_s
    h2d = int(sum(h.h2d_bytes() for h in hosts))
    d2h = int(pipe.d2h_total_bytes() // max(max(args.warmup, 2) + args.steps, 1)) + \
        (int(rec_host.numel() * 4) if shards is not None else 0)
    clocks = sampler.stop()
    if world > 1:                                       # per-step bytes of the busiest rank
        tt = torch.tensor([h2d, d2h], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        h2d, d2h = int(tt[0].item()), int(tt[1].item())

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (all-pairs geometry + vIoU) ------------------------------
    my_shapes = [(v.n_tracklets, v.n_frames) for v in videos]
    windowed = args.geo_layout == "windowed"
    alg_bytes = alg_bytes_windowed(videos) if windowed else alg_bytes_of(my_shapes)      # per step of this rank
    geo_ms_step = geo_ms_total / args.steps
    achieved = alg_bytes / (geo_ms_step / 1e3) / 1e9
    peak, peak_src = measured_peak()
    alone_bytes = alg_bytes_windowed([videos[i] for i in batch_vids[big]]) if windowed else \
        alg_bytes_of([my_shapes[i] for i in batch_vids[big]])
    alone_gbs = alone_bytes / (float(np.mean(alone_ms)) / 1e3) / 1e9
    cfgd.update({
        "batches_per_step_per_gpu": len(hosts), "pairs_per_step": pairs_global,
        "capacities": {str(cc): {"videos": cap.videos, "pairs": cap.pairs, "geo_gb": cap.geo_floats * 4 / 1e9,
                                 "geo_chunk": cap.geo_chunk, "max_n": cap.max_n, "max_t": cap.max_t}
                       for cc, cap in caps.items()},
        "precision": args.precision, "relationness_precision": args.relationness, "geo_layout": args.geo_layout,
        "geo_reserve_sms": cfg.geo_reserve_sms,
        "sharding": ("per video, LPT on N(N-1)T (imbalance %.4f); NCCL all-gather of the [V,200,8] int32 triplet "
                     "records inside the timed region" % imbalance) if shards is not None else
                    "per video, every rank its own videos; no data-path collective in `value`, all-gather of the "
                    "records per step in the e2e loop",
        "l2": "256 MiB buffer zeroed between timed iterations (untimed); each step also writes %.2f GB of outputs"
              % (alg_bytes / 1e9),
        "wall_s_timed_region": t_wall,
        "batch_lanes": len(lanes) if lanes is not None else 1,
        "launch": "eager C-ABI calls" if args.eager else
                  ("one CUDA-graph launch per batch; the graph was captured once per capacity bucket and serves "
                   "every ragged batch packed for it; three branches inside: the persistent all-pairs kernel | "
                   "relationness -> top-K -> surviving-pair rows -> predicate head -> records | per-tracklet terms, "
                   "volumes, span NMS, vIoU finalize"),
        "e2e_pipeline": "tspn_b200.serving.PipelinedStage, depth %d, %d compute stream(s): per batch the used part "
                        "of the pinned input arena goes H2D, results D2H; warm-up and timed steps run back to back, "
                        "the clock covers exactly the timed steps" % (args.depth, args.compute_streams),
        "h2d_transport": "u16 boxes + u8 motion counts (lossless for these inputs), expanded on the device",
        "cpu_affinity": bound, "collective": args.collective if group is not None else None,
    })
    line = {
        "metric": "tracklet pairs scored/sec", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None,
        "dtype": "f32 geometry/features, f64 volume sums, %s heads" % ("bf16 tcgen05" if args.precision == "tensor"
                                                                         else "f32 exact-order"),
        "data": "synthetic", "config": cfgd,
        "roofline": {"bound": "hbm", "kernel": "pair_geo_kernel (CUDA events immediately around each launch)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(args.workload, cfgd["videos"], args.geo_layout),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_step": alg_bytes, "launches_per_step": geo_launches / args.steps,
                     "avg_launch_ms": geo_ms_total / max(geo_launches, 1),
                     "share_of_step": geo_ms_step / float(np.mean(step_ms)),
                     "note": "achieved / frac are measured inside the timed steps, where the side branches' kernels "
                             "co-reside with this kernel on every SM; `alone` is the largest batch's launch with an "
                             "idle device",
                     "alone": {"avg_launch_ms": float(np.mean(alone_ms)), "achieved": alone_gbs,
                               "frac": alone_gbs / peak}},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches_per_step * args.steps),
        "launches_per_step": int(launches_per_step),
        "clocks": clocks,
    }
    if windowed:
        line["roofline"]["kernel"] = "pair_geo_windowed_kernel (CUDA events immediately around each launch)"
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_subprocess(args)
    if world == 1 and not windowed and not args.no_layout_extra and args.precision == "tensor" and sparsify:
        line["windowed_layout"] = windowed_subprocess(args)
    emit(line)


def cpu_baseline_subprocess(args):
    """The CPU leg in a child process (it forks worker processes: not something to do under a live CUDA context)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "1",
           "--warmup", "1", "--span-proposals", str(args.span_proposals)]
    if args.videos:
        cmd += ["--videos", str(args.videos)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
        return json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}


def windowed_subprocess(args):
    """The same workload with the opt-in WINDOWED geometry layout (tspn_pair_geo_viou_windowed), in a child process
    after this one's measurements: reported beside the dense (parity layout) numbers, never instead of them."""
    cmd = [sys.executable, os.path.abspath(__file__), "--geo-layout", "windowed", "--no-cpu-baseline",
           "--workload", args.workload, "--steps", str(args.steps), "--warmup", str(args.warmup),
           "--span-proposals", str(args.span_proposals), "--precision", args.precision, "--depth", str(args.depth),
           "--relationness", args.relationness, "--geo-budget-gb", str(args.geo_budget_gb)]
    if args.videos:
        cmd += ["--videos", str(args.videos)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        d = json.loads(out.stdout.strip().splitlines()[-1])
        r = d["roofline"]
        return {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "e2e": d["e2e"],
                "roofline": {k: r[k] for k in ("kernel", "achieved", "peak", "unit", "frac", "traffic",
                                               "algorithmic_bytes_per_step", "avg_launch_ms", "share_of_step", "alone")},
                "note": "geo rows as [7][Lw] per pair (only the overlap window's frames; bit-identical to the dense rows "
                        "there, tests/test_gpu_windowed.py); every other output of the step is the same tensor"}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "note": "failed: %r" % (e,)}


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
        if args.workload == "baseline_yaml":
            run_baseline_yaml(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
