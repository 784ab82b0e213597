"""Latency of the drop-in call: ``BaseModel.forward([PairList])`` on ONE video, CPU tensors in, CPU tensors out - the
call lib/modeling/predict.py:57 makes (``TEST_BATCH_SIZE: 1``).  Per-video problems are launch- and copy-latency
bound on a B200 ("Hard parts"): this is the number to read beside the batched throughput.

    python tools/latency_basemodel.py [--calls 200]

One JSON line per configuration: p50 / p99 / mean milliseconds per call and the pairs/s they amount to.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tspn_b200  # noqa: E402
from tspn_b200 import synth  # noqa: E402
from tspn_b200.list_pair import PairList  # noqa: E402
from tspn_b200.model import BaseModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calls", type=int, default=200)
    args = ap.parse_args()
    cases = [("vidvrd_single (configs[0]) N=20 T=300", 20, 300, 35, 132, (15.0, 30.0, 45.0, 60.0), 7.5),
             ("vidor_single (configs[2]) N=64 T=2000", 64, 2000, 80, 50, (16.0, 64.0, 256.0, 1024.0), 16.0)]
    for name, n, t, c, r, sizes, stride in cases:
        for mode in ("baseline.yaml: PPN/DPN off, precomputed [P,F] rows, fp32",
                     "full pair stage from tracklets: PPN+DPN, sparsify, tensor heads, span NMS"):
            cfg = tspn_b200.get_default_cfg()
            fdim = synth.feature_dim(c)
            cfg.PREDICT.OBJECT_NUM, cfg.PREDICT.PREDICATE_NUM, cfg.PREDICT.FEATURE_DIM = c, r, fdim
            cfg.RELPN.PPN.IN_CHANNELS = cfg.RELPN.PPN.OUT_CHANNELS = c
            cfg.RELPN.DPN.IN_CHANNELS = 8
            cfg.RELPN.DPN.ANCHOR_SIZES, cfg.RELPN.DPN.ANCHOR_STRIDE = list(sizes), stride
            full = mode.startswith("full")
            cfg.RELPN.USE_PPN = cfg.RELPN.USE_DPN = full
            if full:
                cfg.PREDICT.PRECISION, cfg.PREDICT.SPARSIFY = "tensor", True
            sd = {k: torch.from_numpy(v) for k, v in synth.make_weights(c, r, fdim, dpn_in=8, seed=0).items()}
            model = BaseModel(cfg).eval()
            model.load_state_dict(sd)
            vids = [synth.make_video(n, t, c, seed=s) for s in range(4)]
            if full:
                pls = [PairList.from_tracklets(v.boxes, v.span, v.cls, v.motion) for v in vids]
            else:
                rng = np.random.Generator(np.random.PCG64(0))
                pls = []
                for v in vids:
                    pl = PairList(torch.from_numpy(rng.random((n * (n - 1), fdim), dtype=np.float32)))
                    pl.add_field("track_cls_logits", torch.from_numpy(v.cls))
                    pl.add_field("num_tracklets", n)
                    pls.append(pl)
            times = []
            with torch.no_grad():
                for i in range(args.calls + 10):
                    t0 = time.perf_counter()
                    pp, dp, logits = model([pls[i % 4]], None)
                    _ = float(logits[0][0, 0])                  # the result is on the host
                    if i >= 10:
                        times.append(time.perf_counter() - t0)
            ms = 1e3 * np.asarray(times)
            print(json.dumps({"call": "BaseModel.forward([PairList]) CPU in / CPU out, one video per call",
                              "video": name, "mode": mode, "calls": args.calls, "p50_ms": float(np.percentile(ms, 50)),
                              "p99_ms": float(np.percentile(ms, 99)), "mean_ms": float(ms.mean()),
                              "pairs_per_s_at_p50": n * (n - 1) / (float(np.percentile(ms, 50)) / 1e3)}), flush=True)


if __name__ == "__main__":
    main()
