"""Micro-benchmark of the tensor-core span head (DPNHead at the reference's default width).

    python tools/bench_span_head.py [K] [Cin] [T]

Prints the CUDA-event time of the whole call (weight + input pre-passes, the implicit-GEMM kernel, the finish pass) and
the achieved bf16 TFLOP/s against MEASURED_PEAKS.json, for the one-CTA form (default) and the opt-in CTA-pair form
(``TSPN_SPAN_HEAD_PAIR=1``).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tspn_b200 import ops, synth  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cin = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
t = int(sys.argv[3]) if len(sys.argv) > 3 else 300
a = 4
sd = synth.make_weights(35, 132, 16, dpn_in=cin, n_anchors=a, seed=0)
p = "relpn.duration_proposal_network.dpn_head."
args = [torch.from_numpy(sd[p + n]).cuda() for n in ("conv.weight", "conv.bias", "duration_pred.weight",
                                                     "duration_pred.bias")]
x = torch.randn((k, cin, t), device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for form in ("one_cta", "cta_pair"):
    prec = "tensor"
    os.environ["TSPN_SPAN_HEAD_PAIR"] = "1" if form == "cta_pair" else "0"
    for _ in range(3):
        ops.span_head(x, *args, precision=prec)
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.span_head(x, *args, precision=prec)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    flop = 2.0 * k * t * (3 * cin * cin + 2 * a * cin)
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, sustained = pk["bf16_tflops"], pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    except Exception:  # noqa: BLE001
        peak = sustained = 1590.0
    print(json.dumps({"op": "span_head", "precision": prec, "form": form, "k": k, "cin": cin, "t": t, "ms": ms,
                      "tflops": flop / ms / 1e9, "peak_tflops": peak, "frac": flop / ms / 1e9 / peak,
                      "frac_of_sustained": flop / ms / 1e9 / sustained}))
