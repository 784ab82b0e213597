"""Micro-benchmark of the tensor-core predicate head at the sizes where it is HBM-bound on its rows.

    python tools/bench_predicate.py [rows] [F] [R]        # default: 16 VidOR videos x 4032 pairs, F=11160, R=50

The op reads 2 F (bf16 rows) or 4 F (fp32 rows, tf32 MMA) bytes per row and does 2 F R flop on them: 25-66 flop/B,
far below the ridge, so the roofline is HBM read bandwidth (MEASURED_PEAKS.json hbm_gbs).  One JSON line per form.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tspn_b200 import ops  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 16 * 4032
f = int(sys.argv[2]) if len(sys.argv) > 2 else 11160
r = int(sys.argv[3]) if len(sys.argv) > 3 else 50
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    peak = 6650.0
w = (torch.randn((r, f), device="cuda") * 0.01).contiguous()
b = torch.zeros(r, device="cuda")
packed = ops.pack_predicate_weights(w)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for name, dtype, ld_mult in (("bf16 rows (kind::f16)", torch.bfloat16, 8), ("fp32 rows (kind::tf32)", torch.float32, 4)):
    ld = (f + ld_mult - 1) // ld_mult * ld_mult
    x = torch.randn((m, ld), device="cuda", dtype=torch.float32).to(dtype)[:, :f]
    for _ in range(3):
        ops.predicate_head(x, w, b, precision="tensor", packed=packed)
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.predicate_head(x, w, b, precision="tensor", packed=packed)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    nbytes = m * f * x.element_size() + m * r * 4 + r * f * x.element_size()
    print(json.dumps({"op": "predicate_head", "form": name, "rows": m, "F": f, "R": r, "ms": ms,
                      "GBps": nbytes / ms / 1e6, "peak_GBps": peak, "frac_of_hbm": nbytes / ms / 1e6 / peak,
                      "TFLOPs": 2.0 * m * f * r / ms / 1e9}))
    del x
