"""Stand-alone timing of tspn_survivor_rows on the bench workload (no concurrent pair kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tspn_b200 import ops, synth  # noqa: E402
from tspn_b200.batch import HostBatch  # noqa: E402
from tspn_b200.pipeline import DPN_PREFIX, PairStage, StageConfig  # noqa: E402

spec = synth.CONFIGS["vidor_single"]
c, r, k = spec["classes"], spec["predicates"], spec["topk"]
n, t = spec["n"][0], spec["t"][0]
videos = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = StageConfig(n_classes=c, n_predicates=r, topk=k, sparsify=True, precision="tensor",
                  anchor_sizes=(16.0, 64.0, 256.0, 1024.0), anchor_stride=16.0)
stage = PairStage(cfg)
stage.load_weights(synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0), "cuda")
batch = HostBatch.from_videos([synth.make_video(n, t, c, seed=i) for i in range(videos)], compact=True).to_device("cuda")
scores = ops.relationness(batch, stage.ppn_weights())
idx, val, row = ops.topk_pairs(batch, scores, k, exclude_diagonal=True)
sw = tuple(stage.w[DPN_PREFIX + key] for key in ("conv.weight", "conv.bias", "duration_pred.weight", "duration_pred.bias"))
ov = ops.pair_geometry(batch, write_geo=False)["overlap"]
sel = row.reshape(-1)
win = (ov[sel.clamp_min(0), 1] - ov[sel.clamp_min(0), 0]).float()
print("rows %d, mean overlap window %.0f frames, max %d" % (sel.numel(), win.mean().item(), int(win.max().item())))
for with_spans in (True, False):
    for _ in range(3):
        ops.survivor_rows(batch, row, span_weights=sw if with_spans else None, sizes=stage.sizes_dev, stride=16.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        ops.survivor_rows(batch, row, span_weights=sw if with_spans else None, sizes=stage.sizes_dev, stride=16.0)
    e1.record()
    torch.cuda.synchronize()
    print("survivor_rows%s: %.1f us per launch (includes output allocation)" % (" + spans" if with_spans else "",
                                                                               1e3 * e0.elapsed_time(e1) / 20))
