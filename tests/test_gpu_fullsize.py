"""BASELINE.json's full-size configurations on the GPU, checked through size-independent properties
(the oracle needs minutes for them) plus the oracle on a seeded SAMPLE of pairs:

* configs[2] VidOR-shaped video N=64, T=2000 (C=80, R=50, K=256) - the bench workload;
* configs[4] stress N=256, T=4096, K=1024 (65 280 ordered pairs, 8.6 GB of geometry);
* configs[1] / configs[3] ragged batches (VidVRD-test-shaped N<=40, T<=1200; VidOR-val-shaped N<=64,
  T<=2000) through the whole stage, oracle on a sample of videos.

Properties used: vIoU / tIoU / overlap are symmetric in (s, o); the per-frame IoU and mask channels are
symmetric, the log-ratio channels antisymmetric; the mask sums to the overlap length; the reductions of the
geometry-writing and reductions-only kernels are bit-identical; the matrix form ``cubic_iou`` agrees on
full-span inputs; top-K is sorted, duplicate-free and idempotent; spans are ordered and inside [0, T];
records are sorted by score.
"""
import numpy as np
import pytest
import torch

from oracle import exact, geometry as ogeo
from tspn_b200 import _lib, ops, synth
from tspn_b200.batch import HostBatch
from tspn_b200.pipeline import PairStage, StageConfig

pytestmark = pytest.mark.gpu


def _rows(n):
    pr = ogeo.enumerate_pairs(n)
    s, o = pr[:, 0], pr[:, 1]
    rev = o * (n - 1) + s - (s > o)            # row of (o, s)
    return pr, rev


def _check_geometry_properties(v, batch, out, sample=48, seed=0, rtol=1e-5):
    n, t = v.n_tracklets, v.n_frames
    pr, rev = _rows(n)
    viou, tiou, ov = out["viou"].cpu().numpy(), out["tiou"].cpu().numpy(), out["overlap"].cpu().numpy()
    np.testing.assert_array_equal(viou, viou[rev])
    np.testing.assert_array_equal(tiou, tiou[rev])
    np.testing.assert_array_equal(ov, ov[rev])
    a = np.maximum(v.span[pr[:, 0], 0], v.span[pr[:, 1], 0])
    b = np.minimum(v.span[pr[:, 0], 1], v.span[pr[:, 1], 1])
    has = b > a
    np.testing.assert_array_equal(ov[:, 0], np.where(has, a, 0))          # frame bounds bit-exact
    np.testing.assert_array_equal(ov[:, 1], np.where(has, b, 0))
    assert (viou >= 0).all() and (viou <= 1).all()
    geo = batch.geo_rows(out["geo"], 0)                                    # [P, 8, Tp] on the device
    mask_sum = geo[:, 7].sum(dim=1).cpu().numpy()
    np.testing.assert_array_equal(mask_sum, np.where(has, b - a, 0).astype(np.float32))
    rev_t = torch.from_numpy(rev).to(geo.device)
    assert torch.equal(geo[:, 4], geo[rev_t, 4])                           # per-frame IoU symmetric
    assert torch.equal(geo[:, 7], geo[rev_t, 7])
    for ch in (2, 3):                                                      # log(a/b) = -log(b/a)
        d = (geo[:, ch] + geo[rev_t, ch]).abs().max().item()
        assert d <= 2e-6, (ch, d)
    assert float(geo[:, :, t:].abs().max()) == 0.0 if geo.shape[2] > t else True
    # the oracle on a seeded sample of pairs
    rng = np.random.Generator(np.random.PCG64(seed))
    sel = np.sort(rng.choice(pr.shape[0], size=min(sample, pr.shape[0]), replace=False))
    w_geo, w_viou, w_tiou, w_ov = ogeo.pair_geometry(v.boxes, v.span, pr[sel, 0], pr[sel, 1])
    got = geo[torch.from_numpy(sel).to(geo.device)].cpu().numpy()[:, :, :t]
    np.testing.assert_allclose(got, w_geo, rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(viou[sel], w_viou, rtol=rtol, atol=0)
    np.testing.assert_allclose(tiou[sel], w_tiou, rtol=1e-6, atol=0)
    np.testing.assert_array_equal(ov[sel], w_ov)


@pytest.mark.parametrize("name,sample", [("vidor_single", 64), ("stress", 24)])
def test_geometry_full_size_properties(name, sample):
    spec = synth.CONFIGS[name]
    n, t = spec["n"][0], spec["t"][0]
    v = synth.make_video(n, t, spec["classes"], seed=17)
    batch = HostBatch.from_videos([v]).to_device("cuda")
    out = ops.pair_geometry(batch, write_geo=True)
    red = ops.pair_geometry(batch, write_geo=False)
    torch.cuda.synchronize()
    assert torch.equal(out["viou"], red["viou"]) and torch.equal(out["overlap"], red["overlap"])
    _check_geometry_properties(v, batch, out, sample=sample)
    del out, red
    torch.cuda.empty_cache()


def test_cubic_iou_matrix_agrees_with_pair_kernel_full_span():
    v = synth.make_video(64, 2000, 80, seed=23, full_span=True)
    batch = HostBatch.from_videos([v]).to_device("cuda")
    out = ops.pair_geometry(batch, write_geo=False)
    m = ops.cubic_iou(torch.from_numpy(v.boxes).cuda(), torch.from_numpy(v.boxes).cuda())
    torch.cuda.synchronize()
    pr, _ = _rows(64)
    m = m.cpu().numpy()
    # both kernels sum exactly (integer boxes): identical fp32 results, symmetric, unit diagonal
    np.testing.assert_array_equal(out["viou"].cpu().numpy(), m[pr[:, 0], pr[:, 1]])
    np.testing.assert_array_equal(m, m.T)
    np.testing.assert_array_equal(np.diag(m), np.ones(64, dtype=np.float32))


@pytest.mark.parametrize("name", ["vidor_single", "stress"])
def test_stage_full_size_properties(name):
    spec = synth.CONFIGS[name]
    n, t, c, r, k = spec["n"][0], spec["t"][0], spec["classes"], spec["predicates"], spec["topk"]
    v = synth.make_video(n, t, c, seed=29)
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0)
    sizes, stride = (16.0, 64.0, 256.0, 1024.0), 16.0
    stage = PairStage(StageConfig(n_classes=c, n_predicates=r, topk=k, sparsify=True, precision="fp32",
                                  anchor_sizes=sizes, anchor_stride=stride))
    stage.load_weights(sd, "cuda")
    batch = HostBatch.from_videos([v]).to_device("cuda")
    res = stage.forward(batch)
    torch.cuda.synchronize()
    # relationness + top-K: bit-exact against the C oracle (fp32 exact order), sorted, duplicate-free
    sc = exact.relationness(v.cls, sd)
    np.testing.assert_array_equal(batch.score_view(res.scores, 0).cpu().numpy(), sc)
    sc_nodiag = sc.copy()
    sc_nodiag[np.arange(n), np.arange(n)] = -np.inf
    want = exact.topk(sc_nodiag, k)[:min(k, n * (n - 1))]
    got = res.pair_proposals(0).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    vals = res.topk_score[0, :len(got)].cpu().numpy()
    assert (np.diff(vals) <= 0).all() and len(set(got.tolist())) == len(got)
    # idempotence: selecting again from the selected scores returns them unchanged
    np.testing.assert_array_equal(np.sort(vals)[::-1], vals)
    # spans: ordered, inside the video
    sp = res.spans[0].cpu().numpy()
    assert sp.shape[0] == len(got) and (sp[..., 0] >= 0).all() and (sp[..., 1] <= t).all()
    assert (sp[..., 1] > sp[..., 0]).all()
    # predicate scores are probabilities; records sorted by score, ids inside the video
    lg = res.logits(0).cpu().numpy()
    assert lg.shape == (len(got), r) and (lg >= 0).all() and (lg <= 1).all() and np.isfinite(lg).all()
    rec = res.records[0].cpu().numpy()
    cnt = int(res.record_counts[0])
    assert cnt == min(200, len(got) * 20)
    scores = rec[:cnt, 0].view(np.float32)
    assert (np.diff(scores) <= 0).all()
    assert (rec[:cnt, 4] >= 0).all() and (rec[:cnt, 4] < n).all() and (rec[:cnt, 4] != rec[:cnt, 5]).all()
    assert (rec[:cnt, 2] >= 0).all() and (rec[:cnt, 2] < r).all()
    del res
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name,videos,check", [("vidvrd_test", 200, 6), ("vidor_val", 48, 3)])
def test_ragged_config_batches(name, videos, check):
    """configs[1] (all 200 VidVRD-test-shaped videos in one launch) and a 48-video slice of configs[3]:
    every video's reductions satisfy the symmetry property; `check` seeded videos are compared with the
    oracle (sampled pairs), proposals bit-exact."""
    spec = synth.CONFIGS[name]
    c, r, k = spec["classes"], spec["predicates"], spec["topk"]
    vids = synth.make_config(name, seed=3, videos=videos)
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0)
    stage = PairStage(StageConfig(n_classes=c, n_predicates=r, topk=k, sparsify=True, precision="fp32"))
    stage.load_weights(sd, "cuda")
    batch = HostBatch.from_videos(vids).to_device("cuda")
    res = stage.forward(batch)
    torch.cuda.synchronize()
    viou = res.geom["viou"].cpu().numpy()
    ov = res.geom["overlap"].cpu().numpy()
    for i, v in enumerate(vids):
        n = v.n_tracklets
        sl = batch.pair_slice(i)
        if n < 2:
            assert sl.stop == sl.start
            continue
        _, rev = _rows(n)
        np.testing.assert_array_equal(viou[sl], viou[sl][rev])
        np.testing.assert_array_equal(ov[sl], ov[sl][rev])
    rng = np.random.Generator(np.random.PCG64(1))
    for i in rng.choice(len(vids), size=check, replace=False):
        v = vids[int(i)]
        n, t = v.n_tracklets, v.n_frames
        if n < 2:
            continue
        pr, _ = _rows(n)
        sel = np.sort(rng.choice(pr.shape[0], size=min(32, pr.shape[0]), replace=False))
        w_geo, w_viou, _, w_ov = ogeo.pair_geometry(v.boxes, v.span, pr[sel, 0], pr[sel, 1])
        geo = batch.geo_rows(res.geom["geo"], int(i))
        got = geo[torch.from_numpy(sel).cuda()].cpu().numpy()[:, :, :t]
        np.testing.assert_allclose(got, w_geo, rtol=1e-5, atol=1e-12)
        sl = batch.pair_slice(int(i))
        np.testing.assert_allclose(viou[sl][sel], w_viou, rtol=1e-5, atol=0)
        np.testing.assert_array_equal(ov[sl][sel], w_ov)
        sc = exact.relationness(v.cls, sd)
        sc[np.arange(n), np.arange(n)] = -np.inf
        np.testing.assert_array_equal(res.pair_proposals(int(i)).cpu().numpy(), exact.topk(sc, k)[:min(k, n * (n - 1))])


@pytest.mark.parametrize("name,videos", [("vidor_single", 3), ("vidor_val", 24)])
def test_bench_path_full_size(name, videos, monkeypatch):
    """The bench path (tensor precision, sparsify: heads of the survivors recomputed from the boxes on the side
    branches, persistent pair kernel with reserved SMs, fused scores + top-K) at BASELINE.json sizes - configs[2]
    (N=64, T=2000) and a ragged slice of configs[3] - against the stored-geometry-row path of the same stage: span
    proposals, proposals and reductions bit for bit, predicate scores to the split-K summation order; and against
    the float64 oracle on sampled rows (1e-2 absolute, BASELINE.json)."""
    from oracle import features as ofeat, heads as oheads
    spec = synth.CONFIGS[name]
    c, r, k = spec["classes"], spec["predicates"], spec["topk"]
    vids = synth.make_config(name, seed=5, videos=videos)
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0)
    cfg = StageConfig(n_classes=c, n_predicates=r, topk=k, sparsify=True, precision="tensor",
                      anchor_sizes=(16.0, 64.0, 256.0, 1024.0), anchor_stride=16.0)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("TSPN_SURVIVOR_PATH", mode)
        monkeypatch.setenv("TSPN_FUSED_TOPK", mode)
        stage = PairStage(cfg)
        stage.load_weights(sd, "cuda")
        batch = HostBatch.from_videos(vids).to_device("cuda")
        g = stage.capture(batch)
        g.replay()
        res[mode] = g.replay()
        torch.cuda.synchronize()
    a, b = res["0"], res["1"]
    assert torch.equal(a.topk_idx, b.topk_idx) and torch.equal(a.topk_score, b.topk_score)
    assert torch.equal(a.scores, b.scores)
    for key in ("viou", "tiou", "overlap"):
        assert torch.equal(a.geom[key], b.geom[key]), key
    assert (a.rel_logits - b.rel_logits).abs().max().item() <= 5e-6
    assert torch.equal(a.record_counts, b.record_counts)
    for i in range(len(vids)):
        assert torch.equal(a.spans[i], b.spans[i]), i
    rng = np.random.Generator(np.random.PCG64(2))
    for i in rng.choice(len(vids), size=min(3, len(vids)), replace=False):
        v = vids[int(i)]
        n = v.n_tracklets
        if n < 2:
            continue
        order = b.pair_proposals(int(i)).cpu().numpy()
        pick = np.sort(rng.choice(len(order), size=min(8, len(order)), replace=False))
        s, o = order[pick] // n, order[pick] % n
        geo, _, _, ov = ogeo.pair_geometry(v.boxes, v.span, s, o)
        feats = ofeat.assemble_features(v.cls, v.motion, ofeat.relative_block(geo, ov), np.stack([s, o], axis=1))
        want = oheads.relation_predictor_f64(feats, sd)
        np.testing.assert_allclose(b.logits(int(i)).cpu().numpy()[pick], want, rtol=0, atol=1e-2)
    del res, a, b
    torch.cuda.empty_cache()
