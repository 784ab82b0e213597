"""GPU parity: the CUDA path (through the C ABI) against the oracle on identical seeded inputs.

Tolerances are BASELINE.json's: pair indices, top-K selection and span frame bounds bit-exact in
fp32 mode; features and vIoU within 1e-5 relative; tensor-core (bf16 / tf32) scores within 1e-2
absolute.
"""
import numpy as np
import pytest
import torch

from oracle import exact, features as ofeat, geometry as ogeo, heads as oheads
from tspn_b200 import _lib, ops, synth
from tspn_b200.batch import HostBatch

pytestmark = pytest.mark.gpu

RTOL = 1e-5          # "features and vIoU within 1e-5 relative"


def _batch(videos):
    return HostBatch.from_videos(videos).to_device("cuda")


def _geo_oracle(v, clip=False):
    return ogeo.pair_geometry_chunked(v.boxes, v.span, clip_volumes=clip)


RAGGED = [(20, 300, 0), (5, 37, 11), (2, 1, 3), (1, 9, 4), (0, 5, 5), (3, 3, 6), (7, 513, 7), (9, 512, 8),
          (4, 1030, 9), (12, 64, 2)]


@pytest.mark.parametrize("clip", [False, True])
def test_pair_geometry_ragged_batch(clip):
    vids = [synth.make_video(n, t, 35, seed=s) for n, t, s in RAGGED]
    batch = _batch(vids)
    out = ops.pair_geometry(batch, write_geo=True, clipped=clip)
    torch.cuda.synchronize()
    viou = out["viou"].cpu().numpy()
    tiou = out["tiou"].cpu().numpy()
    ov = out["overlap"].cpu().numpy()
    for i, v in enumerate(vids):
        sl = batch.pair_slice(i)
        if v.n_pairs == 0:
            assert sl.stop == sl.start
            continue
        geo, w_viou, w_tiou, w_ov = _geo_oracle(v, clip)
        got = batch.geo_rows(out["geo"], i).cpu().numpy()
        t = v.n_frames
        np.testing.assert_array_equal(got[:, :, t:], 0)                    # row pad written as zero
        np.testing.assert_allclose(got[:, :, :t], geo, rtol=RTOL, atol=1e-12, err_msg="video %d" % i)
        np.testing.assert_array_equal(got[:, 7, :t], geo[:, 7])            # overlap mask exact
        np.testing.assert_array_equal(ov[sl], w_ov)                         # frame bounds bit-exact
        np.testing.assert_allclose(viou[sl], w_viou, rtol=RTOL, atol=0)
        np.testing.assert_allclose(tiou[sl], w_tiou, rtol=1e-6, atol=0)


@pytest.mark.parametrize("shapes,chunk", [
    ([(20, 300, 0), (5, 37, 11), (9, 512, 8), (2, 1, 3)], 512),          # one 128-thread CTA per row segment
    ([(7, 513, 7), (4, 1024, 9), (3, 700, 1)], 1024),                    # 256 threads
    ([(3, 4100, 2), (70, 40, 3), (4, 2049, 4), (5, 2048, 5)], 2048),     # 512 threads; several chunks per row;
])                                                                       # N-1 > 64: two object groups
@pytest.mark.parametrize("clip", [False, True])
def test_pair_geometry_kernel_shapes(shapes, chunk, clip):
    """Every CTA shape tspn_geo_chunk selects, rows spanning several chunks (forward differences and
    volume sums across chunk boundaries), more objects than one work item holds."""
    vids = [synth.make_video(n, t, 35, seed=s) for n, t, s in shapes]
    batch = _batch(vids)
    assert int(batch.totals[_lib.TOT_GEO_CHUNK]) == chunk
    out = ops.pair_geometry(batch, write_geo=True, clipped=clip)
    red = ops.pair_geometry(batch, write_geo=False, clipped=clip)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out["viou"].cpu().numpy(), red["viou"].cpu().numpy())
    viou, tiou, ov = out["viou"].cpu().numpy(), out["tiou"].cpu().numpy(), out["overlap"].cpu().numpy()
    for i, v in enumerate(vids):
        sl = batch.pair_slice(i)
        geo, w_viou, w_tiou, w_ov = _geo_oracle(v, clip)
        got = batch.geo_rows(out["geo"], i).cpu().numpy()
        t = v.n_frames
        np.testing.assert_array_equal(got[:, :, t:], 0)
        np.testing.assert_allclose(got[:, :, :t], geo, rtol=RTOL, atol=1e-12, err_msg="video %d" % i)
        np.testing.assert_array_equal(ov[sl], w_ov)
        np.testing.assert_allclose(viou[sl], w_viou, rtol=RTOL, atol=0)
        np.testing.assert_allclose(tiou[sl], w_tiou, rtol=1e-6, atol=0)


@pytest.mark.parametrize("shapes", [
    RAGGED,                                                               # ragged batch incl. N = 0 / 1 videos
    [(20, 300, 0), (5, 37, 11), (9, 512, 8), (2, 1, 3)],                  # 128-thread CTAs
    [(7, 513, 7), (4, 1024, 9), (3, 700, 1)],                             # 256-thread CTAs
    [(3, 4100, 2), (70, 40, 3), (4, 2049, 4), (5, 2048, 5)],              # 512-thread CTAs, two object groups
    [(2, 9, 1)] * 700,                                                    # 1400 one-object items, each shorter than
                                                                          # the TMA ring
    [(3, 600, 4), (2, 2100, 5), (40, 2100, 6)],                           # items of 1, 2 and 39 objects interleaved
])
@pytest.mark.parametrize("clip", [False, True])
def test_dense_and_sparse_kernel_shapes_are_bit_identical(shapes, clip):
    """The two occupancy shapes of the pair-geometry kernel (1024 threads per SM / 2-stage ring / 64 registers
    against ~512 threads / 3-stage ring) and its two scheduling forms (persistent CTAs pulling work items from a
    queue - the ring phases run on across items - against one CTA per work item): every output bit for bit."""
    vids = [synth.make_video(n, t, 35, seed=s) for n, t, s in shapes]
    batch = _batch(vids)
    a = ops.pair_geometry(batch, write_geo=True, clipped=clip, dense_ctas=False, persistent=True)
    b = ops.pair_geometry(batch, write_geo=True, clipped=clip, dense_ctas=True, persistent=False)
    c = ops.pair_geometry(batch, write_geo=False, clipped=clip, dense_ctas=False, persistent=True)
    d = ops.pair_geometry(batch, write_geo=True, clipped=clip, dense_ctas=False, persistent=False)
    torch.cuda.synchronize()
    for key in ("geo", "viou", "tiou", "overlap"):
        assert torch.equal(a[key], b[key]), key
        assert torch.equal(a[key], d[key]), key
    for key in ("viou", "tiou", "overlap"):
        assert torch.equal(a[key], c[key]), key


@pytest.mark.parametrize("shapes", [
    [(20, 300, 0), (5, 37, 11), (9, 512, 8), (2, 1, 3), (1, 9, 4), (0, 5, 5)],      # single chunk (512)
    [(12, 2000, 1), (70, 1500, 2)],                                                # single chunk (2048), two groups
    [(3, 4100, 2), (4, 2049, 4), (5, 2048, 5)],                                    # several chunks per row
])
@pytest.mark.parametrize("clip", [False, True])
def test_pair_geometry_phases_on_two_streams_and_stale_sums(shapes, clip):
    """The three phases of tspn_pair_geo_viou issued the way the pipeline issues them - PRE on a side stream under
    MAIN (every (pair, chunk) sum has a single writer, which stores it: nothing is zeroed, also when a row spans
    several chunks), POST on the side stream after both - onto outputs that still hold another batch's sums and
    windows: bit-identical to the one-call form on fresh outputs."""
    vids = [synth.make_video(n, t, 35, seed=s) for n, t, s in shapes]
    batch = _batch(vids)
    want = ops.pair_geometry(batch, write_geo=True, clipped=clip)
    out = ops.pair_geometry_outputs(batch, write_geo=True)
    for key in ("viou", "tiou"):
        out[key].fill_(float("nan"))
    out["overlap"].fill_(-7)
    out["workspace"].view(torch.int64).fill_(0x0123456789abcdef)         # stale volumes and per-pair sums
    torch.cuda.synchronize()
    main, side = torch.cuda.current_stream(), torch.cuda.Stream()
    assert int(batch.totals[_lib.TOT_MAX_CHUNKS]) == max(-(-t // 2048) for _, t, _ in shapes)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        ops.pair_geometry_phase(batch, out, _lib.GEO_PHASE_PRE, clipped=clip)
    ops.pair_geometry_phase(batch, out, _lib.GEO_PHASE_MAIN, clipped=clip)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        ops.pair_geometry_phase(batch, out, _lib.GEO_PHASE_POST, clipped=clip)
    main.wait_stream(side)
    torch.cuda.synchronize()
    for key in ("geo", "viou", "tiou", "overlap"):
        assert torch.equal(want[key], out[key]), key


def test_pair_geometry_reductions_only_and_fractional_boxes():
    v = synth.make_video(11, 700, 35, seed=21, integer_boxes=False)
    batch = _batch([v])
    full = ops.pair_geometry(batch, write_geo=True)
    red = ops.pair_geometry(batch, write_geo=False)
    torch.cuda.synchronize()
    assert red["geo"] is None
    np.testing.assert_array_equal(full["viou"].cpu().numpy(), red["viou"].cpu().numpy())
    geo, w_viou, _, _ = _geo_oracle(v)
    got = batch.geo_view(full["geo"], 0).cpu().numpy()
    # fractional boxes: the inputs themselves carry fp32 rounding, so the difference channels are
    # compared against the channel's scale instead of element-wise
    for ch in range(8):
        scale = max(np.abs(geo[:, ch]).max(), 1e-6)
        assert np.abs(got[:, ch] - geo[:, ch]).max() <= 2e-5 * scale, ch
    np.testing.assert_allclose(red["viou"].cpu().numpy(), w_viou, rtol=RTOL)


@pytest.mark.parametrize("tag", ["A", "S"])
def test_cubic_iou_against_reference_outputs(golden, tag):
    n, t, seed = {"A": (20, 300, 0), "S": (5, 37, 11)}[tag]
    v = synth.make_video(n, t, 35, seed=seed, full_span=True)
    b = torch.from_numpy(v.boxes).cuda()
    got = ops.cubic_iou(b, b).cpu().numpy()
    np.testing.assert_allclose(got, ogeo.cubic_iou_f64(v.boxes, v.boxes), rtol=2e-7)
    np.testing.assert_allclose(got, golden[f"cubic_iou_f32_{tag}"], rtol=RTOL)        # the reference itself
    np.testing.assert_allclose(np.diag(got), 1.0, atol=1e-6)
    np.testing.assert_array_equal(got, got.T)
    cross = ops.cubic_iou(b[: n // 2].contiguous(), b[n // 2:].contiguous()).cpu().numpy()
    np.testing.assert_allclose(cross, golden[f"cubic_iou_cross_{tag}"], rtol=RTOL)
    # the all-pairs kernel reproduces the matrix's off-diagonal entries bit for bit
    batch = _batch([v])
    out = ops.pair_geometry(batch, write_geo=False)
    pr = ogeo.enumerate_pairs(n)
    np.testing.assert_array_equal(out["viou"].cpu().numpy(), got[pr[:, 0], pr[:, 1]])
    assert torch.all(out["tiou"] == 1.0)


def test_viou_v2_v3_against_reference_outputs(golden):
    from tspn_b200 import trajectory
    v = synth.make_video(20, 300, 35, seed=3, full_span=False)
    pr = ogeo.enumerate_pairs(20)
    trajs = [v.boxes[i, v.span[i, 0]:v.span[i, 1]] for i in range(20)]
    durs = [tuple(int(x) for x in v.span[i]) for i in range(20)]
    got2 = trajectory.viou_batch(trajs, durs, pr)
    np.testing.assert_allclose(got2, golden["viou_v2_seed3"], rtol=2e-7, atol=0)
    got3 = trajectory.viou_batch(trajs, durs, pr, clipped=True)
    ref3 = golden["traj_iou_v3_seed3"]
    ok = ~np.isnan(ref3)
    np.testing.assert_allclose(got3[ok], ref3[ok], rtol=RTOL, atol=0)
    # the reference's scalar signature
    s, o = pr[17]
    assert trajectory.viou(trajs[s].tolist(), durs[s], trajs[o].tolist(), durs[o]) == pytest.approx(
        golden["viou_v2_seed3"][17], rel=2e-7)
    batch = _batch([v])
    np.testing.assert_array_equal(ops.pair_geometry(batch, write_geo=False)["viou"].cpu().numpy(), got2)


def test_enumerate_pairs_bit_exact():
    vids = [synth.make_video(n, 4, 35, seed=n) for n in (0, 1, 2, 5, 20, 3)]
    batch = _batch(vids)
    got = ops.enumerate_pairs(batch).cpu().numpy()
    want = np.concatenate([ogeo.enumerate_pairs(v.n_tracklets) for v in vids], axis=0)
    np.testing.assert_array_equal(got, want)


def test_features_layout_and_values(golden):
    vids = [synth.make_video(n, t, 35, seed=s) for n, t, s in [(6, 300, 0), (5, 37, 11), (4, 1200, 5), (3, 3, 6)]]
    vids[0].motion[2, :1000] = 0           # an empty histogram stays zero
    batch = _batch(vids)
    geom = ops.pair_geometry(batch)
    mn = ops.normalize_motion(batch.motion)
    f32, bf = ops.assemble_features(batch, mn, geom["geo"], geom["overlap"], want_fp32=True, want_bf16=True)
    torch.cuda.synchronize()
    f32, bf = f32.cpu().numpy(), bf.float().cpu().numpy()
    mn_all = mn.cpu().numpy()
    for i, v in enumerate(vids):
        geo, _, _, ov = _geo_oracle(v)
        rel = ofeat.relative_block(geo, ov)
        pr = ogeo.enumerate_pairs(v.n_tracklets)
        want = ofeat.assemble_features(v.cls, v.motion, rel, pr)
        # a pooled bin is a mean of signed per-frame values: 1e-5 relative to the magnitude of what
        # is averaged (mean |x| over the bin), element-wise 1e-5 relative everywhere else
        scale = np.abs(ofeat.assemble_features(v.cls, v.motion, ofeat.relative_block(np.abs(geo), ov), pr))
        sl = batch.pair_slice(i)
        err = np.abs(f32[sl] - want)
        assert (err <= RTOL * scale + 1e-12).all(), ("video %d" % i, float((err / (scale + 1e-30)).max()))
        nrel = want.shape[1] - 3000
        np.testing.assert_allclose(f32[sl][:, :nrel], want[:, :nrel], rtol=RTOL, atol=0, err_msg="video %d" % i)
        assert (np.abs(bf[sl] - want) <= 2.0 ** -8 * scale + 1e-12).all()
        np.testing.assert_allclose(mn_all[batch.tracklet_slice(i)], ofeat.normalize_motion_ref(v.motion),
                                   rtol=1e-6, atol=0)
    assert f32.shape[1] == synth.feature_dim(35) == 11070
    # a row subset with a padding row (-1)
    rows = torch.tensor([3, -1, 0, int(batch.total_pairs) - 1], dtype=torch.int64, device="cuda")
    sub, _ = ops.assemble_features(batch, mn, geom["geo"], geom["overlap"], rows=rows)
    sub = sub.cpu().numpy()
    np.testing.assert_array_equal(sub[0], f32[3])
    np.testing.assert_array_equal(sub[1], 0)
    np.testing.assert_array_equal(sub[3], f32[-1])


def _ppn_keys(sd):
    return [torch.from_numpy(sd["relpn.pair_proposal_network.ppn_head." + k]).cuda() for k in ops.PPN_KEYS]


@pytest.mark.parametrize("k", [256, 16, 1024])
def test_relationness_and_topk_bit_exact(golden, k):
    shapes = [(20, 300, 0), (12, 8, 2), (1, 8, 3), (2, 8, 4), (0, 8, 5), (40, 8, 6)]
    vids = [synth.make_video(n, t, 35, seed=s) for n, t, s in shapes]
    sd = synth.make_weights(35, 132, 11070, seed=0)
    batch = _batch(vids)
    scores = ops.relationness(batch, _ppn_keys(sd))
    for excl in (False, True):
        idx, val, row = ops.topk_pairs(batch, scores, k, exclude_diagonal=excl)
        torch.cuda.synchronize()
        for i, v in enumerate(vids):
            n = v.n_tracklets
            want = exact.relationness(v.cls, sd) if n else np.zeros((0, 0), np.float32)
            got = batch.score_view(scores, i).cpu().numpy()
            np.testing.assert_array_equal(got, want)                           # scores bit-exact
            w = want.copy()
            if excl and n:
                w[np.arange(n), np.arange(n)] = -np.inf
            order = exact.topk(w, k)
            k_eff = min(k, n * n - (n if excl else 0))
            order = order[:k_eff]
            np.testing.assert_array_equal(idx[i, :k_eff].cpu().numpy(), order)  # selection + order bit-exact
            assert torch.all(idx[i, k_eff:] == -1)
            np.testing.assert_array_equal(val[i, :k_eff].cpu().numpy(), want.reshape(-1)[order])
            s, o = order // max(n, 1), order % max(n, 1)
            wrow = np.where(s == o, -1, batch.pair_slice(i).start + s * (n - 1) + o - (o > s))
            np.testing.assert_array_equal(row[i, :k_eff].cpu().numpy(), wrow)
    # the reference's own scores / proposals (golden: N=20 video of seed 0 is config A)
    np.testing.assert_allclose(batch.score_view(scores, 0).cpu().numpy(), golden["ppn_scores_A"], atol=2e-6)


def test_topk_ties_go_to_lower_index():
    n = 24
    v = synth.make_video(n, 4, 35, seed=1)
    batch = _batch([v])
    rng = np.random.Generator(np.random.PCG64(5))
    sc = rng.integers(0, 7, size=n * n).astype(np.float32) / 8.0            # heavy ties
    idx, val, _ = ops.topk_pairs(batch, torch.from_numpy(sc).cuda(), 100)
    np.testing.assert_array_equal(idx[0].cpu().numpy(), exact.topk(sc, 100))
    np.testing.assert_array_equal(idx[0].cpu().numpy(), oheads.topk_stable(sc, 100))


@pytest.mark.parametrize("tag", ["A", "V"])
def test_predicate_head_exact_and_reference(golden, tag):
    from tests.golden.make_golden import synth_features
    n, t, c, r, seed = {"A": (20, 300, 35, 132, 0), "V": (12, 64, 80, 50, 2)}[tag]
    fdim = synth.feature_dim(c)
    sd = synth.make_weights(c, r, fdim, seed=seed)
    feats = synth_features(n * (n - 1), fdim, seed)
    w = torch.from_numpy(sd["classifier.rel_predictor.weight"]).cuda()
    b = torch.from_numpy(sd["classifier.rel_predictor.bias"]).cuda() + 0.01
    x = torch.from_numpy(feats).cuda()
    got = ops.predicate_head(x, w, b, "fp32").cpu().numpy()
    sd2 = dict(sd)
    sd2["classifier.rel_predictor.bias"] = b.cpu().numpy()
    np.testing.assert_array_equal(got, exact.predicate(feats, sd2))           # bit-exact vs the fixed order
    got0 = ops.predicate_head(x, w, b - 0.01, "fp32").cpu().numpy()
    np.testing.assert_allclose(got0, golden[f"rel_logits_{tag}"], rtol=0, atol=1e-6)   # the reference
    # row-padded view (stride != F) gives the same bits
    ld = ops.padded(fdim, 4) + 4
    buf = torch.zeros((x.shape[0], ld), device="cuda")
    buf[:, :fdim] = x
    np.testing.assert_array_equal(ops.predicate_head(buf[:, :fdim], w, b, "fp32").cpu().numpy(), got)


@pytest.mark.parametrize("cin,k,t", [(8, 6, 300), (8, 3, 1), (5, 2, 7), (64, 3, 50), (20, 2, 131)])
def test_span_head_exact(golden, cin, k, t):
    sd = synth.make_weights(35, 132, 16, dpn_in=cin, n_anchors=4, seed=9 if cin == 64 else 0)
    rng = np.random.Generator(np.random.PCG64((9 if cin == 64 else 0) + 77))
    x = rng.normal(0, 1, size=(k, cin, t)).astype(np.float32)
    p = "relpn.duration_proposal_network.dpn_head."
    sd[p + "conv.bias"] = rng.normal(0, 0.01, size=cin).astype(np.float32)
    sd[p + "duration_pred.bias"] = rng.normal(0, 0.01, size=8).astype(np.float32)
    args = [torch.from_numpy(sd[p + n]).cuda() for n in ("conv.weight", "conv.bias", "duration_pred.weight",
                                                         "duration_pred.bias")]
    got = ops.span_head(torch.from_numpy(x).cuda(), *args).cpu().numpy()
    np.testing.assert_array_equal(got, exact.span_head(x, sd))                 # bit-exact
    np.testing.assert_allclose(got, oheads.dpn_head_f64(x, sd), rtol=0, atol=2e-6)
    if (cin, k, t) == (8, 6, 300):
        sd0 = synth.make_weights(35, 132, 11070, dpn_in=8, n_anchors=4, seed=0)
        args0 = [torch.from_numpy(sd0[p + n]).cuda() for n in ("conv.weight", "conv.bias", "duration_pred.weight",
                                                               "duration_pred.bias")]
        np.testing.assert_allclose(ops.span_head(torch.from_numpy(x).cuda(), *args0).cpu().numpy(),
                                   golden["dpn_reg_A"], rtol=0, atol=1e-6)    # the reference
    # gathered rows, including a padding row
    rows = torch.tensor([k - 1, -1, 0], dtype=torch.int64, device="cuda")
    sub = ops.span_head(torch.from_numpy(x).cuda(), *args, rows=rows).cpu().numpy()
    np.testing.assert_array_equal(sub[0], got[k - 1])
    np.testing.assert_array_equal(sub[1], 0)
    np.testing.assert_array_equal(sub[2], got[0])


@pytest.mark.parametrize("cin,a,k,t,stride", [(8, 4, 6, 300, 7.5), (8, 4, 3, 1, 7.5), (8, 4, 5, 2000, 16.0),
                                              (8, 4, 2, 37, 1.0), (5, 4, 2, 60, 7.5), (8, 3, 2, 131, 4.0),
                                              (20, 2, 2, 131, 7.5)])
def test_span_proposals_fused_bit_exact(cin, a, k, t, stride):
    """tspn_span_proposals == tspn_span_head(fp32) -> tspn_span_decode == the C oracle, bit for bit."""
    sd = synth.make_weights(35, 132, 16, dpn_in=cin, n_anchors=a, seed=3)
    rng = np.random.Generator(np.random.PCG64(cin * 1000 + t))
    x = rng.normal(0, 1, size=(k, cin, t)).astype(np.float32)
    p = "relpn.duration_proposal_network.dpn_head."
    sd[p + "conv.weight"] = (sd[p + "conv.weight"] * 30).astype(np.float32)       # regressions of O(1)
    sd[p + "duration_pred.weight"] = (sd[p + "duration_pred.weight"] * 30).astype(np.float32)
    sd[p + "conv.bias"] = rng.normal(0, 0.1, size=cin).astype(np.float32)
    sd[p + "duration_pred.bias"] = rng.normal(0, 0.5, size=2 * a).astype(np.float32)
    args = [torch.from_numpy(sd[p + n]).cuda() for n in ("conv.weight", "conv.bias", "duration_pred.weight",
                                                         "duration_pred.bias")]
    sizes = tuple(float(4 ** (i + 1)) for i in range(a))
    sz = torch.tensor(sizes, dtype=torch.float32).cuda()
    xd = torch.from_numpy(x).cuda()
    got = ops.span_proposals(xd, *args, sz, stride).cpu().numpy()
    two = ops.span_decode(ops.span_head(xd, *args), sz, stride).cpu().numpy()
    np.testing.assert_array_equal(got, two)
    np.testing.assert_array_equal(got, exact.span_decode(exact.span_head(x, sd), sizes, stride))
    if t > 1:
        assert len(np.unique(got[..., 1] - got[..., 0])) > 1                      # not a degenerate case
    # gathered rows of a row-padded buffer, including a padding row and a row base
    tp = (t + 3) // 4 * 4
    buf = torch.zeros((k, cin, tp), device="cuda")
    buf[:, :, :t] = xd
    rows = torch.tensor([k - 1 + 100, -1, 100], dtype=torch.int64, device="cuda")
    sub = ops.span_proposals(buf, *args, sz, stride, rows=rows, t=t, row_base=100).cpu().numpy()
    np.testing.assert_array_equal(sub[0], got[k - 1])
    np.testing.assert_array_equal(sub[2], got[0])
    zero = exact.span_decode(np.zeros((1, 2 * a, t), np.float32), sizes, stride)[0]
    np.testing.assert_array_equal(sub[1], zero)


@pytest.mark.parametrize("t,stride,sizes", [(60, 7.5, (15, 30, 45, 60)), (300, 7.5, (15, 30, 45, 60)),
                                            (37, 1.0, (4, 8, 16, 32)), (1, 7.5, (15, 30, 45, 60)),
                                            (2000, 16.0, (16, 64, 256, 1024))])
def test_span_decode_bit_exact(t, stride, sizes):
    rng = np.random.Generator(np.random.PCG64(t))
    reg = (rng.normal(0, 1.0, size=(5, 8, t)) * np.array([1, 2, 1, 2, 1, 2, 1, 6])[None, :, None]).astype(np.float32)
    got = ops.span_decode(torch.from_numpy(reg).cuda(), torch.tensor(sizes, dtype=torch.float32).cuda(),
                          stride).cpu().numpy()
    want = exact.span_decode(reg, sizes, stride)
    np.testing.assert_array_equal(got, want)                                    # frame bounds bit-exact
    assert got.shape[1] == ops.span_num_locations(t, stride) * 4 == exact.n_locations(t, stride) * 4
    assert (got[..., 0] >= 0).all() and (got[..., 1] <= t).all() and (got[..., 1] > got[..., 0]).all()
    f64 = oheads.decode_spans_f64(reg, sizes, stride)
    assert np.abs(got - f64).max() <= 1 and (got != f64).mean() < 0.01
    anc = oheads.grid_anchors(t, sizes, stride)
    assert anc.shape[0] == got.shape[1]


def test_c_abi_error_codes_on_the_device():
    """Error behaviour of the boundary on a real device: negative return code + thread-local message,
    nothing launched, no exception across the ABI (the Python wrapper raises RuntimeError)."""
    lib = _lib.load()
    v = synth.make_video(4, 40, 35, seed=1)
    batch = _batch([v])
    out = ops.pair_geometry(batch, write_geo=True)
    torch.cuda.synchronize()
    tot = batch.totals
    args = [batch.table.data_ptr(), 1, int(tot[_lib.TOT_ITEMS]), int(tot[_lib.TOT_GEO_CHUNK]),
            int(tot[_lib.TOT_MAX_CHUNKS]), batch.total_tracklets,
            batch.total_pairs, int(tot[_lib.TOT_BOXES]), batch.boxes.data_ptr(), batch.span.data_ptr(),
            out["geo"].data_ptr(), out["viou"].data_ptr(), out["tiou"].data_ptr(), out["overlap"].data_ptr(), 0,
            out["workspace"].data_ptr(), _lib.stream_ptr()]

    def call(**patch):
        a = list(args)
        for i, val in patch.items():
            a[int(i[1:])] = val
        return lib.tspn_pair_geo_viou(*a)
    assert call() == _lib.TSPN_OK
    assert call(_8=None) == _lib.TSPN_EBADARG and "null pointer" in _lib.last_error()
    assert call(_8=batch.boxes.data_ptr() + 4) == _lib.TSPN_EALIGN
    assert call(_7=int(tot[_lib.TOT_BOXES]) + 1) == _lib.TSPN_ESHAPE
    assert call(_3=777) == _lib.TSPN_EBADARG and "geo_chunk" in _lib.last_error()
    assert call(_4=0) == _lib.TSPN_EBADARG and "max_chunks" in _lib.last_error()
    assert call(_2=-1) == _lib.TSPN_EBADARG
    assert call(_2=0) == _lib.TSPN_OK                                   # empty batch: nothing to do
    with pytest.raises(RuntimeError, match="TSPN_EBADARG"):
        _lib.check(call(_9=None), "tspn_pair_geo_viou")
    # top-K: k beyond the supported block size, predicate head: tensor precision without packed weights
    scores = torch.rand(16, device="cuda")
    idx = torch.empty((1, 2000), dtype=torch.int64, device="cuda")
    val = torch.empty((1, 2000), dtype=torch.float32, device="cuda")
    assert lib.tspn_topk_pairs(batch.table.data_ptr(), 1, scores.data_ptr(), 2000, 0, idx.data_ptr(), val.data_ptr(),
                               None, _lib.stream_ptr()) < 0
    torch.cuda.synchronize()                                            # no sticky CUDA error was left behind
    again = ops.pair_geometry(batch, write_geo=True)
    torch.cuda.synchronize()
    assert torch.equal(again["geo"], out["geo"])


@pytest.mark.parametrize("exclude", [False, True])
@pytest.mark.parametrize("k", [256, 7, 1024])
def test_fused_scores_topk_equals_the_two_calls(exclude, k):
    """tspn_relationness_topk (one CTA per video computes, writes and ranks the video's scores) against
    tspn_relationness followed by tspn_topk_pairs: scores, indices, top scores and pair rows bit for bit, on a
    ragged batch with empty, one-tracklet and 90-tracklet (8100 candidates: the cache limit) videos."""
    c = 35
    shapes = [(20, 30, 0), (0, 5, 1), (1, 9, 2), (2, 4, 3), (64, 16, 4), (90, 8, 5), (33, 12, 6)]
    vids = [synth.make_video(n, t, c, seed=s) for n, t, s in shapes]
    sd = synth.make_weights(c, 50, synth.feature_dim(c), dpn_in=8, seed=3)
    w = [torch.from_numpy(sd["relpn.pair_proposal_network.ppn_head." + key]).cuda() for key in ops.PPN_KEYS]
    batch = _batch(vids)
    assert ops.relationness_topk_supported(batch)
    scores = ops.relationness(batch, w)
    idx, val, row = ops.topk_pairs(batch, scores, k, exclude_diagonal=exclude)
    f_scores, f_idx, f_val, f_row = ops.relationness_topk(batch, w, k, exclude_diagonal=exclude)
    torch.cuda.synchronize()
    assert torch.equal(scores, f_scores)
    assert torch.equal(idx, f_idx) and torch.equal(val, f_val) and torch.equal(row, f_row)
    big = _batch([synth.make_video(91, 4, c, seed=7)])
    assert not ops.relationness_topk_supported(big)


def test_pair_geometry_with_reserved_sms_is_bit_identical():
    """The persistent pair kernel with SM slots left free for concurrent streams (TSPN_GEO_RESERVE_SHIFT) - fewer
    CTAs pulling from the same work-item queue - against the full grid."""
    vids = [synth.make_video(n, t, 35, seed=s) for n, t, s in [(40, 700, 1), (9, 2100, 2), (3, 64, 3)]]
    batch = _batch(vids)
    want = ops.pair_geometry(batch, write_geo=True)
    for reserve in (8, 140, 255):
        out = ops.pair_geometry_outputs(batch, write_geo=True)
        ops.pair_geometry_phase(batch, out, 0, reserve_sms=reserve)
        torch.cuda.synchronize()
        for key in ("geo", "viou", "tiou", "overlap"):
            assert torch.equal(want[key], out[key]), (reserve, key)
