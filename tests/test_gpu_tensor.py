"""Tensor-core (tcgen05) heads against the float64 oracle: "bf16 MLP scores within 1e-2 absolute"."""
import numpy as np
import pytest
import torch

from oracle import heads as oheads
from tspn_b200 import ops, synth
from tests.golden.make_golden import synth_features

pytestmark = pytest.mark.gpu

ATOL_BF16 = 1e-2      # BASELINE.json: bf16 MLP scores within 1e-2 absolute


@pytest.mark.parametrize("m,c,r,seed", [(380, 35, 132, 0), (132, 80, 50, 2), (4032, 80, 50, 3), (1, 80, 50, 4),
                                        (129, 35, 132, 5)])
@pytest.mark.parametrize("dtype", ["fp32_tf32", "bf16"])
def test_predicate_head_tensor(m, c, r, seed, dtype):
    fdim = synth.feature_dim(c)
    sd = synth.make_weights(c, r, fdim, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed))
    sd["classifier.rel_predictor.bias"] = rng.normal(0, 0.5, size=r).astype(np.float32)
    sd["classifier.rel_predictor.weight"] = (sd["classifier.rel_predictor.weight"] * 5).astype(np.float32)
    feats = synth_features(m, fdim, seed)
    want = oheads.relation_predictor_f64(feats, sd)
    w = torch.from_numpy(sd["classifier.rel_predictor.weight"]).cuda()
    b = torch.from_numpy(sd["classifier.rel_predictor.bias"]).cuda()
    ld = ops.padded(fdim, 8)
    if dtype == "bf16":
        buf = torch.zeros((m, ld), dtype=torch.bfloat16, device="cuda")
        buf[:, :fdim] = torch.from_numpy(feats).cuda().to(torch.bfloat16)
    else:
        buf = torch.zeros((m, ld), dtype=torch.float32, device="cuda")
        buf[:, :fdim] = torch.from_numpy(feats).cuda()
    got = ops.predicate_head(buf[:, :fdim], w, b, "tensor").cpu().numpy()
    err = np.abs(got - want).max()
    assert err <= ATOL_BF16, err
    assert want.std() > 0.05                       # the scores are not trivially 0.5
    # tighter, informative bound: tf32 on fp32 storage is ~10x closer than bf16
    assert err <= (2e-3 if dtype == "fp32_tf32" else ATOL_BF16), err
    exact_fp32 = ops.predicate_head(torch.from_numpy(feats).cuda(), w, b, "fp32").cpu().numpy()
    assert np.abs(got - exact_fp32).max() <= ATOL_BF16


@pytest.mark.parametrize("cin,k,t,a", [(64, 3, 50, 4), (128, 5, 300, 4), (320, 2, 131, 4), (1024, 4, 300, 4),
                                       (72, 3, 9, 2), (256, 2, 1, 4), (512, 7, 257, 8)])
def test_span_head_tensor(cin, k, t, a, monkeypatch):
    """DPNHead (dpn.py:55-73) as a tcgen05 implicit GEMM: bf16 operands, fp32 accumulation.  Channel
    counts cover one chunk, several chunks, a ragged last chunk (320 = 256 + 64) and a ragged K slab
    (72); row tiles straddle pairs for every T here.  Cin >= 256 also runs the opt-in CTA-pair form
    (cta_group::2, ``TSPN_SPAN_HEAD_PAIR=1``): the same bits as the one-CTA form."""
    monkeypatch.delenv("TSPN_SPAN_HEAD_PAIR", raising=False)
    sd = synth.make_weights(35, 132, 16, dpn_in=cin, n_anchors=a, seed=cin + t)
    rng = np.random.Generator(np.random.PCG64(cin * 1000 + t))
    p = "relpn.duration_proposal_network.dpn_head."
    # scale the N(0, 0.01) init so that the hidden units and outputs are O(1): a meaningful 1e-2 test
    sd[p + "conv.weight"] = (sd[p + "conv.weight"] * (100.0 / np.sqrt(3 * cin))).astype(np.float32)
    sd[p + "duration_pred.weight"] = (sd[p + "duration_pred.weight"] * (100.0 / np.sqrt(cin))).astype(np.float32)
    sd[p + "conv.bias"] = rng.normal(0, 0.3, size=cin).astype(np.float32)
    sd[p + "duration_pred.bias"] = rng.normal(0, 0.3, size=2 * a).astype(np.float32)
    x = rng.normal(0, 1, size=(k, cin, t)).astype(np.float32)
    want = oheads.dpn_head_f64(x, sd)
    assert want.std() > 0.3
    args = [torch.from_numpy(sd[p + n]).cuda() for n in ("conv.weight", "conv.bias", "duration_pred.weight",
                                                         "duration_pred.bias")]
    xd = torch.from_numpy(x).cuda()
    got = ops.span_head(xd, *args, precision="tensor").cpu().numpy()
    assert got.shape == want.shape
    err = np.abs(got - want).max()
    assert err <= ATOL_BF16, err
    # gathered rows incl. a padding row, from a row-padded buffer (ld_t > t)
    buf = torch.zeros((k, cin, t + 3), dtype=torch.float32, device="cuda")
    buf[:, :, :t] = xd
    rows = torch.tensor([k - 1, -1, 0], dtype=torch.int64, device="cuda")
    sub = ops.span_head(buf, *args, rows=rows, t=t, precision="tensor").cpu().numpy()
    np.testing.assert_array_equal(sub[0], got[k - 1])
    np.testing.assert_array_equal(sub[1], 0)
    np.testing.assert_array_equal(sub[2], got[0])
    if cin >= 256:
        monkeypatch.setenv("TSPN_SPAN_HEAD_PAIR", "1")
        pair = ops.span_head(xd, *args, precision="tensor").cpu().numpy()
        np.testing.assert_array_equal(pair, got)
        np.testing.assert_array_equal(ops.span_head(buf, *args, rows=rows, t=t, precision="tensor").cpu().numpy(), sub)


def test_decomposed_predicate_head_pieces():
    """tspn_tracklet_rows, tspn_predicate_head_affine (raw) and tspn_assemble_relative against their float64
    definitions: x W^T = A_s[s] + A_o[o] + rel W_rel^T."""
    from oracle import features as ofeat, geometry as ogeo
    from tspn_b200.batch import HostBatch
    c, r = 35, 132
    vids = [synth.make_video(7, 90, c, seed=3), synth.make_video(4, 33, c, seed=4)]
    vids[0].motion[2, :1000] = 0
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=5)
    w = sd["classifier.rel_predictor.weight"].astype(np.float64)
    for compact in (True, False):
        batch = HostBatch.from_videos(vids, compact=compact).to_device("cuda")
        rows_t = ops.tracklet_rows(batch)
        torch.cuda.synchronize()
        cls = np.concatenate([v.cls for v in vids])
        mn = np.concatenate([np.concatenate([ofeat.l1_normalize_ref(v.motion[:, k * 1000:(k + 1) * 1000]) for k in range(4)], axis=1)
                             for v in vids])
        want_rows = np.concatenate([cls, mn], axis=1)
        got_rows = rows_t.float().cpu().numpy()
        assert got_rows.shape[1] % 8 == 0 and (got_rows[:, c + 4000:] == 0).all()
        np.testing.assert_allclose(got_rows[:, :c + 4000], want_rows, rtol=2 ** -8, atol=1e-30)      # bf16 rounding
    wt = torch.from_numpy(sd["classifier.rel_predictor.weight"]).cuda()
    w_s = torch.cat([wt[:, :c], wt[:, 2 * c:2 * c + 4000]], dim=1)
    w_o = torch.cat([wt[:, c:2 * c], wt[:, 2 * c + 4000:2 * c + 8000]], dim=1)
    terms = torch.cat([ops.predicate_head_affine(rows_t, ops.pack_predicate_weights(ws_.contiguous()), r, raw=True,
                                                 k_dim=c + 4000) for ws_ in (w_s, w_o)], dim=1)
    torch.cuda.synchronize()
    want_terms = np.concatenate([want_rows @ np.concatenate([w[:, :c], w[:, 2 * c:2 * c + 4000]], axis=1).T,
                                 want_rows @ np.concatenate([w[:, c:2 * c], w[:, 2 * c + 4000:2 * c + 8000]], axis=1).T], axis=1)
    np.testing.assert_allclose(terms.cpu().numpy(), want_terms, rtol=0, atol=2e-3)
    geom = ops.pair_geometry(batch, write_geo=True)
    n_pairs = batch.total_pairs
    sel = torch.tensor([0, 5, -1, n_pairs - 1, 17, -1, 41], dtype=torch.int64, device="cuda")
    rel16, row_bias = ops.assemble_relative(batch, geom["geo"], geom["overlap"], sel, terms[:, :r].contiguous(),
                                            terms[:, r:].contiguous())
    torch.cuda.synchronize()
    got_rel, got_rb, tm = rel16.float().cpu().numpy(), row_bias.cpu().numpy(), terms.cpu().numpy()
    for i, gp in enumerate(sel.tolist()):
        if gp < 0:
            assert (got_rel[i] == 0).all() and (got_rb[i] == 0).all()
            continue
        vi = 0 if gp < vids[0].n_pairs else 1
        v = vids[vi]
        p = gp - (0 if vi == 0 else vids[0].n_pairs)
        pr = ogeo.enumerate_pairs(v.n_tracklets)[p]
        g, _, _, ov = ogeo.pair_geometry(v.boxes, v.span, pr[None, 0], pr[None, 1])
        want_rel = ofeat.relative_block(g, ov)[0]
        scale = np.abs(want_rel).mean() + 1e-12
        assert np.abs(got_rel[i] - want_rel).max() <= 2 ** -7 * max(np.abs(want_rel).max(), scale)
        t0 = 0 if vi == 0 else vids[0].n_tracklets
        np.testing.assert_array_equal(got_rb[i], tm[t0 + pr[0], :r] + tm[t0 + pr[1], r:])


@pytest.mark.parametrize("sparsify", [True, False])
def test_decomposed_head_matches_materialised_rows_and_oracle(sparsify):
    """Whole stage, tensor precision: the decomposed classifier against the one-GEMM-over-F form
    (StageConfig.materialize_features) and against the float64 oracle (1e-2 absolute, BASELINE.json)."""
    from oracle import features as ofeat, geometry as ogeo, heads as oheads
    from tspn_b200.batch import HostBatch
    from tspn_b200.pipeline import PairStage, StageConfig
    c, r, k = 35, 132, 40
    vids = [synth.make_video(9, 120, c, seed=11), synth.make_video(3, 64, c, seed=12), synth.make_video(6, 200, c, seed=13)]
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=2)
    res = {}
    for mat in (False, True):
        stage = PairStage(StageConfig(n_classes=c, n_predicates=r, topk=k, sparsify=sparsify, precision="tensor",
                                      materialize_features=mat))
        stage.load_weights(sd, "cuda")
        batch = HostBatch.from_videos(vids).to_device("cuda")
        res[mat] = stage.forward(batch)
        torch.cuda.synchronize()
        assert (res[mat].features_bf16 is not None) == mat
    assert torch.equal(res[False].topk_idx, res[True].topk_idx)
    for i, v in enumerate(vids):
        a, b = res[False].logits(i).cpu().numpy(), res[True].logits(i).cpu().numpy()
        assert a.shape == b.shape and np.abs(a - b).max() <= 5e-3
        n = v.n_tracklets
        geo, _, _, ov = ogeo.pair_geometry_chunked(v.boxes, v.span)
        feats = ofeat.assemble_features(v.cls, v.motion, ofeat.relative_block(geo, ov), ogeo.enumerate_pairs(n))
        if sparsify:
            order = res[False].pair_proposals(i).cpu().numpy()
            s, o = order // n, order % n
            feats = feats[s * (n - 1) + o - (o > s)]
        want = oheads.relation_predictor_f64(feats, sd)
        np.testing.assert_allclose(a, want, rtol=0, atol=1e-2)
    # records are built from the decomposed logits
    assert res[False].records.shape == res[True].records.shape


@pytest.mark.parametrize("shapes,k", [
    ([(9, 120, 11), (3, 64, 12), (6, 700, 13), (1, 30, 14), (2, 5, 15)], 24),     # ragged: padding rows, N = 1, tiny T
    ([(12, 2000, 1), (40, 1500, 2)], 96),                                          # bench-shaped rows (4 passes)
    ([(5, 2500, 3), (4, 513, 4)], 20),                                             # tile longer than 2048 frames
])
def test_survivor_rows_bit_identical_to_stored_rows(shapes, k):
    """tspn_survivor_rows (relative block, bias rows and span proposals recomputed from the boxes) against
    tspn_assemble_relative + tspn_span_proposals on the geometry rows the all-pairs kernel stored: every output
    bit for bit, including padding rows and videos shorter than the batch's longest."""
    from tspn_b200.batch import HostBatch
    c, r = 35, 50
    vids = [synth.make_video(n, t, c, seed=s) for n, t, s in shapes]
    vids[0].span[1] = (0, 3)                                   # some pairs without temporal overlap
    vids[0].span[2] = (vids[0].n_frames - 2, vids[0].n_frames)
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=5)
    batch = HostBatch.from_videos(vids).to_device("cuda")
    rng = np.random.Generator(np.random.PCG64(k))
    rows = np.full((len(vids), k), -1, dtype=np.int64)
    off = 0
    for i, v in enumerate(vids):
        kk = min(k, v.n_pairs)
        rows[i, :kk] = off + rng.permutation(v.n_pairs)[:kk]
        off += v.n_pairs
    rows_d = torch.from_numpy(rows).cuda()
    n_trk = batch.total_tracklets
    terms = torch.from_numpy(rng.normal(size=(2, n_trk, r)).astype(np.float32)).cuda()
    sw = tuple(torch.from_numpy(sd["relpn.duration_proposal_network.dpn_head." + key]).cuda()
               for key in ("conv.weight", "conv.bias", "duration_pred.weight", "duration_pred.bias"))
    sizes = torch.tensor([16.0, 64.0, 256.0, 1024.0], device="cuda")
    stride = 16.0
    assert ops.survivor_rows_supported(batch, 4)
    rel, bias, spans = ops.survivor_rows(batch, rows_d, terms[0], terms[1], span_weights=sw, sizes=sizes, stride=stride)
    geom = ops.pair_geometry(batch, write_geo=True)
    want_rel, want_bias = ops.assemble_relative(batch, geom["geo"], geom["overlap"], rows_d.reshape(-1), terms[0], terms[1])
    torch.cuda.synchronize()
    assert torch.equal(rel.view(torch.int16), want_rel.view(torch.int16))
    assert torch.equal(bias, want_bias)
    rel2, none, spans2 = ops.survivor_rows(batch, rows_d, span_weights=sw, sizes=sizes, stride=stride)   # no terms
    bias2 = ops.gather_pair_terms(batch, rows_d.reshape(-1), terms[0], terms[1])
    torch.cuda.synchronize()
    assert none is None and torch.equal(rel2.view(torch.int16), rel.view(torch.int16)) and torch.equal(spans2, spans)
    assert torch.equal(bias2, want_bias)
    a_n = 4
    l_max = ops.span_num_locations(max(batch.t), stride)
    assert spans.shape == (rows.size, l_max * a_n, 2)
    for i, v in enumerate(vids):
        t, tp = v.n_frames, (v.n_frames + 3) // 4 * 4
        l_v = ops.span_num_locations(t, stride)
        got = spans[i * k:(i + 1) * k]
        assert (got[:, l_v * a_n:] == 0).all()
        if v.n_pairs == 0:
            x = torch.zeros((1, 8, tp), device="cuda")
            want = ops.span_proposals(x, *sw, sizes, stride, rows=torch.full((k,), -1, dtype=torch.int64, device="cuda"),
                                      t=t, row_base=0)
        else:
            p0 = batch.pair_slice(i).start
            x = batch.geo_rows(geom["geo"], i)
            want = ops.span_proposals(x, *sw, sizes, stride, rows=rows_d[i], t=t, row_base=p0)
        torch.cuda.synchronize()
        assert torch.equal(got[:, :l_v * a_n], want), "video %d" % i


@pytest.mark.parametrize("use_dpn", [True, False])
def test_stage_survivor_path_equals_stored_row_path(use_dpn, monkeypatch):
    """The whole stage (tensor precision, sparsify) with the heads of the survivors on the side branch
    (tspn_survivor_rows -> affine head -> records, nothing waits for the pair kernel) against the path that
    reads the stored geometry rows after it: logits to split-K summation order, everything else bit for bit;
    eager and captured."""
    from tspn_b200.batch import HostBatch
    from tspn_b200.pipeline import PairStage, StageConfig
    c, r, k = 35, 50, 40
    vids = [synth.make_video(9, 120, c, seed=11), synth.make_video(3, 64, c, seed=12), synth.make_video(6, 700, c, seed=13),
            synth.make_video(1, 40, c, seed=14)]
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=2)
    cfg = StageConfig(n_classes=c, n_predicates=r, topk=k, sparsify=True, precision="tensor", use_dpn=use_dpn,
                      anchor_sizes=(16.0, 64.0, 256.0, 1024.0), anchor_stride=16.0)
    out = {}
    for mode in ("0", "1", "graph"):
        monkeypatch.setenv("TSPN_SURVIVOR_PATH", "0" if mode == "0" else "1")
        stage = PairStage(cfg)
        stage.load_weights(sd, "cuda")
        batch = HostBatch.from_videos(vids).to_device("cuda")
        assert stage._survivor_path(batch, None, True) == (mode != "0")
        if mode == "graph":
            g = stage.capture(batch)
            g.replay()
            res = g.replay()
        else:
            res = stage.forward(batch)
        torch.cuda.synchronize()
        out[mode] = res
    a = out["0"]
    for mode in ("1", "graph"):
        b = out[mode]
        assert torch.equal(a.topk_idx, b.topk_idx) and torch.equal(a.topk_row, b.topk_row)
        for key in ("viou", "tiou", "overlap", "geo"):
            assert torch.equal(a.geom[key], b.geom[key]), key
        assert (a.rel_logits - b.rel_logits).abs().max().item() <= 5e-6
        for i in range(len(vids)):
            if use_dpn:
                assert torch.equal(a.spans[i], b.spans[i]), i
            else:
                assert b.spans is None
        # records: same triplets unless two scores are closer than the summation-order noise
        sa, sb = ops.record_scores(a.records), ops.record_scores(b.records)
        assert torch.equal(a.record_counts, b.record_counts)
        assert (sa - sb).abs().max().item() <= 5e-6
        same = (a.records[..., 1:] == b.records[..., 1:]).all(dim=-1)
        # ... and every slot that differs must sit in a near-tie: a neighbouring rank of the same video scores within
        # the summation-order noise of it (then the two orders are both valid sorts of their own scores)
        gap = torch.full_like(sa, float("inf"))
        gap[:, 1:] = torch.minimum(gap[:, 1:], (sa[:, 1:] - sa[:, :-1]).abs())
        gap[:, :-1] = torch.minimum(gap[:, :-1], (sa[:, 1:] - sa[:, :-1]).abs())
        last = torch.zeros_like(same)                   # the cut at rank topk_per_video is a near-tie with rank + 1,
        cnt = a.record_counts.long()                    # which is not in the records: the last kept slot is exempt
        vid = torch.nonzero(cnt > 0).flatten()
        last[vid, cnt[vid] - 1] = True
        assert bool((same | (gap <= 1e-5) | last).all()), "records differ outside near-ties"
        assert same.float().mean().item() > 0.9
