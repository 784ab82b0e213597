"""Drop-in behaviour of the module mirrors against outputs of the unmodified reference."""
import numpy as np
import pytest
import torch

import tspn_b200
from oracle import exact, features as ofeat, geometry as ogeo, heads as oheads
from tspn_b200 import synth
from tspn_b200.list_pair import PairList
from tspn_b200.model import BaseModel, RelationPredictor
from tspn_b200.relpn import DPNHead, PPNHead
from tests.golden.make_golden import synth_features

pytestmark = pytest.mark.gpu

CASES = {"A": (20, 300, 35, 132, 0), "V": (12, 64, 80, 50, 2)}


def _cfg(c, r, fdim, use_ppn, use_dpn, **kw):
    cfg = tspn_b200.get_default_cfg()
    cfg.PREDICT.OBJECT_NUM, cfg.PREDICT.PREDICATE_NUM, cfg.PREDICT.FEATURE_DIM = c, r, fdim
    cfg.RELPN.PPN.IN_CHANNELS = cfg.RELPN.PPN.OUT_CHANNELS = c
    cfg.RELPN.USE_PPN, cfg.RELPN.USE_DPN = use_ppn, use_dpn
    cfg.RELPN.DPN.IN_CHANNELS = 8
    cfg.RELPN.DPN.NUM_DURATION_PROPOSALS = kw.pop("NUM_SPANS", 0)       # 0: every decoded span, no suppression
    for k, v in kw.items():
        cfg.PREDICT[k] = v
    return cfg


@pytest.mark.parametrize("tag", ["A", "V"])
def test_basemodel_reference_mode_cpu_in_cpu_out(golden, tag):
    """predict.py:57: model(pair_list, _) with CPU tensors and precomputed feature rows."""
    n, t, c, r, seed = CASES[tag]
    fdim = synth.feature_dim(c)
    sd = {k: torch.from_numpy(v) for k, v in synth.make_weights(c, r, fdim, seed=seed).items()}
    vid = synth.make_video(n, t, c, seed=seed)
    plist = PairList(torch.from_numpy(synth_features(n * (n - 1), fdim, seed)))
    plist.add_field("tracklet_pairs", torch.from_numpy(ogeo.enumerate_pairs(n)))
    plist.add_field("track_cls_logits", torch.from_numpy(vid.cls))
    plist.add_field("num_tracklets", n)
    for use_ppn in (False, True):
        model = BaseModel(_cfg(c, r, fdim, use_ppn, False))
        model.load_state_dict(sd)                     # the reference's 14 keys, strict
        model.eval()
        with torch.no_grad():
            pp, dp, logits = model([plist], None)
        assert dp is None and len(logits) == 1 and not logits[0].is_cuda
        np.testing.assert_allclose(logits[0].numpy(), golden[f"basemodel_{tag}_ppn{int(use_ppn)}_logits"],
                                   rtol=0, atol=1e-6)
        if use_ppn:
            ref = golden[f"basemodel_{tag}_ppn1_proposals"]
            assert pp[0].dtype == torch.int64 and pp[0].shape == ref.shape
            ex = exact.topk(exact.relationness(vid.cls, {k: v.numpy() for k, v in sd.items()}), 256)
            np.testing.assert_array_equal(pp[0].numpy(), ex)              # bit-exact vs the fixed order
            srt = np.sort(golden[f"ppn_scores_{tag}"].reshape(-1).astype(np.float64))[::-1]
            if (srt[:-1] - srt[1:])[:len(ref)].min() > 4e-6:              # margins >> score noise
                np.testing.assert_array_equal(pp[0].numpy(), ref)        # == the reference's selection
        else:
            assert pp is None


def test_heads_module_api(golden):
    n, t, c, r, seed = CASES["A"]
    fdim = synth.feature_dim(c)
    sd = {k: torch.from_numpy(v) for k, v in synth.make_weights(c, r, fdim, seed=seed).items()}
    vid = synth.make_video(n, t, c, seed=seed)
    head = PPNHead(c, 64, c).eval()
    head.load_state_dict({k.split("ppn_head.")[1]: v for k, v in sd.items() if "ppn_head" in k})
    cl = torch.from_numpy(vid.cls)
    m = head(cl, cl)
    np.testing.assert_allclose(m.numpy(), golden["ppn_scores_A"], atol=2e-6)
    m2 = head(cl[:7].clone(), cl[9:].clone())                              # different subject / object sets
    np.testing.assert_array_equal(m2.numpy(), m.numpy()[:7, 9:])
    clf = RelationPredictor(fdim, r).eval()
    clf.load_state_dict({k.split("classifier.")[1]: v for k, v in sd.items() if k.startswith("classifier.")})
    y = clf(torch.from_numpy(synth_features(n * (n - 1), fdim, seed)).cuda())
    assert y.is_cuda
    np.testing.assert_allclose(y.cpu().numpy(), golden["rel_logits_A"], atol=1e-6)
    dpn = DPNHead(8, 4).eval()
    dpn.load_state_dict({k.split("dpn_head.")[1]: v for k, v in sd.items() if "dpn_head" in k})
    x = np.random.Generator(np.random.PCG64(seed + 77)).normal(0, 1, size=(6, 8, t)).astype(np.float32)
    np.testing.assert_allclose(dpn(torch.from_numpy(x)).numpy(), golden["dpn_reg_A"], atol=1e-6)


@pytest.mark.parametrize("sparsify", [False, True])
def test_basemodel_full_pair_stage(sparsify):
    """Tracklets in, everything constructed on the GPU; two videos of different shape per call."""
    c, r = 35, 132
    fdim = synth.feature_dim(c)
    sd_np = synth.make_weights(c, r, fdim, dpn_in=8, seed=1)
    sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
    vids = [synth.make_video(9, 120, c, seed=4), synth.make_video(20, 300, c, seed=5)]
    pls = [PairList.from_tracklets(v.boxes, v.span, v.cls, v.motion) for v in vids]
    model = BaseModel(_cfg(c, r, fdim, True, True, SPARSIFY=sparsify)).eval()
    model.load_state_dict(sd)
    with torch.no_grad():
        pp, dp, logits = model(pls)
    sizes, stride = model.stage_config.anchor_sizes, model.stage_config.anchor_stride
    for i, v in enumerate(vids):
        n = v.n_tracklets
        geo, viou, _, ov = ogeo.pair_geometry_chunked(v.boxes, v.span)
        feats = ofeat.assemble_features(v.cls, v.motion, ofeat.relative_block(geo, ov), ogeo.enumerate_pairs(n))
        sc = exact.relationness(v.cls, sd_np)
        if sparsify:
            sc = sc.copy()
            sc[np.arange(n), np.arange(n)] = -np.inf
        k_eff = min(256, n * (n - 1) if sparsify else n * n)
        order = exact.topk(sc, 256)[:k_eff]
        np.testing.assert_array_equal(pp[i].numpy(), order)                        # pair indices bit-exact
        s, o = order // n, order % n
        rows = s * (n - 1) + o - (o > s)
        want_logits = oheads.relation_predictor_f64(feats, sd_np)
        if sparsify:
            np.testing.assert_allclose(logits[i].numpy(), want_logits[rows], rtol=0, atol=2e-6)
        else:
            np.testing.assert_allclose(logits[i].numpy(), want_logits, rtol=0, atol=2e-6)
        # spans: the GPU's own fp32 geometry through the exact-order span head -> bit-exact bounds
        res = model.last_result
        g32 = res.batch.geo_view(res.geom["geo"], i).cpu().numpy()
        valid = s != o
        reg = exact.span_head(np.ascontiguousarray(g32[np.where(valid, rows, 0)]), sd_np)
        want_sp = exact.span_decode(reg, sizes, stride)
        got_sp = dp[i].numpy()
        assert got_sp.shape == want_sp.shape and got_sp.dtype == np.int32
        np.testing.assert_array_equal(got_sp[valid], want_sp[valid])               # frame bounds bit-exact
        np.testing.assert_allclose(res.geom["viou"][res.batch.pair_slice(i)].cpu().numpy(), viou, rtol=1e-5)
    # the same call with RelNMS active (defaults.py:62: 64 proposals, rel_nms.py:10: threshold 0.5): the kept spans
    # are the oracle's greedy temporal NMS of the decoded spans above, bit for bit ([SPEC] s8)
    nms = BaseModel(_cfg(c, r, fdim, True, True, SPARSIFY=sparsify, NUM_SPANS=64)).eval()
    nms.load_state_dict(sd)
    with torch.no_grad():
        pp2, dp2, logits2 = nms(pls)
    for i, v in enumerate(vids):
        n = v.n_tracklets
        assert torch.equal(pp2[i], pp[i]) and torch.equal(logits2[i], logits[i])
        order = pp[i].numpy()
        s, o = order // n, order % n
        valid = s != o
        rows = np.where(valid, s * (n - 1) + o - (o > s), 0)
        ov = model.last_result.geom["overlap"][model.last_result.batch.pair_slice(i)].cpu().numpy()
        want, want_cnt = oheads.select_spans(dp[i].numpy(), ov[rows], 64, 0.5)
        got = dp2[i].numpy()
        assert got.shape == (len(order), 64, 2) and got.dtype == np.int16
        np.testing.assert_array_equal(got[valid], want[valid])
        np.testing.assert_array_equal(nms.last_result.span_count(i).cpu().numpy()[valid], want_cnt[valid])
        assert (got[~valid] == 0).all()


@pytest.mark.parametrize("sparsify", [False, True])
def test_triplet_records_match_predict_py_postprocessing(sparsify):
    """Row N1: top-20 per pair -> top-200 per video -> (score, triplet, tracklet ids), bit-exact."""
    from tspn_b200 import ops
    from tspn_b200.batch import HostBatch
    from tspn_b200.pipeline import PairStage, StageConfig
    c, r = 35, 132
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=2)
    vids = [synth.make_video(14, 90, c, seed=7), synth.make_video(3, 40, c, seed=8), synth.make_video(1, 5, c, seed=9)]
    for mirror in (False, True):
        stage = PairStage(StageConfig(n_classes=c, n_predicates=r, topk=64, sparsify=sparsify, precision="fp32",
                                      mirror_q4=mirror))
        stage.load_weights(sd)
        batch = HostBatch.from_videos(vids).to_device("cuda")
        res = stage.forward(batch)
        torch.cuda.synchronize()
        rec = res.records.cpu().numpy()
        cnt = res.record_counts.cpu().numpy()
        sc = ops.record_scores(res.records).cpu().numpy()
        ov = res.geom["overlap"].cpu().numpy()
        for i, v in enumerate(vids):
            n = v.n_tracklets
            logits = res.logits(i).cpu().numpy()
            if logits.shape[0] == 0:
                assert cnt[i] == 0
                continue
            pr = ogeo.enumerate_pairs(n)
            if sparsify:
                order = res.pair_proposals(i).cpu().numpy()
                s, o = order // n, order % n
                rows = s * (n - 1) + o - (o > s)
                pairs = pr[rows]
            else:
                rows = np.arange(n * (n - 1))
                pairs = pr
            w_score, w_trip, w_tid = oheads.postprocess_ref(logits, v.cls, pairs, 20, 200, fix_q4=True)
            m = len(w_score)
            assert cnt[i] == m
            np.testing.assert_array_equal(sc[i, :m], w_score)                       # scores bit-exact
            np.testing.assert_array_equal(rec[i, :m, 2], w_trip[:, 1])              # predicates
            np.testing.assert_array_equal(rec[i, :m, 4:6], w_tid)                   # tracklet ids
            np.testing.assert_array_equal(rec[i, :m, 1], w_trip[:, 0])              # subject class
            if mirror:       # quirk Q4: object class read from tracklet 0 (or 1 when the object is 0)
                q4 = v.cls[np.where(w_tid[:, 1] == 0, 1, 0)].argmax(axis=1)
                np.testing.assert_array_equal(rec[i, :m, 3], q4)
            else:
                np.testing.assert_array_equal(rec[i, :m, 3], w_trip[:, 2])
            grow = res.batch.pair_slice(i).start + w_tid[:, 0] * (n - 1) + w_tid[:, 1] - (w_tid[:, 1] > w_tid[:, 0])
            np.testing.assert_array_equal(rec[i, :m, 6:8], ov[grow])
            assert (rec[i, m:, 1:6] == -1).all()


@pytest.mark.parametrize("single", [True, False])
@pytest.mark.parametrize("precision", ["fp32", "tensor"])
def test_graphed_and_pipelined_stage_match_eager(precision, single):
    """CUDA-graph replay and the pinned-host serving loop return exactly what the eager call returns,
    for every batch fed through the same slots (stale state would show up on the second batch)."""
    from tspn_b200.batch import HostBatch
    from tspn_b200.pipeline import PairStage, StageConfig
    from tspn_b200.serving import PipelinedStage
    c, r = 35, 132
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=3)
    shapes = [(9, 120), (14, 300), (5, 77)]
    stage = PairStage(StageConfig(n_classes=c, n_predicates=r, topk=64, sparsify=True, precision=precision))
    stage.load_weights(sd, "cuda")
    hosts = [HostBatch.from_videos([synth.make_video(n, t, c, seed=10 * b + i) for i, (n, t) in enumerate(shapes)])
             for b in range(5)]
    want = []
    for h in hosts:
        res = stage.forward(h.to_device("cuda"))
        torch.cuda.synchronize()
        want.append({k: v.cpu().clone() for k, v in res.host_outputs().items()})
    # graph replay on refilled inputs
    batch = hosts[0].to_device("cuda")
    graphed = stage.capture(batch, single=single)
    timers = {}
    for h, w in zip(hosts, want):
        batch.copy_from(h)
        res = graphed.replay(timers=timers)
        torch.cuda.synchronize()
        assert timers["geo"][0].elapsed_time(timers["geo"][1]) > 0.0     # the geometry kernel's own events
        for k, v in res.host_outputs().items():
            assert torch.equal(v.cpu(), w[k]), k
    # pinned-host pipeline, two batches in flight
    pipe = PipelinedStage(stage, hosts[0], device="cuda", depth=2, single_graph=single,
                          compute_streams=2 if single else 1)
    got = [{k: v.clone() for k, v in out.items()} for out in pipe.run(iter(hosts))]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert set(g) == set(w)
        for k in w:
            assert torch.equal(g[k], w[k]), k


def test_compact_transport_is_lossless():
    """u16 boxes + u8 motion counts (the default transport when the values allow it) against fp32 transport:
    every output of the stage bit for bit."""
    from tspn_b200.batch import HostBatch
    from tspn_b200.pipeline import PairStage, StageConfig
    c, r = 35, 132
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=3)
    vids = [synth.make_video(n, t, c, seed=40 + i) for i, (n, t) in enumerate([(9, 120), (14, 300), (2, 5), (5, 77)])]
    vids[1].motion[3, :1000] = 0                       # an empty histogram block stays zero
    vids[1].motion[4, 7] = 255
    for precision in ("fp32", "tensor"):
        stage = PairStage(StageConfig(n_classes=c, n_predicates=r, topk=64, sparsify=True, precision=precision))
        stage.load_weights(sd, "cuda")
        outs = []
        for compact in (True, False):
            host = HostBatch.from_videos(vids, compact=compact)
            assert host.boxes_compact == compact and host.motion_compact == compact
            res = stage.forward(host.to_device("cuda"))
            torch.cuda.synchronize()
            outs.append(({k: v.cpu() for k, v in res.host_outputs().items()}, res.geom["geo"].cpu(),
                         host.h2d_bytes()))
        assert outs[0][2] < 0.5 * outs[1][2]
        assert torch.equal(outs[0][1], outs[1][1])
        for k in outs[1][0]:
            assert torch.equal(outs[0][0][k], outs[1][0][k]), k


@pytest.mark.parametrize("with_capacity", [False, True])
def test_delta_coded_boxes_expand_to_the_same_rows(with_capacity):
    """Span-packed boxes, raw u16 against delta-coded (first box + i8 differences, ``TSPN_PACKED_DELTA``): the dense
    fp32 rows ``tspn_unpack_boxes_spans`` writes are the host's boxes bit for bit either way - spans of 0..3 frames,
    spans longer than several 256-frame scan tiles, a tracklet that cannot be delta-coded between ones that can, steps
    of exactly -128 / +127 - and the refill of a device batch switches coding per tracklet without a new layout."""
    from tests.test_cpu_ragged import _edge_video
    from tspn_b200 import _lib
    from tspn_b200.batch import VT_BOX_OFF, VT_N, VT_T, VT_TB, Capacity, HostBatch
    vids = [synth.make_video(12, 2000, 35, seed=1), _edge_video(), synth.make_video(3, 5, 35, seed=2),
            synth.make_video(5, 777, 35, seed=3, full_span=True)]
    vids[3].boxes[2, 300:, 1] += 300.0                 # raw tracklet in a video of delta tracklets
    vids[3].boxes[2, 300:, 3] += 300.0
    cap = Capacity.for_shapes([(12, 2048), (9, 64), (3, 8), (5, 1024)], 35, videos=5) if with_capacity else None

    hosts = {d: HostBatch.from_videos(vids, capacity=cap, delta=d) for d in (True, False)}
    assert hosts[True].h2d_bytes() < hosts[False].h2d_bytes()
    off = hosts[True].box_off.numpy()
    assert ((off & _lib.PACKED_DELTA) != 0).sum() >= 12 + 5 + 4
    dev = hosts[False].to_device("cuda")
    for d in (False, True, False, True):
        dev.copy_from(hosts[d])
        torch.cuda.synchronize()
        rows = dev.boxes.cpu().numpy().reshape(-1, 4)
        for v, vid in enumerate(vids):
            r = hosts[d].table[v]
            n, t, tb, b0 = int(r[VT_N]), int(r[VT_T]), int(r[VT_TB]), int(r[VT_BOX_OFF])
            got = rows[b0:b0 + n * tb].reshape(n, tb, 4)
            np.testing.assert_array_equal(got[:, :t], vid.boxes, err_msg=f"delta={d} video {v}")
            np.testing.assert_array_equal(got[:, t:], 0)
