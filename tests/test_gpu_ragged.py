"""Ragged batches through capacity buckets, and the span selection ([SPEC] s8), on the GPU.

A batch packed for a ``Capacity`` (buffers, grids and the pair kernel's chunk fixed by upper bounds; the table's
sentinel row carries the true totals to the device) must give, bit for bit, what the same videos give as an
exact-shape batch - eagerly, through ONE graph captured for the capacity, and through the serving loop."""
import numpy as np
import pytest
import torch

from oracle import heads as oheads
from tspn_b200 import _lib, ops, synth
from tspn_b200.batch import Capacity, HostBatch
from tspn_b200.pipeline import PairStage, StageConfig
from tspn_b200.serving import PipelinedStage, host_batches_for

pytestmark = pytest.mark.gpu

C, R = 35, 132
VIDVRD = ((15.0, 30.0, 45.0, 60.0), 7.5)


def _stage(precision="tensor", n_spans=16, topk=64, sparsify=True, **kw):
    sd = synth.make_weights(C, R, synth.feature_dim(C), dpn_in=8, seed=3)
    st = PairStage(StageConfig(n_classes=C, n_predicates=R, topk=topk, sparsify=sparsify, precision=precision,
                               num_span_proposals=n_spans, **kw))
    st.load_weights(sd, "cuda")
    return st


def _outputs(res):
    torch.cuda.synchronize()
    return {k: v.cpu().clone() for k, v in res.host_outputs().items()}


# ---------------------------------------------------------------------------------------------------------
# span selection
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("thr", [0.3, 0.5, 1.0])
@pytest.mark.parametrize("m,a_n,int16", [(504, 4, True), (644, 4, False), (36, 4, True), (1028, 4, True), (96, 3, True)])
def test_span_select_explicit_windows_bit_exact(m, a_n, int16, thr):
    """Every row has M candidates and an explicit window: random spans (no anchor structure at all - the group
    bounds are data-derived), duplicates (ties to the lower index), empty windows."""
    rng = np.random.Generator(np.random.PCG64(m))
    r_n = 37
    s = rng.integers(0, 1900, (r_n, m))
    c = np.stack([s, np.minimum(s + rng.integers(1, 700, (r_n, m)), 2000)], axis=-1).astype(np.int32)
    c[5, 10:40] = c[5, 10]                                               # identical candidates
    c[6] = c[6, 0]
    w = np.stack([rng.integers(0, 1000, r_n), rng.integers(500, 2000, r_n)], axis=1).astype(np.int32)
    w[2] = (700, 700)
    w[3] = (900, 100)
    got, cnt = ops.span_select(torch.from_numpy(c).cuda(), a_n, 16.0, 64, thr, windows=torch.from_numpy(w).cuda(),
                               int16=int16)
    torch.cuda.synchronize()
    want, wcnt = oheads.select_spans(c, w, 64, thr)
    assert got.dtype == (torch.int16 if int16 else torch.int32)
    np.testing.assert_array_equal(cnt.cpu().numpy(), wcnt)
    np.testing.assert_array_equal(got.cpu().numpy().astype(np.int32), want)


def test_span_select_on_decoded_anchors_of_a_ragged_batch():
    """Table mode: candidates = the decoded anchors of the stage (location major, anchor minor), per-video
    location counts, windows from the tracklet spans, padding rows -> count 0."""
    sizes, stride = VIDVRD
    vids = [synth.make_video(9, 500, C, seed=1), synth.make_video(4, 77, C, seed=2), synth.make_video(12, 300, C, seed=3)]
    st = _stage("fp32", n_spans=0, anchor_sizes=sizes, anchor_stride=stride)
    batch = HostBatch.from_videos(vids).to_device("cuda")
    res = st.forward(batch)
    torch.cuda.synchronize()
    raw = res.span_buffers
    st64 = _stage("fp32", n_spans=64, anchor_sizes=sizes, anchor_stride=stride)
    res64 = st64.forward(batch)
    torch.cuda.synchronize()
    ov = res.geom["overlap"].cpu().numpy()
    for i, v in enumerate(vids):
        n = v.n_tracklets
        order = res.pair_proposals(i).cpu().numpy()
        s, o = order // n, order % n
        rows = batch.pair_slice(i).start + s * (n - 1) + o - (o > s)
        cands = res.spans[i].cpu().numpy()
        want, wcnt = oheads.select_spans(cands, ov[rows], 64, 0.5)
        np.testing.assert_array_equal(res64.spans[i].cpu().numpy().astype(np.int32), want)
        np.testing.assert_array_equal(res64.span_count(i).cpu().numpy(), wcnt)
        assert torch.equal(res64.logits(i), res.logits(i))
    assert raw is not None


# ---------------------------------------------------------------------------------------------------------
# capacity batches
# ---------------------------------------------------------------------------------------------------------
SHAPES = [(9, 120), (14, 300), (2, 5), (1, 40), (0, 9), (20, 511)]


@pytest.mark.parametrize("precision,sparsify", [("tensor", True), ("fp32", True), ("fp32", False), ("tensor", False)])
def test_capacity_batch_equals_exact_batch(precision, sparsify):
    vids = [synth.make_video(n, t, C, seed=60 + i) for i, (n, t) in enumerate(SHAPES)]
    st = _stage(precision, sparsify=sparsify)
    exact = st.forward(HostBatch.from_videos(vids).to_device("cuda"))
    want = _outputs(exact)
    want_geo = exact.geom["geo"].cpu()
    # a roomy capacity: more videos, tracklets, pairs, items than the batch has; a longer max T in the same class
    cap = Capacity.for_shapes(SHAPES + [(25, 512), (25, 400), (3, 3)], C, videos=12)
    host = HostBatch.from_videos(vids, capacity=cap)
    batch = host.to_device("cuda")
    assert batch.num_videos == 12 and batch.num_real == len(vids) and batch.total_pairs > batch.actual_pairs
    res = st.forward(batch)
    got = _outputs(res)
    assert set(got) == set(want)
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert torch.equal(got[k], want[k]), k
    n_geo = int(host.actual[_lib.TOT_GEO_FLOATS])
    assert torch.equal(res.geom["geo"][:n_geo].cpu(), want_geo)
    for i in range(len(vids)):
        assert torch.equal(res.logits(i), exact.logits(i))
        if res.spans is not None:
            assert torch.equal(res.spans[i], exact.spans[i])


def test_one_graph_serves_ragged_batches_of_its_capacity():
    """The graph is captured ONCE for the capacity, on one batch; every other batch that fits - fewer or more
    videos, other tracklet and frame counts, a batch of one, an empty batch - replays it and returns exactly what
    its own eager exact-shape call returns."""
    rng = np.random.Generator(np.random.PCG64(5))
    groups = [[(9, 120), (14, 300), (5, 77)], [(20, 500)], [(3, 30), (3, 31), (3, 32), (3, 33), (7, 400), (2, 2)], [],
              [(16, 256), (1, 10)], [(9, 120), (14, 300), (5, 77)]]
    cap = Capacity.for_shapes([(20, 512), (16, 512), (14, 300), (9, 200)], C, videos=8)
    st = _stage("tensor")
    vids = [[synth.make_video(n, t, C, seed=int(rng.integers(1 << 20))) for n, t in g] for g in groups]
    want = [_outputs(st.forward(HostBatch.from_videos(v).to_device("cuda"))) if v else None for v in vids]
    hosts = [HostBatch.from_videos(v, capacity=cap) if v else HostBatch([], [], [], [], capacity=cap) for v in vids]
    batch = hosts[0].to_device("cuda")
    graphed = st.capture(batch)
    for h, w in zip(hosts, want):
        batch.copy_from(h)
        got = _outputs(graphed.replay())
        if w is None:
            assert all(v.shape[0] == 0 for v in got.values())
            continue
        for k in w:
            assert torch.equal(got[k], w[k]), k
    # resident refill (device to device), as the bench's resident-input loop does
    resident = hosts[2].to_device("cuda")
    batch.copy_from_device(resident)
    got = _outputs(graphed.replay())
    for k in want[2]:
        assert torch.equal(got[k], want[2][k]), k


def test_serving_loop_over_buckets_matches_eager():
    """Ragged videos -> batches per chunk class -> one bucket (slots + graph) per class; results in submission
    order equal the eager exact-shape results; nothing is re-captured for later passes."""
    shapes = [(n, t) for n, t in synth.config_shapes("vidvrd_test", 3, 40)]
    vids = [synth.make_video(n, t, C, seed=200 + i) for i, (n, t) in enumerate(shapes)]
    st = _stage("tensor", anchor_sizes=VIDVRD[0], anchor_stride=VIDVRD[1])
    hosts, batch_vids, caps = host_batches_for(vids, C, geo_budget_bytes=48 << 20, max_videos=6, merge_below_bytes=0)
    assert len(caps) >= 2 and len(hosts) > len(caps)
    want = [_outputs(st.forward(HostBatch.from_videos([vids[i] for i in ids]).to_device("cuda"))) for ids in batch_vids]
    pipe = PipelinedStage(st, [hosts[[h.capacity for h in hosts].index(c)] for c in caps.values()], depth=2)
    n_buckets = len(pipe.buckets)
    for _ in range(2):
        got = [{k: v.clone() for k, v in out.items()} for out in pipe.run(iter(hosts))]
        assert len(got) == len(want)
        for g, w in zip(got, want):
            for k in w:
                assert torch.equal(g[k], w[k]), k
    assert len(pipe.buckets) == n_buckets


def test_multi_chunk_rows_have_single_writer_sums():
    """T > 2048: a pair's volume sums come from several work items, each storing its own chunk slot (no zeroing,
    no atomics); a capacity with more chunk slots than the batch needs gives the same bits."""
    vids = [synth.make_video(5, 4100, C, seed=1), synth.make_video(7, 2049, C, seed=2), synth.make_video(3, 300, C, seed=3)]
    exact = HostBatch.from_videos(vids).to_device("cuda")
    a = ops.pair_geometry(exact, write_geo=True)
    cap = Capacity.for_shapes([(8, 6200), (8, 4100), (8, 4100), (4, 300)], C, videos=5)
    assert cap.max_chunks == 4 and int(exact.totals[_lib.TOT_MAX_CHUNKS]) == 3
    padded = HostBatch.from_videos(vids, capacity=cap).to_device("cuda")
    b = ops.pair_geometry(padded, write_geo=True)
    torch.cuda.synchronize()
    p, g = exact.total_pairs, int(exact.totals[_lib.TOT_GEO_FLOATS])
    for k in ("viou", "tiou", "overlap"):
        assert torch.equal(a[k], b[k][:p]), k
    assert torch.equal(a["geo"], b["geo"][:g])
