"""The opt-in WINDOWED geometry layout (tspn_pair_geo_viou_windowed) against the dense parity layout.

Every channel of a pair's dense row is zero outside the pair's overlap window [a, b) and channel 7 is the window's
indicator ([SPEC] s2), so the windowed rows - frames [a & ~3, (b + 3) & ~3) of channels 0..6 - carry the same
information.  Here: the windowed rows equal the dense rows bit for bit on those frames, the dense rows are zero on all
other frames (so nothing was dropped), the offsets equal a numpy prefix sum, and vIoU / tIoU / overlap are unchanged -
on ragged batches, multi-chunk videos (T > 2048), the clipped vIoU variant, a capacity batch refilled under a captured
graph, and the bench step (survivor path)."""
import numpy as np
import pytest
import torch

from tspn_b200 import _lib, ops, synth
from tspn_b200.batch import Capacity, HostBatch
from tspn_b200.pipeline import PairStage, StageConfig

pytestmark = pytest.mark.gpu

C, R = 35, 132


def _want_offsets(vids):
    """numpy restatement of tspn_geo_window_offsets: exclusive prefix sum of 7 * Lw in pair order."""
    lens = []
    for v in vids:
        n = v.n_tracklets
        for s in range(n):
            for o in range(n):
                if o == s:
                    continue
                a, b = max(v.span[s, 0], v.span[o, 0]), min(v.span[s, 1], v.span[o, 1])
                lens.append(7 * ((((b + 3) & ~3) - (a & ~3)) if b > a else 0))
    lens = np.asarray(lens, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(lens)[:-1]]) if len(lens) else lens, int(lens.sum())


def _check_against_dense(batch, vids, dense, win):
    geo_d, geo_w = dense["geo"].cpu().numpy(), win["geo"].cpu().numpy()
    off = win["geo_off"].cpu().numpy()
    ov = win["overlap"].cpu().numpy()
    want_off, want_total = _want_offsets(vids)
    p_all = len(want_off)
    np.testing.assert_array_equal(off[:p_all], want_off)
    assert int(win["geo_total"].item()) == want_total
    for k in ("viou", "tiou", "overlap"):
        np.testing.assert_array_equal(win[k].cpu().numpy()[:p_all], dense[k].cpu().numpy()[:p_all])
    th = batch.table_host
    for i, v in enumerate(vids):
        n, tp = v.n_tracklets, int(th[i, _lib.VT_TP])
        p0, g0 = int(th[i, _lib.VT_PAIR_OFF]), int(th[i, _lib.VT_GEO_OFF])
        rows = geo_d[g0:g0 + n * (n - 1) * 8 * tp].reshape(n * (n - 1), 8, tp)
        for p in range(n * (n - 1)):
            a, b = int(ov[p0 + p, 0]), int(ov[p0 + p, 1])
            a4, b4 = a & ~3, (b + 3) & ~3
            lw = b4 - a4 if b > a else 0
            got = geo_w[off[p0 + p]:off[p0 + p] + 7 * lw].reshape(7, lw)
            np.testing.assert_array_equal(got.view(np.uint32), rows[p, :7, a4:a4 + lw].view(np.uint32))
            outside = np.ones(tp, dtype=bool)
            outside[a4:a4 + lw] = False
            assert not rows[p][:, outside].any()                 # nothing outside the window was dropped
            mask = np.zeros(tp, dtype=np.float32)
            mask[a:b] = 1.0
            np.testing.assert_array_equal(rows[p, 7], mask)      # channel 7 is implied by [a, b)


@pytest.mark.parametrize("clipped", [False, True])
@pytest.mark.parametrize("shapes", [[(9, 500), (4, 77), (12, 300)],        # chunk 512, ragged
                                    [(7, 1000), (3, 5), (2, 1024)],         # chunk 1024, a 5-frame video, N = 2
                                    [(34, 2000), (5, 1300)],                # chunk 2048, two object groups
                                    [(6, 4100), (3, 2049)]])                # multi-chunk videos
def test_windowed_rows_equal_dense_rows(shapes, clipped):
    vids = [synth.make_video(n, t, C, seed=10 + i) for i, (n, t) in enumerate(shapes)]
    batch = HostBatch.from_videos(vids).to_device("cuda")
    dense = ops.pair_geometry(batch, clipped=clipped)
    win = ops.pair_geometry(batch, clipped=clipped, windowed=True)
    torch.cuda.synchronize()
    assert win["geo"].numel() == dense["geo"].numel() // 8 * 7
    _check_against_dense(batch, vids, dense, win)


def test_windowed_disjoint_and_degenerate_spans():
    """Pairs without an overlap window take no room; one-frame windows; windows at frame 0 and at T."""
    v = synth.make_video(6, 257, C, seed=3, full_span=True)
    v.span[:] = [(0, 1), (0, 257), (256, 257), (100, 101), (101, 200), (3, 6)]
    v.boxes[:] = np.where((np.arange(257)[None, :, None] >= v.span[:, :1, None]) &
                          (np.arange(257)[None, :, None] < v.span[:, 1:, None]), v.boxes, 0)
    batch = HostBatch.from_videos([v], compact=False).to_device("cuda")
    dense = ops.pair_geometry(batch)
    win = ops.pair_geometry(batch, windowed=True)
    torch.cuda.synchronize()
    _check_against_dense(batch, [v], dense, win)


def test_windowed_c_abi_rejects_missing_offsets_and_dense_ctas():
    v = synth.make_video(4, 64, C, seed=1)
    batch = HostBatch.from_videos([v]).to_device("cuda")
    out = ops.pair_geometry_outputs(batch, windowed=True)
    lib, tot = _lib.load(), batch.totals

    def call(geo, off, flags):
        return lib.tspn_pair_geo_viou_windowed(
            _lib.ptr(batch.table), batch.num_videos, int(tot[_lib.TOT_ITEMS]), int(tot[_lib.TOT_GEO_CHUNK]),
            int(tot[_lib.TOT_MAX_CHUNKS]), batch.total_tracklets, batch.total_pairs, int(tot[_lib.TOT_BOXES]),
            _lib.ptr(batch.boxes), _lib.ptr(batch.span), geo, off, _lib.ptr(out["viou"]), _lib.ptr(out["tiou"]),
            _lib.ptr(out["overlap"]), flags, _lib.ptr(out["workspace"]), _lib.stream_ptr())
    assert call(_lib.ptr(out["geo"]), None, 0) != 0
    assert call(None, _lib.ptr(out["geo_off"]), 0) != 0
    assert call(_lib.ptr(out["geo"]), _lib.ptr(out["geo_off"]), _lib.GEO_DENSE_CTAS) != 0
    assert call(_lib.ptr(out["geo"]), _lib.ptr(out["geo_off"]), 0) == 0
    torch.cuda.synchronize()


def _stage(layout, **kw):
    sd = synth.make_weights(C, R, synth.feature_dim(C), dpn_in=8, seed=3)
    st = PairStage(StageConfig(n_classes=C, n_predicates=R, topk=64, sparsify=True, precision="tensor",
                               num_span_proposals=16, geo_layout=layout, **kw))
    st.load_weights(sd, "cuda")
    return st


def test_bench_step_windowed_under_a_capacity_graph_refilled():
    """The survivor-path step with geo_layout='windowed' through ONE graph captured for a capacity: after every
    refill the offsets follow the new spans (part of the upload), the rows equal the dense step's rows, and every
    other output of the step is identical to the dense step's."""
    cap = Capacity.for_shapes([(16, 600), (12, 512), (9, 300)], C, videos=3)
    sets = [[synth.make_video(12, 500, C, seed=1), synth.make_video(7, 65, C, seed=2)],
            [synth.make_video(16, 600, C, seed=5), synth.make_video(3, 12, C, seed=6), synth.make_video(9, 300, C, seed=7)]]
    st_d, st_w = _stage("dense"), _stage("windowed")
    batch_d = HostBatch.from_videos(sets[0], capacity=cap).to_device("cuda")
    batch_w = HostBatch.from_videos(sets[0], capacity=cap).to_device("cuda")
    g_d, g_w = st_d.capture(batch_d), st_w.capture(batch_w)
    for vids in sets[::-1] + sets:
        batch_d.copy_from(HostBatch.from_videos(vids, capacity=cap))
        batch_w.copy_from(HostBatch.from_videos(vids, capacity=cap))
        res_d, res_w = g_d.replay(), g_w.replay()
        torch.cuda.synchronize()
        _check_against_dense(batch_w, vids, res_d.geom, res_w.geom)
        out_d, out_w = res_d.host_outputs(), res_w.host_outputs()
        assert out_d.keys() == out_w.keys()
        for k in out_d:
            assert torch.equal(out_d[k], out_w[k]), k
        rows = res_w.geo_window_rows(len(vids) - 1)
        dense_rows = batch_d.geo_rows(res_d.geom["geo"], len(vids) - 1)
        ov = res_w.geom["overlap"][batch_w.pair_slice(len(vids) - 1)].cpu().numpy()
        for p, r in enumerate(rows):
            a4 = int(ov[p, 0]) & ~3
            assert torch.equal(r, dense_rows[p, :7, a4:a4 + r.shape[1]])


def test_windowed_layout_needs_the_survivor_path():
    sd = synth.make_weights(C, R, synth.feature_dim(C), dpn_in=8, seed=3)
    st = PairStage(StageConfig(n_classes=C, n_predicates=R, topk=64, sparsify=False, precision="fp32",
                               geo_layout="windowed"))
    st.load_weights(sd, "cuda")
    batch = HostBatch.from_videos([synth.make_video(5, 100, C, seed=1)]).to_device("cuda")
    with pytest.raises(ValueError, match="windowed"):
        st.forward(batch)
