"""PPNHead on the tensor cores (tcgen05, tf32 operands, fp32 accumulate) against the float64 definition:
scores within 1e-2 absolute (BASELINE.json north_star), selection identical wherever the score gap allows."""
import numpy as np
import pytest
import torch

from oracle import heads as oheads
from tspn_b200 import ops, synth
from tspn_b200.batch import Capacity, HostBatch
from tspn_b200.pipeline import PPN_PREFIX, PairStage, StageConfig

pytestmark = pytest.mark.gpu


def _weights(sd, dev="cuda"):
    return [torch.from_numpy(sd[PPN_PREFIX + k]).to(dev) for k in ops.PPN_KEYS]


@pytest.mark.parametrize("c,shapes", [
    (80, [(64, 16), (64, 16), (3, 16), (1, 16), (0, 16), (33, 16)]),      # VidOR classes; a tile of 128 rows cuts a video
    (35, [(20, 8), (40, 8), (7, 8)]),                                     # C not a multiple of 8: K padding
    (80, [(256, 8), (130, 8), (5, 8)]),                                   # N > 128: two subject tiles per video
    (128, [(90, 8)]),
])
def test_scores_within_tolerance_of_f64_definition(c, shapes):
    sd = synth.make_weights(c, 50, synth.feature_dim(c), seed=1)
    vids = [synth.make_video(n, t, c, seed=10 + i) for i, (n, t) in enumerate(shapes)]
    batch = HostBatch.from_videos(vids).to_device("cuda")
    assert ops.relationness_tc_supported(batch)
    got = ops.relationness(batch, _weights(sd), precision="tensor")
    exact = ops.relationness(batch, _weights(sd), precision="fp32")
    torch.cuda.synchronize()
    worst = 0.0
    for i, v in enumerate(vids):
        want = oheads.ppn_head_f64(v.cls, v.cls, sd)
        g = batch.score_view(got, i).cpu().numpy()
        assert g.shape == want.shape
        if g.size:
            worst = max(worst, float(np.abs(g - want).max()))
            np.testing.assert_allclose(g, want, rtol=0, atol=1e-2)       # the north star's bf16-MLP tolerance
            np.testing.assert_allclose(batch.score_view(exact, i).cpu().numpy(), want, rtol=0, atol=2e-6)
    assert worst < 5e-3                                                  # tf32 operands: far inside it


@pytest.mark.parametrize("exclude", [False, True])
@pytest.mark.parametrize("n_max,k", [(64, 256), (200, 1024)])
def test_fused_tc_scores_topk_selects_like_the_oracle_up_to_the_score_gap(n_max, k, exclude):
    c = 80
    sd = synth.make_weights(c, 50, synth.feature_dim(c), seed=2)
    rng = np.random.Generator(np.random.PCG64(n_max))
    shapes = [(n_max, 8)] + [(int(rng.integers(2, n_max)), 8) for _ in range(4)]
    vids = [synth.make_video(n, t, c, seed=30 + i) for i, (n, t) in enumerate(shapes)]
    cap = Capacity.for_shapes(shapes + [(n_max, 8)], c, videos=8)
    batch = HostBatch.from_videos(vids, capacity=cap).to_device("cuda")      # capacity grid: empty trailing videos
    scores, idx, val, row = ops.relationness_topk(batch, _weights(sd), k, exclude_diagonal=exclude, precision="tensor")
    torch.cuda.synchronize()
    for i, v in enumerate(vids):
        n = v.n_tracklets
        sc = batch.score_view(scores, i).cpu().numpy().astype(np.float64)
        want_sc = oheads.ppn_head_f64(v.cls, v.cls, sd)
        np.testing.assert_allclose(sc, want_sc, rtol=0, atol=1e-2)
        cand = sc.copy()
        if exclude:
            cand[np.arange(n), np.arange(n)] = -np.inf
        k_eff = min(k, n * (n - 1) if exclude else n * n)
        got = idx[i, :k_eff].cpu().numpy()
        # the kernel's selection is exactly the stable top-K of the scores IT produced ...
        np.testing.assert_array_equal(got, oheads.topk_stable(cand.astype(np.float32), k)[:k_eff])
        assert (idx[i, k_eff:] == -1).all()
        np.testing.assert_array_equal(val[i, :k_eff].cpu().numpy(), sc.reshape(-1)[got].astype(np.float32))
        # ... and agrees with the float64 definition's selection wherever the K-th score gap exceeds 2e-2
        ref = want_sc.copy()
        if exclude:
            ref[np.arange(n), np.arange(n)] = -np.inf
        flat = np.sort(ref.reshape(-1))[::-1]
        if k_eff < flat.size and np.isfinite(flat[k_eff]):
            kth = flat[k_eff - 1]
            sure_in = np.flatnonzero(ref.reshape(-1) > kth + 2e-2)
            sure_out = np.flatnonzero(ref.reshape(-1) < kth - 2e-2)
            assert np.isin(sure_in, got).all() and not np.isin(sure_out, got).any()
        s, o = got // n, got % n
        want_row = np.where(s == o, -1, batch.pair_slice(i).start + s * (n - 1) + o - (o > s))
        np.testing.assert_array_equal(row[i, :k_eff].cpu().numpy(), want_row)
    for vpad in range(len(vids), batch.num_videos):
        assert (idx[vpad] == -1).all()


def test_stage_with_tensor_relationness_runs_the_heads_on_its_own_selection():
    c, r = 80, 50
    sd = synth.make_weights(c, r, synth.feature_dim(c), dpn_in=8, seed=0)
    vids = [synth.make_video(24, 300, c, seed=1), synth.make_video(9, 120, c, seed=2)]
    st = PairStage(StageConfig(n_classes=c, n_predicates=r, topk=64, sparsify=True, precision="tensor",
                               relationness_precision="tensor", num_span_proposals=16))
    st.load_weights(sd, "cuda")
    batch = HostBatch.from_videos(vids).to_device("cuda")
    res = st.forward(batch)
    graphed = st.capture(batch)
    rep = graphed.replay()
    torch.cuda.synchronize()
    for k, v in res.host_outputs().items():
        assert torch.equal(v, rep.host_outputs()[k]), k
    for i, v in enumerate(vids):
        n = v.n_tracklets
        sc = batch.score_view(res.scores, i).cpu().numpy()
        np.testing.assert_allclose(sc, oheads.ppn_head_f64(v.cls, v.cls, sd), rtol=0, atol=1e-2)
        cand = sc.copy()
        cand[np.arange(n), np.arange(n)] = -np.inf
        np.testing.assert_array_equal(res.pair_proposals(i).cpu().numpy(),
                                      oheads.topk_stable(cand, 64)[:min(64, n * (n - 1))])
        assert res.logits(i).shape == (min(64, n * (n - 1)), r)
