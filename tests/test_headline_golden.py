"""Outputs of the unmodified reference at the HEADLINE shape (BASELINE.json configs[2]: N=64, T=2000, 80 classes, 50
predicates, F=11160; tests/golden/make_golden.py --headline) against the oracle (CPU) and the CUDA path (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import exact, geometry as ogeo, heads as oheads
from tests.golden.make_golden import synth_features
from tspn_b200 import synth

N, T, C, R, SEED = 64, 2000, 80, 50, 0


@pytest.fixture(scope="module")
def headline():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs_headline.npz"))


def test_oracle_matches_the_reference_at_the_headline_shape(headline):
    v = synth.make_video(N, T, C, seed=SEED, full_span=True)
    np.testing.assert_array_equal(ogeo.cubic_iou_ref(v.boxes, v.boxes), headline["cubic_iou_f32_C"])     # the port
    # the float64 definition against the reference's own fp32 accumulation (3.3e-6 off at T = 2000)
    np.testing.assert_allclose(ogeo.cubic_iou_f64(v.boxes, v.boxes), headline["cubic_iou_f32_C"], rtol=1e-5)
    vr = synth.make_video(N, T, C, seed=SEED + 1)
    rows = headline["viou_v2_C_rows"]
    pr = ogeo.enumerate_pairs(N)[rows]
    np.testing.assert_array_equal(ogeo.viou_pairs_ref((vr.boxes, vr.span, pr)), headline["viou_v2_C"])
    _, viou, _, _ = ogeo.pair_geometry(vr.boxes, vr.span, pr[:, 0], pr[:, 1])
    np.testing.assert_allclose(viou, headline["viou_v2_C"], rtol=1e-12)
    sd = synth.make_weights(C, R, synth.feature_dim(C), dpn_in=8, n_anchors=4, seed=SEED)
    np.testing.assert_allclose(oheads.ppn_head_ref(v.cls, v.cls, sd).numpy(), headline["ppn_scores_C"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(exact.relationness(v.cls, sd), headline["ppn_scores_C"], rtol=0, atol=2e-6)
    feats = synth_features(N * (N - 1), synth.feature_dim(C), SEED)
    np.testing.assert_allclose(oheads.relation_predictor_f64(feats, sd), headline["rel_logits_C"], rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_cuda_path_matches_the_reference_at_the_headline_shape(headline):
    from tspn_b200 import ops
    from tspn_b200.batch import HostBatch
    from tspn_b200.pipeline import CLS_PREFIX, PPN_PREFIX
    v = synth.make_video(N, T, C, seed=SEED, full_span=True)
    b = torch.from_numpy(v.boxes).cuda()
    np.testing.assert_allclose(ops.cubic_iou(b, b).cpu().numpy(), headline["cubic_iou_f32_C"], rtol=1e-5)
    vr = synth.make_video(N, T, C, seed=SEED + 1)
    batch = HostBatch.from_videos([vr]).to_device("cuda")
    geom = ops.pair_geometry(batch, write_geo=False)
    np.testing.assert_allclose(geom["viou"].cpu().numpy()[headline["viou_v2_C_rows"]], headline["viou_v2_C"], rtol=1e-5)
    sd = synth.make_weights(C, R, synth.feature_dim(C), dpn_in=8, n_anchors=4, seed=SEED)
    fb = HostBatch.from_videos([v]).to_device("cuda")
    w = [torch.from_numpy(sd[PPN_PREFIX + k]).cuda() for k in ops.PPN_KEYS]
    np.testing.assert_allclose(ops.relationness(fb, w).view(N, N).cpu().numpy(), headline["ppn_scores_C"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(ops.relationness(fb, w, precision="tensor").view(N, N).cpu().numpy(),
                               headline["ppn_scores_C"], rtol=0, atol=1e-2)
    feats = torch.from_numpy(synth_features(N * (N - 1), synth.feature_dim(C), SEED)).cuda()
    wt, bias = torch.from_numpy(sd[CLS_PREFIX + "weight"]).cuda(), torch.from_numpy(sd[CLS_PREFIX + "bias"]).cuda()
    np.testing.assert_allclose(ops.predicate_head(feats, wt, bias, precision="fp32").cpu().numpy(),
                               headline["rel_logits_C"], rtol=0, atol=2e-6)
    ld = ops.padded(feats.shape[1], 4)
    buf = torch.zeros((feats.shape[0], ld), device="cuda")
    buf[:, :feats.shape[1]] = feats
    np.testing.assert_allclose(ops.predicate_head(buf[:, :feats.shape[1]], wt, bias, precision="tensor").cpu().numpy(),
                               headline["rel_logits_C"], rtol=0, atol=1e-2)
