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
def test_span_head_tensor(cin, k, t, a):
    """DPNHead (dpn.py:55-73) as a tcgen05 implicit GEMM: bf16 operands, fp32 accumulation.  Channel
    counts cover one chunk, several chunks, a ragged last chunk (320 = 256 + 64) and a ragged K slab
    (72); row tiles straddle pairs for every T here."""
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
