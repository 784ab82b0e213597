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
