"""Row N4: the per-segment relation-feature container (vrdataset.py:190-217) -> PairList -> model."""
import numpy as np
import pytest
import torch

from oracle import features as ofeat, geometry as ogeo, heads as oheads
from tspn_b200 import formats, synth


def _segment(tmp_path, n_prop=6, n_gt=2, c=35, seed=0):
    """A segment file as the upstream extractor writes it: N proposals (trackid -1) + M ground-truth tracklets,
    all ordered pairs among the N + M, raw (un-normalised) BoW counts."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = n_prop + n_gt
    trackid = np.concatenate([np.full(n_prop, -1), np.arange(n_gt)]).astype(np.int64)
    perm = rng.permutation(n)                       # ground truth is not necessarily last
    trackid = trackid[perm]
    pairs = ogeo.enumerate_pairs(n)
    f = 2 * c + 11000
    feats = rng.random((pairs.shape[0], f), dtype=np.float32)
    feats[:, 2 * c:2 * c + 8000] = rng.poisson(0.05, size=(pairs.shape[0], 8000))
    feats[3, 2 * c:2 * c + 1000] = 0                # an empty histogram stays zero
    iou = rng.random((n, n), dtype=np.float32)
    path = str(tmp_path / (formats.segment_signature("ILSVRC2015_train_00005003", 0, 30) + "-relation.npz"))
    np.savez(path, trackid=trackid, pairs=pairs, feats=feats, iou=iou)
    cls = synth.make_video(n_prop, 8, c, seed=seed).cls
    return path, trackid, pairs, feats, iou, cls


def test_segment_container_to_pair_list(tmp_path):
    path, trackid, pairs, feats, iou, cls = _segment(tmp_path)
    assert path.endswith("ILSVRC2015_train_00005003-0000-0030-relation.npz")
    got = formats.load_segment(path)
    for a, b in zip(got, (pairs, feats, iou, trackid)):
        np.testing.assert_array_equal(a, b)
    pl = formats.segment_pair_list(*got, cls)
    keep = (trackid[pairs[:, 0]] < 0) & (trackid[pairs[:, 1]] < 0)            # vrdataset.py:140-145
    assert len(pl) == keep.sum() == 6 * 5 and pl.get_field("num_tracklets") == 6
    np.testing.assert_array_equal(pl.get_field("tracklet_pairs").numpy(), pairs[keep])
    want = feats[keep].copy()
    for b0 in range(70, 8070, 1000):                                         # vrdataset.py:227-236
        want[:, b0:b0 + 1000] = ofeat.l1_normalize_ref(want[:, b0:b0 + 1000])
    np.testing.assert_allclose(pl.features.numpy(), want, rtol=1e-6, atol=0)
    np.testing.assert_array_equal(pl.features.numpy()[:, :70], feats[keep][:, :70])
    np.testing.assert_array_equal(pl.features.numpy()[:, 8070:], feats[keep][:, 8070:])
    raw = formats.segment_pair_list(*got, cls, normalize=False)
    np.testing.assert_array_equal(raw.features.numpy(), feats[keep])
    with pytest.raises(ValueError, match="lacks"):
        bad = str(tmp_path / "bad.npz")
        np.savez(bad, pairs=pairs)
        formats.load_segment(bad)
    with pytest.raises(RuntimeError, match="h5py"):
        formats.load_segment(str(tmp_path / "x-relation.h5"))


@pytest.mark.gpu
def test_segment_through_basemodel_reference_mode(tmp_path):
    """configs/baseline.yaml mode (PPN and DPN off): the segment's rows through RelationPredictor on the GPU."""
    import tspn_b200
    from tspn_b200.model import BaseModel
    path, trackid, pairs, feats, iou, cls = _segment(tmp_path, seed=3)
    pl_gpu = formats.segment_pair_list(*formats.load_segment(path), cls, device="cuda")
    pl_cpu = formats.segment_pair_list(*formats.load_segment(path), cls)
    np.testing.assert_allclose(pl_gpu.features.numpy(), pl_cpu.features.numpy(), rtol=1e-6, atol=0)
    cfg = tspn_b200.get_default_cfg()
    cfg.RELPN.USE_PPN = cfg.RELPN.USE_DPN = False
    cfg.RELPN.DPN.IN_CHANNELS = 8                         # the synthetic checkpoint's span-head width
    sd_np = synth.make_weights(35, 132, 11070, seed=1)
    model = BaseModel(cfg).eval()
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd_np.items()})
    with torch.no_grad():
        pp, dp, logits = model([pl_gpu], None)
    assert pp is None and dp is None and logits[0].shape == (30, 132)
    np.testing.assert_allclose(logits[0].numpy(), oheads.relation_predictor_f64(pl_cpu.features.numpy(), sd_np),
                               rtol=0, atol=2e-6)
