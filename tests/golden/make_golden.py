"""Generate tests/golden/reference_outputs.npz by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (sangminwoo/Temporal-Span-Proposal-Network-VidVRD) has no tests and no
golden vectors, but it is Python, so its arithmetic functions on the pair-stage path are
imported live here and executed on seeded synthetic inputs (``tspn_b200.synth``); only
their OUTPUTS are stored — inputs are regenerated from the seeds by the tests.  Import
stubs are needed for two absent third-party modules that contribute no arithmetic:
``dlib`` (only the ``drectangle`` accessor type, lib/modeling/trajectory.py:1) and
``IPython`` (``embed``, pulled in by lib/evaluation/visual_relation_detection.py:5).

Functions executed: ``cubic_iou``/``traj_iou`` (lib/modeling/trajectory.py:127-158),
``viou`` (lib/evaluation/common.py:65-106), ``_traj_iou`` (lib/modeling/association.py:35-48),
``normalize`` (lib/utils/miscellaneous.py:32-35), ``PPNHead.forward``/``PPN._forward_test``
(lib/modeling/relpn/ppn.py:79-112), ``RelationPredictor.forward`` (lib/modeling/model.py:76-88),
``DPNHead.forward`` (lib/modeling/relpn/dpn.py:55-73), ``BaseModel.forward`` in eval mode
(lib/modeling/model.py:53-65).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TSPN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


def _install_stubs():
    dl = types.ModuleType("dlib")

    class drectangle:  # accessor type only (dlib 19.x API used at trajectory.py:148-155)
        def __init__(self, l, t, r, b):
            self._v = (l, t, r, b)

        def left(self):
            return self._v[0]

        def top(self):
            return self._v[1]

        def right(self):
            return self._v[2]

        def bottom(self):
            return self._v[3]

    dl.drectangle = drectangle
    dl.correlation_tracker = object
    sys.modules["dlib"] = dl
    ip = types.ModuleType("IPython")
    ip.embed = lambda *a, **k: None
    sys.modules["IPython"] = ip


def _ns(**kw):
    return types.SimpleNamespace(**kw)


def make_cfg(use_ppn, use_dpn, classes=35, predicates=132, fdim=11070, topk=256, dpn_in=8, anchors=4):
    return _ns(
        RELPN=_ns(USE_PPN=use_ppn, USE_DPN=use_dpn, OBJECT_DIM=1024,
                  PPN=_ns(NUM_PAIR_PROPOSALS=topk, IN_CHANNELS=classes, HIDDEN_CHANNELS=64,
                          OUT_CHANNELS=classes, BATCH_SIZE_PER_SEGMENT=256, POSITIVE_FRACTION=0.25),
                  DPN=_ns(NUM_DURATION_PROPOSALS=64, DPN_ONLY=False, IN_CHANNELS=dpn_in,
                          NUM_ANCHORS_PER_LOCATION=anchors, ANCHOR_SIZES=35, ANCHOR_STRIDE=132)),
        PREDICT=_ns(OBJECT_NUM=classes, PREDICATE_NUM=predicates, TOPK_PER_PAIR=20, TOPK_PER_SEG=200,
                    FEATURE_DIM=fdim))


def synth_features(p, f, seed):
    """Seeded stand-in for the h5 ``feats`` rows: sparse non-negative values."""
    rng = np.random.Generator(np.random.PCG64(seed + 31337))
    x = rng.random((p, f), dtype=np.float32)
    keep = rng.random((p, f), dtype=np.float32) < 0.1
    return np.where(keep, x, np.float32(0)).astype(np.float32)


def main():
    _install_stubs()
    import torch
    from lib.modeling.trajectory import cubic_iou, traj_iou, Trajectory
    from lib.evaluation.common import viou
    from lib.modeling.association import _traj_iou
    from lib.utils.miscellaneous import normalize
    from lib.modeling.model import BaseModel, RelationPredictor
    from lib.modeling.relpn.ppn import PPNHead
    from lib.modeling.relpn.dpn import DPNHead
    from lib.dataset.list_pair import PairList
    from tspn_b200 import synth
    from oracle.geometry import enumerate_pairs

    torch.set_num_threads(1)   # one summation order for the stored torch outputs
    out = {}

    # ---- geometry: config A (N=20, T=300), full spans for V1; ragged small case ------------
    for tag, (n, t, seed) in {"A": (20, 300, 0), "S": (5, 37, 11)}.items():
        v = synth.make_video(n, t, 35, seed=seed, full_span=True)
        out[f"cubic_iou_f32_{tag}"] = cubic_iou(v.boxes, v.boxes)
        b64 = v.boxes.astype(np.float64)
        out[f"cubic_iou_f64in_{tag}"] = cubic_iou(b64, b64)
        half = v.boxes[: n // 2].copy()
        rest = v.boxes[n // 2:].copy()
        out[f"cubic_iou_cross_{tag}"] = cubic_iou(half, rest)
        trajs = [Trajectory(0, t, [tuple(map(float, r)) for r in v.boxes[i]], 1.0, 0, None) for i in range(min(n, 6))]
        out[f"traj_iou_{tag}"] = traj_iou(trajs, trajs)

    # ---- V2 viou / V3 _traj_iou on tracklets with different spans --------------------------
    v = synth.make_video(20, 300, 35, seed=3, full_span=False)
    pairs = enumerate_pairs(20)
    vi = np.zeros(pairs.shape[0], dtype=np.float64)
    v3 = np.full(pairs.shape[0], np.nan, dtype=np.float64)
    lists = [[tuple(int(c) for c in r) for r in v.boxes[i, v.span[i, 0]:v.span[i, 1]]] for i in range(20)]
    trs = [Trajectory(int(v.span[i, 0]), int(v.span[i, 1]), [tuple(map(float, r)) for r in lists[i]], 1.0, 0, None)
           for i in range(20)]
    for r, (s, o) in enumerate(pairs):
        vi[r] = viou(lists[s], tuple(v.span[s]), lists[o], tuple(v.span[o]))
        a, b = (s, o) if v.span[s, 0] <= v.span[o, 0] else (o, s)
        # _traj_iou assumes the earlier-starting trajectory also ends first (association.py:44-46)
        if v.span[a, 1] <= v.span[b, 1] or v.span[a, 1] <= v.span[b, 0]:
            v3[r] = float(_traj_iou(trs[s], trs[o]))
    out["viou_v2_seed3"] = vi
    out["traj_iou_v3_seed3"] = v3

    # ---- normalize (L1 over the BoW blocks) -------------------------------------------------
    m = synth.make_video(6, 10, 35, seed=5).motion
    m[2, :1000] = 0
    out["normalize_l1_seed5"] = np.concatenate(
        [normalize(m[:, k * 1000:(k + 1) * 1000], axis=-1, order=1) for k in range(4)], axis=1)

    # ---- heads ------------------------------------------------------------------------------
    for tag, (n, t, c, r, seed) in {"A": (20, 300, 35, 132, 0), "V": (12, 64, 80, 50, 2)}.items():
        fdim = synth.feature_dim(c)
        sd_np = synth.make_weights(c, r, fdim, dpn_in=8, n_anchors=4, seed=seed)
        sd = {k: torch.from_numpy(vv) for k, vv in sd_np.items()}
        vid = synth.make_video(n, t, c, seed=seed)
        p = n * (n - 1)
        feats = synth_features(p, fdim, seed)

        head = PPNHead(c, 64, c)
        head.load_state_dict({k.split("ppn_head.")[1]: vv for k, vv in sd.items() if "ppn_head" in k})
        with torch.no_grad():
            out[f"ppn_scores_{tag}"] = head(torch.from_numpy(vid.cls), torch.from_numpy(vid.cls)).numpy()

        clf = RelationPredictor(fdim, r)
        clf.load_state_dict({k.split("classifier.")[1]: vv for k, vv in sd.items() if k.startswith("classifier.")})
        with torch.no_grad():
            out[f"rel_logits_{tag}"] = clf(torch.from_numpy(feats)).numpy()

        dpn = DPNHead(8, 4)
        dpn.load_state_dict({k.split("dpn_head.")[1]: vv for k, vv in sd.items() if "dpn_head" in k})
        rng = np.random.Generator(np.random.PCG64(seed + 77))
        x = rng.normal(0, 1, size=(6, 8, t)).astype(np.float32)
        with torch.no_grad():
            out[f"dpn_reg_{tag}"] = dpn(torch.from_numpy(x)).numpy()

        # BaseModel eval forward, baseline.yaml flags (PPN off, DPN off) and PPN on
        plist = PairList(torch.from_numpy(feats))
        plist.add_field("tracklet_pairs", torch.from_numpy(enumerate_pairs(n)))
        plist.add_field("track_cls_logits", torch.from_numpy(vid.cls))
        plist.add_field("num_tracklets", n)
        for flags in ((False, False), (True, False)):
            model = BaseModel(make_cfg(flags[0], flags[1], c, r, fdim, 256, 8, 4))
            model.load_state_dict(sd)
            model.eval()
            with torch.no_grad():
                pp, dp, logits = model([plist], None)
            assert dp is None
            key = f"basemodel_{tag}_ppn{int(flags[0])}"
            out[key + "_logits"] = logits[0].numpy()
            if pp is not None:
                out[key + "_proposals"] = pp[0].numpy()

    # DPNHead at a tensor-core-sized channel count
    sd_np = synth.make_weights(35, 132, 16, dpn_in=64, n_anchors=4, seed=9)
    dpn = DPNHead(64, 4)
    dpn.load_state_dict({k.split("dpn_head.")[1]: torch.from_numpy(vv) for k, vv in sd_np.items() if "dpn_head" in k})
    rng = np.random.Generator(np.random.PCG64(9 + 77))
    x = rng.normal(0, 1, size=(3, 64, 50)).astype(np.float32)
    with torch.no_grad():
        out["dpn_reg_C64"] = dpn(torch.from_numpy(x)).numpy()

    # DPN.forward really is broken in the reference (quirk Q5) — record the fact.
    try:
        BaseModel(make_cfg(True, True, c, r, fdim, 256, 8, 4)).eval()([plist], None)
        out["dpn_forward_error"] = np.array("none")
    except Exception as e:  # noqa: BLE001
        out["dpn_forward_error"] = np.array(type(e).__name__)

    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(vv, "shape", None) for k, vv in out.items()})
    print("size", os.path.getsize(path))


def main_headline():
    """The same functions of the unmodified reference at the HEADLINE shape (BASELINE.json configs[2]: N=64, T=2000,
    80 classes, 50 predicates, F=11160) -> reference_outputs_headline.npz (kept in its own file so that
    reference_outputs.npz regenerates bit for bit)."""
    _install_stubs()
    import torch
    from lib.modeling.trajectory import cubic_iou
    from lib.evaluation.common import viou
    from lib.modeling.model import RelationPredictor
    from lib.modeling.relpn.ppn import PPNHead
    from tspn_b200 import synth
    from oracle.geometry import enumerate_pairs

    torch.set_num_threads(1)
    n, t, c, r, seed = 64, 2000, 80, 50, 0
    out = {}
    v = synth.make_video(n, t, c, seed=seed, full_span=True)
    out["cubic_iou_f32_C"] = cubic_iou(v.boxes, v.boxes)                      # [64, 64], V1
    vr = synth.make_video(n, t, c, seed=seed + 1)                             # ragged spans: V2 on a sample of pairs
    pairs = enumerate_pairs(n)
    sel = np.sort(np.random.Generator(np.random.PCG64(5)).choice(pairs.shape[0], size=96, replace=False))
    lists = [[tuple(int(q) for q in row) for row in vr.boxes[i, vr.span[i, 0]:vr.span[i, 1]]] for i in range(n)]
    out["viou_v2_C_rows"] = sel.astype(np.int64)
    out["viou_v2_C"] = np.array([viou(lists[s], tuple(vr.span[s]), lists[o], tuple(vr.span[o])) for s, o in pairs[sel]])
    fdim = synth.feature_dim(c)
    sd = {k: torch.from_numpy(vv) for k, vv in synth.make_weights(c, r, fdim, dpn_in=8, n_anchors=4, seed=seed).items()}
    head = PPNHead(c, 64, c)
    head.load_state_dict({k.split("ppn_head.")[1]: vv for k, vv in sd.items() if "ppn_head" in k})
    with torch.no_grad():
        out["ppn_scores_C"] = head(torch.from_numpy(v.cls), torch.from_numpy(v.cls)).numpy()
    clf = RelationPredictor(fdim, r)
    clf.load_state_dict({k.split("classifier.")[1]: vv for k, vv in sd.items() if k.startswith("classifier.")})
    with torch.no_grad():
        out["rel_logits_C"] = clf(torch.from_numpy(synth_features(n * (n - 1), fdim, seed))).numpy()    # [4032, 50]
    path = os.path.join(HERE, "reference_outputs_headline.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(vv, "shape", None) for k, vv in out.items()}, "size", os.path.getsize(path))


if __name__ == "__main__":
    if "--headline" in sys.argv:
        main_headline()
    else:
        main()
