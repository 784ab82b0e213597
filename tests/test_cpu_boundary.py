"""CPU-side checks: the C-ABI library loads and exports what include/tspn_b200.h declares, host
logic (batch layout, config, PairList), and the no-fallback rule."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import tspn_b200
from tspn_b200 import _lib, synth
from tspn_b200.batch import HostBatch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "tspn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tspn_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from tspn_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 20
    assert sorted(_lib.SIGNATURES) == declared          # binding table == header
    exported = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    for name in declared:
        assert re.search(r"\bT %s\b" % name, exported), name
        getattr(lib, name)
    assert lib.tspn_version() == _lib.ABI_VERSION == 6


def test_sass_is_blackwell_native():
    """TMA (UTMALDG) in the geometry kernel, tcgen05 (UTC*MMA, LDTM) in the tensor heads."""
    sass = subprocess.check_output(["cuobjdump", "-sass", _lib.LIB_PATH], text=True)
    assert "UTMALDG" in sass and "UTCHMMA" in sass and "LDTM" in sass
    assert "HMMA.16816" not in sass                     # no legacy mma.sync path
    arch = subprocess.check_output(["cuobjdump", "-lelf", _lib.LIB_PATH], text=True)
    assert "sm_100a" in arch


def test_video_table_layout():
    n, t = [20, 1, 0, 5, 2, 40], [300, 10, 7, 37, 1, 1200]
    table, tot = _lib.build_video_table(n, t)
    assert table.shape == (6 + 1, _lib.VT_COLS)             # the videos + the sentinel row
    trk = pairs = geo = boxes = scores = items = 0
    for v in range(6):
        row = table[v]
        tp, tb = (t[v] + 3) // 4 * 4, (t[v] + 7) // 8 * 8
        assert list(row[:4]) == [n[v], t[v], tp, tb]
        assert row[_lib.VT_TRK_OFF] == trk and row[_lib.VT_PAIR_OFF] == pairs and row[_lib.VT_GEO_OFF] == geo
        assert row[_lib.VT_BOX_OFF] == boxes and row[_lib.VT_SCORE_OFF] == scores and row[_lib.VT_ITEM_OFF] == items
        p = n[v] * max(n[v] - 1, 0)
        trk, pairs, geo, boxes, scores = trk + n[v], pairs + p, geo + p * 8 * tp, boxes + n[v] * tb, scores + n[v] ** 2
        chunk = 2048                                         # tspn_geo_chunk(max T = 1200)
        groups, chunks = -(-(n[v] - 1) // _lib.GEO_OBJ_GROUP), -(-t[v] // chunk)
        items += n[v] * groups * chunks if n[v] >= 2 else 0
    assert list(tot[:6]) == [trk, pairs, geo, items, boxes, scores] and tot[6] == 40 and tot[7] == 1200 and tot[8] == 2048
    # the sentinel carries the totals in the offset columns (what the kernels read as the batch's true sizes)
    sent = table[6]
    assert [sent[c] for c in (_lib.VT_TRK_OFF, _lib.VT_PAIR_OFF, _lib.VT_GEO_OFF, _lib.VT_ITEM_OFF, _lib.VT_BOX_OFF,
                              _lib.VT_SCORE_OFF)] == [trk, pairs, geo, items, boxes, scores]
    assert sent[_lib.VT_N] == 0 and tot[_lib.TOT_MAX_CHUNKS] == 1
    # padded to a capacity of 9 table rows with a fixed chunk: empty videos, then the sentinel; same offsets
    padded, tot_p = _lib.build_video_table(n, t, table_rows=9, geo_chunk=2048)
    assert padded.shape == (10, _lib.VT_COLS) and np.array_equal(padded[:6], table[:6]) and np.array_equal(tot_p, tot)
    assert all(padded[v][_lib.VT_N] == 0 and padded[v][_lib.VT_PAIR_OFF] == pairs for v in (6, 7, 8, 9))
    # a smaller chunk than the longest video: more work items and chunk slots, same everything else
    _, tot_c = _lib.build_video_table(n, t, geo_chunk=512)
    assert tot_c[_lib.TOT_GEO_CHUNK] == 512 and tot_c[_lib.TOT_MAX_CHUNKS] == 3 and tot_c[_lib.TOT_ITEMS] > items
    assert _lib.build_video_table([256], [4096])[1][_lib.TOT_MAX_CHUNKS] == 2
    with pytest.raises(RuntimeError, match="TSPN_EBADARG"):
        _lib.build_video_table(n, t, table_rows=3)
    assert [_lib.load().tspn_geo_chunk(x) for x in (1, 512, 513, 1024, 1025, 5000)] == [512, 512, 1024, 1024, 2048, 2048]
    assert _lib.build_video_table([70, 3], [300, 20])[1][_lib.TOT_ITEMS] == 70 * 3 + 3      # three groups of <= 32
    assert boxes % 8 == 0
    with pytest.raises(RuntimeError, match="TSPN_ESHAPE"):
        _lib.build_video_table([3], [0])


def test_host_batch_packing():
    vids = [synth.make_video(4, 10, 35, seed=1), synth.make_video(0, 3, 35, seed=2), synth.make_video(3, 17, 35, seed=3)]
    hb = HostBatch.from_videos(vids, pin=False, compact=False)
    assert hb.boxes.shape == (4 * 16 + 3 * 24, 4) and hb.span.shape == (7, 2)
    b = hb.boxes.numpy()
    np.testing.assert_array_equal(b[:64].reshape(4, 16, 4)[:, :10], vids[0].boxes)
    np.testing.assert_array_equal(b[:64].reshape(4, 16, 4)[:, 10:], 0)
    np.testing.assert_array_equal(b[64:].reshape(3, 24, 4)[:, :17], vids[2].boxes)
    np.testing.assert_array_equal(hb.cls.numpy()[4:], vids[2].cls)
    assert hb.h2d_bytes() > 0
    # compact transport: integer boxes as u16, integer motion counts as u8 - lossless, fewer bytes
    hc = HostBatch.from_videos(vids, pin=False, delta=False)
    assert hc.boxes_compact and hc.motion_compact and hc.h2d_bytes() < hb.h2d_bytes() // 2
    # ... and span-packed: only the frames [pstart, pend) of every tracklet travel, tracklet after tracklet
    # (raw coding here; the delta coding is covered by tests/test_cpu_ragged.py's round trips)
    packed = hc.boxes.numpy().view(np.uint16).astype(np.float32)
    off = hc.box_off.numpy()
    trk = 0
    for v in (vids[0], vids[2]):
        for n in range(v.n_tracklets):
            ps, pe = v.span[n]
            np.testing.assert_array_equal(packed[off[trk]:off[trk] + pe - ps], v.boxes[n, ps:pe])
            trk += 1
    assert hc.packed_boxes == sum(int((v.span[:, 1] - v.span[:, 0]).sum()) for v in vids) < b.shape[0]
    np.testing.assert_array_equal(hc.motion.numpy().astype(np.float32), hb.motion.numpy())
    np.testing.assert_array_equal(hc.cls.numpy(), hb.cls.numpy())
    # values that are not exactly representable travel as fp32
    frac = synth.make_video(3, 9, 35, seed=4, integer_boxes=False)
    frac.motion[0, 0] = 0.5
    hf = HostBatch.from_videos([frac], pin=False)
    assert not hf.boxes_compact and not hf.motion_compact and hf.boxes.dtype == torch.float32
    with pytest.raises(ValueError):
        HostBatch([np.zeros((2, 5, 4), np.float32)], [np.array([[0, 6], [0, 5]], np.int32)], pin=False)


def test_config_matches_reference_defaults_and_yaml():
    cfg = tspn_b200.get_default_cfg()
    assert cfg.PREDICT.OBJECT_NUM == 35 and cfg.PREDICT.PREDICATE_NUM == 132 and cfg.PREDICT.FEATURE_DIM == 11070
    assert cfg.RELPN.PPN.NUM_PAIR_PROPOSALS == 256 and cfg.RELPN.PPN.HIDDEN_CHANNELS == 64
    assert cfg.RELPN.DPN.NUM_ANCHORS_PER_LOCATION == 4 and cfg.PREDICT.TOPK_PER_SEG == 200
    cfg.merge_from_file(os.path.join(ROOT, "configs", "baseline.yaml"))
    assert cfg.RELPN.USE_PPN is False and cfg.RELPN.USE_DPN is False          # configs/baseline.yaml:15-17
    assert cfg.SOLVER.BASE_LR == pytest.approx(1e-2) and cfg.RELPN.PPN.POSITIVE_FRACTION == 0.25
    with pytest.raises(KeyError):
        cfg.merge_from_dict({"PREDICT": {"NOPE": 1}})
    vid = tspn_b200.get_default_cfg()
    vid.merge_from_file(os.path.join(ROOT, "configs", "vidor.yaml"))
    assert vid.PREDICT.OBJECT_NUM == 80 and vid.PREDICT.PREDICATE_NUM == 50 and vid.PREDICT.FEATURE_DIM == 11160


def test_pairlist_api():
    from tspn_b200.list_pair import PairList
    f = torch.arange(12.0).view(6, 2)
    pl = PairList(f)
    pl.add_field("tracklet_pairs", torch.arange(12).view(6, 2))
    assert len(pl) == 6 and pl.has_field("tracklet_pairs") and pl.fields() == ["tracklet_pairs"]
    sub = pl[torch.tensor([0, 2])]
    assert len(sub) == 2 and sub.get_field("tracklet_pairs").shape == (2, 2)
    assert pl.to("cpu").features.equal(f)
    assert len(pl.copy_with_fields("tracklet_pairs")) == 6
    with pytest.raises(KeyError):
        pl.copy_with_fields("missing")
    v = synth.make_video(3, 5, 35, seed=0)
    t = PairList.from_tracklets(v.boxes, v.span, v.cls, v.motion)
    assert t.has_tracklets() and len(t) == 6 and t.features is None


def test_state_dict_keys_are_the_references():
    from tspn_b200.model import BaseModel
    cfg = tspn_b200.get_default_cfg()
    keys = list(BaseModel(cfg).state_dict().keys())
    want = [f"relpn.pair_proposal_network.ppn_head.{b}.{i}.{p}" for b in ("sub_emb", "obj_emb") for i in (0, 2)
            for p in ("weight", "bias")]
    want += [f"relpn.duration_proposal_network.dpn_head.{m}.{p}" for m in ("conv", "duration_pred")
             for p in ("weight", "bias")]
    want += ["classifier.rel_predictor.weight", "classifier.rel_predictor.bias"]
    assert keys == want
    sd = BaseModel(cfg).state_dict()
    assert sd["classifier.rel_predictor.weight"].shape == (132, 11070)
    assert sd["relpn.duration_proposal_network.dpn_head.conv.weight"].shape == (1024, 1024, 3)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    """Without a B200 the product path raises; it never computes on the CPU and never touches oracle/."""
    from tspn_b200.model import BaseModel, RelationPredictor
    from tspn_b200 import ops, trajectory
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        RelationPredictor(16, 4).eval()(torch.zeros(2, 16))
    with pytest.raises(RuntimeError):
        trajectory.cubic_iou(np.zeros((2, 3, 4), np.float32), np.zeros((2, 3, 4), np.float32))
    with pytest.raises(RuntimeError):
        ops.require_device()
    src_dir = os.path.join(ROOT, "temporal-span-proposal-network-vidvrd_b200")
    for dirpath, _, files in os.walk(src_dir):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
