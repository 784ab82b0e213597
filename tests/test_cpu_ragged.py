"""Host logic of the ragged-batch path: capacities, packing, the span-selection oracle, CPU placement.
No GPU: the library is only used for its host-side table builder."""
import numpy as np
import pytest

from oracle import heads as oheads
from tspn_b200 import _lib, affinity, synth
from tspn_b200.batch import Capacity, HostBatch, bucket_capacities, pack_batches, t_class


def test_pack_batches_respects_classes_and_budget():
    shapes = synth.config_shapes("vidor_val", 0, 200)
    batches = pack_batches(shapes, geo_budget_bytes=1 << 30, max_videos=32, merge_below_bytes=0)
    seen = sorted(i for _, vids in batches for i in vids)
    assert seen == list(range(len(shapes)))                              # every video exactly once
    for c, vids in batches:
        assert all(t_class(shapes[i][1]) == c for i in vids) and len(vids) <= 32
        geo = sum(shapes[i][0] * (shapes[i][0] - 1) * ((shapes[i][1] + 3) // 4 * 4) * 32 for i in vids)
        assert geo <= (1 << 30) or len(vids) == 1                        # only a single oversized video may exceed
    assert pack_batches(shapes, 1 << 30, 32, 0) == batches               # deterministic
    # a class with little geometry joins the next larger one (fewer, fuller batches): still every video once,
    # never in a smaller chunk than its own
    merged = pack_batches(shapes[:40], geo_budget_bytes=1 << 30, max_videos=32, merge_below_bytes=1 << 40)
    assert sorted(i for _, vids in merged for i in vids) == list(range(40))
    assert {c for c, _ in merged} == {2048} and all(t_class(shapes[i][1]) <= c for c, vids in merged for i in vids)
    assert [t_class(t) for t in (1, 512, 513, 1024, 1025, 2048, 4096)] == [512, 512, 1024, 1024, 2048, 2048, 2048]


def test_capacity_holds_every_batch_of_its_class():
    shapes = synth.config_shapes("vidvrd_test", 0, 60)
    batches = pack_batches(shapes, geo_budget_bytes=64 << 20, max_videos=8, merge_below_bytes=0)
    caps = bucket_capacities(shapes, batches, 35)
    for c, vids in batches:
        cap = caps[c]
        _, tot = _lib.build_video_table([shapes[i][0] for i in vids], [shapes[i][1] for i in vids], geo_chunk=c)
        assert cap.geo_chunk == c and cap.fits(tot, len(vids))
        assert cap.max_chunks == 1 and cap.max_t == c


def test_host_batch_with_capacity_layout_and_errors():
    c = 35
    vids = [synth.make_video(5, 40, c, seed=1), synth.make_video(3, 70, c, seed=2)]
    cap = Capacity.for_shapes([(9, 100), (6, 100), (4, 30)], c, videos=4)
    a = HostBatch.from_videos(vids, pin=False, capacity=cap)
    b = HostBatch.from_videos(vids[:1], pin=False, capacity=cap)
    assert a.layout == b.layout and a.num_videos == b.num_videos == 4 and (a.num_real, b.num_real) == (2, 1)
    assert a.table.shape == (5, _lib.VT_COLS) and a.n == [5, 3, 0, 0] and list(a.totals) == list(cap.totals())
    assert a.table[4][_lib.VT_PAIR_OFF] == 5 * 4 + 3 * 2 == a.actual[_lib.TOT_PAIRS]
    assert a.h2d_bytes() > b.h2d_bytes() and a.h2d_bytes() < a.arena.numel()
    # the transport dtypes belong to the capacity: data that does not fit raises instead of changing the layout
    frac = synth.make_video(4, 30, c, seed=3, integer_boxes=False)
    with pytest.raises(ValueError, match="u16"):
        HostBatch.from_videos([frac], pin=False, capacity=cap)
    HostBatch.from_videos([frac], pin=False, capacity=cap.grown(boxes_u16=False))
    big = synth.make_video(4, 30, c, seed=4)
    big.motion[0, 0] = 300.0
    with pytest.raises(ValueError, match="u8"):
        HostBatch.from_videos([big], pin=False, capacity=cap)
    with pytest.raises(ValueError, match="does not fit"):
        HostBatch.from_videos([synth.make_video(12, 100, c, seed=5)], pin=False, capacity=cap)


def _brute_select(c, w, n_keep, thr_q10):
    wa, wb = (int(w[0]), int(w[1])) if w[1] > w[0] else (0, 0)
    cand = []
    for i, (s, e) in enumerate(c):
        s, e = int(s), int(e)
        inter = max(0, min(e, wb) - max(s, wa))
        uni = (e - s) + (wb - wa) - inter
        cand.append((-((inter << 15) // uni), i, s, e))
    cand.sort()
    kept = []
    for _, i, s, e in cand:
        if all(not (min(e, ke) - max(s, ks) > 0 and (min(e, ke) - max(s, ks)) * 1024 >
                    thr_q10 * ((e - s) + (ke - ks) - (min(e, ke) - max(s, ks)))) for ks, ke in kept):
            kept.append((s, e))
            if len(kept) == n_keep:
                break
    return kept


@pytest.mark.parametrize("thr", [0.3, 0.5, 0.7, 1.0])
def test_select_spans_oracle_is_greedy_nms(thr):
    """The vectorised oracle against a literal sort-then-scan greedy NMS."""
    rng = np.random.Generator(np.random.PCG64(7))
    r_n, m = 12, 200
    s = rng.integers(0, 900, (r_n, m))
    c = np.stack([s, s + rng.integers(1, 300, (r_n, m))], axis=-1)
    w = np.stack([rng.integers(0, 500, r_n), rng.integers(300, 1100, r_n)], axis=1)
    w[3] = (50, 50)                                                      # an empty window: ranked by index only
    n_cand = rng.integers(1, m + 1, r_n)
    out, cnt = oheads.select_spans(c, w, 16, thr, n_cand=n_cand)
    for r in range(r_n):
        want = _brute_select(c[r, :n_cand[r]], w[r], 16, int(thr * 1024 + 0.5))
        assert cnt[r] == len(want)
        np.testing.assert_array_equal(out[r, :cnt[r]], np.array(want).reshape(-1, 2))
        assert (out[r, cnt[r]:] == 0).all()


def test_cpulist_parsing_and_even_share(monkeypatch):
    assert affinity._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert affinity._parse_cpulist("") == []


# ---------------------------------------------------------------------------------------------------------
# box transport: span-packed u16, per tracklet raw or delta-coded (HostBatch._pack_span_boxes)
# ---------------------------------------------------------------------------------------------------------
def _edge_video():
    """Spans of 0, 1, 2, 3 frames, a span at frame 0, one up to T, a tracklet that jumps by more than 127 pixels
    (raw fallback), one that moves by exactly -128 / +127 (still delta), and coordinates at the u16 limits."""
    from tspn_b200 import synth
    v = synth.make_video(9, 40, 35, seed=5, full_span=True)
    v.span[:] = [(0, 40), (7, 7), (3, 4), (10, 12), (20, 23), (0, 1), (39, 40), (5, 35), (2, 38)]
    v.boxes[7, 20:, 0] += 500.0                       # a jump: this tracklet cannot be delta-coded
    v.boxes[7, 20:, 2] += 500.0
    step = np.where(np.arange(40) % 2 == 0, 127.0, -128.0)
    v.boxes[8, :, 0] = 1000.0 + np.cumsum(step)
    v.boxes[8, :, 2] = 65535.0 - np.arange(40)
    t = np.arange(40)[None, :, None]
    v.boxes[:] = np.where((t >= v.span[:, :1, None]) & (t < v.span[:, 1:, None]), v.boxes, 0.0)
    return v


@pytest.mark.parametrize("delta", [True, False])
def test_box_transport_round_trip_on_the_host(delta):
    from tspn_b200 import _lib, synth
    vids = [synth.make_video(12, 300, 35, seed=1), _edge_video(), synth.make_video(3, 5, 35, seed=2)]
    host = HostBatch.from_videos(vids, pin=False, delta=delta)
    assert host.boxes_compact
    for v, got in zip(vids, host.unpacked_boxes()):
        np.testing.assert_array_equal(got, v.boxes)
    off = host.box_off.numpy()
    is_delta = (off & _lib.PACKED_DELTA) != 0
    if not delta:
        assert not is_delta.any()
        assert host.packed_boxes == sum(int((v.span[:, 1] - v.span[:, 0]).sum()) for v in vids)
    else:
        edge0 = 12                                    # first tracklet of the edge video
        assert is_delta[:12].all()                    # smooth synthetic tracklets
        assert not is_delta[edge0 + 1]                # empty span: nothing to code
        assert not is_delta[edge0 + 7]                # the jump: raw
        assert is_delta[edge0 + 8]                    # -128 / +127 per frame still fits
        lens = np.concatenate([v.span[:, 1] - v.span[:, 0] for v in vids])
        want = np.where(is_delta, 1 + lens // 2, lens).sum()
        assert host.packed_boxes == want
        raw = HostBatch.from_videos(vids, pin=False, delta=False)
        assert host.h2d_bytes() < raw.h2d_bytes()
        assert host.layout == raw.layout              # the arena does not depend on the coding


def test_box_transport_with_a_capacity_and_fractional_boxes():
    from tspn_b200 import synth
    vids = [synth.make_video(6, 120, 35, seed=3), _edge_video()]
    cap = Capacity.for_shapes([(12, 128), (9, 64)], 35, videos=3)
    host = HostBatch.from_videos(vids, pin=False, capacity=cap)
    for v, got in zip(vids, host.unpacked_boxes()):
        np.testing.assert_array_equal(got, v.boxes)
    frac = synth.make_video(4, 50, 35, seed=4, integer_boxes=False)
    host = HostBatch.from_videos([frac], pin=False)
    assert not host.boxes_compact
    np.testing.assert_array_equal(host.unpacked_boxes()[0], frac.boxes)


def _pack_span_boxes_numpy(b, s, trk0, slot0, bview, box_off, delta):
    """The numpy form of ``HostBatch._pack_span_boxes`` (the first implementation; the product runs the same loops in the
    library, ``tspn_host_pack_boxes_spans``): the reference the native packer is compared with, byte for byte."""
    from tspn_b200 import _lib
    n, t = b.shape[0], b.shape[1]
    frame = np.arange(t, dtype=np.int32)[None, :]
    alive = (frame >= s[:, :1]) & (frame < s[:, 1:2])                  # [N, T]
    lens = (s[:, 1] - s[:, 0]).astype(np.int64)
    cnt = int(lens.sum())
    flat = b[alive].astype(np.int32)                                   # [cnt, 4], tracklet-major
    first = np.concatenate([[0], np.cumsum(lens)[:-1]])                # row of every tracklet's first frame
    trk_of = np.repeat(np.arange(n), lens)
    f_in = np.arange(cnt, dtype=np.int64) - first[trk_of]              # frame index inside its tracklet's span
    d = np.zeros_like(flat)
    if cnt > 1:
        d[1:] = flat[1:] - flat[:-1]
    d[f_in == 0] = 0
    is_delta = np.zeros(n, dtype=bool)
    if delta and cnt:
        bad = ((d < -128) | (d > 127)).any(axis=1)
        is_delta = (np.bincount(trk_of, weights=bad, minlength=n) == 0) & (lens > 0)
    slots = np.where(is_delta, 1 + lens // 2, lens)                    # 1 + ceil((L - 1) / 2) = 1 + L // 2
    off = slot0 + np.concatenate([[0], np.cumsum(slots)[:-1]])
    box_off[trk0:trk0 + n] = off | np.where(is_delta, np.int64(_lib.PACKED_DELTA), np.int64(0))
    used = int(slots.sum())
    row_delta = is_delta[trk_of]
    raw = ~row_delta
    bview[(off[trk_of] + f_in)[raw]] = flat[raw]
    head = row_delta & (f_in == 0)
    bview[off[trk_of][head]] = flat[head]
    tail = row_delta & (f_in > 0)
    if tail.any():
        i8 = bview[slot0:slot0 + used].view(np.int8).reshape(-1, 4)    # 4-byte records, two per slot
        i8[((off[trk_of] - slot0 + 1) * 2 + f_in - 1)[tail]] = d[tail].astype(np.int8)
    return used


@pytest.mark.parametrize("delta", [True, False])
def test_native_packer_equals_the_numpy_reference(delta):
    from tspn_b200 import synth
    rng = np.random.default_rng(7)
    vids = [synth.make_video(12, 300, 35, seed=1), _edge_video(), synth.make_video(3, 5, 35, seed=2),
            synth.make_video(20, 1200, 35, seed=9)]
    vids[3].boxes[5, 400:, :] += 200.0                       # one raw tracklet between delta tracklets
    jump = rng.integers(-140, 141, size=(1200, 4)).cumsum(axis=0) + 30000     # steps around the int8 limits
    vids[3].boxes[7] = np.where((np.arange(1200)[:, None] >= vids[3].span[7, 0]) &
                                (np.arange(1200)[:, None] < vids[3].span[7, 1]), jump, 0).astype(np.float32)
    tt = np.arange(1200)[None, :, None]
    vids[3].boxes[:] = np.where((tt >= vids[3].span[:, :1, None]) & (tt < vids[3].span[:, 1:, None]), vids[3].boxes, 0.0)
    host = HostBatch.from_videos(vids, pin=False, delta=delta)
    got = host.boxes.numpy().view(np.uint16)
    want = np.zeros_like(got)
    want_off = np.zeros_like(host.box_off.numpy())
    slot, trk = 0, 0
    for v in vids:
        slot += _pack_span_boxes_numpy(v.boxes, v.span, trk, slot, want, want_off, delta)
        trk += v.n_tracklets
    assert host.packed_boxes == slot
    np.testing.assert_array_equal(host.box_off.numpy(), want_off)
    np.testing.assert_array_equal(got, want)
    for v, dense in zip(vids, host.unpacked_boxes()):
        np.testing.assert_array_equal(dense, v.boxes)


def test_native_packer_rejects_what_it_cannot_ship():
    import ctypes
    from tspn_b200 import _lib
    lib = _lib.load()
    boxes = np.zeros((2, 6, 4), dtype=np.float32)
    boxes[:, :, 2:] = 10.0
    span = np.array([[0, 6], [2, 5]], dtype=np.int32)
    dst = np.zeros((9, 4), dtype=np.uint16)
    off = np.zeros(2, dtype=np.int64)
    used = ctypes.c_int64(-1)

    def pack(b, s, slots=9, slot0=0):
        return lib.tspn_host_pack_boxes_spans(b.ctypes.data, 2, 6, s.ctypes.data, 0, dst.ctypes.data, slots, slot0,
                                              off.ctypes.data, ctypes.addressof(used))
    assert pack(boxes, span) == _lib.TSPN_OK and used.value == 9 and off.tolist() == [0, 6]
    assert pack(boxes, span, slots=8) == _lib.TSPN_ESHAPE and "exceed the arena" in _lib.last_error()
    bad = boxes.copy()
    bad[1, 3, 0] = 0.5                                       # inside tracklet 1's span
    assert pack(bad, span) == _lib.TSPN_ESHAPE and "fractional" in _lib.last_error()
    bad[1, 3, 0] = 0.0
    bad[1, 0, 0] = 0.5                                       # outside its span: never read, never shipped
    assert pack(bad, span) == _lib.TSPN_OK
    assert pack(boxes, np.array([[0, 7], [2, 5]], dtype=np.int32)) == _lib.TSPN_ESHAPE


@pytest.mark.parametrize("seed", range(6))
def test_native_packer_on_random_tracklets(seed):
    """Random spans (empty, one frame, whole video), random walks whose steps straddle the int8 limits, coordinates at
    both ends of the u16 range: native packer = numpy reference byte for byte, and the host decode returns the boxes."""
    rng = np.random.default_rng(100 + seed)
    n, t = int(rng.integers(1, 9)), int(rng.integers(1, 70))
    span = np.sort(rng.integers(0, t + 1, size=(n, 2)), axis=1).astype(np.int32)
    span[rng.integers(0, n)] = (0, t)
    step_hi = int(rng.choice([3, 127, 128, 129, 400]))
    walk = rng.integers(-step_hi, step_hi + 1, size=(n, t, 4)).cumsum(axis=1)
    boxes = np.clip(walk + rng.choice([0, 300, 65535 - 300]), 0, 65535).astype(np.float32)
    tt = np.arange(t)[None, :, None]
    boxes = np.where((tt >= span[:, :1, None]) & (tt < span[:, 1:, None]), boxes, 0.0).astype(np.float32)
    for delta in (True, False):
        host = HostBatch([boxes], [span], pin=False, delta=delta)
        assert host.boxes_compact
        got = host.boxes.numpy().view(np.uint16)
        want, want_off = np.zeros_like(got), np.zeros_like(host.box_off.numpy())
        used = _pack_span_boxes_numpy(boxes, span, 0, 0, want, want_off, delta)
        assert host.packed_boxes == used
        np.testing.assert_array_equal(host.box_off.numpy(), want_off)
        np.testing.assert_array_equal(got[:used], want[:used])
        np.testing.assert_array_equal(host.unpacked_boxes()[0], boxes)
