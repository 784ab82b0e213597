"""ctypes binding of csrc/libtspn_b200.so (the C ABI of include/tspn_b200.h).

No fallback: if the library is missing ``load()`` raises, and on a non-sm_100 device every
compute entry returns ``TSPN_EARCH`` which ``check()`` turns into ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libtspn_b200.so")

TSPN_OK, TSPN_EBADARG, TSPN_ESHAPE, TSPN_EALIGN, TSPN_ECUDA, TSPN_EARCH = 0, -1, -2, -3, -4, -5
ERROR_NAMES = {-1: "TSPN_EBADARG", -2: "TSPN_ESHAPE", -3: "TSPN_EALIGN", -4: "TSPN_ECUDA", -5: "TSPN_EARCH"}

VT_COLS = 12
VT_N, VT_T, VT_TP, VT_TB, VT_TRK_OFF, VT_PAIR_OFF, VT_GEO_OFF, VT_ITEM_OFF, VT_BOX_OFF, VT_SCORE_OFF = range(10)
TOT_COLS = 10
(TOT_TRACKLETS, TOT_PAIRS, TOT_GEO_FLOATS, TOT_ITEMS, TOT_BOXES, TOT_SCORES, TOT_MAX_N, TOT_MAX_T, TOT_GEO_CHUNK,
 TOT_MAX_CHUNKS) = range(10)

ABI_VERSION = 6
GEO_OBJ_GROUP = 32        # include/tspn_b200.h: objects per work item of the pair-geometry kernel
GEO_CHANNELS = 8
MOTION_DIM = 4000
REL_DIM = 3000
VIOU_FULL, VIOU_CLIPPED = 0, 1
GEO_DENSE_CTAS = 2
GEO_PHASE_PRE, GEO_PHASE_MAIN, GEO_PHASE_POST = 8, 16, 32
GEO_PERSISTENT = 128
GEO_RESERVE_SHIFT = 16
PACKED_DELTA = 1 << 62     # include/tspn_b200.h TSPN_PACKED_DELTA: box_off bit of a delta-coded tracklet
TOPK_KEEP_DIAGONAL, TOPK_EXCLUDE_DIAGONAL = 0, 1
PREC_FP32_EXACT, PREC_TENSOR = 0, 1
AFFINE_RAW = 1
AFFINE_BACKGROUND = 2
SPANS_I16 = 1

P = c_void_p      # device pointers travel as integers (tensor.data_ptr())

# name -> (restype, argtypes); the single source of truth for the binding AND for the
# "library exports every declared symbol" test.
SIGNATURES = {
    "tspn_version": (c_int, []),
    "tspn_last_error": (c_int, [c_char_p, c_int]),
    "tspn_check_device": (c_int, []),
    "tspn_build_video_table": (c_int, [c_int, POINTER(c_int32), POINTER(c_int32), c_int, c_int, POINTER(c_int64),
                                       POINTER(c_int64)]),
    "tspn_enumerate_pairs": (c_int, [P, c_int, c_int64, P, P]),
    "tspn_pair_geo_workspace_bytes": (c_int64, [c_int64, c_int64, c_int]),
    "tspn_geo_chunk": (c_int, [c_int64]),
    "tspn_pair_geo_viou": (c_int, [P, c_int, c_int64, c_int, c_int, c_int64, c_int64, c_int64, P, P, P, P, P, P, c_int,
                                   P, P]),
    "tspn_geo_window_offsets": (c_int, [P, c_int, c_int64, P, P, P, P]),
    "tspn_pair_geo_viou_windowed": (c_int, [P, c_int, c_int64, c_int, c_int, c_int64, c_int64, c_int64, P, P, P, P, P, P,
                                            P, c_int, P, P]),
    "tspn_cubic_iou": (c_int, [P, c_int, P, c_int, c_int, P, P]),
    "tspn_viou_pairs": (c_int, [P, P, P, P, P, c_int64, c_int, P, P]),
    "tspn_viou_pairs_workspace_bytes": (c_int64, [c_int64]),
    "tspn_viou_pairs_f64": (c_int, [P, P, P, P, c_int64, P, P, c_int64, c_int, P, P, P]),
    "tspn_normalize_motion": (c_int, [P, c_int64, P, P]),
    "tspn_normalize_motion_u8": (c_int, [P, c_int64, P, P]),
    "tspn_unpack_boxes_u16": (c_int, [P, c_int64, P, P]),
    "tspn_unpack_boxes_spans": (c_int, [P, c_int, c_int64, P, P, P, P, P]),
    "tspn_host_pack_boxes_spans": (c_int, [P, c_int, c_int, P, c_int, P, c_int64, c_int64, P, P]),
    "tspn_assemble_features": (c_int, [P, c_int, c_int64, c_int, P, c_int, P, P, P, P, c_int64, P, c_int64, P, c_int64, P]),
    "tspn_relationness_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "tspn_relationness_tc_supported": (c_int, [c_int, c_int, c_int]),
    "tspn_relationness": (c_int, [P, c_int, c_int64, c_int, P, c_int, c_int, P, P, P, P, P, P, P, P, P, c_int, P, P]),
    "tspn_topk_pairs": (c_int, [P, c_int, P, c_int, c_int, P, P, P, P]),
    "tspn_relationness_topk_supported": (c_int, [c_int, c_int]),
    "tspn_relationness_topk": (c_int, [P, c_int, c_int64, c_int, P, c_int, c_int, P, P, P, P, P, P, P, P, P, c_int, c_int,
                                       c_int, P, P, P, P, P]),
    "tspn_predicate_packed_bytes": (c_int64, [c_int, c_int]),
    "tspn_pack_predicate_weights": (c_int, [P, c_int, c_int, P, P]),
    "tspn_predicate_workspace_bytes": (c_int64, [c_int64, c_int, c_int, c_int]),
    "tspn_predicate_head": (c_int, [P, c_int, c_int64, c_int64, c_int, P, P, P, c_int, P, c_int, P, P]),
    "tspn_tracklet_rows": (c_int, [P, c_int, P, c_int, c_int64, P, c_int64, P]),
    "tspn_predicate_head_affine": (c_int, [P, c_int, c_int64, c_int64, c_int, P, P, P, c_int64, c_int, P, c_int, P, P]),
    "tspn_assemble_relative": (c_int, [P, c_int, c_int64, c_int, P, P, P, c_int64, P, c_int64, P, P, c_int, P, P]),
    "tspn_span_head_workspace_bytes": (c_int64, [c_int64, c_int, c_int, c_int, c_int]),
    "tspn_span_head": (c_int, [P, P, c_int64, c_int64, c_int64, c_int64, c_int, c_int, P, P, P, P, c_int, P, c_int, P, P]),
    "tspn_span_num_locations": (c_int, [c_int, c_float]),
    "tspn_span_decode": (c_int, [P, c_int64, c_int, c_int, P, c_float, P, P]),
    "tspn_span_proposals": (c_int, [P, P, c_int64, c_int64, c_int64, c_int64, c_int, c_int, P, P, P, P, c_int, P,
                                    c_float, P, P]),
    "tspn_span_select": (c_int, [P, c_int, c_int, P, P, c_int64, P, P, c_int64, c_int, c_int, c_float, c_int, c_float,
                                 c_int, P, P, P]),
    "tspn_postprocess_workspace_bytes": (c_int64, [c_int64, c_int]),
    "tspn_postprocess": (c_int, [P, c_int, P, P, P, c_int64, c_int, P, c_int, P, P, c_int, c_int, c_int, P, P, P, P]),
    "tspn_records_scatter": (c_int, [P, c_int64, P, P, P, c_int, c_int, c_int, c_int, P, c_int, P]),
    "tspn_records_wait": (c_int, [P, c_int, c_int, c_int, P, c_int, P]),
    "tspn_records_release": (c_int, [P, c_int, c_int, c_int, c_int, P]),
    "tspn_survivor_rows_supported": (c_int, [c_int, c_int]),
    "tspn_gather_pair_terms": (c_int, [P, c_int, P, c_int64, P, P, c_int, P, P]),
    "tspn_survivor_rows": (c_int, [P, c_int, c_int, P, P, P, c_int64, c_int64, P, c_int64, P, P, c_int, P, P, P, P, P,
                                   c_int, P, c_float, P, c_int64, P]),
}

_lib = None


def load():
    """Load the shared library (built by ``tspn_b200.build``); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libtspn_b200.so not found at %s — build it with `python -m tspn_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.tspn_version() != ABI_VERSION:
        raise RuntimeError("libtspn_b200.so ABI version %d, expected %d" % (lib.tspn_version(), ABI_VERSION))
    _lib = lib
    return lib


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    load().tspn_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != TSPN_OK:
        raise RuntimeError("%s failed: %s: %s" % (what or "libtspn_b200", ERROR_NAMES.get(rc, rc), last_error()))


def ptr(t) -> int:
    """data_ptr of a torch tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def build_video_table(n_tracklets, n_frames, table_rows: int = 0, geo_chunk: int = 0):
    """Host-side batch layout: returns (table int64 [table_rows + 1, VT_COLS] numpy, totals int64 [TOT_COLS]).

    ``table_rows`` (0 = number of videos) pads the table with empty videos; the last row is the sentinel
    that carries the batch's true totals to the device (include/tspn_b200.h).  ``geo_chunk`` (0 = from the
    longest video) fixes the pair kernel's chunk, as a capacity bucket does."""
    import numpy as np
    n = np.ascontiguousarray(n_tracklets, dtype=np.int32)
    t = np.ascontiguousarray(n_frames, dtype=np.int32)
    v = int(n.shape[0])
    rows = int(table_rows) if table_rows else v
    table = np.zeros((rows + 1, VT_COLS), dtype=np.int64)
    totals = np.zeros(TOT_COLS, dtype=np.int64)
    rc = load().tspn_build_video_table(
        v, n.ctypes.data_as(POINTER(c_int32)), t.ctypes.data_as(POINTER(c_int32)), rows, int(geo_chunk),
        table.ctypes.data_as(POINTER(c_int64)), totals.ctypes.data_as(POINTER(c_int64)))
    check(rc, "tspn_build_video_table")
    return table, totals
