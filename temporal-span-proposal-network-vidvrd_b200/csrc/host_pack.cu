// host_pack.cu — HOST side of the compact box transport (include/tspn_b200.h: tspn_unpack_boxes_spans is its device side).
//
// The serving loop is fed from pinned host batches (serving.PipelinedStage); a B200 step takes 0.7 ms for 16 VidOR
// videos, the numpy packer of the first version took 180 ms for the same batch (boolean masks and fancy indexing over
// 33 MB of boxes) - 250 steps of GPU time per batch packed.  These are the same loops in C++: one pass per tracklet to
// validate the coordinates and decide raw / delta, one to write.  Plain CPU code, no CUDA call: usable (and tested)
// without a GPU.
//
//   tspn_host_pack_boxes_spans  one video's [n][t][4] fp32 boxes -> span-packed u16 slots, raw or delta per tracklet
#include "common.cuh"

namespace tspn {

static inline bool exact_u(float v, float hi, int* out) {
    if (!(v >= 0.0f && v <= hi)) return false;           // also false for NaN
    const int iv = (int)v;
    *out = iv;
    return (float)iv == v;
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int tspn_host_pack_boxes_spans(const float* boxes, int n_tracklets, int n_frames, const int32_t* span, int allow_delta,
                               uint16_t* dst, int64_t dst_slots, int64_t slot0, int64_t* box_off,
                               int64_t* slots_used) {
    TSPN_REQUIRE(n_tracklets >= 0 && n_frames >= 0 && slot0 >= 0 && dst_slots >= slot0, TSPN_EBADARG,
                 "tspn_host_pack_boxes_spans: bad size");
    TSPN_REQUIRE(slots_used, TSPN_EBADARG, "tspn_host_pack_boxes_spans: null pointer");
    *slots_used = 0;
    if (n_tracklets == 0) return TSPN_OK;
    TSPN_REQUIRE(boxes && span && dst && box_off, TSPN_EBADARG, "tspn_host_pack_boxes_spans: null pointer");
    TSPN_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 7u) == 0, TSPN_EALIGN,
                 "tspn_host_pack_boxes_spans: dst must be 8-byte aligned (one slot = 4 x u16)");
    int64_t slot = slot0;
    for (int i = 0; i < n_tracklets; ++i) {
        const int ps = span[2 * i], pe = span[2 * i + 1];
        TSPN_REQUIRE(ps >= 0 && ps <= pe && pe <= n_frames, TSPN_ESHAPE,
                     "tspn_host_pack_boxes_spans: tracklet %d: span [%d, %d) outside [0, %d]", i, ps, pe, n_frames);
        const int64_t len = (int64_t)pe - ps;
        const float* b = boxes + ((int64_t)i * n_frames + ps) * 4;
        // pass 1: exactness, and whether every frame-to-frame difference fits an int8
        bool delta = allow_delta != 0 && len > 0;
        int prev[4] = {0, 0, 0, 0};
        for (int64_t f = 0; f < len; ++f) {
            int cur[4];
            for (int c = 0; c < 4; ++c)
                TSPN_REQUIRE(exact_u(b[f * 4 + c], 65535.0f, &cur[c]), TSPN_ESHAPE,
                             "tspn_host_pack_boxes_spans: tracklet %d frame %lld: coordinate %g is fractional or outside "
                             "[0, 65535] (ship this batch as fp32)", i, (long long)(ps + f), (double)b[f * 4 + c]);
            if (f > 0 && delta)
                for (int c = 0; c < 4; ++c) {
                    const int d = cur[c] - prev[c];
                    if (d < -128 || d > 127) delta = false;
                }
            for (int c = 0; c < 4; ++c) prev[c] = cur[c];
        }
        const int64_t need = delta ? 1 + len / 2 : len;       // delta: the first box + ceil((len - 1) / 2) slots of i8
        TSPN_REQUIRE(slot + need <= dst_slots, TSPN_ESHAPE,
                     "tspn_host_pack_boxes_spans: the packed boxes exceed the arena (%lld slots)", (long long)dst_slots);
        box_off[i] = slot | (delta ? TSPN_PACKED_DELTA : 0ll);
        // pass 2: write
        uint16_t* out = dst + slot * 4;
        if (!delta) {
            for (int64_t k = 0; k < len * 4; ++k) out[k] = (uint16_t)(int)b[k];
        } else {
            for (int c = 0; c < 4; ++c) {
                prev[c] = (int)b[c];
                out[c] = (uint16_t)prev[c];
            }
            int8_t* dl = reinterpret_cast<int8_t*>(out + 4);
            for (int64_t f = 1; f < len; ++f)
                for (int c = 0; c < 4; ++c) {
                    const int cur = (int)b[f * 4 + c];
                    dl[(f - 1) * 4 + c] = (int8_t)(cur - prev[c]);
                    prev[c] = cur;
                }
            if (len > 1 && ((len - 1) & 1))                   // the unused half of the last slot
                for (int c = 0; c < 4; ++c) dl[(len - 1) * 4 + c] = 0;
        }
        slot += need;
    }
    *slots_used = slot - slot0;
    return TSPN_OK;
}

}  // extern "C"
