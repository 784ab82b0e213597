// features.cu — per-pair feature rows (lib/dataset/vrdataset.py:219-243).
//
//   [0,C) subject classeme | [C,2C) object classeme | 4000 subject motion BoW (4 blocks of
//   1000, each L1-normalised: vrdataset.py:227-236 + utils/miscellaneous.py:32-35) |
//   4000 object motion BoW | 3000 relative block = adaptive average pooling of the geometry
//   channels (0,1 | 2,3 | 5,6) over the pair's temporal overlap window to 500 bins each
//   ([SPEC] s4; the reference loads these 3000 columns precomputed from h5).
//
// HBM-write-bound: 4*F bytes per row; the per-tracklet blocks are re-read from L2.
#include <cuda_bf16.h>

#include "common.cuh"

namespace tspn {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// one warp per (tracklet, 1000-wide block)
__global__ void __launch_bounds__(128) normalize_motion_kernel(const float* __restrict__ motion, int64_t n_blocks,
                                                               float* __restrict__ out) {
    const int64_t blk = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (blk >= n_blocks) return;
    const int lane = threadIdx.x & 31;
    const float* src = motion + blk * TSPN_MOTION_BLOCK;
    float* dst = out + blk * TSPN_MOTION_BLOCK;
    float s = 0.0f;
    for (int i = lane; i < TSPN_MOTION_BLOCK; i += 32) s += fabsf(__ldg(src + i));
    s = warp_sum_f(s);
    if (s == 0.0f) s = 1.0f;                      // miscellaneous.py:34 — empty histograms stay zero
    for (int i = lane; i < TSPN_MOTION_BLOCK; i += 32) dst[i] = __ldg(src + i) / s;
}

// u8 counts -> L1-normalised fp32 blocks (compact transport of the motion histograms, see batch.py)
__global__ void __launch_bounds__(128) normalize_motion_u8_kernel(const uint8_t* __restrict__ motion, int64_t n_blocks,
                                                                  float* __restrict__ out) {
    const int64_t blk = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (blk >= n_blocks) return;
    const int lane = threadIdx.x & 31;
    const uint8_t* src = motion + blk * TSPN_MOTION_BLOCK;       // 1000 bytes: 4-byte aligned
    float* dst = out + blk * TSPN_MOTION_BLOCK;
    // integer counts: the fp32 sum is exact (<= 255 000) and equals the reference's sum of |x| in any order
    uint32_t isum = 0;
    for (int i = lane; i < TSPN_MOTION_BLOCK / 4; i += 32) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
        isum += (w & 0xffu) + ((w >> 8) & 0xffu) + ((w >> 16) & 0xffu) + (w >> 24);
    }
    isum = __reduce_add_sync(0xffffffffu, isum);
    const float s = isum ? (float)isum : 1.0f;                   // miscellaneous.py:34 - empty histograms stay zero
    for (int i = lane; i < TSPN_MOTION_BLOCK / 4; i += 32) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
        const float4 v = make_float4((float)(w & 0xffu) / s, (float)((w >> 8) & 0xffu) / s,
                                     (float)((w >> 16) & 0xffu) / s, (float)(w >> 24) / s);
        *reinterpret_cast<float4*>(dst + 4 * i) = v;
    }
}

// u16 pixel coordinates -> fp32 boxes (compact transport of integer boxes)
__global__ void __launch_bounds__(256) unpack_boxes_u16_kernel(const uint2* __restrict__ src, int64_t n_boxes,
                                                               float4* __restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_boxes; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 w = __ldg(src + i);
        dst[i] = make_float4((float)(w.x & 0xffffu), (float)(w.x >> 16), (float)(w.y & 0xffffu), (float)(w.y >> 16));
    }
}

// span-packed u16 boxes -> dense fp32 rows: tracklet n's packed boxes (frames [pstart, pend) only, at 8-byte slot
// pk_off[n] of `packed`) go to its dense row at frames pstart .., every other frame of the row is written as zero.
// Two encodings per tracklet (bit 62 of pk_off[n], TSPN_PACKED_DELTA):
//   raw    one slot per frame: 4 x u16 pixel coordinates;
//   delta  slot 0 = the first frame's 4 x u16, then 4 x i8 per further frame (two frames per slot): the coordinate
//          differences to the previous frame - what the host emits for every tracklet whose boxes move by at most
//          [-128, 127] pixels per frame (half the bytes; tracked objects do).  Decoding is a prefix sum over the
//          frames: tiles of 256 frames, a warp scan per coordinate + the warps' totals through shared memory + the
//          running box carried from tile to tile - integer adds, so the boxes are the host's integers exactly.
// One CTA per tracklet; the dense row is box_off(video) + n_local * Tb.
__global__ void __launch_bounds__(256) unpack_boxes_spans_kernel(const int64_t* __restrict__ table, int nv,
                                                                 const int32_t* __restrict__ span,
                                                                 const int64_t* __restrict__ pk_off,
                                                                 const uint2* __restrict__ packed,
                                                                 float4* __restrict__ dst) {
    __shared__ int4 warp_tot[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_trk = table_total(table, nv, TSPN_VT_TRK_OFF);
    for (int64_t trk = blockIdx.x; trk < n_trk; trk += gridDim.x) {
        const int v = find_video(table, nv, TSPN_VT_TRK_OFF, trk);
        const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
        const int tb = (int)row[TSPN_VT_TB];
        float4* out = dst + row[TSPN_VT_BOX_OFF] + (trk - row[TSPN_VT_TRK_OFF]) * tb;
        const int ps = __ldg(span + 2 * trk), pe = __ldg(span + 2 * trk + 1);
        const int64_t po = __ldg(pk_off + trk);
        const int64_t slot = po & ~TSPN_PACKED_DELTA;
        if (!(po & TSPN_PACKED_DELTA)) {
            const uint2* src = packed + slot - ps;
            for (int f = threadIdx.x; f < tb; f += 256) {
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (f >= ps && f < pe) {
                    const uint2 w = __ldg(src + f);
                    b = make_float4((float)(w.x & 0xffffu), (float)(w.x >> 16), (float)(w.y & 0xffffu), (float)(w.y >> 16));
                }
                out[f] = b;
            }
            continue;
        }
        for (int f = threadIdx.x; f < tb; f += 256)
            if (f < ps || f >= pe) out[f] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int n_delta = pe - ps - 1;                          // frames after the first
        if (n_delta < 0) continue;
        const uint2 w0 = __ldg(packed + slot);
        int4 run = make_int4((int)(w0.x & 0xffffu), (int)(w0.x >> 16), (int)(w0.y & 0xffffu), (int)(w0.y >> 16));
        if (threadIdx.x == 0) out[ps] = make_float4((float)run.x, (float)run.y, (float)run.z, (float)run.w);
        const int32_t* dl = reinterpret_cast<const int32_t*>(packed + slot + 1);
        for (int j0 = 0; j0 < n_delta; j0 += 256) {               // uniform trip count: barriers inside
            const int j = j0 + threadIdx.x;
            int4 d = make_int4(0, 0, 0, 0);
            if (j < n_delta) {
                const int32_t w = __ldg(dl + j);
                d = make_int4((int)(int8_t)(w & 0xff), (int)(int8_t)((w >> 8) & 0xff), (int)(int8_t)((w >> 16) & 0xff),
                              (int)(int8_t)((w >> 24) & 0xff));
            }
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int ux = __shfl_up_sync(0xffffffffu, d.x, off), uy = __shfl_up_sync(0xffffffffu, d.y, off);
                const int uz = __shfl_up_sync(0xffffffffu, d.z, off), uw = __shfl_up_sync(0xffffffffu, d.w, off);
                if (lane >= off) { d.x += ux; d.y += uy; d.z += uz; d.w += uw; }
            }
            if (lane == 31) warp_tot[warp] = d;
            __syncthreads();
            int4 pre = run, tot = run;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const int4 t = warp_tot[w];
                if (w < warp) { pre.x += t.x; pre.y += t.y; pre.z += t.z; pre.w += t.w; }
                tot.x += t.x; tot.y += t.y; tot.z += t.z; tot.w += t.w;
            }
            if (j < n_delta)
                out[ps + 1 + j] = make_float4((float)(pre.x + d.x), (float)(pre.y + d.y), (float)(pre.z + d.z),
                                              (float)(pre.w + d.w));
            __syncthreads();                                      // warp_tot is rewritten by the next tile
            run = tot;
        }
    }
}

// Tracklet rows for the decomposed predicate head: [cls (C) | L1-normalised motion (4000) | 0-pad] in bf16,
// one CTA of 128 threads per tracklet (warp w normalises BoW block w).  MOTION_U8: compact transport.
template <bool MOTION_U8>
__global__ void __launch_bounds__(128) tracklet_rows_kernel(const float* __restrict__ cls, int n_classes,
                                                            const void* __restrict__ motion, int64_t n_tracklets,
                                                            __nv_bfloat16* __restrict__ out, int64_t ld) {
    const int64_t trk = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __nv_bfloat16* dst = out + trk * ld;
    for (int i = threadIdx.x; i < n_classes; i += 128) dst[i] = __float2bfloat16(__ldg(cls + trk * n_classes + i));
    for (int64_t i = n_classes + TSPN_MOTION_DIM + threadIdx.x; i < ld; i += 128) dst[i] = __float2bfloat16(0.0f);
    __nv_bfloat16* md = dst + n_classes + warp * TSPN_MOTION_BLOCK;
    if (MOTION_U8) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(motion) + trk * TSPN_MOTION_DIM + warp * TSPN_MOTION_BLOCK;
        const bool vec = ((reinterpret_cast<uintptr_t>(src) & 3) == 0) && ((reinterpret_cast<uintptr_t>(md) & 7) == 0);
        uint32_t isum = 0;
        if (vec) {
            // four counts per load, four bf16 per store: a quarter of the load/store instructions (this kernel runs
            // beside the all-pairs kernel, where every LSU instruction is expensive); same values, same rounding
            const uint32_t* src4 = reinterpret_cast<const uint32_t*>(src);
            for (int i = lane; i < TSPN_MOTION_BLOCK / 4; i += 32) {
                const uint32_t w = __ldg(src4 + i);
                isum += (w & 0xffu) + ((w >> 8) & 0xffu) + ((w >> 16) & 0xffu) + (w >> 24);
            }
            isum = __reduce_add_sync(0xffffffffu, isum);
            const float s = isum ? (float)isum : 1.0f;
            for (int i = lane; i < TSPN_MOTION_BLOCK / 4; i += 32) {
                const uint32_t w = __ldg(src4 + i);
                const __nv_bfloat162 lo = __floats2bfloat162_rn((float)(w & 0xffu) / s, (float)((w >> 8) & 0xffu) / s);
                const __nv_bfloat162 hi = __floats2bfloat162_rn((float)((w >> 16) & 0xffu) / s, (float)(w >> 24) / s);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(md + 4 * i) = pk;
            }
        } else {
            for (int i = lane; i < TSPN_MOTION_BLOCK; i += 32) isum += __ldg(src + i);
            isum = __reduce_add_sync(0xffffffffu, isum);
            const float s = isum ? (float)isum : 1.0f;
            for (int i = lane; i < TSPN_MOTION_BLOCK; i += 32) md[i] = __float2bfloat16((float)__ldg(src + i) / s);
        }
    } else {
        const float* src = reinterpret_cast<const float*>(motion) + trk * TSPN_MOTION_DIM + warp * TSPN_MOTION_BLOCK;
        float s = 0.0f;
        for (int i = lane; i < TSPN_MOTION_BLOCK; i += 32) s += fabsf(__ldg(src + i));
        s = warp_sum_f(s);
        if (s == 0.0f) s = 1.0f;
        for (int i = lane; i < TSPN_MOTION_BLOCK; i += 32) md[i] = __float2bfloat16(__ldg(src + i) / s);
    }
}

constexpr int ASM_THREADS = 256;
constexpr int ASM_POOLED = 6;       // pooled geometry channels: (0,1) position, (2,3) size, (5,6) motion
constexpr int ASM_INV_TAB = 32;     // bin widths with a tabulated reciprocal

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// The row as a function of the column: [0,C) subject classeme | [C,2C) object classeme | subject motion |
// object motion | pooled relative block (shared memory) | zero padding up to the leading dimension.
struct RowSource {
    const float* cls_s; const float* cls_o; const float* mot_s; const float* mot_o; const float* rel;
    int C, m0, m1, r0, F;
    __device__ __forceinline__ float at(int col) const {
        if (col < C) return __ldg(cls_s + col);
        if (col < m0) return __ldg(cls_o + (col - C));
        if (col < m1) return __ldg(mot_s + (col - m0));
        if (col < r0) return __ldg(mot_o + (col - m1));
        if (col < F) return rel[col - r0];
        return 0.0f;
    }
    // four consecutive columns starting at a multiple of 4; every segment boundary is a multiple of 4
    __device__ __forceinline__ float4 at4(int col) const {
        if (col < C) return __ldg(reinterpret_cast<const float4*>(cls_s + col));
        if (col < m0) return __ldg(reinterpret_cast<const float4*>(cls_o + (col - C)));
        if (col < m1) return __ldg(reinterpret_cast<const float4*>(mot_s + (col - m0)));
        if (col < r0) return __ldg(reinterpret_cast<const float4*>(mot_o + (col - m1)));
        if (col < F) return *reinterpret_cast<const float4*>(rel + (col - r0));
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
};

// One CTA per output row.  (1) Bin i of every pooled channel covers the same frames of the pair's overlap
// window, so a thread owns bin i of all six channels: one index computation, six independent load streams
// straight from global memory (a warp's 32 bins are one contiguous run of frames per channel; the bins'
// later frames hit the lines the first ones brought into L1), frames added in ascending order; the 3 x 2 x
// 500 bins go to shared memory.  No staging round trips and no barrier until the block is complete - the
// staged form this replaces spent 14 k warp-instructions per row on load/sync/pool rounds.  (2) The whole
// row leaves as aligned 128-bit stores, each thread gathering the 4 (fp32) / 8 (bf16) columns of its vector
// from the classeme / normalised-motion rows (L2) or the pooled block.
// REL_ONLY (decomposed predicate head): the row is just the pooled relative block in bf16, and the CTA also
// emits the row's bias  A_s[subject] + A_o[object]  from the per-tracklet terms (cls = A_s [Ntrk][R],
// motion_norm = A_o [Ntrk][R], n_classes = R, feat = the [rows][R] bias rows).
template <bool BF16, bool REL_ONLY = false>
__global__ void __launch_bounds__(ASM_THREADS)
assemble_kernel(const int64_t* __restrict__ table, int nv, const float* __restrict__ cls, int n_classes,
                const float* __restrict__ motion_norm, const float* __restrict__ geo,
                const int32_t* __restrict__ overlap, const int64_t* __restrict__ rows, float* __restrict__ feat,
                int64_t ld_feat, __nv_bfloat16* __restrict__ feat_bf16, int64_t ld_bf16) {
    __shared__ __align__(16) float s_rel[TSPN_REL_DIM];
    __shared__ float s_inv[ASM_INV_TAB];
    const int64_t r = blockIdx.x;
    // the grid is sized for a capacity: rows beyond the batch's pairs are padding rows (written as zeros, so that
    // whatever runs over the whole buffer behind this kernel reads initialised memory)
    const int64_t gp = rows ? rows[r] : (r < table_total(table, nv, TSPN_VT_PAIR_OFF) ? r : -1);
    float* out = feat ? feat + r * ld_feat : nullptr;
    __nv_bfloat16* outb = BF16 ? feat_bf16 + r * ld_bf16 : nullptr;
    const int C = n_classes;
    const int F = 2 * C + 2 * TSPN_MOTION_DIM + TSPN_REL_DIM;
    if (REL_ONLY && gp < 0) {                      // padding row: zero relative block, zero bias row
        for (int64_t q = threadIdx.x; q < ld_bf16 / 8; q += ASM_THREADS)
            reinterpret_cast<uint4*>(outb)[q] = make_uint4(0u, 0u, 0u, 0u);
        for (int c = threadIdx.x; c < C; c += ASM_THREADS) out[c] = 0.0f;
        return;
    }
    if (gp < 0) {                                  // padding row (e.g. K_eff < K): zeros
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (out)
            for (int64_t q = threadIdx.x; q < ld_feat / 4; q += ASM_THREADS) reinterpret_cast<float4*>(out)[q] = z;
        if (BF16)
            for (int64_t q = threadIdx.x; q < ld_bf16 / 8; q += ASM_THREADS)
                reinterpret_cast<uint4*>(outb)[q] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const int v = find_video(table, nv, TSPN_VT_PAIR_OFF, gp);
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const int n = (int)row[TSPN_VT_N];
    const int tp = (int)row[TSPN_VT_TP];
    const int p = (int)(gp - row[TSPN_VT_PAIR_OFF]);
    const int s = p / (n - 1);
    const int k = p - s * (n - 1);
    const int o = k + (k >= s ? 1 : 0);
    const int64_t ts = row[TSPN_VT_TRK_OFF] + s, to = row[TSPN_VT_TRK_OFF] + o;

    // ---- relative block ----
    const int a = __ldg(overlap + 2 * gp), b = __ldg(overlap + 2 * gp + 1);
    const uint32_t len = b > a ? (uint32_t)(b - a) : 0u;
    // the six pooled channel rows, at the first frame of the window
    const float* g0 = geo + row[TSPN_VT_GEO_OFF] + (int64_t)p * TSPN_GEO_CHANNELS * tp + a;
    const float* g1 = g0 + tp;
    const float* g2 = g0 + 2 * (int64_t)tp;
    const float* g3 = g0 + 3 * (int64_t)tp;
    const float* g5 = g0 + 5 * (int64_t)tp;
    const float* g6 = g0 + 6 * (int64_t)tp;
    // 1 / (frames per bin) for the few bin widths that occur (one IEEE division per width instead of one per bin)
    if (threadIdx.x < ASM_INV_TAB) s_inv[threadIdx.x] = 1.0f / (float)(threadIdx.x ? threadIdx.x : 1);
    __syncthreads();
    for (int i = threadIdx.x; i < TSPN_REL_BINS; i += ASM_THREADS) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a5 = 0.f, a6 = 0.f;
        if (len > 0) {
            const uint32_t st = ((uint32_t)i * len) / TSPN_REL_BINS;              // len < 2^22: fits 32 bits
            const uint32_t en = ((uint32_t)(i + 1) * len + TSPN_REL_BINS - 1) / TSPN_REL_BINS;
            const uint32_t width = en - st;
            const float inv = width < ASM_INV_TAB ? s_inv[width] : 1.0f / (float)width;
#pragma unroll 2
            for (uint32_t f = st; f < en; ++f) {                                  // ascending frames
                a0 += __ldg(g0 + f);
                a1 += __ldg(g1 + f);
                a2 += __ldg(g2 + f);
                a3 += __ldg(g3 + f);
                a5 += __ldg(g5 + f);
                a6 += __ldg(g6 + f);
            }
            a0 *= inv; a1 *= inv; a2 *= inv; a3 *= inv; a5 *= inv; a6 *= inv;
        }
        s_rel[0 * TSPN_REL_BINS + i] = a0;
        s_rel[1 * TSPN_REL_BINS + i] = a1;
        s_rel[2 * TSPN_REL_BINS + i] = a2;
        s_rel[3 * TSPN_REL_BINS + i] = a3;
        s_rel[4 * TSPN_REL_BINS + i] = a5;
        s_rel[5 * TSPN_REL_BINS + i] = a6;
    }
    __syncthreads();

    if (REL_ONLY) {
        for (int col = 8 * threadIdx.x; col < (int)ld_bf16; col += 8 * ASM_THREADS) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = col + j < TSPN_REL_DIM ? s_rel[col + j] : 0.0f;
            *reinterpret_cast<uint4*>(outb + col) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                                pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
        const float* a_s = cls + ts * C;                     // A_s of the subject tracklet
        const float* a_o = motion_norm + to * C;             // A_o of the object tracklet
        for (int c = threadIdx.x; c < C; c += ASM_THREADS) out[c] = __ldg(a_s + c) + __ldg(a_o + c);
        return;
    }
    // ---- the row, as aligned vectors ----
    RowSource src;
    src.cls_s = cls + ts * C; src.cls_o = cls + to * C;
    src.mot_s = motion_norm + ts * TSPN_MOTION_DIM; src.mot_o = motion_norm + to * TSPN_MOTION_DIM;
    src.rel = s_rel;
    src.C = C; src.m0 = 2 * C; src.m1 = src.m0 + TSPN_MOTION_DIM; src.r0 = src.m1 + TSPN_MOTION_DIM; src.F = F;
    const bool fast = (C & 3) == 0 && ((reinterpret_cast<uintptr_t>(cls) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(motion_norm) & 15) == 0);
    const int ld_max = (int)max(out ? ld_feat : (int64_t)0, BF16 ? ld_bf16 : (int64_t)0);
    for (int col = 8 * threadIdx.x; col < ld_max; col += 8 * ASM_THREADS) {
        float4 lo, hi;
        if (fast) {
            lo = src.at4(col);
            hi = src.at4(col + 4);
        } else {
            lo = make_float4(src.at(col), src.at(col + 1), src.at(col + 2), src.at(col + 3));
            hi = make_float4(src.at(col + 4), src.at(col + 5), src.at(col + 6), src.at(col + 7));
        }
        if (out) {
            if (col + 4 <= ld_feat) *reinterpret_cast<float4*>(out + col) = lo;
            if (col + 8 <= ld_feat) *reinterpret_cast<float4*>(out + col + 4) = hi;
        }
        if (BF16 && col + 8 <= ld_bf16)
            *reinterpret_cast<uint4*>(outb + col) = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w),
                                                                pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
    }
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int tspn_normalize_motion(const float* d_motion, int64_t n_tracklets, float* d_out, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(n_tracklets >= 0, TSPN_EBADARG, "tspn_normalize_motion: negative size");
    if (n_tracklets == 0) return TSPN_OK;
    TSPN_REQUIRE(d_motion && d_out, TSPN_EBADARG, "tspn_normalize_motion: null pointer");
    const int64_t n_blocks = n_tracklets * (TSPN_MOTION_DIM / TSPN_MOTION_BLOCK);
    normalize_motion_kernel<<<(unsigned)((n_blocks + 3) / 4), 128, 0, (cudaStream_t)stream>>>(d_motion, n_blocks,
                                                                                              d_out);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_normalize_motion_u8(const uint8_t* d_motion, int64_t n_tracklets, float* d_out, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(n_tracklets >= 0, TSPN_EBADARG, "tspn_normalize_motion_u8: negative size");
    if (n_tracklets == 0) return TSPN_OK;
    TSPN_REQUIRE(d_motion && d_out, TSPN_EBADARG, "tspn_normalize_motion_u8: null pointer");
    TSPN_REQUIRE((reinterpret_cast<uintptr_t>(d_motion) & 3u) == 0 && aligned16(d_out), TSPN_EALIGN,
                 "tspn_normalize_motion_u8: motion must be 4-byte and out 16-byte aligned");
    const int64_t n_blocks = n_tracklets * (TSPN_MOTION_DIM / TSPN_MOTION_BLOCK);
    normalize_motion_u8_kernel<<<(unsigned)((n_blocks + 3) / 4), 128, 0, (cudaStream_t)stream>>>(d_motion, n_blocks,
                                                                                                 d_out);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_unpack_boxes_u16(const uint16_t* d_src, int64_t n_boxes, float* d_dst, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(n_boxes >= 0, TSPN_EBADARG, "tspn_unpack_boxes_u16: negative size");
    if (n_boxes == 0) return TSPN_OK;
    TSPN_REQUIRE(d_src && d_dst, TSPN_EBADARG, "tspn_unpack_boxes_u16: null pointer");
    TSPN_REQUIRE((reinterpret_cast<uintptr_t>(d_src) & 7u) == 0 && aligned16(d_dst), TSPN_EALIGN,
                 "tspn_unpack_boxes_u16: src must be 8-byte and dst 16-byte aligned");
    int64_t blocks = (n_boxes + 255) / 256;
    const int64_t cap = 16 * (int64_t)num_sms();
    if (blocks > cap) blocks = cap;
    unpack_boxes_u16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint2*>(d_src), n_boxes, reinterpret_cast<float4*>(d_dst));
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_unpack_boxes_spans(const int64_t* d_table, int num_videos, int64_t total_tracklets, const int32_t* d_span,
                            const int64_t* d_packed_off, const uint16_t* d_packed, float* d_dst, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && total_tracklets >= 0, TSPN_EBADARG, "tspn_unpack_boxes_spans: negative size");
    if (total_tracklets == 0) return TSPN_OK;
    TSPN_REQUIRE(d_table && d_span && d_packed_off && d_packed && d_dst, TSPN_EBADARG,
                 "tspn_unpack_boxes_spans: null pointer");
    TSPN_REQUIRE((reinterpret_cast<uintptr_t>(d_packed) & 7u) == 0 && aligned16(d_dst), TSPN_EALIGN,
                 "tspn_unpack_boxes_spans: src must be 8-byte and dst 16-byte aligned");
    int64_t blocks = total_tracklets;
    const int64_t cap = 16 * (int64_t)num_sms();
    if (blocks > cap) blocks = cap;
    unpack_boxes_spans_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_table, num_videos, d_span, d_packed_off, reinterpret_cast<const uint2*>(d_packed),
        reinterpret_cast<float4*>(d_dst));
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_tracklet_rows(const float* d_cls, int n_classes, const void* d_motion, int motion_is_u8,
                       int64_t n_tracklets, void* d_out_bf16, int64_t ld, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(n_tracklets >= 0 && n_classes > 0, TSPN_EBADARG, "tspn_tracklet_rows: bad size");
    if (n_tracklets == 0) return TSPN_OK;
    TSPN_REQUIRE(d_cls && d_motion && d_out_bf16, TSPN_EBADARG, "tspn_tracklet_rows: null pointer");
    TSPN_REQUIRE(ld >= n_classes + TSPN_MOTION_DIM && (ld & 7) == 0, TSPN_ESHAPE,
                 "tspn_tracklet_rows: ld=%lld must be >= C+4000 and a multiple of 8", (long long)ld);
    TSPN_REQUIRE(aligned16(d_out_bf16), TSPN_EALIGN, "tspn_tracklet_rows: output must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    prefer_max_smem(tracklet_rows_kernel<true>);
    prefer_max_smem(tracklet_rows_kernel<false>);
    if (motion_is_u8)
        tracklet_rows_kernel<true><<<(unsigned)n_tracklets, 128, 0, st>>>(d_cls, n_classes, d_motion, n_tracklets,
                                                                          reinterpret_cast<__nv_bfloat16*>(d_out_bf16), ld);
    else
        tracklet_rows_kernel<false><<<(unsigned)n_tracklets, 128, 0, st>>>(d_cls, n_classes, d_motion, n_tracklets,
                                                                           reinterpret_cast<__nv_bfloat16*>(d_out_bf16), ld);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_assemble_relative(const int64_t* d_table, int num_videos, int64_t total_pairs, int max_frames,
                           const float* d_geo, const int32_t* d_overlap, const int64_t* d_rows, int64_t n_rows,
                           void* d_rel_bf16, int64_t ld_rel, const float* d_terms_subject,
                           const float* d_terms_object, int n_outputs, float* d_row_bias, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && total_pairs >= 0 && n_rows >= 0 && n_outputs > 0 && max_frames >= 0, TSPN_EBADARG,
                 "tspn_assemble_relative: bad size");
    if (!d_rows) n_rows = total_pairs;
    if (n_rows == 0) return TSPN_OK;
    TSPN_REQUIRE(d_table && d_geo && d_overlap && d_rel_bf16 && d_terms_subject && d_terms_object && d_row_bias,
                 TSPN_EBADARG,
                 "tspn_assemble_relative: null pointer");
    TSPN_REQUIRE(ld_rel >= TSPN_REL_DIM && (ld_rel & 7) == 0, TSPN_ESHAPE,
                 "tspn_assemble_relative: ld_rel=%lld must be >= 3000 and a multiple of 8", (long long)ld_rel);
    TSPN_REQUIRE(aligned16(d_rel_bf16), TSPN_EALIGN, "tspn_assemble_relative: output must be 16-byte aligned");
    TSPN_REQUIRE(n_rows < (1ll << 31), TSPN_ESHAPE, "tspn_assemble_relative: too many rows");
    prefer_max_smem((assemble_kernel<true, true>));
    assemble_kernel<true, true><<<(unsigned)n_rows, ASM_THREADS, 0, (cudaStream_t)stream>>>(
        d_table, num_videos, d_terms_subject, n_outputs, d_terms_object, d_geo, d_overlap, d_rows, d_row_bias, n_outputs,
        reinterpret_cast<__nv_bfloat16*>(d_rel_bf16), ld_rel);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_assemble_features(const int64_t* d_table, int num_videos, int64_t total_pairs, int max_frames,
                           const float* d_cls, int n_classes, const float* d_motion_norm, const float* d_geo,
                           const int32_t* d_overlap, const int64_t* d_rows, int64_t n_rows, float* d_feat,
                           int64_t ld_feat, void* d_feat_bf16, int64_t ld_bf16, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && total_pairs >= 0 && n_rows >= 0 && n_classes > 0, TSPN_EBADARG,
                 "tspn_assemble_features: bad size");
    if (!d_rows) n_rows = total_pairs;
    if (n_rows == 0) return TSPN_OK;
    const int64_t F = 2 * (int64_t)n_classes + 2 * TSPN_MOTION_DIM + TSPN_REL_DIM;
    TSPN_REQUIRE(d_table && d_cls && d_motion_norm && d_geo && d_overlap && (d_feat || d_feat_bf16), TSPN_EBADARG,
                 "tspn_assemble_features: null pointer");
    TSPN_REQUIRE(!d_feat || (ld_feat >= F && (ld_feat & 3) == 0), TSPN_ESHAPE,
                 "tspn_assemble_features: ld_feat=%lld must be >= F=%lld and a multiple of 4", (long long)ld_feat,
                 (long long)F);
    TSPN_REQUIRE(!d_feat_bf16 || (ld_bf16 >= F && (ld_bf16 & 7) == 0), TSPN_ESHAPE,
                 "tspn_assemble_features: ld_bf16=%lld must be >= F=%lld and a multiple of 8", (long long)ld_bf16,
                 (long long)F);
    TSPN_REQUIRE(aligned16(d_feat) && aligned16(d_feat_bf16), TSPN_EALIGN,
                 "tspn_assemble_features: outputs must be 16-byte aligned");
    TSPN_REQUIRE(n_rows < (1ll << 31), TSPN_ESHAPE, "tspn_assemble_features: too many rows");
    TSPN_REQUIRE(max_frames >= 0, TSPN_EBADARG, "tspn_assemble_features: max_frames=%d", max_frames);
    cudaStream_t st = (cudaStream_t)stream;
    if (d_feat_bf16) {
        assemble_kernel<true><<<(unsigned)n_rows, ASM_THREADS, 0, st>>>(
            d_table, num_videos, d_cls, n_classes, d_motion_norm, d_geo, d_overlap, d_rows, d_feat, ld_feat,
            reinterpret_cast<__nv_bfloat16*>(d_feat_bf16), ld_bf16);
    } else {
        assemble_kernel<false><<<(unsigned)n_rows, ASM_THREADS, 0, st>>>(
            d_table, num_videos, d_cls, n_classes, d_motion_norm, d_geo, d_overlap, d_rows, d_feat, ld_feat, nullptr,
            0);
    }
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // extern "C"
