// survivors.cu — everything the heads need from a SURVIVING pair, recomputed from the boxes.
//
//   tspn_survivor_rows   for every row the top-K kept (lib/modeling/relpn/ppn.py:79-90):
//     * the pooled 3 x 2 x 500 relative block of its feature row ([SPEC] s4; the last 3000 columns of
//       lib/dataset/vrdataset.py:219-243) in bf16 + the bias row A_s[subject] + A_o[object] of the
//       decomposed predicate head (include/tspn_b200.h, a14 decomposed) - what tspn_assemble_relative
//       produces from the stored geometry rows;
//     * the decoded temporal-span proposals (DPNHead, lib/modeling/relpn/dpn.py:55-73, at the anchor
//       columns + [SPEC] s5 decode) - what tspn_span_proposals produces from the stored geometry rows.
//
// Why recompute: the stored rows of the K survivors are 64 KB each (268 MB per 4096 rows); reading them
// back right after the all-pairs kernel has written 4 GB through the L2 makes the feature / span kernels
// HBM-bound behind it (measured: 86 us + 59 us on the critical path of a 0.87 ms step).  The boxes of
// the two tracklets are 2 x 32 KB and L2-resident, the per-frame math is ~3 instructions per frame, and
// nothing here depends on the all-pairs kernel - so this kernel, the predicate head and the records run
// on the side stream UNDERNEATH the all-pairs kernel: a CTA is shaped (128 threads = 4 warps x 96 registers,
// 72 KB of shared memory at T = 2000) to co-reside with that kernel's 512-thread / 104-register / 131 KB CTA.
//
// The per-frame values come from the same device function as the all-pairs kernel (geo_math.cuh), the
// pooling adds frames in ascending order like assemble_kernel, the span chain is span_math.cuh: every
// output is bit-identical to the stored-rows path (tests/test_gpu_tensor.py).
#include <cuda_bf16.h>

#include "common.cuh"
#include "geo_math.cuh"
#include "span_math.cuh"

namespace tspn {

constexpr int SV_THREADS = 128;
constexpr int SV_CHUNK = SV_THREADS * GEO_FPT;          // frames per pass = frames per shared-memory block
constexpr int SV_BLOCK_BYTES = SV_CHUNK * 32;           // a block: 512 subject boxes | 512 object boxes (16 KB), later
                                                        // overwritten in place by its [8 channels][512 frames] tile
constexpr int SV_HALF = SV_CHUNK * 16;                  // offset of the object boxes inside a block
constexpr int SV_CIN = TSPN_GEO_CHANNELS;
constexpr int SV_INV_TAB = 32;
constexpr int SV_MAX_PASSES = 128;                      // 65 536 frames
// dynamic shared memory: two blocks (the pass being computed | the next pass's boxes in flight)
constexpr int SV_SMEM_BYTES = 2 * SV_BLOCK_BYTES;

__host__ __device__ __forceinline__ int span_locations(int t, float stride) {
    // len(torch.arange(0, T+1, step=stride)) = ceil((T+1)/stride), anchor_generator.py:50-52
    const double q = ((double)t + 1.0) / (double)stride;
    int n = (int)q;
    if ((double)n < q) ++n;
    return n;
}

// 16-byte async copy global -> shared (SASS: LDGSTS), L2 only
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ uint32_t swizzle128(uint32_t lin) { return lin ^ (((lin >> 7) & 7u) << 4); }

// The window is STREAMED through shared memory 512 frames (one pass) at a time: while pass p is computed, the
// boxes of pass p+1 are in flight into the other block; the pass's [8][512] tile overwrites its boxes in place and
// is consumed at once - every pooled bin adds the frames of its range that lie in the pass (a bin that straddles
// two passes keeps its partial sums in registers; frames are still added in ascending order, so the result is the
// one-tile result bit for bit), and every anchor location is evaluated in the pass that holds its right-hand tap
// (the two taps before it come from the tile or from the last two columns of the previous pass, carried over).
// 32 KB of shared memory per CTA whatever the video length, instead of 16 KB per 512 frames of the LONGEST video of
// the batch (72 KB at T = 2000): two CTAs fit beside the all-pairs kernel's CTA where one did.
#ifndef TSPN_SV_MIN_CTAS
#define TSPN_SV_MIN_CTAS 6          // 6 CTAs of 128 threads per SM: <= 80 registers (16 bytes of spills) - two CTAs fit in
#endif                              // the 20 480 registers a pair kernel at 88 leaves; 5 gives 96 registers, no spills
template <int A>
__global__ void __launch_bounds__(SV_THREADS, TSPN_SV_MIN_CTAS)
survivor_rows_kernel(const int64_t* __restrict__ table, int nv, const float4* __restrict__ boxes,
                     const int32_t* __restrict__ span, const int64_t* __restrict__ rows, int64_t rows_per_video,
                     __nv_bfloat16* __restrict__ rel, int64_t ld_rel, const float* __restrict__ terms_s,
                     const float* __restrict__ terms_o, int n_out, float* __restrict__ row_bias,
                     const float* __restrict__ conv_w, const float* __restrict__ conv_b,
                     const float* __restrict__ pred_w, const float* __restrict__ pred_b,
                     const float* __restrict__ sizes, float stride, int32_t* __restrict__ spans, int64_t ld_spans) {
    constexpr int A2 = 2 * A;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(16) float4 w_conv[SV_CIN * SV_CIN];                      // [co][ci] -> (w0, w1, w2, -)
    __shared__ __align__(16) float w_pred[SV_CIN * A2];                           // [co][j]
    __shared__ float b_conv[SV_CIN], b_pred[A2];
    __shared__ float s_inv[SV_INV_TAB];
    __shared__ __align__(16) float4 s_halo[2][2];        // [buffer][subject | object]: the box right behind the pass
    __shared__ float s_carry[2][2][SV_CIN];              // [pass & 1][col 510 | 511]: a pass's last two tile columns

    const int tid = threadIdx.x;
    const int64_t r = blockIdx.x;
    const int64_t gp = rows[r];
    // rows are laid out [V][rows_per_video] (include/tspn_b200.h): the video is the row's position - no binary search
    // through the table (four dependent global loads at the head of every CTA, under the all-pairs kernel)
    const int v = (int)(r / rows_per_video);
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const int t_len = (int)row[TSPN_VT_T];
    __nv_bfloat16* outb = rel + r * ld_rel;
    float* bias_row = row_bias ? row_bias + r * n_out : nullptr;
    int32_t* sp_row = spans ? spans + r * ld_spans : nullptr;
    const int n_loc = spans ? span_locations(t_len, stride) : 0;

    if (spans) {
        for (int i = tid; i < SV_CIN * SV_CIN; i += SV_THREADS)
            w_conv[i] = make_float4(__ldg(conv_w + i * 3), __ldg(conv_w + i * 3 + 1), __ldg(conv_w + i * 3 + 2), 0.0f);
        for (int i = tid; i < SV_CIN * A2; i += SV_THREADS) {
            const int co = i / A2, j = i - co * A2;
            w_pred[i] = __ldg(pred_w + j * SV_CIN + co);
        }
        if (tid < SV_CIN) b_conv[tid] = conv_b ? __ldg(conv_b + tid) : 0.0f;
        if (tid < A2) b_pred[tid] = pred_b ? __ldg(pred_b + tid) : 0.0f;
        // columns of the span buffer beyond this video's locations (a batch is sized for its longest video)
        for (int64_t i = (int64_t)n_loc * A2 + tid; i < ld_spans; i += SV_THREADS) sp_row[i] = 0;
    }
    if (tid < SV_INV_TAB) s_inv[tid] = 1.0f / (float)(tid ? tid : 1);
    __syncthreads();                               // weights and tables staged (a row without a window never reaches
                                                   // the barriers of the pass loop)

    if (gp < 0) {                                  // padding row (K_eff < K): zero block, zero bias, zero regressions
        for (int64_t q = tid; q < ld_rel; q += SV_THREADS) outb[q] = __float2bfloat16(0.0f);
        if (row_bias)
            for (int c = tid; c < n_out; c += SV_THREADS) bias_row[c] = 0.0f;
        for (int l = tid; l < n_loc; l += SV_THREADS) {
            float zero[A2];
            int32_t res[A2];
#pragma unroll
            for (int j = 0; j < A2; ++j) zero[j] = 0.0f;
            span_decode_location<A>(zero, sizes, __fmul_rn((float)l, stride), t_len, res);
#pragma unroll
            for (int j = 0; j < A2; ++j) sp_row[l * A2 + j] = res[j];
        }
        return;
    }
    const int n = (int)row[TSPN_VT_N];
    const int64_t tb = row[TSPN_VT_TB];
    const int p = (int)(gp - row[TSPN_VT_PAIR_OFF]);
    const int s = p / (n - 1);
    const int k = p - s * (n - 1);
    const int o = k + (k >= s ? 1 : 0);
    const int64_t ts = row[TSPN_VT_TRK_OFF] + s, to = row[TSPN_VT_TRK_OFF] + o;
    const int ps = __ldg(span + 2 * ts), pe = __ldg(span + 2 * ts + 1);
    const int qs = __ldg(span + 2 * to), qe = __ldg(span + 2 * to + 1);
    const int a = max(ps, qs), b = min(pe, qe);                 // temporal overlap window [a, b)
    const uint32_t len = b > a ? (uint32_t)(b - a) : 0u;
    const int a4 = a & ~3;
    const int w0 = a - a4;                                      // tile column of the window's first frame

    const float4* bs = boxes + row[TSPN_VT_BOX_OFF] + (int64_t)s * tb;
    const float4* bo = boxes + row[TSPN_VT_BOX_OFF] + (int64_t)o * tb;
    const int last = (int)tb - 1;                               // frames >= b are masked: any in-row box will do
    const int n_pass = len ? (b - a4 + SV_CHUNK - 1) / SV_CHUNK : 0;
    const uint32_t base = smem_u32(smem);

    // boxes of pass q (both tracklets, 512 frames + the halo box behind them) -> block q & 1, 128-byte swizzle as
    // the all-pairs kernel's TMA stages (conflict-free LDS.128 at a 64-byte thread stride)
    auto prefetch = [&](int q) {
        const uint32_t blk = base + (uint32_t)(q & 1) * SV_BLOCK_BYTES;
        const int f0 = a4 + q * SV_CHUNK;
        for (int idx = tid; idx < SV_CHUNK; idx += SV_THREADS) {
            const int f = min(f0 + idx, last);
            const uint32_t lin = blk + (uint32_t)idx * 16u;
            cp_async16(swizzle128(lin), bs + f);
            cp_async16(swizzle128(lin + SV_HALF), bo + f);
        }
        if (tid < 2) cp_async16(smem_u32(&s_halo[q & 1][tid]), (tid ? bo : bs) + min(f0 + SV_CHUNK, last));
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // ---- per-thread streaming state ---------------------------------------------------------------------
    // pooled bins tid, tid + 128, ...: bin index, next frame (window-relative) to add, partial sums
    int bin = tid;
    uint32_t bin_f = len ? ((uint32_t)bin * len) / TSPN_REL_BINS : 0u;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (len == 0) {                                // no temporal overlap: every pooled bin is zero
        for (int i = tid; i < TSPN_REL_BINS; i += SV_THREADS)
#pragma unroll
            for (int c = 0; c < 6; ++c) outb[c * TSPN_REL_BINS + i] = __float2bfloat16(0.0f);
    }
    if (n_pass) prefetch(0);
    for (int pass = 0; pass < n_pass; ++pass) {
        const uint32_t blk = base + (uint32_t)(pass & 1) * SV_BLOCK_BYTES;
        float* const tile = reinterpret_cast<float*>(smem + (size_t)(pass & 1) * SV_BLOCK_BYTES);   // [8][512]
        if (pass + 1 < n_pass) {
            prefetch(pass + 1);                     // the other block: its tile (pass - 1) was consumed before the
            asm volatile("cp.async.wait_group 1;" ::: "memory");    // barrier that ended the previous iteration
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int j0 = tid * GEO_FPT;
        float out[TSPN_GEO_CHANNELS][GEO_FPT];
        float fi, fs, fo;
        const float4 halo_s = s_halo[pass & 1][0], halo_o = s_halo[pass & 1][1];
        auto load_s = [blk, halo_s](int j) { return j < SV_CHUNK ? ld_box(blk, j) : halo_s; };
        auto load_o = [blk, halo_o](int j) { return j < SV_CHUNK ? ld_box(blk + SV_HALF, j) : halo_o; };
        geo_step<false>(load_s, load_o, j0, a4 + pass * SV_CHUNK + j0, a, b, out, fi, fs, fo);
        __syncthreads();                                        // every thread of the pass holds its boxes
#pragma unroll
        for (int ch = 0; ch < TSPN_GEO_CHANNELS; ++ch)
            *reinterpret_cast<float4*>(tile + ch * SV_CHUNK + j0) = make_float4(out[ch][0], out[ch][1], out[ch][2], out[ch][3]);
        __syncthreads();

        // ---- relative block: the frames of this pass, added to the bins they belong to, ascending ----
        const uint32_t f_end = (uint32_t)min((int)len, (pass + 1) * SV_CHUNK - w0);    // first frame NOT in this tile
        const int col0 = pass * SV_CHUNK - w0;                  // frame f sits in tile column f - col0
        while (bin < TSPN_REL_BINS) {
            const uint32_t en = ((uint32_t)(bin + 1) * len + TSPN_REL_BINS - 1) / TSPN_REL_BINS;
            const uint32_t stop = min(en, f_end);
            for (uint32_t f = bin_f; f < stop; ++f) {           // ascending frames
                const float* x = tile + ((int)f - col0);
                acc[0] += x[0];
                acc[1] += x[SV_CHUNK];
                acc[2] += x[2 * SV_CHUNK];
                acc[3] += x[3 * SV_CHUNK];
                acc[4] += x[5 * SV_CHUNK];
                acc[5] += x[6 * SV_CHUNK];
            }
            if (stop < en) {                                    // the bin continues in the next pass
                bin_f = max(bin_f, stop);
                break;
            }
            const uint32_t st = ((uint32_t)bin * len) / TSPN_REL_BINS;
            const uint32_t width = en - st;
            const float inv = width < SV_INV_TAB ? s_inv[width] : 1.0f / (float)width;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                outb[c * TSPN_REL_BINS + bin] = __float2bfloat16(acc[c] * inv);
                acc[c] = 0.0f;
            }
            bin += SV_THREADS;
            bin_f = bin < TSPN_REL_BINS ? ((uint32_t)bin * len) / TSPN_REL_BINS : 0u;
            if (bin < TSPN_REL_BINS && bin_f >= f_end) break;   // its first frame is in a later pass
        }

        // ---- span proposals: the locations whose right-hand tap lies in this pass ----
        for (int l = tid; l < n_loc; l += SV_THREADS) {
            const float ac = __fmul_rn((float)l, stride);
            const int t = min((int)floorf(ac), t_len - 1);
            int home = (t + 1 - a4) >> 9;                       // pass of frame t + 1 (floor division by 512)
            home = max(0, min(home, n_pass - 1));
            if (home != pass) continue;
            const bool has_m = t > 0, has_p = t + 1 < t_len;
            float xv[SV_CIN][3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int f = t - 1 + d;
                const bool in = f >= a && f < b;                 // every channel is 0 outside the window
                const int col = f - a4 - pass * SV_CHUNK;        // >= -2 for an in-window frame of this location
#pragma unroll
                for (int ci = 0; ci < SV_CIN; ++ci)
                    xv[ci][d] = !in ? 0.0f : (col >= 0 ? tile[ci * SV_CHUNK + col] : s_carry[(pass - 1) & 1][col + 2][ci]);
            }
            int32_t res[A2];
            span_location<SV_CIN, A>(xv, has_m, has_p, w_conv, w_pred, b_conv, b_pred, sizes, ac, t_len, res);
#pragma unroll
            for (int j = 0; j < A2; ++j) sp_row[l * A2 + j] = res[j];
        }
        // this pass's last two columns for the next pass's left-hand taps (its own copy: the readers of this pass
        // use the other one), before the barrier after which the block may be refilled
        if (tid < 2 * SV_CIN) s_carry[pass & 1][tid >> 3][tid & 7] = tile[(tid & 7) * SV_CHUNK + SV_CHUNK - 2 + (tid >> 3)];
        __syncthreads();                                        // tile consumed: the next iteration refills this block
    }
    for (int64_t q = TSPN_REL_DIM + tid; q < ld_rel; q += SV_THREADS) outb[q] = __float2bfloat16(0.0f);
    if (row_bias) {
        const float* a_s = terms_s + ts * n_out;                 // A_s of the subject tracklet
        const float* a_o = terms_o + to * n_out;                 // A_o of the object tracklet
        for (int c = tid; c < n_out; c += SV_THREADS) bias_row[c] = __ldg(a_s + c) + __ldg(a_o + c);
    }
    if (n_pass == 0) {
        // no temporal overlap: every input of the span head is zero (the window is empty)
        for (int l = tid; l < n_loc; l += SV_THREADS) {
            const float ac = __fmul_rn((float)l, stride);
            const int t = min((int)floorf(ac), t_len - 1);
            float xv[SV_CIN][3];
#pragma unroll
            for (int ci = 0; ci < SV_CIN; ++ci) xv[ci][0] = xv[ci][1] = xv[ci][2] = 0.0f;
            int32_t res[A2];
            span_location<SV_CIN, A>(xv, t > 0, t + 1 < t_len, w_conv, w_pred, b_conv, b_pred, sizes, ac, t_len, res);
#pragma unroll
            for (int j = 0; j < A2; ++j) sp_row[l * A2 + j] = res[j];
        }
    }
}

// bias rows of the decomposed predicate head on their own: one warp per scored row
__global__ void __launch_bounds__(128)
gather_pair_terms_kernel(const int64_t* __restrict__ table, int nv, const int64_t* __restrict__ rows, int64_t n_rows,
                         const float* __restrict__ terms_s, const float* __restrict__ terms_o, int n_out,
                         float* __restrict__ row_bias) {
    const int64_t r = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const int lane = threadIdx.x & 31;
    // the grid is sized for a capacity: rows beyond the batch's pairs are padding rows (zeros)
    const int64_t gp = rows ? rows[r] : (r < table_total(table, nv, TSPN_VT_PAIR_OFF) ? r : -1);
    float* dst = row_bias + r * n_out;
    if (gp < 0) {
        for (int c = lane; c < n_out; c += 32) dst[c] = 0.0f;
        return;
    }
    const int v = find_video(table, nv, TSPN_VT_PAIR_OFF, gp);
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const int n = (int)row[TSPN_VT_N];
    const int p = (int)(gp - row[TSPN_VT_PAIR_OFF]);
    const int s = p / (n - 1);
    const int k = p - s * (n - 1);
    const int o = k + (k >= s ? 1 : 0);
    const float* a_s = terms_s + (row[TSPN_VT_TRK_OFF] + s) * n_out;
    const float* a_o = terms_o + (row[TSPN_VT_TRK_OFF] + o) * n_out;
    for (int c = lane; c < n_out; c += 32) dst[c] = __ldg(a_s + c) + __ldg(a_o + c);
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int tspn_survivor_rows_supported(int max_frames, int n_anchors) {
    return n_anchors == 4 && max_frames > 0 && max_frames <= SV_MAX_PASSES * SV_CHUNK;
}

int tspn_gather_pair_terms(const int64_t* d_table, int num_videos, const int64_t* d_rows, int64_t n_rows,
                           const float* d_terms_subject, const float* d_terms_object, int n_outputs, float* d_row_bias,
                           void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && n_rows >= 0 && n_outputs > 0, TSPN_EBADARG, "tspn_gather_pair_terms: bad size");
    if (n_rows == 0) return TSPN_OK;
    TSPN_REQUIRE(d_table && d_terms_subject && d_terms_object && d_row_bias, TSPN_EBADARG,
                 "tspn_gather_pair_terms: null pointer");
    prefer_max_smem(gather_pair_terms_kernel);
    gather_pair_terms_kernel<<<(unsigned)((n_rows + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        d_table, num_videos, d_rows, n_rows, d_terms_subject, d_terms_object, n_outputs, d_row_bias);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_survivor_rows(const int64_t* d_table, int num_videos, int max_frames, const float* d_boxes,
                       const int32_t* d_span, const int64_t* d_rows, int64_t n_rows, int64_t rows_per_video,
                       void* d_rel_bf16, int64_t ld_rel, const float* d_terms_subject, const float* d_terms_object,
                       int n_outputs, float* d_row_bias, const float* d_conv_w, const float* d_conv_b,
                       const float* d_pred_w, const float* d_pred_b, int n_anchors, const float* d_sizes, float stride,
                       int32_t* d_spans, int64_t ld_spans, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && n_rows >= 0 && n_outputs > 0 && max_frames > 0 && rows_per_video > 0, TSPN_EBADARG,
                 "tspn_survivor_rows: bad size");
    if (n_rows == 0) return TSPN_OK;
    TSPN_REQUIRE(d_table && d_boxes && d_span && d_rows && d_rel_bf16, TSPN_EBADARG, "tspn_survivor_rows: null pointer");
    TSPN_REQUIRE(!d_row_bias || (d_terms_subject && d_terms_object), TSPN_EBADARG,
                 "tspn_survivor_rows: bias rows need the per-tracklet terms");
    TSPN_REQUIRE(ld_rel >= TSPN_REL_DIM, TSPN_ESHAPE, "tspn_survivor_rows: ld_rel=%lld must be >= 3000", (long long)ld_rel);
    TSPN_REQUIRE(aligned16(d_boxes), TSPN_EALIGN, "tspn_survivor_rows: boxes must be 16-byte aligned");
    TSPN_REQUIRE(n_rows < (1ll << 31), TSPN_ESHAPE, "tspn_survivor_rows: too many rows");
    TSPN_REQUIRE(tspn_survivor_rows_supported(max_frames, d_spans ? n_anchors : 4), TSPN_ESHAPE,
                 "tspn_survivor_rows: max_frames=%d / n_anchors=%d not supported (use tspn_assemble_relative + "
                 "tspn_span_proposals)", max_frames, n_anchors);
    if (d_spans) {
        TSPN_REQUIRE(d_conv_w && d_pred_w && d_sizes && stride > 0.0f, TSPN_EBADARG,
                     "tspn_survivor_rows: span head weights / anchors missing");
        TSPN_REQUIRE(ld_spans >= (int64_t)span_locations(max_frames, stride) * 2 * n_anchors, TSPN_ESHAPE,
                     "tspn_survivor_rows: ld_spans=%lld < locations(max_frames) * 2A", (long long)ld_spans);
    }
    const int smem = SV_SMEM_BYTES;
    TSPN_CUDA_OK(cudaFuncSetAttribute(survivor_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    prefer_max_smem(survivor_rows_kernel<4>);
    survivor_rows_kernel<4><<<(unsigned)n_rows, SV_THREADS, smem, (cudaStream_t)stream>>>(
        d_table, num_videos, reinterpret_cast<const float4*>(d_boxes), d_span, d_rows, rows_per_video,
        reinterpret_cast<__nv_bfloat16*>(d_rel_bf16), ld_rel, d_terms_subject, d_terms_object, n_outputs, d_row_bias,
        d_conv_w, d_conv_b, d_pred_w, d_pred_b, d_sizes, stride, d_spans, ld_spans);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // extern "C"
