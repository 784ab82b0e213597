// postproc.cu — relation triplet records (row N1 of SURVEY.md section 8f).
//
//   tspn_postprocess   lib/modeling/predict.py:66-117
//     top `topk_per_pair` (20) predicates of every scored pair  (predict.py:70-73),
//     top `topk_per_video` (200) of those per video             (predict.py:76-81),
//     subject / object class = argmax of the tracklet classeme  (predict.py:88-93, quirk Q4),
//     one fixed 32-byte record per kept triplet — the payload of the multi-GPU all-gather.
//
// Both selections are descending with ties to the lower index ([SPEC] s6); the reference does
// this with two full torch.sort calls and per-element Python list comprehensions.
#include "common.cuh"
#include "topk_block.cuh"

namespace tspn {

constexpr int PP_MAX_R = 256;     // predicates per row handled by one warp (8 per lane)

// one warp per scored row: its topk_per_pair best predicates, in order; SLOTS = ceil(r / 32) values per lane
template <int SLOTS>
__global__ void __launch_bounds__(128)
pair_top_predicates_kernel(const float* __restrict__ logits, const int64_t* __restrict__ rows, int64_t m, int r,
                           int tpp, float* __restrict__ cand_score, int32_t* __restrict__ cand_pred) {
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= m) return;
    const int lane = threadIdx.x & 31;
    const float NEG_INF = __uint_as_float(0xff800000u);
    float* cs = cand_score + row * tpp;
    int32_t* cp = cand_pred + row * tpp;
    if (rows && rows[row] < 0) {                 // padding row: no candidates
        for (int j = lane; j < tpp; j += 32) {
            cs[j] = NEG_INF;
            cp[j] = -1;
        }
        return;
    }
    // Scores as order-preserving unsigned keys (larger float <=> larger key): a round is then the lane's best key,
    // ONE warp REDUX.MAX, and one REDUX.MIN over the predicate indices of the lanes that hold that key (ties to the
    // lower index) - instead of five rounds of two shuffles and a three-way compare.  0 = "no predicate left"
    // (below the key of -inf).
    uint32_t key[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
        const int c = lane + 32 * j;
        uint32_t u = 0u;
        if (c < r) {
            uint32_t b = __float_as_uint(__ldg(logits + row * r + c));
            if (b == 0x80000000u) b = 0u;        // -0.0 ties with +0.0, as in a float compare
            u = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
            if (u == 0u) u = 1u;                 // (only -NaN with all mantissa bits set maps to 0)
        }
        key[j] = u;
    }
    for (int it = 0; it < tpp; ++it) {
        uint32_t best = 0u;
        int bj = 0;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j)
            if (key[j] > best) {                 // ascending j = ascending index: the first maximum wins
                best = key[j];
                bj = j;
            }
        const uint32_t top = __reduce_max_sync(0xffffffffu, best);
        if (top == 0u) {                         // fewer than tpp predicates
            for (int j = it + lane; j < tpp; j += 32) {
                cs[j] = NEG_INF;
                cp[j] = -1;
            }
            break;
        }
        const int bi = (int)__reduce_min_sync(0xffffffffu, best == top ? (unsigned)(lane + 32 * bj) : 0xffffffffu);
        if (lane == 0) {
            const uint32_t b = (top & 0x80000000u) ? (top & 0x7fffffffu) : ~top;
            cs[it] = __uint_as_float(b);
            cp[it] = bi;
        }
        if ((bi & 31) == lane) {
#pragma unroll
            for (int j = 0; j < SLOTS; ++j)
                if (j == (bi >> 5)) key[j] = 0u;
        }
    }
}

__device__ __forceinline__ int argmax_row(const float* __restrict__ x, int c) {
    float best = __ldg(x);
    int bi = 0;
    for (int i = 1; i < c; ++i) {
        const float q = __ldg(x + i);
        if (q > best) {
            best = q;
            bi = i;
        }
    }
    return bi;
}

constexpr int PP_LABEL_CAP = 1024;     // tracklet labels of one video kept in shared memory

// arg-max of one classeme row by a warp: largest value, ties to the lower index (torch.argmax)
__device__ __forceinline__ int warp_argmax_row(const float* __restrict__ x, int c, int lane) {
    float best = __uint_as_float(0xff800000u);
    int bi = 0x7fffffff;
    for (int i = lane; i < c; i += 32) {
        const float q = __ldg(x + i);
        if (q > best) {           // ascending i within a lane: the first maximum wins
            best = q;
            bi = i;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) {
            best = ob;
            bi = oi;
        }
    }
    return bi == 0x7fffffff ? 0 : bi;      // all -inf / NaN rows: index 0, like the sequential scan
}

// one CTA per video: top `tpv` of its (rows x tpp) candidates -> records
__global__ void __launch_bounds__(TOPK_THREADS, 5)      // <= 48 registers: co-resides with the pair kernel
video_top_triplets_kernel(const int64_t* __restrict__ table, int nv, const float* __restrict__ cand_score,
                          const int32_t* __restrict__ cand_pred, const int64_t* __restrict__ rows,
                          const int64_t* __restrict__ row_video_off, int tpp, int tpv,
                          const float* __restrict__ cls, int n_classes, const int32_t* __restrict__ overlap,
                          const int32_t* __restrict__ span, int mirror_q4, int32_t* __restrict__ records,
                          int32_t* __restrict__ counts) {
    __shared__ TopkSmem sm;
    __shared__ int s_label[PP_LABEL_CAP];
    const int v = blockIdx.x;
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const int n = (int)row[TSPN_VT_N];
    // the class label of every tracklet of the video, once (predict.py:88-93 takes the arg-max per kept
    // triplet; 2 x 200 sequential scans of C values were the bulk of this kernel's latency)
    const bool labels_cached = n <= PP_LABEL_CAP;
    if (labels_cached) {
        // four lanes per tracklet: 64 tracklets per round of the block, every lane's loads independent (one memory
        // round trip per round instead of one per tracklet and warp); largest value, ties to the lower index
        const float* cls_v = cls + row[TSPN_VT_TRK_OFF] * n_classes;
        const int sub = threadIdx.x & 3;
        for (int j0 = 0; j0 < n; j0 += TOPK_THREADS / 4) {
            const int j = j0 + (threadIdx.x >> 2);
            float best = __uint_as_float(0xff800000u);
            int bi = 0x7fffffff;
            if (j < n) {
                const float* x = cls_v + (int64_t)j * n_classes;
                for (int i = sub; i < n_classes; i += 4) {
                    const float q = __ldg(x + i);
                    if (q > best) {           // ascending i within a lane: the first maximum wins
                        best = q;
                        bi = i;
                    }
                }
            }
#pragma unroll
            for (int off = 1; off < 4; off <<= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ob > best || (ob == best && oi < bi)) {
                    best = ob;
                    bi = oi;
                }
            }
            if (j < n && sub == 0) s_label[j] = bi == 0x7fffffff ? 0 : bi;   // all -inf / NaN rows: index 0
        }
    }                                       // block_topk synchronises before the labels are read
    const int64_t pair_off = row[TSPN_VT_PAIR_OFF];
    const int64_t r0 = row_video_off ? row_video_off[v] : pair_off;
    const int64_t r1 = row_video_off ? row_video_off[v + 1] : pair_off + (int64_t)n * (n > 0 ? n - 1 : 0);
    const float* cs = cand_score + r0 * tpp;
    const int64_t total = (r1 - r0) * tpp;
    const int k_eff = block_topk(sm, total, tpv, [&](int64_t i) -> float { return __ldg(cs + i); });
    int32_t* rec = records + (int64_t)v * tpv * 8;
    for (int i = threadIdx.x; i < tpv; i += TOPK_THREADS) {
        int32_t out[8] = {0, -1, -1, -1, -1, -1, 0, 0};
        if (i < k_eff) {
            const uint32_t flat = (uint32_t)(sm.sel[i] & 0xffffffffu);
            const int64_t rr = r0 + flat / (uint32_t)tpp;
            const int64_t gp = rows ? rows[rr] : rr;
            const int p = (int)(gp - pair_off);
            const int s = p / (n - 1);
            const int k = p - s * (n - 1);
            const int o = k + (k >= s ? 1 : 0);
            const int64_t trk0 = row[TSPN_VT_TRK_OFF];
            // predict.py:89 reads the object's class from pair row (N-1)*o, whose object is tracklet
            // 0 (or 1 when o == 0): quirk Q4, reproduced only on request
            const int o_src = mirror_q4 ? (o == 0 ? 1 : 0) : o;
            out[0] = __float_as_int(__ldg(cs + flat));
            out[1] = labels_cached ? s_label[s] : argmax_row(cls + (trk0 + s) * n_classes, n_classes);
            out[2] = __ldg(cand_pred + r0 * tpp + flat);
            out[3] = labels_cached ? s_label[o_src] : argmax_row(cls + (trk0 + o_src) * n_classes, n_classes);
            out[4] = s;
            out[5] = o;
            if (overlap) {
                out[6] = __ldg(overlap + 2 * gp);
                out[7] = __ldg(overlap + 2 * gp + 1);
            } else {                 // the same window from the two tracklet spans ([SPEC] s3)
                const int ps = __ldg(span + 2 * (trk0 + s)), pe = __ldg(span + 2 * (trk0 + s) + 1);
                const int qs = __ldg(span + 2 * (trk0 + o)), qe = __ldg(span + 2 * (trk0 + o) + 1);
                const int wa = max(ps, qs), wb = min(pe, qe);
                out[6] = wb > wa ? wa : 0;
                out[7] = wb > wa ? wb : 0;
            }
        }
        int4* dst = reinterpret_cast<int4*>(rec + (int64_t)i * 8);
        dst[0] = make_int4(out[0], out[1], out[2], out[3]);
        dst[1] = make_int4(out[4], out[5], out[6], out[7]);
    }
    if (threadIdx.x == 0) counts[v] = k_eff;
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int64_t tspn_postprocess_workspace_bytes(int64_t m, int topk_per_pair) {
    if (m <= 0 || topk_per_pair <= 0) return 16;
    return m * (int64_t)topk_per_pair * 8 + 16;
}

int tspn_postprocess(const int64_t* d_table, int num_videos, const float* d_logits, const int64_t* d_rows,
                     const int64_t* d_row_video_off, int64_t n_rows, int n_predicates, const float* d_cls,
                     int n_classes, const int32_t* d_overlap, const int32_t* d_span, int topk_per_pair,
                     int topk_per_video, int flags, int32_t* d_records, int32_t* d_counts, void* d_workspace,
                     void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && n_rows >= 0 && n_predicates > 0 && n_classes > 0 && topk_per_pair > 0 &&
                     topk_per_video > 0,
                 TSPN_EBADARG, "tspn_postprocess: bad size");
    TSPN_REQUIRE(n_predicates <= PP_MAX_R, TSPN_ESHAPE, "tspn_postprocess: at most %d predicates", PP_MAX_R);
    TSPN_REQUIRE(topk_per_video <= TOPK_MAX_K, TSPN_ESHAPE, "tspn_postprocess: topk_per_video must be <= %d",
                 TOPK_MAX_K);
    if (num_videos == 0) return TSPN_OK;
    TSPN_REQUIRE(d_table && d_cls && (d_overlap || d_span) && d_records && d_counts && d_workspace &&
                     (d_logits || n_rows == 0),
                 TSPN_EBADARG, "tspn_postprocess: null pointer");
    TSPN_REQUIRE(aligned16(d_records) && aligned16(d_workspace), TSPN_EALIGN,
                 "tspn_postprocess: records/workspace must be 16-byte aligned");
    TSPN_REQUIRE(n_rows * (int64_t)topk_per_pair < (1ll << 31), TSPN_ESHAPE, "tspn_postprocess: too many candidates");
    cudaStream_t st = (cudaStream_t)stream;
    const int tpp = topk_per_pair < n_predicates ? topk_per_pair : n_predicates;
    float* cand_score = reinterpret_cast<float*>(d_workspace);
    int32_t* cand_pred = reinterpret_cast<int32_t*>(cand_score + n_rows * tpp);
    if (n_rows > 0) {
        const unsigned blocks = (unsigned)((n_rows + 3) / 4);
#define TSPN_LAUNCH_PTP(S)                                                                                            \
    do {                                                                                                             \
        prefer_max_smem(pair_top_predicates_kernel<S>);                                                              \
        pair_top_predicates_kernel<S><<<blocks, 128, 0, st>>>(d_logits, d_rows, n_rows, n_predicates, tpp, cand_score, \
                                                              cand_pred);                                            \
    } while (0)
        if (n_predicates <= 32) TSPN_LAUNCH_PTP(1);
        else if (n_predicates <= 64) TSPN_LAUNCH_PTP(2);
        else if (n_predicates <= 128) TSPN_LAUNCH_PTP(4);
        else if (n_predicates <= 160) TSPN_LAUNCH_PTP(5);
        else TSPN_LAUNCH_PTP(8);
#undef TSPN_LAUNCH_PTP
        TSPN_CUDA_OK(cudaGetLastError());
    }
    prefer_max_smem(video_top_triplets_kernel);
    video_top_triplets_kernel<<<(unsigned)num_videos, TOPK_THREADS, 0, st>>>(
        d_table, num_videos, cand_score, cand_pred, d_rows, d_row_video_off, tpp, topk_per_video, d_cls, n_classes,
        d_overlap, d_span, (flags & TSPN_POST_MIRROR_Q4) ? 1 : 0, d_records, d_counts);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // extern "C"
