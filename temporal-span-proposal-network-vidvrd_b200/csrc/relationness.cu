// relationness.cu — PPNHead scores of every ordered pair and their top-K selection.
//
//   tspn_relationness   PPNHead.forward           lib/modeling/relpn/ppn.py:92-112
//   tspn_topk_pairs     PPN._forward_test sort    lib/modeling/relpn/ppn.py:79-90
//
// Arithmetic is the fixed-order fp32 definition of DESIGN.md ("exact-order arithmetic"):
// every linear layer is a k-ascending fma chain starting from the bias, the pair score is a
// c-ascending fma chain from 0, the sigmoid is 1/(1+exp_det(-z)) with an exp made of
// fma/mul/add/rint only — so the scores, and with them the top-K selection, are
// bit-reproducible (and bit-identical to oracle/exact).  The problem is tiny (<= 21 MFLOP at
// N=256): latency-bound, not tensor-bound; it never approaches either roofline.
//
// Top-K is a block-level MSD radix select (8-bit digits on the order-preserving uint32 image
// of the score, warp-aggregated histograms) followed by an ordered compaction of the ties and
// a bitonic sort of the K survivors: descending score, ties to the lower flat index
// ([SPEC] s6 == torch.sort(stable=True)).
#include "common.cuh"
#include "exact_math.cuh"
#include "topk_block.cuh"

namespace tspn {

// ---- embeddings: S = W2s relu(W1s x + b1s) + b2s, O likewise ---------------------------------------
// one CTA per tracklet; threads [0,H) do the subject branch hidden units, [H,2H) the object's.
__global__ void __launch_bounds__(128)
ppn_embed_kernel(const float* __restrict__ cls, int C, int H, const float* __restrict__ sw0,
                 const float* __restrict__ sb0, const float* __restrict__ sw2, const float* __restrict__ sb2,
                 const float* __restrict__ ow0, const float* __restrict__ ob0, const float* __restrict__ ow2,
                 const float* __restrict__ ob2, float* __restrict__ S, float* __restrict__ O) {
    extern __shared__ float sm[];
    float* x = sm;            // [C]
    float* hid = sm + C;      // [2][H]
    const int64_t trk = blockIdx.x;
    for (int i = threadIdx.x; i < C; i += blockDim.x) x[i] = cls[trk * C + i];
    __syncthreads();
    for (int u = threadIdx.x; u < 2 * H; u += blockDim.x) {
        const int br = u / H, j = u - br * H;
        const float* w = (br ? ow0 : sw0) + (int64_t)j * C;
        float acc = (br ? ob0 : sb0)[j];
        for (int i = 0; i < C; ++i) acc = __fmaf_rn(x[i], __ldg(w + i), acc);
        hid[u] = fmaxf(acc, 0.0f);
    }
    __syncthreads();
    for (int u = threadIdx.x; u < 2 * C; u += blockDim.x) {
        const int br = u / C, c = u - br * C;
        const float* w = (br ? ow2 : sw2) + (int64_t)c * H;
        const float* h = hid + br * H;
        float acc = (br ? ob2 : sb2)[c];
        for (int j = 0; j < H; ++j) acc = __fmaf_rn(h[j], __ldg(w + j), acc);
        (br ? O : S)[trk * C + c] = acc;
    }
}

// Tiled variant used when the four weight matrices fit in shared memory (C*H <= 12288): a CTA
// stages them once, transposed so that consecutive threads read consecutive words (the per-thread
// row walk of the kernel above is 32 cache lines per warp load), and embeds EMB_G tracklets.
// Every output is the same k-ascending fma chain -> same bits as ppn_embed_kernel.
#ifndef TSPN_EMB_G
#define TSPN_EMB_G 8
#endif
constexpr int EMB_G = TSPN_EMB_G;
constexpr int EMB_THREADS = 256;
constexpr int EMB_LD = 5;                  // staging: iterations whose loads (4 each) are in flight together
__global__ void __launch_bounds__(EMB_THREADS)
ppn_embed_tiled_kernel(const float* __restrict__ cls, int64_t n_trk, int C, int H, const float* __restrict__ sw0,
                       const float* __restrict__ sb0, const float* __restrict__ sw2,
                       const float* __restrict__ sb2, const float* __restrict__ ow0,
                       const float* __restrict__ ob0, const float* __restrict__ ow2,
                       const float* __restrict__ ob2, float* __restrict__ S, float* __restrict__ O) {
    extern __shared__ float sm[];
    // rows padded by one word: the transposing stores of the staging loop (consecutive threads write a column)
    // would otherwise all hit one bank (stride H = 64 words) or two (stride C = 80)
    const int HP = H + 1, CP = C + 1;
    float* w0t = sm;                      // [2][C][HP]  w0t[br][i][j] = W0_br[j][i]
    float* w2t = w0t + 2 * C * HP;        // [2][H][CP]  w2t[br][j][c] = W2_br[c][j]
    float* x = w2t + 2 * H * CP;          // [G][C]
    float* hid = x + EMB_G * C;           // [G][2][H]
    float* bias = hid + EMB_G * 2 * H;    // [sb0 (H) | ob0 (H) | sb2 (C) | ob2 (C)]
    const int tid = threadIdx.x;
    const int64_t trk0 = (int64_t)blockIdx.x * EMB_G;
    const int g_cnt = (int)min((int64_t)EMB_G, n_trk - trk0);
    // Staging.  Two things make this kernel's time under the all-pairs kernel (it heads the side chain): (1) index
    // arithmetic - the flat-index form spent three integer divisions per word staged; (row, column) now advance
    // incrementally; (2) memory round trips, several microseconds each there - every global load of the kernel is issued
    // here, in batches of EMB_LD x 4 independent branch-free loads per thread (was: one pair of loads per loop iteration,
    // plus one bias load at the head of every fma chain: ~50 exposed round trips per thread).
    const int CH = C * H;
    for (int e = tid; e < 2 * H + 2 * C; e += EMB_THREADS) {       // [sb0 | ob0 | sb2 | ob2]
        const float* src = e < H ? sb0 + e : e < 2 * H ? ob0 + (e - H) : e < 2 * H + C ? sb2 + (e - 2 * H)
                                                                                       : ob2 + (e - 2 * H - C);
        bias[e] = __ldg(src);
    }
    for (int e0 = tid; e0 < g_cnt * C; e0 += EMB_THREADS * 4) {
        float xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = __ldg(cls + trk0 * C + min(e0 + u * EMB_THREADS, g_cnt * C - 1));
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (e0 + u * EMB_THREADS < g_cnt * C) x[e0 + u * EMB_THREADS] = xv[u];
    }
    {
        // W0_br is [H][C] row-major: element rem = j * C + i  ->  w0t[br][i][j]
        // W2_br is [C][H] row-major: element rem = c * H + j2 ->  w2t[br][j2][c]
        const int dj = EMB_THREADS / C, di = EMB_THREADS - dj * C;
        const int dc = EMB_THREADS / H, dj2 = EMB_THREADS - dc * H;
        int j = tid / C, i = tid - j * C;
        int c = tid / H, j2 = tid - c * H;
        for (int rem0 = tid; rem0 < CH; rem0 += EMB_THREADS * EMB_LD) {
            float a0[EMB_LD], a1[EMB_LD], a2[EMB_LD], a3[EMB_LD];
#pragma unroll
            for (int u = 0; u < EMB_LD; ++u) {
                const int rc = min(rem0 + u * EMB_THREADS, CH - 1);
                a0[u] = __ldg(sw0 + rc);
                a1[u] = __ldg(ow0 + rc);
                a2[u] = __ldg(sw2 + rc);
                a3[u] = __ldg(ow2 + rc);
            }
#pragma unroll
            for (int u = 0; u < EMB_LD; ++u) {
                if (rem0 + u * EMB_THREADS < CH) {
                    w0t[i * HP + j] = a0[u];
                    w0t[C * HP + i * HP + j] = a1[u];
                    w2t[j2 * CP + c] = a2[u];
                    w2t[H * CP + j2 * CP + c] = a3[u];
                }
                j += dj;
                i += di;
                if (i >= C) { i -= C; ++j; }
                c += dc;
                j2 += dj2;
                if (j2 >= H) { j2 -= H; ++c; }
            }
        }
    }
    __syncthreads();
    for (int u = tid; u < g_cnt * 2 * H; u += EMB_THREADS) {
        const int g = u / (2 * H), j2 = u - g * 2 * H;
        const int br = j2 / H, j = j2 - br * H;
        const float* w = w0t + br * C * HP + j;
        const float* xg = x + g * C;
        float acc = bias[br * H + j];
        for (int i = 0; i < C; ++i) acc = __fmaf_rn(xg[i], w[i * HP], acc);
        hid[u] = fmaxf(acc, 0.0f);
    }
    __syncthreads();
    for (int u = tid; u < g_cnt * 2 * C; u += EMB_THREADS) {
        const int g = u / (2 * C), c2 = u - g * 2 * C;
        const int br = c2 / C, c = c2 - br * C;
        const float* w = w2t + br * H * CP + c;
        const float* h = hid + (g * 2 + br) * H;
        float acc = bias[2 * H + br * C + c];
        for (int j = 0; j < H; ++j) acc = __fmaf_rn(h[j], w[j * CP], acc);
        (br ? O : S)[(trk0 + g) * C + c] = acc;
    }
}

// ---- scores: M[s][o] = sigmoid(sum_c S[s][c] O[o][c]); one CTA per subject row ---------------------
__global__ void __launch_bounds__(128)
pair_scores_kernel(const int64_t* __restrict__ table, int nv, const float* __restrict__ S,
                   const float* __restrict__ O, int C, float* __restrict__ scores) {
    extern __shared__ float srow[];     // [C]
    const int64_t trk = blockIdx.x;
    if (trk >= table_total(table, nv, TSPN_VT_TRK_OFF)) return;      // the grid is sized for a capacity
    const int v = find_video(table, nv, TSPN_VT_TRK_OFF, trk);
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const int n = (int)row[TSPN_VT_N];
    const int64_t trk0 = row[TSPN_VT_TRK_OFF];
    const int s = (int)(trk - trk0);
    for (int i = threadIdx.x; i < C; i += blockDim.x) srow[i] = S[trk * C + i];
    __syncthreads();
    float* out = scores + row[TSPN_VT_SCORE_OFF] + (int64_t)s * n;
    for (int o = threadIdx.x; o < n; o += blockDim.x) {
        const float* orow = O + (trk0 + o) * C;
        float z = 0.0f;
        for (int c = 0; c < C; ++c) z = __fmaf_rn(srow[c], __ldg(orow + c), z);
        out[o] = sigmoid_det(z);
    }
}

// ---- top-K ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TOPK_THREADS)
topk_kernel(const int64_t* __restrict__ table, int nv, const float* __restrict__ scores, int K, int exclude_diag,
            int64_t* __restrict__ out_idx, float* __restrict__ out_score, int64_t* __restrict__ out_row) {
    __shared__ TopkSmem sm;
    const int v = blockIdx.x;
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const uint32_t n = (uint32_t)row[TSPN_VT_N];
    const int64_t total = (int64_t)n * n;
    const float* sc = scores + row[TSPN_VT_SCORE_OFF];
    const int tid = threadIdx.x;
    const float NEG_INF = __uint_as_float(0xff800000u);

    // the diagonal (flat index a multiple of n + 1) is not a candidate.  Divisibility without a division per candidate
    // (Lemire & Kaser: for 32-bit u, d:  u % d == 0  <=>  u * M <= M - 1 in 64-bit arithmetic, M = 2^64 / d rounded up)
    const uint64_t diag_m = ~0ull / (uint64_t)(n + 1u) + 1ull;
    const int k_eff = block_topk<32, 16>(sm, total, K, [&](int64_t i) -> float {     // big videos: few, deep round trips
        const float val = __ldg(sc + i);                             // (load first: no branch around it)
        return (exclude_diag && (uint64_t)(uint32_t)i * diag_m <= diag_m - 1ull) ? NEG_INF : val;
    });

    int64_t* oi = out_idx + (int64_t)v * K;
    float* os = out_score + (int64_t)v * K;
    int64_t* orow = out_row ? out_row + (int64_t)v * K : nullptr;
    for (int i = tid; i < K; i += TOPK_THREADS) {
        if (i < k_eff) {
            const uint32_t flat = (uint32_t)(sm.sel[i] & 0xffffffffu);
            oi[i] = (int64_t)flat;
            os[i] = __ldg(sc + flat);
            if (orow) {
                const int s = (int)(flat / n), o = (int)(flat % n);
                orow[i] = (s == o) ? -1 : row[TSPN_VT_PAIR_OFF] + (int64_t)s * (n - 1) + o - (o > s ? 1 : 0);
            }
        } else {
            oi[i] = -1;
            os[i] = 0.0f;
            if (orow) orow[i] = -1;
        }
    }
}

// ---- scores + top-K of one video in one CTA ------------------------------------------------------------
// The video's subject and object embeddings are staged in shared memory (rows padded by one word: thread i
// walks row o = i % n conflict-free, row s = i / n is a broadcast), every score is computed once - the same
// c-ascending fma chain as pair_scores_kernel - written to the scores tensor and handed to the block top-K as
// the key it caches.  One launch and one pass over the embeddings instead of two kernels with the scores going
// through memory in between: this pair sits on the chain that decides when the surviving-pair kernel can start.
// Needs n * n <= TOPK_CACHE (block_topk then evaluates every candidate exactly once).
__global__ void __launch_bounds__(TOPK_THREADS, 5)
scores_topk_kernel(const int64_t* __restrict__ table, int nv, const float* __restrict__ S,
                   const float* __restrict__ O, int C, float* __restrict__ scores, int K, int exclude_diag,
                   int64_t* __restrict__ out_idx, float* __restrict__ out_score, int64_t* __restrict__ out_row) {
    __shared__ TopkSmem sm;
    extern __shared__ float emb[];      // [2][n][C + 1]
    const int v = blockIdx.x;
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const uint32_t n = (uint32_t)row[TSPN_VT_N];
    const int64_t trk0 = row[TSPN_VT_TRK_OFF];
    const int64_t total = (int64_t)n * n;
    const int tid = threadIdx.x;
    const int CP = C + 1;
    float* const ss = emb;
    float* const oo = emb + (size_t)n * CP;
    const int n_emb = (int)n * C;
    for (int e0 = tid; e0 < n_emb; e0 += TOPK_THREADS * 4) {       // eight independent loads in flight per thread
        float a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int ec = min(e0 + u * TOPK_THREADS, n_emb - 1);
            a[u] = S[trk0 * C + ec];
            b[u] = O[trk0 * C + ec];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * TOPK_THREADS;
            if (e < n_emb) {
                const int r = e / C, c = e - r * C;
                ss[r * CP + c] = a[u];
                oo[r * CP + c] = b[u];
            }
        }
    }
    __syncthreads();
    float* const sc = scores + row[TSPN_VT_SCORE_OFF];
    const float NEG_INF = __uint_as_float(0xff800000u);

    const int k_eff = block_topk<1, 4>(sm, total, K, [&](int64_t i) -> float {       // n * n <= TOPK_CACHE: always cached
        const uint32_t u = (uint32_t)i;
        const uint32_t s = u / n, o = u - s * n;
        const float* sr = ss + s * CP;
        const float* orow = oo + o * CP;
        float z = 0.0f;
        for (int c = 0; c < C; ++c) z = __fmaf_rn(sr[c], orow[c], z);
        const float val = sigmoid_det(z);
        sc[i] = val;
        return (exclude_diag && s == o) ? NEG_INF : val;            // the diagonal is scored but not a candidate
    });

    int64_t* oi = out_idx + (int64_t)v * K;
    float* os = out_score + (int64_t)v * K;
    int64_t* orow = out_row ? out_row + (int64_t)v * K : nullptr;
    for (int i = tid; i < K; i += TOPK_THREADS) {
        if (i < k_eff) {
            const uint32_t flat = (uint32_t)(sm.sel[i] & 0xffffffffu);
            oi[i] = (int64_t)flat;
            os[i] = sc[flat];                   // written by this CTA before block_topk's barriers
            if (orow) {
                const int s = (int)(flat / n), o = (int)(flat % n);
                orow[i] = (s == o) ? -1 : row[TSPN_VT_PAIR_OFF] + (int64_t)s * (n - 1) + o - (o > s ? 1 : 0);
            }
        } else {
            oi[i] = -1;
            os[i] = 0.0f;
            if (orow) orow[i] = -1;
        }
    }
}

static int launch_embeddings(int64_t total_tracklets, const float* d_cls, int n_classes, int hidden,
                             const float* d_sub_w0, const float* d_sub_b0, const float* d_sub_w2, const float* d_sub_b2,
                             const float* d_obj_w0, const float* d_obj_b0, const float* d_obj_w2, const float* d_obj_b2,
                             float* S, float* O, cudaStream_t st) {
    if ((int64_t)n_classes * hidden <= 12288) {
        const size_t smt = (size_t)(2 * n_classes * (hidden + 1) + 2 * hidden * (n_classes + 1) + EMB_G * n_classes +
                                    EMB_G * 2 * hidden + 2 * hidden + 2 * n_classes) * sizeof(float);
        TSPN_CUDA_OK(cudaFuncSetAttribute(ppn_embed_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smt));
        prefer_max_smem(ppn_embed_tiled_kernel);
        ppn_embed_tiled_kernel<<<(unsigned)((total_tracklets + EMB_G - 1) / EMB_G), EMB_THREADS, smt, st>>>(
            d_cls, total_tracklets, n_classes, hidden, d_sub_w0, d_sub_b0, d_sub_w2, d_sub_b2, d_obj_w0, d_obj_b0,
            d_obj_w2, d_obj_b2, S, O);
    } else {
        const size_t sm1 = (size_t)(n_classes + 2 * hidden) * sizeof(float);
        prefer_max_smem(ppn_embed_kernel);
        ppn_embed_kernel<<<(unsigned)total_tracklets, 128, sm1, st>>>(d_cls, n_classes, hidden, d_sub_w0, d_sub_b0,
                                                                      d_sub_w2, d_sub_b2, d_obj_w0, d_obj_b0,
                                                                      d_obj_w2, d_obj_b2, S, O);
    }
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

// relationness_tc.cu: the tensor-core form (TSPN_PREC_TENSOR)
int relationness_tc_supported(int max_tracklets, int n_classes, int hidden);
int launch_relationness_tc(const int64_t* d_table, int num_videos, int64_t total_tracklets, int max_tracklets,
                           const float* d_cls, int C, int H, const float* sw0, const float* sb0, const float* sw2,
                           const float* sb2, const float* ow0, const float* ob0, const float* ow2, const float* ob2,
                           float* S, float* O, float* d_scores, int k, int exclude_diag, int64_t* d_idx,
                           float* d_val, int64_t* d_row, bool* fused_topk, cudaStream_t st);

}  // namespace tspn

using namespace tspn;

extern "C" {

int64_t tspn_relationness_workspace_bytes(int64_t total_tracklets, int n_classes, int hidden) {
    (void)hidden;
    const int64_t t = total_tracklets > 0 ? total_tracklets : 1;
    return 2 * t * (int64_t)n_classes * (int64_t)sizeof(float);
}

int tspn_relationness_tc_supported(int max_tracklets, int n_classes, int hidden) {
    return relationness_tc_supported(max_tracklets, n_classes, hidden);
}

int tspn_relationness(const int64_t* d_table, int num_videos, int64_t total_tracklets, int max_tracklets,
                      const float* d_cls, int n_classes, int hidden, const float* d_sub_w0, const float* d_sub_b0,
                      const float* d_sub_w2, const float* d_sub_b2, const float* d_obj_w0, const float* d_obj_b0,
                      const float* d_obj_w2, const float* d_obj_b2, float* d_scores, int precision,
                      void* d_workspace, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && total_tracklets >= 0 && n_classes > 0 && hidden > 0, TSPN_EBADARG,
                 "tspn_relationness: bad size");
    if (total_tracklets == 0) return TSPN_OK;
    TSPN_REQUIRE(d_table && d_cls && d_sub_w0 && d_sub_b0 && d_sub_w2 && d_sub_b2 && d_obj_w0 && d_obj_b0 &&
                     d_obj_w2 && d_obj_b2 && d_scores && d_workspace,
                 TSPN_EBADARG, "tspn_relationness: null pointer");
    TSPN_REQUIRE(n_classes <= 4096 && hidden <= 4096, TSPN_ESHAPE, "tspn_relationness: C/H too large");
    cudaStream_t st = (cudaStream_t)stream;
    float* S = reinterpret_cast<float*>(d_workspace);
    float* O = S + total_tracklets * n_classes;
    if (precision == TSPN_PREC_TENSOR) {
        TSPN_REQUIRE(relationness_tc_supported(max_tracklets, n_classes, hidden), TSPN_ESHAPE,
                     "tspn_relationness: tensor precision needs N <= 256, C <= 128, H <= 128 and H %% 8 == 0 "
                     "(got N=%d C=%d H=%d)", max_tracklets, n_classes, hidden);
        bool fused = false;
        return launch_relationness_tc(d_table, num_videos, total_tracklets, max_tracklets, d_cls, n_classes, hidden,
                                      d_sub_w0, d_sub_b0, d_sub_w2, d_sub_b2, d_obj_w0, d_obj_b0, d_obj_w2, d_obj_b2, S,
                                      O, d_scores, 0, 0, nullptr, nullptr, nullptr, &fused, st);
    }
    {
        const int rc = launch_embeddings(total_tracklets, d_cls, n_classes, hidden, d_sub_w0, d_sub_b0, d_sub_w2, d_sub_b2,
                                         d_obj_w0, d_obj_b0, d_obj_w2, d_obj_b2, S, O, st);
        if (rc != TSPN_OK) return rc;
    }
    prefer_max_smem(pair_scores_kernel);
    pair_scores_kernel<<<(unsigned)total_tracklets, 128, (size_t)n_classes * sizeof(float), st>>>(
        d_table, num_videos, S, O, n_classes, d_scores);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_relationness_topk_supported(int max_tracklets, int n_classes) {
    return max_tracklets > 0 && (int64_t)max_tracklets * max_tracklets <= TOPK_CACHE &&
           2 * (int64_t)max_tracklets * (n_classes + 1) * 4 <= 160 * 1024;
}

int tspn_relationness_topk(const int64_t* d_table, int num_videos, int64_t total_tracklets, int max_tracklets,
                           const float* d_cls, int n_classes, int hidden, const float* d_sub_w0,
                           const float* d_sub_b0, const float* d_sub_w2, const float* d_sub_b2, const float* d_obj_w0,
                           const float* d_obj_b0, const float* d_obj_w2, const float* d_obj_b2, float* d_scores, int k,
                           int flags, int precision, int64_t* d_topk_idx, float* d_topk_score, int64_t* d_topk_row,
                           void* d_workspace, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && total_tracklets >= 0 && n_classes > 0 && hidden > 0 && k >= 0, TSPN_EBADARG,
                 "tspn_relationness_topk: bad size");
    TSPN_REQUIRE(k <= TOPK_MAX_K, TSPN_ESHAPE, "tspn_relationness_topk: K=%d exceeds the supported maximum %d", k,
                 TOPK_MAX_K);
    if (num_videos == 0) return TSPN_OK;
    if (precision == TSPN_PREC_TENSOR) {
        // tcgen05 form: embeddings, then scores (+ top-K in the same CTA when every video fits one: N <= 128;
        // larger videos: the block top-K as a third launch on the stored scores)
        TSPN_REQUIRE(relationness_tc_supported(max_tracklets, n_classes, hidden), TSPN_ESHAPE,
                     "tspn_relationness_topk: tensor precision needs N <= 256, C <= 128, H <= 128 and H %% 8 == 0 "
                     "(got N=%d C=%d H=%d)", max_tracklets, n_classes, hidden);
        TSPN_REQUIRE(d_table && d_workspace && d_topk_idx && d_topk_score && (total_tracklets == 0 || (d_cls && d_scores)),
                     TSPN_EBADARG, "tspn_relationness_topk: null pointer");
        cudaStream_t st = (cudaStream_t)stream;
        float* S = reinterpret_cast<float*>(d_workspace);
        float* O = S + total_tracklets * n_classes;
        bool fused = false;
        if (total_tracklets > 0) {
            const int rc = launch_relationness_tc(d_table, num_videos, total_tracklets, max_tracklets, d_cls, n_classes,
                                                  hidden, d_sub_w0, d_sub_b0, d_sub_w2, d_sub_b2, d_obj_w0, d_obj_b0,
                                                  d_obj_w2, d_obj_b2, S, O, d_scores, k,
                                                  (flags & TSPN_TOPK_EXCLUDE_DIAGONAL) ? 1 : 0, d_topk_idx, d_topk_score,
                                                  d_topk_row, &fused, st);
            if (rc != TSPN_OK) return rc;
        }
        if (!fused && k > 0)
            return tspn_topk_pairs(d_table, num_videos, d_scores, k, flags, d_topk_idx, d_topk_score, d_topk_row, stream);
        return TSPN_OK;
    }
    TSPN_REQUIRE(tspn_relationness_topk_supported(max_tracklets, n_classes), TSPN_ESHAPE,
                 "tspn_relationness_topk: max_tracklets=%d not supported (use tspn_relationness + tspn_topk_pairs)",
                 max_tracklets);
    TSPN_REQUIRE(d_table && d_sub_w0 && d_sub_b0 && d_sub_w2 && d_sub_b2 && d_obj_w0 && d_obj_b0 && d_obj_w2 && d_obj_b2 &&
                     d_workspace && d_topk_idx && d_topk_score && (total_tracklets == 0 || (d_cls && d_scores)),
                 TSPN_EBADARG, "tspn_relationness_topk: null pointer");
    TSPN_REQUIRE(n_classes <= 4096 && hidden <= 4096, TSPN_ESHAPE, "tspn_relationness_topk: C/H too large");
    cudaStream_t st = (cudaStream_t)stream;
    float* S = reinterpret_cast<float*>(d_workspace);
    float* O = S + total_tracklets * n_classes;
    if (total_tracklets > 0) {
        const int rc = launch_embeddings(total_tracklets, d_cls, n_classes, hidden, d_sub_w0, d_sub_b0, d_sub_w2, d_sub_b2,
                                         d_obj_w0, d_obj_b0, d_obj_w2, d_obj_b2, S, O, st);
        if (rc != TSPN_OK) return rc;
    }
    if (k == 0) return TSPN_OK;
    const size_t smem = 2 * (size_t)max_tracklets * (n_classes + 1) * sizeof(float);
    TSPN_CUDA_OK(cudaFuncSetAttribute(scores_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prefer_max_smem(scores_topk_kernel);
    scores_topk_kernel<<<(unsigned)num_videos, TOPK_THREADS, smem, st>>>(
        d_table, num_videos, S, O, n_classes, d_scores, k, (flags & TSPN_TOPK_EXCLUDE_DIAGONAL) ? 1 : 0, d_topk_idx,
        d_topk_score, d_topk_row);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_topk_pairs(const int64_t* d_table, int num_videos, const float* d_scores, int k, int flags,
                    int64_t* d_topk_idx, float* d_topk_score, int64_t* d_topk_row, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && k >= 0, TSPN_EBADARG, "tspn_topk_pairs: negative size");
    TSPN_REQUIRE(k <= TOPK_MAX_K, TSPN_ESHAPE, "tspn_topk_pairs: K=%d exceeds the supported maximum %d", k,
                 TOPK_MAX_K);
    if (num_videos == 0 || k == 0) return TSPN_OK;
    TSPN_REQUIRE(d_table && d_scores && d_topk_idx && d_topk_score, TSPN_EBADARG, "tspn_topk_pairs: null pointer");
    prefer_max_smem(topk_kernel);
    topk_kernel<<<(unsigned)num_videos, TOPK_THREADS, 0, (cudaStream_t)stream>>>(
        d_table, num_videos, d_scores, k, (flags & TSPN_TOPK_EXCLUDE_DIAGONAL) ? 1 : 0, d_topk_idx, d_topk_score,
        d_topk_row);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // extern "C"
