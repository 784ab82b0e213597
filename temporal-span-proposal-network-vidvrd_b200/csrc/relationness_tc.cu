// relationness_tc.cu — PPNHead on the tensor cores (TSPN_PREC_TENSOR).
//
//   S = W2s relu(W1s x + b1s) + b2s,  O likewise,  M = sigmoid(S O^T)      lib/modeling/relpn/ppn.py:92-112
//
// All three contractions run as tcgen05.mma with fp32 accumulators in TMEM.  Operands are the fp32 values
// themselves fed as kind::tf32 (10 explicit mantissa bits): the data is a few KB per video, so operand width buys
// nothing, and bf16 operands - measured in a bit-accurate simulation on the gain-4 synthetic weights the parity
// tests use - reach 1.4e-2 absolute on the scores, above the 1e-2 the north star allows; tf32 stays below 2e-3.
// (The header defines TSPN_PREC_TENSOR as "bf16, or tf32 on fp32 storage".)
//
//   ppn_embed_tc_kernel    one CTA per 128 consecutive tracklets: X [128, C] x [W1s; W1o]^T -> [128, 2H] in TMEM
//                          (one MMA chain for both branches), epilogue bias + ReLU back into shared memory as the
//                          A operands of the second layer, two MMA chains -> S, O [128, C] in TMEM, epilogue bias ->
//                          the same fp32 workspace rows the exact-order kernels write.
//   scores_tc_kernel       one CTA per (video, 128 subject rows): S_v O_v^T -> [128, N] in TMEM, epilogue sigmoid ->
//                          scores; when the video fits one CTA (N <= 128) the block top-K of topk_block.cuh runs in
//                          the same CTA on the scores it has just written.
//
// Operand staging is by the CTA's threads (not TMA): they write the canonical SWIZZLE_128B K-major layout - rows
// of 128 bytes (32 fp32 of K), 8-row atoms of 1 KB, the 16-byte chunk index XORed with the row index mod 8 - then
// fence.proxy.async before the MMA reads it.  Not shaped for co-residency with the all-pairs kernel (48-96 KB of
// operands per CTA): under it these CTAs land on the SMs the persistent pair kernel leaves free.
#include "exact_math.cuh"
#include "tc_common.cuh"
#include "topk_block.cuh"

namespace tspn {

constexpr int RT_M = 128;                   // UMMA M: tracklet rows per CTA
constexpr int RT_SLAB_K = 32;               // fp32 elements per 128-byte slab row
constexpr int RT_EMB_THREADS = 128;

__host__ __device__ __forceinline__ int rt_round_up(int a, int b) { return (a + b - 1) / b * b; }

// byte offset of element (row r, k) of a K-major SWIZZLE_128B operand whose slabs are `rows` rows high
__device__ __forceinline__ uint32_t rt_off(int rows, int r, int k) {
    const int slab = k >> 5, kk = k & 31;
    return (uint32_t)slab * (uint32_t)rows * 128u + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u +
           (uint32_t)(((kk >> 2) ^ (r & 7)) << 4) + (uint32_t)(kk & 3) * 4u;
}

// one MMA chain: D[128, n] (+)= A[128, kp] B[n, kp]^T, kp a multiple of 8, operands as laid out by rt_off
__device__ __forceinline__ void rt_mma_chain(uint32_t tmem_d, uint32_t a_addr, int a_rows, uint32_t b_addr, int b_rows,
                                             int kp, uint32_t idesc) {
    for (int ks = 0; ks < kp / 8; ++ks) {
        const int slab = ks >> 2, k4 = ks & 3;
        const uint64_t adesc = umma_smem_desc(a_addr + (uint32_t)slab * (uint32_t)a_rows * 128u, 16, 1024) + 2 * k4;
        const uint64_t bdesc = umma_smem_desc(b_addr + (uint32_t)slab * (uint32_t)b_rows * 128u, 16, 1024) + 2 * k4;
        umma_tf32(tmem_d, adesc, bdesc, idesc, ks != 0);
    }
}

static inline size_t rt_embed_smem(int c, int h) {
    const int kp1 = rt_round_up(c, 8), sl_c = (kp1 + 31) / 32, sl_h = (h + 31) / 32, n1 = 2 * h, n2 = rt_round_up(c, 16);
    const size_t x = (size_t)sl_c * RT_M * 128, w1 = (size_t)sl_c * n1 * 128;
    const size_t a2 = 2 * (size_t)sl_h * RT_M * 128;                       // hidden tiles overlay X | W1
    const size_t front = x + w1 > a2 ? x + w1 : a2;
    const size_t w2 = 2 * (size_t)sl_h * n2 * 128;
    return front + w2 + 1024 /* alignment slack */ + 64;
}

__global__ void __launch_bounds__(RT_EMB_THREADS)
ppn_embed_tc_kernel(const int64_t* __restrict__ table, int nv, const float* __restrict__ cls, int C, int H,
                    const float* __restrict__ sw0, const float* __restrict__ sb0, const float* __restrict__ sw2,
                    const float* __restrict__ sb2, const float* __restrict__ ow0, const float* __restrict__ ob0,
                    const float* __restrict__ ow2, const float* __restrict__ ob2, float* __restrict__ S,
                    float* __restrict__ O) {
    extern __shared__ uint8_t rt_smem_raw[];
    uint8_t* const smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(rt_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int64_t n_trk = table_total(table, nv, TSPN_VT_TRK_OFF);
    const int64_t row0 = (int64_t)blockIdx.x * RT_M;
    if (row0 >= n_trk) return;                                    // the grid is sized for a capacity
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kp1 = rt_round_up(C, 8), sl_c = (kp1 + 31) / 32, sl_h = (H + 31) / 32;
    const int n1 = 2 * H, n2 = rt_round_up(C, 16);
    uint8_t* const x_t = smem;                                    // [128][kp1]
    uint8_t* const w1_t = x_t + (size_t)sl_c * RT_M * 128;        // [2H][kp1]: rows [0,H) = W1s, [H,2H) = W1o
    uint8_t* const a2_t = smem;                                   // [2][128][H], overlays x_t | w1_t after layer 1
    const size_t a2_bytes = 2 * (size_t)sl_h * RT_M * 128, xw_bytes = (size_t)sl_c * (RT_M + n1) * 128;
    uint8_t* const w2_t = smem + (a2_bytes > xw_bytes ? a2_bytes : xw_bytes);     // [2][n2][H]
    uint64_t* const bar = reinterpret_cast<uint64_t*>(w2_t + 2 * (size_t)sl_h * n2 * 128);
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    // ---- stage X, [W1s; W1o], W2s, W2o (zero K / row padding) -------------------------------------
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = tid; i < xw_bytes / 16; i += RT_EMB_THREADS) reinterpret_cast<float4*>(smem)[i] = z4;
    for (size_t i = tid; i < 2 * (size_t)sl_h * n2 * 128 / 16; i += RT_EMB_THREADS) reinterpret_cast<float4*>(w2_t)[i] = z4;
    __syncthreads();
    const int rows_here = (int)min((int64_t)RT_M, n_trk - row0);
    for (int e = tid; e < rows_here * C; e += RT_EMB_THREADS) {
        const int r = e / C, k = e - r * C;
        *reinterpret_cast<float*>(x_t + rt_off(RT_M, r, k)) = __ldg(cls + (row0 + r) * C + k);
    }
    for (int e = tid; e < n1 * C; e += RT_EMB_THREADS) {
        const int j = e / C, k = e - j * C;
        *reinterpret_cast<float*>(w1_t + rt_off(n1, j, k)) = j < H ? __ldg(sw0 + j * C + k) : __ldg(ow0 + (j - H) * C + k);
    }
    for (int e = tid; e < 2 * C * H; e += RT_EMB_THREADS) {
        const int br = e / (C * H), rem = e - br * C * H, c = rem / H, k = rem - c * H;
        *reinterpret_cast<float*>(w2_t + (size_t)br * sl_h * n2 * 128 + rt_off(n2, c, k)) = __ldg((br ? ow2 : sw2) + rem);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // ---- layer 1: [128, 2H] = X [W1s; W1o]^T ------------------------------------------------------
    if (tid == 0) {
        rt_mma_chain(tmem, smem_u32(x_t), RT_M, smem_u32(w1_t), n1, kp1, umma_idesc(UMMA_FMT_TF32, RT_M, (uint32_t)n1, 0, 0));
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    {   // bias + ReLU -> hidden tiles (A operands of layer 2); x_t / w1_t are dead: the MMAs that read them are done
        const int r = warp * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < n1; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int col = c0 + j, br = col >= H ? 1 : 0, u = col - br * H;
                const float hval = fmaxf(v[j] + __ldg((br ? ob0 : sb0) + u), 0.0f);
                *reinterpret_cast<float*>(a2_t + (size_t)br * sl_h * RT_M * 128 + rt_off(RT_M, r, u)) = hval;
            }
        }
    }
    // K padding of the hidden tiles (H not a multiple of 32 leaves stale bytes of X / W1 in the last slab)
    if (H & 31) {
        for (int e = tid; e < 2 * RT_M * (sl_h * 32 - H); e += RT_EMB_THREADS) {
            const int br = e / (RT_M * (sl_h * 32 - H)), rem = e - br * RT_M * (sl_h * 32 - H);
            const int r = rem / (sl_h * 32 - H), k = H + rem - r * (sl_h * 32 - H);
            *reinterpret_cast<float*>(a2_t + (size_t)br * sl_h * RT_M * 128 + rt_off(RT_M, r, k)) = 0.0f;
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- layer 2: S = hid_s W2s^T -> TMEM columns [0, n2), O = hid_o W2o^T -> [128, 128 + n2) ------------
    if (tid == 0) {
        const uint32_t idesc = umma_idesc(UMMA_FMT_TF32, RT_M, (uint32_t)n2, 0, 0);
        rt_mma_chain(tmem, smem_u32(a2_t), RT_M, smem_u32(w2_t), n2, H, idesc);
        rt_mma_chain(tmem + 128u, smem_u32(a2_t + (size_t)sl_h * RT_M * 128), RT_M,
                     smem_u32(w2_t + (size_t)sl_h * n2 * 128), n2, H, idesc);
        umma_commit(bar);
    }
    mbar_wait(bar, 1);
    tc_fence_after();
    {
        const int r = warp * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int br = 0; br < 2; ++br) {
            float* dst = (br ? O : S) + (row0 + r) * C;
            const float* b2 = br ? ob2 : sb2;
            for (int c0 = 0; c0 < n2; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)(br * 128 + c0), v);
                if (r < rows_here) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < C) dst[c0 + j] = v[j] + __ldg(b2 + c0 + j);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

static inline size_t rt_scores_smem(int c, int max_n) {
    const int kp = rt_round_up(c, 8), sl = (kp + 31) / 32, npad = rt_round_up(max_n, 16);
    return (size_t)sl * (RT_M + npad) * 128 + 1024 + 64;
}
__host__ __device__ __forceinline__ uint32_t rt_pow2_cols(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

// scores of one video (+ its top-K when the video fits one CTA).  grid = (videos, ceil(max N / 128)).
template <bool FUSED_TOPK>
__global__ void __launch_bounds__(TOPK_THREADS)
scores_tc_kernel(const int64_t* __restrict__ table, int nv, const float* __restrict__ S, const float* __restrict__ O,
                 int C, int npad_max, float* __restrict__ scores, int K, int exclude_diag,
                 int64_t* __restrict__ out_idx, float* __restrict__ out_score, int64_t* __restrict__ out_row) {
    __shared__ TopkSmem sm;
    extern __shared__ uint8_t rt_smem_raw[];
    uint8_t* const smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(rt_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int v = blockIdx.x;
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const int n = (int)row[TSPN_VT_N];
    const int m0 = blockIdx.y * RT_M;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t trk0 = row[TSPN_VT_TRK_OFF];
    float* const sc = scores + row[TSPN_VT_SCORE_OFF];
    const int kp = rt_round_up(C, 8), sl = (kp + 31) / 32;
    const int npad = rt_round_up(n, 16);
    uint8_t* const a_t = smem;                                    // [128][kp]: subject rows m0 ..
    uint8_t* const b_t = a_t + (size_t)sl * RT_M * 128;           // [npad][kp]: every object row of the video
    uint64_t* const bar = reinterpret_cast<uint64_t*>(b_t + (size_t)sl * rt_round_up(npad_max, 16) * 128);
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const bool active = m0 < n;                                   // uniform per CTA
    const uint32_t cols = rt_pow2_cols(npad);
    if (active) {
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        if (warp == 0) {
            tmem_alloc(tmem_slot, cols);
            tmem_relinquish();
        }
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (size_t i = tid; i < (size_t)sl * (RT_M + npad) * 128 / 16; i += TOPK_THREADS) {
            // A and B are contiguous only when npad == npad_max: zero them separately
            const size_t a_vec = (size_t)sl * RT_M * 128 / 16;
            if (i < a_vec) reinterpret_cast<float4*>(a_t)[i] = z4;
            else reinterpret_cast<float4*>(b_t)[i - a_vec] = z4;
        }
        __syncthreads();
        const int rows_a = min(RT_M, n - m0);
        for (int e = tid; e < rows_a * C; e += TOPK_THREADS) {
            const int r = e / C, k = e - r * C;
            *reinterpret_cast<float*>(a_t + rt_off(RT_M, r, k)) = S[(trk0 + m0 + r) * C + k];
        }
        for (int e = tid; e < n * C; e += TOPK_THREADS) {
            const int r = e / C, k = e - r * C;
            *reinterpret_cast<float*>(b_t + rt_off(npad, r, k)) = O[(trk0 + r) * C + k];
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        const uint32_t tmem = *tmem_slot;
        if (tid == 0) {
            rt_mma_chain(tmem, smem_u32(a_t), RT_M, smem_u32(b_t), npad, kp, umma_idesc(UMMA_FMT_TF32, RT_M, (uint32_t)npad, 0, 0));
            umma_commit(bar);
        }
        mbar_wait(bar, 0);
        tc_fence_after();
        if (warp < 4) {
            const int r = m0 + warp * 32 + lane;
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < npad; c0 += 16) {
                float val[16];
                tmem_ld16(taddr + (uint32_t)c0, val);
                if (r < n) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < n) sc[(int64_t)r * n + c0 + j] = 1.0f / (1.0f + __expf(-val[j]));
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            tmem_dealloc(tmem, cols);
        }
    }
    if (!FUSED_TOPK) return;
    // ---- top-K of the scores this CTA has just written (the video fits one CTA: gridDim.y == 1) ----
    __syncthreads();
    const int64_t total = (int64_t)n * n;
    const float NEG_INF = __uint_as_float(0xff800000u);
    const uint32_t un = (uint32_t)n;
    const int k_eff = block_topk(sm, total, K, [&](int64_t i) -> float {
        const uint32_t u = (uint32_t)i;
        if (exclude_diag && (u / un) == (u % un)) return NEG_INF;
        return sc[i];
    });
    int64_t* oi = out_idx + (int64_t)v * K;
    float* os = out_score + (int64_t)v * K;
    int64_t* orow = out_row ? out_row + (int64_t)v * K : nullptr;
    for (int i = tid; i < K; i += TOPK_THREADS) {
        if (i < k_eff) {
            const uint32_t flat = (uint32_t)(sm.sel[i] & 0xffffffffu);
            oi[i] = (int64_t)flat;
            os[i] = sc[flat];
            if (orow) {
                const int s = (int)(flat / un), o = (int)(flat % un);
                orow[i] = (s == o) ? -1 : row[TSPN_VT_PAIR_OFF] + (int64_t)s * (n - 1) + o - (o > s ? 1 : 0);
            }
        } else {
            oi[i] = -1;
            os[i] = 0.0f;
            if (orow) orow[i] = -1;
        }
    }
}

// host: launch the two kernels (called by relationness.cu)
int relationness_tc_supported(int max_tracklets, int n_classes, int hidden) {
    return max_tracklets > 0 && max_tracklets <= 256 && n_classes > 0 && n_classes <= 128 && hidden > 0 &&
           hidden <= 128 && hidden % 8 == 0 && rt_embed_smem(n_classes, hidden) <= 200 * 1024 &&
           rt_scores_smem(n_classes, max_tracklets) <= 160 * 1024;
}

int launch_relationness_tc(const int64_t* d_table, int num_videos, int64_t total_tracklets, int max_tracklets,
                           const float* d_cls, int C, int H, const float* sw0, const float* sb0, const float* sw2,
                           const float* sb2, const float* ow0, const float* ob0, const float* ow2, const float* ob2,
                           float* S, float* O, float* d_scores, int k, int exclude_diag, int64_t* d_idx,
                           float* d_val, int64_t* d_row, bool* fused_topk, cudaStream_t st) {
    const size_t sm1 = rt_embed_smem(C, H);
    TSPN_CUDA_OK(cudaFuncSetAttribute(ppn_embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
    prefer_max_smem(ppn_embed_tc_kernel);
    ppn_embed_tc_kernel<<<(unsigned)((total_tracklets + RT_M - 1) / RT_M), RT_EMB_THREADS, sm1, st>>>(
        d_table, num_videos, d_cls, C, H, sw0, sb0, sw2, sb2, ow0, ob0, ow2, ob2, S, O);
    TSPN_CUDA_OK(cudaGetLastError());
    const size_t sm2 = rt_scores_smem(C, max_tracklets);
    const int m_tiles = (max_tracklets + RT_M - 1) / RT_M;
    const bool fuse = k > 0 && m_tiles == 1 && d_idx != nullptr;
    *fused_topk = fuse;
    dim3 grid((unsigned)num_videos, (unsigned)m_tiles);
    if (fuse) {
        TSPN_CUDA_OK(cudaFuncSetAttribute(scores_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        prefer_max_smem(scores_tc_kernel<true>);
        scores_tc_kernel<true><<<grid, TOPK_THREADS, sm2, st>>>(d_table, num_videos, S, O, C, max_tracklets, d_scores, k,
                                                                exclude_diag, d_idx, d_val, d_row);
    } else {
        TSPN_CUDA_OK(cudaFuncSetAttribute(scores_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        prefer_max_smem(scores_tc_kernel<false>);
        scores_tc_kernel<false><<<grid, TOPK_THREADS, sm2, st>>>(d_table, num_videos, S, O, C, max_tracklets, d_scores, 0,
                                                                 0, nullptr, nullptr, nullptr);
    }
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // namespace tspn
