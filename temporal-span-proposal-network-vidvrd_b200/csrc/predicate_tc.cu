// predicate_tc.cu — tensor-core predicate classifier (TSPN_PREC_TENSOR).
//
//   y = sigmoid(x W^T + b)      RelationPredictor.forward, lib/modeling/model.py:76-88
//
// The op reads 4*F bytes per row and does 2*F*R flops on them (R = 50 / 132): 25-66 flop/B,
// far below the B200 ridge, so it is HBM-bound on x; the tensor pipe only has to keep up.
//   * x [m, F] (fp32 -> kind::tf32 on the fp32 storage, or bf16 -> kind::f16) and the packed
//     W [Rpad, Fpad] are streamed by TMA (SWIZZLE_128B, 128-byte K slabs) through a ring of
//     mbarrier-guarded stages; OOB rows/columns are zero-filled by TMA, nothing is padded in HBM;
//   * one elected thread issues tcgen05.mma (M=128, N=Rpad, K=32 B per instruction),
//     accumulators live in TMEM (Rpad fp32 columns);
//   * four epilogue warps read TMEM (tcgen05.ld 32x32b), add the bias, apply the sigmoid and
//     store rows; with split-K (needed to fill 148 SMs when m/128 is small) they store partials
//     and a second tiny kernel sums the splits in a fixed order.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue.
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace tspn {

constexpr int PT_BM = 128;                // rows per CTA tile (UMMA M)
constexpr int PT_SLAB = 128;              // bytes of K per stage row (one swizzle atom row)
constexpr int PT_THREADS = 192;
constexpr int PT_MAX_SPLITS = 16;
constexpr int PT_SMEM_BUDGET = 200 * 1024;

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline int pt_rpad(int r) { return round_up(r, 16); }
static inline int pt_kpad(int f, int elem) { return round_up(f, PT_SLAB / elem); }

static int pt_splits(int64_t m, int num_kblocks) {
    const int64_t mtiles = (m + PT_BM - 1) / PT_BM;
    const int sms = num_sms();
    int64_t s = (2 * (int64_t)sms + mtiles - 1) / mtiles;
    if (s > PT_MAX_SPLITS) s = PT_MAX_SPLITS;
    if (s > num_kblocks) s = num_kblocks;
    if (s < 1) s = 1;
    return (int)s;
}

template <bool BF16>
// <= 48 registers: 6 warps are allocated as 8 (warp allocation granularity 4), and 8 x 32 x 48 registers is what
// the pair kernel's 16 warps x 104 registers leave free on an SM (co-residency, TSPN_AFFINE_BACKGROUND)
__global__ void __launch_bounds__(PT_THREADS, 7)
predicate_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, int64_t m,
                    int r, int rpad, int num_kblocks, int kb_per_split, int stages, uint32_t tmem_cols,
                    const float* __restrict__ bias, const float* __restrict__ row_bias, int64_t ld_rb, int raw,
                    float* __restrict__ y, float* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int stage_a = PT_BM * PT_SLAB;
    const int stage_bytes = stage_a + rpad * PT_SLAB;
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    uint64_t* const empty = full + stages;
    uint64_t* const tmem_full = empty + stages;
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, split = blockIdx.y;
    const int kb0 = split * kb_per_split;
    const int nkb = min(num_kblocks, kb0 + kb_per_split) - kb0;
    constexpr int K_ELEMS = BF16 ? 64 : 32;          // elements per 128-byte slab
    constexpr int UMMA_PER_SLAB = 4;                 // 32 bytes of K per tcgen05.mma

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ring position kept incrementally: no division by the run-time stage count in a one-thread role
            int s = 0;
            uint32_t ph = 0;
            bool first_lap = true;
            for (int i = 0; i < nkb; ++i) {
                if (!first_lap) mbar_wait(&empty[s], ph ^ 1u);
                uint8_t* a = smem + (size_t)s * stage_bytes;
                mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                tma_load_2d(a, &map_x, (kb0 + i) * K_ELEMS, m_tile * PT_BM, &full[s]);
                tma_load_2d(a + stage_a, &map_w, (kb0 + i) * K_ELEMS, 0, &full[s]);
                if (++s == stages) { s = 0; ph ^= 1u; first_lap = false; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(BF16 ? UMMA_FMT_BF16 : UMMA_FMT_TF32, PT_BM, (uint32_t)rpad, 0, 0);
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t adesc = umma_smem_desc(a_addr, 16, 1024);
                const uint64_t bdesc = umma_smem_desc(a_addr + stage_a, 16, 1024);
#pragma unroll
                for (int k = 0; k < UMMA_PER_SLAB; ++k) {
                    // advance 32 bytes along K inside the swizzled 128-byte row: +2 in 16-byte units
                    if (BF16) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
                    else      umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
                }
                umma_commit(&empty[s]);          // frees the stage once these MMAs have read it
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(tmem_full);              // accumulator complete
        }
    } else {
        // ---- epilogue: warp w owns TMEM lanes [32*(w%4), 32*(w%4)+32) ----
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int quad = warp & 3;
        const int64_t row = (int64_t)m_tile * PT_BM + quad * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
        const bool direct = gridDim.y == 1;
        for (int c0 = 0; c0 < rpad; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + (uint32_t)c0, v);
            if (direct) {
                if (row < m) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = c0 + j;
                        if (c < r) {
                            const float z = v[j] + (bias ? __ldg(bias + c) : 0.0f) +
                                            (row_bias ? __ldg(row_bias + row * ld_rb + c) : 0.0f);
                            y[row * r + c] = raw ? z : 1.0f / (1.0f + __expf(-z));
                        }
                    }
                }
            } else {
                const int64_t mpad = (int64_t)gridDim.x * PT_BM;
                float4* dst = reinterpret_cast<float4*>(partial + ((int64_t)split * mpad + row) * rpad + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// sum the split-K partials in split order, add bias (+ per-row bias), sigmoid (unless raw).  A 32 x 8 thread block
// covers 8 rows, lane-fastest over the predicates: coalesced, and no 64-bit division per element (the flat-index
// form of this kernel spent its time on `idx / r`).
constexpr int PR_ROWS = 8;
__global__ void __launch_bounds__(32 * PR_ROWS)
predicate_reduce_kernel(const float* __restrict__ partial, int splits, int64_t mpad, int rpad, int64_t m, int r,
                        const float* __restrict__ bias, const float* __restrict__ row_bias, int64_t ld_rb, int raw,
                        float* __restrict__ y) {
    const int64_t row = (int64_t)blockIdx.x * PR_ROWS + threadIdx.y;
    if (row >= m) return;
    const float* prow = partial + row * rpad;
    const int64_t split_stride = mpad * rpad;
    for (int c = threadIdx.x; c < r; c += 32) {
        float z = bias ? __ldg(bias + c) : 0.0f;
        for (int s = 0; s < splits; ++s) z += __ldg(prow + s * split_stride + c);
        if (row_bias) z += __ldg(row_bias + row * ld_rb + c);
        y[row * r + c] = raw ? z : 1.0f / (1.0f + __expf(-z));
    }
}

// packed weights: [bf16 Rpad x Kpad16] then [fp32 Rpad x Kpad32], zero padded
__global__ void __launch_bounds__(256)
pack_predicate_kernel(const float* __restrict__ w, int r, int f, int rpad, int kpad16, int kpad32,
                      __nv_bfloat16* __restrict__ wb, float* __restrict__ wf) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n16 = (int64_t)rpad * kpad16, n32 = (int64_t)rpad * kpad32;
    if (idx < n16) {
        const int row = (int)(idx / kpad16), col = (int)(idx % kpad16);
        wb[idx] = __float2bfloat16((row < r && col < f) ? __ldg(w + (int64_t)row * f + col) : 0.0f);
    }
    if (idx < n32) {
        const int row = (int)(idx / kpad32), col = (int)(idx % kpad32);
        wf[idx] = (row < r && col < f) ? __ldg(w + (int64_t)row * f + col) : 0.0f;
    }
}

static int64_t packed_bf16_bytes(int r, int f) { return (int64_t)pt_rpad(r) * pt_kpad(f, 2) * 2; }
static int64_t packed_f32_bytes(int r, int f) { return (int64_t)pt_rpad(r) * pt_kpad(f, 4) * 4; }

int predicate_head_tensor(const void* d_x, int x_is_bf16, int64_t ld_x, int64_t m, int feature_dim,
                          const void* d_w_packed, const float* d_bias, const float* d_row_bias, int64_t ld_rb, int flags,
                          int n_predicates, float* d_y, void* d_workspace, cudaStream_t st) {
    const int r = n_predicates, f = feature_dim;
    const int raw = (flags & TSPN_AFFINE_RAW) ? 1 : 0;
    const bool background = (flags & TSPN_AFFINE_BACKGROUND) != 0;
    const int rpad = pt_rpad(r);
    TSPN_REQUIRE(rpad <= 256, TSPN_ESHAPE, "tensor predicate head supports at most 256 predicates (got %d)", r);
    const int elem = x_is_bf16 ? 2 : 4;
    TSPN_REQUIRE(aligned16(d_x) && (ld_x * elem) % 16 == 0, TSPN_EALIGN,
                 "tensor predicate head: x must be 16-byte aligned with a row stride that is a multiple of 16 bytes "
                 "(ld=%lld elements)", (long long)ld_x);
    TSPN_REQUIRE(aligned16(d_w_packed), TSPN_EALIGN, "tensor predicate head: packed weights must be 16-byte aligned");
    TSPN_REQUIRE(m < (1ll << 31) - PT_BM, TSPN_ESHAPE, "tensor predicate head: too many rows");
    const int k_elems = PT_SLAB / elem;
    const int num_kblocks = (f + k_elems - 1) / k_elems;
    int splits = pt_splits(m, num_kblocks);
    if (background && splits > 4) splits = 4;           // few long-lived CTAs next to the foreground kernel's
    const int kb_per_split = (num_kblocks + splits - 1) / splits;
    const int eff_splits = (num_kblocks + kb_per_split - 1) / kb_per_split;
    const int64_t mtiles = (m + PT_BM - 1) / PT_BM;
    TSPN_REQUIRE(eff_splits == 1 || d_workspace, TSPN_EBADARG, "tensor predicate head: workspace required");
    TSPN_REQUIRE(eff_splits == 1 || aligned16(d_workspace), TSPN_EALIGN, "tensor predicate head: workspace alignment");

    const int stage_bytes = PT_BM * PT_SLAB + rpad * PT_SLAB;
    int stages = PT_SMEM_BUDGET / stage_bytes;
    if (stages > 8) stages = 8;
    if (background && stages > 3) stages = 3;           // <= 80 KB: fits beside a 134 KB foreground CTA
    if (stages > kb_per_split) stages = kb_per_split < 2 ? 2 : kb_per_split;
    const size_t smem_bytes = (size_t)stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < rpad) tmem_cols <<= 1;

    CUtensorMap map_x, map_w;
    {
        const uint64_t dims[2] = {(uint64_t)f, (uint64_t)m};
        const uint64_t strides[1] = {(uint64_t)ld_x * elem};
        const uint32_t box[2] = {(uint32_t)k_elems, (uint32_t)PT_BM};
        int rc = encode_tensor_map(&map_x, x_is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                   2, d_x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != TSPN_OK) return rc;
    }
    {
        const int kpad = pt_kpad(f, elem);
        const uint8_t* base = reinterpret_cast<const uint8_t*>(d_w_packed) + (x_is_bf16 ? 0 : packed_bf16_bytes(r, f));
        const uint64_t dims[2] = {(uint64_t)kpad, (uint64_t)rpad};
        const uint64_t strides[1] = {(uint64_t)kpad * elem};
        const uint32_t box[2] = {(uint32_t)k_elems, (uint32_t)rpad};
        int rc = encode_tensor_map(&map_w, x_is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                   2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != TSPN_OK) return rc;
    }
    float* partial = reinterpret_cast<float*>(d_workspace);
    dim3 grid((unsigned)mtiles, (unsigned)eff_splits);
    if (x_is_bf16) {
        TSPN_CUDA_OK(cudaFuncSetAttribute(predicate_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem_bytes));
        prefer_max_smem(predicate_tc_kernel<true>);
        predicate_tc_kernel<true><<<grid, PT_THREADS, smem_bytes, st>>>(map_x, map_w, m, r, rpad, num_kblocks,
                                                                        kb_per_split, stages, tmem_cols, d_bias,
                                                                        d_row_bias, ld_rb, raw, d_y, partial);
    } else {
        TSPN_CUDA_OK(cudaFuncSetAttribute(predicate_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem_bytes));
        prefer_max_smem(predicate_tc_kernel<false>);
        predicate_tc_kernel<false><<<grid, PT_THREADS, smem_bytes, st>>>(map_x, map_w, m, r, rpad, num_kblocks,
                                                                         kb_per_split, stages, tmem_cols, d_bias,
                                                                         d_row_bias, ld_rb, raw, d_y, partial);
    }
    TSPN_CUDA_OK(cudaGetLastError());
    if (eff_splits > 1) {
        prefer_max_smem(predicate_reduce_kernel);
        predicate_reduce_kernel<<<(unsigned)((m + PR_ROWS - 1) / PR_ROWS), dim3(32, PR_ROWS), 0, st>>>(
            partial, eff_splits, mtiles * PT_BM, rpad, m, r, d_bias, d_row_bias, ld_rb, raw, d_y);
        TSPN_CUDA_OK(cudaGetLastError());
    }
    return TSPN_OK;
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int64_t tspn_predicate_packed_bytes(int n_predicates, int feature_dim) {
    if (n_predicates <= 0 || feature_dim <= 0) return 0;
    return packed_bf16_bytes(n_predicates, feature_dim) + packed_f32_bytes(n_predicates, feature_dim);
}

int tspn_pack_predicate_weights(const float* d_w, int n_predicates, int feature_dim, void* d_packed, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(n_predicates > 0 && feature_dim > 0, TSPN_EBADARG, "tspn_pack_predicate_weights: bad size");
    TSPN_REQUIRE(d_w && d_packed, TSPN_EBADARG, "tspn_pack_predicate_weights: null pointer");
    TSPN_REQUIRE(aligned16(d_packed), TSPN_EALIGN, "tspn_pack_predicate_weights: output must be 16-byte aligned");
    const int rpad = pt_rpad(n_predicates), k16 = pt_kpad(feature_dim, 2), k32 = pt_kpad(feature_dim, 4);
    __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(d_packed);
    float* wf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d_packed) + packed_bf16_bytes(n_predicates, feature_dim));
    const int64_t n = (int64_t)rpad * (k16 > k32 ? k16 : k32);
    pack_predicate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_w, n_predicates, feature_dim,
                                                                                         rpad, k16, k32, wb, wf);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int64_t tspn_predicate_workspace_bytes(int64_t m, int feature_dim, int n_predicates, int precision) {
    if (precision != TSPN_PREC_TENSOR || m <= 0) return 0;
    (void)feature_dim;
    const int64_t mtiles = (m + PT_BM - 1) / PT_BM;
    const int splits = pt_splits(m, 1 << 30);
    if (splits <= 1) return 0;
    return (int64_t)splits * mtiles * PT_BM * pt_rpad(n_predicates) * (int64_t)sizeof(float);
}

}  // extern "C"
