// predicate.cu — predicate classifier  y = sigmoid(x W^T + b)
//
//   RelationPredictor.forward   lib/modeling/model.py:76-88
//
// Two precisions (include/tspn_b200.h):
//   TSPN_PREC_FP32_EXACT  CUDA-core tiled GEMM whose every output is the k-ascending fma chain
//                         of DESIGN.md ("exact-order arithmetic"): bit-identical to oracle/exact.
//   TSPN_PREC_TENSOR      tcgen05 kernel (predicate_tc.cu): TMA-fed, accumulators in TMEM.
#include "common.cuh"
#include "exact_math.cuh"

namespace tspn {

int predicate_head_tensor(const void* d_x, int x_is_bf16, int64_t ld_x, int64_t m, int feature_dim,
                          const void* d_w_packed, const float* d_bias, const float* d_row_bias, int64_t ld_rb, int flags,
                          int n_predicates, float* d_y, void* d_workspace, cudaStream_t st);

constexpr int PX_BM = 64, PX_BN = 64, PX_BK = 16, PX_THREADS = 256;

__global__ void __launch_bounds__(PX_THREADS)
predicate_exact_kernel(const float* __restrict__ x, int64_t ld_x, int64_t m, int f, const float* __restrict__ w,
                       const float* __restrict__ bias, int r, float* __restrict__ y) {
    __shared__ __align__(16) float As[PX_BK][PX_BM];
    __shared__ __align__(16) float Bs[PX_BK][PX_BN];
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * PX_BM;
    const int col0 = blockIdx.y * PX_BN;
    const int ty = tid >> 4, tx = tid & 15;
    const int lr = tid >> 2, lk = (tid & 3) * 4;       // loader: row/col lr, k offset lk

    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = col0 + tx * 4 + j;
        const float b = (c < r && bias) ? __ldg(bias + c) : 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = b;
    }
    const bool vec_ok = ((ld_x & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int k0 = 0; k0 < f; k0 += PX_BK) {
        // ---- stage x tile (transposed) ----
        {
            const int64_t gr = row0 + lr;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (gr < m) {
                const float* src = x + gr * ld_x + k0 + lk;
                if (vec_ok && k0 + lk + 3 < f) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(src));
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (k0 + lk + i < f) v[i] = __ldg(src + i);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) As[lk + i][lr] = v[i];
        }
        // ---- stage W tile (transposed); rows of W are not 16-byte aligned in general ----
        {
            const int gc = col0 + lr;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (gc < r) {
                const float* src = w + (int64_t)gc * f + k0 + lk;
#pragma unroll
                for (int i = 0; i < 4; ++i) if (k0 + lk + i < f) v[i] = __ldg(src + i);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[lk + i][lr] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < PX_BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gr = row0 + ty * 4 + i;
        if (gr >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = col0 + tx * 4 + j;
            if (c < r) y[gr * r + c] = sigmoid_det(acc[i][j]);
        }
    }
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int tspn_predicate_head(const void* d_x, int x_is_bf16, int64_t ld_x, int64_t m, int feature_dim, const float* d_w,
                        const void* d_w_packed, const float* d_bias, int n_predicates, float* d_y, int precision,
                        void* d_workspace, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(m >= 0 && feature_dim > 0 && n_predicates > 0 && ld_x >= feature_dim, TSPN_EBADARG,
                 "tspn_predicate_head: bad size (m=%lld f=%d r=%d ld=%lld)", (long long)m, feature_dim,
                 n_predicates, (long long)ld_x);
    if (m == 0) return TSPN_OK;
    TSPN_REQUIRE(d_x && d_y, TSPN_EBADARG, "tspn_predicate_head: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == TSPN_PREC_FP32_EXACT) {
        TSPN_REQUIRE(!x_is_bf16 && d_w, TSPN_EBADARG, "tspn_predicate_head: exact mode needs fp32 x and W");
        dim3 grid((unsigned)((m + PX_BM - 1) / PX_BM), (unsigned)((n_predicates + PX_BN - 1) / PX_BN));
        predicate_exact_kernel<<<grid, PX_THREADS, 0, st>>>(reinterpret_cast<const float*>(d_x), ld_x, m,
                                                             feature_dim, d_w, d_bias, n_predicates, d_y);
        TSPN_CUDA_OK(cudaGetLastError());
        return TSPN_OK;
    }
    TSPN_REQUIRE(precision == TSPN_PREC_TENSOR, TSPN_EBADARG, "tspn_predicate_head: unknown precision %d", precision);
    TSPN_REQUIRE(d_w_packed, TSPN_EBADARG, "tspn_predicate_head: tensor mode needs tspn_pack_predicate_weights output");
    return predicate_head_tensor(d_x, x_is_bf16, ld_x, m, feature_dim, d_w_packed, d_bias, nullptr, 0, 0, n_predicates,
                                 d_y, d_workspace, st);
}

int tspn_predicate_head_affine(const void* d_x, int x_is_bf16, int64_t ld_x, int64_t m, int feature_dim,
                               const void* d_w_packed, const float* d_bias, const float* d_row_bias,
                               int64_t ld_row_bias, int n_outputs, float* d_y, int flags, void* d_workspace,
                               void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(m >= 0 && feature_dim > 0 && n_outputs > 0 && ld_x >= feature_dim, TSPN_EBADARG,
                 "tspn_predicate_head_affine: bad size (m=%lld f=%d r=%d ld=%lld)", (long long)m, feature_dim,
                 n_outputs, (long long)ld_x);
    if (m == 0) return TSPN_OK;
    TSPN_REQUIRE(d_x && d_y && d_w_packed, TSPN_EBADARG, "tspn_predicate_head_affine: null pointer");
    TSPN_REQUIRE(!d_row_bias || ld_row_bias >= n_outputs, TSPN_ESHAPE,
                 "tspn_predicate_head_affine: ld_row_bias=%lld < %d", (long long)ld_row_bias, n_outputs);
    return predicate_head_tensor(d_x, x_is_bf16, ld_x, m, feature_dim, d_w_packed, d_bias, d_row_bias, ld_row_bias,
                                 flags, n_outputs, d_y, d_workspace, (cudaStream_t)stream);
}

}  // extern "C"
