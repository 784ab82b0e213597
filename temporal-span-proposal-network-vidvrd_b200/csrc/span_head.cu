// span_head.cu — temporal-span head and span decode.
//
//   tspn_span_head     DPNHead.forward   lib/modeling/relpn/dpn.py:55-73
//                      Conv1d(Cin->Cin, k3, p1) -> ReLU -> Conv1d(Cin->2A, k1)
//   tspn_span_decode   anchors of lib/modeling/relpn/anchor_generator.py:48-104 applied to the
//                      regressions, to integer frame bounds ([SPEC] s5; DPN.forward / RelNMS are
//                      broken / a stub in the reference, relpn/dpn.py:24-28, relpn/rel_nms.py:14-15)
//
// TSPN_PREC_FP32_EXACT: one thread per (pair, frame); the hidden unit h[co] is the
// (ci ascending, tap ascending) fma chain from its bias and is folded into the 2A outputs as
// soon as it is complete (co ascending) — exactly the order of oracle/exact, no hidden tensor
// ever exists in memory.  TSPN_PREC_TENSOR: implicit GEMM on tcgen05 (span_head_tc.cu).
#include "common.cuh"
#include "exact_math.cuh"
#include "span_math.cuh"

namespace tspn {

int span_head_tensor(const float* d_x, const int64_t* d_rows, int64_t row_base, int64_t row_stride, int64_t ld_t,
                     int64_t k, int cin,
                     int t, const float* d_conv_w, const float* d_conv_b, const float* d_pred_w,
                     const float* d_pred_b, int a2, float* d_out, void* d_workspace, cudaStream_t st);

constexpr int SH_THREADS = 128;
constexpr int SH_MAX_A2 = 16;
constexpr int SH_REG_CIN = 16;

__global__ void __launch_bounds__(SH_THREADS)
span_head_exact_kernel(const float* __restrict__ x, const int64_t* __restrict__ rows, int64_t row_base,
                       int64_t row_stride, int64_t ld_t, int cin, int t_len, const float* __restrict__ conv_w,
                       const float* __restrict__ conv_b, const float* __restrict__ pred_w,
                       const float* __restrict__ pred_b, int a2, float* __restrict__ out) {
    const int64_t p = blockIdx.y;
    const int t = blockIdx.x * SH_THREADS + threadIdx.x;
    if (t >= t_len) return;
    float* o = out + p * a2 * (int64_t)t_len + t;
    const int64_t src = rows ? rows[p] - (rows[p] >= 0 ? row_base : 0) : p;
    if (src < 0) {
        for (int j = 0; j < a2; ++j) o[(int64_t)j * t_len] = 0.0f;
        return;
    }
    const float* xr = x + src * row_stride;
    const bool has_m = t > 0, has_p = t + 1 < t_len;
    float acc[SH_MAX_A2];
#pragma unroll
    for (int j = 0; j < SH_MAX_A2; ++j) acc[j] = (j < a2) ? (pred_b ? __ldg(pred_b + j) : 0.0f) : 0.0f;

    float xm[SH_REG_CIN], x0[SH_REG_CIN], xp[SH_REG_CIN];
    const bool cached = cin <= SH_REG_CIN;
    if (cached) {
#pragma unroll
        for (int ci = 0; ci < SH_REG_CIN; ++ci) {
            if (ci < cin) {
                const float* xc = xr + (int64_t)ci * ld_t + t;
                xm[ci] = has_m ? __ldg(xc - 1) : 0.0f;
                x0[ci] = __ldg(xc);
                xp[ci] = has_p ? __ldg(xc + 1) : 0.0f;
            }
        }
    }
    for (int co = 0; co < cin; ++co) {
        float h = conv_b ? __ldg(conv_b + co) : 0.0f;
        const float* w = conv_w + (int64_t)co * cin * 3;
        if (cached) {
#pragma unroll
            for (int ci = 0; ci < SH_REG_CIN; ++ci) {
                if (ci < cin) {
                    if (has_m) h = __fmaf_rn(__ldg(w + ci * 3 + 0), xm[ci], h);
                    h = __fmaf_rn(__ldg(w + ci * 3 + 1), x0[ci], h);
                    if (has_p) h = __fmaf_rn(__ldg(w + ci * 3 + 2), xp[ci], h);
                }
            }
        } else {
            for (int ci = 0; ci < cin; ++ci) {
                const float* xc = xr + (int64_t)ci * ld_t + t;
                if (has_m) h = __fmaf_rn(__ldg(w + ci * 3 + 0), __ldg(xc - 1), h);
                h = __fmaf_rn(__ldg(w + ci * 3 + 1), __ldg(xc), h);
                if (has_p) h = __fmaf_rn(__ldg(w + ci * 3 + 2), __ldg(xc + 1), h);
            }
        }
        h = fmaxf(h, 0.0f);
#pragma unroll
        for (int j = 0; j < SH_MAX_A2; ++j)
            if (j < a2) acc[j] = __fmaf_rn(__ldg(pred_w + (int64_t)j * cin + co), h, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < SH_MAX_A2; ++j)
        if (j < a2) o[(int64_t)j * t_len] = acc[j];
}

// Small-channel specialisation (the pair stage feeds the 8 geometry channels): weights staged in
// shared memory as 128-bit broadcast reads, a thread owns FPT consecutive frames (vector loads and
// stores), same fma order per output as the generic kernel -> same bits.  FPT = 2 keeps the thread
// at ~64 registers (7 CTAs per SM): the kernel streams 64*T bytes per pair and needs the occupancy.
template <int CIN, int A2, int FPT>
__global__ void __launch_bounds__(SH_THREADS)
span_head_small_kernel(const float* __restrict__ x, const int64_t* __restrict__ rows, int64_t row_base,
                       int64_t row_stride, int64_t ld_t, int t_len, const float* __restrict__ conv_w,
                       const float* __restrict__ conv_b, const float* __restrict__ pred_w,
                       const float* __restrict__ pred_b, float* __restrict__ out) {
    static_assert(FPT == 2 || FPT == 4, "FPT must be 2 or 4");
    static_assert(A2 % 4 == 0, "A2 must be a multiple of 4");
    __shared__ __align__(16) float4 w_conv[CIN * CIN];      // [co][ci] -> (w0, w1, w2, -)
    __shared__ __align__(16) float w_pred[CIN * A2];        // [co][j]
    __shared__ float b_conv[CIN], b_pred[A2];
    for (int i = threadIdx.x; i < CIN * CIN; i += SH_THREADS)
        w_conv[i] = make_float4(__ldg(conv_w + i * 3), __ldg(conv_w + i * 3 + 1), __ldg(conv_w + i * 3 + 2), 0.0f);
    for (int i = threadIdx.x; i < CIN * A2; i += SH_THREADS) {
        const int co = i / A2, j = i - co * A2;
        w_pred[i] = __ldg(pred_w + j * CIN + co);
    }
    if (threadIdx.x < CIN) b_conv[threadIdx.x] = conv_b ? __ldg(conv_b + threadIdx.x) : 0.0f;
    if (threadIdx.x < A2) b_pred[threadIdx.x] = pred_b ? __ldg(pred_b + threadIdx.x) : 0.0f;
    __syncthreads();

    const int64_t p = blockIdx.y;
    const int t0 = (blockIdx.x * SH_THREADS + threadIdx.x) * FPT;
    if (t0 >= t_len) return;
    float* o = out + p * A2 * (int64_t)t_len + t0;
    const bool full = t0 + FPT - 1 < t_len;
    const bool vec_out = ((t_len & (FPT - 1)) == 0) && full;
    const int64_t src = rows ? rows[p] - (rows[p] >= 0 ? row_base : 0) : p;
    if (src < 0) {
#pragma unroll
        for (int j = 0; j < A2; ++j)
            for (int i = 0; i < FPT && t0 + i < t_len; ++i) o[(int64_t)j * t_len + i] = 0.0f;
        return;
    }
    const float* xr = x + src * row_stride + t0;
    // xv[ci][0 .. FPT+1] = x[ci][t0-1 .. t0+FPT], zero outside [0, T)
    float xv[CIN][FPT + 2];
    const bool vec_in = ((ld_t & (FPT - 1)) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                        ((row_stride & 3) == 0) && full;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
        const float* xc = xr + (int64_t)ci * ld_t;
        if (vec_in) {
            if (FPT == 4) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(xc));
                xv[ci][1] = q.x; xv[ci][2] = q.y; xv[ci][FPT - 1] = q.z; xv[ci][FPT] = q.w;
            } else {
                const float2 q = __ldg(reinterpret_cast<const float2*>(xc));
                xv[ci][1] = q.x; xv[ci][2] = q.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < FPT; ++i) xv[ci][1 + i] = (t0 + i < t_len) ? __ldg(xc + i) : 0.0f;
        }
        xv[ci][0] = t0 > 0 ? __ldg(xc - 1) : 0.0f;
        xv[ci][FPT + 1] = t0 + FPT < t_len ? __ldg(xc + FPT) : 0.0f;
    }
    float acc[A2][FPT];
#pragma unroll
    for (int j = 0; j < A2; ++j)
#pragma unroll
        for (int i = 0; i < FPT; ++i) acc[j][i] = b_pred[j];
#pragma unroll 1
    for (int co = 0; co < CIN; ++co) {
        float h[FPT];
#pragma unroll
        for (int i = 0; i < FPT; ++i) h[i] = b_conv[co];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            const float4 w = w_conv[co * CIN + ci];
#pragma unroll
            for (int i = 0; i < FPT; ++i) {
                const int t = t0 + i;
                // taps outside [0, T) are skipped, not multiplied by zero (same chain as the oracle)
                if (t > 0) h[i] = __fmaf_rn(w.x, xv[ci][i], h[i]);
                h[i] = __fmaf_rn(w.y, xv[ci][i + 1], h[i]);
                if (t + 1 < t_len) h[i] = __fmaf_rn(w.z, xv[ci][i + 2], h[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < FPT; ++i) h[i] = fmaxf(h[i], 0.0f);
#pragma unroll
        for (int j4 = 0; j4 < A2 / 4; ++j4) {
            const float4 wp = *reinterpret_cast<const float4*>(&w_pred[co * A2 + 4 * j4]);
            const float wv[4] = {wp.x, wp.y, wp.z, wp.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int i = 0; i < FPT; ++i) acc[4 * j4 + jj][i] = __fmaf_rn(wv[jj], h[i], acc[4 * j4 + jj][i]);
        }
    }
#pragma unroll
    for (int j = 0; j < A2; ++j) {
        float* oj = o + (int64_t)j * t_len;
        if (vec_out) {
            if (FPT == 4) *reinterpret_cast<float4*>(oj) = make_float4(acc[j][0], acc[j][1], acc[j][FPT - 2], acc[j][FPT - 1]);
            else          *reinterpret_cast<float2*>(oj) = make_float2(acc[j][0], acc[j][1]);
        } else {
#pragma unroll
            for (int i = 0; i < FPT; ++i)
                if (t0 + i < t_len) oj[i] = acc[j][i];
        }
    }
}

__global__ void __launch_bounds__(256)
span_decode_kernel(const float* __restrict__ reg, int64_t k, int a_n, int t_len, int n_loc,
                   const float* __restrict__ sizes, float stride, int32_t* __restrict__ spans) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = k * n_loc * a_n;
    if (idx >= total) return;
    const int a = (int)(idx % a_n);
    const int l = (int)((idx / a_n) % n_loc);
    const int64_t p = idx / ((int64_t)a_n * n_loc);
    const float ac = __fmul_rn((float)l, stride);
    int col = (int)floorf(ac);
    col = min(col, t_len - 1);
    const float aw = __ldg(sizes + a);
    const float dc = __ldg(reg + (p * 2 * a_n + 2 * a) * t_len + col);
    const float dw = __ldg(reg + (p * 2 * a_n + 2 * a + 1) * t_len + col);
    int32_t lo, hi;
    decode_anchor(dc, dw, aw, ac, t_len, &lo, &hi);
    spans[2 * idx] = lo;
    spans[2 * idx + 1] = hi;
}

// Fused span proposals: the head is evaluated only at the anchor columns floor(l * stride) the
// decode reads (n_loc of the T frames), with the fma chain of span_head_*_kernel per column, and the
// regressions are decoded in registers - the [K, 2A, T] regression tensor never exists.  One thread
// per (pair, anchor location); bit-identical to tspn_span_head(fp32) followed by tspn_span_decode.
template <int CIN, int A>
__global__ void __launch_bounds__(SH_THREADS)
span_proposals_small_kernel(const float* __restrict__ x, const int64_t* __restrict__ rows, int64_t row_base,
                            int64_t row_stride, int64_t ld_t, int t_len, const float* __restrict__ conv_w,
                            const float* __restrict__ conv_b, const float* __restrict__ pred_w,
                            const float* __restrict__ pred_b, const float* __restrict__ sizes, float stride,
                            int n_loc, int32_t* __restrict__ spans) {
    constexpr int A2 = 2 * A;
    __shared__ __align__(16) float4 w_conv[CIN * CIN];      // [co][ci] -> (w0, w1, w2, -)
    __shared__ __align__(16) float w_pred[CIN * A2];        // [co][j]
    __shared__ float b_conv[CIN], b_pred[A2];
    for (int i = threadIdx.x; i < CIN * CIN; i += SH_THREADS)
        w_conv[i] = make_float4(__ldg(conv_w + i * 3), __ldg(conv_w + i * 3 + 1), __ldg(conv_w + i * 3 + 2), 0.0f);
    for (int i = threadIdx.x; i < CIN * A2; i += SH_THREADS) {
        const int co = i / A2, j = i - co * A2;
        w_pred[i] = __ldg(pred_w + j * CIN + co);
    }
    if (threadIdx.x < CIN) b_conv[threadIdx.x] = conv_b ? __ldg(conv_b + threadIdx.x) : 0.0f;
    if (threadIdx.x < A2) b_pred[threadIdx.x] = pred_b ? __ldg(pred_b + threadIdx.x) : 0.0f;
    __syncthreads();

    const int64_t p = blockIdx.y;
    const int l = blockIdx.x * SH_THREADS + threadIdx.x;
    if (l >= n_loc) return;
    const float ac = __fmul_rn((float)l, stride);
    const int t = min((int)floorf(ac), t_len - 1);
    const int64_t src = rows ? rows[p] - (rows[p] >= 0 ? row_base : 0) : p;
    int32_t res[A2];
    if (src < 0) {                                          // padding row: regressions are 0
        float zero[A2];
#pragma unroll
        for (int j = 0; j < A2; ++j) zero[j] = 0.0f;
        span_decode_location<A>(zero, sizes, ac, t_len, res);
    } else {
        const float* xr = x + src * row_stride + t;
        const bool has_m = t > 0, has_p = t + 1 < t_len;
        float xv[CIN][3];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            const float* xc = xr + (int64_t)ci * ld_t;
            xv[ci][0] = has_m ? __ldg(xc - 1) : 0.0f;
            xv[ci][1] = __ldg(xc);
            xv[ci][2] = has_p ? __ldg(xc + 1) : 0.0f;
        }
        span_location<CIN, A>(xv, has_m, has_p, w_conv, w_pred, b_conv, b_pred, sizes, ac, t_len, res);
    }
    int32_t* o = spans + (p * n_loc + l) * A2;               // [k][n_loc * A][2]
    if (A2 % 4 == 0) {
#pragma unroll
        for (int q = 0; q < A2 / 4; ++q)
            reinterpret_cast<int4*>(o)[q] = make_int4(res[4 * q], res[4 * q + 1], res[4 * q + 2], res[4 * q + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < A2; ++j) o[j] = res[j];
    }
}

// any Cin / A (2A <= SH_MAX_A2): same chain, inputs re-read from L1/L2 per hidden unit
__global__ void __launch_bounds__(SH_THREADS)
span_proposals_generic_kernel(const float* __restrict__ x, const int64_t* __restrict__ rows, int64_t row_base,
                              int64_t row_stride, int64_t ld_t, int cin, int t_len,
                              const float* __restrict__ conv_w, const float* __restrict__ conv_b,
                              const float* __restrict__ pred_w, const float* __restrict__ pred_b, int a_n,
                              const float* __restrict__ sizes, float stride, int n_loc, int32_t* __restrict__ spans) {
    const int64_t p = blockIdx.y;
    const int l = blockIdx.x * SH_THREADS + threadIdx.x;
    if (l >= n_loc) return;
    const int a2 = 2 * a_n;
    const float ac = __fmul_rn((float)l, stride);
    const int t = min((int)floorf(ac), t_len - 1);
    const int64_t src = rows ? rows[p] - (rows[p] >= 0 ? row_base : 0) : p;
    float acc[SH_MAX_A2];
#pragma unroll
    for (int j = 0; j < SH_MAX_A2; ++j) acc[j] = 0.0f;
    if (src >= 0) {
#pragma unroll
        for (int j = 0; j < SH_MAX_A2; ++j) acc[j] = (j < a2 && pred_b) ? __ldg(pred_b + j) : 0.0f;
        const float* xr = x + src * row_stride + t;
        const bool has_m = t > 0, has_p = t + 1 < t_len;
        for (int co = 0; co < cin; ++co) {
            float h = conv_b ? __ldg(conv_b + co) : 0.0f;
            const float* w = conv_w + (int64_t)co * cin * 3;
            for (int ci = 0; ci < cin; ++ci) {
                const float* xc = xr + (int64_t)ci * ld_t;
                if (has_m) h = __fmaf_rn(__ldg(w + ci * 3 + 0), __ldg(xc - 1), h);
                h = __fmaf_rn(__ldg(w + ci * 3 + 1), __ldg(xc), h);
                if (has_p) h = __fmaf_rn(__ldg(w + ci * 3 + 2), __ldg(xc + 1), h);
            }
            h = fmaxf(h, 0.0f);
#pragma unroll
            for (int j = 0; j < SH_MAX_A2; ++j)
                if (j < a2) acc[j] = __fmaf_rn(__ldg(pred_w + (int64_t)j * cin + co), h, acc[j]);
        }
    }
    int32_t* o = spans + (p * n_loc + l) * a2;
#pragma unroll
    for (int a = 0; a < SH_MAX_A2 / 2; ++a) {
        if (a < a_n) {
            int32_t lo, hi;
            decode_anchor(acc[2 * a], acc[2 * a + 1], __ldg(sizes + a), ac, t_len, &lo, &hi);
            o[2 * a] = lo;
            o[2 * a + 1] = hi;
        }
    }
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int tspn_span_num_locations(int t, float stride) {
    if (t <= 0 || !(stride > 0.0f)) return 0;
    // len(torch.arange(0, T+1, step=stride)) = ceil((T+1)/stride), anchor_generator.py:50-52
    const double q = ((double)t + 1.0) / (double)stride;
    int n = (int)q;
    if ((double)n < q) ++n;
    return n;
}

int tspn_span_head(const float* d_x, const int64_t* d_rows, int64_t row_base, int64_t row_stride, int64_t ld_t,
                   int64_t k, int cin, int t, const float* d_conv_w, const float* d_conv_b, const float* d_pred_w,
                   const float* d_pred_b, int a2, float* d_out, int precision, void* d_workspace, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(k >= 0 && cin > 0 && t > 0 && a2 > 0 && ld_t >= t && row_stride >= 0, TSPN_EBADARG,
                 "tspn_span_head: bad size");
    TSPN_REQUIRE(a2 <= SH_MAX_A2, TSPN_ESHAPE, "tspn_span_head: 2A=%d exceeds the supported maximum %d", a2,
                 SH_MAX_A2);
    if (k == 0) return TSPN_OK;
    TSPN_REQUIRE(d_x && d_conv_w && d_pred_w && d_out, TSPN_EBADARG, "tspn_span_head: null pointer");
    TSPN_REQUIRE(k < 65536, TSPN_ESHAPE, "tspn_span_head: k=%lld must be < 65536 per call", (long long)k);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == TSPN_PREC_FP32_EXACT) {
        if (cin == 8 && a2 == 8) {
            constexpr int FPT = 2;
            dim3 grid((unsigned)((t + FPT * SH_THREADS - 1) / (FPT * SH_THREADS)), (unsigned)k);
            span_head_small_kernel<8, 8, FPT><<<grid, SH_THREADS, 0, st>>>(d_x, d_rows, row_base, row_stride, ld_t,
                                                                           t, d_conv_w, d_conv_b, d_pred_w, d_pred_b,
                                                                           d_out);
        } else {
            dim3 grid((unsigned)((t + SH_THREADS - 1) / SH_THREADS), (unsigned)k);
            span_head_exact_kernel<<<grid, SH_THREADS, 0, st>>>(d_x, d_rows, row_base, row_stride, ld_t, cin, t,
                                                                d_conv_w, d_conv_b, d_pred_w, d_pred_b, a2, d_out);
        }
        TSPN_CUDA_OK(cudaGetLastError());
        return TSPN_OK;
    }
    TSPN_REQUIRE(precision == TSPN_PREC_TENSOR, TSPN_EBADARG, "tspn_span_head: unknown precision %d", precision);
    return span_head_tensor(d_x, d_rows, row_base, row_stride, ld_t, k, cin, t, d_conv_w, d_conv_b, d_pred_w, d_pred_b, a2,
                            d_out, d_workspace, st);
}

int tspn_span_decode(const float* d_reg, int64_t k, int n_anchors, int t, const float* d_sizes, float stride,
                     int32_t* d_spans, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(k >= 0 && n_anchors > 0 && t > 0 && stride > 0.0f, TSPN_EBADARG, "tspn_span_decode: bad size");
    if (k == 0) return TSPN_OK;
    TSPN_REQUIRE(d_reg && d_sizes && d_spans, TSPN_EBADARG, "tspn_span_decode: null pointer");
    const int n_loc = tspn_span_num_locations(t, stride);
    const int64_t total = k * n_loc * n_anchors;
    span_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_reg, k, n_anchors, t, n_loc, d_sizes, stride, d_spans);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_span_proposals(const float* d_x, const int64_t* d_rows, int64_t row_base, int64_t row_stride, int64_t ld_t,
                        int64_t k, int cin, int t, const float* d_conv_w, const float* d_conv_b,
                        const float* d_pred_w, const float* d_pred_b, int n_anchors, const float* d_sizes,
                        float stride, int32_t* d_spans, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(k >= 0 && cin > 0 && t > 0 && n_anchors > 0 && ld_t >= t && row_stride >= 0 && stride > 0.0f,
                 TSPN_EBADARG, "tspn_span_proposals: bad size");
    TSPN_REQUIRE(2 * n_anchors <= SH_MAX_A2, TSPN_ESHAPE, "tspn_span_proposals: 2A=%d exceeds the supported maximum %d",
                 2 * n_anchors, SH_MAX_A2);
    if (k == 0) return TSPN_OK;
    TSPN_REQUIRE(d_x && d_conv_w && d_pred_w && d_sizes && d_spans, TSPN_EBADARG, "tspn_span_proposals: null pointer");
    TSPN_REQUIRE(k < 65536, TSPN_ESHAPE, "tspn_span_proposals: k=%lld must be < 65536 per call", (long long)k);
    const int n_loc = tspn_span_num_locations(t, stride);
    dim3 grid((unsigned)((n_loc + SH_THREADS - 1) / SH_THREADS), (unsigned)k);
    cudaStream_t st = (cudaStream_t)stream;
    if (cin == 8 && n_anchors == 4 && aligned16(d_spans))
        span_proposals_small_kernel<8, 4><<<grid, SH_THREADS, 0, st>>>(d_x, d_rows, row_base, row_stride, ld_t, t,
                                                                      d_conv_w, d_conv_b, d_pred_w, d_pred_b, d_sizes,
                                                                      stride, n_loc, d_spans);
    else
        span_proposals_generic_kernel<<<grid, SH_THREADS, 0, st>>>(d_x, d_rows, row_base, row_stride, ld_t, cin, t,
                                                                   d_conv_w, d_conv_b, d_pred_w, d_pred_b, n_anchors,
                                                                   d_sizes, stride, n_loc, d_spans);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}


}  // extern "C"
