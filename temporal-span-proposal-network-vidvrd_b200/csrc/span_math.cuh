// span_math.cuh — the temporal-span head at one anchor location and the span decode, as device code shared
// by the fused span-proposal kernel (span_head.cu) and the surviving-pairs kernel (survivors.cu): one fma
// chain, one decode, bit-identical wherever the inputs are bit-identical.
#pragma once

#include "common.cuh"
#include "exact_math.cuh"

namespace tspn {

// [SPEC] s5, every step one correctly rounded fp32 operation (see oracle/exact).
__device__ __forceinline__ void decode_anchor(float dc, float dw, float aw, float ac, int t_len, int32_t* lo_out,
                                              int32_t* hi_out) {
    const float CLAMP = 4.1351666f;                     // fp32 nearest of log(1000/16)
    dw = fminf(dw, CLAMP);
    const float ctr = __fmaf_rn(dc, aw, ac);
    const float w = __fmul_rn(aw, exp_det(dw));
    const float hw = __fmul_rn(0.5f, w);
    float lo = floorf(__fadd_rn(__fadd_rn(ctr, -hw), 0.5f));
    float hi = floorf(__fadd_rn(__fadd_rn(ctr, hw), 0.5f));
    lo = fminf(fmaxf(lo, 0.0f), (float)(t_len - 1));
    hi = fminf(fmaxf(hi, __fadd_rn(lo, 1.0f)), (float)t_len);
    *lo_out = (int32_t)lo;
    *hi_out = (int32_t)hi;
}


// decode the A anchors of one location from their 2A regressions
template <int A>
__device__ __forceinline__ void span_decode_location(const float (&acc)[2 * A], const float* __restrict__ sizes,
                                                     float ac, int t_len, int32_t (&res)[2 * A]) {
#pragma unroll
    for (int a = 0; a < A; ++a)
        decode_anchor(acc[2 * a], acc[2 * a + 1], __ldg(sizes + a), ac, t_len, &res[2 * a], &res[2 * a + 1]);
}

// DPNHead at one column (lib/modeling/relpn/dpn.py:69-73) + decode: xv[ci] = the column's (t-1, t, t+1) inputs,
// w_conv[co * CIN + ci] = (w0, w1, w2, -), w_pred[co * 2A + j]; hidden unit co is the (ci ascending, tap
// ascending) fma chain from its bias, folded into the 2A outputs as soon as it is complete (co ascending) -
// the order of oracle/exact.  Taps outside [0, T) are skipped, not multiplied by zero.
template <int CIN, int A>
__device__ __forceinline__ void span_location(const float (&xv)[CIN][3], bool has_m, bool has_p,
                                              const float4* __restrict__ w_conv, const float* __restrict__ w_pred,
                                              const float* __restrict__ b_conv, const float* __restrict__ b_pred,
                                              const float* __restrict__ sizes, float ac, int t_len,
                                              int32_t (&res)[2 * A]) {
    constexpr int A2 = 2 * A;
    float acc[A2];
#pragma unroll
    for (int j = 0; j < A2; ++j) acc[j] = b_pred[j];
#pragma unroll 2
    for (int co = 0; co < CIN; ++co) {
        float h = b_conv[co];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            const float4 w = w_conv[co * CIN + ci];
            if (has_m) h = __fmaf_rn(w.x, xv[ci][0], h);
            h = __fmaf_rn(w.y, xv[ci][1], h);
            if (has_p) h = __fmaf_rn(w.z, xv[ci][2], h);
        }
        h = fmaxf(h, 0.0f);
#pragma unroll
        for (int j = 0; j < A2; ++j) acc[j] = __fmaf_rn(w_pred[co * A2 + j], h, acc[j]);
    }
    span_decode_location<A>(acc, sizes, ac, t_len, res);
}

}  // namespace tspn
