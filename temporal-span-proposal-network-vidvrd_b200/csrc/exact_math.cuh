// exact_math.cuh — the fixed-order fp32 primitives of the "fp32 exact" mode.
//
// Every operation is a single correctly rounded IEEE fp32 op issued through an intrinsic
// (__fmaf_rn / __fmul_rn / __fadd_rn / __fdiv_rn / rintf), so nvcc can neither contract nor
// reassociate it: the results are bit-reproducible across launches, grid shapes and GPUs,
// which is what makes "top-K selection and span frame bounds bit-exact" a checkable claim.
// DESIGN.md ("exact-order arithmetic") is the specification.
#pragma once

#include <stdint.h>

namespace tspn {

// exp(x): round-to-nearest-even range reduction by ln2 (Cody-Waite split), degree-7 Taylor
// polynomial in Horner form, scaling by 2^n through the exponent field.  |rel err| < 1e-7.
__device__ __forceinline__ float exp_det(float x) {
    const float LOG2E = 1.44269504088896341f;
    const float LN2_HI = 0.693359375f;
    const float LN2_LO = -2.12194440e-4f;
    x = fminf(x, 88.0f);
    x = fmaxf(x, -87.0f);
    const float n = rintf(__fmul_rn(x, LOG2E));
    float r = __fmaf_rn(n, -LN2_HI, x);
    r = __fmaf_rn(n, -LN2_LO, r);
    float p = 1.0f / 5040.0f;
    p = __fmaf_rn(p, r, 1.0f / 720.0f);
    p = __fmaf_rn(p, r, 1.0f / 120.0f);
    p = __fmaf_rn(p, r, 1.0f / 24.0f);
    p = __fmaf_rn(p, r, 1.0f / 6.0f);
    p = __fmaf_rn(p, r, 0.5f);
    p = __fmaf_rn(p, r, 1.0f);
    p = __fmaf_rn(p, r, 1.0f);
    const int e = (int)n;                                   // in [-126, 127]
    const float scale = __uint_as_float((uint32_t)(e + 127) << 23);
    return __fmul_rn(p, scale);
}

__device__ __forceinline__ float sigmoid_det(float z) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, exp_det(-z)));
}

}  // namespace tspn
