// geo_math.cuh — the per-frame pair geometry ([SPEC] s2) as device code shared by the
// all-pairs kernel (geo_viou.cu) and the surviving-pairs kernel (survivors.cu): both evaluate exactly the same
// instruction sequence per (pair, frame), so a channel value recomputed for a surviving pair is bit-identical
// to the one the all-pairs kernel stored.
#pragma once

#include "common.cuh"

namespace tspn {

constexpr int GEO_FPT = 4;                            // frames per thread

// box j of a chunk staged with SWIZZLE_128B: the 16-byte slot index (address bits 4..6) is XORed
// with address bits 7..9 of the shared-memory address, so the pattern is a function of the
// absolute address and stages only need 128-byte alignment.
__device__ __forceinline__ float4 ld_box(uint32_t stage_addr, int j) {
    const uint32_t lin = stage_addr + ((uint32_t)j << 4);
    const uint32_t phys = lin ^ (((lin >> 7) & 7u) << 4);
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(phys));
    return v;
}

__device__ __forceinline__ float rcp_fast(float x) {          // MUFU.RCP, <= 1 ulp
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_fast(float x) {          // MUFU.LG2, abs err 2^-22.6
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// log(a / b) for a, b >= 1.  Near a == b the quotient form loses relative accuracy, so the result is
// 2*atanh(z), z = (a-b)/(a+b), as an odd series (|z| < 0.15: truncation < 1e-9 relative); elsewhere
// ln2 * lg2(a/b), whose absolute error is far below 1e-5 of a result that is at least 0.3.
__device__ __forceinline__ float log_ratio(float a, float b, float rb) {
    const float z = (a - b) * rcp_fast(a + b);
    const float z2 = z * z;
    float p = fmaf(z2, 1.0f / 9.0f, 1.0f / 7.0f);
    p = fmaf(z2, p, 1.0f / 5.0f);
    p = fmaf(z2, p, 1.0f / 3.0f);
    const float near = 2.0f * fmaf(z * z2, p, z);
    const float far = 0.69314718056f * lg2_fast(a * rb);
    return fabsf(z) < 0.15f ? near : far;
}

// ---- one (pair, 4 frames) step of a thread: the eight channels + the three fp32 partial sums ----------
// per-thread partial sums stay in fp32: 4 integer-valued products <= 4 * 2^21 < 2^24 are exact
// load_s(j) / load_o(j): box j of the subject / object chunk (j0 <= j <= j0 + GEO_FPT; shared-memory stages in
// the all-pairs kernel, global memory in the surviving-pairs kernel - the arithmetic is the same code)
// INTERIOR: the caller guarantees a <= t0 and t0 + GEO_FPT < b - all four frames and their forward differences lie
// inside the window, so every predicate below is true and is dropped at compile time.  The arithmetic is the same
// instruction sequence (no product feeds an add without an explicit rounding point: nothing for the compiler to
// contract differently), hence the same bits.
template <bool CLIP, bool INTERIOR = false, typename LoadS, typename LoadO>
__device__ __forceinline__ void geo_step(LoadS load_s, LoadO load_o, int j0, int t0, int a, int b,
                                         float (&out)[TSPN_GEO_CHANNELS][GEO_FPT], float& fsum_i, float& fsum_s,
                                         float& fsum_o) {
    fsum_i = 0.0f; fsum_s = 0.0f; fsum_o = 0.0f;
    if (!INTERIOR) {
#pragma unroll
        for (int ch = 0; ch < TSPN_GEO_CHANNELS; ++ch)
#pragma unroll
            for (int i = 0; i < GEO_FPT; ++i) out[ch][i] = 0.0f;
        if (!(t0 < b && t0 + GEO_FPT > a)) return;
    }
    float dcx[GEO_FPT + 1], dcy[GEO_FPT + 1], wo[GEO_FPT + 1], ho[GEO_FPT + 1];
    float rwo[GEO_FPT + 1], rho[GEO_FPT + 1];
#pragma unroll
    for (int i = 0; i <= GEO_FPT; ++i) {
        const float4 sb = load_s(j0 + i);
        const float4 ob = load_o(j0 + i);
        wo[i] = (ob.z - ob.x) + 1.0f;
        ho[i] = (ob.w - ob.y) + 1.0f;
        rwo[i] = rcp_fast(wo[i]);
        rho[i] = rcp_fast(ho[i]);
        // centre deltas from coordinate differences: exact for integer boxes and
        // free of the cancellation that (x1+x2)/2 - (x1'+x2')/2 would carry
        dcx[i] = 0.5f * ((sb.x - ob.x) + (sb.z - ob.z));
        dcy[i] = 0.5f * ((sb.y - ob.y) + (sb.w - ob.w));
        if (i < GEO_FPT) {
            const int t = t0 + i;
            const bool in = INTERIOR || ((t >= a) && (t < b));
            const float ws = (sb.z - sb.x) + 1.0f, hs = (sb.w - sb.y) + 1.0f;
            const float iw = fmaxf((fminf(sb.z, ob.z) - fmaxf(sb.x, ob.x)) + 1.0f, 0.0f);
            const float ih = fmaxf((fminf(sb.w, ob.w) - fmaxf(sb.y, ob.y)) + 1.0f, 0.0f);
            // explicit rounding points: the volume sums must not depend on whether the
            // compiler contracts these products into the accumulation (template variants)
            const float inter = __fmul_rn(iw, ih);
            const float as = __fmul_rn(ws, hs), ao = __fmul_rn(wo[i], ho[i]);
            if (in) {
                out[0][i] = dcx[i] * rwo[i];
                out[1][i] = dcy[i] * rho[i];
                out[2][i] = log_ratio(ws, wo[i], rwo[i]);
                out[3][i] = log_ratio(hs, ho[i], rho[i]);
                out[4][i] = inter * rcp_fast((as + ao) - inter);
                out[7][i] = 1.0f;
                fsum_i = __fadd_rn(fsum_i, inter);
                if (CLIP) {
                    fsum_s = __fadd_rn(fsum_s, as);
                    fsum_o = __fadd_rn(fsum_o, ao);
                }
            }
        }
    }
    // forward differences in closed form:
    //   c0[t+1]-c0[t] = (dcx[t+1]*wo[t] - dcx[t]*wo[t+1]) / (wo[t]*wo[t+1])
    // (two-product compensation keeps the numerator exact to one rounding)
#pragma unroll
    for (int i = 0; i < GEO_FPT; ++i) {
        const int t = t0 + i;
        if (INTERIOR || (t >= a && t + 1 < b)) {
            float p = dcx[i] * wo[i + 1];
            float e = fmaf(dcx[i], wo[i + 1], -p);
            out[5][i] = (fmaf(dcx[i + 1], wo[i], -p) - e) * (rwo[i] * rwo[i + 1]);
            p = dcy[i] * ho[i + 1];
            e = fmaf(dcy[i], ho[i + 1], -p);
            out[6][i] = (fmaf(dcy[i + 1], ho[i], -p) - e) * (rho[i] * rho[i + 1]);
        }
    }
}

}  // namespace tspn
