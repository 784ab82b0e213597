/*
 * oracle/exact/tspn_exact.c — CPU restatement of the FIXED-ORDER fp32 arithmetic of
 * the pair stage.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): nothing in the
 * product links or loads this file.
 *
 * Why it exists: BASELINE.json asks for pair indices, top-K selection and span frame
 * bounds to be bit-exact in "fp32 mode".  The reference computes its scores with
 * torch's sgemm / conv / sigmoid (lib/modeling/relpn/ppn.py:107-112,
 * lib/modeling/model.py:85-88, lib/modeling/relpn/dpn.py:69-73) whose summation order
 * is unspecified, so bit-exactness is only meaningful against a *defined* order.  This
 * file is that definition (DESIGN.md "exact-order arithmetic"):
 *
 *   linear   y[o] = b[o]; for i ascending: y[o] = fma(x[i], W[o][i], y[o])
 *   dot      z = 0;       for c ascending: z = fma(S[s][c], O[o][c], z)
 *   conv k3  h = b[co];   for ci ascending, for dt in 0..2:
 *                            h = fma(W[co][ci][dt], x[ci][t+dt-1] (0 outside), h)
 *   sigmoid  1 / (1 + exp_det(-z))   with exp_det below (fma/mul/add/rint only)
 *   decode   [SPEC] s5, every operation a single correctly rounded fp32 op
 *
 * PARITY UNPINNED by the reference for the summation order itself (the reference has
 * none); the *values* are pinned to the reference within 1e-6 by tests/golden.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#define TSPN_CLONES __attribute__((target_clones("avx2,fma", "default")))
#else
#define TSPN_CLONES
#endif

/* ---- deterministic exp: round-to-nearest range reduction + degree-7 Horner ---- */
static inline float exp_det(float x) {
    const float LOG2E = 1.44269504088896341f;
    const float LN2_HI = 0.693359375f;            /* exactly representable, 9 bits */
    const float LN2_LO = -2.12194440e-4f;
    if (x > 88.0f) x = 88.0f;
    if (x < -87.0f) x = -87.0f;
    float n = rintf(x * LOG2E);
    float r = fmaf(n, -LN2_HI, x);
    r = fmaf(n, -LN2_LO, r);
    float p = 1.0f / 5040.0f;
    p = fmaf(p, r, 1.0f / 720.0f);
    p = fmaf(p, r, 1.0f / 120.0f);
    p = fmaf(p, r, 1.0f / 24.0f);
    p = fmaf(p, r, 1.0f / 6.0f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int32_t e = (int32_t)n;
    /* split the scale so that 2^e never leaves the normal range: e in [-126, 127] */
    uint32_t bits = (uint32_t)(e + 127) << 23;
    float scale;
    memcpy(&scale, &bits, 4);
    return p * scale;
}

static inline float sigmoid_det(float z) {
    return 1.0f / (1.0f + exp_det(-z));
}

float tspn_exact_exp(float x) { return exp_det(x); }
float tspn_exact_sigmoid(float x) { return sigmoid_det(x); }

/* y[m][o] = act(b[o] + sum_i x[m][i] W[o][i]),  Wt is W transposed: [In][Out] */
TSPN_CLONES
void tspn_exact_linear_t(const float* x, int64_t ldx, const float* Wt, const float* b,
                         float* y, int64_t ldy, int m, int in, int out, int act /*0 none 1 relu 2 sigmoid*/) {
    for (int r = 0; r < m; ++r) {
        float* acc = y + (int64_t)r * ldy;
        for (int o = 0; o < out; ++o) acc[o] = b ? b[o] : 0.0f;
        const float* xr = x + (int64_t)r * ldx;
        for (int i = 0; i < in; ++i) {
            const float xv = xr[i];
            const float* w = Wt + (int64_t)i * out;
            for (int o = 0; o < out; ++o) acc[o] = fmaf(xv, w[o], acc[o]);
        }
        if (act == 1) for (int o = 0; o < out; ++o) acc[o] = acc[o] > 0.0f ? acc[o] : 0.0f;
        if (act == 2) for (int o = 0; o < out; ++o) acc[o] = sigmoid_det(acc[o]);
    }
}

/* M[s][o] = sigmoid(sum_c S[s][c] O[o][c]) */
TSPN_CLONES
void tspn_exact_pair_scores(const float* S, const float* O, float* M, int n, int c) {
    for (int s = 0; s < n; ++s)
        for (int o = 0; o < n; ++o) {
            float z = 0.0f;
            for (int k = 0; k < c; ++k) z = fmaf(S[(int64_t)s * c + k], O[(int64_t)o * c + k], z);
            M[(int64_t)s * n + o] = sigmoid_det(z);
        }
}

/* DPNHead: x [K][Cin][T], conv_w [Cin][Cin][3], pred_w [A2][Cin] -> out [K][A2][T] */
TSPN_CLONES
void tspn_exact_span_head(const float* x, const float* conv_w, const float* conv_b,
                          const float* pred_w, const float* pred_b, float* out,
                          float* hid /* scratch [Cin][T] */, int k, int cin, int t, int a2) {
    for (int p = 0; p < k; ++p) {
        const float* xp = x + (int64_t)p * cin * t;
        for (int co = 0; co < cin; ++co) {
            float* h = hid + (int64_t)co * t;
            for (int f = 0; f < t; ++f) h[f] = conv_b ? conv_b[co] : 0.0f;
            for (int ci = 0; ci < cin; ++ci) {
                const float* xr = xp + (int64_t)ci * t;
                const float* w = conv_w + ((int64_t)co * cin + ci) * 3;
                for (int dt = 0; dt < 3; ++dt) {
                    const float wv = w[dt];
                    /* column f reads x[f + dt - 1]; taps outside [0, T) contribute fma(w, 0, h) == h */
                    int f0 = dt == 0 ? 1 : 0;
                    int f1 = dt == 2 ? t - 1 : t;
                    for (int f = f0; f < f1; ++f) h[f] = fmaf(wv, xr[f + dt - 1], h[f]);
                }
            }
            for (int f = 0; f < t; ++f) h[f] = h[f] > 0.0f ? h[f] : 0.0f;
        }
        for (int j = 0; j < a2; ++j) {
            float* o = out + ((int64_t)p * a2 + j) * t;
            for (int f = 0; f < t; ++f) o[f] = pred_b ? pred_b[j] : 0.0f;
            for (int co = 0; co < cin; ++co) {
                const float wv = pred_w[(int64_t)j * cin + co];
                const float* h = hid + (int64_t)co * t;
                for (int f = 0; f < t; ++f) o[f] = fmaf(wv, h[f], o[f]);
            }
        }
    }
}

/* [SPEC] s5 span decode, reg [K][2A][T] -> spans [K][L*A][2] (int32) */
void tspn_exact_span_decode(const float* reg, const float* sizes, float stride, int32_t* spans,
                            int k, int a, int t, int n_loc) {
    const float CLAMP = 4.1351666f; /* fp32 nearest of log(1000/16) */
    for (int p = 0; p < k; ++p)
        for (int l = 0; l < n_loc; ++l) {
            const float ac = (float)l * stride;
            int col = (int)floorf(ac);
            if (col > t - 1) col = t - 1;
            for (int j = 0; j < a; ++j) {
                const float aw = sizes[j];
                const float dc = reg[((int64_t)p * 2 * a + 2 * j) * t + col];
                float dw = reg[((int64_t)p * 2 * a + 2 * j + 1) * t + col];
                if (dw > CLAMP) dw = CLAMP;
                const float ctr = fmaf(dc, aw, ac);
                const float w = aw * exp_det(dw);
                const float hw = 0.5f * w;
                float lo = floorf((ctr - hw) + 0.5f);
                float hi = floorf((ctr + hw) + 0.5f);
                lo = fminf(fmaxf(lo, 0.0f), (float)(t - 1));
                hi = fminf(fmaxf(hi, lo + 1.0f), (float)t);
                int32_t* o = spans + (((int64_t)p * n_loc + l) * a + j) * 2;
                o[0] = (int32_t)lo;
                o[1] = (int32_t)hi;
            }
        }
}
