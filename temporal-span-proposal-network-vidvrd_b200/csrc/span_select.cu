// span_select.cu — temporal NMS + top-n of the decoded span proposals of every scored pair ([SPEC] s8).
//
//   tspn_span_select   what RelNMS is meant to do (lib/modeling/relpn/rel_nms.py:6-15: nms_threshold 0.5,
//                      top_k_proposals = RELPN.DPN.NUM_DURATION_PROPOSALS, lib/config/defaults.py:62) - the
//                      reference's forward is a stub, so the rule is defined here and in oracle/heads.py:
//     candidate i of a pair = decoded span [s_i, e_i) (anchor-location major, anchor minor: the decode's order);
//     its rank key is q_i = floor(2^15 * |span_i ^ W| / |span_i v W|) against the pair's temporal overlap window
//     W (a relation can only hold while both tracklets exist), ties to the lower candidate index;
//     greedy: repeatedly keep the best live candidate and drop every live candidate whose temporal IoU with it
//     exceeds nms_thr (integers: inter * 1024 > thr_q10 * union), until n_keep are kept or none is left.
//   Output: the kept spans in keep order, int16 (or int32) [n_rows][n_keep][2], zero padded, and their count.
//
// Integer arithmetic only: bit-exact against the oracle by construction.
//
// One warp per pair.  Candidates live in shared memory in an ANCHOR-major layout, 32 per group, so that the 32
// candidates of a group have similar positions and lengths: lane g keeps group g's bounds (min start, max end,
// min / max length) and its best live key.  An iteration is then: arg-max of 32 lane values (one REDUX), a
// bounds test per group (one ballot), and the exact test only on the 2-3 groups a kept span can reach - instead
// of 64 x 504 pair tests per row.  The grouping is only a speed-up: the bounds are data-derived, any candidate
// set gives the oracle's result.
#include "common.cuh"

namespace tspn {

constexpr int SS_WARPS = 4;
constexpr int SS_MAX_GPL = 4;              // groups per lane: up to 128 groups = 4096 candidate slots
constexpr int SS_MAX_KEEP = 256;

__host__ __device__ __forceinline__ int ss_locations(int t, float stride) {
    const double q = ((double)t + 1.0) / (double)stride;
    int n = (int)q;
    if ((double)n < q) ++n;
    return n;
}

// AFIX: the anchor count as a compile-time constant (4, the reference's NUM_ANCHORS_PER_LOCATION) or 0 = run time:
// the candidate index is split into (location, anchor) once per iteration, and the kernel is issue-bound
template <int GPL, bool OUT16, int AFIX>
__global__ void __launch_bounds__(SS_WARPS * 32)
span_select_kernel(const int64_t* __restrict__ table, int nv, const int32_t* __restrict__ trk_span,
                   const int64_t* __restrict__ rows, int64_t n_rows,
                   const int32_t* __restrict__ windows, const int32_t* __restrict__ cand, int64_t ld_cand,
                   int n_cand_fixed, int n_anchors, float stride, int n_keep, int thr_q10, int slots_per_warp,
                   void* __restrict__ out, int32_t* __restrict__ counts) {
    extern __shared__ uint32_t ss_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * SS_WARPS + warp;
    if (r >= n_rows) return;
    uint32_t* const keys = ss_smem + (size_t)warp * (2 * slots_per_warp + SS_MAX_KEEP);
    uint32_t* const se = keys + slots_per_warp;
    uint32_t* const kept_se = se + slots_per_warp;

    // ---- the row: its pair, its video, its window, its candidate count -------------------------------
    const int64_t gp = rows ? rows[r] : r;
    int wa = 0, wb = 0, n_loc = 0;
    const int A = AFIX ? AFIX : n_anchors;
    bool live = gp >= 0;
    if (table) {
        if (live && gp >= table_total(table, nv, TSPN_VT_PAIR_OFF)) live = false;      // beyond the batch (capacity grid)
        if (live) {
            const int v = find_video(table, nv, TSPN_VT_PAIR_OFF, gp);
            const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
            const int n = (int)row[TSPN_VT_N];
            n_loc = ss_locations((int)row[TSPN_VT_T], stride);
            if (!windows) {
                const int p = (int)(gp - row[TSPN_VT_PAIR_OFF]);
                const int s = p / (n - 1);
                const int k = p - s * (n - 1);
                const int o = k + (k >= s ? 1 : 0);
                const int64_t ts = row[TSPN_VT_TRK_OFF] + s, to = row[TSPN_VT_TRK_OFF] + o;
                const int a = max(__ldg(trk_span + 2 * ts), __ldg(trk_span + 2 * to));
                const int b = min(__ldg(trk_span + 2 * ts + 1), __ldg(trk_span + 2 * to + 1));
                if (b > a) { wa = a; wb = b; }
            }
        }
    } else {
        n_loc = n_cand_fixed / A;
    }
    if (windows && live) {
        wa = __ldg(windows + 2 * r);
        wb = __ldg(windows + 2 * r + 1);
        if (wb <= wa) wa = wb = 0;
    }
    const int wlen = wb - wa;
    const int gpa = (n_loc + 31) >> 5;                    // groups per anchor
    const int G = live ? gpa * A : 0;

    // ---- load: slot (group j, lane) = anchor j / gpa, location (j % gpa) * 32 + lane ----------------------
    uint32_t gkey[GPL];
    int g_smin[GPL], g_emax[GPL], g_lmin[GPL], g_lmax[GPL];
#pragma unroll
    for (int u = 0; u < GPL; ++u) { gkey[u] = 0u; g_smin[u] = 0x7fffffff; g_emax[u] = 0; g_lmin[u] = 0x7fffffff; g_lmax[u] = 0; }
    const int32_t* crow = cand + r * ld_cand;
    // pass 1: every load of the lane in flight at once (the row's candidates sit in L2: one round trip, not G)
    for (int a = 0, j = 0; a < (G ? A : 0); ++a) {
        for (int l = lane; l < gpa * 32; l += 32, ++j) {      // group j = a * gpa + l / 32 (no division: issue-bound)
            uint32_t key = 0u, pack = 0u;
            if (l < n_loc) {
                const int i = l * A + a;
                const int2 c = __ldg(reinterpret_cast<const int2*>(crow) + i);
                const int inter = max(0, min(c.y, wb) - max(c.x, wa));
                const int uni = (c.y - c.x) + wlen - inter;
                const uint32_t q = uni > 0 ? ((uint32_t)inter << 15) / (uint32_t)uni : 0u;
                key = (1u << 28) | (q << 12) | (uint32_t)(4095 - i);
                pack = (uint32_t)c.x | ((uint32_t)c.y << 16);
            }
            keys[j * 32 + lane] = key;
            se[j * 32 + lane] = pack;
        }
    }
    // pass 2: the groups' bounds and best keys (each lane re-reads its own slots: no barrier needed)
    for (int j = 0; j < G; ++j) {
        const uint32_t key = keys[j * 32 + lane], pack = se[j * 32 + lane];
        const int s = key ? (int)(pack & 0xffffu) : 0x7fffffff, e = key ? (int)(pack >> 16) : 0;
        const int len_lo = key ? e - s : 0x7fffffff, len_hi = key ? e - s : 0;
        const uint32_t gk = __reduce_max_sync(0xffffffffu, key);
        const int smin = __reduce_min_sync(0xffffffffu, s), emax = __reduce_max_sync(0xffffffffu, e);
        const int lmin = __reduce_min_sync(0xffffffffu, len_lo), lmax = __reduce_max_sync(0xffffffffu, len_hi);
        if (lane == (j & 31)) {
#pragma unroll
            for (int u = 0; u < GPL; ++u)
                if (u == (j >> 5)) { gkey[u] = gk; g_smin[u] = smin; g_emax[u] = emax; g_lmin[u] = lmin; g_lmax[u] = lmax; }
        }
    }
    __syncwarp();

    // ---- greedy selection ---------------------------------------------------------------------------
    int kept = 0;
    while (kept < n_keep) {
        uint32_t m = gkey[0];
#pragma unroll
        for (int u = 1; u < GPL; ++u) m = max(m, gkey[u]);
        m = __reduce_max_sync(0xffffffffu, m);
        if (m == 0u) break;
        const int i = 4095 - (int)(m & 4095u);
        const int l = i / A, a = i - l * A;               // shifts when A is the compile-time 4
        const int jw = a * gpa + (l >> 5);
        const int pw = jw * 32 + (l & 31);
        const uint32_t w = se[pw];
        const int sw = (int)(w & 0xffffu), ew = (int)(w >> 16), lw = ew - sw;
        if (lane == 0) kept_se[kept] = w;
        ++kept;
        if (lane == (pw & 31)) keys[pw] = 0u;            // the winner leaves the live set whatever the threshold
        __syncwarp();
        // which groups can hold a candidate with tIoU(candidate, winner) > thr: the group's hull must overlap the
        // winner, and min(len) / max(len) > thr must be possible for some length in [lmin, lmax]
#pragma unroll
        for (int u = 0; u < GPL; ++u) {
            const bool hit = gkey[u] != 0u && ew > g_smin[u] && sw < g_emax[u] &&
                             g_lmax[u] * 1024 > thr_q10 * lw && lw * 1024 > thr_q10 * min(g_lmin[u], 65536);
            uint32_t mask = __ballot_sync(0xffffffffu, hit);
            if (u == (jw >> 5)) mask |= 1u << (jw & 31);                       // the winner's own group
            while (mask) {
                const int jl = __ffs(mask) - 1;
                mask &= mask - 1u;
                const int j = u * 32 + jl;
                uint32_t key = keys[j * 32 + lane];
                if (key) {
                    const uint32_t c = se[j * 32 + lane];
                    const int s = (int)(c & 0xffffu), e = (int)(c >> 16);
                    const int inter = min(e, ew) - max(s, sw);
                    const int uni = (e - s) + lw - inter;
                    if (inter > 0 && inter * 1024 > thr_q10 * uni) {          // frames < 2^16: products < 2^27
                        key = 0u;
                        keys[j * 32 + lane] = 0u;
                    }
                }
                const uint32_t gk = __reduce_max_sync(0xffffffffu, key);
                if (lane == jl) gkey[u] = gk;
            }
        }
    }
    __syncwarp();

    // ---- output ---------------------------------------------------------------------------------------
    if (lane == 0 && counts) counts[r] = kept;
    if (OUT16) {
        uint32_t* o = reinterpret_cast<uint32_t*>(out) + r * n_keep;             // (s, e) int16 = one 32-bit word
        for (int k = lane; k < n_keep; k += 32) o[k] = k < kept ? kept_se[k] : 0u;
    } else {
        int2* o = reinterpret_cast<int2*>(out) + r * n_keep;
        for (int k = lane; k < n_keep; k += 32) {
            const uint32_t w = k < kept ? kept_se[k] : 0u;
            o[k] = make_int2((int)(w & 0xffffu), (int)(w >> 16));
        }
    }
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int tspn_span_select(const int64_t* d_table, int num_videos, int max_frames, const int32_t* d_span,
                     const int64_t* d_rows, int64_t n_rows, const int32_t* d_windows,
                     const int32_t* d_cand, int64_t ld_cand, int n_cand, int n_anchors, float stride, int n_keep,
                     float nms_threshold, int flags, void* d_out, int32_t* d_counts, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(n_rows >= 0 && n_anchors > 0 && n_keep > 0 && stride > 0.0f, TSPN_EBADARG, "tspn_span_select: bad size");
    TSPN_REQUIRE(n_keep <= SS_MAX_KEEP, TSPN_ESHAPE, "tspn_span_select: n_keep=%d exceeds %d", n_keep, SS_MAX_KEEP);
    TSPN_REQUIRE(nms_threshold > 0.0f && nms_threshold <= 1.0f, TSPN_EBADARG, "tspn_span_select: nms_threshold in (0, 1]");
    if (n_rows == 0) return TSPN_OK;
    TSPN_REQUIRE(d_cand && d_out, TSPN_EBADARG, "tspn_span_select: null pointer");
    TSPN_REQUIRE(d_table ? (d_windows || d_span) && max_frames > 0 : (d_windows && n_cand > 0 && n_cand % n_anchors == 0),
                 TSPN_EBADARG, "tspn_span_select: needs the table + tracklet spans (or windows), or windows + n_cand");
    TSPN_REQUIRE((reinterpret_cast<uintptr_t>(d_cand) & 7) == 0 && (ld_cand & 1) == 0, TSPN_EALIGN,
                 "tspn_span_select: candidates must be 8-byte aligned pairs");
    TSPN_REQUIRE(max_frames < 65536, TSPN_ESHAPE, "tspn_span_select: frame bounds must fit 16 bits (T=%d)", max_frames);
    const int n_loc = d_table ? ss_locations(max_frames, stride) : n_cand / n_anchors;
    TSPN_REQUIRE((int64_t)n_loc * n_anchors <= 4096, TSPN_ESHAPE,
                 "tspn_span_select: %d locations x %d anchors exceed 4096 candidates per pair", n_loc, n_anchors);
    TSPN_REQUIRE(ld_cand >= (int64_t)n_loc * n_anchors * 2 || d_table, TSPN_ESHAPE, "tspn_span_select: ld_cand too small");
    const int groups = ((n_loc + 31) / 32) * n_anchors;
    const int gpl = (groups + 31) / 32;
    TSPN_REQUIRE(gpl <= SS_MAX_GPL, TSPN_ESHAPE, "tspn_span_select: too many candidate groups (%d)", groups);
    const int slots = groups * 32;
    const size_t smem = (size_t)SS_WARPS * (2 * slots + SS_MAX_KEEP) * sizeof(uint32_t);
    TSPN_REQUIRE(smem <= 200 * 1024, TSPN_ESHAPE, "tspn_span_select: candidate set too large for shared memory");
    const int thr_q10 = (int)(nms_threshold * 1024.0f + 0.5f);
    const bool out16 = (flags & TSPN_SPANS_I16) != 0;
    const unsigned blocks = (unsigned)((n_rows + SS_WARPS - 1) / SS_WARPS);
    cudaStream_t st = (cudaStream_t)stream;
#define TSPN_LAUNCH_SS(GPLV, O16, AF)                                                                                 \
    do {                                                                                                             \
        TSPN_CUDA_OK(cudaFuncSetAttribute(span_select_kernel<GPLV, O16, AF>,                                         \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
        prefer_max_smem(span_select_kernel<GPLV, O16, AF>);                                                          \
        span_select_kernel<GPLV, O16, AF><<<blocks, SS_WARPS * 32, smem, st>>>(                                      \
            d_table, num_videos, d_span, d_rows, n_rows, d_windows, d_cand, ld_cand, n_cand,                        \
            n_anchors, stride, n_keep, thr_q10, slots, d_out, d_counts);                                             \
    } while (0)
    if (gpl == 1 && n_anchors == 4) { if (out16) TSPN_LAUNCH_SS(1, true, 4); else TSPN_LAUNCH_SS(1, false, 4); }
    else if (gpl == 1) { if (out16) TSPN_LAUNCH_SS(1, true, 0); else TSPN_LAUNCH_SS(1, false, 0); }
    else if (gpl == 2) { if (out16) TSPN_LAUNCH_SS(2, true, 0); else TSPN_LAUNCH_SS(2, false, 0); }
    else { if (out16) TSPN_LAUNCH_SS(4, true, 0); else TSPN_LAUNCH_SS(4, false, 0); }
#undef TSPN_LAUNCH_SS
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // extern "C"
