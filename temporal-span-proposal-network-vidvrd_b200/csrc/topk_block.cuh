// topk_block.cuh — block-level top-K of up to 2^31 fp32 candidates, K <= 1024.
//
// MSD radix select (8-bit digits on the order-preserving uint32 image of the value, shared
// histogram) finds the K-th largest key; everything above it is collected in any order, the ties
// with the threshold are compacted in index order (warp ballots + running base), and the K
// survivors are bitonic-sorted on the 64-bit composite (~key, index): descending value, ties to
// the lower index ([SPEC] s6 == torch.sort(stable=True)).  Deterministic.
#pragma once

#include "common.cuh"

namespace tspn {

constexpr int TOPK_THREADS = 256;
constexpr int TOPK_MAX_K = 1024;

__device__ __forceinline__ uint32_t order_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);      // larger float <=> larger key
}
constexpr uint32_t KEY_NEG_INF = 0x007FFFFFu;               // order_key(-inf): "not a candidate"

struct TopkSmem {
    uint32_t hist[256];
    uint64_t sel[TOPK_MAX_K];
    uint32_t prefix, need, count, tie_base, n_cand;
    uint32_t warp_cnt[TOPK_THREADS / 32];
};

// Candidate i in [0, total) has value value_of(i); values equal to -inf are not candidates.
// On return sm.sel[0 .. k_eff) holds the winners sorted (low 32 bits = index); returns k_eff =
// min(k, number of candidates).  Must be called by all TOPK_THREADS threads of the block.
template <typename ValueOf>
__device__ int block_topk(TopkSmem& sm, int64_t total, int k, ValueOf value_of) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // -- count candidates ------------------------------------------------------------------------
    if (tid == 0) sm.n_cand = 0;
    __syncthreads();
    {
        uint32_t c = 0;
        for (int64_t i = tid; i < total; i += TOPK_THREADS) c += order_key(value_of(i)) != KEY_NEG_INF;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if (lane == 0 && c) atomicAdd(&sm.n_cand, c);
    }
    __syncthreads();
    const int k_eff = (int)min((uint32_t)k, sm.n_cand);
    if (k_eff == 0) return 0;
    // -- radix select of the k_eff-th largest key ----------------------------------------------------
    if (tid == 0) {
        sm.prefix = 0;
        sm.need = (uint32_t)k_eff;
    }
    uint32_t prefix_mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        sm.hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = sm.prefix;
        for (int64_t i = tid; i < total; i += TOPK_THREADS) {
            const uint32_t key = order_key(value_of(i));
            if (key != KEY_NEG_INF && (key & prefix_mask) == prefix) atomicAdd(&sm.hist[(key >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t need = sm.need, d = 255;
            for (;; --d) {
                const uint32_t c = sm.hist[d];
                if (c >= need) break;
                need -= c;
                if (d == 0) break;
            }
            sm.need = need;
            sm.prefix = prefix | (d << shift);
        }
        prefix_mask |= 0xffu << shift;
        __syncthreads();
    }
    const uint32_t thr = sm.prefix;
    const uint32_t need_ties = sm.need;
    const uint32_t n_above = (uint32_t)k_eff - need_ties;
    if (tid == 0) {
        sm.count = 0;
        sm.tie_base = 0;
    }
    __syncthreads();
    // -- collect ------------------------------------------------------------------------------------------
    for (int64_t base = 0; base < total; base += TOPK_THREADS) {
        const int64_t i = base + tid;
        bool above = false, tie = false;
        uint32_t key = 0;
        if (i < total) {
            key = order_key(value_of(i));
            above = key > thr;                       // thr > KEY_NEG_INF, so -inf is never collected
            tie = key == thr;
        }
        if (above) sm.sel[atomicAdd(&sm.count, 1u)] = ((uint64_t)(~key) << 32) | (uint32_t)i;
        const uint32_t bal = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) sm.warp_cnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = sm.tie_base;
        for (int w = 0; w < warp; ++w) before += sm.warp_cnt[w];
        const uint32_t rank = before + __popc(bal & ((1u << lane) - 1u));
        if (tie && rank < need_ties) sm.sel[n_above + rank] = ((uint64_t)(~key) << 32) | (uint32_t)i;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
            for (int w = 0; w < TOPK_THREADS / 32; ++w) t += sm.warp_cnt[w];
            sm.tie_base += t;
        }
        __syncthreads();
    }
    // -- bitonic sort ----------------------------------------------------------------------------------------
    int m = 1;
    while (m < k_eff) m <<= 1;
    for (int i = k_eff + tid; i < m; i += TOPK_THREADS) sm.sel[i] = ~0ull;
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < m / 2; i += TOPK_THREADS) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint64_t a = sm.sel[lo], b = sm.sel[hi];
                if ((a > b) == up) {
                    sm.sel[lo] = b;
                    sm.sel[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    return k_eff;
}

}  // namespace tspn
