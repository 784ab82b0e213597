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
constexpr int TOPK_CACHE = 8192;          // candidates whose keys are kept in shared memory across the passes

__device__ __forceinline__ uint32_t order_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);      // larger float <=> larger key
}
constexpr uint32_t KEY_NEG_INF = 0x007FFFFFu;               // order_key(-inf): "not a candidate"

struct TopkSmem {
    uint32_t hist[256];
    uint64_t sel[TOPK_MAX_K];
    uint32_t keys[TOPK_CACHE];
    uint32_t prefix, need, count, tie_base, n_cand;
    uint32_t warp_cnt[TOPK_THREADS / 32];
};

// Candidate i in [0, total) has value value_of(i); values equal to -inf are not candidates.
// On return sm.sel[0 .. k_eff) holds the winners sorted (low 32 bits = index); returns k_eff =
// min(k, number of candidates).  Must be called by all TOPK_THREADS threads of the block.

// UNCACHED_BATCH: loads a thread keeps in flight per step of a radix pass when the candidates do not fit the key cache
// (1 for callers whose candidates always fit: their value_of is then not replicated in the code).
// TOPK_EPT: consecutive candidates per thread and step of the collect pass (each step costs a memory round trip and three
// block barriers: kernels that run UNDER the all-pairs kernel see several microseconds per round trip).
template <int UNCACHED_BATCH = 8, int TOPK_EPT = 4, typename ValueOf>
__device__ int block_topk(TopkSmem& sm, int64_t total, int k, ValueOf value_of) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // The candidates are read five times (four radix passes + the collect pass).  Up to TOPK_CACHE of them
    // (one video's N*N relationness scores up to N = 90, or 256 pairs x 20 predicates) are converted to keys
    // once and kept in shared memory; larger problems re-read global memory (L2) every pass.
    const bool cached = total <= TOPK_CACHE;
    if (cached) {
        for (int64_t i = tid; i < total; i += TOPK_THREADS) sm.keys[i] = order_key(value_of(i));
        __syncthreads();
    }
    auto key_of = [&](int64_t i) -> uint32_t { return cached ? sm.keys[i] : order_key(value_of(i)); };
    // -- radix select of the k_eff-th largest key (the first pass also counts the candidates) --------
    if (tid == 0) {
        sm.prefix = 0;
        sm.need = 0;
    }
    uint32_t prefix_mask = 0;
    int k_eff = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        sm.hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = sm.prefix;
        // run-length aggregation: neighbouring scores usually share their leading digits, which
        // would otherwise serialise the shared-memory atomics on one bin
        uint32_t run_bin = 0xffffffffu, run_cnt = 0;
        auto count_key = [&](uint32_t key) {
            if (key != KEY_NEG_INF && (key & prefix_mask) == prefix) {
                const uint32_t bin = (key >> shift) & 0xffu;
                if (bin == run_bin) {
                    ++run_cnt;
                } else {
                    if (run_cnt) atomicAdd(&sm.hist[run_bin], run_cnt);
                    run_bin = bin;
                    run_cnt = 1;
                }
            }
        };
        if (cached) {
            for (int64_t i = tid; i < total; i += TOPK_THREADS) count_key(sm.keys[i]);
        } else {
            // candidates in global memory: UNCACHED_BATCH independent loads in flight per thread, then the (serial)
            // run-length bookkeeping - one load -> use round trip per candidate made a pass over one stress video's
            // 65 536 scores 200 us under the all-pairs kernel (same candidates in the same order: same histogram)
            for (int64_t i0 = tid; i0 < total; i0 += (int64_t)TOPK_THREADS * UNCACHED_BATCH) {
                uint32_t kk[UNCACHED_BATCH];
#pragma unroll
                for (int u = 0; u < UNCACHED_BATCH; ++u) {
                    // no branch around the load (a clamped index, a select afterwards): loads behind branches sit in
                    // separate basic blocks, and ptxas then waits for each one before it issues the next
                    const int64_t i = i0 + (int64_t)u * TOPK_THREADS;
                    const uint32_t key = order_key(value_of(i < total ? i : total - 1));
                    kk[u] = i < total ? key : KEY_NEG_INF;
                }
#pragma unroll
                for (int u = 0; u < UNCACHED_BATCH; ++u) count_key(kk[u]);
            }
        }
        if (run_cnt) atomicAdd(&sm.hist[run_bin], run_cnt);
        __syncthreads();
        {
            // parallel bucket walk: thread t owns bin 255-t; an inclusive scan from the top bin down
            // gives, for every bin, how many candidates sit in it or above it
            const uint32_t mine = sm.hist[255 - tid];
            uint32_t incl = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += up;
            }
            if (lane == 31) sm.warp_cnt[warp] = incl;
            __syncthreads();
            uint32_t before = 0, total_cnt = 0;
            for (int w = 0; w < TOPK_THREADS / 32; ++w) {
                const uint32_t c = sm.warp_cnt[w];
                if (w < warp) before += c;
                total_cnt += c;
            }
            incl += before;
            const uint32_t need = (shift == 24) ? min((uint32_t)k, total_cnt) : sm.need;
            __syncthreads();                        // everyone has read sm.need / warp_cnt
            if (shift == 24 && tid == 0) sm.n_cand = total_cnt;
            // the wanted bin is the first (from the top) whose inclusive count reaches `need`
            if (need > 0 && incl >= need && incl - mine < need) {
                sm.need = need - (incl - mine);      // rank inside the bin
                sm.prefix = prefix | ((uint32_t)(255 - tid) << shift);
            }
            if (need == 0 && tid == 0) sm.need = 0;
        }
        prefix_mask |= 0xffu << shift;
        __syncthreads();
        if (shift == 24) {
            k_eff = (int)min((uint32_t)k, sm.n_cand);
            if (k_eff == 0) return 0;                // uniform: every thread reads the same n_cand
        }
    }
    const uint32_t thr = sm.prefix;
    const uint32_t need_ties = sm.need;
    const uint32_t n_above = (uint32_t)k_eff - need_ties;
    if (tid == 0) {
        sm.count = 0;
        sm.tie_base = 0;
    }
    __syncthreads();
    // -- collect: a thread owns TOPK_EPT consecutive candidates so that ties stay in index order ------
    for (int64_t base = 0; base < total; base += TOPK_THREADS * TOPK_EPT) {
        uint32_t keys[TOPK_EPT];
        uint32_t n_tie = 0;
#pragma unroll
        for (int e = 0; e < TOPK_EPT; ++e) {          // the thread's loads first, all in flight together
            const int64_t i = base + (int64_t)tid * TOPK_EPT + e;
            const uint32_t key = key_of(i < total ? i : total - 1);
            keys[e] = i < total ? key : KEY_NEG_INF;
        }
#pragma unroll
        for (int e = 0; e < TOPK_EPT; ++e) {
            const int64_t i = base + (int64_t)tid * TOPK_EPT + e;
            if (keys[e] > thr) sm.sel[atomicAdd(&sm.count, 1u)] = ((uint64_t)(~keys[e]) << 32) | (uint32_t)i;
            n_tie += keys[e] == thr;                  // thr > KEY_NEG_INF, so -inf is never collected
        }
        // exclusive prefix of the tie counts: within the warp by shuffles, across warps through smem
        uint32_t incl = n_tie;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += up;
        }
        if (lane == 31) sm.warp_cnt[warp] = incl;
        __syncthreads();
        uint32_t rank = sm.tie_base + incl - n_tie;
        for (int w = 0; w < warp; ++w) rank += sm.warp_cnt[w];
#pragma unroll
        for (int e = 0; e < TOPK_EPT; ++e) {
            if (keys[e] == thr) {
                if (rank < need_ties)
                    sm.sel[n_above + rank] = ((uint64_t)(~keys[e]) << 32) | (uint32_t)(base + (int64_t)tid * TOPK_EPT + e);
                ++rank;
            }
        }
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
