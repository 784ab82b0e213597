// geo_windowed.cu — the all-pairs geometry kernel of the WINDOWED layout (tspn_pair_geo_viou_windowed): one WARP per
// tracklet pair, walking only the frames of the pair's temporal overlap window.
//
// Same outputs as the dense kernel of geo_viou.cu (trajectory.py:85-141, common.py:65-106, association.py:35-48 and
// the channels of [SPEC] s2) - the values come from the same device function (geo_math.cuh: geo_step), the volume sums
// are the same 64-bit fixed-point integers - but the rows are stored as [7][Lw] per pair (include/tspn_b200.h).
//
// Why another kernel shape: every channel is zero outside the window, so the windowed layout has ~2.9x fewer bytes to
// write than the dense one on the bench workload.  The dense kernel's shape (a CTA per subject x 32 objects x chunk,
// whole chunks of both tracklets staged by TMA, a thread per 4 frames of the CHUNK) then spends its time on frames that
// produce nothing: 60 % of its warps fall through every object step while the others - 6 of 16 per SM on average -
// carry the latency of the step alone, and all 2048 frames of every object still cross L2 -> shared memory (measured:
// 0.495 ms for 1.45 GB = 0.45 of the HBM peak).  Here the unit of work is the window itself:
//   * a warp owns a pair (pairs are handed out one at a time from a global queue: they cost anything from nothing
//     to T / 128 iterations, and a static split leaves warps idle at the end - 8 pairs per pull: 0.364 ms, 1: 0.323 ms):
//     it reads the two spans, and walks the window [a & ~3, (b + 3) & ~3) in blocks of 128 frames,
//     lane l owning frames 4 l .. 4 l + 3 of the block - the same 4-aligned groups as the dense kernel, so the fp32
//     partial sums (and hence the fixed-point totals) are bit-identical to the dense kernel's;
//   * the block's boxes (129 per tracklet: + 1 halo frame for the forward differences) go global -> shared memory with
//     16-byte cp.async (LDGSTS), coalesced, into a per-warp double buffer with a 128-byte XOR swizzle that makes the
//     lanes' five LDS.128 reads bank-conflict free; block i + 1 is in flight while block i is computed - no block-wide
//     barrier, no TMA descriptor, nothing shared between warps;
//   * every channel leaves as one 128-bit streaming store per lane: a pair's output is ONE contiguous region of
//     7 * Lw floats, written as seven 512-byte-per-iteration streams by a single warp;
//   * the pair's three sums stay in the warp's registers as 64-bit fixed point until the window is done: one writer per
//     pair (chunk slot 0 of the per-(pair, chunk) sums the finalize kernel adds up; the other slots get zeros).
// Algorithmic bytes per pair: 28 * Lw written + 8 (offset) + 24 (sums) + 8 (window).
#include "common.cuh"
#include "geo_math.cuh"

namespace tspn {

constexpr int GW_WARPS = 8;                          // warps (= pairs in flight) per CTA
constexpr int GW_THREADS = GW_WARPS * 32;
constexpr int GW_BLOCK = 32 * GEO_FPT;               // frames per warp iteration
#ifndef TSPN_GW_UNIT
#define TSPN_GW_UNIT 1
#endif
constexpr int GW_UNIT = TSPN_GW_UNIT;                // consecutive pairs per work unit
constexpr int GW_BUF_BYTES = (GW_BLOCK + 8) * 16;    // 128 boxes + halo, whole 128-byte lines (17)
constexpr int GW_WARP_BYTES = 4 * GW_BUF_BYTES;      // {subject, object} x double buffer
constexpr int GW_SMEM_BYTES = GW_WARPS * GW_WARP_BYTES + 128;
#ifndef TSPN_GW_CTAS
#define TSPN_GW_CTAS 2
#endif
#ifndef TSPN_GW_MAXNREG
#define TSPN_GW_MAXNREG 104
#endif

// 16 bytes global -> shared (LDGSTS, cached in L1: the CTA's other warps read the same subject)
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Box j of a per-warp buffer sits at byte 16 * (j ^ ((j >> 3) & 7)): the 16-byte slot inside a 128-byte line is XORed
// with the line index, which makes both the coalesced writes (lane l -> box 32 i + l) and the lanes' reads of their own
// five boxes (4 l .. 4 l + 4) bank-conflict free.  Relative to the buffer (128-byte aligned), so every lane's offsets
// are loop invariants.
__device__ __forceinline__ uint32_t box_off(int j) { return 16u * (uint32_t)(j ^ ((j >> 3) & 7)); }
__device__ __forceinline__ float4 lds_box(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

template <bool CLIP>
__global__ void __maxnreg__(TSPN_GW_MAXNREG)
pair_geo_windowed_kernel(const int64_t* __restrict__ table, int nv, const float4* __restrict__ boxes,
                         const int32_t* __restrict__ span, float* __restrict__ geo, const int64_t* __restrict__ geo_off,
                         unsigned long long* __restrict__ fx, int32_t* __restrict__ overlap,
                         unsigned int* __restrict__ queue, int geo_chunk, int max_chunks) {
    extern __shared__ uint8_t gw_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = ((smem_u32(gw_smem_raw) + 127u) & ~127u) + (uint32_t)warp * GW_WARP_BYTES;
    const int64_t n_pairs = table_total(table, nv, TSPN_VT_PAIR_OFF);
    const int j0 = lane * GEO_FPT;                       // this lane's first box inside a block
    uint32_t wr_off[GEO_FPT], rd_off[GEO_FPT + 1];       // loop invariants: where it writes (copies) and reads
#pragma unroll
    for (int i = 0; i < GEO_FPT; ++i) wr_off[i] = box_off(i * 32 + lane);
#pragma unroll
    for (int i = 0; i <= GEO_FPT; ++i) rd_off[i] = box_off(j0 + i);

    // Work unit = GW_UNIT consecutive pairs, pulled by the warp from a global queue (pairs cost anything between
    // nothing and T / 128 iterations: a static split leaves warps idle at the end).  Lane 0 requests the next unit's
    // number before the current unit is worked on and the warp reads it (the shuffle is what waits for the atomic)
    // only when the unit is done.  The video's table row stays in registers: the queue hands out increasing pair
    // numbers, so a warp's next pair lies in the same video or, usually, the next one - one dependent load instead of
    // a binary search.
    unsigned int pulled = 0;
    if (lane == 0) pulled = atomicAdd(queue, 1u);
    unsigned int unit = __shfl_sync(0xffffffffu, pulled, 0);
    int v = -1;
    int64_t pair_lo = 0, pair_hi = 0, trk_off = 0, box_off0 = 0;     // the current video's row
    int n1 = 1, tb = 0, nchunks = 1;
    while ((int64_t)unit * GW_UNIT < n_pairs) {
    if (lane == 0) pulled = atomicAdd(queue, 1u);
    const int64_t p_end = min((int64_t)(unit + 1) * GW_UNIT, n_pairs);
    for (int64_t p = (int64_t)unit * GW_UNIT; p < p_end; ++p) {
        // ---- the pair (warp-uniform) -----------------------------------------------------------------------
        if (p >= pair_hi) {
            int64_t hi2 = -1;
            if (v >= 0 && v + 1 < nv) hi2 = __ldg(table + (int64_t)(v + 2) * TSPN_VT_COLS + TSPN_VT_PAIR_OFF);
            if (p < hi2) ++v;                                          // the next video (pair_hi <= p < its end)
            else v = find_video(table, nv, TSPN_VT_PAIR_OFF, p);
            const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
            n1 = (int)__ldg(row + TSPN_VT_N) - 1;
            tb = (int)__ldg(row + TSPN_VT_TB);
            trk_off = __ldg(row + TSPN_VT_TRK_OFF);
            box_off0 = __ldg(row + TSPN_VT_BOX_OFF);
            pair_lo = __ldg(row + TSPN_VT_PAIR_OFF);
            pair_hi = __ldg(row + TSPN_VT_COLS + TSPN_VT_PAIR_OFF);     // row v + 1 (the sentinel row after the last)
            nchunks = ((int)__ldg(row + TSPN_VT_T) + geo_chunk - 1) / geo_chunk;
        }
        const int loc = (int)(p - pair_lo);
        const int s = loc / n1, k = loc - s * n1;
        const int o = k + (k >= s ? 1 : 0);
        const int ps = __ldg(span + 2 * (trk_off + s)), pe = __ldg(span + 2 * (trk_off + s) + 1);
        const int qs = __ldg(span + 2 * (trk_off + o)), qe = __ldg(span + 2 * (trk_off + o) + 1);
        const int a = max(ps, qs), b = min(pe, qe);
        const bool has = b > a;
        const int a4 = a & ~3;
        const int lw = has ? ((b + 3) & ~3) - a4 : 0;
        const int nblk = (lw + GW_BLOCK - 1) / GW_BLOCK;
        const float4* sbox = boxes + box_off0 + (int64_t)s * tb + a4;
        const float4* obox = boxes + box_off0 + (int64_t)o * tb + a4;
        float* const grow = geo + __ldg(geo_off + p);

        // boxes of block `it` (window-relative frames [it * 128, it * 128 + 128]) into buffer it & 1.  Frames at or
        // beyond the window's end are not copied: whatever the buffer holds there is read but never used (they
        // lie outside [a, b), and the forward difference of frame t needs t + 1 < b).
        auto fetch = [&](int it) {
            const uint32_t sb = base + (uint32_t)(it & 1) * (2 * GW_BUF_BYTES), ob = sb + GW_BUF_BYTES;
            const int f0 = it * GW_BLOCK + lane;
#pragma unroll
            for (int i = 0; i < GEO_FPT; ++i) {
                if (f0 + i * 32 < lw) {
                    cp_async_16(sb + wr_off[i], sbox + f0 + i * 32);
                    cp_async_16(ob + wr_off[i], obox + f0 + i * 32);
                }
            }
            if (lane == 0 && f0 + GW_BLOCK < lw) {                     // halo: first frame of the next block
                cp_async_16(sb + box_off(GW_BLOCK), sbox + f0 + GW_BLOCK);
                cp_async_16(ob + box_off(GW_BLOCK), obox + f0 + GW_BLOCK);
            }
            cp_async_commit();
        };

        unsigned long long acc_i = 0ull, acc_s = 0ull, acc_o = 0ull;
        if (nblk > 0) fetch(0);
        for (int it = 0; it < nblk; ++it) {
            if (it + 1 < nblk) {
                fetch(it + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();                                              // every lane's copies of block `it` have landed
            const uint32_t sb = base + (uint32_t)(it & 1) * (2 * GW_BUF_BYTES), ob = sb + GW_BUF_BYTES;
            const int rel = it * GW_BLOCK + j0;                        // window-relative first frame of this lane
            float fsum_i = 0.0f, fsum_s = 0.0f, fsum_o = 0.0f;
            float out[TSPN_GEO_CHANNELS][GEO_FPT];
            auto ld_s = [&](int j) { return lds_box(sb + rd_off[j - j0]); };
            auto ld_o = [&](int j) { return lds_box(ob + rd_off[j - j0]); };
            // a block strictly inside the window (every lane: a <= t0 and t0 + 4 < b) takes the predicate-free form of
            // the step; the window's first block (unless a is a multiple of 4) and its last one the general form
            const bool interior = (it > 0 || a == a4) && a4 + (it + 1) * GW_BLOCK < b;
            if (interior) {
                geo_step<CLIP, true>(ld_s, ld_o, j0, a4 + rel, a, b, out, fsum_i, fsum_s, fsum_o);
            } else if (rel < lw) {
                geo_step<CLIP, false>(ld_s, ld_o, j0, a4 + rel, a, b, out, fsum_i, fsum_s, fsum_o);
            }
            if (rel < lw) {
                float* gr = grow + rel;
#pragma unroll
                for (int ch = 0; ch < TSPN_GEO_CHANNELS - 1; ++ch)
                    st_stream_f4(gr + (int64_t)ch * lw, make_float4(out[ch][0], out[ch][1], out[ch][2], out[ch][3]));
                // the dense kernel's conversion, per 4-frame partial: same integers, whatever the order of the adds
                acc_i += __float2ull_rn(fsum_i * 65536.0f);
                if (CLIP) {
                    acc_s += __float2ull_rn(fsum_s * 65536.0f);
                    acc_o += __float2ull_rn(fsum_o * 65536.0f);
                }
            }
            __syncwarp();                                              // buffer it & 1 is free for block it + 2
        }
        acc_i = warp_sum_u64(acc_i);
        if (CLIP) {
            acc_s = warp_sum_u64(acc_s);
            acc_o = warp_sum_u64(acc_o);
        }
        if (lane == 0) {
            unsigned long long* slot = fx + p * max_chunks * 3;
            slot[0] = acc_i;
            slot[1] = acc_s;
            slot[2] = acc_o;
            for (int c = 1; c < nchunks; ++c) {
                slot[3 * c] = 0ull;
                slot[3 * c + 1] = 0ull;
                slot[3 * c + 2] = 0ull;
            }
            *reinterpret_cast<int2*>(overlap + 2 * p) = make_int2(has ? a : 0, has ? b : 0);
        }
    }
    unit = __shfl_sync(0xffffffffu, pulled, 0);
    }
}

// MAIN phase of tspn_pair_geo_viou_windowed (called by geo_viou.cu).  `reserve`: CTA slots (of TSPN_GW_CTAS per SM)
// left to concurrent streams - the grid is persistent (grid-stride over the pairs), so whatever it does not occupy
// stays free for the side branches for the whole launch.
__global__ void reset_unit_queue_kernel(unsigned int* __restrict__ queue) {
    if (threadIdx.x == 0) *queue = 0u;
}

int launch_pair_geo_windowed(const int64_t* d_table, int num_videos, int64_t total_pairs, const float* d_boxes,
                             const int32_t* d_span, float* d_geo, const int64_t* d_geo_off, unsigned long long* fx,
                             int32_t* d_overlap, unsigned int* d_queue, int geo_chunk, int max_chunks, int reserve,
                             bool clip, cudaStream_t st) {
    int64_t grid = (int64_t)num_sms() * TSPN_GW_CTAS - reserve;
    const int64_t need = (total_pairs + GW_WARPS * GW_UNIT - 1) / (GW_WARPS * GW_UNIT);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    reset_unit_queue_kernel<<<1, 32, 0, st>>>(d_queue);   // (a kernel, not a memset node: see reset_queue_kernel)
#define TSPN_LAUNCH_GW(C)                                                                                          \
    do {                                                                                                           \
        TSPN_CUDA_OK(cudaFuncSetAttribute(pair_geo_windowed_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          GW_SMEM_BYTES));                                                         \
        prefer_max_smem(pair_geo_windowed_kernel<C>);                                                              \
        pair_geo_windowed_kernel<C><<<(unsigned)grid, GW_THREADS, GW_SMEM_BYTES, st>>>(                            \
            d_table, num_videos, reinterpret_cast<const float4*>(d_boxes), d_span, d_geo, d_geo_off, fx, d_overlap, \
            d_queue, geo_chunk, max_chunks);                                                                               \
    } while (0)
    if (clip) TSPN_LAUNCH_GW(true); else TSPN_LAUNCH_GW(false);
#undef TSPN_LAUNCH_GW
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // namespace tspn
