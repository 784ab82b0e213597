// geo_viou.cu — all-pairs per-frame geometry + trajectory vIoU (the HBM-write-bound kernel).
//
// Replaces, for every ordered tracklet pair of every video of a batch:
//   _intersect/_union/cubic_iou  lib/modeling/trajectory.py:85-141          (V1)
//   viou                         lib/evaluation/common.py:65-106            (V2, default)
//   _traj_iou                    lib/modeling/association.py:35-48          (V3, CLIP)
// and produces the per-frame channels of [SPEC] s2: box-centre deltas,
// log-scale ratios, per-frame IoU, forward differences, temporal-overlap mask.
//
// Design (sm_100a):
//   * work item = (video, subject s, group of <= 64 objects, chunk of 512 / 1024 / 2048 frames);
//     one CTA per item;
//   * the two tracklets' box rows are staged chunk by chunk (chunk + 1 halo row of 8 boxes) into
//     shared memory by 2-D tiled TMA (cp.async.bulk.tensor -> UTMALDG) with SWIZZLE_128B into a
//     ring on full/empty mbarriers: the subject chunk is loaded once per item, the object chunks
//     stream behind the compute of the previous objects;
//   * a thread owns 4 consecutive frames: the swizzle makes its five LDS.128 box reads
//     bank-conflict free, and every channel leaves as one 128-bit streaming store (a warp writes
//     512 contiguous bytes per channel row);
//   * intersection volumes accumulate as 64-bit fixed point (exact for integer boxes; integer
//     adds -> the result does not depend on any reduction or scheduling order);
//   * per-tracklet volumes come from a small pre-kernel (one CTA per tracklet), vIoU / tIoU from a
//     small post-kernel; the pair kernel itself writes the overlap windows.
// Algorithmic bytes: 32*Tp written + 24 B of reductions per pair; boxes are re-read from L2.
#include "common.cuh"
#include "geo_math.cuh"

namespace tspn {

constexpr int GEO_OG = TSPN_GEO_OBJ_GROUP;
#ifndef TSPN_GEO_RING
#define TSPN_GEO_RING 3
#endif
constexpr int GEO_RING = TSPN_GEO_RING;               // object-chunk stages in flight (non-dense shape)
#ifndef TSPN_GEO_ROTATE
#define TSPN_GEO_ROTATE 1
#endif
constexpr bool GEO_ROTATE = TSPN_GEO_ROTATE != 0;     // rotate the warps' frame blocks with the object index
#ifndef TSPN_GEO_CTAS256
#define TSPN_GEO_CTAS256 2                            // CTAs per SM of the 256-thread (1024-frame chunk) shape
#endif
#ifndef TSPN_GEO_CTAS128
#define TSPN_GEO_CTAS128 3                            // ... and of the 128-thread (512-frame chunk) shape
#endif

// Shape of one CTA: THREADS threads cover a chunk of 4*THREADS frames (512 / 1024 / 2048, chosen per
// batch by tspn_geo_chunk).  HBM absorbs this kernel's store stream best as few, wide streams
// (tools/bench_store_pattern.cu, store-only kernels with this address pattern: 5.8 TB/s with 2 KB row
// segments, 7.0-7.4 TB/s with whole 8 KB rows), so a CTA writes row segments as long as the video allows.
// Two occupancy shapes:
//   DENSE = false  (default) ~512 threads per SM (one 512-thread CTA), 3-stage object ring, 103 registers;
//   DENSE = true   (flag TSPN_GEO_DENSE_CTAS) 1024 threads per SM (two 512-thread CTAs), 2-stage ring, 64
//                  registers (64 B of spills): twice the warps to cover the LDS / MUFU / TMA latencies
//                  between a warp's store bursts.  Measured on the bench workload: 0.919 ms against
//                  0.707 ms for the default - kept, bit-identical and tested, as the record of that A/B
//                  (profiles/r1_geo_kernel_forms_ab.md).
// The shared-memory request pins the occupancy.
template <int THREADS, bool DENSE>
struct GeoCfg {
    static constexpr int CHUNK = THREADS * GEO_FPT;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int RING = DENSE ? 2 : GEO_RING;             // object-chunk stages in flight
    static constexpr int STAGES = 1 + RING;                       // subject chunk + object ring
    static constexpr int ROWS = CHUNK / 8 + 1;                    // rows of 8 boxes + 1 halo row
    static constexpr int SPLIT = ROWS > 256 ? 2 : 1;              // a TMA box has at most 256 rows
    static constexpr int BOX_ROWS = SPLIT == 1 ? ROWS : CHUNK / 16 + 1;   // the second box re-reads one row
    static constexpr int TX_BYTES = SPLIT * BOX_ROWS * 128;       // bytes landing per staged chunk
    static constexpr int STAGE_BYTES = ROWS * 128;                // stages are packed (128-byte aligned)
    static constexpr int MIN_CTAS = DENSE ? 1024 / THREADS
                                          : (THREADS >= 512 ? 1 : (THREADS == 256 ? TSPN_GEO_CTAS256 : TSPN_GEO_CTAS128));
    // barriers, per-object fixed-point sums, per-object overlap windows
    static constexpr int TAIL_BYTES = (2 * RING * 8 + GEO_OG * (3 * 8 + 8) + 16 + 127) / 128 * 128;
    static constexpr int SMEM_USED = STAGES * STAGE_BYTES + TAIL_BYTES;
    // the 512-thread shape is already limited to one CTA per SM by its registers (16 warps x 104): no padding,
    // every byte it does not use is left to co-resident kernels of the side stream
    static constexpr int SMEM_PIN = (THREADS >= 512 && !DENSE) ? 0 : 227 * 1024 / (MIN_CTAS + 1) + 1024;
    static constexpr int SMEM_BYTES = SMEM_USED > SMEM_PIN ? SMEM_USED : SMEM_PIN;
    static_assert((SMEM_BYTES + 1024) * MIN_CTAS <= 228 * 1024, "the stages of MIN_CTAS CTAs must fit one SM");
};


__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// ---- pre-kernel: per-tracklet volume (sum over [pstart, pend) of w*h in fp64; one 128-thread CTA per
// tracklet, four independent loads in flight per thread, fixed combination order) --------------------
__global__ void __launch_bounds__(128) tracklet_volume_kernel(const int64_t* __restrict__ table, int nv,
                                                              const float4* __restrict__ boxes,
                                                              const int32_t* __restrict__ span,
                                                              double* __restrict__ vol) {
    __shared__ double part[4];
    const int64_t total_tracklets = table_total(table, nv, TSPN_VT_TRK_OFF);
    for (int64_t trk = blockIdx.x; trk < total_tracklets; trk += gridDim.x) {
        const int v = find_video(table, nv, TSPN_VT_TRK_OFF, trk);
        const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
        const int64_t n_local = trk - row[TSPN_VT_TRK_OFF];
        const float4* b = boxes + row[TSPN_VT_BOX_OFF] + n_local * row[TSPN_VT_TB];
        const int ps = span[2 * trk], pe = span[2 * trk + 1];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int t = ps + (int)threadIdx.x;
        for (; t + 384 < pe; t += 512) {
            const float4 q0 = __ldg(b + t), q1 = __ldg(b + t + 128), q2 = __ldg(b + t + 256), q3 = __ldg(b + t + 384);
            a0 += (double)(((q0.z - q0.x) + 1.0f) * ((q0.w - q0.y) + 1.0f));
            a1 += (double)(((q1.z - q1.x) + 1.0f) * ((q1.w - q1.y) + 1.0f));
            a2 += (double)(((q2.z - q2.x) + 1.0f) * ((q2.w - q2.y) + 1.0f));
            a3 += (double)(((q3.z - q3.x) + 1.0f) * ((q3.w - q3.y) + 1.0f));
        }
        for (; t < pe; t += 128) {
            const float4 q = __ldg(b + t);
            a0 += (double)(((q.z - q.x) + 1.0f) * ((q.w - q.y) + 1.0f));
        }
        const double w = warp_sum((a0 + a1) + (a2 + a3));
        __syncthreads();                               // the previous tracklet's part[] has been read
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = w;
        __syncthreads();
        if (threadIdx.x == 0) vol[trk] = (part[0] + part[1]) + (part[2] + part[3]);
    }
}

// ---- the pair kernel ----------------------------------------------------------------------------
// Work item = (video, subject s, group of GEO_OG objects, chunk of GEO_CHUNK frames); one CTA per
// item, chunk-fastest in the grid.  The subject chunk is staged once, the object chunks stream
// through a 2-deep TMA ring (full/empty mbarriers, no block-wide barrier in the loop).
//
// Volume sums are accumulated as 64-bit fixed point (2^-16 units): a thread's fp32 partial over its 4
// frames (exact: integer-valued products below 2^24) is converted once, the warp total is two
// REDUX.SUM on 24-bit limbs, lane 0 adds it to the pair's shared-memory accumulator and the CTA adds
// its chunk's total to the pair's global accumulator, all with integer atomics.  Integer addition is
// associative, so the sums - exact for integer boxes, quantised at 2^-17 absolute per partial
// otherwise - do not depend on any reduction or scheduling order.
constexpr float GEO_FX_SCALE = 65536.0f;

__device__ __forceinline__ unsigned long long warp_sum_fx(float v) {
    const unsigned long long fx = __float2ull_rn(v * GEO_FX_SCALE);
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(fx & 0xffffffull));       // < 2^29
    const unsigned hi = __reduce_add_sync(0xffffffffu, (unsigned)(fx >> 24));               // < 2^27
    return (unsigned long long)lo + ((unsigned long long)hi << 24);
}


template <int THREADS, bool WRITE_GEO, bool CLIP, bool DENSE>
__device__ __forceinline__ void
pair_geo_body(const CUtensorMap& box_map, const int64_t* __restrict__ table, int nv,
              const int32_t* __restrict__ span, float* __restrict__ geo, unsigned long long* __restrict__ fx,
              int32_t* __restrict__ overlap, unsigned int* __restrict__ queue, int max_chunks) {
    using Cfg = GeoCfg<THREADS, DENSE>;
    constexpr int GEO_CHUNK = Cfg::CHUNK, GEO_WARPS = Cfg::WARPS, GEO_STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr int GEO_TX_BYTES = Cfg::TX_BYTES, GEO_SPLIT = Cfg::SPLIT, RING = Cfg::RING;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* const s_stage = smem;
    uint8_t* const o_stage0 = smem + GEO_STAGE_BYTES;
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * GEO_STAGE_BYTES);   // [RING]
    uint64_t* const empty = full + RING;                                                       // [RING]
    unsigned long long* const acc = reinterpret_cast<unsigned long long*>(empty + RING);       // [OG][3]
    int2* const owin = reinterpret_cast<int2*>(acc + GEO_OG * 3);                              // [OG] overlap windows
    unsigned int* const s_item = reinterpret_cast<unsigned int*>(owin + GEO_OG);               // next work item

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < RING; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], GEO_WARPS);
        }
        fence_mbar_init();
    }
    const uint32_t ss = smem_u32(s_stage);
    // Ring steps of this CTA so far.  Step g uses stage g % RING in its (g / RING)-th use: the consumers wait for
    // `full` with parity (g / RING) & 1, the producer refills a stage once `empty` has completed its previous use.
    // The count runs on across work items, so a persistent CTA never re-initialises a barrier.
    unsigned int g0 = 0;

    // One work item per CTA (queue == nullptr: item = blockIdx.x), or PERSISTENT CTAs - one per SM slot for the
    // whole launch - that pull items from a global queue.  Persistent CTAs never leave their SM, so kernels of a
    // concurrent stream can only ever co-reside with them (in the registers / shared memory they leave free)
    // instead of taking over SMs between two CTAs and locking the next one out.
    for (;;) {
    unsigned int item;
    if (queue) {
        if (tid == 0) *s_item = atomicAdd(queue, 1u);
        __syncthreads();
        item = *s_item;
    } else {
        item = blockIdx.x;
    }
    // the batch's true work-item count (sentinel row): the grid / queue bound of the launch is only a capacity
    // (re-read per item rather than held in a register: the kernel sits exactly at its register budget)
    if (item >= (unsigned int)table_total(table, nv, TSPN_VT_ITEM_OFF)) break;

    // ---- decode the work item ------------------------------------------------------------------
    const int v = find_video(table, nv, TSPN_VT_ITEM_OFF, (int64_t)item);
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const int n = (int)row[TSPN_VT_N];
    const int t_len = (int)row[TSPN_VT_T];
    const int tp = (int)row[TSPN_VT_TP];
    const int64_t tb = row[TSPN_VT_TB];
    const int64_t trk_off = row[TSPN_VT_TRK_OFF];
    const int groups = (n - 1 + GEO_OG - 1) / GEO_OG;
    const int nchunks = (t_len + GEO_CHUNK - 1) / GEO_CHUNK;
    const int local = (int)((int64_t)item - row[TSPN_VT_ITEM_OFF]);
    const int c = local % nchunks;
    const int sg = local / nchunks;
    const int s = sg / groups;
    const int k0 = (sg - s * groups) * GEO_OG;
    const int nobj = min(GEO_OG, n - 1 - k0);
    const int64_t box_row0 = row[TSPN_VT_BOX_OFF];           // multiple of 8
    const int64_t pair0 = row[TSPN_VT_PAIR_OFF] + (int64_t)s * (n - 1) + k0;

    if (tid < nobj) {       // the loop below touches no global or local memory besides its stores
        const int ps = __ldg(span + 2 * (trk_off + s)), pe = __ldg(span + 2 * (trk_off + s) + 1);
        const int k = k0 + tid;
        const int o = k + (k >= s ? 1 : 0);
        const int qs = __ldg(span + 2 * (trk_off + o)), qe = __ldg(span + 2 * (trk_off + o) + 1);
        owin[tid] = make_int2(max(ps, qs), min(pe, qe));
    }
    for (int i = tid; i < GEO_OG * 3; i += THREADS) acc[i] = 0ull;
    __syncthreads();

    auto issue = [&](int q) {          // thread 0: object q (and, with the first, the subject chunk)
        const int k = k0 + q;
        const int o = k + (k >= s ? 1 : 0);
        const unsigned int g = g0 + (unsigned int)q;
        const int st = (int)(g % RING);
        // the stage's previous use (ring step g - RING) has been read by every warp
        if (g >= (unsigned int)RING) mbar_wait(&empty[st], ((g / RING) - 1u) & 1u);
        mbar_expect_tx(&full[st], q == 0 ? 2 * GEO_TX_BYTES : GEO_TX_BYTES);
#pragma unroll
        for (int h = 0; h < GEO_SPLIT; ++h) {                 // second half: rows CHUNK/16 .. CHUNK/8 (+ halo)
            const int r_off = h * (GEO_CHUNK / 16);
            if (q == 0)
                tma_load_2d(s_stage + r_off * 128, &box_map, 0,
                            (int)((box_row0 + (int64_t)s * tb + (int64_t)c * GEO_CHUNK) >> 3) + r_off, &full[st]);
            tma_load_2d(o_stage0 + st * GEO_STAGE_BYTES + r_off * 128, &box_map, 0,
                        (int)((box_row0 + (int64_t)o * tb + (int64_t)c * GEO_CHUNK) >> 3) + r_off, &full[st]);
        }
    };
    if (tid == 0) {
        for (int q = 0; q < RING && q < nobj; ++q) issue(q);
    }

    float* g = WRITE_GEO ? geo + row[TSPN_VT_GEO_OFF] + ((int64_t)(s * (n - 1) + k0) * TSPN_GEO_CHANNELS) * tp +
                               (int64_t)c * GEO_CHUNK
                         : nullptr;
    int st = (int)(g0 % RING);                               // ring stage of step q and its phase parity
    int ph = (int)((g0 / RING) & 1u);
    for (int q = 0; q < nobj; ++q) {
        const int2 win = owin[q];
        const int a = win.x, b = win.y;                      // overlap window [a, b)
        // Frame ownership rotates with the object: warp w covers the 128-frame block (w + q) mod WARPS.
        // Frames inside the overlap window cost ~10x the frames outside it, and the window sits mostly in
        // the middle of the video, so a fixed block per warp would make the same warps the slow ones for
        // every object of the item while the others idle at the ring (the stage advances with the
        // slowest warp).
        const int j0 = (((warp + (GEO_ROTATE ? q : 0)) & (GEO_WARPS - 1)) * 32 + lane) * GEO_FPT;   // inside the chunk
        const int t0 = c * GEO_CHUNK + j0;                   // first frame of this thread

        mbar_wait(&full[st], ph);

        float fsum_i, fsum_s, fsum_o;
        float out[TSPN_GEO_CHANNELS][GEO_FPT];
        const uint32_t os_addr = smem_u32(o_stage0 + st * GEO_STAGE_BYTES);
        geo_step<CLIP>([ss](int j) { return ld_box(ss, j); }, [os_addr](int j) { return ld_box(os_addr, j); }, j0, t0,
                       a, b, out, fsum_i, fsum_s, fsum_o);
        // this warp is done reading the object stage of step q: hand it back; thread 0 refills it with
        // object q+RING (end of the step) once every warp has done so
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);

        if (WRITE_GEO) {
            if (t0 < tp) {
#pragma unroll
                for (int ch = 0; ch < TSPN_GEO_CHANNELS; ++ch)
                    st_stream_f4(g + (int64_t)ch * tp + j0, make_float4(out[ch][0], out[ch][1], out[ch][2], out[ch][3]));
            }
            g += (int64_t)TSPN_GEO_CHANNELS * tp;
        }
        // the three volume sums over the chunk's frames: integer adds, so neither the warp reduction nor
        // the order of the shared-memory atomics below can change the result
        const unsigned long long tot_i = warp_sum_fx(fsum_i);
        unsigned long long tot_s = 0ull, tot_o = 0ull;
        if (CLIP) {
            tot_s = warp_sum_fx(fsum_s);
            tot_o = warp_sum_fx(fsum_o);
        }
        if (lane == 0) {
            if (tot_i) atomicAdd(acc + q * 3, tot_i);
            if (CLIP) {
                if (tot_s) atomicAdd(acc + q * 3 + 1, tot_s);
                if (tot_o) atomicAdd(acc + q * 3 + 2, tot_o);
            }
        }
        if (tid == 0 && q + RING < nobj) issue(q + RING);    // by now the other warps have normally arrived
        if (++st == RING) { st = 0; ph ^= 1; }
    }
    g0 += (unsigned int)nobj;
    __syncthreads();
    // this chunk's contribution to the pair's sums goes into the pair's own slot for this chunk: one writer per
    // (pair, chunk), plain stores - nothing is zeroed beforehand and no global atomics; the finalize kernel adds
    // the slots of a pair in ascending chunk order
    for (int i3 = tid; i3 < nobj * 3; i3 += THREADS) {
        const int q = i3 / 3, j = i3 - q * 3;
        fx[((pair0 + q) * max_chunks + c) * 3 + j] = acc[i3];
    }
    // the pairs' temporal overlap windows ([SPEC] s3): written here so that feature assembly and the records
    // do not wait for the per-pair finalize
    if (c == 0 && tid < nobj) {
        const int2 w = owin[tid];
        const bool has = w.y > w.x;
        *reinterpret_cast<int2*>(overlap + 2 * (pair0 + tid)) = make_int2(has ? w.x : 0, has ? w.y : 0);
    }
    if (!queue) break;
    }
}

template <int THREADS, bool WRITE_GEO, bool CLIP, bool DENSE>
__global__ void __launch_bounds__(THREADS, GeoCfg<THREADS, DENSE>::MIN_CTAS)
pair_geo_kernel(const __grid_constant__ CUtensorMap box_map, const int64_t* __restrict__ table, int nv,
                const int32_t* __restrict__ span, float* __restrict__ geo, unsigned long long* __restrict__ fx,
                int32_t* __restrict__ overlap, unsigned int* __restrict__ queue, int max_chunks) {
    pair_geo_body<THREADS, WRITE_GEO, CLIP, DENSE>(box_map, table, nv, span, geo, fx, overlap, queue, max_chunks);
}

// The 512-thread shape with the register count PINNED at 104: 16 warps x 104 registers leave exactly the 12 288
// registers one 128-thread x 96-register CTA of the side branch (survivor_rows) needs to co-reside on the SM
// (DESIGN.md section 4).  Under __launch_bounds__(512, 1) ptxas is free to take up to 128 and lands on 104 or 105
// depending on unrelated edits - and 105 is allocated as 112, which locks the side branch out of 140 of the 148
// SMs (measured: survivor_rows 0.42 -> 0.58 ms, step 0.74 -> 0.97 ms).
#ifndef TSPN_GEO_MAXNREG
#define TSPN_GEO_MAXNREG 88
#endif
template <bool WRITE_GEO, bool CLIP>
__global__ void __maxnreg__(TSPN_GEO_MAXNREG)
pair_geo_kernel_r104(const __grid_constant__ CUtensorMap box_map, const int64_t* __restrict__ table, int nv,
                     const int32_t* __restrict__ span, float* __restrict__ geo, unsigned long long* __restrict__ fx,
                     int32_t* __restrict__ overlap, unsigned int* __restrict__ queue, int max_chunks) {
    pair_geo_body<512, WRITE_GEO, CLIP, false>(box_map, table, nv, span, geo, fx, overlap, queue, max_chunks);
}

// ---- post-kernel: per-pair reductions (vIoU, tIoU; the pair kernel writes the overlap windows) ----------
template <bool CLIP>
__global__ void __launch_bounds__(256)
pair_finalize_kernel(const int64_t* __restrict__ table, int nv, const int32_t* __restrict__ span,
                     const double* __restrict__ vol, const unsigned long long* __restrict__ fx, int geo_chunk,
                     int max_chunks, float* __restrict__ viou, float* __restrict__ tiou) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= table_total(table, nv, TSPN_VT_PAIR_OFF)) return;
    const int v = find_video(table, nv, TSPN_VT_PAIR_OFF, p);
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const int n1 = (int)row[TSPN_VT_N] - 1;
    const int loc = (int)(p - row[TSPN_VT_PAIR_OFF]);
    const int s = loc / n1, k = loc - s * n1;
    const int o = k + (k >= s ? 1 : 0);
    const int64_t ts = row[TSPN_VT_TRK_OFF] + s, to = row[TSPN_VT_TRK_OFF] + o;
    const int ps = __ldg(span + 2 * ts), pe = __ldg(span + 2 * ts + 1);
    const int qs = __ldg(span + 2 * to), qe = __ldg(span + 2 * to + 1);
    const int a = max(ps, qs), b = min(pe, qe);
    const bool has = b > a;
    const double inv_scale = 1.0 / (double)GEO_FX_SCALE;
    // the pair's chunk slots, ascending: integer adds, exact
    const int nchunks = ((int)row[TSPN_VT_T] + geo_chunk - 1) / geo_chunk;
    unsigned long long f0 = 0ull, f1 = 0ull, f2 = 0ull;
    for (int c = 0; c < nchunks; ++c) {
        const unsigned long long* slot = fx + (p * max_chunks + c) * 3;
        f0 += slot[0];
        if (CLIP) {
            f1 += slot[1];
            f2 += slot[2];
        }
    }
    const double inter = (double)f0 * inv_scale;
    const double vs = CLIP ? (double)f1 * inv_scale : vol[ts];
    const double vo = CLIP ? (double)f2 * inv_scale : vol[to];
    const double den = vs + vo - inter;
    const int ov = has ? b - a : 0;
    const int tden = (pe - ps) + (qe - qs) - ov;
    viou[p] = (has && den > 0.0) ? (float)(inter / den) : 0.0f;
    tiou[p] = (has && tden > 0) ? (float)((double)ov / (double)tden) : 0.0f;
}

// ---- cubic_iou(bboxes1, bboxes2): one warp per matrix entry -------------------------------------
__global__ void __launch_bounds__(128) cubic_iou_kernel(const float4* __restrict__ b1, int n1,
                                                        const float4* __restrict__ b2, int n2, int t,
                                                        float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e >= (int64_t)n1 * n2) return;
    const int lane = threadIdx.x & 31;
    const int i = (int)(e / n2), j = (int)(e - (int64_t)i * n2);
    const float4* p = b1 + (int64_t)i * t;
    const float4* q = b2 + (int64_t)j * t;
    double si = 0.0, sa = 0.0, sb = 0.0;
    for (int f = lane; f < t; f += 32) {
        const float4 x = __ldg(p + f), y = __ldg(q + f);
        const float iw = fmaxf((fminf(x.z, y.z) - fmaxf(x.x, y.x)) + 1.0f, 0.0f);
        const float ih = fmaxf((fminf(x.w, y.w) - fmaxf(x.y, y.y)) + 1.0f, 0.0f);
        si += (double)(iw * ih);
        sa += (double)(((x.z - x.x) + 1.0f) * ((x.w - x.y) + 1.0f));
        sb += (double)(((y.z - y.x) + 1.0f) * ((y.w - y.y) + 1.0f));
    }
    si = warp_sum(si);
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    if (lane == 0) out[e] = (float)(si / (sa + sb - si));
}

// ---- viou over an explicit pair list (evaluation / association use) --------------------------------
template <bool CLIP>
__global__ void __launch_bounds__(128) viou_pairs_kernel(const float4* __restrict__ pool,
                                                         const int64_t* __restrict__ traj_off,
                                                         const int32_t* __restrict__ traj_span,
                                                         const int32_t* __restrict__ ia,
                                                         const int32_t* __restrict__ ib, int64_t n_pairs,
                                                         float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e >= n_pairs) return;
    const int lane = threadIdx.x & 31;
    const int i = ia[e], j = ib[e];
    const int s1 = traj_span[2 * i], e1 = traj_span[2 * i + 1];
    const int s2 = traj_span[2 * j], e2 = traj_span[2 * j + 1];
    const float4* p = pool + traj_off[i];
    const float4* q = pool + traj_off[j];
    const int a = max(s1, s2), b = min(e1, e2);
    double si = 0.0, sa = 0.0, sb = 0.0;
    for (int f = a + lane; f < b; f += 32) {
        const float4 x = __ldg(p + (f - s1)), y = __ldg(q + (f - s2));
        const float iw = fmaxf((fminf(x.z, y.z) - fmaxf(x.x, y.x)) + 1.0f, 0.0f);
        const float ih = fmaxf((fminf(x.w, y.w) - fmaxf(x.y, y.y)) + 1.0f, 0.0f);
        si += (double)(iw * ih);
        if (CLIP) {
            sa += (double)(((x.z - x.x) + 1.0f) * ((x.w - x.y) + 1.0f));
            sb += (double)(((y.z - y.x) + 1.0f) * ((y.w - y.y) + 1.0f));
        }
    }
    if (!CLIP) {
        for (int f = lane; f < e1 - s1; f += 32) {
            const float4 x = __ldg(p + f);
            sa += (double)(((x.z - x.x) + 1.0f) * ((x.w - x.y) + 1.0f));
        }
        for (int f = lane; f < e2 - s2; f += 32) {
            const float4 y = __ldg(q + f);
            sb += (double)(((y.z - y.x) + 1.0f) * ((y.w - y.y) + 1.0f));
        }
    }
    si = warp_sum(si);
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    if (lane == 0) {
        const double den = sa + sb - si;
        out[e] = (b > a && den > 0.0) ? (float)(si / den) : 0.0f;
    }
}

// ---- viou over an explicit pair list, evaluation form (N3): per-trajectory volumes once, then one
// warp per pair over the overlap window only; sums and the ratio in fp64 so that the host-side
// threshold / arg-max decisions of eval_detection_scores (visual_relation_detection.py:8-36) see
// exactly python's float(v_overlap) / (v1 + v2 - v_overlap) for integer boxes ---------------------
__global__ void __launch_bounds__(128) traj_volume_kernel(const float4* __restrict__ pool,
                                                          const int64_t* __restrict__ traj_off,
                                                          const int32_t* __restrict__ traj_span,
                                                          const int32_t* __restrict__ traj_len, int64_t n_traj,
                                                          double* __restrict__ vol) {
    const int64_t j = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (j >= n_traj) return;
    const int lane = threadIdx.x & 31;
    const float4* p = pool + traj_off[j];
    // common.py:100-105 sums the volume over the whole box LIST, which may be longer than the duration
    const int len = traj_len ? traj_len[j] : traj_span[2 * j + 1] - traj_span[2 * j];
    double acc = 0.0;
    for (int f = lane; f < len; f += 32) {
        const float4 x = __ldg(p + f);
        acc += (double)((x.z - x.x) + 1.0f) * (double)((x.w - x.y) + 1.0f);
    }
    acc = warp_sum(acc);
    if (lane == 0) vol[j] = acc;
}

template <bool CLIP>
__global__ void __launch_bounds__(128) viou_pairs_f64_kernel(const float4* __restrict__ pool,
                                                             const int64_t* __restrict__ traj_off,
                                                             const int32_t* __restrict__ traj_span,
                                                             const double* __restrict__ vol,
                                                             const int32_t* __restrict__ ia,
                                                             const int32_t* __restrict__ ib, int64_t n_pairs,
                                                             double* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e >= n_pairs) return;
    const int lane = threadIdx.x & 31;
    const int i = ia[e], j = ib[e];
    const int s1 = traj_span[2 * i], e1 = traj_span[2 * i + 1];
    const int s2 = traj_span[2 * j], e2 = traj_span[2 * j + 1];
    const float4* p = pool + traj_off[i] + (max(s1, s2) - s1);
    const float4* q = pool + traj_off[j] + (max(s1, s2) - s2);
    const int len = min(e1, e2) - max(s1, s2);
    double si = 0.0, sa = 0.0, sb = 0.0;
    for (int f = lane; f < len; f += 32) {
        const float4 x = __ldg(p + f), y = __ldg(q + f);
        const float iw = fmaxf((fminf(x.z, y.z) - fmaxf(x.x, y.x)) + 1.0f, 0.0f);
        const float ih = fmaxf((fminf(x.w, y.w) - fmaxf(x.y, y.y)) + 1.0f, 0.0f);
        si += (double)iw * (double)ih;
        if (CLIP) {
            sa += (double)((x.z - x.x) + 1.0f) * (double)((x.w - x.y) + 1.0f);
            sb += (double)((y.z - y.x) + 1.0f) * (double)((y.w - y.y) + 1.0f);
        }
    }
    si = warp_sum(si);
    if (CLIP) {
        sa = warp_sum(sa);
        sb = warp_sum(sb);
    } else {
        sa = vol[i];
        sb = vol[j];
    }
    if (lane == 0) {
        const double den = sa + sb - si;
        out[e] = (len > 0 && den > 0.0) ? si / den : 0.0;
    }
}

// ---- pair enumeration ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) enumerate_pairs_kernel(const int64_t* __restrict__ table, int nv,
                                                              int64_t total_pairs, int64_t* __restrict__ pairs) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total_pairs || p >= table_total(table, nv, TSPN_VT_PAIR_OFF)) return;
    const int v = find_video(table, nv, TSPN_VT_PAIR_OFF, p);
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const int64_t n1 = row[TSPN_VT_N] - 1;
    const int64_t loc = p - row[TSPN_VT_PAIR_OFF];
    const int64_t s = loc / n1, k = loc - s * n1;
    pairs[2 * p] = s;
    pairs[2 * p + 1] = k + (k >= s ? 1 : 0);
}

// ---- offsets of the windowed rows: geo_off[p] = 7 * sum of Lw over the pairs before p (floats, a multiple of 4),
// Lw = the pair's overlap window rounded out to multiples of 4 frames.  One CTA walks the pairs in tiles of
// WO_THREADS * 4 with a block scan per tile: a few tens of microseconds per batch, run once per upload (beside the
// box expansion), not in the step.
constexpr int WO_THREADS = 1024;
__global__ void __launch_bounds__(WO_THREADS)
geo_window_offsets_kernel(const int64_t* __restrict__ table, int nv, int64_t total_pairs_cap,
                          const int32_t* __restrict__ span, int64_t* __restrict__ geo_off, int64_t* __restrict__ total) {
    __shared__ int64_t warp_sum[WO_THREADS / 32];
    __shared__ int64_t s_running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_pairs = min(total_pairs_cap, table_total(table, nv, TSPN_VT_PAIR_OFF));
    if (tid == 0) s_running = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_pairs; base += WO_THREADS * 4) {
        int64_t len[4];
        int64_t mine = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t p = base + (int64_t)tid * 4 + e;
            len[e] = 0;
            if (p < n_pairs) {
                const int v = find_video(table, nv, TSPN_VT_PAIR_OFF, p);
                const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
                const int n1 = (int)row[TSPN_VT_N] - 1;
                const int loc = (int)(p - row[TSPN_VT_PAIR_OFF]);
                const int s = loc / n1, k = loc - s * n1;
                const int o = k + (k >= s ? 1 : 0);
                const int64_t ts = row[TSPN_VT_TRK_OFF] + s, to = row[TSPN_VT_TRK_OFF] + o;
                const int a = max(__ldg(span + 2 * ts), __ldg(span + 2 * to));
                const int b = min(__ldg(span + 2 * ts + 1), __ldg(span + 2 * to + 1));
                if (b > a) len[e] = (int64_t)(TSPN_GEO_CHANNELS - 1) * (((b + 3) & ~3) - (a & ~3));
            }
            mine += len[e];
        }
        int64_t incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t up = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += up;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        int64_t before = s_running;
        for (int w = 0; w < warp; ++w) before += warp_sum[w];
        int64_t run = before + incl - mine;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t p = base + (int64_t)tid * 4 + e;
            if (p < n_pairs) geo_off[p] = run;
            run += len[e];
        }
        __syncthreads();
        if (tid == WO_THREADS - 1) s_running = run;
        __syncthreads();
    }
    if (tid == 0) *total = s_running;
}

__global__ void reset_queue_kernel(unsigned int* __restrict__ queue) {
    if (threadIdx.x == 0) *queue = 0u;
}

template <int THREADS, bool DENSE>
static int launch_pair_geo(const int64_t* d_table, int num_videos, int64_t total_items, int64_t total_boxes,
                           const float* d_boxes, const int32_t* d_span, float* d_geo, unsigned long long* fx,
                           int32_t* d_overlap, unsigned int* d_queue, int reserve, bool clip, int max_chunks,
                           cudaStream_t st) {
    using Cfg = GeoCfg<THREADS, DENSE>;
    // boxes viewed as a 2-D tensor: rows of 8 boxes (32 floats = 128 B)
    CUtensorMap map;
    const uint64_t dims[2] = {32, (uint64_t)(total_boxes / 8)};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {32, (uint32_t)Cfg::BOX_ROWS};
    int rc = encode_tensor_map(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_boxes, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != TSPN_OK) return rc;
    unsigned grid = (unsigned)total_items;
    // Shared memory: one CTA per item pins its occupancy with the request (SMEM_BYTES); persistent CTAs are
    // counted out by the grid instead and ask only for what they use, so that on every SM the rest (95 KB beside
    // two 256-thread CTAs, 127 KB beside three 128-thread CTAs) is there for the side branches - with the pin the
    // side branch of a 512- / 1024-frame batch ran on the reserved SMs only (embedding kernel 36 -> 307 us).
    const int smem_bytes = d_queue ? Cfg::SMEM_USED : Cfg::SMEM_BYTES;
    if (d_queue) {                      // persistent: one CTA per SM slot, items from the queue
        int64_t slots = (int64_t)num_sms() * Cfg::MIN_CTAS - reserve;      // SM slots left to concurrent streams
        if (slots < 1) slots = 1;
        if (slots < total_items) grid = (unsigned)slots;
        reset_queue_kernel<<<1, 32, 0, st>>>(d_queue);      // (a kernel, not a memset node: compute-sanitizer's
                                                            // initcheck does not see memset nodes of a replayed graph)
    }
#define TSPN_LAUNCH_GEO(W, C)                                                                                  \
    do {                                                                                                       \
        if (THREADS == 512 && !DENSE) {                                                                        \
            TSPN_CUDA_OK(cudaFuncSetAttribute(pair_geo_kernel_r104<W, C>,                                      \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));  \
            prefer_max_smem(pair_geo_kernel_r104<W, C>);                                                       \
            pair_geo_kernel_r104<W, C><<<grid, THREADS, smem_bytes, st>>>(                                     \
                map, d_table, num_videos, d_span, d_geo, fx, d_overlap, d_queue, max_chunks);                  \
        } else {                                                                                               \
            TSPN_CUDA_OK(cudaFuncSetAttribute(pair_geo_kernel<THREADS, W, C, DENSE>,                           \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));  \
            prefer_max_smem(pair_geo_kernel<THREADS, W, C, DENSE>);                                            \
            pair_geo_kernel<THREADS, W, C, DENSE><<<grid, THREADS, smem_bytes, st>>>(                          \
                map, d_table, num_videos, d_span, d_geo, fx, d_overlap, d_queue, max_chunks);                  \
        }                                                                                                      \
    } while (0)
    if (d_geo) {
        if (clip) TSPN_LAUNCH_GEO(true, true); else TSPN_LAUNCH_GEO(true, false);
    } else {
        if (clip) TSPN_LAUNCH_GEO(false, true); else TSPN_LAUNCH_GEO(false, false);
    }
#undef TSPN_LAUNCH_GEO
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int tspn_enumerate_pairs(const int64_t* d_table, int num_videos, int64_t total_pairs, int64_t* d_pairs,
                         void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(total_pairs >= 0 && num_videos >= 0, TSPN_EBADARG, "tspn_enumerate_pairs: negative size");
    if (total_pairs == 0) return TSPN_OK;
    TSPN_REQUIRE(d_table && d_pairs, TSPN_EBADARG, "tspn_enumerate_pairs: null pointer");
    const int64_t blocks = (total_pairs + 255) / 256;
    enumerate_pairs_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_table, num_videos, total_pairs,
                                                                               d_pairs);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

static inline int64_t geo_ws_vol_bytes(int64_t total_tracklets) {
    return ((total_tracklets > 0 ? total_tracklets : 1) * (int64_t)sizeof(double) + 15) / 16 * 16;
}

int64_t tspn_pair_geo_workspace_bytes(int64_t total_tracklets, int64_t total_pairs, int max_chunks) {
    // per-tracklet volumes | per-(pair, chunk) fixed-point sums | the persistent kernel's work-item queue
    const int64_t mc = max_chunks > 0 ? max_chunks : 1;
    return geo_ws_vol_bytes(total_tracklets) + (total_pairs > 0 ? total_pairs : 1) * mc * 3 * (int64_t)sizeof(uint64_t) + 16;
}

static int pair_geo_viou_impl(const int64_t* d_table, int num_videos, int64_t total_items, int geo_chunk, int max_chunks,
                              int64_t total_tracklets, int64_t total_pairs, int64_t total_boxes, const float* d_boxes,
                              const int32_t* d_span, float* d_geo, const int64_t* d_geo_off, float* d_viou, float* d_tiou,
                              int32_t* d_overlap, int flags, void* d_workspace, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && total_items >= 0 && total_tracklets >= 0 && total_boxes >= 0 && total_pairs >= 0,
                 TSPN_EBADARG, "tspn_pair_geo_viou: negative size");
    if (total_items == 0 || total_pairs == 0) return TSPN_OK;
    TSPN_REQUIRE(d_table && d_boxes && d_span && d_viou && d_tiou && d_overlap && d_workspace, TSPN_EBADARG,
                 "tspn_pair_geo_viou: null pointer");
    TSPN_REQUIRE(aligned16(d_boxes) && aligned16(d_geo) && aligned16(d_workspace), TSPN_EALIGN,
                 "tspn_pair_geo_viou: boxes/geo/workspace must be 16-byte aligned");
    TSPN_REQUIRE((total_boxes & 7) == 0, TSPN_ESHAPE,
                 "tspn_pair_geo_viou: total_boxes (%lld) must be a multiple of 8 (rows padded to Tb)",
                 (long long)total_boxes);
    TSPN_REQUIRE(total_items < (1ll << 31), TSPN_ESHAPE, "tspn_pair_geo_viou: too many work items");
    TSPN_REQUIRE(geo_chunk == 512 || geo_chunk == 1024 || geo_chunk == 2048, TSPN_EBADARG,
                 "tspn_pair_geo_viou: geo_chunk=%d (pass totals[TSPN_TOT_GEO_CHUNK] of tspn_build_video_table)", geo_chunk);
    TSPN_REQUIRE(max_chunks >= 1, TSPN_EBADARG,
                 "tspn_pair_geo_viou: max_chunks=%d (pass totals[TSPN_TOT_MAX_CHUNKS] of tspn_build_video_table)", max_chunks);
    cudaStream_t st = (cudaStream_t)stream;
    double* vol = reinterpret_cast<double*>(d_workspace);
    unsigned long long* fx =
        reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(d_workspace) + geo_ws_vol_bytes(total_tracklets));
    const bool clip = (flags & TSPN_VIOU_CLIPPED) != 0;
    const int phases = flags & (TSPN_GEO_PHASE_PRE | TSPN_GEO_PHASE_MAIN | TSPN_GEO_PHASE_POST);
    const bool all = phases == 0;
    if ((all || (phases & TSPN_GEO_PHASE_PRE)) && !clip && total_tracklets > 0) {
        // per-tracklet volumes over the full spans (the clipped variant sums them per pair in MAIN)
        const int64_t cap = 16 * (int64_t)num_sms();
        const int64_t blocks = total_tracklets < cap ? total_tracklets : cap;          // grid-stride
        prefer_max_smem(tracklet_volume_kernel);
        tracklet_volume_kernel<<<(unsigned)blocks, 128, 0, st>>>(d_table, num_videos,
                                                                reinterpret_cast<const float4*>(d_boxes), d_span, vol);
        TSPN_CUDA_OK(cudaGetLastError());
    }
    if (all || (phases & TSPN_GEO_PHASE_MAIN)) {
        int rc = TSPN_OK;
        const bool dense = (flags & TSPN_GEO_DENSE_CTAS) != 0;
        unsigned int* queue = (flags & TSPN_GEO_PERSISTENT)
                                  ? reinterpret_cast<unsigned int*>(fx + total_pairs * (int64_t)max_chunks * 3)
                                  : nullptr;
        const int reserve = (flags >> TSPN_GEO_RESERVE_SHIFT) & 0xff;
        if (d_geo_off) {
            // the windowed layout's own kernel: a warp per pair over the pair's window (geo_windowed.cu)
            rc = launch_pair_geo_windowed(d_table, num_videos, total_pairs, d_boxes, d_span, d_geo, d_geo_off, fx,
                                          d_overlap, reinterpret_cast<unsigned int*>(fx + total_pairs * (int64_t)max_chunks * 3),
                                          geo_chunk, max_chunks, reserve, clip, st);
            if (rc != TSPN_OK) return rc;
        } else {
#define TSPN_GEO_SHAPE(T)                                                                                          \
    (dense ? launch_pair_geo<T, true>(d_table, num_videos, total_items, total_boxes, d_boxes, d_span, d_geo, fx,   \
                                      d_overlap, queue, reserve, clip, max_chunks, st)                             \
           : launch_pair_geo<T, false>(d_table, num_videos, total_items, total_boxes, d_boxes, d_span, d_geo, fx,  \
                                       d_overlap, queue, reserve, clip, max_chunks, st))
        if (geo_chunk == 512) rc = TSPN_GEO_SHAPE(128);
        else if (geo_chunk == 1024) rc = TSPN_GEO_SHAPE(256);
        else rc = TSPN_GEO_SHAPE(512);
#undef TSPN_GEO_SHAPE
        if (rc != TSPN_OK) return rc;
        }
    }
    if (all || (phases & TSPN_GEO_PHASE_POST)) {
        const unsigned fblocks = (unsigned)((total_pairs + 255) / 256);
        prefer_max_smem(pair_finalize_kernel<true>);
        prefer_max_smem(pair_finalize_kernel<false>);
        if (clip)
            pair_finalize_kernel<true><<<fblocks, 256, 0, st>>>(d_table, num_videos, d_span, vol, fx, geo_chunk,
                                                                max_chunks, d_viou, d_tiou);
        else
            pair_finalize_kernel<false><<<fblocks, 256, 0, st>>>(d_table, num_videos, d_span, vol, fx, geo_chunk,
                                                                 max_chunks, d_viou, d_tiou);
        TSPN_CUDA_OK(cudaGetLastError());
    }
    return TSPN_OK;
}

int tspn_pair_geo_viou(const int64_t* d_table, int num_videos, int64_t total_items, int geo_chunk, int max_chunks,
                       int64_t total_tracklets, int64_t total_pairs, int64_t total_boxes, const float* d_boxes, const int32_t* d_span,
                       float* d_geo, float* d_viou, float* d_tiou, int32_t* d_overlap, int flags, void* d_workspace,
                       void* stream) {
    return pair_geo_viou_impl(d_table, num_videos, total_items, geo_chunk, max_chunks, total_tracklets, total_pairs,
                              total_boxes, d_boxes, d_span, d_geo, nullptr, d_viou, d_tiou, d_overlap, flags, d_workspace,
                              stream);
}

int tspn_pair_geo_viou_windowed(const int64_t* d_table, int num_videos, int64_t total_items, int geo_chunk, int max_chunks,
                                int64_t total_tracklets, int64_t total_pairs, int64_t total_boxes, const float* d_boxes,
                                const int32_t* d_span, float* d_geo, const int64_t* d_geo_off, float* d_viou,
                                float* d_tiou, int32_t* d_overlap, int flags, void* d_workspace, void* stream) {
    TSPN_REQUIRE(d_geo && d_geo_off, TSPN_EBADARG, "tspn_pair_geo_viou_windowed: geo and geo_off are required");
    TSPN_REQUIRE(!(flags & TSPN_GEO_DENSE_CTAS), TSPN_EBADARG, "tspn_pair_geo_viou_windowed: no dense-CTA shape");
    return pair_geo_viou_impl(d_table, num_videos, total_items, geo_chunk, max_chunks, total_tracklets, total_pairs,
                              total_boxes, d_boxes, d_span, d_geo, d_geo_off, d_viou, d_tiou, d_overlap, flags,
                              d_workspace, stream);
}

int tspn_geo_window_offsets(const int64_t* d_table, int num_videos, int64_t total_pairs, const int32_t* d_span,
                            int64_t* d_geo_off, int64_t* d_total, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && total_pairs >= 0, TSPN_EBADARG, "tspn_geo_window_offsets: negative size");
    TSPN_REQUIRE(d_table && d_span && d_geo_off && d_total, TSPN_EBADARG, "tspn_geo_window_offsets: null pointer");
    geo_window_offsets_kernel<<<1, WO_THREADS, 0, (cudaStream_t)stream>>>(d_table, num_videos, total_pairs, d_span,
                                                                          d_geo_off, d_total);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_cubic_iou(const float* d_b1, int n1, const float* d_b2, int n2, int t, float* d_out, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(n1 >= 0 && n2 >= 0 && t >= 0, TSPN_EBADARG, "tspn_cubic_iou: negative size");
    if ((int64_t)n1 * n2 == 0) return TSPN_OK;
    TSPN_REQUIRE(d_b1 && d_b2 && d_out, TSPN_EBADARG, "tspn_cubic_iou: null pointer");
    TSPN_REQUIRE(aligned16(d_b1) && aligned16(d_b2), TSPN_EALIGN, "tspn_cubic_iou: boxes must be 16-byte aligned");
    const int64_t entries = (int64_t)n1 * n2;
    cubic_iou_kernel<<<(unsigned)((entries + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(d_b1), n1, reinterpret_cast<const float4*>(d_b2), n2, t, d_out);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_viou_pairs(const float* d_pool, const int64_t* d_traj_off, const int32_t* d_traj_span, const int32_t* d_a,
                    const int32_t* d_b, int64_t n_pairs, int flags, float* d_out, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(n_pairs >= 0, TSPN_EBADARG, "tspn_viou_pairs: negative size");
    if (n_pairs == 0) return TSPN_OK;
    TSPN_REQUIRE(d_pool && d_traj_off && d_traj_span && d_a && d_b && d_out, TSPN_EBADARG,
                 "tspn_viou_pairs: null pointer");
    TSPN_REQUIRE(aligned16(d_pool), TSPN_EALIGN, "tspn_viou_pairs: pool must be 16-byte aligned");
    const unsigned blocks = (unsigned)((n_pairs + 3) / 4);
    if (flags & TSPN_VIOU_CLIPPED)
        viou_pairs_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(d_pool), d_traj_off, d_traj_span, d_a, d_b, n_pairs, d_out);
    else
        viou_pairs_kernel<false><<<blocks, 128, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(d_pool), d_traj_off, d_traj_span, d_a, d_b, n_pairs, d_out);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int64_t tspn_viou_pairs_workspace_bytes(int64_t n_traj) {
    return ((n_traj > 0 ? n_traj : 1) * (int64_t)sizeof(double) + 15) / 16 * 16;
}

int tspn_viou_pairs_f64(const float* d_pool, const int64_t* d_traj_off, const int32_t* d_traj_span,
                        const int32_t* d_traj_len, int64_t n_traj, const int32_t* d_a, const int32_t* d_b,
                        int64_t n_pairs, int flags, double* d_out, void* d_workspace, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(n_pairs >= 0 && n_traj >= 0, TSPN_EBADARG, "tspn_viou_pairs_f64: negative size");
    if (n_pairs == 0) return TSPN_OK;
    TSPN_REQUIRE(d_pool && d_traj_off && d_traj_span && d_a && d_b && d_out && d_workspace, TSPN_EBADARG,
                 "tspn_viou_pairs_f64: null pointer");
    TSPN_REQUIRE(aligned16(d_pool) && aligned16(d_workspace), TSPN_EALIGN,
                 "tspn_viou_pairs_f64: pool/workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    double* vol = reinterpret_cast<double*>(d_workspace);
    const float4* pool = reinterpret_cast<const float4*>(d_pool);
    const unsigned blocks = (unsigned)((n_pairs + 3) / 4);
    if (flags & TSPN_VIOU_CLIPPED) {
        viou_pairs_f64_kernel<true><<<blocks, 128, 0, st>>>(pool, d_traj_off, d_traj_span, vol, d_a, d_b, n_pairs, d_out);
    } else {
        traj_volume_kernel<<<(unsigned)((n_traj + 3) / 4), 128, 0, st>>>(pool, d_traj_off, d_traj_span, d_traj_len, n_traj,
                                                                          vol);
        TSPN_CUDA_OK(cudaGetLastError());
        viou_pairs_f64_kernel<false><<<blocks, 128, 0, st>>>(pool, d_traj_off, d_traj_span, vol, d_a, d_b, n_pairs, d_out);
    }
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // extern "C"
