// span_head_tc.cu — temporal-span head on the tensor cores (TSPN_PREC_TENSOR, Cin >= 64).
//
//   DPNHead.forward   lib/modeling/relpn/dpn.py:55-73
//   out[k, j, t] = b2[j] + sum_co W2[j, co] * relu(b[co] + sum_{ci, d} W[co, ci, d] * x[k, ci, t + d - 1])
//
// At the reference's default width (Cin = 1024, A = 4: lib/config/defaults.py:64-65) this is
// 6.3 MFLOP per pair-frame and the one kernel of the path that is bound by the tensor pipe
// (1.5 kFLOP/B).  It runs as an implicit GEMM on tcgen05:
//
//   * pre-pass (span_pack_x_kernel): the gathered x rows are transposed and converted once into
//     xt[row][ci] bf16 (K-major), row = 1 + k*(T+2) + 1 + t, with an all-zero row in front of and
//     behind every pair.  The Conv1d(k=3, pad=1) then is three GEMMs over the SAME flattened row
//     space, shifted by -1 / 0 / +1 rows: the zero rows are the convolution's padding, and row
//     tiles may straddle pairs, so there is no per-pair tile quantisation (T = 300 would waste
//     22 % of every third 128-row tile otherwise);
//   * work unit = (128-row tile, chunk of <= 256 output channels); persistent CTAs (one per SM)
//     walk the units round-robin.  Per unit the K loop runs over 3 taps x Cin/64 slabs: TMA
//     (SWIZZLE_128B) brings the shifted A tile [128 rows x 64 ci] and the weight tile
//     [chunk co x 64 ci] of that tap through a 4-stage mbarrier ring; one elected thread issues
//     tcgen05.mma (M = 128, N = chunk, K = 16, bf16 -> fp32) into one of two TMEM accumulators;
//   * the epilogue warps read the finished accumulator (tcgen05.ld), add the conv bias, apply
//     the ReLU and fold the hidden units straight into the 2A outputs (the 1x1 conv) while the
//     MMA warp already works on the next unit in the other TMEM buffer — the hidden tensor
//     never exists in memory;
//   * with several channel chunks the per-chunk partial outputs are summed in chunk order by
//     span_finish_kernel (deterministic), which also adds b2 and writes the [k][2A][T] layout.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue.
//
// What bounded the first version of this kernel (profiles/r2_contraction_kernels.md, limiter experiments with the loads,
// the MMAs and the epilogue switched off one at a time): not L2 and not the tensor pipe (51 % busy) but two serial
// chains in one-thread roles - 256 global-load -> use round trips per unit in the epilogue (three `__ldg` per column
// behind a `co < cin` branch; now one staged slice per unit and LDS broadcasts), and two 64-bit divisions by the
// run-time stage count per K block in the producer and MMA threads (~800 cycles against the 512 cycles of MMA a block
// feeds; now incremental stage / phase counters).  [256, 1024, 300]: main kernel 0.494 -> 0.31 ms.
//
// CTA-PAIR form (span_head_tc2_kernel, cta_group::2, Cin >= 256, opt-in: TSPN_SPAN_HEAD_PAIR=1): a pair of CTAs on the
// two SMs of a TPC shares one M = 256 x N = 256 MMA: each CTA stages its own 128 rows of A and only HALF of the weight
// tile (128 of the 256 output channels), the tensor cores of both SMs read both halves - (128 + 128) x 64 per CTA for
// the same math, 2/3 of the L2 -> SM traffic.  The leader CTA issues the MMAs; both CTAs load (signalling the leader's
// mbarrier), both run the epilogue on their own 128 accumulator rows.  Same results bit for bit; measured 3-7 % SLOWER
// than the one-CTA form at both benchmark shapes once the serial chains were gone (the one-CTA main kernel then runs
// at the rate of the measured cuBLAS bf16 burst), so it stays the opt-in.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace tspn {

constexpr int ST_BM = 128;                 // rows (pair-frames) per tile = UMMA M
constexpr int ST_SLAB = 64;                // bf16 elements of K per stage row (128 bytes)
constexpr int ST_THREADS = 192;
constexpr int ST_MAX_CHUNK = 256;          // output channels per unit = UMMA N
constexpr int ST_SMEM_BUDGET = 224 * 1024;    // stages + the epilogue's slice (227 KB per CTA on sm_100)

static inline int64_t st_round(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline int st_chunk(int cin) { return cin >= ST_MAX_CHUNK ? ST_MAX_CHUNK : (int)st_round(cin, 16); }
static inline int st_a2p(int a2) { return (int)st_round(a2, 4); }

struct StLayout {
    int64_t rows;        // k * (t + 2): flattened pair-frames incl. the two zero rows per pair
    int64_t tiles;       // ceil(rows / 128)
    int chunk, chunks, a2p;
    int64_t off_xt, off_wb, off_w2t, off_partial, total;
};

static StLayout st_layout(int64_t k, int cin, int t, int a2) {
    StLayout L;
    L.rows = k * ((int64_t)t + 2);
    L.tiles = (L.rows + 2 * ST_BM - 1) / (2 * ST_BM) * 2;      // even: a CTA pair works on two row tiles at once
    L.chunk = st_chunk(cin);
    L.chunks = (cin + L.chunk - 1) / L.chunk;
    L.a2p = st_a2p(a2);
    int64_t off = 0;
    L.off_xt = off;
    off += st_round((L.rows + 2) * (int64_t)cin * 2, 256);
    L.off_wb = off;
    off += st_round(3ll * cin * cin * 2, 256);
    L.off_w2t = off;
    off += st_round((int64_t)cin * L.a2p * 4, 256);
    L.off_partial = off;
    off += st_round((int64_t)L.chunks * L.tiles * ST_BM * L.a2p * 4, 256);
    L.total = off;
    return L;
}

// ---- pre-pass 1: weights.  wb[d][co][ci] = bf16(W[co][ci][d]);  w2t[co][j] = W2[j][co] (zero padded) ----
__global__ void __launch_bounds__(256)
span_pack_w_kernel(const float* __restrict__ conv_w, const float* __restrict__ pred_w, int cin, int a2, int a2p,
                   __nv_bfloat16* __restrict__ wb, float* __restrict__ w2t) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)cin * cin;
    if (idx < n) {
        const float* w = conv_w + idx * 3;
#pragma unroll
        for (int d = 0; d < 3; ++d) wb[(int64_t)d * n + idx] = __float2bfloat16(__ldg(w + d));
    }
    if (idx < (int64_t)cin * a2p) {
        const int co = (int)(idx / a2p), j = (int)(idx % a2p);
        w2t[idx] = j < a2 ? __ldg(pred_w + (int64_t)j * cin + co) : 0.0f;
    }
}

// ---- pre-pass 2: x[k][ci][t] fp32 (rows gathered) -> xt[1 + k*(T+2) + 1 + t][ci] bf16 --------------------
// One CTA transposes a [64 ci x 32 t] tile through shared memory: 128-byte coalesced reads along t,
// 128-byte coalesced bf16x2 writes along ci.  Tiles with t0 == 0 / the last t tile also write the
// pair's leading / trailing zero row; pair 0 / the last pair write the buffer's outer zero rows.
__global__ void __launch_bounds__(256)
span_pack_x_kernel(const float* __restrict__ x, const int64_t* __restrict__ rows, int64_t row_base,
                   int64_t row_stride, int64_t ld_t, int cin, int t_len, int64_t k_total,
                   __nv_bfloat16* __restrict__ xt) {
    __shared__ float tile[64][33];
    const int64_t k = blockIdx.z;
    const int ci0 = blockIdx.y * 64, t0 = blockIdx.x * 32;
    const int64_t src = rows ? rows[k] - (rows[k] >= 0 ? row_base : 0) : k;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 8 warps
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = ty + 8 * i, ci = ci0 + c, t = t0 + tx;
        float v = 0.0f;
        if (src >= 0 && ci < cin && t < t_len) v = __ldg(x + src * row_stride + (int64_t)ci * ld_t + t);
        tile[c][tx] = v;
    }
    __syncthreads();
    const int64_t row0 = 1 + k * ((int64_t)t_len + 2);           // the pair's leading zero row
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int tt = ty + 8 * i, t = t0 + tt, ci = ci0 + 2 * tx;
        if (t < t_len && ci < cin) {                              // cin is even (multiple of 8)
            const __nv_bfloat162 v = __floats2bfloat162_rn(tile[2 * tx][tt], tile[2 * tx + 1][tt]);
            *reinterpret_cast<__nv_bfloat162*>(xt + (row0 + 1 + t) * cin + ci) = v;
        }
    }
    // zero rows: 64 ci of this CTA's slab, written by the first warp
    if (ty == 0) {
        const int ci = ci0 + 2 * tx;
        if (ci < cin) {
            const __nv_bfloat162 z = __floats2bfloat162_rn(0.0f, 0.0f);
            if (t0 == 0) {
                *reinterpret_cast<__nv_bfloat162*>(xt + row0 * cin + ci) = z;
                if (k == 0) *reinterpret_cast<__nv_bfloat162*>(xt + ci) = z;
            }
            if (t0 + 32 >= t_len) {
                *reinterpret_cast<__nv_bfloat162*>(xt + (row0 + 1 + t_len) * cin + ci) = z;
                if (k == k_total - 1) *reinterpret_cast<__nv_bfloat162*>(xt + (row0 + 2 + t_len) * cin + ci) = z;
            }
        }
    }
}

// ---- epilogue pieces shared by both kernel forms --------------------------------------------------------------------
// The four epilogue warps (threads 64..191) stage the unit's slice of the 1x1 conv - w2t[co0 .. co0 + chunk) and the
// conv bias - in shared memory once per unit, zero padded to `chunk` columns (columns beyond Cin read accumulators that
// are exactly zero: their weight rows are TMA out-of-bounds fill): the column loop then has no guards and its operands
// are warp-uniform LDS broadcasts.  (First version: three global loads per column behind a `co < cin` branch, which
// ptxas left as 256 serialised load -> use round trips per unit: 24 us of epilogue per unit against 13 us of MMA.)
__device__ __forceinline__ void st_named_bar(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
constexpr int ST_EPI_THREADS = 128;
static inline size_t st_epi_bytes(int chunk, int a2p) { return (size_t)chunk * (a2p + 1) * sizeof(float); }

template <int A2P>
__device__ __forceinline__ void st_stage_slice(float* __restrict__ s_w, float* __restrict__ s_b, int chunk, int co0,
                                               int cin, const float* __restrict__ conv_b,
                                               const float* __restrict__ w2t) {
    const int et = (int)threadIdx.x - 64;
    const int ncols = min(chunk, cin - co0);
    st_named_bar(1, ST_EPI_THREADS);                     // every warp is done with the previous slice
    const float4* src = reinterpret_cast<const float4*>(w2t + (int64_t)co0 * A2P);
    float4* dst = reinterpret_cast<float4*>(s_w);
    for (int i = et; i < chunk * (A2P / 4); i += ST_EPI_THREADS)
        dst[i] = i < ncols * (A2P / 4) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = et; i < chunk; i += ST_EPI_THREADS) s_b[i] = (conv_b && i < ncols) ? __ldg(conv_b + co0 + i) : 0.0f;
    st_named_bar(1, ST_EPI_THREADS);
}

// acc[j] = sum over the chunk's columns of W2[j][co] * relu(accumulator[row][co] + b[co]) for this lane's row
template <int A2P>
__device__ __forceinline__ void st_fold_columns(uint32_t taddr, int chunk, const float* __restrict__ s_w,
                                                const float* __restrict__ s_b, float (&acc)[A2P]) {
#pragma unroll
    for (int j = 0; j < A2P; ++j) acc[j] = 0.0f;
    for (int c0 = 0; c0 < chunk; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_b + c0 + 4 * g);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float h = fmaxf(v[4 * g + e] + bb[e], 0.0f);
                const float4* wq = reinterpret_cast<const float4*>(s_w + (c0 + 4 * g + e) * A2P);
#pragma unroll
                for (int q = 0; q < A2P / 4; ++q) {
                    const float4 w = wq[q];
                    acc[4 * q + 0] = fmaf(w.x, h, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(w.y, h, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(w.z, h, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(w.w, h, acc[4 * q + 3]);
                }
            }
        }
    }
}

// ---- the implicit-GEMM kernel ---------------------------------------------------------------------------
template <int A2P>
__global__ void __launch_bounds__(ST_THREADS, 1)
span_head_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                    int64_t tiles, int cin, int chunk, int chunks, int kslabs, int stages, uint32_t tmem_cols,
                    const float* __restrict__ conv_b, const float* __restrict__ w2t,
                    float* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int stage_a = ST_BM * 128;
    const int stage_bytes = stage_a + chunk * 128;
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    uint64_t* const empty = full + stages;
    uint64_t* const acc_full = empty + stages;          // [2]
    uint64_t* const acc_empty = acc_full + 2;           // [2]
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* const s_w = reinterpret_cast<float*>(tmem_slot + 4);          // [chunk][A2P], 16-byte aligned
    float* const s_b = s_w + (size_t)chunk * A2P;                        // [chunk]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t units = tiles * chunks;
    const int nkb = 3 * kslabs;                         // K blocks per unit: tap-major

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);                // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ring position kept incrementally (stage s, phase bit ph): a division by the run-time stage count per K
            // block - two 64-bit ones in the first version - is a ~400-cycle serial chain in a one-thread role, more
            // than the 512 cycles of MMA a block feeds
            int s = 0;
            uint32_t ph = 0;
            bool first_lap = true;
            for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
                const int64_t tile = u / chunks;
                const int co0 = (int)(u - tile * chunks) * chunk;
                for (int tap = 0; tap < 3; ++tap) {
                    for (int ci0 = 0; ci0 < kslabs * ST_SLAB; ci0 += ST_SLAB) {
                        if (!first_lap) mbar_wait(&empty[s], ph ^ 1u);
                        uint8_t* a = smem + (size_t)s * stage_bytes;
                        mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                        // output row r reads xt rows (1 + r) + tap - 1 = r + tap
                        tma_load_2d(a, &map_x, ci0, (int)(tile * ST_BM + tap), &full[s]);
                        tma_load_3d(a + stage_a, &map_w, ci0, co0, tap, &full[s]);
                        if (++s == stages) { s = 0; ph ^= 1u; first_lap = false; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(UMMA_FMT_BF16, ST_BM, (uint32_t)chunk, 0, 0);
            int64_t ul = 0;
            int s = 0;
            uint32_t ph = 0;
            for (int64_t u = blockIdx.x; u < units; u += gridDim.x, ++ul) {
                const int b = (int)(ul & 1);
                const int64_t buse = ul >> 1;
                if (buse > 0) mbar_wait(&acc_empty[b], (uint32_t)((buse - 1) & 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(b * chunk);
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t adesc = umma_smem_desc(a_addr, 16, 1024);
                    const uint64_t bdesc = umma_smem_desc(a_addr + stage_a, 16, 1024);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)       // 4 x (K = 16 bf16 = 32 bytes) per 128-byte slab
                        umma_f16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (i | kk) != 0);
                    umma_commit(&empty[s]);
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(&acc_full[b]);
            }
        }
    } else {
        // ---- epilogue: warp w owns TMEM lanes [32*(w%4), +32) = rows of the tile ----
        const int quad = warp & 3;
        int64_t ul = 0;
        for (int64_t u = blockIdx.x; u < units; u += gridDim.x, ++ul) {
            const int b = (int)(ul & 1);
            const int64_t tile = u / chunks;
            const int cidx = (int)(u - tile * chunks);
            const int co0 = cidx * chunk;
            st_stage_slice<A2P>(s_w, s_b, chunk, co0, cin, conv_b, w2t);     // (under this unit's MMAs)
            mbar_wait(&acc_full[b], (uint32_t)((ul >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * chunk);
            float acc[A2P];
            st_fold_columns<A2P>(taddr, chunk, s_w, s_b, acc);
            // the accumulator has been read: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[b]);
            const int64_t row = tile * ST_BM + quad * 32 + lane;
            float4* dst = reinterpret_cast<float4*>(partial + ((int64_t)cidx * tiles * ST_BM + row) * A2P);
#pragma unroll
            for (int q = 0; q < A2P / 4; ++q)
                dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---- the same, by CTA pairs (cta_group::2, Cin >= 256: chunk = 256) ---------------------------------------------------
// Work unit = (pair of row tiles 2 u', 2 u' + 1, chunk of 256 output channels); cluster c of the persistent grid walks
// units c, c + clusters, ...  CTA r of the pair owns row tile 2 u' + r (A rows, accumulator rows, epilogue) and stages
// output channels [co0 + 128 r, +128) of the weight tile.  Barriers: full[s] lives in the leader (both CTAs' TMA bytes
// are expected there), empty[s] and acc_full[b] are signalled in both CTAs by the leader's multicast commits,
// acc_empty[b] lives in the leader and counts the epilogue warps of both CTAs.  Every wait is bounded (trap, no hang).
constexpr int ST_PAIR_HALF = ST_MAX_CHUNK / 2;       // weight rows per CTA
constexpr int ST_PAIR_STAGE = ST_BM * 128 + ST_PAIR_HALF * 128;

template <int A2P>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ST_THREADS, 1)
span_head_tc2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     int64_t tiles, int cin, int chunks, int kslabs, int stages,
                     const float* __restrict__ conv_b, const float* __restrict__ w2t,
                     float* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int stage_a = ST_BM * 128;
    constexpr int stage_bytes = ST_PAIR_STAGE;
    constexpr int chunk = ST_MAX_CHUNK;
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    uint64_t* const empty = full + stages;
    uint64_t* const acc_full = empty + stages;          // [2]
    uint64_t* const acc_empty = acc_full + 2;           // [2]
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* const s_w = reinterpret_cast<float*>(tmem_slot + 4);
    float* const s_b = s_w + (size_t)chunk * A2P;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();            // 0 = leader
    const int64_t cluster = blockIdx.x >> 1, clusters = gridDim.x >> 1;
    const int64_t units = (tiles >> 1) * chunks;
    const int nkb = 3 * kslabs;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);                     // the leader's producer (expect_tx of both CTAs' bytes)
            mbar_init(&empty[s], 1);                    // the leader's multicast commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 8);                // four epilogue warps of each CTA (used in the leader only)
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_slot, 2 * chunk);
        tmem_relinquish_pair();
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();                                 // both CTAs' barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            bool first_lap = true;
            const uint32_t full0 = mapa_rank(smem_u32(&full[0]), 0);             // the leader's full[] barriers
            for (int64_t u = cluster; u < units; u += clusters) {
                const int64_t tp = u / chunks;
                const int co0 = (int)(u - tp * chunks) * chunk + (int)rank * ST_PAIR_HALF;
                const int64_t tile = 2 * tp + rank;
                for (int tap = 0; tap < 3; ++tap) {
                    for (int ci0 = 0; ci0 < kslabs * ST_SLAB; ci0 += ST_SLAB) {
                        if (!first_lap) mbar_wait_bounded(&empty[s], ph ^ 1u);
                        uint8_t* a = smem + (size_t)s * stage_bytes;
                        if (rank == 0) mbar_expect_tx(&full[s], 2u * (uint32_t)stage_bytes);
                        const uint32_t bar = full0 + 8u * (uint32_t)s;
                        tma_load_2d_pair(a, &map_x, ci0, (int)(tile * ST_BM + tap), bar);
                        tma_load_3d_pair(a + stage_a, &map_w, ci0, co0, tap, bar);
                        if (++s == stages) { s = 0; ph ^= 1u; first_lap = false; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = umma_idesc(UMMA_FMT_BF16, 2 * ST_BM, (uint32_t)chunk, 0, 0);
            int64_t ul = 0;
            int s = 0;
            uint32_t ph = 0;
            for (int64_t u = cluster; u < units; u += clusters, ++ul) {
                const int b = (int)(ul & 1);
                const int64_t buse = ul >> 1;
                if (buse > 0) mbar_wait_bounded(&acc_empty[b], (uint32_t)((buse - 1) & 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(b * chunk);
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait_bounded(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t adesc = umma_smem_desc(a_addr, 16, 1024);
                    const uint64_t bdesc = umma_smem_desc(a_addr + stage_a, 16, 1024);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_f16_pair(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (i | kk) != 0);
                    umma_commit_pair(&empty[s]);        // the stage is free in both CTAs
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
                umma_commit_pair(&acc_full[b]);         // both CTAs' halves of the accumulator are complete
            }
        }
    } else {
        const int quad = warp & 3;
        const uint32_t acc_empty_leader = mapa_rank(smem_u32(&acc_empty[0]), 0);
        int64_t ul = 0;
        for (int64_t u = cluster; u < units; u += clusters, ++ul) {
            const int b = (int)(ul & 1);
            const int64_t tp = u / chunks;
            const int cidx = (int)(u - tp * chunks);
            const int co0 = cidx * chunk;
            const int64_t tile = 2 * tp + rank;
            st_stage_slice<A2P>(s_w, s_b, chunk, co0, cin, conv_b, w2t);
            mbar_wait_bounded(&acc_full[b], (uint32_t)((ul >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * chunk);
            float acc[A2P];
            st_fold_columns<A2P>(taddr, chunk, s_w, s_b, acc);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty_leader + (uint32_t)b * 8u);
            const int64_t row = tile * ST_BM + quad * 32 + lane;
            float4* dst = reinterpret_cast<float4*>(partial + ((int64_t)cidx * tiles * ST_BM + row) * A2P);
#pragma unroll
            for (int q = 0; q < A2P / 4; ++q)
                dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        }
    }
    // neither CTA leaves (or frees its TMEM) while the other may still read its shared memory or signal its barriers
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 2 * chunk);
    }
}

// ---- finish: sum the channel chunks in order, add b2, write out[k][j][t]; padding pairs -> 0 -------------
__global__ void __launch_bounds__(256)
span_finish_kernel(const float* __restrict__ partial, int chunks, int64_t rows_pad, int a2p, int a2, int t_len,
                   const int64_t* __restrict__ rows, const float* __restrict__ pred_b, float* __restrict__ out) {
    const int64_t k = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_len) return;
    const bool pad = rows && rows[k] < 0;
    const int64_t r = k * ((int64_t)t_len + 2) + 1 + t;
    for (int j = 0; j < a2; ++j) {
        float z = 0.0f;
        if (!pad) {
            z = pred_b ? __ldg(pred_b + j) : 0.0f;
            for (int c = 0; c < chunks; ++c) z += __ldg(partial + ((int64_t)c * rows_pad + r) * a2p + j);
        }
        out[(k * a2 + j) * (int64_t)t_len + t] = z;
    }
}

int span_head_tensor(const float* d_x, const int64_t* d_rows, int64_t row_base, int64_t row_stride, int64_t ld_t,
                     int64_t k, int cin, int t, const float* d_conv_w, const float* d_conv_b, const float* d_pred_w,
                     const float* d_pred_b, int a2, float* d_out, void* d_workspace, cudaStream_t st) {
    TSPN_REQUIRE(cin >= 64 && cin % 8 == 0, TSPN_ESHAPE,
                 "tensor span head needs Cin >= 64 and a multiple of 8 (got %d); use TSPN_PREC_FP32_EXACT", cin);
    TSPN_REQUIRE(a2 <= 16, TSPN_ESHAPE, "tensor span head supports 2A <= 16 (got %d)", a2);
    TSPN_REQUIRE(d_workspace && aligned16(d_workspace), TSPN_EBADARG,
                 "tensor span head: 16-byte aligned workspace of tspn_span_head_workspace_bytes() required");
    const StLayout L = st_layout(k, cin, t, a2);
    TSPN_REQUIRE(L.rows + 2 + ST_BM < (1ll << 31), TSPN_ESHAPE, "tensor span head: k*(t+2) too large");
    uint8_t* ws = reinterpret_cast<uint8_t*>(d_workspace);
    __nv_bfloat16* xt = reinterpret_cast<__nv_bfloat16*>(ws + L.off_xt);
    __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(ws + L.off_wb);
    float* w2t = reinterpret_cast<float*>(ws + L.off_w2t);
    float* partial = reinterpret_cast<float*>(ws + L.off_partial);

    {
        const int64_t n = (int64_t)cin * cin;
        span_pack_w_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_conv_w, d_pred_w, cin, a2, L.a2p, wb, w2t);
        TSPN_CUDA_OK(cudaGetLastError());
        dim3 grid((unsigned)((t + 31) / 32), (unsigned)((cin + 63) / 64), (unsigned)k);
        span_pack_x_kernel<<<grid, 256, 0, st>>>(d_x, d_rows, row_base, row_stride, ld_t, cin, t, k, xt);
        TSPN_CUDA_OK(cudaGetLastError());
    }

    const int kslabs = (cin + ST_SLAB - 1) / ST_SLAB;
    const int stage_bytes = ST_BM * 128 + L.chunk * 128;
    int stages = (int)((ST_SMEM_BUDGET - st_epi_bytes(L.chunk, L.a2p)) / stage_bytes);
    if (stages > 6) stages = 6;
    if (stages < 2) stages = 2;
    const size_t smem_bytes = (size_t)stages * stage_bytes + (2 * stages + 4) * sizeof(uint64_t) + 16 +
                              st_epi_bytes(L.chunk, L.a2p);
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * L.chunk) tmem_cols <<= 1;

    CUtensorMap map_x, map_w;
    {
        const uint64_t dims[2] = {(uint64_t)cin, (uint64_t)(L.rows + 2)};
        const uint64_t strides[1] = {(uint64_t)cin * 2};
        const uint32_t box[2] = {(uint32_t)ST_SLAB, (uint32_t)ST_BM};
        int rc = encode_tensor_map(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xt, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != TSPN_OK) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)cin, (uint64_t)cin, 3};
        const uint64_t strides[2] = {(uint64_t)cin * 2, (uint64_t)cin * cin * 2};
        const uint32_t box[3] = {(uint32_t)ST_SLAB, (uint32_t)L.chunk, 1};
        int rc = encode_tensor_map(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wb, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != TSPN_OK) return rc;
    }
    const int sms = num_sms();
    const char* pair_env = getenv("TSPN_SPAN_HEAD_PAIR");                  // opt-in (read per call: tests run both forms)
    if (L.chunk == ST_MAX_CHUNK && pair_env && pair_env[0] == '1') {
        // CTA pairs: half of the weight tile per CTA
        CUtensorMap map_wh;
        const uint64_t dims[3] = {(uint64_t)cin, (uint64_t)cin, 3};
        const uint64_t strides[2] = {(uint64_t)cin * 2, (uint64_t)cin * cin * 2};
        const uint32_t box[3] = {(uint32_t)ST_SLAB, (uint32_t)ST_PAIR_HALF, 1};
        int rc = encode_tensor_map(&map_wh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wb, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != TSPN_OK) return rc;
        int pstages = (int)((ST_SMEM_BUDGET - st_epi_bytes(ST_MAX_CHUNK, L.a2p)) / ST_PAIR_STAGE);
        if (pstages > 6) pstages = 6;
        const size_t psmem = (size_t)pstages * ST_PAIR_STAGE + (2 * pstages + 4) * sizeof(uint64_t) + 16 +
                             st_epi_bytes(ST_MAX_CHUNK, L.a2p);
        const int64_t punits = (L.tiles / 2) * L.chunks;
        const int64_t clusters = punits < sms / 2 ? punits : sms / 2;
#define TSPN_LAUNCH_ST2(A2P)                                                                                      \
    do {                                                                                                          \
        TSPN_CUDA_OK(cudaFuncSetAttribute(span_head_tc2_kernel<A2P>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)psmem));                                                           \
        span_head_tc2_kernel<A2P><<<(unsigned)(2 * clusters), ST_THREADS, psmem, st>>>(                           \
            map_x, map_wh, L.tiles, cin, L.chunks, kslabs, pstages, d_conv_b, w2t, partial);                      \
    } while (0)
        switch (L.a2p) {
            case 4: TSPN_LAUNCH_ST2(4); break;
            case 8: TSPN_LAUNCH_ST2(8); break;
            case 12: TSPN_LAUNCH_ST2(12); break;
            default: TSPN_LAUNCH_ST2(16); break;
        }
#undef TSPN_LAUNCH_ST2
        TSPN_CUDA_OK(cudaGetLastError());
        dim3 fgrid2((unsigned)((t + 255) / 256), (unsigned)k);
        span_finish_kernel<<<fgrid2, 256, 0, st>>>(partial, L.chunks, L.tiles * ST_BM, L.a2p, a2, t, d_rows, d_pred_b,
                                                   d_out);
        TSPN_CUDA_OK(cudaGetLastError());
        return TSPN_OK;
    }
    const int64_t units = L.tiles * L.chunks;
    const unsigned grid = (unsigned)(units < sms ? units : sms);
#define TSPN_LAUNCH_ST(A2P)                                                                                       \
    do {                                                                                                          \
        TSPN_CUDA_OK(cudaFuncSetAttribute(span_head_tc_kernel<A2P>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          (int)smem_bytes));                                                      \
        span_head_tc_kernel<A2P><<<grid, ST_THREADS, smem_bytes, st>>>(map_x, map_w, L.tiles, cin, L.chunk,       \
                                                                       L.chunks, kslabs, stages, tmem_cols,       \
                                                                       d_conv_b, w2t, partial);                   \
    } while (0)
    switch (L.a2p) {
        case 4: TSPN_LAUNCH_ST(4); break;
        case 8: TSPN_LAUNCH_ST(8); break;
        case 12: TSPN_LAUNCH_ST(12); break;
        default: TSPN_LAUNCH_ST(16); break;
    }
#undef TSPN_LAUNCH_ST
    TSPN_CUDA_OK(cudaGetLastError());
    dim3 fgrid((unsigned)((t + 255) / 256), (unsigned)k);
    span_finish_kernel<<<fgrid, 256, 0, st>>>(partial, L.chunks, L.tiles * ST_BM, L.a2p, a2, t, d_rows, d_pred_b,
                                              d_out);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int64_t tspn_span_head_workspace_bytes(int64_t k, int cin, int t, int a2, int precision) {
    if (precision != TSPN_PREC_TENSOR || k <= 0 || cin <= 0 || t <= 0 || a2 <= 0) return 0;
    return st_layout(k, cin, t, a2).total;
}

}  // extern "C"
