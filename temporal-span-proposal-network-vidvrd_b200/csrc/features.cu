// features.cu — per-pair feature rows (lib/dataset/vrdataset.py:219-243).
//
//   [0,C) subject classeme | [C,2C) object classeme | 4000 subject motion BoW (4 blocks of
//   1000, each L1-normalised: vrdataset.py:227-236 + utils/miscellaneous.py:32-35) |
//   4000 object motion BoW | 3000 relative block = adaptive average pooling of the geometry
//   channels (0,1 | 2,3 | 5,6) over the pair's temporal overlap window to 500 bins each
//   ([SPEC] s4; the reference loads these 3000 columns precomputed from h5).
//
// HBM-write-bound: 4*F bytes per row; the per-tracklet blocks are re-read from L2.
#include <cuda_bf16.h>

#include "common.cuh"

namespace tspn {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// one warp per (tracklet, 1000-wide block)
__global__ void __launch_bounds__(128) normalize_motion_kernel(const float* __restrict__ motion, int64_t n_blocks,
                                                               float* __restrict__ out) {
    const int64_t blk = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (blk >= n_blocks) return;
    const int lane = threadIdx.x & 31;
    const float* src = motion + blk * TSPN_MOTION_BLOCK;
    float* dst = out + blk * TSPN_MOTION_BLOCK;
    float s = 0.0f;
    for (int i = lane; i < TSPN_MOTION_BLOCK; i += 32) s += fabsf(__ldg(src + i));
    s = warp_sum_f(s);
    if (s == 0.0f) s = 1.0f;                      // miscellaneous.py:34 — empty histograms stay zero
    for (int i = lane; i < TSPN_MOTION_BLOCK; i += 32) dst[i] = __ldg(src + i) / s;
}

constexpr int ASM_THREADS = 256;

template <bool BF16>
__device__ __forceinline__ void put(float* out, __nv_bfloat16* outb, int col, float v) {
    if (out) out[col] = v;
    if (BF16) outb[col] = __float2bfloat16(v);
}

// copy `n` floats (n % 4 == 0, src 16-byte aligned) to columns [col0, col0+n) of the row
template <bool BF16>
__device__ __forceinline__ void copy_block(const float* __restrict__ src, int n, float* out, __nv_bfloat16* outb,
                                           int col0, bool vec32, bool vec16) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int q = threadIdx.x; q < n / 4; q += ASM_THREADS) {
        const float4 v = __ldg(s4 + q);
        const int col = col0 + 4 * q;
        if (out) {
            if (vec32) {
                *reinterpret_cast<float4*>(out + col) = v;
            } else {
                out[col] = v.x; out[col + 1] = v.y; out[col + 2] = v.z; out[col + 3] = v.w;
            }
        }
        if (BF16) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            if (vec16) {
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(outb + col) = pk;
            } else {
                outb[col] = lo.x; outb[col + 1] = lo.y; outb[col + 2] = hi.x; outb[col + 3] = hi.y;
            }
        }
    }
}

template <bool BF16>
__global__ void __launch_bounds__(ASM_THREADS)
assemble_kernel(const int64_t* __restrict__ table, int nv, const float* __restrict__ cls, int n_classes,
                const float* __restrict__ motion_norm, const float* __restrict__ geo,
                const int32_t* __restrict__ overlap, const int64_t* __restrict__ rows, float* __restrict__ feat,
                int64_t ld_feat, __nv_bfloat16* __restrict__ feat_bf16, int64_t ld_bf16) {
    const int64_t r = blockIdx.x;
    const int64_t gp = rows ? rows[r] : r;
    float* out = feat ? feat + r * ld_feat : nullptr;
    __nv_bfloat16* outb = BF16 ? feat_bf16 + r * ld_bf16 : nullptr;
    const int C = n_classes;
    const int F = 2 * C + 2 * TSPN_MOTION_DIM + TSPN_REL_DIM;
    if (gp < 0) {                                  // padding row (e.g. K_eff < K): zeros
        for (int64_t col = threadIdx.x; col < (out ? ld_feat : 0); col += ASM_THREADS) out[col] = 0.0f;
        for (int64_t col = threadIdx.x; col < (BF16 ? ld_bf16 : 0); col += ASM_THREADS) outb[col] = __float2bfloat16(0.0f);
        return;
    }
    const int v = find_video(table, nv, TSPN_VT_PAIR_OFF, gp);
    const int64_t* row = table + (int64_t)v * TSPN_VT_COLS;
    const int n = (int)row[TSPN_VT_N];
    const int tp = (int)row[TSPN_VT_TP];
    const int p = (int)(gp - row[TSPN_VT_PAIR_OFF]);
    const int s = p / (n - 1);
    const int k = p - s * (n - 1);
    const int o = k + (k >= s ? 1 : 0);
    const int64_t ts = row[TSPN_VT_TRK_OFF] + s, to = row[TSPN_VT_TRK_OFF] + o;
    const int m0 = 2 * C, m1 = m0 + TSPN_MOTION_DIM, r0 = m1 + TSPN_MOTION_DIM;
    // ---- classemes ----
    for (int col = threadIdx.x; col < m0; col += ASM_THREADS)
        put<BF16>(out, outb, col, __ldg(cls + (col < C ? ts * C + col : to * C + (col - C))));
    // ---- motion blocks: 128-bit copies when the row offset 2C allows it ----
    const bool vec32 = (m0 & 3) == 0, vec16 = (m0 & 3) == 0;       // 16-byte (fp32) / 8-byte (bf16) stores
    copy_block<BF16>(motion_norm + ts * TSPN_MOTION_DIM, TSPN_MOTION_DIM, out, outb, m0, vec32, vec16);
    copy_block<BF16>(motion_norm + to * TSPN_MOTION_DIM, TSPN_MOTION_DIM, out, outb, m1, vec32, vec16);
    // ---- relative block: bin i of every pooled channel shares its frame range ----
    const float* g = geo + row[TSPN_VT_GEO_OFF] + (int64_t)p * TSPN_GEO_CHANNELS * tp;
    const int a = __ldg(overlap + 2 * gp), b = __ldg(overlap + 2 * gp + 1);
    const uint32_t len = b > a ? (uint32_t)(b - a) : 0u;
    for (int i = threadIdx.x; i < TSPN_REL_BINS; i += ASM_THREADS) {
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (len > 0) {
            const uint32_t st = ((uint32_t)i * len) / TSPN_REL_BINS;              // len < 2^22: fits 32 bits
            const uint32_t en = ((uint32_t)(i + 1) * len + TSPN_REL_BINS - 1) / TSPN_REL_BINS;
            const float inv = 1.0f / (float)(en - st);
#pragma unroll
            for (int slot = 0; slot < 6; ++slot) {
                const float* gc = g + (int64_t)(slot < 4 ? slot : slot + 1) * tp + a;
                float sacc = 0.0f;
                for (uint32_t f = st; f < en; ++f) sacc += __ldg(gc + f);
                acc[slot] = sacc * inv;
            }
        }
#pragma unroll
        for (int slot = 0; slot < 6; ++slot) put<BF16>(out, outb, r0 + slot * TSPN_REL_BINS + i, acc[slot]);
    }
    // zero the padding columns so that a padded row can be fed to TMA / vector loads
    if (out)
        for (int64_t col = F + threadIdx.x; col < ld_feat; col += ASM_THREADS) out[col] = 0.0f;
    if (BF16)
        for (int64_t col = F + threadIdx.x; col < ld_bf16; col += ASM_THREADS) outb[col] = __float2bfloat16(0.0f);
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int tspn_normalize_motion(const float* d_motion, int64_t n_tracklets, float* d_out, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(n_tracklets >= 0, TSPN_EBADARG, "tspn_normalize_motion: negative size");
    if (n_tracklets == 0) return TSPN_OK;
    TSPN_REQUIRE(d_motion && d_out, TSPN_EBADARG, "tspn_normalize_motion: null pointer");
    const int64_t n_blocks = n_tracklets * (TSPN_MOTION_DIM / TSPN_MOTION_BLOCK);
    normalize_motion_kernel<<<(unsigned)((n_blocks + 3) / 4), 128, 0, (cudaStream_t)stream>>>(d_motion, n_blocks,
                                                                                              d_out);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_assemble_features(const int64_t* d_table, int num_videos, int64_t total_pairs, const float* d_cls,
                           int n_classes, const float* d_motion_norm, const float* d_geo,
                           const int32_t* d_overlap, const int64_t* d_rows, int64_t n_rows, float* d_feat,
                           int64_t ld_feat, void* d_feat_bf16, int64_t ld_bf16, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(num_videos >= 0 && total_pairs >= 0 && n_rows >= 0 && n_classes > 0, TSPN_EBADARG,
                 "tspn_assemble_features: bad size");
    if (!d_rows) n_rows = total_pairs;
    if (n_rows == 0) return TSPN_OK;
    const int64_t F = 2 * (int64_t)n_classes + 2 * TSPN_MOTION_DIM + TSPN_REL_DIM;
    TSPN_REQUIRE(d_table && d_cls && d_motion_norm && d_geo && d_overlap && (d_feat || d_feat_bf16), TSPN_EBADARG,
                 "tspn_assemble_features: null pointer");
    TSPN_REQUIRE(!d_feat || (ld_feat >= F && (ld_feat & 3) == 0), TSPN_ESHAPE,
                 "tspn_assemble_features: ld_feat=%lld must be >= F=%lld and a multiple of 4", (long long)ld_feat,
                 (long long)F);
    TSPN_REQUIRE(!d_feat_bf16 || (ld_bf16 >= F && (ld_bf16 & 7) == 0), TSPN_ESHAPE,
                 "tspn_assemble_features: ld_bf16=%lld must be >= F=%lld and a multiple of 8", (long long)ld_bf16,
                 (long long)F);
    TSPN_REQUIRE(aligned16(d_feat) && aligned16(d_feat_bf16), TSPN_EALIGN,
                 "tspn_assemble_features: outputs must be 16-byte aligned");
    TSPN_REQUIRE(n_rows < (1ll << 31), TSPN_ESHAPE, "tspn_assemble_features: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_feat_bf16)
        assemble_kernel<true><<<(unsigned)n_rows, ASM_THREADS, 0, st>>>(
            d_table, num_videos, d_cls, n_classes, d_motion_norm, d_geo, d_overlap, d_rows, d_feat, ld_feat,
            reinterpret_cast<__nv_bfloat16*>(d_feat_bf16), ld_bf16);
    else
        assemble_kernel<false><<<(unsigned)n_rows, ASM_THREADS, 0, st>>>(
            d_table, num_videos, d_cls, n_classes, d_motion_norm, d_geo, d_overlap, d_rows, d_feat, ld_feat, nullptr,
            0);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // extern "C"
