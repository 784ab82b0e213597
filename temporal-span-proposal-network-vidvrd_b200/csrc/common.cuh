// common.cuh — shared host/device helpers of libtspn_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tspn_b200.h"

namespace tspn {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_arch();                       // TSPN_OK or TSPN_EARCH (cached per device)
int num_sms();

#define TSPN_REQUIRE(cond, code, ...)                \
    do {                                             \
        if (!(cond)) {                               \
            ::tspn::set_error(__VA_ARGS__);          \
            return (code);                           \
        }                                            \
    } while (0)

#define TSPN_CUDA_OK(expr)                                                          \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            ::tspn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                  \
            return TSPN_ECUDA;                                                      \
        }                                                                           \
    } while (0)

#define TSPN_ARCH_OK()                              \
    do {                                            \
        int _a = ::tspn::check_arch();              \
        if (_a != TSPN_OK) return _a;               \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- video table ----------------------------------------------------------------------
struct VideoRow {
    int64_t c[TSPN_VT_COLS];
};

// Largest v with table[v][col] <= x.  The column is non-decreasing.
__device__ __forceinline__ int find_video(const int64_t* __restrict__ table, int nv, int col, int64_t x) {
    int lo = 0, hi = nv - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(table + (int64_t)mid * TSPN_VT_COLS + col) <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// True total of an offset column (tracklets, pairs, geo floats, work items, boxes, scores) of the batch: the
// sentinel row `nv` of the table (include/tspn_b200.h).  The scalar totals a launch was given are upper bounds.
__device__ __forceinline__ int64_t table_total(const int64_t* __restrict__ table, int nv, int col) {
    return __ldg(table + (int64_t)nv * TSPN_VT_COLS + col);
}

// ---- PTX wrappers: mbarrier, TMA ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 2-D tiled TMA load, completion signalled on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2),
        "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 1-D bulk copy global -> shared (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// streaming 128-bit store (written once, never re-read by this kernel)
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}


// ---- host: shared-memory carve-out ------------------------------------------------------------
// An SM's L1 / shared-memory split is a per-SM mode that only changes when the SM is idle.  The pair kernel needs
// 134 KB per CTA; left to the driver that means a 164 KB carve-out, i.e. 29 KB for anything that wants to
// co-reside with it - and a kernel preferring another split waits for (or takes over) whole SMs.  Every kernel
// of the pair stage therefore asks for the maximum carve-out (228 KB): the pair kernel leaves 93 KB per SM to the
// side stream's kernels, and no launch ever makes an SM drain to switch modes.
template <typename Kernel>
inline void prefer_max_smem(Kernel kernel) {
    cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributePreferredSharedMemoryCarveout,
                         (int)cudaSharedmemCarveoutMaxShared);
}

// ---- host: tensor maps ---------------------------------------------------------------------
// cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint (no link-time libcuda).
int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes /* rank-1 */,
                      const uint32_t* box, CUtensorMapSwizzle swizzle);

// ---- the windowed all-pairs kernel (geo_windowed.cu), launched by the MAIN phase of tspn_pair_geo_viou_windowed ----
int launch_pair_geo_windowed(const int64_t* d_table, int num_videos, int64_t total_pairs, const float* d_boxes,
                             const int32_t* d_span, float* d_geo, const int64_t* d_geo_off, unsigned long long* fx,
                             int32_t* d_overlap, unsigned int* d_queue, int geo_chunk, int max_chunks, int reserve,
                             bool clip, cudaStream_t st);

}  // namespace tspn
