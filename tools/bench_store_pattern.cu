// bench_store_pattern.cu — how fast can a B200 absorb the geometry kernel's store pattern?
// Store-only kernels (no math) with the address stream of pair_geo_kernel and some alternatives,
// to separate "HBM write pattern" limits from the kernel's own latency.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bench_store_pattern.bin tools/bench_store_pattern.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int MODE>
__device__ __forceinline__ void st4(float* p, float4 v) {
    if (MODE == 0) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else if (MODE == 1) asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else if (MODE == 2) asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// A: sequential grid-stride fill
template <int MODE>
__global__ void fill_seq(float* out, int64_t n4) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        st4<MODE>(out + 4 * i, v);
}

// B: the geometry kernel's pattern.  item = (video, s, group of OG objects); chunk-major loop; per
// step every thread stores 8 float4 (one per channel row, rows tp floats apart).
template <int MODE, int THREADS, bool OBJ_MAJOR>
__global__ void __launch_bounds__(THREADS) fill_geo(float* out, int n, int tp, int og, int chunk_frames) {
    const int groups = (n - 1 + og - 1) / og;
    const int64_t item = blockIdx.x;
    const int v = (int)(item / (n * groups));
    const int local = (int)(item - (int64_t)v * n * groups);
    const int s = local / groups, k0 = (local - s * groups) * og;
    const int nobj = min(og, n - 1 - k0);
    const int nchunks = (tp + chunk_frames - 1) / chunk_frames;
    float* base = out + ((int64_t)v * n * (n - 1) + (int64_t)s * (n - 1) + k0) * 8 * tp;
    const float4 val = make_float4(1.f, 2.f, 3.f, (float)threadIdx.x);
    const int steps = nchunks * nobj;
    for (int q = 0; q < steps; ++q) {
        int c, jj;
        if (OBJ_MAJOR) { jj = q / nchunks; c = q - jj * nchunks; } else { c = q / nobj; jj = q - c * nobj; }
        const int t0 = c * chunk_frames + threadIdx.x * 4;
        if (t0 < tp) {
            float* g = base + (int64_t)jj * 8 * tp + t0;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) st4<MODE>(g + (int64_t)ch * tp, val);
        }
    }
}

// D: chunk-split: item = (video, s, group, chunk); adjacent blocks (or the CTAs of one cluster) write
// adjacent chunks of the same rows at about the same time
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) fill_geo_split(float* out, int n, int tp, int og, int nchunks) {
    extern __shared__ float dummy[];
    const int groups = (n - 1 + og - 1) / og;
    const int c = blockIdx.x % nchunks;
    const int64_t item = blockIdx.x / nchunks;
    const int v = (int)(item / (n * groups));
    const int local = (int)(item - (int64_t)v * n * groups);
    const int s = local / groups, k0 = (local - s * groups) * og;
    const int nobj = min(og, n - 1 - k0);
    float* base = out + ((int64_t)v * n * (n - 1) + (int64_t)s * (n - 1) + k0) * 8 * tp;
    const float4 val = make_float4(1.f, 2.f, 3.f, (float)threadIdx.x);
    const int t0 = c * THREADS * 4 + threadIdx.x * 4;
    if (t0 >= tp) return;
    for (int jj = 0; jj < nobj; ++jj) {
        float* g = base + (int64_t)jj * 8 * tp + t0;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) st4<MODE>(g + (int64_t)ch * tp, val);
    }
}

template <typename K, typename... Args>
static void launch_cluster(K kernel, int grid, int threads, int cluster, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kernel, args...));
}

// C: one warp per channel row: a warp writes a whole row of tp floats contiguously (8 KB), a CTA of 8
// warps covers one pair (64 KB contiguous)
template <int MODE>
__global__ void __launch_bounds__(256) fill_rows(float* out, int64_t pairs, int tp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 val = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int64_t p = blockIdx.x; p < pairs; p += gridDim.x) {
        float* g = out + (p * 8 + warp) * tp;
        for (int t = lane * 4; t < tp; t += 128) st4<MODE>(g + t, val);
    }
}

template <typename F>
static float time_it(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return ms / reps;
}

int main() {
    const int V = 16, N = 64, TP = 2000, OG = 8;
    const int64_t pairs = (int64_t)V * N * (N - 1);
    const int64_t floats = pairs * 8 * TP;
    float* out; CK(cudaMalloc(&out, floats * 4));
    const double gb = floats * 4 / 1e9;
    printf("buffer %.3f GB\n", gb);
    auto report = [&](const char* name, float ms) { printf("%-46s %8.3f ms  %8.1f GB/s\n", name, ms, gb / (ms * 1e-3)); };
    const int reps = 10;
    report("A seq fill .cs      grid 148*8x256", time_it([&] { fill_seq<0><<<148 * 8, 256>>>(out, floats / 4); }, reps));
    report("A seq fill default  grid 148*8x256", time_it([&] { fill_seq<1><<<148 * 8, 256>>>(out, floats / 4); }, reps));
    report("A seq fill .wt", time_it([&] { fill_seq<2><<<148 * 8, 256>>>(out, floats / 4); }, reps));
    report("A seq fill no_allocate", time_it([&] { fill_seq<3><<<148 * 8, 256>>>(out, floats / 4); }, reps));
    const int items = V * N * ((N - 1 + OG - 1) / OG);
    report("B geo pattern .cs  chunk-major 128thr", time_it([&] { fill_geo<0, 128, false><<<items, 128>>>(out, N, TP, OG, 512); }, reps));
    report("B geo pattern dflt chunk-major 128thr", time_it([&] { fill_geo<1, 128, false><<<items, 128>>>(out, N, TP, OG, 512); }, reps));
    report("B geo pattern .wt  chunk-major 128thr", time_it([&] { fill_geo<2, 128, false><<<items, 128>>>(out, N, TP, OG, 512); }, reps));
    report("B geo pattern .cs  object-major 128thr", time_it([&] { fill_geo<0, 128, true><<<items, 128>>>(out, N, TP, OG, 512); }, reps));
    report("B geo pattern dflt object-major 128thr", time_it([&] { fill_geo<1, 128, true><<<items, 128>>>(out, N, TP, OG, 512); }, reps));
    report("B geo pattern .cs  chunk-major 256thr/1024", time_it([&] { fill_geo<0, 256, false><<<items, 256>>>(out, N, TP, OG, 1024); }, reps));
    report("B geo pattern .cs  object-major 256thr/1024", time_it([&] { fill_geo<0, 256, true><<<items, 256>>>(out, N, TP, OG, 1024); }, reps));
    report("B geo pattern .cs  chunk-major 512thr/2048", time_it([&] { fill_geo<0, 512, false><<<items, 512>>>(out, N, TP, OG, 2048); }, reps));
    for (int smem_kb : {0, 64, 100, 140}) {
        char nm[96];
        CK(cudaFuncSetAttribute(fill_geo<0, 512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(fill_geo<0, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        snprintf(nm, sizeof nm, "B 512thr/2048 .cs, %d KB smem/CTA", smem_kb);
        report(nm, time_it([&] { fill_geo<0, 512, false><<<items, 512, smem_kb * 1024>>>(out, N, TP, OG, 2048); }, reps));
        snprintf(nm, sizeof nm, "B 256thr/1024 obj-major .cs, %d KB smem/CTA", smem_kb / 2);
        report(nm, time_it([&] { fill_geo<0, 256, true><<<items, 256, smem_kb * 512>>>(out, N, TP, OG, 1024); }, reps));
    }
    report("B 1024thr/4096 .cs", time_it([&] { fill_geo<0, 1024, false><<<items, 1024>>>(out, N, TP, OG, 4096); }, reps));
    report("D split 128thr x4 chunks, no cluster", time_it([&] { fill_geo_split<0, 128><<<items * 4, 128>>>(out, N, TP, OG, 4); }, reps));
    report("D split 128thr x4 chunks, cluster 4", time_it([&] { launch_cluster(fill_geo_split<0, 128>, items * 4, 128, 4, 0, out, N, TP, OG, 4); }, reps));
    report("D split 256thr x2 chunks, no cluster", time_it([&] { fill_geo_split<0, 256><<<items * 2, 256>>>(out, N, TP, OG, 2); }, reps));
    report("D split 256thr x2 chunks, cluster 2", time_it([&] { launch_cluster(fill_geo_split<0, 256>, items * 2, 256, 2, 0, out, N, TP, OG, 2); }, reps));
    CK(cudaFuncSetAttribute(fill_geo_split<0, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    report("D split 128thr x4, cluster 4, 34KB smem", time_it([&] { launch_cluster(fill_geo_split<0, 128>, items * 4, 128, 4, 34 * 1024, out, N, TP, OG, 4); }, reps));
    report("D split 128thr x4, no cluster, 34KB smem", time_it([&] { fill_geo_split<0, 128><<<items * 4, 128, 34 * 1024>>>(out, N, TP, OG, 4); }, reps));
    report("C row-per-warp .cs   grid 148*8", time_it([&] { fill_rows<0><<<148 * 8, 256>>>(out, pairs, TP); }, reps));
    report("C row-per-warp dflt  grid 148*8", time_it([&] { fill_rows<1><<<148 * 8, 256>>>(out, pairs, TP); }, reps));
    report("C row-per-warp .cs   grid 148*4", time_it([&] { fill_rows<0><<<148 * 4, 256>>>(out, pairs, TP); }, reps));
    CK(cudaFree(out));
    return 0;
}
