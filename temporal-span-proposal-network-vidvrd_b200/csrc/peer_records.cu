// peer_records.cu — the path's one exchange step without a collective launch: every rank WRITES its per-video
// triplet records straight into every peer's gather buffer over NVLink (peer-mapped symmetric memory), and a
// flag per (step, producer) tells the consumers when a step's records have landed.
//
// Replaces, in the serving loop, the per-step ncclAllGather of the [V, 200, 8] int32 records (lib/utils/comm.py:48-88
// is what the reference would use).  At 102 KB per rank the collective itself is nothing; what costs is launching an
// NCCL kernel every 0.77 ms beside a persistent kernel that holds every SM (measured at 8 GPUs: +0.14 ms per step,
// profiles/r2_e2e_scaling.md).  Here the exchange is 26 small CTAs of plain stores plus two 32-thread flag kernels.
//
// Protocol (ring of `ring` slots, step s uses slot s % ring; all ranks run the same number of steps):
//   producer rank p, step s:  wait until every consumer c has released step s - ring   (credit[slot][c] on p)
//                             store its records into buf[slot][p] of EVERY rank, fence, then written[slot][p] = s
//                             on every rank (release, system scope)
//   consumer rank c, step s:  wait until written[slot][p] >= s for every p (acquire), copy buf[slot] out,
//                             then credit[slot][c] = s on every rank
// Flags only ever grow, so a late reader sees ">= s".  Every wait is a bounded spin: on expiry the kernel raises the
// error word of the state block and returns instead of hanging the device.
#include "common.cuh"

namespace tspn {

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int PR_MAX_WORLD = 32;

// state: [0] error word (0 = fine, 1 = a credit wait expired, 2 = a records wait expired), [1] reserved,
// [2 + slot] CTA counter of the slot's scatter in flight (steps of different slots may run concurrently on two
// compute streams; two steps of the same slot are `ring` steps apart and ordered by the credits).  The step number is
// the caller's (1, 2, 3, ...: the same sequence on every rank), not a device counter: kernels of consecutive steps
// launched on different streams may run in either order.
__global__ void __launch_bounds__(256)
records_scatter_kernel(const int4* __restrict__ src, int64_t n_vec, const uint64_t* __restrict__ peer_bufs,
                       const uint64_t* __restrict__ peer_flags, const int* __restrict__ local_flags, int world, int rank,
                       int ring, int s, int* __restrict__ state, int max_spin) {
    __shared__ int s_last;
    const int slot = s % ring;
    if ((int)threadIdx.x < world) {      // credits: consumer threadIdx.x has copied this slot's previous step out
        const int* f = local_flags + ((int64_t)(ring + slot)) * world + threadIdx.x;
        int spins = 0;
        while (ld_acquire_sys(f) < s - ring && ++spins < max_spin) __nanosleep(200);
        if (spins >= max_spin) atomicExch(state, 1);
    }
    __syncthreads();
    const int64_t base = ((int64_t)slot * world + rank) * n_vec;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int4 v = __ldg(src + i);
        for (int r = 0; r < world; ++r) reinterpret_cast<int4*>(peer_bufs[r])[base + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(state + 2 + slot, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (s_last) {                        // every CTA's stores are fenced: publish the step
        __threadfence_system();
        if ((int)threadIdx.x < world)
            st_release_sys(reinterpret_cast<int*>(peer_flags[threadIdx.x]) + (int64_t)slot * world + rank, s);
        if (threadIdx.x == 0) state[2 + slot] = 0;
    }
}

__global__ void __launch_bounds__(32)
records_wait_kernel(const int* __restrict__ local_flags, int world, int ring, int s, int* __restrict__ state,
                    int max_spin) {
    const int slot = s % ring;
    if ((int)threadIdx.x < world) {
        const int* f = local_flags + (int64_t)slot * world + threadIdx.x;
        int spins = 0;
        while (ld_acquire_sys(f) < s && ++spins < max_spin) __nanosleep(200);
        if (spins >= max_spin) atomicExch(state, 2);
    }
}

__global__ void __launch_bounds__(32)
records_release_kernel(const uint64_t* __restrict__ peer_flags, int world, int rank, int ring, int s) {
    const int slot = s % ring;
    if ((int)threadIdx.x < world)
        st_release_sys(reinterpret_cast<int*>(peer_flags[threadIdx.x]) + ((int64_t)(ring + slot)) * world + rank, s);
}

}  // namespace tspn

using namespace tspn;

extern "C" {

int tspn_records_scatter(const int32_t* d_records, int64_t n_int32, const uint64_t* d_peer_bufs,
                         const uint64_t* d_peer_flags, const int32_t* d_local_flags, int world, int rank, int ring,
                         int step, int32_t* d_state, int max_spin, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(world >= 1 && world <= PR_MAX_WORLD && rank >= 0 && rank < world && ring >= 2 && n_int32 >= 0 &&
                     max_spin > 0 && step >= 1, TSPN_EBADARG, "tspn_records_scatter: bad argument");
    TSPN_REQUIRE((n_int32 & 3) == 0, TSPN_ESHAPE, "tspn_records_scatter: records must be a multiple of 16 bytes");
    TSPN_REQUIRE(d_records && d_peer_bufs && d_peer_flags && d_local_flags && d_state, TSPN_EBADARG,
                 "tspn_records_scatter: null pointer");
    TSPN_REQUIRE(aligned16(d_records), TSPN_EALIGN, "tspn_records_scatter: records must be 16-byte aligned");
    const int64_t n_vec = n_int32 / 4;
    int64_t blocks = (n_vec + 255) / 256;
    if (blocks > 32) blocks = 32;
    if (blocks < 1) blocks = 1;
    records_scatter_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const int4*>(d_records), n_vec, d_peer_bufs, d_peer_flags, d_local_flags, world, rank, ring,
        step, d_state, max_spin);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_records_wait(const int32_t* d_local_flags, int world, int ring, int step, int32_t* d_state, int max_spin,
                      void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(world >= 1 && world <= PR_MAX_WORLD && ring >= 2 && max_spin > 0 && step >= 1 && d_local_flags &&
                     d_state, TSPN_EBADARG, "tspn_records_wait: bad argument");
    records_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_local_flags, world, ring, step, d_state, max_spin);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

int tspn_records_release(const uint64_t* d_peer_flags, int world, int rank, int ring, int step, void* stream) {
    TSPN_ARCH_OK();
    TSPN_REQUIRE(world >= 1 && world <= PR_MAX_WORLD && rank >= 0 && rank < world && ring >= 2 && step >= 1 &&
                     d_peer_flags, TSPN_EBADARG, "tspn_records_release: bad argument");
    records_release_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_peer_flags, world, rank, ring, step);
    TSPN_CUDA_OK(cudaGetLastError());
    return TSPN_OK;
}

}  // extern "C"
