// allreduce.cu -- one-shot all-reduce (sum) of the small shared-gradient bucket over NVLink peer memory.
//
// The only collective of the frame-sharded energy evaluation is a ~10 KB sum per iteration (seg-MLP gradients +
// loss).  At that size NCCL's ring/tree machinery is pure latency (~25-30 us in-graph on
// 8 GPUs); here every rank stages its bucket in a symmetric (peer-mapped) buffer, raises a flag in every peer's
// flag array, waits for all flags and then sums all ranks' buckets itself, reading them over NVLink with
// L1-bypassing loads in a FIXED rank order -- so every rank gets the bitwise identical result.
// Hazards: the staging area is double buffered by iteration parity (a rank can be at most one barrier ahead of a
// peer), flags carry monotonically increasing epochs (never reset), spinning is bounded so that a
// dead peer produces NaNs instead of a hung GPU.
#include "common.cuh"
#include "kernels.h"

namespace reart {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_volatile_f32(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float4 ld_volatile_f32x4(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// peer_base[r]: base of rank r's symmetric allocation, laid out as [2][n_pad] floats then [world] u32 flags
// (n_pad a multiple of 4, bases 16-byte aligned).  One CTA; every thread owns float4 slots, so a bucket of up to
// 4096 floats is staged, exchanged and summed in a single pass with all peer loads of a slot in flight at once.
__global__ void __launch_bounds__(1024) allreduce_oneshot_kernel(const unsigned long long* __restrict__ peer_base,
                                                                 int rank, int world, int n, int n_pad,
                                                                 unsigned* __restrict__ epoch,
                                                                 float* __restrict__ data) {
    __shared__ int s_ok;
    const unsigned e = *epoch + 1u;
    const int half = (int)(e & 1u);
    const int n4 = (n + 3) / 4;
    float4* mine = reinterpret_cast<float4*>(reinterpret_cast<float*>(peer_base[rank]) + (size_t)half * n_pad);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {                           // stage the local bucket
        float4 v;
        v.x = 4 * i < n ? data[4 * i] : 0.f;         v.y = 4 * i + 1 < n ? data[4 * i + 1] : 0.f;
        v.z = 4 * i + 2 < n ? data[4 * i + 2] : 0.f; v.w = 4 * i + 3 < n ? data[4 * i + 3] : 0.f;
        mine[i] = v;
    }
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < world) {
        __threadfence_system();                                                    // the CTA's staged bucket, system wide
        unsigned* peer_flags = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(peer_base[threadIdx.x]) + 2 * (size_t)n_pad);
        st_release_sys(peer_flags + rank, e);                                      // "rank's bucket of epoch e is staged"
        const unsigned* my_flags = reinterpret_cast<const unsigned*>(reinterpret_cast<float*>(peer_base[rank]) + 2 * (size_t)n_pad);
        long long spins = 0;
        while ((int)(ld_acquire_sys(my_flags + threadIdx.x) - e) < 0) {
            if (++spins > (1LL << 26)) { s_ok = 0; break; }                         // ~ seconds: a peer is gone
        }
    }
    __syncthreads();
    const bool ok = s_ok != 0;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r0 = 0; r0 < world; r0 += 8) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)                           // issue all peer loads first: eight NVLink reads in flight
                v[k] = (r0 + k < world)
                           ? ld_volatile_f32x4(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(peer_base[r0 + k]) + (size_t)half * n_pad) + i)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {                         // fixed rank order => identical bits on every rank
                sum.x += v[k].x; sum.y += v[k].y; sum.z += v[k].z; sum.w += v[k].w;
            }
        }
        const float bad = __int_as_float(0x7fc00000);
        if (4 * i < n) data[4 * i] = ok ? sum.x : bad;
        if (4 * i + 1 < n) data[4 * i + 1] = ok ? sum.y : bad;
        if (4 * i + 2 < n) data[4 * i + 2] = ok ? sum.z : bad;
        if (4 * i + 3 < n) data[4 * i + 3] = ok ? sum.w : bad;
    }
    __syncthreads();
    if (threadIdx.x == 0) *epoch = e;
}

int launch_allreduce_oneshot(const unsigned long long* peer_base, int rank, int world, int64_t n, int64_t n_pad,
                             unsigned* epoch, float* data, cudaStream_t stream) {
    if (n <= 0 || world <= 1) return kOk;
    if (world > 64 || n > n_pad || n > (1 << 20) || (n_pad & 3)) return kErrUnsupported;
    allreduce_oneshot_kernel<<<1, 1024, 0, stream>>>(peer_base, rank, world, (int)n, (int)n_pad, epoch, data);
    REART_CHECK_LAUNCH();
    return kOk;
}

}  // namespace reart
