// fps.cu -- furthest point sampling and ball query for all frames at once.
//
// Replaces the two reachable kernels of the reference's in-tree PointNet++ extension
// (networks/pointnet_lib/src/sampling_gpu.cu:93-209 furthest_point_sampling_kernel,
//  networks/pointnet_lib/src/ball_query_gpu.cu:9-45 ball_query_kernel_fast), bound at
// networks/pointnet_lib/pointnet2_utils.py:29,263 as pointnet2_cuda.*_wrapper.
// FPS is sequential in the number of samples; one CTA per cloud keeps the running min-distances in
// registers and reduces the arg-max with warp shuffles on a packed 64-bit key
// (dist_bits << 32 | ~index) so the lowest index wins exact ties (pinned).
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace reart {

constexpr int kFpsThreads = 512;
constexpr int kFpsMaxPerThread = 64;                          // N <= 32768 per cloud

template <int PPT>
__global__ void __launch_bounds__(kFpsThreads) fps_kernel(const float* __restrict__ xyz, int n, int m,
                                                          int* __restrict__ out) {
    __shared__ u64 warp_best[kFpsThreads / 32];
    __shared__ int s_old;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ pts = xyz + (int64_t)blockIdx.x * n * 3;
    int* __restrict__ o = out + (int64_t)blockIdx.x * m;
    float temp[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) temp[k] = 1e10f;
    int old = 0;
    if (tid == 0 && m > 0) o[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float ox = __ldg(pts + 3 * old), oy = __ldg(pts + 3 * old + 1), oz = __ldg(pts + 3 * old + 2);
        u64 best = 0;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int i = k * kFpsThreads + tid;
            if (i < n) {
                const float d = sqdist_scalar(__ldg(pts + 3 * i), __ldg(pts + 3 * i + 1), __ldg(pts + 3 * i + 2), ox, oy, oz);
                const float d2 = fminf(d, temp[k]);
                temp[k] = d2;
                const u64 key = ((u64)__float_as_uint(d2) << 32) | (u64)(0xffffffffu - (unsigned)i);
                best = key > best ? key : best;
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const u64 other = __shfl_xor_sync(0xffffffffu, best, s);
            best = other > best ? other : best;
        }
        if (lane == 0) warp_best[warp] = best;
        __syncthreads();
        if (warp == 0) {
            u64 b = lane < kFpsThreads / 32 ? warp_best[lane] : 0;
#pragma unroll
            for (int s = 8; s > 0; s >>= 1) {
                const u64 other = __shfl_xor_sync(0xffffffffu, b, s);
                b = other > b ? other : b;
            }
            if (lane == 0) {
                const int idx = (int)(0xffffffffu - (unsigned)(b & 0xffffffffu));
                s_old = idx;
                o[j] = idx;
            }
        }
        __syncthreads();
        old = s_old;
    }
}

int launch_fps(const float* xyz, int64_t B, int64_t N, int64_t m, int* out, cudaStream_t stream) {
    if (B <= 0 || m <= 0) return kOk;
    if (N <= 0 || N > (int64_t)kFpsThreads * kFpsMaxPerThread) return kErrUnsupported;
    const int ppt = (int)ceil_div(N, kFpsThreads);
#define REART_FPS_CASE(P) if (ppt <= P) { fps_kernel<P><<<(unsigned)B, kFpsThreads, 0, stream>>>(xyz, (int)N, (int)m, out); REART_CHECK_LAUNCH(); return kOk; }
    REART_FPS_CASE(1) REART_FPS_CASE(2) REART_FPS_CASE(4) REART_FPS_CASE(8) REART_FPS_CASE(16) REART_FPS_CASE(32)
    REART_FPS_CASE(64)
#undef REART_FPS_CASE
    return kErrUnsupported;
}


// Clouds beyond the register budget of fps_kernel (N > 32 768): the running min-distances live in a caller-provided
// temp [B,N] buffer (what the reference's kernel does for every size, sampling_gpu.cu:93-209; its Python wrapper allocates
// exactly this tensor, networks/pointnet_lib/pointnet2_utils.py:26-29), so there is no size limit.  Same arithmetic, same
// packed-key arg-max with the lowest index on ties, hence the same samples as fps_kernel where both apply (tested).
constexpr int kFpsLargeThreads = 1024;
__global__ void __launch_bounds__(kFpsLargeThreads) fps_large_kernel(const float* __restrict__ xyz, int n, int m,
                                                                     float* __restrict__ temp, int* __restrict__ out) {
    __shared__ u64 warp_best[kFpsLargeThreads / 32];
    __shared__ int s_old;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ pts = xyz + (int64_t)blockIdx.x * n * 3;
    float* __restrict__ tmp = temp + (int64_t)blockIdx.x * n;
    int* __restrict__ o = out + (int64_t)blockIdx.x * m;
    for (int i = tid; i < n; i += kFpsLargeThreads) tmp[i] = 1e10f;
    int old = 0;
    if (tid == 0 && m > 0) o[0] = 0;
    __syncthreads();
    for (int j = 1; j < m; ++j) {
        const float ox = __ldg(pts + 3 * old), oy = __ldg(pts + 3 * old + 1), oz = __ldg(pts + 3 * old + 2);
        u64 best = 0;
        for (int i = tid; i < n; i += kFpsLargeThreads) {
            const float d = sqdist_scalar(__ldg(pts + 3 * i), __ldg(pts + 3 * i + 1), __ldg(pts + 3 * i + 2), ox, oy, oz);
            const float d2 = fminf(d, tmp[i]);
            tmp[i] = d2;
            const u64 key = ((u64)__float_as_uint(d2) << 32) | (u64)(0xffffffffu - (unsigned)i);
            best = key > best ? key : best;
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const u64 other = __shfl_xor_sync(0xffffffffu, best, s);
            best = other > best ? other : best;
        }
        if (lane == 0) warp_best[warp] = best;
        __syncthreads();
        if (warp == 0) {
            u64 b = warp_best[lane];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const u64 other = __shfl_xor_sync(0xffffffffu, b, s);
                b = other > b ? other : b;
            }
            if (lane == 0) {
                const int idx = (int)(0xffffffffu - (unsigned)(b & 0xffffffffu));
                s_old = idx;
                o[j] = idx;
            }
        }
        __syncthreads();
        old = s_old;
    }
}

int launch_fps_large(const float* xyz, int64_t B, int64_t N, int64_t m, float* temp, int* out, cudaStream_t stream) {
    if (B <= 0 || m <= 0) return kOk;
    if (N <= 0 || !temp) return kErrInvalidArg;
    fps_large_kernel<<<(unsigned)B, kFpsLargeThreads, 0, stream>>>(xyz, (int)N, (int)m, temp, out);
    REART_CHECK_LAUNCH();
    return kOk;
}

// one thread per query centre: first `nsample` points with d2 < radius^2, in index order; the first hit
// pre-fills every slot (ball_query_gpu.cu:33-41); idx untouched when the ball is empty (caller zero-fills).
__global__ void ball_query_kernel(int n, int m, float radius2, int nsample, const float* __restrict__ new_xyz,
                                  const float* __restrict__ xyz, int* __restrict__ idx) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float* q = new_xyz + ((int64_t)b * m + i) * 3;
    const float* p = xyz + (int64_t)b * n * 3;
    int* o = idx + ((int64_t)b * m + i) * nsample;
    const float qx = q[0], qy = q[1], qz = q[2];
    int cnt = 0;
    for (int k = 0; k < n; ++k) {
        const float d2 = sqdist_scalar(qx, qy, qz, __ldg(p + 3 * k), __ldg(p + 3 * k + 1), __ldg(p + 3 * k + 2));
        if (d2 < radius2) {
            if (cnt == 0)
                for (int l = 0; l < nsample; ++l) o[l] = k;
            o[cnt] = k;
            if (++cnt >= nsample) break;
        }
    }
}

int launch_ball_query(const float* new_xyz, const float* xyz, int64_t B, int64_t N, int64_t m, float radius,
                      int nsample, int* idx, cudaStream_t stream) {
    if (B <= 0 || m <= 0 || nsample <= 0) return kOk;
    if (B > 65535) return kErrUnsupported;
    dim3 grid((unsigned)ceil_div(m, 128), (unsigned)B);
    ball_query_kernel<<<grid, 128, 0, stream>>>((int)N, (int)m, radius * radius, nsample, new_xyz, xyz, idx);
    REART_CHECK_LAUNCH();
    return kOk;
}

}  // namespace reart
