// skin.cu -- soft-assignment skinning of the canonical cloud, forward and backward (HBM-bound).
//
// Replaces the torch composition of networks/model.py:63-69 (BaseModel), :161-165 (KinematicModel)
// and utils/model_utils.py:54-67 (compute_pc_transform):
//     out[t,n,:] = sum_p W[n,p] * (R[t,p] @ cano[n] + tr[t,p])
// without materialising the [T,P,N,3] intermediate the reference builds with bmm (P x the output).
// Exact zeros in W are skipped (gumbel_softmax(hard=True) / one_hot rows have one non-zero),
// so the common case costs one 3x4 transform per (t,n).
//
// Algorithmic HBM bytes: fwd reads 12N + 4NP + 48TP, writes 12TN (+12TN for the packed copy);
// bwd reads 12TN + 12N + 4NP + 48TP, writes 4NP + 48TP.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace reart {

constexpr int kSkinThreads = 128;
constexpr int kSkinFramesPerBlock = 8;

// grid (ceil(Npad/128), ceil(T/8)); dynamic smem: [8 frames][P][12] transforms
__global__ void __launch_bounds__(kSkinThreads) skin_fwd_kernel(const float* __restrict__ cano,
                                                                const float* __restrict__ W,
                                                                const float* __restrict__ R,
                                                                const float* __restrict__ tr, int T, int N, int P,
                                                                float* __restrict__ out, float* __restrict__ out_packed,
                                                                int n_pad) {
    extern __shared__ float sm_tf[];                          // [frames][P][12]: r00..r22, t0..t2
    const int t0 = blockIdx.y * kSkinFramesPerBlock;
    const int nt = min(kSkinFramesPerBlock, T - t0);
    for (int e = threadIdx.x; e < nt * P * 12; e += blockDim.x) {
        const int f = e / (P * 12), rem = e - f * P * 12, p = rem / 12, k = rem - p * 12;
        const int64_t tp = (int64_t)(t0 + f) * P + p;
        sm_tf[e] = (k < 9) ? R[tp * 9 + k] : tr[tp * 3 + (k - 9)];
    }
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_pad) return;
    const bool real = n < N;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (real) { cx = cano[3 * n]; cy = cano[3 * n + 1]; cz = cano[3 * n + 2]; }
    const float* __restrict__ w = W + (int64_t)(real ? n : 0) * P;
    for (int f = 0; f < nt; ++f) {
        float ax = 0.f, ay = 0.f, az = 0.f;
        if (real) {
            const float* tf = sm_tf + f * P * 12;
            for (int p = 0; p < P; ++p) {
                const float wp = __ldg(w + p);
                if (wp != 0.f) {
                    const float* m = tf + p * 12;
                    const float vx = cx * m[0] + cy * m[1] + cz * m[2] + m[9];
                    const float vy = cx * m[3] + cy * m[4] + cz * m[5] + m[10];
                    const float vz = cx * m[6] + cy * m[7] + cz * m[8] + m[11];
                    ax += wp * vx; ay += wp * vy; az += wp * vz;
                }
            }
            float* o = out + ((int64_t)(t0 + f) * N + n) * 3;
            o[0] = ax; o[1] = ay; o[2] = az;
        } else {
            ax = ay = az = INFINITY;                          // padding of the packed copy
        }
        if (out_packed) {
            float* g = out_packed + (int64_t)(t0 + f) * n_pad * 3 + (int64_t)(n >> 2) * kGroupFloats + (n & 3);
            g[0] = ax; g[4] = ay; g[8] = az;
        }
    }
}

int launch_skin_fwd(const float* cano, const float* W, const float* R, const float* tr, int64_t T, int64_t N, int64_t P,
                    float* out, float* out_packed, cudaStream_t stream) {
    if (T <= 0 || N <= 0) return kOk;
    if (P <= 0 || P > 256) return kErrUnsupported;
    const int64_t n_pad = out_packed ? padded_points(N) : N;
    dim3 grid((unsigned)ceil_div(n_pad, kSkinThreads), (unsigned)ceil_div(T, kSkinFramesPerBlock));
    const size_t smem = (size_t)kSkinFramesPerBlock * P * 12 * sizeof(float);
    skin_fwd_kernel<<<grid, kSkinThreads, smem, stream>>>(cano, W, R, tr, (int)T, (int)N, (int)P, out, out_packed,
                                                          (int)n_pad);
    REART_CHECK_LAUNCH();
    return kOk;
}

// ----------------------------------------------------------------------------- backward
// g [T,N,3] -> gW [N,P] (+=), gR [T,P,9] (+=), gtr [T,P,3] (+=); outputs must be zero on entry.
//   gW[n,p]  = sum_t g[t,n] . (R[t,p] c_n + tr[t,p])                (dense in p: straight-through grads)
//   gR[t,p]  = sum_n W[n,p] g[t,n] c_n^T ;  gtr[t,p] = sum_n W[n,p] g[t,n]   (only where W != 0)
// grid (ceil(N/128), ceil(T/8)).  gW partial sums over the block's 8 frames are kept in shared memory
// ([128][P+1] floats) and added to global with one atomic per (n,p) per block; gR/gtr are reduced in
// shared memory across the block's points, then one atomic per (t,p,k) per block.
__global__ void __launch_bounds__(kSkinThreads) skin_bwd_kernel(const float* __restrict__ cano,
                                                                const float* __restrict__ W,
                                                                const float* __restrict__ R,
                                                                const float* __restrict__ tr,
                                                                const float* __restrict__ g, int T, int N, int P,
                                                                float* __restrict__ gW, float* __restrict__ gR,
                                                                float* __restrict__ gtr) {
    extern __shared__ float sm[];
    float* sm_tf = sm;                                         // [frames][P][12]
    float* sm_acc = sm_tf + kSkinFramesPerBlock * P * 12;      // [frames][P][12] reduction of gR|gtr
    float* sm_gw = sm_acc + kSkinFramesPerBlock * P * 12;      // [128][P+1]
    const int t0 = blockIdx.y * kSkinFramesPerBlock;
    const int nt = min(kSkinFramesPerBlock, T - t0);
    for (int e = threadIdx.x; e < nt * P * 12; e += blockDim.x) {
        const int f = e / (P * 12), rem = e - f * P * 12, p = rem / 12, k = rem - p * 12;
        const int64_t tp = (int64_t)(t0 + f) * P + p;
        sm_tf[e] = (k < 9) ? R[tp * 9 + k] : tr[tp * 3 + (k - 9)];
        sm_acc[e] = 0.f;
    }
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const bool real = n < N;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (real) { cx = cano[3 * n]; cy = cano[3 * n + 1]; cz = cano[3 * n + 2]; }
    float* my_gw = sm_gw + threadIdx.x * (P + 1);
    for (int p = 0; p < P; ++p) my_gw[p] = 0.f;
    if (real) {
        const float* __restrict__ w = W + (int64_t)n * P;
        for (int f = 0; f < nt; ++f) {
            const float* gg = g + ((int64_t)(t0 + f) * N + n) * 3;
            const float gx = gg[0], gy = gg[1], gz = gg[2];
            const float* tf = sm_tf + f * P * 12;
            float* acc = sm_acc + f * P * 12;
            for (int p = 0; p < P; ++p) {
                const float* m = tf + p * 12;
                const float vx = cx * m[0] + cy * m[1] + cz * m[2] + m[9];
                const float vy = cx * m[3] + cy * m[4] + cz * m[5] + m[10];
                const float vz = cx * m[6] + cy * m[7] + cz * m[8] + m[11];
                my_gw[p] += gx * vx + gy * vy + gz * vz;
                const float wp = __ldg(w + p);
                if (wp != 0.f) {
                    float* a = acc + p * 12;
                    const float wx = wp * gx, wy = wp * gy, wz = wp * gz;
                    atomicAdd(a + 0, wx * cx); atomicAdd(a + 1, wx * cy); atomicAdd(a + 2, wx * cz);
                    atomicAdd(a + 3, wy * cx); atomicAdd(a + 4, wy * cy); atomicAdd(a + 5, wy * cz);
                    atomicAdd(a + 6, wz * cx); atomicAdd(a + 7, wz * cy); atomicAdd(a + 8, wz * cz);
                    atomicAdd(a + 9, wx); atomicAdd(a + 10, wy); atomicAdd(a + 11, wz);
                }
            }
        }
        float* o = gW + (int64_t)n * P;
        if (gridDim.y == 1) { for (int p = 0; p < P; ++p) o[p] = my_gw[p]; }
        else { for (int p = 0; p < P; ++p) atomicAdd(o + p, my_gw[p]); }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nt * P * 12; e += blockDim.x) {
        const float v = sm_acc[e];
        if (v != 0.f) {
            const int f = e / (P * 12), rem = e - f * P * 12, p = rem / 12, k = rem - p * 12;
            const int64_t tp = (int64_t)(t0 + f) * P + p;
            if (k < 9) atomicAdd(gR + tp * 9 + k, v);
            else atomicAdd(gtr + tp * 3 + (k - 9), v);
        }
    }
}

int launch_skin_bwd(const float* cano, const float* W, const float* R, const float* tr, const float* g, int64_t T,
                    int64_t N, int64_t P, float* gW, float* gR, float* gtr, cudaStream_t stream) {
    if (P <= 0 || P > 256) return kErrUnsupported;
    if (N * P > 0 && cudaMemsetAsync(gW, 0, sizeof(float) * (size_t)(N * P), stream) != cudaSuccess) return kErrLaunch;
    if (T * P > 0) {
        if (cudaMemsetAsync(gR, 0, sizeof(float) * (size_t)(T * P * 9), stream) != cudaSuccess) return kErrLaunch;
        if (cudaMemsetAsync(gtr, 0, sizeof(float) * (size_t)(T * P * 3), stream) != cudaSuccess) return kErrLaunch;
    }
    if (T <= 0 || N <= 0) return kOk;
    dim3 grid((unsigned)ceil_div(N, kSkinThreads), (unsigned)ceil_div(T, kSkinFramesPerBlock));
    const size_t smem = ((size_t)2 * kSkinFramesPerBlock * P * 12 + (size_t)kSkinThreads * (P + 1)) * sizeof(float);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(skin_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return kErrUnsupported;
    skin_bwd_kernel<<<grid, kSkinThreads, smem, stream>>>(cano, W, R, tr, g, (int)T, (int)N, (int)P, gW, gR, gtr);
    REART_CHECK_LAUNCH();
    return kOk;
}

}  // namespace reart
