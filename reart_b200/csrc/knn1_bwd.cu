// knn1_bwd.cu -- backward of the K=1 search (memory-bound, O(B*P)).
//
// Replaces chamferdist._C.knn_points_backward as the reference calls it (utils/chamfer.py:206-208):
//   diff = 2 * g[b,i] * (p1[b,i] - p2[b,idx[b,i]]);  grad_p1[b,i] += diff;  grad_p2[b,idx] -= diff.
// The scatter side uses no-return float atomics (SASS REDG); contributions to one target arrive in
// nondeterministic order exactly as upstream's atomicAdd does (tests compare at 1e-5 relative).
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace reart {

template <bool ACCUM_P1>
__global__ void knn1_bwd_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                const int64_t* __restrict__ idx, const float* __restrict__ g, int64_t B, int64_t P1,
                                int64_t P2, float* __restrict__ grad_p1, float* __restrict__ grad_p2) {
    const int64_t total = B * P1;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / P1;
        const int64_t j = idx[e];
        const float g2 = 2.0f * g[e];
        const float* a = p1 + e * 3;
        const float* t = p2 + (b * P2 + j) * 3;
        float* go = grad_p2 + (b * P2 + j) * 3;
        const float dx = g2 * (a[0] - t[0]), dy = g2 * (a[1] - t[1]), dz = g2 * (a[2] - t[2]);
        if (ACCUM_P1) {
            atomicAdd(grad_p1 + e * 3 + 0, dx); atomicAdd(grad_p1 + e * 3 + 1, dy); atomicAdd(grad_p1 + e * 3 + 2, dz);
        } else {
            grad_p1[e * 3 + 0] = dx; grad_p1[e * 3 + 1] = dy; grad_p1[e * 3 + 2] = dz;
        }
        atomicAdd(go + 0, -dx); atomicAdd(go + 1, -dy); atomicAdd(go + 2, -dz);
    }
}

// accumulate == 0: grad_p1 is overwritten and grad_p2 is zeroed first (standalone backward);
// accumulate == 1: both are accumulated into (caller zeroed them; used by the bidirectional backward).
int launch_knn1_bwd(const float* p1, const float* p2, const int64_t* idx, const float* grad_dists, int64_t B,
                    int64_t P1, int64_t P2, float* grad_p1, float* grad_p2, int accumulate, cudaStream_t stream) {
    if (!accumulate) {
        if (B * P2 > 0 && cudaMemsetAsync(grad_p2, 0, sizeof(float) * (size_t)(B * P2 * 3), stream) != cudaSuccess)
            return kErrLaunch;
    }
    const int64_t total = B * P1;
    if (total <= 0) return kOk;
    if (P2 <= 0) {
        if (!accumulate && cudaMemsetAsync(grad_p1, 0, sizeof(float) * (size_t)(total * 3), stream) != cudaSuccess)
            return kErrLaunch;
        return kOk;
    }
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>(ceil_div(total, threads), 148 * 16);
    if (accumulate)
        knn1_bwd_kernel<true><<<blocks, threads, 0, stream>>>(p1, p2, idx, grad_dists, B, P1, P2, grad_p1, grad_p2);
    else
        knn1_bwd_kernel<false><<<blocks, threads, 0, stream>>>(p1, p2, idx, grad_dists, B, P1, P2, grad_p1, grad_p2);
    REART_CHECK_LAUNCH();
    return kOk;
}

// Both directions in one launch.  Elements [0, B*N) are src->tgt pairs, [B*N, B*N+B*M) tgt->src pairs.
// grad_src / grad_tgt must be zero on entry; grad_tgt may be null.
__global__ void chamfer_bidir_bwd_kernel(const float* __restrict__ src, const float* __restrict__ tgt,
                                         const int64_t* __restrict__ i_fwd, const int64_t* __restrict__ i_bwd,
                                         const float* __restrict__ g_fwd, const float* __restrict__ g_bwd, int64_t B,
                                         int64_t N, int64_t M, float* __restrict__ grad_src,
                                         float* __restrict__ grad_tgt) {
    const int64_t nf = B * N, total = nf + B * M;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        if (e < nf) {
            const int64_t b = e / N, j = i_fwd[e];
            const float g2 = 2.0f * g_fwd[e];
            const float* a = src + e * 3;
            const float* t = tgt + (b * M + j) * 3;
            const float dx = g2 * (a[0] - t[0]), dy = g2 * (a[1] - t[1]), dz = g2 * (a[2] - t[2]);
            atomicAdd(grad_src + e * 3 + 0, dx); atomicAdd(grad_src + e * 3 + 1, dy); atomicAdd(grad_src + e * 3 + 2, dz);
            if (grad_tgt) {
                float* go = grad_tgt + (b * M + j) * 3;
                atomicAdd(go + 0, -dx); atomicAdd(go + 1, -dy); atomicAdd(go + 2, -dz);
            }
        } else {
            const int64_t f = e - nf;
            const int64_t b = f / M, j = i_bwd[f];
            const float g2 = 2.0f * g_bwd[f];
            const float* a = tgt + f * 3;
            const float* t = src + (b * N + j) * 3;
            const float dx = g2 * (a[0] - t[0]), dy = g2 * (a[1] - t[1]), dz = g2 * (a[2] - t[2]);
            if (grad_tgt) {
                atomicAdd(grad_tgt + f * 3 + 0, dx); atomicAdd(grad_tgt + f * 3 + 1, dy); atomicAdd(grad_tgt + f * 3 + 2, dz);
            }
            float* go = grad_src + (b * N + j) * 3;
            atomicAdd(go + 0, -dx); atomicAdd(go + 1, -dy); atomicAdd(go + 2, -dz);
        }
    }
}

int launch_chamfer_bidir_bwd(const float* src, const float* tgt, const int64_t* i_fwd, const int64_t* i_bwd,
                             const float* g_fwd, const float* g_bwd, int64_t B, int64_t N, int64_t M, float* grad_src,
                             float* grad_tgt, cudaStream_t stream) {
    if (B * N > 0 && cudaMemsetAsync(grad_src, 0, sizeof(float) * (size_t)(B * N * 3), stream) != cudaSuccess)
        return kErrLaunch;
    if (grad_tgt && B * M > 0 && cudaMemsetAsync(grad_tgt, 0, sizeof(float) * (size_t)(B * M * 3), stream) != cudaSuccess)
        return kErrLaunch;
    const int64_t total = B * (N + M);
    if (total <= 0 || N <= 0 || M <= 0) return kOk;
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>(ceil_div(total, threads), 148 * 16);
    chamfer_bidir_bwd_kernel<<<blocks, threads, 0, stream>>>(src, tgt, i_fwd, i_bwd, g_fwd, g_bwd, B, N, M, grad_src,
                                                             grad_tgt);
    REART_CHECK_LAUNCH();
    return kOk;
}

}  // namespace reart
