// knnk.cu -- small-k brute-force k-NN (k <= 8) with Euclidean distances, and the fused flow blend.
//
// Replaces knn_cuda.KNN(k, transpose_mode)(ref, query) (third-party KNN_CUDA 0.2; call sites
// utils/flow_utils.py:158, utils/model_utils.py:42) and the torch tail of blend_anchor_motion
// (utils/flow_utils.py:159-167): clamp 1e-10, inverse-distance weights, blended flow, validity mask.
// Ordering: ascending squared distance, ties -> lowest reference index (oracle_knn).
// One thread per query; reference points are staged through shared memory in SoA tiles.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace reart {

constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 1024;

template <int K>
__device__ __forceinline__ void topk_insert(float (&bd)[K], int (&bi)[K], float d, int j) {
    if (d < bd[K - 1]) {
        bd[K - 1] = d; bi[K - 1] = j;
#pragma unroll
        for (int a = K - 1; a > 0; --a) {
            if (bd[a] < bd[a - 1]) {                          // strict: equal distances keep index order
                const float td = bd[a]; bd[a] = bd[a - 1]; bd[a - 1] = td;
                const int ti = bi[a]; bi[a] = bi[a - 1]; bi[a - 1] = ti;
            }
        }
    }
}

// grid (ceil(m/128), B).  ref_off == null: ref is [B,n,3] dense; else batch b uses rows [ref_off[b], ref_off[b+1]).
template <int K, bool BLEND>
__global__ void __launch_bounds__(kKnnThreads) knn_topk_kernel(const float* __restrict__ ref,
                                                               const float* __restrict__ query,
                                                               const float* __restrict__ flow,
                                                               const int64_t* __restrict__ ref_off, int n_dense, int m,
                                                               float* __restrict__ out_dist,
                                                               int64_t* __restrict__ out_idx,
                                                               float* __restrict__ out_blend,
                                                               unsigned char* __restrict__ out_mask) {
    __shared__ float sx[kKnnTile], sy[kKnnTile], sz[kKnnTile];
    const int b = blockIdx.y;
    const int64_t r0 = ref_off ? ref_off[b] : (int64_t)b * n_dense;
    const int n = ref_off ? (int)(ref_off[b + 1] - ref_off[b]) : n_dense;
    const float* __restrict__ rp = ref + r0 * 3;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (valid) {
        const float* q = query + ((int64_t)b * m + i) * 3;
        qx = q[0]; qy = q[1]; qz = q[2];
    }
    float bd[K];
    int bi[K];
#pragma unroll
    for (int a = 0; a < K; ++a) { bd[a] = INFINITY; bi[a] = 0; }
    for (int base = 0; base < n; base += kKnnTile) {
        const int cnt = min(kKnnTile, n - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            sx[e] = rp[(int64_t)(base + e) * 3]; sy[e] = rp[(int64_t)(base + e) * 3 + 1]; sz[e] = rp[(int64_t)(base + e) * 3 + 2];
        }
        __syncthreads();
#pragma unroll 4
        for (int e = 0; e < cnt; ++e) {
            const float d = sqdist_scalar(qx, qy, qz, sx[e], sy[e], sz[e]);
            topk_insert<K>(bd, bi, d, base + e);
        }
    }
    if (!valid) return;
    const int64_t o = (int64_t)b * m + i;
    if (!BLEND) {
#pragma unroll
        for (int a = 0; a < K; ++a) { out_dist[o * K + a] = sqrtf(bd[a]); out_idx[o * K + a] = bi[a]; }
    } else {
        // utils/flow_utils.py:159-167
        float w[K], ws = 0.f, mind = INFINITY, maxf = -INFINITY;
        float fx[K], fy[K], fz[K];
#pragma unroll
        for (int a = 0; a < K; ++a) {
            float d = sqrtf(bd[a]);
            if (d < 1e-10f) d = 1e-10f;
            w[a] = 1.0f / d;
            ws += w[a];
            mind = fminf(mind, d);
            const float* f = flow + (r0 + bi[a]) * 3;
            fx[a] = f[0]; fy[a] = f[1]; fz[a] = f[2];
            maxf = fmaxf(maxf, fx[a] * fx[a] + fy[a] * fy[a] + fz[a] * fz[a]);
        }
        float ox = 0.f, oy = 0.f, oz = 0.f;
#pragma unroll
        for (int a = 0; a < K; ++a) {
            const float wn = w[a] / ws;
            ox += fx[a] * wn; oy += fy[a] * wn; oz += fz[a] * wn;
        }
        out_blend[o * 3] = ox; out_blend[o * 3 + 1] = oy; out_blend[o * 3 + 2] = oz;
        if (out_mask) out_mask[o] = (mind <= maxf) || (mind <= 0.05f);
    }
}

template <int K>
static int launch_knn_k(const float* ref, const float* query, int64_t B, int64_t n, int64_t m, float* dist,
                        int64_t* idx, cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div(m, kKnnThreads), (unsigned)B);
    knn_topk_kernel<K, false><<<grid, kKnnThreads, 0, stream>>>(ref, query, nullptr, nullptr, (int)n, (int)m, dist, idx,
                                                                nullptr, nullptr);
    REART_CHECK_LAUNCH();
    return kOk;
}

int launch_knn(const float* ref, const float* query, int64_t B, int64_t n, int64_t m, int k, float* dist, int64_t* idx,
               cudaStream_t stream) {
    if (B <= 0 || m <= 0) return kOk;
    if (B > 65535) return kErrUnsupported;
    switch (k) {
        case 1: return launch_knn_k<1>(ref, query, B, n, m, dist, idx, stream);
        case 2: return launch_knn_k<2>(ref, query, B, n, m, dist, idx, stream);
        case 3: return launch_knn_k<3>(ref, query, B, n, m, dist, idx, stream);
        case 4: return launch_knn_k<4>(ref, query, B, n, m, dist, idx, stream);
        case 5: return launch_knn_k<5>(ref, query, B, n, m, dist, idx, stream);
        case 6: return launch_knn_k<6>(ref, query, B, n, m, dist, idx, stream);
        case 7: return launch_knn_k<7>(ref, query, B, n, m, dist, idx, stream);
        case 8: return launch_knn_k<8>(ref, query, B, n, m, dist, idx, stream);
        default: return kErrUnsupported;
    }
}

int launch_knn3_blend(const float* query, const float* ref_cat, const float* flow_cat, const int64_t* ref_off,
                      int64_t T, int64_t m, float* blended, unsigned char* mask, cudaStream_t stream) {
    if (T <= 0 || m <= 0) return kOk;
    if (T > 65535) return kErrUnsupported;
    dim3 grid((unsigned)ceil_div(m, kKnnThreads), (unsigned)T);
    knn_topk_kernel<3, true><<<grid, kKnnThreads, 0, stream>>>(ref_cat, query, flow_cat, ref_off, 0, (int)m, nullptr,
                                                               nullptr, blended, mask);
    REART_CHECK_LAUNCH();
    return kOk;
}

}  // namespace reart
