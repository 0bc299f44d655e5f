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
                                                               unsigned char* __restrict__ out_mask, int squared) {
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
        for (int a = 0; a < K; ++a) { out_dist[o * K + a] = squared ? bd[a] : sqrtf(bd[a]); out_idx[o * K + a] = bi[a]; }
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

// ----------------------------------------------------------------------------- register-blocked variant
// Same results as knn_topk_kernel, built like the K=1 search: each thread owns RQ queries, the CTA transposes a
// tile of reference points into groups of four [x4|y4|z4] in shared memory, distances come from packed f32x2 math,
// and the sorted top-K insertion only runs when the smallest of four fresh distances beats the query's current
// K-th best (after the first few tiles that is rare).  Candidates are offered in increasing index order with a
// strict <, so equal distances keep the lowest index first, exactly like the one-query-per-thread kernel.
// (Round 2 tried a two-phase form -- branch-free sweep recording one trigger bit per group, insertion afterwards from the
// bit masks -- and measured it SLOWER: 1.58 ms vs 1.22 ms on the 64-pair flow workload; the immediate form stays.)
constexpr int kTkThreads = 128;
constexpr int kTkRQ = 4;
constexpr int kTkTile = 512;                                  // reference points per shared-memory tile

template <int K, bool BLEND>
__global__ void __launch_bounds__(kTkThreads) knn_topk_tiled_kernel(const float* __restrict__ ref,
                                                                    const float* __restrict__ query,
                                                                    const float* __restrict__ flow,
                                                                    const int64_t* __restrict__ ref_off, int n_dense,
                                                                    int m, float* __restrict__ out_dist,
                                                                    int64_t* __restrict__ out_idx,
                                                                    float* __restrict__ out_blend,
                                                                    unsigned char* __restrict__ out_mask, int squared) {
    __shared__ __align__(16) float tile[kTkTile * 3];          // groups of 4: [x0..x3|y0..y3|z0..z3]
    const int b = blockIdx.y;
    const int64_t r0 = ref_off ? ref_off[b] : (int64_t)b * n_dense;
    const int n = ref_off ? (int)(ref_off[b + 1] - ref_off[b]) : n_dense;
    const float* __restrict__ rp = ref + r0 * 3;
    const int qbase = blockIdx.x * (kTkThreads * kTkRQ);
    u64 QX[kTkRQ], QY[kTkRQ], QZ[kTkRQ];
    float bd[kTkRQ][K];
    int bi[kTkRQ][K];
#pragma unroll
    for (int r = 0; r < kTkRQ; ++r) {
        const int i = qbase + r * kTkThreads + threadIdx.x;
        float x = 0.f, y = 0.f, z = 0.f;
        if (i < m) { const float* q = query + ((int64_t)b * m + i) * 3; x = q[0]; y = q[1]; z = q[2]; }
        QX[r] = pack2(x, x); QY[r] = pack2(y, y); QZ[r] = pack2(z, z);
#pragma unroll
        for (int a = 0; a < K; ++a) { bd[r][a] = INFINITY; bi[r][a] = 0; }
    }
    for (int base = 0; base < n; base += kTkTile) {
        const int cnt = min(kTkTile, n - base);
        const int cnt4 = (cnt + 3) & ~3;
        __syncthreads();
        for (int e = threadIdx.x; e < cnt4; e += blockDim.x) {
            float x = INFINITY, y = INFINITY, z = INFINITY;       // padding never beats a finite distance
            if (e < cnt) { const float* s = rp + (int64_t)(base + e) * 3; x = s[0]; y = s[1]; z = s[2]; }
            float* g = tile + (e >> 2) * 12 + (e & 3);
            g[0] = x; g[4] = y; g[8] = z;
        }
        __syncthreads();
        const float4* __restrict__ t4 = reinterpret_cast<const float4*>(tile);
        for (int g = 0; g < cnt4 / 4; ++g) {
            const float4 X = t4[3 * g], Y = t4[3 * g + 1], Z = t4[3 * g + 2];
            const u64 X01 = pack2(X.x, X.y), X23 = pack2(X.z, X.w);
            const u64 Y01 = pack2(Y.x, Y.y), Y23 = pack2(Y.z, Y.w);
            const u64 Z01 = pack2(Z.x, Z.y), Z23 = pack2(Z.z, Z.w);
#pragma unroll
            for (int r = 0; r < kTkRQ; ++r) {
                float a0, a1, a2, a3;
                unpack2(sqdist_pair(QX[r], QY[r], QZ[r], X01, Y01, Z01), a0, a1);
                unpack2(sqdist_pair(QX[r], QY[r], QZ[r], X23, Y23, Z23), a2, a3);
                if (fminf(min3(a0, a1, a2), a3) < bd[r][K - 1]) {
                    const int j = base + 4 * g;
                    topk_insert<K>(bd[r], bi[r], a0, j);
                    topk_insert<K>(bd[r], bi[r], a1, j + 1);
                    topk_insert<K>(bd[r], bi[r], a2, j + 2);
                    topk_insert<K>(bd[r], bi[r], a3, j + 3);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kTkRQ; ++r) {
        const int i = qbase + r * kTkThreads + threadIdx.x;
        if (i >= m) continue;
        const int64_t o = (int64_t)b * m + i;
        if (!BLEND) {
#pragma unroll
            for (int a = 0; a < K; ++a) { out_dist[o * K + a] = squared ? bd[r][a] : sqrtf(bd[r][a]); out_idx[o * K + a] = bi[r][a]; }
        } else {
            float w[K], ws = 0.f, mind = INFINITY, maxf = -INFINITY;
            float fx[K], fy[K], fz[K];
#pragma unroll
            for (int a = 0; a < K; ++a) {
                float d = sqrtf(bd[r][a]);
                if (d < 1e-10f) d = 1e-10f;
                w[a] = 1.0f / d;
                ws += w[a];
                mind = fminf(mind, d);
                const float* f = flow + (r0 + bi[r][a]) * 3;
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
}

template <int K>
static int launch_knn_k(const float* ref, const float* query, int64_t B, int64_t n, int64_t m, float* dist,
                        int64_t* idx, int squared, cudaStream_t stream) {
    if (K <= 4 && m >= 2048) {                               // register-blocked packed-math variant for real workloads
        dim3 tg((unsigned)ceil_div(m, kTkThreads * kTkRQ), (unsigned)B);
        knn_topk_tiled_kernel<(K <= 4 ? K : 1), false><<<tg, kTkThreads, 0, stream>>>(ref, query, nullptr, nullptr, (int)n,
                                                                                     (int)m, dist, idx, nullptr, nullptr, squared);
        REART_CHECK_LAUNCH();
        return kOk;
    }
    dim3 grid((unsigned)ceil_div(m, kKnnThreads), (unsigned)B);
    knn_topk_kernel<K, false><<<grid, kKnnThreads, 0, stream>>>(ref, query, nullptr, nullptr, (int)n, (int)m, dist, idx,
                                                                nullptr, nullptr, squared);
    REART_CHECK_LAUNCH();
    return kOk;
}

int launch_knn(const float* ref, const float* query, int64_t B, int64_t n, int64_t m, int k, float* dist, int64_t* idx,
               int squared, cudaStream_t stream) {
    if (B <= 0 || m <= 0) return kOk;
    if (B > 65535) return kErrUnsupported;
    switch (k) {
        case 1: return launch_knn_k<1>(ref, query, B, n, m, dist, idx, squared, stream);
        case 2: return launch_knn_k<2>(ref, query, B, n, m, dist, idx, squared, stream);
        case 3: return launch_knn_k<3>(ref, query, B, n, m, dist, idx, squared, stream);
        case 4: return launch_knn_k<4>(ref, query, B, n, m, dist, idx, squared, stream);
        case 5: return launch_knn_k<5>(ref, query, B, n, m, dist, idx, squared, stream);
        case 6: return launch_knn_k<6>(ref, query, B, n, m, dist, idx, squared, stream);
        case 7: return launch_knn_k<7>(ref, query, B, n, m, dist, idx, squared, stream);
        case 8: return launch_knn_k<8>(ref, query, B, n, m, dist, idx, squared, stream);
        default: return kErrUnsupported;
    }
}

int launch_knn3_blend(const float* query, const float* ref_cat, const float* flow_cat, const int64_t* ref_off,
                      int64_t T, int64_t m, float* blended, unsigned char* mask, cudaStream_t stream) {
    if (T <= 0 || m <= 0) return kOk;
    if (T > 65535) return kErrUnsupported;
    if (m >= 2048) {
        dim3 tg((unsigned)ceil_div(m, kTkThreads * kTkRQ), (unsigned)T);
        knn_topk_tiled_kernel<3, true><<<tg, kTkThreads, 0, stream>>>(ref_cat, query, flow_cat, ref_off, 0, (int)m, nullptr,
                                                                      nullptr, blended, mask, 0);
        REART_CHECK_LAUNCH();
        return kOk;
    }
    dim3 grid((unsigned)ceil_div(m, kKnnThreads), (unsigned)T);
    knn_topk_kernel<3, true><<<grid, kKnnThreads, 0, stream>>>(ref_cat, query, flow_cat, ref_off, 0, (int)m, nullptr,
                                                               nullptr, blended, mask, 0);
    REART_CHECK_LAUNCH();
    return kOk;
}

}  // namespace reart
