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

// ----------------------------------------------------------------------------- windowed exact variant (flow blend)
// The brute force above evaluates every (query, reference) pair; its ceiling is the FP32 pipe and the divergent
// insertions keep it at ~0.3 of that.  The flow loss searches the SAME reference sets every iteration, so they are
// sorted along x once (flow_refs_sort_kernel) and each call only has to look at the x-slab that can still hold one
// of the K nearest:
//   * knnw_order_queries_kernel  buckets the queries of a pair by (x slab, y cell, z cell) (counting sort) so that the
//     128 queries of a warp span a thin x-range and are neighbours in space;
//   * knn_window_kernel          a CTA (4 warps x 128 queries) starts at the 512-reference tiles that overlap its own
//     x-range, then walks outwards to the left and to the right; a warp skips a 64-reference sub-tile, and the CTA
//     stops walking, when the squared x-gap to the nearest reference of the sub-tile / tile is STRICTLY above the
//     largest current K-th best of its queries.  d = fma(dz,dz,fma(dy,dy,dx*dx)) >= dx*dx and float subtraction and
//     multiplication are monotonic, so no skipped reference can enter any top-K, ties included: the results are
//     bit-identical to the brute force (same distance arithmetic; candidates arrive in x order, so the insertion
//     orders (distance, original index) lexicographically instead of relying on arrival order).
// Sorted layout per pair: groups of four references [x4 | y4 | z4 | original index x4] (64 B), +INF / INT_MAX padded;
// pair b starts at group sorted_off[b].
constexpr int kWinSortMax = 16384;                            // references per pair the one-CTA sort handles
constexpr int kWinSortThreads = 1024;
constexpr int kWinThreads = 128;
constexpr int kWinRQ = 4;
constexpr int kWinTileGroups = 128;                           // 512 references per shared-memory tile
constexpr int kNoIndex = 0x7fffffff;

__global__ void __launch_bounds__(kWinSortThreads) flow_refs_sort_kernel(const float* __restrict__ ref,
                                                                         const int64_t* __restrict__ ref_off, int n_pad,
                                                                         float* __restrict__ sorted,
                                                                         int64_t* __restrict__ sorted_off) {
    extern __shared__ __align__(16) unsigned char sort_raw[];
    volatile u64* keys = reinterpret_cast<volatile u64*>(sort_raw);
    const int b = blockIdx.x, tid = threadIdx.x;
    const int64_t r0 = ref_off[b];
    const int n = (int)min((int64_t)n_pad, ref_off[b + 1] - r0);
    const int64_t g0 = ((r0 & ~(int64_t)3) + 8 * (int64_t)b) >> 2;       // disjoint, 4-aligned slots for every pair
    if (tid == 0) sorted_off[b] = g0;
    for (int e = tid; e < n_pad; e += kWinSortThreads)
        keys[e] = e < n ? ((u64)orderable_bits(ref[(r0 + e) * 3]) << 32) | (unsigned)e : ~0ull;
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int e = tid; e < n_pad; e += kWinSortThreads) {
                const int partner = e ^ j;
                if (partner > e) {
                    const u64 a = keys[e], c = keys[partner];
                    const bool up = (e & k) == 0;
                    if ((a > c) == up) { keys[e] = c; keys[partner] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int slots = ((n + 3) >> 2) << 2;
    float* out = sorted + g0 * 16;
    for (int e = tid; e < slots; e += kWinSortThreads) {
        float x = INFINITY, y = INFINITY, z = INFINITY;
        int idx = kNoIndex;
        if (e < n) {
            idx = (int)(keys[e] & 0xffffffffu);
            const float* s = ref + (r0 + idx) * 3;
            x = s[0]; y = s[1]; z = s[2];
        }
        float* g = out + (int64_t)(e >> 2) * 16 + (e & 3);
        g[0] = x; g[4] = y; g[8] = z; g[12] = __int_as_float(idx);
    }
}

// qperm[b][pos] = query index, positions ordered by bucket = (x slab, y cell, z cell): x is the major key, so the 128
// queries of a warp span a thin x-range (narrow slab windows), and inside a slab neighbours in the order are neighbours
// in y,z too, so the lanes of a warp meet their nearest references in the same groups (their insertions coincide instead
// of serialising).  The order inside a bucket is arbitrary: the order only decides which queries share a warp, never a result.
__global__ void __launch_bounds__(kWinSortThreads) knnw_order_queries_kernel(const float* __restrict__ query, int m,
                                                                             int* __restrict__ qperm, int x_bits,
                                                                             int yz_bits) {
    extern __shared__ int hist[];
    const int buckets = 1 << (x_bits + 2 * yz_bits);
    __shared__ float s_lo[3][32], s_hi[3][32];
    __shared__ int s_warp[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ q = query + (int64_t)b * m * 3;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < m; i += kWinSortThreads) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float v = q[(int64_t)i * 3 + k]; lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
        if (lane == 0) { s_lo[k][warp] = lo[k]; s_hi[k][warp] = hi[k]; }
    }
    for (int e = tid; e < buckets; e += kWinSortThreads) hist[e] = 0;
    __syncthreads();
    float scale[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = s_lo[k][0]; hi[k] = s_hi[k][0];
        for (int w = 1; w < 32; ++w) { lo[k] = fminf(lo[k], s_lo[k][w]); hi[k] = fmaxf(hi[k], s_hi[k][w]); }
        const int cells = k == 0 ? (1 << x_bits) : (1 << yz_bits);
        scale[k] = hi[k] > lo[k] ? (float)cells / (hi[k] - lo[k]) : 0.f;
    }
    auto bucket = [&](int i) {
        const int cx = min((1 << x_bits) - 1, max(0, (int)((q[(int64_t)i * 3] - lo[0]) * scale[0])));
        const int cy = min((1 << yz_bits) - 1, max(0, (int)((q[(int64_t)i * 3 + 1] - lo[1]) * scale[1])));
        const int cz = min((1 << yz_bits) - 1, max(0, (int)((q[(int64_t)i * 3 + 2] - lo[2]) * scale[2])));
        // boustrophedon in y and z: consecutive buckets stay neighbours in space
        const int yy = (cx & 1) ? (1 << yz_bits) - 1 - cy : cy;
        const int zz = (yy & 1) ? (1 << yz_bits) - 1 - cz : cz;
        return (cx << (2 * yz_bits)) | (yy << yz_bits) | zz;
    };
    for (int i = tid; i < m; i += kWinSortThreads) atomicAdd(&hist[bucket(i)], 1);
    __syncthreads();
    // exclusive scan of the bucket counts: buckets / 1024 consecutive buckets per thread
    const int kPer = buckets / kWinSortThreads;
    int tsum = 0;
    for (int k = 0; k < kPer; ++k) tsum += hist[tid * kPer + k];
    int incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int base = incl - tsum;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    for (int k = 0; k < kPer; ++k) { const int c = hist[tid * kPer + k]; hist[tid * kPer + k] = base; base += c; }   // the bucket's running cursor
    __syncthreads();
    int* __restrict__ out = qperm + (int64_t)b * m;
    for (int i = tid; i < m; i += kWinSortThreads) out[atomicAdd(&hist[bucket(i)], 1)] = i;
}

// sorted insertion by (distance, original index): candidates do not arrive in index order here
template <int K>
__device__ __forceinline__ void topk_insert_lex(float (&bd)[K], int (&bi)[K], float d, int j) {
    if (d < bd[K - 1] || (d == bd[K - 1] && j < bi[K - 1])) {
        bd[K - 1] = d; bi[K - 1] = j;
#pragma unroll
        for (int a = K - 1; a > 0; --a) {
            if (bd[a] < bd[a - 1] || (bd[a] == bd[a - 1] && bi[a] < bi[a - 1])) {
                const float td = bd[a]; bd[a] = bd[a - 1]; bd[a - 1] = td;
                const int ti = bi[a]; bi[a] = bi[a - 1]; bi[a - 1] = ti;
            }
        }
    }
}

__device__ __forceinline__ float sq_gap(float a, float b) { const float g = a - b; return g * g; }

template <int K>
__global__ void __launch_bounds__(kWinThreads) knn_window_blend_kernel(const float* __restrict__ query,
                                                                       const float* __restrict__ sorted,
                                                                       const int64_t* __restrict__ sorted_off,
                                                                       const float* __restrict__ flow,
                                                                       const int64_t* __restrict__ ref_off,
                                                                       const int* __restrict__ qperm, int m,
                                                                       float* __restrict__ out_blend,
                                                                       unsigned char* __restrict__ out_mask, int sub_groups) {
    __shared__ __align__(16) float tile[kWinTileGroups * 16];
    __shared__ float s_lo[4], s_hi[4], s_r2[4];
    __shared__ int s_cnt[2][4];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t r0 = ref_off[b];
    const int n = (int)(ref_off[b + 1] - r0);
    const float* __restrict__ S = sorted + sorted_off[b] * 16;
    const int qbase = blockIdx.x * (kWinThreads * kWinRQ);
    int qi[kWinRQ];
    bool qvalid[kWinRQ];
    u64 QX[kWinRQ], QY[kWinRQ], QZ[kWinRQ];
    float bd[kWinRQ][K];
    int bi[kWinRQ][K];
    float qlo = INFINITY, qhi = -INFINITY;
#pragma unroll
    for (int r = 0; r < kWinRQ; ++r) {
        const int pos = qbase + warp * (32 * kWinRQ) + r * 32 + lane;
        qvalid[r] = pos < m;
        qi[r] = qperm[(int64_t)b * m + min(pos, m - 1)];        // lanes past the end shadow the last query
        const float* q = query + ((int64_t)b * m + qi[r]) * 3;
        const float x = q[0], y = q[1], z = q[2];
        QX[r] = pack2(x, x); QY[r] = pack2(y, y); QZ[r] = pack2(z, z);
        qlo = fminf(qlo, x); qhi = fmaxf(qhi, x);
#pragma unroll
        for (int a = 0; a < K; ++a) { bd[r][a] = INFINITY; bi[r][a] = kNoIndex; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { qlo = fminf(qlo, __shfl_xor_sync(0xffffffffu, qlo, o)); qhi = fmaxf(qhi, __shfl_xor_sync(0xffffffffu, qhi, o)); }
    if (lane == 0) { s_lo[warp] = qlo; s_hi[warp] = qhi; s_r2[warp] = INFINITY; }
    __syncthreads();
    const float clo = fminf(fminf(s_lo[0], s_lo[1]), fminf(s_lo[2], s_lo[3]));
    const float chi = fmaxf(fmaxf(s_hi[0], s_hi[1]), fmaxf(s_hi[2], s_hi[3]));
    // positions of the CTA's x-range in the sorted references: c0 = #{x < clo}, c1 = #{x <= chi}
    int below = 0, upto = 0;
    for (int p = tid; p < n; p += kWinThreads) {
        const float x = S[(int64_t)(p >> 2) * 16 + (p & 3)];
        below += x < clo; upto += x <= chi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { below += __shfl_xor_sync(0xffffffffu, below, o); upto += __shfl_xor_sync(0xffffffffu, upto, o); }
    if (lane == 0) { s_cnt[0][warp] = below; s_cnt[1][warp] = upto; }
    __syncthreads();
    const int c0 = s_cnt[0][0] + s_cnt[0][1] + s_cnt[0][2] + s_cnt[0][3];
    const int c1 = s_cnt[1][0] + s_cnt[1][1] + s_cnt[1][2] + s_cnt[1][3];
    const int ngroups = (n + 3) >> 2;
    const int ntiles = (ngroups + kWinTileGroups - 1) / kWinTileGroups;
    const int refs_per_tile = kWinTileGroups * 4;

    // tiles in the order: those overlapping the CTA's own x-range (phase 0), then outwards to the left (1), then to the
    // right (2); ONE copy of the tile body (a state machine instead of three loops keeps the register arrays in registers)
    if (n > 0) {
        const int ta = min(ntiles - 1, c0 / refs_per_tile);
        const int tb = min(ntiles - 1, max(ta, (c1 - 1) / refs_per_tile));
        int phase = 0, t = ta;
        for (;;) {
            if (phase == 0 && t > tb) { phase = 1; t = ta - 1; }
            if (phase == 1 && t < 0) { phase = 2; t = tb + 1; }
            if (phase == 2 && t >= ntiles) break;
            __syncthreads();                                   // previous tile consumed, s_r2 of all warps visible
            if (phase != 0) {
                const float r2c = fmaxf(fmaxf(s_r2[0], s_r2[1]), fmaxf(s_r2[2], s_r2[3]));
                if (phase == 1) {                              // the tile's largest x is its last reference
                    const int p = t * refs_per_tile + refs_per_tile - 1;
                    const float x_last = S[(int64_t)(p >> 2) * 16 + (p & 3)];
                    if (x_last < clo && sq_gap(clo, x_last) > r2c) { phase = 2; t = tb + 1; continue; }
                } else {                                       // the tile's smallest x is its first reference
                    const float x_first = S[(int64_t)t * kWinTileGroups * 16];
                    if (x_first > chi && sq_gap(chi, x_first) > r2c) break;
                }
            }
            const int gcount = min(kWinTileGroups, ngroups - t * kWinTileGroups);
            const float4* __restrict__ src = reinterpret_cast<const float4*>(S) + (int64_t)t * kWinTileGroups * 4;
            float4* dst = reinterpret_cast<float4*>(tile);
            for (int e = tid; e < gcount * 4; e += kWinThreads) dst[e] = src[e];
            __syncthreads();
            const float4* __restrict__ t4 = reinterpret_cast<const float4*>(tile);
            for (int g0 = 0; g0 < gcount; g0 += sub_groups) {
                const int g1 = min(gcount, g0 + sub_groups);
                float mine = bd[0][K - 1];
#pragma unroll
                for (int r = 1; r < kWinRQ; ++r) mine = fmaxf(mine, bd[r][K - 1]);
                const float r2w = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(mine)));   // distances >= 0
                const float x_first = tile[g0 * 16], x_last = tile[(g1 - 1) * 16 + 3];
                if ((x_last < qlo && sq_gap(qlo, x_last) > r2w) || (x_first > qhi && sq_gap(qhi, x_first) > r2w)) continue;
                for (int g = g0; g < g1; ++g) {
                    const float4 X = t4[4 * g], Y = t4[4 * g + 1], Z = t4[4 * g + 2];
                    const u64 X01 = pack2(X.x, X.y), X23 = pack2(X.z, X.w);
                    const u64 Y01 = pack2(Y.x, Y.y), Y23 = pack2(Y.z, Y.w);
                    const u64 Z01 = pack2(Z.x, Z.y), Z23 = pack2(Z.z, Z.w);
#pragma unroll
                    for (int r = 0; r < kWinRQ; ++r) {
                        float a0, a1, a2, a3;
                        unpack2(sqdist_pair(QX[r], QY[r], QZ[r], X01, Y01, Z01), a0, a1);
                        unpack2(sqdist_pair(QX[r], QY[r], QZ[r], X23, Y23, Z23), a2, a3);
                        if (fminf(min3(a0, a1, a2), a3) <= bd[r][K - 1]) {
                            const int4 J = *reinterpret_cast<const int4*>(&t4[4 * g + 3]);
                            topk_insert_lex<K>(bd[r], bi[r], a0, J.x);
                            topk_insert_lex<K>(bd[r], bi[r], a1, J.y);
                            topk_insert_lex<K>(bd[r], bi[r], a2, J.z);
                            topk_insert_lex<K>(bd[r], bi[r], a3, J.w);
                        }
                    }
                }
            }
            {
                float mine = bd[0][K - 1];
#pragma unroll
                for (int r = 1; r < kWinRQ; ++r) mine = fmaxf(mine, bd[r][K - 1]);
                const float r2w = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(mine)));
                if (lane == 0) s_r2[warp] = r2w;
            }
            t += phase == 1 ? -1 : 1;
        }
    }
#pragma unroll
    for (int r = 0; r < kWinRQ; ++r) {
        if (!qvalid[r]) continue;
        const int64_t o = (int64_t)b * m + qi[r];
        // utils/flow_utils.py:159-167 (same arithmetic and order as the brute-force kernels above)
        float w[K], ws = 0.f, mind = INFINITY, maxf = -INFINITY;
        float fx[K], fy[K], fz[K];
#pragma unroll
        for (int a = 0; a < K; ++a) {
            float d = sqrtf(bd[r][a]);
            if (d < 1e-10f) d = 1e-10f;
            w[a] = 1.0f / d;
            ws += w[a];
            mind = fminf(mind, d);
            const float* f = flow + (r0 + (bi[r][a] == kNoIndex ? 0 : bi[r][a])) * 3;
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

int64_t flow_refs_sorted_floats(int64_t total_refs, int64_t T) { return (total_refs + 8 * T + 8) * 4; }

int launch_flow_refs_sort(const float* ref_cat, const int64_t* ref_off, int64_t T, int64_t max_refs, float* sorted,
                          int64_t* sorted_off, cudaStream_t stream) {
    if (T <= 0) return kOk;
    if (max_refs > kWinSortMax || T > 0x7fffffff) return kErrUnsupported;
    int n_pad = 32;
    while (n_pad < max_refs) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * sizeof(u64);
    static bool attr_done[64] = {};
    int devid = 0;
    cudaGetDevice(&devid);
    if (smem > 48 * 1024 && (devid < 0 || devid >= 64 || !attr_done[devid])) {
        if (cudaFuncSetAttribute(flow_refs_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kWinSortMax * sizeof(u64))) != cudaSuccess)
            return kErrUnsupported;
        if (devid >= 0 && devid < 64) attr_done[devid] = true;
    }
    flow_refs_sort_kernel<<<(unsigned)T, kWinSortThreads, smem, stream>>>(ref_cat, ref_off, n_pad, sorted, sorted_off);
    REART_CHECK_LAUNCH();
    return kOk;
}

int launch_knn3_blend_sorted(const float* query, const float* sorted, const int64_t* sorted_off, const float* flow_cat,
                             const int64_t* ref_off, int64_t T, int64_t m, int* qperm, float* blended,
                             unsigned char* mask, cudaStream_t stream) {
    if (T <= 0 || m <= 0) return kOk;
    if (T > 65535) return kErrUnsupported;
    // measured on B200 (profiles/r02_flow_blend.md): 32 x slabs x 16 x 16 (y,z) cells and 64-reference skip units are the
    // best of a flat landscape (500 us; 6/3/3 bits with 128-reference units 543 us; x only 689 us)
    const int xb = 5, yzb = 4, sub = 16;
    const size_t hsmem = sizeof(int) << (xb + 2 * yzb);
    if (hsmem > 48 * 1024) cudaFuncSetAttribute(knnw_order_queries_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem);
    knnw_order_queries_kernel<<<(unsigned)T, kWinSortThreads, hsmem, stream>>>(query, (int)m, qperm, xb, yzb);
    REART_CHECK_LAUNCH();
    dim3 grid((unsigned)ceil_div(m, kWinThreads * kWinRQ), (unsigned)T);
    knn_window_blend_kernel<3><<<grid, kWinThreads, 0, stream>>>(query, sorted, sorted_off, flow_cat, ref_off, qperm, (int)m,
                                                                 blended, mask, sub);
    REART_CHECK_LAUNCH();
    return kOk;
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
