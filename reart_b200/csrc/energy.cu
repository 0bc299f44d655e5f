// energy.cu -- fused tail of the energy evaluation: from the (min, chunk) keys of the symmetric search
// to the loss and its gradient w.r.t. the skinned cloud, in one pass.
//
// Fuses what the reference does in four steps: index recovery (inside chamferdist._C.knn_points_idx,
// utils/chamfer.py:174), torch.sum of the per-point distances (networks/loss.py:27-28) and the two
// _knn_points.backward calls (utils/chamfer.py:195-209):
//   loss   = sum_i d(src_i, tgt_nn(i)) + sum_j d(tgt_j, src_nn(j))
//   g_src  = gscale * [ 2 (src_i - tgt_nn(i))  -  sum_{j: nn(j)=i} 2 (tgt_j - src_i) ]
// Memory-bound O(B (N+M)); the scatter uses REDG float atomics like upstream.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace reart {

// Two launches: the row pass WRITES the direct gradient of every skinned point (plain coalesced stores: no zero
// fill of g_src, no atomics), the column pass then scatters the reverse-direction terms with float atomics.
template <bool ROWS>
__global__ void __launch_bounds__(256) energy_bwd_kernel(const EnergyParams p) {
    const int64_t total = ROWS ? (int64_t)p.B * p.N : (int64_t)p.B * p.M;
    float local = 0.f;
    const float g2 = 2.0f * p.gscale;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        if (ROWS) {
            const int64_t b = e / p.N;
            const u64 key = p.keys_a[e];
            const float dmin = __uint_as_float((unsigned)(key >> 32));
            const float* a = p.src + e * 3;
            const float ax = a[0], ay = a[1], az = a[2];
            int j = rescan_chunk(p.tgt_packed + b * (int64_t)p.m_pad * 3, (unsigned)(key & 0xffffffffu),
                                 p.row_chunk_pts, p.m_pad, ax, ay, az, dmin);
            if (j >= p.M) j = 0;
            const float* t = p.tgt + (b * p.M + j) * 3;
            p.g_src[e * 3 + 0] = g2 * (ax - t[0]);
            p.g_src[e * 3 + 1] = g2 * (ay - t[1]);
            p.g_src[e * 3 + 2] = g2 * (az - t[2]);
            local += dmin;
            if (p.d_fwd) p.d_fwd[e] = dmin;
            if (p.i_fwd) p.i_fwd[e] = j;
        } else {
            const int64_t f = e;
            const int64_t b = f / p.M;
            const u64 key = p.keys_b[f];
            const float dmin = __uint_as_float((unsigned)(key >> 32));
            const float* a = p.tgt + f * 3;
            const float ax = a[0], ay = a[1], az = a[2];
            int i;
            if (p.src_perm)
                i = rescan_sorted_chunk(p.src_packed + b * (int64_t)p.n_pad * 3, p.src_perm + b * (int64_t)p.n_pad,
                                        (unsigned)(key & 0xffffffffu), p.col_chunk_pts, ax, ay, az, dmin);
            else
                i = rescan_chunk(p.src_packed + b * (int64_t)p.n_pad * 3, (unsigned)(key & 0xffffffffu),
                                 p.col_chunk_pts, p.n_pad, ax, ay, az, dmin);
            if (i >= p.N) i = 0;
            const float* t = p.src + (b * p.N + i) * 3;
            float* go = p.g_src + (b * p.N + i) * 3;
            atomicAdd(go + 0, -g2 * (ax - t[0]));
            atomicAdd(go + 1, -g2 * (ay - t[1]));
            atomicAdd(go + 2, -g2 * (az - t[2]));
            local += dmin;
            if (p.d_bwd) p.d_bwd[f] = dmin;
            if (p.i_bwd) p.i_bwd[f] = i;
        }
    }
    // block reduction of the loss: warp shuffle, then one double atomic per block
    __shared__ float warp_sums[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += (double)warp_sums[w];
        atomicAdd(p.loss, s);
    }
}

// g_src needs NO zero fill (the row pass overwrites it); loss must be zero on entry.
int launch_energy_bwd(const EnergyParams& p, cudaStream_t stream) {
    if (p.B <= 0) return kOk;
    const int64_t rows = (int64_t)p.B * p.N, cols = (int64_t)p.B * p.M;
    if (rows > 0) {
        if (p.M <= 0) {
            if (cudaMemsetAsync(p.g_src, 0, sizeof(float) * (size_t)rows * 3, stream) != cudaSuccess) return kErrLaunch;
        } else {
            const int blocks = (int)std::min<int64_t>(ceil_div(rows, 256), 148 * 16);
            energy_bwd_kernel<true><<<blocks, 256, 0, stream>>>(p);
            REART_CHECK_LAUNCH();
        }
    }
    if (cols > 0 && p.N > 0) {
        const int blocks = (int)std::min<int64_t>(ceil_div(cols, 256), 148 * 16);
        energy_bwd_kernel<false><<<blocks, 256, 0, stream>>>(p);
        REART_CHECK_LAUNCH();
    }
    return kOk;
}

}  // namespace reart
