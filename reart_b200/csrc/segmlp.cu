// segmlp.cu -- the small per-iteration pieces around the skinning weights, fused so that the FIXED cost of an
// optimisation step (which does not shrink when frames are sharded over GPUs) is a handful of launches:
//   * seg MLP  logits = W2 relu(W0 x + b0)   (networks/blocks.py:99-118 as built at networks/model.py:19:
//     Conv1d(3,H,1,bias) -> ReLU -> Conv1d(H,P,1,no bias); a k=1 Conv1d is a per-point matmul), forward + backward;
//   * straight-through gumbel-softmax (networks/model.py:44, F.gumbel_softmax(seg, tau, hard=True)) forward +
//     backward given torch-drawn Exponential(1) noise (the RNG stays in torch):
//       y = softmax((logits - log e) / tau);  W = (onehot(argmax y) - y) + y   (ones are 1 +- 1 ulp, zeros exact, Q5)
//       dL/dlogits = y * (g - sum_p g_p y_p) / tau.
// Everything here is O(N (H + P)) and launch-latency bound.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace reart {

constexpr int kMlpThreads = 256;

// ----------------------------------------------------------------------------- seg MLP forward
// W0|b0 and W2^T staged in shared memory; kMlpLanes lanes per point (common.cuh segmlp_point)
template <int PMAX>
__global__ void __launch_bounds__(kMlpThreads) segmlp_fwd_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ w0,
                                                                 const float* __restrict__ b0,
                                                                 const float* __restrict__ w2, int N, int H, int P,
                                                                 float* __restrict__ logits) {
    extern __shared__ __align__(16) float sm[];
    float* s0 = sm;                       // [H][4]: w0 row + bias
    float* s2 = sm + H * 4;               // [H][PMAX + 4]: W2 transposed, zero padded
    segmlp_stage<PMAX>(w0, b0, w2, H, P, s0, s2);
    __syncthreads();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = gid / kMlpLanes, part = gid % kMlpLanes;
    const bool real = n < N;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (real) { px = x[3 * n]; py = x[3 * n + 1]; pz = x[3 * n + 2]; }
    float acc[PMAX];
    segmlp_point<PMAX>(s0, s2, H, part, px, py, pz, acc);
    if (real) {
#pragma unroll
        for (int p = 0; p < PMAX; ++p)
            if (p < P && (p % kMlpLanes) == part) logits[(int64_t)n * P + p] = acc[p];
    }
}

// ----------------------------------------------------------------------------- seg MLP backward
// thread = hidden unit k, block = chunk of points staged in shared memory; all sums over the chunk stay in
// registers, one atomic per output element per block.  gw0 [H,3], gb0 [H], gw2 [P,H] must be zero on entry.
constexpr int kMlpBwdChunk = 64;

template <int PMAX>
__global__ void segmlp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w0,
                                  const float* __restrict__ b0, const float* __restrict__ w2,
                                  const float* __restrict__ glogits, int N, int H, int P, float* __restrict__ gw0,
                                  float* __restrict__ gb0, float* __restrict__ gw2) {
    extern __shared__ __align__(16) float sm[];
    float* sx = sm;                                   // [chunk][4]
    float* sg = sm + kMlpBwdChunk * 4;                // [chunk][PMAX]
    const int n0 = blockIdx.x * kMlpBwdChunk;
    const int cnt = min(kMlpBwdChunk, N - n0);
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
        sx[4 * e] = x[3 * (n0 + e)]; sx[4 * e + 1] = x[3 * (n0 + e) + 1]; sx[4 * e + 2] = x[3 * (n0 + e) + 2]; sx[4 * e + 3] = 0.f;
    }
    for (int e = threadIdx.x; e < cnt * PMAX; e += blockDim.x) {
        const int i = e / PMAX, p = e - i * PMAX;
        sg[e] = p < P ? glogits[(int64_t)(n0 + i) * P + p] : 0.f;
    }
    __syncthreads();
    const int k = threadIdx.x;
    if (k >= H) return;
    const float wx = w0[3 * k], wy = w0[3 * k + 1], wz = w0[3 * k + 2], bb = b0[k];
    float w2k[PMAX], a2[PMAX];
#pragma unroll
    for (int p = 0; p < PMAX; ++p) { w2k[p] = p < P ? w2[p * H + k] : 0.f; a2[p] = 0.f; }
    float ax = 0.f, ay = 0.f, az = 0.f, ab = 0.f;
    for (int i = 0; i < cnt; ++i) {
        const float4 pt = reinterpret_cast<const float4*>(sx)[i];
        const float pre = wx * pt.x + wy * pt.y + wz * pt.z + bb;
        const float h = fmaxf(pre, 0.f);
        const float4* g4 = reinterpret_cast<const float4*>(sg + i * PMAX);
        float gh = 0.f;
#pragma unroll
        for (int q = 0; q < PMAX / 4; ++q) {
            const float4 g = g4[q];
            a2[4 * q] += g.x * h; a2[4 * q + 1] += g.y * h; a2[4 * q + 2] += g.z * h; a2[4 * q + 3] += g.w * h;
            gh += g.x * w2k[4 * q] + g.y * w2k[4 * q + 1] + g.z * w2k[4 * q + 2] + g.w * w2k[4 * q + 3];
        }
        if (pre > 0.f) { ax += gh * pt.x; ay += gh * pt.y; az += gh * pt.z; ab += gh; }
    }
    atomicAdd(gw0 + 3 * k, ax); atomicAdd(gw0 + 3 * k + 1, ay); atomicAdd(gw0 + 3 * k + 2, az);
    atomicAdd(gb0 + k, ab);
#pragma unroll
    for (int p = 0; p < PMAX; ++p)
        if (p < P) atomicAdd(gw2 + p * H + k, a2[p]);
}

template <int PMAX>
static int launch_segmlp_p(const float* x, const float* w0, const float* b0, const float* w2, const float* glogits,
                           int64_t N, int64_t H, int64_t P, float* logits, float* gw0, float* gb0, float* gw2,
                           cudaStream_t stream) {
    if (logits) {
        const size_t smem = (size_t)H * (4 + PMAX + 4) * sizeof(float);
        if (smem > 48 * 1024) return kErrUnsupported;
        segmlp_fwd_kernel<PMAX><<<(unsigned)ceil_div(kMlpLanes * N, kMlpThreads), kMlpThreads, smem, stream>>>(
            x, w0, b0, w2, (int)N, (int)H, (int)P, logits);
        REART_CHECK_LAUNCH();
        return kOk;
    }
    if (gb0 == gw0 + H * 3 && gw2 == gb0 + H) {                   // one flat [gw0 | gb0 | gw2] bucket: one memset
        if (cudaMemsetAsync(gw0, 0, sizeof(float) * (size_t)H * (4 + P), stream) != cudaSuccess) return kErrLaunch;
    } else {
        if (cudaMemsetAsync(gw0, 0, sizeof(float) * (size_t)H * 3, stream) != cudaSuccess) return kErrLaunch;
        if (cudaMemsetAsync(gb0, 0, sizeof(float) * (size_t)H, stream) != cudaSuccess) return kErrLaunch;
        if (cudaMemsetAsync(gw2, 0, sizeof(float) * (size_t)H * P, stream) != cudaSuccess) return kErrLaunch;
    }
    const size_t smem = (size_t)kMlpBwdChunk * (4 + PMAX) * sizeof(float);
    const int threads = (int)round_up(H, 32);
    segmlp_bwd_kernel<PMAX><<<(unsigned)ceil_div(N, kMlpBwdChunk), threads, smem, stream>>>(
        x, w0, b0, w2, glogits, (int)N, (int)H, (int)P, gw0, gb0, gw2);
    REART_CHECK_LAUNCH();
    return kOk;
}

int launch_segmlp(const float* x, const float* w0, const float* b0, const float* w2, const float* glogits, int64_t N,
                  int64_t H, int64_t P, float* logits, float* gw0, float* gb0, float* gw2, cudaStream_t stream) {
    if (N <= 0) return kOk;
    if (H <= 0 || H > 1024 || P <= 0 || P > 32) return kErrUnsupported;
    if (P <= 8) return launch_segmlp_p<8>(x, w0, b0, w2, glogits, N, H, P, logits, gw0, gb0, gw2, stream);
    if (P <= 16) return launch_segmlp_p<16>(x, w0, b0, w2, glogits, N, H, P, logits, gw0, gb0, gw2, stream);
    return launch_segmlp_p<32>(x, w0, b0, w2, glogits, N, H, P, logits, gw0, gb0, gw2, stream);
}

// ----------------------------------------------------------------------------- gumbel softmax, straight-through
template <int PMAX, bool BWD>
__global__ void gumbel_st_kernel(const float* __restrict__ logits, const float* __restrict__ expo,
                                 const float* __restrict__ tau_ptr, const float* __restrict__ gW, int N, int P,
                                 float* __restrict__ W, float* __restrict__ ysoft, float* __restrict__ glogits) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float inv_tau = 1.0f / *tau_ptr;
    if (!BWD) {
        float z[PMAX];
        float mx = -INFINITY;
#pragma unroll
        for (int p = 0; p < PMAX; ++p) {
            if (p < P) {
                z[p] = (logits[(int64_t)n * P + p] - logf(expo[(int64_t)n * P + p])) * inv_tau;
                mx = fmaxf(mx, z[p]);
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int p = 0; p < PMAX; ++p)
            if (p < P) { z[p] = expf(z[p] - mx); sum += z[p]; }
        int hot = 0;
        float best = -1.f;
#pragma unroll
        for (int p = 0; p < PMAX; ++p)
            if (p < P) { z[p] = z[p] / sum; if (z[p] > best) { best = z[p]; hot = p; } }
#pragma unroll
        for (int p = 0; p < PMAX; ++p)
            if (p < P) {
                const float y = z[p];
                ysoft[(int64_t)n * P + p] = y;
                W[(int64_t)n * P + p] = ((p == hot ? 1.0f : 0.0f) - y) + y;
            }
    } else {
        float dot = 0.f;
#pragma unroll
        for (int p = 0; p < PMAX; ++p)
            if (p < P) dot += gW[(int64_t)n * P + p] * ysoft[(int64_t)n * P + p];
#pragma unroll
        for (int p = 0; p < PMAX; ++p)
            if (p < P) {
                const float y = ysoft[(int64_t)n * P + p];
                glogits[(int64_t)n * P + p] = y * (gW[(int64_t)n * P + p] - dot) * inv_tau;
            }
    }
}

int launch_gumbel_st(const float* logits, const float* expo, const float* tau, const float* gW, int64_t N, int64_t P,
                     float* W, float* ysoft, float* glogits, cudaStream_t stream) {
    if (N <= 0) return kOk;
    if (P <= 0 || P > 32) return kErrUnsupported;
    const unsigned blocks = (unsigned)ceil_div(N, 128);
    const bool bwd = glogits != nullptr;
#define REART_GST(PM)                                                                                                   \
    do {                                                                                                                \
        if (bwd) gumbel_st_kernel<PM, true><<<blocks, 128, 0, stream>>>(logits, expo, tau, gW, (int)N, (int)P, W, ysoft, glogits); \
        else gumbel_st_kernel<PM, false><<<blocks, 128, 0, stream>>>(logits, expo, tau, gW, (int)N, (int)P, W, ysoft, glogits);    \
    } while (0)
    if (P <= 8) REART_GST(8);
    else if (P <= 16) REART_GST(16);
    else REART_GST(32);
#undef REART_GST
    REART_CHECK_LAUNCH();
    return kOk;
}

}  // namespace reart
