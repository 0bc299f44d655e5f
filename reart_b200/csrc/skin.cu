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
                                                                int n_pad, int fpb) {
    extern __shared__ float sm_tf[];                          // [frames][P][12]: r00..r22, t0..t2
    const int t0 = blockIdx.y * fpb;
    const int nt = min(fpb, T - t0);
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
    if (P <= 0 || P > 32) return kErrUnsupported;
    const int64_t n_pad = out_packed ? padded_points(N) : N;
    // frames per block: 8 amortises the W-row reads, but few frames (a GPU's shard under strong scaling) need more CTAs
    int fpb = kSkinFramesPerBlock;
    while (fpb > 1 && ceil_div(n_pad, kSkinThreads) * ceil_div(T, fpb) < 2 * 148) fpb /= 2;
    dim3 grid((unsigned)ceil_div(n_pad, kSkinThreads), (unsigned)ceil_div(T, fpb));
    const size_t smem = (size_t)fpb * P * 12 * sizeof(float);
    skin_fwd_kernel<<<grid, kSkinThreads, smem, stream>>>(cano, W, R, tr, (int)T, (int)N, (int)P, out, out_packed,
                                                          (int)n_pad, fpb);
    REART_CHECK_LAUNCH();
    return kOk;
}

// ----------------------------------------------------------------------------- forward, x-sorted packed copy
// Same skinning, but the packed copy is written with every block of kSortBlock (= one column chunk of the symmetric
// search: the 256 consecutive points one warp owns) SORTED BY X, plus perm[t][k] = original offset of the point now at
// sorted position k.  The index recovery of the column side (energy.cu) then binary-searches the x window
// |b.x - a.x| <= sqrt(dmin) instead of testing all 256 candidates (~4 candidates survive on surface clouds).
// The AoS output keeps the original order -- it is what the search and every consumer read.
constexpr int kSortBlock = 256;

__device__ __forceinline__ unsigned orderable_bits(float v) {
    const unsigned u = __float_as_uint(v);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}

__global__ void __launch_bounds__(kSortBlock) skin_fwd_sorted_kernel(const float* __restrict__ cano,
                                                                     const float* __restrict__ W,
                                                                     const float* __restrict__ R,
                                                                     const float* __restrict__ tr, int T, int N, int P,
                                                                     float* __restrict__ out,
                                                                     float* __restrict__ out_packed,
                                                                     unsigned char* __restrict__ perm, int n_pad,
                                                                     int fpb) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    u64* keys = reinterpret_cast<u64*>(sm_raw);                               // [256]
    float* sx = reinterpret_cast<float*>(keys + kSortBlock);                  // [3][256]
    float* sm_tf = sx + 3 * kSortBlock;                                       // [fpb][P][12]
    const int t0 = blockIdx.y * fpb;
    const int nt = min(fpb, T - t0);
    for (int e = threadIdx.x; e < nt * P * 12; e += blockDim.x) {
        const int f = e / (P * 12), rem = e - f * P * 12, p = rem / 12, k = rem - p * 12;
        const int64_t tp = (int64_t)(t0 + f) * P + p;
        sm_tf[e] = (k < 9) ? R[tp * 9 + k] : tr[tp * 3 + (k - 9)];
    }
    __syncthreads();
    const int i = threadIdx.x;
    const int base = blockIdx.x * kSortBlock;
    const int n = base + i;
    const bool real = n < N;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (real) { cx = cano[3 * n]; cy = cano[3 * n + 1]; cz = cano[3 * n + 2]; }
    const float* __restrict__ w = W + (int64_t)(real ? n : 0) * P;
    for (int f = 0; f < nt; ++f) {
        float ax = INFINITY, ay = INFINITY, az = INFINITY;                    // padding sorts last
        if (real) {
            ax = ay = az = 0.f;
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
        }
        sx[i] = ax; sx[kSortBlock + i] = ay; sx[2 * kSortBlock + i] = az;
        keys[i] = ((u64)orderable_bits(ax) << 32) | (u64)i;
        __syncthreads();
        // bitonic sort of the 256 (x, offset) keys, ascending
        for (int k = 2; k <= kSortBlock; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const u64 a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
                __syncthreads();
            }
        }
        const int src = (int)(keys[i] & 0xffu);
        if (base + i < n_pad) {
            float* g = out_packed + (int64_t)(t0 + f) * n_pad * 3 + (int64_t)((base + i) >> 2) * kGroupFloats + (i & 3);
            g[0] = sx[src]; g[4] = sx[kSortBlock + src]; g[8] = sx[2 * kSortBlock + src];
            perm[(int64_t)(t0 + f) * n_pad + base + i] = (unsigned char)src;
        }
        __syncthreads();
    }
}

int launch_skin_fwd_sorted(const float* cano, const float* W, const float* R, const float* tr, int64_t T, int64_t N,
                           int64_t P, float* out, float* out_packed, unsigned char* perm, int64_t n_pad,
                           cudaStream_t stream) {
    if (T <= 0 || N <= 0) return kOk;
    if (P <= 0 || P > 32 || n_pad % kSortBlock != 0 || n_pad < N) return kErrUnsupported;
    int fpb = kSkinFramesPerBlock;
    while (fpb > 1 && (n_pad / kSortBlock) * ceil_div(T, fpb) < 2 * 148) fpb /= 2;
    dim3 grid((unsigned)(n_pad / kSortBlock), (unsigned)ceil_div(T, fpb));
    const size_t smem = (size_t)kSortBlock * (8 + 12) + (size_t)fpb * P * 12 * sizeof(float);
    skin_fwd_sorted_kernel<<<grid, kSortBlock, smem, stream>>>(cano, W, R, tr, (int)T, (int)N, (int)P, out, out_packed,
                                                               perm, (int)n_pad, fpb);
    REART_CHECK_LAUNCH();
    return kOk;
}

// ----------------------------------------------------------------------------- backward
// g [T,N,3] -> gW [N,P], gR [T,P,9], gtr [T,P,3]
//   gW[n,p]  = sum_t g[t,n] . (R[t,p] c_n + tr[t,p])                (dense in p: straight-through grads)
//   gR[t,p]  = sum_n W[n,p] g[t,n] c_n^T ;  gtr[t,p] = sum_n W[n,p] g[t,n]
// Two kernels, no shared-memory atomics:
//   skin_bwd_w_kernel     one thread per point, loops over all frames, P accumulators in registers;
//   skin_bwd_pose_kernel  the pose gradients as a skinny GEMM  W^T [P x N] . G [N x 12T],  G[n,(t,k)] = g (x) [c,1]
//                         built on the fly: one thread per column (t,k), P accumulators in registers, the block's
//                         W rows broadcast from shared memory; partial sums of each n-chunk merge with one
//                         float atomic per (t,p,k) per block.
constexpr int kBwdFrames = 8;

constexpr int kBwdWPoints = 32;                              // points per block
constexpr int kBwdWThreads = kBwdWPoints * kBwdFrames;       // 256: thread = (point, frame lane)

template <int PMAX>
__global__ void __launch_bounds__(kBwdWThreads) skin_bwd_w_kernel(const float* __restrict__ cano,
                                                                  const float* __restrict__ R,
                                                                  const float* __restrict__ tr,
                                                                  const float* __restrict__ g, int T, int N, int P,
                                                                  float* __restrict__ gW) {
    extern __shared__ float sm_dyn[];
    float* sm_tf = sm_dyn;                                    // [kBwdFrames][P][12]
    float* sm_red = sm_dyn + kBwdFrames * P * 12;             // [kBwdFrames][kBwdWPoints][PMAX+1]
    const int pl = threadIdx.x % kBwdWPoints, tl = threadIdx.x / kBwdWPoints;
    const int n = blockIdx.x * kBwdWPoints + pl;
    const bool real = n < N;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (real) { cx = cano[3 * n]; cy = cano[3 * n + 1]; cz = cano[3 * n + 2]; }
    float acc[PMAX];
#pragma unroll
    for (int p = 0; p < PMAX; ++p) acc[p] = 0.f;
    for (int t0 = 0; t0 < T; t0 += kBwdFrames) {
        const int nt = min(kBwdFrames, T - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < nt * P * 12; e += blockDim.x) {
            const int f = e / (P * 12), rem = e - f * P * 12, p = rem / 12, k = rem - p * 12;
            const int64_t tp = (int64_t)(t0 + f) * P + p;
            sm_tf[e] = (k < 9) ? R[tp * 9 + k] : tr[tp * 3 + (k - 9)];
        }
        __syncthreads();
        if (real && tl < nt) {
            const float* gg = g + ((int64_t)(t0 + tl) * N + n) * 3;
            const float gx = gg[0], gy = gg[1], gz = gg[2];
            const float* tf = sm_tf + tl * P * 12;
#pragma unroll
            for (int p = 0; p < PMAX; ++p) {
                if (p < P) {
                    const float* m = tf + p * 12;
                    const float vx = cx * m[0] + cy * m[1] + cz * m[2] + m[9];
                    const float vy = cx * m[3] + cy * m[4] + cz * m[5] + m[10];
                    const float vz = cx * m[6] + cy * m[7] + cz * m[8] + m[11];
                    acc[p] += gx * vx + gy * vy + gz * vz;
                }
            }
        }
    }
    float* mine = sm_red + (tl * kBwdWPoints + pl) * (PMAX + 1);
#pragma unroll
    for (int p = 0; p < PMAX; ++p) mine[p] = acc[p];
    __syncthreads();
    // thread (pl, tl) finishes parts p = tl, tl+8, ...: sum over the 8 frame lanes in a fixed order
    if (real) {
        for (int p = tl; p < P; p += kBwdFrames) {
            float s = 0.f;
#pragma unroll
            for (int f = 0; f < kBwdFrames; ++f) s += sm_red[(f * kBwdWPoints + pl) * (PMAX + 1) + p];
            gW[(int64_t)n * P + p] = s;
        }
    }
}

constexpr int kPoseThreads = 256;
constexpr int kPoseChunk = 64;                                // points per block (more CTAs: the loop is latency bound)

template <int PMAX>
__global__ void __launch_bounds__(kPoseThreads) skin_bwd_pose_kernel(const float* __restrict__ cano,
                                                                     const float* __restrict__ W,
                                                                     const float* __restrict__ g, int T, int N, int P,
                                                                     float* __restrict__ gR, float* __restrict__ gtr) {
    extern __shared__ float sm[];
    float* sW = sm;                                            // [kPoseChunk][PMAX]
    float* sC = sm + kPoseChunk * PMAX;                        // [kPoseChunk][3]
    const int n0 = blockIdx.x * kPoseChunk;
    const int cnt = min(kPoseChunk, N - n0);
    for (int e = threadIdx.x; e < cnt * PMAX; e += blockDim.x) {
        const int i = e / PMAX, p = e - i * PMAX;
        sW[e] = p < P ? W[(int64_t)(n0 + i) * P + p] : 0.f;
    }
    for (int e = threadIdx.x; e < cnt * 3; e += blockDim.x) sC[e] = cano[(int64_t)n0 * 3 + e];
    __syncthreads();
    const int col = blockIdx.y * blockDim.x + threadIdx.x;    // column (t,k) of G
    if (col >= T * 12) return;
    const int t = col / 12, k = col - t * 12;
    const int kk = k < 9 ? k / 3 : k - 9;                     // component of g
    const int l = k < 9 ? k % 3 : -1;                         // component of c (or the constant 1)
    const float* __restrict__ gp = g + ((int64_t)t * N + n0) * 3 + kk;
    float acc[PMAX];
#pragma unroll
    for (int p = 0; p < PMAX; ++p) acc[p] = 0.f;
    for (int i = 0; i < cnt; ++i) {
        float v = gp[(int64_t)i * 3];
        if (l >= 0) v *= sC[3 * i + l];
        const float4* w4 = reinterpret_cast<const float4*>(sW + i * PMAX);
#pragma unroll
        for (int q = 0; q < PMAX / 4; ++q) {
            const float4 w = w4[q];
            acc[4 * q] += w.x * v; acc[4 * q + 1] += w.y * v; acc[4 * q + 2] += w.z * v; acc[4 * q + 3] += w.w * v;
        }
    }
#pragma unroll
    for (int p = 0; p < PMAX; ++p) {
        if (p < P && acc[p] != 0.f) {
            const int64_t tp = (int64_t)t * P + p;
            if (k < 9) atomicAdd(gR + tp * 9 + k, acc[p]);
            else atomicAdd(gtr + tp * 3 + (k - 9), acc[p]);
        }
    }
}

template <int PMAX>
static int launch_skin_bwd_p(const float* cano, const float* W, const float* R, const float* tr, const float* g,
                             int64_t T, int64_t N, int64_t P, float* gW, float* gR, float* gtr, cudaStream_t stream) {
    {
        const size_t smem = ((size_t)kBwdFrames * P * 12 + (size_t)kBwdWThreads * (PMAX + 1)) * sizeof(float);
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(skin_bwd_w_kernel<PMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return kErrUnsupported;
        skin_bwd_w_kernel<PMAX><<<(unsigned)ceil_div(N, kBwdWPoints), kBwdWThreads, smem, stream>>>(
            cano, R, tr, g, (int)T, (int)N, (int)P, gW);
        REART_CHECK_LAUNCH();
    }
    {
        const size_t smem = (size_t)kPoseChunk * (PMAX + 3) * sizeof(float);
        dim3 grid((unsigned)ceil_div(N, kPoseChunk), (unsigned)ceil_div(T * 12, kPoseThreads));
        skin_bwd_pose_kernel<PMAX><<<grid, kPoseThreads, smem, stream>>>(cano, W, g, (int)T, (int)N, (int)P, gR, gtr);
        REART_CHECK_LAUNCH();
    }
    return kOk;
}

int launch_skin_bwd(const float* cano, const float* W, const float* R, const float* tr, const float* g, int64_t T,
                    int64_t N, int64_t P, float* gW, float* gR, float* gtr, cudaStream_t stream) {
    if (P <= 0 || P > 32) return kErrUnsupported;
    if (T * P > 0) {
        if (gtr == gR + T * P * 9) {                              // one flat [gR | gtr] buffer: one memset
            if (cudaMemsetAsync(gR, 0, sizeof(float) * (size_t)(T * P * 12), stream) != cudaSuccess) return kErrLaunch;
        } else {
            if (cudaMemsetAsync(gR, 0, sizeof(float) * (size_t)(T * P * 9), stream) != cudaSuccess) return kErrLaunch;
            if (cudaMemsetAsync(gtr, 0, sizeof(float) * (size_t)(T * P * 3), stream) != cudaSuccess) return kErrLaunch;
        }
    }
    if (N <= 0) return kOk;
    if (T <= 0) {
        if (cudaMemsetAsync(gW, 0, sizeof(float) * (size_t)(N * P), stream) != cudaSuccess) return kErrLaunch;
        return kOk;
    }
    if (P <= 8) return launch_skin_bwd_p<8>(cano, W, R, tr, g, T, N, P, gW, gR, gtr, stream);
    if (P <= 16) return launch_skin_bwd_p<16>(cano, W, R, tr, g, T, N, P, gW, gR, gtr, stream);
    return launch_skin_bwd_p<32>(cano, W, R, tr, g, T, N, P, gW, gR, gtr, stream);
}

}  // namespace reart
