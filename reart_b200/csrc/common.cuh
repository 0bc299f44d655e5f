// common.cuh -- shared device helpers for the reart_b200 sm_100a kernels.
//
// PTX wrappers used across kernels: packed f32x2 arithmetic (Blackwell FADD2/FMUL2/FFMA2),
// 3-input min (FMNMX3), mbarrier + 1-D TMA bulk copy (UBLKCP).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace reart {

typedef unsigned long long u64;

// ----------------------------------------------------------------------------- error codes (include/reart_b200.h)
enum : int {
    kOk = 0,
    kErrInvalidArg = -1,
    kErrWorkspace = -2,
    kErrLaunch = -3,
    kErrUnsupported = -4,
};

// the CUDA error behind the most recent kErrLaunch (reart_last_cuda_error() reports it)
inline int& last_cuda_error() {
    static int e = 0;
    return e;
}

#define REART_CHECK_LAUNCH()                                   \
    do {                                                       \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) {                              \
            reart::last_cuda_error() = (int)e__;               \
            return reart::kErrLaunch;                          \
        }                                                      \
    } while (0)

// ----------------------------------------------------------------------------- packed f32x2
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float m;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
    return m;
}

// Squared distance in the pinned contraction order (oracle/reart_oracle.c sqdist3):
//   d = fma(dz,dz, fma(dy,dy, dx*dx)),  dx = q.x - t.x
__device__ __forceinline__ float sqdist_scalar(float qx, float qy, float qz, float tx, float ty, float tz) {
    float dx = __fsub_rn(qx, tx), dy = __fsub_rn(qy, ty), dz = __fsub_rn(qz, tz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}
// Two targets at once for one query (query operand is a scalar broadcast: FADD2 Rq.F32, -Rt.F32x2)
__device__ __forceinline__ u64 sqdist_pair(u64 QX, u64 QY, u64 QZ, u64 TX, u64 TY, u64 TZ) {
    u64 dx = sub2(QX, TX), dy = sub2(QY, TY), dz = sub2(QZ, TZ);
    return fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
}

// ----------------------------------------------------------------------------- mbarrier + TMA bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__host__ __device__ __forceinline__ int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
__host__ __device__ __forceinline__ int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------- packed cloud layout
// A cloud [B,P,3] is re-laid out for the NN search as groups of 4 points
//   [x0 x1 x2 x3 | y0 y1 y2 y3 | z0 z1 z2 z3]   (48 bytes, three float4)
// padded with +INF coordinates to a multiple of kChunk points per batch element.
constexpr int kChunk = 32;               // points per arg-min chunk (and padding granule)
constexpr int kGroupFloats = 12;         // floats per group of 4 points

__host__ __device__ __forceinline__ int64_t padded_points(int64_t P) { return round_up(P > 0 ? P : 1, kChunk); }
__host__ __device__ __forceinline__ int64_t packed_floats_per_batch(int64_t P) { return padded_points(P) * 3; }

// Re-scan one arg-min chunk of a packed cloud for the FIRST point whose distance to q equals dmin
// (same arithmetic as the search loops => guaranteed hit for finite inputs).  Used by the finalize /
// fused-backward kernels to turn (min, chunk) keys into exact lowest-index arg-mins.
// Only the x lane of a group is loaded at first: d = fma(dz,dz,fma(dy,dy,dx*dx)) >= fl(dx*dx) (rounding is
// monotonic), so a candidate with fl(dx*dx) > dmin cannot match and its y/z are never fetched -- the re-scan
// moves ~1/3 of the chunk's bytes through L2.
__device__ __forceinline__ int rescan_chunk(const float* __restrict__ tpacked_b, unsigned chunk, int chunk_pts,
                                            int nt_pad, float qx, float qy, float qz, float dmin) {
    const int start = (int)chunk * chunk_pts;
    const int end = min(start + chunk_pts, nt_pad);
    const float4* __restrict__ cg = reinterpret_cast<const float4*>(tpacked_b) + (start / 4) * 3;
    const int ngroups = (end - start) / 4;
    for (int g = 0; g < ngroups; ++g) {
        const float4 X = __ldg(cg + 3 * g);
        const float dx0 = __fsub_rn(qx, X.x), dx1 = __fsub_rn(qx, X.y), dx2 = __fsub_rn(qx, X.z), dx3 = __fsub_rn(qx, X.w);
        const bool c0 = __fmul_rn(dx0, dx0) <= dmin, c1 = __fmul_rn(dx1, dx1) <= dmin;
        const bool c2 = __fmul_rn(dx2, dx2) <= dmin, c3 = __fmul_rn(dx3, dx3) <= dmin;
        if (c0 | c1 | c2 | c3) {
            const float4 Y = __ldg(cg + 3 * g + 1), Z = __ldg(cg + 3 * g + 2);
            if (c0 && sqdist_scalar(qx, qy, qz, X.x, Y.x, Z.x) == dmin) return start + 4 * g;
            if (c1 && sqdist_scalar(qx, qy, qz, X.y, Y.y, Z.y) == dmin) return start + 4 * g + 1;
            if (c2 && sqdist_scalar(qx, qy, qz, X.z, Y.z, Z.z) == dmin) return start + 4 * g + 2;
            if (c3 && sqdist_scalar(qx, qy, qz, X.w, Y.w, Z.w) == dmin) return start + 4 * g + 3;
        }
    }
    return start;   // unreachable for finite inputs (the minimum was produced by this very arithmetic)
}

// ----------------------------------------------------------------------------- batched index recovery
// The two re-scans above walk their candidates one dependent load at a time; at L2 latency that serial depth (8 probes
// of a binary search + one iteration per window candidate, the slowest lane of a warp deciding) was the whole cost of
// the energy column pass (profiles/r01_small_kernels_ncu.md: 10.5 of 32 lanes active, DRAM 3.9 %).  The batched forms
// below issue their loads in independent groups, so a recovery is 2-4 memory round trips deep.

// 32-target row chunk (chunk_pts == kChunk): the x lanes of two groups are fetched per round trip, y/z only for a
// group that holds a candidate; returns at the first exact match in index order (same answer as rescan_chunk).  Few
// registers on purpose: this pass lives on occupancy (every load is an L2 round trip).
__device__ __forceinline__ int rescan_chunk32(const float* __restrict__ tpacked_b, unsigned chunk, float qx, float qy,
                                              float qz, float dmin) {
    const int start = (int)chunk * kChunk;
    const float4* __restrict__ cg = reinterpret_cast<const float4*>(tpacked_b) + (start / 4) * 3;
#pragma unroll 1
    for (int h = 0; h < 4; ++h) {
        const float4 X0 = __ldg(cg + 6 * h), X1 = __ldg(cg + 6 * h + 3);
        const float a0 = __fsub_rn(qx, X0.x), a1 = __fsub_rn(qx, X0.y), a2 = __fsub_rn(qx, X0.z), a3 = __fsub_rn(qx, X0.w);
        const float b0 = __fsub_rn(qx, X1.x), b1 = __fsub_rn(qx, X1.y), b2 = __fsub_rn(qx, X1.z), b3 = __fsub_rn(qx, X1.w);
        const bool c0 = __fmul_rn(a0, a0) <= dmin, c1 = __fmul_rn(a1, a1) <= dmin, c2 = __fmul_rn(a2, a2) <= dmin, c3 = __fmul_rn(a3, a3) <= dmin;
        const bool e0 = __fmul_rn(b0, b0) <= dmin, e1 = __fmul_rn(b1, b1) <= dmin, e2 = __fmul_rn(b2, b2) <= dmin, e3 = __fmul_rn(b3, b3) <= dmin;
        const bool anyc = c0 | c1 | c2 | c3, anye = e0 | e1 | e2 | e3;
        float4 Y0, Z0, Y1, Z1;
        if (anyc) { Y0 = __ldg(cg + 6 * h + 1); Z0 = __ldg(cg + 6 * h + 2); }
        if (anye) { Y1 = __ldg(cg + 6 * h + 4); Z1 = __ldg(cg + 6 * h + 5); }
        if (anyc) {
            if (c0 && sqdist_scalar(qx, qy, qz, X0.x, Y0.x, Z0.x) == dmin) return start + 8 * h;
            if (c1 && sqdist_scalar(qx, qy, qz, X0.y, Y0.y, Z0.y) == dmin) return start + 8 * h + 1;
            if (c2 && sqdist_scalar(qx, qy, qz, X0.z, Y0.z, Z0.z) == dmin) return start + 8 * h + 2;
            if (c3 && sqdist_scalar(qx, qy, qz, X0.w, Y0.w, Z0.w) == dmin) return start + 8 * h + 3;
        }
        if (anye) {
            if (e0 && sqdist_scalar(qx, qy, qz, X1.x, Y1.x, Z1.x) == dmin) return start + 8 * h + 4;
            if (e1 && sqdist_scalar(qx, qy, qz, X1.y, Y1.y, Z1.y) == dmin) return start + 8 * h + 5;
            if (e2 && sqdist_scalar(qx, qy, qz, X1.z, Y1.z, Z1.z) == dmin) return start + 8 * h + 6;
            if (e3 && sqdist_scalar(qx, qy, qz, X1.w, Y1.w, Z1.w) == dmin) return start + 8 * h + 7;
        }
    }
    return start;   // unreachable for finite inputs
}

// x-SORTED copy of a cloud, read ONLY by the index recovery (the search streams the other cloud and reads this one in
// its original AoS order), so its layout is chosen for the recovery: per block of kSortedChunk = 256 points
//     xs[256] | yz[256][2]          (3 * 256 floats = the same 12 B/point as any packed copy)
// sorted ascending by x (+INF padding last), with perm[k] = original offset of the point at sorted position k and
// a quantile index xq[16] = xs[15], xs[31], ..., xs[255].  A recovery probes the 64-byte quantile line, then the 16
// contiguous x values of the segment (two 16-way probes instead of an 8-step binary search), then walks the window
// eight x values (one 32-byte sector) per round trip and fetches (y,z) only for candidates.
constexpr int kSortedChunk = 256;
constexpr int kQuantiles = 16;
constexpr int kSortedChunkFloats = 3 * kSortedChunk;

__device__ __forceinline__ int count_below(const float4& v, float lo) {
    return (v.x < lo ? 1 : 0) + (v.y < lo ? 1 : 0) + (v.z < lo ? 1 : 0) + (v.w < lo ? 1 : 0);
}

// One batch of the window walk: eight consecutive sorted positions k0 .. k0+7 (x values in X0, X1).  All candidate
// (y,z) loads are issued together (predicated, independent), then tested -- one memory round trip per batch.
__device__ __forceinline__ void sorted_batch_test(const float2* __restrict__ yz, const unsigned char* __restrict__ pm,
                                                  int k0, const float4& X0, const float4& X1, float qx, float qy, float qz,
                                                  float dmin, int& best) {
    const float x[8] = {X0.x, X0.y, X0.z, X0.w, X1.x, X1.y, X1.z, X1.w};
    bool c[8];
    float2 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float d = __fsub_rn(qx, x[e]);
        c[e] = __fmul_rn(d, d) <= dmin;                        // false for +INF padding and for everything outside the window
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (c[e]) v[e] = __ldg(yz + k0 + e);
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (c[e] && sqdist_scalar(qx, qy, qz, x[e], v[e].x, v[e].y) == dmin) best = min(best, (int)pm[k0 + e]);
}

__device__ __forceinline__ int rescan_sorted_chunk_q(const float* __restrict__ sorted_b, const unsigned char* __restrict__ perm_b,
                                                     const float* __restrict__ xq_b, unsigned chunk, float qx, float qy,
                                                     float qz, float dmin) {
    const int base = (int)chunk * kSortedChunk;
    const float* __restrict__ blk = sorted_b + (int64_t)chunk * kSortedChunkFloats;
    const float4* __restrict__ xs4 = reinterpret_cast<const float4*>(blk);                        // 64 float4 of x
    const float2* __restrict__ yz = reinterpret_cast<const float2*>(blk + kSortedChunk);
    const unsigned char* __restrict__ pm = perm_b + base;
    const float r = sqrtf(dmin) * 1.00001f + 1e-30f;
    const float lo = qx - r - fabsf(qx) * 1e-6f, hi = qx + r + fabsf(qx) * 1e-6f;
    const float4* __restrict__ q4 = reinterpret_cast<const float4*>(xq_b + (int64_t)chunk * kQuantiles);
    const float4 q0 = __ldg(q4), q1 = __ldg(q4 + 1), q2 = __ldg(q4 + 2), q3 = __ldg(q4 + 3);
    int s = count_below(q0, lo) + count_below(q1, lo) + count_below(q2, lo) + count_below(q3, lo);
    s = min(s, kQuantiles - 1);
    const float4 a0 = __ldg(xs4 + 4 * s), a1 = __ldg(xs4 + 4 * s + 1), a2 = __ldg(xs4 + 4 * s + 2), a3 = __ldg(xs4 + 4 * s + 3);
    const int a = 16 * s + count_below(a0, lo) + count_below(a1, lo) + count_below(a2, lo) + count_below(a3, lo);
    int best = 0x7fffffff;
    // window walk, two groups (8 positions) per step, the NEXT step's x values already in flight
    const float4 inf4 = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
    constexpr int G = kSortedChunk / 4;
    int g = (a >> 2) & ~1;                                     // even group: the pair (g, g+1) is always inside the block
    float4 X0 = g < G ? __ldg(xs4 + g) : inf4, X1 = g < G ? __ldg(xs4 + g + 1) : inf4;
#pragma unroll 1
    while (g < G) {
        const int gn = g + 2;
        const float4 N0 = gn < G ? __ldg(xs4 + gn) : inf4, N1 = gn < G ? __ldg(xs4 + gn + 1) : inf4;
        sorted_batch_test(yz, pm, 4 * g, X0, X1, qx, qy, qz, dmin, best);
        if (X1.w > hi) break;                                  // sorted: nothing further can match
        X0 = N0; X1 = N1; g = gn;
    }
    return base + (best == 0x7fffffff ? 0 : best);
}

// One axis of a skinned point, R_row . c + t, with the operation order pinned: every kernel that skins (skin.cu, and the
// producer side of chamfer_sym.cu) must give the same bits, because the search holds its rows in registers while the
// energy passes re-read them from memory.  (This is the contraction nvcc chose for the plain expression in round 1.)
__device__ __forceinline__ float skin_axis(float cx, float cy, float cz, float m0, float m1, float m2, float t) {
    return __fadd_rn(__fmaf_rn(cz, m2, __fmaf_rn(cx, m0, __fmul_rn(cy, m1))), t);
}

// Bitonic sort of 256 keys held by ONE warp, 8 per lane (key index e = 8 * lane + r), ascending in e.  Exchange distances
// below 8 are register swaps inside a lane, the others one 64-bit shuffle per key; no shared memory, no barrier.
__device__ __forceinline__ void warp_bitonic_sort256(u64 (&key)[8], int lane) {
#pragma unroll
    for (int k = 2; k <= 256; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 8) {
                const bool lower = (lane & (j >> 3)) == 0;                          // (e & j) == 0
                const bool asc = k >= 256 ? true : ((lane & (k >> 3)) == 0);        // (e & k) == 0
                const bool keep_min = lower == asc;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const u64 other = __shfl_xor_sync(0xffffffffu, key[r], j >> 3);
                    key[r] = keep_min ? (other < key[r] ? other : key[r]) : (other > key[r] ? other : key[r]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if ((r & j) == 0) {                                             // pair (r, r | j), r the lower index
                        const bool asc = k >= 8 ? (k >= 256 ? true : ((lane & (k >> 3)) == 0)) : ((r & k) == 0);
                        const u64 a = key[r], c = key[r | j];
                        const u64 lo = a < c ? a : c, hi = a < c ? c : a;
                        key[r] = asc ? lo : hi;
                        key[r | j] = asc ? hi : lo;
                    }
                }
            }
        }
    }
}

// Writes one point of a sorted block (used by the two kernels that build the copy).
__device__ __forceinline__ void sorted_block_store(float* __restrict__ sorted_b, int64_t chunk, int k, float x, float y, float z) {
    float* blk = sorted_b + chunk * kSortedChunkFloats;
    blk[k] = x;
    reinterpret_cast<float2*>(blk + kSortedChunk)[k] = make_float2(y, z);
}

// ----------------------------------------------------------------------------- block sort for the x-sorted copies
__device__ __forceinline__ unsigned orderable_bits(float v) {
    const unsigned u = __float_as_uint(v);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}

// Bitonic sort of one (x, offset) key per thread across a 256-thread block, ascending by thread index.  Exchange
// distances below 32 stay inside a warp (register shuffles); only the 6 steps with distance >= 32 go through shared
// memory (round 1 ran all 36 steps through shared memory with a barrier each: 62.8 us for 64 x 16k points).
// The exchange buffer must NOT be __restrict__: with it nvcc treats every store but the last as dead (no same-thread
// read in between) and deletes them across the barriers.
__device__ __forceinline__ u64 block_bitonic_sort256(u64 key, volatile u64* xchg) {
    const int i = threadIdx.x;
#pragma unroll
    for (int k = 2; k <= kSortedChunk; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            u64 other;
            if (j >= 32) {
                xchg[i] = key;
                __syncthreads();
                other = xchg[i ^ j];
                __syncthreads();
            } else {
                other = __shfl_xor_sync(0xffffffffu, key, j);
            }
            const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
            key = keep_min ? (other < key ? other : key) : (other > key ? other : key);
        }
    }
    return key;
}


// Fixed-point scale shared by the two energy passes: every reverse-direction gradient term is bounded by
// |2 gscale| * D along an axis, D = sqrt(largest column minimum), and at most M of them meet in one accumulator,
// so 2^k with k = 62 - ceil(log2(M |2 gscale| D)) keeps any sum inside an int64.  Integer addition is associative:
// the scatter becomes independent of the order in which the atomics land.
__device__ __forceinline__ int fixed_point_exponent(unsigned bound_bits, float g2abs, int M) {
    const float D = sqrtf(__uint_as_float(bound_bits)) * 1.0001f;
    const float maxsum = (float)M * g2abs * D;
    if (!(maxsum > 0.f) || !(maxsum < 3.0e38f)) return 0;
    int e;
    frexpf(maxsum, &e);                                                            // maxsum < 2^e
    return max(-100, min(100, 61 - e));
}

// ----------------------------------------------------------------------------- seg MLP forward, shared by segmlp.cu / relax.cu
// logits = W2 relu(W0 x + b0) for ONE point by kMlpLanes consecutive lanes of a warp: lane l covers the hidden units
// k = l, l + kMlpLanes, ...; partial logits meet through xor shuffles, after which every lane of the point holds all of
// them.  s0 [H][4] = (w0 row, bias), s2 [H][kS2] = W2 transposed, zero padded (kS2 = PMAX + 4 keeps float4 alignment and
// spreads the transposing stores over the banks).  Both kernels use this one function, so their logits agree bit for bit.
constexpr int kMlpLanes = 8;

template <int PMAX>
__device__ __forceinline__ void segmlp_stage(const float* __restrict__ w0, const float* __restrict__ b0,
                                             const float* __restrict__ w2, int H, int P, float* __restrict__ s0,
                                             float* __restrict__ s2) {
    constexpr int kS2 = PMAX + 4;
    for (int e = threadIdx.x; e < H; e += blockDim.x) {
        s0[4 * e] = w0[3 * e]; s0[4 * e + 1] = w0[3 * e + 1]; s0[4 * e + 2] = w0[3 * e + 2]; s0[4 * e + 3] = b0[e];
    }
    for (int e = threadIdx.x; e < P * H; e += blockDim.x) {      // coalesced over the [P][H] weights
        const int p = e / H, k = e - p * H;
        s2[k * kS2 + p] = w2[e];
    }
    for (int e = threadIdx.x; e < H * (kS2 - P); e += blockDim.x) {
        const int k = e / (kS2 - P), p = P + (e - k * (kS2 - P));
        s2[k * kS2 + p] = 0.f;
    }
}

template <int PMAX>
__device__ __forceinline__ void segmlp_point(const float* __restrict__ s0, const float* __restrict__ s2, int H, int part,
                                             float px, float py, float pz, float (&acc)[PMAX]) {
    constexpr int kS2 = PMAX + 4;
#pragma unroll
    for (int p = 0; p < PMAX; ++p) acc[p] = 0.f;
    for (int k = part; k < H; k += kMlpLanes) {
        const float4 w = reinterpret_cast<const float4*>(s0)[k];
        const float h = fmaxf(w.x * px + w.y * py + w.z * pz + w.w, 0.f);
        const float4* c4 = reinterpret_cast<const float4*>(s2 + k * kS2);
#pragma unroll
        for (int q = 0; q < PMAX / 4; ++q) {
            const float4 c = c4[q];
            acc[4 * q] += c.x * h; acc[4 * q + 1] += c.y * h; acc[4 * q + 2] += c.z * h; acc[4 * q + 3] += c.w * h;
        }
    }
#pragma unroll
    for (int p = 0; p < PMAX; ++p) {
#pragma unroll
        for (int o = 1; o < kMlpLanes; o <<= 1) acc[p] += __shfl_xor_sync(0xffffffffu, acc[p], o);
    }
}

}  // namespace reart
