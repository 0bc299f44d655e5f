// probe.cu -- FP32 pipe micro-benchmarks used by bench.py to measure the roofline denominator
// on the box it runs on (MEASURED_PEAKS.json carries no FP32 entry).
//
// Variants (each thread runs `iters` iterations of an unrolled body on register-resident data):
//   0  FFMA   scalar fused multiply-add, 16 independent chains          (2 FLOP / lane-op)
//   1  FFMA2  packed f32x2 fused multiply-add, 8 independent chains     (4 FLOP / lane-op)
//   2  KNNMIX the NN inner loop's mix per 2 pairs: 3 FADD2 + FMUL2 + 2 FFMA2 + FMNMX3 (16 alg. FLOP)
//   3  FMNMX3 3-input min only
//   4  FADD2  packed add only
//   5  FFMA2 + FMNMX3 interleaved 1:1 (dual-issue check)
// The launcher times one launch with CUDA events on the given stream and returns milliseconds.
#include "common.cuh"
#include "kernels.h"

namespace reart {

template <int V>
__global__ void __launch_bounds__(256) probe_kernel(const float* __restrict__ in, float* __restrict__ out, int iters) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = in[(tid + i * 37) & 1023];
    const float m = in[(tid + 5) & 1023], c = in[(tid + 11) & 1023];
    float acc = 0.f;
    if (V == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], m, c);
        }
    } else if (V == 1) {
        u64 p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pack2(a[2 * i], a[2 * i + 1]);
        const u64 M = pack2(m, m), C = pack2(c, c + 1.f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], M, C);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) unpack2(p[i], a[2 * i], a[2 * i + 1]);
    } else if (V == 2) {
        u64 QX[8], QY[8], QZ[8];
        float best[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            QX[r] = pack2(a[r], a[r]); QY[r] = pack2(a[8 + r], a[8 + r]); QZ[r] = pack2(a[(3 * r + 1) & 15], a[(3 * r + 1) & 15]);
            best[r] = 3.0e38f;
        }
        u64 TX = pack2(a[12], a[13]), TY = pack2(a[14], a[15]), TZ = pack2(m, c);
        const u64 step = pack2(1e-3f, 2e-3f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    float lo, hi;
                    unpack2(sqdist_pair(QX[r], QY[r], QZ[r], TX, TY, TZ), lo, hi);
                    best[r] = min3(best[r], lo, hi);
                }
                // keep all three target coordinates changing (3 extra FADD2 per 8 query-pairs)
                TX = sub2(TX, step); TY = sub2(TY, step); TZ = sub2(TZ, step);
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) acc += best[r];
    } else if (V == 3) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = min3(a[i], a[(i + 1) & 15], a[(i + 5) & 15]);
        }
    } else if (V == 4) {
        u64 p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pack2(a[2 * i], a[2 * i + 1]);
        const u64 C = pack2(c, m);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = sub2(p[i], C);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) unpack2(p[i], a[2 * i], a[2 * i + 1]);
    } else if (V == 5) {
        u64 p[8];
        float b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { p[i] = pack2(a[2 * i], a[2 * i + 1]); b[i] = a[i] + 1.f; }
        const u64 M = pack2(m, m), C = pack2(c, c + 1.f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    p[i] = fma2(p[i], M, C);
                    b[i] = min3(b[i], b[(i + 1) & 7], b[(i + 3) & 7]);
                }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { unpack2(p[i], a[2 * i], a[2 * i + 1]); acc += b[i]; }
    }
    else if (V == 6 || V == 7 || V == 8) {
        // FFMA2 interleaved 1:1 with: 6 scalar FMNMX, 7 unsigned integer min, 8 3-input unsigned min
        u64 p[8];
        float b[8];
        unsigned ub[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { p[i] = pack2(a[2 * i], a[2 * i + 1]); b[i] = a[i] + 1.f; ub[i] = __float_as_uint(b[i]); }
        const u64 M = pack2(m, m), C = pack2(c, c + 1.f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    p[i] = fma2(p[i], M, C);
                    if (V == 6) b[i] = fminf(b[i], b[(i + 3) & 7]);
                    if (V == 7) ub[i] = min(ub[i], ub[(i + 3) & 7] + 1u);
                    if (V == 8) ub[i] = min(min(ub[i], ub[(i + 3) & 7]), ub[(i + 5) & 7] + 1u);
                }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { unpack2(p[i], a[2 * i], a[2 * i + 1]); acc += b[i] + __uint_as_float(ub[i]); }
    } else if (V == 9) {
        // scalar FFMA interleaved 1:1 with scalar FMNMX
        float b[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] = a[i] + 1.f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    a[i] = __fmaf_rn(a[i], m, c);
                    b[i] = fminf(b[i], b[(i + 3) & 15]);
                }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += b[i];
    } else if (V == 10) {
        // scalar FMNMX alone
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fminf(a[i], a[(i + 3) & 15]);
        }
    } else if (V == 11) {
        // warp-wide unsigned min reduction (REDUX) alone
        unsigned ub[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) ub[i] = __float_as_uint(a[i]);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 16; ++i) ub[i] = __reduce_min_sync(0xffffffffu, ub[i] + (unsigned)i);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += __uint_as_float(ub[i]);
    } else if (V == 12) {
        // FFMA2 interleaved 4:1 with REDUX
        u64 p[8];
        unsigned ub[2];
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pack2(a[2 * i], a[2 * i + 1]);
        ub[0] = __float_as_uint(a[3]); ub[1] = __float_as_uint(a[7]);
        const u64 M = pack2(m, m), C = pack2(c, c + 1.f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], M, C);
                ub[0] = __reduce_min_sync(0xffffffffu, ub[0] + 1u);
                ub[1] = __reduce_min_sync(0xffffffffu, ub[1] + 3u);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) unpack2(p[i], a[2 * i], a[2 * i + 1]);
        acc += __uint_as_float(ub[0]) + __uint_as_float(ub[1]);
    } else if (V == 13) {
        // scalar knn mix: 3 FADD + FMUL + 2 FFMA + FMNMX per pair
        float qx[8], qy[8], qz[8], best[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { qx[r] = a[r]; qy[r] = a[8 + r]; qz[r] = a[(3 * r + 1) & 15]; best[r] = 3.0e38f; }
        float tx = a[12], ty = a[14], tz = m;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int r = 0; r < 8; ++r) best[r] = fminf(best[r], sqdist_scalar(qx[r], qy[r], qz[r], tx, ty, tz));
                tx -= 1e-3f; ty -= 2e-3f; tz -= 1e-3f;
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) acc += best[r];
    }
    else if (V == 14 || V == 15 || V == 16) {
        // FFMA2 x8 interleaved with: 14 two (compare + ballot), 15 two SHFL.BFLY, 16 two REDUX (reference for 12)
        u64 p[8];
        unsigned ub[2];
        float fb[2];
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pack2(a[2 * i], a[2 * i + 1]);
        ub[0] = __float_as_uint(a[3]); ub[1] = __float_as_uint(a[7]); fb[0] = a[5]; fb[1] = a[9];
        const u64 M = pack2(m, m), C = pack2(c, c + 1.f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], M, C);
                if (V == 14) {
                    ub[0] += __ballot_sync(0xffffffffu, ub[0] > ub[1] + (unsigned)u);
                    ub[1] += __ballot_sync(0xffffffffu, ub[1] > ub[0] + 3u);
                } else if (V == 15) {
                    fb[0] += __shfl_xor_sync(0xffffffffu, fb[0], 16);
                    fb[1] += __shfl_xor_sync(0xffffffffu, fb[1], 8);
                } else {
                    ub[0] = __reduce_min_sync(0xffffffffu, ub[0] + 1u);
                    ub[1] = __reduce_min_sync(0xffffffffu, ub[1] + 3u);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) unpack2(p[i], a[2 * i], a[2 * i + 1]);
        acc += __uint_as_float(ub[0]) + __uint_as_float(ub[1]) + fb[0] + fb[1];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += a[i];
    out[tid] = acc;
}

// lane_ops: instructions of the measured kind executed per thread (so the caller can convert to FLOP/s)
int launch_probe(int variant, int iters, int blocks, const float* in, float* out, double* ms, double* ops_per_thread,
                 cudaStream_t stream) {
    if (!in || !out || !ms || !ops_per_thread || iters <= 0 || blocks <= 0) return kErrInvalidArg;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return kErrLaunch;
    for (int rep = 0; rep < 2; ++rep) {
        if (rep == 1) cudaEventRecord(e0, stream);
        double opi = 64.0;
        switch (variant) {
#define REART_PROBE_CASE(V, OPS) case V: probe_kernel<V><<<blocks, 256, 0, stream>>>(in, out, iters); opi = OPS; break;
            REART_PROBE_CASE(0, 64.0) REART_PROBE_CASE(1, 64.0) REART_PROBE_CASE(2, 32.0) REART_PROBE_CASE(3, 64.0)
            REART_PROBE_CASE(4, 64.0) REART_PROBE_CASE(5, 64.0) REART_PROBE_CASE(6, 64.0) REART_PROBE_CASE(7, 64.0)
            REART_PROBE_CASE(8, 64.0) REART_PROBE_CASE(9, 64.0) REART_PROBE_CASE(10, 64.0) REART_PROBE_CASE(11, 64.0)
            REART_PROBE_CASE(12, 64.0) REART_PROBE_CASE(13, 64.0) REART_PROBE_CASE(14, 64.0) REART_PROBE_CASE(15, 64.0)
            REART_PROBE_CASE(16, 64.0)
#undef REART_PROBE_CASE
            default: cudaEventDestroy(e0); cudaEventDestroy(e1); return kErrInvalidArg;
        }
        *ops_per_thread = opi * iters;
        if (rep == 1) cudaEventRecord(e1, stream);
    }
    cudaError_t err = cudaEventSynchronize(e1);
    float t = 0.f;
    if (err == cudaSuccess) err = cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) return kErrLaunch;
    *ms = (double)t;
    return kOk;
}

}  // namespace reart
