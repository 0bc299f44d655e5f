// cull.cu -- the static bounds the culled search schedule (chamfer_sym.cu, CULL = true) tests against.
//
// The brute-force search evaluates every (skinned point, observed point) pair of a frame; an optimisation moves the
// clouds a little per iteration, so the nearest neighbours of the PREVIOUS evaluation are excellent candidates now:
//   ub_i = d(a_i, b[nn_prev(i)])  >=  min_j d(a_i, b_j)          (any real target bounds the minimum from above)
// and likewise per observed point.  One pass turns them into
//   rowbound[b][rc]  = max of ub over the 256 rows of row chunk rc   (one warp of the search owns exactly these rows)
//   colbox[b][cc]    = bounding box of the 32 targets of chunk cc, and the max of ub over them,
// and the search skips a (row chunk, target chunk) pair whose box-to-box squared gap -- a lower bound of every computed
// pair distance in it -- is STRICTLY above both bounds: such a pair can hold neither a minimum nor a tie.  The first
// evaluation (no history, nn = -1) gets infinite bounds and is the brute-force search.  The distances here use the very
// arithmetic of the search (sqdist_scalar), so the bounds bound the COMPUTED minima and the culled keys are bit-identical.
// How much is skipped depends on how compact the chunks are: callers that control the point order (engine.py) sort
// both clouds into k-d leaves once; any order is correct.
#include "common.cuh"
#include "kernels.h"

namespace reart {

constexpr int kCullRowChunk = 256;
constexpr int kCullSeeds = 4;                                 // neighbours on each side whose seeds are also tried

__global__ void __launch_bounds__(kCullRowChunk) cull_row_bounds_kernel(const CullParams p, int row_chunks) {
    __shared__ unsigned s_w[kCullRowChunk / 32];
    const int rc = blockIdx.x, b = blockIdx.y;
    const int i = rc * kCullRowChunk + threadIdx.x;
    unsigned bits = 0u;                                        // distances are >= 0: their bit patterns order like the floats
    if (i < p.na) {
        // candidates: the row's own previous arg-min and those of its neighbours in the caller's point order (spatial
        // neighbours when the caller sorted the cloud, engine.py) -- a point that changed part since the last step has a
        // useless seed of its own, but a neighbour that already belongs to the new part has a good one.  ANY real target
        // bounds the minimum from above, so the smallest of the candidates is still an exact bound.
        float ub = INFINITY;
        if (p.nn_rows) {
            const float* a = p.a + ((int64_t)b * p.na + i) * 3;
            const float ax = a[0], ay = a[1], az = a[2];
#pragma unroll
            for (int d = -kCullSeeds; d <= kCullSeeds; ++d) {
                const int in = min(max(i + d, 0), p.na - 1);
                const int j = p.nn_rows[(int64_t)b * p.na + in];
                if (j >= 0 && j < p.nb) {
                    const float* t = p.b + ((int64_t)b * p.nb + j) * 3;
                    const float c = sqdist_scalar(ax, ay, az, t[0], t[1], t[2]);
                    if (c < ub) ub = c;                        // NaN never passes: no culling on NaN input
                }
            }
        }
        bits = __float_as_uint(ub);
    }
    bits = __reduce_max_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned m = 0u;
#pragma unroll
        for (int w = 0; w < kCullRowChunk / 32; ++w) m = max(m, s_w[w]);
        p.rowbound[(int64_t)b * row_chunks + rc] = __uint_as_float(m);
    }
}

// one warp per 32-target chunk
__global__ void __launch_bounds__(256) cull_col_boxes_kernel(const CullParams p, int chunks_total) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (int64_t)p.B * chunks_total) return;
    const int b = (int)(w / chunks_total), cc = (int)(w - (int64_t)b * chunks_total);
    const int j = cc * kChunk + lane;
    float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
    unsigned bits = 0u;
    if (j < p.nb) {
        const float* t = p.b + ((int64_t)b * p.nb + j) * 3;
        const float x = t[0], y = t[1], z = t[2];
        lx = hx = x; ly = hy = y; lz = hz = z;
        // candidates: the previous arg-min rows of this target and of its neighbours in the caller's order, and the rows
        // next to the own one (a row that changed part moved away; its neighbours in the canonical order mostly did not)
        float ub = INFINITY;
        if (p.nn_cols) {
#pragma unroll
            for (int d = -kCullSeeds; d <= kCullSeeds; ++d) {
                const int jn = min(max(j + d, 0), p.nb - 1);
                const int i0 = p.nn_cols[(int64_t)b * p.nb + jn];
                if (i0 < 0 || i0 >= p.na) continue;
#pragma unroll
                for (int e = -1; e <= 1; ++e) {
                    if (d != 0 && e != 0) continue;            // row neighbours only around the target's own seed
                    const int i = min(max(i0 + e, 0), p.na - 1);
                    const float* a = p.a + ((int64_t)b * p.na + i) * 3;
                    const float c = sqdist_scalar(a[0], a[1], a[2], x, y, z);
                    if (c < ub) ub = c;
                }
            }
        }
        bits = __float_as_uint(ub);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o));
        lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
        hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
    }
    bits = __reduce_max_sync(0xffffffffu, bits);
    if (lane == 0) {
        float4* o = reinterpret_cast<float4*>(p.colbox + w * 8);
        o[0] = make_float4(lx, ly, lz, hx);
        o[1] = make_float4(hy, hz, __uint_as_float(bits), 0.f);
    }
}

int launch_cull_bounds(const CullParams& p, cudaStream_t stream) {
    if (p.B <= 0 || p.na <= 0 || p.nb <= 0) return kOk;
    if (p.B > 65535) return kErrUnsupported;
    const int row_chunks = (int)ceil_div(p.na, kCullRowChunk);
    const int chunks_total = p.nb_pad / kChunk;
    dim3 grid((unsigned)row_chunks, (unsigned)p.B);
    cull_row_bounds_kernel<<<grid, kCullRowChunk, 0, stream>>>(p, row_chunks);
    REART_CHECK_LAUNCH();
    const int64_t warps = (int64_t)p.B * chunks_total;
    cull_col_boxes_kernel<<<(unsigned)ceil_div(warps, 8), 256, 0, stream>>>(p, chunks_total);
    REART_CHECK_LAUNCH();
    return kOk;
}

}  // namespace reart
