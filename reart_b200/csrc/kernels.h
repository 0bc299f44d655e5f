// kernels.h -- internal launcher interface between the .cu translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace reart {

// One search direction: every query of `q` against the packed targets `tpacked`.
struct KnnDir {
    const float* q;               // [B, nq, 3] queries (AoS)
    const float* tpacked;         // [B, nt_pad/4, 12] packed targets (common.cuh layout)
    unsigned long long* keys;     // [B, nq] (dist_bits << 32 | chunk) merge keys
    float* out_dists;             // [B, nq] or null   (finalize)
    int64_t* out_idx;             // [B, nq] or null   (finalize)
    int nq;                       // queries per batch element
    int nt;                       // real targets per batch element
    int nt_pad;                   // padded targets (multiple of kChunk)
    int qblocks;                  // filled by the launcher
    int splits;                   // filled by the launcher
    int keys_preset;              // 1: caller already set keys to 0xff.. (fused pipelines)
    int chunk_pts;                // targets per arg-min chunk recorded in the key (finalize re-scan span)
};

struct KnnParams {
    KnnDir dir[2];
    int ndir;
    int B;
    int items0;                   // number of work items of dir[0]
};

// Symmetric bidirectional search (chamfer_sym.cu): A rows in registers, B columns streamed.
struct SymParams {
    const float* a;               // [B, na, 3]
    const float* b_packed;        // [B, nb_pad/4, 12]
    unsigned long long* keys_a;   // [B, na]  row keys   (chunk = 32 B-points)
    unsigned long long* keys_b;   // [B, nb]  column keys (chunk = col_chunk_pts A-points)
    int B, na, nb, nb_pad;
    int qblocks, splits;          // filled by the launcher
    int col_chunk_pts;            // filled by the launcher (32 * R)
    int keys_preset;
};

int launch_pack_cloud(const float* pts, float* packed, int64_t B, int64_t P, cudaStream_t stream);
int launch_knn1_search(KnnParams& p, cudaStream_t stream);
int launch_knn1_finalize(const KnnParams& p, cudaStream_t stream);
int launch_chamfer_sym(SymParams& p, cudaStream_t stream);

int launch_knn1_bwd(const float* p1, const float* p2, const int64_t* idx, const float* grad_dists, int64_t B,
                    int64_t P1, int64_t P2, float* grad_p1, float* grad_p2, int accumulate, cudaStream_t stream);

int launch_chamfer_bidir_bwd(const float* src, const float* tgt, const int64_t* i_fwd, const int64_t* i_bwd,
                             const float* g_fwd, const float* g_bwd, int64_t B, int64_t N, int64_t M, float* grad_src,
                             float* grad_tgt, cudaStream_t stream);

int launch_probe(int variant, int iters, int blocks, const float* in, float* out, double* ms, double* ops_per_thread,
                 cudaStream_t stream);

}  // namespace reart
