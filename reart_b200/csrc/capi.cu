// capi.cu -- extern "C" boundary of libreart_b200.so (declared in include/reart_b200.h).
// Plain pointers and sizes only; argument validation and workspace carving live here, the
// kernels and their launchers in the other translation units (kernels.h).
#include "../../include/reart_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace reart;

namespace {

constexpr int64_t kAlign = 256;
inline int64_t align_up(int64_t v) { return round_up(v, kAlign); }
inline bool fits_int(int64_t v) { return v >= 0 && v <= 0x7fffffff; }

struct Carver {
    char* base;
    int64_t off;
    int64_t cap;
    bool ok;
    Carver(void* p, int64_t bytes) : base(static_cast<char*>(p)), off(0), cap(bytes), ok(p != nullptr || bytes == 0) {}
    template <typename T>
    T* take(int64_t count) {
        const int64_t bytes = align_up(count * (int64_t)sizeof(T));
        if (!ok || off + bytes > cap) { ok = false; return nullptr; }
        T* r = reinterpret_cast<T*>(base + off);
        off += bytes;
        return r;
    }
};

inline int64_t packed_bytes(int64_t B, int64_t P) { return align_up(B * packed_floats_per_batch(P) * 4); }
inline int64_t keys_bytes(int64_t B, int64_t P) { return align_up(B * P * 8); }

}  // namespace

extern "C" {

const char* reart_version(void) { return "reart_b200 0.1.0 (sm_100a)"; }

const char* reart_error_string(int code) {
    switch (code) {
        case REART_OK: return "ok";
        case REART_ERR_INVALID_ARG: return "invalid argument (null pointer, negative or oversize dimension)";
        case REART_ERR_WORKSPACE: return "workspace too small or null";
        case REART_ERR_LAUNCH: return "CUDA launch or runtime error";
        case REART_ERR_UNSUPPORTED: return "unsupported size";
        default: return "unknown error";
    }
}

int64_t reart_knn1_workspace_bytes(int64_t B, int64_t P1, int64_t P2) {
    if (B < 0 || P1 < 0 || P2 < 0) return -1;
    return packed_bytes(B, P2) + keys_bytes(B, P1) + kAlign;
}

int reart_knn1_fwd(const float* p1, const float* p2, int64_t B, int64_t P1, int64_t P2, float* dists, int64_t* idx,
                   void* workspace, int64_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (B < 0 || P1 < 0 || P2 < 0 || !fits_int(B) || !fits_int(P1) || !fits_int(padded_points(P2)))
        return REART_ERR_INVALID_ARG;
    if (B == 0 || P1 == 0) return REART_OK;
    if (!p1 || !dists || !idx) return REART_ERR_INVALID_ARG;
    if (P2 == 0) {
        if (cudaMemsetAsync(dists, 0, sizeof(float) * (size_t)(B * P1), stream) != cudaSuccess) return REART_ERR_LAUNCH;
        if (cudaMemsetAsync(idx, 0, sizeof(int64_t) * (size_t)(B * P1), stream) != cudaSuccess) return REART_ERR_LAUNCH;
        return REART_OK;
    }
    if (!p2) return REART_ERR_INVALID_ARG;
    Carver ws(workspace, workspace_bytes);
    float* packed = ws.take<float>(B * packed_floats_per_batch(P2));
    u64* keys = ws.take<u64>(B * P1);
    if (!ws.ok) return REART_ERR_WORKSPACE;
    int rc = launch_pack_cloud(p2, packed, B, P2, stream);
    if (rc) return rc;
    KnnParams p = {};
    p.ndir = 1;
    p.B = (int)B;
    p.dir[0] = KnnDir{p1, packed, keys, dists, idx, (int)P1, (int)P2, (int)padded_points(P2), 0, 0, 0, kChunk};
    rc = launch_knn1_search(p, stream);
    if (rc) return rc;
    return launch_knn1_finalize(p, stream);
}

int64_t reart_chamfer_workspace_bytes(int64_t B, int64_t N, int64_t M) {
    if (B < 0 || N < 0 || M < 0) return -1;
    return packed_bytes(B, N) + packed_bytes(B, M) + keys_bytes(B, N) + keys_bytes(B, M) + kAlign;
}

int reart_chamfer_bidir_fwd(const float* src, const float* tgt, int64_t B, int64_t N, int64_t M, float* d_fwd,
                            int64_t* i_fwd, float* d_bwd, int64_t* i_bwd, void* workspace, int64_t workspace_bytes,
                            void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (B < 0 || N < 0 || M < 0 || !fits_int(B) || !fits_int(padded_points(N)) || !fits_int(padded_points(M)))
        return REART_ERR_INVALID_ARG;
    if (B == 0) return REART_OK;
    if (N == 0 || M == 0) {
        // degenerate: each direction falls back to the zero-padding rule of the single search
        int rc = reart_knn1_fwd(src, tgt, B, N, M, d_fwd, i_fwd, workspace, workspace_bytes, stream_);
        if (rc) return rc;
        return reart_knn1_fwd(tgt, src, B, M, N, d_bwd, i_bwd, workspace, workspace_bytes, stream_);
    }
    if (!src || !tgt || !d_fwd || !i_fwd || !d_bwd || !i_bwd) return REART_ERR_INVALID_ARG;
    Carver ws(workspace, workspace_bytes);
    float* psrc = ws.take<float>(B * packed_floats_per_batch(N));
    float* ptgt = ws.take<float>(B * packed_floats_per_batch(M));
    u64* kf = ws.take<u64>(B * N);
    u64* kb = ws.take<u64>(B * M);
    if (!ws.ok) return REART_ERR_WORKSPACE;
    int rc = launch_pack_cloud(src, psrc, B, N, stream);
    if (rc) return rc;
    rc = launch_pack_cloud(tgt, ptgt, B, M, stream);
    if (rc) return rc;
    // One evaluation per (src_i, tgt_j) feeds both directions (chamfer_sym.cu).
    SymParams sp = {};
    sp.a = src; sp.b_packed = ptgt; sp.keys_a = kf; sp.keys_b = kb;
    sp.B = (int)B; sp.na = (int)N; sp.nb = (int)M; sp.nb_pad = (int)padded_points(M);
    rc = launch_chamfer_sym(sp, stream);
    if (rc) return rc;
    KnnParams p = {};
    p.ndir = 2;
    p.B = (int)B;
    p.dir[0] = KnnDir{src, ptgt, kf, d_fwd, i_fwd, (int)N, (int)M, (int)padded_points(M), 0, 0, 0, kChunk};
    p.dir[1] = KnnDir{tgt, psrc, kb, d_bwd, i_bwd, (int)M, (int)N, (int)padded_points(N), 0, 0, 0, sp.col_chunk_pts};
    return launch_knn1_finalize(p, stream);
}

int reart_knn1_bwd(const float* p1, const float* p2, const int64_t* idx, const float* grad_dists, int64_t B,
                   int64_t P1, int64_t P2, float* grad_p1, float* grad_p2, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (B < 0 || P1 < 0 || P2 < 0) return REART_ERR_INVALID_ARG;
    if (B == 0) return REART_OK;
    if ((P1 > 0 && (!p1 || !idx || !grad_dists || !grad_p1)) || (P2 > 0 && (!p2 || !grad_p2))) return REART_ERR_INVALID_ARG;
    return launch_knn1_bwd(p1, p2, idx, grad_dists, B, P1, P2, grad_p1, grad_p2, 0, stream);
}

int reart_chamfer_bidir_bwd(const float* src, const float* tgt, const int64_t* i_fwd, const int64_t* i_bwd,
                            const float* g_fwd, const float* g_bwd, int64_t B, int64_t N, int64_t M, float* grad_src,
                            float* grad_tgt, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (B < 0 || N < 0 || M < 0) return REART_ERR_INVALID_ARG;
    if (B == 0) return REART_OK;
    if (N > 0 && M > 0 && (!src || !tgt || !i_fwd || !i_bwd || !g_fwd || !g_bwd || !grad_src)) return REART_ERR_INVALID_ARG;
    return launch_chamfer_bidir_bwd(src, tgt, i_fwd, i_bwd, g_fwd, g_bwd, B, N, M, grad_src, grad_tgt, stream);
}

int reart_fp32_probe(int variant, int iters, int blocks, const float* scratch_in, float* scratch_out, double* ms,
                     double* ops_per_thread, void* stream_) {
    return launch_probe(variant, iters, blocks, scratch_in, scratch_out, ms, ops_per_thread,
                        static_cast<cudaStream_t>(stream_));
}

}  // extern "C"
