/*
 * reart_b200.h -- C ABI of libreart_b200.so: B200 (sm_100a) kernels for reart's per-iteration
 * energy evaluation (skinning -> bidirectional Chamfer nearest-neighbour search -> backward).
 *
 * Conventions (they follow the reference's in-tree native wrappers,
 * networks/pointnet_lib/src/sampling.cpp:11-21, minus the exit(-1)):
 *   - every pointer is DEVICE memory owned by the caller (torch tensors); the library never
 *     allocates or frees; tensors are row-major contiguous; data float32, indices int64;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and never
 *     synchronise (except reart_fp32_probe, a benchmark helper);
 *   - return value 0 = success, negative = error (reart_error_string);
 *   - `workspace` is scratch sized by the matching *_workspace_bytes() query, 256-byte aligned.
 *   - reentrant; no global state.
 *
 * Each entry point cites the reference interface (file:line under the reference tree) it replaces.
 */
#ifndef REART_B200_H_
#define REART_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define REART_API __attribute__((visibility("default")))
#else
#define REART_API
#endif

#define REART_OK 0
#define REART_ERR_INVALID_ARG (-1)
#define REART_ERR_WORKSPACE (-2)
#define REART_ERR_LAUNCH (-3)
#define REART_ERR_UNSUPPORTED (-4)

REART_API const char* reart_version(void);
REART_API const char* reart_error_string(int code);

/* ---------------------------------------------------------------------------------------------
 * K=1 nearest neighbour, D=3.
 * Replaces chamferdist._C.knn_points_idx(p1, p2, lengths1, lengths2, K=1, version=-1)
 * as called from utils/chamfer.py:174 (via knn_points, utils/chamfer.py:212-286, full lengths).
 *   p1 [B,P1,3], p2 [B,P2,3]  ->  dists [B,P1] squared L2, idx [B,P1] (lowest index on ties).
 * P2 == 0 yields dists = 0, idx = 0 (upstream zero padding).
 * ------------------------------------------------------------------------------------------- */
REART_API int64_t reart_knn1_workspace_bytes(int64_t B, int64_t P1, int64_t P2);
REART_API int reart_knn1_fwd(const float* p1, const float* p2, int64_t B, int64_t P1, int64_t P2, float* dists, int64_t* idx,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* Both directions of ChamferDistance.forward in one pass (utils/chamfer.py:78-94):
 *   src [B,N,3], tgt [B,M,3] -> d_fwd,i_fwd [B,N] (src->tgt) and d_bwd,i_bwd [B,M] (tgt->src). */
REART_API int64_t reart_chamfer_workspace_bytes(int64_t B, int64_t N, int64_t M);
REART_API int reart_chamfer_bidir_fwd(const float* src, const float* tgt, int64_t B, int64_t N, int64_t M, float* d_fwd,
                            int64_t* i_fwd, float* d_bwd, int64_t* i_bwd, void* workspace, int64_t workspace_bytes,
                            void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward of the K=1 search.
 * Replaces chamferdist._C.knn_points_backward(p1, p2, lengths1, lengths2, idx, grad_dists)
 * as called from utils/chamfer.py:206-208:
 *   grad_p1[b,i] = 2 g[b,i] (p1[b,i] - p2[b,idx[b,i]]);  grad_p2[b,idx[b,i]] -= same.
 * Both outputs are fully overwritten.
 * ------------------------------------------------------------------------------------------- */
REART_API int reart_knn1_bwd(const float* p1, const float* p2, const int64_t* idx, const float* grad_dists, int64_t B,
                   int64_t P1, int64_t P2, float* grad_p1, float* grad_p2, void* stream);

/* Backward of both directions at once (what autograd sums over the two _knn_points.backward
 * calls, utils/chamfer.py:195-209): g_fwd [B,N], g_bwd [B,M] -> grad_src [B,N,3], grad_tgt [B,M,3].
 * grad_tgt may be NULL (observed frames need no gradient). */
REART_API int reart_chamfer_bidir_bwd(const float* src, const float* tgt, const int64_t* i_fwd, const int64_t* i_bwd,
                            const float* g_fwd, const float* g_bwd, int64_t B, int64_t N, int64_t M, float* grad_src,
                            float* grad_tgt, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FP32-pipe micro-benchmark (roofline denominator).  Runs variant
 * 0..5 (see csrc/probe.cu) once warm and once timed with CUDA events on `stream`, SYNCHRONISES,
 * and returns milliseconds and the number of measured lane-operations per thread.
 * scratch_in: >= 1024 floats (any finite values), scratch_out: >= blocks*256 floats.
 * ------------------------------------------------------------------------------------------- */
REART_API int reart_fp32_probe(int variant, int iters, int blocks, const float* scratch_in, float* scratch_out, double* ms,
                     double* ops_per_thread, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REART_B200_H_ */
