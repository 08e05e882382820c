/*
 * recsys_b200.h -- C ABI of librecsys_b200.so (hand-written sm_100a kernels for the recsys-lite hot path).
 *
 * The reference (Lunexa-AI/recsys-lite) is pure Python and has no FFI, plugin registry or operator
 * table: its extension seam is the Python API of src/recsys_lite/algo.py.
 * Each entry point below therefore names the reference FUNCTION (file:line) whose arithmetic it
 * replaces; the Python adapter recsys_lite_b200/api.py keeps the reference's validation, error
 * strings and output post-processing and calls these through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; no exception crosses the boundary;
 *   - every function returns 0 on success, a negative rl_status otherwise; rl_last_error(h) gives text;
 *   - "_dev" pointers are DEVICE pointers owned by the caller (cudaMalloc / torch data_ptr()); kernels are
 *     enqueued on the handle's stream (rl_set_stream) and the call returns without synchronising;
 *   - "_host" entry points take HOST pointers, perform H2D / compute / D2H and return synchronised;
 *   - inputs are never modified unless the parameter is documented in/out;
 *   - factor matrices on the device are float32, row-major, with row stride rl_padded_k(k) floats and the
 *     padding columns zero;
 *   - index arrays: indptr int64 (n_rows+1), indices int32, ascending inside a row.
 */
#ifndef RECSYS_B200_H
#define RECSYS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden: only this ABI is exported */
#endif

#define RL_ABI_VERSION 1

typedef struct rl_context* rl_handle;

typedef enum {
  RL_OK = 0,
  RL_ERR_INVALID = -1,   /* bad argument (message says which)            */
  RL_ERR_CUDA = -2,      /* CUDA runtime / launch failure                */
  RL_ERR_NO_DEVICE = -3, /* no usable sm_100 device                      */
  RL_ERR_NOMEM = -4,     /* device or pinned allocation failed           */
  RL_ERR_NUMERIC = -5    /* factorisation broke down (non-SPD system)    */
} rl_status;

/* arithmetic mode of the ALS solve (rl_als_* `precision`) */
#define RL_PREC_FP32_IR 0 /* fp32 Gram + fp32 Cholesky + `ir_steps` refinement steps with an fp64 matrix-free residual */
#define RL_PREC_FP64 1    /* fp64 Gram + fp64 Cholesky (reference arithmetic, algo.py:319-322, LU replaced by Cholesky)  */

/* element types of dense matrices handed to rl_topn_dense_* */
#define RL_F32 0
#define RL_F64 1

/* ---- lifetime --------------------------------------------------------------------------------------- */
int rl_abi_version(void);
/* device < 0: current device. Fails with RL_ERR_NO_DEVICE when no CUDA device / not compute capability 10.x */
int rl_create(int device, rl_handle* out);
int rl_destroy(rl_handle h);
/* text of the last error on this handle (h may be NULL: last error of the calling thread) */
const char* rl_last_error(rl_handle h);
/* use an existing cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream; NULL is the CUDA default stream);
 * RL_OWN_STREAM switches back to the non-blocking stream the handle created for itself */
#define RL_OWN_STREAM ((void*)(intptr_t)-1)
int rl_set_stream(rl_handle h, void* cuda_stream);
int rl_synchronize(rl_handle h);
int rl_device_info(rl_handle h, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem_bytes);
/* number of kernels this library has launched on the handle since creation (bench: "gpu_launches") */
int64_t rl_launch_count(rl_handle h);
/* The *_host entry points allocate their device staging buffers from the device's stream-ordered memory pool and leave
 * them cached there between calls (no cudaMalloc / cudaFree per call).  rl_trim returns the cached memory and the
 * handle's scratch buffer to the driver. */
int rl_trim(rl_handle h);

/* row stride (floats) of device factor matrices for `k` factors: 16, 32, 64 or 128; 0 if k is unsupported */
int rl_padded_k(int k);

/* pinned host memory for the host<->device legs of the *_host entry points */
int rl_pinned_alloc(int64_t bytes, void** out);
int rl_pinned_free(void* p);
/* plain device memory helpers so a ctypes-only caller needs no other CUDA binding */
int rl_dev_alloc(rl_handle h, int64_t bytes, void** out);
int rl_dev_free(rl_handle h, void* p);
/* the same from the device's stream-ordered pool (cudaMallocAsync / cudaFreeAsync on the handle's stream: no device
 * synchronisation, freed blocks stay cached until rl_trim).  Buffers handed out by rl_als_fit_resident_host are of this
 * kind.  Not for rl_ipc_export. */
int rl_dev_alloc_async(rl_handle h, int64_t bytes, void** out);
int rl_dev_free_async(rl_handle h, void* p);
int rl_memcpy_h2d(rl_handle h, void* dst_dev, const void* src_host, int64_t bytes);
int rl_memcpy_d2h(rl_handle h, void* dst_host, const void* src_dev, int64_t bytes);
/* (rows x k) host float64/float32 -> (rows x rl_padded_k(k)) device float32 with zero padding, and back */
int rl_factors_upload(rl_handle h, const void* src_host, int src_dtype, int64_t rows, int k, float* dst_dev);
int rl_factors_download(rl_handle h, const float* src_dev, int64_t rows, int k, void* dst_host, int dst_dtype);

/* ---- ALS half-step: replaces the loops of RecommenderSystem._fit_als, algo.py:314-322 (users) and
 *      algo.py:324-332 (items).  For every row r with observations S_r (ascending):
 *          A_r = w0*G0 + sum_{i in S_r} (c_ri - w0) y_i y_i^T + lambda*I ,  b_r = sum_{i in S_r} c_ri p_ri y_i
 *          mine[r] = A_r^{-1} b_r
 *      conf_dev == pref_dev == NULL, w0 == 0 is the reference (c = p = 1); rows with no observation are left
 *      untouched (algo.py:316-317, :326-327).  w0 != 0 needs gram0_dev = Y^T Y from rl_gram (float64,
 *      kp x kp) -- Hu-Koren-Volinsky mode, NOT reference parity.
 *      other_dev (n_other x kp) is only read; mine_dev (n_rows x kp) is written.
 *      row_order_dev: optional permutation of [0, n_rows) giving the processing order (longest rows first). */
int rl_als_half_step_dev(rl_handle h, int64_t n_rows, int64_t n_other, int k,
                         const int64_t* indptr_dev, const int32_t* indices_dev,
                         const float* conf_dev, const float* pref_dev,
                         const float* other_dev, float* mine_dev,
                         double lambda, double w0, const double* gram0_dev,
                         int precision, int ir_steps, const int32_t* row_order_dev);

/* The same half-step fused with the multi-GPU exchange ("all-gather of the factor matrix once per
 * half-step"): every solved row is stored into mine_dev AND, over NVLink, into the replicas of the other ranks.
 * peer_mine_dev: host array of n_peers (<= RL_MAX_PEERS) device pointers, each the address in a PEER's replica that
 * corresponds to mine_dev (same row range), mapped into this process with rl_ipc_open.  The stores are complete when
 * the kernel is; ranks still need one barrier on the stream (e.g. a one-element all-reduce) before the replica is read. */
#define RL_MAX_PEERS 15
int rl_als_half_step_peers_dev(rl_handle h, int64_t n_rows, int64_t n_other, int k,
                               const int64_t* indptr_dev, const int32_t* indices_dev,
                               const float* conf_dev, const float* pref_dev,
                               const float* other_dev, float* mine_dev,
                               double lambda, double w0, const double* gram0_dev,
                               int precision, int ir_steps, const int32_t* row_order_dev,
                               int n_peers, float* const* peer_mine_dev);
/* CUDA IPC plumbing for the peer replicas (one process per GPU).  rl_ipc_export: 64-byte handle of the allocation that
 * contains dev_ptr + the offset of dev_ptr inside it (framework allocators sub-allocate).  rl_ipc_open: maps a peer's
 * allocation into this process (peer access over NVLink) and returns base + offset; rl_ipc_close takes that pointer. */
#define RL_IPC_HANDLE_BYTES 64
int rl_ipc_export(rl_handle h, const void* dev_ptr, void* handle_out, int64_t* offset_out);
int rl_ipc_open(rl_handle h, const void* handle, int64_t offset, void** dev_ptr_out);
int rl_ipc_close(rl_handle h, void* dev_ptr);

/* G = Y^T Y, float64 (kp x kp row-major), Y (n x kp) float32.  Used for the w0 term above. */
int rl_gram_dev(rl_handle h, const float* y_dev, int64_t n, int k, double* gram_dev);

/* CSR of the transposed observation pattern (rows = items; the reference scans dense columns instead, algo.py:325),
 * built on the device by a stable sort of the (column, row) pairs.  indptr_t_dev: n_cols + 1 int64, indices_t_dev: nnz int32,
 * ascending inside each row.  nnz = indptr[n_rows] is passed by the caller (no device -> host read). */
int rl_csr_transpose_dev(rl_handle h, int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr_dev,
                         const int32_t* indices_dev, int64_t* indptr_t_dev, int32_t* indices_t_dev);

/* whole fit on HOST buffers: replaces RecommenderSystem._fit_als, algo.py:300-337 (minus the RNG draw, which
 * stays with the caller so that np.random semantics are kept).  user_f / item_f: (n x k) float64 (RL_F64) or
 * float32 (RL_F32) host arrays, in: initial factors, out: fitted factors.  *_t is the CSR of the transpose; pass
 * indptr_t = indices_t = NULL to have it built on the device (rl_csr_transpose_dev) underneath the first user half-step. */
int rl_als_fit_host(rl_handle h, int64_t n_users, int64_t n_items, int k,
                    const int64_t* indptr, const int32_t* indices,
                    const int64_t* indptr_t, const int32_t* indices_t,
                    int iterations, double lambda, int precision, int ir_steps,
                    void* user_f, void* item_f, int factor_dtype);
/* The same fit with the model LEFT ON THE DEVICE (what RecommenderSystem.fit of the drop-in surface calls: the reference
 * keeps its model in the object, algo.py:333-337, here it stays in HBM until somebody reads it).  user_f / item_f are
 * only read.  dev_out[0..5] receive device pointers owned by the caller (release with rl_dev_free_async): U (n_users x
 * padded k float32), V, indptr, indices, indptr_t, indices_t of the pattern.  The call returns when the uploads have
 * left the host buffers; the kernels may still be running on the handle's stream. */
int rl_als_fit_resident_host(rl_handle h, int64_t n_users, int64_t n_items, int k,
                             const int64_t* indptr, const int32_t* indices,
                             const int64_t* indptr_t, const int32_t* indices_t,
                             int iterations, double lambda, int precision, int ir_steps,
                             const void* user_f, const void* item_f, int factor_dtype, void** dev_out);

/* Host-side validation of a CSR pattern before it is handed to the kernels (what scipy's has_canonical_format /
 * data.min() checks cost 60 ms for at 50 M entries, single-threaded): bit 0 of *flags_out = every row strictly
 * ascending with column indices in [0, n_cols) and indptr non-decreasing from 0; bit 1 = every stored value > 0
 * (algo.py:311: only positive cells count as observed).  indptr_dtype / data_dtype: RL_I32, RL_I64, RL_F32, RL_F64;
 * data may be NULL (bit 1 set).  Multi-threaded over row ranges (threads <= 0: all cores, at most 32). */
#define RL_I32 2
#define RL_I64 3
int rl_csr_check_host(const void* indptr, int indptr_dtype, const int32_t* indices, const void* data, int data_dtype,
                      int64_t n_rows, int64_t n_cols, int threads, int* flags_out);

/* ---- fused scoring + masking + top-n: replaces predict (algo.py:143-147 `U @ V.T`, or _predict_svd
 *      algo.py:363-368 with the bias terms) followed by top_n (algo.py:533-659) inside recommend
 *      (algo.py:172-201) without materialising the score matrix.
 *      seen_* : CSR of the items to exclude per user (NULL: exclude nothing, algo.py:202-208).
 *      bias: score += global_bias + user_bias[u] + item_bias[i] (NULL pointers = 0).
 *      idx_out (n_users x n) int32, -1 where fewer than n unseen items exist; score_out optional (float64).
 *      Order: descending exact float64 score of the float32 factors; equal scores: lower item index first. */
int rl_score_topn_dev(rl_handle h, const float* user_f_dev, int64_t n_users,
                      const float* item_f_dev, int64_t n_items, int k,
                      const int64_t* seen_indptr_dev, const int32_t* seen_indices_dev,
                      double global_bias, const float* user_bias_dev, const float* item_bias_dev,
                      int n, int32_t* idx_out_dev, double* score_out_dev, int mode);
#define RL_SCORE_AUTO 0   /* best tensor-core candidate pass the shape allows (centred single product for k <= 64,
                             split operands for k <= 128) + exact rescoring, else the exact kernel          */
#define RL_SCORE_EXACT 1  /* CUDA-core float64 scoring of every pair                                       */
#define RL_SCORE_TENSOR 2 /* force the tcgen05 candidate pass, split-operand 3 x fp16 (error if unsupported shape) */
#define RL_SCORE_TENSOR_TF32X1 3 /* same with a single fp16 product: wider candidate window, more exact-kernel fallbacks */
#define RL_SCORE_TENSOR_CENTRED 4 /* force the centred, norm-classed single-product pass (score_tc2.cu; k <= 64) */

/* debug / validation of the tensor-core path: raw tensor-core scores U V^T (terms = 1: one tf32 product, 3: split
 * operands), out (n_users x ld) float32 with ld >= n_items rounded up to 128 (the padding columns receive the
 * zero-filled tail of the last tile). */
int rl_score_dump_dev(rl_handle h, const float* user_f_dev, int64_t n_users, const float* item_f_dev, int64_t n_items,
                      int k, float* out_dev, int64_t ld, int terms);
/* the same for the centred single-product pass: raw accumulators (n_users x ld; column = position in class order,
 * ld >= (ceil(n_items / 128) + 8) * 128), the power-of-two row scales (n_users), the position -> item id map (ld int32,
 * -1 = padding) and the 256-byte class table (see score_tc2.cu: ClassTable). */
int rl_score_dump_centred_dev(rl_handle h, const float* user_f_dev, int64_t n_users, const float* item_f_dev, int64_t n_items,
                              int k, float* out_dev, int64_t ld, float* row_scale_dev, int32_t* perm_dev, void* table_dev);
/* number of user rows the last tensor-core rl_score_topn_dev call handed to the exact kernel (candidate overflow) */
int64_t rl_last_overflow_rows(rl_handle h);

int rl_recommend_host(rl_handle h, const void* user_f, const void* item_f, int factor_dtype,
                      int64_t n_users, int64_t n_items, int k,
                      const int64_t* seen_indptr, const int32_t* seen_indices,
                      double global_bias, const float* user_bias, const float* item_bias,
                      int n, int32_t* idx_out, double* score_out, int mode);
/* recommend from a model that already sits in HBM (rl_als_fit_resident_host): users are scored in chunks and the top-n
 * rows of chunk c travel to idx_out (host, pageable or pinned) while chunk c + 1 is scored. */
int rl_recommend_resident_host(rl_handle h, const float* user_f_dev, int64_t n_users, const float* item_f_dev, int64_t n_items,
                               int k, const int64_t* seen_indptr_dev, const int32_t* seen_indices_dev, int n, int32_t* idx_out,
                               int mode);

/* ---- standalone masked row top-n: replaces _top_n_numba / top_n, algo.py:533-659, on dense est / known.
 *      idx_out (n_users x n) int32 padded with -1; n_valid_out[u] = min(n, #items with known <= 0). */
int rl_topn_dense_dev(rl_handle h, const void* est_dev, int est_dtype, const void* known_dev, int known_dtype,
                      int64_t n_users, int64_t n_items, int n, int32_t* idx_out_dev, int32_t* n_valid_out_dev);
int rl_topn_dense_host(rl_handle h, const void* est, int est_dtype, const void* known, int known_dtype,
                       int64_t n_users, int64_t n_items, int n, int32_t* idx_out, int32_t* n_valid_out);

/* ---- truncated SVD: replaces np.linalg.svd + slicing in svd_reconstruct's dense branch (algo.py:510-513) and in
 *      _fit_svd_dense (algo.py:238-243).  The device computes the float64 Gram of the SHORTER side, its
 *      eigenvectors by one-sided cyclic Jacobi, and returns the rank-k factors in the form
 *          recon = left * right^T        left (m x k) float64, right (n x k) float64, s (k) float64 descending
 *      m >= n:  right = V_k (orthonormal columns), left = M V_k = U_k S_k
 *      m <  n:  left  = U_k (orthonormal columns), right = M^T U_k = V_k S_k
 *      Only reconstructions are comparable with the reference (sign / rotation ambiguity of the factors).
 *      min(m, n) <= 4096. */
int rl_svd_topk_dev(rl_handle h, const float* m_dev, int64_t m, int64_t n, int k,
                    double* s_out_dev, double* left_out_dev, double* right_out_dev);

/* out = a * b^T + global_bias + row_bias[:,None] + col_bias[None,:]  (bias pointers may be NULL).
 *      replaces `u @ (s[:, None] * vt)` (algo.py:516-517), _predict_svd (algo.py:363-368) and the ALS predict
 *      `user_factors @ item_factors.T` (algo.py:147).  in_dtype RL_F64: a (m x k), b (n x k) float64 dense;
 *      in_dtype RL_F32: device factor matrices with row stride rl_padded_k(k).  out (m x n) float32 or float64,
 *      accumulated in float64 and rounded once. */
int rl_lowrank_dev(rl_handle h, const void* a_dev, const void* b_dev, int in_dtype, int64_t m, int64_t n, int k,
                   double global_bias, const float* row_bias_dev, const float* col_bias_dev,
                   void* out_dev, int out_dtype);

/* svd_reconstruct on HOST buffers (algo.py:510-517): mat, recon (m x n) float32; optional s_out (k), left_out (m x k),
 * right_out (n x k) float64 as above; optional bias terms are added to recon (class path, algo.py:363-368). */
int rl_svd_reconstruct_host(rl_handle h, const float* mat, int64_t m, int64_t n, int k, float* recon,
                            double* s_out, double* left_out, double* right_out);

/* ALS predict on HOST buffers (algo.py:143-147): out (n_users x n_items) float64 = U V^T from (n x k) host factors. */
/* SVD class fit on HOST buffers: replaces RecommenderSystem._fit_svd_dense (algo.py:210-248) including its bias stage
 * (algo.py:218-235: global / per-user / per-item means of the entries > 0, centring of EVERY cell) -- the ratings go to
 * the device once.  ratings (m x n) float32; ratings_is_f32 != 0: the means are rounded to float32 as NumPy does for a
 * float32 matrix.  bias_out[0] = global bias (NaN without a positive entry), user_bias_out (m), item_bias_out (n)
 * float64; s / left / right as rl_svd_topk_dev (rank k of the CENTRED matrix; zeros when it is all zero or not finite:
 * status_out[0] = 1 then). */
int rl_svd_fit_host(rl_handle h, const float* ratings, int64_t m, int64_t n, int k, int ratings_is_f32,
                    double* bias_out, double* user_bias_out, double* item_bias_out,
                    double* s_out, double* left_out, double* right_out, int* status_out);
/* the bias stage alone on device buffers: tot_dev[0] / tot_dev[1] = sum / count of the positive entries (the global
 * mean), user_bias_dev (m), item_bias_dev (n) float64, centred_dev (m x n) float32 (may be NULL). */
int rl_bias_centre_dev(rl_handle h, const float* ratings_dev, int64_t m, int64_t n, int ratings_is_f32,
                       double* tot_dev, double* user_bias_dev, double* item_bias_dev, float* centred_dev);

/* G (p x q float64, row stride ldg) = op_x(X)^T op_y(Y) on the tensor cores (tcgen05, csrc/gram_tc.cu): X (n x p, row stride
 * ldx), Y (n x q, row stride ldy) float32 device matrices, contraction over the long dimension n.  op: 0 = x, 1 = x^2 of the
 * entries > 0, 2 = [x > 0], 3 = x [x > 0]; op | 8: the operand is handed over transposed ((p x n) row-major).  terms: 3 =
 * split fp16 operands (fp32-accurate), 1 = one fp16 product (exact for small integers). */
int rl_gram_tc_dev(rl_handle h, const float* x_dev, int ldx, int op_x, const float* y_dev, int ldy, int op_y, int64_t n, int p,
                   int q, double* g_dev, int ldg, int terms);
/* KNN item-item cosine over co-raters (replaces RecommenderSystem._fit_knn, algo.py:339-350) as three masked Gram products,
 * and the neighbour-weighted prediction (algo.py:148-169) as two products with the k-sparse weight matrix.  HOST buffers:
 * ratings (m x n) float32, sim (n x n) float64, pred (m x n) float64. */
int rl_knn_item_sim_host(rl_handle h, const float* ratings, int64_t m, int n, double* sim_out);
int rl_knn_predict_host(rl_handle h, const float* ratings, int64_t m, int n, const double* sim, int neighbors, double* pred_out);

/* precision@k / recall@k / NDCG@k of top-n lists against held-out item sets (reference tools.py:50-110), on HOST buffers:
 * recs (n_users x n int32, -1 padded), relevant items as CSR rows (ascending); out3 = the three user averages. */
int rl_ranking_metrics_host(rl_handle h, const int32_t* recs, int64_t n_users, int n, int k,
                            const int64_t* actual_indptr, const int32_t* actual_indices, double* out3);

int rl_predict_host(rl_handle h, const void* user_f, const void* item_f, int factor_dtype,
                    int64_t n_users, int64_t n_items, int k,
                    double global_bias, const float* user_bias, const float* item_bias,
                    void* out, int out_dtype);

/* ---- compute_rmse / compute_mae on dense HOST matrices (algo.py:662-743): out3 = {sum sq err, sum |err|, #cells with
 *      actual > 0}; flags_out bit 0: an infinite value was seen, bit 1: a NaN (the adapter raises the reference's
 *      ValueError texts). */
int rl_dense_error_host(rl_handle h, const void* pred, int pred_dtype, const void* actual, int actual_dtype,
                        int64_t n_cells, double* out3, int* flags_out);

/* ---- observed-entry error: replaces compute_rmse / compute_mae, algo.py:662-743, for predictions U V^T over a
 *      CSR pattern (values NULL = all ones).  out_host[0] = sum of squared errors, out_host[1] = sum |err|. */
int rl_observed_error_dev(rl_handle h, const float* user_f_dev, const float* item_f_dev, int k,
                          int64_t n_users, const int64_t* indptr_dev, const int32_t* indices_dev,
                          const float* values_dev, double* out2_dev);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* RECSYS_B200_H */
