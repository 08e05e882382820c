// common.cuh -- shared declarations of librecsys_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/recsys_b200.h"

struct rl_context {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t total_mem = 0;
  cudaStream_t stream = nullptr;      // stream kernels are enqueued on
  cudaStream_t own_stream = nullptr;  // created by rl_create
  cudaStream_t copy_stream = nullptr; // host <-> device traffic of the *_host entry points, overlapped with the kernels
  cudaStream_t aux_stream = nullptr;  // second compute stream: the cooperative long-row ALS kernel runs beside the warp-per-row one
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
  cudaEvent_t free_guard = nullptr;   // orders the release of staging buffers behind the copy stream (DevBuf)
  int32_t* long_list = nullptr;       // rows for the cooperative ALS kernel (device), capacity long_list_rows
  int64_t long_list_rows = 0;
  unsigned int* counter = nullptr;    // device work counters (row schedulers), 64 x u32
  int64_t launches = 0;
  std::string err;
  // grow-only device scratch (svd, scoring candidates, ...)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // ALS k = 64 path: does this CSR (keyed by its device indptr pointer and row count) hold rows long enough for the
  // cooperative kernel?  Answered once per pattern by a reduction over indptr; a stale "no" only costs speed.
  struct LongRowsEntry { const void* indptr; int64_t n_rows; bool has_long; };
  std::vector<LongRowsEntry> long_rows_cache;
  // scoring: number of entries of a seen-item pattern (keyed like long_rows_cache), read from the device once
  struct SeenNnzEntry { const void* indptr; int64_t n_rows; int64_t nnz; };
  std::vector<SeenNnzEntry> seen_nnz_cache;
  unsigned int preloaded_exact = 0;  // bit (kp >> 4): fallback kernels of the tensor-core scoring path are loaded
  unsigned int last_overflow = 0;  // rows the last tensor-core scoring call handed to the exact kernel
  // Centred scoring (score_tc2.cu): the preprocessing of the item factors depends on V alone.  The chunked host wrappers score
  // several user ranges against the same V: 1 = armed (the next call prepares V and goes to 2), 2 = the prepared V in the scratch
  // buffer may be reused (same scratch allocation, V pointer, shape), 0 = every call prepares (plain device calls: V may have changed)
  int tc2_prep_state = 0;
  const void* tc2_prep_v = nullptr;
  const void* tc2_prep_scratch = nullptr;
  size_t tc2_prep_scratch_bytes = 0;
  int64_t tc2_prep_items = 0;
  int tc2_prep_kp = 0;
  std::vector<std::pair<void*, void*>> ipc_bases;  // (pointer handed out by rl_ipc_open, base of the mapping)
};

namespace rl {

extern thread_local std::string g_last_error;

int fail(rl_context* h, int code, const char* fmt, ...);
int cuda_fail(rl_context* h, cudaError_t e, const char* what);
// returns a device scratch buffer of at least `bytes` (grow-only, stream-ordered with respect to h->stream)
int scratch(rl_context* h, size_t bytes, void** out);

#define RL_CUDA(h, expr)                                         \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return rl::cuda_fail((h), _e, #expr); \
  } while (0)

#define RL_LAUNCH_CHECK(h, name)                                  \
  do {                                                            \
    cudaError_t _e = cudaGetLastError();                          \
    if (_e != cudaSuccess) return rl::cuda_fail((h), _e, (name)); \
    (h)->launches++;                                              \
  } while (0)

constexpr unsigned FULL_MASK = 0xffffffffu;

__host__ __device__ constexpr int padded_k(int k) {
  return k <= 0 ? 0 : k <= 16 ? 16 : k <= 32 ? 32 : k <= 64 ? 64 : k <= 128 ? 128 : 0;
}

// power-of-two scale s with m * s <= 2^14 (fp16 headroom: hi parts stay finite, lo parts stay normal down to 2^-17 m)
__host__ __device__ __forceinline__ float pow2_scale(float m) {
  if (!(m > 0.f) || !(m < 3.0e38f)) return 1.0f;
  int e;
  frexpf(m, &e);                       // m = f * 2^e, f in [0.5, 1)
  e = 14 - e;
  e = e > 100 ? 100 : (e < -100 ? -100 : e);
  return ldexpf(1.0f, e);
}
// ---- kernels' host launchers (one per .cu) --------------------------------------------------------------
int launch_als_half_step(rl_context* h, int64_t n_rows, int64_t n_other, int k, const int64_t* indptr,
                         const int32_t* indices, const float* conf, const float* pref, const float* other,
                         float* mine, double lambda, double w0, const double* gram0, int precision,
                         int ir_steps, const int32_t* row_order, int n_peers = 0, float* const* peers = nullptr);
int launch_gram(rl_context* h, const float* y, int64_t n, int k, double* gram);
int launch_gram_tc(rl_context* h, const float* x, int ldx, int op_x, const float* y, int ldy, int op_y, int64_t n, int p, int q,
                   double* g, int ldg, int terms);
int launch_csr_transpose(rl_context* h, int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                         int64_t* indptr_t, int32_t* indices_t);
// Callers that score one problem in several row ranges (rl_recommend_host) collect the rows the tensor-core pass hands
// to the exact kernel in ONE list of absolute row numbers and finish them with a single launch_score_overflow at the
// end, instead of a host round trip and a small launch per range.
struct ScoreDefer {
  int32_t* rows;            // device, room for every row of the whole problem
  unsigned int* count;      // device, zeroed by the caller
  int64_t row_offset;       // absolute row number of row 0 of this call
};
int launch_score_topn(rl_context* h, const float* uf, int64_t n_users, const float* vf, int64_t n_items, int k,
                      const int64_t* seen_indptr, const int32_t* seen_indices, double gbias, const float* ubias,
                      const float* ibias, int n, int32_t* idx_out, double* score_out, int mode,
                      const ScoreDefer* defer = nullptr);
// the deferred rows: exact kernel over the whole problem's arrays (all pointers at row 0); n_rows from the counter
int launch_score_overflow(rl_context* h, const float* uf, int64_t n_users, const float* vf, int64_t n_items, int k,
                          const int64_t* seen_indptr, const int32_t* seen_indices, int n, int32_t* idx_out, double* score_out,
                          const int32_t* rows, unsigned int n_rows);
int launch_score_dump(rl_context* h, const float* uf, int64_t n_users, const float* vf, int64_t n_items, int k, float* out,
                      int64_t ld, int terms);
int launch_score_dump_centred(rl_context* h, const float* uf, int64_t n_users, const float* vf, int64_t n_items, int k, float* out,
                              int64_t ld, float* row_scale, int32_t* perm, void* table);
int launch_topn_dense(rl_context* h, const void* est, int est_dtype, const void* known, int known_dtype,
                      int64_t n_users, int64_t n_items, int n, int32_t* idx_out, int32_t* n_valid_out);
int launch_svd_topk(rl_context* h, const float* m_dev, int64_t m, int64_t n, int k, double* s_out, double* left_out,
                    double* right_out);
int launch_lowrank(rl_context* h, const void* a, const void* b, int in_dtype, int64_t m, int64_t n, int k, double gbias,
                   const float* rbias, const float* cbias, void* out, int out_dtype);
int launch_observed_error(rl_context* h, const float* uf, const float* vf, int k, int64_t n_users,
                          const int64_t* indptr, const int32_t* indices, const float* values, double* out2);
int launch_dense_error(rl_context* h, const void* pred, int pdtype, const void* actual, int adtype, int64_t n, double* out4,
                       int* flags);
int launch_knn_item_sim(rl_context* h, const float* r, int64_t m, int n, double* sim);
int launch_knn_predict(rl_context* h, const float* r, int64_t m, int n, const double* sim, int k, double* pred);
int launch_ranking_metrics(rl_context* h, const int32_t* recs, int64_t n_users, int n, int k, const int64_t* a_indptr,
                           const int32_t* a_indices, double* out4);
int launch_bias_centre(rl_context* h, const float* r, int64_t m, int64_t n, int as_float32, double* tot, double* ub, double* ib,
                       float* centred);
int launch_factors_convert(rl_context* h, const void* src, int src_dtype, int64_t rows, int k, int src_ld,
                           void* dst, int dst_dtype, int dst_ld);

// ---- small device helpers --------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace rl
