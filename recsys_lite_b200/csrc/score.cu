// score.cu -- scoring + masking + top-n entry points: exact CUDA-core path, dense top_n, observed-entry error.
#include "score_exact.cuh"

namespace rl {

namespace {

// ---- standalone masked row top-n over dense est / known (reference _top_n_numba, algo.py:533-584) -----------
// One warp per user row; est and known are streamed once with coalesced loads (HBM-bound), the running top-n
// lives in registers.  n_valid = min(n, #items with known <= 0) (algo.py:571-575).
template <typename TE, typename TK, int R>
__global__ void __launch_bounds__(256) topn_dense_kernel(const TE* __restrict__ est, const TK* __restrict__ known,
                                                         int64_t n_users, int64_t n_items, int n,
                                                         int32_t* __restrict__ idx_out, int32_t* __restrict__ n_valid) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_users) return;
  const TE* e = est + row * n_items;
  const TK* kn = known + row * n_items;
  WarpTopK<R> list;
  list.init();
  double tv;
  int tix;
  list.threshold(n, tv, tix);
  int unrated = 0;
  for (int64_t j0 = 0; j0 < n_items; j0 += 32) {
    const int64_t j = j0 + lane;
    double sv = -CUDART_INF;
    bool open = false;
    if (j < n_items) {
      open = !(static_cast<double>(kn[j]) > 0.0);
      if (open) sv = static_cast<double>(e[j]);
    }
    // n_valid counts what can be ranked: unrated items whose estimate is neither -inf nor NaN (the reference's result for
    // such estimates is implementation defined; here they are never returned and never counted)
    unrated += __popc(__ballot_sync(FULL_MASK, open && sv == sv && sv != -CUDART_INF));
    unsigned m = __ballot_sync(FULL_MASK, open && sv != -CUDART_INF && beats(sv, static_cast<int>(j), tv, tix));
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const double cs = __shfl_sync(FULL_MASK, sv, l);
      const int cj = static_cast<int>(j0) + l;
      if (beats(cs, cj, tv, tix)) {
        list.insert(cs, cj, lane);
        list.threshold(n, tv, tix);
      }
    }
  }
  list.store(n, lane, idx_out + row * n, nullptr);
  if (lane == 0 && n_valid) n_valid[row] = min(n, unrated);
}

template <typename TE, typename TK>
int launch_topn_dense_t(rl_context* h, const TE* est, const TK* known, int64_t n_users, int64_t n_items, int n,
                        int32_t* idx_out, int32_t* n_valid) {
  const unsigned grid = static_cast<unsigned>((n_users + 7) / 8);
  if (n <= 32) topn_dense_kernel<TE, TK, 1><<<grid, 256, 0, h->stream>>>(est, known, n_users, n_items, n, idx_out, n_valid);
  else if (n <= 64) topn_dense_kernel<TE, TK, 2><<<grid, 256, 0, h->stream>>>(est, known, n_users, n_items, n, idx_out, n_valid);
  else if (n <= 128) topn_dense_kernel<TE, TK, 4><<<grid, 256, 0, h->stream>>>(est, known, n_users, n_items, n, idx_out, n_valid);
  else return fail(h, RL_ERR_INVALID, "topn_dense: n=%d exceeds the on-chip list capacity of 128", n);
  RL_LAUNCH_CHECK(h, "topn_dense_kernel");
  return RL_OK;
}

// ---- sum of squared / absolute errors of U V^T over a CSR pattern (reference compute_rmse/mae, algo.py:662-743)
template <int KP>
__global__ void __launch_bounds__(256) observed_error_kernel(const float* __restrict__ uf, const float* __restrict__ vf,
                                                             int64_t n_users, const int64_t* __restrict__ indptr,
                                                             const int32_t* __restrict__ indices,
                                                             const float* __restrict__ values, double* __restrict__ out2) {
  // one warp per user; lanes split the factor dimension (KP/32 or fewer entries each) -> coalesced row reads
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  constexpr int Q = (KP + 31) / 32;
  double se = 0.0, ae = 0.0;
  for (int64_t u = warp_global; u < n_users; u += n_warps) {
    double ux[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int c = lane + 32 * q;
      ux[q] = (c < KP) ? static_cast<double>(uf[u * KP + c]) : 0.0;
    }
    for (int64_t p = indptr[u]; p < indptr[u + 1]; ++p) {
      const int64_t it = indices[p];
      double d = 0.0;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int c = lane + 32 * q;
        if (c < KP) d = fma(ux[q], static_cast<double>(vf[it * KP + c]), d);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(FULL_MASK, d, o);
      const double err = d - (values ? static_cast<double>(values[p]) : 1.0);
      se += err * err;
      ae += fabs(err);
    }
  }
  if (lane == 0) {  // every lane holds the same sums
    atomicAdd(out2, se);
    atomicAdd(out2 + 1, ae);
  }
}

// ---- compute_rmse / compute_mae on dense matrices (reference algo.py:662-743): one streaming pass, float64 sums --------
// out4: [0] sum of squared errors, [1] sum |err|, [2] number of cells with actual > 0; flags: bit 0 = inf seen, bit 1 = NaN
template <typename TP, typename TA>
__global__ void __launch_bounds__(256) dense_error_kernel(const TP* __restrict__ pred, const TA* __restrict__ actual, int64_t n,
                                                          double* __restrict__ out4, int* __restrict__ flags) {
  double se = 0.0, ae = 0.0, cnt = 0.0;
  int fl = 0;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n; p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double pv = static_cast<double>(pred[p]), av = static_cast<double>(actual[p]);
    if (isinf(pv) || isinf(av)) fl |= 1;
    if (isnan(pv) || isnan(av)) fl |= 2;
    if (av > 0.0) {
      const double d = pv - av;
      se += d * d;
      ae += fabs(d);
      cnt += 1.0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(FULL_MASK, se, o);
    ae += __shfl_xor_sync(FULL_MASK, ae, o);
    cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
    fl |= __shfl_xor_sync(FULL_MASK, fl, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out4, se);
    atomicAdd(out4 + 1, ae);
    atomicAdd(out4 + 2, cnt);
    if (fl) atomicOr(flags, fl);
  }
}

// ---- ranking metrics of top-n lists against held-out item sets (reference tools.py:50-110): one thread per user,
// recs (n_users x n, -1 padded), relevant items as CSR rows (ascending).  out4: sum over users of precision@k, recall@k
// (capped at 1), dcg@k / idcg, and the number of users counted (users without recommendations or without relevant items
// contribute 0, as in the reference).
__global__ void __launch_bounds__(256) ranking_metrics_kernel(const int32_t* __restrict__ recs, int64_t n_users, int n, int k,
                                                              const int64_t* __restrict__ a_indptr, const int32_t* __restrict__ a_indices,
                                                              double* __restrict__ out4) {
  double sp = 0.0, sr = 0.0, sn = 0.0;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < n_users; u += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t a0 = a_indptr[u], a1 = a_indptr[u + 1];
    int len = 0;
    while (len < n && recs[u * n + len] >= 0) ++len;
    if (len == 0 || a1 == a0) continue;
    const int kk = k < len ? k : len;
    int hits = 0;
    double dcg = 0.0, idcg = 0.0;
    for (int i = 0; i < kk; ++i) {
      const int item = recs[u * n + i];
      int64_t lo = a0, hi = a1;
      while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (a_indices[mid] < item) lo = mid + 1; else hi = mid;
      }
      const double w = 1.0 / log2(static_cast<double>(i) + 2.0);
      if (lo < a1 && a_indices[lo] == item) { ++hits; dcg += w; }
      idcg += w;       // the reference's ideal list counts every position once it has two or more entries (tools.py:100-103)
    }
    if (kk == 1)      // one-entry ideal list of the reference: counted iff NO item id below len(recs) is relevant (tools.py:100-103)
      idcg = a_indices[a0] < len ? 0.0 : 1.0;
    sp += static_cast<double>(hits) / static_cast<double>(k);
    const double rec = static_cast<double>(hits) / static_cast<double>(a1 - a0);
    sr += rec < 1.0 ? rec : 1.0;
    if (idcg > 0.0) sn += dcg / idcg;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sp += __shfl_xor_sync(FULL_MASK, sp, o);
    sr += __shfl_xor_sync(FULL_MASK, sr, o);
    sn += __shfl_xor_sync(FULL_MASK, sn, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out4, sp); atomicAdd(out4 + 1, sr); atomicAdd(out4 + 2, sn); }
}

}  // namespace

int launch_ranking_metrics(rl_context* h, const int32_t* recs, int64_t n_users, int n, int k, const int64_t* a_indptr,
                           const int32_t* a_indices, double* out4) {
  if (!recs || !a_indptr || !out4 || n <= 0 || k <= 0 || n_users < 0) return fail(h, RL_ERR_INVALID, "ranking_metrics: bad argument");
  RL_CUDA(h, cudaMemsetAsync(out4, 0, 4 * sizeof(double), h->stream));
  if (n_users == 0) return RL_OK;
  int64_t grid = (n_users + 255) / 256;
  if (grid > 16 * h->sm_count) grid = 16 * h->sm_count;
  ranking_metrics_kernel<<<static_cast<unsigned>(grid), 256, 0, h->stream>>>(recs, n_users, n, k, a_indptr, a_indices, out4);
  RL_LAUNCH_CHECK(h, "ranking_metrics_kernel");
  return RL_OK;
}

int launch_dense_error(rl_context* h, const void* pred, int pdtype, const void* actual, int adtype, int64_t n, double* out4, int* flags) {
  if (!pred || !actual || !out4 || !flags || n < 0) return fail(h, RL_ERR_INVALID, "dense_error: bad argument");
  RL_CUDA(h, cudaMemsetAsync(out4, 0, 4 * sizeof(double), h->stream));
  RL_CUDA(h, cudaMemsetAsync(flags, 0, sizeof(int), h->stream));
  if (n == 0) return RL_OK;
  int64_t grid = (n + 255) / 256;
  if (grid > 16 * h->sm_count) grid = 16 * h->sm_count;
  const unsigned g = static_cast<unsigned>(grid);
  if (pdtype == RL_F32 && adtype == RL_F32) dense_error_kernel<float, float><<<g, 256, 0, h->stream>>>(static_cast<const float*>(pred), static_cast<const float*>(actual), n, out4, flags);
  else if (pdtype == RL_F32 && adtype == RL_F64) dense_error_kernel<float, double><<<g, 256, 0, h->stream>>>(static_cast<const float*>(pred), static_cast<const double*>(actual), n, out4, flags);
  else if (pdtype == RL_F64 && adtype == RL_F32) dense_error_kernel<double, float><<<g, 256, 0, h->stream>>>(static_cast<const double*>(pred), static_cast<const float*>(actual), n, out4, flags);
  else if (pdtype == RL_F64 && adtype == RL_F64) dense_error_kernel<double, double><<<g, 256, 0, h->stream>>>(static_cast<const double*>(pred), static_cast<const double*>(actual), n, out4, flags);
  else return fail(h, RL_ERR_INVALID, "dense_error: unknown dtype");
  RL_LAUNCH_CHECK(h, "dense_error_kernel");
  return RL_OK;
}

bool score_tc_supported(int kp, int n, int64_t n_users, int64_t n_items, const float* ubias, const float* ibias, double gbias);
int launch_score_tc(rl_context* h, int kp, const ScoreArgs& sa, float* dump, int64_t dump_ld, int terms, const ScoreDefer* defer = nullptr);
bool score_tc2_supported(int kp, int n, int64_t n_users, int64_t n_items, const float* ubias, const float* ibias, double gbias);
int launch_score_tc2(rl_context* h, int kp, const ScoreArgs& sa, const ScoreDefer* defer, float* dump, int64_t dump_ld, float* dump_su,
                     int32_t* dump_perm, void* dump_tab);

int launch_score_dump_centred(rl_context* h, const float* uf, int64_t n_users, const float* vf, int64_t n_items, int k, float* out,
                              int64_t ld, float* row_scale, int32_t* perm, void* table) {
  const int kp = padded_k(k);
  if (!score_tc2_supported(kp, 1, n_users, n_items, nullptr, nullptr, 0.0)) return fail(h, RL_ERR_INVALID, "score_dump_centred: unsupported shape");
  if (!uf || !vf || !out || !row_scale || ld < ((n_items + 127) / 128 + 8) * 128) return fail(h, RL_ERR_INVALID, "score_dump_centred: bad argument");
  ScoreArgs a{uf, vf, n_users, n_items, nullptr, nullptr, 0.0, nullptr, nullptr, 1, nullptr, nullptr, nullptr, 0, false};
  return launch_score_tc2(h, kp, a, nullptr, out, ld, row_scale, perm, table);
}

int launch_score_dump(rl_context* h, const float* uf, int64_t n_users, const float* vf, int64_t n_items, int k, float* out, int64_t ld,
                      int terms) {
  const int kp = padded_k(k);
  if (kp == 0) return fail(h, RL_ERR_INVALID, "score_dump: k=%d outside [1, 128]", k);
  if (!uf || !vf || !out || ld < ((n_items + 127) / 128) * 128) return fail(h, RL_ERR_INVALID, "score_dump: bad argument");
  ScoreArgs a{uf, vf, n_users, n_items, nullptr, nullptr, 0.0, nullptr, nullptr, 1, nullptr, nullptr, nullptr, 0, false};
  return launch_score_tc(h, kp, a, out, ld, terms);
}

int launch_score_topn(rl_context* h, const float* uf, int64_t n_users, const float* vf, int64_t n_items, int k,
                      const int64_t* seen_indptr, const int32_t* seen_indices, double gbias, const float* ubias,
                      const float* ibias, int n, int32_t* idx_out, double* score_out, int mode, const ScoreDefer* defer) {
  const int kp = padded_k(k);
  if (kp == 0) return fail(h, RL_ERR_INVALID, "score_topn: k=%d outside [1, 128]", k);
  if (n <= 0) return fail(h, RL_ERR_INVALID, "score_topn: n must be positive, got %d", n);
  if (n_users < 0 || n_items < 0 || n_items > INT_MAX - 64) return fail(h, RL_ERR_INVALID, "score_topn: bad shape");
  if (n_users == 0) return RL_OK;
  if (!uf || !vf || !idx_out) return fail(h, RL_ERR_INVALID, "score_topn: null pointer");
  if ((seen_indptr == nullptr) != (seen_indices == nullptr) && seen_indptr == nullptr)
    return fail(h, RL_ERR_INVALID, "score_topn: seen_indices without seen_indptr");
  ScoreArgs a{uf, vf, n_users, n_items, seen_indptr, seen_indices, gbias, ubias, ibias, n, idx_out, score_out, nullptr, 0, false};
  const bool tc_ok = score_tc_supported(kp, n, n_users, n_items, ubias, ibias, gbias);
  if (mode < RL_SCORE_AUTO || mode > RL_SCORE_TENSOR_CENTRED) return fail(h, RL_ERR_INVALID, "score_topn: unknown mode %d", mode);
  const bool tc2_ok = score_tc2_supported(kp, n, n_users, n_items, ubias, ibias, gbias);
  if (mode == RL_SCORE_TENSOR_CENTRED && !tc2_ok)
    return fail(h, RL_ERR_INVALID, "score_topn: centred tensor-core path needs k <= 64, n <= 16, >= 1024 items and no bias terms");
  if ((mode == RL_SCORE_AUTO || mode == RL_SCORE_TENSOR_CENTRED) && tc2_ok)
    return launch_score_tc2(h, kp, a, defer, nullptr, 0, nullptr, nullptr, nullptr);
  if ((mode == RL_SCORE_TENSOR || mode == RL_SCORE_TENSOR_TF32X1) && !tc_ok)
    return fail(h, RL_ERR_INVALID, "score_topn: tensor-core path needs k <= 128, n <= 16, >= 1024 items and no bias terms");
  if (mode != RL_SCORE_EXACT && tc_ok) return launch_score_tc(h, kp, a, nullptr, 0, mode == RL_SCORE_TENSOR_TF32X1 ? 1 : 3, defer);
  return launch_score_exact(h, kp, a);
}

int launch_score_overflow(rl_context* h, const float* uf, int64_t n_users, const float* vf, int64_t n_items, int k,
                          const int64_t* seen_indptr, const int32_t* seen_indices, int n, int32_t* idx_out, double* score_out,
                          const int32_t* rows, unsigned int n_rows) {
  if (n_rows == 0) return RL_OK;
  ScoreArgs a{uf, vf, n_users, n_items, seen_indptr, seen_indices, 0.0, nullptr, nullptr, n, idx_out, score_out, rows,
              static_cast<int64_t>(n_rows), true};
  return launch_score_exact(h, padded_k(k), a);
}

int launch_topn_dense(rl_context* h, const void* est, int est_dtype, const void* known, int known_dtype,
                      int64_t n_users, int64_t n_items, int n, int32_t* idx_out, int32_t* n_valid_out) {
  if (n <= 0) return fail(h, RL_ERR_INVALID, "topn_dense: n must be positive, got %d", n);
  if (n_users < 0 || n_items < 0 || n_items > INT_MAX - 64) return fail(h, RL_ERR_INVALID, "topn_dense: bad shape");
  if (n_users == 0) return RL_OK;
  if (!est || !known || !idx_out) return fail(h, RL_ERR_INVALID, "topn_dense: null pointer");
  if (est_dtype == RL_F32 && known_dtype == RL_F32)
    return launch_topn_dense_t(h, static_cast<const float*>(est), static_cast<const float*>(known), n_users, n_items, n, idx_out, n_valid_out);
  if (est_dtype == RL_F32 && known_dtype == RL_F64)
    return launch_topn_dense_t(h, static_cast<const float*>(est), static_cast<const double*>(known), n_users, n_items, n, idx_out, n_valid_out);
  if (est_dtype == RL_F64 && known_dtype == RL_F32)
    return launch_topn_dense_t(h, static_cast<const double*>(est), static_cast<const float*>(known), n_users, n_items, n, idx_out, n_valid_out);
  if (est_dtype == RL_F64 && known_dtype == RL_F64)
    return launch_topn_dense_t(h, static_cast<const double*>(est), static_cast<const double*>(known), n_users, n_items, n, idx_out, n_valid_out);
  return fail(h, RL_ERR_INVALID, "topn_dense: unknown dtype");
}

int launch_observed_error(rl_context* h, const float* uf, const float* vf, int k, int64_t n_users,
                          const int64_t* indptr, const int32_t* indices, const float* values, double* out2) {
  const int kp = padded_k(k);
  if (kp == 0) return fail(h, RL_ERR_INVALID, "observed_error: k=%d outside [1, 128]", k);
  if (!uf || !vf || !indptr || !out2) return fail(h, RL_ERR_INVALID, "observed_error: null pointer");
  RL_CUDA(h, cudaMemsetAsync(out2, 0, 2 * sizeof(double), h->stream));
  if (n_users == 0) return RL_OK;
  int64_t grid = (n_users + 7) / 8;
  if (grid > 8 * h->sm_count) grid = 8 * h->sm_count;
  switch (kp) {
    case 16: observed_error_kernel<16><<<grid, 256, 0, h->stream>>>(uf, vf, n_users, indptr, indices, values, out2); break;
    case 32: observed_error_kernel<32><<<grid, 256, 0, h->stream>>>(uf, vf, n_users, indptr, indices, values, out2); break;
    case 64: observed_error_kernel<64><<<grid, 256, 0, h->stream>>>(uf, vf, n_users, indptr, indices, values, out2); break;
    case 128: observed_error_kernel<128><<<grid, 256, 0, h->stream>>>(uf, vf, n_users, indptr, indices, values, out2); break;
  }
  RL_LAUNCH_CHECK(h, "observed_error_kernel");
  return RL_OK;
}

}  // namespace rl
