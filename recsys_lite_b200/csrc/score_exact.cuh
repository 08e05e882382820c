// score_exact.cuh -- CUDA-core float64 scoring fused with seen-item masking and top-n (exact path).
//
// Replaces predict (reference algo.py:143-147 `U @ V.T`, or algo.py:363-368 with bias terms) + top_n
// (algo.py:533-659) as composed in recommend (algo.py:172-201).  Scores are float64 dot products of the
// float32 factors, so the only difference to the reference's float64 GEMM is summation order (~1e-16 relative).
// One CTA owns 64 users and sweeps the items in tiles of 64; the 64 x 64 score tile lives in shared memory only.
// This kernel is the exact fallback of the tensor-core candidate pass and the path for small / odd shapes.
#pragma once

#include "common.cuh"
#include "topk.cuh"

namespace rl {

struct ScoreArgs {
  const float* uf;
  const float* vf;
  int64_t n_users, n_items;
  const int64_t* seen_indptr;
  const int32_t* seen_indices;
  double gbias;
  const float* ubias;
  const float* ibias;
  int n;
  int32_t* idx_out;
  double* score_out;
  const int32_t* user_list;  // optional: score only these users
  int64_t n_list;
  bool scatter;              // with user_list: write row `user` of idx_out (true) or row `position in list` (false)
  // item slicing (short user lists: one CTA per 64 users would sweep all items alone).  slices > 1: CTA (x, y) scores
  // items [y * slice_items, (y + 1) * slice_items) and writes its list to row (position * slices + y) of part_*.
  int slices = 1;
  int64_t slice_items = 0;
  int32_t* part_idx = nullptr;
  double* part_score = nullptr;
};

constexpr int SE_UT = 64, SE_IT = 64, SE_THREADS = 256, SE_UPW = 8;  // users per warp

template <int KP>
constexpr size_t score_exact_smem() {
  return sizeof(double) * (static_cast<size_t>(KP) * SE_UT + static_cast<size_t>(KP) * SE_IT + SE_UT * (SE_IT + 1));
}

template <int KP, int R>
__global__ void __launch_bounds__(SE_THREADS) score_topn_exact_kernel(const ScoreArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ut = reinterpret_cast<double*>(smem_raw);  // [KP][UT]
  double* Vt = Ut + KP * SE_UT;                      // [KP][IT]
  double* S = Vt + KP * SE_IT;                       // [UT][IT+1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t total = a.user_list ? a.n_list : a.n_users;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * SE_UT;

  // stage the user tile (transposed, float64)
  for (int p = tid; p < SE_UT * (KP / 4); p += SE_THREADS) {
    const int u = p % SE_UT, c4 = p / SE_UT;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t0 + u < total) {
      const int64_t ur = a.user_list ? a.user_list[t0 + u] : t0 + u;
      x = *reinterpret_cast<const float4*>(a.uf + ur * KP + 4 * c4);
    }
    Ut[(4 * c4 + 0) * SE_UT + u] = x.x;
    Ut[(4 * c4 + 1) * SE_UT + u] = x.y;
    Ut[(4 * c4 + 2) * SE_UT + u] = x.z;
    Ut[(4 * c4 + 3) * SE_UT + u] = x.w;
  }

  WarpTopK<R> lists[SE_UPW];
  int64_t cur[SE_UPW], end[SE_UPW];
  double ub[SE_UPW];
#pragma unroll
  for (int uu = 0; uu < SE_UPW; ++uu) {
    lists[uu].init();
    const int64_t t = t0 + warp * SE_UPW + uu;
    cur[uu] = end[uu] = 0;
    ub[uu] = 0.0;
    if (t < total) {
      const int64_t ur = a.user_list ? a.user_list[t] : t;
      if (a.seen_indptr) { cur[uu] = a.seen_indptr[ur]; end[uu] = a.seen_indptr[ur + 1]; }
      if (a.ubias) ub[uu] = a.ubias[ur];
    }
  }
  const int tu = tid >> 4, ti = tid & 15;
  int64_t j_begin = 0, j_end = a.n_items;
  if (a.slices > 1) {
    j_begin = static_cast<int64_t>(blockIdx.y) * a.slice_items;     // a multiple of SE_IT
    j_end = min(a.n_items, j_begin + a.slice_items);
#pragma unroll
    for (int uu = 0; uu < SE_UPW; ++uu) {       // skip the seen items in front of this slice (ascending list): bisection,
      int64_t lo = cur[uu], hi = end[uu];        // heavy users have 10^4 .. 10^5 of them and there are hundreds of slices
      while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (a.seen_indices[mid] < j_begin) lo = mid + 1; else hi = mid;
      }
      cur[uu] = lo;
    }
  }

  for (int64_t j0 = j_begin; j0 < j_end; j0 += SE_IT) {
    __syncthreads();  // previous tile fully consumed (and Ut visible on the first trip)
    for (int p = tid; p < SE_IT * (KP / 4); p += SE_THREADS) {
      const int i = p % SE_IT, c4 = p / SE_IT;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + i < a.n_items) x = *reinterpret_cast<const float4*>(a.vf + (j0 + i) * KP + 4 * c4);
      Vt[(4 * c4 + 0) * SE_IT + i] = x.x;
      Vt[(4 * c4 + 1) * SE_IT + i] = x.y;
      Vt[(4 * c4 + 2) * SE_IT + i] = x.z;
      Vt[(4 * c4 + 3) * SE_IT + i] = x.w;
    }
    __syncthreads();
    double acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
#pragma unroll 4
    for (int c = 0; c < KP; ++c) {
      const double2 u01 = *reinterpret_cast<const double2*>(Ut + c * SE_UT + 4 * tu);
      const double2 u23 = *reinterpret_cast<const double2*>(Ut + c * SE_UT + 4 * tu + 2);
      const double2 v01 = *reinterpret_cast<const double2*>(Vt + c * SE_IT + 4 * ti);
      const double2 v23 = *reinterpret_cast<const double2*>(Vt + c * SE_IT + 4 * ti + 2);
      const double ux[4] = {u01.x, u01.y, u23.x, u23.y};
      const double vx[4] = {v01.x, v01.y, v23.x, v23.y};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fma(ux[x], vx[y], acc[x][y]);
    }
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int64_t j = j0 + 4 * ti + y;
      const double ibv = (a.ibias && j < a.n_items) ? static_cast<double>(a.ibias[j]) : 0.0;
#pragma unroll
      for (int x = 0; x < 4; ++x)
        S[(4 * tu + x) * (SE_IT + 1) + 4 * ti + y] = (j < a.n_items) ? acc[x][y] + ibv + a.gbias : -CUDART_INF;
    }
    __syncthreads();

    // per-warp: mask the seen items of its 8 users, then offer the row to the user's list
#pragma unroll
    for (int uu = 0; uu < SE_UPW; ++uu) {
      const int ul = warp * SE_UPW + uu;
      if (t0 + ul >= total) continue;
      double* Srow = S + ul * (SE_IT + 1);
      for (;;) {
        const int64_t p = cur[uu] + lane;
        const int idx = (p < end[uu]) ? a.seen_indices[p] : INT_MAX;
        const bool in = idx < j0 + SE_IT;
        if (in) Srow[idx - j0] = -CUDART_INF;
        const unsigned b = __ballot_sync(FULL_MASK, in);
        cur[uu] += __popc(b);
        if (b != FULL_MASK) break;
      }
      __syncwarp();
      double tv;
      int tix;
      lists[uu].threshold(a.n, tv, tix);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int col = lane + 32 * half;
        const double sv = Srow[col] + ((Srow[col] == -CUDART_INF) ? 0.0 : ub[uu]);
        const int jj = static_cast<int>(j0) + col;
        unsigned m = __ballot_sync(FULL_MASK, sv != -CUDART_INF && beats(sv, jj, tv, tix));
        while (m) {
          const int l = __ffs(m) - 1;
          m &= m - 1;
          const double cs = __shfl_sync(FULL_MASK, sv, l);
          const int cj = static_cast<int>(j0) + l + 32 * half;
          if (beats(cs, cj, tv, tix)) {
            lists[uu].insert(cs, cj, lane);
            lists[uu].threshold(a.n, tv, tix);
          }
        }
      }
    }
  }
#pragma unroll
  for (int uu = 0; uu < SE_UPW; ++uu) {
    const int64_t t = t0 + warp * SE_UPW + uu;
    if (t < total) {
      if (a.slices > 1) {
        const int64_t prow = t * a.slices + blockIdx.y;
        lists[uu].store(a.n, lane, a.part_idx + prow * a.n, a.part_score + prow * a.n);
      } else {
        const int64_t orow = (a.user_list && a.scatter) ? a.user_list[t] : t;
        lists[uu].store(a.n, lane, a.idx_out + orow * a.n, a.score_out ? a.score_out + orow * a.n : nullptr);
      }
    }
  }
}

// merge of the per-slice lists: one warp per user, same order contract (score desc, index asc)
template <int R>
__global__ void __launch_bounds__(128) score_merge_kernel(const ScoreArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t total = a.user_list ? a.n_list : a.n_users;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (t >= total) return;
  WarpTopK<R> list;
  list.init();
  double tv;
  int tix;
  list.threshold(a.n, tv, tix);
  const int64_t cand = static_cast<int64_t>(a.slices) * a.n;
  for (int64_t c0 = 0; c0 < cand; c0 += 32) {
    const int64_t c = c0 + lane;
    const int id = (c < cand) ? a.part_idx[t * cand + c] : -1;
    const double sc = (c < cand) ? a.part_score[t * cand + c] : -CUDART_INF;
    unsigned m = __ballot_sync(FULL_MASK, id >= 0 && beats(sc, id, tv, tix));
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const double cs = __shfl_sync(FULL_MASK, sc, l);
      const int cj = __shfl_sync(FULL_MASK, id, l);
      if (beats(cs, cj, tv, tix)) {
        list.insert(cs, cj, lane);
        list.threshold(a.n, tv, tix);
      }
    }
  }
  const int64_t orow = (a.user_list && a.scatter) ? a.user_list[t] : t;
  list.store(a.n, lane, a.idx_out + orow * a.n, a.score_out ? a.score_out + orow * a.n : nullptr);
}

template <int KP, int R>
int launch_score_exact_kr(rl_context* h, const ScoreArgs& a_in) {
  ScoreArgs a = a_in;
  auto kern = score_topn_exact_kernel<KP, R>;
  constexpr size_t smem = score_exact_smem<KP>();
  RL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t total = a.user_list ? a.n_list : a.n_users;
  const int64_t grid = (total + SE_UT - 1) / SE_UT;
  if (grid == 0) return RL_OK;
  // few user groups and many items: slice the item range so that the whole chip works on them, then merge
  const int64_t item_tiles = (a.n_items + SE_IT - 1) / SE_IT;
  void* part = nullptr;
  if (grid * 2 <= h->sm_count && item_tiles >= 16) {
    int64_t slices = (2 * static_cast<int64_t>(h->sm_count) + grid - 1) / grid;
    if (slices > item_tiles / 4) slices = item_tiles / 4;
    if (slices > 512) slices = 512;
    const int64_t tiles_per = (item_tiles + slices - 1) / slices;
    a.slice_items = tiles_per * SE_IT;
    a.slices = static_cast<int>((item_tiles + tiles_per - 1) / tiles_per);
    const size_t rows = static_cast<size_t>(total) * a.slices * a.n;
    const size_t idx_bytes = (rows * sizeof(int32_t) + 15) & ~size_t(15);
    RL_CUDA(h, cudaMallocAsync(&part, idx_bytes + rows * sizeof(double), h->stream));
    a.part_idx = static_cast<int32_t*>(part);
    a.part_score = reinterpret_cast<double*>(static_cast<char*>(part) + idx_bytes);
  }
  kern<<<dim3(static_cast<unsigned>(grid), static_cast<unsigned>(a.slices)), SE_THREADS, smem, h->stream>>>(a);
  RL_LAUNCH_CHECK(h, "score_topn_exact_kernel");
  if (a.slices > 1) {
    score_merge_kernel<R><<<static_cast<unsigned>((total + 3) / 4), 128, 0, h->stream>>>(a);
    RL_LAUNCH_CHECK(h, "score_merge_kernel");
    RL_CUDA(h, cudaFreeAsync(part, h->stream));
  }
  return RL_OK;
}

template <int KP>
int launch_score_exact_k(rl_context* h, const ScoreArgs& a) {
  if (a.n <= 32) return launch_score_exact_kr<KP, 1>(h, a);
  if (a.n <= 64) return launch_score_exact_kr<KP, 2>(h, a);
  if (a.n <= 128) return launch_score_exact_kr<KP, 4>(h, a);
  return fail(h, RL_ERR_INVALID, "score_topn: n=%d exceeds the on-chip list capacity of 128", a.n);
}

// Force the (lazily loaded) fallback kernels of the tensor-core path into the context: their first use must not land
// inside somebody's timed step as a multi-millisecond module load.
template <int KP>
inline void preload_score_exact_k() {
  cudaFuncAttributes at;
  cudaFuncGetAttributes(&at, score_topn_exact_kernel<KP, 1>);
  cudaFuncGetAttributes(&at, score_merge_kernel<1>);
}
inline void preload_score_exact(int kp) {
  switch (kp) {
    case 16: preload_score_exact_k<16>(); break;
    case 32: preload_score_exact_k<32>(); break;
    case 64: preload_score_exact_k<64>(); break;
    case 128: preload_score_exact_k<128>(); break;
  }
}

inline int launch_score_exact(rl_context* h, int kp, const ScoreArgs& a) {
  switch (kp) {
    case 16: return launch_score_exact_k<16>(h, a);
    case 32: return launch_score_exact_k<32>(h, a);
    case 64: return launch_score_exact_k<64>(h, a);
    case 128: return launch_score_exact_k<128>(h, a);
  }
  return fail(h, RL_ERR_INVALID, "score_topn: unsupported padded k %d", kp);
}

}  // namespace rl
