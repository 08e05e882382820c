// topk.cuh -- warp-held sorted top-n list (n <= 32*R), shared by the scoring and dense top-n kernels.
//
// Order contract (reference algo.py:579-582 sorts descending; ties are unspecified there):
// descending score, and among equal scores the LOWER item index first.  Empty slots hold (-inf, INT_MAX).
#pragma once

#include <cuda_runtime.h>
#include <limits.h>
#include <math_constants.h>

namespace rl {

__device__ __forceinline__ bool beats(double a, int ai, double b, int bi) {
  return a > b || (a == b && ai < bi);
}

template <int R>
struct WarpTopK {
  double v[R];
  int id[R];

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int r = 0; r < R; ++r) { v[r] = -CUDART_INF; id[r] = INT_MAX; }
  }

  // value / index of list position n-1 (the entry a candidate has to beat once the list is full)
  __device__ __forceinline__ void threshold(int n, double& tv, int& ti) const {
    const int pos = n - 1, src = pos / R, slot = pos % R;
    double x = v[0];
    int y = id[0];
#pragma unroll
    for (int r = 1; r < R; ++r)
      if (r == slot) { x = v[r]; y = id[r]; }
    tv = __shfl_sync(0xffffffffu, x, src);
    ti = __shfl_sync(0xffffffffu, y, src);
  }

  // insert (s, j); all 32 lanes must call with the same (s, j)
  __device__ __forceinline__ void insert(double s, int j, int lane) {
    int p = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) p += __popc(__ballot_sync(0xffffffffu, beats(v[r], id[r], s, j)));
    const double cv = __shfl_up_sync(0xffffffffu, v[R - 1], 1);
    const int ci = __shfl_up_sync(0xffffffffu, id[R - 1], 1);
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
      const int pos = lane * R + r;
      if (pos > p) {
        v[r] = (r > 0) ? v[r > 0 ? r - 1 : 0] : cv;
        id[r] = (r > 0) ? id[r > 0 ? r - 1 : 0] : ci;
      } else if (pos == p) {
        v[r] = s;
        id[r] = j;
      }
    }
  }

  // write the first n entries: idx (-1 for empty) and optionally the scores
  __device__ __forceinline__ void store(int n, int lane, int32_t* idx_out, double* score_out) const {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int pos = lane * R + r;
      if (pos < n) {
        idx_out[pos] = (id[r] == INT_MAX) ? -1 : id[r];
        if (score_out) score_out[pos] = v[r];
      }
    }
  }
};

}  // namespace rl
