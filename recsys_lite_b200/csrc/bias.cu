// bias.cu -- bias stage of the SVD class path on the device.
//
// Replaces the Python list comprehensions of RecommenderSystem._fit_svd_dense (reference algo.py:218-235), 80 % of the
// reference's fit time at the 10 K x 500 configuration:
//     global_bias = mean(R[R > 0]);  user_bias[u] = mean(R[u][R[u] > 0]) - global_bias (0 without positive entry);
//     item_bias likewise per column;  centred = R - global_bias - user_bias[:, None] - item_bias   (every cell).
// Three streaming passes over R (HBM bound: 2 reads + 1 read/write of m n floats): masked row sums / counts (one warp
// per row, coalesced), masked column sums / counts (a CTA per 32-column strip over a row slab, float64 atomics), then
// the centring.  Sums are float64; the means are rounded to the caller's float type on the host side of the ABI.
#include <math_constants.h>

#include "common.cuh"

namespace rl {

namespace {

// one warp per row: sum and count of the entries > 0
__global__ void __launch_bounds__(256) masked_row_sums_kernel(const float* __restrict__ r, int64_t m, int64_t n, double* __restrict__ row_sum,
                                                              int* __restrict__ row_cnt, double* __restrict__ tot /* [sum, count] */) {
  const int lane = threadIdx.x & 31;
  double bs = 0.0, bc = 0.0;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < m; row += static_cast<int64_t>(gridDim.x) * 8) {
    const float* p = r + row * n;
    double s = 0.0;
    int c = 0;
    for (int64_t j = lane; j < n; j += 32) {
      const float x = p[j];
      if (x > 0.f) { s += static_cast<double>(x); ++c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(FULL_MASK, s, o);
      c += __shfl_xor_sync(FULL_MASK, c, o);
    }
    if (lane == 0) { row_sum[row] = s; row_cnt[row] = c; bs += s; bc += c; }
  }
  if (lane == 0 && (bs != 0.0 || bc != 0.0)) { atomicAdd(tot, bs); atomicAdd(tot + 1, bc); }
}

// CTA (x, y): columns [32 x, 32 x + 32) of the row slab y; thread (tx, ty) walks rows ty, ty + 8, ... of the slab
__global__ void __launch_bounds__(256) masked_col_sums_kernel(const float* __restrict__ r, int64_t m, int64_t n, int64_t slab,
                                                              double* __restrict__ col_sum, int* __restrict__ col_cnt) {
  __shared__ double ss[8][32];
  __shared__ int sc[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t j = static_cast<int64_t>(blockIdx.x) * 32 + tx;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * slab, r1 = min(m, r0 + slab);
  double s = 0.0;
  int c = 0;
  if (j < n)
    for (int64_t i = r0 + ty; i < r1; i += 8) {
      const float x = r[i * n + j];
      if (x > 0.f) { s += static_cast<double>(x); ++c; }
    }
  ss[ty][tx] = s;
  sc[ty][tx] = c;
  __syncthreads();
  if (ty == 0 && j < n) {
#pragma unroll
    for (int t = 1; t < 8; ++t) { s += ss[t][tx]; c += sc[t][tx]; }
    if (c) { atomicAdd(col_sum + j, s); atomicAdd(col_cnt + j, c); }
  }
}

// biases from the sums (float64), written as float32 for the centring and as float64 for the caller
__global__ void __launch_bounds__(256) bias_from_sums_kernel(const double* __restrict__ sum, const int* __restrict__ cnt, int64_t len,
                                                             const double* __restrict__ tot, int as_float32, double* __restrict__ bias64,
                                                             float* __restrict__ bias32) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= len) return;
  // global bias in the caller's float type (the reference subtracts the rounded value, algo.py:218-234)
  double gb = tot[1] > 0.0 ? tot[0] / tot[1] : CUDART_NAN;
  if (as_float32) gb = static_cast<double>(static_cast<float>(gb));
  double b = cnt[i] > 0 ? sum[i] / static_cast<double>(cnt[i]) - gb : 0.0;
  if (as_float32) b = static_cast<double>(static_cast<float>(b));
  bias64[i] = b;
  bias32[i] = static_cast<float>(b);
}

__global__ void __launch_bounds__(256) centre_kernel(const float* __restrict__ r, int64_t m, int64_t n, const double* __restrict__ tot,
                                                     int as_float32, const double* __restrict__ ub, const double* __restrict__ ib,
                                                     float* __restrict__ out) {
  double gb = tot[1] > 0.0 ? tot[0] / tot[1] : CUDART_NAN;
  if (as_float32) gb = static_cast<double>(static_cast<float>(gb));
  const int64_t total = m * n;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = p / n, j = p - i * n;
    // the reference evaluates R - gb - ub[:, None] - ib left to right in the matrix's float type, then casts to float32
    if (as_float32) {
      const float x = ((r[p] - static_cast<float>(gb)) - static_cast<float>(ub[i])) - static_cast<float>(ib[j]);
      out[p] = x;
    } else {
      out[p] = static_cast<float>(((static_cast<double>(r[p]) - gb) - ub[i]) - ib[j]);
    }
  }
}

}  // namespace

// r_dev (m x n) float32 -> global bias (tot_dev[0] / tot_dev[1]), user_bias (m) / item_bias (n) float64, centred (m x n) float32.
// as_float32: round the means to float32 as the reference does for float32 input.  work: m + n doubles + m + n ints + 2 doubles.
int launch_bias_centre(rl_context* h, const float* r, int64_t m, int64_t n, int as_float32, double* tot, double* ub, double* ib,
                       float* centred) {
  if (m <= 0 || n <= 0 || !r || !tot || !ub || !ib) return fail(h, RL_ERR_INVALID, "bias_centre: bad argument");
  void* w = nullptr;
  const size_t wb = sizeof(double) * (m + n) + sizeof(int) * (m + n) + sizeof(float) * (m + n) + 64;
  RL_CUDA(h, cudaMallocAsync(&w, wb, h->stream));
  RL_CUDA(h, cudaMemsetAsync(w, 0, wb, h->stream));
  RL_CUDA(h, cudaMemsetAsync(tot, 0, 2 * sizeof(double), h->stream));
  double* rs = static_cast<double*>(w);
  double* cs = rs + m;
  int* rc = reinterpret_cast<int*>(cs + n);
  int* cc = rc + m;
  float* b32 = reinterpret_cast<float*>(cc + n);
  {
    int64_t g = (m + 7) / 8;
    if (g > 16 * h->sm_count) g = 16 * h->sm_count;
    masked_row_sums_kernel<<<static_cast<unsigned>(g), 256, 0, h->stream>>>(r, m, n, rs, rc, tot);
    RL_LAUNCH_CHECK(h, "masked_row_sums_kernel");
  }
  {
    const int64_t strips = (n + 31) / 32;
    int64_t slabs = (4 * static_cast<int64_t>(h->sm_count) + strips - 1) / strips;
    if (slabs > (m + 63) / 64) slabs = (m + 63) / 64;
    if (slabs < 1) slabs = 1;
    if (slabs > 65535) slabs = 65535;
    const int64_t slab = (m + slabs - 1) / slabs;
    masked_col_sums_kernel<<<dim3(static_cast<unsigned>(strips), static_cast<unsigned>((m + slab - 1) / slab)), 256, 0, h->stream>>>(r, m, n, slab, cs, cc);
    RL_LAUNCH_CHECK(h, "masked_col_sums_kernel");
  }
  bias_from_sums_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, h->stream>>>(rs, rc, m, tot, as_float32, ub, b32);
  RL_LAUNCH_CHECK(h, "bias_from_sums_kernel");
  bias_from_sums_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, h->stream>>>(cs, cc, n, tot, as_float32, ib, b32 + m);
  RL_LAUNCH_CHECK(h, "bias_from_sums_kernel");
  if (centred) {
    int64_t g = (m * n + 255) / 256;
    if (g > 16 * h->sm_count) g = 16 * h->sm_count;
    centre_kernel<<<static_cast<unsigned>(g), 256, 0, h->stream>>>(r, m, n, tot, as_float32, ub, ib, centred);
    RL_LAUNCH_CHECK(h, "centre_kernel");
  }
  RL_CUDA(h, cudaFreeAsync(w, h->stream));
  return RL_OK;
}

}  // namespace rl
