// csr.cu -- observation pattern on the device: CSR -> CSR of the transpose (rows = items), ascending inside each row.
//
// The reference finds the users of an item by scanning a dense column (`ratings[:, i] > 0`, algo.py:325); the kernels
// take the transposed pattern as a second CSR instead.  Building it is a stable sort of the (column, row) pairs by
// column: rows are expanded from indptr, sorted with cub::DeviceRadixSort (stable, so rows stay ascending inside a
// column), column counts are a histogram + prefix sum.  Format conversion, not arithmetic of the path: library sort.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace rl {

namespace {

__global__ void __launch_bounds__(256) expand_rows_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, int32_t* __restrict__ rows) {
  // one warp per row: writes the row id over the row's range
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = w0; r < n_rows; r += nw) {
    const int64_t s = indptr[r], e = indptr[r + 1];
    for (int64_t p = s + lane; p < e; p += 32) rows[p] = static_cast<int32_t>(r);
  }
}

__global__ void __launch_bounds__(256) count_cols_kernel(const int32_t* __restrict__ indices, int64_t nnz, unsigned long long* __restrict__ counts) {
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < nnz; p += static_cast<int64_t>(gridDim.x) * blockDim.x)
    atomicAdd(counts + indices[p], 1ull);
}

}  // namespace

// indptr_t (n_cols + 1, int64) and indices_t (nnz, int32) are device outputs; enqueued on h->stream, temporaries from the
// stream-ordered pool.
int launch_csr_transpose(rl_context* h, int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                         int64_t* indptr_t, int32_t* indices_t) {
  if (n_rows < 0 || n_cols < 0 || nnz < 0 || n_rows >= (1ll << 31) || n_cols >= (1ll << 31) || nnz >= (1ll << 31))
    return fail(h, RL_ERR_INVALID, "csr_transpose: sizes must fit 31 bits");
  if (!indptr || !indptr_t || (nnz > 0 && (!indices || !indices_t))) return fail(h, RL_ERR_INVALID, "csr_transpose: null pointer");
  RL_CUDA(h, cudaMemsetAsync(indptr_t, 0, sizeof(int64_t) * (n_cols + 1), h->stream));
  if (nnz == 0) return RL_OK;
  int32_t *rows = nullptr, *keys_out = nullptr;
  void* tmp = nullptr;
  size_t tmp_sort = 0, tmp_scan = 0;
  int end_bit = 1;
  while ((1ll << end_bit) < n_cols) ++end_bit;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, indices, keys_out, rows, indices_t, static_cast<int>(nnz), 0, end_bit, h->stream);
  cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, indptr_t + 1, indptr_t + 1, static_cast<int>(n_cols), h->stream);
  const size_t tmp_bytes = tmp_sort > tmp_scan ? tmp_sort : tmp_scan;
  RL_CUDA(h, cudaMallocAsync(&rows, sizeof(int32_t) * nnz, h->stream));
  RL_CUDA(h, cudaMallocAsync(&keys_out, sizeof(int32_t) * nnz, h->stream));
  RL_CUDA(h, cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, h->stream));
  const int grid = 8 * h->sm_count;
  expand_rows_kernel<<<grid, 256, 0, h->stream>>>(indptr, n_rows, rows);
  RL_LAUNCH_CHECK(h, "expand_rows_kernel");
  count_cols_kernel<<<grid, 256, 0, h->stream>>>(indices, nnz, reinterpret_cast<unsigned long long*>(indptr_t + 1));
  RL_LAUNCH_CHECK(h, "count_cols_kernel");
  size_t t1 = tmp_bytes;
  RL_CUDA(h, cub::DeviceScan::InclusiveSum(tmp, t1, indptr_t + 1, indptr_t + 1, static_cast<int>(n_cols), h->stream));
  size_t t2 = tmp_bytes;
  RL_CUDA(h, cub::DeviceRadixSort::SortPairs(tmp, t2, indices, keys_out, rows, indices_t, static_cast<int>(nnz), 0, end_bit, h->stream));
  h->launches += 2;
  RL_CUDA(h, cudaFreeAsync(tmp, h->stream));
  RL_CUDA(h, cudaFreeAsync(keys_out, h->stream));
  RL_CUDA(h, cudaFreeAsync(rows, h->stream));
  return RL_OK;
}

}  // namespace rl
