// svd.cu -- truncated SVD factors through the Gram matrix of the shorter side, and the low-rank product.
//
// Replaces np.linalg.svd + slicing + `u @ (s * vt)` of the reference (algo.py:510-517, :238-243, :363-368).
//   1. G = A^T A in float64 (A = M if m >= n else M^T), split over row slabs with float64 atomics;
//   2. eigenvectors of G by one-sided cyclic Jacobi (Hestenes) on the columns of W = G, V accumulated: every
//      round rotates n/2 disjoint column pairs (round-robin tournament), one CTA per pair; sweeps repeat until
//      a sweep performs no rotation;  eigenvalue_j = |w_j|, singular value = sqrt(eigenvalue_j);
//   3. top-k columns by eigenvalue -> right (n x k); left = A right (m x k), all float64;
//   4. rl_lowrank: out = left right^T (+ bias terms), accumulated in float64, rounded once to the output type.
// C2 (10 000 x 500, k = 10) is a parity configuration: it is latency bound (about 5 000 tiny launches), not a
// roofline configuration.
#include "common.cuh"

namespace rl {

namespace {

__global__ void transpose_f32_kernel(const float* __restrict__ in, int64_t rows, int64_t cols, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 32, r0 = static_cast<int64_t>(blockIdx.y) * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][i];
  }
}

// G (n x n, row-major == column-major since symmetric) += A[slab]^T A[slab]; lower tiles only, 64 x 64 per CTA
__global__ void __launch_bounds__(256) gram64_kernel(const float* __restrict__ a, int64_t m, int n, int64_t slab,
                                                     double* __restrict__ g) {
  __shared__ double Ai[32][64];
  __shared__ double Aj[32][64];
  int t = blockIdx.x, bi = 0;  // lower-triangular tile index -> (bi, bj)
  while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
  const int bj = t - bi * (bi + 1) / 2;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.y) * slab;
  const int64_t r_end = min(m, r_begin + slab);
  double acc[4][4];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += 32) {
    for (int p = threadIdx.x; p < 32 * 64; p += 256) {
      const int rr = p >> 6, c = p & 63;
      const int64_t r = r0 + rr;
      const int ci = bi * 64 + c, cj = bj * 64 + c;
      Ai[rr][c] = (r < r_end && ci < n) ? static_cast<double>(a[r * n + ci]) : 0.0;
      Aj[rr][c] = (r < r_end && cj < n) ? static_cast<double>(a[r * n + cj]) : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
      double xa[4], xb[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) { xa[x] = Ai[rr][4 * ty + x]; xb[x] = Aj[rr][4 * tx + x]; }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fma(xa[x], xb[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int gi = bi * 64 + 4 * ty + x, gj = bj * 64 + 4 * tx + y;
      if (gi < n && gj < n && gj <= gi) atomicAdd(&g[static_cast<int64_t>(gi) * n + gj], acc[x][y]);
    }
}

// mirror the lower triangle, V = I, trace -> scal[0]
__global__ void jacobi_init_kernel(double* __restrict__ w, double* __restrict__ v, int n, double* __restrict__ scal) {
  const int64_t total = static_cast<int64_t>(n) * n;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < total;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(p / n), j = static_cast<int>(p % n);
    if (j > i) w[p] = w[static_cast<int64_t>(j) * n + i];
    v[p] = (i == j) ? 1.0 : 0.0;
    if (i == j) atomicAdd(scal, w[p]);
  }
}

__device__ __forceinline__ double block_sum3(double& a, double& b, double& c, double* red) {
  // reduces three values over a 256-thread block; results broadcast to all threads
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(FULL_MASK, a, o);
    b += __shfl_xor_sync(FULL_MASK, b, o);
    c += __shfl_xor_sync(FULL_MASK, c, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[warp * 3 + 0] = a; red[warp * 3 + 1] = b; red[warp * 3 + 2] = c; }
  __syncthreads();
  a = b = c = 0.0;
  for (int wi = 0; wi < 8; ++wi) { a += red[wi * 3]; b += red[wi * 3 + 1]; c += red[wi * 3 + 2]; }
  return a;
}

// one tournament round: CTA i rotates columns (p, q) of W and V (column-major, ld = n)
__global__ void __launch_bounds__(256) jacobi_round_kernel(double* __restrict__ w, double* __restrict__ v, int n, int np,
                                                           int round, double tol, const double* __restrict__ scal,
                                                           unsigned int* __restrict__ rotations) {
  __shared__ double red[24];
  const int i = blockIdx.x;
  int p, q;
  if (i == 0) { p = np - 1; q = round; }
  else { p = (round + i) % (np - 1); q = (round - i + (np - 1)) % (np - 1); }
  if (p >= n || q >= n) return;  // phantom column when n is odd
  if (p > q) { const int t = p; p = q; q = t; }
  double* wp = w + static_cast<int64_t>(p) * n;
  double* wq = w + static_cast<int64_t>(q) * n;
  double* vp = v + static_cast<int64_t>(p) * n;
  double* vq = v + static_cast<int64_t>(q) * n;
  double al = 0.0, be = 0.0, ga = 0.0;
  if (n <= 1024) {
    // Columns of up to 1024 entries live in registers for the whole round: all four columns are requested at once (one
    // L2 round trip instead of two in a row around the reduction -- the round is a chain of latencies, not of work).
    double x[4], y[4], c4[4], d4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = threadIdx.x + 256 * e;
      const bool in = r < n;
      x[e] = in ? wp[r] : 0.0; y[e] = in ? wq[r] : 0.0;
      c4[e] = in ? vp[r] : 0.0; d4[e] = in ? vq[r] : 0.0;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { al = fma(x[e], x[e], al); be = fma(y[e], y[e], be); ga = fma(x[e], y[e], ga); }
    block_sum3(al, be, ga, red);
    const double floor2r = (1e-14 * scal[0]) * (1e-14 * scal[0]);
    if (al <= floor2r || be <= floor2r) return;
    if (fabs(ga) <= tol * sqrt(al * be)) return;
    const double zeta = (be - al) / (2.0 * ga);
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = threadIdx.x + 256 * e;
      if (r < n) {
        wp[r] = c * x[e] - sn * y[e];
        wq[r] = sn * x[e] + c * y[e];
        vp[r] = c * c4[e] - sn * d4[e];
        vq[r] = sn * c4[e] + c * d4[e];
      }
    }
    if (threadIdx.x == 0) atomicAdd(rotations, 1u);
    return;
  }
  for (int r = threadIdx.x; r < n; r += 256) {
    const double x = wp[r], y = wq[r];
    al = fma(x, x, al); be = fma(y, y, be); ga = fma(x, y, ga);
  }
  block_sum3(al, be, ga, red);
  // columns whose norm is rounding noise relative to trace(G) carry no direction: leave them alone
  const double floor2 = (1e-14 * scal[0]) * (1e-14 * scal[0]);
  if (al <= floor2 || be <= floor2) return;
  if (fabs(ga) <= tol * sqrt(al * be)) return;
  const double zeta = (be - al) / (2.0 * ga);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
  for (int r = threadIdx.x; r < n; r += 256) {
    const double x = wp[r], y = wq[r];
    wp[r] = c * x - s * y;
    wq[r] = s * x + c * y;
    const double a = vp[r], b = vq[r];
    vp[r] = c * a - s * b;
    vq[r] = s * a + c * b;
  }
  if (threadIdx.x == 0) atomicAdd(rotations, 1u);
}

// eigenvalue_j = |w_j|; rank by (value desc, index asc); the k best columns -> right (n x k row-major), s
__global__ void __launch_bounds__(256) select_topk_kernel(const double* __restrict__ w, const double* __restrict__ v, int n,
                                                          int k, double* __restrict__ lam, double* __restrict__ s_out,
                                                          double* __restrict__ right) {
  // phase 1 (all blocks): column norms
  for (int j = blockIdx.x; j < n; j += gridDim.x) {
    __shared__ double red[24];
    double a = 0.0, b = 0.0, c = 0.0;
    const double* wj = w + static_cast<int64_t>(j) * n;
    for (int r = threadIdx.x; r < n; r += 256) a = fma(wj[r], wj[r], a);
    block_sum3(a, b, c, red);
    if (threadIdx.x == 0) lam[j] = sqrt(a);
    __syncthreads();
  }
  (void)v; (void)k; (void)s_out; (void)right;
}

__global__ void __launch_bounds__(256) gather_topk_kernel(const double* __restrict__ v, const double* __restrict__ lam, int n,
                                                          int k, double* __restrict__ s_out, double* __restrict__ right) {
  // one block per column j: rank of j among the eigenvalues; if rank < k copy the eigenvector
  const int j = blockIdx.x;
  __shared__ int rank_s;
  if (threadIdx.x == 0) rank_s = 0;
  __syncthreads();
  int cnt = 0;
  const double lj = lam[j];
  for (int i = threadIdx.x; i < n; i += 256) {
    const double li = lam[i];
    if (li > lj || (li == lj && i < j)) ++cnt;
  }
  atomicAdd(&rank_s, cnt);
  __syncthreads();
  const int rank = rank_s;
  if (rank >= k) return;
  if (threadIdx.x == 0 && s_out) s_out[rank] = sqrt(lj);
  const double* vj = v + static_cast<int64_t>(j) * n;
  for (int r = threadIdx.x; r < n; r += 256) right[static_cast<int64_t>(r) * k + rank] = vj[r];
}

// left (m x k) = A (m x n float32) * right (n x k float64)
__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ a, int64_t m, int n, int k,
                                                      const double* __restrict__ right, double* __restrict__ left) {
  const int64_t total = m * k;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < total;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = p / k;
    const int c = static_cast<int>(p - r * k);
    const float* ar = a + r * n;
    double acc = 0.0;
    for (int j = 0; j < n; ++j) acc = fma(static_cast<double>(ar[j]), right[static_cast<int64_t>(j) * k + c], acc);
    left[p] = acc;
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) lowrank_kernel(const TI* __restrict__ a, const TI* __restrict__ b, int lda, int64_t m,
                                                      int64_t n, int k, double gbias, const float* __restrict__ rbias,
                                                      const float* __restrict__ cbias, TO* __restrict__ out) {
  // 16 x 16 threads, each 4 x 4 outputs of a 64 x 64 tile; k is consumed in slices of 16 through shared memory
  __shared__ double As[16][64 + 1];
  __shared__ double Bs[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 64, c0 = static_cast<int64_t>(blockIdx.x) * 64;
  double acc[4][4];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
  for (int k0 = 0; k0 < k; k0 += 16) {
    for (int p = threadIdx.x; p < 64 * 16; p += 256) {
      const int rr = p >> 4, kk = p & 15;
      const int64_t r = r0 + rr, c = c0 + rr;
      As[kk][rr] = (r < m && k0 + kk < k) ? static_cast<double>(a[r * lda + k0 + kk]) : 0.0;
      Bs[kk][rr] = (c < n && k0 + kk < k) ? static_cast<double>(b[c * lda + k0 + kk]) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double xa[4], xb[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) { xa[x] = As[kk][4 * ty + x]; xb[x] = Bs[kk][4 * tx + x]; }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fma(xa[x], xb[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int64_t r = r0 + 4 * ty + x;
    if (r >= m) continue;
    const double rb = gbias + (rbias ? static_cast<double>(rbias[r]) : 0.0);
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int64_t c = c0 + 4 * tx + y;
      if (c < n) out[r * n + c] = static_cast<TO>(acc[x][y] + rb + (cbias ? static_cast<double>(cbias[c]) : 0.0));
    }
  }
}

}  // namespace

int launch_lowrank(rl_context* h, const void* a, const void* b, int in_dtype, int64_t m, int64_t n, int k, double gbias,
                   const float* rbias, const float* cbias, void* out, int out_dtype) {
  if (m < 0 || n < 0 || k <= 0) return fail(h, RL_ERR_INVALID, "lowrank: bad shape");
  if (m == 0 || n == 0) return RL_OK;
  if (!a || !b || !out) return fail(h, RL_ERR_INVALID, "lowrank: null pointer");
  const dim3 grid(static_cast<unsigned>((n + 63) / 64), static_cast<unsigned>((m + 63) / 64));
  if (grid.y > 65535) return fail(h, RL_ERR_INVALID, "lowrank: more than 65535*64 rows; tile the call");
  if (in_dtype == RL_F64) {
    const double* ad = static_cast<const double*>(a);
    const double* bd = static_cast<const double*>(b);
    if (out_dtype == RL_F32) lowrank_kernel<double, float><<<grid, 256, 0, h->stream>>>(ad, bd, k, m, n, k, gbias, rbias, cbias, static_cast<float*>(out));
    else if (out_dtype == RL_F64) lowrank_kernel<double, double><<<grid, 256, 0, h->stream>>>(ad, bd, k, m, n, k, gbias, rbias, cbias, static_cast<double*>(out));
    else return fail(h, RL_ERR_INVALID, "lowrank: bad out dtype");
  } else if (in_dtype == RL_F32) {
    const int kp = padded_k(k);
    if (kp == 0) return fail(h, RL_ERR_INVALID, "lowrank: k=%d outside [1, 128] for float32 factor matrices", k);
    const float* af = static_cast<const float*>(a);
    const float* bf = static_cast<const float*>(b);
    if (out_dtype == RL_F32) lowrank_kernel<float, float><<<grid, 256, 0, h->stream>>>(af, bf, kp, m, n, k, gbias, rbias, cbias, static_cast<float*>(out));
    else if (out_dtype == RL_F64) lowrank_kernel<float, double><<<grid, 256, 0, h->stream>>>(af, bf, kp, m, n, k, gbias, rbias, cbias, static_cast<double*>(out));
    else return fail(h, RL_ERR_INVALID, "lowrank: bad out dtype");
  } else {
    return fail(h, RL_ERR_INVALID, "lowrank: bad in dtype");
  }
  RL_LAUNCH_CHECK(h, "lowrank_kernel");
  return RL_OK;
}

int launch_svd_topk(rl_context* h, const float* m_dev, int64_t m, int64_t n, int k, double* s_out, double* left_out,
                    double* right_out) {
  if (m <= 0 || n <= 0) return fail(h, RL_ERR_INVALID, "svd_topk: zero dimension");
  const int64_t small = m < n ? m : n;
  if (k < 1 || k > small) return fail(h, RL_ERR_INVALID, "svd_topk: k=%d outside [1, min(m,n)=%lld]", k, static_cast<long long>(small));
  if (small > 4096) return fail(h, RL_ERR_INVALID, "svd_topk: min(m, n)=%lld exceeds 4096", static_cast<long long>(small));
  if (!m_dev || !left_out || !right_out) return fail(h, RL_ERR_INVALID, "svd_topk: null pointer");
  const bool wide = m < n;
  const int64_t rows = wide ? n : m;       // rows of A (the tall orientation)
  const int nn = static_cast<int>(small);  // columns of A
  // scratch: [A^T copy if wide] + W + V + lam + scal + rotations
  const size_t a_bytes = wide ? sizeof(float) * static_cast<size_t>(m) * n : 0;
  const size_t mat_bytes = sizeof(double) * static_cast<size_t>(nn) * nn;
  const size_t off_w = (a_bytes + 255) & ~size_t(255);
  const size_t off_v = off_w + mat_bytes;
  const size_t off_lam = off_v + mat_bytes;
  const size_t off_scal = off_lam + sizeof(double) * nn;
  const size_t off_rot = off_scal + 64;
  void* base = nullptr;
  int rc = scratch(h, off_rot + 256, &base);
  if (rc != RL_OK) return rc;
  unsigned char* bp = static_cast<unsigned char*>(base);
  const float* a = m_dev;
  if (wide) {
    float* at = reinterpret_cast<float*>(bp);
    const dim3 g(static_cast<unsigned>((n + 31) / 32), static_cast<unsigned>((m + 31) / 32));
    transpose_f32_kernel<<<g, dim3(32, 8), 0, h->stream>>>(m_dev, m, n, at);
    RL_LAUNCH_CHECK(h, "transpose_f32_kernel");
    a = at;
  }
  double* w = reinterpret_cast<double*>(bp + off_w);
  double* v = reinterpret_cast<double*>(bp + off_v);
  double* lam = reinterpret_cast<double*>(bp + off_lam);
  double* scal = reinterpret_cast<double*>(bp + off_scal);
  unsigned int* rot = reinterpret_cast<unsigned int*>(bp + off_rot);
  RL_CUDA(h, cudaMemsetAsync(w, 0, mat_bytes, h->stream));
  RL_CUDA(h, cudaMemsetAsync(scal, 0, 64, h->stream));
  {
    const int nt = (nn + 63) / 64;
    const int tiles = nt * (nt + 1) / 2;
    int splits = static_cast<int>((4 * h->sm_count + tiles - 1) / tiles);
    int64_t slab = (rows + splits - 1) / splits;
    slab = ((slab + 31) / 32) * 32;
    splits = static_cast<int>((rows + slab - 1) / slab);
    gram64_kernel<<<dim3(tiles, splits), 256, 0, h->stream>>>(a, rows, nn, slab, w);
    RL_LAUNCH_CHECK(h, "gram64_kernel");
  }
  jacobi_init_kernel<<<4 * h->sm_count, 256, 0, h->stream>>>(w, v, nn, scal);
  RL_LAUNCH_CHECK(h, "jacobi_init_kernel");
  if (nn > 1) {
    const int np = (nn + 1) & ~1;
    const double tol = 1e-15;
    // One sweep = np - 1 dependent launches of a few microseconds each: launch-bound.  The sweep is captured once into
    // a CUDA graph (counter reset + all rounds) and replayed until a sweep rotates nothing; small problems and a
    // failed capture take the plain launches.
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const bool capturing = np - 1 >= 32 && cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (!capturing) (void)cudaGetLastError();   // e.g. the legacy default stream cannot be captured: plain launches
    if (capturing) {
      bool ok = cudaMemsetAsync(rot, 0, sizeof(unsigned int), h->stream) == cudaSuccess;
      for (int round = 0; ok && round < np - 1; ++round) {
        jacobi_round_kernel<<<np / 2, 256, 0, h->stream>>>(w, v, nn, np, round, tol, scal, rot);
        ok = cudaGetLastError() == cudaSuccess;
      }
      if (cudaStreamEndCapture(h->stream, &graph) != cudaSuccess || !ok || !graph ||
          cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        graph = nullptr;
        exec = nullptr;
        (void)cudaGetLastError();
      }
    }
    int rc_sweeps = RL_OK;
    for (int sweep = 0; sweep < 60; ++sweep) {
      cudaError_t e = cudaSuccess;
      if (exec) {
        e = cudaGraphLaunch(exec, h->stream);
        if (e == cudaSuccess) h->launches += np - 1;
      } else {
        e = cudaMemsetAsync(rot, 0, sizeof(unsigned int), h->stream);
        for (int round = 0; e == cudaSuccess && round < np - 1; ++round) {
          jacobi_round_kernel<<<np / 2, 256, 0, h->stream>>>(w, v, nn, np, round, tol, scal, rot);
          e = cudaGetLastError();
          if (e == cudaSuccess) h->launches++;
        }
      }
      unsigned int done = 0;
      if (e == cudaSuccess) e = cudaMemcpyAsync(&done, rot, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) { rc_sweeps = cuda_fail(h, e, "jacobi sweep"); break; }
      if (done == 0) break;
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (rc_sweeps != RL_OK) return rc_sweeps;
  }
  select_topk_kernel<<<2 * h->sm_count, 256, 0, h->stream>>>(w, v, nn, k, lam, nullptr, nullptr);
  RL_LAUNCH_CHECK(h, "select_topk_kernel");
  // the orthonormal factor goes to `right` when m >= n, to `left` when m < n
  double* ortho = wide ? left_out : right_out;
  double* proj = wide ? right_out : left_out;
  gather_topk_kernel<<<nn, 256, 0, h->stream>>>(v, lam, nn, k, s_out, ortho);
  RL_LAUNCH_CHECK(h, "gather_topk_kernel");
  {
    int64_t grid = (rows * k + 255) / 256;
    if (grid > 32 * h->sm_count) grid = 32 * h->sm_count;
    project_kernel<<<static_cast<unsigned>(grid), 256, 0, h->stream>>>(a, rows, nn, k, ortho, proj);
    RL_LAUNCH_CHECK(h, "project_kernel");
  }
  return RL_OK;
}

}  // namespace rl
