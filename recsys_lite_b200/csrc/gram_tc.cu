// gram_tc.cu -- dense Gram / cross-product  G = X^T Y  on the 5th-generation tensor cores (tcgen05, TMEM, TMA).
//
// north_star kernel (1), dense part: `item_f.T @ item_f` of the weighted ALS normal equations (the w0 Y^T Y term the
// reference's binary mode does not have, reference algo.py:320,330 for the per-row analogue), `M^T M` style products, and
// the three co-rating products of the KNN item-item cosine (reference algo.py:339-350: sum r_ui r_uj, sum r_ui^2 [r_uj > 0],
// number of co-raters) -- all of them "tall matrix transposed times tall matrix", contraction over the LONG dimension n.
//
// X (n x p), Y (n x q) float32 row-major.  The tensor core wants both operands K-major (contraction index contiguous), so
// a first kernel writes X^T and Y^T as fp16 (hi, lo) pairs: x s = hi + lo + r, |r| <= 2^-22 |x s|, s a power of two
// (transpose_split_kernel, 32 x 32 tiles through shared memory).  The GEMM kernel then runs, per CTA, one 128 x 128 output
// tile over a slice of n:
//   warp 0   TMA producer: (X^T hi, lo) and (Y^T hi, lo) tiles of 128 rows x 64 n-values (128B swizzle) into a 3-slot ring;
//   warp 1   MMA issuer: tcgen05.mma.kind::f16 with BOTH operands from shared memory, M = N = 128, K = 16:
//            hi.hi + lo.hi + hi.lo (TERMS = 3; fp16 x fp16 is exact in the fp32 accumulator) or hi.hi alone (TERMS = 1: exact
//            for small integers, e.g. ratings 1..5 and 0/1 masks);
//   warp 2   TMEM allocation;
//   warps 4-7 epilogue: every 32 chunks (2048 values of n -- the fp32 accumulator must not run longer than that) the
//            accumulator is read back with tcgen05.ld, unscaled and added to the float64 result with atomics; the slices of n
//            of different CTAs meet in the same atomics (split-K).
// Accuracy: TERMS = 3 ~ 1e-6 relative to sum |x||y| (checked against float64 in tests/test_gpu_parity.py).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05_ptx.cuh"

namespace rl {

namespace {

using namespace tc;

constexpr int G_T = 128;            // output tile (rows of X^T x rows of Y^T)
constexpr int G_KC = 64;            // n-values per chunk (one 128-byte swizzled row of fp16)
constexpr int G_DRAIN = 32;         // chunks per accumulator drain
constexpr int G_SLOTS = 3;
constexpr int G_THREADS = 256;
constexpr uint32_t G_TILE_BYTES = G_T * G_KC * 2;
constexpr uint32_t G_IDESC = (1u << 4) | (static_cast<uint32_t>(G_T >> 3) << 17) | (static_cast<uint32_t>(G_T >> 4) << 24);

__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// max |x| over an (n x p) matrix with row stride ld
__global__ void __launch_bounds__(256) absmax2_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int ld, float* __restrict__ out) {
  float m = 0.f;
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    m = fmaxf(m, fabsf(x[r * ld + (i - r * cols)]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}

// X (n x p, stride ld) -> hi, lo: (p_pad x n_pad) fp16, zero padded; op: 0 = x, 1 = x^2, 2 = [x > 0], 3 = x [x > 0];
// op | 8: X is handed over already transposed, (p x n) row-major with stride ld (the contraction index contiguous)
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ x, int64_t n, int p, int ld, int op,
                                                              const float* __restrict__ amax, int64_t n_pad, __half* __restrict__ hi,
                                                              __half* __restrict__ lo) {
  __shared__ float t[32][33];
  float sc = 1.0f;
  const bool pre_t = (op & 8) != 0;
  op &= 7;
  if (op == 0 || op == 3) sc = pow2_scale(*amax);
  else if (op == 1) { const float a = *amax; sc = pow2_scale(a * a); }
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int p0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t nn = n0 + r;
    const int pp = p0 + tx;
    float v = 0.f;
    if (nn < n && pp < p) {
      v = pre_t ? x[static_cast<int64_t>(pp) * ld + nn] : x[nn * ld + pp];
      if (op == 1) v = v > 0.f ? v * v : 0.f;       // (squares of the observed entries only: the co-rating sums)
      else if (op == 2) v = v > 0.f ? 1.f : 0.f;
      else if (op == 3) v = v > 0.f ? v : 0.f;
    }
    t[r][tx] = v * sc;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int pp = p0 + r;                     // output row
    const int64_t nn = n0 + tx;                // output column
    if (nn < n_pad) {
      const float v = t[tx][r];
      const __half h = __float2half_rn(v);
      hi[static_cast<int64_t>(pp) * n_pad + nn] = h;
      if (lo) lo[static_cast<int64_t>(pp) * n_pad + nn] = __float2half_rn(v - __half2float(h));
    }
  }
}

struct GramArgs {
  double* g;            // (p x q) float64, row stride ldg; the kernel ADDS into it
  int ldg, p, q;
  int n_chunks;         // n_pad / 64
  int chunks_per_cta;
  const float* amax;    // [0] = max|x|, [1] = max|y| (device): the operands were scaled by powers of two derived from them
  int op_x, op_y, same;
};

__device__ __forceinline__ double gram_scale_of(int op, float a) {
  if (op == 2) return 1.0;
  if (op == 1) return static_cast<double>(pow2_scale(a * a));
  return static_cast<double>(pow2_scale(a));
}

template <int TERMS>
__global__ void __launch_bounds__(G_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
               const __grid_constant__ CUtensorMap map_yh, const __grid_constant__ CUtensorMap map_yl, const GramArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - raw);
  constexpr int PARTS = TERMS == 3 ? 4 : 2;                    // tiles per slot: X hi, Y hi (, X lo, Y lo)
  constexpr uint32_t SLOT_BYTES = PARTS * G_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + G_SLOTS * SLOT_BYTES);
  const uint32_t bar0 = base + G_SLOTS * SLOT_BYTES;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int FULL = 0, EMPTY = G_SLOTS, ACC_FULL = 2 * G_SLOTS, ACC_EMPTY = ACC_FULL + 1;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + ACC_EMPTY + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * a.chunks_per_cta;
  const int c1 = min(a.n_chunks, c0 + a.chunks_per_cta);
  const int pt = blockIdx.y, qt = blockIdx.z;

  if (threadIdx.x == 0) {
    for (int i = 0; i < G_SLOTS; ++i) { mbar_init(BAR(FULL + i), 1); mbar_init(BAR(EMPTY + i), 1); }
    mbar_init(BAR(ACC_FULL), 1);
    mbar_init(BAR(ACC_EMPTY), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_drains = (c1 - c0 + G_DRAIN - 1) / G_DRAIN;

  if (warp == 0) {
    if (lane == 0) {
      for (int c = c0, i = 0; c < c1; ++c, ++i) {
        const uint32_t slot = i % G_SLOTS, ph = (i / G_SLOTS) & 1;
        mbar_wait_relaxed(BAR(EMPTY + slot), ph ^ 1, 32);
        mbar_expect_tx(BAR(FULL + slot), SLOT_BYTES);
        const uint32_t dst = base + slot * SLOT_BYTES;
        tma_load_2d(dst, &map_xh, BAR(FULL + slot), c * G_KC, pt * G_T);
        tma_load_2d(dst + G_TILE_BYTES, &map_yh, BAR(FULL + slot), c * G_KC, qt * G_T);
        if (TERMS == 3) {
          tma_load_2d(dst + 2 * G_TILE_BYTES, &map_xl, BAR(FULL + slot), c * G_KC, pt * G_T);
          tma_load_2d(dst + 3 * G_TILE_BYTES, &map_yl, BAR(FULL + slot), c * G_KC, qt * G_T);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int i = 0;
      for (int d = 0; d < n_drains; ++d) {
        mbar_wait(BAR(ACC_EMPTY), (d & 1) ^ 1);              // the epilogue has read the previous block of the accumulator
        tc_fence_after();
        const int e = min(c1, c0 + (d + 1) * G_DRAIN);
        bool first = true;
        for (int c = c0 + d * G_DRAIN; c < e; ++c, ++i) {
          const uint32_t slot = i % G_SLOTS, ph = (i / G_SLOTS) & 1;
          mbar_wait(BAR(FULL + slot), ph);
          tc_fence_after();
          const uint32_t s0 = base + slot * SLOT_BYTES;
          const uint64_t xh = make_desc(s0), yh = make_desc(s0 + G_TILE_BYTES);
          const uint64_t xl = make_desc(s0 + 2 * G_TILE_BYTES), yl = make_desc(s0 + 3 * G_TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < G_KC / 16; ++kk) {
            const uint64_t off = static_cast<uint64_t>((kk * 32) >> 4);
            mma_f16_ss(tmem_base, xh + off, yh + off, G_IDESC, first ? 0u : 1u);
            first = false;
            if (TERMS == 3) {
              mma_f16_ss(tmem_base, xl + off, yh + off, G_IDESC, 1u);
              mma_f16_ss(tmem_base, xh + off, yl + off, G_IDESC, 1u);
            }
          }
          tc_commit(BAR(EMPTY + slot));
        }
        tc_commit(BAR(ACC_FULL));
      }
    }
  } else if (warp >= 4) {
    const int q4 = warp & 3;
    const int row = pt * G_T + q4 * 32 + lane;                 // output row (TMEM lane)
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    const double unscale = 1.0 / (gram_scale_of(a.op_x, a.amax[0]) * gram_scale_of(a.op_y, a.amax[a.same ? 0 : 1]));
    for (int d = 0; d < n_drains; ++d) {
      mbar_wait(BAR(ACC_FULL), d & 1);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < G_T / 32; ++g) {
        float v[32];
        tc_ld32(t_row + g * 32, v);
        tc_wait_ld_dep(v);
        if (row < a.p) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = qt * G_T + g * 32 + j;
            if (col < a.q && v[j] != 0.f) atomicAdd(a.g + static_cast<int64_t>(row) * a.ldg + col, static_cast<double>(v[j]) * unscale);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(ACC_EMPTY));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map_g(rl_context* h, CUtensorMap* map, const __half* ptr, int64_t rows, int64_t n_pad) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeTiledFn>(p);
  }
  if (!enc) return fail(h, RL_ERR_CUDA, "gram_tc: cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(n_pad), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(n_pad) * 2};
  const cuuint32_t box[2] = {G_KC, G_T};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, RL_ERR_CUDA, "gram_tc: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return RL_OK;
}

}  // namespace

// G (p x q float64, row stride ldg) = op_x(X)^T op_y(Y): X (n x p, stride ldx), Y (n x q, stride ldy) float32 on the device.
// op: 0 = x, 1 = x^2 (entries > 0), 2 = [x > 0], 3 = x [x > 0].  terms: 3 = split operands (fp32-accurate), 1 = one fp16
// product (exact for small integers).  x == y with the same op: the transposed copy is made once.
int launch_gram_tc(rl_context* h, const float* x, int ldx, int op_x, const float* y, int ldy, int op_y, int64_t n, int p, int q,
                   double* g, int ldg, int terms) {
  if (!x || !y || !g || n < 0 || p <= 0 || q <= 0 || ldg < q) return fail(h, RL_ERR_INVALID, "gram_tc: bad argument");
  if (terms != 1 && terms != 3) return fail(h, RL_ERR_INVALID, "gram_tc: terms must be 1 or 3");
  RL_CUDA(h, cudaMemset2DAsync(g, sizeof(double) * ldg, 0, sizeof(double) * q, p, h->stream));
  if (n == 0) return RL_OK;
  const int64_t n_pad = ((n + G_KC - 1) / G_KC) * G_KC;
  const int p_pad = ((p + G_T - 1) / G_T) * G_T, q_pad = ((q + G_T - 1) / G_T) * G_T;
  const bool same = x == y && ldx == ldy && op_x == op_y && p == q;
  const size_t xb = sizeof(__half) * static_cast<size_t>(p_pad) * n_pad, yb = sizeof(__half) * static_cast<size_t>(q_pad) * n_pad;
  const int halves = terms == 3 ? 2 : 1;
  void* w = nullptr;
  RL_CUDA(h, cudaMallocAsync(&w, 256 + halves * (xb + (same ? 0 : yb)), h->stream));
  char* wb = static_cast<char*>(w);
  float* amax = reinterpret_cast<float*>(wb);       // [0] = max|x|, [1] = max|y|
  __half* xh = reinterpret_cast<__half*>(wb + 256);
  __half* xl = terms == 3 ? xh + static_cast<size_t>(p_pad) * n_pad : nullptr;
  __half* yh = same ? xh : reinterpret_cast<__half*>(wb + 256 + halves * xb);
  __half* yl = same ? xl : (terms == 3 ? yh + static_cast<size_t>(q_pad) * n_pad : nullptr);
  RL_CUDA(h, cudaMemsetAsync(w, 0, 256 + halves * (xb + (same ? 0 : yb)), h->stream));
  auto prep = [&](const float* m, int ld, int op, int cols, int cols_pad, float* am, __half* hi, __half* lo) -> int {
    if ((op & 7) != 2) {
      int64_t gsz = (n * cols + 255) / 256;
      if (gsz > 8 * h->sm_count) gsz = 8 * h->sm_count;
      if (op & 8) absmax2_kernel<<<static_cast<unsigned>(gsz), 256, 0, h->stream>>>(m, cols, n, ld, am);
      else absmax2_kernel<<<static_cast<unsigned>(gsz), 256, 0, h->stream>>>(m, n, cols, ld, am);
      RL_LAUNCH_CHECK(h, "absmax2_kernel");
    }
    const dim3 grid(static_cast<unsigned>((n_pad + 31) / 32), static_cast<unsigned>((cols + 31) / 32));
    transpose_split_kernel<<<grid, 256, 0, h->stream>>>(m, n, cols, ld, op, am, n_pad, hi, lo);
    RL_LAUNCH_CHECK(h, "transpose_split_kernel");
    (void)cols_pad;
    return RL_OK;
  };
  int rc = prep(x, ldx, op_x, p, p_pad, amax, xh, xl);
  if (rc == RL_OK && !same) rc = prep(y, ldy, op_y, q, q_pad, amax + 1, yh, yl);
  if (rc != RL_OK) { cudaFreeAsync(w, h->stream); return rc; }
  CUtensorMap mxh, mxl, myh, myl;
  if ((rc = make_map_g(h, &mxh, xh, p_pad, n_pad)) != RL_OK || (rc = make_map_g(h, &myh, yh, q_pad, n_pad)) != RL_OK ||
      (rc = make_map_g(h, &mxl, terms == 3 ? xl : xh, p_pad, n_pad)) != RL_OK ||
      (rc = make_map_g(h, &myl, terms == 3 ? yl : yh, q_pad, n_pad)) != RL_OK) {
    cudaFreeAsync(w, h->stream);
    return rc;
  }
  GramArgs a{};
  a.g = g; a.ldg = ldg; a.p = p; a.q = q;
  a.n_chunks = static_cast<int>(n_pad / G_KC);
  const int tiles = (p_pad / G_T) * (q_pad / G_T);
  int split = (2 * h->sm_count + tiles - 1) / tiles;               // CTAs along n
  if (split > a.n_chunks) split = a.n_chunks;
  if (split < 1) split = 1;
  a.chunks_per_cta = (a.n_chunks + split - 1) / split;
  split = (a.n_chunks + a.chunks_per_cta - 1) / a.chunks_per_cta;
  a.amax = amax; a.op_x = op_x & 7; a.op_y = op_y & 7; a.same = same ? 1 : 0;
  const dim3 grid(static_cast<unsigned>(split), static_cast<unsigned>(p_pad / G_T), static_cast<unsigned>(q_pad / G_T));
  const size_t smem = G_SLOTS * (terms == 3 ? 4 : 2) * G_TILE_BYTES + 256 + 1024;
  if (terms == 3) {
    RL_CUDA(h, cudaFuncSetAttribute(gram_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    gram_tc_kernel<3><<<grid, G_THREADS, smem, h->stream>>>(mxh, mxl, myh, myl, a);
  } else {
    RL_CUDA(h, cudaFuncSetAttribute(gram_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    gram_tc_kernel<1><<<grid, G_THREADS, smem, h->stream>>>(mxh, mxl, myh, myl, a);
  }
  RL_LAUNCH_CHECK(h, "gram_tc_kernel");
  RL_CUDA(h, cudaFreeAsync(w, h->stream));
  return RL_OK;
}

namespace {

// sim[i][j] = num / sqrt(D[i][j] D[j][i]) for i != j with at least one co-rater, else 0 (reference algo.py:343-349)
__global__ void __launch_bounds__(256) knn_sim_finish_kernel(const double* __restrict__ num, const double* __restrict__ dd,
                                                             const double* __restrict__ cnt, int n, double* __restrict__ sim) {
  const int64_t total = static_cast<int64_t>(n) * n;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(p / n), j = static_cast<int>(p - static_cast<int64_t>(i) * n);
    double v = 0.0;
    if (i != j && cnt[p] > 0.5) {
      const double den = sqrt(dd[p] * dd[static_cast<int64_t>(j) * n + i]);
      v = den > 0.0 ? num[p] / den : 0.0;
    }
    sim[p] = v;
  }
}

// Wt[j][i] = sim[i][j] for the `k` items j that np.argsort(sim[i])[-k:] keeps (largest similarities), 0 elsewhere
__global__ void __launch_bounds__(256) knn_weights_kernel(const double* __restrict__ sim, const int32_t* __restrict__ nbr, int n, int k,
                                                          float* __restrict__ wt) {
  const int64_t total = static_cast<int64_t>(n) * k;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(p / k);
    const int j = nbr[p];
    if (j >= 0) wt[static_cast<int64_t>(j) * n + i] = static_cast<float>(sim[static_cast<int64_t>(i) * n + j]);
  }
}

// pred[u][i] = num / den where the user has not rated item i and the rated neighbours carry positive weight (algo.py:151-167)
__global__ void __launch_bounds__(256) knn_predict_finish_kernel(const float* __restrict__ r, const double* __restrict__ num,
                                                                 const double* __restrict__ den, int64_t total, double* __restrict__ pred) {
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += static_cast<int64_t>(gridDim.x) * blockDim.x)
    pred[p] = (r[p] == 0.f && den[p] > 0.0) ? num[p] / den[p] : 0.0;
}

}  // namespace

// item-item cosine over co-raters as three masked Gram products: ratings (m x n) float32 on the device
int launch_knn_item_sim(rl_context* h, const float* r, int64_t m, int n, double* sim) {
  if (!r || !sim || m <= 0 || n <= 0) return fail(h, RL_ERR_INVALID, "knn_item_sim: bad argument");
  void* w = nullptr;
  const size_t nn = static_cast<size_t>(n) * n;
  RL_CUDA(h, cudaMallocAsync(&w, 3 * nn * sizeof(double), h->stream));
  double* num = static_cast<double*>(w);
  double* dd = num + nn;
  double* cnt = dd + nn;
  int rc = launch_gram_tc(h, r, n, 3, r, n, 3, m, n, n, num, n, 3);           // sum r_ui r_uj over the raters of both
  if (rc == RL_OK) rc = launch_gram_tc(h, r, n, 1, r, n, 2, m, n, n, dd, n, 3);   // sum r_ui^2 over the raters of j
  if (rc == RL_OK) rc = launch_gram_tc(h, r, n, 2, r, n, 2, m, n, n, cnt, n, 1);  // number of co-raters (exact)
  if (rc == RL_OK) {
    int64_t g = (static_cast<int64_t>(nn) + 255) / 256;
    if (g > 16 * h->sm_count) g = 16 * h->sm_count;
    knn_sim_finish_kernel<<<static_cast<unsigned>(g), 256, 0, h->stream>>>(num, dd, cnt, n, sim);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(h, e, "knn_sim_finish_kernel"); else h->launches++;
  }
  cudaFreeAsync(w, h->stream);
  return rc;
}

// reference KNN predict (algo.py:148-169) as two products with the k-sparse neighbour weights: num = R+ W^T, den = [R > 0] W^T
int launch_knn_predict(rl_context* h, const float* r, int64_t m, int n, const double* sim, int k, double* pred) {
  if (!r || !sim || !pred || m <= 0 || n <= 0 || k <= 0) return fail(h, RL_ERR_INVALID, "knn_predict: bad argument");
  if (k > n) k = n;
  if (k > 128) return fail(h, RL_ERR_INVALID, "knn_predict: neighbors=%d exceeds the selection capacity of 128", k);
  void* w = nullptr;
  const size_t nn = static_cast<size_t>(n) * n, mn = static_cast<size_t>(m) * n;
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t o_wt = al(sizeof(int32_t) * static_cast<size_t>(n) * k), o_num = o_wt + al(sizeof(float) * nn), o_den = o_num + al(sizeof(double) * mn);
  RL_CUDA(h, cudaMallocAsync(&w, o_den + al(sizeof(double) * mn), h->stream));
  char* wb = static_cast<char*>(w);
  int32_t* nbr = reinterpret_cast<int32_t*>(wb);
  float* wt = reinterpret_cast<float*>(wb + o_wt);
  double* num = reinterpret_cast<double*>(wb + o_num);
  double* den = reinterpret_cast<double*>(wb + o_den);
  RL_CUDA(h, cudaMemsetAsync(wt, 0, sizeof(float) * nn, h->stream));
  // the k largest similarities of every item row (known = nothing masked: pass the zero matrix wt as `known`)
  int rc = launch_topn_dense(h, sim, RL_F64, wt, RL_F32, n, n, k, nbr, nullptr);
  if (rc == RL_OK) {
    int64_t g = (static_cast<int64_t>(n) * k + 255) / 256;
    if (g > 16 * h->sm_count) g = 16 * h->sm_count;
    knn_weights_kernel<<<static_cast<unsigned>(g), 256, 0, h->stream>>>(sim, nbr, n, k, wt);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(h, e, "knn_weights_kernel"); else h->launches++;
  }
  // contraction over the items j: X = R^T (given as R, already "transposed": users x items), Y = W^T (items x items)
  if (rc == RL_OK) rc = launch_gram_tc(h, r, n, 3 | 8, wt, n, 0, n, static_cast<int>(m), n, num, n, 3);
  if (rc == RL_OK) rc = launch_gram_tc(h, r, n, 2 | 8, wt, n, 0, n, static_cast<int>(m), n, den, n, 3);
  if (rc == RL_OK) {
    int64_t g = (static_cast<int64_t>(mn) + 255) / 256;
    if (g > 16 * h->sm_count) g = 16 * h->sm_count;
    knn_predict_finish_kernel<<<static_cast<unsigned>(g), 256, 0, h->stream>>>(r, num, den, static_cast<int64_t>(mn), pred);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(h, e, "knn_predict_finish_kernel"); else h->launches++;
  }
  cudaFreeAsync(w, h->stream);
  return rc;
}

}  // namespace rl
