// als.cu -- ALS half-step: CSR gather -> per-row Gram / RHS -> Cholesky -> solve (+ iterative refinement).
//
// Replaces the per-row loops of RecommenderSystem._fit_als (reference src/recsys_lite/algo.py:314-322 and
// :324-332).  One warp owns one row r of the side being solved:
//
//   1. gather   the rows y_i (i in S_r, ascending) of the fixed factor matrix are streamed into shared
//               memory in chunks of CH rows with cp.async (16-byte pieces, double buffered);
//   2. Gram     A_r = sum w_i y_i y_i^T is accumulated in registers: the lower triangle of the KP x KP
//               matrix is cut into T x T register tiles (8 x 8 tile grid = 28 off-diagonal tiles on lanes
//               0..27, 4 diagonal tiles on lanes 28..31, the other 4 diagonal tiles as 1 x T strips spread over
//               all lanes), so every lane issues the same FMA stream: 72 FMA per gathered row at KP = 64;
//   3. solve    packed lower-triangular Cholesky in shared memory (left-looking, lanes over rows), forward /
//               backward substitution with the right-hand side in registers;
//   4. refine   (fp32 mode) `ir_steps` rounds of x += A~^{-1}(b - A x) with the residual evaluated in float64
//               directly from the re-gathered rows (matrix free), which removes the cond(A) ~ 1e5 error of the
//               fp32 factorisation;
//   5. store    x rounded to float32 into the row of the factor matrix being solved.
//
// Rows are handed out through a global atomic counter (persistent warps), optionally in a caller supplied
// order (longest first) so that skewed degree distributions do not leave a tail.
//
// Algorithmic bytes per row: nnz_r * (kp*4 + 4) gathered + 16 indptr + kp*4 written.

#include <cstdlib>

#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace rl {

namespace {

constexpr size_t ALS_SMEM_MAX = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100

// offset of row i in the packed lower-triangular storage; every row is padded to a multiple of 4 entries so
// that rows start 16-byte aligned (float) / 32-byte aligned (double)
__host__ __device__ constexpr int tri_off(int i) {
  return 8 * (i >> 2) * ((i >> 2) + 1) + (i & 3) * (4 * (i >> 2) + 4);
}

template <int KP>
struct AlsCfg {
  static constexpr int T = KP >= 64 ? 8 : KP / 8;    // register tile edge (2, 4, 8, 8)
  static constexpr int NB = KP / T;                   // tiles per side (8, 8, 8, 16)
  static constexpr int NOFF = NB * (NB - 1) / 2;      // strictly-lower tiles
  static constexpr int NPASS = NB == 8 ? 1 : 4;       // passes of 32 main tiles
  static constexpr int NDIAG_MAIN = NPASS * 32 - NOFF;  // diagonal tiles handled as full tiles
  static constexpr int NDIAG_LEFT = NB - NDIAG_MAIN;    // diagonal tiles handled as strips
  static constexpr int NSTRIP = NDIAG_LEFT * T;
  static constexpr int SPASS = (NSTRIP + 31) / 32;
  static constexpr int Q = (KP + 31) / 32;            // vector entries per lane
  static constexpr int CH = 16;                       // gathered rows per chunk
  static constexpr int LDY = KP + 4;                  // shared-memory row stride of a gathered row (floats)
  static constexpr int TRI = tri_off(KP);
};

template <typename real, int KP>
struct WarpSmem {
  using C = AlsCfg<KP>;
  static constexpr size_t a_bytes = sizeof(real) * C::TRI;
  static constexpr size_t invd_bytes = sizeof(real) * KP;
  static constexpr size_t xs_bytes = sizeof(double) * KP;
  static constexpr size_t ts_bytes = sizeof(double) * 32;
  static constexpr size_t y_bytes = sizeof(float) * 2 * C::CH * C::LDY;
  static constexpr size_t w_bytes = sizeof(float2) * 2 * C::CH;
  static constexpr size_t bytes = a_bytes + invd_bytes + xs_bytes + ts_bytes + y_bytes + w_bytes;
  static_assert(bytes % 16 == 0, "per-warp shared memory must keep 16-byte alignment");
  // warps per CTA (CTAs per SM then follow from shared memory): 4 unless one CTA would not fit
  static constexpr int warps = 4 * bytes <= ALS_SMEM_MAX ? 4 : (2 * bytes <= ALS_SMEM_MAX ? 2 : 1);
  static_assert(bytes <= ALS_SMEM_MAX, "one row does not fit in shared memory");
};

struct AlsArgs {
  const int64_t* indptr;
  const int32_t* indices;
  const float* conf;
  const float* pref;
  const float* other;
  float* mine;
  int64_t n_rows;
  int k;
  double lambda;
  double w0;
  const double* gram0;
  int ir_steps;
  unsigned int* counter;
  const int32_t* row_order;
  // fused exchange (multi-GPU): the solved row is also stored into the replicas of the other ranks, which are mapped
  // into this process (CUDA IPC) and reached over NVLink -- same row offset as `mine`
  int n_peers;
  float* peers[RL_MAX_PEERS];
  const float* other_absmax;   // max |entry| of `other` (device scalar), for the fp16 operand scale of the k = 64 path
  int split_long;              // k = 64 path: rows above K64_LONG observations are left to the cooperative kernel
  const int32_t* long_list;    //   ... which takes them from this list (device), *long_count entries
  const unsigned int* long_count;
};

template <typename real, int T>
__device__ __forceinline__ void load_vec(const float* p, real (&out)[T]) {
  if constexpr (T == 8) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
    out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
  } else if constexpr (T == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
  } else {
    const float2 a = *reinterpret_cast<const float2*>(p);
    out[0] = a.x; out[1] = a.y;
  }
}

__device__ __forceinline__ void decode_tile(int t, int noff, int& bi, int& bj) {
  if (t < noff) {
    int i = 1;
    while ((i * (i + 1)) / 2 <= t) ++i;
    bi = i;
    bj = t - i * (i - 1) / 2;
  } else {
    bi = bj = t - noff;
  }
}

// cp.async gather of chunk c of the row's observation list into buffer `yb`; weights into `wb`
template <int KP, bool WEIGHTED>
__device__ __forceinline__ void gather_chunk(const AlsArgs& a, int64_t s, int nnz, int c, float* yb, float2* wb,
                                             int lane) {
  using C = AlsCfg<KP>;
  const int base = c * C::CH;
  const int cnt = min(C::CH, nnz - base);
  int my = 0;
  if (lane < cnt) {
    my = a.indices[s + base + lane];
    if constexpr (WEIGHTED) {
      const float cf = a.conf ? a.conf[s + base + lane] : 1.0f;
      const float pf = a.pref ? a.pref[s + base + lane] : 1.0f;
      wb[lane] = make_float2(cf - static_cast<float>(a.w0), cf * pf);
    }
  }
  constexpr int PPR = KP / 4;  // 16-byte pieces per gathered row
  const int total = cnt * PPR;
  for (int p0 = 0; p0 < total; p0 += 32) {
    const int p = p0 + lane;
    const int r = min(p / PPR, 31);
    const int idx = __shfl_sync(FULL_MASK, my, r);
    if (p < total) {
      const int c4 = p - r * PPR;
      cp_async16(yb + r * C::LDY + 4 * c4, a.other + static_cast<int64_t>(idx) * KP + 4 * c4);
    }
  }
  cp_async_commit();
}

// left-looking Cholesky of the packed lower triangle, in place; invd[j] = 1 / L_jj.  Returns false on breakdown.
template <typename real, int KP>
__device__ __forceinline__ bool chol_warp(real* A, real* invd, int lane) {
  constexpr int Q = AlsCfg<KP>::Q;
  bool ok = true;
  for (int j = 0; j < KP; ++j) {
    const real* Lj = A + tri_off(j);
    real s[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int i = j + lane + 32 * q;
      s[q] = 0;
      if (i < KP) {
        const real* Li = A + tri_off(i);
        real a0 = Li[j], a1 = 0, a2 = 0, a3 = 0;
        int p = 0;
        if constexpr (sizeof(real) == 4) {
          for (; p + 4 <= j; p += 4) {
            const float4 x = *reinterpret_cast<const float4*>(Li + p);
            const float4 y = *reinterpret_cast<const float4*>(Lj + p);
            a0 = fmaf(-x.x, y.x, a0); a1 = fmaf(-x.y, y.y, a1);
            a2 = fmaf(-x.z, y.z, a2); a3 = fmaf(-x.w, y.w, a3);
          }
        } else {
          for (; p + 4 <= j; p += 4) {
            const double2 x0 = *reinterpret_cast<const double2*>(Li + p);
            const double2 x1 = *reinterpret_cast<const double2*>(Li + p + 2);
            const double2 y0 = *reinterpret_cast<const double2*>(Lj + p);
            const double2 y1 = *reinterpret_cast<const double2*>(Lj + p + 2);
            a0 = fma(-x0.x, y0.x, a0); a1 = fma(-x0.y, y0.y, a1);
            a2 = fma(-x1.x, y1.x, a2); a3 = fma(-x1.y, y1.y, a3);
          }
        }
        for (; p < j; ++p) a0 -= Li[p] * Lj[p];
        s[q] = (a0 + a1) + (a2 + a3);
      }
    }
    const real piv = __shfl_sync(FULL_MASK, s[0], 0);  // lane 0, q = 0 holds row i == j
    if (!(piv > static_cast<real>(0))) ok = false;
    const real d = sqrt(piv > static_cast<real>(1e-30) ? piv : static_cast<real>(1e-30));
    const real inv = static_cast<real>(1) / d;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int i = j + lane + 32 * q;
      if (i < KP) A[tri_off(i) + j] = (i == j) ? d : s[q] * inv;
    }
    if (lane == 0) invd[j] = inv;
    __syncwarp();
  }
  return ok;
}

// v <- (L L^T)^{-1} v ; v[q] is entry lane + 32 q
template <typename real, int KP>
__device__ __forceinline__ void solve_warp(const real* A, const real* invd, real (&v)[AlsCfg<KP>::Q], int lane) {
  constexpr int Q = AlsCfg<KP>::Q;
  int off[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) off[q] = tri_off(lane + 32 * q);
  // forward: L z = v (column oriented)
#pragma unroll
  for (int jq = 0; jq < Q; ++jq) {
#pragma unroll 8
    for (int jl = 0; jl < 32; ++jl) {
      const int j = jq * 32 + jl;
      if (j < KP) {
        const real zj = __shfl_sync(FULL_MASK, v[jq], jl) * invd[j];
        if (lane == jl) v[jq] = zj;
#pragma unroll
        for (int q = jq; q < Q; ++q) {
          const int i = lane + 32 * q;
          if (i > j && i < KP) v[q] -= A[off[q] + j] * zj;
        }
      }
    }
  }
  // backward: L^T x = z (row j of L is contiguous)
#pragma unroll
  for (int jq = Q - 1; jq >= 0; --jq) {
#pragma unroll 8
    for (int jl = 31; jl >= 0; --jl) {
      const int j = jq * 32 + jl;
      if (j < KP) {
        const real xj = __shfl_sync(FULL_MASK, v[jq], jl) * invd[j];
        if (lane == jl) v[jq] = xj;
        const real* Lj = A + tri_off(j);
#pragma unroll
        for (int q = 0; q <= jq; ++q) {
          const int i = lane + 32 * q;
          if (i < j) v[q] -= Lj[i] * xj;
        }
      }
    }
  }
}

template <typename real, int KP, bool WEIGHTED>
__global__ void __launch_bounds__(WarpSmem<real, KP>::warps * 32) als_half_step_kernel(const AlsArgs a) {
  using C = AlsCfg<KP>;
  constexpr int T = C::T, Q = C::Q, CH = C::CH, LDY = C::LDY;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wbase = smem_raw + static_cast<size_t>(warp) * WarpSmem<real, KP>::bytes;
  real* A = reinterpret_cast<real*>(wbase);
  real* invd = A + C::TRI;
  double* xs = reinterpret_cast<double*>(invd + KP);
  double* ts = xs + KP;
  float* ybuf = reinterpret_cast<float*>(ts + 32);
  float2* wbuf = reinterpret_cast<float2*>(ybuf + 2 * CH * LDY);

  // tile ownership of this lane (main tiles per pass, strips per strip pass)
  int tbi[C::NPASS], tbj[C::NPASS];
#pragma unroll
  for (int p = 0; p < C::NPASS; ++p) decode_tile(p * 32 + lane, C::NOFF, tbi[p], tbj[p]);
  int sd[C::SPASS], sr[C::SPASS];
  bool sact[C::SPASS];
#pragma unroll
  for (int p = 0; p < C::SPASS; ++p) {
    const int sidx = p * 32 + lane;
    sact[p] = sidx < C::NSTRIP;
    sd[p] = C::NDIAG_MAIN + (sact[p] ? sidx / T : 0);
    sr[p] = sact[p] ? sidx % T : 0;
  }
  const real lam = static_cast<real>(a.lambda);
  const bool hkv = a.w0 != 0.0;

  for (;;) {
    unsigned int ticket = 0;
    if (lane == 0) ticket = atomicAdd(a.counter, 1u);
    ticket = __shfl_sync(FULL_MASK, ticket, 0);
    if (static_cast<int64_t>(ticket) >= a.n_rows) break;
    const int64_t row = a.row_order ? a.row_order[ticket] : static_cast<int64_t>(ticket);
    const int64_t s = a.indptr[row];
    const int64_t nnz64 = a.indptr[row + 1] - s;
    if (nnz64 <= 0) {
      if (hkv) {  // A = w0 G0 + lambda I, b = 0  ->  x = 0
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int col = lane + 32 * q;
          if (col < KP) {
            a.mine[row * KP + col] = 0.0f;
            for (int pr = 0; pr < a.n_peers; ++pr) a.peers[pr][row * KP + col] = 0.0f;
          }
        }
      }
      continue;  // reference: rows with no observation keep their previous value (algo.py:316-317)
    }
    const int nnz = static_cast<int>(nnz64);
    const int nchunk = (nnz + CH - 1) / CH;

    if constexpr (C::NPASS > 1) {  // accumulators are flushed into A chunk by chunk: start from w0*G0 + lambda*I
      for (int i = 0; i < KP; ++i) {
        for (int j = lane; j <= i; j += 32) {
          real v = hkv ? static_cast<real>(a.w0 * a.gram0[i * KP + j]) : static_cast<real>(0);
          if (i == j) v += (i < a.k) ? lam : static_cast<real>(1);
          A[tri_off(i) + j] = v;
        }
      }
      __syncwarp();
    }

    real acc[T][T];
    real sacc[C::SPASS][T];
    double bq[Q];
#pragma unroll
    for (int x = 0; x < T; ++x)
#pragma unroll
      for (int y = 0; y < T; ++y) acc[x][y] = 0;
#pragma unroll
    for (int p = 0; p < C::SPASS; ++p)
#pragma unroll
      for (int y = 0; y < T; ++y) sacc[p][y] = 0;
#pragma unroll
    for (int q = 0; q < Q; ++q) bq[q] = 0.0;

    gather_chunk<KP, WEIGHTED>(a, s, nnz, 0, ybuf, wbuf, lane);
    for (int c = 0; c < nchunk; ++c) {
      const int buf = c & 1;
      if (c + 1 < nchunk) {
        gather_chunk<KP, WEIGHTED>(a, s, nnz, c + 1, ybuf + (buf ^ 1) * CH * LDY, wbuf + (buf ^ 1) * CH, lane);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const float* yb = ybuf + buf * CH * LDY;
      const float2* wb = wbuf + buf * CH;
      const int cnt = min(CH, nnz - c * CH);

#pragma unroll
      for (int p = 0; p < C::NPASS; ++p) {
        if constexpr (C::NPASS > 1) {
#pragma unroll
          for (int x = 0; x < T; ++x)
#pragma unroll
            for (int y = 0; y < T; ++y) acc[x][y] = 0;
        }
        const int oa = T * tbi[p], ob = T * tbj[p];
        for (int r = 0; r < cnt; ++r) {
          const float* yr = yb + r * LDY;
          real xa[T], xb[T];
          load_vec<real, T>(yr + oa, xa);
          load_vec<real, T>(yr + ob, xb);
          if constexpr (WEIGHTED) {
            const real wg = wb[r].x;
#pragma unroll
            for (int x = 0; x < T; ++x) xa[x] *= wg;
          }
#pragma unroll
          for (int x = 0; x < T; ++x)
#pragma unroll
            for (int y = 0; y < T; ++y) acc[x][y] = fma(xa[x], xb[y], acc[x][y]);
        }
        if constexpr (C::NPASS > 1) {
#pragma unroll
          for (int x = 0; x < T; ++x)
#pragma unroll
            for (int y = 0; y < T; ++y) {
              const int gi = oa + x, gj = ob + y;
              if (gj <= gi) A[tri_off(gi) + gj] += acc[x][y];
            }
        }
      }
      // diagonal strips + right-hand side
      for (int r = 0; r < cnt; ++r) {
        const float* yr = yb + r * LDY;
        real wg = 1, wr = 1;
        if constexpr (WEIGHTED) { wg = wb[r].x; wr = wb[r].y; }
#pragma unroll
        for (int p = 0; p < C::SPASS; ++p) {
          real xb[T];
          load_vec<real, T>(yr + T * sd[p], xb);
          const real xa = static_cast<real>(yr[T * sd[p] + sr[p]]) * wg;
#pragma unroll
          for (int y = 0; y < T; ++y) sacc[p][y] = fma(xa, xb[y], sacc[p][y]);
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int col = lane + 32 * q;
          if (col < KP) bq[q] += static_cast<double>(wr) * static_cast<double>(yr[col]);
        }
      }
      __syncwarp();  // everyone is done with this buffer before it is refilled
    }

    // ---- flush registers into the packed lower triangle --------------------------------------------------
    if constexpr (C::NPASS == 1) {
      const int oa = T * tbi[0], ob = T * tbj[0];
#pragma unroll
      for (int x = 0; x < T; ++x)
#pragma unroll
        for (int y = 0; y < T; ++y) {
          const int gi = oa + x, gj = ob + y;
          if (gj <= gi) {
            real v = acc[x][y];
            if (hkv) v += static_cast<real>(a.w0 * a.gram0[gi * KP + gj]);
            if (gi == gj) v += (gi < a.k) ? lam : static_cast<real>(1);
            A[tri_off(gi) + gj] = v;
          }
        }
    }
#pragma unroll
    for (int p = 0; p < C::SPASS; ++p) {
      if (sact[p]) {
        const int gi = T * sd[p] + sr[p];
#pragma unroll
        for (int y = 0; y < T; ++y) {
          const int gj = T * sd[p] + y;
          if (y <= sr[p]) {
            if constexpr (C::NPASS == 1) {
              real v = sacc[p][y];
              if (hkv) v += static_cast<real>(a.w0 * a.gram0[gi * KP + gj]);
              if (gi == gj) v += (gi < a.k) ? lam : static_cast<real>(1);
              A[tri_off(gi) + gj] = v;
            } else {
              A[tri_off(gi) + gj] += sacc[p][y];
            }
          }
        }
      }
    }
    __syncwarp();

    // ---- factorise and solve ---------------------------------------------------------------------------
    chol_warp<real, KP>(A, invd, lane);
    real v[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) v[q] = static_cast<real>(bq[q]);
    solve_warp<real, KP>(A, invd, v, lane);

    double xd[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) xd[q] = static_cast<double>(v[q]);

    // ---- iterative refinement with a float64 matrix-free residual ------------------------------------------
    for (int it = 0; it < a.ir_steps; ++it) {
      double rq[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int col = lane + 32 * q;
        if (col < KP) {
          xs[col] = xd[q];
          rq[q] = bq[q] - ((col < a.k) ? a.lambda : 1.0) * xd[q];
        } else {
          rq[q] = 0.0;
        }
      }
      __syncwarp();
      if (hkv) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int col = lane + 32 * q;
          if (col < KP) {
            double t = 0.0;
            for (int cc = 0; cc < KP; ++cc) t = fma(a.gram0[col * KP + cc], xs[cc], t);
            rq[q] -= a.w0 * t;
          }
        }
      }
      gather_chunk<KP, WEIGHTED>(a, s, nnz, 0, ybuf, wbuf, lane);
      for (int c = 0; c < nchunk; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunk) {
          gather_chunk<KP, WEIGHTED>(a, s, nnz, c + 1, ybuf + (buf ^ 1) * CH * LDY, wbuf + (buf ^ 1) * CH, lane);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncwarp();
        const float* yb = ybuf + buf * CH * LDY;
        const float2* wb = wbuf + buf * CH;
        const int cnt = min(CH, nnz - c * CH);
        if (lane < cnt) {  // t_i = w_i * <y_i, x>
          const float* yr = yb + lane * LDY;
          double t0 = 0.0, t1 = 0.0;
#pragma unroll 4
          for (int cc = 0; cc < KP; cc += 4) {
            const float4 y4 = *reinterpret_cast<const float4*>(yr + cc);
            t0 = fma(static_cast<double>(y4.x), xs[cc], t0);
            t1 = fma(static_cast<double>(y4.y), xs[cc + 1], t1);
            t0 = fma(static_cast<double>(y4.z), xs[cc + 2], t0);
            t1 = fma(static_cast<double>(y4.w), xs[cc + 3], t1);
          }
          double t = t0 + t1;
          if constexpr (WEIGHTED) t *= static_cast<double>(wb[lane].x);
          ts[lane] = t;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int col = lane + 32 * q;
          if (col < KP) {
            double r = rq[q];
            for (int i = 0; i < cnt; ++i) r = fma(-ts[i], static_cast<double>(yb[i * LDY + col]), r);
            rq[q] = r;
          }
        }
        __syncwarp();
      }
#pragma unroll
      for (int q = 0; q < Q; ++q) v[q] = static_cast<real>(rq[q]);
      solve_warp<real, KP>(A, invd, v, lane);
#pragma unroll
      for (int q = 0; q < Q; ++q) xd[q] += static_cast<double>(v[q]);
    }

#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int col = lane + 32 * q;
      if (col < KP) {
        const float xo = (col < a.k) ? static_cast<float>(xd[q]) : 0.0f;
        a.mine[row * KP + col] = xo;
        for (int pr = 0; pr < a.n_peers; ++pr) a.peers[pr][row * KP + col] = xo;
      }
    }
    __syncwarp();
  }
}

// longest row of a CSR: max over r of indptr[r + 1] - indptr[r], clamped to INT_MAX
__global__ void __launch_bounds__(256) max_degree_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, int* __restrict__ out) {
  int m = 0;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < n_rows; r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t d = indptr[r + 1] - indptr[r];
    m = max(m, static_cast<int>(d > INT_MAX ? INT_MAX : d));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(FULL_MASK, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// max |x| over a float array (non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(256) absmax_kernel(const float4* __restrict__ x, int64_t n4, float* __restrict__ out) {
  float m = 0.f;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n4; p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = x[p];
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}

#include "als_k64.cuh"
#include "als_k128.cuh"

template <typename real, int KP, bool WEIGHTED>
int launch_one(rl_context* h, const AlsArgs& args) {
  auto kern = als_half_step_kernel<real, KP, WEIGHTED>;
  constexpr int ALS_WARPS = WarpSmem<real, KP>::warps;
  const size_t smem = WarpSmem<real, KP>::bytes * ALS_WARPS;
  RL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  RL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ALS_WARPS * 32, smem));
  if (per_sm < 1) return fail(h, RL_ERR_CUDA, "als_half_step: kernel does not fit on an SM (smem %zu)", smem);
  int64_t want = (args.n_rows + ALS_WARPS - 1) / ALS_WARPS;
  int64_t grid = static_cast<int64_t>(h->sm_count) * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  RL_CUDA(h, cudaMemsetAsync(args.counter, 0, sizeof(unsigned int), h->stream));
  kern<<<static_cast<unsigned>(grid), ALS_WARPS * 32, smem, h->stream>>>(args);
  RL_LAUNCH_CHECK(h, "als_half_step_kernel");
  return RL_OK;
}

template <typename real, bool WEIGHTED>
int launch_kp(rl_context* h, int kp, const AlsArgs& args) {
  switch (kp) {
    case 16: return launch_one<real, 16, WEIGHTED>(h, args);
    case 32: return launch_one<real, 32, WEIGHTED>(h, args);
    case 64: return launch_one<real, 64, WEIGHTED>(h, args);
    case 128: return launch_one<real, 128, WEIGHTED>(h, args);
  }
  return fail(h, RL_ERR_INVALID, "als_half_step: unsupported padded k %d", kp);
}

// ---- G = Y^T Y in float64 (w0 term of the general normal equations) ---------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ y, int64_t n, double* __restrict__ g) {
  // each thread owns a (KP/16) x (KP/16) block of G; the CTA streams a slab of rows through shared memory
  constexpr int B = KP / 16;
  constexpr int ROWS = 32;
  __shared__ float tile[ROWS][KP + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[B][B];
#pragma unroll
  for (int i = 0; i < B; ++i)
#pragma unroll
    for (int j = 0; j < B; ++j) acc[i][j] = 0.0;
  for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * ROWS; r0 < n; r0 += static_cast<int64_t>(gridDim.x) * ROWS) {
    const int cnt = static_cast<int>((n - r0) < ROWS ? (n - r0) : ROWS);
    for (int p = threadIdx.x; p < cnt * (KP / 4); p += 256) {
      const int r = p / (KP / 4), c4 = p % (KP / 4);
      *reinterpret_cast<float4*>(&tile[r][4 * c4]) = *reinterpret_cast<const float4*>(y + (r0 + r) * KP + 4 * c4);
    }
    __syncthreads();
    for (int r = 0; r < cnt; ++r) {
      double xa[B], xb[B];
#pragma unroll
      for (int i = 0; i < B; ++i) { xa[i] = tile[r][ty * B + i]; xb[i] = tile[r][tx * B + i]; }
#pragma unroll
      for (int i = 0; i < B; ++i)
#pragma unroll
        for (int j = 0; j < B; ++j) acc[i][j] = fma(xa[i], xb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < B; ++i)
#pragma unroll
    for (int j = 0; j < B; ++j) atomicAdd(&g[(ty * B + i) * KP + tx * B + j], acc[i][j]);
}

}  // namespace

int launch_als_half_step(rl_context* h, int64_t n_rows, int64_t n_other, int k, const int64_t* indptr,
                         const int32_t* indices, const float* conf, const float* pref, const float* other,
                         float* mine, double lambda, double w0, const double* gram0, int precision,
                         int ir_steps, const int32_t* row_order, int n_peers, float* const* peers) {
  const int kp = padded_k(k);
  if (kp == 0) return fail(h, RL_ERR_INVALID, "als_half_step: k=%d outside [1, 128]", k);
  if (n_peers < 0 || n_peers > RL_MAX_PEERS || (n_peers > 0 && !peers))
    return fail(h, RL_ERR_INVALID, "als_half_step: n_peers=%d outside [0, %d] or null peer list", n_peers, RL_MAX_PEERS);
  if (n_rows < 0 || n_other < 0) return fail(h, RL_ERR_INVALID, "als_half_step: negative size");
  if (n_rows >= (1ll << 32) - 1) return fail(h, RL_ERR_INVALID, "als_half_step: n_rows too large for the row ticket");
  if (n_rows == 0) return RL_OK;
  if (!indptr || !other || !mine) return fail(h, RL_ERR_INVALID, "als_half_step: null pointer");
  if (w0 != 0.0 && !gram0) return fail(h, RL_ERR_INVALID, "als_half_step: w0 != 0 needs gram0 (rl_gram_dev)");
  if (!(lambda >= 0.0)) return fail(h, RL_ERR_INVALID, "als_half_step: lambda must be >= 0");
  if (ir_steps < 0 || ir_steps > 8) return fail(h, RL_ERR_INVALID, "als_half_step: ir_steps outside [0, 8]");
  AlsArgs a{indptr, indices, conf, pref, other, mine, n_rows, k, lambda, w0, gram0, ir_steps, h->counter, row_order, n_peers, {}, nullptr, 0, nullptr, nullptr};
  for (int pr = 0; pr < n_peers; ++pr) {
    if (!peers[pr]) return fail(h, RL_ERR_INVALID, "als_half_step: peer pointer %d is null", pr);
    a.peers[pr] = peers[pr];
  }
  a.other_absmax = reinterpret_cast<const float*>(h->counter + 16);
  const bool weighted = conf != nullptr || pref != nullptr || w0 != 0.0;
  if (precision == RL_PREC_FP64) {
    a.ir_steps = 0;
    return weighted ? launch_kp<double, true>(h, kp, a) : launch_kp<double, false>(h, kp, a);
  }
  if (precision == RL_PREC_FP32_IR) {
    if (kp == 64 && !getenv("RL_ALS_GENERIC")) {
      RL_CUDA(h, cudaMemsetAsync(h->counter + 16, 0, sizeof(float), h->stream));
      const int64_t n4 = n_other * (kp / 4);
      int64_t g = (n4 + 255) / 256;
      if (g > 8 * h->sm_count) g = 8 * h->sm_count;
      if (g > 0) {
        absmax_kernel<<<static_cast<unsigned>(g), 256, 0, h->stream>>>(reinterpret_cast<const float4*>(other), n4,
                                                                      reinterpret_cast<float*>(h->counter + 16));
        RL_LAUNCH_CHECK(h, "absmax_kernel");
      }
      // rows above K64_LONG observations need the cooperative kernel: asked once per pattern (see rl_context)
      int has_long = -1;
      if (!getenv("RL_ALS_NO_LONG_CACHE"))      // (tests: always ask the device)
        for (const auto& e : h->long_rows_cache)
          if (e.indptr == indptr && e.n_rows == n_rows) has_long = e.has_long ? 1 : 0;
      if (has_long < 0) {
        int* slot = reinterpret_cast<int*>(h->counter + 17);
        RL_CUDA(h, cudaMemsetAsync(slot, 0, sizeof(int), h->stream));
        int64_t gd = (n_rows + 255) / 256;
        if (gd > 8 * h->sm_count) gd = 8 * h->sm_count;
        max_degree_kernel<<<static_cast<unsigned>(gd), 256, 0, h->stream>>>(indptr, n_rows, slot);
        RL_LAUNCH_CHECK(h, "max_degree_kernel");
        int md = 0;
        RL_CUDA(h, cudaMemcpyAsync(&md, slot, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        RL_CUDA(h, cudaStreamSynchronize(h->stream));
        has_long = md > K64_LONG ? 1 : 0;
        if (h->long_rows_cache.size() >= 64) h->long_rows_cache.clear();
        h->long_rows_cache.push_back({indptr, n_rows, has_long != 0});
      }
      a.split_long = has_long;
      return weighted ? launch_k64<true>(h, a) : launch_k64<false>(h, a);
    }
    if (kp == 128 && !weighted && !getenv("RL_ALS_GENERIC")) {     // 2 x 2 blocks of the k = 64 machinery, four warps per row
      RL_CUDA(h, cudaMemsetAsync(h->counter + 16, 0, sizeof(float), h->stream));
      const int64_t n4 = n_other * (kp / 4);
      int64_t g = (n4 + 255) / 256;
      if (g > 8 * h->sm_count) g = 8 * h->sm_count;
      if (g > 0) {
        absmax_kernel<<<static_cast<unsigned>(g), 256, 0, h->stream>>>(reinterpret_cast<const float4*>(other), n4,
                                                                      reinterpret_cast<float*>(h->counter + 16));
        RL_LAUNCH_CHECK(h, "absmax_kernel");
      }
      return launch_k128(h, a);
    }
    return weighted ? launch_kp<float, true>(h, kp, a) : launch_kp<float, false>(h, kp, a);
  }
  return fail(h, RL_ERR_INVALID, "als_half_step: unknown precision %d", precision);
}

int launch_gram(rl_context* h, const float* y, int64_t n, int k, double* gram) {
  const int kp = padded_k(k);
  if (kp == 0) return fail(h, RL_ERR_INVALID, "gram: k=%d outside [1, 128]", k);
  if (!y || !gram) return fail(h, RL_ERR_INVALID, "gram: null pointer");
  // tall inputs: tcgen05 split-fp16 GEMM (gram_tc.cu); short ones are latency bound either way and keep the float64 FMA kernel
  if (n >= 8192) return launch_gram_tc(h, y, kp, 0, y, kp, 0, n, kp, kp, gram, kp, 3);
  RL_CUDA(h, cudaMemsetAsync(gram, 0, sizeof(double) * kp * kp, h->stream));
  if (n == 0) return RL_OK;
  int64_t grid = (n + 31) / 32;
  if (grid > 2 * h->sm_count) grid = 2 * h->sm_count;
  switch (kp) {
    case 16: gram_kernel<16><<<grid, 256, 0, h->stream>>>(y, n, gram); break;
    case 32: gram_kernel<32><<<grid, 256, 0, h->stream>>>(y, n, gram); break;
    case 64: gram_kernel<64><<<grid, 256, 0, h->stream>>>(y, n, gram); break;
    case 128: gram_kernel<128><<<grid, 256, 0, h->stream>>>(y, n, gram); break;
  }
  RL_LAUNCH_CHECK(h, "gram_kernel");
  return RL_OK;
}

}  // namespace rl
