// als_k128.cuh -- ALS half-step, fp32 + refinement path for 64 < padded k <= 128 (the C5 configuration).
// Included by als.cu after als_k64.cuh (uses its 64 x 64 building blocks).
//
// Same arithmetic as als_k64.cuh (reference algo.py:314-322 / :324-332), one CTA of four warps per row, the 128 x 128
// normal equations as 2 x 2 blocks of 64:
//
//        A = [A11  .  ]      L11 = chol(A11)             (warp 0, k64_chol)
//            [A21  A22]      L21 = A21 L11^-T            (all warps: 16 rows each; update on the tensor cores, row-local solve)
//                            S   = A22 - L21 L21^T       (all warps: tensor cores)
//                            L22 = chol(S)               (warp 0, k64_chol)
//
//   * Gram: every warp reads the same gathered chunk (8 rows, shared buffer) and owns a part of the tiles: warp 0 the
//     lower triangle of A11 (20 tiles), warp 1 that of A22 (20), warps 2 and 3 the upper / lower half of A21 (16 each);
//     fp16 split operands, B fragments are A registers (see als_k64.cuh);
//   * the right-hand side rides along with both factorisations (forward substitution), the coupling term
//     b2 -= L21 z1 with the row-local solve of L21;
//   * back substitution and the float64 matrix-free refinement as in the k = 64 kernel (residual: one thread per
//     factor dimension, no atomics).
// Rows with at most 64 observations (fewer observations than factors: the user side of C5) are solved in the DUAL form,
// als_k128_wdual_kernel below.
// Unweighted (reference) mode only; weights / w0 run the general kernel.

constexpr int K128 = 128;
constexpr int K128_LDY = 132;     // row stride of a gathered row (2 LDY = 8 mod 32: conflict-free fragment reads)
#ifndef K128_CH_
#define K128_CH_ 16
#endif
constexpr int K128_CH = K128_CH_;   // gathered rows per chunk (a multiple of 8: one mma k-step per 8 rows); two CTA barriers per chunk
constexpr int K128_FSZ = 8 * 520; // the full 64 x 64 block A21 / L21: 8 panels of 64 rows x 8 floats, 8 floats of skew each

// element (i, j) of the full block: B21[k128_f0(j >> 3) + 8 i + ((j & 7) ^ k64_swz(i))]
__host__ __device__ constexpr int k128_f0(int K) { return 520 * K; }

struct K128Smem {
  static constexpr size_t p_bytes = sizeof(float) * K64_LSZ;          // P11, P22
  static constexpr size_t f_bytes = sizeof(float) * K128_FSZ;         // B21
  static constexpr size_t invd_bytes = sizeof(float) * K128;
  static constexpr size_t xs_bytes = sizeof(double) * K128;           // x in float64; aliased by the float right-hand side
  static constexpr size_t ts_bytes = sizeof(double) * K128_CH;
  static constexpr size_t y_bytes = sizeof(float) * 2 * K128_CH * K128_LDY;
  static constexpr size_t bytes = 2 * p_bytes + f_bytes + invd_bytes + xs_bytes + ts_bytes + y_bytes;
  static_assert(bytes % 16 == 0, "shared memory layout must keep 16-byte alignment");
};

// ---- rows with at most 64 observations: the DUAL form ---------------------------------------------------------------------
// x = (Y^T Y + lambda I)^-1 Y^T 1 = Y^T (Y Y^T + lambda I)^-1 1  (Y: the nnz gathered rows, nnz x k).  With fewer observations
// than factors -- every user row of the C5 configuration: ~50 observations, k = 128 -- the nnz x nnz system is an eighth of
// the Cholesky work of the 128 x 128 one and fits ONE 64 x 64 factorisation of als_k64.cuh; it is also far better
// conditioned (no eigenvalue sits at lambda), so one refinement step reaches ~1e-7.  Padded rows / columns of the 64 x 64
// system form an identity block that is neither accumulated (16-row tile groups) nor factored (nblk panels).
constexpr int K128_DUAL_MAX = 64;

// One WARP per row, as in the k = 64 kernel.  (A first version gave a row to a CTA of four warps and kept the gathered rows
// in shared memory, 35 KB: four rows in flight per SM, and three of the four warps waited at a barrier while warp 0 factored --
// 40 % of the stall samples, 58 ms per C5 shard against 32 ms for this one.)  Here G = Y Y^T is accumulated from 16-COLUMN chunks of the gathered rows (64 observations x 64 bytes, double
// buffered: 12 KB), so only the 64 x 64 system and its factor live in shared memory (23 KB per warp, nine rows in flight per
// SM), and the three matrix-vector passes of the refinement / the output (x = Y^T z, t = Y x, x = Y^T z) read the rows again
// from L2, where the gather of a moment ago has left them.
#ifndef KD_CW_
#define KD_CW_ 8
#endif
constexpr int KD_CW = KD_CW_;             // factor columns per chunk (8: 32-byte pieces, 12 rows in flight per SM; 16: 64-byte pieces, 9 rows)
constexpr int KD_LD = KD_CW == 8 ? 8 : 24;   // row stride of a chunk (floats): the float2 fragment reads of rows g, columns 2 tq are conflict-free
constexpr int KD_WARPS = 3, KD_CTAS = KD_CW == 8 ? 5 : 3;
struct K128WarpDualSmem {
  static constexpr size_t bytes = sizeof(float) * (K64_LSZ + 64 + 64) + sizeof(double) * 64 + sizeof(int) * 64 + sizeof(float) * 2 * 64 * KD_LD;
  static_assert(bytes % 16 == 0, "per-warp shared memory must keep 16-byte alignment");
};

__global__ void __launch_bounds__(KD_WARPS * 32, KD_CTAS) als_k128_wdual_kernel(const AlsArgs a) {
  constexpr int KP = K128;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  unsigned char* wb = smem_raw + static_cast<size_t>(warp) * K128WarpDualSmem::bytes;
  float* Lt = reinterpret_cast<float*>(wb);
  float* invd = Lt + K64_LSZ;
  float* bs = invd + 64;
  double* zs = reinterpret_cast<double*>(bs + 64);
  int* ids = reinterpret_cast<int*>(zs + 64);
  float* ybuf = reinterpret_cast<float*>(ids + 64);
  const float lam = static_cast<float>(a.lambda);
  const float ysc = pow2_scale(*a.other_absmax);
  const float yunsc = (1.0f / ysc) * (1.0f / ysc);
  for (;;) {
    unsigned int ticket = 0;
    if (lane == 0) ticket = atomicAdd(a.counter + 1, 1u);
    ticket = __shfl_sync(FULL_MASK, ticket, 0);
    if (static_cast<int64_t>(ticket) >= a.n_rows) break;
    const int64_t row = a.row_order ? a.row_order[ticket] : static_cast<int64_t>(ticket);
    const int64_t s = a.indptr[row];
    const int64_t nnz64 = a.indptr[row + 1] - s;
    if (nnz64 <= 0 || nnz64 > K128_DUAL_MAX) continue;     // no observation: the row keeps its value (algo.py:316-317)
    const int nnz = static_cast<int>(nnz64);
    const int nblk = (nnz + 7) >> 3, mtiles = (nnz + 15) >> 4;
    const int id0 = lane < nnz ? a.indices[s + lane] : -1;
    const int id1 = 32 + lane < nnz ? a.indices[s + 32 + lane] : -1;
    ids[lane] = id0;
    ids[32 + lane] = id1;
    bs[lane] = lane < nnz ? 1.0f : 0.0f;
    bs[32 + lane] = 32 + lane < nnz ? 1.0f : 0.0f;
    // chunk c: columns KD_CW c .. KD_CW (c + 1) - 1 of the 64 (zero-filled beyond nnz) gathered rows, in 16-byte pieces
    auto gather = [&](int c, float* yb) {
      constexpr int PPR = KD_CW / 4, RP = 32 / PPR;        // 16-byte pieces per row, rows covered by one pass of the warp
#pragma unroll
      for (int p = 0; p < 2 * PPR; ++p) {
        const int r = lane / PPR + RP * p, q4 = lane % PPR;
        const int idx = __shfl_sync(FULL_MASK, p < PPR ? id0 : id1, lane / PPR + RP * (p % PPR));
        const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(yb + r * KD_LD + 4 * q4));
        if (idx >= 0) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(a.other + static_cast<int64_t>(idx) * KP + KD_CW * c + 4 * q4));
        } else {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 0;\n" ::"r"(dst), "l"(a.other));
        }
      }
      cp_async_commit();
    };
    // ---- G = Y Y^T on the tensor cores: A(m = observation, k = factor), the B fragment of n-tile nj is an A register ----
    float acc[20][4];
#pragma unroll
    for (int t = 0; t < 20; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
    for (int c = -1; c < KP / KD_CW; ++c) {
      const int buf = c & 1;
      if (c + 1 < KP / KD_CW) {
        gather(c + 1, ybuf + (buf ^ 1) * 64 * KD_LD);
        if (c < 0) continue;
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const float* yb = ybuf + buf * 64 * KD_LD;
#pragma unroll
      for (int ks = 0; ks < KD_CW / 8; ++ks) {
        const float* yk = yb + 8 * ks + 2 * tq;
        uint32_t hi[4][2], lo[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          if (mi >= mtiles) { hi[mi][0] = hi[mi][1] = lo[mi][0] = lo[mi][1] = 0u; continue; }
          const float2 y0 = *reinterpret_cast<const float2*>(yk + (16 * mi + g) * KD_LD);
          const float2 y1 = *reinterpret_cast<const float2*>(yk + (16 * mi + 8 + g) * KD_LD);
          k64_split_h2(y0.x * ysc, y0.y * ysc, hi[mi][0], lo[mi][0]);
          k64_split_h2(y1.x * ysc, y1.y * ysc, hi[mi][1], lo[mi][1]);
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
          if (mi < mtiles) {
#pragma unroll
            for (int nj = 0; nj <= 2 * mi + 1; ++nj) mma_f16(acc[k64_tile(mi, nj)], hi[mi][0], hi[mi][1], hi[nj >> 1][nj & 1]);
          }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
          if (mi < mtiles) {
#pragma unroll
            for (int nj = 0; nj <= 2 * mi + 1; ++nj) mma_f16(acc[k64_tile(mi, nj)], hi[mi][0], hi[mi][1], lo[nj >> 1][nj & 1]);
          }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
          if (mi < mtiles) {
#pragma unroll
            for (int nj = 0; nj <= 2 * mi + 1; ++nj) mma_f16(acc[k64_tile(mi, nj)], lo[mi][0], lo[mi][1], hi[nj >> 1][nj & 1]);
          }
      }
      __syncwarp();
    }
    // ---- + lambda I (unit diagonal on the padding), accumulator tiles -> panel storage ----
    {
      const int oc = (2 * tq) ^ (((g >> 2) & 1) << 2);
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int nj = 0; nj <= 2 * mi + 1; ++nj) {
          float (&d)[4] = acc[k64_tile(mi, nj)];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int gi = 16 * mi + g + ((e & 2) ? 8 : 0), gj = 8 * nj + 2 * tq + (e & 1);
            d[e] *= yunsc;
            if (nj >= 2 * mi && gi == gj) d[e] += (gi < nnz) ? lam : 1.0f;
          }
          if (nj <= 2 * mi) *reinterpret_cast<float2*>(Lt + k64_pb0(nj) + 8 * (16 * mi + g) + oc) = make_float2(d[0], d[1]);
          *reinterpret_cast<float2*>(Lt + k64_pb0(nj) + 8 * (16 * mi + 8 + g) + oc) = make_float2(d[2], d[3]);
        }
    }
    __syncwarp();
    k64_chol(Lt, invd, bs, lane, nblk);
    // x = Y^T z for the z in zs: lane owns entries lane + 32 q; the rows come from L2
    auto yt_z = [&](double (&x)[4]) {
      x[0] = x[1] = x[2] = x[3] = 0.0;
      // eight rows (32 loads) in flight per lane: the pass is a chain of L2 latencies, not of work
      for (int i0 = 0; i0 < nnz; i0 += 8) {
        float y[8][4];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int i = min(i0 + e, nnz - 1);
          const float* yr = a.other + static_cast<int64_t>(ids[i]) * KP + lane;
#pragma unroll
          for (int q = 0; q < 4; ++q) y[e][q] = __ldg(yr + 32 * q);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const double z = i0 + e < nnz ? zs[i0 + e] : 0.0;
#pragma unroll
          for (int q = 0; q < 4; ++q) x[q] = fma(static_cast<double>(y[e][q]), z, x[q]);
        }
      }
    };
    double zd0 = 0.0, zd1 = 0.0;
    double x[4];
    for (int it = 0;; ++it) {
      float v0 = bs[lane], v1 = bs[32 + lane];
      __syncwarp();
      k64_backward(Lt, invd, v0, v1, lane, nblk);
      zd0 += static_cast<double>(v0);
      zd1 += static_cast<double>(v1);
      zs[lane] = zd0;
      zs[32 + lane] = zd1;
      __syncwarp();
      yt_z(x);
      if (it >= a.ir_steps) break;
      // r_i = 1 - <y_i, x> - lambda z_i for the observed rows, 0 on the padding
      bs[lane] = 0.0f;
      bs[32 + lane] = 0.0f;
      __syncwarp();
      for (int i0 = 0; i0 < nnz; i0 += 8) {          // eight rows per batch: 32 loads in flight, one transposed reduction
        float y[8][4];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int i = min(i0 + e, nnz - 1);
          const float* yr = a.other + static_cast<int64_t>(ids[i]) * KP + lane;
#pragma unroll
          for (int q = 0; q < 4; ++q) y[e][q] = __ldg(yr + 32 * q);
        }
        double t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          t[e] = 0.0;
#pragma unroll
          for (int q = 0; q < 4; ++q) t[e] = fma(static_cast<double>(y[e][q]), x[q], t[e]);
        }
        // lanes with bit 4 set keep rows 4 .. 7, then bit 3 picks the pair, bit 2 the row: row e = 4 b4 + 2 b3 + b2
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool up = lane & 16;
          const double send = up ? t[e] : t[e + 4], keep = up ? t[e + 4] : t[e];
          t[e] = keep + __shfl_xor_sync(FULL_MASK, send, 16);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool up = lane & 8;
          const double send = up ? t[e] : t[e + 2], keep = up ? t[e + 2] : t[e];
          t[e] = keep + __shfl_xor_sync(FULL_MASK, send, 8);
        }
        {
          const bool up = lane & 4;
          const double send = up ? t[0] : t[1], keep = up ? t[1] : t[0];
          t[0] = keep + __shfl_xor_sync(FULL_MASK, send, 4);
        }
        t[0] += __shfl_xor_sync(FULL_MASK, t[0], 2);
        t[0] += __shfl_xor_sync(FULL_MASK, t[0], 1);
        const int i = i0 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        if ((lane & 3) == 0 && i < nnz) bs[i] = static_cast<float>(1.0 - t[0] - a.lambda * zs[i]);
      }
      __syncwarp();
      k64_forward(Lt, invd, bs, lane, nblk);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float xo = (lane + 32 * q < a.k) ? static_cast<float>(x[q]) : 0.0f;
      a.mine[row * KP + lane + 32 * q] = xo;
      for (int pr = 0; pr < a.n_peers; ++pr) a.peers[pr][row * KP + lane + 32 * q] = xo;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128, 4) als_k128_kernel(const AlsArgs a) {
  constexpr int KP = K128, CH = K128_CH, LDY = K128_LDY;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* P11 = reinterpret_cast<float*>(smem_raw);
  float* P22 = P11 + K64_LSZ;
  float* B21 = P22 + K64_LSZ;
  float* invd = B21 + K128_FSZ;
  double* xs = reinterpret_cast<double*>(invd + K128);
  float* bs = reinterpret_cast<float*>(xs);
  double* ts = xs + K128;
  float* ybuf = reinterpret_cast<float*>(ts + K128_CH);
  __shared__ unsigned int s_ticket;

  const float lam = static_cast<float>(a.lambda);
  const float ysc = pow2_scale(*a.other_absmax);
  const float yunsc = (1.0f / ysc) * (1.0f / ysc);
  const int g = lane >> 2, tq = lane & 3;
  const int sg = ((g >> 2) & 1) << 2;          // swizzle bit of rows 8 m + g
  const int o0 = tq ^ sg, o1 = o0 ^ 4;         // physical positions of panel columns tq and tq + 4 in such a row
  const int oc = (2 * tq) ^ sg;                // ... of the accumulator columns 2 tq, 2 tq + 1

  for (;;) {
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(a.counter, 1u);
    __syncthreads();
    const unsigned int ticket = s_ticket;
    if (static_cast<int64_t>(ticket) >= a.n_rows) break;
    const int64_t row = a.row_order ? a.row_order[ticket] : static_cast<int64_t>(ticket);
    const int64_t s = a.indptr[row];
    const int64_t nnz64 = a.indptr[row + 1] - s;
    if (nnz64 <= 0) continue;   // reference: rows with no observation keep their previous value (algo.py:316-317)
    const int nnz = static_cast<int>(nnz64);
    if (nnz <= K128_DUAL_MAX) continue;          // fewer observations than factors: als_k128_wdual_kernel has the row
    const int nchunk = (nnz + CH - 1) / CH;

    // all threads: cp.async gather of chunk c into buffer `yb` (8 rows x 32 pieces of 16 bytes; short chunks zero-filled)
    auto gather = [&](int c, float* yb) {
      const int cnt = min(CH, nnz - c * CH);
#pragma unroll
      for (int q = 0; q < CH / 4; ++q) {
        const int p = tid + 128 * q;
        const int r = p >> 5, c4 = p & 31;
        const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(yb + r * LDY + 4 * c4));
        if (r < cnt) {
          const int idx = a.indices[s + c * CH + r];
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(a.other + static_cast<int64_t>(idx) * KP + 4 * c4));
        } else {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 0;\n" ::"r"(dst), "l"(a.other));
        }
      }
      cp_async_commit();
    };

    // ---- gather + Gram: warp 0 -> A11, warp 1 -> A22 (lower triangles), warps 2 / 3 -> upper / lower half of A21 ------
    float acc[20][4];
    float bsum = 0.f;                            // right-hand side entry `tid` (fp32: the refinement residual is b-free)
#pragma unroll
    for (int t = 0; t < 20; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
    const int abase = (warp == 0) ? 0 : (warp == 3 ? 6 : 4);   // first 16-column group of the A side
    for (int c = -1; c < nchunk; ++c) {
      const int buf = c & 1;
      if (c + 1 < nchunk) {
        gather(c + 1, ybuf + (buf ^ 1) * CH * LDY);
        if (c < 0) continue;
        cp_async_wait<1>();
      } else {
        if (c < 0) continue;
        cp_async_wait<0>();
      }
      __syncthreads();
      const float* yb = ybuf + buf * CH * LDY;
#pragma unroll
      for (int r = 0; r < CH; ++r) bsum += yb[r * LDY + tid];
#pragma unroll 1
      for (int ks = 0; ks < CH / 8; ++ks) {      // one mma k-step per 8 gathered rows
      const float* y0 = yb + (8 * ks + 2 * tq) * LDY + g;
      const float* y1 = y0 + LDY;
      if (warp < 2) {
        uint32_t hi[4][2], lo[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          k64_split_h2(y0[16 * (abase + mi)] * ysc, y1[16 * (abase + mi)] * ysc, hi[mi][0], lo[mi][0]);
          k64_split_h2(y0[16 * (abase + mi) + 8] * ysc, y1[16 * (abase + mi) + 8] * ysc, hi[mi][1], lo[mi][1]);
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int nj = 0; nj <= 2 * mi + 1; ++nj) mma_f16(acc[k64_tile(mi, nj)], hi[mi][0], hi[mi][1], hi[nj >> 1][nj & 1]);
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int nj = 0; nj <= 2 * mi + 1; ++nj) mma_f16(acc[k64_tile(mi, nj)], hi[mi][0], hi[mi][1], lo[nj >> 1][nj & 1]);
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int nj = 0; nj <= 2 * mi + 1; ++nj) mma_f16(acc[k64_tile(mi, nj)], lo[mi][0], lo[mi][1], hi[nj >> 1][nj & 1]);
      } else {
        uint32_t ahi[2][2], alo[2][2], bhi[4][2], blo[4][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          k64_split_h2(y0[16 * (abase + mi)] * ysc, y1[16 * (abase + mi)] * ysc, ahi[mi][0], alo[mi][0]);
          k64_split_h2(y0[16 * (abase + mi) + 8] * ysc, y1[16 * (abase + mi) + 8] * ysc, ahi[mi][1], alo[mi][1]);
        }
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
          k64_split_h2(y0[16 * mb] * ysc, y1[16 * mb] * ysc, bhi[mb][0], blo[mb][0]);
          k64_split_h2(y0[16 * mb + 8] * ysc, y1[16 * mb + 8] * ysc, bhi[mb][1], blo[mb][1]);
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int nj = 0; nj < 8; ++nj) mma_f16(acc[8 * mi + nj], ahi[mi][0], ahi[mi][1], bhi[nj >> 1][nj & 1]);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int nj = 0; nj < 8; ++nj) mma_f16(acc[8 * mi + nj], ahi[mi][0], ahi[mi][1], blo[nj >> 1][nj & 1]);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int nj = 0; nj < 8; ++nj) mma_f16(acc[8 * mi + nj], alo[mi][0], alo[mi][1], bhi[nj >> 1][nj & 1]);
      }
      }
      __syncthreads();                           // the buffer may be overwritten by the gather after next
    }

    // ---- accumulator tiles -> shared memory (unscaled, + lambda I on the diagonal blocks) ---------------------------
    if (warp < 2) {
      float* P = warp == 0 ? P11 : P22;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int nj = 0; nj <= 2 * mi + 1; ++nj) {
          float (&d)[4] = acc[k64_tile(mi, nj)];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int gi = 16 * mi + g + ((e & 2) ? 8 : 0), gj = 8 * nj + 2 * tq + (e & 1);
            d[e] *= yunsc;
            if (nj >= 2 * mi && gi == gj) d[e] += (64 * warp + gi < a.k) ? lam : 1.0f;
          }
          float* pn = P + k64_pb0(nj);
          if (nj <= 2 * mi) *reinterpret_cast<float2*>(pn + 8 * (16 * mi + g) + oc) = make_float2(d[0], d[1]);
          *reinterpret_cast<float2*>(pn + 8 * (16 * mi + 8 + g) + oc) = make_float2(d[2], d[3]);
        }
    } else {
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int nj = 0; nj < 8; ++nj) {
          const float (&d)[4] = acc[8 * mi + nj];
          float* pn = B21 + k128_f0(nj) + 8 * (16 * (2 * (warp - 2) + mi) + g) + oc;
          *reinterpret_cast<float2*>(pn) = make_float2(d[0] * yunsc, d[1] * yunsc);
          *reinterpret_cast<float2*>(pn + 64) = make_float2(d[2] * yunsc, d[3] * yunsc);
        }
    }
    bs[tid] = bsum;
    __syncthreads();

    // ---- L11 = chol(A11), z1 = L11^-1 b1 ------------------------------------------------------------------------------
    if (warp == 0) k64_chol(P11, invd, bs, lane);
    __syncthreads();

    // ---- L21 = A21 L11^-T (warp w: rows 16 w .. 16 w + 15 of the block), b2 -= L21 z1 -----------------------------------
    {
      const int rr = min(lane, 15);              // lanes 16..31 shadow row 15 (results dropped)
      const int srow = k64_swz(rr);              // 16 w is a multiple of 16: swizzle bit of row 16 w + r is that of r
      float bb = bs[64 + 16 * warp + rr];
#pragma unroll 1
      for (int J = 0; J < 8; ++J) {
        float* fj = B21 + k128_f0(J) + 8 * (16 * warp);
        if (J > 0) {
          float d[4];
          {
            const float2 t0 = *reinterpret_cast<const float2*>(fj + 8 * g + oc);
            const float2 t1 = *reinterpret_cast<const float2*>(fj + 8 * (8 + g) + oc);
            d[0] = t0.x; d[1] = t0.y; d[2] = t1.x; d[3] = t1.y;
          }
#pragma unroll 1
          for (int K = 0; K < J; ++K) {
            const float* pk = P11 + k64_pb0(K);
            const float* pa = B21 + k128_f0(K) + 8 * (16 * warp + g);
            uint32_t bh[2], bl[2], ah[4], al[4];
            k64_split(pk[8 * (8 * J + g) + o0], bh[0], bl[0]);
            k64_split(pk[8 * (8 * J + g) + o1], bh[1], bl[1]);
            k64_split_neg(pa[o0], ah[0], al[0]);
            k64_split_neg(pa[64 + o0], ah[1], al[1]);
            k64_split_neg(pa[o1], ah[2], al[2]);
            k64_split_neg(pa[64 + o1], ah[3], al[3]);
            mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
            mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
            mma_tf32(d, al[0], al[1], al[2], al[3], bh[0], bh[1]);
          }
          *reinterpret_cast<float2*>(fj + 8 * g + oc) = make_float2(d[0], d[1]);
          *reinterpret_cast<float2*>(fj + 8 * (8 + g) + oc) = make_float2(d[2], d[3]);
          __syncwarp();
        }
        // row-local solve against the 8 x 8 diagonal block J of L11 (broadcast loads)
        float av[8];
        {
          const float4 lo = *reinterpret_cast<const float4*>(fj + 8 * rr + srow);
          const float4 hi = *reinterpret_cast<const float4*>(fj + 8 * rr + (srow ^ 4));
          av[0] = lo.x; av[1] = lo.y; av[2] = lo.z; av[3] = lo.w; av[4] = hi.x; av[5] = hi.y; av[6] = hi.z; av[7] = hi.w;
        }
        const float* dj = P11 + k64_pb0(J) + 64 * J;     // row c of the diagonal block at dj + 8 c
        const float4 i0 = *reinterpret_cast<const float4*>(invd + 8 * J), i1 = *reinterpret_cast<const float4*>(invd + 8 * J + 4);
        const float4 z0 = *reinterpret_cast<const float4*>(bs + 8 * J), z1v = *reinterpret_cast<const float4*>(bs + 8 * J + 4);
        const float iv[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
        const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1v.x, z1v.y, z1v.z, z1v.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float accv = av[c];
          if (c > 0) {
            const int sc = k64_swz(c);
            const float4 dl = *reinterpret_cast<const float4*>(dj + 8 * c + sc);
            const float4 dh = *reinterpret_cast<const float4*>(dj + 8 * c + (sc ^ 4));
            const float dv[8] = {dl.x, dl.y, dl.z, dl.w, dh.x, dh.y, dh.z, dh.w};
#pragma unroll
            for (int cp = 0; cp < c; ++cp) accv = fmaf(-av[cp], dv[cp], accv);
          }
          av[c] = accv * iv[c];
          bb = fmaf(-av[c], zv[c], bb);
        }
        if (lane < 16) {
          *reinterpret_cast<float4*>(fj + 8 * rr + srow) = make_float4(av[0], av[1], av[2], av[3]);
          *reinterpret_cast<float4*>(fj + 8 * rr + (srow ^ 4)) = make_float4(av[4], av[5], av[6], av[7]);
        }
        __syncwarp();
      }
      if (lane < 16) bs[64 + 16 * warp + rr] = bb;
    }
    __syncthreads();

    // ---- S = A22 - L21 L21^T on the tensor cores: warp 0: tiles (3, 0..3), 1: (3, 4..7), 2: (2, 0..5), 3: (0, 0..1), (1, 0..3)
    {
      const int ntile = warp < 2 ? 4 : 6;
#pragma unroll 1
      for (int t = 0; t < ntile; ++t) {
        int mi, nj;
        if (warp == 0) { mi = 3; nj = t; }
        else if (warp == 1) { mi = 3; nj = 4 + t; }
        else if (warp == 2) { mi = 2; nj = t; }
        else { mi = t < 2 ? 0 : 1; nj = t < 2 ? t : t - 2; }
        float* pn = P22 + k64_pb0(nj);
        const bool upper_ok = nj <= 2 * mi;      // rows 16 mi + g of the odd tile nj = 2 mi + 1 lie above the diagonal
        float d[4];
        {
          const float2 t0 = *reinterpret_cast<const float2*>(pn + 8 * (16 * mi + g) + oc);
          const float2 t1 = *reinterpret_cast<const float2*>(pn + 8 * (16 * mi + 8 + g) + oc);
          d[0] = t0.x; d[1] = t0.y; d[2] = t1.x; d[3] = t1.y;
        }
#pragma unroll 2
        for (int K = 0; K < 8; ++K) {
          const float* pa = B21 + k128_f0(K) + 8 * (16 * mi + g);
          const float* pb = B21 + k128_f0(K) + 8 * (8 * nj + g);
          uint32_t bh[2], bl[2], ah[4], al[4];
          k64_split(pb[o0], bh[0], bl[0]);
          k64_split(pb[o1], bh[1], bl[1]);
          k64_split_neg(pa[o0], ah[0], al[0]);
          k64_split_neg(pa[64 + o0], ah[1], al[1]);
          k64_split_neg(pa[o1], ah[2], al[2]);
          k64_split_neg(pa[64 + o1], ah[3], al[3]);
          mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
          mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
          mma_tf32(d, al[0], al[1], al[2], al[3], bh[0], bh[1]);
        }
        if (upper_ok) *reinterpret_cast<float2*>(pn + 8 * (16 * mi + g) + oc) = make_float2(d[0], d[1]);
        *reinterpret_cast<float2*>(pn + 8 * (16 * mi + 8 + g) + oc) = make_float2(d[2], d[3]);
      }
    }
    __syncthreads();

    // ---- L22 = chol(S), z2 = L22^-1 (b2 - L21 z1); back substitution; refinement -----------------------------------------
    if (warp == 0) k64_chol(P22, invd + 64, bs + 64, lane);
    // x = L^-T z for z in bs[0..127]  (warp 0): x2 = L22^-T z2, x1 = L11^-T (z1 - L21^T x2)
    auto backsolve = [&](float& x1a, float& x1b, float& x2a, float& x2b) {
      x2a = bs[64 + lane]; x2b = bs[96 + lane];
      x1a = bs[lane]; x1b = bs[32 + lane];
      __syncwarp();
      k64_backward(P22, invd + 64, x2a, x2b, lane);
      const float* c0 = B21 + k128_f0(lane >> 3);          // L21[i][l]      = c0[8 i + ((l & 7) ^ swz(i))]
      const float* c1 = B21 + k128_f0(4 + (lane >> 3));    // L21[i][l + 32]
      const int q0 = lane & 7, q1 = q0 ^ 4;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        const float xi = __shfl_sync(FULL_MASK, x2a, i);
        const int q = (i & 4) ? q1 : q0;
        x1a = fmaf(-c0[8 * i + q], xi, x1a);
        x1b = fmaf(-c1[8 * i + q], xi, x1b);
      }
#pragma unroll 8
      for (int i = 32; i < 64; ++i) {
        const float xi = __shfl_sync(FULL_MASK, x2b, i - 32);
        const int q = (i & 4) ? q1 : q0;
        x1a = fmaf(-c0[8 * i + q], xi, x1a);
        x1b = fmaf(-c1[8 * i + q], xi, x1b);
      }
      k64_backward(P11, invd, x1a, x1b, lane);
    };
    double xd[4] = {0.0, 0.0, 0.0, 0.0};          // warp 0: x entries lane, lane + 32, lane + 64, lane + 96
    // Fewer observations than factors (64 < nnz < k: the dual kernel takes nnz <= 64): k - nnz eigenvalues of the matrix are
    // exactly lambda, its condition number ~1e6 -- one refinement step leaves ~1e-4, so such rows get at least two.
    const int ir_row = (nnz < a.k && a.ir_steps == 1) ? 2 : a.ir_steps;
    for (int it = 0;; ++it) {
      if (warp == 0) {
        float x1a, x1b, x2a, x2b;
        backsolve(x1a, x1b, x2a, x2b);
        xd[0] += static_cast<double>(x1a); xd[1] += static_cast<double>(x1b);
        xd[2] += static_cast<double>(x2a); xd[3] += static_cast<double>(x2b);
      }
      if (it >= ir_row) break;
      if (warp == 0) {
        xs[lane] = xd[0]; xs[32 + lane] = xd[1]; xs[64 + lane] = xd[2]; xs[96 + lane] = xd[3];
      }
      __syncthreads();
      // r = b - A x = sum_i (1 - <y_i, x>) y_i - lambda x : thread `tid` owns entry `tid`
      double r = -((tid < a.k) ? a.lambda : 1.0) * xs[tid];
      for (int c = -1; c < nchunk; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunk) {
          gather(c + 1, ybuf + (buf ^ 1) * CH * LDY);
          if (c < 0) continue;
          cp_async_wait<1>();
        } else {
          if (c < 0) continue;
          cp_async_wait<0>();
        }
        __syncthreads();
        const float* yb = ybuf + buf * CH * LDY;
        const int cnt = min(CH, nnz - c * CH);
#pragma unroll
        for (int rr2 = 0; rr2 < CH / 4; ++rr2) {  // warp w: rows (CH / 4) w .. of the chunk
          const int i = (CH / 4) * warp + rr2;
          const float* yr = yb + i * LDY;
          double t = 0.0;
#pragma unroll
          for (int q = 0; q < 4; ++q) t = fma(static_cast<double>(yr[lane + 32 * q]), xs[lane + 32 * q], t);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL_MASK, t, o);
          if (lane == 0) ts[i] = (i < cnt) ? 1.0 - t : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < CH; ++i) r = fma(ts[i], static_cast<double>(yb[i * LDY + tid]), r);
        __syncthreads();
      }
      bs[tid] = static_cast<float>(r);            // xs is dead from here on: its bytes take the right-hand side
      __syncthreads();
      if (warp == 0) {
        k64_forward(P11, invd, bs, lane);         // z1
        {                                          // r2 -= L21 z1 : lane owns rows lane, lane + 32 of L21
          float b2a = bs[64 + lane], b2b = bs[96 + lane];
          const int s0 = k64_swz(lane);            // swz(lane + 32) is the same
#pragma unroll 1
          for (int K = 0; K < 8; ++K) {
            const float4 z0 = *reinterpret_cast<const float4*>(bs + 8 * K), z1v = *reinterpret_cast<const float4*>(bs + 8 * K + 4);
            const float* fa = B21 + k128_f0(K) + 8 * lane;
            const float4 al0 = *reinterpret_cast<const float4*>(fa + s0), ah0 = *reinterpret_cast<const float4*>(fa + (s0 ^ 4));
            const float4 al1 = *reinterpret_cast<const float4*>(fa + 256 + s0), ah1 = *reinterpret_cast<const float4*>(fa + 256 + (s0 ^ 4));
            b2a -= al0.x * z0.x + al0.y * z0.y + al0.z * z0.z + al0.w * z0.w + ah0.x * z1v.x + ah0.y * z1v.y + ah0.z * z1v.z + ah0.w * z1v.w;
            b2b -= al1.x * z0.x + al1.y * z0.y + al1.z * z0.z + al1.w * z0.w + ah1.x * z1v.x + ah1.y * z1v.y + ah1.z * z1v.z + ah1.w * z1v.w;
          }
          __syncwarp();
          bs[64 + lane] = b2a;
          bs[96 + lane] = b2b;
          __syncwarp();
        }
        k64_forward(P22, invd + 64, bs + 64, lane);   // z2
      }
    }
    if (warp == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float xo = (lane + 32 * q < a.k) ? static_cast<float>(xd[q]) : 0.0f;
        a.mine[row * KP + lane + 32 * q] = xo;
        for (int pr = 0; pr < a.n_peers; ++pr) a.peers[pr][row * KP + lane + 32 * q] = xo;
      }
    }
  }
}

int launch_k128(rl_context* h, const AlsArgs& args) {
  RL_CUDA(h, cudaMemsetAsync(args.counter, 0, 2 * sizeof(unsigned int), h->stream));
  {   // rows with at most 64 observations: the dual form, one warp per row
    auto kern = als_k128_wdual_kernel;
    const size_t smem = K128WarpDualSmem::bytes * KD_WARPS;
    RL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    RL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, KD_WARPS * 32, smem));
    if (per_sm < 1) return fail(h, RL_ERR_CUDA, "als_k128 (dual): kernel does not fit on an SM (smem %zu)", smem);
    int64_t grid = static_cast<int64_t>(h->sm_count) * per_sm;
    const int64_t want = (args.n_rows + KD_WARPS - 1) / KD_WARPS;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    kern<<<static_cast<unsigned>(grid), KD_WARPS * 32, smem, h->stream>>>(args);
    RL_LAUNCH_CHECK(h, "als_k128_wdual_kernel");
  }
  auto kern = als_k128_kernel;
  const size_t smem = K128Smem::bytes;
  RL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  RL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
  if (per_sm < 1) return fail(h, RL_ERR_CUDA, "als_k128: kernel does not fit on an SM (smem %zu)", smem);
  int64_t grid = static_cast<int64_t>(h->sm_count) * per_sm;
  if (grid > args.n_rows) grid = args.n_rows;
  if (grid < 1) grid = 1;
  kern<<<static_cast<unsigned>(grid), 128, smem, h->stream>>>(args);
  RL_LAUNCH_CHECK(h, "als_k128_kernel");
  return RL_OK;
}
