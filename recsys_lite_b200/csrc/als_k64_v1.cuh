// als_k64.cuh -- ALS half-step, fp32 + refinement path for padded k = 64 (the C3 / C4 configurations).
// Included by als.cu inside its anonymous namespace (uses AlsArgs and the cp.async helpers).
//
// Same algorithm as als_half_step_kernel (reference algo.py:314-322 / :324-332), different Gram and solve phases:
//
//   * Gram on the tensor cores: per chunk of 8 gathered rows the 64 x 64 update Y_c^T Y_c is 20 lower-triangle
//     m16n8k8 tiles; every element is split x = hi + lo (hi = upper 19 bits) and each tile gets hi.hi + hi.lo +
//     lo.hi (3 x tf32, fp32 accumulate), which keeps the Gram at fp32 accuracy.  The A
//     and B fragments are the SAME 16 registers (A = Y_c^T, B = Y_c), so a chunk costs 16 shared-memory loads,
//     32 split ops and 60 mma.sync instead of 8 x 72 FMAs + 8 x 36 loads.  (tcgen05 needs M >= 64 tiles fed by
//     shared-memory descriptors and a TMEM round trip per user; at 50 gathered rows per user the per-row
//     warp-level MMA is the better fit -- see DESIGN.md.)
//
//   * the Gram fragments are written once to shared memory (packed lower triangle), then every lane pulls "its" two
//     rows (l and l + 32) into registers; the same 8.7 KB then receive the factor (packed, column-major);
//   * right-looking Cholesky entirely in registers: at column j the pivot lane broadcasts a_jj, everyone scales
//     its rows (L_ij = a_ij * rsqrt(a_jj)), the column goes to shared memory (it is row j of the column-major
//     factor kept for the refinement solves) and the rank-1 update a_ik -= L_ij L_kj runs as straight FMAs on
//     register windows that are shifted by four columns after every group of four pivots, so the same code is
//     reused for all groups of a phase (phase A: columns 0..31, windows 32 / 64; phase B: columns 32..63,
//     window 32) -- no index arithmetic, no triangular packing, ~4.1 K FMA per system instead of ~16 K
//     instructions of the left-looking shared-memory version;
//   * the forward substitution of the right-hand side rides along with the factorisation.
//
// rsqrt.approx is good enough here: the factor is only a preconditioner for the float64-residual refinement.

constexpr int K64O = 64;
constexpr int K64O_TRI = 2176;  // floats of a packed 64 x 64 lower triangle with 4-aligned rows (or 4-column groups)
constexpr int K64O_LDY = 72; // row stride of a gathered row: 8 t + g is a bijection onto the banks for the mma fragments
constexpr int K64O_CH = 16;  // gathered rows per chunk (two k = 8 mma steps)

// Gram staging: row-major packed lower triangle, row i at k64o_row(i) (rows padded to a multiple of 4 entries).
__host__ __device__ constexpr int k64o_row(int i) { return 8 * (i >> 2) * ((i >> 2) + 1) + (i & 3) * (4 * (i >> 2) + 4); }
// Factor: column-major packed lower triangle.  The four columns of group g = j >> 2 all start at row 4g, so that
// column j holds rows 4g .. 63 at k64o_col(j) + (i - 4g): 16-byte aligned, float4 loads of L[4g + 4m ..][j] need no shuffling.
__host__ __device__ constexpr int k64o_col(int j) { return 256 * (j >> 2) - 8 * (j >> 2) * ((j >> 2) - 1) + (j & 3) * (64 - 4 * (j >> 2)); }
static_assert(k64o_row(63) + 64 == K64O_TRI && k64o_col(63) + 4 == K64O_TRI, "packed triangle size");

template <bool WEIGHTED>
struct K64OSmem {
  static constexpr size_t lt_bytes = sizeof(float) * K64O_TRI;
  static constexpr size_t invd_bytes = sizeof(float) * K64O;
  static constexpr size_t xs_bytes = sizeof(double) * K64O;
  static constexpr size_t ts_bytes = sizeof(double) * K64O_CH;
  static constexpr size_t y_bytes = sizeof(float) * 2 * K64O_CH * K64O_LDY;
  static constexpr size_t w_bytes = WEIGHTED ? sizeof(float2) * 2 * K64O_CH : 0;
  static constexpr size_t bytes = lt_bytes + invd_bytes + xs_bytes + ts_bytes + y_bytes + w_bytes;
  static constexpr int warps = 4;   // 3 CTAs per SM: 3 x (4 x 18.4 KB + 1 KB) < 228 KB, registers <= 168
  static_assert(bytes % 16 == 0, "per-warp shared memory must keep 16-byte alignment");
};

// cp.async gather of chunk c (16-byte pieces); rows up to the next multiple of 8 are zero-filled for the k = 8 mma
template <bool WEIGHTED>
__device__ __forceinline__ void k64o_gather(const AlsArgs& a, int64_t s, int nnz, int c, float* yb, float2* wb, int lane) {
  const int base = c * K64O_CH;
  const int cnt = min(K64O_CH, nnz - base);
  int my = 0;
  if (lane < cnt) {
    my = a.indices[s + base + lane];
    if constexpr (WEIGHTED) {
      const float cf = a.conf ? a.conf[s + base + lane] : 1.0f;
      const float pf = a.pref ? a.pref[s + base + lane] : 1.0f;
      wb[lane] = make_float2(cf - static_cast<float>(a.w0), cf * pf);
    }
  }
  constexpr int PPR = K64O / 4;  // 16-byte pieces per gathered row
  const int total = cnt * PPR, padded = ((cnt + 7) & ~7) * PPR;
  for (int p0 = 0; p0 < padded; p0 += 32) {
    const int p = p0 + lane;
    const int r = p / PPR;
    const int idx = __shfl_sync(FULL_MASK, my, r & 31);
    const int c4 = p - r * PPR;
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(yb + r * K64O_LDY + 4 * c4));
    if (p < total) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(a.other + static_cast<int64_t>(idx) * K64O + 4 * c4));
    } else if (p < padded) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 0;\n" ::"r"(dst), "l"(a.other));  // src-size 0: zero fill
    }
  }
  cp_async_commit();
}

__device__ __forceinline__ void mma_tf32_o(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3,
                                         const uint32_t b0, const uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// lower-triangle 16 x 8 tiles of the 64 x 64 Gram: tile (mi, nj) is needed iff nj <= 2 mi + 1  ->  2 + 4 + 6 + 8 = 20 tiles
__device__ __forceinline__ constexpr int k64o_tile(int mi, int nj) { return mi * (mi + 1) + nj; }

// four pivots j0 .. j0+3.  Phase A (j0 < 32): rows l (window a0, 32 wide) and l + 32 (window a1, 64 wide) are live.
// Phase B (j0 >= 32): only rows l + 32 (window a1, first 32 positions).  Window position p holds column j0 + p.
template <bool PHASE_A>
__device__ __forceinline__ void chol_group4(float (&a0)[32], float (&a1)[64], float& b0, float& b1, float* __restrict__ Lt,
                                            float* __restrict__ invd, int j0, int lane) {
  constexpr int W1 = PHASE_A ? 64 : 32;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int j = j0 + c;
    const int owner = j & 31;
    const float piv = __shfl_sync(FULL_MASK, PHASE_A ? a0[c] : a1[c], owner);
    const float inv = rsqrtf(fmaxf(piv, 1e-30f));
    const float l0 = PHASE_A ? a0[c] * inv : 0.f;
    const float l1 = a1[c] * inv;
    float* col = Lt + k64o_col(j) - j0;          // col[i] = L_ij for rows i >= j0
    if (PHASE_A) { if (lane >= j0) col[lane] = l0; }
    if (PHASE_A || lane + 32 >= j0) col[32 + lane] = l1;
    if (lane == 0) invd[j] = inv;
    // forward substitution rides along: z_j = b_j / L_jj, b_i -= L_ij z_j for i > j
    const float zj = __shfl_sync(FULL_MASK, PHASE_A ? b0 : b1, owner) * inv;
    if (PHASE_A) {
      if (lane == owner) b0 = zj;
      else if (lane > j) b0 = fmaf(-l0, zj, b0);
      b1 = fmaf(-l1, zj, b1);
    } else {
      if (lane == owner) b1 = zj;
      else if (lane + 32 > j) b1 = fmaf(-l1, zj, b1);
    }
    __syncwarp();
    // rank-1 update of the remaining window: a_i[p] -= L_ij * L_{j0+p, j}
#pragma unroll
    for (int p = c + 1; p < 4; ++p) {
      const float lk = col[j0 + p];
      if (PHASE_A) a0[p] = fmaf(-l0, lk, a0[p]);
      a1[p] = fmaf(-l1, lk, a1[p]);
    }
#pragma unroll
    for (int m = 1; m < W1 / 4; ++m) {
      // rows j0 + 4m .. of column j; beyond row 63 this reads the next column's storage: finite junk that only
      // ever reaches window slots of the (unused) upper triangle
      const float4 lk = *reinterpret_cast<const float4*>(col + j0 + 4 * m);
      const int d = (c == 3) ? 4 * m - 4 : 4 * m;  // the last pivot of the group writes the window shifted by four
      if (PHASE_A && m < 8) {
        a0[d + 0] = fmaf(-l0, lk.x, a0[4 * m + 0]);
        a0[d + 1] = fmaf(-l0, lk.y, a0[4 * m + 1]);
        a0[d + 2] = fmaf(-l0, lk.z, a0[4 * m + 2]);
        a0[d + 3] = fmaf(-l0, lk.w, a0[4 * m + 3]);
      }
      a1[d + 0] = fmaf(-l1, lk.x, a1[4 * m + 0]);
      a1[d + 1] = fmaf(-l1, lk.y, a1[4 * m + 1]);
      a1[d + 2] = fmaf(-l1, lk.z, a1[4 * m + 2]);
      a1[d + 3] = fmaf(-l1, lk.w, a1[4 * m + 3]);
    }
  }
  // columns that slid into the window belong to the (unused) upper triangle
  if (PHASE_A) { a0[28] = a0[29] = a0[30] = a0[31] = 0.f; }
  a1[W1 - 4] = a1[W1 - 3] = a1[W1 - 2] = a1[W1 - 1] = 0.f;
}

// v <- L^{-T} v   (v0: entry l, v1: entry l + 32); L_ji (j > i) sits in column i: Lt[k64o_col(i) - 4 (i >> 2) + j]
__device__ __forceinline__ void k64o_backward(const float* __restrict__ Lt, const float* __restrict__ invd, float& v0, float& v1, int lane) {
  const float* c0 = Lt + k64o_col(lane) - (lane & ~3);
  const float* c1 = Lt + k64o_col(lane + 32) - ((lane + 32) & ~3);
#pragma unroll 4
  for (int j = 63; j >= 32; --j) {
    const float xj = __shfl_sync(FULL_MASK, v1, j - 32) * invd[j];
    if (lane == j - 32) v1 = xj;
    else if (lane + 32 < j) v1 = fmaf(-c1[j], xj, v1);
    v0 = fmaf(-c0[j], xj, v0);
  }
#pragma unroll 4
  for (int j = 31; j >= 0; --j) {
    const float xj = __shfl_sync(FULL_MASK, v0, j) * invd[j];
    if (lane == j) v0 = xj;
    else if (lane < j) v0 = fmaf(-c0[j], xj, v0);
  }
}

// v <- L^{-1} v ; L_ij (i > j) = Lt[k64o_col(j) - 4 (j >> 2) + i]
__device__ __forceinline__ void k64o_forward(const float* __restrict__ Lt, const float* __restrict__ invd, float& v0, float& v1, int lane) {
#pragma unroll 4
  for (int j = 0; j < 32; ++j) {
    const float* col = Lt + k64o_col(j) - (j & ~3);
    const float zj = __shfl_sync(FULL_MASK, v0, j) * invd[j];
    if (lane == j) v0 = zj;
    else if (lane > j) v0 = fmaf(-col[lane], zj, v0);
    v1 = fmaf(-col[32 + lane], zj, v1);
  }
#pragma unroll 4
  for (int j = 32; j < 64; ++j) {
    const float* col = Lt + k64o_col(j) - (j & ~3);
    const float zj = __shfl_sync(FULL_MASK, v1, j - 32) * invd[j];
    if (lane == j - 32) v1 = zj;
    else if (lane + 32 > j) v1 = fmaf(-col[32 + lane], zj, v1);
  }
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(K64OSmem<WEIGHTED>::warps * 32, 3) als_k64o_kernel(const AlsArgs a) {
  constexpr int KP = K64O, CH = K64O_CH, LDY = K64O_LDY;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wbase = smem_raw + static_cast<size_t>(warp) * K64OSmem<WEIGHTED>::bytes;
  float* Lt = reinterpret_cast<float*>(wbase);
  float* invd = Lt + K64O_TRI;
  double* xs = reinterpret_cast<double*>(invd + K64O);
  double* ts = xs + K64O;
  float* ybuf = reinterpret_cast<float*>(ts + K64O_CH);
  float2* wbuf = reinterpret_cast<float2*>(ybuf + 2 * CH * LDY);

  const float lam = static_cast<float>(a.lambda);
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
      if (hkv) {
        a.mine[row * KP + lane] = 0.0f;
        a.mine[row * KP + 32 + lane] = 0.0f;
        for (int pr = 0; pr < a.n_peers; ++pr) {
          a.peers[pr][row * KP + lane] = 0.0f;
          a.peers[pr][row * KP + 32 + lane] = 0.0f;
        }
      }
      continue;  // reference: rows with no observation keep their previous value (algo.py:316-317)
    }
    const int nnz = static_cast<int>(nnz64);
    const int nchunk = (nnz + CH - 1) / CH;

    // ---- gather + Gram on the tensor cores (3 x tf32 split, fp32 accumulate) -----------------------------------
    float acc[20][4];
    double bq0 = 0.0, bq1 = 0.0;
#pragma unroll
    for (int t = 0; t < 20; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
    const int g = lane >> 2, tq = lane & 3;
    k64o_gather<WEIGHTED>(a, s, nnz, 0, ybuf, wbuf, lane);
    for (int c = 0; c < nchunk; ++c) {
      const int buf = c & 1;
      if (c + 1 < nchunk) {
        k64o_gather<WEIGHTED>(a, s, nnz, c + 1, ybuf + (buf ^ 1) * CH * LDY, wbuf + (buf ^ 1) * CH, lane);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const float* yb = ybuf + buf * CH * LDY;
      const float2* wb = wbuf + buf * CH;
      const int cnt = min(CH, nnz - c * CH);
      for (int r = 0; r < cnt; ++r) {   // right-hand side in float64 (an fp32 sum would be amplified by 1 / lambda)
        const float* yr = yb + r * LDY;
        double wr = 1.0;
        if constexpr (WEIGHTED) wr = static_cast<double>(wb[r].y);
        bq0 += wr * static_cast<double>(yr[lane]);
        bq1 += wr * static_cast<double>(yr[32 + lane]);
      }
      for (int k0 = 0; k0 < cnt; k0 += 8) {
        // fragments: hi[mi][0..3] = A(mi) a0..a3 = Y[k0 + tq (+4)][16 mi + g (+8)]; B(nj = 2mi) = (a0, a2), B(2mi+1) = (a1, a3)
        uint32_t hi[4][4], lo[4][4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          float v[4];
          v[0] = yb[(k0 + tq) * LDY + 16 * mi + g];
          v[1] = yb[(k0 + tq) * LDY + 16 * mi + g + 8];
          v[2] = yb[(k0 + tq + 4) * LDY + 16 * mi + g];
          v[3] = yb[(k0 + tq + 4) * LDY + 16 * mi + g + 8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hi[mi][e] = __float_as_uint(v[e]) & 0xFFFFE000u;
            lo[mi][e] = __float_as_uint(v[e] - __uint_as_float(hi[mi][e]));
          }
        }
        uint32_t ahi[4][4], alo[4][4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int e = 0; e < 4; ++e) { ahi[mi][e] = hi[mi][e]; alo[mi][e] = lo[mi][e]; }
        if constexpr (WEIGHTED) {     // scale the A side by the Gram weight of its gathered row (rows tq and tq + 4)
          const float w0 = (k0 + tq < cnt) ? wb[k0 + tq].x : 0.f;
          const float w1 = (k0 + tq + 4 < cnt) ? wb[k0 + tq + 4].x : 0.f;
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float w = (e < 2) ? w0 : w1;
              const float full = (__uint_as_float(hi[mi][e]) + __uint_as_float(lo[mi][e])) * w;
              ahi[mi][e] = __float_as_uint(full) & 0xFFFFE000u;
              alo[mi][e] = __float_as_uint(full - __uint_as_float(ahi[mi][e]));
            }
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int nj = 0; nj <= 2 * mi + 1; ++nj) {
            const int bm = nj >> 1, bo = nj & 1;   // B(nj) = (a[bo], a[bo + 2]) of m-tile bm
            float (&d)[4] = acc[k64o_tile(mi, nj)];
            mma_tf32_o(d, ahi[mi][0], ahi[mi][1], ahi[mi][2], ahi[mi][3], hi[bm][bo], hi[bm][bo + 2]);
            mma_tf32_o(d, ahi[mi][0], ahi[mi][1], ahi[mi][2], ahi[mi][3], lo[bm][bo], lo[bm][bo + 2]);
            mma_tf32_o(d, alo[mi][0], alo[mi][1], alo[mi][2], alo[mi][3], hi[bm][bo], hi[bm][bo + 2]);
          }
      }
      __syncwarp();
    }

    // ---- Gram fragments -> shared (packed row-major lower triangle) -> two rows per lane in registers ------
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int nj = 0; nj <= 2 * mi + 1; ++nj) {
        const float (&d)[4] = acc[k64o_tile(mi, nj)];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int gi = 16 * mi + g + ((e & 2) ? 8 : 0), gj = 8 * nj + 2 * tq + (e & 1);
          if (gj <= gi) {
            float v = d[e];
            if (hkv) v += static_cast<float>(a.w0 * a.gram0[gi * KP + gj]);
            if (gi == gj) v += (gi < a.k) ? lam : 1.0f;
            Lt[k64o_row(gi) + gj] = v;
          }
        }
      }
    __syncwarp();
    float a0[32], a1[64];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const float4 v = *reinterpret_cast<const float4*>(Lt + k64o_row(lane) + 4 * m);
      a0[4 * m] = v.x; a0[4 * m + 1] = v.y; a0[4 * m + 2] = v.z; a0[4 * m + 3] = v.w;
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const float4 v = *reinterpret_cast<const float4*>(Lt + k64o_row(32 + lane) + 4 * m);
      a1[4 * m] = v.x; a1[4 * m + 1] = v.y; a1[4 * m + 2] = v.z; a1[4 * m + 3] = v.w;
    }
    __syncwarp();

    // ---- Cholesky in registers, forward substitution fused ------------------------------------------------------
    float b0 = static_cast<float>(bq0), b1 = static_cast<float>(bq1);
#pragma unroll 1
    for (int j0 = 0; j0 < 32; j0 += 4) chol_group4<true>(a0, a1, b0, b1, Lt, invd, j0, lane);
#pragma unroll 1
    for (int j0 = 32; j0 < 64; j0 += 4) chol_group4<false>(a0, a1, b0, b1, Lt, invd, j0, lane);
    __syncwarp();
    k64o_backward(Lt, invd, b0, b1, lane);
    double xd0 = static_cast<double>(b0), xd1 = static_cast<double>(b1);

    // ---- iterative refinement with a float64 matrix-free residual ---------------------------------------------
    for (int it = 0; it < a.ir_steps; ++it) {
      xs[lane] = xd0;
      xs[32 + lane] = xd1;
      double r0 = bq0 - ((lane < a.k) ? a.lambda : 1.0) * xd0;
      double r1 = bq1 - ((32 + lane < a.k) ? a.lambda : 1.0) * xd1;
      __syncwarp();
      if (hkv) {
        double t0 = 0.0, t1 = 0.0;
        for (int cc = 0; cc < KP; ++cc) {
          t0 = fma(a.gram0[lane * KP + cc], xs[cc], t0);
          t1 = fma(a.gram0[(32 + lane) * KP + cc], xs[cc], t1);
        }
        r0 -= a.w0 * t0;
        r1 -= a.w0 * t1;
      }
      k64o_gather<WEIGHTED>(a, s, nnz, 0, ybuf, wbuf, lane);
      for (int c = 0; c < nchunk; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunk) {
          k64o_gather<WEIGHTED>(a, s, nnz, c + 1, ybuf + (buf ^ 1) * CH * LDY, wbuf + (buf ^ 1) * CH, lane);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncwarp();
        const float* yb = ybuf + buf * CH * LDY;
        const float2* wb = wbuf + buf * CH;
        const int cnt = min(CH, nnz - c * CH);
        // t_i = w_i <y_i, x>: two lanes per gathered row (each half of the factor dimension), then a shuffle add
        {
          const int i = lane & 15, half = lane >> 4;
          double t = 0.0;
          if (i < cnt) {
            const float* yr = yb + i * LDY + 32 * half;
            const double* xh = xs + 32 * half;
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int cc = 0; cc < 32; cc += 4) {
              const float4 y4 = *reinterpret_cast<const float4*>(yr + cc);
              t0 = fma(static_cast<double>(y4.x), xh[cc], t0);
              t1 = fma(static_cast<double>(y4.y), xh[cc + 1], t1);
              t0 = fma(static_cast<double>(y4.z), xh[cc + 2], t0);
              t1 = fma(static_cast<double>(y4.w), xh[cc + 3], t1);
            }
            t = t0 + t1;
          }
          t += __shfl_xor_sync(FULL_MASK, t, 16);
          if constexpr (WEIGHTED) { if (i < cnt) t *= static_cast<double>(wb[i].x); }
          if (lane < 16) ts[lane] = t;
        }
        __syncwarp();
        for (int i = 0; i < cnt; ++i) {
          const double ti = ts[i];
          r0 = fma(-ti, static_cast<double>(yb[i * LDY + lane]), r0);
          r1 = fma(-ti, static_cast<double>(yb[i * LDY + 32 + lane]), r1);
        }
        __syncwarp();
      }
      float v0 = static_cast<float>(r0), v1 = static_cast<float>(r1);
      k64o_forward(Lt, invd, v0, v1, lane);
      k64o_backward(Lt, invd, v0, v1, lane);
      xd0 += static_cast<double>(v0);
      xd1 += static_cast<double>(v1);
    }
    const float xo0 = (lane < a.k) ? static_cast<float>(xd0) : 0.0f;
    const float xo1 = (32 + lane < a.k) ? static_cast<float>(xd1) : 0.0f;
    a.mine[row * KP + lane] = xo0;
    a.mine[row * KP + 32 + lane] = xo1;
    for (int pr = 0; pr < a.n_peers; ++pr) {      // fused exchange: two 128-byte NVLink stores per peer replica
      a.peers[pr][row * KP + lane] = xo0;
      a.peers[pr][row * KP + 32 + lane] = xo1;
    }
    __syncwarp();
  }
}

template <bool WEIGHTED>
int launch_k64o(rl_context* h, const AlsArgs& args) {
  auto kern = als_k64o_kernel<WEIGHTED>;
  constexpr int W = K64OSmem<WEIGHTED>::warps;
  const size_t smem = K64OSmem<WEIGHTED>::bytes * W;
  RL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  RL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, smem));
  if (per_sm < 1) return fail(h, RL_ERR_CUDA, "als_k64: kernel does not fit on an SM (smem %zu)", smem);
  int64_t want = (args.n_rows + W - 1) / W;
  int64_t grid = static_cast<int64_t>(h->sm_count) * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  RL_CUDA(h, cudaMemsetAsync(args.counter, 0, sizeof(unsigned int), h->stream));
  kern<<<static_cast<unsigned>(grid), W * 32, smem, h->stream>>>(args);
  RL_LAUNCH_CHECK(h, "als_k64o_kernel");
  return RL_OK;
}
