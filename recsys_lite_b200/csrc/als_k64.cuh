// als_k64.cuh -- ALS half-step, fp32 + refinement path for padded k = 64 (the C3 / C4 configurations).
// Included by als.cu inside its anonymous namespace (uses AlsArgs and the cp.async helpers).
//
// Same algorithm as als_half_step_kernel (reference algo.py:314-322 / :324-332): per row with observations S,
//   x = (Y_S^T Y_S + lambda I)^-1 Y_S^T 1,
// one warp per row.  Everything that is a matrix product runs on the tensor cores (mma.sync m16n8k8, every
// operand split x = hi + lo with hi = upper 19 bits, products hi.hi + hi.lo + lo.hi: fp32-accurate):
//
//   * Gram: per 8 gathered rows the 64 x 64 update Y_c^T Y_c is 20 lower-triangle 16 x 8 tiles whose A and B
//     fragments are the SAME 16 registers (A = Y_c^T, B = Y_c).  The right-hand side sum is kept in float64.
//   * Cholesky, blocked left-looking and in place in shared memory, block columns of 8:
//     A[:, J] -= L[:, 0:8J] L[J, 0:8J]^T   is J k-steps of the same m16n8k8 per 16 x 8 tile (fragments read back from
//     the factor), the 8-column panel goes through shared memory to turn "mma layout" into "one row per lane"
//     and is factored row-locally (Crout inside the panel: lane i needs the finished row c of the 8 x 8 diagonal
//     block for column c, eight shuffle rounds per panel).  The forward substitution of the right-hand side rides
//     along while the panel rows are in registers.  ~2.4 K instructions per system instead of ~5.8 K for the
//     register-window rank-1 version, and one copy of the code for all block columns (instruction cache).
//   * the factor is stored panel by panel (block column K: rows 8K .. 63, 8 floats per row), the two 16-byte halves
//     of a row swapped when bit 2 of the row index is set and panel bases skewed by 8 floats: conflict-free for
//     the mma fragment loads (8 rows x 4 columns), for row loads (one row per lane, LDS.128) and for "row i across
//     all columns" (the backward substitution).
//   * refinement: x += L^-T L^-1 (b - A x) with the residual evaluated in float64 directly from the re-gathered
//     rows (matrix-free), ir_steps times.
//
// rsqrt.approx pivots are good enough: the factor is only a preconditioner for the float64-residual refinement.

constexpr int K64 = 64;
constexpr int K64_LDY = 68;   // row stride of a gathered row: the fragment reads (rows 2t, 2t+1; columns g) hit banks 8 t + g (+4)
constexpr int K64_CH = 8;     // gathered rows per chunk (one k = 8 mma step), double buffered
constexpr int K64_LSZ = 2368; // floats of the factor storage (8 panels, 8 floats of skew each)

// element (i, j) of the factor, i >= 8 (j >> 3):  Lt[k64_pb0(j >> 3) + 8 i + ((j & 7) ^ k64_swz(i))]
__host__ __device__ constexpr int k64_pb0(int K) { return K * (488 - 32 * K); }
__host__ __device__ constexpr int k64_swz(int i) { return ((i >> 2) & 1) << 2; }
static_assert(k64_pb0(7) + 8 * 63 + 7 < K64_LSZ && k64_pb0(1) + 8 * 8 == 520, "factor storage layout");

template <bool WEIGHTED>
struct K64Smem {
  static constexpr size_t lt_bytes = sizeof(float) * K64_LSZ;
  static constexpr size_t invd_bytes = sizeof(float) * K64;
  static constexpr size_t xs_bytes = sizeof(double) * K64;       // x in float64 (residual pass); aliased by the float right-hand side
  static constexpr size_t ts_bytes = sizeof(double) * 16;
  static constexpr size_t y_bytes = sizeof(float) * 2 * K64_CH * K64_LDY;
  static constexpr size_t w_bytes = WEIGHTED ? sizeof(float2) * 2 * K64_CH : 0;
  static constexpr size_t bytes = lt_bytes + invd_bytes + xs_bytes + ts_bytes + y_bytes + w_bytes;
  static constexpr int warps = 3;   // 5 CTAs per SM: 5 x (3 x 14.6 KB + 1 KB) < 227 KB, registers <= 136
  static constexpr int ctas = 5;
  static_assert(bytes % 16 == 0, "per-warp shared memory must keep 16-byte alignment");
};

// cp.async gather of chunk c (16-byte pieces); a short last chunk is zero-filled up to 8 rows for the k = 8 mma
template <bool WEIGHTED>
__device__ __forceinline__ void k64_gather(const AlsArgs& a, int64_t s, int nnz, int c, float* yb, float2* wb, int lane) {
  const int base = c * K64_CH;
  const int cnt = min(K64_CH, nnz - base);
  int my = 0;
  if (lane < cnt) {
    my = a.indices[s + base + lane];
    if constexpr (WEIGHTED) {
      const float cf = a.conf ? a.conf[s + base + lane] : 1.0f;
      const float pf = a.pref ? a.pref[s + base + lane] : 1.0f;
      wb[lane] = make_float2(cf - static_cast<float>(a.w0), cf * pf);
    }
  }
  constexpr int PPR = K64 / 4;  // 16-byte pieces per gathered row
#pragma unroll
  for (int p0 = 0; p0 < K64_CH * PPR; p0 += 32) {
    const int p = p0 + lane;
    const int r = p / PPR;
    const int idx = __shfl_sync(FULL_MASK, my, r);
    const int c4 = p - r * PPR;
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(yb + r * K64_LDY + 4 * c4));
    if (r < cnt) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(a.other + static_cast<int64_t>(idx) * K64 + 4 * c4));
    } else {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 0;\n" ::"r"(dst), "l"(a.other));  // src-size 0: zero fill
    }
  }
  cp_async_commit();
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3,
                                         const uint32_t b0, const uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
// (x0, x1) -> fp16 pairs hi = fp16(x), lo = fp16(x - hi): 22 significant bits, first element in the low half
__device__ __forceinline__ void k64_split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// lower-triangle 16 x 8 tiles of a 64 x 64 symmetric matrix: tile (mi, nj) is needed iff nj <= 2 mi + 1  ->  2 + 4 + 6 + 8 = 20
__device__ __forceinline__ constexpr int k64_tile(int mi, int nj) { return mi * (mi + 1) + nj; }

__device__ __forceinline__ void k64_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void k64_split_neg(float x, uint32_t& hi, uint32_t& lo) {   // parts of -x
  const uint32_t h = __float_as_uint(x) & 0xFFFFE000u;
  hi = h ^ 0x80000000u;
  lo = __float_as_uint(__uint_as_float(h) - x);
}

// One panel (block column J): rows 8J .. 63 of columns 8J .. 8J+7 sit in shared memory.  Lane l takes rows 8J + l and
// (TWO, i.e. J < 4) 8J + 32 + l.  FACTOR: the rows hold updated values of A and are factored against the 8 x 8 diagonal
// block (rows of lanes 0..7; Crout inside the panel), the factor is written back and 1 / L_cc goes to invd.  In both
// modes the right-hand side bs[] is carried along:  z_c = b_c / L_cc,  b_i -= L_ic z_c  (forward substitution).
template <bool TWO, bool FACTOR>
__device__ __forceinline__ void k64_panel(float* __restrict__ Lt, float* __restrict__ invd, float* __restrict__ bs, int J, int lane) {
  const int NR = 64 - 8 * J;                   // rows in this panel
  float* pj = Lt + k64_pb0(J) + 64 * J;        // row r of the panel (global row 8J + r) at pj + 8 r
  const int r0 = min(lane, NR - 1);            // lanes beyond the panel shadow its last row (their results are dropped)
  const bool v0 = lane < NR;
  const int r1 = TWO ? min(lane + 32, NR - 1) : 0;
  const bool v1 = TWO && lane + 32 < NR;
  const int s0 = k64_swz(r0), s1 = k64_swz(r1);   // 8J is a multiple of 8: the swizzle bit of row 8J + r is that of r
  float a[8], c1[8], iv[8];
  {
    const float4 lo = *reinterpret_cast<const float4*>(pj + 8 * r0 + s0);
    const float4 hi = *reinterpret_cast<const float4*>(pj + 8 * r0 + (s0 ^ 4));
    a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w; a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
  }
  if (TWO) {
    const float4 lo = *reinterpret_cast<const float4*>(pj + 8 * r1 + s1);
    const float4 hi = *reinterpret_cast<const float4*>(pj + 8 * r1 + (s1 ^ 4));
    c1[0] = lo.x; c1[1] = lo.y; c1[2] = lo.z; c1[3] = lo.w; c1[4] = hi.x; c1[5] = hi.y; c1[6] = hi.z; c1[7] = hi.w;
  }
  if (!FACTOR) {
    const float4 i0 = *reinterpret_cast<const float4*>(invd + 8 * J);
    const float4 i1 = *reinterpret_cast<const float4*>(invd + 8 * J + 4);
    iv[0] = i0.x; iv[1] = i0.y; iv[2] = i0.z; iv[3] = i0.w; iv[4] = i1.x; iv[5] = i1.y; iv[6] = i1.z; iv[7] = i1.w;
  }
  float bb0 = bs[8 * J + r0];
  float bb1 = TWO ? bs[8 * J + r1] : 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float inv;
    if (FACTOR) {
      // column c: v = a_c - sum_{c' < c} a_c' D[c][c'], D = finished rows of the diagonal block (lane c holds row c)
      float acc0 = a[c], acc1 = TWO ? c1[c] : 0.f;
#pragma unroll
      for (int cp = 0; cp < c; ++cp) {
        const float d = __shfl_sync(FULL_MASK, a[cp], c);
        acc0 = fmaf(-a[cp], d, acc0);
        if (TWO) acc1 = fmaf(-c1[cp], d, acc1);
      }
      const float piv = __shfl_sync(FULL_MASK, acc0, c);
      inv = rsqrtf(fmaxf(piv, 1e-30f));
      a[c] = acc0 * inv;                           // lane c: piv * rsqrt(piv) = L_cc
      if (TWO) c1[c] = acc1 * inv;
      iv[c] = inv;
    } else {
      inv = iv[c];
    }
    const float zc = __shfl_sync(FULL_MASK, bb0, c) * inv;
    if (lane == c) bb0 = zc;
    else if (lane > c) bb0 = fmaf(-a[c], zc, bb0);
    if (TWO) bb1 = fmaf(-c1[c], zc, bb1);
  }
  if (FACTOR) {
    if (lane == 0) {
      *reinterpret_cast<float4*>(invd + 8 * J) = make_float4(iv[0], iv[1], iv[2], iv[3]);
      *reinterpret_cast<float4*>(invd + 8 * J + 4) = make_float4(iv[4], iv[5], iv[6], iv[7]);
    }
    if (v0) {
      *reinterpret_cast<float4*>(pj + 8 * r0 + s0) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(pj + 8 * r0 + (s0 ^ 4)) = make_float4(a[4], a[5], a[6], a[7]);
    }
    if (v1) {
      *reinterpret_cast<float4*>(pj + 8 * r1 + s1) = make_float4(c1[0], c1[1], c1[2], c1[3]);
      *reinterpret_cast<float4*>(pj + 8 * r1 + (s1 ^ 4)) = make_float4(c1[4], c1[5], c1[6], c1[7]);
    }
  }
  if (v0) bs[8 * J + r0] = bb0;
  if (v1) bs[8 * J + r1] = bb1;
  __syncwarp();
}

// d[mi] -= L[16 mi .. 16 mi + 15][0 : 8J] L[8J .. 8J + 7][0 : 8J]^T for the tiles mi >= M0 of block column J
template <int M0>
__device__ __forceinline__ void k64_update(float (&d)[4][4], const float* __restrict__ Lt, int J, int g, int o0, int o1) {
#pragma unroll 1
  for (int K = 0; K < J; ++K) {
    const float* pk = Lt + K * (488 - 32 * K);
    uint32_t bh[2], bl[2];                     // B(k, n) = L[8J + n][8K + k]
    k64_split(pk[8 * (8 * J + g) + o0], bh[0], bl[0]);
    k64_split(pk[8 * (8 * J + g) + o1], bh[1], bl[1]);
    uint32_t ah[4][4], al[4][4];               // A(m, k) = -L[16 mi + m][8K + k]
#pragma unroll
    for (int mi = M0; mi < 4; ++mi) {
      const float* pa = pk + 8 * (16 * mi + g);
      k64_split_neg(pa[o0], ah[mi][0], al[mi][0]);
      k64_split_neg(pa[64 + o0], ah[mi][1], al[mi][1]);
      k64_split_neg(pa[o1], ah[mi][2], al[mi][2]);
      k64_split_neg(pa[64 + o1], ah[mi][3], al[mi][3]);
    }
    // term by term across the tiles: consecutive HMMAs are independent
#pragma unroll
    for (int mi = M0; mi < 4; ++mi) mma_tf32(d[mi], ah[mi][0], ah[mi][1], ah[mi][2], ah[mi][3], bh[0], bh[1]);
#pragma unroll
    for (int mi = M0; mi < 4; ++mi) mma_tf32(d[mi], ah[mi][0], ah[mi][1], ah[mi][2], ah[mi][3], bl[0], bl[1]);
#pragma unroll
    for (int mi = M0; mi < 4; ++mi) mma_tf32(d[mi], al[mi][0], al[mi][1], al[mi][2], al[mi][3], bh[0], bh[1]);
  }
}

// Blocked left-looking Cholesky, in place in the panel storage (which holds A on entry, written there from the Gram
// accumulators); L -> Lt, 1 / L_jj -> invd, and bs <- L^-1 bs on the way.  One copy of the code for all block columns
// (the kernel is instruction-cache bound otherwise: every warp of an SM is in a different phase).
// nblk < 8: only the leading 8 nblk columns are factored (callers whose trailing rows / columns are an identity block).
__device__ __forceinline__ void k64_chol(float* __restrict__ Lt, float* __restrict__ invd, float* __restrict__ bs, int lane, int nblk = 8) {
  const int g = lane >> 2, tq = lane & 3;
  const int sg = ((g >> 2) & 1) << 2;          // swizzle bit of rows 8 m + g
  const int o0 = tq ^ sg, o1 = o0 ^ 4;         // physical positions of panel columns tq and tq + 4 in such a row
  const int oc = (2 * tq) ^ sg;                // ... of the accumulator columns 2 tq, 2 tq + 1
#pragma unroll 1
  for (int J = 0; J < nblk; ++J) {
    float* pj = Lt + k64_pb0(J);
    const int m0 = J >> 1;
    float d[4][4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      if (mi >= m0) {       // rows 16 mi + g of the odd-J first tile lie above the panel: junk in, never stored
        const float2 t0 = *reinterpret_cast<const float2*>(pj + 8 * (16 * mi + g) + oc);
        const float2 t1 = *reinterpret_cast<const float2*>(pj + 8 * (16 * mi + 8 + g) + oc);
        d[mi][0] = t0.x; d[mi][1] = t0.y; d[mi][2] = t1.x; d[mi][3] = t1.y;
      }
    }
    switch (m0) {      // the set of 16-row tiles below the diagonal is fixed per block column: no predicates in the K loop
      case 0: k64_update<0>(d, Lt, J, g, o0, o1); break;
      case 1: k64_update<1>(d, Lt, J, g, o0, o1); break;
      case 2: k64_update<2>(d, Lt, J, g, o0, o1); break;
      default: k64_update<3>(d, Lt, J, g, o0, o1); break;
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      if (mi >= m0) {
        if (mi > m0 || !(J & 1)) *reinterpret_cast<float2*>(pj + 8 * (16 * mi + g) + oc) = make_float2(d[mi][0], d[mi][1]);
        *reinterpret_cast<float2*>(pj + 8 * (16 * mi + 8 + g) + oc) = make_float2(d[mi][2], d[mi][3]);
      }
    }
    __syncwarp();
    if (J < 4) k64_panel<true, true>(Lt, invd, bs, J, lane);
    else k64_panel<false, true>(Lt, invd, bs, J, lane);
  }
}

// bs <- L^-1 bs with the stored factor (the sweep of k64_panel with the rows loaded instead of computed)
__device__ __forceinline__ void k64_forward(float* __restrict__ Lt, float* __restrict__ invd, float* __restrict__ bs, int lane, int nblk = 8) {
#pragma unroll 1
  for (int J = 0; J < 4 && J < nblk; ++J) k64_panel<true, false>(Lt, invd, bs, J, lane);
#pragma unroll 1
  for (int J = 4; J < nblk; ++J) k64_panel<false, false>(Lt, invd, bs, J, lane);
}

// v <- L^-T v   (v0: entry l, v1: entry l + 32), row i of the factor read across the lanes.
// Step i: x_i = w_i / L_ii, then w_j -= L_ij x_i for j < i.  Lane i's entry is final (as w_i) from step i on, so the
// division is left to the end (one multiply per lane instead of a select per step) and the updates are predicated on
// j < i (the storage above the diagonal holds junk).  Blocks of eight rows: the swizzle bit of row i is a constant of the
// unrolled body.
// nblk < 8: entries from 8 nblk on are taken as zero (not read) and returned as zero.
__device__ __forceinline__ void k64_backward(const float* __restrict__ Lt, const float* __restrict__ invd, float& v0, float& v1, int lane,
                                             int nblk = 8) {
  const float* c0 = Lt + k64_pb0(lane >> 3);            // L[i][l]      = c0[8 i + ((l & 7) ^ swz(i))]
  const float* c1 = Lt + k64_pb0(4 + (lane >> 3));      // L[i][l + 32] = c1[8 i + ((l & 7) ^ swz(i))]
  const int q0 = lane & 7, q1 = q0 ^ 4;
  const float inv0 = lane < 8 * nblk ? invd[lane] : 0.f, inv1 = 32 + lane < 8 * nblk ? invd[32 + lane] : 0.f;
#pragma unroll 1
  for (int b = nblk - 1; b >= 4; --b) {
    const float* r1 = c1 + 64 * b;
    const float* r0 = c0 + 64 * b;
#pragma unroll
    for (int r = 7; r >= 0; --r) {
      const int i = 8 * b + r;
      const float xi = __shfl_sync(FULL_MASK, v1 * inv1, i - 32);     // the owner lane's product is the one that counts
      const float l1 = r1[8 * r + ((r & 4) ? q1 : q0)];
      const float l0 = r0[8 * r + ((r & 4) ? q1 : q0)];
      if (lane < i - 32) v1 = fmaf(-l1, xi, v1);
      v0 = fmaf(-l0, xi, v0);
    }
  }
#pragma unroll 1
  for (int b = min(nblk - 1, 3); b >= 0; --b) {
    const float* r0 = c0 + 64 * b;
#pragma unroll
    for (int r = 7; r >= 0; --r) {
      const int i = 8 * b + r;
      const float xi = __shfl_sync(FULL_MASK, v0 * inv0, i);
      const float l0 = r0[8 * r + ((r & 4) ? q1 : q0)];               // lanes >= i read below their panel: finite junk, not used
      if (lane < i) v0 = fmaf(-l0, xi, v0);
    }
  }
  v0 *= inv0;
  v1 *= inv1;
}

// rows longer than this are solved by a whole CTA (als_k64_kernel<.., K64_COOP_WARPS>) instead of one warp: on power-law
// data a single row can hold a large part of all observations, and one warp gathers ~6 M rows/s
constexpr int K64_LONG = 2048;
constexpr int K64_COOP_WARPS = 16;
// ... and rows longer than this by a thread-block cluster of K64_CLUSTER such CTAs (partial sums meet in the leader CTA's
// shared memory over DSMEM): one CTA gathers ~100 M rows/s, a row with a million observations would hold it for 20 ms
constexpr int K64_GIANT = 32768;
constexpr int K64_CLUSTER = 8;

// the rows the cooperative kernels work on: long rows in list[0 .. count[0]), giant rows in list[n_rows - count[1] .. n_rows)
__global__ void __launch_bounds__(256) k64_find_long_rows_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, int32_t* __restrict__ list,
                                                                 unsigned int* __restrict__ count) {
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < n_rows; r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t d = indptr[r + 1] - indptr[r];
    if (d > K64_GIANT) list[n_rows - 1 - atomicAdd(count + 1, 1u)] = static_cast<int32_t>(r);
    else if (d > K64_LONG) list[atomicAdd(count, 1u)] = static_cast<int32_t>(r);
  }
}

// NW = 1: one warp per row, independent warps (the common case).  NW > 1: the NW warps of a CTA share one (long) row:
// every warp takes the chunks c = w (mod NW), partial Gram tiles and residuals meet in shared memory (atomics), warp 0
// factors and solves.
// NCTA > 1 (with NW > 1): the CTAs of a thread-block cluster share one (giant) row the same way; their totals are added
// into the leader CTA's shared memory through distributed shared memory, cluster barriers in place of __syncthreads.
template <bool WEIGHTED, int NW, int NCTA>
__global__ void __launch_bounds__(NW == 1 ? K64Smem<WEIGHTED>::warps * 32 : NW * 32, NW == 1 ? K64Smem<WEIGHTED>::ctas : 1)
als_k64_kernel(const AlsArgs a) {
  constexpr int KP = K64, CH = K64_CH, LDY = K64_LDY;
  constexpr bool COOP = NW > 1;
  constexpr bool CLUSTER = NCTA > 1;
  static_assert(!CLUSTER || COOP, "clusters only in the cooperative mode");
  namespace cg = cooperative_groups;
  const unsigned int crank = CLUSTER ? cg::this_cluster().block_rank() : 0u;   // 0: the leader CTA of the cluster
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // solo: [Lt | invd | xs | ts | y | w] per warp.  coop: [Lt | invd | xs | rs] once, then [ts | y | w] per warp.
  constexpr size_t SH_BYTES = K64Smem<WEIGHTED>::lt_bytes + K64Smem<WEIGHTED>::invd_bytes + K64Smem<WEIGHTED>::xs_bytes;
  constexpr size_t PW_BYTES = K64Smem<WEIGHTED>::ts_bytes + K64Smem<WEIGHTED>::y_bytes + K64Smem<WEIGHTED>::w_bytes;
  unsigned char* shb = COOP ? smem_raw : smem_raw + static_cast<size_t>(warp) * K64Smem<WEIGHTED>::bytes;
  unsigned char* pwb = COOP ? smem_raw + SH_BYTES + sizeof(double) * K64 + static_cast<size_t>(warp) * PW_BYTES : shb + SH_BYTES;
  float* Lt = reinterpret_cast<float*>(shb);
  float* invd = Lt + K64_LSZ;
  double* xs = reinterpret_cast<double*>(invd + K64);
  float* bs = reinterpret_cast<float*>(xs);               // the float right-hand side of a solve lives in the same 512 bytes
  double* rs = xs + K64;                                  // coop only: residual accumulator
  double* ts = reinterpret_cast<double*>(pwb);
  float* ybuf = reinterpret_cast<float*>(ts + 16);
  float2* wbuf = reinterpret_cast<float2*>(ybuf + 2 * CH * LDY);
  __shared__ unsigned int s_ticket;

  const float lam = static_cast<float>(a.lambda);
  const bool hkv = WEIGHTED && a.w0 != 0.0;   // w0 != 0 always runs the WEIGHTED instantiation
  // fp16 split of the Gram operands: a power-of-two scale puts max |y| at <= 2^14 (the largest |entry| of `other` was
  // reduced by absmax_kernel just before this launch); Gram tiles are unscaled by its exact inverse square
  const float ysc = pow2_scale(*a.other_absmax);
  const float yunsc = (1.0f / ysc) * (1.0f / ysc);
  const int g = lane >> 2, tq = lane & 3;
  const bool lead = !COOP || (warp == 0 && crank == 0);   // the warp that factors and solves
  // the leader CTA's copies of the shared accumulators (DSMEM)
  float* Lt0 = Lt;
  float* bs0 = bs;
  double* xs0 = xs;
  double* rs0 = rs;
  if constexpr (CLUSTER) {
    cg::cluster_group cluster = cg::this_cluster();
    Lt0 = cluster.map_shared_rank(Lt, 0);
    bs0 = cluster.map_shared_rank(bs, 0);
    xs0 = cluster.map_shared_rank(xs, 0);
    rs0 = cluster.map_shared_rank(rs, 0);
  }
  auto block_or_cluster_sync = [&]() {
    if constexpr (CLUSTER) cg::this_cluster().sync(); else __syncthreads();
  };
  unsigned int giant_pos = CLUSTER ? blockIdx.x / NCTA : 0u;   // cluster c takes giant rows c, c + #clusters, ...

  for (;;) {
    int64_t row;
    if constexpr (!COOP) {
      unsigned int ticket = 0;
      if (lane == 0) ticket = atomicAdd(a.counter, 1u);
      ticket = __shfl_sync(FULL_MASK, ticket, 0);
      if (static_cast<int64_t>(ticket) >= a.n_rows) break;
      row = a.row_order ? a.row_order[ticket] : static_cast<int64_t>(ticket);
    } else {
      if constexpr (CLUSTER) {
        if (giant_pos >= a.long_count[1]) break;
        row = a.long_list[a.n_rows - 1 - giant_pos];
        giant_pos += gridDim.x / NCTA;
      } else {
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = atomicAdd(a.counter + 1, 1u);
        __syncthreads();
        if (s_ticket >= a.long_count[0]) break;
        row = a.long_list[s_ticket];
      }
    }
    const int64_t s = a.indptr[row];
    const int64_t nnz64 = a.indptr[row + 1] - s;
    if constexpr (!COOP) {
      if (a.split_long && nnz64 > K64_LONG) continue;   // the cooperative kernel has it
    }
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
    const int nchunk_all = (nnz + CH - 1) / CH;
    // chunks of this warp: w, w + NW, ...
    constexpr int NWT = NW * NCTA;               // warps sharing the row
    const int gw = COOP ? static_cast<int>(crank) * NW + warp : 0;
    const int nchunk = COOP ? (nchunk_all > gw ? (nchunk_all - gw + NWT - 1) / NWT : 0) : nchunk_all;
    auto chunk_id = [&](int c) { return COOP ? gw + c * NWT : c; };
    if constexpr (COOP) {                       // the accumulators of the row start from zero
      for (int p = threadIdx.x; p < K64_LSZ; p += NW * 32) Lt[p] = 0.f;
      if (threadIdx.x < K64) bs[threadIdx.x] = 0.f;
      block_or_cluster_sync();
    }

    // ---- gather + Gram on the tensor cores (fp16 split operands, fp32 accumulate) --------------------------------
    float acc[20][4];
    float bf0 = 0.f, bf1 = 0.f;      // right-hand side in fp32: only the first solve uses it, the refinement residual is b-free
#pragma unroll
    for (int t = 0; t < 20; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
    for (int c = -1; c < nchunk; ++c) {       // one gather call site: chunk c + 1 is requested, then chunk c is consumed
      const int buf = c & 1;
      if (c + 1 < nchunk) {
        k64_gather<WEIGHTED>(a, s, nnz, chunk_id(c + 1), ybuf + (buf ^ 1) * CH * LDY, wbuf + (buf ^ 1) * CH, lane);
        if (c < 0) continue;
        cp_async_wait<1>();
      } else {
        if (c < 0) continue;
        cp_async_wait<0>();
      }
      __syncwarp();
      const float* yb = ybuf + buf * CH * LDY;
      const float2* wb = wbuf + buf * CH;
      const int cnt = min(CH, nnz - chunk_id(c) * CH);
#pragma unroll
      for (int r = 0; r < CH; ++r) {        // zero-filled rows add nothing
        const float* yr = yb + r * LDY;
        if constexpr (WEIGHTED) {
          const float wr = (r < cnt) ? wb[r].y : 0.f;
          bf0 = fmaf(wr, yr[lane], bf0);
          bf1 = fmaf(wr, yr[32 + lane], bf1);
        } else {
          bf0 += yr[lane];
          bf1 += yr[32 + lane];
        }
      }
      {
        // m16n8k8 fp16 fragments, k = gathered row: A(mi) = (a0, a1) = Y[2 tq, 2 tq + 1][16 mi + g (+8)] packed in pairs;
        // B(nj) = Y[2 tq, 2 tq + 1][8 nj + g] is ONE of those registers: a[nj >> 1][nj & 1] -- no operand moves.
        uint32_t hi[4][2], lo[4][2];
        const float* y0 = yb + (2 * tq) * LDY + g;
        const float* y1 = y0 + LDY;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          k64_split_h2(y0[16 * mi] * ysc, y1[16 * mi] * ysc, hi[mi][0], lo[mi][0]);
          k64_split_h2(y0[16 * mi + 8] * ysc, y1[16 * mi + 8] * ysc, hi[mi][1], lo[mi][1]);
        }
        uint32_t ahi[4][2], alo[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int e = 0; e < 2; ++e) { ahi[mi][e] = hi[mi][e]; alo[mi][e] = lo[mi][e]; }
        if constexpr (WEIGHTED) {     // scale the A side by the Gram weight of its gathered row (rows 2 tq and 2 tq + 1)
          const float w0 = (2 * tq < cnt) ? wb[2 * tq].x * ysc : 0.f;
          const float w1 = (2 * tq + 1 < cnt) ? wb[2 * tq + 1].x * ysc : 0.f;
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) {
            k64_split_h2(y0[16 * mi] * w0, y1[16 * mi] * w1, ahi[mi][0], alo[mi][0]);
            k64_split_h2(y0[16 * mi + 8] * w0, y1[16 * mi + 8] * w1, ahi[mi][1], alo[mi][1]);
          }
        }
        // term by term: consecutive HMMAs are independent
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int nj = 0; nj <= 2 * mi + 1; ++nj) mma_f16(acc[k64_tile(mi, nj)], ahi[mi][0], ahi[mi][1], hi[nj >> 1][nj & 1]);
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int nj = 0; nj <= 2 * mi + 1; ++nj) mma_f16(acc[k64_tile(mi, nj)], ahi[mi][0], ahi[mi][1], lo[nj >> 1][nj & 1]);
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int nj = 0; nj <= 2 * mi + 1; ++nj) mma_f16(acc[k64_tile(mi, nj)], alo[mi][0], alo[mi][1], hi[nj >> 1][nj & 1]);
      }
      __syncwarp();
    }

    // ---- + lambda I (+ w0 Y^T Y), accumulator tiles -> panel storage: element e of tile (mi, nj) is
    //      A[16 mi + g + 8 (e >> 1)][8 nj + 2 tq + (e & 1)]; the first 8 rows of the odd tiles nj = 2 mi + 1 lie above the diagonal
    // (cooperative modes: the warps, then the CTAs of the cluster, add their parts one after the other -- a fixed
    //  order, so that the result does not depend on timing)
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
            d[e] *= yunsc;                         // the operands were scaled by ysc (a power of two): exact
            if (lead) {
              if (hkv) d[e] += static_cast<float>(a.w0 * a.gram0[gi * KP + gj]);
              if (nj >= 2 * mi && gi == gj) d[e] += (gi < a.k) ? lam : 1.0f;
            }
          }
        }
#pragma unroll 1
      for (int w = 0; w < NW; ++w) {
        if (!COOP || warp == w) {
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int nj = 0; nj <= 2 * mi + 1; ++nj) {
              const float (&d)[4] = acc[k64_tile(mi, nj)];
              float2* p0 = reinterpret_cast<float2*>(Lt + k64_pb0(nj) + 8 * (16 * mi + g) + oc);
              float2* p1 = reinterpret_cast<float2*>(Lt + k64_pb0(nj) + 8 * (16 * mi + 8 + g) + oc);
              if constexpr (!COOP) {
                if (nj <= 2 * mi) *p0 = make_float2(d[0], d[1]);
                *p1 = make_float2(d[2], d[3]);
              } else {
                if (nj <= 2 * mi) { const float2 o = *p0; *p0 = make_float2(o.x + d[0], o.y + d[1]); }
                const float2 o = *p1;
                *p1 = make_float2(o.x + d[2], o.y + d[3]);
              }
            }
          if constexpr (COOP) { bs[lane] += bf0; bs[32 + lane] += bf1; }
        }
        if constexpr (COOP) __syncthreads();
      }
    }
    // ---- blocked Cholesky on the tensor cores, forward substitution fused ------------------------------------
    if constexpr (!COOP) {
      bs[lane] = bf0;
      bs[32 + lane] = bf1;
      __syncwarp();
    } else if constexpr (CLUSTER) {             // every CTA's totals -> the leader's, CTA after CTA (DSMEM)
      cg::this_cluster().sync();                // the leader's own warps are done with its copy
#pragma unroll 1
      for (unsigned int cta = 1; cta < NCTA; ++cta) {
        if (crank == cta) {
          for (int p = threadIdx.x; p < K64_LSZ; p += NW * 32) Lt0[p] += Lt[p];
          if (threadIdx.x < K64) bs0[threadIdx.x] += bs[threadIdx.x];
        }
        cg::this_cluster().sync();
      }
    }
    if (lead) k64_chol(Lt, invd, bs, lane);
    // ---- back substitution, then refinement steps  x += L^-T L^-1 (b - A x)  with a float64 matrix-free residual ----
    double xd0 = 0.0, xd1 = 0.0;
    for (int it = 0;; ++it) {
      if (lead) {
        float v0 = bs[lane], v1 = bs[32 + lane];    // L^-1 (right-hand side or residual)
        __syncwarp();
        k64_backward(Lt, invd, v0, v1, lane);
        xd0 += static_cast<double>(v0);
        xd1 += static_cast<double>(v1);
      }
      if (it >= a.ir_steps) break;
      if (lead) {
        xs[lane] = xd0;
        xs[32 + lane] = xd1;
      }
      if constexpr (COOP) { if (warp == 0) { rs[lane] = 0.0; rs[32 + lane] = 0.0; } }
      if constexpr (COOP) block_or_cluster_sync(); else __syncwarp();
      if constexpr (CLUSTER) {                  // every CTA works from its own copy of x
        if (crank != 0 && threadIdx.x < K64) xs[threadIdx.x] = xs0[threadIdx.x];
        __syncthreads();
      }
      // r = b - A x = sum_i (p_i - w_i <y_i, x>) y_i - lambda x (- w0 G x): the right-hand side never appears on its own
      double r0 = 0.0, r1 = 0.0;
      if (lead) {
        r0 = -((lane < a.k) ? a.lambda : 1.0) * xd0;
        r1 = -((32 + lane < a.k) ? a.lambda : 1.0) * xd1;
      }
      if (hkv && lead) {
        double t0 = 0.0, t1 = 0.0;
        for (int cc = 0; cc < KP; ++cc) {
          t0 = fma(a.gram0[lane * KP + cc], xs[cc], t0);
          t1 = fma(a.gram0[(32 + lane) * KP + cc], xs[cc], t1);
        }
        r0 -= a.w0 * t0;
        r1 -= a.w0 * t1;
      }
      for (int c = -1; c < nchunk; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunk) {
          k64_gather<WEIGHTED>(a, s, nnz, chunk_id(c + 1), ybuf + (buf ^ 1) * CH * LDY, wbuf + (buf ^ 1) * CH, lane);
          if (c < 0) continue;
          cp_async_wait<1>();
        } else {
          if (c < 0) continue;
          cp_async_wait<0>();
        }
        __syncwarp();
        const float* yb = ybuf + buf * CH * LDY;
        const float2* wb = wbuf + buf * CH;
        const int cnt = min(CH, nnz - chunk_id(c) * CH);
        // t_i = w_i <y_i, x>: four lanes per gathered row (a quarter of the factor dimension each), then two shuffle adds
        {
          const int i = lane & 7, quarter = lane >> 3;
          const float* yr = yb + i * LDY + 16 * quarter;
          const double* xh = xs + 16 * quarter;
          double t0 = 0.0, t1 = 0.0;
#pragma unroll
          for (int cc = 0; cc < 16; cc += 4) {
            const float4 y4 = *reinterpret_cast<const float4*>(yr + cc);
            t0 = fma(static_cast<double>(y4.x), xh[cc], t0);
            t1 = fma(static_cast<double>(y4.y), xh[cc + 1], t1);
            t0 = fma(static_cast<double>(y4.z), xh[cc + 2], t0);
            t1 = fma(static_cast<double>(y4.w), xh[cc + 3], t1);
          }
          double t = t0 + t1;
          t += __shfl_xor_sync(FULL_MASK, t, 8);
          t += __shfl_xor_sync(FULL_MASK, t, 16);
          if constexpr (WEIGHTED) t = (i < cnt) ? static_cast<double>(wb[i].y) - static_cast<double>(wb[i].x) * t : 0.0;
          else t = (i < cnt) ? 1.0 - t : 0.0;
          if (lane < 8) ts[lane] = t;     // coefficient of y_i in the residual; zero-filled rows: 0
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const double ti = ts[i];
          r0 = fma(ti, static_cast<double>(yb[i * LDY + lane]), r0);
          r1 = fma(ti, static_cast<double>(yb[i * LDY + 32 + lane]), r1);
        }
        __syncwarp();
      }
      if constexpr (COOP) {                     // partial residuals: warp after warp, then CTA after CTA (fixed order)
#pragma unroll 1
        for (int w = 0; w < NW; ++w) {
          if (warp == w) { rs[lane] += r0; rs[32 + lane] += r1; }
          __syncthreads();
        }
        if constexpr (CLUSTER) {
          cg::this_cluster().sync();
#pragma unroll 1
          for (unsigned int cta = 1; cta < NCTA; ++cta) {
            if (crank == cta && threadIdx.x < K64) rs0[threadIdx.x] += rs[threadIdx.x];
            cg::this_cluster().sync();
          }
        }
        r0 = rs[lane];
        r1 = rs[32 + lane];
        __syncthreads();                        // everybody has read xs / rs before warp 0 reuses the bytes as bs
      }
      if (lead) {
        bs[lane] = static_cast<float>(r0);      // xs is dead from here on: its bytes take the right-hand side
        bs[32 + lane] = static_cast<float>(r1);
        __syncwarp();
        k64_forward(Lt, invd, bs, lane);
      }
    }
    if (lead) {
      const float xo0 = (lane < a.k) ? static_cast<float>(xd0) : 0.0f;
      const float xo1 = (32 + lane < a.k) ? static_cast<float>(xd1) : 0.0f;
      a.mine[row * KP + lane] = xo0;
      a.mine[row * KP + 32 + lane] = xo1;
      for (int pr = 0; pr < a.n_peers; ++pr) {      // fused exchange: two 128-byte NVLink stores per peer replica
        a.peers[pr][row * KP + lane] = xo0;
        a.peers[pr][row * KP + 32 + lane] = xo1;
      }
    }
    if constexpr (COOP) block_or_cluster_sync(); else __syncwarp();
  }
}

template <bool WEIGHTED>
int launch_k64(rl_context* h, const AlsArgs& args_in) {
  AlsArgs args = args_in;
  RL_CUDA(h, cudaMemsetAsync(args.counter, 0, 2 * sizeof(unsigned int), h->stream));
  if (args.split_long) {
    // Rows with more than K64_LONG observations: listed by a small kernel, then one CTA per row by ticket.  Both run on
    // the handle's second compute stream and are launched first, so that the cooperative kernel takes one CTA slot per
    // SM and the warp-per-row kernel fills the rest of the chip beside it (the two write disjoint rows).
    if (args.n_rows >= (1ll << 31)) return fail(h, RL_ERR_INVALID, "als_k64: too many rows for the long-row list");
    if (h->long_list_rows < args.n_rows) {
      RL_CUDA(h, cudaStreamSynchronize(h->aux_stream));
      if (h->long_list) RL_CUDA(h, cudaFree(h->long_list));
      h->long_list = nullptr;
      h->long_list_rows = 0;
      RL_CUDA(h, cudaMalloc(&h->long_list, sizeof(int32_t) * static_cast<size_t>(args.n_rows)));
      h->long_list_rows = args.n_rows;
    }
    args.long_list = h->long_list;
    args.long_count = args.counter + 2;
    RL_CUDA(h, cudaMemsetAsync(args.counter + 2, 0, 2 * sizeof(unsigned int), h->stream));
    RL_CUDA(h, cudaEventRecord(h->aux_fork, h->stream));
    RL_CUDA(h, cudaStreamWaitEvent(h->aux_stream, h->aux_fork, 0));
    int64_t gf = (args.n_rows + 255) / 256;
    if (gf > 4 * h->sm_count) gf = 4 * h->sm_count;
    k64_find_long_rows_kernel<<<static_cast<unsigned>(gf), 256, 0, h->aux_stream>>>(args.indptr, args.n_rows, h->long_list, args.counter + 2);
    RL_LAUNCH_CHECK(h, "k64_find_long_rows_kernel");
    auto coop = als_k64_kernel<WEIGHTED, K64_COOP_WARPS, 1>;
    auto giant = als_k64_kernel<WEIGHTED, K64_COOP_WARPS, K64_CLUSTER>;
    constexpr size_t sh = K64Smem<WEIGHTED>::lt_bytes + K64Smem<WEIGHTED>::invd_bytes + K64Smem<WEIGHTED>::xs_bytes + sizeof(double) * K64;
    constexpr size_t pw = K64Smem<WEIGHTED>::ts_bytes + K64Smem<WEIGHTED>::y_bytes + K64Smem<WEIGHTED>::w_bytes;
    const size_t smem = sh + pw * K64_COOP_WARPS;
    RL_CUDA(h, cudaFuncSetAttribute(coop, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    {  // giant rows first (clusters of K64_CLUSTER CTAs), then the long ones
      RL_CUDA(h, cudaFuncSetAttribute(giant, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(static_cast<unsigned>((h->sm_count / K64_CLUSTER) * K64_CLUSTER));
      cfg.blockDim = dim3(K64_COOP_WARPS * 32);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = h->aux_stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = K64_CLUSTER;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      RL_CUDA(h, cudaLaunchKernelEx(&cfg, giant, args));
      h->launches++;
    }
    coop<<<static_cast<unsigned>(h->sm_count), K64_COOP_WARPS * 32, smem, h->aux_stream>>>(args);
    RL_LAUNCH_CHECK(h, "als_k64_kernel(coop)");
    RL_CUDA(h, cudaEventRecord(h->aux_join, h->aux_stream));
  }
  auto kern = als_k64_kernel<WEIGHTED, 1, 1>;
  constexpr int W = K64Smem<WEIGHTED>::warps;
  const size_t smem = K64Smem<WEIGHTED>::bytes * W;
  RL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  RL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, smem));
  if (per_sm < 1) return fail(h, RL_ERR_CUDA, "als_k64: kernel does not fit on an SM (smem %zu)", smem);
  int64_t want = (args.n_rows + W - 1) / W;
  int64_t grid = static_cast<int64_t>(h->sm_count) * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  kern<<<static_cast<unsigned>(grid), W * 32, smem, h->stream>>>(args);
  RL_LAUNCH_CHECK(h, "als_k64_kernel");
  if (args.split_long) RL_CUDA(h, cudaStreamWaitEvent(h->stream, h->aux_join, 0));
  return RL_OK;
}
