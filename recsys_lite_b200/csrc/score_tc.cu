// score_tc.cu -- tensor-core scoring fused with seen-item masking and top-n selection (tcgen05 / TMEM / TMA).
//
// Replaces predict + top_n of the reference (algo.py:143-147 `U @ V.T`, algo.py:533-659) for ALS factors without
// ever writing the n_users x n_items score matrix.  One persistent CTA per SM owns 128 users at a time (one TMEM
// lane = one user) and sweeps all items in tiles of 128:
//
//   warp 0   TMA producer: U tile(s) (128 x KP fp32) once per user tile, V tiles (128 x KP fp32) through a
//            shared-memory slot ring (128B swizzle, K-major), mbarrier complete_tx signalling;
//   warp 1   MMA issuer: tcgen05.mma.kind::tf32 (M=128, N=128, K=8) straight from the fp32 tiles (the tensor core
//            reads the upper 19 bits of each operand), accumulators in TMEM, four 128-column stages;
//   warp 2   TMEM allocation (512 columns);
//   warps 4-7 epilogue: tcgen05.ld 32 columns at a time, 3-input max over groups of 8 scores and ONE compare
//            against the row threshold per group (the common case costs ~0.6 instructions per score); rare hits
//            are checked against the row's seen-item cursor (CSR, ascending) and appended to a per-row candidate
//            buffer in shared memory.
//
// Precision (TERMS).  A single tf32 product is only good to 2^-9 |u||v|: with 10^5 items the top scores of ALS
// factors are ~3e-5 apart, so the candidate window would hold hundreds of items.  TERMS = 3 is the split-operand
// scheme: x = x_hi + x_lo with x_hi = the 19 upper bits (what the tensor core reads anyway) and x_lo = x - x_hi
// (exact in fp32, precomputed once per call), and  u.v ~ u_hi.v_hi + u_lo.v_hi + u_hi.v_lo  accumulated in the
// same TMEM tile (3 x KP/8 MMAs per item tile).  The dropped u_lo.v_lo term and the re-truncation of the lo parts
// are each below 2^-20 |u||v|.
//
// Exactness.  The approximate scores s~ satisfy |s~ - s| <= eps_u = ERR * |u|_2 * max_i |v_i|_2 (ERR covers the
// operand terms above plus fp32 accumulation).  Each row tracks the n best approximate scores of UNSEEN items; a
// score is kept iff it exceeds (n-th best so far) - 2 eps_u.  Every item of the exact top-n therefore is among
// the candidates; they are re-scored in float64 (same summation order as the exact kernel) and ranked by (score
// desc, index asc).  Rows whose candidate buffer overflows are listed for the exact CUDA-core kernel.  The
// result is identical to rl_score_topn_dev(mode = RL_SCORE_EXACT).
#include <cuda.h>

#include "common.cuh"
#include "score_exact.cuh"

namespace rl {

namespace {

constexpr int TC_BM = 128;      // users per tile (TMEM lanes)
constexpr int TC_BN = 128;      // items per tile (TMEM columns per accumulator stage)
constexpr int TC_ACC = 4;       // TMEM accumulator stages (4 x 128 columns = all of TMEM)
constexpr int TC_SLOTS = 3;     // shared-memory slots of the V ring
constexpr int TC_THREADS = 256;
constexpr int TC_CAND = 40;     // candidate slots per row
constexpr int TC_NTOP = 16;     // largest n served by this path
// bounds on |tensor-core dot - exact dot| / (|u| |v|), checked by tests/test_gpu_parity.py::test_tensor_core_raw_scores
constexpr float TC_ERR1 = 2.5e-3f;   // one tf32 product: 2^-9 operand truncation (+ accumulation)
constexpr float TC_ERR3 = 1.2e-5f;   // split operands: 3 * 2^-20 operand terms + fp32 accumulation over 3*KP products

template <int KP, int TERMS>
struct TcSmem {
  static constexpr int NKH = KP / 32;                          // 128-byte K sub-tiles
  static constexpr int NA = TERMS == 3 ? 2 : 1;                // U tile, U_lo tile
  static constexpr int SPT = TERMS == 3 ? 2 : 1;               // ring slots per item tile (V, V_lo)
  static constexpr uint32_t A_BYTES = TC_BM * KP * 4;          // one U tile
  static constexpr uint32_t B_BYTES = TC_BN * KP * 4;          // one ring slot
  static constexpr uint32_t off_a = 0;
  static constexpr uint32_t off_b = off_a + NA * A_BYTES;
  static constexpr uint32_t off_cs = off_b + TC_SLOTS * B_BYTES;    // float [TC_CAND][128]
  static constexpr uint32_t off_ci = off_cs + TC_CAND * TC_BM * 4;  // int   [TC_CAND][128]
  static constexpr uint32_t off_top = off_ci + TC_CAND * TC_BM * 4; // float [TC_NTOP][128]
  static constexpr uint32_t off_state = off_top + TC_NTOP * TC_BM * 4;  // per-row cursors / counters, [field][128]
  static constexpr uint32_t off_bar = off_state + 10 * TC_BM * 4;
  static constexpr uint32_t total = off_bar + 256 + 1024;           // + alignment slack
  static_assert(total <= 227 * 1024, "tensor-core scoring kernel exceeds shared memory");
};

struct TcArgs {
  const float* uf;
  const float* vf;
  int64_t n_users;
  int n_items;
  const int64_t* seen_indptr;
  const int32_t* seen_indices;
  int n;
  int32_t* idx_out;
  double* score_out;
  const float* vmax;        // max_i |v_i|_2
  float err;                // relative error bound of the approximate scores (TC_ERR1 / TC_ERR3)
  int32_t* overflow_rows;   // rows that need the exact kernel
  unsigned int* overflow_count;
  float* dump;              // debug: raw tensor-core scores, (n_users_padded x dump_ld)
  int64_t dump_ld;
  int n_user_tiles, n_item_tiles;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // bounded: a protocol bug must end in a trap, not in a hung GPU
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 26)) {
      printf("score_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, rows packed at 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF) | (static_cast<uint64_t>(1) << 16) |
         (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
}

// instruction descriptor: D fp32, A/B tf32, both K-major, N = TC_BN, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(TC_BN >> 3) << 17) |
                              (static_cast<uint32_t>(TC_BM >> 4) << 24);

// ---- per-row selection state: thread-private slices of shared memory ([field][row], conflict free) ------------
struct RowMem {
  float* cs;        // [TC_CAND][128] candidate approximate scores
  int* ci;          // [TC_CAND][128] candidate item ids
  float* top;       // [TC_NTOP][128] n best approximate scores of unseen items, descending
  int* cnt;         // [128] candidates stored
  int* nxt;         // [128] next seen item id >= current column (INT_MAX: none)
  int* nxt2;        // [128] the one after (prefetched)
  int* cur_lo;      // [128] low / high words of the cursor into seen_indices
  int* cur_hi;
  int* end_lo;
  int* end_hi;
  float* delta;     // [128] 2 * eps_u
  int* flags;       // [128] bit 0: candidate buffer overflowed
};

// one score above the row threshold: returns the (possibly raised) threshold
__device__ __noinline__ float row_hit(float s, int col, int row, float tau, RowMem m, const int32_t* __restrict__ seen_indices,
                                      int n_items, int n) {
  int nxt = m.nxt[row];
  if (nxt < col) {  // move the seen cursor up to this column
    int64_t cur = (static_cast<int64_t>(m.cur_hi[row]) << 32) | static_cast<uint32_t>(m.cur_lo[row]);
    const int64_t end = (static_cast<int64_t>(m.end_hi[row]) << 32) | static_cast<uint32_t>(m.end_lo[row]);
    int nxt2 = m.nxt2[row];
    while (nxt < col) {
      nxt = nxt2;
      ++cur;
      nxt2 = (cur + 1 < end) ? seen_indices[cur + 1] : INT_MAX;
    }
    m.nxt[row] = nxt;
    m.nxt2[row] = nxt2;
    m.cur_lo[row] = static_cast<int>(cur & 0xffffffff);
    m.cur_hi[row] = static_cast<int>(cur >> 32);
  }
  if (col == nxt || col >= n_items) return tau;
  int cnt = m.cnt[row];
  if (cnt == TC_CAND) {  // drop everything the current threshold has made obsolete
    int j = 0;
    for (int i = 0; i < TC_CAND; ++i) {
      const float v = m.cs[i * TC_BM + row];
      if (v > tau) {
        m.cs[j * TC_BM + row] = v;
        m.ci[j * TC_BM + row] = m.ci[i * TC_BM + row];
        ++j;
      }
    }
    cnt = j;
  }
  if (cnt < TC_CAND) {
    m.cs[cnt * TC_BM + row] = s;
    m.ci[cnt * TC_BM + row] = col;
    ++cnt;
  } else {
    m.flags[row] |= 1;
  }
  m.cnt[row] = cnt;
  if (s > m.top[(n - 1) * TC_BM + row]) {  // sorted insertion into the n best approximate scores
    int i = n - 1;
    while (i > 0 && m.top[(i - 1) * TC_BM + row] < s) {
      m.top[i * TC_BM + row] = m.top[(i - 1) * TC_BM + row];
      --i;
    }
    m.top[i * TC_BM + row] = s;
    tau = m.top[(n - 1) * TC_BM + row] - m.delta[row];
  }
  return tau;
}

template <int KP, int TERMS, bool DUMP>
__global__ void __launch_bounds__(TC_THREADS, 1)
score_topn_tc_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_ulo,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_vlo, const TcArgs a) {
  using S = TcSmem<KP, TERMS>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // 128B-swizzled tiles need 1024-byte alignment
  unsigned char* sm = smem_raw + (base - raw);
  RowMem rm;
  rm.cs = reinterpret_cast<float*>(sm + S::off_cs);
  rm.ci = reinterpret_cast<int*>(sm + S::off_ci);
  rm.top = reinterpret_cast<float*>(sm + S::off_top);
  {
    int* st = reinterpret_cast<int*>(sm + S::off_state);
    rm.cnt = st; rm.nxt = st + TC_BM; rm.nxt2 = st + 2 * TC_BM; rm.cur_lo = st + 3 * TC_BM; rm.cur_hi = st + 4 * TC_BM;
    rm.end_lo = st + 5 * TC_BM; rm.end_hi = st + 6 * TC_BM; rm.delta = reinterpret_cast<float*>(st + 7 * TC_BM);
    rm.flags = st + 8 * TC_BM;
  }
  int* ci = rm.ci;
  float* top = rm.top;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::off_bar);
  // barrier slots: [0,3) slot full  [3,6) slot empty  6 a_full  7 a_empty  [8,12) tmem_full  [12,16) tmem_empty ; 16: tmem base
  const uint32_t bar0 = base + S::off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int B_FULL = 0, B_EMPTY = TC_SLOTS, A_FULL = 2 * TC_SLOTS, A_EMPTY = A_FULL + 1, T_FULL = A_FULL + 2, T_EMPTY = T_FULL + TC_ACC;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + T_EMPTY + TC_ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_SLOTS; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 1); }
    for (int i = 0; i < TC_ACC; ++i) { mbar_init(BAR(T_FULL + i), 1); mbar_init(BAR(T_EMPTY + i), 4); }
    mbar_init(BAR(A_FULL), 1);
    mbar_init(BAR(A_EMPTY), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================================== TMA producer ==============================================
    if (lane == 0) {
      uint32_t slot_global = 0, ut_local = 0;
      for (int ut = blockIdx.x; ut < a.n_user_tiles; ut += gridDim.x, ++ut_local) {
        mbar_wait(BAR(A_EMPTY), (ut_local & 1) ^ 1);  // U tile(s) free (first pass succeeds immediately)
        mbar_expect_tx(BAR(A_FULL), S::NA * S::A_BYTES);
#pragma unroll
        for (int kh = 0; kh < S::NKH; ++kh) {
          tma_load_2d(base + S::off_a + kh * (TC_BM * 128), &map_u, BAR(A_FULL), kh * 32, ut * TC_BM);
          if (TERMS == 3) tma_load_2d(base + S::off_a + S::A_BYTES + kh * (TC_BM * 128), &map_ulo, BAR(A_FULL), kh * 32, ut * TC_BM);
        }
        for (int it = 0; it < a.n_item_tiles; ++it) {
#pragma unroll
          for (int part = 0; part < S::SPT; ++part, ++slot_global) {
            const uint32_t slot = slot_global % TC_SLOTS, ph = (slot_global / TC_SLOTS) & 1;
            mbar_wait(BAR(B_EMPTY + slot), ph ^ 1);
            mbar_expect_tx(BAR(B_FULL + slot), S::B_BYTES);
#pragma unroll
            for (int kh = 0; kh < S::NKH; ++kh)
              tma_load_2d(base + S::off_b + slot * S::B_BYTES + kh * (TC_BN * 128), part == 0 ? &map_v : &map_vlo,
                          BAR(B_FULL + slot), kh * 32, it * TC_BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================== MMA issuer ================================================
    if (lane == 0) {
      uint32_t slot_global = 0, it_global = 0, ut_local = 0;
      for (int ut = blockIdx.x; ut < a.n_user_tiles; ut += gridDim.x, ++ut_local) {
        mbar_wait(BAR(A_FULL), ut_local & 1);
        for (int it = 0; it < a.n_item_tiles; ++it, ++it_global) {
          const uint32_t acc = it_global % TC_ACC, aph = (it_global / TC_ACC) & 1;
          const uint32_t d_tmem = tmem_base + acc * TC_BN;
          mbar_wait(BAR(T_EMPTY + acc), aph ^ 1);    // accumulator stage drained by the epilogue
#pragma unroll
          for (int part = 0; part < S::SPT; ++part, ++slot_global) {
            const uint32_t slot = slot_global % TC_SLOTS, ph = (slot_global / TC_SLOTS) & 1;
            mbar_wait(BAR(B_FULL + slot), ph);       // V (part 0) or V_lo (part 1) tile landed
            tc_fence_after();
            const uint32_t b_base = base + S::off_b + slot * S::B_BYTES;
            // part 0: U.V (+ U_lo.V when TERMS == 3)      part 1: U.V_lo
            const int n_a = (part == 0) ? S::NA : 1;
            for (int ta = 0; ta < n_a; ++ta) {
              const uint32_t a_base = base + S::off_a + ta * S::A_BYTES;
#pragma unroll
              for (int kk = 0; kk < KP / 8; ++kk) {
                const uint32_t kin = (kk & 3) * 32u;
                const uint64_t ad = make_desc(a_base + (kk >> 2) * (TC_BM * 128) + kin);
                const uint64_t bd = make_desc(b_base + (kk >> 2) * (TC_BN * 128) + kin);
                tc_mma_tf32(d_tmem, ad, bd, TC_IDESC, (part | ta | kk) != 0 ? 1u : 0u);
              }
            }
            tc_commit(BAR(B_EMPTY + slot));          // slot reusable once these MMAs retire
          }
          tc_commit(BAR(T_FULL + acc));              // accumulator ready for the epilogue
        }
        tc_commit(BAR(A_EMPTY));                     // U tile(s) reusable
      }
    }
  } else if (warp >= 4) {
    // =========================================== epilogue ==================================================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;          // row inside the user tile == TMEM lane
    const float vmax = DUMP ? 0.f : *a.vmax;
    uint32_t it_global = 0;
    for (int ut = blockIdx.x; ut < a.n_user_tiles; ut += gridDim.x) {
      const int64_t user = static_cast<int64_t>(ut) * TC_BM + row;
      const bool valid = user < a.n_users;
      float tau = -CUDART_INF_F;
      if (!DUMP) {
#pragma unroll
        for (int i = 0; i < TC_NTOP; ++i) top[i * TC_BM + row] = -CUDART_INF_F;
        int64_t cur = 0, end = 0;
        int nxt = INT_MAX, nxt2 = INT_MAX;
        float delta = 0.f;
        if (valid) {
          float nn = 0.f;
          const float4* up = reinterpret_cast<const float4*>(a.uf + user * KP);
#pragma unroll
          for (int c = 0; c < KP / 4; ++c) {
            const float4 x = up[c];
            nn += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
          }
          delta = 2.0f * a.err * sqrtf(nn) * vmax + 1e-30f;
          if (a.seen_indptr) {
            cur = a.seen_indptr[user];
            end = a.seen_indptr[user + 1];
            if (cur < end) nxt = a.seen_indices[cur];
            if (cur + 1 < end) nxt2 = a.seen_indices[cur + 1];
          }
        }
        rm.cnt[row] = 0; rm.nxt[row] = nxt; rm.nxt2[row] = nxt2; rm.delta[row] = delta; rm.flags[row] = 0;
        rm.cur_lo[row] = static_cast<int>(cur & 0xffffffff); rm.cur_hi[row] = static_cast<int>(cur >> 32);
        rm.end_lo[row] = static_cast<int>(end & 0xffffffff); rm.end_hi[row] = static_cast<int>(end >> 32);
      }
      for (int it = 0; it < a.n_item_tiles; ++it, ++it_global) {
        const uint32_t acc = it_global % TC_ACC, aph = (it_global / TC_ACC) & 1;
        mbar_wait(BAR(T_FULL + acc), aph);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * TC_BN;
#pragma unroll 1
        for (int g = 0; g < TC_BN / 32; ++g) {
          float s[32];
          tc_ld32(t_row + g * 32, s);
          const int col0 = it * TC_BN + g * 32;
          if (DUMP) {
            if (valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) a.dump[user * a.dump_ld + col0 + j] = s[j];
            }
          } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const float m8 = max3(max3(s[8 * b], s[8 * b + 1], s[8 * b + 2]), max3(s[8 * b + 3], s[8 * b + 4], s[8 * b + 5]),
                                    fmaxf(s[8 * b + 6], s[8 * b + 7]));
              if (m8 > tau) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (s[8 * b + j] > tau) tau = row_hit(s[8 * b + j], col0 + 8 * b + j, row, tau, rm, a.seen_indices, a.n_items, a.n);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(T_EMPTY + acc));
      }
      if (DUMP) continue;
      // ---- exact float64 re-scoring of the candidates, ranking by (score desc, index asc) ------------------
      if (valid) {
        const int cnt = rm.cnt[row];
        if (rm.flags[row] & 1) {
          const unsigned slot = atomicAdd(a.overflow_count, 1u);
          a.overflow_rows[slot] = static_cast<int32_t>(user);
        } else {
          double ex[TC_CAND];
          const float* up = a.uf + user * KP;
          for (int c = 0; c < cnt; ++c) {
            const float* vp = a.vf + static_cast<int64_t>(ci[c * TC_BM + row]) * KP;
            double accd = 0.0;
#pragma unroll 4
            for (int k4 = 0; k4 < KP / 4; ++k4) {
              const float4 x = *reinterpret_cast<const float4*>(up + 4 * k4);
              const float4 y = *reinterpret_cast<const float4*>(vp + 4 * k4);
              accd = fma(static_cast<double>(x.x), static_cast<double>(y.x), accd);
              accd = fma(static_cast<double>(x.y), static_cast<double>(y.y), accd);
              accd = fma(static_cast<double>(x.z), static_cast<double>(y.z), accd);
              accd = fma(static_cast<double>(x.w), static_cast<double>(y.w), accd);
            }
            ex[c] = accd;
          }
          for (int p = 0; p < a.n; ++p) {
            int best = -1;
            for (int c = 0; c < cnt; ++c) {
              const int id = ci[c * TC_BM + row];
              if (id < 0) continue;
              if (best < 0 || beats(ex[c], id, ex[best], ci[best * TC_BM + row])) best = c;
            }
            if (best >= 0) {
              a.idx_out[user * a.n + p] = ci[best * TC_BM + row];
              if (a.score_out) a.score_out[user * a.n + p] = ex[best];
              ci[best * TC_BM + row] = -1;
            } else {
              a.idx_out[user * a.n + p] = -1;
              if (a.score_out) a.score_out[user * a.n + p] = -CUDART_INF;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// x_lo = x - (x with the 13 low mantissa bits cleared): the part of x the tf32 datapath does not see
__global__ void __launch_bounds__(256) split_lo_kernel(const float* __restrict__ x, int64_t n4, float* __restrict__ lo) {
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n4; p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[p];
    float4 r;
    r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    reinterpret_cast<float4*>(lo)[p] = r;
  }
}

__global__ void __launch_bounds__(256) max_row_norm_kernel(const float* __restrict__ v, int64_t n, int kp, float* __restrict__ out) {
  float best = 0.f;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float nn = 0.f;
    const float4* p = reinterpret_cast<const float4*>(v + r * kp);
    for (int c = 0; c < kp / 4; ++c) {
      const float4 x = p[c];
      nn += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    best = fmaxf(best, nn);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(FULL_MASK, best, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(sqrtf(best) * 1.0000002f));  // non-negative floats order as ints
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map(rl_context* h, CUtensorMap* map, const float* ptr, int64_t rows, int kp, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(h, RL_ERR_CUDA, "score_tc: cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kp), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kp) * 4};
  const cuuint32_t box[2] = {32, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, RL_ERR_CUDA, "score_tc: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return RL_OK;
}

template <int KP, int TERMS, bool DUMP>
int launch_tc(rl_context* h, const CUtensorMap& mu, const CUtensorMap& mulo, const CUtensorMap& mv, const CUtensorMap& mvlo,
              const TcArgs& a) {
  auto kern = score_topn_tc_kernel<KP, TERMS, DUMP>;
  const size_t smem = TcSmem<KP, TERMS>::total;
  RL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int grid = a.n_user_tiles < h->sm_count ? a.n_user_tiles : h->sm_count;
  kern<<<grid, TC_THREADS, smem, h->stream>>>(mu, mulo, mv, mvlo, a);
  RL_LAUNCH_CHECK(h, "score_topn_tc_kernel");
  return RL_OK;
}

template <int KP>
int launch_tc_k(rl_context* h, int terms, bool dump, const CUtensorMap& mu, const CUtensorMap& mulo, const CUtensorMap& mv,
                const CUtensorMap& mvlo, const TcArgs& a) {
  if (terms == 3) return dump ? launch_tc<KP, 3, true>(h, mu, mulo, mv, mvlo, a) : launch_tc<KP, 3, false>(h, mu, mulo, mv, mvlo, a);
  return dump ? launch_tc<KP, 1, true>(h, mu, mulo, mv, mvlo, a) : launch_tc<KP, 1, false>(h, mu, mulo, mv, mvlo, a);
}

}  // namespace

bool score_tc_supported(int kp, int n, int64_t n_users, int64_t n_items, const float* ubias, const float* ibias, double gbias) {
  return (kp == 32 || kp == 64) && n <= TC_NTOP && n_items >= 1024 && n_users >= 1 && !ubias && !ibias && gbias == 0.0;
}

// tensor-core candidate pass + in-kernel exact re-scoring; overflow rows are finished by the exact kernel.
// terms: 1 = single tf32 product, 3 = split operands (default).
int launch_score_tc(rl_context* h, int kp, const ScoreArgs& sa, float* dump, int64_t dump_ld, int terms) {
  if (terms != 1 && terms != 3) return fail(h, RL_ERR_INVALID, "score_tc: terms must be 1 or 3");
  // scratch: vmax (4 B) | overflow count (4 B) | overflow rows (n_users x 4 B) | U_lo | V_lo
  const size_t rows_bytes = (sizeof(int32_t) * static_cast<size_t>(sa.n_users) + 255) & ~size_t(255);
  const size_t ulo_bytes = terms == 3 ? sizeof(float) * static_cast<size_t>(sa.n_users) * kp : 0;
  const size_t vlo_bytes = terms == 3 ? sizeof(float) * static_cast<size_t>(sa.n_items) * kp : 0;
  void* sc = nullptr;
  int rc = scratch(h, 256 + rows_bytes + ulo_bytes + vlo_bytes, &sc);
  if (rc != RL_OK) return rc;
  char* scb = static_cast<char*>(sc);
  float* vmax = reinterpret_cast<float*>(scb);
  unsigned int* ocount = reinterpret_cast<unsigned int*>(scb + 4);
  int32_t* orows = reinterpret_cast<int32_t*>(scb + 256);
  float* ulo = reinterpret_cast<float*>(scb + 256 + rows_bytes);
  float* vlo = reinterpret_cast<float*>(scb + 256 + rows_bytes + ulo_bytes);
  RL_CUDA(h, cudaMemsetAsync(sc, 0, 256, h->stream));
  if (terms == 3) {
    const int64_t nu4 = sa.n_users * kp / 4, nv4 = sa.n_items * kp / 4;
    const int64_t cap = 8 * static_cast<int64_t>(h->sm_count);
    int64_t g = (nu4 + 255) / 256;
    split_lo_kernel<<<static_cast<unsigned>(g < cap ? g : cap), 256, 0, h->stream>>>(sa.uf, nu4, ulo);
    RL_LAUNCH_CHECK(h, "split_lo_kernel");
    g = (nv4 + 255) / 256;
    split_lo_kernel<<<static_cast<unsigned>(g < cap ? g : cap), 256, 0, h->stream>>>(sa.vf, nv4, vlo);
    RL_LAUNCH_CHECK(h, "split_lo_kernel");
  }
  CUtensorMap mu, mulo, mv, mvlo;
  if ((rc = make_map(h, &mu, sa.uf, sa.n_users, kp, TC_BM)) != RL_OK) return rc;
  if ((rc = make_map(h, &mv, sa.vf, sa.n_items, kp, TC_BN)) != RL_OK) return rc;
  if ((rc = make_map(h, &mulo, terms == 3 ? ulo : sa.uf, sa.n_users, kp, TC_BM)) != RL_OK) return rc;
  if ((rc = make_map(h, &mvlo, terms == 3 ? vlo : sa.vf, sa.n_items, kp, TC_BN)) != RL_OK) return rc;
  TcArgs a{};
  a.uf = sa.uf; a.vf = sa.vf; a.n_users = sa.n_users; a.n_items = static_cast<int>(sa.n_items);
  a.seen_indptr = sa.seen_indptr; a.seen_indices = sa.seen_indices; a.n = sa.n; a.idx_out = sa.idx_out; a.score_out = sa.score_out;
  a.vmax = vmax; a.err = terms == 3 ? TC_ERR3 : TC_ERR1;
  a.overflow_rows = orows; a.overflow_count = ocount; a.dump = dump; a.dump_ld = dump_ld;
  a.n_user_tiles = static_cast<int>((sa.n_users + TC_BM - 1) / TC_BM);
  a.n_item_tiles = static_cast<int>((sa.n_items + TC_BN - 1) / TC_BN);
  if (!dump) {
    int64_t g = (sa.n_items + 255) / 256;
    if (g > 4 * h->sm_count) g = 4 * h->sm_count;
    max_row_norm_kernel<<<static_cast<unsigned>(g), 256, 0, h->stream>>>(sa.vf, sa.n_items, kp, vmax);
    RL_LAUNCH_CHECK(h, "max_row_norm_kernel");
  }
  if (kp == 64) rc = launch_tc_k<64>(h, terms, dump != nullptr, mu, mulo, mv, mvlo, a);
  else if (kp == 32) rc = launch_tc_k<32>(h, terms, dump != nullptr, mu, mulo, mv, mvlo, a);
  else return fail(h, RL_ERR_INVALID, "score_tc: padded k %d not supported by the tensor-core path", kp);
  if (rc != RL_OK || dump) return rc;
  // rows whose candidate buffer overflowed: exact kernel on the list (needs the count on the host)
  unsigned int n_over = 0;
  RL_CUDA(h, cudaMemcpyAsync(&n_over, ocount, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  h->last_overflow = n_over;
  if (n_over > 0) {
    ScoreArgs ea = sa;
    ea.user_list = orows;
    ea.n_list = n_over;
    ea.scatter = true;
    return launch_score_exact(h, kp, ea);
  }
  return RL_OK;
}

}  // namespace rl
