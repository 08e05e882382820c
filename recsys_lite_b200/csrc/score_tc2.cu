// score_tc2.cu -- single-product tensor-core scoring on CENTRED, norm-classed item factors, fused with seen-item
// masking and exact top-n (tcgen05 / TMEM / TMA).  Default path of rl_score_topn_dev for k <= 64.
//
// Replaces predict + top_n of the reference (algo.py:143-147 `U @ V.T`, algo.py:533-659 as composed in
// recommend, algo.py:172-201) for ALS factors without ever writing the n_users x n_items score matrix.
//
// Why centring.  The reference's ALS has no negative term, so its factors converge towards "every score = 1": after
// 25 iterations at C4 the scores of one user are 0.9977 +- 9e-3 and the top ten sit 3e-5 apart, while |u||v| ~ 1.
// An error bound relative to |u||v| (what the split-operand kernel of score_tc.cu uses) then lets hundreds of items
// into the candidate window.  Ranking a user's items by u.v_i is ranking them by u.(v_i - vbar): the constant
// u.vbar drops out.  With w_i = v_i - vbar the operand of the GEMM shrinks with the spread of the scores, and so
// does the rounding error of ONE fp16 product:
//     |fp16(u).fp16(w_i) - u.w_i|  <=  |u - u_h| |w_i| + |u_h| |w_i - w_h,i|  (+ fp32 accumulation)
// -- computed from what the roundings actually dropped, per user row and per item class.  Items are binned by
// |w_i| into factor-of-two classes (scan order: see class_scan_kernel; each class padded to whole 128-item tiles, ascending item id
// inside a class), so that the 8 % of slow-converging items with |w_i| ~ 1 do not set the window of the 92 % with
// |w_i| ~ 0.015.  One product instead of three (score_tc.cu) is a third of the tensor work, and the kernel is
// power-bound on exactly that work.
//
// Exactness.  All quantities in "raw" units (scores times s_u s, both powers of two).  For item i of class c:
// a_i = tensor-core value, T_i = s_u s u.(v_i - vbar) the true centred score, |a_i - T_i| <= eps_c (eps_c covers the
// two operand roundings, the rounding of the fp32 subtraction v_i - vbar, fp32 accumulation, and the float64
// summation error of the final re-scoring).  Each row tracks the n best LOWER bounds L_i = a_i - eps_c of unseen
// items; Lth = the n-th best is a lower bound of the n-th best true score.  An item is a candidate iff its UPPER
// bound H_i = a_i + eps_c reaches Lth: every member of the exact top-n is a candidate.  Candidates are re-scored in
// float64 from the original u, v_i (same summation order as the exact kernel) and ranked by (score desc, index
// asc): the result is identical to rl_score_topn_dev(mode = RL_SCORE_EXACT).
//
// One persistent CTA per SM owns TWO user tiles of 128 at a time ("halves": every V tile fetched from L2 feeds two
// MMA groups) and sweeps the items in tiles of 128:
//   warp 0       TMA producer: W_hi tiles (128 x 64 fp16, 128B swizzle, K-major) through a shared-memory ring;
//   warp 1       MMA issuer: tcgen05.mma.kind::f16 (M=128, N=128, K=16), A (a user tile) from TMEM, B from shared
//                memory, fp32 accumulators in three 128-column TMEM stages shared by the two halves; also issues the
//                512-byte bulk copy of the tile's item ids (position -> item id) that completes on the same mbarrier;
//   warp 2       TMEM allocation;
//   warps 4-11   scan warps, one per (half, TMEM lane quarter): tcgen05.ld 32 columns at a time, software pipelined,
//                3-input max over groups of 8 scores, ONE compare per group against the row threshold; scores above it
//                are pushed as (score, item id) into the row's ring in shared memory;
//   warps 12-19  selection warps, one thread per user row: drain the row's ring in order -- seen-item cursor
//                (ascending CSR row, restarted at class boundaries), top-n tracker of lower bounds in registers,
//                candidate list in shared memory, threshold published for the scan thread; at the end of a sweep
//                they write the candidate ids to global memory and put the next-but-one user tile into TMEM
//                (tcgen05.st of the fp16-rounded, power-of-two scaled row -- no shared-memory copy of U).
// A second small kernel re-scores the candidates exactly and writes the top-n.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdlib>

#include "common.cuh"
#include "score_exact.cuh"
#include "tcgen05_ptx.cuh"

namespace rl {

namespace {

using namespace tc;

constexpr int T2_BM = 128;        // users per half (TMEM lanes)
constexpr int T2_BN = 128;        // items per tile
constexpr int T2_ACC = 3;         // TMEM accumulator stages
constexpr int T2_NCLS = 8;        // item classes by |w_i| (factor of two each)
#ifndef T2_SUBS
#define T2_SUBS 2               // scan warps per (half, TMEM lane quarter): 1 = one warp reads all 128 columns of an item tile
#endif                          // (32 columns per tcgen05.ld), 2 = two warps, 64 columns each (16 per tcgen05.ld, 72 registers)
constexpr int T2_THREADS = 32 * (4 + 8 * T2_SUBS + 8);   // 4 control + scan + 8 selection warps
constexpr int T2_RING = 16 / T2_SUBS;   // ring slots per row and producer
#ifndef T2_CHUNK
#define T2_CHUNK 32              // columns per tcgen05.ld of a scan warp (one chunk of the compare loop)
#endif
constexpr int T2_CH = T2_CHUNK;         // columns per tcgen05.ld (16 or 32)
constexpr int T2_CAND = 40;       // candidate slots per row (shared memory)
#ifndef T2_SPILL_BLOCKS
#define T2_SPILL_BLOCKS 6
#endif
constexpr int T2_SPILL_MAX = T2_SPILL_BLOCKS;   // ... plus up to this many pool blocks of T2_CAND positions in global memory when the list runs over
#ifndef T2_THR_PER_CHUNK
#define T2_THR_PER_CHUNK 1      // re-read the row threshold for every 16 columns (1) or once per item tile (0)
#endif
#ifndef T2_PUSH_COUNTED
#define T2_PUSH_COUNTED 1       // slow path: count the hits of a group, wait once for that much room, straight predicated stores
#endif
#ifndef T2_JOINT
#define T2_JOINT 1              // candidate compaction / seen-window refills run jointly for the 32 rows of a selection warp
#endif
#ifndef T2_POLL_NS
#define T2_POLL_NS 1500         // sleep of an idle selection warp between two looks at its rings
#endif
#ifndef T2_PERIOD_MIN
#define T2_PERIOD_MIN 3000      // interval at the start of a sweep (thresholds move fast); grows by 150 clocks per item tile
#endif
#ifndef T2_PERIOD_MAX
#define T2_PERIOD_MAX 16000     // largest interval (clocks) between two selection passes of a warp
#endif
constexpr int T2_WIN = 32;        // seen positions of a row kept in shared memory at a time
constexpr int T2_EMPTY = INT_MIN; // ring slot free
constexpr int T2_END = -100;      // ring marker: end of the sweep (records carry item POSITIONS >= 0)
constexpr int T2_PERIOD = 3000;   // clocks between selection passes when no ring is filling up

// what the preprocessing kernels leave for the main kernel
struct ClassTable {
  int first_tile[T2_NCLS + 1];    // first item tile of class c; first_tile[T2_NCLS] = n_tiles
  int n_tiles;
  int pad0;
  float wc[T2_NCLS];              // >= max over class c of |w_i| s                    (raw units, inflated)
  float rc[T2_NCLS];              // >= max over class c of |w_i s - fp16(w_i s)|
  float scale;                    // s: power of two, max |w_ij| s <= 2^14
  float vmax_raw;                 // >= max_i |v_i| s
  unsigned wmax_bits, amax_bits, vmax_bits;   // reductions in progress (non-negative floats order as unsigned)
  unsigned wc_bits[T2_NCLS], rc_bits[T2_NCLS];
  int count[T2_NCLS];             // items of class c (its last tile is padded with zero rows)
  int label_of[T2_NCLS];          // norm bin (class_of) -> class label = place in the scan order
};

template <int KP>
struct T2Smem {
  static constexpr int SLOTS = KP > 64 ? 2 : 3;
  static constexpr uint32_t B_BYTES = T2_BN * KP * 2;
  static constexpr uint32_t off_b = 0;
  static constexpr uint32_t off_ring = off_b + SLOTS * B_BYTES;                    // [2 producers][RING][256] (score, position)
  static constexpr uint32_t off_cand = off_ring + T2_SUBS * T2_RING * 2 * T2_BM * 8;     // [CAND][256] (upper bound, position)
  static constexpr uint32_t off_thr = off_cand + T2_CAND * 2 * T2_BM * 8;          // [256] (threshold, sweep tag)
  static constexpr uint32_t off_win = off_thr + 2 * T2_BM * 8;                     // [T2_WIN][256] next seen positions of a row
  static constexpr uint32_t off_eps = off_win + T2_WIN * 2 * T2_BM * 4;            // [2 sweep parities][NCLS][256] windows
  static constexpr uint32_t off_tab = off_eps + 2 * T2_NCLS * 2 * T2_BM * 4;       // ClassTable copy
  static constexpr uint32_t off_bar = off_tab + 256;
  static constexpr uint32_t total = off_bar + 256 + 1024;
  static_assert(sizeof(ClassTable) <= 256, "class table");
  static_assert(total <= 227 * 1024, "score_tc2 exceeds shared memory");
};

struct T2Args {
  const float* uf;
  int ldf;
  int64_t n_users;
  const int64_t* seen_indptr;
  int n;
  const ClassTable* tab;
  const int32_t* seen_pos;      // the seen items of every row as POSITIONS, ascending per row (same indptr as seen_indices)
  int64_t seen_cap;             // entries seen_pos has room for (rows beyond it go to the exact kernel)
  int32_t* cand_ids;            // [n_users][T2_CAND] candidate positions; spilled rows: [n blocks | block ids | block counts]
  int32_t* cand_cnt;            // [n_users]: number of candidates, -1 = overflow (row goes to the exact kernel), -2 = spilled
  int32_t* pool;                // [pool_blocks][T2_CAND] spilled candidate positions
  unsigned int* pool_next;      // next free pool block
  unsigned int pool_blocks;
  float acc_rel;                // allowance for the fp32 accumulation of the tensor core, relative to |u_h||w_h|
  float* dump;                  // debug: raw accumulators (n_users x dump_ld), dump_su: the row scales
  float* dump_su;
  int64_t dump_ld;
  int n_user_tiles;
  long long* dbg;
  int dbg_flags;                // RL_TC2_FLAGS: 1 = the scan pushes nothing, 2 = selection skips the seen cursor, 4 = stages released unread, 8 = loads only
};

__device__ __forceinline__ void ring_put(uint32_t addr, float score, int id) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(__float_as_int(score)), "r"(id) : "memory");
}
__device__ __forceinline__ void ring_get(uint32_t addr, float& score, int& id) {
  int sb;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(sb), "=r"(id) : "r"(addr) : "memory");
  score = __int_as_float(sb);
}
__device__ __forceinline__ int lds_volatile(uint32_t addr) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// The kernel is instruction-fetch sensitive (28 warps in six different roles per SM): the cold paths of the selection and
// scan loops are real function calls.
// seen-position window of a row <- seen[woff .. woff + T2_WIN)
__device__ __noinline__ void t2_refill_window(const int32_t* __restrict__ seen, int seen_len, int woff, uint32_t win_row, uint32_t es) {
#pragma unroll 8
  for (int i = 0; i < T2_WIN; ++i) {
    const int v = (woff + i < seen_len) ? seen[woff + i] : INT_MAX;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(win_row + i * es), "r"(v) : "memory");
  }
}
// drops the candidates whose upper bound fell below the row's threshold; returns the new length of the list
__device__ __noinline__ int t2_compact_list(uint32_t cand_row, uint32_t rs, int cnt, float lth) {
  int jw = 0;
  for (int f = 0; f < cnt; ++f) {
    float hb; int id;
    ring_get(cand_row + f * rs, hb, id);
    if (hb >= lth) { ring_put(cand_row + jw * rs, hb, id); ++jw; }
  }
  return jw;
}
// a scan thread waits for the selection thread to free ring slots (in order)
__device__ __noinline__ void t2_wait_slot(uint32_t addr) {
  uint32_t spin = 0;
  while (lds_volatile(addr) != T2_EMPTY) {
    __nanosleep(64);
    if (++spin > (1u << 24)) { printf("score_tc2: ring timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}

// rare path of the selection threads, kept out of line: the row's candidate list (cnt positions in shared memory, entry f at
// cand_row + f * rs) moves to a fresh pool block; the row's header in cand_ids takes the block id and the count
__device__ __noinline__ bool t2_spill_block(int32_t* pool, unsigned int* pool_next, unsigned int pool_blocks, int32_t* hdr, uint32_t cand_row,
                                            uint32_t rs, int cnt, int n_blk) {
  const unsigned int b = atomicAdd(pool_next, 1u);
  if (b >= pool_blocks) return false;
  int32_t* dst = pool + static_cast<size_t>(b) * T2_CAND;
  for (int f = 0; f < cnt; ++f) {
    float hb; int p;
    ring_get(cand_row + f * rs, hb, p);
    dst[f] = p;
  }
  hdr[1 + n_blk] = static_cast<int32_t>(b);
  hdr[1 + T2_SPILL_MAX + n_blk] = cnt;
  return true;
}

// instruction descriptor: D fp32, A/B fp16, both K-major, N = 128, M = 128
constexpr uint32_t T2_IDESC = (1u << 4) | (static_cast<uint32_t>(T2_BN >> 3) << 17) | (static_cast<uint32_t>(T2_BM >> 4) << 24);

template <int KP, int NT, bool DUMP, bool DBG>
__global__ void __launch_bounds__(T2_THREADS, 1)
score_tc2_kernel(const __grid_constant__ CUtensorMap map_w, const T2Args a) {
  using S = T2Smem<KP>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - raw);
  constexpr int SLOTS = S::SLOTS;
  // barriers: B slot full / empty, user tile full / empty (two TMEM buffers), then PER HALF: accumulator stage full / empty.
  // A parity wait is only sound when its waiter sees every phase of the barrier in order; with one MMA issuer per half the
  // accumulator barriers therefore belong to one half each (T_FULL[h][s]: committed by issuer h, awaited by the scan warps of
  // half h; T_EMPTY[h][s]: released by those warps, awaited by whichever issuer uses stage s next -- the schedule is static,
  // so every issuer knows who used a stage last and how often).
  constexpr int B_FULL = 0, B_EMPTY = SLOTS, A_FULL = 2 * SLOTS, A_EMPTY = A_FULL + 2, T_FULL = A_FULL + 4, T_EMPTY = T_FULL + 2 * T2_ACC;
  const uint32_t bar0 = base + S::off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  auto wait_bar = [&](uint32_t bar, uint32_t parity, long long& acc) {     // DBG: the time spent waiting is counted
    if constexpr (DBG) {
      const long long t0 = clock64();
      mbar_wait_sleep(bar, parity);
      acc += clock64() - t0;
    } else {
      mbar_wait_sleep(bar, parity);
    }
  };
  auto dflag = [&](int bit) { return DBG && (a.dbg_flags & bit) != 0; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::off_bar);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + T_EMPTY + 2 * T2_ACC);
  ClassTable* tab = reinterpret_cast<ClassTable*>(sm + S::off_tab);
  constexpr int A_BUFS = KP <= 64 ? 2 : 1;
  constexpr uint32_t A_COLS = KP / 2;           // TMEM columns of one user tile (two fp16 per column)
  constexpr uint32_t ACC_COL0 = 128;
  constexpr uint32_t RS = 2 * T2_BM * 8;        // byte stride between ring / candidate slots of one row
  constexpr uint32_t ES = 2 * T2_BM * 4;        // byte stride between classes of the window table / seen-window entries

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous range of 128-user tiles of this CTA, taken two at a time ("halves")
  const int tile_begin = static_cast<int>(static_cast<int64_t>(a.n_user_tiles) * blockIdx.x / gridDim.x);
  const int tile_end = static_cast<int>(static_cast<int64_t>(a.n_user_tiles) * (blockIdx.x + 1) / gridDim.x);
  const int n_sweeps = (tile_end - tile_begin + 1) / 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 2); }
    for (int i = 0; i < 2 * T2_ACC; ++i) { mbar_init(BAR(T_FULL + i), 1); mbar_init(BAR(T_EMPTY + i), 4 * T2_SUBS); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(A_FULL + i), 8); mbar_init(BAR(A_EMPTY + i), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(ClassTable) / 4); i += T2_THREADS)
    reinterpret_cast<int*>(tab)[i] = reinterpret_cast<const int*>(a.tab)[i];
  for (int i = threadIdx.x; i < T2_SUBS * T2_RING * 2 * T2_BM; i += T2_THREADS) ring_put(base + S::off_ring + 8u * i, 0.f, T2_EMPTY);
  for (int i = threadIdx.x; i < 2 * T2_BM; i += T2_THREADS) ring_put(base + S::off_thr + 8u * i, 0.f, -1);
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = tab->n_tiles;
  auto a_buf = [](int sweep) { return A_BUFS == 2 ? static_cast<uint32_t>(sweep & 1) : 0u; };
  auto a_phase = [](int sweep) { return A_BUFS == 2 ? static_cast<uint32_t>((sweep >> 1) & 1) : static_cast<uint32_t>(sweep & 1); };

  if (warp == 0) {
    // =========================================== TMA producer ==============================================
    if (lane == 0) {
      uint32_t slot_global = 0;
      for (int j = 0; j < n_sweeps; ++j) {
        for (int it = 0; it < n_tiles; ++it, ++slot_global) {
          const uint32_t slot = slot_global % SLOTS, ph = (slot_global / SLOTS) & 1;
          mbar_wait_relaxed(BAR(B_EMPTY + slot), ph ^ 1, 64);
          mbar_expect_tx(BAR(B_FULL + slot), S::B_BYTES);
#pragma unroll
          for (int kh = 0; kh < KP / 64; ++kh)
            tma_load_2d(base + S::off_b + slot * S::B_BYTES + kh * (T2_BN * 128), &map_w, BAR(B_FULL + slot), kh * 64, it * T2_BN);
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ============================ MMA issuers: one thread per half (the bookkeeping and the commit of one group run while
    //                              the tensor pipe works on the other half's group: one issuer alone reaches 80 % of the pipe)
    if (lane == 0) {
      const int my = warp >> 1;
      uint32_t slot_global = 0, acc = 0;
      uint32_t upar = 0;          // bit (g * 3 + s): parity of the number of uses of stage s by half g so far
      uint32_t last = 0;          // 2 bits per stage: 0 = never used, 1 + g = last used by half g
      long long w_a = 0, w_t = 0, w_b = 0;
      const long long t_start = DBG ? clock64() : 0;
      for (int j = 0; j < n_sweeps; ++j) {
        const int halves = (tile_begin + 2 * j + 1 < tile_end) ? 2 : 1;
        wait_bar(BAR(A_FULL + a_buf(j)), a_phase(j), w_a);
        tc_fence_after();
        const uint32_t a_tmem = tmem_base + a_buf(j) * (2 * A_COLS) + my * A_COLS;
        for (int it = 0; it < n_tiles; ++it, ++slot_global) {
          const uint32_t slot = slot_global % SLOTS, ph = (slot_global / SLOTS) & 1;
          wait_bar(BAR(B_FULL + slot), ph, w_b);     // (also the idle issuer of a one-half sweep: it must not run ahead of the ring)
          tc_fence_after();
          const uint64_t bd0 = make_desc(base + S::off_b + slot * S::B_BYTES);
#define RL_BOFF(kk) (static_cast<uint64_t>((((kk) >> 2) * (T2_BN * 128) + ((kk) & 3) * 32) >> 4))
          for (int hh = 0; hh < halves; ++hh) {
            if (hh == my) {
              const uint32_t prev = (last >> (2 * acc)) & 3u;
              if (prev) {           // the stage has to be drained by the scan warps of the half that used it last
                const uint32_t g = prev - 1;
                wait_bar(BAR(T_EMPTY + g * T2_ACC + acc), ((upar >> (g * 3 + acc)) & 1u) ^ 1u, w_t);
                tc_fence_after();
              }
              const uint32_t d_tmem = tmem_base + ACC_COL0 + acc * T2_BN;
#pragma unroll
              for (int kk = 0; kk < KP / 16; ++kk)
                tc_mma_f16_ts(d_tmem, a_tmem + kk * 8, bd0 + RL_BOFF(kk), T2_IDESC, kk != 0 ? 1u : 0u);
              tc_commit(BAR(T_FULL + my * T2_ACC + acc));
            }
            upar ^= 1u << (hh * 3 + acc);
            last = (last & ~(3u << (2 * acc))) | (static_cast<uint32_t>(hh + 1) << (2 * acc));
            acc = acc == T2_ACC - 1 ? 0 : acc + 1;
          }
#undef RL_BOFF
          if (my < halves) tc_commit(BAR(B_EMPTY + slot));
          else mbar_arrive(BAR(B_EMPTY + slot));        // (the other issuer's commit tells when the MMAs are done with the slot)
        }
        if (my < halves) tc_commit(BAR(A_EMPTY + a_buf(j)));
        else mbar_arrive(BAR(A_EMPTY + a_buf(j)));
      }
      if (DBG && a.dbg && blockIdx.x == 0) { long long* d = a.dbg + 4 * my; d[0] = clock64() - t_start; d[1] = w_a; d[2] = w_t; d[3] = w_b; }
    }
  } else if (warp >= 4 && warp < 4 + 8 * T2_SUBS) {
    // ==================== scan warps: (half, column part, TMEM lane quarter) ==================================
    // The kernel is bound by the integer / compare pipe: everything in the tile and chunk loops is kept in registers and
    // advanced incrementally (no divisions, no re-derived addresses), one copy of the chunk body.
    const int idx = warp - 4, hh = idx / (4 * T2_SUBS), sub = (idx >> 2) % T2_SUBS, q = idx & 3;
    const int rowg = hh * T2_BM + q * 32 + lane;                 // row inside the CTA's pair of user tiles
    const uint32_t ring_row = base + S::off_ring + sub * (T2_RING * RS) + 8u * rowg;   // this producer's ring of the row
    const uint32_t thr_addr = base + S::off_thr + 8u * rowg;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ACC_COL0 + sub * (T2_BN / T2_SUBS);
    const uint32_t bar_full = BAR(T_FULL + hh * T2_ACC), bar_empty = BAR(T_EMPTY + hh * T2_ACC);
    uint32_t tail = 0;            // records pushed so far (ring slot = tail % T2_RING)
    uint32_t acc = 0;             // accumulator stage of the next half-tile of this half
    uint32_t fpar = 0;            // bit s: parity of this half's next use of accumulator stage s
    long long w_full = 0, n_push = 0, n_slow = 0, c_ld = 0, c_scan = 0;
    const long long t_start = DBG ? clock64() : 0;
    auto wait_slot = [&](uint32_t addr) { t2_wait_slot(addr); };   // ring full: the selection thread frees slots in order
    for (int j = 0; j < n_sweeps; ++j) {
      const int halves = (tile_begin + 2 * j + 1 < tile_end) ? 2 : 1;
      if (hh >= halves) { acc = (acc + static_cast<uint32_t>(n_tiles)) % T2_ACC; continue; }     // (only the last sweep)
      const uint32_t step = halves == 2 ? 2u : 1u;
      if (halves == 2 && hh == 1) acc = acc == T2_ACC - 1 ? 0 : acc + 1;      // half 1 takes the stage after half 0's
      const int64_t user = static_cast<int64_t>(tile_begin + 2 * j + hh) * T2_BM + q * 32 + lane;
      const bool valid = user < a.n_users;
      const float floor_thr = (valid && !dflag(1)) ? -CUDART_INF_F : CUDART_INF_F;    // rows past the end never push
      const uint32_t eps_row = base + S::off_eps + (static_cast<uint32_t>(j & 1) * T2_NCLS * 2 * T2_BM + rowg) * 4u;
      // the windows of this sweep were written by the selection warps before they released the user tile: they are read
      // after an accumulator barrier of the sweep (which orders them), never before
      int cls = -1, next_cls_tile = 0;
      float eps_c = 0.f;
      int pos = sub * (T2_BN / T2_SUBS);             // position of the first column of the next chunk
      for (int it = 0; it < n_tiles; ++it) {
        const uint32_t aph = (fpar >> acc) & 1u;
        fpar ^= 1u << acc;
        wait_bar(bar_full + 8u * acc, aph, w_full);
        tc_fence_after();
        if (it >= next_cls_tile) {
          do { ++cls; next_cls_tile = tab->first_tile[cls + 1]; } while (it >= next_cls_tile);
          eps_c = __int_as_float(lds_volatile(eps_row + cls * ES));
        }
        uint32_t t_addr = t_lane + acc * T2_BN;
        if (dflag(4)) {                              // debug: MMA / TMA ceiling (stages released unread)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_empty + 8u * acc);
        } else {
          long long tq0 = 0;
          if constexpr (DBG) tq0 = clock64();
#pragma unroll 1
          for (int ch = 0; ch < T2_BN / T2_SUBS / T2_CH; ++ch, t_addr += T2_CH, pos += T2_CH) {
            float sg[T2_CH];
            tc_ld(t_addr, sg);
            // the selection thread publishes the n-th best LOWER bound of this sweep; an item of this class can only
            // matter if its upper bound reaches it.  Re-read for every chunk: the bound moves fastest right after the
            // sweep's start (while the load is in flight)
            float lth;
            int tag;
            ring_get(thr_addr, lth, tag);
            if (tag != j) lth = -CUDART_INF_F;
            const float thr = fmaxf(__fsub_rd(lth, eps_c), floor_thr);
            tc_wait_dep(sg);
            if constexpr (DBG) { const long long t1 = clock64(); c_ld += t1 - tq0; tq0 = t1; }
            if (ch == T2_BN / T2_SUBS / T2_CH - 1) {     // the last chunk sits in registers: the stage goes back to the MMA
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_empty + 8u * acc);
            }
            if (DUMP) {
              if (valid) {
#pragma unroll
                for (int jj = 0; jj < T2_CH; ++jj) a.dump[user * a.dump_ld + pos + jj] = sg[jj];
              }
            } else if (!dflag(8)) {
              bool any = false;
#pragma unroll
              for (int b = 0; b < T2_CH / 8; ++b) {
                const float m8 = max3(max3(sg[8 * b], sg[8 * b + 1], sg[8 * b + 2]), max3(sg[8 * b + 3], sg[8 * b + 4], sg[8 * b + 5]),
                                      fmaxf(sg[8 * b + 6], sg[8 * b + 7]));
                if (m8 > thr) {
                  // (score, position) records for the selection thread: count the hits, wait once for that much room (a
                  // slot is free again once its position reads EMPTY), straight predicated stores
                  uint32_t c = 0;
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) c += sg[8 * b + jj] > thr ? 1u : 0u;
                  {
                    const uint32_t last = ring_row + ((tail + c - 1) % T2_RING) * RS + 4;
                    if (lds_volatile(last) != T2_EMPTY) wait_slot(last);
                  }
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) {
                    if (sg[8 * b + jj] > thr) {
                      ring_put(ring_row + (tail % T2_RING) * RS, sg[8 * b + jj], pos + 8 * b + jj);
                      ++tail;
                    }
                  }
                  if constexpr (DBG) n_push += c;
                  any = true;
                }
              }
              if constexpr (DBG) { if (any) ++n_slow; }
              __syncwarp();     // the pushes diverge; tcgen05.ld / wait need the whole warp
            }
            if constexpr (DBG) { const long long t1 = clock64(); c_scan += t1 - tq0; tq0 = t1; }
          }
        }
        pos += T2_BN - T2_BN / T2_SUBS;              // to this warp's columns of the next item tile
        acc += step;
        if (acc >= T2_ACC) acc -= T2_ACC;
      }
      if (halves == 2 && hh == 1) acc = acc == 0 ? T2_ACC - 1 : acc - 1;      // back to the common count of half-tiles
      if (!DUMP && valid) {
        const uint32_t addr = ring_row + (tail % T2_RING) * RS;
        if (lds_volatile(addr + 4) != T2_EMPTY) wait_slot(addr + 4);
        ring_put(addr, 0.f, T2_END);
        ++tail;
      }
      __syncwarp();
    }
    if (DBG && a.dbg && blockIdx.x == 0 && lane == 0) {
      long long* d = a.dbg + 8 + 4 * idx;
      d[0] = clock64() - t_start; d[1] = w_full; d[2] = c_ld; d[3] = c_scan;
    }
  } else if (warp >= 4 + 8 * T2_SUBS) {
    // ======================== selection warps: user-tile loaders + per-row selection ==========================
    const int hh = (warp - 4 - 8 * T2_SUBS) >> 2, q = warp & 3;
    const int rowg = hh * T2_BM + q * 32 + lane;
    const uint32_t ring_a = base + S::off_ring + 8u * rowg, ring_b = ring_a + T2_RING * RS;   // the two producers of the row
    const uint32_t cand_row = base + S::off_cand + 8u * rowg;
    const uint32_t thr_addr = base + S::off_thr + 8u * rowg;
    const uint32_t win_row = base + S::off_win + 4u * rowg;       // seen-position window: entry i at + i * ES
    const uint32_t t_row0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // writes the user tile of sweep j (this thread: one row of half hh) into TMEM and the row's windows into shared memory
    auto load_tile = [&](int j) {
      const int tile = tile_begin + 2 * j + hh;
      const int64_t user = static_cast<int64_t>(tile) * T2_BM + q * 32 + lane;
      const bool valid = tile < tile_end && user < a.n_users;
      mbar_wait_relaxed(BAR(A_EMPTY + a_buf(j)), a_phase(j) ^ 1, 100);
      tc_fence_after();
      const uint32_t t_row = t_row0 + a_buf(j) * (2 * A_COLS) + hh * A_COLS;
      const float4* up = reinterpret_cast<const float4*>(a.uf + user * a.ldf);
      float su = 1.0f;
      if (valid) {
        float mx = 0.f;
        for (int c = 0; c < a.ldf / 4; ++c) {
          const float4 x = up[c];
          mx = fmaxf(fmaxf(mx, fmaxf(fabsf(x.x), fabsf(x.y))), fmaxf(fabsf(x.z), fabsf(x.w)));
        }
        su = pow2_scale(mx);
      }
      float nn = 0.f, rr = 0.f, hn = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < KP; c0 += 64) {
        uint32_t hi[32];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid && c0 + 4 * c < a.ldf) v = up[c0 / 4 + c];
          v.x *= su; v.y *= su; v.z *= su; v.w *= su;
          const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
          const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
          nn += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          hn += f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y;
          const float r0 = v.x - f0.x, r1 = v.y - f0.y, r2 = v.z - f1.x, r3 = v.w - f1.y;   // exact in fp32
          rr += r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
          hi[2 * c] = *reinterpret_cast<const uint32_t*>(&h0);
          hi[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&h1);
        }
        tc_st32(t_row + c0 / 2, hi);
      }
      if (DUMP && valid) a.dump_su[user] = su;
      // the windows of this row, one per item class, for the scan threads and the selection thread of sweep j
      const float e1 = sqrtf(rr) * 1.0001f, e2 = sqrtf(hn) * 1.0001f, e3 = sqrtf(nn) * 1.0001f;
      const uint32_t eps_row = base + S::off_eps + (static_cast<uint32_t>(j & 1) * T2_NCLS * 2 * T2_BM + rowg) * 4u;
#pragma unroll
      for (int c = 0; c < T2_NCLS; ++c) {
        const float wcc = tab->wc[c], rcc = tab->rc[c];
        const float eps = (e1 * wcc + e2 * rcc + a.acc_rel * e2 * (wcc + rcc) + e3 * (6.1e-8f * wcc + 6.0e-14f * tab->vmax_raw)) * 1.0001f + 1e-30f;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(eps_row + c * ES), "r"(__float_as_int(eps)) : "memory");
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(A_FULL + a_buf(j)));
    };
    for (int j = 0; j < A_BUFS && j < n_sweeps; ++j) load_tile(j);
    if (DUMP) {
      for (int j = A_BUFS; j < n_sweeps; ++j) load_tile(j);
    } else {
      uint32_t head_a = 0, head_b = 0;
      long long c_idle = 0, n_rec = 0, n_cmp = 0, n_pass = 0, n_iter = 0, c_pass = 0, n_refill = 0;
      const long long t_start = clock64();
      for (int j = 0; j < n_sweeps; ++j) {
        const int halves = (tile_begin + 2 * j + 1 < tile_end) ? 2 : 1;
        const int64_t user = static_cast<int64_t>(tile_begin + 2 * j + hh) * T2_BM + q * 32 + lane;
        const bool valid = hh < halves && user < a.n_users;
        const uint32_t eps_row = base + S::off_eps + (static_cast<uint32_t>(j & 1) * T2_NCLS * 2 * T2_BM + rowg) * 4u;
        // ---- per-row state of this sweep (registers) ----
        float tk[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) tk[i] = i < NT - a.n ? CUDART_INF_F : -CUDART_INF_F;
        float lth = -CUDART_INF_F, eps = 0.f;
        int cnt = 0, cls_lo_pos = 0, cls_hi_pos = 0, real_end_pos = 0;
        bool over = false, done_a = !valid, done_b = T2_SUBS == 1 || !valid;
        // the row's seen positions (ascending); T2_WIN of them, from offset woff, live in shared memory; entries before
        // win[wp] are below every position that can still arrive
        const int32_t* seen = a.seen_pos;
        int seen_len = 0, woff = 0, wp = 0;
        if (valid && a.seen_indptr) {
          const int64_t b0 = a.seen_indptr[user], e0 = a.seen_indptr[user + 1];
          if (e0 > a.seen_cap || b0 > e0 || e0 - b0 > INT_MAX / 2) {      // stale capacity: the exact kernel takes the row
            over = true; lth = CUDART_INF_F;
            ring_put(thr_addr, lth, j);
          } else {
            seen += b0;
            seen_len = static_cast<int>(e0 - b0);
          }
        }
        auto refill = [&]() {                                     // window <- seen[woff .. woff + T2_WIN)
          t2_refill_window(seen, seen_len, woff, win_row, ES);
          wp = 0;
          if constexpr (DBG) ++n_refill;
        };
        if (valid) refill();
        auto compact = [&]() {
          cnt = t2_compact_list(cand_row, RS, cnt, lth);
          if constexpr (DBG) ++n_cmp;
        };
        // The list is still full after a compaction (thresholds rise slowly while the densest classes go by): its positions
        // move to a pool block in global memory and the list starts again; they are re-scored with the rest at the end.
        // No block left: the exact kernel takes the row.
        int n_blk = 0;
        auto spill = [&]() -> bool {
          if (n_blk == T2_SPILL_MAX) return false;
          if (!t2_spill_block(a.pool, a.pool_next, a.pool_blocks, a.cand_ids + user * T2_CAND, cand_row, RS, cnt, n_blk)) return false;
          ++n_blk;
          cnt = 0;
          return true;
        };
        auto full_list = [&]() {
          if (!spill()) { over = true; cnt = 0; lth = CUDART_INF_F; ring_put(thr_addr, lth, j); }
        };
        long long last_pass = clock64();
        int period = T2_PERIOD_MIN;                          // clocks between passes when no ring is filling up; grows with the sweep
        // Passes are run in lock step by the 32 rows of the warp (a row's record costs the same few hundred instructions of
        // the warp whether one lane or all of them have one, so records are left to accumulate): when a ring is half full, or
        // every `period` clocks.  Refills of the seen window and compactions of the candidate list run for all rows at once.
        while (!__all_sync(FULL_MASK, done_a && done_b)) {
          bool pending = false, urgent = false;
          if (!done_a) {
            pending = lds_volatile(ring_a + (head_a % T2_RING) * RS + 4) != T2_EMPTY;
            urgent = lds_volatile(ring_a + ((head_a + T2_RING / 2 - 1) % T2_RING) * RS + 4) != T2_EMPTY;
          }
          if (!done_b) {
            pending |= lds_volatile(ring_b + (head_b % T2_RING) * RS + 4) != T2_EMPTY;
            urgent |= lds_volatile(ring_b + ((head_b + T2_RING / 2 - 1) % T2_RING) * RS + 4) != T2_EMPTY;
          }
          const bool go = __any_sync(FULL_MASK, urgent) || (__any_sync(FULL_MASK, pending) && clock64() - last_pass > period);
          if (!go) {
            const long long t0 = DBG ? clock64() : 0;
            __nanosleep(T2_POLL_NS);
            if constexpr (DBG) c_idle += clock64() - t0;
            continue;
          }
          if constexpr (DBG) ++n_pass;
          const long long t_pass0 = DBG ? clock64() : 0;
          if (T2_JOINT && __any_sync(FULL_MASK, wp >= 3 * T2_WIN / 4)) {     // seen windows running low: refilled together
            if (wp >= T2_WIN / 4) { woff += wp; refill(); }
          }
          int last_pos = 0;
          for (;;) {
            // the record with the smaller position of the two producers' next ones (they are at most two tiles apart)
            float sc = 0.f, sc_b = 0.f;
            int pos = T2_EMPTY, pos_b = T2_EMPTY;
            const uint32_t addr_a = ring_a + (head_a % T2_RING) * RS, addr_b = ring_b + (head_b % T2_RING) * RS;
            if (!done_a) ring_get(addr_a, sc, pos);
            if (!done_b) ring_get(addr_b, sc_b, pos_b);
            const bool take_b = pos_b != T2_EMPTY && (pos == T2_EMPTY || pos_b < pos);
            if (take_b) { pos = pos_b; sc = sc_b; }
            if (!__any_sync(FULL_MASK, pos != T2_EMPTY)) break;
            if constexpr (DBG) ++n_iter;
            if (pos == T2_EMPTY) continue;
            if (take_b) { ring_put(addr_b, 0.f, T2_EMPTY); ++head_b; }
            else { ring_put(addr_a, 0.f, T2_EMPTY); ++head_a; }
            if (pos == T2_END) {
              if (take_b) done_b = true; else done_a = true;
              continue;
            }
            if constexpr (DBG) ++n_rec;
            if (over) continue;
            last_pos = pos;
            if (pos < cls_lo_pos || pos >= cls_hi_pos) {          // record of another item class: its window, its extent
              int c = 0;
              while (pos >= tab->first_tile[c + 1] * T2_BN) ++c;
              cls_lo_pos = tab->first_tile[c] * T2_BN;
              cls_hi_pos = tab->first_tile[c + 1] * T2_BN;
              real_end_pos = cls_lo_pos + tab->count[c];
              eps = __int_as_float(lds_volatile(eps_row + c * ES));
            }
            if (pos >= real_end_pos) continue;        // padding position at the end of a class
            if (!dflag(2)) {
              // seen items: nothing below (tile - 2) can arrive any more; look ahead from there without committing
              const int low = ((pos >> 7) - 2) << 7;
              int nx = lds_volatile(win_row + wp * ES);
              while (nx < low) {
                if (++wp == T2_WIN) { woff += T2_WIN; refill(); }
                nx = lds_volatile(win_row + wp * ES);
              }
              int i = wp;
              while (nx < pos) {
                if (++i == T2_WIN) {                  // window exhausted inside the look-ahead (heavy users): bisection
                  int lo = woff + T2_WIN, hi = seen_len;
                  while (lo < hi) {
                    const int mid = lo + ((hi - lo) >> 1);
                    if (seen[mid] < pos) lo = mid + 1; else hi = mid;
                  }
                  nx = lo < seen_len ? seen[lo] : INT_MAX;
                  break;
                }
                nx = lds_volatile(win_row + i * ES);
              }
              if (nx == pos) continue;                // already seen by this user
            }
            const float lb = __fsub_rd(sc, eps), hb = __fadd_ru(sc, eps);
            if (lb > tk[NT - 1]) {
#pragma unroll
              for (int i = NT - 1; i > 0; --i) tk[i] = fmaxf(tk[i], fminf(tk[i - 1], lb));
              tk[0] = fmaxf(tk[0], lb);
              lth = tk[NT - 1];
              // published at once: while the threshold is low the scan threads feed this pass faster than it drains
              ring_put(thr_addr, lth, j);
            }
            if (hb >= lth) {
              ring_put(cand_row + cnt * RS, hb, pos);
              if (++cnt == T2_CAND) {                 // emergency: normally the joint compaction below keeps the list short
                compact();
                if (cnt > T2_CAND - 4) full_list();
              }
            }
          }
          if (T2_JOINT && __any_sync(FULL_MASK, cnt >= T2_CAND - 10)) {     // candidate lists filling up: compacted together
            if (cnt > 0 && !over) {
              compact();
              if (cnt > T2_CAND - 4) full_list();
            }
          }
          // thresholds settle as the sweep goes on: let more records accumulate per pass
          const int t_now = __reduce_max_sync(FULL_MASK, last_pos) >> 7;
          period = min(T2_PERIOD_MAX, max(period, T2_PERIOD_MIN + 150 * t_now));
          last_pass = clock64();
          if constexpr (DBG) c_pass += last_pass - t_pass0;
        }
        if (valid) {
          // ---- end of the sweep: final candidates (positions) to global memory ----
          if (!over) compact();
          if (!over && n_blk > 0 && cnt > 0 && !spill()) over = true;
          int32_t* out = a.cand_ids + user * T2_CAND;
          if (!over && n_blk > 0) {
            out[0] = n_blk;
            a.cand_cnt[user] = -2;
          } else {
            for (int f = 0; f < cnt; ++f) {
              float hb; int pos;
              ring_get(cand_row + f * RS, hb, pos);
              out[f] = pos;
            }
            a.cand_cnt[user] = over ? -1 : cnt;
          }
        }
        // the user tile of the sweep after next goes into the TMEM buffer this sweep has just finished with
        __syncwarp();       // tcgen05.st needs the whole warp
        if (j + A_BUFS < n_sweeps) load_tile(j + A_BUFS);
      }
      if (DBG && a.dbg && blockIdx.x == 0 && lane == 0) {
        long long* d = a.dbg + 80 + 8 * (warp - 4 - 8 * T2_SUBS);
        d[0] = clock64() - t_start; d[1] = c_idle; d[2] = n_rec; d[3] = n_cmp; d[4] = n_pass; d[5] = n_iter; d[6] = c_pass; d[7] = n_refill;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- exact re-scoring of the candidates: one warp per user, float64 dot products in the exact kernel's order -------
// candidate c of the row has item id fetch(c); PER = candidates per lane
template <int KP, int PER, typename Fetch>
__device__ __forceinline__ void rescore_row(const float* __restrict__ uf, const float* __restrict__ vf, int64_t user, int n, int cnt, Fetch fetch,
                                            int32_t* __restrict__ idx_out, double* __restrict__ score_out, int lane) {
  int id[PER];
  double sc[PER];
  const float* up = uf + user * KP;
#pragma unroll
  for (int p = 0; p < PER; ++p) {
    const int c = lane + 32 * p;
    id[p] = c < cnt ? fetch(c) : -1;
    sc[p] = -CUDART_INF;
    if (id[p] >= 0) {
      const float* vp = vf + static_cast<int64_t>(id[p]) * KP;
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
      sc[p] = accd;
    }
  }
  // rank of every candidate = number of candidates that beat it (score desc, index asc)
  int rank[PER];
#pragma unroll
  for (int p = 0; p < PER; ++p) rank[p] = 0;
  for (int c = 0; c < cnt; ++c) {
    double os = 0.0;
    int oi = 0;
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const double s_p = __shfl_sync(FULL_MASK, sc[p], c & 31);
      const int i_p = __shfl_sync(FULL_MASK, id[p], c & 31);
      if ((c >> 5) == p) { os = s_p; oi = i_p; }
    }
#pragma unroll
    for (int p = 0; p < PER; ++p)
      if (id[p] >= 0 && beats(os, oi, sc[p], id[p])) ++rank[p];
  }
#pragma unroll
  for (int p = 0; p < PER; ++p) {
    if (id[p] >= 0 && rank[p] < n) {
      idx_out[user * n + rank[p]] = id[p];
      if (score_out) score_out[user * n + rank[p]] = sc[p];
    }
  }
  for (int r = cnt + lane; r < n; r += 32) {       // fewer than n unseen items
    idx_out[user * n + r] = -1;
    if (score_out) score_out[user * n + r] = -CUDART_INF;
  }
}

template <int KP>
__global__ void __launch_bounds__(256) rescore_kernel(const float* __restrict__ uf, const float* __restrict__ vf, int64_t n_users, int n,
                                                      const int32_t* __restrict__ cand_ids, const int32_t* __restrict__ cand_cnt,
                                                      const int32_t* __restrict__ perm, int32_t* __restrict__ idx_out, double* __restrict__ score_out,
                                                      int32_t* __restrict__ overflow_rows, unsigned int* __restrict__ overflow_count,
                                                      int64_t row_offset, int32_t* __restrict__ spill_rows, unsigned int* __restrict__ spill_count) {
  const int lane = threadIdx.x & 31;
  const int64_t user = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (user >= n_users) return;
  const int cnt = cand_cnt[user];
  if (cnt == -2) {            // spilled list: re-scored by rescore_spill_kernel
    if (lane == 0) spill_rows[atomicAdd(spill_count, 1u)] = static_cast<int32_t>(user);
    return;
  }
  if (cnt < 0) {
    if (lane == 0) overflow_rows[atomicAdd(overflow_count, 1u)] = static_cast<int32_t>(user + row_offset);
    return;
  }
  const int32_t* cp = cand_ids + user * T2_CAND;
  rescore_row<KP, (T2_CAND + 31) / 32>(uf, vf, user, n, cnt, [&](int c) { return perm[cp[c]]; }, idx_out, score_out, lane);
}

// rows whose candidates went to pool blocks: the concatenation of the blocks, up to T2_SPILL_MAX * T2_CAND candidates
template <int KP>
__global__ void __launch_bounds__(256) rescore_spill_kernel(const float* __restrict__ uf, const float* __restrict__ vf, int n,
                                                            const int32_t* __restrict__ cand_ids, const int32_t* __restrict__ pool,
                                                            const int32_t* __restrict__ perm, int32_t* __restrict__ idx_out,
                                                            double* __restrict__ score_out, const int32_t* __restrict__ spill_rows,
                                                            const unsigned int* __restrict__ spill_count) {
  const int lane = threadIdx.x & 31;
  const unsigned int total = *spill_count;
  for (unsigned int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < total; i += gridDim.x * 8) {
    const int64_t user = spill_rows[i];
    const int32_t* hdr = cand_ids + user * T2_CAND;
    const int nb = hdr[0];
    int cnt = 0;
    for (int b = 0; b < nb; ++b) cnt += hdr[1 + T2_SPILL_MAX + b];
    auto fetch = [&](int c) {
      int b = 0;
      while (c >= hdr[1 + T2_SPILL_MAX + b]) { c -= hdr[1 + T2_SPILL_MAX + b]; ++b; }
      return perm[pool[static_cast<size_t>(hdr[1 + b]) * T2_CAND + c]];
    };
    rescore_row<KP, (T2_CAND * (T2_SPILL_MAX > 0 ? T2_SPILL_MAX : 1) + 31) / 32>(uf, vf, user, n, cnt, fetch, idx_out, score_out, lane);
  }
}

// ---- preprocessing of the item factors ------------------------------------------------------------------------------
constexpr int PP_ROWS = 1024;     // items per block of the placement kernels

// column sums in float64: deterministic two-stage reduction (partial[block][kp], then one block)
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ v, int64_t n, int kp, double* __restrict__ partial) {
  __shared__ double red[256 * 4];
  const int c4n = kp / 4, c4 = threadIdx.x % c4n, rl = threadIdx.x / c4n, rstep = 256 / c4n;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = per * blockIdx.x, r1 = min(n, r0 + per);
  for (int64_t r = r0 + rl; r < r1; r += rstep) {
    const float4 x = *reinterpret_cast<const float4*>(v + r * kp + 4 * c4);
    s0 += x.x; s1 += x.y; s2 += x.z; s3 += x.w;
  }
  red[threadIdx.x * 4 + 0] = s0; red[threadIdx.x * 4 + 1] = s1; red[threadIdx.x * 4 + 2] = s2; red[threadIdx.x * 4 + 3] = s3;
  __syncthreads();
  if (rl == 0) {
    for (int o = 1; o < rstep; ++o) {
      s0 += red[(o * c4n + c4) * 4 + 0]; s1 += red[(o * c4n + c4) * 4 + 1];
      s2 += red[(o * c4n + c4) * 4 + 2]; s3 += red[(o * c4n + c4) * 4 + 3];
    }
    double* p = partial + static_cast<int64_t>(blockIdx.x) * kp + 4 * c4;
    p[0] = s0; p[1] = s1; p[2] = s2; p[3] = s3;
  }
}
__global__ void __launch_bounds__(128) colsum_final_kernel(const double* __restrict__ partial, int nblk, int kp, int64_t n, float* __restrict__ vbar,
                                                           ClassTable* __restrict__ tab) {
  const int c = threadIdx.x;
  if (c < kp) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partial[static_cast<int64_t>(b) * kp + c];
    vbar[c] = static_cast<float>(s / static_cast<double>(n));
  }
  if (c == 0) {
    tab->wmax_bits = tab->amax_bits = tab->vmax_bits = 0u;
    for (int i = 0; i < T2_NCLS; ++i) tab->wc_bits[i] = tab->rc_bits[i] = 0u;
  }
}

// w_i = fl32(v_i - vbar): norms, largest norm, largest entry, largest |v_i| (kp / 4 lanes per row)
__global__ void __launch_bounds__(256) centre_norm_kernel(const float* __restrict__ v, int64_t n, int kp, const float* __restrict__ vbar,
                                                          float* __restrict__ wnorm, ClassTable* __restrict__ tab) {
  const int c4n = kp / 4;
  const int64_t total = n * c4n;
  float wmax = 0.f, amax = 0.f, vmax = 0.f;
  for (int64_t p0 = static_cast<int64_t>(blockIdx.x) * blockDim.x; p0 < total; p0 += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p = p0 + threadIdx.x;
    float nn = 0.f, vv = 0.f, am = 0.f;
    int64_t r = 0;
    if (p < total) {
      r = p / c4n;
      const int c4 = static_cast<int>(p - r * c4n);
      const float4 x = *reinterpret_cast<const float4*>(v + r * kp + 4 * c4);
      const float4 m = *reinterpret_cast<const float4*>(vbar + 4 * c4);
      const float w0 = x.x - m.x, w1 = x.y - m.y, w2 = x.z - m.z, w3 = x.w - m.w;
      nn = w0 * w0 + w1 * w1 + w2 * w2 + w3 * w3;
      vv = x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
      am = fmaxf(fmaxf(fabsf(w0), fabsf(w1)), fmaxf(fabsf(w2), fabsf(w3)));
    }
    for (int o = c4n >> 1; o > 0; o >>= 1) {      // c4n is a power of two <= 32 and rows do not straddle warps
      nn += __shfl_xor_sync(FULL_MASK, nn, o);
      vv += __shfl_xor_sync(FULL_MASK, vv, o);
    }
    if (p < total) {
      const float wn = sqrtf(nn) * 1.00001f;
      if ((threadIdx.x & (c4n - 1)) == 0) wnorm[r] = wn;
      wmax = fmaxf(wmax, wn);
      vmax = fmaxf(vmax, sqrtf(vv) * 1.00001f);
      amax = fmaxf(amax, am);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wmax = fmaxf(wmax, __shfl_xor_sync(FULL_MASK, wmax, o));
    amax = fmaxf(amax, __shfl_xor_sync(FULL_MASK, amax, o));
    vmax = fmaxf(vmax, __shfl_xor_sync(FULL_MASK, vmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&tab->wmax_bits, __float_as_uint(wmax));
    atomicMax(&tab->amax_bits, __float_as_uint(amax));
    atomicMax(&tab->vmax_bits, __float_as_uint(vmax));
  }
}

__device__ __forceinline__ int class_of(float wn, float wmax) {
  if (!(wn > 0.f) || !(wmax > 0.f)) return 0;
  const float ratio = wmax / wn;                     // >= 1 up to rounding
  int c = ilogbf(ratio);
  c = c < 0 ? 0 : (c > T2_NCLS - 1 ? T2_NCLS - 1 : c);
  return T2_NCLS - 1 - c;                            // bin 0 = smallest norms = narrowest window
}

// class of every item, per-block class counts
__global__ void __launch_bounds__(PP_ROWS) classify_kernel(const float* __restrict__ wnorm, int64_t n, const ClassTable* __restrict__ tab,
                                                           unsigned char* __restrict__ cls, int* __restrict__ block_counts) {
  __shared__ int cnt[T2_NCLS];
  if (threadIdx.x < T2_NCLS) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t r = static_cast<int64_t>(blockIdx.x) * PP_ROWS + threadIdx.x;
  if (r < n) {
    const int c = class_of(wnorm[r], __uint_as_float(tab->wmax_bits));
    cls[r] = static_cast<unsigned char>(c);
    atomicAdd(&cnt[c], 1);
  }
  __syncthreads();
  if (threadIdx.x < T2_NCLS) block_counts[blockIdx.x * T2_NCLS + threadIdx.x] = cnt[threadIdx.x];
}

// The scan order of the classes.  A row's candidate records are the items that beat its running n-th best: scanned in
// an order in which the scores tend to RISE (after 20 ALS iterations the far-out items of the widest norm bins score
// lowest, the members of the final top-n sit in the middle bins) every class restarts the harmonic sum n ln(size / n) of
// records.  The exact top-n of T2_SAMPLE evenly spaced users says where the winners come from: classes are scanned by
// falling density (sampled members / class size); bins nobody's top-n touched follow by rising norm.  Any order is exact.
constexpr int T2_SAMPLE = 64;
__global__ void __launch_bounds__(T2_SAMPLE) sample_users_kernel(int64_t n_users, int n_sample, int32_t* __restrict__ list) {
  if (static_cast<int>(threadIdx.x) < n_sample) list[threadIdx.x] = static_cast<int32_t>((n_users * threadIdx.x) / n_sample);
}

// class bases (whole tiles) in scan order, per-block offsets inside the classes, the scale
__global__ void __launch_bounds__(32) class_scan_kernel(int* __restrict__ block_counts, int nblk, ClassTable* __restrict__ tab,
                                                        const int32_t* __restrict__ sample_idx, int n_sample_entries,
                                                        const float* __restrict__ wnorm, int64_t n_items) {
  __shared__ int total[T2_NCLS];
  __shared__ int hist[T2_NCLS];
  const int c = threadIdx.x;
  if (c < T2_NCLS) {
    int run = 0;
    for (int b = 0; b < nblk; ++b) {
      const int x = block_counts[b * T2_NCLS + c];
      block_counts[b * T2_NCLS + c] = run;
      run += x;
    }
    total[c] = run;
    hist[c] = 0;
  }
  __syncwarp();
  const float wmax = __uint_as_float(tab->wmax_bits);
  for (int e = c; e < n_sample_entries; e += 32) {
    const int id = sample_idx[e];
    if (id >= 0 && id < n_items) atomicAdd(&hist[class_of(wnorm[id], wmax)], 1);
  }
  __syncwarp();
  if (c == 0) {
    int order[T2_NCLS];
    for (int i = 0; i < T2_NCLS; ++i) order[i] = i;
    for (int i = 1; i < T2_NCLS; ++i) {            // stable insertion sort by falling hist / total
      const int x = order[i];
      int j = i;
      while (j > 0) {
        const int y = order[j - 1];
        if (static_cast<long long>(hist[x]) * total[y] > static_cast<long long>(hist[y]) * total[x]) { order[j] = y; --j; }
        else break;
      }
      order[j] = x;
    }
    int tile = 0;
    for (int l = 0; l < T2_NCLS; ++l) {
      const int nc = order[l];
      tab->label_of[nc] = l;
      tab->first_tile[l] = tile;
      tab->count[l] = total[nc];
      tile += (total[nc] + T2_BN - 1) / T2_BN;
    }
    tab->first_tile[T2_NCLS] = tile;
    tab->n_tiles = tile;
    tab->scale = pow2_scale(__uint_as_float(tab->amax_bits));
  }
}

// stable placement: position of item r = class base + (number of items of its class with a smaller id); writes the
// position -> id map and the fp16 row of the centred, scaled factors; per-class maxima of |w s| and |w s - fp16(w s)|.
// cls[] holds norm bins on entry and class labels (scan order) on exit.
template <int KPH>
__global__ void __launch_bounds__(PP_ROWS) place_kernel(const float* __restrict__ v, int64_t n, int kp, const float* __restrict__ vbar,
                                                        const float* __restrict__ wnorm, unsigned char* __restrict__ cls,
                                                        const int* __restrict__ block_offsets, ClassTable* __restrict__ tab,
                                                        int32_t* __restrict__ perm, int32_t* __restrict__ inv, __half* __restrict__ whi) {
  __shared__ int warp_cnt[32][T2_NCLS];
  __shared__ int pos_s[PP_ROWS];
  __shared__ float rmax_s[T2_NCLS], wmax_s[T2_NCLS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * PP_ROWS + tid;
  const int c = r < n ? cls[r] : -1;
  int my_rank = 0;
#pragma unroll
  for (int k = 0; k < T2_NCLS; ++k) {
    const unsigned m = __ballot_sync(FULL_MASK, c == k);
    if (c == k) my_rank = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) warp_cnt[warp][k] = __popc(m);
  }
  if (tid < T2_NCLS) { rmax_s[tid] = 0.f; wmax_s[tid] = 0.f; }
  __syncthreads();
  int pos = -1;
  if (c >= 0) {
    int before = 0;
    for (int wv = 0; wv < warp; ++wv) before += warp_cnt[wv][c];
    pos = tab->first_tile[tab->label_of[c]] * T2_BN + block_offsets[blockIdx.x * T2_NCLS + c] + before + my_rank;
    perm[pos] = static_cast<int32_t>(r);
    inv[r] = pos;
  }
  pos_s[tid] = pos;
  __syncthreads();
  // rows of this block, KPH / 4 lanes per row
  const float s = tab->scale;
  constexpr int C4N = KPH / 4;
  for (int p = tid; p < PP_ROWS * C4N; p += PP_ROWS) {
    const int rl = p / C4N, c4 = p % C4N;
    const int64_t rr = static_cast<int64_t>(blockIdx.x) * PP_ROWS + rl;
    float res = 0.f;
    const int ps = pos_s[rl];
    if (rr < n) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f), m = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * c4 < kp) {
        x = *reinterpret_cast<const float4*>(v + rr * kp + 4 * c4);
        m = *reinterpret_cast<const float4*>(vbar + 4 * c4);
      }
      const float w0 = (x.x - m.x) * s, w1 = (x.y - m.y) * s, w2 = (x.z - m.z) * s, w3 = (x.w - m.w) * s;
      const __half2 h0 = __floats2half2_rn(w0, w1), h1 = __floats2half2_rn(w2, w3);
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      const float d0 = w0 - f0.x, d1 = w1 - f0.y, d2 = w2 - f1.x, d3 = w3 - f1.y;
      res = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      uint2 packed;
      packed.x = *reinterpret_cast<const uint32_t*>(&h0);
      packed.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(whi + static_cast<int64_t>(ps) * KPH + 4 * c4) = packed;
    }
    for (int o = C4N >> 1; o > 0; o >>= 1) res += __shfl_xor_sync(FULL_MASK, res, o);
    if (rr < n && c4 == 0) {
      const int cc = tab->label_of[cls[rr]];
      atomicMax(reinterpret_cast<unsigned*>(&rmax_s[cc]), __float_as_uint(sqrtf(res) * 1.0001f));
      atomicMax(reinterpret_cast<unsigned*>(&wmax_s[cc]), __float_as_uint(wnorm[rr] * s * 1.0001f));
    }
  }
  __syncthreads();
  if (c >= 0) cls[r] = static_cast<unsigned char>(tab->label_of[c]);
  if (tid < T2_NCLS) {
    atomicMax(&tab->rc_bits[tid], __float_as_uint(rmax_s[tid]));
    atomicMax(&tab->wc_bits[tid], __float_as_uint(wmax_s[tid]));
  }
}

// The seen items of every row as positions, ascending: ids ascend inside a row and inside a class, so a stable
// partition of the row by class IS the sort by position.  One warp per row, two passes (count, place).
__global__ void __launch_bounds__(256) seen_positions_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                             int64_t n_rows, int64_t cap, const int32_t* __restrict__ inv,
                                                             const unsigned char* __restrict__ cls, int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < n_rows; row += static_cast<int64_t>(gridDim.x) * 8) {
    const int64_t beg = indptr[row], end = indptr[row + 1];
    if (end > cap || beg > end) continue;
    int cnt[T2_NCLS];
#pragma unroll
    for (int k = 0; k < T2_NCLS; ++k) cnt[k] = 0;
    for (int64_t p0 = beg; p0 < end; p0 += 32) {
      const int64_t p = p0 + lane;
      const int c = p < end ? cls[indices[p]] : -1;
#pragma unroll
      for (int k = 0; k < T2_NCLS; ++k) cnt[k] += __popc(__ballot_sync(FULL_MASK, c == k));
    }
    int run[T2_NCLS];
    int acc = 0;
#pragma unroll
    for (int k = 0; k < T2_NCLS; ++k) { run[k] = acc; acc += cnt[k]; }
    for (int64_t p0 = beg; p0 < end; p0 += 32) {
      const int64_t p = p0 + lane;
      const int id = p < end ? indices[p] : 0;
      const int c = p < end ? cls[id] : -1;
#pragma unroll
      for (int k = 0; k < T2_NCLS; ++k) {
        const unsigned m = __ballot_sync(FULL_MASK, c == k);
        if (c == k) out[beg + run[k] + __popc(m & lt)] = inv[id];
        run[k] += __popc(m);
      }
    }
  }
}

__global__ void class_finish_kernel(ClassTable* __restrict__ tab) {
  if (threadIdx.x == 0) {
    for (int c = 0; c < T2_NCLS; ++c) {
      tab->wc[c] = __uint_as_float(tab->wc_bits[c]);
      tab->rc[c] = __uint_as_float(tab->rc_bits[c]);
    }
    tab->vmax_raw = __uint_as_float(tab->vmax_bits) * tab->scale * 1.0001f;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode2() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map2(rl_context* h, CUtensorMap* map, const __half* ptr, int64_t rows, int kp) {
  EncodeTiledFn enc = get_encode2();
  if (!enc) return fail(h, RL_ERR_CUDA, "score_tc2: cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kp), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kp) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(T2_BN)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, RL_ERR_CUDA, "score_tc2: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return RL_OK;
}

template <int KP, int NT, bool DUMP, bool DBG>
int launch_tc2_k(rl_context* h, const CUtensorMap& mw, const T2Args& a) {
  const size_t smem = T2Smem<KP>::total;
  const int grid = a.n_user_tiles < 2 * h->sm_count ? (a.n_user_tiles + 1) / 2 : h->sm_count;
  auto kern = score_tc2_kernel<KP, NT, DUMP, DBG>;
  RL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<grid, T2_THREADS, smem, h->stream>>>(mw, a);
  RL_LAUNCH_CHECK(h, "score_tc2_kernel");
  return RL_OK;
}

template <int KP, bool DUMP>
int launch_tc2_n(rl_context* h, const CUtensorMap& mw, const T2Args& a) {
  if (!DUMP && a.dbg && a.n <= 10) return launch_tc2_k<KP, 10, false, true>(h, mw, a);   // RL_TC_DEBUG: counters, RL_TC2_FLAGS
  if (a.n <= 10) return launch_tc2_k<KP, 10, DUMP, false>(h, mw, a);
  return launch_tc2_k<KP, 16, DUMP, false>(h, mw, a);
}

}  // namespace

bool score_tc2_supported(int kp, int n, int64_t n_users, int64_t n_items, const float* ubias, const float* ibias, double gbias) {
  return (kp == 16 || kp == 32 || kp == 64) && n <= 16 && n_items >= 1024 && n_items < (INT64_C(1) << 30) && n_users >= 1 && !ubias &&
         !ibias && gbias == 0.0;
}

// Centred single-product candidate pass + exact re-scoring.  Rows whose candidate list overflows are appended to the
// overflow list (defer, or the call's own list which is then finished by the exact kernel).
// dump != nullptr: debug mode, writes the raw accumulators (n_users x dump_ld, positions in class order), the row scales
// (dump_su), the position -> id map (dump_perm, dump_ld entries) and the class table (dump_tab, 64 floats/ints).
int launch_score_tc2(rl_context* h, int kp, const ScoreArgs& sa, const ScoreDefer* defer, float* dump, int64_t dump_ld, float* dump_su,
                     int32_t* dump_perm, void* dump_tab) {
  const int kph = 64;
  const int64_t n_items = sa.n_items;
  const int64_t n_pos = ((n_items + T2_BN - 1) / T2_BN + T2_NCLS) * T2_BN;
  const int nblk = static_cast<int>((n_items + PP_ROWS - 1) / PP_ROWS);
  const int csum_blocks = 64;
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  // scratch layout
  size_t off = 0;
  const size_t o_tab = off; off += al(sizeof(ClassTable));
  const size_t o_cnt = off; off += al(256);                                            // overflow count (+ debug counters)
  const size_t o_dbg = off; off += al(160 * sizeof(long long));
  const size_t o_vbar = off; off += al(sizeof(float) * 128);
  const size_t o_part = off; off += al(sizeof(double) * csum_blocks * kp);
  const size_t o_wn = off; off += al(sizeof(float) * n_items);
  const size_t o_cls = off; off += al(n_items);
  const size_t o_bc = off; off += al(sizeof(int) * nblk * T2_NCLS);
  const size_t o_perm = off; off += al(sizeof(int32_t) * n_pos);
  const size_t o_inv = off; off += al(sizeof(int32_t) * n_items);
  const size_t o_whi = off; off += al(sizeof(__half) * n_pos * kph);
  const size_t o_orow = off; off += al(sizeof(int32_t) * sa.n_users);
  const size_t o_ccnt = off; off += al(sizeof(int32_t) * sa.n_users);
  const size_t o_cids = off; off += al(sizeof(int32_t) * sa.n_users * T2_CAND);
  // seen positions: as many entries as the pattern has (indptr[n_users]: read from the device once per pattern and handle)
  int64_t seen_cap = 0;
  if (sa.seen_indptr && !dump) {
    bool hit = false;
    for (const auto& e : h->seen_nnz_cache)
      if (e.indptr == sa.seen_indptr && e.n_rows == sa.n_users) { seen_cap = e.nnz; hit = true; break; }
    if (!hit) {
      RL_CUDA(h, cudaMemcpyAsync(&seen_cap, sa.seen_indptr + sa.n_users, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
      RL_CUDA(h, cudaStreamSynchronize(h->stream));
      if (h->seen_nnz_cache.size() >= 64) h->seen_nnz_cache.clear();
      h->seen_nnz_cache.push_back({sa.seen_indptr, sa.n_users, seen_cap});
    }
  }
  const size_t o_spos = off; off += al(sizeof(int32_t) * static_cast<size_t>(seen_cap));
  const unsigned int pool_blocks = static_cast<unsigned int>(sa.n_users / 8 + 4096);
  const size_t o_pool = off; off += al(sizeof(int32_t) * static_cast<size_t>(pool_blocks) * T2_CAND);
  const size_t o_srow = off; off += al(sizeof(int32_t) * sa.n_users);
  const size_t o_slist = off; off += al(sizeof(int32_t) * T2_SAMPLE);
  const size_t o_sidx = off; off += al(sizeof(int32_t) * T2_SAMPLE * 16);
  void* sc = nullptr;
  int rc = scratch(h, off, &sc);
  if (rc != RL_OK) return rc;
  char* scb = static_cast<char*>(sc);
  ClassTable* tab = reinterpret_cast<ClassTable*>(scb + o_tab);
  unsigned int* ocount = reinterpret_cast<unsigned int*>(scb + o_cnt);
  long long* dbg = reinterpret_cast<long long*>(scb + o_dbg);
  float* vbar = reinterpret_cast<float*>(scb + o_vbar);
  double* partial = reinterpret_cast<double*>(scb + o_part);
  float* wnorm = reinterpret_cast<float*>(scb + o_wn);
  unsigned char* cls = reinterpret_cast<unsigned char*>(scb + o_cls);
  int* bc = reinterpret_cast<int*>(scb + o_bc);
  int32_t* perm = reinterpret_cast<int32_t*>(scb + o_perm);
  int32_t* inv = reinterpret_cast<int32_t*>(scb + o_inv);
  int32_t* spos = reinterpret_cast<int32_t*>(scb + o_spos);
  __half* whi = reinterpret_cast<__half*>(scb + o_whi);
  int32_t* orows = reinterpret_cast<int32_t*>(scb + o_orow);
  int32_t* ccnt = reinterpret_cast<int32_t*>(scb + o_ccnt);
  int32_t* cids = reinterpret_cast<int32_t*>(scb + o_cids);

  if (!dump && !(h->preloaded_exact & (1u << (kp >> 4)))) {
    // first use of the fallback (module load, attribute set-up, pool growth) must not land inside a later, timed call
    h->preloaded_exact |= 1u << (kp >> 4);
    RL_CUDA(h, cudaMemsetAsync(orows, 0, sizeof(int32_t), h->stream));
    ScoreArgs ea = sa;
    ea.user_list = orows;
    ea.n_list = 1;
    ea.scatter = true;
    if ((rc = launch_score_exact(h, kp, ea)) != RL_OK) return rc;
  }
  RL_CUDA(h, cudaMemsetAsync(scb + o_cnt, 0, 256 + al(160 * sizeof(long long)), h->stream));
  // the V-dependent part (everything up to the class table) is reused between the chunks of one host-wrapper call
  const bool reuse = !dump && h->tc2_prep_state == 2 && h->tc2_prep_v == sa.vf && h->tc2_prep_scratch == sc &&
                     h->tc2_prep_scratch_bytes == h->scratch_bytes && h->tc2_prep_items == n_items && h->tc2_prep_kp == kp;
  if (!reuse) {
    RL_CUDA(h, cudaMemsetAsync(perm, 0xFF, sizeof(int32_t) * n_pos, h->stream));
    RL_CUDA(h, cudaMemsetAsync(whi, 0, sizeof(__half) * n_pos * kph, h->stream));
    colsum_partial_kernel<<<csum_blocks, 256, 0, h->stream>>>(sa.vf, n_items, kp, partial);
    RL_LAUNCH_CHECK(h, "colsum_partial_kernel");
    colsum_final_kernel<<<1, 128, 0, h->stream>>>(partial, csum_blocks, kp, n_items, vbar, tab);
    RL_LAUNCH_CHECK(h, "colsum_final_kernel");
    {
      int64_t g = (n_items * (kp / 4) + 255) / 256;
      if (g > 8 * h->sm_count) g = 8 * h->sm_count;
      centre_norm_kernel<<<static_cast<unsigned>(g), 256, 0, h->stream>>>(sa.vf, n_items, kp, vbar, wnorm, tab);
      RL_LAUNCH_CHECK(h, "centre_norm_kernel");
    }
    classify_kernel<<<nblk, PP_ROWS, 0, h->stream>>>(wnorm, n_items, tab, cls, bc);
    RL_LAUNCH_CHECK(h, "classify_kernel");
    // exact top-n of a few evenly spaced users: where the winners come from decides the scan order of the classes
    const int n_sample = static_cast<int>(sa.n_users < T2_SAMPLE ? sa.n_users : T2_SAMPLE);
    int32_t* slist = reinterpret_cast<int32_t*>(scb + o_slist);
    int32_t* sidx = reinterpret_cast<int32_t*>(scb + o_sidx);
    {
      sample_users_kernel<<<1, T2_SAMPLE, 0, h->stream>>>(sa.n_users, n_sample, slist);
      RL_LAUNCH_CHECK(h, "sample_users_kernel");
      ScoreArgs ea = sa;
      ea.user_list = slist;
      ea.n_list = n_sample;
      ea.scatter = false;
      ea.idx_out = sidx;
      ea.score_out = nullptr;
      if ((rc = launch_score_exact(h, kp, ea)) != RL_OK) return rc;
    }
    class_scan_kernel<<<1, 32, 0, h->stream>>>(bc, nblk, tab, sidx, n_sample * sa.n, wnorm, n_items);
    RL_LAUNCH_CHECK(h, "class_scan_kernel");
    place_kernel<64><<<nblk, PP_ROWS, 0, h->stream>>>(sa.vf, n_items, kp, vbar, wnorm, cls, bc, tab, perm, inv, whi);
    RL_LAUNCH_CHECK(h, "place_kernel");
    class_finish_kernel<<<1, 32, 0, h->stream>>>(tab);
    RL_LAUNCH_CHECK(h, "class_finish_kernel");
    if (!dump && h->tc2_prep_state >= 1) {
      h->tc2_prep_state = 2;
      h->tc2_prep_v = sa.vf; h->tc2_prep_scratch = sc; h->tc2_prep_scratch_bytes = h->scratch_bytes;
      h->tc2_prep_items = n_items; h->tc2_prep_kp = kp;
    }
  }
  if (sa.seen_indptr && !dump && seen_cap > 0) {
    int64_t g = (sa.n_users + 7) / 8;
    if (g > 16 * h->sm_count) g = 16 * h->sm_count;
    seen_positions_kernel<<<static_cast<unsigned>(g), 256, 0, h->stream>>>(sa.seen_indptr, sa.seen_indices, sa.n_users, seen_cap, inv, cls, spos);
    RL_LAUNCH_CHECK(h, "seen_positions_kernel");
  }

  CUtensorMap mw;
  if ((rc = make_map2(h, &mw, whi, n_pos, kph)) != RL_OK) return rc;
  T2Args a{};
  a.uf = sa.uf; a.ldf = kp; a.n_users = sa.n_users; a.seen_indptr = dump ? nullptr : sa.seen_indptr; a.n = sa.n;
  a.seen_pos = spos; a.seen_cap = seen_cap;
  a.tab = tab; a.cand_ids = cids; a.cand_cnt = ccnt;
  unsigned int* pool_next = ocount + 4;
  unsigned int* spill_count = ocount + 8;
  int32_t* pool = reinterpret_cast<int32_t*>(scb + o_pool);
  int32_t* srows = reinterpret_cast<int32_t*>(scb + o_srow);
  a.pool = pool; a.pool_next = pool_next; a.pool_blocks = pool_blocks;
  a.acc_rel = static_cast<float>(kph) * 1.1920929e-7f;      // k * 2^-23
  a.dump = dump; a.dump_su = dump_su; a.dump_ld = dump_ld;
  a.n_user_tiles = static_cast<int>((sa.n_users + T2_BM - 1) / T2_BM);
  const bool want_dbg = getenv("RL_TC_DEBUG") != nullptr;
  a.dbg = want_dbg ? dbg : nullptr;
  a.dbg_flags = getenv("RL_TC2_FLAGS") ? atoi(getenv("RL_TC2_FLAGS")) : 0;
  if (dump) {
    if ((rc = launch_tc2_n<64, true>(h, mw, a)) != RL_OK) return rc;
    if (dump_perm) RL_CUDA(h, cudaMemcpyAsync(dump_perm, perm, sizeof(int32_t) * (dump_ld < n_pos ? dump_ld : n_pos), cudaMemcpyDeviceToDevice, h->stream));
    if (dump_tab) RL_CUDA(h, cudaMemcpyAsync(dump_tab, tab, 256, cudaMemcpyDeviceToDevice, h->stream));
    return RL_OK;
  }
  if ((rc = launch_tc2_n<64, false>(h, mw, a)) != RL_OK) return rc;
  int32_t* o_rows = defer ? defer->rows : orows;
  unsigned int* o_count = defer ? defer->count : ocount;
  const int64_t row_offset = defer ? defer->row_offset : 0;
  {
    const unsigned g = static_cast<unsigned>((sa.n_users + 7) / 8);
    const unsigned gs = static_cast<unsigned>(2 * h->sm_count);
#define RL_RESCORE(KPV)                                                                                                                        \
  rescore_kernel<KPV><<<g, 256, 0, h->stream>>>(sa.uf, sa.vf, sa.n_users, sa.n, cids, ccnt, perm, sa.idx_out, sa.score_out, o_rows, o_count, \
                                                row_offset, srows, spill_count);                                                             \
  rescore_spill_kernel<KPV><<<gs, 256, 0, h->stream>>>(sa.uf, sa.vf, sa.n, cids, pool, perm, sa.idx_out, sa.score_out, srows, spill_count)
    switch (kp) {
      case 16: RL_RESCORE(16); break;
      case 32: RL_RESCORE(32); break;
      default: RL_RESCORE(64); break;
    }
#undef RL_RESCORE
    h->launches++;
    RL_LAUNCH_CHECK(h, "rescore_kernel");
  }
  if (want_dbg) {
    long long d[160];
    RL_CUDA(h, cudaMemcpyAsync(d, dbg, sizeof(d), cudaMemcpyDeviceToHost, h->stream));
    RL_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 2; ++i)
      fprintf(stderr, "score_tc2 dbg (block 0, cycles): mma issuer %d total %lld wait_A %lld wait_Tempty %lld wait_Bfull %lld\n", i, d[4 * i], d[4 * i + 1],
              d[4 * i + 2], d[4 * i + 3]);
    for (int w = 0; w < 16; w += 3)
      fprintf(stderr, "  scan warp %d: total %lld wait_Tfull %lld wait_ld %lld scan %lld\n", w, d[8 + 4 * w], d[9 + 4 * w], d[10 + 4 * w],
              d[11 + 4 * w]);
    for (int w = 0; w < 8; w += 2)
      fprintf(stderr, "  sel warp %d: total %lld idle %lld records(lane0) %lld compactions %lld passes %lld iterations %lld in-pass %lld refills %lld\n", w,
              d[80 + 8 * w], d[81 + 8 * w], d[82 + 8 * w], d[83 + 8 * w], d[84 + 8 * w], d[85 + 8 * w], d[86 + 8 * w], d[87 + 8 * w]);
  }
  if (defer) return RL_OK;
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
