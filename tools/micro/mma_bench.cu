// mma_bench.cu -- what one SM sustains: tcgen05.mma groups (A from TMEM, B from shared memory, fp16, M = 128) issued
// by one thread, with W warps reading the accumulator region back with tcgen05.ld at the same time.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I recsys_lite_b200/csrc -o tools/micro/mma_bench tools/micro/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tcgen05_ptx.cuh"
using namespace rl::tc;

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}

// PER MMAs of K = 16 per commit group; the groups rotate over the accumulator regions
template <int N, int PER>
__global__ void __launch_bounds__(640, 1) k(int groups, int ld_warps, int ld_mode, int issuers, long long* out, uint32_t* sink) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  __shared__ int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar[0]), issuers);
    for (int i = 1; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  constexpr uint32_t IDESC = (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
  constexpr int NACC = N == 256 ? 1 : (N == 128 ? 3 : 6);
  if (warp < issuers && lane == 0) {
    const uint64_t bd0 = make_desc(base);
    const long long t0 = clock64();
    for (int g = warp; g < groups; g += issuers) {
      const uint32_t d = tb + 128 + (g % NACC) * N;
#pragma unroll
      for (int kk = 0; kk < PER; ++kk)
        tc_mma_f16_ts(d, tb + (kk & 3) * 8, bd0 + static_cast<uint64_t>(((kk & 3) * 32) >> 4), IDESC, kk != 0 ? 1u : 0u);
      if (g + issuers < groups) tc_commit(smem_u32(&bar[1 + warp]));
    }
    const long long t1 = clock64();
    tc_commit(smem_u32(&bar[0]));           // completes when everything this thread issued has finished
    if (warp == 0) {
      mbar_wait(smem_u32(&bar[0]), 0);
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
      stop = 1;
    }
  } else if (warp >= 4 && warp < 4 + ld_warps) {
    const uint32_t trow = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 128;
    uint32_t acc = 0;
    long long n = 0;
    const long long t0 = clock64();
    while (!*reinterpret_cast<volatile int*>(&stop)) {
      if (ld_mode == 32) {
        uint32_t v[32];
        const uint32_t col = static_cast<uint32_t>(((warp >> 2) * 4 + (n & 3)) * 32) % 384u;
        tc_ld32(trow + col, reinterpret_cast<float(&)[32]>(v));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
        n += 32;
      } else {
        uint32_t v[16], w[16];
        const uint32_t col = static_cast<uint32_t>(((warp >> 2) * 4 + (n & 3)) * 32) % 384u;
        ld16(trow + col, v);
        ld16(trow + col + 16, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= v[i] ^ w[i];
        n += 32;
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[4 + (warp - 4) * 2] = n; out[5 + (warp - 4) * 2] = t1 - t0; }
    if (acc == 0x12345u) sink[threadIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

template <int N, int PER>
void run(int groups, int ld_warps, int ld_mode, int issuers, long long* out, uint32_t* sink) {
  cudaFuncSetAttribute(k<N, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  cudaMemset(out, 0, 64 * 8);
  for (int it = 0; it < 2; ++it) {
    k<N, PER><<<148, 640, 80 * 1024>>>(groups, ld_warps, ld_mode, issuers, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long c[64];
  cudaMemcpy(c, out, sizeof(c), cudaMemcpyDeviceToHost);
  double cols = 0, clk = 1;
  for (int w = 0; w < ld_warps; ++w) { cols += (double)c[4 + 2 * w]; clk = (double)c[5 + 2 * w]; }
  printf("N=%3d %2d MMAs/commit issuers %d | ld warps %2d (x%d): MMA %6.1f clk each (issue %6.1f), tensor work %5.1f%% of 8192 flop/clk | LDTM %6.1f B/clk/SM\n", N, PER,
         issuers, ld_warps, ld_mode, (double)c[1] / groups / PER, (double)c[0] / groups / PER, 100.0 * (2.0 * 128 * N * 16) * groups * PER / c[1] / 8192.0,
         cols * 32 * 4 / clk);
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 4096);
  const int G = 20000;
  run<128, 4>(G, 0, 32, 1, out, sink);
  run<128, 12>(G, 0, 32, 1, out, sink);
  run<128, 4>(G, 0, 32, 2, out, sink);
  run<256, 4>(G, 0, 32, 1, out, sink);
  run<256, 8>(G, 0, 32, 1, out, sink);
  run<64, 4>(G, 0, 32, 1, out, sink);
  for (int w : {4, 8, 16}) {
    run<128, 4>(G, w, 32, 1, out, sink);
    run<128, 4>(G, w, 16, 1, out, sink);
    run<256, 4>(G, w, 32, 1, out, sink);
    run<256, 4>(G, w, 16, 1, out, sink);
  }
  run<128, 4>(G, 16, 16, 2, out, sink);
  return 0;
}
