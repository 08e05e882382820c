// tmem_ld_bench.cu -- how fast can the epilogue warps of one SM read TMEM (tcgen05.ld.32x32b.x32)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/tmem_ld_bench tools/micro/tmem_ld_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tc_ld32_pack16(uint32_t taddr, uint32_t (&r)[32]) {
  // 64 columns of 16-bit data (low halves of the 32-bit cells), packed two per register
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.pack::16b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int active_warps, int reps, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = slot;
  const int q = warp & 3;
  const uint32_t trow = tb + ((uint32_t)(q * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < active_warps) {
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t col = (uint32_t)(((warp >> 2) * 4 + c) * 32) & 511u;
        if (MODE == 2) {
          if (c & 1) continue;
          uint32_t v[32], w[32];
          tc_ld32(trow + col, v);
          tc_ld32(trow + ((col + 32) & 511u), w);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= v[i] ^ w[i];
        } else if (MODE == 0) {
          uint32_t v[32];
          tc_ld32(trow + col, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= v[i];
        } else {
          uint32_t v[32];
          tc_ld32_pack16(trow + ((col * 2) & 511u), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= v[i];
        }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4096);
  const int reps = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int aw : {1, 4, 8, 16}) {
      for (int it = 0; it < 2; ++it) {
        if (mode == 0) k<0><<<148, 512>>>(aw, reps, out, sink); else if (mode == 1) k<1><<<148, 512>>>(aw, reps, out, sink); else k<2><<<148, 512>>>(aw, reps, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      const double cols = (double)aw * reps * 4 * 32 * (mode == 1 ? 2 : 1);           // 32-lane x 32-column pieces
      const double bytes = cols * 32 * 4;                        // as 32-bit TMEM cells
      printf("mode %s warps %2d: %lld clk, %.1f clk per x32 load per warp, %.1f B/clk/SM (32-bit cells)\n", mode == 1 ? "pack16" : mode == 2 ? "b32x2" : "b32", aw, c,
             (double)c / (reps * 4), bytes / c);
    }
  return 0;
}
