// api.cu -- the C ABI of librecsys_b200.so (include/recsys_b200.h): handle, memory helpers, host-buffer wrappers.
#include <cuda.h>
#include <cstdarg>
#include <algorithm>
#include <thread>
#include <vector>

#include <cmath>

#include "common.cuh"

namespace rl {

thread_local std::string g_last_error;

int fail(rl_context* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (h) h->err = buf;
  return code;
}

int cuda_fail(rl_context* h, cudaError_t e, const char* what) {
  const int code = (e == cudaErrorMemoryAllocation) ? RL_ERR_NOMEM : RL_ERR_CUDA;
  return fail(h, code, "CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
}

int scratch(rl_context* h, size_t bytes, void** out) {
  if (bytes > h->scratch_bytes) {
    if (h->scratch) {
      RL_CUDA(h, cudaStreamSynchronize(h->stream));
      RL_CUDA(h, cudaFree(h->scratch));
      h->scratch = nullptr;
      h->scratch_bytes = 0;
    }
    const size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    RL_CUDA(h, cudaMalloc(&h->scratch, want));
    h->scratch_bytes = want;
  }
  *out = h->scratch;
  return RL_OK;
}

namespace {

// (rows x k, stride src_ld) -> (rows x dst_ld) with zero fill of the padding columns
template <typename TS, typename TD>
__global__ void convert_kernel(const TS* __restrict__ src, int64_t rows, int k, int src_ld, TD* __restrict__ dst, int dst_ld) {
  const int64_t total = rows * dst_ld;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < total;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = p / dst_ld;
    const int c = static_cast<int>(p - r * dst_ld);
    dst[p] = (c < k) ? static_cast<TD>(src[r * src_ld + c]) : static_cast<TD>(0);
  }
}

// rows without observations keep their initial values (reference algo.py:316-317): copied from the uploaded initial
// matrix after the half-step, so that the upload itself can run underneath the half-step
__global__ void __launch_bounds__(256) fill_empty_rows_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, int kp,
                                                              const float* __restrict__ init, float* __restrict__ out) {
  const int per = kp / 4;
  const int64_t total = n_rows * per;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = p / per;
    if (indptr[r + 1] == indptr[r])
      reinterpret_cast<float4*>(out)[p] = reinterpret_cast<const float4*>(init)[p];
  }
}

// RAII device allocation for the host wrappers.  Stream-ordered (cudaMallocAsync on the handle's stream): the device's
// default memory pool keeps freed blocks (release threshold set in rl_create), so repeated host-buffer calls neither
// pay cudaMalloc / cudaFree nor synchronise the device.
struct DevBuf {
  void* p = nullptr;
  rl_context* ctx = nullptr;
  // The free is ordered on the compute stream; transfers of the host wrappers run on the copy stream.  On the regular
  // paths the compute stream has already waited for them; on an early error return it has not, so the free itself is made
  // to wait for whatever the copy stream still has in flight.
  ~DevBuf() {
    if (!p) return;
    if (ctx->copy_stream && ctx->copy_stream != ctx->stream) {
      if (!ctx->free_guard) cudaEventCreateWithFlags(&ctx->free_guard, cudaEventDisableTiming);
      if (ctx->free_guard && cudaEventRecord(ctx->free_guard, ctx->copy_stream) == cudaSuccess)
        cudaStreamWaitEvent(ctx->stream, ctx->free_guard, 0);
    }
    cudaFreeAsync(p, ctx->stream);
  }
  int alloc(rl_context* h, size_t bytes) {
    ctx = h;
    RL_CUDA(h, cudaMallocAsync(&p, bytes ? bytes : 16, h->stream));
    return RL_OK;
  }
  template <typename T> T* as() { return static_cast<T*>(p); }
  void* release() { void* q = p; p = nullptr; return q; }      // the caller of the C ABI owns it from here on
};

#define RL_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != RL_OK) return _rc; \
  } while (0)

int upload(rl_context* h, DevBuf& b, const void* host, size_t bytes) {
  RL_TRY(b.alloc(h, bytes));
  if (bytes) RL_CUDA(h, cudaMemcpyAsync(b.p, host, bytes, cudaMemcpyHostToDevice, h->stream));
  return RL_OK;
}

// run a piece of the host wrappers on another stream of the handle (launchers enqueue on h->stream)
struct StreamScope {
  rl_context* h;
  cudaStream_t saved;
  StreamScope(rl_context* ctx, cudaStream_t s) : h(ctx), saved(ctx->stream) { h->stream = s; }
  ~StreamScope() { h->stream = saved; }
};

// the chunks of one host-wrapper call score against the same V: its preprocessing (score_tc2.cu) is done once
struct PrepScope {
  rl_context* h;
  explicit PrepScope(rl_context* ctx) : h(ctx) { h->tc2_prep_state = 1; }
  ~PrepScope() { h->tc2_prep_state = 0; }
};

struct Event {  // ordering between the handle's compute and copy streams
  cudaEvent_t e = nullptr;
  ~Event() { if (e) cudaEventDestroy(e); }
  int record(rl_context* h, cudaStream_t s) {
    if (!e) RL_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    RL_CUDA(h, cudaEventRecord(e, s));
    return RL_OK;
  }
  int wait(rl_context* h, cudaStream_t s) {
    RL_CUDA(h, cudaStreamWaitEvent(s, e, 0));
    return RL_OK;
  }
};

}  // namespace

int launch_factors_convert(rl_context* h, const void* src, int src_dtype, int64_t rows, int k, int src_ld,
                           void* dst, int dst_dtype, int dst_ld);

namespace {

// host (rows x k, f32 / f64) -> device (rows x padded k, f32) on h->stream, no synchronisation.  A float32 matrix whose
// k is already the padded width is copied straight into place; everything else goes through a staging buffer.
int factors_h2d(rl_context* h, const void* src, int dtype, int64_t rows, int k, float* dst, DevBuf& stage) {
  if (rows == 0) return RL_OK;
  const int kp = rl::padded_k(k);
  const size_t bytes = static_cast<size_t>(rows) * k * (dtype == RL_F64 ? 8 : 4);
  if (dtype == RL_F32 && k == kp) {
    RL_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return RL_OK;
  }
  RL_TRY(stage.alloc(h, bytes));
  RL_CUDA(h, cudaMemcpyAsync(stage.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return launch_factors_convert(h, stage.p, dtype, rows, k, k, dst, RL_F32, kp);
}

int factors_d2h(rl_context* h, const float* src, int64_t rows, int k, void* dst, int dtype, DevBuf& stage) {
  if (rows == 0) return RL_OK;
  const int kp = rl::padded_k(k);
  const size_t bytes = static_cast<size_t>(rows) * k * (dtype == RL_F64 ? 8 : 4);
  if (dtype == RL_F32 && k == kp) {
    RL_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    return RL_OK;
  }
  RL_TRY(stage.alloc(h, bytes));
  RL_TRY(launch_factors_convert(h, src, RL_F32, rows, k, kp, stage.p, dtype, k));
  RL_CUDA(h, cudaMemcpyAsync(dst, stage.p, bytes, cudaMemcpyDeviceToHost, h->stream));
  return RL_OK;
}

}  // namespace

int launch_factors_convert(rl_context* h, const void* src, int src_dtype, int64_t rows, int k, int src_ld,
                           void* dst, int dst_dtype, int dst_ld) {
  if (rows == 0) return RL_OK;
  const int64_t total = rows * dst_ld;
  int64_t grid = (total + 255) / 256;
  if (grid > 16 * h->sm_count) grid = 16 * h->sm_count;
  if (src_dtype == RL_F64 && dst_dtype == RL_F32)
    convert_kernel<double, float><<<grid, 256, 0, h->stream>>>(static_cast<const double*>(src), rows, k, src_ld, static_cast<float*>(dst), dst_ld);
  else if (src_dtype == RL_F32 && dst_dtype == RL_F32)
    convert_kernel<float, float><<<grid, 256, 0, h->stream>>>(static_cast<const float*>(src), rows, k, src_ld, static_cast<float*>(dst), dst_ld);
  else if (src_dtype == RL_F32 && dst_dtype == RL_F64)
    convert_kernel<float, double><<<grid, 256, 0, h->stream>>>(static_cast<const float*>(src), rows, k, src_ld, static_cast<double*>(dst), dst_ld);
  else
    return fail(h, RL_ERR_INVALID, "factors_convert: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
  RL_LAUNCH_CHECK(h, "convert_kernel");
  return RL_OK;
}

}  // namespace rl

using namespace rl;

#define RL_NEED(h)                                                     \
  do {                                                                 \
    if (!(h)) return rl::fail(nullptr, RL_ERR_INVALID, "null handle"); \
    cudaError_t _e = cudaSetDevice((h)->device);                       \
    if (_e != cudaSuccess) return rl::cuda_fail((h), _e, "cudaSetDevice"); \
  } while (0)

extern "C" {

int rl_abi_version(void) { return RL_ABI_VERSION; }

int rl_padded_k(int k) { return rl::padded_k(k); }

const char* rl_last_error(rl_handle h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int rl_create(int device, rl_handle* out) {
  if (!out) return fail(nullptr, RL_ERR_INVALID, "rl_create: out is null");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(nullptr, RL_ERR_NO_DEVICE, "rl_create: no CUDA device (%s); this library has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  if (device >= count) return fail(nullptr, RL_ERR_INVALID, "rl_create: device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(nullptr, RL_ERR_NO_DEVICE, "rl_create: device %d is sm_%d%d; kernels are built for sm_100a only", device,
                prop.major, prop.minor);
  rl_context* h = new rl_context();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->cc_major = prop.major;
  h->cc_minor = prop.minor;
  h->total_mem = prop.totalGlobalMem;
  if ((e = cudaSetDevice(device)) != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaSetDevice"); }
  if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaStreamCreate"); }
  h->stream = h->own_stream;
  if ((e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) {
    cudaStreamDestroy(h->own_stream);
    delete h;
    return cuda_fail(nullptr, e, "cudaStreamCreate(copy)");
  }
  if ((e = cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&h->aux_fork, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&h->aux_join, cudaEventDisableTiming)) != cudaSuccess) {
    cudaStreamDestroy(h->own_stream);
    cudaStreamDestroy(h->copy_stream);
    delete h;
    return cuda_fail(nullptr, e, "cudaStreamCreate(aux)");
  }
  if ((e = cudaMalloc(&h->counter, 64 * sizeof(unsigned int))) != cudaSuccess) {
    cudaStreamDestroy(h->own_stream);
    delete h;
    return cuda_fail(nullptr, e, "cudaMalloc(counter)");
  }
  cudaMemset(h->counter, 0, 64 * sizeof(unsigned int));
  {  // keep the stream-ordered pool's memory between calls (see DevBuf); rl_trim() gives it back
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, h->device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  *out = h;
  return RL_OK;
}

int rl_destroy(rl_handle h) {
  if (!h) return RL_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->scratch) cudaFree(h->scratch);
  if (h->counter) cudaFree(h->counter);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->aux_stream) { cudaStreamSynchronize(h->aux_stream); cudaStreamDestroy(h->aux_stream); }
  if (h->long_list) cudaFree(h->long_list);
  if (h->aux_fork) cudaEventDestroy(h->aux_fork);
  if (h->aux_join) cudaEventDestroy(h->aux_join);
  if (h->free_guard) cudaEventDestroy(h->free_guard);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return RL_OK;
}

int rl_trim(rl_handle h) {
  RL_NEED(h);
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->scratch) {
    RL_CUDA(h, cudaFree(h->scratch));
    h->scratch = nullptr;
    h->scratch_bytes = 0;
  }
  cudaMemPool_t pool;
  RL_CUDA(h, cudaDeviceGetDefaultMemPool(&pool, h->device));
  RL_CUDA(h, cudaMemPoolTrimTo(pool, 0));
  return RL_OK;
}

int rl_set_stream(rl_handle h, void* cuda_stream) {
  RL_NEED(h);
  h->stream = (cuda_stream == RL_OWN_STREAM) ? h->own_stream : static_cast<cudaStream_t>(cuda_stream);
  return RL_OK;
}

int rl_synchronize(rl_handle h) {
  RL_NEED(h);
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  return RL_OK;
}

int rl_device_info(rl_handle h, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem_bytes) {
  RL_NEED(h);
  if (sm_count) *sm_count = h->sm_count;
  if (cc_major) *cc_major = h->cc_major;
  if (cc_minor) *cc_minor = h->cc_minor;
  if (total_mem_bytes) *total_mem_bytes = static_cast<int64_t>(h->total_mem);
  return RL_OK;
}

int64_t rl_launch_count(rl_handle h) { return h ? h->launches : -1; }

int rl_pinned_alloc(int64_t bytes, void** out) {
  if (!out || bytes < 0) return fail(nullptr, RL_ERR_INVALID, "rl_pinned_alloc: bad argument");
  cudaError_t e = cudaMallocHost(out, static_cast<size_t>(bytes ? bytes : 16));
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaMallocHost");
  return RL_OK;
}

int rl_pinned_free(void* p) {
  if (p) cudaFreeHost(p);
  return RL_OK;
}

int rl_dev_alloc(rl_handle h, int64_t bytes, void** out) {
  RL_NEED(h);
  if (!out || bytes < 0) return fail(h, RL_ERR_INVALID, "rl_dev_alloc: bad argument");
  RL_CUDA(h, cudaMalloc(out, static_cast<size_t>(bytes ? bytes : 16)));
  return RL_OK;
}

int rl_dev_free(rl_handle h, void* p) {
  RL_NEED(h);
  if (p) RL_CUDA(h, cudaFree(p));
  return RL_OK;
}

int rl_memcpy_h2d(rl_handle h, void* dst_dev, const void* src_host, int64_t bytes) {
  RL_NEED(h);
  if (bytes > 0) RL_CUDA(h, cudaMemcpyAsync(dst_dev, src_host, static_cast<size_t>(bytes), cudaMemcpyHostToDevice, h->stream));
  return RL_OK;
}

int rl_memcpy_d2h(rl_handle h, void* dst_host, const void* src_dev, int64_t bytes) {
  RL_NEED(h);
  if (bytes > 0) {
    RL_CUDA(h, cudaMemcpyAsync(dst_host, src_dev, static_cast<size_t>(bytes), cudaMemcpyDeviceToHost, h->stream));
    RL_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return RL_OK;
}

int rl_factors_upload(rl_handle h, const void* src_host, int src_dtype, int64_t rows, int k, float* dst_dev) {
  RL_NEED(h);
  const int kp = rl::padded_k(k);
  if (kp == 0) return fail(h, RL_ERR_INVALID, "rl_factors_upload: k=%d outside [1, 128]", k);
  if (src_dtype != RL_F32 && src_dtype != RL_F64) return fail(h, RL_ERR_INVALID, "rl_factors_upload: bad dtype");
  if (rows == 0) return RL_OK;
  const size_t es = src_dtype == RL_F64 ? 8 : 4;
  const size_t bytes = static_cast<size_t>(rows) * k * es;
  void* stage = nullptr;
  RL_TRY(scratch(h, bytes, &stage));
  RL_CUDA(h, cudaMemcpyAsync(stage, src_host, bytes, cudaMemcpyHostToDevice, h->stream));
  return launch_factors_convert(h, stage, src_dtype, rows, k, k, dst_dev, RL_F32, kp);
}

int rl_factors_download(rl_handle h, const float* src_dev, int64_t rows, int k, void* dst_host, int dst_dtype) {
  RL_NEED(h);
  const int kp = rl::padded_k(k);
  if (kp == 0) return fail(h, RL_ERR_INVALID, "rl_factors_download: k=%d outside [1, 128]", k);
  if (dst_dtype != RL_F32 && dst_dtype != RL_F64) return fail(h, RL_ERR_INVALID, "rl_factors_download: bad dtype");
  if (rows == 0) return RL_OK;
  const size_t es = dst_dtype == RL_F64 ? 8 : 4;
  const size_t bytes = static_cast<size_t>(rows) * k * es;
  void* stage = nullptr;
  RL_TRY(scratch(h, bytes, &stage));
  RL_TRY(launch_factors_convert(h, src_dev, RL_F32, rows, k, kp, stage, dst_dtype, k));
  RL_CUDA(h, cudaMemcpyAsync(dst_host, stage, bytes, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  return RL_OK;
}

int rl_als_half_step_dev(rl_handle h, int64_t n_rows, int64_t n_other, int k, const int64_t* indptr_dev,
                         const int32_t* indices_dev, const float* conf_dev, const float* pref_dev,
                         const float* other_dev, float* mine_dev, double lambda, double w0, const double* gram0_dev,
                         int precision, int ir_steps, const int32_t* row_order_dev) {
  RL_NEED(h);
  return launch_als_half_step(h, n_rows, n_other, k, indptr_dev, indices_dev, conf_dev, pref_dev, other_dev, mine_dev,
                              lambda, w0, gram0_dev, precision, ir_steps, row_order_dev);
}

int rl_als_half_step_peers_dev(rl_handle h, int64_t n_rows, int64_t n_other, int k, const int64_t* indptr_dev,
                               const int32_t* indices_dev, const float* conf_dev, const float* pref_dev,
                               const float* other_dev, float* mine_dev, double lambda, double w0, const double* gram0_dev,
                               int precision, int ir_steps, const int32_t* row_order_dev, int n_peers,
                               float* const* peer_mine_dev) {
  RL_NEED(h);
  return launch_als_half_step(h, n_rows, n_other, k, indptr_dev, indices_dev, conf_dev, pref_dev, other_dev, mine_dev,
                              lambda, w0, gram0_dev, precision, ir_steps, row_order_dev, n_peers, peer_mine_dev);
}

int rl_ipc_export(rl_handle h, const void* dev_ptr, void* handle_out, int64_t* offset_out) {
  RL_NEED(h);
  if (!dev_ptr || !handle_out || !offset_out) return fail(h, RL_ERR_INVALID, "rl_ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == RL_IPC_HANDLE_BYTES, "IPC handle size");
  // driver entry point looked up at run time: the library itself does not link against libcuda
  typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static GetRangeFn get_range = nullptr;
  if (!get_range) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return fail(h, RL_ERR_CUDA, "rl_ipc_export: cuMemGetAddressRange not available from the driver");
    get_range = reinterpret_cast<GetRangeFn>(fp);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (get_range(&base, &size, reinterpret_cast<CUdeviceptr>(dev_ptr)) != CUDA_SUCCESS)
    return fail(h, RL_ERR_CUDA, "rl_ipc_export: cuMemGetAddressRange failed (not a device allocation?)");
  cudaIpcMemHandle_t hd;
  RL_CUDA(h, cudaIpcGetMemHandle(&hd, reinterpret_cast<void*>(base)));
  memcpy(handle_out, &hd, sizeof(hd));
  *offset_out = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(dev_ptr) - base);
  return RL_OK;
}

int rl_ipc_open(rl_handle h, const void* handle, int64_t offset, void** dev_ptr_out) {
  RL_NEED(h);
  if (!handle || !dev_ptr_out || offset < 0) return fail(h, RL_ERR_INVALID, "rl_ipc_open: bad argument");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle, sizeof(hd));
  void* base = nullptr;
  RL_CUDA(h, cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
  h->ipc_bases.push_back({static_cast<char*>(base) + offset, base});
  *dev_ptr_out = static_cast<char*>(base) + offset;
  return RL_OK;
}

int rl_ipc_close(rl_handle h, void* dev_ptr) {
  RL_NEED(h);
  for (size_t i = 0; i < h->ipc_bases.size(); ++i) {
    if (h->ipc_bases[i].first == dev_ptr) {
      void* base = h->ipc_bases[i].second;
      h->ipc_bases.erase(h->ipc_bases.begin() + i);
      RL_CUDA(h, cudaStreamSynchronize(h->stream));
      RL_CUDA(h, cudaIpcCloseMemHandle(base));
      return RL_OK;
    }
  }
  return fail(h, RL_ERR_INVALID, "rl_ipc_close: pointer was not returned by rl_ipc_open");
}

int rl_csr_transpose_dev(rl_handle h, int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr_dev,
                         const int32_t* indices_dev, int64_t* indptr_t_dev, int32_t* indices_t_dev) {
  RL_NEED(h);
  return launch_csr_transpose(h, n_rows, n_cols, nnz, indptr_dev, indices_dev, indptr_t_dev, indices_t_dev);
}

int rl_gram_dev(rl_handle h, const float* y_dev, int64_t n, int k, double* gram_dev) {
  RL_NEED(h);
  return launch_gram(h, y_dev, n, k, gram_dev);
}

namespace {
// keep == nullptr: fitted factors go back to user_f / item_f.  keep != nullptr: nothing is downloaded, the six device
// buffers of the model and its pattern are handed to the caller.
int als_fit_host_impl(rl_handle h, int64_t n_users, int64_t n_items, int k, const int64_t* indptr, const int32_t* indices,
                      const int64_t* indptr_t, const int32_t* indices_t, int iterations, double lambda, int precision,
                      int ir_steps, void* user_f, void* item_f, int factor_dtype, void** keep) {
  const int kp = rl::padded_k(k);
  if (kp == 0) return fail(h, RL_ERR_INVALID, "rl_als_fit_host: k=%d outside [1, 128]", k);
  if (n_users < 0 || n_items < 0 || iterations < 0) return fail(h, RL_ERR_INVALID, "rl_als_fit_host: negative size");
  if (!indptr || !user_f || !item_f) return fail(h, RL_ERR_INVALID, "rl_als_fit_host: null pointer");
  if ((indptr_t == nullptr) != (indices_t == nullptr) && indptr[n_users] > 0)
    return fail(h, RL_ERR_INVALID, "rl_als_fit_host: pass both transposed arrays or neither");
  const bool device_transpose = indptr_t == nullptr;     // built on the device underneath the first user half-step
  const int64_t nnz = indptr[n_users];
  if (!device_transpose && nnz != indptr_t[n_items])
    return fail(h, RL_ERR_INVALID, "rl_als_fit_host: CSR and transposed CSR disagree on nnz");
  // Host traffic runs on the copy stream, ordered against the kernels by events:
  //   copy:    V, user CSR | U (initial), item CSR ................... | U back (during the last item half-step)
  //   compute:              | user half-step, empty rows <- initial U | item half-step ... | V back
  // The first user half-step overwrites every row that has observations, so it does not wait for the initial U: that
  // upload runs underneath it into a second buffer, from which only the rows WITHOUT observations are taken afterwards.
  DevBuf d_ip, d_ix, d_tp, d_tx, d_u, d_u0, d_v, st_u, st_v, st_ud, st_vd;
  Event ev_user_ready, ev_u0_ready, ev_item_ready, ev_u_done, ev_copy_done;
  cudaStream_t cs = h->copy_stream, ks = h->stream;
  RL_TRY(d_ip.alloc(h, sizeof(int64_t) * (n_users + 1)));
  RL_TRY(d_ix.alloc(h, sizeof(int32_t) * (nnz ? nnz : 1)));
  RL_TRY(d_tp.alloc(h, sizeof(int64_t) * (n_items + 1)));
  RL_TRY(d_tx.alloc(h, sizeof(int32_t) * (nnz ? nnz : 1)));
  RL_TRY(d_u.alloc(h, sizeof(float) * (n_users ? n_users : 1) * kp));
  RL_TRY(d_v.alloc(h, sizeof(float) * (n_items ? n_items : 1) * kp));
  // The first user half-step overwrites every row that has observations: when every row has some, the initial U is
  // never read and its upload (the largest transfer of the call) is skipped.
  bool any_empty = iterations == 0;
  for (int64_t r = 0; r < n_users && !any_empty; ++r) any_empty = indptr[r + 1] == indptr[r];
  const bool lazy_u = iterations > 0 && n_users > 0;
  float* u_upload = d_u.as<float>();
  if (lazy_u && any_empty) {
    RL_TRY(d_u0.alloc(h, sizeof(float) * n_users * kp));
    u_upload = d_u0.as<float>();
  }
  // Row ranges of the first user half-step: 1/8, 1/8, 1/4, 1/2 of the observations (large problems only), so that it
  // starts after an eighth of the index upload instead of all of it.
  // ... from PAGEABLE host memory (what the Python surface passes) the upload is slower than the half-step and holds the
  // host thread: eight equal parts, each launched as soon as it has gone up.
  int n_parts = 1;
  int64_t part_row[9] = {0, n_users, n_users, n_users, n_users, n_users, n_users, n_users, n_users};
  if (iterations > 0 && nnz >= (int64_t(1) << 22) && n_users >= 4096) {
    cudaPointerAttributes pa{};
    const bool pinned = cudaPointerGetAttributes(&pa, indices) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const double frac4[3] = {0.125, 0.25, 0.5};
    n_parts = pinned ? 4 : 8;
    for (int c = 0; c + 1 < n_parts; ++c) {
      const double f = pinned ? frac4[c] : (c + 1) / 8.0;
      const int64_t target = static_cast<int64_t>(f * static_cast<double>(nnz));
      part_row[c + 1] = std::lower_bound(indptr, indptr + n_users + 1, target) - indptr;
      if (part_row[c + 1] > n_users) part_row[c + 1] = n_users;
      if (part_row[c + 1] < part_row[c]) part_row[c + 1] = part_row[c];
    }
    part_row[n_parts] = n_users;
  }
  Event ev_part[8];
  Event ev_alloc;
  RL_TRY(ev_alloc.record(h, ks));      // the buffers come from the stream-ordered pool on the compute stream
  RL_TRY(ev_alloc.wait(h, cs));
  {
    StreamScope on_copy(h, cs);
    RL_TRY(factors_h2d(h, item_f, factor_dtype, n_items, k, d_v.as<float>(), st_v));
    RL_CUDA(h, cudaMemcpyAsync(d_ip.p, indptr, sizeof(int64_t) * (n_users + 1), cudaMemcpyHostToDevice, cs));
    for (int c = 0; c < n_parts; ++c) {       // the first user half-step starts on part 0 while the other parts travel
      const int64_t o0 = indptr[part_row[c]], o1 = indptr[part_row[c + 1]];
      if (o1 > o0) RL_CUDA(h, cudaMemcpyAsync(d_ix.as<int32_t>() + o0, indices + o0, sizeof(int32_t) * (o1 - o0), cudaMemcpyHostToDevice, cs));
      RL_TRY(ev_part[c].record(h, cs));
      if (iterations > 0 && n_parts > 1) {
        // launched here, between the copies: a copy from PAGEABLE memory holds the host thread until its data has left, so
        // kernels enqueued after the whole loop would start only when every part has gone up
        RL_TRY(ev_part[c].wait(h, ks));
        const int64_t r0 = part_row[c], r1 = part_row[c + 1];
        if (r1 > r0) {
          StreamScope on_compute(h, ks);
          RL_TRY(launch_als_half_step(h, r1 - r0, n_items, k, d_ip.as<int64_t>() + r0, d_ix.as<int32_t>(), nullptr, nullptr,
                                      d_v.as<float>(), d_u.as<float>() + r0 * kp, lambda, 0.0, nullptr, precision, ir_steps, nullptr));
        }
      }
    }
    RL_TRY(ev_user_ready.record(h, cs));
    if (any_empty) RL_TRY(factors_h2d(h, user_f, factor_dtype, n_users, k, u_upload, st_u));
    RL_TRY(ev_u0_ready.record(h, cs));
    if (!device_transpose) {
      RL_CUDA(h, cudaMemcpyAsync(d_tp.p, indptr_t, sizeof(int64_t) * (n_items + 1), cudaMemcpyHostToDevice, cs));
      if (nnz) RL_CUDA(h, cudaMemcpyAsync(d_tx.p, indices_t, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, cs));
      RL_TRY(ev_item_ready.record(h, cs));
    }
  }
  if (device_transpose) {   // on the second compute stream, beside the first user half-step
    RL_TRY(ev_user_ready.wait(h, h->aux_stream));
    StreamScope on_aux(h, h->aux_stream);
    RL_TRY(launch_csr_transpose(h, n_users, n_items, nnz, d_ip.as<int64_t>(), d_ix.as<int32_t>(), d_tp.as<int64_t>(), d_tx.as<int32_t>()));
    RL_TRY(ev_item_ready.record(h, h->aux_stream));
  }
  if (iterations == 0 || n_parts == 1) RL_TRY(ev_user_ready.wait(h, ks));
  if (!lazy_u) RL_TRY(ev_u0_ready.wait(h, ks));
  bool u_on_copy = false;
  for (int it = 0; it < iterations; ++it) {
    if (it == 0 && n_parts > 1) {
      // (the row ranges of the first user half-step were launched between the uploads of their index parts, above)
    } else {
      RL_TRY(launch_als_half_step(h, n_users, n_items, k, d_ip.as<int64_t>(), d_ix.as<int32_t>(), nullptr, nullptr,
                                  d_v.as<float>(), d_u.as<float>(), lambda, 0.0, nullptr, precision, ir_steps, nullptr));
    }
    if (it == 0) {
      RL_TRY(ev_u0_ready.wait(h, ks));
      int64_t g = (n_users * (kp / 4) + 255) / 256;
      if (g > 8 * h->sm_count) g = 8 * h->sm_count;
      if (g > 0 && any_empty) {        // (no users: nothing to restore, and a grid of 0 blocks is an invalid launch)
        fill_empty_rows_kernel<<<static_cast<unsigned>(g), 256, 0, ks>>>(d_ip.as<int64_t>(), n_users, kp, d_u0.as<float>(), d_u.as<float>());
        RL_LAUNCH_CHECK(h, "fill_empty_rows_kernel");
      }
      RL_TRY(ev_item_ready.wait(h, ks));
    }
    if (it == iterations - 1 && !keep) {        // U is final: send it home while the item half-step runs
      RL_TRY(ev_u_done.record(h, ks));
      RL_TRY(ev_u_done.wait(h, cs));
      StreamScope on_copy(h, cs);
      RL_TRY(factors_d2h(h, d_u.as<float>(), n_users, k, user_f, factor_dtype, st_ud));
      u_on_copy = true;
    }
    RL_TRY(launch_als_half_step(h, n_items, n_users, k, d_tp.as<int64_t>(), d_tx.as<int32_t>(), nullptr, nullptr,
                                d_u.as<float>(), d_v.as<float>(), lambda, 0.0, nullptr, precision, ir_steps, nullptr));
  }
  if (keep) {
    // the host buffers have been read once the copy stream is idle; the kernels go on running on the compute stream
    RL_CUDA(h, cudaStreamSynchronize(cs));
    RL_TRY(ev_copy_done.record(h, h->aux_stream));       // the device-side transposition
    RL_TRY(ev_copy_done.wait(h, ks));
    keep[0] = d_u.release(); keep[1] = d_v.release(); keep[2] = d_ip.release();
    keep[3] = d_ix.release(); keep[4] = d_tp.release(); keep[5] = d_tx.release();
    return RL_OK;
  }
  if (!u_on_copy) RL_TRY(factors_d2h(h, d_u.as<float>(), n_users, k, user_f, factor_dtype, st_ud));
  RL_TRY(factors_d2h(h, d_v.as<float>(), n_items, k, item_f, factor_dtype, st_vd));
  RL_TRY(ev_copy_done.record(h, cs));  // the buffers are freed on the compute stream: it has to see the copy stream's work
  RL_TRY(ev_copy_done.wait(h, ks));
  RL_CUDA(h, cudaStreamSynchronize(ks));
  return RL_OK;
}
}  // namespace

int rl_als_fit_host(rl_handle h, int64_t n_users, int64_t n_items, int k, const int64_t* indptr, const int32_t* indices,
                    const int64_t* indptr_t, const int32_t* indices_t, int iterations, double lambda, int precision,
                    int ir_steps, void* user_f, void* item_f, int factor_dtype) {
  RL_NEED(h);
  return als_fit_host_impl(h, n_users, n_items, k, indptr, indices, indptr_t, indices_t, iterations, lambda, precision, ir_steps,
                           user_f, item_f, factor_dtype, nullptr);
}

int rl_als_fit_resident_host(rl_handle h, int64_t n_users, int64_t n_items, int k, const int64_t* indptr, const int32_t* indices,
                             const int64_t* indptr_t, const int32_t* indices_t, int iterations, double lambda, int precision,
                             int ir_steps, const void* user_f, const void* item_f, int factor_dtype, void** dev_out) {
  RL_NEED(h);
  if (!dev_out) return fail(h, RL_ERR_INVALID, "rl_als_fit_resident_host: null dev_out");
  for (int i = 0; i < 6; ++i) dev_out[i] = nullptr;
  return als_fit_host_impl(h, n_users, n_items, k, indptr, indices, indptr_t, indices_t, iterations, lambda, precision, ir_steps,
                           const_cast<void*>(user_f), const_cast<void*>(item_f), factor_dtype, dev_out);
}

int rl_dev_alloc_async(rl_handle h, int64_t bytes, void** out) {
  RL_NEED(h);
  if (!out || bytes < 0) return fail(h, RL_ERR_INVALID, "rl_dev_alloc_async: bad argument");
  RL_CUDA(h, cudaMallocAsync(out, static_cast<size_t>(bytes ? bytes : 16), h->stream));
  return RL_OK;
}

int rl_dev_free_async(rl_handle h, void* p) {
  RL_NEED(h);
  if (p) RL_CUDA(h, cudaFreeAsync(p, h->stream));
  return RL_OK;
}

int rl_score_topn_dev(rl_handle h, const float* user_f_dev, int64_t n_users, const float* item_f_dev, int64_t n_items,
                      int k, const int64_t* seen_indptr_dev, const int32_t* seen_indices_dev, double global_bias,
                      const float* user_bias_dev, const float* item_bias_dev, int n, int32_t* idx_out_dev,
                      double* score_out_dev, int mode) {
  RL_NEED(h);
  return launch_score_topn(h, user_f_dev, n_users, item_f_dev, n_items, k, seen_indptr_dev, seen_indices_dev, global_bias,
                           user_bias_dev, item_bias_dev, n, idx_out_dev, score_out_dev, mode);
}

int rl_score_dump_dev(rl_handle h, const float* user_f_dev, int64_t n_users, const float* item_f_dev, int64_t n_items, int k,
                      float* out_dev, int64_t ld, int terms) {
  RL_NEED(h);
  return launch_score_dump(h, user_f_dev, n_users, item_f_dev, n_items, k, out_dev, ld, terms);
}

int rl_score_dump_centred_dev(rl_handle h, const float* user_f_dev, int64_t n_users, const float* item_f_dev, int64_t n_items, int k,
                              float* out_dev, int64_t ld, float* row_scale_dev, int32_t* perm_dev, void* table_dev) {
  RL_NEED(h);
  return launch_score_dump_centred(h, user_f_dev, n_users, item_f_dev, n_items, k, out_dev, ld, row_scale_dev, perm_dev, table_dev);
}

int64_t rl_last_overflow_rows(rl_handle h) { return h ? static_cast<int64_t>(h->last_overflow) : -1; }

int rl_recommend_host(rl_handle h, const void* user_f, const void* item_f, int factor_dtype, int64_t n_users,
                      int64_t n_items, int k, const int64_t* seen_indptr, const int32_t* seen_indices, double global_bias,
                      const float* user_bias, const float* item_bias, int n, int32_t* idx_out, double* score_out, int mode) {
  RL_NEED(h);
  const int kp = rl::padded_k(k);
  if (kp == 0) return fail(h, RL_ERR_INVALID, "rl_recommend_host: k=%d outside [1, 128]", k);
  if (n <= 0) return fail(h, RL_ERR_INVALID, "rl_recommend_host: n must be positive, got %d", n);
  if (n_users < 0 || n_items < 0) return fail(h, RL_ERR_INVALID, "rl_recommend_host: negative size");
  if (n_users == 0) return RL_OK;
  if (!user_f || !item_f || !idx_out) return fail(h, RL_ERR_INVALID, "rl_recommend_host: null pointer");
  // Users are processed in a few chunks (whole sweeps of the scoring kernel: multiples of 128 users x SM count) so that
  // the copy stream uploads chunk c + 1 (factor rows + its part of the seen lists) while chunk c is being scored, and
  // the results of chunk c travel back during chunk c + 1.
  DevBuf d_u, d_v, d_sp, d_sx, d_ub, d_ib, d_idx, d_sc, st_v;
  const int64_t nnz = seen_indptr ? seen_indptr[n_users] : 0;
  RL_TRY(d_u.alloc(h, sizeof(float) * n_users * kp));
  RL_TRY(d_v.alloc(h, sizeof(float) * (n_items ? n_items : 1) * kp));
  if (seen_indptr) {
    RL_TRY(d_sp.alloc(h, sizeof(int64_t) * (n_users + 1)));
    RL_TRY(d_sx.alloc(h, sizeof(int32_t) * (nnz ? nnz : 1)));
  }
  RL_TRY(d_idx.alloc(h, sizeof(int32_t) * n_users * n));
  if (score_out) RL_TRY(d_sc.alloc(h, sizeof(double) * n_users * n));
  if (user_bias) RL_TRY(upload(h, d_ub, user_bias, sizeof(float) * n_users));
  if (item_bias) RL_TRY(upload(h, d_ib, item_bias, sizeof(float) * n_items));
  cudaStream_t cs = h->copy_stream, ks = h->stream;
  // chunk boundaries, never a fraction of a sweep; large calls start with two small chunks (2 and 6 sweeps) so that
  // the first scoring kernel waits for a small upload only and each later upload hides behind the chunk before it
  const int64_t sweep = static_cast<int64_t>(128) * h->sm_count;
  const int64_t n_sweeps = (n_users + sweep - 1) / sweep;
  std::vector<int64_t> bound{0};
  if (n_sweeps >= 16) {
    bound.push_back(2 * sweep);
    bound.push_back(8 * sweep);
    const int64_t rest = n_sweeps - 8, per3 = ((rest + 5) / 6) * 2;      // even: the centred kernel takes user tiles in pairs
    for (int c = 1; c <= 2; ++c) bound.push_back((8 + c * per3) * sweep);
  } else if (n_sweeps >= 2) {
    const int64_t per4 = (n_sweeps + 3) / 4;
    for (int c = 1; c * per4 < n_sweeps; ++c) bound.push_back(c * per4 * sweep);
  }
  while (bound.back() >= n_users) bound.pop_back();
  if (bound.empty()) bound.push_back(0);
  bound.push_back(n_users);
  const int n_chunks = static_cast<int>(bound.size()) - 1;
  std::vector<Event> ready(n_chunks), done(n_chunks);
  std::vector<DevBuf> stage(n_chunks);
  Event ev_alloc, ev_copy_done;
  RL_TRY(ev_alloc.record(h, ks));
  RL_TRY(ev_alloc.wait(h, cs));
  {
    StreamScope on_copy(h, cs);
    RL_TRY(factors_h2d(h, item_f, factor_dtype, n_items, k, d_v.as<float>(), st_v));
    if (seen_indptr) RL_CUDA(h, cudaMemcpyAsync(d_sp.p, seen_indptr, sizeof(int64_t) * (n_users + 1), cudaMemcpyHostToDevice, cs));
    for (int c = 0; c < n_chunks; ++c) {
      const int64_t r0 = bound[c], r1 = bound[c + 1];
      const size_t es = factor_dtype == RL_F64 ? 8 : 4;
      RL_TRY(factors_h2d(h, static_cast<const char*>(user_f) + static_cast<size_t>(r0) * k * es, factor_dtype, r1 - r0, k,
                         d_u.as<float>() + r0 * kp, stage[c]));
      if (seen_indptr && seen_indptr[r1] > seen_indptr[r0])
        RL_CUDA(h, cudaMemcpyAsync(d_sx.as<int32_t>() + seen_indptr[r0], seen_indices + seen_indptr[r0],
                                   sizeof(int32_t) * (seen_indptr[r1] - seen_indptr[r0]), cudaMemcpyHostToDevice, cs));
      RL_TRY(ready[c].record(h, cs));
    }
  }
  // rows the tensor-core pass cannot finish are collected over all chunks and scored exactly once at the end: no host
  // round trip and no small launch per chunk
  DevBuf d_over;
  RL_TRY(d_over.alloc(h, 256 + sizeof(int32_t) * n_users));
  RL_CUDA(h, cudaMemsetAsync(d_over.p, 0, 256, ks));
  rl::ScoreDefer defer{reinterpret_cast<int32_t*>(static_cast<char*>(d_over.p) + 256), static_cast<unsigned int*>(d_over.p), 0};
  unsigned int overflow_total = 0;
  PrepScope prep(h);
  for (int c = 0; c < n_chunks; ++c) {
    const int64_t r0 = bound[c], r1 = bound[c + 1];
    RL_TRY(ready[c].wait(h, ks));
    h->last_overflow = 0;
    defer.row_offset = r0;
    // seen_indptr holds absolute offsets into the full seen_indices array: pass its window, keep the indices whole
    RL_TRY(launch_score_topn(h, d_u.as<float>() + r0 * kp, r1 - r0, d_v.as<float>(), n_items, k,
                             seen_indptr ? d_sp.as<int64_t>() + r0 : nullptr, seen_indptr ? d_sx.as<int32_t>() : nullptr, global_bias,
                             user_bias ? d_ub.as<float>() + r0 : nullptr, item_bias ? d_ib.as<float>() : nullptr, n,
                             d_idx.as<int32_t>() + r0 * n, score_out ? d_sc.as<double>() + r0 * n : nullptr, mode, &defer));
    overflow_total += h->last_overflow;
    RL_TRY(done[c].record(h, ks));
    RL_TRY(done[c].wait(h, cs));
    RL_CUDA(h, cudaMemcpyAsync(idx_out + r0 * n, d_idx.as<int32_t>() + r0 * n, sizeof(int32_t) * (r1 - r0) * n, cudaMemcpyDeviceToHost, cs));
    if (score_out)
      RL_CUDA(h, cudaMemcpyAsync(score_out + r0 * n, d_sc.as<double>() + r0 * n, sizeof(double) * (r1 - r0) * n, cudaMemcpyDeviceToHost, cs));
  }
  unsigned int n_over = 0;
  RL_CUDA(h, cudaMemcpyAsync(&n_over, defer.count, sizeof(unsigned int), cudaMemcpyDeviceToHost, ks));
  RL_CUDA(h, cudaStreamSynchronize(ks));
  if (n_over > 0) {
    RL_TRY(launch_score_overflow(h, d_u.as<float>(), n_users, d_v.as<float>(), n_items, k, seen_indptr ? d_sp.as<int64_t>() : nullptr,
                                 seen_indptr ? d_sx.as<int32_t>() : nullptr, n, d_idx.as<int32_t>(), score_out ? d_sc.as<double>() : nullptr,
                                 defer.rows, n_over));
    // their rows were copied home before they were final: send them again (a handful of rows, or everything)
    std::vector<int32_t> rows(n_over <= 256 ? n_over : 0);
    if (!rows.empty()) RL_CUDA(h, cudaMemcpyAsync(rows.data(), defer.rows, sizeof(int32_t) * n_over, cudaMemcpyDeviceToHost, ks));
    RL_TRY(ev_copy_done.record(h, cs));          // the chunk downloads come first
    RL_TRY(ev_copy_done.wait(h, ks));
    RL_CUDA(h, cudaStreamSynchronize(ks));
    if (!rows.empty()) {
      for (int32_t r : rows) {
        RL_CUDA(h, cudaMemcpyAsync(idx_out + static_cast<int64_t>(r) * n, d_idx.as<int32_t>() + static_cast<int64_t>(r) * n,
                                   sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ks));
        if (score_out)
          RL_CUDA(h, cudaMemcpyAsync(score_out + static_cast<int64_t>(r) * n, d_sc.as<double>() + static_cast<int64_t>(r) * n,
                                     sizeof(double) * n, cudaMemcpyDeviceToHost, ks));
      }
    } else {
      RL_CUDA(h, cudaMemcpyAsync(idx_out, d_idx.p, sizeof(int32_t) * n_users * n, cudaMemcpyDeviceToHost, ks));
      if (score_out) RL_CUDA(h, cudaMemcpyAsync(score_out, d_sc.p, sizeof(double) * n_users * n, cudaMemcpyDeviceToHost, ks));
    }
  }
  h->last_overflow = overflow_total + n_over;
  RL_TRY(ev_copy_done.record(h, cs));
  RL_TRY(ev_copy_done.wait(h, ks));
  RL_CUDA(h, cudaStreamSynchronize(ks));
  return RL_OK;
}

int rl_recommend_resident_host(rl_handle h, const float* user_f_dev, int64_t n_users, const float* item_f_dev, int64_t n_items, int k,
                               const int64_t* seen_indptr_dev, const int32_t* seen_indices_dev, int n, int32_t* idx_out, int mode) {
  RL_NEED(h);
  const int kp = rl::padded_k(k);
  if (kp == 0) return fail(h, RL_ERR_INVALID, "rl_recommend_resident_host: k=%d outside [1, 128]", k);
  if (n <= 0) return fail(h, RL_ERR_INVALID, "rl_recommend_resident_host: n must be positive, got %d", n);
  if (n_users < 0 || n_items < 0) return fail(h, RL_ERR_INVALID, "rl_recommend_resident_host: negative size");
  if (n_users == 0) return RL_OK;
  if (!user_f_dev || !item_f_dev || !idx_out) return fail(h, RL_ERR_INVALID, "rl_recommend_resident_host: null pointer");
  if ((seen_indptr_dev == nullptr) != (seen_indices_dev == nullptr))
    return fail(h, RL_ERR_INVALID, "rl_recommend_resident_host: pass both seen arrays or neither");
  // chunks of whole kernel sweeps; the rows of chunk c go home on the copy stream while chunk c + 1 is scored
  cudaStream_t cs = h->copy_stream, ks = h->stream;
  const int64_t sweep = static_cast<int64_t>(128) * h->sm_count;
  const int64_t n_sweeps = (n_users + sweep - 1) / sweep;
  const int64_t per = n_sweeps >= 8 ? ((n_sweeps + 7) / 8) * 2 : n_sweeps;      // even: the centred kernel takes user tiles in pairs
  std::vector<int64_t> bound{0};
  for (int64_t b = per * sweep; b < n_users; b += per * sweep) bound.push_back(b);
  bound.push_back(n_users);
  const int n_chunks = static_cast<int>(bound.size()) - 1;
  std::vector<Event> done(n_chunks);
  Event ev_copy_done;
  DevBuf d_idx, d_over;
  RL_TRY(d_idx.alloc(h, sizeof(int32_t) * n_users * n));
  RL_TRY(d_over.alloc(h, 256 + sizeof(int32_t) * n_users));
  RL_CUDA(h, cudaMemsetAsync(d_over.p, 0, 256, ks));
  rl::ScoreDefer defer{reinterpret_cast<int32_t*>(static_cast<char*>(d_over.p) + 256), static_cast<unsigned int*>(d_over.p), 0};
  unsigned int overflow_total = 0;
  PrepScope prep(h);
  for (int c = 0; c < n_chunks; ++c) {
    const int64_t r0 = bound[c], r1 = bound[c + 1];
    h->last_overflow = 0;
    defer.row_offset = r0;
    RL_TRY(launch_score_topn(h, user_f_dev + r0 * kp, r1 - r0, item_f_dev, n_items, k, seen_indptr_dev ? seen_indptr_dev + r0 : nullptr,
                             seen_indices_dev, 0.0, nullptr, nullptr, n, d_idx.as<int32_t>() + r0 * n, nullptr, mode, &defer));
    overflow_total += h->last_overflow;
    RL_TRY(done[c].record(h, ks));
    RL_TRY(done[c].wait(h, cs));
    RL_CUDA(h, cudaMemcpyAsync(idx_out + r0 * n, d_idx.as<int32_t>() + r0 * n, sizeof(int32_t) * (r1 - r0) * n, cudaMemcpyDeviceToHost, cs));
  }
  unsigned int n_over = 0;
  RL_CUDA(h, cudaMemcpyAsync(&n_over, defer.count, sizeof(unsigned int), cudaMemcpyDeviceToHost, ks));
  RL_CUDA(h, cudaStreamSynchronize(ks));
  if (n_over > 0) {      // rows the tensor-core pass could not finish: one exact launch, then they travel again
    RL_TRY(launch_score_overflow(h, user_f_dev, n_users, item_f_dev, n_items, k, seen_indptr_dev, seen_indices_dev, n, d_idx.as<int32_t>(),
                                 nullptr, defer.rows, n_over));
    std::vector<int32_t> rows(n_over <= 256 ? n_over : 0);
    if (!rows.empty()) RL_CUDA(h, cudaMemcpyAsync(rows.data(), defer.rows, sizeof(int32_t) * n_over, cudaMemcpyDeviceToHost, ks));
    RL_TRY(ev_copy_done.record(h, cs));
    RL_TRY(ev_copy_done.wait(h, ks));
    RL_CUDA(h, cudaStreamSynchronize(ks));
    if (!rows.empty()) {
      for (int32_t r : rows)
        RL_CUDA(h, cudaMemcpyAsync(idx_out + static_cast<int64_t>(r) * n, d_idx.as<int32_t>() + static_cast<int64_t>(r) * n,
                                   sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ks));
    } else {
      RL_CUDA(h, cudaMemcpyAsync(idx_out, d_idx.p, sizeof(int32_t) * n_users * n, cudaMemcpyDeviceToHost, ks));
    }
  }
  h->last_overflow = overflow_total + n_over;
  RL_TRY(ev_copy_done.record(h, cs));
  RL_TRY(ev_copy_done.wait(h, ks));
  RL_CUDA(h, cudaStreamSynchronize(ks));
  return RL_OK;
}

// ---- host-side CSR validation (no device work) -------------------------------------------------------------------------
extern "C++" {
namespace {
template <typename TP, typename TD>
int csr_check_rows(const TP* indptr, const int32_t* indices, const TD* data, int64_t r0, int64_t r1, int64_t n_cols) {
  int ok = 3;
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t s = static_cast<int64_t>(indptr[r]), e = static_cast<int64_t>(indptr[r + 1]);
    if (e < s) return 0;
    int64_t prev = -1;
    bool sorted = true, pos = true;
    for (int64_t p = s; p < e; ++p) {
      const int64_t c = indices[p];
      sorted &= c > prev;
      prev = c;
      if (data) pos &= data[p] > TD(0);
    }
    if (!sorted || prev >= n_cols) ok &= ~1;
    if (!pos) ok &= ~2;
    if (ok == 0) return 0;
  }
  return ok;
}
template <typename TP>
int csr_check_dispatch(const TP* indptr, const int32_t* indices, const void* data, int data_dtype, int64_t r0, int64_t r1, int64_t n_cols) {
  switch (data_dtype) {
    case RL_F32: return csr_check_rows(indptr, indices, static_cast<const float*>(data), r0, r1, n_cols);
    case RL_F64: return csr_check_rows(indptr, indices, static_cast<const double*>(data), r0, r1, n_cols);
    case RL_I32: return csr_check_rows(indptr, indices, static_cast<const int32_t*>(data), r0, r1, n_cols);
    default: return csr_check_rows(indptr, indices, static_cast<const int64_t*>(data), r0, r1, n_cols);
  }
}
}  // namespace
}  // extern "C++"

int rl_csr_check_host(const void* indptr, int indptr_dtype, const int32_t* indices, const void* data, int data_dtype, int64_t n_rows,
                      int64_t n_cols, int threads, int* flags_out) {
  if (!indptr || !flags_out || n_rows < 0 || n_cols < 0) return fail(nullptr, RL_ERR_INVALID, "rl_csr_check_host: bad argument");
  if ((indptr_dtype != RL_I32 && indptr_dtype != RL_I64) || (data && (data_dtype < RL_F32 || data_dtype > RL_I64)))
    return fail(nullptr, RL_ERR_INVALID, "rl_csr_check_host: bad dtype");
  const auto at = [&](int64_t r) -> int64_t {
    return indptr_dtype == RL_I32 ? static_cast<const int32_t*>(indptr)[r] : static_cast<const int64_t*>(indptr)[r];
  };
  *flags_out = 0;
  if (at(0) != 0) return RL_OK;
  const int64_t nnz = at(n_rows);
  if (nnz > 0 && !indices) return fail(nullptr, RL_ERR_INVALID, "rl_csr_check_host: null indices");
  int nt = threads > 0 ? threads : static_cast<int>(std::thread::hardware_concurrency());
  nt = std::max(1, std::min(nt, 32));
  if (nnz < (int64_t(1) << 20)) nt = 1;
  std::vector<int> part(nt, 3);
  std::vector<int64_t> cut(nt + 1, n_rows);
  cut[0] = 0;
  for (int t = 1; t < nt; ++t) {       // row ranges with about the same number of entries
    const int64_t target = nnz / nt * t;
    int64_t lo = cut[t - 1], hi = n_rows;
    while (lo < hi) { const int64_t mid = lo + (hi - lo) / 2; if (at(mid) < target) lo = mid + 1; else hi = mid; }
    cut[t] = lo;
  }
  const auto work = [&](int t) {
    part[t] = indptr_dtype == RL_I32
                  ? csr_check_dispatch(static_cast<const int32_t*>(indptr), indices, data, data_dtype, cut[t], cut[t + 1], n_cols)
                  : csr_check_dispatch(static_cast<const int64_t*>(indptr), indices, data, data_dtype, cut[t], cut[t + 1], n_cols);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  int ok = 3;
  for (int t = 0; t < nt; ++t) ok &= part[t];
  *flags_out = ok;
  return RL_OK;
}

int rl_topn_dense_dev(rl_handle h, const void* est_dev, int est_dtype, const void* known_dev, int known_dtype,
                      int64_t n_users, int64_t n_items, int n, int32_t* idx_out_dev, int32_t* n_valid_out_dev) {
  RL_NEED(h);
  return launch_topn_dense(h, est_dev, est_dtype, known_dev, known_dtype, n_users, n_items, n, idx_out_dev, n_valid_out_dev);
}

int rl_topn_dense_host(rl_handle h, const void* est, int est_dtype, const void* known, int known_dtype, int64_t n_users,
                       int64_t n_items, int n, int32_t* idx_out, int32_t* n_valid_out) {
  RL_NEED(h);
  if (n <= 0) return fail(h, RL_ERR_INVALID, "rl_topn_dense_host: n must be positive, got %d", n);
  if (n_users < 0 || n_items < 0) return fail(h, RL_ERR_INVALID, "rl_topn_dense_host: negative size");
  if (n_users == 0 || n_items == 0) return RL_OK;
  if (!est || !known || !idx_out) return fail(h, RL_ERR_INVALID, "rl_topn_dense_host: null pointer");
  const size_t cells = static_cast<size_t>(n_users) * n_items;
  DevBuf d_e, d_k, d_idx, d_nv;
  RL_TRY(upload(h, d_e, est, cells * (est_dtype == RL_F64 ? 8 : 4)));
  RL_TRY(upload(h, d_k, known, cells * (known_dtype == RL_F64 ? 8 : 4)));
  RL_TRY(d_idx.alloc(h, sizeof(int32_t) * n_users * n));
  RL_TRY(d_nv.alloc(h, sizeof(int32_t) * n_users));
  RL_TRY(launch_topn_dense(h, d_e.p, est_dtype, d_k.p, known_dtype, n_users, n_items, n, d_idx.as<int32_t>(), d_nv.as<int32_t>()));
  RL_CUDA(h, cudaMemcpyAsync(idx_out, d_idx.p, sizeof(int32_t) * n_users * n, cudaMemcpyDeviceToHost, h->stream));
  if (n_valid_out) RL_CUDA(h, cudaMemcpyAsync(n_valid_out, d_nv.p, sizeof(int32_t) * n_users, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  return RL_OK;
}

int rl_svd_topk_dev(rl_handle h, const float* m_dev, int64_t m, int64_t n, int k, double* s_out_dev, double* left_out_dev,
                    double* right_out_dev) {
  RL_NEED(h);
  return launch_svd_topk(h, m_dev, m, n, k, s_out_dev, left_out_dev, right_out_dev);
}

int rl_lowrank_dev(rl_handle h, const void* a_dev, const void* b_dev, int in_dtype, int64_t m, int64_t n, int k,
                   double global_bias, const float* row_bias_dev, const float* col_bias_dev, void* out_dev, int out_dtype) {
  RL_NEED(h);
  return launch_lowrank(h, a_dev, b_dev, in_dtype, m, n, k, global_bias, row_bias_dev, col_bias_dev, out_dev, out_dtype);
}

int rl_svd_reconstruct_host(rl_handle h, const float* mat, int64_t m, int64_t n, int k, float* recon, double* s_out,
                            double* left_out, double* right_out) {
  RL_NEED(h);
  if (m <= 0 || n <= 0) return fail(h, RL_ERR_INVALID, "rl_svd_reconstruct_host: zero dimension");
  if (k < 1 || k > (m < n ? m : n)) return fail(h, RL_ERR_INVALID, "rl_svd_reconstruct_host: k=%d out of range", k);
  if (!mat) return fail(h, RL_ERR_INVALID, "rl_svd_reconstruct_host: null pointer");
  DevBuf d_m, d_r, d_s, d_l, d_rt;
  const size_t cells = static_cast<size_t>(m) * n;
  RL_TRY(upload(h, d_m, mat, cells * sizeof(float)));
  RL_TRY(d_s.alloc(h, sizeof(double) * k));
  RL_TRY(d_l.alloc(h, sizeof(double) * m * k));
  RL_TRY(d_rt.alloc(h, sizeof(double) * n * k));
  RL_TRY(launch_svd_topk(h, d_m.as<float>(), m, n, k, d_s.as<double>(), d_l.as<double>(), d_rt.as<double>()));
  if (recon) {
    RL_TRY(d_r.alloc(h, cells * sizeof(float)));
    RL_TRY(launch_lowrank(h, d_l.p, d_rt.p, RL_F64, m, n, k, 0.0, nullptr, nullptr, d_r.p, RL_F32));
    RL_CUDA(h, cudaMemcpyAsync(recon, d_r.p, cells * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  }
  if (s_out) RL_CUDA(h, cudaMemcpyAsync(s_out, d_s.p, sizeof(double) * k, cudaMemcpyDeviceToHost, h->stream));
  if (left_out) RL_CUDA(h, cudaMemcpyAsync(left_out, d_l.p, sizeof(double) * m * k, cudaMemcpyDeviceToHost, h->stream));
  if (right_out) RL_CUDA(h, cudaMemcpyAsync(right_out, d_rt.p, sizeof(double) * n * k, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  return RL_OK;
}

// finite and not all zero?  flags[0] bit 0: non-finite seen, bit 1: non-zero seen
__global__ void __launch_bounds__(256) finite_any_kernel(const float* __restrict__ x, int64_t n, int* __restrict__ flags) {
  int f = 0;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n; p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = x[p];
    if (!isfinite(v)) f |= 1;
    if (v != 0.f) f |= 2;
  }
  f |= __shfl_xor_sync(0xffffffffu, f, 16); f |= __shfl_xor_sync(0xffffffffu, f, 8); f |= __shfl_xor_sync(0xffffffffu, f, 4);
  f |= __shfl_xor_sync(0xffffffffu, f, 2); f |= __shfl_xor_sync(0xffffffffu, f, 1);
  if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

int rl_gram_tc_dev(rl_handle h, const float* x_dev, int ldx, int op_x, const float* y_dev, int ldy, int op_y, int64_t n, int p, int q,
                   double* g_dev, int ldg, int terms) {
  RL_NEED(h);
  return launch_gram_tc(h, x_dev, ldx, op_x, y_dev, ldy, op_y, n, p, q, g_dev, ldg, terms);
}

int rl_knn_item_sim_host(rl_handle h, const float* ratings, int64_t m, int n, double* sim_out) {
  RL_NEED(h);
  if (!ratings || !sim_out || m <= 0 || n <= 0) return fail(h, RL_ERR_INVALID, "rl_knn_item_sim_host: bad argument");
  DevBuf d_r, d_s;
  RL_TRY(upload(h, d_r, ratings, sizeof(float) * m * n));
  RL_TRY(d_s.alloc(h, sizeof(double) * n * n));
  RL_TRY(launch_knn_item_sim(h, d_r.as<float>(), m, n, d_s.as<double>()));
  RL_CUDA(h, cudaMemcpyAsync(sim_out, d_s.p, sizeof(double) * n * n, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  return RL_OK;
}

int rl_knn_predict_host(rl_handle h, const float* ratings, int64_t m, int n, const double* sim, int neighbors, double* pred_out) {
  RL_NEED(h);
  if (!ratings || !sim || !pred_out || m <= 0 || n <= 0 || neighbors <= 0) return fail(h, RL_ERR_INVALID, "rl_knn_predict_host: bad argument");
  DevBuf d_r, d_s, d_p;
  RL_TRY(upload(h, d_r, ratings, sizeof(float) * m * n));
  RL_TRY(upload(h, d_s, sim, sizeof(double) * n * n));
  RL_TRY(d_p.alloc(h, sizeof(double) * m * n));
  RL_TRY(launch_knn_predict(h, d_r.as<float>(), m, n, d_s.as<double>(), neighbors, d_p.as<double>()));
  RL_CUDA(h, cudaMemcpyAsync(pred_out, d_p.p, sizeof(double) * m * n, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  return RL_OK;
}

int rl_ranking_metrics_host(rl_handle h, const int32_t* recs, int64_t n_users, int n, int k, const int64_t* actual_indptr,
                            const int32_t* actual_indices, double* out3) {
  RL_NEED(h);
  if (n_users < 0 || n <= 0 || k <= 0 || !out3) return fail(h, RL_ERR_INVALID, "rl_ranking_metrics_host: bad argument");
  out3[0] = out3[1] = out3[2] = 0.0;
  if (n_users == 0) return RL_OK;
  if (!recs || !actual_indptr) return fail(h, RL_ERR_INVALID, "rl_ranking_metrics_host: null pointer");
  const int64_t nnz = actual_indptr[n_users];
  DevBuf d_r, d_p, d_x, d_o;
  RL_TRY(upload(h, d_r, recs, sizeof(int32_t) * n_users * n));
  RL_TRY(upload(h, d_p, actual_indptr, sizeof(int64_t) * (n_users + 1)));
  RL_TRY(upload(h, d_x, actual_indices, sizeof(int32_t) * nnz));
  RL_TRY(d_o.alloc(h, 4 * sizeof(double)));
  RL_TRY(launch_ranking_metrics(h, d_r.as<int32_t>(), n_users, n, k, d_p.as<int64_t>(), d_x.as<int32_t>(), d_o.as<double>()));
  double o[4];
  RL_CUDA(h, cudaMemcpyAsync(o, d_o.p, sizeof(o), cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 3; ++i) out3[i] = o[i] / static_cast<double>(n_users);
  return RL_OK;
}

int rl_bias_centre_dev(rl_handle h, const float* ratings_dev, int64_t m, int64_t n, int ratings_is_f32, double* tot_dev,
                       double* user_bias_dev, double* item_bias_dev, float* centred_dev) {
  RL_NEED(h);
  return launch_bias_centre(h, ratings_dev, m, n, ratings_is_f32, tot_dev, user_bias_dev, item_bias_dev, centred_dev);
}

int rl_svd_fit_host(rl_handle h, const float* ratings, int64_t m, int64_t n, int k, int ratings_is_f32, double* bias_out,
                    double* user_bias_out, double* item_bias_out, double* s_out, double* left_out, double* right_out, int* status_out) {
  RL_NEED(h);
  if (m <= 0 || n <= 0) return fail(h, RL_ERR_INVALID, "rl_svd_fit_host: zero dimension");
  if (k < 1 || k > (m < n ? m : n)) return fail(h, RL_ERR_INVALID, "rl_svd_fit_host: k=%d out of range", k);
  if (!ratings || !bias_out || !user_bias_out || !item_bias_out || !s_out || !left_out || !right_out || !status_out)
    return fail(h, RL_ERR_INVALID, "rl_svd_fit_host: null pointer");
  DevBuf d_r, d_c, d_tot, d_ub, d_ib, d_s, d_l, d_rt, d_fl;
  const size_t cells = static_cast<size_t>(m) * n;
  RL_TRY(upload(h, d_r, ratings, cells * sizeof(float)));
  RL_TRY(d_c.alloc(h, cells * sizeof(float)));
  RL_TRY(d_tot.alloc(h, 2 * sizeof(double)));
  RL_TRY(d_ub.alloc(h, sizeof(double) * m));
  RL_TRY(d_ib.alloc(h, sizeof(double) * n));
  RL_TRY(d_fl.alloc(h, sizeof(int)));
  RL_TRY(launch_bias_centre(h, d_r.as<float>(), m, n, ratings_is_f32, d_tot.as<double>(), d_ub.as<double>(), d_ib.as<double>(), d_c.as<float>()));
  RL_CUDA(h, cudaMemsetAsync(d_fl.p, 0, sizeof(int), h->stream));
  {
    int64_t g = (static_cast<int64_t>(cells) + 255) / 256;
    if (g > 8 * h->sm_count) g = 8 * h->sm_count;
    finite_any_kernel<<<static_cast<unsigned>(g), 256, 0, h->stream>>>(d_c.as<float>(), static_cast<int64_t>(cells), d_fl.as<int>());
    RL_LAUNCH_CHECK(h, "finite_any_kernel");
  }
  int flags = 0;
  double tot[2] = {0.0, 0.0};
  RL_CUDA(h, cudaMemcpyAsync(&flags, d_fl.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaMemcpyAsync(tot, d_tot.p, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaMemcpyAsync(user_bias_out, d_ub.p, sizeof(double) * m, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaMemcpyAsync(item_bias_out, d_ib.p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  double gb = tot[1] > 0.0 ? tot[0] / tot[1] : std::nan("");
  if (ratings_is_f32) gb = static_cast<double>(static_cast<float>(gb));
  bias_out[0] = gb;
  const bool degenerate = (flags & 1) || !(flags & 2);
  status_out[0] = degenerate ? 1 : 0;
  if (degenerate) {
    for (int i = 0; i < k; ++i) s_out[i] = 0.0;
    for (int64_t i = 0; i < m * k; ++i) left_out[i] = 0.0;
    for (int64_t i = 0; i < n * k; ++i) right_out[i] = 0.0;
    return RL_OK;
  }
  RL_TRY(d_s.alloc(h, sizeof(double) * k));
  RL_TRY(d_l.alloc(h, sizeof(double) * m * k));
  RL_TRY(d_rt.alloc(h, sizeof(double) * n * k));
  RL_TRY(launch_svd_topk(h, d_c.as<float>(), m, n, k, d_s.as<double>(), d_l.as<double>(), d_rt.as<double>()));
  RL_CUDA(h, cudaMemcpyAsync(s_out, d_s.p, sizeof(double) * k, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaMemcpyAsync(left_out, d_l.p, sizeof(double) * m * k, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaMemcpyAsync(right_out, d_rt.p, sizeof(double) * n * k, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  return RL_OK;
}

int rl_predict_host(rl_handle h, const void* user_f, const void* item_f, int factor_dtype, int64_t n_users, int64_t n_items,
                    int k, double global_bias, const float* user_bias, const float* item_bias, void* out, int out_dtype) {
  RL_NEED(h);
  if (k <= 0) return fail(h, RL_ERR_INVALID, "rl_predict_host: k must be positive");
  if (n_users < 0 || n_items < 0) return fail(h, RL_ERR_INVALID, "rl_predict_host: negative size");
  if (n_users == 0 || n_items == 0) return RL_OK;
  if (!user_f || !item_f || !out) return fail(h, RL_ERR_INVALID, "rl_predict_host: null pointer");
  if (out_dtype != RL_F32 && out_dtype != RL_F64) return fail(h, RL_ERR_INVALID, "rl_predict_host: bad out dtype");
  if (factor_dtype != RL_F32 && factor_dtype != RL_F64) return fail(h, RL_ERR_INVALID, "rl_predict_host: bad factor dtype");
  // factors go to the device as dense float64 (n x k): no padding, no k limit, one rounding at the output
  DevBuf d_u, d_v, d_raw, d_ub, d_ib, d_o;
  const size_t es = factor_dtype == RL_F64 ? 8 : 4;
  if (factor_dtype == RL_F64) {
    RL_TRY(upload(h, d_u, user_f, es * n_users * k));
    RL_TRY(upload(h, d_v, item_f, es * n_items * k));
  } else {
    RL_TRY(d_u.alloc(h, sizeof(double) * n_users * k));
    RL_TRY(d_v.alloc(h, sizeof(double) * n_items * k));
    RL_TRY(d_raw.alloc(h, es * (n_users > n_items ? n_users : n_items) * k));
    RL_CUDA(h, cudaMemcpyAsync(d_raw.p, user_f, es * n_users * k, cudaMemcpyHostToDevice, h->stream));
    RL_TRY(launch_factors_convert(h, d_raw.p, RL_F32, n_users, k, k, d_u.p, RL_F64, k));
    RL_CUDA(h, cudaMemcpyAsync(d_raw.p, item_f, es * n_items * k, cudaMemcpyHostToDevice, h->stream));
    RL_TRY(launch_factors_convert(h, d_raw.p, RL_F32, n_items, k, k, d_v.p, RL_F64, k));
  }
  if (user_bias) RL_TRY(upload(h, d_ub, user_bias, sizeof(float) * n_users));
  if (item_bias) RL_TRY(upload(h, d_ib, item_bias, sizeof(float) * n_items));
  const size_t bytes = static_cast<size_t>(n_users) * n_items * (out_dtype == RL_F64 ? 8 : 4);
  RL_TRY(d_o.alloc(h, bytes));
  RL_TRY(launch_lowrank(h, d_u.p, d_v.p, RL_F64, n_users, n_items, k, global_bias, user_bias ? d_ub.as<float>() : nullptr,
                        item_bias ? d_ib.as<float>() : nullptr, d_o.p, out_dtype));
  RL_CUDA(h, cudaMemcpyAsync(out, d_o.p, bytes, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  return RL_OK;
}

int rl_dense_error_host(rl_handle h, const void* pred, int pred_dtype, const void* actual, int actual_dtype, int64_t n_cells,
                        double* out3, int* flags_out) {
  RL_NEED(h);
  if (n_cells < 0 || !out3 || !flags_out) return fail(h, RL_ERR_INVALID, "rl_dense_error_host: bad argument");
  out3[0] = out3[1] = out3[2] = 0.0;
  *flags_out = 0;
  if (n_cells == 0) return RL_OK;
  if (!pred || !actual) return fail(h, RL_ERR_INVALID, "rl_dense_error_host: null pointer");
  DevBuf d_p, d_a, d_o;
  RL_TRY(upload(h, d_p, pred, static_cast<size_t>(n_cells) * (pred_dtype == RL_F64 ? 8 : 4)));
  RL_TRY(upload(h, d_a, actual, static_cast<size_t>(n_cells) * (actual_dtype == RL_F64 ? 8 : 4)));
  RL_TRY(d_o.alloc(h, 64));
  double* o4 = d_o.as<double>();
  int* fl = reinterpret_cast<int*>(o4 + 4);
  RL_TRY(launch_dense_error(h, d_p.p, pred_dtype, d_a.p, actual_dtype, n_cells, o4, fl));
  double host[5];
  RL_CUDA(h, cudaMemcpyAsync(host, d_o.p, 40, cudaMemcpyDeviceToHost, h->stream));
  RL_CUDA(h, cudaStreamSynchronize(h->stream));
  out3[0] = host[0]; out3[1] = host[1]; out3[2] = host[2];
  memcpy(flags_out, &host[4], sizeof(int));
  return RL_OK;
}

int rl_observed_error_dev(rl_handle h, const float* user_f_dev, const float* item_f_dev, int k, int64_t n_users,
                          const int64_t* indptr_dev, const int32_t* indices_dev, const float* values_dev, double* out2_dev) {
  RL_NEED(h);
  return launch_observed_error(h, user_f_dev, item_f_dev, k, n_users, indptr_dev, indices_dev, values_dev, out2_dev);
}

}  // extern "C"
