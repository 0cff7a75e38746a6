// Error channel and device queries of the C ABI (include/mvs_b200.h).
#include "common.cuh"

#include <cstring>
#include <mutex>

namespace mvs {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* get_error() { return g_err; }

cudaError_t pool_malloc(void** p, size_t bytes, cudaStream_t st) {
  static std::mutex m;
  static bool done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !done[dev]) {
    std::lock_guard<std::mutex> l(m);
    if (!done[dev]) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
      done[dev] = true;
    }
  }
  return cudaMallocAsync(p, bytes, st);
}

}  // namespace mvs

extern "C" const char* mvs_last_error(void) { return mvs::get_error(); }

extern "C" int mvs_abi_version(void) { return MVS_ABI_VERSION; }

extern "C" int mvs_device_info(char* name, int cap, int* sm_count, int* cc_major,
                               int* cc_minor) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    mvs::set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0"
                                                          : cudaGetErrorString(e));
    return MVS_ERR_NO_DEVICE;
  }
  int dev = 0;
  MVS_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MVS_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (name && cap > 0) {
    strncpy(name, prop.name, cap - 1);
    name[cap - 1] = 0;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return MVS_OK;
}

// Lets bindings verify their struct layouts against the compiled ones.
extern "C" int mvs_struct_sizes(int* view_xform_bytes, int* chunk_bytes) {
  if (view_xform_bytes) *view_xform_bytes = (int)sizeof(mvs_view_xform);
  if (chunk_bytes) *chunk_bytes = (int)sizeof(mvs_chunk);
  return MVS_OK;
}
