// Shared helpers for the engine's translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "mvs_b200.h"

namespace mvs {

// thread-local last-error buffer behind mvs_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

// cudaMallocAsync from the device's default pool, whose release threshold is raised once per device:
// with the default of 0 the pool hands its memory back to the OS at every synchronisation, and the
// unmap / re-map of even a few kilobytes was measured to stall the next launches for 250-570 ms
// every few calls next to multi-GB allocations (the C3 pair-preparation launches).
cudaError_t pool_malloc(void** p, size_t bytes, cudaStream_t st);

#define MVS_CHECK_CUDA(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::mvs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                       __FILE__, __LINE__);                                         \
      return MVS_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define MVS_REQUIRE(cond, code, ...)                                                \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ::mvs::set_error(__VA_ARGS__);                                                \
      return (code);                                                                \
    }                                                                               \
  } while (0)

__host__ __device__ inline size_t dtype_size(int dt) {
  return dt == MVS_U8 ? 1 : (dt == MVS_U16 ? 2 : 4);
}

__device__ __forceinline__ float load_as_float(const void* p, int dt, int64_t i) {
  if (dt == MVS_F32) return __ldg(reinterpret_cast<const float*>(p) + i);
  if (dt == MVS_U16) return (float)__ldg(reinterpret_cast<const unsigned short*>(p) + i);
  return (float)__ldg(reinterpret_cast<const unsigned char*>(p) + i);
}

// float32 -> output dtype the way `np.nan_to_num(fused).astype(dtype)` does for
// in-range values (fusion/_core.py:1713): NaN -> 0, truncation toward zero.
// Out-of-range values (undefined behaviour in numpy) saturate here.
__device__ __forceinline__ void store_from_float(void* p, int dt, int64_t i, float v) {
  if (v != v) v = 0.0f;
  if (dt == MVS_F32) {
    reinterpret_cast<float*>(p)[i] = v;
  } else if (dt == MVS_U16) {
    int q = __float2int_rz(v);
    q = q < 0 ? 0 : (q > 65535 ? 65535 : q);
    reinterpret_cast<unsigned short*>(p)[i] = (unsigned short)q;
  } else {
    int q = __float2int_rz(v);
    q = q < 0 ? 0 : (q > 255 ? 255 : q);
    reinterpret_cast<unsigned char*>(p)[i] = (unsigned char)q;
  }
}

// 32-bit integer mixer (murmur3 finaliser) -- identical on host and device.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

__host__ __device__ __forceinline__ uint32_t hash3(uint32_t seed, int64_t z, int64_t y,
                                                   int64_t x) {
  uint32_t h = mix32(seed ^ 0x9e3779b9u);
  h = mix32(h ^ (uint32_t)(z & 0xffffffff) ^ (uint32_t)((uint64_t)z >> 32) * 0x27d4eb2fu);
  h = mix32(h ^ (uint32_t)(y & 0xffffffff) ^ (uint32_t)((uint64_t)y >> 32) * 0x165667b1u);
  h = mix32(h ^ (uint32_t)(x & 0xffffffff) ^ (uint32_t)((uint64_t)x >> 32) * 0xd3a2646cu);
  return h;
}

}  // namespace mvs
