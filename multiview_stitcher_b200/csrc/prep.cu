// Mean-binning on the device: the coarsening register_pair_of_msims applies before
// cropping (registration.py:1732-1743) and the level-to-level step of
// the output pyramid (ngff_utils.py:1284-1285, msi_utils.py:49-60).
// The crop itself is a strided window (no kernel) and the resampling onto the
// fixed view's grid is mvs_resample_views (fuse.cu).
#include "common.cuh"

#include <algorithm>

namespace mvs {

// One thread per output voxel; the window is walked in C order with a float64
// accumulator: integer sums are exact, so `mean(dtype=float64).astype(dtype)`
// is reproduced bit for bit; float32 windows skip NaNs like xarray's skipna mean
// (SKIP) or propagate them like np.mean (pyramid levels).
template <typename T, bool SKIP>
__global__ void __launch_bounds__(256)
bin_mean_kernel(const T* __restrict__ in, int64_t sz, int64_t sy, int64_t sx, int nz, int ny,
                int nx, int bz, int by, int bx, T* __restrict__ out) {
  // one warp per 32-output row segment, 32-bit index arithmetic (no 64-bit divisions per output)
  const unsigned rows = (unsigned)nz * (unsigned)ny, xt = ((unsigned)nx + 31) / 32;
  const unsigned nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (unsigned u = blockIdx.x * nwarps + (threadIdx.x >> 5); u < rows * xt; u += gridDim.x * nwarps) {
    const unsigned row = u / xt, xu = (u - row * xt) * 32 + lane;
    if (xu >= (unsigned)nx) continue;
    const int z = (int)(row / (unsigned)ny), y = (int)(row - (unsigned)z * (unsigned)ny), x = (int)xu;
    const int64_t i = (int64_t)row * nx + x;
    const T* p = in + (int64_t)z * bz * sz + (int64_t)y * by * sy + (int64_t)x * bx * sx;
    double acc = 0.0;
    int cnt = 0;
    for (int dz = 0; dz < bz; ++dz)
      for (int dy = 0; dy < by; ++dy) {
        const T* row = p + dz * sz + dy * sy;
        for (int dx = 0; dx < bx; ++dx) {
          const T v = __ldg(row + dx * sx);
          if (!SKIP || v == v) { acc += (double)v; ++cnt; }
        }
      }
    T res;
    if (sizeof(T) == 4) {  // float32
      const double m = cnt ? acc / (double)cnt : (double)__int_as_float(0x7fc00000);
      res = (T)m;
    } else {
      res = (T)(long long)(acc / (double)cnt);  // truncation like astype
    }
    out[i] = res;
  }
}

}  // namespace mvs

extern "C" int mvs_bin_mean(const void* d_in, int dtype, const int32_t shape[3],
                            const int64_t stride[3], const int32_t bin[3], int skip_nan,
                            void* d_out, void* stream) {
  using namespace mvs;
  MVS_REQUIRE(d_in && d_out && shape && stride && bin, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(dtype == MVS_U8 || dtype == MVS_U16 || dtype == MVS_F32, MVS_ERR_UNSUPPORTED,
              "dtype %d", dtype);
  for (int d = 0; d < 3; ++d)
    MVS_REQUIRE(bin[d] >= 1 && shape[d] >= 1, MVS_ERR_INVALID, "axis %d: bin %d, shape %d", d,
                bin[d], shape[d]);
  const int nz = shape[0] / bin[0], ny = shape[1] / bin[1], nx = shape[2] / bin[2];  // boundary="trim"
  const long long N = (long long)nz * ny * nx;
  if (N <= 0) return MVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)std::min<long long>((N + 255) / 256, 148 * 8);
#define MVS_BIN_LAUNCH(T, SKIP)                                                              \
  bin_mean_kernel<T, SKIP><<<grid, 256, 0, st>>>((const T*)d_in, stride[0], stride[1], stride[2], \
                                                 nz, ny, nx, bin[0], bin[1], bin[2], (T*)d_out)
  if (dtype == MVS_U8) MVS_BIN_LAUNCH(unsigned char, false);
  else if (dtype == MVS_U16) MVS_BIN_LAUNCH(unsigned short, false);
  else if (skip_nan) MVS_BIN_LAUNCH(float, true);
  else MVS_BIN_LAUNCH(float, false);
#undef MVS_BIN_LAUNCH
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("bin_mean launch: %s", cudaGetErrorString(e)); return MVS_ERR_CUDA; }
  return MVS_OK;
}
