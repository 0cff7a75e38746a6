// Post-resample stage of the fusion path on (V, *chunk) float32 stacks (sm_100a):
//   * normalize_weights                       weights.py:325-345
//   * content_based (Preibisch) weights       weights.py:22-74
//       nan_gaussian_filter                   weights.py:293-322
//       scipy.ndimage.gaussian_filter(mode="reflect", truncate=4): one correlate1d
//       pass per axis (z, y, x), float64 accumulation in scipy's symmetric order,
//       float32 between passes
//   * weighted_average_fusion / max_fusion / simple_average_fusion on stacks
//                                             fusion/_core.py:42-131
//   * trim + nan_to_num + cast                fusion/_core.py:1687-1713
// These are the arithmetic of the reference's fusion_func / weights_func hooks
// (post-resample level) and the multi-pass half of content-weighted
// fusion.  A stack is V contiguous volumes of N = nz*ny*nx voxels, NaN = outside.

#include <vector>

#include "common.cuh"

namespace mvs {

static int ew_grid(long long n) {
  long long b = (n + 255) / 256;
  return (int)(b < 148 * 16 ? (b < 1 ? 1 : b) : 148 * 16);
}

// w[v][i] /= nansum_v w[v][i]   (0 -> 1), weights.py:340-345
__global__ void __launch_bounds__(256)
normalize_weights_kernel(float* __restrict__ w, int V, long long N) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int v = 0; v < V; ++v) {
      const float x = w[(long long)v * N + i];
      s = __fadd_rn(s, x != x ? 0.f : x);
    }
    if (s == 0.f) s = 1.f;
    for (int v = 0; v < V; ++v) w[(long long)v * N + i] = __fdiv_rn(w[(long long)v * N + i], s);
  }
}

// views[bw < thresh] = NaN   (weights.py:53-54); also writes the NaN mask arrays
// V0 = nan->0 and W0 = 1 / 0 used by nan_gaussian_filter (weights.py:305-312)
__global__ void __launch_bounds__(256)
mask_split_kernel(const float* __restrict__ views, const float* __restrict__ bw, float thresh,
                  float* __restrict__ masked, float* __restrict__ v0, float* __restrict__ w0,
                  long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float v = views[i];
    if (bw != nullptr && bw[i] < thresh) v = NAN;
    const bool nan = v != v;
    if (masked) masked[i] = v;
    v0[i] = nan ? 0.f : v;
    w0[i] = nan ? 0.f : 1.f;
  }
}

// One scipy.ndimage.correlate1d pass with a symmetric kernel along `axis` of a
// batch of (nz, ny, nx) volumes, mode="reflect".  weights: fw[0..radius], fw[j] =
// weight at distance j from the centre (float64).  scipy (ni_filters.c,
// NI_Correlate1D, symmetric branch): tmp = in[0]*fw[0]; for j = radius..1:
// tmp += (in[-j] + in[+j]) * fw[j]; all float64; output rounded to float32.
__global__ void __launch_bounds__(256)
gauss1d_kernel(const float* __restrict__ in, float* __restrict__ out, int nz, int ny, int nx,
               int axis, const double* __restrict__ fw, int radius, long long total) {
  const int n = axis == 0 ? nz : (axis == 1 ? ny : nx);
  const long long stride = axis == 0 ? (long long)ny * nx : (axis == 1 ? nx : 1);
  const int period = 2 * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)((i / ((long long)nx * ny)) % nz);
    const int c = axis == 0 ? z : (axis == 1 ? y : x);
    const float* line = in + (i - (long long)c * stride);
    auto at = [&](int k) -> double {
      // half-sample symmetric reflection: d c b a | a b c d | d c b a
      if (k < 0 || k >= n) {
        k %= period;
        if (k < 0) k += period;
        if (k >= n) k = period - 1 - k;
      }
      return (double)__ldg(line + (long long)k * stride);
    };
    double tmp = __dmul_rn(at(c), fw[0]);
    for (int j = radius; j >= 1; --j)
      tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(at(c - j), at(c + j)), fw[j]));
    out[i] = (float)tmp;
  }
}

// Z = VV / WW with WW[nan] = 1, Z[nan] = NaN (weights.py:314-320); `ref` carries
// the NaN pattern.  With SQDIFF: out = (ref - Z)^2 in float32 (weights.py:57-65).
template <bool SQDIFF>
__global__ void __launch_bounds__(256)
nan_divide_kernel(const float* __restrict__ vv, const float* __restrict__ ww,
                  const float* __restrict__ ref, float* __restrict__ out, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float r = ref[i];
    float z;
    if (r != r) z = NAN;
    else z = __fdiv_rn(vv[i], ww[i]);
    if (SQDIFF) {
      const float d = __fsub_rn(r, z);
      z = __fmul_rn(d, d);
    }
    out[i] = z;
  }
}

// weighted_average_fusion (fusion/_core.py:61-94) on stacks; fw may be NULL.
__global__ void __launch_bounds__(256)
weighted_average_kernel(const float* __restrict__ views, const float* __restrict__ bw,
                        const float* __restrict__ fw, int V, long long N,
                        float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    if (fw) {
      for (int v = 0; v < V; ++v) {
        const float a = __fmul_rn(bw[(long long)v * N + i], fw[(long long)v * N + i]);
        s = __fadd_rn(s, a != a ? 0.f : a);
      }
      if (s == 0.f) s = 1.f;
    }
    float acc = 0.f;
    for (int v = 0; v < V; ++v) {
      float a = bw[(long long)v * N + i];
      if (fw) a = __fdiv_rn(__fmul_rn(a, fw[(long long)v * N + i]), s);
      const float p = __fmul_rn(views[(long long)v * N + i], a);
      acc = __fadd_rn(acc, p != p ? 0.f : p);
    }
    out[i] = acc;
  }
}

// max_fusion (:42-58) / simple_average_fusion (:97-131) on stacks
template <bool MEAN>
__global__ void __launch_bounds__(256)
reduce_views_kernel(const float* __restrict__ views, int V, long long N, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = MEAN ? 0.f : NAN, cnt = 0.f;
    for (int v = 0; v < V; ++v) {
      const float x = views[(long long)v * N + i];
      if (x != x) continue;
      if (MEAN) { acc = __fadd_rn(acc, x); cnt = __fadd_rn(cnt, 1.f); }
      else acc = (acc != acc) ? x : fmaxf(acc, x);
    }
    if (MEAN) acc = cnt > 0.f ? __fdiv_rn(acc, cnt) : NAN;
    out[i] = acc;
  }
}

// fused[trim:-trim] -> nan_to_num -> astype (fusion/_core.py:1687-1713)
__global__ void __launch_bounds__(256)
trim_cast_kernel(const float* __restrict__ in, int nz, int ny, int nx, int tz, int ty, int tx,
                 void* __restrict__ out, int dtype, int64_t sz, int64_t sy, int64_t sx) {
  const int oz = nz - 2 * tz, oy = ny - 2 * ty, ox = nx - 2 * tx;
  const long long total = (long long)oz * oy * ox;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % ox), y = (int)((i / ox) % oy), z = (int)(i / ((long long)ox * oy));
    const float v = in[((long long)(z + tz) * ny + (y + ty)) * nx + (x + tx)];
    store_from_float(out, dtype, z * sz + y * sy + x * sx, v);
  }
}

static int check_stack(const void* a, int V, const int32_t shape[3]) {
  MVS_REQUIRE(a != nullptr, MVS_ERR_INVALID, "NULL stack pointer");
  MVS_REQUIRE(V >= 1, MVS_ERR_INVALID, "V = %d", V);
  MVS_REQUIRE(shape && shape[0] >= 1 && shape[1] >= 1 && shape[2] >= 1, MVS_ERR_INVALID,
              "bad stack shape");
  return MVS_OK;
}

// gaussian_filter of `batch` volumes: src -> dst, using tmp as the ping-pong buffer
static int gaussian_batch(const float* src, float* dst, float* tmp, int batch, const int32_t shape[3],
                          int ndim, const double* d_fw, int radius, cudaStream_t st) {
  const long long total = (long long)batch * shape[0] * shape[1] * shape[2];
  const int grid = ew_grid(total);
  // axes in scipy's order (z, y, x); an odd number of passes must end in dst
  const int first_axis = 3 - ndim;
  const float* cur = src;
  for (int axis = first_axis; axis < 3; ++axis) {
    const int remaining = 3 - axis;  // passes left including this one
    float* o = (remaining % 2 == 1) ? dst : tmp;
    gauss1d_kernel<<<grid, 256, 0, st>>>(cur, o, shape[0], shape[1], shape[2], axis, d_fw, radius,
                                         total);
    MVS_CHECK_CUDA(cudaGetLastError());
    cur = o;
  }
  return MVS_OK;
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_normalize_weights(float* d_weights, int V, int64_t N, void* stream) {
  MVS_REQUIRE(d_weights && V >= 1 && N >= 0, MVS_ERR_INVALID, "bad arguments");
  if (N == 0) return MVS_OK;
  normalize_weights_kernel<<<ew_grid(N), 256, 0, (cudaStream_t)stream>>>(d_weights, V, N);
  MVS_CHECK_CUDA(cudaGetLastError());
  return MVS_OK;
}

extern "C" int mvs_gaussian_filter(const float* d_in, float* d_out, int batch,
                                   const int32_t shape[3], int ndim, const double* weights,
                                   int radius, void* stream) {
  int rc = check_stack(d_in, batch, shape);
  if (rc) return rc;
  MVS_REQUIRE(d_out && weights && radius >= 0, MVS_ERR_INVALID, "bad arguments");
  MVS_REQUIRE(ndim == 2 || ndim == 3, MVS_ERR_INVALID, "ndim must be 2 or 3");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)batch * shape[0] * shape[1] * shape[2];
  double* d_fw = nullptr;
  float* tmp = nullptr;
  MVS_CHECK_CUDA(cudaMalloc(&d_fw, sizeof(double) * (radius + 1)));
  cudaError_t e = cudaMalloc(&tmp, sizeof(float) * total);
  if (e != cudaSuccess) { cudaFree(d_fw); set_error("cudaMalloc: %s", cudaGetErrorString(e)); return MVS_ERR_CUDA; }
  cudaMemcpyAsync(d_fw, weights, sizeof(double) * (radius + 1), cudaMemcpyHostToDevice, st);
  rc = gaussian_batch(d_in, d_out, tmp, batch, shape, ndim, d_fw, radius, st);
  cudaStreamSynchronize(st);
  cudaFree(d_fw);
  cudaFree(tmp);
  return rc;
}

extern "C" int mvs_content_based(const float* d_views, const float* d_blending, int V,
                                 const int32_t shape[3], int ndim, const double* w1, int r1,
                                 const double* w2, int r2, float* d_out_weights, void* stream) {
  int rc = check_stack(d_views, V, shape);
  if (rc) return rc;
  MVS_REQUIRE(d_blending && d_out_weights && w1 && w2 && r1 >= 0 && r2 >= 0, MVS_ERR_INVALID,
              "bad arguments");
  MVS_REQUIRE(ndim == 2 || ndim == 3, MVS_ERR_INVALID, "ndim must be 2 or 3");
  MVS_REQUIRE(ndim == 3 || shape[0] == 1, MVS_ERR_INVALID, "2-D stack needs shape[0] == 1");
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = (long long)shape[0] * shape[1] * shape[2];
  const long long total = N * V;
  const int grid = ew_grid(total);
  // workspace: masked views M, and a batch of 2V volumes [V0 | W0] plus two
  // filter buffers of the same size
  float* ws = nullptr;
  double* d_fw = nullptr;
  MVS_CHECK_CUDA(cudaMalloc(&d_fw, sizeof(double) * (r1 + r2 + 2)));
  cudaError_t e = cudaMalloc(&ws, sizeof(float) * total * 7);
  if (e != cudaSuccess) {
    cudaFree(d_fw);
    set_error("content_based workspace (%lld bytes): %s", (long long)sizeof(float) * total * 7,
              cudaGetErrorString(e));
    return MVS_ERR_CUDA;
  }
  float* M = ws;                // masked views
  float* VW = ws + total;       // [V0 | W0]  (2*total)
  float* F = ws + 3 * total;    // filtered   (2*total)
  float* T = ws + 5 * total;    // ping-pong  (2*total)
  double* d_w1 = d_fw;
  double* d_w2 = d_fw + r1 + 1;
  cudaMemcpyAsync(d_w1, w1, sizeof(double) * (r1 + 1), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_w2, w2, sizeof(double) * (r2 + 1), cudaMemcpyHostToDevice, st);
  auto fail = [&](int code) { cudaStreamSynchronize(st); cudaFree(ws); cudaFree(d_fw); return code; };

  // transformed_views[blending_weights < 1e-7] = NaN; split into V0 / W0
  mask_split_kernel<<<grid, 256, 0, st>>>(d_views, d_blending, 1e-7f, M, VW, VW + total, total);
  // inner nan-gaussian (sigma_1) and squared difference
  if ((rc = gaussian_batch(VW, F, T, 2 * V, shape, ndim, d_w1, r1, st))) return fail(rc);
  nan_divide_kernel<true><<<grid, 256, 0, st>>>(F, F + total, M, T, total);  // T = (M - Z)^2
  // outer nan-gaussian (sigma_2)
  mask_split_kernel<<<grid, 256, 0, st>>>(T, nullptr, 0.f, nullptr, VW, VW + total, total);
  // keep the NaN pattern of the squared difference (== pattern of M)
  if ((rc = gaussian_batch(VW, F, T, 2 * V, shape, ndim, d_w2, r2, st))) return fail(rc);
  nan_divide_kernel<false><<<grid, 256, 0, st>>>(F, F + total, M, d_out_weights, total);
  normalize_weights_kernel<<<ew_grid(N), 256, 0, st>>>(d_out_weights, V, N);
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("content_based launch: %s", cudaGetErrorString(e)); return fail(MVS_ERR_CUDA); }
  return fail(MVS_OK);
}

extern "C" int mvs_fuse_stack(const float* d_views, const float* d_blending,
                              const float* d_fusion_weights, int V, int64_t N, int fusion_mode,
                              float* d_out, void* stream) {
  MVS_REQUIRE(d_views && d_out && V >= 1 && N >= 0, MVS_ERR_INVALID, "bad arguments");
  if (N == 0) return MVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ew_grid(N);
  if (fusion_mode == MVS_FUSE_WAVG) {
    MVS_REQUIRE(d_blending != nullptr, MVS_ERR_INVALID, "weighted average needs blending weights");
    weighted_average_kernel<<<grid, 256, 0, st>>>(d_views, d_blending, d_fusion_weights, V, N, d_out);
  } else if (fusion_mode == MVS_FUSE_MAX) {
    reduce_views_kernel<false><<<grid, 256, 0, st>>>(d_views, V, N, d_out);
  } else if (fusion_mode == MVS_FUSE_MEAN) {
    reduce_views_kernel<true><<<grid, 256, 0, st>>>(d_views, V, N, d_out);
  } else {
    set_error("unknown fusion mode %d", fusion_mode);
    return MVS_ERR_INVALID;
  }
  MVS_CHECK_CUDA(cudaGetLastError());
  return MVS_OK;
}

extern "C" int mvs_trim_cast(const float* d_in, const int32_t shape[3], const int32_t trim[3],
                             void* d_out, int out_dtype, const int64_t out_stride[3],
                             void* stream) {
  MVS_REQUIRE(d_in && d_out && shape && trim && out_stride, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(out_dtype >= MVS_U8 && out_dtype <= MVS_F32, MVS_ERR_INVALID, "bad dtype");
  for (int d = 0; d < 3; ++d)
    MVS_REQUIRE(trim[d] >= 0 && shape[d] - 2 * trim[d] >= 0, MVS_ERR_INVALID, "bad trim");
  const long long total = (long long)(shape[0] - 2 * trim[0]) * (shape[1] - 2 * trim[1]) *
                          (shape[2] - 2 * trim[2]);
  if (total == 0) return MVS_OK;
  trim_cast_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(
      d_in, shape[0], shape[1], shape[2], trim[0], trim[1], trim[2], d_out, out_dtype,
      out_stride[0], out_stride[1], out_stride[2]);
  MVS_CHECK_CUDA(cudaGetLastError());
  return MVS_OK;
}
