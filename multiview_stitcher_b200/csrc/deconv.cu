// Multi-view deconvolution fusion (fusion/mv_deconv.py:251-500) on the GPU: the per-view,
// per-iteration pair of PSF-sized convolutions of the Richardson-Lucy update with their
// element-wise steps fused into the epilogues.
//
//   forward      wr  = 1 + w_v * (ratio - 1),  ratio = covered ? img_v / max(psi (*) PSF_v, eps) : 1
//                (scipy.ndimage.convolve(psi, kernel1, mode="mirror"), :444-462)
//   back         psi = clamp(reg(psi * (wr (*) kernel2_v)))
//                (convolve(weighted_ratio, kernel2, mode="constant", cval=1), :463-486)
//
// conv3_kernel: a CTA owns an 8 x 8 x 32 output tile; the input tile with its halo is staged
// once in shared memory (boundary rule applied while staging) together with the flipped
// kernel; a thread owns one (y, x) column of 8 outputs and, per (ky, kx), keeps the z line in
// registers so every shared-memory load feeds up to KZ multiply-adds.  Compute-bound by
// design (729 taps per voxel for the default 9^3 PSF): no HBM roofline applies.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace mvs {

constexpr int kCvTZ = 8, kCvTY = 8, kCvTX = 32;
constexpr int kCvMaxK = 15;

struct ConvArgs {
  const float* in;
  float* out;
  const float* wcorr;   // flipped kernel (correlation weights), k[0]*k[1]*k[2]
  int n[3];
  int k[3];
  int mode;             // 0 mirror, 1 constant
  float cval;
  int epi;              // 0 plain, 1 forward (weighted ratio), 2 back (psi update)
  const float* view;    // epi 1: transformed view (NaN = not covered)
  const float* weight;  // epi 1: blending weight of the view
  const float* psi;     // epi 2: current estimate
  float min_value, lambda_reg, max_intensity;
};

__device__ __forceinline__ int mirror_index(int i, int n) {
  if (n == 1) return 0;
  // scipy "mirror": d c b | a b c d | c b a  (period 2n - 2)
  const int p = 2 * n - 2;
  i = i % p;
  if (i < 0) i += p;
  return i < n ? i : p - i;
}

template <int KZ>
__global__ void __launch_bounds__(256)
conv3_kernel(ConvArgs a) {
  extern __shared__ float sm[];
  const int KY = a.k[1], KX = a.k[2];
  const int tz = kCvTZ + KZ - 1, ty = kCvTY + KY - 1, tx = kCvTX + KX - 1;
  float* tile = sm;                       // [tz][ty][tx]
  float* wk = sm + tz * ty * tx;          // [KZ][KY][KX]
  const int nbx = (a.n[2] + kCvTX - 1) / kCvTX, nby = (a.n[1] + kCvTY - 1) / kCvTY;
  int b = blockIdx.x;
  const int bx = b % nbx; b /= nbx;
  const int by = b % nby;
  const int bz = b / nby;
  const int x0 = bx * kCvTX, y0 = by * kCvTY, z0 = bz * kCvTZ;
  const int cz = KZ / 2, cy = KY / 2, cx = KX / 2;
  const int64_t sy = a.n[2], sz = (int64_t)a.n[1] * a.n[2];

  for (int i = threadIdx.x; i < KZ * KY * KX; i += 256) wk[i] = __ldg(a.wcorr + i);
  for (int i = threadIdx.x; i < tz * ty * tx; i += 256) {
    const int lx = i % tx, ly = (i / tx) % ty, lz = i / (tx * ty);
    int gz = z0 + lz - cz, gy = y0 + ly - cy, gx = x0 + lx - cx;
    float v;
    if (a.mode == 0) {
      gz = mirror_index(gz, a.n[0]); gy = mirror_index(gy, a.n[1]); gx = mirror_index(gx, a.n[2]);
      v = __ldg(a.in + gz * sz + gy * sy + gx);
    } else {
      const bool inside = gz >= 0 && gz < a.n[0] && gy >= 0 && gy < a.n[1] && gx >= 0 && gx < a.n[2];
      v = inside ? __ldg(a.in + gz * sz + gy * sy + gx) : a.cval;
    }
    tile[i] = v;
  }
  __syncthreads();

  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  float acc[kCvTZ];
#pragma unroll
  for (int o = 0; o < kCvTZ; ++o) acc[o] = 0.f;
  for (int ky = 0; ky < KY; ++ky) {
    for (int kx = 0; kx < KX; ++kx) {
      float col[kCvTZ + KZ - 1];
      const float* p = tile + (ly + ky) * tx + (lx + kx);
#pragma unroll
      for (int z = 0; z < kCvTZ + KZ - 1; ++z) col[z] = p[z * ty * tx];
#pragma unroll
      for (int kz = 0; kz < KZ; ++kz) {
        const float w = wk[(kz * KY + ky) * KX + kx];
#pragma unroll
        for (int o = 0; o < kCvTZ; ++o) acc[o] = fmaf(w, col[o + kz], acc[o]);
      }
    }
  }
  const int x = x0 + lx, y = y0 + ly;
  if (x >= a.n[2] || y >= a.n[1]) return;
#pragma unroll
  for (int o = 0; o < kCvTZ; ++o) {
    const int z = z0 + o;
    if (z >= a.n[0]) break;
    const int64_t idx = z * sz + y * sy + x;
    float r = acc[o];
    if (a.epi == 1) {
      // mv_deconv.py:448-462 (float32 arithmetic)
      const float img = a.view[idx];
      const bool covered = img == img;
      const float ratio = covered ? __fdiv_rn(img, fmaxf(r, a.min_value)) : 1.0f;
      r = __fadd_rn(1.0f, __fmul_rn(a.weight[idx], __fsub_rn(ratio, 1.0f)));
    } else if (a.epi == 2) {
      // mv_deconv.py:465-486
      float value = __fmul_rn(a.psi[idx], r);
      if (a.lambda_reg > 0.f) {
        const float xr = __fdiv_rn(fmaxf(value, 0.0f), a.max_intensity);
        const float s = __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(2.0f * a.lambda_reg, xr)));
        value = __fmul_rn(__fdiv_rn(__fsub_rn(s, 1.0f), a.lambda_reg), a.max_intensity);
      }
      r = value != value ? a.min_value : fmaxf(value, a.min_value);
    }
    a.out[idx] = r;
  }
}

// psi0 = clip(nansum_v(nan_to_num(view_v) * w_v), min_value)   (mv_deconv.py:420-421)
__global__ void deconv_init_kernel(const float* __restrict__ views, const float* __restrict__ weights, int V,
                                   int64_t N, float min_value, float* __restrict__ psi) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int v = 0; v < V; ++v) {
      float t = views[v * N + i];
      if (t != t) t = 0.f;
      const float p = __fmul_rn(t, weights[v * N + i]);
      if (p == p) s = __fadd_rn(s, p);
    }
    psi[i] = fmaxf(s, min_value);
  }
}

// union coverage mask, one erosion step (face neighbours, border value 1), final masking
__global__ void deconv_union_kernel(const float* __restrict__ views, int V, int64_t N, unsigned char* __restrict__ m) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned char any = 0;
    for (int v = 0; v < V; ++v) {
      const float t = views[v * N + i];
      any |= (t == t) ? 1 : 0;
    }
    m[i] = any;
  }
}
__global__ void deconv_erode_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int nz,
                                    int ny, int nx, int ndim) {
  const int64_t N = (int64_t)nz * ny * nx;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((int64_t)nx * ny));
    unsigned char r = in[i];
    if (x > 0) r &= in[i - 1];
    if (x < nx - 1) r &= in[i + 1];
    if (y > 0) r &= in[i - nx];
    if (y < ny - 1) r &= in[i + nx];
    if (ndim == 3) {
      if (z > 0) r &= in[i - (int64_t)nx * ny];
      if (z < nz - 1) r &= in[i + (int64_t)nx * ny];
    }
    out[i] = r;
  }
}
__global__ void deconv_mask_kernel(float* __restrict__ psi, const unsigned char* __restrict__ m, int64_t N) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    if (!m[i]) psi[i] = 0.f;
}

static cudaError_t launch_conv(const ConvArgs& a, cudaStream_t st) {
  const int64_t nb = (int64_t)((a.n[2] + kCvTX - 1) / kCvTX) * ((a.n[1] + kCvTY - 1) / kCvTY) * ((a.n[0] + kCvTZ - 1) / kCvTZ);
  const size_t smem = sizeof(float) * ((size_t)(kCvTZ + a.k[0] - 1) * (kCvTY + a.k[1] - 1) * (kCvTX + a.k[2] - 1) +
                                       (size_t)a.k[0] * a.k[1] * a.k[2]);
#define MVS_CONV_CASE(KZ)                                                                           \
  case KZ: {                                                                                        \
    cudaError_t e = cudaFuncSetAttribute(conv3_kernel<KZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                 \
    conv3_kernel<KZ><<<(unsigned)nb, 256, smem, st>>>(a);                                           \
    break;                                                                                          \
  }
  switch (a.k[0]) {
    MVS_CONV_CASE(1) MVS_CONV_CASE(3) MVS_CONV_CASE(5) MVS_CONV_CASE(7)
    MVS_CONV_CASE(9) MVS_CONV_CASE(11) MVS_CONV_CASE(13) MVS_CONV_CASE(15)
    default: return cudaErrorInvalidValue;
  }
#undef MVS_CONV_CASE
  return cudaGetLastError();
}

}  // namespace mvs

extern "C" int mvs_convolve(const float* d_in, float* d_out, const int32_t shape[3], const float* kernel,
                            const int32_t kshape[3], int mode, float cval, void* stream) {
  using namespace mvs;
  MVS_REQUIRE(d_in && d_out && shape && kernel && kshape, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(d_in != d_out, MVS_ERR_INVALID, "in-place convolution is not supported");
  MVS_REQUIRE(mode == 0 || mode == 1, MVS_ERR_INVALID, "mode must be 0 (mirror) or 1 (constant)");
  ConvArgs a{};
  int64_t kn = 1;
  for (int d = 0; d < 3; ++d) {
    MVS_REQUIRE(shape[d] >= 1, MVS_ERR_INVALID, "empty extent");
    MVS_REQUIRE(kshape[d] >= 1 && kshape[d] <= kCvMaxK && (kshape[d] & 1), MVS_ERR_UNSUPPORTED,
                "kernel extent %d (odd, <= %d)", kshape[d], kCvMaxK);
    a.n[d] = shape[d]; a.k[d] = kshape[d];
    kn *= kshape[d];
  }
  cudaStream_t st = (cudaStream_t)stream;
  // flipped copy: convolution as correlation (scipy.ndimage.convolve, odd sizes)
  std::vector<float> flipped((size_t)kn);
  for (int64_t i = 0; i < kn; ++i) flipped[(size_t)i] = kernel[kn - 1 - i];
  float* d_w = nullptr;
  MVS_CHECK_CUDA(mvs::pool_malloc((void**)&d_w, sizeof(float) * kn, st));
  MVS_CHECK_CUDA(cudaMemcpyAsync(d_w, flipped.data(), sizeof(float) * kn, cudaMemcpyHostToDevice, st));
  a.in = d_in; a.out = d_out; a.wcorr = d_w; a.mode = mode; a.cval = cval; a.epi = 0;
  cudaError_t e = launch_conv(a, st);
  cudaFreeAsync(d_w, st);
  if (e != cudaSuccess) { set_error("convolve launch: %s", cudaGetErrorString(e)); return MVS_ERR_CUDA; }
  return MVS_OK;
}

extern "C" int mvs_mv_deconvolution(const float* d_views, const float* d_weights, int V, const int32_t shape[3],
                                    int ndim, const float* kernels1, const float* kernels2,
                                    const int32_t kshape[3], int n_iterations, float lambda_reg,
                                    float min_value, int erosion_px, float* d_out, void* stream) {
  using namespace mvs;
  MVS_REQUIRE(d_views && d_weights && shape && kernels1 && kernels2 && kshape && d_out, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(V >= 1 && n_iterations >= 0 && erosion_px >= 0, MVS_ERR_INVALID, "bad count");
  MVS_REQUIRE(ndim == 2 || ndim == 3, MVS_ERR_INVALID, "ndim must be 2 or 3");
  ConvArgs base{};
  int64_t kn = 1, N = 1;
  for (int d = 0; d < 3; ++d) {
    MVS_REQUIRE(shape[d] >= 1, MVS_ERR_INVALID, "empty extent");
    MVS_REQUIRE(kshape[d] >= 1 && kshape[d] <= kCvMaxK && (kshape[d] & 1), MVS_ERR_UNSUPPORTED,
                "PSF extent %d (odd, <= %d)", kshape[d], kCvMaxK);
    base.n[d] = shape[d]; base.k[d] = kshape[d];
    kn *= kshape[d]; N *= shape[d];
  }
  cudaStream_t st = (cudaStream_t)stream;
  // flipped kernels for all views, both kinds
  std::vector<float> flipped((size_t)(2 * V * kn));
  for (int v = 0; v < V; ++v)
    for (int64_t i = 0; i < kn; ++i) {
      flipped[(size_t)(v * kn + i)] = kernels1[v * kn + (kn - 1 - i)];
      flipped[(size_t)((V + v) * kn + i)] = kernels2[v * kn + (kn - 1 - i)];
    }
  float *d_w = nullptr, *d_psi = nullptr, *d_tmp = nullptr, *d_wr = nullptr;
  MVS_CHECK_CUDA(mvs::pool_malloc((void**)&d_w, sizeof(float) * 2 * V * kn, st));
  MVS_CHECK_CUDA(cudaMemcpyAsync(d_w, flipped.data(), sizeof(float) * 2 * V * kn, cudaMemcpyHostToDevice, st));
  MVS_CHECK_CUDA(mvs::pool_malloc((void**)&d_psi, sizeof(float) * N, st));
  MVS_CHECK_CUDA(mvs::pool_malloc((void**)&d_tmp, sizeof(float) * N, st));
  MVS_CHECK_CUDA(mvs::pool_malloc((void**)&d_wr, sizeof(float) * N, st));
  const unsigned grid = (unsigned)std::min<int64_t>((N + 255) / 256, 148 * 16);
  deconv_init_kernel<<<grid, 256, 0, st>>>(d_views, d_weights, V, N, min_value, d_psi);
  MVS_CHECK_CUDA(cudaGetLastError());
  // max_intensity = psi.max() (mv_deconv.py:423-425): a tiny reduction on the host side of the ABI
  float max_intensity = 1.f;
  if (lambda_reg > 0.f) {
    std::vector<float> h((size_t)N);
    MVS_CHECK_CUDA(cudaMemcpyAsync(h.data(), d_psi, sizeof(float) * N, cudaMemcpyDeviceToHost, st));
    MVS_CHECK_CUDA(cudaStreamSynchronize(st));
    max_intensity = *std::max_element(h.begin(), h.end());
    if (!(max_intensity > 0.f)) max_intensity = 1.f;
  }
  cudaError_t e = cudaSuccess;
  for (int it = 0; it < n_iterations && e == cudaSuccess; ++it) {
    for (int v = 0; v < V && e == cudaSuccess; ++v) {
      ConvArgs f = base;
      f.in = d_psi; f.out = d_wr; f.wcorr = d_w + v * kn; f.mode = 0; f.cval = 0.f; f.epi = 1;
      f.view = d_views + v * N; f.weight = d_weights + v * N; f.min_value = min_value;
      e = launch_conv(f, st);
      if (e != cudaSuccess) break;
      ConvArgs g = base;
      g.in = d_wr; g.out = d_tmp; g.wcorr = d_w + (V + v) * kn; g.mode = 1; g.cval = 1.f; g.epi = 2;
      g.psi = d_psi; g.min_value = min_value; g.lambda_reg = lambda_reg; g.max_intensity = max_intensity;
      e = launch_conv(g, st);
      std::swap(d_psi, d_tmp);
    }
  }
  if (e == cudaSuccess && erosion_px > 0) {
    unsigned char *m0 = nullptr, *m1 = nullptr;
    MVS_CHECK_CUDA(mvs::pool_malloc((void**)&m0, (size_t)N, st));
    MVS_CHECK_CUDA(mvs::pool_malloc((void**)&m1, (size_t)N, st));
    deconv_union_kernel<<<grid, 256, 0, st>>>(d_views, V, N, m0);
    for (int k = 0; k < erosion_px; ++k) {
      deconv_erode_kernel<<<grid, 256, 0, st>>>(m0, m1, shape[0], shape[1], shape[2], ndim);
      std::swap(m0, m1);
    }
    deconv_mask_kernel<<<grid, 256, 0, st>>>(d_psi, m0, N);
    e = cudaGetLastError();
    cudaFreeAsync(m0, st);
    cudaFreeAsync(m1, st);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_out, d_psi, sizeof(float) * N, cudaMemcpyDeviceToDevice, st);
  cudaFreeAsync(d_w, st);
  cudaFreeAsync(d_psi, st);
  cudaFreeAsync(d_tmp, st);
  cudaFreeAsync(d_wr, st);
  if (e != cudaSuccess) { set_error("deconvolution: %s", cudaGetErrorString(e)); return MVS_ERR_CUDA; }
  return MVS_OK;
}
