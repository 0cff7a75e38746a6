// DCT-entropy content weights (weights.content_based_dct, weights.py:77-290) on the GPU.
//
// The reference cuts every view into blocks of dct_size^ndim voxels, takes the orthonormal
// DCT-II of each block (scipy.fftpack.dctn), and scores the block by the Shannon entropy of
// the coefficients inside the OTF support (L1 frequency index < r_o) normalised by the
// block's L2 norm; the per-block scores are normalised over the views, interpolated back to
// voxel resolution (scipy affine_transform, order 1, mode "nearest") and normalised again.
//
//   dct_quality_kernel   one CTA per (view, block): the block is staged once in shared
//                        memory (NaN-filled like the reference), and the separable DCT is
//                        evaluated IN PLACE along x, y, z -- only for the coefficients
//                        inside the OTF support (816 of 32768 for 32^3, r_o = 16), a thread
//                        per line with the line in registers; entropy by a block reduction.
//   dct_qnorm_kernel     per block: subtract the minimum over the views, normalise.
//   dct_weights_kernel   per voxel: multilinear lookup in the normalised score maps at
//                        scipy's float64 coordinates, normalise over the views.
// HBM traffic: every view voxel is read once (twice through L2 for the NaN statistics) and
// V weights are written per voxel.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace mvs {

constexpr int kDctMax = 32;            // largest block edge
constexpr int kDctPitch = kDctMax + 1; // row pitch of the staged block (bank-conflict free)

struct DctArgs {
  const float* views;  // V volumes of n[0]*n[1]*n[2]
  float* quality;      // V * nb[0]*nb[1]*nb[2]
  int V;
  int n[3];            // volume extent z, y, x
  int bs[3];           // block size z, y, x (<= 32)
  int nb[3];           // blocks per axis
  float r_o;           // OTF support radius (L1 index), < 0: all coefficients (L1-mean mode)
  float exponent;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  // 256 threads; red: 8 floats
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}
__device__ __forceinline__ float block_min(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = INFINITY;
  for (int w = 0; w < 8; ++w) s = fminf(s, red[w]);
  return s;
}

__global__ void __launch_bounds__(256)
dct_quality_kernel(DctArgs a) {
  extern __shared__ float sm[];
  float* buf = sm;                                     // [32][32][33]
  float* ctab = sm + kDctMax * kDctMax * kDctPitch;    // [3][32][32]: ctab[ax][k][n]
  __shared__ float red[8];

  int b = blockIdx.x;
  const int cx = b % a.nb[2]; b /= a.nb[2];
  const int cy = b % a.nb[1]; b /= a.nb[1];
  const int cz = b % a.nb[0];
  const int v = b / a.nb[0];
  const int z0 = cz * a.bs[0], y0 = cy * a.bs[1], x0 = cx * a.bs[2];
  const int ez = min(a.bs[0], a.n[0] - z0), ey = min(a.bs[1], a.n[1] - y0), ex = min(a.bs[2], a.n[2] - x0);
  const int cnt = ez * ey * ex;
  const float* src = a.views + (int64_t)v * a.n[0] * a.n[1] * a.n[2];
  const int64_t sy = a.n[2], sz = (int64_t)a.n[1] * a.n[2];
  float* qout = a.quality + (int64_t)blockIdx.x;

  // ---- NaN statistics: valid count, minimum of the valid voxels ----
  float nvalid = 0.f, vmin = INFINITY;
  for (int i = threadIdx.x; i < cnt; i += 256) {
    const int x = i % ex, y = (i / ex) % ey, z = i / (ex * ey);
    const float t = __ldg(src + (z0 + z) * sz + (y0 + y) * sy + (x0 + x));
    if (t == t) { nvalid += 1.f; vmin = fminf(vmin, t); }
  }
  nvalid = block_sum(nvalid, red);
  vmin = block_min(vmin, red);
  if (nvalid < 0.2f * (float)cnt) {  // mostly invalid: score stays 0 (weights.py:205-207)
    if (threadIdx.x == 0) *qout = 0.f;
    return;
  }
  const float fill = vmin > 0.0001f ? vmin : 0.0f;

  // ---- stage the block (NaN-filled), L2 norm, cosine tables ----
  float ss = 0.f;
  for (int i = threadIdx.x; i < cnt; i += 256) {
    const int x = i % ex, y = (i / ex) % ey, z = i / (ex * ey);
    float t = __ldg(src + (z0 + z) * sz + (y0 + y) * sy + (x0 + x));
    if (t != t) t = fill;
    buf[(z * kDctMax + y) * kDctPitch + x] = t;
    ss = fmaf(t, t, ss);
  }
  for (int i = threadIdx.x; i < 3 * kDctMax * kDctMax; i += 256) {
    const int ax = i / (kDctMax * kDctMax), k = (i / kDctMax) % kDctMax, n = i % kDctMax;
    const int N = ax == 0 ? ez : (ax == 1 ? ey : ex);
    float c = 0.f;
    if (k < N && n < N) {
      // orthonormal DCT-II: s_k cos(pi (2n+1) k / (2N)), s_0 = sqrt(1/N), s_k = sqrt(2/N)
      c = cospif((float)((2 * n + 1) * k) / (float)(2 * N)) * sqrtf((k == 0 ? 1.f : 2.f) / (float)N);
    }
    ctab[i] = c;
  }
  ss = block_sum(ss, red);  // (contains the __syncthreads the staging needs)
  const bool otf = a.r_o >= 0.f;
  const float l2 = sqrtf(ss);
  if (otf && l2 == 0.f) {
    if (threadIdx.x == 0) *qout = 0.f;
    return;
  }
  // coefficients needed per axis: index sums < r_o
  const int KX = otf ? min(ex, (int)ceilf(a.r_o)) : ex;
  const int KY = otf ? min(ey, (int)ceilf(a.r_o)) : ey;
  const int KZ = otf ? min(ez, (int)ceilf(a.r_o)) : ez;
  const float* cxt = ctab + 2 * kDctMax * kDctMax;
  const float* cyt = ctab + 1 * kDctMax * kDctMax;
  const float* czt = ctab;
  float line[kDctMax];

  // ---- along x: one thread per (z, y) row, in place ----
  for (int r = threadIdx.x; r < ez * ey; r += 256) {
    float* row = buf + ((r / ey) * kDctMax + (r % ey)) * kDctPitch;
#pragma unroll
    for (int n = 0; n < kDctMax; ++n) line[n] = n < ex ? row[n] : 0.f;
    for (int k = 0; k < KX; ++k) {
      float s = 0.f;
#pragma unroll
      for (int n = 0; n < kDctMax; ++n) s = fmaf(cxt[k * kDctMax + n], line[n], s);
      row[k] = s;
    }
  }
  __syncthreads();
  // ---- along y: one thread per (z, kx) column ----
  for (int r = threadIdx.x; r < ez * KX; r += 256) {
    const int z = r / KX, kx = r % KX;
    float* col = buf + (z * kDctMax) * kDctPitch + kx;
#pragma unroll
    for (int n = 0; n < kDctMax; ++n) line[n] = n < ey ? col[n * kDctPitch] : 0.f;
    for (int k = 0; k < KY; ++k) {
      if (otf && (float)(kx + k) >= a.r_o) break;
      float s = 0.f;
#pragma unroll
      for (int n = 0; n < kDctMax; ++n) s = fmaf(cyt[k * kDctMax + n], line[n], s);
      col[k * kDctPitch] = s;
    }
  }
  __syncthreads();
  // ---- along z: one thread per (ky, kx) column; OTF mode accumulates the entropy directly ----
  float h = 0.f, l1 = 0.f;
  for (int r = threadIdx.x; r < KY * KX; r += 256) {
    const int ky = r / KX, kx = r % KX;
    if (otf && (float)(kx + ky) >= a.r_o) continue;
    float* col = buf + ky * kDctPitch + kx;
#pragma unroll
    for (int n = 0; n < kDctMax; ++n) line[n] = n < ez ? col[n * kDctMax * kDctPitch] : 0.f;
    for (int k = 0; k < KZ; ++k) {
      if (otf && (float)(kx + ky + k) >= a.r_o) break;
      float s = 0.f;
#pragma unroll
      for (int n = 0; n < kDctMax; ++n) s = fmaf(czt[k * kDctMax + n], line[n], s);
      if (otf) {
        const float p = fabsf(s) / l2;
        if (p > 0.f) h -= p * log2f(p);
      } else {
        col[k * kDctMax * kDctPitch] = fabsf(s);
        l1 += fabsf(s);
      }
    }
  }
  float q;
  if (otf) {
    h = block_sum(h, red);
    q = (2.0f / (a.r_o * a.r_o)) * h;
    q = q > 0.f ? powf(q, a.exponent) : (q < 0.f ? -powf(-q, a.exponent) : 0.f);
  } else {
    // L1-mean normalisation over ALL coefficients (weights.py:232-243)
    l1 = block_sum(l1, red);
    const float dsl1 = l1 / (float)cnt;
    if (dsl1 == 0.f) {
      if (threadIdx.x == 0) *qout = 0.f;
      return;
    }
    for (int i = threadIdx.x; i < cnt; i += 256) {
      const int x = i % ex, y = (i / ex) % ey, z = i / (ex * ey);
      const float p = buf[(z * kDctMax + y) * kDctPitch + x] / dsl1;
      if (p > 0.f) h -= p * log2f(p);
    }
    h = block_sum(h, red);
    q = powf(dsl1 * h, a.exponent);
  }
  if (threadIdx.x == 0) *qout = q;
}

// quality[v][b] <- (q - min_v q) / sum_v(q - min_v q)   (sum == 0 -> 1; weights.py:246-248)
__global__ void dct_qnorm_kernel(float* q, int V, int64_t nblocks) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  float mn = INFINITY;
  for (int v = 0; v < V; ++v) {
    const float t = q[v * nblocks + b];
    if (t == t) mn = fminf(mn, t);  // nanmin
  }
  float s = 0.f;
  for (int v = 0; v < V; ++v) {
    const float t = q[v * nblocks + b] - mn;
    q[v * nblocks + b] = t;
    if (t == t) s += t;  // nansum
  }
  if (s == 0.f) s = 1.f;
  for (int v = 0; v < V; ++v) q[v * nblocks + b] = q[v * nblocks + b] / s;
}

// weights[v][voxel] = multilinear(quality[v]) at q = p / ds - (ds - 1) / (2 ds), mode "nearest",
// then normalised over v (weights.py:250-272)
__global__ void __launch_bounds__(256)
dct_weights_kernel(const float* __restrict__ q, int V, int nz, int ny, int nx, int bz, int by, int bx,
                   int nbz, int nby, int nbx, float* __restrict__ out) {
  const int64_t N = (int64_t)nz * ny * nx;
  const int64_t nblocks = (int64_t)nbz * nby * nbx;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((int64_t)nx * ny));
    int i0[3], i1[3];
    float t[3];
    const int p[3] = {z, y, x}, ds[3] = {bz, by, bx}, nb[3] = {nbz, nby, nbx};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      // scipy: matrix * p + offset in float64
      double c = __dadd_rn(__dmul_rn(1.0 / (double)ds[d], (double)p[d]), -((double)(ds[d] - 1)) / (2.0 * (double)ds[d]));
      c = fmin(fmax(c, 0.0), (double)(nb[d] - 1));  // mode "nearest"
      const double f = floor(c);
      i0[d] = (int)f;
      i1[d] = min(i0[d] + 1, nb[d] - 1);
      t[d] = (float)(c - f);
    }
    float sum = 0.f;
    float w[32];
    for (int v0 = 0; v0 < V; v0 += 32) {
      const int vn = min(32, V - v0);
      for (int k = 0; k < vn; ++k) {
        const float* m = q + (int64_t)(v0 + k) * nblocks;
        auto at = [&](int a, int b_, int c) { return __ldg(m + ((int64_t)a * nby + b_) * nbx + c); };
        const float a00 = fmaf(t[2], at(i0[0], i0[1], i1[2]) - at(i0[0], i0[1], i0[2]), at(i0[0], i0[1], i0[2]));
        const float a01 = fmaf(t[2], at(i0[0], i1[1], i1[2]) - at(i0[0], i1[1], i0[2]), at(i0[0], i1[1], i0[2]));
        const float a10 = fmaf(t[2], at(i1[0], i0[1], i1[2]) - at(i1[0], i0[1], i0[2]), at(i1[0], i0[1], i0[2]));
        const float a11 = fmaf(t[2], at(i1[0], i1[1], i1[2]) - at(i1[0], i1[1], i0[2]), at(i1[0], i1[1], i0[2]));
        const float b0 = fmaf(t[1], a01 - a00, a00), b1 = fmaf(t[1], a11 - a10, a10);
        w[k] = fmaf(t[0], b1 - b0, b0);
        sum += w[k];
      }
      if (V <= 32) break;
      // more than 32 views: two passes (sum first)
      for (int k = 0; k < vn; ++k) out[(int64_t)(v0 + k) * N + i] = w[k];
    }
    if (V <= 32) {
      const float s = sum == 0.f ? 1.f : sum;
      for (int k = 0; k < V; ++k) out[(int64_t)k * N + i] = w[k] / s;
    } else {
      const float s = sum == 0.f ? 1.f : sum;
      for (int k = 0; k < V; ++k) out[(int64_t)k * N + i] = out[(int64_t)k * N + i] / s;
    }
  }
}

}  // namespace mvs

extern "C" int mvs_content_based_dct(const float* d_views, int V, const int32_t shape[3], int ndim,
                                     const int32_t block[3], float r_o, float exponent,
                                     float* d_out_weights, void* stream) {
  using namespace mvs;
  MVS_REQUIRE(d_views && shape && block && d_out_weights, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(V >= 1, MVS_ERR_INVALID, "V = %d", V);
  MVS_REQUIRE(ndim == 2 || ndim == 3, MVS_ERR_INVALID, "ndim must be 2 or 3");
  DctArgs a;
  a.views = d_views;
  a.V = V;
  a.r_o = r_o;
  a.exponent = exponent;
  int64_t nblocks = 1;
  for (int d = 0; d < 3; ++d) {
    MVS_REQUIRE(shape[d] >= 1 && block[d] >= 1, MVS_ERR_INVALID, "empty extent");
    MVS_REQUIRE(block[d] <= kDctMax, MVS_ERR_UNSUPPORTED, "dct_size %d > %d", block[d], kDctMax);
    a.n[d] = shape[d];
    a.bs[d] = std::min(block[d], shape[d]);
    a.nb[d] = (shape[d] + a.bs[d] - 1) / a.bs[d];
    nblocks *= a.nb[d];
  }
  MVS_REQUIRE(ndim == 3 || shape[0] == 1, MVS_ERR_INVALID, "2-D stack with z extent %d", shape[0]);
  MVS_REQUIRE(nblocks * V < (1ll << 31), MVS_ERR_UNSUPPORTED, "too many blocks");
  cudaStream_t st = (cudaStream_t)stream;
  float* d_q = nullptr;
  MVS_CHECK_CUDA(mvs::pool_malloc((void**)&d_q, sizeof(float) * nblocks * V, st));
  a.quality = d_q;
  const size_t smem = sizeof(float) * (kDctMax * kDctMax * kDctPitch + 3 * kDctMax * kDctMax);
  MVS_CHECK_CUDA(cudaFuncSetAttribute(dct_quality_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dct_quality_kernel<<<(unsigned)(nblocks * V), 256, smem, st>>>(a);
  MVS_CHECK_CUDA(cudaGetLastError());
  dct_qnorm_kernel<<<(unsigned)((nblocks + 255) / 256), 256, 0, st>>>(d_q, V, nblocks);
  MVS_CHECK_CUDA(cudaGetLastError());
  const int64_t N = (int64_t)shape[0] * shape[1] * shape[2];
  const unsigned grid = (unsigned)std::min<int64_t>((N + 255) / 256, 148 * 16);
  dct_weights_kernel<<<grid, 256, 0, st>>>(d_q, V, a.n[0], a.n[1], a.n[2], a.bs[0], a.bs[1], a.bs[2], a.nb[0],
                                           a.nb[1], a.nb[2], d_out_weights);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(d_q, st);
  if (e != cudaSuccess) { set_error("dct weights launch: %s", cudaGetErrorString(e)); return MVS_ERR_CUDA; }
  return MVS_OK;
}
