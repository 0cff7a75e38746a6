// Candidate disambiguation for phase-correlation registration (sm_100a):
// the loop of registration.phase_correlation_registration over translation
// candidates (registration.py:493-556), batched over (pair, candidate):
//   im1t   = scipy.ndimage.affine_transform(im1, translation, order=1,
//            mode="constant", cval=NaN)                                  (:494-500)
//   mask   = ~isnan(im1t) & ~isnan(im0), its count, bbox of ~isnan(im1t) (:501-528)
//   SSIM   = skimage.metrics.structural_similarity on the bbox slices    (:535-548)
//   quality= scipy.stats.spearmanr(im0[mask], im1t[mask])                (:551-553)
// im1t is recomputed on the fly wherever it is needed (never stored), with
// scipy's float64 tap arithmetic so masks and values are bit-identical.

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <vector>

#include "common.cuh"

struct mvs_pc_plan;

namespace mvs {
const float* pc_r0(const mvs_pc_plan* p, int pair);
const float* pc_r1(const mvs_pc_plan* p, int pair);
int pc_ndim(const mvs_pc_plan* p);
const int* pc_shape(const mvs_pc_plan* p);
int pc_loaded(const mvs_pc_plan* p);
int pc_scratch(mvs_pc_plan* p, size_t bytes, void** out);

struct Cand {
  const float* r0;
  const float* r1;
  double t[3];
  int lo[3];   // slice start (SSIM)
  int len[3];  // slice extent
  int win;
  int tiles[3];
  long long tile_base;  // first tile index of this candidate in the flat tile list
  long long mat_off;    // offset of this candidate's materialised im1t[slices]
};

// scipy's mirrored tap index for mode="constant" splines (ni_interpolation.c)
__device__ __forceinline__ int mirror_idx(int idx, int len) {
  if (len <= 1) return 0;
  const int s2 = 2 * len - 2;
  if (idx < 0) {
    idx = s2 * (-idx / s2) + idx;
    return idx <= 1 - len ? idx + s2 : -idx;
  }
  if (idx >= len) {
    idx -= s2 * (idx / s2);
    if (idx >= len) idx = s2 - idx;
  }
  return idx;
}

// im1t at integer voxel (z, y, x): order-1 spline of r1 at (o + t), NaN outside
// [0, len-1]; float64 tap products in scipy's order, result rounded to float32.
template <int NDIM>
__device__ __forceinline__ float shifted_value(const float* __restrict__ r1, int n0, int n1,
                                               int n2, const double* t, int z, int y, int x) {
  const double cx = (double)x + t[2], cy = (double)y + t[1];
  const double cz = NDIM == 3 ? (double)z + t[0] : 0.0;
  if (cx < 0.0 || cx > (double)(n2 - 1) || cy < 0.0 || cy > (double)(n1 - 1)) return NAN;
  if (NDIM == 3 && (cz < 0.0 || cz > (double)(n0 - 1))) return NAN;
  const double fx = floor(cx), fy = floor(cy), fz = floor(cz);
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  double wx[2], wy[2], wz[2];
  wx[0] = 1.0 - (cx - fx); wx[1] = 1.0 - wx[0];
  wy[0] = 1.0 - (cy - fy); wy[1] = 1.0 - wy[0];
  wz[0] = 1.0 - (cz - fz); wz[1] = 1.0 - wz[0];
  int xs[2] = {ix, ix + 1 < n2 ? ix + 1 : mirror_idx(ix + 1, n2)};
  int ys[2] = {iy, iy + 1 < n1 ? iy + 1 : mirror_idx(iy + 1, n1)};
  int zs[2] = {iz, iz + 1 < n0 ? iz + 1 : mirror_idx(iz + 1, n0)};
  double acc = 0.0;
  if (NDIM == 3) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double v = (double)__ldg(r1 + ((long long)zs[a] * n1 + ys[b]) * n2 + xs[c]);
          v = __dmul_rn(__dmul_rn(__dmul_rn(v, wz[a]), wy[b]), wx[c]);
          acc = __dadd_rn(acc, v);
        }
  } else {
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double v = (double)__ldg(r1 + (long long)ys[b] * n2 + xs[c]);
        v = __dmul_rn(__dmul_rn(v, wy[b]), wx[c]);
        acc = __dadd_rn(acc, v);
      }
  }
  return (float)acc;
}

// The same value with the (z, y) part of the work hoisted: a warp that walks along x keeps the row's
// validity, tap rows and weights in registers (identical arithmetic, so identical results).
constexpr unsigned kRowSeg = 256;  // voxels of a row one warp walks with the hoisted (z, y) part

template <int NDIM>
struct ShiftRow {
  bool valid;
  long long base[4];  // offsets of the (z tap, y tap) rows: [a * 2 + b]
  double wzy[4][2];   // wz[a], wy[b] per row tap
  __device__ __forceinline__ void init(int n0, int n1, int n2, const double* t, int z, int y) {
    const double cy = (double)y + t[1];
    const double cz = NDIM == 3 ? (double)z + t[0] : 0.0;
    valid = !(cy < 0.0 || cy > (double)(n1 - 1)) && !(NDIM == 3 && (cz < 0.0 || cz > (double)(n0 - 1)));
    const double fy = floor(cy), fz = floor(cz);
    const int iy = (int)fy, iz = (int)fz;
    double wy[2], wz[2];
    wy[0] = 1.0 - (cy - fy); wy[1] = 1.0 - wy[0];
    wz[0] = 1.0 - (cz - fz); wz[1] = 1.0 - wz[0];
    const int ys[2] = {iy, iy + 1 < n1 ? iy + 1 : mirror_idx(iy + 1, n1)};
    const int zs[2] = {iz, iz + 1 < n0 ? iz + 1 : mirror_idx(iz + 1, n0)};
#pragma unroll
    for (int a = 0; a < (NDIM == 3 ? 2 : 1); ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        base[a * 2 + b] = valid ? ((long long)(NDIM == 3 ? zs[a] : 0) * n1 + ys[b]) * n2 : 0;
        wzy[a * 2 + b][0] = wz[a];
        wzy[a * 2 + b][1] = wy[b];
      }
  }
  __device__ __forceinline__ float at(const float* __restrict__ r1, int n2, const double* t, int x) const {
    const double cx = (double)x + t[2];
    if (!valid || cx < 0.0 || cx > (double)(n2 - 1)) return NAN;
    const double fx = floor(cx);
    const int ix = (int)fx;
    double wx[2];
    wx[0] = 1.0 - (cx - fx); wx[1] = 1.0 - wx[0];
    const int xs[2] = {ix, ix + 1 < n2 ? ix + 1 : mirror_idx(ix + 1, n2)};
    double acc = 0.0;
#pragma unroll
    for (int ab = 0; ab < (NDIM == 3 ? 4 : 2); ++ab)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double v = (double)__ldg(r1 + base[ab] + xs[c]);
        if (NDIM == 3) v = __dmul_rn(v, wzy[ab][0]);
        v = __dmul_rn(__dmul_rn(v, wzy[ab][1]), wx[c]);
        acc = __dadd_rn(acc, v);
      }
    return (float)acc;
  }
};

// ---- stage C: mask statistics ------------------------------------------------

constexpr int kStatBlocks = 32;

template <int NDIM>
__global__ void __launch_bounds__(256)
cand_stats_kernel(const Cand* __restrict__ cands, int n0, int n1, int n2,
                  long long* __restrict__ partial /* [cand][block][8] */) {
  const Cand c = cands[blockIdx.y];
  const long long N = (long long)n0 * n1 * n2;
  long long nmask = 0, nvalid = 0;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {-1, -1, -1};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % n2), y = (int)((i / n2) % n1), z = (int)(i / ((long long)n1 * n2));
    float v = shifted_value<NDIM>(c.r1, n0, n1, n2, c.t, z, y, x);
    if (v == v) {
      ++nvalid;
      lo[0] = min(lo[0], z); lo[1] = min(lo[1], y); lo[2] = min(lo[2], x);
      hi[0] = max(hi[0], z); hi[1] = max(hi[1], y); hi[2] = max(hi[2], x);
      float f = __ldg(c.r0 + i);
      if (f == f) ++nmask;
    }
  }
  __shared__ long long s_cnt[2][256];
  __shared__ int s_lo[3][256], s_hi[3][256];
  const int t = threadIdx.x;
  s_cnt[0][t] = nmask; s_cnt[1][t] = nvalid;
  for (int d = 0; d < 3; ++d) { s_lo[d][t] = lo[d]; s_hi[d][t] = hi[d]; }
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (t < s) {
      s_cnt[0][t] += s_cnt[0][t + s]; s_cnt[1][t] += s_cnt[1][t + s];
      for (int d = 0; d < 3; ++d) {
        s_lo[d][t] = min(s_lo[d][t], s_lo[d][t + s]);
        s_hi[d][t] = max(s_hi[d][t], s_hi[d][t + s]);
      }
    }
    __syncthreads();
  }
  if (t == 0) {
    long long* p = partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 8;
    p[0] = s_cnt[0][0]; p[1] = s_cnt[1][0];
    for (int d = 0; d < 3; ++d) { p[2 + d] = s_lo[d][0]; p[5 + d] = s_hi[d][0]; }
  }
}

// ---- stage D: SSIM -------------------------------------------------------------
// Uniform win^ndim window, sample covariance, K1 = 0.01, K2 = 0.03, data_range 1
// (images are rescaled to [0,1]); scipy.ndimage.uniform_filter semantics: one
// pass per axis (z, y, x), float64 line sums, float32 between passes.
// im1t is materialised once per candidate (slice region only) by
// materialize_kernel; the tiles then read plain float32 windows.

// Tiles march along the first filter axis (z in 3-D, y in 2-D) with float64
// running sums per column, so the halo of that axis is paid once per tile and
// the per-voxel work does not grow with the window:
//   2-D: a CTA owns up to 128 output columns x 64 output rows; thread = column.
//   3-D: a CTA owns 32 x 8 output columns x 32 output planes; thread = up to 3
//        of the (32+6) x (8+6) input columns; each plane then gets its y and x
//        passes through shared memory.
constexpr int kMaxWin = 7;
// 2-D tile width (knob).  Measured: 150 output columns (156 of 160 threads busy, 2 instead of 3 tiles across
// C2's 305-px slices) is no faster than 128 (1145 vs 1108 us per 20-pair batch): the kernel is bound by the
// per-row barrier / shared-memory latency chain, not by lane utilisation
#ifndef MVS_S2_TX
#define MVS_S2_TX 128
#endif
// What bounds the kernel (ncu, profiles/r02_ssim2d_ncu.txt): the XU pipe at 62 % -- the 15 float32 -> float64 and
// 10 float64 -> float32 conversions per thread and row that keep the window sums exact run at a quarter of the
// ALU rate.  Measured and neutral: tile shapes 64..150 x 32..64 (1110-1165 us per 20-pair batch), two outputs per
// thread in the x pass with 128-bit shared loads (14 -> 4 loads per output pair: 1132 us), occupancy hints.
constexpr int kS2TX = MVS_S2_TX, kS2TY = 64, kS2Threads = 160, kS2Cols = kS2TX + kMaxWin - 1;
static_assert(kS2Cols <= kS2Threads, "one thread per input column");
#ifndef MVS_S3_THREADS
#define MVS_S3_THREADS 288  // 9 warps: the (32+6) x (8+6) = 532 input columns of a tile take 2 per thread (3 with 256)
#endif
constexpr int kS3TX = 32, kS3TY = 8, kS3TZ = 32, kS3Threads = MVS_S3_THREADS;
constexpr int kS3WX = kS3TX + kMaxWin - 1, kS3WY = kS3TY + kMaxWin - 1, kS3P = kS3WX + 1;

// im1t[slices] of every candidate, packed one after the other (offset mat_off)
template <int NDIM>
__global__ void __launch_bounds__(256)
materialize_kernel(const Cand* __restrict__ cands, int n0, int n1, int n2,
                   float* __restrict__ mat) {
  const Cand c = cands[blockIdx.y];
  const long long n = (long long)c.len[0] * c.len[1] * c.len[2];
  float* out = mat + c.mat_off;
  // one warp per row of the slice (32-bit index arithmetic; the row's z / y taps and weights stay in
  // registers while the lanes walk along x)
  const unsigned rows = (unsigned)c.len[0] * (unsigned)c.len[1];
  const unsigned leny = (unsigned)c.len[1];
  const int lenx = c.len[2];
  const unsigned nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
  const unsigned xt = ((unsigned)lenx + kRowSeg - 1) / kRowSeg;  // x segments per row: the units of work
  for (unsigned u = blockIdx.x * nwarps + (threadIdx.x >> 5); u < rows * xt; u += gridDim.x * nwarps) {
    const unsigned row = u / xt, x0 = (u - row * xt) * kRowSeg;
    const unsigned z = row / leny, y = row - z * leny;
    ShiftRow<NDIM> R;
    R.init(n0, n1, n2, c.t, c.lo[0] + (int)z, c.lo[1] + (int)y);
    float* orow = out + (long long)row * lenx;
    const int x1 = min(lenx, (int)(x0 + kRowSeg));
    for (int x = (int)(x0 + lane); x < x1; x += 32) orow[x] = R.at(c.r1, n2, c.t, c.lo[2] + x);
  }
}

// float32 SSIM of one window from its five float32 means (skimage
// structural_similarity, gaussian_weights=False, use_sample_covariance=True)
__device__ __forceinline__ float ssim_value(float ux, float uy, float uxx, float uyy, float uxy,
                                            float cov_norm) {
  const float C1 = __fmul_rn(0.01f, 0.01f);  // (K1 * R)^2 with R = 1 in float32
  const float C2 = __fmul_rn(0.03f, 0.03f);
  const float vx = __fmul_rn(cov_norm, __fsub_rn(uxx, __fmul_rn(ux, ux)));
  const float vy = __fmul_rn(cov_norm, __fsub_rn(uyy, __fmul_rn(uy, uy)));
  const float vxy = __fmul_rn(cov_norm, __fsub_rn(uxy, __fmul_rn(ux, uy)));
  const float A1 = __fadd_rn(__fmul_rn(__fmul_rn(2.f, ux), uy), C1);
  const float A2 = __fadd_rn(__fmul_rn(2.f, vxy), C2);
  const float B1 = __fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), C1);
  const float B2 = __fadd_rn(__fadd_rn(vx, vy), C2);
  return __fdiv_rn(__fmul_rn(A1, A2), __fmul_rn(B1, B2));
}

__device__ __forceinline__ int find_cand(const Cand* __restrict__ cands, int n_cand, long long tile) {
  int ci = 0, chi = n_cand - 1;
  while (ci < chi) {
    const int mid = (ci + chi + 1) >> 1;
    if (cands[mid].tile_base <= tile) ci = mid; else chi = mid - 1;
  }
  return ci;
}

// the five filtered quantities a, b, aa, bb, ab (products in float32 like skimage)
__device__ __forceinline__ void five(float a, float b, double* q) {
  q[0] = (double)a; q[1] = (double)b;
  q[2] = (double)__fmul_rn(a, a); q[3] = (double)__fmul_rn(b, b); q[4] = (double)__fmul_rn(a, b);
}

template <int THREADS>
__device__ __forceinline__ void ssim_block_reduce(double sum, float vmax, long long tile,
                                                  double* __restrict__ tile_sum,
                                                  float* __restrict__ tile_max) {
  __shared__ double s_red[THREADS / 32];
  __shared__ float s_max[THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = sum; s_max[threadIdx.x >> 5] = vmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    float m = -INFINITY;
    for (int i = 0; i < THREADS / 32; ++i) { s += s_red[i]; m = fmaxf(m, s_max[i]); }
    tile_sum[tile] = s;
    tile_max[tile] = m;
  }
}

// (occupancy hints measured on the B200: minBlocks 8 / 10 -> 1295 / 1438 us against 1108 us for the
// compiler's own choice of 56 registers)
template <int WIN>
__global__ void __launch_bounds__(kS2Threads)
ssim2d_kernel(const Cand* __restrict__ cands, int n_cand, int n1, int n2,
              const float* __restrict__ mat, double* __restrict__ tile_sum,
              float* __restrict__ tile_max) {
  __shared__ double rowbuf[2][5][kS2Cols + 2];
  const long long tile = blockIdx.x;
  const Cand c = cands[find_cand(cands, n_cand, tile)];
  const long long local = tile - c.tile_base;
  const int tx_i = (int)(local % c.tiles[2]), ty_i = (int)(local / c.tiles[2]);
  const int ox = tx_i * kS2TX, oy = ty_i * kS2TY;  // output origin in the slice
  constexpr int win = WIN;  // all candidates of a launch share the window
  const int leny = c.len[1], lenx = c.len[2];
  const int nxo = min(kS2TX, lenx - win + 1 - ox), nyo = min(kS2TY, leny - win + 1 - oy);
  const int ncols = nxo + win - 1;
  const int t = threadIdx.x;
  const bool col_ok = t < ncols;
  const float* pa = c.r0 + (long long)(c.lo[1] + oy) * n2 + c.lo[2] + ox + (col_ok ? t : 0);
  const float* pb = mat + c.mat_off + (long long)oy * lenx + ox + (col_ok ? t : 0);
  const double inv = 1.0 / (double)win;
  const int np = win * win;
  const float cov_norm = (float)((double)np / (double)(np - 1));
  double S[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  double sum = 0.0;
  float vmax = -INFINITY;
  const int nrows = nyo + win - 1;
  // software pipeline: the loads of row r + 1 are in flight while row r is processed
  float na = 0.f, nb = 0.f, oa = 0.f, ob = 0.f;
  if (col_ok) { na = __ldg(pa); nb = __ldg(pb); }
  for (int r = 0; r < nrows; ++r) {
    float a = na, b = nb, a0 = oa, b0 = ob;
    if (col_ok && r + 1 < nrows) {
      na = __ldg(pa + (long long)(r + 1) * n2);
      nb = __ldg(pb + (long long)(r + 1) * lenx);
      if (r + 1 >= win) {
        oa = __ldg(pa + (long long)(r + 1 - win) * n2);
        ob = __ldg(pb + (long long)(r + 1 - win) * lenx);
      }
    }
    if (col_ok) {
      if (b == b) vmax = fmaxf(vmax, b); else b = 0.f;
      if (a != a) a = 0.f;
      double qn[5];
      five(a, b, qn);
      if (r >= win) {
        if (b0 != b0) b0 = 0.f;
        if (a0 != a0) a0 = 0.f;
        double qo[5];
        five(a0, b0, qo);
#pragma unroll
        for (int q = 0; q < 5; ++q) S[q] += qn[q] - qo[q];
      } else {
#pragma unroll
        for (int q = 0; q < 5; ++q) S[q] += qn[q];
      }
    }
    if (r >= win - 1) {
      double (*buf)[kS2Cols + 2] = rowbuf[r & 1];
      if (col_ok) {
#pragma unroll
        for (int q = 0; q < 5; ++q) buf[q][t] = (double)(float)(S[q] * inv);  // float32 between passes
      }
      __syncthreads();
      if (t < nxo) {
        float U[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < win; ++k) s += buf[q][t + k];
          U[q] = (float)(s * inv);
        }
        sum += (double)ssim_value(U[0], U[1], U[2], U[3], U[4], cov_norm);
      }
    }
  }
  ssim_block_reduce<kS2Threads>(sum, vmax, tile, tile_sum, tile_max);
}

// 3 CTAs of 9 warps per SM (72 registers, no spills): 12 C3 face pairs 67 -> 48 ms against 2 CTAs at 96
#ifndef MVS_S3_MINB
#define MVS_S3_MINB 3
#endif
template <int WIN>
__global__ void __launch_bounds__(kS3Threads, MVS_S3_MINB)
ssim3d_kernel(const Cand* __restrict__ cands, int n_cand, int n1, int n2,
              const float* __restrict__ mat, double* __restrict__ tile_sum,
              float* __restrict__ tile_max) {
  // float32-rounded values kept as doubles (no conversions in the running sums)
  __shared__ double zf[5][kS3WY][kS3P];  // z-filtered plane
  __shared__ double yf[5][kS3TY][kS3P];  // then y-filtered
  __shared__ float xf[5][kS3TY][kS3TX + 1];  // then x-filtered: the five means per output
  const long long tile = blockIdx.x;
  const Cand c = cands[find_cand(cands, n_cand, tile)];
  long long local = tile - c.tile_base;
  const int tx_i = (int)(local % c.tiles[2]);
  const int ty_i = (int)((local / c.tiles[2]) % c.tiles[1]);
  const int tz_i = (int)(local / ((long long)c.tiles[2] * c.tiles[1]));
  const int ox = tx_i * kS3TX, oy = ty_i * kS3TY, oz = tz_i * kS3TZ;
  constexpr int win = WIN;  // all candidates of a launch share the window
  const int lenz = c.len[0], leny = c.len[1], lenx = c.len[2];
  const int nxo = min(kS3TX, lenx - win + 1 - ox), nyo = min(kS3TY, leny - win + 1 - oy);
  const int nzo = min(kS3TZ, lenz - win + 1 - oz);
  const int t = threadIdx.x;
  const double inv = 1.0 / (double)win;
  const int np = win * win * win;
  const float cov_norm = (float)((double)np / (double)(np - 1));
  constexpr int NCOL = kS3WX * kS3WY;                       // input columns of the tile
  constexpr int CPT = (NCOL + kS3Threads - 1) / kS3Threads;  // columns per thread
  int cy[CPT], cx[CPT];
  bool ok[CPT];
  long long offa[CPT], offb[CPT];
  double S[CPT][5];
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int col = t + i * kS3Threads;
    cy[i] = col / kS3WX; cx[i] = col - cy[i] * kS3WX;
    ok[i] = col < NCOL && oy + cy[i] < leny && ox + cx[i] < lenx;
    if (col >= NCOL) { cy[i] = 0; cx[i] = 0; }
    offa[i] = ok[i] ? ((long long)(c.lo[0] + oz) * n1 + c.lo[1] + oy + cy[i]) * n2 + c.lo[2] + ox + cx[i] : 0;
    offb[i] = ok[i] ? ((long long)oz * leny + oy + cy[i]) * lenx + ox + cx[i] : 0;
#pragma unroll
    for (int q = 0; q < 5; ++q) S[i][q] = 0.0;
  }
  const long long pla = (long long)n1 * n2, plb = (long long)leny * lenx;
  const float* ra = c.r0;
  const float* rb = mat + c.mat_off;
  double sum = 0.0;
  float vmax = -INFINITY;
  const int nplanes = nzo + win - 1;
  const int yo = t / kS3TX, xo = t - yo * kS3TX;  // this thread's output in the x pass
  // software pipeline: the loads of plane r + 1 are in flight while plane r is processed
  float na[CPT], nb[CPT], oa[CPT], ob[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    na[i] = ok[i] ? __ldg(ra + offa[i]) : 0.f;
    nb[i] = ok[i] ? __ldg(rb + offb[i]) : 0.f;
    oa[i] = 0.f; ob[i] = 0.f;
  }
  for (int r = 0; r < nplanes; ++r) {
    float ca[CPT], cb[CPT], c0a[CPT], c0b[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) { ca[i] = na[i]; cb[i] = nb[i]; c0a[i] = oa[i]; c0b[i] = ob[i]; }
    if (r + 1 < nplanes) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        if (!ok[i]) continue;
        na[i] = __ldg(ra + offa[i] + (r + 1) * pla);
        nb[i] = __ldg(rb + offb[i] + (r + 1) * plb);
        if (r + 1 >= win) {
          oa[i] = __ldg(ra + offa[i] + (r + 1 - win) * pla);
          ob[i] = __ldg(rb + offb[i] + (r + 1 - win) * plb);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      if (!ok[i]) continue;
      float a = ca[i], b = cb[i];
      if (b == b) vmax = fmaxf(vmax, b); else b = 0.f;
      if (a != a) a = 0.f;
      double qn[5];
      five(a, b, qn);
      if (r >= win) {
        float a0 = c0a[i], b0 = c0b[i];
        if (b0 != b0) b0 = 0.f;
        if (a0 != a0) a0 = 0.f;
        double qo[5];
        five(a0, b0, qo);
#pragma unroll
        for (int q = 0; q < 5; ++q) S[i][q] += qn[q] - qo[q];
      } else {
#pragma unroll
        for (int q = 0; q < 5; ++q) S[i][q] += qn[q];
      }
    }
    if (r < win - 1) continue;
    // z-filtered plane (float32 between passes); columns outside the slice hold 0
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      if (t + i * kS3Threads < NCOL) {
#pragma unroll
        for (int q = 0; q < 5; ++q) zf[q][cy[i]][cx[i]] = ok[i] ? (double)(float)(S[i][q] * inv) : 0.0;
      }
    }
    __syncthreads();
    // y pass: thread = (quantity, x column), a running window sum down the kS3TY outputs
    if (t < 5 * kS3WX) {
      const int q = t / kS3WX, x = t - q * kS3WX;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < win; ++k) s += zf[q][k][x];
#pragma unroll
      for (int y = 0; y < kS3TY; ++y) {
        yf[q][y][x] = (double)(float)(s * inv);
        if (y + 1 < kS3TY) s += zf[q][y + win][x] - zf[q][y][x];
      }
    }
    __syncthreads();
    // x pass: thread = (quantity, row, 8-output segment), a running window sum along x
    if (t < 5 * kS3TY * (kS3TX / 8)) {
      const int q = t / (kS3TY * (kS3TX / 8));
      const int r2 = t - q * (kS3TY * (kS3TX / 8));
      const int y = r2 / (kS3TX / 8), x0 = (r2 - y * (kS3TX / 8)) * 8;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < win; ++k) s += yf[q][y][x0 + k];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xf[q][y][x0 + j] = (float)(s * inv);
        if (j + 1 < 8) s += yf[q][y][x0 + j + win] - yf[q][y][x0 + j];
      }
    }
    __syncthreads();
    // SSIM of this thread's output voxel
    if (yo < nyo && xo < nxo)
      sum += (double)ssim_value(xf[0][yo][xo], xf[1][yo][xo], xf[2][yo][xo], xf[3][yo][xo], xf[4][yo][xo], cov_norm);
  }
  ssim_block_reduce<kS3Threads>(sum, vmax, tile, tile_sum, tile_max);
}

// ---- stage E: Spearman ---------------------------------------------------------
// Ranks of up to four pairs are computed with ONE radix sort per image: the key is
// (pair slot << 30) | 30 order-preserving bits of the float, so every pair occupies
// its own N-element segment of the sorted array.

// Both ranked images live in a known interval -- im0 is rescaled to [0, 1] and the reference ranks
// im1t - 1, i.e. [-1, 0] -- so an order-preserving key needs 30 bits: the float's bits for [0, 1],
// 0x3F800000 minus the bits of |v| for [-1, 0] (-0 and +0 coincide).  With the slot of a pair in the
// bits above, four pairs sort together on 32-bit keys (4 radix passes over 8 bytes per element
// instead of 5 over 12 with 64-bit keys).
constexpr unsigned kKeyOne = 0x3F800000u;     // bits of 1.0f: the largest valid key
constexpr unsigned kKeyMasked = 0x3FFFFFFFu;  // sorts behind every valid key of its slot
constexpr int kKeyBits = 30;
constexpr int kSortSlots = 4;

__device__ __forceinline__ unsigned key_unit(float v) {  // v in [0, 1]
  return min(__float_as_uint(v + 0.0f) & 0x7fffffffu, kKeyOne);
}
__device__ __forceinline__ unsigned key_neg_unit(float v) {  // v in [-1, 0]
  return kKeyOne - min(__float_as_uint(v) & 0x7fffffffu, kKeyOne);
}

// Stage E without scatters: sort (a-key, b-key) by a; the position in the sorted
// segment is a's rank, written (doubled, so tie averages stay integers) as the
// payload of a second sort by b; in b-sorted order both ranks are at hand and the
// Pearson sums stream out.
template <int NDIM>
__global__ void __launch_bounds__(256)
spearman_keys_kernel(const Cand* __restrict__ cands, int n0, int n1, int n2,
                     unsigned* __restrict__ ka, unsigned* __restrict__ vb) {
  const int slot = blockIdx.y;
  const Cand c = cands[slot];
  const long long N = (long long)n0 * n1 * n2;
  const unsigned hi = (unsigned)slot << kKeyBits;
  // one warp per row (see materialize_kernel)
  const unsigned rows = (unsigned)n0 * (unsigned)n1;
  const unsigned nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
  const unsigned xt = ((unsigned)n2 + kRowSeg - 1) / kRowSeg;
  for (unsigned u = blockIdx.x * nwarps + (threadIdx.x >> 5); u < rows * xt; u += gridDim.x * nwarps) {
    const unsigned row = u / xt, x0 = (u - row * xt) * kRowSeg;
    const unsigned zu = row / (unsigned)n1;
    ShiftRow<NDIM> R;
    R.init(n0, n1, n2, c.t, (int)zu, (int)(row - zu * (unsigned)n1));
    const int x1 = min(n2, (int)(x0 + kRowSeg));
    for (int x = (int)(x0 + lane); x < x1; x += 32) {
      const long long i = (long long)row * n2 + x;
      const float b = R.at(c.r1, n2, c.t, x);
      const float a = __ldg(c.r0 + i);
      const bool m = (a == a) && (b == b);
      // the reference ranks `im1t[mask] - 1` in float32 (registration.py:551-553):
      // the subtraction merges values below ~3e-8 into ties, which changes ranks
      const long long e = (long long)slot * N + i;
      ka[e] = hi | (m ? key_unit(a) : kKeyMasked);
      vb[e] = m ? key_neg_unit(__fsub_rn(b, 1.0f)) : kKeyMasked;
    }
  }
}

// Runs of equal keys (ties) without per-element searches: run heads / tails are
// flagged with their own index, an inclusive max-scan carries the head index
// forward, a min-scan over the reversed tail flags carries the tail index back.
// (Integer-valued microscopy data is tie-heavy; a binary search per element
// costs ~40 dependent L2 reads.)
__global__ void __launch_bounds__(256)
run_flags_kernel(const unsigned* __restrict__ sorted, long long E,
                 unsigned* __restrict__ head, unsigned* __restrict__ tail_rev) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < E;
       j += (long long)gridDim.x * blockDim.x) {
    const unsigned v = sorted[j];
    head[j] = (j == 0 || sorted[j - 1] != v) ? (unsigned)j : 0u;
    tail_rev[E - 1 - j] = (j == E - 1 || sorted[j + 1] != v) ? (unsigned)j : 0xffffffffu;
  }
}

struct MaxOp { __device__ __forceinline__ unsigned operator()(unsigned a, unsigned b) const { return a > b ? a : b; } };
struct MinOp { __device__ __forceinline__ unsigned operator()(unsigned a, unsigned b) const { return a < b ? a : b; } };

// a-sorted order -> keys / payload of the sort by b.  Payload = 2 * average rank of a
// (scipy.stats.rankdata "average": ties share the mean rank).
__global__ void __launch_bounds__(256)
rank_a_kernel(const unsigned* __restrict__ first, const unsigned* __restrict__ last_rev,
              const unsigned* __restrict__ vb_sorted, long long N, long long E,
              const long long* __restrict__ nmask, unsigned* __restrict__ kb,
              unsigned* __restrict__ ra2) {
  const int slot = blockIdx.y;
  const long long n = nmask[slot];
  const long long base = (long long)slot * N;
  const unsigned hi = (unsigned)slot << kKeyBits;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N;
       j += (long long)gridDim.x * blockDim.x) {
    const long long e = base + j;
    if (j < n) {
      const long long f = (long long)first[e] - base, l = (long long)last_rev[E - 1 - e] - base;  // inclusive
      kb[e] = hi | vb_sorted[e];
      ra2[e] = (unsigned)(f + l + 2);  // 2 * ((f + l) / 2 + 1)
    } else {
      kb[e] = hi | kKeyMasked;
      ra2[e] = 0u;
    }
  }
}

constexpr int kPearsonBlocks = 64;

// b-sorted order: rank of b from the position, rank of a from the payload
__global__ void __launch_bounds__(256)
pearson_sorted_kernel(const unsigned* __restrict__ first, const unsigned* __restrict__ last_rev,
                      const unsigned* __restrict__ ra2, long long N, long long E,
                      const long long* __restrict__ nmask,
                      double* __restrict__ partial /* [slot][block][3] of this sub-batch */) {
  const int slot = blockIdx.y;
  const long long n = nmask[slot];
  const double mean = 0.5 * (double)(n + 1);
  const long long base = (long long)slot * N;
  double sab = 0.0, saa = 0.0, sbb = 0.0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    const long long e = base + j;
    const long long f = (long long)first[e] - base, l = (long long)last_rev[E - 1 - e] - base;
    const double a = 0.5 * (double)ra2[e] - mean;
    const double b = 0.5 * (double)(f + l + 2) - mean;
    sab += a * b; saa += a * a; sbb += b * b;
  }
  __shared__ double s[3][256];
  const int t = threadIdx.x;
  s[0][t] = sab; s[1][t] = saa; s[2][t] = sbb;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (t < k) { s[0][t] += s[0][t + k]; s[1][t] += s[1][t + k]; s[2][t] += s[2][t + k]; }
    __syncthreads();
  }
  if (t == 0) {
    double* p = partial + ((long long)slot * gridDim.x + blockIdx.x) * 3;
    p[0] = s[0][0]; p[1] = s[1][0]; p[2] = s[2][0];
  }
}

}  // namespace mvs

using namespace mvs;

static int fill_cands(mvs_pc_plan* p, int n_cand, const int32_t* cand_pair, const double* cand_t,
                      std::vector<Cand>& out) {
  MVS_REQUIRE(p && cand_pair && cand_t, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(n_cand >= 1, MVS_ERR_INVALID, "n_cand = %d", n_cand);
  out.resize(n_cand);
  for (int i = 0; i < n_cand; ++i) {
    MVS_REQUIRE(cand_pair[i] >= 0 && cand_pair[i] < pc_loaded(p), MVS_ERR_INVALID,
                "candidate %d: pair %d not loaded", i, cand_pair[i]);
    Cand& c = out[i];
    memset(&c, 0, sizeof(c));
    c.r0 = pc_r0(p, cand_pair[i]);
    c.r1 = pc_r1(p, cand_pair[i]);
    for (int d = 0; d < 3; ++d) c.t[d] = cand_t[3 * i + d];
  }
  return MVS_OK;
}

extern "C" int mvs_pc_candidate_stats(mvs_pc_plan* p, int n_cand, const int32_t* cand_pair,
                                      const double* cand_t, int64_t* stats_host, void* stream) {
  MVS_REQUIRE(stats_host, MVS_ERR_INVALID, "NULL pointer");
  std::vector<Cand> cands;
  int rc = fill_cands(p, n_cand, cand_pair, cand_t, cands);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int* sh = pc_shape(p);
  const size_t cbytes = sizeof(Cand) * n_cand;
  const size_t pbytes = sizeof(long long) * 8 * kStatBlocks * n_cand;
  void* scratch;
  if ((rc = pc_scratch(p, cbytes + pbytes + 256, &scratch))) return rc;
  Cand* d_c = (Cand*)scratch;
  long long* d_p = (long long*)((char*)scratch + ((cbytes + 255) / 256) * 256);
  MVS_CHECK_CUDA(cudaMemcpyAsync(d_c, cands.data(), cbytes, cudaMemcpyHostToDevice, st));
  dim3 grid(kStatBlocks, n_cand);
  if (pc_ndim(p) == 3)
    cand_stats_kernel<3><<<grid, 256, 0, st>>>(d_c, sh[0], sh[1], sh[2], d_p);
  else
    cand_stats_kernel<2><<<grid, 256, 0, st>>>(d_c, sh[0], sh[1], sh[2], d_p);
  MVS_CHECK_CUDA(cudaGetLastError());
  std::vector<long long> part((size_t)8 * kStatBlocks * n_cand);
  MVS_CHECK_CUDA(cudaMemcpyAsync(part.data(), d_p, pbytes, cudaMemcpyDeviceToHost, st));
  MVS_CHECK_CUDA(cudaStreamSynchronize(st));
  for (int c = 0; c < n_cand; ++c) {
    int64_t* s = stats_host + (size_t)c * 8;
    s[0] = s[1] = 0;
    for (int d = 0; d < 3; ++d) { s[2 + d] = INT_MAX; s[5 + d] = -1; }
    for (int b = 0; b < kStatBlocks; ++b) {
      const long long* q = part.data() + ((size_t)c * kStatBlocks + b) * 8;
      s[0] += q[0]; s[1] += q[1];
      for (int d = 0; d < 3; ++d) {
        s[2 + d] = std::min<int64_t>(s[2 + d], q[2 + d]);
        s[5 + d] = std::max<int64_t>(s[5 + d], q[5 + d]);
      }
    }
  }
  return MVS_OK;
}

extern "C" int mvs_pc_candidate_ssim(mvs_pc_plan* p, int n_cand, const int32_t* cand_pair,
                                     const double* cand_t, const int32_t* slices,
                                     const int32_t* win, double* out_host, void* stream) {
  MVS_REQUIRE(slices && win && out_host, MVS_ERR_INVALID, "NULL pointer");
  std::vector<Cand> cands;
  int rc = fill_cands(p, n_cand, cand_pair, cand_t, cands);
  if (rc) return rc;
  const int ndim = pc_ndim(p);
  const int* sh = pc_shape(p);
  const int TZ = ndim == 3 ? kS3TZ : 1, TY = ndim == 3 ? kS3TY : kS2TY, TX = ndim == 3 ? kS3TX : kS2TX;
  long long total_tiles = 0;
  for (int i = 0; i < n_cand; ++i) {
    Cand& c = cands[i];
    c.win = win[i];
    MVS_REQUIRE(c.win >= 3 && c.win <= kMaxWin && (c.win & 1), MVS_ERR_INVALID,
                "candidate %d: SSIM window %d (odd, 3..7)", i, c.win);
    for (int d = 0; d < 3; ++d) {
      c.lo[d] = slices[6 * i + d];
      c.len[d] = slices[6 * i + 3 + d] - slices[6 * i + d];
      MVS_REQUIRE(c.lo[d] >= 0 && c.len[d] >= 1 && c.lo[d] + c.len[d] <= sh[d], MVS_ERR_INVALID,
                  "candidate %d: slice out of range on axis %d", i, d);
      MVS_REQUIRE(d < 3 - ndim || c.len[d] >= c.win, MVS_ERR_INVALID,
                  "candidate %d: window exceeds slice extent", i);
    }
    const int T[3] = {TZ, TY, TX};
    for (int d = 0; d < 3; ++d) {
      int nout = (d < 3 - ndim) ? 1 : c.len[d] - c.win + 1;
      c.tiles[d] = (nout + T[d] - 1) / T[d];
    }
    c.tile_base = total_tiles;
    total_tiles += (long long)c.tiles[0] * c.tiles[1] * c.tiles[2];
  }
  MVS_REQUIRE(total_tiles < (1LL << 31), MVS_ERR_UNSUPPORTED, "too many SSIM tiles");
  long long mat_total = 0;
  for (int i = 0; i < n_cand; ++i) {
    cands[i].mat_off = mat_total;
    mat_total += (long long)cands[i].len[0] * cands[i].len[1] * cands[i].len[2];
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cbytes = ((sizeof(Cand) * n_cand + 255) / 256) * 256;
  const size_t sbytes = ((sizeof(double) * total_tiles + 255) / 256) * 256;
  const size_t mbytes = ((sizeof(float) * total_tiles + 255) / 256) * 256;
  const size_t matbytes = sizeof(float) * (size_t)mat_total;
  void* scratch;
  if ((rc = pc_scratch(p, cbytes + sbytes + mbytes + matbytes + 256, &scratch))) return rc;
  Cand* d_c = (Cand*)scratch;
  double* d_sum = (double*)((char*)scratch + cbytes);
  float* d_max = (float*)((char*)scratch + cbytes + sbytes);
  float* d_mat = (float*)((char*)scratch + cbytes + sbytes + mbytes);
  MVS_CHECK_CUDA(cudaMemcpyAsync(d_c, cands.data(), sizeof(Cand) * n_cand,
                                 cudaMemcpyHostToDevice, st));
  {
    dim3 g(148 * 2, n_cand);
    if (ndim == 3) materialize_kernel<3><<<g, 256, 0, st>>>(d_c, sh[0], sh[1], sh[2], d_mat);
    else materialize_kernel<2><<<g, 256, 0, st>>>(d_c, sh[0], sh[1], sh[2], d_mat);
    MVS_CHECK_CUDA(cudaGetLastError());
  }
  // one launch per window size present (3 / 5 / 7): the kernels are specialised on it
  {
    int w0 = cands[0].win;
    bool same = true;
    for (int i = 1; i < n_cand; ++i) same = same && cands[i].win == w0;
    MVS_REQUIRE(same, MVS_ERR_UNSUPPORTED, "SSIM candidates of one call must share the window size");
    const unsigned g = (unsigned)total_tiles;
#define MVS_SSIM_LAUNCH(W)                                                                         \
  if (ndim == 3) ssim3d_kernel<W><<<g, kS3Threads, 0, st>>>(d_c, n_cand, sh[1], sh[2], d_mat, d_sum, d_max); \
  else ssim2d_kernel<W><<<g, kS2Threads, 0, st>>>(d_c, n_cand, sh[1], sh[2], d_mat, d_sum, d_max);
    if (w0 == 7) { MVS_SSIM_LAUNCH(7) }
    else if (w0 == 5) { MVS_SSIM_LAUNCH(5) }
    else { MVS_SSIM_LAUNCH(3) }
#undef MVS_SSIM_LAUNCH
  }
  MVS_CHECK_CUDA(cudaGetLastError());
  std::vector<double> hs(total_tiles);
  std::vector<float> hm(total_tiles);
  MVS_CHECK_CUDA(cudaMemcpyAsync(hs.data(), d_sum, sizeof(double) * total_tiles,
                                 cudaMemcpyDeviceToHost, st));
  MVS_CHECK_CUDA(cudaMemcpyAsync(hm.data(), d_max, sizeof(float) * total_tiles,
                                 cudaMemcpyDeviceToHost, st));
  MVS_CHECK_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < n_cand; ++i) {
    const Cand& c = cands[i];
    const long long nt = (long long)c.tiles[0] * c.tiles[1] * c.tiles[2];
    double s = 0.0;
    float m = -INFINITY;
    for (long long k = 0; k < nt; ++k) {
      s += hs[c.tile_base + k];
      m = std::max(m, hm[c.tile_base + k]);
    }
    double cnt = 1.0;
    for (int d = 3 - ndim; d < 3; ++d) cnt *= (double)(c.len[d] - c.win + 1);
    out_host[2 * i] = s / cnt;
    out_host[2 * i + 1] = (m == -INFINITY) ? NAN : (double)m;
  }
  return MVS_OK;
}

extern "C" int mvs_pc_spearman_batch(mvs_pc_plan* p, int n, const int32_t* pairs, const double* ts,
                                     const int64_t* n_mask, double* rho_host, void* stream) {
  MVS_REQUIRE(p && pairs && ts && n_mask && rho_host, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(n >= 1, MVS_ERR_INVALID, "n = %d", n);
  std::vector<Cand> cands;
  int rc = fill_cands(p, n, pairs, ts, cands);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int ndim = pc_ndim(p);
  const int* sh = pc_shape(p);
  const long long N = (long long)sh[0] * sh[1] * sh[2];
  // sub-batches of up to kSortSlots pairs sorted together on 32-bit keys (<= 2^26 elements)
  int B = (int)std::max<long long>(1, std::min<long long>(std::min(n, kSortSlots), (1LL << 26) / N));
  MVS_REQUIRE((long long)B * N < (1LL << 31), MVS_ERR_UNSUPPORTED, "pair volume too large for the rank sort");
  const long long E = (long long)B * N;
  size_t temp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                  (const unsigned*)nullptr, (unsigned*)nullptr, (int)E, 0, 32, st);
  {
    size_t scan_bytes = 0;
    cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                   MaxOp(), (int)E, st);
    temp_bytes = std::max(temp_bytes, scan_bytes);
  }
  auto al = [](size_t b) { return ((b + 255) / 256) * 256; };
  const size_t ub = al(sizeof(unsigned) * E), pb = al(sizeof(double) * 3 * kPearsonBlocks * n),
               cb = al(sizeof(Cand) * n), nb = al(sizeof(long long) * n);
  void* scratch;
  if ((rc = pc_scratch(p, 8 * ub + pb + cb + nb + al(temp_bytes), &scratch))) return rc;
  char* w = (char*)scratch;
  unsigned* k1 = (unsigned*)w; w += ub;     // keys in
  unsigned* k2 = (unsigned*)w; w += ub;     // keys sorted
  unsigned* v1 = (unsigned*)w; w += ub;     // payload in
  unsigned* v2 = (unsigned*)w; w += ub;     // payload sorted
  unsigned* f_in = (unsigned*)w; w += ub;   // run-head flags / scanned
  unsigned* f_out = (unsigned*)w; w += ub;
  unsigned* l_in = (unsigned*)w; w += ub;   // run-tail flags (reversed) / scanned
  unsigned* l_out = (unsigned*)w; w += ub;
  double* part = (double*)w; w += pb;       // [n][kPearsonBlocks][3]
  Cand* d_c = (Cand*)w; w += cb;
  long long* d_n = (long long*)w; w += nb;
  void* temp = w;
  const int gx = (int)std::min<long long>((N + 255) / 256, 148 * 4);
  std::vector<long long> hn(n_mask, n_mask + n);
  // everything is enqueued for all sub-batches (the scratch buffers are reused in stream order);
  // the host waits once, for the partial sums of all pairs
  MVS_CHECK_CUDA(cudaMemcpyAsync(d_c, cands.data(), sizeof(Cand) * n, cudaMemcpyHostToDevice, st));
  MVS_CHECK_CUDA(cudaMemcpyAsync(d_n, hn.data(), sizeof(long long) * n, cudaMemcpyHostToDevice, st));
  for (int b0 = 0; b0 < n; b0 += B) {
    const int nb_ = std::min(B, n - b0);
    int seg_bits = 0;
    while ((1 << seg_bits) < nb_) ++seg_bits;
    const int end_bit = kKeyBits + seg_bits;
    dim3 grid(gx, nb_);
    if (ndim == 3) spearman_keys_kernel<3><<<grid, 256, 0, st>>>(d_c + b0, sh[0], sh[1], sh[2], k1, v1);
    else spearman_keys_kernel<2><<<grid, 256, 0, st>>>(d_c + b0, sh[0], sh[1], sh[2], k1, v1);
    MVS_CHECK_CUDA(cudaGetLastError());
    const int items = (int)((long long)nb_ * N);
    // by a: (a key, b key) -> k2, v2
    MVS_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k1, k2, v1, v2, items, 0, end_bit, st));
    auto run_bounds = [&]() -> int {  // ties of k2 -> f_out (first), l_out (last, reversed order)
      run_flags_kernel<<<148 * 8, 256, 0, st>>>(k2, items, f_in, l_in);
      MVS_CHECK_CUDA(cub::DeviceScan::InclusiveScan(temp, temp_bytes, f_in, f_out, MaxOp(), items, st));
      MVS_CHECK_CUDA(cub::DeviceScan::InclusiveScan(temp, temp_bytes, l_in, l_out, MinOp(), items, st));
      return MVS_OK;
    };
    if ((rc = run_bounds())) return rc;
    rank_a_kernel<<<grid, 256, 0, st>>>(f_out, l_out, v2, N, items, d_n + b0, k1, v1);  // -> (b key, 2 rank_a)
    // by b: -> k2, v2
    MVS_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k1, k2, v1, v2, items, 0, end_bit, st));
    dim3 pg(kPearsonBlocks, nb_);
    if ((rc = run_bounds())) return rc;
    pearson_sorted_kernel<<<pg, 256, 0, st>>>(f_out, l_out, v2, N, items, d_n + b0,
                                              part + (size_t)3 * kPearsonBlocks * b0);
    MVS_CHECK_CUDA(cudaGetLastError());
  }
  std::vector<double> hp((size_t)3 * kPearsonBlocks * n);
  MVS_CHECK_CUDA(cudaMemcpyAsync(hp.data(), part, sizeof(double) * 3 * kPearsonBlocks * n, cudaMemcpyDeviceToHost, st));
  MVS_CHECK_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < n; ++i) {
    if (hn[i] < 2) { rho_host[i] = NAN; continue; }
    double sab = 0, saa = 0, sbb = 0;
    const double* q = hp.data() + (size_t)3 * kPearsonBlocks * i;
    for (int k = 0; k < kPearsonBlocks; ++k) { sab += q[3 * k]; saa += q[3 * k + 1]; sbb += q[3 * k + 2]; }
    rho_host[i] = (saa > 0 && sbb > 0) ? sab / sqrt(saa * sbb) : NAN;
  }
  return MVS_OK;
}
