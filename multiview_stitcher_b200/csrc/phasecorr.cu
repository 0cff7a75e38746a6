// Batched 2-D/3-D phase correlation (sm_100a): the FFT half of
// registration.phase_correlation_registration (registration.py:353-443), i.e.
// rescale_intensity -> two skimage.registration.phase_cross_correlation calls
// (normalization "phase" and None) including the upsampled-DFT refinement.
//
// Data flow per pair (all pairs of a plan share one shape, batched per launch):
//   r0, r1   = rescale_intensity(fixed / moving) to [0,1], NaN kept    (:382-389)
//   Z        = FFT_nd(nan_to_num(r0) + i * nan_to_num(r1))             packed:
//              F = (Z_k + conj Z_-k)/2,  M = (Z_k - conj Z_-k)/(2i)
//   P        = F * conj(M);  Pn = P / max(|P|, 100 eps)                 cross power
//   Q        = s P + i * Pn;  cc = IFFT_nd(Q):  Re = cc(None), Im = cc("phase")
//              (P, Pn are Hermitian, so both correlation surfaces are real and
//               ONE complex inverse transform yields both; s = 1/N^2 keeps
//               |s P| <= 1 next to the unit-modulus Pn; only the integer peaks
//               are read off this packed transform)
//   peaks    = first argmax of |Re cc|, |Im cc|
//   C        = conj(upsampled_dft(conj(P or Pn))) on ceil(1.5 u)^ndim samples
//              around each peak (float64 accumulation) from the plain P
// Passes over HBM (fft_pass.cuh), N complex voxels of 8 bytes per pair:
//   forward x   reads r0, r1 (8N)  writes Z (8N)
//   forward y   [3-D] in place (16N)
//   forward y/z + cross power: reads Z (8N) writes Q and P (16N) -- Z's last axis
//               and the cross-power spectrum in one kernel (mirror lines paired per CTA)
//   inverse     first axis in place (16N), [3-D] middle axis in place (16N)
//   inverse x + argmax: reads Q (8N), stores nothing
//   upsampled DFT: reads P (8N)
// i.e. (32 ndim + 8) N bytes, the "pass model" of SURVEY.md 8d that the roofline
// figure is quoted on.

#include <climits>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "fft_pass.cuh"

namespace mvs {

// ---------------------------------------------------------------------------
// per-length tables
// ---------------------------------------------------------------------------

static std::mutex g_fft_mutex;
static std::map<int, AxisFft*> g_fft_cache;

static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

static void host_fft_pow2(std::vector<double>& re, std::vector<double>& im) {
  // plain iterative radix-2 in float64 (table construction only)
  const int m = (int)re.size();
  for (int i = 1, j = 0; i < m; ++i) {
    int bit = m >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  for (int len = 2; len <= m; len <<= 1) {
    for (int i = 0; i < m; i += len)
      for (int k = 0; k < len / 2; ++k) {
        double ang = -2.0 * M_PI * k / len;
        double wr = cos(ang), wi = sin(ang);
        int a = i + k, b = i + k + len / 2;
        double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] += xr; im[a] += xi;
      }
  }
}

// float64 O(m^2) DFT for the one non-power-of-two kernel length (table construction only)
static void host_dft(std::vector<double>& re, std::vector<double>& im) {
  const int m = (int)re.size();
  std::vector<double> yr(m), yi(m);
  for (int k = 0; k < m; ++k) {
    double sr = 0.0, si = 0.0;
    for (int j = 0; j < m; ++j) {
      const double ang = -2.0 * M_PI * (double)(((long long)j * k) % m) / (double)m;
      const double c = cos(ang), sn = sin(ang);
      sr += re[j] * c - im[j] * sn;
      si += re[j] * sn + im[j] * c;
    }
    yr[k] = sr; yi[k] = si;
  }
  re.swap(yr); im.swap(yi);
}

// Kernel length of the chirp-z transform of n points: the next power of two >= 2n-1, or
// (MVS_BLUESTEIN_SMOOTH=1 in the environment, read once) the 640-point kernel where it
// fits -- 513 <= 2n-1 <= 640, i.e. 257 <= n <= 320 -- at 0.58x the flops of 1024 points.
static int bluestein_length(int n) {
  static const bool smooth = [] {
    const char* e = getenv("MVS_BLUESTEIN_SMOOTH");
    return e && e[0] == '1';
  }();
  int m = 1;
  while (m < 2 * n - 1) m <<= 1;
  if (smooth && m == 1024 && 2 * n - 1 <= 640) m = 640;
  return m;
}

const AxisFft* get_axis_fft(int n) {
  std::lock_guard<std::mutex> lock(g_fft_mutex);
  auto it = g_fft_cache.find(n);
  if (it != g_fft_cache.end()) return it->second;
  if (n < 1) { set_error("FFT length %d", n); return nullptr; }
  AxisFft ax{};
  ax.n = n;
  if (is_pow2(n)) { ax.m = n; ax.bluestein = 0; }
  else { ax.m = bluestein_length(n); ax.bluestein = 1; }
  if (ax.m > 8192) {
    set_error("axis length %d needs a %d-point FFT (max 8192)", n, ax.m);
    return nullptr;
  }
  const int m = ax.m;
  std::vector<float2> tw(m);
  for (int k = 0; k < m; ++k) {
    double ang = -2.0 * M_PI * (double)k / (double)m;
    tw[k] = make_float2((float)cos(ang), (float)sin(ang));
  }
  float2 *d_tw = nullptr, *d_chirp = nullptr, *d_bhat = nullptr;
  auto up = [&](float2** d, const std::vector<float2>& h) {
    if (cudaMalloc(d, sizeof(float2) * h.size()) != cudaSuccess) return false;
    return cudaMemcpy(*d, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice) ==
           cudaSuccess;
  };
  bool ok = up(&d_tw, tw);
  if (ok && ax.bluestein) {
    std::vector<float2> chirp(n), bhat(m);
    std::vector<double> cr(n), ci(n), br(m, 0.0), bi(m, 0.0);
    for (int k = 0; k < n; ++k) {
      long long k2 = ((long long)k * k) % (2LL * n);  // exact phase reduction
      double ang = -M_PI * (double)k2 / (double)n;
      cr[k] = cos(ang); ci[k] = sin(ang);
      chirp[k] = make_float2((float)cr[k], (float)ci[k]);
    }
    for (int k = 0; k < n; ++k) {
      br[k] = cr[k]; bi[k] = -ci[k];
      if (k) { br[m - k] = cr[k]; bi[m - k] = -ci[k]; }
    }
    if (is_pow2(m)) host_fft_pow2(br, bi); else host_dft(br, bi);
    for (int k = 0; k < m; ++k)
      bhat[k] = make_float2((float)(br[k] / m), (float)(bi[k] / m));
    ok = up(&d_chirp, chirp) && up(&d_bhat, bhat);
  }
  if (!ok) {
    set_error("FFT table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_tw); cudaFree(d_chirp); cudaFree(d_bhat);
    return nullptr;
  }
  ax.tw = d_tw; ax.chirp = d_chirp; ax.bhat = d_bhat;
  AxisFft* p = new AxisFft(ax);
  g_fft_cache[n] = p;
  return p;
}

// lines per CTA / shared-memory geometry of one pass (fft_pass.cuh)
struct PassGeom { int L, line_stride, threads; size_t smem; };

static PassGeom pass_geom(int m, bool contiguous, bool paired, long long nlines) {
  const int E = fft_values_per_thread(m), T = m / E;
  int L;
  if (contiguous) {
    L = T >= 128 ? 1 : 128 / T;  // ~128 threads per CTA: several CTAs per SM in different phases
  } else {
    // runs of >= 32 bytes along the fastest axis
    L = T <= 32 ? 256 / T : (T <= 64 ? 8 : (T <= 128 ? 4 : (T <= 256 ? 2 : 1)));
  }
  if (L > 32) L = 32;
  if (m == 640 && L > 8) L = 8;  // 256-thread CTAs (the kernel's launch bound)
  while (L > 1 && (long long)(L / 2) >= nlines) L >>= 1;  // few lines: smaller CTAs
  if (paired && L < 2) L = 2;
  while (L * T < 32) L <<= 1;  // whole warps
  PassGeom g;
  g.L = L;
  const int padm = m + (m >> 4);
  // skew successive line buffers so that the "lines fastest" thread mapping spreads
  // a half-warp over all banks
  const int skew = contiguous ? (T < 16 ? T : 0) : (L <= 16 ? 16 / L : 1);
  g.line_stride = padm + skew;
  g.threads = L * T;
  g.smem = sizeof(float2) * (size_t)L * g.line_stride;
  return g;
}

// ---------------------------------------------------------------------------
// element-wise kernels
// ---------------------------------------------------------------------------

constexpr int kRedBlocks = 64;  // partial-reduction blocks per image

// per-image partial min / max / NaN count / bbox of non-NaN voxels / sum of the values
__global__ void __launch_bounds__(256)
stats_kernel(const float* const* __restrict__ imgs, long long N, int n1, int n2,
             double* __restrict__ partial /* [img][block][10] */) {
  const float* im = imgs[blockIdx.y];
  float mn = INFINITY, mx = -INFINITY;
  long long nan = 0;
  double vsum = 0.0;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {-1, -1, -1};
  // one warp per 256-voxel row segment: (z, y) are known per segment, so no per-voxel divisions
  // (three 64-bit divisions per voxel cost several times the reduction itself)
  const unsigned rows = (unsigned)(N / n2), xt = ((unsigned)n2 + 255) / 256;
  const unsigned nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (unsigned u = blockIdx.x * nwarps + (threadIdx.x >> 5); u < rows * xt; u += gridDim.x * nwarps) {
    const unsigned row = u / xt, x0 = (u - row * xt) * 256;
    const int z = (int)(row / (unsigned)n1), y = (int)(row - (unsigned)z * (unsigned)n1);
    const float* rp = im + (long long)row * n2;
    const int x1 = min(n2, (int)x0 + 256);
    bool any = false;
    for (int x = (int)(x0 + lane); x < x1; x += 32) {
      const float v = __ldg(rp + x);
      if (v != v) { ++nan; continue; }
      mn = fminf(mn, v); mx = fmaxf(mx, v);
      vsum += (double)v;
      lo[2] = min(lo[2], x); hi[2] = max(hi[2], x);
      any = true;
    }
    if (any) {
      lo[0] = min(lo[0], z); lo[1] = min(lo[1], y);
      hi[0] = max(hi[0], z); hi[1] = max(hi[1], y);
    }
  }
  __shared__ float s_mn[256], s_mx[256];
  __shared__ long long s_nan[256];
  __shared__ double s_sum[256];
  __shared__ int s_lo[3][256], s_hi[3][256];
  const int t = threadIdx.x;
  s_mn[t] = mn; s_mx[t] = mx; s_nan[t] = nan; s_sum[t] = vsum;
  for (int d = 0; d < 3; ++d) { s_lo[d][t] = lo[d]; s_hi[d][t] = hi[d]; }
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (t < s) {
      s_mn[t] = fminf(s_mn[t], s_mn[t + s]); s_mx[t] = fmaxf(s_mx[t], s_mx[t + s]);
      s_nan[t] += s_nan[t + s];
      s_sum[t] += s_sum[t + s];
      for (int d = 0; d < 3; ++d) {
        s_lo[d][t] = min(s_lo[d][t], s_lo[d][t + s]);
        s_hi[d][t] = max(s_hi[d][t], s_hi[d][t + s]);
      }
    }
    __syncthreads();
  }
  if (t == 0) {
    double* p = partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 10;
    p[0] = s_mn[0]; p[1] = s_mx[0]; p[2] = (double)s_nan[0];
    for (int d = 0; d < 3; ++d) { p[3 + d] = s_lo[d][0]; p[6 + d] = s_hi[d][0]; }
    p[9] = s_sum[0];
  }
}

// skimage.exposure.rescale_intensity(im, in_range=(nanmin, nanmax), out_range=(0,1))
// in float32: (im - imin) / float32(imax - imin); NaN stays NaN.
__global__ void __launch_bounds__(256)
rescale_kernel(const float* const* __restrict__ imgs, const float* __restrict__ mn,
               const float* __restrict__ scale, float* __restrict__ out0,
               float* __restrict__ out1, long long N) {
  const int img = blockIdx.y;  // 2*pair + which
  const float* im = imgs[img];
  float* out = ((img & 1) ? out1 : out0) + (long long)(img >> 1) * N;
  const float lo = mn[img], sc = scale[img];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    float v = __ldg(im + i);
    float r;
    if (sc > 0.f) r = __fdiv_rn(__fsub_rn(v, lo), sc);
    else r = fminf(fmaxf(v, 0.f), 1.f);  // constant image: clip to out_range
    out[i] = (v != v) ? v : r;
  }
}

// ---------------------------------------------------------------------------
// upsampled DFT around the integer peaks
// ---------------------------------------------------------------------------
// cc'[a..] = sum_k P_k prod_d exp(+2 pi i (a_d - off_d) ks_d(k_d) / (n_d u)),
// off_d = fix(R/2) - shift_d u, ks = numpy.fft.fftfreq ordering, a_d < R.
// Along x the kernel factor splits into a peak-dependent ramp H[k] (folded into
// the spectrum) and the peak-independent geometric sequence g[k]^a, so the x
// contraction needs no table lookups: acc[a] += w; w *= g.

// Decodes the peaks, applies skimage's wrap (shift > fix(n/2) -> shift - n) and
// fills, per (pair, norm): H_x[n2], E_y[R][n1], E_z[R][n0]; block 0 also fills
// g[n2] = exp(+2 pi i ks / (n2 u)).
__global__ void __launch_bounds__(256)
updft_setup_kernel(const unsigned long long* __restrict__ keys, int n0, int n1, int n2,
                   int ndim, int R, int u, int* __restrict__ peaks /* [pair][2][3] */,
                   float2* __restrict__ E, long long e_stride /* per (pair,norm) */,
                   int e_off1, int e_off0, float2* __restrict__ G) {
  // grid (pair*2 + norm, slices): the table entries of one (pair, norm) are spread
  // over gridDim.y blocks
  const int pn = blockIdx.x;
  const int tid = blockIdx.y * blockDim.x + threadIdx.x, nth = gridDim.y * blockDim.x;
  const unsigned long long key = keys[pn];
  const long long idx = (long long)(0xffffffffull - (key & 0xffffffffull));
  int p[3];
  p[2] = (int)(idx % n2); p[1] = (int)((idx / n2) % n1); p[0] = (int)(idx / ((long long)n1 * n2));
  const int nn[3] = {n0, n1, n2};
  int sh[3];
  for (int d = 0; d < 3; ++d) sh[d] = p[d] > nn[d] / 2 ? p[d] - nn[d] : p[d];
  if (blockIdx.y == 0 && threadIdx.x < 3) peaks[pn * 3 + threadIdx.x] = sh[threadIdx.x];
  float2* e = E + (long long)pn * e_stride;
  {
    // off - R/2: the x contraction runs over centred powers g^(a - R/2) (updft_x_kernel)
    const double off = (double)(R / 2) - (double)sh[2] * u - (double)(R / 2);
    for (int k = tid; k < n2; k += nth) {
      const int ks = (k <= (n2 - 1) / 2) ? k : k - n2;
      double s, c;
      sincospi(-2.0 * off * (double)ks / ((double)n2 * (double)u), &s, &c);
      e[k] = make_float2((float)c, (float)s);
      if (pn == 0) {
        sincospi(2.0 * (double)ks / ((double)n2 * (double)u), &s, &c);
        G[k] = make_float2((float)c, (float)s);
      }
    }
  }
  const int eoff[2] = {e_off0, e_off1};
  for (int d = 3 - ndim; d < 2; ++d) {
    const int n = nn[d];
    const double off = (double)(R / 2) - (double)sh[d] * u;
    float2* ed = e + eoff[d];
    for (int i = tid; i < R * n; i += nth) {
      const int a = i / n, k = i - a * n;
      const int ks = (k <= (n - 1) / 2) ? k : k - n;
      double s, c;
      sincospi(2.0 * ((double)a - off) * (double)ks / ((double)n * (double)u), &s, &c);
      ed[i] = make_float2((float)c, (float)s);
    }
  }
}

// Contract the x axis for both normalisations: T[pn][line][a] = sum_x P_pn[line, x] H_pn[x] g[x]^(a - R/2)
// (the setup kernel folds g^(R/2) into H, so the powers are centred).  |g| = 1, hence
// g^-j = conj(g^j): with w = P H and g^j = c + i s the four real sums
//   A = sum wr c, B = sum wi s, C = sum wr s, D = sum wi c
// give BOTH outputs of a +-j pair -- (A - B) + i (C + D) and (A + B) + i (D - C) -- for four
// FMAs per element and norm, against the twelve instructions per output of a plain
// accumulate-and-rotate loop.  One warp per line; float partials per lane, combined in float64
// in lane order.
constexpr int kUpWarps = 4;

template <int R>
__global__ void __launch_bounds__(kUpWarps * 32)
updft_x_kernel(const float2* __restrict__ Pg, int n0, int n1, int n2, long long N,
               const float2* __restrict__ E, long long e_stride, const float2* __restrict__ G,
               double2* __restrict__ T) {
  constexpr int C0 = R / 2, JP = R - 1 - C0, JM = C0 > JP ? C0 : JP;  // negative side reaches C0
  __shared__ float2 s_part[kUpWarps][2 * R][33];
  const int pair = blockIdx.y;
  const long long nlines = (long long)n0 * n1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long line = (long long)blockIdx.x * kUpWarps + warp;
  if (line < nlines) {
    const float2* row = Pg + (long long)pair * N + line * n2;
    const float2* h0 = E + (long long)(2 * pair) * e_stride;
    const float2* h1 = h0 + e_stride;
    const float tiny = 100.0f * 1.1920929e-07f;  // 100 * eps(float32)
    float2 cen0 = make_float2(0.f, 0.f), cen1 = make_float2(0.f, 0.f);
    float A0[JM > 0 ? JM : 1], B0[JM > 0 ? JM : 1], C0s[JM > 0 ? JM : 1], D0[JM > 0 ? JM : 1];
    float A1[JM > 0 ? JM : 1], B1[JM > 0 ? JM : 1], C1s[JM > 0 ? JM : 1], D1[JM > 0 ? JM : 1];
#pragma unroll
    for (int j = 0; j < JM; ++j) { A0[j] = B0[j] = C0s[j] = D0[j] = 0.f; A1[j] = B1[j] = C1s[j] = D1[j] = 0.f; }
    // two x per iteration: the power chains of the two elements are independent
    for (int x = lane; x < n2; x += 64) {
      const int x2 = x + 32;
      const bool has2 = x2 < n2;
      const int xb = has2 ? x2 : x;
      const float2 PsA = row[x];
      const float2 PsB = has2 ? row[x2] : make_float2(0.f, 0.f);
      const float magA = fmaxf(hypotf(PsA.x, PsA.y), tiny), magB = fmaxf(hypotf(PsB.x, PsB.y), tiny);
      const float2 PnA = make_float2(__fdiv_rn(PsA.x, magA), __fdiv_rn(PsA.y, magA));
      const float2 PnB = make_float2(__fdiv_rn(PsB.x, magB), __fdiv_rn(PsB.y, magB));
      const float2 gA = __ldg(G + x), gB = __ldg(G + xb);
      const float2 w0A = cmul(PsA, __ldg(h0 + x)), w1A = cmul(PnA, __ldg(h1 + x));
      const float2 w0B = cmul(PsB, __ldg(h0 + xb)), w1B = cmul(PnB, __ldg(h1 + xb));
      cen0 = cadd(cen0, cadd(w0A, w0B));
      cen1 = cadd(cen1, cadd(w1A, w1B));
      float2 qA = gA, qB = gB;
#pragma unroll
      for (int j = 0; j < JM; ++j) {
        A0[j] = fmaf(w0A.x, qA.x, A0[j]); A0[j] = fmaf(w0B.x, qB.x, A0[j]);
        B0[j] = fmaf(w0A.y, qA.y, B0[j]); B0[j] = fmaf(w0B.y, qB.y, B0[j]);
        C0s[j] = fmaf(w0A.x, qA.y, C0s[j]); C0s[j] = fmaf(w0B.x, qB.y, C0s[j]);
        D0[j] = fmaf(w0A.y, qA.x, D0[j]); D0[j] = fmaf(w0B.y, qB.x, D0[j]);
        A1[j] = fmaf(w1A.x, qA.x, A1[j]); A1[j] = fmaf(w1B.x, qB.x, A1[j]);
        B1[j] = fmaf(w1A.y, qA.y, B1[j]); B1[j] = fmaf(w1B.y, qB.y, B1[j]);
        C1s[j] = fmaf(w1A.x, qA.y, C1s[j]); C1s[j] = fmaf(w1B.x, qB.y, C1s[j]);
        D1[j] = fmaf(w1A.y, qA.x, D1[j]); D1[j] = fmaf(w1B.y, qB.x, D1[j]);
        if (j + 1 < JM) { qA = cmul(qA, gA); qB = cmul(qB, gB); }
      }
    }
    s_part[warp][C0][lane] = cen0;
    s_part[warp][R + C0][lane] = cen1;
#pragma unroll
    for (int j = 0; j < JM; ++j) {
      if (j < JP) {
        s_part[warp][C0 + 1 + j][lane] = make_float2(A0[j] - B0[j], C0s[j] + D0[j]);
        s_part[warp][R + C0 + 1 + j][lane] = make_float2(A1[j] - B1[j], C1s[j] + D1[j]);
      }
      if (j < C0) {
        s_part[warp][C0 - 1 - j][lane] = make_float2(A0[j] + B0[j], D0[j] - C0s[j]);
        s_part[warp][R + C0 - 1 - j][lane] = make_float2(A1[j] + B1[j], D1[j] - C1s[j]);
      }
    }
  }
  __syncwarp();
  if (line < nlines) {
    for (int o = lane; o < 2 * R; o += 32) {
      double re = 0.0, im = 0.0;
      for (int i = 0; i < 32; ++i) { const float2 v = s_part[warp][o][i]; re += v.x; im += v.y; }
      const int norm = o / R, b = o - norm * R;
      T[((long long)(2 * pair + norm) * nlines + line) * R + b] = make_double2(re, im);
    }
  }
}

// Contract one more axis: out[pn][o][a][r] = sum_k E[a][k] * in[pn][o][k][r].
// One CTA per (o, a, pn): threads = (k-slice, r), float64 partials combined in
// slice order (deterministic).
__global__ void __launch_bounds__(256)
updft_axis_kernel(const double2* __restrict__ in, double2* __restrict__ out, int outer, int n,
                  int inner, int R, const float2* __restrict__ E, long long e_stride, int e_off) {
  __shared__ double2 s_part[256];
  const int pn = blockIdx.y;
  const int o = blockIdx.x / R, a = blockIdx.x - o * R;
  const int slices = 256 / inner;  // inner = R^(axes done) <= 225
  const int sl = threadIdx.x / inner, r = threadIdx.x - sl * inner;
  double re = 0.0, im = 0.0;
  if (sl < slices) {
    const float2* e = E + (long long)pn * e_stride + e_off + (long long)a * n;
    const double2* src = in + ((long long)pn * outer + o) * n * inner + r;
    // eight independent loads in flight per thread (the loop is latency-bound otherwise)
    for (int k0 = sl; k0 < n; k0 += 8 * slices) {
      double2 v[8];
      float2 w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + j * slices;
        const int kc = k < n ? k : sl;
        v[j] = src[(long long)kc * inner];
        w[j] = __ldg(e + kc);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (k0 + j * slices < n) {
          re += v[j].x * w[j].x - v[j].y * w[j].y;
          im += v[j].x * w[j].y + v[j].y * w[j].x;
        }
      }
    }
  }
  s_part[threadIdx.x] = make_double2(re, im);
  __syncthreads();
  if (threadIdx.x < inner) {
    double sr = 0.0, si = 0.0;
    for (int i = 0; i < slices; ++i) { const double2 v = s_part[i * inner + threadIdx.x]; sr += v.x; si += v.y; }
    out[((long long)pn * outer + o) * R * inner + (long long)a * inner + threadIdx.x] = make_double2(sr, si);
  }
}

}  // namespace mvs

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------

using namespace mvs;

struct mvs_pc_plan {
  int ndim = 2;
  int shape[3] = {1, 1, 1};
  long long N = 0;
  int max_pairs = 0;
  int upsample = 10;
  int R = 15;
  AxisFft ax[3];
  float *r0 = nullptr, *r1 = nullptr;
  float2 *Z = nullptr, *Q = nullptr;
  const float** d_imgs = nullptr;  // [2*max_pairs]
  double* d_partial = nullptr;     // stats partials
  float *d_mn = nullptr, *d_scale = nullptr;
  float* d_cp = nullptr;  // per pair: scale of P in the packed spectrum
  unsigned long long* d_keys = nullptr;
  int* d_peaks = nullptr;
  float2* d_E = nullptr;
  float2* d_G = nullptr;  // [n2] geometric-sequence base of the x contraction
  long long e_stride = 0;
  int e_off[3] = {0, 0, 0};
  double2 *d_T0 = nullptr, *d_T1 = nullptr;
  long long t_stride = 0;
  int loaded = 0;
  // scratch for the disambiguation stages (disambig.cu)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
};

namespace mvs {
// accessors used by disambig.cu
const float* pc_r0(const mvs_pc_plan* p, int pair) { return p->r0 + (long long)pair * p->N; }
const float* pc_r1(const mvs_pc_plan* p, int pair) { return p->r1 + (long long)pair * p->N; }
int pc_ndim(const mvs_pc_plan* p) { return p->ndim; }
const int* pc_shape(const mvs_pc_plan* p) { return p->shape; }
int pc_loaded(const mvs_pc_plan* p) { return p->loaded; }
int pc_scratch(mvs_pc_plan* p, size_t bytes, void** out) {
  if (p->scratch_bytes < bytes) {
    cudaFree(p->scratch);
    p->scratch = nullptr; p->scratch_bytes = 0;
    MVS_CHECK_CUDA(cudaMalloc(&p->scratch, bytes));
    p->scratch_bytes = bytes;
  }
  *out = p->scratch;
  return MVS_OK;
}
}  // namespace mvs

extern "C" int mvs_pc_plan_destroy(mvs_pc_plan* p) {
  if (!p) return MVS_OK;
  cudaFree(p->r0); cudaFree(p->r1); cudaFree(p->Z); cudaFree(p->Q);
  cudaFree((void*)p->d_imgs); cudaFree(p->d_partial); cudaFree(p->d_mn); cudaFree(p->d_scale); cudaFree(p->d_cp);
  cudaFree(p->d_keys); cudaFree(p->d_peaks); cudaFree(p->d_E); cudaFree(p->d_G); cudaFree(p->d_T0);
  cudaFree(p->d_T1); cudaFree(p->scratch);
  delete p;
  return MVS_OK;
}

extern "C" int mvs_pc_plan_create(mvs_pc_plan** plan, int ndim, const int32_t shape[3],
                                  int max_pairs, int upsample_factor) {
  MVS_REQUIRE(plan && shape, MVS_ERR_INVALID, "NULL pointer");
  *plan = nullptr;
  MVS_REQUIRE(ndim == 2 || ndim == 3, MVS_ERR_INVALID, "ndim must be 2 or 3");
  MVS_REQUIRE(max_pairs >= 1, MVS_ERR_INVALID, "max_pairs must be >= 1");
  MVS_REQUIRE(upsample_factor >= 1 && upsample_factor <= 10, MVS_ERR_UNSUPPORTED,
              "upsample_factor %d not supported (1..10)", upsample_factor);
  MVS_REQUIRE(ndim == 3 || shape[0] == 1, MVS_ERR_INVALID, "2-D plan needs shape[0] == 1");
  for (int d = 0; d < 3; ++d)
    MVS_REQUIRE(shape[d] >= 1, MVS_ERR_INVALID, "shape[%d] = %d", d, shape[d]);
  mvs_pc_plan* p = new mvs_pc_plan();
  p->ndim = ndim;
  p->N = 1;
  for (int d = 0; d < 3; ++d) { p->shape[d] = shape[d]; p->N *= shape[d]; }
  if (p->N >= (1LL << 31)) {
    set_error("pair volume of %lld voxels exceeds 2^31", p->N);
    delete p;
    return MVS_ERR_UNSUPPORTED;
  }
  p->max_pairs = max_pairs;
  p->upsample = upsample_factor;
  p->R = (int)ceil(upsample_factor * 1.5);
  for (int d = 3 - ndim; d < 3; ++d) {
    const AxisFft* ax = get_axis_fft(shape[d]);
    if (!ax) { delete p; return MVS_ERR_UNSUPPORTED; }
    p->ax[d] = *ax;
  }
  const long long NP = p->N * max_pairs;
  // per (pair, norm): H_x[n2], E_y[R][n1], E_z[R][n0] (updft_setup_kernel)
  int off = shape[2];
  p->e_off[2] = 0;
  p->e_off[1] = off; off += p->R * shape[1];
  p->e_off[0] = off; if (ndim == 3) off += p->R * shape[0];
  p->e_stride = off;
  const long long lines = p->N / shape[2];
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** ptr, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(ptr, bytes ? bytes : 16);
  };
  alloc((void**)&p->r0, sizeof(float) * NP);
  alloc((void**)&p->r1, sizeof(float) * NP);
  alloc((void**)&p->Z, sizeof(float2) * NP);
  alloc((void**)&p->Q, sizeof(float2) * NP);
  alloc((void**)&p->d_imgs, sizeof(float*) * 2 * max_pairs);
  alloc((void**)&p->d_partial, sizeof(double) * 10 * kRedBlocks * 2 * max_pairs);
  alloc((void**)&p->d_cp, sizeof(float) * max_pairs);
  alloc((void**)&p->d_mn, sizeof(float) * 2 * max_pairs);
  alloc((void**)&p->d_scale, sizeof(float) * 2 * max_pairs);
  alloc((void**)&p->d_keys, sizeof(unsigned long long) * 2 * max_pairs);
  alloc((void**)&p->d_peaks, sizeof(int) * 6 * max_pairs);
  alloc((void**)&p->d_E, sizeof(float2) * p->e_stride * 2 * max_pairs);
  alloc((void**)&p->d_G, sizeof(float2) * shape[2]);
  {
    const long long R = p->R;
    long long t = lines * R;
    t = std::max(t, (long long)shape[0] * R * R);
    t = std::max(t, R * R * R);
    p->t_stride = t;
  }
  alloc((void**)&p->d_T0, sizeof(double2) * p->t_stride * 2 * max_pairs);
  alloc((void**)&p->d_T1, sizeof(double2) * p->t_stride * 2 * max_pairs);
  if (e != cudaSuccess) {
    set_error("phase-correlation plan allocation failed: %s", cudaGetErrorString(e));
    mvs_pc_plan_destroy(p);
    return MVS_ERR_CUDA;
  }
  *plan = p;
  return MVS_OK;
}

static int grid_for(long long N) {
  long long b = (N + 255) / 256;
  return (int)(b < 148 * 8 ? b : 148 * 8);
}

extern "C" int mvs_pc_load_pairs(mvs_pc_plan* p, int n, const float* const* fixed,
                                 const float* const* moving, double* stats_host, void* stream) {
  MVS_REQUIRE(p && fixed && moving && stats_host, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(n >= 1 && n <= p->max_pairs, MVS_ERR_INVALID, "n = %d outside 1..%d", n,
              p->max_pairs);
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<const float*> imgs(2 * n);
  for (int i = 0; i < n; ++i) {
    MVS_REQUIRE(fixed[i] && moving[i], MVS_ERR_INVALID, "pair %d: NULL image", i);
    imgs[2 * i] = fixed[i];
    imgs[2 * i + 1] = moving[i];
  }
  MVS_CHECK_CUDA(cudaMemcpyAsync((void*)p->d_imgs, imgs.data(), sizeof(float*) * 2 * n,
                                 cudaMemcpyHostToDevice, st));
  dim3 g(kRedBlocks, 2 * n);
  stats_kernel<<<g, 256, 0, st>>>(p->d_imgs, p->N, p->shape[1], p->shape[2], p->d_partial);
  MVS_CHECK_CUDA(cudaGetLastError());
  std::vector<double> part((size_t)10 * kRedBlocks * 2 * n);
  MVS_CHECK_CUDA(cudaMemcpyAsync(part.data(), p->d_partial, sizeof(double) * part.size(),
                                 cudaMemcpyDeviceToHost, st));
  MVS_CHECK_CUDA(cudaStreamSynchronize(st));
  std::vector<float> mn(2 * n), sc(2 * n), cp(n);
  std::vector<double> rsum(2 * n);
  for (int img = 0; img < 2 * n; ++img) {
    double* s = stats_host + (size_t)img * 9;
    double vsum = 0.0;
    s[0] = INFINITY; s[1] = -INFINITY; s[2] = 0;
    for (int d = 0; d < 3; ++d) { s[3 + d] = 2147483647.0; s[6 + d] = -1; }
    for (int b = 0; b < kRedBlocks; ++b) {
      const double* q = part.data() + ((size_t)img * kRedBlocks + b) * 10;
      s[0] = std::min(s[0], q[0]); s[1] = std::max(s[1], q[1]); s[2] += q[2];
      vsum += q[9];
      for (int d = 0; d < 3; ++d) {
        s[3 + d] = std::min(s[3 + d], q[3 + d]);
        s[6 + d] = std::max(s[6 + d], q[6 + d]);
      }
    }
    mn[img] = (float)s[0];
    // (imax - imin) evaluated in float64, used as a float32 divisor (skimage)
    sc[img] = (s[1] > s[0]) ? (float)(s[1] - s[0]) : 0.0f;
    // sum of the rescaled image = DC term of its spectrum
    const double nvalid = (double)p->N - s[2];
    rsum[img] = (s[1] > s[0]) ? (vsum - nvalid * s[0]) / (s[1] - s[0]) : nvalid;
  }
  // Packed spectrum Q = s P + i Pn: s puts the largest |P| (the DC product) at 2^12,
  // so the unit-modulus Pn keeps >= 12 bits next to it everywhere while P keeps a
  // 2^36 dynamic range above the float32 rounding of Pn -- both correlation
  // surfaces come out of ONE inverse transform at full working precision.
  for (int i = 0; i < n; ++i) {
    const double dc = rsum[2 * i] * rsum[2 * i + 1];
    cp[i] = (float)(dc > 1e-30 ? 4096.0 / dc : 1.0);
  }
  MVS_CHECK_CUDA(cudaMemcpyAsync(p->d_cp, cp.data(), sizeof(float) * n, cudaMemcpyHostToDevice, st));
  MVS_CHECK_CUDA(cudaMemcpyAsync(p->d_mn, mn.data(), sizeof(float) * 2 * n,
                                 cudaMemcpyHostToDevice, st));
  MVS_CHECK_CUDA(cudaMemcpyAsync(p->d_scale, sc.data(), sizeof(float) * 2 * n,
                                 cudaMemcpyHostToDevice, st));
  dim3 g2(grid_for(p->N), 2 * n);
  rescale_kernel<<<g2, 256, 0, st>>>(p->d_imgs, p->d_mn, p->d_scale, p->r0, p->r1, p->N);
  MVS_CHECK_CUDA(cudaGetLastError());
  MVS_CHECK_CUDA(cudaStreamSynchronize(st));  // mn/sc staging vectors die here
  p->loaded = n;
  return MVS_OK;
}

enum PassKind { PASS_LOAD_REAL, PASS_PLAIN, PASS_PAIRED, PASS_ARGMAX };

// The dynamic shared-memory cap of a kernel is per-function state shared by every
// thread of the process (crop-shape groups are registered from concurrent threads
// with different line counts): raise it once to the largest size any launch uses.
constexpr int kFftSmemCap = 200 * 1024;

template <int M, bool BLUE>
static cudaError_t fft_kernel_ready() {
  static std::once_flag once;
  static cudaError_t status = cudaSuccess;
  std::call_once(once, [] {
    status = cudaFuncSetAttribute(fft_reg_pass_kernel<M, BLUE>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kFftSmemCap);
  });
  return status;
}

template <int M>
static int launch_pass_m(const FftPassArgs& a, bool blue, dim3 grid, int threads, size_t smem,
                         cudaStream_t st) {
  MVS_REQUIRE(smem <= (size_t)kFftSmemCap, MVS_ERR_UNSUPPORTED, "FFT pass needs %zu bytes of shared memory", smem);
  if (blue) {
    MVS_CHECK_CUDA((fft_kernel_ready<M, true>()));
    fft_reg_pass_kernel<M, true><<<grid, threads, smem, st>>>(a);
  } else {
    MVS_CHECK_CUDA((fft_kernel_ready<M, false>()));
    fft_reg_pass_kernel<M, false><<<grid, threads, smem, st>>>(a);
  }
  MVS_CHECK_CUDA(cudaGetLastError());
  return MVS_OK;
}

// One pass along `axis` over the n loaded pairs.  sign -1 forward / +1 inverse.
static int launch_pass(mvs_pc_plan* p, int n, int axis, int sign, PassKind kind, const float2* src,
                       float2* dst, cudaStream_t st, float2* dst2 = nullptr) {
  FftPassArgs a{};
  const AxisFft& ax = p->ax[axis];
  a.src = src; a.dst = dst; a.dst2 = dst2;
  a.re = p->r0; a.im = p->r1;
  a.n = p->shape[axis];
  long long inner = 1, outer = 1;
  for (int d = axis + 1; d < 3; ++d) inner *= p->shape[d];
  for (int d = 0; d < axis; ++d) outer *= p->shape[d];
  a.inner = inner; a.outer = outer;
  a.batch_stride = p->N;
  a.sign = sign;
  a.load_real = kind == PASS_LOAD_REAL;
  a.paired = kind == PASS_PAIRED;
  a.argmax = kind == PASS_ARGMAX;
  a.contig = (axis == 2);
  a.keys = p->d_keys;
  a.tw = ax.tw; a.chirp = ax.chirp; a.bhat = ax.bhat;
  const long long nlines = outer * inner;
  const PassGeom g = pass_geom(ax.m, a.contig != 0, a.paired != 0, nlines);
  a.L = g.L; a.line_stride = g.line_stride;
  MVS_REQUIRE(g.smem <= (size_t)kFftSmemCap && g.threads <= 512, MVS_ERR_UNSUPPORTED,
              "axis length %d: pass geometry out of range", a.n);
  dim3 grid((unsigned)((nlines + g.L - 1) / g.L), n);
  if (a.paired) {
    // lines along the first data axis, indexed (y', x); a CTA owns L/2 lines
    // x0.. and their mirrors (-y', -x)
    MVS_REQUIRE(outer == 1 && sign < 0 && axis < 2, MVS_ERR_INVALID, "paired pass on a wrong axis");
    a.n2 = p->shape[2];
    a.n1p = (int)(inner / p->shape[2]);
    a.items_x = p->shape[2] / 2 + 1;
    a.xblocks = (a.items_x + g.L / 2 - 1) / (g.L / 2);
    a.cp_scales = p->d_cp;  // per pair, set by mvs_pc_load_pairs
    grid.x = (unsigned)((long long)a.xblocks * a.n1p);
  }
  const bool blue = ax.bluestein != 0;
  switch (ax.m) {
    case 1: return launch_pass_m<1>(a, blue, grid, g.threads, g.smem, st);
    case 2: return launch_pass_m<2>(a, blue, grid, g.threads, g.smem, st);
    case 4: return launch_pass_m<4>(a, blue, grid, g.threads, g.smem, st);
    case 8: return launch_pass_m<8>(a, blue, grid, g.threads, g.smem, st);
    case 16: return launch_pass_m<16>(a, blue, grid, g.threads, g.smem, st);
    case 32: return launch_pass_m<32>(a, blue, grid, g.threads, g.smem, st);
    case 64: return launch_pass_m<64>(a, blue, grid, g.threads, g.smem, st);
    case 128: return launch_pass_m<128>(a, blue, grid, g.threads, g.smem, st);
    case 256: return launch_pass_m<256>(a, blue, grid, g.threads, g.smem, st);
    case 512: return launch_pass_m<512>(a, blue, grid, g.threads, g.smem, st);
    case 640: return launch_pass_m<640>(a, blue, grid, g.threads, g.smem, st);
    case 1024: return launch_pass_m<1024>(a, blue, grid, g.threads, g.smem, st);
    case 2048: return launch_pass_m<2048>(a, blue, grid, g.threads, g.smem, st);
    case 4096: return launch_pass_m<4096>(a, blue, grid, g.threads, g.smem, st);
    case 8192: return launch_pass_m<8192>(a, blue, grid, g.threads, g.smem, st);
  }
  set_error("no FFT kernel for length %d", ax.m);
  return MVS_ERR_UNSUPPORTED;
}

template <int R>
static int launch_updft_x(mvs_pc_plan* p, int n, cudaStream_t st) {
  const long long lines = p->N / p->shape[2];
  dim3 grid((unsigned)((lines + kUpWarps - 1) / kUpWarps), n);
  updft_x_kernel<R><<<grid, kUpWarps * 32, 0, st>>>(p->Z, p->shape[0], p->shape[1], p->shape[2],
                                                    p->N, p->d_E, p->e_stride, p->d_G, p->d_T0);
  MVS_CHECK_CUDA(cudaGetLastError());
  return MVS_OK;
}

extern "C" int mvs_pc_correlate(mvs_pc_plan* p, int n, int32_t* peaks_host, double* updft_host,
                                void* stream) {
  MVS_REQUIRE(p && peaks_host && updft_host, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(n >= 1 && n <= p->loaded, MVS_ERR_INVALID, "n = %d but %d pairs loaded", n,
              p->loaded);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  MVS_CHECK_CUDA(cudaMemsetAsync(p->d_keys, 0, sizeof(unsigned long long) * 2 * n, st));
  // forward: x (packing the real pair), [y], then the first data axis fused with
  // the cross-power spectrum: Z -> Q
  if ((rc = launch_pass(p, n, 2, -1, PASS_LOAD_REAL, nullptr, p->Z, st))) return rc;
  if (p->ndim == 3 && (rc = launch_pass(p, n, 1, -1, PASS_PLAIN, p->Z, p->Z, st))) return rc;
  const int first = 3 - p->ndim;
  // (Z is overwritten in place by the plain cross-power spectrum P)
  if ((rc = launch_pass(p, n, first, -1, PASS_PAIRED, p->Z, p->Q, st, p->Z))) return rc;
  // inverse of Q in place: first data axis, [y], x + argmax
  if ((rc = launch_pass(p, n, first, +1, PASS_PLAIN, p->Q, p->Q, st))) return rc;
  if (p->ndim == 3 && (rc = launch_pass(p, n, 1, +1, PASS_PLAIN, p->Q, p->Q, st))) return rc;
  // (test hook: MVS_PC_DEBUG_STORE keeps the correlation surfaces in Q)
  float2* surf = getenv("MVS_PC_DEBUG_STORE") ? p->Q : nullptr;
  if ((rc = launch_pass(p, n, 2, +1, PASS_ARGMAX, p->Q, surf, st))) return rc;
  // slot 0 = |Re| = normalization None, slot 1 = |Im| = "phase"
  updft_setup_kernel<<<dim3(2 * n, 16), 256, 0, st>>>(p->d_keys, p->shape[0], p->shape[1], p->shape[2],
                                            p->ndim, p->R, p->upsample, p->d_peaks, p->d_E,
                                            p->e_stride, p->e_off[1], p->e_off[0], p->d_G);
  MVS_CHECK_CUDA(cudaGetLastError());
  const int R = p->R;
  double2* result = nullptr;
  int rn = 1;  // R^ndim
  if (p->upsample > 1) {
    switch (R) {
      case 2: rc = launch_updft_x<2>(p, n, st); break;
      case 3: rc = launch_updft_x<3>(p, n, st); break;
      case 5: rc = launch_updft_x<5>(p, n, st); break;
      case 6: rc = launch_updft_x<6>(p, n, st); break;
      case 8: rc = launch_updft_x<8>(p, n, st); break;
      case 9: rc = launch_updft_x<9>(p, n, st); break;
      case 11: rc = launch_updft_x<11>(p, n, st); break;
      case 12: rc = launch_updft_x<12>(p, n, st); break;
      case 14: rc = launch_updft_x<14>(p, n, st); break;
      case 15: rc = launch_updft_x<15>(p, n, st); break;
      default:
        set_error("upsampled region size %d not instantiated", R);
        return MVS_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
    // contract y: in [z][y][R] -> out [z][R_y][R_x]
    {
      const int outer = p->shape[0], nn = p->shape[1], inner = R;
      dim3 grid(outer * R, 2 * n);
      updft_axis_kernel<<<grid, 256, 0, st>>>(p->d_T0, p->d_T1, outer, nn, inner, R, p->d_E,
                                              p->e_stride, p->e_off[1]);
      MVS_CHECK_CUDA(cudaGetLastError());
      result = p->d_T1;
      rn = R * R;
    }
    if (p->ndim == 3) {
      const int outer = 1, nn = p->shape[0], inner = R * R;
      dim3 grid(outer * R, 2 * n);
      updft_axis_kernel<<<grid, 256, 0, st>>>(p->d_T1, p->d_T0, outer, nn, inner, R, p->d_E,
                                              p->e_stride, p->e_off[0]);
      MVS_CHECK_CUDA(cudaGetLastError());
      result = p->d_T0;
      rn = R * R * R;
    }
  }
  MVS_CHECK_CUDA(cudaMemcpyAsync(peaks_host, p->d_peaks, sizeof(int) * 6 * n,
                                 cudaMemcpyDeviceToHost, st));
  if (result) {
    // per (pair,norm) block stride differs between the 2-D and 3-D layouts
    const long long blk = (p->ndim == 3) ? (long long)rn : (long long)p->shape[0] * rn;
    MVS_CHECK_CUDA(cudaMemcpy2DAsync(updft_host, sizeof(double2) * rn, result,
                                     sizeof(double2) * blk, sizeof(double2) * rn, 2 * n,
                                     cudaMemcpyDeviceToHost, st));
  }
  MVS_CHECK_CUDA(cudaStreamSynchronize(st));
  return MVS_OK;
}

// test hook: copies pair `pair` of the plan's complex work buffers to the host
// (which = 0: Z / plain cross power P after mvs_pc_correlate, 1: Q after the
// inverse passes that store)
extern "C" int mvs_pc_debug_copy(const mvs_pc_plan* p, int which, int pair, float* host) {
  MVS_REQUIRE(p && host && pair >= 0 && pair < p->max_pairs, MVS_ERR_INVALID, "bad arguments");
  const float2* src = (which == 0 ? p->Z : p->Q) + (long long)pair * p->N;
  MVS_CHECK_CUDA(cudaMemcpy(host, src, sizeof(float2) * p->N, cudaMemcpyDeviceToHost));
  return MVS_OK;
}

extern "C" int mvs_pc_plan_info(const mvs_pc_plan* p, int* region, int64_t* voxels,
                                int* launches_per_correlate) {
  MVS_REQUIRE(p, MVS_ERR_INVALID, "plan is NULL");
  if (region) *region = p->R;
  if (voxels) *voxels = p->N;
  if (launches_per_correlate)
    *launches_per_correlate = 2 * p->ndim + 1 + (p->upsample > 1 ? p->ndim : 0);
  return MVS_OK;
}
