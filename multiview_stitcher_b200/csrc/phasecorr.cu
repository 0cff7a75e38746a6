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
//   Q        = P + i * Pn;   cc = IFFT_nd(Q):  Re = cc(None), Im = cc("phase")
//              (P, Pn are Hermitian, so both correlation surfaces are real and
//               ONE complex inverse transform yields both)
//   peaks    = first argmax of |Re cc|, |Im cc|
//   C        = conj(upsampled_dft(conj(P or Pn))) on ceil(1.5 u)^ndim samples
//              around each peak (float64 accumulation)
// HBM traffic: ndim forward + ndim inverse passes of 8N+8N bytes plus the
// cross-power pass -- the (32 ndim + 8) N "pass model" of SURVEY.md 8d.

#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "fft.cuh"

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

const AxisFft* get_axis_fft(int n) {
  std::lock_guard<std::mutex> lock(g_fft_mutex);
  auto it = g_fft_cache.find(n);
  if (it != g_fft_cache.end()) return it->second;
  if (n < 1) { set_error("FFT length %d", n); return nullptr; }
  AxisFft ax{};
  ax.n = n;
  if (is_pow2(n)) { ax.m = n; ax.bluestein = 0; }
  else { int m = 1; while (m < 2 * n - 1) m <<= 1; ax.m = m; ax.bluestein = 1; }
  if (ax.m > 16384) {
    set_error("axis length %d needs a %d-point shared-memory FFT (max 16384)", n, ax.m);
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
    host_fft_pow2(br, bi);
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

int fft_lines_per_cta(const AxisFft& ax, bool contiguous) {
  int L = 2048 / ax.m;           // 32 KB for the two buffers -> several CTAs per SM
  if (L < 1) L = 1;
  if (L > 16) L = 16;
  if (!contiguous && L < 4 && ax.m <= 1024) L = 4;  // 32-byte runs on strided axes
  if (!contiguous && L < 2 && ax.m <= 4096) L = 2;
  return L;
}

size_t fft_smem_bytes(const AxisFft& ax, int L, int* mp_out) {
  int pad = L > 1 ? (16 / L > 0 ? 16 / L : 1) : 0;
  int mp = ax.m + pad;
  if (mp_out) *mp_out = mp;
  return sizeof(float2) * 2 * (size_t)L * mp;
}

// ---------------------------------------------------------------------------
// FFT pass kernel
// ---------------------------------------------------------------------------

struct FftPass {
  float2* data;        // complex volume(s), in place
  const float* re;     // LOAD_REAL: real sources (NaN -> 0)
  const float* im;
  long long outer, inner;  // lines = outer * inner; element k at o*n*inner + in + k*inner
  long long batch_stride;  // elements between pairs
  int n, L, mp;
  AxisFft ax;
};

template <bool LOAD_REAL, int SIGN>
__global__ void __launch_bounds__(256) fft_pass_kernel(FftPass P) {
  extern __shared__ float2 smem[];
  const int L = P.L, n = P.n, m = P.ax.m, mp = P.mp;
  float2* a = smem;
  float2* b = smem + (size_t)L * mp;
  const long long nlines = P.outer * P.inner;
  const long long line0 = (long long)blockIdx.x * L;
  const long long boff = (long long)blockIdx.y * P.batch_stride;
  const bool contiguous = (P.inner == 1);

  for (int idx = threadIdx.x; idx < L * m; idx += blockDim.x) {
    int line, k;
    if (contiguous) { line = idx / m; k = idx - line * m; }
    else { k = idx / L; line = idx - k * L; }
    float2 v = make_float2(0.f, 0.f);
    const long long q = line0 + line;
    if (k < n && q < nlines) {
      const long long o = q / P.inner, in = q - o * P.inner;
      const long long g = boff + o * (long long)n * P.inner + in + (long long)k * P.inner;
      if (LOAD_REAL) {
        float x = __ldg(P.re + g), y = __ldg(P.im + g);
        v = make_float2(x != x ? 0.f : x, y != y ? 0.f : y);
      } else {
        v = P.data[g];
      }
    }
    a[line * mp + k] = v;
  }
  __syncthreads();
  float2* res = fft_lines(a, b, P.ax, mp, L, SIGN);
  for (int idx = threadIdx.x; idx < L * n; idx += blockDim.x) {
    int line, k;
    if (contiguous) { line = idx / n; k = idx - line * n; }
    else { k = idx / L; line = idx - k * L; }
    const long long q = line0 + line;
    if (q < nlines) {
      const long long o = q / P.inner, in = q - o * P.inner;
      const long long g = boff + o * (long long)n * P.inner + in + (long long)k * P.inner;
      P.data[g] = res[line * mp + k];
    }
  }
}

// ---------------------------------------------------------------------------
// element-wise kernels
// ---------------------------------------------------------------------------

constexpr int kRedBlocks = 64;  // partial-reduction blocks per image

// per-image partial min / max / NaN count / bbox of non-NaN voxels
__global__ void __launch_bounds__(256)
stats_kernel(const float* const* __restrict__ imgs, long long N, int n1, int n2,
             double* __restrict__ partial /* [img][block][9] */) {
  const float* im = imgs[blockIdx.y];
  float mn = INFINITY, mx = -INFINITY;
  long long nan = 0;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {-1, -1, -1};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    float v = __ldg(im + i);
    if (v != v) { ++nan; continue; }
    mn = fminf(mn, v); mx = fmaxf(mx, v);
    int x = (int)(i % n2), y = (int)((i / n2) % n1), z = (int)(i / ((long long)n1 * n2));
    lo[0] = min(lo[0], z); lo[1] = min(lo[1], y); lo[2] = min(lo[2], x);
    hi[0] = max(hi[0], z); hi[1] = max(hi[1], y); hi[2] = max(hi[2], x);
  }
  __shared__ float s_mn[256], s_mx[256];
  __shared__ long long s_nan[256];
  __shared__ int s_lo[3][256], s_hi[3][256];
  const int t = threadIdx.x;
  s_mn[t] = mn; s_mx[t] = mx; s_nan[t] = nan;
  for (int d = 0; d < 3; ++d) { s_lo[d][t] = lo[d]; s_hi[d][t] = hi[d]; }
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (t < s) {
      s_mn[t] = fminf(s_mn[t], s_mn[t + s]); s_mx[t] = fmaxf(s_mx[t], s_mx[t + s]);
      s_nan[t] += s_nan[t + s];
      for (int d = 0; d < 3; ++d) {
        s_lo[d][t] = min(s_lo[d][t], s_lo[d][t + s]);
        s_hi[d][t] = max(s_hi[d][t], s_hi[d][t + s]);
      }
    }
    __syncthreads();
  }
  if (t == 0) {
    double* p = partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 9;
    p[0] = s_mn[0]; p[1] = s_mx[0]; p[2] = (double)s_nan[0];
    for (int d = 0; d < 3; ++d) { p[3 + d] = s_lo[d][0]; p[6 + d] = s_hi[d][0]; }
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

// Q_k = P_k + i Pn_k from the packed spectrum Z (see file header).
__global__ void __launch_bounds__(256)
cross_power_kernel(const float2* __restrict__ Z, float2* __restrict__ Q, int n0, int n1, int n2,
                   long long N) {
  const float2* z = Z + (long long)blockIdx.y * N;
  float2* q = Q + (long long)blockIdx.y * N;
  const float tiny = 100.0f * 1.1920929e-07f;  // 100 * eps(float32)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % n2), y = (int)((i / n2) % n1), zz = (int)(i / ((long long)n1 * n2));
    int mx = x ? n2 - x : 0, my = y ? n1 - y : 0, mz = zz ? n0 - zz : 0;
    float2 a = z[i];
    float2 bm = z[((long long)mz * n1 + my) * n2 + mx];
    float2 bc = make_float2(bm.x, -bm.y);                     // conj Z_-k
    float2 F = make_float2(0.5f * (a.x + bc.x), 0.5f * (a.y + bc.y));
    float2 d = make_float2(a.x - bc.x, a.y - bc.y);
    float2 M = make_float2(0.5f * d.y, -0.5f * d.x);          // d / (2i)
    float2 P = make_float2(F.x * M.x + F.y * M.y, F.y * M.x - F.x * M.y);  // F conj(M)
    float mag = fmaxf(hypotf(P.x, P.y), tiny);
    float2 Pn = make_float2(__fdiv_rn(P.x, mag), __fdiv_rn(P.y, mag));
    q[i] = make_float2(P.x - Pn.y, P.y + Pn.x);
  }
}

// first argmax of |Re| (slot 0: normalization None) and |Im| (slot 1: "phase")
__global__ void __launch_bounds__(256)
argmax_kernel(const float2* __restrict__ Q, long long N, unsigned long long* __restrict__ keys) {
  const float2* q = Q + (long long)blockIdx.y * N;
  unsigned long long k0 = 0, k1 = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    float2 v = q[i];
    unsigned long long lowbits = 0xffffffffull - (unsigned long long)i;
    unsigned long long a = ((unsigned long long)__float_as_uint(fabsf(v.x)) << 32) | lowbits;
    unsigned long long b = ((unsigned long long)__float_as_uint(fabsf(v.y)) << 32) | lowbits;
    k0 = a > k0 ? a : k0;
    k1 = b > k1 ? b : k1;
  }
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long a = __shfl_xor_sync(0xffffffffu, k0, o);
    unsigned long long b = __shfl_xor_sync(0xffffffffu, k1, o);
    k0 = a > k0 ? a : k0;
    k1 = b > k1 ? b : k1;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(keys + 2 * blockIdx.y + 0, k0);
    atomicMax(keys + 2 * blockIdx.y + 1, k1);
  }
}

// ---------------------------------------------------------------------------
// upsampled DFT around the integer peaks
// ---------------------------------------------------------------------------

// Decodes the peaks, applies skimage's wrap (shift > fix(n/2) -> shift - n) and
// fills E[pair][norm][d][a][k] = exp(+2 pi i (a - off_d) ks(k) / (n_d u)),
// off_d = fix(R/2) - shift_d * u,  ks = numpy.fft.fftfreq ordering.
__global__ void __launch_bounds__(256)
updft_setup_kernel(const unsigned long long* __restrict__ keys, int n0, int n1, int n2,
                   int ndim, int R, int u, int* __restrict__ peaks /* [pair][2][3] */,
                   float2* __restrict__ E, long long e_stride /* per (pair,norm) */,
                   int e_off1, int e_off2) {
  const int pn = blockIdx.x;  // pair*2 + norm
  const unsigned long long key = keys[pn];
  const long long idx = (long long)(0xffffffffull - (key & 0xffffffffull));
  int p[3];
  p[2] = (int)(idx % n2); p[1] = (int)((idx / n2) % n1); p[0] = (int)(idx / ((long long)n1 * n2));
  const int nn[3] = {n0, n1, n2};
  int sh[3];
  for (int d = 0; d < 3; ++d) sh[d] = p[d] > nn[d] / 2 ? p[d] - nn[d] : p[d];
  if (threadIdx.x < 3) peaks[pn * 3 + threadIdx.x] = sh[threadIdx.x];
  const int eoff[3] = {0, e_off1, e_off2};
  for (int d = 3 - ndim; d < 3; ++d) {
    const int n = nn[d];
    const double off = (double)(R / 2) - (double)sh[d] * u;
    float2* e = E + (long long)pn * e_stride + eoff[d];
    for (int i = threadIdx.x; i < R * n; i += blockDim.x) {
      const int a = i / n, k = i - a * n;
      const int ks = (k <= (n - 1) / 2) ? k : k - n;
      double s, c;
      sincospi(2.0 * ((double)a - off) * (double)ks / ((double)n * (double)u), &s, &c);
      e[i] = make_float2((float)c, (float)s);
    }
  }
}

// Contract the x axis: T[pn][line][b] = sum_x Pnorm[line, x] * Ex[b][x],
// one warp per line, float partials per lane, float64 across lanes.
template <int R>
__global__ void __launch_bounds__(256)
updft_x_kernel(const float2* __restrict__ Z, int n0, int n1, int n2, long long N,
               const float2* __restrict__ E, long long e_stride, int e_off2,
               double2* __restrict__ T) {
  const int pair = blockIdx.y, norm = blockIdx.z;
  const int pn = pair * 2 + norm;
  const long long nlines = (long long)n0 * n1;
  const long long line = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (line >= nlines) return;
  const int lane = threadIdx.x & 31;
  const int y = (int)(line % n1), zz = (int)(line / n1);
  const int my = y ? n1 - y : 0, mz = zz ? n0 - zz : 0;
  const float2* z = Z + (long long)pair * N;
  const float2* row = z + line * n2;
  const float2* mrow = z + ((long long)mz * n1 + my) * n2;
  const float2* ex = E + (long long)pn * e_stride + e_off2;
  const float tiny = 100.0f * 1.1920929e-07f;
  float2 acc[R];
#pragma unroll
  for (int b = 0; b < R; ++b) acc[b] = make_float2(0.f, 0.f);
  for (int x = lane; x < n2; x += 32) {
    const int mx = x ? n2 - x : 0;
    float2 a = row[x], bm = mrow[mx];
    float2 bc = make_float2(bm.x, -bm.y);
    float2 F = make_float2(0.5f * (a.x + bc.x), 0.5f * (a.y + bc.y));
    float2 d = make_float2(a.x - bc.x, a.y - bc.y);
    float2 M = make_float2(0.5f * d.y, -0.5f * d.x);
    float2 P = make_float2(F.x * M.x + F.y * M.y, F.y * M.x - F.x * M.y);
    if (norm == 1) {
      float mag = fmaxf(hypotf(P.x, P.y), tiny);
      P = make_float2(__fdiv_rn(P.x, mag), __fdiv_rn(P.y, mag));
    }
#pragma unroll
    for (int b = 0; b < R; ++b) {
      float2 e = __ldg(ex + b * n2 + x);
      acc[b].x += P.x * e.x - P.y * e.y;
      acc[b].y += P.x * e.y + P.y * e.x;
    }
  }
#pragma unroll
  for (int b = 0; b < R; ++b) {
    double re = acc[b].x, im = acc[b].y;
    for (int o = 16; o > 0; o >>= 1) {
      re += __shfl_xor_sync(0xffffffffu, re, o);
      im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if (lane == 0) T[((long long)pn * nlines + line) * R + b] = make_double2(re, im);
  }
}

// Contract one more axis: out[pn][o][a][r] = sum_k E[a][k] * in[pn][o][k][r]
// (one thread per output, serial float64 sum -> deterministic).
__global__ void __launch_bounds__(128)
updft_axis_kernel(const double2* __restrict__ in, double2* __restrict__ out, int outer, int n,
                  int inner, int R, const float2* __restrict__ E, long long e_stride, int e_off) {
  const int pn = blockIdx.y;
  const int total = outer * R * inner;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int r = i % inner, a = (i / inner) % R, o = i / (inner * R);
  const float2* e = E + (long long)pn * e_stride + e_off + (long long)a * n;
  const double2* src = in + ((long long)pn * outer + o) * n * inner + r;
  double re = 0.0, im = 0.0;
  for (int k = 0; k < n; ++k) {
    double2 v = src[(long long)k * inner];
    float2 w = __ldg(e + k);
    re += v.x * w.x - v.y * w.y;
    im += v.x * w.y + v.y * w.x;
  }
  out[((long long)pn * outer + o) * R * inner + (long long)a * inner + r] = make_double2(re, im);
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
  unsigned long long* d_keys = nullptr;
  int* d_peaks = nullptr;
  float2* d_E = nullptr;
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
  cudaFree((void*)p->d_imgs); cudaFree(p->d_partial); cudaFree(p->d_mn); cudaFree(p->d_scale);
  cudaFree(p->d_keys); cudaFree(p->d_peaks); cudaFree(p->d_E); cudaFree(p->d_T0);
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
  int off = 0;
  for (int d = 3 - ndim; d < 3; ++d) { p->e_off[d] = off; off += p->R * shape[d]; }
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
  alloc((void**)&p->d_partial, sizeof(double) * 9 * kRedBlocks * 2 * max_pairs);
  alloc((void**)&p->d_mn, sizeof(float) * 2 * max_pairs);
  alloc((void**)&p->d_scale, sizeof(float) * 2 * max_pairs);
  alloc((void**)&p->d_keys, sizeof(unsigned long long) * 2 * max_pairs);
  alloc((void**)&p->d_peaks, sizeof(int) * 6 * max_pairs);
  alloc((void**)&p->d_E, sizeof(float2) * p->e_stride * 2 * max_pairs);
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
  std::vector<double> part((size_t)9 * kRedBlocks * 2 * n);
  MVS_CHECK_CUDA(cudaMemcpyAsync(part.data(), p->d_partial, sizeof(double) * part.size(),
                                 cudaMemcpyDeviceToHost, st));
  MVS_CHECK_CUDA(cudaStreamSynchronize(st));
  std::vector<float> mn(2 * n), sc(2 * n);
  for (int img = 0; img < 2 * n; ++img) {
    double* s = stats_host + (size_t)img * 9;
    s[0] = INFINITY; s[1] = -INFINITY; s[2] = 0;
    for (int d = 0; d < 3; ++d) { s[3 + d] = 2147483647.0; s[6 + d] = -1; }
    for (int b = 0; b < kRedBlocks; ++b) {
      const double* q = part.data() + ((size_t)img * kRedBlocks + b) * 9;
      s[0] = std::min(s[0], q[0]); s[1] = std::max(s[1], q[1]); s[2] += q[2];
      for (int d = 0; d < 3; ++d) {
        s[3 + d] = std::min(s[3 + d], q[3 + d]);
        s[6 + d] = std::max(s[6 + d], q[6 + d]);
      }
    }
    mn[img] = (float)s[0];
    // (imax - imin) evaluated in float64, used as a float32 divisor (skimage)
    sc[img] = (s[1] > s[0]) ? (float)(s[1] - s[0]) : 0.0f;
  }
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

template <bool LOAD_REAL, int SIGN>
static int launch_pass(mvs_pc_plan* p, int n, int axis, float2* data, cudaStream_t st) {
  FftPass a{};
  a.data = data;
  a.re = p->r0; a.im = p->r1;
  a.n = p->shape[axis];
  a.ax = p->ax[axis];
  long long inner = 1, outer = 1;
  for (int d = axis + 1; d < 3; ++d) inner *= p->shape[d];
  for (int d = 0; d < axis; ++d) outer *= p->shape[d];
  a.inner = inner; a.outer = outer;
  a.batch_stride = p->N;
  a.L = fft_lines_per_cta(a.ax, inner == 1);
  size_t smem = fft_smem_bytes(a.ax, a.L, &a.mp);
  MVS_REQUIRE(smem <= 227 * 1024, MVS_ERR_UNSUPPORTED,
              "axis length %d needs %zu bytes of shared memory", a.n, smem);
  auto kern = fft_pass_kernel<LOAD_REAL, SIGN>;
  MVS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  const long long nlines = outer * inner;
  dim3 grid((unsigned)((nlines + a.L - 1) / a.L), n);
  kern<<<grid, 256, smem, st>>>(a);
  MVS_CHECK_CUDA(cudaGetLastError());
  return MVS_OK;
}

template <int R>
static int launch_updft_x(mvs_pc_plan* p, int n, cudaStream_t st) {
  const long long lines = p->N / p->shape[2];
  dim3 grid((unsigned)((lines + 7) / 8), n, 2);
  updft_x_kernel<R><<<grid, 256, 0, st>>>(p->Z, p->shape[0], p->shape[1], p->shape[2], p->N,
                                          p->d_E, p->e_stride, p->e_off[2], p->d_T0);
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
  // forward: x (from the real pair), then y, then z
  if ((rc = launch_pass<true, -1>(p, n, 2, p->Z, st))) return rc;
  if ((rc = launch_pass<false, -1>(p, n, 1, p->Z, st))) return rc;
  if (p->ndim == 3 && (rc = launch_pass<false, -1>(p, n, 0, p->Z, st))) return rc;
  dim3 g(grid_for(p->N), n);
  cross_power_kernel<<<g, 256, 0, st>>>(p->Z, p->Q, p->shape[0], p->shape[1], p->shape[2], p->N);
  MVS_CHECK_CUDA(cudaGetLastError());
  // inverse: z, y, x
  if (p->ndim == 3 && (rc = launch_pass<false, +1>(p, n, 0, p->Q, st))) return rc;
  if ((rc = launch_pass<false, +1>(p, n, 1, p->Q, st))) return rc;
  if ((rc = launch_pass<false, +1>(p, n, 2, p->Q, st))) return rc;
  MVS_CHECK_CUDA(cudaMemsetAsync(p->d_keys, 0, sizeof(unsigned long long) * 2 * n, st));
  argmax_kernel<<<g, 256, 0, st>>>(p->Q, p->N, p->d_keys);
  MVS_CHECK_CUDA(cudaGetLastError());
  // slot 0 = |Re| = normalization None, slot 1 = |Im| = "phase"
  updft_setup_kernel<<<2 * n, 256, 0, st>>>(p->d_keys, p->shape[0], p->shape[1], p->shape[2],
                                            p->ndim, p->R, p->upsample, p->d_peaks, p->d_E,
                                            p->e_stride, p->e_off[1], p->e_off[2]);
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
      dim3 grid((outer * R * inner + 127) / 128, 2 * n);
      updft_axis_kernel<<<grid, 128, 0, st>>>(p->d_T0, p->d_T1, outer, nn, inner, R, p->d_E,
                                              p->e_stride, p->e_off[1]);
      MVS_CHECK_CUDA(cudaGetLastError());
      result = p->d_T1;
      rn = R * R;
    }
    if (p->ndim == 3) {
      const int outer = 1, nn = p->shape[0], inner = R * R;
      dim3 grid((outer * R * inner + 127) / 128, 2 * n);
      updft_axis_kernel<<<grid, 128, 0, st>>>(p->d_T1, p->d_T0, outer, nn, inner, R, p->d_E,
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

extern "C" int mvs_pc_plan_info(const mvs_pc_plan* p, int* region, int64_t* voxels,
                                int* launches_per_correlate) {
  MVS_REQUIRE(p, MVS_ERR_INVALID, "plan is NULL");
  if (region) *region = p->R;
  if (voxels) *voxels = p->N;
  if (launches_per_correlate)
    *launches_per_correlate = 2 * p->ndim + 3 + (p->upsample > 1 ? p->ndim : 0);
  return MVS_OK;
}
