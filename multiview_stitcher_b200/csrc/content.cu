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

#include <climits>
#include <mutex>
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

// ---- fast path of the same pass -------------------------------------------------
// A CTA of 4 warps owns a tile: 32 lanes across a non-filter axis x 128 outputs
// along the filter axis, staged once (with its 2 r' halo rows, reflect-resolved) in
// shared memory as [k][lane] so every read is conflict-free; each warp then produces
// a 32-output sub-range.  A lane computes 4 consecutive outputs at a time from two
// 4-deep rotating register windows (left / right taps): 2 shared-memory reads feed
// 4 x (DADD, DMUL, DADD) -- scipy's exact float64 operation order.  r' = radius
// rounded up to a multiple of 4 with zero weights (x + 0.0 * finite == x).  Tiles
// whose inputs are all one value (the 0/1 validity mask of nan_gaussian_filter away
// from edges, empty space) collapse to one chain.  `boxes` (optional) restricts a
// volume to the bounding box of its valid voxels: inputs outside read as 0, outputs
// outside are not produced (they are exact zeros that nobody reads,
// weights.py:314-320 divides only where the view exists).

#ifndef MVS_GTA
#define MVS_GTA 128
#endif
constexpr int kGTA = MVS_GTA;  // outputs per tile along the filter axis (= threads per CTA)
constexpr int kGSub = 32;     // ... per warp
constexpr int kGWarps = kGTA / kGSub;
constexpr int kGPitch = 33;
constexpr int kGB = 8;        // loads requested per batch

struct GaussArgs {
  const float* in;
  float* out;
  int nz, ny, nx, batch;
  const double* fw;  // [rp + 1], zero beyond the true radius
  int rp;            // radius rounded up to a multiple of 4, >= 4
  const int* boxes;  // [box_mod][6]: lo z,y,x then hi z,y,x (inclusive); nullptr = whole volume
  int box_mod;
};

__device__ __forceinline__ int reflect_idx(int k, int n) {
  // half-sample symmetric reflection: d c b a | a b c d | d c b a
  if (k < 0 || k >= n) {
    const int period = 2 * n;
    k %= period;
    if (k < 0) k += period;
    if (k >= n) k = period - 1 - k;
  }
  return k;
}

#define MVS_GSTEP(JJ, A4)                                                                      \
  {                                                                                            \
    const double w = wts[(JJ)];                                                                \
    _Pragma("unroll") for (int i = 0; i < 4; ++i)                                              \
      acc[i] = __dadd_rn(acc[i], __dmul_rn(__dadd_rn(Lw[(i - (A4)) & 3], Rw[((A4) + i) & 3]), w)); \
    Lw[(4 - (A4)) & 3] = (double)sp[(4 - (JJ)) * kGPitch];                                     \
    Rw[((A4) + 3) & 3] = (double)sp[((JJ)-1) * kGPitch];                                       \
  }

template <int AXIS>
__global__ void __launch_bounds__(kGWarps * 32) gauss_tile_kernel(const GaussArgs A) {
  extern __shared__ double gs_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rp = A.rp, span = kGTA + 2 * rp;
  double* wts = gs_smem;
  float* strip = reinterpret_cast<float*>(wts + rp + 1);
  float* otile = strip + span * kGPitch;  // AXIS == 2 only
  for (int i = tid; i <= rp; i += blockDim.x) wts[i] = A.fw[i];
  const int nA = AXIS == 0 ? A.nz : (AXIS == 1 ? A.ny : A.nx);
  const int nL = AXIS == 2 ? A.ny : A.nx;
  const int nT = AXIS == 0 ? A.ny : A.nz;
  const int tilesA = (nA + kGTA - 1) / kGTA, tilesL = (nL + 31) / 32;
  // tile index decode in 32-bit arithmetic (64-bit divisions cost ~100 instructions each)
  const unsigned per_row = (unsigned)tilesA * (unsigned)tilesL;
  const unsigned per_vol = (unsigned)nT * per_row;
  const unsigned tiles = per_vol * (unsigned)A.batch;  // < 2^31 (checked by the launcher)
  const long long sy = A.nx, sz = (long long)A.ny * A.nx;
  const long long strideA = AXIS == 0 ? sz : (AXIS == 1 ? sy : 1);
  const long long strideT = AXIS == 0 ? sy : sz;
  for (unsigned su = blockIdx.x; su < tiles; su += gridDim.x) {
    const int b = (int)(su / per_vol);
    unsigned rem = su - (unsigned)b * per_vol;
    const int th = (int)(rem / per_row);
    rem -= (unsigned)th * per_row;
    const int ta = (int)(rem / (unsigned)tilesL);
    const int tl = (int)(rem - (unsigned)ta * (unsigned)tilesL);
    const int a0 = ta * kGTA, l0 = tl * 32;
    int loA = 0, hiA = nA - 1, loL = 0, hiL = nL - 1, loT = 0, hiT = nT - 1;
    if (A.boxes) {
      const int* bx = A.boxes + 6 * (b % A.box_mod);
      const int iA = AXIS, iL = AXIS == 2 ? 1 : 2, iT = AXIS == 0 ? 1 : 0;
      loA = max(loA, bx[iA]); hiA = min(hiA, bx[3 + iA]);
      loL = max(loL, bx[iL]); hiL = min(hiL, bx[3 + iL]);
      loT = max(loT, bx[iT]); hiT = min(hiT, bx[3 + iT]);
    }
    if (th < loT || th > hiT || a0 > hiA || a0 + kGTA - 1 < loA || l0 > hiL || l0 + 31 < loL)
      continue;  // nothing of this tile is wanted (CTA-uniform)
    const long long voff = (long long)b * A.nz * sz + (long long)th * strideT;
    const float* vin = A.in + voff;
    float* vout = A.out + voff;
    __syncthreads();  // the previous tile's readers are done (also publishes wts)
    // ---- stage the tile: loads issued kGB at a time (clamped addresses) ----
    bool same = true;
    float first = 0.f;
    if (AXIS != 2) {
      // warp w stages rows w, w + 4, ...; lanes along x
      const int xl = l0 + lane;
      const bool lane_ok = xl >= loL && xl <= hiL;
      const int myrows = (span - warp + kGWarps - 1) / kGWarps;
      for (int i0 = 0; i0 < myrows; i0 += kGB) {
        float tv[kGB];
#pragma unroll
        for (int u = 0; u < kGB; ++u) {
          const int kk = warp + kGWarps * min(i0 + u, myrows - 1);
          const int a = reflect_idx(a0 - rp + kk, nA);
          const bool ok = lane_ok && a >= loA && a <= hiA;
          const float x = __ldg(ok ? vin + (long long)a * strideA + xl : A.in);
          tv[u] = ok ? x : 0.f;
        }
        if (i0 == 0) first = tv[0];
#pragma unroll
        for (int u = 0; u < kGB; ++u) {
          if (i0 + u < myrows) {
            strip[(warp + kGWarps * (i0 + u)) * kGPitch + lane] = tv[u];
            same = same && (tv[u] == first);
          }
        }
      }
    } else {
      // consecutive threads read consecutive x of one row: coalesced; transposed into [k][row]
      const int spad = (span + 31) & ~31;
      const int total = 32 * spad;
      for (int e0 = 0; e0 < total; e0 += kGB * kGTA) {
        float tv[kGB];
#pragma unroll
        for (int u = 0; u < kGB; ++u) {
          const int e = e0 + u * kGTA + tid;
          const int row = e / spad, kk = e - row * spad;
          const int y = l0 + row;
          const int a = reflect_idx(a0 - rp + kk, nA);
          const bool ok = e < total && kk < span && y >= loL && y <= hiL && a >= loA && a <= hiA;
          const float x = __ldg(ok ? vin + (long long)y * sy + a : A.in);
          tv[u] = ok ? x : 0.f;
        }
        if (e0 == 0) first = tv[0];
#pragma unroll
        for (int u = 0; u < kGB; ++u) {
          const int e = e0 + u * kGTA + tid;
          const int row = e / spad, kk = e - row * spad;
          if (e < total && kk < span) {
            strip[kk * kGPitch + row] = tv[u];
            same = same && (tv[u] == first);
          }
        }
      }
    }
    __syncthreads();
    const float f0 = strip[0];
    const bool uniform = __syncthreads_and(same && first == f0) != 0;
    const int lpos = l0 + lane;  // this lane's coordinate on the lane axis
    const bool lane_out = lpos >= loL && lpos <= hiL;
    const int s0 = a0 + warp * kGSub;  // this warp's first output
    if (uniform) {
      double tmp = __dmul_rn((double)f0, wts[0]);
      const double two = __dadd_rn((double)f0, (double)f0);
      for (int j = rp; j >= 1; --j) tmp = __dadd_rn(tmp, __dmul_rn(two, wts[j]));
      const float o = (float)tmp;
      if (AXIS != 2) {
        if (lane_out)
          for (int k = max(s0, loA); k <= min(s0 + kGSub - 1, hiA); ++k)
            vout[(long long)k * strideA + lpos] = o;
      } else {
        const int a = a0 + tid;
        if (a >= loA && a <= hiA)
          for (int row = max(l0, loL); row <= min(l0 + 31, hiL); ++row) vout[(long long)row * sy + a] = o;
      }
      continue;
    }
    for (int g = 0; g < kGSub / 4; ++g) {
      const int c0 = s0 + 4 * g;
      if (c0 > hiA || c0 + 3 < loA) continue;  // warp-uniform
      const float* sp = strip + (rp + warp * kGSub + 4 * g) * kGPitch + lane;
      double acc[4], Lw[4], Rw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i] = __dmul_rn((double)sp[i * kGPitch], wts[0]);
        Lw[i] = (double)sp[(i - rp) * kGPitch];
        Rw[i] = (double)sp[(i + rp) * kGPitch];
      }
      for (int j = rp; j >= 4; j -= 4) {
        MVS_GSTEP(j, 0)
        MVS_GSTEP(j - 1, 3)
        MVS_GSTEP(j - 2, 2)
        MVS_GSTEP(j - 3, 1)
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int a = c0 + i;
        if (AXIS != 2) {
          if (lane_out && a >= loA && a <= hiA) vout[(long long)a * strideA + lpos] = (float)acc[i];
        } else {
          otile[(warp * kGSub + 4 * g + i) * kGPitch + lane] = (float)acc[i];
        }
      }
    }
    if (AXIS == 2) {
      __syncthreads();
      const int a = a0 + tid;  // 128 consecutive x per row: coalesced
      if (a >= loA && a <= hiA)
        for (int row = max(l0, loL); row <= min(l0 + 31, hiL); ++row)
          vout[(long long)row * sy + a] = otile[tid * kGPitch + (row - l0)];
    }
  }
}
#undef MVS_GSTEP

// bounding boxes (lo z,y,x / hi z,y,x) of the non-NaN voxels of every volume
__global__ void box_init_kernel(int* __restrict__ boxes, int V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * V) boxes[i] = (i % 6) < 3 ? INT_MAX : -1;
}

__global__ void __launch_bounds__(256)
box_kernel(const float* __restrict__ vols, int nz, int ny, int nx, int* __restrict__ boxes) {
  const long long N = (long long)nz * ny * nx;
  const float* v = vols + (long long)blockIdx.y * N;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {-1, -1, -1};
  // one warp per 256-voxel row segment: (z, y) per segment, no per-voxel 64-bit divisions
  const unsigned rows = (unsigned)nz * (unsigned)ny, xt = ((unsigned)nx + 255) / 256;
  const unsigned nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (unsigned u = blockIdx.x * nwarps + (threadIdx.x >> 5); u < rows * xt; u += gridDim.x * nwarps) {
    const unsigned row = u / xt, x0 = (u - row * xt) * 256;
    const int z = (int)(row / (unsigned)ny), y = (int)(row - (unsigned)z * (unsigned)ny);
    const float* rp = v + (long long)row * nx;
    const int x1 = min(nx, (int)x0 + 256);
    bool any = false;
    for (int xi = (int)(x0 + lane); xi < x1; xi += 32) {
      const float x = rp[xi];
      if (x == x) { lo[2] = min(lo[2], xi); hi[2] = max(hi[2], xi); any = true; }
    }
    if (any) { lo[0] = min(lo[0], z); hi[0] = max(hi[0], z); lo[1] = min(lo[1], y); hi[1] = max(hi[1], y); }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  if ((threadIdx.x & 31) == 0) {
    int* bx = boxes + 6 * blockIdx.y;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (lo[d] != INT_MAX) atomicMin(bx + d, lo[d]);
      if (hi[d] >= 0) atomicMax(bx + 3 + d, hi[d]);
    }
  }
}

// T = (M - VV/WW)^2 with the NaN pattern of M (weights.py:57-65) split straight
// into the zero-filled values / validity mask of the next nan_gaussian_filter
__global__ void __launch_bounds__(256)
sqdiff_split_kernel(const float* __restrict__ vv, const float* __restrict__ ww,
                    const float* __restrict__ ref, float* __restrict__ v0, float* __restrict__ w0,
                    long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float r = ref[i];
    float t = NAN;
    if (r == r) {
      const float d = __fsub_rn(r, __fdiv_rn(vv[i], ww[i]));
      t = __fmul_rn(d, d);
    }
    const bool nan = t != t;
    v0[i] = nan ? 0.f : t;
    w0[i] = nan ? 0.f : 1.f;
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

static int round_up4(int r) { return r < 4 ? 4 : ((r + 3) / 4) * 4; }

// shared memory of one tile (0 in *fits: use the plain kernel)
static size_t tile_smem(int rp, int axis, bool* fits) {
  const size_t bytes = sizeof(double) * (rp + 1) +
                       sizeof(float) * ((size_t)(kGTA + 2 * rp) * kGPitch + (axis == 2 ? kGTA * kGPitch : 0));
  *fits = bytes <= 200 * 1024;
  return bytes;
}

template <int AXIS>
static int launch_tile(const GaussArgs& a, size_t smem, cudaStream_t st) {
  auto kern = gauss_tile_kernel<AXIS>;
  MVS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nA = AXIS == 0 ? a.nz : (AXIS == 1 ? a.ny : a.nx);
  const int nL = AXIS == 2 ? a.ny : a.nx;
  const int nT = AXIS == 0 ? a.ny : a.nz;
  const long long tiles = (long long)a.batch * nT * ((nA + kGTA - 1) / kGTA) * ((nL + 31) / 32);
  MVS_REQUIRE(tiles < (1LL << 31), MVS_ERR_UNSUPPORTED, "stack too large for one Gaussian pass");
  long long blocks = tiles;
  const long long cap = 148LL * 16;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, kGWarps * 32, smem, st>>>(a);
  MVS_CHECK_CUDA(cudaGetLastError());
  return MVS_OK;
}

// gaussian_filter of `batch` volumes: src -> dst, using tmp as the ping-pong buffer.
// d_fw holds rp + 1 weights (zero beyond `radius`).  boxes: see gauss_tile_kernel.
static int gaussian_batch(const float* src, float* dst, float* tmp, int batch, const int32_t shape[3],
                          int ndim, const double* d_fw, int radius, int rp, const int* boxes,
                          int box_mod, cudaStream_t st) {
  const long long total = (long long)batch * shape[0] * shape[1] * shape[2];
  const int grid = ew_grid(total);
  // axes in scipy's order (z, y, x); an odd number of passes must end in dst
  const int first_axis = 3 - ndim;
  const float* cur = src;
  for (int axis = first_axis; axis < 3; ++axis) {
    const int remaining = 3 - axis;  // passes left including this one
    float* o = (remaining % 2 == 1) ? dst : tmp;
    bool fits = false;
    const size_t smem = tile_smem(rp, axis, &fits);
    if (fits) {
      GaussArgs a{cur, o, shape[0], shape[1], shape[2], batch, d_fw, rp, boxes, box_mod};
      int rc = axis == 0 ? launch_tile<0>(a, smem, st)
                         : (axis == 1 ? launch_tile<1>(a, smem, st) : launch_tile<2>(a, smem, st));
      if (rc) return rc;
    } else {
      // very wide kernels: one thread per output, taps from global memory (computes
      // the whole volume, which is a superset of any box)
      gauss1d_kernel<<<grid, 256, 0, st>>>(cur, o, shape[0], shape[1], shape[2], axis, d_fw, radius,
                                           total);
      MVS_CHECK_CUDA(cudaGetLastError());
    }
    cur = o;
  }
  return MVS_OK;
}

// Grow-only device workspaces of the multi-pass entry points, one per stream in use (up to
// kWsSlots; least recently used beyond that).  Calls are serialised by g_ws_mutex while they
// ENQUEUE; on one stream the work of two calls is ordered by the stream itself, so nothing waits
// for the GPU, and calls on different streams (e.g. the chunks of a content-weighted fusion
// alternating between two streams, so that the FP64-bound Gaussian passes of one chunk overlap the
// HBM-bound element-wise passes of the next) do not share a buffer.  A slot that changes hands first
// waits for its previous stream; growing a buffer frees it (cudaFree synchronises the device).
static std::mutex g_ws_mutex;
constexpr int kWsSlots = 3;
struct WsSlot {
  void* p = nullptr;
  size_t bytes = 0;
  cudaStream_t st = nullptr;
  bool used = false;
  unsigned long long tick = 0;
};
static WsSlot g_ws[kWsSlots];
static unsigned long long g_ws_tick = 0;

static int workspace(size_t bytes, void** out, cudaStream_t st) {
  WsSlot* e = nullptr;
  for (auto& w : g_ws)
    if (w.used && w.st == st) e = &w;
  if (!e) {
    for (auto& w : g_ws)
      if (!w.used) { e = &w; break; }
    if (!e) {
      e = &g_ws[0];
      for (auto& w : g_ws)
        if (w.tick < e->tick) e = &w;
      if (cudaStreamSynchronize(e->st) != cudaSuccess) cudaGetLastError();  // a stream that no longer exists
    }
    e->st = st;
    e->used = true;
  }
  e->tick = ++g_ws_tick;
  if (e->bytes < bytes) {
    if (e->p) cudaFree(e->p);
    e->p = nullptr;
    e->bytes = 0;
    cudaError_t err = cudaMalloc(&e->p, bytes);
    if (err != cudaSuccess) {
      set_error("workspace of %zu bytes: %s", bytes, cudaGetErrorString(err));
      return MVS_ERR_CUDA;
    }
    e->bytes = bytes;
  }
  *out = e->p;
  return MVS_OK;
}

static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

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
  const int rp = round_up4(radius);
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  void* ws = nullptr;
  const size_t wbytes = align256(sizeof(double) * (rp + 1));
  if ((rc = workspace(wbytes + sizeof(float) * total, &ws, st))) return rc;
  double* d_fw = (double*)ws;
  float* tmp = (float*)((char*)ws + wbytes);
  std::vector<double> hw(rp + 1, 0.0);
  for (int i = 0; i <= radius; ++i) hw[i] = weights[i];
  MVS_CHECK_CUDA(cudaMemcpyAsync(d_fw, hw.data(), sizeof(double) * (rp + 1), cudaMemcpyHostToDevice, st));
  return gaussian_batch(d_in, d_out, tmp, batch, shape, ndim, d_fw, radius, rp, nullptr, 1, st);
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
  const int rp1 = round_up4(r1), rp2 = round_up4(r2);
  // workspace: filter weights, per-view boxes, masked views M, a batch of 2V
  // volumes [V0 | W0] and two filter buffers of the same size
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  const size_t wbytes = align256(sizeof(double) * (rp1 + rp2 + 2));
  const size_t bbytes = align256(sizeof(int) * 6 * V);
  void* wsp = nullptr;
  if ((rc = workspace(wbytes + bbytes + sizeof(float) * total * 7, &wsp, st))) return rc;
  double* d_w1 = (double*)wsp;
  double* d_w2 = d_w1 + rp1 + 1;
  int* d_boxes = (int*)((char*)wsp + wbytes);
  float* ws = (float*)((char*)wsp + wbytes + bbytes);
  float* M = ws;                // masked views
  float* VW = ws + total;       // [V0 | W0]  (2*total)
  float* F = ws + 3 * total;    // filtered   (2*total)
  float* T = ws + 5 * total;    // ping-pong  (2*total)
  std::vector<double> hw(rp1 + rp2 + 2, 0.0);
  for (int i = 0; i <= r1; ++i) hw[i] = w1[i];
  for (int i = 0; i <= r2; ++i) hw[rp1 + 1 + i] = w2[i];
  MVS_CHECK_CUDA(cudaMemcpyAsync(d_w1, hw.data(), sizeof(double) * hw.size(), cudaMemcpyHostToDevice, st));
  auto fail = [&](int code) { return code; };  // stream-ordered: nothing to wait for

  // transformed_views[blending_weights < 1e-7] = NaN; split into V0 / W0
  mask_split_kernel<<<grid, 256, 0, st>>>(d_views, d_blending, 1e-7f, M, VW, VW + total, total);
  // per-view bounding box of what is left: all filtering is confined to it
  box_init_kernel<<<(6 * V + 255) / 256, 256, 0, st>>>(d_boxes, V);
  {
    const long long bb = (N + 255) / 256;
    dim3 bg((unsigned)(bb < 296 ? (bb < 1 ? 1 : bb) : 296), V);
    box_kernel<<<bg, 256, 0, st>>>(M, shape[0], shape[1], shape[2], d_boxes);
  }
  // inner nan-gaussian (sigma_1), squared difference, split again
  if ((rc = gaussian_batch(VW, F, T, 2 * V, shape, ndim, d_w1, r1, rp1, d_boxes, V, st))) return fail(rc);
  sqdiff_split_kernel<<<grid, 256, 0, st>>>(F, F + total, M, VW, VW + total, total);
  // outer nan-gaussian (sigma_2); the NaN pattern is still that of M
  if ((rc = gaussian_batch(VW, F, T, 2 * V, shape, ndim, d_w2, r2, rp2, d_boxes, V, st))) return fail(rc);
  nan_divide_kernel<false><<<grid, 256, 0, st>>>(F, F + total, M, d_out_weights, total);
  normalize_weights_kernel<<<ew_grid(N), 256, 0, st>>>(d_out_weights, V, N);
  cudaError_t e = cudaGetLastError();
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
