// Fused affine resample + blending weights + weighted accumulation (sm_100a).
//
// One launch covers every output chunk of a plan.  Per output voxel and per
// contributing view the kernel evaluates, in registers:
//   * the sample position  x = M * o + off  in float64 with scipy's operation
//     order (ni_interpolation.c NI_GeometricTransform: sequential
//     multiply-adds over the output coordinates, then the shift), the
//     "outside" predicate x < 0 || x > len-1 and floor() -- so validity and
//     nearest-neighbour picks are bit-identical to
//     scipy.ndimage.affine_transform(mode="constant") as called by
//     transformation.py:136-139;
//   * order-0 / order-1 interpolation of the view (float32 arithmetic);
//   * the blending weight: multilinear lookup in the view's 5^ndim EDT support
//     table at  u = Mw * o + offw  (weights.py:465-481), cosine ramp and clip
//     (weights.py:502-509) -- no weight volume ever touches HBM;
//   * mask, normalisation and the fusion function with the reference's
//     float32 operation order (fusion/_core.py:1647-1649, weights.py:340-345,
//     fusion/_core.py:61-94 / :42-58 / :97-131), NaN->0 and the output cast
//     (fusion/_core.py:1713).
// HBM traffic is the algorithmic minimum: every contributing view voxel is read
// (L1/L2 absorb the 2^ndim-tap reuse), every output voxel written once.

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mvs {

constexpr int kBX = 128;      // output block extent along x
constexpr int kBY = 8;        // ... along y (one warp per row)
constexpr int kThreads = 256;
constexpr int kVPT = kBX / 32;  // voxels per thread, strided by 32 along x
constexpr int kMaxXforms = 1024;  // per chunk

}  // namespace mvs
#include "fuse_stencil.cuh"
#include "fuse_stencil3.cuh"
namespace mvs {

struct ViewEval {
  float v;  // interpolated value (undefined if !valid)
  float b;  // blending weight (0 if !valid)
  bool valid;
};

__device__ __forceinline__ double affine_coord(const double* __restrict__ m, double off,
                                               double pre, double cx) {
  // ((c0*m0 + c1*m1) + c2*m2) + shift, no FMA contraction (matches scipy/C)
  return __dadd_rn(__dadd_rn(pre, __dmul_rn(cx, m[2])), off);
}

template <int NDIM>
__device__ __forceinline__ double affine_pre(const double* __restrict__ m, double cz,
                                             double cy) {
  if (NDIM == 3) return __dadd_rn(__dmul_rn(cz, m[0]), __dmul_rn(cy, m[1]));
  return __dmul_rn(cy, m[1]);
}

__device__ __forceinline__ float lerp(float a, float b, float t) {
  return fmaf(t, b, fmaf(-t, a, a));
}

// Blending weight at table coordinates (uz, uy, ux); table is 5^NDIM, cval 0.
template <int NDIM>
__device__ __forceinline__ float blend_weight(const float* __restrict__ tab, double uz,
                                              double uy, double ux) {
  if (ux < 0.0 || ux > 4.0 || uy < 0.0 || uy > 4.0) return 0.0f;
  if (NDIM == 3 && (uz < 0.0 || uz > 4.0)) return 0.0f;
  int ix = __double2int_rd(ux), iy = __double2int_rd(uy);  // in [0, 4]: round-down == floor
  float tx = (float)(ux - (double)ix), ty = (float)(uy - (double)iy);
  int ix1 = min(ix + 1, 4), iy1 = min(iy + 1, 4);
  float w;
  if (NDIM == 3) {
    int iz = __double2int_rd(uz);
    float tz = (float)(uz - (double)iz);
    int iz1 = min(iz + 1, 4);
    const float* p0 = tab + iz * 25;
    const float* p1 = tab + iz1 * 25;
    float a0 = lerp(lerp(__ldg(p0 + iy * 5 + ix), __ldg(p0 + iy * 5 + ix1), tx),
                    lerp(__ldg(p0 + iy1 * 5 + ix), __ldg(p0 + iy1 * 5 + ix1), tx), ty);
    float a1 = lerp(lerp(__ldg(p1 + iy * 5 + ix), __ldg(p1 + iy * 5 + ix1), tx),
                    lerp(__ldg(p1 + iy1 * 5 + ix), __ldg(p1 + iy1 * 5 + ix1), tx), ty);
    w = lerp(a0, a1, tz);
  } else {
    w = lerp(lerp(__ldg(tab + iy * 5 + ix), __ldg(tab + iy * 5 + ix1), tx),
             lerp(__ldg(tab + iy1 * 5 + ix), __ldg(tab + iy1 * 5 + ix1), tx), ty);
  }
  // weights.py:502-507: x<1 -> (cos((1-x)*pi)+1)/2 in float32, then clip
  if (w < 1.0f) {
    float a = __fmul_rn(__fsub_rn(1.0f, w), 3.14159274101257324f);
    w = __fdiv_rn(__fadd_rn(cosf(a), 1.0f), 2.0f);
  }
  return fminf(fmaxf(w, 0.0f), 1.0f);
}

// Order-0 / order-1 sample of view X at the (valid) window position (xz, xy, xx):
// scipy's tap indices (edge tap clamped), float32 interpolation.
template <int NDIM, int ORDER, typename T>
__device__ __forceinline__ float sample_view(const mvs_view_xform& X, double xz, double xy, double xx) {
  const T* base = reinterpret_cast<const T*>(X.data);
  const int nz = X.shape[0], ny = X.shape[1], nx = X.shape[2];
  const int64_t sz = X.stride[0], sy = X.stride[1], sx = X.stride[2];
  if (ORDER == 0) {
    const int64_t ix = (int64_t)floor(__dadd_rn(xx, 0.5));
    const int64_t iy = (int64_t)floor(__dadd_rn(xy, 0.5));
    const int64_t iz = NDIM == 3 ? (int64_t)floor(__dadd_rn(xz, 0.5)) : 0;
    return (float)__ldg(base + iz * sz + iy * sy + ix * sx);
  }
  // positions are valid (>= 0): round-down conversion == floor
  const int ix = __double2int_rd(xx), iy = __double2int_rd(xy);
  const float tx = (float)(xx - (double)ix), ty = (float)(xy - (double)iy);
  const int64_t ox0 = ix * sx, ox1 = (ix + 1 > nx - 1 ? ix : ix + 1) * sx;
  const int64_t oy0 = iy * sy, oy1 = (iy + 1 > ny - 1 ? iy : iy + 1) * sy;
  if (NDIM == 3) {
    const int iz = __double2int_rd(xz);
    const float tz = (float)(xz - (double)iz);
    const T* p00 = base + iz * sz + oy0;
    const T* p01 = base + iz * sz + oy1;
    const T* p10 = base + (iz + 1 > nz - 1 ? iz : iz + 1) * sz + oy0;
    const T* p11 = p10 - oy0 + oy1;
    const float v000 = (float)__ldg(p00 + ox0), v001 = (float)__ldg(p00 + ox1);
    const float v010 = (float)__ldg(p01 + ox0), v011 = (float)__ldg(p01 + ox1);
    const float v100 = (float)__ldg(p10 + ox0), v101 = (float)__ldg(p10 + ox1);
    const float v110 = (float)__ldg(p11 + ox0), v111 = (float)__ldg(p11 + ox1);
    const float a0 = lerp(lerp(v000, v001, tx), lerp(v010, v011, tx), ty);
    const float a1 = lerp(lerp(v100, v101, tx), lerp(v110, v111, tx), ty);
    return lerp(a0, a1, tz);
  }
  const T* p0 = base + oy0;
  const T* p1 = base + oy1;
  const float v00 = (float)__ldg(p0 + ox0), v01 = (float)__ldg(p0 + ox1);
  const float v10 = (float)__ldg(p1 + ox0), v11 = (float)__ldg(p1 + ox1);
  return lerp(lerp(v00, v01, tx), lerp(v10, v11, tx), ty);
}

// Evaluates view X at output sample index (cz, cy, cx) (already halo-shifted).
template <int NDIM, int ORDER, bool WANT_V, bool WANT_B>
__device__ __forceinline__ ViewEval eval_view(const mvs_view_xform& X,
                                              const float* __restrict__ tables, double cz,
                                              double cy, double cx) {
  ViewEval r;
  r.v = 0.0f;
  r.b = 0.0f;
  const double* m = X.matrix;
  double xz = 0.0;
  if (NDIM == 3) xz = affine_coord(m + 0, X.offset[0], affine_pre<NDIM>(m + 0, cz, cy), cx);
  double xy = affine_coord(m + 3, X.offset[1], affine_pre<NDIM>(m + 3, cz, cy), cx);
  double xx = affine_coord(m + 6, X.offset[2], affine_pre<NDIM>(m + 6, cz, cy), cx);
  const int nz = X.shape[0], ny = X.shape[1], nx = X.shape[2];
  bool valid = !(xx < 0.0 || xx > (double)(nx - 1) || xy < 0.0 || xy > (double)(ny - 1));
  if (NDIM == 3) valid = valid && !(xz < 0.0 || xz > (double)(nz - 1));
  r.valid = valid;
  if (!valid) return r;

  if (WANT_V) {
    // one (warp-uniform) dtype switch per view instead of one per tap
    const int dt = X.dtype;
    if (dt == MVS_U16) r.v = sample_view<NDIM, ORDER, unsigned short>(X, xz, xy, xx);
    else if (dt == MVS_F32) {
      r.v = sample_view<NDIM, ORDER, float>(X, xz, xy, xx);
      // NaN data inside a float view counts as "outside" for this voxel: the reference zeroes
      // the weight where the transformed view is NaN (fusion/_core.py:1648) and its fusion
      // functions are nan-aware, so another view's valid data survives
      if (r.v != r.v) { r.valid = false; return r; }
    } else r.v = sample_view<NDIM, ORDER, unsigned char>(X, xz, xy, xx);
  }
  if (WANT_B) {
    const double* w = X.wmatrix;
    double uz = 0.0;
    if (NDIM == 3)
      uz = affine_coord(w + 0, X.woffset[0], affine_pre<NDIM>(w + 0, cz, cy), cx);
    double uy = affine_coord(w + 3, X.woffset[1], affine_pre<NDIM>(w + 3, cz, cy), cx);
    double ux = affine_coord(w + 6, X.woffset[2], affine_pre<NDIM>(w + 6, cz, cy), cx);
    r.b = blend_weight<NDIM>(tables + (int64_t)X.table * 125, uz, uy, ux);
  }
  return r;
}

// Conservative test: can view X be valid anywhere in the sample-index box?
template <int NDIM>
__device__ bool view_touches_box(const mvs_view_xform& X, const double lo[3],
                                 const double hi[3]) {
  for (int d = (NDIM == 3 ? 0 : 1); d < 3; ++d) {
    double a = X.offset[d], b = X.offset[d];
    for (int j = (NDIM == 3 ? 0 : 1); j < 3; ++j) {
      double m = X.matrix[d * 3 + j];
      double p = m * lo[j], q = m * hi[j];
      a += fmin(p, q);
      b += fmax(p, q);
    }
    if (b < -1e-6 || a > (double)(X.shape[d] - 1) + 1e-6) return false;
  }
  return true;
}

}  // namespace mvs
#include "fuse_affine.cuh"
namespace mvs {

template <int NDIM, int ORDER, int MODE, bool PARTIAL>
__global__ void __launch_bounds__(kThreads)
fuse_kernel(const mvs_chunk* __restrict__ chunks, const int64_t* __restrict__ block_start,
            int n_chunks, const mvs_view_xform* __restrict__ xforms,
            const float* __restrict__ tables, int64_t block_begin, int64_t block_end) {
  __shared__ unsigned char s_flag[kMaxXforms];
  __shared__ int s_nact;
  __shared__ int s_single;

  const int64_t bid = block_begin + (int64_t)blockIdx.x + (int64_t)blockIdx.y * gridDim.x;
  if (bid >= block_end) return;
  int lo = 0, hi = n_chunks - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (__ldg(block_start + mid) <= bid) lo = mid; else hi = mid - 1;
  }
  const mvs_chunk& ck = chunks[lo];
  const int64_t local = bid - __ldg(block_start + lo);
  const int sh_y = ck.shape[1], sh_x = ck.shape[2];
  const int nbx = (sh_x + kBX - 1) / kBX, nby = (sh_y + kBY - 1) / kBY;
  const int x0 = (int)(local % nbx) * kBX;
  const int y0 = (int)((local / nbx) % nby) * kBY;
  const int z = (int)(local / ((int64_t)nbx * nby));
  const int first = ck.first_xform, nxf = ck.n_xforms;

  // ---- per-block culling of the chunk's view list (order preserved) ----
  if (threadIdx.x == 0) { s_nact = 0; s_single = -1; }
  __syncthreads();
  {
    double blo[3], bhi[3];
    blo[0] = (double)(z + ck.halo[0]);
    bhi[0] = blo[0];
    blo[1] = (double)(y0 + ck.halo[1]);
    bhi[1] = (double)(min(y0 + kBY, sh_y) - 1 + ck.halo[1]);
    blo[2] = (double)(x0 + ck.halo[2]);
    bhi[2] = (double)(min(x0 + kBX, sh_x) - 1 + ck.halo[2]);
    for (int i = threadIdx.x; i < nxf; i += kThreads) {
      bool t = view_touches_box<NDIM>(xforms[first + i], blo, bhi);
      s_flag[i] = t ? 1 : 0;
      if (t) { atomicAdd(&s_nact, 1); atomicMax(&s_single, i); }
    }
  }
  __syncthreads();
  const int nact = s_nact;
  const int single = s_single;  // the only active view when nact == 1

  const int lane = threadIdx.x & 31;
  const int y = y0 + (threadIdx.x >> 5);
  if (y >= sh_y) return;
  const double cz = (double)(z + ck.halo[0]);
  const double cy = (double)(y + ck.halo[1]);

  float res[kVPT];
  float den[kVPT];
#pragma unroll
  for (int k = 0; k < kVPT; ++k) { res[k] = 0.0f; den[k] = 0.0f; }

  if (MODE == MVS_FUSE_WAVG) {
    if (nact == 1 && !PARTIAL) {
      // single contributing view: w/w == 1 exactly where b > 0, else 0
      const mvs_view_xform& X = xforms[first + single];
#pragma unroll
      for (int k = 0; k < kVPT; ++k) {
        int x = x0 + lane + 32 * k;
        if (x < sh_x) {
          ViewEval e = eval_view<NDIM, ORDER, true, true>(X, tables, cz, cy,
                                                          (double)(x + ck.halo[2]));
          res[k] = (e.valid && e.b > 0.0f) ? __fmul_rn(e.v, __fdiv_rn(e.b, e.b)) : 0.0f;
        }
      }
    } else if (nact >= 1) {
      // One pass over the views: acc = sum_i v_i b_i, s = sum_i b_i (sequential
      // float32, view order), out = acc / s.  A voxel that exactly one view weights
      // keeps that view's value untouched (the reference's b/b == 1), so single-view
      // voxels stay bit-exact; blended voxels differ from the reference's
      // sum_i v_i (b_i / s) by float32 rounding only.
      float s[kVPT], vone[kVPT];
      int npos[kVPT];
#pragma unroll
      for (int k = 0; k < kVPT; ++k) { s[k] = 0.0f; vone[k] = 0.0f; npos[k] = 0; }
      for (int i = 0; i < nxf; ++i) {
        if (!s_flag[i]) continue;
        const mvs_view_xform& X = xforms[first + i];
#pragma unroll
        for (int k = 0; k < kVPT; ++k) {
          int x = x0 + lane + 32 * k;
          if (x < sh_x) {
            ViewEval e = eval_view<NDIM, ORDER, true, true>(X, tables, cz, cy,
                                                            (double)(x + ck.halo[2]));
            if (e.valid) {
              s[k] = __fadd_rn(s[k], e.b);
              res[k] = __fadd_rn(res[k], __fmul_rn(e.v, e.b));
              if (e.b > 0.0f) { vone[k] = e.v; ++npos[k]; }
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kVPT; ++k) {
        if (PARTIAL) den[k] = s[k];
        else res[k] = npos[k] == 0 ? 0.0f : (npos[k] == 1 ? vone[k] : __fdiv_rn(res[k], s[k]));
      }
    }
  } else {
    float cnt[kVPT];
    bool any[kVPT];
#pragma unroll
    for (int k = 0; k < kVPT; ++k) { cnt[k] = 0.0f; any[k] = false; }
    for (int i = 0; i < nxf; ++i) {
      if (!s_flag[i]) continue;
      const mvs_view_xform& X = xforms[first + i];
#pragma unroll
      for (int k = 0; k < kVPT; ++k) {
        int x = x0 + lane + 32 * k;
        if (x < sh_x) {
          ViewEval e = eval_view<NDIM, ORDER, true, false>(X, tables, cz, cy,
                                                           (double)(x + ck.halo[2]));
          if (e.valid) {
            if (MODE == MVS_FUSE_MAX) {
              res[k] = any[k] ? fmaxf(res[k], e.v) : e.v;
            } else {
              res[k] = __fadd_rn(res[k], e.v);
              cnt[k] = __fadd_rn(cnt[k], 1.0f);
            }
            any[k] = true;
          }
        }
      }
    }
    if (MODE == MVS_FUSE_MEAN) {
#pragma unroll
      for (int k = 0; k < kVPT; ++k) res[k] = any[k] ? __fdiv_rn(res[k], cnt[k]) : 0.0f;
    }
  }

  const int64_t row = (int64_t)z * ck.stride[0] + (int64_t)y * ck.stride[1];
#pragma unroll
  for (int k = 0; k < kVPT; ++k) {
    int x = x0 + lane + 32 * k;
    if (x < sh_x) {
      int64_t o = row + (int64_t)x * ck.stride[2];
      if (PARTIAL) {
        ck.acc_num[o] = res[k];
        ck.acc_den[o] = den[k];
      } else {
        store_from_float(ck.out, ck.out_dtype, o, res[k]);
      }
    }
  }
}

// transform_sim of every view of one chunk (NaN outside) plus the un-normalised
// blending weights (cosine ramp x validity): the (V, *chunk) stacks fuse_np hands
// to weights_func / fusion_func (fusion/_core.py:1621-1651).
template <int NDIM, int ORDER>
__global__ void __launch_bounds__(256)
resample_views_kernel(mvs_chunk ck, const mvs_view_xform* __restrict__ xforms, int n_views,
                      const float* __restrict__ tables, float* __restrict__ out_views,
                      float* __restrict__ out_weights) {
  const long long N = (long long)ck.shape[0] * ck.shape[1] * ck.shape[2];
  const int v = blockIdx.y;
  const mvs_view_xform& X = xforms[v];
  // one warp per 32-voxel row segment, 32-bit index arithmetic (no 64-bit divisions per voxel)
  const unsigned rows = (unsigned)ck.shape[0] * (unsigned)ck.shape[1], nx = (unsigned)ck.shape[2];
  const unsigned xt = (nx + 31) / 32, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (unsigned u = blockIdx.x * nwarps + (threadIdx.x >> 5); u < rows * xt; u += gridDim.x * nwarps) {
    const unsigned row = u / xt, xu = (u - row * xt) * 32 + lane;
    if (xu >= nx) continue;
    const unsigned zu = row / (unsigned)ck.shape[1];
    const int x = (int)xu, y = (int)(row - zu * (unsigned)ck.shape[1]), z = (int)zu;
    const long long i = (long long)row * nx + x;
    ViewEval e;
    if (out_weights)
      e = eval_view<NDIM, ORDER, true, true>(X, tables, (double)(z + ck.halo[0]),
                                             (double)(y + ck.halo[1]), (double)(x + ck.halo[2]));
    else
      e = eval_view<NDIM, ORDER, true, false>(X, tables, (double)(z + ck.halo[0]),
                                              (double)(y + ck.halo[1]), (double)(x + ck.halo[2]));
    out_views[(long long)v * N + i] = e.valid ? e.v : NAN;
    if (out_weights) out_weights[(long long)v * N + i] = e.valid ? e.b : 0.f;
  }
}

__global__ void finalize_kernel(const float* __restrict__ num, const float* __restrict__ den,
                                void* out, int out_dtype, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    float d = den[i];
    if (d == 0.0f) d = 1.0f;
    store_from_float(out, out_dtype, i, __fdiv_rn(num[i], d));
  }
}

// One block column per box: out[box voxel] = cast(num / den) from the box's packed
// (C-order) accumulators; 4 consecutive x voxels per thread.
__global__ void __launch_bounds__(256)
finalize_boxes_kernel(const mvs_chunk* __restrict__ boxes) {
  const mvs_chunk& b = boxes[blockIdx.y];
  const int nx = b.shape[2], ny = b.shape[1];
  const int64_t n = (int64_t)b.shape[0] * ny * nx;
  const float* __restrict__ num = b.acc_num;
  const float* __restrict__ den = b.acc_den;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % nx);
    const int64_t r = i / nx;
    const int y = (int)(r % ny), z = (int)(r / ny);
    float d = den[i];
    if (d == 0.0f) d = 1.0f;
    store_from_float(b.out, b.out_dtype, z * b.stride[0] + y * b.stride[1] + x * b.stride[2],
                     __fdiv_rn(num[i], d));
  }
}

}  // namespace mvs

struct mvs_fuse_plan {
  // general-affine schedule
  mvs_chunk* d_chunks = nullptr;
  int64_t* d_block_start = nullptr;
  int n_chunks = 0;
  int64_t total_blocks = 0;
  // translation (stencil) schedule
  mvs_chunk* d_chunks_st = nullptr;
  int64_t* d_block_start_st = nullptr;
  int n_chunks_st = 0;
  int64_t total_blocks_st = 0;
  mvs::StencilXform* d_sxf = nullptr;
  CUtensorMap* d_tmaps = nullptr;
  mvs::BlockRec* d_recs = nullptr;  // per-block schedule of the stencil path
  unsigned long long* d_counter = nullptr;  // dynamic block scheduler
  // original chunk order -> position in the two schedules
  std::vector<int> prefix_st;          // stencil chunks among chunks[0..c)
  std::vector<int64_t> h_bs_st, h_bs_gen;  // host copies of the block schedules
  int n_chunks_total = 0;
  int64_t run_st[2] = {0, 0}, run_gen[2] = {0, 0};  // block ranges of the current run
  int stencil_dtype = MVS_F32;
  // general affine with TMA-staged bricks (fuse_affine.cuh)
  mvs_chunk* d_chunks_aff = nullptr;
  int64_t* d_block_start_aff = nullptr;
  int n_chunks_aff = 0;
  int64_t total_blocks_aff = 0;
  mvs::AffInfo* d_ainfo = nullptr;
  CUtensorMap* d_tmaps_aff = nullptr;
  size_t aff_smem = 0;
  std::vector<int> prefix_aff;
  std::vector<int64_t> h_bs_aff;
  int64_t run_aff[2] = {0, 0};
  bool s3 = false;  // 3-D chunks run the z-marching column kernel (fuse_stencil3.cuh)
  int sm_count = 148;
  mvs_view_xform* d_xforms = nullptr;
  float* d_tables = nullptr;
  int n_xforms = 0, n_tables = 0;
  int ndim = 2, order = 1, mode = 0;
  bool partial = false;
  int64_t out_voxels = 0;
};

using namespace mvs;

// ---- translation fast path: host-side classification ---------------------------

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                             const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tensorMapEncodeTiled get_tensor_map_encoder() {
  static PFN_tensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_tensorMapEncodeTiled)p;
  }
  return fn;
}

// TMA descriptor of one view window: box = one stencil block footprint (bz == 0), or
// the z-marching kernel's boxes of bz planes (fuse_stencil3.cuh).
static bool make_tensor_map(const mvs_view_xform& X, int ndim, CUtensorMap* out, int bz = 0) {
  PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t es = (cuuint64_t)dtype_size(X.dtype);
  const CUtensorMapDataType dt = X.dtype == MVS_F32   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : X.dtype == MVS_U16 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                                      : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  const cuuint32_t bw = (cuuint32_t)((bz ? S3::BX : (ndim == 3 ? SBlock<3>::BX : SBlock<2>::BX)) + 16 / es);
  cuuint64_t gdim[3] = {(cuuint64_t)X.shape[2], (cuuint64_t)X.shape[1], (cuuint64_t)X.shape[0]};
  cuuint64_t gstr[2] = {(cuuint64_t)X.stride[1] * es, (cuuint64_t)X.stride[0] * es};
  cuuint32_t box[3] = {bw, (cuuint32_t)(ndim == 3 ? SBlock<3>::ROWS_Y : SBlock<2>::ROWS_Y),
                       (cuuint32_t)(ndim == 3 ? SBlock<3>::ROWS_Z : 1)};
  if (bz) { box[1] = S3::ROWS; box[2] = (cuuint32_t)bz; }
  cuuint32_t estr[3] = {1, 1, 1};
  // a rank-2 map needs a sane stride even for single-row windows
  if (X.shape[1] == 1) gstr[0] = ((gdim[0] * es + 15) / 16) * 16;
  if (ndim == 3 && X.shape[0] == 1) gstr[1] = gstr[0] * gdim[1];
  CUresult r = enc(out, dt, (cuuint32_t)ndim, const_cast<void*>(X.data), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// TMA descriptor of a view with an explicit box (z, y, x elements): affine bricks.
static bool make_tensor_map_box(const mvs_view_xform& X, int ndim, const int box_zyx[3], CUtensorMap* out) {
  PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t es = (cuuint64_t)dtype_size(X.dtype);
  const CUtensorMapDataType dt = X.dtype == MVS_F32   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : X.dtype == MVS_U16 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                                      : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  cuuint64_t gdim[3] = {(cuuint64_t)X.shape[2], (cuuint64_t)X.shape[1], (cuuint64_t)X.shape[0]};
  cuuint64_t gstr[2] = {(cuuint64_t)X.stride[1] * es, (cuuint64_t)X.stride[0] * es};
  cuuint32_t box[3] = {(cuuint32_t)box_zyx[2], (cuuint32_t)box_zyx[1], (cuuint32_t)box_zyx[0]};
  cuuint32_t estr[3] = {1, 1, 1};
  if (X.shape[1] == 1) gstr[0] = ((gdim[0] * es + 15) / 16) * 16;
  if (ndim == 3 && X.shape[0] == 1) gstr[1] = gstr[0] * gdim[1];
  CUresult r = enc(out, dt, (cuuint32_t)ndim, const_cast<void*>(X.data), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Brick extent of pairing X for the affine path's output block; false when the view cannot
// be staged (alignment, brick too large for shared memory).
static bool affine_box(const mvs_view_xform& X, int ndim, int box_zyx[3]) {
  const int64_t es = (int64_t)dtype_size(X.dtype);
  if (X.stride[2] != 1 || ((uintptr_t)X.data) % 16 || (X.stride[1] * es) % 16) return false;
  if (ndim == 3 && (X.stride[0] * es) % 16) return false;
  const int A = (int)(16 / es);
  const int bz = ndim == 3 ? ABlock<3>::BZ : 1, by = ndim == 3 ? ABlock<3>::BY : ABlock<2>::BY,
            bx = ndim == 3 ? ABlock<3>::BX : ABlock<2>::BX;
  int64_t bytes = es;
  for (int d = 0; d < 3; ++d) {
    if (ndim == 2 && d == 0) { box_zyx[0] = 1; continue; }
    double ext = fabs(X.matrix[3 * d + 1]) * (by - 1) + fabs(X.matrix[3 * d + 2]) * (bx - 1);
    if (ndim == 3) ext += fabs(X.matrix[3 * d + 0]) * (bz - 1);
    if (!(ext < 1000.0)) return false;
    int n = (int)ceil(ext) + 4;  // -1 margin below, second tap, rounding slack
    if (d == 2) {
      n = ((n + A - 1 + A - 1) / A) * A;  // + alignment slack, whole 16-byte units
      if (((n / A) & 1) == 0) n += A;     // odd number of 16-byte units per row ...
    } else if (d == 1 && (n & 1) == 0) {
      ++n;                                // ... and an odd row count: lanes that step along y or z
    }                                     // of the brick (rotated views) spread over 8 banks, not 4
    if (n > 256) return false;
    box_zyx[d] = n;
    bytes *= n;
  }
  return bytes <= kAffMaxBrickBytes;
}

// Fills S and returns true when pairing X is a pure translation that the
// bulk-copy stencil kernel can serve.
static bool make_stencil(const mvs_view_xform& X, int ndim, int order, int dtype,
                         StencilXform& S) {
  memset(&S, 0, sizeof(S));
  for (int d = 0; d < 3; ++d)
    for (int j = 0; j < 3; ++j) {
      if (X.matrix[d * 3 + j] != (d == j ? 1.0 : 0.0)) return false;
      if (d != j && X.wmatrix[d * 3 + j] != 0.0) return false;
    }
  if (X.dtype != dtype || X.stride[2] != 1) return false;
  const int64_t esize = (int64_t)dtype_size(dtype);
  if (((uintptr_t)X.data) % 16 || (X.stride[1] * esize) % 16) return false;
  if (ndim == 3 && (X.stride[0] * esize) % 16) return false;
  for (int d = 0; d < 3; ++d) {
    S.omin[d] = INT_MIN / 2; S.omax[d] = INT_MAX / 2;
    S.wm[d] = X.wmatrix[d * 3 + d]; S.woff[d] = X.woffset[d];
  }
  for (int d = 3 - ndim; d < 3; ++d) {
    const double off = X.offset[d];
    const int n = X.shape[d];
    if (!(fabs(off) < 1e9)) return false;
    if (order == 0) {
      S.shift[d] = (int)floor(off + 0.5); S.t[d] = 0.f; S.d1[d] = 0;
    } else {
      const double f = floor(off);
      S.shift[d] = (int)f; S.t[d] = (float)(off - f);
      if (S.t[d] >= 1.0f) return false;  // cannot happen after the 1e-6 snap
      S.d1[d] = S.t[d] != 0.f ? 1 : 0;
    }
    // valid sample range: scipy's predicate is fl(o + off) < 0 || fl(o + off) > n - 1
    long long lo = (long long)ceil(-off), hi = (long long)floor((double)(n - 1) - off);
    while ((double)lo + off < 0.0) ++lo;
    while ((double)(lo - 1) + off >= 0.0) --lo;
    while ((double)hi + off > (double)(n - 1)) --hi;
    while ((double)(hi + 1) + off <= (double)(n - 1)) ++hi;
    if (lo < INT_MIN / 2 || hi > INT_MAX / 2) return false;
    S.omin[d] = (int)lo; S.omax[d] = (int)hi;
    // order 0 picks floor(o + off + 0.5): must stay inside the window
  }
  return true;
}

template <int NDIM, typename T, int MODE, bool PARTIAL>
static cudaError_t launch_stencil(const mvs_fuse_plan* p, cudaStream_t st) {
  const int64_t nb = p->run_st[1] - p->run_st[0];
  auto kern = fuse_stencil_kernel<NDIM, T, MODE, PARTIAL>;
  const size_t smem = stencil_smem_bytes<NDIM, T>();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(p->d_counter, 0, sizeof(unsigned long long), st)) != cudaSuccess) return e;
  // persistent: 2-3 CTAs per SM, each walks blocks bid, bid + grid, ...
  const int grid = (int)std::min<int64_t>(nb, (int64_t)p->sm_count * (NDIM == 2 ? 4 : 2));
  kern<<<grid, SBlock<NDIM>::THREADS, smem, st>>>(p->d_chunks_st, p->d_block_start_st, p->n_chunks_st,
                                            p->d_xforms, p->d_sxf, p->d_tables, p->d_tmaps, p->d_recs, p->d_counter,
                                            p->run_st[0], p->run_st[1]);
  return cudaGetLastError();
}

template <typename T, int MODE, bool PARTIAL>
static cudaError_t launch_stencil3(const mvs_fuse_plan* p, cudaStream_t st) {
  const int64_t nb = p->run_st[1] - p->run_st[0];
  auto kern = fuse_stencil3_kernel<T, MODE, PARTIAL>;
  const size_t smem = stencil3_smem_bytes<T>();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(p->d_counter, 0, sizeof(unsigned long long), st)) != cudaSuccess) return e;
  const int grid = (int)std::min<int64_t>(nb, (int64_t)p->sm_count * 2);  // persistent, 2 CTAs per SM
  kern<<<grid, S3::THREADS, smem, st>>>(p->d_chunks_st, p->d_block_start_st, p->n_chunks_st, p->d_xforms,
                                        p->d_sxf, p->d_tables, p->d_tmaps, p->d_recs, p->d_counter,
                                        p->run_st[0], p->run_st[1]);
  return cudaGetLastError();
}

template <int NDIM, typename T>
static cudaError_t dispatch_stencil_mode(const mvs_fuse_plan* p, cudaStream_t st) {
  if (NDIM == 3 && p->s3) {
    switch (p->mode) {
      case MVS_FUSE_WAVG:
        return p->partial ? launch_stencil3<T, MVS_FUSE_WAVG, true>(p, st)
                          : launch_stencil3<T, MVS_FUSE_WAVG, false>(p, st);
      case MVS_FUSE_MAX:
        return launch_stencil3<T, MVS_FUSE_MAX, false>(p, st);
      default:
        return launch_stencil3<T, MVS_FUSE_MEAN, false>(p, st);
    }
  }
  switch (p->mode) {
    case MVS_FUSE_WAVG:
      return p->partial ? launch_stencil<NDIM, T, MVS_FUSE_WAVG, true>(p, st)
                        : launch_stencil<NDIM, T, MVS_FUSE_WAVG, false>(p, st);
    case MVS_FUSE_MAX:
      return launch_stencil<NDIM, T, MVS_FUSE_MAX, false>(p, st);
    default:
      return launch_stencil<NDIM, T, MVS_FUSE_MEAN, false>(p, st);
  }
}

template <int NDIM>
static cudaError_t dispatch_stencil(const mvs_fuse_plan* p, cudaStream_t st) {
  switch (p->stencil_dtype) {
    case MVS_U8: return dispatch_stencil_mode<NDIM, unsigned char>(p, st);
    case MVS_U16: return dispatch_stencil_mode<NDIM, unsigned short>(p, st);
    default: return dispatch_stencil_mode<NDIM, float>(p, st);
  }
}

template <int NDIM, int ORDER, int MODE, bool PARTIAL>
static cudaError_t launch_fuse(const mvs_fuse_plan* p, cudaStream_t st) {
  const int64_t nb = p->run_gen[1] - p->run_gen[0];
  const int64_t gx = std::min<int64_t>(nb, 1 << 30);
  const int64_t gy = (nb + gx - 1) / gx;
  dim3 grid((unsigned)gx, (unsigned)gy);
  fuse_kernel<NDIM, ORDER, MODE, PARTIAL><<<grid, kThreads, 0, st>>>(
      p->d_chunks, p->d_block_start, p->n_chunks, p->d_xforms, p->d_tables, p->run_gen[0],
      p->run_gen[1]);
  return cudaGetLastError();
}

template <int NDIM, int ORDER, int MODE, bool PARTIAL>
static cudaError_t launch_affine(const mvs_fuse_plan* p, cudaStream_t st) {
  const int64_t nb = p->run_aff[1] - p->run_aff[0];
  const int64_t gx = std::min<int64_t>(nb, 1 << 30);
  const int64_t gy = (nb + gx - 1) / gx;
  auto kern = fuse_affine_kernel<NDIM, ORDER, MODE, PARTIAL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * p->aff_smem));
  if (e != cudaSuccess) return e;
  kern<<<dim3((unsigned)gx, (unsigned)gy), ABlock<NDIM>::THREADS, 2 * p->aff_smem, st>>>(
      p->d_chunks_aff, p->d_block_start_aff, p->n_chunks_aff, p->d_xforms, p->d_ainfo, p->d_tables,
      p->d_tmaps_aff, p->run_aff[0], p->run_aff[1], (int)p->aff_smem);
  return cudaGetLastError();
}

template <int NDIM, int ORDER>
static cudaError_t dispatch_affine(const mvs_fuse_plan* p, cudaStream_t st) {
  switch (p->mode) {
    case MVS_FUSE_WAVG:
      return p->partial ? launch_affine<NDIM, ORDER, MVS_FUSE_WAVG, true>(p, st)
                        : launch_affine<NDIM, ORDER, MVS_FUSE_WAVG, false>(p, st);
    case MVS_FUSE_MAX:
      return launch_affine<NDIM, ORDER, MVS_FUSE_MAX, false>(p, st);
    default:
      return launch_affine<NDIM, ORDER, MVS_FUSE_MEAN, false>(p, st);
  }
}

template <int NDIM, int ORDER>
static cudaError_t dispatch_mode(const mvs_fuse_plan* p, cudaStream_t st) {
  switch (p->mode) {
    case MVS_FUSE_WAVG:
      return p->partial ? launch_fuse<NDIM, ORDER, MVS_FUSE_WAVG, true>(p, st)
                        : launch_fuse<NDIM, ORDER, MVS_FUSE_WAVG, false>(p, st);
    case MVS_FUSE_MAX:
      return launch_fuse<NDIM, ORDER, MVS_FUSE_MAX, false>(p, st);
    default:
      return launch_fuse<NDIM, ORDER, MVS_FUSE_MEAN, false>(p, st);
  }
}

extern "C" int mvs_fuse_plan_create(mvs_fuse_plan** plan, const mvs_chunk* chunks,
                                    int n_chunks, const mvs_view_xform* xforms, int n_xforms,
                                    const float* tables, int n_tables, int ndim, int order,
                                    int fusion_mode, void* stream) {
  MVS_REQUIRE(plan != nullptr, MVS_ERR_INVALID, "plan is NULL");
  *plan = nullptr;
  MVS_REQUIRE(ndim == 2 || ndim == 3, MVS_ERR_INVALID, "ndim must be 2 or 3, got %d", ndim);
  MVS_REQUIRE(order == 0 || order == 1, MVS_ERR_UNSUPPORTED,
              "interpolation order %d not supported (0 or 1)", order);
  MVS_REQUIRE(fusion_mode >= MVS_FUSE_WAVG && fusion_mode <= MVS_FUSE_MEAN, MVS_ERR_INVALID,
              "unknown fusion mode %d", fusion_mode);
  MVS_REQUIRE(n_chunks >= 0 && n_xforms >= 0 && n_tables >= 0, MVS_ERR_INVALID,
              "negative count");
  MVS_REQUIRE(n_chunks == 0 || chunks != nullptr, MVS_ERR_INVALID, "chunks is NULL");
  MVS_REQUIRE(n_xforms == 0 || xforms != nullptr, MVS_ERR_INVALID, "xforms is NULL");

  int64_t out_voxels = 0;
  bool partial = false, any_out = false;
  for (int c = 0; c < n_chunks; ++c) {
    const mvs_chunk& ck = chunks[c];
    for (int d = 0; d < 3; ++d)
      MVS_REQUIRE(ck.shape[d] >= 0 && ck.halo[d] >= 0, MVS_ERR_INVALID,
                  "chunk %d: negative extent/halo", c);
    MVS_REQUIRE(ndim == 3 || ck.shape[0] <= 1, MVS_ERR_INVALID,
                "chunk %d: 2-D plan with z extent %d", c, ck.shape[0]);
    MVS_REQUIRE(ck.n_xforms >= 0 && ck.n_xforms <= kMaxXforms, MVS_ERR_UNSUPPORTED,
                "chunk %d: %d views (max %d per chunk)", c, ck.n_xforms, kMaxXforms);
    MVS_REQUIRE(ck.first_xform >= 0 && ck.first_xform + ck.n_xforms <= n_xforms,
                MVS_ERR_INVALID, "chunk %d: xform range out of bounds", c);
    MVS_REQUIRE(ck.out_dtype >= MVS_U8 && ck.out_dtype <= MVS_F32, MVS_ERR_INVALID,
                "chunk %d: bad out dtype", c);
    const bool has_acc = ck.acc_num != nullptr || ck.acc_den != nullptr;
    if (has_acc) {
      MVS_REQUIRE(ck.acc_num && ck.acc_den, MVS_ERR_INVALID,
                  "chunk %d: acc_num and acc_den must both be set", c);
      MVS_REQUIRE(fusion_mode == MVS_FUSE_WAVG, MVS_ERR_UNSUPPORTED,
                  "partial accumulators need MVS_FUSE_WAVG");
      partial = true;
    } else {
      MVS_REQUIRE(ck.out != nullptr || ck.shape[0] * ck.shape[1] * ck.shape[2] == 0,
                  MVS_ERR_INVALID, "chunk %d: out is NULL", c);
      any_out = true;
    }
    out_voxels += (int64_t)ck.shape[0] * ck.shape[1] * ck.shape[2];
  }
  MVS_REQUIRE(!(partial && any_out), MVS_ERR_UNSUPPORTED,
              "a plan must be all-partial or all-final");
  for (int i = 0; i < n_xforms; ++i) {
    const mvs_view_xform& X = xforms[i];
    MVS_REQUIRE(X.data != nullptr, MVS_ERR_INVALID, "xform %d: data is NULL", i);
    MVS_REQUIRE(X.dtype >= MVS_U8 && X.dtype <= MVS_F32, MVS_ERR_INVALID,
                "xform %d: bad dtype %d", i, X.dtype);
    MVS_REQUIRE(X.shape[0] >= 1 && X.shape[1] >= 1 && X.shape[2] >= 1, MVS_ERR_INVALID,
                "xform %d: empty window", i);
    MVS_REQUIRE(fusion_mode != MVS_FUSE_WAVG || (X.table >= 0 && X.table < n_tables),
                MVS_ERR_INVALID, "xform %d: table index %d out of range", i, X.table);
  }
  MVS_REQUIRE(n_tables == 0 || tables != nullptr, MVS_ERR_INVALID, "tables is NULL");

  // classify pairings / chunks: translation stencil path vs general affine path
  const bool allow_stencil = getenv("MVS_FUSE_GENERIC") == nullptr;
  // 3-D: the z-marching column kernel (fuse_stencil3.cuh) is an experiment that measured
  // SLOWER than the block kernel on the B200 (C3: 12.0 vs 7.2 ms, profiles/r02_stencil3_*):
  // opt-in with MVS_STENCIL3=1
  const bool s3 = ndim == 3 && getenv("MVS_STENCIL3") != nullptr;
  const int stencil_dtype = n_xforms ? xforms[0].dtype : MVS_F32;
  std::vector<StencilXform> sxf(n_xforms);
  std::vector<char> xf_ok(n_xforms, 0);
  for (int i = 0; i < n_xforms; ++i)
    xf_ok[i] = allow_stencil && make_stencil(xforms[i], ndim, order, stencil_dtype, sxf[i]);
  std::vector<CUtensorMap> tmaps;
  {
    std::vector<int> owner;  // xform index that created tmaps[k]
    for (int i = 0; i < n_xforms; ++i) {
      if (!xf_ok[i]) continue;
      int found = -1;
      for (size_t k = 0; k < owner.size() && found < 0; ++k) {
        const mvs_view_xform& Y = xforms[owner[k]];
        const mvs_view_xform& X = xforms[i];
        if (Y.data == X.data && !memcmp(Y.shape, X.shape, sizeof(X.shape)) &&
            !memcmp(Y.stride, X.stride, sizeof(X.stride)) && Y.dtype == X.dtype)
          found = (int)k;
      }
      if (found < 0) {
        CUtensorMap m;
        if (s3) {
          // two boxes per view: the step's 4 new planes and the single plane below them
          CUtensorMap m1;
          if (!make_tensor_map(xforms[i], ndim, &m, S3::PZ) || !make_tensor_map(xforms[i], ndim, &m1, 1)) {
            xf_ok[i] = 0;
            continue;
          }
          tmaps.push_back(m);
          tmaps.push_back(m1);
        } else {
          if (!make_tensor_map(xforms[i], ndim, &m)) { xf_ok[i] = 0; continue; }
          tmaps.push_back(m);
        }
        owner.push_back(i);
        found = (int)owner.size() - 1;
      }
      sxf[i].tmap = found;
    }
  }
  // affine pairings: brick extents, tensor maps (one per distinct view + box), float32 matrices
  const bool allow_affine = getenv("MVS_FUSE_GATHER") == nullptr;
  std::vector<AffInfo> ainfo(n_xforms);
  std::vector<char> af_ok(n_xforms, 0);
  std::vector<CUtensorMap> tmaps_aff;
  {
    struct Key { const void* data; int32_t shape[3]; int64_t stride[3]; int dtype; int box[3]; };
    std::vector<Key> keys;
    for (int i = 0; i < n_xforms; ++i) {
      AffInfo& a = ainfo[i];
      memset(&a, 0, sizeof(a));
      a.tmap = -1;
      for (int q = 0; q < 9; ++q) { a.m[q] = (float)xforms[i].matrix[q]; a.wm[q] = (float)xforms[i].wmatrix[q]; }
      if (!allow_affine || xf_ok[i]) continue;  // translations take the stencil path
      int box[3];
      if (!affine_box(xforms[i], ndim, box)) continue;
      const mvs_view_xform& X = xforms[i];
      int found = -1;
      for (size_t k = 0; k < keys.size() && found < 0; ++k)
        if (keys[k].data == X.data && !memcmp(keys[k].shape, X.shape, sizeof(X.shape)) &&
            !memcmp(keys[k].stride, X.stride, sizeof(X.stride)) && keys[k].dtype == X.dtype &&
            !memcmp(keys[k].box, box, sizeof(box)))
          found = (int)k;
      if (found < 0) {
        CUtensorMap m;
        if (!make_tensor_map_box(X, ndim, box, &m)) continue;
        Key k;
        k.data = X.data; memcpy(k.shape, X.shape, sizeof(X.shape)); memcpy(k.stride, X.stride, sizeof(X.stride));
        k.dtype = X.dtype; memcpy(k.box, box, sizeof(box));
        keys.push_back(k);
        tmaps_aff.push_back(m);
        found = (int)keys.size() - 1;
      }
      a.tmap = found;
      memcpy(a.box, box, sizeof(box));
      af_ok[i] = 1;
    }
  }
  std::vector<mvs_chunk> ch_st, ch_gen, ch_aff;
  std::vector<int64_t> bs_st(1, 0), bs_gen(1, 0), bs_aff(1, 0);
  std::vector<int> prefix_st(n_chunks + 1, 0), prefix_aff(n_chunks + 1, 0);
  size_t aff_smem = 0;
  for (int c = 0; c < n_chunks; ++c) {
    const mvs_chunk& ck = chunks[c];
    bool ok = allow_stencil && ck.n_xforms <= 32 && ck.stride[2] == 1;
    for (int i = 0; ok && i < ck.n_xforms; ++i) ok = xf_ok[ck.first_xform + i];
    // affine-staged: every pairing stageable (translations of a mixed chunk included: they
    // were skipped above, so a chunk mixing both kinds falls through to the gather kernel)
    bool aok = !ok && allow_affine && ck.n_xforms > 0;
    for (int i = 0; aok && i < ck.n_xforms; ++i) aok = af_ok[ck.first_xform + i];
    prefix_st[c + 1] = prefix_st[c] + (ok ? 1 : 0);
    prefix_aff[c + 1] = prefix_aff[c] + (aok ? 1 : 0);
    if (ok) {
      const int BX = s3 ? S3::BX : (ndim == 3 ? SBlock<3>::BX : SBlock<2>::BX);
      const int BY = s3 ? S3::BY : (ndim == 3 ? SBlock<3>::BY : SBlock<2>::BY);
      const int BZ = s3 ? std::max(ck.shape[0], 1) : (ndim == 3 ? SBlock<3>::BZ : 1);  // s3: one unit per column
      const int64_t nb = (int64_t)((ck.shape[2] + BX - 1) / BX) * ((ck.shape[1] + BY - 1) / BY) *
                         ((ck.shape[0] + BZ - 1) / BZ);
      ch_st.push_back(ck);
      bs_st.push_back(bs_st.back() + nb);
    } else if (aok) {
      const int BX = ndim == 3 ? ABlock<3>::BX : ABlock<2>::BX, BY = ndim == 3 ? ABlock<3>::BY : ABlock<2>::BY,
                BZ = ndim == 3 ? ABlock<3>::BZ : 1;
      const int64_t nb = (int64_t)((ck.shape[2] + BX - 1) / BX) * ((ck.shape[1] + BY - 1) / BY) *
                         ((ck.shape[0] + BZ - 1) / BZ);
      ch_aff.push_back(ck);
      bs_aff.push_back(bs_aff.back() + nb);
      for (int i = 0; i < ck.n_xforms; ++i) {
        const AffInfo& a = ainfo[ck.first_xform + i];
        aff_smem = std::max(aff_smem, (size_t)a.box[0] * a.box[1] * a.box[2] * dtype_size(xforms[ck.first_xform + i].dtype));
      }
    } else {
      const int64_t nbx = (ck.shape[2] + kBX - 1) / kBX, nby = (ck.shape[1] + kBY - 1) / kBY;
      ch_gen.push_back(ck);
      bs_gen.push_back(bs_gen.back() + nbx * nby * (int64_t)ck.shape[0]);
    }
  }
  aff_smem = ((aff_smem + 127) / 128) * 128;

  cudaStream_t st = (cudaStream_t)stream;
  mvs_fuse_plan* p = new mvs_fuse_plan();
  p->n_chunks = (int)ch_gen.size(); p->n_chunks_st = (int)ch_st.size();
  p->n_xforms = n_xforms; p->n_tables = n_tables;
  p->ndim = ndim; p->order = order; p->mode = fusion_mode; p->partial = partial;
  p->total_blocks = bs_gen.back(); p->total_blocks_st = bs_st.back();
  p->stencil_dtype = stencil_dtype;
  p->s3 = s3;
  p->prefix_st = prefix_st; p->h_bs_st = bs_st; p->h_bs_gen = bs_gen; p->n_chunks_total = n_chunks;
  p->prefix_aff = prefix_aff; p->h_bs_aff = bs_aff; p->n_chunks_aff = (int)ch_aff.size();
  p->total_blocks_aff = bs_aff.back(); p->aff_smem = aff_smem;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (p->sm_count <= 0) p->sm_count = 148;
  }
  p->out_voxels = out_voxels;
  auto fail = [&](cudaError_t e, const char* what) {
    set_error("%s failed: %s", what, cudaGetErrorString(e));
    mvs_fuse_plan_destroy(p);
    return (int)MVS_ERR_CUDA;
  };
  cudaError_t e;
  auto upload = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    if (bytes == 0) return cudaSuccess;
    cudaError_t err = cudaMalloc(dst, bytes);
    if (err != cudaSuccess) return err;
    return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, st);
  };
  if ((e = upload((void**)&p->d_chunks, ch_gen.data(), sizeof(mvs_chunk) * ch_gen.size())) !=
      cudaSuccess)
    return fail(e, "upload chunks");
  if ((e = upload((void**)&p->d_chunks_st, ch_st.data(), sizeof(mvs_chunk) * ch_st.size())) !=
      cudaSuccess)
    return fail(e, "upload stencil chunks");
  if ((e = upload((void**)&p->d_sxf, sxf.data(), sizeof(StencilXform) * sxf.size())) !=
      cudaSuccess)
    return fail(e, "upload stencil constants");
  if ((e = upload((void**)&p->d_block_start_st, bs_st.data(), sizeof(int64_t) * bs_st.size())) !=
      cudaSuccess)
    return fail(e, "upload stencil block schedule");
  if ((e = upload((void**)&p->d_tmaps, tmaps.data(), sizeof(CUtensorMap) * tmaps.size())) !=
      cudaSuccess)
    return fail(e, "upload tensor maps");
  if ((e = upload((void**)&p->d_xforms, xforms, sizeof(mvs_view_xform) * n_xforms)) !=
      cudaSuccess)
    return fail(e, "upload xforms");
  if ((e = upload((void**)&p->d_tables, tables, sizeof(float) * 125 * n_tables)) != cudaSuccess)
    return fail(e, "upload tables");
  if ((e = upload((void**)&p->d_block_start, bs_gen.data(), sizeof(int64_t) * bs_gen.size())) !=
      cudaSuccess)
    return fail(e, "upload block schedule");
  if (!ch_aff.empty()) {
    if ((e = upload((void**)&p->d_chunks_aff, ch_aff.data(), sizeof(mvs_chunk) * ch_aff.size())) != cudaSuccess)
      return fail(e, "upload affine chunks");
    if ((e = upload((void**)&p->d_block_start_aff, bs_aff.data(), sizeof(int64_t) * bs_aff.size())) != cudaSuccess)
      return fail(e, "upload affine block schedule");
    if ((e = upload((void**)&p->d_ainfo, ainfo.data(), sizeof(AffInfo) * ainfo.size())) != cudaSuccess)
      return fail(e, "upload affine pairings");
    if ((e = upload((void**)&p->d_tmaps_aff, tmaps_aff.data(), sizeof(CUtensorMap) * tmaps_aff.size())) != cudaSuccess)
      return fail(e, "upload affine tensor maps");
  }
  // per-block schedule of the stencil path (views touching each block + weight classes)
  if (p->total_blocks_st > 0) {
    if ((e = cudaMalloc((void**)&p->d_recs, sizeof(BlockRec) * p->total_blocks_st)) != cudaSuccess)
      return fail(e, "allocate block schedule");
    if ((e = cudaMalloc((void**)&p->d_counter, sizeof(unsigned long long))) != cudaSuccess)
      return fail(e, "allocate block counter");
    const float* tabs = fusion_mode == MVS_FUSE_WAVG ? p->d_tables : nullptr;
    const unsigned grid = (unsigned)((p->total_blocks_st + 7) / 8);
    if (s3)
      stencil3_classify_kernel<<<grid, 256, 0, st>>>(p->d_chunks_st, p->d_block_start_st, p->n_chunks_st,
                                                     p->d_xforms, p->d_sxf, tabs, p->d_recs);
    else if (ndim == 3)
      stencil_classify_kernel<3><<<grid, 256, 0, st>>>(p->d_chunks_st, p->d_block_start_st, p->n_chunks_st,
                                                     p->d_xforms, p->d_sxf, tabs, p->d_recs);
    else
      stencil_classify_kernel<2><<<grid, 256, 0, st>>>(p->d_chunks_st, p->d_block_start_st, p->n_chunks_st,
                                                     p->d_xforms, p->d_sxf, tabs, p->d_recs);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "classify launch");
  }
  // host staging buffers die with this call: make the copies complete
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e, "cudaStreamSynchronize");
  *plan = p;
  return MVS_OK;
}

extern "C" int mvs_fuse_plan_run_chunks(mvs_fuse_plan* p, int first_chunk, int n_chunks,
                                        void* stream) {
  MVS_REQUIRE(p != nullptr, MVS_ERR_INVALID, "plan is NULL");
  MVS_REQUIRE(first_chunk >= 0 && n_chunks >= 0 && first_chunk + n_chunks <= p->n_chunks_total,
              MVS_ERR_INVALID, "chunk range [%d, %d) outside the plan's %d chunks", first_chunk,
              first_chunk + n_chunks, p->n_chunks_total);
  const int c0 = first_chunk, c1 = first_chunk + n_chunks;
  const int s0 = p->prefix_st[c0], s1 = p->prefix_st[c1];
  const int a0 = p->prefix_aff[c0], a1 = p->prefix_aff[c1];
  const int g0 = c0 - s0 - a0, g1 = c1 - s1 - a1;
  p->run_aff[0] = p->h_bs_aff[a0]; p->run_aff[1] = p->h_bs_aff[a1];
  p->run_st[0] = p->h_bs_st[s0]; p->run_st[1] = p->h_bs_st[s1];
  p->run_gen[0] = p->h_bs_gen[g0]; p->run_gen[1] = p->h_bs_gen[g1];
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  if (p->run_st[1] > p->run_st[0])
    e = p->ndim == 2 ? dispatch_stencil<2>(p, st) : dispatch_stencil<3>(p, st);
  if (e == cudaSuccess && p->run_aff[1] > p->run_aff[0]) {
    if (p->ndim == 2)
      e = p->order == 0 ? dispatch_affine<2, 0>(p, st) : dispatch_affine<2, 1>(p, st);
    else
      e = p->order == 0 ? dispatch_affine<3, 0>(p, st) : dispatch_affine<3, 1>(p, st);
  }
  if (e == cudaSuccess && p->run_gen[1] > p->run_gen[0]) {
    if (p->ndim == 2)
      e = p->order == 0 ? dispatch_mode<2, 0>(p, st) : dispatch_mode<2, 1>(p, st);
    else
      e = p->order == 0 ? dispatch_mode<3, 0>(p, st) : dispatch_mode<3, 1>(p, st);
  }
  if (e != cudaSuccess) {
    set_error("fuse kernel launch failed: %s", cudaGetErrorString(e));
    return MVS_ERR_CUDA;
  }
  return MVS_OK;
}

extern "C" int mvs_fuse_plan_run(mvs_fuse_plan* p, void* stream) {
  MVS_REQUIRE(p != nullptr, MVS_ERR_INVALID, "plan is NULL");
  return mvs_fuse_plan_run_chunks(p, 0, p->n_chunks_total, stream);
}

extern "C" int mvs_fuse_plan_info(const mvs_fuse_plan* p, int* launches, int64_t* blocks,
                                  int64_t* out_voxels) {
  MVS_REQUIRE(p != nullptr, MVS_ERR_INVALID, "plan is NULL");
  if (launches) *launches = (p->total_blocks > 0 ? 1 : 0) + (p->total_blocks_st > 0 ? 1 : 0) + (p->total_blocks_aff > 0 ? 1 : 0);
  if (blocks) *blocks = p->total_blocks + p->total_blocks_st + p->total_blocks_aff;
  if (out_voxels) *out_voxels = p->out_voxels;
  return MVS_OK;
}

extern "C" int mvs_fuse_plan_destroy(mvs_fuse_plan* p) {
  if (!p) return MVS_OK;
  cudaFree(p->d_chunks);
  cudaFree(p->d_chunks_st);
  cudaFree(p->d_block_start_st);
  cudaFree(p->d_sxf);
  cudaFree(p->d_tmaps);
  cudaFree(p->d_recs);
  cudaFree(p->d_counter);
  cudaFree(p->d_xforms);
  cudaFree(p->d_tables);
  cudaFree(p->d_block_start);
  cudaFree(p->d_chunks_aff);
  cudaFree(p->d_block_start_aff);
  cudaFree(p->d_ainfo);
  cudaFree(p->d_tmaps_aff);
  delete p;
  return MVS_OK;
}

extern "C" int mvs_fuse_finalize(const float* acc_num, const float* acc_den, void* out,
                                 int out_dtype, int64_t n, void* stream) {
  MVS_REQUIRE(n >= 0, MVS_ERR_INVALID, "negative n");
  if (n == 0) return MVS_OK;
  MVS_REQUIRE(acc_num && acc_den && out, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(out_dtype >= MVS_U8 && out_dtype <= MVS_F32, MVS_ERR_INVALID, "bad out dtype");
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  finalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(acc_num, acc_den, out, out_dtype, n);
  MVS_CHECK_CUDA(cudaGetLastError());
  return MVS_OK;
}

extern "C" int mvs_fuse_finalize_boxes(const mvs_chunk* boxes, int n_boxes, void* stream) {
  MVS_REQUIRE(n_boxes >= 0, MVS_ERR_INVALID, "negative n_boxes");
  if (n_boxes == 0) return MVS_OK;
  MVS_REQUIRE(boxes != nullptr, MVS_ERR_INVALID, "boxes is NULL");
  MVS_REQUIRE(n_boxes <= 65535, MVS_ERR_UNSUPPORTED, "%d boxes (max 65535 per call)", n_boxes);
  int64_t biggest = 0;
  for (int i = 0; i < n_boxes; ++i) {
    const mvs_chunk& b = boxes[i];
    MVS_REQUIRE(b.out && b.acc_num && b.acc_den, MVS_ERR_INVALID, "box %d: NULL pointer", i);
    MVS_REQUIRE(b.out_dtype >= MVS_U8 && b.out_dtype <= MVS_F32, MVS_ERR_INVALID,
                "box %d: bad out dtype", i);
    MVS_REQUIRE(b.shape[0] >= 0 && b.shape[1] >= 0 && b.shape[2] >= 0, MVS_ERR_INVALID,
                "box %d: negative extent", i);
    biggest = std::max<int64_t>(biggest, (int64_t)b.shape[0] * b.shape[1] * b.shape[2]);
  }
  if (biggest == 0) return MVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  mvs_chunk* d_boxes = nullptr;
  MVS_CHECK_CUDA(mvs::pool_malloc((void**)&d_boxes, sizeof(mvs_chunk) * n_boxes, st));
  // pageable source: staged before the call returns
  MVS_CHECK_CUDA(cudaMemcpyAsync(d_boxes, boxes, sizeof(mvs_chunk) * n_boxes,
                                 cudaMemcpyHostToDevice, st));
  const unsigned gx = (unsigned)std::min<int64_t>((biggest + 255) / 256, 148 * 8);
  finalize_boxes_kernel<<<dim3(gx, (unsigned)n_boxes), 256, 0, st>>>(d_boxes);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(d_boxes, st);
  if (e != cudaSuccess) { set_error("finalize launch: %s", cudaGetErrorString(e)); return MVS_ERR_CUDA; }
  return MVS_OK;
}

extern "C" int mvs_resample_views(const mvs_view_xform* xforms, int n_views, const float* tables,
                                  int n_tables, const int32_t shape[3], const int32_t halo[3],
                                  int ndim, int order, float* d_views, float* d_weights,
                                  void* stream) {
  MVS_REQUIRE(xforms && shape && halo && d_views, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(n_views >= 1, MVS_ERR_INVALID, "n_views = %d", n_views);
  MVS_REQUIRE(ndim == 2 || ndim == 3, MVS_ERR_INVALID, "ndim must be 2 or 3");
  MVS_REQUIRE(order == 0 || order == 1, MVS_ERR_UNSUPPORTED, "interpolation order %d", order);
  MVS_REQUIRE(!d_weights || (tables && n_tables >= 1), MVS_ERR_INVALID, "weights need tables");
  for (int i = 0; i < n_views; ++i) {
    MVS_REQUIRE(xforms[i].data != nullptr, MVS_ERR_INVALID, "xform %d: data is NULL", i);
    MVS_REQUIRE(!d_weights || (xforms[i].table >= 0 && xforms[i].table < n_tables),
                MVS_ERR_INVALID, "xform %d: table index out of range", i);
  }
  mvs_chunk ck{};
  for (int d = 0; d < 3; ++d) { ck.shape[d] = shape[d]; ck.halo[d] = halo[d]; }
  const long long N = (long long)shape[0] * shape[1] * shape[2];
  if (N <= 0) return MVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // Parameter buffer from the stream-ordered pool of the CURRENT device: nothing is shared
  // between devices, streams or threads (per-group streams of register_pairs run
  // concurrently).  The host arrays are pageable, so the async copies have staged them when
  // they return.
  const size_t xbytes = ((sizeof(mvs_view_xform) * n_views + 255) / 256) * 256;
  const size_t tbytes = d_weights ? sizeof(float) * 125 * n_tables : 0;
  void* d_buf = nullptr;
  MVS_CHECK_CUDA(mvs::pool_malloc(&d_buf, xbytes + tbytes, st));
  mvs_view_xform* d_x = (mvs_view_xform*)d_buf;
  float* d_t = d_weights ? (float*)((char*)d_buf + xbytes) : nullptr;
  if (d_weights)
    MVS_CHECK_CUDA(cudaMemcpyAsync(d_t, tables, tbytes, cudaMemcpyHostToDevice, st));
  MVS_CHECK_CUDA(cudaMemcpyAsync(d_x, xforms, sizeof(mvs_view_xform) * n_views, cudaMemcpyHostToDevice, st));
  dim3 grid((unsigned)std::min<long long>((N + 255) / 256, 148 * 16), n_views);
  if (ndim == 2) {
    if (order == 0) resample_views_kernel<2, 0><<<grid, 256, 0, st>>>(ck, d_x, n_views, d_t, d_views, d_weights);
    else resample_views_kernel<2, 1><<<grid, 256, 0, st>>>(ck, d_x, n_views, d_t, d_views, d_weights);
  } else {
    if (order == 0) resample_views_kernel<3, 0><<<grid, 256, 0, st>>>(ck, d_x, n_views, d_t, d_views, d_weights);
    else resample_views_kernel<3, 1><<<grid, 256, 0, st>>>(ck, d_x, n_views, d_t, d_views, d_weights);
  }
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(d_buf, st);
  if (e != cudaSuccess) { set_error("resample launch: %s", cudaGetErrorString(e)); return MVS_ERR_CUDA; }
  return MVS_OK;
}
