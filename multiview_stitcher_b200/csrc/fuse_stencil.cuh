// Translation fast path of the fused resample-blend kernel (sm_100a).
//
// For the tile-stitching case the pixel matrix handed to scipy is exactly the
// identity (transformation.py:56 with equal spacings), so the order-1 resample
// of a view is a constant-coefficient 2^ndim-tap stencil on a shifted window:
//   x_in = o + off,  floor(x_in) = o + floor(off),  frac = off - floor(off).
// A CTA owns an output block (BZ x BY x 128).  For every contributing view it
//   1. pulls the block's input footprint (rows of <= 160 elements) into shared
//      memory with 1-D bulk async copies (cp.async.bulk -> UBLKCP, 16-byte
//      aligned, completion on an mbarrier) -- HBM is read in full coalesced
//      row segments, never gathered;
//   2. meanwhile evaluates the separable parts of the blending weight (table
//      cell + fraction per column / row / plane, float64) into shared memory;
//   3. interpolates from shared memory (lanes along x: conflict-free), reusing
//      the x-interpolated rows between neighbouring output rows / planes;
//   4. blends with the reference's float32 operation order.
// Results are bit-identical to the general kernel except for the last ulp of
// the interpolation fraction.
#pragma once

#include "common.cuh"

namespace mvs {

struct StencilXform {
  int shift[3];   // window px of tap 0 = sample index + shift
  float t[3];     // interpolation fraction per axis
  int d1[3];      // offset of the second tap (0 when the fraction is 0 / order 0)
  int omin[3];    // valid sample-index range per axis (inclusive)
  int omax[3];
  int always_pos; // blending weight > 0 on every valid voxel (single-view shortcut)
  double wm[3];   // diagonal of wmatrix
  double woff[3];
};

template <int NDIM>
struct SBlock {
  static constexpr int BX = 128;
  static constexpr int BY = NDIM == 3 ? 8 : 32;
  static constexpr int BZ = NDIM == 3 ? 4 : 1;
  static constexpr int ROWP = 160;  // staged row pitch in elements (>= BX + 1 + 2*15)
  static constexpr int ROWS_Y = BY + 1;
  static constexpr int ROWS_Z = NDIM == 3 ? BZ + 1 : 1;
  static constexpr int NROWS = ROWS_Y * ROWS_Z;
  static constexpr int OUTS = 16;  // outputs per thread
  static constexpr int NW = BX + BY + BZ;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  unsigned spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) __trap();  // never hang the GPU on a lost copy
  }
}

__device__ __forceinline__ float lerp_s(float a, float b, float t) {
  return fmaf(t, b, fmaf(-t, a, a));
}

__device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// Separable blending-weight coordinates of this block for one view.
template <int NDIM>
__device__ __forceinline__ void fill_weight_axes(const StencilXform& S, int x0s, int y0s, int z0s,
                                                 int* s_wi, float* s_wt) {
  using B = SBlock<NDIM>;
  for (int i = threadIdx.x; i < B::NW; i += blockDim.x) {
    int d, o;
    if (i < B::BX) { d = 2; o = x0s + i; }
    else if (i < B::BX + B::BY) { d = 1; o = y0s + (i - B::BX); }
    else { d = 0; o = z0s + (i - B::BX - B::BY); }
    const double u = __dadd_rn(__dmul_rn((double)o, S.wm[d]), S.woff[d]);
    int cell = -1;
    float fr = 0.f;
    if (!(u < 0.0 || u > 4.0)) {
      const double f = floor(u);
      cell = (int)f;
      fr = (float)(u - f);
    }
    s_wi[i] = cell;
    s_wt[i] = fr;
  }
}

template <int NDIM>
__device__ __forceinline__ float stencil_weight(const float* __restrict__ tab, const int* s_wi,
                                                const float* s_wt, int ix_, int iy_, int iz_) {
  using B = SBlock<NDIM>;
  const int ix = s_wi[ix_], iy = s_wi[B::BX + iy_];
  if (ix < 0 || iy < 0) return 0.f;
  const float tx = s_wt[ix_], ty = s_wt[B::BX + iy_];
  const int ix1 = min(ix + 1, 4), iy1 = min(iy + 1, 4);
  float w;
  if (NDIM == 3) {
    const int iz = s_wi[B::BX + B::BY + iz_];
    if (iz < 0) return 0.f;
    const float tz = s_wt[B::BX + B::BY + iz_];
    const int iz1 = min(iz + 1, 4);
    const float* p0 = tab + iz * 25;
    const float* p1 = tab + iz1 * 25;
    float a0 = lerp_s(lerp_s(p0[iy * 5 + ix], p0[iy * 5 + ix1], tx),
                      lerp_s(p0[iy1 * 5 + ix], p0[iy1 * 5 + ix1], tx), ty);
    float a1 = lerp_s(lerp_s(p1[iy * 5 + ix], p1[iy * 5 + ix1], tx),
                      lerp_s(p1[iy1 * 5 + ix], p1[iy1 * 5 + ix1], tx), ty);
    w = lerp_s(a0, a1, tz);
  } else {
    w = lerp_s(lerp_s(tab[iy * 5 + ix], tab[iy * 5 + ix1], tx),
               lerp_s(tab[iy1 * 5 + ix], tab[iy1 * 5 + ix1], tx), ty);
  }
  if (w < 1.0f) {
    float a = __fmul_rn(__fsub_rn(1.0f, w), 3.14159274101257324f);
    w = __fdiv_rn(__fadd_rn(cosf(a), 1.0f), 2.0f);
  }
  return fminf(fmaxf(w, 0.0f), 1.0f);
}

template <int NDIM, typename T, int MODE, bool PARTIAL>
__global__ void __launch_bounds__(256)
fuse_stencil_kernel(const mvs_chunk* __restrict__ chunks, const int64_t* __restrict__ block_start,
                    int n_chunks, const mvs_view_xform* __restrict__ xforms,
                    const StencilXform* __restrict__ sxf, const float* __restrict__ tables) {
  using B = SBlock<NDIM>;
  constexpr int A = 16 / (int)sizeof(T);  // elements per 16 bytes
  __shared__ __align__(128) T stage[B::NROWS * B::ROWP];
  __shared__ float s_tab[125];
  __shared__ int s_wi[B::NW];
  __shared__ float s_wt[B::NW];
  __shared__ unsigned char s_flag[kMaxXforms];
  __shared__ int s_nact, s_single;
  __shared__ __align__(8) unsigned long long s_bar;

  const int64_t bid = (int64_t)blockIdx.x + (int64_t)blockIdx.y * gridDim.x;
  if (bid >= block_start[n_chunks]) return;
  int lo = 0, hi = n_chunks - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (__ldg(block_start + mid) <= bid) lo = mid; else hi = mid - 1;
  }
  const mvs_chunk& ck = chunks[lo];
  const int64_t local = bid - __ldg(block_start + lo);
  const int sh_z = ck.shape[0], sh_y = ck.shape[1], sh_x = ck.shape[2];
  const int nbx = (sh_x + B::BX - 1) / B::BX, nby = (sh_y + B::BY - 1) / B::BY;
  const int x0 = (int)(local % nbx) * B::BX;
  const int y0 = (int)((local / nbx) % nby) * B::BY;
  const int z0 = (int)(local / ((int64_t)nbx * nby)) * B::BZ;
  const int first = ck.first_xform, nxf = ck.n_xforms;
  // block origin / extent in sample-index space
  const int x0s = x0 + ck.halo[2], y0s = y0 + ck.halo[1], z0s = z0 + ck.halo[0];
  const int x1s = min(x0 + B::BX, sh_x) - 1 + ck.halo[2];
  const int y1s = min(y0 + B::BY, sh_y) - 1 + ck.halo[1];
  const int z1s = min(z0 + B::BZ, sh_z) - 1 + ck.halo[0];

  if (threadIdx.x == 0) {
    s_nact = 0; s_single = -1;
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nxf; i += blockDim.x) {
    const StencilXform& S = sxf[first + i];
    bool t = S.omax[2] >= x0s && S.omin[2] <= x1s && S.omax[1] >= y0s && S.omin[1] <= y1s;
    if (NDIM == 3) t = t && S.omax[0] >= z0s && S.omin[0] <= z1s;
    s_flag[i] = t ? 1 : 0;
    if (t) { atomicAdd(&s_nact, 1); atomicMax(&s_single, i); }
  }
  __syncthreads();
  const int nact = s_nact;
  const int single = s_single;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = warp & 3, half = warp >> 2;
  const int jx = cg * 32 + lane;  // block-local output column
  uint32_t phase = 0;

  float acc[B::OUTS], s[B::OUTS];
  unsigned anymask = 0;  // MAX / MEAN: bit k set once a valid view was seen
#pragma unroll
  for (int k = 0; k < B::OUTS; ++k) { acc[k] = 0.f; s[k] = 0.f; }

  // block-local (jz, jy) of output k of this thread
  auto out_y = [&](int k) { return NDIM == 3 ? (k & 7) : half * 16 + k; };
  auto out_z = [&](int k) { return NDIM == 3 ? half * 2 + (k >> 3) : 0; };

  // valid bits of this thread's outputs for view S
  auto valid_bits = [&](const StencilXform& S) -> unsigned {
    unsigned m = 0;
    const int sx = x0s + jx;
    if (sx < S.omin[2] || sx > S.omax[2] || x0 + jx >= sh_x) return 0u;
#pragma unroll
    for (int k = 0; k < B::OUTS; ++k) {
      const int sy = y0s + out_y(k), sz = z0s + out_z(k);
      bool v = sy >= S.omin[1] && sy <= S.omax[1] && y0 + out_y(k) < sh_y;
      if (NDIM == 3) v = v && sz >= S.omin[0] && sz <= S.omax[0] && z0 + out_z(k) < sh_z;
      m |= v ? (1u << k) : 0u;
    }
    return m;
  };

  // stages the footprint of view `xi` (and its weight axes / table when WANT_W)
  auto stage_view = [&](int xi, bool want_w, bool want_data) {
    const mvs_view_xform& X = xforms[first + xi];
    const StencilXform& S = sxf[first + xi];
    __syncthreads();  // previous consumers of stage / weight axes are done
    if (want_w) {
      fill_weight_axes<NDIM>(S, x0s, y0s, z0s, s_wi, s_wt);
      const float* tab = tables + (int64_t)X.table * 125;
      for (int i = threadIdx.x; i < (NDIM == 3 ? 125 : 25); i += blockDim.x) s_tab[i] = __ldg(tab + i);
    }
    uint32_t total = 0;
    if (want_data) {
      const int nx = X.shape[2], ny = X.shape[1], nz = X.shape[0];
      const int x0g = x0s + S.shift[2], y0g = y0s + S.shift[1], z0g = z0s + S.shift[0];
      const int xa_u = floor_div(x0g, A) * A;
      const int xb_u = floor_div(x0g + B::BX + 1 + A - 1, A) * A;
      const int xa = max(xa_u, 0), xb = min(xb_u, nx);
      const int row_bytes = xb > xa ? (xb - xa) * (int)sizeof(T) : 0;
      const int nyv = max(0, min(y0g + B::ROWS_Y - 1, ny - 1) - max(y0g, 0) + 1);
      const int nzv = NDIM == 3 ? max(0, min(z0g + B::ROWS_Z - 1, nz - 1) - max(z0g, 0) + 1) : 1;
      total = (uint32_t)row_bytes * nyv * nzv;
      if (total) {
        if (threadIdx.x == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(&s_bar, total);
        }
        if (threadIdx.x < B::NROWS) {
          const int r = threadIdx.x;
          const int rz = r / B::ROWS_Y, ry = r - rz * B::ROWS_Y;
          const int gy = y0g + ry, gz = NDIM == 3 ? z0g + rz : 0;
          if (gy >= 0 && gy < ny && gz >= 0 && gz < nz) {
            const T* src = reinterpret_cast<const T*>(X.data) + (int64_t)gz * X.stride[0] +
                           (int64_t)gy * X.stride[1] + xa;
            bulk_g2s(stage + r * B::ROWP + (xa - xa_u), src, (uint32_t)row_bytes, &s_bar);
          }
        }
      }
    }
    __syncthreads();  // weight axes / table visible
    if (total) { mbar_wait(&s_bar, phase); phase ^= 1; }
  };

  // interpolated values of this thread's outputs from the staged footprint
  auto values = [&](const StencilXform& S, float* val) {
    const int x0g = x0s + S.shift[2];
    const int xa_u = floor_div(x0g, A) * A;
    const int c0 = (x0g - xa_u) + jx;
    const int c1 = c0 + S.d1[2];
    const float tx = S.t[2], ty = S.t[1], tz = S.t[0];
    const bool dy = S.d1[1] != 0, dz = S.d1[0] != 0;
    if (NDIM == 2) {
      const int rb = half * 16;
      const T* p = stage + rb * B::ROWP;
      float hprev = lerp_s((float)p[c0], (float)p[c1], tx);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        p += B::ROWP;
        const float hn = lerp_s((float)p[c0], (float)p[c1], tx);
        val[k] = dy ? lerp_s(hprev, hn, ty) : hprev;
        hprev = hn;
      }
    } else {
      float gprev[8];
#pragma unroll
      for (int pz = 0; pz < 3; ++pz) {
        const T* p = stage + ((half * 2 + pz) * B::ROWS_Y) * B::ROWP;
        float g[8];
        float hprev = lerp_s((float)p[c0], (float)p[c1], tx);
#pragma unroll
        for (int y = 0; y < 8; ++y) {
          p += B::ROWP;
          const float hn = lerp_s((float)p[c0], (float)p[c1], tx);
          g[y] = dy ? lerp_s(hprev, hn, ty) : hprev;
          hprev = hn;
        }
        if (pz > 0) {
#pragma unroll
          for (int y = 0; y < 8; ++y)
            val[(pz - 1) * 8 + y] = dz ? lerp_s(gprev[y], g[y], tz) : gprev[y];
        }
#pragma unroll
        for (int y = 0; y < 8; ++y) gprev[y] = g[y];
      }
    }
  };

  auto weight_of = [&](int k) {
    return stencil_weight<NDIM>(s_tab, s_wi, s_wt, jx, out_y(k), out_z(k));
  };

  if (MODE == MVS_FUSE_WAVG) {
    const bool fast_single = (nact == 1) && !PARTIAL && sxf[first + max(single, 0)].always_pos;
    if (nact >= 1 && fast_single) {
      const StencilXform& S = sxf[first + single];
      stage_view(single, false, true);
      float val[B::OUTS];
      values(S, val);
      const unsigned vm = valid_bits(S);
#pragma unroll
      for (int k = 0; k < B::OUTS; ++k) acc[k] = (vm >> k) & 1 ? val[k] : 0.f;
    } else if (nact >= 1) {
      // pass A: s = sum_i b_i * valid_i (float32, view order)
      for (int i = 0; i < nxf; ++i) {
        if (!s_flag[i]) continue;
        const StencilXform& S = sxf[first + i];
        stage_view(i, true, false);
        const unsigned vm = valid_bits(S);
#pragma unroll
        for (int k = 0; k < B::OUTS; ++k)
          if ((vm >> k) & 1) s[k] = __fadd_rn(s[k], weight_of(k));
      }
      if (!PARTIAL) {
#pragma unroll
        for (int k = 0; k < B::OUTS; ++k) if (s[k] == 0.f) s[k] = 1.f;
      }
      // pass B: sum_i v_i * (b_i / s)
      for (int i = 0; i < nxf; ++i) {
        if (!s_flag[i]) continue;
        const StencilXform& S = sxf[first + i];
        stage_view(i, true, true);
        float val[B::OUTS];
        values(S, val);
        const unsigned vm = valid_bits(S);
#pragma unroll
        for (int k = 0; k < B::OUTS; ++k) {
          if ((vm >> k) & 1) {
            const float b = weight_of(k);
            const float w = PARTIAL ? b : __fdiv_rn(b, s[k]);
            acc[k] = __fadd_rn(acc[k], __fmul_rn(val[k], w));
          }
        }
      }
    }
  } else {
    for (int i = 0; i < nxf; ++i) {
      if (!s_flag[i]) continue;
      const StencilXform& S = sxf[first + i];
      stage_view(i, false, true);
      float val[B::OUTS];
      values(S, val);
      const unsigned vm = valid_bits(S);
#pragma unroll
      for (int k = 0; k < B::OUTS; ++k) {
        if ((vm >> k) & 1) {
          if (MODE == MVS_FUSE_MAX) {
            acc[k] = (anymask >> k) & 1 ? fmaxf(acc[k], val[k]) : val[k];
          } else {
            acc[k] = __fadd_rn(acc[k], val[k]);
            s[k] = __fadd_rn(s[k], 1.0f);
          }
          anymask |= 1u << k;
        }
      }
    }
    if (MODE == MVS_FUSE_MEAN) {
#pragma unroll
      for (int k = 0; k < B::OUTS; ++k)
        acc[k] = (anymask >> k) & 1 ? __fdiv_rn(acc[k], s[k]) : 0.f;
    }
  }

  const int xo = x0 + jx;
  if (xo < sh_x) {
#pragma unroll
    for (int k = 0; k < B::OUTS; ++k) {
      const int yo = y0 + out_y(k), zo = z0 + out_z(k);
      if (yo < sh_y && zo < sh_z) {
        const int64_t o = (int64_t)zo * ck.stride[0] + (int64_t)yo * ck.stride[1] +
                          (int64_t)xo * ck.stride[2];
        if (PARTIAL) {
          ck.acc_num[o] = acc[k];
          ck.acc_den[o] = s[k];
        } else {
          store_from_float(ck.out, ck.out_dtype, o, acc[k]);
        }
      }
    }
  }
}

}  // namespace mvs
