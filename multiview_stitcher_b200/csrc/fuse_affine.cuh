// General-affine path of the fused resample-blend kernel with TMA-staged source bricks
// (sm_100a).  Replaces, for views whose affine is NOT a pure translation (multi-view
// light-sheet: rotations, tilt, anisotropic scale), the per-voxel global-memory gather of
// fuse_kernel (fuse.cu), which ran at ~2 % of the HBM roofline.
//
// An output block is compact (3-D: 8 x 8 x 32 voxels, 2-D: 16 x 64) so that its pre-image
// under any affine is a small brick of the view.  Per (block, view):
//   * one thread computes the brick origin from the block's corner coordinates (float64,
//     scipy's operation order) and issues ONE TMA tensor copy of the brick into shared
//     memory (cp.async.bulk.tensor -> UTMALDG; out-of-view elements are zero-filled).  Two
//     brick buffers: the copy of the NEXT view is in flight while the current one is sampled;
//   * eight threads evaluate the block's corners: all inside the view -> no per-voxel
//     validity tests; raw blending weight >= 1 at all of them -> every weight in the block
//     is exactly 1 and the table lookup is skipped (the views' interiors: most blocks);
//   * every thread walks its 8 (4) voxels: brick-local sample coordinates by float32 FMAs
//     from the exact corner coordinate, taps from shared memory (no 64-bit address
//     arithmetic, no cache misses), float32 interpolation; blending weight from the view's
//     5^ndim table in shared memory with the polynomial cosine ramp.
//   * exactness where it decides something: a voxel whose fast coordinate lies within 1e-3
//     px of the view's border (the outside predicate x < 0 || x > n-1 of
//     scipy.ndimage.affine_transform, transformation.py:136-139) or whose table coordinate
//     lies within 1e-3 of the table's border (where the reference's weight underflows to an
//     exact 0) is re-evaluated in float64 with scipy's operation order and the reference's
//     float32 cosine formula; order-0 picks are always computed in float64.  So validity,
//     nearest-neighbour picks and single-view voxels stay bit-identical to the reference.
#pragma once

#include "fuse_stencil.cuh"

namespace mvs {

template <int NDIM>
struct ABlock {
  static constexpr int BX = NDIM == 3 ? 32 : 64;
  static constexpr int BY = NDIM == 3 ? 8 : 16;
  static constexpr int BZ = NDIM == 3 ? 8 : 1;
  static constexpr int THREADS = 256;
  static constexpr int VPT = BX * BY * BZ / THREADS;  // 8 (3-D) / 4 (2-D)
};

constexpr int kAffMaxBrickBytes = 40 * 1024;

// per (chunk, view) pairing, host-built
struct AffInfo {
  int tmap;       // tensor map of the view with this pairing's box
  int box[3];     // staged brick extent (z, y, x) in elements
  float m[9];     // float32 copy of the pixel matrix (brick-local coordinates)
  float wm[9];    // ... of the table matrix
};

// exact sample coordinate, scipy's operation order: ((cz*m0 + cy*m1) + cx*m2) + off
__device__ __forceinline__ double aff_exact(const double* __restrict__ m, double off, double cz,
                                            double cy, double cx, bool three_d) {
  double pre = three_d ? __dadd_rn(__dmul_rn(cz, m[0]), __dmul_rn(cy, m[1])) : __dmul_rn(cy, m[1]);
  return __dadd_rn(__dadd_rn(pre, __dmul_rn(cx, m[2])), off);
}

template <typename T>
__device__ __forceinline__ float brick_tap(const T* __restrict__ b, int i) { return tofl(b[i]); }

// reference blending weight at exact table coordinates (weights.py:475-509), table in smem
template <int NDIM>
__device__ float aff_weight_exact(const float* __restrict__ tab, double uz, double uy, double ux) {
  float w = raw_table_value_s<NDIM>(tab, uz, uy, ux);
  if (w < 1.0f) {
    const float a = __fmul_rn(__fsub_rn(1.0f, w), 3.14159274101257324f);
    w = __fdiv_rn(__fadd_rn(cosf(a), 1.0f), 2.0f);
  }
  return fminf(fmaxf(w, 0.0f), 1.0f);
}

// per staged view, in shared memory (double-buffered)
struct AffMeta {
  float tab[128];
  float m[18];       // float32 pixel matrix, table matrix
  float base[3];     // corner sample coordinate minus brick origin
  float wbase[3];    // corner table coordinate
  float lim[6];      // validity bounds in brick-local coordinates: lo z,y,x, hi z,y,x
  float cmin[8];     // raw weight at the block's corners
  int cin[8];        // corner at least 1e-3 px inside the view
  int org[3];        // brick origin in view pixels (z, y, x)
  int box[3];        // brick extent
  int dtype, xi;     // view dtype, pairing index
};

#ifndef MVS_AFF_MINB
#define MVS_AFF_MINB 3
#endif
template <int NDIM, int ORDER, int MODE, bool PARTIAL>
__global__ void __launch_bounds__(256, MVS_AFF_MINB)
fuse_affine_kernel(const mvs_chunk* __restrict__ chunks, const int64_t* __restrict__ block_start,
                   int n_chunks, const mvs_view_xform* __restrict__ xforms,
                   const AffInfo* __restrict__ ainfo, const float* __restrict__ tables,
                   const CUtensorMap* __restrict__ tmaps, int64_t block_begin, int64_t block_end,
                   int brick_bytes) {
  using B = ABlock<NDIM>;
  constexpr int VPT = B::VPT;
  extern __shared__ __align__(128) unsigned char brick_raw[];  // two bricks of brick_bytes
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ AffMeta s_meta[2];
  __shared__ unsigned short s_list[kMaxXforms];
  __shared__ unsigned char s_flag[kMaxXforms];
  __shared__ int s_nact;

  const int64_t bid = block_begin + (int64_t)blockIdx.x + (int64_t)blockIdx.y * gridDim.x;
  if (bid >= block_end) return;
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
  const int x0s = x0 + ck.halo[2], y0s = y0 + ck.halo[1], z0s = z0 + ck.halo[0];
  const int x1s = min(x0 + B::BX, sh_x) - 1 + ck.halo[2], y1s = min(y0 + B::BY, sh_y) - 1 + ck.halo[1],
            z1s = min(z0 + B::BZ, sh_z) - 1 + ck.halo[0];

  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    double blo[3] = {(double)z0s, (double)y0s, (double)x0s}, bhi[3] = {(double)z1s, (double)y1s, (double)x1s};
    for (int i = threadIdx.x; i < nxf; i += B::THREADS)
      s_flag[i] = view_touches_box<NDIM>(xforms[first + i], blo, bhi) ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 0; i < nxf; ++i)
      if (s_flag[i]) s_list[n++] = (unsigned short)i;
    s_nact = n;
  }
  __syncthreads();
  const int nact = s_nact;

  // stages view j of the active list into buffer b (roles by thread range)
  auto stage = [&](int j, int b) {
    const int xi = first + (int)s_list[j];
    const mvs_view_xform& X = xforms[xi];
    const AffInfo& AI = ainfo[xi];
    AffMeta& M = s_meta[b];
    if (threadIdx.x == 0) {
      int org[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (NDIM == 2 && d == 0) { org[0] = 0; M.base[0] = 0.f; M.wbase[0] = 0.f; continue; }
        const double c = aff_exact(X.matrix + 3 * d, X.offset[d], (double)z0s, (double)y0s, (double)x0s, NDIM == 3);
        double mn = c;
        if (NDIM == 3) mn += fmin(0.0, X.matrix[3 * d + 0] * (B::BZ - 1));
        mn += fmin(0.0, X.matrix[3 * d + 1] * (B::BY - 1));
        mn += fmin(0.0, X.matrix[3 * d + 2] * (B::BX - 1));
        mn = fmax(fmin(mn, 2.0e9), -2.0e9);
        int o = (int)floor(mn) - 1;
        if (d == 2) {
          const int A = 16 / (int)dtype_size(X.dtype);
          o = floor_div(o, A) * A;  // TMA: 16-byte aligned innermost start
        }
        org[d] = o;
        M.base[d] = (float)(c - (double)o);
        M.wbase[d] = (float)aff_exact(X.wmatrix + 3 * d, X.woffset[d], (double)z0s, (double)y0s, (double)x0s, NDIM == 3);
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        M.org[d] = org[d];
        M.lim[d] = (float)(-org[d]);
        M.lim[3 + d] = (float)(X.shape[d] - 1 - org[d]);
        M.box[d] = AI.box[d];
      }
      M.dtype = X.dtype;
      M.xi = xi;
      const uint32_t bytes = (uint32_t)(AI.box[0] * AI.box[1] * AI.box[2]) * (uint32_t)dtype_size(X.dtype);
      mbar_expect_tx(&bar[b], bytes);
      if (NDIM == 3) tma_load_3d(brick_raw + (size_t)b * brick_bytes, tmaps + AI.tmap, org[2], org[1], org[0], &bar[b]);
      else tma_load_2d(brick_raw + (size_t)b * brick_bytes, tmaps + AI.tmap, org[2], org[1], &bar[b]);
    }
    if (MODE == MVS_FUSE_WAVG && threadIdx.x >= 32 && threadIdx.x < 32 + 125)
      M.tab[threadIdx.x - 32] = __ldg(tables + (int64_t)X.table * 125 + (threadIdx.x - 32));
    if (threadIdx.x >= 192 && threadIdx.x < 192 + 18)
      M.m[threadIdx.x - 192] = threadIdx.x < 201 ? AI.m[threadIdx.x - 192] : AI.wm[threadIdx.x - 201];
    if (threadIdx.x >= 224 && threadIdx.x < 232) {
      // the block's 8 corners, exactly: inside the view? raw weight?
      const int c = threadIdx.x - 224;
      const double cz = (double)((c & 4) ? z1s : z0s), cy = (double)((c & 2) ? y1s : y0s), cx = (double)((c & 1) ? x1s : x0s);
      const double ex = aff_exact(X.matrix + 6, X.offset[2], cz, cy, cx, NDIM == 3);
      const double ey = aff_exact(X.matrix + 3, X.offset[1], cz, cy, cx, NDIM == 3);
      const double ez = NDIM == 3 ? aff_exact(X.matrix + 0, X.offset[0], cz, cy, cx, true) : 0.0;
      bool in = ex >= 1e-3 && ex <= (double)(X.shape[2] - 1) - 1e-3 && ey >= 1e-3 && ey <= (double)(X.shape[1] - 1) - 1e-3;
      if (NDIM == 3) in = in && ez >= 1e-3 && ez <= (double)(X.shape[0] - 1) - 1e-3;
      M.cin[c] = in ? 1 : 0;
      float raw = 0.f;
      if (MODE == MVS_FUSE_WAVG) {
        const double ux = aff_exact(X.wmatrix + 6, X.woffset[2], cz, cy, cx, NDIM == 3);
        const double uy = aff_exact(X.wmatrix + 3, X.woffset[1], cz, cy, cx, NDIM == 3);
        const double uz = NDIM == 3 ? aff_exact(X.wmatrix + 0, X.woffset[0], cz, cy, cx, true) : 0.0;
        raw = raw_table_value<NDIM>(tables + (int64_t)X.table * 125, uz, uy, ux);
      }
      M.cmin[c] = raw;
    }
  };

  // thread -> voxels: 3-D: x = lane, y = warp, z = k;  2-D: x = tid & 63, y = (tid >> 6) + 4k
  const int dx = NDIM == 3 ? (threadIdx.x & 31) : (threadIdx.x & 63);
  const int dy0 = NDIM == 3 ? (threadIdx.x >> 5) : (threadIdx.x >> 6);
  const bool in_x = x0 + dx < sh_x;

  float res[VPT], sw[VPT], vone[VPT];
  int npos[VPT];
#pragma unroll
  for (int k = 0; k < VPT; ++k) { res[k] = 0.f; sw[k] = 0.f; vone[k] = 0.f; npos[k] = 0; }

  if (nact > 0) stage(0, 0);
  for (int j = 0; j < nact; ++j) {
    const int b = j & 1;
    __syncthreads();  // view j's record is complete; view j-1 has been consumed (its buffer is free)
    if (j + 1 < nact) stage(j + 1, b ^ 1);
    mbar_wait(&bar[b], (unsigned)(j >> 1) & 1u);
    const AffMeta& M = s_meta[b];
    const mvs_view_xform& X = xforms[M.xi];
    const unsigned char* brick = brick_raw + (size_t)b * brick_bytes;

    const int bxn = M.box[2], byn = M.box[1], bzn = M.box[0];
    const int oz = M.org[0], oy = M.org[1], ox = M.org[2];
    const float loz = M.lim[0], loy = M.lim[1], lox = M.lim[2], hiz = M.lim[3], hiy = M.lim[4], hix = M.lim[5];
    const bool all_in = (M.cin[0] & M.cin[1] & M.cin[2] & M.cin[3] & M.cin[4] & M.cin[5] & M.cin[6] & M.cin[7]) != 0;
    bool unit_w = false;
    if (MODE == MVS_FUSE_WAVG) {
      const float cm = fminf(fminf(fminf(M.cmin[0], M.cmin[1]), fminf(M.cmin[2], M.cmin[3])),
                             fminf(fminf(M.cmin[4], M.cmin[5]), fminf(M.cmin[6], M.cmin[7])));
      unit_w = cm >= 1.0f;
    }
    const float* m = M.m;
    const float* wm = M.m + 9;
    const float fx = (float)dx;
    const int dt = M.dtype;

#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int dz = NDIM == 3 ? k : 0;
      const int dy = NDIM == 3 ? dy0 : dy0 + 4 * k;
      if (!in_x || y0 + dy >= sh_y || z0 + dz >= sh_z) continue;
      const float fz = (float)dz, fy = (float)dy;
      float lz = 0.f;
      if (NDIM == 3) lz = fmaf(fx, m[2], fmaf(fy, m[1], fmaf(fz, m[0], M.base[0])));
      float ly = fmaf(fx, m[5], fmaf(fy, m[4], (NDIM == 3 ? fmaf(fz, m[3], M.base[1]) : M.base[1])));
      float lx = fmaf(fx, m[8], fmaf(fy, m[7], (NDIM == 3 ? fmaf(fz, m[6], M.base[2]) : M.base[2])));
      constexpr float EPS = 1e-3f;
      bool need_exact = false;
      if (!all_in) {
        bool out = lx < lox - EPS || lx > hix + EPS || ly < loy - EPS || ly > hiy + EPS;
        bool sure = lx > lox + EPS && lx < hix - EPS && ly > loy + EPS && ly < hiy - EPS;
        if (NDIM == 3) {
          out = out || lz < loz - EPS || lz > hiz + EPS;
          sure = sure && lz > loz + EPS && lz < hiz - EPS;
        }
        if (out) continue;
        need_exact = !sure;
      }
      if (ORDER == 0 && !need_exact) {
        // nearest neighbour from the fast coordinate unless it sits on a rounding boundary
        const float px = lx + 0.5f, py = ly + 0.5f, pz = lz + 0.5f;
        const float qx = px - floorf(px), qy = py - floorf(py), qz = pz - floorf(pz);
        need_exact = qx < EPS || qx > 1.f - EPS || qy < EPS || qy > 1.f - EPS;
        if (NDIM == 3) need_exact = need_exact || qz < EPS || qz > 1.f - EPS;
        lx = floorf(px); ly = floorf(py); lz = floorf(pz);
      }
#ifdef MVS_AFF_EXACT_X
      need_exact = true;
#endif
      const double cz = (double)(z0s + dz), cy = (double)(y0s + dy), cx = (double)(x0s + dx);
      if (need_exact) {
        // exact predicate (and, for order 0, exact picks) in float64
        const double ex = aff_exact(X.matrix + 6, X.offset[2], cz, cy, cx, NDIM == 3);
        const double ey = aff_exact(X.matrix + 3, X.offset[1], cz, cy, cx, NDIM == 3);
        const double ez = NDIM == 3 ? aff_exact(X.matrix + 0, X.offset[0], cz, cy, cx, true) : 0.0;
        bool valid = !(ex < 0.0 || ex > (double)(X.shape[2] - 1) || ey < 0.0 || ey > (double)(X.shape[1] - 1));
        if (NDIM == 3) valid = valid && !(ez < 0.0 || ez > (double)(X.shape[0] - 1));
        if (!valid) continue;
        if (ORDER == 0) {
          lx = (float)((long long)floor(__dadd_rn(ex, 0.5)) - ox);
          ly = (float)((long long)floor(__dadd_rn(ey, 0.5)) - oy);
          if (NDIM == 3) lz = (float)((long long)floor(__dadd_rn(ez, 0.5)) - oz);
        } else {
          lx = (float)(ex - (double)ox);
          ly = (float)(ey - (double)oy);
          if (NDIM == 3) lz = (float)(ez - (double)oz);
        }
      }
      // ---- sample the brick ----
      float v;
      {
        const float flx = floorf(lx), fly = floorf(ly), flz = floorf(lz);
        constexpr int LAST = ORDER == 0 ? 1 : 2;  // order 1 also reads the element after
        const int ix = min(max((int)flx, 0), bxn - LAST), iy = min(max((int)fly, 0), byn - LAST);
        const int iz = NDIM == 3 ? min(max((int)flz, 0), bzn - LAST) : 0;
        const float tx = lx - (float)ix, ty = ly - (float)iy, tz = lz - (float)iz;
        const int base = (iz * byn + iy) * bxn + ix;
        const int sy_ = bxn, sz_ = byn * bxn;
        auto fetch = [&](auto* bp) -> float {
          if (ORDER == 0) return brick_tap(bp, base);
          const float a00 = lerp_s(brick_tap(bp, base), brick_tap(bp, base + 1), tx);
          const float a01 = lerp_s(brick_tap(bp, base + sy_), brick_tap(bp, base + sy_ + 1), tx);
          const float a0 = lerp_s(a00, a01, ty);
          if (NDIM == 2) return a0;
          const float a10 = lerp_s(brick_tap(bp, base + sz_), brick_tap(bp, base + sz_ + 1), tx);
          const float a11 = lerp_s(brick_tap(bp, base + sz_ + sy_), brick_tap(bp, base + sz_ + sy_ + 1), tx);
          return lerp_s(a0, lerp_s(a10, a11, ty), tz);
        };
        if (dt == MVS_U16) v = fetch(reinterpret_cast<const unsigned short*>(brick));
        else if (dt == MVS_F32) {
          v = fetch(reinterpret_cast<const float*>(brick));
          if (v != v) continue;  // NaN data = outside for this voxel (fusion/_core.py:1648)
        } else v = fetch(reinterpret_cast<const unsigned char*>(brick));
      }
      if (MODE == MVS_FUSE_MAX) {
        res[k] = npos[k] ? fmaxf(res[k], v) : v;
        npos[k] = 1;
        continue;
      }
      if (MODE == MVS_FUSE_MEAN) {
        res[k] = __fadd_rn(res[k], v);
        sw[k] = __fadd_rn(sw[k], 1.0f);
        npos[k] = 1;
        continue;
      }
      // ---- blending weight ----
      float bw;
      if (unit_w) {
        bw = 1.0f;
      } else {
        const float* tab = M.tab;
        float uz = 0.f;
        if (NDIM == 3) uz = fmaf(fx, wm[2], fmaf(fy, wm[1], fmaf(fz, wm[0], M.wbase[0])));
        const float uy = fmaf(fx, wm[5], fmaf(fy, wm[4], (NDIM == 3 ? fmaf(fz, wm[3], M.wbase[1]) : M.wbase[1])));
        const float ux = fmaf(fx, wm[8], fmaf(fy, wm[7], (NDIM == 3 ? fmaf(fz, wm[6], M.wbase[2]) : M.wbase[2])));
        constexpr float WEPS = 4e-3f;
        bool near = ux < WEPS || ux > 4.f - WEPS || uy < WEPS || uy > 4.f - WEPS;
        if (NDIM == 3) near = near || uz < WEPS || uz > 4.f - WEPS;
#ifdef MVS_AFF_EXACT_W
        near = true;
#endif
        if (near) {
          const double eux = aff_exact(X.wmatrix + 6, X.woffset[2], cz, cy, cx, NDIM == 3);
          const double euy = aff_exact(X.wmatrix + 3, X.woffset[1], cz, cy, cx, NDIM == 3);
          const double euz = NDIM == 3 ? aff_exact(X.wmatrix + 0, X.woffset[0], cz, cy, cx, true) : 0.0;
          bw = aff_weight_exact<NDIM>(tab, euz, euy, eux);
        } else {
          const float fux = floorf(ux), fuy = floorf(uy), fuz = floorf(uz);
          const int ix = (int)fux, iy = (int)fuy, iz = NDIM == 3 ? (int)fuz : 0;
          const float tx = ux - fux, ty = uy - fuy, tz = uz - fuz;
          const int ix1 = min(ix + 1, 4), iy1 = min(iy + 1, 4), iz1 = min(iz + 1, 4);
          const float* p0 = tab + iz * 25;
          float w = lerp_s(lerp_s(p0[iy * 5 + ix], p0[iy * 5 + ix1], tx), lerp_s(p0[iy1 * 5 + ix], p0[iy1 * 5 + ix1], tx), ty);
          if (NDIM == 3) {
            const float* p1 = tab + iz1 * 25;
            const float w1 = lerp_s(lerp_s(p1[iy * 5 + ix], p1[iy * 5 + ix1], tx), lerp_s(p1[iy1 * 5 + ix], p1[iy1 * 5 + ix1], tx), ty);
            w = lerp_s(w, w1, tz);
          }
          if (w < 0.02f) {
            // tiny weights: the reference's float32 (cos + 1) / 2 is quantised there, and where
            // ALL views are near their borders the quantised values decide the blend
            const float a = __fmul_rn(__fsub_rn(1.0f, w), 3.14159274101257324f);
            bw = __fdiv_rn(__fadd_rn(cosf(a), 1.0f), 2.0f);
          } else {
            bw = w < 1.0f ? cosine_ramp(w) : 1.0f;
          }
          bw = fminf(fmaxf(bw, 0.f), 1.f);
        }
      }
      sw[k] = __fadd_rn(sw[k], bw);
      res[k] = __fadd_rn(res[k], __fmul_rn(v, bw));
      if (bw > 0.0f) { vone[k] = v; ++npos[k]; }
    }
  }

  // ---- finalise and store ----
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int dz = NDIM == 3 ? k : 0;
    const int dy = NDIM == 3 ? dy0 : dy0 + 4 * k;
    if (!in_x || y0 + dy >= sh_y || z0 + dz >= sh_z) continue;
    float r = res[k], d = sw[k];
    if (MODE == MVS_FUSE_WAVG) {
      if (!PARTIAL) r = npos[k] == 0 ? 0.0f : (npos[k] == 1 ? vone[k] : __fdiv_rn(res[k], sw[k]));
    } else if (MODE == MVS_FUSE_MEAN) {
      r = npos[k] ? __fdiv_rn(res[k], sw[k]) : 0.0f;
    } else {
      r = npos[k] ? res[k] : 0.0f;
    }
    const int64_t o = (int64_t)(z0 + dz) * ck.stride[0] + (int64_t)(y0 + dy) * ck.stride[1] + (int64_t)(x0 + dx) * ck.stride[2];
    if (PARTIAL) {
      ck.acc_num[o] = r;
      ck.acc_den[o] = d;
    } else {
      store_from_float(ck.out, ck.out_dtype, o, r);
    }
  }
}

}  // namespace mvs
