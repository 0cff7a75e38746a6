// Translation fast path of the fused resample-blend kernel (sm_100a).
//
// For the tile-stitching case the pixel matrix handed to scipy is exactly the
// identity (transformation.py:56 with equal spacings), so the order-1 resample
// of a view is a constant-coefficient 2^ndim-tap stencil on a shifted window:
//   x_in = o + off,  floor(x_in) = o + floor(off),  frac = off - floor(off).
//
// Persistent, warp-specialised kernel.  CTAs pull output blocks (2-D: 32 x 64,
// 3-D: 4 x 32 x 32 voxels) from a global counter:
//   * warp 8, the producer, culls the chunk's views against the block,
//     classifies their blending weights (all ones / all positive / general) from
//     the weight at the corners of block x valid-box, and pulls each contributing
//     view's input footprint into a ring of shared-memory slots with ONE TMA
//     tensor copy (cp.async.bulk.tensor -> UTMALDG; out-of-bounds elements are
//     zero-filled by the hardware), completion signalled on an mbarrier;
//   * warps 0-7 consume the slots: lanes along x (conflict-free shared-memory
//     reads whatever the sub-vector misalignment of the box), x-interpolated
//     rows reused between neighbouring output rows / planes, blending with the
//     reference's float32 semantics, coalesced 128-byte row-segment stores.
// HBM is read in full row segments through TMA and written once.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace mvs {

struct StencilXform {
  int shift[3];   // window px of tap 0 = sample index + shift
  float t[3];     // interpolation fraction per axis
  int d1[3];      // offset of the second tap (0 when the fraction is 0 / order 0)
  int omin[3];    // valid sample-index range per axis (inclusive)
  int omax[3];
  int tmap;       // index of the view's tensor map
  double wm[3];   // diagonal of wmatrix
  double woff[3];
};

template <int NDIM>
struct SBlock {
  // 3-D blocks are 64 wide: a tile overlap (tens of pixels) then makes fewer
  // block columns two-view blocks than with 128-wide blocks
  // Block widths along x (compile-time knobs).  Narrow blocks mean that a tile
  // overlap (tens of pixels) turns fewer block columns into two-view blocks:
  // measured on C3, 4x8x128 -> 7.84 ms, 4x16x64 -> 6.55 ms, 4x32x32 -> 6.20 ms.
#ifndef MVS_BX3
#define MVS_BX3 32
#endif
#ifndef MVS_BX2
#define MVS_BX2 64  // C2: 16x128 -> 0.204 ms, 32x64 -> 0.186 ms, 64x32 -> 0.187 ms
#endif
  static constexpr int BX = NDIM == 3 ? MVS_BX3 : MVS_BX2;
  static constexpr int BY = NDIM == 3 ? 1024 / MVS_BX3 : 2048 / MVS_BX2;
  static constexpr int BZ = NDIM == 3 ? 4 : 1;
  static constexpr int ROWS_Y = BY + 1;
  static constexpr int ROWS_Z = NDIM == 3 ? BZ + 1 : 1;
  static constexpr int NROWS = ROWS_Y * ROWS_Z;
  static constexpr int OUTS = 8;  // outputs per thread: one column x 8 rows
  // consumer warps: 2-D 4 column groups x 2 row groups; 3-D 2 column groups x 2 row groups x 4 planes
  static constexpr int CWARPS = NDIM == 3 ? 16 : 8;
  static constexpr int THREADS = (CWARPS + 1) * 32;
  static constexpr int NW = BX + BY + BZ;
};

// staged row pitch in elements: >= BX + 1 and a multiple of 16 bytes
template <int NDIM, typename T>
struct BoxW { static constexpr int value = SBlock<NDIM>::BX + 16 / (int)sizeof(T); };

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
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// suspend-time hint of try_wait: a waiting warp sleeps in hardware instead of spinning through
// issue slots the working warps need (the kernels are issue-bound)
#ifndef MVS_MBAR_SUSPEND_NS
#define MVS_MBAR_SUSPEND_NS 20000u
#endif
#ifndef MVS_MBAR_SLEEP_NS
#define MVS_MBAR_SLEEP_NS 0
#endif
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  unsigned spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(MVS_MBAR_SUSPEND_NS)
        : "memory");
#if MVS_MBAR_SLEEP_NS > 0
    if (!done) __nanosleep(MVS_MBAR_SLEEP_NS);  // free the issue slots for the warps that have data
#endif
    if (!done && ++spins > (1u << 26)) __trap();  // never hang the GPU on a lost copy
  }
}
// TMA tile copy global -> shared, 2-D / 3-D (coordinates innermost first)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int cx, int cy,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(cx), "r"(cy), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int cx, int cy,
                                            int cz, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(cx), "r"(cy), "r"(cz), "r"(smem_u32(bar))
      : "memory");
}

// staged element -> float.  Integer samples avoid the quarter-rate I2F pipe: the value is
// OR-ed into the mantissa of 2^23 and 2^23 is subtracted (exact for < 2^23; LOP3 + FADD)
#ifndef MVS_MAGIC_CVT
#define MVS_MAGIC_CVT 0  // measured: no difference (the kernels are not I2F-bound)
#endif
__device__ __forceinline__ float tofl(float v) { return v; }
__device__ __forceinline__ float tofl(unsigned short v) {
#if MVS_MAGIC_CVT
  return __uint_as_float(0x4B000000u | (unsigned)v) - 8388608.0f;
#else
  return (float)v;
#endif
}
__device__ __forceinline__ float tofl(unsigned char v) {
#if MVS_MAGIC_CVT
  return __uint_as_float(0x4B000000u | (unsigned)v) - 8388608.0f;
#else
  return (float)v;
#endif
}
__device__ __forceinline__ float lerp_s(float a, float b, float t) {
  return fmaf(t, b, fmaf(-t, a, a));
}

__device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// raw (pre-cosine) table value at float64 table coordinates; 0 outside [0,4]
template <int NDIM>
__device__ float raw_table_value(const float* __restrict__ tab, double uz, double uy, double ux) {
  if (ux < 0.0 || ux > 4.0 || uy < 0.0 || uy > 4.0) return 0.f;
  if (NDIM == 3 && (uz < 0.0 || uz > 4.0)) return 0.f;
  const double fx = floor(ux), fy = floor(uy), fz = floor(uz);
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  const float tx = (float)(ux - fx), ty = (float)(uy - fy), tz = (float)(uz - fz);
  const int ix1 = min(ix + 1, 4), iy1 = min(iy + 1, 4), iz1 = min(iz + 1, 4);
  if (NDIM == 3) {
    const float* p0 = tab + iz * 25;
    const float* p1 = tab + iz1 * 25;
    float a0 = lerp_s(lerp_s(__ldg(p0 + iy * 5 + ix), __ldg(p0 + iy * 5 + ix1), tx),
                      lerp_s(__ldg(p0 + iy1 * 5 + ix), __ldg(p0 + iy1 * 5 + ix1), tx), ty);
    float a1 = lerp_s(lerp_s(__ldg(p1 + iy * 5 + ix), __ldg(p1 + iy * 5 + ix1), tx),
                      lerp_s(__ldg(p1 + iy1 * 5 + ix), __ldg(p1 + iy1 * 5 + ix1), tx), ty);
    return lerp_s(a0, a1, tz);
  }
  return lerp_s(lerp_s(__ldg(tab + iy * 5 + ix), __ldg(tab + iy * 5 + ix1), tx),
                lerp_s(__ldg(tab + iy1 * 5 + ix), __ldg(tab + iy1 * 5 + ix1), tx), ty);
}

// same with plain loads (table staged in shared memory)
template <int NDIM>
__device__ float raw_table_value_s(const float* __restrict__ tab, double uz, double uy, double ux) {
  if (ux < 0.0 || ux > 4.0 || uy < 0.0 || uy > 4.0) return 0.f;
  if (NDIM == 3 && (uz < 0.0 || uz > 4.0)) return 0.f;
  const double fx = floor(ux), fy = floor(uy), fz = floor(uz);
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  const float tx = (float)(ux - fx), ty = (float)(uy - fy), tz = (float)(uz - fz);
  const int ix1 = min(ix + 1, 4), iy1 = min(iy + 1, 4), iz1 = min(iz + 1, 4);
  if (NDIM == 3) {
    const float* p0 = tab + iz * 25;
    const float* p1 = tab + iz1 * 25;
    float a0 = lerp_s(lerp_s(p0[iy * 5 + ix], p0[iy * 5 + ix1], tx), lerp_s(p0[iy1 * 5 + ix], p0[iy1 * 5 + ix1], tx), ty);
    float a1 = lerp_s(lerp_s(p1[iy * 5 + ix], p1[iy * 5 + ix1], tx), lerp_s(p1[iy1 * 5 + ix], p1[iy1 * 5 + ix1], tx), ty);
    return lerp_s(a0, a1, tz);
  }
  return lerp_s(lerp_s(tab[iy * 5 + ix], tab[iy * 5 + ix1], tx), lerp_s(tab[iy1 * 5 + ix], tab[iy1 * 5 + ix1], tx), ty);
}

// Smallest pre-cosine table value x for which (cos((1-x)*pi)+1)/2 is safely > 0
// in float32 (delta^2/2 >= ~3 ulp of 2^-24 with delta = pi*x).
#define MVS_POSITIVE_X 1.8e-4f

// weight classes of a (block, view) pairing
enum { VIEW_GENERAL = 1, VIEW_POSITIVE = 2, VIEW_UNIT = 3 };
// item flags
enum { ITEM_FIRST = 1, ITEM_LAST = 2, ITEM_EMPTY = 4, ITEM_STOP = 8, ITEM_SIMPLE = 16 };

// four consecutive staged elements as floats (16-byte aligned for every T)
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const unsigned short* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  return make_float4((float)(v.x & 0xffffu), (float)(v.x >> 16), (float)(v.y & 0xffffu),
                     (float)(v.y >> 16));
}
__device__ __forceinline__ float4 load4(const unsigned char* p) {
  const unsigned v = *reinterpret_cast<const unsigned*>(p);
  return make_float4((float)(v & 0xffu), (float)((v >> 8) & 0xffu), (float)((v >> 16) & 0xffu),
                     (float)(v >> 24));
}

// float32 -> output dtype like np.nan_to_num(x).astype(dtype) for in-range values
__device__ __forceinline__ float fix_nan(float v) { return v != v ? 0.f : v; }
__device__ __forceinline__ unsigned to_u16(float v) {
  int q = __float2int_rz(fix_nan(v));
  return (unsigned)(q < 0 ? 0 : (q > 65535 ? 65535 : q));
}
__device__ __forceinline__ unsigned to_u8(float v) {
  int q = __float2int_rz(fix_nan(v));
  return (unsigned)(q < 0 ? 0 : (q > 255 ? 255 : q));
}

// float32 -> integer like np.nan_to_num(x).astype(dtype) for in-range values, in ONE
// instruction: PTX float-to-integer conversion truncates (rzi), maps NaN to 0 and saturates
__device__ __forceinline__ unsigned cvt_u16_sat(float v) {
  unsigned short r;
  asm("cvt.rzi.u16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ unsigned cvt_u8_sat(float v) {
  unsigned r;
  asm("{\n\t.reg .u8 t;\n\tcvt.rzi.u8.f32 t, %1;\n\tcvt.u32.u8 %0, t;\n\t}" : "=r"(r) : "f"(v));
  return r;
}
// (cos((1 - x) pi) + 1) / 2 == sin^2(pi x / 2) for x in [0, 1): the square of an odd
// near-minimax polynomial of sin(pi/2 t) (4e-7 relative in float32), i.e. WITHOUT the
// cancellation of the reference's float32 (cos + 1) / 2 near x = 0, whose own rounding noise
// there (6e-8 absolute) is larger than this error.  Callers use it where several views are
// blended (tolerance 1e-4 relative) and fall back to the reference's formula right at a view
// border (x < 0.02), where that formula underflows to an exact 0 and the view is ignored.
__device__ __forceinline__ float cosine_ramp(float x) {
  const float t = fminf(fmaxf(x, 0.f), 1.f);
  const float t2 = t * t;
  float p = -3.431829309e-6f;
  p = fmaf(p, t2, 1.602546911e-4f);
  p = fmaf(p, t2, -4.681657796e-3f);
  p = fmaf(p, t2, 7.969260372e-2f);
  p = fmaf(p, t2, -6.459640956e-1f);
  p = fmaf(p, t2, 1.570796327f);
  const float sn = p * t;
  return sn * sn;
}

// stores 4 consecutive outputs of one row; `nvalid` = how many lie inside the chunk
__device__ __forceinline__ void store_row4(void* out, int dtype, int64_t o, const float* v,
                                           int nvalid) {
  if (dtype == MVS_F32) {
    float* p = reinterpret_cast<float*>(out) + o;
    if (nvalid == 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      *reinterpret_cast<float4*>(p) = make_float4(fix_nan(v[0]), fix_nan(v[1]), fix_nan(v[2]), fix_nan(v[3]));
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) if (c < nvalid) p[c] = fix_nan(v[c]);
    }
  } else if (dtype == MVS_U16) {
    unsigned short* p = reinterpret_cast<unsigned short*>(out) + o;
    if (nvalid == 4 && (reinterpret_cast<uintptr_t>(p) & 7) == 0) {
      *reinterpret_cast<uint2*>(p) = make_uint2(to_u16(v[0]) | (to_u16(v[1]) << 16), to_u16(v[2]) | (to_u16(v[3]) << 16));
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) if (c < nvalid) p[c] = (unsigned short)to_u16(v[c]);
    }
  } else {
    unsigned char* p = reinterpret_cast<unsigned char*>(out) + o;
    if (nvalid == 4 && (reinterpret_cast<uintptr_t>(p) & 3) == 0) {
      *reinterpret_cast<unsigned*>(p) = to_u8(v[0]) | (to_u8(v[1]) << 8) | (to_u8(v[2]) << 16) | (to_u8(v[3]) << 24);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) if (c < nvalid) p[c] = (unsigned char)to_u8(v[c]);
    }
  }
}
__device__ __forceinline__ void store_row4_f32(float* p, const float* v, int nvalid) {
  if (nvalid == 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < nvalid) p[c] = v[c];
  }
}

// Per-block schedule record, computed once per plan by stencil_classify_kernel:
// which of the chunk's views touch the block and how their weights behave there.
struct BlockRec {
  int chunk, first;     // chunk index, its first_xform
  int x0, y0, z0;       // block origin (chunk-local voxels)
  unsigned active;      // bit per view of the chunk (<= 32 views on this path)
  unsigned codes_lo, codes_hi;  // 2 bits per view: VIEW_GENERAL / POSITIVE / UNIT
};

// One warp per output block: cull the chunk's views against the block and
// classify their blending weights.  8 lanes per view evaluate the (pre-cosine)
// weight at the corners of block x valid-box, where it is smallest:
//   min >= 1          -> every weight in the block is exactly 1   (VIEW_UNIT)
//   min >= POSITIVE_X -> every weight in the block is > 0         (VIEW_POSITIVE)
template <int NDIM>
__global__ void __launch_bounds__(256)
stencil_classify_kernel(const mvs_chunk* __restrict__ chunks, const int64_t* __restrict__ block_start,
                        int n_chunks, const mvs_view_xform* __restrict__ xforms,
                        const StencilXform* __restrict__ sxf, const float* __restrict__ tables,
                        BlockRec* __restrict__ recs) {
  using B = SBlock<NDIM>;
  const int lane = threadIdx.x & 31;
  const int64_t nblocks = block_start[n_chunks];
  const int64_t bid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bid >= nblocks) return;
  int lo = 0, hi = n_chunks - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (__ldg(block_start + mid) <= bid) lo = mid; else hi = mid - 1;
  }
  const mvs_chunk& ck = chunks[lo];
  const int local = (int)(bid - __ldg(block_start + lo));
  const int sh_z = ck.shape[0], sh_y = ck.shape[1], sh_x = ck.shape[2];
  const int nbx = (sh_x + B::BX - 1) / B::BX, nby = (sh_y + B::BY - 1) / B::BY;
  const int bz = local / (nbx * nby);
  const int rem = local - bz * (nbx * nby);
  const int by = rem / nbx;
  const int x0 = (rem - by * nbx) * B::BX, y0 = by * B::BY, z0 = bz * B::BZ;
  const int first = ck.first_xform, nxf = ck.n_xforms;
  const int x0s = x0 + ck.halo[2], y0s = y0 + ck.halo[1], z0s = z0 + ck.halo[0];
  const int x1s = min(x0 + B::BX, sh_x) - 1 + ck.halo[2];
  const int y1s = min(y0 + B::BY, sh_y) - 1 + ck.halo[1];
  const int z1s = min(z0 + B::BZ, sh_z) - 1 + ck.halo[0];
  unsigned active = 0;
  unsigned long long codes = 0;
  for (int base = 0; base < nxf; base += 4) {
    const int vi = base + (lane >> 3), c = lane & 7;
    float raw = INFINITY;
    bool act = false;
    if (vi < nxf) {
      const StencilXform& S = sxf[first + vi];
      act = S.omax[2] >= x0s && S.omin[2] <= x1s && S.omax[1] >= y0s && S.omin[1] <= y1s;
      if (NDIM == 3) act = act && S.omax[0] >= z0s && S.omin[0] <= z1s;
      if (act && tables != nullptr) {
        const int ox = (c & 1) ? min(x1s, S.omax[2]) : max(x0s, S.omin[2]);
        const int oy = (c & 2) ? min(y1s, S.omax[1]) : max(y0s, S.omin[1]);
        const int oz = NDIM == 3 ? ((c & 4) ? min(z1s, S.omax[0]) : max(z0s, S.omin[0])) : 0;
        const double ux = __dadd_rn(__dmul_rn((double)ox, S.wm[2]), S.woff[2]);
        const double uy = __dadd_rn(__dmul_rn((double)oy, S.wm[1]), S.woff[1]);
        const double uz = NDIM == 3 ? __dadd_rn(__dmul_rn((double)oz, S.wm[0]), S.woff[0]) : 0.0;
        raw = fmaxf(raw_table_value<NDIM>(tables + (int64_t)xforms[first + vi].table * 125, uz, uy, ux), 0.f);
      }
    }
    raw = fminf(raw, __shfl_xor_sync(0xffffffffu, raw, 1));
    raw = fminf(raw, __shfl_xor_sync(0xffffffffu, raw, 2));
    raw = fminf(raw, __shfl_xor_sync(0xffffffffu, raw, 4));
    int code = 0;
    if (act) {
      code = VIEW_GENERAL;
      if (tables != nullptr) {
        if (raw >= 1.0f) code = VIEW_UNIT;
        else if (raw >= MVS_POSITIVE_X) code = VIEW_POSITIVE;
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int cg_ = __shfl_sync(0xffffffffu, code, g * 8);
      if (base + g < nxf && cg_) {
        active |= 1u << (base + g);
        codes |= (unsigned long long)cg_ << (2 * (base + g));
      }
    }
  }
  if (lane == 0) {
    BlockRec r;
    r.chunk = lo; r.first = first; r.x0 = x0; r.y0 = y0; r.z0 = z0;
    r.active = active; r.codes_lo = (unsigned)codes; r.codes_hi = (unsigned)(codes >> 32);
    recs[bid] = r;
  }
}

// Stores the 16 outputs of one consumer thread (one column; 2-D: 16 consecutive
// rows, 3-D: 2 planes x 8 rows).  Lanes run along x, so every warp store
// instruction writes one contiguous row segment.  CHECK_NAN: float inputs may
// carry NaN data (-> 0 like np.nan_to_num); integer inputs cannot.
template <int NDIM, typename OT, bool CHECK_NAN>
__device__ __forceinline__ void store_column(OT* __restrict__ p, int64_t sy, int64_t sz, int ylim,
                                             int zlim, const float* v) {
  auto conv = [](float x) -> OT {
    if (sizeof(OT) == 4) return (OT)(CHECK_NAN ? fix_nan(x) : x);
    // one saturating convert: truncation toward zero, NaN -> 0, clamped to the type's range
    if (sizeof(OT) == 2) return (OT)cvt_u16_sat(x);
    return (OT)cvt_u8_sat(x);
  };
  if (NDIM == 2) {
    constexpr int NR = SBlock<2>::OUTS;
    if (ylim >= NR) {
#pragma unroll
      for (int k = 0; k < NR; ++k) { *p = conv(v[k]); p += sy; }
    } else {
#pragma unroll
      for (int k = 0; k < NR; ++k) { if (k < ylim) *p = conv(v[k]); p += sy; }
    }
  } else {
    // 3-D: this thread owns one plane (zlim > 0 checked by the caller) x 8 rows
    if (ylim >= 8) {
#pragma unroll
      for (int y = 0; y < 8; ++y) { *p = conv(v[y]); p += sy; }
    } else {
#pragma unroll
      for (int y = 0; y < 8; ++y) { if (y < ylim) *p = conv(v[y]); p += sy; }
    }
  }
}

// One pipeline slot: the staged footprint of one (block, view) item plus what the
// consumer warps need to blend it.
template <int NDIM, typename T>
struct alignas(128) StencilSlot {
  using B = SBlock<NDIM>;
  T stage[B::NROWS * BoxW<NDIM, T>::value];
  float tab[128];
  int wi[B::NW];
  float wt[B::NW];
  StencilXform sx;
  int chunk, x0, y0, z0;  // block origin (chunk-local voxels)
  int flags, wmode, xoff, pad1;  // xoff: staged column of the block's first tap
};

constexpr int kStencilMaxViews = 32;  // views per chunk on this path

#ifndef MVS_NS3
#define MVS_NS3 3
#endif
#ifndef MVS_PREFETCH
#define MVS_PREFETCH 0  // measured: C2 0.197 -> 0.200 ms, C3 unchanged, C5 row 34.8 -> 33.4 ms: the producer is not the bottleneck
#endif
#ifndef MVS_STATIC_SCHED
#define MVS_STATIC_SCHED 0  // measured: static round-robin loses L2 locality (C5 row 35 -> 70 ms)
#endif
template <int NDIM, typename T>
struct StencilStages {
  static constexpr int value = NDIM == 3 ? MVS_NS3 : 4;
};

template <int NDIM, typename T, int MODE, bool PARTIAL>
__global__ void __launch_bounds__(SBlock<NDIM>::THREADS, NDIM == 2 ? 4 : 2)
fuse_stencil_kernel(const mvs_chunk* __restrict__ chunks, const int64_t* __restrict__ block_start,
                    int n_chunks, const mvs_view_xform* __restrict__ xforms,
                    const StencilXform* __restrict__ sxf, const float* __restrict__ tables,
                    const CUtensorMap* __restrict__ tmaps, const BlockRec* __restrict__ recs,
                    unsigned long long* __restrict__ next_block, int64_t block_begin,
                    int64_t block_end) {
  using B = SBlock<NDIM>;
  using Slot = StencilSlot<NDIM, T>;
  constexpr int NS = StencilStages<NDIM, T>::value;
  constexpr int BW = BoxW<NDIM, T>::value;
  constexpr int A = 16 / (int)sizeof(T);  // elements per 16 bytes
  constexpr uint32_t kBoxBytes = (uint32_t)(B::NROWS * BW * sizeof(T));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Slot* slots = reinterpret_cast<Slot*>(smem_raw);
  __shared__ __align__(8) unsigned long long full_bar[NS], empty_bar[NS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], B::CWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t nblocks = min(block_end, (int64_t)block_start[n_chunks]);

  if (warp == B::CWARPS) {
    // =========================== producer warp ===========================
    int it = 0;  // item counter (slot = it % NS)
    auto acquire = [&]() -> Slot& {
      const int s = it % NS;
      mbar_wait(&empty_bar[s], ((it / NS) & 1) ^ 1);
      return slots[s];
    };
    // dynamic schedule: every producer pulls the next block from a global counter,
    // so CTAs stay busy whatever the mix of single- and multi-view blocks, and
    // concurrently processed blocks are neighbours (halo rows hit in L2)
    const int4* recs4 = reinterpret_cast<const int4*>(recs);
#if MVS_STATIC_SCHED
    // static round-robin schedule: block ids are known ahead, so the NEXT block's record is
    // requested while the current block is being issued (the dependent chain atomic ->
    // record -> view constants otherwise costs the lone producer warp ~2 us per block)
    int64_t bid = block_begin + blockIdx.x, bid_n = 0;
    int4 ra = make_int4(0, 0, 0, 0), rb = ra, ra_n = ra, rb_n = ra;
    if (bid < nblocks) { ra = __ldg(recs4 + 2 * bid); rb = __ldg(recs4 + 2 * bid + 1); }
    for (; bid < nblocks; bid = bid_n, ra = ra_n, rb = rb_n) {
      bid_n = bid + gridDim.x;
      if (bid_n < nblocks) { ra_n = __ldg(recs4 + 2 * bid_n); rb_n = __ldg(recs4 + 2 * bid_n + 1); }
#elif MVS_PREFETCH
    // dynamic schedule, two deep: the id of block k+2 is requested and the record of block k+1 is
    // loaded while block k is issued, so the dependent chain atomic -> record -> view constants
    // (~2 us for the lone producer warp) is off the critical path; a CTA over-grabs at most two
    // ids past the end (ids only grow, so an id past the end is never followed by a valid one)
    unsigned long long nb0 = 0, nb1 = 0, nb2 = 0;
    if (lane == 0) { nb0 = atomicAdd(next_block, 1ull); nb1 = atomicAdd(next_block, 1ull); }
    int64_t bid = block_begin + (int64_t)__shfl_sync(0xffffffffu, nb0, 0), bid1 = 0;
    int4 ra = make_int4(0, 0, 0, 0), rb = ra, ra1 = ra, rb1 = ra;
    if (bid < nblocks) { ra = __ldg(recs4 + 2 * bid); rb = __ldg(recs4 + 2 * bid + 1); }
    for (; bid < nblocks; bid = bid1, ra = ra1, rb = rb1, nb1 = nb2) {
      bid1 = block_begin + (int64_t)__shfl_sync(0xffffffffu, nb1, 0);
      if (lane == 0) nb2 = atomicAdd(next_block, 1ull);
      if (bid1 < nblocks) { ra1 = __ldg(recs4 + 2 * bid1); rb1 = __ldg(recs4 + 2 * bid1 + 1); }
#else
    for (;;) {
      unsigned long long nb = 0;
      if (lane == 0) nb = atomicAdd(next_block, 1ull);
      const int64_t bid = block_begin + (int64_t)__shfl_sync(0xffffffffu, nb, 0);
      if (bid >= nblocks) break;
      const int4 ra = __ldg(recs4 + 2 * bid), rb = __ldg(recs4 + 2 * bid + 1);
#endif
      const int ci = ra.x, first = ra.y, x0 = ra.z, y0 = ra.w, z0 = rb.x;
      const unsigned active = (unsigned)rb.y;
      const unsigned long long codes = (unsigned long long)(unsigned)rb.z | ((unsigned long long)(unsigned)rb.w << 32);
      const mvs_chunk& ck = chunks[ci];
      const int x0s = x0 + ck.halo[2], y0s = y0 + ck.halo[1], z0s = z0 + ck.halo[0];
      const int nact = __popc(active);
      // every active view has unit weights (or the mode needs no weights): the
      // consumers can use plain sums (acc = sum v, den = count)
      bool simple = true;
      if (MODE == MVS_FUSE_WAVG)
        for (unsigned rest = active; rest; rest &= rest - 1)
          simple = simple && ((codes >> (2 * (__ffs(rest) - 1))) & 3) == VIEW_UNIT;

      if (nact == 0) {
        Slot& sl = acquire();
        if (lane == 0) {
          sl.chunk = ci; sl.x0 = x0; sl.y0 = y0; sl.z0 = z0;
          sl.flags = ITEM_FIRST | ITEM_LAST | ITEM_EMPTY; sl.wmode = 0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[it % NS]);
        ++it;
        continue;
      }
      int seen = 0;
      for (unsigned rest = active; rest; rest &= rest - 1) {
        const int vi = __ffs(rest) - 1;
        const int code = (int)((codes >> (2 * vi)) & 3);
        const StencilXform& S = sxf[first + vi];
        // weights: 0 = not needed (out = v), 1 = all ones, 2 = table lookup
        int wmode = 0;
        if (MODE == MVS_FUSE_WAVG) {
          if (nact == 1 && !PARTIAL && code >= VIEW_POSITIVE) wmode = 0;
          else wmode = code == VIEW_UNIT ? 1 : 2;
        }
        Slot& sl = acquire();
        {
          const int* src = reinterpret_cast<const int*>(&S);
          int* dst = reinterpret_cast<int*>(&sl.sx);
          for (int q = lane; q < (int)(sizeof(StencilXform) / 4); q += 32) dst[q] = src[q];
          if (lane == 0) {
            sl.chunk = ci; sl.x0 = x0; sl.y0 = y0; sl.z0 = z0;
            sl.flags = (seen == 0 ? ITEM_FIRST : 0) | (seen == nact - 1 ? ITEM_LAST : 0) |
                       (simple ? ITEM_SIMPLE : 0);
            sl.wmode = wmode;
          }
        }
        if (wmode == 2) {
          for (int q = lane; q < B::NW; q += 32) {
            int d, o;
            if (q < B::BX) { d = 2; o = x0s + q; }
            else if (q < B::BX + B::BY) { d = 1; o = y0s + (q - B::BX); }
            else { d = 0; o = z0s + (q - B::BX - B::BY); }
            const double wm = d == 2 ? S.wm[2] : (d == 1 ? S.wm[1] : S.wm[0]);
            const double wo = d == 2 ? S.woff[2] : (d == 1 ? S.woff[1] : S.woff[0]);
            const double u = __dadd_rn(__dmul_rn((double)o, wm), wo);
            int cell = -1;
            float fr = 0.f;
            if (!(u < 0.0 || u > 4.0)) { const double f = floor(u); cell = (int)f; fr = (float)(u - f); }
            sl.wi[q] = cell; sl.wt[q] = fr;
          }
          const float* tab = tables + (int64_t)xforms[first + vi].table * 125;
          for (int q = lane; q < (NDIM == 3 ? 125 : 25); q += 32) sl.tab[q] = __ldg(tab + q);
        }
        // TMA needs a 16-byte aligned innermost start coordinate
        const int x0g = x0s + S.shift[2];
        const int xa = floor_div(x0g, A) * A;
        if (lane == 0) sl.xoff = x0g - xa;
        __syncwarp();
        if (lane == 0) {
          unsigned long long* fb = &full_bar[it % NS];
          mbar_expect_tx(fb, kBoxBytes);
          const CUtensorMap* map = tmaps + S.tmap;
          if (NDIM == 3)
            tma_load_3d(sl.stage, map, xa, y0s + S.shift[1], z0s + S.shift[0], fb);
          else
            tma_load_2d(sl.stage, map, xa, y0s + S.shift[1], fb);
        }
        ++it;
        ++seen;
      }
    }
    // stop item
    {
      Slot& sl = acquire();
      if (lane == 0) sl.flags = ITEM_STOP;
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[it % NS]);
    }
    return;
  }

  // ============================ consumer warps ============================
  // lanes run along x (conflict-free shared-memory reads for any sub-vector
  // misalignment of the staged box): thread = one column x 16 rows.
  // 2-D: column (w&3)*32 + lane, rows (w>>2)*8 + k.   3-D: plane w>>2, rows k.
  // 2-D: column group warp & 3, row group warp >> 2.  3-D: column group warp & 1,
  // row group (warp >> 1) & 1, plane warp >> 2.
  constexpr int CG = B::BX / 32, RG = B::BY / B::OUTS;  // column / row groups of a plane
  const int cg = warp % CG;
  const int zpl = NDIM == 3 ? warp / (CG * RG) : 0;
  const int yoff = ((warp / CG) % RG) * B::OUTS;
  const int jx = cg * 32 + lane;

  // writes this thread's 16 outputs of the block at (x0, y0, z0)
  // clean: no output of this thread can be NaN (the lone-view path has already zeroed them)
  auto store_block = [&](const mvs_chunk& ck, int x0, int y0, int z0, const float* v, const float* d,
                         bool clean = false) {
    const int xo = x0 + jx;
    if (xo >= ck.shape[2]) return;
    const int64_t sy = ck.stride[1], sz = ck.stride[0];
    const int zrow0 = zpl, yrow0 = yoff;
    const int64_t o0 = (int64_t)(z0 + zrow0) * sz + (int64_t)(y0 + yrow0) * sy + (int64_t)xo;
    const int ylim = ck.shape[1] - y0 - yrow0;
    const int zlim = NDIM == 3 ? ck.shape[0] - z0 - zrow0 : 1;
    if (zlim <= 0) return;
    constexpr bool kNan = sizeof(T) == 4;  // float32 views may hold NaN data
    if (PARTIAL) {
      store_column<NDIM, float, false>(ck.acc_num + o0, sy, sz, ylim, zlim, v);
      store_column<NDIM, float, false>(ck.acc_den + o0, sy, sz, ylim, zlim, d);
    } else if (ck.out_dtype == MVS_F32) {
      if (clean) store_column<NDIM, float, false>(reinterpret_cast<float*>(ck.out) + o0, sy, sz, ylim, zlim, v);
      else store_column<NDIM, float, kNan>(reinterpret_cast<float*>(ck.out) + o0, sy, sz, ylim, zlim, v);
    } else if (ck.out_dtype == MVS_U16) {
      store_column<NDIM, unsigned short, kNan>(reinterpret_cast<unsigned short*>(ck.out) + o0, sy, sz, ylim, zlim, v);
    } else {
      store_column<NDIM, unsigned char, kNan>(reinterpret_cast<unsigned char*>(ck.out) + o0, sy, sz, ylim, zlim, v);
    }
  };

  constexpr unsigned kAll = (1u << B::OUTS) - 1u;
  float acc[B::OUTS], den[B::OUTS];
  unsigned anymask = 0;    // bit k: a valid view was seen for output k
  unsigned multimask = 0;  // bit k: at least two valid views (WAVG: acc is weighted)

  for (int it = 0;; ++it) {
    const int s = it % NS;
    Slot& sl = slots[s];
    mbar_wait(&full_bar[s], (it / NS) & 1);
    const int flags = sl.flags;
    if (flags & ITEM_STOP) break;
    const int wmode = sl.wmode;
    const mvs_chunk& ck = chunks[sl.chunk];
    const int x0 = sl.x0, y0 = sl.y0, z0 = sl.z0;
    const int sh_z = ck.shape[0], sh_y = ck.shape[1], sh_x = ck.shape[2];
    if (flags & ITEM_FIRST) {
      anymask = 0; multimask = 0;
#pragma unroll
      for (int k = 0; k < B::OUTS; ++k) { acc[k] = 0.f; den[k] = 0.f; }
    }
    const bool lone = (flags & (ITEM_FIRST | ITEM_LAST | ITEM_EMPTY)) == (ITEM_FIRST | ITEM_LAST) &&
                      (MODE != MVS_FUSE_WAVG || (wmode == 0 && !PARTIAL));
    if (!(flags & ITEM_EMPTY)) {
      const StencilXform& S = sl.sx;
      const int x0s = x0 + ck.halo[2], y0s = y0 + ck.halo[1], z0s = z0 + ck.halo[0];
      // ---- valid bits of this thread's 16 outputs (ranges -> masks) ----
      unsigned vm;
      {
        const int sx = x0s + jx;
        const bool vx = sx >= S.omin[2] && sx <= S.omax[2] && x0 + jx < sh_x;
        const int ya = max(S.omin[1] - y0s, 0), yb = min(S.omax[1] - y0s, sh_y - 1 - y0);
        if (NDIM == 2) {
          const int ka = max(ya - yoff, 0), kb = min(yb - yoff, B::OUTS - 1);
          vm = (vx && kb >= ka) ? ((kAll >> (B::OUTS - 1 - kb)) & (kAll << ka) & kAll) : 0u;
        } else {
          const int ka = max(ya - yoff, 0), kb = min(yb - yoff, 7);
          const unsigned ym = kb >= ka ? ((0xffu >> (7 - kb)) & (0xffu << ka)) : 0u;
          const int za = max(S.omin[0] - z0s, 0), zb = min(S.omax[0] - z0s, sh_z - 1 - z0);
          vm = (vx && zpl >= za && zpl <= zb) ? ym : 0u;
        }
      }
      const float tx = S.t[2], ty = S.t[1], tz = NDIM == 3 ? S.t[0] : 0.f;
      const int c0 = sl.xoff + jx;
      // float32 views: the second tap always takes part, like in scipy, whose order-1 spline
      // multiplies a NaN neighbour by its zero weight at integer positions (NaN data spreads
      // one voxel towards lower indices); integer views skip it when the fraction is 0
      constexpr bool kF = sizeof(T) == 4;
      const int c1 = c0 + (kF ? 1 : S.d1[2]);
      const bool dy = kF || S.d1[1] != 0, dz = NDIM == 3 && (kF || S.d1[0] != 0);

      // ---- interpolate this thread's outputs from shared memory ----
      float val[B::OUTS];
      if (NDIM == 2) {
        const T* p = sl.stage + yoff * BW;
        float hprev = lerp_s(tofl(p[c0]), tofl(p[c1]), tx);
#pragma unroll
        for (int k = 0; k < B::OUTS; ++k) {
          p += BW;
          const float hn = lerp_s(tofl(p[c0]), tofl(p[c1]), tx);
          val[k] = dy ? lerp_s(hprev, hn, ty) : hprev;
          hprev = hn;
        }
      } else {
        // plane `half` and the one above it: 2 x 9 staged rows -> 8 outputs
        float g0[8];
#pragma unroll
        for (int pz = 0; pz < 2; ++pz) {
          const T* p = sl.stage + ((zpl + pz) * B::ROWS_Y + yoff) * BW;
          float hprev = lerp_s(tofl(p[c0]), tofl(p[c1]), tx);
#pragma unroll
          for (int y = 0; y < 8; ++y) {
            p += BW;
            const float hn = lerp_s(tofl(p[c0]), tofl(p[c1]), tx);
            const float g = dy ? lerp_s(hprev, hn, ty) : hprev;
            if (pz == 0) g0[y] = g;
            else val[y] = dz ? lerp_s(g0[y], g, tz) : g0[y];
            hprev = hn;
          }
        }
      }

      // NaN data inside a float view counts as "outside" for that voxel (the reference zeroes
      // the weight where the transformed view is NaN, fusion/_core.py:1648; nan-aware fusion)
      // One test per thread finds the rare case: a NaN among the outputs makes their sum NaN (so may
      // inf - inf: the exact per-output test below then finds nothing, which is harmless).
      if (sizeof(T) == 4) {
        float nsum = val[0];
#pragma unroll
        for (int k = 1; k < B::OUTS; ++k) nsum += val[k];
        if (__any_sync(0xffffffffu, nsum != nsum)) {
#pragma unroll
          for (int k = 0; k < B::OUTS; ++k)
            if (val[k] != val[k]) vm &= ~(1u << k);
        }
      }

      // ---- combine ----
      if (lone) {
        // the block's only view (weight positive everywhere): out = v
        if (!__all_sync(0xffffffffu, vm == kAll)) {
#pragma unroll
          for (int k = 0; k < B::OUTS; ++k) val[k] = (vm >> k) & 1 ? val[k] : 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);  // slot consumed
        store_block(ck, x0, y0, z0, val, nullptr, true);
        continue;
      } else if (MODE == MVS_FUSE_MAX) {
#pragma unroll
        for (int k = 0; k < B::OUTS; ++k)
          if ((vm >> k) & 1) acc[k] = (anymask >> k) & 1 ? fmaxf(acc[k], val[k]) : val[k];
        anymask |= vm;
      } else if (MODE == MVS_FUSE_MEAN || (flags & ITEM_SIMPLE)) {
        // unit weights: acc = sum of valid values, den = number of valid views
#pragma unroll
        for (int k = 0; k < B::OUTS; ++k) {
          const bool valid = (vm >> k) & 1;
          acc[k] = __fadd_rn(acc[k], valid ? val[k] : 0.f);
          den[k] = __fadd_rn(den[k], valid ? 1.f : 0.f);
        }
      } else {
        // General weights.  A thread owns one column (and plane): the x (and z)
        // parts of the multilinear table lookup are constant for its 8 outputs, so
        // the table is first reduced to the <= 3 y-nodes the block can touch
        // (A[0..2]); each output then costs one lerp (+ the cosine near borders).
        int ix = 0, iz = 0, iyA = 0;
        float A0 = 0.f, A1 = 0.f, A2 = 0.f;
        bool colin = true;
        if (wmode == 2) {
          ix = sl.wi[jx];
          const float wtx = sl.wt[jx];
          const int ixc = max(ix, 0), ix1 = min(ixc + 1, 4);
          const int ky0 = yoff;
          iyA = max(sl.wi[B::BX + ky0], 0);
          const int n0 = min(iyA, 4), n1 = min(iyA + 1, 4), n2 = min(iyA + 2, 4);
          if (NDIM == 3) {
            iz = sl.wi[B::BX + B::BY + zpl];
            const float wtz = sl.wt[B::BX + B::BY + zpl];
            const int izc = max(iz, 0), iz1 = min(izc + 1, 4);
            const float* p0 = sl.tab + izc * 25;
            const float* p1 = sl.tab + iz1 * 25;
            A0 = lerp_s(lerp_s(p0[n0 * 5 + ixc], p0[n0 * 5 + ix1], wtx), lerp_s(p1[n0 * 5 + ixc], p1[n0 * 5 + ix1], wtx), wtz);
            A1 = lerp_s(lerp_s(p0[n1 * 5 + ixc], p0[n1 * 5 + ix1], wtx), lerp_s(p1[n1 * 5 + ixc], p1[n1 * 5 + ix1], wtx), wtz);
            A2 = lerp_s(lerp_s(p0[n2 * 5 + ixc], p0[n2 * 5 + ix1], wtx), lerp_s(p1[n2 * 5 + ixc], p1[n2 * 5 + ix1], wtx), wtz);
          } else {
            A0 = lerp_s(sl.tab[n0 * 5 + ixc], sl.tab[n0 * 5 + ix1], wtx);
            A1 = lerp_s(sl.tab[n1 * 5 + ixc], sl.tab[n1 * 5 + ix1], wtx);
            A2 = lerp_s(sl.tab[n2 * 5 + ixc], sl.tab[n2 * 5 + ix1], wtx);
          }
          colin = ix >= 0 && iz >= 0;
        }
#pragma unroll
        for (int k = 0; k < B::OUTS; ++k) {
          const int ky = yoff + k;
          const bool valid = (vm >> k) & 1;
          float b = valid ? 1.f : 0.f;
          if (wmode == 2) {
            const int iy = sl.wi[B::BX + ky];
            const float wty = sl.wt[B::BX + ky];
            const int d = max(iy, 0) - iyA;  // 0 or 1 (table cells are wider than a block)
            const bool inside = colin && iy >= 0;
            float w = d == 0 ? lerp_s(A0, A1, wty) : lerp_s(A1, A2, wty);
            if (d > 1 || d < 0) {
              // tiny views (table cell narrower than the block): full lookup
              const float wtx = sl.wt[jx];
              const int ixc = max(ix, 0), ix1 = min(ixc + 1, 4);
              const int iyc = max(iy, 0), iy1 = min(iyc + 1, 4);
              if (NDIM == 3) {
                const float wtz = sl.wt[B::BX + B::BY + zpl];
                const int izc = max(iz, 0), iz1 = min(izc + 1, 4);
                const float* p0 = sl.tab + izc * 25;
                const float* p1 = sl.tab + iz1 * 25;
                const float a0 = lerp_s(lerp_s(p0[iyc * 5 + ixc], p0[iyc * 5 + ix1], wtx),
                                        lerp_s(p1[iyc * 5 + ixc], p1[iyc * 5 + ix1], wtx), wtz);
                const float a1 = lerp_s(lerp_s(p0[iy1 * 5 + ixc], p0[iy1 * 5 + ix1], wtx),
                                        lerp_s(p1[iy1 * 5 + ixc], p1[iy1 * 5 + ix1], wtx), wtz);
                w = lerp_s(a0, a1, wty);
              } else {
                w = lerp_s(lerp_s(sl.tab[iyc * 5 + ixc], sl.tab[iyc * 5 + ix1], wtx),
                           lerp_s(sl.tab[iy1 * 5 + ixc], sl.tab[iy1 * 5 + ix1], wtx), wty);
              }
            }
            if (__any_sync(0xffffffffu, valid && w < 1.0f)) {
              // weights.py:502-507 cosine ramp, only near view borders
              float cw = cosine_ramp(w);
              if (__any_sync(0xffffffffu, valid && w < 0.02f)) {
                const float a = __fmul_rn(__fsub_rn(1.0f, w), 3.14159274101257324f);
                const float ce = __fmul_rn(__fadd_rn(cosf(a), 1.0f), 0.5f);
                cw = w < 0.02f ? ce : cw;
              }
              w = w < 1.0f ? cw : w;
            }
            w = fminf(fmaxf(w, 0.0f), 1.0f);
            b = (valid && inside) ? w : 0.f;
          }
          // acc holds the raw value while only one view has contributed (its
          // normalised weight is exactly 1); the first view is weighted
          // retroactively once a second one arrives.
          if (valid) {
            if (!((anymask >> k) & 1)) { acc[k] = val[k]; den[k] = b; }
            else {
              const float first_term = ((multimask >> k) & 1) ? acc[k] : __fmul_rn(acc[k], den[k]);
              acc[k] = __fadd_rn(first_term, __fmul_rn(val[k], b));
              den[k] = __fadd_rn(den[k], b);
              multimask |= 1u << k;
            }
          }
        }
        anymask |= vm;
      }
    }
    // slot fully consumed by this warp
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (!(flags & ITEM_LAST)) continue;

    // ---- finalise (registers only) ----
    if (flags & ITEM_EMPTY) {
      // acc holds zeros
    } else if (MODE == MVS_FUSE_MAX) {
#pragma unroll
      for (int k = 0; k < B::OUTS; ++k) acc[k] = (anymask >> k) & 1 ? acc[k] : 0.f;
    } else if (MODE == MVS_FUSE_MEAN || (flags & ITEM_SIMPLE)) {
      if (!PARTIAL) {
#pragma unroll
        for (int k = 0; k < B::OUTS; ++k)
          if (__any_sync(0xffffffffu, den[k] > 1.f))
            acc[k] = den[k] > 1.f ? __fdiv_rn(acc[k], den[k]) : acc[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < B::OUTS; ++k) {
        const bool had = (anymask >> k) & 1, multi = (multimask >> k) & 1;
        if (PARTIAL) {
          acc[k] = !had ? 0.f : (multi ? acc[k] : __fmul_rn(acc[k], den[k]));
        } else {
          float r = (had && den[k] > 0.f) ? acc[k] : 0.f;
          if (__any_sync(0xffffffffu, multi))
            r = multi ? __fdiv_rn(acc[k], den[k] == 0.f ? 1.f : den[k]) : r;
          acc[k] = r;
        }
      }
    }

    store_block(ck, x0, y0, z0, acc, den);
  }
}

template <int NDIM, typename T>
constexpr size_t stencil_smem_bytes() {
  return sizeof(StencilSlot<NDIM, T>) * StencilStages<NDIM, T>::value;
}

}  // namespace mvs
