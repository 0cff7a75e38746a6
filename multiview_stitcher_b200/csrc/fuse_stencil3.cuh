// 3-D translation fast path of the fused resample-blend kernel (sm_100a): the
// z-marching variant.
//
// Same contract as fuse_stencil_kernel (fuse_stencil.cuh) -- the pixel matrix is the
// identity, the order-1 resample is a constant-coefficient 8-tap stencil -- but the
// work unit is a COLUMN of the output chunk: 64 (x) x 16 (y) voxels marched along z
// in steps of 4 planes.
//   * producer warp: per step and contributing view ONE TMA tensor copy of the 4 new
//     input planes (box 72 x 17 x 4 for uint16; hardware zero fill out of bounds)
//     into a ring of shared-memory slots.  A column fed by a single view carries the
//     x/y-interpolated last plane in registers from step to step, so no input plane
//     is staged twice (no z halo); columns with several views also stage the plane
//     below (a second, one-plane box) and stay self-contained per step.
//   * consumer warps (8): a thread owns 2 adjacent columns x 2 rows and marches the
//     4 planes: three staged rows per plane are read as aligned 32-bit words (two
//     uint16 taps per load; lanes run along x: conflict-free), x-lerped once, reused
//     for both rows; the z-lerp partner is the previous plane's value in registers.
//     Outputs leave as packed pairs (one 32-bit store per 2 uint16 voxels).
// Per output voxel this is ~15 instructions against ~45 of the block-of-4-planes
// kernel, which was issue-bound at 17-19 % of the HBM roofline on C3 / C5.
#pragma once

#include "fuse_stencil.cuh"

namespace mvs {

struct S3 {
  static constexpr int BX = 64, BY = 16, PZ = 4;
  static constexpr int ROWS = BY + 1;
  static constexpr int CWARPS = 8;  // 2 output rows each
  static constexpr int THREADS = (CWARPS + 1) * 32;
  static constexpr int OUTS = PZ * 4;  // per thread and step: 4 planes x 2 rows x 2 columns
  static constexpr int NW = BX + BY + PZ;
};
enum { ITEM_CARRY = 32 };

template <typename T>
struct S3Pitch { static constexpr int value = S3::BX + 16 / (int)sizeof(T); };

template <typename T>
struct alignas(128) Slot3 {
  static constexpr int BW = S3Pitch<T>::value;
  T main[S3::PZ * S3::ROWS * BW];       // planes z+1 .. z+4 of the step (one TMA box)
  alignas(128) T prime[S3::ROWS * BW];  // plane z (second box; skipped when carried)
  alignas(16) float tab[128];
  int wi[S3::NW];
  float wt[S3::NW];
  // item record, written by the producer (the consumers touch no global metadata)
  unsigned long long obase;   // address of the output voxel (z0, y0, x0): out, or acc_num if partial
  unsigned long long dbase;   // acc_den counterpart (partial mode)
  long long osy, osz;         // element strides of the output rows / planes
  int nx, ny, nz;             // output voxels of this step inside the chunk (from x0, y0, z0)
  int odt;                    // output dtype
  int vx0, vx1, vy0, vy1;     // block-local index range in which the view is valid
  int vzm;                    // bit p: plane z0 + p is valid
  int d1x, d1y, d1z;          // second tap present (fraction != 0)
  float tx, ty, tz;           // interpolation fractions
  int flags, wmode, xoff, pad1;
};

template <typename T>
struct S3Stages { static constexpr int value = sizeof(T) == 4 ? 4 : 5; };

template <typename T>
constexpr size_t stencil3_smem_bytes() { return sizeof(Slot3<T>) * S3Stages<T>::value; }

// three consecutive staged elements c0, c0+1, c0+2 as floats
__device__ __forceinline__ void load3(const float* row, int c0, float& a, float& b, float& c) {
  a = row[c0]; b = row[c0 + 1]; c = row[c0 + 2];
}
__device__ __forceinline__ void load3(const unsigned char* row, int c0, float& a, float& b, float& c) {
  a = (float)row[c0]; b = (float)row[c0 + 1]; c = (float)row[c0 + 2];
}
// uint16: two aligned 32-bit words hold the three taps; which halves depends on the
// (block-uniform) parity of c0, selected with byte permutes
__device__ __forceinline__ void load3(const unsigned short* row, int c0, float& a, float& b, float& c) {
  const unsigned* w = reinterpret_cast<const unsigned*>(row) + (c0 >> 1);
  const unsigned w0 = w[0], w1 = w[1];
  const bool odd = c0 & 1;
  const unsigned ab = __byte_perm(w0, w1, odd ? 0x5432u : 0x3210u);
  const unsigned cw = odd ? (w1 >> 16) : (w1 & 0xffffu);
  a = (float)(ab & 0xffffu);
  b = (float)(ab >> 16);
  c = (float)cw;
}

// two adjacent outputs of one row; nvalid = how many lie inside the chunk (1 or 2)
template <typename OT, bool CHECK_NAN>
__device__ __forceinline__ void store_pair(OT* __restrict__ p, float a, float b, int nvalid) {
  if (CHECK_NAN) { a = fix_nan(a); b = fix_nan(b); }
  if (sizeof(OT) == 4) {
    float* q = reinterpret_cast<float*>(p);
    if (nvalid == 2 && (reinterpret_cast<uintptr_t>(q) & 7) == 0) {
      *reinterpret_cast<float2*>(q) = make_float2(a, b);
    } else {
      q[0] = a;
      if (nvalid == 2) q[1] = b;
    }
  } else if (sizeof(OT) == 2) {
    int qa = __float2int_rz(a), qb = __float2int_rz(b);
    qa = qa < 0 ? 0 : (qa > 65535 ? 65535 : qa);
    qb = qb < 0 ? 0 : (qb > 65535 ? 65535 : qb);
    unsigned short* q = reinterpret_cast<unsigned short*>(p);
    if (nvalid == 2 && (reinterpret_cast<uintptr_t>(q) & 3) == 0) {
      *reinterpret_cast<unsigned*>(q) = (unsigned)qa | ((unsigned)qb << 16);
    } else {
      q[0] = (unsigned short)qa;
      if (nvalid == 2) q[1] = (unsigned short)qb;
    }
  } else {
    int qa = __float2int_rz(a), qb = __float2int_rz(b);
    qa = qa < 0 ? 0 : (qa > 255 ? 255 : qa);
    qb = qb < 0 ? 0 : (qb > 255 ? 255 : qb);
    unsigned char* q = reinterpret_cast<unsigned char*>(p);
    if (nvalid == 2 && (reinterpret_cast<uintptr_t>(q) & 1) == 0) {
      *reinterpret_cast<unsigned short*>(q) = (unsigned short)(qa | (qb << 8));
    } else {
      q[0] = (unsigned char)qa;
      if (nvalid == 2) q[1] = (unsigned char)qb;
    }
  }
}

// One warp per output COLUMN (64 x 16 voxels x the chunk's z extent): cull the chunk's
// views against it and classify their blending weights (see stencil_classify_kernel).
__global__ void __launch_bounds__(256)
stencil3_classify_kernel(const mvs_chunk* __restrict__ chunks, const int64_t* __restrict__ col_start,
                         int n_chunks, const mvs_view_xform* __restrict__ xforms,
                         const StencilXform* __restrict__ sxf, const float* __restrict__ tables,
                         BlockRec* __restrict__ recs) {
  const int lane = threadIdx.x & 31;
  const int64_t ncols = col_start[n_chunks];
  const int64_t bid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bid >= ncols) return;
  int lo = 0, hi = n_chunks - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (__ldg(col_start + mid) <= bid) lo = mid; else hi = mid - 1;
  }
  const mvs_chunk& ck = chunks[lo];
  const int local = (int)(bid - __ldg(col_start + lo));
  const int sh_z = ck.shape[0], sh_y = ck.shape[1], sh_x = ck.shape[2];
  const int nbx = (sh_x + S3::BX - 1) / S3::BX;
  const int by = local / nbx;
  const int x0 = (local - by * nbx) * S3::BX, y0 = by * S3::BY;
  const int first = ck.first_xform, nxf = ck.n_xforms;
  const int x0s = x0 + ck.halo[2], y0s = y0 + ck.halo[1], z0s = ck.halo[0];
  const int x1s = min(x0 + S3::BX, sh_x) - 1 + ck.halo[2];
  const int y1s = min(y0 + S3::BY, sh_y) - 1 + ck.halo[1];
  const int z1s = sh_z - 1 + ck.halo[0];
  unsigned active = 0;
  unsigned long long codes = 0;
  for (int base = 0; base < nxf; base += 4) {
    const int vi = base + (lane >> 3), c = lane & 7;
    float raw = INFINITY;
    bool act = false;
    if (vi < nxf) {
      const StencilXform& S = sxf[first + vi];
      act = S.omax[2] >= x0s && S.omin[2] <= x1s && S.omax[1] >= y0s && S.omin[1] <= y1s &&
            S.omax[0] >= z0s && S.omin[0] <= z1s;
      if (act && tables != nullptr) {
        const int ox = (c & 1) ? min(x1s, S.omax[2]) : max(x0s, S.omin[2]);
        const int oy = (c & 2) ? min(y1s, S.omax[1]) : max(y0s, S.omin[1]);
        const int oz = (c & 4) ? min(z1s, S.omax[0]) : max(z0s, S.omin[0]);
        const double ux = __dadd_rn(__dmul_rn((double)ox, S.wm[2]), S.woff[2]);
        const double uy = __dadd_rn(__dmul_rn((double)oy, S.wm[1]), S.woff[1]);
        const double uz = __dadd_rn(__dmul_rn((double)oz, S.wm[0]), S.woff[0]);
        raw = fmaxf(raw_table_value<3>(tables + (int64_t)xforms[first + vi].table * 125, uz, uy, ux), 0.f);
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
    r.chunk = lo; r.first = first; r.x0 = x0; r.y0 = y0; r.z0 = 0;
    r.active = active; r.codes_lo = (unsigned)codes; r.codes_hi = (unsigned)(codes >> 32);
    recs[bid] = r;
  }
}

template <typename T, int MODE, bool PARTIAL>
__global__ void __launch_bounds__(S3::THREADS, 2)
fuse_stencil3_kernel(const mvs_chunk* __restrict__ chunks, const int64_t* __restrict__ col_start,
                     int n_chunks, const mvs_view_xform* __restrict__ xforms,
                     const StencilXform* __restrict__ sxf, const float* __restrict__ tables,
                     const CUtensorMap* __restrict__ tmaps, const BlockRec* __restrict__ recs,
                     unsigned long long* __restrict__ next_col, int64_t col_begin, int64_t col_end) {
  using Slot = Slot3<T>;
  constexpr int NS = S3Stages<T>::value;
  constexpr int BW = S3Pitch<T>::value;
  constexpr int PZ = S3::PZ;
  constexpr int A = 16 / (int)sizeof(T);
  constexpr uint32_t kMainBytes = (uint32_t)(PZ * S3::ROWS * BW * sizeof(T));
  constexpr uint32_t kPrimeBytes = (uint32_t)(S3::ROWS * BW * sizeof(T));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Slot* slots = reinterpret_cast<Slot*>(smem_raw);
  __shared__ __align__(8) unsigned long long full_bar[NS], empty_bar[NS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], S3::CWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t ncols = min(col_end, (int64_t)col_start[n_chunks]);

  if (warp == S3::CWARPS) {
    // =========================== producer warp ===========================
    int it = 0;
    auto acquire = [&]() -> Slot& {
      const int s = it % NS;
      mbar_wait(&empty_bar[s], ((it / NS) & 1) ^ 1);
      return slots[s];
    };
    const int4* recs4 = reinterpret_cast<const int4*>(recs);
    for (;;) {
      unsigned long long nb = 0;
      if (lane == 0) nb = atomicAdd(next_col, 1ull);
      const int64_t bid = col_begin + (int64_t)__shfl_sync(0xffffffffu, nb, 0);
      if (bid >= ncols) break;
      const int4 ra = __ldg(recs4 + 2 * bid), rb = __ldg(recs4 + 2 * bid + 1);
      const int ci = ra.x, first = ra.y, x0 = ra.z, y0 = ra.w;
      const unsigned active = (unsigned)rb.y;
      const unsigned long long codes = (unsigned long long)(unsigned)rb.z | ((unsigned long long)(unsigned)rb.w << 32);
      const mvs_chunk& ck = chunks[ci];
      const int sh_z = ck.shape[0], sh_y = ck.shape[1], sh_x = ck.shape[2];
      const int x0s = x0 + ck.halo[2], y0s = y0 + ck.halo[1], hz = ck.halo[0];
      const long long osy = ck.stride[1], osz = ck.stride[0];
      const int odt = ck.out_dtype;
      const size_t oes = PARTIAL ? 4 : dtype_size(odt);
      const unsigned long long ob0 = (unsigned long long)(PARTIAL ? (void*)ck.acc_num : ck.out) +
                                     (unsigned long long)(((long long)y0 * osy + x0) * (long long)oes);
      const unsigned long long db0 = PARTIAL ? (unsigned long long)ck.acc_den + (unsigned long long)(((long long)y0 * osy + x0) * 4) : 0ull;
      auto fill_out = [&](Slot& sl, int z0) {
        sl.obase = ob0 + (unsigned long long)((long long)z0 * osz * (long long)oes);
        sl.dbase = db0 + (unsigned long long)((long long)z0 * osz * 4);
        sl.osy = osy; sl.osz = osz; sl.odt = odt;
        sl.nx = sh_x - x0; sl.ny = sh_y - y0; sl.nz = sh_z - z0;
      };
      const int x1s = min(x0 + S3::BX, sh_x) - 1 + ck.halo[2];
      const int y1s = min(y0 + S3::BY, sh_y) - 1 + ck.halo[1];
      int prev_single = -1;  // view whose interpolated last plane the consumers carry
      for (int z0 = 0; z0 < sh_z; z0 += PZ) {
        const int z0s = z0 + hz, z1s = min(z0 + PZ, sh_z) - 1 + hz;
        unsigned act = 0;
        for (unsigned rest = active; rest; rest &= rest - 1) {
          const int vi = __ffs(rest) - 1;
          const StencilXform& S = sxf[first + vi];
          if (S.omax[0] >= z0s && S.omin[0] <= z1s) act |= 1u << vi;
        }
        const int nact = __popc(act);
        if (nact == 0) {
          Slot& sl = acquire();
          if (lane == 0) {
            fill_out(sl, z0);
            sl.flags = ITEM_FIRST | ITEM_LAST | ITEM_EMPTY; sl.wmode = 0;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[it % NS]);
          ++it;
          prev_single = -1;
          continue;
        }
        // Weight classes of THIS step (the column-level classes are too coarse along z: a
        // view's top / bottom ramp lies somewhere in most columns).  8 lanes per view evaluate
        // the raw weight at the corners of step-box x valid-box, 4 views per pass.
        unsigned long long scodes = 0;
        if (MODE == MVS_FUSE_WAVG) {
          unsigned rest = act;
          while (rest) {
            int myv = -1;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int v = rest ? __ffs(rest) - 1 : -1;
              rest &= rest - 1;
              if ((lane >> 3) == g) myv = v;
            }
            float raw = INFINITY;
            int ccode = 0;
            if (myv >= 0) {
              ccode = (int)((codes >> (2 * myv)) & 3);
              if (ccode != VIEW_UNIT) {
                const StencilXform& S = sxf[first + myv];
                const int c = lane & 7;
                const int ox = (c & 1) ? min(x1s, S.omax[2]) : max(x0s, S.omin[2]);
                const int oy = (c & 2) ? min(y1s, S.omax[1]) : max(y0s, S.omin[1]);
                const int oz = (c & 4) ? min(z1s, S.omax[0]) : max(z0s, S.omin[0]);
                const double ux = __dadd_rn(__dmul_rn((double)ox, S.wm[2]), S.woff[2]);
                const double uy = __dadd_rn(__dmul_rn((double)oy, S.wm[1]), S.woff[1]);
                const double uz = __dadd_rn(__dmul_rn((double)oz, S.wm[0]), S.woff[0]);
                raw = fmaxf(raw_table_value<3>(tables + (int64_t)xforms[first + myv].table * 125, uz, uy, ux), 0.f);
              }
            }
            raw = fminf(raw, __shfl_xor_sync(0xffffffffu, raw, 1));
            raw = fminf(raw, __shfl_xor_sync(0xffffffffu, raw, 2));
            raw = fminf(raw, __shfl_xor_sync(0xffffffffu, raw, 4));
            int code = ccode;
            if (myv >= 0 && ccode != VIEW_UNIT)
              code = raw >= 1.0f ? VIEW_UNIT : (raw >= MVS_POSITIVE_X ? VIEW_POSITIVE : VIEW_GENERAL);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int v = __shfl_sync(0xffffffffu, myv, g * 8);
              const int cd = __shfl_sync(0xffffffffu, code, g * 8);
              if (v >= 0) scodes |= (unsigned long long)cd << (2 * v);
            }
          }
        }
        bool simple = true;
        if (MODE == MVS_FUSE_WAVG)
          for (unsigned rest = act; rest; rest &= rest - 1)
            simple = simple && ((scodes >> (2 * (__ffs(rest) - 1))) & 3) == VIEW_UNIT;
        int seen = 0, last_vi = -1;
        for (unsigned rest = act; rest; rest &= rest - 1) {
          const int vi = __ffs(rest) - 1;
          last_vi = vi;
          const int code = MODE == MVS_FUSE_WAVG ? (int)((scodes >> (2 * vi)) & 3) : (int)((codes >> (2 * vi)) & 3);
          const StencilXform& S = sxf[first + vi];
          int wmode = 0;
          if (MODE == MVS_FUSE_WAVG) {
            if (nact == 1 && !PARTIAL && code >= VIEW_POSITIVE) wmode = 0;
            else wmode = code == VIEW_UNIT ? 1 : 2;
          }
          const bool carry = nact == 1 && prev_single == vi;
          Slot& sl = acquire();
          if (lane == 0) {
            fill_out(sl, z0);
            // validity as block-local index ranges / plane bits (sample index = voxel + halo)
            sl.vx0 = max(S.omin[2] - x0s, 0); sl.vx1 = min(S.omax[2] - x0s, sh_x - 1 - x0);
            sl.vy0 = max(S.omin[1] - y0s, 0); sl.vy1 = min(S.omax[1] - y0s, sh_y - 1 - y0);
            int vzm = 0;
            for (int pp = 0; pp < PZ; ++pp)
              if (z0s + pp >= S.omin[0] && z0s + pp <= S.omax[0] && z0 + pp < sh_z) vzm |= 1 << pp;
            sl.vzm = vzm;
            sl.d1x = S.d1[2]; sl.d1y = S.d1[1]; sl.d1z = S.d1[0];
            sl.tx = S.t[2]; sl.ty = S.t[1]; sl.tz = S.t[0];
            sl.flags = (seen == 0 ? ITEM_FIRST : 0) | (seen == nact - 1 ? ITEM_LAST : 0) |
                       (simple ? ITEM_SIMPLE : 0) | (carry ? ITEM_CARRY : 0);
            sl.wmode = wmode;
          }
          if (wmode == 2) {
            for (int q = lane; q < S3::NW; q += 32) {
              int d, o;
              if (q < S3::BX) { d = 2; o = x0s + q; }
              else if (q < S3::BX + S3::BY) { d = 1; o = y0s + (q - S3::BX); }
              else { d = 0; o = z0s + (q - S3::BX - S3::BY); }
              const double wm = d == 2 ? S.wm[2] : (d == 1 ? S.wm[1] : S.wm[0]);
              const double wo = d == 2 ? S.woff[2] : (d == 1 ? S.woff[1] : S.woff[0]);
              const double u = __dadd_rn(__dmul_rn((double)o, wm), wo);
              int cell = -1;
              float fr = 0.f;
              if (!(u < 0.0 || u > 4.0)) { const double f = floor(u); cell = (int)f; fr = (float)(u - f); }
              sl.wi[q] = cell; sl.wt[q] = fr;
            }
            const float* tab = tables + (int64_t)xforms[first + vi].table * 125;
            for (int q = lane; q < 125; q += 32) sl.tab[q] = __ldg(tab + q);
          }
          const int x0g = x0s + S.shift[2];
          const int xa = floor_div(x0g, A) * A;
          if (lane == 0) sl.xoff = x0g - xa;
          __syncwarp();
          if (lane == 0) {
            unsigned long long* fb = &full_bar[it % NS];
            mbar_expect_tx(fb, kMainBytes + (carry ? 0u : kPrimeBytes));
            const CUtensorMap* map = tmaps + 2 * S.tmap;
            const int ys = y0s + S.shift[1], zs = z0s + S.shift[0];
            tma_load_3d(sl.main, map, xa, ys, zs + 1, fb);
            if (!carry) tma_load_3d(sl.prime, map + 1, xa, ys, zs, fb);
          }
          ++it;
          ++seen;
        }
        prev_single = nact == 1 ? last_vi : -1;
      }
    }
    {
      Slot& sl = acquire();
      if (lane == 0) sl.flags = ITEM_STOP;
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[it % NS]);
    }
    return;
  }

  // ============================ consumer warps ============================
  // thread = columns 2*lane, 2*lane+1 x rows 2*warp, 2*warp+1, marching the step's planes;
  // output index k = p*4 + r*2 + c
  constexpr int OUTS = S3::OUTS;
  const int jx = 2 * lane, jy = 2 * warp;
  float acc[OUTS], den[OUTS];
  float gc[4] = {0.f, 0.f, 0.f, 0.f};  // carried x/y-interpolated plane
  unsigned anymask = 0, multimask = 0;
  constexpr bool kNan = sizeof(T) == 4;  // float32 views may hold NaN data

  // item record -> registers (read before the slot is released)
  struct OutRec { unsigned long long obase, dbase; long long osy, osz; int nx, ny, nz, odt; };
  auto rows_of = [&](auto* base, const OutRec& o, const float* v, auto conv_store) {
    // base: typed pointer to voxel (z0, y0 + jy, x0 + jx)
    using PT = decltype(base);
    const int nval = min(2, o.nx - jx);
    const int np = min(PZ, o.nz), nr = min(2, o.ny - jy);
#pragma unroll
    for (int p = 0; p < PZ; ++p) {
      if (p < np) {
        PT q = base + (long long)p * o.osz;
        conv_store(q, v[p * 4], v[p * 4 + 1], nval);
        if (nr > 1) conv_store(q + o.osy, v[p * 4 + 2], v[p * 4 + 3], nval);
      }
    }
  };
  auto store_step = [&](const OutRec& o, const float* v, const float* d) {
    if (jx >= o.nx || jy >= o.ny) return;
    const long long off = (long long)jy * o.osy + jx;
    if (PARTIAL) {
      auto st = [](float* q, float a, float b, int n) { store_pair<float, false>(q, a, b, n); };
      rows_of(reinterpret_cast<float*>(o.obase) + off, o, v, st);
      rows_of(reinterpret_cast<float*>(o.dbase) + off, o, d, st);
    } else if (o.odt == MVS_U16) {
      rows_of(reinterpret_cast<unsigned short*>(o.obase) + off, o, v,
              [](unsigned short* q, float a, float b, int n) { store_pair<unsigned short, kNan>(q, a, b, n); });
    } else if (o.odt == MVS_F32) {
      rows_of(reinterpret_cast<float*>(o.obase) + off, o, v,
              [](float* q, float a, float b, int n) { store_pair<float, kNan>(q, a, b, n); });
    } else {
      rows_of(reinterpret_cast<unsigned char*>(o.obase) + off, o, v,
              [](unsigned char* q, float a, float b, int n) { store_pair<unsigned char, kNan>(q, a, b, n); });
    }
  };

  for (int it = 0;; ++it) {
    const int s = it % NS;
    Slot& sl = slots[s];
    mbar_wait(&full_bar[s], (it / NS) & 1);
    const int flags = sl.flags;
    if (flags & ITEM_STOP) break;
    const int wmode = sl.wmode;
    OutRec orec;
    if (flags & ITEM_LAST) {
      orec.obase = sl.obase; orec.dbase = sl.dbase; orec.osy = sl.osy; orec.osz = sl.osz;
      orec.nx = sl.nx; orec.ny = sl.ny; orec.nz = sl.nz; orec.odt = sl.odt;
    }
    if (flags & ITEM_FIRST) {
      anymask = 0; multimask = 0;
#pragma unroll
      for (int k = 0; k < OUTS; ++k) { acc[k] = 0.f; den[k] = 0.f; }
    }
    const bool lone = (flags & (ITEM_FIRST | ITEM_LAST | ITEM_EMPTY)) == (ITEM_FIRST | ITEM_LAST) &&
                      (MODE != MVS_FUSE_WAVG || (wmode == 0 && !PARTIAL));
    if (!(flags & ITEM_EMPTY)) {
      // ---- valid bits of this thread's 16 outputs: bit k = p*4 + r*2 + c ----
      unsigned vm;
      {
        const int vx0 = sl.vx0, vx1 = sl.vx1, vy0 = sl.vy0, vy1 = sl.vy1;
        const unsigned vxm = (jx >= vx0 && jx <= vx1 ? 1u : 0u) | (jx + 1 >= vx0 && jx + 1 <= vx1 ? 2u : 0u);
        const unsigned rc = (jy >= vy0 && jy <= vy1 ? vxm : 0u) | (jy + 1 >= vy0 && jy + 1 <= vy1 ? vxm << 2 : 0u);
        const unsigned vzm = (unsigned)sl.vzm;
        vm = ((vzm & 1u) ? rc : 0u) | ((vzm & 2u) ? rc << 4 : 0u) | ((vzm & 4u) ? rc << 8 : 0u) | ((vzm & 8u) ? rc << 12 : 0u);
      }
      const float tx = sl.tx, ty = sl.ty, tz = sl.tz;
      const int c0 = sl.xoff + jx;

      // ---- interpolate: x/y per plane, z against the previous plane ----
      float val[OUTS];
      auto plane_g = [&](const T* base, float* g) {
        float h[3][2];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          float a, b, c;
          load3(base + (jy + rr) * BW, c0, a, b, c);
          // the second tap always takes part (scipy multiplies a NaN neighbour by its zero
          // weight at integer positions; for finite data lerp(a, b, 0) == a exactly)
          h[rr][0] = lerp_s(a, b, tx);
          h[rr][1] = lerp_s(b, c, tx);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int c = 0; c < 2; ++c)
            g[r * 2 + c] = lerp_s(h[r][c], h[r + 1][c], ty);
      };
      if (!(flags & ITEM_CARRY)) plane_g(sl.prime, gc);
#pragma unroll
      for (int p = 0; p < PZ; ++p) {
        float g[4];
        plane_g(sl.main + p * (S3::ROWS * BW), g);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          val[p * 4 + k] = lerp_s(gc[k], g[k], tz);
          gc[k] = g[k];
        }
      }

      if (kNan) {  // NaN data = outside for that voxel (fusion/_core.py:1648)
#pragma unroll
        for (int k = 0; k < OUTS; ++k)
          if (val[k] != val[k]) vm &= ~(1u << k);
      }

      // ---- combine ----
      if (lone) {
        if (!__all_sync(0xffffffffu, vm == 0xffffu)) {
#pragma unroll
          for (int k = 0; k < OUTS; ++k) val[k] = (vm >> k) & 1 ? val[k] : 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        store_step(orec, val, nullptr);
        continue;
      } else if (MODE == MVS_FUSE_MAX) {
#pragma unroll
        for (int k = 0; k < OUTS; ++k)
          if ((vm >> k) & 1) acc[k] = (anymask >> k) & 1 ? fmaxf(acc[k], val[k]) : val[k];
        anymask |= vm;
      } else if (MODE == MVS_FUSE_MEAN || (flags & ITEM_SIMPLE)) {
#pragma unroll
        for (int k = 0; k < OUTS; ++k) {
          const bool valid = (vm >> k) & 1;
          acc[k] = __fadd_rn(acc[k], valid ? val[k] : 0.f);
          den[k] = __fadd_rn(den[k], valid ? 1.f : 0.f);
        }
      } else {
        // General weights.  (row, column) are fixed per thread: the table is first reduced
        // to the <= 3 z-nodes the step can touch, per (r, c); each output then costs one
        // lerp along z (+ the cosine near view borders).
        float B[3][4];
        int izA = 0;
        bool rcin[4] = {true, true, true, true};
        if (wmode == 2) {
          izA = max(sl.wi[S3::BX + S3::BY], 0);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int iy = sl.wi[S3::BX + jy + r];
            const float wty = sl.wt[S3::BX + jy + r];
            const int iyc = max(iy, 0), iy1 = min(iyc + 1, 4);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int ix = sl.wi[jx + c];
              const float wtx = sl.wt[jx + c];
              const int ixc = max(ix, 0), ix1 = min(ixc + 1, 4);
              rcin[r * 2 + c] = ix >= 0 && iy >= 0;
#pragma unroll
              for (int n = 0; n < 3; ++n) {
                const float* pl = sl.tab + min(izA + n, 4) * 25;
                B[n][r * 2 + c] = lerp_s(lerp_s(pl[iyc * 5 + ixc], pl[iyc * 5 + ix1], wtx),
                                         lerp_s(pl[iy1 * 5 + ixc], pl[iy1 * 5 + ix1], wtx), wty);
              }
            }
          }
        }
#pragma unroll
        for (int p = 0; p < PZ; ++p) {
          int iz = 0;
          float wtz = 0.f;
          if (wmode == 2) { iz = sl.wi[S3::BX + S3::BY + p]; wtz = sl.wt[S3::BX + S3::BY + p]; }
          const int d = max(iz, 0) - izA;  // 0 or 1 (table cells are wider than a step)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = p * 4 + q;
            const bool valid = (vm >> k) & 1;
            float b = valid ? 1.f : 0.f;
            if (wmode == 2) {
              float w = d == 0 ? lerp_s(B[0][q], B[1][q], wtz) : lerp_s(B[1][q], B[2][q], wtz);
              if (d > 1 || d < 0) {
                // tiny views (table cell thinner than a step): full lookup
                const int r = q >> 1, c = q & 1;
                const int iy = sl.wi[S3::BX + jy + r], ix = sl.wi[jx + c];
                const float wty = sl.wt[S3::BX + jy + r], wtx = sl.wt[jx + c];
                const int iyc = max(iy, 0), iy1 = min(iyc + 1, 4), ixc = max(ix, 0), ix1 = min(ixc + 1, 4);
                const int izc = max(iz, 0), iz1 = min(izc + 1, 4);
                const float* p0 = sl.tab + izc * 25;
                const float* p1 = sl.tab + iz1 * 25;
                const float a0 = lerp_s(lerp_s(p0[iyc * 5 + ixc], p0[iyc * 5 + ix1], wtx),
                                        lerp_s(p0[iy1 * 5 + ixc], p0[iy1 * 5 + ix1], wtx), wty);
                const float a1 = lerp_s(lerp_s(p1[iyc * 5 + ixc], p1[iyc * 5 + ix1], wtx),
                                        lerp_s(p1[iy1 * 5 + ixc], p1[iy1 * 5 + ix1], wtx), wty);
                w = lerp_s(a0, a1, wtz);
              }
              if (__any_sync(0xffffffffu, valid && w < 1.0f)) {
                // weights.py:502-507 cosine ramp, only near view borders
                float cw = cosine_ramp(w);
                if (__any_sync(0xffffffffu, valid && w < 0.02f)) {
                  // right at a view border the reference's float32 (cos + 1) / 2 underflows to an
                  // exact 0 (such a view is then ignored): same formula there
                  const float a = __fmul_rn(__fsub_rn(1.0f, w), 3.14159274101257324f);
                  const float ce = __fmul_rn(__fadd_rn(cosf(a), 1.0f), 0.5f);
                  cw = w < 0.02f ? ce : cw;
                }
                w = w < 1.0f ? cw : w;
              }
              w = fminf(fmaxf(w, 0.0f), 1.0f);
              b = (valid && rcin[q] && iz >= 0) ? w : 0.f;
            }
            // acc holds the raw value while only one view has contributed (its normalised
            // weight is exactly 1); the first view is weighted retroactively once a second
            // one arrives.
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
        }
        anymask |= vm;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (!(flags & ITEM_LAST)) continue;

    // ---- finalise (registers only) ----
    if (flags & ITEM_EMPTY) {
      // acc holds zeros
    } else if (MODE == MVS_FUSE_MAX) {
#pragma unroll
      for (int k = 0; k < OUTS; ++k) acc[k] = (anymask >> k) & 1 ? acc[k] : 0.f;
    } else if (MODE == MVS_FUSE_MEAN || (flags & ITEM_SIMPLE)) {
      if (!PARTIAL) {
#pragma unroll
        for (int k = 0; k < OUTS; ++k)
          if (__any_sync(0xffffffffu, den[k] > 1.f))
            acc[k] = den[k] > 1.f ? __fdiv_rn(acc[k], den[k]) : acc[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < OUTS; ++k) {
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
    store_step(orec, acc, den);
  }
}

}  // namespace mvs
