// One pass of the separable n-D transform of phase correlation: L lines of one
// axis per CTA, each line transformed in registers (fft_reg.cuh), with the
// element-wise neighbours of the reference's pipeline fused into the passes:
//   * LOAD_REAL  first forward pass packs the two rescaled real crops of a pair
//                into one complex line (Z = r0 + i r1, NaN -> 0);
//   * PAIRED     last forward pass: a CTA owns lines q and mirror(q) together, so
//                the cross-power spectrum  Q_k = s P_k + i P_k / max(|P_k|, 100 eps),
//                P = F conj(M), F = (Z_k + conj Z_-k)/2, M = (Z_k - conj Z_-k)/(2i)
//                (skimage phase_cross_correlation, both normalisations packed,
//                registration.py:416-431) is formed in the epilogue from shared
//                memory (for the inverse transform), next to the plain P that the
//                upsampled DFT refines on (overwriting Z in place);
//   * ARGMAX     last inverse pass: the correlation surfaces are only needed for
//                their first maxima, so the pass reduces (|value|, ~index) keys
//                and stores nothing.
// The per-thread pieces are __host__ __device__ (tests/csrc/fft_emul.cu runs the
// same code thread by thread on the CPU).
#pragma once

#include <cmath>
#include <cstring>

#include "fft_reg.cuh"

namespace mvs {

struct FftPassArgs {
  const float2* src;       // complex input (unused with load_real)
  float2* dst;             // complex output, may equal src; nullptr: nothing stored
  float2* dst2;            // paired: the plain cross-power spectrum P (may equal src)
  const float* re;         // load_real: real sources (NaN -> 0)
  const float* im;
  long long outer, inner;  // lines = outer * inner; element k of line (o, in) at (o n + k) inner + in
  long long batch_stride;  // elements between pairs (blockIdx.y)
  int n;                   // logical transform length
  int L;                   // lines per CTA
  int line_stride;         // float2 elements between two line buffers in shared memory
  int contig;              // thread mapping: 1 = line-major (contiguous axis), 0 = lines fastest
  int sign;                // -1 forward, +1 inverse (unnormalised)
  int load_real;
  int paired;              // cross-power epilogue (forward only, outer == 1, L even)
  int argmax;              // reduce keys (contiguous axis only)
  int n2, n1p, items_x, xblocks;  // paired: innermost length, inner / n2, n2/2 + 1, CTAs along x
  const float* cp_scales;  // [pair] s: puts max |s P| at 2^12 next to the unit-modulus normalised spectrum
  unsigned long long* keys;  // [pair][2]: slot 0 = |Re| (normalization None), 1 = |Im| ("phase")
  const float2* tw;
  const float2* chirp;
  const float2* bhat;
};

// order-preserving 32-bit image of |v| (NaN ranks above every number, like numpy's argmax)
MVS_HD unsigned fft_abs_bits(float v) {
#ifdef __CUDA_ARCH__
  const unsigned u = __float_as_uint(v) & 0x7fffffffu;
#else
  unsigned u;
  memcpy(&u, &v, 4);
  u &= 0x7fffffffu;
#endif
  return u;
}
// larger |v| wins, then the LOWER flat index (first maximum in C order)
MVS_HD unsigned long long fft_key(unsigned abs_bits, unsigned idx) {
  return ((unsigned long long)abs_bits << 32) | (unsigned long long)(0xffffffffu - idx);
}

template <int M, bool BLUE>
struct PassThread {
  using Sc = FftSched<M>;
  static constexpr int E = Sc::E, T = Sc::T;
  int t, l;
  bool valid;
  long long base;   // element k of this thread's line lives at base + k * inner
  long long flat0;  // argmax: flat index of element 0 of the line within the pair
  float2 v[E];

  MVS_HD void init(const FftPassArgs& P, int tid, long long bx, long long by) {
    if (P.contig) { l = tid / T; t = tid - l * T; }
    else { t = tid / P.L; l = tid - t * P.L; }
    const long long boff = by * P.batch_stride;
    if (P.paired) {
      const int H = P.L >> 1;
      const long long yb = bx / P.xblocks;
      const int xb = (int)(bx - yb * P.xblocks);
      const int half = l / H, i = l - half * H;
      const int xi = xb * H + i;
      int x = xi, y = (int)yb;
      // columns that mirror onto themselves (x = 0, x = n2/2) pair the rows y' and
      // -y': only the smaller row index owns the pair, so every line is read and
      // written by exactly one CTA (P may then overwrite Z in place)
      const bool self_col = xi == 0 || 2 * xi == P.n2;
      valid = xi < P.items_x && !(self_col && y != 0 && P.n1p - y < y);
      if (half) { x = xi ? P.n2 - xi : 0; y = y ? P.n1p - y : 0; }
      base = valid ? boff + (long long)y * P.n2 + x : boff;
      flat0 = 0;
    } else {
      const long long q = bx * P.L + l;
      valid = q < P.outer * P.inner;
      const long long o = q / P.inner, in = q - o * P.inner;
      base = valid ? boff + o * (long long)P.n * P.inner + in : boff;
      flat0 = o * (long long)P.n * P.inner + in;
    }
  }

  // All global loads of the thread are issued back to back (clamped addresses, no
  // branches) before any of them is consumed: E independent requests in flight.
  MVS_HD void load(const FftPassArgs& P) {
    const float sgn = P.sign > 0 ? -1.f : 1.f;  // inverse = conj(forward(conj(x)))
    if (P.load_real) {
      float a[E], b[E];
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int k = t + q * T;
        const long long g = base + (long long)(k < P.n ? k : 0) * P.inner;
        a[q] = MVS_LDG(P.re + g);
        b[q] = MVS_LDG(P.im + g);
      }
#pragma unroll
      for (int q = 0; q < E; ++q) v[q] = make_float2(a[q] != a[q] ? 0.f : a[q], b[q] != b[q] ? 0.f : b[q]);
    } else {
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int k = t + q * T;
        v[q] = P.src[base + (long long)(k < P.n ? k : 0) * P.inner];
      }
    }
    if (BLUE) {
      float2 c[E];
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int k = t + q * T;
        c[q] = MVS_LDG(P.chirp + (k < P.n ? k : 0));
      }
#pragma unroll
      for (int q = 0; q < E; ++q) v[q] = cmul(make_float2(v[q].x, sgn * v[q].y), c[q]);
    } else {
#pragma unroll
      for (int q = 0; q < E; ++q) v[q].y *= sgn;
    }
#pragma unroll
    for (int q = 0; q < E; ++q)
      if (!valid || t + q * T >= P.n) v[q] = make_float2(0.f, 0.f);
  }

  // Bluestein: between the two power-of-two transforms (registers only)
  MVS_HD void mid(const FftPassArgs& P) {
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int k = t + q * T;
      const float2 y = cmul(v[q], MVS_LDG(P.bhat + k));
      v[q] = make_float2(y.x, -y.y);  // conj: the second transform runs forward
    }
  }

  MVS_HD void post(const FftPassArgs& P) {
    const float sgn = P.sign > 0 ? -1.f : 1.f;
    if (BLUE) {
      float2 c[E];
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int k = t + q * T;
        c[q] = MVS_LDG(P.chirp + (k < P.n ? k : 0));
      }
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const float2 y = cmul(make_float2(v[q].x, -v[q].y), c[q]);
        v[q] = (t + q * T < P.n) ? make_float2(y.x, sgn * y.y) : make_float2(0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int q = 0; q < E; ++q) v[q].y *= sgn;
    }
  }

  // paired epilogue, step 1: publish the line's spectrum
  MVS_HD void publish(float2* sline) const {
#pragma unroll
    for (int q = 0; q < E; ++q) sline[fft_pad(t + q * T)] = v[q];
  }
  // step 2: Q_k from Z_k (registers) and Z_-k (the partner line, reversed)
  MVS_HD void cross_power(const FftPassArgs& P, const float2* partner, float cps) {
    const float tiny = 100.0f * 1.1920929e-07f;  // 100 * eps(float32)
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int k = t + q * T;
      if (k < P.n) {
        const float2 a = v[q];
        const float2 bm = partner[fft_pad(k ? P.n - k : 0)];
        const float2 bc = make_float2(bm.x, -bm.y);                   // conj Z_-k
        const float2 F = make_float2(0.5f * (a.x + bc.x), 0.5f * (a.y + bc.y));
        const float2 d = make_float2(a.x - bc.x, a.y - bc.y);
        const float2 Mm = make_float2(0.5f * d.y, -0.5f * d.x);       // d / (2i)
        const float2 Pw = make_float2(F.x * Mm.x + F.y * Mm.y, F.y * Mm.x - F.x * Mm.y);
        // |P| = |F| |M| (the squares of P itself could overflow for N ~ 2^31)
        const float mag = fmaxf(sqrtf(F.x * F.x + F.y * F.y) * sqrtf(Mm.x * Mm.x + Mm.y * Mm.y), tiny);
        const float2 Pn = make_float2(Pw.x / mag, Pw.y / mag);
        v[q] = make_float2(cps * Pw.x - Pn.y, cps * Pw.y + Pn.x);
        if (valid && P.dst2) P.dst2[base + (long long)k * P.inner] = Pw;
      }
    }
  }

  MVS_HD void store(const FftPassArgs& P) const {
    if (!valid || !P.dst) return;
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int k = t + q * T;
      if (k < P.n) P.dst[base + (long long)k * P.inner] = v[q];
    }
  }

  // this thread's best (|value| bits, flat index) per surface: x = Re, y = Im
  MVS_HD void best(const FftPassArgs& P, unsigned* bits, unsigned* idx) const {
    bits[0] = bits[1] = 0u;
    idx[0] = idx[1] = 0xffffffffu;
    if (!valid) return;
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int k = t + q * T;
      if (k < P.n) {
        const unsigned i = (unsigned)(flat0 + (long long)k * P.inner);
        const unsigned bx = fft_abs_bits(v[q].x), by = fft_abs_bits(v[q].y);
        if (bx > bits[0] || (bx == bits[0] && i < idx[0])) { bits[0] = bx; idx[0] = i; }
        if (by > bits[1] || (by == bits[1] && i < idx[1])) { bits[1] = by; idx[1] = i; }
      }
    }
  }
  MVS_HD void keys(const FftPassArgs& P, unsigned long long& k0, unsigned long long& k1) const {
    unsigned bits[2], idx[2];
    best(P, bits, idx);
    k0 = idx[0] == 0xffffffffu ? 0ull : fft_key(bits[0], idx[0]);
    k1 = idx[1] == 0xffffffffu ? 0ull : fft_key(bits[1], idx[1]);
  }
};

#ifdef __CUDACC__
// CTAs have at most 512 threads (pass_geom); the 20-values-per-thread 640-point kernel
// runs at most 8 lines x 32 threads and gets the registers that frees (no spills; its
// register budget / occupancy is untuned -- a second launch-bound argument changes the
// code generated for every other length, so that tuning needs its own kernel entry)
// MVS_FFT_MINB (experiment): minimum CTAs per SM for the plain passes.  Measured with 2 (64 registers, a few
// spilled bytes): power-of-two crops 4 % faster, C2's stage 3 % slower.  Left undefined by default: even an
// explicit 1 changes the register allocation of every length (76 -> 113 for 2048 points).
#ifdef MVS_FFT_MINB
#define MVS_FFT_BOUNDS(M, BLUE) __launch_bounds__((M) == 640 ? 256 : 512, (!(BLUE) && (M) != 640) ? MVS_FFT_MINB : 1)
#else
#define MVS_FFT_BOUNDS(M, BLUE) __launch_bounds__((M) == 640 ? 256 : 512)
#endif
template <int M, bool BLUE>
__global__ void MVS_FFT_BOUNDS(M, BLUE) fft_reg_pass_kernel(const FftPassArgs P) {
  extern __shared__ float2 fft_smem[];
  __shared__ unsigned long long s_keys[2][16];
  PassThread<M, BLUE> th;
  th.init(P, threadIdx.x, blockIdx.x, blockIdx.y);
  float2* sline = fft_smem + th.l * P.line_stride;
  th.load(P);
  fft_line_reg<M>(th.v, th.t, sline, P.tw);
  if (BLUE) {
    th.mid(P);
    fft_line_reg<M>(th.v, th.t, sline, P.tw);
  }
  th.post(P);
  if (P.paired) {
    __syncthreads();
    th.publish(sline);
    __syncthreads();
    const int lp = th.l + (th.l < (P.L >> 1) ? (P.L >> 1) : -(P.L >> 1));
    th.cross_power(P, fft_smem + lp * P.line_stride, P.cp_scales[blockIdx.y]);
  }
  th.store(P);
  if (P.argmax) {
    unsigned bits[2], idx[2];
    th.best(P, bits, idx);
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      // warp: largest |value|, then the lowest index among the lanes holding it
      const unsigned m = __reduce_max_sync(0xffffffffu, bits[s]);
      const unsigned i = __reduce_min_sync(0xffffffffu, bits[s] == m ? idx[s] : 0xffffffffu);
      if ((threadIdx.x & 31) == 0)
        s_keys[s][threadIdx.x >> 5] = i == 0xffffffffu ? 0ull : fft_key(m, i);
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      const int nw = (blockDim.x + 31) >> 5;
      unsigned long long m = 0;
      for (int i = 0; i < nw; ++i) {
        const unsigned long long c = s_keys[threadIdx.x][i];
        m = c > m ? c : m;
      }
      if (m) atomicMax(P.keys + 2 * blockIdx.y + threadIdx.x, m);
    }
  }
}
#endif

}  // namespace mvs
