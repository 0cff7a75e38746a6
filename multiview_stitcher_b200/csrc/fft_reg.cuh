// Register-resident batched 1-D complex FFT lines (sm_100a), no cuFFT.
//
// A line of M = 2^p points is owned by T = M/E threads, E = min(M, 16) values
// per thread: thread t holds the elements {t + q T, q < E} in registers before
// AND after every stage, so
//   * the first stage reads global memory straight into registers and the last
//     stage's results leave from registers (coalesced along t),
//   * a Stockham stage of radix R <= 16 is E/R register butterflies,
//   * only the autosort permutation between two stages goes through (padded,
//     conflict-free) shared memory: ceil(p/4) - 1 round trips per transform
//     instead of one per radix-4 stage.
// One smooth non-power-of-two length is built on the same scheme: M = 640 = 20*4*4*2
// with E = 20 values per thread and T = 32 threads (one warp) per line -- the
// Bluestein length for 257 <= n <= 320 (C2's 307-px overlap), 0.58x the flops of
// the 1024-point transform it replaces (enabled with -DMVS_BLUESTEIN_SMOOTH).
// Any other length n runs Bluestein's chirp-z on the same core with
// m = 2^k >= 2n-1 (exact length-n DFT -- zero padding would change the circular
// correlation the reference computes, SURVEY.md 7 hard part 1); the product
// with the chirp spectrum happens in registers because the output layout of the
// forward transform is the input layout of the inverse one.
//
// The per-thread pieces are __host__ __device__ so that tests/csrc/fft_emul.cu
// can execute the very same index arithmetic thread by thread on the CPU.
#pragma once

#include <cuda_runtime.h>

namespace mvs {

#define MVS_HD __host__ __device__ __forceinline__

#ifdef __CUDA_ARCH__
#define MVS_LDG(p) __ldg(p)
#else
#define MVS_LDG(p) (*(p))
#endif

struct AxisFft {
  int n;          // logical transform length
  int m;          // power-of-two kernel length (== n when n is 2^k)
  int bluestein;  // 1 -> chirp-z
  const float2* tw;     // [m]  exp(-2 pi i k / m)
  const float2* chirp;  // [n]  exp(-i pi k^2 / n)
  const float2* bhat;   // [m]  FFT_m(conj chirp, wrapped) / m
};

MVS_HD float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
MVS_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
MVS_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
MVS_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
MVS_HD float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// ---- forward DFTs of R register values, natural order in and out -------------

template <int R>
struct DftReg;

template <>
struct DftReg<1> {
  static MVS_HD void run(float2*) {}
};

template <>
struct DftReg<2> {
  static MVS_HD void run(float2* x) {
    const float2 a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
  }
};

MVS_HD void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = cmul_mi(csub(x1, x3));
  x0 = cadd(t0, t2);
  x1 = cadd(t1, t3);
  x2 = csub(t0, t2);
  x3 = csub(t1, t3);
}

template <>
struct DftReg<4> {
  static MVS_HD void run(float2* x) { dft4(x[0], x[1], x[2], x[3]); }
};

// n = i + 2m, k = a + 4c:  X[a + 4c] = y0[a] + (-1)^c W8^a y1[a],  y_i = DFT4_m x[i + 2m]
template <>
struct DftReg<8> {
  static MVS_HD void run(float2* x) {
    const float h = 0.70710678118654752440f;
    dft4(x[0], x[2], x[4], x[6]);  // y0[a] in x[2a]
    dft4(x[1], x[3], x[5], x[7]);  // y1[a] in x[2a + 1]
    const float2 y0 = x[1];
    const float2 y1 = cmul(x[3], make_float2(h, -h));
    const float2 y2 = cmul_mi(x[5]);
    const float2 y3 = cmul(x[7], make_float2(-h, -h));
    const float2 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
    x[0] = cadd(e0, y0); x[4] = csub(e0, y0);
    x[1] = cadd(e1, y1); x[5] = csub(e1, y1);
    x[2] = cadd(e2, y2); x[6] = csub(e2, y2);
    x[3] = cadd(e3, y3); x[7] = csub(e3, y3);
  }
};

// n = i + 4m, k = a + 4c:  X[a + 4c] = DFT4_i ( W16^(i a) * DFT4_m x[i + 4m] )
template <>
struct DftReg<16> {
  static MVS_HD void run(float2* x) {
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
    const float h = 0.70710678118654752440f;
#pragma unroll
    for (int i = 0; i < 4; ++i) dft4(x[i], x[i + 4], x[i + 8], x[i + 12]);  // y_i[a] in x[i + 4a]
    // twiddles W16^(i a), W16^k = (cos(pi k / 8), -sin(pi k / 8))
    x[5] = cmul(x[5], make_float2(c1, -s1));     // i=1,a=1 : W^1
    x[9] = cmul(x[9], make_float2(h, -h));       // i=1,a=2 : W^2
    x[13] = cmul(x[13], make_float2(s1, -c1));   // i=1,a=3 : W^3
    x[6] = cmul(x[6], make_float2(h, -h));       // i=2,a=1 : W^2
    x[10] = cmul_mi(x[10]);                      // i=2,a=2 : W^4
    x[14] = cmul(x[14], make_float2(-h, -h));    // i=2,a=3 : W^6
    x[7] = cmul(x[7], make_float2(s1, -c1));     // i=3,a=1 : W^3
    x[11] = cmul(x[11], make_float2(-h, -h));    // i=3,a=2 : W^6
    x[15] = cmul(x[15], make_float2(-c1, s1));   // i=3,a=3 : W^9
#pragma unroll
    for (int a = 0; a < 4; ++a) dft4(x[4 * a], x[4 * a + 1], x[4 * a + 2], x[4 * a + 3]);
    // now X[a + 4c] sits in x[4a + c]: transpose the 4x4 register tile
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = a + 1; c < 4; ++c) {
        const float2 tmp = x[4 * a + c];
        x[4 * a + c] = x[4 * c + a];
        x[4 * c + a] = tmp;
      }
  }
};

// 5-point DFT (forward), natural order in and out
MVS_HD void dft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;  // cos(2 pi/5), cos(4 pi/5)
  const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;   // sin(2 pi/5), sin(4 pi/5)
  const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  const float2 m1 = make_float2(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
  const float2 m2 = make_float2(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
  const float2 u1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
  const float2 u2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
  x0 = make_float2(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
  const float2 r1 = cmul_mi(u1), r2 = cmul_mi(u2);  // -i u
  x1 = cadd(m1, r1);
  x4 = csub(m1, r1);
  x2 = cadd(m2, r2);
  x3 = csub(m2, r2);
}

template <>
struct DftReg<5> {
  static MVS_HD void run(float2* x) { dft5(x[0], x[1], x[2], x[3], x[4]); }
};

// n = 5 n1 + n2, k = k1 + 4 k2:  X[k1 + 4 k2] = DFT5_n2 ( W20^(n2 k1) * DFT4_n1 x[5 n1 + n2] )
template <>
struct DftReg<20> {
  static MVS_HD void run(float2* x) {
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) dft4(x[n2], x[5 + n2], x[10 + n2], x[15 + n2]);  // y[n2][k1] in x[5 k1 + n2]
    // W20^e = (cos(pi e / 10), -sin(pi e / 10))
    const float c1 = 0.95105651629515357212f, s1 = 0.30901699437494742410f;  // e = 1
    const float c2 = 0.80901699437494742410f, s2 = 0.58778525229247312917f;  // e = 2
    const float c3 = s2, s3 = c2;                                            // e = 3
    const float c4 = s1, s4 = c1;                                            // e = 4
    // k1 = 1: e = n2
    x[6] = cmul(x[6], make_float2(c1, -s1));
    x[7] = cmul(x[7], make_float2(c2, -s2));
    x[8] = cmul(x[8], make_float2(c3, -s3));
    x[9] = cmul(x[9], make_float2(c4, -s4));
    // k1 = 2: e = 2 n2
    x[11] = cmul(x[11], make_float2(c2, -s2));
    x[12] = cmul(x[12], make_float2(c4, -s4));
    x[13] = cmul(x[13], make_float2(-c4, -s4));   // e = 6
    x[14] = cmul(x[14], make_float2(-c2, -s2));   // e = 8
    // k1 = 3: e = 3 n2
    x[16] = cmul(x[16], make_float2(c3, -s3));
    x[17] = cmul(x[17], make_float2(-c4, -s4));   // e = 6
    x[18] = cmul(x[18], make_float2(-c1, -s1));   // e = 9
    x[19] = cmul(x[19], make_float2(-c2, s2));    // e = 12
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft5(x[5 * k1], x[5 * k1 + 1], x[5 * k1 + 2], x[5 * k1 + 3], x[5 * k1 + 4]);
    // X[k1 + 4 k2] sits in x[5 k1 + k2]: reorder to natural order
    float2 y[20];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
      for (int k2 = 0; k2 < 5; ++k2) y[k1 + 4 * k2] = x[5 * k1 + k2];
#pragma unroll
    for (int i = 0; i < 20; ++i) x[i] = y[i];
  }
};

// ---- stage schedule ------------------------------------------------------------

template <int M>
struct FftSched {
  static constexpr int log2i(int v) { return v <= 1 ? 0 : 1 + log2i(v >> 1); }
  static constexpr int P = log2i(M);
  static constexpr int E = M < 16 ? M : 16;   // values per thread
  static constexpr int T = M / E;             // threads per line
  static constexpr int NST = (P + 3) / 4;     // stages (radix 16 ... 16, remainder last)
  static constexpr int radix(int s) { return s < NST - 1 ? 16 : (1 << (P - 4 * (NST - 1))); }
  static constexpr int ns(int s) { return s == 0 ? 1 : 16 * ns(s - 1); }  // product of earlier radices
  static constexpr int PADM = M + (M >> 4);   // padded line length in shared memory
};

// 640 = 20 * 4 * 4 * 2: one warp per line, 20 values per thread
template <>
struct FftSched<640> {
  static constexpr int E = 20;
  static constexpr int T = 32;
  static constexpr int NST = 4;
  static constexpr int radix(int s) { return s == 0 ? 20 : (s == 3 ? 2 : 4); }
  static constexpr int ns(int s) { return s == 0 ? 1 : (s == 1 ? 20 : (s == 2 ? 80 : 320)); }
  static constexpr int PADM = 640 + (640 >> 4);
};

// values per thread of the M-point kernel (host-side launch geometry)
constexpr int fft_values_per_thread(int m) { return m == 640 ? 20 : (m < 16 ? m : 16); }

MVS_HD int fft_pad(int i) { return i + (i >> 4); }

// Stage S of the transform of one line, thread t of T.
template <int M, int S>
struct FftStage {
  using Sc = FftSched<M>;
  static constexpr int E = Sc::E, T = Sc::T;
  static constexpr int R = Sc::radix(S);
  static constexpr int B = E / R;       // butterflies per thread
  static constexpr int NS = Sc::ns(S);  // Stockham Ns

  // twiddle + butterflies; afterwards output r of butterfly b is in v[b + r B]
  static MVS_HD void compute(float2* v, int t, const float2* __restrict__ tw) {
#pragma unroll
    for (int b = 0; b < B; ++b) {
      float2 x[R];
#pragma unroll
      for (int r = 0; r < R; ++r) x[r] = v[b + r * B];
      if (NS > 1) {
        const int j = t + b * T;
        const int k = (NS & (NS - 1)) == 0 ? (j & (NS - 1)) : (j % NS);
        constexpr int tstep = M / (NS * R);
#pragma unroll
        for (int r = 1; r < R; ++r) x[r] = cmul(x[r], MVS_LDG(tw + k * r * tstep));
      }
      DftReg<R>::run(x);
#pragma unroll
      for (int r = 0; r < R; ++r) v[b + r * B] = x[r];
    }
  }
  // autosort permutation: results to their Stockham positions in the line buffer.
  // The padded index of (base + r NS) is pad(base) + r * const whenever the stride
  // keeps the low four bits, so the stores use immediate offsets.
  static MVS_HD void scatter(const float2* v, int t, float2* sline) {
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int j = t + b * T;
      const int k = (NS & (NS - 1)) == 0 ? (j & (NS - 1)) : (j % NS);
      const int base = (j - k) * R + k;
      if constexpr (NS % 16 == 0) {
        float2* p = sline + fft_pad(base);
#pragma unroll
        for (int r = 0; r < R; ++r) p[r * (NS + NS / 16)] = v[b + r * B];
      } else if constexpr (NS == 1 && R == 16) {
        float2* p = sline + 17 * j;  // pad(16 j + r) = 17 j + r
#pragma unroll
        for (int r = 0; r < R; ++r) p[r] = v[b + r * B];
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) sline[fft_pad(base + r * NS)] = v[b + r * B];
      }
    }
  }
};

// the register layout every stage starts from: v[q] = line[t + q T]
template <int M>
MVS_HD void fft_gather(float2* v, int t, const float2* sline) {
  using Sc = FftSched<M>;
  if constexpr (Sc::T % 16 == 0) {
    const float2* p = sline + fft_pad(t);
#pragma unroll
    for (int q = 0; q < Sc::E; ++q) v[q] = p[q * (Sc::T + Sc::T / 16)];
  } else {
#pragma unroll
    for (int q = 0; q < Sc::E; ++q) v[q] = sline[fft_pad(t + q * Sc::T)];
  }
}

#ifdef __CUDACC__
// Forward transform (e^{-i...}, unnormalised) of the line held in v; all
// threads of the CTA must call (CTA-wide barriers).
template <int M, int S = 0>
__device__ __forceinline__ void fft_line_reg(float2* v, int t, float2* sline,
                                             const float2* __restrict__ tw) {
  using Sc = FftSched<M>;
  if constexpr (S < Sc::NST) {
    FftStage<M, S>::compute(v, t, tw);
    if constexpr (S + 1 < Sc::NST) {
      __syncthreads();  // earlier readers of the line buffer are done
      FftStage<M, S>::scatter(v, t, sline);
      __syncthreads();
      fft_gather<M>(v, t, sline);
      fft_line_reg<M, S + 1>(v, t, sline, tw);
    }
  }
}
#endif

}  // namespace mvs
