// Batched 1-D complex FFT lines in shared memory (sm_100a), no cuFFT.
//
// * power-of-two lengths: Stockham autosort, radix-4 stages (+ one radix-2),
//   ping-pong between two shared-memory buffers, twiddles from a per-length
//   table computed in float64 on the host;
// * any other length n: Bluestein chirp-z on the same Stockham kernel with
//   m = 2^k >= 2n-1 (exact length-n DFT -- zero padding to 2^k would change
//   the circular correlation the reference computes, SURVEY.md 7 hard part 1).
//
// A CTA transforms L lines at once so that lines along a strided axis are
// loaded as runs of L adjacent elements.
#pragma once

#include "common.cuh"

namespace mvs {

struct AxisFft {
  int n;          // logical transform length
  int m;          // power-of-two kernel length (== n when n is 2^k)
  int bluestein;  // 1 -> chirp-z
  const float2* tw;     // [m]  exp(-2 pi i k / m)
  const float2* chirp;  // [n]  exp(-i pi k^2 / n)
  const float2* bhat;   // [m]  FFT_m(conj chirp, wrapped) / m
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// L lines of length m (line stride mp) from buffer a, result pointer returned
// (a or b).  sign = -1: forward (e^{-i..}), +1: inverse (unnormalised).
// All threads of the CTA must call; ends with a __syncthreads().
__device__ inline float2* stockham_lines(float2* a, float2* b, int m, int mp, int L,
                                         const float2* __restrict__ tw, int sign) {
  for (int Ns = 1; Ns < m;) {
    const int R = (Ns * 4 <= m) ? 4 : 2;
    const int q = m / R;
    const int qshift = 31 - __clz(q);  // m, q are powers of two
    const int tstep = m / (Ns * R);
    const int total = L * q;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int line = idx >> qshift;
      const int j = idx & (q - 1);
      const int k = j & (Ns - 1);
      const float2* in = a + line * mp;
      float2* out = b + line * mp;
      if (R == 4) {
        float2 v0 = in[j], v1 = in[j + q], v2 = in[j + 2 * q], v3 = in[j + 3 * q];
        if (k) {
          float2 w1 = __ldg(tw + k * tstep), w2 = __ldg(tw + 2 * k * tstep),
                 w3 = __ldg(tw + 3 * k * tstep);
          if (sign > 0) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
          v1 = cmul(v1, w1); v2 = cmul(v2, w2); v3 = cmul(v3, w3);
        }
        float2 t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3), d = csub(v1, v3);
        // (v1 - v3) * exp(sign * i pi/2)
        float2 t3 = sign < 0 ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);
        const int base = (j - k) * 4 + k;
        out[base] = cadd(t0, t2);
        out[base + Ns] = cadd(t1, t3);
        out[base + 2 * Ns] = csub(t0, t2);
        out[base + 3 * Ns] = csub(t1, t3);
      } else {
        float2 v0 = in[j], v1 = in[j + q];
        if (k) {
          float2 w1 = __ldg(tw + k * tstep);
          if (sign > 0) w1.y = -w1.y;
          v1 = cmul(v1, w1);
        }
        const int base = (j - k) * 2 + k;
        out[base] = cadd(v0, v1);
        out[base + Ns] = csub(v0, v1);
      }
    }
    __syncthreads();
    float2* t = a; a = b; b = t;
    Ns *= R;
  }
  return a;
}

// Length-n DFT of L lines already staged in `a` (n valid entries per line,
// entries [n, m) must be zero for Bluestein).  Returns the buffer holding the
// n results per line.  sign as above.
__device__ inline float2* fft_lines(float2* a, float2* b, const AxisFft& ax, int mp, int L,
                                    int sign) {
  if (!ax.bluestein) return stockham_lines(a, b, ax.m, mp, L, ax.tw, sign);
  const int n = ax.n, m = ax.m;
  // a_k = x_k * chirp_s(k)
  for (int idx = threadIdx.x; idx < L * n; idx += blockDim.x) {
    const int line = idx / n, k = idx - line * n;
    float2 c = __ldg(ax.chirp + k);
    if (sign > 0) c.y = -c.y;
    a[line * mp + k] = cmul(a[line * mp + k], c);
  }
  __syncthreads();
  float2* A = stockham_lines(a, b, m, mp, L, ax.tw, -1);
  float2* other = (A == a) ? b : a;
  for (int idx = threadIdx.x; idx < L * m; idx += blockDim.x) {
    const int line = idx / m, k = idx - line * m;
    float2 h = __ldg(ax.bhat + k);
    if (sign > 0) h.y = -h.y;
    A[line * mp + k] = cmul(A[line * mp + k], h);
  }
  __syncthreads();
  float2* Y = stockham_lines(A, other, m, mp, L, ax.tw, +1);
  for (int idx = threadIdx.x; idx < L * n; idx += blockDim.x) {
    const int line = idx / n, k = idx - line * n;
    float2 c = __ldg(ax.chirp + k);
    if (sign > 0) c.y = -c.y;
    Y[line * mp + k] = cmul(Y[line * mp + k], c);
  }
  __syncthreads();
  return Y;
}

// Host side: per-length tables, cached for the life of the process.
// Returns nullptr (and sets the error) on failure.
const AxisFft* get_axis_fft(int n);
int fft_lines_per_cta(const AxisFft& ax, bool contiguous);
size_t fft_smem_bytes(const AxisFft& ax, int L, int* mp_out);

}  // namespace mvs
