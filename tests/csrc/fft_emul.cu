// CPU execution of the register-FFT index arithmetic (fft_reg.cuh): every
// "thread" is run in turn, phase by phase, against a float64 O(n^2) DFT.
// Build: nvcc -std=c++17 -I multiview_stitcher_b200/csrc tests/csrc/fft_emul.cu -o /tmp/fft_emul
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft_reg.cuh"

using namespace mvs;

template <int M, int S>
struct Runner {
  static void run(std::vector<std::vector<float2>>& regs, std::vector<float2>& sline,
                  const float2* tw) {
    using Sc = FftSched<M>;
    if constexpr (S < Sc::NST) {
      for (int t = 0; t < Sc::T; ++t) FftStage<M, S>::compute(regs[t].data(), t, tw);
      if constexpr (S + 1 < Sc::NST) {
        for (auto& s : sline) s = make_float2(NAN, NAN);
        for (int t = 0; t < Sc::T; ++t) FftStage<M, S>::scatter(regs[t].data(), t, sline.data());
        for (int t = 0; t < Sc::T; ++t) fft_gather<M>(regs[t].data(), t, sline.data());
        Runner<M, S + 1>::run(regs, sline, tw);
      }
    }
  }
};

template <int M>
double check() {
  using Sc = FftSched<M>;
  std::vector<float2> tw(M), x(M);
  for (int k = 0; k < M; ++k) {
    double a = -2.0 * M_PI * k / M;
    tw[k] = make_float2((float)cos(a), (float)sin(a));
    x[k] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
  }
  std::vector<std::vector<float2>> regs(Sc::T, std::vector<float2>(Sc::E));
  for (int t = 0; t < Sc::T; ++t)
    for (int q = 0; q < Sc::E; ++q) regs[t][q] = x[t + q * Sc::T];
  std::vector<float2> sline(Sc::PADM + 16);
  Runner<M, 0>::run(regs, sline, tw.data());
  double err = 0, nrm = 0;
  for (int k = 0; k < M; ++k) {
    double re = 0, im = 0;
    for (int n = 0; n < M; ++n) {
      double a = -2.0 * M_PI * (double)((long long)n * k % M) / M;
      re += x[n].x * cos(a) - x[n].y * sin(a);
      im += x[n].x * sin(a) + x[n].y * cos(a);
    }
    const float2 got = regs[k % Sc::T][k / Sc::T];
    err = fmax(err, hypot(got.x - re, got.y - im));
    nrm = fmax(nrm, hypot(re, im));
  }
  printf("M=%5d E=%2d T=%4d stages=%d  max err / max |X| = %.3e\n", M, Sc::E, Sc::T, Sc::NST, err / nrm);
  return err / nrm;
}

int main() {
  double worst = 0;
  worst = fmax(worst, check<1>());
  worst = fmax(worst, check<2>());
  worst = fmax(worst, check<4>());
  worst = fmax(worst, check<8>());
  worst = fmax(worst, check<16>());
  worst = fmax(worst, check<32>());
  worst = fmax(worst, check<64>());
  worst = fmax(worst, check<128>());
  worst = fmax(worst, check<256>());
  worst = fmax(worst, check<512>());
  worst = fmax(worst, check<640>());
  worst = fmax(worst, check<1024>());
  worst = fmax(worst, check<2048>());
  worst = fmax(worst, check<4096>());
  worst = fmax(worst, check<8192>());
  if (!(worst < 5e-6)) { printf("FAIL\n"); return 1; }
  printf("OK\n");
  return 0;
}
