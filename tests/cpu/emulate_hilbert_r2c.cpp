// CPU emulation of the real-input Hilbert engine (hilbert_r2c_core.cuh): the same __host__ __device__
// pass functions the CUDA kernel runs, all 128 threads of a line up to each barrier, on NaN-poisoned
// planes, against a float64 Hilbert transform computed from the definition (DFT, -j sgn f, inverse DFT).
//
// usage: emulate_hilbert_r2c [lines]
#include "../../fftconv_b200/csrc/hilbert_r2c_tables.h"

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

using namespace fc;
using namespace fc::hr2c;

// float64 reference: radix-2 FFT of the 8192-point line
static void fft(std::vector<std::complex<double>> &a, int sign) {
  const size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const double ang = sign * 2.0 * M_PI / (double)len;
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; ++k) {
        const std::complex<double> w(std::cos(ang * (double)k), std::sin(ang * (double)k));
        const auto u = a[i + k], v = a[i + k + len / 2] * w;
        a[i + k] = u + v;
        a[i + k + len / 2] = u - v;
      }
  }
}

int main(int argc, char **argv) {
  const int lines = argc > 1 ? atoi(argv[1]) : 3;
  std::vector<Quad> tw1(kTw1Quads), tw2(kTw2Quads);
  std::vector<Pair> mir(kMirPairs);
  fill_tables(tw1.data(), tw2.data(), mir.data());
  std::mt19937 rng(99);
  std::normal_distribution<float> nd(0.f, 1.f);
  double worst = 0;
  // every word of the planes must be written before it is read, and every point owned by exactly one thread
  {
    std::vector<int> owner(kGroupWords / 2, 0);
    for (int d0 = 0; d0 < 16; ++d0)
      for (int d1 = 0; d1 < 16; ++d1)
        for (int d2 = 0; d2 < 16; ++d2) ++owner[alpha(d0) + beta(d1) + d2];
    for (int v : owner)
      if (v > 1) { std::printf("layout collision\n"); return 1; }
    std::vector<int> seen(4096, 0);
    for (int s = 0; s < kThreads; ++s) {
      const P3Map m = p3_map(s);
      ++seen[m.k0a * 16 + m.k1a];
      ++seen[m.k0b * 16 + m.k1b];
    }
    for (int i = 0; i < 256; ++i)
      if (seen[i] != 1) { std::printf("P3 map: butterfly %d owned %d times\n", i, seen[i]); return 1; }
  }
  for (int line = 0; line < lines; ++line) {
    std::vector<float> x_store(kRow + 4);
    float *x = x_store.data();
    while (reinterpret_cast<unsigned long long>(x) & 15ULL) ++x;
    for (int i = 0; i < kRow; ++i) x[i] = nd(rng) + (line == 1 ? 0.5f : 0.f);
    std::vector<float> planes(kGroupWords, NAN);
    std::vector<Regs> regs(kThreads);
    std::vector<ThreadMaps> maps(kThreads);
    for (int t = 0; t < kThreads; ++t) maps[t] = thread_maps(t, tw1.data(), tw2.data(), mir.data());
    for (int t = 0; t < kThreads; ++t) {
      load_line(t, regs[t], x);
      dft16<-1>(regs[t]);
      fwd_a_pre(regs[t], planes.data(), maps[t]);
    }
    // between the two group barriers a warp only synchronises with itself (__syncwarp): run warp after warp,
    // so that a warp that needed another warp's P2 output would read stale data and fail the comparison
    for (int w = 0; w < kThreads / 32; ++w) {
      for (int t = 32 * w; t < 32 * w + 32; ++t) {
        fwd_a_post(regs[t], planes.data(), maps[t]);
        dft16<-1>(regs[t]);
        fwd_b_pre(regs[t], planes.data(), maps[t]);
      }
      for (int t = 32 * w; t < 32 * w + 32; ++t) {
        fwd_b_post(regs[t], planes.data(), maps[t]);
        dft16<-1>(regs[t]);
        fwd_c(t, regs[t], maps[t], mir.data());
        dft16<+1>(regs[t]);
        inv_a_pre(regs[t], planes.data(), maps[t]);
      }
      for (int t = 32 * w; t < 32 * w + 32; ++t) {
        inv_a_post(regs[t], planes.data(), maps[t]);
        dft16<+1>(regs[t]);
        inv_b_pre(regs[t], planes.data(), maps[t]);
      }
    }
    std::vector<float> h(kRow, NAN);
    for (int t = 0; t < kThreads; ++t) {
      inv_b_post(regs[t], planes.data(), maps[t]);
      dft16<+1>(regs[t]);
      for (int a = 0; a < 16; ++a) { // slot16(a): g[256 a + 2 t + half]
        const int s = slot16(a), m = 256 * a + 2 * t;
        h[2 * m] = regs[t].re[s].x;
        h[2 * m + 1] = regs[t].im[s].x;
        h[2 * m + 2] = regs[t].re[s].y;
        h[2 * m + 3] = regs[t].im[s].y;
      }
    }
    std::vector<std::complex<double>> X(kRow);
    for (int i = 0; i < kRow; ++i) X[i] = x[i];
    fft(X, -1);
    for (int k = 0; k < kRow; ++k) {
      if (k == 0 || k == kRow / 2) X[k] = 0;
      else if (k < kRow / 2) X[k] *= std::complex<double>(0, -1);
      else X[k] *= std::complex<double>(0, 1);
    }
    fft(X, +1);
    double num = 0, den = 0;
    for (int i = 0; i < kRow; ++i) {
      const double want = X[i].real() / kRow, d = (double)h[i] - want;
      num += d * d;
      den += want * want;
    }
    const double err = std::sqrt(num / den);
    if (!(err < 1e-6)) { std::printf("line %d: rel l2 %.3e\n", line, err); return 1; }
    worst = std::max(worst, err);
  }
  std::printf("%.3e ALL PASSED\n", worst);
  return 0;
}
