// bench_c1.cpp -- BASELINE.json config 1 through the reference's own C++ call: the reference's
// benchmark case (n = 4352, K = 65; 17.2 us per call on an i7 with FFTW,
// /root/reference/README.md:99-102, /root/reference/test/test_script.cpp:29-34) with
// std::vector buffers (pageable host memory) handed to oaconvolve_fftw / convolve_fftw.
// Prints one JSON line: microseconds per call, host pointers in, host pointers out.
#include "fftconv/fftconv.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

template <class F> static double us_per_call(F &&f, int reps) {
  for (int i = 0; i < 200; ++i) f();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) f();
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::micro>(t1 - t0).count() / reps;
}

int main() {
  const std::size_t n = 4352, K = 65;
  std::mt19937 rng(1);
  std::normal_distribution<float> nd;
  std::vector<float> a(n), k(K), full(n + K - 1), same(n);
  for (auto &v : a) v = nd(rng);
  for (auto &v : k) v = nd(rng);
  const std::span<const float> as(a), ks(k);
  try {
    const double oa_full = us_per_call([&] { fftconv::oaconvolve_fftw<float, fftconv::Full>(as, ks, std::span<float>(full)); }, 3000);
    const double oa_same = us_per_call([&] { fftconv::oaconvolve_fftw<float, fftconv::Same>(as, ks, std::span<float>(same)); }, 3000);
    const double cv_full = us_per_call([&] { fftconv::convolve_fftw<float, fftconv::Full>(as, ks, std::span<float>(full)); }, 3000);
    // parity against the definition, double accumulation
    double num = 0, den = 0;
    for (std::size_t o = 0; o < n + K - 1; ++o) {
      double s = 0;
      for (std::size_t j = 0; j < K; ++j)
        if (o >= j && o - j < n) s += (double)k[j] * (double)a[o - j];
      num += (s - full[o]) * (s - full[o]);
      den += s * s;
    }
    std::printf("{\"oaconvolve_full_us\": %.2f, \"oaconvolve_same_us\": %.2f, \"convolve_full_us\": %.2f, "
                "\"rel_l2_vs_definition\": %.3e, \"buffers\": \"std::vector (pageable)\", \"n\": 4352, \"k\": 65}\n",
                oa_full, oa_same, cv_full, std::sqrt(num / den));
  } catch (const std::exception &e) {
    std::printf("{\"error\": \"%s\"}\n", e.what());
    return 1;
  }
  return 0;
}
