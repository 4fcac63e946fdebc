// CPU emulation of the register-blocked FFT engine (fast_fft.cuh): the same
// __host__ __device__ pass functions, run for all threads pass by pass (a pass
// boundary = the kernel's barrier).  Checks circular convolution through
// fwd passes -> spectrum multiply (position order) -> mirrored inverse passes,
// for every supported size and both precisions, against a direct evaluation.
#include "../../fftconv_b200/csrc/fast_fft.cuh"

#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

using namespace fc::fast;

template <int LOG2N, int J, typename T> struct HostFwd {
  static void run(std::vector<cx<T>> &buf, const cx<scalar_t<T>> *W) {
    if constexpr (J <= Plan<LOG2N>::P - 2) {
      for (int t = 0; t < Plan<LOG2N>::THREADS; ++t) fwd_mid<LOG2N, J>(t, buf.data(), W);
      HostFwd<LOG2N, J + 1, T>::run(buf, W);
    }
  }
};
template <int LOG2N, int J, typename T> struct HostInv {
  static void run(std::vector<cx<T>> &buf, const cx<scalar_t<T>> *W) {
    if constexpr (J >= 1) {
      for (int t = 0; t < Plan<LOG2N>::THREADS; ++t) inv_mid<LOG2N, J>(t, buf.data(), W);
      HostInv<LOG2N, J - 1, T>::run(buf, W);
    }
  }
};

template <int LOG2N, typename T> double run_case() {
  using P = Plan<LOG2N>;
  constexpr int N = P::N;
  const long double pi = 3.14159265358979323846264338327950288L;
  std::mt19937 rng(LOG2N * 7 + sizeof(T));
  std::normal_distribution<double> nd(0, 1);
  const int K = 37;
  std::vector<std::complex<double>> x(N);
  std::vector<double> h(K);
  for (auto &v : x) v = {nd(rng), nd(rng)};
  for (auto &v : h) v = nd(rng);
  std::vector<cx<T>> W(P::TW_ELEMS), H(N);
  fill_pass_twiddles<LOG2N, T>(W.data(), [&](long long num, long long den) {
    const long double a = 2 * pi * (long double)(num % den) / (long double)den;
    return cx<T>{(T)cosl(a), (T)-sinl(a)};
  });
  for (int idx = 0; idx < N; ++idx) {
    const int pos = idx;
    const unsigned k = freq_of_pos(pos_of_table_index(idx, LOG2N), LOG2N);
    long double re = 0, im = 0;
    for (int j = 0; j < K; ++j) {
      const long double a = 2 * pi * ((long long)j * k % N) / N;
      re += h[j] * cosl(a);
      im -= h[j] * sinl(a);
    }
    H[pos] = {(T)(re / N), (T)(im / N)};
  }
  std::vector<cx<T>> buf(P::SMEM_ELEMS);
  std::vector<std::vector<cx<T>>> regs(P::THREADS, std::vector<cx<T>>(16));
  for (int t = 0; t < P::THREADS; ++t)
    for (int s = 0; s < 16; ++s) {
      const int i = pass0_index<LOG2N>(t, s);
      regs[t][s] = {(T)x[i].real(), (T)x[i].imag()};
    }
  for (int t = 0; t < P::THREADS; ++t) fwd_pass0<LOG2N>(t, regs[t].data(), buf.data(), W.data());
  HostFwd<LOG2N, 1, T>::run(buf, W.data());
  for (int t = 0; t < P::THREADS; ++t) last_fused<LOG2N>(t, buf.data(), TableSpectrum<T>{H.data(), P::THREADS});
  HostInv<LOG2N, P::P - 2, T>::run(buf, W.data());
  for (int t = 0; t < P::THREADS; ++t) inv_pass0<LOG2N>(t, regs[t].data(), buf.data(), W.data());
  // compare against direct circular convolution at 4096 sampled outputs (all for small N)
  double num = 0, den = 0;
  std::vector<bool> seen(N, false);
  for (int t = 0; t < P::THREADS; ++t)
    for (int s = 0; s < 16; ++s) {
      const int i = pass0_index<LOG2N>(t, s);
      if (seen[i]) return 1e9; // the slot map must be a bijection
      seen[i] = true;
      if (N > 4096 && (i * 2654435761u >> 20) % (N / 4096) != 0) continue;
      std::complex<double> acc = 0;
      for (int j = 0; j < K; ++j) acc += h[j] * x[(i - j + N) % N];
      const std::complex<double> got(regs[t][s].x, regs[t][s].y);
      num += std::norm(got - acc);
      den += std::norm(acc);
    }
  return std::sqrt(num / den);
}

// packed data type f2: two independent complex transforms in the halves of every value
template <int LOG2N> double run_case_packed() {
  using P = Plan<LOG2N>;
  using fc::f2;
  constexpr int N = P::N;
  const long double pi = 3.14159265358979323846264338327950288L;
  std::mt19937 rng(LOG2N * 11);
  std::normal_distribution<double> nd(0, 1);
  const int K = 41;
  std::vector<std::complex<double>> x[2] = {std::vector<std::complex<double>>(N), std::vector<std::complex<double>>(N)};
  std::vector<double> h(K);
  for (int half = 0; half < 2; ++half)
    for (auto &v : x[half]) v = {nd(rng), nd(rng)};
  for (auto &v : h) v = nd(rng);
  std::vector<cx<float>> W(P::TW_ELEMS), H(N);
  fill_pass_twiddles<LOG2N, float>(W.data(), [&](long long num, long long den) {
    const long double a = 2 * pi * (long double)(num % den) / (long double)den;
    return cx<float>{(float)cosl(a), (float)-sinl(a)};
  });
  for (int idx = 0; idx < N; ++idx) {
    const unsigned k = freq_of_pos(pos_of_table_index(idx, LOG2N), LOG2N);
    long double re = 0, im = 0;
    for (int j = 0; j < K; ++j) {
      const long double a = 2 * pi * ((long long)j * k % N) / N;
      re += h[j] * cosl(a);
      im -= h[j] * sinl(a);
    }
    H[idx] = {(float)(re / N), (float)(im / N)};
  }
  std::vector<cx<f2>> buf(P::SMEM_ELEMS);
  std::vector<std::vector<cx<f2>>> regs(P::THREADS, std::vector<cx<f2>>(16));
  for (int t = 0; t < P::THREADS; ++t)
    for (int s = 0; s < 16; ++s) {
      const int i = pass0_index<LOG2N>(t, s);
      regs[t][s].x = fc::make_f2((float)x[0][i].real(), (float)x[1][i].real());
      regs[t][s].y = fc::make_f2((float)x[0][i].imag(), (float)x[1][i].imag());
    }
  for (int t = 0; t < P::THREADS; ++t) fwd_pass0<LOG2N>(t, regs[t].data(), buf.data(), W.data());
  HostFwd<LOG2N, 1, f2>::run(buf, W.data());
  for (int t = 0; t < P::THREADS; ++t) last_fused<LOG2N>(t, buf.data(), TableSpectrum<float>{H.data(), P::THREADS});
  HostInv<LOG2N, P::P - 2, f2>::run(buf, W.data());
  for (int t = 0; t < P::THREADS; ++t) inv_pass0<LOG2N>(t, regs[t].data(), buf.data(), W.data());
  double num = 0, den = 0;
  for (int t = 0; t < P::THREADS; ++t)
    for (int s = 0; s < 16; ++s) {
      const int i = pass0_index<LOG2N>(t, s);
      if (N > 4096 && (i * 2654435761u >> 20) % (N / 4096) != 0) continue;
      for (int half = 0; half < 2; ++half) {
        std::complex<double> acc = 0;
        for (int j = 0; j < K; ++j) acc += h[j] * x[half][(i - j + N) % N];
        const std::complex<double> got(half ? regs[t][s].x.y : regs[t][s].x.x, half ? regs[t][s].y.y : regs[t][s].y.x);
        num += std::norm(got - acc);
        den += std::norm(acc);
      }
    }
  return std::sqrt(num / den);
}

template <int LOG2N> int check() {
  const double ef = run_case<LOG2N, float>(), ed = run_case<LOG2N, double>(), ep = run_case_packed<LOG2N>();
  std::printf("N=%d  float %.3e  double %.3e  packed f32x2 %.3e\n", 1 << LOG2N, ef, ed, ep);
  return (ef < 2e-6 && ed < 2e-15 * 10 && ep < 2e-6) ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += check<9>();
  bad += check<10>();
  bad += check<11>();
  bad += check<12>();
  bad += check<13>();
  bad += check<14>();
  std::printf(bad ? "FAILED\n" : "ALL PASSED\n");
  return bad;
}
