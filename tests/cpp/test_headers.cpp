// C++ drop-in test: the reference's GTest cases for the hot path, restated without
// GTest/Armadillo, compiled against include/fftconv/*.hpp of THIS repo.
//   test/test_fftconv.cpp:19-43   copy_to_padded_buffer
//   test/test_fftconv.cpp:93-124  FFTConvEngine<double>::get_for_ksize(95).oaconvolve<Mode>
//   test/test_fftconv.cpp:127-218 convolve_fftw / oaconvolve_fftw, f32 + f64, Full + Same,
//                                 planner flags accepted
//   test/test_hilbert.cpp:6-24    Hilbert known-answer vector
// Oracle = direct convolution in double (what arma::conv computes); tolerances from
// test/test_common.hpp:9-18 (1e-9 double, 1e-5 float).  Needs a GPU.
#include <fftconv/fftconv.hpp>
#include <fftconv/hilbert.hpp>
#include <fftconv/multi_gpu.hpp>

#include <array>
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <random>
#include <thread>
#include <vector>

using fftconv::ConvMode;

static int g_fail = 0;
#define EXPECT(c)                                                     \
  do {                                                                \
    if (!(c)) {                                                       \
      std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c);         \
      ++g_fail;                                                       \
    }                                                                 \
  } while (0)

template <class T> std::vector<T> randn(size_t n, unsigned seed) {
  std::mt19937 g(seed);
  std::normal_distribution<double> d(0, 1);
  std::vector<T> v(n);
  for (auto &x : v) x = (T)d(g);
  return v;
}
template <class T> std::vector<double> direct(const std::vector<T> &a, const std::vector<T> &k, ConvMode m) {
  const size_t n = a.size(), K = k.size();
  std::vector<double> full(n + K - 1, 0.0);
  for (size_t i = 0; i < n; ++i)
    for (size_t j = 0; j < K; ++j) full[i + j] += (double)a[i] * (double)k[j];
  if (m == ConvMode::Full) return full;
  return std::vector<double>(full.begin() + K / 2, full.begin() + K / 2 + n);
}
template <class T> constexpr double tol() { return std::is_same_v<T, float> ? 1e-5 : 1e-9; }
template <class T> void near(const std::vector<T> &got, const std::vector<double> &want, double scale = 1.0) {
  EXPECT(got.size() == want.size());
  double worst = 0;
  for (size_t i = 0; i < want.size() && i < got.size(); ++i) worst = std::max(worst, std::abs((double)got[i] - want[i]));
  if (!(worst < tol<T>() * scale)) {
    std::printf("  worst abs error %.3e (tol %.1e)\n", worst, tol<T>() * scale);
    ++g_fail;
  }
}

template <class T, ConvMode Mode, class F> void test_conv(F f) {
  for (int size = 4; size < 100; size += 9) {
    auto a = randn<T>(size, 100 + size), k = randn<T>(size, 200 + size);
    std::vector<T> out(Mode == ConvMode::Full ? 2 * size - 1 : size, T(0));
    f(std::span<const T>(a), std::span<const T>(k), std::span<T>(out));
    near<T>(out, direct(a, k, Mode), std::is_same_v<T, float> ? 10.0 : 1.0);
  }
}
template <class T, ConvMode Mode, class F> void test_oaconv(F f) {
  auto k = randn<T>(95, 7);
  for (int size = 200; size < 501; size += 150) {
    auto a = randn<T>(size, 300 + size);
    std::vector<T> out(Mode == ConvMode::Full ? size + 94 : size, T(0));
    f(std::span<const T>(a), std::span<const T>(k), std::span<T>(out));
    near<T>(out, direct(a, k, Mode), std::is_same_v<T, float> ? 10.0 : 1.0);
  }
}

// fftw::Plan<T> (round 2): the reference's 1-D plan kinds over fftconv_cuda_fft_*
template <typename T> void test_fftw_plans(double tol) {
  using Cx = fftw::Complex<T>;
  for (int n : {1, 2, 20, 64, 100, 243}) {
    const int rows = 3;
    auto re = randn<T>(n * rows, 7 * n), im = randn<T>(n * rows, 9 * n + 1);
    std::vector<T> z(2 * n * rows), Z(2 * n * rows), back(2 * n * rows);
    for (int i = 0; i < n * rows; ++i) { z[2 * i] = re[i]; z[2 * i + 1] = im[i]; }
    auto fwd = fftw::Plan<T>::many_dft(1, &n, rows, reinterpret_cast<Cx *>(z.data()), nullptr, 1, n,
                                       reinterpret_cast<Cx *>(Z.data()), nullptr, 1, n, FFTW_FORWARD, FFTW_ESTIMATE);
    fwd.execute();
    double worst = 0, scale = 0;
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < n; ++k) { // direct DFT in double
        double sr = 0, si = 0;
        for (int j = 0; j < n; ++j) {
          const double a = -2.0 * M_PI * (double)((long long)j * k % n) / n, c = std::cos(a), s = std::sin(a);
          sr += re[r * n + j] * c - im[r * n + j] * s;
          si += re[r * n + j] * s + im[r * n + j] * c;
        }
        worst = std::max({worst, std::abs(sr - Z[2 * (r * n + k)]), std::abs(si - Z[2 * (r * n + k) + 1])});
        scale = std::max({scale, std::abs(sr), std::abs(si)});
      }
    EXPECT(worst <= tol * std::max(scale, 1.0) * 8);
    auto bwd = fftw::Plan<T>::dft_1d(n, reinterpret_cast<Cx *>(Z.data()), reinterpret_cast<Cx *>(back.data()), FFTW_BACKWARD,
                                     FFTW_ESTIMATE);
    bwd.execute(); // first row only; unnormalised: back = n z
    for (int i = 0; i < 2 * n; ++i) EXPECT(std::abs(back[i] - n * z[i]) <= tol * n * 8 * std::max(scale, 1.0));
    // r2c -> c2r round trip through the new-array execute functions
    std::vector<T> half(2 * (n / 2 + 1)), rt(n);
    auto r2c = fftw::Plan<T>::dft_r2c_1d(n, re.data(), reinterpret_cast<Cx *>(half.data()), FFTW_ESTIMATE);
    auto c2r = fftw::Plan<T>::dft_c2r_1d(n, reinterpret_cast<Cx *>(half.data()), rt.data(), FFTW_ESTIMATE);
    r2c.execute_dft_r2c(re.data() + n, reinterpret_cast<Cx *>(half.data())); // the second row, not the planned one
    c2r.execute();
    for (int i = 0; i < n; ++i) EXPECT(std::abs(rt[i] - n * re[n + i]) <= tol * n * 8 * std::max(scale, 1.0));
    bool threw = false;
    try { r2c.execute_dft(reinterpret_cast<Cx *>(z.data()), reinterpret_cast<Cx *>(Z.data())); } catch (const std::logic_error &) { threw = true; }
    EXPECT(threw);
  }
  { // the engines of fftw.hpp:575-600 on their own aligned buffers: r2c, the reference's Hilbert multiplier, c2r
    const std::size_t n = 10;
    auto &eng = fftw::EngineR2C1D<T>::get(n);
    EXPECT(&eng == &fftw::EngineR2C1D<T>::get(n));                       // cached per length and thread
    EXPECT((reinterpret_cast<std::uintptr_t>(eng.buf.in) & 63) == 0);
    const double x[10] = {-0.999984, -0.736924, 0.511211, -0.0826997, 0.0655345, -0.562082, -0.905911, 0.357729, 0.358593, 0.869386};
    for (std::size_t i = 0; i < n; ++i) eng.buf.in[i] = (T)x[i];
    eng.forward();
    double dc = 0;
    for (double v : x) dc += v;
    EXPECT(std::abs(eng.buf.out[0][0] - dc) < 1e-5 && std::abs(eng.buf.out[0][1]) < 1e-5);
    eng.backward();
    for (std::size_t i = 0; i < n; ++i) EXPECT(std::abs(eng.buf.in[i] - n * x[i]) < 1e-4);
    auto &c2c = fftw::EngineDFT1D<T>::get(8);
    for (int i = 0; i < 8; ++i) { c2c.buf.in[i][0] = (T)(i == 1); c2c.buf.in[i][1] = 0; }
    c2c.forward(); // a shifted impulse: exp(-2 pi j k / 8)
    for (int k = 0; k < 8; ++k)
      EXPECT(std::abs(c2c.buf.out[k][0] - std::cos(2 * M_PI * k / 8)) < 1e-6 && std::abs(c2c.buf.out[k][1] + std::sin(2 * M_PI * k / 8)) < 1e-6);
  }
  bool threw = false;
  const int n = 16;
  std::vector<T> a(64), b(64);
  try {
    (void)fftw::Plan<T>::many_dft(1, &n, 2, reinterpret_cast<Cx *>(a.data()), nullptr, 2, n, reinterpret_cast<Cx *>(b.data()),
                                  nullptr, 1, n, FFTW_FORWARD, FFTW_ESTIMATE); // strided: not supported
  } catch (const std::invalid_argument &) { threw = true; }
  EXPECT(threw);
}

int main() {
  fftw::WisdomSetup wisdom(false);
  test_fftw_plans<double>(1e-13);
  test_fftw_plans<float>(1e-6);
  { // copy_to_padded_buffer
    const std::vector<int> src = {1, 2, 3};
    std::vector<int> dst(5, -1);
    fftconv::internal::copy_to_padded_buffer<int>(std::span<const int>(src), std::span<int>(dst));
    EXPECT(dst[0] == 1 && dst[1] == 2 && dst[2] == 3 && dst[3] == 0 && dst[4] == 0);
  }
  EXPECT(fftconv::internal::get_optimal_fft_size(95) == 512);
  EXPECT(fftconv::internal::get_optimal_fft_size(245) == 2048);
  { // engine interface
    using T = double;
    auto k = randn<T>(95, 1);
    auto &plans = fftconv::FFTConvEngine<T>::get_for_ksize(k.size());
    for (int n : {1000, 2000, 3000}) {
      auto a = randn<T>(n, n);
      std::vector<T> full(n + 94), same(n);
      plans.oaconvolve<ConvMode::Full>(a, k, full);
      plans.oaconvolve<ConvMode::Same>(a, k, same);
      near<T>(full, direct(a, k, ConvMode::Full));
      near<T>(same, direct(a, k, ConvMode::Same));
    }
  }
  test_conv<double, ConvMode::Full>(fftconv::convolve_fftw<double, ConvMode::Full>);
  test_conv<float, ConvMode::Full>(fftconv::convolve_fftw<float, ConvMode::Full>);
  test_conv<double, ConvMode::Same>(fftconv::convolve_fftw<double, ConvMode::Same>);
  test_conv<float, ConvMode::Same>(fftconv::convolve_fftw<float, ConvMode::Same>);
  test_conv<double, ConvMode::Full>(fftconv::convolve_fftw<double, ConvMode::Full, FFTW_PATIENT>);
  test_conv<double, ConvMode::Full>(fftconv::convolve_fftw<double, ConvMode::Full, FFTW_EXHAUSTIVE>);
  test_conv<double, ConvMode::Full>(fftconv::oaconvolve_fftw<double, ConvMode::Full>);
  test_conv<float, ConvMode::Full>(fftconv::oaconvolve_fftw<float, ConvMode::Full>);
  test_oaconv<double, ConvMode::Same>(fftconv::oaconvolve_fftw<double, ConvMode::Same>);
  test_oaconv<float, ConvMode::Same>(fftconv::oaconvolve_fftw<float, ConvMode::Same>);
  test_oaconv<double, ConvMode::Same>(fftconv::oaconvolve_fftw<double, ConvMode::Same, FFTW_MEASURE>);
  test_oaconv<double, ConvMode::Same>(fftconv::oaconvolve<double, ConvMode::Same>);
  { // default Mode is Same
    auto a = randn<double>(300, 5), k = randn<double>(31, 6);
    std::vector<double> out(300);
    fftconv::oaconvolve_fftw<double>(std::span<const double>(a), std::span<const double>(k), std::span<double>(out));
    near<double>(out, direct(a, k, ConvMode::Same));
  }
  { // batched forms
    const size_t batch = 5, n = 3000;
    auto a = randn<float>(batch * n, 9), k = randn<float>(245, 10);
    std::vector<float> out(batch * (n + 244));
    fftconv::oaconvolve_batched<float, ConvMode::Full>(std::span<const float>(a), n, std::span<const float>(k),
                                                       std::span<float>(out));
    for (size_t b = 0; b < batch; ++b) {
      std::vector<float> row(a.begin() + b * n, a.begin() + (b + 1) * n);
      std::vector<float> got(out.begin() + b * (n + 244), out.begin() + (b + 1) * (n + 244));
      near<float>(got, direct(row, k, ConvMode::Full), 100.0);
    }
  }
  { // filter handle: the plain call's samples (to rounding: a handle is planned for large jobs and may
    // use a different block size than a small plain call), taps transformed once
    const size_t batch = 3, n = 5000;
    auto a = randn<float>(batch * n, 19), k = randn<float>(245, 20);
    std::vector<float> want(batch * (n + 244)), got(batch * (n + 244));
    fftconv::oaconvolve_batched<float, ConvMode::Full>(std::span<const float>(a), n, std::span<const float>(k),
                                                       std::span<float>(want));
    fftconv::Filter<float> f{std::span<const float>(k)};
    f.oaconvolve_batched<ConvMode::Full>(std::span<const float>(a), n, std::span<float>(got));
    EXPECT(f.size() == 245);
    for (size_t i = 0; i < got.size(); ++i) EXPECT(std::abs(got[i] - want[i]) < 1e-4);
  }
  { // Hilbert known-answer vector, double and float, abs 1e-6
    const std::array<double, 10> inp = {-0.999984, -0.736924, 0.511211, -0.0826997, 0.0655345,
                                        -0.562082, -0.905911, 0.357729, 0.358593,   0.869386};
    const std::array<double, 10> expect = {1.45197493, 1.15365169, 0.54703078, 0.27346519, 0.15097965,
                                           0.83696245, 1.1476185,  0.71885109, 0.46089151, 1.07384968};
    std::array<double, 10> out{};
    fftconv::hilbert<double>(inp, out);
    for (int i = 0; i < 10; ++i) EXPECT(std::abs(out[i] - expect[i]) < 1e-6);
    std::array<float, 10> inf{}, outf{};
    for (int i = 0; i < 10; ++i) inf[i] = (float)inp[i];
    fftconv::hilbert<float>(inf, outf);
    for (int i = 0; i < 10; ++i) EXPECT(std::abs(outf[i] - expect[i]) < 1e-6 * 5);
  }
  { // envelope with fused post-processing: decimated + log-compressed samples of hilbert()
    const size_t batch = 3, n = 1000, dec = 3;
    auto x = randn<double>(batch * n, 23);
    std::vector<double> env(batch * n);
    fftconv::hilbert_batched<double>(std::span<const double>(x), n, std::span<double>(env));
    fftconv::EnvelopeOptions opt;
    opt.decimate = dec;
    opt.log_compress = true;
    opt.ref = 2.0;
    opt.dynamic_range_db = 40.0;
    const size_t m = opt.out_len(n);
    EXPECT(m == 334);
    std::vector<double> got(batch * m);
    fftconv::envelope_batched<double>(std::span<const double>(x), n, std::span<double>(got), opt);
    for (size_t b = 0; b < batch; ++b)
      for (size_t j = 0; j < m; ++j) {
        double want = 1.0 + 20.0 * std::log10(env[b * n + j * dec] / 2.0) / 40.0;
        want = want < 0 ? 0 : (want > 1 ? 1 : want);
        EXPECT(std::abs(got[b * m + j] - want) < 1e-12);
      }
    // rf_envelope == envelope of the Same-mode overlap-add
    auto k = randn<double>(65, 24);
    std::vector<double> filt(batch * n), want2(batch * n), got2(batch * n);
    fftconv::oaconvolve_batched<double, ConvMode::Same>(std::span<const double>(x), n, std::span<const double>(k),
                                                        std::span<double>(filt));
    fftconv::hilbert_batched<double>(std::span<const double>(filt), n, std::span<double>(want2));
    fftconv::rf_envelope_batched<double>(std::span<const double>(x), n, std::span<const double>(k),
                                         std::span<double>(got2));
    for (size_t i = 0; i < got2.size(); ++i) EXPECT(std::abs(got2[i] - want2[i]) < 1e-12 * 50);
  }
  { // errors surface as exceptions (the reference only asserts)
    bool threw = false;
    std::vector<double> a(10), k(3), out(10);
    try {
      fftconv::convolve_fftw<double, ConvMode::Full>(std::span<const double>(a), std::span<const double>(k),
                                                     std::span<double>(out));
    } catch (const std::runtime_error &) { threw = true; }
    EXPECT(threw);
  }
  { // Full mode into an oversized output: the reference zero-fills `out` (fftconv.hpp:385) and
    // accepts output.size() >= n + K - 1 (:340); the samples past n + K - 1 must come back zero
    auto a = randn<double>(500, 31), k = randn<double>(33, 32);
    std::vector<double> out(500 + 32 + 7, 123.0);
    fftconv::oaconvolve_fftw<double, ConvMode::Full>(std::span<const double>(a), std::span<const double>(k),
                                                     std::span<double>(out));
    std::vector<double> head(out.begin(), out.begin() + 532);
    near<double>(head, direct(a, k, ConvMode::Full));
    for (size_t i = 532; i < out.size(); ++i) EXPECT(out[i] == 0.0);
  }
  { // one process, every visible GPU (fftconv_cuda_mgpu_*): rows sharded; one long signal cut
    // into chunks with the K-1 tails exchanged over NCCL when there is more than one device
    fftconv::MultiGpu mg;
    std::printf("multi-GPU communicator over %d device(s)\n", mg.size());
    const size_t batch = 11, n = 6000;
    auto a = randn<float>(batch * n, 41), k = randn<float>(245, 42);
    std::vector<float> got(batch * (n + 244)), want(batch * (n + 244));
    fftconv::oaconvolve_batched<float, ConvMode::Full>(std::span<const float>(a), n, std::span<const float>(k),
                                                       std::span<float>(want));
    mg.oaconvolve_batched<float, ConvMode::Full>(std::span<const float>(a), n, std::span<const float>(k),
                                                 std::span<float>(got));
    // (not bit-identical across device counts: which rows share a complex FFT depends on the shard)
    for (size_t i = 0; i < got.size(); ++i) EXPECT(std::abs(got[i] - want[i]) < 1e-4);
    const size_t nl = 1 << 20;
    auto x = randn<double>(nl, 43), kd = randn<double>(245, 44);
    std::vector<double> yl(nl + 244);
    mg.oaconvolve_long<double>(std::span<const double>(x), std::span<const double>(kd), std::span<double>(yl));
    // exact check around every cut and at both ends
    std::vector<size_t> probe = {0, 1, 243, 244, nl - 1, nl, nl + 243};
    for (int d = 1; d < mg.size(); ++d) {
      size_t lo = 0, hi = 0;
      fftconv_cuda_mgpu_chunk_bounds(nl, mg.size(), d, &lo, &hi);
      for (size_t o = lo - 250; o < lo + 250; ++o) probe.push_back(o);
    }
    for (size_t o : probe) {
      double s = 0;
      for (size_t j = 0; j < 245; ++j)
        if (o >= j && o - j < nl) s += kd[j] * x[o - j];
      EXPECT(std::abs(yl[o] - s) < 1e-9);
    }
    if (mg.size() >= 2) { // two host threads, two devices, host-pointer calls: they must not serialise on a lock
      auto worker = [&](int dev, std::vector<float> *out) {
        fftconv::MultiGpu one{std::vector<int>{dev}};
        one.oaconvolve_batched<float, ConvMode::Full>(std::span<const float>(a), n, std::span<const float>(k),
                                                      std::span<float>(*out));
      };
      std::vector<float> o0(want.size()), o1(want.size());
      std::thread t0(worker, 0, &o0), t1(worker, 1, &o1);
      t0.join();
      t1.join();
      for (size_t i = 0; i < want.size(); ++i) EXPECT(o0[i] == want[i] && o1[i] == want[i]); // same call, same device count
    }
  }
  std::printf(g_fail ? "FAILED %d\n" : "ALL PASSED\n", g_fail);
  return g_fail ? 1 : 0;
}
