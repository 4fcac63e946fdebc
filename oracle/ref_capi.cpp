/*
 * oracle/ref_capi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C entry points around the UNMODIFIED reference headers, compiled where they
 * lie (-I/root/reference/include) on top of oracle/fftw_shim.cpp.  Nothing of
 * the reference is copied: this file only instantiates and calls
 *   fftconv::oaconvolve_fftw   (include/fftconv/fftconv.hpp:474-480)
 *   fftconv::convolve_fftw     (include/fftconv/fftconv.hpp:453-461)
 *   fftconv::hilbert           (include/fftconv/hilbert.hpp:16-66)
 *   fftconv::internal::get_optimal_fft_size (include/fftconv/fftconv.hpp:53-66)
 * The batched variants loop the reference's single-signal call over rows,
 * sharded across std::thread -- one thread_local engine per thread, which is
 * how the reference is meant to be threaded (include/fftconv/fftw.hpp:443-455).
 *
 * Caveat printed with every number made from this library: the FFT underneath
 * is oracle/fftw_shim.cpp, not FFTW's SIMD codelets.
 *
 * Output: oracle/_ref/libfftconv_ref.so (git-ignored, ships to the GPU box).
 */
// <algorithm> first: the reference's AVX2 multiply (fftconv.hpp:173) uses
// std::min(initializer_list) without including it itself.
#include <algorithm>

#include <fftconv/fftconv.hpp>
#include <fftconv/hilbert.hpp>

#include <cstddef>
#include <span>
#include <thread>
#include <vector>

namespace {

template <class F> void shard(size_t batch, int nthreads, F &&fn) {
  size_t nt = nthreads > 0 ? (size_t)nthreads : std::max(1u, std::thread::hardware_concurrency());
  nt = std::max<size_t>(1, std::min(nt, batch));
  if (nt == 1) {
    for (size_t b = 0; b < batch; ++b) fn(b);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(nt);
  for (size_t t = 0; t < nt; ++t) {
    const size_t lo = batch * t / nt, hi = batch * (t + 1) / nt;
    pool.emplace_back([lo, hi, &fn]() {
      for (size_t b = lo; b < hi; ++b) fn(b);
    });
  }
  for (auto &th : pool) th.join();
}

template <class T>
void oa(const T *a, size_t n, const T *k, size_t klen, T *out, size_t out_len, int mode) {
  std::span<const T> sa(a, n), sk(k, klen);
  std::span<T> so(out, out_len);
  if (mode == 1) fftconv::oaconvolve_fftw<T, fftconv::ConvMode::Same>(sa, sk, so);
  else fftconv::oaconvolve_fftw<T, fftconv::ConvMode::Full>(sa, sk, so);
}

template <class T>
void cv(const T *a, size_t n, const T *k, size_t klen, T *out, size_t out_len, int mode) {
  std::span<const T> sa(a, n), sk(k, klen);
  std::span<T> so(out, out_len);
  if (mode == 1) fftconv::convolve_fftw<T, fftconv::ConvMode::Same>(sa, sk, so);
  else fftconv::convolve_fftw<T, fftconv::ConvMode::Full>(sa, sk, so);
}

template <class T> void hb(const T *x, T *env, size_t n) {
  fftconv::hilbert<T>(std::span<const T>(x, n), std::span<T>(env, n));
}

} // namespace

extern "C" {

size_t ref_optimal_fft_size(size_t klen) { return fftconv::internal::get_optimal_fft_size(klen); }
const char *ref_version(void) { return FFTCONV_VERSION; }
int ref_hardware_threads(void) { return (int)std::max(1u, std::thread::hardware_concurrency()); }

#define REF_API(SFX, T)                                                                      \
  void ref_oaconvolve_##SFX(const T *a, size_t n, const T *k, size_t klen, T *out,           \
                            size_t out_len, int mode) {                                      \
    oa<T>(a, n, k, klen, out, out_len, mode);                                                \
  }                                                                                          \
  void ref_convolve_##SFX(const T *a, size_t n, const T *k, size_t klen, T *out,             \
                          size_t out_len, int mode) {                                        \
    cv<T>(a, n, k, klen, out, out_len, mode);                                                \
  }                                                                                          \
  void ref_hilbert_##SFX(const T *x, T *env, size_t n) { hb<T>(x, env, n); }                 \
  void ref_oaconvolve_batched_##SFX(const T *a, size_t batch, size_t n, size_t a_stride,     \
                                    const T *k, size_t klen, T *out, size_t out_len,         \
                                    size_t out_stride, int mode, int nthreads) {             \
    shard(batch, nthreads, [&](size_t b) {                                                   \
      oa<T>(a + b * a_stride, n, k, klen, out + b * out_stride, out_len, mode);              \
    });                                                                                      \
  }                                                                                          \
  void ref_convolve_batched_##SFX(const T *a, size_t batch, size_t n, size_t a_stride,       \
                                  const T *k, size_t klen, T *out, size_t out_len,           \
                                  size_t out_stride, int mode, int nthreads) {               \
    shard(batch, nthreads, [&](size_t b) {                                                   \
      cv<T>(a + b * a_stride, n, k, klen, out + b * out_stride, out_len, mode);              \
    });                                                                                      \
  }                                                                                          \
  void ref_hilbert_batched_##SFX(const T *x, size_t batch, size_t n, size_t x_stride,        \
                                 T *env, size_t env_stride, int nthreads) {                  \
    shard(batch, nthreads, [&](size_t b) { hb<T>(x + b * x_stride, env + b * env_stride, n); }); \
  }

REF_API(f32, float)
REF_API(f64, double)

/* raw shim FFT, exported so tests can pin it against the reference's own
 * 20-point DFT known-answer vectors (test/test_fftw.cpp:261-299,347-393). */
void ref_rfft_f64(const double *x, size_t n, double *out_interleaved) {
  auto *in = fftw_alloc_real(n);
  auto *out = fftw_alloc_complex(n / 2 + 1);
  std::copy(x, x + n, in);
  auto p = fftw_plan_dft_r2c_1d((int)n, in, out, FFTW_ESTIMATE);
  fftw_execute(p);
  std::copy(&out[0][0], &out[0][0] + 2 * (n / 2 + 1), out_interleaved);
  fftw_destroy_plan(p); fftw_free(in); fftw_free(out);
}
void ref_irfft_f64(const double *in_interleaved, size_t n, double *x) {
  auto *out = fftw_alloc_real(n);
  auto *in = fftw_alloc_complex(n / 2 + 1);
  std::copy(in_interleaved, in_interleaved + 2 * (n / 2 + 1), &in[0][0]);
  auto p = fftw_plan_dft_c2r_1d((int)n, in, out, FFTW_ESTIMATE);
  fftw_execute(p);
  std::copy(out, out + n, x);
  fftw_destroy_plan(p); fftw_free(in); fftw_free(out);
}
void ref_rfft_f32(const float *x, size_t n, float *out_interleaved) {
  auto *in = fftwf_alloc_real(n);
  auto *out = fftwf_alloc_complex(n / 2 + 1);
  std::copy(x, x + n, in);
  auto p = fftwf_plan_dft_r2c_1d((int)n, in, out, FFTW_ESTIMATE);
  fftwf_execute(p);
  std::copy(&out[0][0], &out[0][0] + 2 * (n / 2 + 1), out_interleaved);
  fftwf_destroy_plan(p); fftwf_free(in); fftwf_free(out);
}
void ref_irfft_f32(const float *in_interleaved, size_t n, float *x) {
  auto *out = fftwf_alloc_real(n);
  auto *in = fftwf_alloc_complex(n / 2 + 1);
  std::copy(in_interleaved, in_interleaved + 2 * (n / 2 + 1), &in[0][0]);
  auto p = fftwf_plan_dft_c2r_1d((int)n, in, out, FFTW_ESTIMATE);
  fftwf_execute(p);
  std::copy(out, out + n, x);
  fftwf_destroy_plan(p); fftwf_free(in); fftwf_free(out);
}
}
