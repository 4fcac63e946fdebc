// include/fftconv/hilbert.hpp -- Hilbert-transform envelope, B200 back end.
//
// Same signature as the reference (/root/reference/include/fftconv/hilbert.hpp:16-66):
//   template <fftw::Floating T> void hilbert(std::span<const T> x, std::span<T> env)
// env[i] = sqrt(x[i]^2 + H{x}[i]^2), the magnitude of the analytic signal, i.e.
// np.abs(scipy.signal.hilbert(x)).  Forwards to libfftconv_cuda; the reference's
// optional Intel IPP variant (:68-181) is out of scope.
#pragma once

#include "fftconv.hpp"

namespace fftconv {

template <fftw::Floating T> void hilbert(const std::span<const T> x, const std::span<T> env) {
  if (x.size() != env.size()) throw std::runtime_error("fftconv: hilbert needs x.size() == env.size()");
  if constexpr (std::is_same_v<T, float>)
    internal::check(fftconv_cuda_hilbert_f32(x.data(), env.data(), x.size()));
  else
    internal::check(fftconv_cuda_hilbert_f64(x.data(), env.data(), x.size()));
}

// batch rows of n samples each, row-major; host or device spans (see fftconv.hpp).
template <fftw::Floating T>
void hilbert_batched(std::span<const T> x, std::size_t n, std::span<T> env, void *stream = nullptr) {
  if (n == 0 || x.size() % n != 0 || env.size() != x.size())
    throw std::runtime_error("fftconv: hilbert_batched needs x.size() == env.size() == batch * n");
  const std::size_t batch = x.size() / n;
  if constexpr (std::is_same_v<T, float>)
    internal::check(fftconv_cuda_hilbert_batched_f32(x.data(), batch, n, n, env.data(), n, stream));
  else
    internal::check(fftconv_cuda_hilbert_batched_f64(x.data(), batch, n, n, env.data(), n, stream));
}

// ---- envelope with fused post-processing (new) ---------------------
// What callers of hilbert() do next -- decimation and log compression of the envelope -- done in
// the store of the Hilbert kernels (fftconv_cuda_envelope_*).  Default-constructed options give
// exactly hilbert_batched.  Output rows hold ceil(n / decimate) samples.
struct EnvelopeOptions {
  bool log_compress = false;       // y = clamp(1 + 20 log10(env / ref) / dynamic_range_db, 0, 1)
  double ref = 1.0;
  double dynamic_range_db = 60.0;
  std::size_t decimate = 1;        // y[j] = env[j * decimate]
  [[nodiscard]] std::size_t out_len(std::size_t n) const { return (n + (decimate ? decimate : 1) - 1) / (decimate ? decimate : 1); }
  [[nodiscard]] fftconv_cuda_envelope_opts c_opts() const {
    return fftconv_cuda_envelope_opts{log_compress ? 1 : 0, ref, dynamic_range_db, decimate};
  }
};

template <fftw::Floating T>
void envelope_batched(std::span<const T> x, std::size_t n, std::span<T> out, const EnvelopeOptions &opt = {},
                      void *stream = nullptr) {
  if (n == 0 || x.size() % n != 0) throw std::runtime_error("fftconv: envelope_batched needs x.size() == batch * n");
  const std::size_t batch = x.size() / n, m = opt.out_len(n);
  if (out.size() != batch * m) throw std::runtime_error("fftconv: envelope_batched needs out.size() == batch * ceil(n / decimate)");
  const fftconv_cuda_envelope_opts o = opt.c_opts();
  if constexpr (std::is_same_v<T, float>)
    internal::check(fftconv_cuda_envelope_batched_f32(x.data(), batch, n, n, out.data(), m, m, &o, stream));
  else
    internal::check(fftconv_cuda_envelope_batched_f64(x.data(), batch, n, n, out.data(), m, m, &o, stream));
}

// band-pass oaconvolve_fftw<T, Same>(row, kernel), then the envelope of the filtered row
template <fftw::Floating T>
void rf_envelope_batched(std::span<const T> x, std::size_t n, std::span<const T> kernel, std::span<T> out,
                         const EnvelopeOptions &opt = {}, void *stream = nullptr) {
  if (n == 0 || x.size() % n != 0) throw std::runtime_error("fftconv: rf_envelope_batched needs x.size() == batch * n");
  const std::size_t batch = x.size() / n, m = opt.out_len(n);
  if (out.size() != batch * m) throw std::runtime_error("fftconv: rf_envelope_batched needs out.size() == batch * ceil(n / decimate)");
  const fftconv_cuda_envelope_opts o = opt.c_opts();
  if constexpr (std::is_same_v<T, float>)
    internal::check(fftconv_cuda_rf_envelope_batched_f32(x.data(), batch, n, n, kernel.data(), kernel.size(), out.data(), m, m, &o, stream));
  else
    internal::check(fftconv_cuda_rf_envelope_batched_f64(x.data(), batch, n, n, kernel.data(), kernel.size(), out.data(), m, m, &o, stream));
}

} // namespace fftconv
