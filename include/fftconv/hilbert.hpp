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

} // namespace fftconv
