// include/fftconv/fftconv.hpp -- drop-in C++20 front end of the B200 back end.
//
// Same names, template parameters, defaults and span signatures as the reference
// (/root/reference/include/fftconv/fftconv.hpp):
//   enum ConvMode { Full, Same }                                   :23
//   isSIMDAligned<A>(ptr)                                           :26-28
//   internal::get_optimal_fft_size(K)                               :53-66
//   internal::copy_to_padded_buffer(src, dst)                       :70-82
//   FFTConvEngine<T, Flag>::get / get_for_ksize / convolve<Mode> /
//       oaconvolve<Mode> / oaconvolve_same / oaconvolve_full        :314-445
//   convolve_fftw<T, Mode = Same, PlannerFlag>(in, kernel, out)     :453-461
//   oaconvolve_fftw<T, Mode = Same, PlannerFlag>(in, kernel, out)   :474-480
// plus the north-star's short names (convolve, oaconvolve, FloatOrDouble) and the
// batched forms the reference does not have.  Every function forwards to the C ABI
// of libfftconv_cuda (include/fftconv_cuda.h); nothing is computed on the host and
// there is no CPU fallback.  The reference's functions are void and only assert;
// here a failed call (bad sizes, no CUDA device, unsupported shape) throws
// std::runtime_error carrying fftconv_cuda_last_error().
#pragma once

#include "../fftconv_cuda.h"
#include "aligned_vector.hpp"
#include "fftw.hpp"

#include <algorithm>
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <span>
#include <stdexcept>
#include <string>

#define FFTCONV_VERSION "0.5.1" // API level of the reference this header mirrors (fftconv.hpp:18)

namespace fftconv {

using fftw::FloatOrDouble;
using fftw::Floating;

enum ConvMode { Full, Same };

template <std::size_t Alignment> bool isSIMDAligned(const void *ptr) {
  return reinterpret_cast<std::uintptr_t>(ptr) % Alignment == 0;
}

namespace internal {

inline void check(int rc) {
  if (rc != FFTCONV_CUDA_OK) throw std::runtime_error(std::string("fftconv: ") + fftconv_cuda_last_error());
}

inline auto get_optimal_fft_size(const std::size_t filter_size) -> std::size_t {
  return fftconv_cuda_optimal_fft_size(filter_size);
}

template <typename T> inline void copy_to_padded_buffer(const std::span<const T> src, const std::span<T> dst) {
  assert(src.size() <= dst.size());
  std::copy(src.begin(), src.end(), dst.begin());
  std::fill(dst.begin() + static_cast<std::ptrdiff_t>(src.size()), dst.end(), T(0));
}

constexpr int c_mode(ConvMode m) { return m == ConvMode::Full ? FFTCONV_CUDA_FULL : FFTCONV_CUDA_SAME; }

template <Floating T>
inline void oa(const T *a, std::size_t batch, std::size_t n, std::size_t a_stride, std::span<const T> k, T *out,
               std::size_t out_len, std::size_t out_stride, ConvMode mode, void *stream) {
  if constexpr (std::is_same_v<T, float>)
    check(fftconv_cuda_oaconvolve_batched_f32(a, batch, n, a_stride, k.data(), k.size(), out, out_len,
                                              out_stride, c_mode(mode), stream));
  else
    check(fftconv_cuda_oaconvolve_batched_f64(a, batch, n, a_stride, k.data(), k.size(), out, out_len,
                                              out_stride, c_mode(mode), stream));
}
template <Floating T>
inline void cv(const T *a, std::size_t batch, std::size_t n, std::size_t a_stride, std::span<const T> k, T *out,
               std::size_t out_len, std::size_t out_stride, ConvMode mode, void *stream) {
  if constexpr (std::is_same_v<T, float>)
    check(fftconv_cuda_convolve_batched_f32(a, batch, n, a_stride, k.data(), k.size(), out, out_len,
                                            out_stride, c_mode(mode), stream));
  else
    check(fftconv_cuda_convolve_batched_f64(a, batch, n, a_stride, k.data(), k.size(), out, out_len,
                                            out_stride, c_mode(mode), stream));
}

} // namespace internal

// The reference's engine object owns FFTW plans and buffers per FFT length; here
// it is a stateless handle (the plan cache lives behind the C ABI), kept so code
// written as FFTConvEngine<T>::get_for_ksize(k).oaconvolve<Mode>(a, k, out)
// (test/test_fftconv.cpp:97-103) compiles unchanged.
template <typename T, int PlannerFlag = FFTW_ESTIMATE> struct FFTConvEngine {
  using View = const std::span<const T>;
  using MutView = std::span<T>;

  static auto get(std::size_t /*padded_length*/) -> FFTConvEngine<T, PlannerFlag> & {
    static thread_local FFTConvEngine<T, PlannerFlag> e;
    return e;
  }
  static auto get_for_ksize(std::size_t /*ksize*/) -> FFTConvEngine<T> & { return FFTConvEngine<T>::get(0); }

  template <ConvMode Mode = ConvMode::Full> void convolve(View a, View k, MutView out) {
    internal::cv<T>(a.data(), 1, a.size(), a.size(), k, out.data(), out.size(), out.size(), Mode, nullptr);
  }
  template <ConvMode Mode = ConvMode::Full> void oaconvolve(View a, View k, MutView out) {
    internal::oa<T>(a.data(), 1, a.size(), a.size(), k, out.data(), out.size(), out.size(), Mode, nullptr);
  }
  void oaconvolve_same(View a, View k, MutView out) { oaconvolve<ConvMode::Same>(a, k, out); }
  void oaconvolve_full(View a, View k, MutView out) { oaconvolve<ConvMode::Full>(a, k, out); }
};

// 1-D convolution with one FFT of length >= n + K - 1.  NB the default Mode is
// Same, as in the reference.  Full needs output.size() == n + K - 1, Same needs
// output.size() == n; out[i] = full[i + K/2] in Same mode.
template <Floating T, ConvMode Mode = ConvMode::Same, int PlannerFlag = FFTW_ESTIMATE>
void convolve_fftw(const std::span<const T> input, const std::span<const T> kernel, std::span<T> output) {
  internal::cv<T>(input.data(), 1, input.size(), input.size(), kernel, output.data(), output.size(),
                  output.size(), Mode, nullptr);
}

// 1-D overlap-add convolution; segment FFT size from get_optimal_fft_size(K).
template <Floating T, ConvMode Mode = ConvMode::Same, int PlannerFlag = FFTW_ESTIMATE>
void oaconvolve_fftw(std::span<const T> input, std::span<const T> kernel, std::span<T> output) {
  internal::oa<T>(input.data(), 1, input.size(), input.size(), kernel, output.data(), output.size(),
                  output.size(), Mode, nullptr);
}

// README.md:29 / BASELINE.json north_star spell the entry points without _fftw.
template <FloatOrDouble T, ConvMode Mode = ConvMode::Same>
void convolve(std::span<const T> input, std::span<const T> kernel, std::span<T> output) {
  convolve_fftw<T, Mode>(input, kernel, output);
}
template <FloatOrDouble T, ConvMode Mode = ConvMode::Same>
void oaconvolve(std::span<const T> input, std::span<const T> kernel, std::span<T> output) {
  oaconvolve_fftw<T, Mode>(input, kernel, output);
}

// ---- batched forms (new) -------------------------------------------------------
// `signals` holds batch rows of n samples (row-major, contiguous); `output` holds
// batch rows of n + K - 1 (Full) or n (Same) samples.  Spans may view host memory
// or device memory (then the work is enqueued on `stream`, a cudaStream_t, and the
// call returns without synchronising).
template <FloatOrDouble T, ConvMode Mode = ConvMode::Same>
void oaconvolve_batched(std::span<const T> signals, std::size_t n, std::span<const T> kernel, std::span<T> output,
                        void *stream = nullptr) {
  if (n == 0 || signals.size() % n != 0) throw std::runtime_error("fftconv: signals.size() must be a multiple of n");
  const std::size_t batch = signals.size() / n;
  const std::size_t out_len = Mode == ConvMode::Full ? n + kernel.size() - 1 : n;
  if (output.size() != batch * out_len) throw std::runtime_error("fftconv: output.size() must be batch * row length");
  internal::oa<T>(signals.data(), batch, n, n, kernel, output.data(), out_len, out_len, Mode, stream);
}
template <FloatOrDouble T, ConvMode Mode = ConvMode::Same>
void convolve_batched(std::span<const T> signals, std::size_t n, std::span<const T> kernel, std::span<T> output,
                      void *stream = nullptr) {
  if (n == 0 || signals.size() % n != 0) throw std::runtime_error("fftconv: signals.size() must be a multiple of n");
  const std::size_t batch = signals.size() / n;
  const std::size_t out_len = Mode == ConvMode::Full ? n + kernel.size() - 1 : n;
  if (output.size() != batch * out_len) throw std::runtime_error("fftconv: output.size() must be batch * row length");
  internal::cv<T>(signals.data(), batch, n, n, kernel, output.data(), out_len, out_len, Mode, stream);
}

// ---- filter handle (new) --------------------------------------
// The reference transforms the kernel on every call (fftconv.hpp:391-392 there).  A Filter
// owns the taps and their spectrum on the current device, so repeated calls launch the
// overlap-add kernel alone.  Move-only; spans may view host or device memory.
template <FloatOrDouble T> class Filter {
public:
  explicit Filter(std::span<const T> kernel) : klen_(kernel.size()) {
    if constexpr (std::is_same_v<T, float>)
      internal::check(fftconv_cuda_filter_create_f32(kernel.data(), kernel.size(), &h_));
    else
      internal::check(fftconv_cuda_filter_create_f64(kernel.data(), kernel.size(), &h_));
  }
  Filter(const Filter &) = delete;
  Filter &operator=(const Filter &) = delete;
  Filter(Filter &&o) noexcept : h_(o.h_), klen_(o.klen_) { o.h_ = nullptr; }
  ~Filter() { fftconv_cuda_filter_destroy(h_); }
  [[nodiscard]] std::size_t size() const { return klen_; }

  template <ConvMode Mode = ConvMode::Same>
  void oaconvolve(std::span<const T> input, std::span<T> output, void *stream = nullptr) const {
    oaconvolve_batched<Mode>(input, input.size(), output, stream);
  }
  template <ConvMode Mode = ConvMode::Same>
  void oaconvolve_batched(std::span<const T> signals, std::size_t n, std::span<T> output, void *stream = nullptr) const {
    if (n == 0 || signals.size() % n != 0) throw std::runtime_error("fftconv: signals.size() must be a multiple of n");
    const std::size_t batch = signals.size() / n;
    const std::size_t out_len = Mode == ConvMode::Full ? n + klen_ - 1 : n;
    if (output.size() != batch * out_len) throw std::runtime_error("fftconv: output.size() must be batch * row length");
    if constexpr (std::is_same_v<T, float>)
      internal::check(fftconv_cuda_filter_oaconvolve_f32(h_, signals.data(), batch, n, n, output.data(), out_len,
                                                         out_len, internal::c_mode(Mode), stream));
    else
      internal::check(fftconv_cuda_filter_oaconvolve_f64(h_, signals.data(), batch, n, n, output.data(), out_len,
                                                         out_len, internal::c_mode(Mode), stream));
  }

private:
  void *h_ = nullptr;
  std::size_t klen_;
};

} // namespace fftconv
