// include/fftconv/multi_gpu.hpp -- the hot path across the GPUs of one box, from one C++ process.
//
// The reference is single-device and single-threaded (one signal per call,
// /root/reference/include/fftconv/fftconv.hpp:474-480); BASELINE.json's north star adds the two
// partitions implemented behind fftconv_cuda_mgpu_* (include/fftconv_cuda.h):
//   * batches of signals: rows sharded over the devices, no communication;
//   * one very long signal: contiguous chunks, the K-1 sample tail of each chunk
//     (FFTConvEngine::oaconvolve's carry between blocks, fftconv.hpp:387-410) goes to the
//     neighbouring GPU by NCCL send / recv.
// Host spans in, host spans out; same error convention as fftconv.hpp (std::runtime_error).
#pragma once

#include "fftconv.hpp"

#include <vector>

namespace fftconv {

class MultiGpu {
public:
  // devices: CUDA ordinals; empty = every visible device
  explicit MultiGpu(const std::vector<int> &devices = {}) {
    int n = static_cast<int>(devices.size());
    if (n == 0) n = visible_devices();
    internal::check(fftconv_cuda_mgpu_init(n, devices.empty() ? nullptr : devices.data(), &h_));
  }
  MultiGpu(const MultiGpu &) = delete;
  MultiGpu &operator=(const MultiGpu &) = delete;
  MultiGpu(MultiGpu &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  ~MultiGpu() { fftconv_cuda_mgpu_destroy(h_); }

  [[nodiscard]] int size() const { return fftconv_cuda_mgpu_device_count(h_); }

  // batch rows of n samples each, row-major; output rows of n + K - 1 (Full) or n (Same)
  template <FloatOrDouble T, ConvMode Mode = ConvMode::Same>
  void oaconvolve_batched(std::span<const T> signals, std::size_t n, std::span<const T> kernel, std::span<T> output) {
    if (n == 0 || signals.size() % n != 0) throw std::runtime_error("fftconv: signals.size() must be a multiple of n");
    const std::size_t batch = signals.size() / n;
    const std::size_t out_len = Mode == ConvMode::Full ? n + kernel.size() - 1 : n;
    if (output.size() != batch * out_len) throw std::runtime_error("fftconv: output.size() must be batch * row length");
    if constexpr (std::is_same_v<T, float>)
      internal::check(fftconv_cuda_mgpu_oaconvolve_batched_f32(h_, signals.data(), batch, n, n, kernel.data(), kernel.size(),
                                                               output.data(), out_len, out_len, internal::c_mode(Mode)));
    else
      internal::check(fftconv_cuda_mgpu_oaconvolve_batched_f64(h_, signals.data(), batch, n, n, kernel.data(), kernel.size(),
                                                               output.data(), out_len, out_len, internal::c_mode(Mode)));
  }

  template <FloatOrDouble T> void hilbert_batched(std::span<const T> x, std::size_t n, std::span<T> env) {
    if (n == 0 || x.size() % n != 0 || env.size() != x.size()) throw std::runtime_error("fftconv: bad hilbert batch shape");
    if constexpr (std::is_same_v<T, float>)
      internal::check(fftconv_cuda_mgpu_hilbert_batched_f32(h_, x.data(), x.size() / n, n, n, env.data(), n));
    else
      internal::check(fftconv_cuda_mgpu_hilbert_batched_f64(h_, x.data(), x.size() / n, n, n, env.data(), n));
  }

  // one long signal, Full mode: output.size() == input.size() + kernel.size() - 1
  template <FloatOrDouble T>
  void oaconvolve_long(std::span<const T> input, std::span<const T> kernel, std::span<T> output) {
    if (output.size() != input.size() + kernel.size() - 1) throw std::runtime_error("fftconv: output.size() must be n + K - 1");
    if constexpr (std::is_same_v<T, float>)
      internal::check(fftconv_cuda_mgpu_oaconvolve_long_host_f32(h_, input.data(), input.size(), kernel.data(), kernel.size(), output.data()));
    else
      internal::check(fftconv_cuda_mgpu_oaconvolve_long_host_f64(h_, input.data(), input.size(), kernel.data(), kernel.size(), output.data()));
  }

private:
  static int visible_devices() {
    // probe by asking for a communicator of one device more until it fails would be silly:
    // the C ABI reports the count through a failed init's message only, so count here by trial
    int n = 0;
    void *probe = nullptr;
    while (n < 64) {
      const int dev = n;
      if (fftconv_cuda_mgpu_init(1, &dev, &probe) != FFTCONV_CUDA_OK) break;
      fftconv_cuda_mgpu_destroy(probe);
      ++n;
    }
    if (n == 0) throw std::runtime_error(std::string("fftconv: ") + fftconv_cuda_last_error());
    return n;
  }
  void *h_ = nullptr;
};

} // namespace fftconv
