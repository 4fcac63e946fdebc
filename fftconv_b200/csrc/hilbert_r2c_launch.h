// hilbert_r2c_launch.h -- what the dispatcher (fftconv_cuda.cu) needs from the translation unit that
// holds the real-input Hilbert kernel (fftconv_cuda_hr2c.cu).
#pragma once

#include <cstddef>

namespace fc {
namespace hr2c {

constexpr size_t kLineSamples = 8192; // the line length this engine serves
// bytes of the device tables (twiddles, mirror cos / sin), and a host-side builder for them
size_t table_bytes();
void build_tables(void *host_buf);
// Envelope of `batch` float32 lines of 8192 samples (16-byte aligned rows).  `tables`: device copy of
// build_tables' buffer.  0 on success, otherwise a cudaError_t.  `sms`: CTAs to launch.
int launch(const float *in, float *out, long long batch, long long in_stride, long long out_stride,
           const void *tables, int sms, void *stream);

} // namespace hr2c
} // namespace fc
