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
// envelope post-processing in the store (the fields of fc::EnvEpilogue<float>, envelope_epilogue.cuh)
struct Epilogue {
  int plain, log, dec;
  unsigned magic;
  float a, b;
};
// Envelope of `batch` float32 lines of 8192 samples (input rows 16-byte aligned; output rows too when the
// epilogue is plain).  `tables`: device copy of build_tables' buffer.  0 on success, otherwise a cudaError_t.
// `sms`: CTAs to launch.
int launch(const float *in, float *out, long long batch, long long in_stride, long long out_stride,
           const void *tables, const Epilogue &epi, int sms, void *stream);

} // namespace hr2c
} // namespace fc
