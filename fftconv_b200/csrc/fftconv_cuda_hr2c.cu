// fftconv_cuda_hr2c.cu -- translation unit of the real-input Hilbert kernel (hilbert_r2c_*.cuh).
// Kept apart from fftconv_cuda.cu so that the two compile in parallel.
#include "hilbert_r2c_kernel.cuh"
#include "hilbert_r2c_launch.h"
#include "hilbert_r2c_tables.h"

#include <cuda_runtime.h>

namespace fc {
namespace hr2c {

size_t table_bytes() { return kSmemTablesPadded; }

void build_tables(void *host_buf) {
  Quad *tw1 = static_cast<Quad *>(host_buf);
  Quad *tw2 = tw1 + kTw1Quads;
  Pair *mir = reinterpret_cast<Pair *>(tw2 + kTw2Quads);
  fill_tables(tw1, tw2, mir);
}

int launch(const float *in, float *out, long long batch, long long in_stride, long long out_stride,
           const void *tables, int sms, void *stream) {
  static bool configured[64] = {false}; // per device: the opt-in to > 48 KB of dynamic shared memory
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    e = cudaFuncSetAttribute(hilbert_r2c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes());
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  HrParams p;
  p.in = in;
  p.out = out;
  p.batch = batch;
  p.in_stride = in_stride;
  p.out_stride = out_stride;
  p.tw1 = static_cast<const Quad *>(tables);
  p.tw2 = p.tw1 + kTw1Quads;
  p.mir = reinterpret_cast<const Pair *>(p.tw2 + kTw2Quads);
  hilbert_r2c_kernel<<<(unsigned)sms, kGroups * kThreads, smem_bytes(), static_cast<cudaStream_t>(stream)>>>(p);
  return (int)cudaGetLastError();
}

} // namespace hr2c
} // namespace fc
