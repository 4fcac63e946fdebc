// fftconv_cuda_hr2c.cu -- translation unit of the real-input Hilbert kernel (hilbert_r2c_*.cuh).
// Kept apart from fftconv_cuda.cu so that the two compile in parallel.
#include "hilbert_r2c_kernel.cuh"
#include "hilbert_r2c_launch.h"
#include "hilbert_r2c_tables.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

namespace fc {
namespace hr2c {

size_t table_bytes() { return kSmemTablesPadded; }

void build_tables(void *host_buf) {
  Quad *tw1 = static_cast<Quad *>(host_buf);
  Quad *tw2 = tw1 + kTw1Quads;
  Pair *mir = reinterpret_cast<Pair *>(tw2 + kTw2Quads);
  fill_tables(tw1, tw2, mir);
}

template <int GROUPS, int EPI, int ENVB = 2>
static int launch_variant(const HrParams &p, long long batch, int sms, void *stream) {
  static bool configured[64] = {false}; // per device: the opt-in to > 48 KB of dynamic shared memory
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  auto kern = hilbert_r2c_kernel<GROUPS, EPI, ENVB>;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(GROUPS));
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const long long ctas = std::max<long long>(1, std::min<long long>(sms, (batch + GROUPS - 1) / GROUPS));
  kern<<<(unsigned)ctas, GROUPS * kThreads, smem_bytes(GROUPS), static_cast<cudaStream_t>(stream)>>>(p);
  return (int)cudaGetLastError();
}

int launch(const float *in, float *out, long long batch, long long in_stride, long long out_stride,
           const void *tables, const Epilogue &epi, int sms, void *stream) {
  HrParams p;
  p.in = in;
  p.out = out;
  p.batch = batch;
  p.in_stride = in_stride;
  p.out_stride = out_stride;
  p.tw1 = static_cast<const Quad *>(tables);
  p.tw2 = p.tw1 + kTw1Quads;
  p.mir = reinterpret_cast<const Pair *>(p.tw2 + kTw2Quads);
  // FFTCONV_CUDA_HR2C_GROUPS=3 / 5: lines in flight per SM (A/B measurements; 4 = 128 registers per thread)
  p.epi = EnvEpilogue<float>{epi.plain, epi.log, epi.dec, epi.magic, epi.a, epi.b};
  if (!epi.plain && epi.dec % 4 == 0) { // only the first of a thread's four samples can survive
    p.epi.dec = epi.dec / 4;
    p.epi.magic = p.epi.dec > 1 ? (unsigned)(4294967296ULL / (unsigned)p.epi.dec) + 1u : 0u; // exact for j < 2048
    return launch_variant<kGroups, 2>(p, batch, sms, stream);
  }
  if (!epi.plain) return launch_variant<kGroups, 1>(p, batch, sms, stream);
  const char *g = std::getenv("FFTCONV_CUDA_HR2C_GROUPS");
  const int groups = g ? std::atoi(g) : kGroups;
  if (groups == 3) return launch_variant<3, 0>(p, batch, sms, stream);
  if (groups == 5) return launch_variant<5, 0>(p, batch, sms, stream);
  return launch_variant<kGroups, 0>(p, batch, sms, stream);
}

} // namespace hr2c
} // namespace fc
