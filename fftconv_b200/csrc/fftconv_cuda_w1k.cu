// fftconv_cuda_w1k.cu -- translation unit of the warp-per-FFT overlap-add kernel (oa_w1k_*.cuh).
// Kept apart from fftconv_cuda.cu so that the two compile in parallel.
#include "oa_w1k_kernel.cuh"
#include "oa_w1k_launch.h"
#include "oa_w1k_tables.h"

#include <cuda_runtime.h>

#include <cstdlib>

namespace fc {
namespace w1k {

size_t launch_smem_bytes() { return smem_bytes(20); }

void build_twiddles(cpair *tw) { fill_tw(tw); }

template <int WARPS, bool SHARED_DFT, bool FULL_TW = false> static int launch_variant(const W1kParams &p, int sms, void *stream) {
  static bool configured[64] = {false}; // per device: the opt-in to > 48 KB of dynamic shared memory
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  auto kern = oa_w1k_kernel<WARPS, SHARED_DFT, FULL_TW>;
  const size_t smem = smem_bytes(WARPS);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  kern<<<(unsigned)sms, 32 * WARPS, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return (int)cudaGetLastError();
}

int warps_per_cta() {
  const char *w = std::getenv("FFTCONV_CUDA_W1K_WARPS");
  return (w && std::atoi(w) == 20) ? 20 : 16; // measured: 20 warps of 96 registers (72 B of spills) are 10 % slower
}

int launch(const W1kParams &p, int sms, void *stream) {
  // FFTCONV_CUDA_W1K_UNROLLED=1: four inlined dft16 copies (the first version; kept for A/B measurements)
  // FFTCONV_CUDA_W1K_WARPS=20: twenty warps of 96 registers instead of sixteen of 128 (A/B only)
  const char *v = std::getenv("FFTCONV_CUDA_W1K_UNROLLED");
  if (v && *v && *v != '0') return launch_variant<16, false>(p, sms, stream);
  // All fifteen pair twiddles come from shared memory; FFTCONV_CUDA_W1K_TWTABLE=0 recomputes eleven of them
  // from four (the round-2 first version; measured 0.4-0.7 % slower: the FP32 pipe is the fuller one)
  const char *tt = std::getenv("FFTCONV_CUDA_W1K_TWTABLE");
  const bool table = !(tt && *tt == '0');
  if (warps_per_cta() == 16) return table ? launch_variant<16, true, true>(p, sms, stream) : launch_variant<16, true>(p, sms, stream);
  return launch_variant<20, true>(p, sms, stream);
}

} // namespace w1k
} // namespace fc
