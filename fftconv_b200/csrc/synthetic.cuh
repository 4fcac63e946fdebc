// synthetic.cuh -- counter-based synthetic signal generator:
//   x[i] = g(hash64(seed, i)),  i = global element index
// so that any rank / the host can recompute any sample of a signal that never crosses PCIe
// (config 5: one signal of 2^31 samples split across GPUs -- every rank must be able to
// recompute its left neighbour's last K-1 inputs to check the exchanged tail exactly).
//
// g is bit-reproducible across CPU and GPU by construction: the four 16-bit fields of the
// hash are summed as integers (Irwin-Hall, n = 4: a bell-shaped density on [-3.46, 3.46],
// mean 0, variance 1), converted to floating point exactly (|v| <= 131070 < 2^24) and scaled
// by ONE multiplication with a constant -- one IEEE rounding, the same everywhere.  No log,
// cos or division.  The numpy twin is fftconv_b200/synthetic.py.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define FC_SYN_HD __host__ __device__ __forceinline__
#else
#define FC_SYN_HD inline
#endif

namespace fc {
namespace synthetic {

FC_SYN_HD uint64_t hash64(uint64_t seed, uint64_t i) { // splitmix64 of the (seed, counter) pair
  uint64_t z = seed + (i + 1ULL) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
FC_SYN_HD int irwin_hall4(uint64_t h) { // in [-131070, 131070]
  return (int)((h & 0xFFFFULL) + ((h >> 16) & 0xFFFFULL) + ((h >> 32) & 0xFFFFULL) + (h >> 48)) - 131070;
}
// 1 / sqrt(4 * (65536^2 - 1) / 12): unit variance
constexpr double kScale = 2.6428997921303014e-05;
template <typename T> FC_SYN_HD T sample(uint64_t seed, uint64_t i, T scale) {
  return (T)irwin_hall4(hash64(seed, i)) * scale;
}

#if defined(__CUDACC__)
// dst[i] = sample(seed, first + i) * 1, i < n.  `scale` = (T)(kScale * gain), rounded once on the host.
template <typename T>
__global__ void fill_kernel(T *__restrict__ dst, unsigned long long n, unsigned long long seed,
                            unsigned long long first, T scale) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = sample<T>(seed, first + i, scale);
}
#endif

} // namespace synthetic
} // namespace fc
