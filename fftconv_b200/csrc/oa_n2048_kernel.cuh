// oa_n2048_kernel.cuh -- the sm_100a overlap-add kernel for float32, FFT size 2048.
//
// Persistent grid: one CTA per SM, G groups of 128 threads per CTA.  Every group
// owns an equal, contiguous slice of the (stream-quad, segment) work list and
// walks it with oa_n2048_group.cuh::run_group.  Per segment quad a group issues
// 64 coalesced 4-byte loads and 64 coalesced 4-byte stores per thread (one
// 128-byte line per warp instruction), ~1160 packed-FP32 instructions per
// thread, and crosses four 128-thread named barriers.
//
// Shared memory per CTA: tw1 16 KB + hs 16 KB + tw2 1 KB, then per group a
// 32 KB exchange buffer and two (K-1)-point tail buffers.
#pragma once

#include "oa_n2048_group.cuh"

namespace fc {
namespace n2048 {

struct DeviceExec {
  const OaParams &p;
  int t;
  int bar_id;
  pt4 *xb;
  pt4 *tails;
  int tpts;
  const cpair *tw1, *tw2, *hs;

  __device__ __forceinline__ void sync() const {
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
  }
  __device__ __forceinline__ void iteration(const Lane (&ln)[4], long long j, bool warm,
                                            bool add_tail, int parity) {
    Regs v;
    io_load(t, v, p, ln, j);
    phase0(t, v, xb, tw1);
    sync();
    phase1(t, v, xb, tw2);
    sync();
    phase2(t, v, xb, hs);
    sync();
    phase3(t, v, xb, tw2);
    sync();
    phase4(t, v, xb, tw1);
    io_store(t, v, p, ln, j, tails + (parity ^ 1) * tpts, tails + parity * tpts, add_tail, warm);
  }
};

inline size_t smem_bytes(int groups, int klen) {
  return (size_t)(2048 + 2048 + 128) * sizeof(cpair) +
         (size_t)groups * (2048 + 2 * tail_pts(klen)) * sizeof(pt4);
}

template <int G>
__global__ void __launch_bounds__(128 * G, 1) oa_n2048_kernel(const __grid_constant__ OaParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpair *tw1 = reinterpret_cast<cpair *>(smem_raw);
  cpair *hs = tw1 + 2048;
  cpair *tw2 = hs + 2048;
  pt4 *grp = reinterpret_cast<pt4 *>(tw2 + 128);
  const int tpts = tail_pts(p.klen);

  for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
    tw1[i] = p.tw1[i];
    hs[i] = p.hs[i];
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) tw2[i] = p.tw2[i];
  __syncthreads();

  const int g = threadIdx.x >> 7;
  const long long slot = (long long)blockIdx.x * G + g;
  const long long n_slots = (long long)gridDim.x * G;
  // equal contiguous slices; 128-bit intermediate not needed: total_work * n_slots < 2^63
  const long long w0 = p.total_work / n_slots * slot + min(p.total_work % n_slots, slot);
  const long long w1 = w0 + p.total_work / n_slots + (slot < p.total_work % n_slots ? 1 : 0);

  pt4 *base = grp + (size_t)g * (2048 + 2 * tpts);
  DeviceExec ex{p, (int)(threadIdx.x & 127), g + 1, base, base + 2048, tpts, tw1, tw2, hs};
  run_group(p, w0, w1, ex);
}

} // namespace n2048
} // namespace fc
