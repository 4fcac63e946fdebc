// hilbert_r2c_kernel.cuh -- the sm_100a kernel of the real-input Hilbert engine (hilbert_r2c_core.cuh).
//
// Persistent grid, one CTA of 4 x 128 threads per SM.  Each 128-thread group owns two shared-memory
// planes (34 KB) and one named barrier and works through its lines on its own: the four groups of an
// SM drift apart, so one group's shared-memory phases overlap another's arithmetic and global loads.
// Per line and thread:
//   16 x LDG.128 (the line, coalesced) | P1 | 32 x STS.64 | bar | 32 x LDS.64 | P2 | 32 x STS.64 | bar |
//   64 x LDS.32 | P3 | mirror pass in registers | P3' | 64 x STS.32 | bar | 32 x LDS.64 | P2' |
//   32 x STS.64 | bar | 32 x LDS.64 | P1' | 16 x LDG.128 (the line again, from L2) | sqrt | 16 x STG.128
// The forward and the inverse dft16 each exist ONCE, in three-trip loops: six inlined copies would not
// fit the 32 KB instruction cache.
#pragma once

#include "hilbert_r2c_core.cuh"

namespace fc {
namespace hr2c {

struct HrParams {
  const float *in;
  float *out;
  long long batch, in_stride, out_stride;
  const Quad *tw1, *tw2; // kTw1Quads, kTw2Quads
  const Pair *mir;       // kMirPairs
};

constexpr size_t kSmemTables = (kTw1Quads + kTw2Quads) * sizeof(Quad) + kMirPairs * sizeof(Pair);
constexpr size_t kSmemTablesPadded = (kSmemTables + 15) & ~size_t(15);
constexpr size_t smem_bytes() { return kSmemTablesPadded + (size_t)kGroups * kGroupWords * sizeof(float); }

#if defined(__CUDACC__)
__device__ __forceinline__ void group_barrier(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kThreads) : "memory");
}
__device__ __forceinline__ float sqrt_approx(float x) { // MUFU.SQRT
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// slot16(a) of v holds g[256 a + 2 t + h] = h[2m] + j h[2m+1]: env = sqrt(x^2 + h^2), four samples per store
__device__ __forceinline__ void envelope_store(int t, const Regs &v, const float *x, float *out) {
  const float4 *xi = reinterpret_cast<const float4 *>(x) + t;
  float4 *eo = reinterpret_cast<float4 *>(out) + t;
#pragma unroll
  for (int half = 0; half < 2; ++half) { // eight loads in flight at a time
    float4 q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = __ldg(xi + 128 * (8 * half + i));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int s = slot16(8 * half + i);
      float4 e;
      e.x = sqrt_approx(fmaf(q[i].x, q[i].x, v.re[s].x * v.re[s].x));
      e.y = sqrt_approx(fmaf(q[i].y, q[i].y, v.im[s].x * v.im[s].x));
      e.z = sqrt_approx(fmaf(q[i].z, q[i].z, v.re[s].y * v.re[s].y));
      e.w = sqrt_approx(fmaf(q[i].w, q[i].w, v.im[s].y * v.im[s].y));
      eo[128 * (8 * half + i)] = e;
    }
  }
}

__global__ void __launch_bounds__(kGroups *kThreads, 1) hilbert_r2c_kernel(const __grid_constant__ HrParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Quad *tw1 = reinterpret_cast<Quad *>(smem_raw);
  Quad *tw2 = tw1 + kTw1Quads;
  Pair *mir = reinterpret_cast<Pair *>(tw2 + kTw2Quads);
  float *planes_all = reinterpret_cast<float *>(smem_raw + kSmemTablesPadded);

  for (int i = threadIdx.x; i < kTw1Quads; i += blockDim.x) tw1[i] = p.tw1[i];
  for (int i = threadIdx.x; i < kTw2Quads; i += blockDim.x) tw2[i] = p.tw2[i];
  for (int i = threadIdx.x; i < kMirPairs; i += blockDim.x) mir[i] = p.mir[i];
  __syncthreads();

  const int g = threadIdx.x >> 7, t = threadIdx.x & (kThreads - 1);
  float *planes = planes_all + g * kGroupWords;
  const ThreadMaps m = thread_maps(t, tw1, tw2, mir);
  const int bar = g + 1;
  const long long n_groups = (long long)gridDim.x * kGroups;

#pragma unroll 1
  for (long long line = (long long)blockIdx.x * kGroups + g; line < p.batch; line += n_groups) {
    const float *x = p.in + line * p.in_stride;
    Regs v;
    load_line(t, v, x);
#pragma unroll 1
    for (int trip = 0; trip < 3; ++trip) {
      dft16<-1>(v);
      if (trip == 0) {
        fwd_a_pre(v, planes, m);
        group_barrier(bar);
        fwd_a_post(v, planes, m);
      } else if (trip == 1) {
        fwd_b_pre(v, planes, m);
        group_barrier(bar);
        fwd_b_post(v, planes, m);
      } else {
        fwd_c(t, v, m, mir);
      }
    }
#pragma unroll 1
    for (int trip = 0; trip < 3; ++trip) {
      dft16<+1>(v);
      if (trip == 0) {
        inv_a_pre(v, planes, m);
        group_barrier(bar);
        inv_a_post(v, planes, m);
      } else if (trip == 1) {
        inv_b_pre(v, planes, m);
        group_barrier(bar);
        inv_b_post(v, planes, m);
      } else {
        envelope_store(t, v, x, p.out + line * p.out_stride);
      }
    }
  }
}
#endif // __CUDACC__

} // namespace hr2c
} // namespace fc
