// hilbert_r2c_kernel.cuh -- the sm_100a kernel of the real-input Hilbert engine (hilbert_r2c_core.cuh).
//
// Persistent grid, one CTA of 4 x 128 threads per SM.  Each 128-thread group owns two shared-memory
// planes (34 KB) and one named barrier and works through its lines on its own: the four groups of an
// SM drift apart, so one group's shared-memory phases overlap another's arithmetic and global loads.
// Per line and thread:
//   16 x LDG.128 (the line, coalesced) | P1 | 32 x STS.64 | bar | 32 x LDS.64 | P2 | 32 x STS.64 | bar |
//   64 x LDS.32 | P3 | mirror pass in registers | P3' | 64 x STS.32 | bar | 32 x LDS.64 | P2' |
//   32 x STS.64 | bar | 32 x LDS.64 | P1' | 16 x LDG.128 (the line again, from L2) | sqrt | 16 x STG.128
// Each direction's dft16 exists twice: once in a two-trip loop (P1 / P2, P2' / P1') and once in the
// straight-line middle section (P3, mirror, P3').
#pragma once

#include "envelope_epilogue.cuh"
#include "hilbert_r2c_core.cuh"

namespace fc {
namespace hr2c {

struct HrParams {
  const float *in;
  float *out;
  long long batch, in_stride, out_stride;
  const Quad *tw1, *tw2; // kTw1Quads, kTw2Quads
  const Pair *mir;       // kMirPairs
  EnvEpilogue<float> epi; // envelope post-processing in the store (envelope_epilogue.cuh); plain = the bare envelope
};

constexpr size_t kSmemTables = (kTw1Quads + kTw2Quads) * sizeof(Quad) + kMirPairs * sizeof(Pair);
constexpr size_t kSmemTablesPadded = (kSmemTables + 15) & ~size_t(15);
constexpr size_t smem_bytes(int groups = kGroups) { return kSmemTablesPadded + (size_t)groups * kGroupWords * sizeof(float); }

#if defined(__CUDACC__)
__device__ __forceinline__ void group_barrier(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kThreads) : "memory");
}
// Where two code paths deliver a thread's 32 points to the same dft16 (a trip that loads from global memory
// and one that loads from shared memory; the two mirror variants), the values cross the join as sixteen 64-bit
// registers per part.  As 32-bit halves the compiler picks ONE layout for the join -- that of the 128-bit
// global loads, (re A, im A, re B, im B) -- and re-pairs every value with MOVs on both sides of it; a 64-bit
// value can only live in an aligned register pair, which is what the packed instructions want.
struct Packed {
  unsigned long long re[16], im[16];
};
__device__ __forceinline__ unsigned long long pack64(f2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ f2 unpack64(unsigned long long v) {
  f2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ void pack_regs(const Regs &v, Packed &q) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    q.re[i] = pack64(v.re[i]);
    q.im[i] = pack64(v.im[i]);
  }
}
__device__ __forceinline__ void unpack_regs(const Packed &q, Regs &v) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v.re[i] = unpack64(q.re[i]);
    v.im[i] = unpack64(q.im[i]);
  }
}
__device__ __forceinline__ float sqrt_approx(float x) { // MUFU.SQRT
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// slot16(a) of v holds g[256 a + 2 t + h] = h[2m] + j h[2m+1]: env = sqrt(x^2 + h^2), four samples per store.
// The line is read again (from L2) in batches of four 128-bit loads, two batches in flight.
template <bool PLAIN>
__device__ __forceinline__ void envelope_store(int t, const Regs &v, const float *x, float *out,
                                               const EnvEpilogue<float> &epi) {
  const float4 *xi = reinterpret_cast<const float4 *>(x) + t;
  float4 *eo = reinterpret_cast<float4 *>(out) + t;
  float4 q[2][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) q[0][i] = __ldg(xi + 128 * i);
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    if (b < 3) {
#pragma unroll
      for (int i = 0; i < 4; ++i) q[(b + 1) & 1][i] = __ldg(xi + 128 * (4 * (b + 1) + i));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int s = slot16(4 * b + i);
      const float4 xv = q[b & 1][i];
      float4 e;
      e.x = sqrt_approx(fmaf(xv.x, xv.x, v.re[s].x * v.re[s].x));
      e.y = sqrt_approx(fmaf(xv.y, xv.y, v.im[s].x * v.im[s].x));
      e.z = sqrt_approx(fmaf(xv.z, xv.z, v.re[s].y * v.re[s].y));
      e.w = sqrt_approx(fmaf(xv.w, xv.w, v.im[s].y * v.im[s].y));
      if (PLAIN) {
        eo[128 * (4 * b + i)] = e;
      } else { // decimation / log compression: sample by sample
        const long long i0 = 4LL * (128 * (4 * b + i) + t);
        env_store<false>(epi, out, i0, e.x);
        env_store<false>(epi, out, i0 + 1, e.y);
        env_store<false>(epi, out, i0 + 2, e.z);
        env_store<false>(epi, out, i0 + 3, e.w);
      }
    }
  }
}
// ask for the group's next line to be brought into L2 while the current one is transformed
__device__ __forceinline__ void prefetch_line_l2(const float *x) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(x), "r"(kRow * 4) : "memory");
}

template <int GROUPS, bool PLAIN>
__global__ void __launch_bounds__(GROUPS *kThreads, 1) hilbert_r2c_kernel(const __grid_constant__ HrParams p) {
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
  ThreadMaps m = thread_maps(t, tw1, tw2, mir);
  // P3's 32-bit accesses must stay 32-bit: knowing that base3a / base3b are even, the compiler merges the loads
  // of two neighbouring points of ONE butterfly into a 64-bit load and then needs ~200 MOVs to pair each of them
  // with the other butterfly's point
  asm volatile("" : "+r"(m.base3a), "+r"(m.base3b));
  const int bar = g + 1;
  const long long n_groups = (long long)gridDim.x * GROUPS;

#pragma unroll 1
  for (long long line = (long long)blockIdx.x * GROUPS + g; line < p.batch; line += n_groups) {
    const float *x = p.in + line * p.in_stride;
    if (t == 0 && line + n_groups < p.batch) prefetch_line_l2(x + n_groups * p.in_stride);
    // Nothing but addresses lives across a trip boundary: every trip starts with a load (global or shared)
    // and ends with a store.  (With the 32 points carried in registers around three-trip loops the compiler
    // re-paired them with ~1000 MOVs per line.)
#pragma unroll 1
    for (int trip = 0; trip < 2; ++trip) { // P1, P2
      Regs v;
      Packed q;
      if (trip == 0) {
        load_line(t, v, x);
        pack_regs(v, q);
      } else {
        fwd_a_post(v, planes, m);
        pack_regs(v, q);
      }
      unpack_regs(q, v);
      dft16<-1>(v);
      if (trip == 0) {
        fwd_a_pre(v, planes, m);
        group_barrier(bar);
      } else {
        fwd_b_pre(v, planes, m);
        __syncwarp(); // P2 -> P3 stays inside the warp (p2_k0)
      }
    }
    { // P3, the mirror pass, P3': no exchange in between
      Regs v;
      fwd_b_post(v, planes, m);
      dft16<-1>(v);
      Packed q;
      if (t == kSelfPaired) {
        mirror_self(v, m.cs, mir + kThreads);
        pack_regs(v, q);
      } else {
        mirror(v, m.cs);
        pack_regs(v, q);
      }
      unpack_regs(q, v);
      dft16<+1>(v);
      inv_a_pre(v, planes, m);
      __syncwarp(); // P3' -> P2' stays inside the warp
    }
#pragma unroll 1
    for (int trip = 0; trip < 2; ++trip) { // P2', P1'
      Regs v;
      Packed q;
      if (trip == 0) {
        inv_a_post(v, planes, m);
        pack_regs(v, q);
      } else {
        inv_b_post(v, planes, m);
        pack_regs(v, q);
      }
      unpack_regs(q, v);
      dft16<+1>(v);
      if (trip == 0) {
        inv_b_pre(v, planes, m);
        group_barrier(bar);
      } else {
        envelope_store<PLAIN>(t, v, x, p.out + line * p.out_stride, p.epi);
      }
    }
  }
}
#endif // __CUDACC__

} // namespace hr2c
} // namespace fc
