// hilbert_r2c_kernel.cuh -- the sm_100a kernel of the real-input Hilbert engine (hilbert_r2c_core.cuh).
//
// Persistent grid, one CTA of 4 x 128 threads per SM.  Each 128-thread group owns two shared-memory
// planes (34 KB) and one named barrier and works through its lines on its own: the four groups of an
// SM drift apart, so one group's shared-memory phases overlap another's arithmetic and global loads.
// Per line and thread:
//   16 x LDG.128 (the line, coalesced) | P1 | 32 x STS.64 | bar | 32 x LDS.64 | P2 | 32 x STS.64 | bar |
//   64 x LDS.32 | P3 | mirror pass in registers | P3' | 64 x STS.32 | bar | 32 x LDS.64 | P2' |
//   32 x STS.64 | bar | 32 x LDS.64 | P1' | 16 x LDG.128 (the line again, from L2) | sqrt | 16 x STG.128
// All six passes are straight-line code (2.6 k instructions; the instruction cache holds it: hit rate in
// profiles/).
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
// While the envelope is stored the registers of v free up, four slots per batch: the NEXT line's samples are
// requested into them (nq, consumed by the next iteration's first pass), so that no load latency is exposed at
// the top of the line loop.
// EPI: 0 = plain envelope, 1 = decimation / log compression sample by sample, 2 = decimation by a multiple of
// four: only the first sample of a thread's four can survive (epi.dec = decimation / 4, epi.magic for it)
template <int EPI, int ENVB = 2>
__device__ __forceinline__ void envelope_store(int t, const Regs &v, const float *x, float *out,
                                               const EnvEpilogue<float> &epi, float4 (&nq)[16], const float *x_next) {
  const float4 *xn = reinterpret_cast<const float4 *>(x_next) + t;
  const float4 *xi = reinterpret_cast<const float4 *>(x) + t;
  float4 *eo = reinterpret_cast<float4 *>(out) + t;
  // batches of two slots: two batches of the second read in flight (16 registers), and the next line's
  // samples take over the registers of the slots just consumed
  constexpr int B = ENVB;
  float4 q[2][B];
#pragma unroll
  for (int i = 0; i < B; ++i) q[0][i] = __ldg(xi + 128 * i);
#pragma unroll
  for (int b = 0; b < 16 / B; ++b) {
    if (b + 1 < 16 / B) {
#pragma unroll
      for (int i = 0; i < B; ++i) q[(b + 1) & 1][i] = __ldg(xi + 128 * (B * (b + 1) + i));
    }
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int a = B * b + i, s = slot16(a);
      const float4 xv = q[b & 1][i];
      if (EPI == 2) {
        const unsigned j = (unsigned)(128 * a + t);
        const unsigned qd = epi.dec > 1 ? __umulhi(j, epi.magic) : j;
        if (qd * (unsigned)epi.dec == j) {
          float y = sqrt_approx(fmaf(xv.x, xv.x, v.re[s].x * v.re[s].x));
          if (epi.log) {
            y = epi.a * env_log2(y) + epi.b; // env = 0 -> -inf -> 0
            y = y < 0.0f ? 0.0f : (y > 1.0f ? 1.0f : y);
            if (!(y == y)) y = 0.0f;
          }
          out[qd] = y;
        }
        continue;
      }
      float4 e;
      e.x = sqrt_approx(fmaf(xv.x, xv.x, v.re[s].x * v.re[s].x));
      e.y = sqrt_approx(fmaf(xv.y, xv.y, v.im[s].x * v.im[s].x));
      e.z = sqrt_approx(fmaf(xv.z, xv.z, v.re[s].y * v.re[s].y));
      e.w = sqrt_approx(fmaf(xv.w, xv.w, v.im[s].y * v.im[s].y));
      if (EPI == 0) {
        eo[128 * a] = e;
      } else { // decimation / log compression: sample by sample
        const long long i0 = 4LL * (128 * a + t);
        env_store<false>(epi, out, i0, e.x);
        env_store<false>(epi, out, i0 + 1, e.y);
        env_store<false>(epi, out, i0 + 2, e.z);
        env_store<false>(epi, out, i0 + 3, e.w);
      }
    }
    // unconditional (the caller passes the current line again when there is no next one): a conditional
    // assignment would keep the old nq alive through the whole iteration -- 64 more registers
#pragma unroll
    for (int i = 0; i < B; ++i) nq[B * b + i] = xn[128 * (B * b + i)];
  }
}
// the first line of a group: all sixteen loads at once
__device__ __forceinline__ void load_quads(int t, float4 (&nq)[16], const float *x) {
  const float4 *q = reinterpret_cast<const float4 *>(x) + t;
#pragma unroll
  for (int a = 0; a < 16; ++a) nq[a] = q[128 * a];
}
// P1 input: slot a = points m = 256 a + 2 t and + 1 (low / high halves)
__device__ __forceinline__ void quads_to_regs(const float4 (&nq)[16], Regs &v) {
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    v.re[a] = make_f2(nq[a].x, nq[a].z);
    v.im[a] = make_f2(nq[a].y, nq[a].w);
  }
}
// ask for the group's next line to be brought into L2 while the current one is transformed
__device__ __forceinline__ void prefetch_line_l2(const float *x) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(x), "r"(kRow * 4) : "memory");
}

template <int GROUPS, int EPI, int ENVB = 2>
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
  // (measured with the second read of the line in batches of 1, 2 and 4 loads: no difference beyond the run-to-run
  // drift of the clocks; 2 it is)
  const int bar = g + 1;
  const long long n_groups = (long long)gridDim.x * GROUPS;

  float4 nq[16]; // the samples of the line about to be transformed
  {
    const long long first = (long long)blockIdx.x * GROUPS + g;
    if (first < p.batch) load_quads(t, nq, p.in + first * p.in_stride);
  }
#pragma unroll 1
  for (long long line = (long long)blockIdx.x * GROUPS + g; line < p.batch; line += n_groups) {
    const float *x = p.in + line * p.in_stride;
    const float *x_next = line + n_groups < p.batch ? x + n_groups * p.in_stride : nullptr;
    if (t == 0 && x_next) prefetch_line_l2(x_next);
    // Straight-line code, six inlined dft16: with the passes in two- or three-trip loops the compiler either
    // re-paired the 32 points with ~1000 MOVs per line (points carried in registers around the loop) or kept the
    // next line's 64 prefetched registers alive through whole trips (anything assigned in one branch of a trip).
    { // P1: straight-line, so that nq is dead before anything else needs its registers
      Regs v;
      quads_to_regs(nq, v);
      dft16<-1>(v);
      fwd_a_pre(v, planes, m);
      group_barrier(bar);
    }
    { // P2
      Regs v;
      fwd_a_post(v, planes, m);
      dft16<-1>(v);
      fwd_b_pre(v, planes, m);
      __syncwarp(); // P2 -> P3 stays inside the warp (p2_k0)
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
    { // P2'
      Regs v;
      inv_a_post(v, planes, m);
      dft16<+1>(v);
      inv_b_pre(v, planes, m);
      group_barrier(bar);
    }
    { // P1', envelope; the next line's samples are requested as registers free up
      Regs v;
      inv_b_post(v, planes, m);
      dft16<+1>(v);
      envelope_store<EPI, ENVB>(t, v, x, p.out + line * p.out_stride, p.epi, nq, x_next ? x_next : x);
    }
  }
}
#endif // __CUDACC__

} // namespace hr2c
} // namespace fc
