// oa_w2_kernel.cuh -- the sm_100a overlap-add kernel built on the warp-pair engine
// (oa_w2_core.cuh): float32, FFT size 2048.
//
// Persistent grid: one CTA per SM, G pairs of warps per CTA (G = 4: 8 warps per SM, up to
// 255 registers per thread -- 128 of them hold the thread's 32 packed complex points).
// Every pair owns an equal, contiguous slice of the (stream-quad, segment) work list.
//
// Time line of a pair per segment quad:
//   mbarrier wait      four input segments have landed in the staging area (cp.async.bulk / TMA,
//                      issued one iteration ahead), which aliases the two exchange buffers
//   LDS.32             both warps read every input: radix-2 stage in registers
//   bar(64) A          the staging area may now be overwritten
//   S1 | syncwarp | S2 (x H) | syncwarp | S1'      each warp on its own 16.5 KB buffer
//   hand-over: 16 points per thread to the partner warp      bar(64) B, read, bar(64) C
//   thread 0           requests the NEXT inputs (the buffers are idle from here on)
//   E +- O, tail hand-over (K-1 samples, ping-pong buffers), 128 coalesced 4-byte stores per
//   thread straight from registers; the two warp leaders describe the iteration after next
//
// Shared memory per CTA: hs 16 KB + tw 1.5 KB (+pad), then per pair 33 KB of exchange
// buffers, two (K-1)-point tail buffers, three packets, scheduler state, one mbarrier.
#pragma once

#include "oa_n2048_kernel.cuh"
#include "oa_w2_group.cuh"

namespace fc {
namespace w2 {

#if defined(__CUDACC__)
namespace dev = n2048::dev;

__device__ __forceinline__ void bar_pair(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// One segment quad on one warp (w = 0 / 1: even / odd bins).  ONE instruction stream for both
// warps; the 1024-point transform body is one code block run twice (forward, then the inverse
// as a forward transform of the swapped data).
__device__ __forceinline__ void pair_iteration(const OaParams &p, const Packet *cur, int t, int w, pt4 *xb_own,
                                               const pt4 *xb_partner, const float *stage, int LS,
                                               const cpair *tw, const cpair *hs_p,
                                               const pt4 *tail_prev, pt4 *tail_next, uint32_t mbar,
                                               uint32_t &in_phase, int bar_id, const Packet *next, float *stage_w) {
  const bool warm = cur->warm != 0;
  const bool add_tail = cur->first == 0;
  const bool in_fast = all4(cur->in_fast), out_fast = all4(cur->out_fast);
  const float sgn = w ? -1.0f : 1.0f;
  Regs v;
  if (in_fast) {
    int sh[4] = {cur->in_shift[0], cur->in_shift[1], cur->in_shift[2], cur->in_shift[3]};
    while (!dev::mbar_try_wait(mbar, in_phase)) {
    }
    in_phase ^= 1;
    fast_load(t, v, stage, LS, sh, p.step, sgn);
  } else {
    Lane ln[4] = {cur->ln[0], cur->ln[1], cur->ln[2], cur->ln[3]};
    io_load(t, v, p, ln, cur->j, sgn);
  }
  bar_pair(bar_id); // A: both warps have read their inputs out of the staging area
  if (w) twiddle64(v);
#pragma unroll 1
  for (int i = 1; i <= 4; ++i) {
    if (i == 3) __syncwarp(); // every lane has read its rows (trip 1) before columns are rewritten
    if (trip_head(i, w, t, v, xb_own, tw, hs_p)) {
      __syncwarp();
      get_rows(t, v, xb_own);
    }
  }
  combine_put(t, v, xb_own);
  bar_pair(bar_id); // B
  Out o;
  combine_get(t, v, xb_partner, sgn, o);
  bar_pair(bar_id); // C: the exchange buffers are idle
  if (w == 0 && t == 0) n2048::issue_loads(next, stage_w, LS, mbar);
  if (out_fast || warm) {
    const int off = t + 512 * w;
    fast_store(t, w, o, cur->out_dst[0] + off, cur->out_dst[1] + off, cur->out_dst[2] + off, cur->out_dst[3] + off,
               p.klen, p.step, tail_prev, tail_next, add_tail, out_fast);
  } else {
    Lane ln[4] = {cur->ln[0], cur->ln[1], cur->ln[2], cur->ln[3]};
    io_store(t, w, o, p, ln, cur->j, tail_prev, tail_next, add_tail, warm);
  }
}

template <int G>
__global__ void __launch_bounds__(64 * G, 1) oa_w2_kernel(const __grid_constant__ OaParams p,
                                                          const cpair *__restrict__ tw_g,
                                                          const cpair *__restrict__ hs_g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cpair *hs = reinterpret_cast<cpair *>(smem_raw);
  cpair *tw = hs + 2048;
  unsigned char *pair0 = reinterpret_cast<unsigned char *>(tw + kTwElems);

  const PairLayout L = pair_layout(p.klen);
  const int g = threadIdx.x >> 6, w = (threadIdx.x >> 5) & 1, t = threadIdx.x & 31;
  unsigned char *base = pair0 + (size_t)g * L.total;
  pt4 *xb0 = reinterpret_cast<pt4 *>(base + L.xb);
  pt4 *xb_own = xb0 + w * kXbPts, *xb_partner = xb0 + (w ^ 1) * kXbPts;
  float *stage = reinterpret_cast<float *>(base + L.xb); // input staging aliases both buffers
  pt4 *tails = reinterpret_cast<pt4 *>(base + L.tails);
  const int pk_stride = ((int)sizeof(Packet) + 15) & ~15;
  unsigned char *pk_base = base + L.packets;
  auto PK = [&](int i) -> Packet * { return reinterpret_cast<Packet *>(pk_base + i * pk_stride); };
  const uint32_t mbar = dev::smem_u32(base + L.mbar);
  const int tpts = tail_pts(p.klen);
  const int LS = L.lane_stride;

  for (int i = threadIdx.x; i < 2048; i += blockDim.x) hs[i] = hs_g[i];
  for (int i = threadIdx.x; i < kTwElems; i += blockDim.x) tw[i] = tw_g[i];

  const long long slot = (long long)blockIdx.x * G + g;
  const long long n_slots = (long long)gridDim.x * G;
  const long long w0 = p.total_work / n_slots * slot + min(p.total_work % n_slots, slot);
  const long long w1 = w0 + p.total_work / n_slots + (slot < p.total_work % n_slots ? 1 : 0);

  // the two warp leaders of a pair are its scheduler: leader w describes lanes 2 w and 2 w + 1
  const bool leader = t == 0;
  SchedSlot *sslot = reinterpret_cast<SchedSlot *>(base + L.sched) + 2 * w;
  if (leader) {
#pragma unroll 1
    for (int l = 0; l < 2; ++l) {
      sslot[l].it.init(w0, w1);
      fill_packet_lane(p, sslot[l], PK(0), 2 * w + l);
      fill_packet_lane(p, sslot[l], PK(1), 2 * w + l);
    }
  }
  if (w == 0 && t == 0) {
    dev::mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (w == 0 && t == 0) n2048::issue_loads(PK(0), stage, LS, mbar);

  const int bar_id = g + 1;
  uint32_t in_phase = 0;
  int parity = 0;
  int ci = 0;
  const cpair *hs_p = hs + w * 1024;

  while (PK(ci)->valid) {
    const Packet *cur = PK(ci);
    const int ni = ci == 2 ? 0 : ci + 1, nni = ni == 2 ? 0 : ni + 1;
    pt4 *tail_prev = tails + (parity ^ 1) * tpts, *tail_next = tails + parity * tpts;
    pair_iteration(p, cur, t, w, xb_own, xb_partner, stage, LS, tw, hs_p, tail_prev, tail_next,
                   mbar, in_phase, bar_id, PK(ni), stage);
    if (leader && PK(ni)->valid) {
#pragma unroll 1
      for (int l = 0; l < 2; ++l) fill_packet_lane(p, sslot[l], PK(nni), 2 * w + l);
    }
    parity ^= 1;
    ci = ni;
  }
}
#endif // __CUDACC__

} // namespace w2
} // namespace fc
