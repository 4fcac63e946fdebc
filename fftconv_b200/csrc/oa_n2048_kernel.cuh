// oa_n2048_kernel.cuh -- the sm_100a overlap-add kernel for float32, FFT size 2048.
//
// Persistent grid: one CTA per SM, G groups of 128 threads per CTA.  Every group
// owns an equal, contiguous slice of the (stream-quad, segment) work list.
//
// Per segment quad a group
//   * waits on an mbarrier for the four input segments that thread 0 requested
//     one iteration earlier with cp.async.bulk (TMA, global -> shared), and reads
//     them with 4-byte shared loads at immediate offsets;
//   * runs phases P0..P4 of oa_n2048_core.cuh (~1100 packed-FP32 instructions per
//     thread) across four 128-thread named barriers;
//   * writes the 4 x step results into a shared staging buffer that thread 0
//     drains with cp.async.bulk (shared -> global) during the next iteration.
// Thread 0 of each group is also the group's scheduler: it walks the work slice
// one iteration ahead and publishes a small descriptor ("packet") in shared
// memory, so the other 127 threads execute no 64-bit control arithmetic at all.
//
// Shared memory per CTA: hs 16 KB + tw1 16 KB + tw2 1 KB, then per group a 32 KB
// exchange buffer, input and output staging (4 lanes each), two tail buffers,
// two packets and one mbarrier.
#pragma once

#include "oa_n2048_group.cuh"

#include <cstdint>

namespace fc {
namespace n2048 {

#if defined(__CUDACC__)
namespace dev {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bar_sync(int id) {
  asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(mbar), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared, completion counted in bytes on an mbarrier (TMA, SASS UBLKCP)
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

} // namespace dev

// Thread 0: describe the next iteration and, if it is a fast-path load, start the bulk copies.
__device__ __forceinline__ void schedule_next(const OaParams &p, WorkIter &it, const Packet *prev, Packet *pk,
                                              float *in_stage, int lane_stride, uint32_t mbar) {
  InPlan ip;
  fill_packet(p, it, prev, pk, ip);
  if (ip.fast) {
    dev::fence_proxy_async(); // the staging buffer was read through the generic proxy
    dev::mbar_expect_tx(mbar, (uint32_t)(ip.bytes[0] + ip.bytes[1] + ip.bytes[2] + ip.bytes[3]));
#pragma unroll
    for (int l = 0; l < 4; ++l)
      dev::bulk_load(dev::smem_u32(in_stage + l * lane_stride), ip.src[l], (uint32_t)ip.bytes[l], mbar);
  }
}

// Thread 0: drain the staged outputs of the previous iteration to global memory.
__device__ __forceinline__ void flush_out(const Packet *pk, const float *out_stage, int lane_stride, int step) {
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const int d = pk->out_shift[l];
    const FlushSplit f = flush_split(d, step);
    const int head = f.head, nb = f.nb;
    float *dst = pk->out_dst[l];
    const float *src = out_stage + l * lane_stride + d;
    dev::bulk_store(dst + head, dev::smem_u32(src + head), (uint32_t)nb * 4u);
    for (int i = 0; i < head; ++i) dst[i] = src[i];
    for (int i = head + nb; i < step; ++i) dst[i] = src[i];
  }
  dev::bulk_commit();
}

template <int G>
__global__ void __launch_bounds__(128 * G, 1) oa_n2048_kernel(const __grid_constant__ OaParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cpair *hs = reinterpret_cast<cpair *>(smem_raw);
  cpair *tw1 = hs + 2048;
  cpair *tw2 = tw1 + 2048;
  unsigned char *grp0 = reinterpret_cast<unsigned char *>(tw2 + 128);

  const GroupLayout L = group_layout(p.klen);
  const int g = threadIdx.x >> 7, t = threadIdx.x & 127;
  unsigned char *base = grp0 + (size_t)g * L.total;
  pt4 *xb = reinterpret_cast<pt4 *>(base + L.xb);
  float *in_stage = reinterpret_cast<float *>(base + L.in);
  float *out_stage = reinterpret_cast<float *>(base + L.out);
  pt4 *tails = reinterpret_cast<pt4 *>(base + L.tails);
  const int pk_stride = ((int)sizeof(Packet) + 15) & ~15;
  Packet *pk0 = reinterpret_cast<Packet *>(base + L.packets);
  Packet *pk1 = reinterpret_cast<Packet *>(base + L.packets + pk_stride);
  const uint32_t mbar = dev::smem_u32(base + L.mbar);
  const int tpts = tail_pts(p.klen);
  const int LS = L.lane_stride;

  for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
    tw1[i] = p.tw1[i];
    hs[i] = p.hs[i];
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) tw2[i] = p.tw2[i];

  const long long slot = (long long)blockIdx.x * G + g;
  const long long n_slots = (long long)gridDim.x * G;
  const long long w0 = p.total_work / n_slots * slot + min(p.total_work % n_slots, slot);
  const long long w1 = w0 + p.total_work / n_slots + (slot < p.total_work % n_slots ? 1 : 0);

  WorkIter &it = *reinterpret_cast<WorkIter *>(base + L.sched); // used by thread 0 only
  if (t == 0) {
    dev::mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    it.init(w0, w1);
    schedule_next(p, it, pk1, pk0, in_stage, LS, mbar);
  }
  __syncthreads();

  const ThreadMasks masks = make_masks(t, p);
  const int bar_id = g + 1;
  uint32_t in_phase = 0;
  int parity = 0;
  bool out_pending = false; // a staged output quad waits to be flushed (uniform)
  Packet *cur = pk0, *nxt = pk1;

  while (cur->valid) {
    const bool warm = cur->warm != 0;
    const bool add_tail = cur->first == 0;
    const bool in_fast = cur->in_fast != 0, out_fast = cur->out_fast != 0;
    Regs v;
    if (in_fast) {
      int sh[4] = {cur->in_shift[0], cur->in_shift[1], cur->in_shift[2], cur->in_shift[3]};
      while (!dev::mbar_try_wait(mbar, in_phase)) {
      }
      in_phase ^= 1;
      fast_load(t, v, in_stage, LS, sh, masks.in_zero);
    } else {
      Lane ln[4] = {cur->ln[0], cur->ln[1], cur->ln[2], cur->ln[3]};
      io_load(t, v, p, ln, cur->j);
    }
    phase0(t, v, xb, tw1);
    dev::bar_sync(bar_id);
    if (t == 0) {
      if (out_pending) flush_out(nxt, out_stage, LS, p.step); // nxt still holds the previous packet
      schedule_next(p, it, cur, nxt, in_stage, LS, mbar);
    }
    phase1(t, v, xb, tw2);
    dev::bar_sync(bar_id);
    phase2(t, v, xb, hs);
    dev::bar_sync(bar_id);
    phase3(t, v, xb, tw2);
    if (t == 0 && out_pending) dev::bulk_wait_read0(); // staging buffer is free again
    dev::bar_sync(bar_id);
    phase4(t, v, xb, tw1);
    pt4 *tail_prev = tails + (parity ^ 1) * tpts, *tail_next = tails + parity * tpts;
    if (out_fast || warm) {
      int sh[4] = {cur->out_shift[0], cur->out_shift[1], cur->out_shift[2], cur->out_shift[3]};
      fast_store(t, v, out_stage, LS, sh, masks, tail_prev, tail_next, add_tail, out_fast, p.step);
      if (out_fast) dev::fence_proxy_async(); // make the staged values visible to the bulk-copy engine
    } else {
      Lane ln[4] = {cur->ln[0], cur->ln[1], cur->ln[2], cur->ln[3]};
      io_store(t, v, p, ln, cur->j, tail_prev, tail_next, add_tail, warm);
    }
    out_pending = out_fast;
    parity ^= 1;
    Packet *tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  dev::bar_sync(bar_id);
  if (t == 0) {
    if (out_pending) flush_out(nxt, out_stage, LS, p.step);
    dev::bulk_wait0();
  }
}
#endif // __CUDACC__

} // namespace n2048
} // namespace fc
