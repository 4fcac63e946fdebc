// oa_n2048_kernel.cuh -- the sm_100a overlap-add kernel for float32, FFT size 2048.
//
// Persistent grid: one CTA per SM, G groups of 128 threads per CTA (G = 4: 16
// warps per SM, 128 registers per thread).  Every group owns a slice of `budget`
// consecutive blocks of the (stream-quad, block) work list; where its slice
// starts comes from the plan's table (oa_n2048_tables.h: plan_slices).
//
// Shared memory is the scarce resource (the 32 KB exchange buffer per group is
// irreducible), so the global <-> shared staging areas ALIAS the exchange
// buffer and the group's time line per segment quad is
//   mbarrier wait      four input segments have landed in xb (cp.async.bulk,
//                      TMA, issued by thread 0 at the end of the previous
//                      iteration)
//   LDS.32 x 64        x[t + 128 m] of the four lanes -> packed registers
//   P0 compute | bar | P0 put | bar | P1 | warp | P2 | warp | P3 | bar | P4 reads | bar
//   thread 0           requests the NEXT inputs (the buffer is idle from here on)
//                      and then builds the packet of the iteration after next
//   P4 DFT, tail hand-over (K-1 samples, separate small ping-pong buffers),
//   64 coalesced 4-byte stores per thread straight from registers
// While one group waits for its copies the other three groups keep the FP32
// pipe busy.  Thread 0 of each group is the group's scheduler: the other 127
// threads execute no 64-bit control arithmetic at all.
//
// Shared memory per CTA: hs 16 KB + tw1 16 KB + tw2 1 KB, then per group the
// 32 KB exchange buffer, two (K-1)-point tail buffers, two packets, the
// scheduler state and one mbarrier (~40.6 KB for K = 245).
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

// Thread 0: start the bulk copies of a fast-path packet into the staging area
// (= the exchange buffer, which must be idle).
template <typename S>
__device__ __forceinline__ void issue_loads(const PacketT<S> *pk, S *stage, int lane_stride, uint32_t mbar) {
  constexpr int NL = OaParamsT<S>::NL;
  if (pk->valid && all4(pk->in_fast)) {
    dev::fence_proxy_async(); // the area was last touched through the generic proxy
    uint32_t bytes = 0;
#pragma unroll
    for (int l = 0; l < NL; ++l) bytes += (uint32_t)pk->in_bytes[l];
    dev::mbar_expect_tx(mbar, bytes);
#pragma unroll
    for (int l = 0; l < NL; ++l)
      dev::bulk_load(dev::smem_u32(stage + l * lane_stride), pk->in_src[l], (uint32_t)pk->in_bytes[l], mbar);
  }
}

// D = f2 (float32, four segments per group iteration) or double (float64, two).
template <typename D, int G, bool RECOMP, bool WIDE = false>
__global__ void __launch_bounds__(128 * G, 1) oa_n2048_kernel(const __grid_constant__ OaParamsT<scalar_of_t<D>> p) {
  using S = scalar_of_t<D>;
  using cpair = CPairT<S>;
  using Packet = PacketT<S>;
  using SchedSlot = SchedSlotT<S>;
  using Lane = LaneT<S>;
  using Regs = RegsT<D>;
  using pt = PtT<D>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cpair *hs = reinterpret_cast<cpair *>(smem_raw);
  cpair *tw1 = hs + 2048;                         // RECOMP: rows k2 = 1, 2, 4, 8 only
  cpair *tw2 = tw1 + tw1_rows(RECOMP) * 128;
  unsigned char *grp0 = reinterpret_cast<unsigned char *>(tw2 + 128);

  const GroupLayout L = group_layout<S>(p.klen);
  const int g = threadIdx.x >> 7, t = threadIdx.x & 127;
  unsigned char *base = grp0 + (size_t)g * L.total;
  pt *xb = reinterpret_cast<pt *>(base + L.xb);
  S *stage = reinterpret_cast<S *>(base + L.xb); // input staging aliases xb
  pt *tails = reinterpret_cast<pt *>(base + L.tails);
  const int pk_stride = ((int)sizeof(Packet) + 15) & ~15;
  // ring of three packets, addressed arithmetically so that every access stays a
  // shared-memory access (an array of pointers would live in local memory)
  unsigned char *pk_base = base + L.packets;
  auto PK = [&](int i) -> Packet * { return reinterpret_cast<Packet *>(pk_base + i * pk_stride); };
  const uint32_t mbar = dev::smem_u32(base + L.mbar);
  const int tpts = tail_pts(p.klen);
  const int LS = L.lane_stride;

  for (int i = threadIdx.x; i < 2048; i += blockDim.x) hs[i] = p.hs[i];
  for (int i = threadIdx.x; i < tw1_rows(RECOMP) * 128; i += blockDim.x)
    tw1[i] = RECOMP ? p.tw1[(128 << (i >> 7)) + (i & 127)] : p.tw1[i];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) tw2[i] = p.tw2[i];

  const long long slot = (long long)blockIdx.x * G + g;
  const SlotStart start = slot < p.n_slots ? p.slots[slot] : SlotStart{p.n_quads, 0, 0};

  // the four warp leaders of a group are its scheduler: leader w describes lane w
  const bool leader = (t & 31) == 0;
  const int my_lane = t >> 5;
  SchedSlot &sslot = reinterpret_cast<SchedSlot *>(base + L.sched)[my_lane];
  if (leader) {
    sslot.it.init(start.quad, start.u, (int)start.budget);
    fill_packet_lane(p, sslot, PK(0), my_lane); // iteration 0
    fill_packet_lane(p, sslot, PK(1), my_lane); // iteration 1
  }
  if (t == 0) {
    dev::mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (t == 0) issue_loads(PK(0), stage, LS, mbar);

  const ThreadMasks masks = make_masks(t, p);
  const int bar_id = g + 1;
  uint32_t in_phase = 0;
  int parity = 0;
  int ci = 0; // ring index of the current packet

  while (PK(ci)->valid) {
    const Packet *cur = PK(ci);
    const bool skip = cur->skip != 0;
    const bool add_tail = cur->first == 0;
    const bool in_fast = all4(cur->in_fast), out_fast = all4(cur->out_fast);
    Regs v;
    if (in_fast) {
      int sh[4] = {cur->in_shift[0], cur->in_shift[1], cur->in_shift[2], cur->in_shift[3]};
      while (!dev::mbar_try_wait(mbar, in_phase)) {
      }
      in_phase ^= 1;
      fast_load<WIDE>(t, v, stage, LS, sh, masks.in_zero);
    } else {
      Lane ln[4] = {cur->ln[0], cur->ln[1], cur->ln[2], cur->ln[3]};
      io_load(t, v, p, ln, cur->u);
    }
    phase0_compute<RECOMP, RECOMP>(t, v, tw1);
    dev::bar_sync(bar_id); // every thread has read its inputs out of the staging area
    phase0_put(t, v, xb);
    dev::bar_sync(bar_id);
    // P1 -> P2 -> P3 exchange inside each warp's own 512 points (oa_n2048_core.cuh, p2_col): a warp
    // barrier suffices.  Measured: float32 2 % faster with it (the warps of a group drift apart and
    // overlap), float64 2 % slower (0.55 -> 0.56 ms at 2048 x 65536) -- that one keeps the group barrier.
    auto mid_sync = [&]() {
      if constexpr (sizeof(S) == 4) __syncwarp();
      else dev::bar_sync(bar_id);
    };
    phase1<RECOMP>(t, v, xb, tw2);
    mid_sync();
    phase2(t, v, xb, hs);
    mid_sync();
    phase3<RECOMP>(t, v, xb, tw2);
    dev::bar_sync(bar_id);
    phase4_read<RECOMP, RECOMP>(t, v, xb, tw1);
    dev::bar_sync(bar_id); // every thread has read its points: the exchange buffer is idle
    const int ni = ci == 2 ? 0 : ci + 1, nni = ni == 2 ? 0 : ni + 1;
    // request the next inputs now: they land while this iteration's last DFT and stores run
    if (t == 0) issue_loads(PK(ni), stage, LS, mbar);
    phase4_compute(v);
    pt *tail_prev = tails + (parity ^ 1) * tpts, *tail_next = tails + parity * tpts;
    if (out_fast) {
      // coalesced element stores straight from registers (full 128-byte lines per warp instruction)
      S *const dst[4] = {cur->out_dst[0] + t, cur->out_dst[1] + t, cur->out_dst[2] + t, cur->out_dst[3] + t};
      fast_store_to<WIDE>(t, v, dst, masks, tail_prev, tail_next, add_tail, skip, p.step);
    } else {
      Lane ln[4] = {cur->ln[0], cur->ln[1], cur->ln[2], cur->ln[3]};
      io_store(t, v, p, ln, cur->u, tail_prev, tail_next, add_tail, skip);
    }
    // describe the iteration after next, one lane per warp leader (no warp is later
    // than the others, and the copies requested above are still in flight)
    if (leader && PK(ni)->valid) fill_packet_lane(p, sslot, PK(nni), my_lane);
    parity ^= 1;
    ci = ni;
  }
}
#endif // __CUDACC__

} // namespace n2048
} // namespace fc
