// oa_w1k_kernel.cuh -- the sm_100a kernel of the warp-per-FFT overlap-add engine (float32,
// 1024-point blocks; oa_w1k_core.cuh has the method, oa_w1k_group.cuh the work decomposition).
//
// Persistent grid, one CTA of kWarps warps per SM; every warp is on its own: it owns an exchange
// buffer of 9216 bytes (which also receives its input blocks), one mbarrier, and a contiguous range of
// units.  Per unit:
//   mbarrier wait | 64 x LDS.32 (staged inputs) | warp sync
//   P0: radix-2 step (scalar), dft16 (packed), twiddles | 64 x STS.32 | warp sync | 16 x LDS.128
//   P1: dft16, radix-2 step, spectrum multiply, radix-2 step, dft16 | 16 x STS.128 | warp sync | 64 x LDS.32
//   warp sync | lane 0 requests the next unit's two input blocks (cp.async.bulk -> the buffer, idle now)
//   P2: twiddles, dft16, radix-2 step | 64 coalesced 4-byte stores
// There is no block-level barrier after the prologue.
#pragma once

#include "oa_n2048_kernel.cuh" // the PTX wrappers (namespace fc::n2048::dev)
#include "oa_w1k_group.cuh"

namespace fc {
namespace w1k {

constexpr int kWarps = 16;
constexpr int kTwCount = kTwRows * 32;
// shared memory: spectrum (512 x 32 B), lane twiddles, one mbarrier per warp, then the warps' buffers
constexpr size_t kSmemHs = 1024 * sizeof(Quad); // folded spectrum tables (oa_w1k_core.cuh)
constexpr size_t kSmemTw = kTwCount * sizeof(cpair);
constexpr size_t kSmemBar = 256; // up to 32 mbarriers
constexpr size_t smem_bytes(int warps = kWarps) {
  return kSmemHs + kSmemTw + kSmemBar + (size_t)warps * kWarpFloats * sizeof(float);
}

#if defined(__CUDACC__)
namespace dev = fc::n2048::dev;

// the general description of a unit, out of line: it is needed for the last block of a row and other
// rarities only, and inlined three times it would push the steady-state code out of the instruction cache
__device__ __noinline__ UnitDesc describe_unit_cold(const W1kParams &p, int row0, int seg0, int row1, int seg1) {
  return describe_unit_at(p, row0, seg0, row1, seg1);
}

struct Pos { // (row, block) of the two items of a unit; the engine takes batch, blocks per row < 2^31
  int row[2], seg[2];
};
// (row, block) of a unit's two items from its index: 64-bit divisions, out of line (pool units only)
__device__ __noinline__ Pos position_of_unit_cold(long long u, long long segs) {
  Pos q;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const long long it = 2 * u + i;
    q.row[i] = (int)(it / segs);
    q.seg[i] = (int)(it - (long long)q.row[i] * segs);
  }
  return q;
}
__device__ __forceinline__ void advance(Pos &q, int segs) { // to the next unit: two items further
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    q.seg[i] += 2;
    while (q.seg[i] >= segs) {
      q.seg[i] -= segs;
      ++q.row[i];
    }
  }
}
// What a unit needs when its inputs are read and its outputs stored; worked out one unit ahead, when
// its input blocks are requested.
struct Carry { // packed into one register: the state lives across a whole iteration
  // bit 0: bulk, bits 1-2: fast_out of item 0 / 1, bits 3-4 / 5-6: shifts, bits 7-17 / 18-28: zero_from of item 0 / 1
  unsigned code;
  __device__ __forceinline__ bool bulk() const { return code & 1u; }
  __device__ __forceinline__ bool fast_out(int i) const { return (code >> (1 + i)) & 1u; }
  __device__ __forceinline__ int sh_a() const { return (int)((code >> 3) & 3u); }
  __device__ __forceinline__ int sh_b() const { return (int)((code >> 5) & 3u); }
  __device__ __forceinline__ int zero_from(int i) const { return (int)((code >> (7 + 11 * i)) & 2047u); }
};

// SHARED_DFT: one copy of each direction's dft16 in a two-trip loop (see oa_w1k_core.cuh)
template <int WARPS, bool SHARED_DFT, bool FULL_TW = false>
__global__ void __launch_bounds__(32 * WARPS, 1) oa_w1k_kernel(const __grid_constant__ W1kParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const SpectrumTables hs4{reinterpret_cast<Quad *>(smem_raw)};
  cpair *tw = reinterpret_cast<cpair *>(smem_raw + kSmemHs);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw + kSmemHs + kSmemTw);
  float *xb_all = reinterpret_cast<float *>(smem_raw + kSmemHs + kSmemTw + kSmemBar);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *xb = xb_all + warp * kWarpFloats;
  const uint32_t mbar = dev::smem_u32(bars + warp);

  if (p.hs) { // a filter handle: the spectrum is resident
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
      fold_spectrum_entry(p.tw[tw_index(kTwRowCross + (i >> 5), 0)], p.hs[2 * i], p.hs[2 * i + 1], hs4.q0[i], hs4.q1()[i]);
    }
  }
  for (int i = threadIdx.x; i < kTwCount; i += blockDim.x) tw[i] = p.tw[i];
  if (lane == 0) {
    dev::mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // The spectrum is needed for the first time in the middle of a warp's first unit.  The last warp transforms the
  // taps (oa_w1k_group.cuh: spectrum_*) and ARRIVES at a named barrier; the other warps start their first unit at
  // once -- its bulk copies are in flight while the taps are transformed -- and wait at that barrier just before
  // their first spectrum multiply (a warp without units only arrives).
  constexpr int kSpectrumBarrier = 1;
  bool wait_spectrum = false;
  if (!p.hs) {
    if (warp == WARPS - 1) {
      spectrum_stage(lane, p.taps, p.klen, xb);
      __syncwarp();
      Regs v;
      load_inputs(lane, v, xb, 0, 0);
      __syncwarp();
      phase0(lane, v, tw);
      t1_write(lane, v, xb);
      __syncwarp();
      t1_read(lane, v, xb);
      spectrum_finish(lane, v, hs4);
      __syncwarp(); // its own first request reuses the buffer
      asm volatile("bar.arrive %0, %1;" ::"n"(kSpectrumBarrier), "n"(32 * WARPS) : "memory");
    } else {
      wait_spectrum = true;
    }
  }
  const long long n_warps = (long long)gridDim.x * WARPS;
  long long u0, u1;
  // Ranges are dealt out warp-major (range index = warp * CTAs + CTA): the n_units mod n_warps ranges that are one
  // unit longer then sit in the first warps of EVERY CTA instead of filling the first CTAs -- the warps of an SM
  // share its pipes, so what counts is units per SM (512 rows: 148 on the busiest SM instead of 160) -- and the
  // last warp, which transformed the taps, is the last to get a longer range.
  const long long range_id = p.cta_major ? (long long)blockIdx.x * WARPS + warp : (long long)warp * gridDim.x + blockIdx.x;
  const bool dynamic = p.sched != nullptr; // fixed ranges of static_per_warp units, then single units from the pool
  if (dynamic) {
    u0 = range_id * p.static_per_warp;
    u1 = u0 + p.static_per_warp;
  } else {
    warp_range(p.n_units, n_warps, range_id, u0, u1);
  }
  // a warp's last act in a dynamic launch: count itself out; the last one of the grid rewinds the pool counter
  // (atomicInc wraps the count of finished warps to zero by itself), so the next launch finds both at zero
  auto sign_off = [&]() { // (nothing of the set-up above is kept alive for this)
    if (p.sched != nullptr && (threadIdx.x & 31) == 0) {
      const unsigned last = gridDim.x * WARPS - 1;
      __threadfence();
      if (atomicInc(p.sched + 1, last) == last) p.sched[0] = 0u;
    }
  };
  if (u0 >= u1) {
    if (wait_spectrum) asm volatile("bar.arrive %0, %1;" ::"n"(kSpectrumBarrier), "n"(32 * WARPS) : "memory");
    sign_off();
    return;
  }
  StoreMasks masks = make_store_masks(lane, p);
  const int segs = (int)p.segs, step = p.step, tail = p.klen - 1;
  const float *tensor_end = p.in + (p.batch - 1) * p.in_stride + p.n;

  Pos q;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const long long it = 2 * u0 + i;
    q.row[i] = (int)(it / p.segs);
    q.seg[i] = (int)(it - (long long)q.row[i] * p.segs);
  }
  // Request the input blocks of the unit at `at` (the buffer must be idle) and work out what the unit will need
  // later (oa_w1k_group.cuh: item_request).
  // Interior blocks -- not a row's first, and ending at least three samples before the row's end, so that the
  // 16-byte rounding of the copy stays inside the row -- need none of item_request's case analysis: the copy is
  // kN samples (+ 16 bytes when the block start is not 16-byte aligned), nothing is zeroed, all `step` outputs lie
  // inside the row (oa_w1k_group.cuh: interior_seg_hi has the argument).  83 of config 2's 85 blocks per row.
  const int seg_hi = interior_seg_hi(p);
  const int batch_i = (int)p.batch;
  auto request = [&](const Pos &at) -> Carry {
    Carry c;
    if (at.seg[0] >= 1 && at.seg[0] <= seg_hi && at.seg[1] >= 1 && at.seg[1] <= seg_hi && at.row[1] < batch_i) {
      const float *f0 = p.in + ((long long)at.row[0] * p.in_stride + ((long long)at.seg[0] * step - tail));
      const float *f1 = p.in + ((long long)at.row[1] * p.in_stride + ((long long)at.seg[1] * step - tail));
      const unsigned i0 = (unsigned)align_shift4(f0), i1 = (unsigned)align_shift4(f1);
      c.code = 7u | (i0 << 3) | (i1 << 5) | ((unsigned)(kN + 4) << 7) | ((unsigned)(kN + 4) << 18);
      if (lane == 0) {
        const uint32_t b0 = 4u * kN + (i0 ? 16u : 0u), b1 = 4u * kN + (i1 ? 16u : 0u);
        dev::fence_proxy_async(); // the buffer was last touched through the generic proxy
        dev::mbar_expect_tx(mbar, b0 + b1);
        dev::bulk_load(dev::smem_u32(xb), f0 - i0, b0, mbar);
        dev::bulk_load(dev::smem_u32(xb + kPlane), f1 - i1, b1, mbar);
      }
      return c;
    }
    const ItemReq r0 = item_request(p, at.row[0], at.seg[0], tensor_end);
    const ItemReq r1 = item_request(p, at.row[1], at.seg[1], tensor_end);
    const bool in_ok = r0.in_ok && r1.in_ok;
    // a unit with a block that cannot be copied in bulk (misaligned row starts, the tensor's end, the odd last
    // item) is staged by the lanes at the start of its iteration instead (manual_fill): nothing to request now
    c.code = (in_ok ? 1u : 0u) | (r0.out_ok ? 2u : 0u) | (r1.out_ok ? 4u : 0u) | ((unsigned)r0.shift << 3) |
             ((unsigned)r1.shift << 5) | ((unsigned)r0.zero_from << 7) | ((unsigned)r1.zero_from << 18);
    if (in_ok) {
      if (lane == 0) {
        dev::fence_proxy_async(); // the buffer was last touched through the generic proxy
        dev::mbar_expect_tx(mbar, (uint32_t)(r0.bytes + r1.bytes));
        dev::bulk_load(dev::smem_u32(xb + r0.dst_first), r0.src, (uint32_t)r0.bytes, mbar);
        dev::bulk_load(dev::smem_u32(xb + kPlane + r1.dst_first), r1.src, (uint32_t)r1.bytes, mbar);
      }
      for (int i = lane; i < r0.head_zero; i += 32) xb[i] = 0.0f;
      for (int i = lane; i < r1.head_zero; i += 32) xb[kPlane + i] = 0.0f;
    }
    return c;
  };

  Carry nxt = request(q);
  uint32_t phase = wait_spectrum ? 2u : 0u; // bit 0: the mbarrier's phase parity; bit 1: the spectrum barrier is still ahead

  int static_left = (int)(u1 - u0) - 1;    // units of the fixed range after the current one
  unsigned fetched = 0;                 // lane 0: the pool unit the counter handed out
#pragma unroll 1
  for (;;) {
    // the unit after this one comes from the pool: ask for it now, the answer is needed half a unit from here
    const bool from_pool = p.sched != nullptr && static_left == 0;
    if (from_pool && lane == 0) fetched = atomicAdd(p.sched, 1u);
    const Carry cur = nxt;
    const Pos at = q;
    Regs v;
    bool has_next = false;
    // moves q to the unit after this one; false: there is none
    auto next_unit = [&]() -> bool {
      if (!from_pool) {
        if (static_left <= 0) return false; // fixed ranges only, and this was the last unit
        --static_left;
        advance(q, segs);
        return true;
      }
      const long long nu = p.pool_first + (long long)__shfl_sync(0xffffffffu, fetched, 0);
      if (nu >= p.n_units) return false;
      q = position_of_unit_cold(nu, p.segs);
      return true;
    };
    int sh_a = cur.sh_a(), sh_b = cur.sh_b();
    if (cur.bulk()) {
      while (!dev::mbar_try_wait(mbar, phase & 1u)) {
      }
      phase ^= 1;
      // blocks clipped at their row's end: zeros behind the copied samples
      zero_tail(lane, cur.zero_from(0), xb);
      zero_tail(lane, cur.zero_from(1), xb + kPlane);
    } else { // no data registers are live here: the out-of-line call costs no spills
      const UnitDesc d = describe_unit_cold(p, at.row[0], at.seg[0], at.row[1], at.seg[1]);
      manual_fill(lane, p, d.item[0], xb);
      manual_fill(lane, p, d.item[1], xb + kPlane);
      sh_a = d.item[0].shift;
      sh_b = d.item[1].shift;
    }
    __syncwarp(); // the lanes' own fills (zeros, manual) are visible to the whole warp
    load_inputs(lane, v, xb, sh_a, sh_b);
    __syncwarp(); // every lane has its inputs: the buffer turns into the exchange buffer
    if constexpr (SHARED_DFT) {
      p0_pre(lane, v, tw);
#pragma unroll 1
      for (int trip = 0; trip < 2; ++trip) {
        dft16<-1>(v);
        if (trip == 0) {
          p0_post<FULL_TW>(lane, v, tw);
          t1_write(lane, v, xb);
          __syncwarp();
          t1_read(lane, v, xb);
        } else {
          if (phase & 2u) { // first unit only
            asm volatile("bar.sync %0, %1;" ::"n"(kSpectrumBarrier), "n"(32 * WARPS) : "memory");
            phase &= 1u;
          }
          p1_mid(lane, v, hs4);
        }
      }
#pragma unroll 1
      for (int trip = 0; trip < 2; ++trip) {
        dft16<+1>(v);
        if (trip == 0) {
          t2_write(lane, v, xb); // a lane's own row, which only it read in T1: no barrier needed before
          __syncwarp();
          t2_read(lane, v, xb);
          __syncwarp(); // the buffer is idle: the next inputs may land while P2 and the stores run
          has_next = next_unit();
          if (has_next) nxt = request(q);
          p2_pre<FULL_TW>(lane, v, tw);
        } else {
          p2_post(lane, v, tw);
        }
      }
    } else {
      phase0(lane, v, tw);
      t1_write(lane, v, xb);
      __syncwarp();
      t1_read(lane, v, xb);
      if (phase & 2u) {
        asm volatile("bar.sync %0, %1;" ::"n"(kSpectrumBarrier), "n"(32 * WARPS) : "memory");
        phase &= 1u;
      }
      phase1(lane, v, hs4);
      t2_write(lane, v, xb); // a lane's own row, which only it read in T1: no barrier needed before
      __syncwarp();
      t2_read(lane, v, xb);
      __syncwarp(); // the buffer is idle: the next inputs may land while P2 and the stores run
      has_next = next_unit();
      if (has_next) nxt = request(q);
      phase2(lane, v, tw);
    }
    asm volatile("" : "+r"(masks.keep_from)); // not loop-invariant as far as the compiler knows: no hoisted predicates
    // where y[0] of each item goes: row's output + block start - pad
    if (cur.fast_out(0))
      store_item_fast<0>(lane, v, p.out + (long long)at.row[0] * p.out_stride + ((long long)at.seg[0] * step - tail - p.pad), masks);
    else
      store_item_slow<0>(lane, v, p, at.row[0], at.seg[0]);
    if (cur.fast_out(1))
      store_item_fast<1>(lane, v, p.out + (long long)at.row[1] * p.out_stride + ((long long)at.seg[1] * step - tail - p.pad), masks);
    else
      store_item_slow<1>(lane, v, p, at.row[1], at.seg[1]);
    if (!has_next) break;
  }
  sign_off();
}
#endif // __CUDACC__

} // namespace w1k
} // namespace fc
