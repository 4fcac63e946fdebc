// oa_w2_group.cuh -- I/O phases of the warp-pair engine (oa_w2_core.cuh).  The work
// decomposition (streams, quads, warm-up segments, packets) is the one of
// oa_n2048_group.cuh and is reused unchanged; only the thread <-> sample mapping differs:
// thread t of warp p consumes z[t + 32 m] and z[t + 32 m + 1024] for m = 0..31 (both warps
// read every input) and produces y[t + 32 (j + 16 p)] and y[t + 32 (j + 16 p) + 1024], j = 0..15.
//
// Reference semantics being replaced: FFTConvEngine::oaconvolve
// (/root/reference/include/fftconv/fftconv.hpp:382-437).
#pragma once

#include "oa_n2048_group.cuh"
#include "oa_w2_core.cuh"

namespace fc {
namespace w2 {

using n2048::all4;
using n2048::align_shift;
using n2048::fill_packet_lane;
using n2048::Lane;
using n2048::OaParams;
using n2048::Packet;
using n2048::SchedSlot;
using n2048::tail_pts;

// lane 0 -> re, half A    lane 1 -> im, half A    lane 2 -> re, half B    lane 3 -> im, half B
// radix-2 stage: u = z[n] + sgn z[n + 1024], sgn = +1 (p = 0) or -1 (p = 1)
FC_HD void radix2_point(Regs &v, int m, const float (&lo)[4], const float (&hi)[4], bool hi_zero, float sgn) {
  const f2 lr = make_f2(lo[0], lo[2]), li = make_f2(lo[1], lo[3]);
  if (hi_zero) {
    v.re[m] = lr;
    v.im[m] = li;
  } else {
    v.re[m] = fmas2(make_f2(hi[0], hi[2]), sgn, lr);
    v.im[m] = fmas2(make_f2(hi[1], hi[3]), sgn, li);
  }
}

// step >= 2048 - 409 = 1639: z[1024 + 32 m + t] is certainly data for m <= 18 and certainly
// padding for m >= 26 (step <= 2048 - 220 = 1828 < 1024 + 32 * 26).
constexpr int kHiAlways = 19, kHiNever = 26;

// ---- fast path: staged through shared memory (4 lanes of `ls` floats) --------------------
FC_HD void fast_load(int t, Regs &v, const float *in, int ls, const int (&shift)[4], int step, float sgn) {
  const float *b0 = in + 0 * ls + shift[0] + t;
  const float *b1 = in + 1 * ls + shift[1] + t;
  const float *b2 = in + 2 * ls + shift[2] + t;
  const float *b3 = in + 3 * ls + shift[3] + t;
  const int lim = step - kHalf - t; // z[1024 + 32 m + t] is data iff 32 m < lim
#pragma unroll
  for (int m = 0; m < kPts; ++m) {
    const float lo[4] = {b0[32 * m], b1[32 * m], b2[32 * m], b3[32 * m]};
    float hi[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < kHiNever) {
      const bool on = (m < kHiAlways) || (32 * m < lim);
      if (on) {
        hi[0] = b0[kHalf + 32 * m];
        hi[1] = b1[kHalf + 32 * m];
        hi[2] = b2[kHalf + 32 * m];
        hi[3] = b3[kHalf + 32 * m];
      }
    }
    radix2_point(v, m, lo, hi, m >= kHiNever, sgn);
  }
}

// ---- slow path: direct, predicated global access ----------------------------------------
FC_HD void io_load(int t, Regs &v, const OaParams &p, const Lane (&ln)[4], long long j, float sgn) {
  const float *src[4];
  int ilen[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const long long seg = ln[l].seg0 + j;
    const bool on = ln[l].active && seg >= 0 && j < ln[l].cnt;
    const long long pos = seg * p.step;
    long long len = on ? p.n - pos : 0;
    if (len > p.step) len = p.step;
    src[l] = ln[l].in + (on ? pos : 0);
    ilen[l] = (int)len;
  }
#pragma unroll
  for (int m = 0; m < kPts; ++m) {
    float lo[4], hi[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int i = t + 32 * m;
#if defined(__CUDA_ARCH__)
      lo[l] = (i < ilen[l]) ? __ldg(src[l] + i) : 0.0f;
      hi[l] = (m < kHiNever && i + kHalf < ilen[l]) ? __ldg(src[l] + i + kHalf) : 0.0f;
#else
      lo[l] = (i < ilen[l]) ? src[l][i] : 0.0f;
      hi[l] = (m < kHiNever && i + kHalf < ilen[l]) ? src[l][i + kHalf] : 0.0f;
#endif
    }
    radix2_point(v, m, lo, hi, m >= kHiNever, sgn);
  }
}

// registers -> the four lanes' outputs (+ tail hand-over).  b_l points at y[t + 512 p] of lane l
// in global memory (one coalesced 128-byte line per warp instruction); the warp owns
// y[t + 32 j + 512 p] and y[t + 32 j + 512 p + 1024], j = 0..15.
FC_HD void fast_store(int t, int p, const Out &o, float *b0, float *b1, float *b2, float *b3, int klen, int step,
                      const pt4 *tail_prev, pt4 *tail_next, bool add_tail, bool stage) {
  // p = 0: y[t + 32 j] takes an incoming tail iff 32 j < tail_lim
  const int tail_lim = (p == 0 && add_tail) ? klen - 1 - t : 0;
  // p = 1: y[1536 + 32 j + t] is stored iff 32 j < own_lim, else it belongs to the outgoing tail
  const int own_lim = p ? step - (kHalf + 512) - t : 1 << 20;
  pt4 *tail_dst = tail_next + (t + kHalf + 512 - step); // p = 1 only
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    f2 yr = o.lo_re[j], yi = o.lo_im[j];
    if (j < 13) { // K - 1 <= 409 < 416
      if (32 * j < tail_lim) {
        f2 tr, ti;
        get(tail_prev, t + 32 * j, tr, ti);
        yr = add2(yr, tr);
        yi = add2(yi, ti);
      }
    }
    if (stage) {
      b0[32 * j] = yr.x;
      b1[32 * j] = yi.x;
      b2[32 * j] = yr.y;
      b3[32 * j] = yi.y;
    }
    yr = o.hi_re[j];
    yi = o.hi_im[j];
    const bool in_tail = j >= 3 && !(32 * j < own_lim); // 1536 + 96 = 1632 < 1639 <= step
    if (in_tail) {
      put(tail_dst, 32 * j, yr, yi);
    } else if (stage) {
      b0[kHalf + 32 * j] = yr.x;
      b1[kHalf + 32 * j] = yi.x;
      b2[kHalf + 32 * j] = yr.y;
      b3[kHalf + 32 * j] = yi.y;
    }
  }
}

FC_HD void io_store(int t, int p_warp, const Out &o, const OaParams &p, const Lane (&ln)[4], long long j,
                    const pt4 *tail_prev, pt4 *tail_next, bool add_tail, bool warm) {
  int limit[4];
  long long obase[4]; // output index of y[0]
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const long long seg = ln[l].seg0 + j;
    const bool on = !warm && ln[l].active && j < ln[l].cnt;
    const long long pos = seg * p.step;
    long long lim = p.step;
    if (seg == p.segs - 1) { // the row's final segment also owns its tail
      lim = p.out_len + p.pad - pos;
      if (lim > kN) lim = kN;
    }
    limit[l] = on ? (int)lim : 0;
    obase[l] = pos - p.pad;
  }
  const int tail_len = p.klen - 1;
#pragma unroll
  for (int jj = 0; jj < 32; ++jj) {
    const int i = t + 32 * (jj & 15) + 512 * p_warp + ((jj >> 4) ? kHalf : 0);
    f2 yr = (jj >> 4) ? o.hi_re[jj & 15] : o.lo_re[jj & 15];
    f2 yi = (jj >> 4) ? o.hi_im[jj & 15] : o.lo_im[jj & 15];
    if (add_tail && i < tail_len) {
      f2 tr, ti;
      get(tail_prev, i, tr, ti);
      yr = add2(yr, tr);
      yi = add2(yi, ti);
    }
    if (i >= p.step) put(tail_next, i - p.step, yr, yi);
    const float y[4] = {yr.x, yi.x, yr.y, yi.y};
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const long long oi = obase[l] + i;
      if (i < limit[l] && oi >= 0) ln[l].out[oi] = y[l];
    }
  }
}

// ---- shared-memory plan of one warp pair ---------------------------------------------------
struct PairLayout {
  int lane_stride;     // floats per lane of the input staging area (which aliases xb0|xb1)
  int xb, tails, packets, sched, mbar, total;
};
FC_HD PairLayout pair_layout(int klen) {
  PairLayout g;
  const int step = kN - (klen - 1);
  g.lane_stride = (3 + step + 3) & ~3;
  int off = 0;
  g.xb = off;      off += 2 * kXbPts * (int)sizeof(pt4);          // 33792 B >= 4 * 1832 * 4
  g.tails = off;   off += 2 * tail_pts(klen) * (int)sizeof(pt4);
  g.packets = off; off += 3 * (((int)sizeof(Packet) + 15) & ~15);
  g.sched = off;   off += 4 * (int)sizeof(SchedSlot);
  g.mbar = off;    off += 16;
  g.total = (off + 127) & ~127;
  return g;
}
inline size_t smem_bytes(int pairs, int klen) {
  return (size_t)(2048 + kTwElems) * sizeof(cpair) + (size_t)pairs * pair_layout(klen).total;
}

} // namespace w2
} // namespace fc
