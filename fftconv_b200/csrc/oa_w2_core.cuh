// oa_w2_core.cuh -- second-generation fused overlap-add segment engine for FFT size
// 2048, float32 (the reference's segment size for 221 <= K < 411 taps:
// /root/reference/include/fftconv/fftconv.hpp:38-66).  Replaces, per segment, the
// reference's copy_to_padded_buffer + r2c + multiply_cx + c2r + normalize_add
// (fftconv.hpp:400-410 / :417-432).
//
// Why a second engine: oa_n2048_core.cuh (16 points per thread, 2048 = 16 x 16 x 8)
// needs FOUR full passes of the data through shared memory per segment quad and is
// bound by the shared-memory crossbar.  Here a thread holds 32
// points and a WARP owns a whole 1024-point transform, so the data crosses the
// crossbar twice (one transpose per direction) plus one half-size hand-over:
//
//   A "pair" of two warps transforms FOUR real segments (two ride in re/im of one
//   complex FFT, two such FFTs ride in the halves of every f32x2 register).
//   2048 = 2 x (32 x 32), decimation in frequency, z[n] = a[n] + i b[n]:
//     warp p = 0:  u[n] = z[n] + z[n + 1024]                    -> Z[2 k']
//     warp p = 1:  u[n] = (z[n] - z[n + 1024]) W_2048^n         -> Z[2 k' + 1]
//   (both warps read all inputs from the TMA staging area; the radix-2 stage costs no
//   exchange).  Each warp then runs a private 1024-point FFT, n = t + 32 m, k' = k2 + 32 k1:
//     S1 (thread t):   32-point DFT over m -> k2, twiddle W_1024^{t k2} [W_2048^t folded in for p = 1]
//     transpose inside the warp (16.5 KB, __syncwarp only)
//     S2 (thread k2):  32-point DFT over t -> k1, times H[2 k' + p] / 2048,
//                      inverse 32-point DFT back to t
//     transpose back, S1' (thread t): conjugate twiddle, inverse 32-point DFT -> m
//   y[n] = E[n] + W_2048^{-n} O[n], y[n + 1024] = E[n] - W_2048^{-n} O[n]  (E from warp 0,
//   O from warp 1): the warps swap HALF of their points (m >= 16 / m < 16) and each
//   produces 1024 output samples.
//
// W_2048^{t + 32 m} = W_2048^t W_64^m: the W_64^m are compile-time constants, W_2048^t
// is folded into the per-thread twiddle chain.
//
// Everything here is __host__ __device__ so tests can emulate the 64 threads phase
// by phase on the CPU (tests/cpu/emulate_oa_w2.cpp).
#pragma once

#include "oa_n2048_core.cuh"

namespace fc {
namespace w2 {

using n2048::bfly4;
using n2048::cpair;
using n2048::cpair_mul;
using n2048::kC8;
using n2048::pt4;
using n2048::rot90;

constexpr int kN = 2048;
constexpr int kHalf = 1024;
constexpr int kPts = 32;        // points per thread
constexpr int kRow = 33;        // exchange buffer row stride in points (32 + 1 pad: conflict-free both ways)
constexpr int kXbPts = 32 * kRow;

struct Regs {
  f2 re[kPts];
  f2 im[kPts];
};

// cos(2 pi e / 64), sin(2 pi e / 64) as literals (folded after unrolling)
FC_HD constexpr float q64(int e) { // e = 0..16
  switch (e) {
  case 0: return 1.0f;
  case 1: return 9.95184726672196928732e-01f;
  case 2: return 9.80785280403230430579e-01f;
  case 3: return 9.56940335732208824382e-01f;
  case 4: return 9.23879532511286738483e-01f;
  case 5: return 8.81921264348355049556e-01f;
  case 6: return 8.31469612302545235671e-01f;
  case 7: return 7.73010453362736993377e-01f;
  case 8: return 7.07106781186547572737e-01f;
  case 9: return 6.34393284163645487794e-01f;
  case 10: return 5.55570233019602288671e-01f;
  case 11: return 4.71396736825997808573e-01f;
  case 12: return 3.82683432365089837290e-01f;
  case 13: return 2.90284677254462331053e-01f;
  case 14: return 1.95090322016128331351e-01f;
  case 15: return 9.80171403295607701622e-02f;
  default: return 0.0f;
  }
}
FC_HD constexpr float cos64(int e) {
  e &= 63;
  if (e > 32) e = 64 - e;
  return e > 16 ? -q64(32 - e) : q64(e);
}
FC_HD constexpr float sin64(int e) { return cos64(e + 48); }

// (re + i im) *= (wr + i wi): four packed instructions
FC_HD void cmul(f2 &re, f2 &im, float wr, float wi) {
  const f2 t = muls2(im, -wi);
  const f2 u = muls2(im, wr);
  const f2 nre = fmas2(re, wr, t);
  im = fmas2(re, wi, u);
  re = nre;
}

// 8-point DFT on slots B + ST j (j = 0..7), natural order in, output k in slot B + ST slot8(k)
FC_HD constexpr int slot8(int k) { return 2 * (k & 3) + (k >> 2); }
template <int S, int B, int ST> FC_HD void dft8(Regs &v) {
#define FC_W2_R(j) v.re[B + ST * (j)], v.im[B + ST * (j)]
  bfly4<S>(FC_W2_R(0), FC_W2_R(2), FC_W2_R(4), FC_W2_R(6));
  bfly4<S>(FC_W2_R(1), FC_W2_R(3), FC_W2_R(5), FC_W2_R(7));
  const float s = (S < 0) ? -1.0f : 1.0f;
  cmul(FC_W2_R(3), kC8, s * kC8);
  rot90<S>(FC_W2_R(5));
  cmul(FC_W2_R(7), -kC8, s * kC8);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    const f2 ar = v.re[B + ST * (2 * k1)], ai = v.im[B + ST * (2 * k1)];
    const f2 br = v.re[B + ST * (2 * k1 + 1)], bi = v.im[B + ST * (2 * k1 + 1)];
    v.re[B + ST * (2 * k1)] = add2(ar, br);
    v.im[B + ST * (2 * k1)] = add2(ai, bi);
    v.re[B + ST * (2 * k1 + 1)] = sub2(ar, br);
    v.im[B + ST * (2 * k1 + 1)] = sub2(ai, bi);
  }
#undef FC_W2_R
}

// 32-point DFT, sign S, natural order in, output k in slot slot32(k).
//   m = m0 + 4 m1, k = k1 + 8 k0:
//   X[k1 + 8 k0] = sum_m0 W4^{m0 k0} W32^{m0 k1} sum_m1 x[m0 + 4 m1] W8^{m1 k1}
FC_HD constexpr int slot32(int k) { return 4 * slot8(k & 7) + (k >> 3); }
template <int S> FC_HD void dft32(Regs &v) {
  dft8<S, 0, 4>(v);
  dft8<S, 1, 4>(v);
  dft8<S, 2, 4>(v);
  dft8<S, 3, 4>(v); // A[m0][k1] in slot m0 + 4 slot8(k1)
  const float s = (S < 0) ? -1.0f : 1.0f;
#pragma unroll
  for (int m0 = 1; m0 < 4; ++m0)
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) {
      const int e = m0 * k1, sl = m0 + 4 * slot8(k1); // W32^e = W64^{2e}
      if (e == 8) rot90<S>(v.re[sl], v.im[sl]);
      else cmul(v.re[sl], v.im[sl], cos64(2 * e), s * sin64(2 * e));
    }
#pragma unroll
  for (int q = 0; q < 8; ++q)
    bfly4<S>(v.re[4 * q], v.im[4 * q], v.re[4 * q + 1], v.im[4 * q + 1], v.re[4 * q + 2], v.im[4 * q + 2],
             v.re[4 * q + 3], v.im[4 * q + 3]);
}

// exchange buffer of one warp: point (row r, column c) at r * 33 + c
using n2048::get;
using n2048::put;

// The transform of one warp is a short LOOP over four "trips" around ONE copy of the 32-point
// DFT -- the unrolled code of a 32-point-per-thread FFT is large, the SM's instruction cache is
// 32 KB, and a kernel whose hot path does not fit re-fetches every line on every iteration
// (measured: DESIGN.md section 4).  Both warps of a pair run the same instruction stream; p only
// selects tables, a sign and two warp-uniform branches:
//     (p = 1 only)  u[t + 32 m] *= W_64^m                   constants, before the loop
//     trip 1        dft32, chain, put columns, get rows      forward 1024-point DFT, first half
//     trip 2        dft32, times H                           second half; into the "swapped" domain
//     trip 3        dft32, chain, put columns, get rows      inverse (IDFT = swap . DFT . swap), first half
//     trip 4        dft32 [, chain]                          second half [p = 1: times W_64^{m'}]
// "chain" multiplies the point in slot slot32(s) by b prod_j g_j^{s_j} (s_j = bits of s), with
// the six factors read from a per-trip table and the products formed in registers:
//     trip 1: g_j = W_1024^{lane 2^j},          b = 1 (p = 0) or W_2048^lane (p = 1)
//     trip 3: g_j = W_2048^{(2 lane + p) 2^j},  b = 1 (p = 0) or (-1)^lane   (p = 1: moves e[m] to slot32(m ^ 16))
//     trip 4: index m' = m ^ 16 in slot32(m): g = W_64^{1, 2, 4, 8, -16}, b = W_64^{16}
// Table ids: p = 0: trip 1 -> 0, trip 3 -> 1;   p = 1: trips 1, 3, 4 -> 2, 3, 4.
// Spectrum table: hs[(p * 32 + k1) * 32 + k2] = H[2 (k2 + 32 k1) + p] / 2048.
constexpr int kTwPerTable = 6 * 32;
constexpr int kTwTables = 5;
constexpr int kTwElems = kTwTables * kTwPerTable;
FC_HD int w2_freq(int p, int k1, int k2) { return 2 * (k2 + 32 * k1) + p; }
FC_HD int trip_table(int i, int p) { return p ? (i == 1 ? 2 : i == 3 ? 3 : 4) : (i == 1 ? 0 : 1); }

// Applies use(s, w(s)) for s = 0..31, w(s) = b prod_j g_j^{s_j}, depth-first so that few partial
// products are live at a time.
template <class Use> FC_HD void for_each_twiddle(const cpair *tw, int lane, Use &&use) {
  const cpair g0 = tw[0 * 32 + lane], g1 = tw[1 * 32 + lane], g2 = tw[2 * 32 + lane], g3 = tw[3 * 32 + lane],
              g4 = tw[4 * 32 + lane], b = tw[5 * 32 + lane];
  use(0, b);
  const cpair w16 = cpair_mul(b, g4);
  use(16, w16);
  {
    const cpair w24 = cpair_mul(w16, g3);
    use(24, w24);
    {
      const cpair w28 = cpair_mul(w24, g2);
      use(28, w28);
      const cpair w30 = cpair_mul(w28, g1);
      use(30, w30);
      use(31, cpair_mul(w30, g0));
      use(29, cpair_mul(w28, g0));
    }
    const cpair w26 = cpair_mul(w24, g1);
    use(26, w26);
    use(27, cpair_mul(w26, g0));
    use(25, cpair_mul(w24, g0));
  }
  {
    const cpair w20 = cpair_mul(w16, g2);
    use(20, w20);
    const cpair w22 = cpair_mul(w20, g1);
    use(22, w22);
    use(23, cpair_mul(w22, g0));
    use(21, cpair_mul(w20, g0));
  }
  {
    const cpair w18 = cpair_mul(w16, g1);
    use(18, w18);
    use(19, cpair_mul(w18, g0));
    use(17, cpair_mul(w16, g0));
  }
  const cpair w8 = cpair_mul(b, g3);
  use(8, w8);
  {
    const cpair w12 = cpair_mul(w8, g2);
    use(12, w12);
    const cpair w14 = cpair_mul(w12, g1);
    use(14, w14);
    use(15, cpair_mul(w14, g0));
    use(13, cpair_mul(w12, g0));
    const cpair w10 = cpair_mul(w8, g1);
    use(10, w10);
    use(11, cpair_mul(w10, g0));
    use(9, cpair_mul(w8, g0));
  }
  const cpair w4 = cpair_mul(b, g2);
  use(4, w4);
  {
    const cpair w6 = cpair_mul(w4, g1);
    use(6, w6);
    use(7, cpair_mul(w6, g0));
    use(5, cpair_mul(w4, g0));
  }
  const cpair w2 = cpair_mul(b, g1);
  use(2, w2);
  use(3, cpair_mul(w2, g0));
  use(1, cpair_mul(b, g0));
}

// ---- building blocks of a trip ---------------------------------------------------------------
// p = 1, before the loop: times W_64^m on the point of index m (natural order)
FC_HD void twiddle64(Regs &v) {
#pragma unroll
  for (int m = 1; m < kPts; ++m) {
    if (m == 16) rot90<-1>(v.re[m], v.im[m]);
    else cmul(v.re[m], v.im[m], cos64(m), -sin64(m));
  }
}
FC_HD void chain(int lane, Regs &v, const cpair *table) {
  w2::for_each_twiddle(table, lane, [&](int s, cpair w) {
    const int sl = slot32(s);
    cmul(v.re[sl], v.im[sl], w.re, w.im);
  });
}
// the exchange buffer of the warp is written by columns and read by rows (row stride 33 points:
// conflict-free both ways)
FC_HD void put_columns(int lane, const Regs &v, pt4 *xb) {
#pragma unroll
  for (int s = 0; s < kPts; ++s) put(xb, s * kRow + lane, v.re[slot32(s)], v.im[slot32(s)]);
}
FC_HD void get_rows(int lane, Regs &v, const pt4 *xb) {
#pragma unroll
  for (int s = 0; s < kPts; ++s) get(xb, lane * kRow + s, v.re[s], v.im[s]);
}
// times H, back to natural slot order, and into the swapped domain (re <-> im)
FC_HD void spectrum_multiply(int lane, Regs &v, const cpair *hs_p /* hs + p * 1024 */) {
  Regs y;
#pragma unroll
  for (int k1 = 0; k1 < kPts; ++k1) {
    const cpair h = hs_p[k1 * 32 + lane];
    f2 re = v.re[slot32(k1)], im = v.im[slot32(k1)];
    cmul(re, im, h.re, h.im);
    y.re[k1] = im;
    y.im[k1] = re;
  }
  v = y;
}
// One trip up to its exchange (returns true when the caller must run get_rows after a warp sync).
FC_HD bool trip_head(int i, int p, int lane, Regs &v, pt4 *xb, const cpair *tw, const cpair *hs_p) {
  dft32<-1>(v);
  if (i == 2) {
    spectrum_multiply(lane, v, hs_p);
    return false;
  }
  if (i == 4 && !p) return false;
  chain(lane, v, tw + trip_table(i, p) * kTwPerTable);
  if (i == 4) return false;
  put_columns(lane, v, xb);
  return true;
}

// After trip 4 lane t holds (swapped) e[t + 32 m] in slot slot32(m) for p = 0 and in slot
// slot32(m ^ 16) for p = 1: every warp keeps the
// slots of m < 16 and hands the slots of m >= 16 to its partner, through row `lane` of its own
// buffer (addresses this lane alone has just read).
FC_HD void combine_put(int lane, const Regs &v, pt4 *xb_own) {
#pragma unroll
  for (int j = 0; j < 16; ++j) put(xb_own, lane * kRow + j, v.re[slot32(j + 16)], v.im[slot32(j + 16)]);
}
// afterwards (un-swapped) lo[j] = y[t + 32 (j + 16 p)], hi[j] = y[t + 32 (j + 16 p) + 1024];
// sigma = +1 (p = 0: mine = E, partner = O') or -1 (p = 1: mine = O', partner = E)
struct Out {
  f2 lo_re[16], lo_im[16], hi_re[16], hi_im[16];
};
FC_HD void combine_get(int lane, const Regs &v, const pt4 *xb_partner, float sigma, Out &o) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    f2 pr, pi;
    get(xb_partner, lane * kRow + j, pr, pi);
    const f2 mr = v.re[slot32(j)], mi = v.im[slot32(j)];
    o.lo_im[j] = add2(mr, pr); // swapped domain: its real part is the imaginary part of y
    o.lo_re[j] = add2(mi, pi);
    o.hi_im[j] = muls2(sub2(mr, pr), sigma);
    o.hi_re[j] = muls2(sub2(mi, pi), sigma);
  }
}

} // namespace w2
} // namespace fc
