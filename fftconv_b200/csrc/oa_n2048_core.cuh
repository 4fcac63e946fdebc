// oa_n2048_core.cuh -- the fused overlap-add segment engine for FFT size 2048
// (the reference's segment size for 221 <= K < 411 taps, e.g. K = 245:
// /root/reference/include/fftconv/fftconv.hpp:38-66), float32.
//
// One "group" of 128 threads transforms FOUR real segments at once:
//   * two segments ride in the real and imaginary parts of one complex FFT --
//     the taps are real, so IFFT(FFT(a + i b) * H) = (a conv h) + i (b conv h)
//     and no r2c/c2r untangling pass exists at all;
//   * two such complex FFTs ride in the two halves of every packed f32x2
//     register (packed_f32x2.cuh), so each butterfly is ONE instruction.
// This replaces, per segment, the reference's copy_to_padded_buffer + r2c +
// multiply_cx + c2r + normalize_add (fftconv.hpp:400-410 / :417-432).
//
// 2048 = 16 x 16 x 8, 16 points per thread, decimation in frequency:
//   n = t + 128 m            (thread t loads x[t + 128 m], m = 0..15, coalesced)
//   t = 8 t1 + t0
//   k = 16 k1a + 256 k1b + k2
//   W_N^{nk} = W_16^{m k2} W_2048^{t k2}  W_16^{t1 k1a} W_128^{t0 k1a}  W_8^{t0 k1b}
//   P0 (thread t):        radix-16 over m  -> k2,  twiddle W_2048^{t k2}
//   P1 (thread t0 + 8k2): radix-16 over t1 -> k1a, twiddle W_128^{t0 k1a}
//   P2 (thread t: columns c = p2_col(t, 0), p2_col(t, 1); c = 16 k2 + k1a): radix-8 over
//                         t0 -> k1b, times H[k], inverse radix-8 back to t0
//   P3 / P4 mirror P1 / P0 with conjugated twiddles; 1/N is folded into H.
// The three exchanges go through ONE 32 KB shared buffer of 16-byte points
// (reA, reB, imA, imB).  Its layouts are chosen so that every phase writes
// exactly the addresses it has just read, so only four barriers per segment
// quad are needed, and every 128-bit access is bank-conflict free:
//   X1: point (t, k2)        at k2*128 + 8 t1 + ((t0 + t1) & 7)
//   X2: point (c, t0)        at c*8 + ((t0 + c) & 7)
//
// Everything here is __host__ __device__ so tests can emulate the 128 threads
// phase by phase on the CPU (tests/cpu/emulate_oa_n2048.cpp).
//
// The engine is a template over its data type D (packed_f32x2.cuh, DataTraits):
//   D = f2      float32, FOUR real segments per group (as described above);
//   D = double  float64, TWO real segments per group (real / imaginary part), the
//               same 16-byte points, layouts, phases and barrier count.
#pragma once

#include "packed_f32x2.cuh"

namespace fc {
namespace n2048 {

constexpr int kN = 2048;
constexpr int kThreads = 128; // per group
constexpr int kPts = 16;      // points per thread

template <typename D> struct PtT { // one complex point (16 B): f2 -> (reA, reB, imA, imB)
  D re, im;
};
using pt4 = PtT<f2>;

template <typename D> struct RegsT {
  D re[kPts];
  D im[kPts];
};
using Regs = RegsT<f2>;

constexpr float kC8 = 0.70710678118654752440f;  // cos(pi/4)
constexpr float kC16 = 0.92387953251128675613f; // cos(pi/8)
constexpr float kS16 = 0.38268343236508977173f; // sin(pi/8)
template <typename S> struct Consts;
template <> struct Consts<float> {
  static constexpr float c8 = kC8, c16 = kC16, s16 = kS16;
};
template <> struct Consts<double> {
  static constexpr double c8 = 0.70710678118654752440084436210485, c16 = 0.92387953251128675612818318939679,
                          s16 = 0.38268343236508977172845998403040;
};
template <typename D> using scalar_of_t = typename DataTraits<D>::S;

// 4-point DFT, sign S (-1 forward, +1 inverse), in place on 4 named slots.
template <int S, typename D>
FC_HD void bfly4(D &r0, D &i0, D &r1, D &i1, D &r2, D &i2, D &r3, D &i3) {
  const D s0r = add2(r0, r2), s0i = add2(i0, i2);
  const D d0r = sub2(r0, r2), d0i = sub2(i0, i2);
  const D s1r = add2(r1, r3), s1i = add2(i1, i3);
  const D d1r = sub2(r1, r3), d1i = sub2(i1, i3);
  r0 = add2(s0r, s1r); i0 = add2(s0i, s1i);
  r2 = sub2(s0r, s1r); i2 = sub2(s0i, s1i);
  if (S < 0) { // X1 = d0 - j d1, X3 = d0 + j d1
    r1 = add2(d0r, d1i); i1 = sub2(d0i, d1r);
    r3 = sub2(d0r, d1i); i3 = add2(d0i, d1r);
  } else {     // X1 = d0 + j d1, X3 = d0 - j d1
    r1 = sub2(d0r, d1i); i1 = add2(d0i, d1r);
    r3 = add2(d0r, d1i); i3 = sub2(d0i, d1r);
  }
}

// multiply by -j (S<0) or +j (S>0)
template <int S, typename D> FC_HD void rot90(D &r, D &i) {
  using Sc = scalar_of_t<D>;
  const D t = r;
  if (S < 0) { r = i; i = muls2(t, (Sc)-1); }
  else       { r = muls2(i, (Sc)-1); i = t; }
}

// 16-point DFT of v.re/im[0..15], natural order in, natural order out.
//   m = m0 + 4 m1, k = k1 + 4 k0:
//   X[k1 + 4 k0] = sum_m0 W4^{m0 k0} W16^{m0 k1} sum_m1 x[m0 + 4 m1] W4^{m1 k1}
template <int S, typename D> FC_HD void dft16(RegsT<D> &v) {
  using Sc = scalar_of_t<D>;
  constexpr Sc kC8 = Consts<Sc>::c8, kC16 = Consts<Sc>::c16, kS16 = Consts<Sc>::s16;
  // step A: 4-point DFTs over m1 for each m0; result A[m0][k1] stays in slot m0 + 4 k1
#pragma unroll
  for (int m0 = 0; m0 < 4; ++m0)
    bfly4<S>(v.re[m0], v.im[m0], v.re[m0 + 4], v.im[m0 + 4], v.re[m0 + 8], v.im[m0 + 8],
             v.re[m0 + 12], v.im[m0 + 12]);
  // step B: twiddles W16^{m0 k1} on slot m0 + 4 k1   (imag sign follows S)
  const Sc s = (S < 0) ? (Sc)-1 : (Sc)1;
  cmul2(v.re[1 + 4], v.im[1 + 4], kC16, s * kS16);    // e = 1
  cmul2(v.re[2 + 4], v.im[2 + 4], kC8, s * kC8);      // e = 2
  cmul2(v.re[3 + 4], v.im[3 + 4], kS16, s * kC16);    // e = 3
  cmul2(v.re[1 + 8], v.im[1 + 8], kC8, s * kC8);      // e = 2
  rot90<S>(v.re[2 + 8], v.im[2 + 8]);                 // e = 4
  cmul2(v.re[3 + 8], v.im[3 + 8], -kC8, s * kC8);     // e = 6
  cmul2(v.re[1 + 12], v.im[1 + 12], kS16, s * kC16);  // e = 3
  cmul2(v.re[2 + 12], v.im[2 + 12], -kC8, s * kC8);   // e = 6
  cmul2(v.re[3 + 12], v.im[3 + 12], -kC16, -s * kS16); // e = 9
  // step C: 4-point DFTs over m0 for each k1; X[k1 + 4 k0] lands in slot 4 k1 + k0
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
    bfly4<S>(v.re[4 * k1], v.im[4 * k1], v.re[4 * k1 + 1], v.im[4 * k1 + 1], v.re[4 * k1 + 2],
             v.im[4 * k1 + 2], v.re[4 * k1 + 3], v.im[4 * k1 + 3]);
}
// slot that holds output k of dft16 (k = k1 + 4 k0 sits in slot 4 k1 + k0)
FC_HD constexpr int slot16(int k) { return 4 * (k & 3) + (k >> 2); }

// 8-point DFT on slots base..base+7, natural in, output k in slot base + slot8(k).
//   m = m0 + 2 m1, k = k1 + 4 k0:
//   X[k1 + 4 k0] = sum_m0 W2^{m0 k0} W8^{m0 k1} sum_m1 x[m0 + 2 m1] W4^{m1 k1}
template <int S, typename D> FC_HD void dft8(D *re, D *im) {
  using Sc = scalar_of_t<D>;
  constexpr Sc kC8 = Consts<Sc>::c8;
  bfly4<S>(re[0], im[0], re[2], im[2], re[4], im[4], re[6], im[6]); // m0 = 0: A[0][k1] in slot 2 k1
  bfly4<S>(re[1], im[1], re[3], im[3], re[5], im[5], re[7], im[7]); // m0 = 1: A[1][k1] in slot 2 k1 + 1
  const Sc s = (S < 0) ? (Sc)-1 : (Sc)1;
  cmul2(re[3], im[3], kC8, s * kC8);    // W8^1
  rot90<S>(re[5], im[5]);               // W8^2
  cmul2(re[7], im[7], -kC8, s * kC8);   // W8^3
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {      // X[k1] -> slot 2 k1, X[k1 + 4] -> slot 2 k1 + 1
    const D ar = re[2 * k1], ai = im[2 * k1], br = re[2 * k1 + 1], bi = im[2 * k1 + 1];
    re[2 * k1] = add2(ar, br); im[2 * k1] = add2(ai, bi);
    re[2 * k1 + 1] = sub2(ar, br); im[2 * k1 + 1] = sub2(ai, bi);
  }
}
FC_HD constexpr int slot8(int k) { return 2 * (k & 3) + (k >> 2); }

// shared-memory layouts ------------------------------------------------------
FC_HD int x1_addr(int t, int k2) { // t = 8 t1 + t0
  const int t1 = t >> 3, t0 = t & 7;
  return k2 * 128 + 8 * t1 + ((t0 + t1) & 7);
}
FC_HD int x2_addr(int c, int t0) { return c * 8 + ((t0 + c) & 7); }

template <typename D> FC_HD void put(PtT<D> *buf, int addr, D re, D im) {
  PtT<D> p;
  p.re = re;
  p.im = im;
  buf[addr] = p; // one 128-bit store
}
template <typename D> FC_HD void get(const PtT<D> *buf, int addr, D &re, D &im) {
  const PtT<D> p = buf[addr];
  re = p.re;
  im = p.im;
}

// Twiddle / spectrum tables (float2-like pairs), shared by all groups of a CTA:
//   tw1[k2 * 128 + t]   = W_2048^{t k2}       k2 = 0..15
//   tw2[k1a * 8 + t0]   = W_128^{t0 k1a}      k1a = 0..15
//   hs[slot * 128 + t]  = H[k] / 2048, slot = 8 h + k1b, c = p2_col(t, h),
//                         k = 16 (c & 15) + 256 k1b + (c >> 4)
template <typename S> struct CPairT { S re, im; };
using cpair = CPairT<float>;

// Phase P2: thread t handles two columns c (h = 0, 1) of 8 points each.  Warp w = t >> 5 takes
// c in [64 w, 64 w + 64): exactly the points [512 w, 512 w + 512) that the same warp wrote in P1
// (c = 16 k2 + k1a, k2 = 4 w .. 4 w + 3) and reads back in P3, so the exchanges P1 -> P2 -> P3
// are warp-local and need __syncwarp() only; the group barrier is kept where data crosses warps
// (P0 -> P1, P3 -> P4, and around the aliased staging area).
FC_HD int p2_col(int t, int h) { return 64 * (t >> 5) + 32 * h + (t & 31); }
// frequency handled by thread t in register slot s of phase P2
FC_HD int p2_freq(int t, int s) {
  const int h = s >> 3, k1b = s & 7, c = p2_col(t, h);
  return 16 * (c & 15) + 256 * k1b + (c >> 4);
}

// Twiddles w[k], k = 1..15, of one thread: either all 15 read from the table, or
// (RECOMP) only k = 1, 2, 4, 8 read and the rest formed as products in registers --
// the shared-memory crossbar is this kernel's bottleneck and the FP32 pipe has
// slack, so 11 fewer 8-byte loads are worth 44 scalar multiply-adds.  Products are
// at most three deep (k = 15 = 8 * (4 * (1 * 2))): relative twiddle error <= ~2e-7.
template <typename S> FC_HD CPairT<S> cpair_mul(CPairT<S> a, CPairT<S> b) {
  CPairT<S> r;
  r.re = a.re * b.re - a.im * b.im;
  r.im = a.re * b.im + a.im * b.re;
  return r;
}
// COMPACT (with RECOMP): the table holds only the four rows that are read, row r = entry k = 2^r.
template <bool RECOMP, bool COMPACT = false, typename S, class Use>
FC_HD void for_each_twiddle(const CPairT<S> *col /* entry k = 0 */, int stride, Use &&use) {
  using cpair = CPairT<S>;
  static_assert(RECOMP || !COMPACT, "the compact table only serves twiddle recomputation");
  if (!RECOMP) {
#pragma unroll
    for (int k = 1; k < 16; ++k) use(k, col[k * stride]);
  } else {
    const cpair w1 = col[(COMPACT ? 0 : 1) * stride], w2 = col[(COMPACT ? 1 : 2) * stride],
                w4 = col[(COMPACT ? 2 : 4) * stride], w8 = col[(COMPACT ? 3 : 8) * stride];
    use(1, w1); use(2, w2); use(4, w4); use(8, w8);
    const cpair w3 = cpair_mul(w1, w2); use(3, w3);
    const cpair w5 = cpair_mul(w4, w1); use(5, w5);
    const cpair w6 = cpair_mul(w4, w2); use(6, w6);
    const cpair w7 = cpair_mul(w4, w3); use(7, w7);
    use(9, cpair_mul(w8, w1));
    use(10, cpair_mul(w8, w2));
    use(11, cpair_mul(w8, w3));
    use(12, cpair_mul(w8, w4));
    use(13, cpair_mul(w8, w5));
    use(14, cpair_mul(w8, w6));
    use(15, cpair_mul(w8, w7));
  }
}

// ---- phases (called between group barriers) ---------------------------------
// P0: v holds x[t + 128 m] in slot m.  Split in two so that a barrier can sit
// between the last read of the (aliased) input staging area and the first write
// of the exchange buffer.
template <bool RECOMP = false, bool COMPACT = false, typename D>
FC_HD void phase0_compute(int t, RegsT<D> &v, const CPairT<scalar_of_t<D>> *tw1) {
  using cpair = CPairT<scalar_of_t<D>>;
  dft16<-1>(v);
  for_each_twiddle<RECOMP, COMPACT>(tw1 + t, 128, [&](int k2, cpair w) {
    const int s = slot16(k2);
    cmul2(v.re[s], v.im[s], w.re, w.im);
  });
}
template <typename D> FC_HD void phase0_put(int t, const RegsT<D> &v, PtT<D> *xb) {
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) put(xb, x1_addr(t, k2), v.re[slot16(k2)], v.im[slot16(k2)]);
}
template <typename D> FC_HD void phase0(int t, RegsT<D> &v, PtT<D> *xb, const CPairT<scalar_of_t<D>> *tw1) {
  phase0_compute<false, false>(t, v, tw1);
  phase0_put(t, v, xb);
}

template <bool RECOMP = false, typename D>
FC_HD void phase1(int tau, RegsT<D> &v, PtT<D> *xb, const CPairT<scalar_of_t<D>> *tw2) {
  using cpair = CPairT<scalar_of_t<D>>;
  const int t0 = tau & 7, k2 = tau >> 3;
#pragma unroll
  for (int t1 = 0; t1 < 16; ++t1) get(xb, x1_addr(8 * t1 + t0, k2), v.re[t1], v.im[t1]);
  dft16<-1>(v);
  for_each_twiddle<RECOMP>(tw2 + t0, 8, [&](int k1a, cpair w) {
    const int s = slot16(k1a);
    cmul2(v.re[s], v.im[s], w.re, w.im);
  });
#pragma unroll
  for (int k1a = 0; k1a < 16; ++k1a) {
    // same address set as the reads above: k2*128 + 8 j + ((t0 + j) & 7)
    put(xb, x2_addr(k2 * 16 + k1a, t0), v.re[slot16(k1a)], v.im[slot16(k1a)]);
  }
}

template <typename D> FC_HD void phase2(int c0, RegsT<D> &v, PtT<D> *xb, const CPairT<scalar_of_t<D>> *hs) {
  using cpair = CPairT<scalar_of_t<D>>;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = p2_col(c0, h);
#pragma unroll
    for (int t0 = 0; t0 < 8; ++t0) get(xb, x2_addr(c, t0), v.re[8 * h + t0], v.im[8 * h + t0]);
    dft8<-1>(&v.re[8 * h], &v.im[8 * h]);
    // spectrum multiply; dft8 output k1b sits in slot slot8(k1b).  The inverse
    // dft8 wants natural order in, so gather while multiplying.
    D yr[8], yi[8];
#pragma unroll
    for (int k1b = 0; k1b < 8; ++k1b) {
      const cpair w = hs[(8 * h + k1b) * 128 + c0];
      yr[k1b] = v.re[8 * h + slot8(k1b)];
      yi[k1b] = v.im[8 * h + slot8(k1b)];
      cmul2(yr[k1b], yi[k1b], w.re, w.im);
    }
    dft8<+1>(yr, yi);
#pragma unroll
    for (int t0 = 0; t0 < 8; ++t0) put(xb, x2_addr(c, t0), yr[slot8(t0)], yi[slot8(t0)]);
  }
}

template <bool RECOMP = false, typename D>
FC_HD void phase3(int tau, RegsT<D> &v, PtT<D> *xb, const CPairT<scalar_of_t<D>> *tw2) {
  using cpair = CPairT<scalar_of_t<D>>;
  const int t0 = tau & 7, k2 = tau >> 3;
#pragma unroll
  for (int k1a = 0; k1a < 16; ++k1a) get(xb, x2_addr(k2 * 16 + k1a, t0), v.re[k1a], v.im[k1a]);
  for_each_twiddle<RECOMP>(tw2 + t0, 8, [&](int k1a, cpair w) { cmul2(v.re[k1a], v.im[k1a], w.re, -w.im); });
  dft16<+1>(v);
#pragma unroll
  for (int t1 = 0; t1 < 16; ++t1)
    put(xb, x1_addr(8 * t1 + t0, k2), v.re[slot16(t1)], v.im[slot16(t1)]);
}

// P4: leaves y[t + 128 m] in slot slot16(m).  Split so that a barrier (and the
// request for the next inputs, which land in the exchange buffer) can follow the
// last read of the buffer.
template <bool RECOMP = false, bool COMPACT = false, typename D>
FC_HD void phase4_read(int t, RegsT<D> &v, const PtT<D> *xb, const CPairT<scalar_of_t<D>> *tw1) {
  using cpair = CPairT<scalar_of_t<D>>;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) get(xb, x1_addr(t, k2), v.re[k2], v.im[k2]);
  for_each_twiddle<RECOMP, COMPACT>(tw1 + t, 128, [&](int k2, cpair w) { cmul2(v.re[k2], v.im[k2], w.re, -w.im); });
}
template <typename D> FC_HD void phase4_compute(RegsT<D> &v) { dft16<+1>(v); }
template <typename D> FC_HD void phase4(int t, RegsT<D> &v, const PtT<D> *xb, const CPairT<scalar_of_t<D>> *tw1) {
  phase4_read<false, false>(t, v, xb, tw1);
  phase4_compute(v);
}

} // namespace n2048
} // namespace fc
