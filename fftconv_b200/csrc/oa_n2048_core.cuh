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

// 4-point DFT of (y0, w1 y1, w2 y2, w3 y3) with the constant twiddles folded into the butterflies
// (Linzer-Feig): w_i = c_i (1 + j t_i) with t_i = tan(arg w_i), so that
//   p_i = y_i (1 + j t_i)                       2 FMA
//   s0, d0 = y0 +- c2 p2                        4 FMA
//   s1', d1' = p1 +- (c3 / c1) p3               4 FMA      (s1 = c1 s1', d1 = c1 d1')
//   X0, X2 = s0 +- c1 s1'   X1, X3 = d0 -+ j c1 d1'   8 FMA
// 22 fused multiply-adds instead of 3 complex multiplications + 16 additions = 28 instructions.
// The t_i are given for the FORWARD transform (S < 0); the inverse conjugates them.  |t|, |c3/c1|
// <= tan(3 pi / 8) = 2.414 for every twiddle set used here, so nothing is amplified.
template <int S, typename D, typename Sc>
FC_HD void bfly4_tw(D &r0, D &i0, D &r1, D &i1, D &r2, D &i2, D &r3, D &i3, Sc c1, Sc t1f, Sc c2, Sc t2f, Sc t3f,
                    Sc rho /* c3 / c1 */) {
  const Sc t1 = S < 0 ? t1f : -t1f, t2 = S < 0 ? t2f : -t2f, t3 = S < 0 ? t3f : -t3f;
  const D p2r = fmas2(i2, -t2, r2), p2i = fmas2(r2, t2, i2);
  const D s0r = fmas2(p2r, c2, r0), s0i = fmas2(p2i, c2, i0);
  const D d0r = fmas2(p2r, -c2, r0), d0i = fmas2(p2i, -c2, i0);
  const D p1r = fmas2(i1, -t1, r1), p1i = fmas2(r1, t1, i1);
  const D p3r = fmas2(i3, -t3, r3), p3i = fmas2(r3, t3, i3);
  const D s1r = fmas2(p3r, rho, p1r), s1i = fmas2(p3i, rho, p1i);
  const D d1r = fmas2(p3r, -rho, p1r), d1i = fmas2(p3i, -rho, p1i);
  r0 = fmas2(s1r, c1, s0r); i0 = fmas2(s1i, c1, s0i);
  r2 = fmas2(s1r, -c1, s0r); i2 = fmas2(s1i, -c1, s0i);
  if (S < 0) { // X1 = d0 - j c1 d1', X3 = d0 + j c1 d1'
    r1 = fmas2(d1i, c1, d0r); i1 = fmas2(d1r, -c1, d0i);
    r3 = fmas2(d1i, -c1, d0r); i3 = fmas2(d1r, c1, d0i);
  } else {
    r1 = fmas2(d1i, -c1, d0r); i1 = fmas2(d1r, c1, d0i);
    r3 = fmas2(d1i, c1, d0r); i3 = fmas2(d1r, -c1, d0i);
  }
}
// the same for the twiddle set (W8, W4, W8^3) = (c8 (1 - j), -j, c8 (-1 - j)) (forward): only
// additions until the last level -- 20 instructions
template <int S, typename D, typename Sc>
FC_HD void bfly4_w8(D &r0, D &i0, D &r1, D &i1, D &r2, D &i2, D &r3, D &i3, Sc c8) {
  // w2 y2 = -+ j y2
  D s0r, s0i, d0r, d0i, p1r, p1i, p3r, p3i;
  if (S < 0) {
    s0r = add2(r0, i2); s0i = sub2(i0, r2); d0r = sub2(r0, i2); d0i = add2(i0, r2);
    p1r = add2(r1, i1); p1i = sub2(i1, r1);   // y1 (1 - j)
    p3r = sub2(r3, i3); p3i = add2(i3, r3);   // y3 (1 + j);  w3 y3 = -c8 p3
  } else {
    s0r = sub2(r0, i2); s0i = add2(i0, r2); d0r = add2(r0, i2); d0i = sub2(i0, r2);
    p1r = sub2(r1, i1); p1i = add2(i1, r1);   // y1 (1 + j)
    p3r = add2(r3, i3); p3i = sub2(i3, r3);   // y3 (1 - j);  w3 y3 = -c8 p3
  }
  const D s1r = sub2(p1r, p3r), s1i = sub2(p1i, p3i); // s1 = c8 s1'
  const D d1r = add2(p1r, p3r), d1i = add2(p1i, p3i);
  r0 = fmas2(s1r, c8, s0r); i0 = fmas2(s1i, c8, s0i);
  r2 = fmas2(s1r, -c8, s0r); i2 = fmas2(s1i, -c8, s0i);
  if (S < 0) {
    r1 = fmas2(d1i, c8, d0r); i1 = fmas2(d1r, -c8, d0i);
    r3 = fmas2(d1i, -c8, d0r); i3 = fmas2(d1r, c8, d0i);
  } else {
    r1 = fmas2(d1i, -c8, d0r); i1 = fmas2(d1r, c8, d0i);
    r3 = fmas2(d1i, c8, d0r); i3 = fmas2(d1r, -c8, d0i);
  }
}

// 16-point DFT of v.re/im[0..15], natural order in, natural order out.
//   m = m0 + 4 m1, k = k1 + 4 k0:
//   X[k1 + 4 k0] = sum_m0 W4^{m0 k0} W16^{m0 k1} sum_m1 x[m0 + 4 m1] W4^{m1 k1}
// 144 instructions: the twiddles W16^{m0 k1} between the two levels never appear as complex
// multiplications, they ride in the fused multiply-adds of the second level (bfly4_tw / bfly4_w8).
template <int S, typename D> FC_HD void dft16(RegsT<D> &v) {
  using Sc = scalar_of_t<D>;
  constexpr Sc kC8 = Consts<Sc>::c8, kC16 = Consts<Sc>::c16, kS16 = Consts<Sc>::s16;
  constexpr Sc kT16 = kS16 / kC16;     // tan(pi / 8)
  constexpr Sc kT316 = kC16 / kS16;    // tan(3 pi / 8)
  // step A: 4-point DFTs over m1 for each m0; result A[m0][k1] stays in slot m0 + 4 k1
#pragma unroll
  for (int m0 = 0; m0 < 4; ++m0)
    bfly4<S>(v.re[m0], v.im[m0], v.re[m0 + 4], v.im[m0 + 4], v.re[m0 + 8], v.im[m0 + 8],
             v.re[m0 + 12], v.im[m0 + 12]);
  // steps B + C: 4-point DFTs over m0 for each k1 with the twiddles W16^{m0 k1} (forward: e^{-j pi m0 k1 / 8})
  // folded in; X[k1 + 4 k0] lands in slot 4 k1 + k0
  bfly4<S>(v.re[0], v.im[0], v.re[1], v.im[1], v.re[2], v.im[2], v.re[3], v.im[3]);
  // k1 = 1: w = (W16, W16^2, W16^3) = (c16 (1 - j t16), c8 (1 - j), s16 (1 - j t316))
  bfly4_tw<S>(v.re[4], v.im[4], v.re[5], v.im[5], v.re[6], v.im[6], v.re[7], v.im[7],
              kC16, -kT16, kC8, (Sc)-1, -kT316, kS16 / kC16);
  // k1 = 2: w = (W16^2, W16^4, W16^6) = (c8 (1 - j), -j, c8 (-1 - j))
  bfly4_w8<S>(v.re[8], v.im[8], v.re[9], v.im[9], v.re[10], v.im[10], v.re[11], v.im[11], kC8);
  // k1 = 3: w = (W16^3, W16^6, W16^9) = (s16 (1 - j t316), -c8 (1 + j), -c16 (1 - j t16))
  bfly4_tw<S>(v.re[12], v.im[12], v.re[13], v.im[13], v.re[14], v.im[14], v.re[15], v.im[15],
              kS16, -kT316, -kC8, (Sc)1, -kT16, -kC16 / kS16);
}
// slot that holds output k of dft16 (k = k1 + 4 k0 sits in slot 4 k1 + k0)
FC_HD constexpr int slot16(int k) { return 4 * (k & 3) + (k >> 2); }

// 8-point DFT on slots base..base+7, natural in, output k in slot base + slot8(k).
//   m = m0 + 2 m1, k = k1 + 4 k0:
//   X[k1 + 4 k0] = sum_m0 W2^{m0 k0} W8^{m0 k1} sum_m1 x[m0 + 2 m1] W4^{m1 k1}
// 52 instructions: the W8 twiddles ride in the fused multiply-adds of the last level.
template <int S, typename D> FC_HD void dft8(D *re, D *im) {
  using Sc = scalar_of_t<D>;
  constexpr Sc kC8 = Consts<Sc>::c8;
  bfly4<S>(re[0], im[0], re[2], im[2], re[4], im[4], re[6], im[6]); // m0 = 0: A[0][k1] in slot 2 k1
  bfly4<S>(re[1], im[1], re[3], im[3], re[5], im[5], re[7], im[7]); // m0 = 1: A[1][k1] in slot 2 k1 + 1
  // X[k1] = A0[k1] + W8^{k1} A1[k1] -> slot 2 k1,  X[k1 + 4] = A0[k1] - W8^{k1} A1[k1] -> slot 2 k1 + 1
  { // k1 = 0
    const D ar = re[0], ai = im[0], br = re[1], bi = im[1];
    re[0] = add2(ar, br); im[0] = add2(ai, bi); re[1] = sub2(ar, br); im[1] = sub2(ai, bi);
  }
  { // k1 = 1: W8 = c8 (1 -+ j): u = b (1 -+ j), X = a +- c8 u
    const D ar = re[2], ai = im[2], br = re[3], bi = im[3];
    const D ur = S < 0 ? add2(br, bi) : sub2(br, bi), ui = S < 0 ? sub2(bi, br) : add2(bi, br);
    re[2] = fmas2(ur, kC8, ar); im[2] = fmas2(ui, kC8, ai); re[3] = fmas2(ur, -kC8, ar); im[3] = fmas2(ui, -kC8, ai);
  }
  { // k1 = 2: W8^2 = -+ j
    const D ar = re[4], ai = im[4], br = re[5], bi = im[5];
    if (S < 0) { re[4] = add2(ar, bi); im[4] = sub2(ai, br); re[5] = sub2(ar, bi); im[5] = add2(ai, br); }
    else       { re[4] = sub2(ar, bi); im[4] = add2(ai, br); re[5] = add2(ar, bi); im[5] = sub2(ai, br); }
  }
  { // k1 = 3: W8^3 = -c8 (1 +- j): u = b (1 +- j), X = a -+ c8 u
    const D ar = re[6], ai = im[6], br = re[7], bi = im[7];
    const D ur = S < 0 ? sub2(br, bi) : add2(br, bi), ui = S < 0 ? add2(bi, br) : sub2(bi, br);
    re[6] = fmas2(ur, -kC8, ar); im[6] = fmas2(ui, -kC8, ai); re[7] = fmas2(ur, kC8, ar); im[7] = fmas2(ui, kC8, ai);
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
