// hilbert_r2c_core.cuh -- Hilbert envelope of 8192-sample float32 lines through ONE 4096-point
// complex FFT per line (real-input formulation), 128 threads per line, four independent lines per SM.
//
// Replaces, per line, the reference's r2c FFT + (-j sgn f) multiplier + c2r FFT + magnitude loop
// (/root/reference/include/fftconv/hilbert.hpp:16-66).
//
// Why a second Hilbert engine (BASELINE config 4: 8192 lines x 8192 samples): the register-blocked
// kernel (fast_kernels.cuh) runs one 8192-point packed FFT of four lines on a whole SM -- eight
// passes of the data through shared memory (ncu: LSU pipe 59 %, FP32 pipe 42 %, issue 38 %, all
// sixteen warps in the same phase, shared-memory and arithmetic phases back to back).  Here
//   * a line x[0 .. N), N = 8192, is the complex sequence z[m] = x[2m] + j x[2m+1], m < M = 4096;
//     its FFT Z gives the spectrum of the Hilbert transform h = H{x}, packed the same way
//     (g[m] = h[2m] + j h[2m+1]), with one multiply-add pass over mirrored bins:
//         G[k] = cos(t_k) conj(Z[M - k]) + j sin(t_k) Z[k],  t_k = 2 pi k / N,  G[0] = 0
//     (derivation in DESIGN.md section 3.11), so a line costs two 4096-point complex FFTs;
//   * 4096 = 16 x 16 x 16: three radix-16 passes per direction, FOUR exchanges through shared
//     memory per line in all -- the first pass reads global memory, the last one leaves the
//     envelope in registers;
//   * a thread holds 32 points: TWO radix-16 butterflies, one in each half of every packed
//     f32x2 register (SASS FADD2 / FFMA2), 128 threads per line, 4 lines per SM in flight,
//     each with its own named barrier (4 barriers per line);
//   * in the third pass the two butterflies of a thread are chosen as MIRROR PARTNERS: the
//     thread that holds bins k0 + 16 k1 + 256 k2 (k2 = 0..15) in its low halves holds bins
//     M - (...) in its high halves, so the mirror pass happens between the halves of the
//     thread's own registers -- no exchange, no barrier.
//
// Index algebra.  m = 256 a + 16 b + c, k = k0 + 16 k1 + 256 k2, W = exp(-2 pi j / 4096):
//   W^{mk} = W16^{a k0} . W^{(16 b + c) k0} . W16^{b k1} . W256^{c k1} . W16^{c k2}
//   P1 (thread t: b = t >> 3, c = 2 (t & 7) + h):  DFT16 over a -> k0, times W^{(16 b + c) k0}
//   P2 (thread u: k0 = p2_k0(u), c = 2 (u & 7) + h): DFT16 over b -> k1, times W256^{c k1}
//   P3 (thread s: (k0, k1) and its mirror partner): DFT16 over c -> k2
// h = 0 / 1 is the low / high half of the packed registers.  The inverse runs the same passes
// backwards (P3', P2', P1') with conjugated twiddles applied BEFORE each butterfly; 1 / M rides in
// the mirror pass's cos / sin table.  Point (d0, d1, d2) -- digits (a|k0, b|k1, c|k2) -- lives at
// word  alpha(d0) + beta(d1) + d2  of two planes (real parts, imaginary parts); every pass rewrites
// exactly the words it read, so one barrier per exchange is enough.
//
// Everything here is __host__ __device__: tests/cpu/emulate_hilbert_r2c.cpp runs the same
// functions thread by thread on the CPU against a float64 Hilbert transform.
#pragma once

#include "oa_n2048_core.cuh"

namespace fc {
namespace hr2c {

using n2048::cpair;
using n2048::dft16;
using n2048::slot16;
using Regs = n2048::RegsT<f2>;

constexpr int kRow = 8192;      // samples per line
constexpr int kM = kRow / 2;    // complex points
constexpr int kThreads = 128;   // per line
constexpr int kGroups = 4;      // lines in flight per CTA

// plane layout: block of digit 0 every 272 words; rows of 16 words, two spare words after every
// second row (so that 32-bit column reads of 16 rows spread over 16 banks and 64-bit reads of two
// neighbouring blocks over all 32)
constexpr int kBlockStride = 272;
constexpr int kPlane = 16 * kBlockStride; // words per plane
FC_HD constexpr int alpha(int d0) { return kBlockStride * d0; }
FC_HD constexpr int beta(int d1) { return 16 * d1 + 2 * (d1 >> 1); }
constexpr int kGroupWords = 2 * kPlane;   // real plane, imaginary plane

struct alignas(8) Pair {
  float x, y;
};
struct alignas(16) Quad {
  float a, b, c, d;
};

// ---- which butterflies a thread runs ---------------------------------------------------------------
// P1 / P1': thread t, digits (b, c) = (t >> 3, 2 (t & 7) + h)
FC_HD int p1_base(int t) { return beta(t >> 3) + 2 * (t & 7); }
// P2 / P2': thread u, digits (k0, c) = (p2_k0(u), 2 (u & 7) + h).  Warp w owns a set of k0 that is closed under
// the mirror map k0 -> 16 - k0: {2w+1, 2w+2, 15-2w, 14-2w} (w < 3), {7, 8, 9, 0} -- exactly the blocks its
// threads read in P3, where a thread holds a butterfly and its mirror partner.  The exchanges P2 -> P3 and
// P3' -> P2' therefore stay inside a warp and need __syncwarp() only; neighbours in a half-warp have k0 of
// opposite parity (blocks 16 words apart mod 32: conflict-free 64-bit accesses).
FC_HD int p2_k0(int u) {
  const int w = u >> 5, q = (u >> 3) & 3;
  if (w < 3) return q < 2 ? 2 * w + 1 + q : 16 - (2 * w + 1 + (q - 2));
  return q == 3 ? 0 : 7 + q;
}
FC_HD int p2_base(int u) { return alpha(p2_k0(u)) + 2 * (u & 7); }
// P3 / P3': thread s holds butterfly A = (k0, k1) in the low halves and its mirror partner B in the
// high halves: bins M - (k0 + 16 k1 + 256 k2) = k0' + 16 k1' + 256 (15 - k2).
//   s < 112:        A = (1 + (s >> 4), s & 15)   B = (16 - k0, 15 - k1)
//   112 <= s < 120: A = (8, s - 112)             B = (8, 15 - k1)
//   120 <= s < 127: A = (0, s - 119)             B = (0, 16 - k1)
//   s = 127:        A = (0, 0), B = (0, 8): both their own partners (k2 <-> 16 - k2 resp. 15 - k2)
struct P3Map {
  int k0a, k1a, k0b, k1b;
};
FC_HD P3Map p3_map(int s) {
  P3Map m;
  if (s < 112) {
    m.k0a = 1 + (s >> 4); m.k1a = s & 15; m.k0b = 16 - m.k0a; m.k1b = 15 - m.k1a;
  } else if (s < 120) {
    m.k0a = 8; m.k1a = s - 112; m.k0b = 8; m.k1b = 15 - m.k1a;
  } else if (s < 127) {
    m.k0a = 0; m.k1a = s - 119; m.k0b = 0; m.k1b = 16 - m.k1a;
  } else {
    m.k0a = 0; m.k1a = 0; m.k0b = 0; m.k1b = 8;
  }
  return m;
}
constexpr int kSelfPaired = 127; // the one thread whose two butterflies are their own mirror images

// ---- twiddles ----------------------------------------------------------------------------------------
// A packed twiddle: different values in the two halves (the two butterflies of a thread differ in c).
struct TwP {
  f2 re, im;
};
FC_HD TwP twp_load(const Quad *q) { // table entry (re lo, re hi, im lo, im hi)
  const Quad v = *q;
  TwP w;
  w.re = make_f2(v.a, v.b);
  w.im = make_f2(v.c, v.d);
  return w;
}
FC_HD TwP twp_mul(TwP a, TwP b) {
  TwP r;
  r.re = fnma2(a.im, b.im, mul2(a.re, b.re));
  r.im = fma2(a.re, b.im, mul2(a.im, b.re));
  return r;
}
// (re, im) *= w  or  *= conj(w)
template <bool CONJ> FC_HD void twp_apply(f2 &re, f2 &im, TwP w) {
  const f2 a = mul2(re, w.re), b = mul2(im, w.re);
  if (!CONJ) {
    const f2 nre = fnma2(im, w.im, a); // re wr - im wi
    im = fma2(re, w.im, b);            // re wi + im wr
    re = nre;
  } else {
    const f2 nre = fma2(im, w.im, a);  // re wr + im wi
    im = fnma2(re, w.im, b);           // im wr - re wi
    re = nre;
  }
}
// Slot j (natural) or slot16(j) of v times w^j, j = 1..15; the table holds w^1, w^2, w^4, w^8, the other
// powers are products formed in registers, at most three deep (as in oa_n2048_core.cuh).
template <bool CONJ, bool SLOT16> FC_HD void apply_powers(Regs &v, const Quad *tab, int stride) {
  const TwP w1 = twp_load(tab), w2 = twp_load(tab + stride), w4 = twp_load(tab + 2 * stride),
            w8 = twp_load(tab + 3 * stride);
  auto use = [&](int j, TwP w) {
    const int s = SLOT16 ? slot16(j) : j;
    twp_apply<CONJ>(v.re[s], v.im[s], w);
  };
  // order: every power is used -- and its product with w^8 formed and used -- as soon as it exists, so that
  // at most five twiddles are alive beside the data
  use(1, w1); use(9, twp_mul(w8, w1));
  use(2, w2); use(10, twp_mul(w8, w2));
  const TwP w3 = twp_mul(w1, w2); use(3, w3); use(11, twp_mul(w8, w3));
  const TwP w5 = twp_mul(w4, w1); use(5, w5); use(13, twp_mul(w8, w5));
  const TwP w6 = twp_mul(w4, w2); use(6, w6); use(14, twp_mul(w8, w6));
  const TwP w7 = twp_mul(w4, w3); use(7, w7); use(15, twp_mul(w8, w7));
  use(4, w4); use(12, twp_mul(w8, w4));
  use(8, w8);
}
// tables (hilbert_r2c_tables.h):
//   tw1[p * 128 + t] = W4096^{r 2^p},  r = 16 (t >> 3) + 2 (t & 7) + h      (thread t of P1 / P1')
//   tw2[p * 8 + j]   = W256^{c 2^p},   c = 2 j + h                            (thread u of P2 / P2', j = u & 7)
//   mir[k2 * 129 + s] = (cos, sin)(2 pi k / 8192) / 4096, k = the bin of thread s's low half, slot k2;
//                       column 128: the high half of the self-paired thread
// (thread-minor, so that a warp reads consecutive entries)
constexpr int kTw1Quads = kThreads * 4;
constexpr int kTw2Quads = 8 * 4;
constexpr int kMirStride = kThreads + 1;
constexpr int kMirPairs = kMirStride * 16;

// ---- global memory <-> registers ---------------------------------------------------------------------
// P1 input: thread t reads x[512 a + 4 t .. + 4) = (re, im) of points m = 256 a + 2 t and + 1, a = 0..15
FC_HD void load_line(int t, Regs &v, const float *x) {
  const Quad *q = reinterpret_cast<const Quad *>(x) + t;
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    const Quad p = q[128 * a];
    v.re[a] = make_f2(p.a, p.c);
    v.im[a] = make_f2(p.b, p.d);
  }
}

// ---- shared memory <-> registers ---------------------------------------------------------------------
// 64-bit accesses: both halves of a register at once (the two butterflies are neighbours in the last digit)
template <bool SLOT16> FC_HD void put64(const Regs &v, float *planes, int base, int j, int off) {
  const int s = SLOT16 ? slot16(j) : j;
  *reinterpret_cast<Pair *>(planes + base + off) = Pair{v.re[s].x, v.re[s].y};
  *reinterpret_cast<Pair *>(planes + kPlane + base + off) = Pair{v.im[s].x, v.im[s].y};
}
FC_HD void get64(Regs &v, const float *planes, int base, int j, int off) {
  const Pair r = *reinterpret_cast<const Pair *>(planes + base + off);
  const Pair i = *reinterpret_cast<const Pair *>(planes + kPlane + base + off);
  v.re[j] = make_f2(r.x, r.y);
  v.im[j] = make_f2(i.x, i.y);
}
// after P1 / before P1': digit 0 varies over the slots
FC_HD void p1_store(Regs &v, float *planes, int base) {
#pragma unroll
  for (int j = 0; j < 16; ++j) put64<true>(v, planes, base, j, alpha(j));
}
FC_HD void p1_fetch(Regs &v, const float *planes, int base) {
#pragma unroll
  for (int j = 0; j < 16; ++j) get64(v, planes, base, j, alpha(j));
}
// around P2 / P2': digit 1 varies
FC_HD void p2_store(Regs &v, float *planes, int base) {
#pragma unroll
  for (int j = 0; j < 16; ++j) put64<true>(v, planes, base, j, beta(j));
}
FC_HD void p2_fetch(Regs &v, const float *planes, int base) {
#pragma unroll
  for (int j = 0; j < 16; ++j) get64(v, planes, base, j, beta(j));
}
// around P3 / P3': digit 2 varies; the two halves belong to different butterflies (32-bit accesses)
FC_HD void p3_fetch(Regs &v, const float *planes, int base_a, int base_b) {
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    v.re[c] = make_f2(planes[base_a + c], planes[base_b + c]);
    v.im[c] = make_f2(planes[kPlane + base_a + c], planes[kPlane + base_b + c]);
  }
}
// 64-bit stores of two neighbouring points of ONE butterfly (base_a, base_b are even): the halves have to be
// re-paired in registers (MOVs, on the idle integer pipe), but 32-bit stores from a warp whose 32 butterflies
// all sit at even words would hit only half of the banks (two-way conflict: 128 wavefronts instead of 64)
FC_HD void p3_store(const Regs &v, float *planes, int base_a, int base_b) {
#pragma unroll
  for (int c = 0; c < 16; c += 2) {
    const int s0 = slot16(c), s1 = slot16(c + 1);
    *reinterpret_cast<Pair *>(planes + base_a + c) = Pair{v.re[s0].x, v.re[s1].x};
    *reinterpret_cast<Pair *>(planes + base_b + c) = Pair{v.re[s0].y, v.re[s1].y};
    *reinterpret_cast<Pair *>(planes + kPlane + base_a + c) = Pair{v.im[s0].x, v.im[s1].x};
    *reinterpret_cast<Pair *>(planes + kPlane + base_b + c) = Pair{v.im[s0].y, v.im[s1].y};
  }
}

// ---- the mirror pass -----------------------------------------------------------------------------------
// After P3 slot16(k2) holds Z[k] (low half: butterfly A, high half: B).  G[k] = C conj(Z[M - k]) + j S Z[k];
// for the partner bin k' = M - k: cos -> -C, sin -> S.  Results go to NATURAL slots (k2 for A, 15 - k2 for
// B's bin M - k), which is where the inverse dft16 wants its inputs.
FC_HD void mirror(Regs &v, const Pair *cs /* this thread's column of the (C, S) table */) {
  Regs w;
  // slots k2 and 15 - k2 together: the four points they hold are consumed and the four results produced in
  // one step, so nothing but the data itself is alive (taken one k2 at a time half of the inputs stay alive
  // for eight steps beside the results: spills)
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    const int j2 = 15 - k2, sa = slot16(k2), sb = slot16(j2);
    const float C = cs[k2 * kMirStride].x, S = cs[k2 * kMirStride].y;   // bin k of A, slot k2
    const float D = cs[j2 * kMirStride].x, T = cs[j2 * kMirStride].y;   // bin of A, slot 15 - k2
    const float axr = v.re[sa].x, axi = v.im[sa].x, ayr = v.re[sa].y, ayi = v.im[sa].y;
    const float bxr = v.re[sb].x, bxi = v.im[sb].x, byr = v.re[sb].y, byi = v.im[sb].y;
    // A's bin in slot k2 pairs with B's bin in slot 15 - k2 (by), A's bin in slot 15 - k2 with B's in slot k2 (ay)
    w.re[k2].x = C * byr - S * axi;
    w.im[k2].x = S * axr - C * byi;
    w.re[j2].y = -C * axr - S * byi;
    w.im[j2].y = C * axi + S * byr;
    w.re[j2].x = D * ayr - T * bxi;
    w.im[j2].x = T * bxr - D * ayi;
    w.re[k2].y = -D * bxr - T * ayi;
    w.im[k2].y = D * bxi + T * ayr;
  }
  v = w;
}
// the self-paired thread: A = bins 256 k2 (partner 256 (16 - k2); G[0] = 0: its table entry is (0, 0)),
// B = bins 128 + 256 k2 (partner 128 + 256 (15 - k2)); cs_b: the extra table row
FC_HD void mirror_self(Regs &v, const Pair *cs_a, const Pair *cs_b) {
  Regs w;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const int s = slot16(k2), pa = slot16((16 - k2) & 15), pb = slot16(15 - k2);
    const float Ca = cs_a[k2 * kMirStride].x, Sa = cs_a[k2 * kMirStride].y;
    const float Cb = cs_b[k2 * kMirStride].x, Sb = cs_b[k2 * kMirStride].y;
    w.re[k2].x = Ca * v.re[pa].x - Sa * v.im[s].x;
    w.im[k2].x = Sa * v.re[s].x - Ca * v.im[pa].x;
    w.re[k2].y = Cb * v.re[pb].y - Sb * v.im[s].y;
    w.im[k2].y = Sb * v.re[s].y - Cb * v.im[pb].y;
  }
  v = w;
}

// ---- the passes, cut at the dft16 calls ------------------------------------------------------------------
//   forward:  load_line | dft16 | fwd_a | dft16 | fwd_b | dft16 | fwd_c          (| = the kernel's loop trips)
//   inverse:            | dft16 | inv_a | dft16 | inv_b | dft16 | envelope
// `sync` is the line's barrier (the CPU emulation runs all threads up to each barrier instead).
struct ThreadMaps { // what a thread keeps across lines
  int base1, base2, base3a, base3b;
  const Quad *tw1, *tw2; // its columns of the tables
  const Pair *cs;        // its column of the mirror table
};
FC_HD ThreadMaps thread_maps(int t, const Quad *tw1, const Quad *tw2, const Pair *mir) {
  ThreadMaps m;
  m.base1 = p1_base(t);
  m.base2 = p2_base(t);
  const P3Map p = p3_map(t);
  m.base3a = alpha(p.k0a) + beta(p.k1a);
  m.base3b = alpha(p.k0b) + beta(p.k1b);
  m.tw1 = tw1 + t;
  m.tw2 = tw2 + (t & 7);
  m.cs = mir + t;
  return m;
}
FC_HD void fwd_a_pre(Regs &v, float *planes, const ThreadMaps &m) {  // after P1's dft16, up to the barrier
  apply_powers<false, true>(v, m.tw1, kThreads);
  p1_store(v, planes, m.base1);
}
FC_HD void fwd_a_post(Regs &v, const float *planes, const ThreadMaps &m) { p2_fetch(v, planes, m.base2); }
FC_HD void fwd_b_pre(Regs &v, float *planes, const ThreadMaps &m) {  // after P2's dft16
  apply_powers<false, true>(v, m.tw2, 8);
  p2_store(v, planes, m.base2);
}
FC_HD void fwd_b_post(Regs &v, const float *planes, const ThreadMaps &m) { p3_fetch(v, planes, m.base3a, m.base3b); }
FC_HD void fwd_c(int t, Regs &v, const ThreadMaps &m, const Pair *mir) { // after P3's dft16: no exchange
  if (t == kSelfPaired) mirror_self(v, m.cs, mir + kThreads);
  else mirror(v, m.cs);
}
FC_HD void inv_a_pre(Regs &v, float *planes, const ThreadMaps &m) {  // after P3''s dft16
  p3_store(v, planes, m.base3a, m.base3b);
}
FC_HD void inv_a_post(Regs &v, const float *planes, const ThreadMaps &m) {
  p2_fetch(v, planes, m.base2);
  apply_powers<true, false>(v, m.tw2, 8);
}
FC_HD void inv_b_pre(Regs &v, float *planes, const ThreadMaps &m) {  // after P2''s dft16
  p2_store(v, planes, m.base2);
}
FC_HD void inv_b_post(Regs &v, const float *planes, const ThreadMaps &m) {
  p1_fetch(v, planes, m.base1);
  apply_powers<true, false>(v, m.tw1, kThreads);
}

} // namespace hr2c
} // namespace fc
