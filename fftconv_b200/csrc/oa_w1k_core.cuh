// oa_w1k_core.cuh -- overlap-add segments of 1024 points, ONE WARP per complex FFT.
//
// Why a second engine beside oa_n2048_*: that kernel is bound by the shared-memory data pipe
// (ncu, round 1: LSU pipe 75 % busy, 512 of its 751 shared wavefronts per warp iteration are the
// four exchanges of a 2048 = 16 x 16 x 8 decomposition).  A thread that holds 32 points instead
// of 16 needs only two exchanges: 1024 = 32 x 32, and both are transposes inside one warp --
// no block barrier anywhere, sixteen fully independent warps per SM.  Round 1 tried 32 points
// per thread with FOUR real segments per point (oa_w2): 128 data registers, 8 warps per SM,
// latency bound.  Here a thread holds 32 points of ONE complex FFT (two real segments in the real
// and imaginary parts), 64 data registers, and uses the packed f32x2 arithmetic ACROSS the points:
//     register pair i  =  (point i, point i + 16)      ("half packing")
// A 32-point DFT is then one radix-2 step between the halves of each pair (scalar instructions)
// and one 16-point DFT on all pairs at once (packed instructions, the dft16 of oa_n2048_core.cuh).
//
// Segments: the reference's overlap-add (fftconv.hpp:387-410) pads a block of step = N - (K-1)
// samples with zeros and adds the K-1 tail into the next block.  The same linear convolution is
// obtained by overlapping on the INPUT side: a block reads N samples starting K-1 before its first
// output and its first K-1 (circularly wrapped) outputs are dropped.  Every block is then
// independent of every other -- no tail hand-over, no ordering between blocks, any warp can take
// any block -- and every output sample is still written exactly once (deterministic, no atomics).
//
// 1024 = 32 x 32, n = l + 32 m, k = 32 k1 + k2, lane l, W = exp(-2 pi i / N):
//   P0  lane l:   A_l[k2] = sum_m z[l + 32 m] W_32^{m k2};   B_l[k2] = A_l[k2] W_1024^{l k2}
//   T1  transpose: lane k2 receives B_l[k2], l = 0..31
//   P1  lane k2:  X[32 k1 + k2] = sum_l B_l[k2] W_32^{l k1};  Y = X H;
//                 C_k2[l] = sum_k1 Y[32 k1 + k2] W_32^{-l k1}
//   T2  transpose: lane l receives C_k2[l], k2 = 0..31
//   P2  lane l:   y[l + 32 m] = sum_k2 C_k2[l] conj(W_1024^{l k2}) W_32^{-m k2}      (1/N folded into H)
//
// The four 32-point DFTs and which points share a register pair (dft16 works on pair index 0..15):
//   P0  forward, decimation in frequency:  in (m, m+16)    -> radix-2 step, dft16 -> out (k2 = 2j, 2j+1)
//   P1  forward, decimation in time:       in (l = 2i, 2i+1) -> dft16, radix-2 step -> out (k1 = j, j+16)
//   P1' inverse, decimation in frequency:  in (k1 = j, j+16) -> radix-2 step, dft16 -> out (l = 2i, 2i+1)
//   P2  inverse, decimation in time:       in (k2 = 2j, 2j+1) -> dft16, radix-2 step -> out (m, m+16)
// so that the transposes move 16-byte quantities on one side (rows of consecutive l) and the pair
// twiddles W_1024^{l k2} are (w2^j, w2^j w), w2 = w^2: a broadcast scalar per pair plus one extra
// scalar multiply on the odd half.
//
// Everything is __host__ __device__: tests/cpu/emulate_oa_w1k.cpp runs the same functions lane by lane.
#pragma once

#include "oa_n2048_core.cuh"

#include <cmath>

namespace fc {
namespace w1k {

using n2048::CPairT;
using n2048::cpair;
using n2048::dft16;
using n2048::RegsT;
using n2048::slot16;
using Regs = RegsT<f2>;

constexpr int kN = 1024;
constexpr int kRowStride = 36;              // floats per exchange row (32 + 4: conflict-free 128-bit rows, 32-bit columns)
constexpr int kPlane = 32 * kRowStride;     // floats per plane (re / im); also one row's input staging area
constexpr int kWarpFloats = 2 * kPlane;     // exchange buffer of one warp
constexpr int kMaxTaps = 400;               // step >= 625: below that the 2048-point kernel is the better choice

// cos / sin of 2 pi i / 32, i = 0..15
struct W32Table {
  float c[16], s[16];
};
FC_HD constexpr W32Table w32_table() {
  return W32Table{{1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                   0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
                   0.0f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                   -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f},
                  {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                   0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
                   1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                   0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f}};
}

// (dr, di) *= exp(S 2 pi j i / 32), i a compile-time constant 0..15
template <int S, int I> FC_HD void rot32(float &dr, float &di) {
  constexpr W32Table T = w32_table();
  if (I == 0) return;
  if (I == 8) { // +- j
    const float t = dr;
    if (S < 0) { dr = di; di = -t; }
    else       { dr = -di; di = t; }
    return;
  }
  const float c = T.c[I], s = (S < 0 ? -T.s[I] : T.s[I]);
  const float nr = dr * c - di * s;
  di = dr * s + di * c;
  dr = nr;
}

// Radix-2 step BETWEEN the halves of every pair, decimation in frequency: pair i holds points
// (i, i + 16); afterwards its low half holds x[i] + x[i+16] (input i of the even-output DFT16) and its
// high half (x[i] - x[i+16]) exp(S 2 pi j i / 32) (input i of the odd-output DFT16).  Pairs in natural slots.
template <int S, int I = 0> FC_HD void cross_dif(Regs &v) {
  if constexpr (I < 16) {
    const float sr = v.re[I].x + v.re[I].y, si = v.im[I].x + v.im[I].y;
    float dr = v.re[I].x - v.re[I].y, di = v.im[I].x - v.im[I].y;
    rot32<S, I>(dr, di);
    v.re[I].x = sr; v.im[I].x = si;
    v.re[I].y = dr; v.im[I].y = di;
    cross_dif<S, I + 1>(v);
  }
}
// Radix-2 step between the halves, decimation in time: the pair in slot slot16(j) holds (E[j], O[j]),
// the j-th outputs of the DFT16s of the even / odd inputs; afterwards (X[j], X[j + 16]) with
// X[j] = E[j] + exp(S 2 pi j j / 32) O[j], X[j+16] = E[j] - exp(..) O[j].
template <int S, int J = 0> FC_HD void cross_dit(Regs &v) {
  if constexpr (J < 16) {
    constexpr int s = slot16(J);
    float orr = v.re[s].y, oi = v.im[s].y;
    rot32<S, J>(orr, oi);
    const float er = v.re[s].x, ei = v.im[s].x;
    v.re[s].x = er + orr; v.im[s].x = ei + oi;
    v.re[s].y = er - orr; v.im[s].y = ei - oi;
    cross_dit<S, J + 1>(v);
  }
}

// The same two steps with the lane's odd-point factor w = W_1024^l folded into the step's twiddles:
// ct[i * 32 + lane] = exp(-2 pi j i / 32) w (forward); the inverse uses the conjugates.  This replaces
// sixteen constant multiplications plus sixteen multiplications by w with sixteen multiplications.
// Table rows come in pairs: rows 2 q and 2 q + 1 of a lane are one 16-byte entry (tw_index), so the sixteen
// twiddles of a step are eight 128-bit loads.
struct alignas(16) Quad { float a, b, c, d; };
FC_HD constexpr int tw_index(int row, int lane) { return (((row >> 1) * 32 + lane) << 1) + (row & 1); }
FC_HD Quad tw_pair(const cpair *tw, int even_row, int lane) { return *reinterpret_cast<const Quad *>(tw + tw_index(even_row, lane)); }
constexpr int kTwRowCross = 4;   // rows 4 .. 19: exp(-2 pi j i / 32) W_1024^l, i = 0 .. 15
constexpr int kTwRowPowers = 19; // rows 20 .. 34: W_512^{l j}, j = 1 .. 15 (row 19 + j); row 35 pads the last pair
constexpr int kTwRows = 36;

template <bool CONJ, int I> FC_HD void cross_dif_one(Regs &v, float wr, float wim) {
  const float wi = CONJ ? -wim : wim;
  const float sr = v.re[I].x + v.re[I].y, si = v.im[I].x + v.im[I].y;
  const float dr = v.re[I].x - v.re[I].y, di = v.im[I].x - v.im[I].y;
  v.re[I].x = sr; v.im[I].x = si;
  v.re[I].y = dr * wr - di * wi;
  v.im[I].y = dr * wi + di * wr;
}
template <bool CONJ, int I = 0> FC_HD void cross_dif_tw(Regs &v, const cpair *tw, int lane) {
  if constexpr (I < 16) {
    const Quad w = tw_pair(tw, kTwRowCross + I, lane);
    cross_dif_one<CONJ, I>(v, w.a, w.b);
    cross_dif_one<CONJ, I + 1>(v, w.c, w.d);
    cross_dif_tw<CONJ, I + 2>(v, tw, lane);
  }
}
template <bool CONJ, int J> FC_HD void cross_dit_one(Regs &v, float wr, float wim) {
  constexpr int s = slot16(J);
  const float wi = CONJ ? -wim : wim;
  const float orr = v.re[s].y, oi = v.im[s].y;
  const float tr = orr * wr - oi * wi, ti = orr * wi + oi * wr;
  const float er = v.re[s].x, ei = v.im[s].x;
  v.re[s].x = er + tr; v.im[s].x = ei + ti;
  v.re[s].y = er - tr; v.im[s].y = ei - ti;
}
template <bool CONJ, int J = 0> FC_HD void cross_dit_tw(Regs &v, const cpair *tw, int lane) {
  if constexpr (J < 16) {
    const Quad w = tw_pair(tw, kTwRowCross + J, lane);
    cross_dit_one<CONJ, J>(v, w.a, w.b);
    cross_dit_one<CONJ, J + 1>(v, w.c, w.d);
    cross_dit_tw<CONJ, J + 2>(v, tw, lane);
  }
}

// ---- pair twiddles W_1024^{l k2}, k2 = 2 j + h -------------------------------------------------
// per lane: w2^1, w2^2, w2^4, w2^8 (w2 = W_512^l): table rows r = 0..3 (entry tw_index(r, l)); the other powers of
// w2 are products formed in registers (as in oa_n2048_core.cuh).  Rows 4 .. 19 of the same table hold
// the radix-2 step's twiddles with w = W_1024^l folded in: tw[(4 + i) * 32 + l] = exp(-2 pi j i / 32) w.
struct LaneTwiddles {
  cpair p[16]; // p[j] = w2^j (p[0] unused)
  cpair w;
};
// FULL: all fifteen powers come from the table (rows 20 .. 34: tw[(19 + j) * 32 + l] = w2^j) -- 15 shared loads
// instead of 4 loads and 44 scalar multiply-adds; which is better depends on which pipe has room (A/B switch).
template <bool FULL = false> FC_HD LaneTwiddles lane_twiddles(const cpair *tw, int lane) {
  using n2048::cpair_mul;
  LaneTwiddles t;
  if (FULL) {
    t.w = cpair{1.0f, 0.0f};
    t.p[0] = cpair{1.0f, 0.0f};
#pragma unroll
    for (int j = 1; j < 16; j += 2) { // rows 19 + j and 20 + j: one 128-bit load
      const Quad q = tw_pair(tw, kTwRowPowers + j, lane);
      t.p[j] = cpair{q.a, q.b};
      if (j + 1 < 16) t.p[j + 1] = cpair{q.c, q.d};
    }
    return t;
  }
  t.p[1] = tw[tw_index(0, lane)];
  t.p[2] = tw[tw_index(1, lane)];
  t.p[4] = tw[tw_index(2, lane)];
  t.p[8] = tw[tw_index(3, lane)];
  t.w = cpair{1.0f, 0.0f};
  t.p[0] = cpair{1.0f, 0.0f};
  t.p[3] = cpair_mul(t.p[1], t.p[2]);
  t.p[5] = cpair_mul(t.p[4], t.p[1]);
  t.p[6] = cpair_mul(t.p[4], t.p[2]);
  t.p[7] = cpair_mul(t.p[4], t.p[3]);
  t.p[9] = cpair_mul(t.p[8], t.p[1]);
  t.p[10] = cpair_mul(t.p[8], t.p[2]);
  t.p[11] = cpair_mul(t.p[8], t.p[3]);
  t.p[12] = cpair_mul(t.p[8], t.p[4]);
  t.p[13] = cpair_mul(t.p[8], t.p[5]);
  t.p[14] = cpair_mul(t.p[8], t.p[6]);
  t.p[15] = cpair_mul(t.p[8], t.p[7]);
  return t;
}
// the pair for index j sits in slot SLOT(j); CONJ = inverse direction.  Both points of pair j get
// w2^j; the odd point's extra factor w rides in the neighbouring radix-2 step (cross_*_tw).
template <bool CONJ, bool SLOT16> FC_HD void apply_pair_twiddles(Regs &v, const LaneTwiddles &t) {
#pragma unroll
  for (int j = 1; j < 16; ++j) {
    const int s = SLOT16 ? slot16(j) : j;
    cmul2(v.re[s], v.im[s], t.p[j].re, CONJ ? -t.p[j].im : t.p[j].im);
  }
}

// ---- exchange buffer ---------------------------------------------------------------------------
// two planes (re, im) of 32 rows x kRowStride floats; entry (row r, column c) holds the value that
// lane c produced for lane r (T1) or lane r produced for lane c (T2).
FC_HD int ex_addr(int row, int col) { return row * kRowStride + col; }

// T1 write: lane l holds B_l[k2] in slot16(j), k2 = 2 j + h  ->  entry (k2, l)
FC_HD void t1_write(int lane, const Regs &v, float *xb) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int s = slot16(j);
    xb[ex_addr(2 * j, lane)] = v.re[s].x;
    xb[ex_addr(2 * j + 1, lane)] = v.re[s].y;
    xb[kPlane + ex_addr(2 * j, lane)] = v.im[s].x;
    xb[kPlane + ex_addr(2 * j + 1, lane)] = v.im[s].y;
  }
}
// T1 read: lane k2 reads its row: pair i = (l = 2 i, 2 i + 1) into natural slot i
FC_HD void t1_read(int lane, Regs &v, const float *xb) {
  const Quad *re = reinterpret_cast<const Quad *>(xb + ex_addr(lane, 0));
  const Quad *im = reinterpret_cast<const Quad *>(xb + kPlane + ex_addr(lane, 0));
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const Quad r = re[c], q = im[c];
    v.re[2 * c] = make_f2(r.a, r.b);
    v.re[2 * c + 1] = make_f2(r.c, r.d);
    v.im[2 * c] = make_f2(q.a, q.b);
    v.im[2 * c + 1] = make_f2(q.c, q.d);
  }
}
// T2 write: lane k2 holds C_k2[l] in slot16(i), l = 2 i + h  ->  its own row, 16 bytes at a time
FC_HD void t2_write(int lane, const Regs &v, float *xb) {
  Quad *re = reinterpret_cast<Quad *>(xb + ex_addr(lane, 0));
  Quad *im = reinterpret_cast<Quad *>(xb + kPlane + ex_addr(lane, 0));
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int s0 = slot16(2 * c), s1 = slot16(2 * c + 1);
    re[c] = Quad{v.re[s0].x, v.re[s0].y, v.re[s1].x, v.re[s1].y};
    im[c] = Quad{v.im[s0].x, v.im[s0].y, v.im[s1].x, v.im[s1].y};
  }
}
// T2 read: lane l reads column l: k2 = 2 j + h into natural slot j
FC_HD void t2_read(int lane, Regs &v, const float *xb) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    v.re[j] = make_f2(xb[ex_addr(2 * j, lane)], xb[ex_addr(2 * j + 1, lane)]);
    v.im[j] = make_f2(xb[kPlane + ex_addr(2 * j, lane)], xb[kPlane + ex_addr(2 * j + 1, lane)]);
  }
}

// ---- spectrum ------------------------------------------------------------------------------------
// Lane k2 holds, after P1's forward dft16, the pair (E[j], O[j]) in slot16(j): the j-th outputs of the DFT16s of
// its even / odd inputs.  The reference's three steps -- last radix-2 step of the forward transform, multiply by
// the spectrum, first radix-2 step of the inverse (fftconv.hpp:398-400: forward, multiply_cx, backward) --
//   X_lo = E + w O, X_hi = E - w O            (w = exp(-2 pi i j / 32); frequencies k_lo = 32 j + k2, k_hi = k_lo + 512)
//   Y_lo = X_lo H_lo, Y_hi = X_hi H_hi
//   s = Y_lo + Y_hi,  d = (Y_lo - Y_hi) conj(w)
// collapse into one 2 x 2 complex matrix per (j, k2), because w conj(w) = 1:
//   s = E A + O B,   d = E C + O A,      A = H_lo + H_hi,  B = w (H_lo - H_hi),  C = conj(w) (H_lo - H_hi)
// 16 fused multiply-adds per pair instead of 8 + 8 (two packed complex products) + 8; 1 / N is folded in.  (Measured against
// tabulating only A and C and forming B = w^2 C with four more instructions per pair -- a third less table
// traffic -- : this form is 0.8 % faster, profiles/r02_w1k_session4_ab.jsonl.)
// A spectrum given as complex entries (filter handles, tests) is indexed by e = 2 (j * 32 + k2) + half:
FC_HD constexpr unsigned freq_of_index(unsigned e) {
  return 32u * ((e >> 6) + 16u * (e & 1u)) + ((e >> 1) & 31u);
}
// In registers the pair (s, d) is one packed value per component, so the eight products per component pair up:
//   (s.re, d.re) = E.re (A.re, C.re) - E.im (A.im, C.im) + O.re (B.re, A.re) - O.im (B.im, A.im)
//   (s.im, d.im) = E.re (A.im, C.im) + E.im (A.re, C.re) + O.re (B.im, A.im) + O.im (B.re, A.re)
// eight packed multiply-adds with a broadcast scalar instead of sixteen scalar ones; the table holds the four
// constant pairs per (j, k2): q0 = (A.re, C.re, A.im, C.im), q1 = (B.re, A.re, B.im, A.im) -- 32 bytes.
struct SpectrumTables { // one pointer: 512 entries q0, then 512 entries q1
  Quad *q0;
  FC_HD Quad *q1() const { return q0 + 512; }
};
// w = exp(-2 pi i j / 32): row 4 + j of the lane-twiddle table at lane 0 (oa_w1k_tables.h: fill_tw)
FC_HD void fold_spectrum_entry(cpair w, cpair h_lo, cpair h_hi, Quad &q0, Quad &q1) {
  const float wr = w.re, wi = w.im;
  const float ar = h_lo.re + h_hi.re, ai = h_lo.im + h_hi.im;
  const float dr = h_lo.re - h_hi.re, di = h_lo.im - h_hi.im;
  const float br = dr * wr - di * wi, bi = dr * wi + di * wr;   // B = w D
  const float cr = dr * wr + di * wi, ci = di * wr - dr * wi;   // C = conj(w) D
  q0 = Quad{ar, cr, ai, ci};
  q1 = Quad{br, ar, bi, ai};
}
// one pair: (E, O) in -> (s, d) out
FC_HD void spectrum_pair(int lane, int j, f2 inr, f2 ini, f2 &outr, f2 &outi, const SpectrumTables &h) {
  const Quad a = h.q0[j * 32 + lane], b = h.q1()[j * 32 + lane];
  const f2 p0 = make_f2(a.a, a.b), p1 = make_f2(a.c, a.d), p2 = make_f2(b.a, b.b), p3 = make_f2(b.c, b.d);
  const float er = inr.x, ei = ini.x, orr = inr.y, oi = ini.y;
  outr = fmas2(p0, er, fmas2(p1, -ei, fmas2(p2, orr, muls2(p3, -oi))));
  outi = fmas2(p1, er, fmas2(p0, ei, fmas2(p3, orr, muls2(p2, oi))));
}
// v: (E, O) in slot16(j)  ->  (s, d) in slot j, where the inverse dft16 wants them.  slot16 is an involution
// (a 4 x 4 transpose of the slot index), so the move happens in place: j and slot16(j) trade places.
FC_HD void multiply_spectrum_fused(int lane, Regs &v, const SpectrumTables &h) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int t = slot16(j);
    if (t == j) {
      spectrum_pair(lane, j, v.re[j], v.im[j], v.re[j], v.im[j], h);
    } else if (j < t) {
      const f2 ar = v.re[t], ai = v.im[t], br = v.re[j], bi = v.im[j]; // inputs of outputs j and t
      spectrum_pair(lane, j, ar, ai, v.re[j], v.im[j], h);
      spectrum_pair(lane, t, br, bi, v.re[t], v.im[t], h);
    }
  }
}

// ---- the phases of one warp iteration (between __syncwarp()s) -------------------------------------
// P0: slots hold z[l + 32 i] (low) and z[l + 32 (i + 16)] (high); leaves B_l[k2] in slot16(j)
FC_HD void phase0(int lane, Regs &v, const cpair *tw) {
  cross_dif_tw<false>(v, tw, lane); // the odd outputs' factor w rides in the step's twiddles
  dft16<-1>(v);
  apply_pair_twiddles<false, true>(v, lane_twiddles(tw, lane));
}
// P1: slots hold B_l[k2] for l = 2 i, 2 i + 1; leaves C_k2[l], l = 2 i + h, in slot16(i)
FC_HD void phase1(int lane, Regs &v, const SpectrumTables &h) {
  dft16<-1>(v);
  multiply_spectrum_fused(lane, v, h);
  dft16<+1>(v);
}
// The same phases cut at the two dft16 calls of each direction, so that a kernel can run ONE copy of
// the forward and ONE copy of the inverse dft16 in two-trip loops (the unrolled code of four copies
// does not fit the 32 KB instruction cache: measured 87 % hit rate, 0.64 stall cycles per issue).
//   phase0 = p0_pre, dft16<-1>, p0_post        phase1 = dft16<-1>, p1_mid, dft16<+1>
//   phase2 = p2_pre, dft16<+1>, p2_post
FC_HD void p0_pre(int lane, Regs &v, const cpair *tw) { cross_dif_tw<false>(v, tw, lane); }
template <bool FULL = false> FC_HD void p0_post(int lane, Regs &v, const cpair *tw) { apply_pair_twiddles<false, true>(v, lane_twiddles<FULL>(tw, lane)); }
// between P1's two transforms: radix-2 step, spectrum, radix-2 step of the inverse in one (multiply_spectrum_fused)
FC_HD void p1_mid(int lane, Regs &v, const SpectrumTables &h) { multiply_spectrum_fused(lane, v, h); }
template <bool FULL = false> FC_HD void p2_pre(int lane, Regs &v, const cpair *tw) { apply_pair_twiddles<true, false>(v, lane_twiddles<FULL>(tw, lane)); }
FC_HD void p2_post(int lane, Regs &v, const cpair *tw) { cross_dit_tw<true>(v, tw, lane); }

// P2: slots hold C_k2[l] for k2 = 2 j, 2 j + 1; leaves y[l + 32 m], y[l + 32 (m + 16)] in slot16(m)
FC_HD void phase2(int lane, Regs &v, const cpair *tw) {
  apply_pair_twiddles<true, false>(v, lane_twiddles(tw, lane));
  dft16<+1>(v);
  cross_dit_tw<true>(v, tw, lane); // conj(w) of the odd inputs commutes with their dft16
}

} // namespace w1k
} // namespace fc
