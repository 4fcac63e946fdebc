// oa_w1k_group.cuh -- work decomposition and I/O of the warp-per-FFT engine (oa_w1k_core.cuh).
// Host + device, so that tests/cpu/emulate_oa_w1k.cpp runs the very same code.
//
// The job is a list of ITEMS (row r, block s), s < segs = ceil((n + K - 1) / step): block s owns the
// outputs f in [s step, (s + 1) step) of the row's full convolution and reads the N = 1024 input
// samples [s step - (K - 1), s step - (K - 1) + N) -- zeros outside the row.  Two consecutive items
// ride in the real and imaginary part of one complex FFT: a UNIT.  Units are independent; warp w of
// W takes the contiguous range [w U / W, (w + 1) U / W).
//
// Input: the block's samples are staged in the warp's exchange buffer (plane 0: real part's row,
// plane 1: imaginary part's row) by one bulk asynchronous copy per item (TMA, SASS UBLKCP), requested
// by lane 0 one unit ahead.  Sample with row index r sits at staging index (r - block start) + shift,
// shift in 0..3 chosen so that source and destination of the copy are both 16-byte aligned.  The first
// block of a row (K - 1 zeros in front) is copied from the row start and the lanes zero the head; a
// block that reaches past the row end, a misaligned row start, or a span that would leave the tensor
// is filled by the lanes with plain loads instead.
// Output: y[i], i = lane + 32 m, is output f = block start + i for i >= K - 1: 4-byte stores, one full
// 128-byte line per warp instruction.
#pragma once

#include "oa_w1k_core.cuh"

#include <cstddef>

namespace fc {
namespace w1k {

struct W1kParams {
  const float *in;
  float *out;
  long long batch, n, in_stride, out_stride, out_len;
  int klen, step, pad; // step = N - (K - 1); pad = 0 (Full) or K / 2 (Same)
  long long full_len;  // n + K - 1
  long long segs;      // blocks per row
  long long n_items;   // batch * segs
  long long n_units;   // ceil(n_items / 2)
  const cpair *tw;     // kTwElems per-lane twiddles
  const cpair *hs;     // 1024 spectrum entries (freq_of_index order), 1/N folded in; or nullptr:
  const float *taps;   // ... the kernel transforms the taps itself (klen device floats), see spectrum_by_warp
  int cta_major;       // A/B switch: deal unit ranges CTA by CTA (the first version) instead of warp-major
  // Dynamic tail (kernel only): warp w owns units [w S, (w + 1) S), S = static_per_warp, and then takes one unit at a
  // time from the pool [pool_first, n_units) through an atomic counter, so that SMs that run a few per cent faster
  // than others (measured: 934 k .. 972 k active cycles of 1 M) take more units.  sched[0]: next pool unit,
  // sched[1]: warps that have finished; both return to zero at the end of every launch.  nullptr: fixed ranges.
  unsigned *sched;
  long long static_per_warp, pool_first;
};

inline void fill_params(W1kParams &p, const float *in, float *out, long long batch, long long n, long long in_stride,
                        long long out_stride, long long out_len, int klen, int mode) {
  p.in = in; p.out = out;
  p.batch = batch; p.n = n;
  p.in_stride = in_stride; p.out_stride = out_stride; p.out_len = out_len;
  p.klen = klen;
  p.step = kN - (klen - 1);
  p.pad = mode ? klen / 2 : 0;
  p.full_len = n + klen - 1;
  p.segs = (p.full_len + p.step - 1) / p.step;
  p.n_items = batch * p.segs;
  p.n_units = (p.n_items + 1) / 2;
  p.tw = nullptr;
  p.hs = nullptr;
  p.taps = nullptr;
  p.cta_major = 0;
  p.sched = nullptr;
  p.static_per_warp = 0;
  p.pool_first = 0;
}

FC_HD void warp_range(long long n_units, long long n_warps, long long w, long long &u0, long long &u1) {
  const long long base = n_units / n_warps, rem = n_units % n_warps;
  u0 = w * base + (w < rem ? w : rem);
  u1 = u0 + base + (w < rem ? 1 : 0);
}

struct ItemDesc {
  const float *row_in;  // the row's first sample
  float *row_out;       // the row's first output
  long long bs;         // row index of the block's first input sample (negative for a row's first block)
  const float *src;     // bulk copy: 16-byte aligned source, 0 bytes = nothing to copy
  int bytes;
  int dst_first;        // staging index (floats) the copy starts at
  int zero_upto;        // the lanes zero staging [0, zero_upto) (row's first block)
  int shift;            // sample r sits at staging index (r - bs) + shift
  int nout;             // outputs owned: y[K-1 .. K-1+nout)
  int active;
  int bulk_ok;
};
struct UnitDesc {
  ItemDesc item[2];
  int bulk;     // both items are staged by bulk copies
  int fast_out; // both items own `step` outputs that all lie inside their rows
};

FC_HD int align_shift4(const float *ptr) { return (int)((reinterpret_cast<unsigned long long>(ptr) >> 2) & 3ULL); }

// What the kernel needs to stage one block by a bulk copy and to store its outputs, worked out one unit ahead
// (the kernel and tests/cpu/emulate_oa_w1k.cpp both use this).  Besides interior blocks this covers
//   * a row's first block: copied from the row start, the lanes zero the K - 1 samples in front (head_zero);
//   * a block that reaches past the row's end (every row's last block; on short rows -- 8192-sample A-lines --
//     one block in eleven): the copy stops at the row's end (rounded up to 16 bytes: it may bring up to three
//     floats of whatever follows the row) and the lanes zero the staging area from `zero_from` on AFTER the
//     copy has landed.
// Still staged by the lanes with plain loads (manual_fill): misaligned first blocks, copies that would leave the
// tensor (the last row's end when it is not a multiple of 16 bytes), inactive items.
struct ItemReq {
  const float *src; // 16-byte aligned
  int bytes;        // multiple of 16
  int dst_first;    // staging index the copy starts at (multiple of 4)
  int shift;        // sample r sits at staging index (r - block start) + shift
  int head_zero;    // the lanes zero staging [0, head_zero) at request time
  int zero_from;    // the lanes zero staging [zero_from, kN + 4) after the copy has landed (kN + 4: nothing)
  bool in_ok, out_ok;
};
FC_HD ItemReq item_request(const W1kParams &p, int row, int seg, const float *tensor_end) {
  ItemReq r;
  const int tail = p.klen - 1;
  const long long bs = (long long)seg * p.step - tail;
  const int head = seg == 0 ? tail : 0; // a row's first block: K - 1 zeros in front
  const float *first = p.in + (long long)row * p.in_stride + (bs + head);
  const int ish = align_shift4(first);
  r.shift = (ish - head) & 3;
  r.dst_first = head + r.shift - ish;
  r.src = first - ish;
  const long long avail = p.n - (bs + head); // samples of the row from `first` on (>= 1 for every block)
  const int want = kN - head;
  const int nval = avail < want ? (int)avail : want;
  r.bytes = ((ish + nval) * 4 + 15) & ~15;
  r.head_zero = head > 0 ? r.dst_first : 0;
  r.zero_from = nval < want ? head + r.shift + nval : kN + 4;
  const bool active = row < p.batch;
  r.in_ok = active && (head == 0 || ish == 0) && r.src + r.bytes / 4 <= tensor_end && r.src >= p.in;
  const long long o0 = bs + tail - p.pad; // output index of the block's first kept sample
  r.out_ok = active && (long long)(seg + 1) * p.step <= p.full_len && o0 >= 0 && o0 + p.step <= p.out_len;
  return r;
}
// Blocks 1 <= seg <= interior_seg_hi(p) of any row are INTERIOR: item_request gives them in_ok, out_ok, head_zero = 0,
// zero_from = kN + 4, shift = the source's misalignment, dst_first = 0, bytes = 4 kN (+ 16 when misaligned).  Why:
// the block reads row samples [bs, bs + kN) with bs = seg step - (K - 1) >= step - (K - 1) > 0 and
// bs + kN + 3 <= n, so the copy -- rounded out to 16-byte boundaries, at most 3 floats either side -- stays inside
// the row; its outputs are [seg step, (seg + 1) step) with (seg + 1) step = bs + kN <= n - 3 <= out_len + pad - 3.
// (tests/cpu/emulate_oa_w1k.cpp checks the claim against item_request for every block of every case.)
FC_HD int interior_seg_hi(const W1kParams &p) {
  const long long hi = (p.n + (p.klen - 1) - kN - 3) / p.step; // floor for non-negative numerators; negative: none
  if (p.n + (p.klen - 1) - kN - 3 < 0) return 0;
  return (int)(hi < 0x7fffffff ? hi : 0x7fffffff);
}

// lanes: zeros behind a block that was clipped at its row's end (after the bulk copy has landed)
FC_HD void zero_tail(int lane, int zero_from, float *plane) {
  for (int i = zero_from + lane; i < kN + 4; i += 32) plane[i] = 0.0f;
}

FC_HD ItemDesc describe_item(const W1kParams &p, long long row, long long seg) {
  ItemDesc d;
  d.active = row < p.batch;
  if (!d.active) { row = 0; seg = 0; }
  d.row_in = p.in + row * p.in_stride;
  d.row_out = p.out + row * p.out_stride;
  d.bs = seg * p.step - (p.klen - 1);
  const long long left = p.full_len - seg * p.step;
  d.nout = d.active ? (int)(left < p.step ? left : p.step) : 0;
  // valid input rows [r0, r1)
  const long long r0 = d.bs < 0 ? 0 : d.bs;
  long long r1 = d.bs + kN;
  const bool tail_clipped = r1 > p.n;
  if (tail_clipped) r1 = p.n;
  const float *first = d.row_in + r0;
  const int ish = align_shift4(first);
  const int head = (int)(r0 - d.bs); // zeros in front
  d.shift = (ish - head) & 3;
  d.dst_first = head + d.shift - ish;
  d.src = first - ish;
  const long long nvalid = r1 > r0 ? r1 - r0 : 0;
  d.bytes = (int)(((ish + nvalid) * 4 + 15) & ~15LL);
  d.zero_upto = head > 0 ? d.dst_first : 0;
  const float *tensor_end = p.in + (p.batch - 1) * p.in_stride + p.n;
  d.bulk_ok = d.active && !tail_clipped && nvalid > 0 && (head == 0 || ish == 0) &&
              (d.src + d.bytes / 4 <= tensor_end) && (d.src >= p.in);
  return d;
}

FC_HD UnitDesc describe_unit_at(const W1kParams &p, long long row0, long long seg0, long long row1, long long seg1) {
  UnitDesc u;
  u.item[0] = describe_item(p, row0, seg0);
  u.item[1] = describe_item(p, row1, seg1);
  u.bulk = u.item[0].bulk_ok && u.item[1].bulk_ok;
  u.fast_out = 1;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const ItemDesc &d = u.item[i];
    const long long o0 = d.bs - p.pad + (p.klen - 1); // output index of y[K - 1]
    u.fast_out = u.fast_out && d.active && d.nout == p.step && o0 >= 0 && o0 + p.step <= p.out_len;
  }
  return u;
}
FC_HD UnitDesc describe_unit(const W1kParams &p, long long u) {
  const long long it0 = 2 * u, it1 = 2 * u + 1;
  const long long r0 = it0 / p.segs, r1 = it1 < p.n_items ? it1 / p.segs : p.batch;
  return describe_unit_at(p, r0, it0 - r0 * p.segs, r1, it1 < p.n_items ? it1 - r1 * p.segs : 0);
}

// lanes: zero the head of a row's first block (bulk path)
FC_HD void zero_fill(int lane, const ItemDesc &d, float *plane) {
  for (int i = lane; i < d.zero_upto; i += 32) plane[i] = 0.0f;
}
// lanes: stage one block with plain loads (rare: row ends, misaligned row starts, inactive items)
FC_HD void manual_fill(int lane, const W1kParams &p, const ItemDesc &d, float *plane) {
  for (int i = lane; i < kN + 4; i += 32) {
    const long long r = d.bs + (i - d.shift);
    float v = 0.0f;
    if (d.active && i >= d.shift && r >= 0 && r < p.n) v = d.row_in[r];
    plane[i] = v;
  }
}

// staged samples -> registers: pair i = (z[lane + 32 i], z[lane + 32 (i + 16)]), re from plane 0, im from plane 1
FC_HD void load_inputs(int lane, Regs &v, const float *xb, int shift_a, int shift_b) {
  const float *a = xb + shift_a + lane, *b = xb + kPlane + shift_b + lane;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v.re[i] = make_f2(a[32 * i], a[32 * (i + 16)]);
    v.im[i] = make_f2(b[32 * i], b[32 * (i + 16)]);
  }
}

// y[lane + 32 m] is kept iff lane + 32 m >= K - 1 (the block's first K - 1 outputs are the wrapped part: dropped),
// i.e. iff 32 m >= keep_from with keep_from = K - 1 - lane: one register, one compare against an immediate per store.
// (The first version kept a bit mask; the compiler turned it into nine loop-invariant registers -- and spills.)
struct StoreMasks {
  int keep_from;
};
FC_HD StoreMasks make_store_masks(int lane, const W1kParams &p) {
  StoreMasks k;
  k.keep_from = p.klen - 1 - lane;
  return k;
}

// y[lane + 32 m] sits in slot16(m & 15), low half for m < 16; ya / yb: where y[0] of the two items goes
FC_HD void store_fast_to(int lane, const Regs &v, float *ya, float *yb, const StoreMasks &k) {
  float *a = ya + lane, *b = yb + lane;
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const int s = slot16(m);
    if (32 * m >= k.keep_from) {
      a[32 * m] = v.re[s].x;
      b[32 * m] = v.im[s].x;
    }
    // m + 16 >= 16: lane + 512 >= K - 1 for every K this engine serves
    a[32 * (m + 16)] = v.re[s].y;
    b[32 * (m + 16)] = v.im[s].y;
  }
}
// one item of a unit (IT = 0: real parts, 1: imaginary parts); y: where y[0] of the item goes
template <int IT> FC_HD void store_item_fast(int lane, const Regs &v, float *y, const StoreMasks &k) {
  float *a = y + lane;
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const int s = slot16(m);
    if (32 * m >= k.keep_from) a[32 * m] = IT ? v.im[s].x : v.re[s].x;
    a[32 * (m + 16)] = IT ? v.im[s].y : v.re[s].y; // lane + 512 >= K - 1 for every K this engine serves
  }
}
// predicated stores of one item given by its (row, block): row ends, Same-mode edges
template <int IT> FC_HD void store_item_slow(int lane, const Regs &v, const W1kParams &p, long long row, long long seg) {
  if (row >= p.batch) return;
  float *row_out = p.out + row * p.out_stride;
  const long long left = p.full_len - seg * p.step;
  const int nout = (int)(left < p.step ? left : p.step);
  const long long obase = seg * p.step - (p.klen - 1) - p.pad; // output index of y[0]
  // y[i] is stored iff K - 1 <= i < K - 1 + nout and 0 <= obase + i < out_len: one 32-bit range [lo, hi)
  long long lo = p.klen - 1, hi = p.klen - 1 + nout;
  if (-obase > lo) lo = -obase;
  if (p.out_len - obase < hi) hi = p.out_len - obase;
  const int ilo = (int)lo, ihi = (int)(hi > lo ? hi : lo);
  float *a = row_out + obase + lane;
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const int s = slot16(m);
    const int i0 = lane + 32 * m, i1 = i0 + 512;
    if (i0 >= ilo && i0 < ihi) a[32 * m] = IT ? v.im[s].x : v.re[s].x;
    if (i1 >= ilo && i1 < ihi) a[32 * (m + 16)] = IT ? v.im[s].y : v.re[s].y;
  }
}
FC_HD void store_fast(int lane, const Regs &v, const UnitDesc &d, const W1kParams &p, const StoreMasks &k) {
  store_fast_to(lane, v, d.item[0].row_out + (d.item[0].bs - p.pad), d.item[1].row_out + (d.item[1].bs - p.pad), k);
}
// predicated stores of a unit given by its items' (row, block): what the output side needs is little
// (no alignment, no staging), so this does not go through describe_item
FC_HD void store_slow_at(int lane, const Regs &v, const W1kParams &p, long long row0, long long seg0, long long row1,
                         long long seg1) {
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const long long row = it ? row1 : row0, seg = it ? seg1 : seg0;
    if (row >= p.batch) continue;
    float *row_out = p.out + row * p.out_stride;
    const long long left = p.full_len - seg * p.step;
    const int nout = (int)(left < p.step ? left : p.step);
    const long long obase = seg * p.step - (p.klen - 1) - p.pad; // output index of y[0]
    const int lo = p.klen - 1, hi = p.klen - 1 + nout;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int s = slot16(m);
      const int i0 = lane + 32 * m, i1 = i0 + 512;
      const float y0 = it ? v.im[s].x : v.re[s].x, y1 = it ? v.im[s].y : v.re[s].y;
      const long long o0 = obase + i0, o1 = obase + i1;
      if (i0 >= lo && i0 < hi && o0 >= 0 && o0 < p.out_len) row_out[o0] = y0;
      if (i1 >= lo && i1 < hi && o1 >= 0 && o1 < p.out_len) row_out[o1] = y1;
    }
  }
}
FC_HD void store_slow(int lane, const Regs &v, const W1kParams &p, const UnitDesc &d) {
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const ItemDesc &I = d.item[it];
    if (!I.active) continue;
    const long long obase = I.bs - p.pad; // output index of y[0]
    const int lo = p.klen - 1, hi = p.klen - 1 + I.nout;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int s = slot16(m);
      const int i0 = lane + 32 * m, i1 = i0 + 512;
      const float y0 = it ? v.im[s].x : v.re[s].x, y1 = it ? v.im[s].y : v.re[s].y;
      const long long o0 = obase + i0, o1 = obase + i1;
      if (i0 >= lo && i0 < hi && o0 >= 0 && o0 < p.out_len) I.row_out[o0] = y0;
      if (i1 >= lo && i1 < hi && o1 >= 0 && o1 < p.out_len) I.row_out[o1] = y1;
    }
  }
}

// ---- the kernel spectrum by the engine's own forward transform ----------------------------------------------
// The reference transforms the (zero-padded) taps in every call (fftconv.hpp:391-392).  Here one warp of every
// CTA does it with the forward half of a unit -- the taps as the real part of a block, P0, T1, P1's dft16 and
// radix-2 step -- which leaves lane k2 with exactly the frequencies, in exactly the registers, in which
// multiply_spectrum_fused consumes them (its tables follow from E and O directly).  No separate spectrum
// kernel, no launch gap in front of the segment kernel (measured: 11 us of a 90 us step at 512 rows per GPU).
// stage: lanes fill the warp's buffer with the taps (plane 0) and zeros
FC_HD void spectrum_stage(int lane, const float *taps, int klen, float *xb) {
  for (int i = lane; i < kN + 4; i += 32) {
    xb[i] = i < klen ? taps[i] : 0.0f;
    xb[kPlane + i] = 0.0f;
  }
}
// after load_inputs, phase0, t1_write / t1_read: the forward dft16 leaves (E[j], O[j]) of the taps' transform in
// slot16(j), and the folded tables of multiply_spectrum_fused follow without the radix-2 step: with
// H_lo, H_hi = (E +- w O) / N:  A = 2 E / N,  C = conj(w) (H_lo - H_hi) = 2 O / N,  B = w^2 C,  w^2 = exp(-2 pi i j / 16).
template <int J = 0> FC_HD void spectrum_emit(int lane, const Regs &v, const SpectrumTables &h) {
  if constexpr (J < 16) {
    constexpr int s = slot16(J);
    constexpr float sc = 2.0f / (float)kN;
    const float ar = v.re[s].x * sc, ai = v.im[s].x * sc, cr = v.re[s].y * sc, ci = v.im[s].y * sc;
    float br = cr, bi = ci;
    rot32<-1, (2 * J) & 15>(br, bi);
    if (J >= 8) { br = -br; bi = -bi; }
    h.q0[J * 32 + lane] = Quad{ar, cr, ai, ci};
    h.q1()[J * 32 + lane] = Quad{br, ar, bi, ai};
    spectrum_emit<J + 1>(lane, v, h);
  }
}
FC_HD void spectrum_finish(int lane, Regs &v, const SpectrumTables &h) {
  dft16<-1>(v);
  spectrum_emit(lane, v, h);
}

} // namespace w1k
} // namespace fc
