// oa_n2048_group.cuh -- work decomposition and I/O phases around the N = 2048
// segment engine (oa_n2048_core.cuh).  Host+device, so tests can emulate it.
//
// Reference semantics being replaced: FFTConvEngine::oaconvolve
// (/root/reference/include/fftconv/fftconv.hpp:382-437): zero `out`, then for
// every block pos = 0, step, 2 step, ... add y = ifft(fft(pad(a[pos:pos+len])) H)/N
// into out at pos (Full) or pos - K/2 (Same), clipped to out.size().
//
// Here nothing is zeroed and nothing is accumulated in global memory.  A "stream" is one row (or,
// when there are fewer rows than lanes, one of a few "virtual rows" a row is cut into); NL streams
// form a quad that one 128-thread group walks in lockstep, block after block.  Inside a walk the
// K-1 sample tail of block s stays in shared memory and is added to the head of block s+1 before
// that sample is stored, so every output sample is written exactly once, with at most two addends
// -> deterministic, no atomics.
//
// Work is divided among the groups of the grid in equal numbers of BLOCKS, not rows: a group's
// slice may begin in the middle of a row.  It then starts K-1 samples early and does not store the
// first K-1 outputs of its first block -- they lack the contributions of the samples before the
// block and belong to the previous slice, which in turn drops its last tail.  (Round 1 instead
// replayed the predecessor block with stores off: one whole extra block per slice, 12 % of the
// work when a slice is only 8 blocks long, e.g. config 2 sharded over 8 GPUs.)  Because a start
// in mid-row costs K-1 samples, slice boundaries do not fall on a regular grid; the host walks the
// same recurrence once per (batch, n, K, grid) and keeps the table of slice starts on the device
// (oa_n2048_tables.h: plan_slices -- part of the plan cache).
//
// Global memory is touched in two ways:
//   * fast path (all four lanes hold a full interior segment): the segment is
//     staged through shared memory by bulk asynchronous copies (TMA, 16-byte
//     granularity, issued by one thread, prefetched one segment ahead); threads
//     only do 4-byte shared-memory accesses with immediate offsets.  Misaligned
//     rows are handled by copying the enclosing 16-byte-aligned span and
//     indexing it with a per-lane shift.
//   * slow path (first / last / ragged segments, partly filled quads): plain
//     predicated 4-byte global loads and stores.
#pragma once

#include "oa_n2048_core.cuh"

#include <cstddef>

namespace fc {
namespace n2048 {

// S = scalar type (float: four streams per quad, double: two).  The word "quad" is kept for
// the set of NL streams that one group transforms together.
struct SlotStart {      // the slice of one group
  long long quad;       // quad index; >= n_quads: nothing to do
  long long u;          // first sample (local to the quad's streams) whose output the slice owns
  long long budget;     // blocks in the slice
};

template <typename S> struct OaParamsT {
  static constexpr int NL = DataTraits<typename DataOf<S>::D>::NL;
  const S *in;
  S *out;
  long long batch;      // rows
  long long n;          // samples per row
  long long in_stride;  // elements between rows
  long long out_stride;
  long long out_len;    // n + K - 1 (Full) or n (Same)
  int klen;             // K
  int step;             // N - (K - 1)
  int pad;              // 0 (Full) or K / 2 (Same)
  long long vrows;      // virtual rows per row (1 unless batch < NL)
  long long vlen;       // samples per virtual row (the last of a row may be shorter)
  long long n_streams;  // batch * vrows
  long long n_quads;    // ceil(n_streams / NL)
  long long n_slots;    // slices = groups of the grid
  const SlotStart *slots; // n_slots entries (device memory for the kernel)
  const CPairT<S> *tw1, *tw2, *hs; // global tables (see oa_n2048_core.cuh)
};
using OaParams = OaParamsT<float>;

template <typename S> struct LaneT {
  const S *in;          // the ROW's first sample
  S *out;               // the row's first output
  long long a0;         // first sample of the stream inside its row
  long long len;        // samples of the stream
  bool active;
  bool row_end;         // the stream ends where the row ends: its last block also owns the final tail
};
using Lane = LaneT<float>;

template <typename S> FC_HD void make_lanes(const OaParamsT<S> &p, long long quad, LaneT<S> (&ln)[4]) {
  constexpr int NL = OaParamsT<S>::NL;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const long long sid = quad * NL + l;
    ln[l].active = l < NL && sid < p.n_streams;
    const long long row = ln[l].active ? sid / p.vrows : 0;
    const long long v = ln[l].active ? sid - row * p.vrows : 0;
    ln[l].a0 = v * p.vlen;
    long long len = p.n - ln[l].a0;
    if (len > p.vlen) len = p.vlen;
    if (len <= 0) { len = 0; ln[l].active = false; }
    ln[l].len = ln[l].active ? len : 0;
    ln[l].row_end = ln[l].a0 + len >= p.n;
    ln[l].in = p.in + row * p.in_stride;
    ln[l].out = p.out + row * p.out_stride;
  }
}
// samples the quad's longest stream has, and whether any of its streams starts in mid-row
template <typename S> FC_HD long long quad_len(const LaneT<S> (&ln)[4]) {
  long long m = 0;
#pragma unroll
  for (int l = 0; l < 4; ++l) m = (ln[l].active && ln[l].len > m) ? ln[l].len : m;
  return m;
}
template <typename S> FC_HD bool quad_has_prev(const LaneT<S> (&ln)[4]) {
  bool w = false;
#pragma unroll
  for (int l = 0; l < 4; ++l) w = w || (ln[l].active && ln[l].a0 > 0);
  return w;
}

// ---- slow path: direct global access -------------------------------------------
// Load x[t + 128 m] of the block that starts at local sample u (may be negative: the K-1 samples
// before a stream that begins a row are zeros) of the four lanes into the packed registers.
//   lane 0 -> re, half A    lane 1 -> im, half A
//   lane 2 -> re, half B    lane 3 -> im, half B
template <typename D, typename S>
FC_HD void io_load(int t, RegsT<D> &v, const OaParamsT<S> &p, const LaneT<S> (&ln)[4], long long u) {
  constexpr int NL = DataTraits<D>::NL;
  S x[NL][kPts];
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    const bool on = ln[l].active && u < ln[l].len;
    const long long pos = ln[l].a0 + u;   // row coordinate of x[0]
    // valid block-local indices: [lo, hi)
    long long lo = on ? -pos : (long long)kN, hi = on ? p.n - pos : 0;
    if (lo < 0) lo = 0;
    if (hi > p.step) hi = p.step;
    const S *src = ln[l].in + (on ? pos : 0);
    const int ilo = (int)lo, ihi = (int)hi;
#pragma unroll
    for (int m = 0; m < kPts; ++m) {
      const int i = t + 128 * m;
#if defined(__CUDA_ARCH__)
      x[l][m] = (i >= ilo && i < ihi) ? __ldg(src + i) : (S)0;
#else
      x[l][m] = (i >= ilo && i < ihi) ? src[i] : (S)0;
#endif
    }
  }
#pragma unroll
  for (int m = 0; m < kPts; ++m) {
    S xm[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) xm[l] = x[l][m];
    DataTraits<D>::pack(xm, v.re[m], v.im[m]);
  }
}

// Which y[i] of the block at local start u a lane owns: [lo, hi).  The first block of a slice that
// starts in mid-row (skip) leaves its first K-1 outputs to the previous slice; a stream's last block
// stops at the stream's end, plus the final K-1 tail where that is the row's end.
template <typename S>
FC_HD void owned_range(const OaParamsT<S> &p, const LaneT<S> &ln, long long u, bool skip, int &lo, int &hi) {
  const bool on = ln.active && u < ln.len;
  lo = skip ? p.klen - 1 : 0;
  long long h = p.step;
  if (u + p.step >= ln.len) {
    h = ln.len - u + (ln.row_end ? p.klen - 1 : 0);
    if (h > kN) h = kN;
  }
  hi = on ? (int)h : 0;
}

// v holds y[t + 128 m] in slot slot16(m).  Adds the incoming tail, saves the
// outgoing tail, stores the samples this block owns.
template <typename D, typename S>
FC_HD void io_store(int t, const RegsT<D> &v, const OaParamsT<S> &p, const LaneT<S> (&ln)[4], long long u,
                    const PtT<D> *tail_prev, PtT<D> *tail_next, bool add_tail, bool skip) {
  constexpr int NL = DataTraits<D>::NL;
  int lo[NL], hi[NL];
  long long obase[NL]; // output index of y[0]
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    owned_range(p, ln[l], u, skip, lo[l], hi[l]);
    obase[l] = ln[l].a0 + u - p.pad;
  }
  const int tail_len = p.klen - 1;
#pragma unroll
  for (int m = 0; m < kPts; ++m) {
    const int i = t + 128 * m;
    const int s = slot16(m);
    D yr = v.re[s], yi = v.im[s];
    if (add_tail && i < tail_len) {
      D tr, ti;
      get(tail_prev, i, tr, ti);
      yr = add2(yr, tr);
      yi = add2(yi, ti);
    }
    if (i >= p.step) put(tail_next, i - p.step, yr, yi);
    S y[NL];
    DataTraits<D>::unpack(yr, yi, y);
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const long long o = obase[l] + i;
      if (i >= lo[l] && i < hi[l] && o >= 0 && o < p.out_len) ln[l].out[o] = y[l];
    }
  }
}

// ---- fast path: staged through shared memory ------------------------------------
// staging buffers hold NL lanes of `ls` elements each, ls = (A - 1 + step + A - 1) & ~(A - 1) with
// A = 16 / sizeof(S) elements per 16-byte unit (oa_n2048_kernel.cuh)

// elements between `ptr` and the 16-byte boundary below it
FC_HD int align_shift(const float *ptr) { return (int)((reinterpret_cast<unsigned long long>(ptr) >> 2) & 3ULL); }
FC_HD int align_shift(const double *ptr) { return (int)((reinterpret_cast<unsigned long long>(ptr) >> 3) & 1ULL); }

// Per-thread constants of the fast path (depend on t, K only).
struct ThreadMasks {
  unsigned in_zero;   // bit m: t + 128 m >= step  (padding region: value is 0, no load)
  unsigned tail_add;  // bit m: t + 128 m <  K - 1 (incoming tail is added)
};
template <typename S> FC_HD ThreadMasks make_masks(int t, const OaParamsT<S> &p) {
  ThreadMasks k;
  k.in_zero = 0;
  k.tail_add = 0;
#pragma unroll
  for (int m = 0; m < kPts; ++m) {
    const int i = t + 128 * m;
    if (i >= p.step) k.in_zero |= 1u << m;
    if (i < p.klen - 1) k.tail_add |= 1u << m;
  }
  return k;
}

// staged input -> registers.  `in` is the group's input staging buffer.
// The kernel is specialised for the table row it was built for, 8 <= K <= 410 (WIDE = false): the
// incoming tail touches only y[i], i < 512 (register slots m < 4), the zero padding / outgoing tail
// only i >= 1536 (m >= 12).  WIDE = true serves 411 <= K <= 1024 on the same 2048-point segments
// (m < 8 / m >= 8): for large jobs that beats the reference's larger segment on the generic engine
// (fftconv_cuda.cu, throughput_oa_size).  1024 is the end: the tail (K - 1 samples) must not be
// longer than the step (2049 - K), or it would reach past the next segment and the one-segment
// warm-up of a stream that starts in mid-row would not recover it.
template <bool WIDE> FC_HD constexpr int tail_in_slots() { return WIDE ? 8 : 4; }
template <bool WIDE> FC_HD constexpr int pad_first_slot() { return WIDE ? 8 : 12; }
constexpr int kMaxTapsNarrow = 410, kMaxTapsWide = 1024;

template <bool WIDE = false, typename D, typename S>
FC_HD void fast_load(int t, RegsT<D> &v, const S *in, int ls, const int (&shift)[4], unsigned in_zero) {
  constexpr int NL = DataTraits<D>::NL;
  const S *b[NL];
#pragma unroll
  for (int l = 0; l < NL; ++l) b[l] = in + l * ls + shift[l] + t;
#pragma unroll
  for (int m = 0; m < kPts; ++m) {
    // step >= 2048 - 409 (narrow): only m >= 12 can fall into the padding region
    const bool z = (m >= pad_first_slot<WIDE>()) && ((in_zero >> m) & 1u);
    S x[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) x[l] = (S)0;
    if (!z) {
#pragma unroll
      for (int l = 0; l < NL; ++l) x[l] = b[l][128 * m];
    }
    DataTraits<D>::pack(x, v.re[m], v.im[m]);
  }
}

// registers -> the lanes' outputs in GLOBAL memory (+ tail hand-over).  b_l points at y[t] of
// lane l (one coalesced 128-byte line per warp instruction); element y[t + 128 m] goes to b_l[128 m].
// b[l] points at y[t] of lane l (entries l >= NL are ignored)
template <bool WIDE = false, typename D, typename S>
FC_HD void fast_store_to(int t, const RegsT<D> &v, S *const (&b)[4], const ThreadMasks &k, const PtT<D> *tail_prev,
                         PtT<D> *tail_next, bool add_tail, bool skip_head, int step) {
  constexpr int NL = DataTraits<D>::NL;
#pragma unroll
  for (int m = 0; m < kPts; ++m) {
    const int i = t + 128 * m;
    const int s = slot16(m);
    D yr = v.re[s], yi = v.im[s];
    if (m < tail_in_slots<WIDE>()) { // K - 1 <= 409 < 512 (narrow)
      if (add_tail && ((k.tail_add >> m) & 1u)) {
        D tr, ti;
        get(tail_prev, i, tr, ti);
        yr = add2(yr, tr);
        yi = add2(yi, ti);
      }
    }
    const bool in_tail = (m >= pad_first_slot<WIDE>()) && ((k.in_zero >> m) & 1u); // i >= step
    // the first block of a slice that starts in mid-row: y[i], i < K - 1, belongs to the previous slice
    const bool head = skip_head && m < tail_in_slots<WIDE>() && ((k.tail_add >> m) & 1u);
    if (in_tail) {
      put(tail_next, i - step, yr, yi);
    } else if (!head) {
      S y[NL];
      DataTraits<D>::unpack(yr, yi, y);
#pragma unroll
      for (int l = 0; l < NL; ++l) b[l][128 * m] = y[l]; // (streaming st.global.cs was measured: 1 % slower)
    }
  }
}

// ---- uniform (per group) control flow ---------------------------------------------
struct IterDesc {
  long long quad;
  long long u;    // local sample at which the block's input starts (negative: zeros before a row)
  bool skip;      // first block of a slice that starts in mid-row: its first K-1 outputs are not stored
  bool first;     // first block of a walk: no incoming tail
};

// Enumerates the blocks of one slice: `budget` blocks from (quad, ua) on, quad after quad.  The host
// runs the same recurrence to find where every slice starts (oa_n2048_tables.h, plan_slices).
struct WorkIter {
  long long quad, ua, u, qlen;
  int budget;
  bool new_walk;

  FC_HD void init(long long quad0, long long ua0, int budget_) {
    quad = quad0;
    ua = ua0;
    u = qlen = 0;
    budget = budget_;
    new_walk = true;
  }
  // returns false when the slice is exhausted; `ln` is refreshed on a new quad
  template <typename S> FC_HD bool next(const OaParamsT<S> &p, LaneT<S> (&ln)[4], IterDesc &d) {
    if (budget <= 0 || quad >= p.n_quads) return false;
    if (new_walk) {
      make_lanes(p, quad, ln);
      qlen = quad_len(ln);
      const bool overlap = ua > 0 || quad_has_prev(ln);
      u = overlap ? ua - (p.klen - 1) : 0;
      d.first = true;
      d.skip = overlap;
      new_walk = false;
    } else {
      d.first = false;
      d.skip = false;
    }
    d.quad = quad;
    d.u = u;
    u += p.step;
    --budget;
    if (u >= qlen) { // the quad's streams are exhausted: the slice goes on with the next quad
      ++quad;
      ua = 0;
      new_walk = true;
    }
    return true;
  }
};

template <typename S> struct SchedSlotT { // private scheduler state of one warp leader
  WorkIter it;
  LaneT<S> ln[4];
  long long pad_;
};
using SchedSlot = SchedSlotT<float>;

// tail buffers: two (ping-pong) arrays of tail_pts(K) points per group
FC_HD int tail_pts(int klen) { return ((klen - 1) + 7) & ~7; }

template <typename S> struct PacketT { // one iteration of one group, one lane per warp leader
  int valid;           // 0: the slice is exhausted
  int skip;            // the first K-1 outputs of this block are not stored (IterDesc::skip)
  int first;           // no incoming tail
  int in_fast[4];      // per lane; the quad takes the fast path when all lanes are set
  int out_fast[4];     //   (slots >= NL of a two-lane quad are preset to 1)
  int in_shift[4];
  int out_shift[4];
  int in_bytes[4];     // bulk-load size per lane (fast path)
  const S *in_src[4];
  S *out_dst[4];
  LaneT<S> ln[4];      // slow path only
  long long u;
};
using Packet = PacketT<float>;

struct GroupLayout {   // byte offsets inside a group's shared-memory block
  int lane_stride;     // floats per lane of the staging area (which aliases xb)
  int xb, tails, packets, sched, mbar, total;
};
// The input and output staging areas alias the exchange buffer: NL lanes x lane_stride
// elements, e.g. 4 * 1832 * 4 B = 29312 B (float, K = 245) <= 32768 B.
template <typename S = float> FC_HD GroupLayout group_layout(int klen) {
  constexpr int A = 16 / (int)sizeof(S);
  GroupLayout g;
  const int step = kN - (klen - 1);
  g.lane_stride = (A - 1 + step + A - 1) & ~(A - 1);
  int off = 0;
  g.xb = off;      off += kN * 16;
  g.tails = off;   off += 2 * tail_pts(klen) * 16;
  g.packets = off; off += 3 * (((int)sizeof(PacketT<S>) + 15) & ~15);
  g.sched = off;   off += 4 * (int)sizeof(SchedSlotT<S>);
  g.mbar = off;    off += 16;
  g.total = off;
  return g;
}
// tw1 rows kept in shared memory: all 16, or with twiddle recomputation only k2 = 1, 2, 4, 8
// (row r of the compact table holds k2 = 2^r)
FC_HD constexpr int tw1_rows(bool recomp) { return recomp ? 4 : 16; }
template <typename S = float> inline size_t smem_bytes(int groups, int klen, bool recomp = false) {
  return (size_t)(2048 + tw1_rows(recomp) * 128 + 128) * sizeof(CPairT<S>) + (size_t)groups * group_layout<S>(klen).total;
}

// One lane of the description of the next iteration.  The group's four warp
// leaders each own one lane (`slot` holds that leader's private copy of the work
// iterator and of the quad's lanes, so the leaders never wait for one another);
// lane 0 also writes the fields shared by the quad.
template <typename S>
FC_HD void fill_packet_lane(const OaParamsT<S> &p, SchedSlotT<S> &slot, PacketT<S> *pk, int lane) {
  constexpr int NL = OaParamsT<S>::NL;
  IterDesc d;
  if (!slot.it.next(p, slot.ln, d)) {
    if (lane == 0) pk->valid = 0;
    return;
  }
  if (lane >= NL) { // no such stream in a two-lane quad: never blocks the fast path
    pk->in_fast[lane] = pk->out_fast[lane] = 1;
    pk->in_shift[lane] = pk->out_shift[lane] = pk->in_bytes[lane] = 0;
    pk->in_src[lane] = p.in;
    pk->out_dst[lane] = p.out;
    pk->ln[lane] = slot.ln[lane];
    return;
  }
  const LaneT<S> ln = slot.ln[lane];
  const bool on = ln.active && d.u < ln.len;
  const long long pos = ln.a0 + d.u;
  // input: a full block inside the row whose 16-byte aligned span stays inside the tensor
  const S *tensor_end = p.in + (p.batch - 1) * p.in_stride + p.n;
  const bool full = on && pos >= 0 && p.n - pos >= p.step;
  const S *a = ln.in + (full ? pos : 0);
  const int ish = align_shift(a);
  const int bytes = ((ish + p.step) * (int)sizeof(S) + 15) & ~15;
  pk->in_shift[lane] = ish;
  pk->in_src[lane] = a - ish;
  pk->in_bytes[lane] = bytes;
  pk->in_fast[lane] = full && (a - ish + bytes / (int)sizeof(S) <= tensor_end);
  // output: the block owns exactly y[lo .. step) with lo = 0 or K - 1, all of it inside the row
  int lo, hi;
  owned_range(p, ln, d.u, d.skip, lo, hi);
  const long long o0 = pos - p.pad;
  const bool whole = on && hi == p.step && o0 + lo >= 0 && o0 + p.step <= p.out_len;
  S *dst = ln.out + (whole ? o0 : 0);
  pk->out_dst[lane] = dst;
  pk->out_shift[lane] = align_shift(dst);
  pk->out_fast[lane] = whole;
  pk->ln[lane] = ln;
  if (lane == 0) {
    pk->valid = 1;
    pk->skip = d.skip;
    pk->first = d.first;
    pk->u = d.u;
  }
}
FC_HD bool all4(const int (&f)[4]) { return (f[0] & f[1] & f[2] & f[3]) != 0; }

// How a staged output lane [0, step) splits into a 16-byte aligned bulk part and
// up to 3 + 3 boundary samples.
struct FlushSplit { int head, nb; };
FC_HD FlushSplit flush_split(int shift, int step) {
  FlushSplit f;
  f.head = (4 - shift) & 3;
  f.nb = (step - f.head) & ~3;
  return f;
}

} // namespace n2048
} // namespace fc
