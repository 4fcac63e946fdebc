// oa_n2048_group.cuh -- work decomposition and I/O phases around the N = 2048
// segment engine (oa_n2048_core.cuh).  Host+device, so tests can emulate it.
//
// Reference semantics being replaced: FFTConvEngine::oaconvolve
// (/root/reference/include/fftconv/fftconv.hpp:382-437): zero `out`, then for
// every block pos = 0, step, 2 step, ... add y = ifft(fft(pad(a[pos:pos+len])) H)/N
// into out at pos (Full) or pos - K/2 (Same), clipped to out.size().
//
// Here nothing is zeroed and nothing is accumulated in global memory: a "stream"
// (one row, or one chunk of consecutive segments of a row) is walked in order by
// one lane of a 128-thread group; the K-1 sample tail of segment s stays in
// shared memory and is added to the head of segment s+1 before that sample is
// stored, so every output sample is written exactly once, with at most two
// addends -> deterministic, no atomics.  A stream that starts in the middle of
// a row first runs its predecessor segment with stores disabled ("warm-up") to
// obtain the incoming tail.
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
  long long segs;       // segments per row = ceil(n / step)
  long long chunks;     // streams per row
  long long segs_per_chunk;
  long long n_streams;  // batch * chunks
  long long n_quads;    // ceil(n_streams / NL)
  long long total_work; // n_quads * segs_per_chunk
  const CPairT<S> *tw1, *tw2, *hs; // global tables (see oa_n2048_core.cuh)
};
using OaParams = OaParamsT<float>;

template <typename S> struct LaneT {
  const S *in;
  S *out;
  long long seg0;
  long long cnt;
  bool active;
};
using Lane = LaneT<float>;

template <typename S> FC_HD void make_lanes(const OaParamsT<S> &p, long long quad, LaneT<S> (&ln)[4]) {
  constexpr int NL = OaParamsT<S>::NL;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const long long sid = quad * NL + l;
    if (l >= NL) { // unused slots of a two-lane (float64) quad
      ln[l].active = false;
      ln[l].seg0 = ln[l].cnt = 0;
      ln[l].in = p.in;
      ln[l].out = p.out;
      continue;
    }
    ln[l].active = sid < p.n_streams;
    const long long row = ln[l].active ? sid / p.chunks : 0;
    const long long chunk = ln[l].active ? sid - row * p.chunks : 0;
    ln[l].seg0 = chunk * p.segs_per_chunk;
    const long long rem = p.segs - ln[l].seg0;
    ln[l].cnt = rem < p.segs_per_chunk ? rem : p.segs_per_chunk;
    ln[l].in = p.in + row * p.in_stride;
    ln[l].out = p.out + row * p.out_stride;
  }
}

// does any lane of this quad have a predecessor segment at chunk-local index j?
template <typename S> FC_HD bool needs_warmup(const LaneT<S> (&ln)[4], long long j) {
  bool w = false;
#pragma unroll
  for (int l = 0; l < 4; ++l) w = w || (ln[l].active && j < ln[l].cnt && ln[l].seg0 + j > 0);
  return w;
}

// ---- slow path: direct global access -------------------------------------------
// Load x[t + 128 m] of the four lanes' segments (chunk-local index j, may be -1
// for a warm-up) into the packed registers, zero padded.
//   lane 0 -> re, half A    lane 1 -> im, half A
//   lane 2 -> re, half B    lane 3 -> im, half B
template <typename D, typename S>
FC_HD void io_load(int t, RegsT<D> &v, const OaParamsT<S> &p, const LaneT<S> (&ln)[4], long long j) {
  constexpr int NL = DataTraits<D>::NL;
  S x[NL][kPts];
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    const long long seg = ln[l].seg0 + j;
    const bool on = ln[l].active && seg >= 0 && j < ln[l].cnt;
    const long long pos = seg * p.step;
    long long len = on ? p.n - pos : 0;
    if (len > p.step) len = p.step;
    const S *src = ln[l].in + (on ? pos : 0);
    const int ilen = (int)len;
#pragma unroll
    for (int m = 0; m < kPts; ++m) {
      const int i = t + 128 * m;
#if defined(__CUDA_ARCH__)
      x[l][m] = (i < ilen) ? __ldg(src + i) : (S)0;
#else
      x[l][m] = (i < ilen) ? src[i] : (S)0;
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

// v holds y[t + 128 m] in slot slot16(m).  Adds the incoming tail, saves the
// outgoing tail, stores the samples this segment owns.
template <typename D, typename S>
FC_HD void io_store(int t, const RegsT<D> &v, const OaParamsT<S> &p, const LaneT<S> (&ln)[4], long long j,
                    const PtT<D> *tail_prev, PtT<D> *tail_next, bool add_tail, bool warm) {
  constexpr int NL = DataTraits<D>::NL;
  int limit[NL];
  long long obase[NL]; // output index of y[0]
#pragma unroll
  for (int l = 0; l < NL; ++l) {
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
      if (i < limit[l] && o >= 0) ln[l].out[o] = y[l];
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
                         PtT<D> *tail_next, bool add_tail, bool stage, int step) {
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
    if (in_tail) {
      put(tail_next, i - step, yr, yi);
    } else if (stage) {
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
  long long j;    // chunk-local segment index (jfirst - 1 for a warm-up)
  bool warm;
  bool first;     // first iteration of a quad visit: no incoming tail
};

// Enumerates the iterations of the work slice [w0, w1): for every quad visit an
// optional warm-up iteration, then the visit's segments in order.
struct WorkIter {
  long long w, w1;
  long long quad, j, jend, jfirst;
  bool in_visit;

  FC_HD void init(long long w0_, long long w1_) {
    w = w0_;
    w1 = w1_;
    in_visit = false;
    quad = j = jend = jfirst = 0;
  }
  // returns false when the slice is exhausted; `ln` is refreshed on a new quad
  template <typename S> FC_HD bool next(const OaParamsT<S> &p, LaneT<S> (&ln)[4], IterDesc &d) {
    if (!in_visit) {
      if (w >= w1) return false;
      quad = w / p.segs_per_chunk;
      jfirst = w - quad * p.segs_per_chunk;
      jend = jfirst + (w1 - w);
      if (jend > p.segs_per_chunk) jend = p.segs_per_chunk;
      make_lanes(p, quad, ln);
      w += jend - jfirst;
      j = needs_warmup(ln, jfirst) ? jfirst - 1 : jfirst;
      in_visit = true;
      d.first = true;
    } else {
      d.first = false;
    }
    d.quad = quad;
    d.j = j;
    d.warm = j < jfirst;
    ++j;
    if (j >= jend) in_visit = false;
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
  int warm;
  int first;           // no incoming tail
  int in_fast[4];      // per lane; the quad takes the fast path when all lanes are set
  int out_fast[4];     //   (slots >= NL of a two-lane quad are preset to 1)
  int in_shift[4];
  int out_shift[4];
  int in_bytes[4];     // bulk-load size per lane (fast path)
  const S *in_src[4];
  S *out_dst[4];
  LaneT<S> ln[4];      // slow path only
  long long j;
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
  constexpr int NL = OaParamsT<S>::NL, A = 16 / (int)sizeof(S);
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
  const long long seg = ln.seg0 + d.j;
  const bool on = ln.active && seg >= 0 && d.j < ln.cnt;
  const long long pos = seg * p.step;
  // input: a full segment whose 16-byte aligned span stays inside the tensor
  const S *tensor_end = p.in + (p.batch - 1) * p.in_stride + p.n;
  const bool full = on && (p.n - pos >= p.step);
  const S *a = ln.in + (full ? pos : 0);
  const int ish = align_shift(a);
  const int bytes = ((ish + p.step) * (int)sizeof(S) + 15) & ~15;
  pk->in_shift[lane] = ish;
  pk->in_src[lane] = a - ish;
  pk->in_bytes[lane] = bytes;
  pk->in_fast[lane] = full && (a - ish + bytes / (int)sizeof(S) <= tensor_end);
  (void)A;
  // output: an interior segment owns exactly y[0..step), all of it inside the row
  const long long o0 = pos - p.pad;
  const bool interior = on && seg < p.segs - 1 && o0 >= 0;
  S *dst = ln.out + (interior ? o0 : 0);
  pk->out_dst[lane] = dst;
  pk->out_shift[lane] = align_shift(dst);
  pk->out_fast[lane] = interior && !d.warm;
  pk->ln[lane] = ln;
  if (lane == 0) {
    pk->valid = 1;
    pk->warm = d.warm;
    pk->first = d.first;
    pk->j = d.j;
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
