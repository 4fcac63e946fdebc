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

struct OaParams {
  const float *in;
  float *out;
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
  long long n_quads;    // ceil(n_streams / 4)
  long long total_work; // n_quads * segs_per_chunk
  const cpair *tw1, *tw2, *hs; // global tables (see oa_n2048_core.cuh)
};

struct Lane {
  const float *in;
  float *out;
  long long seg0;
  long long cnt;
  bool active;
};

FC_HD void make_lanes(const OaParams &p, long long quad, Lane (&ln)[4]) {
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const long long sid = quad * 4 + l;
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
FC_HD bool needs_warmup(const Lane (&ln)[4], long long j) {
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
FC_HD void io_load(int t, Regs &v, const OaParams &p, const Lane (&ln)[4], long long j) {
  float x[4][kPts];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const long long seg = ln[l].seg0 + j;
    const bool on = ln[l].active && seg >= 0 && j < ln[l].cnt;
    const long long pos = seg * p.step;
    long long len = on ? p.n - pos : 0;
    if (len > p.step) len = p.step;
    const float *src = ln[l].in + (on ? pos : 0);
    const int ilen = (int)len;
#pragma unroll
    for (int m = 0; m < kPts; ++m) {
      const int i = t + 128 * m;
#if defined(__CUDA_ARCH__)
      x[l][m] = (i < ilen) ? __ldg(src + i) : 0.0f;
#else
      x[l][m] = (i < ilen) ? src[i] : 0.0f;
#endif
    }
  }
#pragma unroll
  for (int m = 0; m < kPts; ++m) {
    v.re[m] = make_f2(x[0][m], x[2][m]);
    v.im[m] = make_f2(x[1][m], x[3][m]);
  }
}

// v holds y[t + 128 m] in slot slot16(m).  Adds the incoming tail, saves the
// outgoing tail, stores the samples this segment owns.
FC_HD void io_store(int t, const Regs &v, const OaParams &p, const Lane (&ln)[4], long long j,
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
  for (int m = 0; m < kPts; ++m) {
    const int i = t + 128 * m;
    const int s = slot16(m);
    f2 yr = v.re[s], yi = v.im[s];
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
      const long long o = obase[l] + i;
      if (i < limit[l] && o >= 0) ln[l].out[o] = y[l];
    }
  }
}

// ---- fast path: staged through shared memory ------------------------------------
// staging buffers hold 4 lanes of `ls` floats each, ls = (3 + step + 3) & ~3 (oa_n2048_kernel.cuh)

FC_HD int align_shift(const void *ptr) { return (int)((reinterpret_cast<unsigned long long>(ptr) >> 2) & 3ULL); }

// Per-thread constants of the fast path (depend on t, K only).
struct ThreadMasks {
  unsigned in_zero;   // bit m: t + 128 m >= step  (padding region: value is 0, no load)
  unsigned tail_add;  // bit m: t + 128 m <  K - 1 (incoming tail is added)
};
FC_HD ThreadMasks make_masks(int t, const OaParams &p) {
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
FC_HD void fast_load(int t, Regs &v, const float *in, int ls, const int (&shift)[4], unsigned in_zero) {
  const float *b0 = in + 0 * ls + shift[0] + t;
  const float *b1 = in + 1 * ls + shift[1] + t;
  const float *b2 = in + 2 * ls + shift[2] + t;
  const float *b3 = in + 3 * ls + shift[3] + t;
#pragma unroll
  for (int m = 0; m < kPts; ++m) {
    // step >= 2048 - 409: only m >= 12 can fall into the padding region
    const bool z = (m >= 12) && ((in_zero >> m) & 1u);
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
    if (!z) {
      x0 = b0[128 * m];
      x1 = b1[128 * m];
      x2 = b2[128 * m];
      x3 = b3[128 * m];
    }
    v.re[m] = make_f2(x0, x2);
    v.im[m] = make_f2(x1, x3);
  }
}

// registers -> the four lanes' outputs (+ tail hand-over).  b_l points at y[t] of
// lane l -- in a shared staging buffer or directly in global memory (one coalesced
// 128-byte line per warp instruction); element y[t + 128 m] goes to b_l[128 m].
FC_HD void fast_store_to(int t, const Regs &v, float *b0, float *b1, float *b2, float *b3, const ThreadMasks &k,
                         const pt4 *tail_prev, pt4 *tail_next, bool add_tail, bool stage, int step);

// registers -> staged output (+ tail hand-over).  `out` is the output staging buffer.
FC_HD void fast_store(int t, const Regs &v, float *out, int ls, const int (&shift)[4], const ThreadMasks &k,
                      const pt4 *tail_prev, pt4 *tail_next, bool add_tail, bool stage, int step) {
  fast_store_to(t, v, out + 0 * ls + shift[0] + t, out + 1 * ls + shift[1] + t, out + 2 * ls + shift[2] + t,
                out + 3 * ls + shift[3] + t, k, tail_prev, tail_next, add_tail, stage, step);
}

FC_HD void fast_store_to(int t, const Regs &v, float *b0, float *b1, float *b2, float *b3, const ThreadMasks &k,
                         const pt4 *tail_prev, pt4 *tail_next, bool add_tail, bool stage, int step) {
#pragma unroll
  for (int m = 0; m < kPts; ++m) {
    const int i = t + 128 * m;
    const int s = slot16(m);
    f2 yr = v.re[s], yi = v.im[s];
    if (m < 4) { // K - 1 <= 409 < 512
      if (add_tail && ((k.tail_add >> m) & 1u)) {
        f2 tr, ti;
        get(tail_prev, i, tr, ti);
        yr = add2(yr, tr);
        yi = add2(yi, ti);
      }
    }
    const bool in_tail = (m >= 12) && ((k.in_zero >> m) & 1u); // i >= step
    if (in_tail) {
      put(tail_next, i - step, yr, yi);
    } else if (stage) {
      b0[128 * m] = yr.x;
      b1[128 * m] = yi.x;
      b2[128 * m] = yr.y;
      b3[128 * m] = yi.y;
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
  FC_HD bool next(const OaParams &p, Lane (&ln)[4], IterDesc &d) {
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

struct SchedSlot {      // private scheduler state of one warp leader
  WorkIter it;
  Lane ln[4];
  long long pad_;
};

// tail buffers: two (ping-pong) arrays of tail_pts(K) points per group
FC_HD int tail_pts(int klen) { return ((klen - 1) + 7) & ~7; }

struct Packet {        // one iteration of one group, produced by thread 0
  int valid;           // 0: the slice is exhausted
  int warm;
  int first;           // no incoming tail
  int in_fast[4];      // per lane; the quad takes the fast path when all four are set
  int out_fast[4];
  int in_shift[4];
  int out_shift[4];
  int in_bytes[4];     // bulk-load size per lane (fast path)
  const float *in_src[4];
  float *out_dst[4];
  Lane ln[4];          // slow path only
  long long j;
};

struct GroupLayout {   // byte offsets inside a group's shared-memory block
  int lane_stride;     // floats per lane of the staging area (which aliases xb)
  int xb, tails, packets, sched, mbar, total;
};
// The input and output staging areas alias the exchange buffer: 4 lanes x
// lane_stride floats <= 4 * 1832 * 4 B = 29312 B <= 32768 B.
FC_HD GroupLayout group_layout(int klen) {
  GroupLayout g;
  const int step = kN - (klen - 1);
  g.lane_stride = (3 + step + 3) & ~3;
  int off = 0;
  g.xb = off;      off += kN * (int)sizeof(pt4);
  g.tails = off;   off += 2 * tail_pts(klen) * (int)sizeof(pt4);
  g.packets = off; off += 3 * (((int)sizeof(Packet) + 15) & ~15);
  g.sched = off;   off += 4 * (int)sizeof(SchedSlot);
  g.mbar = off;    off += 16;
  g.total = off;
  return g;
}
inline size_t smem_bytes(int groups, int klen) {
  return (size_t)(2048 + 2048 + 128) * sizeof(cpair) + (size_t)groups * group_layout(klen).total;
}

// One lane of the description of the next iteration.  The group's four warp
// leaders each own one lane (`slot` holds that leader's private copy of the work
// iterator and of the quad's lanes, so the leaders never wait for one another);
// lane 0 also writes the fields shared by the quad.
FC_HD void fill_packet_lane(const OaParams &p, SchedSlot &slot, Packet *pk, int lane) {
  IterDesc d;
  if (!slot.it.next(p, slot.ln, d)) {
    if (lane == 0) pk->valid = 0;
    return;
  }
  const Lane ln = slot.ln[lane];
  const long long seg = ln.seg0 + d.j;
  const bool on = ln.active && seg >= 0 && d.j < ln.cnt;
  const long long pos = seg * p.step;
  // input: a full segment whose 16-byte aligned span stays inside the tensor
  const float *tensor_end = p.in + (p.batch - 1) * p.in_stride + p.n;
  const bool full = on && (p.n - pos >= p.step);
  const float *a = ln.in + (full ? pos : 0);
  const int ish = align_shift(a);
  const int bytes = ((ish + p.step) * 4 + 15) & ~15;
  pk->in_shift[lane] = ish;
  pk->in_src[lane] = a - ish;
  pk->in_bytes[lane] = bytes;
  pk->in_fast[lane] = full && (a - ish + bytes / 4 <= tensor_end);
  // output: an interior segment owns exactly y[0..step), all of it inside the row
  const long long o0 = pos - p.pad;
  const bool interior = on && seg < p.segs - 1 && o0 >= 0;
  float *dst = ln.out + (interior ? o0 : 0);
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
