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
#pragma once

#include "oa_n2048_core.cuh"

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

// Uniform (per group) control flow.  Exec::iteration(lanes, j, warm, add_tail,
// parity) runs load -> P0..P4 -> store for all 128 threads of the group.
template <class Exec>
FC_HD void run_group(const OaParams &p, long long w0, long long w1, Exec &ex) {
  int parity = 0;
  long long w = w0;
  while (w < w1) {
    const long long quad = w / p.segs_per_chunk;
    long long j = w - quad * p.segs_per_chunk;
    long long jend = j + (w1 - w);
    if (jend > p.segs_per_chunk) jend = p.segs_per_chunk;
    Lane ln[4];
    make_lanes(p, quad, ln);
    // one call site (keeps a single copy of the ~1600-instruction body in the
    // instruction cache): a warm-up is simply iteration j - 1 with stores off.
    const long long jfirst = j;
    long long jj = needs_warmup(ln, j) ? j - 1 : j;
    bool add_tail = false;
    w += jend - j;
    for (; jj < jend; ++jj) {
      ex.iteration(ln, jj, /*warm=*/jj < jfirst, add_tail, parity);
      parity ^= 1;
      add_tail = true;
    }
  }
}

// tail buffers: two (ping-pong) arrays of tail_pts(K) points per group
FC_HD int tail_pts(int klen) { return ((klen - 1) + 7) & ~7; }

} // namespace n2048
} // namespace fc
