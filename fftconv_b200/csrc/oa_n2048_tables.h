// oa_n2048_tables.h -- host-side construction of the N = 2048 kernel's twiddle
// tables and launch geometry.  Angles are evaluated in long double and rounded
// once to float (no fast-math sincos anywhere on the path).
#pragma once

#include "oa_n2048_group.cuh"

#include <cmath>
#include <vector>

namespace fc {
namespace n2048 {

template <typename S> inline CPairT<S> unit_root(long long num, long long den) { // exp(-2 pi i num / den)
  num %= den;
  const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)num / (long double)den;
  CPairT<S> w;
  w.re = (S)cosl(a);
  w.im = (S)(-sinl(a));
  return w;
}
inline cpair unit_root_f(long long num, long long den) { return unit_root<float>(num, den); }

template <typename S> inline void fill_tw1(CPairT<S> *tw1) { // [k2][t] = W_2048^{t k2}
  for (int k2 = 0; k2 < 16; ++k2)
    for (int t = 0; t < 128; ++t) tw1[k2 * 128 + t] = unit_root<S>((long long)t * k2, 2048);
}
template <typename S> inline void fill_tw2(CPairT<S> *tw2) { // [k1a][t0] = W_128^{t0 k1a}
  for (int k1a = 0; k1a < 16; ++k1a)
    for (int t0 = 0; t0 < 8; ++t0) tw2[k1a * 8 + t0] = unit_root<S>((long long)t0 * k1a, 128);
}

// Host reference for the spectrum table (tests only; the library computes it on
// the device): hs[s * 128 + c0] = sum_j taps[j] W_2048^{j k} / 2048, k = p2_freq(c0, s).
template <typename S> inline void fill_hs_host(const S *taps, int klen, CPairT<S> *hs) {
  for (int s = 0; s < 16; ++s)
    for (int c0 = 0; c0 < 128; ++c0) {
      const int k = p2_freq(c0, s);
      long double re = 0, im = 0;
      for (int j = 0; j < klen; ++j) {
        const long long e = ((long long)j * k) % 2048;
        const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)e / 2048.0L;
        re += (long double)taps[j] * cosl(a);
        im -= (long double)taps[j] * sinl(a);
      }
      hs[s * 128 + c0].re = (S)(re / 2048.0L);
      hs[s * 128 + c0].im = (S)(im / 2048.0L);
    }
}

// ---- launch geometry (host) --------------------------------------------------------------------
// Streams: NL consecutive streams form a quad.  With at least NL rows a stream is a row; with fewer,
// every row is cut into `vrows` virtual rows so that the lanes of a quad are not left empty (one
// 2^31-sample signal becomes four streams of 2^29 samples).
// Slices: every group of the grid gets `budget` or `budget` + 1 blocks.  A slice that starts in
// mid-stream begins K-1 samples early (its first K-1 outputs belong to the slice before), so slice
// starts do not lie on a regular grid: plan_slices walks the same recurrence as WorkIter and records
// them.  The groups with the extra block are spread evenly over the grid (at most one per SM while
// they are a quarter of all groups or fewer): such a group runs its last block alone on its SM, at a
// fraction of the time a block takes while four groups share the SM -- a 9th block for an eighth of
// the groups costs far less than a 9th block for all (config 2 sharded over 8 GPUs: 8.13 blocks each).
struct SlicePlan {
  long long vrows = 1, vlen = 0, n_streams = 0, n_quads = 0;
  long long budget = 0, extra = 0; // `extra` of the slots have budget + 1 blocks
  std::vector<SlotStart> starts;   // n_slots entries
};
// does slot g carry one of the `extra` extra blocks?  (evenly spaced: Bresenham)
inline bool slot_has_extra(long long g, long long extra, long long n_slots) {
  return ((g + 1) * extra) / n_slots != (g * extra) / n_slots;
}

template <typename S>
inline void stream_geometry(long long batch, long long n, int klen, long long &vrows, long long &vlen) {
  constexpr int NL = OaParamsT<S>::NL;
  constexpr long long A = 16 / (long long)sizeof(S);
  vrows = 1;
  vlen = n;
  if (batch < NL) {
    long long v = (NL + batch - 1) / batch;
    // a virtual row must hold the K-1 samples the next one starts with, and be worth a few blocks
    while (v > 1 && n / v < 4LL * kN) --v;
    if (v > 1) {
      vlen = ((n + v - 1) / v + A - 1) / A * A;
      vrows = (n + vlen - 1) / vlen;
    }
  }
  (void)klen;
}

namespace detail {
// blocks a walk needs from local position u on, in a quad of length qlen
inline long long blocks_to_end(long long qlen, long long u, int step) { return (qlen - u + step - 1) / step; }
} // namespace detail

template <typename S>
inline bool simulate_slices(const OaParamsT<S> &p, long long budget, long long extra, long long n_slots,
                            std::vector<SlotStart> *starts) {
  constexpr int NL = OaParamsT<S>::NL;
  long long q = 0, ua = 0;
  if (starts) starts->resize((size_t)n_slots);
  // quad geometry without touching memory: lanes sid = q NL + l, row = sid / vrows, v = sid % vrows
  auto quad_info = [&](long long quad, long long &qlen, bool &has_prev) {
    qlen = 0;
    has_prev = false;
    for (int l = 0; l < NL; ++l) {
      const long long sid = quad * NL + l;
      if (sid >= p.n_streams) break;
      const long long v = sid % p.vrows, a0 = v * p.vlen;
      long long len = p.n - a0;
      if (len > p.vlen) len = p.vlen;
      if (len <= 0) continue;
      if (len > qlen) qlen = len;
      if (a0 > 0) has_prev = true;
    }
  };
  for (long long g = 0; g < n_slots; ++g) {
    long long left = budget + (slot_has_extra(g, extra, n_slots) ? 1 : 0);
    if (starts) (*starts)[(size_t)g] = SlotStart{q, ua, left};
    while (left > 0 && q < p.n_quads) {
      long long qlen;
      bool has_prev;
      quad_info(q, qlen, has_prev);
      long long u = (ua > 0 || has_prev) ? ua - (p.klen - 1) : 0;
      const long long need = qlen > u ? detail::blocks_to_end(qlen, u, p.step) : 1; // (an empty quad still costs its one block)
      const long long take = need < left ? need : left;
      u += take * p.step;
      left -= take;
      if (u >= qlen) {
        ++q;
        ua = 0;
      } else {
        ua = u;
      }
    }
  }
  return q >= p.n_quads;
}

// mode: 0 = Full, 1 = Same.  n_slots: groups in the grid.  p.slots is left for the caller (device copy).
template <typename S>
inline SlicePlan plan_slices(OaParamsT<S> &p, const S *in, S *out, long long batch, long long n,
                             long long in_stride, long long out_stride, long long out_len, int klen,
                             int mode, long long n_slots) {
  constexpr int NL = OaParamsT<S>::NL;
  p.in = in; p.out = out;
  p.batch = batch; p.n = n;
  p.in_stride = in_stride; p.out_stride = out_stride; p.out_len = out_len;
  p.klen = klen;
  p.step = kN - (klen - 1);
  p.pad = mode ? klen / 2 : 0;
  stream_geometry<S>(batch, n, klen, p.vrows, p.vlen);
  p.n_streams = batch * p.vrows;
  p.n_quads = (p.n_streams + NL - 1) / NL;
  if (n_slots < 1) n_slots = 1;
  p.n_slots = n_slots;
  // lower bound: every stream's samples once; each slice start costs at most K-1 more
  const long long blocks_min = p.n_quads * ((p.vlen + p.step - 1) / p.step);
  long long budget = blocks_min / n_slots, extra = 0;
  if (budget < 1) budget = 1;
  for (bool found = false; !found; ++budget) {
    for (int sixteenth = 0; sixteenth <= 16 && !found; ++sixteenth) {
      extra = n_slots * sixteenth / 16;
      found = simulate_slices(p, budget, extra, n_slots, nullptr);
    }
    if (found) break;
  }
  SlicePlan plan;
  plan.vrows = p.vrows; plan.vlen = p.vlen; plan.n_streams = p.n_streams; plan.n_quads = p.n_quads;
  plan.budget = budget;
  plan.extra = extra;
  simulate_slices(p, budget, extra, n_slots, &plan.starts);
  p.slots = nullptr;
  return plan;
}

} // namespace n2048
} // namespace fc
