// oa_n2048_tables.h -- host-side construction of the N = 2048 kernel's twiddle
// tables and launch geometry.  Angles are evaluated in long double and rounded
// once to float (no fast-math sincos anywhere on the path).
#pragma once

#include "oa_n2048_group.cuh"

#include <cmath>

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

// mode: 0 = Full, 1 = Same.  chunks: streams per row (>= 1).
template <typename S>
inline void fill_params(OaParamsT<S> &p, const S *in, S *out, long long batch, long long n,
                        long long in_stride, long long out_stride, long long out_len, int klen,
                        int mode, long long chunks) {
  constexpr int NL = OaParamsT<S>::NL;
  p.in = in; p.out = out;
  p.batch = batch; p.n = n;
  p.in_stride = in_stride; p.out_stride = out_stride; p.out_len = out_len;
  p.klen = klen;
  p.step = kN - (klen - 1);
  p.pad = mode ? klen / 2 : 0;
  p.segs = (n + p.step - 1) / p.step;
  if (chunks < 1) chunks = 1;
  if (chunks > p.segs) chunks = p.segs;
  p.segs_per_chunk = (p.segs + chunks - 1) / chunks;
  p.chunks = (p.segs + p.segs_per_chunk - 1) / p.segs_per_chunk;
  p.n_streams = batch * p.chunks;
  p.n_quads = (p.n_streams + NL - 1) / NL;
  p.total_work = p.n_quads * p.segs_per_chunk;
}

} // namespace n2048
} // namespace fc
