// hilbert_r2c_tables.h -- host-side tables of the real-input Hilbert engine (hilbert_r2c_core.cuh),
// evaluated in long double and rounded once.
#pragma once

#include "hilbert_r2c_core.cuh"

#include <cmath>

namespace fc {
namespace hr2c {

inline void fill_tables(Quad *tw1, Quad *tw2, Pair *mir) {
  const long double two_pi = 2.0L * 3.14159265358979323846264338327950288L;
  auto root = [&](long long num, long long den, float &re, float &im) { // exp(-2 pi j num / den)
    const long double a = two_pi * (long double)(num % den) / (long double)den;
    re = (float)cosl(a);
    im = (float)(-sinl(a));
  };
  for (int t = 0; t < kThreads; ++t)
    for (int p = 0; p < 4; ++p) {
      const long long r = 16 * (t >> 3) + 2 * (t & 7);
      Quad q;
      root(r << p, 4096, q.a, q.c);
      root((r + 1) << p, 4096, q.b, q.d);
      tw1[p * kThreads + t] = q;
    }
  for (int j = 0; j < 8; ++j)
    for (int p = 0; p < 4; ++p) {
      Quad q;
      root((long long)(2 * j) << p, 256, q.a, q.c);
      root((long long)(2 * j + 1) << p, 256, q.b, q.d);
      tw2[p * 8 + j] = q;
    }
  auto cs = [&](long long k) {
    Pair v;
    if (k % kM == 0) return Pair{0.0f, 0.0f}; // G[0] = 0
    const long double a = two_pi * (long double)k / (long double)kRow;
    v.x = (float)(cosl(a) / (long double)kM);
    v.y = (float)(sinl(a) / (long double)kM);
    return v;
  };
  for (int s = 0; s < kThreads; ++s) {
    const P3Map m = p3_map(s);
    for (int k2 = 0; k2 < 16; ++k2) mir[k2 * kMirStride + s] = cs(m.k0a + 16 * m.k1a + 256 * k2);
  }
  const P3Map self = p3_map(kSelfPaired);
  for (int k2 = 0; k2 < 16; ++k2) mir[k2 * kMirStride + kThreads] = cs(self.k0b + 16 * self.k1b + 256 * k2);
}

} // namespace hr2c
} // namespace fc
