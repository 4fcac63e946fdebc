// oa_w2_tables.h -- host-side construction of the warp-pair engine's twiddle table and of a
// reference spectrum table (tests).  Angles in long double, rounded once to float.
#pragma once

#include "oa_n2048_tables.h"
#include "oa_w2_group.cuh"

namespace fc {
namespace w2 {

// the five chain tables of oa_w2_core.cuh: T[j * 32 + lane] = g_j (j = 0..4), T[160 + lane] = b
inline void fill_tw(cpair *tw) {
  auto root = [](long long num, long long den) { return n2048::unit_root_f(((num % den) + den) % den, den); };
  for (int id = 0; id < kTwTables; ++id) {
    cpair *T = tw + id * kTwPerTable;
    for (int lane = 0; lane < 32; ++lane) {
      cpair one;
      one.re = 1.0f;
      one.im = 0.0f;
      cpair b = one;
      for (int j = 0; j < 5; ++j) {
        cpair g = one;
        switch (id) {
        case 0: case 1: g = root((long long)(2 * lane) << j, 2048); break;       // p = 0, trips 1 and 3
        case 2: g = root((long long)(2 * lane) << j, 2048); break;               // p = 1, trip 1
        case 3: g = root((long long)(2 * lane + 1) << j, 2048); break;           // p = 1, trip 3
        case 4: g = j < 4 ? root(1 << j, 64) : root(-16, 64); break;             // p = 1, trip 4
        }
        T[j * 32 + lane] = g;
      }
      if (id == 2) b = root(lane, 2048);
      if (id == 3) b.re = (lane & 1) ? -1.0f : 1.0f;
      if (id == 4) b = root(16, 64);
      T[5 * 32 + lane] = b;
    }
  }
}

// hs[(p * 32 + k1) * 32 + k2] = sum_j taps[j] W_2048^{j k} / 2048, k = w2_freq(p, k1, k2)
inline void fill_hs_host(const float *taps, int klen, cpair *hs) {
  for (int idx = 0; idx < 2048; ++idx) {
    const int k = w2_freq(idx >> 10, (idx >> 5) & 31, idx & 31);
    long double re = 0, im = 0;
    for (int j = 0; j < klen; ++j) {
      const long long e = ((long long)j * k) % 2048;
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)e / 2048.0L;
      re += (long double)taps[j] * cosl(a);
      im -= (long double)taps[j] * sinl(a);
    }
    hs[idx].re = (float)(re / 2048.0L);
    hs[idx].im = (float)(im / 2048.0L);
  }
}

} // namespace w2
} // namespace fc
