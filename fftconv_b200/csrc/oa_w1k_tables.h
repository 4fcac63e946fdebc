// oa_w1k_tables.h -- host-side tables of the warp-per-FFT engine (oa_w1k_core.cuh): per-lane
// twiddles, evaluated in long double and rounded once, and (tests only) the spectrum table.
#pragma once

#include "oa_n2048_tables.h"
#include "oa_w1k_core.cuh"

namespace fc {
namespace w1k {

constexpr int kTwElems = kTwRows * 32;
// row r of lane l sits at tw_index(r, l) (rows in pairs, 16 bytes per lane and pair).  r = 0..3: W_512^{l 2^r};
// r = 4 + i: exp(-2 pi j i / 32) W_1024^l = W_1024^{32 i + l}; r = 19 + j, j = 1..15: W_512^{l j} (every power, for the
// kernel variant that does not recompute them); r = 35: padding
inline void fill_tw(cpair *tw) {
  for (int l = 0; l < 32; ++l) {
    for (int r = 0; r < 4; ++r) tw[tw_index(r, l)] = n2048::unit_root<float>((long long)l << r, 512);
    for (int i = 0; i < 16; ++i) tw[tw_index(kTwRowCross + i, l)] = n2048::unit_root<float>(32 * i + l, 1024);
    for (int j = 1; j < 16; ++j) tw[tw_index(kTwRowPowers + j, l)] = n2048::unit_root<float>((long long)l * j, 512);
    tw[tw_index(35, l)] = cpair{1.0f, 0.0f};
  }
}

// Host reference for the spectrum (tests only; the library computes it on the device with
// spectrum_kernel, layout LAYOUT_W1K): entry e holds H[freq_of_index(e)] / 1024 as (re, im).
inline void fill_hs_host(const float *taps, int klen, cpair *hs) {
  for (unsigned e = 0; e < 1024; ++e) {
    const unsigned k = freq_of_index(e);
    long double re = 0, im = 0;
    for (int j = 0; j < klen; ++j) {
      const long long m = ((long long)j * k) % 1024;
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / 1024.0L;
      re += (long double)taps[j] * cosl(a);
      im -= (long double)taps[j] * sinl(a);
    }
    hs[e].re = (float)(re / 1024.0L);
    hs[e].im = (float)(im / 1024.0L);
  }
}
// the kernel's shared-memory form: the folded tables of multiply_spectrum_fused
inline void pack_hs(const cpair *hs, Quad *q0, Quad *q1) {
  for (int q = 0; q < 512; ++q)
    fold_spectrum_entry(n2048::unit_root<float>(32 * (q >> 5), 1024), hs[2 * q], hs[2 * q + 1], q0[q], q1[q]);
}

} // namespace w1k
} // namespace fc
