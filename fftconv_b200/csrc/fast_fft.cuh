// fast_fft.cuh -- register-blocked shared-memory FFT engine for power-of-two
// lengths 512 .. 16384, float and double.  Host+device (tests emulate it on the CPU).
//
// One complex FFT of length N = 2^LOG2N is owned by N/16 threads, 16 points per
// thread per pass.  The transform is a sequence of in-place decimation-in-frequency
// passes: first a radix-2^R0 pass (R0 = LOG2N mod 4, or 4), then radix-16 passes.
// Every pass reads 16 points into registers, runs 16-point (or smaller) DFTs there,
// multiplies by twiddles and writes the SAME shared-memory addresses back, so one
// barrier per pass suffices and no reordering pass exists: the spectrum table is
// stored in the resulting digit-reversed position order, and the inverse is the
// exact mirror (conjugate twiddles first, then the inverse small DFT).
//
// First pass inputs come straight from global memory into registers and the last
// inverse pass leaves the outputs in registers at the same (thread, slot) places, so
// the signal itself never round-trips through shared memory for I/O.  The last
// forward pass, the spectrum multiply and the first inverse pass are fused in
// registers.
//
// Shared-memory layout: element i at i + (i >> 4) (one pad element per 16), which
// makes every pass bank-conflict free for 8- and 16-byte elements.
#pragma once

#include "oa_n2048_core.cuh" // the fused-multiply-add forms of the 8- and 16-point DFTs (dft8, dft16)
#include "packed_f32x2.cuh"

#if defined(__CUDACC__)
#define FCF_HD __host__ __device__ __forceinline__
#else
#define FCF_HD inline
#endif

namespace fc {
namespace fast {

// A complex number whose parts are of the DATA type D: float, double, or the packed pair f2
// (two independent float transforms per thread, one in each half of every 64-bit register
// pair: SASS FADD2 / FMUL2 / FFMA2).  Twiddles and spectra are always of the SCALAR type
// scalar_t<D> and are broadcast to both halves.
template <typename D> struct cx {
  D x, y;
};
template <typename D> struct scalar_of { using type = D; };
template <> struct scalar_of<f2> { using type = float; };
template <typename D> using scalar_t = typename scalar_of<D>::type;

FCF_HD f2 operator+(f2 a, f2 b) { return add2(a, b); }
FCF_HD f2 operator-(f2 a, f2 b) { return sub2(a, b); }
FCF_HD f2 operator-(f2 a) { return muls2(a, -1.0f); }

template <typename D> FCF_HD cx<D> operator+(cx<D> a, cx<D> b) { return {a.x + b.x, a.y + b.y}; }
template <typename D> FCF_HD cx<D> operator-(cx<D> a, cx<D> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> FCF_HD cx<T> cmul(cx<T> a, cx<T> w) { return {a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x}; }
template <typename T> FCF_HD cx<T> cmulc(cx<T> a, cx<T> w) { return {a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y}; }
// packed data, scalar twiddle: four packed instructions
FCF_HD cx<f2> cmul(cx<f2> a, cx<float> w) {
  return {fmas2(a.x, w.x, muls2(a.y, -w.y)), fmas2(a.x, w.y, muls2(a.y, w.x))};
}
FCF_HD cx<f2> cmulc(cx<f2> a, cx<float> w) {
  return {fmas2(a.x, w.x, muls2(a.y, w.y)), fmas2(a.x, -w.y, muls2(a.y, w.x))};
}
// multiply by -j (S < 0) or +j (S > 0)
template <int S, typename D> FCF_HD cx<D> rot90(cx<D> a) {
  return S < 0 ? cx<D>{a.y, -a.x} : cx<D>{-a.y, a.x};
}
// j g v for a real scalar g
template <typename T> FCF_HD cx<T> mul_j(cx<T> v, T g) { return {-g * v.y, g * v.x}; }
FCF_HD cx<f2> mul_j(cx<f2> v, float g) { return {muls2(v.y, -g), muls2(v.x, g)}; }
// a + (-j b) (S < 0) / a + (+j b) (S > 0), and a - (...): no explicit negation
template <int S, typename D> FCF_HD cx<D> add_rot(cx<D> a, cx<D> b) {
  return S < 0 ? cx<D>{a.x + b.y, a.y - b.x} : cx<D>{a.x - b.y, a.y + b.x};
}
template <int S, typename D> FCF_HD cx<D> sub_rot(cx<D> a, cx<D> b) {
  return S < 0 ? cx<D>{a.x - b.y, a.y + b.x} : cx<D>{a.x + b.y, a.y - b.x};
}

template <int LOG2N> struct Plan {
  static_assert(LOG2N >= 9 && LOG2N <= 14, "fast engine covers N = 512 .. 16384");
  static constexpr int N = 1 << LOG2N;
  static constexpr int R0 = (LOG2N % 4 == 0) ? 4 : (LOG2N % 4); // log2 of the first radix
  static constexpr int P = (LOG2N - R0) / 4 + 1;                // passes
  static constexpr int THREADS = N / 16;                        // per FFT
  static constexpr int SMEM_ELEMS = N + N / 16;                 // padded
  // Per-pass twiddle tables, s-major so that the threads of a warp read consecutive
  // entries: pass 0 holds (R-1) rows of N/R entries, TW0[(s-1) (N/R) + p] = W_N^{p s};
  // middle pass J holds 15 rows of m/16 entries, TWJ[(s-1) (m/16) + p] = W_m^{p s}.
  FCF_HD static constexpr int tw_offset(int pass) { // first entry of pass `pass` (0 .. P-2)
    int off = 0;
    if (pass == 0) return 0;
    off += ((1 << R0) - 1) * (N >> R0);
    for (int j = 1; j < pass; ++j) off += 15 * ((N >> (R0 + 4 * (j - 1))) >> 4);
    return off;
  }
  static constexpr int TW_ELEMS = tw_offset(P - 1);
};

FCF_HD constexpr int phys(int i) { return i + (i >> 4); }

// spectrum-table index (s-major, see last_fused) -> position after the forward passes
FCF_HD constexpr unsigned pos_of_table_index(unsigned idx, int log2n) {
  const unsigned threads = (1u << log2n) / 16;
  return 16 * (idx % threads) + idx / threads;
}

// position (after all forward passes) -> frequency index
FCF_HD constexpr unsigned freq_of_pos(unsigned pos, int log2n) {
  const int r0 = (log2n % 4 == 0) ? 4 : (log2n % 4);
  unsigned k = 0, mult = 1;
  int lm = log2n; // log2 of the current sub-transform length
  int r = r0;
  while (lm > 0) {
    const unsigned stride = 1u << (lm - r);
    const unsigned digit = pos / stride;
    pos -= digit * stride;
    k += mult * digit;
    mult <<= r;
    lm -= r;
    r = 4;
  }
  return k;
}

// ---- small DFTs in registers: natural order in, natural order out -----------------
template <int S, typename D> FCF_HD void dft2(cx<D> *v) {
  const cx<D> a = v[0], b = v[1];
  v[0] = a + b;
  v[1] = a - b;
}
template <int S, typename D> FCF_HD void dft4(cx<D> *v) {
  const cx<D> s0 = v[0] + v[2], d0 = v[0] - v[2], s1 = v[1] + v[3], d1 = v[1] - v[3];
  v[0] = s0 + s1;
  v[1] = add_rot<S>(d0, d1); // d0 - j d1 (forward) / d0 + j d1 (inverse)
  v[2] = s0 - s1;
  v[3] = sub_rot<S>(d0, d1);
}
// 8- and 16-point DFTs: the fused-multiply-add forms of oa_n2048_core.cuh (52 and 144 instructions;
// the constant twiddles ride in the butterflies), wrapped to natural order in, natural order out --
// the permutations are register renamings.
template <int S, typename D> FCF_HD void dft8(cx<D> *v) {
  D re[8], im[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    re[i] = v[i].x;
    im[i] = v[i].y;
  }
  n2048::dft8<S>(re, im);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = cx<D>{re[n2048::slot8(k)], im[n2048::slot8(k)]};
}
template <int S, typename D> FCF_HD void dft16(cx<D> *v) {
  n2048::RegsT<D> r;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    r.re[i] = v[i].x;
    r.im[i] = v[i].y;
  }
  n2048::dft16<S>(r);
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = cx<D>{r.re[n2048::slot16(k)], r.im[n2048::slot16(k)]};
}
template <int S, int LOGR, typename D> FCF_HD void dft_small(cx<D> *v) {
  if (LOGR == 1) dft2<S>(v);
  else if (LOGR == 2) dft4<S>(v);
  else if (LOGR == 3) dft8<S>(v);
  else dft16<S>(v);
}

// ---- twiddles of one thread and pass ---------------------------------------------------
// tw[s] = W^{p s}, s = 1 .. R-1: only the powers of two are loaded (TW[(s - 1) * stride + p]) and
// the rest are formed as products in registers (at most three deep, relative error <= 3 eps):
// the loads come from global memory through L1 and their latency, not arithmetic, is what a
// pass waits for (ncu: DESIGN.md section 4).
template <typename T> FCF_HD cx<T> cxmul(cx<T> a, cx<T> b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <int R, bool RECOMP, typename T> FCF_HD void load_twiddles(cx<T> *tw, const cx<T> *TW, int stride, int p) {
  if (!RECOMP) {
#pragma unroll
    for (int s = 1; s < R; ++s) tw[s] = TW[(s - 1) * stride + p];
  } else {
#pragma unroll
    for (int s = 1; s < R; s <<= 1) tw[s] = TW[(s - 1) * stride + p];
#pragma unroll
    for (int s = 3; s < R; ++s)
      if (s & (s - 1)) { // not a power of two: highest bit times the rest
        int hb = 1;
        while (2 * hb <= s) hb <<= 1;
        tw[s] = cxmul(tw[hb], tw[s - hb]);
      }
  }
}
template <typename D> struct recomp_twiddles { static constexpr bool value = true; };

// ---- passes --------------------------------------------------------------------
// Register slot convention: pass 0 holds butterfly u (of U = 16 >> R0) in slots
// [u R, u R + R), element q = logical index p_u + q (N >> R0), p_u = tid + u THREADS.
template <int LOG2N> FCF_HD int pass0_index(int tid, int slot) {
  using P = Plan<LOG2N>;
  constexpr int R = 1 << P::R0;
  const int u = slot / R, q = slot % R;
  return tid + u * P::THREADS + q * (P::N >> P::R0);
}

// forward pass 0: v = inputs (slot convention above) -> shared memory
template <int LOG2N, typename D> FCF_HD void fwd_pass0(int tid, cx<D> *v, cx<D> *buf, const cx<scalar_t<D>> *TW) {
  using P = Plan<LOG2N>;
  constexpr int R = 1 << P::R0, U = 16 / R, S0 = P::N >> P::R0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int p = tid + u * P::THREADS;
    cx<scalar_t<D>> tw[R];
    load_twiddles<R, recomp_twiddles<D>::value>(tw, TW, S0, p); // W_N^{p s}
    dft_small<-1, P::R0>(v + u * R);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      cx<D> y = v[u * R + s];
      if (s > 0) y = cmul(y, tw[s]);
      buf[phys(p + s * S0)] = y;
    }
  }
}
// inverse pass 0: shared memory -> v = outputs (same slot convention)
template <int LOG2N, typename D> FCF_HD void inv_pass0(int tid, cx<D> *v, const cx<D> *buf, const cx<scalar_t<D>> *TW) {
  using P = Plan<LOG2N>;
  constexpr int R = 1 << P::R0, U = 16 / R, S0 = P::N >> P::R0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int p = tid + u * P::THREADS;
    cx<scalar_t<D>> tw[R];
    load_twiddles<R, recomp_twiddles<D>::value>(tw, TW, S0, p);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      cx<D> y = buf[phys(p + s * S0)];
      if (s > 0) y = cmulc(y, tw[s]);
      v[u * R + s] = y;
    }
    dft_small<+1, P::R0>(v + u * R);
  }
}

// middle radix-16 pass J (1 <= J <= P - 2): sub-transform length m = N >> (R0 + 4 (J - 1))
template <int LOG2N, int J, typename D> FCF_HD void fwd_mid(int tid, cx<D> *buf, const cx<scalar_t<D>> *TWALL) {
  using P = Plan<LOG2N>;
  constexpr int LM = LOG2N - P::R0 - 4 * (J - 1), LS = LM - 4;
  constexpr int OFF = P::tw_offset(J);
  const cx<scalar_t<D>> *TW = TWALL + OFF;
  const int p = tid & ((1 << LS) - 1), base = ((tid >> LS) << LM) + p;
  cx<D> v[16];
  cx<scalar_t<D>> tw[16];
  load_twiddles<16, recomp_twiddles<D>::value>(tw, TW, 1 << LS, p); // W_m^{p s}
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = buf[phys(base + (q << LS))];
  dft16<-1>(v);
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    cx<D> y = v[s];
    if (s > 0) y = cmul(y, tw[s]);
    buf[phys(base + (s << LS))] = y;
  }
}
template <int LOG2N, int J, typename D> FCF_HD void inv_mid(int tid, cx<D> *buf, const cx<scalar_t<D>> *TWALL) {
  using P = Plan<LOG2N>;
  constexpr int LM = LOG2N - P::R0 - 4 * (J - 1), LS = LM - 4;
  constexpr int OFF = P::tw_offset(J);
  const cx<scalar_t<D>> *TW = TWALL + OFF;
  const int p = tid & ((1 << LS) - 1), base = ((tid >> LS) << LM) + p;
  cx<D> v[16];
  cx<scalar_t<D>> tw[16];
  load_twiddles<16, recomp_twiddles<D>::value>(tw, TW, 1 << LS, p);
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    cx<D> y = buf[phys(base + (s << LS))];
    if (s > 0) y = cmulc(y, tw[s]);
    v[s] = y;
  }
  dft16<+1>(v);
#pragma unroll
  for (int q = 0; q < 16; ++q) buf[phys(base + (q << LS))] = v[q];
}

// last forward pass (m = 16) + spectrum multiply + first inverse pass, in registers.
// H is s-major: H[s THREADS + tid] belongs to position 16 tid + s (coalesced reads).
// The spectrum multiplier applied between the transforms: a table (the filter's spectrum), or
// the Hilbert transformer -j sgn(f) / N, which needs no table: position 16 tid + s holds
// frequency (N / 16) s + (lower digits of tid), so s < 8 are the positive frequencies, and DC /
// Nyquist are (tid, s) = (0, 0) / (0, 8)  (/root/reference/include/fftconv/hilbert.hpp:39-51, :57).
template <typename T> struct TableSpectrum {
  const cx<T> *H;
  int threads;
  template <typename D> FCF_HD cx<D> apply(cx<D> v, int tid, int s) const { return cmul(v, H[s * threads + tid]); }
};
template <typename T> struct HilbertSpectrum {
  T inv_n;
  template <typename D> FCF_HD cx<D> apply(cx<D> v, int tid, int s) const {
    const T g = (tid == 0 && (s & 7) == 0) ? (T)0 : (s < 8 ? -inv_n : inv_n); // times j g
    return mul_j(v, g);
  }
};
template <int LOG2N, typename D, class Spectrum> FCF_HD void last_fused(int tid, cx<D> *buf, const Spectrum &H) {
  cx<D> v[16];
  const int base = 16 * tid;
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = buf[phys(base + q)];
  dft16<-1>(v);
#pragma unroll
  for (int s = 0; s < 16; ++s) v[s] = H.apply(v[s], tid, s);
  dft16<+1>(v);
#pragma unroll
  for (int q = 0; q < 16; ++q) buf[phys(base + q)] = v[q];
}

// Synchronisation between two passes.  A pass whose sub-transforms have 2^LM <= 512 points keeps
// every warp inside its own 512-point region ([512 w, 512 w + 512), w = warp index within the FFT:
// mid pass blocks tid >> (LM - 4), last pass points 16 tid ...), so an exchange whose wider side is
// such a pass is warp-local: `sync.warp()` (__syncwarp) replaces the barrier among all the FFT's
// threads.  For N = 8192 that is two of six boundaries (mid2 -> last, last -> inverse mid2), for
// N <= 4096 two of four: the warps of an FFT drift apart there and one warp's shared-memory phase
// overlaps another's arithmetic.
// Between those extremes the threads that must meet are the 2^LM / 16 threads of one sub-transform
// (after the radix-2 first pass of N = 8192 the two 4096-point halves are independent until the
// mirrored last pass): `sync.sub<LM>(tid)` lets a kernel give every such block its own named barrier.
template <int LM, class Sync> FCF_HD void pass_sync(Sync &sync, int tid) {
  if constexpr (LM <= 9) sync.warp();
  else sync.template sub<LM>(tid);
}
template <int LOG2N, int J> struct MidLm { // log2 of the sub-transform length of middle pass J
  static constexpr int value = LOG2N - Plan<LOG2N>::R0 - 4 * (J - 1);
};

// compile-time loops over the middle passes
template <int LOG2N, int J, typename T> struct FwdMids { // T = data type
  template <class Sync> static FCF_HD void run(int tid, cx<T> *buf, const cx<scalar_t<T>> *W, Sync &sync) {
    if constexpr (J <= Plan<LOG2N>::P - 2) {
      fwd_mid<LOG2N, J>(tid, buf, W);
      pass_sync<MidLm<LOG2N, J>::value>(sync, tid); // the next pass is narrower
      FwdMids<LOG2N, J + 1, T>::run(tid, buf, W, sync);
    }
  }
};
template <int LOG2N, int J, typename T> struct InvMids {
  template <class Sync> static FCF_HD void run(int tid, cx<T> *buf, const cx<scalar_t<T>> *W, Sync &sync) {
    if constexpr (J >= 1) {
      inv_mid<LOG2N, J>(tid, buf, W);
      pass_sync<(J > 1 ? MidLm<LOG2N, (J > 1 ? J - 1 : 1)>::value : LOG2N)>(sync, tid); // the next pass is wider
      InvMids<LOG2N, J - 1, T>::run(tid, buf, W, sync);
    }
  }
};

// Whole circular convolution with the spectrum H: v (inputs, pass-0 slot convention)
// -> v (outputs, same convention).  `sync` is the barrier among the FFT's threads.
template <int LOG2N, typename T, class Spectrum, class Sync>
FCF_HD void conv_core_with(int tid, cx<T> *v, cx<T> *buf, const cx<scalar_t<T>> *W /* pass twiddle tables */,
                           const Spectrum &H, Sync &sync) {
  fwd_pass0<LOG2N>(tid, v, buf, W);
  pass_sync<LOG2N>(sync, tid);
  FwdMids<LOG2N, 1, T>::run(tid, buf, W, sync);
  last_fused<LOG2N>(tid, buf, H);
  pass_sync<8>(sync, tid); // the last middle pass always has 256-point sub-transforms
  InvMids<LOG2N, Plan<LOG2N>::P - 2, T>::run(tid, buf, W, sync);
  inv_pass0<LOG2N>(tid, v, buf, W);
}
template <int LOG2N, typename T, class Sync>
FCF_HD void conv_core(int tid, cx<T> *v, cx<T> *buf, const cx<scalar_t<T>> *W /* pass twiddle tables */,
                      const cx<scalar_t<T>> *Htab, Sync &sync) {
  const TableSpectrum<scalar_t<T>> H{Htab, Plan<LOG2N>::THREADS};
  conv_core_with<LOG2N>(tid, v, buf, W, H, sync);
}

// Host: fill the per-pass twiddle tables (TW_ELEMS entries) in long double.
template <int LOG2N, typename T, class RootFn> void fill_pass_twiddles(cx<T> *tw, RootFn root /* (num, den) -> exp(-2 pi i num/den) */) {
  using P = Plan<LOG2N>;
  constexpr int R = 1 << P::R0, S0 = P::N >> P::R0;
  for (int s = 1; s < R; ++s)
    for (int p = 0; p < S0; ++p) tw[(s - 1) * S0 + p] = root((long long)p * s, (long long)P::N);
  int lm = LOG2N - P::R0;
  for (int j = 1; j <= P::P - 2; ++j, lm -= 4) {
    cx<T> *t = tw + P::tw_offset(j);
    const int ls = lm - 4;
    for (int s = 1; s < 16; ++s)
      for (int p = 0; p < (1 << ls); ++p) t[((s - 1) << ls) + p] = root((long long)p * s, 1LL << lm);
  }
}

} // namespace fast
} // namespace fc
