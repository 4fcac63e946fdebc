/*
 * oracle/fftw_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A small CPU FFT that implements, with FFTW 3.3's *documented* semantics, the
 * nine FFTW entry points per precision that the reference's convolution hot
 * path calls (see oracle/fftw3.h for the call sites).  FFTW itself is an
 * un-vendored third-party dependency (vcpkg.json:9-23, fftw3 3.3.10) that is
 * absent from /root/reference and from this image, so the reference headers
 * are compiled UNMODIFIED on top of this shim (oracle/ref_capi.cpp).
 *
 * Semantics restated from the FFTW manual (sections "One-Dimensional DFTs of
 * Real Data", "What FFTW Really Computes", "New-array Execute Functions"):
 *   r2c:  Y[k] = sum_j x[j] exp(-2 pi i j k / n),  k = 0 .. n/2      (no scale)
 *   c2r:  x[j] = sum_k Y[k] exp(+2 pi i j k / n) with Y Hermitian-extended,
 *         i.e. c2r(r2c(x)) = n * x; the input array may be overwritten.
 *   planning never touches the arrays with FFTW_ESTIMATE; with other flags the
 *   arrays may be clobbered (we never touch them).
 *
 * Algorithm (ours, not FFTW's): power-of-two complex FFT = Stockham autosort,
 * radix-4 passes plus one radix-2 pass, twiddles tabulated in double and
 * rounded once; any other length = Bluestein chirp-z over the power-of-two
 * engine (lengths <= 64: direct O(n^2) sum accumulated in double).  Even-length
 * real transforms use the half-length complex packing, odd lengths go through
 * a full complex transform.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load the library this file is linked into.
 */
#include "fftw3.h"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <vector>

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;

inline bool is_pow2(size_t n) { return n != 0 && (n & (n - 1)) == 0; }

template <class R> using Cx = std::complex<R>;

// exp(sign * 2 pi i * num / den) evaluated in long double, octant-reduced.
template <class R> Cx<R> unit_root(long long num, long long den, int sign) {
  num %= den;
  if (num < 0) num += den;
  const long double a = 2.0L * static_cast<long double>(kPi) *
                        static_cast<long double>(num) /
                        static_cast<long double>(den);
  return Cx<R>(static_cast<R>(cosl(a)), static_cast<R>(sign * sinl(a)));
}

// ---------------------------------------------------------------------------
// Power-of-two complex FFT (Stockham autosort).
// ---------------------------------------------------------------------------
template <class R> struct Pow2Engine {
  size_t n;
  std::vector<Cx<R>> w; // w[k] = exp(-2 pi i k / n)

  explicit Pow2Engine(size_t n_) : n(n_), w(n_) {
    for (size_t k = 0; k < n; ++k) w[k] = unit_root<R>((long long)k, (long long)n, -1);
  }

  // Transforms x (length n) in place; y is scratch of length n.
  // sign = -1 forward, +1 backward (unnormalised).
  void run(Cx<R> *x, Cx<R> *y, int sign) const {
    if (n == 1) return;
    Cx<R> *src = x, *dst = y;
    size_t len = n, s = 1;
    const bool fwd = sign < 0;
    while (len >= 4) {
      const size_t q4 = len / 4;
      for (size_t p = 0; p < q4; ++p) {
        Cx<R> w1 = w[p * s], w2 = w[2 * p * s], w3 = w[3 * p * s];
        if (!fwd) { w1 = std::conj(w1); w2 = std::conj(w2); w3 = std::conj(w3); }
        const Cx<R> *a = src + s * p, *b = src + s * (p + q4),
                    *c = src + s * (p + 2 * q4), *d = src + s * (p + 3 * q4);
        Cx<R> *o = dst + s * 4 * p;
        for (size_t q = 0; q < s; ++q) {
          const R ar = a[q].real(), ai = a[q].imag(), br = b[q].real(), bi = b[q].imag();
          const R cr = c[q].real(), ci = c[q].imag(), dr = d[q].real(), di = d[q].imag();
          const R apc_r = ar + cr, apc_i = ai + ci, amc_r = ar - cr, amc_i = ai - ci;
          const R bpd_r = br + dr, bpd_i = bi + di, bmd_r = br - dr, bmd_i = bi - di;
          // j*(b-d) = (-bmd_i, bmd_r)
          R t1r, t1i, t3r, t3i;
          if (fwd) { // k=1: amc - j*bmd ; k=3: amc + j*bmd
            t1r = amc_r + bmd_i; t1i = amc_i - bmd_r;
            t3r = amc_r - bmd_i; t3i = amc_i + bmd_r;
          } else {
            t1r = amc_r - bmd_i; t1i = amc_i + bmd_r;
            t3r = amc_r + bmd_i; t3i = amc_i - bmd_r;
          }
          const R t2r = apc_r - bpd_r, t2i = apc_i - bpd_i;
          o[q] = Cx<R>(apc_r + bpd_r, apc_i + bpd_i);
          o[q + s] = Cx<R>(t1r * w1.real() - t1i * w1.imag(), t1r * w1.imag() + t1i * w1.real());
          o[q + 2 * s] = Cx<R>(t2r * w2.real() - t2i * w2.imag(), t2r * w2.imag() + t2i * w2.real());
          o[q + 3 * s] = Cx<R>(t3r * w3.real() - t3i * w3.imag(), t3r * w3.imag() + t3i * w3.real());
        }
      }
      std::swap(src, dst);
      len /= 4;
      s *= 4;
    }
    if (len == 2) {
      for (size_t q = 0; q < s; ++q) {
        const Cx<R> a = src[q], b = src[q + s];
        dst[q] = a + b;
        dst[q + s] = a - b;
      }
      std::swap(src, dst);
    }
    if (src != x) std::memcpy(x, src, n * sizeof(Cx<R>));
  }
};

// ---------------------------------------------------------------------------
// Arbitrary-length complex FFT.
// ---------------------------------------------------------------------------
template <class R> struct CfftEngine {
  size_t n;
  enum Kind { POW2, DIRECT, BLUESTEIN } kind;
  std::unique_ptr<Pow2Engine<R>> p2;          // POW2: size n; BLUESTEIN: size m
  std::vector<std::complex<double>> roots;    // DIRECT: exp(-2 pi i k / n)
  size_t m = 0;                               // BLUESTEIN padded length
  std::vector<Cx<R>> chirp;                   // exp(-pi i k^2 / n)
  std::vector<Cx<R>> chirp_spec;              // FFT_m of conj(chirp), wrapped

  explicit CfftEngine(size_t n_) : n(n_) {
    if (is_pow2(n)) {
      kind = POW2;
      p2 = std::make_unique<Pow2Engine<R>>(n);
    } else if (n <= 64) {
      kind = DIRECT;
      roots.resize(n);
      for (size_t k = 0; k < n; ++k) roots[k] = unit_root<double>((long long)k, (long long)n, -1);
    } else {
      kind = BLUESTEIN;
      m = 1;
      while (m < 2 * n - 1) m *= 2;
      p2 = std::make_unique<Pow2Engine<R>>(m);
      chirp.resize(n);
      for (size_t k = 0; k < n; ++k) {
        // k^2 mod 2n keeps the angle small and exact
        const long long k2 = (long long)((unsigned long long)k * k % (2 * n));
        chirp[k] = unit_root<R>(k2, 2 * (long long)n, -1);
      }
      chirp_spec.assign(m, Cx<R>(0, 0));
      chirp_spec[0] = std::conj(chirp[0]);
      for (size_t k = 1; k < n; ++k) chirp_spec[k] = chirp_spec[m - k] = std::conj(chirp[k]);
      std::vector<Cx<R>> tmp(m);
      p2->run(chirp_spec.data(), tmp.data(), -1);
    }
  }

  size_t scratch_len() const { return kind == BLUESTEIN ? 2 * m : n; }

  // In-place transform of x[0..n); scratch has scratch_len() entries.
  void run(Cx<R> *x, Cx<R> *scratch, int sign) const {
    switch (kind) {
    case POW2: p2->run(x, scratch, sign); break;
    case DIRECT: {
      for (size_t k = 0; k < n; ++k) {
        double sr = 0, si = 0;
        for (size_t j = 0; j < n; ++j) {
          std::complex<double> w = roots[(j * k) % n];
          if (sign > 0) w = std::conj(w);
          sr += (double)x[j].real() * w.real() - (double)x[j].imag() * w.imag();
          si += (double)x[j].real() * w.imag() + (double)x[j].imag() * w.real();
        }
        scratch[k] = Cx<R>((R)sr, (R)si);
      }
      std::memcpy(x, scratch, n * sizeof(Cx<R>));
      break;
    }
    case BLUESTEIN: {
      // backward transform = conj(forward(conj(x)))
      Cx<R> *a = scratch, *tmp = scratch + m;
      for (size_t j = 0; j < n; ++j) {
        const Cx<R> v = sign < 0 ? x[j] : std::conj(x[j]);
        a[j] = v * chirp[j];
      }
      for (size_t j = n; j < m; ++j) a[j] = Cx<R>(0, 0);
      p2->run(a, tmp, -1);
      for (size_t j = 0; j < m; ++j) a[j] *= chirp_spec[j];
      p2->run(a, tmp, +1);
      const R inv = (R)(1.0 / (double)m);
      for (size_t k = 0; k < n; ++k) {
        const Cx<R> v = a[k] * chirp[k] * inv;
        x[k] = sign < 0 ? v : std::conj(v);
      }
      break;
    }
    }
  }
};

// ---------------------------------------------------------------------------
// Real transforms of length n on top of the complex engine.
// ---------------------------------------------------------------------------
template <class R> struct RealEngine {
  size_t n;
  bool even;
  std::unique_ptr<CfftEngine<R>> c;    // length n/2 (even) or n (odd)
  std::vector<Cx<R>> w;                // even: exp(-2 pi i k / n), k <= n/2

  explicit RealEngine(size_t n_) : n(n_), even(n_ % 2 == 0 && n_ >= 2) {
    c = std::make_unique<CfftEngine<R>>(even ? n / 2 : n);
    if (even) {
      w.resize(n / 2 + 1);
      for (size_t k = 0; k <= n / 2; ++k) w[k] = unit_root<R>((long long)k, (long long)n, -1);
    }
  }
  size_t work_len() const { return (even ? n / 2 : n) + c->scratch_len(); }

  void r2c(const R *in, Cx<R> *out, Cx<R> *work) const {
    if (even) {
      const size_t h = n / 2;
      Cx<R> *z = work, *scr = work + h;
      for (size_t j = 0; j < h; ++j) z[j] = Cx<R>(in[2 * j], in[2 * j + 1]);
      c->run(z, scr, -1);
      for (size_t k = 0; k <= h; ++k) {
        const Cx<R> zk = z[k % h], zc = std::conj(z[(h - k) % h]);
        const Cx<R> e = (zk + zc) * (R)0.5;
        const Cx<R> d = (zk - zc) * (R)0.5;              // = i * O[k]
        const Cx<R> o(d.imag(), -d.real());              // O[k] = d / i
        out[k] = e + w[k] * o;
      }
    } else {
      Cx<R> *z = work, *scr = work + n;
      for (size_t j = 0; j < n; ++j) z[j] = Cx<R>(in[j], 0);
      c->run(z, scr, -1);
      for (size_t k = 0; k <= n / 2; ++k) out[k] = z[k];
    }
  }

  void c2r(const Cx<R> *in, R *out, Cx<R> *work) const {
    if (even) {
      const size_t h = n / 2;
      Cx<R> *z = work, *scr = work + h;
      for (size_t k = 0; k < h; ++k) {
        // FFTW ignores the imaginary parts of the DC and Nyquist bins.
        Cx<R> xk = in[k], xc = std::conj(in[h - k]);
        if (k == 0) { xk = Cx<R>(in[0].real(), 0); xc = Cx<R>(in[h].real(), 0); }
        const Cx<R> e = xk + xc;
        const Cx<R> o = (xk - xc) * std::conj(w[k]);
        z[k] = e + Cx<R>(-o.imag(), o.real());           // e + i*o
      }
      c->run(z, scr, +1);
      for (size_t j = 0; j < h; ++j) { out[2 * j] = z[j].real(); out[2 * j + 1] = z[j].imag(); }
    } else {
      Cx<R> *z = work, *scr = work + n;
      z[0] = Cx<R>(in[0].real(), 0);
      for (size_t k = 1; k <= n / 2; ++k) { z[k] = in[k]; z[n - k] = std::conj(in[k]); }
      c->run(z, scr, +1);
      for (size_t j = 0; j < n; ++j) out[j] = z[j].real();
    }
  }
};

template <class R> std::shared_ptr<const RealEngine<R>> get_engine(size_t n) {
  static std::mutex mu;
  static std::map<size_t, std::shared_ptr<const RealEngine<R>>> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto &e = cache[n];
  if (!e) e = std::make_shared<const RealEngine<R>>(n);
  return e;
}

template <class R> struct PlanImpl {
  std::shared_ptr<const RealEngine<R>> eng;
  bool forward;        // true: r2c, false: c2r
  R *real;
  Cx<R> *cplx;
  std::vector<Cx<R>> work;
};

template <class R> PlanImpl<R> *make_plan(int n, R *real, void *cplx, bool forward) {
  if (n <= 0) return nullptr;
  auto *p = new PlanImpl<R>();
  p->eng = get_engine<R>((size_t)n);
  p->forward = forward;
  p->real = real;
  p->cplx = reinterpret_cast<Cx<R> *>(cplx);
  p->work.resize(p->eng->work_len());
  return p;
}

template <class R> void run_plan(PlanImpl<R> *p, R *real, void *cplx) {
  if (p->forward) p->eng->r2c(real, reinterpret_cast<Cx<R> *>(cplx), p->work.data());
  else p->eng->c2r(reinterpret_cast<const Cx<R> *>(cplx), real, p->work.data());
}

void *aligned64(size_t bytes) {
  if (bytes == 0) bytes = 64;
  void *p = nullptr;
  if (posix_memalign(&p, 64, (bytes + 63) / 64 * 64) != 0) return nullptr;
  return p;
}

} // namespace

#define SHIM_API(X, R, C)                                                      \
  X(plan) X(plan_dft_r2c_1d)(int n, R *in, C *out, unsigned) {                 \
    return reinterpret_cast<X(plan)>(make_plan<R>(n, in, out, true));          \
  }                                                                            \
  X(plan) X(plan_dft_c2r_1d)(int n, C *in, R *out, unsigned) {                 \
    return reinterpret_cast<X(plan)>(make_plan<R>(n, out, in, false));         \
  }                                                                            \
  void X(execute)(const X(plan) p) {                                           \
    auto *q = reinterpret_cast<PlanImpl<R> *>(p);                              \
    run_plan<R>(q, q->real, q->cplx);                                          \
  }                                                                            \
  void X(execute_dft_r2c)(const X(plan) p, R *in, C *out) {                    \
    run_plan<R>(reinterpret_cast<PlanImpl<R> *>(p), in, out);                  \
  }                                                                            \
  void X(execute_dft_c2r)(const X(plan) p, C *in, R *out) {                    \
    run_plan<R>(reinterpret_cast<PlanImpl<R> *>(p), out, in);                  \
  }                                                                            \
  void X(destroy_plan)(X(plan) p) {                                            \
    delete reinterpret_cast<PlanImpl<R> *>(p);                                 \
  }                                                                            \
  R *X(alloc_real)(size_t n) {                                                 \
    return static_cast<R *>(aligned64(n * sizeof(R)));                         \
  }                                                                            \
  C *X(alloc_complex)(size_t n) {                                              \
    return static_cast<C *>(aligned64(n * sizeof(C)));                         \
  }                                                                            \
  void *X(malloc)(size_t n) { return aligned64(n); }                           \
  void X(free)(void *p) { std::free(p); }                                      \
  void X(make_planner_thread_safe)(void) {}                                    \
  int X(import_wisdom_from_filename)(const char *) { return 0; }               \
  int X(export_wisdom_to_filename)(const char *) { return 0; }

extern "C" {
SHIM_API(ORACLE_MANGLE_D, double, fftw_complex)
SHIM_API(ORACLE_MANGLE_F, float, fftwf_complex)
}
