// include/fftconv/fftw.hpp -- what is left of the reference's FFTW wrapper on a GPU.
//
// The reference header (/root/reference/include/fftconv/fftw.hpp) wraps the whole
// FFTW planner API; the convolution hot path only reaches its r2c/c2r plans, the
// thread_local plan cache and the normalise loops (:134-147, :180-200, :396-461,
// :595-652).  All of that is replaced by libfftconv_cuda (plan/workspace cache +
// fused kernels), so this header keeps only the names callers of the hot path see:
//   fftw::Floating              fftw.hpp:47-48  (+ the README's FloatOrDouble alias)
//   fftw::WisdomSetup           fftw.hpp:30-45  -> no-op: there is nothing to tune
//   FFTW_ESTIMATE / MEASURE / PATIENT / EXHAUSTIVE: accepted as the PlannerFlag
//   template argument of convolve_fftw / oaconvolve_fftw and ignored.
// Round 2: fftw::Plan<T> is back for the 1-D transforms the
// C ABI serves (fftconv_cuda_fft_*): dft_1d, dft_r2c_1d, dft_c2r_1d and their contiguous
// batched ("many", rank 1, unit stride) forms, with execute() and the new-array execute
// functions -- same names, argument order and conventions (unnormalised, FFTW_FORWARD = -1)
// as the reference's wrapper (fftw.hpp:134-147, :158-160, :180-182, :197-199, :244-278,
// :396-423).  A plan holds no tables: it records the shape and the pointers it was planned
// with; the library caches twiddles per device.  Pointers may be host or device pointers.
// 2-D / 3-D, r2r, strided "many" and guru plans stay out of scope.
#pragma once

#include "../fftconv_cuda.h"

#include <cstddef>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>

// planner flag values as published in the FFTW 3.3 manual
#ifndef FFTW_ESTIMATE
#define FFTW_MEASURE (0U)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)
#endif
#ifndef FFTW_FORWARD
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#endif

namespace fftw {

template <typename T>
concept Floating = std::is_same_v<T, float> || std::is_same_v<T, double>;
template <typename T>
concept FloatOrDouble = Floating<T>;

// RAII object the reference's mains construct first (test/test_fftw.cpp:399,
// test/test_script.cpp:71).  FFTW wisdom has no GPU counterpart: the kernels are
// fixed and the tables are rebuilt in microseconds.
struct WisdomSetup {
  explicit WisdomSetup(bool /*threadSafe*/) {}
  WisdomSetup(const WisdomSetup &) = delete;
  WisdomSetup &operator=(const WisdomSetup &) = delete;
};

// FFTW's complex type: T[2] = (re, im), fftw.hpp:50-75
template <Floating T> using Complex = T[2];

// fftw::Plan<T> (fftw.hpp:134-147): the 1-D kinds, contiguous batches
template <Floating T> struct Plan {
  enum class Kind { none, c2c, r2c, c2r };
  Kind kind = Kind::none;
  std::size_t n = 0, howmany = 1;
  int sign = FFTW_FORWARD;
  const void *in = nullptr;
  void *out = nullptr;

  Plan() = default;
  Plan(const Plan &) = delete;
  Plan(Plan &&) = default;
  Plan &operator=(const Plan &) = delete;
  Plan &operator=(Plan &&) = default;

  // fftw.hpp:158-160, :180-182, :197-199 (flags are accepted and ignored: there is nothing to plan)
  [[nodiscard]] static Plan dft_1d(int n, Complex<T> *in, Complex<T> *out, int sign, unsigned /*flags*/) {
    return make(Kind::c2c, n, 1, in, out, sign);
  }
  [[nodiscard]] static Plan dft_r2c_1d(int n, T *in, Complex<T> *out, unsigned /*flags*/) {
    return make(Kind::r2c, n, 1, in, out, FFTW_FORWARD);
  }
  [[nodiscard]] static Plan dft_c2r_1d(int n, Complex<T> *in, T *out, unsigned /*flags*/) {
    return make(Kind::c2r, n, 1, in, out, FFTW_BACKWARD);
  }
  // fftw.hpp:244-278: rank 1, unit stride, rows back to back (idist / odist = the row length) -- the only
  // "many" layout the batched kernels take; anything else throws
  [[nodiscard]] static Plan many_dft(int rank, const int *n, int howmany, Complex<T> *in, const int *inembed,
                                     int istride, int idist, Complex<T> *out, const int *onembed, int ostride,
                                     int odist, int sign, unsigned /*flags*/) {
    contiguous(rank, n, inembed, istride, idist, n[0], onembed, ostride, odist, n[0]);
    return make(Kind::c2c, n[0], howmany, in, out, sign);
  }
  [[nodiscard]] static Plan many_dft_r2c(int rank, const int *n, int howmany, T *in, const int *inembed, int istride,
                                         int idist, Complex<T> *out, const int *onembed, int ostride, int odist,
                                         unsigned /*flags*/) {
    contiguous(rank, n, inembed, istride, idist, n[0], onembed, ostride, odist, n[0] / 2 + 1);
    return make(Kind::r2c, n[0], howmany, in, out, FFTW_FORWARD);
  }
  [[nodiscard]] static Plan many_dft_c2r(int rank, const int *n, int howmany, Complex<T> *in, const int *inembed,
                                         int istride, int idist, T *out, const int *onembed, int ostride, int odist,
                                         unsigned /*flags*/) {
    contiguous(rank, n, inembed, istride, idist, n[0] / 2 + 1, onembed, ostride, odist, n[0]);
    return make(Kind::c2r, n[0], howmany, in, out, FFTW_BACKWARD);
  }

  // fftw.hpp:396-423
  void execute() const { run(in, out); }
  void execute_dft(const Complex<T> *i, Complex<T> *o) const { expect(Kind::c2c); run(i, o); }
  void execute_dft_r2c(const T *i, Complex<T> *o) const { expect(Kind::r2c); run(i, o); }
  void execute_dft_c2r(const Complex<T> *i, T *o) const { expect(Kind::c2r); run(i, o); }

private:
  static Plan make(Kind k, int n, int howmany, const void *in, void *out, int sign) {
    if (n <= 0 || howmany <= 0) throw std::invalid_argument("fftw::Plan: n and howmany must be positive");
    Plan p;
    p.kind = k;
    p.n = (std::size_t)n;
    p.howmany = (std::size_t)howmany;
    p.sign = sign;
    p.in = in;
    p.out = out;
    return p;
  }
  static void contiguous(int rank, const int *n, const int *inembed, int istride, int idist, int irow,
                         const int *onembed, int ostride, int odist, int orow) {
    if (rank != 1 || !n || istride != 1 || ostride != 1 || idist != irow || odist != orow ||
        (inembed && inembed[0] != irow && inembed[0] != n[0]) || (onembed && onembed[0] != orow && onembed[0] != n[0]))
      throw std::invalid_argument("fftw::Plan: only rank-1, unit-stride batches with rows back to back are supported");
  }
  void expect(Kind k) const {
    if (kind != k) throw std::logic_error("fftw::Plan: executed with the arrays of another plan kind");
  }
  void run(const void *i, void *o) const {
    int rc = FFTCONV_CUDA_ERR_ARG;
    const T *ti = static_cast<const T *>(i);
    T *to = static_cast<T *>(o);
    if constexpr (std::is_same_v<T, float>) {
      if (kind == Kind::c2c) rc = fftconv_cuda_fft_c2c_f32(ti, to, n, howmany, sign, nullptr);
      if (kind == Kind::r2c) rc = fftconv_cuda_fft_r2c_f32(ti, to, n, howmany, nullptr);
      if (kind == Kind::c2r) rc = fftconv_cuda_fft_c2r_f32(ti, to, n, howmany, nullptr);
    } else {
      if (kind == Kind::c2c) rc = fftconv_cuda_fft_c2c_f64(ti, to, n, howmany, sign, nullptr);
      if (kind == Kind::r2c) rc = fftconv_cuda_fft_r2c_f64(ti, to, n, howmany, nullptr);
      if (kind == Kind::c2r) rc = fftconv_cuda_fft_c2r_f64(ti, to, n, howmany, nullptr);
    }
    if (rc != 0) throw std::runtime_error(std::string("fftw::Plan: ") + fftconv_cuda_last_error());
  }
};

// ---- host buffers and the per-length engines built on Plan (fftw.hpp:77-104, :441-600) -----------------------
// fftw_malloc's 64-byte aligned host arrays; the transforms copy them to the device and back
// (fftconv_cuda_fft_* with host pointers), so code written against the reference's EngineDFT1D /
// EngineR2C1D -- buffers `buf.in` / `buf.out`, forward(), backward(), the thread_local cache behind
// get(n) -- runs unchanged.
inline void *malloc(std::size_t bytes) {
  void *p = std::aligned_alloc(64, (bytes + 63) / 64 * 64);
  if (!p) throw std::bad_alloc();
  return p;
}
template <Floating T> T *alloc_real(std::size_t n) { return static_cast<T *>(fftw::malloc(n * sizeof(T))); }
template <Floating T> Complex<T> *alloc_complex(std::size_t n) {
  return static_cast<Complex<T> *>(fftw::malloc(n * sizeof(Complex<T>)));
}
template <Floating T> void free(void *p) { std::free(p); }

template <typename T, bool InPlace = false> struct C2CBuffer {
  using Cx = Complex<T>;
  Cx *in, *out;
  explicit C2CBuffer(std::size_t n) : in(alloc_complex<T>(n)), out(InPlace ? in : alloc_complex<T>(n)) {}
  C2CBuffer(const C2CBuffer &) = delete;
  C2CBuffer &operator=(const C2CBuffer &) = delete;
  ~C2CBuffer() noexcept {
    if (!InPlace) fftw::free<T>(out);
    fftw::free<T>(in);
  }
};
template <typename T> struct R2CBuffer {
  using Cx = Complex<T>;
  T *in;
  Cx *out;
  explicit R2CBuffer(std::size_t n) : in(alloc_real<T>(n)), out(alloc_complex<T>(n / 2 + 1)) {}
  R2CBuffer(const R2CBuffer &) = delete;
  R2CBuffer &operator=(const R2CBuffer &) = delete;
  ~R2CBuffer() noexcept {
    fftw::free<T>(in);
    fftw::free<T>(out);
  }
};

// one engine per transform length and thread, created on first use (fftw.hpp:443-461)
template <typename Child> struct cache_mixin {
  static Child &get(std::size_t n) {
    thread_local std::unordered_map<std::size_t, std::unique_ptr<Child>> cache;
    auto &slot = cache[n];
    if (!slot) slot = std::make_unique<Child>(n);
    return *slot;
  }
};

template <Floating T, int PlannerFlag = FFTW_ESTIMATE, bool InPlace = false>
struct EngineDFT1D : public cache_mixin<EngineDFT1D<T, PlannerFlag, InPlace>> {
  using Cx = Complex<T>;
  C2CBuffer<T, InPlace> buf;
  Plan<T> plan_forward, plan_backward;
  explicit EngineDFT1D(std::size_t n)
      : buf(n), plan_forward(Plan<T>::dft_1d((int)n, buf.in, buf.out, FFTW_FORWARD, PlannerFlag)),
        plan_backward(Plan<T>::dft_1d((int)n, buf.out, buf.in, FFTW_BACKWARD, PlannerFlag)) {}
  void forward() { plan_forward.execute(); }
  void forward(const Cx *in, Cx *out) const { plan_forward.execute_dft(in, out); }
  void backward() { plan_backward.execute(); }
  void backward(const Cx *in, Cx *out) const { plan_backward.execute_dft(in, out); }
};

template <Floating T, int PlannerFlag = FFTW_ESTIMATE>
struct EngineR2C1D : public cache_mixin<EngineR2C1D<T, PlannerFlag>> {
  using Cx = Complex<T>;
  R2CBuffer<T> buf;
  Plan<T> plan_forward, plan_backward;
  explicit EngineR2C1D(std::size_t n)
      : buf(n), plan_forward(Plan<T>::dft_r2c_1d((int)n, buf.in, buf.out, PlannerFlag)),
        plan_backward(Plan<T>::dft_c2r_1d((int)n, buf.out, buf.in, PlannerFlag)) {}
  void forward() { plan_forward.execute(); }
  void forward(const T *in, Cx *out) const { plan_forward.execute_dft_r2c(in, out); }
  void backward() { plan_backward.execute(); }
  void backward(const Cx *in, T *out) const { plan_backward.execute_dft_c2r(in, out); }
};

} // namespace fftw
