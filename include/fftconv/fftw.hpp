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
// The c2c / r2r / many / guru wrappers of the reference are out of scope (SURVEY.md
// section 2, row 2).
#pragma once

#include <type_traits>

// planner flag values as published in the FFTW 3.3 manual
#ifndef FFTW_ESTIMATE
#define FFTW_MEASURE (0U)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)
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

} // namespace fftw
