/*
 * oracle/fftw3.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Declarations-only stand-in for FFTW3's public header.  FFTW 3.3.10 (pulled
 * by the reference's vcpkg manifest, vcpkg.json:5-23) is a third-party
 * dependency that is NOT vendored in /root/reference and is not installed in
 * this image.  The reference headers (include/fftconv/fftw.hpp:9) only need a
 * header that declares FFTW's documented C API; the nine entry points per
 * precision that the convolution hot path really calls
 * (include/fftconv/fftconv.hpp:327-330,345,350,360,363,374,392,404,406,421,423;
 * include/fftconv/hilbert.hpp:27,35,54) are implemented in fftw_shim.cpp with
 * FFTW's published semantics (unnormalised, forward sign -1, r2c returns
 * n/2+1 bins, c2r may destroy its input).
 *
 * Everything declared here follows the FFTW 3.3 manual ("FFTW Reference"),
 * not FFTW's source.
 */
#ifndef ORACLE_FFTW3_STUB_H
#define ORACLE_FFTW3_STUB_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_UNALIGNED (1U << 1)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

typedef enum {
  FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2, FFTW_REDFT00 = 3,
  FFTW_REDFT01 = 4, FFTW_REDFT10 = 5, FFTW_REDFT11 = 6, FFTW_RODFT00 = 7,
  FFTW_RODFT01 = 8, FFTW_RODFT10 = 9, FFTW_RODFT11 = 10
} fftw_r2r_kind_do_not_use_me;

struct oracle_iodim { int n, is, os; };
struct oracle_iodim64 { ptrdiff_t n, is, os; };

#define ORACLE_FFTW_API(X, R, C)                                               \
  typedef R C[2];                                                              \
  typedef struct X(plan_s) * X(plan);                                          \
  typedef struct oracle_iodim X(iodim);                                        \
  typedef struct oracle_iodim64 X(iodim64);                                    \
  typedef fftw_r2r_kind_do_not_use_me X(r2r_kind);                             \
  /* the nine entry points the hot path uses (implemented in fftw_shim.cpp) */ \
  X(plan) X(plan_dft_r2c_1d)(int n, R *in, C *out, unsigned flags);            \
  X(plan) X(plan_dft_c2r_1d)(int n, C *in, R *out, unsigned flags);            \
  void X(execute)(const X(plan) p);                                            \
  void X(execute_dft_r2c)(const X(plan) p, R *in, C *out);                     \
  void X(execute_dft_c2r)(const X(plan) p, C *in, R *out);                     \
  void X(destroy_plan)(X(plan) p);                                             \
  R *X(alloc_real)(size_t n);                                                  \
  C *X(alloc_complex)(size_t n);                                               \
  void X(free)(void *p);                                                       \
  /* named by non-dependent code in fftw.hpp:30-45,99 (also in the shim) */    \
  void *X(malloc)(size_t n);                                                   \
  void X(make_planner_thread_safe)(void);                                      \
  int X(import_wisdom_from_filename)(const char *filename);                    \
  int X(export_wisdom_to_filename)(const char *filename);                      \
  /* rest of the planner API: only named inside never-instantiated member   */ \
  /* templates of fftw::Plan<T> (fftw.hpp:158-394); declared, never defined */ \
  X(plan) X(plan_dft_1d)(int, C *, C *, int, unsigned);                        \
  X(plan) X(plan_dft_2d)(int, int, C *, C *, int, unsigned);                   \
  X(plan) X(plan_dft_3d)(int, int, int, C *, C *, int, unsigned);              \
  X(plan) X(plan_dft)(int, const int *, C *, C *, int, unsigned);              \
  X(plan) X(plan_dft_r2c_2d)(int, int, R *, C *, unsigned);                    \
  X(plan) X(plan_dft_r2c_3d)(int, int, int, R *, C *, unsigned);               \
  X(plan) X(plan_dft_r2c)(int, const int *, R *, C *, unsigned);               \
  X(plan) X(plan_dft_c2r_2d)(int, int, C *, R *, unsigned);                    \
  X(plan) X(plan_dft_c2r_3d)(int, int, int, C *, R *, unsigned);               \
  X(plan) X(plan_dft_c2r)(int, const int *, C *, R *, unsigned);               \
  X(plan) X(plan_r2r_1d)(int, R *, R *, X(r2r_kind), unsigned);                \
  X(plan) X(plan_r2r_2d)(int, int, R *, R *, X(r2r_kind), X(r2r_kind),         \
                         unsigned);                                            \
  X(plan) X(plan_r2r_3d)(int, int, int, R *, R *, X(r2r_kind), X(r2r_kind),    \
                         X(r2r_kind), unsigned);                               \
  X(plan) X(plan_r2r)(int, const int *, R *, R *, const X(r2r_kind) *,         \
                      unsigned);                                               \
  X(plan) X(plan_many_dft)(int, const int *, int, C *, const int *, int, int,  \
                           C *, const int *, int, int, int, unsigned);         \
  X(plan) X(plan_many_dft_r2c)(int, const int *, int, R *, const int *, int,   \
                               int, C *, const int *, int, int, unsigned);     \
  X(plan) X(plan_many_dft_c2r)(int, const int *, int, C *, const int *, int,   \
                               int, R *, const int *, int, int, unsigned);     \
  X(plan) X(plan_many_r2r)(int, const int *, int, R *, const int *, int, int,  \
                           R *, const int *, int, int, const X(r2r_kind) *,    \
                           unsigned);                                          \
  X(plan) X(plan_guru_dft)(int, const X(iodim) *, int, const X(iodim) *, C *,  \
                           C *, int, unsigned);                                \
  X(plan) X(plan_guru_split_dft)(int, const X(iodim) *, int,                   \
                                 const X(iodim) *, R *, R *, R *, R *,         \
                                 unsigned);                                    \
  X(plan) X(plan_guru_dft_r2c)(int, const X(iodim) *, int, const X(iodim) *,   \
                               R *, C *, unsigned);                            \
  X(plan) X(plan_guru_split_dft_r2c)(int, const X(iodim) *, int,               \
                                     const X(iodim) *, R *, R *, R *,          \
                                     unsigned);                                \
  X(plan) X(plan_guru_dft_c2r)(int, const X(iodim) *, int, const X(iodim) *,   \
                               C *, R *, unsigned);                            \
  X(plan) X(plan_guru_split_dft_c2r)(int, const X(iodim) *, int,               \
                                     const X(iodim) *, R *, R *, R *,          \
                                     unsigned);                                \
  X(plan) X(plan_guru_r2r)(int, const X(iodim) *, int, const X(iodim) *, R *,  \
                           R *, const X(r2r_kind) *, unsigned);                \
  X(plan) X(plan_guru64_dft)(int, const X(iodim64) *, int,                     \
                             const X(iodim64) *, C *, C *, int, unsigned);     \
  X(plan) X(plan_guru64_split_dft)(int, const X(iodim64) *, int,               \
                                   const X(iodim64) *, R *, R *, R *, R *,     \
                                   unsigned);                                  \
  X(plan) X(plan_guru64_dft_r2c)(int, const X(iodim64) *, int,                 \
                                 const X(iodim64) *, R *, C *, unsigned);      \
  X(plan) X(plan_guru64_split_dft_r2c)(int, const X(iodim64) *, int,           \
                                       const X(iodim64) *, R *, R *, R *,      \
                                       unsigned);                              \
  X(plan) X(plan_guru64_dft_c2r)(int, const X(iodim64) *, int,                 \
                                 const X(iodim64) *, C *, R *, unsigned);      \
  X(plan) X(plan_guru64_split_dft_c2r)(int, const X(iodim64) *, int,           \
                                       const X(iodim64) *, R *, R *, R *,      \
                                       unsigned);                              \
  X(plan) X(plan_guru64_r2r)(int, const X(iodim64) *, int,                     \
                             const X(iodim64) *, R *, R *,                     \
                             const X(r2r_kind) *, unsigned);                   \
  void X(execute_dft)(const X(plan), C *, C *);                                \
  void X(execute_split_dft)(const X(plan), R *, R *, R *, R *);                \
  void X(execute_split_dft_r2c)(const X(plan), R *, R *, R *);                 \
  void X(execute_split_dft_c2r)(const X(plan), R *, R *, R *);                 \
  void X(execute_r2r)(const X(plan), R *, R *);

#define ORACLE_MANGLE_D(name) fftw_##name
#define ORACLE_MANGLE_F(name) fftwf_##name

ORACLE_FFTW_API(ORACLE_MANGLE_D, double, fftw_complex)
ORACLE_FFTW_API(ORACLE_MANGLE_F, float, fftwf_complex)

#ifdef __cplusplus
}
#endif
#endif
