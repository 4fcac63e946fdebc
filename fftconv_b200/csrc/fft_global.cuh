// fft_global.cuh -- the general 1-D FFT surface: batched complex FFTs of ANY
// length through global memory.  This is the wide, slow lane of the library: the fused kernels
// (oa_*, fast_*, long_fft_*) serve the hot path; these serve what they do not reach -- standalone
// transforms (the reference's FFTW wrappers dft_1d / dft_r2c_1d / dft_c2r_1d and their "many" forms,
// /root/reference/include/fftconv/fftw.hpp:134-200, :396-423) and Hilbert envelopes of lines longer
// than the on-chip and split transforms cover (the reference takes any n,
// /root/reference/include/fftconv/hilbert.hpp:16-22).
//
//   power-of-two n:  Stockham autosort passes, radix 4 (+ one radix-2 pass when log2 n is odd), one
//                    kernel per pass, ping-pong between two buffers; natural order in and out.
//                    Twiddles come from a table exp(-2 pi i k / n) evaluated in long double on the host.
//   any other n:     chirp-z (Bluestein): X[k] = c[k] sum_j (x[j] c[j]) conj(c)[k - j], c[j] = exp(-pi i j^2 / n),
//                    the convolution as a circular one of power-of-two length M >= 2 n - 1.
// Conventions are FFTW's: unnormalised, exp(-2 pi i jk / n) forward (sign = -1), exp(+...) backward.
#pragma once

#include <cuda_runtime.h>

namespace fc {
namespace gfft {

template <typename T> struct cxg {
  T x, y;
};
template <typename T> __device__ __forceinline__ cxg<T> cmul(cxg<T> a, cxg<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T> __device__ __forceinline__ cxg<T> conj_if(cxg<T> a, bool c) { return {a.x, c ? -a.y : a.y}; }

// One Stockham pass of radix R (2 or 4).  The array holds `batch` rows of n complex points; ns is the
// length of the sub-transforms already done.  Thread j of a row: k = j mod ns,
//   v[r] = in[j + r n / R] W_{ns R}^{k r};  DFT_R;  out[(j - k) R + k + r ns] = v[r]
// W[e] = exp(-2 pi i e / n); sign > 0 conjugates the twiddles and the small DFT.
template <typename T, int R>
__global__ void stockham_pass(const cxg<T> *__restrict__ in, cxg<T> *__restrict__ out, const cxg<T> *__restrict__ W,
                              long long n, long long ns, int sign, long long batch) {
  const long long per_row = n / R;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_row * batch) return;
  const long long row = idx / per_row, j = idx - row * per_row;
  const long long k = j & (ns - 1);
  const cxg<T> *src = in + row * n;
  cxg<T> *dst = out + row * n;
  const bool inv = sign > 0;
  const long long tstep = n / (ns * R); // W_{ns R}^{k r} = W[k r tstep]
  cxg<T> v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    v[r] = src[j + r * per_row];
    if (r > 0) v[r] = cmul(v[r], conj_if(W[k * r * tstep], inv));
  }
  if (R == 2) {
    const cxg<T> a = v[0], b = v[1];
    v[0] = {a.x + b.x, a.y + b.y};
    v[1] = {a.x - b.x, a.y - b.y};
  } else {
    const cxg<T> s0 = {v[0].x + v[2].x, v[0].y + v[2].y}, d0 = {v[0].x - v[2].x, v[0].y - v[2].y};
    const cxg<T> s1 = {v[1].x + v[3].x, v[1].y + v[3].y}, d1 = {v[1].x - v[3].x, v[1].y - v[3].y};
    v[0] = {s0.x + s1.x, s0.y + s1.y};
    v[2] = {s0.x - s1.x, s0.y - s1.y};
    if (!inv) { // X1 = d0 - j d1, X3 = d0 + j d1
      v[1] = {d0.x + d1.y, d0.y - d1.x};
      v[3] = {d0.x - d1.y, d0.y + d1.x};
    } else {
      v[1] = {d0.x - d1.y, d0.y + d1.x};
      v[3] = {d0.x + d1.y, d0.y - d1.x};
    }
  }
  const long long base = (j - k) * R + k;
#pragma unroll
  for (int r = 0; r < R; ++r) dst[base + r * ns] = v[r];
}

// ---- glue kernels (all: one thread per element of a batch of rows) --------------------------------
// rows of n reals (row stride `stride`) -> rows of m complex points, zero padded; two real rows may be
// packed as real / imaginary part (rows 2 p and 2 p + 1 -> complex row p)
template <typename T>
__global__ void real_to_complex(const T *__restrict__ in, long long stride, long long n, long long rows,
                                cxg<T> *__restrict__ out, long long m, int pack_pairs) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long crow_count = pack_pairs ? (rows + 1) / 2 : rows;
  if (idx >= crow_count * m) return;
  const long long p = idx / m, i = idx - p * m;
  cxg<T> v = {(T)0, (T)0};
  if (i < n) {
    if (pack_pairs) {
      v.x = in[(2 * p) * stride + i];
      if (2 * p + 1 < rows) v.y = in[(2 * p + 1) * stride + i];
    } else {
      v.x = in[p * stride + i];
    }
  }
  out[idx] = v;
}
// complex rows of m points -> the first `keep` complex points of each row (e.g. n / 2 + 1 of an r2c)
template <typename T>
__global__ void complex_take(const cxg<T> *__restrict__ in, long long m, long long rows, cxg<T> *__restrict__ out,
                             long long keep) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * keep) return;
  const long long p = idx / keep, i = idx - p * keep;
  out[idx] = in[p * m + i];
}
// half spectra (n / 2 + 1 points per row, Hermitian) -> full complex rows of n points
template <typename T>
__global__ void hermitian_expand(const cxg<T> *__restrict__ in, long long n, long long rows, cxg<T> *__restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * n) return;
  const long long half = n / 2 + 1, p = idx / n, k = idx - p * n;
  cxg<T> v = k < half ? in[p * half + k] : conj_if(in[p * half + (n - k)], true);
  if (k == 0 || 2 * k == n) v.y = (T)0; // c2r ignores the imaginary parts of DC and Nyquist, like FFTW
  out[idx] = v;
}
template <typename T>
__global__ void complex_real_part(const cxg<T> *__restrict__ in, long long n, long long rows, T *__restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < rows * n) out[idx] = in[idx].x;
}

// chirp-z: a[i] = x[i] c[i] (i < n), 0 up to m.  conj_in: the backward transform is conj o forward o conj
template <typename T>
__global__ void chirp_in(const cxg<T> *__restrict__ x, long long n, long long rows, const cxg<T> *__restrict__ chirp,
                         cxg<T> *__restrict__ a, long long m, int conj_in) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * m) return;
  const long long p = idx / m, i = idx - p * m;
  cxg<T> v = {(T)0, (T)0};
  if (i < n) v = cmul(conj_if(x[p * n + i], conj_in != 0), chirp[i]);
  a[idx] = v;
}
template <typename T>
__global__ void pointwise_mul(cxg<T> *__restrict__ a, const cxg<T> *__restrict__ b, long long m, long long rows) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * m) return;
  a[idx] = cmul(a[idx], b[idx % m]);
}
// X[k] = c[k] y[k] / m, k < n
template <typename T>
__global__ void chirp_out(const cxg<T> *__restrict__ y, long long m, long long n, long long rows,
                          const cxg<T> *__restrict__ chirp, cxg<T> *__restrict__ out, T inv_m, int conj_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * n) return;
  const long long p = idx / n, k = idx - p * n;
  cxg<T> v = cmul(y[p * m + k], chirp[k]);
  v.x *= inv_m;
  v.y *= inv_m;
  out[idx] = conj_if(v, conj_out != 0);
}

// Hilbert transformer on full spectra: X[k] *= -j sgn(k) / n (DC and Nyquist to zero)
// (/root/reference/include/fftconv/hilbert.hpp:39-51, :57)
template <typename T> __global__ void hilbert_multiplier(cxg<T> *__restrict__ x, long long n, long long rows) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * n) return;
  const long long k = idx % n;
  const cxg<T> v = x[idx];
  const T g = (T)1 / (T)n;
  cxg<T> r = {(T)0, (T)0};
  if (k != 0 && 2 * k != n) {
    if (2 * k < n) r = {v.y * g, -v.x * g};  // -j v / n
    else r = {-v.y * g, v.x * g};            // +j v / n
  }
  x[idx] = r;
}
// env = sqrt(x^2 + h^2): complex row p holds the Hilbert transforms of real rows 2 p (re) and 2 p + 1 (im)
// with the envelope post-processing of fftconv_cuda_envelope_* : keep every dec-th sample, optional log map
template <typename T>
__global__ void envelope_from_pairs(const T *__restrict__ in, long long stride, long long n, long long rows,
                                    const cxg<T> *__restrict__ h, T *__restrict__ out, long long out_stride,
                                    long long dec, int log_map, T la, T lb) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * n) return;
  const long long r = idx / n, i = idx - r * n;
  if (dec > 1 && i % dec != 0) return;
  const cxg<T> hv = h[(r / 2) * n + i];
  const T x = in[r * stride + i], hh = (r & 1) ? hv.y : hv.x;
  T e = sqrt(x * x + hh * hh);
  if (log_map) {
    e = la * log2(e) + lb;
    e = e < (T)0 ? (T)0 : (e > (T)1 ? (T)1 : e);
  }
  out[r * out_stride + (dec > 1 ? i / dec : i)] = e;
}

} // namespace gfft
} // namespace fc
