// generic_fft.cuh -- shared-memory FFT building blocks for every case the
// specialised kernels do not cover: any power-of-two segment size, float or
// double, overlap-add / single-FFT convolution and the Hilbert envelope.
//
// One CTA owns one complex FFT of length N held in shared memory; its real and
// imaginary parts carry two independent real streams (two rows, or two chunks
// of a row) -- real taps / a real Hilbert kernel commute with that packing, so
// no r2c / c2r untangle pass is needed.  The forward transform is an in-place
// decimation-in-frequency (natural order in, digit-reversed order out), the
// spectrum table is stored in the same digit-reversed order, and the inverse is
// the mirrored decimation-in-time (digit-reversed in, natural out): no
// reordering pass either.  Twiddles come from a table evaluated in long double
// on the host.
#pragma once

#include <cuda_runtime.h>

namespace fc {
namespace generic {

template <typename T> struct cx {
  T x, y;
};

template <typename T> __device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> w) {
  cx<T> r;
  r.x = a.x * w.x - a.y * w.y;
  r.y = a.x * w.y + a.y * w.x;
  return r;
}
template <typename T> __device__ __forceinline__ cx<T> cmulc(cx<T> a, cx<T> w) { // a * conj(w)
  cx<T> r;
  r.x = a.x * w.x + a.y * w.y;
  r.y = a.y * w.x - a.x * w.y;
  return r;
}

// position (in the DIF output order) -> frequency index
__host__ __device__ inline unsigned dif_freq_of_pos(unsigned pos, unsigned n) {
  unsigned k = 0, mult = 1;
  while (n >= 4) {
    const unsigned quarter = n >> 2;
    const unsigned q = pos / quarter;
    pos -= q * quarter;
    k += mult * q;
    mult <<= 2;
    n = quarter;
  }
  if (n == 2) k += mult * pos;
  return k;
}

// In-place forward DIF over buf[0..N), N = 1 << log2n, W[k] = exp(-2 pi i k / N).
// Ends with a barrier.
template <typename T>
__device__ void fft_dif(cx<T> *buf, const cx<T> *__restrict__ W, int log2n) {
  const int N = 1 << log2n;
  int lm = log2n; // log2 of current sub-transform length m
  for (; lm >= 2; lm -= 2) {
    const int lq = lm - 2;           // log2(m / 4)
    const int qmask = (1 << lq) - 1;
    const int tshift = log2n - lm;   // twiddle stride N / m
    for (int b = threadIdx.x; b < (N >> 2); b += blockDim.x) {
      const int pq = b & qmask;
      const int base = ((b >> lq) << lm) + pq;
      const int quarter = 1 << lq;
      const cx<T> a = buf[base], bb = buf[base + quarter], c = buf[base + 2 * quarter],
                  d = buf[base + 3 * quarter];
      const cx<T> s0 = {a.x + c.x, a.y + c.y}, d0 = {a.x - c.x, a.y - c.y};
      const cx<T> s1 = {bb.x + d.x, bb.y + d.y}, d1 = {bb.x - d.x, bb.y - d.y};
      const cx<T> y0 = {s0.x + s1.x, s0.y + s1.y};
      cx<T> y2 = {s0.x - s1.x, s0.y - s1.y};
      cx<T> y1 = {d0.x + d1.y, d0.y - d1.x}; // d0 - j d1
      cx<T> y3 = {d0.x - d1.y, d0.y + d1.x}; // d0 + j d1
      if (pq != 0) {
        const int e = pq << tshift;
        y1 = cmul(y1, W[e]);
        y2 = cmul(y2, W[2 * e]);
        y3 = cmul(y3, W[3 * e]);
      }
      buf[base] = y0;
      buf[base + quarter] = y1;
      buf[base + 2 * quarter] = y2;
      buf[base + 3 * quarter] = y3;
    }
    __syncthreads();
  }
  if (lm == 1) {
    for (int b = threadIdx.x; b < (N >> 1); b += blockDim.x) {
      const cx<T> a = buf[2 * b], bb = buf[2 * b + 1];
      buf[2 * b] = {a.x + bb.x, a.y + bb.y};
      buf[2 * b + 1] = {a.x - bb.x, a.y - bb.y};
    }
    __syncthreads();
  }
}

// In-place inverse (unnormalised) DIT, exact mirror of fft_dif.  Ends with a barrier.
template <typename T>
__device__ void fft_dit_inv(cx<T> *buf, const cx<T> *__restrict__ W, int log2n) {
  const int N = 1 << log2n;
  int lm = 2;
  if (log2n & 1) {
    for (int b = threadIdx.x; b < (N >> 1); b += blockDim.x) {
      const cx<T> a = buf[2 * b], bb = buf[2 * b + 1];
      buf[2 * b] = {a.x + bb.x, a.y + bb.y};
      buf[2 * b + 1] = {a.x - bb.x, a.y - bb.y};
    }
    __syncthreads();
    lm = 3;
  }
  for (; lm <= log2n; lm += 2) {
    const int lq = lm - 2;
    const int qmask = (1 << lq) - 1;
    const int tshift = log2n - lm;
    for (int b = threadIdx.x; b < (N >> 2); b += blockDim.x) {
      const int pq = b & qmask;
      const int base = ((b >> lq) << lm) + pq;
      const int quarter = 1 << lq;
      const cx<T> y0 = buf[base];
      cx<T> t1 = buf[base + quarter], t2 = buf[base + 2 * quarter], t3 = buf[base + 3 * quarter];
      if (pq != 0) {
        const int e = pq << tshift;
        t1 = cmulc(t1, W[e]);
        t2 = cmulc(t2, W[2 * e]);
        t3 = cmulc(t3, W[3 * e]);
      }
      const cx<T> s0 = {y0.x + t2.x, y0.y + t2.y}, d0 = {y0.x - t2.x, y0.y - t2.y};
      const cx<T> s1 = {t1.x + t3.x, t1.y + t3.y}, d1 = {t1.x - t3.x, t1.y - t3.y};
      buf[base] = {s0.x + s1.x, s0.y + s1.y};
      buf[base + quarter] = {d0.x - d1.y, d0.y + d1.x};     // d0 + j d1
      buf[base + 2 * quarter] = {s0.x - s1.x, s0.y - s1.y};
      buf[base + 3 * quarter] = {d0.x + d1.y, d0.y - d1.x}; // d0 - j d1
    }
    __syncthreads();
  }
}

// buf[i] *= H[i]; ends with a barrier.
template <typename T>
__device__ void spectrum_mul(cx<T> *buf, const cx<T> *__restrict__ H, int N) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) buf[i] = cmul(buf[i], H[i]);
  __syncthreads();
}

} // namespace generic
} // namespace fc
