// long_fft_kernels.cuh -- one FFT longer than a CTA's shared memory: N_big = 16 N_sub.
//
// convolve_fftw (/root/reference/include/fftconv/fftconv.hpp:453-461, :338-380) uses ONE
// transform of length n + K - 1 however long the row is, and hilbert
// (include/fftconv/hilbert.hpp:16-66) one of length n.  For rows beyond the
// single-CTA engine (16384 float / 8192 double points) the transform is split once
// more, decimation in frequency, through a scratch array in global memory:
//   outer_fwd_kernel   16-point DFTs over the slowest index (stride N_sub) in
//                      registers, twiddle W_Nbig^{p s}, scratch[pair][s][p]
//   inner_kernel       the 16 sub-transforms of length N_sub per row pair on the
//                      register-blocked engine (fast_fft.cuh): forward passes,
//                      spectrum multiply, inverse passes, in place in scratch
//   outer_inv_kernel   conjugate twiddle, inverse 16-point DFTs, outputs
// Two real rows ride in the real and imaginary parts, as everywhere else.  The
// spectrum table is laid out [s][fast-engine table order]; frequency of entry
// (s, idx) is s + 16 * freq_sub(idx).
#pragma once

#include "envelope_epilogue.cuh"
#include "fast_fft.cuh"

namespace fc {
namespace longfft {

using fast::cx;

template <typename T> struct LongArgs {
  const T *in;
  T *out;
  long long batch, n, in_stride, out_stride, out_len;
  int pad;            // Same-mode offset (0 for Full / Hilbert)
  int log2sub;        // log2 N_sub
  long long pair0;    // first row pair handled by this launch
  long long pairs;    // row pairs in this launch
  cx<T> *scratch;     // pairs * 16 * N_sub
  const cx<T> *TWbig; // 15 * N_sub: [(s-1) N_sub + p] = W_Nbig^{p s}
  int hilbert;        // outer_inv: write the envelope instead of the convolution
  EnvEpilogue<T> epi; // hilbert only
  // Chirp-z (Bluestein) Hilbert envelope of a line whose length n is not a power of two, N_big =
  // M >= 2 n - 1: two convolutions with the chirp filter (generic_kernels.cuh, hilbert_bluestein_kernel,
  // is the single-CTA form of the same recipe):
  //   bluestein = 1: load (x0 + i x1) chirp, store u = chirp conj(Hilbert(chirp * y)) into cbuf
  //   bluestein = 2: load cbuf, store the envelope from conj(chirp * y) / n
  int bluestein;
  const cx<T> *chirp; // n: exp(-pi i j^2 / n)
  cx<T> *cbuf;        // pairs * N_big
};

template <typename T> __global__ void outer_fwd_kernel(const LongArgs<T> a) {
  const int nsub = 1 << a.log2sub;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nsub) return;
  const long long pair = blockIdx.y;
  const long long r0 = (a.pair0 + pair) * 2, r1 = r0 + 1;
  const bool has1 = r1 < a.batch;
  const T *x0 = a.in + r0 * a.in_stride;
  const T *x1 = a.in + (has1 ? r1 : r0) * a.in_stride;
  cx<T> v[16];
  if (a.bluestein == 2) {
    const cx<T> *src = a.cbuf + pair * 16 * nsub + p;
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = src[(long long)q * nsub];
  } else {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const long long i = p + (long long)q * nsub;
      v[q].x = i < a.n ? x0[i] : (T)0;
      v[q].y = (has1 && i < a.n) ? x1[i] : (T)0;
      if (a.bluestein == 1 && i < a.n) v[q] = fast::cmul(v[q], a.chirp[i]);
    }
  }
  fast::dft16<-1>(v);
  cx<T> *dst = a.scratch + pair * 16 * nsub + p;
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    cx<T> y = v[s];
    if (s > 0) y = fast::cmul(y, a.TWbig[(long long)(s - 1) * nsub + p]);
    dst[(long long)s * nsub] = y;
  }
}

template <typename T> __global__ void outer_inv_kernel(const LongArgs<T> a) {
  const int nsub = 1 << a.log2sub;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nsub) return;
  const long long pair = blockIdx.y;
  const long long r0 = (a.pair0 + pair) * 2, r1 = r0 + 1;
  const bool has1 = r1 < a.batch;
  const cx<T> *src = a.scratch + pair * 16 * nsub + p;
  cx<T> v[16];
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    cx<T> y = src[(long long)s * nsub];
    if (s > 0) y = fast::cmulc(y, a.TWbig[(long long)(s - 1) * nsub + p]);
    v[s] = y;
  }
  fast::dft16<+1>(v);
  T *o0 = a.out + r0 * a.out_stride;
  T *o1 = a.out + (has1 ? r1 : r0) * a.out_stride;
  if (a.bluestein == 1) {
    // X[k] = chirp[k] v[k]; Z = -j X (0 < k < n/2), +j X (k > n/2), 0 at DC / Nyquist
    // (/root/reference/include/fftconv/hilbert.hpp:39-51); the inverse DFT is conj(DFT(conj(Z))) / n
    cx<T> *dst = a.cbuf + pair * 16 * nsub + p;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const long long k = p + (long long)q * nsub;
      cx<T> u = {(T)0, (T)0};
      if (k < a.n && k != 0 && 2 * k != a.n) {
        const cx<T> c = a.chirp[k];
        const cx<T> X = fast::cmul(v[q], c);
        cx<T> Zc; // conj(Z)
        if (2 * k < a.n) { Zc.x = X.y; Zc.y = X.x; }
        else             { Zc.x = -X.y; Zc.y = -X.x; }
        u = fast::cmul(Zc, c);
      }
      dst[(long long)q * nsub] = u;
    }
  } else if (a.bluestein == 2) {
    const T *x0 = a.in + r0 * a.in_stride;
    const T *x1 = a.in + (has1 ? r1 : r0) * a.in_stride;
    const T inv_n = (T)(1.0 / (double)a.n);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const long long m = p + (long long)q * nsub;
      if (m < a.n) {
        const cx<T> w = fast::cmul(v[q], a.chirp[m]); // conj(h) * n
        const T h0 = w.x * inv_n, h1 = -w.y * inv_n;
        const T u = x0[m];
        env_store(a.epi, o0, m, sqrt(u * u + h0 * h0));
        if (has1) {
          const T t = x1[m];
          env_store(a.epi, o1, m, sqrt(t * t + h1 * h1));
        }
      }
    }
  } else if (a.hilbert) {
    const T *x0 = a.in + r0 * a.in_stride;
    const T *x1 = a.in + (has1 ? r1 : r0) * a.in_stride;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const long long i = p + (long long)q * nsub;
      if (i < a.n) {
        const T u = x0[i];
        env_store(a.epi, o0, i, sqrt(u * u + v[q].x * v[q].x));
        if (has1) {
          const T w = x1[i];
          env_store(a.epi, o1, i, sqrt(w * w + v[q].y * v[q].y));
        }
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const long long o = p + (long long)q * nsub - a.pad;
      if (o >= 0 && o < a.out_len) {
        o0[o] = v[q].x;
        if (has1) o1[o] = v[q].y;
      }
    }
  }
}

struct CtaSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
  __device__ __forceinline__ void warp() const { __syncwarp(); } // warp-local exchange (fast_fft.cuh, pass_sync)
  template <int LM> __device__ __forceinline__ void sub(int) const { __syncthreads(); }
};

// one sub-transform per FFT slot, FPC slots per CTA; H = Hbig + (sub index) * N_sub
template <typename T, int LOG2N, int FPC>
__global__ void __launch_bounds__(fast::Plan<LOG2N>::THREADS *FPC)
    inner_kernel(cx<T> *scratch, long long n_subs, const cx<T> *W, const cx<T> *Hbig) {
  using P = fast::Plan<LOG2N>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int f = threadIdx.x / P::THREADS, tid = threadIdx.x % P::THREADS;
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw) + (size_t)f * P::SMEM_ELEMS;
  long long g = (long long)blockIdx.x * FPC + f;
  const bool on = g < n_subs;
  if (!on) g = n_subs - 1; // keep the barriers uniform; results of the duplicate are dropped
  cx<T> *x = scratch + g * P::N;
  const cx<T> *H = Hbig + (g & 15) * P::N;
  cx<T> v[16];
#pragma unroll
  for (int s = 0; s < 16; ++s) v[s] = x[fast::pass0_index<LOG2N>(tid, s)];
  CtaSync sync;
  fast::conv_core<LOG2N>(tid, v, buf, W, H, sync);
  if (on) {
#pragma unroll
    for (int s = 0; s < 16; ++s) x[fast::pass0_index<LOG2N>(tid, s)] = v[s];
  }
}

} // namespace longfft
} // namespace fc
