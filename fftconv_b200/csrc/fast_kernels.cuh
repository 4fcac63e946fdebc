// fast_kernels.cuh -- overlap-add / single-FFT convolution and the Hilbert envelope
// on the register-blocked engine of fast_fft.cuh, for FFT sizes 512 .. 16384, float
// and double.  Same stream / warm-up / tail scheme as oa_n2048_group.cuh, two real
// streams per complex FFT (real and imaginary part), FPC independent FFTs per CTA.
//   oaconvolve_fftw  /root/reference/include/fftconv/fftconv.hpp:474-480, :382-437
//   convolve_fftw    /root/reference/include/fftconv/fftconv.hpp:453-461, :338-380
//   hilbert          /root/reference/include/fftconv/hilbert.hpp:16-66
#pragma once

#include "fast_fft.cuh"
#include "generic_kernels.cuh"

namespace fc {
namespace fast {

// occupancy target: ~768 threads per SM (<= 85 registers) for float, ~512 (<= 128) for double
template <typename T, int CTA_THREADS> constexpr int min_blocks() {
  constexpr int target = sizeof(T) == 4 ? 768 : 512;
  return CTA_THREADS >= target ? 1 : target / CTA_THREADS;
}

struct CtaSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

template <typename T, int LOG2N, int FPC>
__global__ void __launch_bounds__(Plan<LOG2N>::THREADS *FPC, min_blocks<T, Plan<LOG2N>::THREADS * FPC>())
    oa_fast_kernel(const generic::OaArgs<T> p) {
  using P = Plan<LOG2N>;
  constexpr int N = P::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int f = threadIdx.x / P::THREADS, tid = threadIdx.x % P::THREADS;
  const int tail_len = p.klen - 1;
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw) + (size_t)f * P::SMEM_ELEMS;
  cx<T> *tails = reinterpret_cast<cx<T> *>(smem_raw) + (size_t)FPC * P::SMEM_ELEMS + (size_t)f * 2 * tail_len;
  const cx<T> *W = reinterpret_cast<const cx<T> *>(p.W);
  const cx<T> *H = reinterpret_cast<const cx<T> *>(p.H);

  // CTA-uniform loop bounds over all FPC FFTs (barriers are CTA-wide)
  long long max_cnt = 0;
  bool warm_needed = false;
  const T *src[2] = {p.in, p.in};
  T *dst[2] = {p.out, p.out};
  long long seg0[2] = {0, 0}, cnt[2] = {0, 0};
#pragma unroll
  for (int ff = 0; ff < FPC; ++ff)
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      const long long sid = ((long long)blockIdx.x * FPC + ff) * 2 + l;
      const bool active = sid < p.n_streams;
      const long long row = active ? sid / p.chunks : 0;
      const long long chunk = active ? sid - row * p.chunks : 0;
      const long long s0 = chunk * p.segs_per_chunk;
      const long long rem = p.segs - s0;
      const long long c = active ? (rem < p.segs_per_chunk ? rem : p.segs_per_chunk) : 0;
      if (c > max_cnt) max_cnt = c;
      warm_needed = warm_needed || (active && s0 > 0);
      if (ff == f) {
        seg0[l] = s0;
        cnt[l] = c;
        src[l] = p.in + row * p.in_stride;
        dst[l] = p.out + row * p.out_stride;
      }
    }

  CtaSync sync;
  int parity = 0;
  bool add_tail = false;
  for (long long j = warm_needed ? -1 : 0; j < max_cnt; ++j) {
    const bool warm = j < 0;
    long long pos[2];
    int len[2], limit[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      const long long seg = seg0[l] + j;
      const bool on = seg >= 0 && j < cnt[l];
      pos[l] = seg * p.step;
      long long ll = on ? p.n - pos[l] : 0;
      if (ll > p.step) ll = p.step;
      len[l] = (int)ll;
      long long lim = p.step;
      if (seg == p.segs - 1) {
        lim = p.out_len + p.pad - pos[l];
        if (lim > N) lim = N;
      }
      limit[l] = (on && !warm) ? (int)lim : 0;
    }
    cx<T> v[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      const int i = pass0_index<LOG2N>(tid, s);
      v[s].x = i < len[0] ? src[0][pos[0] + i] : (T)0;
      v[s].y = i < len[1] ? src[1][pos[1] + i] : (T)0;
    }
    conv_core<LOG2N>(tid, v, buf, W, H, sync);
    cx<T> *tail_prev = tails + (parity ^ 1) * tail_len;
    cx<T> *tail_next = tails + parity * tail_len;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      const int i = pass0_index<LOG2N>(tid, s);
      cx<T> y = v[s];
      if (add_tail && i < tail_len) y = y + tail_prev[i];
      if (i >= p.step) tail_next[i - p.step] = y;
      const long long o0 = pos[0] - p.pad + i, o1 = pos[1] - p.pad + i;
      if (i < limit[0] && o0 >= 0) dst[0][o0] = y.x;
      if (i < limit[1] && o1 >= 0) dst[1][o1] = y.y;
    }
    __syncthreads(); // tails and the exchange buffer are reused by the next segment
    parity ^= 1;
    add_tail = true;
  }
}

// Hilbert envelope of rows of length N (power of two): rows 2b, 2b+1 per FFT.
// G (position order) = -j sgn(f) / N with DC and Nyquist zeroed (hilbert.hpp:39-51).
template <typename T, int LOG2N, int FPC>
__global__ void __launch_bounds__(Plan<LOG2N>::THREADS *FPC, min_blocks<T, Plan<LOG2N>::THREADS * FPC>())
    hilbert_fast_kernel(const generic::HilbertArgs<T> p) {
  using P = Plan<LOG2N>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int f = threadIdx.x / P::THREADS, tid = threadIdx.x % P::THREADS;
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw) + (size_t)f * P::SMEM_ELEMS;
  const cx<T> *W = reinterpret_cast<const cx<T> *>(p.W);
  const cx<T> *G = reinterpret_cast<const cx<T> *>(p.G);
  const long long r0 = ((long long)blockIdx.x * FPC + f) * 2, r1 = r0 + 1;
  const bool has0 = r0 < p.batch, has1 = r1 < p.batch;
  const T *x0 = p.in + (has0 ? r0 : 0) * p.in_stride;
  const T *x1 = p.in + (has1 ? r1 : 0) * p.in_stride;
  // float keeps the inputs in registers for the envelope; double re-reads them
  // (16 more complex doubles would not fit 128 registers per thread)
  constexpr bool kKeep = sizeof(T) == 4;
  cx<T> v[16], x[kKeep ? 16 : 1];
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    const int i = pass0_index<LOG2N>(tid, s);
    v[s].x = has0 ? x0[i] : (T)0;
    v[s].y = has1 ? x1[i] : (T)0;
    if constexpr (kKeep) x[s] = v[s];
  }
  CtaSync sync;
  conv_core<LOG2N>(tid, v, buf, W, G, sync);
  T *e0 = p.out + (has0 ? r0 : 0) * p.out_stride;
  T *e1 = p.out + (has1 ? r1 : 0) * p.out_stride;
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    const int i = pass0_index<LOG2N>(tid, s);
    T a, b;
    if constexpr (kKeep) {
      a = x[s].x;
      b = x[s].y;
    } else {
      a = has0 ? x0[i] : (T)0;
      b = has1 ? x1[i] : (T)0;
    }
    if (has0) e0[i] = sqrt(a * a + v[s].x * v[s].x);
    if (has1) e1[i] = sqrt(b * b + v[s].y * v[s].y);
  }
}

} // namespace fast
} // namespace fc
