// generic_kernels.cuh -- fallback CUDA kernels (any power-of-two FFT size, float
// or double) for the three reference entry points:
//   oaconvolve_fftw  /root/reference/include/fftconv/fftconv.hpp:474-480, :382-437
//   convolve_fftw    /root/reference/include/fftconv/fftconv.hpp:453-461, :338-380
//   hilbert          /root/reference/include/fftconv/hilbert.hpp:16-66
// The f32 / N = 2048 overlap-add case has its own kernel (oa_n2048_kernel.cuh).
#pragma once

#include "envelope_epilogue.cuh"
#include "fast_fft.cuh"
#include "generic_fft.cuh"

namespace fc {
namespace generic {

using fc::EnvEpilogue;
using fc::env_store;

// ---------------------------------------------------------------------------
// Spectrum of the taps, evaluated directly in double:
//   H[pos] = scale * sum_j taps[j] exp(-2 pi i j k(pos) / N)
// k(pos) follows the layout of the consuming kernel.  wd[e] = exp(-2 pi i e / N)
// in double.  O(N K) work -- K <= a few thousand taps by construction.  (j k fits
// 32 bits: j < K <= N / 2, k < N <= 2^15.)
// ---------------------------------------------------------------------------
enum SpectrumLayout { LAYOUT_DIF = 0, LAYOUT_N2048 = 1, LAYOUT_FAST = 2, LAYOUT_LONG = 3, LAYOUT_W1K = 4 };

__device__ __forceinline__ unsigned n2048_freq_of_index(unsigned idx) {
  // idx = slot * 128 + t (see fc::n2048::p2_freq, p2_col)
  const unsigned s = idx >> 7, t = idx & 127;
  const unsigned h = s >> 3, k1b = s & 7, c = 64 * (t >> 5) + 32 * h + (t & 31);
  return 16 * (c & 15) + 256 * k1b + (c >> 4);
}

// One warp per spectrum bin; the lanes split the taps and reduce with shuffles.
constexpr int kSpectrumWarps = 8;
template <typename T>
__global__ void spectrum_kernel(const T *__restrict__ taps, int klen, int N,
                                const double2 *__restrict__ wd, int layout, double scale,
                                cx<T> *__restrict__ H) {
  const int lane = threadIdx.x & 31;
  const int pos = blockIdx.x * kSpectrumWarps + (threadIdx.x >> 5);
  if (pos >= N) return;
  unsigned k;
  if (layout == LAYOUT_N2048) k = n2048_freq_of_index(pos);
  else if (layout == LAYOUT_W1K) // entry e = 2 (j * 32 + k2) + half (fc::w1k::freq_of_index)
    k = 32u * (((unsigned)pos >> 6) + 16u * ((unsigned)pos & 1u)) + (((unsigned)pos >> 1) & 31u);
  else if (layout == LAYOUT_FAST) k = fc::fast::freq_of_pos(fc::fast::pos_of_table_index(pos, 31 - __clz(N)), 31 - __clz(N));
  else if (layout == LAYOUT_LONG) {
    // N = 16 N_sub (long_fft_kernels.cuh): entry (s, idx_sub) holds frequency s + 16 freq_sub(idx_sub)
    const int l2s = 31 - __clz(N) - 4;
    const unsigned sub = (unsigned)pos & ((1u << l2s) - 1);
    k = ((unsigned)pos >> l2s) + 16u * fc::fast::freq_of_pos(fc::fast::pos_of_table_index(sub, l2s), l2s);
  } else k = dif_freq_of_pos(pos, N);
  const unsigned mask = (unsigned)(N - 1);
  double re = 0.0, im = 0.0;
  for (int j = lane; j < klen; j += 32) {
    const double2 w = wd[((unsigned)j * k) & mask]; // (j k) mod N, N a power of two
    const double v = (double)taps[j];
    re += v * w.x;
    im += v * w.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    re += __shfl_xor_sync(0xffffffffu, re, o);
    im += __shfl_xor_sync(0xffffffffu, im, o);
  }
  if (lane == 0) {
    H[pos].x = (T)(re * scale);
    H[pos].y = (T)(im * scale);
  }
}

// ---------------------------------------------------------------------------
// Overlap-add (and, with one segment covering the whole row, single-FFT
// convolution).  CTA b walks streams 2b (real part) and 2b+1 (imaginary part).
// ---------------------------------------------------------------------------
template <typename T> struct OaArgs {
  const T *in;
  T *out;
  long long batch, n, in_stride, out_stride, out_len;
  int klen, log2n, step, pad;
  long long segs, chunks, segs_per_chunk, n_streams;
  const cx<T> *W; // N twiddles
  const cx<T> *H; // N spectrum values, DIF order, 1/N folded in
};

template <typename T> __global__ void oa_generic_kernel(const OaArgs<T> p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int N = 1 << p.log2n;
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *tails = buf + N; // 2 x (klen - 1)
  const int tail_len = p.klen - 1;

  const T *src[2];
  T *dst[2];
  long long seg0[2], cnt[2];
  bool active[2];
  long long max_cnt = 0;
  bool warm_needed = false;
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const long long sid = (long long)blockIdx.x * 2 + l;
    active[l] = sid < p.n_streams;
    const long long row = active[l] ? sid / p.chunks : 0;
    const long long chunk = active[l] ? sid - row * p.chunks : 0;
    seg0[l] = chunk * p.segs_per_chunk;
    const long long rem = p.segs - seg0[l];
    cnt[l] = active[l] ? (rem < p.segs_per_chunk ? rem : p.segs_per_chunk) : 0;
    src[l] = p.in + row * p.in_stride;
    dst[l] = p.out + row * p.out_stride;
    if (cnt[l] > max_cnt) max_cnt = cnt[l];
    warm_needed = warm_needed || (active[l] && seg0[l] > 0);
  }

  int parity = 0;
  bool add_tail = false;
  for (long long j = warm_needed ? -1 : 0; j < max_cnt; ++j) {
    const bool warm = j < 0;
    long long pos[2];
    int len[2];
    bool on[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      const long long seg = seg0[l] + j;
      on[l] = active[l] && seg >= 0 && j < cnt[l];
      pos[l] = seg * p.step;
      long long ll = on[l] ? p.n - pos[l] : 0;
      if (ll > p.step) ll = p.step;
      len[l] = (int)ll;
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      cx<T> v;
      v.x = i < len[0] ? src[0][pos[0] + i] : (T)0;
      v.y = i < len[1] ? src[1][pos[1] + i] : (T)0;
      buf[i] = v;
    }
    __syncthreads();
    fft_dif(buf, p.W, p.log2n);
    spectrum_mul(buf, p.H, N);
    fft_dit_inv(buf, p.W, p.log2n);

    cx<T> *tail_prev = tails + (parity ^ 1) * tail_len;
    cx<T> *tail_next = tails + parity * tail_len;
    int limit[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      long long lim = p.step;
      if (seg0[l] + j == p.segs - 1) {
        lim = p.out_len + p.pad - pos[l];
        if (lim > N) lim = N;
      }
      limit[l] = (on[l] && !warm) ? (int)lim : 0;
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      cx<T> y = buf[i];
      if (add_tail && i < tail_len) {
        y.x += tail_prev[i].x;
        y.y += tail_prev[i].y;
      }
      if (i >= p.step) tail_next[i - p.step] = y;
      const long long o0 = pos[0] - p.pad + i, o1 = pos[1] - p.pad + i;
      if (i < limit[0] && o0 >= 0) dst[0][o0] = y.x;
      if (i < limit[1] && o1 >= 0) dst[1][o1] = y.y;
    }
    __syncthreads();
    parity ^= 1;
    add_tail = true;
  }
}

// ---------------------------------------------------------------------------
// Hilbert envelope, power-of-two length n = N: rows 2b and 2b+1 in one complex
// FFT; G (DIF order) holds the Hilbert multiplier -j sgn(f) / n with DC and
// Nyquist zeroed (hilbert.hpp:39-51), so buf becomes h0 + i h1; env = sqrt(x^2 + h^2).
// ---------------------------------------------------------------------------
template <typename T> struct HilbertArgs {
  const T *in;
  T *out;
  long long batch, n, in_stride, out_stride;
  int log2n;
  const cx<T> *W;
  const cx<T> *G;
  EnvEpilogue<T> epi;
};

template <typename T> __global__ void hilbert_pow2_kernel(const HilbertArgs<T> p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int N = 1 << p.log2n;
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw);
  const long long r0 = (long long)blockIdx.x * 2, r1 = r0 + 1;
  const bool has1 = r1 < p.batch;
  const T *x0 = p.in + r0 * p.in_stride;
  const T *x1 = p.in + (has1 ? r1 : r0) * p.in_stride;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    cx<T> v;
    v.x = x0[i];
    v.y = has1 ? x1[i] : (T)0;
    buf[i] = v;
  }
  __syncthreads();
  fft_dif(buf, p.W, p.log2n);
  spectrum_mul(buf, p.G, N);
  fft_dit_inv(buf, p.W, p.log2n);
  T *e0 = p.out + r0 * p.out_stride;
  T *e1 = p.out + (has1 ? r1 : r0) * p.out_stride;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const cx<T> h = buf[i];
    const T a = x0[i];
    env_store(p.epi, e0, i, sqrt(a * a + h.x * h.x));
    if (has1) {
      const T b = x1[i];
      env_store(p.epi, e1, i, sqrt(b * b + h.y * h.y));
    }
  }
}

// ---------------------------------------------------------------------------
// Hilbert envelope, any length n, by Bluestein's chirp-z identity on top of the
// power-of-two engine (M >= 2 n - 1):
//   DFT_n(x)[k] = c[k] * ((x c) (*)_M conj-chirp)[k],   c[j] = exp(-pi i j^2 / n)
// Rows 2b and 2b+1 again ride in the real and imaginary parts (the Hilbert
// transformer is a real filter).  Tables: chirp[n]; B[M] = FFT_M(wrapped
// conj(chirp)) / M in DIF order.  The inverse DFT is conj(DFT(conj(.))) / n.
// (/root/reference/include/fftconv/hilbert.hpp:16-66 accepts any n > 0.)
// ---------------------------------------------------------------------------
template <typename T> struct HilbertBluesteinArgs {
  const T *in;
  T *out;
  long long batch, n, in_stride, out_stride;
  int log2m;
  const cx<T> *W;     // M twiddles
  const cx<T> *B;     // M, DIF order, 1/M folded in
  const cx<T> *chirp; // n
  EnvEpilogue<T> epi;
};

template <typename T> __global__ void hilbert_bluestein_kernel(const HilbertBluesteinArgs<T> p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int M = 1 << p.log2m;
  const int n = (int)p.n;
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw);
  const long long r0 = (long long)blockIdx.x * 2, r1 = r0 + 1;
  const bool has1 = r1 < p.batch;
  const T *x0 = p.in + r0 * p.in_stride;
  const T *x1 = p.in + (has1 ? r1 : r0) * p.in_stride;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    cx<T> v = {(T)0, (T)0};
    if (i < n) {
      v.x = x0[i];
      v.y = has1 ? x1[i] : (T)0;
      v = cmul(v, p.chirp[i]);
    }
    buf[i] = v;
  }
  __syncthreads();
  fft_dif(buf, p.W, p.log2m);
  spectrum_mul(buf, p.B, M);
  fft_dit_inv(buf, p.W, p.log2m);
  // X[k] = c[k] buf[k]; Hilbert multiplier (hilbert.hpp:39-51); set up the inverse
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    cx<T> u = {(T)0, (T)0};
    if (k < n && k != 0 && 2 * k != n) {
      const cx<T> c = p.chirp[k];
      const cx<T> X = cmul(buf[k], c);
      // Z = -j X for 0 < k < n/2, +j X for k > n/2
      cx<T> Z;
      if (2 * k < n) { Z.x = X.y; Z.y = -X.x; }
      else           { Z.x = -X.y; Z.y = X.x; }
      const cx<T> Zc = {Z.x, -Z.y};
      u = cmul(Zc, c);
    }
    buf[k] = u;
  }
  __syncthreads();
  fft_dif(buf, p.W, p.log2m);
  spectrum_mul(buf, p.B, M);
  fft_dit_inv(buf, p.W, p.log2m);
  T *e0 = p.out + r0 * p.out_stride;
  T *e1 = p.out + (has1 ? r1 : r0) * p.out_stride;
  const T inv_n = (T)(1.0 / (double)n);
  for (int m = threadIdx.x; m < n; m += blockDim.x) {
    const cx<T> w = cmul(buf[m], p.chirp[m]); // conj(h) * n
    const T h0 = w.x * inv_n, h1 = -w.y * inv_n;
    const T a = x0[m];
    env_store(p.epi, e0, m, sqrt(a * a + h0 * h0));
    if (has1) {
      const T b = x1[m];
      env_store(p.epi, e1, m, sqrt(b * b + h1 * h1));
    }
  }
}

// dst[i] += src[i]  (tail hand-over between neighbouring chunks of a long signal)
template <typename T> __global__ void add_inplace_kernel(T *dst, const T *src, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

// Taps too long for one segment FFT are split into blocks k_p = k[p Kb : (p + 1) Kb]:
//   a * k = sum_p shift(a * k_p, p Kb)
// `part` holds rows of part_len = n + len(k_p) - 1 samples (the Full convolution with block p);
// out[o] (+)= part[o + pad - shift], summed in block order -> deterministic.  The first block
// assigns (and zero-fills what it does not reach), like the reference's fill(out, 0)
// (/root/reference/include/fftconv/fftconv.hpp:385).
template <typename T>
__global__ void partition_accumulate_kernel(const T *__restrict__ part, long long rows, long long part_len,
                                            T *__restrict__ out, long long out_len, long long out_stride,
                                            long long shift_minus_pad, int first) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * out_len) return;
  const long long r = idx / out_len, o = idx - r * out_len;
  const long long i = o - shift_minus_pad;
  const T v = (i >= 0 && i < part_len) ? part[r * part_len + i] : (T)0;
  T *dst = out + r * out_stride + o;
  *dst = first ? v : *dst + v;
}

// Streaming overlap-add across calls: `full` holds rows of n + K - 1 samples (this
// call's Full convolution).  out[r][i] = full[r][i] + carry[r][i] (i < K-1), and the new
// carry is what sticks out past n (plus the part of the old carry not yet consumed).
template <typename T>
__global__ void stream_merge_kernel(const T *__restrict__ full, long long rows, long long n, int tail_len,
                                    const T *__restrict__ carry, T *__restrict__ carry_next, T *__restrict__ out,
                                    long long out_stride) {
  const long long w = n + tail_len;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * w) return;
  const long long r = idx / w, i = idx - r * w;
  T v = full[idx];
  if (i < tail_len) v += carry[r * tail_len + i];
  if (i < n) out[r * out_stride + i] = v;
  else carry_next[r * tail_len + (i - n)] = v;
}

} // namespace generic
} // namespace fc
