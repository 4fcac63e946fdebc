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
#include "oa_n2048_kernel.cuh" // the TMA / mbarrier wrappers (n2048::dev)

#ifndef FC_ENV_BATCH
#define FC_ENV_BATCH 8 // envelope: input samples re-read per batch (x 4 rows in flight per thread); measured on config 4: 4 -> 0.214 ms, 8 -> 0.209, 16 -> 0.210
#endif

namespace fc {
namespace fast {

// occupancy target: ~768 threads per SM (<= 85 registers) for float, ~512 (<= 128) for double
// and for the packed float kernels (D = f2: four real streams per FFT, 64 data registers)
template <typename D, int CTA_THREADS> constexpr int min_blocks() {
  constexpr int target = sizeof(D) == 4 ? 768 : 512;
  return CTA_THREADS >= target ? 1 : target / CTA_THREADS;
}

struct CtaSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
  __device__ __forceinline__ void warp() const { __syncwarp(); } // warp-local exchange (fast_fft.cuh, pass_sync)
  template <int LM> __device__ __forceinline__ void sub(int) const { __syncthreads(); }
};

// Barrier among the threads of ONE of the FPC FFTs that share a CTA (named barrier f of 16): the
// FFTs of a CTA are independent, so they need not wait for one another -- while one is in a
// shared-memory exchange another is in its butterflies.  Kernels that use it must not call
// __syncthreads() in the same phase (that is barrier 0).
struct FftSync {
  int id, threads;
  int sub_base = -1; // >= 0: named barriers sub_base, sub_base + 1, ... are free for the sub-transforms of this FFT
  __device__ __forceinline__ void operator()() const {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
  }
  __device__ __forceinline__ void warp() const { __syncwarp(); }
  // the 2^LM / 16 threads that share one sub-transform of 2^LM points (mid pass: block tid >> (LM - 4))
  template <int LM> __device__ __forceinline__ void sub(int tid) const {
    constexpr int TB = (1 << LM) / 16;
    if (sub_base >= 0 && TB < threads) {
      asm volatile("bar.sync %0, %1;" ::"r"(sub_base + tid / TB), "r"(TB) : "memory");
    } else {
      (*this)();
    }
  }
};

// The real streams ("lanes") that ride in one complex value of data type D:
//   D = T:   lane 0 = real part, lane 1 = imaginary part
//   D = f2:  lane 0 = re half A, lane 1 = im half A, lane 2 = re half B, lane 3 = im half B
template <typename D> struct Lanes {
  static constexpr int N = 2;
  static __device__ __forceinline__ cx<D> make(const D (&a)[2]) { return {a[0], a[1]}; }
  static __device__ __forceinline__ void split(cx<D> c, D (&a)[2]) {
    a[0] = c.x;
    a[1] = c.y;
  }
};
template <> struct Lanes<f2> {
  static constexpr int N = 4;
  static __device__ __forceinline__ cx<f2> make(const float (&a)[4]) {
    return {make_f2(a[0], a[2]), make_f2(a[1], a[3])};
  }
  static __device__ __forceinline__ void split(cx<f2> c, float (&a)[4]) {
    a[0] = c.x.x;
    a[1] = c.y.x;
    a[2] = c.x.y;
    a[3] = c.y.y;
  }
};

// D = T (two streams per FFT) or D = f2 with T = float (four streams per FFT)
template <typename D, int LOG2N, int FPC>
__global__ void __launch_bounds__(Plan<LOG2N>::THREADS *FPC, min_blocks<D, Plan<LOG2N>::THREADS * FPC>())
    oa_fast_kernel(const generic::OaArgs<scalar_t<D>> p) {
  using P = Plan<LOG2N>;
  using T = scalar_t<D>;
  using L = Lanes<D>;
  constexpr int N = P::N, NL = L::N;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int f = threadIdx.x / P::THREADS, tid = threadIdx.x % P::THREADS;
  const int tail_len = p.klen - 1;
  cx<D> *buf = reinterpret_cast<cx<D> *>(smem_raw) + (size_t)f * P::SMEM_ELEMS;
  cx<D> *tails = reinterpret_cast<cx<D> *>(smem_raw) + (size_t)FPC * P::SMEM_ELEMS + (size_t)f * 2 * tail_len;
  const cx<T> *W = reinterpret_cast<const cx<T> *>(p.W);
  const cx<T> *H = reinterpret_cast<const cx<T> *>(p.H);

  // CTA-uniform loop bounds over all FPC FFTs (barriers are CTA-wide)
  long long max_cnt = 0;
  bool warm_needed = false;
  const T *src[NL];
  T *dst[NL];
  long long seg0[NL], cnt[NL];
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    src[l] = p.in;
    dst[l] = p.out;
    seg0[l] = cnt[l] = 0;
  }
#pragma unroll
  for (int ff = 0; ff < FPC; ++ff)
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const long long sid = ((long long)blockIdx.x * FPC + ff) * NL + l;
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

  // one barrier domain per FFT (the trip count stays CTA-uniform, so every domain executes the same
  // number of barriers); a single FFT per CTA is the CTA barrier itself
  const FftSync sync{FPC > 1 ? f : 0, FPC > 1 ? P::THREADS : P::THREADS * FPC, FPC == 1 ? 1 : -1};
  int parity = 0;
  bool add_tail = false;
  for (long long j = warm_needed ? -1 : 0; j < max_cnt; ++j) {
    const bool warm = j < 0;
    long long pos[NL];
    int len[NL], limit[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
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
    cx<D> v[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      const int i = pass0_index<LOG2N>(tid, s);
      T x[NL];
#pragma unroll
      for (int l = 0; l < NL; ++l) x[l] = i < len[l] ? src[l][pos[l] + i] : (T)0;
      v[s] = L::make(x);
    }
    conv_core<LOG2N>(tid, v, buf, W, H, sync);
    cx<D> *tail_prev = tails + (parity ^ 1) * tail_len;
    cx<D> *tail_next = tails + parity * tail_len;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      const int i = pass0_index<LOG2N>(tid, s);
      cx<D> y = v[s];
      if (add_tail && i < tail_len) y = y + tail_prev[i];
      if (i >= p.step) tail_next[i - p.step] = y;
      T ys[NL];
      L::split(y, ys);
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const long long o = pos[l] - p.pad + i;
        if (i < limit[l] && o >= 0) dst[l][o] = ys[l];
      }
    }
    sync(); // tails and the exchange buffer (both per FFT) are reused by the next segment
    parity ^= 1;
    add_tail = true;
  }
}

// Hilbert envelope of rows of length N (power of two): rows 2b, 2b+1 (or 4b .. 4b+3) per FFT.
// G (position order) = -j sgn(f) / N with DC and Nyquist zeroed (hilbert.hpp:39-51).
template <typename D, int LOG2N, int FPC, bool PLAIN = true>
__global__ void __launch_bounds__(Plan<LOG2N>::THREADS *FPC, min_blocks<D, Plan<LOG2N>::THREADS * FPC>())
    hilbert_fast_kernel(const generic::HilbertArgs<scalar_t<D>> p) {
  using P = Plan<LOG2N>;
  using T = scalar_t<D>;
  using L = Lanes<D>;
  constexpr int NL = L::N;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int f = threadIdx.x / P::THREADS, tid = threadIdx.x % P::THREADS;
  cx<D> *buf = reinterpret_cast<cx<D> *>(smem_raw) + (size_t)f * P::SMEM_ELEMS;
  const cx<T> *W = reinterpret_cast<const cx<T> *>(p.W);
  bool has[NL];
  const T *x[NL];
  T *e[NL];
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    const long long r = ((long long)blockIdx.x * FPC + f) * NL + l;
    has[l] = r < p.batch;
    x[l] = p.in + (has[l] ? r : 0) * p.in_stride;
    e[l] = p.out + (has[l] ? r : 0) * p.out_stride;
  }
  // plain float keeps the inputs in registers for the envelope; double and packed float
  // re-read them (16 more complex values would not fit 128 registers per thread)
  constexpr bool kKeep = sizeof(D) == 4;
  cx<D> v[16], keep[kKeep ? 16 : 1];
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    const int i = pass0_index<LOG2N>(tid, s);
    T a[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) a[l] = has[l] ? x[l][i] : (T)0;
    v[s] = L::make(a);
    if constexpr (kKeep) keep[s] = v[s];
  }
  const FftSync sync{FPC > 1 ? f : 0, FPC > 1 ? P::THREADS : P::THREADS * FPC, FPC == 1 ? 1 : -1};
  const HilbertSpectrum<T> G{(T)(1.0 / (double)P::N)};
  conv_core_with<LOG2N>(tid, v, buf, W, G, sync);
  // envelope; the inputs are re-read (L2) in batches so that their latencies overlap
  constexpr int kBatch = kKeep ? 16 : (NL == 4 ? 4 : 8);
#pragma unroll
  for (int s0 = 0; s0 < 16; s0 += kBatch) {
    T a[kBatch][NL];
#pragma unroll
    for (int s = 0; s < kBatch; ++s) {
      if constexpr (kKeep) {
        L::split(keep[s0 + s], a[s]);
      } else {
        const int i = pass0_index<LOG2N>(tid, s0 + s);
#pragma unroll
        for (int l = 0; l < NL; ++l) a[s][l] = has[l] ? __ldg(x[l] + i) : (T)0;
      }
    }
#pragma unroll
    for (int s = 0; s < kBatch; ++s) {
      const int i = pass0_index<LOG2N>(tid, s0 + s);
      T h[NL];
      L::split(v[s0 + s], h);
#pragma unroll
      for (int l = 0; l < NL; ++l)
        if (has[l]) env_store<PLAIN>(p.epi, e[l], i, (T)sqrt(a[s][l] * a[s][l] + h[l] * h[l]));
    }
  }
}

// sqrt(a^2 + h^2) for the envelope: the float kernel uses the hardware approximation (MUFU.SQRT,
// <= 1 ulp; the parity bound is 1e-5 relative L2), which has no slow path and no branch.
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Persistent, TMA-fed Hilbert envelope for packed float (four rows per FFT, FPC FFTs per CTA, one
// CTA per SM).  ncu of the one-shot kernel: half of a CTA's life is I/O --
// every warp waits for its 64 scattered 4-byte loads at the start and nothing overlaps the
// envelope / store phase at the end.  Here
//   * the rows arrive by bulk asynchronous copies (cp.async.bulk, one thread) into the exchange
//     buffer, which they alias ([4][N] floats = N 16-byte points), and are consumed with 4-byte
//     shared loads;
//   * the NEXT unit's rows are requested as soon as the last inverse pass has read the buffer,
//     so they land while this unit's envelope is computed and stored.
// Requires 16-byte aligned rows (base and stride); other layouts take hilbert_fast_kernel.
template <int LOG2N, int FPC, bool PLAIN = true, typename D = f2, int CTAS_PER_SM = 1>
__global__ void __launch_bounds__(Plan<LOG2N>::THREADS *FPC, CTAS_PER_SM)
    hilbert_pk_persistent_kernel(const generic::HilbertArgs<float> p) {
  using P = Plan<LOG2N>;
  namespace dev = n2048::dev;
  constexpr int N = P::N;
  constexpr int NL = Lanes<D>::N; // rows per FFT: 4 (packed) or 2 (D = float: two CTAs per SM, see the launcher)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int f = threadIdx.x / P::THREADS, tid = threadIdx.x % P::THREADS;
  cx<D> *buf = reinterpret_cast<cx<D> *>(smem_raw) + (size_t)f * P::SMEM_ELEMS;
  const float *stage = reinterpret_cast<const float *>(buf); // [NL][N], aliases the exchange buffer
  const uint32_t mbar = dev::smem_u32(smem_raw + (size_t)FPC * P::SMEM_ELEMS * sizeof(cx<D>));
  const cx<float> *W = reinterpret_cast<const cx<float> *>(p.W);
  const long long units = (p.batch + NL * FPC - 1) / (NL * FPC);

  auto request = [&](long long unit) { // thread 0: bulk loads of every valid row of `unit`
    const long long r0 = unit * NL * FPC;
    long long rows = p.batch - r0;
    if (rows > NL * FPC) rows = NL * FPC;
    dev::fence_proxy_async();
    dev::mbar_expect_tx(mbar, (uint32_t)(rows * N * 4));
    for (int r = 0; r < (int)rows; ++r) {
      unsigned char *dst = smem_raw + (size_t)(r / NL) * P::SMEM_ELEMS * sizeof(cx<D>) + (size_t)(r % NL) * N * 4;
      dev::bulk_load(dev::smem_u32(dst), p.in + (r0 + r) * p.in_stride, (uint32_t)(N * 4), mbar);
    }
  };
  if (threadIdx.x == 0) {
    dev::mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long unit = blockIdx.x;
  if (threadIdx.x == 0 && unit < units) request(unit);
  uint32_t phase = 0;
  // inside the transform every FFT of the CTA has its own barrier (ids 1 .. FPC; barrier 0 stays the
  // CTA barrier around the staged rows, which all FFTs share); sixteen FFTs would need one id too many
  constexpr bool kOwnBarrier = FPC > 1 && FPC <= 15;
  // a single FFT per CTA (8192-sample lines): its independent sub-transforms get barriers 1, 2, ...
  const FftSync sync{kOwnBarrier ? f + 1 : 0, kOwnBarrier ? P::THREADS : P::THREADS * FPC, FPC == 1 ? 1 : -1};
  const HilbertSpectrum<float> G{(float)(1.0 / (double)N)};
  for (; unit < units; unit += gridDim.x) {
    bool has[NL];
    const float *x[NL];
    float *e[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const long long r = (unit * FPC + f) * NL + l;
      has[l] = r < p.batch;
      x[l] = p.in + (has[l] ? r : 0) * p.in_stride;
      e[l] = p.out + (has[l] ? r : 0) * p.out_stride;
    }
    while (!dev::mbar_try_wait(mbar, phase)) {
    }
    phase ^= 1;
    cx<D> v[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      const int i = pass0_index<LOG2N>(tid, s);
      float a[NL];
#pragma unroll
      for (int l = 0; l < NL; ++l) a[l] = has[l] ? stage[l * N + i] : 0.0f; // rows past the batch were not loaded
      v[s] = Lanes<D>::make(a);
    }
    __syncthreads(); // the staged rows have been read; pass 0 overwrites them
    conv_core_with<LOG2N>(tid, v, buf, W, G, sync);
    __syncthreads(); // the last inverse pass has read the buffer: it is idle until the next unit
    if (threadIdx.x == 0 && unit + gridDim.x < units) request(unit + gridDim.x);
    bool all = true;
#pragma unroll
    for (int l = 0; l < NL; ++l) all = all && has[l];
    if (all) {
      constexpr int EB = PLAIN ? FC_ENV_BATCH : 4; // (the post-processing store needs the registers: 8 measured slower there)
#pragma unroll
      for (int s0 = 0; s0 < 16; s0 += EB) {
        float a[EB][NL];
#pragma unroll
        for (int s = 0; s < EB; ++s) {
          const int i = pass0_index<LOG2N>(tid, s0 + s);
#pragma unroll
          for (int l = 0; l < NL; ++l) a[s][l] = __ldg(x[l] + i);
        }
#pragma unroll
        for (int s = 0; s < EB; ++s) {
          const int i = pass0_index<LOG2N>(tid, s0 + s);
          float h[NL];
          Lanes<D>::split(v[s0 + s], h);
#pragma unroll
          for (int l = 0; l < NL; ++l) env_store<PLAIN>(p.epi, e[l], i, fast_sqrt(a[s][l] * a[s][l] + h[l] * h[l]));
        }
      }
    } else {
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        const int i = pass0_index<LOG2N>(tid, s);
        float h[NL];
        Lanes<D>::split(v[s], h);
#pragma unroll
        for (int l = 0; l < NL; ++l)
          if (has[l]) {
            const float a = __ldg(x[l] + i);
            env_store<PLAIN>(p.epi, e[l], i, fast_sqrt(a * a + h[l] * h[l]));
          }
      }
    }
  }
}

} // namespace fast
} // namespace fc
