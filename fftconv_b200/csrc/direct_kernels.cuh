// direct_kernels.cuh -- time-domain convolution for very short kernels (dispatched for K <= 16 taps
// in float32, K <= 8 in float64).
//
// The reference convolves everything through FFTs, down to 16-point segments for K < 7
// (/root/reference/include/fftconv/fftconv.hpp:38-49).  On this machine a kernel of a few taps
// is cheaper evaluated directly: 2 K flops per output sample stay below the HBM time up to a few
// dozen taps, every input sample is read once and every output sample written once, and there is
// no segment overlap to pay for.  Same linear convolution, Full and Same
// (out[o] = sum_j k[j] a[o + pad - j], pad = 0 or K / 2, fftconv.hpp:400-431), deterministic.
//
// A CTA of 256 threads produces tiles of 4096 consecutive outputs of one row: the inputs a tile
// needs (4096 + KP samples, KP = taps padded to a multiple of four) are fetched with aligned
// 128-bit loads -- all of a thread's loads are issued before the first is consumed, the bytes in
// flight are what a streaming kernel lives on -- and laid out in shared memory at their logical
// offset, so whatever the alignment of the row every thread reads its windows with aligned 128-bit
// shared loads; four rounds of four outputs per thread, each stored as one 128-bit store (the
// row's first 0..3 outputs are peeled off so that tiles start on a 16-byte boundary of the output
// row).  Grid-stride loop over (row, tile).
//
// ncu (profiles/r01_ncu_v12_direct_conv_*): LSU data pipe 78 %, DRAM 61 % of ncu's peak -- the
// scattered scalar stores that lay the window out cost a 4-way bank conflict each.  The obvious
// cure (load both aligned vectors a misaligned logical vector straddles, assemble it in registers,
// one aligned 128-bit store) was measured and is SLOWER: 0.50-0.52 ms against 0.40-0.42 (float32),
// 0.85 against 0.44 (float64) -- twice the global loads in flight per thread costs more than the
// conflicts.
#pragma once

#include <cstdint>

namespace fc {
namespace direct {

constexpr int kThreads = 256;
constexpr int kRounds = 4;                    // groups of four outputs per thread and tile
constexpr int kTile = 4 * kThreads * kRounds; // outputs per tile

template <typename T> struct DirectArgs {
  const T *in;
  T *out;
  const T *taps;
  long long batch, n, in_stride, out_stride, out_len;
  int klen, pad;
  long long tiles_per_row;
};

template <typename T, int KP>
__global__ void __launch_bounds__(kThreads) direct_conv_kernel(const DirectArgs<T> p) {
  constexpr int V = 16 / (int)sizeof(T);        // 4 floats or 2 doubles per 128-bit access
  constexpr int W = kTile + KP;                 // logical window: inputs o0 + pad - KP + [0, W)... see below
  constexpr int SM = W + 2 * V;                 // + slack for the aligned over-fetch
  __shared__ __align__(16) T xs[SM];
  __shared__ T ks[KP];
  const int tid = threadIdx.x;
  if (tid < KP) ks[tid] = tid < p.klen ? p.taps[tid] : (T)0;
  const long long total = p.batch * p.tiles_per_row;
  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const long long row = w / p.tiles_per_row, tile = w - row * p.tiles_per_row;
    const T *a = p.in + row * p.in_stride;
    T *out = p.out + row * p.out_stride;
    // outputs [0, head) of a row are peeled so that tiles start 16-byte aligned in `out`
    const int head = (int)((V - ((reinterpret_cast<uintptr_t>(out) / sizeof(T)) & (V - 1))) & (V - 1));
    const long long o0 = head + tile * kTile;            // first output of this tile
    // logical window: xs[i] = a[s0 + i], s0 = o0 + pad - KP  (zero outside [0, n))
    const long long s0 = o0 + p.pad - KP;
    // aligned over-fetch: vectors start at the 16-byte boundary at or below a + s0
    const long long addr0 = (long long)(reinterpret_cast<uintptr_t>(a) / sizeof(T)) + s0;
    const int shift = (int)(((addr0 % V) + V) % V);      // a + s0 - shift is 16-byte aligned
    __syncthreads();                                      // previous tile's window is no longer read
    constexpr int NV = (W + V + V * kThreads - 1) / (V * kThreads); // vectors per thread (upper bound)
    T e[NV][V];
#pragma unroll
    for (int u = 0; u < NV; ++u) {                         // every load first ...
      const int v = tid + u * kThreads;
      const long long g = s0 - shift + (long long)v * V;  // first element of this vector
#pragma unroll
      for (int c = 0; c < V; ++c) e[u][c] = (T)0;
      if (v * V < W + shift) {
        if (g >= 0 && g + V <= p.n) {
          if constexpr (sizeof(T) == 4) {
            const float4 q = *reinterpret_cast<const float4 *>(a + g);
            e[u][0] = q.x; e[u][1] = q.y; e[u][2] = q.z; e[u][3] = q.w;
          } else {
            const double2 q = *reinterpret_cast<const double2 *>(a + g);
            e[u][0] = q.x; e[u][1] = q.y;
          }
        } else {
#pragma unroll
          for (int c = 0; c < V; ++c)
            if (g + c >= 0 && g + c < p.n) e[u][c] = a[g + c];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < NV; ++u) {                         // ... then the window at its logical offsets
      const int v = tid + u * kThreads;
#pragma unroll
      for (int c = 0; c < V; ++c) {
        const int i = v * V + c - shift;
        if (i >= 0 && i < W) xs[i] = e[u][c];
      }
    }
    __syncthreads();
    // outputs o = o0 + 1024 q + 4 tid + r:  sum_j k[j] a[o + pad - j] = sum_j k[j] xs[1024 q + 4 tid + r + KP - j]
#pragma unroll
    for (int q = 0; q < kRounds; ++q) {
      const int base = 4 * kThreads * q + 4 * tid;
      T acc[4] = {(T)0, (T)0, (T)0, (T)0};
      T win[KP + 4];
#pragma unroll
      for (int c = 0; c < (KP + 4) / V; ++c) {
        if constexpr (sizeof(T) == 4) {
          const float4 w4 = *reinterpret_cast<const float4 *>(xs + base + c * 4);
          win[4 * c] = w4.x; win[4 * c + 1] = w4.y; win[4 * c + 2] = w4.z; win[4 * c + 3] = w4.w;
        } else {
          const double2 w2 = *reinterpret_cast<const double2 *>(xs + base + c * 2);
          win[2 * c] = w2.x; win[2 * c + 1] = w2.y;
        }
      }
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        const T kj = ks[j];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] += kj * win[r + KP - j];
      }
      const long long o = o0 + base;
      if (o + 4 <= p.out_len) {
        if constexpr (sizeof(T) == 4) {
          *reinterpret_cast<float4 *>(out + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        } else {
          *reinterpret_cast<double2 *>(out + o) = make_double2(acc[0], acc[1]);
          *reinterpret_cast<double2 *>(out + o + 2) = make_double2(acc[2], acc[3]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (o + r < p.out_len) out[o + r] = acc[r];
      }
    }
    // the peeled head of the row (at most 3 outputs), by the row's first tile
    if (tile == 0 && tid < head && tid < p.out_len) {
      T s = (T)0;
      for (int j = 0; j < p.klen; ++j) {
        const long long i = (long long)tid + p.pad - j;
        if (i >= 0 && i < p.n) s += ks[j] * a[i];
      }
      out[tid] = s;
    }
  }
}

} // namespace direct
} // namespace fc

// ---- small jobs: one thread per output sample ----------------------------------------------------
// A single short signal (the reference's own benchmark case is n = 4352, K = 65, 17.2 us per call
// on a CPU, /root/reference/README.md:99-102) is bound by launch latency, not by arithmetic: the
// FFT path needs two launches (kernel spectrum, segments) and a few dependent passes.  When the
// whole job is at most a few million multiply-adds the definition is evaluated directly, one
// thread per output, accumulating in double; taps of up to kInlineTaps ride in the kernel
// arguments so that a host-pointer call needs no separate copy for them.
namespace fc {
namespace direct {

constexpr int kInlineTaps = 96;

template <typename T> struct SmallArgs {
  const T *in;
  T *out;
  const T *taps;     // used when klen > kInlineTaps or the taps already live on the device
  long long batch, n, in_stride, out_stride, out_len;
  int klen, pad, inline_taps;
  T tap[kInlineTaps];
};

template <typename T>
__global__ void __launch_bounds__(256) direct_small_kernel(const __grid_constant__ SmallArgs<T> p) {
  __shared__ T ks[256];
  for (int j = threadIdx.x; j < p.klen; j += blockDim.x) ks[j] = p.inline_taps ? p.tap[j] : p.taps[j];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.batch * p.out_len) return;
  const long long row = idx / p.out_len, o = idx - row * p.out_len;
  const T *a = p.in + row * p.in_stride;
  // out[o] = sum_j k[j] a[o + pad - j], 0 <= o + pad - j < n
  const long long hi = o + p.pad;                              // index for j = 0
  int j0 = hi >= p.n ? (int)(hi - p.n + 1) : 0;
  int j1 = hi < p.klen - 1 ? (int)hi : p.klen - 1;             // inclusive
  double acc = 0.0;
  for (int j = j0; j <= j1; ++j) acc += (double)ks[j] * (double)a[hi - j];
  p.out[row * p.out_stride + o] = (T)acc;
}

// The same job for a HOST-pointer call (config 1 through the reference's span API): the signal
// sits in a page-locked buffer mapped into the device address space, the kernel reads it across
// PCIe exactly once -- each CTA stages the 256 + K - 1 inputs of its 256 outputs in shared memory
// with coalesced loads -- and writes the outputs straight back into mapped host memory: one
// launch and one synchronisation per call, no cudaMemcpy in either direction.
template <typename T>
__global__ void __launch_bounds__(256) direct_tiny_kernel(const __grid_constant__ SmallArgs<T> p, int tiles_per_row,
                                                          unsigned *counter, unsigned *flag, unsigned epoch) {
  __shared__ T ks[256];
  __shared__ T xs[512];
  const int row = blockIdx.x / tiles_per_row, tile = blockIdx.x - row * tiles_per_row;
  const T *a = p.in + (long long)row * p.in_stride;
  const long long o0 = (long long)tile * 256;
  const long long lo = o0 + p.pad - (p.klen - 1);              // input index of xs[0]
  for (int j = threadIdx.x; j < p.klen; j += 256) ks[j] = p.inline_taps ? p.tap[j] : p.taps[j];
  for (int i = threadIdx.x; i < 256 + p.klen - 1; i += 256) {
    const long long src = lo + i;
    xs[i] = (src >= 0 && src < p.n) ? a[src] : (T)0;
  }
  __syncthreads();
  const long long o = o0 + threadIdx.x;
  if (o < p.out_len) {
    // out[o] = sum_j k[j] a[o + pad - j]; a[o + pad - j] = xs[threadIdx.x + K - 1 - j]
    double acc = 0.0;
    const T *w = xs + threadIdx.x + p.klen - 1;
    for (int j = 0; j < p.klen; ++j) acc += (double)ks[j] * (double)w[-j];
    p.out[(long long)row * p.out_stride + o] = (T)acc;
  }
  // completion: every CTA's outputs are fenced system-wide before it is counted; the last CTA
  // resets the counter for the next call and publishes the epoch to the host
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(counter, 1u) + 1u;
    if (done == gridDim.x) {
      *counter = 0u;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned *>(flag) = epoch;
    }
  }
}

// The K - 1 outputs past the end of a chunk's Full convolution depend only on the chunk's last
// K - 1 samples: tail[i] = sum_{j > i} k[j] last[K - 1 + i - j].  One thread per output, double
// accumulation -- (K - 1)^2 / 2 multiply-adds in all.  Used when one long signal is cut across
// GPUs: the tail is produced first and travels to the neighbour while the chunk is convolved.
template <typename T>
__global__ void oa_tail_kernel(const T *__restrict__ last, const T *__restrict__ k, int klen, T *__restrict__ tail) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= klen - 1) return;
  double acc = 0.0;
  for (int j = i + 1; j < klen; ++j) acc += (double)k[j] * (double)last[klen - 1 + i - j];
  tail[i] = (T)acc;
}

} // namespace direct
} // namespace fc
