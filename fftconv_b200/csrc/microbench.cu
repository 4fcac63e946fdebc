// microbench.cu -- measures the B200 facts the overlap-add kernel design rests on:
// issue / pipe throughput of scalar FFMA vs packed FFMA2 / FADD2, how much
// shared-memory traffic rides along for free, and 4-byte vs 16-byte streaming
// copy bandwidth.  Run on the GPU box: fftconv_b200/lib/microbench
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      std::exit(1);                                                                    \
    }                                                                                  \
  } while (0)

constexpr int kChains = 8;

__global__ void k_ffma(float *out, int iters, float a, float b) {
  float x[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b,
                                                   unsigned long long c) {
  unsigned long long r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}

__global__ void k_ffma2(unsigned long long *out, int iters, float a, float b) {
  unsigned long long x[kChains];
  const unsigned long long pa = pk(a, a), pb = pk(b, b);
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = pk(threadIdx.x * 0.001f + i, 1.0f + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) x[i] = fma2(x[i], pa, pb);
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_fadd2(unsigned long long *out, int iters, float b) {
  unsigned long long x[kChains];
  const unsigned long long pb = pk(b, -b);
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = pk(threadIdx.x * 0.001f + i, 1.0f + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) x[i] = add2(x[i], pb);
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 with LDS.128 + STS.128 mixed in: `ratio` packed FMAs per 16-byte smem access
template <int RATIO>
__global__ void k_ffma2_smem(unsigned long long *out, int iters, float a, float b) {
  extern __shared__ float4 sm[];
  unsigned long long x[kChains];
  const unsigned long long pa = pk(a, a), pb = pk(b, b);
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = pk(threadIdx.x * 0.001f + i, 1.0f + i);
  sm[threadIdx.x] = make_float4(1, 2, 3, 4);
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < RATIO / kChains; ++r) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) x[i] = fma2(x[i], pa, pb);
    }
    const float4 v = sm[(threadIdx.x + it) & (blockDim.x - 1)];
    acc.x += v.x;
    sm[(threadIdx.x + 2 * it + 1) & (blockDim.x - 1)] = acc;
  }
  unsigned long long s = __float_as_uint(acc.x);
#pragma unroll
  for (int i = 0; i < kChains; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename V> __global__ void k_copy(const V *__restrict__ in, V *__restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

template <typename F> float time_ms(F &&launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  std::printf("device %s sms %d clock_khz %d\n", prop.name, sms, clk_khz);
  const int threads = 512, blocks = sms * 4, iters = 4096;
  void *buf;
  CK(cudaMalloc(&buf, (size_t)blocks * threads * 16));

  {
    const float ms = time_ms([&] { k_ffma<<<blocks, threads>>>((float *)buf, iters, 1.0001f, 0.5f); });
    const double inst = (double)blocks * threads * iters * kChains;
    std::printf("ffma_scalar: %.3f ms  %.2f TFLOP/s  %.1f lane-inst/clk/SM@max\n", ms,
                2 * inst / ms / 1e9, inst / (ms * 1e-3) / sms / (clk_khz * 1e3));
  }
  {
    const float ms = time_ms([&] { k_ffma2<<<blocks, threads>>>((unsigned long long *)buf, iters, 1.0001f, 0.5f); });
    const double inst = (double)blocks * threads * iters * kChains;
    std::printf("ffma2_packed: %.3f ms  %.2f TFLOP/s  %.1f lane-inst/clk/SM@max (x2 flops each)\n", ms,
                4 * inst / ms / 1e9, inst / (ms * 1e-3) / sms / (clk_khz * 1e3));
  }
  {
    const float ms = time_ms([&] { k_fadd2<<<blocks, threads>>>((unsigned long long *)buf, iters, 0.5f); });
    const double inst = (double)blocks * threads * iters * kChains;
    std::printf("fadd2_packed: %.3f ms  %.2f Tadd/s  %.1f lane-inst/clk/SM@max\n", ms,
                2 * inst / ms / 1e9, inst / (ms * 1e-3) / sms / (clk_khz * 1e3));
  }
  {
    const float ms8 = time_ms([&] {
      k_ffma2_smem<8><<<blocks, threads, threads * 16>>>((unsigned long long *)buf, iters, 1.0001f, 0.5f);
    });
    const float ms16 = time_ms([&] {
      k_ffma2_smem<16><<<blocks, threads, threads * 16>>>((unsigned long long *)buf, iters / 2, 1.0001f, 0.5f);
    });
    const float ms32 = time_ms([&] {
      k_ffma2_smem<32><<<blocks, threads, threads * 16>>>((unsigned long long *)buf, iters / 4, 1.0001f, 0.5f);
    });
    std::printf("ffma2 + (LDS.128+STS.128) per 8/16/32 FFMA2: %.3f / %.3f / %.3f ms (same FFMA2 count as ffma2_packed)\n",
                ms8, ms16, ms32);
  }
  {
    const size_t bytes = (size_t)2 << 30;
    void *a, *b;
    CK(cudaMalloc(&a, bytes));
    CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes));
    for (int per_sm : {8, 16, 32}) {
      const float m4 = time_ms([&] { k_copy<float><<<sms * per_sm, 512>>>((const float *)a, (float *)b, bytes / 4); });
      const float m16 = time_ms([&] { k_copy<float4><<<sms * per_sm, 512>>>((const float4 *)a, (float4 *)b, bytes / 16); });
      std::printf("copy 2 GiB, %d CTAs/SM x 512 thr: LDG.32 %.1f GB/s   LDG.128 %.1f GB/s (read+write)\n",
                  per_sm, 2.0 * bytes / m4 / 1e6, 2.0 * bytes / m16 / 1e6);
    }
    CK(cudaFree(a));
    CK(cudaFree(b));
  }
  CK(cudaFree(buf));
  return 0;
}
