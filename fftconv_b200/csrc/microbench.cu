// microbench.cu -- measures the B200 facts the overlap-add kernel design rests on:
// issue / pipe throughput of scalar FFMA vs packed FFMA2 / FADD2, how much
// shared-memory traffic rides along for free, and 4-byte vs 16-byte streaming
// copy bandwidth.  Run on the GPU box: fftconv_b200/lib/microbench
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <vector>
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

// ---- host-call latency floor (BASELINE config 1 through host pointers) --------------------------------
__global__ void k_flag(volatile unsigned *flag, unsigned epoch) {
  __threadfence_system();
  if (threadIdx.x == 0 && blockIdx.x == 0) *flag = epoch;
}
__global__ void k_touch(const float *in, float *out, int n, volatile unsigned *flag, unsigned epoch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * 2.0f;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *flag = epoch; // (not a completion protocol: a latency probe)
}
template <class F> static double host_us(F &&f, int reps) {
  for (int i = 0; i < 200; ++i) f();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) f();
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
}
static void host_latency_floor() {
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  void *hm = nullptr, *dm = nullptr;
  CK(cudaHostAlloc(&hm, 1 << 20, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer(&dm, hm, 0));
  volatile unsigned *flag_h = (volatile unsigned *)hm;
  unsigned *flag_d = (unsigned *)dm;
  float *in_d = (float *)dm + 1024, *out_d = (float *)dm + 1024 + 8192;
  std::vector<float> pageable(4352, 1.0f);
  unsigned epoch = 0;
  const double t_attr = host_us([&] {
    cudaPointerAttributes a;
    cudaPointerGetAttributes(&a, pageable.data());
  }, 20000);
  const double t_sync = host_us([&] {
    k_flag<<<1, 32, 0, st>>>(flag_d, ++epoch);
    cudaStreamSynchronize(st);
  }, 5000);
  const double t_poll = host_us([&] {
    k_flag<<<1, 32, 0, st>>>(flag_d, ++epoch);
    while (*flag_h != epoch) {
    }
  }, 5000);
  const double t_touch = host_us([&] {
    k_touch<<<18, 256, 0, st>>>(in_d, out_d, 4416, flag_d, ++epoch);
    while (*flag_h != epoch) {
    }
  }, 5000);
  float *dd;
  CK(cudaMalloc((void **)&dd, 1 << 20));
  std::vector<float> back(4416);
  const double t_copy = host_us([&] {
    cudaMemcpyAsync(dd, pageable.data(), 4352 * 4, cudaMemcpyHostToDevice, st);
    k_flag<<<1, 32, 0, st>>>(flag_d, ++epoch);
    cudaMemcpyAsync(back.data(), dd, 4416 * 4, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
  }, 3000);
  std::printf("host-call latency floor (us): cudaPointerGetAttributes(pageable) %.2f | empty kernel + "
              "cudaStreamSynchronize %.2f | empty kernel + poll of a mapped flag %.2f | 18-CTA kernel reading 17 KB / "
              "writing 17 KB of mapped host memory + poll %.2f | cudaMemcpyAsync H2D 17 KB + kernel + D2H 17 KB + sync %.2f\n",
              t_attr, t_sync, t_poll, t_touch, t_copy);
}

int main() {
  host_latency_floor();
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
