// microbench_rf.cu -- do packed FP32 instructions keep their 2-cycle rate when all three 64-bit operands are
// distinct registers (register-file bandwidth), as in the FFT butterflies?  Compare with scalar FFMA / FADD.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_rf tools/microbench_rf.cu
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("%s\n", cudaGetErrorString(e)); return 1; } } while (0)
typedef unsigned long long u64;
constexpr int R = 24;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fmas2(u64 a, float s, u64 c) { u64 r; asm volatile("{.reg .b64 pb; mov.b64 pb, {%2, %2}; fma.rn.f32x2 %0, %1, pb, %3;}" : "=l"(r) : "l"(a), "f"(s), "l"(c)); return r; }
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int MODE> __global__ void __launch_bounds__(512, 1) k_packed(u64 *out, int iters, float s0) {
  u64 x[R];
#pragma unroll
  for (int i = 0; i < R; ++i) x[i] = pk(threadIdx.x * 1e-3f + i, 1.0f + i * 1e-3f);
  float s[4] = {s0, s0 * 0.5f, s0 * 0.25f, s0 * 0.125f};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      if (MODE == 0) x[i] = fma2(x[(i + 5) % R], x[(i + 11) % R], x[i]);        // three distinct 64-bit operands
      if (MODE == 1) x[i] = fmas2(x[(i + 5) % R], s[i & 3], x[i]);              // pair, broadcast scalar, pair
      if (MODE == 2) x[i] = add2(x[(i + 5) % R], x[(i + 11) % R]);              // two distinct pairs
      if (MODE == 3) x[i] = mul2(x[(i + 5) % R], x[(i + 11) % R]);
      if (MODE == 4) x[i] = fma2(x[i], x[(i + 11) % R], x[i]);                  // two distinct pairs (a == c)
      if (MODE == 5) x[i] = fma2(x[(i + 5) % R], x[(i + 11) % R], x[(i + 17) % R]); // dest differs from all sources
    }
  }
  u64 t = 0;
#pragma unroll
  for (int i = 0; i < R; ++i) t ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int MODE> __global__ void __launch_bounds__(512, 1) k_scalar(float *out, int iters, float s0) {
  float x[2 * R];
#pragma unroll
  for (int i = 0; i < 2 * R; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) {
      if (MODE == 0) x[i] = fmaf(x[(i + 5) % (2 * R)], x[(i + 11) % (2 * R)], x[i]);
      if (MODE == 1) x[i] = x[(i + 5) % (2 * R)] + x[(i + 11) % (2 * R)];
      if (MODE == 2) x[i] = fmaf(x[(i + 5) % (2 * R)], s0, x[i]);
    }
  }
  float t = 0;
#pragma unroll
  for (int i = 0; i < 2 * R; ++i) t += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <class F> static float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount; int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  void *buf; CK(cudaMalloc(&buf, (size_t)sms * 4 * 512 * 8));
  const int iters = 8192;
  for (int threads : {128, 256, 512}) {
    const char *names[] = {"fma2 3 distinct pairs", "fma2 pair*scalar+pair", "add2 2 distinct pairs", "mul2 2 distinct pairs", "fma2 a*b+a", "fma2 dest != sources"};
    float ms[6];
    ms[0] = time_ms([&] { k_packed<0><<<sms, threads>>>((u64 *)buf, iters, 1e-3f); });
    ms[1] = time_ms([&] { k_packed<1><<<sms, threads>>>((u64 *)buf, iters, 1e-3f); });
    ms[2] = time_ms([&] { k_packed<2><<<sms, threads>>>((u64 *)buf, iters, 1e-3f); });
    ms[3] = time_ms([&] { k_packed<3><<<sms, threads>>>((u64 *)buf, iters, 1e-3f); });
    ms[4] = time_ms([&] { k_packed<4><<<sms, threads>>>((u64 *)buf, iters, 1e-3f); });
    ms[5] = time_ms([&] { k_packed<5><<<sms, threads>>>((u64 *)buf, iters, 1e-3f); });
    for (int m = 0; m < 6; ++m) {
      const double warp_inst = (double)(threads / 32) * iters * R; // per SM
      const double cyc = ms[m] * 1e-3 * clk * 1e3;
      std::printf("threads %3d  %-24s %.3f ms  %.2f cycles per warp-instruction per scheduler\n", threads, names[m], ms[m], cyc / (warp_inst / 4));
    }
    const char *sn[] = {"ffma 3 distinct", "fadd 2 distinct", "ffma reg*scalar+reg"};
    float t[3];
    t[0] = time_ms([&] { k_scalar<0><<<sms, threads>>>((float *)buf, iters, 1e-3f); });
    t[1] = time_ms([&] { k_scalar<1><<<sms, threads>>>((float *)buf, iters, 1e-3f); });
    t[2] = time_ms([&] { k_scalar<2><<<sms, threads>>>((float *)buf, iters, 1e-3f); });
    for (int m = 0; m < 3; ++m) {
      const double warp_inst = (double)(threads / 32) * iters * 2 * R;
      const double cyc = t[m] * 1e-3 * clk * 1e3;
      std::printf("threads %3d  %-24s %.3f ms  %.2f cycles per warp-instruction per scheduler\n", threads, sn[m], t[m], cyc / (warp_inst / 4));
    }
  }
  return 0;
}
