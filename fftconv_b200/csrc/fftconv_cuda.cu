// fftconv_cuda.cu -- libfftconv_cuda: C ABI (include/fftconv_cuda.h), plan /
// workspace cache and kernel dispatch for the B200 convolution hot path.
//
// What the cache replaces: the reference keeps one FFTW plan pair plus three
// aligned buffers per FFT length in a thread_local map
// (/root/reference/include/fftconv/fftw.hpp:443-461, fftconv.hpp:288-336).
// Here a "plan" is keyed by (device, FFT length, dtype) and owns the twiddle
// tables in device memory; per (device, stream) there is a small workspace for
// the taps and their spectrum; per device there are staging buffers and three
// streams for the host-pointer path.  The segment length comes from the same
// table as the reference (fftconv.hpp:38-66).
#include "../../include/fftconv_cuda.h"

#include "direct_kernels.cuh"
#include "fast_kernels.cuh"
#include "fft_global.cuh"
#include "generic_kernels.cuh"
#include "long_fft_kernels.cuh"
#include "oa_n2048_kernel.cuh"
#include "oa_n2048_tables.h"
#include "hilbert_r2c_launch.h"
#include "oa_w1k_launch.h"
#include "synthetic.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <type_traits>
#include <vector>

namespace {

using fc::generic::cx;

thread_local std::string g_err;
// Locking (the reference is lock-free per thread through thread_local plan caches,
// /root/reference/include/fftconv/fftw.hpp:443-455; device memory and streams are shared, so here):
//   g_map_mu       the device -> context map, held for the lookup only
//   DeviceCtx::mu  one device's plan / table / workspace cache, held while a call looks up or
//                  builds its plan and ENQUEUES its kernels -- never across a host transfer or a
//                  stream synchronisation
//   DeviceCtx::host_mu  one device's host-pointer pipeline (three streams + staging buffers),
//                  held for the duration of a host-pointer call on that device
// Calls on different devices never share a lock; device-pointer calls on one device only
// serialise their enqueue.
std::mutex g_map_mu;
std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_reserved_sms{0}; // SMs the persistent kernels leave free (fftconv_cuda_set_reserved_sms)

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CU_TRY(expr)                                                                        \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(FFTCONV_CUDA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

// /root/reference/include/fftconv/fftconv.hpp:38-66
constexpr size_t kOaTable[9][2] = {{7, 16},    {12, 32},    {21, 64},    {36, 128},  {65, 256},
                                   {120, 512}, {221, 1024}, {411, 2048}, {769, 4096}};
constexpr size_t kMaxOaFft = 8192;

size_t optimal_fft_size(size_t klen) {
  for (const auto &row : kOaTable)
    if (klen < row[0]) return row[1];
  size_t size = kMaxOaFft;
  while (size < klen * 2) size *= 2;
  return size;
}

int ilog2(size_t n) {
  int l = 0;
  while ((size_t(1) << l) < n) ++l;
  return l;
}

struct Tables {           // per (FFT length, dtype)
  void *W = nullptr;      // cx<T>[N]   exp(-2 pi i k / N)
  double2 *Wd = nullptr;  // double2[N] same, double (spectrum kernel)
  void *G = nullptr;      // cx<T>[N]   Hilbert multiplier, DIF order (lazy)
  void *Gfast = nullptr;  // cx<T>[N]   Hilbert multiplier, fast-engine table order (lazy)
  void *TWfast = nullptr; // cx<T>[..]  fast-engine per-pass twiddle tables (lazy)
  void *TWbig = nullptr;  // cx<T>[15 N/16] outer-pass twiddles of the long FFT, N = 16 N_sub (lazy)
  void *Glong = nullptr;  // cx<T>[N]   Hilbert multiplier in the long-FFT table order (lazy)
};

struct BluesteinTables {  // per (n, dtype): Hilbert of a non-power-of-two length
  size_t M = 0;
  void *chirp = nullptr;  // cx<T>[n]
  void *B = nullptr;      // cx<T>[M], DIF order, 1/M folded in
};

struct StreamWs {         // per stream: everything a call in flight on that stream may still be using
  void *taps = nullptr;
  size_t taps_bytes = 0;
  void *H = nullptr;
  size_t h_bytes = 0;
  void *rf = nullptr;     // filtered rows between the two launches of rf_envelope
  size_t rf_bytes = 0;
  void *part = nullptr;   // partial convolution with one block of a long kernel
  size_t part_bytes = 0;
  void *scratch = nullptr; // long-FFT intermediate (grow only): per stream, so that launches on
  size_t scratch_bytes = 0; // different streams (the host pipeline's three) never share it
  unsigned *sched = nullptr; // two counters of the warp-per-FFT kernel's dynamic tail (oa_w1k_group.cuh), zero between launches
};
void free_ws(StreamWs &w) {
  cudaFree(w.taps);
  cudaFree(w.H);
  cudaFree(w.rf);
  cudaFree(w.part);
  cudaFree(w.scratch);
  cudaFree(w.sched);
  w = StreamWs();
}

constexpr int kPipe = 3;
constexpr int kRing = 8;  // worker threads (each: stream + device staging + page-locked in / out) for pageable callers
struct RingSlot {
  cudaStream_t st = nullptr;
  void *din = nullptr, *dout = nullptr, *dtmp = nullptr, *hin = nullptr, *hout = nullptr;
  size_t cap_din = 0, cap_dout = 0, cap_dtmp = 0, cap_hin = 0, cap_hout = 0;
};

struct DeviceCtx {
  std::mutex mu;      // cache + enqueue
  std::mutex host_mu; // host-pointer pipeline
  int device = -1;
  int sm_count = 0;
  size_t smem_optin = 0;
  std::map<std::pair<size_t, int>, Tables> tables;
  std::map<std::pair<size_t, int>, BluesteinTables> bluestein;
  fc::n2048::cpair *tw1 = nullptr, *tw2 = nullptr;
  fc::n2048::cpair *tw_w1k = nullptr; // per-lane twiddles of the warp-per-FFT engine (oa_w1k_core.cuh)
  // slice plans of the N = 2048 kernel, keyed by (dtype, batch, n, K, slices): where every group of
  // the grid starts (oa_n2048_tables.h, plan_slices) -- computed once per shape, kept on the device
  struct SlicePlanDev {
    fc::n2048::SlotStart *starts = nullptr;
    long long vrows = 1, vlen = 0, n_streams = 0, n_quads = 0;
  };
  std::map<std::array<long long, 5>, SlicePlanDev> slice_plans;
  // general FFT surface (fft_global.cuh): twiddle tables per (length, dtype); chirp-z plans per (n, dtype)
  struct ChirpPlan {
    size_t M = 0;
    void *chirp = nullptr; // cx[n]  exp(-pi i j^2 / n)
    void *Bf = nullptr;    // cx[M]  FFT_M of the wrapped conjugate chirp
  };
  void *hr2c_tables = nullptr; // the real-input Hilbert engine's twiddle / mirror tables (hilbert_r2c_*.cuh)
  std::map<std::pair<size_t, int>, void *> gfft_w;
  std::map<std::pair<size_t, int>, ChirpPlan> gfft_chirp;
  fc::n2048::CPairT<double> *tw1d = nullptr, *tw2d = nullptr; // the float64 instantiation's tables
  std::map<cudaStream_t, StreamWs> ws;
  cudaStream_t pipe[kPipe] = {nullptr, nullptr, nullptr};
  void *stage_in[kPipe] = {nullptr, nullptr, nullptr};
  void *stage_out[kPipe] = {nullptr, nullptr, nullptr};
  void *stage_tmp[kPipe] = {nullptr, nullptr, nullptr}; // filtered rows between the two steps of OP_RF
  size_t cap_in[kPipe] = {0, 0, 0}, cap_out[kPipe] = {0, 0, 0}, cap_tmp[kPipe] = {0, 0, 0};
  std::map<const void *, size_t> smem_set; // kernel -> opted-in dynamic smem
  // host-pointer path: pinned ring for pageable callers, mapped buffer for tiny calls
  RingSlot ring[kRing];
  void *tiny_host = nullptr, *tiny_dev = nullptr; // one mapped (zero-copy) page-locked buffer
  size_t tiny_bytes = 0;
  cudaStream_t tiny_stream = nullptr;
  unsigned *tiny_counter = nullptr;               // device: CTAs of the tiny kernel that have finished
  unsigned tiny_epoch = 0;
};

std::map<int, std::unique_ptr<DeviceCtx>> g_ctx; // contexts never move: calls keep plain pointers

int get_ctx(DeviceCtx **out) {
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_map_mu);
  auto it = g_ctx.find(dev);
  if (it == g_ctx.end()) {
    std::unique_ptr<DeviceCtx> c(new DeviceCtx());
    c->device = dev;
    int v = 0;
    CU_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    c->sm_count = v;
    CU_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    c->smem_optin = (size_t)v;
    it = g_ctx.emplace(dev, std::move(c)).first;
  }
  *out = it->second.get();
  return 0;
}

void free_ctx(DeviceCtx &c) {
  for (auto &kv : c.tables) {
    cudaFree(kv.second.W);
    cudaFree(kv.second.Wd);
    cudaFree(kv.second.G);
    cudaFree(kv.second.Gfast);
    cudaFree(kv.second.TWfast);
    cudaFree(kv.second.TWbig);
    cudaFree(kv.second.Glong);
  }
  c.tables.clear();
  for (auto &kv : c.bluestein) {
    cudaFree(kv.second.chirp);
    cudaFree(kv.second.B);
  }
  c.bluestein.clear();
  cudaFree(c.tw1);
  cudaFree(c.tw2);
  cudaFree(c.tw1d);
  cudaFree(c.tw2d);
  cudaFree(c.tw_w1k);
  c.tw1 = c.tw2 = c.tw_w1k = nullptr;
  for (auto &kv : c.slice_plans) cudaFree(kv.second.starts);
  c.slice_plans.clear();
  cudaFree(c.hr2c_tables);
  c.hr2c_tables = nullptr;
  for (auto &kv : c.gfft_w) cudaFree(kv.second);
  c.gfft_w.clear();
  for (auto &kv : c.gfft_chirp) {
    cudaFree(kv.second.chirp);
    cudaFree(kv.second.Bf);
  }
  c.gfft_chirp.clear();
  c.tw1d = c.tw2d = nullptr;
  for (auto &kv : c.ws) free_ws(kv.second);
  c.ws.clear();
  c.smem_set.clear();
  for (int i = 0; i < kPipe; ++i) {
    if (c.pipe[i]) cudaStreamDestroy(c.pipe[i]);
    cudaFree(c.stage_in[i]);
    cudaFree(c.stage_out[i]);
    cudaFree(c.stage_tmp[i]);
    c.pipe[i] = nullptr;
    c.stage_in[i] = c.stage_out[i] = c.stage_tmp[i] = nullptr;
    c.cap_in[i] = c.cap_out[i] = c.cap_tmp[i] = 0;
  }
  for (RingSlot &r : c.ring) {
    if (r.st) cudaStreamDestroy(r.st);
    cudaFree(r.din);
    cudaFree(r.dout);
    cudaFree(r.dtmp);
    cudaFreeHost(r.hin);
    cudaFreeHost(r.hout);
    r = RingSlot();
  }
  cudaFreeHost(c.tiny_host);
  cudaFree(c.tiny_counter);
  c.tiny_counter = nullptr;
  c.tiny_host = c.tiny_dev = nullptr;
  c.tiny_bytes = 0;
  if (c.tiny_stream) cudaStreamDestroy(c.tiny_stream);
  c.tiny_stream = nullptr;
}

template <typename T> constexpr int dtype_id() { return std::is_same<T, float>::value ? 0 : 1; }

template <typename T> int get_tables(DeviceCtx &c, size_t N, Tables **out) {
  auto key = std::make_pair(N, dtype_id<T>());
  auto it = c.tables.find(key);
  if (it == c.tables.end()) {
    std::vector<cx<T>> w(N);
    std::vector<double2> wd(N);
    for (size_t k = 0; k < N; ++k) {
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)N;
      const long double cr = cosl(a), si = -sinl(a);
      w[k].x = (T)cr;
      w[k].y = (T)si;
      wd[k].x = (double)cr;
      wd[k].y = (double)si;
    }
    Tables t;
    CU_TRY(cudaMalloc(&t.W, N * sizeof(cx<T>)));
    CU_TRY(cudaMalloc((void **)&t.Wd, N * sizeof(double2)));
    CU_TRY(cudaMemcpy(t.W, w.data(), N * sizeof(cx<T>), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(t.Wd, wd.data(), N * sizeof(double2), cudaMemcpyHostToDevice));
    it = c.tables.emplace(key, t).first;
  }
  *out = &it->second;
  return 0;
}

template <typename T> int get_hilbert_table(DeviceCtx &c, size_t N, Tables &t, bool fast_layout = false) {
  void *&slot = fast_layout ? t.Gfast : t.G;
  if (slot) return 0;
  // full-spectrum multiplier of the Hilbert transformer, 1/n folded in
  // (/root/reference/include/fftconv/hilbert.hpp:39-51, :57)
  std::vector<cx<T>> g(N);
  for (size_t pos = 0; pos < N; ++pos) {
    const size_t k = fast_layout
                         ? fc::fast::freq_of_pos(fc::fast::pos_of_table_index((unsigned)pos, ilog2(N)), ilog2(N))
                         : fc::generic::dif_freq_of_pos((unsigned)pos, (unsigned)N);
    T im = 0;
    if (k != 0 && 2 * k != N) im = (T)((2 * k < N ? -1.0 : 1.0) / (double)N);
    g[pos].x = 0;
    g[pos].y = im;
  }
  CU_TRY(cudaMalloc(&slot, N * sizeof(cx<T>)));
  CU_TRY(cudaMemcpy(slot, g.data(), N * sizeof(cx<T>), cudaMemcpyHostToDevice));
  (void)c;
  return 0;
}

// in-place radix-2 FFT in long double (host, table construction only)
void host_fft_ld(std::vector<long double> &re, std::vector<long double> &im) {
  const size_t n = re.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      std::swap(re[i], re[j]);
      std::swap(im[i], im[j]);
    }
  }
  const long double pi = 3.14159265358979323846264338327950288L;
  std::vector<long double> cr(n / 2 + 1), ci(n / 2 + 1); // exp(-2 pi i j / n), evaluated once
  for (size_t j = 0; j < n / 2; ++j) {
    const long double a = -2.0L * pi * (long double)j / (long double)n;
    cr[j] = cosl(a);
    ci[j] = sinl(a);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const size_t stride = n / len;
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; ++k) {
        const long double wr = cr[k * stride], wi = ci[k * stride];
        const size_t u = i + k, v = i + k + len / 2;
        const long double tr = re[v] * wr - im[v] * wi, ti = re[v] * wi + im[v] * wr;
        re[v] = re[u] - tr;
        im[v] = im[u] - ti;
        re[u] += tr;
        im[u] += ti;
      }
  }
}

template <typename T> int get_bluestein(DeviceCtx &c, size_t n, BluesteinTables **out, bool long_layout = false) {
  auto key = std::make_pair(n, dtype_id<T>());
  auto it = c.bluestein.find(key);
  if (it == c.bluestein.end()) {
    size_t M = 1;
    while (M < 2 * n - 1) M *= 2;
    const long double pi = 3.14159265358979323846264338327950288L;
    std::vector<long double> cr(n), ci(n);
    std::vector<cx<T>> chirp(n);
    for (size_t j = 0; j < n; ++j) {
      const unsigned long long j2 = (unsigned long long)j * j % (2ULL * n); // angle reduced exactly
      const long double a = pi * (long double)j2 / (long double)n;
      cr[j] = cosl(a);
      ci[j] = -sinl(a);
      chirp[j].x = (T)cr[j];
      chirp[j].y = (T)ci[j];
    }
    std::vector<long double> br(M, 0.0L), bi(M, 0.0L);
    br[0] = cr[0];
    bi[0] = -ci[0];
    for (size_t j = 1; j < n; ++j) {
      br[j] = br[M - j] = cr[j];
      bi[j] = bi[M - j] = -ci[j];
    }
    host_fft_ld(br, bi);
    std::vector<cx<T>> B(M);
    const size_t nsub = M / 16;
    const int l2s = ilog2(M) - 4;
    for (size_t pos = 0; pos < M; ++pos) {
      // long layout (long_fft_kernels.cuh): entry (s, idx_sub) holds frequency s + 16 freq_sub(idx_sub)
      const size_t k = long_layout
                           ? pos / nsub + 16 * (size_t)fc::fast::freq_of_pos(
                                                   fc::fast::pos_of_table_index((unsigned)(pos % nsub), l2s), l2s)
                           : fc::generic::dif_freq_of_pos((unsigned)pos, (unsigned)M);
      B[pos].x = (T)(br[k] / (long double)M);
      B[pos].y = (T)(bi[k] / (long double)M);
    }
    BluesteinTables t;
    t.M = M;
    CU_TRY(cudaMalloc(&t.chirp, n * sizeof(cx<T>)));
    CU_TRY(cudaMalloc(&t.B, M * sizeof(cx<T>)));
    CU_TRY(cudaMemcpy(t.chirp, chirp.data(), n * sizeof(cx<T>), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(t.B, B.data(), M * sizeof(cx<T>), cudaMemcpyHostToDevice));
    it = c.bluestein.emplace(key, t).first;
  }
  *out = &it->second;
  return 0;
}

template <typename S>
int get_n2048_tables(DeviceCtx &c, const fc::n2048::CPairT<S> **tw1_out, const fc::n2048::CPairT<S> **tw2_out) {
  using CP = fc::n2048::CPairT<S>;
  CP **tw1, **tw2;
  if constexpr (std::is_same<S, float>::value) {
    tw1 = &c.tw1;
    tw2 = &c.tw2;
  } else {
    tw1 = &c.tw1d;
    tw2 = &c.tw2d;
  }
  if (!*tw1) {
    std::vector<CP> t1(2048), t2(128);
    fc::n2048::fill_tw1(t1.data());
    fc::n2048::fill_tw2(t2.data());
    CU_TRY(cudaMalloc((void **)tw1, t1.size() * sizeof(CP)));
    CU_TRY(cudaMalloc((void **)tw2, t2.size() * sizeof(CP)));
    CU_TRY(cudaMemcpy(*tw1, t1.data(), t1.size() * sizeof(CP), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(*tw2, t2.data(), t2.size() * sizeof(CP), cudaMemcpyHostToDevice));
  }
  *tw1_out = *tw1;
  *tw2_out = *tw2;
  return 0;
}

int grow(void **p, size_t *cap, size_t bytes) {
  if (*cap >= bytes) return 0;
  if (*p) CU_TRY(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  CU_TRY(cudaMalloc(p, bytes));
  *cap = bytes;
  return 0;
}

template <typename K> int ensure_smem(DeviceCtx &c, K kernel, size_t bytes) {
  const void *key = (const void *)kernel;
  auto it = c.smem_set.find(key);
  if (it != c.smem_set.end() && it->second >= bytes) return 0;
  if (bytes > c.smem_optin)
    return fail(FFTCONV_CUDA_ERR_UNSUPPORTED, "needs " + std::to_string(bytes) +
                                                 " B of shared memory per CTA, device allows " +
                                                 std::to_string(c.smem_optin));
  CU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  c.smem_set[key] = bytes;
  return 0;
}

bool is_device_ptr(const void *p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

// taps -> device (per-stream buffer) -> spectrum H in the layout the kernel wants
// A filter handle (fftconv_cuda_filter_*): taps + their spectrum, resident on one device.
struct Filter {
  int dtype = 0;
  int device = 0;
  size_t klen = 0, N = 0;
  int layout = 0;
  void *taps = nullptr;
  void *H = nullptr;
};
thread_local const Filter *tl_filter = nullptr; // set around a call made through a handle

// Envelope post-processing of the call in flight (fftconv_cuda_envelope_* / rf_envelope_*): set
// around entry(), read where the Hilbert kernels are launched.  dec == 0 means "plain hilbert".
struct EnvOpts {
  int log = 0;
  size_t dec = 0;
  double a = 0.0, b = 0.0;
};
thread_local EnvOpts tl_env;
template <typename T> fc::EnvEpilogue<T> current_epilogue(size_t n) {
  if (tl_env.dec == 0 || (tl_env.dec == 1 && !tl_env.log)) return fc::plain_epilogue<T>();
  // i / dec == umulhi(i, floor(2^32 / dec) + 1) for every i < n when n * dec < 2^32
  unsigned magic = 0;
  if (tl_env.dec > 1 && (double)n * (double)tl_env.dec < 4294967296.0)
    magic = (unsigned)(4294967296ULL / tl_env.dec) + 1u;
  return fc::EnvEpilogue<T>{0, tl_env.log, (int)tl_env.dec, magic, (T)tl_env.a, (T)tl_env.b};
}
size_t env_out_len(size_t n) { return tl_env.dec > 1 ? (n + tl_env.dec - 1) / tl_env.dec : n; }

template <typename T>
int prepare_spectrum(DeviceCtx &c, cudaStream_t st, const T *k, size_t klen, size_t N, int layout,
                     Tables &tab, const cx<T> **H_out) {
  if (tl_filter && tl_filter->dtype == dtype_id<T>() && tl_filter->N == N && tl_filter->layout == layout &&
      tl_filter->klen == klen && tl_filter->device == c.device && tl_filter->taps == (const void *)k) {
    *H_out = static_cast<const cx<T> *>(tl_filter->H); // spectrum reuse: no kernel
    return 0;
  }
  StreamWs &ws = c.ws[st];
  const T *taps_dev = k;
  if (!is_device_ptr(k)) {
    if (int r = grow(&ws.taps, &ws.taps_bytes, klen * sizeof(T))) return r;
    CU_TRY(cudaMemcpyAsync(ws.taps, k, klen * sizeof(T), cudaMemcpyHostToDevice, st));
    taps_dev = static_cast<const T *>(ws.taps);
  }
  if (int r = grow(&ws.H, &ws.h_bytes, N * sizeof(cx<T>))) return r;
  const int warps = fc::generic::kSpectrumWarps;
  fc::generic::spectrum_kernel<T><<<(unsigned)((N + warps - 1) / warps), warps * 32, 0, st>>>(
      taps_dev, (int)klen, (int)N, tab.Wd, layout, 1.0 / (double)N, static_cast<cx<T> *>(ws.H));
  CU_TRY(cudaGetLastError());
  ++g_launches;
  *H_out = static_cast<const cx<T> *>(ws.H);
  return 0;
}

// ---- fast engine (fast_fft.cuh): N = 512 .. 16384 ------------------------------------
template <typename T> constexpr int fast_fpc(int log2n) {
  // FFTs per CTA: 256 threads (float) / 128 threads (double) where one FFT has fewer
  const int target = sizeof(T) == 4 ? 256 : 128;
  const int threads = (1 << log2n) / 16;
  return threads >= target ? 1 : target / threads;
}
template <typename T> size_t fast_smem(int log2n, size_t klen) {
  const size_t n = size_t(1) << log2n;
  return (size_t)fast_fpc<T>(log2n) * (n + n / 16 + 2 * (klen ? klen - 1 : 0)) * sizeof(cx<T>);
}
template <typename T> bool fast_ok(const DeviceCtx &c, size_t N, size_t klen) {
  if (N < 512 || N > 16384 || env_int("FFTCONV_CUDA_FORCE_GENERIC", 0)) return false;
  if (sizeof(T) == 8 && N > 8192) return false; // 1024 threads x complex double does not fit the register file
  return fast_smem<T>(ilog2(N), klen) <= c.smem_optin;
}

// Packed float (D = f2: four real streams per FFT, 16-byte shared-memory points): 512 threads
// per CTA where one FFT has fewer, one CTA per SM, for FFT sizes <= 8192.
constexpr int packed_fpc(int log2n) {
  const int threads = (1 << log2n) / 16;
  return threads >= 512 ? 1 : 512 / threads;
}
size_t packed_smem(int log2n, size_t klen) {
  const size_t n = size_t(1) << log2n;
  return (size_t)packed_fpc(log2n) * (n + n / 16 + 2 * (klen ? klen - 1 : 0)) * sizeof(cx<fc::f2>);
}
template <typename T> bool packed_ok(const DeviceCtx &c, size_t N, size_t klen) {
  if (!std::is_same<T, float>::value || N < 512 || N > 8192) return false;
  if (env_int("FFTCONV_CUDA_FORCE_GENERIC", 0) || env_int("FFTCONV_CUDA_NO_PACKED", 0)) return false;
  return packed_smem(ilog2(N), klen) <= c.smem_optin;
}

template <typename T, int LOG2N>
int launch_oa_fast(DeviceCtx &c, cudaStream_t st, const fc::generic::OaArgs<T> &p, size_t klen) {
  if constexpr (std::is_same<T, float>::value && LOG2N <= 13) {
    // four streams per FFT pay off once the streams fill the machine; a few short signals are
    // latency-bound and better spread over more, smaller CTAs
    if (packed_ok<T>(c, size_t(1) << LOG2N, klen) && p.n_streams >= 2LL * packed_fpc(LOG2N) * c.sm_count) {
      constexpr int FPC = packed_fpc(LOG2N);
      const size_t smem = packed_smem(LOG2N, klen);
      auto kern = fc::fast::oa_fast_kernel<fc::f2, LOG2N, FPC>;
      if (int r = ensure_smem(c, kern, smem)) return r;
      const long long grid = (p.n_streams + 4 * FPC - 1) / (4 * FPC);
      if (grid > 0x7fffffffLL) return fail(FFTCONV_CUDA_ERR_UNSUPPORTED, "batch too large for one launch");
      kern<<<(unsigned)grid, fc::fast::Plan<LOG2N>::THREADS * FPC, smem, st>>>(p);
      CU_TRY(cudaGetLastError());
      ++g_launches;
      return 0;
    }
  }
  constexpr int FPC = fast_fpc<T>(LOG2N);
  const size_t smem = fast_smem<T>(LOG2N, klen);
  auto kern = fc::fast::oa_fast_kernel<T, LOG2N, FPC>;
  if (int r = ensure_smem(c, kern, smem)) return r;
  const long long grid = (p.n_streams + 2 * FPC - 1) / (2 * FPC);
  if (grid > 0x7fffffffLL) return fail(FFTCONV_CUDA_ERR_UNSUPPORTED, "batch too large for one launch");
  kern<<<(unsigned)grid, fc::fast::Plan<LOG2N>::THREADS * FPC, smem, st>>>(p);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}
template <typename T, int LOG2N>
int launch_hilbert_fast(DeviceCtx &c, cudaStream_t st, const fc::generic::HilbertArgs<T> &p) {
  if constexpr (std::is_same<T, float>::value && LOG2N <= 13) {
    if (packed_ok<T>(c, size_t(1) << LOG2N, 0) && p.batch >= 2LL * packed_fpc(LOG2N) * c.sm_count) {
      constexpr int FPC = packed_fpc(LOG2N);
      const size_t smem = packed_smem(LOG2N, 0);
      const bool aligned = (reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && (p.in_stride & 3) == 0;
      // experiment (FFTCONV_CUDA_HILBERT_2CTA=1, 8192-point lines): two rows per FFT in plain float and
      // TWO independent CTAs per SM, so that one CTA's shared-memory phases overlap the other's arithmetic
      if constexpr (LOG2N == 13) {
        if (aligned && p.epi.plain && env_int("FFTCONV_CUDA_HILBERT_2CTA", 0)) {
          auto pk2 = fc::fast::hilbert_pk_persistent_kernel<LOG2N, 1, true, float, 2>;
          const size_t smem2 = fc::fast::Plan<LOG2N>::SMEM_ELEMS * sizeof(cx<float>) + 16;
          if (int r = ensure_smem(c, pk2, smem2)) return r;
          const long long units = (p.batch + 1) / 2;
          pk2<<<(unsigned)std::min<long long>(units, 2LL * c.sm_count), fc::fast::Plan<LOG2N>::THREADS, smem2, st>>>(p);
          CU_TRY(cudaGetLastError());
          ++g_launches;
          return 0;
        }
      }
      if (aligned && !env_int("FFTCONV_CUDA_NO_PERSISTENT_HILBERT", 0)) {
        auto pk = p.epi.plain ? fc::fast::hilbert_pk_persistent_kernel<LOG2N, FPC, true>
                              : fc::fast::hilbert_pk_persistent_kernel<LOG2N, FPC, false>;
        if (int r = ensure_smem(c, pk, smem + 16)) return r;
        const long long units = (p.batch + 4 * FPC - 1) / (4 * FPC);
        pk<<<(unsigned)std::min<long long>(units, c.sm_count), fc::fast::Plan<LOG2N>::THREADS * FPC, smem + 16, st>>>(p);
        CU_TRY(cudaGetLastError());
        ++g_launches;
        return 0;
      }
      auto kern = p.epi.plain ? fc::fast::hilbert_fast_kernel<fc::f2, LOG2N, FPC, true>
                              : fc::fast::hilbert_fast_kernel<fc::f2, LOG2N, FPC, false>;
      if (int r = ensure_smem(c, kern, smem)) return r;
      const long long grid = (p.batch + 4 * FPC - 1) / (4 * FPC);
      kern<<<(unsigned)grid, fc::fast::Plan<LOG2N>::THREADS * FPC, smem, st>>>(p);
      CU_TRY(cudaGetLastError());
      ++g_launches;
      return 0;
    }
  }
  constexpr int FPC = fast_fpc<T>(LOG2N);
  const size_t smem = fast_smem<T>(LOG2N, 0);
  auto kern = p.epi.plain ? fc::fast::hilbert_fast_kernel<T, LOG2N, FPC, true>
                          : fc::fast::hilbert_fast_kernel<T, LOG2N, FPC, false>;
  if (int r = ensure_smem(c, kern, smem)) return r;
  const long long grid = (p.batch + 2 * FPC - 1) / (2 * FPC);
  kern<<<(unsigned)grid, fc::fast::Plan<LOG2N>::THREADS * FPC, smem, st>>>(p);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}
template <typename T, int LOG2N> int build_fast_twiddles(Tables &t) {
  using P = fc::fast::Plan<LOG2N>;
  std::vector<fc::fast::cx<T>> tw(P::TW_ELEMS);
  fc::fast::fill_pass_twiddles<LOG2N, T>(tw.data(), [](long long num, long long den) {
    const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)(num % den) / (long double)den;
    return fc::fast::cx<T>{(T)cosl(a), (T)(-sinl(a))};
  });
  CU_TRY(cudaMalloc(&t.TWfast, tw.size() * sizeof(fc::fast::cx<T>)));
  CU_TRY(cudaMemcpy(t.TWfast, tw.data(), tw.size() * sizeof(fc::fast::cx<T>), cudaMemcpyHostToDevice));
  return 0;
}

#define FAST_SWITCH(FN, log2n, ...)                                        \
  switch (log2n) {                                                         \
  case 9: return FN<T, 9>(__VA_ARGS__);                                    \
  case 10: return FN<T, 10>(__VA_ARGS__);                                  \
  case 11: return FN<T, 11>(__VA_ARGS__);                                  \
  case 12: return FN<T, 12>(__VA_ARGS__);                                  \
  case 13: return FN<T, 13>(__VA_ARGS__);                                  \
  case 14: return FN<T, 14>(__VA_ARGS__);                                  \
  default: return fail(FFTCONV_CUDA_ERR_UNSUPPORTED, "fast engine: bad size"); \
  }

// Is the dedicated N = 2048 kernel (oa_n2048_kernel: 16 points per thread, 128-thread groups) in use?
// (Round 1 also carried a warp-pair engine, oa_w2: measured 12 % slower, removed -- DESIGN.md 3.6.)
int oa2048_engine() {
  if (env_int("FFTCONV_CUDA_FORCE_GENERIC", 0) || env_int("FFTCONV_CUDA_NO_N2048", 0)) return 0;
  return 1;
}

// engine for one call: the dedicated kernel takes N = 2048, 8 <= K <= 410 (staging area and tail
// buffers are sized for that) and, in its WIDE instantiations, up to 1024 taps in float32 / 640 in
// float64 -- what three groups' tail buffers leave room for; those are reached through the
// throughput rule only, the table gives such kernels larger segments
template <typename T> size_t oa2048_max_taps() {
  if (env_int("FFTCONV_CUDA_NO_N2048_WIDE", 0)) return fc::n2048::kMaxTapsNarrow;
  return std::is_same<T, float>::value ? (size_t)fc::n2048::kMaxTapsWide : 640;
}
template <typename T> int oa2048_engine_for(size_t N, size_t klen) {
  if (N != 2048 || klen < 8 || klen > oa2048_max_taps<T>()) return 0;
  const int e = oa2048_engine();
  if (!std::is_same<T, float>::value) return (e && !env_int("FFTCONV_CUDA_NO_N2048_F64", 0)) ? 1 : 0;
  return e;
}

// The warp-per-FFT engine (oa_w1k_*.cuh): float32, 1024-point blocks, 8 <= K <= w1k::kMaxTaps.  It serves
// the table's own N = 1024 row and, through the throughput rule, large jobs with shorter or longer
// kernels (FFTCONV_CUDA_NO_W1K=1 switches it off; FFTCONV_CUDA_W1K_MAX_TAPS moves the upper end of
// what the throughput rule sends to it: 360 taps, where the 2048-point kernel takes over -- measured on 4096 x 65536,
// sustained: K = 340 0.592 against 0.615 ms, K = 360 0.606 / 0.622, K = 400 0.636 / 0.626; profiles/r02_w1k_tap_limit.jsonl).
template <typename T> bool w1k_engine_for(size_t N, size_t klen) {
  if (!std::is_same<T, float>::value || N != 1024) return false;
  if (klen < 8 || klen > (size_t)fc::w1k::kMaxTaps) return false;
  return !(env_int("FFTCONV_CUDA_FORCE_GENERIC", 0) || env_int("FFTCONV_CUDA_NO_W1K", 0));
}
size_t w1k_rule_max_taps() { return (size_t)std::min(fc::w1k::kMaxTaps, std::max(0, env_int("FFTCONV_CUDA_W1K_MAX_TAPS", 360))); }

// The plan of one launch of the N = 2048 kernel: stream geometry + where each group's slice starts.
template <typename T>
int get_slice_plan(DeviceCtx &c, fc::n2048::OaParamsT<T> &p, const T *a, T *out, size_t batch, size_t n, size_t a_stride,
                   size_t out_stride, size_t out_len, size_t klen, int mode, long long n_slots) {
  const std::array<long long, 5> key = {dtype_id<T>(), (long long)batch, (long long)n, (long long)klen, n_slots};
  auto it = c.slice_plans.find(key);
  if (it == c.slice_plans.end()) {
    if (c.slice_plans.size() >= 512) { // shapes keep changing: start over (tables may still be read by kernels in flight)
      CU_TRY(cudaDeviceSynchronize());
      for (auto &kv : c.slice_plans) cudaFree(kv.second.starts);
      c.slice_plans.clear();
    }
    const fc::n2048::SlicePlan plan = fc::n2048::plan_slices<T>(p, a, out, (long long)batch, (long long)n, (long long)a_stride,
                                                               (long long)out_stride, (long long)out_len, (int)klen, mode, n_slots);
    DeviceCtx::SlicePlanDev d;
    d.vrows = plan.vrows; d.vlen = plan.vlen; d.n_streams = plan.n_streams; d.n_quads = plan.n_quads;
    CU_TRY(cudaMalloc((void **)&d.starts, plan.starts.size() * sizeof(fc::n2048::SlotStart)));
    CU_TRY(cudaMemcpy(d.starts, plan.starts.data(), plan.starts.size() * sizeof(fc::n2048::SlotStart), cudaMemcpyHostToDevice));
    it = c.slice_plans.emplace(key, d).first;
  }
  const DeviceCtx::SlicePlanDev &d = it->second;
  p.in = a; p.out = out;
  p.batch = (long long)batch; p.n = (long long)n;
  p.in_stride = (long long)a_stride; p.out_stride = (long long)out_stride; p.out_len = (long long)out_len;
  p.klen = (int)klen;
  p.step = fc::n2048::kN - ((int)klen - 1);
  p.pad = mode == FFTCONV_CUDA_SAME ? (int)(klen / 2) : 0;
  p.vrows = d.vrows; p.vlen = d.vlen; p.n_streams = d.n_streams; p.n_quads = d.n_quads;
  p.n_slots = n_slots;
  p.slots = d.starts;
  return 0;
}

// ---- overlap-add on device pointers -------------------------------------------------
template <typename T>
int oa_device(DeviceCtx &c, cudaStream_t st, const T *a, size_t batch, size_t n, size_t a_stride,
              const T *k, size_t klen, T *out, size_t out_len, size_t out_stride, int mode,
              size_t N) {
  Tables *tab = nullptr;
  if (int r = get_tables<T>(c, N, &tab)) return r;
  const size_t step = N - (klen - 1);
  const long long segs = (long long)((n + step - 1) / step);

  if constexpr (std::is_same<T, float>::value) {
    if (w1k_engine_for<T>(N, klen)) {
      if (!c.tw_w1k) {
        std::vector<fc::n2048::cpair> tw(fc::w1k::kTwiddleElems);
        fc::w1k::build_twiddles(tw.data());
        CU_TRY(cudaMalloc((void **)&c.tw_w1k, tw.size() * sizeof(fc::n2048::cpair)));
        CU_TRY(cudaMemcpy(c.tw_w1k, tw.data(), tw.size() * sizeof(fc::n2048::cpair), cudaMemcpyHostToDevice));
      }
      // The spectrum: resident in a filter handle, else transformed by the kernel itself (one warp per CTA runs the
      // forward half of a unit on the taps: no separate spectrum kernel, no launch gap).  FFTCONV_CUDA_W1K_SPECTRUM_KERNEL=1
      // restores the separate kernel (double-precision DFT of the taps) for A/B measurements.
      const cx<float> *H = nullptr;
      const float *taps_dev = nullptr;
      const bool handle = tl_filter && tl_filter->dtype == dtype_id<float>() && tl_filter->N == N &&
                          tl_filter->layout == fc::generic::LAYOUT_W1K && tl_filter->klen == klen &&
                          tl_filter->device == c.device && tl_filter->taps == (const void *)k;
      if (handle || env_int("FFTCONV_CUDA_W1K_SPECTRUM_KERNEL", 0)) {
        if (int r = prepare_spectrum<float>(c, st, k, klen, N, fc::generic::LAYOUT_W1K, *tab, &H)) return r;
      } else {
        taps_dev = k;
        if (!is_device_ptr(k)) {
          StreamWs &ws = c.ws[st];
          if (int r = grow(&ws.taps, &ws.taps_bytes, klen * sizeof(float))) return r;
          CU_TRY(cudaMemcpyAsync(ws.taps, k, klen * sizeof(float), cudaMemcpyHostToDevice, st));
          taps_dev = static_cast<const float *>(ws.taps);
        }
      }
      fc::w1k::W1kParams p{};
      fc::w1k::fill_params(p, a, out, (long long)batch, (long long)n, (long long)a_stride, (long long)out_stride,
                           (long long)out_len, (int)klen, mode == FFTCONV_CUDA_SAME ? 1 : 0);
      p.tw = c.tw_w1k;
      p.hs = reinterpret_cast<const fc::n2048::cpair *>(H);
      p.taps = taps_dev;
      p.cta_major = env_int("FFTCONV_CUDA_W1K_CTA_MAJOR", 0);
      if (fc::w1k::launch_smem_bytes() > c.smem_optin)
        return fail(FFTCONV_CUDA_ERR_UNSUPPORTED, "warp-per-FFT kernel: not enough shared memory per CTA on this device");
      const int sms_avail = std::max(1, c.sm_count - std::max(0, g_reserved_sms.load()));
      const int wpc = fc::w1k::warps_per_cta();
      const long long ctas = std::max<long long>(1, std::min<long long>(sms_avail, (p.n_units + wpc - 1) / wpc));
      // FFTCONV_CUDA_W1K_POOL_PCT=p (default 0: fixed ranges): the last p per cent of the units of a large job are
      // handed out one at a time (W1kParams::sched); the counters live with the stream, whose launches follow each
      // other.  Measured (profiles/r02_w1k_session4_ab.jsonl): the SMs' active cycles differ by 4 % under fixed
      // ranges, yet p = 5 .. 50 changes nothing (+-0.3 %) -- the sustained rate is set by the 1 kW power cap, and
      // an SM that idles at the end of a launch lends its power to the others.  Kept as an A/B switch.
      const long long n_warps = ctas * wpc;
      const int pool_pct = std::min(90, std::max(0, env_int("FFTCONV_CUDA_W1K_POOL_PCT", 0)));
      if (pool_pct > 0 && p.n_units >= 16 * n_warps) {
        StreamWs &ws = c.ws[st];
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (!ws.sched && cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
          CU_TRY(cudaMalloc((void **)&ws.sched, 2 * sizeof(unsigned)));
          CU_TRY(cudaMemset(ws.sched, 0, 2 * sizeof(unsigned)));
        }
        if (ws.sched) {
          p.sched = ws.sched;
          p.static_per_warp = std::max<long long>(1, p.n_units * (100 - pool_pct) / 100 / n_warps);
          p.pool_first = p.static_per_warp * n_warps;
        }
      }
      if (const int e = fc::w1k::launch(p, (int)ctas, st))
        return fail(FFTCONV_CUDA_ERR_CUDA, std::string("oa_w1k_kernel: ") + cudaGetErrorString((cudaError_t)e));
      ++g_launches;
      return 0;
    }
  }
  const int engine2048 = oa2048_engine_for<T>(N, klen);
  if (engine2048 == 1) {
    using D = typename fc::DataOf<T>::D;
    constexpr int NL = fc::DataTraits<D>::NL;
    const fc::n2048::CPairT<T> *tw1 = nullptr, *tw2 = nullptr;
    if (int r = get_n2048_tables<T>(c, &tw1, &tw2)) return r;
    const cx<T> *H = nullptr;
    if (int r = prepare_spectrum<T>(c, st, k, klen, N, fc::generic::LAYOUT_N2048, *tab, &H)) return r;
    int G = env_int("FFTCONV_CUDA_N2048_GROUPS", 4);
    if (G < 1 || G > 4) G = 4;
    const bool want_recomp = env_int("FFTCONV_CUDA_TW_RECOMPUTE", 1) != 0;
    // float64 with the longest kernels (K > ~390: two 16-byte tail buffers of K - 1 points per
    // group) does not fit four groups: drop to three
    const bool wide_k = klen > (size_t)fc::n2048::kMaxTapsNarrow || env_int("FFTCONV_CUDA_N2048_FORCE_WIDE", 0);
    if (wide_k && G < 3) G = 4; // the wide instantiations exist for four and three groups, always with twiddle recomputation
    while (G > (wide_k ? 3 : 1) && fc::n2048::smem_bytes<T>(G, (int)klen, wide_k || (want_recomp && G == 4)) > c.smem_optin) --G;
    // a caller that runs a collective beside this kernel (the K-1 tail exchange of a long signal)
    // asks for a few SMs to be left free: a persistent grid that owns every SM would make the
    // collective's kernel wait for it, or wait for the collective's kernel on the SM it took
    const int sms = std::max(1, c.sm_count - std::max(0, g_reserved_sms.load()));
    // the grid: one CTA per SM unless there is less than one block per group to hand out
    const long long blocks_total =
        (long long)((batch + NL - 1) / NL) * (long long)((n + (N - klen)) / (N - klen + 1));
    const long long grid = std::max<long long>(1, std::min<long long>(sms, (blocks_total + G - 1) / G));
    fc::n2048::OaParamsT<T> p{};
    if (int r = get_slice_plan<T>(c, p, a, out, batch, n, a_stride, out_stride, out_len, klen, mode, grid * G)) return r;
    p.tw1 = tw1;
    p.tw2 = tw2;
    p.hs = reinterpret_cast<const fc::n2048::CPairT<T> *>(H);
    const bool wide = wide_k;
    const bool recomp = wide || (want_recomp && G == 4);
    const size_t smem = fc::n2048::smem_bytes<T>(G, (int)klen, recomp);
#define FC_LAUNCH_N2048(GG, RR, WW)                                                            \
  {                                                                                            \
    if (int r = ensure_smem(c, fc::n2048::oa_n2048_kernel<D, GG, RR, WW>, smem)) return r;     \
    fc::n2048::oa_n2048_kernel<D, GG, RR, WW><<<(unsigned)grid, 128 * GG, smem, st>>>(p);      \
  }
    if (wide) {
      if (G == 4) FC_LAUNCH_N2048(4, true, true)
      else FC_LAUNCH_N2048(3, true, true)
    } else if (G == 4 && recomp) FC_LAUNCH_N2048(4, true, false)
    else if (G == 4) FC_LAUNCH_N2048(4, false, false)
    else if (G == 3) FC_LAUNCH_N2048(3, false, false)
    else if (G == 2) FC_LAUNCH_N2048(2, false, false)
    else FC_LAUNCH_N2048(1, false, false)
#undef FC_LAUNCH_N2048
    CU_TRY(cudaGetLastError());
    ++g_launches;
    return 0;
  }

  const bool use_fast = fast_ok<T>(c, N, klen);
  const cx<T> *H = nullptr;
  if (int r = prepare_spectrum<T>(c, st, k, klen, N,
                                  use_fast ? fc::generic::LAYOUT_FAST : fc::generic::LAYOUT_DIF, *tab, &H))
    return r;
  fc::generic::OaArgs<T> p{};
  p.in = a;
  p.out = out;
  p.batch = (long long)batch;
  p.n = (long long)n;
  p.in_stride = (long long)a_stride;
  p.out_stride = (long long)out_stride;
  p.out_len = (long long)out_len;
  p.klen = (int)klen;
  p.log2n = ilog2(N);
  p.step = (int)step;
  p.pad = mode == FFTCONV_CUDA_SAME ? (int)(klen / 2) : 0;
  p.segs = segs;
  const long long target_streams = 8LL * c.sm_count;
  long long chunks = 1;
  if ((long long)batch < target_streams) {
    chunks = (target_streams + (long long)batch - 1) / (long long)batch;
    // >= 8 segments per chunk keeps the warm-up segment under ~12 %; when even that
    // cannot fill a quarter of the SMs (single short signals) go down to 2
    long long min_segs = 8;
    if ((long long)batch * std::max<long long>(1, segs / 8) < c.sm_count / 4) min_segs = 2;
    chunks = std::min(chunks, std::max<long long>(1, segs / min_segs));
  }
  p.segs_per_chunk = (segs + chunks - 1) / chunks;
  p.chunks = (segs + p.segs_per_chunk - 1) / p.segs_per_chunk;
  p.n_streams = (long long)batch * p.chunks;
  p.W = static_cast<const cx<T> *>(tab->W);
  p.H = H;
  if (use_fast) {
    if (!tab->TWfast) {
      auto build = [&]() -> int { FAST_SWITCH(build_fast_twiddles, p.log2n, *tab) };
      if (int r = build()) return r;
    }
    p.W = static_cast<const cx<T> *>(tab->TWfast);
    FAST_SWITCH(launch_oa_fast, p.log2n, c, st, p, klen)
  }
  const size_t smem = (N + 2 * (klen - 1)) * sizeof(cx<T>);
  if (int r = ensure_smem(c, fc::generic::oa_generic_kernel<T>, smem)) return r;
  const long long grid = (p.n_streams + 1) / 2;
  if (grid > 0x7fffffffLL) return fail(FFTCONV_CUDA_ERR_UNSUPPORTED, "batch too large for one launch");
  const int threads = (int)std::min<size_t>(512, std::max<size_t>(32, N / 4));
  fc::generic::oa_generic_kernel<T><<<(unsigned)grid, threads, smem, st>>>(p);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

template <typename T> size_t max_generic_fft(const DeviceCtx &c, size_t klen) {
  size_t N = 1;
  while ((2 * N + 2 * (klen - 1)) * sizeof(cx<T>) <= c.smem_optin) N *= 2;
  return N;
}

// ---- one FFT longer than a CTA: N_big = 16 N_sub through a global scratch ----------------
template <typename T> bool long_ok(const DeviceCtx &c, size_t Nbig) {
  if (Nbig < 16 * 512 || env_int("FFTCONV_CUDA_FORCE_GENERIC", 0)) return false;
  const size_t nsub = Nbig / 16;
  if (nsub > 16384 || (sizeof(T) == 8 && nsub > 8192)) return false;
  return fast_smem<T>(ilog2(nsub), 0) <= c.smem_optin;
}

template <typename T, int LOG2N>
int launch_inner(DeviceCtx &c, cudaStream_t st, cx<T> *scratch, long long n_subs, const cx<T> *W, const cx<T> *Hbig) {
  constexpr int FPC = fast_fpc<T>(LOG2N);
  const size_t smem = fast_smem<T>(LOG2N, 0);
  auto kern = fc::longfft::inner_kernel<T, LOG2N, FPC>;
  if (int r = ensure_smem(c, kern, smem)) return r;
  const long long grid = (n_subs + FPC - 1) / FPC;
  kern<<<(unsigned)grid, fc::fast::Plan<LOG2N>::THREADS * FPC, smem, st>>>(
      reinterpret_cast<fc::fast::cx<T> *>(scratch), n_subs, reinterpret_cast<const fc::fast::cx<T> *>(W),
      reinterpret_cast<const fc::fast::cx<T> *>(Hbig));
  CU_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

// Hbig: spectrum table of N_big entries in LAYOUT_LONG order (1/N_big folded in)
template <typename T>
int long_device(DeviceCtx &c, cudaStream_t st, const T *a, size_t batch, size_t n, size_t a_stride, T *out,
                size_t out_len, size_t out_stride, int pad, size_t Nbig, const cx<T> *Hbig, bool hilbert,
                const cx<T> *chirp = nullptr /* chirp-z Hilbert: Hbig is the chirp filter's spectrum */) {
  const size_t nsub = Nbig / 16;
  const int l2s = ilog2(nsub);
  Tables *tbig = nullptr, *tsub = nullptr;
  if (int r = get_tables<T>(c, Nbig, &tbig)) return r;
  if (int r = get_tables<T>(c, nsub, &tsub)) return r;
  if (!tsub->TWfast) {
    auto build = [&]() -> int { FAST_SWITCH(build_fast_twiddles, l2s, *tsub) };
    if (int r = build()) return r;
  }
  if (!tbig->TWbig) {
    std::vector<cx<T>> tw(15 * nsub);
    for (size_t s = 1; s < 16; ++s)
      for (size_t p = 0; p < nsub; ++p) {
        const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)((p * s) % Nbig) / (long double)Nbig;
        tw[(s - 1) * nsub + p].x = (T)cosl(ang);
        tw[(s - 1) * nsub + p].y = (T)(-sinl(ang));
      }
    CU_TRY(cudaMalloc(&tbig->TWbig, tw.size() * sizeof(cx<T>)));
    CU_TRY(cudaMemcpy(tbig->TWbig, tw.data(), tw.size() * sizeof(cx<T>), cudaMemcpyHostToDevice));
  }
  const size_t pairs_total = (batch + 1) / 2;
  const size_t pair_bytes = Nbig * sizeof(cx<T>);
  const size_t budget = (size_t)env_int("FFTCONV_CUDA_SCRATCH_MB", 512) << 20;
  const size_t bufs = chirp ? 2 : 1; // chirp-z keeps the sequence between its two convolutions
  size_t pairs_per = std::max<size_t>(1, std::min<size_t>(pairs_total, budget / (bufs * pair_bytes)));
  pairs_per = std::min<size_t>(pairs_per, 65535);
  // per stream: launches on different streams -- the host pipeline's three, or two
  // user streams -- must not write the same intermediate; work on ONE stream is ordered
  StreamWs &wsl = c.ws[st];
  if (wsl.scratch_bytes < bufs * pairs_per * pair_bytes) CU_TRY(cudaStreamSynchronize(st)); // before the old one is freed
  if (int r = grow(&wsl.scratch, &wsl.scratch_bytes, bufs * pairs_per * pair_bytes)) return r;
  fc::longfft::LongArgs<T> la{};
  la.in = a;
  la.out = out;
  la.batch = (long long)batch;
  la.n = (long long)n;
  la.in_stride = (long long)a_stride;
  la.out_stride = (long long)out_stride;
  la.out_len = (long long)out_len;
  la.pad = pad;
  la.log2sub = l2s;
  la.scratch = reinterpret_cast<fc::fast::cx<T> *>(wsl.scratch);
  la.TWbig = reinterpret_cast<const fc::fast::cx<T> *>(tbig->TWbig);
  la.hilbert = hilbert ? 1 : 0;
  la.epi = current_epilogue<T>(n);
  la.bluestein = 0;
  la.chirp = reinterpret_cast<const fc::fast::cx<T> *>(chirp);
  la.cbuf = reinterpret_cast<fc::fast::cx<T> *>(wsl.scratch) + pairs_per * Nbig;
  const int threads = 256;
  for (size_t p0 = 0; p0 < pairs_total; p0 += pairs_per) {
    const size_t np = std::min(pairs_per, pairs_total - p0);
    la.pair0 = (long long)p0;
    la.pairs = (long long)np;
    const dim3 grid((unsigned)((nsub + threads - 1) / threads), (unsigned)np);
    auto inner = [&]() -> int {
      FAST_SWITCH(launch_inner, l2s, c, st, static_cast<cx<T> *>(wsl.scratch), (long long)np * 16,
                  static_cast<const cx<T> *>(tsub->TWfast), Hbig)
    };
    for (int pass = 0; pass < (chirp ? 2 : 1); ++pass) {
      la.bluestein = chirp ? pass + 1 : 0;
      fc::longfft::outer_fwd_kernel<T><<<grid, threads, 0, st>>>(la);
      CU_TRY(cudaGetLastError());
      ++g_launches;
      if (int r = inner()) return r;
      fc::longfft::outer_inv_kernel<T><<<grid, threads, 0, st>>>(la);
      CU_TRY(cudaGetLastError());
      ++g_launches;
    }
  }
  return 0;
}

enum Op { OP_OA, OP_CONV, OP_HILBERT, OP_RF /* overlap-add (Same) then Hilbert envelope */ };

// Segment FFT size for overlap-add.  The reference's table (fftconv.hpp:38-66) is a CPU cost
// model, N (log2 N + 1) / (N - K + 1) per output sample; on this machine a segment also has a
// fixed cost (barriers, exposed latencies, the fraction of a CTA it keeps busy) that the model
// does not see, and tiny segments are slower by an order of magnitude (measured, DESIGN.md
// section 4: K = 20 -> N = 64 takes 4.7 ms at 4096 x 65536, the N = 2048 kernel 0.58 ms).  The
// table therefore stays the rule for small jobs -- where it also decides the latency of one
// short call -- and for every K it was measured best; for large jobs short kernels are moved to
// the size whose kernel is fastest.  Any N > K - 1 computes the same linear convolution.
// FFTCONV_CUDA_STRICT_TABLE=1 restores the reference's choice, FFTCONV_CUDA_OA_FFT_SIZE forces one.
template <typename T> size_t throughput_oa_size(const DeviceCtx &c, size_t klen, size_t batch, size_t n, size_t table_n) {
  if (env_int("FFTCONV_CUDA_STRICT_TABLE", 0)) return table_n;
  const int forced = env_int("FFTCONV_CUDA_OA_FFT_SIZE", 0);
  if (forced >= 16 && (forced & (forced - 1)) == 0 && (size_t)forced >= 2 * klen && (size_t)forced >= table_n / 64)
    return (size_t)forced;
  if ((double)batch * (double)n < 4194304.0) return table_n;
  (void)c;
  if (std::is_same<T, float>::value) {
    if (klen >= 8 && klen <= w1k_rule_max_taps() && w1k_engine_for<T>(1024, klen)) return 1024;
    if (klen >= 8 && klen < 221 && oa2048_engine() == 1) return 2048;
    if (klen < 8) return 512;
  } else {
    if (klen < 221) return 2048;                       // measured: 0.68-0.72 ms against 0.79-0.92 (N = 1024)
  }
  // 411 .. 1024 (float64: 640) taps: 2048-point segments on the dedicated kernel rather than the
  // table's 4096 / 8192 points on the register-blocked engine (measured)
  if (klen > 410 && klen <= oa2048_max_taps<T>() && oa2048_engine() != 0 && oa2048_engine_for<T>(2048, klen)) return 2048;
  if (!std::is_same<T, float>::value && table_n == 8192 && klen <= 1024)
    return 4096;  // 1.15 ms against 1.44: the 8192-point float64 CTA is register-starved
  return table_n;
}

// Kernels longer than a segment FFT can hold (N > max_generic_fft): the taps are cut into blocks of
// Kb taps, every block is convolved with segment FFTs of 4 Kb points (75 % of each segment is new
// output) into a per-stream temporary, and the partial results are summed in block order.  The
// reference handles such kernels with one ever larger FFT (fftconv.hpp:61-65); the linear
// convolution is the same.
template <typename T>
int partitioned_device(DeviceCtx &c, cudaStream_t st, const T *a, size_t batch, size_t n, size_t a_stride, const T *k,
                       size_t klen, T *out, size_t out_len, size_t out_stride, int mode) {
  const size_t Kb = std::is_same<T, float>::value ? 4096 : 2048, N = 4 * Kb;
  if (!fast_ok<T>(c, N, Kb)) return fail(FFTCONV_CUDA_ERR_UNSUPPORTED, "long kernel: block FFT does not fit this device");
  const size_t part_max = n + Kb - 1;
  StreamWs &ws = c.ws[st];
  if (ws.part_bytes < batch * part_max * sizeof(T)) CU_TRY(cudaStreamSynchronize(st));
  if (int r = grow(&ws.part, &ws.part_bytes, batch * part_max * sizeof(T))) return r;
  T *part = static_cast<T *>(ws.part);
  const long long pad = mode == FFTCONV_CUDA_SAME ? (long long)(klen / 2) : 0;
  const long long total = (long long)batch * (long long)out_len;
  const int threads = 256;
  if ((total + threads - 1) / threads > 0x7fffffffLL) return fail(FFTCONV_CUDA_ERR_UNSUPPORTED, "batch too large for one launch");
  for (size_t p0 = 0, blk = 0; p0 < klen; p0 += Kb, ++blk) {
    const size_t kp = std::min(Kb, klen - p0), plen = n + kp - 1;
    if (int r = oa_device<T>(c, st, a, batch, n, a_stride, k + p0, kp, part, plen, plen, FFTCONV_CUDA_FULL, N)) return r;
    fc::generic::partition_accumulate_kernel<T><<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(
        part, (long long)batch, (long long)plen, out, (long long)out_len, (long long)out_stride, (long long)p0 - pad,
        blk == 0 ? 1 : 0);
    CU_TRY(cudaGetLastError());
    ++g_launches;
  }
  return 0;
}

// Short kernels: direct time-domain evaluation (direct_kernels.cuh), one launch, no spectrum.
// Measured at 4096 x 65536 float32 / 2048 x 65536 float64, Full, direct
// against the FFT path: K <= 8: 0.42 / 0.44 ms against 0.53-0.89 / 0.54-0.68; 9..16: 0.47 / 0.54
// against 0.51-0.53 / 0.53-0.54; 17..24: 0.53 / 0.68 against 0.50-0.53 / 0.53-0.54.  So float32 goes
// direct up to 16 taps and float64 up to 8 (FFTCONV_CUDA_DIRECT_MAX_TAPS overrides, at most 16).
template <typename T> size_t direct_max_taps() {
  const int dflt = std::is_same<T, float>::value ? 16 : 8;
  return (size_t)std::min(16, std::max(0, env_int("FFTCONV_CUDA_DIRECT_MAX_TAPS", dflt)));
}
template <typename T>
int direct_device(DeviceCtx &c, cudaStream_t st, const T *a, size_t batch, size_t n, size_t a_stride, const T *k,
                  size_t klen, T *out, size_t out_len, size_t out_stride, int mode) {
  StreamWs &ws = c.ws[st];
  const T *taps_dev = k;
  if (!is_device_ptr(k)) {
    if (int r = grow(&ws.taps, &ws.taps_bytes, klen * sizeof(T))) return r;
    CU_TRY(cudaMemcpyAsync(ws.taps, k, klen * sizeof(T), cudaMemcpyHostToDevice, st));
    taps_dev = static_cast<const T *>(ws.taps);
  }
  fc::direct::DirectArgs<T> p{};
  p.in = a;
  p.out = out;
  p.taps = taps_dev;
  p.batch = (long long)batch;
  p.n = (long long)n;
  p.in_stride = (long long)a_stride;
  p.out_stride = (long long)out_stride;
  p.out_len = (long long)out_len;
  p.klen = (int)klen;
  p.pad = mode == FFTCONV_CUDA_SAME ? (int)(klen / 2) : 0;
  p.tiles_per_row = ((long long)out_len + fc::direct::kTile - 1) / fc::direct::kTile;
  const long long total = p.batch * p.tiles_per_row;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(total, 8LL * c.sm_count));
  if (klen <= 8) fc::direct::direct_conv_kernel<T, 8><<<grid, fc::direct::kThreads, 0, st>>>(p);
  else fc::direct::direct_conv_kernel<T, 16><<<grid, fc::direct::kThreads, 0, st>>>(p);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

// Small jobs (at most 4 Mi multiply-adds, at most 256 taps): the definition, one thread per output,
// one launch (direct_kernels.cuh, direct_small_kernel) -- config 1 is such a job.
template <typename T>
int direct_small_device(DeviceCtx &c, cudaStream_t st, const T *a, size_t batch, size_t n, size_t a_stride, const T *k,
                        size_t klen, T *out, size_t out_len, size_t out_stride, int mode) {
  fc::direct::SmallArgs<T> p{};
  p.in = a;
  p.out = out;
  p.batch = (long long)batch;
  p.n = (long long)n;
  p.in_stride = (long long)a_stride;
  p.out_stride = (long long)out_stride;
  p.out_len = (long long)out_len;
  p.klen = (int)klen;
  p.pad = mode == FFTCONV_CUDA_SAME ? (int)(klen / 2) : 0;
  if (is_device_ptr(k)) {
    p.taps = k;
  } else if (klen <= (size_t)fc::direct::kInlineTaps) {
    p.inline_taps = 1;
    std::memcpy(p.tap, k, klen * sizeof(T));
  } else {
    StreamWs &ws = c.ws[st];
    if (int r = grow(&ws.taps, &ws.taps_bytes, klen * sizeof(T))) return r;
    CU_TRY(cudaMemcpyAsync(ws.taps, k, klen * sizeof(T), cudaMemcpyHostToDevice, st));
    p.taps = static_cast<const T *>(ws.taps);
  }
  const long long total = p.batch * p.out_len;
  fc::direct::direct_small_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

template <typename T>
int conv_device(DeviceCtx &c, cudaStream_t st, Op op, const T *a, size_t batch, size_t n,
                size_t a_stride, const T *k, size_t klen, T *out, size_t out_len,
                size_t out_stride, int mode) {
  const bool fft_forced = env_int("FFTCONV_CUDA_NO_DIRECT", 0) || env_int("FFTCONV_CUDA_STRICT_TABLE", 0) ||
                          env_int("FFTCONV_CUDA_FORCE_GENERIC", 0) || env_int("FFTCONV_CUDA_OA_FFT_SIZE", 0) ||
                          env_int("FFTCONV_CUDA_CONV_SINGLE_FFT", 0);
  if (!tl_filter && !fft_forced && klen <= 256 && !env_int("FFTCONV_CUDA_NO_SMALL_DIRECT", 0) &&
      (double)batch * (double)out_len * (double)klen <= 4194304.0)
    return direct_small_device<T>(c, st, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode);
  if (klen <= direct_max_taps<T>() && !tl_filter && !env_int("FFTCONV_CUDA_NO_DIRECT", 0) &&
      !env_int("FFTCONV_CUDA_STRICT_TABLE", 0) && !env_int("FFTCONV_CUDA_FORCE_GENERIC", 0) &&
      !env_int("FFTCONV_CUDA_OA_FFT_SIZE", 0) && !env_int("FFTCONV_CUDA_CONV_SINGLE_FFT", 0))
    return direct_device<T>(c, st, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode);
  size_t N;
  if (op == OP_OA) {
    N = optimal_fft_size(klen);
    if (tl_filter && tl_filter->taps == (const void *)k) N = tl_filter->N; // the handle's spectrum
    else N = throughput_oa_size<T>(c, klen, batch, n, N);
  } else {
    // single FFT of length >= n + K - 1 (any such length gives the same linear
    // convolution as the reference's exact-length transform, fftconv.hpp:458)
    N = 16;
    while (N < n + klen - 1) N *= 2;
    if (N > max_generic_fft<T>(c, klen)) {
      // Beyond one CTA's shared memory the same linear convolution is computed with segment
      // FFTs of the table size, entirely on chip (measured 2.3x faster than the split
      // transform at 1024 x 16384 float64, K = 165).  One long FFT split over three kernels
      // with a global scratch (long_fft_kernels.cuh) -- the reference's own method,
      // fftconv.hpp:338-380 -- serves taps too long for a segment FFT, or everything with
      // FFTCONV_CUDA_CONV_SINGLE_FFT=1, while the direct O(N K) spectrum evaluation stays cheap.
      const bool oa_feasible = optimal_fft_size(klen) <= max_generic_fft<T>(c, klen);
      const bool prefer_long = env_int("FFTCONV_CUDA_CONV_SINGLE_FFT", 0) || !oa_feasible;
      const bool want_long = prefer_long && !env_int("FFTCONV_CUDA_CONV_LONG_AS_OA", 0) && long_ok<T>(c, N) &&
                             (double)N * (double)klen <= 536870912.0;
      if (want_long) {
        Tables *tbig = nullptr;
        if (int r = get_tables<T>(c, N, &tbig)) return r;
        const cx<T> *H = nullptr;
        if (int r = prepare_spectrum<T>(c, st, k, klen, N, fc::generic::LAYOUT_LONG, *tbig, &H)) return r;
        return long_device<T>(c, st, a, batch, n, a_stride, out, out_len, out_stride,
                              mode == FFTCONV_CUDA_SAME ? (int)(klen / 2) : 0, N, H, false);
      }
      N = throughput_oa_size<T>(c, klen, batch, n, optimal_fft_size(klen));
    }
  }
  // beyond one CTA's shared memory -- or inside the narrow window where the segment fits only the
  // plain radix-4 kernel (measured 7.8 ms against ~3 ms at 2048 x 65536 float64, K = 3000)
  const bool slow_window = N >= 4096 && !fast_ok<T>(c, N, klen) && !(tl_filter && tl_filter->taps == (const void *)k) &&
                           !env_int("FFTCONV_CUDA_FORCE_GENERIC", 0);
  if (N > max_generic_fft<T>(c, klen) || slow_window) {
    if (tl_filter && tl_filter->taps == (const void *)k)
      return fail(FFTCONV_CUDA_ERR_UNSUPPORTED, "filter handle: kernel too long for a segment FFT");
    return partitioned_device<T>(c, st, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode);
  }
  return oa_device<T>(c, st, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode, N);
}


// ---- general FFT surface (fft_global.cuh) ---------------------------------------
// Batched complex transforms of any length through global memory, FFTW conventions.  Used by the
// fftconv_cuda_fft_* entry points and by Hilbert envelopes of lines beyond the on-chip / split kernels.
using fc::gfft::cxg;
inline unsigned gblocks(long long total) { return (unsigned)((total + 255) / 256); }

template <typename T> int gfft_twiddles(DeviceCtx &c, size_t m, const cxg<T> **out) {
  auto key = std::make_pair(m, dtype_id<T>());
  auto it = c.gfft_w.find(key);
  if (it == c.gfft_w.end()) {
    std::vector<cxg<T>> w(m);
    for (size_t k = 0; k < m; ++k) {
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)m;
      w[k].x = (T)cosl(a);
      w[k].y = (T)(-sinl(a));
    }
    void *d = nullptr;
    CU_TRY(cudaMalloc(&d, m * sizeof(cxg<T>)));
    CU_TRY(cudaMemcpy(d, w.data(), m * sizeof(cxg<T>), cudaMemcpyHostToDevice));
    it = c.gfft_w.emplace(key, d).first;
  }
  *out = static_cast<const cxg<T> *>(it->second);
  return 0;
}

// rows x m complex points, m a power of two; in, out, tmp distinct (in is not modified)
template <typename T>
int gfft_pow2(DeviceCtx &c, cudaStream_t st, const cxg<T> *in, cxg<T> *out, cxg<T> *tmp, size_t m, size_t rows, int sign) {
  if (m == 1) {
    CU_TRY(cudaMemcpyAsync(out, in, rows * sizeof(cxg<T>), cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  const cxg<T> *W = nullptr;
  if (int r = gfft_twiddles<T>(c, m, &W)) return r;
  const int l2 = ilog2(m);
  const int passes = l2 / 2 + (l2 & 1);
  const cxg<T> *src = in;
  size_t ns = 1;
  for (int p = 1; p <= passes; ++p) {
    cxg<T> *dst = ((passes - p) % 2 == 0) ? out : tmp;
    const bool radix2 = (l2 & 1) && p == passes; // the odd factor of two goes last
    if (radix2) {
      fc::gfft::stockham_pass<T, 2><<<gblocks((long long)(rows * m / 2)), 256, 0, st>>>(src, dst, W, (long long)m, (long long)ns, sign, (long long)rows);
      ns *= 2;
    } else {
      fc::gfft::stockham_pass<T, 4><<<gblocks((long long)(rows * m / 4)), 256, 0, st>>>(src, dst, W, (long long)m, (long long)ns, sign, (long long)rows);
      ns *= 4;
    }
    CU_TRY(cudaGetLastError());
    ++g_launches;
    src = dst;
  }
  return 0;
}

template <typename T> int gfft_chirp_plan(DeviceCtx &c, cudaStream_t st, size_t n, DeviceCtx::ChirpPlan **out) {
  auto key = std::make_pair(n, dtype_id<T>());
  auto it = c.gfft_chirp.find(key);
  if (it == c.gfft_chirp.end()) {
    size_t M = 1;
    while (M < 2 * n - 1) M *= 2;
    const long double pi = 3.14159265358979323846264338327950288L;
    std::vector<cxg<T>> chirp(n), b(M);
    for (size_t i = 0; i < M; ++i) b[i] = {(T)0, (T)0};
    for (size_t j = 0; j < n; ++j) {
      const unsigned long long j2 = (unsigned long long)((unsigned __int128)j * j % (2ULL * n)); // angle reduced exactly
      const long double a = pi * (long double)j2 / (long double)n;
      chirp[j] = {(T)cosl(a), (T)(-sinl(a))};
      const cxg<T> cj = {chirp[j].x, -chirp[j].y};
      b[j] = cj;
      if (j) b[M - j] = cj;
    }
    DeviceCtx::ChirpPlan pl;
    pl.M = M;
    void *braw = nullptr, *tmp = nullptr;
    CU_TRY(cudaMalloc(&pl.chirp, n * sizeof(cxg<T>)));
    CU_TRY(cudaMalloc(&pl.Bf, M * sizeof(cxg<T>)));
    CU_TRY(cudaMalloc(&braw, M * sizeof(cxg<T>)));
    CU_TRY(cudaMalloc(&tmp, M * sizeof(cxg<T>)));
    CU_TRY(cudaMemcpy(pl.chirp, chirp.data(), n * sizeof(cxg<T>), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(braw, b.data(), M * sizeof(cxg<T>), cudaMemcpyHostToDevice));
    const int r = gfft_pow2<T>(c, st, static_cast<const cxg<T> *>(braw), static_cast<cxg<T> *>(pl.Bf), static_cast<cxg<T> *>(tmp), M, 1, -1);
    CU_TRY(cudaStreamSynchronize(st));
    cudaFree(braw);
    cudaFree(tmp);
    if (r) return r;
    it = c.gfft_chirp.emplace(key, pl).first;
  }
  *out = &it->second;
  return 0;
}

// work-buffer elements (complex) a transform of `rows` rows of n points needs beside in and out
template <typename T> size_t gfft_work_elems(size_t n, size_t rows) {
  if ((n & (n - 1)) == 0) return rows * n;
  size_t M = 1;
  while (M < 2 * n - 1) M *= 2;
  return 3 * rows * M;
}
// rows x n complex points, any n; in and out distinct; `work`: gfft_work_elems elements
template <typename T>
int gfft_any(DeviceCtx &c, cudaStream_t st, const cxg<T> *in, cxg<T> *out, cxg<T> *work, size_t n, size_t rows, int sign) {
  if ((n & (n - 1)) == 0) return gfft_pow2<T>(c, st, in, out, work, n, rows, sign);
  DeviceCtx::ChirpPlan *pl = nullptr;
  if (int r = gfft_chirp_plan<T>(c, st, n, &pl)) return r;
  const size_t M = pl->M;
  cxg<T> *a = work, *A = work + rows * M, *t = work + 2 * rows * M;
  const cxg<T> *chirp = static_cast<const cxg<T> *>(pl->chirp);
  const int cj = sign > 0 ? 1 : 0; // backward = conj o forward o conj
  fc::gfft::chirp_in<T><<<gblocks((long long)(rows * M)), 256, 0, st>>>(in, (long long)n, (long long)rows, chirp, a, (long long)M, cj);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  if (int r = gfft_pow2<T>(c, st, a, A, t, M, rows, -1)) return r;
  fc::gfft::pointwise_mul<T><<<gblocks((long long)(rows * M)), 256, 0, st>>>(A, static_cast<const cxg<T> *>(pl->Bf), (long long)M, (long long)rows);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  if (int r = gfft_pow2<T>(c, st, A, a, t, M, rows, +1)) return r;
  fc::gfft::chirp_out<T><<<gblocks((long long)(rows * n)), 256, 0, st>>>(a, (long long)M, (long long)n, (long long)rows, chirp, out, (T)(1.0 / (double)M), cj);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

// Hilbert envelope of lines of any length (device pointers): pairs of rows as real / imaginary part
template <typename T>
int hilbert_general_device(DeviceCtx &c, cudaStream_t st, const T *x, size_t batch, size_t n, size_t x_stride, T *env,
                           size_t env_stride) {
  const size_t pairs_total = (batch + 1) / 2;
  const size_t per_pair = 2 * n + gfft_work_elems<T>(n, 1); // Z, H, work
  const size_t budget = ((size_t)env_int("FFTCONV_CUDA_SCRATCH_MB", 512) << 20) / sizeof(cxg<T>);
  const size_t pairs_per = std::max<size_t>(1, std::min(pairs_total, budget / per_pair));
  StreamWs &ws = c.ws[st];
  const size_t need = pairs_per * per_pair * sizeof(cxg<T>);
  if (ws.scratch_bytes < need) CU_TRY(cudaStreamSynchronize(st));
  if (int r = grow(&ws.scratch, &ws.scratch_bytes, need)) return r;
  cxg<T> *Z = static_cast<cxg<T> *>(ws.scratch), *Hh = Z + pairs_per * n, *work = Hh + pairs_per * n;
  const fc::EnvEpilogue<T> epi = current_epilogue<T>(n);
  for (size_t p0 = 0; p0 < pairs_total; p0 += pairs_per) {
    const size_t np = std::min(pairs_per, pairs_total - p0);
    const size_t r0 = 2 * p0, rows = std::min(batch - r0, 2 * np);
    fc::gfft::real_to_complex<T><<<gblocks((long long)(np * n)), 256, 0, st>>>(x + r0 * x_stride, (long long)x_stride, (long long)n, (long long)rows, Hh, (long long)n, 1);
    CU_TRY(cudaGetLastError());
    ++g_launches;
    if (int r = gfft_any<T>(c, st, Hh, Z, work, n, np, -1)) return r;
    fc::gfft::hilbert_multiplier<T><<<gblocks((long long)(np * n)), 256, 0, st>>>(Z, (long long)n, (long long)np);
    CU_TRY(cudaGetLastError());
    ++g_launches;
    if (int r = gfft_any<T>(c, st, Z, Hh, work, n, np, +1)) return r;
    fc::gfft::envelope_from_pairs<T><<<gblocks((long long)(rows * n)), 256, 0, st>>>(
        x + r0 * x_stride, (long long)x_stride, (long long)n, (long long)rows, Hh, env + r0 * env_stride, (long long)env_stride,
        (long long)(epi.plain ? 1 : epi.dec), epi.plain ? 0 : epi.log, epi.a, epi.b);
    CU_TRY(cudaGetLastError());
    ++g_launches;
  }
  return 0;
}

// fftconv_cuda_fft_*: batched 1-D transforms with FFTW's conventions (unnormalised; sign -1 forward, +1 backward;
// r2c writes n / 2 + 1 points per row, c2r reads as many and ignores the imaginary parts of DC and Nyquist).
// kind 0: c2c, 1: r2c, 2: c2r.  Device pointers; rows contiguous.  c.mu held by the caller.
template <typename T>
int fft_device(DeviceCtx &c, cudaStream_t st, int kind, const void *in, void *out, size_t n, size_t batch, int sign) {
  const size_t half = n / 2 + 1;
  // complex elements per row beside the caller's arrays: work + (r2c / c2r, or in-place c2c: two staging rows)
  const bool staged = kind != 0 || in == out;
  const size_t per_row = gfft_work_elems<T>(n, 1) + (kind != 0 ? 2 * n : (staged ? n : 0));
  const size_t budget = ((size_t)env_int("FFTCONV_CUDA_SCRATCH_MB", 512) << 20) / sizeof(cxg<T>);
  const size_t rows_per = std::max<size_t>(1, std::min(batch, budget / std::max<size_t>(1, per_row)));
  StreamWs &ws = c.ws[st];
  const size_t need = std::max<size_t>(16, rows_per * per_row * sizeof(cxg<T>));
  if (ws.scratch_bytes < need) CU_TRY(cudaStreamSynchronize(st));
  if (int r = grow(&ws.scratch, &ws.scratch_bytes, need)) return r;
  cxg<T> *work = static_cast<cxg<T> *>(ws.scratch);
  cxg<T> *za = work + rows_per * gfft_work_elems<T>(n, 1), *zb = za + rows_per * n;
  for (size_t r0 = 0; r0 < batch; r0 += rows_per) {
    const size_t rows = std::min(rows_per, batch - r0);
    if (kind == 0) {
      const cxg<T> *src = static_cast<const cxg<T> *>(in) + r0 * n;
      cxg<T> *dst = static_cast<cxg<T> *>(out) + r0 * n;
      if (staged) {
        CU_TRY(cudaMemcpyAsync(za, src, rows * n * sizeof(cxg<T>), cudaMemcpyDeviceToDevice, st));
        src = za;
      }
      if (int r = gfft_any<T>(c, st, src, dst, work, n, rows, sign)) return r;
    } else if (kind == 1) {
      fc::gfft::real_to_complex<T><<<gblocks((long long)(rows * n)), 256, 0, st>>>(
          static_cast<const T *>(in) + r0 * n, (long long)n, (long long)n, (long long)rows, za, (long long)n, 0);
      CU_TRY(cudaGetLastError());
      ++g_launches;
      if (int r = gfft_any<T>(c, st, za, zb, work, n, rows, -1)) return r;
      fc::gfft::complex_take<T><<<gblocks((long long)(rows * half)), 256, 0, st>>>(
          zb, (long long)n, (long long)rows, static_cast<cxg<T> *>(out) + r0 * half, (long long)half);
      CU_TRY(cudaGetLastError());
      ++g_launches;
    } else {
      fc::gfft::hermitian_expand<T><<<gblocks((long long)(rows * n)), 256, 0, st>>>(
          static_cast<const cxg<T> *>(in) + r0 * half, (long long)n, (long long)rows, za);
      CU_TRY(cudaGetLastError());
      ++g_launches;
      if (int r = gfft_any<T>(c, st, za, zb, work, n, rows, +1)) return r;
      fc::gfft::complex_real_part<T><<<gblocks((long long)(rows * n)), 256, 0, st>>>(zb, (long long)n, (long long)rows,
                                                                                    static_cast<T *>(out) + r0 * n);
      CU_TRY(cudaGetLastError());
      ++g_launches;
    }
  }
  return 0;
}

template <typename T>
int hilbert_device(DeviceCtx &c, cudaStream_t st, const T *x, size_t batch, size_t n,
                   size_t x_stride, T *env, size_t env_stride) {
  if ((n & (n - 1)) != 0) {
    size_t M = 1;
    while (M < 2 * n - 1) M *= 2;
    const size_t smem_b = M * sizeof(cx<T>);
    if (smem_b > c.smem_optin) {
      // the chirp filter's two convolutions as split FFTs of M = 16 N_sub points through global scratch
      if (!long_ok<T>(c, M)) // beyond the split (16 x 16384) FFT: the general transform through global memory
        return hilbert_general_device<T>(c, st, x, batch, n, x_stride, env, env_stride);
      BluesteinTables *btl = nullptr;
      if (int r = get_bluestein<T>(c, n, &btl, true)) return r;
      return long_device<T>(c, st, x, batch, n, x_stride, env, env_out_len(n), env_stride, 0, M,
                            static_cast<const cx<T> *>(btl->B), true, static_cast<const cx<T> *>(btl->chirp));
    }
    BluesteinTables *bt = nullptr;
    if (int r = get_bluestein<T>(c, n, &bt)) return r;
    Tables *tabm = nullptr;
    if (int r = get_tables<T>(c, M, &tabm)) return r;
    fc::generic::HilbertBluesteinArgs<T> p{};
    p.in = x;
    p.out = env;
    p.batch = (long long)batch;
    p.n = (long long)n;
    p.in_stride = (long long)x_stride;
    p.out_stride = (long long)env_stride;
    p.log2m = ilog2(M);
    p.W = static_cast<const cx<T> *>(tabm->W);
    p.B = static_cast<const cx<T> *>(bt->B);
    p.chirp = static_cast<const cx<T> *>(bt->chirp);
    p.epi = current_epilogue<T>(n);
    if (int r = ensure_smem(c, fc::generic::hilbert_bluestein_kernel<T>, std::max<size_t>(smem_b, 16))) return r;
    const long long gridb = ((long long)batch + 1) / 2;
    const int threadsb = (int)std::min<size_t>(512, std::max<size_t>(32, M / 4));
    fc::generic::hilbert_bluestein_kernel<T><<<(unsigned)gridb, threadsb, std::max<size_t>(smem_b, 16), st>>>(p);
    CU_TRY(cudaGetLastError());
    ++g_launches;
    return 0;
  }
  if constexpr (std::is_same<T, float>::value) {
    // 8192-sample float32 lines (BASELINE config 4): the real-input engine -- one 4096-point complex FFT per
    // line, 128 threads per line, four independent lines per SM (hilbert_r2c_core.cuh).  Needs 16-byte aligned
    // input rows (and output rows, for the plain epilogue) and enough lines to give every group a few;
    // FFTCONV_CUDA_NO_HR2C=1 switches it off.
    const fc::EnvEpilogue<T> epi = current_epilogue<T>(n);
    const bool in16 = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (x_stride & 3) == 0;
    const bool out16 = (reinterpret_cast<uintptr_t>(env) & 15) == 0 && (env_stride & 3) == 0;
    if (n == fc::hr2c::kLineSamples && in16 && (out16 || !epi.plain) && batch >= 64 &&
        fc::hr2c::table_bytes() + 4 * 2 * 4352 * sizeof(float) <= c.smem_optin && !env_int("FFTCONV_CUDA_NO_HR2C", 0) &&
        !env_int("FFTCONV_CUDA_FORCE_GENERIC", 0)) {
      if (!c.hr2c_tables) {
        std::vector<unsigned char> host(fc::hr2c::table_bytes());
        fc::hr2c::build_tables(host.data());
        CU_TRY(cudaMalloc(&c.hr2c_tables, host.size()));
        CU_TRY(cudaMemcpy(c.hr2c_tables, host.data(), host.size(), cudaMemcpyHostToDevice));
      }
      const long long ctas = std::max<long long>(1, std::min<long long>(c.sm_count, ((long long)batch + 3) / 4));
      const fc::hr2c::Epilogue he{epi.plain, epi.log, epi.dec, epi.magic, (float)epi.a, (float)epi.b};
      if (const int e = fc::hr2c::launch(x, env, (long long)batch, (long long)x_stride, (long long)env_stride,
                                         c.hr2c_tables, he, (int)ctas, st))
        return fail(FFTCONV_CUDA_ERR_CUDA, std::string("hilbert_r2c_kernel: ") + cudaGetErrorString((cudaError_t)e));
      ++g_launches;
      return 0;
    }
  }
  if (!fast_ok<T>(c, n, 0) && long_ok<T>(c, n)) {
    Tables *tbig = nullptr;
    if (int r = get_tables<T>(c, n, &tbig)) return r;
    if (!tbig->Glong) {
      const int l2s = ilog2(n) - 4;
      const size_t nsub = n / 16;
      std::vector<cx<T>> g(n);
      for (size_t idx = 0; idx < n; ++idx) {
        const size_t kf = idx / nsub + 16 * (size_t)fc::fast::freq_of_pos(
                                               fc::fast::pos_of_table_index((unsigned)(idx % nsub), l2s), l2s);
        T im = 0;
        if (kf != 0 && 2 * kf != n) im = (T)((2 * kf < n ? -1.0 : 1.0) / (double)n);
        g[idx].x = 0;
        g[idx].y = im;
      }
      CU_TRY(cudaMalloc(&tbig->Glong, n * sizeof(cx<T>)));
      CU_TRY(cudaMemcpy(tbig->Glong, g.data(), n * sizeof(cx<T>), cudaMemcpyHostToDevice));
    }
    return long_device<T>(c, st, x, batch, n, x_stride, env, env_out_len(n), env_stride, 0, n,
                          static_cast<const cx<T> *>(tbig->Glong), true);
  }
  const size_t smem = n * sizeof(cx<T>);
  if (smem > c.smem_optin) // beyond the single-CTA and the split (16 x 16384) FFT
    return hilbert_general_device<T>(c, st, x, batch, n, x_stride, env, env_stride);
  Tables *tab = nullptr;
  if (int r = get_tables<T>(c, n, &tab)) return r;
  const bool use_fast = fast_ok<T>(c, n, 0);
  if (int r = get_hilbert_table<T>(c, n, *tab, use_fast)) return r;
  fc::generic::HilbertArgs<T> p{};
  p.in = x;
  p.out = env;
  p.batch = (long long)batch;
  p.n = (long long)n;
  p.in_stride = (long long)x_stride;
  p.out_stride = (long long)env_stride;
  p.log2n = ilog2(n);
  p.W = static_cast<const cx<T> *>(tab->W);
  p.G = static_cast<const cx<T> *>(use_fast ? tab->Gfast : tab->G);
  p.epi = current_epilogue<T>(n);
  if (use_fast) {
    if (!tab->TWfast) {
      auto build = [&]() -> int { FAST_SWITCH(build_fast_twiddles, p.log2n, *tab) };
      if (int r = build()) return r;
    }
    p.W = static_cast<const cx<T> *>(tab->TWfast);
    FAST_SWITCH(launch_hilbert_fast, p.log2n, c, st, p)
  }
  if (int r = ensure_smem(c, fc::generic::hilbert_pow2_kernel<T>, std::max<size_t>(smem, 16))) return r;
  const long long grid = ((long long)batch + 1) / 2;
  const int threads = (int)std::min<size_t>(512, std::max<size_t>(32, n / 4));
  fc::generic::hilbert_pow2_kernel<T><<<(unsigned)grid, threads, std::max<size_t>(smem, 16), st>>>(p);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

// ---- host-pointer path ------------------------------------------------------------------
// Kind of memory behind a caller's pointer.  The reference's callers hand in std::vector / NumPy
// arrays, i.e. PAGEABLE memory; cudaMemcpyAsync on pageable memory is staged by the driver and
// blocks the calling thread, so nothing would overlap.
enum PtrKind { PTR_DEVICE, PTR_PINNED, PTR_PAGEABLE };
PtrKind ptr_kind(const void *p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return PTR_PAGEABLE;
  }
  if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) return PTR_DEVICE;
  return attr.type == cudaMemoryTypeHost ? PTR_PINNED : PTR_PAGEABLE;
}

// one row block of the host pipeline on device buffers: kernels only, c.mu held by the caller
template <typename T>
int block_device(DeviceCtx &c, cudaStream_t st, Op op, void **tmp_buf, size_t *tmp_cap, size_t rows_cap, const T *din,
                 size_t rows, size_t n, const T *k, size_t klen, T *dout, size_t out_len, int mode) {
  if (op == OP_HILBERT) return hilbert_device<T>(c, st, din, rows, n, n, dout, out_len);
  if (op == OP_RF) {
    if (int g = grow(tmp_buf, tmp_cap, rows_cap * n * sizeof(T))) return g;
    T *tmp = static_cast<T *>(*tmp_buf);
    if (int r = conv_device<T>(c, st, OP_OA, din, rows, n, n, k, klen, tmp, n, n, FFTCONV_CUDA_SAME)) return r;
    return hilbert_device<T>(c, st, tmp, rows, n, n, dout, out_len);
  }
  return conv_device<T>(c, st, op, din, rows, n, n, k, klen, dout, out_len, out_len, mode);
}

// Pinned (or registered) host memory: row blocks pipelined over three streams; every copy is one
// cudaMemcpyAsync (or one 2-D copy for strided rows) that returns at once.
template <typename T>
int run_host_pinned(DeviceCtx &c, Op op, const T *a, size_t batch, size_t n, size_t a_stride, const T *k,
                    size_t klen, T *out, size_t out_len, size_t out_stride, int mode) {
  const size_t row_bytes = std::max(n, out_len) * sizeof(T);
  const size_t target = (size_t)env_int("FFTCONV_CUDA_HOST_BLOCK_MB", 32) << 20;
  size_t rows_per_block = std::max<size_t>(1, target / std::max<size_t>(1, row_bytes));
  rows_per_block = std::min(rows_per_block, batch);
  // a medium batch (a 512-row shard is four 32 MB blocks) would barely pipeline: aim for >= 12 blocks, but
  // not below 4 MB, where the per-copy overhead shows
  while (rows_per_block > 1 && batch / rows_per_block < 12 && rows_per_block * row_bytes > (size_t)(4 << 20))
    rows_per_block = (rows_per_block + 1) / 2;
  // Block schedule: full blocks in the middle, blocks of 1/8, 1/4, 1/2 of that at both ends -- the
  // first upload and the last download are the only copies nothing else overlaps with.
  std::vector<size_t> blocks;
  {
    std::vector<size_t> ramp;
    if (rows_per_block >= 8 && batch >= 6 * rows_per_block && !env_int("FFTCONV_CUDA_HOST_NO_RAMP", 0))
      for (size_t r = rows_per_block / 8; r < rows_per_block; r *= 2) ramp.push_back(r);
    size_t ramp_rows = 0;
    for (size_t r : ramp) ramp_rows += r;
    blocks = ramp;
    for (size_t left = batch - 2 * ramp_rows; left > 0;) {
      const size_t rows = std::min(rows_per_block, left);
      blocks.push_back(rows);
      left -= rows;
    }
    blocks.insert(blocks.end(), ramp.rbegin(), ramp.rend());
  }
  size_t r0 = 0;
  for (size_t b = 0; b < blocks.size(); r0 += blocks[b], ++b) {
    const int s = (int)(b % kPipe);
    const size_t rows = blocks[b];
    if (int r = grow(&c.stage_in[s], &c.cap_in[s], rows_per_block * n * sizeof(T))) return r;
    if (int r = grow(&c.stage_out[s], &c.cap_out[s], rows_per_block * out_len * sizeof(T))) return r;
    T *din = static_cast<T *>(c.stage_in[s]);
    T *dout = static_cast<T *>(c.stage_out[s]);
    if (a_stride == n) // contiguous rows: one linear copy
      CU_TRY(cudaMemcpyAsync(din, a + r0 * n, rows * n * sizeof(T), cudaMemcpyHostToDevice, c.pipe[s]));
    else
      CU_TRY(cudaMemcpy2DAsync(din, n * sizeof(T), a + r0 * a_stride, a_stride * sizeof(T),
                               n * sizeof(T), rows, cudaMemcpyHostToDevice, c.pipe[s]));
    {
      std::lock_guard<std::mutex> lock(c.mu); // plan cache + enqueue only
      if (int r = block_device<T>(c, c.pipe[s], op, &c.stage_tmp[s], &c.cap_tmp[s], rows_per_block, din, rows, n, k, klen,
                                  dout, out_len, mode))
        return r;
    }
    if (out_stride == out_len)
      CU_TRY(cudaMemcpyAsync(out + r0 * out_len, dout, rows * out_len * sizeof(T), cudaMemcpyDeviceToHost, c.pipe[s]));
    else
      CU_TRY(cudaMemcpy2DAsync(out + r0 * out_stride, out_stride * sizeof(T), dout,
                               out_len * sizeof(T), out_len * sizeof(T), rows,
                               cudaMemcpyDeviceToHost, c.pipe[s]));
  }
  for (int i = 0; i < kPipe; ++i) CU_TRY(cudaStreamSynchronize(c.pipe[i]));
  return 0;
}

// Pageable host memory through a page-locked ring: up to kRing worker threads, each with its own
// stream, device staging and a pinned in / out buffer.  A worker copies a row block into its pinned
// buffer (plain memcpy -- the part the driver would otherwise do single-threaded and synchronously),
// uploads, computes, downloads and copies out; the workers overlap one another's phases.  Measured on
// the headline shape (profiles/r02_hostab.jsonl): driver staging 155 ms, cudaHostRegister per call
// 162 ms, 3 workers x 16 MB 100 ms, and one numpy-style single-threaded copy of 1 GB alone takes 71 ms
// -- the host memcpy is the bound, hence many workers.
template <typename T> int ring_alloc(void **p, size_t *cap, size_t bytes) {
  if (*cap >= bytes) return 0;
  cudaFreeHost(*p);
  *p = nullptr;
  *cap = 0;
  CU_TRY(cudaHostAlloc(p, bytes, cudaHostAllocPortable));
  *cap = bytes;
  return 0;
}
template <typename T>
int run_host_ring(DeviceCtx &c, Op op, const T *a, size_t batch, size_t n, size_t a_stride, const T *k,
                  size_t klen, T *out, size_t out_len, size_t out_stride, int mode) {
  const size_t row_bytes = std::max(n, out_len) * sizeof(T);
  const size_t target = (size_t)env_int("FFTCONV_CUDA_HOST_RING_BLOCK_MB", 32) << 20;
  int max_workers = env_int("FFTCONV_CUDA_HOST_RING_THREADS", kRing);
  max_workers = std::max(1, std::min(max_workers, kRing));
  const unsigned hw = std::thread::hardware_concurrency();
  if (hw) max_workers = std::min<int>(max_workers, (int)std::max(1u, hw / 2));
  size_t rows_per_block = std::min(batch, std::max<size_t>(1, target / std::max<size_t>(1, row_bytes)));
  // at least two blocks per worker, so that one worker's copies overlap its neighbours' transfers
  while (rows_per_block > 1 && (batch + rows_per_block - 1) / rows_per_block < (size_t)(2 * max_workers) &&
         rows_per_block * row_bytes > (size_t)(2 << 20))
    rows_per_block = (rows_per_block + 1) / 2;
  const size_t n_blocks = (batch + rows_per_block - 1) / rows_per_block;
  const int workers = (int)std::min<size_t>((size_t)max_workers, n_blocks);
  for (int s = 0; s < workers; ++s) {
    RingSlot &r = c.ring[s];
    if (!r.st) CU_TRY(cudaStreamCreateWithFlags(&r.st, cudaStreamNonBlocking));
    if (int e = grow(&r.din, &r.cap_din, rows_per_block * n * sizeof(T))) return e;
    if (int e = grow(&r.dout, &r.cap_dout, rows_per_block * out_len * sizeof(T))) return e;
    if (int e = ring_alloc<T>(&r.hin, &r.cap_hin, rows_per_block * n * sizeof(T))) return e;
    if (int e = ring_alloc<T>(&r.hout, &r.cap_hout, rows_per_block * out_len * sizeof(T))) return e;
  }
  int status[kRing] = {0};
  std::string msgs[kRing];
  const Filter *parent_filter = tl_filter; // thread_local call context: hand it to the workers
  const EnvOpts parent_env = tl_env;
  auto worker = [&](int s) {
    auto body = [&]() -> int {
      tl_filter = parent_filter;
      tl_env = parent_env;
      CU_TRY(cudaSetDevice(c.device));
      RingSlot &r = c.ring[s];
      T *din = static_cast<T *>(r.din), *dout = static_cast<T *>(r.dout);
      T *hin = static_cast<T *>(r.hin), *hout = static_cast<T *>(r.hout);
      for (size_t b = (size_t)s; b < n_blocks; b += (size_t)workers) {
        const size_t r0 = b * rows_per_block, rows = std::min(rows_per_block, batch - r0);
        if (a_stride == n) std::memcpy(hin, a + r0 * n, rows * n * sizeof(T));
        else
          for (size_t i = 0; i < rows; ++i) std::memcpy(hin + i * n, a + (r0 + i) * a_stride, n * sizeof(T));
        CU_TRY(cudaMemcpyAsync(din, hin, rows * n * sizeof(T), cudaMemcpyHostToDevice, r.st));
        {
          std::lock_guard<std::mutex> lock(c.mu);
          if (int e = block_device<T>(c, r.st, op, &r.dtmp, &r.cap_dtmp, rows_per_block, din, rows, n, k, klen, dout,
                                      out_len, mode))
            return e;
        }
        CU_TRY(cudaMemcpyAsync(hout, dout, rows * out_len * sizeof(T), cudaMemcpyDeviceToHost, r.st));
        CU_TRY(cudaStreamSynchronize(r.st));
        if (out_stride == out_len) std::memcpy(out + r0 * out_len, hout, rows * out_len * sizeof(T));
        else
          for (size_t i = 0; i < rows; ++i) std::memcpy(out + (r0 + i) * out_stride, hout + i * out_len, out_len * sizeof(T));
      }
      return 0;
    };
    status[s] = body();
    if (status[s]) msgs[s] = g_err; // g_err is thread_local: carry the message back
    tl_filter = nullptr;
    tl_env = EnvOpts();
  };
  std::thread th[kRing];
  for (int s = 1; s < workers; ++s) th[s] = std::thread(worker, s);
  worker(0);
  tl_filter = parent_filter; // worker(0) ran on the calling thread
  tl_env = parent_env;
  for (int s = 1; s < workers; ++s) th[s].join();
  for (int s = 0; s < workers; ++s)
    if (status[s]) return fail(status[s], msgs[s]);
  return 0;
}

// A tiny convolution through the host API (BASELINE config 1: n = 4352, K = 65; the reference needs
// 17.2 us per call, README.md:99-102): no cudaMemcpy at all -- the signal is copied into a mapped
// page-locked buffer, one kernel reads it across PCIe and writes the result back the same way.
constexpr size_t kTinyBytes = 256 << 10;
template <typename T> bool tiny_ok(Op op, size_t batch, size_t n, size_t klen, size_t out_len) {
  if (op != OP_OA && op != OP_CONV) return false;
  if (tl_filter || klen > 256 || batch > 64) return false;
  if ((double)batch * (double)out_len * (double)klen > 4194304.0) return false;
  if (batch * (n + out_len) * sizeof(T) + klen * sizeof(T) > kTinyBytes) return false;
  return !(env_int("FFTCONV_CUDA_NO_DIRECT", 0) || env_int("FFTCONV_CUDA_STRICT_TABLE", 0) ||
           env_int("FFTCONV_CUDA_FORCE_GENERIC", 0) || env_int("FFTCONV_CUDA_OA_FFT_SIZE", 0) ||
           env_int("FFTCONV_CUDA_CONV_SINGLE_FFT", 0) || env_int("FFTCONV_CUDA_NO_SMALL_DIRECT", 0) ||
           env_int("FFTCONV_CUDA_NO_TINY_HOST", 0));
}
template <typename T>
int run_host_tiny(DeviceCtx &c, const T *a, size_t batch, size_t n, size_t a_stride, const T *k, size_t klen, T *out,
                  size_t out_len, size_t out_stride, int mode) {
  if (!c.tiny_host) {
    CU_TRY(cudaHostAlloc(&c.tiny_host, kTinyBytes + 64, cudaHostAllocMapped | cudaHostAllocPortable));
    CU_TRY(cudaHostGetDevicePointer(&c.tiny_dev, c.tiny_host, 0));
    CU_TRY(cudaStreamCreateWithFlags(&c.tiny_stream, cudaStreamNonBlocking));
    CU_TRY(cudaMalloc((void **)&c.tiny_counter, sizeof(unsigned)));
    CU_TRY(cudaMemset(c.tiny_counter, 0, sizeof(unsigned)));
    std::memset(static_cast<char *>(c.tiny_host) + kTinyBytes, 0, 64);
    c.tiny_bytes = kTinyBytes;
  }
  T *h = static_cast<T *>(c.tiny_host), *d = static_cast<T *>(c.tiny_dev);
  // completion flag: the last 64 bytes of the mapped buffer.  The kernel's last CTA writes the
  // call's epoch there after a system-wide fence; polling it costs less than a stream synchronise.
  volatile unsigned *flag_h = reinterpret_cast<volatile unsigned *>(static_cast<char *>(c.tiny_host) + kTinyBytes);
  unsigned *flag_d = reinterpret_cast<unsigned *>(static_cast<char *>(c.tiny_dev) + kTinyBytes);
  const unsigned epoch = ++c.tiny_epoch ? c.tiny_epoch : ++c.tiny_epoch; // never 0
  for (size_t r = 0; r < batch; ++r) std::memcpy(h + r * n, a + r * a_stride, n * sizeof(T));
  const size_t out_off = batch * n, taps_off = out_off + batch * out_len;
  fc::direct::SmallArgs<T> p{};
  p.in = d;
  p.out = d + out_off;
  p.batch = (long long)batch;
  p.n = (long long)n;
  p.in_stride = (long long)n;
  p.out_stride = (long long)out_len;
  p.out_len = (long long)out_len;
  p.klen = (int)klen;
  p.pad = mode == FFTCONV_CUDA_SAME ? (int)(klen / 2) : 0;
  if (klen <= (size_t)fc::direct::kInlineTaps) {
    p.inline_taps = 1;
    std::memcpy(p.tap, k, klen * sizeof(T));
  } else {
    std::memcpy(h + taps_off, k, klen * sizeof(T));
    p.taps = d + taps_off;
  }
  const int tiles = (int)((out_len + 255) / 256);
  fc::direct::direct_tiny_kernel<T><<<(unsigned)(batch * tiles), 256, 0, c.tiny_stream>>>(p, tiles, c.tiny_counter, flag_d, epoch);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  if (env_int("FFTCONV_CUDA_TINY_SYNC", 0)) {
    CU_TRY(cudaStreamSynchronize(c.tiny_stream));
  } else {
    // spin on the flag; every 4096 polls ask the stream, so that a failed launch cannot hang the caller
    for (unsigned spins = 1; *flag_h != epoch; ++spins) {
      if ((spins & 4095u) == 0) {
        const cudaError_t q = cudaStreamQuery(c.tiny_stream);
        if (q == cudaSuccess) break; // finished: the flag is visible after the stream drained
        if (q != cudaErrorNotReady) return fail(FFTCONV_CUDA_ERR_CUDA, std::string("tiny kernel: ") + cudaGetErrorString(q));
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  for (size_t r = 0; r < batch; ++r) std::memcpy(out + r * out_stride, h + out_off + r * out_len, out_len * sizeof(T));
  return 0;
}

template <typename T>
int run_host(DeviceCtx &c, Op op, PtrKind kind_in, PtrKind kind_out, const T *a, size_t batch, size_t n, size_t a_stride,
             const T *k, size_t klen, T *out, size_t out_len, size_t out_stride, int mode) {
  std::lock_guard<std::mutex> host_lock(c.host_mu);
  if (tiny_ok<T>(op, batch, n, klen, out_len) && ptr_kind(k) != PTR_DEVICE)
    return run_host_tiny<T>(c, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode);
  for (int i = 0; i < kPipe; ++i)
    if (!c.pipe[i]) CU_TRY(cudaStreamCreateWithFlags(&c.pipe[i], cudaStreamNonBlocking));
  const size_t in_bytes = ((batch - 1) * a_stride + n) * sizeof(T), out_bytes = ((batch - 1) * out_stride + out_len) * sizeof(T);
  const bool pageable = kind_in == PTR_PAGEABLE || kind_out == PTR_PAGEABLE;
  // FFTCONV_CUDA_HOST_PAGEABLE: 0 = plain cudaMemcpyAsync on pageable memory (driver staging),
  // 1 = page-lock the caller's arrays for the duration of the call (cudaHostRegister), 2 = ring
  const int strategy = env_int("FFTCONV_CUDA_HOST_PAGEABLE", 2);
  if (pageable && in_bytes + out_bytes >= (size_t)(8 << 20)) {
    if (strategy == 2) return run_host_ring<T>(c, op, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode);
    if (strategy == 1) {
      const bool reg_in = kind_in == PTR_PAGEABLE &&
                          cudaHostRegister(const_cast<T *>(a), in_bytes, cudaHostRegisterPortable | cudaHostRegisterReadOnly) == cudaSuccess;
      const bool reg_out = kind_out == PTR_PAGEABLE && cudaHostRegister(out, out_bytes, cudaHostRegisterPortable) == cudaSuccess;
      cudaGetLastError(); // a failed registration falls back to the driver's staging
      const int r = run_host_pinned<T>(c, op, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode);
      if (reg_in) cudaHostUnregister(const_cast<T *>(a));
      if (reg_out) cudaHostUnregister(out);
      return r;
    }
  }
  return run_host_pinned<T>(c, op, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode);
}

template <typename T>
int entry(Op op, const T *a, size_t batch, size_t n, size_t a_stride, const T *k, size_t klen,
          T *out, size_t out_len, size_t out_stride, int mode, void *stream) {
  if (!a || !out) return fail(FFTCONV_CUDA_ERR_ARG, "null data pointer");
  if (batch == 0) return 0;
  if (n == 0) return fail(FFTCONV_CUDA_ERR_ARG, "empty signal");
  if (a_stride < n || out_stride < out_len) return fail(FFTCONV_CUDA_ERR_ARG, "row stride shorter than row");
  size_t zero_tail = 0; // Full mode, oversized output: samples past n + K - 1 are zeroed
  if (op == OP_RF) {
    if (!k || klen == 0) return fail(FFTCONV_CUDA_ERR_ARG, "empty kernel");
    if (out_len != env_out_len(n))
      return fail(FFTCONV_CUDA_ERR_ARG, "rf_envelope: output length " + std::to_string(out_len) + ", expected " +
                                            std::to_string(env_out_len(n)));
  } else if (op != OP_HILBERT) {
    if (!k || klen == 0) return fail(FFTCONV_CUDA_ERR_ARG, "empty kernel");
    if (mode != FFTCONV_CUDA_FULL && mode != FFTCONV_CUDA_SAME)
      return fail(FFTCONV_CUDA_ERR_ARG, "Unsupported convolution mode: " + std::to_string(mode));
    const size_t want = mode == FFTCONV_CUDA_FULL ? n + klen - 1 : n;
    // the reference zero-fills `out` and accepts a Full output longer than n + K - 1
    // (/root/reference/include/fftconv/fftconv.hpp:340, :385, :467)
    if (mode == FFTCONV_CUDA_FULL && out_len > want) {
      zero_tail = out_len - want;
      out_len = want;
    }
    if (out_len != want)
      return fail(FFTCONV_CUDA_ERR_ARG, "output length " + std::to_string(out_len) + ", expected " +
                                            std::to_string(want));
  } else if (out_len != env_out_len(n)) {
    return fail(FFTCONV_CUDA_ERR_ARG, tl_env.dec > 1 ? "envelope: output length must be ceil(n / decimate)"
                                                     : "hilbert: output length must equal input length");
  }
  DeviceCtx *c = nullptr;
  if (int r = get_ctx(&c)) return r;
  if (tl_filter && tl_filter->device != c->device)
    return fail(FFTCONV_CUDA_ERR_ARG, "filter handle lives on device " + std::to_string(tl_filter->device) +
                                          ", the current device is " + std::to_string(c->device));
  const PtrKind kind_in = ptr_kind(a), kind_out = ptr_kind(out);
  const bool dev_in = kind_in == PTR_DEVICE, dev_out = kind_out == PTR_DEVICE;
  if (dev_in != dev_out)
    return fail(FFTCONV_CUDA_ERR_ARG, "input and output must both be host or both be device pointers");
  if (dev_in) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::lock_guard<std::mutex> lock(c->mu);
    if (zero_tail)
      CU_TRY(cudaMemset2DAsync(out + out_len, out_stride * sizeof(T), 0, zero_tail * sizeof(T), batch, st));
    if (op == OP_HILBERT) return hilbert_device<T>(*c, st, a, batch, n, a_stride, out, out_stride);
    if (op == OP_RF) {
      // the filtered rows live in a per-stream workspace (grow-only, stream-ordered reuse)
      StreamWs &ws = c->ws[st];
      if (ws.rf_bytes < batch * n * sizeof(T)) CU_TRY(cudaStreamSynchronize(st)); // before the old one is freed
      if (int g = grow(&ws.rf, &ws.rf_bytes, batch * n * sizeof(T))) return g;
      T *tmp = static_cast<T *>(ws.rf);
      if (int r = conv_device<T>(*c, st, OP_OA, a, batch, n, a_stride, k, klen, tmp, n, n, FFTCONV_CUDA_SAME)) return r;
      return hilbert_device<T>(*c, st, tmp, batch, n, n, out, out_stride);
    }
    return conv_device<T>(*c, st, op, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode);
  }
  if (zero_tail)
    for (size_t r = 0; r < batch; ++r) std::memset(out + r * out_stride + out_len, 0, zero_tail * sizeof(T));
  return run_host<T>(*c, op, kind_in, kind_out, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode);
}

// fftconv_cuda_envelope_* / fftconv_cuda_rf_envelope_*: validate the options, publish them to
// the launch sites through tl_env, run the Hilbert (or filter + Hilbert) path.
template <typename T>
int envelope_entry(Op op, const T *x, size_t batch, size_t n, size_t x_stride, const T *k, size_t klen, T *out,
                   size_t out_len, size_t out_stride, const fftconv_cuda_envelope_opts *opts, void *stream) {
  EnvOpts e;
  e.dec = 1;
  if (opts) {
    e.dec = opts->decimate ? opts->decimate : 1;
    e.log = opts->log_compress ? 1 : 0;
    if (e.log) {
      if (!(opts->ref > 0.0) || !(opts->dynamic_range_db > 0.0))
        return fail(FFTCONV_CUDA_ERR_ARG, "envelope: log compression needs ref > 0 and dynamic_range_db > 0");
      // 1 + 20 log10(env / ref) / DR  =  a log2(env) + b
      e.a = 20.0 * std::log10(2.0) / opts->dynamic_range_db;
      e.b = 1.0 - 20.0 * std::log10(opts->ref) / opts->dynamic_range_db;
    }
    if (e.dec > 0x7fffffffULL) return fail(FFTCONV_CUDA_ERR_ARG, "envelope: decimation factor too large");
  }
  tl_env = e;
  const int r = entry<T>(op, x, batch, n, x_stride, k, klen, out, out_len, out_stride, FFTCONV_CUDA_SAME, stream);
  tl_env = EnvOpts();
  return r;
}

template <typename T> int add_entry(T *dst, const T *src, size_t n, void *stream) {
  if (n == 0) return 0;
  if (!dst || !src) return fail(FFTCONV_CUDA_ERR_ARG, "null pointer");
  const int threads = 256;
  fc::generic::add_inplace_kernel<T><<<(unsigned)((n + threads - 1) / threads), threads, 0,
                                       static_cast<cudaStream_t>(stream)>>>(dst, src, (long long)n);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

template <typename T> int oa_tail_entry(const T *last, const T *k, size_t klen, T *tail, void *stream) {
  if (klen <= 1) return 0;
  if (!last || !k || !tail) return fail(FFTCONV_CUDA_ERR_ARG, "null pointer");
  if (klen > 0x7fffffffULL) return fail(FFTCONV_CUDA_ERR_ARG, "oa_tail: kernel too long");
  fc::direct::oa_tail_kernel<T><<<(unsigned)((klen - 1 + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      last, k, (int)klen, tail);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

// x[i] = synthetic sample (seed, first + i): the measurement harness's signal generator
// (synthetic.cuh; numpy twin fftconv_b200/synthetic.py)
template <typename T>
int fill_synthetic_entry(T *dst, size_t n, unsigned long long seed, unsigned long long first, double gain, void *stream) {
  if (n == 0) return 0;
  if (!dst) return fail(FFTCONV_CUDA_ERR_ARG, "null pointer");
  if (ptr_kind(dst) != PTR_DEVICE) return fail(FFTCONV_CUDA_ERR_ARG, "fill_synthetic: device pointer expected");
  DeviceCtx *c = nullptr;
  if (int r = get_ctx(&c)) return r;
  const int threads = 256;
  const unsigned grid = (unsigned)std::min<size_t>((n + threads - 1) / threads, (size_t)c->sm_count * 16);
  fc::synthetic::fill_kernel<T><<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      dst, (unsigned long long)n, seed, first, (T)(fc::synthetic::kScale * gain));
  CU_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

// the general FFT surface: the reference's FFTW wrappers dft_1d / dft_r2c_1d /
// dft_c2r_1d and their batched forms (/root/reference/include/fftconv/fftw.hpp:152-195, :396-423)
template <typename T> int fft_entry(int kind, const void *in, void *out, size_t n, size_t batch, int sign, void *stream) {
  if (batch == 0) return 0;
  if (!in || !out) return fail(FFTCONV_CUDA_ERR_ARG, "fft: null pointer");
  if (n == 0) return fail(FFTCONV_CUDA_ERR_ARG, "fft: empty transform");
  if (kind == 0 && sign != -1 && sign != 1) return fail(FFTCONV_CUDA_ERR_ARG, "fft: sign must be -1 (forward) or +1 (backward)");
  if (n > ((size_t)1 << 30)) return fail(FFTCONV_CUDA_ERR_UNSUPPORTED, "fft: transforms beyond 2^30 points are not covered");
  DeviceCtx *c = nullptr;
  if (int r = get_ctx(&c)) return r;
  const size_t half = n / 2 + 1;
  const size_t in_bytes = batch * (kind == 1 ? n * sizeof(T) : (kind == 2 ? half : n) * sizeof(cxg<T>));
  const size_t out_bytes = batch * (kind == 2 ? n * sizeof(T) : (kind == 1 ? half : n) * sizeof(cxg<T>));
  const bool dev_in = ptr_kind(in) == PTR_DEVICE, dev_out = ptr_kind(out) == PTR_DEVICE;
  if (dev_in != dev_out) return fail(FFTCONV_CUDA_ERR_ARG, "fft: input and output must both be host or both be device pointers");
  if (dev_in) {
    std::lock_guard<std::mutex> lock(c->mu);
    return fft_device<T>(*c, static_cast<cudaStream_t>(stream), kind, in, out, n, batch, sign);
  }
  // host pointers: upload, transform, download (the slow lane: no pipelining)
  void *din = nullptr, *dout = nullptr;
  auto body = [&]() -> int {
    CU_TRY(cudaMalloc(&din, in_bytes));
    CU_TRY(cudaMalloc(&dout, out_bytes));
    CU_TRY(cudaMemcpy(din, in, in_bytes, cudaMemcpyHostToDevice));
    {
      std::lock_guard<std::mutex> lock(c->mu);
      if (int r = fft_device<T>(*c, nullptr, kind, din, dout, n, batch, sign)) return r;
    }
    CU_TRY(cudaMemcpy(out, dout, out_bytes, cudaMemcpyDeviceToHost));
    return 0;
  };
  const int r = body();
  cudaFree(din);
  cudaFree(dout);
  return r;
}

template <typename T> int oa_layout_for(const DeviceCtx &c, size_t N, size_t klen) {
  if (w1k_engine_for<T>(N, klen)) return fc::generic::LAYOUT_W1K;
  if (oa2048_engine_for<T>(N, klen)) return fc::generic::LAYOUT_N2048;
  return fast_ok<T>(c, N, klen) ? fc::generic::LAYOUT_FAST : fc::generic::LAYOUT_DIF;
}

template <typename T> int filter_create(const T *k, size_t klen, void **out) {
  if (!k || klen == 0 || !out) return fail(FFTCONV_CUDA_ERR_ARG, "filter_create: empty kernel or null output");
  DeviceCtx *c = nullptr;
  if (int r = get_ctx(&c)) return r;
  std::lock_guard<std::mutex> lock(c->mu);
  // a handle is for repeated, large calls: the segment size of the throughput rule
  const size_t N = throughput_oa_size<T>(*c, klen, 1 << 16, 1 << 16, optimal_fft_size(klen));
  if (N > max_generic_fft<T>(*c, klen))
    return fail(FFTCONV_CUDA_ERR_UNSUPPORTED, "filter_create: " + std::to_string(klen) + " taps need FFT size " +
                                                 std::to_string(N) + ", beyond the single-CTA FFT");
  Tables *tab = nullptr;
  if (int r = get_tables<T>(*c, N, &tab)) return r;
  std::unique_ptr<Filter> f(new Filter());
  f->dtype = dtype_id<T>();
  f->device = c->device;
  f->klen = klen;
  f->N = N;
  f->layout = oa_layout_for<T>(*c, N, klen);
  auto body = [&]() -> int {
    CU_TRY(cudaMalloc(&f->taps, klen * sizeof(T)));
    CU_TRY(cudaMalloc(&f->H, N * sizeof(cx<T>)));
    CU_TRY(cudaMemcpy(f->taps, k, klen * sizeof(T), cudaMemcpyDefault));
    const int warps = fc::generic::kSpectrumWarps;
    fc::generic::spectrum_kernel<T><<<(unsigned)((N + warps - 1) / warps), warps * 32>>>(
        static_cast<const T *>(f->taps), (int)klen, (int)N, tab->Wd, f->layout, 1.0 / (double)N,
        static_cast<cx<T> *>(f->H));
    CU_TRY(cudaGetLastError());
    ++g_launches;
    CU_TRY(cudaDeviceSynchronize());
    return 0;
  };
  if (int r = body()) { // no leak on any failure path
    cudaFree(f->taps);
    cudaFree(f->H);
    return r;
  }
  *out = f.release();
  return 0;
}

template <typename T>
int filter_oaconvolve(const void *filter, const T *a, size_t batch, size_t n, size_t a_stride, T *out,
                      size_t out_len, size_t out_stride, int mode, void *stream) {
  const Filter *f = static_cast<const Filter *>(filter);
  if (!f || f->dtype != dtype_id<T>()) return fail(FFTCONV_CUDA_ERR_ARG, "filter handle of the wrong precision");
  tl_filter = f;
  const int r = entry<T>(OP_OA, a, batch, n, a_stride, static_cast<const T *>(f->taps), f->klen, out, out_len,
                         out_stride, mode, stream);
  tl_filter = nullptr;
  return r;
}

struct StreamState {
  std::mutex mu; // push / flush of ONE state are serialised (they reorder the carry buffers)
  const Filter *f = nullptr;
  size_t rows = 0;
  void *carry[2] = {nullptr, nullptr};
  int cur = 0;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
};

// the current device must be the filter's: the carry buffers and every later push live there
int check_filter_device(const Filter *f, const char *what) {
  int dev = -1;
  CU_TRY(cudaGetDevice(&dev));
  if (dev != f->device)
    return fail(FFTCONV_CUDA_ERR_ARG, std::string(what) + ": the filter lives on device " + std::to_string(f->device) +
                                          ", the current device is " + std::to_string(dev));
  return 0;
}

template <typename T> int stream_create(const void *filter, size_t rows, void **out) {
  const Filter *f = static_cast<const Filter *>(filter);
  if (!f || f->dtype != dtype_id<T>() || rows == 0 || !out)
    return fail(FFTCONV_CUDA_ERR_ARG, "stream_create: bad filter handle or zero rows");
  if (int r = check_filter_device(f, "stream_create")) return r;
  std::unique_ptr<StreamState> s(new StreamState());
  s->f = f;
  s->rows = rows;
  const size_t bytes = std::max<size_t>(1, rows * (f->klen - 1)) * sizeof(T);
  auto body = [&]() -> int {
    for (int i = 0; i < 2; ++i) {
      CU_TRY(cudaMalloc(&s->carry[i], bytes));
      CU_TRY(cudaMemset(s->carry[i], 0, bytes));
    }
    CU_TRY(cudaDeviceSynchronize());
    return 0;
  };
  if (int r = body()) {
    cudaFree(s->carry[0]);
    cudaFree(s->carry[1]);
    return r;
  }
  *out = s.release();
  return 0;
}

template <typename T>
int stream_push(void *state, const T *chunk, size_t n, size_t chunk_stride, T *out, size_t out_stride, void *stream) {
  auto *s = static_cast<StreamState *>(state);
  if (!s || s->f->dtype != dtype_id<T>()) return fail(FFTCONV_CUDA_ERR_ARG, "stream state of the wrong precision");
  if (n == 0) return 0;
  if (!is_device_ptr(chunk) || !is_device_ptr(out))
    return fail(FFTCONV_CUDA_ERR_ARG, "stream_push: device pointers only");
  if (int r = check_filter_device(s->f, "stream_push")) return r;
  std::lock_guard<std::mutex> lock(s->mu);
  const size_t tail_len = s->f->klen - 1, w = n + tail_len;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (s->tmp_bytes < s->rows * w * sizeof(T)) CU_TRY(cudaStreamSynchronize(st)); // before the old one is freed
  if (int r = grow(&s->tmp, &s->tmp_bytes, s->rows * w * sizeof(T))) return r;
  if (int r = filter_oaconvolve<T>(s->f, chunk, s->rows, n, chunk_stride, static_cast<T *>(s->tmp), w, w,
                                   FFTCONV_CUDA_FULL, stream))
    return r;
  const long long total = (long long)(s->rows * w);
  const int threads = 256;
  fc::generic::stream_merge_kernel<T><<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(
      static_cast<const T *>(s->tmp), (long long)s->rows, (long long)n, (int)tail_len,
      static_cast<const T *>(s->carry[s->cur]), static_cast<T *>(s->carry[s->cur ^ 1]), out, (long long)out_stride);
  CU_TRY(cudaGetLastError());
  ++g_launches;
  s->cur ^= 1;
  return 0;
}

template <typename T> int stream_flush(void *state, T *tail_out, size_t tail_stride, void *stream) {
  auto *s = static_cast<StreamState *>(state);
  if (!s || s->f->dtype != dtype_id<T>()) return fail(FFTCONV_CUDA_ERR_ARG, "stream state of the wrong precision");
  const size_t tail_len = s->f->klen - 1;
  if (tail_len == 0) return 0;
  if (int r = check_filter_device(s->f, "stream_flush")) return r;
  std::lock_guard<std::mutex> lock(s->mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU_TRY(cudaMemcpy2DAsync(tail_out, tail_stride * sizeof(T), s->carry[s->cur], tail_len * sizeof(T),
                           tail_len * sizeof(T), s->rows, cudaMemcpyDefault, st));
  CU_TRY(cudaMemsetAsync(s->carry[s->cur], 0, s->rows * tail_len * sizeof(T), st));
  return 0;
}

} // namespace

extern "C" {

const char *fftconv_cuda_version(void) { return FFTCONV_CUDA_VERSION; }
const char *fftconv_cuda_last_error(void) { return g_err.c_str(); }
// for the library's other translation units (fftconv_cuda_mgpu.cu); not declared in the public header
void fftconv_cuda_internal_set_error(const char *msg) { g_err = msg ? msg : ""; }
size_t fftconv_cuda_optimal_fft_size(size_t klen) { return optimal_fft_size(klen); }
unsigned long long fftconv_cuda_launch_count(void) { return g_launches.load(); }
int fftconv_cuda_set_reserved_sms(int sms) {
  const int prev = g_reserved_sms.exchange(sms < 0 ? 0 : sms);
  return prev;
}

int fftconv_cuda_clear_cache(void) {
  // Contexts are emptied, not destroyed: a call on another thread that already holds a
  // DeviceCtx pointer keeps a valid object and simply rebuilds what it needs.
  std::lock_guard<std::mutex> map_lock(g_map_mu);
  int cur = 0;
  if (cudaGetDevice(&cur) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  for (auto &kv : g_ctx) {
    DeviceCtx &c = *kv.second;
    std::lock_guard<std::mutex> host_lock(c.host_mu);
    std::lock_guard<std::mutex> lock(c.mu);
    cudaSetDevice(kv.first);
    cudaDeviceSynchronize();
    free_ctx(c);
  }
  cudaSetDevice(cur);
  return 0;
}

int fftconv_cuda_release_stream(void *stream) {
  DeviceCtx *c = nullptr;
  if (int r = get_ctx(&c)) return r;
  std::lock_guard<std::mutex> lock(c->mu);
  auto it = c->ws.find(static_cast<cudaStream_t>(stream));
  if (it == c->ws.end()) return 0;
  CU_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream))); // work in flight may still use the buffers
  free_ws(it->second);
  c->ws.erase(it);
  return 0;
}

#define FFTCONV_FFT_API(SFX, T)                                                                              \
  int fftconv_cuda_fft_c2c_##SFX(const T *in, T *out, size_t n, size_t batch, int sign, void *stream) {      \
    return fft_entry<T>(0, in, out, n, batch, sign, stream);                                                 \
  }                                                                                                          \
  int fftconv_cuda_fft_r2c_##SFX(const T *in, T *out, size_t n, size_t batch, void *stream) {                \
    return fft_entry<T>(1, in, out, n, batch, -1, stream);                                                   \
  }                                                                                                          \
  int fftconv_cuda_fft_c2r_##SFX(const T *in, T *out, size_t n, size_t batch, void *stream) {                \
    return fft_entry<T>(2, in, out, n, batch, +1, stream);                                                   \
  }
FFTCONV_FFT_API(f32, float)
FFTCONV_FFT_API(f64, double)

#define FFTCONV_API(SFX, T)                                                                      \
  int fftconv_cuda_oaconvolve_##SFX(const T *a, size_t n, const T *k, size_t klen, T *out,       \
                                    size_t out_len, int mode) {                                  \
    return entry<T>(OP_OA, a, 1, n, n, k, klen, out, out_len, out_len, mode, nullptr);           \
  }                                                                                              \
  int fftconv_cuda_oaconvolve_batched_##SFX(const T *a, size_t batch, size_t n, size_t a_stride, \
                                            const T *k, size_t klen, T *out, size_t out_len,     \
                                            size_t out_stride, int mode, void *stream) {         \
    return entry<T>(OP_OA, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode, stream); \
  }                                                                                              \
  int fftconv_cuda_convolve_##SFX(const T *a, size_t n, const T *k, size_t klen, T *out,         \
                                  size_t out_len, int mode) {                                    \
    return entry<T>(OP_CONV, a, 1, n, n, k, klen, out, out_len, out_len, mode, nullptr);         \
  }                                                                                              \
  int fftconv_cuda_convolve_batched_##SFX(const T *a, size_t batch, size_t n, size_t a_stride,   \
                                          const T *k, size_t klen, T *out, size_t out_len,       \
                                          size_t out_stride, int mode, void *stream) {           \
    return entry<T>(OP_CONV, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode, stream); \
  }                                                                                              \
  int fftconv_cuda_hilbert_##SFX(const T *x, T *env, size_t n) {                                 \
    return entry<T>(OP_HILBERT, x, 1, n, n, nullptr, 0, env, n, n, 0, nullptr);                  \
  }                                                                                              \
  int fftconv_cuda_hilbert_batched_##SFX(const T *x, size_t batch, size_t n, size_t x_stride,    \
                                         T *env, size_t env_stride, void *stream) {              \
    return entry<T>(OP_HILBERT, x, batch, n, x_stride, nullptr, 0, env, n, env_stride, 0, stream); \
  }                                                                                              \
  int fftconv_cuda_add_##SFX(T *dst, const T *src, size_t n, void *stream) {                     \
    return add_entry<T>(dst, src, n, stream);                                                    \
  }                                                                                              \
  int fftconv_cuda_oa_tail_##SFX(const T *last, const T *k, size_t klen, T *tail, void *stream) { \
    return oa_tail_entry<T>(last, k, klen, tail, stream);                                        \
  }                                                                                              \
  int fftconv_cuda_fill_synthetic_##SFX(T *dst, size_t n, unsigned long long seed,               \
                                        unsigned long long first, double gain, void *stream) {   \
    return fill_synthetic_entry<T>(dst, n, seed, first, gain, stream);                           \
  }

FFTCONV_API(f32, float)
FFTCONV_API(f64, double)

#define FFTCONV_ENVELOPE_API(SFX, T)                                                                     \
  int fftconv_cuda_envelope_batched_##SFX(const T *x, size_t batch, size_t n, size_t x_stride, T *out,   \
                                          size_t out_len, size_t out_stride,                             \
                                          const fftconv_cuda_envelope_opts *opts, void *stream) {        \
    return envelope_entry<T>(OP_HILBERT, x, batch, n, x_stride, nullptr, 0, out, out_len, out_stride,    \
                             opts, stream);                                                              \
  }                                                                                                      \
  int fftconv_cuda_rf_envelope_batched_##SFX(const T *x, size_t batch, size_t n, size_t x_stride,        \
                                             const T *k, size_t klen, T *out, size_t out_len,            \
                                             size_t out_stride, const fftconv_cuda_envelope_opts *opts,  \
                                             void *stream) {                                             \
    return envelope_entry<T>(OP_RF, x, batch, n, x_stride, k, klen, out, out_len, out_stride, opts,      \
                             stream);                                                                    \
  }
FFTCONV_ENVELOPE_API(f32, float)
FFTCONV_ENVELOPE_API(f64, double)

#define FFTCONV_HANDLE_API(SFX, T)                                                                   \
  int fftconv_cuda_filter_create_##SFX(const T *k, size_t klen, void **filter_out) {                \
    return filter_create<T>(k, klen, filter_out);                                                   \
  }                                                                                                  \
  int fftconv_cuda_filter_oaconvolve_##SFX(const void *filter, const T *a, size_t batch, size_t n,  \
                                           size_t a_stride, T *out, size_t out_len,                  \
                                           size_t out_stride, int mode, void *stream) {              \
    return filter_oaconvolve<T>(filter, a, batch, n, a_stride, out, out_len, out_stride, mode, stream); \
  }                                                                                                  \
  int fftconv_cuda_stream_create_##SFX(const void *filter, size_t rows, void **state_out) {         \
    return stream_create<T>(filter, rows, state_out);                                               \
  }                                                                                                  \
  int fftconv_cuda_stream_push_##SFX(void *state, const T *chunk, size_t n, size_t chunk_stride,    \
                                     T *out, size_t out_stride, void *stream) {                      \
    return stream_push<T>(state, chunk, n, chunk_stride, out, out_stride, stream);                  \
  }                                                                                                  \
  int fftconv_cuda_stream_flush_##SFX(void *state, T *tail_out, size_t tail_stride, void *stream) { \
    return stream_flush<T>(state, tail_out, tail_stride, stream);                                   \
  }
FFTCONV_HANDLE_API(f32, float)
FFTCONV_HANDLE_API(f64, double)

int fftconv_cuda_filter_destroy(void *filter) {
  auto *f = static_cast<Filter *>(filter);
  if (!f) return 0;
  cudaFree(f->taps);
  cudaFree(f->H);
  delete f;
  return 0;
}
int fftconv_cuda_stream_destroy(void *state) {
  auto *s = static_cast<StreamState *>(state);
  if (!s) return 0;
  cudaFree(s->carry[0]);
  cudaFree(s->carry[1]);
  cudaFree(s->tmp);
  delete s;
  return 0;
}

} // extern "C"
