// fftconv_cuda_mgpu.cu -- one process, several GPUs (include/fftconv_cuda.h, "mgpu" section).
//
// A client of the single-device C ABI plus NCCL's C library.  What is partitioned:
//   * batches: rows are independent (the reference handles one signal per call,
//     /root/reference/include/fftconv/fftconv.hpp:474-480) -> contiguous row blocks, one host
//     thread per device, no collective;
//   * one long signal: the blocks of FFTConvEngine::oaconvolve
//     (/root/reference/include/fftconv/fftconv.hpp:387-410) are independent except for the K-1
//     sample tail block s adds into block s+1 -> contiguous chunks, one per device, and ONE
//     neighbour exchange of K-1 samples per cut: ncclSend / ncclRecv in one group.
// libnccl.so.2 is loaded with dlopen when the first communicator is created; the shared library
// itself has no link-time dependency on NCCL, and single-device users never touch it.
#include "../../include/fftconv_cuda.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

extern "C" void fftconv_cuda_internal_set_error(const char *msg); // fftconv_cuda.cu

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

namespace {

// ---- the slice of NCCL's C API this file uses (nccl.h, 2.x; types are ABI-stable) ----------
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
enum { kNcclFloat32 = 7, kNcclFloat64 = 8 }; // ncclDataType_t
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

// failures go into the library's one thread-local message (fftconv_cuda_last_error)
int mfail(int code, const std::string &msg) {
  fftconv_cuda_internal_set_error(msg.c_str());
  return code;
}
#define MCU(expr)                                                                                   \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return mfail(FFTCONV_CUDA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));     \
  } while (0)
#define MNCCL(expr)                                                                                 \
  do {                                                                                              \
    ncclResult_t r__ = (expr);                                                                      \
    if (r__ != 0)                                                                                   \
      return mfail(FFTCONV_CUDA_ERR_CUDA, std::string(#expr) + ": " +                               \
                                              (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "NCCL error")); \
  } while (0)
// a failure of the single-device ABI: its message is already in place
#define MFC(expr)                        \
  do {                                   \
    int r__ = (expr);                    \
    if (r__ != 0) return r__;            \
  } while (0)

int load_nccl() {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.handle) return 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return mfail(FFTCONV_CUDA_ERR_UNSUPPORTED, std::string("multi-GPU needs NCCL: ") + dlerror());
  NcclApi a;
  a.handle = h;
#define SYM(field, name)                                                                       \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));                               \
  if (!a.field) return mfail(FFTCONV_CUDA_ERR_UNSUPPORTED, std::string("libnccl lacks ") + name);
  SYM(CommInitAll, "ncclCommInitAll")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl = a;
  return 0;
}

struct DevSlot {
  int device = -1;
  cudaStream_t main = nullptr, side = nullptr;
  cudaEvent_t ready = nullptr, exchanged = nullptr;
  void *taps = nullptr;     // K elements
  void *tail = nullptr;     // K-1 elements out
  void *recv = nullptr;     // K-1 elements in
  size_t small_bytes = 0;
  void *in = nullptr, *out = nullptr; // chunk buffers of the host form (grow only)
  size_t in_bytes = 0, out_bytes = 0;
};

struct Comm {
  std::mutex mu; // one operation at a time per communicator
  std::vector<DevSlot> dev;
  std::vector<ncclComm_t> nccl; // empty for a single device
};

class DeviceGuard {
  int prev_ = -1;

public:
  DeviceGuard() { cudaGetDevice(&prev_); }
  ~DeviceGuard() {
    if (prev_ >= 0) cudaSetDevice(prev_);
  }
};

void bounds_rows(size_t batch, int ndev, int d, size_t *lo, size_t *hi) {
  const size_t base = batch / (size_t)ndev, rem = batch % (size_t)ndev;
  *lo = (size_t)d * base + std::min<size_t>((size_t)d, rem);
  *hi = *lo + base + ((size_t)d < rem ? 1 : 0);
}
size_t cut_at(size_t n, int ndev, int r) { // interior cuts at multiples of 4 samples (16-byte aligned float chunks)
  if (r <= 0) return 0;
  if (r >= ndev) return n;
  const unsigned __int128 c = (unsigned __int128)n * (unsigned)r / (unsigned)ndev;
  return std::min<size_t>(n, (size_t)c / 4 * 4);
}

int grow_dev(void **p, size_t *cap, size_t bytes) {
  if (*cap >= bytes) return 0;
  if (*p) MCU(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  MCU(cudaMalloc(p, bytes));
  *cap = bytes;
  return 0;
}

// run f(d) on one host thread per device (the calling thread takes device 0); first error wins
template <class F> int per_device(Comm &c, F &&f) {
  const int nd = (int)c.dev.size();
  std::vector<int> rc(nd, 0);
  std::vector<std::string> msg(nd);
  auto run = [&](int d) {
    DeviceGuard guard;
    if (cudaSetDevice(c.dev[d].device) != cudaSuccess) {
      rc[d] = FFTCONV_CUDA_ERR_CUDA;
      msg[d] = "cudaSetDevice failed";
      return;
    }
    rc[d] = f(d);
    if (rc[d]) msg[d] = fftconv_cuda_last_error(); // thread-local: carry it to the caller's thread
  };
  std::vector<std::thread> th;
  for (int d = 1; d < nd; ++d) th.emplace_back(run, d);
  run(0);
  for (auto &t : th) t.join();
  for (int d = 0; d < nd; ++d)
    if (rc[d]) return mfail(rc[d], "device " + std::to_string(c.dev[d].device) + ": " + msg[d]);
  return 0;
}

template <typename T> struct Api;
template <> struct Api<float> {
  static constexpr int nccl_type = kNcclFloat32;
  static int oa(const float *a, size_t b, size_t n, size_t as, const float *k, size_t kl, float *o, size_t ol, size_t os,
                int m, void *st) {
    return fftconv_cuda_oaconvolve_batched_f32(a, b, n, as, k, kl, o, ol, os, m, st);
  }
  static int hilbert(const float *x, size_t b, size_t n, size_t xs, float *e, size_t es, void *st) {
    return fftconv_cuda_hilbert_batched_f32(x, b, n, xs, e, es, st);
  }
  static int tail(const float *last, const float *k, size_t kl, float *t, void *st) {
    return fftconv_cuda_oa_tail_f32(last, k, kl, t, st);
  }
  static int add(float *d, const float *s, size_t n, void *st) { return fftconv_cuda_add_f32(d, s, n, st); }
};
template <> struct Api<double> {
  static constexpr int nccl_type = kNcclFloat64;
  static int oa(const double *a, size_t b, size_t n, size_t as, const double *k, size_t kl, double *o, size_t ol,
                size_t os, int m, void *st) {
    return fftconv_cuda_oaconvolve_batched_f64(a, b, n, as, k, kl, o, ol, os, m, st);
  }
  static int hilbert(const double *x, size_t b, size_t n, size_t xs, double *e, size_t es, void *st) {
    return fftconv_cuda_hilbert_batched_f64(x, b, n, xs, e, es, st);
  }
  static int tail(const double *last, const double *k, size_t kl, double *t, void *st) {
    return fftconv_cuda_oa_tail_f64(last, k, kl, t, st);
  }
  static int add(double *d, const double *s, size_t n, void *st) { return fftconv_cuda_add_f64(d, s, n, st); }
};

template <typename T>
int batched(Comm &c, bool hilbert, const T *a, size_t batch, size_t n, size_t a_stride, const T *k, size_t klen, T *out,
            size_t out_len, size_t out_stride, int mode) {
  if (!a || !out) return mfail(FFTCONV_CUDA_ERR_ARG, "null data pointer");
  std::lock_guard<std::mutex> lock(c.mu);
  const int nd = (int)c.dev.size();
  return per_device(c, [&](int d) -> int {
    size_t lo, hi;
    bounds_rows(batch, nd, d, &lo, &hi);
    if (hi == lo) return 0;
    // host pointers: the single-device call runs its own pipelined transfer on this device
    if (hilbert) MFC(Api<T>::hilbert(a + lo * a_stride, hi - lo, n, a_stride, out + lo * out_stride, out_stride, nullptr));
    else MFC(Api<T>::oa(a + lo * a_stride, hi - lo, n, a_stride, k, klen, out + lo * out_stride, out_len, out_stride, mode, nullptr));
    return 0;
  });
}

// device-resident long signal: everything is enqueued from the calling thread
template <typename T>
int long_device(Comm &c, const T *const *a_chunks, const size_t *chunk_len, const T *k, size_t klen,
                T *const *out_chunks) {
  const int nd = (int)c.dev.size();
  if (!a_chunks || !chunk_len || !k || !out_chunks || klen == 0) return mfail(FFTCONV_CUDA_ERR_ARG, "null argument");
  const size_t tail_len = klen - 1;
  for (int d = 0; d < nd; ++d) {
    if (!a_chunks[d] || !out_chunks[d]) return mfail(FFTCONV_CUDA_ERR_ARG, "null chunk pointer");
    if (chunk_len[d] < std::max<size_t>(klen, 1))
      return mfail(FFTCONV_CUDA_ERR_ARG, "chunk " + std::to_string(d) + " is shorter than the kernel");
  }
  DeviceGuard guard;
  const bool exchange = nd > 1 && tail_len > 0;
  // 1. per device: taps to the device; the tail of the chunk, on the side stream
  for (int d = 0; d < nd; ++d) {
    DevSlot &s = c.dev[d];
    MCU(cudaSetDevice(s.device));
    if (s.small_bytes < klen * sizeof(T)) {
      MCU(cudaStreamSynchronize(s.side));
      MCU(cudaStreamSynchronize(s.main));
      cudaFree(s.taps);
      cudaFree(s.tail);
      cudaFree(s.recv);
      s.taps = s.tail = s.recv = nullptr;
      s.small_bytes = 0;
      MCU(cudaMalloc(&s.taps, klen * sizeof(T)));
      MCU(cudaMalloc(&s.tail, klen * sizeof(T)));
      MCU(cudaMalloc(&s.recv, klen * sizeof(T)));
      s.small_bytes = klen * sizeof(T);
    }
    // the side stream starts after everything already queued on the main stream (the previous
    // call's add reads s.recv; the caller's producer of a_chunks[d] ran on the main stream or earlier)
    MCU(cudaEventRecord(s.ready, s.main));
    MCU(cudaStreamWaitEvent(s.side, s.ready, 0));
    MCU(cudaMemcpyAsync(s.taps, k, klen * sizeof(T), cudaMemcpyHostToDevice, s.side));
    if (exchange && d < nd - 1)
      MFC(Api<T>::tail(a_chunks[d] + (chunk_len[d] - tail_len), static_cast<const T *>(s.taps), klen,
                       static_cast<T *>(s.tail), s.side));
  }
  // 2. one neighbour exchange, all ranks of this process in one NCCL group
  if (exchange) {
    MNCCL(g_nccl.GroupStart());
    for (int d = 0; d < nd; ++d) {
      DevSlot &s = c.dev[d];
      if (d < nd - 1) MNCCL(g_nccl.Send(s.tail, tail_len, Api<T>::nccl_type, d + 1, c.nccl[d], s.side));
      if (d > 0) MNCCL(g_nccl.Recv(s.recv, tail_len, Api<T>::nccl_type, d - 1, c.nccl[d], s.side));
    }
    MNCCL(g_nccl.GroupEnd());
  }
  // 3. the chunks themselves on the main streams (persistent kernel: leave the exchange its SMs)
  const int prev_reserved = exchange ? fftconv_cuda_set_reserved_sms(nd == 2 ? 1 : 2) : 0;
  int rc = 0;
  for (int d = 0; d < nd && !rc; ++d) {
    DevSlot &s = c.dev[d];
    if (cudaSetDevice(s.device) != cudaSuccess) {
      rc = mfail(FFTCONV_CUDA_ERR_CUDA, "cudaSetDevice failed");
      break;
    }
    const size_t m = chunk_len[d];
    rc = Api<T>::oa(a_chunks[d], 1, m, m, k, klen, out_chunks[d], m + tail_len, m + tail_len, FFTCONV_CUDA_FULL, s.main);
  }
  if (exchange) fftconv_cuda_set_reserved_sms(prev_reserved);
  if (rc) return rc;
  // 4. the received tail is added to the head of the chunk
  for (int d = 0; d < nd; ++d) {
    DevSlot &s = c.dev[d];
    MCU(cudaSetDevice(s.device));
    MCU(cudaEventRecord(s.exchanged, s.side));
    MCU(cudaStreamWaitEvent(s.main, s.exchanged, 0));
    if (exchange && d > 0) MFC(Api<T>::add(out_chunks[d], static_cast<const T *>(s.recv), tail_len, s.main));
  }
  return 0;
}

int sync_all(Comm &c) {
  DeviceGuard guard;
  for (auto &s : c.dev) {
    MCU(cudaSetDevice(s.device));
    MCU(cudaStreamSynchronize(s.side));
    MCU(cudaStreamSynchronize(s.main));
  }
  return 0;
}

template <typename T> int long_host(Comm &c, const T *a, size_t n, const T *k, size_t klen, T *out) {
  if (!a || !out || !k || klen == 0) return mfail(FFTCONV_CUDA_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lock(c.mu);
  const int nd = (int)c.dev.size();
  const size_t tail_len = klen - 1;
  std::vector<size_t> lo(nd), len(nd);
  for (int d = 0; d < nd; ++d) {
    lo[d] = cut_at(n, nd, d);
    len[d] = cut_at(n, nd, d + 1) - lo[d];
    if (len[d] < klen) return mfail(FFTCONV_CUDA_ERR_ARG, "signal too short to split across " + std::to_string(nd) + " devices");
  }
  // upload (one thread per device: eight PCIe links at once)
  if (int r = per_device(c, [&](int d) -> int {
        DevSlot &s = c.dev[d];
        if (int g = grow_dev(&s.in, &s.in_bytes, len[d] * sizeof(T))) return g;
        if (int g = grow_dev(&s.out, &s.out_bytes, (len[d] + tail_len) * sizeof(T))) return g;
        MCU(cudaMemcpyAsync(s.in, a + lo[d], len[d] * sizeof(T), cudaMemcpyHostToDevice, s.main));
        return 0;
      }))
    return r;
  std::vector<const T *> ac(nd);
  std::vector<T *> oc(nd);
  for (int d = 0; d < nd; ++d) {
    ac[d] = static_cast<const T *>(c.dev[d].in);
    oc[d] = static_cast<T *>(c.dev[d].out);
  }
  if (int r = long_device<T>(c, ac.data(), len.data(), k, klen, oc.data())) return r;
  return per_device(c, [&](int d) -> int {
    DevSlot &s = c.dev[d];
    const size_t cnt = len[d] + (d == nd - 1 ? tail_len : 0);
    MCU(cudaMemcpyAsync(out + lo[d], s.out, cnt * sizeof(T), cudaMemcpyDeviceToHost, s.main));
    MCU(cudaStreamSynchronize(s.main));
    return 0;
  });
}

void destroy(Comm *c) {
  DeviceGuard guard;
  for (size_t d = 0; d < c->dev.size(); ++d) {
    DevSlot &s = c->dev[d];
    if (cudaSetDevice(s.device) != cudaSuccess) continue;
    if (s.main) cudaStreamSynchronize(s.main);
    if (s.side) cudaStreamSynchronize(s.side);
    if (d < c->nccl.size() && c->nccl[d] && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl[d]);
    cudaFree(s.taps);
    cudaFree(s.tail);
    cudaFree(s.recv);
    cudaFree(s.in);
    cudaFree(s.out);
    if (s.ready) cudaEventDestroy(s.ready);
    if (s.exchanged) cudaEventDestroy(s.exchanged);
    if (s.main) cudaStreamDestroy(s.main);
    if (s.side) cudaStreamDestroy(s.side);
  }
  delete c;
}

int init(int ndev, const int *devices, void **out) {
  if (ndev < 1 || !out) return mfail(FFTCONV_CUDA_ERR_ARG, "mgpu_init: ndev < 1 or null output");
  int visible = 0;
  MCU(cudaGetDeviceCount(&visible));
  std::vector<int> ids(ndev);
  for (int d = 0; d < ndev; ++d) {
    ids[d] = devices ? devices[d] : d;
    if (ids[d] < 0 || ids[d] >= visible)
      return mfail(FFTCONV_CUDA_ERR_ARG, "mgpu_init: device " + std::to_string(ids[d]) + " of " + std::to_string(visible) + " visible");
    for (int e = 0; e < d; ++e)
      if (ids[e] == ids[d]) return mfail(FFTCONV_CUDA_ERR_ARG, "mgpu_init: device listed twice");
  }
  DeviceGuard guard;
  Comm *c = new Comm();
  c->dev.resize(ndev);
  auto body = [&]() -> int {
    for (int d = 0; d < ndev; ++d) {
      DevSlot &s = c->dev[d];
      s.device = ids[d];
      MCU(cudaSetDevice(s.device));
      MCU(cudaStreamCreateWithFlags(&s.main, cudaStreamNonBlocking));
      MCU(cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking));
      MCU(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
      MCU(cudaEventCreateWithFlags(&s.exchanged, cudaEventDisableTiming));
    }
    if (ndev > 1) {
      if (int r = load_nccl()) return r;
      c->nccl.assign(ndev, nullptr);
      MNCCL(g_nccl.CommInitAll(c->nccl.data(), ndev, ids.data()));
    }
    return 0;
  };
  if (int r = body()) {
    const std::string keep = fftconv_cuda_last_error();
    destroy(c);
    return mfail(r, keep);
  }
  *out = c;
  return 0;
}

} // namespace

extern "C" {

int fftconv_cuda_mgpu_init(int ndev, const int *devices, void **comm_out) { return init(ndev, devices, comm_out); }
int fftconv_cuda_mgpu_destroy(void *comm) {
  if (comm) destroy(static_cast<Comm *>(comm));
  return 0;
}
int fftconv_cuda_mgpu_device_count(const void *comm) { return comm ? (int)static_cast<const Comm *>(comm)->dev.size() : 0; }
int fftconv_cuda_mgpu_sync(void *comm) {
  if (!comm) return mfail(FFTCONV_CUDA_ERR_ARG, "null communicator");
  return sync_all(*static_cast<Comm *>(comm));
}
void fftconv_cuda_mgpu_chunk_bounds(size_t n, int ndev, int d, size_t *lo, size_t *hi) {
  if (lo) *lo = cut_at(n, ndev, d);
  if (hi) *hi = cut_at(n, ndev, d + 1);
}
void fftconv_cuda_mgpu_row_bounds(size_t batch, int ndev, int d, size_t *lo, size_t *hi) {
  size_t l = 0, h = 0;
  if (ndev > 0 && d >= 0 && d < ndev) bounds_rows(batch, ndev, d, &l, &h);
  if (lo) *lo = l;
  if (hi) *hi = h;
}

#define FFTCONV_MGPU_API(SFX, T)                                                                                \
  int fftconv_cuda_mgpu_oaconvolve_batched_##SFX(void *comm, const T *a, size_t batch, size_t n, size_t a_stride, \
                                                 const T *k, size_t klen, T *out, size_t out_len,              \
                                                 size_t out_stride, int mode) {                                \
    if (!comm) return mfail(FFTCONV_CUDA_ERR_ARG, "null communicator");                                        \
    return batched<T>(*static_cast<Comm *>(comm), false, a, batch, n, a_stride, k, klen, out, out_len, out_stride, mode); \
  }                                                                                                            \
  int fftconv_cuda_mgpu_hilbert_batched_##SFX(void *comm, const T *x, size_t batch, size_t n, size_t x_stride, \
                                              T *env, size_t env_stride) {                                     \
    if (!comm) return mfail(FFTCONV_CUDA_ERR_ARG, "null communicator");                                        \
    return batched<T>(*static_cast<Comm *>(comm), true, x, batch, n, x_stride, nullptr, 0, env, n, env_stride, 0); \
  }                                                                                                            \
  int fftconv_cuda_mgpu_oaconvolve_long_##SFX(void *comm, const T *const *a_chunks, const size_t *chunk_len,   \
                                              const T *k, size_t klen, T *const *out_chunks) {                 \
    if (!comm) return mfail(FFTCONV_CUDA_ERR_ARG, "null communicator");                                        \
    Comm &c = *static_cast<Comm *>(comm);                                                                      \
    std::lock_guard<std::mutex> lock(c.mu);                                                                    \
    return long_device<T>(c, a_chunks, chunk_len, k, klen, out_chunks);                                        \
  }                                                                                                            \
  int fftconv_cuda_mgpu_oaconvolve_long_host_##SFX(void *comm, const T *a, size_t n, const T *k, size_t klen,  \
                                                   T *out) {                                                   \
    if (!comm) return mfail(FFTCONV_CUDA_ERR_ARG, "null communicator");                                        \
    return long_host<T>(*static_cast<Comm *>(comm), a, n, k, klen, out);                                       \
  }
FFTCONV_MGPU_API(f32, float)
FFTCONV_MGPU_API(f64, double)

} // extern "C"
