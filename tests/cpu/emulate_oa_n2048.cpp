// CPU emulation of the N = 2048 overlap-add kernel: the very same
// __host__ __device__ phase functions the CUDA kernel runs, executed thread by
// thread with the barriers replaced by loop boundaries.  Checks the FFT
// decomposition, shared-memory layouts, twiddle/spectrum table layouts and the
// stream / warm-up / tail logic against direct convolution in double.
//
// usage: emulate_oa_n2048 batch n K mode(0 full,1 same) chunks n_slots  -> prints rel-L2 error
#include "../../fftconv_b200/csrc/oa_n2048_group.cuh"
#include "../../fftconv_b200/csrc/oa_n2048_tables.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

using namespace fc;
using namespace fc::n2048;

struct HostExec {
  const OaParams &p;
  std::vector<pt4> xb, tails;
  std::vector<Regs> regs;
  int tpts;
  explicit HostExec(const OaParams &pp)
      : p(pp), xb(kN), tails(2 * tail_pts(pp.klen)), regs(kThreads), tpts(tail_pts(pp.klen)) {}
  void iteration(const Lane (&ln)[4], long long j, bool warm, bool add_tail, int parity) {
    for (int t = 0; t < kThreads; ++t) { io_load(t, regs[t], p, ln, j); phase0(t, regs[t], xb.data(), p.tw1); }
    for (int t = 0; t < kThreads; ++t) phase1(t, regs[t], xb.data(), p.tw2);
    for (int t = 0; t < kThreads; ++t) phase2(t, regs[t], xb.data(), p.hs);
    for (int t = 0; t < kThreads; ++t) phase3(t, regs[t], xb.data(), p.tw2);
    // tail reads of this iteration must all see the previous buffer: it is the
    // other half of the ping-pong, so thread order does not matter.
    for (int t = 0; t < kThreads; ++t) {
      phase4(t, regs[t], xb.data(), p.tw1);
      io_store(t, regs[t], p, ln, j, tails.data() + (parity ^ 1) * tpts, tails.data() + parity * tpts,
               add_tail, warm);
    }
  }
};

int main(int argc, char **argv) {
  if (argc < 7) { std::fprintf(stderr, "usage\n"); return 2; }
  const long long batch = atoll(argv[1]), n = atoll(argv[2]);
  const int K = atoi(argv[3]), mode = atoi(argv[4]);
  const long long chunks = atoll(argv[5]), n_slots = atoll(argv[6]);

  std::mt19937 rng(1234);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float> a(batch * n), k(K);
  for (auto &v : a) v = nd(rng);
  for (auto &v : k) v = nd(rng);

  OaParams p{};
  const long long out_len = mode ? n : n + K - 1;
  std::vector<float> out(batch * out_len, NAN);
  std::vector<cpair> tw1(2048), tw2(128), hs(2048);
  fill_tw1(tw1.data());
  fill_tw2(tw2.data());
  fill_hs_host(k.data(), K, hs.data());
  fill_params(p, a.data(), out.data(), batch, n, n, out_len, out_len, K, mode, chunks);
  p.tw1 = tw1.data(); p.tw2 = tw2.data(); p.hs = hs.data();

  for (long long s = 0; s < n_slots; ++s) {
    const long long w0 = p.total_work * s / n_slots, w1 = p.total_work * (s + 1) / n_slots;
    HostExec ex(p);
    run_group(p, w0, w1, ex);
  }

  double num = 0, den = 0;
  const int pad = mode ? K / 2 : 0;
  for (long long b = 0; b < batch; ++b)
    for (long long o = 0; o < out_len; ++o) {
      const long long f = o + pad; // index in the full convolution
      double acc = 0;
      for (int j = 0; j < K; ++j) {
        const long long ai = f - j;
        if (ai >= 0 && ai < n) acc += (double)a[b * n + ai] * (double)k[j];
      }
      const double d = (double)out[b * out_len + o] - acc;
      num += d * d; den += acc * acc;
    }
  const double err = std::sqrt(num / (den > 0 ? den : 1));
  std::printf("%.3e\n", err);
  return (err < 2e-6 && std::isfinite(err)) ? 0 : 1;
}
