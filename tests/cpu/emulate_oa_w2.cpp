// CPU emulation of the warp-pair overlap-add kernel (oa_w2_kernel.cuh): the very same
// __host__ __device__ functions the CUDA kernel runs (32-point DFTs, exchange layouts,
// radix-2 split across the two warps, packet scheduler, fast/slow I/O paths), executed
// thread by thread with the barriers replaced by loop boundaries and the bulk (TMA) copies
// by memcpy.  Checked against direct convolution in double.
//
// usage: emulate_oa_w2 batch n K mode(0 full,1 same) chunks n_slots [misalign]
#include "../../fftconv_b200/csrc/oa_w2_group.cuh"
#include "../../fftconv_b200/csrc/oa_w2_tables.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

using fc::f2;
using namespace fc::w2;
using fc::n2048::fill_params;

static long long g_fast_in = 0, g_fast_out = 0, g_iters = 0;

struct HostPair {
  const OaParams &p;
  const cpair *tw, *hs;
  PairLayout L;
  std::vector<pt4> xb, tails; // xb = both warps' buffers, contiguous as in the kernel
  float *stage;
  std::vector<Regs> regs;
  std::vector<Out> outs;
  int tpts;

  HostPair(const OaParams &pp, const cpair *tw_, const cpair *hs_)
      : p(pp), tw(tw_), hs(hs_), L(pair_layout(pp.klen)), xb(2 * kXbPts), tails(2 * tail_pts(pp.klen)), regs(64),
        outs(64), tpts(tail_pts(pp.klen)) {
    stage = reinterpret_cast<float *>(xb.data());
    if (4 * L.lane_stride * 4 > 2 * kXbPts * (int)sizeof(pt4)) { std::printf("staging does not fit\n"); std::exit(4); }
  }
  void poison() {
    for (int i = 0; i < 2 * kXbPts * 4; ++i) stage[i] = NAN;
  }
  void issue(const Packet *q) {
    if (q->valid && all4(q->in_fast))
      for (int l = 0; l < 4; ++l) std::memcpy(stage + l * L.lane_stride, q->in_src[l], q->in_bytes[l]);
  }
  template <class F> void each(F &&f) { // f(w, t, thread index)
    for (int w = 0; w < 2; ++w)
      for (int t = 0; t < 32; ++t) f(w, t, w * 32 + t);
  }

  void run(long long w0, long long w1) {
    SchedSlot slots[4];
    Packet ring[3];
    std::memset(ring, 0, sizeof(ring));
    std::memset(slots, 0, sizeof(slots));
    for (int l = 0; l < 4; ++l) {
      slots[l].it.init(w0, w1);
      fill_packet_lane(p, slots[l], &ring[0], l);
      fill_packet_lane(p, slots[l], &ring[1], l);
    }
    poison();
    issue(&ring[0]);
    int parity = 0, ci = 0;
    pt4 *xbw[2] = {xb.data(), xb.data() + kXbPts};
    while (ring[ci].valid) {
      const Packet *cur = &ring[ci];
      ++g_iters;
      const bool in_fast = all4(cur->in_fast), out_fast = all4(cur->out_fast);
      g_fast_in += in_fast;
      g_fast_out += out_fast;
      const bool warm = cur->warm != 0, add_tail = cur->first == 0;
      each([&](int w, int t, int i) {
        const float sgn = w ? -1.0f : 1.0f;
        if (in_fast) {
          int sh[4] = {cur->in_shift[0], cur->in_shift[1], cur->in_shift[2], cur->in_shift[3]};
          fast_load(t, regs[i], stage, L.lane_stride, sh, p.step, sgn);
        } else {
          io_load(t, regs[i], p, cur->ln, cur->j, sgn);
        }
      });
      poison(); // bar A
      each([&](int w, int t, int idx) {
        if (w) twiddle64(regs[idx]);
      });
      for (int i = 1; i <= 4; ++i) { // the lanes of a warp run a trip in lockstep (__syncwarp around the exchange)
        bool ex[64];
        each([&](int w, int t, int idx) { ex[idx] = trip_head(i, w, t, regs[idx], xbw[w], tw, hs + w * 1024); });
        each([&](int w, int t, int idx) {
          if (ex[idx]) get_rows(t, regs[idx], xbw[w]);
        });
      }
      // the hand-over writes only addresses the writing lane has just read itself
      each([&](int w, int t, int i) { combine_put(t, regs[i], xbw[w]); });
      each([&](int w, int t, int i) { combine_get(t, regs[i], xbw[w ^ 1], w ? -1.0f : 1.0f, outs[i]); }); // bar B
      const int ni = ci == 2 ? 0 : ci + 1, nni = ni == 2 ? 0 : ni + 1;
      poison(); // bar C
      issue(&ring[ni]);
      pt4 *tail_prev = tails.data() + (parity ^ 1) * tpts, *tail_next = tails.data() + parity * tpts;
      each([&](int w, int t, int i) {
        if (out_fast || warm) {
          const int off = t + 512 * w;
          fast_store(t, w, outs[i], cur->out_dst[0] + off, cur->out_dst[1] + off, cur->out_dst[2] + off,
                     cur->out_dst[3] + off, p.klen, p.step, tail_prev, tail_next, add_tail, out_fast);
        } else {
          io_store(t, w, outs[i], p, cur->ln, cur->j, tail_prev, tail_next, add_tail, warm);
        }
      });
      if (ring[ni].valid)
        for (int l = 0; l < 4; ++l) fill_packet_lane(p, slots[l], &ring[nni], l);
      parity ^= 1;
      ci = ni;
    }
  }
};

int main(int argc, char **argv) {
  if (argc < 7) { std::fprintf(stderr, "usage\n"); return 2; }
  const long long batch = atoll(argv[1]), n = atoll(argv[2]);
  const int K = atoi(argv[3]), mode = atoi(argv[4]);
  const long long chunks = atoll(argv[5]), n_slots = atoll(argv[6]);
  const int mis = argc > 7 ? atoi(argv[7]) : 0;

  std::mt19937 rng(1234);
  std::normal_distribution<float> nd(0.f, 1.f);
  const long long out_len = mode ? n : n + K - 1;
  std::vector<float> a_store(batch * n + 8), out_store(batch * out_len + 8, NAN), k(K);
  float *a = a_store.data(), *out = out_store.data();
  while (reinterpret_cast<unsigned long long>(a) & 15ULL) ++a;
  while (reinterpret_cast<unsigned long long>(out) & 15ULL) ++out;
  a += mis;
  out += mis;
  for (long long i = 0; i < batch * n; ++i) a[i] = nd(rng);
  for (auto &v : k) v = nd(rng);

  OaParams p{};
  std::vector<cpair> tw(kTwElems), hs(2048);
  fill_tw(tw.data());
  fc::w2::fill_hs_host(k.data(), K, hs.data());
  fill_params(p, a, out, batch, n, n, out_len, out_len, K, mode, chunks);

  for (long long s = 0; s < n_slots; ++s) {
    const long long w0 = p.total_work * s / n_slots, w1 = p.total_work * (s + 1) / n_slots;
    HostPair pr(p, tw.data(), hs.data());
    pr.run(w0, w1);
  }

  double num = 0, den = 0;
  const int pad = mode ? K / 2 : 0;
  for (long long b = 0; b < batch; ++b)
    for (long long o = 0; o < out_len; ++o) {
      const long long f = o + pad;
      double acc = 0;
      for (int j = 0; j < K; ++j) {
        const long long ai = f - j;
        if (ai >= 0 && ai < n) acc += (double)a[b * n + ai] * (double)k[j];
      }
      const double d = (double)out[b * out_len + o] - acc;
      num += d * d; den += acc * acc;
    }
  const double err = std::sqrt(num / (den > 0 ? den : 1));
  std::printf("%.3e iters %lld fast_in %lld fast_out %lld\n", err, g_iters, g_fast_in, g_fast_out);
  return (err < 2e-6 && std::isfinite(err)) ? 0 : 1;
}
