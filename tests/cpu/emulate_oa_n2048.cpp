// CPU emulation of the N = 2048 overlap-add kernel: the very same
// __host__ __device__ functions the CUDA kernel runs (FFT phases, shared-memory
// layouts, packet scheduler, fast/slow I/O paths), executed thread by thread
// with the barriers replaced by loop boundaries and the bulk (TMA) copies by
// memcpy.  Checked against direct convolution in double.
//
// usage: emulate_oa_n2048 batch n K mode(0 full,1 same) unused n_slots [misalign [recomp [f64]]]
//   n_slots:  groups the work is sliced over (the kernel: 4 per SM); the slice starts come from
//             plan_slices exactly as in the library
//   misalign: shifts the input/output base pointers by that many elements (0..3 float, 0..1 double)
//   recomp:   1 (default) = twiddles k = 3, 5..7, 9..15 formed as products in registers
//   f64:      1 = the float64 instantiation (two streams per group); its error bound is 1e-13
#include "../../fftconv_b200/csrc/oa_n2048_group.cuh"
#include "../../fftconv_b200/csrc/oa_n2048_tables.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

using namespace fc;
using namespace fc::n2048;

static long long g_fast_in = 0, g_fast_out = 0, g_iters = 0;

template <typename D> struct HostGroup {
  using S = scalar_of_t<D>;
  using OaParams = OaParamsT<S>;
  using Packet = PacketT<S>;
  using SchedSlot = SchedSlotT<S>;
  using pt = PtT<D>;
  static constexpr int NL = DataTraits<D>::NL;
  const OaParams &p;
  GroupLayout L;
  std::vector<pt> xb, tails;
  S *stage; // aliases xb, exactly as in the kernel
  std::vector<RegsT<D>> regs;
  int tpts;

  explicit HostGroup(const OaParams &pp)
      : p(pp), L(group_layout<S>(pp.klen)), xb(kN), tails(2 * tail_pts(pp.klen)), regs(kThreads),
        tpts(tail_pts(pp.klen)) {
    stage = reinterpret_cast<S *>(xb.data());
    if (NL * L.lane_stride * (int)sizeof(S) > kN * 16) { std::printf("staging does not fit xb\n"); std::exit(4); }
  }

  void poison() { // whatever was in the exchange buffer must never be consumed as data
    for (int i = 0; i < kN * 16 / (int)sizeof(S); ++i) stage[i] = (S)NAN;
  }
  void issue(const Packet *q) { // the bulk (TMA) loads of a fast-path packet
    if (q->valid && all4(q->in_fast))
      for (int l = 0; l < NL; ++l) std::memcpy(stage + l * L.lane_stride, q->in_src[l], q->in_bytes[l]);
  }

  template <bool RECOMP, bool WIDE> void run(SlotStart start) {
    SchedSlot slots[4];
    Packet ring[3];
    std::memset(ring, 0, sizeof(ring));
    std::memset(slots, 0, sizeof(slots));
    for (int l = 0; l < 4; ++l) {
      slots[l].it.init(start.quad, start.u, (int)start.budget);
      fill_packet_lane(p, slots[l], &ring[0], l);
      fill_packet_lane(p, slots[l], &ring[1], l);
    }
    poison();
    issue(&ring[0]);
    int parity = 0, ci = 0;
    std::vector<ThreadMasks> masks(kThreads);
    for (int t = 0; t < kThreads; ++t) masks[t] = make_masks(t, p);
    while (ring[ci].valid) {
      const Packet *cur = &ring[ci];
      ++g_iters;
      const bool in_fast = all4(cur->in_fast), out_fast = all4(cur->out_fast);
      g_fast_in += in_fast;
      g_fast_out += out_fast;
      const bool skip = cur->skip != 0, add_tail = cur->first == 0;
      for (int t = 0; t < kThreads; ++t) {
        if (in_fast) {
          int sh[4] = {cur->in_shift[0], cur->in_shift[1], cur->in_shift[2], cur->in_shift[3]};
          fast_load<WIDE>(t, regs[t], stage, L.lane_stride, sh, masks[t].in_zero);
        } else {
          io_load(t, regs[t], p, cur->ln, cur->u);
        }
        phase0_compute<RECOMP>(t, regs[t], p.tw1);
      }
      poison();
      for (int t = 0; t < kThreads; ++t) phase0_put(t, regs[t], xb.data());
      for (int t = 0; t < kThreads; ++t) phase1<RECOMP>(t, regs[t], xb.data(), p.tw2);
      for (int t = 0; t < kThreads; ++t) phase2(t, regs[t], xb.data(), p.hs);
      for (int t = 0; t < kThreads; ++t) phase3<RECOMP>(t, regs[t], xb.data(), p.tw2);
      for (int t = 0; t < kThreads; ++t) phase4_read<RECOMP>(t, regs[t], xb.data(), p.tw1);
      // --- barrier; thread 0: request the next inputs, describe the iteration after next ---
      const int ni = ci == 2 ? 0 : ci + 1, nni = ni == 2 ? 0 : ni + 1;
      poison();
      issue(&ring[ni]);

      pt *tail_prev = tails.data() + (parity ^ 1) * tpts, *tail_next = tails.data() + parity * tpts;
      for (int t = 0; t < kThreads; ++t) {
        phase4_compute(regs[t]);
        if (out_fast) {
          S *const dst[4] = {cur->out_dst[0] + t, cur->out_dst[1] + t, cur->out_dst[2] + t, cur->out_dst[3] + t};
          fast_store_to<WIDE>(t, regs[t], dst, masks[t], tail_prev, tail_next, add_tail, skip, p.step);
        } else {
          io_store(t, regs[t], p, cur->ln, cur->u, tail_prev, tail_next, add_tail, skip);
        }
      }
      if (ring[ni].valid)
        for (int l = 0; l < 4; ++l) fill_packet_lane(p, slots[l], &ring[nni], l);
      parity ^= 1;
      ci = ni;
    }
  }
};

template <typename D>
int run_case(long long batch, long long n, int K, int mode, long long chunks, long long n_slots, int mis, int recomp,
             double bound) {
  using S = scalar_of_t<D>;
  using cpair = CPairT<S>;
  std::mt19937 rng(1234);
  std::normal_distribution<S> nd((S)0, (S)1);
  const long long out_len = mode ? n : n + K - 1;
  // 16-byte aligned backing stores, then an optional misalignment of the views
  std::vector<S> a_store(batch * n + 8), out_store(batch * out_len + 8, (S)NAN), k(K);
  S *a = a_store.data(), *out = out_store.data();
  while (reinterpret_cast<unsigned long long>(a) & 15ULL) ++a;
  while (reinterpret_cast<unsigned long long>(out) & 15ULL) ++out;
  a += mis;
  out += mis;
  for (long long i = 0; i < batch * n; ++i) a[i] = nd(rng);
  for (auto &v : k) v = nd(rng);

  OaParamsT<S> p{};
  std::vector<cpair> tw1(2048), tw2(128), hs(2048);
  fill_tw1(tw1.data());
  fill_tw2(tw2.data());
  fill_hs_host(k.data(), K, hs.data());
  (void)chunks;
  const SlicePlan plan = plan_slices(p, (const S *)a, out, batch, n, n, out_len, out_len, K, mode, n_slots);
  p.slots = plan.starts.data();
  p.tw1 = tw1.data(); p.tw2 = tw2.data(); p.hs = hs.data();

  for (long long s = 0; s < n_slots; ++s) {
    HostGroup<D> grp(p);
    // kernels of more than 410 taps take the WIDE instantiation, exactly as the dispatcher does
    if (K > kMaxTapsNarrow) grp.template run<true, true>(plan.starts[s]);
    else if (recomp) grp.template run<true, false>(plan.starts[s]);
    else grp.template run<false, false>(plan.starts[s]);
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
  std::printf("%.3e iters %lld fast_in %lld fast_out %lld budget %lld (+1 for %lld of %lld) vrows %lld\n", err, g_iters,
              g_fast_in, g_fast_out, plan.budget, plan.extra, n_slots, plan.vrows);
  return (err < bound && std::isfinite(err)) ? 0 : 1;
}

int main(int argc, char **argv) {
  if (argc < 7) { std::fprintf(stderr, "usage\n"); return 2; }
  const long long batch = atoll(argv[1]), n = atoll(argv[2]);
  const int K = atoi(argv[3]), mode = atoi(argv[4]);
  const long long chunks = atoll(argv[5]), n_slots = atoll(argv[6]);
  const int mis = argc > 7 ? atoi(argv[7]) : 0;
  const int recomp = argc > 8 ? atoi(argv[8]) : 1; // twiddle recomputation (the kernel's default)
  const int f64 = argc > 9 ? atoi(argv[9]) : 0;
  if (f64) return run_case<double>(batch, n, K, mode, chunks, n_slots, mis, recomp, 1e-13);
  return run_case<f2>(batch, n, K, mode, chunks, n_slots, mis, recomp, 2e-6);
}
