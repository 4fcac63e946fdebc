// CPU emulation of the warp-per-FFT overlap-add engine (oa_w1k_*.cuh): the same __host__ __device__
// phase functions the CUDA kernel runs, executed lane by lane with __syncwarp() replaced by loop
// boundaries and the bulk copies by memcpy, against direct convolution in double.
//
// usage: emulate_oa_w1k batch n K mode(0 full, 1 same) n_warps [misalign]
#include "../../fftconv_b200/csrc/oa_w1k_group.cuh"
#include "../../fftconv_b200/csrc/oa_w1k_tables.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

using namespace fc;
using namespace fc::w1k;

static long long g_units = 0, g_bulk = 0, g_fast_out = 0;

static void run_warp(const W1kParams &p, long long u0, long long u1, const cpair *tw, const Quad *hs4) {
  std::vector<float> xb(kWarpFloats + 8);
  std::vector<Regs> regs(32);
  for (long long u = u0; u < u1; ++u) {
    ++g_units;
    const UnitDesc d = describe_unit(p, u);
    for (auto &v : xb) v = NAN;                       // nothing left over may be consumed
    if (d.bulk) {                                      // what lane 0 asks the TMA for + the lanes' zero fill
      ++g_bulk;
      for (int it = 0; it < 2; ++it) {
        const ItemDesc &I = d.item[it];
        if (I.bytes) std::memcpy(xb.data() + it * kPlane + I.dst_first, I.src, I.bytes);
        for (int lane = 0; lane < 32; ++lane) zero_fill(lane, I, xb.data() + it * kPlane);
      }
    } else {
      for (int it = 0; it < 2; ++it)
        for (int lane = 0; lane < 32; ++lane) manual_fill(lane, p, d.item[it], xb.data() + it * kPlane);
    }
    for (int lane = 0; lane < 32; ++lane) load_inputs(lane, regs[lane], xb.data(), d.item[0].shift, d.item[1].shift);
    for (auto &v : xb) v = NAN;
    // odd units run the phases as the kernel's two-trip loops cut them, even units the fused forms
    const bool split = (u & 1) != 0;
    for (int lane = 0; lane < 32; ++lane) {
      if (split) { p0_pre(lane, regs[lane], tw); dft16<-1>(regs[lane]); p0_post(lane, regs[lane], tw); }
      else phase0(lane, regs[lane], tw);
      t1_write(lane, regs[lane], xb.data());
    }
    for (int lane = 0; lane < 32; ++lane) {
      t1_read(lane, regs[lane], xb.data());
      if (split) { dft16<-1>(regs[lane]); p1_mid(lane, regs[lane], hs4); dft16<+1>(regs[lane]); }
      else phase1(lane, regs[lane], hs4);
    }
    // T2 writes a lane's own row, which only that lane read in T1: no barrier in between
    for (int lane = 0; lane < 32; ++lane) t2_write(lane, regs[lane], xb.data());
    for (int lane = 0; lane < 32; ++lane) t2_read(lane, regs[lane], xb.data());
    for (int lane = 0; lane < 32; ++lane) {
      if (split) { p2_pre(lane, regs[lane], tw); dft16<+1>(regs[lane]); p2_post(lane, regs[lane], tw); }
      else phase2(lane, regs[lane], tw);
      const StoreMasks m = make_store_masks(lane, p);
      if (d.fast_out) store_fast(lane, regs[lane], d, p, m);
      else if (split) store_slow(lane, regs[lane], p, d);
      else {
        const long long it0 = 2 * u, it1 = 2 * u + 1;
        const long long r0 = it0 / p.segs, r1 = it1 < p.n_items ? it1 / p.segs : p.batch;
        store_slow_at(lane, regs[lane], p, r0, it0 - r0 * p.segs, r1, it1 < p.n_items ? it1 - r1 * p.segs : 0);
      }
    }
    g_fast_out += d.fast_out;
  }
}

int main(int argc, char **argv) {
  if (argc < 6) { std::fprintf(stderr, "usage\n"); return 2; }
  const long long batch = atoll(argv[1]), n = atoll(argv[2]);
  const int K = atoi(argv[3]), mode = atoi(argv[4]);
  const long long n_warps = atoll(argv[5]);
  const int mis = argc > 6 ? atoi(argv[6]) : 0;
  std::mt19937 rng(4321);
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

  std::vector<cpair> tw(kTwElems), hs(1024);
  std::vector<Quad> hs4(512);
  fill_tw(tw.data());
  fill_hs_host(k.data(), K, hs.data());
  pack_hs(hs.data(), hs4.data());
  W1kParams p{};
  fill_params(p, a, out, batch, n, n, out_len, out_len, K, mode);
  for (long long w = 0; w < n_warps; ++w) {
    long long u0, u1;
    warp_range(p.n_units, n_warps, w, u0, u1);
    run_warp(p, u0, u1, tw.data(), hs4.data());
  }
  double num = 0, den = 0;
  const int pad = mode ? K / 2 : 0;
  long long bad = 0;
  for (long long b = 0; b < batch; ++b)
    for (long long o = 0; o < out_len; ++o) {
      const long long f = o + pad;
      double acc = 0;
      for (int j = 0; j < K; ++j) {
        const long long ai = f - j;
        if (ai >= 0 && ai < n) acc += (double)a[b * n + ai] * (double)k[j];
      }
      const double got = (double)out[b * out_len + o];
      if (!std::isfinite(got)) ++bad;
      const double d = got - acc;
      num += d * d; den += acc * acc;
    }
  // guard bands: nothing outside the outputs may be written
  for (float *q = out_store.data(); q < out; ++q) if (!std::isnan(*q)) ++bad;
  for (float *q = out + batch * out_len; q < out_store.data() + out_store.size(); ++q) if (!std::isnan(*q)) ++bad;
  const double err = std::sqrt(num / (den > 0 ? den : 1));
  std::printf("%.3e units %lld bulk %lld fast_out %lld bad %lld\n", err, g_units, g_bulk, g_fast_out, bad);
  return (err < 2e-6 && std::isfinite(err) && bad == 0) ? 0 : 1;
}
