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

static long long g_units = 0, g_bulk = 0, g_fast_out = 0, g_interior = 0;

static void run_warp(const W1kParams &p, long long u0, long long u1, const cpair *tw, const SpectrumTables &hs4) {
  std::vector<float> xb(kWarpFloats + 8);
  std::vector<Regs> regs(32);
  for (long long u = u0; u < u1; ++u) {
    ++g_units;
    const UnitDesc d = describe_unit(p, u);
    for (auto &v : xb) v = NAN;                       // nothing left over may be consumed
    // the kernel's own request logic (item_request): bulk copies where it says so, lanes otherwise
    const long long it0 = 2 * u, it1 = 2 * u + 1;
    const int row0 = (int)(it0 / p.segs), seg0 = (int)(it0 - (long long)row0 * p.segs);
    const int row1 = it1 < p.n_items ? (int)(it1 / p.segs) : (int)p.batch;
    const int seg1 = it1 < p.n_items ? (int)(it1 - (long long)row1 * p.segs) : 0;
    const float *tensor_end = p.in + (p.batch - 1) * p.in_stride + p.n;
    const ItemReq rq[2] = {item_request(p, row0, seg0, tensor_end), item_request(p, row1, seg1, tensor_end)};
    const bool bulk = rq[0].in_ok && rq[1].in_ok;
    { // the kernel's shortcut for interior blocks (interior_seg_hi) must agree with item_request
      const int hi = interior_seg_hi(p), rows[2] = {row0, row1}, segs2[2] = {seg0, seg1};
      for (int it = 0; it < 2; ++it) {
        if (!(segs2[it] >= 1 && segs2[it] <= hi && rows[it] < p.batch)) continue;
        const float *f = p.in + ((long long)rows[it] * p.in_stride + ((long long)segs2[it] * p.step - (p.klen - 1)));
        const int ish = align_shift4(f);
        const ItemReq &r = rq[it];
        if (!(r.in_ok && r.out_ok && r.head_zero == 0 && r.zero_from == kN + 4 && r.shift == ish && r.dst_first == 0 &&
              r.src == f - ish && r.bytes == 4 * kN + (ish ? 16 : 0))) { std::printf("interior shortcut disagrees\n"); std::exit(1); }
        ++g_interior;
      }
    }
    int shift[2] = {d.item[0].shift, d.item[1].shift};
    if (bulk) {                                        // what lane 0 asks the TMA for + the lanes' fills
      ++g_bulk;
      for (int it = 0; it < 2; ++it) {
        float *plane = xb.data() + it * kPlane;
        if ((reinterpret_cast<unsigned long long>(rq[it].src) & 15ULL) || (rq[it].bytes & 15) || (rq[it].dst_first & 3) ||
            rq[it].src < p.in || rq[it].src + rq[it].bytes / 4 > tensor_end) { std::printf("bad bulk request\n"); std::exit(1); }
        for (int i = 0; i < rq[it].head_zero; ++i) plane[i] = 0.0f;          // at request time
        std::memcpy(plane + rq[it].dst_first, rq[it].src, rq[it].bytes);   // the copy lands
        for (int lane = 0; lane < 32; ++lane) zero_tail(lane, rq[it].zero_from, plane);
        shift[it] = rq[it].shift;
      }
    } else {
      for (int it = 0; it < 2; ++it)
        for (int lane = 0; lane < 32; ++lane) manual_fill(lane, p, d.item[it], xb.data() + it * kPlane);
    }
    for (int lane = 0; lane < 32; ++lane) load_inputs(lane, regs[lane], xb.data(), shift[0], shift[1]);
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
      if (split) { // the kernel's per-item paths
        if (rq[0].out_ok) store_item_fast<0>(lane, regs[lane], p.out + (long long)row0 * p.out_stride + ((long long)seg0 * p.step - (p.klen - 1) - p.pad), m);
        else store_item_slow<0>(lane, regs[lane], p, row0, seg0);
        if (rq[1].out_ok) store_item_fast<1>(lane, regs[lane], p.out + (long long)row1 * p.out_stride + ((long long)seg1 * p.step - (p.klen - 1) - p.pad), m);
        else store_item_slow<1>(lane, regs[lane], p, row1, seg1);
      } else if (d.fast_out) store_fast(lane, regs[lane], d, p, m);
      else store_slow_at(lane, regs[lane], p, row0, seg0, row1, seg1);
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

  std::vector<Quad> tw_store(kTwElems / 2); // 16-byte aligned: the kernel reads pairs of rows as one Quad
  struct { cpair *p; cpair *data() const { return p; } } tw{reinterpret_cast<cpair *>(tw_store.data())};
  std::vector<cpair> hs(1024);
  std::vector<Quad> hab(1024); // 512 entries q0, then 512 entries q1 (SpectrumTables)
  Quad *hc = hab.data() + 512;
  fill_tw(tw.data());
  fill_hs_host(k.data(), K, hs.data());
  pack_hs(hs.data(), hab.data(), hc);
  { // the kernel's own spectrum (one warp transforms the taps) against the long-double DFT above; then use it
    std::vector<float> xb(kWarpFloats + 8, NAN);
    std::vector<Regs> regs(32);
    std::vector<Quad> own(1024);
    const SpectrumTables own_t{own.data()};
    for (int lane = 0; lane < 32; ++lane) spectrum_stage(lane, k.data(), K, xb.data());
    for (int lane = 0; lane < 32; ++lane) load_inputs(lane, regs[lane], xb.data(), 0, 0);
    for (int lane = 0; lane < 32; ++lane) { phase0(lane, regs[lane], tw.data()); t1_write(lane, regs[lane], xb.data()); }
    for (int lane = 0; lane < 32; ++lane) { t1_read(lane, regs[lane], xb.data()); spectrum_finish(lane, regs[lane], own_t); }
    double num = 0, den = 0;
    const float *a = reinterpret_cast<const float *>(own.data()), *b = reinterpret_cast<const float *>(hab.data());
    for (int i = 0; i < 1024 * 4; ++i) { num += ((double)a[i] - b[i]) * ((double)a[i] - b[i]); den += (double)b[i] * b[i]; }
    const double serr = std::sqrt(num / (den > 0 ? den : 1));
    if (!(serr < 1e-6)) { std::printf("spectrum by the engine: rel l2 %.3e\n", serr); return 1; }
    if (K % 2) hab = own; // odd K: the engine's own tables; even K: the host's (both paths stay covered)
  }
  const SpectrumTables hs4{hab.data()};
  W1kParams p{};
  fill_params(p, a, out, batch, n, n, out_len, out_len, K, mode);
  for (long long w = 0; w < n_warps; ++w) {
    long long u0, u1;
    warp_range(p.n_units, n_warps, w, u0, u1);
    run_warp(p, u0, u1, tw.data(), hs4);
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
  std::printf("%.3e units %lld bulk %lld fast_out %lld interior %lld bad %lld\n", err, g_units, g_bulk, g_fast_out, g_interior, bad);
  return (err < 2e-6 && std::isfinite(err) && bad == 0) ? 0 : 1;
}
