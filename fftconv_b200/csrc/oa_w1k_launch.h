// oa_w1k_launch.h -- what the dispatcher (fftconv_cuda.cu) needs from the translation unit that holds
// the warp-per-FFT kernel (fftconv_cuda_w1k.cu): parameter block, table builder, launcher.
#pragma once

#include "oa_w1k_group.cuh"

#include <cstddef>

namespace fc {
namespace w1k {

size_t launch_smem_bytes();
int warps_per_cta();
// 0 on success, otherwise a cudaError_t.  `sms`: CTAs to launch (one per SM, fewer for small jobs).
int launch(const W1kParams &p, int sms, void *stream);
// fills kTwElems entries (oa_w1k_tables.h: fill_tw)
void build_twiddles(cpair *tw);
constexpr int kTwiddleElems = 36 * 32;

} // namespace w1k
} // namespace fc
