// envelope_epilogue.cuh -- the store of every Hilbert envelope kernel.
#pragma once

namespace fc {

// Envelope post-processing fused into the store of every Hilbert kernel (SURVEY.md section 8f
// item 3: what the callers of fftconv::hilbert do next -- log compression and decimation of the
// detected envelope; the reference's dead scale_and_magnitude* helpers,
// /root/reference/include/fftconv/fftw.hpp:656-924, show the same intent).  With `plain` set the
// store is env itself at index i: the reference's hilbert (hilbert.hpp:57-63).
//   dec > 1:  only samples i = 0, dec, 2 dec, ... are kept, at index i / dec
//   log:      out = clamp(a * log2(env) + b, 0, 1), i.e. 1 + 20 log10(env / ref) / DR with
//             a = 20 log10(2) / DR and b = 1 - 20 log10(ref) / DR
template <typename T> struct EnvEpilogue {
  int plain, log, dec;
  unsigned magic; // floor(2^32 / dec) + 1 when i / dec == umulhi(i, magic) for every i of a row, else 0
  T a, b;
};
template <typename T> __host__ inline EnvEpilogue<T> plain_epilogue() { return EnvEpilogue<T>{1, 0, 1, 0u, (T)0, (T)0}; }

__device__ __forceinline__ float env_log2(float x) { return __log2f(x); } // MUFU.LG2
__device__ __forceinline__ double env_log2(double x) { return log2(x); }

// PLAIN = true compiles to the bare store (the throughput kernels are instantiated both ways and
// the launcher picks; the small-size kernels branch at run time).
template <bool PLAIN = false, typename T>
__device__ __forceinline__ void env_store(const EnvEpilogue<T> &e, T *row, long long i, T env) {
  if (PLAIN || e.plain) {
    row[i] = env;
    return;
  }
  unsigned j = (unsigned)i;
  if (e.dec > 1) {
    const unsigned q = e.magic ? __umulhi(j, e.magic) : j / (unsigned)e.dec;
    if (q * (unsigned)e.dec != j) return;
    j = q;
  }
  T v = env;
  if (e.log) {
    v = e.a * env_log2(env) + e.b; // env = 0 -> -inf -> 0
    v = v < (T)0 ? (T)0 : (v > (T)1 ? (T)1 : v);
    if (!(v == v)) v = (T)0;
  }
  row[j] = v;
}

} // namespace fc
