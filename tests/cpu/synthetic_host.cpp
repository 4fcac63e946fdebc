// Host build of the device generator (fftconv_b200/csrc/synthetic.cuh): prints samples for the
// CPU test that pins the numpy twin (fftconv_b200/synthetic.py) to it.
//   synthetic_host <seed> <first> <count> <gain> <f32|f64>
#include "../../fftconv_b200/csrc/synthetic.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char **argv) {
  if (argc < 6) return 2;
  const unsigned long long seed = std::strtoull(argv[1], nullptr, 0), first = std::strtoull(argv[2], nullptr, 0);
  const long count = std::atol(argv[3]);
  const double gain = std::atof(argv[4]);
  const bool f32 = std::strcmp(argv[5], "f32") == 0;
  for (long i = 0; i < count; ++i) {
    if (f32) std::printf("%.9g\n", (double)fc::synthetic::sample<float>(seed, first + i, (float)(fc::synthetic::kScale * gain)));
    else std::printf("%.17g\n", fc::synthetic::sample<double>(seed, first + i, fc::synthetic::kScale * gain));
  }
  return 0;
}
