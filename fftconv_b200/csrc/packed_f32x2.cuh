// packed_f32x2.cuh -- two-lane FP32 arithmetic for sm_100a.
//
// Blackwell has packed FP32 instructions (PTX add/sub/mul/fma .f32x2 -> SASS
// FADD2 / FMUL2 / FFMA2) that process a 64-bit register pair per lane in one
// issue slot, and FMUL2/FFMA2 accept a 32-bit scalar operand broadcast to both
// halves.  The overlap-add kernels run TWO independent FFTs per thread, one in
// each half of every register pair, so every butterfly is one packed
// instruction for both FFTs and every twiddle is a broadcast scalar.
//
// The same functions compile for the host (plain float math) so that the whole
// kernel can be emulated thread-by-thread on the CPU in tests.
#pragma once

#if defined(__CUDACC__)
#define FC_HD __host__ __device__ __forceinline__
#else
#define FC_HD inline
#endif

namespace fc {

struct f2 {
  float x, y; // lane A, lane B
};

FC_HD f2 make_f2(float a, float b) {
  f2 r;
  r.x = a;
  r.y = b;
  return r;
}

#if defined(__CUDA_ARCH__)
#define FC_PK2(a, b)                                                           \
  "{.reg .b64 pa, pb, pc;\n\t"                                                 \
  "mov.b64 pa, {%2, %3};\n\t"                                                  \
  "mov.b64 pb, {%4, %5};\n\t"
FC_HD f2 add2(f2 a, f2 b) {
  f2 r;
  asm(FC_PK2(a, b) "add.rn.f32x2 pc, pa, pb;\n\t"
                   "mov.b64 {%0, %1}, pc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
FC_HD f2 sub2(f2 a, f2 b) {
  f2 r;
  asm(FC_PK2(a, b) "sub.rn.f32x2 pc, pa, pb;\n\t"
                   "mov.b64 {%0, %1}, pc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
// a * s (scalar broadcast to both lanes)
FC_HD f2 muls2(f2 a, float s) {
  f2 r;
  asm("{.reg .b64 pa, pb, pc;\n\t"
      "mov.b64 pa, {%2, %3};\n\t"
      "mov.b64 pb, {%4, %4};\n\t"
      "mul.rn.f32x2 pc, pa, pb;\n\t"
      "mov.b64 {%0, %1}, pc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(s));
  return r;
}
// a * s + c
FC_HD f2 fmas2(f2 a, float s, f2 c) {
  f2 r;
  asm("{.reg .b64 pa, pb, pc, pd;\n\t"
      "mov.b64 pa, {%2, %3};\n\t"
      "mov.b64 pb, {%4, %4};\n\t"
      "mov.b64 pc, {%5, %6};\n\t"
      "fma.rn.f32x2 pd, pa, pb, pc;\n\t"
      "mov.b64 {%0, %1}, pd;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(s), "f"(c.x), "f"(c.y));
  return r;
}
// a * s - c
FC_HD f2 fmss2(f2 a, float s, f2 c) {
  f2 r;
  asm("{.reg .b64 pa, pb, pc, pd;\n\t"
      "mov.b64 pa, {%2, %3};\n\t"
      "mov.b64 pb, {%4, %4};\n\t"
      "mov.b64 pc, {%5, %6};\n\t"
      "mul.rn.f32x2 pd, pa, pb;\n\t"
      "sub.rn.f32x2 pd, pd, pc;\n\t"
      "mov.b64 {%0, %1}, pd;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(s), "f"(c.x), "f"(c.y));
  return r;
}
// full-width forms: both operands are pairs
FC_HD f2 mul2(f2 a, f2 b) {
  f2 r;
  asm(FC_PK2(a, b) "mul.rn.f32x2 pc, pa, pb;\n\t"
                   "mov.b64 {%0, %1}, pc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
// a * b + c
FC_HD f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("{.reg .b64 pa, pb, pc, pd;\n\t"
      "mov.b64 pa, {%2, %3};\n\t"
      "mov.b64 pb, {%4, %5};\n\t"
      "mov.b64 pc, {%6, %7};\n\t"
      "fma.rn.f32x2 pd, pa, pb, pc;\n\t"
      "mov.b64 {%0, %1}, pd;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
// c - a * b (the sign flips are integer-pipe instructions)
FC_HD f2 fnma2(f2 a, f2 b, f2 c) { return fma2(make_f2(-a.x, -a.y), b, c); }
#undef FC_PK2
#else
FC_HD f2 mul2(f2 a, f2 b) { return make_f2(a.x * b.x, a.y * b.y); }
FC_HD f2 fma2(f2 a, f2 b, f2 c) { return make_f2(a.x * b.x + c.x, a.y * b.y + c.y); }
FC_HD f2 fnma2(f2 a, f2 b, f2 c) { return make_f2(c.x - a.x * b.x, c.y - a.y * b.y); }
FC_HD f2 add2(f2 a, f2 b) { return make_f2(a.x + b.x, a.y + b.y); }
FC_HD f2 sub2(f2 a, f2 b) { return make_f2(a.x - b.x, a.y - b.y); }
FC_HD f2 muls2(f2 a, float s) { return make_f2(a.x * s, a.y * s); }
FC_HD f2 fmas2(f2 a, float s, f2 c) { return make_f2(a.x * s + c.x, a.y * s + c.y); }
FC_HD f2 fmss2(f2 a, float s, f2 c) { return make_f2(a.x * s - c.x, a.y * s - c.y); }
#endif

// The same vocabulary for a plain double: the N = 2048 overlap-add engine is a template over
// its data type (f2: two float transforms per thread; double: one double transform).
FC_HD double add2(double a, double b) { return a + b; }
FC_HD double sub2(double a, double b) { return a - b; }
FC_HD double muls2(double a, double s) { return a * s; }
FC_HD double fmas2(double a, double s, double c) { return a * s + c; }
FC_HD double fmss2(double a, double s, double c) { return a * s - c; }
FC_HD void cmul2(double &re, double &im, double wr, double wi) {
  const double nre = re * wr - im * wi;
  im = re * wi + im * wr;
  re = nre;
}

// ... and for a plain float (the register-blocked engine's two-rows-per-FFT instantiation)
FC_HD float add2(float a, float b) { return a + b; }
FC_HD float sub2(float a, float b) { return a - b; }
FC_HD float muls2(float a, float s) { return a * s; }
FC_HD float fmas2(float a, float s, float c) { return a * s + c; }
FC_HD float fmss2(float a, float s, float c) { return a * s - c; }

// Data-type traits of the overlap-add engines: scalar type and how many real streams ("lanes")
// ride in one complex value -- f2: lane 0 = re half A, 1 = im half A, 2 = re half B, 3 = im half B;
// double: lane 0 = re, lane 1 = im.
template <typename D> struct DataTraits;
template <> struct DataTraits<f2> {
  using S = float;
  static constexpr int NL = 4;
  static FC_HD void pack(const float (&x)[4], f2 &re, f2 &im) {
    re = make_f2(x[0], x[2]);
    im = make_f2(x[1], x[3]);
  }
  static FC_HD void unpack(f2 re, f2 im, float (&y)[4]) {
    y[0] = re.x; y[1] = im.x; y[2] = re.y; y[3] = im.y;
  }
};
template <> struct DataTraits<double> {
  using S = double;
  static constexpr int NL = 2;
  static FC_HD void pack(const double (&x)[2], double &re, double &im) {
    re = x[0];
    im = x[1];
  }
  static FC_HD void unpack(double re, double im, double (&y)[2]) {
    y[0] = re; y[1] = im;
  }
};
template <> struct DataTraits<float> { // scalar float: one transform per thread, two real streams (re, im)
  using S = float;
  static constexpr int NL = 2;
};
template <typename S> struct DataOf;            // scalar type -> data type of the engine
template <> struct DataOf<float> { using D = f2; };
template <> struct DataOf<double> { using D = double; };

// (re + i im) *= (wr + i wi), both lanes, scalar twiddle.
FC_HD void cmul2(f2 &re, f2 &im, float wr, float wi) {
  const f2 t = muls2(im, -wi);     // the negation is on the scalar (free for constants)
  const f2 u = muls2(im, wr);
  const f2 nre = fmas2(re, wr, t); // re*wr - im*wi
  im = fmas2(re, wi, u);           // re*wi + im*wr
  re = nre;
}

} // namespace fc
