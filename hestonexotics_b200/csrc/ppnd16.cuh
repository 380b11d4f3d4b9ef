// hestonexotics_b200/csrc/ppnd16.cuh
//
// Uniform map and inverse normal CDF of the reference's RNG wrapper, inlined for
// the device: src/RNG.cpp:31 (u64 -> [0,1]) and Wichura's AS241 / PPND16 as the
// reference carries it in src/as241.f90:15-119.
//
// Two arithmetic modes (include/hexo_gpu.h, hexo_normal_mode):
//   F64 -- the algorithm in double precision throughout;
//   F32 -- what the reference computes AS BUILT: as241.f90:20-25 declares all
//          locals and coefficients default REAL, only P and the result are
//          double, so Q/R are rounded to single and the rational functions are
//          evaluated in single precision.
#pragma once
#include <stdint.h>

namespace hexo {

// tests/test_ppnd16.py parses the table below and re-checks the AS241 "hash
// sums" (as241.f90:45,64,83), so keep one coefficient per PPND_COEF line.
#define PPND_COEF(name, value) static constexpr double name = value;
struct Ppnd {
  // p close to 1/2 (as241.f90:31-44)
  PPND_COEF(A0, 3.3871328727963666080e+0)
  PPND_COEF(A1, 1.3314166789178437745e+2)
  PPND_COEF(A2, 1.9715909503065514427e+3)
  PPND_COEF(A3, 1.3731693765509461125e+4)
  PPND_COEF(A4, 4.5921953931549871457e+4)
  PPND_COEF(A5, 6.7265770927008700853e+4)
  PPND_COEF(A6, 3.3430575583588128105e+4)
  PPND_COEF(A7, 2.5090809287301226727e+3)
  PPND_COEF(B1, 4.2313330701600911252e+1)
  PPND_COEF(B2, 6.8718700749205790830e+2)
  PPND_COEF(B3, 5.3941960214247511077e+3)
  PPND_COEF(B4, 2.1213794301586595867e+4)
  PPND_COEF(B5, 3.9307895800092710610e+4)
  PPND_COEF(B6, 2.8729085735721942674e+4)
  PPND_COEF(B7, 5.2264952788528545610e+3)
  // p neither close to 0, 1/2 nor 1 (as241.f90:49-63)
  PPND_COEF(C0, 1.42343711074968357734e+0)
  PPND_COEF(C1, 4.63033784615654529590e+0)
  PPND_COEF(C2, 5.76949722146069140550e+0)
  PPND_COEF(C3, 3.64784832476320460504e+0)
  PPND_COEF(C4, 1.27045825245236838258e+0)
  PPND_COEF(C5, 2.41780725177450611770e-1)
  PPND_COEF(C6, 2.27238449892691845833e-2)
  PPND_COEF(C7, 7.74545014278341407640e-4)
  PPND_COEF(D1, 2.05319162663775882187e+0)
  PPND_COEF(D2, 1.67638483018380384940e+0)
  PPND_COEF(D3, 6.89767334985100004550e-1)
  PPND_COEF(D4, 1.48103976427480074590e-1)
  PPND_COEF(D5, 1.51986665636164571966e-2)
  PPND_COEF(D6, 5.47593808499534494600e-4)
  PPND_COEF(D7, 1.05075007164441684324e-9)
  // p near 0 or 1 (as241.f90:68-82)
  PPND_COEF(E0, 6.65790464350110377720e+0)
  PPND_COEF(E1, 5.46378491116411436990e+0)
  PPND_COEF(E2, 1.78482653991729133580e+0)
  PPND_COEF(E3, 2.96560571828504891230e-1)
  PPND_COEF(E4, 2.65321895265761230930e-2)
  PPND_COEF(E5, 1.24266094738807843860e-3)
  PPND_COEF(E6, 2.71155556874348757815e-5)
  PPND_COEF(E7, 2.01033439929228813265e-7)
  PPND_COEF(F1, 5.99832206555887937690e-1)
  PPND_COEF(F2, 1.36929880922735805310e-1)
  PPND_COEF(F3, 1.48753612908506148525e-2)
  PPND_COEF(F4, 7.86869131145613259100e-4)
  PPND_COEF(F5, 1.84631831751005468180e-5)
  PPND_COEF(F6, 1.42151175831644588870e-7)
  PPND_COEF(F7, 2.04426310338993978564e-15)
  static constexpr double SPLIT1 = 0.425, SPLIT2 = 5.0, CONST1 = 0.180625, CONST2 = 1.6;
};
#undef PPND_COEF

// src/RNG.cpp:31: (double)u64 / (double)(2^64-1).  The divisor rounds to 2^64,
// so the result is RN(u64) * 2^-64, in [0,1] INCLUSIVE.
__device__ __forceinline__ double u64_to_unit(uint64_t x) {
  return __ull2double_rn(x) * 5.42101086242752217e-20;  // 2^-64, exact scaling
}

template <typename T>
__device__ __forceinline__ T horner8(T r, T c7, T c6, T c5, T c4, T c3, T c2, T c1, T c0) {
  T v = c7;
  v = v * r + c6;
  v = v * r + c5;
  v = v * r + c4;
  v = v * r + c3;
  v = v * r + c2;
  v = v * r + c1;
  v = v * r + c0;
  return v;
}

// as241.f90:85-118 in double precision.  p in {0,1} returns 0 like the
// reference does with IFAULT=1 (:99-103; the assert at RNG.cpp:40 is compiled
// out in the release build).
__device__ __forceinline__ double ppnd16_f64(double p) {
  using P = Ppnd;
  const double q = p - 0.5;
  if (fabs(q) <= P::SPLIT1) {
    const double r = P::CONST1 - q * q;
    return q * horner8<double>(r, P::A7, P::A6, P::A5, P::A4, P::A3, P::A2, P::A1, P::A0) /
           horner8<double>(r, P::B7, P::B6, P::B5, P::B4, P::B3, P::B2, P::B1, 1.0);
  }
  double r = (q < 0.0) ? p : 1.0 - p;
  if (r <= 0.0) return 0.0;
  r = sqrt(-log(r));
  double z;
  if (r <= P::SPLIT2) {
    r -= P::CONST2;
    z = horner8<double>(r, P::C7, P::C6, P::C5, P::C4, P::C3, P::C2, P::C1, P::C0) /
        horner8<double>(r, P::D7, P::D6, P::D5, P::D4, P::D3, P::D2, P::D1, 1.0);
  } else {
    r -= P::SPLIT2;
    z = horner8<double>(r, P::E7, P::E6, P::E5, P::E4, P::E3, P::E2, P::E1, P::E0) /
        horner8<double>(r, P::F7, P::F6, P::F5, P::F4, P::F3, P::F2, P::F1, 1.0);
  }
  return (q < 0.0) ? -z : z;
}

// The as-built single-precision evaluation.  P stays double where the Fortran
// mixes it into an expression (Q = P - HALF, R = ONE - P are double operations
// whose result is rounded into a REAL local).
__device__ __forceinline__ double ppnd16_f32(double p) {
  using P = Ppnd;
#define PF(x) ((float)(P::x))
  const float q = (float)(p - 0.5);
  if (fabsf(q) <= PF(SPLIT1)) {
    const float r = PF(CONST1) - q * q;
    const float num = horner8<float>(r, PF(A7), PF(A6), PF(A5), PF(A4), PF(A3), PF(A2), PF(A1), PF(A0));
    const float den = horner8<float>(r, PF(B7), PF(B6), PF(B5), PF(B4), PF(B3), PF(B2), PF(B1), 1.0f);
    return (double)(q * num / den);
  }
  float r = (q < 0.0f) ? (float)p : (float)(1.0 - p);
  if (r <= 0.0f) return 0.0;
  r = sqrtf(-logf(r));
  float z;
  if (r <= PF(SPLIT2)) {
    r -= PF(CONST2);
    z = horner8<float>(r, PF(C7), PF(C6), PF(C5), PF(C4), PF(C3), PF(C2), PF(C1), PF(C0)) /
        horner8<float>(r, PF(D7), PF(D6), PF(D5), PF(D4), PF(D3), PF(D2), PF(D1), 1.0f);
  } else {
    r -= PF(SPLIT2);
    z = horner8<float>(r, PF(E7), PF(E6), PF(E5), PF(E4), PF(E3), PF(E2), PF(E1), PF(E0)) /
        horner8<float>(r, PF(F7), PF(F6), PF(F5), PF(F4), PF(F3), PF(F2), PF(F1), 1.0f);
  }
#undef PF
  return (double)((q < 0.0f) ? -z : z);
}

template <int NORMAL_MODE>
__device__ __forceinline__ double ppnd16(double p) {
  return NORMAL_MODE == 1 ? ppnd16_f64(p) : ppnd16_f32(p);
}

}  // namespace hexo
