// hestonexotics_b200/csrc/normals.cuh
//
// Batched inverse-normal transform for the path kernel: raw shishua words ->
// standard normals, i.e. RNG::setup_u + RNG::setup_g of the reference
// (src/RNG.cpp:28-43: u64 -> [0,1] -> ppnd16) without the 8 MiB buffers.
//
// F32 mode (the reference AS BUILT, as241.f90:20-25): two draws are evaluated
// together in packed single precision (fma.rn.f32x2 -> FFMA2, Blackwell's
// two-wide FP32 FMA), the AS241 coefficients appear as immediates, and the
// central (|q| <= 0.425) and intermediate-tail (r <= 5) rational functions are
// BOTH evaluated and selected, because at 15 % tail probability per draw a warp
// would execute both sides of a branch anyway.  Only the far tail (p < 1.4e-11)
// is a real branch.  q is formed from the high word of the draw (a perturbation
// of < 2^-32, far below single precision), the tail argument min(p, 1-p) from all
// 64 bits.
//
// F64 mode: AS241 in double precision; central region for all draws first, then
// each lane loops over its own tail draws.
#pragma once
#include <stdint.h>

#include "fastmath.cuh"
#include "ppnd16.cuh"

namespace hexo {

// ---- packed f32x2 helpers ---------------------------------------------------
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
#define HEXO_BC(x) pack2((float)(x), (float)(x))

// Horner for two arguments at once, coefficients broadcast as immediates
__device__ __forceinline__ uint64_t horner8x2(uint64_t r, float c7, float c6, float c5, float c4,
                                              float c3, float c2, float c1, float c0) {
  uint64_t v = ffma2(HEXO_BC(c7), r, HEXO_BC(c6));
  v = ffma2(v, r, HEXO_BC(c5));
  v = ffma2(v, r, HEXO_BC(c4));
  v = ffma2(v, r, HEXO_BC(c3));
  v = ffma2(v, r, HEXO_BC(c2));
  v = ffma2(v, r, HEXO_BC(c1));
  v = ffma2(v, r, HEXO_BC(c0));
  return v;
}

__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// far tail of as241.f90:110-114 (r > 5, i.e. p < 1.4e-11) and the p in {0,1} case
// (:99-103): rare, evaluated in plain scalar code
__device__ __noinline__ float ppnd_far_tail_f32(float q, float t) {
  using P = Ppnd;
  if (!(t < 3.0e38f)) return 0.0f;  // v == 0: the reference returns 0 with IFAULT = 1
  const float r = sqrtf(t) - (float)P::SPLIT2;
  const float z =
      horner8<float>(r, (float)P::E7, (float)P::E6, (float)P::E5, (float)P::E4, (float)P::E3,
                     (float)P::E2, (float)P::E1, (float)P::E0) /
      horner8<float>(r, (float)P::F7, (float)P::F6, (float)P::F5, (float)P::F4, (float)P::F3,
                     (float)P::F2, (float)P::F1, 1.0f);
  return q < 0.0f ? -z : z;
}

// Two draws -> two normals, single precision (as-built AS241)
__device__ __forceinline__ void normal2_f32(uint64_t w0, uint64_t w1, float& z0, float& z1) {
  using P = Ppnd;
  const uint32_t hi0 = (uint32_t)(w0 >> 32), lo0 = (uint32_t)w0;
  const uint32_t hi1 = (uint32_t)(w1 >> 32), lo1 = (uint32_t)w1;
  // q = p - 1/2 from the high word
  const uint64_t q2 = fmul2(pack2((float)(int32_t)(hi0 ^ 0x80000000u),
                                  (float)(int32_t)(hi1 ^ 0x80000000u)),
                            HEXO_BC(2.3283064365386963e-10f));  // 2^-32
  float q0, q1;
  unpack2(q2, q0, q1);
  // v = p (q < 0) or ~p ~ 1 - p (q >= 0) as a 64-bit fraction; vf = v 2^-32
  const uint32_t f0 = (uint32_t)((int32_t)hi0 >> 31), f1 = (uint32_t)((int32_t)hi1 >> 31);
  const uint64_t vf2 = ffma2(pack2((float)(lo0 ^ f0), (float)(lo1 ^ f1)),
                             HEXO_BC(2.3283064365386963e-10f),
                             pack2((float)(hi0 ^ f0), (float)(hi1 ^ f1)));
  float vf0, vf1;
  unpack2(vf2, vf0, vf1);
  // t = -ln(v 2^-64) = (32 - lg2(vf)) ln2  >= 0
  const uint64_t t2 = ffma2(pack2(mufu_lg2(vf0), mufu_lg2(vf1)), HEXO_BC(-0.69314718055994530942f),
                            HEXO_BC(22.180709777918249f));
  float t0, t1;
  unpack2(t2, t0, t1);
  const float r0 = mufu_sqrt(t0), r1 = mufu_sqrt(t1);
  // central: q A(rc)/B(rc), rc = 0.180625 - q^2     (as241.f90:88-92)
  const uint64_t rc = pack2(fmaf(-q0, q0, (float)P::CONST1), fmaf(-q1, q1, (float)P::CONST1));
  uint64_t numc = horner8x2(rc, (float)P::A7, (float)P::A6, (float)P::A5, (float)P::A4,
                            (float)P::A3, (float)P::A2, (float)P::A1, (float)P::A0);
  const uint64_t denc = horner8x2(rc, (float)P::B7, (float)P::B6, (float)P::B5, (float)P::B4,
                                  (float)P::B3, (float)P::B2, (float)P::B1, 1.0f);
  numc = fmul2(numc, q2);
  // intermediate tail: C(r-1.6)/D(r-1.6)             (as241.f90:104-109)
  const uint64_t rm = fadd2(pack2(r0, r1), HEXO_BC(-(float)P::CONST2));
  const uint64_t numm = horner8x2(rm, (float)P::C7, (float)P::C6, (float)P::C5, (float)P::C4,
                                  (float)P::C3, (float)P::C2, (float)P::C1, (float)P::C0);
  const uint64_t denm = horner8x2(rm, (float)P::D7, (float)P::D6, (float)P::D5, (float)P::D4,
                                  (float)P::D3, (float)P::D2, (float)P::D1, 1.0f);
  float nc0, nc1, dc0, dc1, nm0, nm1, dm0, dm1;
  unpack2(numc, nc0, nc1);
  unpack2(denc, dc0, dc1);
  unpack2(numm, nm0, nm1);
  unpack2(denm, dm0, dm1);
  // sign of q onto the (positive) tail value           (as241.f90:116)
  nm0 = __uint_as_float(__float_as_uint(nm0) ^ (__float_as_uint(q0) & 0x80000000u));
  nm1 = __uint_as_float(__float_as_uint(nm1) ^ (__float_as_uint(q1) & 0x80000000u));
  const bool c0 = fabsf(q0) <= (float)P::SPLIT1, c1 = fabsf(q1) <= (float)P::SPLIT1;
  const uint64_t z2 = fmul2(pack2(c0 ? nc0 : nm0, c1 ? nc1 : nm1),
                            pack2(mufu_rcp(c0 ? dc0 : dm0), mufu_rcp(c1 ? dc1 : dm1)));
  unpack2(z2, z0, z1);
  if (fmaxf(t0, t1) > 25.0f) {  // r > 5, i.e. p < 1.4e-11: essentially never
    if (t0 > 25.0f) z0 = ppnd_far_tail_f32(q0, t0);
    if (t1 > 25.0f) z1 = ppnd_far_tail_f32(q1, t1);
  }
}

// ---- double precision ---------------------------------------------------------

// central region only; *tail is set when the draw needs the tail formula
__device__ __forceinline__ double normal_central_f64(uint64_t w, bool& tail) {
  using P = Ppnd;
  const double q = u64_to_unit(w) - 0.5;
  tail = fabs(q) > P::SPLIT1;
  const double r = P::CONST1 - q * q;
  return q * horner8<double>(r, P::A7, P::A6, P::A5, P::A4, P::A3, P::A2, P::A1, P::A0) *
         fast_rcp(horner8<double>(r, P::B7, P::B6, P::B5, P::B4, P::B3, P::B2, P::B1, 1.0));
}

// as241.f90:94-116 for a draw already known to be outside the central region
__device__ __forceinline__ double normal_tail_f64(uint64_t w) {
  using P = Ppnd;
  const double p = u64_to_unit(w);
  const double q = p - 0.5;
  double r = (q < 0.0) ? p : 1.0 - p;
  if (r <= 0.0) return 0.0;
  r = fast_sqrt(-log(r));
  double z;
  if (r <= P::SPLIT2) {
    r -= P::CONST2;
    z = horner8<double>(r, P::C7, P::C6, P::C5, P::C4, P::C3, P::C2, P::C1, P::C0) *
        fast_rcp(horner8<double>(r, P::D7, P::D6, P::D5, P::D4, P::D3, P::D2, P::D1, 1.0));
  } else {
    r -= P::SPLIT2;
    z = horner8<double>(r, P::E7, P::E6, P::E5, P::E4, P::E3, P::E2, P::E1, P::E0) /
        horner8<double>(r, P::F7, P::F6, P::F5, P::F4, P::F3, P::F2, P::F1, 1.0);
  }
  return (q < 0.0) ? -z : z;
}

}  // namespace hexo
