// hestonexotics_b200/csrc/normals.cuh
//
// Batched inverse-normal transform for the path kernel: raw shishua words ->
// standard normals, i.e. RNG::setup_u + RNG::setup_g of the reference
// (src/RNG.cpp:28-43: u64 -> [0,1] -> ppnd16) without the 8 MiB buffers.
//
// A branch per draw would make nearly every warp execute both sides (15 % tail
// probability per draw), so a generator round (16 draws per thread) is transformed
// in two phases: the central formula (|q| <= 0.425) for all 16 draws, then each
// lane loops over its OWN tail draws (2.4 on average, so the warp runs the tail
// code ~6 times per round instead of 16).
//
// F32 mode (the reference AS BUILT, as241.f90:20-25): the central phase handles two
// draws per packed single-precision instruction (fma.rn.f32x2 -> FFMA2, Blackwell's
// two-wide FP32 FMA) with the AS241 coefficients as immediates.  q is formed from
// the high word of the draw (a perturbation of < 2^-32, far below single
// precision), the tail argument min(p, 1-p) from all 64 bits.
//
// F64 mode: AS241 in double precision, same two phases.
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

__device__ __forceinline__ uint64_t horner4x2(uint64_t r, float c3, float c2, float c1, float c0) {
  uint64_t v = ffma2(HEXO_BC(c3), r, HEXO_BC(c2));
  v = ffma2(v, r, HEXO_BC(c1));
  return ffma2(v, r, HEXO_BC(c0));
}

// AS241's single-precision routine PPND7 (Wichura 1988: the same algorithm with 7-digit
// coefficients; its printed hash sums are AB 32.3184577772, CD 15.7614929821, checked in
// tests/test_ppnd16.py): degree 3/3 and 3/2 rational functions on the same regions as PPND16.
// The reference carries only PPND16 (src/as241.f90) and evaluates it in single precision as
// built; PPND7 is what the algorithm itself prescribes for single precision.  Optional mode
// HEXO_NORMAL_F32_PPND7; the far tail (p < 1.4e-11) keeps PPND16's E / F coefficients.
struct Ppnd7 {
  static constexpr float A0 = 3.3871327179e+00f, A1 = 5.0434271938e+01f, A2 = 1.5929113202e+02f,
                         A3 = 5.9109374720e+01f, B1 = 1.7895169469e+01f, B2 = 7.8757757664e+01f,
                         B3 = 6.7187563600e+01f;
  static constexpr float C0 = 1.4234372777e+00f, C1 = 2.7568153900e+00f, C2 = 1.3067284816e+00f,
                         C3 = 1.7023821103e-01f, D1 = 7.3700164250e-01f, D2 = 1.2021132975e-01f;
};

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

// Phase 1 (F32 mode): the central formula for two draws at once (as241.f90:88-92).
// rb0/rb1 return words whose SIGN BIT says that the draw is outside the central region and gets
// its value from normal2_tail_mid_f32 afterwards (z0/z1 then hold a finite placeholder).
//   P7 (AS241's own single-precision routine): rb = r = 0.180625 - q^2 (as241.f90:89) --
//   CONST1 = SPLIT1^2, so |q| > 0.425 (:88) exactly when r is negative.
//   PPND16's coefficients in single precision (the reference as built): rb = r + 0.021875, i.e.
//   the central rational function is used up to |q| = 0.45.  AS241 switches formulas at 0.425
//   because that is where BOTH reach 1e-16; the central one degrades smoothly beyond its region
//   (truncation error 3.7e-14 at |q| = 0.43, 5.9e-12 at 0.44, 2.5e-10 at 0.45: a five-hundredth
//   of half an ulp of the single-precision result), while in single precision the tail formula
//   pays for a logarithm and a square root first.  Measured against the as-built oracle on 2^24
//   words the distance does not grow (6.1 single-precision ulps of z at most, 4.7 from the double
//   oracle; with the split at 0.425: 6.2 / 5.6; at 0.46 the rounding of the evaluation itself
//   reaches the stated 8 ulps, so 0.45 it is) -- and a third of the tail draws (15 % -> 10 % of all
//   draws) never enter the tail phase: +3.6 % on the whole kernel.  F64 mode keeps 0.425.
struct F32Split {
  static constexpr float kShift = 0.45f * 0.45f - 0.180625f;  // 0.021875
};
template <bool P7 = false>
__device__ __forceinline__ void normal2_central_f32(uint64_t w0, uint64_t w1, float& z0, float& z1,
                                                    uint32_t& rb0, uint32_t& rb1) {
  using P = Ppnd;
  const uint32_t hi0 = (uint32_t)(w0 >> 32), hi1 = (uint32_t)(w1 >> 32);
  // q = p - 1/2 from the high word of the draw
  const uint64_t q2 = fmul2(pack2((float)(int32_t)(hi0 ^ 0x80000000u),
                                  (float)(int32_t)(hi1 ^ 0x80000000u)),
                            HEXO_BC(2.3283064365386963e-10f));  // 2^-32
  float q0, q1;
  unpack2(q2, q0, q1);
  const float rr0 = fmaf(-q0, q0, (float)P::CONST1), rr1 = fmaf(-q1, q1, (float)P::CONST1);
  const uint64_t rc = pack2(rr0, rr1);
  if (P7) {
    rb0 = __float_as_uint(rr0);
    rb1 = __float_as_uint(rr1);
  } else {
    float s0, s1;
    unpack2(fadd2(rc, HEXO_BC(F32Split::kShift)), s0, s1);
    rb0 = __float_as_uint(s0);
    rb1 = __float_as_uint(s1);
  }
  uint64_t num = P7 ? horner4x2(rc, Ppnd7::A3, Ppnd7::A2, Ppnd7::A1, Ppnd7::A0)
                    : horner8x2(rc, (float)P::A7, (float)P::A6, (float)P::A5, (float)P::A4,
                                (float)P::A3, (float)P::A2, (float)P::A1, (float)P::A0);
  const uint64_t den = P7 ? horner4x2(rc, Ppnd7::B3, Ppnd7::B2, Ppnd7::B1, 1.0f)
                          : horner8x2(rc, (float)P::B7, (float)P::B6, (float)P::B5, (float)P::B4,
                                      (float)P::B3, (float)P::B2, (float)P::B1, 1.0f);
  num = fmul2(num, q2);
  float d0, d1;
  unpack2(den, d0, d1);
  unpack2(fmul2(num, pack2(mufu_rcp(d0), mufu_rcp(d1))), z0, z1);
}

// Phase 2 (F32 mode): draws outside the central region (as241.f90:94-116), two per call.  The
// tail argument min(p, 1-p) uses all 64 bits of the draw: v = p (p < 1/2) or ~p ~ 1 - p as a 64-bit
// fraction, t = -ln(v 2^-64) = (32 - lg2(v 2^-32)) ln2 >= 0, r = sqrt(t), then the intermediate-
// tail formula (:104-109), branch-free; t tells the caller whether the far tail (t > 25, i.e.
// r > 5, p < 1.4e-11, or p in {0,1}) has to replace the value.  Every step that has a packed form
// is done for both draws at once: v 2^-32, t, r - 1.6, the C / D Horner chains (14 FFMA2 instead
// of 28 FFMA) and the final product; fma.rn.f32x2 is two IEEE fmas.
template <bool P7 = false>
__device__ __forceinline__ void normal2_tail_mid_f32(uint64_t w0, uint64_t w1, float& z0, float& z1,
                                                     float& t0, float& t1) {
  using P = Ppnd;
  const uint32_t hi0 = (uint32_t)(w0 >> 32), lo0 = (uint32_t)w0;
  const uint32_t hi1 = (uint32_t)(w1 >> 32), lo1 = (uint32_t)w1;
  const uint32_t f0 = (uint32_t)((int32_t)hi0 >> 31), f1 = (uint32_t)((int32_t)hi1 >> 31);
  float v0, v1;
  unpack2(ffma2(pack2((float)(lo0 ^ f0), (float)(lo1 ^ f1)), HEXO_BC(2.3283064365386963e-10f),
                pack2((float)(hi0 ^ f0), (float)(hi1 ^ f1))),
          v0, v1);
  const uint64_t t2 = ffma2(pack2(mufu_lg2(v0), mufu_lg2(v1)), HEXO_BC(-0.69314718055994530942f),
                            HEXO_BC(22.180709777918249f));
  unpack2(t2, t0, t1);
  const uint64_t r = fadd2(pack2(mufu_sqrt(t0), mufu_sqrt(t1)), HEXO_BC(-(float)P::CONST2));
  const uint64_t num = P7 ? horner4x2(r, Ppnd7::C3, Ppnd7::C2, Ppnd7::C1, Ppnd7::C0)
                          : horner8x2(r, (float)P::C7, (float)P::C6, (float)P::C5, (float)P::C4,
                                      (float)P::C3, (float)P::C2, (float)P::C1, (float)P::C0);
  const uint64_t den =
      P7 ? ffma2(ffma2(HEXO_BC(Ppnd7::D2), r, HEXO_BC(Ppnd7::D1)), r, HEXO_BC(1.0f))
         : horner8x2(r, (float)P::D7, (float)P::D6, (float)P::D5, (float)P::D4, (float)P::D3,
                     (float)P::D2, (float)P::D1, 1.0f);
  float d0, d1, a0, a1;
  unpack2(den, d0, d1);
  unpack2(fmul2(num, pack2(mufu_rcp(d0), mufu_rcp(d1))), a0, a1);
  uint32_t s0, s1;  // sign of q = p - 1/2 (as241.f90:116), one lop3 each: z ^ (~hi & 0x80000000)
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD2;" : "=r"(s0) : "r"(__float_as_uint(a0)), "r"(hi0));
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD2;" : "=r"(s1) : "r"(__float_as_uint(a1)), "r"(hi1));
  z0 = __uint_as_float(s0);
  z1 = __uint_as_float(s1);
}

// The far tail for the same draw (rare: t > 25, i.e. min(p, 1-p) < 1.4e-11, or p in {0, 1}).
// Out here the reference's own arithmetic matters: it forms p = RN(w) 2^-64 in double first
// (RNG.cpp:31), so the top 1024 words ARE p = 1 (-> 0 with IFAULT, as241.f90:99-103) and 1 - p is
// a multiple of 2^-53, while the bit pattern ~w used above is exact.  The scalar as-built routine
// on the reference's p reproduces all of that.
static __device__ __noinline__ float normal_tail_far_f32(uint64_t w, float t) {
  (void)t;
  return (float)ppnd16_f32(u64_to_unit(w));
}

// ---- double precision ---------------------------------------------------------

// Central formula (as241.f90:88-92) for the two draws of a step at once: z = q A(r) / B(r) with
// ONE reciprocal for both -- 1 / (B_v B_x), then times the other draw's B -- which trades a MUFU
// seed (and the move that zeroes its low word) for two multiplications it saves elsewhere.
// rhi0 / rhi1 return the high words of r = 0.180625 - q^2, whose sign bit says that the draw
// needs the tail formula (|q| > 0.425 <=> r < 0); such draws get a finite placeholder here.
__device__ __forceinline__ void normal2_central_f64(uint64_t w0, uint64_t w1, double& z0, double& z1,
                                                    uint32_t& rhi0, uint32_t& rhi1) {
  using P = Ppnd;
  const double q0 = u64_to_unit(w0) - 0.5, q1 = u64_to_unit(w1) - 0.5;
  const double r0 = fma(-q0, q0, P::CONST1), r1 = fma(-q1, q1, P::CONST1);
  rhi0 = (uint32_t)__double2hiint(r0);
  rhi1 = (uint32_t)__double2hiint(r1);
  const double n0 = q0 * horner8<double>(r0, P::A7, P::A6, P::A5, P::A4, P::A3, P::A2, P::A1, P::A0);
  const double n1 = q1 * horner8<double>(r1, P::A7, P::A6, P::A5, P::A4, P::A3, P::A2, P::A1, P::A0);
  const double d0 = horner8<double>(r0, P::B7, P::B6, P::B5, P::B4, P::B3, P::B2, P::B1, 1.0);
  const double d1 = horner8<double>(r1, P::B7, P::B6, P::B5, P::B4, P::B3, P::B2, P::B1, 1.0);
  const double y = fast_rcp(d0 * d1);  // 0.0021 < B < 100 for every |q| <= 1/2, tail draws included
  z0 = n0 * (y * d1);
  z1 = n1 * (y * d0);
}

// AS241's intermediate-tail coefficients (as241.f90:49-63) in the constant bank: there they are
// free `c[3][..]` operands of DFMA; as literals every one of them costs two moves per use (a
// double cannot be an immediate), 34 instructions per evaluation of the tail formula.
struct PpndTailConst {
  double c[8];  // C7 .. C0
  double d[7];  // D7 .. D1
  double const2;
};
static __constant__ PpndTailConst kPpndTail = {
    {Ppnd::C7, Ppnd::C6, Ppnd::C5, Ppnd::C4, Ppnd::C3, Ppnd::C2, Ppnd::C1, Ppnd::C0},
    {Ppnd::D7, Ppnd::D6, Ppnd::D5, Ppnd::D4, Ppnd::D3, Ppnd::D2, Ppnd::D1},
    Ppnd::CONST2};

// as241.f90:94-116 for a draw already known to be outside the central region (|q| > 0.425):
// branch-free intermediate tail (:104-109); *rr = sqrt(-ln(min(p,1-p))) tells the caller whether
// the far tail (rr >= 5) or the p in {0,1} case (rr ~ 26) has to replace the value.
//   logtab: shared-space address of the block's logarithm table (fast_neglog_tab)
// The sign of q = p - 1/2 is the top bit of the word (for |q| > 0.425 the rounding of RNG.cpp:31
// cannot change it), so the selection of min(p, 1-p) (:94-98) and the final sign (:116) are
// integer operations; v = 0 (p in {0, 1}) is lifted to the smallest normal double by an integer
// max on its high word, which sends it to the far-tail path like any p < 1.4e-11.
__device__ __forceinline__ double normal_tail_mid_f64(uint64_t w, double& rr, uint32_t logtab) {
  // p = RN(w) 2^-64 (RNG.cpp:31) and 1 - p side by side: the scaling is exact, so the fused form
  // rounds once, exactly like 1.0 - p
  const double dw = __ull2double_rn(w);
  const double p = dw * 5.42101086242752217e-20;
  const double pc = fma(-dw, 5.42101086242752217e-20, 1.0);
  const bool upper = (int32_t)(w >> 32) < 0;  // q > 0
  const int vhi = upper ? __double2hiint(pc) : __double2hiint(p);
  const int vlo = upper ? __double2loint(pc) : __double2loint(p);
  const double v = __hiloint2double(max(vhi, 0x00100000), vlo);
  rr = fast_sqrt(fast_neglog_tab(v, logtab));
  const double r = rr - kPpndTail.const2;
  double num = fma(r, kPpndTail.c[0], kPpndTail.c[1]);
  double den = fma(r, kPpndTail.d[0], kPpndTail.d[1]);
#pragma unroll
  for (int i = 2; i < 8; ++i) num = fma(num, r, kPpndTail.c[i]);
#pragma unroll
  for (int i = 2; i < 7; ++i) den = fma(den, r, kPpndTail.d[i]);
  den = fma(den, r, 1.0);
  const double z = num * fast_rcp(den);
  // -z for q < 0: flip the sign bit where the word's top bit is clear
  return __hiloint2double(__double2hiint(z) ^ (~(int)(w >> 32) & 0x80000000), __double2loint(z));
}

// far tail (:110-114) and p in {0,1} (:99-103) for the same draw (rare)
static __device__ __noinline__ double normal_tail_far_f64(uint64_t w, double rr) {
  using P = Ppnd;
  const double p = u64_to_unit(w);
  const double q = p - 0.5;
  const double v = (q < 0.0) ? p : 1.0 - p;
  if (v <= 0.0) return 0.0;
  const double r = rr - P::SPLIT2;
  const double z = horner8<double>(r, P::E7, P::E6, P::E5, P::E4, P::E3, P::E2, P::E1, P::E0) /
                   horner8<double>(r, P::F7, P::F6, P::F5, P::F4, P::F3, P::F2, P::F1, 1.0);
  return (q < 0.0) ? -z : z;
}

}  // namespace hexo
