// hestonexotics_b200/csrc/fastmath.cuh
//
// Branch-free double-precision reciprocal, square root and exponential for the
// QE step.  CUDA's own `/`, sqrt() and exp() are correctly rounded / <1 ulp but
// each carries a slow path (denormals, huge arguments) behind a divergent branch
// and a call; on the path kernel those branches cost more issue slots than the
// arithmetic.  The versions here start from the 20-bit MUFU seeds
// (rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64 -> MUFU.RCP64H / MUFU.RSQ64H) and
// apply ONE third-order correction, which lands within ~1 ulp for the normal,
// well-scaled arguments the stepper produces (tests/test_gpu_parity.py pins the
// error through hexo_gpu_selftest_math).
#pragma once
#include <stdint.h>

namespace hexo {

// Double-precision literals cannot be instruction immediates (unless their low
// word is zero), and under register pressure ptxas re-materialises them with two
// moves per use inside the step loop.  Living in the constant bank they become
// free `c[3][..]` operands of DFMA/DADD.
struct FmConst {
  double inv40320, inv5040;            // Taylor coefficients of e^r ...
  double inv720, inv120, inv24, inv6;  // ...
  double invL;                         // 32/ln2
  double nLhi, nLlo;                   // -(ln2/32) split in two
  double magic;                        // 1.5 * 2^52
  double tiny;                         // 1e-300: keeps sqrt arguments off exact zero
  double u_max;                        // largest double below 1
  double log_c[9];                     // 1/19, 1/17, ... 1/3: atanh series of fast_log
  double ln2_hi, ln2_lo;               // ln2 split so that e * ln2_hi is exact
  double em1[6];                       // (e^x - 1)/x on |x| <= 0.08: coefficients of x^6 .. x^1
};
// static: one copy per translation unit (the kernels are built in several)
static __constant__ FmConst kFm = {
    1.0 / 40320.0, 1.0 / 5040.0,
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,
    46.166241308446828384,
    -2.16608493865351192653e-02, -5.96317165397058656257e-12,
    6755399441055744.0,
    1e-300,
    0.99999999999999988898,
    {1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0,
     1.0 / 3.0},
    6.93147180369123816490e-01, 1.90821492927058770002e-10,
    // degree-6 interpolant of (e^x - 1)/x at the Chebyshev nodes of [-0.08, 0.08]: e^x = 1 + x p(x)
    // to 7e-16 (relative) on the whole interval, 3e-16 for the |x| < 0.01 of a typical step --
    // one multiply-add fewer than the Taylor polynomial of the same accuracy (degree 7)
    {0.0001984435648549994893, 0.0013891666913593416117, 0.0083333332345585629478,
     0.041666665777675055694, 0.16666666666674568706, 0.50000000000071119961},
};

__device__ __forceinline__ double mufu_rcp64(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  return y;
}
__device__ __forceinline__ double mufu_rsqrt64(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  return y;
}

// 1/a for normal a (not 0, inf, nan, denormal).  y0 has relative error e ~ 2^-20;
// y0 (1 + e + e^2) leaves e^3 ~ 2^-60.
__device__ __forceinline__ double fast_rcp(double a) {
  const double y0 = mufu_rcp64(a);
  const double e = fma(-a, y0, 1.0);
  const double t = fma(e, e, e);
  return fma(y0, t, y0);
}

// sqrt(a) for normal a > 0.  With g = a y0 and r = 1 - g y0, sqrt(a) = g (1-r)^(-1/2)
// = g (1 + r/2 + 3 r^2/8 + O(r^3)), r ~ 2^-19.
__device__ __forceinline__ double fast_sqrt(double a) {
  const double y0 = mufu_rsqrt64(a);
  const double g = a * y0;
  const double r = fma(-g, y0, 1.0);
  const double q = fma(0.375, r, 0.5);
  return fma(g, q * r, g);
}

// sign(a) sqrt(|a|): the same correction seeded from |a| -- the seed instruction only reads the
// high word, so the absolute value is one integer AND, not an FP64 instruction.  For arguments
// that are non-negative up to rounding: a slightly negative one gives a slightly negative
// result instead of NaN.
__device__ __forceinline__ double fast_sqrt_signed(double a) {
  const double y0 = mufu_rsqrt64(__hiloint2double(__double2hiint(a) & 0x7fffffff, 0));
  const double g = a * y0;
  const double r = fma(-fabs(g), y0, 1.0);
  const double q = fma(0.375, r, 0.5);
  return fma(g, q * r, g);
}

// exp(x) for |x| < 700: x = (32 k + j) ln2/32 + r, exp(x) = 2^k 2^(j/32) e^r with
// |r| <= ln2/64 and a degree-6 Taylor polynomial (remainder r^7/5040 < 4e-18).
// `tab_saddr` is the shared-space byte address of the 32-entry table 2^(j/32)
// (filled by exp_table_init).
__device__ __forceinline__ double fast_exp(double x, uint32_t tab_saddr) {
  const double t = fma(x, kFm.invL, kFm.magic);
  const int n = __double2loint(t);                   // round-to-nearest integer of x*32/ln2
  const double nd = t - kFm.magic;
  double r = fma(nd, kFm.nLhi, x);
  r = fma(nd, kFm.nLlo, r);
  double p = fma(r, kFm.inv720, kFm.inv120);
  p = fma(p, r, kFm.inv24);
  p = fma(p, r, kFm.inv6);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = p * r;                                         // e^r - 1
  double T;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(tab_saddr + ((n & 31) << 3)));
  const double v = fma(T, p, T);
  // scale by 2^k: add k to the exponent field (v is in [1,2), k in [-1010,1010])
  const int k = n >> 5;
  return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
}

// ln(y) for normal y > 0, branch-free: y = 2^e f with f in [sqrt(1/2), sqrt(2)),
// ln f = 2 atanh(s), s = (f-1)/(f+1), |s| <= 0.1716, odd series up to s^19 (next term
// s^21/21 < 5e-18); ~1 ulp away from f = 1, relative error ~2e-16 overall.
__device__ __forceinline__ double fast_log(double y) {
  int hi = __double2hiint(y);
  hi += 0x3ff00000 - 0x3fe6a09e;                     // move the split point to sqrt(1/2)
  const int e = (hi >> 20) - 0x3ff;
  hi = (hi & 0x000fffff) + 0x3fe6a09e;
  const double f = __hiloint2double(hi, __double2loint(y));
  const double s = (f - 1.0) * fast_rcp(f + 1.0);
  const double z = s * s;
  double p = fma(z, kFm.log_c[0], kFm.log_c[1]);
#pragma unroll
  for (int i = 2; i < 9; ++i) p = fma(p, z, kFm.log_c[i]);
  const double lf = fma(p * z, s + s, s + s);        // 2 s (1 + z/3 + z^2/5 + ...)
  const double ed = (double)e;
  // e ln2 with ln2 split so that e * hi part is exact
  return fma(ed, kFm.ln2_hi, fma(ed, kFm.ln2_lo, lf));
}

__device__ __forceinline__ void exp_table_init(double* tab, int tid, int nthreads) {
  for (int j = tid; j < 32; j += nthreads) tab[j] = exp2((double)j / 32.0);
}

}  // namespace hexo
