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
  double log1p_c[2];                   // -1/6, 1/5: the top of log1p's series in fast_neglog_tab
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
    {-1.0 / 6.0, 0.2},
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

// ln(y) for normal y > 0 on the rare psi >= 1.5 path (one or two lanes of a warp at a time, so a
// constant-bank table indexed per lane costs nothing): y = 2^e f, f in [1, 2), table entry i =
// (1/c_i rounded to double, -ln of that rounded value) for c_i = 1 + (i + 1/2)/128, so that
// r = f / c_i - 1 comes out of ONE fma with |r| <= 0.0039 and ln f = ln c_i + log1p(r) with
// log1p's series up to r^6 (next term 2e-18).  No reciprocal, half the instructions of fast_log.
// Absolute error <= 1.7e-16 max(1, |ln y|); relative 2.6e-16 for |ln y| > 0.01 (for y -> 1 the
// cancellation between e ln2 + ln c_i and log1p(r) leaves <= 2e-16 absolute, which is what
// matters for V').
struct LogTabEntry {
  double invc, lnc;
};
static __constant__ LogTabEntry kLogTab[128] = {
    {0x1.fe01fe01fe020p-1, 0x1.ff00aa2b10ba0p-9},
    {0x1.fa11caa01fa12p-1, 0x1.7dc475f810a69p-7},
    {0x1.f6310aca0dbb5p-1, 0x1.3cea44346a584p-6},
    {0x1.f25f644230ab5p-1, 0x1.b9fc027af919ap-6},
    {0x1.ee9c7f8458e02p-1, 0x1.1b0d98923d97fp-5},
    {0x1.eae807aba01ebp-1, 0x1.58a5bafc8e4d3p-5},
    {0x1.e741aa59750e4p-1, 0x1.95c830ec8e3f2p-5},
    {0x1.e3a9179dc1a73p-1, 0x1.d276b8adb0b56p-5},
    {0x1.e01e01e01e01ep-1, 0x1.075983598e471p-4},
    {0x1.dca01dca01dcap-1, 0x1.253f62f0a1417p-4},
    {0x1.d92f2231e7f8ap-1, 0x1.42edcbea646eep-4},
    {0x1.d5cac807572b2p-1, 0x1.60658a93750c4p-4},
    {0x1.d272ca3fc5b1ap-1, 0x1.7da766d7b12d0p-4},
    {0x1.cf26e5c44bfc6p-1, 0x1.9ab42462033aep-4},
    {0x1.cbe6d9601cbe7p-1, 0x1.b78c82bb0eda0p-4},
    {0x1.c8b265afb8a42p-1, 0x1.d4313d66cb35dp-4},
    {0x1.c5894d10d4986p-1, 0x1.f0a30c01162a4p-4},
    {0x1.c26b5392ea01cp-1, 0x1.0671512ca596fp-3},
    {0x1.bf583ee868d8bp-1, 0x1.14785846742acp-3},
    {0x1.bc4fd65883e7bp-1, 0x1.2266f190a5acdp-3},
    {0x1.b951e2b18ff23p-1, 0x1.303d718e47fd5p-3},
    {0x1.b65e2e3beee05p-1, 0x1.3dfc2b0ecc62ap-3},
    {0x1.b37484ad806cep-1, 0x1.4ba36f39a55e5p-3},
    {0x1.b094b31d922a4p-1, 0x1.59338d9982085p-3},
    {0x1.adbe87f94905ep-1, 0x1.66acd4272ad51p-3},
    {0x1.aaf1d2f87ebfdp-1, 0x1.740f8f54037a3p-3},
    {0x1.a82e65130e159p-1, 0x1.815c0a14357e9p-3},
    {0x1.a574107688a4ap-1, 0x1.8e928de886d41p-3},
    {0x1.a2c2a87c51ca0p-1, 0x1.9bb362e7dfb85p-3},
    {0x1.a01a01a01a01ap-1, 0x1.a8becfc882f19p-3},
    {0x1.9d79f176b682dp-1, 0x1.b5b519e8fb5a6p-3},
    {0x1.9ae24ea5510dap-1, 0x1.c2968558c18c2p-3},
    {0x1.9852f0d8ec0ffp-1, 0x1.cf6354e09c5ddp-3},
    {0x1.95cbb0be377aep-1, 0x1.dc1bca0abec7bp-3},
    {0x1.934c67f9b2ce6p-1, 0x1.e8c0252aa5a60p-3},
    {0x1.90d4f120190d5p-1, 0x1.f550a564b7b37p-3},
    {0x1.8e6527af1373fp-1, 0x1.00e6c45ad501dp-2},
    {0x1.8bfce8062ff3ap-1, 0x1.071b85fcd590dp-2},
    {0x1.899c0f601899cp-1, 0x1.0d46b579ab74bp-2},
    {0x1.87427bcc092b9p-1, 0x1.136870293a8b0p-2},
    {0x1.84f00c2780614p-1, 0x1.1980d2dd4236fp-2},
    {0x1.82a4a0182a4a0p-1, 0x1.1f8ff9e48a2f3p-2},
    {0x1.8060180601806p-1, 0x1.2596010df763ap-2},
    {0x1.7e225515a4f1dp-1, 0x1.2b9303ab89d25p-2},
    {0x1.7beb3922e017cp-1, 0x1.31871c9544185p-2},
    {0x1.79baa6bb6398bp-1, 0x1.3772662bfd85cp-2},
    {0x1.77908119ac60dp-1, 0x1.3d54fa5c1f710p-2},
    {0x1.756cac201756dp-1, 0x1.432ef2a04e813p-2},
    {0x1.734f0c541fe8dp-1, 0x1.49006804009d0p-2},
    {0x1.713786d9c7c09p-1, 0x1.4ec9732600269p-2},
    {0x1.6f26016f26017p-1, 0x1.548a2c3add263p-2},
    {0x1.6d1a62681c861p-1, 0x1.5a42ab0f4cfe2p-2},
    {0x1.6b1490aa31a3dp-1, 0x1.5ff3070a793d4p-2},
    {0x1.691473a88d0c0p-1, 0x1.659b57303e1f2p-2},
    {0x1.6719f3601671ap-1, 0x1.6b3bb2235943dp-2},
    {0x1.6524f853b4aa3p-1, 0x1.70d42e2789236p-2},
    {0x1.63356b88ac0dep-1, 0x1.7664e1239dbcfp-2},
    {0x1.614b36831ae94p-1, 0x1.7bede0a37afbfp-2},
    {0x1.5f66434292dfcp-1, 0x1.816f41da0d495p-2},
    {0x1.5d867c3ece2a5p-1, 0x1.86e919a330ba1p-2},
    {0x1.5babcc647fa91p-1, 0x1.8c5b7c858b48bp-2},
    {0x1.59d61f123ccaap-1, 0x1.91c67eb45a83ep-2},
    {0x1.5805601580560p-1, 0x1.972a341135159p-2},
    {0x1.56397ba7c52e2p-1, 0x1.9c86b02dc0862p-2},
    {0x1.54725e6bb82fep-1, 0x1.a1dc064d5b995p-2},
    {0x1.52aff56a8054bp-1, 0x1.a72a4966bd9e9p-2},
    {0x1.50f22e111c4c5p-1, 0x1.ac718c258b0e5p-2},
    {0x1.4f38f62dd4c9bp-1, 0x1.b1b1e0ebdfc5ap-2},
    {0x1.4d843bedc2c4cp-1, 0x1.b6eb59d3cf35cp-2},
    {0x1.4bd3edda68fe1p-1, 0x1.bc1e08b0dad0ap-2},
    {0x1.4a27fad76014ap-1, 0x1.c149ff115f027p-2},
    {0x1.4880522014880p-1, 0x1.c66f4e3ff6ff9p-2},
    {0x1.46dce34596066p-1, 0x1.cb8e0744d7acap-2},
    {0x1.453d9e2c776cap-1, 0x1.d0a63ae721e64p-2},
    {0x1.43a2730abee4dp-1, 0x1.d5b7f9ae2c684p-2},
    {0x1.420b5265e5951p-1, 0x1.dac353e2c5955p-2},
    {0x1.40782d10e6566p-1, 0x1.dfc859906d5b5p-2},
    {0x1.3ee8f42a5af07p-1, 0x1.e4c71a8687704p-2},
    {0x1.3d5d991aa75c6p-1, 0x1.e9bfa659861f5p-2},
    {0x1.3bd60d9232955p-1, 0x1.eeb20c640ddf3p-2},
    {0x1.3a524387ac822p-1, 0x1.f39e5bc811e5dp-2},
    {0x1.38d22d366088ep-1, 0x1.f884a36fe9ec1p-2},
    {0x1.3755bd1c945eep-1, 0x1.fd64f20f61571p-2},
    {0x1.35dce5f9f2af8p-1, 0x1.011fab125ff8ap-1},
    {0x1.34679ace01346p-1, 0x1.0389eefce633cp-1},
    {0x1.32f5ced6a1dfap-1, 0x1.05f14bd26459cp-1},
    {0x1.3187758e9ebb6p-1, 0x1.0855c884b450ep-1},
    {0x1.301c82ac40260p-1, 0x1.0ab76bece14d2p-1},
    {0x1.2eb4ea1fed14bp-1, 0x1.0d163ccb9d6b8p-1},
    {0x1.2d50a012d50a0p-1, 0x1.0f7241c9b497dp-1},
    {0x1.2bef98e5a3711p-1, 0x1.11cb81787ccf8p-1},
    {0x1.2a91c92f3c105p-1, 0x1.1422025243d45p-1},
    {0x1.293725bb804a5p-1, 0x1.1675cababa60ep-1},
    {0x1.27dfa38a1ce4dp-1, 0x1.18c6e0ff5cf07p-1},
    {0x1.268b37cd60127p-1, 0x1.1b154b57da29ep-1},
    {0x1.2539d7e9177b2p-1, 0x1.1d610fe677003p-1},
    {0x1.23eb79717605bp-1, 0x1.1faa34b87094cp-1},
    {0x1.22a0122a0122ap-1, 0x1.21f0bfc65beecp-1},
    {0x1.21579804855e6p-1, 0x1.2434b6f483934p-1},
    {0x1.2012012012012p-1, 0x1.26762013430e0p-1},
    {0x1.1ecf43c7fb84cp-1, 0x1.28b500df60783p-1},
    {0x1.1d8f5672e4abdp-1, 0x1.2af15f02640acp-1},
    {0x1.1c522fc1ce059p-1, 0x1.2d2b4012edc9dp-1},
    {0x1.1b17c67f2bae3p-1, 0x1.2f62a99509546p-1},
    {0x1.19e0119e0119ep-1, 0x1.3197a0fa7fe6ap-1},
    {0x1.18ab083902bdbp-1, 0x1.33ca2ba328994p-1},
    {0x1.1778a191bd684p-1, 0x1.35fa4edd36ea0p-1},
    {0x1.1648d50fc3201p-1, 0x1.38280fe58797fp-1},
    {0x1.151b9a3fdd5c9p-1, 0x1.3a5373e7ebdf9p-1},
    {0x1.13f0e8d344724p-1, 0x1.3c7c7fff73206p-1},
    {0x1.12c8b89edc0acp-1, 0x1.3ea33936b2f5bp-1},
    {0x1.11a3019a74826p-1, 0x1.40c7a4880dceap-1},
    {0x1.107fbbe011080p-1, 0x1.42e9c6ddf80bfp-1},
    {0x1.0f5edfab325a2p-1, 0x1.4509a5133bb0ap-1},
    {0x1.0e40655826011p-1, 0x1.472743f33aaadp-1},
    {0x1.0d24456359e3ap-1, 0x1.4942a83a2fc07p-1},
    {0x1.0c0a7868b4171p-1, 0x1.4b5bd6956e273p-1},
    {0x1.0af2f722eecb5p-1, 0x1.4d72d3a39fd01p-1},
    {0x1.09ddba6af8360p-1, 0x1.4f87a3f5026e9p-1},
    {0x1.08cabb37565e2p-1, 0x1.519a4c0ba3446p-1},
    {0x1.07b9f29b8eae2p-1, 0x1.53aad05b99b7cp-1},
    {0x1.06ab59c7912fbp-1, 0x1.55b9354b40bcep-1},
    {0x1.059eea0727586p-1, 0x1.57c57f336f191p-1},
    {0x1.04949cc1664c5p-1, 0x1.59cfb25fae87fp-1},
    {0x1.038c6b78247fcp-1, 0x1.5bd7d30e71c73p-1},
    {0x1.02864fc7729e9p-1, 0x1.5ddde57149923p-1},
    {0x1.0182436517a37p-1, 0x1.5fe1edad18919p-1},
    {0x1.0080402010080p-1, 0x1.61e3efda46467p-1}};
__device__ __forceinline__ double fast_log_pos(double y) {
  const int hi = __double2hiint(y);
  const int e = (hi >> 20) - 1023;
  const LogTabEntry t = kLogTab[(hi >> 13) & 127];
  const double f = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(y));
  const double r = fma(f, t.invc, -1.0);
  double p = fma(r, -1.0 / 6.0, 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, kFm.log_c[8]);  // 1/3
  p = fma(p, r, -0.5);
  p = fma(p, r, 1.0);
  const double ed = (double)e;
  return fma(ed, kFm.ln2_hi, fma(ed, kFm.ln2_lo, fma(r, p, t.lnc)));
}

// -ln(y) by the same algorithm for any normal y > 0 with the table in SHARED memory at byte
// address `tab_saddr` (a copy of kLogTab): for code that all 32 lanes run with different table
// indices, which the constant bank would serialise.  The negation rides on operand signs (no
// instruction).  Absolute error <= 1.7e-16 max(1, |ln y|).
__device__ __forceinline__ double fast_neglog_tab(double y, uint32_t tab_saddr) {
  const int hi = __double2hiint(y);
  const int ne = 1023 - (hi >> 20);
  double invc, lnc;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
               : "=d"(invc), "=d"(lnc)
               : "r"(tab_saddr + ((hi >> 9) & (127 << 4))));
  const double f = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(y));
  const double r = fma(f, invc, -1.0);
  double p = fma(r, kFm.log1p_c[0], kFm.log1p_c[1]);  // -1/6, 1/5
  p = fma(p, r, -0.25);
  p = fma(p, r, kFm.log_c[8]);  // 1/3
  p = fma(p, r, -0.5);
  p = fma(p, r, 1.0);
  const double ned = (double)ne;
  return fma(ned, kFm.ln2_hi, fma(ned, kFm.ln2_lo, fma(-r, p, -lnc)));
}

__device__ __forceinline__ void exp_table_init(double* tab, int tid, int nthreads) {
  for (int j = tid; j < 32; j += nthreads) tab[j] = exp2((double)j / 32.0);
}

}  // namespace hexo
