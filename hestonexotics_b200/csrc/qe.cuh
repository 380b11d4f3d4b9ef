// hestonexotics_b200/csrc/qe.cuh
//
// Andersen quadratic-exponential (QE) step of the Heston SDE, psi_c = 1.5,
// gamma_1 = gamma_2 = 1/2, no drift and no martingale correction -- the
// discretisation of HQEAnderson::operator++ (src/HSimulation.tpp:52-86,
// PSI_C src/inc/HSimulation.h:12).
//
// Same map (V, ln X, Z_V | U_V, Z_X) -> (V', ln X') as the reference, arranged for
// the FP64 pipe:
//  * everything that depends only on (HParams, step width) is hoisted into
//    SegConst on the host; the reference recomputes it every step (:58,:75-79);
//  * the quadratic branch is rewritten without any division.  With
//    w = m^2 - s^2/2 (so that sqrt(w) = m sqrt(1 - psi/2)):
//        b^2 = 2/psi - 1 + sqrt(2/psi (2/psi - 1))   (:64)
//        a   = m / (1 + b^2)                          (:66)
//    give  a b^2 = sqrt(w),  a = m - sqrt(w),  a b = sqrt(sqrt(w) (m - sqrt(w))),  hence
//        V' = a (b + Z)^2 = sqrt(w) + Z (2 sqrt(sqrt(w) (m - sqrt(w))) + (m - sqrt(w)) Z)   (:68)
//    -- two square roots instead of three divisions and two square roots, on the
//    branch-free fast_sqrt of fastmath.cuh; psi < 1.5 is tested as 3 w > s^2/2.
// The result agrees with the reference arithmetic to ~1e-15 per step
// (tests: tape replay <= 1e-12 on final values).
#pragma once
#include <stdint.h>

#include "fastmath.cuh"

namespace hexo {

// One maturity ("segment") of a price<>() call.  While heading for expiry k the
// reference steps with h = expiry_k/steps (AsianContract.h:35-38); n_steps and w
// come from the host schedule (hexo_gpu_schedule / hexo_gpu_schedule_exact).
struct SegConst {
  double h, w, expiry;
  // Asian trapezoid bookkeeping.  Reference schedule: hcarry = h/2 (the trapezoid of the step
  // that crossed the previous expiry is added with THIS segment's h, HSimulation.tpp:42-44) and
  // hs = 0.  Exact schedule (HEXO_SCHEDULE_EXACT): hcarry = h_{k-1}/2, w = 0, hs = h/2.
  double hcarry, hs;
  double D;         // exp(-kappa h)                                      (:58)
  double m0;        // theta (1 - D):  m = V D + m0                        (:59)
  double c1h, c2h;  // s^2/2 = |V c1h + c2h|                               (:60)
  double K0, K1, K2, K3;  // K4 == K3 because gamma_1 == gamma_2           (:75-79)
  // HEXO_DRIFT_MARTINGALE only: A = K2 + K4/2 and its double, and -K3/2, the factor of V that
  // is left of K1 V once K0* = -ln M - (K1 + K3/2) V takes the place of K0
  double A, A2, K1m;
  // HEXO_CV_GEOMETRIC only: 1 / (sum of the trapezoid weights applied up to this expiry), which
  // normalises the geometric average (the weights sum to T except on the reference grid when the
  // steps land on the expiry, SURVEY finding 6)
  double inv_logw;
  uint32_t n_steps, first_opt, n_strikes, pad;
};

// The step is split in two halves so that the path kernel can software-pipeline
// them: the variance recursion V -> V' is the only long loop-carried dependency;
// the log-spot / exp half of step i hangs off (V_i, V_{i+1}) and can overlap the
// variance half of step i+1.

// The variance half (src/HSimulation.tpp:59-73) comes in two pieces so that the path kernel can keep the straight-line
// part of BOTH halves of its pipelined iteration in one basic block and send the rare cases of
// either half through one shared branch (a convergence barrier in the middle of the block
// stops ptxas from interleaving the two dependency chains).
struct QeVarMid {
  double m, s2h;  // :59, :60 (s^2/2)
  bool rare;      // psi >= 1.5 is POSSIBLE (decided on high words): qe_rare_exact() settles it, and
                  // then the quadratic value has to be replaced by qe_variance_rare
  int hi_sw, hi_dm;  // high words of sqrt(w) and of m - sqrt(w): hi_sw < hi_dm proves psi >= 1.5
  double k0;      // MART only: K0* + (K1 + K3/2) V = -ln M of this step
};

// quadratic branch, evaluated unconditionally (NaN when psi > 2; replaced where mid.rare)
//   MART: also mid.k0 = -ln M = -(A sw / (1 - 2 A dm)) + ln(1 - 2 A dm) / 2   (a b^2 = sw, a = dm)
template <bool MART = false>
__device__ __forceinline__ double qe_variance_quad(const SegConst& g, const double V,
                                                   const double zv, QeVarMid& mid) {
  const double m = fma(V, g.D, g.m0);                       // :59
  const double s2h = fabs(fma(V, g.c1h, g.c2h));            // :60  (s^2/2)
  const double w = fma(m, m, -s2h);                         // m^2 (1 - psi/2)
  const double sw = fast_sqrt(w);                           // a b^2
  const double dm = m - sw;                                 // a
  // a b.  sw dm >= 0 in exact arithmetic; with psi below ~1e-15 (sigma ~1e-8) the rounded sw can
  // exceed m by an ulp and the product is -1e-19: fast_sqrt_signed seeds from |x|, so the result
  // is -3e-10 instead of NaN -- against sw ~ m that is the degenerate a = 0 which the
  // reference's a = m/(1+b^2) rounds to.  The 1e-300 keeps the argument off exact zero.
  const double me = fast_sqrt_signed(fma(sw, dm, kFm.tiny));
  mid.m = m;
  mid.s2h = s2h;
  // :63  psi < 1.5  <=>  3 w > s^2/2  <=>  sqrt(w) > m/2  <=>  sw > dm.  The hot path only compares
  // the HIGH WORDS of sw and dm on the integer pipe (both are positive doubles, so their bit
  // patterns order like their values; a NaN sw -- psi > 2 -- makes dm the same NaN): "sw's high
  // word is larger" proves psi < 1.5; anything else, including equal high words (sw and dm within
  // 2^-20 of each other, one step in a million), goes to the rare block, where qe_rare_exact
  // takes the decision with the exact comparison.  No FP64 instruction for the test per step.
  mid.hi_sw = __double2hiint(sw);
  mid.hi_dm = __double2hiint(dm);
  mid.rare = mid.hi_sw <= mid.hi_dm;
  if (MART) {
    const double d = fma(-g.A2, dm, 1.0);                   // 1 - 2 A a
    const double k0 = fma(0.5, fast_log(d), -(g.A * sw) * fast_rcp(d));
    // M does not exist for 1 - 2 A a <= 0: keep the reference drift K0 + (K1 + K3/2) V
    mid.k0 = d > 0.0 ? k0 : fma(g.K1 - g.K1m, V, g.K0);
  }
  return fma(zv, fma(dm, zv, me + me), sw);                 // :64-68
}

// The exact decision psi >= 1.5  <=>  3 w - s^2/2 <= 0 (read off the sign of the high word; a
// positive denormal counts as zero, where both branches are valid), for code that already sits
// behind the cheap test mid.rare.
__device__ __forceinline__ bool qe_rare_exact(const QeVarMid& mid) {
  if (mid.hi_sw < mid.hi_dm) return true;  // clearly smaller: no FP64 needed (a NaN pair is equal)
  const double w = fma(mid.m, mid.m, -mid.s2h);
  return __double2hiint(fma(3.0, w, -mid.s2h)) <= 0;
}

// Exponential / zero-mass branch (:70-73), a few per cent of the warp-steps, straight-line
// code as well.  With q = m^2 (psi + 1) = m^2 + s^2:
//   p = (psi-1)/(psi+1),  p < U  <=>  s^2 - m^2 < U q
//   beta = 2/(m (psi+1)) = 2 m / q,  1 - p = 2 m^2 / q
//   V' = ln((1-p)/(1-U)) / beta = -q/(2m) ln((1-U) q / (2 m^2))
//   uv : callable returning the variance UNIFORM of the same draw
//   MART: mid.k0 = -ln M,  M = p + beta (1-p) / (beta - A) = ((s^2 - m^2) + 4 m^3 / (2 m - A q)) / q
template <bool MART = false, class UniformFn>
__device__ __forceinline__ double qe_variance_rare(const SegConst& g, const double V, QeVarMid& mid,
                                                   const UniformFn& uv) {
  const double m = mid.m;
  const double m2 = m * m, s2 = mid.s2h + mid.s2h;
  const double q = m2 + s2;
  const double u0 = uv();                                   // :72
  // U = 1.0 (probability 2^-54) would make the reference take log(x/0) = inf; treat it as the
  // largest double below 1 instead, i.e. 1 - U >= 2^-53 -- an integer max on the high word (1 - U
  // is in [0, 1] and exactly 0 only for U = 1), where fmax costs five instructions for its NaN
  // rules.
  const double omu0 = 1.0 - u0;
  const double omu = __hiloint2double(max(__double2hiint(omu0), 0x3ca00000), __double2loint(omu0));
  double v = 0.0;
  // Everything that does not depend on U is formed while the uniform is still on its way (shared
  // memory, 64-bit integer conversion): 1/m, c = q/(2 m^2) = 1/(1-p) and q/(2 m) = 1/beta.
  // Behind U the chain is then (1-U) c -> logarithm -> times 1/beta; the reciprocal of q (1-U) that
  // the literal form ln((1-p)/(1-U)) needs would sit in the middle of it.
  const double rm = fast_rcp(m);
  const double ib = 0.5 * q * rm;                           // 1 / beta
  const double c = ib * rm;                                 // 1 / (1 - p)
  // :73  p < U  <=>  (1-U)/(1-p) < 1, read off the high word of the product the logarithm needs
  // anyway (p >= 0.2 here, and a warp rarely has more than one lane on this path: the test skips
  // the logarithm, the longest dependency chain of the kernel, about as often)
  const double x = omu * c;
  if (__double2hiint(x) < 0x3ff00000) {
    v = -ib * fast_log_pos(x);                              // ln((1-p)/(1-U)) / beta
  }
  if (MART) {
    const double d = fma(-g.A, q, m + m);                   // (beta - A) q
    const double M = ((s2 - m2) + 4.0 * m2 * m * fast_rcp(d)) * fast_rcp(q);
    mid.k0 = d > 0.0 ? -fast_log(M) : fma(g.K1 - g.K1m, V, g.K0);
  }
  return v;
}

// Variance half, src/HSimulation.tpp:59-73.
//   zv : variance normal (used when psi < 1.5)
//   uv : callable returning the variance UNIFORM of the same draw (psi >= 1.5)
//   k0 (MART): receives -ln M of the step
template <bool MART = false, class UniformFn>
__device__ __forceinline__ double qe_variance(const SegConst& g, const double V, const double zv,
                                              const UniformFn& uv, double* k0 = nullptr) {
  QeVarMid mid;
  double Vn = qe_variance_quad<MART>(g, V, zv, mid);
  if (mid.rare) {
    if (qe_rare_exact(mid)) Vn = qe_variance_rare<MART>(g, V, mid, uv);
  }
  if (MART) *k0 = mid.k0;
  return Vn;
}

// Log-spot half, src/HSimulation.tpp:75-80 with K3 == K4: the log-return of the step,
// ln X' - ln X = K0 + K1 V + K2 V' + sqrt(K3 (V + V')) Z_X.
__device__ __forceinline__ double qe_logreturn(const SegConst& g, const double V, const double Vn,
                                               const double zx) {
  // V + V' can be exactly 0 (both steps in the zero-mass branch); the 1e-300 keeps the
  // rsqrt seed finite and changes nothing otherwise
  const double sq = fast_sqrt(fma(g.K3, V + Vn, kFm.tiny));
  return fma(sq, zx, fma(g.K2, Vn, fma(g.K1, V, g.K0)));
}
// HEXO_DRIFT_MARTINGALE: K0* + K1 V = k0 - (K3/2) V with k0 = -ln M of the step
__device__ __forceinline__ double qe_logreturn_mart(const SegConst& g, const double V,
                                                    const double Vn, const double zx,
                                                    const double k0) {
  const double sq = fast_sqrt(fma(g.K3, V + Vn, kFm.tiny));
  return fma(sq, zx, fma(g.K2, Vn, fma(g.K1m, V, k0)));
}

// X' = exp(ln X + delta) (HSimulation.tpp:82) as X e^delta.  One-step log-returns are
// small, so e^delta - 1 is delta times a degree-6 polynomial for |delta| <= 0.08 (interpolant at
// the Chebyshev nodes, error < 7e-16) -- no range reduction, no table look-up -- and the
// table-based fast_exp otherwise (rare: 6 standard deviations of a daily step at 20 % volatility).
__device__ __forceinline__ double grow_spot_poly(const double X, const double delta) {
  double p = fma(delta, kFm.em1[0], kFm.em1[1]);
  p = fma(p, delta, kFm.em1[2]);
  p = fma(p, delta, kFm.em1[3]);
  p = fma(p, delta, kFm.em1[4]);
  p = fma(p, delta, kFm.em1[5]);
  p = fma(p, delta, 1.0);
  return fma(X, p * delta, X);
}
// |delta| > 0.08, tested on the high word (threshold 0.0799999982: the polynomial is as good
// there) on the integer pipe
__device__ __forceinline__ bool grow_spot_is_rare(const double delta) {
  return (__double2hiint(delta) & 0x7fffffff) > 0x3fb47ae0;
}
__device__ __forceinline__ double grow_spot_rare(const double X, const double delta,
                                                 const uint32_t exptab_saddr) {
  return X * fast_exp(delta, exptab_saddr);
}
__device__ __forceinline__ double grow_spot(const double X, const double delta,
                                            const uint32_t exptab_saddr) {
  double Xn = grow_spot_poly(X, delta);
  if (grow_spot_is_rare(delta)) Xn = grow_spot_rare(X, delta, exptab_saddr);
  return Xn;
}

}  // namespace hexo
