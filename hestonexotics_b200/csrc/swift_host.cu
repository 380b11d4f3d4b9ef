// hestonexotics_b200/csrc/swift_host.cu  (host-only code; .cu so it joins the one nvcc build)
//
// the semi-analytic European pricer that feeds and
// benchmarks the Monte-Carlo path -- the Heston characteristic function with its
// analytic parameter gradient and the SWIFT (Shannon wavelet inverse Fourier)
// pricer, without FFTW or Eigen.  Restates, in our own code:
//   HDistribution::chf / chf_chf_grad / cumulants / int_error  src/HDistribution.cpp:8-113
//     (same functions, own derivation: decaying-exponential form of the Riccati solution with
//     one directional-derivative routine; cumulants from the moment cascade of the model)
//   SwiftParameters                                             src/SWIFT.cpp:18-35
//   SWIFT::SWIFT (density coefficients via two Fourier sums)    src/SWIFT.cpp:37-79
//   SWIFT::CacheEntry (prices and gradients of one chain)       src/SWIFT.cpp:102-118
// Pinned by the reference's own known-answer tests (40 prices and 200 partials to a
// summed error < 1e-9, src/UnitTest.cpp:219-497) in tests/test_swift.py.
// J <= 512, so the two transforms are evaluated as direct sums with an exact
// twiddle table (~10^5 complex multiplies) instead of pulling in an FFT library.
#include <math.h>

#include <complex>
#include <vector>

#include "../../include/hexo_gpu.h"

namespace hexo {
namespace swift {

using cd = std::complex<double>;
static const cd I(0.0, 1.0);

// ---------------------------------------------------------------------------------------------
// Characteristic function and parameter gradient.
//
// The reference evaluates chf(u) = E[exp(-i u x_tau)], x the log-return at r = 0, in the
// sinh / cosh arrangement of Cui et al. with a hand-expanded derivative per parameter
// (src/HDistribution.cpp:12-88).  Same function here, derived from the Riccati solution in
// its decaying-exponential form.  With s = -u,
//     f = s^2 + i s,   beta = kappa - rho sigma i s,   d = sqrt(beta^2 + sigma^2 f)  (Re d >= 0),
//     e = exp(-d tau),  c = beta - d,  N = d (1 + e) + beta (1 - e),
// and because (beta - d)(beta + d) = -sigma^2 f the two pieces of the exponent collapse to
//     ln chf = kappa theta / sigma^2 * G  -  v_0 f H,
//     G = c tau - 2 ln(N / (2 d)),     H = (1 - e) / N.
// Only e = exp(-d tau) with Re d >= 0 appears, so nothing overflows for large |u| tau, and
// N / (2 d) -> 1 as u -> 0, so the principal logarithm is the continuous branch.
//
// Gradient: v_0 and theta enter linearly.  rho, kappa and sigma enter through beta and d only,
// so ONE routine differentiates (G, H) along a direction (d beta, d(sigma^2 f)/2):
//     d' = (beta beta' + [sigma] sigma f) / d,          e' = -tau e d',
//     N' = d' (1 + e) + beta' (1 - e) + c tau e d',
//     G' = (beta' - d') tau - 2 (N'/N - d'/d),          H' = (tau e d' N - (1 - e) N') / N^2.
// No division by u anywhere (the reference's B_kappa has one), so u = 0 is a regular point.
// ---------------------------------------------------------------------------------------------
struct LogChf {
  cd f, beta, d, e, c, N, G, H;
};

static LogChf log_chf_parts(const hexo_hparams& p, cd u, double tau) {
  LogChf t;
  const cd s = -u;
  t.f = s * s + I * s;
  t.beta = p.kappa - p.rho * p.sigma * I * s;
  t.d = std::sqrt(t.beta * t.beta + p.sigma * p.sigma * t.f);
  t.e = std::exp(-t.d * tau);
  t.c = t.beta - t.d;
  t.N = t.d * (1. + t.e) + t.beta * (1. - t.e);
  t.G = t.c * tau - 2. * std::log(t.N / (2. * t.d));
  t.H = (1. - t.e) / t.N;
  return t;
}

static cd log_chf(const hexo_hparams& p, const LogChf& t) {
  return p.kappa * p.v_m / (p.sigma * p.sigma) * t.G - p.v_0 * t.f * t.H;
}

static cd chf_value(const hexo_hparams& p, cd u, double tau) {
  return std::exp(log_chf(p, log_chf_parts(p, u, tau)));
}

// chf and its partial derivatives in HParams order (v_0, v_m, rho, kappa, sigma)
static void chf_and_grad(const hexo_hparams& p, cd u, double tau, cd out[6]) {
  const LogChf t = log_chf_parts(p, u, tau);
  const cd s = -u;
  const double s2 = p.sigma * p.sigma, w = p.kappa * p.v_m / s2;
  // derivative of ln chf along (beta', half of d(sigma^2 f)) -- without the explicit
  // dependence of the prefactor kappa theta / sigma^2 on the parameter
  auto along = [&](cd dbeta, cd half_ds2f) {
    const cd dd = (t.beta * dbeta + half_ds2f) / t.d;
    const cd dN = dd * (1. + t.e) + dbeta * (1. - t.e) + t.c * tau * t.e * dd;
    const cd dG = (dbeta - dd) * tau - 2. * (dN / t.N - dd / t.d);
    const cd dH = (tau * t.e * dd * t.N - (1. - t.e) * dN) / (t.N * t.N);
    return w * dG - p.v_0 * t.f * dH;
  };
  const cd phi = std::exp(log_chf(p, t));
  out[0] = phi;
  out[1] = phi * (-t.f * t.H);                                              // v_0
  out[2] = phi * (p.kappa / s2 * t.G);                                      // v_m
  out[3] = phi * along(-p.sigma * I * s, 0.);                               // rho
  out[4] = phi * (p.v_m / s2 * t.G + along(1., 0.));                        // kappa
  out[5] = phi * (-2. * w / p.sigma * t.G + along(-p.rho * I * s, p.sigma * t.f));  // sigma
}

// SWIFT truncation error bound for wavelet scale m (HDistribution::int_error,
// src/HDistribution.cpp:8-11): |chf(2^m pi) + chf(-2^m pi)| / (4 2^m pi^2 tau).  For real u the
// two values are complex conjugates, so the numerator is 2 |Re chf(2^m pi)|.
static double truncation_error(const hexo_hparams& p, double tau, unsigned m) {
  const double e = std::exp2((double)m);
  return std::fabs(chf_value(p, e * M_PI, tau).real()) / (2 * e * M_PI * M_PI * tau);
}

// ---------------------------------------------------------------------------------------------
// Cumulants of the log-return for the integration range (HDistribution::first/second/
// fourth_order_moment, src/HDistribution.cpp:90-113, which print closed forms of the
// stationary-start case v_0 = v_m).  Here they come out of the model instead of a formula sheet:
// E[exp(w x)] = exp(A(t, w) + B(t, w) v_0) with  B' = sigma^2/2 B^2 + (rho sigma w - kappa) B
// + (w^2 - w)/2,  A' = kappa theta B.  Expanding B = sum b_n w^n turns the Riccati equation into
// the linear cascade
//     b_n' = -kappa b_n + rho sigma b_{n-1} + sigma^2/2 sum_{j=1}^{n-1} b_j b_{n-j} + ([n=2]-[n=1])/2,
// whose solutions are finite sums of t^k exp(-j kappa t).  ExpPoly holds such a sum exactly and
// implements the three operations the cascade needs; cumulant n = n! (a_n + b_n v_m).
// ---------------------------------------------------------------------------------------------
struct ExpPoly {
  static constexpr int J = 6, K = 6;  // exponents j kappa (j < J), powers t^k (k < K)
  double c[J][K] = {};
  double kappa = 0.0;
  explicit ExpPoly(double kap) : kappa(kap) {}
  double eval(double t) const {
    double s = 0.0;
    for (int j = 0; j < J; ++j) {
      double poly = 0.0;
      for (int k = K - 1; k >= 0; --k) poly = poly * t + c[j][k];
      s += poly * std::exp(-j * kappa * t);
    }
    return s;
  }
  void axpy(double a, const ExpPoly& x) {
    for (int j = 0; j < J; ++j)
      for (int k = 0; k < K; ++k) c[j][k] += a * x.c[j][k];
  }
  ExpPoly times(const ExpPoly& x) const {
    ExpPoly r(kappa);
    for (int j = 0; j < J; ++j)
      for (int k = 0; k < K; ++k)
        if (c[j][k] != 0.0)
          for (int j2 = 0; j + j2 < J; ++j2)
            for (int k2 = 0; k + k2 < K; ++k2) r.c[j + j2][k + k2] += c[j][k] * x.c[j2][k2];
    return r;
  }
  // exp(-lambda t) int_0^t exp(lambda s) g(s) ds with lambda = shift kappa: shift = 1 solves
  // y' = -kappa y + g, y(0) = 0; shift = 0 is the plain integral
  ExpPoly integrate(int shift) const {
    ExpPoly r(kappa);
    for (int j = 0; j < J; ++j)
      for (int k = 0; k < K; ++k) {
        const double g = c[j][k];
        if (g == 0.0) continue;
        const double alpha = (shift - j) * kappa;  // int s^k exp(alpha s) ds
        if (shift == j) {
          r.c[shift][k + 1] += g / (k + 1);
          continue;
        }
        // exp(alpha s) sum_i (-1)^i k!/(k-i)! s^(k-i) / alpha^(i+1), between 0 and t
        double fac = 1.0, pw = alpha;
        for (int i = 0; i <= k; ++i) {
          const double term = ((i & 1) ? -g : g) * fac / pw;
          r.c[j][k - i] += term;               // upper limit: exp(alpha t) exp(-shift kappa t)
          if (i == k) r.c[shift][0] -= term;   // lower limit: only the s^0 term survives at s = 0
          fac *= (k - i);
          pw *= alpha;
        }
      }
    return r;
  }
};

// cumulants 1, 2 and 4 of the log-return at time tau for v_0 = v_m
static void cumulants_124(const hexo_hparams& p, double tau, double out[3]) {
  const double kap = p.kappa, half_s2 = .5 * p.sigma * p.sigma, rs = p.rho * p.sigma;
  std::vector<ExpPoly> b(5, ExpPoly(kap));
  double cum[5] = {};
  double fact = 1.0;
  for (int n = 1; n <= 4; ++n) {
    ExpPoly g(kap);  // right-hand side of b_n' + kappa b_n
    if (n == 1) g.c[0][0] = -.5;
    if (n == 2) g.c[0][0] = .5;
    if (n > 1) g.axpy(rs, b[n - 1]);
    for (int j = 1; j < n; ++j) g.axpy(half_s2, b[j].times(b[n - j]));
    b[n] = g.integrate(1);
    ExpPoly a = b[n].integrate(0);  // a_n / (kappa theta)
    fact *= n;
    cum[n] = fact * (kap * p.v_m * a.eval(tau) + p.v_m * b[n].eval(tau));
  }
  out[0] = cum[1];
  out[1] = cum[2];
  out[2] = cum[4];
}

// frequency of wavelet coefficient i, src/SWIFT.cpp:18-20
static double freq(const hexo_swift_params& q, unsigned i) {
  return M_PI * (2 * (double)i + 1) / (2. * (double)q.J) * (double)q.exp2_m;
}

// Density coefficients, src/SWIFT.cpp:37-79.  The reference runs a backward complex FFT of
// size 2J over the J payoff terms and a real-to-complex FFT of size 4J over at most
// k_2-k_1+1 non-zero samples; both are evaluated here as direct sums over the non-zero terms.
static std::vector<cd> density_coeffs(const hexo_swift_params& q) {
  const unsigned J = q.J, N2 = 2 * J, N4 = 4 * J;
  std::vector<cd> tw(N4);  // tw[t] = exp(+2 pi i t / 4J)
  for (unsigned t = 0; t < N4; ++t) {
    const double ang = 2.0 * M_PI * (double)t / (double)N4;
    tw[t] = cd(std::cos(ang), std::sin(ang));
  }
  const double lo = std::max(q.lower, 0.0);
  const double exp_upper = std::exp(q.upper), exp_lower = std::exp(lo);
  auto H = [&](double y, double exp_y, unsigned j) {
    const double uj = freq(q, j);
    return -I * std::exp(-I * uj * y) * (1. / uj - exp_y / (I + uj));
  };
  std::vector<cd> payoff(J);
  for (unsigned j = 0; j < J; ++j) payoff[j] = H(q.upper, exp_upper, j) - H(lo, exp_lower, j);
  const double scale = q.sqrt_exp2_m / (double)J;
  // density_in[n] for n = (4J + i) mod 4J, i in [k_1, k_2]
  std::vector<double> din(N4, 0.0);
  std::vector<unsigned> nz;
  for (int i = q.k_1; i <= q.k_2; ++i) {
    const unsigned k = (unsigned)((int)N2 + i) & (N2 - 1);  // J is a power of two
    const unsigned n = (unsigned)((int)N4 + i) & (N4 - 1);
    cd acc(0.0, 0.0);  // backward DFT bin k of size 2J: sum_j payoff[j] e^{+2 pi i j k / 2J}
    for (unsigned j = 0; j < J; ++j) acc += payoff[j] * tw[(2u * j * k) & (N4 - 1)];
    const double ang = (double)i * M_PI / (double)N2;
    din[n] = (cd(std::cos(ang), std::sin(ang)) * acc).real() * scale;
    nz.push_back(n);
  }
  std::vector<cd> coeffs(J);
  for (unsigned j = 0; j < J; ++j) {
    const unsigned f = 2 * j + 1;
    cd acc(0.0, 0.0);  // forward DFT bin f of size 4J: sum_n din[n] e^{-2 pi i f n / 4J}
    for (unsigned n : nz) acc += din[n] * std::conj(tw[(f * n) & (N4 - 1)]);
    coeffs[j] = std::conj(acc) * scale;
  }
  return coeffs;
}

}  // namespace swift
}  // namespace hexo

using namespace hexo::swift;

extern "C" {

int hexo_heston_chf(const hexo_hparams* p, double tau, double u_re, double u_im, double out[12]) {
  if (!p || !out || !(tau > 0)) return HEXO_ERR_INVALID_ARGUMENT;
  cd v[6];
  chf_and_grad(*p, cd(u_re, u_im), tau, v);
  for (int j = 0; j < 6; ++j) {
    out[2 * j] = v[j].real();
    out[2 * j + 1] = v[j].imag();
  }
  return HEXO_OK;
}

int hexo_heston_cumulants(const hexo_hparams* p, double tau, double out[3]) {
  if (!p || !out || !(tau > 0) || !(p->kappa > 0)) return HEXO_ERR_INVALID_ARGUMENT;
  cumulants_124(*p, tau, out);
  return HEXO_OK;
}

// SwiftParameters(distr, S, chain), src/SWIFT.cpp:21-35.  truncation_precision: the reference
// uses 1e-7 in its release build and 1e-10 in debug (SWIFT.cpp:12-16); <= 0 selects 1e-7.
int hexo_swift_default_params(const hexo_hparams* p, double tau, double risk_free, double S,
                              double min_strike, double max_strike, double truncation_precision,
                              hexo_swift_params* out) {
  if (!p || !out || !(tau > 0) || !(S > 0) || !(min_strike > 0) || !(max_strike >= min_strike))
    return HEXO_ERR_INVALID_ARGUMENT;
  const double prec = truncation_precision > 0 ? truncation_precision : 1e-7;
  unsigned m = 0;
  while (truncation_error(*p, tau, m++) > prec && m < 30) {
  }
  const double hi = risk_free * tau + std::log(S / min_strike);
  const double lo = risk_free * tau + std::log(S / max_strike);
  double cum[3];
  cumulants_124(*p, tau, cum);
  const double c = std::fabs(cum[0]) + 10. * std::sqrt(std::fabs(cum[1]) + std::sqrt(std::fabs(cum[2])));
  out->m = m;
  out->exp2_m = (uint32_t)std::exp2((double)m);
  out->sqrt_exp2_m = std::sqrt((double)out->exp2_m);
  out->lower = lo - c;
  out->upper = hi + c;
  out->k_1 = (int32_t)std::ceil(out->exp2_m * out->lower);
  out->k_2 = (int32_t)std::floor(out->exp2_m * out->upper);
  const double iota = std::ceil(std::log2(M_PI * std::abs((double)out->k_1 - (double)out->k_2))) - 1;
  out->J = (uint32_t)std::exp2(iota - 1);
  return HEXO_OK;
}

// SWIFT prices (and gradients) of one option chain: SWIFT::price_opts / price_opts_grad,
// src/SWIFT.cpp:87-118.  grad_out (or NULL) is [n_strikes][5] in HParams order.
int hexo_swift_price_chain(const hexo_swift_params* q, const hexo_hparams* p, double tau,
                           double risk_free, double S, const double* strikes, uint32_t n_strikes,
                           double* prices_out, double* grad_out) {
  if (!q || !p || !strikes || !prices_out || n_strikes == 0 || !(tau > 0) || q->J == 0 ||
      (q->J & (q->J - 1)) != 0 || q->J > (1u << 16))
    return HEXO_ERR_INVALID_ARGUMENT;
  const std::vector<cd> dc = density_coeffs(*q);
  const unsigned J = q->J;
  std::vector<cd> weights(6 * (size_t)J);  // chf / gradient at u_i times the density coefficient
  for (unsigned i = 0; i < J; ++i) {
    cd v[6];
    chf_and_grad(*p, freq(*q, i), tau, v);
    for (int j = 0; j < 6; ++j) weights[(size_t)j * J + i] = v[j] * dc[i];
  }
  const double discount = std::exp(-risk_free * tau);
  for (uint32_t s = 0; s < n_strikes; ++s) {
    const double K = strikes[s];
    const double x = risk_free * tau + std::log(S / K);
    cd acc[6];
    for (unsigned i = 0; i < J; ++i) {
      const cd e = discount * K * std::exp(-I * freq(*q, i) * x);
      for (int j = 0; j < (grad_out ? 6 : 1); ++j) acc[j] += weights[(size_t)j * J + i] * e;
    }
    prices_out[s] = acc[0].real();
    if (grad_out)
      for (int j = 1; j < 6; ++j) grad_out[5 * (size_t)s + (j - 1)] = acc[j].real();
  }
  return HEXO_OK;
}

}  // extern "C"
