// hestonexotics_b200/csrc/swift_host.cu  (host-only code; .cu so it joins the one nvcc build)
//
// the semi-analytic European pricer that feeds and
// benchmarks the Monte-Carlo path -- the Heston characteristic function with its
// analytic parameter gradient and the SWIFT (Shannon wavelet inverse Fourier)
// pricer, without FFTW or Eigen.  Restates, in our own code:
//   HDistribution::chf / chf_chf_grad / cumulants / int_error  src/HDistribution.cpp:8-113
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

// intermediate terms of the characteristic function (Cui et al. form,
// arXiv:2103.01570), src/HDistribution.cpp:12-27
struct ChfTerms {
  cd xi, d, sinh_v, cosh_v, A1, A2, A, B, D;
  double exp_kappa_tau;
};

static ChfTerms chf_terms(const hexo_hparams& p, cd u, double tau) {
  ChfTerms t;
  t.xi = p.kappa + p.sigma * p.rho * I * u;
  const cd fac = u * u - I * u;
  t.d = std::sqrt(t.xi * t.xi + p.sigma * p.sigma * fac);
  t.sinh_v = std::sinh(t.d * tau * .5);
  t.cosh_v = std::cosh(t.d * tau * .5);
  t.A1 = fac * t.sinh_v;
  t.A2 = t.d / p.v_0 * t.cosh_v + t.xi / p.v_0 * t.sinh_v;
  t.A = t.A1 / t.A2;
  t.exp_kappa_tau = std::exp(p.kappa * tau * .5);
  t.B = t.d * t.exp_kappa_tau / (p.v_0 * t.A2);
  t.D = std::log((2. * t.d) / (t.d + t.xi + (t.d - t.xi) * std::exp(-t.d * tau))) +
        (p.kappa - t.d) * tau * .5;
  return t;
}

// src/HDistribution.cpp:32-38
static cd chf_value(const hexo_hparams& p, cd u, double tau, const ChfTerms& t) {
  return std::exp((p.kappa * p.v_m * p.rho * tau * u * I) / p.sigma - t.A +
                  (2. * t.D * p.kappa * p.v_m) / (p.sigma * p.sigma));
}

// chf and its partial derivatives in HParams order (v_0, v_m, rho, kappa, sigma),
// src/HDistribution.cpp:57-88
static void chf_and_grad(const hexo_hparams& p, cd u, double tau, cd out[6]) {
  const ChfTerms t = chf_terms(p, u, tau);
  const cd phi = chf_value(p, u, tau, t);
  const cd iu = I * u, fac = u * u - iu;
  const cd d_rho = t.xi * p.sigma * iu / t.d;
  const cd A2_rho = p.sigma * iu * (2. + t.xi * tau) / (2. * t.d * p.v_0) *
                    (t.xi * t.cosh_v + t.d * t.sinh_v);
  const cd B_rho = t.exp_kappa_tau / p.v_0 * (d_rho / t.A2 - t.d / (t.A2 * t.A2) * A2_rho);
  const cd A1_rho = (iu * fac * tau * t.xi * p.sigma) / (2. * t.d) * t.cosh_v;
  const cd A_rho = A1_rho / t.A2 - t.A / t.A2 * A2_rho;
  const cd B_kappa = -I / (p.sigma * u) * B_rho + t.B * tau * .5;
  const cd d_sigma = (p.rho / p.sigma - 1. / t.xi) * d_rho + p.sigma * u * u / t.d;
  const cd A1_sigma = fac * .5 * tau * d_sigma * t.cosh_v;
  const cd A2_sigma = p.rho / p.sigma * A2_rho +
                      (2. + tau * t.xi) / (p.v_0 * tau * t.xi * iu) * A1_rho +
                      p.sigma * tau * t.A1 / p.v_0 * .5;
  const cd A_sigma = A1_sigma / t.A2 - t.A / t.A2 * A2_sigma;
  const cd tiu_vm_s = p.v_m * tau * iu / p.sigma;
  const double s2 = p.sigma * p.sigma;
  const double kvm2_s2 = 2. * p.kappa * p.v_m / s2;
  const cd h_v0 = -t.A / p.v_0;
  const cd h_vm = 2. * p.kappa / s2 * t.D + p.kappa * p.rho * tau * iu / p.sigma;
  const cd h_sigma = -A_sigma - 2. * kvm2_s2 / p.sigma * t.D +
                     kvm2_s2 / t.d * (d_sigma - t.d / t.A2 * A2_sigma) -
                     tiu_vm_s / p.sigma * p.rho * p.kappa;
  const cd h_kappa = -A_rho / (p.sigma * iu) + 2. * p.v_m / s2 * t.D + kvm2_s2 / t.B * B_kappa +
                     tiu_vm_s * p.rho;
  const cd h_rho = -A_rho + kvm2_s2 / t.d * (d_rho - t.d / t.A2 * A2_rho) + tiu_vm_s * p.kappa;
  out[0] = phi;
  out[1] = phi * h_v0;
  out[2] = phi * h_vm;
  out[3] = phi * h_rho;
  out[4] = phi * h_kappa;
  out[5] = phi * h_sigma;
}

// SWIFT truncation error bound for wavelet scale m, src/HDistribution.cpp:8-11
static double truncation_error(const hexo_hparams& p, double tau, unsigned m) {
  const double e = std::exp2((double)m);
  const cd a = chf_value(p, e * M_PI, tau, chf_terms(p, e * M_PI, tau));
  const cd b = chf_value(p, -e * M_PI, tau, chf_terms(p, -e * M_PI, tau));
  return std::abs(a + b) / (4 * e * M_PI * M_PI * tau);
}

// cumulants of the log-return, src/HDistribution.cpp:90-113
static double cumulant1(const hexo_hparams& p, double tau) { return -.5 * p.v_m * tau; }
static double cumulant2(const hexo_hparams& p, double t) {
  const double s2 = p.v_m, r = p.rho, a = p.kappa, k = p.sigma;
  const double a2 = a * a, a3 = a2 * a, k2 = k * k;
  return s2 / (8 * a3) *
         (-k2 * std::exp(-2 * a * t) + 4 * k * std::exp(-a * t) * (k - 2 * a * r) +
          2 * a * t * (4 * a2 + k2 - 4 * a * k * r) + k * (8 * a * r - 3 * k));
}
static double cumulant4(const hexo_hparams& p, double t) {
  const double s2 = p.v_m, r = p.rho, a = p.kappa, k = p.sigma;
  const double a2 = a * a, a3 = a2 * a, a4 = a3 * a;
  const double k2 = k * k, k3 = k2 * k, k4 = k3 * k;
  const double t2 = t * t, r2 = r * r;
  return (3 * k2 * s2) / (64 * std::pow(a, 7)) *
         (-3 * k4 * std::exp(-4 * a * t) -
          8 * k2 * std::exp(-3 * a * t) *
              (2 * a * k * t * (k - 2 * a * r) + 4 * a2 + k2 - 6 * a * k * r) -
          4 * std::exp(-2 * a * t) *
              (4 * a2 * k2 * t2 * std::pow(k - 2 * a * r, 2) +
               2 * a * k * t * (k3 - 16 * a3 * r - 12 * a * k2 * r + 4 * a2 * k * (3 + 4 * r2)) +
               8 * a4 - 3 * k4 - 32 * a3 * k * r + 8 * a * k3 * r + 16 * a2 * k2 * r2) -
          8 * std::exp(-a * t) *
              (-2 * a2 * k * t2 * std::pow(k - 2 * a * r, 3) -
               8 * a * t *
                   (k4 - 7 * a * k3 * r + 4 * a4 * r2 - 8 * a3 * k * r * (1 + r2) +
                    a2 * k2 * (3 + 14 * r2)) -
               9 * k4 + 70 * a * k3 * r + 32 * a3 * k * r * (4 + 3 * r2) - 16 * a4 * (1 + 4 * r2) -
               4 * a2 * k2 * (9 + 40 * r2)) +
          4 * a * t *
              (5 * k4 - 40 * a * k3 * r - 32 * a3 * k * r * (3 + 2 * r2) + 16 * a4 * (1 + 4 * r2) +
               24 * a2 * k2 * (1 + 4 * r2)) -
          73 * k4 + 544 * a * k3 * r + 128 * a3 * k * r * (7 + 6 * r2) - 32 * a4 * (3 + 16 * r2) -
          64 * a2 * k2 * (4 + 19 * r2));
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
  const double c = std::abs(cumulant1(*p, tau)) +
                   10. * std::sqrt(std::fabs(cumulant2(*p, tau)) + std::sqrt(std::abs(cumulant4(*p, tau))));
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
