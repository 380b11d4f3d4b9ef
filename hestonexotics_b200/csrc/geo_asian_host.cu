// hestonexotics_b200/csrc/geo_asian_host.cu  (host-only code; .cu so it joins the one nvcc build)
//
// the known mean of the geometric-Asian control variate the
// reference suggests at src/inc/HSimulation.h:51.
//
// The path kernel accumulates, next to the arithmetic average A = sum_i w_i X_i / T of
// AAsianCallNonAdaptive (src/inc/AsianContract.h:25-34), the same linear functional of ln X:
//     Y = sum_i w_i ln X_i / W,  W = sum_i w_i,   G = exp(Y)
// (the geometric average on A's own weights w_i, including the reference grid's last-step rule,
// normalised so that G is a mean of the path even where the reference's weights sum to T (1 - 1/n))
// and uses c_j = max(G - K_j, 0) as the control of option j.  Its mean is the price of a
// discretely monitored geometric-Asian call in the Heston model at r = 0, which is available in
// closed form up to one Fourier integral because Y is linear in the log-spot at the grid dates:
//
//  * Y = Omega_0 ln S + sum_m Omega_m D_m with D_m the log-return of step m and
//    Omega_m = sum_{i >= m} w_i / W, so  E exp(z Y)  follows from the affine transform of ONE
//    Heston step,  E[exp(a D + b v') | v] = exp(A(h; a, b) + B(h; a, b) v), applied backwards
//    from the last step with a_m = z Omega_m and b_m = B of the step after it.  With
//    beta = kappa - rho sigma a,  D = sqrt(beta^2 - sigma^2 (a^2 - a)),  B+- = (beta +- D)/sigma^2:
//        y0 = (b - B-)/(b - B+),   y = y0 exp(-D h),
//        B  = (B- - B+ y)/(1 - y),
//        A  = kappa theta (B- h - 2/sigma^2 ln((1 - y)/(1 - y0)))
//    (the Riccati equation B' = sigma^2/2 (B - B+)(B - B-) is linear in (B - B-)/(B - B+));
//  * the call price from the transform by Lewis's formula with F = E[G] = E exp(Y):
//        E max(G - K, 0) = F - sqrt(F K)/pi  int_0^inf Re[ exp(i u ln(F/K)) phi(u - i/2) ] du/(u^2 + 1/4),
//    phi(z) = E exp(i z (Y - ln F)); the integrand is smooth and decays like |phi| / u^2.
//
// The law used is that of the exact Heston process at the grid dates; the simulated process is
// its QE discretisation, so the control's simulated mean differs from this value by the scheme's
// (small) weak error -- the estimator removes that part of the discretisation error which the
// arithmetic and the geometric payoff share.
#include <math.h>

#include <complex>
#include <vector>

#include "../../include/hexo_gpu.h"
#include "host_internal.h"

namespace hexo {
namespace geo {

using cd = std::complex<double>;

// weights of Y on ln X_0 .. ln X_N for maturity k (normalised to sum one), and the width of
// every step; mirrors the path kernel's `integral` bookkeeping (path_kernel.cuh)
static void log_weights(const std::vector<SegConst>& segs, uint32_t k, std::vector<double>& w,
                        std::vector<double>& hstep) {
  uint32_t total = 0;
  for (uint32_t s = 0; s <= k; ++s) total += segs[s].n_steps;
  w.assign(total + 1, 0.0);
  hstep.assign(total + 1, 0.0);  // hstep[m]: width of step m (1-based)
  uint32_t cur = 0, prev = 0;
  for (uint32_t s = 0; s <= k; ++s) {
    const SegConst& g = segs[s];
    const uint32_t n = g.n_steps;
    if (n == 0) continue;
    if (s > 0) {  // trapezoid of the step that crossed the previous expiry
      w[cur] += g.hcarry;
      w[prev] += g.hcarry;
    }
    const uint32_t a = cur, b = cur + n;
    for (uint32_t m = a + 1; m <= b; ++m) hstep[m] = g.h;
    // h/2 (L_a - L_{b-1} + 2 sum_{a < i < b} L_i): all trapezoids but the crossing step's
    w[a] += .5 * g.h;
    w[b - 1] -= .5 * g.h;
    for (uint32_t i = a + 1; i < b; ++i) w[i] += g.h;
    cur = b;
    prev = b - 1;
  }
  const SegConst& g = segs[k];
  w[cur] += g.w + g.hs;  // (L_cur - L_prev) w + hs (L_cur + L_prev)
  w[prev] += g.hs - g.w;
  for (double& x : w) x *= g.inv_logw;  // normalised: the weights of Y sum to one
}

struct Grid {
  std::vector<double> omega;  // Omega_m, m = 0..N
  std::vector<double> h;      // step widths, 1-based
  double lnS, v0, kappa, theta, sigma, rho;
};

// ln E exp(z Y)
static cd log_mgf(const Grid& G, cd z) {
  const double s2 = G.sigma * G.sigma;
  cd b = 0.0, acc = z * G.omega[0] * G.lnS;
  for (size_t m = G.h.size() - 1; m >= 1; --m) {
    const cd a = z * G.omega[m];
    const cd beta = G.kappa - G.rho * G.sigma * a;
    const cd D = std::sqrt(beta * beta - s2 * (a * a - a));
    const cd Bp = (beta + D) / s2, Bm = (beta - D) / s2;
    const cd y0 = (b - Bm) / (b - Bp);
    const cd y = y0 * std::exp(-D * G.h[m]);
    acc += G.kappa * G.theta * (Bm * G.h[m] - 2.0 / s2 * std::log((1.0 - y) / (1.0 - y0)));
    b = (Bm - Bp * y) / (1.0 - y);
  }
  return acc + b * G.v0;
}

// 16-point Gauss-Legendre on [-1, 1]
static const double kGlX[8] = {0.0950125098376374401853193, 0.2816035507792589132304605,
                               0.4580167776572273863424194, 0.6178762444026437484466718,
                               0.7554044083550030338951012, 0.8656312023878317438804679,
                               0.9445750230732325760779884, 0.9894009349916499325961542};
static const double kGlW[8] = {0.1894506104550684962853967, 0.1826034150449235888667637,
                               0.1691565193950025381893121, 0.1495959888165767320815017,
                               0.1246289712555338720524763, 0.0951585116824927848099251,
                               0.0622535239386478928628438, 0.0271524594117540948517806};

}  // namespace geo

// E max(G - K_j, 0) for every option of the request (chain-major like the prices)
int geometric_asian_means(const hexo_price_request* r, const std::vector<SegConst>& segs,
                          std::vector<double>& out) {
  using namespace geo;
  const uint32_t n_opts = r->strike_offsets[r->n_chains];
  out.assign(n_opts, 0.0);
  Grid G;
  G.lnS = log(r->S);
  G.v0 = r->p.v_0;
  G.kappa = r->p.kappa;
  G.theta = r->p.v_m;
  G.sigma = r->p.sigma;
  G.rho = r->p.rho;
  for (uint32_t k = 0; k < r->n_chains; ++k) {
    const uint32_t j0 = r->strike_offsets[k], j1 = r->strike_offsets[k + 1];
    if (j0 == j1) continue;
    std::vector<double> w;
    log_weights(segs, k, w, G.h);
    G.omega.assign(w.size(), 0.0);
    double run = 0.0;
    for (size_t i = w.size(); i-- > 0;) {
      run += w[i];
      G.omega[i] = run;
    }
    const double lnF = log_mgf(G, cd(1.0, 0.0)).real();
    const double F = exp(lnF);
    std::vector<double> integral(j1 - j0, 0.0), lnFK(j1 - j0);
    for (uint32_t j = j0; j < j1; ++j) {
      if (!(r->strikes[j] > 0.0)) lnFK[j - j0] = 0.0;  // K <= 0: E[G - K] = F - K, no integral
      else lnFK[j - j0] = lnF - log(r->strikes[j]);
    }
    // panels [0,1/2], [1/2,1], [1,2], [2,4], ... until a panel no longer contributes; a panel is
    // cut into pieces short enough for exp(i u ln(F/K)) of the farthest strike (16 nodes resolve
    // about 8 radians comfortably)
    double osc = 0.0;
    for (double x : lnFK) osc = std::max(osc, fabs(x));
    double lo = 0.0, hi = 0.5;
    for (int panel = 0; panel < 40; ++panel) {
      double biggest = 0.0;
      const int pieces = 1 + (int)std::min(4096.0, (hi - lo) * osc / 8.0);
      const double half = .5 * (hi - lo) / pieces;
      for (int piece = 0; piece < pieces; ++piece) {
        const double mid = lo + (2 * piece + 1) * half;
        for (int q = 0; q < 16; ++q) {
          const double x = q < 8 ? -kGlX[7 - q] : kGlX[q - 8];
          const double wq = (q < 8 ? kGlW[7 - q] : kGlW[q - 8]) * half;
          const double u = mid + half * x;
          // phi(u - i/2) = E exp(i (u - i/2)(Y - ln F)) = exp(log_mgf(1/2 + i u) - (1/2 + i u) ln F)
          const cd z(0.5, u);
          const cd phi = std::exp(log_mgf(G, z) - z * lnF);
          const double den = u * u + 0.25;
          for (uint32_t j = 0; j < j1 - j0; ++j) {
            const double ang = u * lnFK[j];
            const double term = (phi.real() * cos(ang) - phi.imag() * sin(ang)) / den * wq;
            integral[j] += term;
            biggest = std::max(biggest, fabs(term));
          }
        }
      }
      if (panel > 4 && biggest < 1e-17) break;
      lo = hi;
      hi = hi < 1.0 ? 1.0 : 2.0 * hi;
    }
    for (uint32_t j = j0; j < j1; ++j) {
      const double K = r->strikes[j];
      // (a call price: quadrature noise of 1e-9 must not make a far out-of-the-money one negative)
      out[j] = K > 0.0 ? std::max(0.0, F - sqrt(F * K) / M_PI * integral[j - j0]) : F - K;
      if (!std::isfinite(out[j])) return HEXO_ERR_INVALID_ARGUMENT;  // degenerate transform
    }
  }
  return HEXO_OK;
}

}  // namespace hexo
