// hestonexotics_b200/csrc/qe.cuh
//
// Andersen quadratic-exponential (QE) step of the Heston SDE, psi_c = 1.5,
// gamma_1 = gamma_2 = 1/2, no drift and no martingale correction -- the
// discretisation of HQEAnderson::operator++ (src/HSimulation.tpp:52-86,
// PSI_C src/inc/HSimulation.h:12).  Everything that depends only on
// (HParams, step width) is hoisted into SegConst on the host; the reference
// recomputes it every step (:58,:75-79).
#pragma once
#include <stdint.h>

#include "ppnd16.cuh"

namespace hexo {

// One maturity ("segment") of a price<>() call.  While heading for expiry k the
// reference steps with h = expiry_k/steps (AsianContract.h:35-38); n_steps and w
// come from the host schedule (hexo_gpu_schedule).
struct SegConst {
  double h, w, expiry;
  double D;       // exp(-kappa h)                                  (:58)
  double c1, c2;  // s^2 = |V c1 + c2|                               (:60)
  double K0, K1, K2, K3, K4;  //                                     (:75-79)
  uint32_t n_steps, first_opt, n_strikes, pad;
};

// Draws for one step taken from two raw 64-bit words of the stream: the first
// word is the variance draw (normal when psi < psi_c, else the uniform of the
// same word), the second the log-spot normal (:67,:72,:80).
template <int NORMAL_MODE>
struct WordDraws {
  uint64_t wv, wx;
  __device__ __forceinline__ double variance_normal() const {
    return ppnd16<NORMAL_MODE>(u64_to_unit(wv));
  }
  __device__ __forceinline__ double variance_uniform() const { return u64_to_unit(wv); }
  __device__ __forceinline__ double spot_normal() const {
    return ppnd16<NORMAL_MODE>(u64_to_unit(wx));
  }
};

// Draws read from a tape (K4 replay kernel)
struct TapeDraws {
  double zv, uv, zx;
  __device__ __forceinline__ double variance_normal() const { return zv; }
  __device__ __forceinline__ double variance_uniform() const { return uv; }
  __device__ __forceinline__ double spot_normal() const { return zx; }
};

// (V, ln X) -> next step.  src/HSimulation.tpp:59-80.
template <class Draws>
__device__ __forceinline__ void qe_step(const SegConst& g, const double theta, double& V,
                                        double& lnX, const Draws& d) {
  const double m = theta + (V - theta) * g.D;              // :59
  const double sp2 = fabs(V * g.c1 + g.c2);                // :60
  const double psi = sp2 / (m * m);                        // :61
  double Vn;
  if (psi < 1.5) {                                         // :63
    const double ip = 2.0 / psi;
    const double bp2 = ip - 1.0 + sqrt(ip * (ip - 1.0));   // :64
    const double b = sqrt(bp2);                            // :65
    const double a = m / (1.0 + bp2);                      // :66
    const double bz = b + d.variance_normal();             // :67
    Vn = a * bz * bz;                                      // :68
  } else {
    const double p = (psi - 1.0) / (psi + 1.0);            // :70
    const double beta = 2.0 / (m * (psi + 1.0));           // :71
    // U = 1.0 (probability 2^-54) would make the reference take log(x/0) = inf;
    // clamp to the largest double below 1 instead.
    const double u = fmin(d.variance_uniform(), 0.99999999999999988898);  // :72
    Vn = p < u ? log((1.0 - p) / (1.0 - u)) / beta : 0.0;  // :73
  }
  lnX = lnX + g.K0 + g.K1 * V + g.K2 * Vn + sqrt(g.K3 * V + g.K4 * Vn) * d.spot_normal();  // :80
  V = Vn;
}

}  // namespace hexo
