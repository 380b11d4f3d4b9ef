// hestonexotics_b200/cpp/hexo_gpu_adapter.hpp
//
// Header-only C++ adapter that puts the reference's own call signature back on
// top of the C ABI (include/hexo_gpu.h), so that the reference's host code can
// switch its Monte-Carlo pricing to the GPU by changing one identifier:
//
//   src/Main.cpp:88
//   - price<HSimulation::HQEAnderson<ffloat,AAsianCallNonAdaptive>>(p, S, chains, 1e+5, n, 1e+3);
//   + HSimulation::price_gpu<HSimulation::HQEAnderson<ffloat,AAsianCallNonAdaptive>>(p, S, chains, 1e+5, n, 1e+3);
//
// It is compiled against the reference's unchanged headers (Types.h,
// HDistribution.h, HSimulation.h, AsianContract.h, VanillaContract.h) and links
// libhexo_gpu.so.  Argument meaning, result order (chain-major, then option
// order, HSimulation.tpp:39-40) and the exception-on-error behaviour
// (std::runtime_error, like AsianContract.h:30) follow the reference.
#pragma once

#include <list>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "AsianContract.h"    // reference header
#include "HSimulation.h"      // reference header
#include "VanillaContract.h"  // reference header
#include "hexo_gpu.h"

namespace HSimulation {

namespace gpu_detail {
template <class Scheme>
struct payoff_of;
template <>
struct payoff_of<HQEAnderson<ffloat, AAsianCallNonAdaptive>> {
  static constexpr int value = HEXO_PAYOFF_ASIAN;
};
template <>
struct payoff_of<HQEAnderson<ffloat, EuropeanCallNonAdaptive>> {
  static constexpr int value = HEXO_PAYOFF_EUROPEAN;
};
}  // namespace gpu_detail

struct GpuPriceOptions {
  uint64_t seed = 1;                      // the reference's thread-0 seed (HSimulation.tpp:28)
  int normal_mode = HEXO_NORMAL_F32;      // the reference as built (as241.f90:20-25); optional:
                                          // HEXO_NORMAL_F64, HEXO_NORMAL_F32_PPND7
  int rng_mode = HEXO_RNG_SHISHUA;        // the reference's generator; HEXO_RNG_PHILOX optional
  int schedule_mode = HEXO_SCHEDULE_REFERENCE;  // the reference's time grid, quirks included
  int control_variate = HEXO_CV_NONE;     // HEXO_CV_GEOMETRIC (Asian): the control HSimulation.h:51
                                          // suggests; HEXO_CV_UNDERLYING: c = final value - S
  int drift_mode = HEXO_DRIFT_REFERENCE;  // HEXO_DRIFT_MARTINGALE: Andersen's K0* per step
  uint64_t n_streams = 0;                 // 0 = sized for the device(s)
  int n_gpus = 1;                         // devices of this process to spread over; 0 = all
  std::vector<ffloat>* stderr_out = nullptr;  // optional Monte-Carlo standard errors
};

/**
 * Drop-in for HSimulation::price<Scheme> (src/inc/HSimulation.h:61-63).
 * Same six arguments; `n_opts` must equal the number of options in all_chains
 * (the reference computes it that way, src/Main.cpp:53-57).
 */
template <class Scheme>
std::vector<ffloat> price_gpu(const HParams& p, const ffloat S,
                              const std::list<options_chain>& all_chains,
                              unsigned int n_simulations, unsigned int n_opts, unsigned int steps,
                              const GpuPriceOptions& opt = GpuPriceOptions()) {
  std::vector<double> expiries, strikes;
  std::vector<uint32_t> offsets(1, 0u);
  for (const options_chain& chain : all_chains) {
    expiries.push_back(chain.time_to_expiry);
    for (const option& o : chain.options) strikes.push_back(o.strike);
    offsets.push_back(static_cast<uint32_t>(strikes.size()));
  }
  if (strikes.size() != n_opts)
    throw std::runtime_error("price_gpu: n_opts does not match the option chains");
  hexo_price_request req{};
  req.p = hexo_hparams{p.v_0, p.v_m, p.rho, p.kappa, p.sigma};
  req.S = S;
  req.payoff = gpu_detail::payoff_of<Scheme>::value;
  req.n_chains = static_cast<uint32_t>(expiries.size());
  req.expiries = expiries.data();
  req.strike_offsets = offsets.data();
  req.strikes = strikes.data();
  req.n_paths = n_simulations;
  req.steps = steps;
  req.seed = opt.seed;
  req.normal_mode = opt.normal_mode;
  req.rng_mode = opt.rng_mode;
  req.schedule_mode = opt.schedule_mode;
  req.control_variate = opt.control_variate;
  req.drift_mode = opt.drift_mode;
  req.n_streams = opt.n_streams;
  std::vector<ffloat> prices(n_opts);
  if (opt.stderr_out) opt.stderr_out->assign(n_opts, 0.0);
  double* se = opt.stderr_out ? opt.stderr_out->data() : nullptr;
  const int rc = opt.n_gpus == 1 ? hexo_gpu_price(&req, prices.data(), se, nullptr)
                                 : hexo_gpu_price_multi(&req, opt.n_gpus, prices.data(), se, nullptr);
  if (rc != HEXO_OK)
    throw std::runtime_error(std::string("price_gpu: ") + hexo_gpu_last_error());
  return prices;
}

}  // namespace HSimulation
