/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * C entry points around the reference's OWN Monte-Carlo sources, which are
 * compiled verbatim from where they lie (-I/root/reference/src/inc; no source
 * is copied into this repo):
 *   src/HSimulation.tpp (via inc/HSimulation.h), src/RNG.cpp, inc/SDE.h,
 *   inc/AsianContract.h, inc/VanillaContract.h, inc/Types.h, inc/HDistribution.h
 * Two pieces the reference does not carry are supplied by the oracle:
 *   shishua.h  -> oracle/shishua.h (un-vendored dependency, see its header)
 *   ppnd16     -> oracle/hexo_oracle.c restatement of src/as241.f90 (no Fortran
 *                 compiler in this image); as-built REAL*4 mode by default.
 * The result, oracle/_ref/libhexo_ref.so, is used (a) to pin the plain-C
 * restatement, (b) to generate tests/golden fixtures and (c) as the CPU
 * baseline of bench.py ("kind": "reference").
 */
#include <omp.h>

#include <list>
#include <vector>

#include "AsianContract.h"
#include "HSimulation.h"
#include "RNG.h"
#include "VanillaContract.h"
#include "hexo_oracle.h"

static int g_normal_mode = ORACLE_NORMAL_F32;

/* the symbol src/RNG.cpp:6 binds (Fortran BIND(C,NAME='ppnd16'), as241.f90:15) */
extern "C" double ppnd16(double *p, int *ifault) {
  return g_normal_mode == ORACLE_NORMAL_F64 ? oracle_ppnd16_f64(*p, ifault)
                                            : oracle_ppnd16_f32(*p, ifault);
}

extern "C" {

void ref_set_normal_mode(int mode) { g_normal_mode = mode; }
void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_max_threads(void) { return omp_get_max_threads(); }

/* HSimulation::price<HQEAnderson<ffloat,Policy>>, src/HSimulation.tpp:10-51,
 * called exactly like src/Main.cpp:88 does */
int ref_price(int payoff, const double hparams[5], double S, unsigned n_chains,
              const double *expiries, const unsigned *strike_offsets, const double *strikes,
              unsigned n_simulations, unsigned steps, double *prices_out) {
  HParams p = {hparams[0], hparams[1], hparams[2], hparams[3], hparams[4]};
  std::list<options_chain> all_chains;
  unsigned n_opts = 0;
  for (unsigned k = 0; k < n_chains; ++k) {
    all_chains.emplace_back(static_cast<unsigned>(expiries[k] * trading_days), expiries[k]);
    options_chain &ch = all_chains.back();
    for (unsigned j = strike_offsets[k]; j < strike_offsets[k + 1]; ++j) {
      ch.options.push_back({0., 0., strikes[j], 0});
      ++n_opts;
    }
  }
  std::vector<ffloat> res;
  try {
    if (payoff == ORACLE_ASIAN)
      res = HSimulation::price<HSimulation::HQEAnderson<ffloat, AAsianCallNonAdaptive>>(
          p, S, all_chains, n_simulations, n_opts, steps);
    else
      res = HSimulation::price<HSimulation::HQEAnderson<ffloat, EuropeanCallNonAdaptive>>(
          p, S, all_chains, n_simulations, n_opts, steps);
  } catch (...) {
    return -1;
  }
  for (unsigned i = 0; i < n_opts; ++i) prices_out[i] = res[i];
  return 0;
}

/* RNG(size, seed) then a caller-chosen interleaving of get_urand / get_grand
 * (kinds[i] != 0 -> uniform), src/RNG.cpp:8-43, src/inc/RNG.h:39-50 */
int ref_rng_sequence(size_t size, unsigned seed, size_t n, const unsigned char *kinds,
                     double *out) {
  RNG rng(size, seed);
  for (size_t i = 0; i < n; ++i) out[i] = kinds[i] ? rng.get_urand() : rng.get_grand();
  return 0;
}

} /* extern "C" */
