/*
 * oracle/dropin_main.cpp -- TEST INFRASTRUCTURE: the drop-in demonstrated in the
 * reference's own language.  Compiled against the reference's unchanged headers
 * (from /root/reference/src/inc) and the adapter
 * hestonexotics_b200/cpp/hexo_gpu_adapter.hpp, it does what the PRICE case of
 * src/Main.cpp:75-96 does -- with a synthetic option chain and fixed HParams in
 * place of WebAPI / ParamsDB (no network) -- once through the reference's CPU
 * price<> and once through price_gpu<>, and prints both as JSON.
 * Built by oracle/Makefile into oracle/_ref/hexo_dropin; run by
 * tests/test_gpu_dropin.py.
 */
#include <omp.h>

#include <cstdio>
#include <cstdlib>
#include <list>
#include <vector>

#include "AsianContract.h"
#include "HSimulation.h"
#include "VanillaContract.h"
#include "hexo_gpu_adapter.hpp"

using AsianScheme = HSimulation::HQEAnderson<ffloat, AAsianCallNonAdaptive>;
using EuroScheme = HSimulation::HQEAnderson<ffloat, EuropeanCallNonAdaptive>;

static void print_vec(const char* name, const std::vector<ffloat>& v, bool last = false) {
  std::printf("\"%s\": [", name);
  for (size_t i = 0; i < v.size(); ++i) std::printf("%s%.17g", i ? ", " : "", v[i]);
  std::printf("]%s", last ? "" : ", ");
}

int main(int argc, char** argv) {
  const char* kind = argc > 1 ? argv[1] : "asian";
  const unsigned n_cpu = argc > 2 ? std::atoi(argv[2]) : 20000;
  const unsigned n_gpu = argc > 3 ? std::atoi(argv[3]) : 400000;
  const unsigned steps = argc > 4 ? std::atoi(argv[4]) : 100;
  const HParams p = {0.04, 0.04, -0.7, 2.0, 0.5};
  const ffloat S = 100.0;
  std::list<options_chain> all_chains;  // what WebAPI::get_all_option_chains would deliver
  unsigned n_opts = 0;
  for (ffloat T : {0.25, 0.5, 1.0}) {
    all_chains.emplace_back(static_cast<unsigned>(T * trading_days), T);
    for (ffloat K : {90.0, 100.0, 110.0}) {
      all_chains.back().options.push_back({0., 0., K, 0});
      ++n_opts;
    }
  }
  omp_set_num_threads(1);  // the reference's accumulation is racy with more (HSimulation.tpp:40)
  std::vector<ffloat> ref, gpu, se;
  HSimulation::GpuPriceOptions opt;
  opt.stderr_out = &se;
  try {
    if (kind[0] == 'a') {
      ref = HSimulation::price<AsianScheme>(p, S, all_chains, n_cpu, n_opts, steps);
      gpu = HSimulation::price_gpu<AsianScheme>(p, S, all_chains, n_gpu, n_opts, steps, opt);
    } else {
      ref = HSimulation::price<EuroScheme>(p, S, all_chains, n_cpu, n_opts, steps);
      gpu = HSimulation::price_gpu<EuroScheme>(p, S, all_chains, n_gpu, n_opts, steps, opt);
    }
  } catch (const std::exception& e) {
    std::printf("{\"error\": \"%s\"}\n", e.what());
    return 1;
  }
  std::printf("{\"n_cpu\": %u, \"n_gpu\": %u, ", n_cpu, n_gpu);
  print_vec("reference", ref);
  print_vec("gpu", gpu);
  print_vec("gpu_stderr", se, true);
  std::printf("}\n");
  return 0;
}
