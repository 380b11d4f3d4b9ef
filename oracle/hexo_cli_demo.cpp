/*
 * oracle/hexo_cli_demo.cpp -- TEST / DEMO INFRASTRUCTURE.
 *
 * `hexo -p asian all <SYM>` (reference src/Main.cpp:75-96) with the two network/database
 * dependencies replaced by files: a synthetic option chain instead of
 * WebAPI::get_all_option_chains (src/WebAPI.cpp:62-116) and HParams from the command line
 * instead of ParamsDB::fetch / calibrate (src/DB.cpp:7-23, src/Main.cpp:82-86).  The pricing
 * call is the reference's line with one identifier changed (price -> price_gpu) and the output
 * loop is the reference's own format (src/Main.cpp:89-93), using the reference's imp_vol
 * (src/BSM.cpp, compiled verbatim).
 *
 *   hexo_cli_demo <chain.csv> [v_0 v_m rho kappa sigma] [n_simulations steps] [geometric]
 * a trailing `geometric` prices with the geometric-Asian control variate (the one the reference
 * suggests at src/inc/HSimulation.h:51) through GpuPriceOptions -- same output format.
 * chain.csv:  first line `S,<spot>`; then `days_to_expiry,strike,bid,ask,volume` per option,
 * grouped by expiry in increasing order.
 */
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <list>
#include <sstream>
#include <string>
#include <vector>

#include "AsianContract.h"
#include "BSM.h"
#include "HSimulation.h"
#include "hexo_gpu_adapter.hpp"

typedef struct UD {
  ffloat S;
  std::list<options_chain> all_chains;
} underlying_data;

static unsigned int length(const std::list<options_chain>& all_chains) {  // Main.cpp:53-57
  unsigned int n = 0;
  for (const options_chain& c : all_chains) n += c.options.size();
  return n;
}

static underlying_data load_chain_file(const char* path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error(std::string("cannot open ") + path);
  underlying_data d{0., {}};
  std::string line;
  int last_days = -1;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::stringstream ss(line);
    std::string tok;
    std::vector<std::string> f;
    while (std::getline(ss, tok, ',')) f.push_back(tok);
    if (f.size() == 2 && f[0] == "S") {
      d.S = std::stod(f[1]);
      continue;
    }
    if (f.size() != 5) throw std::runtime_error("bad chain line: " + line);
    const int days = std::stoi(f[0]);
    if (days != last_days) {
      d.all_chains.emplace_back(static_cast<unsigned>(days), days / trading_days);
      last_days = days;
    }
    options_chain& ch = d.all_chains.back();
    option o{std::stod(f[3]), std::stod(f[2]), std::stod(f[1]), std::stoll(f[4])};
    ch.options.push_back(o);
    ch.min_strike = std::min(ch.min_strike, o.strike);
    ch.max_strike = std::max(ch.max_strike, o.strike);
  }
  return d;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::cerr << "usage: hexo_cli_demo <chain.csv> [v_0 v_m rho kappa sigma] [n_simulations steps] "
                 "[geometric]\n";
    return 2;
  }
  try {
    underlying_data ddata = load_chain_file(argv[1]);
    HParams p = {0.04, 0.04, -0.7, 2.0, 0.5};
    if (argc >= 7) p = {atof(argv[2]), atof(argv[3]), atof(argv[4]), atof(argv[5]), atof(argv[6])};
    const unsigned n_sim = argc >= 8 ? atoi(argv[7]) : 100000;  // 1e+5, Main.cpp:88
    const unsigned steps = argc >= 9 ? atoi(argv[8]) : 1000;    // 1e+3
    std::cout << "params, v0: " << p.v_0 << "\tv_m: " << p.v_m << "\trho: " << p.rho
              << "\tkappa: " << p.kappa << "\tsigma: " << p.sigma << std::endl;  // Main.cpp:87
    HSimulation::GpuPriceOptions opt;
    if (argc >= 10 && std::string(argv[9]) == "geometric") opt.control_variate = HEXO_CV_GEOMETRIC;
    std::vector<ffloat> results =
        HSimulation::price_gpu<HSimulation::HQEAnderson<ffloat, AAsianCallNonAdaptive>>(
            p, ddata.S, ddata.all_chains, n_sim, length(ddata.all_chains), steps, opt);
    unsigned int i = 0;
    for (const options_chain& opt_chain : ddata.all_chains)
      for (const option& opt : opt_chain.options)  // Main.cpp:90-93
        std::cout << "S: " << std::setw(10) << std::right << std::setfill(' ') << std::fixed
                  << std::setprecision(2) << ddata.S << "\tstrike: " << opt.strike
                  << "\tbid: " << opt.bid << "\task: " << opt.price
                  << "\tasian-option-price: " << results[i++] << "\tvolume: " << opt.volume
                  << "\timp vol: " << imp_vol(ddata.S, opt, opt_chain.time_to_expiry) << "\tlb: "
                  << ddata.S - std::exp(-yearly_risk_free * opt_chain.time_to_expiry) * opt.strike
                  << "\texpiry time: " << opt_chain.time_to_expiry * trading_days << '\n';
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
