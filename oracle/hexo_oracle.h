/*
 * oracle/hexo_oracle.h -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Plain-C CPU restatement of the reference's Monte-Carlo hot path
 * (MartinErhardt/HestonExotics): shishua wrapper (src/RNG.cpp), AS241/PPND16
 * (src/as241.f90), the Andersen-QE stepper and price driver
 * (src/HSimulation.tpp) and the Asian / European payoff policies
 * (src/inc/AsianContract.h, src/inc/VanillaContract.h).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (hestonexotics_b200/csrc) never links or calls it.
 *
 * PARITY STATUS: the reference holds no golden vectors for the MC path
 * (SURVEY.md finding 3).  This restatement is pinned against the reference's
 * OWN sources compiled verbatim into oracle/_ref/libhexo_ref.so (see
 * oracle/Makefile, tests/test_oracle_vs_ref.py) and against fixtures generated
 * from that build (tests/golden/).  The one unpinned piece is the raw shishua
 * byte stream versus upstream shishua (source absent, see oracle/shishua.h).
 */
#ifndef HEXO_ORACLE_H
#define HEXO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_ASIAN = 0, ORACLE_EUROPEAN = 1 };
/* inverse-normal arithmetic: as-built (REAL*4 internals, as241.f90:20-25) or
 * as-intended double precision */
enum { ORACLE_NORMAL_F32 = 0, ORACLE_NORMAL_F64 = 1 };

/* field order = HParams, src/inc/HDistribution.h:9-24 */
typedef struct {
  double v_0, v_m, rho, kappa, sigma;
} oracle_hparams;

typedef struct {
  oracle_hparams p;
  double S;
  int32_t payoff;                 /* ORACLE_ASIAN / ORACLE_EUROPEAN */
  uint32_t n_chains;              /* maturities, strictly increasing */
  const double *expiries;         /* [n_chains] time_to_expiry in years */
  const uint32_t *strike_offsets; /* [n_chains+1] prefix sums of chain sizes */
  const double *strikes;          /* [n_opts] chain-major */
  uint32_t steps;                 /* the reference's `steps` argument */
  int32_t drift_mode;             /* 0 = the reference's K0; 1 = Andersen's martingale-corrected
                                     K0* (NOT in the reference; include/hexo_gpu.h,
                                     HEXO_DRIFT_MARTINGALE) */
} oracle_contract;

/* ---- shishua (oracle/shishua.h) ------------------------------------------ */
/* fill `n_bytes` (multiple of 128) of the stream seeded with seed[4] */
int oracle_shishua_fill(const uint64_t seed[4], uint8_t *out, size_t n_bytes);

/* ---- uniform map, src/RNG.cpp:31 ----------------------------------------- */
double oracle_u64_to_unit(uint64_t x);

/* ---- AS241 / PPND16, src/as241.f90:15-119 -------------------------------- */
double oracle_ppnd16_f64(double p, int *ifault);
double oracle_ppnd16_f32(double p, int *ifault); /* as built */
/* z[i] = ppnd16(u64_to_unit(words[i])): RNG::setup_u + setup_g (src/RNG.cpp:31,39) on
 * caller-supplied raw words */
void oracle_normals_from_words(const uint64_t *words, double *z, size_t n, int normal_mode);
/* sums of the decimal coefficients, to compare with as241.f90:45,64,83 */
void oracle_ppnd16_hash_sums(double out[3]);

/* ---- RNG wrapper, src/RNG.cpp:8-43 + src/inc/RNG.h:39-50 ------------------ */
typedef struct oracle_rng oracle_rng;
oracle_rng *oracle_rng_new(size_t size, unsigned int seed, int normal_mode);
double oracle_rng_grand(oracle_rng *r);
double oracle_rng_urand(oracle_rng *r);
void oracle_rng_free(oracle_rng *r);

/* ---- price driver, src/HSimulation.tpp:10-51 ------------------------------
 * Race-free restatement: emulates `nthreads` OpenMP threads one after another
 * (thread t: seed 1<<t, n_sims/nthreads paths, :27-28) and divides by n_sims
 * (:40).  sum / sumsq (may be NULL) receive the raw payoff sums per option. */
int oracle_price_ref(const oracle_contract *c, unsigned int n_sims, unsigned int nthreads,
                     size_t rand_buf_size, int normal_mode, double *prices, double *sum,
                     double *sumsq);

/* ---- same model, GPU stream convention ------------------------------------
 * Stream s is the shishua stream seeded {seed, s, 0, 0}; its u64 outputs
 * x_0,x_1,... are consumed two per step: x_{2n} is the variance draw of the
 * stream's n-th step (normal if psi<1.5, else the uniform of the same word),
 * x_{2n+1} the log-spot normal.  The job's n_paths are split over
 * n_streams_total streams (stream s gets n_paths/n_streams_total paths, +1 if
 * s < n_paths % n_streams_total); this call runs streams
 * [stream_begin, stream_begin+stream_count). */
int oracle_price_stream(const oracle_contract *c, uint64_t seed, uint64_t n_paths,
                        uint64_t n_streams_total, uint64_t stream_begin, uint64_t stream_count,
                        int normal_mode, double *sum, double *sumsq);

/* Philox4x32-10 (Salmon et al., SC'11): optional generator of the GPU path, not in the reference */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* oracle_price_stream with a choice of generator: rng_mode 0 = shishua (as above), 1 = Philox:
 * the n-th stepper call of stream s uses Philox4x32-10(counter {n_lo,n_hi,s_lo,s_hi}, key
 * {seed_lo,seed_hi}); variance word = c0|c1<<32, spot word = c2|c3<<32. */
int oracle_price_stream_rng(const oracle_contract *c, int rng_mode, uint64_t seed, uint64_t n_paths,
                            uint64_t n_streams_total, uint64_t stream_begin, uint64_t stream_count,
                            int normal_mode, double *sum, double *sumsq);

/* The corrected time grid of include/hexo_gpu.h (HEXO_SCHEDULE_EXACT) -- NOT the reference's
 * loop: segment k covers (T_{k-1}, T_k] with n_k = max(1, round((T_k - T_{k-1}) steps / T_k))
 * equal steps, full trapezoid rule for the Asian average, X at the last step for the European
 * payoff.  Same stepper, same stream convention as oracle_price_stream_rng. */
int oracle_price_stream_exact(const oracle_contract *c, int rng_mode, uint64_t seed,
                              uint64_t n_paths, uint64_t n_streams_total, uint64_t stream_begin,
                              uint64_t stream_count, int normal_mode, double *sum, double *sumsq);

/* Stream pricer with the sums of the control variate c = final value - S (NOT in the
 * reference, which only suggests one at src/inc/HSimulation.h:51): out holds
 * [sum pf | sum pf^2 | sum pf c] per option, then [sum c | sum c^2] per maturity
 * (3 n_opts + 2 n_chains doubles). */
int oracle_price_stream_cv(const oracle_contract *c, int rng_mode, int exact_grid, uint64_t seed,
                           uint64_t n_paths, uint64_t n_streams_total, uint64_t stream_begin,
                           uint64_t stream_count, int normal_mode, double *out);

/* The same with the geometric-Asian control variate (NOT in the reference either): Asian
 * contracts only; out = [sum pf | sum pf^2 | sum pf c | sum c | sum c^2], n_opts entries each,
 * c_j = max(G - K_j, 0), G = exp(the Asian policy's accumulation applied to ln X). */
int oracle_price_stream_geo(const oracle_contract *c, int rng_mode, int exact_grid, uint64_t seed,
                            uint64_t n_paths, uint64_t n_streams_total, uint64_t stream_begin,
                            uint64_t stream_count, int normal_mode, double *out);

/* ---- tape replay ------------------------------------------------------------
 * tape[path][step][3] = {Z_V, U_V, Z_X}; the stepper takes Z_V or U_V according
 * to its branch.  finals[path][chain] receives the policy's final_value (Asian
 * average or interpolated X_T).  Returns the number of steps each path took,
 * or -1 if tape_steps was too short. */
int oracle_replay(const oracle_contract *c, const double *tape, uint64_t n_paths,
                  uint32_t tape_steps, double *finals);

/* number of stepper invocations a path needs until its last maturity is paid
 * (excludes the reference's trailing extra `++`, HSimulation.tpp:35) */
uint32_t oracle_steps_to_last_expiry(const oracle_contract *c);

/* ---- martingale correction (not in the reference) --------------------------------------------
 * K0* of one step of width h from variance V (Andersen 2008, Prop. 9), the constant term that
 * replaces K0 in HSimulation.tpp:80 when drift_mode = 1.  *branch = 0 quadratic, 1 exponential;
 * *corrected = 0 when M does not exist and the step keeps the reference drift. */
double oracle_qe_k0_star(const oracle_hparams *p, double h, double V, int *branch, int *corrected);

#ifdef __cplusplus
}
#endif
#endif
