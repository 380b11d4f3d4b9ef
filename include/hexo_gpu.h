/*
 * include/hexo_gpu.h -- C ABI of the B200-native Heston Monte-Carlo hot path.
 *
 * Drop-in boundary for MartinErhardt/HestonExotics ("hexo").  The reference has
 * no FFI: the "API" of its MC path is one C++ template,
 *     HSimulation::price<Scheme>(const HParams&, ffloat S,
 *         const std::list<options_chain>&, unsigned n_simulations,
 *         unsigned n_opts, unsigned steps) -> std::vector<ffloat>
 * (declared src/inc/HSimulation.h:61-63, defined src/HSimulation.tpp:10-51,
 * only production call site src/Main.cpp:88).  This header is what a binding
 * for that call binds; hestonexotics_b200/cpp/hexo_gpu_adapter.hpp wraps it
 * back into the template signature.
 *
 * Conventions: plain pointers and sizes, no exceptions across the boundary.
 * Every function returns HEXO_OK (0) or a negative hexo_status;
 * hexo_gpu_last_error() gives the message for the calling thread.  All buffers
 * are caller-owned HOST memory unless the name says `_device`.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with
 * HEXO_ERR_NO_DEVICE.
 */
#ifndef HEXO_GPU_H
#define HEXO_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEXO_GPU_ABI_VERSION 4 /* 2: rng_mode, schedule_mode, control_variate; 3: drift_mode;
                                 * 4: normals_from_words, replay steps_used_out, geometric control */

typedef enum {
  HEXO_OK = 0,
  HEXO_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, zero size, bad enum            */
  HEXO_ERR_NOT_INCREASING = -2,   /* expiries not strictly increasing; the         */
                                  /* reference asserts this, HSimulation.tpp:15-21 */
  HEXO_ERR_NO_DEVICE = -3,        /* no usable CUDA device                         */
  HEXO_ERR_CUDA = -4,             /* a CUDA runtime call or kernel failed          */
  HEXO_ERR_TOO_LARGE = -5,        /* n_opts does not fit the kernel's shared memory */
  HEXO_ERR_TAPE_TOO_SHORT = -6    /* replay tape has fewer steps than the schedule */
} hexo_status;

/* Scheme template argument of price<>: which OptionPolicy is plugged into
 * HQEAnderson (src/inc/AsianContract.h:14-48, src/inc/VanillaContract.h:14-40). */
typedef enum { HEXO_PAYOFF_ASIAN = 0, HEXO_PAYOFF_EUROPEAN = 1 } hexo_payoff;

/* Arithmetic of the inverse normal.  F32 is what the reference computes AS
 * BUILT: src/as241.f90:20-25 declares every local and coefficient default REAL
 * and nothing in Makefile.am promotes them.  The kernel's values lie within
 * 2e-6 (|z| <= 4) and 8 single-precision ulps of z of that arithmetic
 * (tests/test_normals_gpu.py); it switches from AS241's central rational function
 * to the tail formula at |q| = 0.45 instead of 0.425 (the central function's own
 * error is 2.5e-10 there, far below single precision).  F64 is the documented
 * AS241 accuracy ("1 part in 10**16", as241.f90:4), regions as in AS241. */
/* F32_PPND7 (optional, faster): single precision with the coefficients of PPND7, the routine the
 * same algorithm AS241 prescribes for single precision (degree 3/3 and 3/2 rational functions on
 * the same regions).  Its values agree with the as-built reference within the tolerance stated
 * for F32 (tests/test_normals_gpu.py).  Available for the default generator, drift and plain
 * sums; other combinations are refused. */
typedef enum {
  HEXO_NORMAL_F32 = 0,
  HEXO_NORMAL_F64 = 1,
  HEXO_NORMAL_F32_PPND7 = 2
} hexo_normal_mode;

/* Generator of the u64 words.  SHISHUA is the reference's (src/RNG.cpp:24-29, stream s seeded
 * {seed,s,0,0}).  PHILOX is an optional counter-based mode that the reference does not have:
 * the n-th stepper call of stream s draws Philox4x32-10(counter {n_lo,n_hi,s_lo,s_hi}, key
 * {seed_lo,seed_hi}) = (c0..c3); variance word c0|c1<<32, spot word c2|c3<<32. */
typedef enum { HEXO_RNG_SHISHUA = 0, HEXO_RNG_PHILOX = 1 } hexo_rng_mode;

/* Time grid.  REFERENCE reproduces the reference's loop bit for bit in its quirks (SURVEY
 * finding 6, Appendix B-3..5): time advances by repeated double addition, so (T=1, steps=252)
 * takes 253 stepper calls while (T=1, steps=1024) takes 1024 with the last trapezoid of the Asian
 * average replaced by X_N - X_{N-1}, and after an expiry the step width switches to T_next/steps
 * from wherever the grid stands.  EXACT is the corrected grid: segment k covers (T_{k-1}, T_k]
 * with n_k = max(1, round((T_k - T_{k-1}) steps / T_k)) steps of equal width, the Asian average
 * is the full trapezoid rule and the European payoff reads X at the last step. */
typedef enum { HEXO_SCHEDULE_REFERENCE = 0, HEXO_SCHEDULE_EXACT = 1 } hexo_schedule_mode;

/* Control variate (the reference only suggests one, src/inc/HSimulation.h:51).  UNDERLYING uses
 * what the path already holds: c = final value - S, i.e. the arithmetic average (Asian) or the
 * terminal spot (European) minus the initial spot.  E[c] = 0 because the model has no drift
 * (HSimulation.tpp:75-80), so  price = mean(payoff) - beta mean(c)  with
 * beta = cov(payoff, c) / var(c) estimated from the same paths; the standard error shrinks by
 * sqrt(1 - corr(payoff, c)^2).  The sums gain [sum payoff c] per option and [sum c | sum c^2]
 * per maturity (hexo_gpu_sums_len). */
/* GEOMETRIC (Asian payoff) is the control the reference names: next to the arithmetic average
 * A = sum w_i X_i / T the kernel accumulates Y = sum w_i ln X_i / sum w_i with the same weights
 * (including the reference grid's last-step rule) and uses, per option, c_j = max(exp(Y) - K_j, 0).
 * E[c_j] is the price of the discretely monitored geometric-Asian call under Heston at r = 0,
 * evaluated on the host from the affine transform of ln X at the grid dates and one Fourier
 * integral (hexo_heston_geometric_asian).  That mean belongs to the exact process, the simulation
 * to its QE discretisation: the estimator also removes the part of the discretisation error the
 * two payoffs share.  Sums: [sum pf | sum pf^2 | sum pf c | sum c | sum c^2], n_opts each. */
typedef enum {
  HEXO_CV_NONE = 0,
  HEXO_CV_UNDERLYING = 1,
  HEXO_CV_GEOMETRIC = 2
} hexo_control_variate;

/* Drift of the log-spot step.  REFERENCE is the reference's (HSimulation.tpp:75-80): the constant
 * K0 of Andersen's scheme, under which the simulated spot is only approximately a martingale
 * (SURVEY finding 7).  MARTINGALE replaces K0 step by step with Andersen's K0* (L. Andersen,
 * "Simple and efficient simulation of the Heston stochastic volatility model", 2008, Prop. 9):
 *   K0* = -ln M - (K1 + K3/2) V,   M = E[exp(A V') | V],   A = K2 + K4/2,
 *   psi <  1.5:  M = exp(A b^2 a / (1 - 2 A a)) / sqrt(1 - 2 A a)
 *   psi >= 1.5:  M = p + beta (1 - p) / (beta - A)
 * so that E[X' | X, V] = X exactly and E[X_t] = S on every grid.  Where M does not exist
 * (A >= 1/(2a) resp. A >= beta; not reachable with rho <= 0) the step keeps the reference drift. */
typedef enum { HEXO_DRIFT_REFERENCE = 0, HEXO_DRIFT_MARTINGALE = 1 } hexo_drift_mode;

/* HParams, src/inc/HDistribution.h:9-24 -- same field order, same meaning */
typedef struct {
  double v_0;   /* initial variance            */
  double v_m;   /* long-term variance (theta)  */
  double rho;   /* spot/variance correlation   */
  double kappa; /* mean-reversion rate         */
  double sigma; /* vol of variance (epsilon)   */
} hexo_hparams;

/* One price<Scheme>() call.  std::list<options_chain> (src/inc/Types.h:37-58) is
 * flattened: chain k has expiry expiries[k] and strikes
 * strikes[strike_offsets[k] .. strike_offsets[k+1]).  n_opts is
 * strike_offsets[n_chains].  Output order = the reference's: chain-major, then
 * option order (HSimulation.tpp:39-40). */
typedef struct {
  hexo_hparams p;
  double S;                       /* spot, `S` of price<>                       */
  int32_t payoff;                 /* hexo_payoff                                */
  uint32_t n_chains;
  const double *expiries;         /* [n_chains] options_chain::time_to_expiry   */
  const uint32_t *strike_offsets; /* [n_chains+1]                               */
  const double *strikes;          /* [n_opts] option::strike                    */
  uint64_t n_paths;               /* `n_simulations` (the reference: unsigned)  */
  uint32_t steps;                 /* `steps`: step width = expiry/steps         */
  uint64_t seed;                  /* stream s is shishua seeded {seed,s,0,0}    */
  int32_t normal_mode;            /* hexo_normal_mode                           */
  int32_t rng_mode;               /* hexo_rng_mode; 0 = the reference's shishua */
  uint64_t n_streams;             /* independent RNG streams the n_paths are    */
                                  /* split over; 0 = pick for the device(s).    */
                                  /* For a fixed (seed,n_paths,n_streams) the   */
                                  /* sums do not depend on how streams are      */
                                  /* sharded over GPUs.                         */
  int32_t schedule_mode;          /* hexo_schedule_mode; 0 = the reference's    */
  int32_t control_variate;        /* hexo_control_variate; 0 = none             */
  int32_t drift_mode;             /* hexo_drift_mode; 0 = the reference's       */
} hexo_price_request;

typedef struct {
  uint64_t n_streams;      /* streams actually used for the whole job          */
  uint64_t steps_per_path; /* stepper invocations per path (N*, e.g. 253)      */
  uint64_t path_steps;     /* n_paths x `steps` of this call (metric unit)     */
  uint32_t grid, block;    /* launch geometry of the path kernel               */
  uint32_t smem_bytes;
  uint32_t kernel_launches; /* kernels this call launched                      */
  float kernel_ms;          /* CUDA-event time of the path kernel + finalize   */
} hexo_gpu_stats;

/* Step schedule of one price<>() call: the reference advances time by repeated
 * double addition and compares with the expiry (HSimulation.tpp:35-36,83-84,
 * SDE.h:30), so the number of steps and the final interpolation weight are a
 * property of (expiries, steps) in double arithmetic.  Host-only, no GPU. */
typedef struct {
  uint32_t n_steps; /* stepper invocations made while heading for this expiry  */
  double h;         /* step width = expiry/steps (AsianContract.h:35-38)       */
  double w;         /* (expiry - prev_time)/h at the paying step (:31-32)      */
  double expiry;
} hexo_segment;

/* ---- lifecycle ------------------------------------------------------------ */
int hexo_gpu_abi_version(void);
/* select the CUDA device for the calling process (one process per GPU) */
int hexo_gpu_init(int device);
int hexo_gpu_shutdown(void);
int hexo_gpu_device_count(void);
const char *hexo_gpu_last_error(void);

/* ---- host-side schedule (no GPU needed) ------------------------------------ */
int hexo_gpu_schedule(const double *expiries, uint32_t n_chains, uint32_t steps,
                      hexo_segment *segments_out /* [n_chains] */);
/* the same for HEXO_SCHEDULE_EXACT (w is 1 for every segment) */
int hexo_gpu_schedule_exact(const double *expiries, uint32_t n_chains, uint32_t steps,
                            hexo_segment *segments_out);

/* ---- K1: the fused pricing path --------------------------------------------
 * Replaces HSimulation::price<Scheme> (HSimulation.tpp:10-51).  prices_out
 * [n_opts] = sum(payoff)/n_paths (:40).  stderr_out (or NULL) = Monte-Carlo
 * standard error per option, which the reference does not report. */
int hexo_gpu_price(const hexo_price_request *req, double *prices_out, double *stderr_out,
                   hexo_gpu_stats *stats);

/* n_reqs independent price<>() calls in one submission -- the shape of Monte-Carlo pricing
 * inside a calibration loop (src/Main.cpp:88 evaluated for many HParams).  Request i runs on CUDA
 * stream i % n_lanes (0 = 16 lanes) of the current device so that small jobs overlap; each job's
 * result is exactly hexo_gpu_price's for the same request.  prices_out / stderr_out (or NULL)
 * are request-major: request i's n_opts values follow request i-1's.  stats (or NULL) is
 * [n_reqs]; kernel_ms there is the device time of the whole batch. */
int hexo_gpu_price_batch(const hexo_price_request *reqs, uint32_t n_reqs, uint32_t n_lanes,
                         double *prices_out, double *stderr_out, hexo_gpu_stats *stats);

/* Number of doubles a shard's sums hold for this request: 2 n_opts ([sum payoff | sum payoff^2]),
 * 3 n_opts + 2 n_chains with HEXO_CV_UNDERLYING ([.. | sum payoff c | sum c | sum c^2]), 5 n_opts
 * with HEXO_CV_GEOMETRIC.
 * Sums of disjoint stream ranges add up; hexo_gpu_finish turns the total into prices and
 * standard errors exactly as hexo_gpu_price does (host only, no GPU). */
size_t hexo_gpu_sums_len(const hexo_price_request *req);
int hexo_gpu_finish(const hexo_price_request *req, const double *sums, double *prices_out,
                    double *stderr_out);
/* Known means of the geometric-Asian control (HEXO_CV_GEOMETRIC): means_out[j] =
 * E max(G - K_j, 0) for every option of an Asian request, G the geometric average on the
 * request's time grid.  Host only, no GPU. */
int hexo_heston_geometric_asian(const hexo_price_request *req, double *means_out);

/* The same call spread over the first n_gpus devices of this process (n_gpus <= 0: all
 * visible devices): the single-process counterpart of the one-rank-per-GPU path, for callers
 * like the reference's CLI.  Streams are split over devices like ranks split them; the sums
 * are added on the host. */
int hexo_gpu_price_multi(const hexo_price_request *req, int n_gpus, double *prices_out,
                         double *stderr_out, hexo_gpu_stats *stats);

/* The same path for one shard of the job: streams [stream_begin,
 * stream_begin+stream_count) of req->n_streams (which must be non-zero here).
 * sums_out[0..n_opts) = sum of payoffs, sums_out[n_opts..2 n_opts) = sum of
 * squared payoffs over this shard's paths (hexo_gpu_sums_len(req) doubles in all:
 * a control variate appends its own sums); summing shards (e.g. by an all-reduce)
 * and dividing by n_paths gives the price -- hexo_gpu_finish does exactly that. */
int hexo_gpu_price_shard(const hexo_price_request *req, uint64_t stream_begin,
                         uint64_t stream_count, double *sums_out, hexo_gpu_stats *stats);

/* Same, asynchronous: enqueues on `cuda_stream` (a cudaStream_t, 0 = default)
 * and leaves the hexo_gpu_sums_len(req) sums in DEVICE memory at sums_device, ready for an
 * NCCL all-reduce on the same stream.  No host synchronisation. */
int hexo_gpu_price_shard_device(const hexo_price_request *req, uint64_t stream_begin,
                                uint64_t stream_count, double *sums_device, void *cuda_stream,
                                hexo_gpu_stats *stats);

/* ---- prepared launches -------------------------------------------------------
 * A plan holds everything one shard needs in device memory (segment constants,
 * strikes, per-block partials, sums), so that repeated launches move no input
 * data: hexo_gpu_plan_launch only enqueues the path kernel and the reduction on
 * `cuda_stream`.  sums_device == NULL leaves the sums in the plan's own buffer
 * (hexo_gpu_plan_sums_device). */
typedef struct hexo_gpu_plan hexo_gpu_plan;
int hexo_gpu_plan_create(const hexo_price_request *req, uint64_t stream_begin,
                         uint64_t stream_count, hexo_gpu_plan **plan_out);
int hexo_gpu_plan_launch(hexo_gpu_plan *plan, double *sums_device, void *cuda_stream);
double *hexo_gpu_plan_sums_device(hexo_gpu_plan *plan);
int hexo_gpu_plan_stats(const hexo_gpu_plan *plan, hexo_gpu_stats *stats);
int hexo_gpu_plan_destroy(hexo_gpu_plan *plan);

/* default stream count for a job on `n_gpus` devices like the current one */
uint64_t hexo_gpu_default_streams(uint64_t n_paths, uint32_t n_opts, int n_gpus);

/* ---- K2: raw shishua bytes (replaces prng_init + prng_gen, RNG.cpp:24,29) --- */
int hexo_gpu_shishua_fill(const uint64_t seed[4], uint8_t *bytes_out, size_t n_bytes);
/* many streams at once: stream i is seeded {seed, first_stream+i, 0, 0};
 * bytes_out[i*bytes_per_stream ...] (bytes_per_stream multiple of 128) */
int hexo_gpu_shishua_streams(uint64_t seed, uint64_t first_stream, uint32_t n_streams,
                             uint8_t *bytes_out, size_t bytes_per_stream);

/* Optional Philox mode (hexo_rng_mode): n Philox4x32-10 blocks, counters[4n], keys[2n] ->
 * out[4n] (Random123 known-answer vectors), and the first `words_per_stream` (multiple of 16)
 * u64 words streams [first_stream, first_stream+n_streams) hand to the stepper. */
int hexo_gpu_philox4x32(const uint32_t *counters, const uint32_t *keys, uint32_t *out, size_t n);
int hexo_gpu_philox_streams(uint64_t seed, uint64_t first_stream, uint32_t n_streams,
                            uint64_t *words_out, size_t words_per_stream);

/* ---- K3: uniform map and inverse normal (RNG.cpp:31, as241.f90:15-119) ------ */
int hexo_gpu_u64_to_unit(const uint64_t *bits_in, double *u_out, size_t n);
int hexo_gpu_ppnd16(const double *u_in, double *z_out, size_t n, int normal_mode);
/* The same transform EXACTLY as the fused kernel K1 runs it (RNG.cpp:31,39 + as241.f90:85-118
 * on raw generator words): the batched central phase and the warp-cooperative tail phase of
 * the kernel's shared-memory ring, fed with caller-supplied words instead of generator rounds.
 * z_out[i] is the normal K1 would hand to the stepper for word words_in[i]; compare with
 * ppnd16(u64_to_unit(word)).  hexo_gpu_ppnd16 above evaluates the scalar routine, which K1 does
 * not call. */
int hexo_gpu_normals_from_words(const uint64_t *words_in, double *z_out, size_t n,
                                int normal_mode);

/* ---- K4: tape replay of stepper + payoff policy ------------------------------
 * tape[path][step][3] = {Z_V, U_V, Z_X} (normals/uniform the reference's RNG
 * would have handed to HSimulation.tpp:67,72,80).  finals_out[path][chain] =
 * the policy's final_value (Asian average / interpolated X_T).  Uses
 * req->{p,S,payoff,n_chains,expiries,steps}; strikes are not needed. */
int hexo_gpu_replay(const hexo_price_request *req, const double *tape, uint64_t n_paths,
                    uint32_t tape_steps, double *finals_out,
                    uint32_t *steps_used_out /* stepper calls per path, or NULL */);

/* ---- semi-analytic European benchmark (host only) ---------
 * The Heston characteristic function with its analytic gradient and the SWIFT
 * pricer of the reference, used to benchmark the Monte-Carlo European price and
 * (in the reference) to calibrate the HParams the path consumes.  No GPU needed. */

/* swift_parameters, src/inc/SWIFT.h:24-48 */
typedef struct {
  uint32_t m;         /* wavelet scale                      */
  uint32_t exp2_m;    /* 2^m                                */
  double sqrt_exp2_m; /* sqrt(2^m)                          */
  double lower;       /* lower integration bound            */
  double upper;       /* upper integration bound            */
  int32_t k_1, k_2;   /* wavelet translation range          */
  uint32_t J;         /* Fourier size (power of two)        */
} hexo_swift_params;

/* HDistribution::chf_chf_grad(u) (src/HDistribution.cpp:52-88): out = (re,im) pairs of
 * chf and its partials in HParams order v_0, v_m, rho, kappa, sigma; risk-free rate is not
 * part of the reference's chf. */
int hexo_heston_chf(const hexo_hparams *p, double tau, double u_re, double u_im, double out[12]);
/* Cumulants 1, 2 and 4 of the log-return at time tau for a stationary start (v_0 := v_m), the
 * quantities HDistribution::first/second/fourth_order_moment return (src/HDistribution.cpp:90-113)
 * and SwiftParameters turns into the integration range (src/SWIFT.cpp:21-35). */
int hexo_heston_cumulants(const hexo_hparams *p, double tau, double out[3]);

/* SwiftParameters(distr, S, chain) (src/SWIFT.cpp:21-35); truncation_precision <= 0
 * selects the reference's release value 1e-7 (SWIFT.cpp:12-16). */
int hexo_swift_default_params(const hexo_hparams *p, double tau, double risk_free, double S,
                              double min_strike, double max_strike, double truncation_precision,
                              hexo_swift_params *params_out);

/* SWIFT::price_opts / price_opts_grad for one chain (src/SWIFT.cpp:87-118).
 * grad_out (or NULL): [n_strikes][5], partials in HParams order. */
int hexo_swift_price_chain(const hexo_swift_params *params, const hexo_hparams *p, double tau,
                           double risk_free, double S, const double *strikes, uint32_t n_strikes,
                           double *prices_out, double *grad_out);

/* ---- FP64 pipe peak (roofline denominator; not in MEASURED_PEAKS.json) ------
 * Runs a register-resident DFMA chain kernel; returns FP64 flop/s (FMA = 2). */
int hexo_gpu_measure_fp64_peak(double *flops_out, float *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* HEXO_GPU_H */
