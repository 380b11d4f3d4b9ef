/*
 * oracle/hexo_oracle.c -- TEST INFRASTRUCTURE (oracle), not product code.
 * See oracle/hexo_oracle.h for scope and parity status.
 *
 * Build: strict IEEE double, no fast-math, no FMA contraction (oracle/Makefile).
 * Every function names the reference lines it restates.
 */
#include "hexo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ppnd16_coef.h"
#include "shishua.h"

/* ========================================================================== */
/* shishua                                                                    */
/* ========================================================================== */

int oracle_shishua_fill(const uint64_t seed[4], uint8_t *out, size_t n_bytes) {
  if (n_bytes & 127u) return -1;
  prng_state s;
  uint64_t sd[4] = {seed[0], seed[1], seed[2], seed[3]};
  prng_init(&s, sd); /* call site: src/RNG.cpp:24 */
  prng_gen(&s, out, n_bytes); /* call site: src/RNG.cpp:29 */
  return 0;
}

/* src/RNG.cpp:31: *x = (ffloat)u64 / (ffloat)(fuint)-1.  (double)(2^64-1) rounds
 * to 2^64, so this is RN(u64) * 2^-64 exactly, in [0, 1] inclusive. */
double oracle_u64_to_unit(uint64_t x) { return (double)x / (double)UINT64_MAX; }

/* ========================================================================== */
/* AS241 / PPND16                                                             */
/* ========================================================================== */

/* src/as241.f90:85-118 evaluated in double precision ("as intended") */
double oracle_ppnd16_f64(double p, int *ifault) {
  double q, r, z;
  if (ifault) *ifault = 0;
  q = p - 0.5;                                   /* :86 */
  if (fabs(q) <= PPND_SPLIT1) {                  /* :88 */
    r = PPND_CONST1 - q * q;                     /* :89 */
    return q *
           (((((((PPND_A7 * r + PPND_A6) * r + PPND_A5) * r + PPND_A4) * r + PPND_A3) * r +
              PPND_A2) * r + PPND_A1) * r + PPND_A0) /
           (((((((PPND_B7 * r + PPND_B6) * r + PPND_B5) * r + PPND_B4) * r + PPND_B3) * r +
              PPND_B2) * r + PPND_B1) * r + 1.0); /* :90-91 */
  }
  r = (q < 0.0) ? p : 1.0 - p;                   /* :94-98 */
  if (r <= 0.0) {                                /* :99-103 */
    if (ifault) *ifault = 1;
    return 0.0;
  }
  r = sqrt(-log(r));                             /* :104 */
  if (r <= PPND_SPLIT2) {                        /* :105 */
    r -= PPND_CONST2;                            /* :106 */
    z = (((((((PPND_C7 * r + PPND_C6) * r + PPND_C5) * r + PPND_C4) * r + PPND_C3) * r +
           PPND_C2) * r + PPND_C1) * r + PPND_C0) /
        (((((((PPND_D7 * r + PPND_D6) * r + PPND_D5) * r + PPND_D4) * r + PPND_D3) * r +
           PPND_D2) * r + PPND_D1) * r + 1.0);   /* :107-109 */
  } else {
    r -= PPND_SPLIT2;                            /* :111 */
    z = (((((((PPND_E7 * r + PPND_E6) * r + PPND_E5) * r + PPND_E4) * r + PPND_E3) * r +
           PPND_E2) * r + PPND_E1) * r + PPND_E0) /
        (((((((PPND_F7 * r + PPND_F6) * r + PPND_F5) * r + PPND_F4) * r + PPND_F3) * r +
           PPND_F2) * r + PPND_F1) * r + 1.0);   /* :112-114 */
  }
  return (q < 0.0) ? -z : z;                     /* :116 */
}

/* The routine AS BUILT by the reference: every local and every coefficient is
 * default REAL (as241.f90:20-25, no -fdefault-real-8 in Makefile.am:62-91), only
 * the dummy argument P and the function result are C_DOUBLE (:15-19).  So mixed
 * expressions with P are evaluated in double and then rounded into the REAL
 * locals Q / R; everything else is single precision. */
#define F(x) ((float)(x))
double oracle_ppnd16_f32(double p, int *ifault) {
  float q, r, z;
  if (ifault) *ifault = 0;
  q = (float)(p - (double)0.5f);                 /* :86  Q = P - HALF */
  if (fabsf(q) <= F(PPND_SPLIT1)) {              /* :88 */
    r = F(PPND_CONST1) - q * q;                  /* :89 */
    z = q *
        (((((((F(PPND_A7) * r + F(PPND_A6)) * r + F(PPND_A5)) * r + F(PPND_A4)) * r +
             F(PPND_A3)) * r + F(PPND_A2)) * r + F(PPND_A1)) * r + F(PPND_A0)) /
        (((((((F(PPND_B7) * r + F(PPND_B6)) * r + F(PPND_B5)) * r + F(PPND_B4)) * r +
             F(PPND_B3)) * r + F(PPND_B2)) * r + F(PPND_B1)) * r + 1.0f);
    return (double)z;
  }
  r = (q < 0.0f) ? (float)p : (float)((double)1.0f - p); /* :94-98 */
  if (r <= 0.0f) {
    if (ifault) *ifault = 1;
    return 0.0;
  }
  r = sqrtf(-logf(r));                           /* :104 */
  if (r <= F(PPND_SPLIT2)) {
    r -= F(PPND_CONST2);
    z = (((((((F(PPND_C7) * r + F(PPND_C6)) * r + F(PPND_C5)) * r + F(PPND_C4)) * r +
           F(PPND_C3)) * r + F(PPND_C2)) * r + F(PPND_C1)) * r + F(PPND_C0)) /
        (((((((F(PPND_D7) * r + F(PPND_D6)) * r + F(PPND_D5)) * r + F(PPND_D4)) * r +
           F(PPND_D3)) * r + F(PPND_D2)) * r + F(PPND_D1)) * r + 1.0f);
  } else {
    r -= F(PPND_SPLIT2);
    z = (((((((F(PPND_E7) * r + F(PPND_E6)) * r + F(PPND_E5)) * r + F(PPND_E4)) * r +
           F(PPND_E3)) * r + F(PPND_E2)) * r + F(PPND_E1)) * r + F(PPND_E0)) /
        (((((((F(PPND_F7) * r + F(PPND_F6)) * r + F(PPND_F5)) * r + F(PPND_F4)) * r +
           F(PPND_F3)) * r + F(PPND_F2)) * r + F(PPND_F1)) * r + 1.0f);
  }
  if (q < 0.0f) z = -z;
  return (double)z;
}
#undef F

static double ppnd16(double p, int normal_mode) {
  int ifault;
  return normal_mode == ORACLE_NORMAL_F64 ? oracle_ppnd16_f64(p, &ifault)
                                          : oracle_ppnd16_f32(p, &ifault);
}

/* RNG::setup_u + RNG::setup_g on caller-supplied raw words (src/RNG.cpp:31,39): the oracle
 * side of hexo_gpu_normals_from_words */
void oracle_normals_from_words(const uint64_t *words, double *z, size_t n, int normal_mode) {
  for (size_t i = 0; i < n; ++i) z[i] = ppnd16(oracle_u64_to_unit(words[i]), normal_mode);
}

/* ========================================================================== */
/* RNG wrapper: two ring buffers fed by ONE shishua state                     */
/* ========================================================================== */

struct oracle_rng {
  double *buf_start_u, *buf_end_u, *buf_cur_u;
  double *buf_start_g, *buf_end_g, *buf_cur_g;
  int normal_mode;
  prng_state s;
};

/* src/RNG.cpp:28-33 */
static double *rng_setup_u(oracle_rng *r, double *start, double *end) {
  prng_gen(&r->s, (uint8_t *)start, sizeof(double) * (size_t)(end - start));
  for (double *x = start; x < end; ++x) {
    uint64_t bits;
    memcpy(&bits, x, 8);
    *x = oracle_u64_to_unit(bits);
  }
  return start;
}

/* src/RNG.cpp:34-43 */
static double *rng_setup_g(oracle_rng *r) {
  rng_setup_u(r, r->buf_start_g, r->buf_end_g);
  for (double *x = r->buf_start_g; x != r->buf_end_g; ++x) *x = ppnd16(*x, r->normal_mode);
  return r->buf_start_g;
}

/* src/RNG.cpp:8-27: the U buffer is filled first, then the G buffer */
oracle_rng *oracle_rng_new(size_t size, unsigned int seed, int normal_mode) {
  if ((size * sizeof(double)) & 127u) return NULL; /* :10 */
  oracle_rng *r = (oracle_rng *)calloc(1, sizeof(*r));
  if (!r) return NULL;
  r->normal_mode = normal_mode;
  r->buf_start_g = (double *)aligned_alloc(128, size * sizeof(double));
  r->buf_start_u = (double *)aligned_alloc(128, size * sizeof(double));
  r->buf_end_u = r->buf_start_u + size;
  r->buf_end_g = r->buf_start_g + size;
  uint64_t my_seed[4] = {seed, 0, 0, 0}; /* src/inc/RNG.h:24, src/RNG.cpp:9 */
  prng_init(&r->s, my_seed);
  r->buf_cur_u = rng_setup_u(r, r->buf_start_u, r->buf_end_u);
  r->buf_cur_g = rng_setup_g(r);
  return r;
}

/* src/inc/RNG.h:39-42 */
double oracle_rng_grand(oracle_rng *r) {
  if (r->buf_cur_g == r->buf_end_g) r->buf_cur_g = rng_setup_g(r);
  return *(r->buf_cur_g++);
}

/* src/inc/RNG.h:47-50 */
double oracle_rng_urand(oracle_rng *r) {
  if (r->buf_cur_u == r->buf_end_u) r->buf_cur_u = rng_setup_u(r, r->buf_start_u, r->buf_end_u);
  return *(r->buf_cur_u++);
}

void oracle_rng_free(oracle_rng *r) {
  if (!r) return;
  free(r->buf_start_g);
  free(r->buf_start_u);
  free(r);
}

/* ========================================================================== */
/* draw sources                                                               */
/* ========================================================================== */

typedef struct draw_src {
  /* one call per stepper invocation, BEFORE the branch is known: makes the
   * step's variance normal, variance uniform and log-spot normal available */
  double (*variance_normal)(struct draw_src *);
  double (*variance_uniform)(struct draw_src *);
  double (*spot_normal)(struct draw_src *);
  void (*end_step)(struct draw_src *);
  void *ctx;
} draw_src;

/* (a) the reference's RNG object: draws are pulled on demand */
static double refsrc_g(draw_src *d) { return oracle_rng_grand((oracle_rng *)d->ctx); }
static double refsrc_u(draw_src *d) { return oracle_rng_urand((oracle_rng *)d->ctx); }
static void refsrc_end(draw_src *d) { (void)d; }

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += W0;
    k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* (b) GPU stream convention: two consecutive u64 words per step */
typedef struct {
  prng_state s;
  uint64_t block[16];
  int pos; /* next unread word in block, 16 = empty */
  int normal_mode;
  uint64_t w_var, w_spot;
  int have;
  int rng_mode;          /* 0 shishua, 1 philox */
  uint64_t seed, stream, n_step;
} stream_ctx;

static void stream_fetch(stream_ctx *c) {
  if (c->have) return;
  if (c->rng_mode == 1) {
    const uint32_t ctr[4] = {(uint32_t)c->n_step, (uint32_t)(c->n_step >> 32), (uint32_t)c->stream,
                             (uint32_t)(c->stream >> 32)};
    const uint32_t key[2] = {(uint32_t)c->seed, (uint32_t)(c->seed >> 32)};
    uint32_t o[4];
    oracle_philox4x32_10(ctr, key, o);
    c->w_var = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
    c->w_spot = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
    c->n_step++;
    c->have = 1;
    return;
  }
  if (c->pos == 16) {
    prng_gen(&c->s, (uint8_t *)c->block, 128);
    c->pos = 0;
  }
  uint8_t *b = (uint8_t *)c->block;
  uint64_t w[2];
  for (int k = 0; k < 2; ++k) { /* little-endian words, order o0[0..3] o1.. o2.. o3.. */
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | b[8 * (c->pos + k) + i];
    w[k] = v;
  }
  c->w_var = w[0];
  c->w_spot = w[1];
  c->pos += 2;
  c->have = 1;
}
static double strsrc_gv(draw_src *d) {
  stream_ctx *c = (stream_ctx *)d->ctx;
  stream_fetch(c);
  return ppnd16(oracle_u64_to_unit(c->w_var), c->normal_mode);
}
static double strsrc_uv(draw_src *d) {
  stream_ctx *c = (stream_ctx *)d->ctx;
  stream_fetch(c);
  return oracle_u64_to_unit(c->w_var);
}
static double strsrc_gx(draw_src *d) {
  stream_ctx *c = (stream_ctx *)d->ctx;
  stream_fetch(c);
  return ppnd16(oracle_u64_to_unit(c->w_spot), c->normal_mode);
}
static void strsrc_end(draw_src *d) { ((stream_ctx *)d->ctx)->have = 0; }

/* (c) tape */
typedef struct {
  const double *tape; /* [steps][3] for the current path */
  uint32_t step, n_steps;
  int overrun;
} tape_ctx;
static double tapesrc_get(draw_src *d, int k) {
  tape_ctx *c = (tape_ctx *)d->ctx;
  if (c->step >= c->n_steps) {
    c->overrun = 1;
    return 0.0;
  }
  return c->tape[3 * (size_t)c->step + k];
}
static double tapesrc_gv(draw_src *d) { return tapesrc_get(d, 0); }
static double tapesrc_uv(draw_src *d) { return tapesrc_get(d, 1); }
static double tapesrc_gx(draw_src *d) { return tapesrc_get(d, 2); }
static void tapesrc_end(draw_src *d) { ((tape_ctx *)d->ctx)->step++; }

/* ========================================================================== */
/* SDE state, QE stepper, payoff policies                                     */
/* ========================================================================== */

#define PSI_C 1.5 /* src/inc/HSimulation.h:12 */

/* src/inc/SDE.h:9-14 + the stepper's private log_X (HSimulation.h:23) */
typedef struct {
  double cur_X, cur_V, prev_X, prev_V, cur_time, prev_time, log_X;
  double prev_log_X; /* NOT in the reference: log_X before the step (geometric control variate) */
} sde_state;

/* AsianContract.h:15-18 / VanillaContract.h:15-17 */
typedef struct {
  int payoff;
  double accumulated_value, final_value, earliest_unpriced_expi, init_step_size;
  /* NOT in the reference (which only suggests a control variate, src/inc/HSimulation.h:51): the
   * same accumulation applied to ln X instead of X -- the log of the geometric average with the
   * arithmetic average's own weights (include/hexo_gpu.h, HEXO_CV_GEOMETRIC) */
  double accumulated_log, final_log;
  double accumulated_weight; /* sum of the trapezoid weights applied so far: normalises final_log */
} policy;

/* AsianContract.h:35-38, VanillaContract.h:32-35 */
static void update_earliest(policy *o, double expiry, double steps) {
  o->init_step_size = expiry / steps;
  o->earliest_unpriced_expi = expiry;
}
/* AsianContract.h:39-42 (the European reset is empty, VanillaContract.h:36) */
static void policy_reset(policy *o) {
  if (o->payoff == ORACLE_ASIAN) {
    o->accumulated_value = 0.0;
    o->final_value = 0.0;
    o->accumulated_log = 0.0;
    o->final_log = 0.0;
    o->accumulated_weight = 0.0;
  }
}
/* AsianContract.h:25-28 (European: empty, VanillaContract.h:24-26) */
static void accumulate_value(policy *o, const sde_state *s) {
  if (o->payoff == ORACLE_ASIAN) {
    o->accumulated_value += o->init_step_size * .5 * (s->cur_X + s->prev_X);
    o->accumulated_log += o->init_step_size * .5 * (s->log_X + s->prev_log_X);
    o->accumulated_weight += o->init_step_size;
  }
}
/* AsianContract.h:29-34 / VanillaContract.h:28-31 */
static void accumulate_final_value(policy *o, const sde_state *s) {
  if (o->payoff == ORACLE_ASIAN) {
    double step_interpolation =
        (s->cur_X - s->prev_X) * (o->earliest_unpriced_expi - s->prev_time) / o->init_step_size;
    o->final_value = (o->accumulated_value + step_interpolation) / o->earliest_unpriced_expi;
    double log_interpolation = (s->log_X - s->prev_log_X) *
                               (o->earliest_unpriced_expi - s->prev_time) / o->init_step_size;
    /* the interpolation term has weights +x and -x: it does not change their sum */
    o->final_log = o->accumulated_weight > 0.0
                       ? (o->accumulated_log + log_interpolation) / o->accumulated_weight
                       : 0.0; /* one-step grid: no trapezoid was applied, G := 1 like the kernel */
  } else {
    o->final_value = s->prev_X + (s->cur_X - s->prev_X) *
                                     (o->earliest_unpriced_expi - s->prev_time) /
                                     o->init_step_size;
  }
}
/* AsianContract.h:43-47 / VanillaContract.h:37-39 */
static double final_payoff(const policy *o, double strike) {
  double v = o->final_value - strike;
  return v > 0.0 ? v : 0.0;
}

/* HQEAnderson::operator++, src/HSimulation.tpp:52-86 */
/* drift_mode 0 is the reference; 1 swaps K_0 for Andersen's K0* (oracle_qe_k0_star below) */
static void qe_step_drift(const oracle_hparams *p, sde_state *st, const policy *o, draw_src *d,
                          int drift_mode) {
  const double theta = p->v_m, rho = p->rho, kappa = p->kappa, eps = p->sigma; /* :54 */
  double delta = o->init_step_size;                                           /* :55 */
  double gamma_1 = .5, gamma_2 = .5;                                          /* :56-57 */
  double discount = exp(-kappa * delta);                                      /* :58 */
  double m = theta + (st->cur_V - theta) * discount;                          /* :59 */
  double sp2 = fabs(st->cur_V * eps * eps * discount / kappa * (1 - discount) +
                    theta * eps * eps / (2 * kappa) * (1. - discount) * (1. - discount)); /* :60 */
  double Psi = sp2 / (m * m);                                                 /* :61 */
  st->prev_V = st->cur_V;                                                     /* :62 */
  if (Psi < PSI_C) {                                                          /* :63 */
    double bp2 = 2 / Psi - 1 + sqrt(2 / Psi * (2 / Psi - 1));                 /* :64 */
    double b = sqrt(bp2);                                                     /* :65 */
    double a = m / (1 + bp2);                                                 /* :66 */
    double Z_V = d->variance_normal(d);                                       /* :67 */
    st->cur_V = a * (b + Z_V) * (b + Z_V);                                    /* :68 */
  } else {
    double pp = (Psi - 1) / (Psi + 1);                                        /* :70 */
    double beta = 2 / (m * (Psi + 1));                                        /* :71 */
    double U_V = d->variance_uniform(d);                                      /* :72 */
    st->cur_V = pp < U_V ? log((1 - pp) / (1 - U_V)) / beta : 0.;             /* :73 */
  }
  double K_0 = -rho * kappa * theta / eps * delta;                            /* :75 */
  double K_1 = gamma_1 * delta * (kappa * rho / eps - .5) - rho / eps;        /* :76 */
  double K_2 = gamma_2 * delta * (kappa * rho / eps - .5) + rho / eps;        /* :77 */
  double K_3 = gamma_1 * delta * (1 - rho * rho);                             /* :78 */
  double K_4 = gamma_2 * delta * (1 - rho * rho);                             /* :79 */
  if (drift_mode == 1) /* NOT the reference: Andersen's K0* in place of K0 */
    K_0 = oracle_qe_k0_star(p, delta, st->prev_V, NULL, NULL);
  st->prev_log_X = st->log_X;
  st->log_X = st->log_X + K_0 + K_1 * st->prev_V + K_2 * st->cur_V +
              sqrt(K_3 * st->prev_V + K_4 * st->cur_V) * d->spot_normal(d);   /* :80 */
  st->prev_X = st->cur_X;                                                     /* :81 */
  st->cur_X = exp(st->log_X);                                                 /* :82 */
  st->prev_time = st->cur_time;                                               /* :83 */
  st->cur_time += delta;                                                      /* :84 */
  d->end_step(d);
}

/* Andersen (2008), Proposition 9, written with the scheme's own quantities a, b^2, p, beta
 * (HSimulation.tpp:59-71); gamma_1 = gamma_2 = 1/2.  Not part of the reference. */
double oracle_qe_k0_star(const oracle_hparams *p, double h, double V, int *branch, int *corrected) {
  const double theta = p->v_m, rho = p->rho, kappa = p->kappa, eps = p->sigma;
  const double discount = exp(-kappa * h);
  const double m = theta + (V - theta) * discount;
  const double sp2 = fabs(V * eps * eps * discount / kappa * (1 - discount) +
                          theta * eps * eps / (2 * kappa) * (1. - discount) * (1. - discount));
  const double Psi = sp2 / (m * m);
  const double K_0 = -rho * kappa * theta / eps * h;
  const double K_1 = .5 * h * (kappa * rho / eps - .5) - rho / eps;
  const double K_2 = .5 * h * (kappa * rho / eps - .5) + rho / eps;
  const double K_3 = .5 * h * (1 - rho * rho), K_4 = K_3;
  const double A = K_2 + .5 * K_4;
  double lnM;
  int ok;
  if (Psi < PSI_C) {
    const double bp2 = 2 / Psi - 1 + sqrt(2 / Psi * (2 / Psi - 1));
    const double a = m / (1 + bp2);
    ok = 1 - 2 * A * a > 0;
    lnM = ok ? A * bp2 * a / (1 - 2 * A * a) - .5 * log(1 - 2 * A * a) : 0;
    if (branch) *branch = 0;
  } else {
    const double pp = (Psi - 1) / (Psi + 1);
    const double beta = 2 / (m * (Psi + 1));
    ok = beta > A;
    lnM = ok ? log(pp + beta * (1 - pp) / (beta - A)) : 0;
    if (branch) *branch = 1;
  }
  if (corrected) *corrected = ok;
  return ok ? -lnM - (K_1 + .5 * K_3) * V : K_0;
}

typedef void (*pay_fn)(void *ctx, uint32_t chain, const policy *o);

/* One path: the body of the `for i` loop, src/HSimulation.tpp:31-45.
 * `trailing_step` reproduces the for-loop's final `++heston_sde` (:35), which
 * runs one more stepper invocation (two more draws) after the last payment.
 * Returns the number of stepper invocations up to the last payment. */
static uint32_t simulate_path(const oracle_contract *c, draw_src *d, int trailing_step,
                              pay_fn pay, void *pay_ctx) {
  const uint32_t n_opts = c->strike_offsets[c->n_chains];
  uint32_t opts_priced = 0, chain = 0, n_steps = 0;                 /* :31-32 */
  policy o;
  memset(&o, 0, sizeof(o));
  o.payoff = c->payoff;
  update_earliest(&o, c->expiries[0], (double)c->steps);            /* :33 */
  policy_reset(&o);                                                 /* :34 */
  /* heston_sde = initial_state: HSimulation.tpp:26, :87-94 */
  sde_state st = {c->S, c->p.v_0, c->S, c->p.v_0, 0.0, 0.0, log(c->S), log(c->S)};
  if (n_opts == 0) return 0;
  qe_step_drift(&c->p, &st, &o, d, c->drift_mode);                  /* :35 ++(sde=init) */
  n_steps = 1;
  for (;;) {
    while (st.cur_time >= c->expiries[chain] && opts_priced < n_opts) { /* :36, SDE.h:30 */
      accumulate_final_value(&o, &st);                              /* :38 */
      pay(pay_ctx, chain, &o);                                      /* :39-40 */
      opts_priced = c->strike_offsets[chain + 1];                   /* :41 */
      if (opts_priced < n_opts) {
        ++chain;
        update_earliest(&o, c->expiries[chain], (double)c->steps);  /* :42 */
      }
    }
    accumulate_value(&o, &st);                                      /* :44 */
    if (opts_priced >= n_opts) {
      if (trailing_step) qe_step_drift(&c->p, &st, &o, d, c->drift_mode); /* :35 ++sde, exit */
      break;
    }
    qe_step_drift(&c->p, &st, &o, d, c->drift_mode);                /* :35 ++sde */
    ++n_steps;
  }
  return n_steps;
}

/* ---- payment sinks --------------------------------------------------------- */

typedef struct {
  const oracle_contract *c;
  double *prices, *sum, *sumsq;
  double n_sims;
  /* control variate c = final value - S (include/hexo_gpu.h, HEXO_CV_UNDERLYING); may be NULL */
  double *cross, *ctl, *ctl2;
  /* geometric != 0 (HEXO_CV_GEOMETRIC): the control is per option, c_j = max(G - K_j, 0) with
   * G = exp(final_log); cross / ctl / ctl2 then hold n_opts entries each */
  int geometric;
} sum_sink;

static void pay_sums(void *ctx, uint32_t chain, const policy *o) {
  sum_sink *k = (sum_sink *)ctx;
  if (k->geometric) {
    const double G = exp(o->final_log);
    for (uint32_t j = k->c->strike_offsets[chain]; j < k->c->strike_offsets[chain + 1]; ++j) {
      const double pf = final_payoff(o, k->c->strikes[j]);
      const double cg = G - k->c->strikes[j] > 0.0 ? G - k->c->strikes[j] : 0.0;
      k->sum[j] += pf;
      k->sumsq[j] += pf * pf;
      k->cross[j] += pf * cg;
      k->ctl[j] += cg;
      k->ctl2[j] += cg * cg;
    }
    return;
  }
  for (uint32_t j = k->c->strike_offsets[chain]; j < k->c->strike_offsets[chain + 1]; ++j) {
    double pf = final_payoff(o, k->c->strikes[j]);
    if (k->prices) k->prices[j] += pf / k->n_sims;                  /* :40 */
    if (k->sum) k->sum[j] += pf;
    if (k->sumsq) k->sumsq[j] += pf * pf;
    if (k->cross) k->cross[j] += pf * (o->final_value - k->c->S);
  }
  if (k->ctl) k->ctl[chain] += o->final_value - k->c->S;
  if (k->ctl2) k->ctl2[chain] += (o->final_value - k->c->S) * (o->final_value - k->c->S);
}

typedef struct {
  double *finals; /* [n_chains] for the current path */
} final_sink;
static void pay_finals(void *ctx, uint32_t chain, const policy *o) {
  ((final_sink *)ctx)->finals[chain] = o->final_value;
}

static int check_contract(const oracle_contract *c) {
  if (!c || c->n_chains == 0 || c->steps == 0) return -1;
  for (uint32_t k = 1; k < c->n_chains; ++k)
    if (!(c->expiries[k - 1] < c->expiries[k])) return -2;          /* :15-21 */
  if (!(c->expiries[0] > 0.0)) return -2;
  return 0;
}

/* ========================================================================== */
/* drivers                                                                    */
/* ========================================================================== */

int oracle_price_ref(const oracle_contract *c, unsigned int n_sims, unsigned int nthreads,
                     size_t rand_buf_size, int normal_mode, double *prices, double *sum,
                     double *sumsq) {
  int rc = check_contract(c);
  if (rc) return rc;
  if (nthreads == 0) return -1;
  const uint32_t n_opts = c->strike_offsets[c->n_chains];
  if (prices) memset(prices, 0, n_opts * sizeof(double));           /* :22 */
  if (sum) memset(sum, 0, n_opts * sizeof(double));
  if (sumsq) memset(sumsq, 0, n_opts * sizeof(double));
  sum_sink sink = {c, prices, sum, sumsq, (double)n_sims, NULL, NULL, NULL, 0};
  for (unsigned int tid = 0; tid < nthreads; ++tid) {               /* :23 */
    const unsigned int local_sims = n_sims / nthreads;              /* :27 */
    oracle_rng *rng = oracle_rng_new(rand_buf_size, 1u << tid, normal_mode); /* :28 */
    if (!rng) return -3;
    draw_src d = {refsrc_g, refsrc_u, refsrc_g, refsrc_end, rng};
    for (unsigned int i = 0; i < local_sims; ++i)                   /* :30 */
      simulate_path(c, &d, 1, pay_sums, &sink);
    oracle_rng_free(rng);
  }
  return 0;
}

int oracle_price_stream(const oracle_contract *c, uint64_t seed, uint64_t n_paths,
                        uint64_t n_streams_total, uint64_t stream_begin, uint64_t stream_count,
                        int normal_mode, double *sum, double *sumsq) {
  return oracle_price_stream_rng(c, 0, seed, n_paths, n_streams_total, stream_begin, stream_count,
                                 normal_mode, sum, sumsq);
}

int oracle_price_stream_rng(const oracle_contract *c, int rng_mode, uint64_t seed, uint64_t n_paths,
                            uint64_t n_streams_total, uint64_t stream_begin, uint64_t stream_count,
                            int normal_mode, double *sum, double *sumsq) {
  int rc = check_contract(c);
  if (rc) return rc;
  if (n_streams_total == 0 || stream_begin + stream_count > n_streams_total) return -1;
  const uint32_t n_opts = c->strike_offsets[c->n_chains];
  if (sum) memset(sum, 0, n_opts * sizeof(double));
  if (sumsq) memset(sumsq, 0, n_opts * sizeof(double));
  sum_sink sink = {c, NULL, sum, sumsq, (double)n_paths, NULL, NULL, NULL, 0};
  const uint64_t base = n_paths / n_streams_total, rem = n_paths % n_streams_total;
  for (uint64_t s = stream_begin; s < stream_begin + stream_count; ++s) {
    stream_ctx sc;
    memset(&sc, 0, sizeof(sc));
    uint64_t sd[4] = {seed, s, 0, 0};
    prng_init(&sc.s, sd);
    sc.pos = 16;
    sc.normal_mode = normal_mode;
    sc.rng_mode = rng_mode;
    sc.seed = seed;
    sc.stream = s;
    sc.n_step = 0;
    draw_src d = {strsrc_gv, strsrc_uv, strsrc_gx, strsrc_end, &sc};
    const uint64_t my_paths = base + (s < rem ? 1 : 0);
    for (uint64_t i = 0; i < my_paths; ++i) simulate_path(c, &d, 0, pay_sums, &sink);
  }
  return 0;
}

/* ---- corrected time grid (NOT the reference's loop; include/hexo_gpu.h, HEXO_SCHEDULE_EXACT) --
 * Segment k covers (T_{k-1}, T_k] with n_k = max(1, round((T_k - T_{k-1}) steps / T_k)) steps
 * of equal width; the Asian average is the full trapezoid rule over [0, T_k], the European
 * payoff reads X at the last step.  The stepper itself is the reference's (qe_step). */
static uint32_t simulate_path_exact(const oracle_contract *c, draw_src *d, pay_fn pay,
                                    void *pay_ctx) {
  policy o;
  memset(&o, 0, sizeof(o));
  o.payoff = c->payoff;
  sde_state st = {c->S, c->p.v_0, c->S, c->p.v_0, 0.0, 0.0, log(c->S), log(c->S)};
  double integral = 0.0, integral_log = 0.0, t_prev = 0.0;
  uint32_t total = 0;
  for (uint32_t k = 0; k < c->n_chains; ++k) {
    const double span = c->expiries[k] - t_prev;
    long long n = llround(span * (double)c->steps / c->expiries[k]);
    if (n < 1) n = 1;
    o.init_step_size = span / (double)n; /* qe_step reads its step width here */
    o.earliest_unpriced_expi = c->expiries[k];
    for (long long j = 0; j < n; ++j) {
      qe_step_drift(&c->p, &st, &o, d, c->drift_mode);
      integral += o.init_step_size * .5 * (st.cur_X + st.prev_X);
      integral_log += o.init_step_size * .5 * (st.log_X + st.prev_log_X);
      ++total;
    }
    o.final_value = c->payoff == ORACLE_ASIAN ? integral / c->expiries[k] : st.cur_X;
    o.final_log = integral_log / c->expiries[k]; /* full trapezoid rule: the weights sum to T_k */
    pay(pay_ctx, k, &o);
    t_prev = c->expiries[k];
  }
  return total;
}

int oracle_price_stream_exact(const oracle_contract *c, int rng_mode, uint64_t seed,
                              uint64_t n_paths, uint64_t n_streams_total, uint64_t stream_begin,
                              uint64_t stream_count, int normal_mode, double *sum, double *sumsq) {
  int rc = check_contract(c);
  if (rc) return rc;
  if (n_streams_total == 0 || stream_begin + stream_count > n_streams_total) return -1;
  const uint32_t n_opts = c->strike_offsets[c->n_chains];
  if (sum) memset(sum, 0, n_opts * sizeof(double));
  if (sumsq) memset(sumsq, 0, n_opts * sizeof(double));
  sum_sink sink = {c, NULL, sum, sumsq, (double)n_paths, NULL, NULL, NULL, 0};
  const uint64_t base = n_paths / n_streams_total, rem = n_paths % n_streams_total;
  for (uint64_t s = stream_begin; s < stream_begin + stream_count; ++s) {
    stream_ctx sc;
    memset(&sc, 0, sizeof(sc));
    uint64_t sd[4] = {seed, s, 0, 0};
    prng_init(&sc.s, sd);
    sc.pos = 16;
    sc.normal_mode = normal_mode;
    sc.rng_mode = rng_mode;
    sc.seed = seed;
    sc.stream = s;
    draw_src d = {strsrc_gv, strsrc_uv, strsrc_gx, strsrc_end, &sc};
    const uint64_t my_paths = base + (s < rem ? 1 : 0);
    for (uint64_t i = 0; i < my_paths; ++i) simulate_path_exact(c, &d, pay_sums, &sink);
  }
  return 0;
}

/* Stream pricer with the control-variate sums (include/hexo_gpu.h, HEXO_CV_UNDERLYING):
 * out = [sum pf | sum pf^2 | sum pf c] per option, then [sum c | sum c^2] per maturity,
 * c = final value - S.  exact_grid selects the corrected time grid. */
int oracle_price_stream_cv(const oracle_contract *c, int rng_mode, int exact_grid, uint64_t seed,
                           uint64_t n_paths, uint64_t n_streams_total, uint64_t stream_begin,
                           uint64_t stream_count, int normal_mode, double *out) {
  int rc = check_contract(c);
  if (rc) return rc;
  if (n_streams_total == 0 || stream_begin + stream_count > n_streams_total || !out) return -1;
  const uint32_t n_opts = c->strike_offsets[c->n_chains];
  memset(out, 0, (3 * (size_t)n_opts + 2 * (size_t)c->n_chains) * sizeof(double));
  sum_sink sink = {c, NULL, out, out + n_opts, (double)n_paths, out + 2 * (size_t)n_opts,
                   out + 3 * (size_t)n_opts, out + 3 * (size_t)n_opts + c->n_chains, 0};
  const uint64_t base = n_paths / n_streams_total, rem = n_paths % n_streams_total;
  for (uint64_t s = stream_begin; s < stream_begin + stream_count; ++s) {
    stream_ctx sc;
    memset(&sc, 0, sizeof(sc));
    uint64_t sd[4] = {seed, s, 0, 0};
    prng_init(&sc.s, sd);
    sc.pos = 16;
    sc.normal_mode = normal_mode;
    sc.rng_mode = rng_mode;
    sc.seed = seed;
    sc.stream = s;
    draw_src d = {strsrc_gv, strsrc_uv, strsrc_gx, strsrc_end, &sc};
    const uint64_t my_paths = base + (s < rem ? 1 : 0);
    for (uint64_t i = 0; i < my_paths; ++i) {
      if (exact_grid)
        simulate_path_exact(c, &d, pay_sums, &sink);
      else
        simulate_path(c, &d, 0, pay_sums, &sink);
    }
  }
  return 0;
}

/* Stream pricer with the sums of the geometric-Asian control variate (include/hexo_gpu.h,
 * HEXO_CV_GEOMETRIC; NOT in the reference): out = [sum pf | sum pf^2 | sum pf c | sum c |
 * sum c^2], n_opts entries each, c_j = max(G - K_j, 0), G = exp of the policy's accumulation
 * applied to ln X.  Asian contracts only. */
int oracle_price_stream_geo(const oracle_contract *c, int rng_mode, int exact_grid, uint64_t seed,
                            uint64_t n_paths, uint64_t n_streams_total, uint64_t stream_begin,
                            uint64_t stream_count, int normal_mode, double *out) {
  int rc = check_contract(c);
  if (rc) return rc;
  if (c->payoff != ORACLE_ASIAN) return -1;
  if (n_streams_total == 0 || stream_begin + stream_count > n_streams_total || !out) return -1;
  const size_t n_opts = c->strike_offsets[c->n_chains];
  memset(out, 0, 5 * n_opts * sizeof(double));
  sum_sink sink = {c, NULL, out, out + n_opts, (double)n_paths, out + 2 * n_opts,
                   out + 3 * n_opts, out + 4 * n_opts, 1};
  const uint64_t base = n_paths / n_streams_total, rem = n_paths % n_streams_total;
  for (uint64_t s = stream_begin; s < stream_begin + stream_count; ++s) {
    stream_ctx sc;
    memset(&sc, 0, sizeof(sc));
    uint64_t sd[4] = {seed, s, 0, 0};
    prng_init(&sc.s, sd);
    sc.pos = 16;
    sc.normal_mode = normal_mode;
    sc.rng_mode = rng_mode;
    sc.seed = seed;
    sc.stream = s;
    draw_src d = {strsrc_gv, strsrc_uv, strsrc_gx, strsrc_end, &sc};
    const uint64_t my_paths = base + (s < rem ? 1 : 0);
    for (uint64_t i = 0; i < my_paths; ++i) {
      if (exact_grid)
        simulate_path_exact(c, &d, pay_sums, &sink);
      else
        simulate_path(c, &d, 0, pay_sums, &sink);
    }
  }
  return 0;
}

int oracle_replay(const oracle_contract *c, const double *tape, uint64_t n_paths,
                  uint32_t tape_steps, double *finals) {
  int rc = check_contract(c);
  if (rc) return rc;
  uint32_t n_steps = 0;
  for (uint64_t i = 0; i < n_paths; ++i) {
    tape_ctx tc = {tape + 3 * (size_t)tape_steps * i, 0, tape_steps, 0};
    draw_src d = {tapesrc_gv, tapesrc_uv, tapesrc_gx, tapesrc_end, &tc};
    final_sink sink = {finals + (size_t)c->n_chains * i};
    n_steps = simulate_path(c, &d, 0, pay_finals, &sink);
    if (tc.overrun) return -1;
  }
  return (int)n_steps;
}

static double zero_draw(draw_src *d) {
  (void)d;
  return 0.0;
}
static void nop_pay(void *ctx, uint32_t chain, const policy *o) {
  (void)ctx;
  (void)chain;
  (void)o;
}

uint32_t oracle_steps_to_last_expiry(const oracle_contract *c) {
  if (check_contract(c)) return 0;
  draw_src d = {zero_draw, zero_draw, zero_draw, refsrc_end, NULL};
  return simulate_path(c, &d, 0, nop_pay, NULL);
}
