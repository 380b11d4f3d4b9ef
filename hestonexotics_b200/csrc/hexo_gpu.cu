// hestonexotics_b200/csrc/hexo_gpu.cu
//
// B200 (sm_100a) Heston Monte-Carlo hot path behind the C ABI of
// include/hexo_gpu.h.  One fused kernel replaces the reference's
// HSimulation::price<Scheme> (src/HSimulation.tpp:10-51) together with
// everything it calls per draw: the shishua wrapper (src/RNG.cpp), PPND16
// (src/as241.f90), the QE stepper (HSimulation.tpp:52-86) and the payoff
// policies (src/inc/AsianContract.h, src/inc/VanillaContract.h).
//
// Design:
//  * one thread = one independent shishua stream; RNG state, variance, log-spot
//    and the running integral stay in registers; no path data touches HBM;
//  * a generator round (16 words) is parked in a per-thread column of shared
//    memory so the step loop can consume two words per step without being
//    unrolled around the generator;
//  * when a maturity is reached the 32 final values of a warp are exchanged
//    through shared memory and each lane owns a strided subset of the strikes,
//    so per-option sums are accumulated without atomics and in a fixed order;
//  * warps -> block partials (shared memory), blocks -> sums (second tiny
//    kernel), both in fixed order: results are reproducible for a given launch
//    geometry.
// There is deliberately no CPU fallback in this file.

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/hexo_gpu.h"
#include "host_internal.h"
#include "path_kernels.h"

namespace hexo {

// ---------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define HEXO_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver             \
                      ? HEXO_ERR_NO_DEVICE                                                 \
                      : HEXO_ERR_CUDA,                                                     \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// ---------------------------------------------------------------------------
// device contexts: one per visible GPU, set up on first use.  A process normally drives one
// GPU (one rank per GPU); hexo_gpu_price_multi drives several from one process, and each of
// them needs its own SM count, shared-memory limit and memory-pool setting.
// ---------------------------------------------------------------------------
struct Context {
  bool ready = false;
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
};
constexpr int kMaxDevices = 64;
static Context g_ctxs[kMaxDevices];
static thread_local Context* g_cur = nullptr;  // context of the calling thread's current device
#define g_ctx (*g_cur)

// points g_cur at the context of the current device, initialising it if need be
static int ensure_context() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(HEXO_ERR_NO_DEVICE, "no CUDA device: %s",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  int dev = 0;
  HEXO_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "device %d: at most %d devices", dev, kMaxDevices);
  g_cur = &g_ctxs[dev];
  if (g_ctx.ready) return HEXO_OK;
  // single attributes, not cudaGetDeviceProperties: that call gathers every property of the
  // device and was measured at 10-380 ms per call while other processes keep their GPUs busy
  int sms = 0, smem = 0;
  HEXO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  HEXO_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // keep freed blocks of the stream-ordered allocator in the pool: a pricing call allocates and
  // frees one small block, and returning it to the driver at every synchronisation makes the
  // next call pay a real allocation (slow once peer access is enabled, as it is under NCCL)
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  g_ctx.device = dev;
  g_ctx.sm_count = sms;
  g_ctx.smem_optin = (size_t)smem;
  g_ctx.ready = true;
  return HEXO_OK;
}

// blocks -> sums, fixed order
__global__ void reduce_partials_kernel(const double* __restrict__ partials, uint32_t n_blocks,
                                       uint32_t n2, double* __restrict__ out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n2) return;
  double s = 0.0;
  for (uint32_t b = 0; b < n_blocks; ++b) s += partials[(size_t)b * n2 + j];
  out[j] = s;
}

// ---------------------------------------------------------------------------
// K2: raw shishua bytes
// ---------------------------------------------------------------------------
__global__ void shishua_streams_kernel(uint64_t seed0, uint64_t seed1_first, uint64_t seed2,
                                       uint64_t seed3, uint32_t n_streams, uint64_t* out,
                                       size_t words_per_stream) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_streams) return;
  Shishua rng;
  uint64_t o[16];
  rng.init(seed0, seed1_first + i, seed2, seed3, o);
  uint64_t* dst = out + (size_t)i * words_per_stream;
  for (size_t off = 0; off < words_per_stream; off += 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[off + j] = o[j];
    rng.round(o);
  }
}

// K2b: Philox4x32-10 blocks (known-answer tests) and the words a stream hands to the stepper
__global__ void philox_kernel(const uint32_t* ctr, const uint32_t* key, uint32_t* out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t o[4];
  philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i],
                key[2 * i + 1], o);
#pragma unroll
  for (int j = 0; j < 4; ++j) out[4 * i + j] = o[j];
}
__global__ void philox_streams_kernel(uint64_t seed, uint64_t first_stream, uint32_t n_streams,
                                      uint64_t* out, size_t words_per_stream) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_streams) return;
  PhiloxGen rng;
  uint64_t o[16];
  rng.init(seed, first_stream + i, 0, 0, o);
  uint64_t* dst = out + (size_t)i * words_per_stream;
  for (size_t off = 0; off < words_per_stream; off += 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[off + j] = o[j];
    rng.round(o);
  }
}

// ---------------------------------------------------------------------------
// K3: uniform map, inverse normal
// ---------------------------------------------------------------------------
__global__ void u64_to_unit_kernel(const uint64_t* in, double* out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = u64_to_unit(in[i]);
}
template <int NORMAL_MODE>
__global__ void ppnd16_kernel(const double* in, double* out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ppnd16<NORMAL_MODE>(in[i]);
}

// K3b: the normal transform EXACTLY as the path kernel runs it.  ppnd16_kernel above evaluates
// the scalar AS241 routine; K1 does not call that one but ring_refill (path_kernel.cuh): batched
// central phase (two draws per FFMA2, q from the high word), then the warp-cooperative tail
// phase through the shared-memory ring.  This kernel feeds caller-supplied words through
// ring_refill in place of generator rounds and reads the normals back from the z ring the way
// the step loop does.  Thread c transforms words [c W, (c + 1) W), W = 2 ring_steps; word 2 s of
// a chunk is the variance word of step s, word 2 s + 1 its spot word.
struct WordSource {
  const uint64_t* p;  // nullptr: central filler words (threads past the end of the input)
  __device__ __forceinline__ void round(uint64_t (&o)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = p ? p[j] : 0x8000000000000000ull;
    if (p) p += 16;
  }
};
template <int NORMAL_MODE>
__global__ void __launch_bounds__(kMaxBlock)
normals_from_words_kernel(const uint64_t* __restrict__ words, double* __restrict__ z,
                          uint64_t n_chunks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Ring = ZRing<NORMAL_MODE>;
  constexpr int kRing = ring_steps(NORMAL_MODE), kWords = 2 * kRing;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const RingAddr ra = ring_addr<NORMAL_MODE>(
      smem_addr(smem_raw), tid, T,
      smem_addr(smem_raw) + (uint32_t)ring_smem(T, NORMAL_MODE) + kTailListBytes * warp);
  ring_logtab_init<NORMAL_MODE>(smem_raw, tid, T);
  __syncthreads();
  const uint64_t chunk = (uint64_t)blockIdx.x * T + tid;
  WordSource src{chunk < n_chunks ? words + chunk * kWords : nullptr};
  uint64_t o[16];
  ring_refill<NORMAL_MODE>(src, o, false, ra, lane);
  if (chunk >= n_chunks) return;
  for (int s = 0; s < kRing; ++s) {
    double zv, zx;
    Ring::get(ra.zcol + s * ra.zstride, zv, zx);
    z[chunk * kWords + 2 * s] = zv;
    z[chunk * kWords + 2 * s + 1] = zx;
  }
}

// ---------------------------------------------------------------------------
// K4: tape replay (one thread per path), same stepper and policy arithmetic
// ---------------------------------------------------------------------------
template <int PAYOFF, bool MART>
__global__ void qe_replay_kernel(double v0, double S, double lnS, uint32_t n_seg,
                                 const SegConst* segs, const double* tape, uint64_t n_paths,
                                 uint32_t tape_steps, double* finals) {
  __shared__ double exptab_mem[32];
  exp_table_init(exptab_mem, threadIdx.x, blockDim.x);
  __syncthreads();
  const uint32_t exptab = smem_addr(exptab_mem);
  const uint64_t path = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (path >= n_paths) return;
  const double* t = tape + (size_t)path * tape_steps * 3;
  double V = v0, lnX = lnS, X = S, Xprev = S, integral = 0.0;
  uint32_t step = 0;
  for (uint32_t k = 0; k < n_seg; ++k) {
    const SegConst g = segs[k];
    const uint32_t n = g.n_steps;
    if (PAYOFF == HEXO_PAYOFF_ASIAN && k > 0 && n > 0) integral += g.hcarry * (X + Xprev);
    const double Xa = X;
    double sumX = 0.0;
    for (uint32_t i = 0; i < n; ++i, ++step) {
      const double uv = t[3 * step + 1];
      double k0 = 0.0;
      const double Vn = qe_variance<MART>(g, V, t[3 * step], [uv]() { return uv; }, &k0);
      const double delta = MART ? qe_logreturn_mart(g, V, Vn, t[3 * step + 2], k0)
                                : qe_logreturn(g, V, Vn, t[3 * step + 2]);
      V = Vn;
      if (PAYOFF == HEXO_PAYOFF_ASIAN) {  // same arithmetic as the path kernel
        Xprev = X;
        X = grow_spot(X, delta, exptab);
        if (i + 1 < n) sumX += X;
      } else {
        lnX += delta;
        if (i + 2 >= n) {
          Xprev = X;
          X = fast_exp(lnX, exptab);
        }
      }
    }
    if (PAYOFF == HEXO_PAYOFF_ASIAN && n > 0) integral += g.h * 0.5 * (Xa - Xprev + 2.0 * sumX);
    const double dx = X - Xprev;
    finals[path * n_seg + k] =
        PAYOFF == HEXO_PAYOFF_ASIAN ? (integral + dx * g.w + g.hs * (X + Xprev)) / g.expiry
                                    : Xprev + dx * g.w;
  }
}

// ---------------------------------------------------------------------------
// FP64 pipe peak: 8 independent DFMA chains per thread, registers only
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ---------------------------------------------------------------------------
// host: schedule and segment constants
// ---------------------------------------------------------------------------

// Replays the reference's time bookkeeping in plain double arithmetic:
// cur_time += delta per stepper call (HSimulation.tpp:83-84), delta =
// expiry_k/steps while chain k is the earliest unpriced one
// (AsianContract.h:35-38), payment when cur_time >= expiry_k
// (HSimulation.tpp:36, SDE.h:30), possibly several chains on one step.
static int build_schedule(const double* expiries, uint32_t n_chains, uint32_t steps,
                          hexo_segment* seg) {
  if (!expiries || !seg || n_chains == 0 || steps == 0)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "schedule: need expiries, n_chains>0, steps>0");
  if (!(expiries[0] > 0.0) || !std::isfinite(expiries[0]))
    return fail(HEXO_ERR_NOT_INCREASING, "first expiry must be positive and finite");
  for (uint32_t k = 1; k < n_chains; ++k)
    if (!(expiries[k - 1] < expiries[k]) || !std::isfinite(expiries[k]))
      return fail(HEXO_ERR_NOT_INCREASING, "expiries must be strictly increasing (chain %u)", k);
  volatile double cur_time = 0.0, prev_time = 0.0;  // volatile: no x87/FMA surprises
  uint32_t k = 0, since_last = 0;
  double h = expiries[0] / (double)steps;
  // ++(sde = initial_state): the first step is taken before any check
  prev_time = cur_time;
  cur_time = cur_time + h;
  since_last = 1;
  for (;;) {
    while (k < n_chains && cur_time >= expiries[k]) {
      seg[k].n_steps = since_last;
      seg[k].h = h;
      seg[k].w = (expiries[k] - prev_time) / h;
      seg[k].expiry = expiries[k];
      since_last = 0;
      if (++k < n_chains) h = expiries[k] / (double)steps;
    }
    if (k >= n_chains) break;
    prev_time = cur_time;
    cur_time = cur_time + h;
    ++since_last;
  }
  return HEXO_OK;
}

// HEXO_SCHEDULE_EXACT: a time grid that ends on every expiry.  Segment k covers
// (T_{k-1}, T_k] with n_k = max(1, round((T_k - T_{k-1}) steps / T_k)) steps of width
// (T_k - T_{k-1}) / n_k -- `steps` keeps the reference's meaning "steps per the expiry currently
// headed for" (AsianContract.h:35-38) -- and the last step counts fully (w = 1).
static int build_schedule_exact(const double* expiries, uint32_t n_chains, uint32_t steps,
                                hexo_segment* seg) {
  if (!expiries || !seg || n_chains == 0 || steps == 0)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "schedule: need expiries, n_chains>0, steps>0");
  if (!(expiries[0] > 0.0) || !std::isfinite(expiries[0]))
    return fail(HEXO_ERR_NOT_INCREASING, "first expiry must be positive and finite");
  double prev = 0.0;
  for (uint32_t k = 0; k < n_chains; ++k) {
    if (k > 0 && (!(expiries[k - 1] < expiries[k]) || !std::isfinite(expiries[k])))
      return fail(HEXO_ERR_NOT_INCREASING, "expiries must be strictly increasing (chain %u)", k);
    const double span = expiries[k] - prev;
    long long n = llround(span * (double)steps / expiries[k]);
    if (n < 1) n = 1;
    seg[k].n_steps = (uint32_t)n;
    seg[k].h = span / (double)n;
    seg[k].w = 1.0;
    seg[k].expiry = expiries[k];
    prev = expiries[k];
  }
  return HEXO_OK;
}

static int check_request(const hexo_price_request* r, bool need_strikes) {
  if (!r) return fail(HEXO_ERR_INVALID_ARGUMENT, "request is NULL");
  if (r->n_chains == 0 || !r->expiries) return fail(HEXO_ERR_INVALID_ARGUMENT, "no chains");
  if (r->steps == 0) return fail(HEXO_ERR_INVALID_ARGUMENT, "steps must be > 0");
  if (r->payoff != HEXO_PAYOFF_ASIAN && r->payoff != HEXO_PAYOFF_EUROPEAN)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "unknown payoff %d", r->payoff);
  if (r->normal_mode != HEXO_NORMAL_F32 && r->normal_mode != HEXO_NORMAL_F64 &&
      r->normal_mode != HEXO_NORMAL_F32_PPND7)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "unknown normal_mode %d", r->normal_mode);
  if (r->normal_mode == HEXO_NORMAL_F32_PPND7 &&
      (r->rng_mode != HEXO_RNG_SHISHUA || r->control_variate != HEXO_CV_NONE ||
       r->drift_mode != HEXO_DRIFT_REFERENCE))
    return fail(HEXO_ERR_INVALID_ARGUMENT,
                "HEXO_NORMAL_F32_PPND7 is built for the shishua generator, the reference drift "
                "and plain sums only");
  if (r->rng_mode != HEXO_RNG_SHISHUA && r->rng_mode != HEXO_RNG_PHILOX)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "unknown rng_mode %d", r->rng_mode);
  if (r->schedule_mode != HEXO_SCHEDULE_REFERENCE && r->schedule_mode != HEXO_SCHEDULE_EXACT)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "unknown schedule_mode %d", r->schedule_mode);
  if (r->control_variate != HEXO_CV_NONE && r->control_variate != HEXO_CV_UNDERLYING &&
      r->control_variate != HEXO_CV_GEOMETRIC)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "unknown control_variate %d", r->control_variate);
  if (r->control_variate == HEXO_CV_GEOMETRIC &&
      (r->payoff != HEXO_PAYOFF_ASIAN || r->rng_mode != HEXO_RNG_SHISHUA))
    return fail(HEXO_ERR_INVALID_ARGUMENT,
                "HEXO_CV_GEOMETRIC is the control of the Asian payoff (shishua generator)");
  if (r->drift_mode != HEXO_DRIFT_REFERENCE && r->drift_mode != HEXO_DRIFT_MARTINGALE)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "unknown drift_mode %d", r->drift_mode);
  // The reference divides by kappa and sigma (HSimulation.tpp:60,75-77) and takes log(S) (:90);
  // it would silently produce NaN prices.  Refuse instead.
  const hexo_hparams& p = r->p;
  if (!(std::isfinite(r->S) && r->S > 0.0))
    return fail(HEXO_ERR_INVALID_ARGUMENT, "spot S must be positive and finite");
  if (!(std::isfinite(p.v_0) && p.v_0 >= 0.0 && std::isfinite(p.v_m) && p.v_m >= 0.0))
    return fail(HEXO_ERR_INVALID_ARGUMENT, "v_0 and v_m must be finite and non-negative");
  if (!(std::isfinite(p.kappa) && p.kappa > 0.0 && std::isfinite(p.sigma) && p.sigma > 0.0))
    return fail(HEXO_ERR_INVALID_ARGUMENT, "kappa and sigma must be positive and finite");
  if (!(p.rho >= -1.0 && p.rho <= 1.0))
    return fail(HEXO_ERR_INVALID_ARGUMENT, "rho must lie in [-1, 1]");
  if (need_strikes) {
    if (!r->strike_offsets || !r->strikes)
      return fail(HEXO_ERR_INVALID_ARGUMENT, "strike_offsets / strikes is NULL");
    if (r->strike_offsets[0] != 0) return fail(HEXO_ERR_INVALID_ARGUMENT, "strike_offsets[0] != 0");
    for (uint32_t k = 0; k < r->n_chains; ++k)
      if (r->strike_offsets[k + 1] < r->strike_offsets[k])
        return fail(HEXO_ERR_INVALID_ARGUMENT, "strike_offsets must be non-decreasing");
    if (r->strike_offsets[r->n_chains] == 0) return fail(HEXO_ERR_INVALID_ARGUMENT, "no options");
    if (r->n_paths == 0) return fail(HEXO_ERR_INVALID_ARGUMENT, "n_paths must be > 0");
    for (uint32_t j = 0; j < r->strike_offsets[r->n_chains]; ++j)
      if (!std::isfinite(r->strikes[j]))
        return fail(HEXO_ERR_INVALID_ARGUMENT, "strike %u is not finite", j);
  }
  return HEXO_OK;
}

static int build_segments(const hexo_price_request* r, bool with_strikes,
                          std::vector<SegConst>& out, uint64_t* steps_per_path) {
  std::vector<hexo_segment> sched(r->n_chains);
  const bool exact = r->schedule_mode == HEXO_SCHEDULE_EXACT;
  int rc = exact ? build_schedule_exact(r->expiries, r->n_chains, r->steps, sched.data())
                 : build_schedule(r->expiries, r->n_chains, r->steps, sched.data());
  if (rc) return rc;
  const double theta = r->p.v_m, rho = r->p.rho, kappa = r->p.kappa, eps = r->p.sigma;
  out.resize(r->n_chains);
  uint64_t total = 0;
  for (uint32_t k = 0; k < r->n_chains; ++k) {
    SegConst& g = out[k];
    const double h = sched[k].h;
    g.h = h;
    g.w = sched[k].w;
    g.hcarry = h * 0.5;  // HSimulation.tpp:42-44: the crossing trapezoid takes the NEW step width
    g.hs = 0.0;
    if (exact) {
      g.hcarry = k > 0 ? sched[k - 1].h * 0.5 : 0.0;
      if (r->payoff == HEXO_PAYOFF_ASIAN) {  // full trapezoid of the last step instead of dx * w
        g.w = 0.0;
        g.hs = h * 0.5;
      }
    }
    g.expiry = sched[k].expiry;
    g.n_steps = sched[k].n_steps;
    total += g.n_steps;
    const double D = exp(-kappa * h);                                   // HSimulation.tpp:58
    g.D = D;
    g.m0 = theta * (1 - D);                                             // :59
    g.c1h = 0.5 * (eps * eps * D / kappa * (1 - D));                    // :60 (halved)
    g.c2h = 0.5 * (theta * eps * eps / (2 * kappa) * (1. - D) * (1. - D));
    g.K0 = -rho * kappa * theta / eps * h;                              // :75
    g.K1 = .5 * h * (kappa * rho / eps - .5) - rho / eps;               // :76
    g.K2 = .5 * h * (kappa * rho / eps - .5) + rho / eps;               // :77
    g.K3 = .5 * h * (1 - rho * rho);                                    // :78 (= K4, :79)
    g.A = g.K2 + .5 * g.K3;  // HEXO_DRIFT_MARTINGALE (Andersen 2008, Prop. 9)
    g.A2 = 2. * g.A;
    g.K1m = -.5 * g.K3;
    // The division-free variance step (qe.cuh) forms a = m - sqrt(m^2 - s^2/2) by subtraction:
    // its relative rounding error is ~4e-16 / psi, psi ~ sigma^2 h / V, and it reaches the
    // log-spot through rho/sigma (K1, K2).  Per step that is about
    // |rho| 2.8e-16 V^1.5 / (sigma^2 sqrt(h)): 5e-17 for the BASELINE parameters, 1e-12 at
    // sigma = 0.01 -- and O(1) at sigma = 1e-8, where the reference's a = m / (1 + b^2) is still
    // fine.  Refuse the (near-deterministic-variance) corner instead of pricing it inaccurately.
    {
      const double vmax = std::max(r->p.v_0, theta);
      const double est = fabs(rho) * 2.8e-16 * vmax * sqrt(vmax) / (eps * eps * sqrt(h));
      if (est > 1e-9)
        return fail(HEXO_ERR_INVALID_ARGUMENT,
                    "sigma = %g is too small for this step width (h = %g): the division-free QE "
                    "step would lose ~%.1e per step in the log-spot; sigma >= %.2e is supported here",
                    eps, h, est, eps * sqrt(est / 1e-9));
    }
    g.first_opt = with_strikes ? r->strike_offsets[k] : 0;
    g.n_strikes = with_strikes ? r->strike_offsets[k + 1] - r->strike_offsets[k] : 0;
    g.pad = 0;
  }
  {  // sum of the Asian policy's trapezoid weights per expiry (see control_means)
    double weight = 0.0;
    for (uint32_t k = 0; k < r->n_chains; ++k) {
      SegConst& g = out[k];
      if (g.n_steps > 0) {
        if (k > 0) weight += 2.0 * g.hcarry;
        weight += g.h * (double)(g.n_steps - 1);
      }
      const double w = weight + 2.0 * g.hs;
      g.inv_logw = w > 0.0 ? 1.0 / w : 0.0;
    }
  }
  if (steps_per_path) *steps_per_path = total;
  return HEXO_OK;
}

// sums per request: [sum pf | sum pf^2] per option, plus with the control variate
// [sum pf c] per option and [sum c | sum c^2] per maturity
static size_t sums_len(const hexo_price_request* r) {
  const size_t n_opts = r->strike_offsets[r->n_chains];
  if (r->control_variate == HEXO_CV_GEOMETRIC) return 5 * n_opts;
  return r->control_variate ? 3 * n_opts + 2 * (size_t)r->n_chains : 2 * n_opts;
}

// Known mean of the control c = final value - S of every maturity.  The spot is a martingale
// (r = 0; exactly under HEXO_DRIFT_MARTINGALE, up to the QE drift error otherwise), so
// E[X_j] = S on every grid point and E[final value] = S x (sum of the weights the policy
// actually applies) / T.  European: the interpolated X_T, weights sum to 1, E[c] = 0.  Asian:
// the trapezoid weights of the schedule.  On the reference grid they do NOT sum to T when the
// steps land exactly on the expiry (SURVEY finding 6: the last trapezoid is replaced by
// (X_N - X_{N-1}) w, mean zero), e.g. E[average] = S (1 - 1/steps) for a power-of-two step count.
static void control_means(const hexo_price_request* r, const std::vector<SegConst>& segs,
                          std::vector<double>& ec) {
  ec.assign(r->n_chains, 0.0);
  if (r->payoff != HEXO_PAYOFF_ASIAN) return;
  for (uint32_t k = 0; k < r->n_chains; ++k) {
    const SegConst& g = segs[k];  // inv_logw = 1 / (sum of the weights), build_segments
    if (!(g.inv_logw > 0.0)) continue;
    ec[k] = r->S * (1.0 / (g.inv_logw * g.expiry) - 1.0);  // dx * w has mean zero
  }
}

// sums -> prices and standard errors.  Plain: mean payoff (HSimulation.tpp:40 divides by
// n_simulations) and its standard error.  Control variate: c = final value - S with the known
// mean E[c] of control_means, so  price = mean(pf) - beta (mean(c) - E[c]),
// beta = cov(pf, c)/var(c), and the variance shrinks by 1 - corr(pf, c)^2.
static int finish_prices(const hexo_price_request* r, const double* sums, double* prices,
                         double* se) {
  std::vector<double> ec(r->n_chains, 0.0);
  const uint32_t n_opts = r->strike_offsets[r->n_chains];
  if (r->control_variate) {
    std::vector<SegConst> segs;
    const int rc = build_segments(r, true, segs, nullptr);
    if (rc) return rc;
    if (r->control_variate == HEXO_CV_GEOMETRIC) {
      // per-option control c_j = max(G - K_j, 0) with its semi-analytic mean (geo_asian_host.cu);
      // sums = [pf | pf^2 | pf c | c | c^2], n_opts each
      std::vector<double> eg;
      const int rg = geometric_asian_means(r, segs, eg);
      if (rg) return fail(rg, "the geometric-Asian control's mean is not finite for these parameters");
      const double n = (double)r->n_paths;
      const double *sp = sums, *sq = sums + n_opts, *sx = sums + 2 * (size_t)n_opts,
                   *sc = sums + 3 * (size_t)n_opts, *sc2 = sums + 4 * (size_t)n_opts;
      for (uint32_t j = 0; j < n_opts; ++j) {
        const double mean = sp[j] / n, mean_c = sc[j] / n;
        double var = n > 1 ? std::max(0.0, (sq[j] - n * mean * mean) / (n - 1)) : 0.0;
        const double var_c = n > 1 ? std::max(0.0, (sc2[j] - n * mean_c * mean_c) / (n - 1)) : 0.0;
        double price = mean;
        if (var_c > 0.0) {
          const double cov = (sx[j] - n * mean * mean_c) / (n - 1);
          const double beta = cov / var_c;
          price = mean - beta * (mean_c - eg[j]);
          var = std::max(0.0, var - beta * cov);
        }
        prices[j] = price;
        if (se) se[j] = sqrt(var / n);
      }
      return HEXO_OK;
    }
    control_means(r, segs, ec);
  }
  const double n = (double)r->n_paths;
  const double* sp = sums;
  const double* sq = sums + n_opts;
  const double* sx = sums + 2 * (size_t)n_opts;
  const double* sc = sums + 3 * (size_t)n_opts;
  const double* sc2 = sc + r->n_chains;
  for (uint32_t k = 0; k < r->n_chains; ++k) {
    double mean_c = 0.0, var_c = 0.0;
    if (r->control_variate) {
      mean_c = sc[k] / n;
      var_c = n > 1 ? std::max(0.0, (sc2[k] - n * mean_c * mean_c) / (n - 1)) : 0.0;
    }
    for (uint32_t j = r->strike_offsets[k]; j < r->strike_offsets[k + 1]; ++j) {
      const double mean = sp[j] / n;
      double var = n > 1 ? std::max(0.0, (sq[j] - n * mean * mean) / (n - 1)) : 0.0;
      double price = mean;
      if (r->control_variate && var_c > 0.0) {
        const double cov = (sx[j] - n * mean * mean_c) / (n - 1);
        const double beta = cov / var_c;
        price = mean - beta * (mean_c - ec[k]);
        var = std::max(0.0, var - beta * cov);
      }
      prices[j] = price;
      if (se) se[j] = sqrt(var / n);
    }
  }
  return HEXO_OK;
}

// ---------------------------------------------------------------------------
// host: a prepared launch ("plan") of the path kernel for one shard
// ---------------------------------------------------------------------------
struct Plan {
  PathArgs args{};
  int payoff = 0, normal_mode = 0, rng_mode = 0;
  uint32_t grid = 0, block = 0, smem = 0, n_opts = 0, n_sums = 0;
  uint64_t steps_per_path = 0, path_steps = 0, n_streams = 0;
  void* blob = nullptr;       // [segs | strikes | partials | sums]
  double* sums_dev = nullptr; // inside blob unless caller-supplied
  size_t gacc_bytes = 0;
  int cv = 0;  // hexo_control_variate (template parameter CV of the path kernel)
  bool mart = false;  // HEXO_DRIFT_MARTINGALE (template parameter MART)
};

// The path-kernel instantiations live in their own translation units (path_kernels_*.cu), which
// build in parallel; path_kernels.h declares the selectors.
static PathKernel pick_kernel(int payoff, int normal_mode, uint32_t n_seg, int rng_mode = 0,
                              int cv = 0, bool mart = false) {
  const int segs = n_seg == 1                       ? kSegsSingle
                   : n_seg <= (uint32_t)kInlineSegs ? kSegsInline
                                                    : kSegsGlobal;
  if (cv == HEXO_CV_GEOMETRIC) return path_kernel_shishua_geo(normal_mode, segs, mart);
  if (mart)
    return rng_mode == HEXO_RNG_PHILOX ? path_kernel_philox_mart(payoff, normal_mode, segs, cv)
                                       : path_kernel_shishua_mart(payoff, normal_mode, segs, cv);
  if (rng_mode == HEXO_RNG_PHILOX) return path_kernel_philox(payoff, normal_mode, segs, cv);
  if (normal_mode == HEXO_NORMAL_F32_PPND7) return path_kernel_shishua_ppnd7(payoff, segs);
  return cv ? path_kernel_shishua_cv(payoff, normal_mode, segs)
            : path_kernel_shishua(payoff, normal_mode, segs);
}

static uint64_t default_streams(uint64_t n_paths, int n_gpus) {
  // one wave of resident threads per GPU
  const uint64_t per_gpu = (uint64_t)(g_cur && g_ctx.ready ? g_ctx.sm_count : 148) *
                           (uint64_t)(kMinBlocksPerSM * kMaxBlock);
  uint64_t s = per_gpu * (uint64_t)std::max(n_gpus, 1);
  if (s > n_paths) s = n_paths;
  return std::max<uint64_t>(s, 1);
}

static int plan_destroy(Plan* p, cudaStream_t st) {
  if (p->blob) cudaFreeAsync(p->blob, st);
  p->blob = nullptr;
  return HEXO_OK;
}

static int plan_fill(const hexo_price_request* r, uint64_t stream_begin, uint64_t stream_count,
                     cudaStream_t st, Plan* p);
// builds the plan; on failure nothing stays allocated
static int plan_create(const hexo_price_request* r, uint64_t stream_begin, uint64_t stream_count,
                       cudaStream_t st, Plan* p) {
  const int rc = plan_fill(r, stream_begin, stream_count, st, p);
  if (rc) plan_destroy(p, st);
  return rc;
}
static int plan_fill(const hexo_price_request* r, uint64_t stream_begin, uint64_t stream_count,
                     cudaStream_t st, Plan* p) {
  int rc = check_request(r, true);
  if (rc) return rc;
  rc = ensure_context();
  if (rc) return rc;
  if (r->n_streams == 0) return fail(HEXO_ERR_INVALID_ARGUMENT, "n_streams must be set for a shard");
  if (stream_count == 0 || stream_begin + stream_count > r->n_streams)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "stream range [%llu,+%llu) outside n_streams=%llu",
                (unsigned long long)stream_begin, (unsigned long long)stream_count,
                (unsigned long long)r->n_streams);
  std::vector<SegConst> segs;
  rc = build_segments(r, true, segs, &p->steps_per_path);
  if (rc) return rc;
  const uint32_t n_opts = r->strike_offsets[r->n_chains];
  // Per-warp option accumulators live in shared memory while that leaves room for
  // kMinBlocksPerSM blocks per SM; larger chains accumulate in (L2-resident) device memory.
  int block = kMaxBlock;
#ifdef HEXO_DEV_PROBES  // development builds only (build.py --dev): never in the shipped library
  if (const char* e = getenv("HEXO_BLOCK")) {  // 32..256, multiple of 32
    const int b = atoi(e);
    if (b >= 32 && b <= kMaxBlock && b % 32 == 0) block = b;
  }
#endif
  p->rng_mode = r->rng_mode;
  const uint32_t n_sums = (uint32_t)sums_len(r);
  p->n_sums = n_sums;
  const size_t smem_budget =
      std::min(g_ctx.smem_optin, (size_t)(227 * 1024) / kMinBlocksPerSM);
  const int streams_per_block = block;  // path-owning threads per block
  auto smem_of = [&](bool acc) { return path_kernel_smem(block, n_sums, r->normal_mode, acc); };
  const bool acc_in_smem = smem_of(true) <= smem_budget;
  p->payoff = r->payoff;
  p->normal_mode = r->normal_mode;
  p->n_opts = n_opts;
  p->block = (uint32_t)block;
  p->smem = (uint32_t)smem_of(acc_in_smem);
  const uint64_t grid64 = (stream_count + streams_per_block - 1) / streams_per_block;
  if (grid64 > 0x7fffffffull) return fail(HEXO_ERR_TOO_LARGE, "too many streams for one launch");
  p->grid = (uint32_t)grid64;
  p->n_streams = r->n_streams;
  p->path_steps = r->n_paths * (uint64_t)r->steps;

  const size_t seg_bytes = segs.size() * sizeof(SegConst);
  const size_t strike_bytes = (size_t)n_opts * sizeof(double);
  const size_t part_bytes = (size_t)p->grid * n_sums * sizeof(double);
  const size_t sums_bytes = (size_t)n_sums * sizeof(double);
  const size_t gacc_bytes =
      acc_in_smem ? 0 : (size_t)p->grid * (streams_per_block / 32) * n_sums * sizeof(double);
  if (gacc_bytes > ((size_t)8 << 30))
    return fail(HEXO_ERR_TOO_LARGE, "%u options x %u blocks need %zu bytes of accumulators", n_opts,
                p->grid, gacc_bytes);
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t off_strikes = up(seg_bytes), off_part = off_strikes + up(strike_bytes),
               off_sums = off_part + up(part_bytes), off_gacc = off_sums + up(sums_bytes),
               total = off_gacc + up(gacc_bytes);
  HEXO_CUDA(cudaMallocAsync(&p->blob, total, st));
  unsigned char* base = static_cast<unsigned char*>(p->blob);
  HEXO_CUDA(cudaMemcpyAsync(base, segs.data(), seg_bytes, cudaMemcpyHostToDevice, st));
  HEXO_CUDA(cudaMemcpyAsync(base + off_strikes, r->strikes, strike_bytes, cudaMemcpyHostToDevice, st));
  p->sums_dev = reinterpret_cast<double*>(base + off_sums);

  PathArgs& a = p->args;
  a.v0 = r->p.v_0;
  a.S = r->S;
  a.lnS = log(r->S);  // HSimulation.tpp:90
  a.seed = r->seed;
  a.stream_begin = stream_begin;
  a.stream_count = stream_count;
  a.base_paths = r->n_paths / r->n_streams;
  a.rem_streams = r->n_paths % r->n_streams;
  a.n_seg = r->n_chains;
  a.n_opts = n_opts;
  p->cv = r->control_variate;
  p->mart = r->drift_mode == HEXO_DRIFT_MARTINGALE;
  a.segs = reinterpret_cast<const SegConst*>(base);
  for (size_t k = 0; k < segs.size() && k < (size_t)kInlineSegs; ++k) a.seg_inline[k] = segs[k];
  a.strikes = reinterpret_cast<const double*>(base + off_strikes);
  a.partials = reinterpret_cast<double*>(base + off_part);
#ifdef HEXO_DEV_PROBES
  {
    const char* e = getenv("HEXO_NO_REFILL");
    a.dev_no_refill = e ? (uint32_t)atoi(e) : 0u;
  }
#endif
  a.gacc = acc_in_smem ? nullptr : reinterpret_cast<double*>(base + off_gacc);
  p->gacc_bytes = gacc_bytes;

  // The attribute belongs to the kernel function, not to the plan, and a later, smaller value
  // would overwrite it: always ask for the device maximum, so that plans with different option
  // counts (and host threads) can coexist.  It is a cap; occupancy follows the size actually
  // passed at launch.
  PathKernel kern = pick_kernel(p->payoff, p->normal_mode, p->args.n_seg, p->rng_mode, p->cv,
                                p->mart);
  HEXO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)g_ctx.smem_optin));
  return HEXO_OK;
}

// enqueue path kernel + reduction; sums land in `sums_out_dev` (or the plan's own buffer)
static int plan_launch(const Plan* p, cudaStream_t st, double* sums_out_dev) {
  if (p->args.gacc) HEXO_CUDA(cudaMemsetAsync(p->args.gacc, 0, p->gacc_bytes, st));
  PathKernel kern = pick_kernel(p->payoff, p->normal_mode, p->args.n_seg, p->rng_mode, p->cv,
                                p->mart);
  kern<<<p->grid, p->block, p->smem, st>>>(p->args);
  HEXO_CUDA(cudaGetLastError());
  const uint32_t n2 = p->n_sums;
  reduce_partials_kernel<<<(n2 + 127) / 128, 128, 0, st>>>(p->args.partials, p->grid, n2,
                                                            sums_out_dev ? sums_out_dev : p->sums_dev);
  HEXO_CUDA(cudaGetLastError());
  return HEXO_OK;
}

static void fill_stats(const Plan& p, float ms, hexo_gpu_stats* s) {
  if (!s) return;
  s->n_streams = p.n_streams;
  s->steps_per_path = p.steps_per_path;
  s->path_steps = p.path_steps;
  s->grid = p.grid;
  s->block = p.block;
  s->smem_bytes = p.smem;
  s->kernel_launches = 2;
  s->kernel_ms = ms;
}

int build_request_segments(const hexo_price_request* r, std::vector<SegConst>& segs) {
  return build_segments(r, true, segs, nullptr);
}

}  // namespace hexo

// ===========================================================================
// C ABI
// ===========================================================================
using namespace hexo;

extern "C" {

int hexo_gpu_abi_version(void) { return HEXO_GPU_ABI_VERSION; }

const char* hexo_gpu_last_error(void) { return g_err; }

int hexo_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int hexo_gpu_init(int device) {
  int n = hexo_gpu_device_count();
  if (n == 0) return fail(HEXO_ERR_NO_DEVICE, "no CUDA device visible");
  if (device < 0 || device >= n)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", device, n);
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != device) HEXO_CUDA(cudaSetDevice(device));
  return ensure_context();  // cheap when the device is already set up: callers may init per call
}

int hexo_gpu_shutdown(void) {
  for (int d = 0; d < kMaxDevices; ++d) g_ctxs[d].ready = false;
  g_cur = nullptr;
  return HEXO_OK;
}

int hexo_gpu_schedule(const double* expiries, uint32_t n_chains, uint32_t steps,
                      hexo_segment* segments_out) {
  return build_schedule(expiries, n_chains, steps, segments_out);
}

int hexo_gpu_schedule_exact(const double* expiries, uint32_t n_chains, uint32_t steps,
                            hexo_segment* segments_out) {
  return build_schedule_exact(expiries, n_chains, steps, segments_out);
}

uint64_t hexo_gpu_default_streams(uint64_t n_paths, uint32_t n_opts, int n_gpus) {
  (void)n_opts;
  ensure_context();
  return default_streams(n_paths, n_gpus);
}

struct hexo_gpu_plan {
  Plan p;
};

int hexo_gpu_plan_create(const hexo_price_request* req, uint64_t stream_begin,
                         uint64_t stream_count, hexo_gpu_plan** plan_out) {
  if (!plan_out) return fail(HEXO_ERR_INVALID_ARGUMENT, "plan_out is NULL");
  hexo_gpu_plan* h = new hexo_gpu_plan();
  int rc = plan_create(req, stream_begin, stream_count, 0, &h->p);
  if (rc == HEXO_OK) {
    cudaError_t e = cudaStreamSynchronize(0);  // inputs resident before the first launch
    if (e != cudaSuccess) rc = fail(HEXO_ERR_CUDA, "plan upload: %s", cudaGetErrorString(e));
  }
  if (rc) {
    plan_destroy(&h->p, 0);
    delete h;
    return rc;
  }
  *plan_out = h;
  return HEXO_OK;
}

int hexo_gpu_plan_launch(hexo_gpu_plan* plan, double* sums_device, void* cuda_stream) {
  if (!plan) return fail(HEXO_ERR_INVALID_ARGUMENT, "plan is NULL");
  return plan_launch(&plan->p, static_cast<cudaStream_t>(cuda_stream), sums_device);
}

double* hexo_gpu_plan_sums_device(hexo_gpu_plan* plan) { return plan ? plan->p.sums_dev : nullptr; }

int hexo_gpu_plan_stats(const hexo_gpu_plan* plan, hexo_gpu_stats* stats) {
  if (!plan || !stats) return fail(HEXO_ERR_INVALID_ARGUMENT, "plan / stats is NULL");
  fill_stats(plan->p, 0.f, stats);
  return HEXO_OK;
}

int hexo_gpu_plan_destroy(hexo_gpu_plan* plan) {
  if (!plan) return HEXO_OK;
  cudaDeviceSynchronize();
  plan_destroy(&plan->p, 0);
  delete plan;
  return HEXO_OK;
}

int hexo_gpu_price_shard_device(const hexo_price_request* req, uint64_t stream_begin,
                                uint64_t stream_count, double* sums_device, void* cuda_stream,
                                hexo_gpu_stats* stats) {
  if (!sums_device) return fail(HEXO_ERR_INVALID_ARGUMENT, "sums_device is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  Plan p;
  int rc = plan_create(req, stream_begin, stream_count, st, &p);
  if (rc) return rc;
  rc = plan_launch(&p, st, sums_device);
  plan_destroy(&p, st);  // stream-ordered free: runs after the kernels
  fill_stats(p, 0.f, stats);
  return rc;
}

int hexo_gpu_price_shard(const hexo_price_request* req, uint64_t stream_begin,
                         uint64_t stream_count, double* sums_out, hexo_gpu_stats* stats) {
  if (!sums_out) return fail(HEXO_ERR_INVALID_ARGUMENT, "sums_out is NULL");
  cudaStream_t st = 0;
  Plan p;
  int rc = plan_create(req, stream_begin, stream_count, st, &p);
  if (rc) return rc;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t ee = cudaEventCreate(&e0);
  if (ee == cudaSuccess) ee = cudaEventCreate(&e1);
  if (ee == cudaSuccess) ee = cudaEventRecord(e0, st);
  if (ee != cudaSuccess) {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    plan_destroy(&p, st);
    return fail(HEXO_ERR_CUDA, "timing events: %s", cudaGetErrorString(ee));
  }
  rc = plan_launch(&p, st, nullptr);
  if (rc == HEXO_OK) {
    cudaEventRecord(e1, st);
    cudaError_t e = cudaMemcpyAsync(sums_out, p.sums_dev, (size_t)p.n_sums * sizeof(double),
                                    cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(HEXO_ERR_CUDA, "path kernel failed: %s", cudaGetErrorString(e));
  }
  float ms = 0.f;
  if (rc == HEXO_OK) cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  plan_destroy(&p, st);
  fill_stats(p, ms, stats);
  return rc;
}

int hexo_gpu_price(const hexo_price_request* req, double* prices_out, double* stderr_out,
                   hexo_gpu_stats* stats) {
  int rc = check_request(req, true);
  if (rc) return rc;
  if (!prices_out) return fail(HEXO_ERR_INVALID_ARGUMENT, "prices_out is NULL");
  rc = ensure_context();
  if (rc) return rc;
  hexo_price_request r = *req;
  if (r.n_streams == 0) r.n_streams = default_streams(r.n_paths, 1);
  std::vector<double> sums(sums_len(&r));
  rc = hexo_gpu_price_shard(&r, 0, r.n_streams, sums.data(), stats);
  if (rc) return rc;
  return finish_prices(&r, sums.data(), prices_out, stderr_out);
}

size_t hexo_gpu_sums_len(const hexo_price_request* req) {
  if (!req || !req->strike_offsets || req->n_chains == 0) return 0;
  return sums_len(req);
}

int hexo_heston_geometric_asian(const hexo_price_request* req, double* means_out) {
  int rc = check_request(req, true);
  if (rc) return rc;
  if (!means_out) return fail(HEXO_ERR_INVALID_ARGUMENT, "means_out is NULL");
  if (req->payoff != HEXO_PAYOFF_ASIAN)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "the geometric average belongs to the Asian payoff");
  std::vector<SegConst> segs;
  rc = build_segments(req, true, segs, nullptr);
  if (rc) return rc;
  std::vector<double> eg;
  rc = geometric_asian_means(req, segs, eg);
  if (rc) return fail(rc, "the geometric-Asian control's mean is not finite for these parameters");
  for (size_t j = 0; j < eg.size(); ++j) means_out[j] = eg[j];
  return HEXO_OK;
}

int hexo_gpu_finish(const hexo_price_request* req, const double* sums, double* prices_out,
                    double* stderr_out) {
  int rc = check_request(req, true);
  if (rc) return rc;
  if (!sums || !prices_out) return fail(HEXO_ERR_INVALID_ARGUMENT, "finish: sums / prices_out is NULL");
  return finish_prices(req, sums, prices_out, stderr_out);
}

// Many independent price<>() calls in one submission (Monte-Carlo inside a
// calibration loop prices the same chains for many HParams).  Request i runs on CUDA stream
// i % n_lanes of the current device, so launches, the tail of one job and the head of the next
// overlap; every job computes exactly what hexo_gpu_price computes for it (same streams, same
// sums).  One device-to-host copy of all sums at the end.
int hexo_gpu_price_batch(const hexo_price_request* reqs, uint32_t n_reqs, uint32_t n_lanes,
                         double* prices_out, double* stderr_out, hexo_gpu_stats* stats) {
  if (!reqs || n_reqs == 0) return fail(HEXO_ERR_INVALID_ARGUMENT, "empty batch");
  if (!prices_out) return fail(HEXO_ERR_INVALID_ARGUMENT, "prices_out is NULL");
  std::vector<size_t> off(n_reqs + 1, 0);
  for (uint32_t i = 0; i < n_reqs; ++i) {
    const int rc = check_request(&reqs[i], true);
    if (rc) return rc;
    off[i + 1] = off[i] + sums_len(&reqs[i]);
  }
  int rc = ensure_context();
  if (rc) return rc;
  if (n_lanes == 0) n_lanes = 16;
  n_lanes = std::min(n_lanes, std::min<uint32_t>(n_reqs, 32));
  std::vector<cudaStream_t> lanes(n_lanes, nullptr);
  std::vector<cudaEvent_t> done(n_lanes, nullptr);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double* all_sums = nullptr;
  std::vector<Plan> plans(n_reqs);
  std::vector<double> sums(off[n_reqs]);
  float ms = 0.f;
  auto cleanup = [&]() {
    for (auto st : lanes)
      if (st) cudaStreamDestroy(st);
    for (auto ev : done)
      if (ev) cudaEventDestroy(ev);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (all_sums) cudaFree(all_sums);
  };
  cudaError_t e = cudaMalloc(&all_sums, off[n_reqs] * sizeof(double));
  for (uint32_t l = 0; l < n_lanes && e == cudaSuccess; ++l) {
    e = cudaStreamCreateWithFlags(&lanes[l], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[l], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e != cudaSuccess) {
    cleanup();
    return fail(HEXO_ERR_CUDA, "batch setup: %s", cudaGetErrorString(e));
  }
  cudaEventRecord(e0, lanes[0]);
  for (uint32_t l = 1; l < n_lanes; ++l) cudaStreamWaitEvent(lanes[l], e0, 0);
  for (uint32_t i = 0; i < n_reqs && rc == HEXO_OK; ++i) {
    cudaStream_t st = lanes[i % n_lanes];
    hexo_price_request r = reqs[i];
    if (r.n_streams == 0) r.n_streams = default_streams(r.n_paths, 1);
    rc = plan_create(&r, 0, r.n_streams, st, &plans[i]);
    if (rc) break;
    rc = plan_launch(&plans[i], st, all_sums + off[i]);
    plan_destroy(&plans[i], st);  // stream-ordered: freed after the job's kernels
  }
  for (uint32_t l = 1; l < n_lanes; ++l) {
    cudaEventRecord(done[l], lanes[l]);
    cudaStreamWaitEvent(lanes[0], done[l], 0);
  }
  cudaEventRecord(e1, lanes[0]);
  e = cudaStreamSynchronize(lanes[0]);
  for (uint32_t l = 1; l < n_lanes; ++l) cudaStreamSynchronize(lanes[l]);
  if (rc == HEXO_OK && e == cudaSuccess)
    e = cudaMemcpy(sums.data(), all_sums, sums.size() * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  cleanup();
  if (rc) return rc;
  if (e != cudaSuccess) return fail(HEXO_ERR_CUDA, "batch: %s", cudaGetErrorString(e));
  size_t out_off = 0;
  for (uint32_t i = 0; i < n_reqs; ++i) {
    rc = finish_prices(&reqs[i], sums.data() + off[i], prices_out + out_off,
                       stderr_out ? stderr_out + out_off : nullptr);
    if (rc) return rc;
    out_off += reqs[i].strike_offsets[reqs[i].n_chains];
    if (stats) fill_stats(plans[i], ms, &stats[i]);  // kernel_ms = the whole batch
  }
  return HEXO_OK;
}

// One process, several GPUs: the single-process form of the multi-GPU path for callers like the
// reference's CLI, which is one process.  Streams are split over the first n_gpus devices
// exactly like hx.price_distributed splits them over ranks; the 2*n_opts sums of each device come
// back over PCIe and are added on the host (16 bytes per option -- no collective needed).
int hexo_gpu_price_multi(const hexo_price_request* req, int n_gpus, double* prices_out,
                         double* stderr_out, hexo_gpu_stats* stats) {
  int rc = check_request(req, true);
  if (rc) return rc;
  if (!prices_out) return fail(HEXO_ERR_INVALID_ARGUMENT, "prices_out is NULL");
  const int n_dev = hexo_gpu_device_count();
  if (n_dev == 0) return fail(HEXO_ERR_NO_DEVICE, "no CUDA device visible");
  if (n_gpus <= 0) n_gpus = n_dev;
  if (n_gpus > n_dev)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "%d GPUs requested, %d visible", n_gpus, n_dev);
  int home = 0;
  HEXO_CUDA(cudaGetDevice(&home));
  rc = ensure_context();
  if (rc) return rc;
  hexo_price_request r = *req;
  if (r.n_streams == 0) r.n_streams = default_streams(r.n_paths, n_gpus);
  std::vector<Plan> plans(n_gpus);
  std::vector<int> used(n_gpus, 0);
  std::vector<cudaEvent_t> ev0(n_gpus), ev1(n_gpus);
  const uint64_t base = r.n_streams / n_gpus, rem = r.n_streams % n_gpus;
  for (int g = 0; g < n_gpus && rc == HEXO_OK; ++g) {  // enqueue everywhere first
    const uint64_t begin = g * base + std::min<uint64_t>(g, rem), count = base + (g < (int)rem);
    if (count == 0) continue;
    if (cudaSetDevice(g) != cudaSuccess) {
      rc = fail(HEXO_ERR_CUDA, "cudaSetDevice(%d) failed", g);
      break;
    }
    rc = plan_create(&r, begin, count, 0, &plans[g]);
    if (rc) break;
    used[g] = 1;
    cudaEventCreate(&ev0[g]);
    cudaEventCreate(&ev1[g]);
    cudaEventRecord(ev0[g], 0);
    rc = plan_launch(&plans[g], 0, nullptr);
    cudaEventRecord(ev1[g], 0);
  }
  std::vector<double> sums(sums_len(&r), 0.0), part(sums_len(&r));
  float ms_max = 0.f;
  for (int g = 0; g < n_gpus; ++g) {  // then collect
    if (!used[g]) continue;
    cudaSetDevice(g);
    if (rc == HEXO_OK) {
      cudaError_t e = cudaMemcpy(part.data(), plans[g].sums_dev, part.size() * sizeof(double),
                                 cudaMemcpyDeviceToHost);
      if (e != cudaSuccess)
        rc = fail(HEXO_ERR_CUDA, "device %d: %s", g, cudaGetErrorString(e));
      else
        for (size_t j = 0; j < part.size(); ++j) sums[j] += part[j];
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev0[g], ev1[g]) == cudaSuccess) ms_max = std::max(ms_max, ms);
    } else {
      cudaDeviceSynchronize();
    }
    cudaEventDestroy(ev0[g]);
    cudaEventDestroy(ev1[g]);
    plan_destroy(&plans[g], 0);
  }
  cudaSetDevice(home);
  if (rc) return rc;
  fill_stats(plans[0], ms_max, stats);
  if (stats) stats->kernel_launches = 2 * (uint32_t)n_gpus;
  return finish_prices(&r, sums.data(), prices_out, stderr_out);
}

int hexo_gpu_shishua_streams(uint64_t seed, uint64_t first_stream, uint32_t n_streams,
                             uint8_t* bytes_out, size_t bytes_per_stream) {
  if (!bytes_out || n_streams == 0 || bytes_per_stream == 0 || (bytes_per_stream & 127))
    return fail(HEXO_ERR_INVALID_ARGUMENT, "shishua: need a buffer and a multiple of 128 bytes");
  int rc = ensure_context();
  if (rc) return rc;
  uint64_t* d = nullptr;
  const size_t total = (size_t)n_streams * bytes_per_stream;
  HEXO_CUDA(cudaMalloc(&d, total));
  shishua_streams_kernel<<<(n_streams + 63) / 64, 64>>>(seed, first_stream, 0, 0, n_streams, d,
                                                        bytes_per_stream / 8);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(bytes_out, d, total, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(HEXO_ERR_CUDA, "shishua kernel: %s", cudaGetErrorString(e));
  return HEXO_OK;
}

int hexo_gpu_philox4x32(const uint32_t* counters, const uint32_t* keys, uint32_t* out, size_t n) {
  if (!counters || !keys || !out || n == 0)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "philox: bad args");
  int rc = ensure_context();
  if (rc) return rc;
  uint32_t *dc = nullptr, *dk = nullptr, *dout = nullptr;
  HEXO_CUDA(cudaMalloc(&dc, n * 16));
  HEXO_CUDA(cudaMalloc(&dk, n * 8));
  HEXO_CUDA(cudaMalloc(&dout, n * 16));
  cudaError_t e = cudaMemcpy(dc, counters, n * 16, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dk, keys, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    philox_kernel<<<(unsigned)((n + 127) / 128), 128>>>(dc, dk, dout, n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, dout, n * 16, cudaMemcpyDeviceToHost);
  cudaFree(dc); cudaFree(dk); cudaFree(dout);
  if (e != cudaSuccess) return fail(HEXO_ERR_CUDA, "philox kernel: %s", cudaGetErrorString(e));
  return HEXO_OK;
}

int hexo_gpu_philox_streams(uint64_t seed, uint64_t first_stream, uint32_t n_streams,
                            uint64_t* words_out, size_t words_per_stream) {
  if (!words_out || n_streams == 0 || words_per_stream == 0 || (words_per_stream & 15))
    return fail(HEXO_ERR_INVALID_ARGUMENT, "philox: need a buffer and a multiple of 16 words");
  int rc = ensure_context();
  if (rc) return rc;
  uint64_t* d = nullptr;
  const size_t total = (size_t)n_streams * words_per_stream * 8;
  HEXO_CUDA(cudaMalloc(&d, total));
  philox_streams_kernel<<<(n_streams + 63) / 64, 64>>>(seed, first_stream, n_streams, d,
                                                       words_per_stream);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(words_out, d, total, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(HEXO_ERR_CUDA, "philox kernel: %s", cudaGetErrorString(e));
  return HEXO_OK;
}

int hexo_gpu_shishua_fill(const uint64_t seed[4], uint8_t* bytes_out, size_t n_bytes) {
  if (!seed || !bytes_out || n_bytes == 0 || (n_bytes & 127))
    return fail(HEXO_ERR_INVALID_ARGUMENT, "shishua: need a buffer and a multiple of 128 bytes");
  int rc = ensure_context();
  if (rc) return rc;
  uint64_t* d = nullptr;
  HEXO_CUDA(cudaMalloc(&d, n_bytes));
  shishua_streams_kernel<<<1, 32>>>(seed[0], seed[1], seed[2], seed[3], 1, d, n_bytes / 8);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(bytes_out, d, n_bytes, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(HEXO_ERR_CUDA, "shishua kernel: %s", cudaGetErrorString(e));
  return HEXO_OK;
}

int hexo_gpu_u64_to_unit(const uint64_t* bits_in, double* u_out, size_t n) {
  if (!bits_in || !u_out || n == 0) return fail(HEXO_ERR_INVALID_ARGUMENT, "u64_to_unit: bad args");
  int rc = ensure_context();
  if (rc) return rc;
  uint64_t* din = nullptr;
  double* dout = nullptr;
  HEXO_CUDA(cudaMalloc(&din, n * 8));
  HEXO_CUDA(cudaMalloc(&dout, n * 8));
  cudaError_t e = cudaMemcpy(din, bits_in, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    u64_to_unit_kernel<<<(unsigned)((n + 255) / 256), 256>>>(din, dout, n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(u_out, dout, n * 8, cudaMemcpyDeviceToHost);
  cudaFree(din);
  cudaFree(dout);
  if (e != cudaSuccess) return fail(HEXO_ERR_CUDA, "u64_to_unit: %s", cudaGetErrorString(e));
  return HEXO_OK;
}

int hexo_gpu_ppnd16(const double* u_in, double* z_out, size_t n, int normal_mode) {
  if (!u_in || !z_out || n == 0) return fail(HEXO_ERR_INVALID_ARGUMENT, "ppnd16: bad args");
  if (normal_mode != HEXO_NORMAL_F32 && normal_mode != HEXO_NORMAL_F64)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "unknown normal_mode %d", normal_mode);
  int rc = ensure_context();
  if (rc) return rc;
  double *din = nullptr, *dout = nullptr;
  HEXO_CUDA(cudaMalloc(&din, n * 8));
  HEXO_CUDA(cudaMalloc(&dout, n * 8));
  cudaError_t e = cudaMemcpy(din, u_in, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (normal_mode == HEXO_NORMAL_F64)
      ppnd16_kernel<1><<<grid, 256>>>(din, dout, n);
    else
      ppnd16_kernel<0><<<grid, 256>>>(din, dout, n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(z_out, dout, n * 8, cudaMemcpyDeviceToHost);
  cudaFree(din);
  cudaFree(dout);
  if (e != cudaSuccess) return fail(HEXO_ERR_CUDA, "ppnd16: %s", cudaGetErrorString(e));
  return HEXO_OK;
}

int hexo_gpu_normals_from_words(const uint64_t* words_in, double* z_out, size_t n, int normal_mode) {
  if (!words_in || !z_out || n == 0)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "normals_from_words: bad args");
  if (normal_mode != HEXO_NORMAL_F32 && normal_mode != HEXO_NORMAL_F64 &&
      normal_mode != HEXO_NORMAL_F32_PPND7)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "unknown normal_mode %d", normal_mode);
  int rc = ensure_context();
  if (rc) return rc;
  const size_t kw = 2 * (size_t)ring_steps(normal_mode);
  const size_t n_chunks = (n + kw - 1) / kw, padded = n_chunks * kw;
  const int block = kMaxBlock;
  const size_t smem = ring_smem(block, normal_mode) + (size_t)kTailListBytes * (block / 32);
  uint64_t* dw = nullptr;
  double* dz = nullptr;
  HEXO_CUDA(cudaMalloc(&dw, padded * 8));
  HEXO_CUDA(cudaMalloc(&dz, padded * 8));
  // pad the last chunk with central words (p = 1/2)
  std::vector<uint64_t> pad(padded - n, 0x8000000000000000ull);
  cudaError_t e = cudaMemcpy(dw, words_in, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !pad.empty())
    e = cudaMemcpy(dw + n, pad.data(), pad.size() * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)((n_chunks + block - 1) / block);
    auto kern = normal_mode == HEXO_NORMAL_F64         ? normals_from_words_kernel<HEXO_NORMAL_F64>
                : normal_mode == HEXO_NORMAL_F32_PPND7 ? normals_from_words_kernel<HEXO_NORMAL_F32_PPND7>
                                                       : normals_from_words_kernel<HEXO_NORMAL_F32>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)g_ctx.smem_optin);
    if (e == cudaSuccess) kern<<<grid, block, smem>>>(dw, dz, n_chunks);
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(z_out, dz, n * 8, cudaMemcpyDeviceToHost);
  cudaFree(dw);
  cudaFree(dz);
  if (e != cudaSuccess) return fail(HEXO_ERR_CUDA, "normals_from_words: %s", cudaGetErrorString(e));
  return HEXO_OK;
}

int hexo_gpu_replay(const hexo_price_request* req, const double* tape, uint64_t n_paths,
                    uint32_t tape_steps, double* finals_out, uint32_t* steps_used_out) {
  int rc = check_request(req, false);
  if (rc) return rc;
  if (!tape || !finals_out || n_paths == 0)
    return fail(HEXO_ERR_INVALID_ARGUMENT, "replay: tape / finals_out / n_paths");
  std::vector<SegConst> segs;
  uint64_t steps_per_path = 0;
  rc = build_segments(req, false, segs, &steps_per_path);
  if (rc) return rc;
  if (steps_per_path > tape_steps)
    return fail(HEXO_ERR_TAPE_TOO_SHORT, "schedule needs %llu steps per path, tape has %u",
                (unsigned long long)steps_per_path, tape_steps);
  rc = ensure_context();
  if (rc) return rc;
  const size_t tape_bytes = (size_t)n_paths * tape_steps * 3 * sizeof(double);
  const size_t fin_bytes = (size_t)n_paths * req->n_chains * sizeof(double);
  SegConst* dseg = nullptr;
  double *dtape = nullptr, *dfin = nullptr;
  HEXO_CUDA(cudaMalloc(&dseg, segs.size() * sizeof(SegConst)));
  HEXO_CUDA(cudaMalloc(&dtape, tape_bytes));
  HEXO_CUDA(cudaMalloc(&dfin, fin_bytes));
  cudaError_t e = cudaMemcpy(dseg, segs.data(), segs.size() * sizeof(SegConst), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dtape, tape, tape_bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)((n_paths + 127) / 128);
    const bool mart = req->drift_mode == HEXO_DRIFT_MARTINGALE;
    auto kern = req->payoff == HEXO_PAYOFF_ASIAN
                    ? (mart ? qe_replay_kernel<HEXO_PAYOFF_ASIAN, true>
                            : qe_replay_kernel<HEXO_PAYOFF_ASIAN, false>)
                    : (mart ? qe_replay_kernel<HEXO_PAYOFF_EUROPEAN, true>
                            : qe_replay_kernel<HEXO_PAYOFF_EUROPEAN, false>);
    kern<<<grid, 128>>>(req->p.v_0, req->S, log(req->S), req->n_chains, dseg, dtape, n_paths,
                        tape_steps, dfin);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(finals_out, dfin, fin_bytes, cudaMemcpyDeviceToHost);
  cudaFree(dseg);
  cudaFree(dtape);
  cudaFree(dfin);
  if (e != cudaSuccess) return fail(HEXO_ERR_CUDA, "replay: %s", cudaGetErrorString(e));
  if (steps_used_out) *steps_used_out = (uint32_t)steps_per_path;
  return HEXO_OK;
}

int hexo_gpu_measure_fp64_peak(double* flops_out, float* ms_out) {
  if (!flops_out) return fail(HEXO_ERR_INVALID_ARGUMENT, "flops_out is NULL");
  int rc = ensure_context();
  if (rc) return rc;
  const int block = 256, grid = g_ctx.sm_count * 8, iters = 4096;
  double* d = nullptr;
  HEXO_CUDA(cudaMalloc(&d, (size_t)grid * block * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<grid, block>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (e != cudaSuccess) return fail(HEXO_ERR_CUDA, "dfma peak: %s", cudaGetErrorString(e));
  const double fmas = (double)grid * block * (double)iters * 64.0;
  *flops_out = 2.0 * fmas / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return HEXO_OK;
}

}  // extern "C"
