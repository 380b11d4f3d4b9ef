// hestonexotics_b200/csrc/path_kernel_il.cuh
//
// K1, "interleaved look-ahead" variant.  Same arithmetic, streams and results as
// heston_qe_paths_kernel (path_kernel.cuh); what changes is WHEN the normals are made.
// The uniform kernel alternates between two phases per thread -- 8 steps of FP64 work, then
// a burst of integer/FP32 work that turns the next generator round into normals -- so at any
// moment some warps queue for the FP64 pipe while others leave it idle.  Here the central part
// of the inverse-normal transform of round r+1 is spread over the 8 steps of round r: every
// step iteration consumes one (Z_V, Z_X) pair of the current round and produces one pair of
// the next round.  The FP32/integer instructions are independent of the FP64 dependency chain
// of the step, so they issue in its latency shadows.  Only the generator round itself and the
// per-lane tail loop remain as a separate phase at round boundaries.
//
// Shared memory (T threads): wring [2][8][T] raw word pairs, zring [2][8][T] normal pairs
// (double-buffered: slot r&1 holds round r), exptab, fvbuf, acc as in path_kernel.cuh.
#pragma once
#include <stdint.h>

#include <type_traits>

#include "path_kernel.cuh"

namespace hexo {

__host__ __device__ inline size_t path_kernel_il_smem(int block, uint32_t n_opts, int normal_mode,
                                                      bool acc_in_smem) {
  const int warps = block / 32;
  const size_t zb = normal_mode == HEXO_NORMAL_F64 ? 16 : 8;
  return 2 * zb * kStepsPerRound * block + (size_t)2 * 16 * kStepsPerRound * block + 32 * 8 +
         (size_t)32 * 8 * warps + (acc_in_smem ? (size_t)warps * 2 * n_opts * 8 : 0);
}

template <int PAYOFF, int NORMAL_MODE, bool INLINE_SEGS>
__global__ void __launch_bounds__(kMaxBlock, kMinBlocksPerSM)
heston_qe_paths_il_kernel(const __grid_constant__ PathArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = blockDim.x, nwarps = T >> 5;
  using Ring = ZRing<NORMAL_MODE>;
  constexpr bool kAsian = PAYOFF == HEXO_PAYOFF_ASIAN;

  unsigned char* sp = smem_raw;
  const uint32_t zstride = pin32(Ring::kBytesPerStep * T), wstride = pin32(16 * T);
  const uint32_t wcol = pin32(smem_addr(sp) + 16 * tid);  // slot 0; slot 1 is 8*wstride further
  sp += (size_t)2 * 16 * kStepsPerRound * T;
  const uint32_t zcol = pin32(smem_addr(sp) + Ring::kBytesPerStep * tid);
  sp += (size_t)2 * Ring::kBytesPerStep * kStepsPerRound * T;
  double* exptab = reinterpret_cast<double*>(sp);
  const uint32_t exptab_s = pin32(smem_addr(sp));
  sp += 32 * 8;
  double* fvbuf = reinterpret_cast<double*>(sp) + 32 * warp;
  sp += (size_t)32 * 8 * nwarps;
  double* acc_all = a.gacc ? a.gacc + (size_t)blockIdx.x * nwarps * 2 * a.n_opts
                           : reinterpret_cast<double*>(sp);
  double* my_sum = acc_all + (size_t)warp * 2 * a.n_opts;  // lane-owned slots
  double* my_sq = my_sum + a.n_opts;
  if (!a.gacc)
    for (uint32_t j = lane; j < 2 * a.n_opts; j += 32) my_sum[j] = 0.0;
  exp_table_init(exptab, tid, T);

  const uint64_t slot = (uint64_t)blockIdx.x * T + tid;
  const uint64_t sid = a.stream_begin + slot;
  const uint64_t my_paths =
      slot < a.stream_count ? a.base_paths + (sid < a.rem_streams ? 1u : 0u) : 0u;
  const uint64_t warp_paths = __shfl_sync(0xffffffffu, my_paths, 0);

  // ---- ring state -------------------------------------------------------------------------
  Shishua rng;
  const uint32_t wslot = 8 * wstride, zslot = 8 * zstride;
  auto store_words = [&](const uint64_t (&o)[16], uint32_t wc) {
#pragma unroll
    for (int s = 0; s < kStepsPerRound; ++s) sts_b64x2(wc + s * wstride, o[2 * s], o[2 * s + 1]);
  };
  // running addresses: (za, wa) = this step of the current round, (zn, wn) = same step of the
  // next round; `tails` collects the tail bits of the next round
  uint32_t za, wa, zn, wn, pos = 0, cur = 0, tails = 0;
  {
    uint64_t o[16];
    rng.init(a.seed, sid, 0, 0, o);
    store_words(o, wcol);
    Ring::fill(o, wcol, wstride, zcol, zstride);  // round 0: central + tails, all at once
    rng.round(o);
    store_words(o, wcol + wslot);                 // round 1: raw words only
  }
  za = zcol, wa = wcol, zn = zcol + zslot, wn = wcol + wslot;
  // round boundary: finish the next round (tails), make it current, generate the one after
  auto boundary = [&]() {
    const uint32_t nxt = cur ^ 1u;
    Ring::tail_phase(tails, wcol + nxt * wslot, wstride, zcol + nxt * zslot, zstride);
    tails = 0;
    uint64_t o[16];
    rng.round(o);
    store_words(o, wcol + cur * wslot);  // the slot of the round just consumed
    cur = nxt;
    pos = 0;
    za = zcol + cur * zslot, wa = wcol + cur * wslot;
    zn = zcol + (cur ^ 1u) * zslot, wn = wcol + (cur ^ 1u) * wslot;
  };
  // normals of the current step + central transform of the same step of the next round
  auto fetch = [&](double& zv, double& zx, uint32_t& ua) {
    if (pos == kStepsPerRound) boundary();
    Ring::get(za, zv, zx);
    ua = wa;
    tails |= Ring::central_step(wn, zn) << (2 * pos);
    ++pos;
    za += zstride, wa += wstride, zn += zstride, wn += wstride;
  };
  __syncthreads();  // exptab

  for (uint64_t p = 0; p < warp_paths; ++p) {
    const bool active = p < my_paths;
    // HQEAnderson::operator=(initial_state), HSimulation.tpp:26,87-94
    double V = a.v0, lnX = a.lnS, X = a.S, Xprev = a.S;
    double integral = 0.0;  // AAsianCallNonAdaptive::accumulated_value, reset per path (:34)
    for (uint32_t k = 0; k < a.n_seg; ++k) {
      SegConst g = INLINE_SEGS ? a.seg_inline[k] : a.segs[k];
      g.D = pin(g.D); g.m0 = pin(g.m0); g.c1h = pin(g.c1h); g.c2h = pin(g.c2h);
      g.K0 = pin(g.K0); g.K1 = pin(g.K1); g.K2 = pin(g.K2); g.K3 = pin(g.K3);
      if (active) {
        const uint32_t n = g.n_steps;
        if (kAsian && k > 0 && n > 0) integral += g.hcarry * (X + Xprev);  // HSimulation.tpp:42-44
        const double Xa = X;
        double sumX = 0.0;
        auto spot_half = [&](double Vfrom, double Vto, double zx, auto with_x) {
          const double delta = qe_logreturn(g, Vfrom, Vto, zx);
          if (kAsian) {
            Xprev = X;
            X = grow_spot(X, delta, exptab_s);
            sumX += X;
          } else {
            lnX += delta;
            if (decltype(with_x)::value) {
              Xprev = X;
              X = fast_exp(lnX, exptab_s);
            }
          }
        };
        // `count` (> 0) steps, software-pipelined as in path_kernel.cuh
        auto run = [&](uint32_t count, auto with_x) {
          double zv, zx_pend;
          uint32_t ua;
          fetch(zv, zx_pend, ua);
          double Vold = V;
          V = qe_variance(g, Vold, zv, [ua]() { return u64_to_unit(lds_b64(ua)); });
          for (uint32_t i = 1; i < count; ++i) {
            double zx;
            fetch(zv, zx, ua);
            spot_half(Vold, V, zx_pend, with_x);
            const double Vn = qe_variance(g, V, zv, [ua]() { return u64_to_unit(lds_b64(ua)); });
            Vold = V;
            V = Vn;
            zx_pend = zx;
          }
          spot_half(Vold, V, zx_pend, with_x);
        };
        if (kAsian) {
          if (n > 0) run(n, std::true_type{});
          if (n > 0) integral += g.h * 0.5 * (Xa - Xprev + 2.0 * (sumX - X));
        } else {
          if (n > 2) run(n - 2, std::false_type{});
          if (n > 0) run(min(n, 2u), std::true_type{});
        }
      }
      // accumulate_final_value, AsianContract.h:29-34 / VanillaContract.h:28-31
      const double dx = X - Xprev;
      const double fv =
          kAsian ? (integral + dx * g.w + g.hs * (X + Xprev)) / g.expiry : Xprev + dx * g.w;
      __syncwarp();
      fvbuf[lane] = fv;
      const unsigned amask = __ballot_sync(0xffffffffu, active);
      __syncwarp();
      for (uint32_t j = lane; j < g.n_strikes; j += 32) {
        const double K = __ldg(a.strikes + g.first_opt + j);
        double s = 0.0, q = 0.0;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
          if ((amask >> l) & 1u) {
            const double pf = fmax(fvbuf[l] - K, 0.0);
            s += pf;
            q = fma(pf, pf, q);
          }
        }
        my_sum[g.first_opt + j] += s;
        my_sq[g.first_opt + j] += q;
      }
    }
  }

  __syncthreads();
  const uint32_t n2 = 2 * a.n_opts;
  for (uint32_t j = tid; j < n2; j += T) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += acc_all[(size_t)w * n2 + j];
    a.partials[(size_t)blockIdx.x * n2 + j] = s;
  }
}

}  // namespace hexo
