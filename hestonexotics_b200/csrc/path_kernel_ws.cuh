// hestonexotics_b200/csrc/path_kernel_ws.cuh
//
// K1, warp-specialised variant.  Same arithmetic, same streams and same results as
// heston_qe_paths_kernel (path_kernel.cuh), different division of labour: the first half
// of a block's warps are CONSUMERS that only run the FP64 part (QE stepper + payoffs), the
// second half are PRODUCERS that only run the integer / FP32 part (shishua rounds + inverse
// normals).  Producer warp w feeds consumer warp w lane by lane through a two-slot ring in
// shared memory, hand-shaken with mbarriers.  The two instruction streams lean on different
// pipes (FP64 vs FMA/ALU/XU), so mixing them on one SM sub-partition fills issue slots that a
// uniform kernel leaves empty while all its warps queue for the same half-rate pipe.
#pragma once
#include <stdint.h>

#include <type_traits>

#include "path_kernel.cuh"

namespace hexo {

#ifndef HEXO_WS_RATIO
#define HEXO_WS_RATIO 1
#endif
constexpr int kWsRatio = HEXO_WS_RATIO;   // consumer warps fed by one producer warp
constexpr int kWsProducerWarps = 2;
constexpr int kWsConsumerWarps = kWsProducerWarps * kWsRatio;
constexpr int kWsBlock = 32 * (kWsProducerWarps + kWsConsumerWarps);  // ratio 2: 192 threads
constexpr int kWsConsumers = kWsConsumerWarps * 32;
// blocks per SM the register budget is cut for: 65536 / (kWsBlock * this) registers per thread
constexpr int kWsMinBlocks = kWsRatio == 1 ? 4 : kWsRatio == 2 ? 3 : 2;

// ---- mbarrier helpers (shared-space addresses) ---------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(addr)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n"
      " bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(addr), "r"(parity)
      : "memory");
}

// Shared memory (C = consumer threads per block, W = consumer warps):
//   zring [2][8][C] (Z_V, Z_X) pairs of two generator rounds, float2 / double2
//   uring [2][8][C] raw variance words of the same rounds (uniform of the psi >= 1.5 branch)
//   oddw  [8][P]    raw spot words of the round being produced (tail phase of the normals),
//                   P = producer threads
//   sring [8][C]    shishua state of every stream, parked between rounds (a producer thread
//                   serves kWsRatio streams in turn)
//   exptab[32], bars[W][4] (full0, full1, empty0, empty1), fvbuf[W][32], acc[W][2][n_opts]
__host__ __device__ inline size_t path_kernel_ws_smem(uint32_t n_opts, int normal_mode,
                                                      bool acc_in_smem) {
  const int C = kWsConsumers, W = C / 32;
  const size_t zb = normal_mode == HEXO_NORMAL_F64 ? 16 : 8;
  return 2 * zb * 8 * C + 2 * 8 * 8 * C + 8 * 8 * (kWsProducerWarps * 32) + 16 * 8 * C + 32 * 8 +
         (size_t)W * 4 * 8 +
         (size_t)32 * 8 * W + (acc_in_smem ? (size_t)W * 2 * n_opts * 8 : 0);
}

// producer side of ZRing: same two phases as ZRing<>::fill, but the raw words of a step live in
// two places (variance word in the consumer-visible uring slot, spot word in oddw)
template <int NORMAL_MODE>
__device__ __forceinline__ void ws_fill(const uint64_t (&o)[16], uint32_t ucol, uint32_t ustride,
                                        uint32_t ocol, uint32_t ostride, uint32_t zcol,
                                        uint32_t zstride) {
  uint32_t tails = 0;
  if (NORMAL_MODE == HEXO_NORMAL_F32) {
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      float zv, zx;
      bool t0, t1;
      normal2_central_f32(o[2 * s], o[2 * s + 1], zv, zx, t0, t1);
      sts_b64(zcol + s * zstride, pack2(zv, zx));
      if (t0) tails |= 1u << (2 * s);
      if (t1) tails |= 2u << (2 * s);
    }
    while (tails) {
      const int j0 = __ffs(tails) - 1;
      tails &= tails - 1;
      const bool two = tails != 0;
      const int j1 = two ? __ffs(tails) - 1 : j0;
      tails &= tails - 1;
      const uint64_t w0 = lds_b64((j0 & 1) ? ocol + (j0 >> 1) * ostride : ucol + (j0 >> 1) * ustride);
      const uint64_t w1 = lds_b64((j1 & 1) ? ocol + (j1 >> 1) * ostride : ucol + (j1 >> 1) * ustride);
      float t0, t1;
      float z0 = normal_tail_mid_f32(w0, t0), z1 = normal_tail_mid_f32(w1, t1);
      if (fmaxf(t0, t1) > 25.0f) {  // far tail: essentially never
        if (t0 > 25.0f) z0 = normal_tail_far_f32(w0, t0);
        if (t1 > 25.0f) z1 = normal_tail_far_f32(w1, t1);
      }
      sts_f32(zcol + (j0 >> 1) * zstride + (j0 & 1) * 4, z0);
      if (two) sts_f32(zcol + (j1 >> 1) * zstride + (j1 & 1) * 4, z1);
    }
  } else {
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      bool t0, t1;
      const double zv = normal_central_f64(o[2 * s], t0);
      const double zx = normal_central_f64(o[2 * s + 1], t1);
      sts_f64x2(zcol + s * zstride, zv, zx);
      if (t0) tails |= 1u << (2 * s);
      if (t1) tails |= 2u << (2 * s);
    }
    while (tails) {
      const int j = __ffs(tails) - 1;
      tails &= tails - 1;
      const uint64_t w = lds_b64((j & 1) ? ocol + (j >> 1) * ostride : ucol + (j >> 1) * ustride);
      sts_f64(zcol + (j >> 1) * zstride + (j & 1) * 8, normal_tail_f64(w));
    }
  }
}

template <int PAYOFF, int NORMAL_MODE, bool INLINE_SEGS>
__global__ void __launch_bounds__(kWsBlock, kWsMinBlocks)
heston_qe_paths_ws_kernel(const __grid_constant__ PathArgs a, const uint32_t steps_per_path) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int C = kWsConsumers, W = C / 32;
  constexpr bool kAsian = PAYOFF == HEXO_PAYOFF_ASIAN;
  constexpr uint32_t ZB = NORMAL_MODE == HEXO_NORMAL_F64 ? 16 : 8;
  using Ring = ZRing<NORMAL_MODE>;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp >= W;
  // consumer: its own warp index; producer: the first consumer warp it feeds
  const int cw = producer ? (warp - W) * kWsRatio : warp;
  const int c = cw * 32 + lane;               // consumer thread index within the block

  unsigned char* sp = smem_raw;
  const uint32_t zstride = ZB * C, ustride = 8 * C;
  const uint32_t zslot = ZB * 8 * C, uslot = 8 * 8 * C;  // bytes per ring slot
  const uint32_t zbase = smem_addr(sp) + ZB * c;
  sp += (size_t)2 * ZB * 8 * C;
  const uint32_t ubase = smem_addr(sp) + 8 * c;
  sp += (size_t)2 * 8 * 8 * C;
  constexpr int PT = kWsProducerWarps * 32;  // producer threads
  const uint32_t ostride = 8 * PT;
  const uint32_t obase = smem_addr(sp) + 8 * (producer ? (warp - W) * 32 + lane : 0);
  sp += (size_t)8 * 8 * PT;
  const uint32_t sbase = smem_addr(sp) + 16 * c;  // parked generator states, stride 16 C
  sp += (size_t)16 * 8 * C;
  double* exptab = reinterpret_cast<double*>(sp);
  const uint32_t exptab_s = smem_addr(sp);
  sp += 32 * 8;
  const uint32_t bars0 = smem_addr(sp);
  const uint32_t bars = bars0 + cw * 32;  // full0, full1, empty0, empty1
  sp += (size_t)W * 4 * 8;
  double* fvbuf = reinterpret_cast<double*>(sp) + 32 * cw;
  sp += (size_t)32 * 8 * W;
  double* acc_all = a.gacc ? a.gacc + (size_t)blockIdx.x * W * 2 * a.n_opts
                           : reinterpret_cast<double*>(sp);
  double* my_sum = acc_all + (size_t)cw * 2 * a.n_opts;
  double* my_sq = my_sum + a.n_opts;
  if (!producer && !a.gacc)
    for (uint32_t j = lane; j < 2 * a.n_opts; j += 32) my_sum[j] = 0.0;
  exp_table_init(exptab, tid, kWsBlock);
  if (tid < W) {
    const uint32_t b = bars0 + tid * 32;
    mbar_init(b + 0, 32);
    mbar_init(b + 8, 32);
    mbar_init(b + 16, 32);
    mbar_init(b + 24, 32);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  // paths of consumer thread (consumer warp cw + q, this lane); q = 0 for a consumer itself
  auto paths_of = [&](int q, uint64_t& sid_out) {
    const uint64_t slot_id = (uint64_t)blockIdx.x * C + c + 32 * q;
    sid_out = a.stream_begin + slot_id;
    return slot_id < a.stream_count ? a.base_paths + (sid_out < a.rem_streams ? 1u : 0u) : 0u;
  };
  uint64_t sid;
  const uint64_t my_paths = paths_of(0, sid);
  const uint64_t warp_paths = __shfl_sync(0xffffffffu, my_paths, 0);

  if (producer) {
    // ------------------------------------------------------------------ producer warps
    // rounds each of the kWsRatio consumer warps will consume
    uint64_t rounds_q[kWsRatio], max_rounds = 0;
#pragma unroll
    for (int q = 0; q < kWsRatio; ++q) {
      uint64_t sq;
      const uint64_t wp = __shfl_sync(0xffffffffu, paths_of(q, sq), 0);
      rounds_q[q] = (wp * steps_per_path + kStepsPerRound - 1) / kStepsPerRound;
      max_rounds = max(max_rounds, rounds_q[q]);
    }
    auto publish = [&](int q, uint64_t r, const uint64_t (&o)[16]) {
      const uint32_t s = (uint32_t)r & 1u, bq = bars + 32 * q;
      if (r >= 2) mbar_wait(bq + 16 + 8 * s, (uint32_t)((r >> 1) - 1) & 1u);  // empty[s]
      const uint32_t ucol = ubase + 8 * 32 * q + s * uslot, zcol = zbase + ZB * 32 * q + s * zslot;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sts_b64(ucol + i * ustride, o[2 * i]);
        sts_b64(obase + i * ostride, o[2 * i + 1]);
      }
      ws_fill<NORMAL_MODE>(o, ucol, ustride, obase, ostride, zcol, zstride);
      mbar_arrive(bq + 8 * s);  // full[s]; release semantics order the stores above
    };
    uint32_t rounds_done = 0;  // generator rounds completed so far (same for every stream)
    // round 0: seed every stream, publish its first 16 words, park its state
#pragma unroll 1
    for (int q = 0; q < kWsRatio; ++q) {
      Shishua rng;
      uint64_t o[16], sq;
      paths_of(q, sq);
      rng.init(a.seed, sq, 0, 0, o);
      rounds_done = rng.rounds;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        sts_b64x2(sbase + 16 * 32 * q + i * 16 * C, rng.s[2 * i], rng.s[2 * i + 1]);
      if (rounds_q[q] > 0) publish(q, 0, o);
    }
    for (uint64_t r = 1; r < max_rounds; ++r) {
#pragma unroll 1
      for (int q = 0; q < kWsRatio; ++q) {
        if (r >= rounds_q[q]) continue;
        Shishua rng;
        uint64_t o[16];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          lds_b64x2(sbase + 16 * 32 * q + i * 16 * C, rng.s[2 * i], rng.s[2 * i + 1]);
        rng.rounds = rounds_done;
        rng.round(o);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          sts_b64x2(sbase + 16 * 32 * q + i * 16 * C, rng.s[2 * i], rng.s[2 * i + 1]);
        publish(q, r, o);
      }
      ++rounds_done;
    }
  } else {
    // ------------------------------------------------------------------ consumer warps
    uint64_t r = 0;            // next round to acquire
    uint32_t pos = kStepsPerRound;
    uint32_t ucol = 0, zcol = 0;
    bool holding = false;
    auto next_round = [&]() {
      if (holding) mbar_arrive(bars + 16 + 8 * ((uint32_t)(r - 1) & 1u));  // empty[slot]
      const uint32_t s = (uint32_t)r & 1u;
      mbar_wait(bars + 8 * s, (uint32_t)(r >> 1) & 1u);                    // full[slot]
      ucol = ubase + s * uslot;
      zcol = zbase + s * zslot;
      holding = true;
      ++r;
      pos = 0;
    };
    for (uint64_t p = 0; p < warp_paths; ++p) {
      const bool active = p < my_paths;  // inactive lanes still step (uniform consumption)
      double V = a.v0, lnX = a.lnS, X = a.S, Xprev = a.S;
      double integral = 0.0;
      for (uint32_t k = 0; k < a.n_seg; ++k) {
        SegConst g = INLINE_SEGS ? a.seg_inline[k] : a.segs[k];
        g.D = pin(g.D); g.m0 = pin(g.m0); g.c1h = pin(g.c1h); g.c2h = pin(g.c2h);
        g.K0 = pin(g.K0); g.K1 = pin(g.K1); g.K2 = pin(g.K2); g.K3 = pin(g.K3);
        const uint32_t n = g.n_steps;
        if (kAsian && k > 0 && n > 0) integral += g.hcarry * (X + Xprev);
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
        auto run = [&](uint32_t count, auto with_x) {
          double Vold = V, zx_pend = 0.0;
          bool first = true;
          while (count) {
            if (pos == kStepsPerRound) next_round();
            uint32_t m = min(kStepsPerRound - pos, count);
            count -= m;
            uint32_t za = zcol + pos * zstride, ua = ucol + pos * ustride;
            pos += m;
            if (first) {
              first = false;
              double zv;
              Ring::get(za, zv, zx_pend);
              Vold = V;
              V = qe_variance(g, Vold, zv, [ua]() { return u64_to_unit(lds_b64(ua)); });
              --m, za += zstride, ua += ustride;
            }
            for (; m; --m, za += zstride, ua += ustride) {
              double zv, zx;
              Ring::get(za, zv, zx);
              spot_half(Vold, V, zx_pend, with_x);
              const double Vn = qe_variance(g, V, zv, [ua]() { return u64_to_unit(lds_b64(ua)); });
              Vold = V;
              V = Vn;
              zx_pend = zx;
            }
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
  }

  // consumer warps -> block partial, fixed order
  __syncthreads();
  const uint32_t n2 = 2 * a.n_opts;
  for (uint32_t j = tid; j < n2; j += kWsBlock) {
    double s = 0.0;
    for (int w = 0; w < W; ++w) s += acc_all[(size_t)w * n2 + j];
    a.partials[(size_t)blockIdx.x * n2 + j] = s;
  }
}

}  // namespace hexo
