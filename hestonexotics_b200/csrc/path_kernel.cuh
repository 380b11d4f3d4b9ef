// hestonexotics_b200/csrc/path_kernel.cuh
//
// K1: the fused path kernel.  Replaces the reference's HSimulation::price<Scheme>
// (src/HSimulation.tpp:10-51) and everything it calls per draw: the shishua
// wrapper (src/RNG.cpp), PPND16 (src/as241.f90), the QE stepper
// (HSimulation.tpp:52-86) and the payoff policies (src/inc/AsianContract.h,
// src/inc/VanillaContract.h).
//
//  * one thread = one independent shishua stream (seed {seed, stream, 0, 0});
//    RNG state, variance, log-spot and the running integral stay in registers;
//    no path data touches HBM;
//  * every ring_steps() steps (16 with F32 normals, 8 with F64) a thread advances its
//    generator (one round = 16 words = 8 steps), turns the words into (Z_V, Z_X) pairs
//    in one batch (normals.cuh) and parks them in its own column of shared memory; the
//    step loop reads one pair per step;
//  * when a maturity is reached the 32 final values of a warp are exchanged
//    through shared memory and each lane owns a strided subset of the strikes,
//    so per-option sums are accumulated without atomics and in a fixed order;
//  * warps -> block partials (shared memory), blocks -> sums (second tiny
//    kernel), both in fixed order: results are reproducible for a given launch
//    geometry.
#pragma once
#include <stdint.h>

#include <type_traits>

#include "../../include/hexo_gpu.h"
#include "normals.cuh"
#include "qe.cuh"
#include "philox.cuh"
#include "shishua.cuh"

namespace hexo {

constexpr int kMaxBlock = 256;
#ifndef HEXO_MIN_BLOCKS
#define HEXO_MIN_BLOCKS 2
#endif
constexpr int kMinBlocksPerSM = HEXO_MIN_BLOCKS;  // register budget: 65536 / (256 * this)
constexpr int kStepsPerRound = 8;  // one shishua round = 16 words = 8 steps
// Steps a thread transforms per refill: two generator rounds.  The tail phase then runs once per
// 32 draws of a lane (154 +- 11 tail draws per warp, dealt out evenly over the lanes) and the
// refill overhead is paid half as often as with one round.
__host__ __device__ constexpr int ring_steps(int /*normal_mode*/) { return 2 * kStepsPerRound; }
// Shared memory of a thread's ring per step.  F32 modes: the raw variance and spot words (two
// planes of 8 bytes) and the pair (Z_V, Z_X) as float2.  F64 mode: the pair is a double2, so only
// the variance words get a plane of their own (the psi >= 1.5 branch reads them); the spot word
// of a TAIL draw waits in its own 8-byte z slot until the tail phase replaces it by the normal.
// That keeps a 16-step ring at 24 bytes per step and thread in both cases: two blocks per SM.
__host__ __device__ constexpr int ring_raw_planes(int normal_mode) {
  return normal_mode == HEXO_NORMAL_F64 ? 1 : 2;
}
__host__ __device__ constexpr int ring_z_bytes(int normal_mode) {
  return normal_mode == HEXO_NORMAL_F64 ? 16 : 8;
}
// F64 mode: the logarithm of the tail formula reads a 128-entry table (fast_log_tab) from shared
// memory -- 32 lanes with 32 different indices would serialise on the constant bank
__host__ __device__ constexpr int ring_logtab_bytes(int normal_mode) {
  return normal_mode == HEXO_NORMAL_F64 ? 128 * 16 : 0;
}
constexpr int kInlineSegs = 8;     // maturities whose constants travel as kernel parameters
// Entries of a warp's tail list (16 bits each).  A refill of 16 steps has 32 x 32 draws per
// warp, 15 % of them tails: 154 +- 11.  A refill with more tail draws than the list holds takes
// the per-lane loop instead.
constexpr int kTailListCap = 256;
constexpr int kTailListBytes = 2 * kTailListCap;
#ifndef HEXO_STEP_UNROLL
#define HEXO_STEP_UNROLL 2  // the loop-carried rotation (Vold <- V <- V') needs an even count
#endif
constexpr int kStepUnroll = HEXO_STEP_UNROLL;
#ifndef HEXO_TAIL_COOP
#define HEXO_TAIL_COOP 1  // 0: every lane loops over its own tail draws (the round-1 scheme)
#endif


struct PathArgs {
  double v0, S, lnS;
  uint64_t seed;
  uint64_t stream_begin;  // first global stream id of this launch
  uint64_t stream_count;  // streams in this launch (one per thread)
  uint64_t base_paths;    // every stream runs base_paths paths ...
  uint64_t rem_streams;   // ... and global streams < rem_streams one more
  uint32_t n_seg, n_opts;
  const SegConst* segs;                // all n_seg segments (device memory)
  SegConst seg_inline[kInlineSegs];    // the first min(n_seg, kInlineSegs) again, in the
                                       // kernel's constant bank: uniform operands, no registers
  const double* strikes;
  // Accumulators per warp (= sums per launch), n_acc.  Plain: [sum pf | sum pf^2] = 2 n_opts.
  // With the control variate (template parameter CV) the control c = final value - S (zero
  // mean: the spot is a martingale, r = 0) adds [sum pf c] per option and [sum c | sum c^2] per
  // maturity: 3 n_opts + 2 n_seg.
  double* partials;  // [gridDim.x][n_acc]
#ifdef HEXO_DEV_PROBES
  uint32_t dev_no_refill;  // development build only (wrong prices!).  HEXO_NO_REFILL=1: the step
                           // loop alone (ring reused with flipped signs); 2: generator and
                           // normal transform alone (the step loop only sums the normals)
#endif
  double* gacc;      // nullptr: per-warp accumulators in shared memory; else zero-initialised
                     // [gridDim.x][warps][n_acc] in device memory (large option chains,
                     // where shared-memory accumulators would cost occupancy)
};

// Shared memory of the path kernel (per block), T = threads, W = warps, R = ring_steps:
//   wring  [P][R][T] the raw words of the R steps a refill covers, P = ring_raw_planes: variance
//                 words first, then (F32 modes) spot words.  Kept because the tail phase of
//                 the normal transform re-reads them and because the psi >= 1.5
//                 branch needs the UNIFORM of the variance draw (HSimulation.tpp:72)
//   zring  [R][T] pairs (Z_V, Z_X) of the refill, float2 (F32 modes) / double2 (F64)
//   logtab [128]  F64 mode only: (1/c_i, ln c_i) of fast_log_tab
//   exptab [32]   2^(j/32)
//   fvbuf  [W][32] final values of a warp at a maturity
//   tlist  [W][kTailListCap] 16-bit entries: the warp's tail draws of a refill (tail_phase_coop)
//   acc    [W][n_acc] lane-owned payoff sums / sums of squares (/ control-variate sums)
__host__ __device__ inline size_t ring_smem(int block, int normal_mode) {
  return (size_t)(8 * ring_raw_planes(normal_mode) + ring_z_bytes(normal_mode)) *
             ring_steps(normal_mode) * block +
         ring_logtab_bytes(normal_mode);
}
__host__ __device__ inline size_t path_kernel_smem(int block, uint32_t n_acc, int normal_mode,
                                                   bool acc_in_smem = true) {
  const int warps = block / 32;
  return ring_smem(block, normal_mode) + 32 * 8 + (size_t)32 * 8 * warps +
         (size_t)kTailListBytes * warps + (acc_in_smem ? (size_t)warps * n_acc * 8 : 0);
}

// ---- shared-space accessors (32-bit addresses: no generic-pointer arithmetic
// in the step loop) -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void sts_b64(uint32_t addr, uint64_t v) {
  asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v));
}
__device__ __forceinline__ uint64_t lds_b64(uint32_t addr) {
  uint64_t v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_b64x2(uint32_t addr, uint64_t a, uint64_t b) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b));
}
__device__ __forceinline__ void lds_b64x2(uint32_t addr, uint64_t& a, uint64_t& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float a) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a));
}
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b));
}
__device__ __forceinline__ void lds_f64x2(uint32_t addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double a) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(a));
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double a;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(addr));
  return a;
}

// makes a value opaque to the compiler and pins it in a register
__device__ __forceinline__ double pin(double x) {
  asm volatile("" : "+d"(x));
  return x;
}

__device__ __forceinline__ uint32_t pin32(uint32_t x) {
  asm volatile("" : "+r"(x));
  return x;
}

// index of the highest set bit (x != 0): one FLO
__device__ __forceinline__ uint32_t bfind32(uint32_t x) {
  uint32_t r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}

__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v));
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// Shared-memory ring of one thread (addresses in the shared window).  Raw words live in planes
// [variance words | spot words] of RING steps each, so word j of a refill (j = RING d + step,
// d = 0 variance / 1 spot) sits at ucol + j ustride; the (Z_V, Z_X) pair of step s sits at
// zcol + s zstride.  Lane l of a warp owns column l: ucol = (lane 0's ucol) + 8 l.  (F64 mode has
// no spot plane: the spot word of a tail draw sits in the second half of its step's z pair.)
struct RingAddr {
  uint32_t ucol, ustride, zcol, zstride;
  uint32_t tlist;   // this warp's tail list (kTailListBytes), see tail_phase_coop
  uint32_t logtab;  // F64 mode: the block's table of fast_neglog_tab
  uint32_t ublock, zblock;  // thread 0's ucol / zcol (block-uniform: they can live in uniform
                            // registers, which costs the tail loop nothing)
};
// carves the ring out of the block's dynamic shared memory at `base` (ring_smem bytes)
template <int NORMAL_MODE>
__device__ __forceinline__ RingAddr ring_addr(uint32_t base, int tid, int T, uint32_t tlist) {
  constexpr uint32_t kP = ring_raw_planes(NORMAL_MODE), kZ = ring_z_bytes(NORMAL_MODE);
  constexpr uint32_t kR = ring_steps(NORMAL_MODE);
  RingAddr ra;
  ra.ucol = base + 8 * tid;
  ra.ustride = 8 * T;
  ra.zcol = base + 8 * kP * kR * T + kZ * tid;
  ra.zstride = kZ * T;
  ra.tlist = tlist;
  ra.logtab = base + (8 * kP + kZ) * kR * T;
  ra.ublock = base;
  ra.zblock = base + 8 * kP * kR * T;
  return ra;
}
// fills the logarithm table of the F64 tail formula (no-op in the F32 modes); the block must
// synchronise before the first refill
template <int NORMAL_MODE>
__device__ __forceinline__ void ring_logtab_init(unsigned char* smem_base, int tid, int T) {
  if (ring_logtab_bytes(NORMAL_MODE) == 0) return;
  constexpr size_t kOff = (size_t)(8 * ring_raw_planes(NORMAL_MODE) + ring_z_bytes(NORMAL_MODE)) *
                          ring_steps(NORMAL_MODE);
  LogTabEntry* tab = reinterpret_cast<LogTabEntry*>(smem_base + kOff * T);
  for (int j = tid; j < 128; j += T) tab[j] = kLogTab[j];
}

// Raw words -> normals in two phases (normals.cuh): the central formula for all draws of a
// generator round, then the tail draws of the whole refill.
//   central_round_planar : one generator round (8 steps); shifts the draws' tail flags into two
//                          accumulators (ring_refill turns them into the mask whose bit j
//                          marks word j of the refill)
//   tail_phase_planar    : every lane loops over its OWN tail draws, two per iteration
//   tail_pair            : two tail draws -> their slots of the z ring
template <int NORMAL_MODE>
struct ZRing;

template <bool P7>
struct ZRingF32 {
  static constexpr int kBytesPerStep = 8;  // float2 (Z_V, Z_X)
  static constexpr int kRawPlanes = 2;     // variance and spot words
  static constexpr bool kTailPairs = true; // the tail formula takes two draws per packed chain
  // One generator round.  The tail flag of a draw is the SIGN BIT of a word the central formula
  // hands back (r = 0.180625 - q^2, shifted for the as-built mode: normal2_central_f32): each
  // flag is shifted into an accumulator with one funnel shift (no compare, no select); `tv`
  // collects the variance draws' flags, `tx` the spot draws', in step order -- ring_refill turns
  // the two into the mask the tail phase reads.
  static __device__ __forceinline__ void central_round_planar(const uint64_t (&o)[16], int step0,
                                                              uint32_t zcol, uint32_t zstride,
                                                              uint32_t& tv, uint32_t& tx) {
#pragma unroll
    for (int s = 0; s < kStepsPerRound; ++s) {
      float zv, zx;
      uint32_t r0, r1;
      normal2_central_f32<P7>(o[2 * s], o[2 * s + 1], zv, zx, r0, r1);
      sts_b64(zcol + (step0 + s) * zstride, pack2(zv, zx));
      tv = __funnelshift_l(r0, tv, 1);
      tx = __funnelshift_l(r1, tx, 1);
    }
  }
  // two tail draws (words w0, w1; `two` false: w1 is w0 again) -> z ring slots za0, za1
  static __device__ __forceinline__ void tail_pair(uint64_t w0, uint64_t w1, bool two, uint32_t za0,
                                                   uint32_t za1) {
    float t0, t1, z0, z1;
    normal2_tail_mid_f32<P7>(w0, w1, z0, z1, t0, t1);
    if (fmaxf(t0, t1) > 25.0f) {  // far tail: essentially never
      if (t0 > 25.0f) z0 = normal_tail_far_f32(w0, t0);
      if (t1 > 25.0f) z1 = normal_tail_far_f32(w1, t1);
    }
    sts_f32(za0, z0);
    if (two) sts_f32(za1, z1);
  }
  static __device__ __forceinline__ void get(uint32_t addr, double& zv, double& zx) {
    float a, b;
    unpack2(lds_b64(addr), a, b);
    zv = (double)a;
    zx = (double)b;
  }
};

template <>
struct ZRing<HEXO_NORMAL_F32> : ZRingF32<false> {};
template <>
struct ZRing<HEXO_NORMAL_F32_PPND7> : ZRingF32<true> {};

template <>
struct ZRing<HEXO_NORMAL_F64> {
  static constexpr int kBytesPerStep = 16;   // double2 (Z_V, Z_X)
  static constexpr int kRawPlanes = 1;       // variance words only, see ring_raw_planes
  static constexpr bool kTailPairs = false;  // one tail draw per lane and iteration
  static __device__ __forceinline__ void central_round_planar(const uint64_t (&o)[16], int step0,
                                                              uint32_t zcol, uint32_t zstride,
                                                              uint32_t& tv, uint32_t& tx) {
#pragma unroll
    for (int s = 0; s < kStepsPerRound; ++s) {
      uint32_t r0, r1;  // high words of r = 0.180625 - q^2: the sign bit is the tail flag
      double zv, zx;
      normal2_central_f64(o[2 * s], o[2 * s + 1], zv, zx, r0, r1);
      const uint32_t za = zcol + (step0 + s) * zstride;
      sts_f64x2(za, zv, zx);
      // a spot draw outside the central region: park its raw word where its normal will go
      if ((int32_t)r1 < 0) sts_b64(za + 8, o[2 * s + 1]);
      tv = __funnelshift_l(r0, tv, 1);
      tx = __funnelshift_l(r1, tx, 1);
    }
  }
  // one tail draw (word w) -> its z ring slot
  static __device__ __forceinline__ void tail_one(uint64_t w, uint32_t za, uint32_t logtab) {
    double r;
    double z = normal_tail_mid_f64(w, r, logtab);
    // far tail (r > 5, as241.f90:105): essentially never.  Tested on the high word of r (5.0 is
    // 0x40140000:00000000; at r = 5 itself the two formulas agree to the algorithm's 1e-16)
    if (__double2hiint(r) >= 0x40140000) z = normal_tail_far_f64(w, r);
    sts_f64(za, z);
  }
  // two tail draws side by side (two independent dependency chains); `two` false: w1 is a copy
  // of w0 and nothing is stored for it
  static __device__ __forceinline__ void tail_two(uint64_t w0, uint64_t w1, bool two, uint32_t za0,
                                                  uint32_t za1, uint32_t logtab) {
    double r0, r1;
    double z0 = normal_tail_mid_f64(w0, r0, logtab), z1 = normal_tail_mid_f64(w1, r1, logtab);
    if (max(__double2hiint(r0), __double2hiint(r1)) >= 0x40140000) {  // far tail, see tail_one
      if (__double2hiint(r0) >= 0x40140000) z0 = normal_tail_far_f64(w0, r0);
      if (__double2hiint(r1) >= 0x40140000) z1 = normal_tail_far_f64(w1, r1);
    }
    sts_f64(za0, z0);
    if (two) sts_f64(za1, z1);
  }
  static __device__ __forceinline__ void get(uint32_t addr, double& zv, double& zx) {
    lds_f64x2(addr, zv, zx);
  }
};

// Where word j of a lane's refill and its normal live (zcol / ucol: that lane's columns).
template <int NORMAL_MODE, int RING>
__device__ __forceinline__ uint32_t tail_z_addr(uint32_t zcol, uint32_t zstride, uint32_t j) {
  constexpr int kLog = RING == 16 ? 4 : RING == 8 ? 3 : RING == 4 ? 2 : 1;
  constexpr uint32_t kHalf = ZRing<NORMAL_MODE>::kBytesPerStep / 2;
  return zcol + (j & (RING - 1)) * zstride + (j >> kLog) * kHalf;
}
template <int NORMAL_MODE, int RING>
__device__ __forceinline__ uint32_t tail_word_addr(uint32_t ucol, uint32_t ustride, uint32_t zaddr,
                                                   uint32_t j) {
  if (ZRing<NORMAL_MODE>::kRawPlanes == 2) return ucol + j * ustride;
  return j < (uint32_t)RING ? ucol + j * ustride : zaddr;  // F64: a spot word waits in its z slot
}

// Tail phase, per lane: each lane loops over its own tail draws (F32 modes: two per iteration,
// independent evaluations hide the MUFU latencies).  The warp runs as long as its unluckiest lane.
template <int NORMAL_MODE, int RING>
__device__ __forceinline__ void tail_phase_planar(uint32_t tails, const RingAddr& ra) {
  using Ring = ZRing<NORMAL_MODE>;
  while (tails) {
    const uint32_t j0 = bfind32(tails);
    const uint32_t b0 = 1u << j0, rest = tails ^ b0;
    const uint32_t za0 = tail_z_addr<NORMAL_MODE, RING>(ra.zcol, ra.zstride, j0);
    const uint64_t w0 = lds_b64(tail_word_addr<NORMAL_MODE, RING>(ra.ucol, ra.ustride, za0, j0));
    if constexpr (Ring::kTailPairs) {
      const bool two = rest != 0;
      const uint32_t j1 = bfind32(two ? rest : b0);  // j0 again when it was the last one
      tails = rest & ~(1u << j1);
      const uint32_t za1 = tail_z_addr<NORMAL_MODE, RING>(ra.zcol, ra.zstride, j1);
      const uint64_t w1 = lds_b64(tail_word_addr<NORMAL_MODE, RING>(ra.ucol, ra.ustride, za1, j1));
      Ring::tail_pair(w0, w1, two, za0, za1);
    } else {
      tails = rest;
      Ring::tail_one(w0, za0, ra.logtab);
    }
  }
}

// Tail phase, warp-cooperative: the tail draws of all 32 lanes (154 +- 11 of the 1024 draws of a
// 16-step refill with AS241's split at |q| = 0.425; 102 +- 10 in the as-built F32 mode, which
// leaves the central formula at 0.45) are listed in shared memory and dealt out evenly.  F32
// modes: two draws per lane and iteration, ceil(total / 64) = 3 (2) iterations instead of as
// many as the unluckiest lane needs (4.8 on average).  F64 mode: ceil(total / 32) = 5 draws per
// lane, taken two at a time (96 % of the lane slots do work; dealt out in pairs it would be
// 80 %).  All 32 lanes of the warp must call this together.
//   list entry = (owner lane << 5) | word index j, 16 bits; lane l takes entries [l c, l c + c),
//   c = entries per lane, so that lanes working side by side read different owners' columns
template <int NORMAL_MODE, int RING>
__device__ __forceinline__ void tail_phase_coop(uint32_t tails, const RingAddr& ra, uint32_t lane) {
  using Ring = ZRing<NORMAL_MODE>;
  constexpr uint32_t kB = Ring::kBytesPerStep;
  const uint32_t cnt = __popc(tails);
  uint32_t incl = cnt;  // inclusive prefix sum of the lanes' counts
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl += up;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total > (uint32_t)kTailListCap) {  // warp-uniform; never in practice
    tail_phase_planar<NORMAL_MODE, RING>(tails, ra);
    return;
  }
  uint32_t la = ra.tlist + 2 * (incl - cnt);
  const uint32_t ubase = ra.ucol - 8 * lane, zbase = ra.zcol - kB * lane;  // lane 0's columns
  if constexpr (Ring::kTailPairs) {
    const uint32_t tag = lane << 5;
    while (tails) {
      const uint32_t j = bfind32(tails);
      tails ^= 1u << j;
      sts_u16(la, tag | j);
      la += 2;
    }
    __syncwarp();
    const uint32_t c = ((total + 63) >> 6) << 1;
    uint32_t i = lane * c;
    for (uint32_t t = 0; t < c; t += 2, i += 2) {
      if (i < total) {
        const uint32_t e2 = lds_u32(ra.tlist + 2 * i);
        const bool two = i + 1 < total;
        const uint32_t e0 = e2 & 0xffffu, e1 = two ? e2 >> 16 : e0;
        const uint32_t l0 = e0 >> 5, j0 = e0 & 31u, l1 = e1 >> 5, j1 = e1 & 31u;
        const uint32_t za0 = tail_z_addr<NORMAL_MODE, RING>(zbase + kB * l0, ra.zstride, j0);
        const uint32_t za1 = tail_z_addr<NORMAL_MODE, RING>(zbase + kB * l1, ra.zstride, j1);
        const uint64_t w0 =
            lds_b64(tail_word_addr<NORMAL_MODE, RING>(ubase + 8 * l0, ra.ustride, za0, j0));
        const uint64_t w1 =
            lds_b64(tail_word_addr<NORMAL_MODE, RING>(ubase + 8 * l1, ra.ustride, za1, j1));
        Ring::tail_pair(w0, w1, two, za0, za1);
      }
    }
  } else {
    // F64 mode.  Entry = (z slot of the draw - zblock) / 8 = 2 owner thread + 2 T (j mod RING) +
    // (j div RING): the z slot is zblock + 8 entry; an odd entry is a spot draw, whose raw word
    // waits in that very slot, an even one a variance draw with its word at ublock + 4 entry (the
    // variance plane has half the z ring's pitch).  Five instructions of decoding per draw.
    static_assert(kB == 16, "entry arithmetic of the F64 ring");
    // pin32: otherwise ptxas re-derives these from %ntid and the shared window base inside the
    // two loops (a constant load in the serial list loop)
    const uint32_t t2 = ra.zstride >> 3, tag = pin32((ra.zcol - ra.zblock) >> 3);
    const uint32_t spotfix = pin32(RING * t2 - 1);  // entry(j) = tag + j t2 - (j div RING) spotfix
    const uint32_t zblock = pin32(ra.zblock), ublock = pin32(ra.ublock);
    while (tails) {
      const uint32_t j = bfind32(tails);
      tails ^= 1u << j;
      uint32_t e = j * t2 + tag;
      if (j >= (uint32_t)RING) e -= spotfix;
      sts_u16(la, e);
      la += 2;
    }
    __syncwarp();
    // c entries per lane (5 as a rule): two at a time -- two independent dependency chains per
    // lane, the formula is one long chain of FP64 latencies -- and the odd one alone.  c is the
    // same for all lanes, so no lane slot is spent on a draw that does not exist (except in the
    // lane where the list ends).
    const uint32_t c = (total + 31) >> 5;
    uint32_t i = lane * c;
    auto entry = [&](uint32_t k, uint32_t& za) {
      const uint32_t e = lds_u16(ra.tlist + 2 * k);
      za = zblock + 8 * e;
      return lds_b64((e & 1u) ? za : ublock + 4 * e);
    };
    uint32_t t = 0;
    for (; t + 1 < c; t += 2, i += 2) {
      if (i < total) {
        const bool two = i + 1 < total;
        uint32_t za0, za1;
        const uint64_t w0 = entry(i, za0);
        const uint64_t w1 = entry(two ? i + 1 : i, za1);
        Ring::tail_two(w0, w1, two, za0, za1, ra.logtab);
      }
    }
    if (t < c && i < total) {
      uint32_t za;
      const uint64_t w = entry(i, za);
      Ring::tail_one(w, za, ra.logtab);
    }
  }
  __syncwarp();
}

// Refill of a thread's whole ring: RING / 8 generator rounds -- raw words parked in their planes,
// central normals in the z ring -- then ONE tail phase over all their draws.  `o` holds the
// first round already when have_first.  This is the code the path kernel runs per refill AND
// what hexo_gpu_normals_from_words exposes for the parity test of the transform.
template <int NORMAL_MODE, class Gen>
__device__ __forceinline__ void ring_refill(Gen& rng, uint64_t (&o)[16], bool have_first,
                                            const RingAddr& ra, uint32_t lane) {
  using Ring = ZRing<NORMAL_MODE>;
  constexpr int kRing = ring_steps(NORMAL_MODE);
  const uint32_t uplane = kRing * ra.ustride;
  uint32_t tv = 0, tx = 0;
#pragma unroll
  for (int r = 0; r < kRing / kStepsPerRound; ++r) {
    if (r > 0 || !have_first) rng.round(o);
    const uint32_t u0 = ra.ucol + r * kStepsPerRound * ra.ustride;
#pragma unroll
    for (int s = 0; s < kStepsPerRound; ++s) {
      sts_b64(u0 + s * ra.ustride, o[2 * s]);
      if (Ring::kRawPlanes == 2) sts_b64(u0 + s * ra.ustride + uplane, o[2 * s + 1]);
    }
    Ring::central_round_planar(o, r * kStepsPerRound, ra.zcol, ra.zstride, tv, tx);
  }
  // After kRing shifts the flag of step s sits at bit kRing - 1 - s of its accumulator; the tail
  // phase wants bit j = word index (variance word of step s: j = s, spot word: j = kRing + s).
  const uint32_t tails = __brev((tv << (32 - kRing)) | (tx << (32 - 2 * kRing)));
#if HEXO_TAIL_COOP
  tail_phase_coop<NORMAL_MODE, kRing>(tails, ra, lane);
#else
  tail_phase_planar<NORMAL_MODE, kRing>(tails, ra);
#endif
}

// Payoff accumulation of one maturity with the control variate c = final value - S: the walk
// over the warp's 32 final values also collects sum pf c per strike, and lane 0 adds sum c and
// sum c^2 of the maturity (in lane order).  Once per path and maturity, so it is kept out of line:
// the step loop's register allocation must not pay for it.
__device__ __forceinline__ void accumulate_with_control(const double* fvbuf, unsigned amask, int lane,
                                                     const double* strikes, uint32_t n_strikes,
                                                     double S, double* sum, double* sq,
                                                     double* cross, double* ctl, double* ctl2) {
  for (uint32_t j = lane; j < n_strikes; j += 32) {
    const double K = __ldg(strikes + j);
    double s = 0.0, q = 0.0, x = 0.0;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) {
      if ((amask >> l) & 1u) {
        const double f = fvbuf[l];
        const double pf = fmax(f - K, 0.0);
        s += pf;
        q = fma(pf, pf, q);
        x = fma(pf, f - S, x);
      }
    }
    sum[j] += s;
    sq[j] += q;
    cross[j] += x;
  }
  if (lane == 0) {
    double c1 = 0.0, c2 = 0.0;
    for (int l = 0; l < 32; ++l) {
      if ((amask >> l) & 1u) {
        const double c = fvbuf[l] - S;
        c1 += c;
        c2 = fma(c, c, c2);
      }
    }
    *ctl += c1;
    *ctl2 += c2;
  }
}

// The same walk with the geometric-Asian control: per strike the payoff of the geometric average
// c = max(G - K, 0) next to the payoff of the arithmetic one, and their sums
// [sum pf | sum pf^2 | sum pf c | sum c | sum c^2].
__device__ __forceinline__ void accumulate_with_geometric(const double* fvbuf, const double* gbuf,
                                                          unsigned amask, int lane,
                                                          const double* strikes, uint32_t n_strikes,
                                                          double* sum, double* sq, double* cross,
                                                          double* csum, double* csq) {
  for (uint32_t j = lane; j < n_strikes; j += 32) {
    const double K = __ldg(strikes + j);
    double s = 0.0, q = 0.0, x = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll 4
    for (int l = 0; l < 32; ++l) {
      if ((amask >> l) & 1u) {
        const double pf = fmax(fvbuf[l] - K, 0.0);
        const double cg = fmax(gbuf[l] - K, 0.0);
        s += pf;
        q = fma(pf, pf, q);
        x = fma(pf, cg, x);
        c1 += cg;
        c2 = fma(cg, cg, c2);
      }
    }
    sum[j] += s;
    sq[j] += q;
    cross[j] += x;
    csum[j] += c1;
    csq[j] += c2;
  }
}

// Where the per-maturity constants are read from: global memory (any number of maturities), the
// kernel parameter bank indexed by the maturity (up to kInlineSegs), or -- one maturity, the
// benchmark shape -- fixed parameter-bank addresses, which the compiler can keep in uniform
// registers: a DFMA takes one uniform-register operand for free, so every DFMA of the step loop
// then reads at most two register pairs.
enum : int { kSegsGlobal = 0, kSegsInline = 1, kSegsSingle = 2 };

// CV: hexo_control_variate -- 0 plain sums; 1 the control c = final value - S; 2 (Asian) the
// geometric-Asian control c_j = max(G - K_j, 0), G = exp of the same accumulation applied to ln X
template <int PAYOFF, int NORMAL_MODE, int SEGS, class Gen = Shishua, int CV = 0,
          bool MART = false>
__global__ void __launch_bounds__(kMaxBlock, kMinBlocksPerSM)
heston_qe_paths_kernel(const __grid_constant__ PathArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = blockDim.x, nwarps = T >> 5;
  using Ring = ZRing<NORMAL_MODE>;
  constexpr bool kAsian = PAYOFF == HEXO_PAYOFF_ASIAN;
  constexpr int kRing = ring_steps(NORMAL_MODE);  // steps per refill

  unsigned char* sp = smem_raw;
  // pin32: keep the ring addresses in registers; otherwise ptxas re-derives them from
  // %tid / %ntid / the shared window base inside the step and tail loops (~20 instructions)
  // raw words in planes, [variance words | spot words], each [kRing][T]: the two 64-bit
  // stores of a step are not adjacent (adjacent, ptxas fuses them into one 128-bit store and
  // pays four moves to line the words up in an aligned register quad) and a warp's 64-bit
  // accesses touch consecutive 8-byte slots
  const uint32_t zstride = pin32(Ring::kBytesPerStep * T), ustride = pin32(8 * T);
  const uint32_t ucol = pin32(smem_addr(sp) + 8 * tid);  // variance word of step 0
  const uint32_t ublock0 = smem_addr(sp);
  sp += (size_t)8 * Ring::kRawPlanes * kRing * T;
  const uint32_t zcol = pin32(smem_addr(sp) + Ring::kBytesPerStep * tid);
  const uint32_t zblock0 = smem_addr(sp);
  sp += (size_t)Ring::kBytesPerStep * kRing * T;
  const uint32_t logtab = ring_logtab_bytes(NORMAL_MODE) ? pin32(smem_addr(sp)) : 0u;
  ring_logtab_init<NORMAL_MODE>(smem_raw, tid, T);
  sp += ring_logtab_bytes(NORMAL_MODE);
  double* exptab = reinterpret_cast<double*>(sp);
  const uint32_t exptab_s = pin32(smem_addr(sp));
  sp += 32 * 8;
  double* fvbuf = reinterpret_cast<double*>(sp) + 32 * warp;
  sp += (size_t)32 * 8 * nwarps;
  const uint32_t tlist = pin32(smem_addr(sp) + kTailListBytes * warp);
  sp += (size_t)kTailListBytes * nwarps;
#define HEXO_N_ACC (CV == 2 ? 5 * a.n_opts : CV ? 3 * a.n_opts + 2 * a.n_seg : 2 * a.n_opts)
  // The plain case spells its offsets out as the 64-bit product chain it has always been: the
  // register allocation of the whole kernel (114 registers, 1 % faster step loop) hangs on it.
  double* acc_all = a.gacc ? a.gacc + (CV ? (size_t)blockIdx.x * nwarps * HEXO_N_ACC
                                         : (size_t)blockIdx.x * nwarps * 2 * a.n_opts)
                           : reinterpret_cast<double*>(sp);
  double* my_sum =  // lane-owned slots
      acc_all + (CV ? (size_t)warp * HEXO_N_ACC : (size_t)warp * 2 * a.n_opts);
  double* my_sq = my_sum + a.n_opts;
  double* my_cross = my_sq + a.n_opts;    // cv only: sum pf c per option ...
  double* my_ctl = my_cross + a.n_opts;   // ... and sum c, sum c^2 per maturity (lane 0); with
                                          // the geometric control per option: [n_opts | n_opts]
  constexpr bool kGeo = kAsian && CV == 2;
  // the warp's geometric averages at a maturity share the tail list's storage (idle then)
  double* gbuf = reinterpret_cast<double*>(smem_raw + (tlist - smem_addr(smem_raw)));
  if (!a.gacc)
    for (uint32_t j = lane; j < HEXO_N_ACC; j += 32) my_sum[j] = 0.0;
  exp_table_init(exptab, tid, T);

  const uint64_t slot = (uint64_t)blockIdx.x * T + tid;
  const uint64_t sid = a.stream_begin + slot;
  const uint64_t my_paths =
      slot < a.stream_count ? a.base_paths + (sid < a.rem_streams ? 1u : 0u) : 0u;
  // Path counts are non-increasing in the stream id, so the block's first stream holds the
  // block's maximum.  It depends on blockIdx and kernel parameters only: every loop below is
  // uniform across the block as far as the compiler can prove, which lets it keep loop counters,
  // strides and constants in uniform registers.  Lanes whose own stream has fewer paths run
  // along and are masked out where payoffs are accumulated.
  const uint64_t block_first = (uint64_t)blockIdx.x * T;
  const uint64_t block_paths =
      block_first < a.stream_count
          ? a.base_paths + (a.stream_begin + block_first < a.rem_streams ? 1u : 0u)
          : 0u;

  Gen rng;  // Shishua (the reference's generator) or PhiloxGen (optional counter mode)
  const RingAddr ra = {ucol, ustride, zcol, zstride, tlist, logtab, ublock0, zblock0};
  if (ring_logtab_bytes(NORMAL_MODE)) __syncthreads();  // logtab: the first refill reads it
  {
    uint64_t o[16];
    rng.init(a.seed, sid, 0, 0, o);
    ring_refill<NORMAL_MODE>(rng, o, true, ra, lane);
  }
  uint32_t pos = 0;  // next unread step of the round
  __syncthreads();   // exptab

  for (uint64_t p = 0; p < block_paths; ++p) {
    const bool active = p < my_paths;
    // HQEAnderson::operator=(initial_state), HSimulation.tpp:26,87-94
    double V = a.v0, lnX = a.lnS, X = a.S, Xprev = a.S;
    double integral = 0.0;  // AAsianCallNonAdaptive::accumulated_value, reset per path (:34)
    // geometric control: the same bookkeeping for L = ln X (lnX itself; Lprev, integralL)
    double Lprev = a.lnS, integralL = 0.0;
    for (uint32_t k = 0; k < (SEGS == kSegsSingle ? 1u : a.n_seg); ++k) {
      SegConst g = SEGS == kSegsSingle   ? a.seg_inline[0]
                   : SEGS == kSegsInline ? a.seg_inline[k]
                                         : a.segs[k];
      // keep the per-step constants in registers: otherwise ptxas re-loads each of them from
      // the constant bank (LDC) at every use inside the step loop
      if (SEGS != kSegsSingle) {
        g.D = pin(g.D); g.m0 = pin(g.m0); g.c1h = pin(g.c1h); g.c2h = pin(g.c2h);
        g.K0 = pin(g.K0); g.K1 = pin(g.K1); g.K2 = pin(g.K2); g.K3 = pin(g.K3);
      }
      {
        const uint32_t n = g.n_steps;
        if (kAsian && k > 0 && n > 0) {
          // The trapezoid of the step that crossed the previous expiry is added
          // AFTER update_earliest switched the step size (HSimulation.tpp:42-44),
          // i.e. with this segment's h.
          integral += g.hcarry * (X + Xprev);
          if (kGeo) integralL += g.hcarry * (lnX + Lprev);
        }
        const double Xa = X, La = lnX;
        double sumX = 0.0, sumL = 0.0;
        // log-spot half of a step.  Asian: the spot itself is advanced multiplicatively
        // (X is needed every step, ln X never); European: ln X is accumulated and X = exp(ln X)
        // only where with_x says so (HSimulation.tpp:80-82)
        // keep_prev: remember X before the step (only the last step of a run needs it)
        // k0 (MART): -ln M of that step (HEXO_DRIFT_MARTINGALE)
        auto spot_half = [&](double Vfrom, double Vto, double zx, double k0, auto with_x,
                             bool keep_prev) {
          const double delta = MART ? qe_logreturn_mart(g, Vfrom, Vto, zx, k0)
                                    : qe_logreturn(g, Vfrom, Vto, zx);
          if (kAsian) {
            if (keep_prev) Xprev = X;
            X = grow_spot(X, delta, exptab_s);
            sumX += X;
            if (kGeo) {
              if (keep_prev) Lprev = lnX;
              lnX += delta;
              sumL += lnX;
            }
          } else {
            lnX += delta;
            if (decltype(with_x)::value) {
              Xprev = X;
              X = fast_exp(lnX, exptab_s);
            }
          }
        };
        // `count` (> 0) steps of this segment, software-pipelined: iteration j does the
        // log-spot / exp half of step j-1 next to the variance half of step j.
        // with_x: also X = exp(ln X) (HSimulation.tpp:81-82)
        auto run = [&](uint32_t count, auto with_x) {
          double Vold = V, zx_pend = 0.0;  // (V_{j-1}, Z_X of step j-1) of the pending half
          double k0_pend = 0.0;            // MART: -ln M of step j-1
          bool first = true;
          while (count) {
            if (pos == kRing) {
#ifdef HEXO_DEV_PROBES
              if (a.dev_no_refill == 1) {
                // probe "step loop alone": no generator, no transform -- the ring is reused with
                // every normal's sign flipped (the draws stay N(0,1), the paths stay in their
                // usual regime; reusing them unchanged drives V into the psi >= 1.5 branch)
                for (int s = 0; s < kRing; ++s) {
                  const uint32_t za = zcol + s * zstride;
                  if (NORMAL_MODE == HEXO_NORMAL_F64) {
                    sts_b64(za, lds_b64(za) ^ 0x8000000000000000ull);
                    sts_b64(za + 8, lds_b64(za + 8) ^ 0x8000000000000000ull);
                  } else {
                    sts_b64(za, lds_b64(za) ^ 0x8000000080000000ull);
                  }
                }
              } else
#endif
              {
                uint64_t o[16];
                ring_refill<NORMAL_MODE>(rng, o, false, ra, lane);
              }
              pos = 0;
            }
            uint32_t m = min(kRing - pos, count);
            count -= m;
            uint32_t za = zcol + pos * zstride;
            // the variance word of the step whose normals sit at `z` (read on the psi >= 1.5
            // path only: derived there instead of carried through the loop)
            auto uniform_at = [&](uint32_t z) {
              return u64_to_unit(lds_b64(ucol + (z - zcol) / (Ring::kBytesPerStep / 8)));
            };
            pos += m;
            if (first) {  // prologue: variance half of the first step
              first = false;
              double zv;
              Ring::get(za, zv, zx_pend);
              Vold = V;
              V = qe_variance<MART>(g, Vold, zv, [&]() { return uniform_at(za); }, &k0_pend);
              --m, za += zstride;
            }
            // unrolled by two so that the loop-carried rotation (Vold <- V <- V', Z_X) becomes
            // register renaming instead of moves
#pragma unroll kStepUnroll
            for (; m; --m, za += zstride) {
              double zv, zx;
              Ring::get(za, zv, zx);
#ifdef HEXO_DEV_PROBES
              if (a.dev_no_refill == 2) {  // probe "generator + transform alone": trivial consumer
                sumX += zv + zx;
                continue;
              }
#endif
              // Second half of the previous step and first half of this one, straight-line
              // parts first: two independent dependency chains in ONE basic block.  The rare
              // cases of both (|log-return| > 0.08; psi >= 1.5) share a single branch behind
              // them -- a branch per half would put a convergence barrier between the chains
              // and ptxas then runs them one after the other.
              const double delta = MART ? qe_logreturn_mart(g, Vold, V, zx_pend, k0_pend)
                                        : qe_logreturn(g, Vold, V, zx_pend);
              QeVarMid mid;
              double Vn = qe_variance_quad<MART>(g, V, zv, mid);
              if (kAsian) {
                double Xn = grow_spot_poly(X, delta);
                const bool rare_x = grow_spot_is_rare(delta);
                if (rare_x | mid.rare) {
                  if (rare_x) Xn = grow_spot_rare(X, delta, exptab_s);
                  if (mid.rare) {
                    if (qe_rare_exact(mid))
                      Vn = qe_variance_rare<MART>(g, V, mid, [&]() { return uniform_at(za); });
                  }
                }
                X = Xn;
                sumX += X;
                if (kGeo) {
                  lnX += delta;
                  sumL += lnX;
                }
              } else {
                lnX += delta;
                if (decltype(with_x)::value) {
                  Xprev = X;
                  X = fast_exp(lnX, exptab_s);
                }
                if (mid.rare) {
                  if (qe_rare_exact(mid))
                    Vn = qe_variance_rare<MART>(g, V, mid, [&]() { return uniform_at(za); });
                }
              }
              Vold = V;
              V = Vn;
              zx_pend = zx;
              if (MART) k0_pend = mid.k0;
            }
          }
          // epilogue: second half of the last step
          spot_half(Vold, V, zx_pend, k0_pend, with_x, true);
        };
        if (kAsian) {
          if (n > 0) run(n, std::true_type{});
          // trapezoids of all but the crossing step: h/2 sum_{j<n} (X_j + X_{j-1}),
          // AsianContract.h:25-28; sumX includes the crossing step's X, take it out
          if (n > 0) integral += g.h * 0.5 * (Xa - Xprev + 2.0 * (sumX - X));
          if (kGeo && n > 0) integralL += g.h * 0.5 * (La - Lprev + 2.0 * (sumL - lnX));
        } else {
          // European: X is only read at the expiry, so only the last two steps need it
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
      if (kGeo)  // geometric average on the arithmetic average's own weights, normalised
        gbuf[lane] = fast_exp(
            (integralL + (lnX - Lprev) * g.w + g.hs * (lnX + Lprev)) * g.inv_logw, exptab_s);
      const unsigned amask = __ballot_sync(0xffffffffu, active);
      __syncwarp();
      // final_payoff for every strike of this chain (HSimulation.tpp:39-40): lane
      // l owns strikes l, l+32, ... and walks the warp's 32 final values.
      if (!CV) {
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
      } else if (CV == 2) {
        accumulate_with_geometric(fvbuf, gbuf, amask, lane, a.strikes + g.first_opt, g.n_strikes,
                                  my_sum + g.first_opt, my_sq + g.first_opt,
                                  my_cross + g.first_opt, my_ctl + g.first_opt,
                                  my_ctl + a.n_opts + g.first_opt);
        __syncwarp();  // gbuf is the tail list's storage: the next refill writes it again
      } else {
        accumulate_with_control(fvbuf, amask, lane, a.strikes + g.first_opt, g.n_strikes, a.S,
                                my_sum + g.first_opt, my_sq + g.first_opt, my_cross + g.first_opt,
                                my_ctl + k, my_ctl + a.n_seg + k);
      }
    }
  }

  // warps -> block partial, fixed order
  __syncthreads();
  const uint32_t n2 = HEXO_N_ACC;
  for (uint32_t j = tid; j < n2; j += T) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += acc_all[(size_t)w * n2 + j];
    a.partials[(size_t)blockIdx.x * n2 + j] = s;
  }
#undef HEXO_N_ACC
}

}  // namespace hexo
