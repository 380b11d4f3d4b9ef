// hestonexotics_b200/csrc/philox.cuh
//
// Optional counter-based generator for the path kernel (BASELINE north-star: "optionally a Philox
// mode"): Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random
// numbers: as easy as 1, 2, 3", SC'11).  Not part of the reference, which only has shishua; it
// is pinned by the Random123 known-answer vectors (tests/test_philox.py).
//
// Stream convention: stream s with seed k draws, for its n-th stepper call,
//   (c0,c1,c2,c3) = Philox4x32-10(counter = {n_lo, n_hi, s_lo, s_hi}, key = {k_lo, k_hi}),
//   variance word = c0 | c1 << 32,  spot word = c2 | c3 << 32.
// Any step of any stream is addressable without generating the ones before it.
#pragma once
#include <stdint.h>

namespace hexo {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += W0;
    k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Same interface as Shishua (shishua.cuh): init() and round() hand out 16 words = 8 steps.
struct PhiloxGen {
  uint64_t n;  // next stepper call of this stream
  uint32_t k0, k1, s0, s1;

  __device__ __forceinline__ void round(uint64_t (&o)[16]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t c[4];
      const uint64_t m = n + i;
      philox4x32_10((uint32_t)m, (uint32_t)(m >> 32), s0, s1, k0, k1, c);
      o[2 * i] = (uint64_t)c[0] | ((uint64_t)c[1] << 32);
      o[2 * i + 1] = (uint64_t)c[2] | ((uint64_t)c[3] << 32);
    }
    n += 8;
  }
  __device__ __forceinline__ void init(uint64_t seed, uint64_t stream, uint64_t, uint64_t,
                                       uint64_t (&o)[16]) {
    n = 0;
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
    s0 = (uint32_t)stream; s1 = (uint32_t)(stream >> 32);
    round(o);
  }
};

}  // namespace hexo
