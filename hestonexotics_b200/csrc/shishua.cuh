// hestonexotics_b200/csrc/shishua.cuh
//
// SHISHUA generator for one CUDA thread: the whole 1024-bit state lives in
// registers and every 32-bit-word rotation of the published algorithm is a
// compile-time register renaming, so a round is pure 64-bit add / shift / xor
// on the integer pipe.  Replaces prng_init / prng_gen as the reference calls
// them (src/RNG.cpp:24,29; shishua itself is an un-vendored dependency fetched
// at reference Makefile.am:87-89).  Output is bit-identical, per stream, to the
// CPU restatement the tests compare against.
//
// Row r, lane i of the state is s[4*r+i].  `rounds` counts completed rounds;
// the generator's counter lanes are rounds*{7,5,3,1}.
#pragma once
#include <stdint.h>

namespace hexo {

struct Shishua {
  uint64_t s[16];
  uint32_t rounds;

  // word w (0..7) of row `row`, as a compile-time selection
  template <int W>
  __device__ __forceinline__ uint32_t word(const uint64_t* row) const {
    return (W & 1) ? (uint32_t)(row[W >> 1] >> 32) : (uint32_t)row[W >> 1];
  }
  // lane L of the row rotated by K 32-bit words
  template <int K, int L>
  __device__ __forceinline__ uint64_t rot(const uint64_t* row) const {
    return (uint64_t)word<(2 * L + K) & 7>(row) | ((uint64_t)word<(2 * L + 1 + K) & 7>(row) << 32);
  }

  // Advance one round; o[16] receives the 128 bytes this round latches (they
  // are what the published generator hands out on its NEXT call).
  __device__ __forceinline__ void round(uint64_t (&o)[16]) {
    const uint64_t n = rounds;
    const uint64_t c[4] = {n * 7ull, n * 5ull, n * 3ull, n};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s[4 + i] += c[i];
      s[12 + i] += c[i];
    }
    ++rounds;
    uint64_t t[16], u[16];
    t[0] = rot<5, 0>(s + 0);  t[1] = rot<5, 1>(s + 0);  t[2] = rot<5, 2>(s + 0);  t[3] = rot<5, 3>(s + 0);
    t[4] = rot<3, 0>(s + 4);  t[5] = rot<3, 1>(s + 4);  t[6] = rot<3, 2>(s + 4);  t[7] = rot<3, 3>(s + 4);
    t[8] = rot<5, 0>(s + 8);  t[9] = rot<5, 1>(s + 8);  t[10] = rot<5, 2>(s + 8); t[11] = rot<5, 3>(s + 8);
    t[12] = rot<3, 0>(s + 12); t[13] = rot<3, 1>(s + 12); t[14] = rot<3, 2>(s + 12); t[15] = rot<3, 3>(s + 12);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      u[i] = s[i] >> 1;
      u[4 + i] = s[4 + i] >> 3;
      u[8 + i] = s[8 + i] >> 1;
      u[12 + i] = s[12 + i] >> 3;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = t[i] + u[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i] = u[i] ^ t[4 + i];          // o0 = u0 ^ t1
      o[4 + i] = u[8 + i] ^ t[12 + i]; // o1 = u2 ^ t3
      o[8 + i] = s[i] ^ s[12 + i];     // o2 = s0 ^ s3
      o[12 + i] = s[8 + i] ^ s[4 + i]; // o3 = s2 ^ s1
    }
  }

  // prng_init: phi digits xor seed, 13 self-feeding rounds.  On return o[16]
  // holds the FIRST 128 bytes of the stream.
  __device__ __forceinline__ void init(uint64_t seed0, uint64_t seed1, uint64_t seed2,
                                       uint64_t seed3, uint64_t (&o)[16]) {
    s[0] = 0x9E3779B97F4A7C15ull ^ seed0;  s[1] = 0xF39CC0605CEDC834ull;
    s[2] = 0x1082276BF3A27251ull ^ seed1;  s[3] = 0xF86C6A11D0C18E95ull;
    s[4] = 0x2767F0B153D27B7Full ^ seed2;  s[5] = 0x0347045B5BF1827Full;
    s[6] = 0x01886F0928403002ull ^ seed3;  s[7] = 0xC1D64BA40F335E36ull;
    s[8] = 0xF06AD7AE9717877Eull ^ seed2;  s[9] = 0x85839D6EFFBD7DC6ull;
    s[10] = 0x64D325D1C5371682ull ^ seed3; s[11] = 0xCADD0CCCFDFFBBE1ull;
    s[12] = 0x626E33B8D04B4331ull ^ seed0; s[13] = 0xBBF73C790D94F79Dull;
    s[14] = 0x471C4AB3ED3D82A5ull ^ seed1; s[15] = 0xFEC507705E4AE6E5ull;
    rounds = 0;
#pragma unroll 1
    for (int r = 0; r < 13; ++r) {
      round(o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] = o[12 + j];
        s[4 + j] = o[8 + j];
        s[8 + j] = o[4 + j];
        s[12 + j] = o[j];
      }
    }
  }
};

}  // namespace hexo
