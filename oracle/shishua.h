/*
 * oracle/shishua.h -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Portable scalar restatement of the SHISHUA pseudo-random generator
 * (github.com/espadrine/shishua, the 128-byte-per-round "shishua" variant).
 *
 * Why this file exists: the reference does not vendor shishua -- it is cloned,
 * unpinned, from GitHub at `make deps` time (reference Makefile.am:87-89) and
 * patched only to add `inline` (reference shishua_inline_patch.diff:1-22).  The
 * source is absent from /root/reference and there is no network, so the
 * published algorithm is restated here.  The API is exactly what the reference's
 * call sites bind: `prng_state` (src/inc/RNG.h:25), `prng_init(prng_state*,
 * uint64_t seed[4])` (src/RNG.cpp:24, signature corroborated by
 * shishua_inline_patch.diff:7), `prng_gen(prng_state*, uint8_t*, size_t)`
 * (src/RNG.cpp:29).  The last row of the phi table, `memset`, `STEPS 1` and
 * `ROUNDS 13` are corroborated by shishua_inline_patch.diff:4-11.
 *
 * PARITY NOTE: byte-level equivalence with upstream shishua is UNPINNED (no
 * upstream golden vector exists inside the reference); the CUDA generator is
 * proven bit-exact against THIS restatement.
 *
 * State layout: four 256-bit rows s0..s3, each four 64-bit lanes:
 *   state[0..3]=s0, state[4..7]=s1, state[8..11]=s2, state[12..15]=s3.
 * One round emits 128 bytes (the previous round's output rows o0..o3, little
 * endian) and then advances the state.
 */
#ifndef HEXO_ORACLE_SHISHUA_H
#define HEXO_ORACLE_SHISHUA_H

#include <assert.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef struct prng_state {
  uint64_t state[16];
  uint64_t output[16];
  uint64_t counter[4];
} prng_state;

/* hex digits of the golden ratio */
static const uint64_t shishua_phi[16] = {
    0x9E3779B97F4A7C15ull, 0xF39CC0605CEDC834ull, 0x1082276BF3A27251ull, 0xF86C6A11D0C18E95ull,
    0x2767F0B153D27B7Full, 0x0347045B5BF1827Full, 0x01886F0928403002ull, 0xC1D64BA40F335E36ull,
    0xF06AD7AE9717877Eull, 0x85839D6EFFBD7DC6ull, 0x64D325D1C5371682ull, 0xCADD0CCCFDFFBBE1ull,
    0x626E33B8D04B4331ull, 0xBBF73C790D94F79Dull, 0x471C4AB3ED3D82A5ull, 0xFEC507705E4AE6E5ull,
};

/* 32-bit word `w` (0..7) of a 4x64-bit row */
static inline uint32_t shishua_word(const uint64_t row[4], unsigned w) {
  return (uint32_t)(row[w >> 1] >> ((w & 1u) * 32u));
}

/* Rotate a 256-bit row by `k` 32-bit words: out.word[i] = in.word[(i+k) mod 8]. */
static inline void shishua_rot32w(uint64_t out[4], const uint64_t in[4], unsigned k) {
  for (unsigned lane = 0; lane < 4; ++lane) {
    uint64_t lo = shishua_word(in, (2 * lane + k) & 7u);
    uint64_t hi = shishua_word(in, (2 * lane + 1 + k) & 7u);
    out[lane] = lo | (hi << 32);
  }
}

/* Advance the state by one round and latch the next 128 output bytes. */
static inline void shishua_round(prng_state *s) {
  uint64_t *s0 = &s->state[0], *s1 = &s->state[4], *s2 = &s->state[8], *s3 = &s->state[12];
  uint64_t *o0 = &s->output[0], *o1 = &s->output[4], *o2 = &s->output[8], *o3 = &s->output[12];
  uint64_t t0[4], t1[4], t2[4], t3[4], u0[4], u1[4], u2[4], u3[4];
  for (int i = 0; i < 4; ++i) {
    s1[i] += s->counter[i];
    s3[i] += s->counter[i];
    s->counter[i] += (uint64_t)(7 - 2 * i); /* 7,5,3,1 */
  }
  shishua_rot32w(t0, s0, 5);
  shishua_rot32w(t1, s1, 3);
  shishua_rot32w(t2, s2, 5);
  shishua_rot32w(t3, s3, 3);
  for (int i = 0; i < 4; ++i) {
    u0[i] = s0[i] >> 1;
    u1[i] = s1[i] >> 3;
    u2[i] = s2[i] >> 1;
    u3[i] = s3[i] >> 3;
  }
  for (int i = 0; i < 4; ++i) {
    s0[i] = t0[i] + u0[i];
    s1[i] = t1[i] + u1[i];
    s2[i] = t2[i] + u2[i];
    s3[i] = t3[i] + u3[i];
    o0[i] = u0[i] ^ t1[i];
    o1[i] = u2[i] ^ t3[i];
  }
  for (int i = 0; i < 4; ++i) {
    o2[i] = s0[i] ^ s3[i];
    o3[i] = s2[i] ^ s1[i];
  }
}

/* `size` must be a multiple of 128.  `buf` may be NULL (advance only). */
static inline void prng_gen(prng_state *s, uint8_t *buf, size_t size) {
  assert((size & 127u) == 0);
  for (size_t off = 0; off < size; off += 128) {
    if (buf != NULL) {
      for (int j = 0; j < 16; ++j) {
        uint64_t v = s->output[j];
        for (int b = 0; b < 8; ++b) buf[off + 8 * j + b] = (uint8_t)(v >> (8 * b));
      }
    }
    shishua_round(s);
  }
}

static inline void prng_init(prng_state *s, uint64_t seed[4]) {
  memset(s, 0, sizeof(*s));
  memcpy(s->state, shishua_phi, sizeof(shishua_phi));
  /* seed goes into even lanes only, so half of the state is never user-controlled */
  s->state[0] ^= seed[0];
  s->state[2] ^= seed[1];
  s->state[4] ^= seed[2];
  s->state[6] ^= seed[3];
  s->state[8] ^= seed[2];
  s->state[10] ^= seed[3];
  s->state[12] ^= seed[0];
  s->state[14] ^= seed[1];
  for (int r = 0; r < 13; ++r) {
    prng_gen(s, NULL, 128);
    for (int j = 0; j < 4; ++j) {
      s->state[j + 0] = s->output[j + 12];
      s->state[j + 4] = s->output[j + 8];
      s->state[j + 8] = s->output[j + 4];
      s->state[j + 12] = s->output[j + 0];
    }
  }
}

#endif /* HEXO_ORACLE_SHISHUA_H */
