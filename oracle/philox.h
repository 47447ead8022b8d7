/*
 * oracle/philox.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
 *
 * Philox4x32-10 counter-based RNG, restated from the published algorithm
 * (Salmon, Moraes, Dror, Shaw: "Parallel Random Numbers: As Easy as 1, 2, 3", SC'11).
 * It replaces cc.mallet.util.Randoms on the GPU path (reference call
 * sites `new Randoms()` cmu_ron/TrainAndPredict.java:34, cmu/TrainAndPredict.java:50).
 * Pinned by the Random123 known-answer vectors in tests/test_oracle_rng.py.
 */
#ifndef B200LDA_ORACLE_PHILOX_H
#define B200LDA_ORACLE_PHILOX_H
#include <stdint.h>

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
    uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHILOX_W0; k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/*
 * The sampler's per-token draw (DESIGN.md "sampling spec"): key = (seed_lo, seed_hi),
 * counter = (token_lo, token_hi, sweep, stream). Word 0 gives the 24-bit uniform.
 */
static inline void b200lda_token_random(uint64_t seed, uint64_t global_token, uint32_t sweep,
                                        uint32_t stream, uint32_t out[4]) {
  uint32_t ctr[4] = {(uint32_t)global_token, (uint32_t)(global_token >> 32), sweep, stream};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  philox4x32_10(ctr, key, out);
}

static inline float b200lda_u24(uint32_t x) { return (float)(x >> 8) * 0x1p-24f; }

#endif
