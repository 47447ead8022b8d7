/*
 * oracle/java_random.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
 *
 * java.util.Random (48-bit LCG) + cc.mallet.util.Randoms.nextUniform, restated from the
 * published JDK algorithm so the Mallet-faithful oracle consumes randomness the way
 * Mallet 2.0.7 would after setRandomSeed(s) (reference call sites
 * `new Randoms()` cmu_ron/TrainAndPredict.java:34, cmu/TrainAndPredict.java:50,353).
 * Pinned by JDK known answers (seed 42) in tests/test_oracle_rng.py.
 */
#ifndef B200LDA_ORACLE_JAVA_RANDOM_H
#define B200LDA_ORACLE_JAVA_RANDOM_H
#include <stdint.h>

typedef struct { uint64_t state; } java_random;

static inline void jr_seed(java_random* r, int64_t seed) {
  r->state = ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
}
static inline int32_t jr_next(java_random* r, int bits) {
  r->state = (r->state * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
  return (int32_t)((int64_t)r->state >> (48 - bits));
}
static inline int32_t jr_next_int(java_random* r) { return jr_next(r, 32); }
static inline int32_t jr_next_int_bound(java_random* r, int32_t n) {
  if ((n & -n) == n) return (int32_t)(((int64_t)n * (int64_t)jr_next(r, 31)) >> 31);
  int32_t bits, val;
  do {
    bits = jr_next(r, 31);
    val = bits % n;
  } while ((int32_t)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);
  return val;
}
/* Random.nextDouble() and Randoms.nextUniform() are the same 53-bit construction. */
static inline double jr_next_uniform(java_random* r) {
  int64_t hi = (int64_t)jr_next(r, 26);
  int64_t lo = (int64_t)jr_next(r, 27);
  return (double)((hi << 27) + lo) * 0x1p-53;
}
#endif
