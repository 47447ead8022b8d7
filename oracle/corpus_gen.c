/*
 * oracle/corpus_gen.c — TEST INFRASTRUCTURE. Seeded synthetic corpus, SURVEY.md §8(d):
 * the reference ships no corpus (its inputs are absolute laptop paths,
 * cmu_ron/TrainAndPredict.java:203-205) so every config is generated. A document is a bag of
 * word ids exactly like the FeatureSequence the reference's importer produces
 * (cmu_ron/InstanceImporter.java:58-68: doc = test id, token = source-file path).
 *
 * Generative model: k_true topics; phi_k ∝ Dirichlet(0.01) sample × Zipf(1.07) base measure;
 * theta_d ~ Dirichlet(0.1); L_d = max(1, round(LogNormal(mu, 0.6))) with E[L_d] = mean_len;
 * token: topic ~ theta_d, word ~ phi_topic (alias tables).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lda_oracle.h"

typedef struct { uint64_t s[4]; } xo256;

static uint64_t splitmix64(uint64_t* x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static void xo_seed(xo256* r, uint64_t seed) {
  for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&seed);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t xo_next(xo256* r) {
  uint64_t* s = r->s;
  uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
  return result;
}
static double xo_unif(xo256* r) { return (double)(xo_next(r) >> 11) * 0x1p-53; }
static double xo_unif_open(xo256* r) { return ((double)(xo_next(r) >> 11) + 0.5) * 0x1p-53; }
static double xo_normal(xo256* r) {
  double u1 = xo_unif_open(r), u2 = xo_unif(r);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
/* Marsaglia-Tsang; for a < 1 boost through Gamma(a+1) * U^(1/a). */
static double xo_gamma(xo256* r, double a) {
  if (a < 1.0) {
    double g = xo_gamma(r, a + 1.0);
    return g * pow(xo_unif_open(r), 1.0 / a);
  }
  double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double x = xo_normal(r), v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = xo_unif_open(r);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
  }
}

typedef struct { float* prob; int32_t* alias; int32_t n; } alias_table;

static void alias_build(alias_table* t, const double* p, int32_t n) {
  t->n = n;
  t->prob = (float*)malloc(sizeof(float) * (size_t)n);
  t->alias = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  double* sc = (double*)malloc(sizeof(double) * (size_t)n);
  int32_t* small = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  int32_t* large = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  int32_t ns = 0, nl = 0;
  for (int32_t i = 0; i < n; ++i) {
    sc[i] = p[i] * n;
    if (sc[i] < 1.0) small[ns++] = i; else large[nl++] = i;
  }
  while (ns > 0 && nl > 0) {
    int32_t s = small[--ns], l = large[--nl];
    t->prob[s] = (float)sc[s];
    t->alias[s] = l;
    sc[l] = (sc[l] + sc[s]) - 1.0;
    if (sc[l] < 1.0) small[ns++] = l; else large[nl++] = l;
  }
  while (nl > 0) { int32_t l = large[--nl]; t->prob[l] = 1.0f; t->alias[l] = l; }
  while (ns > 0) { int32_t s = small[--ns]; t->prob[s] = 1.0f; t->alias[s] = s; }
  free(sc); free(small); free(large);
}
static int32_t alias_draw(const alias_table* t, xo256* r) {
  uint64_t x = xo_next(r);
  int32_t i = (int32_t)(((x >> 32) * (uint64_t)t->n) >> 32);
  float u = (float)(x & 0xFFFFFF) * 0x1p-24f;
  return u < t->prob[i] ? i : t->alias[i];
}
static void alias_free(alias_table* t) { free(t->prob); free(t->alias); }

int64_t oracle_gen_corpus(int64_t D, int32_t V, double mean_len, int32_t k_true, uint64_t seed,
                          int64_t* doc_ptr, int32_t* tok_word, int64_t tok_cap) {
  xo256 rl;
  xo_seed(&rl, seed ^ 0xA5A5A5A5DEADBEEFULL);
  const double sigma = 0.6, mu = log(mean_len) - 0.5 * sigma * sigma;
  /* pass 1: lengths (own stream, so N does not depend on whether tokens are generated) */
  doc_ptr[0] = 0;
  for (int64_t d = 0; d < D; ++d) {
    double len = floor(exp(mu + sigma * xo_normal(&rl)) + 0.5);
    if (len < 1.0) len = 1.0;
    if (len > 65535.0) len = 65535.0;
    doc_ptr[d + 1] = doc_ptr[d] + (int64_t)len;
  }
  const int64_t N = doc_ptr[D];
  if (!tok_word) return N;
  if (tok_cap < N) return -N;

  xo256 r;
  xo_seed(&r, seed);
  double* zipf = (double*)malloc(sizeof(double) * (size_t)V);
  for (int32_t w = 0; w < V; ++w) zipf[w] = pow((double)(w + 1), -1.07);
  alias_table* phi = (alias_table*)malloc(sizeof(alias_table) * (size_t)k_true);
  double* p = (double*)malloc(sizeof(double) * (size_t)V);
  for (int32_t k = 0; k < k_true; ++k) {
    double tot = 0.0;
    for (int32_t w = 0; w < V; ++w) {
      p[w] = xo_gamma(&r, 0.01) * zipf[w];
      tot += p[w];
    }
    if (!(tot > 0.0)) {
      tot = 0.0;
      for (int32_t w = 0; w < V; ++w) { p[w] = zipf[w]; tot += p[w]; }
    }
    for (int32_t w = 0; w < V; ++w) p[w] /= tot;
    alias_build(&phi[k], p, V);
  }
  double* theta = (double*)malloc(sizeof(double) * (size_t)k_true);
  for (int64_t d = 0; d < D; ++d) {
    double tot = 0.0;
    for (int32_t k = 0; k < k_true; ++k) { theta[k] = xo_gamma(&r, 0.1); tot += theta[k]; }
    if (!(tot > 0.0)) { theta[0] = 1.0; tot = 1.0; }
    alias_table th;
    for (int32_t k = 0; k < k_true; ++k) theta[k] /= tot;
    alias_build(&th, theta, k_true);
    for (int64_t i = doc_ptr[d]; i < doc_ptr[d + 1]; ++i) {
      int32_t k = alias_draw(&th, &r);
      tok_word[i] = alias_draw(&phi[k], &r);
    }
    alias_free(&th);
  }
  for (int32_t k = 0; k < k_true; ++k) alias_free(&phi[k]);
  free(phi); free(p); free(theta); free(zipf);
  return N;
}
