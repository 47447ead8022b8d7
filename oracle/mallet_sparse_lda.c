#define _POSIX_C_SOURCE 200809L
/*
 * oracle/mallet_sparse_lda.c — TEST INFRASTRUCTURE (CPU oracle, part a). PARITY UNPINNED (see
 * lda_oracle.h): restated from Mallet 2.0.7's published algorithm (SURVEY.md Appendix A,
 * written from memory because cc.mallet:mallet:2.0.7 — reference pom.xml:107-111 — is not
 * vendored and no JVM exists here). Never linked into or called from the product library.
 *
 * What it stands in for, by reference call site:
 *   mallet_create            new ParallelTopicModel(K, alphaSum, beta)  cmu_ron/TrainAndPredict.java:160, cmu/TrainAndPredict.java:259
 *   mallet_add_instances     addInstances(InstanceList)                cmu_ron/…:162,174   cmu/…:260,271
 *   mallet_set_num_threads   setNumThreads(4)                          cmu_ron/…:164       cmu/…:262
 *   mallet_estimate          estimate()  (optimizeInterval pinned 0)   cmu_ron/…:166,175   cmu/…:265,272
 *   mallet_model_log_likelihood  modelLogLikelihood()                  cmu_ron/…:234       cmu/…:436
 *   mallet_get_topic_probabilities getTopicProbabilities(LabelSequence) cmu_ron/…:143      cmu/…:113
 *   mallet_infer             getInferencer().getSampledDistribution(inst,100,10,10)  cmu_ron/…:144  cmu/…:114
 *
 * Semantics kept: packed (count<<topicBits|topic) type-topic rows sorted descending; SparseLDA
 * s/r/q masses in double with bucket test order q -> r -> s; java.util.Random stream;
 * T worker replicas over contiguous D/T document ranges (last takes the remainder), each
 * rebuilding its replica from its own documents after the sweep, then sum + copy-back.
 * This is also the CPU baseline bench.py times (cpu_baseline.kind = "port").
 *
 * What does pin it, short of the jar (tests/test_oracle.py): java.util.Random known answers; the
 * first draw of a sweep against the textbook conditional over 72 000 seeds; the chain's long-run
 * state frequencies against the enumerated collapsed posterior on a 243-state corpus; the
 * inferencer against the closed form of a one-token document; the LL against scipy gammaln.
 * Those fix the mathematics Mallet implements, not Mallet's byte-for-byte output.
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "java_random.h"
#include "lda_oracle.h"

struct mallet_model {
  int32_t K, V;
  double alpha_sum, beta, beta_sum;
  double* alpha;
  int32_t topic_mask, topic_bits;
  int64_t D, N;
  int64_t* doc_ptr;
  int32_t* tok;
  int32_t* z;
  int32_t** ttc;     /* typeTopicCounts[V][len] */
  int32_t* ttc_len;  /* min(K, corpus frequency) */
  int32_t* tpt;      /* tokensPerTopic[K] */
  int32_t num_threads;
  int32_t random_seed; /* -1 = clock */
  java_random random;
  /* wall-clock split of estimate(): building the worker replicas, the (parallel) sampling phase,
   * the serial sumTypeTopicCounts + copy-back. Reported by bench.py's CPU baseline. */
  double t_setup, t_sample, t_merge;
};

static double now_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
  mallet_model* m;
  int64_t start_doc, num_docs;
  int32_t** ttc;
  int32_t* tpt;
  int owns_counts;
  java_random random;
  java_random* rng; /* points at own `random` or at the model's (T == 1) */
  double smoothing_only_mass;
  double* cached_coefficients;
  int32_t *local_counts, *local_index;
  double* term_scores;
} worker;

void oracle_java_random_ints(int64_t seed, int32_t n, int32_t bound, int32_t* out) {
  java_random r;
  jr_seed(&r, seed);
  for (int32_t i = 0; i < n; ++i) out[i] = bound > 0 ? jr_next_int_bound(&r, bound) : jr_next_int(&r);
}
void oracle_java_random_uniforms(int64_t seed, int32_t n, double* out) {
  java_random r;
  jr_seed(&r, seed);
  for (int32_t i = 0; i < n; ++i) out[i] = jr_next_uniform(&r);
}

static int32_t bit_count(uint32_t x) { return __builtin_popcount(x); }

mallet_model* mallet_create(int32_t K, double alpha_sum, double beta) {
  mallet_model* m = (mallet_model*)calloc(1, sizeof(*m));
  m->K = K;
  m->alpha_sum = alpha_sum;
  m->beta = beta;
  m->alpha = (double*)malloc(sizeof(double) * (size_t)K);
  for (int32_t k = 0; k < K; ++k) m->alpha[k] = alpha_sum / K;
  if (bit_count((uint32_t)K) == 1) {
    m->topic_mask = K - 1;
  } else {
    int32_t hi = 1;
    while (hi * 2 <= K) hi *= 2; /* Integer.highestOneBit */
    m->topic_mask = hi * 2 - 1;
  }
  m->topic_bits = bit_count((uint32_t)m->topic_mask);
  m->tpt = (int32_t*)calloc((size_t)K, sizeof(int32_t));
  m->num_threads = 1;
  m->random_seed = -1;
  jr_seed(&m->random, (int64_t)time(NULL));
  return m;
}

void mallet_destroy(mallet_model* m) {
  if (!m) return;
  for (int32_t w = 0; w < m->V; ++w) free(m->ttc[w]);
  free(m->ttc);
  free(m->ttc_len);
  free(m->tpt);
  free(m->alpha);
  free(m->doc_ptr);
  free(m->tok);
  free(m->z);
  free(m);
}

void mallet_set_random_seed(mallet_model* m, int32_t seed) {
  m->random_seed = seed;
  jr_seed(&m->random, (int64_t)seed);
}
void mallet_set_num_threads(mallet_model* m, int32_t t) { m->num_threads = t < 1 ? 1 : t; }
int64_t mallet_num_tokens(const mallet_model* m) { return m->N; }

/* insert one (type, topic) occurrence into a packed row: ++count, bubble towards the front */
static int row_increment(int32_t* row, int32_t len, int32_t topic, int32_t mask, int32_t bits) {
  int32_t index = 0;
  while (index < len && row[index] > 0 && (row[index] & mask) != topic) ++index;
  if (index == len) return -1;
  if (row[index] == 0) {
    row[index] = (1 << bits) + topic;
  } else {
    int32_t value = row[index] >> bits;
    row[index] = ((value + 1) << bits) + topic;
    while (index > 0 && row[index] > row[index - 1]) {
      int32_t t = row[index];
      row[index] = row[index - 1];
      row[index - 1] = t;
      --index;
    }
  }
  return 0;
}

static void build_counts_from(const mallet_model* m, int32_t** ttc, int32_t* tpt, int64_t d0,
                              int64_t d1) {
  memset(tpt, 0, sizeof(int32_t) * (size_t)m->K);
  for (int32_t w = 0; w < m->V; ++w) {
    int32_t* row = ttc[w];
    for (int32_t i = 0; i < m->ttc_len[w] && row[i] > 0; ++i) row[i] = 0;
  }
  for (int64_t i = m->doc_ptr[d0]; i < m->doc_ptr[d1]; ++i) {
    int32_t topic = m->z[i];
    tpt[topic]++;
    row_increment(ttc[m->tok[i]], m->ttc_len[m->tok[i]], topic, m->topic_mask, m->topic_bits);
  }
}

int mallet_add_instances(mallet_model* m, int64_t D, int32_t V, const int64_t* doc_ptr,
                         const int32_t* tok_word, const int32_t* z_init) {
  const int64_t addN = doc_ptr[D];
  const int64_t newD = m->D + D, newN = m->N + addN;
  m->doc_ptr = (int64_t*)realloc(m->doc_ptr, sizeof(int64_t) * (size_t)(newD + 1));
  m->tok = (int32_t*)realloc(m->tok, sizeof(int32_t) * (size_t)(newN > 0 ? newN : 1));
  m->z = (int32_t*)realloc(m->z, sizeof(int32_t) * (size_t)(newN > 0 ? newN : 1));
  if (m->D == 0) m->doc_ptr[0] = 0;
  for (int64_t d = 0; d < D; ++d) m->doc_ptr[m->D + d + 1] = m->N + doc_ptr[d + 1];
  memcpy(m->tok + m->N, tok_word, sizeof(int32_t) * (size_t)addN);
  for (int64_t i = 0; i < addN; ++i) {
    if (tok_word[i] < 0 || tok_word[i] >= V) return -1;
    m->z[m->N + i] = z_init ? z_init[i] : jr_next_int_bound(&m->random, m->K);
  }
  /* alphabet may have grown (updateModel, cmu_ron/TrainAndPredict.java:173-177) */
  if (V > m->V) {
    m->ttc = (int32_t**)realloc(m->ttc, sizeof(int32_t*) * (size_t)V);
    m->ttc_len = (int32_t*)realloc(m->ttc_len, sizeof(int32_t) * (size_t)V);
    for (int32_t w = m->V; w < V; ++w) {
      m->ttc[w] = NULL;
      m->ttc_len[w] = 0;
    }
    m->V = V;
  }
  m->D = newD;
  m->N = newN;
  m->beta_sum = m->beta * m->V;
  /* row length = min(K, corpus frequency of the type) */
  int64_t* freq = (int64_t*)calloc((size_t)m->V, sizeof(int64_t));
  for (int64_t i = 0; i < m->N; ++i) freq[m->tok[i]]++;
  for (int32_t w = 0; w < m->V; ++w) {
    int32_t len = (int32_t)(freq[w] < m->K ? freq[w] : m->K);
    free(m->ttc[w]);
    m->ttc[w] = (int32_t*)calloc((size_t)(len > 0 ? len : 1), sizeof(int32_t));
    m->ttc_len[w] = len;
  }
  free(freq);
  build_counts_from(m, m->ttc, m->tpt, 0, m->D); /* buildInitialTypeTopicCounts */
  return 0;
}

/* WorkerRunnable.sampleTopicsForOneDoc. */
static void sample_topics_for_one_doc(worker* wk, int64_t doc) {
  mallet_model* m = wk->m;
  const int32_t K = m->K, mask = m->topic_mask, bits = m->topic_bits;
  const double beta = m->beta, beta_sum = m->beta_sum;
  const double* alpha = m->alpha;
  int32_t* tpt = wk->tpt;
  double* coef = wk->cached_coefficients;
  int32_t* local = wk->local_counts;
  int32_t* lidx = wk->local_index;
  double* scores = wk->term_scores;
  const int64_t b = m->doc_ptr[doc], e = m->doc_ptr[doc + 1];

  memset(local, 0, sizeof(int32_t) * (size_t)K);
  for (int64_t i = b; i < e; ++i) local[m->z[i]]++;
  int32_t nz = 0;
  for (int32_t k = 0; k < K; ++k)
    if (local[k] != 0) lidx[nz++] = k;
  double topic_beta_mass = 0.0;
  for (int32_t di = 0; di < nz; ++di) {
    int32_t k = lidx[di], n = local[k];
    topic_beta_mass += beta * n / (tpt[k] + beta_sum);
    coef[k] = (alpha[k] + n) / (tpt[k] + beta_sum);
  }

  for (int64_t pos = b; pos < e; ++pos) {
    const int32_t type = m->tok[pos];
    const int32_t old_topic = m->z[pos];
    int32_t* row = wk->ttc[type];
    const int32_t len = m->ttc_len[type];

    /* remove the token from the document-side statistics */
    wk->smoothing_only_mass -= alpha[old_topic] * beta / (tpt[old_topic] + beta_sum);
    topic_beta_mass -= beta * local[old_topic] / (tpt[old_topic] + beta_sum);
    local[old_topic]--;
    if (local[old_topic] == 0) {
      int32_t di = 0;
      while (lidx[di] != old_topic) ++di;
      while (di < nz) {
        if (di < K - 1) lidx[di] = lidx[di + 1];
        ++di;
      }
      --nz;
    }
    tpt[old_topic]--;
    wk->smoothing_only_mass += alpha[old_topic] * beta / (tpt[old_topic] + beta_sum);
    topic_beta_mass += beta * local[old_topic] / (tpt[old_topic] + beta_sum);
    coef[old_topic] = (alpha[old_topic] + local[old_topic]) / (tpt[old_topic] + beta_sum);

    /* one walk over the type's packed row: decrement in place, score the rest */
    int32_t index = 0;
    int already_decremented = 0;
    double topic_term_mass = 0.0;
    while (index < len && row[index] > 0) {
      int32_t cur_topic = row[index] & mask;
      int32_t cur_value = row[index] >> bits;
      if (!already_decremented && cur_topic == old_topic) {
        cur_value--;
        row[index] = cur_value == 0 ? 0 : (cur_value << bits) + old_topic;
        int32_t sub = index;
        while (sub < len - 1 && row[sub] < row[sub + 1]) {
          int32_t t = row[sub];
          row[sub] = row[sub + 1];
          row[sub + 1] = t;
          ++sub;
        }
        already_decremented = 1;
      } else {
        double score = coef[cur_topic] * cur_value;
        topic_term_mass += score;
        scores[index] = score;
        ++index;
      }
    }
    const int32_t nscored = index;

    double sample = jr_next_uniform(wk->rng) * (wk->smoothing_only_mass + topic_beta_mass + topic_term_mass);
    int32_t new_topic = -1;
    if (sample < topic_term_mass) {
      int32_t i = -1;
      while (sample > 0 && i + 1 < nscored) {
        ++i;
        sample -= scores[i];
      }
      if (i < 0) i = 0;
      new_topic = row[i] & mask;
      int32_t value = row[i] >> bits;
      row[i] = ((value + 1) << bits) + new_topic;
      while (i > 0 && row[i] > row[i - 1]) {
        int32_t t = row[i];
        row[i] = row[i - 1];
        row[i - 1] = t;
        --i;
      }
    } else {
      sample -= topic_term_mass;
      if (sample < topic_beta_mass) {
        sample /= beta;
        for (int32_t di = 0; di < nz; ++di) {
          int32_t k = lidx[di];
          sample -= local[k] / (tpt[k] + beta_sum);
          if (sample <= 0.0) {
            new_topic = k;
            break;
          }
        }
      } else {
        sample -= topic_beta_mass;
        sample /= beta;
        new_topic = 0;
        sample -= alpha[new_topic] / (tpt[new_topic] + beta_sum);
        while (sample > 0.0 && new_topic < K - 1) {
          ++new_topic;
          sample -= alpha[new_topic] / (tpt[new_topic] + beta_sum);
        }
      }
      if (new_topic == -1) new_topic = K - 1; /* Mallet: "sampling error", falls back to K-1 */
      if (row_increment(row, len, new_topic, mask, bits) != 0) {
        fprintf(stderr, "mallet oracle: type-topic row overflow\n");
      }
    }

    m->z[pos] = new_topic;
    wk->smoothing_only_mass -= alpha[new_topic] * beta / (tpt[new_topic] + beta_sum);
    topic_beta_mass -= beta * local[new_topic] / (tpt[new_topic] + beta_sum);
    local[new_topic]++;
    if (local[new_topic] == 1) {
      int32_t di = nz;
      while (di > 0 && lidx[di - 1] > new_topic) {
        lidx[di] = lidx[di - 1];
        --di;
      }
      lidx[di] = new_topic;
      ++nz;
    }
    tpt[new_topic]++;
    coef[new_topic] = (alpha[new_topic] + local[new_topic]) / (tpt[new_topic] + beta_sum);
    wk->smoothing_only_mass += alpha[new_topic] * beta / (tpt[new_topic] + beta_sum);
    topic_beta_mass += beta * local[new_topic] / (tpt[new_topic] + beta_sum);
  }
  for (int32_t di = 0; di < nz; ++di) {
    int32_t k = lidx[di];
    coef[k] = alpha[k] / (tpt[k] + beta_sum);
  }
}

static void* worker_run(void* arg) {
  worker* wk = (worker*)arg;
  mallet_model* m = wk->m;
  wk->smoothing_only_mass = 0.0;
  for (int32_t k = 0; k < m->K; ++k) {
    wk->smoothing_only_mass += m->alpha[k] * m->beta / (wk->tpt[k] + m->beta_sum);
    wk->cached_coefficients[k] = m->alpha[k] / (wk->tpt[k] + m->beta_sum);
  }
  for (int64_t d = wk->start_doc; d < wk->start_doc + wk->num_docs; ++d) sample_topics_for_one_doc(wk, d);
  if (wk->owns_counts) /* buildLocalTypeTopicCounts: replica := counts of own documents only */
    build_counts_from(m, wk->ttc, wk->tpt, wk->start_doc, wk->start_doc + wk->num_docs);
  return NULL;
}

/* sumTypeTopicCounts + copy-back. */
static void sum_type_topic_counts(mallet_model* m, worker* ws, int32_t T) {
  memset(m->tpt, 0, sizeof(int32_t) * (size_t)m->K);
  for (int32_t w = 0; w < m->V; ++w) {
    int32_t* row = m->ttc[w];
    for (int32_t i = 0; i < m->ttc_len[w] && row[i] > 0; ++i) row[i] = 0;
  }
  for (int32_t t = 0; t < T; ++t) {
    for (int32_t k = 0; k < m->K; ++k) m->tpt[k] += ws[t].tpt[k];
    for (int32_t w = 0; w < m->V; ++w) {
      const int32_t* src = ws[t].ttc[w];
      int32_t* dst = m->ttc[w];
      const int32_t len = m->ttc_len[w];
      for (int32_t s = 0; s < len && src[s] > 0; ++s) {
        int32_t topic = src[s] & m->topic_mask, count = src[s] >> m->topic_bits;
        int32_t ti = 0;
        while (ti < len && dst[ti] > 0 && (dst[ti] & m->topic_mask) != topic) ++ti;
        int32_t cur = (ti < len && dst[ti] > 0) ? (dst[ti] >> m->topic_bits) : 0;
        dst[ti] = ((cur + count) << m->topic_bits) + topic;
        while (ti > 0 && dst[ti] > dst[ti - 1]) {
          int32_t x = dst[ti];
          dst[ti] = dst[ti - 1];
          dst[ti - 1] = x;
          --ti;
        }
      }
    }
  }
  for (int32_t t = 0; t < T; ++t) {
    memcpy(ws[t].tpt, m->tpt, sizeof(int32_t) * (size_t)m->K);
    for (int32_t w = 0; w < m->V; ++w)
      memcpy(ws[t].ttc[w], m->ttc[w], sizeof(int32_t) * (size_t)m->ttc_len[w]);
  }
}

void mallet_get_timers(const mallet_model* m, double* setup_s, double* sample_s, double* merge_s) {
  *setup_s = m->t_setup;
  *sample_s = m->t_sample;
  *merge_s = m->t_merge;
}

int mallet_estimate(mallet_model* m, int32_t iterations) {
  const double t_enter = now_seconds();
  const int32_t T = m->num_threads;
  worker* ws = (worker*)calloc((size_t)T, sizeof(worker));
  const int64_t docs_per_thread = m->D / T;
  int64_t offset = 0;
  for (int32_t t = 0; t < T; ++t) {
    worker* wk = &ws[t];
    wk->m = m;
    wk->start_doc = offset;
    wk->num_docs = (t == T - 1) ? (m->D - offset) : docs_per_thread;
    offset += docs_per_thread;
    wk->cached_coefficients = (double*)malloc(sizeof(double) * (size_t)m->K);
    wk->local_counts = (int32_t*)malloc(sizeof(int32_t) * (size_t)m->K);
    wk->local_index = (int32_t*)malloc(sizeof(int32_t) * (size_t)m->K);
    wk->term_scores = (double*)malloc(sizeof(double) * (size_t)m->K);
    if (T > 1) {
      wk->owns_counts = 1;
      wk->tpt = (int32_t*)malloc(sizeof(int32_t) * (size_t)m->K);
      memcpy(wk->tpt, m->tpt, sizeof(int32_t) * (size_t)m->K);
      wk->ttc = (int32_t**)malloc(sizeof(int32_t*) * (size_t)m->V);
      for (int32_t w = 0; w < m->V; ++w) {
        int32_t len = m->ttc_len[w];
        wk->ttc[w] = (int32_t*)malloc(sizeof(int32_t) * (size_t)(len > 0 ? len : 1));
        memcpy(wk->ttc[w], m->ttc[w], sizeof(int32_t) * (size_t)len);
      }
      /* every worker gets `new Randoms(randomSeed)`: identical streams when a seed is set */
      jr_seed(&wk->random, m->random_seed == -1 ? (int64_t)time(NULL) + t : (int64_t)m->random_seed);
      wk->rng = &wk->random;
    } else {
      wk->owns_counts = 0;
      wk->tpt = m->tpt;
      wk->ttc = m->ttc;
      wk->rng = &m->random; /* single thread shares the model's stream (init consumed it first) */
    }
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)T);
  m->t_setup += now_seconds() - t_enter;
  for (int32_t it = 1; it <= iterations; ++it) {
    const double t0 = now_seconds();
    if (T > 1) {
      for (int32_t t = 0; t < T; ++t) pthread_create(&th[t], NULL, worker_run, &ws[t]);
      for (int32_t t = 0; t < T; ++t) pthread_join(th[t], NULL);
      const double t1 = now_seconds();
      sum_type_topic_counts(m, ws, T);
      m->t_sample += t1 - t0;
      m->t_merge += now_seconds() - t1;
    } else {
      worker_run(&ws[0]);
      m->t_sample += now_seconds() - t0;
    }
  }
  free(th);
  for (int32_t t = 0; t < T; ++t) {
    worker* wk = &ws[t];
    free(wk->cached_coefficients);
    free(wk->local_counts);
    free(wk->local_index);
    free(wk->term_scores);
    if (wk->owns_counts) {
      free(wk->tpt);
      for (int32_t w = 0; w < m->V; ++w) free(wk->ttc[w]);
      free(wk->ttc);
    }
  }
  free(ws);
  return 0;
}

double mallet_model_log_likelihood(const mallet_model* m) {
  const int32_t K = m->K;
  double ll = 0.0;
  int32_t* counts = (int32_t*)calloc((size_t)K, sizeof(int32_t));
  double* lga = (double*)malloc(sizeof(double) * (size_t)K);
  double alpha_sum = 0.0;
  for (int32_t k = 0; k < K; ++k) {
    lga[k] = oracle_log_gamma_stirling(m->alpha[k]);
    alpha_sum += m->alpha[k];
  }
  for (int64_t d = 0; d < m->D; ++d) {
    for (int64_t i = m->doc_ptr[d]; i < m->doc_ptr[d + 1]; ++i) counts[m->z[i]]++;
    for (int32_t k = 0; k < K; ++k)
      if (counts[k] > 0) {
        ll += oracle_log_gamma_stirling(m->alpha[k] + counts[k]) - lga[k];
        counts[k] = 0;
      }
    ll -= oracle_log_gamma_stirling(alpha_sum + (double)(m->doc_ptr[d + 1] - m->doc_ptr[d]));
  }
  ll += (double)m->D * oracle_log_gamma_stirling(alpha_sum);
  int64_t nonzero = 0;
  for (int32_t w = 0; w < m->V; ++w) {
    const int32_t* row = m->ttc[w];
    for (int32_t i = 0; i < m->ttc_len[w] && row[i] > 0; ++i) {
      ++nonzero;
      ll += oracle_log_gamma_stirling(m->beta + (double)(row[i] >> m->topic_bits));
    }
  }
  for (int32_t k = 0; k < K; ++k) ll -= oracle_log_gamma_stirling(m->beta * m->V + m->tpt[k]);
  ll += (double)K * oracle_log_gamma_stirling(m->beta * m->V);
  ll -= (double)nonzero * oracle_log_gamma_stirling(m->beta);
  free(counts);
  free(lga);
  return ll;
}

void mallet_get_assignments(const mallet_model* m, int32_t* z) {
  memcpy(z, m->z, sizeof(int32_t) * (size_t)m->N);
}

void mallet_get_counts(const mallet_model* m, int32_t* nwk, int32_t* nk) {
  memset(nwk, 0, sizeof(int32_t) * (size_t)m->V * (size_t)m->K);
  for (int32_t w = 0; w < m->V; ++w) {
    const int32_t* row = m->ttc[w];
    for (int32_t i = 0; i < m->ttc_len[w] && row[i] > 0; ++i)
      nwk[(size_t)w * m->K + (row[i] & m->topic_mask)] = row[i] >> m->topic_bits;
  }
  memcpy(nk, m->tpt, sizeof(int32_t) * (size_t)m->K);
}

void mallet_get_topic_probabilities(const mallet_model* m, int64_t doc, double* theta) {
  oracle_theta(m->K, m->z + m->doc_ptr[doc], m->doc_ptr[doc + 1] - m->doc_ptr[doc], m->alpha, theta);
}

/* TopicInferencer.getSampledDistribution: frozen rows and n_k. */
void mallet_infer(const mallet_model* m, const int32_t* words, int32_t n, int32_t iters,
                  int32_t thinning, int32_t burn_in, int32_t seed, double* theta) {
  const int32_t K = m->K, mask = m->topic_mask, bits = m->topic_bits;
  const double beta = m->beta, beta_sum = m->beta_sum;
  const double* alpha = m->alpha;
  java_random rng;
  jr_seed(&rng, seed == -1 ? (int64_t)time(NULL) : (int64_t)seed);
  int32_t* tok = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  int32_t* z = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  int32_t len = 0;
  for (int32_t i = 0; i < n; ++i)
    if (words[i] >= 0 && words[i] < m->V) tok[len++] = words[i]; /* unknown types dropped */
  int32_t* local = (int32_t*)calloc((size_t)K, sizeof(int32_t));
  int32_t* lidx = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
  double* coef = (double*)malloc(sizeof(double) * (size_t)K);
  double* scores = (double*)malloc(sizeof(double) * (size_t)K);
  double smoothing_only_mass = 0.0;
  for (int32_t k = 0; k < K; ++k) {
    smoothing_only_mass += alpha[k] * beta / (m->tpt[k] + beta_sum);
    coef[k] = alpha[k] / (m->tpt[k] + beta_sum);
  }
  for (int32_t i = 0; i < len; ++i) {
    z[i] = jr_next_int_bound(&rng, K);
    local[z[i]]++;
  }
  int32_t nz = 0;
  for (int32_t k = 0; k < K; ++k)
    if (local[k] != 0) lidx[nz++] = k;
  double topic_beta_mass = 0.0;
  for (int32_t di = 0; di < nz; ++di) {
    int32_t k = lidx[di];
    topic_beta_mass += beta * local[k] / (m->tpt[k] + beta_sum);
    coef[k] = (alpha[k] + local[k]) / (m->tpt[k] + beta_sum);
  }
  for (int32_t k = 0; k < K; ++k) theta[k] = 0.0;
  double sum = 0.0;
  for (int32_t it = 1; it <= iters; ++it) {
    for (int32_t pos = 0; pos < len; ++pos) {
      const int32_t type = tok[pos], old_topic = z[pos];
      const int32_t* row = m->ttc[type];
      const int32_t rlen = m->ttc_len[type];
      topic_beta_mass -= beta * local[old_topic] / (m->tpt[old_topic] + beta_sum);
      local[old_topic]--;
      if (local[old_topic] == 0) {
        int32_t di = 0;
        while (lidx[di] != old_topic) ++di;
        while (di < nz) {
          if (di < K - 1) lidx[di] = lidx[di + 1];
          ++di;
        }
        --nz;
      }
      topic_beta_mass += beta * local[old_topic] / (m->tpt[old_topic] + beta_sum);
      coef[old_topic] = (alpha[old_topic] + local[old_topic]) / (m->tpt[old_topic] + beta_sum);
      double topic_term_mass = 0.0;
      int32_t index = 0;
      while (index < rlen && row[index] > 0) {
        double score = coef[row[index] & mask] * (row[index] >> bits);
        topic_term_mass += score;
        scores[index] = score;
        ++index;
      }
      double sample = jr_next_uniform(&rng) * (smoothing_only_mass + topic_beta_mass + topic_term_mass);
      int32_t new_topic = -1;
      if (sample < topic_term_mass) {
        int32_t i = -1;
        while (sample > 0 && i + 1 < index) {
          ++i;
          sample -= scores[i];
        }
        if (i < 0) i = 0;
        new_topic = row[i] & mask;
      } else {
        sample -= topic_term_mass;
        if (sample < topic_beta_mass) {
          sample /= beta;
          for (int32_t di = 0; di < nz; ++di) {
            int32_t k = lidx[di];
            sample -= local[k] / (m->tpt[k] + beta_sum);
            if (sample <= 0.0) {
              new_topic = k;
              break;
            }
          }
        } else {
          sample -= topic_beta_mass;
          sample /= beta;
          new_topic = 0;
          sample -= alpha[new_topic] / (m->tpt[new_topic] + beta_sum);
          while (sample > 0.0 && new_topic < K - 1) {
            ++new_topic;
            sample -= alpha[new_topic] / (m->tpt[new_topic] + beta_sum);
          }
        }
        if (new_topic == -1) new_topic = K - 1;
      }
      z[pos] = new_topic;
      topic_beta_mass -= beta * local[new_topic] / (m->tpt[new_topic] + beta_sum);
      local[new_topic]++;
      if (local[new_topic] == 1) {
        int32_t di = nz;
        while (di > 0 && lidx[di - 1] > new_topic) {
          lidx[di] = lidx[di - 1];
          --di;
        }
        lidx[di] = new_topic;
        ++nz;
      }
      coef[new_topic] = (alpha[new_topic] + local[new_topic]) / (m->tpt[new_topic] + beta_sum);
      topic_beta_mass += beta * local[new_topic] / (m->tpt[new_topic] + beta_sum);
    }
    if (it > burn_in && (it - burn_in) % thinning == 0) {
      for (int32_t k = 0; k < K; ++k) {
        theta[k] += alpha[k] + local[k];
        sum += alpha[k] + local[k];
      }
    }
  }
  if (sum == 0.0) {
    for (int32_t k = 0; k < K; ++k) {
      theta[k] = alpha[k] + local[k];
      sum += theta[k];
    }
  }
  for (int32_t k = 0; k < K; ++k) theta[k] /= sum;
  free(tok);
  free(z);
  free(local);
  free(lidx);
  free(coef);
  free(scores);
}

/* ---- hyper-parameter optimisation (enabled by setOptimizeInterval(20) at
 * reference cmu_ron/TrainAndPredict.java:163, cmu/TrainAndPredict.java:261) -------------------- */

/* Mallet's Dirichlet.digamma: recurrence up to z >= 6-ish then the asymptotic series. */
double oracle_digamma(double z) {
  const double EULER_MASCHERONI = -0.5772156649015328606065121;
  const double DIGAMMA_COEF_1 = 1.0 / 12, DIGAMMA_COEF_2 = 1.0 / 120, DIGAMMA_COEF_3 = 1.0 / 252,
               DIGAMMA_COEF_4 = 1.0 / 240, DIGAMMA_COEF_5 = 1.0 / 132, DIGAMMA_COEF_6 = 691.0 / 32760,
               DIGAMMA_COEF_7 = 1.0 / 12;
  const double DIGAMMA_LARGE = 9.5, DIGAMMA_SMALL = .000001;
  double psi = 0;
  if (z < DIGAMMA_SMALL) return EULER_MASCHERONI - (1 / z);
  while (z < DIGAMMA_LARGE) {
    psi -= 1 / z;
    z++;
  }
  double invZ = 1 / z, invZSquared = invZ * invZ;
  psi += log(z) - .5 * invZ -
         invZSquared * (DIGAMMA_COEF_1 - invZSquared * (DIGAMMA_COEF_2 - invZSquared * (DIGAMMA_COEF_3 -
         invZSquared * (DIGAMMA_COEF_4 - invZSquared * (DIGAMMA_COEF_5 - invZSquared * (DIGAMMA_COEF_6 -
         invZSquared * DIGAMMA_COEF_7))))));
  return psi;
}

/* Dirichlet.learnParameters(parameters, observations, observationLengths, 1.00001, 1.0, 200):
 * Minka/Wallach fixed point on histograms. observations is K rows of `width` counts
 * (observations[k][n] = documents in which topic k occurs n times), observationLengths[n] =
 * documents of length n. parameters (alpha) is updated in place; returns their sum. */
double oracle_learn_parameters(double* parameters, int32_t K, const int32_t* observations, int32_t width,
                               const int32_t* observation_lengths, double shape, double scale,
                               int32_t num_iterations) {
  double parameters_sum = 0;
  for (int32_t k = 0; k < K; ++k) parameters_sum += parameters[k];
  int32_t* non_zero_limits = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
  for (int32_t k = 0; k < K; ++k) {
    non_zero_limits[k] = -1;
    const int32_t* h = observations + (size_t)k * width;
    for (int32_t n = 0; n < width; ++n)
      if (h[n] > 0) non_zero_limits[k] = n;
  }
  for (int32_t it = 0; it < num_iterations; ++it) {
    double denominator = 0, current_digamma = 0;
    for (int32_t i = 1; i < width; ++i) {
      current_digamma += 1 / (parameters_sum + i - 1);
      denominator += observation_lengths[i] * current_digamma;
    }
    denominator -= 1 / scale;
    parameters_sum = 0;
    for (int32_t k = 0; k < K; ++k) {
      const int32_t limit = non_zero_limits[k];
      const double old = parameters[k];
      const int32_t* h = observations + (size_t)k * width;
      double acc = 0;
      current_digamma = 0;
      for (int32_t i = 1; i <= limit; ++i) {
        current_digamma += 1 / (old + i - 1);
        acc += h[i] * current_digamma;
      }
      parameters[k] = old * (acc + shape) / denominator;
      parameters_sum += parameters[k];
    }
  }
  free(non_zero_limits);
  return parameters_sum;
}

/* Dirichlet.learnSymmetricConcentration(countHistogram, observationLengths, numDimensions,
 * currentValue): count_histogram[c] = (type, topic) cells holding c tokens, observation_lengths[n]
 * = topics holding n tokens; returns the new betaSum. */
double oracle_learn_symmetric_concentration(const int64_t* count_histogram, int64_t n_counts,
                                            const int64_t* observation_lengths, int64_t n_lengths,
                                            int32_t num_dimensions, double current_value) {
  int64_t largest_non_zero_count = 0;
  for (int64_t i = 0; i < n_counts; ++i)
    if (count_histogram[i] > 0) largest_non_zero_count = i;
  int64_t* non_zero_length_index = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_lengths > 0 ? n_lengths : 1));
  int64_t dense_size = 0;
  for (int64_t i = 0; i < n_lengths; ++i)
    if (observation_lengths[i] > 0) non_zero_length_index[dense_size++] = i;
  for (int iteration = 1; iteration <= 200; ++iteration) {
    const double current_parameter = current_value / num_dimensions;
    double current_digamma = 0, numerator = 0;
    for (int64_t index = 1; index <= largest_non_zero_count; ++index) {
      current_digamma += 1.0 / (current_parameter + index - 1);
      numerator += count_histogram[index] * current_digamma;
    }
    current_digamma = 0;
    double denominator = 0;
    int64_t previous_length = 0;
    const double cached_digamma = oracle_digamma(current_value);
    for (int64_t di = 0; di < dense_size; ++di) {
      const int64_t length = non_zero_length_index[di];
      if (length - previous_length > 20) {
        current_digamma = oracle_digamma(current_value + length) - cached_digamma;
      } else {
        for (int64_t index = previous_length; index < length; ++index) current_digamma += 1.0 / (current_value + index);
      }
      denominator += current_digamma * observation_lengths[length];
      previous_length = length; /* (Mallet omits this update; with it the shortcut is exact) */
    }
    current_value = current_parameter * numerator / denominator;
  }
  free(non_zero_length_index);
  return current_value;
}
