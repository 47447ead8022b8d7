/*
 * oracle/lda_oracle.h — TEST INFRASTRUCTURE. CPU oracle for the collapsed-Gibbs LDA hot path.
 *
 * PARITY UNPINNED: the reference (qianjinding/LDAGibbsSampling) holds no tests, golden vectors
 * or recorded outputs for this path, and the arithmetic lives in the un-vendored dependency
 * cc.mallet:mallet:2.0.7 (reference pom.xml:107-111) which is absent from /root/reference and
 * cannot be run here (no JVM). This oracle therefore restates
 *   (a) Mallet 2.0.7's published SparseLDA / AD-LDA semantics (mallet_sparse_lda.c), anchored on
 *       the reference call sites cmu_ron/TrainAndPredict.java:159-171 and
 *       cmu/TrainAndPredict.java:258-269, and
 *   (b) the sampling spec of the B200 kernels (spec_sampler.c; DESIGN.md "sampling spec"), i.e.
 *       the textbook conditional (n_wk+b)(n_dk+a)/(n_k+Vb) evaluated in a fixed fp32 order.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product (libb200lda.so) never links or calls it.
 */
#ifndef B200LDA_ORACLE_H
#define B200LDA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic corpus (SURVEY.md §8(d) generator), corpus_gen.c ---- */
/* Returns N (token count). tok_word may be NULL to query N only (doc_ptr is still filled). */
int64_t oracle_gen_corpus(int64_t D, int32_t V, double mean_len, int32_t k_true, uint64_t seed,
                          int64_t* doc_ptr, int32_t* tok_word, int64_t tok_cap);

/* ---- RNG probes (pinned by known answers) ---- */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void oracle_java_random_ints(int64_t seed, int32_t n, int32_t bound, int32_t* out);
void oracle_java_random_uniforms(int64_t seed, int32_t n, double* out);

/* ---- shared helpers, spec_sampler.c ---- */
/* z[i] = floor(K * x0 / 2^32), x0 = Philox(token = global_off + i, sweep = 0, stream = 1). */
void oracle_init_z_philox(int64_t N, int32_t K, uint64_t seed, int64_t global_off, int32_t* z);
void oracle_count(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr, const int32_t* tok_word,
                  const int32_t* z, int32_t* nwk /*V*K*/, int32_t* nk /*K*/);
/* Sorted sparse doc rows: row d occupies [row_ptr[d], row_ptr[d]+nnz[d]); row_ptr[d] = sum of
 * min(K, L_d') over d' < d. */
void oracle_ndk_csr(int64_t D, int32_t K, const int64_t* doc_ptr, const int32_t* z,
                    int64_t* row_ptr /*D+1*/, int32_t* nnz /*D*/, int32_t* topic, int32_t* count);

/* fp32 32-lane Kogge-Stone tile scan with sequential carry: the scan order of the prior rows. */
void oracle_tile_scan_f32(const float* in, int64_t n, float* out);

/* Doc-bucket prefix order: slot j on lane j mod 32; each lane sums its slots tile by tile, one
 * Kogge-Stone scan over the 32 lane totals; *total = scanned total of lane 31. */
void oracle_lane_strided_prefix_f32(const float* in, int64_t n, float* out, float* total);

/* Per-sweep tables from the sweep-start snapshot. prior is V*K inclusive prefix rows. */
void oracle_spec_tables(int32_t V, int32_t K, const int32_t* nwk, const int32_t* nk,
                        const double* alpha, double beta, float* invden /*K*/, float* ab /*K*/,
                        float* prior /*V*K*/, float* q /*V*/);
/* Hierarchical (fan-out 32) search: index the spec selects in one prefix row for target s. */
int32_t oracle_spec_hsearch(const float* row, int32_t K, float s);

/* One token under the spec. slots = doc's sorted non-zero topics INCLUDING the current token.
 * Returns the new topic. */
int32_t oracle_spec_select(int32_t K, const int32_t* slot_topic, const int32_t* slot_count,
                           int32_t nslots, const int32_t* nwk_row /*K ints of word w*/,
                           const float* invden, const float* ab, const float* prior_row, float q_w,
                           float beta_f, int32_t old_topic, float u);

/* FROZEN mode: no count moves at all; every token sees the snapshot (own token excluded).
 * uniforms may be NULL (Philox(seed, global_off+i, sweep, 0)). */
void oracle_spec_frozen(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr,
                        const int32_t* tok_word, const int32_t* z_in, const double* alpha,
                        double beta, uint64_t seed, uint32_t sweep, int64_t global_off,
                        const float* uniforms, int32_t* z_out);

/* DEFERRED mode chain: n_wk / n_k frozen per sweep, doc rows live inside a document.
 * Sweeps are numbered first_sweep .. first_sweep+n_sweeps-1. z updated in place.
 * Result is independent of how documents are sharded (Philox is keyed by global token). */
void oracle_spec_sweeps(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr,
                        const int32_t* tok_word, int32_t* z, const double* alpha, double beta,
                        uint64_t seed, uint32_t first_sweep, int32_t n_sweeps, int64_t global_off);

/* live != 0: sequential rendering of LIVE mode (n_wk moves immediately; tables stay stale). */
void oracle_spec_sweeps_mode(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr,
                             const int32_t* tok_word, int32_t* z, const double* alpha, double beta,
                             uint64_t seed, uint32_t first_sweep, int32_t n_sweeps,
                             int64_t global_off, int32_t live);

/* One sweep of a document set against GIVEN (global) counts; see spec_sampler.c. */
void oracle_spec_sweep_given_counts(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr,
                                    const int32_t* tok_word, int32_t* z, int32_t* nwk, const int32_t* nk,
                                    const double* alpha, double beta, uint64_t seed, uint32_t sweep,
                                    int64_t global_off, int32_t live, int32_t exclude_self,
                                    int32_t* delta_nwk, int32_t* delta_nk);
/* Held-out inference under the spec (frozen counts); theta is D*K. */
void oracle_spec_infer(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr, const int32_t* tok_word,
                       const int32_t* nwk, const int32_t* nk, const double* alpha, double beta,
                       int32_t iterations, int32_t thinning, int32_t burn_in, uint64_t seed, double* theta);

/* Exact (double) conditional of the textbook formula for one token, n_k NOT excluding the
 * token (as in the spec), own token excluded from n_wk and n_dk. p has K entries, sums to 1. */
void oracle_exact_conditional(int32_t K, int32_t V, const int32_t* ndk_dense /*K*/,
                              const int32_t* nwk_row /*K*/, const int32_t* nk, const double* alpha,
                              double beta, int32_t old_topic, double* p);

/* modelLogLikelihood. stirling != 0 uses Mallet's logGammaStirling. */
double oracle_loglik(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr,
                     const int32_t* tok_word, const int32_t* z, const double* alpha, double beta,
                     int32_t stirling);
double oracle_log_gamma_stirling(double z);

/* theta_d (getTopicProbabilities) and phi (K x V, row-major by topic). */
void oracle_theta(int32_t K, const int32_t* z_doc, int64_t len, const double* alpha, double* out);
void oracle_phi(int32_t V, int32_t K, const int32_t* nwk, const int32_t* nk, double beta,
                double* out /*K*V*/);

/* ---- Mallet-faithful SparseLDA + AD-LDA threads, mallet_sparse_lda.c ---- */
typedef struct mallet_model mallet_model;
mallet_model* mallet_create(int32_t K, double alpha_sum, double beta);
void mallet_destroy(mallet_model* m);
void mallet_set_random_seed(mallet_model* m, int32_t seed);
void mallet_set_num_threads(mallet_model* m, int32_t t);
/* addInstances: z_init may be NULL (java.util.Random nextInt(K) per token, Mallet's init). */
int mallet_add_instances(mallet_model* m, int64_t D, int32_t V, const int64_t* doc_ptr,
                         const int32_t* tok_word, const int32_t* z_init);
/* estimate(): n iterations, hyper-parameter optimisation off (optimizeInterval = 0). */
int mallet_estimate(mallet_model* m, int32_t iterations);
/* cumulative wall-clock seconds of estimate(): worker set-up, sampling phase, merge (T > 1) */
void mallet_get_timers(const mallet_model* m, double* setup_s, double* sample_s, double* merge_s);
double mallet_model_log_likelihood(const mallet_model* m);
void mallet_get_assignments(const mallet_model* m, int32_t* z);
void mallet_get_counts(const mallet_model* m, int32_t* nwk /*V*K dense*/, int32_t* nk);
void mallet_get_topic_probabilities(const mallet_model* m, int64_t doc, double* theta);
/* TopicInferencer.getSampledDistribution(instance, iters, thinning, burnIn). */
void mallet_infer(const mallet_model* m, const int32_t* words, int32_t n, int32_t iters,
                  int32_t thinning, int32_t burn_in, int32_t seed, double* theta);
int64_t mallet_num_tokens(const mallet_model* m);

/* Hyper-parameter optimisation pieces of cc.mallet.types.Dirichlet. */
double oracle_digamma(double z);
double oracle_learn_parameters(double* parameters, int32_t K, const int32_t* observations, int32_t width,
                               const int32_t* observation_lengths, double shape, double scale,
                               int32_t num_iterations);
double oracle_learn_symmetric_concentration(const int64_t* count_histogram, int64_t n_counts,
                                            const int64_t* observation_lengths, int64_t n_lengths,
                                            int32_t num_dimensions, double current_value);

#ifdef __cplusplus
}
#endif
#endif
