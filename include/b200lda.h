/*
 * b200lda.h — C ABI of libb200lda.so: a B200-native (sm_100a) collapsed-Gibbs LDA sampler.
 *
 * Drop-in boundary. The reference (qianjinding/LDAGibbsSampling) has no FFI of its own; its hot
 * path is the used subset of Mallet 2.0.7's cc.mallet.topics.ParallelTopicModel /
 * TopicInferencer (un-vendored jar, reference pom.xml:107-111) driven from exactly two places:
 * cmu_ron/TrainAndPredict.java:159-177 and cmu/TrainAndPredict.java:258-274. Every entry point
 * below names the Java call it stands behind; java/B200TopicModel.java (Panama FFM) and
 * ldagibbssampling_b200/topic_model.py (ctypes) bind exactly these symbols.
 *
 * Conventions
 *  - plain C, no exceptions across the boundary; every call returns B200LDA_OK (0) or a negative
 *    b200lda_status; b200lda_last_error() gives the message of the calling thread's last failure;
 *  - every pointer argument is a HOST pointer unless its name starts with d_; input buffers are
 *    borrowed for the duration of the call only, output buffers are caller-allocated and
 *    caller-owned;
 *  - a context is single-caller (one host thread drives it) and owns one GPU; AD-LDA across
 *    GPUs = one context per GPU (one process per GPU, or several contexts in one process) plus
 *    one integer all-reduce per sweep between b200lda_sweep_begin and b200lda_sweep_end;
 *  - there is NO CPU fallback: b200lda_create fails with B200LDA_ENODEV without an sm_100 GPU.
 *
 * Sampling semantics ("sampling spec"): per token
 *      p(z=k) ∝ (n_wk^{-i} + beta) (n_dk^{-i} + alpha_k) / (n_k + V beta)
 * i.e. the conditional Mallet's WorkerRunnable.sampleTopicsForOneDoc draws from, evaluated as
 * a sparse doc bucket n_dk(n_wk+beta)/(n_k+V beta) over the document's non-zero topics plus a
 * per-word prior bucket alpha_k(n_wk+beta)/(n_k+V beta) resolved through a prefix table built at
 * the start of the sweep; n_k is the sweep-start snapshot. Randomness: Philox4x32-10,
 * key = seed, counter = (global token index, sweep, stream) — independent of sharding.
 */
#ifndef B200LDA_H
#define B200LDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LDA_ABI_VERSION 2

typedef enum {
  B200LDA_OK = 0,
  B200LDA_EINVAL = -1,   /* bad argument (IllegalArgumentException in the Java shim)      */
  B200LDA_ENODEV = -2,   /* no sm_100 device / device ordinal out of range                */
  B200LDA_ENOMEM = -3,   /* device or pinned-host allocation failed                       */
  B200LDA_ECUDA = -4,    /* a CUDA call or kernel failed; context is unusable afterwards  */
  B200LDA_ESTATE = -5,   /* call order violated (e.g. sweep before load_corpus)           */
  B200LDA_ERANGE = -6    /* word id >= V, topic >= K, document longer than 65535 tokens   */
} b200lda_status;

typedef enum {
  /* n_wk is updated in place with integer atomics while other documents read it (closest to
   * Mallet's single-thread chain; run-to-run non-deterministic across documents). */
  B200LDA_MODE_LIVE = 0,
  /* n_wk / n_k are frozen for the whole sweep and the moves are applied at its end (AD-LDA with
   * one "processor" per document). Bit-reproducible, independent of the number of GPUs, and
   * reproduced exactly by the CPU oracle. */
  B200LDA_MODE_DEFERRED = 1
} b200lda_mode;

typedef struct b200lda_ctx b200lda_ctx;

typedef struct {
  int32_t struct_size;          /* = sizeof(b200lda_config); guards ABI drift                  */
  int32_t num_topics;           /* K, 1..65536       ParallelTopicModel(int K, ...)            */
  int32_t num_types;            /* V = alphabet size at addInstances                          */
  int32_t mode;                 /* b200lda_mode                                               */
  double alpha_sum;             /* ParallelTopicModel ctor arg 2 (alpha_k = alpha_sum / K)    */
  double beta;                  /* ParallelTopicModel ctor arg 3                              */
  uint64_t seed;                /* setRandomSeed(int) analogue (Philox key)                   */
  int32_t device;               /* CUDA ordinal                                               */
  int32_t rank;                 /* AD-LDA shard id, 0..world_size-1                           */
  int32_t world_size;           /* number of shards (setNumThreads(n) analogue), >= 1         */
  int32_t table_refresh;        /* LIVE mode: rebuilds of the per-sweep tables per sweep (1..32);
                                   0 = auto: as many as cost <= ~5 % of a sweep, 16 at most    */
  int64_t global_token_offset;  /* global index of this shard's first token (Philox counter)  */
  int64_t global_doc_offset;    /* global index of this shard's first document                */
  void* stream;                 /* cudaStream_t to enqueue on, or NULL to create a private one */
} b200lda_config;

typedef struct {
  int64_t num_docs, num_tokens;
  int64_t sweeps_done;            /* total sweeps sampled so far                              */
  int64_t kernel_launches;        /* kernels launched by this context since create            */
  int64_t tokens_sampled;         /* sum over sweeps of tokens resampled                      */
  double last_sweep_ms;           /* device time of the last sweep (tables+sample+finish)     */
  double last_tables_ms;          /* ... of which per-sweep table build                       */
  double last_sample_ms;          /* ... of which the sampling kernel                         */
  double last_finish_ms;          /* ... of which delta/apply                                 */
  double mean_doc_topics;         /* token-weighted mean #non-zero doc topics (K_d bar)       */
  int64_t tokens_moved_last;      /* tokens whose topic changed in the last sweep             */
  int64_t prior_bucket_last;      /* tokens resolved through the prior table in the last sweep */
  int64_t device_bytes;           /* device memory held                                       */
  int32_t smem_bytes_per_cta, warps_per_cta, ctas, slot_capacity;
  /* accumulated since create / b200lda_reset_stats (device times from CUDA events) */
  int64_t cum_sweeps;
  double cum_tables_ms, cum_sample_ms, cum_finish_ms;
  int64_t cum_tokens_moved, cum_prior_bucket;
  int64_t cum_doc_topics;         /* sum over sampled tokens of the document's non-zero topics */
  /* row-width classes: slot_capacity/ctas/smem above describe the narrowest class (the bulk);
   * these the widest one (the longest documents): its document count, row slots, CTAs */
  int64_t long_docs;
  int32_t long_slot_capacity, long_ctas;
  int32_t row_classes;
  /* LIVE mode table refresh: rebuilds per sweep used by the last sweep (config table_refresh or
   * the auto choice), the number of hot words, and the prior rows rebuilt during the last sweep */
  int32_t table_refresh_last;
  int64_t hot_words, rows_refreshed_last;
} b200lda_stats;

/* Message of the calling thread's most recent failure ("" if none). Never NULL. */
const char* b200lda_last_error(void);
int b200lda_abi_version(void);
/* Number of visible sm_100 devices (0 if none / no driver); never fails. */
int b200lda_device_count(void);

/* new ParallelTopicModel(K, alphaSum, beta)                cmu_ron/TrainAndPredict.java:160,
 *                                                          cmu/TrainAndPredict.java:259 */
int b200lda_create(const b200lda_config* cfg, b200lda_ctx** out);
void b200lda_destroy(b200lda_ctx* ctx);

/* model.addInstances(InstanceList) — corpus half           cmu_ron/TrainAndPredict.java:162,
 * Flattened FeatureSequences: doc d = tok_word[doc_ptr[d] .. doc_ptr[d+1]).   cmu/…:260
 * Copies to the device and packs the doc->token CSR order (row offsets, longest-first visiting
 * order and CSR/word-id validation are computed on the device). Replaces any
 * corpus already loaded (updateModel: pass old + new documents, cmu_ron/…:173-177). */
int b200lda_load_corpus(b200lda_ctx* ctx, int64_t num_docs, const int64_t* doc_ptr,
                        const int32_t* tok_word);
/* model.addInstances(InstanceList) — assignment half: z == NULL draws the uniform random
 * initial topics on the device (Mallet: random.nextInt(K) per token); z != NULL installs the
 * caller's topics (seed-matched init from another sampler, or a restored checkpoint). Builds
 * n_wk, n_k and the sparse n_dk rows. */
int b200lda_init_assignments(b200lda_ctx* ctx, const int32_t* z);
/* The same with the topics in the device's own width (K <= 65536): half the bytes over the bus
 * and no widening pass. For hosts that are not bound to Java's int[] (z must not be NULL). */
int b200lda_init_assignments_u16(b200lda_ctx* ctx, const uint16_t* z);

/* model.estimate() for n sweeps (setNumIterations(n))      cmu_ron/TrainAndPredict.java:165-166,
 * Single-shard contexts only (world_size == 1); blocks until done.            cmu/…:263,265 */
int b200lda_sweep(b200lda_ctx* ctx, int32_t n);

/* One AD-LDA sweep of a multi-shard model with the caller's own all-reduce (setNumThreads(n),
 * cmu_ron/…:164, cmu/…:262; replaces WorkerRunnable.run + ParallelTopicModel.sumTypeTopicCounts):
 *   sweep_begin  enqueues table build + sampling;
 *   the caller sums the exchange buffer (int32, *count elements, device memory) over all shards
 *     in place — ncclAllReduce(ncclInt32, ncclSum) on the context's stream, or any equivalent;
 *     the buffer is the shard's n_wk replica itself ("sweep-start counts + own moves", n_k moves
 *     in the K-cell tail): nothing is copied to form it;
 *   sweep_end    turns the sum into the new global counts (sum - (n-1) x sweep-start counts).
 * Neither call synchronises; use b200lda_synchronize. Also valid with world_size == 1. */
int b200lda_sweep_begin(b200lda_ctx* ctx);
int b200lda_exchange_buffer(b200lda_ctx* ctx, void** d_buf, int64_t* count);
int b200lda_sweep_end(b200lda_ctx* ctx);
/* Once after b200lda_init_assignments on every shard of a multi-shard model (Mallet's
 * sumTypeTopicCounts at start-up): each shard built n_wk / n_k from its own documents only;
 *   counts_sync_begin  puts its n_k into the exchange buffer's tail,
 *   the caller sums the exchange buffer over all shards (same all-reduce as per sweep),
 *   counts_sync_end    installs the sum as the n_wk / n_k replica and its sweep-start snapshot. */
int b200lda_counts_sync_begin(b200lda_ctx* ctx);
int b200lda_counts_sync_end(b200lda_ctx* ctx);
int b200lda_synchronize(b200lda_ctx* ctx);
int b200lda_get_stream(b200lda_ctx* ctx, void** stream);
/* The same all-reduce(sum) done by the library itself for hosts that keep all n shards in ONE
 * process (a JVM after setNumThreads(4), the C++ mirror): sums the exchange buffers (or the
 * hyper-parameter histograms) of the n contexts in place through peer copies. Blocking; the
 * NCCL path above is the fast one. */
#define B200LDA_NCCL_ID_BYTES 128
#define B200LDA_BUFFER_EXCHANGE 0
#define B200LDA_BUFFER_HYPER 1
int b200lda_group_allreduce(b200lda_ctx** ctxs, int32_t n, int32_t which);

/* NCCL inside the library (loaded at run time; B200LDA_ENODEV when there is no libnccl.so.2).
 * setNumThreads(n) of the reference (cmu_ron/TrainAndPredict.java:164, cmu/TrainAndPredict.java:262)
 * = n shards = n GPUs, either
 *  - one process per GPU: rank 0 calls b200lda_nccl_unique_id, ships the 128 bytes to the other
 *    ranks by whatever channel the host has, every rank calls b200lda_comm_init on its context;
 *    from then on b200lda_sweep(ctx, n) runs whole AD-LDA sweeps (collective: every rank calls it)
 *    and b200lda_group_sync_counts(&ctx, 1) is the start-up count sum; or
 *  - all n contexts in ONE process (a JVM, the C++ mirror): create them with rank i / world_size n
 *    on n different devices, call b200lda_group_comm_init once, then b200lda_group_sync_counts
 *    after the assignments are installed and b200lda_group_sweep(ctxs, n, sweeps) to sample: one
 *    host thread drives every shard, the all-reduces are grouped (ncclGroupStart/End).
 * The exchange is in place: every shard all-reduces "sweep-start counts + its own moves" (the
 * exchange buffer IS its n_wk replica, n_k moves in the K-cell tail) and subtracts (n-1) x the
 * sweep-start counts; it runs in 8 slabs, slab i applied while slab i+1 is reduced.
 * Without communicators (several contexts on one device) the group calls fall back to
 * b200lda_group_allreduce's peer copies. */
int b200lda_nccl_unique_id(void* id /* B200LDA_NCCL_ID_BYTES */);
int b200lda_comm_init(b200lda_ctx* ctx, const void* id);
int b200lda_group_comm_init(b200lda_ctx** ctxs, int32_t n);
int b200lda_group_sync_counts(b200lda_ctx** ctxs, int32_t n);
int b200lda_group_sweep(b200lda_ctx** ctxs, int32_t n, int32_t sweeps);

/* Frozen-snapshot parity mode (north star; no Java counterpart): resample every token once
 * against the current counts WITHOUT moving any count. uniforms == NULL uses Philox with the
 * given sweep number; otherwise uniforms[i] in [0,1) is token i's draw. */
int b200lda_sample_frozen(b200lda_ctx* ctx, const float* uniforms, uint32_t sweep, int32_t* z_out);

/* model.getInferencer().getSampledDistribution(instance, iterations, thinning, burnIn)
 *                                                          cmu_ron/TrainAndPredict.java:144,
 * batched: num_docs held-out documents in one device pass against the frozen   cmu/…:114
 * n_wk / n_k / alpha / beta of this context (word ids must be < V: drop unknown types first, as
 * Mallet does). theta: num_docs * K doubles, row d = normalised average of (alpha_k + n_dk) over
 * the samples kept at iterations it > burn_in with (it - burn_in) % thinning == 0 (the final
 * state if none). Does not touch the training chain. */
int b200lda_infer(b200lda_ctx* ctx, int64_t num_docs, const int64_t* doc_ptr, const int32_t* tok_word,
                  int32_t iterations, int32_t thinning, int32_t burn_in, uint64_t seed, double* theta);

/* model.modelLogLikelihood()                               cmu_ron/TrainAndPredict.java:234,
 * world_size > 1: *out is this shard's document part plus the (replicated)    cmu/…:436
 * word/topic part; b200lda_loglik_parts separates them so the caller can sum document parts. */
int b200lda_loglik(b200lda_ctx* ctx, double* out);
int b200lda_loglik_parts(b200lda_ctx* ctx, double* doc_part, double* word_part);

/* Count conservation, checked on the device (north star: sum n_wk = sum n_dk = token count):
 * out[0] = sum_k n_k, out[1] = sum n_wk, out[2] = number of topics whose n_wk column sum differs
 * from n_k (must be 0), out[3] = sum of this shard's n_dk. With shards, out[0] and out[1] are
 * global token counts (every replica holds the global counts), out[3] the shard's own tokens. */
int b200lda_check_invariants(b200lda_ctx* ctx, int64_t* out /* 4 */);

/* model.data.get(d).topicSequence.getFeatures()            cmu_ron/TrainAndPredict.java:135-143 */
int b200lda_get_assignments(b200lda_ctx* ctx, int32_t* z /* num_tokens, document order */);
int b200lda_get_assignments_u16(b200lda_ctx* ctx, uint16_t* z /* num_tokens, document order */);
/* typeTopicCounts / tokensPerTopic (dense)                 used by getInferencer(), cmu_ron/…:169 */
int b200lda_get_nwk(b200lda_ctx* ctx, int32_t* nwk /* V*K, row-major by word */);
int b200lda_get_nk(b200lda_ctx* ctx, int32_t* nk /* K */);
/* Sparse doc-topic rows, topics ascending: row d = [row_ptr[d], row_ptr[d]+nnz). Pass topic ==
 * NULL to only size: row_ptr (num_docs+1 entries) is filled either way. */
int b200lda_get_ndk_csr(b200lda_ctx* ctx, int64_t* row_ptr, int32_t* topic, int32_t* count);
/* The word -> token CSR order of the loaded corpus (built on the device on first request; the
 * sampler itself counts n_wk straight from the doc order): word w's tokens are
 * word_tokens[word_ptr[w] .. word_ptr[w+1]) as indices into tok_word; order inside a word is
 * unspecified. word_ptr has V+1 entries; word_tokens (num_tokens entries) may be NULL. */
int b200lda_get_word_order(b200lda_ctx* ctx, int64_t* word_ptr, int64_t* word_tokens);
/* model.getTopicProbabilities(topicSequence)               cmu_ron/TrainAndPredict.java:143,
 * theta for documents [doc_begin, doc_end): (doc_end-doc_begin)*K doubles.    cmu/…:113 */
int b200lda_get_theta(b200lda_ctx* ctx, int64_t doc_begin, int64_t doc_end, double* theta);
/* phi_kw = (n_wk+beta)/(n_k+V beta): K*V doubles, row-major by topic (printTopWords input,
 * cmu_ron/TrainAndPredict.java:231). */
int b200lda_get_phi(b200lda_ctx* ctx, double* phi);

/* Hyper-parameters. */
int b200lda_set_alpha(b200lda_ctx* ctx, const double* alpha /* K */);
int b200lda_get_alpha(b200lda_ctx* ctx, double* alpha /* K */);
int b200lda_set_beta(b200lda_ctx* ctx, double beta);
int b200lda_get_beta(b200lda_ctx* ctx, double* beta);

/* Hyper-parameter optimisation: model.setOptimizeInterval(20)   cmu_ron/TrainAndPredict.java:163,
 * (ParallelTopicModel.optimizeAlpha / optimizeBeta inside estimate()).          cmu/…:261
 *   hyper_begin(width)  allocate + clear the histograms; width > longest document over ALL shards
 *   hyper_collect       add the current n_dk rows: docLengthCounts[n], topicDocCounts[k][n]
 *                       (Mallet's workers do this on iterations % saveSampleInterval == 0)
 *   hyper_buffer        device address + int32 count ((K+1)*width) to sum over shards first
 *   hyper_get           host copies of the histograms (either pointer may be NULL)
 *   optimize_alpha      Dirichlet.learnParameters(alpha, topicDocCounts, docLengthCounts)
 *                       (Gamma(1.00001, 1) prior, 200 rounds); installs alpha, clears histograms
 *   optimize_beta       Dirichlet.learnSymmetricConcentration over n_wk cell values and n_k;
 *                       installs beta = betaSum / V. */
int b200lda_hyper_begin(b200lda_ctx* ctx, int32_t width);
int b200lda_hyper_collect(b200lda_ctx* ctx);
int b200lda_hyper_buffer(b200lda_ctx* ctx, void** d_buf, int64_t* count);
int b200lda_hyper_get(b200lda_ctx* ctx, int32_t* topic_doc_counts /* K*width */, int32_t* doc_length_counts /* width */);
int b200lda_optimize_alpha(b200lda_ctx* ctx);
int b200lda_optimize_beta(b200lda_ctx* ctx);

/* Checkpoint / resume (replaces the Java serialisation of the trained model,
 * cmu_ron/TrainAndPredict.java:179-200, and its skip-training-if-the-file-exists path, :215-226):
 * the chain's whole state as one host blob = header (K, V, D, N, mode, sweep counter, Philox seed,
 * beta, shard rank/offset, checksum of the loaded doc_ptr / tok_word) + alpha[K] + z[N] (uint16).
 * set_state needs the SAME corpus loaded (checked by size and checksum); it installs alpha, beta,
 * the seed and the sweep counter and rebuilds every count from z, so a restored DEFERRED chain
 * continues bit-identically. Multi-shard models: one blob per shard, then the count sync. */
int b200lda_state_size(b200lda_ctx* ctx, int64_t* bytes);
int b200lda_get_state(b200lda_ctx* ctx, void* buf, int64_t bytes);
int b200lda_set_state(b200lda_ctx* ctx, const void* buf, int64_t bytes);
/* Sweep counter alone (continues the Philox stream of a chain restored through init_assignments). */
int b200lda_set_sweep_counter(b200lda_ctx* ctx, int64_t sweeps_done);

int b200lda_get_stats(b200lda_ctx* ctx, b200lda_stats* out);
int b200lda_reset_stats(b200lda_ctx* ctx);

/* Pinned host staging memory for callers whose runtime cannot pin its own (JVM heaps). */
int b200lda_host_alloc(void** out, size_t bytes);
int b200lda_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* B200LDA_H */
