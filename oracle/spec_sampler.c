/*
 * oracle/spec_sampler.c — TEST INFRASTRUCTURE (CPU oracle, part b). PARITY UNPINNED (see
 * lda_oracle.h). Never linked into or called from the product library.
 *
 * Sequential restatement of the sampling spec the sm_100a kernels implement (DESIGN.md
 * "sampling spec"): the collapsed conditional
 *      p(z=k) ∝ (n_wk^{-i} + β)(n_dk^{-i} + α_k) / (n_k + Vβ)
 * that Mallet's WorkerRunnable.sampleTopicsForOneDoc evaluates (reference call sites
 * cmu_ron/TrainAndPredict.java:166, cmu/TrainAndPredict.java:265), split in
 * a sparse doc bucket  n_dk (n_wk+β)/(n_k+Vβ)  and a per-word prior bucket  α_k (n_wk+β)/(n_k+Vβ),
 * all in fp32 with every operation individually rounded (compile with -ffp-contract=off) and all
 * prefix sums in the order a warp produces them: 32-lane Kogge-Stone tiles with a sequential carry
 * for the per-word prior rows, lane-strided sums + one Kogge-Stone scan for a document's row.
 * Same counts + same uniforms  =>  same topic index as the GPU, bit for bit.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lda_oracle.h"
#include "philox.h"

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr, key, out);
}

void oracle_init_z_philox(int64_t N, int32_t K, uint64_t seed, int64_t global_off, int32_t* z) {
  for (int64_t i = 0; i < N; ++i) {
    uint32_t r[4];
    b200lda_token_random(seed, (uint64_t)(global_off + i), 0u, 1u, r);
    z[i] = (int32_t)(((uint64_t)r[0] * (uint64_t)K) >> 32);
  }
}

void oracle_count(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr, const int32_t* tok_word,
                  const int32_t* z, int32_t* nwk, int32_t* nk) {
  memset(nwk, 0, sizeof(int32_t) * (size_t)V * (size_t)K);
  memset(nk, 0, sizeof(int32_t) * (size_t)K);
  int64_t N = doc_ptr[D];
  for (int64_t i = 0; i < N; ++i) {
    nwk[(size_t)tok_word[i] * K + z[i]]++;
    nk[z[i]]++;
  }
}

void oracle_ndk_csr(int64_t D, int32_t K, const int64_t* doc_ptr, const int32_t* z,
                    int64_t* row_ptr, int32_t* nnz, int32_t* topic, int32_t* count) {
  int32_t* dense = (int32_t*)calloc((size_t)K, sizeof(int32_t));
  int64_t off = 0;
  for (int64_t d = 0; d < D; ++d) {
    int64_t len = doc_ptr[d + 1] - doc_ptr[d];
    row_ptr[d] = off;
    for (int64_t i = doc_ptr[d]; i < doc_ptr[d + 1]; ++i) dense[z[i]]++;
    int32_t n = 0;
    for (int32_t k = 0; k < K; ++k)
      if (dense[k]) {
        topic[off + n] = k;
        count[off + n] = dense[k];
        dense[k] = 0;
        ++n;
      }
    nnz[d] = n;
    off += (len < K) ? len : K;
  }
  row_ptr[D] = off;
  free(dense);
}

void oracle_tile_scan_f32(const float* in, int64_t n, float* out) {
  float carry = 0.0f;
  for (int64_t base = 0; base < n; base += 32) {
    float x[32], y[32];
    for (int l = 0; l < 32; ++l) x[l] = (base + l < n) ? in[base + l] : 0.0f;
    for (int d = 1; d < 32; d <<= 1) {
      for (int l = 0; l < 32; ++l) y[l] = (l >= d) ? (x[l] + x[l - d]) : x[l];
      memcpy(x, y, sizeof(x));
    }
    for (int l = 0; l < 32 && base + l < n; ++l) out[base + l] = carry + x[l];
    carry = carry + x[31];
  }
}

/* Doc-bucket prefix order of the spec ("lane-strided") over a row laid out in nt tiles of 32
 * slots: slot (g, l) = tile g, lane l, sits at in[32 g + l]. Each lane sums ITS slots tile by tile
 * starting from +0 (lane-local inclusive sums), the 32 lane totals go through the Kogge-Stone
 * scan, and slot (g, l) gets  P = E_l + (lane-local inclusive sum), E_l = scanned total of lane
 * l-1 (0 for lane 0). The cumulative order is therefore lane-major: (lane 0: tiles 0, 1, ...),
 * (lane 1: ...), ... *total = the scanned total of lane 31. One warp scan per token whatever the
 * row width. local (may be NULL) receives the lane-local inclusive sums. */
static void lane_prefix_tiles(const float* in, int32_t nt, float* out, float* local, float* E_out,
                              float* total) {
  float T[32], y[32];
  for (int l = 0; l < 32; ++l) {
    float run = 0.0f;
    for (int32_t g = 0; g < nt; ++g) {
      run = run + in[32 * g + l];
      out[32 * g + l] = run; /* lane-local for now */
    }
    T[l] = run;
  }
  if (local) memcpy(local, out, sizeof(float) * 32u * (size_t)nt);
  for (int d = 1; d < 32; d <<= 1) {
    for (int l = 0; l < 32; ++l) y[l] = (l >= d) ? (T[l] + T[l - d]) : T[l];
    memcpy(T, y, sizeof(T));
  }
  for (int l = 0; l < 32; ++l) {
    const float E = l == 0 ? 0.0f : T[l - 1];
    if (E_out) E_out[l] = E;
    for (int32_t g = 0; g < nt; ++g) out[32 * g + l] = E + out[32 * g + l];
  }
  *total = T[31];
}

/* Exported building block (tests/test_oracle.py): n slots in positional layout (slot j on lane
 * j mod 32, tile j div 32), n/32 + 1 tiles, the slots past n weighing +0. */
void oracle_lane_strided_prefix_f32(const float* in, int64_t n, float* out, float* total) {
  const int32_t nt = (int32_t)(n / 32 + 1);
  float* a = (float*)calloc(32u * (size_t)nt, sizeof(float));
  float* S = (float*)malloc(sizeof(float) * 32u * (size_t)nt);
  memcpy(a, in, sizeof(float) * (size_t)n);
  lane_prefix_tiles(a, nt, S, NULL, NULL, total);
  memcpy(out, S, sizeof(float) * (size_t)n);
  free(a);
  free(S);
}

void oracle_spec_tables(int32_t V, int32_t K, const int32_t* nwk, const int32_t* nk,
                        const double* alpha, double beta, float* invden, float* ab, float* prior,
                        float* q) {
  const float beta_f = (float)beta;
  const float vbeta = (float)V * beta_f;
  for (int32_t k = 0; k < K; ++k) {
    invden[k] = 1.0f / ((float)nk[k] + vbeta);
    ab[k] = (float)alpha[k] * invden[k];
  }
  float* b = (float*)malloc(sizeof(float) * (size_t)K);
  for (int32_t w = 0; w < V; ++w) {
    const int32_t* row = nwk + (size_t)w * K;
    for (int32_t k = 0; k < K; ++k) b[k] = ((float)row[k] + beta_f) * ab[k];
    oracle_tile_scan_f32(b, K, prior + (size_t)w * K);
    q[w] = prior[(size_t)w * K + K - 1];
  }
  free(b);
}

/* Levels: L0 = row (size K); L(i+1)[m] = L(i)[min(32m+31, size_i-1)]; top level has <= 32
 * entries. Search picks, top-down, the first entry > s inside the current 32-block, or the
 * block's last entry when none is. */
int32_t oracle_spec_hsearch(const float* row, int32_t K, float s) {
  int32_t sizes[8];
  int nl = 0;
  sizes[nl++] = K;
  while (sizes[nl - 1] > 32) {
    sizes[nl] = (sizes[nl - 1] + 31) / 32;
    ++nl;
  }
  int32_t block = 0; /* index of the 32-block inside the current level */
  for (int lev = nl - 1; lev >= 0; --lev) {
    int32_t lo = block * 32;
    int32_t hi = lo + 32;
    if (hi > sizes[lev]) hi = sizes[lev];
    int32_t pick = hi - 1;
    for (int32_t m = lo; m < hi; ++m) {
      /* value of level-`lev` entry m = fine entry min((m+1)*32^lev - 1, K-1) */
      int64_t fine = (int64_t)(m + 1);
      for (int t = 0; t < lev; ++t) fine *= 32;
      fine -= 1;
      if (fine > K - 1) fine = K - 1;
      if (row[fine] > s) {
        pick = m;
        break;
      }
    }
    block = pick;
  }
  return block;
}

/* ---- a document's row during one visit (DESIGN.md "sampling spec", v2) ---------------------
 * The GPU keeps the row in REGISTERS: nt tiles of 32 slots, slot (g, l) on lane l. A slot is
 * live (count > 0) or dead (count 0: weighs +0). Slots never shift:
 *   - visit start: the ascending compact list of the document's n non-zero topics is split evenly
 *     over nt = n/32 + 1 tiles (q = n / nt, r = n % nt: tile g takes q + [g < r] consecutive sorted
 *     positions starting at g q + min(g, r), on lanes 0, 1, ...); bound[g] = the first topic of
 *     tile g (g >= 1; K when the tile starts empty);
 *   - every live slot caches  wt = invden[t] * (float)n_dk  (recomputed, with that product,
 *     whenever its count changes); the token being resampled weighs  wt - invden[o]  at its slot;
 *   - a topic leaving the document kills its slot in place; a topic entering it takes the lowest
 *     dead lane of its preferred tile g* = #{g >= 1 : topic >= bound[g]} or, when that tile is
 *     full, the lowest lane that has a dead slot in ANY tile, at that lane's lowest dead tile (the
 *     just-vacated slot included); when no slot is dead, an empty tile is appended (bound = K) and
 *     its lane 0 is taken. (Every lane keeps a bit mask of its dead slots: the search is one or
 *     two ballots.)
 * Tiles therefore stay (mostly) topic ranges, which is what keeps a tile's 32 n_wk gathers on few
 * 128-byte lines; nothing else depends on the placement. At the end of the visit the live slots
 * go back to the document's packed row in ascending topic order. */
typedef struct {
  int32_t nt, cap_tiles;
  int32_t* topic; /* [32 cap_tiles] */
  int32_t* count;
  int32_t* bound; /* [cap_tiles] */
  float *wt, *a, *S, *loc;
} spec_row;

static void row_alloc(spec_row* r, int64_t max_slots) {
  r->cap_tiles = (int32_t)(max_slots / 32 + 2);
  const size_t n = 32u * (size_t)r->cap_tiles;
  r->topic = (int32_t*)calloc(n, sizeof(int32_t));
  r->count = (int32_t*)calloc(n, sizeof(int32_t));
  r->bound = (int32_t*)calloc((size_t)r->cap_tiles, sizeof(int32_t));
  r->wt = (float*)calloc(n, sizeof(float));
  r->a = (float*)calloc(n, sizeof(float));
  r->S = (float*)calloc(n, sizeof(float));
  r->loc = (float*)calloc(n, sizeof(float));
  r->nt = 1;
}
static void row_free(spec_row* r) {
  free(r->topic);
  free(r->count);
  free(r->bound);
  free(r->wt);
  free(r->a);
  free(r->S);
  free(r->loc);
}

/* visit start: ns sorted (topic, count) pairs */
static void row_init(spec_row* r, int32_t K, const float* invden, const int32_t* st, const int32_t* sc,
                     int32_t ns) {
  const int32_t nt = ns / 32 + 1;
  const int32_t q = ns / nt, rem = ns % nt;
  r->nt = nt;
  memset(r->topic, 0, sizeof(int32_t) * 32u * (size_t)nt);
  memset(r->count, 0, sizeof(int32_t) * 32u * (size_t)nt);
  memset(r->wt, 0, sizeof(float) * 32u * (size_t)nt);
  for (int32_t g = 0; g < nt; ++g) {
    const int32_t b = g * q + (g < rem ? g : rem), e = b + q + (g < rem ? 1 : 0);
    for (int32_t j = b; j < e; ++j) {
      r->topic[32 * g + (j - b)] = st[j];
      r->count[32 * g + (j - b)] = sc[j];
      r->wt[32 * g + (j - b)] = invden[st[j]] * (float)sc[j];
    }
    r->bound[g] = (g == 0) ? 0 : (b < e ? st[b] : K);
  }
}

/* One token: returns the new topic. Doc bucket weights, own token excluded from n_wk and n_dk:
 *   a = ((float)n_wk' + beta) * wt'   with wt' = wt - invden[o] at the old topic's slot */
static int32_t row_select(const spec_row* r, int32_t K, const int32_t* nwk_row, const float* invden,
                          const float* ab, const float* prior_row, float q_w, float beta_f,
                          int32_t old_topic, float u) {
  const int32_t nt = r->nt, n = 32 * nt;
  for (int32_t j = 0; j < n; ++j) {
    if (r->count[j] == 0) {
      r->a[j] = 0.0f;
      continue;
    }
    const int32_t t = r->topic[j];
    int32_t m = nwk_row[t];
    float wgt = r->wt[j];
    if (t == old_topic) {
      m -= 1;
      wgt = wgt - invden[old_topic];
    }
    if (m < 0) m = 0;
    const float x = (float)m + beta_f;
    r->a[j] = x * wgt;
  }
  float A = 0.0f, E[32];
  lane_prefix_tiles(r->a, nt, r->S, r->loc, E, &A);
  const float delta = ab[old_topic];
  float qp = q_w - delta;
  if (qp < 0.0f) qp = 0.0f;
  const float T = A + qp;
  const float x = u * T;
  if (x < A) {
    /* first slot in cumulative (lane-major) order whose prefix exceeds x. The scanned lane offsets
     * and the lane-local sums round differently, so within an ulp of a lane boundary that slot can
     * be one that adds no weight (a dead slot, or the old topic's slot when the token is its only
     * one), or there is none at all (x within an ulp of A): the token then keeps its topic. */
    int32_t pick = -1;
    for (int l = 0; l < 32 && pick < 0; ++l)
      for (int32_t g = 0; g < nt; ++g)
        if (r->S[32 * g + l] > x) {
          pick = 32 * g + l;
          break;
        }
    if (pick < 0 || r->count[pick] == 0) return old_topic;
    if (r->topic[pick] == old_topic && r->count[pick] == 1) return old_topic;
    return r->topic[pick];
  }
  const float y = x - A;
  const float po = prior_row[old_topic];
  const float pod = po - delta;
  const float s = (y < pod) ? y : (y + delta);
  return oracle_spec_hsearch(prior_row, K, s);
}

/* the token moves from topic o to topic n != o */
static void row_move(spec_row* r, int32_t K, const float* invden, int32_t o, int32_t n) {
  const int32_t slots = 32 * r->nt;
  int32_t jo = 0;
  while (!(r->count[jo] > 0 && r->topic[jo] == o)) ++jo;
  r->count[jo] -= 1;
  r->wt[jo] = invden[o] * (float)r->count[jo];
  for (int32_t j = 0; j < slots; ++j)
    if (r->count[j] > 0 && r->topic[j] == n) {
      r->count[j] += 1;
      r->wt[j] = invden[n] * (float)r->count[j];
      return;
    }
  int32_t gstar = 0;
  for (int32_t g = 1; g < r->nt; ++g)
    if (n >= r->bound[g]) ++gstar;
  for (int l = 0; l < 32; ++l) /* lowest dead lane of the preferred tile */
    if (r->count[32 * gstar + l] == 0) {
      r->topic[32 * gstar + l] = n;
      r->count[32 * gstar + l] = 1;
      r->wt[32 * gstar + l] = invden[n];
      return;
    }
  for (int l = 0; l < 32; ++l) /* lowest lane with a dead slot anywhere, its lowest dead tile */
    for (int32_t g = 0; g < r->nt; ++g)
      if (r->count[32 * g + l] == 0) {
        r->topic[32 * g + l] = n;
        r->count[32 * g + l] = 1;
        r->wt[32 * g + l] = invden[n];
        return;
      }
  /* every slot is live: append an empty tile, take its lane 0 */
  if (r->nt >= r->cap_tiles) abort();
  memset(r->topic + 32 * r->nt, 0, sizeof(int32_t) * 32);
  memset(r->count + 32 * r->nt, 0, sizeof(int32_t) * 32);
  memset(r->wt + 32 * r->nt, 0, sizeof(float) * 32);
  r->bound[r->nt] = K;
  r->topic[32 * r->nt] = n;
  r->count[32 * r->nt] = 1;
  r->wt[32 * r->nt] = invden[n];
  r->nt += 1;
}

/* One token against a row given as its visit-start list (slots = the document's sorted non-zero
 * topics INCLUDING the current token). Returns the new topic. */
int32_t oracle_spec_select(int32_t K, const int32_t* slot_topic, const int32_t* slot_count,
                           int32_t nslots, const int32_t* nwk_row, const float* invden,
                           const float* ab, const float* prior_row, float q_w, float beta_f,
                           int32_t old_topic, float u) {
  spec_row r;
  row_alloc(&r, nslots);
  row_init(&r, K, invden, slot_topic, slot_count, nslots);
  const int32_t res = row_select(&r, K, nwk_row, invden, ab, prior_row, q_w, beta_f, old_topic, u);
  row_free(&r);
  return res;
}

/* ---- drivers ---------------------------------------------------------------------------- */

typedef struct {
  float *invden, *ab, *prior, *q;
} spec_tables;

static spec_tables tables_alloc(int32_t V, int32_t K) {
  spec_tables t;
  t.invden = (float*)malloc(sizeof(float) * (size_t)K);
  t.ab = (float*)malloc(sizeof(float) * (size_t)K);
  t.prior = (float*)malloc(sizeof(float) * (size_t)V * (size_t)K);
  t.q = (float*)malloc(sizeof(float) * (size_t)V);
  return t;
}
static void tables_free(spec_tables* t) {
  free(t->invden);
  free(t->ab);
  free(t->prior);
  free(t->q);
}

static float token_uniform(uint64_t seed, int64_t g, uint32_t sweep) {
  uint32_t r[4];
  b200lda_token_random(seed, (uint64_t)g, sweep, 0u, r);
  return b200lda_u24(r[0]);
}

void oracle_spec_frozen(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr,
                        const int32_t* tok_word, const int32_t* z_in, const double* alpha,
                        double beta, uint64_t seed, uint32_t sweep, int64_t global_off,
                        const float* uniforms, int32_t* z_out) {
  int32_t* nwk = (int32_t*)malloc(sizeof(int32_t) * (size_t)V * (size_t)K);
  int32_t* nk = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
  oracle_count(D, V, K, doc_ptr, tok_word, z_in, nwk, nk);
  spec_tables t = tables_alloc(V, K);
  oracle_spec_tables(V, K, nwk, nk, alpha, beta, t.invden, t.ab, t.prior, t.q);
  const float beta_f = (float)beta;
  int32_t* dense = (int32_t*)calloc((size_t)K, sizeof(int32_t));
  int32_t* st = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
  int32_t* sc = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
  spec_row r;
  row_alloc(&r, K);
  for (int64_t d = 0; d < D; ++d) {
    for (int64_t i = doc_ptr[d]; i < doc_ptr[d + 1]; ++i) dense[z_in[i]]++;
    int32_t ns = 0;
    for (int32_t k = 0; k < K; ++k)
      if (dense[k]) {
        st[ns] = k;
        sc[ns] = dense[k];
        dense[k] = 0;
        ++ns;
      }
    row_init(&r, K, t.invden, st, sc, ns);
    for (int64_t i = doc_ptr[d]; i < doc_ptr[d + 1]; ++i) {
      int32_t w = tok_word[i];
      float u = uniforms ? uniforms[i] : token_uniform(seed, global_off + i, sweep);
      z_out[i] = row_select(&r, K, nwk + (size_t)w * K, t.invden, t.ab, t.prior + (size_t)w * K, t.q[w],
                            beta_f, z_in[i], u);
    }
  }
  row_free(&r);
  free(dense);
  free(st);
  free(sc);
  tables_free(&t);
  free(nwk);
  free(nk);
}

void oracle_spec_sweeps(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr,
                        const int32_t* tok_word, int32_t* z, const double* alpha, double beta,
                        uint64_t seed, uint32_t first_sweep, int32_t n_sweeps, int64_t global_off) {
  oracle_spec_sweeps_mode(D, V, K, doc_ptr, tok_word, z, alpha, beta, seed, first_sweep, n_sweeps,
                          global_off, 0);
}

/* One sweep over a set of documents against GIVEN counts (the shard's replica of the global
 * n_wk / n_k at sweep start). z moves in place; delta_nwk / delta_nk (may be NULL) receive
 * counts_after - counts_before for these documents: what the shard contributes to the all-reduce.
 * live != 0 additionally applies every move to nwk immediately (nwk must then be writable).
 * exclude_self == 0 is held-out inference: the documents' tokens are not part of the counts. */
void oracle_spec_sweep_given_counts(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr,
                                    const int32_t* tok_word, int32_t* z, int32_t* nwk, const int32_t* nk,
                                    const double* alpha, double beta, uint64_t seed, uint32_t sweep,
                                    int64_t global_off, int32_t live, int32_t exclude_self,
                                    int32_t* delta_nwk, int32_t* delta_nk) {
  spec_tables t = tables_alloc(V, K);
  const float beta_f = (float)beta;
  int32_t* dense = (int32_t*)calloc((size_t)K, sizeof(int32_t));
  int32_t* st = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
  int32_t* sc = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
  int32_t* shifted = NULL; /* inference: present n_wk + 1 at the old topic so the spec's -1 cancels */
  oracle_spec_tables(V, K, nwk, nk, alpha, beta, t.invden, t.ab, t.prior, t.q);
  if (!exclude_self) {
    for (int32_t k = 0; k < K; ++k) t.ab[k] = 0.0f; /* delta = 0: nothing of ours is in the table */
    shifted = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
  }
  if (delta_nwk) memset(delta_nwk, 0, sizeof(int32_t) * (size_t)V * (size_t)K);
  if (delta_nk) memset(delta_nk, 0, sizeof(int32_t) * (size_t)K);
  spec_row r;
  row_alloc(&r, K);
  for (int64_t d = 0; d < D; ++d) {
    for (int64_t i = doc_ptr[d]; i < doc_ptr[d + 1]; ++i) dense[z[i]]++;
    int32_t ns = 0;
    for (int32_t k = 0; k < K; ++k)
      if (dense[k]) {
        st[ns] = k;
        sc[ns] = dense[k];
        dense[k] = 0;
        ++ns;
      }
    row_init(&r, K, t.invden, st, sc, ns);
    for (int64_t i = doc_ptr[d]; i < doc_ptr[d + 1]; ++i) {
      const int32_t w = tok_word[i];
      const int32_t o = z[i];
      const float u = token_uniform(seed, global_off + i, sweep);
      const int32_t* row = nwk + (size_t)w * K;
      if (!exclude_self) {
        memcpy(shifted, row, sizeof(int32_t) * (size_t)K);
        shifted[o] += 1;
        row = shifted;
      }
      const int32_t n = row_select(&r, K, row, t.invden, t.ab, t.prior + (size_t)w * K, t.q[w], beta_f, o, u);
      if (n != o) {
        row_move(&r, K, t.invden, o, n);
        z[i] = n;
        if (live) {
          nwk[(size_t)w * K + o]--;
          nwk[(size_t)w * K + n]++;
        }
        if (delta_nwk) {
          delta_nwk[(size_t)w * K + o]--;
          delta_nwk[(size_t)w * K + n]++;
        }
        if (delta_nk) {
          delta_nk[o]--;
          delta_nk[n]++;
        }
      }
    }
  }
  row_free(&r);
  free(shifted);
  free(dense);
  free(st);
  free(sc);
  tables_free(&t);
}

/* live != 0: n_wk moves immediately (a sequential rendering of the GPU's LIVE mode: the prior
 * table and n_k still date from the sweep start). live == 0: DEFERRED mode. */
void oracle_spec_sweeps_mode(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr,
                             const int32_t* tok_word, int32_t* z, const double* alpha, double beta,
                             uint64_t seed, uint32_t first_sweep, int32_t n_sweeps,
                             int64_t global_off, int32_t live) {
  int32_t* nwk = (int32_t*)malloc(sizeof(int32_t) * (size_t)V * (size_t)K);
  int32_t* nk = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
  for (int32_t it = 0; it < n_sweeps; ++it) {
    /* frozen per-sweep snapshot: recounting from z == applying the summed deltas */
    oracle_count(D, V, K, doc_ptr, tok_word, z, nwk, nk);
    oracle_spec_sweep_given_counts(D, V, K, doc_ptr, tok_word, z, nwk, nk, alpha, beta, seed,
                                   first_sweep + (uint32_t)it, global_off, live, 1, NULL, NULL);
  }
  free(nwk);
  free(nk);
}

/* TopicInferencer.getSampledDistribution under the spec (reference call
 * cmu_ron/TrainAndPredict.java:144): frozen trained counts, uniform Philox init, theta = mean of
 * (alpha_k + n_dk) over the kept samples, normalised. theta: D*K doubles. */
void oracle_spec_infer(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr, const int32_t* tok_word,
                       const int32_t* nwk, const int32_t* nk, const double* alpha, double beta,
                       int32_t iterations, int32_t thinning, int32_t burn_in, uint64_t seed, double* theta) {
  const int64_t N = doc_ptr[D];
  int32_t* z = (int32_t*)malloc(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
  int64_t* acc = (int64_t*)calloc((size_t)D * (size_t)K, sizeof(int64_t));
  oracle_init_z_philox(N, K, seed ^ 0x9E3779B97F4A7C15ULL, 0, z);
  int samples = 0;
  for (int32_t it = 1; it <= iterations; ++it) {
    oracle_spec_sweep_given_counts(D, V, K, doc_ptr, tok_word, z, (int32_t*)nwk, nk, alpha, beta, seed,
                                   (uint32_t)it, 0, 0, 0, NULL, NULL);
    if (it > burn_in && (it - burn_in) % thinning == 0) {
      for (int64_t d = 0; d < D; ++d)
        for (int64_t i = doc_ptr[d]; i < doc_ptr[d + 1]; ++i) acc[(size_t)d * K + z[i]]++;
      ++samples;
    }
  }
  if (samples == 0) {
    for (int64_t d = 0; d < D; ++d)
      for (int64_t i = doc_ptr[d]; i < doc_ptr[d + 1]; ++i) acc[(size_t)d * K + z[i]]++;
    samples = 1;
  }
  double alpha_sum = 0.0;
  for (int32_t k = 0; k < K; ++k) alpha_sum += alpha[k];
  for (int64_t d = 0; d < D; ++d) {
    const double len = (double)(doc_ptr[d + 1] - doc_ptr[d]);
    for (int32_t k = 0; k < K; ++k)
      theta[(size_t)d * K + k] = ((double)samples * alpha[k] + (double)acc[(size_t)d * K + k]) /
                                 ((double)samples * (alpha_sum + len));
  }
  free(z);
  free(acc);
}

void oracle_exact_conditional(int32_t K, int32_t V, const int32_t* ndk_dense,
                              const int32_t* nwk_row, const int32_t* nk, const double* alpha,
                              double beta, int32_t old_topic, double* p) {
  double tot = 0.0;
  for (int32_t k = 0; k < K; ++k) {
    int32_t ex = (k == old_topic) ? 1 : 0;
    double v = ((double)(nwk_row[k] - ex) + beta) * ((double)(ndk_dense[k] - ex) + alpha[k]) /
               ((double)nk[k] + (double)V * beta);
    p[k] = v;
    tot += v;
  }
  for (int32_t k = 0; k < K; ++k) p[k] /= tot;
}

/* ---- log-likelihood, theta, phi ----------------------------------------------------------- */

double oracle_log_gamma_stirling(double z) {
  int shift = 0;
  while (z < 2.0) {
    z += 1.0;
    shift++;
  }
  const double half_log_2pi = 0.91893853320467274178;
  double result = half_log_2pi + (z - 0.5) * log(z) - z + 1.0 / (12.0 * z) -
                  1.0 / (360.0 * z * z * z) + 1.0 / (1260.0 * z * z * z * z * z);
  while (shift > 0) {
    shift--;
    z -= 1.0;
    result -= log(z);
  }
  return result;
}

double oracle_loglik(int64_t D, int32_t V, int32_t K, const int64_t* doc_ptr,
                     const int32_t* tok_word, const int32_t* z, const double* alpha, double beta,
                     int32_t stirling) {
  double (*lg)(double) = stirling ? oracle_log_gamma_stirling : lgamma;
  double alpha_sum = 0.0;
  for (int32_t k = 0; k < K; ++k) alpha_sum += alpha[k];
  double* lg_alpha = (double*)malloc(sizeof(double) * (size_t)K);
  for (int32_t k = 0; k < K; ++k) lg_alpha[k] = lg(alpha[k]);
  int32_t* dense = (int32_t*)calloc((size_t)K, sizeof(int32_t));
  double ll = 0.0;
  for (int64_t d = 0; d < D; ++d) {
    int64_t len = doc_ptr[d + 1] - doc_ptr[d];
    for (int64_t i = doc_ptr[d]; i < doc_ptr[d + 1]; ++i) dense[z[i]]++;
    for (int32_t k = 0; k < K; ++k)
      if (dense[k] > 0) {
        ll += lg(alpha[k] + (double)dense[k]) - lg_alpha[k];
        dense[k] = 0;
      }
    ll -= lg(alpha_sum + (double)len);
  }
  ll += (double)D * lg(alpha_sum);
  int32_t* nwk = (int32_t*)malloc(sizeof(int32_t) * (size_t)V * (size_t)K);
  int32_t* nk = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
  oracle_count(D, V, K, doc_ptr, tok_word, z, nwk, nk);
  int64_t nonzero = 0;
  for (size_t i = 0; i < (size_t)V * (size_t)K; ++i)
    if (nwk[i] > 0) {
      ++nonzero;
      ll += lg(beta + (double)nwk[i]);
    }
  for (int32_t k = 0; k < K; ++k) ll -= lg(beta * (double)V + (double)nk[k]);
  ll += (double)K * lg(beta * (double)V);
  ll -= (double)nonzero * lg(beta);
  free(nwk);
  free(nk);
  free(dense);
  free(lg_alpha);
  return ll;
}

void oracle_theta(int32_t K, const int32_t* z_doc, int64_t len, const double* alpha, double* out) {
  double alpha_sum = 0.0;
  for (int32_t k = 0; k < K; ++k) {
    alpha_sum += alpha[k];
    out[k] = 0.0;
  }
  for (int64_t i = 0; i < len; ++i) out[z_doc[i]] += 1.0;
  for (int32_t k = 0; k < K; ++k) out[k] = (out[k] + alpha[k]) / ((double)len + alpha_sum);
}

void oracle_phi(int32_t V, int32_t K, const int32_t* nwk, const int32_t* nk, double beta,
                double* out) {
  for (int32_t k = 0; k < K; ++k)
    for (int32_t w = 0; w < V; ++w)
      out[(size_t)k * V + w] =
          ((double)nwk[(size_t)w * K + k] + beta) / ((double)nk[k] + (double)V * beta);
}
