// sweep_kernel.cuh — K3/K7: the per-token resampling kernel (one warp per document).
//
// Replaces cc.mallet.topics.WorkerRunnable.sampleTopicsForOneDoc, reached through
// model.estimate() at reference cmu_ron/TrainAndPredict.java:166 and cmu/TrainAndPredict.java:265
// (SURVEY.md §8 rows a4, a5). It is NOT a translation of SparseLDA's s/r/q walk over packed rows:
//   * a warp owns a document for one visit; the document's sparse topic row lives in REGISTERS for
//     the whole visit: nt tiles of 32 slots, lane l holds slot (g, l) of every tile g as
//     (topic << 16 | count) plus the cached weight  wt = invden[topic] * count.  Slots never shift:
//     a topic that leaves the document kills its slot in place, a topic that enters it takes a dead
//     slot of its preferred tile (tiles are topic ranges fixed at visit start, so a tile's 32 n_wk
//     gathers stay on few 128-byte lines). A token step therefore edits at most two slots, each by
//     the lane that owns it: no shared-memory row, no shuffles, no shifting;
//   * doc bucket: every lane gathers n_wk[w, topic] for its slots (4 B each), forms
//     (n_wk + beta) * wt, sums its own slots and ONE warp scan runs over the 32 lane totals
//     (lane-strided prefix); the bucket search is a per-lane count plus one ballot;
//   * prior bucket: alpha_k (n_wk + beta) / (n_k + V beta) is word-only, so its mass and prefix
//     table are built once per sweep per word (table_kernels.cuh); a draw that lands there is
//     resolved by a fan-out-32 search = one coalesced 128-byte line per level;
//   * the 32 tokens of a batch are fetched one per lane (word, old topic, Philox draw, Q_w, P_w[o],
//     invden[o], ab[o]) and staged in shared memory, so a token step starts with two broadcast
//     128-bit shared loads instead of a shuffle per field;
//   * count moves: integer RED atomics carry the n_wk moves, n_k moves are accumulated per CTA;
//   * visit end: the live slots go back to the packed row in ascending topic order through a
//     K-bit bitmap in shared memory (rank = popcount of the lower bits).
// Every lane keeps a bit mask of its dead slots, so placing a topic that is new to the document is
// one ballot (two when its preferred tile is full).
// token_step<NT> is straight-line code specialised on the tile count (NT = 1..8); rows that can
// exceed 8 tiles (documents with more than 255 tokens when K > 255) run token_step_wide, the same
// algorithm with the row in shared memory and loops over tiles.
// MODE_UPDATE serves LIVE (LIVE=true: n_wk read through L1/L2 and written in place) and DEFERRED
// (LIVE=false: frozen n_wk through the read-only path, moves go to a second buffer);
// MODE_FROZEN moves nothing (north-star parity mode). MODE_INFER is held-out inference
// (TopicInferencer.getSampledDistribution): n_wk / n_k are frozen and do not contain the
// document, so documents are independent chains and a warp runs ALL iterations of its document in
// one kernel visit (one row visit per iteration), adding the row to the document's sample
// accumulator at every saved iteration.
#pragma once
#include "device_common.cuh"
#include "table_kernels.cuh"

namespace b200lda {

enum { MODE_UPDATE = 0, MODE_FROZEN = 1, MODE_INFER = 2 };

struct SweepParams {
  const int32_t* doc_order;   // [D] document ids, longest first
  int64_t order_begin, order_end;  // this launch's slice of doc_order
  const int64_t* doc_ptr;     // [D+1] token offsets (document order)
  const int32_t* tok_word;    // [N]
  uint16_t* z;                // [N] current topics (updated in MODE_UPDATE)
  int32_t* z_out;             // [N] MODE_FROZEN output
  const int64_t* row_ptr;     // [D+1] offsets into rows (capacity min(K, L_d) per doc)
  int32_t* row_nnz;           // [D]
  uint32_t* rows;             // packed (topic << 16 | count), ascending topic
  const int32_t* nwk_read;    // [V*K]
  int32_t* nwk_write;         // [V*K] (== nwk_read when LIVE; nullptr: no count writes)
  int32_t* nk_delta;          // [K]
  const float* invden;        // [K]  1 / (n_k + V beta)
  const float* ab;            // [K]  alpha_k * invden_k
  const float* prior;         // [V * layout.stride]
  const float* q;             // [V]  prior bucket mass per word
  const float* uniforms;      // [N] or nullptr (Philox)
  // LIVE mode: two copies of every word's prior row and Q_w (copy c of word w at row 2 w + c);
  // prior_sel[w] = the current copy. Beside a bulk launch runs k_prior_refresher: a few small CTAs
  // that rebuild refresh_rows rows of the hot words from the live counts, paced by the document
  // scheduler's progress.
  int32_t* prior_sel;         // [V] or nullptr (one copy)
  unsigned* row_cursor;       // rows claimed so far in this launch
  const int32_t* hot_words;   // [hot_count] the words that carry ~90 % of the tokens: the rows worth rebuilding
  int hot_count;
  unsigned refresh_rows;
  unsigned long long* refresh_count;  // statistics: rows rebuilt
  int sampler_warps;          // warps of the sampling launch the refreshers run beside
  int V;
  PriorLayout layout;
  int K;
  int cap_tiles;              // wide class: tiles of the shared-memory row (0 in the register classes)
  int doc_chunk;              // documents fetched per scheduler atomic (strided through doc_order)
  float beta_f;
  uint64_t seed;
  uint32_t sweep;
  int64_t global_tok_off;
  unsigned long long* doc_counter;  // dynamic document scheduler: next chunk index (starts at 0 for each launch)
  unsigned long long chunk_begin, chunk_end;  // this launch's range of the class's scheduler chunks (a segment of the sweep)
  unsigned long long* stats;        // [0] moved, [1] prior-bucket draws, [2] sum of nnz over tokens (this pass)
  // MODE_INFER: iterations 1..infer_iters per document (Philox sweep key = iteration); a sample is
  // saved when it > burn_in and (it - burn_in) % thinning == 0, or after the last iteration when
  // no iteration qualifies (infer_samples == 0): acc[d, k] += n_dk.
  int infer_iters, infer_burn_in, infer_thinning, infer_samples;
  int32_t* infer_acc;               // [D * K]
};

#ifndef B200LDA_SWEEP_GROUP
#define B200LDA_SWEEP_GROUP 4  // wide path: tiles whose gathers are in flight together
#endif
constexpr int kGroup = B200LDA_SWEEP_GROUP;

// ROWCLASS selects the token-step variants a kernel instance carries, so that the register
// allocation (one per kernel) of the narrow classes is not dictated by the 8-tile variant:
//   0: documents of <= 95 tokens (NT <= 3)   1: <= 159 (NT <= 5)   2: <= 255 (NT <= 8)
//   3: longer documents: row in shared memory, loops over tiles (token_step_wide)
// A row of n live slots starts a visit with n/32 + 1 tiles and grows a tile only when all of its
// slots are live, so a document of L tokens (at most min(K, L) topics) never needs more than
// min(K, L)/32 + 1 tiles.
constexpr int kRowClasses = 4;
constexpr int kWideClass = 3;
__host__ __device__ constexpr int rowclass_max_tiles(int rc) { return rc == 0 ? 3 : rc == 1 ? 5 : rc == 2 ? 8 : 0; }
__host__ __device__ constexpr int rowclass_max_len(int rc) { return 32 * rowclass_max_tiles(rc) - 1; }

#ifndef B200LDA_TOP_EARLY_NT
#define B200LDA_TOP_EARLY_NT 3  // rows of up to this many tiles request the prior's top level at the start of the token step
#endif
#ifndef B200LDA_RC0_MIN_CTAS
#define B200LDA_RC0_MIN_CTAS 6  // 40 registers, no spills: 48 warps per SM (measured: C4 29.9 -> 28.5 ms)
#endif
#ifndef B200LDA_RC1_MIN_CTAS
#define B200LDA_RC1_MIN_CTAS 5
#endif
#ifndef B200LDA_RC2_MIN_CTAS
#define B200LDA_RC2_MIN_CTAS 4
#endif
#ifndef B200LDA_RC3_MIN_CTAS
#define B200LDA_RC3_MIN_CTAS 4
#endif
__host__ __device__ constexpr int rowclass_min_ctas(int rc) {
  return rc == 0 ? B200LDA_RC0_MIN_CTAS : rc == 1 ? B200LDA_RC1_MIN_CTAS : rc == 2 ? B200LDA_RC2_MIN_CTAS : B200LDA_RC3_MIN_CTAS;
}

// Shared memory (32-bit words). Per CTA: [invden | ab | n_k delta] (3K, rounded to 4, when
// TABLES_IN_SMEM). Per warp: the token batch (32 tokens x 8 words), the write-back bitmap and its
// word prefix (2 x ceil(K/32)), and in the wide class the row: slots and lane-local prefixes
// (32 cap_tiles each) and the tile bounds (cap_tiles); a wide row's weights invden[t] * n_dk are
// recomputed per use, which keeps 8 warps x 16 tiles + the tables at 4 CTAs per SM.
constexpr int kBatchWords = 256;
__host__ __device__ constexpr int bitmap_words(int K) { return (K + 31) >> 5; }
__host__ __device__ constexpr int sweep_table_words(int K) { return (3 * K + 3) & ~3; }
__host__ __device__ constexpr int sweep_warp_words(int K, int cap_tiles) {
  return (kBatchWords + 2 * bitmap_words(K) + 65 * cap_tiles + 3) & ~3;
}

// The kernel's dynamic shared memory, addressed by WORD OFFSET everywhere: indexing the extern
// array keeps every access a plain LDS/STS/ATOMS with an immediate base, whereas pointers carried in
// a struct decay to generic addresses and cost an address-space conversion per access.
extern __shared__ __align__(16) unsigned char b200lda_smem_raw[];
__device__ __forceinline__ uint32_t& smem_u32(int word) { return reinterpret_cast<uint32_t*>(b200lda_smem_raw)[word]; }
__device__ __forceinline__ float& smem_f32(int word) { return reinterpret_cast<float*>(b200lda_smem_raw)[word]; }
__device__ __forceinline__ int* smem_i32_ptr(int word) { return reinterpret_cast<int*>(b200lda_smem_raw) + word; }
__device__ __forceinline__ uint32_t* smem_u32_ptr(int word) { return reinterpret_cast<uint32_t*>(b200lda_smem_raw) + word; }
__device__ __forceinline__ uint4& smem_u128(int word) { return reinterpret_cast<uint4*>(b200lda_smem_raw)[word >> 2]; }

// Shared-memory accesses of the token step go through the 32-bit shared-window address kept in a
// register (WarpCtx::sbase / batch_addr): left to itself the compiler rematerialises the window base
// (S2R CgaCtaId, ULEA) and the per-warp offset arithmetic at every use (~20 instructions per token
// in the first register-path build, profiles/r02_tuning.md).
__device__ __forceinline__ uint32_t smem_window_addr() {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(b200lda_smem_raw);
  asm volatile("" : "+r"(a));  // opaque: keep it in a register
  return a;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void reds_add_i32(uint32_t addr, int v) {
  asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

#ifndef B200LDA_LIVE_L1
#define B200LDA_LIVE_L1 1
#endif
// LIVE-mode read of an n_wk cell that other warps update with RED atomics. Through L1 (ld.global.ca):
// the issuing SM's own atomics invalidate its L1 line, so a warp always sees its own moves (pinned by
// tests/test_gpu_parity.py::test_live_mode_equals_sequential_oracle_when_documents_do_not_interact);
// moves made on other SMs become visible when the line is refetched, and L1 is invalidated at every
// launch, so nothing is older than the sweep. LIVE is racy across documents by design; what the L1
// path buys is the hot words' rows at small K (C2, K = 100: 40.1 -> 24.8 ms per sweep).
// B200LDA_LIVE_L1=0 reads at L2 (ld.global.cg), the point of coherence of the atomics.
template <bool LIVE>
__device__ __forceinline__ int count_load(const int32_t* cell) {
  if (!LIVE) return __ldg(cell);
#if B200LDA_LIVE_L1
  return __ldca(cell);
#else
  return __ldcg(cell);
#endif
}

// Per-warp constants and counters shared by the token-step variants.
struct WarpCtx {
  int lane;
  float beta_f;
  int K;
  // word offsets into the kernel's shared memory
  int tab;     // [invden | ab | n_k delta] (TABLES_IN_SMEM): invden at tab, ab at tab + K, deltas at tab + 2K
  int batch;   // this warp's token batch
  int bm;      // write-back bitmap, pf = bm + bmw: exclusive popcount prefix per bitmap word
  int bmw;
  int row;     // wide class: slots at row, prefixes at row + 32 capT, bounds at row + 64 capT
  int capT;
  bool nkd_in_smem;
  int top_lane;  // offset of this lane's entry of the top search level inside a word's prior block, -1: none
  unsigned st_moved, st_prior;
  uint32_t sbase;       // shared-window address of word 0 of the kernel's shared memory
  uint32_t batch_addr;  // shared-window address of this warp's token batch
  uint32_t row_bytes;   // 4 K: bytes of one n_wk row
};

// Prior bucket: skip the own-token mass delta at topic o, then the fan-out-32 search.
// The loads that depend only on (w, o) are taken off the dependent chain: P_w[o] is gathered once
// per 32-token batch (one lane per token) and, for narrow rows, this lane's entry of the top search
// level is requested at the start of the token step, before the bucket is known. A draw that lands
// in the prior bucket then pays one dependent memory access per remaining level only.
// LIVE: rows are rewritten during the sweep (the copy that is not current), so they are read at L2,
// never through the non-coherent path.
#ifndef B200LDA_PRIOR_CG
#define B200LDA_PRIOR_CG 1
#endif
template <bool LIVE>
__device__ __forceinline__ float prior_load(const float* a) { return (LIVE && B200LDA_PRIOR_CG) ? __ldcg(a) : __ldg(a); }
// rw = the row index: the word (one copy) or 2 word + copy (LIVE: two copies per word)
__device__ __forceinline__ const float* prior_row_ptr(const SweepParams& p, uint32_t rw) {
  const char* r = reinterpret_cast<const char*>(p.prior) + (size_t)rw * (size_t)(4u * (uint32_t)p.layout.stride);
  asm volatile("" : "+l"(r));  // opaque: one IMAD.WIDE per token, every level load is an offset from it
  return reinterpret_cast<const float*>(r);
}
template <bool LIVE>
__device__ __forceinline__ float prior_top_entry(const float* prow, const WarpCtx& c) {
  return (c.top_lane >= 0) ? prior_load<LIVE>(prow + (uint32_t)c.top_lane) : 0.0f;
}
template <bool LIVE>
__device__ __forceinline__ int prior_search(const SweepParams& p, const float* prow, int lane, int K, float po, float y,
                                            float delta, float vtop) {
  const float pod = fsub(po, delta);
  const float s = (y < pod) ? y : fadd(y, delta);
  const int top = p.layout.nlev - 1;
  const int ntop = p.layout.size[top];  // <= 32
  const unsigned bt = __ballot_sync(kFullMask, (lane < ntop) && (vtop > s));
  int block = bt ? (__ffs(bt) - 1) : (ntop - 1);
  for (int lev = top - 1; lev >= 1; --lev) {  // middle levels (K > 1024 only)
    const int lo = block << 5;
    const int nvalid = min(32, p.layout.size[lev] - lo);
    float v = 0.0f;
    if (lane < nvalid) v = prior_load<LIVE>(prow + (uint32_t)(p.layout.off[lev] + lo + lane));
    const unsigned b = __ballot_sync(kFullMask, (lane < nvalid) && (v > s));
    block = lo + (b ? (__ffs(b) - 1) : (nvalid - 1));
  }
  if (top >= 1) {  // level 0: the K prefix sums themselves, at offset 0 of the word's block
    const int lo = block << 5;
    const int nvalid = min(32, K - lo);
    float v = 0.0f;
    if (lane < nvalid) v = prior_load<LIVE>(prow + (uint32_t)(lo + lane));
    const unsigned b = __ballot_sync(kFullMask, (lane < nvalid) && (v > s));
    block = lo + (b ? (__ffs(b) - 1) : (nvalid - 1));
  }
  return block;
}

// Count moves of one token: -1 at the old topic from lane 0, +1 at the new topic from lane 1.
// Word-topic totals: integer RED atomics on the word's n_wk row (order-independent sums). Topic
// totals: every move in the sweep would hit the same K addresses, so they are accumulated per CTA
// in shared memory and flushed once at the end of the kernel.
template <bool LIVE>
__device__ __forceinline__ int32_t* write_row(const SweepParams& p, const int32_t* nrow, int w, int K) {
  if (LIVE) return const_cast<int32_t*>(nrow);  // in place: nwk_write == nwk_read
  return p.nwk_write ? p.nwk_write + (size_t)w * K : nullptr;
}
__device__ __forceinline__ void count_moves(const SweepParams& p, const WarpCtx& c, int32_t* wrow, int o, int newt) {
  if (c.lane < 2 && wrow != nullptr) {
    const int topic = c.lane == 0 ? o : newt;
    const int val = c.lane == 0 ? -1 : 1;
    atomicAdd(wrow + topic, val);
    if (c.nkd_in_smem) {
      atomicAdd(smem_i32_ptr(c.tab + 2 * c.K + topic), val);
    } else {
      atomicAdd(p.nk_delta + topic, val);
    }
  }
}

__device__ __forceinline__ int warp_scan_inclusive_i32(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int y = __shfl_up_sync(kFullMask, v, d);
    if (lane >= d) v += y;
  }
  return v;
}

// Visit-start split of a packed row of n slots over nt = n/32 + 1 tiles: tile g takes
// q + [g < r] consecutive sorted positions starting at g q + min(g, r)  (q = n / nt, r = n % nt).
struct RowSplit {
  int q, r;
  __device__ __forceinline__ int begin(int g) const { return g * q + min(g, r); }
  __device__ __forceinline__ int size(int g) const { return q + (g < r ? 1 : 0); }
};
__device__ __forceinline__ RowSplit row_split(int n, int nt) {
  RowSplit s;
  s.q = n / nt;
  s.r = n - s.q * nt;
  return s;
}

// Write-back bitmap, step 2 and 3 (after every live slot has set its topic's bit): exclusive
// popcount prefix per bitmap word; returns the number of live slots.
__device__ __forceinline__ int bitmap_prefix(const WarpCtx& c) {
  int running = 0;
  for (int base = 0; base < c.bmw; base += 32) {
    const int j = base + c.lane;
    const int cnt = j < c.bmw ? __popc(smem_u32(c.bm + j)) : 0;
    const int incl = warp_scan_inclusive_i32(cnt, c.lane);
    if (j < c.bmw) smem_u32(c.bm + c.bmw + j) = (uint32_t)(running + incl - cnt);
    running += __shfl_sync(kFullMask, incl, 31);
  }
  return running;
}
__device__ __forceinline__ int bitmap_rank(const WarpCtx& c, int topic) {
  const int wd = topic >> 5;
  return (int)smem_u32(c.bm + c.bmw + wd) + __popc(smem_u32(c.bm + wd) & ((1u << (topic & 31)) - 1u));
}

// ---- register path ------------------------------------------------------------------------------
// sv / wt hold MAXNT tiles; tiles at or beyond nt are all-dead (0 / +0). NT == nt.
// nv holds the n_wk counts of THIS token's word at the row's topics, requested one token ago; the
// step requests the next token's counts from the row as it stands (before this token's move is
// known) and leaves them in nv: a move changes the row's topic set in at most one slot, which is
// re-gathered by the lane that takes it. The gather latency is thereby off the token-to-token
// dependent chain. In LIVE mode the request precedes this token's own count moves, so when the
// next token is of the same word the two affected counts are fixed up in registers.
// Branch conditions that are warp-uniform by construction (bucket, moved, new-to-row) go through
// votes: the compiler then emits uniform branches without divergence bookkeeping around the warp
// collectives inside.
template <int NT, int MAXNT, int MODE, bool LIVE, bool TS, int TE>
__device__ __forceinline__ int token_step(const SweepParams& p, WarpCtx& c, uint32_t (&sv)[MAXNT], float (&wt)[MAXNT],
                                          int (&nv)[MAXNT], int& nt, unsigned& dead, int bnd, uint32_t tok_addr) {
  constexpr bool EXCL = MODE != MODE_INFER;
  const int lane = c.lane;
  const uint4 ta = lds_u128(tok_addr);       // prior row index (LIVE: 2 word + copy, else word), old topic, uniform, Q_w
  const uint4 tb = lds_u128(tok_addr + 16);  // invden[o], ab[o] (0 in inference), P_w[o], next token's word
  const uint32_t w = LIVE ? ta.x >> 1 : ta.x;
  const int o = (int)ta.y;
  const float u = __uint_as_float(ta.z), qw = __uint_as_float(ta.w);
  const float inv_o = __uint_as_float(tb.x), delta = __uint_as_float(tb.y);
  const char* nwk_bytes = reinterpret_cast<const char*>(p.nwk_read);
  const char* nrow_next_b = nwk_bytes + (size_t)tb.w * (size_t)c.row_bytes;
  asm volatile("" : "+l"(nrow_next_b));  // opaque: per-gather address = one IMAD.WIDE off this pointer
  const int32_t* nrow_next = reinterpret_cast<const int32_t*>(nrow_next_b);
  int nvn[NT];
#pragma unroll
  for (int g = 0; g < NT; ++g)  // a dead slot reads a valid cell; its weight is +0
    nvn[g] = count_load<LIVE>(nrow_next + (sv[g] >> 16));
  const float* prow = prior_row_ptr(p, ta.x);
  constexpr bool kTopEarly = NT <= TE;  // request the prior's top level before the bucket is known
  float vtop = 0.0f;
  if (kTopEarly) vtop = prior_top_entry<LIVE>(prow, c);
  // a live slot of topic o reads (o << 16) + count with 1 <= count <= 0xffff
  const uint32_t okey = ((uint32_t)o << 16) + 1u;
  // Lane-strided prefix: the lane sums its own slots tile by tile (from +0), ONE warp scan runs over
  // the 32 lane totals, slot (g, lane) gets P = E_lane + s[g]. The cumulative order is lane-major.
  float s[NT];
  float run = 0.0f;
#pragma unroll
  for (int g = 0; g < NT; ++g) {
    const bool is_old = (sv[g] - okey) < 0xffffu;
    int n = nv[g];
    if (EXCL && is_old) n -= 1;
    n = max(n, 0);
    const float wv = is_old ? fsub(wt[g], inv_o) : wt[g];
    run = fadd(run, fmul(fadd((float)n, c.beta_f), wv));
    s[g] = run;
  }
  const float incl = warp_scan_inclusive(run, lane);
  const float E = shfl_up1_or_zero(incl);
  const float A = __shfl_sync(kFullMask, incl, 31);
  float qp = fsub(qw, delta);
  qp = qp < 0.0f ? 0.0f : qp;
  const float x = fmul(u, fadd(A, qp));

  int newt = o;
  const bool doc_bucket = __any_sync(kFullMask, x < A);
  if (doc_bucket) {
    // First slot in cumulative (lane-major) order whose prefix exceeds x: the lowest lane with a
    // hit and, prefixes being non-decreasing inside a lane, its count of non-hits. Within an ulp of
    // a lane boundary that slot can be one that adds no weight (dead, or the token's own slot when
    // the token is its only one), or there is none at all: the token then keeps its topic.
    int cnt = 0;
#pragma unroll
    for (int g = 0; g < NT; ++g) cnt += (fadd(E, s[g]) > x) ? 0 : 1;
    const unsigned b = __ballot_sync(kFullMask, cnt < NT);
    uint32_t mine = sv[0];
#pragma unroll
    for (int g = 1; g < NT; ++g) mine = (cnt == g) ? sv[g] : mine;
    const uint32_t pk = __shfl_sync(kFullMask, mine, __ffs(b) - 1);  // b == 0: lane 31's slot, rejected below unless valid... (see guard)
    if (b != 0u && (pk & 0xffffu) != 0u && pk != okey) newt = (int)(pk >> 16);
  } else {
    ++c.st_prior;
    if (!kTopEarly) vtop = prior_top_entry<LIVE>(prow, c);
    newt = prior_search<LIVE>(p, prow, lane, c.K, __uint_as_float(tb.z), fsub(x, A), delta, vtop);
  }

  if (MODE != MODE_FROZEN && __any_sync(kFullMask, newt != o)) {
    ++c.st_moved;
    const float inv_n = TS ? lds_f32(c.sbase + ((uint32_t)newt << 2)) : __ldg(p.invden + newt);
    const uint32_t nkey = ((uint32_t)newt << 16) + 1u;
    const bool same_word = LIVE && tb.w == w;
    // Each lane edits its own slots: -1 at the old topic's slot (it dies in place at count 0),
    // +1 at the new topic's slot when the document already has it.
    bool has_new = false;
#pragma unroll
    for (int g = 0; g < NT; ++g) {
      const bool io = (sv[g] - okey) < 0xffffu;
      const bool in = (sv[g] - nkey) < 0xffffu;
      if (io | in) {
        sv[g] += in ? 1u : 0xffffffffu;
        wt[g] = fmul(in ? inv_n : inv_o, (float)(sv[g] & 0xffffu));
        if ((sv[g] & 0xffffu) == 0u) dead |= 1u << g;  // the old topic left the document: its slot is dead
        if (same_word) nvn[g] += in ? 1 : -1;  // the request preceded this token's own count moves
      }
      has_new = has_new || in;
    }
    if (!doc_bucket && !__any_sync(kFullMask, has_new)) {
      // The topic is new to the document: lowest dead lane of its preferred tile (tiles are the
      // topic ranges fixed at visit start: bnd = first topic of tile `lane`), else the lowest lane
      // with a dead slot anywhere, at its lowest dead tile; none: a new tile's lane 0.
      const int gstar = __popc(__ballot_sync(kFullMask, newt >= bnd));
      unsigned b = __ballot_sync(kFullMask, (dead >> gstar) & 1u);
      int mytile = gstar;
      if (b == 0u) {
        b = __ballot_sync(kFullMask, dead != 0u);
        mytile = __ffs(dead) - 1;
        if (b == 0u) {  // every slot live: append a tile (its registers are already dead slots)
          if (NT < MAXNT) dead |= 1u << NT;
          b = NT < MAXNT ? 1u : 0u;  // NT == MAXNT: nt = MAXNT + 1 tells the caller to continue in shared memory
          mytile = NT;
          nt = NT + 1;
        }
      }
      if (lane == __ffs(b) - 1) {
        // the next token's count at the slot's new topic, requested before this token's +1 is issued
        const int fresh = count_load<LIVE>(nrow_next + (uint32_t)newt) + (same_word ? 1 : 0);
        dead &= ~(1u << mytile);
#pragma unroll
        for (int g = 0; g < (NT < MAXNT ? NT + 1 : NT); ++g) {
          if (g == mytile) {
            sv[g] = nkey;
            wt[g] = inv_n;
            if (g < NT) nvn[g] = fresh; else nv[g < MAXNT ? g : 0] = fresh;
          }
        }
      }
    }
    // count moves: -1 at the old topic from lane 0, +1 at the new topic from lane 1
    if (LIVE || p.nwk_write != nullptr) {
      if (lane < 2) {
        const uint32_t topic = (uint32_t)(lane == 0 ? o : newt);
        const int val = lane == 0 ? -1 : 1;
        char* wbase = LIVE ? const_cast<char*>(nwk_bytes) : reinterpret_cast<char*>(p.nwk_write);
        atomicAdd(reinterpret_cast<int32_t*>(wbase + (size_t)w * (size_t)c.row_bytes) + topic, val);
        if (TS) {
          reds_add_i32(c.sbase + 2u * c.row_bytes + (topic << 2), val);
        } else {
          atomicAdd(p.nk_delta + topic, val);
        }
      }
    }
  }
#pragma unroll
  for (int g = 0; g < NT; ++g) nv[g] = nvn[g];
  return newt;
}

// Wide path: a topic new to the document takes the lowest dead lane of its preferred tile, else the
// lowest lane with a dead slot anywhere (at its lowest dead tile), else lane 0 of an appended tile.
__device__ __forceinline__ void wide_insert(const WarpCtx& c, int& nt, int newt) {
  const int lane = c.lane;
  const int SV = c.row, BD = c.row + 64 * c.capT;
  const uint32_t nkey = ((uint32_t)newt << 16) + 1u;
  int ge = 0;
  for (int g = 1 + lane; g < nt; g += 32) ge += (newt >= (int)smem_u32(BD + g)) ? 1 : 0;
  const int gstar = __reduce_add_sync(kFullMask, ge);
  int mytile = gstar;
  unsigned b = __ballot_sync(kFullMask, (smem_u32(SV + (gstar << 5) + lane) & 0xffffu) == 0u);
  if (b == 0u) {
    mytile = -1;
    for (int g = nt - 1; g >= 0; --g)
      if ((smem_u32(SV + (g << 5) + lane) & 0xffffu) == 0u) mytile = g;
    b = __ballot_sync(kFullMask, mytile >= 0);
    if (b == 0u) {  // every slot live: append an empty tile (nt < capT by the class's document lengths)
      smem_u32(SV + (nt << 5) + lane) = 0u;
      if (lane == 0) smem_u32(BD + nt) = (uint32_t)c.K;
      mytile = nt;
      b = 1u;
      nt += 1;
      __syncwarp();
    }
  }
  if (lane == __ffs(b) - 1) smem_u32(SV + (mytile << 5) + lane) = nkey;
  __syncwarp();
}

// ---- wide path: row in shared memory, loops over tiles ---------------------------------------------
template <int MODE, bool LIVE, bool TS>
__device__ __forceinline__ int token_step_wide(const SweepParams& p, WarpCtx& c, int& nt, int tok) {
  constexpr bool EXCL = MODE != MODE_INFER;
  const int lane = c.lane;
  const int SV = c.row, PS = c.row + 32 * c.capT;
  const uint4 ta = smem_u128(tok);
  const uint32_t w = LIVE ? ta.x >> 1 : ta.x;
  const int o = (int)ta.y;
  const float u = __uint_as_float(ta.z), qw = __uint_as_float(ta.w);
  const int32_t* nrow = p.nwk_read + (size_t)w * (size_t)(uint32_t)c.K;
  const float* prow = prior_row_ptr(p, ta.x);
  const uint4 tb = smem_u128(tok + 4);
  const float inv_o = __uint_as_float(tb.x), delta = __uint_as_float(tb.y);
  const uint32_t okey = ((uint32_t)o << 16) + 1u;
  // Tiles go in groups of kGroup: all of a group's n_wk gathers are issued before any is consumed,
  // so a wide row pays one memory latency per group, not per tile.
  float run = 0.0f;
  for (int t0 = 0; t0 < nt; t0 += kGroup) {
    uint32_t svg[kGroup];
    int nvg[kGroup];
#pragma unroll
    for (int g = 0; g < kGroup; ++g) {
      svg[g] = 0u;
      nvg[g] = 0;
      if (t0 + g < nt) {
        svg[g] = smem_u32(SV + ((t0 + g) << 5) + lane);
        nvg[g] = count_load<LIVE>(nrow + (svg[g] >> 16));
      }
    }
#pragma unroll
    for (int g = 0; g < kGroup; ++g) {
      if (t0 + g < nt) {
        const int j = ((t0 + g) << 5) + lane;
        const bool is_old = (svg[g] - okey) < 0xffffu;
        int n = nvg[g];
        if (EXCL && is_old) n -= 1;
        n = max(n, 0);
        const int topic = (int)(svg[g] >> 16);  // dead slot: count 0, weight +0
        const float wslot = fmul(TS ? smem_f32(c.tab + topic) : __ldg(p.invden + topic), (float)(svg[g] & 0xffffu));
        const float wv = is_old ? fsub(wslot, inv_o) : wslot;
        run = fadd(run, fmul(fadd((float)n, c.beta_f), wv));
        smem_f32(PS + j) = run;
      }
    }
  }
  const float incl = warp_scan_inclusive(run, lane);
  const float E = shfl_up1_or_zero(incl);
  const float A = __shfl_sync(kFullMask, incl, 31);
  float qp = fsub(qw, delta);
  qp = qp < 0.0f ? 0.0f : qp;
  const float x = fmul(u, fadd(A, qp));

  int newt = o;
  const bool doc_bucket = x < A;
  if (doc_bucket) {
    // the lane's largest prefix is at its last tile, where the local sum is `run`
    const unsigned b = __ballot_sync(kFullMask, fadd(E, run) > x);
    if (b) {
      const int src = __ffs(b) - 1;
      uint32_t pk = 0u;
      if (lane == src) {
        int j = lane;
        while (!(fadd(E, smem_f32(PS + j)) > x)) j += 32;  // ends at the lane's last tile at the latest
        pk = smem_u32(SV + j);
      }
      pk = __shfl_sync(kFullMask, pk, src);
      if ((pk & 0xffffu) != 0u && pk != okey) newt = (int)(pk >> 16);
    }
  } else {
    ++c.st_prior;
    newt = prior_search<LIVE>(p, prow, lane, c.K, __uint_as_float(tb.z), fsub(x, A), delta, prior_top_entry<LIVE>(prow, c));
  }

  if (MODE != MODE_FROZEN && newt != o) {
    ++c.st_moved;
    const uint32_t nkey = ((uint32_t)newt << 16) + 1u;
    bool has_new = false;
    for (int g = 0; g < nt; ++g) {
      const int j = (g << 5) + lane;
      const uint32_t v = smem_u32(SV + j);
      const bool io = (v - okey) < 0xffffu;
      const bool in = (v - nkey) < 0xffffu;
      if (io | in) smem_u32(SV + j) = v + (in ? 1u : 0xffffffffu);
      has_new = has_new || in;
    }
    __syncwarp();
    if (!doc_bucket && !__any_sync(kFullMask, has_new)) wide_insert(c, nt, newt);
    count_moves(p, c, write_row<LIVE>(p, nrow, (int)w, c.K), o, newt);
  }
  return newt;
}

// LIVE mode table refresh: a small kernel launched on a side stream right before a bulk sampling
// launch, whose grid is that many CTAs short of filling the GPU. Each warp claims the
// next row index, waits (bounded sleeps) until the samplers have completed that fraction of the
// launch's document chunks (read off the scheduler counter: every sampler warp's next fetch says
// its previous chunk is complete), rebuilds the prior row of the hot word in turn from the live
// counts into the copy that is not current and makes it current. It only READS the scheduler
// counter, leaves as soon as the samplers have finished, and gives up when the counter has not
// moved for ~100 us (the samplers are not running beside it): no sampler ever waits for it and it
// never waits for them longer than that. Readers picked a copy through prior_sel before reading; the copy
// they read stays untouched until the word's NEXT rebuild, a full pass over the hot list later.
__global__ void __launch_bounds__(256, 8) k_prior_refresher(const SweepParams p, unsigned long long nchunks) {
  const int lane = threadIdx.x & 31;
  const unsigned long long sampler_warps = (unsigned long long)p.sampler_warps;
  for (;;) {
    unsigned idx = 0;
    if (lane == 0) idx = atomicAdd(p.row_cursor, 1u);
    idx = __shfl_sync(kFullMask, idx, 0);
    if (idx >= p.refresh_rows) return;
    unsigned long long seen = ~0ull;
    for (int idle = 0;;) {
      unsigned long long done = 0;
      if (lane == 0) done = *reinterpret_cast<volatile unsigned long long*>(p.doc_counter);
      done = __shfl_sync(kFullMask, done, 0);
      // Samplers that run beside this kernel move the counter every few hundred ns (thousands of
      // warps fetch chunks). A counter that stands still for ~100 us means they are NOT running
      // beside it - two streams that share a hardware queue are serialised, and then the samplers
      // are waiting for THIS kernel to end: leave at once (no refresh this launch, no stall) ...
      // ... once the counter has moved the samplers ARE running (a small launch hands all its chunks
      // out at once and is then quiet until the first warp finishes one): only the hard bound applies.
      idle = done == seen ? idle + 1 : 0;
      seen = done;
      if (idle > (done == 0ull ? 50 : 10000)) return;
      done = done > sampler_warps ? done - sampler_warps : 0ull;
      if (done >= nchunks) return;  // the samplers are done: nothing left to refresh for
      if (done * (unsigned long long)p.refresh_rows >= (unsigned long long)idx * nchunks) break;
      __nanosleep(2000);
    }
    const unsigned w = (unsigned)__ldg(p.hot_words + idx % (unsigned)p.hot_count);
    const int copy = 1 - __ldcg(p.prior_sel + w);
    const size_t rw = 2 * (size_t)w + (size_t)copy;
    float* out = const_cast<float*>(p.prior) + rw * (size_t)p.layout.stride;
    const float Q = build_prior_row<true>(p.K, p.nwk_read + (size_t)w * (size_t)p.K, p.ab, p.beta_f, p.layout, out, lane);
    if (lane == 0) const_cast<float*>(p.q)[rw] = Q;
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      *reinterpret_cast<volatile int32_t*>(p.prior_sel + w) = copy;
      atomicAdd(p.refresh_count, 1ull);
    }
  }
}

template <int MODE, bool LIVE, bool TABLES_IN_SMEM, int ROWCLASS>
__global__ void __launch_bounds__(256, rowclass_min_ctas(ROWCLASS)) k_gibbs_sweep(const SweepParams p) {
  // The wide class is a hybrid: a visit keeps the row in registers while it fits 8 tiles and moves it
  // to shared memory when it starts wider or grows a 9th tile (long documents on few topics, and
  // C3's 333-token documents with ~160 topics, stay on the register path).
  constexpr bool HYBRID = ROWCLASS == kWideClass;
  constexpr int MAXNT = HYBRID ? 8 : rowclass_max_tiles(ROWCLASS);
  // Rows up to kTE tiles request the prior's top search level at the start of the token step.
  constexpr int kTE = ROWCLASS <= 1 ? 5 : B200LDA_TOP_EARLY_NT;
  int lane = threadIdx.x & 31;
  asm volatile("" : "+r"(lane));  // opaque: otherwise rematerialised (S2R + LOP) at every use in the token step
  const int warp = __shfl_sync(kFullMask, (int)(threadIdx.x >> 5), 0);  // tells the compiler it is warp-uniform
  const int K = p.K;

  const int tab_words = TABLES_IN_SMEM ? sweep_table_words(K) : 0;
  if (TABLES_IN_SMEM) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      smem_f32(k) = p.invden[k];
      smem_f32(K + k) = p.ab[k];
      smem_u32(2 * K + k) = 0u;
    }
    __syncthreads();
  }
  WarpCtx c;
  c.lane = lane;
  c.beta_f = p.beta_f;
  c.K = K;
  c.tab = 0;
  c.bmw = bitmap_words(K);
  c.capT = p.cap_tiles;
  c.batch = tab_words + warp * sweep_warp_words(K, p.cap_tiles);
  c.bm = c.batch + kBatchWords;
  c.row = c.bm + 2 * c.bmw;
  c.nkd_in_smem = TABLES_IN_SMEM;
  c.top_lane = (lane < p.layout.size[p.layout.nlev - 1]) ? p.layout.off[p.layout.nlev - 1] + lane : -1;
  c.st_moved = 0;
  c.st_prior = 0;
  c.sbase = smem_window_addr();
  c.batch_addr = c.sbase + 4u * (uint32_t)c.batch;
  asm volatile("" : "+r"(c.batch_addr));
  c.row_bytes = 4u * (uint32_t)K;
  asm volatile("" : "+r"(c.row_bytes));
  const int SV = c.row, BD = c.row + 64 * c.capT;

  unsigned long long st_moved = 0, st_prior = 0, st_nnz = 0;
  const unsigned long long ndocs = (unsigned long long)(p.order_end - p.order_begin);

  // Dynamic scheduler: one atomic fetches a chunk of doc_chunk documents. Documents are ordered
  // longest first, so a chunk is STRIDED through that order (chunk ci = documents ci, ci + nchunks,
  // ci + 2 nchunks, ...): every chunk holds one document of each length band instead of the first
  // chunk holding the doc_chunk longest documents of the class (measured on C1, where that chunk
  // alone was the whole sweep: 1.6 -> 0.6 ms).
  const unsigned long long nchunks = (ndocs + (unsigned long long)p.doc_chunk - 1) / (unsigned long long)p.doc_chunk;
  for (;;) {
    unsigned long long ci = 0;
    if (lane == 0) ci = p.chunk_begin + atomicAdd(p.doc_counter, 1ull);
    ci = __shfl_sync(kFullMask, ci, 0);
    if (ci >= p.chunk_end) break;

    for (unsigned long long di = ci; di < ndocs; di += nchunks) {
      // the document's header, read by lane 0 and broadcast: warp-uniform for the compiler too
      const int64_t d = (int64_t)__shfl_sync(kFullMask, __ldg(p.doc_order + p.order_begin + (int64_t)di), 0);
      long long h_tb = 0, h_te = 0, h_rp = 0;
      int h_nnz = 0;
      if (lane == 0) {
        h_tb = p.doc_ptr[d];
        h_te = p.doc_ptr[d + 1];
        h_rp = p.row_ptr[d];
        h_nnz = p.row_nnz[d];
      }
      const int64_t tb = __shfl_sync(kFullMask, h_tb, 0), te = __shfl_sync(kFullMask, h_te, 0);
      if (te == tb) continue;
      const int64_t rp = __shfl_sync(kFullMask, h_rp, 0);
      int nnz = __shfl_sync(kFullMask, h_nnz, 0);
      const int nnz_start = nnz;

      const int iters = MODE == MODE_INFER ? p.infer_iters : 1;
      for (int it = 1; it <= iters; ++it) {
        const uint32_t sweep_key = MODE == MODE_INFER ? (uint32_t)it : p.sweep;

        // ---- visit start: the packed row (ascending topics) split evenly over nnz/32 + 1 tiles
        int nt = (nnz >> 5) + 1;
        const RowSplit sp = row_split(nnz, nt);
        uint32_t sv[MAXNT];
        float wt[MAXNT];
        int nv[MAXNT];
        int bnd = 0x7fffffff;  // register path: first topic of tile `lane` (1 <= lane < nt)
        unsigned dead = 0u;    // register path: bit g = this lane's slot of tile g (g < nt) is dead
        bool wide = HYBRID && nt > MAXNT;
        if (!wide) {
#pragma unroll
          for (int g = 0; g < MAXNT; ++g) {
            sv[g] = 0u;
            wt[g] = 0.0f;
            nv[g] = 0;
            if (g < nt && lane >= sp.size(g)) dead |= 1u << g;
            if (g < nt && lane < sp.size(g)) {
              sv[g] = p.rows[rp + sp.begin(g) + lane];
              const int topic = (int)(sv[g] >> 16);
              const float inv = TABLES_IN_SMEM ? smem_f32(c.tab + topic) : __ldg(p.invden + topic);
              wt[g] = fmul(inv, (float)(sv[g] & 0xffffu));
            }
          }
          if (lane >= 1 && lane < nt) bnd = sp.size(lane) > 0 ? (int)(p.rows[rp + sp.begin(lane)] >> 16) : K;
        } else {
          for (int g = 0; g < nt; ++g) {
            uint32_t v = 0u;
            if (lane < sp.size(g)) v = p.rows[rp + sp.begin(g) + lane];
            smem_u32(SV + (g << 5) + lane) = v;
          }
          for (int g = 1 + lane; g < nt; g += 32)
            smem_u32(BD + g) = sp.size(g) > 0 ? (p.rows[rp + sp.begin(g)] >> 16) : (uint32_t)K;
          __syncwarp();
        }

        for (int64_t base = tb; base < te; base += 32) {
          const int64_t i = base + lane;
          const bool valid = i < te;
          const int w_l = valid ? __ldg(p.tok_word + i) : 0;
          const int o_l = valid ? (int)p.z[i] : 0;
          float u_l = 0.0f;
          if (valid) {
            if (p.uniforms) {
              u_l = __ldg(p.uniforms + i);
            } else {
              u_l = u24(token_random(p.seed, (uint64_t)(p.global_tok_off + i), sweep_key, 0u).x);
            }
          }
          uint32_t rw_l = (uint32_t)w_l;  // the word's current copy of its prior row / Q_w
          if (LIVE) rw_l = 2u * (uint32_t)w_l + (valid ? (uint32_t)__ldcg(p.prior_sel + w_l) : 0u);
          const float q_l = valid ? prior_load<LIVE>(p.q + rw_l) : 0.0f;
          const float po_l = valid ? prior_load<LIVE>(p.prior + (size_t)rw_l * p.layout.stride + o_l) : 0.0f;  // P_w[o]
          const float inv_l = TABLES_IN_SMEM ? smem_f32(c.tab + o_l) : __ldg(p.invden + o_l);
          float dl_l = 0.0f;  // inference: nothing of the document is in the table
          if (MODE != MODE_INFER) dl_l = TABLES_IN_SMEM ? smem_f32(c.tab + K + o_l) : __ldg(p.ab + o_l);
          const int cnt = (int)min((int64_t)32, te - base);
          // the word whose counts the step requests for the NEXT token (its own word again at the
          // batch's last token: a harmless extra request)
          int wn_l = __shfl_down_sync(kFullMask, w_l, 1);
          if (lane + 1 >= cnt) wn_l = w_l;
          __syncwarp();  // the previous batch has been consumed
          smem_u128(c.batch + 8 * lane) = make_uint4(rw_l, (uint32_t)o_l, __float_as_uint(u_l), __float_as_uint(q_l));
          smem_u128(c.batch + 8 * lane + 4) =
              make_uint4(__float_as_uint(inv_l), __float_as_uint(dl_l), __float_as_uint(po_l), (uint32_t)wn_l);
          __syncwarp();
          int new_l = o_l;
          if (!wide) {  // the batch's first token: nobody requested its counts
            const int w0 = __shfl_sync(kFullMask, w_l, 0);
            const int32_t* nrow0 = p.nwk_read + (size_t)w0 * K;
#pragma unroll
            for (int g = 0; g < MAXNT; ++g)
              if (g < nt) nv[g] = count_load<LIVE>(nrow0 + (sv[g] >> 16));
          }

          uint32_t tok_addr = c.batch_addr;
          for (int t = 0; t < cnt; ++t, tok_addr += 32u) {
            int newt;
            if (wide) {
              newt = token_step_wide<MODE, LIVE, TABLES_IN_SMEM>(p, c, nt, c.batch + 8 * t);
            } else {
              // nt is uniform across the warp; instantiations beyond the class's widest row are not generated
#define B200LDA_STEP(N) token_step<(N <= MAXNT ? N : 1), MAXNT, MODE, LIVE, TABLES_IN_SMEM, kTE>(p, c, sv, wt, nv, nt, dead, bnd, tok_addr)
              if (nt == 1) newt = B200LDA_STEP(1);
              else if (nt == 2) newt = B200LDA_STEP(2);
              else if (nt == 3 || MAXNT == 3) newt = B200LDA_STEP(3);
              else if (nt == 4) newt = B200LDA_STEP(4);
              else if (nt == 5 || MAXNT == 5) newt = B200LDA_STEP(5);
              else if (nt == 6) newt = B200LDA_STEP(6);
              else if (nt == 7) newt = B200LDA_STEP(7);
              else newt = B200LDA_STEP(8);
#undef B200LDA_STEP
              if (HYBRID && nt > MAXNT) {
                // the row needs a 9th tile: it moves to shared memory for the rest of the visit and
                // the topic the step could not place is inserted there
                nt = MAXNT;
#pragma unroll
                for (int g = 0; g < MAXNT; ++g) smem_u32(SV + (g << 5) + lane) = sv[g];
                if (lane < MAXNT) smem_u32(BD + lane) = (uint32_t)bnd;
                __syncwarp();
                wide_insert(c, nt, newt);
                wide = true;
              }
            }
            if (lane == t) new_l = newt;
          }

          if (valid) {
            if (MODE == MODE_FROZEN) {
              p.z_out[i] = new_l;
            } else if (new_l != o_l) {
              p.z[i] = (uint16_t)new_l;
            }
          }
        }

        // ---- visit end: live slots back to the packed row in ascending topic order
        if (MODE != MODE_FROZEN) {
          for (int j = lane; j < c.bmw; j += 32) smem_u32(c.bm + j) = 0u;
          __syncwarp();
          if (!wide) {
#pragma unroll
            for (int g = 0; g < MAXNT; ++g)
              if (sv[g] & 0xffffu) atomicOr(smem_u32_ptr(c.bm + (sv[g] >> 21)), 1u << ((sv[g] >> 16) & 31u));
          } else {
            for (int g = 0; g < nt; ++g) {
              const uint32_t v = smem_u32(SV + (g << 5) + lane);
              if (v & 0xffffu) atomicOr(smem_u32_ptr(c.bm + (v >> 21)), 1u << ((v >> 16) & 31u));
            }
          }
          __syncwarp();
          nnz = bitmap_prefix(c);
          __syncwarp();
          const bool save = MODE == MODE_INFER &&
                            (p.infer_samples == 0 ? it == iters
                                                  : (it > p.infer_burn_in && (it - p.infer_burn_in) % p.infer_thinning == 0));
          int32_t* acc = MODE == MODE_INFER ? p.infer_acc + (size_t)d * K : nullptr;  // this warp owns document d
          if (!wide) {
#pragma unroll
            for (int g = 0; g < MAXNT; ++g)
              if (sv[g] & 0xffffu) {
                p.rows[rp + bitmap_rank(c, (int)(sv[g] >> 16))] = sv[g];
                if (save) acc[sv[g] >> 16] += (int32_t)(sv[g] & 0xffffu);
              }
          } else {
            for (int g = 0; g < nt; ++g) {
              const uint32_t v = smem_u32(SV + (g << 5) + lane);
              if (v & 0xffffu) {
                p.rows[rp + bitmap_rank(c, (int)(v >> 16))] = v;
                if (save) acc[v >> 16] += (int32_t)(v & 0xffffu);
              }
            }
          }
          __syncwarp();  // the row is read back by other lanes at the next iteration's visit start
        }
      }

      if (MODE != MODE_FROZEN && lane == 0) p.row_nnz[d] = nnz;
      // statistics: sum over the document's tokens of its non-zero topics, taken as the mean of the
      // row width at visit start and end (the kernel does not track the width token by token)
      st_nnz += (unsigned long long)(te - tb) * (unsigned long long)(nnz_start + nnz) * (unsigned long long)iters / 2ull;
      st_moved += c.st_moved;
      st_prior += c.st_prior;
      c.st_moved = 0;
      c.st_prior = 0;
    }
  }

  if (TABLES_IN_SMEM && MODE == MODE_UPDATE && p.nwk_write != nullptr) {
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const int dv = (int)smem_u32(2 * K + k);
      if (dv != 0) atomicAdd(p.nk_delta + k, dv);
    }
  }
  if (lane == 0) {
    if (st_moved) atomicAdd(p.stats + 0, st_moved);
    if (st_prior) atomicAdd(p.stats + 1, st_prior);
    if (st_nnz) atomicAdd(p.stats + 2, st_nnz);
  }
}

}  // namespace b200lda
