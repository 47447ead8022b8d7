// sweep_kernel.cuh — K3/K7: the per-token resampling kernel (one warp per document).
//
// Replaces cc.mallet.topics.WorkerRunnable.sampleTopicsForOneDoc, reached through
// model.estimate() at reference cmu_ron/TrainAndPredict.java:166 and cmu/TrainAndPredict.java:265
// (SURVEY.md §8 rows a4, a5). It is NOT a translation of SparseLDA's s/r/q walk over packed rows:
//   * a warp owns a document; the document's sparse topic row (topic<<16 | count, ascending
//     topic) lives in shared memory and is the only state that changes token to token;
//   * doc bucket: lane j holds non-zero topic j, gathers n_wk[w, topic_j] (4 B each),
//     forms  n_dk (n_wk + beta) / (n_k + V beta)  and the warp scans it (shuffle prefix sum);
//   * prior bucket: alpha_k (n_wk + beta) / (n_k + V beta) is word-only, so its mass and prefix
//     table are built once per sweep per word (table_kernels.cuh); a draw that lands there is
//     resolved by a fan-out-32 search = one coalesced 128-byte line per level, and the loads that
//     depend only on the token (P_w[o], the top level) are requested before the bucket is known;
//   * randomness: Philox keyed by (seed; global token, sweep), 32 tokens per warp batch, one
//     lane each, so the RNG costs ~2 instructions per token;
//   * count moves: the row edit happens in shared memory (two count updates, a one-slot shift only
//     when a topic enters or leaves the document), integer RED atomics carry the n_wk / n_k moves.
// The kernel is limited by instruction issue (72-77 % of issue slots) with L1TEX second
// (profiles/r01_sweep_v7_ncu_summary.txt, profiles/r01_tuning.md), so the token step is specialised
// on the number of 32-slot tiles of the row: token_step_tiles<NT> is straight-line code (rows
// zero-padded to whole tiles so nothing is predicated per lane, prefixes in registers, no inner
// loops); token_step_generic keeps the loop form for rows wider than kRegTiles tiles. Slots stay in
// tile order (lane = slot mod 32) on purpose: 32 consecutive sorted topics per gather instruction
// touch ~13 of the word row's 32 lines, a lane-blocked order touches 32 and is L1TEX-bound.
// Documents are visited through doc_order (longest first); the host launches the kernel once per
// row-width class so that per-warp shared memory follows the row width.
// MODE_UPDATE serves LIVE (LIVE=true: n_wk read through L1/L2 and written in place) and DEFERRED
// (LIVE=false: frozen n_wk through the read-only path, moves go to a second buffer);
// MODE_FROZEN moves nothing (north-star parity mode). MODE_INFER is held-out inference
// (TopicInferencer.getSampledDistribution): n_wk / n_k are frozen, so documents are independent
// chains and a warp runs ALL iterations of its document in one visit, the row staying in shared
// memory, adding the row to the document's sample accumulator at every saved iteration.
#pragma once
#include "device_common.cuh"

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
  PriorLayout layout;
  int K;
  int slot_cap;               // shared-memory slots per warp (multiple of 32)
  int doc_chunk;              // documents fetched per scheduler atomic (strided through doc_order)
  int exclude_self;           // 1: the token being resampled is counted in n_wk (training);
                              // 0: held-out inference against frozen counts (TopicInferencer)
  float beta_f;
  uint64_t seed;
  uint32_t sweep;
  int64_t global_tok_off;
  unsigned long long* doc_counter;  // dynamic document scheduler: next chunk index (starts at 0 for each launch)
  unsigned long long* stats;        // [0] moved, [1] prior-bucket draws, [2] sum of nnz over tokens
  unsigned long long* stats_cum;    // same three, accumulated until b200lda_reset_stats
  // MODE_INFER: iterations 1..infer_iters per document (Philox sweep key = iteration); a sample is
  // saved when it > burn_in and (it - burn_in) % thinning == 0, or after the last iteration when
  // no iteration qualifies (infer_samples == 0): acc[d, k] += n_dk.
  int infer_iters, infer_burn_in, infer_thinning, infer_samples;
  int32_t* infer_acc;               // [D * K]
};

#ifndef B200LDA_SWEEP_GROUP
#define B200LDA_SWEEP_GROUP 4  // measured on B200: 4 > 3 > 2 > 1 (profiles/r01_tuning.md)
#endif
constexpr int kGroup = B200LDA_SWEEP_GROUP;  // generic path: tiles whose gathers are in flight together
#ifndef B200LDA_REG_TILES
#define B200LDA_REG_TILES 8  // measured: C3 (rows ~150 slots) 1.11 -> 1.41 Gtok/s, C4 unchanged
#endif
constexpr int kRegTiles = B200LDA_REG_TILES;  // rows of up to this many tiles take the straight-line register path

// Shared memory per warp: slot_cap x {uint32 row slot, float prefix} = 8 bytes per slot, plus
// kRowPad zeroed spare slots behind the row (the register path reads whole tiles).
constexpr int kSmemBytesPerSlot = 8;
constexpr int kRowPad = 32;
__host__ __device__ constexpr size_t sweep_smem_per_warp(int slot_cap) {
  return (size_t)kSmemBytesPerSlot * (size_t)slot_cap + sizeof(uint32_t) * kRowPad;
}

#ifndef B200LDA_TOP_EARLY_NT
#define B200LDA_TOP_EARLY_NT 3  // rows of up to this many tiles request the prior's top level at the start of the token step
#endif
#ifndef B200LDA_SWEEP_MIN_CTAS
#define B200LDA_SWEEP_MIN_CTAS 4   // 8-warp CTAs per SM the register allocation must allow
#endif
#ifndef B200LDA_ROWCLASS0_MIN_CTAS
#define B200LDA_ROWCLASS0_MIN_CTAS 4
#endif
#ifndef B200LDA_ROWCLASS1_MIN_CTAS
#define B200LDA_ROWCLASS1_MIN_CTAS 4
#endif

// The kernel's dynamic shared memory, addressed by WORD OFFSET everywhere: indexing the extern
// array keeps every access a plain LDS/STS/ATOMS with an immediate base, whereas pointers carried in
// a struct decay to generic addresses and cost an address-space conversion per access.
extern __shared__ __align__(16) unsigned char b200lda_smem_raw[];
__device__ __forceinline__ uint32_t& smem_u32(int word) { return reinterpret_cast<uint32_t*>(b200lda_smem_raw)[word]; }
__device__ __forceinline__ float& smem_f32(int word) { return reinterpret_cast<float*>(b200lda_smem_raw)[word]; }
__device__ __forceinline__ int* smem_i32_ptr(int word) { return reinterpret_cast<int*>(b200lda_smem_raw) + word; }

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
__device__ __forceinline__ int live_load(const int32_t* cell) {
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
  int excl;
  int K;
  // word offsets into the kernel's shared memory
  int slots;   // this warp's row
  int pref;    // this warp's prefix scratch (generic path)
  int tab;     // [invden | ab | n_k delta] (TABLES_IN_SMEM): invden at tab, ab at tab + K, deltas at tab + 2K
  bool nkd_in_smem;
  int top_lane;  // offset of this lane's entry of the top search level inside a word's prior block, -1: none
  unsigned st_moved, st_prior;
};

// Prior bucket: skip the own-token mass delta at topic o, then the fan-out-32 search.
// The loads that depend only on (w, o) are taken off the dependent chain: P_w[o] is gathered once
// per 32-token batch (one lane per token) and, for narrow rows, this lane's entry of the top search
// level is requested at the start of the token step, before the bucket is known. A draw that lands
// in the prior bucket then pays one dependent memory access per remaining level only.
__device__ __forceinline__ float prior_top_entry(const SweepParams& p, const WarpCtx& c, int w) {
  const float* prow = p.prior + (size_t)w * p.layout.stride;
  return (c.top_lane >= 0) ? __ldg(prow + c.top_lane) : 0.0f;
}
__device__ __forceinline__ int prior_search(const SweepParams& p, int lane, int K, int w, float po, float y, float delta,
                                            float vtop) {
  const float* prow = p.prior + (size_t)w * p.layout.stride;
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
    if (lane < nvalid) v = __ldg(prow + p.layout.off[lev] + lo + lane);
    const unsigned b = __ballot_sync(kFullMask, (lane < nvalid) && (v > s));
    block = lo + (b ? (__ffs(b) - 1) : (nvalid - 1));
  }
  if (top >= 1) {  // level 0: the K prefix sums themselves, at offset 0 of the word's block
    const int lo = block << 5;
    const int nvalid = min(32, K - lo);
    float v = 0.0f;
    if (lane < nvalid) v = __ldg(prow + lo + lane);
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

// ---- register path: rows that fit NT tiles with room for one more slot (nnz + 1 <= 32 NT) -------
// The warp's row in shared memory is ZERO-PADDED past nnz up to 32 NT slots (kRowPad spare slots
// behind slot_cap): a padded slot reads topic 0 / count 0, weighs exactly +0 and can never be the
// old topic's slot, so loads, weights and the bucket search carry no per-lane validity predicate.
// Prefixes stay in registers; the row edit happens in shared memory.
template <int NT, int MODE, bool LIVE, bool TS, int TE>
__device__ __forceinline__ int token_step_tiles(const SweepParams& p, WarpCtx& c, int& nnz, int w, int o, float u,
                                                float qw, float po_l, int t) {
  const int lane = c.lane;
  const int32_t* nrow = p.nwk_read + (size_t)w * c.K;
  constexpr bool kTopEarly = NT <= TE;  // request the prior's top level before the bucket is known
  float vtop = 0.0f;
  if (kTopEarly) vtop = prior_top_entry(p, c, w);
  uint32_t sv[NT];
  int nv[NT];
#pragma unroll
  for (int g = 0; g < NT; ++g) {
    sv[g] = smem_u32(c.slots + (g << 5) + lane);
    const int32_t* cell = nrow + (sv[g] >> 16);  // padded slots read n_wk[w, 0]: harmless, weight is 0
    nv[g] = LIVE ? live_load(cell) : __ldg(cell);
  }
  // a live slot of topic o reads (o << 16) + count with 1 <= count <= 0xffff
  const uint32_t okey = ((uint32_t)o << 16) + 1u;
  // Lane-strided prefix: the lane sums its own slots tile by tile (from +0), ONE warp scan runs over
  // the 32 lane totals, slot (g, lane) gets P = E_lane + s[g]. The cumulative order is lane-major.
  float s[NT];
  float run = 0.0f;
  int myjo = -1;
#pragma unroll
  for (int g = 0; g < NT; ++g) {
    const int topic = (int)(sv[g] >> 16);
    const bool is_old = (sv[g] - okey) < 0xffffu;
    if (is_old) myjo = (g << 5) + lane;
    const int cc = (int)(sv[g] & 0xffffu) - (int)is_old;
    const int n = max(nv[g] - ((int)is_old & c.excl), 0);
    const float inv = TS ? smem_f32(c.tab + topic) : __ldg(p.invden + topic);
    const float a = fmul(fmul(fadd((float)n, c.beta_f), inv), (float)cc);
    run = fadd(run, a);
    s[g] = run;
  }
  const float incl = warp_scan_inclusive(run, lane);
  const float E = shfl_up1_or_zero(incl);
  const float A = __shfl_sync(kFullMask, incl, 31);
  const int jo = __reduce_max_sync(kFullMask, myjo);
  float delta = TS ? smem_f32(c.tab + c.K + o) : __ldg(p.ab + o);
  delta = c.excl ? delta : 0.0f;
  float qp = fsub(qw, delta);
  qp = qp < 0.0f ? 0.0f : qp;
  const float x = fmul(u, fadd(A, qp));

  int newt;
  int jn = -1;  // slot of newt when it already has one
  if (x < A) {
    // First slot in cumulative (lane-major) order whose prefix exceeds x: the lowest lane with a
    // hit and, prefixes being non-decreasing inside a lane, its count of non-hits. A padded slot
    // repeats the prefix of the lane's slot before it, so it is never a lane's first hit unless the
    // lane holds no slot of the row at all (the j < nnz test). No hit (x within an ulp of A): last slot.
    int cnt = 0;
#pragma unroll
    for (int g = 0; g < NT; ++g) cnt += (fadd(E, s[g]) > x) ? 0 : 1;
    const int jc = (cnt << 5) + lane;
    const unsigned b = __ballot_sync(kFullMask, (cnt < NT) && (jc < nnz));
    jn = b ? __shfl_sync(kFullMask, jc, __ffs(b) - 1) : nnz - 1;
    newt = (int)(smem_u32(c.slots + jn) >> 16);  // shared memory still holds the row as loaded
  } else {
    ++c.st_prior;
    if (!kTopEarly) vtop = prior_top_entry(p, c, w);
    newt = prior_search(p, lane, c.K, w, __shfl_sync(kFullMask, po_l, t), fsub(x, A), delta, vtop);
  }

  if (MODE != MODE_FROZEN && newt != o) {
    ++c.st_moved;
    int pos = 0;  // #slots with topic < newt (old slot still present); only needed when newt is new to the row
    if (jn < 0) {
      // packed compares that a padded slot (0) fails: live slot of topic < newt, live slot of topic == newt
      const uint32_t nkey = (uint32_t)newt << 16;
      int less = 0, eqj = -1;
#pragma unroll
      for (int g = 0; g < NT; ++g) {
        less += ((sv[g] - 1u) < nkey) ? 1 : 0;
        if ((sv[g] - nkey - 1u) < 0xffffu) eqj = (g << 5) + lane;
      }
      pos = __reduce_add_sync(kFullMask, less);
      jn = __reduce_max_sync(kFullMask, eqj);
    }
    // the old slot's count decides whether the slot disappears
    const bool del = (smem_u32(c.slots + jo) & 0xffffu) == 1u;
    __syncwarp();  // every lane has read the row before any lane rewrites it
    // The edit happens in shared memory: two single-lane count updates, then (only when a slot
    // appears or disappears) a one-slot shift of [lo, lo + width] read as a whole before it is
    // written back. Measured against editing the register copy with shuffles: fewer instructions
    // per tile (range test + LDS + STS) and no second copy of the row in registers.
    if (lane == 0 && jn >= 0) smem_u32(c.slots + jn) += 1u;
    if (lane == 1 && !del) smem_u32(c.slots + jo) -= 1u;
    if (jn < 0 || del) {
      // destinations [lo, lo + width] take the slot at +off; ins gets the new slot (outside the range)
      int lo = 0x7fffffff, width = 0, off = 0, ins = -1;
      if (jn >= 0) {             // old slot empties, newt has one already: close the gap (slot nnz-1 takes the padding's 0)
        lo = jo; width = nnz - 1 - jo; off = 1;
      } else if (!del) {         // new slot, old one stays: open a gap at pos
        lo = pos + 1; width = nnz - pos - 1; off = -1; ins = pos;
      } else if (pos <= jo) {    // old slot empties, new one appears at or below it
        lo = pos + 1; width = jo - pos - 1; off = -1; ins = pos;
      } else {                   // ... or above it
        lo = jo; width = pos - 2 - jo; off = 1; ins = pos - 1;
      }
      if (width < 0) {  // empty range: a bare replacement of the old slot
        lo = 0x7fffffff;
        width = 0;
      }
      __syncwarp();  // count updates visible
      uint32_t mv[NT];
#pragma unroll
      for (int g = 0; g < NT; ++g) {
        const int j = (g << 5) + lane;
        mv[g] = 0u;
        if ((unsigned)(j - lo) <= (unsigned)width) mv[g] = smem_u32(c.slots + j + off);
      }
      __syncwarp();  // whole range read before any of it is overwritten
#pragma unroll
      for (int g = 0; g < NT; ++g) {
        const int j = (g << 5) + lane;
        if ((unsigned)(j - lo) <= (unsigned)width) smem_u32(c.slots + j) = mv[g];
      }
      if (lane == 0 && ins >= 0) smem_u32(c.slots + ins) = ((uint32_t)newt << 16) | 1u;
    }
    nnz += (jn < 0 ? 1 : 0) - (del ? 1 : 0);
    __syncwarp();
    count_moves(p, c, write_row<LIVE>(p, nrow, w, c.K), o, newt);
  }
  return newt;
}

// Rows [a, b) move one slot up (to [a+1, b+1)); chunks from the top so nothing is overwritten.
__device__ __forceinline__ void row_shift_up(uint32_t* slots, int a, int b, int lane) {
  for (int hi = b - 1; hi >= a; hi -= 32) {
    const int j = hi - lane;
    uint32_t v = 0u;
    if (j >= a) v = slots[j];
    __syncwarp();
    if (j >= a) slots[j + 1] = v;
    __syncwarp();
  }
}
// Rows [a, b) move one slot down (to [a-1, b-1)); chunks from the bottom.
__device__ __forceinline__ void row_shift_down(uint32_t* slots, int a, int b, int lane) {
  for (int lo = a; lo < b; lo += 32) {
    const int j = lo + lane;
    uint32_t v = 0u;
    if (j < b) v = slots[j];
    __syncwarp();
    if (j < b) slots[j - 1] = v;
    __syncwarp();
  }
}

// ---- generic path: rows of any width, loop form ----------------------------------------------------
template <int MODE, bool LIVE, bool TS>
__device__ __forceinline__ int token_step_generic(const SweepParams& p, WarpCtx& c, int& nnz, int w, int o, float u,
                                               float qw, float po_l, int t) {
  const int lane = c.lane;
  uint32_t* slots = &smem_u32(c.slots);
  float* pref = &smem_f32(c.pref);
  const int32_t* nrow = p.nwk_read + (size_t)w * c.K;
  // Tiles go in groups of kGroup: all of a group's n_wk gathers are issued before any is
  // consumed, so a wide row pays one memory latency per group, not per tile. Same lane-strided
  // prefix as the register path: pref[j] holds the lane-local inclusive sum s of slot j.
  const int ntiles = (nnz + 31) >> 5;
  const uint32_t okey = ((uint32_t)o << 16) + 1u;
  float run = 0.0f;
  int myjo = -1;
  for (int t0 = 0; t0 < ntiles; t0 += kGroup) {
    uint32_t sv[kGroup];
    int nv[kGroup];
#pragma unroll
    for (int g = 0; g < kGroup; ++g) {
      const int j = ((t0 + g) << 5) + lane;
      sv[g] = 0u;
      nv[g] = 0;
      if (j < nnz) {
        sv[g] = slots[j];
        const int32_t* cell = nrow + (sv[g] >> 16);
        nv[g] = LIVE ? live_load(cell) : __ldg(cell);
      }
    }
#pragma unroll
    for (int g = 0; g < kGroup; ++g) {
      const int j = ((t0 + g) << 5) + lane;
      if (j < nnz) {
        const int topic = (int)(sv[g] >> 16);
        const bool is_old = (sv[g] - okey) < 0xffffu;
        if (is_old) myjo = j;
        const int cc = (int)(sv[g] & 0xffffu) - (int)is_old;
        const int n = max(nv[g] - ((int)is_old & c.excl), 0);
        const float inv = TS ? smem_f32(c.tab + topic) : __ldg(p.invden + topic);
        const float a = fmul(fmul(fadd((float)n, c.beta_f), inv), (float)cc);
        run = fadd(run, a);
        pref[j] = run;
      }
    }
  }
  __syncwarp();
  const float incl = warp_scan_inclusive(run, lane);
  const float E = shfl_up1_or_zero(incl);
  const float A = __shfl_sync(kFullMask, incl, 31);
  const int jo = __reduce_max_sync(kFullMask, myjo);
  float delta = TS ? smem_f32(c.tab + c.K + o) : __ldg(p.ab + o);
  delta = c.excl ? delta : 0.0f;
  float qp = fsub(qw, delta);
  qp = qp < 0.0f ? 0.0f : qp;
  const float x = fmul(u, fadd(A, qp));

  int newt;
  int jn = -1;
  if (x < A) {
    jn = nnz - 1;
    // the lane's largest prefix is at its last slot of the row, where the local sum is `run`
    const bool hitlane = (lane < nnz) && (fadd(E, run) > x);
    const unsigned b = __ballot_sync(kFullMask, hitlane);
    if (b) {
      const int src = __ffs(b) - 1;
      int cand = 0;
      if (lane == src) {
        int j = lane;
        while (!(fadd(E, pref[j]) > x)) j += 32;  // ends at the lane's last slot at the latest
        cand = j;
      }
      jn = __shfl_sync(kFullMask, cand, src);
    }
    newt = (int)(slots[jn] >> 16);
  } else {
    ++c.st_prior;
    newt = prior_search(p, lane, c.K, w, __shfl_sync(kFullMask, po_l, t), fsub(x, A), delta, prior_top_entry(p, c, w));
  }

  if (MODE != MODE_FROZEN && newt != o) {
    ++c.st_moved;
    // Where does newt live (or go) in the row as it stands, old slot still present?
    int pos = 0;
    if (jn < 0) {
      for (int tile = 0; (tile << 5) < nnz; ++tile) {
        const int j = (tile << 5) + lane;
        const bool act = j < nnz;
        const int topic = act ? (int)(slots[j] >> 16) : 0x7fffffff;
        const unsigned less = __ballot_sync(kFullMask, topic < newt);
        const unsigned eq = __ballot_sync(kFullMask, topic == newt);
        pos += __popc(less);
        if (eq) jn = (tile << 5) + __ffs(eq) - 1;
        if (eq || less != kFullMask) break;
      }
    }
    const uint32_t so = slots[jo];
    const bool del = (so & 0xffffu) == 1u;
    __syncwarp();
    if (jn >= 0) {                 // newt already has a slot: bump it
      if (lane == 0) {
        slots[jn] += 1u;
        if (!del) slots[jo] = so - 1u;
      }
      if (del) {                   // ... and close the gap the old topic leaves
        __syncwarp();
        row_shift_down(slots, jo + 1, nnz, lane);
        --nnz;
        if (lane == 0) slots[nnz] = 0u;  // keep the zero padding behind the row
      }
    } else if (!del) {             // new slot, old one stays
      if (lane == 0) slots[jo] = so - 1u;
      __syncwarp();
      row_shift_up(slots, pos, nnz, lane);
      if (lane == 0) slots[pos] = ((uint32_t)newt << 16) | 1u;
      ++nnz;
    } else if (pos <= jo) {        // old slot empties, new one appears below it
      row_shift_up(slots, pos, jo, lane);
      if (lane == 0) slots[pos] = ((uint32_t)newt << 16) | 1u;
    } else {                       // ... or above it
      row_shift_down(slots, jo + 1, pos, lane);
      if (lane == 0) slots[pos - 1] = ((uint32_t)newt << 16) | 1u;
    }
    __syncwarp();
    count_moves(p, c, write_row<LIVE>(p, nrow, w, c.K), o, newt);
  }
  return newt;
}

// ROWCLASS selects which token-step variants a kernel instance carries, so that the register
// allocation (one per kernel) of the narrow-row classes is not dictated by the 8-tile variant:
//   0: rows <= 64 slots  (NT <= 3)     1: rows <= 128 slots (NT <= 5)     2: any row (NT <= 8 + loop form)
constexpr int kRowClasses = 3;
__host__ __device__ constexpr int rowclass_max_tiles(int rc) { return rc == 0 ? 3 : rc == 1 ? 5 : 8; }
__host__ __device__ constexpr int rowclass_min_ctas(int rc) {
  return rc == 0 ? B200LDA_ROWCLASS0_MIN_CTAS : rc == 1 ? B200LDA_ROWCLASS1_MIN_CTAS : B200LDA_SWEEP_MIN_CTAS;
}

template <int MODE, bool LIVE, bool TABLES_IN_SMEM, int ROWCLASS>
__global__ void __launch_bounds__(256, rowclass_min_ctas(ROWCLASS)) k_gibbs_sweep(const SweepParams p) {
  constexpr int MAXNT = rowclass_max_tiles(ROWCLASS);
  // Rows up to kTE tiles request the prior's top search level at the start of the token step. The
  // <= 128-slot classes have the register for it at any width; in the wide class it costs more
  // than it hides beyond 3 tiles (measured: C4 +1.3 % / C3 -2 % when applied everywhere).
  constexpr int kTE = ROWCLASS <= 1 ? 5 : B200LDA_TOP_EARLY_NT;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int K = p.K;

  // shared memory (words): [invden | ab | n_k delta] (3K, when TABLES_IN_SMEM), the warps' rows, the warps' prefixes
  const int tab_words = TABLES_IN_SMEM ? 3 * K : 0;
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
  c.excl = p.exclude_self;
  c.K = K;
  c.tab = 0;
  const int row_words = p.slot_cap + kRowPad;
  c.slots = tab_words + warp * row_words;
  c.pref = tab_words + nwarps * row_words + warp * p.slot_cap;
  c.nkd_in_smem = TABLES_IN_SMEM;
  c.top_lane = (lane < p.layout.size[p.layout.nlev - 1]) ? p.layout.off[p.layout.nlev - 1] + lane : -1;
  c.st_moved = 0;
  c.st_prior = 0;
  uint32_t* slots = &smem_u32(c.slots);

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
    if (lane == 0) ci = atomicAdd(p.doc_counter, 1ull);
    ci = __shfl_sync(kFullMask, ci, 0);
    if (ci >= nchunks) break;

    for (unsigned long long di = ci; di < ndocs; di += nchunks) {
      const int64_t d = (int64_t)__ldg(p.doc_order + p.order_begin + (int64_t)di);
      const int64_t tb = p.doc_ptr[d], te = p.doc_ptr[d + 1];
      if (te == tb) continue;
      const int64_t rp = p.row_ptr[d];
      int nnz = p.row_nnz[d];
      for (int j = lane; j < row_words; j += 32) slots[j] = j < nnz ? p.rows[rp + j] : 0u;  // zero-padded
      __syncwarp();
      unsigned doc_nnz = 0;

      const int iters = MODE == MODE_INFER ? p.infer_iters : 1;
      for (int it = 1; it <= iters; ++it) {
        const uint32_t sweep_key = MODE == MODE_INFER ? (uint32_t)it : p.sweep;
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
          const float q_l = valid ? __ldg(p.q + w_l) : 0.0f;
          const float po_l = valid ? __ldg(p.prior + (size_t)w_l * p.layout.stride + o_l) : 0.0f;  // P_w[o]
          int new_l = o_l;
          const int cnt = (int)min((int64_t)32, te - base);

          for (int t = 0; t < cnt; ++t) {
            const int w = __shfl_sync(kFullMask, w_l, t);
            const int o = __shfl_sync(kFullMask, o_l, t);
            const float u = __shfl_sync(kFullMask, u_l, t);
            const float qw = __shfl_sync(kFullMask, q_l, t);
            doc_nnz += (unsigned)nnz;
            int newt;
            const int tile_case = nnz >> 5;  // tiles needed for nnz + 1 slots, minus one (uniform across the warp)
            if (tile_case == 0) newt = token_step_tiles<1, MODE, LIVE, TABLES_IN_SMEM, kTE>(p, c, nnz, w, o, u, qw, po_l, t);
            else if (tile_case == 1) newt = token_step_tiles<2, MODE, LIVE, TABLES_IN_SMEM, kTE>(p, c, nnz, w, o, u, qw, po_l, t);
            else if (tile_case == 2 || MAXNT == 3) newt = token_step_tiles<3, MODE, LIVE, TABLES_IN_SMEM, kTE>(p, c, nnz, w, o, u, qw, po_l, t);
            else if (tile_case == 3) newt = token_step_tiles<(MAXNT >= 4 ? 4 : 1), MODE, LIVE, TABLES_IN_SMEM, kTE>(p, c, nnz, w, o, u, qw, po_l, t);
            else if (tile_case == 4 || MAXNT == 5) newt = token_step_tiles<(MAXNT >= 5 ? 5 : 1), MODE, LIVE, TABLES_IN_SMEM, kTE>(p, c, nnz, w, o, u, qw, po_l, t);
            else if (tile_case == 5) newt = token_step_tiles<(MAXNT >= 6 ? 6 : 1), MODE, LIVE, TABLES_IN_SMEM, kTE>(p, c, nnz, w, o, u, qw, po_l, t);
            else if (tile_case == 6) newt = token_step_tiles<(MAXNT >= 7 ? 7 : 1), MODE, LIVE, TABLES_IN_SMEM, kTE>(p, c, nnz, w, o, u, qw, po_l, t);
            else if (tile_case == 7) newt = token_step_tiles<(MAXNT >= 8 ? 8 : 1), MODE, LIVE, TABLES_IN_SMEM, kTE>(p, c, nnz, w, o, u, qw, po_l, t);
            else newt = token_step_generic<MODE, LIVE, TABLES_IN_SMEM>(p, c, nnz, w, o, u, qw, po_l, t);
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

        if (MODE == MODE_INFER) {
          st_nnz += doc_nnz;  // per iteration: 100 iterations of a long document overflow 32 bits
          doc_nnz = 0;
          const bool save = p.infer_samples == 0
                                ? it == iters
                                : (it > p.infer_burn_in && (it - p.infer_burn_in) % p.infer_thinning == 0);
          if (save) {  // this warp owns document d: plain read-modify-write
            __syncwarp();
            int32_t* acc = p.infer_acc + (size_t)d * K;
            for (int j = lane; j < nnz; j += 32) {
              const uint32_t sl = slots[j];
              acc[sl >> 16] += (int32_t)(sl & 0xffffu);
            }
          }
        }
      }

      if (MODE != MODE_FROZEN) {
        for (int j = lane; j < nnz; j += 32) p.rows[rp + j] = slots[j];
        if (lane == 0) p.row_nnz[d] = nnz;
      }
      st_nnz += doc_nnz;
      st_moved += c.st_moved;
      st_prior += c.st_prior;
      c.st_moved = 0;
      c.st_prior = 0;
      __syncwarp();
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
    if (st_moved) {
      atomicAdd(p.stats + 0, st_moved);
      atomicAdd(p.stats_cum + 0, st_moved);
    }
    if (st_prior) {
      atomicAdd(p.stats + 1, st_prior);
      atomicAdd(p.stats_cum + 1, st_prior);
    }
    if (st_nnz) {
      atomicAdd(p.stats + 2, st_nnz);
      atomicAdd(p.stats_cum + 2, st_nnz);
    }
  }
}

}  // namespace b200lda
