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
//     resolved by a fan-out-32 search = one coalesced 128-byte line per level;
//   * randomness: Philox keyed by (seed; global token, sweep), 32 tokens per warp batch, one
//     lane each, so the RNG costs ~2 instructions per token;
//   * count moves: one pass over the part of the 16-bit row between the slot that empties and
//     the slot that appears, integer RED atomics on n_wk / n_k deltas.
// Documents are visited through doc_order (longest first); the host launches the kernel once per
// document class so that short documents get small per-warp rows and therefore high occupancy.
// MODE_UPDATE serves LIVE (LIVE=true: n_wk read through L2 and written in place) and DEFERRED
// (LIVE=false: frozen n_wk through the read-only path, moves go to a second buffer);
// MODE_FROZEN moves nothing (north-star parity mode).
#pragma once
#include "device_common.cuh"

namespace b200lda {

enum { MODE_UPDATE = 0, MODE_FROZEN = 1 };

struct SweepParams {
  const int32_t* doc_order;   // [D] document ids, class by class, longest first
  int64_t order_begin, order_end;  // this launch's slice of doc_order
  const int64_t* doc_ptr;     // [D+1] token offsets (document order)
  const int32_t* tok_word;    // [N]
  uint16_t* z;                // [N] current topics (updated in MODE_UPDATE)
  int32_t* z_out;             // [N] MODE_FROZEN output
  const int64_t* row_ptr;     // [D+1] offsets into rows (capacity min(K, L_d) per doc)
  int32_t* row_nnz;           // [D]
  uint32_t* rows;             // packed (topic << 16 | count), ascending topic
  const int32_t* nwk_read;    // [V*K]
  int32_t* nwk_write;         // [V*K] (== nwk_read when LIVE)
  int32_t* nk_delta;          // [K]
  const float* invden;        // [K]  1 / (n_k + V beta)
  const float* ab;            // [K]  alpha_k * invden_k
  const float* prior;         // [V * layout.stride]
  const float* q;             // [V]  prior bucket mass per word
  const float* uniforms;      // [N] or nullptr (Philox)
  PriorLayout layout;
  int K;
  int slot_cap;               // shared-memory slots per warp (multiple of 32)
  int doc_chunk;              // documents fetched per scheduler atomic
  int exclude_self;           // 1: the token being resampled is counted in n_wk (training);
                              // 0: held-out inference against frozen counts (TopicInferencer)
  float beta_f;
  uint64_t seed;
  uint32_t sweep;
  int64_t global_tok_off;
  unsigned long long* doc_counter;  // dynamic document scheduler (starts at 0 for each launch)
  unsigned long long* stats;        // [0] moved, [1] prior-bucket draws, [2] sum of nnz over tokens
  unsigned long long* stats_cum;    // same three, accumulated until b200lda_reset_stats
};

#ifndef B200LDA_SWEEP_GROUP
#define B200LDA_SWEEP_GROUP 4  // measured on B200: 4 > 3 > 2 > 1 (profiles/r01_tuning.md)
#endif
constexpr int kGroup = B200LDA_SWEEP_GROUP;  // tiles whose gathers are in flight together

// Shared memory per warp: slot_cap x {uint32 row slot, float prefix} = 8 bytes per slot.
constexpr int kSmemBytesPerSlot = 8;

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

#ifndef B200LDA_SWEEP_MIN_CTAS
#define B200LDA_SWEEP_MIN_CTAS 4   // 8-warp CTAs per SM the register allocation must allow
#endif

template <int MODE, bool LIVE, bool TABLES_IN_SMEM>
__global__ void __launch_bounds__(256, B200LDA_SWEEP_MIN_CTAS) k_gibbs_sweep(const SweepParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int K = p.K;

  float* s_tab = reinterpret_cast<float*>(smem_raw);
  const int tab_floats = TABLES_IN_SMEM ? 2 * K : 0;
  uint32_t* slots = reinterpret_cast<uint32_t*>(s_tab + tab_floats) + (size_t)warp * p.slot_cap;
  float* pref = reinterpret_cast<float*>(reinterpret_cast<uint32_t*>(s_tab + tab_floats) +
                                         (size_t)nwarps * p.slot_cap) + (size_t)warp * p.slot_cap;
  if (TABLES_IN_SMEM) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      s_tab[k] = p.invden[k];
      s_tab[K + k] = p.ab[k];
    }
    __syncthreads();
  }
  // fp32 masks that make the Kogge-Stone step "if (lane >= d) v += y" one FFMA: v = y*m + v is
  // exactly v + y (m = 1, one rounding) or exactly v (m = 0).
  const float m1 = lane >= 1 ? 1.0f : 0.0f, m2 = lane >= 2 ? 1.0f : 0.0f, m4 = lane >= 4 ? 1.0f : 0.0f,
              m8 = lane >= 8 ? 1.0f : 0.0f, m16 = lane >= 16 ? 1.0f : 0.0f;

  const float beta_f = p.beta_f;
  unsigned long long st_moved = 0, st_prior = 0, st_nnz = 0;
  const unsigned long long ndocs = (unsigned long long)(p.order_end - p.order_begin);

  for (;;) {
    unsigned long long c0 = 0;
    if (lane == 0) c0 = atomicAdd(p.doc_counter, (unsigned long long)p.doc_chunk);
    c0 = __shfl_sync(kFullMask, c0, 0);
    if (c0 >= ndocs) break;
    const unsigned long long c1 = min(c0 + (unsigned long long)p.doc_chunk, ndocs);

    for (unsigned long long ci = c0; ci < c1; ++ci) {
      const int64_t d = (int64_t)__ldg(p.doc_order + p.order_begin + (int64_t)ci);
      const int64_t tb = p.doc_ptr[d], te = p.doc_ptr[d + 1];
      if (te == tb) continue;
      const int64_t rp = p.row_ptr[d];
      int nnz = p.row_nnz[d];
      for (int j = lane; j < nnz; j += 32) slots[j] = p.rows[rp + j];
      __syncwarp();

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
            u_l = u24(token_random(p.seed, (uint64_t)(p.global_tok_off + i), p.sweep, 0u).x);
          }
        }
        const float q_l = valid ? __ldg(p.q + w_l) : 0.0f;
        int new_l = o_l;
        const int cnt = (int)min((int64_t)32, te - base);

        for (int t = 0; t < cnt; ++t) {
          const int w = __shfl_sync(kFullMask, w_l, t);
          const int o = __shfl_sync(kFullMask, o_l, t);
          const float u = __shfl_sync(kFullMask, u_l, t);
          const float qw = __shfl_sync(kFullMask, q_l, t);
          const int32_t* nrow = p.nwk_read + (size_t)w * K;
          st_nnz += (unsigned)nnz;

          // ---- doc bucket: weights + tile scan --------------------------------------------
          // Tiles go in groups of kGroup: all of a group's n_wk gathers are issued before any
          // is consumed, so a multi-tile row pays one memory latency per group, not per tile.
          const int ntiles = (nnz + 31) >> 5;
          float carry = 0.0f, P = 0.0f;
          int jo = 0;
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
                nv[g] = LIVE ? __ldcg(cell) : __ldg(cell);
              }
            }
#pragma unroll
            for (int g = 0; g < kGroup; ++g) {
              if (t0 + g < ntiles) {
                const int j = ((t0 + g) << 5) + lane;
                const bool act = j < nnz;
                const int topic = (int)(sv[g] >> 16);
                const bool is_old = act && (topic == o);
                const unsigned bo = __ballot_sync(kFullMask, is_old);
                if (bo) jo = ((t0 + g) << 5) + __ffs(bo) - 1;
                const int c = (int)(sv[g] & 0xffffu) - (int)is_old;
                const int n = max(nv[g] - ((int)is_old & p.exclude_self), 0);
                const float inv = TABLES_IN_SMEM ? s_tab[topic] : __ldg(p.invden + topic);
                float a = fmul(fmul(fadd((float)n, beta_f), inv), (float)c);  // c == 0 past nnz
                a = __fmaf_rn(__shfl_up_sync(kFullMask, a, 1), m1, a);
                a = __fmaf_rn(__shfl_up_sync(kFullMask, a, 2), m2, a);
                a = __fmaf_rn(__shfl_up_sync(kFullMask, a, 4), m4, a);
                a = __fmaf_rn(__shfl_up_sync(kFullMask, a, 8), m8, a);
                a = __fmaf_rn(__shfl_up_sync(kFullMask, a, 16), m16, a);
                P = fadd(carry, a);
                if (ntiles > 1) pref[j] = P;
                carry = __shfl_sync(kFullMask, P, 31);
              }
            }
          }
          const float A = __shfl_sync(kFullMask, P, (nnz - 1) & 31);
          float delta = TABLES_IN_SMEM ? s_tab[K + o] : __ldg(p.ab + o);
          delta = p.exclude_self ? delta : 0.0f;
          float qp = fsub(qw, delta);
          qp = qp < 0.0f ? 0.0f : qp;
          const float T = fadd(A, qp);
          const float x = fmul(u, T);

          int newt;
          int jn = -1;  // slot of newt when it is already known to be in the row
          if (x < A) {
            jn = nnz - 1;
            if (ntiles == 1) {
              const unsigned b = __ballot_sync(kFullMask, (lane < nnz) && (P > x));
              if (b) jn = __ffs(b) - 1;
            } else {
              __syncwarp();
              for (int tile = 0; tile < ntiles; ++tile) {
                const int j = (tile << 5) + lane;
                const bool hit = (j < nnz) && (pref[j] > x);
                const unsigned b = __ballot_sync(kFullMask, hit);
                if (b) {
                  jn = (tile << 5) + __ffs(b) - 1;
                  break;
                }
              }
            }
            newt = (int)(slots[jn] >> 16);
          } else {
            // ---- prior bucket: skip the own-token mass delta at topic o, then search ----------
            ++st_prior;
            const float y = fsub(x, A);
            const float* prow = p.prior + (size_t)w * p.layout.stride;
            const float po = __ldg(prow + o);  // level 0 sits at offset 0
            const float pod = fsub(po, delta);
            const float s = (y < pod) ? y : fadd(y, delta);
            int block = 0;
            for (int lev = p.layout.nlev - 1; lev >= 0; --lev) {
              const int lo = block << 5;
              const int nvalid = min(32, p.layout.size[lev] - lo);
              float v = 0.0f;
              if (lane < nvalid) v = __ldg(prow + p.layout.off[lev] + lo + lane);
              const unsigned b = __ballot_sync(kFullMask, (lane < nvalid) && (v > s));
              block = lo + (b ? (__ffs(b) - 1) : (nvalid - 1));
            }
            newt = block;
          }

          if (MODE == MODE_UPDATE && newt != o) {
            ++st_moved;
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
            // word-topic and topic totals: integer RED atomics (order-independent sums)
            if (lane == 0 && p.nwk_write != nullptr) {
              int32_t* wrow = p.nwk_write + (size_t)w * K;
              atomicAdd(wrow + o, -1);
              atomicAdd(wrow + newt, 1);
              atomicAdd(p.nk_delta + o, -1);
              atomicAdd(p.nk_delta + newt, 1);
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

      if (MODE == MODE_UPDATE) {
        for (int j = lane; j < nnz; j += 32) p.rows[rp + j] = slots[j];
        if (lane == 0) p.row_nnz[d] = nnz;
      }
      __syncwarp();
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
