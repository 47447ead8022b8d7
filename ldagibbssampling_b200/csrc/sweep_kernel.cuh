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
//   * count moves: 16-bit row edit in shared memory, integer RED atomics on n_wk / n_k deltas.
// MODE_UPDATE serves both LIVE (nwk_read == nwk_write) and DEFERRED (distinct buffers);
// MODE_FROZEN moves nothing (north-star parity mode).
#pragma once
#include "device_common.cuh"

namespace b200lda {

enum { MODE_UPDATE = 0, MODE_FROZEN = 1 };

struct SweepParams {
  int64_t num_docs;
  const int64_t* doc_ptr;     // [D+1] token offsets (document order)
  const int32_t* tok_word;    // [N]
  uint16_t* z;                // [N] current topics (updated in MODE_UPDATE)
  int32_t* z_out;             // [N] MODE_FROZEN output
  const int64_t* row_ptr;     // [D+1] offsets into rows (capacity min(K, L_d) per doc)
  int32_t* row_nnz;           // [D]
  uint32_t* rows;             // packed (topic << 16 | count), ascending topic
  const int32_t* nwk_read;    // [V*K]
  int32_t* nwk_write;         // [V*K] (== nwk_read in LIVE mode)
  int32_t* nk_delta;          // [K]
  const float* invden;        // [K]  1 / (n_k + V beta)
  const float* ab;            // [K]  alpha_k * invden_k
  const float* prior;         // [V * layout.stride]
  const float* q;             // [V]  prior bucket mass per word
  const float* uniforms;      // [N] or nullptr (Philox)
  PriorLayout layout;
  int K;
  int slot_cap;               // shared-memory slots per warp (multiple of 32)
  float beta_f;
  uint64_t seed;
  uint32_t sweep;
  int64_t global_tok_off;
  unsigned long long* doc_counter;  // dynamic document scheduler
  unsigned long long* stats;        // [0] moved, [1] prior-bucket draws, [2] sum of nnz over tokens
  unsigned long long* stats_cum;    // same three, accumulated until b200lda_reset_stats
};

constexpr int kDocChunk = 4;

template <int MODE, bool TABLES_IN_SMEM>
__global__ void __launch_bounds__(256) k_gibbs_sweep(const SweepParams p) {
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
  const float* t_invden;
  const float* t_ab;
  if (TABLES_IN_SMEM) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      s_tab[k] = p.invden[k];
      s_tab[K + k] = p.ab[k];
    }
    __syncthreads();
    t_invden = s_tab;
    t_ab = s_tab + K;
  } else {
    t_invden = p.invden;
    t_ab = p.ab;
  }

  const bool live = (p.nwk_read == p.nwk_write);
  const float beta_f = p.beta_f;
  unsigned long long st_moved = 0, st_prior = 0, st_nnz = 0;

  for (;;) {
    unsigned long long d0 = 0;
    if (lane == 0) d0 = atomicAdd(p.doc_counter, (unsigned long long)kDocChunk);
    d0 = __shfl_sync(kFullMask, d0, 0);
    if ((int64_t)d0 >= p.num_docs) break;
    const int64_t d1 = min((int64_t)d0 + kDocChunk, p.num_docs);

    for (int64_t d = (int64_t)d0; d < d1; ++d) {
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
          const int ntiles = (nnz + 31) >> 5;
          float carry = 0.0f, P = 0.0f;
          int jo = -1;
          for (int tile = 0; tile < ntiles; ++tile) {
            const int j = (tile << 5) + lane;
            const bool act = j < nnz;
            const uint32_t s = act ? slots[j] : 0u;
            const int topic = (int)(s >> 16);
            int c = (int)(s & 0xffffu);
            int n = 0;
            if (act) n = live ? __ldcg(nrow + topic) : __ldg(nrow + topic);
            const bool is_old = act && (topic == o);
            const unsigned bo = __ballot_sync(kFullMask, is_old);
            if (bo) jo = (tile << 5) + __ffs(bo) - 1;
            if (is_old) {
              c -= 1;
              n = max(n - 1, 0);
            }
            float a = 0.0f;
            if (act) a = fmul(fmul(fadd((float)n, beta_f), t_invden[topic]), (float)c);
            a = warp_scan_inclusive(a, lane);
            P = fadd(carry, a);
            if (ntiles > 1 && act) pref[j] = P;
            carry = __shfl_sync(kFullMask, P, 31);
          }
          const float A = __shfl_sync(kFullMask, P, (nnz - 1) & 31);
          const float delta = t_ab[o];
          float qp = fsub(qw, delta);
          qp = qp < 0.0f ? 0.0f : qp;
          const float T = fadd(A, qp);
          const float x = fmul(u, T);

          int newt;
          if (x < A) {
            int jj = nnz - 1;
            if (ntiles == 1) {
              const unsigned b = __ballot_sync(kFullMask, (lane < nnz) && (P > x));
              if (b) jj = __ffs(b) - 1;
            } else {
              __syncwarp();
              for (int tile = 0; tile < ntiles; ++tile) {
                const int j = (tile << 5) + lane;
                const bool hit = (j < nnz) && (pref[j] > x);
                const unsigned b = __ballot_sync(kFullMask, hit);
                if (b) {
                  jj = (tile << 5) + __ffs(b) - 1;
                  break;
                }
              }
            }
            newt = (int)(slots[jj] >> 16);
          } else {
            // ---- prior bucket: skip the own-token mass delta at topic o, then search ----------
            ++st_prior;
            const float y = fsub(x, A);
            const float* prow = p.prior + (size_t)w * p.layout.stride;
            const float po = __ldg(prow + p.layout.off[0] + o);
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
            // 1) take the token out of topic o (delete the slot if it empties)
            const uint32_t so = slots[jo];
            __syncwarp();
            if ((so & 0xffffu) == 1u) {
              for (int lo = jo; lo < nnz - 1; lo += 32) {
                const int j = lo + lane;
                uint32_t v = 0u;
                if (j < nnz - 1) v = slots[j + 1];
                __syncwarp();
                if (j < nnz - 1) slots[j] = v;
                __syncwarp();
              }
              --nnz;
            } else if (lane == 0) {
              slots[jo] = so - 1u;
            }
            __syncwarp();
            // 2) put it into topic newt (insert a slot, keeping ascending topic order)
            int pos = 0, found = -1;
            for (int tile = 0; (tile << 5) < nnz; ++tile) {
              const int j = (tile << 5) + lane;
              const bool act = j < nnz;
              const int topic = act ? (int)(slots[j] >> 16) : 0x7fffffff;
              const unsigned less = __ballot_sync(kFullMask, act && topic < newt);
              const unsigned eq = __ballot_sync(kFullMask, act && topic == newt);
              pos += __popc(less);
              if (eq) found = (tile << 5) + __ffs(eq) - 1;
              if (eq || less != kFullMask) break;
            }
            if (found >= 0) {
              if (lane == 0) slots[found] += 1u;
            } else {
              for (int hi = nnz - 1; hi >= pos; hi -= 32) {
                const int j = hi - lane;
                uint32_t v = 0u;
                if (j >= pos) v = slots[j];
                __syncwarp();
                if (j >= pos) slots[j + 1] = v;
                __syncwarp();
              }
              if (lane == 0) slots[pos] = ((uint32_t)newt << 16) | 1u;
              ++nnz;
            }
            __syncwarp();
            // 3) word-topic and topic totals: integer RED atomics (order-independent sums)
            if (lane == 0) {
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
