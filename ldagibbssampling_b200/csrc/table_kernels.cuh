// table_kernels.cuh — K2 (per-sweep tables) and K4 (AD-LDA delta / apply).
//
// K2 replaces the smoothingOnlyMass / cachedCoefficients set-up at the top of Mallet's
// WorkerRunnable.run (reached through estimate(), reference cmu_ron/TrainAndPredict.java:166):
// everything about the conditional that does not depend on the document is folded, once per
// sweep, into  invden_k = 1/(n_k + V beta),  ab_k = alpha_k * invden_k  and, per word, the
// inclusive prefix table of  (n_wk + beta) * ab_k  with its fan-out-32 search levels.
// K4 replaces ParallelTopicModel.sumTypeTopicCounts + the copy-back into every worker replica
// (setNumThreads(4), reference cmu_ron/TrainAndPredict.java:164, cmu/TrainAndPredict.java:262):
// each shard forms delta = counts_after - counts_before, the caller all-reduces it, and every
// shard applies the sum.
#pragma once
#include "device_common.cuh"

namespace b200lda {

__global__ void k_topic_tables(int K, const int32_t* __restrict__ nk, const float* __restrict__ alpha_f,
                               float vbeta, float* __restrict__ invden, float* __restrict__ ab) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float inv = __fdiv_rn(1.0f, fadd((float)nk[k], vbeta));
  invden[k] = inv;
  ab[k] = fmul(alpha_f[k], inv);
}

// One word's prior row: coalesced 128-byte reads of the n_wk row, tile scan, coalesced writes of
// the prefix row, then the upper search levels by sub-sampling the level below. Warp-collective.
// CG: the counts are live (other warps move them with atomics): read them at L2.
template <bool CG>
__device__ __forceinline__ float build_prior_row(int K, const int32_t* row, const float* __restrict__ ab, float beta_f,
                                                 const PriorLayout& L, float* out, int lane) {
  // Four tiles per round: their loads are in flight together and their scans are independent, so a
  // row costs K/128 memory latencies, not K/32 (the rows of rare words come from DRAM). The
  // arithmetic is the oracle's: Kogge-Stone inside a tile, tiles chained by carry + x.
  constexpr int G = 4;
  float carry = 0.0f;
  for (int base = 0; base < K; base += 32 * G) {
    float b[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int k = base + 32 * j + lane;
      b[j] = 0.0f;
      if (k < K) b[j] = fmul(fadd((float)(CG ? max(__ldcg(row + k), 0) : row[k]), beta_f), __ldg(ab + k));
    }
#pragma unroll
    for (int j = 0; j < G; ++j) b[j] = warp_scan_inclusive(b[j], lane);
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int k = base + 32 * j + lane;
      if (base + 32 * j < K) {  // uniform
        const float P = fadd(carry, b[j]);
        if (k < K) {
          if (CG) __stcs(out + L.off[0] + k, P); else out[L.off[0] + k] = P;
        }
        carry = __shfl_sync(kFullMask, P, 31);
      }
    }
  }
  __syncwarp();
  for (int lev = 1; lev < L.nlev; ++lev) {
    const float* lower = out + L.off[lev - 1];
    float* upper = out + L.off[lev];
    const int nl = L.size[lev - 1];
    for (int m = lane; m < L.size[lev]; m += 32) {
      const int src = min(32 * m + 31, nl - 1);
      upper[m] = CG ? __ldcg(lower + src) : lower[src];  // CG: never through L1 (the copy was rewritten by this warp)
    }
    __syncwarp();
  }
  return carry;  // Q_w = the row's last prefix
}

// One warp per word (grid-stride). sel (may be null): two copies of every row (rows 2 w and 2 w + 1),
// sel[w] = the current one; the build writes the other copy and flips (no sampling kernel is running).
__global__ void __launch_bounds__(256)
k_prior_rows(int V, int K, const int32_t* __restrict__ nwk, const float* __restrict__ ab, float beta_f,
             PriorLayout L, float* __restrict__ prior, float* __restrict__ q, int32_t* __restrict__ sel) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * warps_per_block;
  for (int64_t w = gw; w < V; w += nw) {
    const int copy = sel ? 1 - sel[w] : 0;
    const size_t rw = sel ? 2 * (size_t)w + (size_t)copy : (size_t)w;  // two interleaved copies per word, or one
    const float Q = build_prior_row<false>(K, nwk + (size_t)w * K, ab, beta_f, L, prior + rw * L.stride, lane);
    if (lane == 0) {
      q[rw] = Q;
      if (sel) sel[w] = copy;
    }
  }
}

// cum[0..3) += last[0..3)  (sweep statistics, accumulated until b200lda_reset_stats)
__global__ void k_accumulate_stats(const unsigned long long* __restrict__ last, unsigned long long* __restrict__ cum) {
  if (threadIdx.x < 3) cum[threadIdx.x] += last[threadIdx.x];
}

// Token count of every word (row sums of n_wk), one warp per word.
__global__ void k_word_counts(int V, int K, const int32_t* __restrict__ nwk, int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = gw; w < V; w += nw) {
    int acc = 0;
    for (int k = lane; k < K; k += 32) acc += nwk[(size_t)w * K + k];
    acc = __reduce_add_sync(kFullMask, acc);
    if (lane == 0) out[w] = acc;
  }
}
// hot[0 .. *n) = the words with at least `threshold` tokens (order unspecified)
__global__ void k_hot_words(int V, const int32_t* __restrict__ counts, int threshold, int32_t* __restrict__ hot,
                            int* __restrict__ n) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < V && counts[w] >= threshold) hot[atomicAdd(n, 1)] = w;
}

// acc[i] += add[i]   (in-process reduction of the shards' exchange buffers)
__global__ void k_add_i32(size_t n, int32_t* __restrict__ acc, const int32_t* __restrict__ add) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = n >> 2;
  int4* a4 = reinterpret_cast<int4*>(acc);
  const int4* b4 = reinterpret_cast<const int4*>(add);
  for (size_t i = tid; i < n4; i += stride) {
    int4 a = a4[i];
    const int4 b = b4[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    a4[i] = a;
  }
  for (size_t i = (n4 << 2) + tid; i < n; i += stride) acc[i] += add[i];
}

// nk += nk_delta; nk_delta = 0   (single-shard sweep finish)
__global__ void k_apply_nk(int K, int32_t* __restrict__ nk, int32_t* __restrict__ nk_delta) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  nk[k] += nk_delta[k];
  nk_delta[k] = 0;
}

// AD-LDA exchange, in place. Every shard starts a sweep with the same global counts G (kept in
// `snap`), samples, and holds A_r = G + (its own moves) in `sum` (plus its n_k moves in the K-cell
// tail behind the V x K cells). The caller all-reduces `sum` (tail included) over the N shards:
// sum = sum_r A_r. The new global counts are  sum - (N - 1) G ; they go to BOTH buffers (the next
// sweep's live / written copy and its snapshot). Replaces Mallet's sumTypeTopicCounts + copy-back
// (setNumThreads(n), reference cmu_ron/TrainAndPredict.java:164) without a delta buffer:
// two reads and two writes per cell around one collective. [i0, i1) = the slab of cells to apply
// (the library pipelines slabs behind the all-reduce of the next one).
__global__ void k_apply_sum(size_t i0, size_t i1, int nm1, int32_t* __restrict__ sum, int32_t* __restrict__ snap) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t a0 = (i0 + 3) & ~(size_t)3, a1 = i1 & ~(size_t)3;  // int4 body, scalar edges
  if (a0 < a1) {
    int4* s4 = reinterpret_cast<int4*>(sum);
    int4* g4 = reinterpret_cast<int4*>(snap);
    for (size_t i = (a0 >> 2) + tid; i < (a1 >> 2); i += stride) {
      const int4 s = s4[i], g = g4[i];
      const int4 n = make_int4(s.x - nm1 * g.x, s.y - nm1 * g.y, s.z - nm1 * g.z, s.w - nm1 * g.w);
      s4[i] = n;
      g4[i] = n;
    }
    for (size_t i = i0 + tid; i < a0; i += stride) sum[i] = snap[i] = sum[i] - nm1 * snap[i];
    for (size_t i = a1 + tid; i < i1; i += stride) sum[i] = snap[i] = sum[i] - nm1 * snap[i];
  } else {
    for (size_t i = i0 + tid; i < i1; i += stride) sum[i] = snap[i] = sum[i] - nm1 * snap[i];
  }
}
// nk[k] += tail[k] (the summed n_k moves); both buffers' tails are cleared for the next sweep.
__global__ void k_apply_nk_tail(int K, int32_t* __restrict__ nk, int32_t* __restrict__ tail_sum, int32_t* __restrict__ tail_snap) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  nk[k] += tail_sum[k];
  tail_sum[k] = 0;
  tail_snap[k] = 0;
}
// start-up: nk[k] = tail[k] (the summed per-shard n_k), tails cleared
__global__ void k_install_nk_tail(int K, int32_t* __restrict__ nk, int32_t* __restrict__ tail_sum, int32_t* __restrict__ tail_snap) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  nk[k] = tail_sum[k];
  tail_sum[k] = 0;
  tail_snap[k] = 0;
}

}  // namespace b200lda
