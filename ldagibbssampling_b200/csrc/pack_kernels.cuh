// pack_kernels.cuh — K1: corpus packing and count construction.
//
// Replaces ParallelTopicModel.addInstances + buildInitialTypeTopicCounts (reference call sites
// cmu_ron/TrainAndPredict.java:162,174 and cmu/TrainAndPredict.java:260,271):
// uniform random initial topics, n_wk / n_k histograms and the sparse per-document topic rows.
// Mallet keeps int[L_d] topics + packed int rows on the Java heap; here z is uint16, n_wk a dense
// int32 V x K matrix (row = word, so one word's topics are contiguous for the sampler's gathers)
// and n_dk one packed (topic<<16 | count) row per document with ascending topics.
#pragma once
#include "device_common.cuh"

namespace b200lda {

// z[i] = floor(K * x / 2^32), x = Philox(seed; global token, sweep 0, stream 1).x
__global__ void k_init_z(int64_t N, int K, uint64_t seed, int64_t global_off, uint16_t* __restrict__ z) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const uint32_t x = token_random(seed, (uint64_t)(global_off + i), 0u, 1u).x;
    z[i] = (uint16_t)(((uint64_t)x * (uint64_t)K) >> 32);
  }
}

// int32 host layout -> uint16 device layout; flags out-of-range topics.
__global__ void k_narrow_z(int64_t N, int K, const int32_t* __restrict__ in, uint16_t* __restrict__ out,
                           int* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const int32_t v = in[i];
    if (v < 0 || v >= K) {
      *bad = 1;
      out[i] = 0;
    } else {
      out[i] = (uint16_t)v;
    }
  }
}

// Order-dependent checksum of an array: sum over i of mix(i, x[i]) (commutative, so a parallel
// reduction gives one well-defined value). Guards a state blob against the wrong corpus.
__device__ __forceinline__ unsigned long long hash_mix(unsigned long long i, unsigned long long x) {
  unsigned long long z = (i + 1ull) * 0x9E3779B97F4A7C15ull ^ (x + 0xD1B54A32D192ED03ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <typename T>
__device__ __forceinline__ void hash_array(int64_t n, const T* __restrict__ x, unsigned long long* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    acc += hash_mix((unsigned long long)i, (unsigned long long)(long long)x[i]);
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}
__global__ void k_hash_i64(int64_t n, const int64_t* __restrict__ x, unsigned long long* __restrict__ out) { hash_array(n, x, out); }
__global__ void k_hash_i32(int64_t n, const int32_t* __restrict__ x, unsigned long long* __restrict__ out) { hash_array(n, x, out); }

// Validation of topics handed over in their device width (16 bits): flags any z >= K (and clamps
// it to 0 so that nothing downstream reads out of bounds before the error is reported).
__global__ void k_validate_z(int64_t N, int K, uint16_t* __restrict__ z, int* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool any = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    if ((int)z[i] >= K) {
      any = true;
      z[i] = 0;
    }
  }
  if (any) *bad = 1;
}

__global__ void k_widen_z(int64_t N, const uint16_t* __restrict__ in, int32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) out[i] = (int32_t)in[i];
}

// Streaming validation of word ids (no atomics): flags any id outside [0, V).
__global__ void k_validate_words(int64_t N, int V, const int32_t* __restrict__ tok_word, int* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool any = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const int32_t w = tok_word[i];
    any |= (w < 0) | (w >= V);
  }
  if (any) *bad = 1;
}

// Document plan, pass 1: validates the CSR (monotone, length <= 65535; *bad = 1 + first bad doc)
// and histograms document lengths (65536 + 1 bins).
__global__ void k_doc_lengths(int64_t D, const int64_t* __restrict__ doc_ptr, unsigned long long* __restrict__ len_hist,
                              long long* __restrict__ bad_doc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += stride) {
    const int64_t len = doc_ptr[d + 1] - doc_ptr[d];
    if (len < 0 || len > 65535) {
      atomicMin(bad_doc, (long long)d);
    } else {
      atomicAdd(len_hist + len, 1ull);
    }
  }
}

// Document plan, pass 2 (two-level exclusive scan of row capacities min(len, K) -> row_ptr):
// block sums, then (after a single-block scan of the sums) per-block rescan + offset.
constexpr int kScanBlock = 1024;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* s_warp, long long* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long y = __shfl_up_sync(kFullMask, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    long long t = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long y = __shfl_up_sync(kFullMask, t, d);
      if (lane >= d) t += y;
    }
    s_warp[lane] = t;
  }
  __syncthreads();
  const long long off = warp ? s_warp[warp - 1] : 0;
  *total = s_warp[(blockDim.x >> 5) - 1];
  return off + x - v;
}

__global__ void __launch_bounds__(kScanBlock)
k_row_block_sums(int64_t D, int K, const int64_t* __restrict__ doc_ptr, unsigned long long* __restrict__ block_sum) {
  __shared__ long long s_warp[32];
  const int64_t d = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  long long v = 0;
  if (d < D) v = min((long long)(doc_ptr[d + 1] - doc_ptr[d]), (long long)K);
  long long total;
  block_exclusive_scan(v, s_warp, &total);
  if (threadIdx.x == 0) block_sum[blockIdx.x] = (unsigned long long)total;
}

__global__ void __launch_bounds__(kScanBlock)
k_row_ptr_apply(int64_t D, int K, const int64_t* __restrict__ doc_ptr, const long long* __restrict__ block_off,
                int64_t* __restrict__ row_ptr) {
  __shared__ long long s_warp[32];
  const int64_t d = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  long long v = 0;
  if (d < D) v = min((long long)(doc_ptr[d + 1] - doc_ptr[d]), (long long)K);
  long long total;
  const long long ex = block_exclusive_scan(v, s_warp, &total);
  const long long base = block_off[blockIdx.x];
  if (d < D) row_ptr[d] = base + ex;
  if (d == D - 1) row_ptr[D] = base + ex + v;
}

// b200lda_get_ndk_csr: compact row offsets = exclusive scan of the rows' nnz (same two-level scan),
// then the packed (topic << 16 | count) slots unpacked into the caller's two arrays.
__global__ void __launch_bounds__(kScanBlock)
k_nnz_block_sums(int64_t D, const int32_t* __restrict__ nnz, unsigned long long* __restrict__ block_sum) {
  __shared__ long long s_warp[32];
  const int64_t d = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  long long total;
  block_exclusive_scan(d < D ? (long long)nnz[d] : 0, s_warp, &total);
  if (threadIdx.x == 0) block_sum[blockIdx.x] = (unsigned long long)total;
}
__global__ void __launch_bounds__(kScanBlock)
k_nnz_ptr_apply(int64_t D, const int32_t* __restrict__ nnz, const long long* __restrict__ block_off,
                int64_t* __restrict__ out_ptr) {
  __shared__ long long s_warp[32];
  const int64_t d = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const long long v = d < D ? (long long)nnz[d] : 0;
  long long total;
  const long long ex = block_exclusive_scan(v, s_warp, &total);
  const long long base = block_off[blockIdx.x];
  if (d < D) out_ptr[d] = base + ex;
  if (d == D - 1) out_ptr[D] = base + ex + v;
}
__global__ void k_unpack_rows(int64_t D, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ nnz,
                              const uint32_t* __restrict__ rows, const int64_t* __restrict__ out_ptr,
                              int32_t* __restrict__ topic, int32_t* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t d = gw; d < D; d += nw) {
    const int64_t rp = row_ptr[d], op = out_ptr[d];
    const int n = nnz[d];
    for (int j = lane; j < n; j += 32) {
      const uint32_t v = rows[rp + j];
      topic[op + j] = (int32_t)(v >> 16);
      count[op + j] = (int32_t)(v & 0xffffu);
    }
  }
}

// Document plan, pass 3: visiting order = longest document first. len_start[L] = number of
// documents longer than L (from the length histogram); ties are ordered by arrival.
__global__ void k_doc_order_scatter(int64_t D, const int64_t* __restrict__ doc_ptr, const long long* __restrict__ len_start,
                                    unsigned long long* __restrict__ cursor, int32_t* __restrict__ doc_order) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += stride) {
    const int64_t len = doc_ptr[d + 1] - doc_ptr[d];
    const unsigned long long r = atomicAdd(cursor + len, 1ull);
    doc_order[len_start[len] + (long long)r] = (int32_t)d;
  }
}

// Validates word ids and histograms them (word -> token CSR, pass 1).
__global__ void k_word_hist(int64_t N, int V, const int32_t* __restrict__ tok_word,
                            unsigned long long* __restrict__ word_count, int* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const int32_t w = tok_word[i];
    if (w < 0 || w >= V) {
      *bad = 1;
    } else {
      atomicAdd(word_count + w, 1ull);
    }
  }
}

// Single-block exclusive scan of word_count into word_ptr[V+1] (V is at most a few million).
__global__ void k_exclusive_scan_u64(int n, const unsigned long long* __restrict__ in,
                                     long long* __restrict__ out) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    unsigned long long v = (i < n) ? in[i] : 0ull;
    unsigned long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long y = __shfl_up_sync(kFullMask, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned long long t = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0ull;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(kFullMask, t, d);
        if (lane >= d) t += y;
      }
      s_warp[lane] = t;
    }
    __syncthreads();
    const unsigned long long warp_off = warp ? s_warp[warp - 1] : 0ull;
    const unsigned long long carry = s_carry;
    if (i < n) out[i] = (long long)(carry + warp_off + x - v);
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + warp_off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = (long long)s_carry;
}

// word -> token CSR, pass 2: wtok[word_ptr[w] + r] = token index; order inside a word is by
// cursor arrival (any order is a valid CSR; the counts built from it are order-independent).
__global__ void k_word_scatter(int64_t N, const int32_t* __restrict__ tok_word, const long long* __restrict__ word_ptr,
                               unsigned long long* __restrict__ cursor, int64_t* __restrict__ wtok) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const int32_t w = tok_word[i];
    const unsigned long long r = atomicAdd(cursor + w, 1ull);
    wtok[word_ptr[w] + (long long)r] = i;
  }
}

// n_wk straight from the doc -> token order (no word order needed): streaming reads, random REDs.
__global__ void __launch_bounds__(256)
k_count_direct(int64_t N, int K, const int32_t* __restrict__ tok_word, const uint16_t* __restrict__ z,
               int32_t* __restrict__ nwk) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride)
    atomicAdd(nwk + (size_t)tok_word[i] * K + z[i], 1);
}

// n_k = column sums of n_wk: each block reduces a slab of rows, one thread per topic column.
__global__ void __launch_bounds__(256)
k_col_sums(int V, int K, int rows_per_block, const int32_t* __restrict__ nwk, int32_t* __restrict__ nk) {
  const int w0 = blockIdx.y * rows_per_block;
  const int w1 = min(V, w0 + rows_per_block);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
    int acc = 0;
    for (int w = w0; w < w1; ++w) acc += nwk[(size_t)w * K + k];
    if (acc) atomicAdd(nk + k, acc);
  }
}

// Sparse document rows: per warp a dense K-entry histogram (shared memory when it fits, else a
// per-warp global scratch row), filled from the document's topics, then compacted in ascending
// topic order with ballot/popc and cleared again.
__global__ void __launch_bounds__(256)
k_build_doc_rows(int64_t D, int K, const int64_t* __restrict__ doc_ptr, const uint16_t* __restrict__ z,
                 const int64_t* __restrict__ row_ptr, int32_t* __restrict__ row_nnz, uint32_t* __restrict__ rows,
                 uint32_t* __restrict__ scratch /* nullptr => shared memory */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  uint32_t* hist = scratch ? scratch + (size_t)gw * K : reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warp * K;
  for (int k = lane; k < K; k += 32) hist[k] = 0u;
  __syncwarp();
  for (int64_t d = gw; d < D; d += nw) {
    const int64_t tb = doc_ptr[d], te = doc_ptr[d + 1];
    for (int64_t i = tb + lane; i < te; i += 32) atomicAdd(hist + z[i], 1u);
    __syncwarp();
    const int64_t rp = row_ptr[d];
    int n = 0;
    for (int base = 0; base < K; base += 32) {
      const int k = base + lane;
      const uint32_t c = (k < K) ? hist[k] : 0u;
      const unsigned m = __ballot_sync(kFullMask, c != 0u);
      if (c != 0u) {
        rows[rp + n + __popc(m & ((1u << lane) - 1u))] = ((uint32_t)k << 16) | c;
        hist[k] = 0u;
      }
      n += __popc(m);
    }
    if (lane == 0) row_nnz[d] = n;
    __syncwarp();
  }
}

}  // namespace b200lda
