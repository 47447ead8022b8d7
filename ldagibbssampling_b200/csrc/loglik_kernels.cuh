// loglik_kernels.cuh — K5 (log-likelihood reduction) and K6 (theta / phi export).
//
// K5 replaces ParallelTopicModel.modelLogLikelihood() (reference cmu_ron/TrainAndPredict.java:234,
// cmu/TrainAndPredict.java:436; formula SURVEY.md §8 a6):
//   sum_d [ sum_{k: n_dk>0} (lgG(alpha_k + n_dk) - lgG(alpha_k)) - lgG(sum alpha + L_d) ] + D lgG(sum alpha)
//   + sum_{w,k: n_wk>0} lgG(beta + n_wk) - sum_k lgG(V beta + n_k) + K lgG(V beta) - nnz(n_wk) lgG(beta)
// in fp64 with CUDA's lgamma (Mallet uses a Stirling series that agrees to ~1e-10 relative).
// Reduction order is fixed (per-thread strided partials -> block tree -> one final block), so
// the value is reproducible run to run.
// K6 replaces getTopicProbabilities (cmu_ron/TrainAndPredict.java:143) and the phi that
// printTopWords / getInferencer read (cmu_ron/TrainAndPredict.java:231,169).
#pragma once
#include "device_common.cuh"

namespace b200lda {

__device__ __forceinline__ double block_reduce_sum(double v, double* s_buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(kFullMask, v, d);
  if (lane == 0) s_buf[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = (lane < (int)(blockDim.x >> 5)) ? s_buf[lane] : 0.0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) r += __shfl_down_sync(kFullMask, r, d);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// Document part: one thread per packed row slot + one term per document.
__global__ void __launch_bounds__(256)
k_loglik_docs(int64_t D, const int64_t* __restrict__ doc_ptr, const int64_t* __restrict__ row_ptr,
              const int32_t* __restrict__ row_nnz, const uint32_t* __restrict__ rows,
              const double* __restrict__ alpha, const double* __restrict__ lg_alpha, double alpha_sum,
              double* __restrict__ partial) {
  __shared__ double s_buf[32];
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  double acc = 0.0;
  for (int64_t d = gw; d < D; d += nw) {
    const int64_t rp = row_ptr[d];
    const int n = row_nnz[d];
    for (int j = lane; j < n; j += 32) {
      const uint32_t s = rows[rp + j];
      const int k = (int)(s >> 16);
      acc += lgamma(alpha[k] + (double)(s & 0xffffu)) - lg_alpha[k];
    }
    if (lane == 0) acc -= lgamma(alpha_sum + (double)(doc_ptr[d + 1] - doc_ptr[d]));
  }
  const double r = block_reduce_sum(acc, s_buf);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

// Word part: lgG(beta + n_wk) over non-zero cells, and the count of those cells.
__global__ void __launch_bounds__(256)
k_loglik_words(size_t VK, const int32_t* __restrict__ nwk, double beta, double* __restrict__ partial,
               unsigned long long* __restrict__ nonzero) {
  __shared__ double s_buf[32];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  unsigned long long nz = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < VK; i += stride) {
    const int32_t n = nwk[i];
    if (n > 0) {
      acc += lgamma(beta + (double)n);
      ++nz;
    }
  }
  const double r = block_reduce_sum(acc, s_buf);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) nz += __shfl_down_sync(kFullMask, nz, d);
  if ((threadIdx.x & 31) == 0 && nz) atomicAdd(nonzero, nz);
}

// Final, single block: out[0] = sum(partial[0..n)) in a fixed order; optionally minus the
// per-topic terms sum_k lgG(V beta + n_k).
__global__ void __launch_bounds__(256)
k_loglik_final(int n, const double* __restrict__ partial, int K, const int32_t* __restrict__ nk, double vbeta,
               double* __restrict__ out) {
  __shared__ double s_buf[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  for (int k = threadIdx.x; k < K; k += blockDim.x) acc -= lgamma(vbeta + (double)nk[k]);
  const double r = block_reduce_sum(acc, s_buf);
  if (threadIdx.x == 0) out[0] = r;
}

// theta_dk = (n_dk + alpha_k) / (L_d + sum alpha) for documents [d0, d1): one warp per document.
__global__ void __launch_bounds__(256)
k_theta(int64_t d0, int64_t d1, int K, const int64_t* __restrict__ doc_ptr, const int64_t* __restrict__ row_ptr,
        const int32_t* __restrict__ row_nnz, const uint32_t* __restrict__ rows, const double* __restrict__ alpha,
        double alpha_sum, double* __restrict__ theta) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t d = d0 + gw; d < d1; d += nw) {
    double* out = theta + (size_t)(d - d0) * K;
    const double den = (double)(doc_ptr[d + 1] - doc_ptr[d]) + alpha_sum;
    for (int k = lane; k < K; k += 32) out[k] = alpha[k] / den;
    __syncwarp();
    const int64_t rp = row_ptr[d];
    const int n = row_nnz[d];
    for (int j = lane; j < n; j += 32) {
      const uint32_t s = rows[rp + j];
      const int k = (int)(s >> 16);
      out[k] = ((double)(s & 0xffffu) + alpha[k]) / den;
    }
    __syncwarp();
  }
}

// Held-out inference (TopicInferencer.getSampledDistribution): acc[d, k] += n_dk of the current
// sample; theta follows as (S alpha_k + acc) / (S (sum alpha + L_d)) after S samples.
__global__ void __launch_bounds__(256)
k_infer_accumulate(int64_t D, int K, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ row_nnz,
                   const uint32_t* __restrict__ rows, int32_t* __restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t d = gw; d < D; d += nw) {
    const int64_t rp = row_ptr[d];
    const int n = row_nnz[d];
    for (int j = lane; j < n; j += 32) {
      const uint32_t s = rows[rp + j];
      acc[(size_t)d * K + (s >> 16)] += (int32_t)(s & 0xffffu);
    }
  }
}

__global__ void __launch_bounds__(256)
k_infer_theta(int64_t D, int K, int samples, const int64_t* __restrict__ doc_ptr, const int32_t* __restrict__ acc,
              const double* __restrict__ alpha, double alpha_sum, double* __restrict__ theta) {
  const size_t total = (size_t)D * K;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t d = (int64_t)(i / K);
    const int k = (int)(i % K);
    const double len = (double)(doc_ptr[d + 1] - doc_ptr[d]);
    theta[i] = ((double)samples * alpha[k] + (double)acc[i]) / ((double)samples * (alpha_sum + len));
  }
}

// phi[k, w] = (n_wk + beta) / (n_k + V beta) for topics [k0, k1): 32x32 shared-memory transpose
// so both the n_wk reads (row = word) and the phi writes (row = topic) are coalesced.
__global__ void __launch_bounds__(256)
k_phi(int V, int K, int k0, int k1, const int32_t* __restrict__ nwk, const int32_t* __restrict__ nk,
      double beta, double* __restrict__ phi) {
  __shared__ int32_t tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int wbase = blockIdx.x * 32, kbase = k0 + blockIdx.y * 32;
  for (int r = ty; r < 32; r += 8) {
    const int w = wbase + r, k = kbase + tx;
    tile[r][tx] = (w < V && k < k1) ? nwk[(size_t)w * K + k] : 0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int k = kbase + r, w = wbase + tx;
    if (k < k1 && w < V)
      phi[(size_t)(k - k0) * V + w] = ((double)tile[tx][r] + beta) / ((double)nk[k] + (double)V * beta);
  }
}

// Count invariants (b200lda_check_invariants): out[0] += sum of n_k, out[2] += #topics whose n_wk
// column sum differs from n_k; out[1] += sum of n_wk; out[3] += sum of the packed n_dk counts.
__global__ void k_check_nk(int K, const int32_t* __restrict__ nk, const int32_t* __restrict__ colsum,
                           unsigned long long* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  atomicAdd(out + 0, (unsigned long long)(long long)nk[k]);
  if (nk[k] != colsum[k]) atomicAdd(out + 2, 1ull);
}
__global__ void k_sum_i32(size_t n, const int32_t* __restrict__ x, unsigned long long* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  long long acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += x[i];
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, (unsigned long long)acc);
}
__global__ void k_sum_rows(int64_t D, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ row_nnz,
                           const uint32_t* __restrict__ rows, unsigned long long* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long acc = 0;
  for (int64_t d = gw; d < D; d += nw) {
    const int64_t rp = row_ptr[d];
    const int nnz = row_nnz[d];
    for (int j = lane; j < nnz; j += 32) acc += rows[rp + j] & 0xffffu;
  }
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane == 0 && acc) atomicAdd(out, acc);
}

}  // namespace b200lda
