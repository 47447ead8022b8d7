// hyper_kernels.cuh — statistics for hyper-parameter optimisation.
//
// Replaces the statistics Mallet's workers collect for ParallelTopicModel.optimizeAlpha /
// optimizeBeta (enabled by setOptimizeInterval(20) at reference cmu_ron/TrainAndPredict.java:163,
// cmu/TrainAndPredict.java:261):
//   * docLengthCounts[n] / topicDocCounts[k][n] — histograms over documents, read here straight
//     off the packed n_dk rows (one integer atomic per non-zero doc topic);
//   * the beta update needs sum over non-zero cells of  psi(b + n_wk) - psi(b): Mallet walks a
//     histogram of cell values on the host; here each fixed-point iteration is one streaming pass
//     over n_wk with the digamma difference evaluated per cell in fp64 (no histogram, no bound on
//     the largest count).
#pragma once
#include "device_common.cuh"
#include "loglik_kernels.cuh"

namespace b200lda {

// hist layout: (K + 1) rows of `width` int32; row k < K: topicDocCounts[k][count], row K:
// docLengthCounts[length]. Counts/lengths >= width cannot occur (width > longest document).
__global__ void __launch_bounds__(256)
k_hyper_collect(int64_t D, int K, int width, const int64_t* __restrict__ doc_ptr, const int64_t* __restrict__ row_ptr,
                const int32_t* __restrict__ row_nnz, const uint32_t* __restrict__ rows, int32_t* __restrict__ hist) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t d = gw; d < D; d += nw) {
    const int64_t rp = row_ptr[d];
    const int n = row_nnz[d];
    for (int j = lane; j < n; j += 32) {
      const uint32_t s = rows[rp + j];
      atomicAdd(hist + (size_t)(s >> 16) * width + (s & 0xffffu), 1);
    }
    if (lane == 0) atomicAdd(hist + (size_t)K * width + (int)(doc_ptr[d + 1] - doc_ptr[d]), 1);
  }
}

// psi(x + n) - psi(x) for integer n >= 1: the exact finite sum for small n (what Mallet's running
// sum computes), the asymptotic digamma series for large n.
__device__ __forceinline__ double digamma_asym(double z) {  // z >= 10
  const double iz = 1.0 / z, iz2 = iz * iz;
  return log(z) - 0.5 * iz -
         iz2 * (1.0 / 12 - iz2 * (1.0 / 120 - iz2 * (1.0 / 252 - iz2 * (1.0 / 240 - iz2 * (1.0 / 132)))));
}
__device__ __forceinline__ double digamma_rise(double x, int n) {
  if (n <= 24) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += 1.0 / (x + i);
    return s;
  }
  // lift both arguments above 10 with the recurrence psi(z) = psi(z + 1) - 1/z
  double lo = 0.0, a = x;
  while (a < 10.0) {
    lo -= 1.0 / a;
    a += 1.0;
  }
  return digamma_asym(x + n) - (lo + digamma_asym(a));
}

// partial[b] = sum over this block's non-zero cells of  psi(param + n) - psi(param)
__global__ void __launch_bounds__(256)
k_beta_numerator(size_t VK, const int32_t* __restrict__ nwk, double param, double* __restrict__ partial) {
  __shared__ double s_buf[32];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < VK; i += stride) {
    const int32_t n = nwk[i];
    if (n > 0) acc += digamma_rise(param, n);
  }
  const double r = block_reduce_sum(acc, s_buf);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

}  // namespace b200lda
