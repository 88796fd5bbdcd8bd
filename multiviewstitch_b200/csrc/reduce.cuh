// reduce.cuh — deterministic block / grid sum reductions in double precision.
//
// Pattern used by every reduction on the path (Umeyama sums, PCA covariance, residual scoring):
// warp-shuffle tree -> shared memory across warps -> one partial row per CTA in global memory ->
// the LAST CTA to finish (atomic ticket) adds the rows in CTA order.  The summation tree is fixed
// by (grid, block) alone, so results are bit-reproducible run to run.
#pragma once
#include "common.cuh"

namespace mvs {

template <int N>
__device__ __forceinline__ void warp_reduce_sum(double (&v)[N]) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
  }
}

// After the call thread 0 holds the block sums in v[].  smem: at least 32*N doubles.
template <int N>
__device__ __forceinline__ void block_reduce_sum(double (&v)[N], double* smem) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  warp_reduce_sum<N>(v);
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < N; ++j) smem[warp * N + j] = v[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; ++w) {
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] += smem[w * N + j];
    }
  }
  __syncthreads();
}

// Grid-wide sum.  Every CTA calls this with its per-thread partials.  Returns true for all
// threads of the last CTA, after out[0..N) (global memory) holds the totals.  `partials` must
// hold gridDim.x*N doubles, `ticket` one zero-initialised unsigned (reset to zero on exit).
template <int N>
__device__ __forceinline__ bool grid_reduce_sum(double (&v)[N], double* smem, double* partials,
                                                unsigned* ticket, double* out) {
  __shared__ bool is_last;
  block_reduce_sum<N>(v, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < N; ++j) partials[(size_t)blockIdx.x * N + j] = v[j];
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  // every thread takes the partial rows b = t, t + blockDim, ... (independent loads, one memory
  // latency) and the CTA adds them with the same fixed tree as above
  double w[N];
#pragma unroll
  for (int j = 0; j < N; ++j) w[j] = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int j = 0; j < N; ++j) w[j] += __ldcg(partials + (size_t)b * N + j);
  }
  block_reduce_sum<N>(w, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < N; ++j) out[j] = w[j];
    *ticket = 0u;
  }
  __threadfence();
  __syncthreads();
  return true;
}

}  // namespace mvs
