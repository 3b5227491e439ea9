// reduce.cuh -- K4: deterministic FP64 quadrature sums (warp shuffle -> block -> fixed-order
// final pass; no floating-point atomics anywhere).  Replaces the serial `exc += ... * W(p)`
// tails of mcpdft.cc:91-93,332-334 and functional.cc:26,80,86,92,170,302.
#pragma once
#include <cuda_runtime.h>

namespace ordm {

// slots of the device result vector
enum { RES_N = 0, RES_NA, RES_NB, RES_TRN, RES_TRNA, RES_TRNB, RES_EX, RES_EC, RES_COUNT };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sums of NV per-thread values; thread 0 writes them to out[0..NV).
// `scratch` = NV * 32 doubles of shared memory.  All threads of the block must call.
template <int NV>
__device__ __forceinline__ void block_sum_store(const double (&v)[NV], double* scratch, double* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) scratch[k * 32 + wid] = s;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = (lane < nw) ? scratch[k * 32 + lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) out[k] = s;
    }
  }
}

// Final pass: one block.  partial is [nrows][NV] row-major; res[slot0 + k] (+)= sum over rows.
// Rows are summed in a fixed (thread-strided, then tree) order => bitwise reproducible for a
// given launch geometry.
template <int NV>
__global__ void __launch_bounds__(256) k4_finalize(const double* __restrict__ partial, int nrows,
                                                   double* __restrict__ res, int slot0, int accumulate) {
  __shared__ double scratch[NV * 32];
  double v[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = 0.0;
  for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] += partial[(size_t)r * NV + k];
  }
  __shared__ double tot[NV];
  block_sum_store<NV>(v, scratch, tot);
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = tot[threadIdx.x];
    if (accumulate) s += res[slot0 + threadIdx.x];
    res[slot0 + threadIdx.x] = s;
  }
}

}  // namespace ordm
