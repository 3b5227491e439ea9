// Microbenchmark: how much DMMA pipe time does an interleaved FP64 FMA/MUL cost on B200?
// 12 warps per SM: warps 4..11 (2 per sub-partition) stream DMMA.8x8x4; warps 0..3 (1 per
// sub-partition) issue R FP64 multiplies per DMMA-loop iteration.  Also: same-warp interleave.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template <int R>
__global__ void __launch_bounds__(384) mix(double* out, int iters) {
  const int wid = threadIdx.x >> 5;
  if (wid >= 4) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i][0] = c[i][1] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-9;
    const double m = 1.0000001;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < R; r++) asm volatile("mul.f64 %0, %0, %1;" : "+d"(v[r & 7]) : "d"(m));
      if (R == 0) asm volatile("" ::: "memory");
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

template <int R>
__global__ void __launch_bounds__(256) mix_same(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c[16][2], v[8];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i][0] = c[i][1] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = 1.0 + i * 1e-3;
  const double m = 1.0000001;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
    for (int r = 0; r < R; r++) asm volatile("mul.f64 %0, %0, %1;" : "+d"(v[r & 7]) : "d"(m));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 8; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int R> void run(double* out, int sms) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 10000; float ms;
  mix<R><<<sms, 384>>>(out, 100); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); mix<R><<<sms, 384>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
  double cyc = ms * 1e-3 * 1.965e9 / iters;
  printf("other-warp  R=%3d DMUL/iter: %.3f ms  %.1f clk/iter (DMMA-only ideal 512)  extra per DMUL = %.2f clk\n", R, ms, cyc, R ? (cyc - 512.0) / R : 0.0);
  mix_same<R><<<sms, 256>>>(out, 100); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); mix_same<R><<<sms, 256>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
  cyc = ms * 1e-3 * 1.965e9 / iters;
  printf("same-warp   R=%3d DMUL/iter: %.3f ms  %.1f clk/iter (ideal 512)  extra per DMUL = %.2f clk\n", R, ms, cyc, R ? (cyc - 512.0) / (2 * R) : 0.0);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 512));
  run<0>(out, p.multiProcessorCount); run<1>(out, p.multiProcessorCount); run<2>(out, p.multiProcessorCount);
  run<4>(out, p.multiProcessorCount); run<8>(out, p.multiProcessorCount); run<16>(out, p.multiProcessorCount); run<32>(out, p.multiProcessorCount);
  return 0;
}
