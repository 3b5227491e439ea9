// Microbenchmark: FP64 pipe ceilings on B200 (sm_100a).
//   - DMMA.8x8x4 issue rate (mma.sync.m8n8k4.f64), register-resident operands
//   - DFMA issue rate
//   - streaming read bandwidth (LDG.128)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int NACC>
__global__ void __launch_bounds__(384) dmma_rate(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int NACC>
__global__ void __launch_bounds__(256) dfma_rate(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) stream_read(const double2* __restrict__ in, size_t n2, double* out) {
  double s = 0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    double2 v0 = __ldg(in + i), v1 = __ldg(in + i + stride), v2 = __ldg(in + i + 2 * stride), v3 = __ldg(in + i + 3 * stride);
    s += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
  }
  for (; i < n2; i += stride) { double2 v = __ldg(in + i); s += v.x + v.y; }
  if (s == 123.456) out[0] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s SMs %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int ctas = 1; ctas <= 4; ctas *= 2) {
    int iters = 20000;
    int grid = sms * ctas;
    dmma_rate<8><<<grid, 256>>>(out, 100);
    CK(cudaDeviceSynchronize());
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0); dmma_rate<8><<<grid, 256>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      double flop = (double)grid * 8 /*warps*/ * iters * 8.0 * 512.0;
      printf("DMMA.8x8x4 NACC=8 ctas/SM=%d: %.3f ms  %.2f TFLOP/s\n", ctas, ms, flop / ms * 1e-9);
    }
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0); dmma_rate<16><<<grid, 256>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      double flop = (double)grid * 8 * iters * 16.0 * 512.0;
      printf("DMMA.8x8x4 NACC=16 ctas/SM=%d: %.3f ms  %.2f TFLOP/s\n", ctas, ms, flop / ms * 1e-9);
    }
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0); dfma_rate<16><<<grid, 256>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      double flop = (double)grid * 256 * iters * 16.0 * 2.0;
      printf("DFMA NACC=16 ctas/SM=%d: %.3f ms  %.2f TFLOP/s\n", ctas, ms, flop / ms * 1e-9);
    }
  }
  // can ONE warp per SM sub-partition keep the DMMA pipe full?  (128-thread CTA, 1 CTA/SM)
  for (int rep = 0; rep < 3; rep++) {
    int iters = 20000;
    cudaEventRecord(e0); dmma_rate<16><<<sms, 128>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    double flop = (double)sms * 4 * iters * 16.0 * 512.0;
    printf("DMMA.8x8x4 NACC=16 ONE warp/SMSP: %.3f ms  %.2f TFLOP/s\n", ms, flop / ms * 1e-9);
  }
  for (int rep = 0; rep < 2; rep++) {
    int iters = 20000;
    cudaEventRecord(e0); dmma_rate<16><<<sms, 384>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    double flop = (double)sms * 12 * iters * 16.0 * 512.0;
    printf("DMMA.8x8x4 NACC=16 THREE warps/SMSP: %.3f ms  %.2f TFLOP/s\n", ms, flop / ms * 1e-9);
  }
  // streaming read
  size_t bytes = (size_t)8 << 30;
  double2* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
  for (int ctas = 2; ctas <= 8; ctas *= 2) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0); stream_read<<<sms * ctas, 256>>>(buf, bytes / 16, out); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      printf("stream_read 8GiB ctas/SM=%d: %.3f ms  %.1f GB/s\n", ctas, ms, bytes / ms * 1e-6);
    }
  }
  return 0;
}
