// Microbenchmark: issue rates on sm_100a of the candidates for the int8 K2's epilogue arithmetic
//   mad.wide.s32 (IMAD.WIDE: s32 x s32 + s64), DFMA, and the exact int32 -> double conversion pattern (LOP3 + MOV + DADD) + DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_wide_rate imad_wide_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template <int NACC>
__global__ void __launch_bounds__(256) imad_wide(long long* out, int iters, int a0, int b0) {
  long long c[NACC];
  int a = a0 + threadIdx.x, b = b0 - threadIdx.x;
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(c[i]) : "r"(a + i), "r"(b));
  }
  long long s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma(double* out, int iters) {
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

// the fold's per-element pattern: v -> exact double (xor, hi word, DADD) -> DFMA into one of NACC chains
template <int NACC>
__global__ void __launch_bounds__(256) cvt_dfma(double* out, int iters, unsigned v0) {
  double phi = 1.0 + threadIdx.x * 1e-9;
  double c[NACC];
  unsigned v[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i] = i; v[i] = v0 + i * 977u + threadIdx.x; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      const double zd = __hiloint2double(0x43300000, (int)(v[i] ^ 0x80000000u)) - 4503601774854144.0;
      c[i] = fma(phi, zd, c[i]);
      v[i] += 0x9e3779b9u;   // one more integer op per element stands in for the loop's bookkeeping
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount, iters = 4096;
  void* out; CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double clk = p.clockRate * 1e3;
  for (int cps = 1; cps <= 4; cps *= 2) {
    float ms;
    auto report = [&](const char* name, double per_thread) {
      const double lanes = per_thread * 256.0 * sms * cps / (ms * 1e-3);
      printf("%-28s ctas/SM=%d: %.3f ms  %.1f lanes/clk/SM (%.2f clk per warp instruction per sub-partition)\n", name, cps, ms, lanes / clk / sms, 32.0 * 4 / (lanes / clk / sms));
    };
    for (int rep = 0; rep < 2; ++rep) { CK(cudaEventRecord(e0)); imad_wide<16><<<sms * cps, 256>>>((long long*)out, iters, 3, 7); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms, e0, e1)); }
    report("IMAD.WIDE (s32*s32+s64)", 16.0 * iters);
    for (int rep = 0; rep < 2; ++rep) { CK(cudaEventRecord(e0)); dfma<16><<<sms * cps, 256>>>((double*)out, iters); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms, e0, e1)); }
    report("DFMA", 16.0 * iters);
    for (int rep = 0; rep < 2; ++rep) { CK(cudaEventRecord(e0)); cvt_dfma<16><<<sms * cps, 256>>>((double*)out, iters, 12345u); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms, e0, e1)); }
    report("xor+hi+DADD+DFMA (per elem)", 16.0 * iters);
  }
  return 0;
}
