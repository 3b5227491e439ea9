// Microbenchmark: K0 (openrdm_b200/csrc/k0_ao2mo.cuh, the library's own DMMA GEMM of the AO -> MO transform) against
// cuBLAS DGEMM on the same device-resident operands: C[m x N] = A[m x K] B[K x N], column-major.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o k0_bench k0_bench.cu -lcublas
// Run:   ./k0_bench [m=262144] [K=861] [N=103]
#include <cublas_v2.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../openrdm_b200/csrc/k0_ao2mo.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int NT> void launch(const k0::Params& p) {
  CK(cudaFuncSetAttribute(k0::k0_ao2mo<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k0::Smem<NT>::bytes));
  dim3 grid((unsigned)((p.m + k0::BM - 1) / k0::BM), (unsigned)((p.N + 8 * NT - 1) / (8 * NT)));
  k0::k0_ao2mo<NT><<<grid, k0::threads_of<NT>(), k0::Smem<NT>::bytes>>>(p);
}

int main(int argc, char** argv) {
  const size_t m = argc > 1 ? atoll(argv[1]) : 262144;
  const int K = argc > 2 ? atoi(argv[2]) : 861, N = argc > 3 ? atoi(argv[3]) : 103;
  std::vector<double> hA(m * K), hB((size_t)K * N);
  for (size_t i = 0; i < hA.size(); ++i) hA[i] = (double)((i * 2654435761ull) % 2001) / 1000.0 - 1.0;
  for (size_t i = 0; i < hB.size(); ++i) hB[i] = (double)((i * 40503ull) % 1999) / 1000.0 - 1.0;
  double *A, *B, *C1, *C2;
  CK(cudaMalloc(&A, hA.size() * 8)); CK(cudaMalloc(&B, hB.size() * 8)); CK(cudaMalloc(&C1, m * N * 8)); CK(cudaMalloc(&C2, m * N * 8));
  CK(cudaMemcpy(A, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));
  k0::Params p{A, m, 0, B, C1, m, m, K, N};
  cublasHandle_t h; cublasCreate(&h);
  const double one = 1.0, zero = 0.0;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto ours = [&]() { if (N <= 32) launch<4>(p); else if (N <= 64) launch<8>(p); else launch<13>(p); };
  auto theirs = [&]() { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, (int)m, N, K, &one, A, (int)m, B, K, &zero, C2, (int)m); };
  const double flops = 2.0 * m * K * N;
  for (int which = 0; which < 2; ++which) {
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      CK(cudaEventRecord(e0)); if (which) theirs(); else ours(); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep) best = std::min(best, ms);
    }
    printf("%-14s m %zu K %d N %d: %.3f ms  %.2f TFLOP/s (A streamed at %.0f GB/s)\n", which ? "cuBLAS DGEMM" : "k0_ao2mo", m, K, N, best, flops / best * 1e-9, m * K * 8.0 / best * 1e-6);
  }
  std::vector<double> c1(m * N), c2(m * N);
  CK(cudaMemcpy(c1.data(), C1, c1.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(c2.data(), C2, c2.size() * 8, cudaMemcpyDeviceToHost));
  double d = 0.0; for (size_t i = 0; i < c1.size(); ++i) d = std::max(d, std::abs(c1[i] - c2[i]));
  printf("max |k0 - cuBLAS| = %.3e (entries are sums of %d products of O(1) numbers)\n", d, K);
  return 0;
}
