// ovlab: how fast can the density stream (K1) run UNDERNEATH the tensor-bound K2 on the same SMs?
// Launches k2_ontop_ws<false,2,true> (lean registers) on one stream and a K1 variant on another,
// config-4 shapes (83 core + 20 active orbitals, GGA), and prints both durations and the makespan.
// Not part of the product.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../openrdm_b200/csrc/k1_density.cuh"
#include "../../openrdm_b200/csrc/k2_ontop.cuh"
#include "k12_fused.cuh"
#include "../../openrdm_b200/csrc/rdm_pack.h"
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

// variant: pure streaming read of the same bytes (no FP64): XOR of the loaded words
template <int UNR>
__global__ void __launch_bounds__(128) stream_only(const k1::Params P) {
  const size_t ld = P.ld;
  unsigned long long acc = 0;
  for (size_t p = (size_t)blockIdx.x * 128 + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * 128) {
    const int n = P.ncore + P.nact;
    int i = 0;
    for (; i + UNR <= n; i += UNR) {
      unsigned long long v[UNR][4];
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const size_t o = (size_t)(i + j) * ld + p;
        v[j][0] = __double_as_longlong(__ldcs(P.phi + o)); v[j][1] = __double_as_longlong(__ldcs(P.phix + o));
        v[j][2] = __double_as_longlong(__ldcs(P.phiy + o)); v[j][3] = __double_as_longlong(__ldcs(P.phiz + o));
      }
#pragma unroll
      for (int j = 0; j < UNR; ++j) acc ^= v[j][0] ^ v[j][1] ^ v[j][2] ^ v[j][3];
    }
    for (; i < n; ++i) { const size_t o = (size_t)i * ld + p; acc ^= __double_as_longlong(__ldcs(P.phi + o)); }
  }
  if (acc == 0x1234567ull) P.rho[0] = 1.0;
}
// variant: core stream with FP64 FMAs, UNR orbitals in flight
template <int UNR>
__global__ void __launch_bounds__(128) core_fma(const k1::Params P) {
  const size_t ld = P.ld;
  for (size_t p = (size_t)blockIdx.x * 128 + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * 128) {
    double rc = 0, cx = 0, cy = 0, cz = 0;
    const int n = P.ncore + P.nact;
    int i = 0;
    for (; i + UNR <= n; i += UNR) {
      double v[UNR][4];
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const size_t o = (size_t)(i + j) * ld + p;
        v[j][0] = __ldcs(P.phi + o); v[j][1] = __ldcs(P.phix + o); v[j][2] = __ldcs(P.phiy + o); v[j][3] = __ldcs(P.phiz + o);
      }
#pragma unroll
      for (int j = 0; j < UNR; ++j) { rc = fma(v[j][0], v[j][0], rc); cx = fma(v[j][0], v[j][1], cx); cy = fma(v[j][0], v[j][2], cy); cz = fma(v[j][0], v[j][3], cz); }
    }
    P.rho[p] = rc; P.gx[p] = cx; P.gy[p] = cy; P.gz[p] = cz;
  }
}

template <int UNR, int MINB>
__global__ void __launch_bounds__(128, MINB) core_fma_light(const k1::Params P) {
  const size_t ld = P.ld;
  for (size_t p = (size_t)blockIdx.x * 128 + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * 128) {
    double rc = 0, cx = 0, cy = 0, cz = 0;
    const int n = P.ncore + P.nact;
    int i = 0;
    for (; i + UNR <= n; i += UNR) {
      double v[UNR][4];
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const size_t o = (size_t)(i + j) * ld + p;
        v[j][0] = __ldcs(P.phi + o); v[j][1] = __ldcs(P.phix + o); v[j][2] = __ldcs(P.phiy + o); v[j][3] = __ldcs(P.phiz + o);
      }
#pragma unroll
      for (int j = 0; j < UNR; ++j) { rc = fma(v[j][0], v[j][0], rc); cx = fma(v[j][0], v[j][1], cx); cy = fma(v[j][0], v[j][2], cy); cz = fma(v[j][0], v[j][3], cz); }
    }
    P.rho[p] = rc; P.gx[p] = cx; P.gy[p] = cy; P.gz[p] = cz;
  }
}

// the product's split: core orbitals only (what would run underneath K2), UNR orbitals in flight
template <int UNR, int MINB>
__global__ void __launch_bounds__(128, MINB) core_only(const k1::Params P) {
  const size_t ld = P.ld;
  for (size_t p = (size_t)blockIdx.x * 128 + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * 128) {
    double rc = 0, cx = 0, cy = 0, cz = 0;
    int i = 0;
    for (; i + UNR <= P.ncore; i += UNR) {
      double v[UNR][4];
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const size_t o = (size_t)(i + j) * ld + p;
        v[j][0] = __ldcs(P.phi + o); v[j][1] = __ldcs(P.phix + o); v[j][2] = __ldcs(P.phiy + o); v[j][3] = __ldcs(P.phiz + o);
      }
#pragma unroll
      for (int j = 0; j < UNR; ++j) { rc = fma(v[j][0], v[j][0], rc); cx = fma(v[j][0], v[j][1], cx); cy = fma(v[j][0], v[j][2], cy); cz = fma(v[j][0], v[j][3], cz); }
    }
    for (; i < P.ncore; ++i) {
      const size_t o = (size_t)i * ld + p;
      const double v = __ldcs(P.phi + o);
      rc = fma(v, v, rc); cx = fma(v, __ldcs(P.phix + o), cx); cy = fma(v, __ldcs(P.phiy + o), cy); cz = fma(v, __ldcs(P.phiz + o), cz);
    }
    P.core_rc[p] = rc; P.core_cx[p] = cx; P.core_cy[p] = cy; P.core_cz[p] = cz;
  }
}

// core stream with UNR x 4 INDEPENDENT accumulators: every loaded value feeds its own FMA chain, so a
// warp that wins the FP64 pipe can issue 4*UNR DFMAs back to back before it stalls on a dependency
template <int UNR, int MINB>
__global__ void __launch_bounds__(128, MINB) core_fma_ilp(const k1::Params P) {
  const size_t ld = P.ld;
  for (size_t p = (size_t)blockIdx.x * 128 + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * 128) {
    double rc[UNR], cx[UNR], cy[UNR], cz[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) rc[j] = cx[j] = cy[j] = cz[j] = 0.0;
    const int n = P.ncore + P.nact;
    int i = 0;
    for (; i + UNR <= n; i += UNR) {
      double v[UNR][4];
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const size_t o = (size_t)(i + j) * ld + p;
        v[j][0] = __ldcs(P.phi + o); v[j][1] = __ldcs(P.phix + o); v[j][2] = __ldcs(P.phiy + o); v[j][3] = __ldcs(P.phiz + o);
      }
#pragma unroll
      for (int j = 0; j < UNR; ++j) { rc[j] = fma(v[j][0], v[j][0], rc[j]); cx[j] = fma(v[j][0], v[j][1], cx[j]); cy[j] = fma(v[j][0], v[j][2], cy[j]); cz[j] = fma(v[j][0], v[j][3], cz[j]); }
    }
    double a = 0, b = 0, c = 0, d = 0;
#pragma unroll
    for (int j = 0; j < UNR; ++j) { a += rc[j]; b += cx[j]; c += cy[j]; d += cz[j]; }
    P.rho[p] = a; P.gx[p] = b; P.gy[p] = c; P.gz[p] = d;
  }
}

int main(int argc, char** argv) {
  const int ncore = 83, nact = 20, n = ncore + nact;
  size_t npts = argc > 1 ? atol(argv[1]) : 5000000;
  int k1grid_mult = argc > 2 ? atoi(argv[2]) : 16;
  int k1_first = argc > 3 ? atoi(argv[3]) : 0;
  int M = nact * (nact + 1) / 2;
  std::vector<double> Dt((size_t)M * M);
  for (size_t i = 0; i < Dt.size(); ++i) Dt[i] = 1e-3 * ((i * 2654435761u) % 1000) / 1000.0;
  for (int i = 0; i < M; ++i) for (int j = 0; j < i; ++j) Dt[(size_t)i * M + j] = Dt[(size_t)j * M + i];
  size_t ld = (npts + 63) / 64 * 64;
  double* phi[4];
  for (int k = 0; k < 4; ++k) { CK(cudaMalloc(&phi[k], ld * n * 8)); CK(cudaMemset(phi[k], 0, ld * n * 8)); }
  {
    std::vector<double> h(ld * 4); for (size_t i = 0; i < h.size(); ++i) h[i] = ((i * 40503u) % 997) / 997.0 - 0.5;
    for (int k = 0; k < 4; ++k) for (int c = 0; c < n; c += 4) CK(cudaMemcpy(phi[k] + (size_t)c * ld, h.data(), std::min<size_t>(4, n - c) * ld * 8, cudaMemcpyHostToDevice));
  }
  double *w, *out[8], *partial;
  CK(cudaMalloc(&w, ld * 8)); CK(cudaMemset(w, 0, ld * 8));
  for (auto& o : out) { CK(cudaMalloc(&o, ld * 8)); }
  CK(cudaMalloc(&partial, 148 * 16 * 8 * 8));
  std::vector<uint16_t> tab; ordm::pair_table(nact, tab);
  uint16_t* dpairs; CK(cudaMalloc(&dpairs, tab.size() * 2)); CK(cudaMemcpy(dpairs, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice));
  ordm::DfragLayout lay; ordm::pack_dfrag(M, Dt, 8, k2::PREFETCH, lay);
  double* dfrag; void* dsegs;
  CK(cudaMalloc(&dfrag, lay.data.size() * 8)); CK(cudaMemcpy(dfrag, lay.data.data(), lay.data.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dsegs, lay.segs.size() * 8)); CK(cudaMemcpy(dsegs, lay.segs.data(), lay.segs.size() * 8, cudaMemcpyHostToDevice));
  k2::Params p2{}; p2.phi_act = phi[0] + (size_t)ncore * ld; p2.ld = ld; p2.npts = npts; p2.nact = nact; p2.M = M; p2.nblk = lay.nblk;
  p2.dfrag = dfrag; p2.segs = (const k2::SegDesc*)dsegs; p2.pairs = dpairs; p2.pi = out[4]; p2.pi_core = nullptr;
  for (int i = 0; i <= 8; ++i) { p2.seg_ofs[i] = lay.seg_ofs[i]; p2.blk_ofs[i] = lay.blk_ofs[i]; }
  p2.ntiles = (int)((npts + 31) / 32);
  size_t sm2 = k2::WsSmem::bytes(lay.nblk * 32, nact);
  CK(cudaFuncSetAttribute(k2::k2_ontop_ws<false, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  k1::Params p1{}; p1.phi = phi[0]; p1.phix = phi[1]; p1.phiy = phi[2]; p1.phiz = phi[3]; p1.w = w; p1.ld = ld; p1.npts = npts; p1.ncore = ncore; p1.nact = nact;
  p1.rho = out[0]; p1.gx = out[1]; p1.gy = out[2]; p1.gz = out[3]; p1.pi_core = out[5]; p1.partial = partial;
  std::vector<double> Ds(k1::MAX_ACT * k1::MAX_ACT, 0.01);
  CK(cudaMemcpyToSymbol(k1::c_Dsa, Ds.data(), Ds.size() * 8)); CK(cudaMemcpyToSymbol(k1::c_Dsb, Ds.data(), Ds.size() * 8));
  cudaStream_t sa, sb; CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
  cudaEvent_t a0, a1, b0, b1, fork; cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1); cudaEventCreate(&fork);
  const int g1 = 148 * k1grid_mult;
  const double gb = (double)npts * n * 4 * 8 * 1e-9;
  auto launch_k1 = [&](int variant, cudaStream_t s) {
    switch (variant) {
      case 0: k1::k1_density<20, true><<<g1, 128, 0, s>>>(p1); break;
      case 1: stream_only<4><<<g1, 128, 0, s>>>(p1); break;
      case 2: stream_only<8><<<g1, 128, 0, s>>>(p1); break;
      case 3: core_fma<4><<<g1, 128, 0, s>>>(p1); break;
      case 4: core_fma<8><<<g1, 128, 0, s>>>(p1); break;
      case 5: stream_only<16><<<g1, 128, 0, s>>>(p1); break;
      case 6: core_fma_light<2, 12><<<g1, 128, 0, s>>>(p1); break;   // <= 40 regs: 3 CTAs next to K2
      case 7: core_fma_light<1, 16><<<g1, 128, 0, s>>>(p1); break;   // <= 32 regs: 4 CTAs next to K2
      case 8: core_fma_light<2, 8><<<g1, 128, 0, s>>>(p1); break;    // <= 64 regs: 2 CTAs next to K2
      case 9: core_fma_ilp<4, 4><<<g1, 128, 0, s>>>(p1); break;      // 16 independent chains
      case 10: core_fma_ilp<8, 4><<<g1, 128, 0, s>>>(p1); break;     // 32 independent chains
      case 11: core_only<2, 8><<<g1, 128, 0, s>>>(p1); break;
      case 12: core_only<4, 8><<<g1, 128, 0, s>>>(p1); break;
      case 13: k1::k1_core<true><<<g1 / 2, k1::CORE_THREADS, 0, s>>>(p1); break;
    }
  };
  // a K1 launched BEFORE K2 must leave the SM's shared-memory carve-out large enough for K2's 147 KB
  CK(cudaFuncSetAttribute(k1::k1_density<20, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute(stream_only<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute(stream_only<8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute(stream_only<16>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute(core_fma<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute(core_fma<8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  p1.core_rc = out[0]; p1.core_cx = out[1]; p1.core_cy = out[2]; p1.core_cz = out[3];
  const char* names[] = {"k1_density<20,gga>", "stream_only<4>", "stream_only<8>", "core_fma<4>", "core_fma<8>", "stream_only<16>", "core_fma_light<2,12>", "core_fma_light<1,16>", "core_fma_light<2,8>", "core_fma_ilp<4>", "core_fma_ilp<8>", "core_only<2> (83 orb)", "core_only<4> (83 orb)", "k1_core (product, 256 thr)"};
  for (int variant = (argc > 4 ? atoi(argv[4]) : 0); variant < 14; ++variant) {
    float alone = 0, tk2 = 0, tk1 = 0, span = 0;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(b0, sb); launch_k1(variant, sb); cudaEventRecord(b1, sb); CK(cudaEventSynchronize(b1)); cudaEventElapsedTime(&alone, b0, b1); }
    for (int r = 0; r < 3; ++r) {
      CK(cudaDeviceSynchronize());
      cudaEventRecord(fork, sa); cudaStreamWaitEvent(sb, fork, 0);
      if (k1_first) { cudaEventRecord(b0, sb); launch_k1(variant, sb); cudaEventRecord(b1, sb); }
      cudaEventRecord(a0, sa); k2::k2_ontop_ws<false, 2, true><<<148, k2::WS_THREADS, sm2, sa>>>(p2); cudaEventRecord(a1, sa);
      if (!k1_first) { cudaEventRecord(b0, sb); launch_k1(variant, sb); cudaEventRecord(b1, sb); }
      CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&tk2, a0, a1); cudaEventElapsedTime(&tk1, b0, b1); cudaEventElapsedTime(&span, k1_first ? b0 : a0, k1_first ? a1 : b1);
    }
    printf("%-20s alone %.3f ms (%.0f GB/s) | with K2: K2 %.3f ms, K1 %.3f ms (%.0f GB/s), a0->b1 %.3f ms\n", names[variant], alone, gb / alone * 1e3, tk2, tk1, gb / tk1 * 1e3, span);
  }
  {  // the remaining active-space part on the core sums, alone
    float t = 0;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(b0, sb); k1::k1_density<20, true><<<g1, 128, 0, sb>>>(p1); cudaEventRecord(b1, sb); CK(cudaEventSynchronize(b1)); cudaEventElapsedTime(&t, b0, b1); }
    printf("k1_density<20,gga> with core sums given (active part only): %.3f ms\n", t);
    p1.core_rc = nullptr;
  }
  {  // the fused kernel: density warps inside the K2 CTA
    CK(cudaFuncSetAttribute(k12::k12_fused<20, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
    float t = 0;
    for (int r = 0; r < 4; ++r) {
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a0, sa); k12::k12_fused<20, true><<<148, k12::THREADS, sm2, sa>>>(p2, p1); cudaEventRecord(a1, sa);
      CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&t, a0, a1);
    }
    printf("k12_fused<20,gga>: %.3f ms\n", t);
    p1.core_rc = out[0]; p1.core_cx = out[1]; p1.core_cy = out[2]; p1.core_cz = out[3];
    CK(cudaFuncSetAttribute(k12::k12_fused_core<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
    CK(cudaFuncSetAttribute(k12::k12_fused_core<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
    for (int r = 0; r < 4; ++r) {
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a0, sa); k12::k12_fused_core<true, 8><<<148, k12::THREADS, sm2, sa>>>(p2, p1); cudaEventRecord(a1, sa);
      CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&t, a0, a1);
    }
    printf("k12_fused_core<gga,8>: %.3f ms\n", t);
    for (int r = 0; r < 4; ++r) {
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a0, sa); k12::k12_fused_core<true, 4><<<148, k12::THREADS, sm2, sa>>>(p2, p1); cudaEventRecord(a1, sa);
      CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&t, a0, a1);
    }
    printf("k12_fused_core<gga,4>: %.3f ms\n", t);
    CK(cudaFuncSetAttribute(k12::k12_fused_core<true, 8, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
    for (int r = 0; r < 4; ++r) {
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a0, sa); k12::k12_fused_core<true, 8, 12><<<148, k12::THREADS, sm2, sa>>>(p2, p1); cudaEventRecord(a1, sa);
      CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&t, a0, a1);
    }
    printf("k12_fused_core<gga,8>, density = warps 8-11, builders = warps 12-15: %.3f ms\n", t);
    p1.ncore = 0;  // no density work at all: the fused shell alone
    for (int r = 0; r < 3; ++r) {
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a0, sa); k12::k12_fused_core<true, 4><<<148, k12::THREADS, sm2, sa>>>(p2, p1); cudaEventRecord(a1, sa);
      CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&t, a0, a1);
    }
    printf("k12_fused_core, ncore=0 (stores only): %.3f ms\n", t);
  }
  return 0;
}
