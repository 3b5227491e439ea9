// k2lab: stand-alone timing harness for the K2 kernels with experiment switches (-DK2_EXP_*).
// Not part of the product; used to find what bounds the DMMA stream.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#ifndef K2LAB_LEAN
#define K2LAB_LEAN false
#endif
#include "../../openrdm_b200/csrc/k2_ontop.cuh"
#include "../../openrdm_b200/csrc/rdm_pack.h"
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
int main(int argc, char** argv) {
  int nact = argc > 1 ? atoi(argv[1]) : 20;
  size_t npts = argc > 2 ? atol(argv[2]) : 2000000;
  int M = nact * (nact + 1) / 2;
  std::vector<double> Dt((size_t)M * M);
  for (size_t i = 0; i < Dt.size(); ++i) Dt[i] = 1e-3 * ((i * 2654435761u) % 1000) / 1000.0;
  for (int i = 0; i < M; ++i) for (int j = 0; j < i; ++j) Dt[(size_t)i * M + j] = Dt[(size_t)j * M + i];
  size_t ld = (npts + 63) / 64 * 64;
  double *phi, *pi; CK(cudaMalloc(&phi, ld * nact * 8)); CK(cudaMalloc(&pi, ld * 8));
  std::vector<double> h(ld * nact); for (size_t i = 0; i < h.size(); ++i) h[i] = ((i * 40503u) % 997) / 997.0 - 0.5;
  CK(cudaMemcpy(phi, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  std::vector<uint16_t> tab; ordm::pair_table(nact, tab);
  uint16_t* dpairs; CK(cudaMalloc(&dpairs, tab.size() * 2)); CK(cudaMemcpy(dpairs, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ws = 1; ws >= 0; --ws) {
    ordm::DfragLayout lay; ordm::pack_dfrag(M, Dt, ws ? 8 : 4, k2::PREFETCH, lay);
    double* dfrag; void* dsegs;
    CK(cudaMalloc(&dfrag, lay.data.size() * 8)); CK(cudaMemcpy(dfrag, lay.data.data(), lay.data.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dsegs, lay.segs.size() * 8)); CK(cudaMemcpy(dsegs, lay.segs.data(), lay.segs.size() * 8, cudaMemcpyHostToDevice));
    k2::Params p{}; p.phi_act = phi; p.ld = ld; p.npts = npts; p.nact = nact; p.M = M; p.nblk = lay.nblk;
    p.dfrag = dfrag; p.segs = (const k2::SegDesc*)dsegs; p.pairs = dpairs; p.pi = pi; p.pi_core = nullptr;
    for (int i = 0; i <= 8; ++i) { p.seg_ofs[i] = lay.seg_ofs[i]; p.blk_ofs[i] = lay.blk_ofs[i]; }
    int Mpad = lay.nblk * 32; float ms = 0;
    if (ws) {
      size_t sm = k2::WsSmem::bytes(Mpad, nact);
      CK(cudaFuncSetAttribute(k2::k2_ontop_ws<false, 2, K2LAB_LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      p.ntiles = (int)((npts + 31) / 32);
      for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k2::k2_ontop_ws<false, 2, K2LAB_LEAN><<<148, k2::WS_THREADS, sm>>>(p); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); }
    } else {
      size_t sm = k2::Smem<64>::bytes(Mpad, nact);
      CK(cudaFuncSetAttribute(k2::k2_ontop<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      p.ntiles = (int)((npts + 63) / 64);
      for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k2::k2_ontop<2, 4><<<148, k2::THREADS, sm>>>(p); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); }
    }
    double dmma = (double)npts / 8 * lay.subtiles;
    printf("%s nact=%d npts=%zu: %.3f ms  executed %.2f TFLOP/s (%.1f%% of 37.1)  subtiles=%lu maxlane=%lu\n", ws ? "WS   " : "plain", nact, npts, ms,
           dmma * 512 / ms * 1e-9, dmma * 512 / ms * 1e-9 / 37.1 * 100, (unsigned long)lay.subtiles, (unsigned long)lay.max_lane_cost);
#ifdef K2_EXP_TRACE
    if (ws) {
      static long long tr[12][64][4]; CK(cudaMemcpyFromSymbol(tr, k2::g_trace, sizeof(tr)));
      long long t0 = tr[0][8][0];
      for (int k = 8; k < 12; ++k) {
        printf("tile %d\n", k);
        for (int w = 0; w < 12; ++w) printf("  warp %2d (%s): %8lld %8lld %8lld %8lld   seg/stream=%lld\n", w, w >= 8 ? "build" : "mma  ", tr[w][k][0] - t0, tr[w][k][1] - t0, tr[w][k][2] - t0, tr[w][k][3] - t0, tr[w][k][2] - tr[w][k][1]);
      }
      for (int l = 0; l < 8; ++l) { long c = 0; for (uint32_t sgi = lay.seg_ofs[l]; sgi < lay.seg_ofs[l + 1]; ++sgi) c += lay.segs[sgi].nblk; printf("lane %d: %u segs, %ld blocks\n", l, lay.seg_ofs[l + 1] - lay.seg_ofs[l], c); }
    }
#endif
    std::vector<double> out(8); CK(cudaMemcpy(out.data(), pi, 64, cudaMemcpyDeviceToHost)); printf("   pi[0..2] = %.6e %.6e %.6e\n", out[0], out[1], out[2]);
  }
  return 0;
}
