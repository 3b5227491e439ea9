// k2i_lab.cu -- lab harness for the product kernel openrdm_b200/csrc/k2_int8.cuh: random active-space orbitals and a dense symmetric D~,
// CPU reference in long double on a sample of points, timing with CUDA events.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o k2i_lab k2i_lab.cu
// run:   timeout 60 ./k2i_lab [npts=2000000] [ncheck=2048]
#define K2I_PROF 1
#include "../../openrdm_b200/csrc/k2_int8.cuh"
#include "../../openrdm_b200/csrc/rdm_pack.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
  const size_t npts = argc > 1 ? (size_t)atoll(argv[1]) : 2000000;
  const size_t ncheck = std::min(npts, argc > 2 ? (size_t)atoll(argv[2]) : (size_t)2048);
  constexpr int NA = 20, M = NA * (NA + 1) / 2, KP = 224;
  std::mt19937_64 rng(42);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::uniform_real_distribution<double> ud(-1.0, 1.0);
  // orbitals: a per-point envelope spanning ~6 decades times per-orbital amplitudes
  std::vector<double> phi((size_t)NA * npts);
  for (size_t p = 0; p < npts; ++p) {
    const double env = std::exp(-14.0 * (double)((p * 2654435761ull) % 1000003ull) / 1000003.0);
    for (int t = 0; t < NA; ++t) phi[(size_t)t * npts + p] = env * 0.5 * nd(rng) * (1.0 + 0.3 * t);
  }
  std::vector<double> D((size_t)M * M);
  for (int i = 0; i < M; ++i) for (int j = i; j < M; ++j) { const double v = ud(rng) * std::exp(-3.0 * std::fabs(ud(rng))); D[(size_t)i * M + j] = D[(size_t)j * M + i] = v; }
  std::vector<uint16_t> tu(704, 0);
  { int I = 0; for (int t = 0; t < NA; ++t) for (int u = t; u < NA; ++u) tu[I++] = (uint16_t)(t | (u << 8)); }
  std::vector<uint8_t> timg; int t_exp = 0;
  const bool pair = false;   // the CTA-pair kernel has its own harness: k2p_lab.cu
  ordm::pack_t_int8(M, D, KP, timg, t_exp);
  double *dphi, *dpi; uint8_t* dt; int* dfail;
  CK(cudaMalloc(&dphi, phi.size() * 8)); CK(cudaMalloc(&dpi, npts * 8)); CK(cudaMalloc(&dt, timg.size())); CK(cudaMalloc(&dfail, 8));
  CK(cudaMemcpy(dphi, phi.data(), phi.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt, timg.data(), timg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dfail, 0, 8)); CK(cudaMemset(dpi, 0, npts * 8));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const size_t smem = k2i::Smem::bytes(KP, NA);
  printf("smem %zu B (opt-in max %zu), T image %zu B, t_exp %d\n", smem, (size_t)prop.sharedMemPerBlockOptin, timg.size(), t_exp);
  CK(cudaFuncSetAttribute(k2i::k2i_ontop<NA, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2i::Smem::bytes(KP, NA)));
  CK(cudaFuncSetAttribute(k2i::k2i_ontop<NA, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2i::Smem::bytes(KP, NA)));
    const int gmin = argc > 3 ? atoi(argv[3]) : 4;
  k2i::Params P{dphi, npts, npts, dt, t_exp, nullptr, dpi, dfail, 0.0, 0.0};
  auto launch = [&](int grid, size_t smem) { if (gmin == 4) k2i::k2i_ontop<NA, 4><<<grid, k2i::THREADS, smem>>>(P); else k2i::k2i_ontop<NA, 5><<<grid, k2i::THREADS, smem>>>(P); };
  const int grid = (int)std::min<size_t>((size_t)prop.multiProcessorCount, (npts + k2i::TILE - 1) / k2i::TILE);
  launch(grid, smem);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0)); launch(grid, smem); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
  }
  std::vector<double> pi(npts); int fail;
  CK(cudaMemcpy(pi.data(), dpi, npts * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&fail, dfail, 4, cudaMemcpyDeviceToHost));
  double emax = 0.0, emax_f64 = 0.0, pimax = 0.0; size_t worst = 0;
  std::vector<long double> x(M);
  for (size_t c = 0; c < ncheck; ++c) {
    const size_t p = (c * 7919) % npts;
    for (int I = 0; I < M; ++I) x[I] = (long double)(phi[(size_t)(tu[I] & 255) * npts + p] * phi[(size_t)(tu[I] >> 8) * npts + p]);
    long double ref = 0.0L; double f64 = 0.0;
    for (int I = 0; I < M; ++I) { long double y = 0.0L; double yd = 0.0; for (int J = 0; J < M; ++J) { y += x[J] * (long double)D[(size_t)I * M + J]; yd += (double)x[J] * D[(size_t)I * M + J]; } ref += x[I] * y; f64 += (double)x[I] * yd; }
    const double sc = std::max(1.0, (double)fabsl(ref));
    const double e = std::fabs((double)((long double)pi[p] - ref)) / sc;
    if (e > emax) { emax = e; worst = p; }
    emax_f64 = std::max(emax_f64, std::fabs((double)((long double)f64 - ref)) / sc); pimax = std::max(pimax, (double)fabsl(ref));
  }
  { long long pf[3][8]; CK(cudaMemcpyFromSymbol(pf, k2i::k2i_prof, sizeof(pf)));
    const double nt = (double)((npts + 127) / 128 + grid - 1) / grid;
    const char* who[3] = {"left-half thread", "right-half thread", "MMA lane"};
    const char* nm[8] = {"S1 phi", "build+S2", "wait drain", "issue", "fold", "wait full", "acc load", "S5"};
    for (int w = 0; w < 3; ++w) { printf("%s, clk per tile:", who[w]); for (int i = 0; i < 8; ++i) printf(" %s %.0f;", nm[i], (double)pf[w][i] / nt); printf("\n"); } }
  if (!pair) { long long tr[3][40]; CK(cudaMemcpyFromSymbol(tr, k2i::k2i_trace, sizeof(tr)));
    const long long t0 = tr[2][0];
    printf("timeline of CTA 0's tile 40 (clk after the tile start; groups in issue order)\n  MMA lane: build done %lld;", tr[2][1] - t0);
    for (int g = 0; g < (gmin == 4 ? 7 : 6); ++g) printf(" g%d R[wait %lld issued %lld] L[wait %lld issued %lld];", g, tr[2][2 + 4 * g] - t0, tr[2][3 + 4 * g] - t0, tr[2][4 + 4 * g] - t0, tr[2][5 + 4 * g] - t0);
    printf(" end %lld\n", tr[2][38] - t0);
    for (int w = 0; w < 2; ++w) { printf("  %s thread: build done %lld;", w ? "right" : "left", tr[w][1] - t0); for (int g = 0; g < (gmin == 4 ? 7 : 6); ++g) printf(" g%d [full %lld folded %lld];", g, tr[w][2 + 2 * g] - t0, tr[w][3 + 2 * g] - t0); printf(" end %lld\n", tr[w][38] - t0); } }
  const double flops = (double)npts * ((double)M * (M + 1) + 2.0 * M);
  printf("GMIN %d: npts %zu: %.3f ms -> %.1f M pts/s (FP64-equivalent %.1f TFLOP/s of the symmetric contraction); barrier timeouts %d\n", gmin, npts, best, npts / best / 1e3, flops / best / 1e9, fail);
  printf("max |dPi| / max(1,|Pi|) over %zu points: int8 path %.3e (worst point %zu: got %.17g), plain FP64 loops %.3e; |Pi|max %.3g\n", ncheck, emax, worst, pi[worst], emax_f64, pimax);
  return (fail || !(emax < 1e-12)) ? 2 : 0;
}
