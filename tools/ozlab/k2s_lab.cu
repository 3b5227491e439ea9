// lab harness for k2i_ontop_stream (csrc/k2_int8_stream.cuh, 21 .. 26 active orbitals): random orbitals and a dense
// symmetric D~, CPU reference in long double on a sample of points, timing with CUDA events.
#define K2S_PROF 1
#include "k2_int8_stream.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#ifndef LAB_NA
#define LAB_NA 26
#endif
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
  const size_t npts = argc > 1 ? (size_t)atoll(argv[1]) : 500000;
  const size_t ncheck = std::min(npts, argc > 2 ? (size_t)atoll(argv[2]) : (size_t)512);
  constexpr int NA = LAB_NA, M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32;
  std::mt19937_64 rng(42);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::uniform_real_distribution<double> ud(-1.0, 1.0);
  std::vector<double> phi((size_t)NA * npts);
  for (size_t p = 0; p < npts; ++p) {
    const double env = (p % 10) ? 1.0 : std::exp(-14.0 * (double)((p * 2654435761ull) % 1000003ull) / 1000003.0);
    for (int t = 0; t < NA; ++t) phi[(size_t)t * npts + p] = env * 0.5 * ud(rng);
  }
  std::vector<double> D((size_t)M * M);
  for (int i = 0; i < M; ++i) for (int j = i; j < M; ++j) { const double v = 0.05 * nd(rng) * std::exp(-2.0 * std::fabs(ud(rng))) + (i == j ? 0.3 : 0.0); D[(size_t)i * M + j] = D[(size_t)j * M + i] = v; }
  std::vector<uint16_t> tu(M);
  { int I = 0; for (int t = 0; t < NA; ++t) for (int u = t; u < NA; ++u) tu[I++] = (uint16_t)(t | (u << 8)); }
  std::vector<uint8_t> timg; int t_exp = 0;
  ordm::pack_t_int8_stream(M, D, KP, timg, t_exp);
  double g1, g2; ordm::int8_guard_constants(M, D, t_exp, 4, g1, g2);
  double *dphi, *dpi; uint8_t* dt; int* dfail;
  CK(cudaMalloc(&dphi, phi.size() * 8)); CK(cudaMalloc(&dpi, npts * 8)); CK(cudaMalloc(&dt, timg.size())); CK(cudaMalloc(&dfail, 8));
  CK(cudaMemcpy(dphi, phi.data(), phi.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt, timg.data(), timg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dfail, 0, 8)); CK(cudaMemset(dpi, 0, npts * 8));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const size_t smem = k2i::StreamSmem::bytes(KP);
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k2i::k2i_ontop_stream<NA, 4>));
  printf("NA %d M %d KP %d: regs %d, local %zu B, smem %zu B dyn + %zu static (opt-in max %zu), X units %d smem + %d TMEM, T image %zu B, t_exp %d\n", NA, M, KP, fa.numRegs,
         fa.localSizeBytes, smem, fa.sharedSizeBytes, (size_t)prop.sharedMemPerBlockOptin, k2i::StreamGeom::smem_units(KP), k2i::StreamGeom::tmem_units(KP), timg.size(), t_exp);
  CK(cudaFuncSetAttribute(k2i::k2i_ontop_stream<NA, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k2i::Params P{dphi, npts, npts, dt, t_exp, nullptr, dpi, dfail, g1, g2};
  int grid = std::max(2, (int)std::min<size_t>((size_t)prop.multiProcessorCount, 2 * ((npts + 255) / 256)) / 2 * 2);
  if (argc > 3) grid = std::max(2, atoi(argv[3]) / 2 * 2);
  auto launch = [&]() { k2i::k2i_ontop_stream<NA, 4><<<grid, k2i::STREAM_THREADS, smem>>>(P); };
  launch();
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(dfail, 0, 8));
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
  }
  std::vector<double> pi(npts); int fail[2];
  CK(cudaMemcpy(pi.data(), dpi, npts * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(fail, dfail, 8, cudaMemcpyDeviceToHost));
  double emax = 0.0, pimax = 0.0; size_t worst = 0, nbad = 0;
  std::vector<long double> x(M);
  for (size_t c = 0; c < ncheck; ++c) {
    const size_t p = (c * 7919) % npts;
    for (int I = 0; I < M; ++I) x[I] = (long double)(phi[(size_t)(tu[I] & 255) * npts + p] * phi[(size_t)(tu[I] >> 8) * npts + p]);
    long double ref = 0.0L;
    for (int I = 0; I < M; ++I) { long double y = 0.0L; for (int J = 0; J < M; ++J) y += x[J] * (long double)D[(size_t)I * M + J]; ref += x[I] * y; }
    const double sc = std::max(1.0, (double)fabsl(ref));
    const double e = std::fabs((double)((long double)pi[p] - ref)) / sc;
    if (e > emax) { emax = e; worst = p; }
    if (e > 1e-12) ++nbad;
    pimax = std::max(pimax, (double)fabsl(ref));
  }
  { long long pf[3][8]; CK(cudaMemcpyFromSymbol(pf, k2i::k2s_prof, sizeof(pf)));
    const double nt = (double)((npts + 255) / 256 + grid / 2 - 1) / (grid / 2);   // laps of the last launch
    const char* who[3] = {"compute thread 0", "MMA lane", "producer lane"};
    const char* nm[3][8] = {{"phi load", "build", "cluster sync", "wait acc full", "fold", "partials+store", "", ""},
                            {"tile sync", "wait drain", "wait T full", "wait T peer", "issue", "commit", "", ""},
                            {"tile sync", "wait empty", "issue TMA", "", "", "", "", ""}};
    for (int w = 0; w < 3; ++w) { printf("%s, clk per tile:", who[w]); double tot = 0; for (int i = 0; i < 6; ++i) { if (nm[w][i][0]) printf(" %s %.0f;", nm[w][i], (double)pf[w][i] / nt); tot += (double)pf[w][i] / nt; } printf(" total %.0f\n", tot); } }
  const double flops = (double)npts * ((double)M * (M + 1) + 2.0 * M);
  printf("NA %d: npts %zu grid %d: %.3f ms -> %.1f M pts/s (FP64-equivalent %.1f TFLOP/s of the symmetric contraction); barrier timeouts %d, guard-flagged points %d\n", NA, npts, grid, best,
         npts / best / 1e3, flops / best / 1e9, fail[0], fail[1]);
  printf("max |dPi| / max(1,|Pi|) over %zu points: %.3e (worst point %zu: got %.17g; %zu above 1e-12); |Pi|max %.3g\n", ncheck, emax, worst, pi[worst], nbad, pimax);
  return fail[0] ? 2 : 0;
}
