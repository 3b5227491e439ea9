// lab harness for k2i_ontop_pair (csrc/k2_int8_pair.cuh): random active-space orbitals and a dense symmetric D~,
// CPU reference in long double on a sample of points, timing with CUDA events.
#define K2I_PROF 1
#include "../../openrdm_b200/csrc/k2_int8_pair.cuh"
#include "../../openrdm_b200/csrc/rdm_pack.h"
#include "../../openrdm_b200/csrc/k1_density.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#ifndef LAB_EW
#define LAB_EW 3
#endif
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// lab: the core-orbital stream's loads without its FP64 arithmetic (the values are xor-ed together): separates what the
// stream costs the kernel beside it through HBM / L2 / issue slots from what its DFMAs cost
__global__ void __launch_bounds__(128, 4) core_stream_nofp64(const k1::Params P) {
  const size_t ld = P.ld;
  for (size_t p = (size_t)blockIdx.x * 128 + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * 128) {
    unsigned long long a = 0, b = 0, c = 0, d = 0;
    int i = 0;
    for (; i + 8 <= P.ncore; i += 8) {
      double v[8], vx[8], vy[8], vz[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const size_t o = (size_t)(i + j) * ld + p;
        v[j] = __ldcs(P.phi + o); vx[j] = __ldcs(P.phix + o); vy[j] = __ldcs(P.phiy + o); vz[j] = __ldcs(P.phiz + o);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { a ^= __double_as_longlong(v[j]); b ^= __double_as_longlong(vx[j]); c ^= __double_as_longlong(vy[j]); d ^= __double_as_longlong(vz[j]); }
    }
    P.core_rc[p] = __longlong_as_double(a ^ b ^ c ^ d);
  }
}

int main(int argc, char** argv) {
  const size_t npts = argc > 1 ? (size_t)atoll(argv[1]) : 2000000;
  const size_t ncheck = std::min(npts, argc > 2 ? (size_t)atoll(argv[2]) : (size_t)2048);
  const int gmin = argc > 3 ? atoi(argv[3]) : 4;
  constexpr int NA = 20, M = NA * (NA + 1) / 2, KP = 224, EW = LAB_EW;
  std::mt19937_64 rng(42);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::uniform_real_distribution<double> ud(-1.0, 1.0);
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
  ordm::pack_t_int8_pair(M, D, KP, timg, t_exp);
  double g1, g2; ordm::int8_guard_constants(M, D, t_exp, gmin, g1, g2);
  double *dphi, *dpi; uint8_t* dt; int* dfail;
  CK(cudaMalloc(&dphi, phi.size() * 8)); CK(cudaMalloc(&dpi, npts * 8)); CK(cudaMalloc(&dt, timg.size())); CK(cudaMalloc(&dfail, 8));
  CK(cudaMemcpy(dphi, phi.data(), phi.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt, timg.data(), timg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dfail, 0, 8)); CK(cudaMemset(dpi, 0, npts * 8));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const size_t smem = k2i::PairSmem::bytes(KP);
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k2i::k2i_ontop_pair<NA, 4, EW>));
  printf("EW %d maxreg %d: regs %d, local %zu B, smem %zu B dyn + %zu static (opt-in max %zu), t_exp %d g1 %.3e g2 %.3e\n", EW, K2P_MAXREG, fa.numRegs, fa.localSizeBytes, smem, fa.sharedSizeBytes, (size_t)prop.sharedMemPerBlockOptin, t_exp, g1, g2);
  CK(cudaFuncSetAttribute(k2i::k2i_ontop_pair<NA, 4, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k2i::k2i_ontop_pair<NA, 5, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k2i::Params P{dphi, npts, npts, dt, t_exp, nullptr, dpi, dfail, g1, g2};
  const int threads = (4 * EW + 1) * 32;
  int grid = std::max(2, (int)std::min<size_t>((size_t)prop.multiProcessorCount, 2 * ((npts + 255) / 256)) / 2 * 2);
  if (argc > 4) grid = std::max(2, atoi(argv[4]) / 2 * 2);   // fewer CTA pairs: per-tile clocks with most of the chip idle
  // LAB_CORE=<ncore>: the product's core-orbital stream (k1::k1_core_light<true, 8>, one CTA per SM, launched first on a
  // second stream) beside the kernel, as ordm_energy runs it; LAB_CORE_MODE=1: the same loads without the DFMAs' results
  // mattering is not possible here, so mode 1 = half as many core orbitals (half the bytes and DFMAs)
  const int lab_core = getenv("LAB_CORE") ? atoi(getenv("LAB_CORE")) : 0;
  const int core_unr = getenv("LAB_CORE_UNR") ? atoi(getenv("LAB_CORE_UNR")) : 8, core_ctas = getenv("LAB_CORE_CTAS") ? atoi(getenv("LAB_CORE_CTAS")) : 1;
  cudaStream_t s2 = nullptr;
  k1::Params kp{};
  cudaEvent_t c0 = nullptr, c1 = nullptr;
  if (lab_core > 0) {
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    double* core; CK(cudaMalloc(&core, (size_t)4 * lab_core * npts * 8 + 4 * npts * 8));
    CK(cudaMemset(core, 0, (size_t)4 * lab_core * npts * 8));
    kp.phi = core; kp.phix = core + (size_t)lab_core * npts; kp.phiy = core + (size_t)2 * lab_core * npts; kp.phiz = core + (size_t)3 * lab_core * npts;
    kp.ld = npts; kp.npts = npts; kp.ncore = lab_core; kp.nact = 0;
    kp.core_rc = core + (size_t)4 * lab_core * npts; kp.core_cx = kp.core_rc + npts; kp.core_cy = kp.core_cx + npts; kp.core_cz = kp.core_cy + npts;
    CK(cudaFuncSetAttribute(k1::k1_core_light<true, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(k1::k1_core_light<true, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(k1::k1_core_light<true, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(core_stream_nofp64, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaEventCreate(&c0)); CK(cudaEventCreate(&c1));
  }
  auto launch = [&]() {
    if (lab_core > 0) {
      CK(cudaEventRecord(c0, s2));
      if (core_unr == 0) core_stream_nofp64<<<prop.multiProcessorCount * core_ctas, 128, 0, s2>>>(kp);
      else if (core_unr == 8) k1::k1_core_light<true, 8><<<prop.multiProcessorCount * core_ctas, k1::LIGHT_THREADS, 0, s2>>>(kp);
      else if (core_unr == 4) k1::k1_core_light<true, 4><<<prop.multiProcessorCount * core_ctas, k1::LIGHT_THREADS, 0, s2>>>(kp);
      else k1::k1_core_light<true, 2><<<prop.multiProcessorCount * core_ctas, k1::LIGHT_THREADS, 0, s2>>>(kp);
      CK(cudaEventRecord(c1, s2));
    }
    if (gmin == 4) k2i::k2i_ontop_pair<NA, 4, EW><<<grid, threads, smem>>>(P); else k2i::k2i_ontop_pair<NA, 5, EW><<<grid, threads, smem>>>(P); };
  launch();
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(dfail, 0, 8));
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
    if (lab_core > 0) { float mc; CK(cudaEventElapsedTime(&mc, c0, c1)); printf("  rep %d: K2i %.3f ms beside the core stream of %d orbitals (%.3f ms, %.0f GB/s)\n", rep, ms, lab_core, mc, 32.0 * lab_core * npts / mc / 1e6); }
  }
  std::vector<double> pi(npts); int fail[2];
  CK(cudaMemcpy(pi.data(), dpi, npts * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(fail, dfail, 8, cudaMemcpyDeviceToHost));
  double emax = 0.0, emax_f64 = 0.0, pimax = 0.0; size_t worst = 0, nbad = 0;
  std::vector<long double> x(M);
  for (size_t c = 0; c < ncheck; ++c) {
    const size_t p = (c * 7919) % npts;
    for (int I = 0; I < M; ++I) x[I] = (long double)(phi[(size_t)(tu[I] & 255) * npts + p] * phi[(size_t)(tu[I] >> 8) * npts + p]);
    long double ref = 0.0L; double f64 = 0.0;
    for (int I = 0; I < M; ++I) { long double y = 0.0L; double yd = 0.0; for (int J = 0; J < M; ++J) { y += x[J] * (long double)D[(size_t)I * M + J]; yd += (double)x[J] * D[(size_t)I * M + J]; } ref += x[I] * y; f64 += (double)x[I] * yd; }
    const double sc = std::max(1.0, (double)fabsl(ref));
    const double e = std::fabs((double)((long double)pi[p] - ref)) / sc;
    if (e > emax) { emax = e; worst = p; }
    if (e > 1e-12) ++nbad;
    emax_f64 = std::max(emax_f64, std::fabs((double)((long double)f64 - ref)) / sc); pimax = std::max(pimax, (double)fabsl(ref));
  }
  { long long pf[3][8]; CK(cudaMemcpyFromSymbol(pf, k2i::k2p_prof, sizeof(pf)));
    const double nt = (double)((npts + 255) / 256 + grid / 2 - 1) / (grid / 2);
    const char* who[3] = {"epilogue warp 0 (sub-partition of the MMA warp)", "epilogue warp 1 (another sub-partition)", "MMA lane"};
    const char* nm[8] = {"phi load", "build+cluster sync", "wait drain", "issue", "partials+store", "wait full", "fold", "tile head"};
    for (int w = 0; w < 3; ++w) { printf("%s, clk per tile:", who[w]); double tot = 0; for (int i = 0; i < 8; ++i) { printf(" %s %.0f;", nm[i], (double)pf[w][i] / nt); tot += (double)pf[w][i] / nt; } printf(" total %.0f\n", tot); } }
  const double flops = (double)npts * ((double)M * (M + 1) + 2.0 * M);
  printf("EW %d GMIN %d: npts %zu: %.3f ms -> %.1f M pts/s (FP64-equivalent %.1f TFLOP/s of the symmetric contraction); barrier timeouts %d, guard-flagged points %d\n", EW, gmin, npts, best, npts / best / 1e3, flops / best / 1e9, fail[0], fail[1]);
  printf("max |dPi| / max(1,|Pi|) over %zu points: int8 path %.3e (worst point %zu: got %.17g; %zu above 1e-12), plain FP64 loops %.3e; |Pi|max %.3g\n", ncheck, emax, worst, pi[worst], nbad, emax_f64, pimax);
  return fail[0] ? 2 : 0;
}
