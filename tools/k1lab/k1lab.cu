// k1lab -- lab harness for K1's active-space part: k1::k1_density on the core partials (the round-1 kernel) against
// k1t::k1_tile (TMA-staged tiles, sum / difference roles).  Synthetic data, parity of every output and event times.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I../../include -I../../openrdm_b200/csrc -I. -o k1lab k1lab.cu
//   (add -DK1LAB_FUSED -o k1lab_fused for the fused-XC lab variant; K1T_FUNC / K1T_GGA / K1T_SLOTS in the environment)
//   ./k1lab [npts] [nact] [nlead] [reps]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
// -DK1LAB_FUSED: the lab variant with K3 fused behind the density (k1_tile_fused.cuh, 13 warps) instead of the product kernel
#ifdef K1LAB_FUSED
#include "k1_tile_fused.cuh"
#else
#include "k1_tile.cuh"
#endif
#include "k3_xc.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void fill(double* p, size_t n, unsigned long long seed, double scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long x = (i + 1) * 0x9E3779B97F4A7C15ull ^ seed;
    x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29; x *= 0x94D049BB133111EBull; x ^= x >> 32;
    p[i] = scale * ((double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5);
  }
}

static bool g_gga = true;
template <int NP>
void launch_tile(const k1t::Maps& maps, const k1t::Params& P, int grid, size_t smem) {
#ifdef K1LAB_FUSED
  if (g_gga) {
    CK(cudaFuncSetAttribute(k1t::k1_tile<NP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k1t::k1_tile<NP, true><<<grid, k1t::THREADS, smem>>>(maps, P);
  } else {
    CK(cudaFuncSetAttribute(k1t::k1_tile<NP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k1t::k1_tile<NP, false><<<grid, k1t::THREADS, smem>>>(maps, P);
  }
#else
  CK(cudaFuncSetAttribute(k1t::k1_tile<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k1t::k1_tile<NP><<<grid, k1t::THREADS, smem>>>(maps, P);
#endif
}
template <int NP>
void launch_old(const k1::Params& p, int grid) { if (g_gga) k1::k1_density<NP, true><<<grid, k1::THREADS>>>(p); else k1::k1_density<NP, false><<<grid, k1::THREADS>>>(p); }

int main(int argc, char** argv) {
  const size_t npts = argc > 1 ? strtoull(argv[1], 0, 10) : 2000000;
  const int nact = argc > 2 ? atoi(argv[2]) : 20;
  const int nlead = argc > 3 ? atoi(argv[3]) : 5;
  const int reps = argc > 4 ? atoi(argv[4]) : 5;
  g_gga = !(getenv("K1T_GGA") && atoi(getenv("K1T_GGA")) == 0);
  const int c0 = 3, ncore = c0 + nlead, n = ncore + nact, ncols = nlead + nact;
  const size_t ld = (npts + 255) / 256 * 256;
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* phi[4];
  for (int a = 0; a < 4; ++a) { CK(cudaMalloc(&phi[a], ld * n * 8)); CK(cudaMemset(phi[a], 0, ld * n * 8)); }
  for (int a = 0; a < 4; ++a)
    for (int c = 0; c < n; ++c) fill<<<592, 256>>>(phi[a] + (size_t)c * ld, npts, 1234 + 100 * a + c, a ? 0.3 : 1.0);
  double *w, *core[4], *out_old[5], *out_new[5], *part_old, *part_new;
  CK(cudaMalloc(&w, ld * 8)); CK(cudaMemset(w, 0, ld * 8)); fill<<<592, 256>>>(w, npts, 77, 1e-3);
  for (int i = 0; i < 4; ++i) { CK(cudaMalloc(&core[i], ld * 8)); CK(cudaMemset(core[i], 0, ld * 8)); fill<<<592, 256>>>(core[i], npts, 900 + i, i ? 0.5 : 2.0); }
  for (int i = 0; i < 5; ++i) { CK(cudaMalloc(&out_old[i], ld * 8)); CK(cudaMalloc(&out_new[i], ld * 8)); CK(cudaMemset(out_old[i], 0, ld * 8)); CK(cudaMemset(out_new[i], 0, ld * 8)); }
  CK(cudaMalloc(&part_old, sms * 16 * 3 * 8)); CK(cudaMalloc(&part_new, sms * 16 * 3 * 8));
  // symmetric 1-RDM-like matrices
  std::vector<double> Da(k1::MAX_ACT * k1::MAX_ACT, 0.0), Db(Da), D0(Da), D1(Da);
  srand(5);
  for (int t = 0; t < nact; ++t)
    for (int u = t; u < nact; ++u) {
      const double a = (rand() / (double)RAND_MAX - 0.5) * (t == u ? 2.0 : 0.4), b = (rand() / (double)RAND_MAX - 0.5) * (t == u ? 2.0 : 0.4);
      Da[t * k1::MAX_ACT + u] = Da[u * k1::MAX_ACT + t] = a;
      Db[t * k1::MAX_ACT + u] = Db[u * k1::MAX_ACT + t] = b;
    }
  for (int t = 0; t < nact; ++t)
    for (int u = 0; u < nact; ++u) {
      D0[t * k1::MAX_ACT + u] = Da[t * k1::MAX_ACT + u] + Db[t * k1::MAX_ACT + u];
      const double d = Da[t * k1::MAX_ACT + u] - Db[t * k1::MAX_ACT + u];
      D1[u * k1::MAX_ACT + t] = d;
    }
  CK(cudaMemcpyToSymbol(k1::c_Dsa, Da.data(), Da.size() * 8));
  CK(cudaMemcpyToSymbol(k1::c_Dsb, Db.data(), Db.size() * 8));
  double* dmat;
  CK(cudaMalloc(&dmat, 2 * D0.size() * 8));
  CK(cudaMemcpy(dmat, D0.data(), D0.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dmat + D0.size(), D1.data(), D1.size() * 8, cudaMemcpyHostToDevice));

  // ---- old kernel
  k1::Params po{};
  po.phi = phi[0]; po.phix = phi[1]; po.phiy = phi[2]; po.phiz = phi[3]; po.w = w; po.ld = ld; po.npts = npts; po.ncore = ncore; po.nact = nact;
  po.rho = out_old[0]; po.gx = out_old[1]; po.gy = out_old[2]; po.gz = out_old[3]; po.pi_core = out_old[4];
  po.partial = part_old; po.core_rc = core[0]; po.core_cx = core[1]; po.core_cy = core[2]; po.core_cz = core[3]; po.core_done = c0;
  const int grid_old = (int)std::min<size_t>((npts + 127) / 128, (size_t)sms * 16);
  // ---- new kernel
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
  if (!enc || qres != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  k1t::Maps maps;
  for (int a = 0; a < 4; ++a) {
    cuuint64_t gdim[2] = {(cuuint64_t)npts, (cuuint64_t)ncols}, gstr[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)k1t::TILE, (cuuint32_t)ncols}, estr[2] = {1, 1};
    CUresult r = enc(&maps.m[a], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, phi[a] + (size_t)c0 * ld, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  k1t::Params pn{};
  pn.w = w; pn.core_rc = core[0]; pn.core_cx = core[1]; pn.core_cy = core[2]; pn.core_cz = core[3];
  pn.npts = npts; pn.ntiles = (int)((npts + k1t::TILE - 1) / k1t::TILE); pn.nlead = nlead; pn.nact = nact; pn.have_diff = 1;
  pn.dmat = dmat; pn.rho = out_new[0]; pn.gx = out_new[1]; pn.gy = out_new[2]; pn.gz = out_new[3]; pn.pi_core = out_new[4]; pn.partial = part_new;
  const int func = getenv("K1T_FUNC") ? atoi(getenv("K1T_FUNC")) : 1;
  const int np = (nact + 3) / 4 * 4;
#ifdef K1LAB_FUSED
  int nslots = (int)((232448 - 16 * 28 - 3072 - (k1t::EXCH_DOUBLES + k1t::dmat_doubles(np)) * 8) / (k1t::slot_doubles(ncols, g_gga) * 8));
#else
  if (!g_gga) { printf("the product kernel serves GGA layouts only (build with -DK1LAB_FUSED for the lab variant)\n"); return 1; }
  int nslots = (int)((232448 - 16 * 12 - 2048 - k1t::dmat_doubles(np) * 8) / (k1t::slot_doubles(ncols) * 8));
#endif
  if (getenv("K1T_SLOTS")) nslots = atoi(getenv("K1T_SLOTS"));
  nslots = std::min(nslots, 12);
  pn.nslots = nslots;
#ifdef K1LAB_FUSED
  const size_t smem = k1t::smem_bytes(ncols, g_gga, nslots, np);
#else
  const size_t smem = k1t::smem_bytes(ncols, nslots, np);
#endif
  const int grid_new = std::min(pn.ntiles, sms);
  printf("npts %zu nact %d nlead %d: old grid %d, new grid %d, slots %d, smem %zu, functional %d\n", npts, nact, nlead, grid_old, grid_new, nslots, smem, func);
  // Pi_act and the K3 of the separate path
  double *pi_in, *pi_old, *pi_new, *part3_old, *part3_new;
  CK(cudaMalloc(&pi_in, ld * 8)); CK(cudaMalloc(&pi_old, ld * 8)); CK(cudaMalloc(&pi_new, ld * 8));
  CK(cudaMemset(pi_in, 0, ld * 8));
  fill<<<592, 256>>>(pi_in, npts, 4242, 0.5);   // both signs: the pi < 0 branch of the translation is hit
  CK(cudaMalloc(&part3_old, sms * 16 * 5 * 8)); CK(cudaMalloc(&part3_new, sms * 16 * 5 * 8));
  k3::Params p3{};
  p3.rho = out_old[0]; p3.pi = pi_old; p3.pi_core = out_old[4]; p3.pi_out = pi_old; p3.w = w; p3.gx = g_gga ? out_old[1] : nullptr; p3.gy = g_gga ? out_old[2] : nullptr; p3.gz = g_gga ? out_old[3] : nullptr;
  p3.npts = npts; p3.omega = 0.3; p3.partial = part3_old;
  const int grid3 = (int)std::min<size_t>((npts + k3::THREADS - 1) / k3::THREADS, (size_t)sms * 8);
  auto run_k3 = [&]() {
    switch (func) {
      case 0: k3::k3_translate_xc<0, false><<<grid3, k3::THREADS>>>(p3); break;
      case 1: k3::k3_translate_xc<1, false><<<grid3, k3::THREADS>>>(p3); break;
      case 3: k3::k3_translate_xc<3, false><<<grid3, k3::THREADS>>>(p3); break;
      default: k3::k3_translate_xc<5, false><<<grid3, k3::THREADS>>>(p3); break;
    }
  };
#ifdef K1LAB_FUSED
  pn.xc = -1; pn.pi = pi_new; pn.partial3 = part3_new; pn.omega = 0.3;
#endif
  auto run_old = [&]() { switch (np) { case 8: launch_old<8>(po, grid_old); break; case 20: launch_old<20>(po, grid_old); break; case 28: launch_old<28>(po, grid_old); break; default: launch_old<24>(po, grid_old); } };
  auto run_new = [&]() { switch (np) { case 8: launch_tile<8>(maps, pn, grid_new, smem); break; case 20: launch_tile<20>(maps, pn, grid_new, smem); break; case 28: launch_tile<28>(maps, pn, grid_new, smem); break; default: launch_tile<24>(maps, pn, grid_new, smem); } };
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const char* what[4] = {"k1_density        ", "k1_density + k3   ", "k1_tile           ", "k1_tile, fused XC "};
#ifdef K1LAB_FUSED
  const int nwhich = 4;
#else
  const int nwhich = 3;
#endif
  for (int which = 0; which < nwhich; ++which) {
    float best = 1e9f, sum = 0.f;
#ifdef K1LAB_FUSED
    pn.xc = which == 3 ? func : -1;
#endif
    for (int r = 0; r < reps + 1; ++r) {
      if (which < 2) CK(cudaMemcpy(pi_old, pi_in, ld * 8, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpy(pi_new, pi_in, ld * 8, cudaMemcpyDeviceToDevice));
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      if (which >= 2) run_new(); else { run_old(); if (which == 1) run_k3(); }
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r) { best = std::min(best, ms); sum += ms; }
    }
    const double bytes = (double)npts * (8.0 * (4 * ncols + 5) + 40.0);
    printf("%s: best %.4f ms, mean %.4f ms  (K1 part: %.0f GB/s on %.0f B/point)\n", what[which], best, sum / reps, bytes / best * 1e-6, bytes / npts);
    if (which == 2 || which == 3) {
      // ---- parity against the separate kernels (which ran last in which = 1)
      std::vector<double> a(npts), b(npts);
      const char* names[6] = {"rho", "gx", "gy", "gz", "pi_core", "pi"};
      for (int i = 0; i < 6; ++i) {
        if (which == 3 && i == 4) continue;
        if (which == 2 && i == 5) continue;
        CK(cudaMemcpy(a.data(), i == 5 ? pi_old : out_old[i], npts * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), i == 5 ? pi_new : out_new[i], npts * 8, cudaMemcpyDeviceToHost));
        double md = 0.0, ma = 0.0;
        for (size_t p = 0; p < npts; ++p) { md = std::max(md, fabs(a[p] - b[p])); ma = std::max(ma, fabs(a[p])); }
        printf("  %-8s max |old - new| = %.3e (max |old| = %.3e)\n", names[i], md, ma);
      }
      std::vector<double> po_h(grid_old * 3), pn_h(grid_new * 3);
      CK(cudaMemcpy(po_h.data(), part_old, po_h.size() * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(pn_h.data(), part_new, pn_h.size() * 8, cudaMemcpyDeviceToHost));
      for (int k = 0; k < 3; ++k) {
        long double so = 0, sn = 0;
        for (int r = 0; r < grid_old; ++r) so += po_h[r * 3 + k];
        for (int r = 0; r < grid_new; ++r) sn += pn_h[r * 3 + k];
        printf("  sum[%d]: old %.15Le new %.15Le diff %.3Le\n", k, so, sn, so - sn);
      }
      if (which == 3) {
        std::vector<double> qo(grid3 * 5), qn(grid_new * 5);
        CK(cudaMemcpy(qo.data(), part3_old, qo.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(qn.data(), part3_new, qn.size() * 8, cudaMemcpyDeviceToHost));
        for (int k = 0; k < 5; ++k) {
          long double so = 0, sn = 0;
          for (int r = 0; r < grid3; ++r) so += qo[r * 5 + k];
          for (int r = 0; r < grid_new; ++r) sn += qn[r * 5 + k];
          printf("  xc sum[%d]: old %.15Le new %.15Le diff %.3Le\n", k, so, sn, so - sn);
        }
      }
    }
  }
  return 0;
}
