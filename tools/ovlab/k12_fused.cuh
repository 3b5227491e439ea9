// k12_fused.cuh -- K1 (densities) fused underneath K2 (on-top pair density) in ONE persistent CTA.
//
// K2 is bound by the FP64 tensor pipe and leaves HBM idle; K1 is a pure HBM stream with ~5 % of
// K2's FP64 work.  Run as two kernels they cost t(K1) + t(K2); run as two co-resident kernels the
// density kernel's DFMAs starve behind the DMMA stream (the FP64 pipe arbiter serves the older /
// lower-numbered MMA warps first: measured 1.6 TB/s, tools/ovlab).  Here the density stream is a
// fourth warpgroup of the K2 CTA -- warps 12..15, the highest warp ids, which the sub-partition
// arbiter favours -- so its DFMAs slot into the gaps of the DMMA stream while its loads use the
// otherwise idle HBM bandwidth:
//
//   warps 0..7    MMA lanes      (k2::ws_cta, 160 registers)
//   warps 8..11   builders       (TMA, X build, Pi store; 56 registers)
//   warps 12..15  density warps  (k1::density_point on the CTA's own 32-point tiles, lane = point,
//                                 one tile per warp round-robin; 136 registers)
//
// The density warps are decoupled from the other two groups (no barrier between them after the
// prologue): they write rho, grad rho and the closed-form core part of Pi to HBM; K3 adds that
// core part to K2's active-space Pi.  Per-CTA partial sums of the integrated densities go to
// dp.partial[blockIdx.x][3] in a fixed order.
#pragma once
#include "../../openrdm_b200/csrc/k1_density.cuh"
#include "../../openrdm_b200/csrc/k2_ontop.cuh"

namespace k12 {

constexpr int THREADS = k2::WS_THREADS + 128;
constexpr int DWARPS = 4;

template <int NP, bool GGA>
__global__ void __launch_bounds__(THREADS, 1) k12_fused(const k2::Params prm, const k1::Params dp) {
  __shared__ double dred[DWARPS][3];
  k2::ws_cta<false, 2, 160>(prm);  // warps 0..11 run K2 to completion; warps 12..15 fall through
  if (threadIdx.x < k2::WS_THREADS) return;
  k2::reg_inc<136>();  // 8 x 160 (MMA) + 4 x 56 (builders) + 4 x 136 = 2048 warp-registers: the whole file
  const int lane = threadIdx.x & 31, dw = (threadIdx.x - k2::WS_THREADS) >> 5;
  const int ntiles = prm.ntiles;
  const int nmine = (ntiles > (int)blockIdx.x) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int k = dw; k < nmine; k += DWARPS) {
    const size_t p = (size_t)(blockIdx.x + (size_t)k * gridDim.x) * k2::WS_P + lane;
    if (p < dp.npts) k1::density_point<NP, GGA>(dp, p, acc);
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double s = ordm::warp_sum(acc[j]);
    if (lane == 0) dred[dw][j] = s;
  }
  asm volatile("bar.sync 2, %0;" ::"n"(DWARPS * 32) : "memory");
  if (dw == 0 && lane < 3) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < DWARPS; ++w) s += dred[w][lane];
    dp.partial[(size_t)blockIdx.x * 3 + lane] = s;
  }
}

// Core-only variant: the density warps stream just the ncore doubly-occupied columns (80 % of K1's
// bytes for config 4) into the partial sums rc = sum phi_i^2, c{x,y,z} = sum phi_i d phi_i (exactly
// what k1::k1_core produces); the light active-space part of K1 then runs as its own kernel on those
// partials.  UNR orbitals x 4 arrays = 4*UNR independent 8-byte loads in flight per thread.
template <bool GGA, int UNR, int BW0 = k2::WARPS>
__global__ void __launch_bounds__(THREADS, 1) k12_fused_core(const k2::Params prm, const k1::Params dp) {
  k2::ws_cta<false, 2, 160, BW0>(prm);
  const int dw0 = BW0 == k2::WARPS ? 12 : 8;  // first density warp
  if ((int)(threadIdx.x >> 5) < dw0 || (int)(threadIdx.x >> 5) >= dw0 + 4) return;
  const int lane = threadIdx.x & 31, dw = (int)(threadIdx.x >> 5) - dw0;
  const int ntiles = prm.ntiles;
  const int nmine = (ntiles > (int)blockIdx.x) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const size_t ld = dp.ld;
  for (int k = dw; k < nmine; k += DWARPS) {
    const size_t p = (size_t)(blockIdx.x + (size_t)k * gridDim.x) * k2::WS_P + lane;
    if (p >= dp.npts) continue;
    double rc = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
    int i = 0;
    for (; i + UNR <= dp.ncore; i += UNR) {
      double v[UNR], vx[UNR], vy[UNR], vz[UNR];
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const size_t o = (size_t)(i + j) * ld + p;
        v[j] = __ldcs(dp.phi + o);
        if (GGA) { vx[j] = __ldcs(dp.phix + o); vy[j] = __ldcs(dp.phiy + o); vz[j] = __ldcs(dp.phiz + o); }
      }
#ifdef K12_EXP_NOFMA  // lab: same loads, integer accumulation (no FP64 pipe)
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        rc = __longlong_as_double(__double_as_longlong(rc) ^ __double_as_longlong(v[j]) ^ __double_as_longlong(vx[j]) ^ __double_as_longlong(vy[j]) ^ __double_as_longlong(vz[j]));
      }
#else
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        rc = fma(v[j], v[j], rc);
        if (GGA) { cx = fma(v[j], vx[j], cx); cy = fma(v[j], vy[j], cy); cz = fma(v[j], vz[j], cz); }
      }
#endif
    }
    for (; i < dp.ncore; ++i) {
      const size_t o = (size_t)i * ld + p;
      const double v = __ldcs(dp.phi + o);
      rc = fma(v, v, rc);
      if (GGA) { cx = fma(v, __ldcs(dp.phix + o), cx); cy = fma(v, __ldcs(dp.phiy + o), cy); cz = fma(v, __ldcs(dp.phiz + o), cz); }
    }
    dp.core_rc[p] = rc;
    if (GGA) { dp.core_cx[p] = cx; dp.core_cy[p] = cy; dp.core_cz[p] = cz; }
  }
}

}  // namespace k12
