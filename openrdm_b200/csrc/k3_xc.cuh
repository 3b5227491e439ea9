// k3_xc.cuh -- K3: R(r), density translation and the translated on-top functional, fused
// into one elementwise pass with the weighted quadrature partial sums.
// Replaces MCPDFT::build_R (mcpdft.cc:272-285), translate_density (294-344),
// translate_density_gradients (346-416) and Functional::EX_LSDA+EC_VWN3 | EX_PBE+EC_PBE
// (functional.cc:12-309) as dispatched by mcpdft_energy (energy.cc:110-122).
//
// Energy-only traffic: 24 B/point (rho, Pi, w) for tSVWN, 48 B/point (+ grad rho) for tPBE;
// with diagnostics on, R, rho~a, rho~b (and sigma~aa/ab/bb) are also written (+8 B each).
// No intermediate per-point vector is re-read: everything between the loads and the five
// partial sums stays in registers.
#pragma once
#include <cuda_runtime.h>
#include "reduce.cuh"
#include "xc_math.cuh"

namespace k3 {

constexpr int THREADS = 256;
#ifndef K3_MINB
#define K3_MINB 4  // 64 registers, 4 CTAs per SM: 0.294 -> 0.280 ms on config 4 (5 gave it back in spills)
#endif

struct Params {
  const double* rho;
  const double* pi;
  const double* pi_core;  // nullable: closed-form core part still to be added to pi (overlapped K1 || K2 mode)
  double* pi_out;         // nullable: where the completed Pi is written back in that mode
  const double* w;
  const double* gx;  // total density gradient (GGA only)
  const double* gy;
  const double* gz;
  const double* px;  // grad Pi (fully-translated GGA only)
  const double* py;
  const double* pz;
  const double* R_in;  // nullable: R installed by the caller (MCPDFT::set_R) instead of build_R's 4 Pi / rho^2; may alias R
  size_t npts;
  double omega;      // MCPDFT_OMEGA for FUNC 5 (WPBE)
  // optional diagnostics (null = not written)
  double* R;
  double* tra;
  double* trb;
  double* tsaa;
  double* tsab;
  double* tsbb;
  double* partial;  // [gridDim.x][5] : trN, trNa, trNb, Ex, Ec
};

// FUNC: 0 tSVWN, 1 tPBE (the two libfunc choices, energy.cc:114-120), 2 trevPBE, 3 tBLYP, 4 tBOP, 5 twPBE
//       (the Psi4 plugin's, mcpdft_solver.cc:742-787).  Everything but SVWN consumes sigma.
// FT:   MCPDFT::fully_translate (mcpdft.cc:418-606) instead of MCPDFT::translate.
template <int FUNC, bool FT = false>
__global__ void __launch_bounds__(THREADS, K3_MINB) k3_translate_xc(const Params P) {
  constexpr bool GGA = FUNC != 0;
  __shared__ double scratch[5 * 32];
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (size_t p = (size_t)blockIdx.x * THREADS + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * THREADS) {
    const double rho = __ldg(P.rho + p), w = __ldg(P.w + p);
    double pi = P.pi[p];
    if (P.pi_core) {  // same association as K2's own store: (sum over lanes) + core part
      pi += __ldg(P.pi_core + p);
      if (P.pi_out) P.pi_out[p] = pi;
    }
    double gx = 0.0, gy = 0.0, gz = 0.0;
    if (GGA && P.gx) { gx = __ldg(P.gx + p); gy = __ldg(P.gy + p); gz = __ldg(P.gz + p); }
    const double R = P.R_in ? P.R_in[p] : xc::r_factor(rho, pi);   // build_R, or the caller's set_R
    xc::Translated t;
    if (FT) {
      double px = 0.0, py = 0.0, pz = 0.0;
      if (GGA && P.px) { px = __ldg(P.px + p); py = __ldg(P.py + p); pz = __ldg(P.pz + p); }
      t = xc::fully_translate_pt<GGA>(rho, pi, gx, gy, gz, px, py, pz, R);
    } else {
      t = xc::translate_pt<GGA>(rho, pi, gx, gy, gz, R);
    }
    double ex, ec;
    if (FUNC == 1) {
      xc::pbe_pair_pt(t.ra, t.rb, t.saa, t.sab, t.sbb, xc::PBE_KAPPA, ex, ec);
    } else if (FUNC == 2) {
      xc::pbe_pair_pt(t.ra, t.rb, t.saa, t.sab, t.sbb, xc::REVPBE_KAPPA, ex, ec);
    } else if (FUNC == 3) {
      ex = xc::ex_b88_pt(t.ra, t.rb, t.saa, t.sbb);
      ec = xc::ec_lyp_pt(t.ra, t.rb, t.saa, t.sab, t.sbb);
    } else if (FUNC == 4) {
      ex = xc::ex_b88_pt(t.ra, t.rb, t.saa, t.sbb);
      ec = xc::ec_b88_op_pt(t.ra, t.rb, t.saa, t.sbb);
    } else if (FUNC == 5) {
      ex = xc::ex_wpbe_pt(t.ra, t.rb, t.saa, t.sbb, P.omega);
      ec = xc::ec_pbe_pt(t.ra, t.rb, t.saa, t.sab, t.sbb);
    } else {
      ex = xc::ex_lsda_pt(t.ra, t.rb);
      ec = xc::ec_vwn3_pt(t.ra, t.rb);
    }
    acc[0] = fma(t.ra + t.rb, w, acc[0]);  // mcpdft.cc:332-334
    acc[1] = fma(t.ra, w, acc[1]);
    acc[2] = fma(t.rb, w, acc[2]);
    acc[3] = fma(ex, w, acc[3]);
    acc[4] = fma(ec, w, acc[4]);
    if (P.R) P.R[p] = t.R;
    if (P.tra) P.tra[p] = t.ra;
    if (P.trb) P.trb[p] = t.rb;
    if (GGA) {
      if (P.tsaa) P.tsaa[p] = t.saa;
      if (P.tsab) P.tsab[p] = t.sab;
      if (P.tsbb) P.tsbb[p] = t.sbb;
    }
  }
  ordm::block_sum_store<5>(acc, scratch, P.partial + (size_t)blockIdx.x * 5);
}

// libfunc's four entry points (functional.h:17-43) and the Psi4 plugin's four (ordm_xc_term 4-9) on caller-provided vectors:
// one weighted sum per call, partial[blockIdx.x] = sum over this block's points.
struct FuncParams {
  const double* ra;
  const double* rb;
  const double* saa;
  const double* sab;
  const double* sbb;
  const double* w;
  size_t npts;
  double* partial;  // [gridDim.x][1]
  double omega;     // TERM 8 only
};

template <int TERM>
__global__ void __launch_bounds__(THREADS) k3_functional(const FuncParams P) {
  __shared__ double scratch[32];
  double acc[1] = {0.0};
  for (size_t p = (size_t)blockIdx.x * THREADS + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * THREADS) {
    const double ra = __ldg(P.ra + p), rb = __ldg(P.rb + p);
    double e;
    if (TERM == 0) e = xc::ex_lsda_pt(ra, rb);
    else if (TERM == 1) e = xc::ec_vwn3_pt(ra, rb);
    else if (TERM == 2) e = xc::ex_pbe_pt(ra, rb, __ldg(P.saa + p), __ldg(P.sbb + p));
    else if (TERM == 3) e = xc::ec_pbe_pt(ra, rb, __ldg(P.saa + p), __ldg(P.sab + p), __ldg(P.sbb + p));
    else if (TERM == 4) e = xc::ex_pbe_pt(ra, rb, __ldg(P.saa + p), __ldg(P.sbb + p), xc::REVPBE_KAPPA);
    else if (TERM == 5) e = xc::ex_b88_pt(ra, rb, __ldg(P.saa + p), __ldg(P.sbb + p));
    else if (TERM == 6) e = xc::ec_lyp_pt(ra, rb, __ldg(P.saa + p), __ldg(P.sab + p), __ldg(P.sbb + p));
    else if (TERM == 7) e = xc::ec_b88_op_pt(ra, rb, __ldg(P.saa + p), __ldg(P.sbb + p));
    else if (TERM == 8) e = xc::ex_wpbe_pt(ra, rb, __ldg(P.saa + p), __ldg(P.sbb + p), P.omega);
    else e = xc::ec_pw92_pt(ra, rb);
    acc[0] = fma(e, __ldg(P.w + p), acc[0]);
  }
  ordm::block_sum_store<1>(acc, scratch, P.partial + blockIdx.x);
}

}  // namespace k3
