// k1_density.cuh -- K1: spin densities, their gradients and sigma on the grid.
// Replaces MCPDFT::build_density_functions (mcpdft.cc:42-105) and build_density_gradients
// (mcpdft.cc:107-176).
//
// HBM layout: phi, phi_x, phi_y, phi_z are column-major [n_occ][ld] FP64 (point index fastest,
// exactly arma::mat(npts,nbfs)); one thread owns one grid point, so every load instruction of
// a warp reads 256 contiguous bytes of one orbital column.  The kernel is a pure stream over
// 8*n_occ*(1+3*GGA) bytes per point.
//
// Orbitals are [ncore | nact].  Core orbitals have D1a = D1b = 1 on the diagonal, so
//     rho_c = sum phi_i^2,      grad rho_c = 2 sum phi_i grad phi_i          (closed form)
// The active block uses the symmetrised 1-RDMs Ds = (D + D^T)/2 held in constant memory
// (a DFMA reads its coefficient straight from the constant bank, no load instruction):
//     T_s = Ds_s . phi_act(p),  rho_t,s = T_s . phi_act,  grad rho_t,s = 2 T_s . grad phi_act
// Orbital values are loaded with ld.global.cs (evict-first): they are read exactly once here, and
// when K1 runs underneath K2 they must not push K2's L2-resident Dt fragments out.
// The T-form reproduces mcpdft.cc:83-84 and :148-153 for arbitrary, non-symmetric D1
// (SURVEY.md Appendix B.2).  Also emitted for K2/K3: pi_core = rc^2 + rc*rtb + rta*rc
// (Appendix B.3) and the total gradient grad rho = grad rho_a + grad rho_b (mcpdft.cc:381-383).
#pragma once
#include <cuda_runtime.h>
#include "reduce.cuh"

namespace k1 {

constexpr int MAX_ACT = 32;      // active orbitals handled by the register-tiled kernel
constexpr int THREADS = 128;

__constant__ double c_Dsa[MAX_ACT * MAX_ACT];
__constant__ double c_Dsb[MAX_ACT * MAX_ACT];

struct Params {
  const double* phi;   // [n][ld]
  const double* phix;  // nullable when !GGA
  const double* phiy;
  const double* phiz;
  const double* w;
  size_t ld;
  size_t npts;
  int ncore, nact;
  // outputs (each npts long, any may be null)
  double* rho;      // total density
  double* gx;       // total gradient
  double* gy;
  double* gz;
  double* pi_core;  // closed-form core part of Pi (null when ncore == 0 => treated as 0)
  double* pcg[3];   // closed-form core part of grad Pi (null unless grad Pi is requested with core orbitals)
  double* rhoa;     // diagnostics
  double* rhob;
  double* grad[6];  // ax ay az bx by bz
  double* saa;
  double* sab;
  double* sbb;
  double* partial;  // [gridDim.x][3] : sum rho*w, rho_a*w, rho_b*w
  // generic path only: symmetrised active 1-RDMs in global memory
  const double* Dsa_g;
  const double* Dsb_g;
  // core partial sums produced by k1_core (rc = sum phi_i^2, c{x,y,z} = sum phi_i d phi_i); when
  // core_rc is set the kernels below skip their own core loop and read these instead
  int core_done;     // with core_rc set: the partials cover core orbitals [0, core_done); the density kernels sum [core_done, ncore) themselves
  double* core_rc;
  double* core_cx;
  double* core_cy;
  double* core_cz;
};

// K1a: the core-orbital stream.  80 % of K1's bytes for config 4 are core columns whose
// contribution is a plain sum of squares / products; running them in their own light-weight
// kernel (about 40 registers, 16 independent 8-byte loads in flight per thread, full occupancy)
// keeps that part of the stream at HBM speed instead of at the occupancy of the
// register-heavy active-space kernel.
constexpr int CORE_THREADS = 256;
template <bool GGA>
__global__ void __launch_bounds__(CORE_THREADS) k1_core(const Params P) {
  const size_t ld = P.ld;
  for (size_t p = (size_t)blockIdx.x * CORE_THREADS + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * CORE_THREADS) {
    double rc = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
    int i = 0;
    const int nc = P.core_done > 0 ? P.core_done : P.ncore;
    for (; i + 4 <= nc; i += 4) {
      double v[4], vx[4], vy[4], vz[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t o = (size_t)(i + j) * ld + p;
        v[j] = __ldcs(P.phi + o);
        if (GGA) { vx[j] = __ldcs(P.phix + o); vy[j] = __ldcs(P.phiy + o); vz[j] = __ldcs(P.phiz + o); }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rc = fma(v[j], v[j], rc);
        if (GGA) { cx = fma(v[j], vx[j], cx); cy = fma(v[j], vy[j], cy); cz = fma(v[j], vz[j], cz); }
      }
    }
    for (; i < nc; ++i) {
      const size_t o = (size_t)i * ld + p;
      const double v = __ldcs(P.phi + o);
      rc = fma(v, v, rc);
      if (GGA) {
        cx = fma(v, __ldcs(P.phix + o), cx);
        cy = fma(v, __ldcs(P.phiy + o), cy);
        cz = fma(v, __ldcs(P.phiz + o), cz);
      }
    }
    P.core_rc[p] = rc;
    if (GGA) { P.core_cx[p] = cx; P.core_cy[p] = cy; P.core_cz[p] = cz; }
  }
}

// The same core sums in a register-light shape (<= 64 registers, 128 threads, 8 loads in flight per
// thread) for running UNDERNEATH the tensor-bound K2: next to K2's register-trimmed CTA an SM has
// 16 K registers left, which hold two or three of these CTAs but only one of the 128-register full
// K1 -- and what limits a co-resident kernel is how often its warps catch a free FP64 slot between
// K2's DMMAs, so the FMA-light / byte-heavy core part (332 DFMAs for 2 656 bytes per point on config 4)
// is the part to run there: it finishes about when K2 does (8.6 vs 8.2 ms; the full K1: 10.1 ms).
// UNR = core orbitals per batch of loads.  UNR = 2: the shape described above (<= 64 registers).  UNR = 8 (<= 128
// registers, 32 eight-byte loads in flight per thread): for running underneath the int8 K2 (k2_int8.cuh), which
// leaves 17 K registers and no shared memory -- one such CTA per SM has to keep ~32 KB in flight on its own.
constexpr int LIGHT_THREADS = 128;
template <bool GGA, int UNR>
__global__ void __launch_bounds__(LIGHT_THREADS, UNR <= 2 ? 8 : 4) k1_core_light(const Params P) {
  const size_t ld = P.ld;
  for (size_t p = (size_t)blockIdx.x * LIGHT_THREADS + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * LIGHT_THREADS) {
    double rc = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
    int i = 0;
    const int nc = P.core_done > 0 ? P.core_done : P.ncore;
    for (; i + UNR <= nc; i += UNR) {
      double v[UNR], vx[UNR], vy[UNR], vz[UNR];
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const size_t o = (size_t)(i + j) * ld + p;
        v[j] = __ldcs(P.phi + o);
        if (GGA) { vx[j] = __ldcs(P.phix + o); vy[j] = __ldcs(P.phiy + o); vz[j] = __ldcs(P.phiz + o); }
      }
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        rc = fma(v[j], v[j], rc);
        if (GGA) { cx = fma(v[j], vx[j], cx); cy = fma(v[j], vy[j], cy); cz = fma(v[j], vz[j], cz); }
      }
    }
    for (; i < nc; ++i) {
      const size_t o = (size_t)i * ld + p;
      const double v = __ldcs(P.phi + o);
      rc = fma(v, v, rc);
      if (GGA) { cx = fma(v, __ldcs(P.phix + o), cx); cy = fma(v, __ldcs(P.phiy + o), cy); cz = fma(v, __ldcs(P.phiz + o), cz); }
    }
    P.core_rc[p] = rc;
    if (GGA) { P.core_cx[p] = cx; P.core_cy[p] = cy; P.core_cz[p] = cz; }
  }
}

// The deep shape with 16-byte loads: a thread owns two adjacent points, 4 orbitals x 4 arrays = 16 LDG.128 in flight (the
// same 256 bytes per thread as UNR = 8 above) for half as many load instructions -- beside the int8 K2, whose epilogue
// warps are bound by instruction issue, every instruction this kernel does not issue is one they can
// (profiles/r02_k2p_core_stream_fp64_vs_loads_only.log: a loads-only stream slows K2 as much as the full kernel does).
// Points are paired (p, p + 1) with p even; ld is a multiple of 256 and the arrays are ld long, so the odd tail is in bounds.
__device__ __forceinline__ double2 ldcs2(const double* p) {
  double2 v;
  asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
template <bool GGA>
__global__ void __launch_bounds__(LIGHT_THREADS, 4) k1_core_pairs(const Params P) {
  const size_t ld = P.ld;
  const int nc = P.core_done > 0 ? P.core_done : P.ncore;
  for (size_t p = 2 * ((size_t)blockIdx.x * LIGHT_THREADS + threadIdx.x); p < P.npts; p += 2 * (size_t)gridDim.x * LIGHT_THREADS) {
    double2 rc = {0.0, 0.0}, cx = {0.0, 0.0}, cy = {0.0, 0.0}, cz = {0.0, 0.0};
    int i = 0;
    for (; i + 4 <= nc; i += 4) {
      double2 v[4], vx[4], vy[4], vz[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t o = (size_t)(i + j) * ld + p;
        v[j] = ldcs2(P.phi + o);
        if (GGA) { vx[j] = ldcs2(P.phix + o); vy[j] = ldcs2(P.phiy + o); vz[j] = ldcs2(P.phiz + o); }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rc.x = fma(v[j].x, v[j].x, rc.x); rc.y = fma(v[j].y, v[j].y, rc.y);
        if (GGA) {
          cx.x = fma(v[j].x, vx[j].x, cx.x); cx.y = fma(v[j].y, vx[j].y, cx.y);
          cy.x = fma(v[j].x, vy[j].x, cy.x); cy.y = fma(v[j].y, vy[j].y, cy.y);
          cz.x = fma(v[j].x, vz[j].x, cz.x); cz.y = fma(v[j].y, vz[j].y, cz.y);
        }
      }
    }
    for (; i < nc; ++i) {
      const size_t o = (size_t)i * ld + p;
      const double2 v = ldcs2(P.phi + o);
      rc.x = fma(v.x, v.x, rc.x); rc.y = fma(v.y, v.y, rc.y);
      if (GGA) {
        const double2 a = ldcs2(P.phix + o), b = ldcs2(P.phiy + o), c = ldcs2(P.phiz + o);
        cx.x = fma(v.x, a.x, cx.x); cx.y = fma(v.y, a.y, cx.y);
        cy.x = fma(v.x, b.x, cy.x); cy.y = fma(v.y, b.y, cy.y);
        cz.x = fma(v.x, c.x, cz.x); cz.y = fma(v.y, c.y, cz.y);
      }
    }
    *reinterpret_cast<double2*>(P.core_rc + p) = rc;
    if (GGA) {
      *reinterpret_cast<double2*>(P.core_cx + p) = cx;
      *reinterpret_cast<double2*>(P.core_cy + p) = cy;
      *reinterpret_cast<double2*>(P.core_cz + p) = cz;
    }
  }
}

// ... and with 32-byte loads (LDG.E.256, new with sm_100): a thread owns four adjacent points, 2 orbitals x 4 arrays = 8 loads
// in flight (again 256 bytes per thread), a warp instruction covers 1 KB of one column.  Half the load instructions of
// k1_core_pairs for the same bytes and DFMAs.  Points are grouped (p .. p + 3) with p a multiple of 4; ld is a multiple of 256
// and the arrays are ld long and 32-byte aligned (checked by the launcher), so the tail group is in bounds.
struct double4v { double x, y, z, w; };
__device__ __forceinline__ double4v ldcs4(const double* p) {
  double4v v;
  asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stcg4(double* p, const double4v& v) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ void fma4(double4v& a, const double4v& x, const double4v& y) {
  a.x = fma(x.x, y.x, a.x); a.y = fma(x.y, y.y, a.y); a.z = fma(x.z, y.z, a.z); a.w = fma(x.w, y.w, a.w);
}
template <bool GGA>
__global__ void __launch_bounds__(LIGHT_THREADS, 4) k1_core_quads(const Params P) {
  const size_t ld = P.ld;
  const int nc = P.core_done > 0 ? P.core_done : P.ncore;
  for (size_t p = 4 * ((size_t)blockIdx.x * LIGHT_THREADS + threadIdx.x); p < P.npts; p += 4 * (size_t)gridDim.x * LIGHT_THREADS) {
    double4v rc = {0.0, 0.0, 0.0, 0.0}, cx = rc, cy = rc, cz = rc;
    int i = 0;
    for (; i + 2 <= nc; i += 2) {
      double4v v[2], vx[2], vy[2], vz[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const size_t o = (size_t)(i + j) * ld + p;
        v[j] = ldcs4(P.phi + o);
        if (GGA) { vx[j] = ldcs4(P.phix + o); vy[j] = ldcs4(P.phiy + o); vz[j] = ldcs4(P.phiz + o); }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        fma4(rc, v[j], v[j]);
        if (GGA) { fma4(cx, v[j], vx[j]); fma4(cy, v[j], vy[j]); fma4(cz, v[j], vz[j]); }
      }
    }
    for (; i < nc; ++i) {
      const size_t o = (size_t)i * ld + p;
      const double4v v = ldcs4(P.phi + o);
      fma4(rc, v, v);
      if (GGA) {
        const double4v a = ldcs4(P.phix + o), b = ldcs4(P.phiy + o), c = ldcs4(P.phiz + o);
        fma4(cx, v, a); fma4(cy, v, b); fma4(cz, v, c);
      }
    }
    stcg4(P.core_rc + p, rc);
    if (GGA) { stcg4(P.core_cx + p, cx); stcg4(P.core_cy + p, cy); stcg4(P.core_cz + p, cz); }
  }
}

__device__ __forceinline__ void st(double* p, size_t i, double v) {
  if (p) p[i] = v;
}

// One grid point of K1.  NP = nact rounded up to a multiple of 4 (coefficients beyond nact are
// zero-padded).  acc += (rho, rho_a, rho_b) * w.  Shared by the stand-alone kernel below and by the
// density warps of the fused K1+K2 kernel (k2_ontop.cuh).
template <int NP, bool GGA>
__device__ __forceinline__ void density_point(const Params& P, const size_t p, double (&acc)[3]) {
  const size_t ld = P.ld;
  // ---- core orbitals: closed form (from k1_core's partials when those were produced) ----
  double rc = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
  int i0 = 0;
  if (P.core_rc) {
    rc = P.core_rc[p];
    if (GGA) { cx = P.core_cx[p]; cy = P.core_cy[p]; cz = P.core_cz[p]; }
    i0 = P.core_done > 0 ? P.core_done : P.ncore;
  }
  {
    for (int i = i0; i < P.ncore; ++i) {
      const size_t o = (size_t)i * ld + p;
      const double v = __ldcs(P.phi + o);
      rc = fma(v, v, rc);
      if (GGA) {
        cx = fma(v, __ldcs(P.phix + o), cx);
        cy = fma(v, __ldcs(P.phiy + o), cy);
        cz = fma(v, __ldcs(P.phiz + o), cz);
      }
    }
  }
  // ---- active orbitals: T = Ds . phi_act in registers ----
  double f[NP];
#pragma unroll
  for (int t = 0; t < NP; ++t) f[t] = (t < P.nact) ? __ldcs(P.phi + (size_t)(P.ncore + t) * ld + p) : 0.0;
  double ra = 0.0, rb = 0.0, ax = 0.0, ay = 0.0, az = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
  // four rows of T at a time: 8 independent FMA chains (a single row pair leaves the FP64
  // pipe latency-bound at this kernel's occupancy)
#pragma unroll
  for (int t0 = 0; t0 < NP; t0 += 4) {
    double Ta[4] = {0.0, 0.0, 0.0, 0.0}, Tb[4] = {0.0, 0.0, 0.0, 0.0};
    double fx[4], fy[4], fz[4];
    if (GGA) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool live = (t0 + j) < P.nact;
        const size_t o = (size_t)(P.ncore + (live ? t0 + j : 0)) * ld + p;
        fx[j] = live ? __ldcs(P.phix + o) : 0.0;
        fy[j] = live ? __ldcs(P.phiy + o) : 0.0;
        fz[j] = live ? __ldcs(P.phiz + o) : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < NP; ++u) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Ta[j] = fma(c_Dsa[(t0 + j) * MAX_ACT + u], f[u], Ta[j]);
        Tb[j] = fma(c_Dsb[(t0 + j) * MAX_ACT + u], f[u], Tb[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ra = fma(Ta[j], f[t0 + j], ra);
      rb = fma(Tb[j], f[t0 + j], rb);
      if (GGA) {
        ax = fma(Ta[j], fx[j], ax); ay = fma(Ta[j], fy[j], ay); az = fma(Ta[j], fz[j], az);
        bx = fma(Tb[j], fx[j], bx); by = fma(Tb[j], fy[j], by); bz = fma(Tb[j], fz[j], bz);
      }
    }
  }
  const double rhoa = rc + ra, rhob = rc + rb;
  const double rho = rhoa + rhob;  // mcpdft.cc:89
  st(P.rho, p, rho);
  st(P.rhoa, p, rhoa);
  st(P.rhob, p, rhob);
  st(P.pi_core, p, rc * rc + rc * rb + ra * rc);
  if (GGA) {
    if (P.pcg[0]) {  // grad(rc^2 + rc (rta + rtb)) with grad rc = 2 c, grad rt = 2 (T . grad phi)
      const double rt = ra + rb;
      P.pcg[0][p] = 2.0 * rc * (2.0 * cx) + (2.0 * cx) * rt + rc * (2.0 * ax + 2.0 * bx);
      P.pcg[1][p] = 2.0 * rc * (2.0 * cy) + (2.0 * cy) * rt + rc * (2.0 * ay + 2.0 * by);
      P.pcg[2][p] = 2.0 * rc * (2.0 * cz) + (2.0 * cz) * rt + rc * (2.0 * az + 2.0 * bz);
    }
    ax = 2.0 * (cx + ax); ay = 2.0 * (cy + ay); az = 2.0 * (cz + az);
    bx = 2.0 * (cx + bx); by = 2.0 * (cy + by); bz = 2.0 * (cz + bz);
    st(P.gx, p, ax + bx);  // mcpdft.cc:381-383
    st(P.gy, p, ay + by);
    st(P.gz, p, az + bz);
    st(P.grad[0], p, ax); st(P.grad[1], p, ay); st(P.grad[2], p, az);
    st(P.grad[3], p, bx); st(P.grad[4], p, by); st(P.grad[5], p, bz);
    st(P.saa, p, (ax * ax) + (ay * ay) + (az * az));  // mcpdft.cc:162-164
    st(P.sbb, p, (bx * bx) + (by * by) + (bz * bz));
    st(P.sab, p, (ax * bx) + (ay * by) + (az * bz));
  }
  const double wp = __ldg(P.w + p);
  acc[0] = fma(rhoa + rhob, wp, acc[0]);  // mcpdft.cc:91-93
  acc[1] = fma(rhoa, wp, acc[1]);
  acc[2] = fma(rhob, wp, acc[2]);
}

// (min blocks per SM pinned to what the kernel had before the core loop moved in beside the partials: 128 registers up to
//  24 active orbitals, 168 above)
template <int NP, bool GGA>
__global__ void __launch_bounds__(THREADS, NP <= 24 ? 4 : 3) k1_density(const Params P) {
  __shared__ double scratch[3 * 32];
  double acc[3] = {0.0, 0.0, 0.0};
  for (size_t p = (size_t)blockIdx.x * THREADS + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * THREADS)
    density_point<NP, GGA>(P, p, acc);
  ordm::block_sum_store<3>(acc, scratch, P.partial + (size_t)blockIdx.x * 3);
}

// Polyradical analysis, MCPDFTSolver::polyradical_analysis (src/libpsi4/mcpdft_solver.cc:3001-3240): the "effectively
// unpaired electron" densities  U(r), D(r), N(r) = sum_ij phi_i(r) phi_j(r) {U, D, N}_ij  with
//   U_ij = 1 - |1 - n_ij|,  D_ij = n_ij (2 - n_ij),  N_ij = n_ij^2 (2 - n_ij)^2,  n = D1a + D1b   (:3104-3108)
// and their quadrature traces (:3185-3187).  Doubly occupied core orbitals (n = 2) contribute nothing, so only the
// active block is contracted: the same shape as the active part of K1, three matrices instead of two.  The matrices
// (host-built, row-major [3][nact][nact]) are read through the read-only path: every lane of a warp reads the same
// element.  Not on the energy path: a diagnostic the plugin prints and writes to mygrids.txt.
struct PolyParams {
  const double* phi_act;   // [nact][ld]
  const double* w;
  const double* udn;       // [3][nact][nact]
  size_t ld, npts;
  int nact;
  double *U, *D, *N;       // per-point outputs (nullable)
  double* partial;         // [gridDim.x][3]: sum w U, sum w D, sum w N
};

template <int NP>
__global__ void __launch_bounds__(THREADS) k1_polyradical(const PolyParams P) {
  __shared__ double scratch[3 * 32];
  double acc[3] = {0.0, 0.0, 0.0};
  const int na = P.nact;
  const double *Um = P.udn, *Dm = P.udn + (size_t)na * na, *Nm = P.udn + (size_t)2 * na * na;
  for (size_t p = (size_t)blockIdx.x * THREADS + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * THREADS) {
    double f[NP];
#pragma unroll
    for (int t = 0; t < NP; ++t) f[t] = (t < na) ? __ldcs(P.phi_act + (size_t)t * P.ld + p) : 0.0;
    double du = 0.0, dd = 0.0, dn = 0.0;
#pragma unroll
    for (int t = 0; t < NP; ++t) {
      if (t < na) {
        double ru = 0.0, rd = 0.0, rn = 0.0;   // row t of the three matrices against phi
#pragma unroll
        for (int u = 0; u < NP; ++u) {
          if (u < na) {
            ru = fma(__ldg(Um + t * na + u), f[u], ru);
            rd = fma(__ldg(Dm + t * na + u), f[u], rd);
            rn = fma(__ldg(Nm + t * na + u), f[u], rn);
          }
        }
        du = fma(ru, f[t], du); dd = fma(rd, f[t], dd); dn = fma(rn, f[t], dn);
      }
    }
    st(P.U, p, du); st(P.D, p, dd); st(P.N, p, dn);
    const double wp = __ldg(P.w + p);
    acc[0] = fma(du, wp, acc[0]); acc[1] = fma(dd, wp, acc[1]); acc[2] = fma(dn, wp, acc[2]);
  }
  ordm::block_sum_store<3>(acc, scratch, P.partial + (size_t)blockIdx.x * 3);
}

// Generic path for nact > MAX_ACT: same math, coefficients from global memory (L1/L2
// resident), active orbital values re-read through L1.  Correct for any size; not tuned.
template <bool GGA>
__global__ void __launch_bounds__(THREADS) k1_density_generic(const Params P) {
  __shared__ double scratch[3 * 32];
  double acc[3] = {0.0, 0.0, 0.0};
  const size_t ld = P.ld;
  const int na = P.nact;
  for (size_t p = (size_t)blockIdx.x * THREADS + threadIdx.x; p < P.npts; p += (size_t)gridDim.x * THREADS) {
    double rc = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
    int i0 = 0;
    if (P.core_rc) {
      rc = P.core_rc[p];
      if (GGA) { cx = P.core_cx[p]; cy = P.core_cy[p]; cz = P.core_cz[p]; }
      i0 = P.core_done > 0 ? P.core_done : P.ncore;
    }
    {
      for (int i = i0; i < P.ncore; ++i) {
        const size_t o = (size_t)i * ld + p;
        const double v = __ldcs(P.phi + o);
        rc = fma(v, v, rc);
        if (GGA) {
          cx = fma(v, __ldcs(P.phix + o), cx);
          cy = fma(v, __ldcs(P.phiy + o), cy);
          cz = fma(v, __ldcs(P.phiz + o), cz);
        }
      }
    }
    double ra = 0.0, rb = 0.0, ax = 0.0, ay = 0.0, az = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
    const double* fa = P.phi + (size_t)P.ncore * ld + p;
    for (int t = 0; t < na; ++t) {
      double Ta = 0.0, Tb = 0.0;
      for (int u = 0; u < na; ++u) {
        const double fu = __ldg(fa + (size_t)u * ld);
        Ta = fma(__ldg(P.Dsa_g + t * na + u), fu, Ta);
        Tb = fma(__ldg(P.Dsb_g + t * na + u), fu, Tb);
      }
      const double ft = __ldg(fa + (size_t)t * ld);
      ra = fma(Ta, ft, ra);
      rb = fma(Tb, ft, rb);
      if (GGA) {
        const size_t o = (size_t)(P.ncore + t) * ld + p;
        const double fx = __ldcs(P.phix + o), fy = __ldcs(P.phiy + o), fz = __ldcs(P.phiz + o);
        ax = fma(Ta, fx, ax); ay = fma(Ta, fy, ay); az = fma(Ta, fz, az);
        bx = fma(Tb, fx, bx); by = fma(Tb, fy, by); bz = fma(Tb, fz, bz);
      }
    }
    const double rhoa = rc + ra, rhob = rc + rb, rho = rhoa + rhob;
    st(P.rho, p, rho);
    st(P.rhoa, p, rhoa);
    st(P.rhob, p, rhob);
    st(P.pi_core, p, rc * rc + rc * rb + ra * rc);
    if (GGA) {
      if (P.pcg[0]) {
        const double rt = ra + rb;
        P.pcg[0][p] = 2.0 * rc * (2.0 * cx) + (2.0 * cx) * rt + rc * (2.0 * ax + 2.0 * bx);
        P.pcg[1][p] = 2.0 * rc * (2.0 * cy) + (2.0 * cy) * rt + rc * (2.0 * ay + 2.0 * by);
        P.pcg[2][p] = 2.0 * rc * (2.0 * cz) + (2.0 * cz) * rt + rc * (2.0 * az + 2.0 * bz);
      }
      ax = 2.0 * (cx + ax); ay = 2.0 * (cy + ay); az = 2.0 * (cz + az);
      bx = 2.0 * (cx + bx); by = 2.0 * (cy + by); bz = 2.0 * (cz + bz);
      st(P.gx, p, ax + bx); st(P.gy, p, ay + by); st(P.gz, p, az + bz);
      st(P.grad[0], p, ax); st(P.grad[1], p, ay); st(P.grad[2], p, az);
      st(P.grad[3], p, bx); st(P.grad[4], p, by); st(P.grad[5], p, bz);
      st(P.saa, p, (ax * ax) + (ay * ay) + (az * az));
      st(P.sbb, p, (bx * bx) + (by * by) + (bz * bz));
      st(P.sab, p, (ax * bx) + (ay * by) + (az * bz));
    }
    const double wp = __ldg(P.w + p);
    acc[0] = fma(rhoa + rhob, wp, acc[0]);
    acc[1] = fma(rhoa, wp, acc[1]);
    acc[2] = fma(rhob, wp, acc[2]);
  }
  ordm::block_sum_store<3>(acc, scratch, P.partial + (size_t)blockIdx.x * 3);
}

}  // namespace k1
