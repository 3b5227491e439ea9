// k1_tile.cuh -- K1t: the active-space part of K1 on TMA-staged point tiles (energy path, GGA layouts).
// Same quantities as k1::k1_density on the core partials (MCPDFT::build_density_functions, mcpdft.cc:42-105, and
// build_density_gradients, mcpdft.cc:107-176, restricted to what mcpdft_energy consumes: rho, grad rho, the closed-form
// core part of Pi and the three integrated densities).  Re-planned around what the ncu capture of k1_density showed
// (profiles/r02B_ncu_k1_active_k3.txt: every stall sample a long-scoreboard wait at the first use of a batch of
// thread-private loads, 16 warps per SM, FP64 pipe 35 % busy, 2.5 issued instructions per DFMA):
//   * one persistent 9-warp CTA per SM; warp 8's elected lane is the producer.  A 128-point tile arrives as one chunk
//     per orbital array: ONE cp.async.bulk.tensor.2d each (box = 128 points x all staged columns; out-of-range points
//     are zero-filled by the TMA unit), the phi chunk also carrying five 1 KB bulk copies (core partial sums, weights).
//     Chunks go round a ring of 6-9 shared-memory slots with full / empty mbarriers, so most of the ring is in flight
//     while a tile is being consumed -- the memory stream no longer depends on what compute warps hold in registers;
//   * the energy needs the SUM density only:  T = (Dsa + Dsb) phi,  rho_t = T.phi,  grad rho_t = 2 T.grad phi  --
//     400 + 80 DFMAs instead of 2 x (400 + 80) -- plus the integrated spin densities, for which
//     phi^T (Dsa - Dsb) phi is enough (400 + 20 DFMAs, no gradient columns);
//   * the two jobs go to two warps per 32 points (sum role, difference role), so a point's FP64 work is spread over two
//     sub-partition schedulers and every thread has 20 independent DFMA chains;
//   * coefficients sit in shared memory ([u][t], copied once per CTA): a row is NP / 2 broadcast LDS.128 into ordinary
//     registers and the loop over u stays rolled (0.5 KB of code) -- see the note at dmat_doubles() for the
//     constant-bank form that was tried first;
//   * everything that needs the finished T (rho_t and the three gradient dots) is one straight-line block after all four
//     chunks of the tile have been waited for: eight independent chains whose shared-memory loads can move to the top.
// Measured (tools/k1lab, 2 M points, 20 active + 5 staged core columns, profiles/r02_k1_tile_lab.log): 0.317 ms = 5.55 TB/s
// = 85 % of the measured HBM copy bandwidth, against 0.342 ms for k1_density; config 4's 5 M points: 0.78 vs 0.85 ms.
// Columns staged per array: [c0, n_occ) = the core columns the side kernel left behind (k1::Params::core_done) followed
// by the active block.  Diagnostic vectors (rho_a, rho_b, spin gradients, sigma) and the core part of grad Pi are not
// produced here: those requests keep k1_density, as do grids without orbital gradients (one 20-column chunk per tile
// streams no faster than the thread-per-point kernel).
// Tried and dropped, kept as a lab variant (tools/k1lab/k1_tile_fused.cuh): K3 fused behind the density (R, translation and
// functional in extra warps fed through shared memory).  A point's XC is a ~10 k-cycle dependent FP64 chain that wants the
// 32 warps per SM k3_translate_xc runs with; with 8 XC warps per SM the fused kernel took 1.26 ms on config 4 against
// 0.78 + 0.29 ms for the two launches.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "k1_density.cuh"
#include "reduce.cuh"

namespace k1t {

#ifndef K1T_NAP_NS
#define K1T_NAP_NS 200
#endif
constexpr int TILE = 128;             // points per chunk row
constexpr int COMPUTE_WARPS = 8;      // 4 point quarters x 2 roles
constexpr int THREADS = 32 * (COMPUTE_WARPS + 1);
constexpr int MAX_ACT = k1::MAX_ACT;
#ifndef K1T_MAX_LEAD
#define K1T_MAX_LEAD 12
#endif
constexpr int MAX_LEAD = K1T_MAX_LEAD;   // staged core columns ahead of the active block

// Coefficients: Params::dmat = [2][MAX_ACT * MAX_ACT] doubles in global memory, role 0: Dsa + Dsb, role 1: Dsa - Dsb (both
// symmetric, element (t, u) at [u * MAX_ACT + t], zero beyond nact); every CTA copies the NP x NP corners into shared memory
// once.  (Constant-bank operands were tried first: 400 LDCU.128 per point and role overran the uniform register file --
// UMOV / MOV.SPILL traffic, a short-scoreboard stall on most DFMAs -- and the unrolled 400-DFMA body missed the instruction
// cache: 0.48 ms per 2 M points.  From shared memory the row of a coefficient matrix is ten broadcast LDS.128 and the
// loop over u stays rolled.)
__host__ __device__ constexpr int dmat_doubles(int np) { return 2 * np * np; }

struct alignas(64) Maps { CUtensorMap m[4]; };   // phi, phi_x, phi_y, phi_z: dims {npts, n_occ - c0}, box {TILE, ncols}

struct Params {
  const double* w;
  const double *core_rc, *core_cx, *core_cy, *core_cz;   // partial sums over core orbitals [0, c0) (null: there are none)
  size_t npts;
  int ntiles;
  int nlead;       // staged core columns ahead of the active block (= ncore - c0) <= MAX_LEAD
  int nact;
  int have_diff;   // Dsa != Dsb
  int nslots;      // ring slots of slot_doubles() each
  double *rho, *gx, *gy, *gz, *pi_core;   // outputs, npts long (pi_core nullable)
  double* partial;                         // [gridDim.x][3]: sum rho w, sum rho_a w, sum rho_b w
  const double* dmat;                      // [2][MAX_ACT * MAX_ACT], see above
};

constexpr int NVEC = 5;   // staged vectors behind the phi columns: rc cx cy cz w
__host__ __device__ constexpr size_t slot_doubles(int ncols) { return (size_t)TILE * (ncols + NVEC); }
__host__ __device__ constexpr size_t smem_bytes(int ncols, int nslots, int np) {
  return (nslots * slot_doubles(ncols) + dmat_doubles(np)) * sizeof(double) + 16 * nslots;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// (waiting warps nap between polls instead of spinning on the barrier word)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  while (!mbar_try(bar, parity)) __nanosleep(K1T_NAP_NS);
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

// position of a warp (or of the producer) in the ring: every party visits every chunk, in the same order -- a parity wait
// on an mbarrier cannot tell the phase it needs from one two phases older, so nobody may skip a use of a slot
struct RingPos {
  int slot, phase;
  __device__ __forceinline__ void next(int nslots) { if (++slot == nslots) { slot = 0; phase ^= 1; } }
};

// What every role needs to find its way round the ring.  (The ring itself is reached from the extern shared array inside
// every role: a pointer handed through this struct would be a generic one, and the orbital reads LD.E instead of LDS.)
struct Ctx {
  size_t sdoubles;     // doubles per ring slot; the coefficient block follows the nslots slots
  uint32_t bar0;       // full[nslots], empty[nslots]
  int nslots, ncols, nlocal;
};

extern __shared__ __align__(1024) unsigned char smem_raw[];
__device__ __forceinline__ double* ring_base() { return reinterpret_cast<double*>(smem_raw); }

// T = D . phi_act(p):  T[t] = sum_u D[u][t] phi_u  with D in shared memory as [u][NP] (symmetric), phi_u = A[u * TILE].
// The loop over u stays rolled (one LDS.64 + NP / 2 broadcast LDS.128 + NP DFMAs on NP independent chains per round;
// fully unrolled it took 0.43 instead of 0.35 ms per 2 M points -- the instruction cache).
template <int NP>
__device__ __forceinline__ void matvec(const double* __restrict__ dm, const double* __restrict__ A, const int na, double (&T)[NP]) {
#pragma unroll
  for (int t = 0; t < NP; ++t) T[t] = 0.0;
#pragma unroll 2
  for (int u = 0; u < na; ++u) {
    const double fu = A[u * TILE];
    const double2* row = reinterpret_cast<const double2*>(dm + u * NP);
#pragma unroll
    for (int t2 = 0; t2 < NP / 2; ++t2) {
      const double2 d = row[t2];
      T[2 * t2] = fma(d.x, fu, T[2 * t2]);
      T[2 * t2 + 1] = fma(d.y, fu, T[2 * t2 + 1]);
    }
  }
}

// The roles are separate, non-inlined functions on purpose: each gets its own register allocation (inlined into one kernel
// body, ptxas hoisted one role's loop invariants at the other's expense).  They copy Params / Ctx into registers first: the
// mbarrier asm's memory clobbers would otherwise force a re-load of every field at every use.

// ---------------------------------------------------------------- producer (one lane)
__device__ __noinline__ void producer_role(const Maps& maps, const Params& P_, const Ctx& c_) {
  const Params P = P_;
  const Ctx c = c_;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const uint32_t abytes = (uint32_t)(c.ncols * TILE * 8);
  const uint32_t bytes0 = abytes + (uint32_t)((P.core_rc ? NVEC : 1) * TILE * 8);
  RingPos rp{0, 0};
  bool wrapped = false;
  for (int it = 0; it < c.nlocal; ++it) {
    const int x = ((int)blockIdx.x + it * (int)gridDim.x) * TILE;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wrapped) mbar_wait(c.bar0 + 8 * (c.nslots + rp.slot), (uint32_t)(rp.phase ^ 1));
      const uint32_t full = c.bar0 + 8 * rp.slot;
      mbar_expect_tx(full, a == 0 ? bytes0 : abytes);
      const uint32_t dst = smem_u32(ring_base() + (size_t)rp.slot * c.sdoubles);
      tma_2d(dst, &maps.m[a], x, 0, full, pol);
      if (a == 0) {
        const uint32_t vdst = dst + abytes;
        tma_1d(vdst + (uint32_t)((NVEC - 1) * TILE * 8), P.w + x, TILE * 8, full, pol);
        if (P.core_rc) {
          tma_1d(vdst, P.core_rc + x, TILE * 8, full, pol);
          tma_1d(vdst + 1 * TILE * 8, P.core_cx + x, TILE * 8, full, pol);
          tma_1d(vdst + 2 * TILE * 8, P.core_cy + x, TILE * 8, full, pol);
          tma_1d(vdst + 3 * TILE * 8, P.core_cz + x, TILE * 8, full, pol);
        }
      }
      if (rp.slot == c.nslots - 1) wrapped = true;
      rp.next(c.nslots);
    }
  }
}

// ---------------------------------------------------------------- sum role: rho, grad rho, core part of Pi
// NP = nact rounded up to a multiple of 4 (coefficient rows / columns beyond nact are zero)
template <int NP>
__device__ __noinline__ void sum_role(const Params& P_, const Ctx& c_, const int quarter, double (&acc_out)[3]) {
  const Params P = P_;
  const Ctx c = c_;
  double acc[3] = {0.0, 0.0, 0.0};
  const int lane = threadIdx.x & 31, pt = quarter * 32 + lane;
  const int na = P.nact, nl = P.nlead, nslots = c.nslots, ncols = c.ncols;
  const bool have_core = P.core_rc != nullptr;
  const double* dm = ring_base() + (size_t)nslots * c.sdoubles;   // Dsa + Dsb
  RingPos rp{0, 0};
  for (int it = 0; it < c.nlocal; ++it) {
    const size_t p = (size_t)((int)blockIdx.x + it * (int)gridDim.x) * TILE + pt;
    const bool inside = p < P.npts;
    // ---- chunk 0: phi columns + vectors
    mbar_wait(c.bar0 + 8 * rp.slot, (uint32_t)rp.phase);
    const double* F = ring_base() + (size_t)rp.slot * c.sdoubles + pt;   // phi column k of this point: F[k * TILE]
    const double* V = F + (size_t)ncols * TILE;                          // rc cx cy cz w
    const double wp = inside ? V[(NVEC - 1) * TILE] : 0.0;
    const double* A = F + (size_t)nl * TILE;
    double rc = 0.0, cc[3] = {0.0, 0.0, 0.0};
    if (have_core && inside) { rc = V[0]; cc[0] = V[TILE]; cc[1] = V[2 * TILE]; cc[2] = V[3 * TILE]; }
    double T[NP];
    matvec<NP>(dm, A, na, T);
    // ---- chunks 1..3: grad phi.  One straight-line block for everything that needs the finished T -- rho_t = T.phi, the
    // three gradient dots and the staged core columns -- eight independent chains with all their shared-memory loads free
    // to move to the top (three separate wait -> load -> chain phases cost as much as the 400-DFMA product itself)
    RingPos rq[4];
    rq[0] = rp;
#pragma unroll
    for (int a = 1; a < 4; ++a) { rq[a] = rq[a - 1]; rq[a].next(nslots); }
    const double* G[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mbar_wait(c.bar0 + 8 * rq[a + 1].slot, (uint32_t)rq[a + 1].phase);
      G[a] = ring_base() + (size_t)rq[a + 1].slot * c.sdoubles + pt;
    }
    if (nl > 0) {   // core columns the side kernel left behind (the phi chunk is still held)
#pragma unroll
      for (int k = 0; k < MAX_LEAD; ++k) {
        if (k < nl) {
          const double v = F[k * TILE];
          rc = fma(v, v, rc);
#pragma unroll
          for (int a = 0; a < 3; ++a) cc[a] = fma(v, G[a][k * TILE], cc[a]);
        }
      }
    }
    double s0 = 0.0, s1 = 0.0, ga[3] = {0.0, 0.0, 0.0}, gb[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int t = 0; t < NP; t += 2) {
      if (t < na) {
        s0 = fma(T[t], A[t * TILE], s0);
#pragma unroll
        for (int a = 0; a < 3; ++a) ga[a] = fma(T[t], G[a][(nl + t) * TILE], ga[a]);
      }
      if (t + 1 < na) {
        s1 = fma(T[t + 1], A[(t + 1) * TILE], s1);
#pragma unroll
        for (int a = 0; a < 3; ++a) gb[a] = fma(T[t + 1], G[a][(nl + t + 1) * TILE], gb[a]);
      }
    }
    double g[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) g[a] = 2.0 * ((cc[a] + cc[a]) + (ga[a] + gb[a]));   // grad rho = grad rho_a + grad rho_b (mcpdft.cc:381-383), grad rho_s = 2 (c + T_s . grad phi)
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a) mbar_arrive(c.bar0 + 8 * (nslots + rq[a].slot));
    }
    rp = rq[3];
    rp.next(nslots);
    const double st = s0 + s1;                  // rho_t,a + rho_t,b
    const double rho = (rc + rc) + st;          // mcpdft.cc:89 (rho_a + rho_b)
    if (inside) {
      P.rho[p] = rho;
      P.gx[p] = g[0]; P.gy[p] = g[1]; P.gz[p] = g[2];
      if (P.pi_core) P.pi_core[p] = rc * (rc + st);   // rc^2 + rc rho_t,b + rho_t,a rc
    }
    const double half = fma(0.5, st, rc);       // (rho_a + rho_b) / 2
    acc[0] = fma(rho, wp, acc[0]);              // mcpdft.cc:91-93
    acc[1] = fma(half, wp, acc[1]);
    acc[2] = fma(half, wp, acc[2]);
  }
  acc_out[0] = acc[0]; acc_out[1] = acc[1]; acc_out[2] = acc[2];
}

// ---------------------------------------------------------------- difference role: (rho_a - rho_b) w for the integrated spin densities
template <int NP>
__device__ __noinline__ void diff_role(const Params& P_, const Ctx& c_, const int quarter, double (&acc_out)[3]) {
  const Params P = P_;
  const Ctx c = c_;
  double acc1 = 0.0;
  const int lane = threadIdx.x & 31, pt = quarter * 32 + lane;
  const int na = P.nact, nl = P.nlead, nslots = c.nslots, ncols = c.ncols;
  const double* dm = ring_base() + (size_t)nslots * c.sdoubles + NP * NP;   // Dsa - Dsb
  RingPos rp{0, 0};
  for (int it = 0; it < c.nlocal; ++it) {
    const size_t p = (size_t)((int)blockIdx.x + it * (int)gridDim.x) * TILE + pt;
    mbar_wait(c.bar0 + 8 * rp.slot, (uint32_t)rp.phase);
    if (P.have_diff) {
      const double* F = ring_base() + (size_t)rp.slot * c.sdoubles + pt;
      const double wp = p < P.npts ? F[(size_t)(ncols + NVEC - 1) * TILE] : 0.0;
      const double* A = F + (size_t)nl * TILE;
      double T[NP];
      matvec<NP>(dm, A, na, T);
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int t = 0; t < NP; t += 2) {
        if (t < na) d0 = fma(T[t], A[t * TILE], d0);
        if (t + 1 < na) d1 = fma(T[t + 1], A[(t + 1) * TILE], d1);
      }
      acc1 = fma(0.5 * (d0 + d1), wp, acc1);   // (rho_a - rho_b) w / 2
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {   // this role reads nothing of the gradient chunks; passing their full barriers keeps its arrivals in phase
      if (a > 0) mbar_wait(c.bar0 + 8 * rp.slot, (uint32_t)rp.phase);
      __syncwarp();
      if (lane == 0) mbar_arrive(c.bar0 + 8 * (nslots + rp.slot));
      rp.next(nslots);
    }
  }
  acc_out[1] = acc1; acc_out[2] = -acc1;
}

template <int NP>
__global__ void __launch_bounds__(THREADS, 1) k1_tile(const __grid_constant__ Maps maps, const __grid_constant__ Params P) {
  __shared__ double scratch[3 * 32];
  Ctx c;
  c.ncols = P.nlead + P.nact; c.nslots = P.nslots;
  c.sdoubles = slot_doubles(c.ncols);
  double* const dm = ring_base() + (size_t)c.nslots * c.sdoubles;
  c.bar0 = smem_u32(dm + dmat_doubles(NP));
  c.nlocal = P.ntiles > (int)blockIdx.x ? (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < dmat_doubles(NP); i += THREADS) {   // [role][u][t] <- dmat[role][u * MAX_ACT + t]
    const int role = i / (NP * NP), u = (i % (NP * NP)) / NP, t = i % NP;
    dm[i] = P.dmat[(size_t)role * MAX_ACT * MAX_ACT + u * MAX_ACT + t];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.nslots; ++s) { mbar_init(c.bar0 + 8 * s, 1); mbar_init(c.bar0 + 8 * (c.nslots + s), COMPUTE_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double acc[3] = {0.0, 0.0, 0.0};
  if (warp == COMPUTE_WARPS) {
    if (lane == 0) producer_role(maps, P, c);
  } else if (warp < 4) {
    sum_role<NP>(P, c, warp & 3, acc);
  } else {
    diff_role<NP>(P, c, warp & 3, acc);
  }
  __syncthreads();
  ordm::block_sum_store<3>(acc, scratch, P.partial + (size_t)blockIdx.x * 3);
}

}  // namespace k1t
