// k1_tile.cuh -- K1t: the active-space part of K1 on TMA-staged point tiles, with K3 fused behind it (energy path).
// Same quantities as k1::k1_density on the core partials (MCPDFT::build_density_functions, mcpdft.cc:42-105, and
// build_density_gradients, mcpdft.cc:107-176, restricted to what mcpdft_energy consumes: rho, grad rho, the closed-form
// core part of Pi and the three integrated densities) and, when asked, as k3::k3_translate_xc right behind them
// (build_R, translate_density(+_gradients), the functional: mcpdft.cc:272-416, functional.cc:12-309).  Re-planned around
// what the ncu capture of k1_density showed (profiles/r02B_ncu_k1_active_k3.csv: every stall sample a long-scoreboard
// wait at the first use of a batch of thread-private loads, 16 warps per SM, FP64 pipe 35 % busy, 2.5 issued
// instructions per DFMA):
//   * one persistent 13-warp CTA per SM; warp 12's elected lane is the producer.  A 128-point tile arrives as one chunk
//     per orbital array: ONE cp.async.bulk.tensor.2d each (box = 128 points x all staged columns; out-of-range points
//     are zero-filled by the TMA unit), the phi chunk also carrying five 1 KB bulk copies (core partial sums, weights).
//     Chunks go round a ring of 6-9 shared-memory slots with full / empty mbarriers, so ~85 % of the ring is in flight
//     while one chunk is being consumed -- the memory stream no longer depends on what compute warps hold in registers;
//   * the energy needs the SUM density only:  T = (Dsa + Dsb) phi,  rho_t = T.phi,  grad rho_t = 2 T.grad phi  --
//     400 + 80 DFMAs instead of 2 x (400 + 80) -- plus the integrated spin densities, for which
//     phi^T (Dsa - Dsb) phi is enough (400 + 20 DFMAs, no gradient columns);
//   * the two jobs go to different warps per 32 points (sum role: 4 warps; difference role: 8 warps, each owning every
//     other tile of its point quarter), so a point's FP64 work is spread over several schedulers and every thread has 20
//     independent DFMA chains;
//   * coefficients sit in shared memory ([u][t], copied once per CTA): a row is NP / 2 broadcast LDS.128, the loop over u
//     stays rolled (0.5 KB of code) -- see the note at dmat_doubles() for the constant-bank form that was tried first;
//   * fused XC (Params::xc >= 0): the sum role hands rho, grad rho and the core part of Pi of its 32 points to the
//     difference-role warp that owns the tile through a double-buffered shared-memory block (mbarriers per quarter and
//     parity); that warp, whose density job is the lighter one, runs R, the translation and the functional on them, reads
//     Pi_act (K2's output), writes the completed Pi back and accumulates the five XC sums.  The kernel streams at HBM
//     speed with the FP64 pipe about half busy, so K3's ~500 FP64 instructions per point ride along: no K3 launch, no
//     pi_core round trip, rho / grad rho are written for the getters but not re-read.
// Columns staged per array: [c0, n_occ) = the core columns the side kernel left behind (k1::Params::core_done) followed
// by the active block.  Diagnostic vectors (rho_a, rho_b, spin gradients, sigma, R, translated densities) and the
// core part of grad Pi are not produced here: those requests keep k1_density / k3_translate_xc.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "k1_density.cuh"
#include "reduce.cuh"
#include "xc_math.cuh"

namespace k1t {

#ifndef K1T_NAP_NS
#define K1T_NAP_NS 200
#endif
constexpr int TILE = 128;             // points per chunk row
constexpr int COMPUTE_WARPS = 12;     // 4 point quarters x (sum role + two difference / XC warps)
constexpr int EMPTY_ARRIVALS = 8;     // per ring slot: phi chunk 4 (sum) + 4 (owning difference warps); gradient chunks 4 x 2 (sum)
constexpr int THREADS = 32 * (COMPUTE_WARPS + 1);
constexpr int MAX_ACT = k1::MAX_ACT;
constexpr int MAX_LEAD = 8;           // staged core columns ahead of the active block
constexpr int EXCH_DOUBLES = 2 * 5 * TILE;   // role 0 -> role 1 hand-over: 2 buffers x (rho gx gy gz pi_core) x 128 points

// Coefficients: Params::dmat = [2][MAX_ACT * MAX_ACT] doubles in global memory, role 0: Dsa + Dsb, role 1: Dsa - Dsb (both
// symmetric, element (t, u) at [u * MAX_ACT + t], zero beyond nact); every CTA copies the NP x NP corners into shared memory
// once.  (Constant-bank operands were tried first: 400 LDCU.128 per point and role overran the uniform register file --
// UMOV / MOV.SPILL traffic, a short-scoreboard stall on most DFMAs -- and the unrolled 400-DFMA body missed the instruction
// cache, profiles/r02K_ncu_k1_tile_ldcu.txt.  From shared memory the row of a coefficient matrix is ten broadcast LDS.128
// into ordinary registers and the loop over u stays rolled: 0.5 KB of code per role.)
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
  double *rho, *gx, *gy, *gz, *pi_core;   // outputs, npts long (pi_core nullable; not written in fused mode)
  double* partial;                         // [gridDim.x][3]: sum rho w, sum rho_a w, sum rho_b w
  const double* dmat;                      // [2][MAX_ACT * MAX_ACT], see above
  // fused XC
  int xc;           // -1: none; else the functional id of k3::k3_translate_xc (0 tSVWN, 1 tPBE, 2 trevPBE, 3 tBLYP, 4 tBOP, 5 twPBE)
  double omega;
  double* pi;       // in: Pi_act (K2), out: Pi_act + Pi_core
  double* partial3; // [gridDim.x][5]: trN, trNa, trNb, Ex, Ec
};

__host__ __device__ constexpr size_t slot_doubles(int ncols, bool gga) { return (size_t)TILE * (ncols + (gga ? 5 : 2)); }
__host__ __device__ constexpr size_t smem_bytes(int ncols, bool gga, int nslots, int np) {
  return (nslots * slot_doubles(ncols, gga) + EXCH_DOUBLES + dmat_doubles(np)) * sizeof(double) + 16 * (nslots + 16);
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
// Waiting warps nap between polls: a role that runs ahead (the difference role, the producer) would otherwise spin on the
// barrier word and take shared-memory / issue bandwidth from the warps it is waiting for (6 M polls per launch, tile period
// 2x longer: profiles/r02G_ncu_k1_tile_spin.txt)
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

// position of a warp (or of the producer) in the ring: every party visits the chunks in the same order
struct RingPos {
  int slot, phase;
  __device__ __forceinline__ void next(int nslots) { if (++slot == nslots) { slot = 0; phase ^= 1; } }
};

// R, translation and functional of one point (k3_translate_xc's body); acc += (trN, trNa, trNb, Ex, Ec) w
template <bool GGA>
__device__ __forceinline__ void xc_point(int func, double omega, double rho, double pi, double gx, double gy, double gz, double w, double (&acc)[5]) {
  const double R = xc::r_factor(rho, pi);
  const xc::Translated t = xc::translate_pt<GGA>(rho, pi, gx, gy, gz, R);
  double ex, ec;
  switch (func) {
    case 1: xc::pbe_pair_pt(t.ra, t.rb, t.saa, t.sab, t.sbb, xc::PBE_KAPPA, ex, ec); break;
    case 2: xc::pbe_pair_pt(t.ra, t.rb, t.saa, t.sab, t.sbb, xc::REVPBE_KAPPA, ex, ec); break;
    case 3: ex = xc::ex_b88_pt(t.ra, t.rb, t.saa, t.sbb); ec = xc::ec_lyp_pt(t.ra, t.rb, t.saa, t.sab, t.sbb); break;
    case 4: ex = xc::ex_b88_pt(t.ra, t.rb, t.saa, t.sbb); ec = xc::ec_b88_op_pt(t.ra, t.rb, t.saa, t.sbb); break;
    case 5: ex = xc::ex_wpbe_pt(t.ra, t.rb, t.saa, t.sbb, omega); ec = xc::ec_pbe_pt(t.ra, t.rb, t.saa, t.sab, t.sbb); break;
    default: ex = xc::ex_lsda_pt(t.ra, t.rb); ec = xc::ec_vwn3_pt(t.ra, t.rb); break;
  }
  acc[0] = fma(t.ra + t.rb, w, acc[0]);  // mcpdft.cc:332-334
  acc[1] = fma(t.ra, w, acc[1]);
  acc[2] = fma(t.rb, w, acc[2]);
  acc[3] = fma(ex, w, acc[3]);
  acc[4] = fma(ec, w, acc[4]);
}

// What every role needs to find its way round the ring
// (the ring itself is reached from the extern shared array inside every role: a pointer handed through this struct
//  would be a generic one, and the orbital reads LD.E instead of LDS)
struct Ctx {
  size_t sdoubles;     // doubles per ring slot; the hand-over block follows the nslots slots
  uint32_t bar0;       // full[nslots], empty[nslots]
  uint32_t xbar0;      // ready[2][4], taken[2][4], phi_here[2][4], phi_seen[2][4] (tile parity, point quarter)
  int nslots, ncols, nlocal;
};

// The roles are separate, non-inlined functions on purpose: inlined into one kernel body, ptxas hoisted the difference
// role's loop-invariant coefficients into the uniform register file and left the sum role ONE uniform register for its
// 400 coefficients -- every DFMA then waited for its own LDCU (short-scoreboard stall on each, 0.57 instead of 0.31 ms per
// 2 M points, profiles/r02E_ncu_k1_tile_v2.txt).  Per function, each role gets its own register allocation.

// ---------------------------------------------------------------- producer (one lane)
extern __shared__ __align__(1024) unsigned char smem_raw[];
__device__ __forceinline__ double* ring_base() { return reinterpret_cast<double*>(smem_raw); }

// T = D . phi_act(p):  T[t] = sum_u D[u][t] phi_u  with D in shared memory as [u][NP] (symmetric), phi_u = A[u * TILE].
// The loop over u stays rolled (one LDS.64 + NP / 2 broadcast LDS.128 + NP DFMAs on NP independent chains per round).
template <int NP>
__device__ __forceinline__ void matvec(const double* __restrict__ dm, const double* __restrict__ A, const int na, double (&T)[NP]) {
  constexpr int H = NP / 4;   // double2 per half row
#pragma unroll
  for (int t = 0; t < NP; ++t) T[t] = 0.0;
  // software pipeline at half-row granularity (a full second row does not fit the 128-register budget of a 13-warp CTA, and
  // ptxas then sinks every LDS.128 to just above its two DFMAs): while one half row is multiplied the other is in flight.
  // The loads of round na run one row past the matrix: still inside the CTA's shared memory, never used.
  double2 ha[H], hb[H];
  const double2* row = reinterpret_cast<const double2*>(dm);
#pragma unroll
  for (int k = 0; k < H; ++k) ha[k] = row[k];
  double fu = A[0];
#pragma unroll 1
  for (int u = 0; u < na; ++u) {
#pragma unroll
    for (int k = 0; k < H; ++k) hb[k] = row[H + k];
    const double fn = A[(u + 1) * TILE];
#pragma unroll
    for (int k = 0; k < H; ++k) { T[2 * k] = fma(ha[k].x, fu, T[2 * k]); T[2 * k + 1] = fma(ha[k].y, fu, T[2 * k + 1]); }
    row += NP / 2;
#pragma unroll
    for (int k = 0; k < H; ++k) ha[k] = row[k];
#pragma unroll
    for (int k = 0; k < H; ++k) { T[2 * (H + k)] = fma(hb[k].x, fu, T[2 * (H + k)]); T[2 * (H + k) + 1] = fma(hb[k].y, fu, T[2 * (H + k) + 1]); }
    fu = fn;
  }
}

template <bool GGA>
__device__ __noinline__ void producer_role(const Maps& maps, const Params& P_, const Ctx& c_) {
  const Params P = P_;   // register copies: the mbarrier asm's memory clobbers would otherwise force a re-load of every field per use
  const Ctx c = c_;
  constexpr int NARR = GGA ? 4 : 1, NVEC = GGA ? 5 : 2;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const uint32_t abytes = (uint32_t)(c.ncols * TILE * 8);
  const uint32_t bytes0 = abytes + (uint32_t)((P.core_rc ? NVEC : 1) * TILE * 8);
  RingPos rp{0, 0};
  bool wrapped = false;
  for (int it = 0; it < c.nlocal; ++it) {
    const int x = ((int)blockIdx.x + it * (int)gridDim.x) * TILE;
#pragma unroll
    for (int a = 0; a < NARR; ++a) {
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
          if (GGA) {
            tma_1d(vdst + 1 * TILE * 8, P.core_cx + x, TILE * 8, full, pol);
            tma_1d(vdst + 2 * TILE * 8, P.core_cy + x, TILE * 8, full, pol);
            tma_1d(vdst + 3 * TILE * 8, P.core_cz + x, TILE * 8, full, pol);
          }
        }
      }
      if (rp.slot == c.nslots - 1) wrapped = true;
      rp.next(c.nslots);
    }
  }
}

// ---------------------------------------------------------------- role 0: sum density, gradient, core part of Pi
// NP = nact rounded up to a multiple of 4 (coefficient rows / columns beyond nact are zero)
template <int NP, bool GGA>
__device__ __noinline__ void sum_role(const Params& P_, const Ctx& c_, const int quarter, double (&acc_out)[3]) {
  const Params P = P_;   // (register copies, see producer_role)
  const Ctx c = c_;
  constexpr int NVEC = GGA ? 5 : 2;
  double acc[3] = {0.0, 0.0, 0.0};
  const int lane = threadIdx.x & 31, pt = quarter * 32 + lane;
  const int na = P.nact, nl = P.nlead, nslots = c.nslots, ncols = c.ncols;
  const bool have_core = P.core_rc != nullptr, fused = P.xc >= 0;
  const double* dm = ring_base() + (size_t)nslots * c.sdoubles + EXCH_DOUBLES;   // Dsa + Dsb
  RingPos rp{0, 0};
  for (int it = 0; it < c.nlocal; ++it) {
    const size_t p = (size_t)((int)blockIdx.x + it * (int)gridDim.x) * TILE + pt;
    const bool inside = p < P.npts;
    // ---- chunk 0: phi columns + vectors
    mbar_wait(c.bar0 + 8 * rp.slot, (uint32_t)rp.phase);
    // tell the difference-role warp that owns this tile: it visits only every other phi chunk, and a parity wait on a full
    // barrier whose phases it has skipped cannot tell an old phase from the one it needs.  For the same reason this role
    // may not signal a tile before the partner has seen the previous signal (phi_seen): a waiter two phases behind blocks
    // on a parity that has already come round again (deadlock seen with 1 chunk per tile and a 7-tile ring).
    if (it >= 2) mbar_wait(c.xbar0 + 8 * (24 + 4 * (it & 1) + quarter), (uint32_t)(((it >> 1) - 1) & 1));
    mbar_arrive(c.xbar0 + 8 * (16 + 4 * (it & 1) + quarter));
    const double* F = ring_base() + (size_t)rp.slot * c.sdoubles + pt;   // phi column k of this point: F[k * TILE]
    const double* V = F + (size_t)ncols * TILE;                      // rc (cx cy cz) w
    const double wp = inside ? V[(NVEC - 1) * TILE] : 0.0;
    const double* A = F + (size_t)nl * TILE;
    double rc = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
    if (have_core && inside) {
      rc = V[0];
      if (GGA) { cx = V[TILE]; cy = V[2 * TILE]; cz = V[3 * TILE]; }
    }
    if (nl > 0) {   // core columns the side kernel left behind
#pragma unroll
      for (int k = 0; k < MAX_LEAD; ++k) if (k < nl) { const double v = F[k * TILE]; rc = fma(v, v, rc); }
    }
    double T[NP];
    matvec<NP>(dm, A, na, T);
    // One straight-line block for everything that needs the finished T: rho_t = T.phi and the three gradient dots, eight
    // independent chains with all their shared-memory loads free to move to the top (three separate wait -> load -> chain
    // phases cost as much as the 400-DFMA product itself: LDS latency and 10-deep dependent chains, one warp per scheduler)
    RingPos rq[4];
    rq[0] = rp;
#pragma unroll
    for (int a = 1; a < 4; ++a) { rq[a] = rq[a - 1]; rq[a].next(nslots); }
    const double* G[3] = {F, F, F};
    if (GGA) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {   // chunks 1..3: grad phi
        mbar_wait(c.bar0 + 8 * rq[a + 1].slot, (uint32_t)rq[a + 1].phase);
        G[a] = ring_base() + (size_t)rq[a + 1].slot * c.sdoubles + pt;
      }
    }
    double s0 = 0.0, s1 = 0.0, ga[3] = {0.0, 0.0, 0.0}, gb[3] = {0.0, 0.0, 0.0}, cc[3] = {cx, cy, cz};
    if (GGA && nl > 0) {
#pragma unroll
      for (int k = 0; k < MAX_LEAD; ++k) {
        if (k < nl) {
#pragma unroll
          for (int a = 0; a < 3; ++a) cc[a] = fma(F[k * TILE], G[a][k * TILE], cc[a]);   // (the phi chunk is still held)
        }
      }
    }
#pragma unroll
    for (int t = 0; t < NP; t += 2) {
      if (t < na) {
        s0 = fma(T[t], A[t * TILE], s0);
        if (GGA) {
#pragma unroll
          for (int a = 0; a < 3; ++a) ga[a] = fma(T[t], G[a][(nl + t) * TILE], ga[a]);
        }
      }
      if (t + 1 < na) {
        s1 = fma(T[t + 1], A[(t + 1) * TILE], s1);
        if (GGA) {
#pragma unroll
          for (int a = 0; a < 3; ++a) gb[a] = fma(T[t + 1], G[a][(nl + t + 1) * TILE], gb[a]);
        }
      }
    }
    double g[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) g[a] = 2.0 * ((cc[a] + cc[a]) + (ga[a] + gb[a]));   // grad rho = grad rho_a + grad rho_b (mcpdft.cc:381-383), grad rho_s = 2 (c + T_s . grad phi)
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(c.bar0 + 8 * (nslots + rq[0].slot));
#pragma unroll
      for (int a = 1; a < (GGA ? 4 : 1); ++a) {   // (twice: the difference warps do not visit the gradient chunks)
        mbar_arrive(c.bar0 + 8 * (nslots + rq[a].slot));
        mbar_arrive(c.bar0 + 8 * (nslots + rq[a].slot));
      }
    }
    rp = GGA ? rq[3] : rq[0];
    rp.next(nslots);
    const double st = s0 + s1;                  // rho_t,a + rho_t,b
    const double rho = (rc + rc) + st;          // mcpdft.cc:89 (rho_a + rho_b)
    const double pic = rc * (rc + st);          // rc^2 + rc rho_t,b + rho_t,a rc
    if (inside) {
      P.rho[p] = rho;
      if (GGA) { P.gx[p] = g[0]; P.gy[p] = g[1]; P.gz[p] = g[2]; }
      if (!fused && P.pi_core) P.pi_core[p] = pic;
    }
    if (fused) {   // hand the point to the partner warp (which is one tile behind)
      const int b = it & 1, j = it >> 1;        // j-th use of buffer b
      if (j > 0) mbar_wait(c.xbar0 + 8 * (8 + 4 * b + quarter), (uint32_t)((j - 1) & 1));   // it has taken tile it - 2 out of this buffer
      double* E = ring_base() + (size_t)nslots * c.sdoubles + (size_t)b * 5 * TILE + pt;
      E[0] = rho; E[TILE] = g[0]; E[2 * TILE] = g[1]; E[3 * TILE] = g[2]; E[4 * TILE] = pic;
      mbar_arrive(c.xbar0 + 8 * (4 * b + quarter));
    }
    const double half = fma(0.5, st, rc);       // (rho_a + rho_b) / 2
    acc[0] = fma(rho, wp, acc[0]);              // mcpdft.cc:91-93
    acc[1] = fma(half, wp, acc[1]);
    acc[2] = fma(half, wp, acc[2]);
  }
  acc_out[0] = acc[0]; acc_out[1] = acc[1]; acc_out[2] = acc[2];
}

// ---------------------------------------------------------------- role 1: spin-difference density and XC of every other tile
// Eight warps: (quarter, parity).  A warp owns the tiles with it % 2 == parity of its quarter: the difference density from the
// tile's phi chunk (the only chunk it touches; the sum role arrives twice on the gradient chunks' empty barriers instead),
// then -- fused mode -- R, translation and functional of the same tile once its sum-role partner has handed over rho, grad rho
// and the core part of Pi.  One warp's XC is a ~7 k-cycle dependent chain against a ~5 k-cycle tile period at HBM speed,
// hence two warps per quarter (with four, the fused kernel took 0.57 instead of 0.30 ms per 2 M points).
template <int NP, bool GGA>
__device__ __noinline__ void diff_role(const Params& P_, const Ctx& c_, const int quarter, const int parity, double (&acc_out)[3], double (&acc3_out)[5]) {
  const Params P = P_;   // (register copies, see producer_role)
  const Ctx c = c_;
  constexpr int NARR = GGA ? 4 : 1, NVEC = GGA ? 5 : 2;
  double acc[3] = {0.0, 0.0, 0.0}, acc3[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int lane = threadIdx.x & 31, pt = quarter * 32 + lane;
  const int na = P.nact, nl = P.nlead, nslots = c.nslots, ncols = c.ncols;
  const bool fused = P.xc >= 0;
  const double* dm = ring_base() + (size_t)nslots * c.sdoubles + EXCH_DOUBLES + NP * NP;   // Dsa - Dsb
  for (int it = parity; it < c.nlocal; it += 2) {
    const size_t p = (size_t)((int)blockIdx.x + it * (int)gridDim.x) * TILE + pt;
    const bool inside = p < P.npts;
    double pi_act = 0.0;                       // K2's Pi_act: in flight during the difference density
    if (fused && inside) pi_act = P.pi[p];
    const int slot = (NARR * it) % nslots;      // the tile's phi chunk; it has landed when the sum-role partner says so
    mbar_wait(c.xbar0 + 8 * (16 + 4 * parity + quarter), (uint32_t)((it >> 1) & 1));
    mbar_arrive(c.xbar0 + 8 * (24 + 4 * parity + quarter));
    const double* F = ring_base() + (size_t)slot * c.sdoubles + pt;
    const double wp = inside ? F[(size_t)(ncols + NVEC - 1) * TILE] : 0.0;
    if (P.have_diff) {
      const double* A = F + (size_t)nl * TILE;
      double T[NP];
      matvec<NP>(dm, A, na, T);
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int t = 0; t < NP; t += 2) {
        if (t < na) d0 = fma(T[t], A[t * TILE], d0);
        if (t + 1 < na) d1 = fma(T[t + 1], A[(t + 1) * TILE], d1);
      }
      const double hd = 0.5 * (d0 + d1) * wp;     // (rho_a - rho_b) w / 2
      acc[1] += hd;
      acc[2] -= hd;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(c.bar0 + 8 * (nslots + slot));
    if (fused) {
      const int j = it >> 1;                      // j-th use of hand-over buffer `parity`
      mbar_wait(c.xbar0 + 8 * (4 * parity + quarter), (uint32_t)(j & 1));
      const double* E = ring_base() + (size_t)nslots * c.sdoubles + (size_t)parity * 5 * TILE + pt;
      const double rho = E[0], gx = E[TILE], gy = E[2 * TILE], gz = E[3 * TILE], pic = E[4 * TILE];
      mbar_arrive(c.xbar0 + 8 * (8 + 4 * parity + quarter));
      const double pi = pi_act + pic;             // same association as K2's own store / k3_translate_xc
      if (inside) { P.pi[p] = pi; xc_point<GGA>(P.xc, P.omega, rho, pi, gx, gy, gz, wp, acc3); }
    }
  }
  acc_out[1] = acc[1]; acc_out[2] = acc[2];
#pragma unroll
  for (int k = 0; k < 5; ++k) acc3_out[k] = acc3[k];
}

template <int NP, bool GGA>
__global__ void __launch_bounds__(THREADS, 1) k1_tile(const __grid_constant__ Maps maps, const __grid_constant__ Params P) {
  __shared__ double scratch[5 * 32];
  Ctx c;
  c.ncols = P.nlead + P.nact; c.nslots = P.nslots;
  c.sdoubles = slot_doubles(c.ncols, GGA);
  double* const dm = ring_base() + (size_t)c.nslots * c.sdoubles + EXCH_DOUBLES;
  c.bar0 = smem_u32(dm + dmat_doubles(NP));
  c.xbar0 = c.bar0 + 16 * c.nslots;
  c.nlocal = P.ntiles > (int)blockIdx.x ? (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < dmat_doubles(NP); i += THREADS) {   // [role][u][t] <- dmat[role][u * MAX_ACT + t]
    const int role = i / (NP * NP), u = (i % (NP * NP)) / NP, t = i % NP;
    dm[i] = P.dmat[(size_t)role * MAX_ACT * MAX_ACT + u * MAX_ACT + t];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.nslots; ++s) { mbar_init(c.bar0 + 8 * s, 1); mbar_init(c.bar0 + 8 * (c.nslots + s), EMPTY_ARRIVALS); }
    for (int s = 0; s < 32; ++s) mbar_init(c.xbar0 + 8 * s, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double acc[3] = {0.0, 0.0, 0.0};
  double acc3[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (warp == COMPUTE_WARPS) {
    if (lane == 0) producer_role<GGA>(maps, P, c);
  } else if (warp < 4) {
    sum_role<NP, GGA>(P, c, warp & 3, acc);
  } else {
    diff_role<NP, GGA>(P, c, warp & 3, (warp - 4) >> 2, acc, acc3);
  }
  __syncthreads();
  ordm::block_sum_store<3>(acc, scratch, P.partial + (size_t)blockIdx.x * 3);
  if (P.xc >= 0) {
    __syncthreads();
    ordm::block_sum_store<5>(acc3, scratch, P.partial3 + (size_t)blockIdx.x * 5);
  }
}

}  // namespace k1t
