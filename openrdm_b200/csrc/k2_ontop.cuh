// k2_ontop.cuh -- K2: on-top pair density Pi(r) on the FP64 tensor pipe.
// Replaces MCPDFT::build_ontop_pair_density (mcpdft.cc:185-212): the reference evaluates
//     Pi(p) = sum_{mu nu la si} phi_mu phi_nu phi_la phi_si D2ab(nu*n+mu, si*n+la)     (n^4 / point)
// Here the orbital-pair products X(p,I) = phi_t(p) phi_u(p), I=(t<=u),
// M = n(n+1)/2, are formed once per point tile in shared memory and contracted with the
// host-folded symmetric matrix Dt (M x M):   Pi = rowsum( X o (X . Dt) ).
// Because Dt is symmetric only the block-upper-triangular part is multiplied: Dt is cut
// into 32x32 blocks, off-diagonal blocks are doubled, blocks below the diagonal are skipped
// (M(M+1) instead of 2M^2 flops per point).
//
// Mapping to sm_100a: there is no FP64 kind of tcgen05.mma (nor wgmma); FP64 tensor work is
// mma.sync, which ptxas lowers to DMMA.8x8x4 for every .f64 shape on sm_100a -- so the
// m8n8k4 shape is used directly.  Measured ceiling on B200: 37.1 TFLOP/s (DMMA issue),
// cuBLAS DGEMM 35.5 TFLOP/s (profiles/r01_fp64_peaks_microbench.log).
//
//  * one persistent CTA (8 warps) per SM walks point tiles of P = 64/32/16/8 points;
//  * the active orbital columns of the tile arrive by TMA bulk copies (cp.async.bulk +
//    mbarrier, one 8*P-byte copy per column; next tile prefetched during the math);
//  * X lives in shared memory as [I][P+4] (the +4 pad makes the DMMA A-fragment loads
//    bank-conflict free: lane -> (I0 + lane%4, p0 + lane/4));
//  * the unit of work is a "task" = one 32x32 block (ib<=jb) of Dt' applied to a 32-point
//    group: C(32x32) = X[:,ib] . Dt'[ib,jb] as 8 k-steps x 16 DMMA.8x8x4, followed by the
//    fused epilogue  pi[p] += sum_j C[p,j] * X[p, jb*32+j]  on the accumulator fragments;
//  * Dt' is pre-packed on the host in DMMA B-fragment order, per task lane, so each k-step
//    of B is one coalesced 1 KB read (2 x LDG.128 per thread) straight from L2 into
//    registers, software-pipelined 3 k-steps ahead -- Dt' (<= 1 MB) stays L2-resident;
//  * per-point results: quad shuffle reduction, then a fixed-order cross-warp sum in smem.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace k2 {

constexpr int BLK = 32;            // Dt block edge
constexpr int KSTEPS = BLK / 4;    // DMMA k-steps per task
constexpr int NTILE = BLK / 8;     // DMMA n-tiles per task
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int STEP_DOUBLES = 32 * NTILE;            // B doubles per k-step (1 KB)
constexpr int TASK_DOUBLES = KSTEPS * STEP_DOUBLES; // B doubles per task (8 KB)
constexpr int PREFETCH = 3;                         // B k-steps in flight
constexpr int RING = PREFETCH + 1;                  // register ring slots (power of two)
static_assert((RING & (RING - 1)) == 0 && KSTEPS % RING == 0, "ring must tile the k-steps");
constexpr int MAX_LANES = 8;

struct Params {
  const double* phi_act;  // first active column, [nact][ld]
  size_t ld;
  size_t npts;
  int nact;
  int M;       // nact(nact+1)/2
  int nblk;    // ceil(M/32)
  int ntask;   // nblk(nblk+1)/2
  int ntiles;  // ceil(npts/P)
  const double* dfrag;            // packed Dt' (fragment order, per lane, zero-padded tail)
  const uint16_t* pairs;          // [Mpad][2] : (t,u) of pair I, (0,0) beyond M
  uint32_t lane_ofs[MAX_LANES];   // first task slot of each task lane in dfrag
  const double* pi_core;          // nullable: closed-form core part added to the result
  double* pi;                     // out, npts
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// shared-memory carve-up (doubles unless noted)
template <int P>
struct Smem {
  static constexpr int PS = P + 4;  // X row stride
  static __host__ __device__ size_t xs_doubles(int Mpad) { return (size_t)Mpad * PS; }
  static __host__ __device__ size_t bytes(int Mpad, int nact) {
    return sizeof(double) * (xs_doubles(Mpad) + (size_t)nact * P + (size_t)WARPS * P) + sizeof(uint16_t) * 2 * Mpad + 16;
  }
};

// NPG point groups per tile (P = 8*MT*NPG points), MT DMMA m-tiles per warp.
// warp w: point group w % NPG ... task lane w / NPG  (NL = WARPS/NPG task lanes).
template <int NPG, int MT>
__global__ void __launch_bounds__(THREADS, 1) k2_ontop(const Params prm) {
  constexpr int P = 8 * MT * NPG;
  constexpr int PS = P + 4;
  constexpr int NL = WARPS / NPG;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int Mpad = prm.nblk * BLK;
  double* Xs = reinterpret_cast<double*>(smem_raw);
  double* phis = Xs + (size_t)Mpad * PS;                 // [nact][P]
  double* red = phis + (size_t)prm.nact * P;             // [NL][P] cross-warp partials
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + (size_t)WARPS * P);
  uint16_t* pairs = reinterpret_cast<uint16_t*>(bar + 2);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int pg = wid % NPG, tl = wid / NPG;
  const int qrow = lane >> 2, qcol = lane & 3;
  const uint32_t tile_bytes = (uint32_t)(prm.nact * P * sizeof(double));

  for (int i = tid; i < 2 * Mpad; i += THREADS) pairs[i] = prm.pairs[i];
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  int tile = blockIdx.x;
  if (tid == 0 && tile < prm.ntiles) {
    mbar_expect_tx(bar, tile_bytes);
    for (int t = 0; t < prm.nact; ++t)
      tma_bulk_g2s(phis + (size_t)t * P, prm.phi_act + (size_t)t * prm.ld + (size_t)tile * P, P * sizeof(double), bar);
  }
  // number of tasks of this warp's lane: tasks tl, tl+NL, ...
  const int my_tasks = (prm.ntask > tl) ? (prm.ntask - tl + NL - 1) / NL : 0;
  const double* bbase = prm.dfrag + (size_t)prm.lane_ofs[tl] * TASK_DOUBLES + lane * NTILE;

  uint32_t parity = 0;
  for (; tile < prm.ntiles; tile += gridDim.x) {
    mbar_wait(bar, parity);
    parity ^= 1;
    // ---- X[I][p] = phi_t[p] * phi_u[p] ----
    for (int idx = tid; idx < Mpad * P; idx += THREADS) {
      const int I = idx / P, p = idx % P;
      const int t = pairs[2 * I], u = pairs[2 * I + 1];
      const double v = (I < prm.M) ? phis[t * P + p] * phis[u * P + p] : 0.0;
      Xs[(size_t)I * PS + p] = v;
    }
    __syncthreads();
    // phis is free again: prefetch the next tile's columns while this one is contracted
    {
      const int nxt = tile + gridDim.x;
      if (tid == 0 && nxt < prm.ntiles) {
        mbar_expect_tx(bar, tile_bytes);
        for (int t = 0; t < prm.nact; ++t)
          tma_bulk_g2s(phis + (size_t)t * P, prm.phi_act + (size_t)t * prm.ld + (size_t)nxt * P, P * sizeof(double), bar);
      }
    }
    // ---- contraction over this warp's tasks ----
    double piacc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) piacc[m] = 0.0;
    const double* xrow = Xs + pg * (8 * MT) + qrow;  // + I*PS + m*8
    const double* bptr = bbase;
    double bq[RING][NTILE];
#pragma unroll
    for (int s = 0; s < PREFETCH; ++s) {
      const double2 lo = __ldg(reinterpret_cast<const double2*>(bptr + s * STEP_DOUBLES));
      const double2 hi = __ldg(reinterpret_cast<const double2*>(bptr + s * STEP_DOUBLES) + 1);
      bq[s][0] = lo.x; bq[s][1] = lo.y; bq[s][2] = hi.x; bq[s][3] = hi.y;
    }
    int ib = 0, jb = 0;
    {  // (ib,jb) of task index tl in the order jb outer, ib<=jb inner
      int rem = tl;
      while (rem > jb) { rem -= jb + 1; ++jb; }
      ib = rem;
    }
    for (int task = 0; task < my_tasks; ++task) {
      double c[MT][NTILE][2];
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NTILE; ++n) c[m][n][0] = c[m][n][1] = 0.0;
      const double* xa = xrow + (size_t)(ib * BLK + qcol) * PS;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        {  // prefetch B of k-step ks+PREFETCH (runs into the zero-padded tail at the very end)
          const double2* src = reinterpret_cast<const double2*>(bptr + (ks + PREFETCH) * STEP_DOUBLES);
          const double2 lo = __ldg(src), hi = __ldg(src + 1);
          bq[(ks + PREFETCH) & (RING - 1)][0] = lo.x; bq[(ks + PREFETCH) & (RING - 1)][1] = lo.y;
          bq[(ks + PREFETCH) & (RING - 1)][2] = hi.x; bq[(ks + PREFETCH) & (RING - 1)][3] = hi.y;
        }
        double a[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) a[m] = xa[(size_t)(ks * 4) * PS + m * 8];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
          for (int n = 0; n < NTILE; ++n) dmma884(c[m][n][0], c[m][n][1], a[m], bq[ks & (RING - 1)][n]);
      }
      bptr += TASK_DOUBLES;
      // fused epilogue: pi[p] += sum_j C[p][j] * X[jb*32 + j][p]; C fragment = rows qrow, cols 2*qcol+{0,1}
      const double* xe = xrow + (size_t)(jb * BLK + 2 * qcol) * PS;
#pragma unroll
      for (int n = 0; n < NTILE; ++n)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          piacc[m] = fma(c[m][n][0], xe[(size_t)(n * 8) * PS + m * 8], piacc[m]);
          piacc[m] = fma(c[m][n][1], xe[(size_t)(n * 8 + 1) * PS + m * 8], piacc[m]);
        }
      // next task of this lane
      ib += NL;
      while (ib > jb) { ib -= jb + 1; ++jb; }
    }
    // ---- reduce: 4 lanes of a quad hold the same rows ----
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      double v = piacc[m];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (qcol == 0) red[(size_t)tl * P + pg * (8 * MT) + m * 8 + qrow] = v;
    }
    __syncthreads();  // all warps done with Xs and red complete
    if (tid < P) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < NL; ++l) s += red[(size_t)l * P + tid];
      const size_t gp = (size_t)tile * P + tid;
      if (gp < prm.npts) {
        if (prm.pi_core) s += prm.pi_core[gp];
        prm.pi[gp] = s;
      }
    }
    // the next iteration's X build must not start before `red` was consumed: the
    // __syncthreads() after the X build orders it (red is only rewritten after that).
  }
}

}  // namespace k2
