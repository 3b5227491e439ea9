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
//  * the unit of work is a "segment" = a run of row blocks ib0..ib0+n-1 of one 32-wide block
//    column jb of Dt' applied to a 32-point group: C(32x32) += X[:,ib] . Dt'[ib,jb], 8 k-steps
//    x 16 DMMA.8x8x4 per block, all blocks of the run accumulated in registers, followed by
//    ONE fused epilogue  pi[p] += sum_j C[p,j] * X[p, jb*32+j]  on the accumulator fragments.
//    The diagonal block only issues the sub-tiles that touch the upper triangle and a ragged
//    last column only its live n-tiles (both compile-time patterns: the DMMA stream of a
//    segment is branch-free).  Segments are balanced over the task lanes on the host;
//  * Dt' is pre-packed on the host in DMMA B-fragment order, per task lane, so each k-step
//    of B is one coalesced 1 KB read (2 x LDG.128 per thread) straight from L2 into
//    registers, software-pipelined 3 k-steps ahead -- Dt' (<= 1 MB) stays L2-resident;
//    A fragments are double-buffered one k-step ahead;
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

struct SegDesc {  // same layout as ordm::SegDesc (rdm_pack.h)
  uint16_t jb, ib0, nblk;
  uint8_t has_diag, ntl;
};

struct Params {
  const double* phi_act;  // first active column, [nact][ld]
  size_t ld;
  size_t npts;
  int nact;
  int M;       // nact(nact+1)/2
  int nblk;    // ceil(M/32)
  int ntiles;  // ceil(npts/P)
  const double* dfrag;              // packed Dt' (fragment order, per task lane, zero-padded tail)
  const SegDesc* segs;              // segment descriptors, lane-major (+1 sentinel)
  const uint16_t* pairs;            // [M][2]: (t,u), t<=u, of packed pair index I (row-major upper triangle)
  uint32_t seg_ofs[MAX_LANES + 1];  // lane l owns segments [seg_ofs[l], seg_ofs[l+1])
  uint32_t blk_ofs[MAX_LANES + 1];  // first 8 KB block slot of lane l in dfrag
  const double* pi_core;            // nullable: closed-form core part added to the result
  double* pi;                       // out, npts
  // grad-Pi variant only (k2_ontop_ws<true, *>): active columns of grad phi, closed-form core
  // parts of grad Pi (nullable) and the three outputs
  const double* dphi_act[3];
  const double* pi_core_g[3];
  double* pi_g[3];
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
  static __host__ __device__ size_t xs_doubles(int Mpad) { return (size_t)(Mpad + 4) * PS; }  // +4 rows: A prefetch overrun
  static __host__ __device__ size_t bytes(int Mpad, int nact) {
    return sizeof(double) * (xs_doubles(Mpad) + (size_t)nact * P + (size_t)WARPS * P) + 16 + sizeof(uint32_t) * Mpad;
  }
};

// One 32-row block of a segment: 8 k-steps; B ring prefetched PREFETCH k-steps ahead, A fragments
// one k-step ahead (the row pointer simply runs on into the next block of the run).
// NTL = live n-tiles of the block column, DIAG = diagonal block (static sub-tile pattern).
template <int MT, int PS, int NTL, bool DIAG>
__device__ __forceinline__ void kblock(double (&c)[MT][NTILE][2], double (&bq)[RING][NTILE], double (&a)[2][MT],
                                       const double*& xa, const double*& bptr) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
#ifndef K2_EXP_NOBLOAD  // (tools/k2lab experiment switches; never defined in the product build)
    {
      const double2* src = reinterpret_cast<const double2*>(bptr + (ks + PREFETCH) * STEP_DOUBLES);
      const double2 lo = __ldg(src), hi = __ldg(src + 1);
      bq[(ks + PREFETCH) & (RING - 1)][0] = lo.x; bq[(ks + PREFETCH) & (RING - 1)][1] = lo.y;
      bq[(ks + PREFETCH) & (RING - 1)][2] = hi.x; bq[(ks + PREFETCH) & (RING - 1)][3] = hi.y;
    }
#endif
#ifndef K2_EXP_NOALOAD
#pragma unroll
    for (int m = 0; m < MT; ++m) a[(ks + 1) & 1][m] = xa[(size_t)((ks + 1) * 4) * PS + m * 8];
#endif
#pragma unroll
    for (int n = 0; n < NTL; ++n)
#ifdef K2_EXP_NODIAGSKIP
      {
#else
      if (!DIAG || 4 * ks <= 8 * n + 7) {
#endif
#pragma unroll
        for (int m = 0; m < MT; ++m) dmma884(c[m][n][0], c[m][n][1], a[ks & 1][m], bq[ks & (RING - 1)][n]);
      }
  }
  xa += (size_t)BLK * PS;
  bptr += TASK_DOUBLES;
}

template <int MT, int PS, int NTL, bool GRAD = false>
__device__ __forceinline__ void run_segment(const SegDesc sd, double (&piacc)[GRAD ? 4 : 1][MT], double (&bq)[RING][NTILE],
                                            const double* xrow, int qcol, const double*& bptr,
                                            const double* ph = nullptr, const uint32_t* pofs = nullptr, int cstride = 0) {
  double c[MT][NTILE][2];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NTILE; ++n) c[m][n][0] = c[m][n][1] = 0.0;
  const double* xa = xrow + (size_t)(sd.ib0 * BLK + qcol) * PS;
  double a[2][MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) a[0][m] = xa[m * 8];
  const int nfull = (int)sd.nblk - (int)sd.has_diag;
  for (int b = 0; b < nfull; ++b) kblock<MT, PS, NTL, false>(c, bq, a, xa, bptr);
  if (sd.has_diag) kblock<MT, PS, NTL, true>(c, bq, a, xa, bptr);
  // fused epilogue: pi[p] += sum_j C[p][j] * X[jb*32 + j][p].  The C fragment of a lane is row
  // qrow, DMMA columns 2*qcol+{0,1} = (host-permuted) J offsets qcol and qcol+4 of the n-tile.
  const double* xe = xrow + (size_t)(sd.jb * BLK + qcol) * PS;
#ifdef K2_EXP_NOEPI
#pragma unroll
  for (int n = 0; n < NTL; ++n)
#pragma unroll
    for (int m = 0; m < MT; ++m) piacc[0][m] += c[m][n][0] + c[m][n][1];
  (void)xe;
#else
  if (!GRAD) {
#pragma unroll
    for (int n = 0; n < NTL; ++n)
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        piacc[0][m] = fma(c[m][n][0], xe[(size_t)(n * 8) * PS + m * 8], piacc[0][m]);
        piacc[0][m] = fma(c[m][n][1], xe[(size_t)(n * 8 + 4) * PS + m * 8], piacc[0][m]);
      }
  } else {
    // C = Y[:, jb] (full Dt).  Pi += Y o X and d_c Pi += Y o d_c X with
    // d_c X(p, J=(t,u)) = d_c phi_t phi_u + phi_t d_c phi_u formed on the fly from the phi /
    // grad phi tiles (ph: [4][nact][P], component stride cstride); the factor 2 of
    // grad Pi = 2 rowsum(dX o Y) is applied once at the store.
    (void)xe;
#pragma unroll
    for (int n = 0; n < NTL; ++n)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t po = pofs[sd.jb * BLK + n * 8 + 4 * h + qcol];
        const double* ft = ph + (po & 0xffffu);
        const double* fu = ph + (po >> 16);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const double cc = c[m][n][h];
          const double a = ft[m * 8], b = fu[m * 8];
          piacc[0][m] = fma(cc, a * b, piacc[0][m]);
#pragma unroll
          for (int d = 1; d < 4; ++d) {
            const double da = ft[d * cstride + m * 8], db = fu[d * cstride + m * 8];
            piacc[GRAD ? d : 0][m] = fma(cc, fma(da, b, a * db), piacc[GRAD ? d : 0][m]);
          }
        }
      }
  }
#endif
}

// NPG point groups per tile (P = 8*MT*NPG points), MT DMMA m-tiles per warp.
// warp w: point group w % NPG, task lane w / NPG  (NL = WARPS/NPG task lanes).
template <int NPG, int MT>
__global__ void __launch_bounds__(THREADS, 1) k2_ontop(const Params prm) {
  constexpr int P = 8 * MT * NPG;
  constexpr int PS = P + 4;
  constexpr int NL = WARPS / NPG;
  constexpr int G = THREADS / P;  // pair rows covered per pass of the X build
  constexpr int XU = 8;           // independent products in flight per thread
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int Mpad = prm.nblk * BLK;
  double* Xs = reinterpret_cast<double*>(smem_raw);
  double* phis = Xs + (size_t)(Mpad + 4) * PS;           // [nact][P]
  double* red = phis + (size_t)prm.nact * P;             // [NL][P] cross-warp partials
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + (size_t)WARPS * P);
  uint32_t* pofs = reinterpret_cast<uint32_t*>(bar + 2);  // [M]: (t*P) | (u*P) << 16

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int pg = wid % NPG, tl = wid / NPG;
  const int qrow = lane >> 2, qcol = lane & 3;
  const int nact = prm.nact;
  const uint32_t tile_bytes = (uint32_t)(nact * P * sizeof(double));

  // rows M..Mpad+3 of X never change: zero them once
  for (int idx = prm.M * P + tid; idx < (Mpad + 4) * P; idx += THREADS) Xs[(size_t)(idx / P) * PS + (idx % P)] = 0.0;
  for (int I = tid; I < prm.M; I += THREADS)
    pofs[I] = (uint32_t)prm.pairs[2 * I] * P | ((uint32_t)prm.pairs[2 * I + 1] * P) << 16;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  int tile = blockIdx.x;
  if (tid == 0 && tile < prm.ntiles) {
    mbar_expect_tx(bar, tile_bytes);
    for (int t = 0; t < nact; ++t)
      tma_bulk_g2s(phis + (size_t)t * P, prm.phi_act + (size_t)t * prm.ld + (size_t)tile * P, P * sizeof(double), bar);
  }
  const int seg0 = (int)prm.seg_ofs[tl], nseg = (int)prm.seg_ofs[tl + 1] - seg0;
  const double* bbase = prm.dfrag + (size_t)prm.blk_ofs[tl] * TASK_DOUBLES + lane * NTILE;
  const SegDesc* sdesc = prm.segs + seg0;
  const int xp = tid % P, xg = tid / P;

  uint32_t parity = 0;
  for (; tile < prm.ntiles; tile += gridDim.x) {
    mbar_wait(bar, parity);
    parity ^= 1;
    // ---- X[I][p] = phi_t[p] * phi_u[p]: thread owns point xp and pair rows xg, xg+G, ...;
    //      XU independent (table -> two phi loads -> product -> store) chains in flight ----
    for (int I0 = xg; I0 < prm.M; I0 += G * XU) {
      uint32_t po[XU];
      double ft[XU], fu[XU];
#pragma unroll
      for (int j = 0; j < XU; ++j) po[j] = (I0 + j * G < prm.M) ? pofs[I0 + j * G] : 0u;
#pragma unroll
      for (int j = 0; j < XU; ++j) { ft[j] = phis[(po[j] & 0xffffu) + xp]; fu[j] = phis[(po[j] >> 16) + xp]; }
#pragma unroll
      for (int j = 0; j < XU; ++j)
        if (I0 + j * G < prm.M) Xs[(size_t)(I0 + j * G) * PS + xp] = ft[j] * fu[j];
    }
    __syncthreads();
    // phis is free again: prefetch the next tile's columns while this one is contracted
    {
      const int nxt = tile + gridDim.x;
      if (tid == 0 && nxt < prm.ntiles) {
        mbar_expect_tx(bar, tile_bytes);
        for (int t = 0; t < nact; ++t)
          tma_bulk_g2s(phis + (size_t)t * P, prm.phi_act + (size_t)t * prm.ld + (size_t)nxt * P, P * sizeof(double), bar);
      }
    }
    // ---- contraction over this warp's segments ----
    double piacc[1][MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) piacc[0][m] = 0.0;
    const double* xrow = Xs + pg * (8 * MT) + qrow;  // + I*PS + m*8
    const double* bptr = bbase;
    double bq[RING][NTILE];
#pragma unroll
    for (int s = 0; s < PREFETCH; ++s) {
      const double2 lo = __ldg(reinterpret_cast<const double2*>(bptr + s * STEP_DOUBLES));
      const double2 hi = __ldg(reinterpret_cast<const double2*>(bptr + s * STEP_DOUBLES) + 1);
      bq[s][0] = lo.x; bq[s][1] = lo.y; bq[s][2] = hi.x; bq[s][3] = hi.y;
    }
    SegDesc sd = sdesc[0];
    for (int sg = 0; sg < nseg; ++sg) {
      const SegDesc sd_next = sdesc[sg + 1];  // sentinel-terminated
      switch (sd.ntl) {
        case 4: run_segment<MT, PS, 4>(sd, piacc, bq, xrow, qcol, bptr); break;
        case 3: run_segment<MT, PS, 3>(sd, piacc, bq, xrow, qcol, bptr); break;
        case 2: run_segment<MT, PS, 2>(sd, piacc, bq, xrow, qcol, bptr); break;
        default: run_segment<MT, PS, 1>(sd, piacc, bq, xrow, qcol, bptr); break;
      }
      sd = sd_next;
    }
    // ---- reduce: 4 lanes of a quad hold the same rows ----
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      double v = piacc[0][m];
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
    // `red` is only rewritten after the next tile's post-X-build __syncthreads(), which the
    // readers above also have to pass: no extra barrier needed.
  }
}

// =============================================================================================
// Warp-specialised variant (used whenever its X buffers fit shared memory).
//
// 12 warps = 3 warpgroups.  Warpgroup 2 ("builders", registers trimmed with setmaxnreg) owns
// everything that is not a DMMA: TMA loads of the phi tile, the X = phi_t*phi_u build into a
// (double-)buffered X, the final fixed-order cross-lane sum and the store of Pi.  Warpgroups 0-1
// (8 "MMA" warps, registers raised with setmaxnreg) only stream segments: each warp is one task
// lane over all 32 points of the tile.  Builders and MMA warps hand X buffers back and forth
// through mbarriers (full[b]: X[b] ready; done[b]: all 8 lanes finished with X[b], partials in
// red[b]), so the X build of tile k+1 and the epilogue/store of tile k-1 overlap the DMMA stream
// of tile k and no CTA-wide barrier is left in the steady state.
//
// GRAD = true additionally produces grad Pi (MCPDFT::build_ontop_pair_density_gradients,
// mcpdft.cc:214-270) from the same product Y = X . Dt:  d_c Pi = 2 rowsum(d_c X o Y)
// (SURVEY.md App. B.1).  That needs Y itself, so the whole symmetric Dt is streamed (2 M^2
// flops per point instead of M(M+1)); the grad phi tiles arrive by TMA next to phi and d_c X is
// formed on the fly in the segment epilogue.  NBUF = 1 (single X buffer) is the fallback when
// two do not fit shared memory.
// =============================================================================================
#ifdef K2_EXP_TRACE
__device__ long long g_trace[12][64][4];  // [warp][tile][event] clock64 stamps of block 0
#define K2_TRACE(ev) do { if (blockIdx.x == 0 && lane == 0 && k < 64) g_trace[wid][k][ev] = clock64(); } while (0)
#else
#define K2_TRACE(ev) do { } while (0)
#endif
constexpr int WS_THREADS = 384;
constexpr int WS_P = 32;         // points per tile
constexpr int WS_PS = WS_P + 4;  // X row stride
constexpr int WS_BUILDERS = 128;

struct WsSmem {
  static __host__ __device__ size_t x_doubles(int Mpad) { return (size_t)(Mpad + 4) * WS_PS; }
  // grad: the phi / grad phi tile rows are padded to WS_PS doubles -- the epilogue that forms dX on the
  // fly reads four different orbital rows per instruction, and with a 32-double stride they all fall
  // on the same banks (4-way conflicts, 170 M per launch in ncu); stride 36 makes it 2-way, the minimum
  static __host__ __device__ size_t bytes(int Mpad, int nact, bool grad = false, int nbuf = 2) {
    const size_t nc = grad ? 4 : 1, ph = grad ? WS_PS : WS_P;
    return sizeof(double) * ((size_t)nbuf * x_doubles(Mpad) + (size_t)nbuf * nc * nact * ph + (size_t)nbuf * nc * WARPS * WS_P) + 64 +
           sizeof(uint32_t) * Mpad;
  }
};

template <int NREG> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(NREG)); }
template <int NREG> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(NREG)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// The CTA body, warps 0..11.  MMA_REGS = register budget of the MMA warps (224 when the CTA owns
// the whole register file, 160 in the fused K1+K2 kernel, which the DMMA stream does not exceed).
// Threads beyond WS_THREADS (the density warps of the fused kernel) only take part in the prologue
// barrier and return.
template <bool GRAD, int NBUF, int MMA_REGS, int BW0 = WARPS>  // BW0 = first builder warp (lab: 12 puts the builders last)
__device__ __forceinline__ void ws_cta(const Params& prm) {
  constexpr int P = WS_P, PS = WS_PS, MT = 4;
  constexpr int NC = GRAD ? 4 : 1;  // phi (+ grad phi) components staged per tile; result rows per point
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int Mpad = prm.nblk * BLK;
  const int nact = prm.nact;
  double* Xs0 = reinterpret_cast<double*>(smem_raw);
  const size_t xsz = WsSmem::x_doubles(Mpad);
  double* phis0 = Xs0 + NBUF * xsz;                               // [NBUF][NC][nact][P]
  constexpr int PH = GRAD ? PS : P;                               // row stride of the phi tiles (see WsSmem::bytes)
  double* red0 = phis0 + (size_t)NBUF * NC * nact * PH;           // [NBUF][8 lanes][NC][P]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red0 + (size_t)NBUF * NC * WARPS * P);
  uint64_t* full = bars;      // [2] builders -> MMA   (count WS_BUILDERS/32 = 4 warp arrivals)
  uint64_t* done = bars + 2;  // [2] MMA -> builders   (count 8 warp arrivals)
  uint64_t* tma = bars + 4;   // [2] phi tile landed   (count 1 + tx bytes)
  uint32_t* pofs = reinterpret_cast<uint32_t*>(bars + 8);  // [Mpad]: (t*P) | (u*P) << 16, 0 beyond M

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cstride = nact * PH;  // component stride inside a phi tile buffer
  const uint32_t tile_bytes = (uint32_t)(NC * nact * P * sizeof(double));
  const int ntiles = prm.ntiles;
  // tiles of this CTA: blockIdx.x + k*gridDim.x, k = 0..nmine-1
  const int nmine = (ntiles > (int)blockIdx.x) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid < WS_THREADS) {
    for (int I = tid; I < Mpad; I += WS_THREADS)
      pofs[I] = I < prm.M ? ((uint32_t)prm.pairs[2 * I] * PH | ((uint32_t)prm.pairs[2 * I + 1] * PH) << 16) : 0u;
    for (int b = 0; b < NBUF; ++b)  // rows M..Mpad+3 of the X buffers stay zero for the whole launch
      for (int idx = prm.M * P + tid; idx < (Mpad + 4) * P; idx += WS_THREADS) Xs0[b * xsz + (size_t)(idx / P) * PS + (idx % P)] = 0.0;
  }
  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) { mbar_init(&full[b], WS_BUILDERS / 32); mbar_init(&done[b], WARPS); mbar_init(&tma[b], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (BW0 == WARPS ? tid >= WS_THREADS : (wid >= WARPS && wid < BW0)) return;

  // The builders' few FP64 multiplies share the FP64/DMMA pipe with the MMA warps' DMMA stream and
  // queue behind it (~260 cycles per DMUL, whichever warp ids the builders get: K2_EXP_TRACE shows
  // the X build taking ~13 k of the 14.5 k-cycle tile period either way), so X for tile k+1 is
  // ready just as the lanes finish tile k -- see DESIGN.md section 4 for what was tried against it.
  if (wid >= WARPS) {
    // ================================ builders ================================
    reg_dec<56>();
    const int btid = tid - BW0 * 32;  // 0..127 within the builder warpgroup
    auto issue_tma = [&](int k) {  // phi (+ grad phi) tile of my k-th tile into phis[k % NBUF]
      const int b = k % NBUF;
      const size_t t0 = (size_t)(blockIdx.x + (size_t)k * gridDim.x) * P;
      mbar_expect_tx(&tma[b], tile_bytes);
      double* dst = phis0 + (size_t)b * NC * cstride;
      for (int t = 0; t < nact; ++t) {
        tma_bulk_g2s(dst + (size_t)t * PH, prm.phi_act + (size_t)t * prm.ld + t0, P * sizeof(double), &tma[b]);
        if (GRAD) {
#pragma unroll
          for (int d = 0; d < 3; ++d)
            tma_bulk_g2s(dst + (size_t)(d + 1) * cstride + (size_t)t * PH, prm.dphi_act[d] + (size_t)t * prm.ld + t0, P * sizeof(double), &tma[b]);
        }
      }
    };
    auto flush_tile = [&](int k) {  // fixed-order sum of the 8 lanes' partials of tile k, store Pi
      const int b = k % NBUF;
      mbar_wait(&done[b], (uint32_t)((k / NBUF) & 1));
      if (btid < NC * P) {
        const int comp = btid / P, pt = btid % P;
        const double* r = red0 + (size_t)b * NC * WARPS * P + (size_t)comp * P + pt;
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < WARPS; ++l) s += r[(size_t)l * NC * P];
        const size_t gp = (size_t)(blockIdx.x + (size_t)k * gridDim.x) * P + pt;
        if (gp < prm.npts) {
          if (comp == 0) {
            if (prm.pi_core) s += prm.pi_core[gp];
            prm.pi[gp] = s;
          } else {
            s *= 2.0;
            if (prm.pi_core_g[comp - 1]) s += prm.pi_core_g[comp - 1][gp];
            prm.pi_g[comp - 1][gp] = s;
          }
        }
      }
    };
    // plain: the phi tile is only read by the X build, so the next-but-one tile is requested as
    // soon as the builders are through with it.  GRAD: the MMA warps' epilogue reads the tile too,
    // so phis[b] is refilled only after done[b] (inside the loop, right after flush_tile).
    if (btid == 0) { if (nmine > 0) issue_tma(0); if (NBUF > 1 && nmine > 1) issue_tma(1); }
    const int xp = btid % P, xg = btid / P;
    constexpr int G = WS_BUILDERS / P, XU = 8;
    for (int k = 0; k < nmine; ++k) {
      const int b = k % NBUF;
      K2_TRACE(0);
      if (k >= NBUF) {
        flush_tile(k - NBUF);  // also proves X[b] and red[b] (GRAD: phis[b]) are free again
        if (GRAD && btid == 0) issue_tma(k);
      }
      K2_TRACE(1);
      mbar_wait(&tma[b], (uint32_t)((k / NBUF) & 1));
      K2_TRACE(2);
      double* Xs = Xs0 + (size_t)b * xsz;
      const double* phis = phis0 + (size_t)b * NC * cstride;
      for (int I0 = xg; I0 < prm.M; I0 += G * XU) {
        uint32_t po[XU];
        double ft[XU], fu[XU];
#pragma unroll
        for (int j = 0; j < XU; ++j) po[j] = (I0 + j * G < prm.M) ? pofs[I0 + j * G] : 0u;
#pragma unroll
        for (int j = 0; j < XU; ++j) { ft[j] = phis[(po[j] & 0xffffu) + xp]; fu[j] = phis[(po[j] >> 16) + xp]; }
#pragma unroll
        for (int j = 0; j < XU; ++j)
          if (I0 + j * G < prm.M) Xs[(size_t)(I0 + j * G) * PS + xp] = ft[j] * fu[j];
      }
      __syncwarp();
      K2_TRACE(3);
      if (lane == 0) mbar_arrive(&full[b]);
      if (!GRAD) {
        // phis[b] may be refilled once every builder warp has read it
        asm volatile("bar.sync 1, %0;" ::"n"(WS_BUILDERS) : "memory");
        if (btid == 0 && k + NBUF < nmine) issue_tma(k + NBUF);
      }
    }
    for (int k = (nmine > NBUF ? nmine - NBUF : 0); k < nmine; ++k) flush_tile(k);
  } else {
    // ================================ MMA warps ================================
    reg_inc<MMA_REGS>();
    const int tl = wid;  // task lane 0..7
    const int qrow = lane >> 2, qcol = lane & 3;
    const int seg0 = (int)prm.seg_ofs[tl], nseg = (int)prm.seg_ofs[tl + 1] - seg0;
    const double* bbase = prm.dfrag + (size_t)prm.blk_ofs[tl] * TASK_DOUBLES + lane * NTILE;
    const SegDesc* sdesc = prm.segs + seg0;
    for (int k = 0; k < nmine; ++k) {
      const int b = k % NBUF;
      const double* bptr = bbase;
      double bq[RING][NTILE];
#pragma unroll
      for (int s = 0; s < PREFETCH; ++s) {  // B does not depend on the tile: fetch before waiting for X
        const double2 lo = __ldg(reinterpret_cast<const double2*>(bptr + s * STEP_DOUBLES));
        const double2 hi = __ldg(reinterpret_cast<const double2*>(bptr + s * STEP_DOUBLES) + 1);
        bq[s][0] = lo.x; bq[s][1] = lo.y; bq[s][2] = hi.x; bq[s][3] = hi.y;
      }
      SegDesc sd = sdesc[0];
      K2_TRACE(0);
      mbar_wait(&full[b], (uint32_t)((k / NBUF) & 1));
      K2_TRACE(1);
      const double* xrow = Xs0 + (size_t)b * xsz + qrow;
      const double* ph = phis0 + (size_t)b * NC * cstride + qrow;
      double piacc[NC][MT];
#pragma unroll
      for (int d = 0; d < NC; ++d)
#pragma unroll
        for (int m = 0; m < MT; ++m) piacc[d][m] = 0.0;
      for (int sg = 0; sg < nseg; ++sg) {
        const SegDesc sd_next = sdesc[sg + 1];
        switch (sd.ntl) {
          case 4: run_segment<MT, PS, 4, GRAD>(sd, piacc, bq, xrow, qcol, bptr, ph, pofs, cstride); break;
          case 3: run_segment<MT, PS, 3, GRAD>(sd, piacc, bq, xrow, qcol, bptr, ph, pofs, cstride); break;
          case 2: run_segment<MT, PS, 2, GRAD>(sd, piacc, bq, xrow, qcol, bptr, ph, pofs, cstride); break;
          default: run_segment<MT, PS, 1, GRAD>(sd, piacc, bq, xrow, qcol, bptr, ph, pofs, cstride); break;
        }
        sd = sd_next;
      }
      K2_TRACE(2);
      double* red = red0 + ((size_t)b * WARPS + tl) * NC * P;
#pragma unroll
      for (int d = 0; d < NC; ++d)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          double v = piacc[d][m];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (qcol == 0) red[d * P + m * 8 + qrow] = v;
        }
      __syncwarp();
      if (lane == 0) mbar_arrive(&done[b]);
      K2_TRACE(3);
    }
  }
}

// LEAN = true trims the register footprint of the CTA to 384 x 128 = 48 K registers (launch cap 128
// per thread via the 512-thread launch bound; builders 56, MMA warps 160) so that one 128-thread
// CTA of another kernel can stay co-resident on the SM.
template <bool GRAD, int NBUF, bool LEAN = false>
__global__ void __launch_bounds__(LEAN ? 512 : WS_THREADS, 1) k2_ontop_ws(const Params prm) {
  ws_cta<GRAD, NBUF, LEAN ? 160 : 224>(prm);
}

}  // namespace k2
