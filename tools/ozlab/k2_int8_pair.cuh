// k2_int8_pair.cuh -- LAB (tools/ozlab; not linked into the product): the CTA-pair form of the int8 tensor-core K2
// (openrdm_b200/csrc/k2_int8.cuh; same arithmetic, same operands):
// two SMs of a cluster run tcgen05.mma.cta_group::2 (M = 256: 128 points per CTA), each CTA holds only HALF of the rows
// of every T chunk (84 KB instead of 168 KB).  The shared memory that frees takes X slices 0..4 (140 KB) out of TMEM,
// and TMEM then has room for TWO full-width accumulators (2 x 224 columns) + X slice 5:
//   * one MMA per (slice pair, K chunk) over the chunk's full column range (7 per pair instead of 11 half-width ones);
//   * consecutive groups alternate between the two accumulators: the epilogue of group G overlaps the MMAs of group
//     G - 1 completely, and group G - 2 waits only for G's drain.
// Only the leader CTA's MMA lane issues; tcgen05.commit multicasts "accumulator full" to both CTAs, both CTAs' epilogue
// threads arrive on the leader's "drained" barrier (mbarrier.arrive.shared::cluster), and one cluster barrier per tile
// says "both CTAs' operands are built".
#pragma once
#include <vector>
#include "../../openrdm_b200/csrc/k2_int8.cuh"
#include "../../openrdm_b200/csrc/rdm_pack.h"

namespace ordm {
// The same image for the CTA-pair kernel (below): block c (c = 0, 1) is what CTA rank c of a cluster keeps
// in shared memory -- of every K chunk only rows [c rows/2, (c+1) rows/2) (tcgen05.mma.cta_group::2 takes half of B's
// N from each CTA), each half chunk a K-major no-swizzle operand of rows/2 rows.  Slice stride and chunk offsets
// are half of pack_t_int8's.
inline void pack_t_int8_pair(int M, const std::vector<double>& Dt, int kp, std::vector<uint8_t>& img, int& t_exp) {
  std::vector<uint8_t> whole;
  pack_t_int8(M, Dt, kp, whole, t_exp);
  const int nslice = 6;
  const size_t slice = (size_t)t_int8_slice_bytes(kp), half = slice / 2;
  img.assign((size_t)2 * nslice * half, 0);
  for (int l = 0; l < nslice; ++l)
    for (int kc = 0; kc < kp / 32; ++kc) {
      const int rows = kp - 32 * kc, hr = rows / 2;
      for (int n = 0; n < rows; ++n)
        for (int k = 0; k < 32; ++k) {
          const size_t from = (size_t)l * slice + (size_t)t_int8_chunk_off(kp, kc) + (size_t)((k >> 4) * (rows >> 3) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 15);
          const int c = n / hr, nl = n % hr;
          const size_t to = ((size_t)c * nslice + l) * half + (size_t)t_int8_chunk_off(kp, kc) / 2 + (size_t)((k >> 4) * (hr >> 3) + (nl >> 3)) * 128 + (nl & 7) * 16 + (k & 15);
          img[to] = whole[from];
        }
    }
}

}  // namespace ordm

namespace k2i {

// shared-memory plan (dynamic) of one CTA of the pair: T half image | X slices 0..4 (the partial sums [2][128] of the
// tile's end alias the first 2 KB of the X slices, which are dead by then)
struct PairSmem {
  __host__ __device__ static size_t t_bytes(int kp) { return (size_t)NSLICE * Geom::slice_bytes(kp) / 2; }
  __host__ __device__ static size_t x_off(int kp) { return t_bytes(kp); }
  __host__ __device__ static size_t part_off(int kp) { return x_off(kp); }
  __host__ __device__ static size_t bytes(int kp) { return x_off(kp) + (size_t)5 * TILE * kp; }
};

__device__ __forceinline__ void mma2_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar) {   // arrives on the barrier at this offset in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// slices 0..4 -> shared memory (K-major no-swizzle operand of 128 rows: 16 K bytes of a row are contiguous, so four
// slice words go out as one 16-byte store, conflict-free across the lanes of a warp), slice 5 -> TMEM
template <int NA, int R>
__device__ __forceinline__ void build_half_pair(const double (&phi)[NA], double scale, uint8_t* x_row, uint32_t tmem_x5_row) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, EPT = KP / 2;
  uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, w[6][4];
  auto emit = [&](int j) {   // slice words of elements 4j .. 4j+3 of this half
    const uint32_t t0 = __byte_perm(lo[0], lo[1], 0x5140), t1 = __byte_perm(lo[2], lo[3], 0x5140);
    const uint32_t t2 = __byte_perm(lo[0], lo[1], 0x7362), t3 = __byte_perm(lo[2], lo[3], 0x7362);
    const uint32_t h0 = __byte_perm(hi[0], hi[1], 0x5140), h1 = __byte_perm(hi[2], hi[3], 0x5140);
    w[0][j & 3] = __byte_perm(t0, t1, 0x5410) ^ 0x80808080u;
    w[1][j & 3] = __byte_perm(t0, t1, 0x7632) ^ 0x80808080u;
    w[2][j & 3] = __byte_perm(t2, t3, 0x5410) ^ 0x80808080u;
    w[3][j & 3] = __byte_perm(t2, t3, 0x7632) ^ 0x80808080u;
    w[4][j & 3] = __byte_perm(h0, h1, 0x5410) ^ 0x80808080u;
    w[5][j & 3] = __byte_perm(h0, h1, 0x7632);
    if ((j & 3) == 3) {
      const int k = EPT * R + 4 * (j - 3);   // first K byte of the four words: a multiple of 16
#pragma unroll
      for (int s = 0; s < 5; ++s)
        *(uint4*)(x_row + (size_t)s * TILE * KP + (k >> 4) * (TILE / 8) * 128) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tmem_x5_row + (uint32_t)((EPT / 4) * R + j - 3)),
                   "r"(w[5][0]), "r"(w[5][1]), "r"(w[5][2]), "r"(w[5][3]) : "memory");
    }
  };
  for_my_pairs<NA, R>([&](int i, int t, int u) {
    const unsigned long long v = (unsigned long long)__double_as_longlong(fma(phi[t] * scale, phi[u], 6755399441055744.0)) + 0x0000008080808080ull;
    lo[i & 3] = (uint32_t)v; hi[i & 3] = (uint32_t)(v >> 32);
    if ((i & 3) == 3) emit(i >> 2);
  });
  constexpr int last = M - EPT * R;
  if constexpr (last < EPT) {
#pragma unroll
    for (int i = last; i < EPT; ++i) {
      lo[i & 3] = 0x80808080u; hi[i & 3] = 0x00000080u;
      if ((i & 3) == 3) emit(i >> 2);
    }
  }
}

// ---- epilogue of one group, TMEM-latency version.  tcgen05.wait::ld waits for every outstanding load, so the depth of
// the load pipeline is one batch; beside continuously running MMAs a TMEM round trip costs ~400 clk, and seven small
// batches (fold_group) made the epilogue the critical path.  Here the half row is read in four batches of Q / 4 columns
// from each of its four regions (28 or 24 columns per batch), the next batch in flight while one is folded.
template <int CPR>
__device__ __forceinline__ void tmem_ld_region(uint32_t taddr, uint32_t* v) {   // CPR = 6 or 7 consecutive columns
  tmem_ld4(taddr, v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[4]), "=r"(v[5]) : "r"(taddr + 4));
  if constexpr (CPR == 7) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[6]) : "r"(taddr + 6));
}

template <int NA, int R, int ROUND, int CPR, int J>   // J = 4 e + k: column e of the batch in region k
__device__ __forceinline__ void fold_round_elem(const double (&phi)[NA], double (&s)[NA], const uint32_t (&v)[4 * CPR]) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, EPT = KP / 2, Q = EPT / 4;
  constexpr int e = J / 4, k = J % 4, I = EPT * R + Q * k + CPR * ROUND + e;
  if constexpr (I < M) {
    constexpr int t = pair_t_of(NA, I), u = pair_u_of(NA, I);
    const double zd = __hiloint2double(0x43300000, (int)(v[CPR * k + e] ^ 0x80000000u)) - 4503601774854144.0;  // 2^52 + 2^31
    s[t] = fma(phi[u], zd, s[t]);
  }
  if constexpr (J + 1 < 4 * CPR) fold_round_elem<NA, R, ROUND, CPR, J + 1>(phi, s, v);
}

// consecutive DFMAs go to different regions = (mostly) different orbitals' chains
template <int NA, int R, int ROUND, int CPR>
__device__ __forceinline__ void fold_round(const double (&phi)[NA], double (&s)[NA], const uint32_t (&v)[4 * CPR]) {
  fold_round_elem<NA, R, ROUND, CPR, 0>(phi, s, v);
}

template <int NA, int R, int ROUND, int CPR>
__device__ __forceinline__ void fold_rounds(const double (&phi)[NA], double (&s)[NA], uint32_t tmem_half, uint32_t (&cur)[4 * CPR], uint32_t (&nxt)[4 * CPR],
                                            uint32_t drain_bar) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, EPT = KP / 2, Q = EPT / 4;
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if constexpr (ROUND + 1 < 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) tmem_ld_region<CPR>(tmem_half + (uint32_t)(Q * k + CPR * (ROUND + 1)), &nxt[CPR * k]);
  } else {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(drain_bar) : "memory");
  }
#ifndef K2I_LAB_NO_FOLD_MATH   // lab switch: loads and barriers only
  fold_round<NA, R, ROUND, CPR>(phi, s, cur);
#else
  if (cur[0] == 0x12345678u) s[0] += 1.0;
#endif
  if constexpr (ROUND + 1 < 4) fold_rounds<NA, R, ROUND + 1, CPR>(phi, s, tmem_half, nxt, cur, drain_bar);
}

template <int NA, int R>
__device__ __forceinline__ double fold_group_rounds(const double (&phi)[NA], uint32_t tmem_half, uint32_t drain_bar) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, EPT = KP / 2, Q = EPT / 4, CPR = Q / 4;
  static_assert(Q % 4 == 0 && (CPR == 6 || CPR == 7), "four regions, four batches of 6 or 7 columns each");
  double s[NA];
#pragma unroll
  for (int t = 0; t < NA; ++t) s[t] = 0.0;
  uint32_t va[4 * CPR], vb[4 * CPR];
#pragma unroll
  for (int k = 0; k < 4; ++k) tmem_ld_region<CPR>(tmem_half + (uint32_t)(Q * k), &va[CPR * k]);
  fold_rounds<NA, R, 0, CPR>(phi, s, tmem_half, va, vb, drain_bar);
  double part0 = 0.0, part1 = 0.0;
#pragma unroll
  for (int t = 0; t < NA; t += 2) {
    part0 = fma(phi[t], s[t], part0);
    if (t + 1 < NA) part1 = fma(phi[t + 1], s[t + 1], part1);
  }
  return part0 + part1;
}

// every MMA of group G into the accumulator at tmem_acc: one per (slice pair, K chunk), N = the chunk's column range
template <int G, int KP>
__device__ __forceinline__ void issue_group_pair(uint32_t tmem_acc, uint32_t tmem_x5, uint64_t tdesc0, uint64_t xdesc0, uint32_t idesc0) {
  constexpr int NCH = KP / 32, KMAX = G < NSLICE - 1 ? G : NSLICE - 1, KMIN = G > NSLICE - 1 ? G - (NSLICE - 1) : 0;
#pragma unroll
  for (int k = KMAX; k >= KMIN; --k) {
    const int l = G - k;
#pragma unroll
    for (int kc = 0; kc < NCH; ++kc) {
      const int n = Geom::chunk_rows(KP, kc);   // columns [32 kc, KP); each CTA holds n / 2 rows of the chunk
      const uint32_t idesc = idesc0 | ((uint32_t)(n >> 3) << 17);
      const uint64_t bd = tdesc0 + (uint64_t)((l * (Geom::slice_bytes(KP) / 2) + Geom::chunk_off(KP, kc) / 2) >> 4) + ((uint64_t)((n / 2) * 16 >> 4) << 16);
      const uint32_t first = (k == KMAX && kc == 0) ? 0u : 1u;
      if (k == 5) mma2_ts(tmem_acc + 32 * kc, tmem_x5 + 8 * kc, bd, idesc, first);
      else mma2_ss(tmem_acc + 32 * kc, xdesc0 + (uint64_t)((k * TILE * KP + kc * 2 * (TILE / 8) * 128) >> 4), bd, idesc, first);
    }
  }
}

template <int KP>
__device__ __forceinline__ void issue_group_pair_of(int G, uint32_t tmem_acc, uint32_t tmem_x5, uint64_t tdesc0, uint64_t xdesc0, uint32_t idesc0) {
  switch (G) {
    case 10: issue_group_pair<10, KP>(tmem_acc, tmem_x5, tdesc0, xdesc0, idesc0); break;
    case 9: issue_group_pair<9, KP>(tmem_acc, tmem_x5, tdesc0, xdesc0, idesc0); break;
    case 8: issue_group_pair<8, KP>(tmem_acc, tmem_x5, tdesc0, xdesc0, idesc0); break;
    case 7: issue_group_pair<7, KP>(tmem_acc, tmem_x5, tdesc0, xdesc0, idesc0); break;
    case 6: issue_group_pair<6, KP>(tmem_acc, tmem_x5, tdesc0, xdesc0, idesc0); break;
    case 5: issue_group_pair<5, KP>(tmem_acc, tmem_x5, tdesc0, xdesc0, idesc0); break;
    default: issue_group_pair<4, KP>(tmem_acc, tmem_x5, tdesc0, xdesc0, idesc0); break;
  }
}

// Measured (B200, 2 M points, 26 slice pairs): 1.87-1.90 ms against 2.00 ms for the single-CTA kernel.  The MMAs
// take 15 k clk per tile (21 k before) and run almost back to back, but the epilogue is now the critical path: 3.1 k clk
// per group and thread (1.1 k of TMEM round trips + 2.0 k for 2 x 112 dependent-latency-bound FP64 instructions with two
// warps per SM sub-partition), 22 k per tile.  See DESIGN.md section 7.
// P.timg here is the PAIR image (ordm::pack_t_int8_pair): CTA rank c of the cluster reads block c
template <int NA, int GMIN>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(K2I_MAXREG) k2i_ontop_pair(Params P) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32;
  constexpr int EPT = KP / 2;
  constexpr int ACC_COLS = KP, X5_COL = 2 * KP;   // TMEM: accumulator b at [b KP, (b + 1) KP), X slice 5 at [2 KP, 2 KP + KP / 4)
  static_assert(2 * KP + KP / 4 <= 512, "TMEM columns");
  static_assert(NSLICE == 6 && (GMIN == 4 || GMIN == 5), "six slices, groups 10 .. GMIN");
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_s[4];   // [0..1] accumulator b complete (multicast commit); [2..3] (leader's copy) accumulator b drained by both CTAs
  uint8_t* s_t = smem;
  uint8_t* s_x = smem + PairSmem::x_off(KP);
  double* s_part = (double*)(smem + PairSmem::part_off(KP));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = warp < CWARPS;
  const int q = warp & 3, r = (warp >> 2) & 1, p = 32 * q + lane;  // lane quadrant, K half, point within the CTA's tile
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));

  {
    const int4* src = (const int4*)(P.timg + (size_t)crank * PairSmem::t_bytes(KP));
    for (int i = tid; i < (int)(PairSmem::t_bytes(KP) / 16); i += THREADS) ((int4*)s_t)[i] = src[i];
  }
  const uint32_t bar_full = (uint32_t)__cvta_generic_to_shared(&bar_s[0]), bar_drain = (uint32_t)__cvta_generic_to_shared(&bar_s[2]);
  if (tid == 0) {
    for (int h = 0; h < 2; ++h) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_full + 8 * h));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_drain + 8 * h), "r"(2 * CWARPS * 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * TILE) >> 4) << 24);   // M = 256 over the pair
  uint8_t* x_row = s_x + (p >> 3) * 128 + (p & 7) * 16;
  const uint64_t tdesc0 = smem_desc((uint32_t)__cvta_generic_to_shared(s_t), 0, 128);
  const uint64_t xdesc0 = smem_desc((uint32_t)__cvta_generic_to_shared(s_x), (TILE / 8) * 128, 128);
  // the leader's "drained" barriers, as seen from this CTA
  uint32_t drain_remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(drain_remote) : "r"(bar_drain), "r"(0));
#ifdef K2I_PROF
  long long pf_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc_ = clock64();
#endif
  uint32_t n = 0;   // groups issued / consumed so far (both accumulators see one group every other n)
  bool ok = true;

  const size_t ntiles = (P.npts + 2 * TILE - 1) / (2 * TILE);   // pair tiles of 256 points
  for (size_t tile = blockIdx.x / 2; tile < ntiles; tile += gridDim.x / 2) {
    const size_t p0 = tile * 2 * TILE + (size_t)crank * TILE;
    double phi[NA];
    int e_raw = 534;
    K2I_LAP(7);
    if (compute) {
      uint32_t mh = 0;
#pragma unroll
      for (int t = 0; t < NA; ++t) {
        phi[t] = (p0 + p < P.npts) ? __ldcs(P.phi_act + (size_t)t * P.ld + p0 + p) : 0.0;
        mh = max(mh, (uint32_t)__double2hiint(phi[t]) & 0x7FFFFFFFu);
      }
      e_raw = min(max((int)(mh >> 20), 534), 1500);
      const double scale = __hiloint2double((3113 - 2 * e_raw + (FIX_BITS - 46)) << 20, 0);  // 2^(FIX_BITS - 2 em)
      const uint32_t tmem_x5_row = tmem + lane_addr + (uint32_t)X5_COL;
      if (r == 0) build_half_pair<NA, 0>(phi, scale, x_row, tmem_x5_row); else build_half_pair<NA, 1>(phi, scale, x_row, tmem_x5_row);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      // next pair tile's orbitals towards L2 while this one is contracted
      const size_t pn = p0 + (size_t)(gridDim.x / 2) * 2 * TILE + p;
      if (pn < P.npts) {
#pragma unroll
        for (int t = r; t < NA; t += 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.phi_act + (size_t)t * P.ld + pn));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    cluster_sync_all();   // both CTAs' operands are built (and the previous tile's partial sums have been consumed)
    K2I_LAP(1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    double pi_acc = 0.0;
#pragma unroll 1
    for (int G = 2 * (NSLICE - 1); G >= GMIN; --G, ++n) {
      const uint32_t b = n & 1, use = n >> 1;   // accumulator, and how often it has been used before
      if (!compute) {
        if (lane == 0 && crank == 0) {
          if (use > 0) ok = wait_parity(bar_drain + 8 * b, (use - 1) & 1) && ok;   // both CTAs have read this accumulator's previous group
          K2I_LAP(2);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_group_pair_of<KP>(G, tmem + b * ACC_COLS, tmem + X5_COL, tdesc0, xdesc0, idesc0);
          commit2(bar_full + 8 * b);
          K2I_LAP(3);
        }
        __syncwarp();
      } else {
        ok = wait_parity(bar_full + 8 * b, use & 1) && ok;
        K2I_LAP(5);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_half = tmem + lane_addr + (uint32_t)(b * ACC_COLS + EPT * r);
        const double part = r == 0 ? fold_group_rounds<NA, 0>(phi, tmem_half, drain_remote + 8 * b) : fold_group_rounds<NA, 1>(phi, tmem_half, drain_remote + 8 * b);
        pi_acc = fma(part, __hiloint2double((1023 + 8 * (G - GMIN)) << 20, 0), pi_acc);
        K2I_LAP(6);
      }
    }
    // every thread has seen the last group's "full": all MMAs of the tile are done, the X slices may be rebuilt
    if (compute) s_part[r * TILE + p] = pi_acc;
    __syncthreads();
    if (compute && r == 0 && p0 + p < P.npts) {
      const double s = s_part[p] + s_part[TILE + p];
      const double f1 = __hiloint2double((2 * e_raw - 1067 - (FIX_BITS - 46)) << 20, 0);        // 2^(2 em - FIX_BITS)
      const double f2 = __hiloint2double((1023 + P.t_exp - FIX_BITS + 8 * GMIN + 1) << 20, 0);  // 2 * 2^(t_exp - FIX_BITS + 8 GMIN)
      const double pi_act = s * f1 * f2;
      P.pi[p0 + p] = P.pi_core ? pi_act + P.pi_core[p0 + p] : pi_act;
    }
    __syncthreads();   // the partial sums alias the X slices: read before the next tile's build writes them
  }
#ifdef K2I_PROF
  if (blockIdx.x == 0 && (tid & 127) == 0) for (int i = 0; i < 8; ++i) k2i_prof[tid >> 7][i] = pf_[i];
#endif
  if (!ok && P.fail) atomicAdd(P.fail, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace k2i
