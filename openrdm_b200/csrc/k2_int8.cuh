// k2_int8.cuh -- K2 on the 5th-generation tensor cores: the active-space on-top pair density
//     Pi_act(p) = sum_{I,J} X_pI D~_IJ X_pJ,   X_pI = phi_t(p) phi_u(p)   (I = packed pair t <= u),
// i.e. MCPDFT::build_ontop_pair_density (/root/reference/src/libmcpdft/mcpdft.cc:185-212) restricted to the active
// orbitals, contracted by tcgen05.mma.kind::i8 through an Ozaki-style exact integer split:
//   * D~ = T + T^T with T block-upper-triangular (32 x 32 blocks, diagonal blocks halved), so Pi = 2 sum_J X_J (X T)_J;
//   * every X row (against (max_t |phi_t|)^2) and T (against max |T|) becomes a 46-bit fixed-point image cut into
//     six balanced signed bytes; the slice pairs (k, l) with k + l >= GMIN are multiplied exactly (INT32
//     accumulators in TMEM), one accumulator per group G = k + l, highest group first;
//   * X slices 0..4 live in TMEM (A operand from TMEM, written with tcgen05.st), slice 5 and all of T in shared
//     memory (K-major, no swizzle: 8 x 16-byte core matrices), the accumulator takes the remaining TMEM columns;
//   * the accumulator is two independent column halves (own MMA chains and barriers): while the epilogue dots one
//     half with the FP64 X row, straight from TMEM, the tensor cores fill the other.
// Error against the FP64 contraction: the dropped pairs (k + l < GMIN) and the 46-bit images, |dPi| ~ 2e-14 (GMIN = 4)
// or ~1e-13 (GMIN = 5) of max(1, |Pi|) on config-4-like data (tools/ozlab/oz_accuracy.py); the parity bar is 1e-12.
// One CTA per SM, persistent over 128-point tiles: 8 compute warps (4 TMEM lane quadrants x 2 K halves; a thread
// owns one point and half of its pairs, the pair -> (t, u) map resolved at compile time) + 1 warp issuing the MMAs.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace k2i {

constexpr int MIN_ACT = 18, MAX_ACT = 20;   // active-space sizes this path is built for (TMEM columns: KP + 5 KP/4 <= 512)

constexpr int TILE = 128;       // points per tile = MMA M
constexpr int NSLICE = 6;       // byte slices per operand
constexpr int CWARPS = 8;       // compute warps (build + epilogue); one more warp issues the MMAs
constexpr int THREADS = (CWARPS + 1) * 32;
#ifndef K2I_MAXREG
#define K2I_MAXREG 128   // 9 warps are allocated as 12: at 128 registers the CTA leaves 16 K registers to a co-resident density CTA
#endif
constexpr int FIX_BITS = 46;    // |fixed-point image| <= 2^46 (one bit of headroom for the balancing carry)

struct Geom {                   // everything derived from the padded pair count KP (multiple of 32)
  __host__ __device__ static constexpr int chunk_rows(int kp, int kc) { return kp - 32 * kc; }                 // N of K-chunk kc
  __host__ __device__ static constexpr int chunk_off(int kp, int kc) { return 32 * (kc * kp - 16 * kc * (kc - 1)); }  // bytes before chunk kc
  __host__ __device__ static constexpr int slice_bytes(int kp) { return chunk_off(kp, kp / 32); }
};

struct Params {
  const double* phi_act;   // [nact][ld]: orbital t of point p at phi_act[t * ld + p]
  size_t ld, npts;
  const uint8_t* timg;     // NSLICE x slice_bytes(kp): balanced byte slices of T in shared-memory order (ordm::pack_t_int8)
  int t_exp;               // max |T| < 2^t_exp
  const double* pi_core;   // optional: closed-form core part, added here (nullptr: K3 adds it)
  double* pi;              // [npts] out
  int* fail;               // fail[0]: bumped if an MMA-completion barrier ever times out; fail[1]: points flagged by the error guard
  double guard_g1, guard_g2;   // constants of the a-posteriori error estimate (k2_int8_pair.cuh: guard_flag; ordm::int8_guard_constants)
};

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
#ifndef K2I_WAIT_NAP_NS
#define K2I_WAIT_NAP_NS 0
#endif
__device__ __forceinline__ bool wait_parity(uint32_t bar, uint32_t parity) {
  for (long spin = 0; spin < (1l << 24); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
#if K2I_WAIT_NAP_NS > 0
    __nanosleep(K2I_WAIT_NAP_NS);   // (lab switch: the step runs under the board's power cap; does a quieter wait buy clocks?)
#endif
  }
  return false;
}
// ---- a-posteriori error estimate of the fixed-point scheme for one point ("guard") -------------
// With q_x = 2^(2 em - 46) the unit of the point's X image, q_t = 2^(t_exp - 46) that of T, independent uniform rounding
// errors and the final dot taken with the unquantised X:
//   var(dPi) <= q_x^2/3 |T|_F^2 |X|^2  +  q_t^2/3 |X|^4  +  4 d^2 2^(4 em) M |X|^2,     |X|^2 <= (sum_t phi_t^2)^2,
// the last term being the dropped slice pairs (d = their r.m.s. per element of X T: 2^-54.6 (GMIN 4), 2^-46.4 (GMIN 5)
// of 2^(2 em + t_exp)).  The host folds the three constants and z^2 (z = 4 standard deviations) into g1, g2
// (ordm::int8_guard_constants):   z^2 var = 2^(4 em) g1 S2^2 + g2 S2^4,  S2 = sum_t phi_t^2.
// The point is flagged when that exceeds (1e-12 max(1, |Pi|))^2.  tests/test_int8_scheme.py pins the model against the
// exact-integer restatement of the arithmetic (measured |dPi| <= 2.5 standard deviations on every case tried).
template <int NA>
__device__ __forceinline__ double sum_squares(const double (&phi)[NA]) {
  double s2 = 0.0;
#pragma unroll
  for (int t = 0; t < NA; ++t) s2 = fma(phi[t], phi[t], s2);
  return s2;
}
__device__ __forceinline__ bool guard_flag(double s2, int e_raw, double pi_abs, double g1, double g2) {
  const double xs2 = __hiloint2double((4 * e_raw - 3065) << 20, 0);   // 2^(4 em), em = e_raw - 1022
  const double x2 = s2 * s2;
  const double bar = 1e-12 * fmax(1.0, pi_abs);
  return fma(xs2 * g1, x2, g2 * x2 * x2) > bar * bar;
}

// shared-memory plan (dynamic): T image | X slice 5 (K-major no-swizzle operand, 128 rows x KP bytes) |
// partial sums [2][128] doubles | the next tile's orbitals [nact][128] doubles (cp.async)
struct Smem {
  __host__ __device__ static size_t t_bytes(int kp) { return (size_t)NSLICE * Geom::slice_bytes(kp); }
  __host__ __device__ static size_t a5_off(int kp) { return t_bytes(kp); }
  __host__ __device__ static size_t part_off(int kp) { return a5_off(kp) + (size_t)TILE * kp; }
  __host__ __device__ static size_t next_off(int kp) { return part_off(kp) + 2 * TILE * 8; }
  __host__ __device__ static size_t bytes(int kp, int nact) { return next_off(kp) + (size_t)nact * TILE * 8; }
};

// Thread layout: 8 compute warps = 4 lane quadrants x 2 K halves: a thread owns one point of the tile and EPT = KP/2
// consecutive pairs, keeps the point's NA orbital values in registers (the pair -> (t, u) map is resolved at compile
// time by full unrolling), builds their byte slices and later folds its half of the accumulator row.
// Warp 8 (one lane) only issues the MMAs and waits for them.
template <int NA, int R, class F>
__device__ __forceinline__ void for_my_pairs(F&& f) {   // f(I, t, u) for the pairs I of K half R, in order
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, EPT = KP / 2;
  int I = 0;
#pragma unroll
  for (int t = 0; t < NA; ++t)
#pragma unroll
    for (int u = t; u < NA; ++u) {
      if (I >= EPT * R && I < EPT * (R + 1)) f(I - EPT * R, t, u);
      ++I;
    }
}

template <int NA, int R>
__device__ __forceinline__ void build_half(const double (&phi)[NA], double scale, uint32_t tmem_row, uint8_t* a5_row) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, EPT = KP / 2, SL_COLS = KP / 4;
  uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, w[5][4];
  auto emit = [&](int j) {   // slice words of elements 4j .. 4j+3 of this half
    const uint32_t t0 = __byte_perm(lo[0], lo[1], 0x5140), t1 = __byte_perm(lo[2], lo[3], 0x5140);
    const uint32_t t2 = __byte_perm(lo[0], lo[1], 0x7362), t3 = __byte_perm(lo[2], lo[3], 0x7362);
    const uint32_t h0 = __byte_perm(hi[0], hi[1], 0x5140), h1 = __byte_perm(hi[2], hi[3], 0x5140);
    w[0][j & 3] = __byte_perm(t0, t1, 0x5410) ^ 0x80808080u;
    w[1][j & 3] = __byte_perm(t0, t1, 0x7632) ^ 0x80808080u;
    w[2][j & 3] = __byte_perm(t2, t3, 0x5410) ^ 0x80808080u;
    w[3][j & 3] = __byte_perm(t2, t3, 0x7632) ^ 0x80808080u;
    w[4][j & 3] = __byte_perm(h0, h1, 0x5410) ^ 0x80808080u;
    const int k = EPT * R + 4 * j;   // K byte of the word; canonical 128-row operand: ((k>>4)*16 + rb)*128 + (r&7)*16 + (k&15)
    *(uint32_t*)(a5_row + (k >> 4) * (TILE / 8) * 128 + (k & 15)) = __byte_perm(h0, h1, 0x7632);
    if ((j & 3) == 3) {
#pragma unroll
      for (int s = 0; s < 5; ++s)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tmem_row + (uint32_t)(SL_COLS * s + (EPT / 4) * R + j - 3)),
                     "r"(w[s][0]), "r"(w[s][1]), "r"(w[s][2]), "r"(w[s][3]) : "memory");
    }
  };
  for_my_pairs<NA, R>([&](int i, int t, int u) {
    const unsigned long long v = (unsigned long long)__double_as_longlong(fma(phi[t] * scale, phi[u], 6755399441055744.0)) + 0x0000008080808080ull;
    lo[i & 3] = (uint32_t)v; hi[i & 3] = (uint32_t)(v >> 32);
    if ((i & 3) == 3) emit(i >> 2);
  });
  // padding pairs of the last half: zero digits (the image of 0 is 0x..80 80 80 80 80 before re-centring)
  constexpr int last = M - EPT * R;   // number of real pairs in this half if it is the padded one
  if constexpr (last < EPT) {
#pragma unroll
    for (int i = last; i < EPT; ++i) {
      lo[i & 3] = 0x80808080u; hi[i & 3] = 0x00000080u;
      if ((i & 3) == 3) emit(i >> 2);
    }
  }
}

// ---- epilogue of one group: sum_J X_J acc_J over this thread's half row, straight from TMEM.
// Factored as sum_t phi_t (sum_u phi_u acc_tu): one exact int -> double (bit trick + DADD) and one DFMA per pair.
// The half row is read as four regions of EPT/4 columns, four columns of each per step, so that consecutive DFMAs
// feed different orbitals' chains (the FP64 pipe needs ~4 independent chains per warp), and the loads of step s + 1
// are in flight while step s is folded.
__host__ __device__ constexpr int pair_t_of(int na, int I) { int t = 0; while (I >= na - t) { I -= na - t; ++t; } return t; }
__host__ __device__ constexpr int pair_u_of(int na, int I) { int t = 0; while (I >= na - t) { I -= na - t; ++t; } return t + I; }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}

template <int NA, int R, int STEP, int J>
__device__ __forceinline__ void fold_elem(const double (&phi)[NA], double (&s)[NA], const uint32_t (&v)[16]) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, EPT = KP / 2, Q = EPT / 4;
  constexpr int I = EPT * R + Q * (J / 4) + 4 * STEP + (J % 4);   // region J / 4, column 4 STEP + J % 4 of it
  if constexpr (I < M) {
    constexpr int t = pair_t_of(NA, I), u = pair_u_of(NA, I);
    const double zd = __hiloint2double(0x43300000, (int)(v[J] ^ 0x80000000u)) - 4503601774854144.0;  // 2^52 + 2^31
    s[t] = fma(phi[u], zd, s[t]);
  }
}

template <int NA, int R, int STEP>
__device__ __forceinline__ void fold_steps(const double (&phi)[NA], double (&s)[NA], uint32_t tmem_half, uint32_t (&cur)[16], uint32_t (&nxt)[16],
                                           uint32_t drain_bar) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, EPT = KP / 2, Q = EPT / 4, NSTEP = Q / 4;
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");   // step STEP is in `cur`
  if constexpr (STEP + 1 < NSTEP) {
#pragma unroll
    for (int k = 0; k < 4; ++k) tmem_ld4(tmem_half + (uint32_t)(Q * k + 4 * (STEP + 1)), &nxt[4 * k]);
  } else {   // every column of this half has been read: the next group's MMAs may overwrite it
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(drain_bar) : "memory");
  }
  fold_elem<NA, R, STEP, 0>(phi, s, cur); fold_elem<NA, R, STEP, 4>(phi, s, cur); fold_elem<NA, R, STEP, 8>(phi, s, cur); fold_elem<NA, R, STEP, 12>(phi, s, cur);
  fold_elem<NA, R, STEP, 1>(phi, s, cur); fold_elem<NA, R, STEP, 5>(phi, s, cur); fold_elem<NA, R, STEP, 9>(phi, s, cur); fold_elem<NA, R, STEP, 13>(phi, s, cur);
  fold_elem<NA, R, STEP, 2>(phi, s, cur); fold_elem<NA, R, STEP, 6>(phi, s, cur); fold_elem<NA, R, STEP, 10>(phi, s, cur); fold_elem<NA, R, STEP, 14>(phi, s, cur);
  fold_elem<NA, R, STEP, 3>(phi, s, cur); fold_elem<NA, R, STEP, 7>(phi, s, cur); fold_elem<NA, R, STEP, 11>(phi, s, cur); fold_elem<NA, R, STEP, 15>(phi, s, cur);
  if constexpr (STEP + 1 < NSTEP) fold_steps<NA, R, STEP + 1>(phi, s, tmem_half, nxt, cur, drain_bar);
}

// tmem_half: this thread's lane, first column of its half of the accumulator
template <int NA, int R>
__device__ __forceinline__ double fold_group(const double (&phi)[NA], uint32_t tmem_half, uint32_t drain_bar) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, EPT = KP / 2, Q = EPT / 4;
  static_assert(Q % 4 == 0, "four regions of whole 4-column steps");
  double s[NA];
#pragma unroll
  for (int t = 0; t < NA; ++t) s[t] = 0.0;
  uint32_t va[16], vb[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) tmem_ld4(tmem_half + (uint32_t)(Q * k), &va[4 * k]);
  fold_steps<NA, R, 0>(phi, s, tmem_half, va, vb, drain_bar);
  double part0 = 0.0, part1 = 0.0;
#pragma unroll
  for (int t = 0; t < NA; t += 2) {
    part0 = fma(phi[t], s[t], part0);
    if (t + 1 < NA) part1 = fma(phi[t + 1], s[t + 1], part1);
  }
  return part0 + part1;
}

// The MMAs of group G = k + l (X slice k, T slice l) that feed one half of the accumulator columns -- H = 1: columns
// [KP/2, KP), H = 0: columns [0, KP/2).  The two halves are separate MMA chains with their own completion barriers, so
// one is drained by the epilogue while the tensor cores run the other.  K chunk kc of T only reaches columns >= 32 kc
// (block-upper-triangular).  Every descriptor is a compile-time offset from a base: a few instructions per MMA.
template <int G, int KP, int H>
__device__ __forceinline__ void issue_half(uint32_t tmem, uint64_t bdesc0, uint64_t adesc5, uint32_t idesc0) {
  constexpr int NCH = KP / 32, SL_COLS = KP / 4, A_COL = KP, HALF = KP / 2;
  constexpr int KMAX = G < NSLICE - 1 ? G : NSLICE - 1, KMIN = G > NSLICE - 1 ? G - (NSLICE - 1) : 0;
#pragma unroll
  for (int k = KMAX; k >= KMIN; --k) {
    const int l = G - k;
#pragma unroll
    for (int kc = 0; kc < NCH; ++kc) {
      if (H == 0 && 32 * kc >= HALF) continue;
      const int c0 = H ? (32 * kc > HALF ? 32 * kc : HALF) : 32 * kc;   // first accumulator column written
      const int n = (H ? KP : HALF) - c0, n0 = c0 - 32 * kc, rows = Geom::chunk_rows(KP, kc);
      const uint32_t idesc = idesc0 | ((uint32_t)(n >> 3) << 17);
      const uint64_t bd = bdesc0 + (uint64_t)((l * Geom::slice_bytes(KP) + Geom::chunk_off(KP, kc) + (n0 >> 3) * 128) >> 4) + ((uint64_t)(rows * 16 >> 4) << 16);
      const uint32_t first = (k == KMAX && kc == 0) ? 0u : 1u;
      if (k < 5) mma_ts(tmem + c0, tmem + (uint32_t)(A_COL + SL_COLS * k + 8 * kc), bd, idesc, first);
      else mma_ss(tmem + c0, adesc5 + (uint64_t)((kc * 2 * (TILE / 8) * 128) >> 4), bd, idesc, first);
    }
  }
}

template <int KP, int H>
__device__ __forceinline__ void issue_half_of(int G, uint32_t tmem, uint64_t bdesc0, uint64_t adesc5, uint32_t idesc0) {
  switch (G) {
    case 10: issue_half<10, KP, H>(tmem, bdesc0, adesc5, idesc0); break;
    case 9: issue_half<9, KP, H>(tmem, bdesc0, adesc5, idesc0); break;
    case 8: issue_half<8, KP, H>(tmem, bdesc0, adesc5, idesc0); break;
    case 7: issue_half<7, KP, H>(tmem, bdesc0, adesc5, idesc0); break;
    case 6: issue_half<6, KP, H>(tmem, bdesc0, adesc5, idesc0); break;
    case 5: issue_half<5, KP, H>(tmem, bdesc0, adesc5, idesc0); break;
    default: issue_half<4, KP, H>(tmem, bdesc0, adesc5, idesc0); break;
  }
}

#ifdef K2I_PROF   // lab only (tools/ozlab): clock64 laps of CTA 0's left-half thread 0, right-half thread 128 and the MMA lane
__device__ long long k2i_prof[3][8];
__device__ long long k2i_trace[3][40];   // clock64 stamps of CTA 0's 40th tile: [0] left-half thread, [1] right-half thread, [2] MMA lane
#define K2I_STAMP(i) do { if (blockIdx.x == 0 && (tid & 127) == 0 && ntile_ == 40) k2i_trace[tid >> 7][i] = clock64(); } while (0)
#define K2I_LAP(i) do { if (blockIdx.x == 0 && (tid & 127) == 0) { const long long t_ = clock64(); pf_[i] += t_ - tc_; tc_ = t_; } } while (0)
#else
#define K2I_LAP(i) do { } while (0)
#define K2I_STAMP(i) do { } while (0)
#endif

// GMIN: the slice pairs with k + l >= GMIN are kept (4: 26 pairs in 7 groups; 5: 21 pairs in 6 groups)
template <int NA, int GMIN>
__global__ void __maxnreg__(K2I_MAXREG) k2i_ontop(Params P) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32;
  constexpr int EPT = KP / 2;
  static_assert(EPT % 16 == 0, "geometry");
  constexpr int A_COL = KP;           // TMEM: accumulator columns [0, KP), X slice k < 5 at A_COL + (KP/4) k
  constexpr int SL_COLS = KP / 4;
  static_assert(A_COL + 5 * SL_COLS <= 512, "TMEM columns");
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_s[4];   // [0..1] accumulator half H complete (tcgen05.commit); [2..3] half H drained by its 128 threads
  uint8_t* s_t = smem;
  uint8_t* s_a5 = smem + Smem::a5_off(KP);
  double* s_part = (double*)(smem + Smem::part_off(KP));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = warp < CWARPS;
  const int q = warp & 3, r = (warp >> 2) & 1, p = 32 * q + lane;  // lane quadrant, K half, point within the tile

  for (int i = tid; i < (int)(Smem::t_bytes(KP) / 16); i += THREADS) ((int4*)s_t)[i] = ((const int4*)P.timg)[i];
  const uint32_t bar_full = (uint32_t)__cvta_generic_to_shared(&bar_s[0]), bar_drain = (uint32_t)__cvta_generic_to_shared(&bar_s[2]);
  if (tid == 0) {
    for (int h = 0; h < 2; ++h) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_full + 8 * h));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_drain + 8 * h), "r"(CWARPS * 32 / 2));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const uint32_t st_addr = (uint32_t)__cvta_generic_to_shared(s_t), a5_addr = (uint32_t)__cvta_generic_to_shared(s_a5);
  const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE >> 4) << 24);
  uint8_t* a5_row = s_a5 + (p >> 3) * 128 + (p & 7) * 16;
  const uint64_t bdesc0 = smem_desc(st_addr, 0, 128), adesc5 = smem_desc(a5_addr, (TILE / 8) * 128, 128);
  static_assert(NSLICE == 6 && GMIN >= 4 && GMIN <= 10, "the group dispatch below is written for six slices, groups 10 .. 4");
#ifdef K2I_PROF
  long long pf_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc_ = clock64();
  int ntile_ = -1;
#endif
  uint32_t phase = 0;      // parity of the current group's barriers (one completion of each per group)
  bool first_group = true, ok = true;
  int nflag = 0;   // points of this thread whose error estimate exceeds the parity bar (guard_flag)

  const size_t ntiles = (P.npts + TILE - 1) / TILE;
  double* s_nxt = (double*)(smem + Smem::next_off(KP));   // [NA][TILE]: the next tile's orbitals (cp.async)
  auto prefetch = [&](size_t tile) {   // K half r fetches the orbitals t = r, r + 2, ... of point p
    const size_t pp = tile * TILE + p;
#pragma unroll
    for (int t = r; t < NA; t += 2) {
      if (tile < ntiles && pp < P.npts)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(s_nxt + t * TILE + p)), "l"(P.phi_act + (size_t)t * P.ld + pp) : "memory");
      else
        s_nxt[t * TILE + p] = 0.0;
    }
  };
  if (compute) prefetch(blockIdx.x);
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t p0 = tile * TILE;
    double phi[NA];
    int e_raw = 534;
#ifdef K2I_PROF
    ++ntile_;
    int gi_ = 0;
#endif
    K2I_STAMP(0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();   // S1: this tile's orbitals are in shared memory
    K2I_LAP(0);
    if (compute) {
      uint32_t mh = 0;   // max over t of the high word of |phi_t|: its exponent field is that of max |phi_t| (integer max: no FP64 chain)
#pragma unroll
      for (int t = 0; t < NA; ++t) { phi[t] = s_nxt[t * TILE + p]; mh = max(mh, (uint32_t)__double2hiint(phi[t]) & 0x7FFFFFFFu); }
      // row scale: max |X| = (max_t |phi_t|)^2 < 2^(2 em); the clamps keep the two scale factors finite normal numbers
      e_raw = min(max((int)(mh >> 20), 534), 1500);
      const double scale = __hiloint2double((3113 - 2 * e_raw + (FIX_BITS - 46)) << 20, 0);  // 2^(FIX_BITS - 2 em)
      const uint32_t tmem_row = tmem + lane_addr + (uint32_t)A_COL;
      if (r == 0) build_half<NA, 0>(phi, scale, tmem_row, a5_row); else build_half<NA, 1>(phi, scale, tmem_row, a5_row);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // S2: operands ready (and every thread has read its orbitals)
    K2I_LAP(1);
    K2I_STAMP(1);
    if (compute) prefetch(tile + gridDim.x);
    double pi_acc = 0.0;
#pragma unroll 1
    for (int G = 2 * (NSLICE - 1); G >= GMIN; --G) {
      if (!compute) {
        if (lane == 0) {
#pragma unroll
          for (int h = 1; h >= 0; --h) {   // right half first: its drain overlaps the left half's MMAs
            if (!first_group) ok = wait_parity(bar_drain + 8 * h, phase ^ 1) && ok;   // the previous group has left this half
            K2I_LAP(2);
            K2I_STAMP(2 + 4 * gi_ + 2 * (1 - h));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (h) issue_half_of<KP, 1>(G, tmem, bdesc0, adesc5, idesc0); else issue_half_of<KP, 0>(G, tmem, bdesc0, adesc5, idesc0);
            commit(bar_full + 8 * h);
            K2I_LAP(3);
            K2I_STAMP(3 + 4 * gi_ + 2 * (1 - h));
          }
        }
        __syncwarp();
      } else {
        ok = wait_parity(bar_full + 8 * r, phase) && ok;
        K2I_LAP(5);
        K2I_STAMP(2 + 2 * gi_);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_half = tmem + lane_addr + (uint32_t)(EPT * r);
        const double part = r == 0 ? fold_group<NA, 0>(phi, tmem_half, bar_drain) : fold_group<NA, 1>(phi, tmem_half, bar_drain + 8);
        pi_acc = fma(part, __hiloint2double((1023 + 8 * (G - GMIN)) << 20, 0), pi_acc);
        K2I_LAP(6);
        K2I_STAMP(3 + 2 * gi_);
        if (G == GMIN) ok = wait_parity(bar_full + 8 * (r ^ 1), phase) && ok;   // every MMA of the tile is done: the X slices may be rebuilt
        K2I_LAP(5);
      }
      phase ^= 1;
      first_group = false;
#ifdef K2I_PROF
      ++gi_;
#endif
    }
    if (compute) s_part[r * TILE + p] = pi_acc;
    K2I_LAP(4);
    __syncthreads();   // S5
    K2I_LAP(7);
    K2I_STAMP(38);
    // ---- Pi = 2 * 2^(2 em - FIX) * 2^(t_exp - FIX) * 2^(8 GMIN) * (sum over the two K halves)
    if (compute && r == 0 && p0 + p < P.npts) {
      const double s = s_part[p] + s_part[TILE + p];
      const double f1 = __hiloint2double((2 * e_raw - 1067 - (FIX_BITS - 46)) << 20, 0);        // 2^(2 em - FIX_BITS)
      const double f2 = __hiloint2double((1023 + P.t_exp - FIX_BITS + 8 * GMIN + 1) << 20, 0);  // 2 * 2^(t_exp - FIX_BITS + 8 GMIN)
      const double pi_act = s * f1 * f2;
      const double pi_tot = P.pi_core ? pi_act + P.pi_core[p0 + p] : pi_act;
      P.pi[p0 + p] = pi_tot;
      if (guard_flag(sum_squares<NA>(phi), e_raw, fabs(pi_tot), P.guard_g1, P.guard_g2)) ++nflag;
    }
  }
#ifdef K2I_PROF
  if (blockIdx.x == 0 && (tid & 127) == 0) for (int i = 0; i < 8; ++i) k2i_prof[tid >> 7][i] = pf_[i];
#endif
  if (!ok && P.fail) atomicAdd(P.fail, 1);
  if (nflag && P.fail) atomicAdd(P.fail + 1, nflag);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace k2i
