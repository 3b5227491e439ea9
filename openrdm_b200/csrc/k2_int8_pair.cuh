// k2_int8_pair.cuh -- K2 on the 5th-generation tensor cores, CTA-pair form: the active-space on-top pair density
//     Pi_act(p) = sum_{I,J} X_pI D~_IJ X_pJ,   X_pI = phi_t(p) phi_u(p)   (I = packed pair t <= u),
// i.e. MCPDFT::build_ontop_pair_density (/root/reference/src/libmcpdft/mcpdft.cc:185-212) restricted to the active
// orbitals.  Same arithmetic as k2_int8.cuh (exact int8 products of the byte slices of 46-bit fixed-point images,
// D~ = T + T^T with T block-upper-triangular, slice pairs k + l >= GMIN, final dot with the unquantised X -- here with
// the orbital values to 51 bits; DESIGN.md section 4 "K2i"); what changes is where things live and who does what:
//   * two SMs of a cluster run tcgen05.mma.cta_group::2 (M = 256: 128 points per CTA); each CTA keeps only HALF of the
//     rows of every T chunk in shared memory (84 KB), which makes room for X slices 0..4 there (140 KB);
//   * TMEM then holds TWO full-width accumulators (2 x KP columns) + X slice 5: consecutive slice-pair groups alternate
//     between them, so the epilogue of one group runs under the MMAs of the next;
//   * the epilogue is spread over EW warps per TMEM lane quadrant (= per SM sub-partition): each thread owns one point
//     and a contiguous range of 16-pair units of its accumulator row, read with tcgen05.ld.x16 one unit ahead and folded
//     as sum_t phi_t (sum_u phi_u acc_tu) with the pair -> (t, u) map resolved at compile time -- on the INTEGER and FP32
//     pipes (ifold_elem below): while tcgen05.mma runs the SM's FP64 pipe issues 13x slower, so no FP64 instruction is
//     issued between a tile's first and last MMA;
//   * the same threads build the X slices of their pair range: one DFMA per pair leaves the balanced image in the
//     mantissa (the re-centring constant rides in the FMA addend), byte transposes cut it into slices;
//   * every point carries an a-posteriori error estimate of the 46-bit scheme (guard_flag below): points whose estimate
//     exceeds the parity bar 1e-12 max(1, |Pi|) are counted, and the host repeats the stage with the FP64 kernels.
// Only the leader CTA's MMA lane issues; tcgen05.commit multicasts "accumulator full" to both CTAs, both CTAs' epilogue
// threads arrive on the leader's "drained" barrier, and one cluster barrier per tile says "both CTAs' operands are built".
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <utility>
#include "k2_int8.cuh"

namespace k2i {

// shared-memory plan (dynamic) of one CTA of the pair: T half image | X slices 0..4
struct PairSmem {
  __host__ __device__ static size_t t_bytes(int kp) { return (size_t)NSLICE * Geom::slice_bytes(kp) / 2; }
  __host__ __device__ static size_t x_off(int kp) { return t_bytes(kp); }
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

// ---- build: the X slices of pairs [16 U0, 16 U1) of this thread's point --------------------------------------------
// slices 0..4 -> shared memory (K-major no-swizzle operand of 128 rows: 16 K bytes of a row are contiguous, so the
// four slice words of 16 pairs go out as one 16-byte store, conflict-free across the lanes of a warp), slice 5 -> TMEM
struct BuildRegs { uint32_t lo[4], hi[4], w[6][4]; };

template <int NA, int I>
__device__ __forceinline__ void build_elem(const double (&phi)[NA], const double (&ps)[NA], BuildRegs& r, uint8_t* x_row, uint32_t tmem_x5_row) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32;
  if constexpr (I < M) {
    constexpr int t = pair_t_of(NA, I), u = pair_u_of(NA, I);
    // bits(x + 1.5 2^52 + C) hold the integer round(x) + C in the low 48 bits: C = 0x8080808080 re-centres bytes 0..4
    const unsigned long long v = (unsigned long long)__double_as_longlong(fma(ps[t], phi[u], 6755399441055744.0 + 551911719040.0));
    r.lo[I & 3] = (uint32_t)v; r.hi[I & 3] = (uint32_t)(v >> 32);
  } else {   // padding pairs: zero digits (the image of 0 is 0x..80 80 80 80 80 before re-centring)
    r.lo[I & 3] = 0x80808080u; r.hi[I & 3] = 0x00000080u;
  }
  if constexpr ((I & 3) == 3) {
    constexpr int j = (I >> 2) & 3;
    const uint32_t t0 = __byte_perm(r.lo[0], r.lo[1], 0x5140), t1 = __byte_perm(r.lo[2], r.lo[3], 0x5140);
    const uint32_t t2 = __byte_perm(r.lo[0], r.lo[1], 0x7362), t3 = __byte_perm(r.lo[2], r.lo[3], 0x7362);
    const uint32_t h0 = __byte_perm(r.hi[0], r.hi[1], 0x5140), h1 = __byte_perm(r.hi[2], r.hi[3], 0x5140);
    r.w[0][j] = __byte_perm(t0, t1, 0x5410) ^ 0x80808080u;
    r.w[1][j] = __byte_perm(t0, t1, 0x7632) ^ 0x80808080u;
    r.w[2][j] = __byte_perm(t2, t3, 0x5410) ^ 0x80808080u;
    r.w[3][j] = __byte_perm(t2, t3, 0x7632) ^ 0x80808080u;
    r.w[4][j] = __byte_perm(h0, h1, 0x5410) ^ 0x80808080u;
    r.w[5][j] = __byte_perm(h0, h1, 0x7632);
    if constexpr (j == 3) {
      constexpr int k = I - 15;   // first K byte of the four words: a multiple of 16
#pragma unroll
      for (int s = 0; s < 5; ++s)
        *(uint4*)(x_row + (size_t)s * TILE * KP + (k >> 4) * (TILE / 8) * 128) = make_uint4(r.w[s][0], r.w[s][1], r.w[s][2], r.w[s][3]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tmem_x5_row + (uint32_t)(k >> 2)),
                   "r"(r.w[5][0]), "r"(r.w[5][1]), "r"(r.w[5][2]), "r"(r.w[5][3]) : "memory");
    }
  }
}

template <int NA, int I0, int... Is>
__device__ __forceinline__ void build_seq(const double (&phi)[NA], const double (&ps)[NA], BuildRegs& r, uint8_t* x_row, uint32_t tmem_x5_row,
                                          std::integer_sequence<int, Is...>) {
  (build_elem<NA, I0 + Is>(phi, ps, r, x_row, tmem_x5_row), ...);
}

template <int NA, int U0, int U1>
__device__ __forceinline__ void build_units(const double (&phi)[NA], double scale, uint8_t* x_row, uint32_t tmem_x5_row) {
  BuildRegs r;
  double ps[NA];
#pragma unroll
  for (int t = 0; t < NA; ++t) ps[t] = phi[t] * scale;
  build_seq<NA, 16 * U0>(phi, ps, r, x_row, tmem_x5_row, std::make_integer_sequence<int, 16 * (U1 - U0)>{});
}

// ---- epilogue of one group over the 16-pair units [U0, U1) of this thread's accumulator row ---------------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}

// last pair of orbital row t inside [lo, hi) (the row's running sum is folded into the total right after it)
__host__ __device__ constexpr bool row_ends_at(int na, int I, int hi, int m) {
  return I + 1 >= hi || I + 1 >= m || pair_t_of(na, I + 1) != pair_t_of(na, I);
}

// The FP64 form of the fold (round 2's first version; still used by the streamed-T lab kernel, tools/ozlab/
// k2_int8_stream.cuh; the product kernel below folds with ifold_elem).  Per accumulator element v (an exact int32) of pair (t, u):
//   groups G >  K2P_F32_MAX:  FP64.  v -> double exactly (2^52 bit trick: xor, hi word, DADD; every K2P_I2F_EVERY-th
//       element through I2F.F64.S32 instead, which runs on the otherwise idle conversion pipe), then one DFMA into the
//       row's running sum  sum_u phi_u v.
//   groups G <= K2P_F32_MAX:  FP32.  v -> float (I2FP, rounding 2^-24), one FFMA with phi_u 2^-em as a float.  These
//       groups weigh <= 2^-32 of the top group, so 2^-22-grade row sums change Pi by < 1e-15 of its scale
//       (tests/test_int8_scheme.py restates both paths; K2P_F32_MAX = 7 would cost 1e-13).
// Beside the running MMAs the SM issues roughly one instruction per 3 clk per sub-partition whatever the pipe
// (profiles/r02*_k2p_*.log: tensor and CUDA-core work add up instead of overlapping), so the instruction count is what
// is minimised here.  Two chains per row (even / odd pair); a row's sums are folded into the FP64 outer sum
// sum_t phi_t row_t as soon as the row ends, so few are live at any time.
#ifndef K2P_F32_MAX
#define K2P_F32_MAX 6
#endif
#ifndef K2P_I2F_EVERY
#define K2P_I2F_EVERY 3
#endif
#ifndef K2P_INTQ_MIN
#define K2P_INTQ_MIN 8   // groups >= this carry the FP32 remainder terms in the integer fold (ifold_elem)
#endif
struct RowAcc { double d0, d1; float f0, f1; };

// FLYF (the streaming kernel for > 20 active orbitals, k2_int8_stream.cuh): no float copy of phi is kept -- 26 orbitals
// as doubles AND floats do not fit the 128-register budget -- and the FP32 groups convert phi_u on the fly, unscaled
// (amplitudes outside float's range only occur where the whole group is far below the parity bar)
template <int NA, int I, int IEND, bool F32, bool FLYF = false>
__device__ __forceinline__ void fold_elem(const double (&phi)[NA], const float (&phif)[NA], RowAcc& r, double& tot, uint32_t v) {
  constexpr int M = NA * (NA + 1) / 2;
  if constexpr (I < M) {
    constexpr int t = pair_t_of(NA, I), u = pair_u_of(NA, I);
    if constexpr (F32) {
      const float vf = __int2float_rn((int)v);
      const float pf = FLYF ? __double2float_rn(phi[u]) : phif[u];
      if constexpr (I & 1) r.f1 = fmaf(pf, vf, r.f1); else r.f0 = fmaf(pf, vf, r.f0);
      if constexpr (row_ends_at(NA, I, IEND, M)) { tot = fma(phi[t], (double)(r.f0 + r.f1), tot); r.f0 = r.f1 = 0.f; }
    } else {
      double zd;
      if constexpr (K2P_I2F_EVERY > 0 && (I % (K2P_I2F_EVERY > 0 ? K2P_I2F_EVERY : 1)) == 0) zd = __int2double_rn((int)v);
      else zd = __hiloint2double(0x43300000, (int)(v ^ 0x80000000u)) - 4503601774854144.0;  // exact int32 -> double (2^52 + 2^31)
      if constexpr (I & 1) r.d1 = fma(phi[u], zd, r.d1); else r.d0 = fma(phi[u], zd, r.d0);
      if constexpr (row_ends_at(NA, I, IEND, M)) { tot = fma(phi[t], r.d0 + r.d1, tot); r.d0 = r.d1 = 0.0; }
    }
  }
}

template <int NA, int U, int UEND, bool F32, bool FLYF, int... Js>
__device__ __forceinline__ void fold_unit(const double (&phi)[NA], const float (&phif)[NA], RowAcc& r, double& tot, const uint32_t (&v)[16],
                                          std::integer_sequence<int, Js...>) {
  (fold_elem<NA, 16 * U + Js, 16 * UEND, F32, FLYF>(phi, phif, r, tot, v[Js]), ...);
}

// tcgen05.wait::ld with the loaded registers as in/out operands: nothing that reads them can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :: "memory");
}

template <int NA, int U, int U1, bool F32, bool FLYF = false>
__device__ __forceinline__ void fold_units(const double (&phi)[NA], const float (&phif)[NA], RowAcc& r, double& tot, uint32_t tmem_row,
                                           uint32_t (&cur)[16], uint32_t (&nxt)[16], uint32_t drain_bar) {
#ifndef K2P_LAB_NOLOAD   // lab switch (tools/ozlab): no TMEM loads, arithmetic on whatever the registers hold
  tmem_wait_ld16(cur);   // unit U is in `cur`
#endif
  if constexpr (U + 1 < U1) {
#ifndef K2P_LAB_NOLOAD
    tmem_ld16(tmem_row + (uint32_t)(16 * (U + 1)), nxt);
#endif
  } else {   // every column of this thread's range has been read: the group after next may overwrite this accumulator
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(drain_bar) : "memory");
  }
#ifndef K2P_LAB_NOMATH   // lab switch: loads and barriers only
  fold_unit<NA, U, U1, F32, FLYF>(phi, phif, r, tot, cur, std::make_integer_sequence<int, 16>{});
#else
  if (cur[0] == 0x12345678u && cur[15] == 0x9abcdefu) tot += 1.0;
#endif
  if constexpr (U + 1 < U1) fold_units<NA, U + 1, U1, F32, FLYF>(phi, phif, r, tot, tmem_row, nxt, cur, drain_bar);
}

template <int NA, int U0, int U1, bool F32, bool FLYF = false>
__device__ __forceinline__ double fold_group_units(const double (&phi)[NA], const float (&phif)[NA], uint32_t tmem_row, uint32_t drain_bar) {
  RowAcc r{0.0, 0.0, 0.f, 0.f};
  double tot = 0.0;
  uint32_t va[16], vb[16];
#ifndef K2P_LAB_NOLOAD
  tmem_ld16(tmem_row + (uint32_t)(16 * U0), va);
#else
#pragma unroll
  for (int j = 0; j < 16; ++j) { va[j] = tmem_row * (j + 1); vb[j] = tmem_row * (j + 17); }
#endif
  fold_units<NA, U0, U1, F32, FLYF>(phi, phif, r, tot, tmem_row, va, vb, drain_bar);
  return tot;
}

// ---- the fold without the FP64 pipe ---------------------------------------------------------------------------------------
// Measured (tools/microbench/mma_math_duty.cu, profiles/r02_mma_math_duty.log): while tcgen05.mma.kind::i8 runs, DFMA / DADD /
// DMUL of the same SM issue 13x slower (0.29 -> 0.022 per clk per sub-partition) -- the FP64 pipe and the tensor core
// exclude each other -- whereas FFMA, IMAD, IMAD.WIDE, I2FP and the conversion pipe keep their rate.  With the FP64 fold
// above the kernel therefore time-shares: MMAs, then fold, never both.  The fold below does the same sums on the integer
// and FP32 pipes; FP64 is left for three instructions per group.
//   phi'_t = phi_t 2^-em (|.| < 1)  =  h_t 2^-27 + r_t,   h_t = rint(phi'_t 2^27) (int32),   |r_t| <= 2^-28 (float q_t = r_t 2^-27)
//   row t of group G:   S_t = sum_u v h_u  (IMAD.WIDE, exact: |v| < 2^24, 20 terms -> |S| < 2^56)
//                       s_t = sum_u v q_u  (I2FP + FFMA; a 2^-28 correction, so FP32's 2^-24 is 2^-52 of the group)
//   group:  sum_t phi'_t (S_t 2^-27 + s_t 2^27)
//         = 2^-54 sum_t h_t S_t                      exact in two int64: S_t = Sh 2^32 + Sl (Sl signed), PH += Sh h_t, PL += Sl h_t
//         + 2^27 sum_t phi'_t s_t + sum_t q_t S_t    FP32 (Ca, Cb)
// FOLD_INT_Q: as above (groups 10, 9, 8, which weigh >= 2^-16 of the total and need 2^-31 or better inside the group);
// FOLD_INT: without the q terms (group 7: weight 2^-24, 27 bits of phi' are plenty);  FOLD_F32: FP32 throughout (groups
// <= K2P_F32_MAX: weight <= 2^-32), row ends included.  The results are in units of 2^(2 em).
enum { FOLD_INT_Q = 0, FOLD_INT = 1, FOLD_F32 = 2 };
struct IRow { long long S, S1; float s, f0, f1; };   // two IMAD.WIDE chains per row (even / odd pair): one alone is latency-bound
struct IGroup { long long PH, PL; float Ca, Cb; };

template <int NA, int I, int IEND, int MODE>
__device__ __forceinline__ void ifold_elem(const int (&h)[NA], const float (&q)[NA], const float (&phif)[NA], IRow& r, IGroup& g, uint32_t v) {
  constexpr int M = NA * (NA + 1) / 2;
  if constexpr (I < M) {
    constexpr int t = pair_t_of(NA, I), u = pair_u_of(NA, I);
    if constexpr (MODE == FOLD_F32) {
      const float vf = __int2float_rn((int)v);
      if constexpr (I & 1) r.f1 = fmaf(phif[u], vf, r.f1); else r.f0 = fmaf(phif[u], vf, r.f0);
      if constexpr (row_ends_at(NA, I, IEND, M)) { g.Ca = fmaf(phif[t], r.f0 + r.f1, g.Ca); r.f0 = r.f1 = 0.f; }
    } else {
      if constexpr (I & 1) r.S1 += (long long)(int)v * (long long)h[u]; else r.S += (long long)(int)v * (long long)h[u];
      if constexpr (MODE == FOLD_INT_Q) r.s = fmaf(__int2float_rn((int)v), q[u], r.s);
      if constexpr (row_ends_at(NA, I, IEND, M)) {
        r.S += r.S1;
        r.S1 = 0;
        const int sl = (int)(unsigned)r.S;
        const int sh = (int)(r.S >> 32) - (sl >> 31);   // S = sh 2^32 + sl with sl signed
        g.PH += (long long)sh * (long long)h[t];
        g.PL += (long long)sl * (long long)h[t];
        if constexpr (MODE == FOLD_INT_Q) {
          g.Ca = fmaf(phif[t], r.s, g.Ca);
          g.Cb = fmaf(q[t], fmaf(__int2float_rn(sh), 4294967296.0f, __int2float_rn(sl)), g.Cb);   // (float)S from its words: no 64-bit conversion
          r.s = 0.f;
        }
        r.S = 0;
      }
    }
  }
}

template <int NA, int U, int UEND, int MODE, int... Js>
__device__ __forceinline__ void ifold_unit(const int (&h)[NA], const float (&q)[NA], const float (&phif)[NA], IRow& r, IGroup& g, const uint32_t (&v)[16],
                                           std::integer_sequence<int, Js...>) {
  (ifold_elem<NA, 16 * U + Js, 16 * UEND, MODE>(h, q, phif, r, g, v[Js]), ...);
}

template <int NA, int U, int U1, int MODE>
__device__ __forceinline__ void ifold_units(const int (&h)[NA], const float (&q)[NA], const float (&phif)[NA], IRow& r, IGroup& g, uint32_t tmem_row,
                                            uint32_t (&cur)[16], uint32_t (&nxt)[16], uint32_t drain_bar) {
#ifndef K2P_LAB_NOLOAD   // lab switch (tools/ozlab): no TMEM loads, arithmetic on whatever the registers hold
  tmem_wait_ld16(cur);   // unit U is in `cur`
#endif
  if constexpr (U + 1 < U1) {
#ifndef K2P_LAB_NOLOAD
    tmem_ld16(tmem_row + (uint32_t)(16 * (U + 1)), nxt);
#endif
  } else {   // every column of this thread's range has been read: the group after next may overwrite this accumulator
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(drain_bar) : "memory");
  }
#ifndef K2P_LAB_NOMATH   // lab switch: loads and barriers only
  ifold_unit<NA, U, U1, MODE>(h, q, phif, r, g, cur, std::make_integer_sequence<int, 16>{});
#else
  if (cur[0] == 0x12345678u && cur[15] == 0x9abcdefu) g.Ca += 1.f;
#endif
  if constexpr (U + 1 < U1) ifold_units<NA, U + 1, U1, MODE>(h, q, phif, r, g, tmem_row, nxt, cur, drain_bar);
}

// What a thread carries from group to group within a tile: not one FP64 instruction is issued between the tile's first
// and last MMA (three DFMAs per group were a fifth of all stall samples, profiles/r02_ncu_full_k2i_intfold.csv).
//   A: sum over the integer groups of (PH 2^32 + PL) 2^(8 (G - 7)), a 128-bit integer (< 2^112);
//   C: the FP32 parts, same weights: remainder terms of groups >= K2P_INTQ_MIN and the groups <= K2P_F32_MAX whole.
// value of the tile's groups in units of 2^(2 em) 2^(8 * 7):  A 2^-54 + C  (tile_acc_value, after the last MMA).
struct TileAcc { __int128 A; float C; };

__device__ __forceinline__ double tile_acc_value(const TileAcc& a) {
  const long long hi = (long long)(a.A >> 64);
  const unsigned long long lo = (unsigned long long)a.A;
  const double v = fma(__ll2double_rn(hi), 18446744073709551616.0, __ull2double_rn(lo));   // |A| < 2^112: exact to 2^-53
  return fma(v, 5.5511151231257827e-17 /* 2^-54 */, (double)a.C);
}

// one group (G) over this thread's units, added into the tile's accumulators
template <int NA, int U0, int U1, int MODE>
__device__ __forceinline__ void ifold_group_units(const int (&h)[NA], const float (&q)[NA], const float (&phif)[NA], uint32_t tmem_row, uint32_t drain_bar,
                                                  int G, TileAcc& acc) {
  IRow r{0, 0, 0.f, 0.f, 0.f};
  IGroup g{0, 0, 0.f, 0.f};
  uint32_t va[16], vb[16];
#ifndef K2P_LAB_NOLOAD
  tmem_ld16(tmem_row + (uint32_t)(16 * U0), va);
#else
#pragma unroll
  for (int j = 0; j < 16; ++j) { va[j] = tmem_row * (j + 1); vb[j] = tmem_row * (j + 17); }
#endif
  ifold_units<NA, U0, U1, MODE>(h, q, phif, r, g, tmem_row, va, vb, drain_bar);
  const float w = __int_as_float((127 + 8 * (G - 7)) << 23);   // 2^(8 (G - 7))
  if constexpr (MODE == FOLD_F32) {
    acc.C = fmaf(g.Ca, w, acc.C);
  } else {
    if constexpr (MODE == FOLD_INT_Q) acc.C = fmaf(fmaf(g.Ca, 134217728.0f, g.Cb), w, acc.C);   // (2^27 Ca + Cb) 2^(8 (G - 7))
    acc.A += (((__int128)g.PH << 32) + (__int128)g.PL) << (8 * (G - 7));
  }
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

// unit range of epilogue part `part` of EW: the KP / 16 units split as evenly as whole units allow, larger parts first
__host__ __device__ constexpr int part_u0(int nunits, int ew, int part) { return part * (nunits / ew) + (part < nunits % ew ? part : nunits % ew); }

// Group order within a tile.  Two accumulators ping-pong: the MMAs of group n + 2 wait for the fold of group n.  Groups
// differ both ways -- MMAs: 1 .. 6 slice pairs (0.5 k .. 3.1 k clk); fold: integer + remainder terms (10, 9, 8), integer
// (7), FP32 (<= 6).  The tile starts with the single-pair group (the epilogue starts 0.5 k clk in), its heaviest MMA group
// follows (it covers the first folds), and it ends with an FP32 group: the fold nothing covers is the cheapest, and it is
// the one during which the next tile's orbitals are loaded (pair_compute_role).  Eight orders measured on 2 M points
// (profiles/r02_k2p_group_orders.log): 1.33 .. 1.38 ms; the old order by descending weight, 10 9 8 7 4 6 5: 1.40.
#ifndef K2P_ORDER4
#define K2P_ORDER4 0x467895Au   // 10 5 9 8 7 6 4 (first group in the low nibble)
#endif
#ifndef K2P_ORDER5
#define K2P_ORDER5 0x67895Au    // 10 5 9 8 7 6
#endif
__host__ __device__ constexpr int group_at_c(int i, int gmin) { return (int)(((gmin == 4 ? K2P_ORDER4 : K2P_ORDER5) >> (4 * i)) & 15u); }
__device__ __forceinline__ int group_at(int i, int gmin) {
  // GMIN 4: 10 5 9 6 7 8 4;  GMIN 5: 10 5 9 7 8 6   (4-bit entries, first group in the low nibble)
  return (int)(((gmin == 4 ? K2P_ORDER4 : K2P_ORDER5) >> (4 * i)) & 15u);
}

// register cap: 13 warps (EW = 3) are allocated as 16; at 96 registers the CTA leaves 16 K registers to a co-resident
// density CTA (k1_core_light), at 128 it owns the register file
#ifndef K2P_MAXREG
#define K2P_MAXREG 128
#endif

#ifdef K2I_PROF
__device__ long long k2p_prof[3][8];
#define K2P_LAP(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && prof_slot >= 0) { const long long t_ = clock64(); pf_[i] += t_ - tc_; tc_ = t_; } } while (0)
#else
#define K2P_LAP(i) do { } while (0)
#endif

// what every thread of the CTA knows after the prologue
struct PairCtx {
  uint32_t tmem, lane_addr, crank;
  uint32_t bar_full, bar_drain, drain_remote;   // shared-memory addresses; drain_remote = the leader's copy as seen from this CTA
  uint8_t* x_row;        // this thread's row of the X operand (slice 0) in shared memory
  double* s_part;        // [EW - 1][TILE] partial sums of parts 1 .. EW - 1
  uint64_t tdesc0, xdesc0;
  uint32_t idesc0;
  int p;                 // point within the CTA's 128-point tile
};

// ---- the epilogue / build role: one instantiation per unit range, so that a thread keeps in registers only the
// orbital values its own rows and columns need
template <int NA, int GMIN, int EW, int PART>
__device__ __forceinline__ void pair_compute_role(const Params& P, const PairCtx& c, bool& ok, int& nflag) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, NUNITS = KP / 16;
  constexpr int U0 = part_u0(NUNITS, EW, PART), U1 = part_u0(NUNITS, EW, PART + 1);
  constexpr int ACC_COLS = KP, X5_COL = 2 * KP;
  constexpr int NGROUPS = 2 * (NSLICE - 1) - GMIN + 1;
  constexpr int ETHREADS = 4 * EW * 32;
  const int p = c.p;
#ifdef K2I_PROF
  long long pf_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc_ = clock64();
  const int prof_slot = (threadIdx.x == 0) ? 0 : (threadIdx.x == 32 ? 1 : -1);   // warp 0 shares its sub-partition with the MMA warp, warp 1 does not
#endif
  uint32_t n = 0;   // groups consumed so far (each accumulator sees every other one)
  const size_t ntiles = (P.npts + 2 * TILE - 1) / (2 * TILE);   // pair tiles of 256 points
  const size_t tstride = gridDim.x / 2;
  // The orbital values of a tile are loaded during the LAST group of the tile before it: that group is folded in FP32
  // (static_assert below), which needs neither h nor q, so their 40 registers are free for the 20 doubles and the load
  // latency (1.4 k clk of every tile alone, 3 k beside the core-orbital stream) disappears behind that group's MMAs.
  static_assert(group_at_c(NGROUPS - 1, GMIN) <= K2P_F32_MAX, "the tile's last group must be an FP32 one");
  double phi[NA];
  auto load_phi = [&](size_t tile) {
    const size_t pp = tile * 2 * TILE + (size_t)c.crank * TILE + p;
#pragma unroll
    for (int t = 0; t < NA; ++t) phi[t] = (tile < ntiles && pp < P.npts) ? __ldg(P.phi_act + (size_t)t * P.ld + pp) : 0.0;
  };
  load_phi(blockIdx.x / 2);
  for (size_t tile = blockIdx.x / 2; tile < ntiles; tile += tstride) {
    const size_t p0 = tile * 2 * TILE + (size_t)c.crank * TILE;
    double s2 = 0.0;   // sum_t phi_t^2 for the error guard
    K2P_LAP(7);
    uint32_t mh = 0;   // max over t of the high word of |phi_t|: its exponent field is that of max |phi_t|
#pragma unroll
    for (int t = 0; t < NA; ++t) mh = max(mh, (uint32_t)__double2hiint(phi[t]) & 0x7FFFFFFFu);
    // row scale: max |X| = (max_t |phi_t|)^2 < 2^(2 em); the clamps keep the scale factors finite normal numbers
    const int e_raw = min(max((int)(mh >> 20), 534), 1500);
    const double scale = __hiloint2double((3113 - 2 * e_raw + (FIX_BITS - 46)) << 20, 0);  // 2^(FIX_BITS - 2 em)
    K2P_LAP(0);
    build_units<NA, U0, U1>(phi, scale, c.x_row, c.tmem + c.lane_addr + (uint32_t)X5_COL);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if constexpr (PART == 0) s2 = sum_squares<NA>(phi);
    // phi'_t = phi_t 2^-em (|.| < 1) in the three forms the fold uses (ifold_elem): float, 27-bit integer part, remainder
    float phif[NA], q[NA];
    int h[NA];
    {
      const double sc0 = __hiloint2double((2045 - e_raw) << 20, 0);   // 2^-em, em = e_raw - 1022
#pragma unroll
      for (int t = 0; t < NA; ++t) {
        // ps + 1.5 2^25 has ulp 2^-27: its low mantissa word is h = rint(ps 2^27) (two's complement), and taking the
        // constant off again leaves h 2^-27 exactly -- no F2I / I2F (16 per clk per SM) in the tile's prologue
        const double ps = phi[t] * sc0;
        const double hd = ps + 50331648.0;
        h[t] = __double2loint(hd);
        const float rf = (float)(ps - (hd - 50331648.0));         // |r| <= 2^-28, 24 bits of it
        q[t] = rf * 7.450580596923828125e-9f;                      // r 2^-27
        phif[t] = fmaf(__int2float_rn(h[t]), 7.450580596923828125e-9f, rf);
      }
    }
    {   // next pair tile's orbitals towards L2 while this one is contracted
      const size_t pn = p0 + tstride * 2 * TILE + p;
      if (pn < P.npts) {
#pragma unroll
        for (int t = PART; t < NA; t += EW) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.phi_act + (size_t)t * P.ld + pn));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    cluster_sync_all();   // both CTAs' operands are built
    K2P_LAP(1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TileAcc acc{0, 0.f};
#pragma unroll 1
    for (int gi = 0; gi < NGROUPS - 1; ++gi, ++n) {
      const int G = group_at(gi, GMIN);
      const uint32_t b = n & 1, use = n >> 1;   // accumulator, and how often it has been used before
      ok = wait_parity(c.bar_full + 8 * b, use & 1) && ok;
      K2P_LAP(5);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_row = c.tmem + c.lane_addr + (uint32_t)(b * ACC_COLS), db = c.drain_remote + 8 * b;
      if (G > K2P_INTQ_MIN - 1) ifold_group_units<NA, U0, U1, FOLD_INT_Q>(h, q, phif, tmem_row, db, G, acc);
      else if (G > K2P_F32_MAX) ifold_group_units<NA, U0, U1, FOLD_INT>(h, q, phif, tmem_row, db, G, acc);
      else ifold_group_units<NA, U0, U1, FOLD_F32>(h, q, phif, tmem_row, db, G, acc);
      K2P_LAP(6);
    }
    // the next tile's orbitals (and this tile's core term) are asked for now, the tile's last group is folded meanwhile
    load_phi(tile + tstride);
    double pi_core_p = 0.0;
    if constexpr (PART == 0) pi_core_p = (P.pi_core && p0 + p < P.npts) ? __ldg(P.pi_core + p0 + p) : 0.0;
    {
      constexpr int G = group_at_c(NGROUPS - 1, GMIN);
      const uint32_t b = n & 1, use = n >> 1;
      ok = wait_parity(c.bar_full + 8 * b, use & 1) && ok;
      K2P_LAP(5);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      ifold_group_units<NA, U0, U1, FOLD_F32>(h, q, phif, c.tmem + c.lane_addr + (uint32_t)(b * ACC_COLS), c.drain_remote + 8 * b, G, acc);
      ++n;
      K2P_LAP(6);
    }
    // every epilogue thread has seen the last group's "full": all MMAs of the tile are done, the X slices are dead
    // (FP64 again: the tensor pipe is idle from here to the next tile's first MMA)
    double pi_acc = tile_acc_value(acc) * __hiloint2double((2 * e_raw - 1021 + 8 * (7 - GMIN)) << 20, 0);   // 2^(2 em) 2^(8 (7 - GMIN)), em = e_raw - 1022
    if constexpr (PART > 0) c.s_part[(PART - 1) * TILE + p] = pi_acc;
    asm volatile("bar.sync 1, %0;" ::"r"(ETHREADS) : "memory");
    if constexpr (PART == 0) {
      double s = pi_acc;
#pragma unroll
      for (int k = 1; k < EW; ++k) s += c.s_part[(k - 1) * TILE + p];
      // Pi = 2 * 2^(2 em - FIX) * 2^(t_exp - FIX) * 2^(8 GMIN) * (sum over the parts)
      const double f1 = __hiloint2double((2 * e_raw - 1067 - (FIX_BITS - 46)) << 20, 0);        // 2^(2 em - FIX_BITS)
      const double f2 = __hiloint2double((1023 + P.t_exp - FIX_BITS + 8 * GMIN + 1) << 20, 0);  // 2 * 2^(t_exp - FIX_BITS + 8 GMIN)
      const double pi_act = s * f1 * f2;
      if (p0 + p < P.npts) {
        const double pi_tot = pi_act + pi_core_p;
        P.pi[p0 + p] = pi_tot;
        if (guard_flag(s2, e_raw, fabs(pi_tot), P.guard_g1, P.guard_g2)) ++nflag;
      }
    }
    // (s_part has its own shared memory: its next writer is a whole tile and a cluster barrier away)
    K2P_LAP(4);
  }
#ifdef K2I_PROF
  if (blockIdx.x == 0 && prof_slot >= 0) for (int i = 0; i < 8; ++i) k2p_prof[prof_slot][i] = pf_[i];
#endif
}

// ---- the MMA role: one warp; lane 0 of the leader CTA issues, the other CTA's warp only keeps the cluster barriers company
template <int NA, int GMIN>
__device__ __forceinline__ void pair_mma_role(const Params& P, const PairCtx& c, bool& ok) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32;
  constexpr int ACC_COLS = KP, X5_COL = 2 * KP;
  constexpr int NGROUPS = 2 * (NSLICE - 1) - GMIN + 1;
  // the issuing lane is chosen by elect.sync, not by lane id: inside a region ptxas knows to be single-lane the MMA's
  // descriptor operands move to uniform registers with one R2UR each; behind `lane == 0` every tcgen05.mma costs a
  // broadcast-and-retry loop (ELECT, R2UR.BROADCAST, BRA.U.ANY) -- instructions that compete with the epilogue warps of
  // the same sub-partition for issue slots (tools/microbench/mma_math_duty.cu: that competition is what slows the MMAs)
  uint32_t elected;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
#ifdef K2I_PROF
  long long pf_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc_ = clock64();
  const int prof_slot = elected ? 2 : -1;
#endif
  uint32_t n = 0;
  const size_t ntiles = (P.npts + 2 * TILE - 1) / (2 * TILE);
  for (size_t tile = blockIdx.x / 2; tile < ntiles; tile += gridDim.x / 2) {
    K2P_LAP(7);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    cluster_sync_all();   // both CTAs' operands are built
    K2P_LAP(1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int gi = 0; gi < NGROUPS; ++gi, ++n) {
      const int G = group_at(gi, GMIN);
      const uint32_t b = n & 1, use = n >> 1;
      if (elected && c.crank == 0) {
        if (use > 0) ok = wait_parity(c.bar_drain + 8 * b, (use - 1) & 1) && ok;   // both CTAs have read this accumulator's previous group
        K2P_LAP(2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifndef K2P_LAB_NOMMA   // lab switch: the epilogue without tensor-core traffic beside it
        issue_group_pair_of<KP>(G, c.tmem + b * ACC_COLS, c.tmem + X5_COL, c.tdesc0, c.xdesc0, c.idesc0);
#endif
        commit2(c.bar_full + 8 * b);
        K2P_LAP(3);
      }
      __syncwarp();
    }
  }
#ifdef K2I_PROF
  if (blockIdx.x == 0 && prof_slot >= 0) for (int i = 0; i < 8; ++i) k2p_prof[prof_slot][i] = pf_[i];
#endif
}

// P.timg here is the PAIR image (ordm::pack_t_int8_pair): CTA rank c of the cluster reads block c.
// P.guard_g1/g2: constants of guard_flag; P.fail[0] counts barrier time-outs, P.fail[1] flagged points.
template <int NA, int GMIN, int EW>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(K2P_MAXREG) k2i_ontop_pair(Params P) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, NUNITS = KP / 16;
  constexpr int ETHREADS = 4 * EW * 32;
  static_assert(2 * KP + KP / 4 <= 512, "TMEM columns");
  static_assert(NSLICE == 6 && (GMIN == 4 || GMIN == 5), "six slices, groups 10 .. GMIN");
  static_assert(K2P_F32_MAX >= 6 && K2P_INTQ_MIN > K2P_F32_MAX, "the integer groups are weighted 2^(8 (G - 7)) >= 1 in TileAcc");
  static_assert(EW >= 2 && EW <= 3 && NUNITS >= EW, "epilogue warps per lane quadrant (EW = 4: no room for s_part_s)");
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ double s_part_s[(EW - 1) * TILE];   // the other parts' per-point sums on their way to part 0
  __shared__ __align__(8) uint64_t bar_s[4];   // [0..1] accumulator b complete (multicast commit); [2..3] (leader's copy) accumulator b drained by both CTAs
  uint8_t* s_t = smem;
  uint8_t* s_x = smem + PairSmem::x_off(KP);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, part = warp >> 2;   // lane quadrant, unit range (part == EW: the MMA warp)
  PairCtx c;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(c.crank));
  {
    const int4* src = (const int4*)(P.timg + (size_t)c.crank * PairSmem::t_bytes(KP));
    for (int i = tid; i < (int)(PairSmem::t_bytes(KP) / 16); i += blockDim.x) ((int4*)s_t)[i] = src[i];
  }
  c.bar_full = (uint32_t)__cvta_generic_to_shared(&bar_s[0]);
  c.bar_drain = (uint32_t)__cvta_generic_to_shared(&bar_s[2]);
  if (tid == 0) {
    for (int h = 0; h < 2; ++h) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.bar_full + 8 * h));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c.bar_drain + 8 * h), "r"(2 * ETHREADS));
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
  c.tmem = tmem_base_s;
  c.lane_addr = (uint32_t)(q * 32) << 16;
  c.idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * TILE) >> 4) << 24);   // M = 256 over the pair
  c.p = 32 * q + lane;
  c.x_row = s_x + (c.p >> 3) * 128 + (c.p & 7) * 16;
  c.s_part = s_part_s;
  c.tdesc0 = smem_desc((uint32_t)__cvta_generic_to_shared(s_t), 0, 128);
  c.xdesc0 = smem_desc((uint32_t)__cvta_generic_to_shared(s_x), (TILE / 8) * 128, 128);
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(c.drain_remote) : "r"(c.bar_drain), "r"(0));
  bool ok = true;
  int nflag = 0;   // points of this thread whose error estimate exceeds the parity bar (guard_flag)
  if (part == 0) pair_compute_role<NA, GMIN, EW, 0>(P, c, ok, nflag);
  else if (part == 1) pair_compute_role<NA, GMIN, EW, 1>(P, c, ok, nflag);
  else if (EW > 2 && part == 2) pair_compute_role<NA, GMIN, EW, (EW > 2 ? 2 : 0)>(P, c, ok, nflag);
  else if (EW > 3 && part == 3) pair_compute_role<NA, GMIN, EW, (EW > 3 ? 3 : 0)>(P, c, ok, nflag);
  else pair_mma_role<NA, GMIN>(P, c, ok);
  if (P.fail) {
    if (!ok) atomicAdd(P.fail, 1);
    if (nflag) atomicAdd(P.fail + 1, nflag);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(c.tmem) : "memory");
}

}  // namespace k2i
