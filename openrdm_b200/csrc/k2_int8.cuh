// k2_int8.cuh -- K2 on the 5th-generation tensor cores: the active-space on-top pair density
//     Pi_act(p) = sum_{I,J} X_pI D~_IJ X_pJ,   X_pI = phi_t(p) phi_u(p)   (I = packed pair t <= u),
// i.e. MCPDFT::build_ontop_pair_density (/root/reference/src/libmcpdft/mcpdft.cc:185-212) restricted to the active
// orbitals, contracted by tcgen05.mma.kind::i8 through an Ozaki-style exact integer split:
//   * D~ = T + T^T with T block-upper-triangular (32 x 32 blocks, diagonal blocks halved), so Pi = 2 sum_J X_J (X T)_J;
//   * every X row (against (max_t |phi_t|)^2) and T (against max |T|) becomes a 48-bit fixed-point image cut into
//     six balanced signed bytes; the slice pairs (k, l) with k + l >= GMIN are multiplied exactly (INT32
//     accumulators in TMEM), one accumulator per group G = k + l, highest group first;
//   * X slices 0..4 live in TMEM (A operand from TMEM, written with tcgen05.st), slice 5 and all of T in shared
//     memory (K-major, no swizzle: 8 x 16-byte core matrices), the accumulator takes the remaining TMEM columns;
//   * the epilogue folds two groups in INT32 (Z = 256 acc_G + acc_{G-1}) and dots them with the FP64 X row.
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
constexpr int FIX_BITS = 46;    // |fixed-point image| <= 2^46 (one bit of headroom for the balancing carry)

struct Geom {                   // everything derived from the padded pair count KP (multiple of 32)
  int kp, nchunk;
  __host__ __device__ static constexpr int chunk_rows(int kp, int kc) { return kp - 32 * kc; }                 // N of K-chunk kc
  __host__ __device__ static constexpr int chunk_off(int kp, int kc) { return 32 * (kc * kp - 16 * kc * (kc - 1)); }  // bytes before chunk kc
  __host__ __device__ static constexpr int slice_bytes(int kp) { return chunk_off(kp, kp / 32); }
};

// byte offset of element (row r, K-byte k) in a K-major no-swizzle canonical operand with R rows
__host__ __device__ inline int canon(int r, int k, int R) { return ((k >> 4) * (R >> 3) + (r >> 3)) * 128 + (r & 7) * 16 + (k & 15); }

struct Params {
  const double* phi_act;   // [nact][ld]: orbital t of point p at phi_act[t * ld + p]
  size_t ld, npts;
  const uint8_t* timg;     // NSLICE x slice_bytes(kp): balanced byte slices of T in shared-memory order (ordm::pack_t_int8)
  int t_exp;               // max |T| < 2^t_exp
  const double* pi_core;   // optional: closed-form core part, added here (nullptr: K3 adds it)
  double* pi;              // [npts] out
  int* fail;               // bumped if an MMA-completion barrier ever times out
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
__device__ __forceinline__ bool wait_parity(uint32_t bar, uint32_t parity) {
  for (long spin = 0; spin < (1l << 24); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}

// shared-memory plan (dynamic): T image | X slice 5 (canonical 128 x KP) | partial sums [2][128]  (pass nact = 0)
struct Smem {
  __host__ __device__ static size_t t_bytes(int kp) { return (size_t)NSLICE * Geom::slice_bytes(kp); }
  __host__ __device__ static size_t a5_off(int kp) { return t_bytes(kp); }
  __host__ __device__ static size_t phi_off(int kp) { return a5_off(kp) + (size_t)TILE * kp; }
  __host__ __device__ static size_t part_off(int kp, int nact) { return phi_off(kp) + (size_t)nact * TILE * 8; }
  __host__ __device__ static size_t bytes(int kp, int nact_next) { return part_off(kp, 0) + 2 * TILE * 8 + (size_t)nact_next * TILE * 8; }
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

template <int NA, int R>
__device__ __forceinline__ double fold_half(const double (&phi)[NA], const int* z) {
  double part0 = 0.0, part1 = 0.0;
  for_my_pairs<NA, R>([&](int i, int t, int u) {
    const double zd = __hiloint2double(0x43300000, (int)((uint32_t)z[i] ^ 0x80000000u)) - 4503601774854144.0;  // 2^52 + 2^31
    if (i & 1) part1 = fma(phi[t] * phi[u], zd, part1); else part0 = fma(phi[t] * phi[u], zd, part0);
  });
  return part0 + part1;
}

// all MMAs of group G = k + l (X slice k, T slice l), every descriptor a compile-time offset from a base:
// the single issuing lane spends a handful of instructions per MMA
template <int G, int KP>
__device__ __forceinline__ void issue_group(uint32_t tmem, uint64_t bdesc0, uint64_t adesc5, uint32_t idesc0) {
  constexpr int NCH = KP / 32, SL_COLS = KP / 4, A_COL = KP, KMAX = G < NSLICE - 1 ? G : NSLICE - 1, KMIN = G > NSLICE - 1 ? G - (NSLICE - 1) : 0;
#pragma unroll
  for (int k = KMAX; k >= KMIN; --k) {
    const int l = G - k;
#pragma unroll
    for (int kc = 0; kc < NCH; ++kc) {
      const int nrows = Geom::chunk_rows(KP, kc);
      const uint32_t idesc = idesc0 | ((uint32_t)(nrows >> 3) << 17);
      const uint64_t bd = bdesc0 + (uint64_t)((l * Geom::slice_bytes(KP) + Geom::chunk_off(KP, kc)) >> 4) + ((uint64_t)(nrows * 16 >> 4) << 16);
      const uint32_t first = (k == KMAX && kc == 0) ? 0u : 1u;
      if (k < 5) mma_ts(tmem + 32 * kc, tmem + (uint32_t)(A_COL + SL_COLS * k + 8 * kc), bd, idesc, first);
      else mma_ss(tmem + 32 * kc, adesc5 + (uint64_t)((kc * 2 * (TILE / 8) * 128) >> 4), bd, idesc, first);
    }
  }
}

// GMIN: the slice pairs with k + l >= GMIN are kept (4: 26 pairs in 7 groups; 5: 21 pairs in 6 groups)
template <int NA, int GMIN>
__global__ void __launch_bounds__(THREADS, 1) k2i_ontop(Params P) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32;
  constexpr int EPT = KP / 2;
  static_assert(EPT % 16 == 0, "geometry");
  constexpr int A_COL = KP;           // TMEM: accumulator columns [0, KP), X slice k < 5 at A_COL + (KP/4) k
  constexpr int SL_COLS = KP / 4;
  static_assert(A_COL + 5 * SL_COLS <= 512, "TMEM columns");
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_s;
  uint8_t* s_t = smem;
  uint8_t* s_a5 = smem + Smem::a5_off(KP);
  double* s_part = (double*)(smem + Smem::part_off(KP, 0));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = warp < CWARPS;
  const int q = warp & 3, r = (warp >> 2) & 1, p = 32 * q + lane;  // lane quadrant, K half, point within the tile

  for (int i = tid; i < (int)(Smem::t_bytes(KP) / 16); i += THREADS) ((int4*)s_t)[i] = ((const int4*)P.timg)[i];
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bar_s);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
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
  static_assert(NSLICE == 6 && (GMIN == 4 || GMIN == 5), "the group dispatch below is written for six slices");
  uint32_t phase = 0;
  bool ok = true;

  const size_t ntiles = (P.npts + TILE - 1) / TILE;
  double* s_nxt = (double*)(smem + Smem::part_off(KP, 0) + 2 * TILE * 8);   // [NA][TILE]: the next tile's orbitals (cp.async)
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
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();   // S1: this tile's orbitals are in shared memory
    if (compute) {
      double m = 0.0;
#pragma unroll
      for (int t = 0; t < NA; ++t) { phi[t] = s_nxt[t * TILE + p]; m = fmax(m, fabs(phi[t])); }
      // row scale: max |X| = (max_t |phi_t|)^2 < 2^(2 em)
      e_raw = max((__double2hiint(m) >> 20) & 0x7FF, 534);
      const double scale = __hiloint2double((3113 - 2 * e_raw + (FIX_BITS - 46)) << 20, 0);  // 2^(FIX_BITS - 2 em)
      const uint32_t tmem_row = tmem + lane_addr + (uint32_t)A_COL;
      if (r == 0) build_half<NA, 0>(phi, scale, tmem_row, a5_row); else build_half<NA, 1>(phi, scale, tmem_row, a5_row);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // S2: operands ready (and every thread has read its orbitals)
    if (compute) prefetch(tile + gridDim.x);
    double pi_acc = 0.0;
    int z[EPT];
    auto fold = [&](int Gdone) {   // the dot of a finished pair of groups with the FP64 X row
      const double part = r == 0 ? fold_half<NA, 0>(phi, z) : fold_half<NA, 1>(phi, z);
      pi_acc = fma(part, __hiloint2double((1023 + 8 * (Gdone - GMIN)) << 20, 0), pi_acc);
    };
#pragma unroll 1
    for (int G = 2 * (NSLICE - 1); G >= GMIN; --G) {
      const bool upper = ((2 * (NSLICE - 1) - G) & 1) == 0;   // G = 10, 8, 6, (4) start a pair of groups; G = 9, 7, 5 finish it
      if (!compute) {
        if (lane == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          switch (G) {
            case 10: issue_group<10, KP>(tmem, bdesc0, adesc5, idesc0); break;
            case 9: issue_group<9, KP>(tmem, bdesc0, adesc5, idesc0); break;
            case 8: issue_group<8, KP>(tmem, bdesc0, adesc5, idesc0); break;
            case 7: issue_group<7, KP>(tmem, bdesc0, adesc5, idesc0); break;
            case 6: issue_group<6, KP>(tmem, bdesc0, adesc5, idesc0); break;
            case 5: issue_group<5, KP>(tmem, bdesc0, adesc5, idesc0); break;
            default: issue_group<4, KP>(tmem, bdesc0, adesc5, idesc0); break;
          }
          commit(bar);
          ok = wait_parity(bar, phase) && ok;
        }
        phase ^= 1;
        __syncwarp();
      } else if (upper && G != 2 * (NSLICE - 1)) {
        fold(G + 1);   // previous pair of groups, while this group's MMAs run
      }
      __syncthreads();   // S3: accumulator complete
      if (compute) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (upper) {
#pragma unroll
          for (int j0 = 0; j0 < EPT; j0 += 8) tmem_ld8(tmem + lane_addr + (uint32_t)(EPT * r + j0), (uint32_t*)&z[j0]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int h = 0; h < EPT; h += 16) {
            uint32_t v[16];
            tmem_ld8(tmem + lane_addr + (uint32_t)(EPT * r + h), &v[0]);
            tmem_ld8(tmem + lane_addr + (uint32_t)(EPT * r + h + 8), &v[8]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) z[h + j] = z[h + j] * 256 + (int)v[j];
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
      __syncthreads();   // S4: the accumulator may be overwritten by the next group
    }
    if (compute) {
      fold(GMIN);
      s_part[r * TILE + p] = pi_acc;
    }
    __syncthreads();   // S5
    // ---- Pi = 2 * 2^(2 em - FIX) * 2^(t_exp - FIX) * 2^(8 GMIN) * (sum over the two K halves)
    if (compute && r == 0 && p0 + p < P.npts) {
      const double s = s_part[p] + s_part[TILE + p];
      const double f1 = __hiloint2double((2 * e_raw - 1067 - (FIX_BITS - 46)) << 20, 0);        // 2^(2 em - FIX_BITS)
      const double f2 = __hiloint2double((1023 + P.t_exp - FIX_BITS + 8 * GMIN + 1) << 20, 0);  // 2 * 2^(t_exp - FIX_BITS + 8 GMIN)
      const double pi_act = s * f1 * f2;
      P.pi[p0 + p] = P.pi_core ? pi_act + P.pi_core[p0 + p] : pi_act;
    }
  }
  if (!ok && P.fail) atomicAdd(P.fail, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace k2i
