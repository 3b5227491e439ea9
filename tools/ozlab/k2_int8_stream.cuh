// k2_int8_stream.cuh -- LAB (tools/ozlab; not linked into the product).  A NEGATIVE RESULT, kept with its measurements:
// the int8 tensor-core on-top kernel for 21 .. 26 active orbitals with the T image STREAMED through a shared-memory ring.
// It is correct (max |dPi| 1.2e-14 of max(1, |Pi|) on 2 M points, no barrier time-outs) but 2.3x SLOWER than the FP64
// DMMA kernel it was meant to replace for (26e,26o): 17.1 ms against 7.5 ms per 2 M points (profiles/r02_k2s_stream.log).
// Why: at KP = 352 the X slices alone need 264 KB (188 KB of shared memory + the 160 spare TMEM columns), which leaves
// 36 KB for T blocks in flight, while the MMA stream consumes 32 B/clk per CTA and one ring round trip (tcgen05.commit
// -> producer -> cp.async.bulk from L2 -> full barrier, + the remote arrive that tells the leader CTA about the other
// CTA's half) is ~3 k clk: 9 slots cover a third of it, and the measured pace is ~820 clk per MMA instead of ~110.
// Two point tiles per T block would halve the T rate but need a second 352-column accumulator (704 > 512 TMEM columns)
// and a second copy of X.  So (26e,26o) stays on the DMMA kernel; what follows is the kernel as measured.
//
// k2_int8_stream.cuh -- K2 on the 5th-generation tensor cores for 21 .. 26 active orbitals (KP = 256 .. 352 padded
// orbital pairs; BASELINE.json's largest config, (26e,26o), has M = 351): the same exact int8 slice arithmetic and the
// same CTA-pair MMA shape as k2_int8_pair.cuh (which see), re-planned around two facts:
//   * the byte image of T no longer fits shared memory (KP = 352: 6 slices x 66 KB = 405 KB, 203 KB per CTA of a pair),
//     so T is STREAMED: it stays in global memory (L2-resident, read 26 x per tile), and a producer warp in each CTA
//     copies the (slice, K chunk) block its next MMA needs -- one contiguous <= 5.6 KB piece of the pair image -- into
//     a ring of NSLOT shared-memory slots with cp.async.bulk (TMA), an mbarrier per slot counting the bytes;
//   * TMEM holds ONE full-width accumulator (KP columns) -- on this part tensor-core work and CUDA-core work of one SM
//     do not overlap anyway (they add up) -- and its remaining 512 - KP columns take as many
//     (slice, chunk) units of the X operand as fit (8 columns each); the other units live in shared memory.
// Roles (10 warps per CTA): warps 0..7 build the X slices and fold the accumulator exactly as in the pair kernel
// (EW = 2 epilogue warps per TMEM lane quadrant); warp 8 issues the MMAs in the leader CTA and, in the other CTA,
// forwards "my half of slot s has landed" to the leader (tcgen05.mma.cta_group::2 reads B from both CTAs, the leader's
// MMA lane can only wait on barriers in its own shared memory); warp 9 is the TMA producer.  "Slot consumed" comes back
// to both producers by a multicast tcgen05.commit per MMA.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <utility>
#include <algorithm>
#include <vector>
#include "../../openrdm_b200/csrc/k2_int8_pair.cuh"
#include "../../openrdm_b200/csrc/rdm_pack.h"

namespace ordm {
// The image for the streaming kernel (k2_int8_stream.cuh, 21 .. 26 active orbitals): per CTA rank c and slice l, the
// T blocks in MMA issue order.  K chunk kc (columns [32 kc, kp)) is one block, or two ([32 kc, 32 kc + 256) and the
// rest) when it is wider than the 256 columns one CTA-pair MMA takes; CTA c keeps rows [c w/2, (c+1) w/2) of a
// w-column block as a K-major no-swizzle operand of w/2 rows x 32 K bytes.  Same total size as pack_t_int8_pair.
inline void pack_t_int8_stream(int M, const std::vector<double>& Dt, int kp, std::vector<uint8_t>& img, int& t_exp) {
  std::vector<uint8_t> whole;
  pack_t_int8(M, Dt, kp, whole, t_exp);
  const int nslice = 6, maxn = 256;
  const size_t slice = (size_t)t_int8_slice_bytes(kp), half = slice / 2;
  img.assign((size_t)2 * nslice * half, 0);
  for (int l = 0; l < nslice; ++l) {
    size_t off = 0;
    for (int kc = 0; kc < kp / 32; ++kc) {
      const int rows = kp - 32 * kc;
      for (int n0 = 0; n0 < rows; n0 += maxn) {
        const int w = std::min(maxn, rows - n0), hw = w / 2;
        for (int n = 0; n < w; ++n)
          for (int k = 0; k < 32; ++k) {
            const int nt = n0 + n;   // row of the whole chunk
            const size_t from = (size_t)l * slice + (size_t)t_int8_chunk_off(kp, kc) + (size_t)((k >> 4) * (rows >> 3) + (nt >> 3)) * 128 + (nt & 7) * 16 + (k & 15);
            const int c = n / hw, nl = n % hw;
            const size_t to = ((size_t)c * nslice + l) * half + off + (size_t)((k >> 4) * (hw >> 3) + (nl >> 3)) * 128 + (nl & 7) * 16 + (k & 15);
            img[to] = whole[from];
          }
        off += (size_t)16 * w;
      }
    }
  }
}

}  // namespace ordm

namespace k2i {

constexpr int STREAM_MIN_ACT = 21, STREAM_MAX_ACT = 26;
#ifndef K2S_NSLOT
#define K2S_NSLOT 8
#endif
#ifndef K2S_BATCH
#define K2S_BATCH 4
#endif
constexpr int NSLOT = K2S_NSLOT;         // T ring depth (4 KB slots)
// Slots are handed back to the producers in batches of SLOT_BATCH, one tcgen05.commit per batch: a multicast commit of
// the CTA pair completes only about once per 1.1 k clk however many are queued (measured, profiles/r02*_k2s_*.log), so
// one commit per MMA paces the whole MMA stream at that rate
constexpr int SLOT_BATCH = K2S_BATCH;
static_assert(NSLOT % SLOT_BATCH == 0 && NSLOT / SLOT_BATCH >= 2, "the ring is a whole number (>= 2) of batches");
constexpr int STREAM_EW = 2;
constexpr int STREAM_THREADS = (4 * STREAM_EW + 2) * 32;

struct StreamGeom {
  __host__ __device__ static constexpr int nch(int kp) { return kp / 32; }
  __host__ __device__ static constexpr int units(int kp) { return NSLICE * nch(kp); }                  // (slice, chunk) units of X
  __host__ __device__ static constexpr int tmem_units(int kp) { return (512 - kp) / 8 < units(kp) ? (512 - kp) / 8 : units(kp); }
  __host__ __device__ static constexpr int smem_units(int kp) { return units(kp) - tmem_units(kp); }
  // unit index of (slice k, chunk kc); the HIGHEST indices live in TMEM
  __host__ __device__ static constexpr int unit(int kp, int k, int kc) { return k * nch(kp) + kc; }
  __host__ __device__ static constexpr bool in_tmem(int kp, int k, int kc) { return unit(kp, k, kc) >= smem_units(kp); }
  __host__ __device__ static constexpr int tmem_col(int kp, int k, int kc) { return kp + 8 * (unit(kp, k, kc) - smem_units(kp)); }
  __host__ __device__ static constexpr int smem_off(int kp, int k, int kc) { return unit(kp, k, kc) * TILE * 32; }   // 4 KB per unit
  // An MMA of the CTA pair takes N <= 256 columns: K chunk kc, which reaches columns [32 kc, kp), is issued as one
  // block, or as two ([32 kc, 32 kc + 256) and the rest) when it is wider.  A block of w columns is w / 2 rows of 32
  // bytes in each CTA's half of the image, blocks in issue order (ordm::pack_t_int8_stream).
  __host__ __device__ static constexpr int max_n() { return 256; }
  __host__ __device__ static constexpr int slot_bytes(int kp) { return 16 * (kp < max_n() ? kp : max_n()); }
  __host__ __device__ static constexpr int slice_half_bytes(int kp) { return Geom::slice_bytes(kp) / 2; }   // sum over blocks of 16 w
};

// shared-memory plan (dynamic): X units | T ring
struct StreamSmem {
  __host__ __device__ static size_t x_bytes(int kp) { return (size_t)StreamGeom::smem_units(kp) * TILE * 32; }
  __host__ __device__ static size_t ring_off(int kp) { return x_bytes(kp) < 2048 ? 2048 : x_bytes(kp); }   // the partial sums [EW][128] alias the X units
  __host__ __device__ static size_t bytes(int kp) { return ring_off(kp) + (size_t)NSLOT * StreamGeom::slot_bytes(kp); }
};

// once a barrier has timed out the kernel's results are void anyway: later waits give up at once instead of
// spinning 2^24 times each (a protocol bug must not turn into minutes of GPU time)
__device__ __forceinline__ bool wait_parity_ok(uint32_t bar, uint32_t parity, bool& ok) {
  if (ok) { ok = wait_parity(bar, parity); return ok; }
  for (int spin = 0; spin < 64; ++spin) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}

// four slice words (16 K bytes) of slice S, chunk KC, half HALF of this thread's row -> TMEM or shared memory
template <int KP, int KC, int HALF, int S>
__device__ __forceinline__ void store_slice(const BuildRegs& r, uint8_t* x_row, uint32_t tmem_row) {
  if constexpr (StreamGeom::in_tmem(KP, S, KC)) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tmem_row + (uint32_t)(StreamGeom::tmem_col(KP, S, KC) + 4 * HALF)),
                 "r"(r.w[S][0]), "r"(r.w[S][1]), "r"(r.w[S][2]), "r"(r.w[S][3]) : "memory");
  } else {   // K-major no-swizzle operand of 128 rows x 32 bytes: the two 16-byte K halves are 2 KB apart
    *(uint4*)(x_row + StreamGeom::smem_off(KP, S, KC) + HALF * (TILE / 8) * 128) = make_uint4(r.w[S][0], r.w[S][1], r.w[S][2], r.w[S][3]);
  }
}
template <int KP, int KC, int HALF, int... Ss>
__device__ __forceinline__ void store_slices(const BuildRegs& r, uint8_t* x_row, uint32_t tmem_row, std::integer_sequence<int, Ss...>) {
  (store_slice<KP, KC, HALF, Ss>(r, x_row, tmem_row), ...);
}

// Latency-critical single-lane waits of the T ring (producer, forwarder, MMA issuer): mbarrier.test_wait in a tight loop.
// try_wait may suspend the thread for a system-dependent time when the phase is not complete yet, and three such naps
// per ring slot set the pace of the MMA stream (measured: ~1 k clk per MMA with try_wait).
__device__ __forceinline__ bool spin_parity_ok(uint32_t bar, uint32_t parity, bool& ok) {
  const long limit = ok ? (1l << 26) : 64;
  for (long spin = 0; spin < limit; ++spin) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
  }
  ok = false;
  return false;
}

// ---- build: the X slices of pairs [16 U0, 16 U1) of this thread's point, each (slice, chunk) unit to where it lives
template <int NA, int I>
__device__ __forceinline__ void build_elem_stream(const double (&phi)[NA], const double (&ps)[NA], BuildRegs& r, uint8_t* x_row, uint32_t tmem_row) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32;
  if constexpr (I < M) {
    constexpr int t = pair_t_of(NA, I), u = pair_u_of(NA, I);
    const unsigned long long v = (unsigned long long)__double_as_longlong(fma(ps[t], phi[u], 6755399441055744.0 + 551911719040.0));
    r.lo[I & 3] = (uint32_t)v; r.hi[I & 3] = (uint32_t)(v >> 32);
  } else {
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
      constexpr int kb = I - 15;   // first K byte of the four words: 16 K bytes = half of chunk kb / 32
      store_slices<KP, kb / 32, (kb / 16) & 1>(r, x_row, tmem_row, std::make_integer_sequence<int, NSLICE>{});
    }
  }
}

template <int NA, int I0, int... Is>
__device__ __forceinline__ void build_seq_stream(const double (&phi)[NA], const double (&ps)[NA], BuildRegs& r, uint8_t* x_row, uint32_t tmem_row,
                                                 std::integer_sequence<int, Is...>) {
  (build_elem_stream<NA, I0 + Is>(phi, ps, r, x_row, tmem_row), ...);
}

template <int NA, int U0, int U1>
__device__ __forceinline__ void build_units_stream(const double (&phi)[NA], double scale, uint8_t* x_row, uint32_t tmem_row) {
  BuildRegs r;
  double ps[NA];
#pragma unroll
  for (int t = 0; t < NA; ++t) ps[t] = phi[t] * scale;
  build_seq_stream<NA, 16 * U0>(phi, ps, r, x_row, tmem_row, std::make_integer_sequence<int, 16 * (U1 - U0)>{});
}

#ifdef K2S_PROF   // lab only (tools/ozlab/k2s_lab): clock64 laps of CTA 0: [0] compute thread 0, [1] MMA lane, [2] producer lane
__device__ long long k2s_prof[3][8];
#define K2S_LAP(w, i) do { const long long t_ = clock64(); pf_[i] += t_ - tc_; tc_ = t_; } while (0)
#define K2S_T0 long long tc_ = clock64(), pf_[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define K2S_DUMP(w) do { if (blockIdx.x == 0) for (int i_ = 0; i_ < 8; ++i_) k2s_prof[w][i_] = pf_[i_]; } while (0)
#else
#define K2S_LAP(w, i) do { } while (0)
#define K2S_T0 do { } while (0)
#define K2S_DUMP(w) do { } while (0)
#endif

struct StreamCtx {
  uint32_t tmem, lane_addr, crank;
  uint32_t bar_acc_full, bar_acc_drain, drain_remote;   // accumulator complete (multicast commit) / drained by both CTAs (leader's copy)
  uint32_t bar_t_full, bar_t_empty, bar_t_peer;         // [NSLOT] each: my half landed / slot consumed (multicast commit) / peer's half landed (leader's copy)
  uint32_t peer_remote;                                 // the leader's bar_t_peer as seen from this CTA
  uint32_t ring;                                        // shared-memory address of the T ring
  uint8_t* x_row;
  double* s_part;
  uint64_t xdesc0;
  uint32_t idesc0;
  int p;
};

// ---- compute role (build + epilogue), one instantiation per unit range
template <int NA, int GMIN, int PART>
__device__ __forceinline__ void stream_compute_role(const Params& P, const StreamCtx& c, bool& ok, int& nflag) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, NUNITS = KP / 16, EW = STREAM_EW;
  constexpr int U0 = part_u0(NUNITS, EW, PART), U1 = part_u0(NUNITS, EW, PART + 1);
  constexpr int NGROUPS = 2 * (NSLICE - 1) - GMIN + 1;
  constexpr int ETHREADS = 4 * EW * 32;
  const int p = c.p;
  uint32_t n = 0;   // groups consumed so far
  const size_t ntiles = (P.npts + 2 * TILE - 1) / (2 * TILE);
#ifdef K2S_PROF
  const bool prof = threadIdx.x == 0;
#define K2S_CLAP(i) do { if (prof) K2S_LAP(0, i); } while (0)
#else
#define K2S_CLAP(i) do { } while (0)
#endif
  K2S_T0;
  for (size_t tile = blockIdx.x / 2; tile < ntiles; tile += gridDim.x / 2) {
    const size_t p0 = tile * 2 * TILE + (size_t)c.crank * TILE;
    double phi[NA];
    const float phif_unused[NA] = {};
    double s2 = 0.0;
    uint32_t mh = 0;
#pragma unroll
    for (int t = 0; t < NA; ++t) {
      phi[t] = (p0 + p < P.npts) ? __ldg(P.phi_act + (size_t)t * P.ld + p0 + p) : 0.0;
      mh = max(mh, (uint32_t)__double2hiint(phi[t]) & 0x7FFFFFFFu);
    }
    const int e_raw = min(max((int)(mh >> 20), 534), 1500);
    const double scale = __hiloint2double((3113 - 2 * e_raw + (FIX_BITS - 46)) << 20, 0);  // 2^(FIX_BITS - 2 em)
    K2S_CLAP(0);
    build_units_stream<NA, U0, U1>(phi, scale, c.x_row, c.tmem + c.lane_addr);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if constexpr (PART == 0) s2 = sum_squares<NA>(phi);
    {   // next pair tile's orbitals towards L2 while this one is contracted
      const size_t pn = p0 + (size_t)(gridDim.x / 2) * 2 * TILE + p;
      if (pn < P.npts) {
#pragma unroll
        for (int t = PART; t < NA; t += EW) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.phi_act + (size_t)t * P.ld + pn));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    K2S_CLAP(1);
    cluster_sync_all();   // both CTAs' operands are built
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    K2S_CLAP(2);
    double pi_acc = 0.0, pi_acc32 = 0.0;
#pragma unroll 1
    for (int gi = 0; gi < NGROUPS; ++gi, ++n) {
      const int G = group_at(gi, GMIN);
      wait_parity_ok(c.bar_acc_full, n & 1, ok);
      K2S_CLAP(3);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_row = c.tmem + c.lane_addr;
      const double wG = __hiloint2double((1023 + 8 * (G - GMIN)) << 20, 0);
      if (G > K2P_F32_MAX) pi_acc = fma(fold_group_units<NA, U0, U1, false, true>(phi, phif_unused, tmem_row, c.drain_remote), wG, pi_acc);
      else pi_acc32 = fma(fold_group_units<NA, U0, U1, true, true>(phi, phif_unused, tmem_row, c.drain_remote), wG, pi_acc32);
      K2S_CLAP(4);
    }
    pi_acc += pi_acc32;
    if constexpr (PART > 0) c.s_part[PART * TILE + p] = pi_acc;
    asm volatile("bar.sync 1, %0;" ::"r"(ETHREADS) : "memory");
    if constexpr (PART == 0) {
      double s = pi_acc;
#pragma unroll
      for (int k = 1; k < EW; ++k) s += c.s_part[k * TILE + p];
      const double f1 = __hiloint2double((2 * e_raw - 1067 - (FIX_BITS - 46)) << 20, 0);        // 2^(2 em - FIX_BITS)
      const double f2 = __hiloint2double((1023 + P.t_exp - FIX_BITS + 8 * GMIN + 1) << 20, 0);  // 2 * 2^(t_exp - FIX_BITS + 8 GMIN)
      const double pi_act = s * f1 * f2;
      if (p0 + p < P.npts) {
        const double pi_tot = P.pi_core ? pi_act + P.pi_core[p0 + p] : pi_act;
        P.pi[p0 + p] = pi_tot;
        if (guard_flag(s2, e_raw, fabs(pi_tot), P.guard_g1, P.guard_g2)) ++nflag;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(ETHREADS) : "memory");   // the partial sums alias the X units
    K2S_CLAP(5);
  }
#ifdef K2S_PROF
  if (prof) K2S_DUMP(0);
#endif
}

// the MMA sequence of one tile, shared by the issuer, the forwarder and the producers:
// f(first_of_group, last_of_group, k, l, kc, col0, w, byte offset of the block inside a slice half, overwrite)
template <int KP, int GMIN, class F>
__device__ __forceinline__ void for_each_mma(F&& f) {
  constexpr int NCH = KP / 32, NGROUPS = 2 * (NSLICE - 1) - GMIN + 1;
#pragma unroll 1
  for (int gi = 0; gi < NGROUPS; ++gi) {
    const int G = group_at(gi, GMIN);
    const int kmax = G < NSLICE - 1 ? G : NSLICE - 1, kmin = G > NSLICE - 1 ? G - (NSLICE - 1) : 0;
#pragma unroll 1
    for (int k = kmax; k >= kmin; --k) {
      int off = 0;
#pragma unroll 1
      for (int kc = 0; kc < NCH; ++kc) {
        const int w = KP - 32 * kc, w0 = w > StreamGeom::max_n() ? StreamGeom::max_n() : w;
        const bool over = k == kmax && kc == 0;   // the group's first touch of these accumulator columns
        f(over, k == kmin && kc == NCH - 1 && w0 == w, k, G - k, kc, 32 * kc, w0, off, over);
        off += 16 * w0;
        if (w0 < w) {
          f(false, k == kmin && kc == NCH - 1, k, G - k, kc, 32 * kc + w0, w - w0, off, over);
          off += 16 * (w - w0);
        }
      }
    }
  }
}

// A operand of unit (k, kc): compile-time placement, run-time indices -> small dispatch
template <int KP>
__device__ __forceinline__ void issue_one(const StreamCtx& c, int k, int kc, int col0, uint64_t bdesc, uint32_t idesc, uint32_t acc_flag) {
  const int unit = StreamGeom::unit(KP, k, kc);
  if (unit >= StreamGeom::smem_units(KP)) mma2_ts(c.tmem + (uint32_t)col0, c.tmem + (uint32_t)(KP + 8 * (unit - StreamGeom::smem_units(KP))), bdesc, idesc, acc_flag);
  else mma2_ss(c.tmem + (uint32_t)col0, c.xdesc0 + (uint64_t)((unit * TILE * 32) >> 4), bdesc, idesc, acc_flag);
}

template <int NA, int GMIN>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(128) k2i_ontop_stream(Params P) {
  constexpr int M = NA * (NA + 1) / 2, KP = (M + 31) / 32 * 32, EW = STREAM_EW;
  constexpr int ETHREADS = 4 * EW * 32;
  static_assert(NA >= STREAM_MIN_ACT && NA <= STREAM_MAX_ACT && KP <= 352, "active-space sizes of the streaming kernel");
  static_assert(NSLICE == 6 && (GMIN == 4 || GMIN == 5), "six slices, groups 10 .. GMIN");
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_s[2 + 3 * NSLOT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, part = warp >> 2;   // part 0, 1: compute; warp 8: MMA issuer / forwarder; warp 9: TMA producer
  StreamCtx c;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(c.crank));
  c.bar_acc_full = (uint32_t)__cvta_generic_to_shared(&bar_s[0]);
  c.bar_acc_drain = (uint32_t)__cvta_generic_to_shared(&bar_s[1]);
  c.bar_t_full = (uint32_t)__cvta_generic_to_shared(&bar_s[2]);
  c.bar_t_empty = (uint32_t)__cvta_generic_to_shared(&bar_s[2 + NSLOT]);
  c.bar_t_peer = (uint32_t)__cvta_generic_to_shared(&bar_s[2 + 2 * NSLOT]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.bar_acc_full));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c.bar_acc_drain), "r"(2 * ETHREADS));
    for (int s = 0; s < NSLOT; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.bar_t_full + 8 * s));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.bar_t_empty + 8 * s));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.bar_t_peer + 8 * s));
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
  c.x_row = smem + (c.p >> 3) * 128 + (c.p & 7) * 16;
  c.s_part = (double*)smem;
  c.xdesc0 = smem_desc((uint32_t)__cvta_generic_to_shared(smem), (TILE / 8) * 128, 128);
  c.ring = (uint32_t)__cvta_generic_to_shared(smem + StreamSmem::ring_off(KP));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(c.drain_remote) : "r"(c.bar_acc_drain), "r"(0));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(c.peer_remote) : "r"(c.bar_t_peer), "r"(0));
  bool ok = true;
  int nflag = 0;
  const size_t ntiles = (P.npts + 2 * TILE - 1) / (2 * TILE);
  size_t my_tiles = 0;
  for (size_t tile = blockIdx.x / 2; tile < ntiles; tile += gridDim.x / 2) ++my_tiles;

  if (part == 0) stream_compute_role<NA, GMIN, 0>(P, c, ok, nflag);
  else if (part == 1) stream_compute_role<NA, GMIN, 1>(P, c, ok, nflag);
  else if (warp == 4 * EW + 1) {
    // ---- TMA producer: the T block of every MMA, from this CTA's half of the image -> ring slot, NSLOT ahead
    const uint8_t* img = P.timg + (size_t)c.crank * (size_t)NSLICE * StreamGeom::slice_half_bytes(KP);
    uint32_t i = 0;
    K2S_T0;
    for (size_t t = 0; t < my_tiles; ++t) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      cluster_sync_all();   // (every thread of the cluster takes part in the per-tile barrier)
      if (lane == 0) {
        K2S_LAP(2, 0);
        for_each_mma<KP, GMIN>([&](bool, bool, int, int l, int, int, int w, int off, bool) {
          const uint32_t s = i % NSLOT, use = i / NSLOT;
          if (use > 0 && s % SLOT_BATCH == 0) spin_parity_ok(c.bar_t_empty + 8 * (s / SLOT_BATCH), (use - 1) & 1, ok);   // the MMAs that read this batch of slots last have completed
          K2S_LAP(2, 1);
          const uint32_t bytes = 16u * (uint32_t)w;                              // (w / 2) rows x 32 bytes
          const uint8_t* src = img + (size_t)l * StreamGeom::slice_half_bytes(KP) + off;
          const uint32_t bar = c.bar_t_full + 8 * s, dst = c.ring + s * (uint32_t)StreamGeom::slot_bytes(KP);
#ifndef K2S_LAB_NOTMA
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
#else   // lab switch (wrong results): the slot is declared full without copying anything: what does the ring cost without TMA?
          if (src == nullptr && dst == 0) ok = false;
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
#endif
          K2S_LAP(2, 2);
          ++i;
        });
      }
      __syncwarp();
    }
    if (lane == 0) K2S_DUMP(2);
  } else {
    // ---- warp 8: MMA issuer (leader CTA) / forwarder of "my half has landed" (the other CTA)
    uint32_t i = 0, n = 0;
    K2S_T0;
    for (size_t t = 0; t < my_tiles; ++t) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      cluster_sync_all();   // both CTAs' operands are built
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        K2S_LAP(1, 0);
        if (c.crank == 0) {
          for_each_mma<KP, GMIN>([&](bool first, bool last, int k, int, int kc, int col0, int w, int, bool over) {
            const uint32_t s = i % NSLOT, use = i / NSLOT;
            if (first && n > 0) wait_parity_ok(c.bar_acc_drain, (n - 1) & 1, ok);   // both CTAs have read the previous group out of the accumulator
            K2S_LAP(1, 1);
            spin_parity_ok(c.bar_t_full + 8 * s, use & 1, ok);
            K2S_LAP(1, 2);
            if (s % SLOT_BATCH == 0) spin_parity_ok(c.bar_t_peer + 8 * (s / SLOT_BATCH), use & 1, ok);   // the other CTA's halves of this batch of slots
            K2S_LAP(1, 3);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = c.idesc0 | ((uint32_t)(w >> 3) << 17);
            const uint64_t bdesc = smem_desc(c.ring + s * (uint32_t)StreamGeom::slot_bytes(KP), (uint32_t)(w / 2) * 16, 128);
#ifndef K2S_LAB_NOMMA
            issue_one<KP>(c, k, kc, col0, bdesc, idesc, over ? 0u : 1u);
#endif
            K2S_LAP(1, 4);
            if (s % SLOT_BATCH == SLOT_BATCH - 1) commit2(c.bar_t_empty + 8 * (s / SLOT_BATCH));
            if (last) { commit2(c.bar_acc_full); ++n; }
            K2S_LAP(1, 5);
            ++i;
          });
        } else {
          for_each_mma<KP, GMIN>([&](bool, bool, int, int, int, int, int, int, bool) {
            const uint32_t s = i % NSLOT, use = i / NSLOT;
            spin_parity_ok(c.bar_t_full + 8 * s, use & 1, ok);
            // one remote arrive per batch of slots.  No data crosses the CTA boundary with it -- the other SM's tensor core
            // reads its own shared memory, which the TMA completed before its local barrier flipped -- so it carries no
            // release semantics: a release.cluster arrive per MMA was measured to pace the whole MMA stream (~800 clk each)
            if (s % SLOT_BATCH == SLOT_BATCH - 1)
              asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(c.peer_remote + 8 * (s / SLOT_BATCH)) : "memory");
            ++i;
          });
        }
      }
      __syncwarp();
    }
    if (lane == 0 && c.crank == 0) K2S_DUMP(1);
  }
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
