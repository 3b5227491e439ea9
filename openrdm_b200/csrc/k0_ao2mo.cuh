// k0_ao2mo.cuh -- K0: the AO -> MO transform of the orbital values, the step directly in front of the path:
//     phi_mo(p, i) = sum_mu phi_ao(p, mu) C(mu, i)
// TransformPhiMatrixAOMO (/root/reference/src/libpsi4/mcpdft_solver.cc:619-635: two F_DGEMMs per irrep) /
// mo_values = np.matmul(ao_values, C_mo) (src/libpyscf/libpyscf.py:608).  A plain FP64 GEMM  C[m x N] = A[m x K] B[K x N],
// all column-major, with m = grid points of a batch (large), K = nao (hundreds), N = occupied MOs (~100): tall and
// skinny, A read once, B L2-resident.  FP64 has no tcgen05 kind; the tensor path is DMMA (mma.sync.m8n8k4.f64), as in K2.
//
// CTA = 128 rows x (8 NT) columns, 16 warps: warp (r, h) owns rows [16 r, 16 r + 16) x column half h (ceil(NT / 2) n-tiles):
// 2 x ceil(NT / 2) accumulator fragments per thread.  (Round 2's first version gave a warp all NT n-tiles: 170 registers,
// one 8-warp CTA per SM and two CTA barriers per chunk -- 21.2 TFLOP/s, 0.76 of cuBLAS; twice the warps at ~100 registers,
// a three-stage pipeline and one barrier per chunk hide the fragment-load and cp.async latencies.)  K is streamed in
// chunks of 16 through a three-stage cp.async pipeline: A chunk [16][128 + 8] (the + 8 makes the A-fragment loads
// conflict-free: 4 k-rows x 8 rows hit 32 distinct 8-byte banks), B chunk [8 NT][16 + 4] (ditto for B fragments).
// Rows beyond m and K beyond nao are zero-filled in shared memory; columns beyond N are not stored.
// A may also be given ROW-major (element (p, mu) at A[p * lda + mu], the layout numpy / data.h5 hold the AO values
// in): the chunk is then gathered with 8-byte loads straight into the same shared-memory layout.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace k0 {

constexpr int BM = 128, KC = 16;
// pipeline depth: three stages for wide CTAs (one per SM anyway), two for narrow ones (N <= 32 streams A at ~3 TB/s and
// wants five CTAs per SM rather than deeper buffers)
template <int NT> __host__ __device__ constexpr int stages_of() { return NT >= 8 ? 3 : 2; }
// WC = column groups of warps (8 warps each): 2 for wide CTAs (NT >= 8), 1 for narrow ones (NT = 4: two n-tiles per warp
// are too little work per barrier: 14.8 vs 16.3 TFLOP/s at N = 20)
template <int NT> __host__ __device__ constexpr int wc_of() { return NT >= 8 ? 2 : 1; }
template <int NT> __host__ __device__ constexpr int threads_of() { return 256 * wc_of<NT>(); }
constexpr int LDA_S = BM + 8, LDB_S = KC + 4;

struct Params {
  const double* A;   // m x K
  size_t lda;
  int a_row_major;   // 0: column-major (A[p + mu * lda]); 1: row-major (A[p * lda + mu])
  const double* B;   // K x N column-major, ldb = K
  double* C;         // m x N column-major
  size_t ldc;
  size_t m;
  int K, N;
};

template <int NT>
struct Smem {
  static constexpr int A_DOUBLES = KC * LDA_S, B_DOUBLES = 8 * NT * LDB_S;
  static constexpr size_t bytes = (size_t)stages_of<NT>() * (A_DOUBLES + B_DOUBLES) * sizeof(double);
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int sz = pred ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool pred) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(threads_of<NT>()) k0_ao2mo(const Params P) {
  constexpr int THREADS = threads_of<NT>(), WC = wc_of<NT>(), NSTAGE = stages_of<NT>();
  extern __shared__ __align__(16) double smem_d[];
  double *sA[NSTAGE], *sB[NSTAGE];
#pragma unroll
  for (int i = 0; i < NSTAGE; ++i) { sA[i] = smem_d + i * Smem<NT>::A_DOUBLES; sB[i] = smem_d + NSTAGE * Smem<NT>::A_DOUBLES + i * Smem<NT>::B_DOUBLES; }
  const int tid = threadIdx.x, warp = (tid >> 5) & 7, half = tid >> 8, lane = tid & 31;
  constexpr int NTW = (NT + WC - 1) / WC;   // n-tiles per warp
  const int b0 = half * NTW;
  const int g = lane >> 2, q = lane & 3;   // fragment coordinates: A row / B column g, k index q; C row g, columns 2q, 2q + 1
  const size_t row0 = (size_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * 8 * NT;
  const int nchunks = (P.K + KC - 1) / KC;

  auto load_chunk = [&](int kc, int buf) {
    const int k0 = kc * KC;
    if (!P.a_row_major) {   // 16 columns x 128 rows: 16-byte pieces, two rows each
      for (int i = tid; i < KC * (BM / 2); i += THREADS) {
        const int k = i / (BM / 2), r = (i % (BM / 2)) * 2;
        const bool ok = (k0 + k) < P.K && (row0 + r) < P.m;   // m and lda are even here (host pads), so a pair is in or out together
        cp_async16(sA[buf] + k * LDA_S + r, P.A + (size_t)(ok ? k0 + k : 0) * P.lda + (ok ? row0 + r : 0), ok);
      }
    } else {                // row-major source: 128 rows x 16 consecutive k: 8-byte gathers, transposed into the same layout
      for (int i = tid; i < KC * BM; i += THREADS) {
        const int r = i / KC, k = i % KC;
        const bool ok = (k0 + k) < P.K && (row0 + r) < P.m;
        cp_async8(sA[buf] + k * LDA_S + r, P.A + (ok ? (row0 + r) * P.lda + (size_t)(k0 + k) : 0), ok);
      }
    }
    for (int i = tid; i < 8 * NT * (KC / 2); i += THREADS) {   // 8 NT columns x 16 k: 16-byte pieces
      const int n = i / (KC / 2), k = (i % (KC / 2)) * 2;
      // K may be odd: a pair that straddles the end is fetched element-wise
      const bool col_ok = (n0 + n) < P.N;
      const bool ok2 = col_ok && (k0 + k + 1) < P.K && (((size_t)(n0 + n) * P.K + k0 + k) & 1) == 0;
      if (ok2) cp_async16(sB[buf] + n * LDB_S + k, P.B + (size_t)(n0 + n) * P.K + k0 + k, true);
      else {
        const bool o0 = col_ok && (k0 + k) < P.K, o1 = col_ok && (k0 + k + 1) < P.K;
        cp_async8(sB[buf] + n * LDB_S + k, P.B + (o0 ? (size_t)(n0 + n) * P.K + k0 + k : 0), o0);
        cp_async8(sB[buf] + n * LDB_S + k + 1, P.B + (o1 ? (size_t)(n0 + n) * P.K + k0 + k + 1 : 0), o1);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  double acc[2][NTW][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < NTW; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  // (an empty commit keeps the group count in step when there is nothing left to load)
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nchunks) load_chunk(s, s);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kc = 0; kc < nchunks; ++kc) {
    const int buf = kc % NSTAGE;
    asm volatile("cp.async.wait_group %0;" ::"n"(NSTAGE - 2) : "memory");   // chunk kc has landed (this thread's part)
    __syncthreads();   // ... everybody's part; and everybody is done with chunk kc - 1, whose buffer is refilled next
    if (kc + NSTAGE - 1 < nchunks) load_chunk(kc + NSTAGE - 1, (kc + NSTAGE - 1) % NSTAGE);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const double* a_s = sA[buf] + 16 * warp + g;
    const double* b_s = sB[buf] + (size_t)(b0 * 8 + g) * LDB_S;
#pragma unroll
    for (int ks = 0; ks < KC; ks += 4) {
      const double a0 = a_s[(ks + q) * LDA_S], a1 = a_s[(ks + q) * LDA_S + 8];
#pragma unroll
      for (int b = 0; b < NTW; ++b) {
        if (b0 + b < NT) {
          const double bf = b_s[b * 8 * LDB_S + ks + q];
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(acc[0][b][0]), "+d"(acc[0][b][1]) : "d"(a0), "d"(bf));
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(acc[1][b][0]), "+d"(acc[1][b][1]) : "d"(a1), "d"(bf));
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const size_t r = row0 + 16 * warp + 8 * a + g;
    if (r >= P.m) continue;
#pragma unroll
    for (int b = 0; b < NTW; ++b) {
      if (b0 + b >= NT) continue;
      const int c = n0 + 8 * (b0 + b) + 2 * q;
      if (c < P.N) P.C[(size_t)c * P.ldc + r] = acc[a][b][0];
      if (c + 1 < P.N) P.C[(size_t)(c + 1) * P.ldc + r] = acc[a][b][1];
    }
  }
}

}  // namespace k0
