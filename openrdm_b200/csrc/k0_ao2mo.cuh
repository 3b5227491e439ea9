// k0_ao2mo.cuh -- K0: the AO -> MO transform of the orbital values, the step directly in front of the path:
//     phi_mo(p, i) = sum_mu phi_ao(p, mu) C(mu, i)
// TransformPhiMatrixAOMO (/root/reference/src/libpsi4/mcpdft_solver.cc:619-635: two F_DGEMMs per irrep) /
// mo_values = np.matmul(ao_values, C_mo) (src/libpyscf/libpyscf.py:608).  A plain FP64 GEMM  C[m x N] = A[m x K] B[K x N],
// all column-major, with m = grid points of a batch (large), K = nao (hundreds), N = occupied MOs (~100): tall and
// skinny, A read once, B L2-resident.  FP64 has no tcgen05 kind; the tensor path is DMMA (mma.sync.m8n8k4.f64), as in K2.
//
// CTA = 128 rows x (8 NT) columns, 8 warps, warp w owns rows [16 w, 16 w + 16) x all 8 NT columns: 2 x NT accumulator
// fragments (4 NT doubles per thread), so a k-step of 4 costs 2 + NT fragment loads for 2 NT DMMAs.  K is streamed in
// chunks of 16 through a two-stage cp.async pipeline: A chunk [16][128 + 8] (the + 8 makes the A-fragment loads
// conflict-free: 4 k-rows x 8 rows hit 32 distinct 8-byte banks), B chunk [8 NT][16 + 4] (ditto for B fragments).
// Rows beyond m and K beyond nao are zero-filled in shared memory; columns beyond N are not stored.
// A may also be given ROW-major (element (p, mu) at A[p * lda + mu], the layout numpy / data.h5 hold the AO values
// in): the chunk is then gathered with 8-byte loads straight into the same shared-memory layout.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace k0 {

constexpr int BM = 128, KC = 16, THREADS = 256;
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
  static constexpr size_t bytes = (size_t)2 * (A_DOUBLES + B_DOUBLES) * sizeof(double);
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
__global__ void __launch_bounds__(THREADS) k0_ao2mo(const Params P) {
  extern __shared__ __align__(16) double smem_d[];
  double* sA[2] = {smem_d, smem_d + Smem<NT>::A_DOUBLES};
  double* sB[2] = {smem_d + 2 * Smem<NT>::A_DOUBLES, smem_d + 2 * Smem<NT>::A_DOUBLES + Smem<NT>::B_DOUBLES};
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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

  double acc[2][NT][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  load_chunk(0, 0);
  for (int kc = 0; kc < nchunks; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < nchunks) {
      load_chunk(kc + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const double* a_s = sA[buf] + 16 * warp + g;
    const double* b_s = sB[buf] + g * LDB_S;
#pragma unroll
    for (int ks = 0; ks < KC; ks += 4) {
      const double a0 = a_s[(ks + q) * LDA_S], a1 = a_s[(ks + q) * LDA_S + 8];
#pragma unroll
      for (int b = 0; b < NT; ++b) {
        const double bf = b_s[b * 8 * LDB_S + ks + q];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(acc[0][b][0]), "+d"(acc[0][b][1]) : "d"(a0), "d"(bf));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(acc[1][b][0]), "+d"(acc[1][b][1]) : "d"(a1), "d"(bf));
      }
    }
    __syncthreads();   // the buffer is refilled two iterations from now, by the load issued at the top of the next one
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const size_t r = row0 + 16 * warp + 8 * a + g;
    if (r >= P.m) continue;
#pragma unroll
    for (int b = 0; b < NT; ++b) {
      const int c = n0 + 8 * b + 2 * q;
      if (c < P.N) P.C[(size_t)c * P.ldc + r] = acc[a][b][0];
      if (c + 1 < P.N) P.C[(size_t)(c + 1) * P.ldc + r] = acc[a][b][1];
    }
  }
}

}  // namespace k0
