// i8mma_ts_bench.cu -- (A operand from TMEM, written by tcgen05.st; ragged-N check) lab microbenchmark (not product code): what does tcgen05.mma.kind::i8 (int8 x int8 -> int32,
// the only tensor-core path on sm_100a that can carry FP64-grade contractions, via Ozaki slicing)
// sustain from one CTA per SM with both operands resident in shared memory, and is the result exact?
//
//   A = point-major operand, 128 rows x 224 K-bytes per slice, K-major "interleaved" (no swizzle) canonical layout
//   B = 224 rows (N) x 224 K-bytes per slice, same layout; D[m][n] = sum_k A[m][k] * B[n][k] in TMEM (int32)
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o i8mma_bench i8mma_bench.cu
// run:   timeout 60 ./i8mma_bench [N=224] [pairs=4] [iters=200]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int MROWS = 128, KB = 224, NMAX = 224;
constexpr int A_SLICES = 2, B_SLICES = 3;
constexpr int A_BYTES = MROWS * KB, B_BYTES = NMAX * KB;
constexpr int SMEM_BYTES = A_SLICES * A_BYTES + B_SLICES * B_BYTES + 1024;

// canonical K-major no-swizzle layout: 8-row x 16-byte core matrices (128 contiguous bytes each);
// core (rb, kb) of an R-row operand sits at (kb * R/8 + rb) * 128
__host__ __device__ inline size_t canon(int r, int k, int R) { return ((size_t)(k >> 4) * (R >> 3) + (r >> 3)) * 128 + (r & 7) * 16 + (k & 15); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool wait_parity(uint32_t bar, uint32_t parity) {
  for (long spin = 0; spin < (1l << 26); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

struct Args {
  const int8_t* a; const int8_t* b; int32_t* out; long long* clocks; int* fail;
  long long* ldlat;   // ldlat[2]: tcgen05.ld round trip (clk) of warp 1 while the MMAs run / when the tensor pipe is idle
  int n, pairs, iters, ragged;
};

__global__ void __launch_bounds__(128, 1) i8mma_kernel(Args g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bars[2];
  uint8_t* sa = smem;
  uint8_t* sb = smem + A_SLICES * A_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < B_SLICES * B_BYTES / 16; i += 128) ((int4*)sb)[i] = ((const int4*)g.b)[i];
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bars[0]), bar1 = (uint32_t)__cvta_generic_to_shared(&bars[1]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // A slices: lane = row (this thread's row is tid), K bytes 4j..4j+3 of slice sl in column 256 + 56*sl + j
  for (int sl = 0; sl < A_SLICES; ++sl) {
    const uint32_t* row = (const uint32_t*)(g.a + ((size_t)sl * MROWS + tid) * KB);
    for (int j0 = 0; j0 < KB / 4; j0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = row[j0 + j];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(256 + 56 * sl + j0);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                   "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.n >> 3) << 17) | ((uint32_t)(MROWS >> 4) << 24);
  const uint32_t sa_addr = (uint32_t)__cvta_generic_to_shared(sa), sb_addr = (uint32_t)__cvta_generic_to_shared(sb);
  constexpr uint32_t A_LBO = (MROWS / 8) * 128, B_LBO = (NMAX / 8) * 128, SBO = 128;
  bool ok = true;
  long long t0 = 0, t1 = 0;
  if (warp == 1 && g.ldlat) {   // TMEM round trips from another lane quadrant, unused columns, while warp 0's lane issues the MMAs
    long long acc = 0; const int nld = g.iters * g.pairs * 2;
    for (int i = 0; i < nld; ++i) {
      uint32_t v[16];
      const long long c0 = clock64();
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tmem + ((uint32_t)32 << 16) + 400u));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += clock64() - c0;
      if (v[0] == 0x7fffffffu) acc += 1;
    }
    if (tid == 32 && blockIdx.x == 0) g.ldlat[0] = acc / nld;
  }
  if (tid == 0) {
    t0 = clock64();
    for (int it = 0; it < g.iters; ++it) {
      // one accumulator: sum over `pairs` slice pairs (A_{p%2} from TMEM . B_{p%3}), 7 K-steps of 32 bytes each;
      // ragged: K-step ks only feeds columns [32 ks, n) (block-upper-triangular B), i.e. rows 32 ks.. of B
      for (int p = 0; p < g.pairs; ++p) {
        const uint32_t a_t = tmem + 256 + 56 * (p % A_SLICES), bbase = sb_addr + (p % B_SLICES) * B_BYTES;
        if (g.ragged == 2) {   // column halves: right half [112, n) from every K-step, then left half [32 ks, 112) from K-steps 0..3
          for (int ks = 0; ks < KB / 32; ++ks) {
            const int c0 = max(112, 32 * ks);
            const uint32_t id = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((g.n - c0) >> 3) << 17);
            mma_i8_ts(tmem + c0, a_t + 8 * ks, smem_desc(bbase + ks * 2 * B_LBO + (c0 >> 3) * 128, B_LBO, SBO), id, (it | p | ks) ? 1u : 0u);
          }
          for (int ks = 0; ks < 4; ++ks) {
            const int c0 = 32 * ks;
            const uint32_t id = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((112 - c0) >> 3) << 17);
            mma_i8_ts(tmem + c0, a_t + 8 * ks, smem_desc(bbase + ks * 2 * B_LBO + (c0 >> 3) * 128, B_LBO, SBO), id, (it | p | ks) ? 1u : 0u);
          }
          continue;
        }
#pragma unroll
        for (int ks = 0; ks < KB / 32; ++ks) {
          const int c0 = g.ragged ? 32 * ks : 0;
          if (c0 >= g.n) continue;
          const uint32_t id = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((g.n - c0) >> 3) << 17);
          mma_i8_ts(tmem + c0, a_t + 8 * ks, smem_desc(bbase + ks * 2 * B_LBO + (c0 >> 3) * 128, B_LBO, SBO), id, (it | p | ks) ? 1u : 0u);
        }
      }
    }
    commit(bar0);
    ok = ok && wait_parity(bar0, 0);
    t1 = clock64();
    if (!ok) atomicAdd(g.fail, 1);
    g.clocks[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 1 && g.ldlat) {   // the same with the tensor pipe idle
    long long acc = 0;
    for (int i = 0; i < 1000; ++i) {
      uint32_t v[16];
      const long long c0 = clock64();
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tmem + ((uint32_t)32 << 16) + 400u));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += clock64() - c0;
      if (v[0] == 0x7fffffffu) acc += 1;
    }
    if (tid == 32 && blockIdx.x == 0) g.ldlat[1] = acc / 1000;
  }
  __syncthreads();
  // epilogue: the last two chains' accumulators -> global (block 0 only), lanes 32*warp.. of each
  if (blockIdx.x == 0) {
    for (int acc = 0; acc < 1; ++acc) {
      for (int c0 = 0; c0 < g.n; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 256 + c0);  // acc = 0
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) g.out[((size_t)acc * MROWS + tid) * NMAX + c0 + j] = (int32_t)v[j];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 224, pairs = argc > 2 ? atoi(argv[2]) : 4, iters = argc > 3 ? atoi(argv[3]) : 200, ragged = argc > 4 ? atoi(argv[4]) : 0;
  if (n % 16 || n < 16 || n > NMAX) { printf("N must be a multiple of 16 in [16, %d]\n", NMAX); return 1; }
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  std::vector<int8_t> A((size_t)A_SLICES * MROWS * KB), B((size_t)B_SLICES * NMAX * KB), ca(A.size()), cb(B.size());
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int8_t)((int)((s >> 16) % 129u) - 64); };
  for (auto& x : A) x = rnd();
  for (auto& x : B) x = rnd();
  for (int sl = 0; sl < A_SLICES; ++sl) for (int r = 0; r < MROWS; ++r) for (int k = 0; k < KB; ++k)
    ca[(size_t)sl * A_BYTES + canon(r, k, MROWS)] = A[((size_t)sl * MROWS + r) * KB + k];
  for (int sl = 0; sl < B_SLICES; ++sl) for (int r = 0; r < NMAX; ++r) for (int k = 0; k < KB; ++k)
    cb[(size_t)sl * B_BYTES + canon(r, k, NMAX)] = B[((size_t)sl * NMAX + r) * KB + k];
  int8_t *da, *db; int32_t* dout; long long* dclk; int* dfail;
  CK(cudaMalloc(&da, A.size())); CK(cudaMalloc(&db, cb.size())); CK(cudaMalloc(&dout, sizeof(int32_t) * 2 * MROWS * NMAX));
  CK(cudaMalloc(&dclk, sizeof(long long) * nsm)); CK(cudaMalloc(&dfail, sizeof(int)));
  CK(cudaMemcpy(da, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, cb.data(), cb.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, sizeof(int32_t) * 2 * MROWS * NMAX)); CK(cudaMemset(dfail, 0, sizeof(int)));
  CK(cudaFuncSetAttribute(i8mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  long long* dlat; CK(cudaMalloc(&dlat, 16)); CK(cudaMemset(dlat, 0, 16));
  Args g{da, db, dout, dclk, dfail, dlat, n, pairs, 1, ragged};
  Args gt = g; gt.iters = iters;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  i8mma_kernel<<<nsm, 128, SMEM_BYTES>>>(g); CK(cudaDeviceSynchronize());
  std::vector<int32_t> out((size_t)2 * MROWS * NMAX);
  CK(cudaMemcpy(out.data(), dout, sizeof(int32_t) * out.size(), cudaMemcpyDeviceToHost));
  CK(cudaEventRecord(e0)); i8mma_kernel<<<nsm, 128, SMEM_BYTES>>>(gt); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> clk(nsm); int fail;
  CK(cudaMemcpy(clk.data(), dclk, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&fail, dfail, sizeof(int), cudaMemcpyDeviceToHost));
  // check: every chain computes sum_p A_{p%2} . B_{p%3}^T
  long bad = 0;
  for (int m = 0; m < MROWS; ++m) for (int j = 0; j < n; ++j) {
    long long want = 0;
    for (int p = 0; p < pairs; ++p) {
      const int8_t* ar = &A[((size_t)(p % A_SLICES) * MROWS + m) * KB]; const int8_t* br = &B[((size_t)(p % B_SLICES) * NMAX + j) * KB];
      for (int k = 0; k < (ragged ? (j / 32 + 1) * 32 : KB); ++k) want += (int)ar[k] * (int)br[k];
    }
    for (int acc = 0; acc < 1; ++acc) if (out[((size_t)acc * MROWS + m) * NMAX + j] != (int32_t)want) { if (bad < 5) printf("mismatch acc %d m %d n %d: got %d want %lld\n", acc, m, j, out[((size_t)acc * MROWS + m) * NMAX + j], want); ++bad; }
  }
  double cmean = 0; for (auto c : clk) cmean += (double)c; cmean /= nsm;
  double mma_per_cta = 0, mac = 0;
  for (int ks = 0; ks < KB / 32; ++ks) { const int c0 = ragged ? 32 * ks : 0; if (c0 < n) { mma_per_cta += (double)iters * pairs * (ragged == 2 && ks < 4 ? 2 : 1); mac += (double)iters * pairs * MROWS * (n - c0) * 32.0; } }
  printf("TS%s N=%d pairs=%d iters=%d: %s, barrier timeouts %d; %.1f clk per MMA (ideal %.1f at 8192 MAC/clk/SM); kernel %.3f ms -> %.1f TOP/s int8 on %d SMs\n",
         ragged ? " ragged" : "", n, pairs, iters, bad ? "MISMATCH" : "exact", fail, cmean / mma_per_cta, mac / mma_per_cta / 8192.0, ms, 2.0 * mac * nsm / (ms * 1e-3) / 1e12, nsm);
  { long long lat[2]; CK(cudaMemcpy(lat, dlat, 16, cudaMemcpyDeviceToHost));
    printf("tcgen05.ld.32x32b.x16 + wait::ld round trip from another warp: %lld clk while the MMAs run, %lld clk with the tensor pipe idle\n", lat[0], lat[1]); }
  return bad || fail ? 2 : 0;
}
