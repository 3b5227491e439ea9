// i8mma_2cta_bench.cu -- lab microbenchmark (not product code), CTA-pair variant (cta_group::2, M = 256, each CTA
// holds its own 128 rows of A and HALF of the B rows): what does tcgen05.mma.kind::i8 (int8 x int8 -> int32,
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
constexpr int BHALF = NMAX / 2;   // B rows held by each CTA of the pair
constexpr int A_BYTES = MROWS * KB, B_BYTES = BHALF * KB;
constexpr int SMEM_BYTES = A_SLICES * A_BYTES + B_SLICES * B_BYTES + 1024;

// canonical K-major no-swizzle layout: 8-row x 16-byte core matrices (128 contiguous bytes each);
// core (rb, kb) of an R-row operand sits at (kb * R/8 + rb) * 128
__host__ __device__ inline size_t canon(int r, int k, int R) { return ((size_t)(k >> 4) * (R >> 3) + (r >> 3)) * 128 + (r & 7) * 16 + (k & 15); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
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
  int n, pairs, iters;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) i8mma_kernel(Args g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bars[2];
  uint8_t* sa = smem;
  uint8_t* sb = smem + A_SLICES * A_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  for (int i = tid; i < A_SLICES * A_BYTES / 16; i += 128) ((int4*)sa)[i] = ((const int4*)(g.a + (size_t)crank * A_SLICES * A_BYTES))[i];
  for (int i = tid; i < B_SLICES * B_BYTES / 16; i += 128) ((int4*)sb)[i] = ((const int4*)(g.b + (size_t)crank * B_SLICES * B_BYTES))[i];
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bars[0]), bar1 = (uint32_t)__cvta_generic_to_shared(&bars[1]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // both CTAs' operands and barriers are ready
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.n >> 3) << 17) | ((uint32_t)((2 * MROWS) >> 4) << 24);   // M = 256 over the pair
  const uint32_t sa_addr = (uint32_t)__cvta_generic_to_shared(sa), sb_addr = (uint32_t)__cvta_generic_to_shared(sb);
  constexpr uint32_t A_LBO = (MROWS / 8) * 128, B_LBO = (BHALF / 8) * 128, SBO = 128;
  bool ok = true;
  long long t0 = 0, t1 = 0;
  if (tid == 0 && crank == 0) {
    t0 = clock64();
    for (int it = 0; it < g.iters; ++it) {
      // accumulator (it & 1): sum over `pairs` slice pairs (A_{p%2} . B_{p%3}), 7 K-steps of 32 bytes each
      const uint32_t d = tmem + (uint32_t)(it & 1) * 256;
      if (it >= 2) ok = ok && wait_parity((it & 1) ? bar1 : bar0, ((it - 2) >> 1) & 1);  // the accumulator's previous chain is done
      for (int p = 0; p < g.pairs; ++p) {
        const uint32_t abase = sa_addr + (p % A_SLICES) * A_BYTES, bbase = sb_addr + (p % B_SLICES) * B_BYTES;
#pragma unroll
        for (int ks = 0; ks < KB / 32; ++ks)
          mma_i8(d, smem_desc(abase + ks * 2 * A_LBO, A_LBO, SBO), smem_desc(bbase + ks * 2 * B_LBO, B_LBO, SBO), idesc, (p | ks) ? 1u : 0u);
      }
      commit((it & 1) ? bar1 : bar0);
    }
    // drain both
    for (int it = max(0, g.iters - 2); it < g.iters; ++it) ok = ok && wait_parity((it & 1) ? bar1 : bar0, (it >> 1) & 1);
    t1 = clock64();
    if (!ok) atomicAdd(g.fail, 1);
    g.clocks[blockIdx.x] = t1 - t0;
  } else if (tid == 0) {   // the other CTA of the pair: the multicast commits flip its barriers too
    for (int it = 0; it < g.iters; ++it) ok = ok && wait_parity((it & 1) ? bar1 : bar0, (it >> 1) & 1);
    if (!ok) atomicAdd(g.fail, 1);
    g.clocks[blockIdx.x] = 0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: the last two chains' accumulators -> global (block 0 only), lanes 32*warp.. of each
  if (blockIdx.x < 2) {
    for (int acc = 0; acc < 2; ++acc) {
      for (int c0 = 0; c0 < g.n; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 256 + c0);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) g.out[((size_t)acc * 2 * MROWS + crank * MROWS + tid) * NMAX + c0 + j] = (int32_t)v[j];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 224, pairs = argc > 2 ? atoi(argv[2]) : 4, iters = argc > 3 ? atoi(argv[3]) : 200;
  if (n != NMAX) { printf("this variant runs N = %d only\n", NMAX); return 1; }
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount / 2 * 2;
  std::vector<int8_t> A((size_t)A_SLICES * 2 * MROWS * KB), B((size_t)B_SLICES * NMAX * KB), ca(A.size()), cb(B.size());
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int8_t)((int)((s >> 16) % 129u) - 64); };
  for (auto& x : A) x = rnd();
  for (auto& x : B) x = rnd();
  for (int c = 0; c < 2; ++c) for (int sl = 0; sl < A_SLICES; ++sl) for (int r = 0; r < MROWS; ++r) for (int k = 0; k < KB; ++k)
    ca[((size_t)c * A_SLICES + sl) * A_BYTES + canon(r, k, MROWS)] = A[((size_t)sl * 2 * MROWS + c * MROWS + r) * KB + k];
  for (int c = 0; c < 2; ++c) for (int sl = 0; sl < B_SLICES; ++sl) for (int r = 0; r < BHALF; ++r) for (int k = 0; k < KB; ++k)
    cb[((size_t)c * B_SLICES + sl) * B_BYTES + canon(r, k, BHALF)] = B[((size_t)sl * NMAX + c * BHALF + r) * KB + k];
  int8_t *da, *db; int32_t* dout; long long* dclk; int* dfail;
  CK(cudaMalloc(&da, ca.size())); CK(cudaMalloc(&db, cb.size())); CK(cudaMalloc(&dout, sizeof(int32_t) * 4 * MROWS * NMAX));
  CK(cudaMalloc(&dclk, sizeof(long long) * nsm)); CK(cudaMalloc(&dfail, sizeof(int)));
  CK(cudaMemcpy(da, ca.data(), ca.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, cb.data(), cb.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, sizeof(int32_t) * 4 * MROWS * NMAX)); CK(cudaMemset(dfail, 0, sizeof(int)));
  CK(cudaFuncSetAttribute(i8mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  Args g{da, db, dout, dclk, dfail, n, pairs, iters};
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  i8mma_kernel<<<nsm, 128, SMEM_BYTES>>>(g); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); i8mma_kernel<<<nsm, 128, SMEM_BYTES>>>(g); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> clk(nsm); std::vector<int32_t> out((size_t)4 * MROWS * NMAX); int fail;
  CK(cudaMemcpy(clk.data(), dclk, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out.data(), dout, sizeof(int32_t) * out.size(), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&fail, dfail, sizeof(int), cudaMemcpyDeviceToHost));
  // check: every chain computes sum_p A_{p%2} . B_{p%3}^T
  long bad = 0;
  for (int m = 0; m < 2 * MROWS; ++m) for (int j = 0; j < n; ++j) {
    long long want = 0;
    for (int p = 0; p < pairs; ++p) {
      const int8_t* ar = &A[((size_t)(p % A_SLICES) * 2 * MROWS + m) * KB]; const int8_t* br = &B[((size_t)(p % B_SLICES) * NMAX + j) * KB];
      for (int k = 0; k < KB; ++k) want += (int)ar[k] * (int)br[k];
    }
    for (int acc = 0; acc < 2; ++acc) if (out[((size_t)acc * 2 * MROWS + m) * NMAX + j] != (int32_t)want) { if (bad < 5) printf("mismatch acc %d m %d n %d: got %d want %lld\n", acc, m, j, out[((size_t)acc * 2 * MROWS + m) * NMAX + j], want); ++bad; }
  }
  double cmean = 0; for (auto c : clk) cmean += (double)c; cmean /= (nsm / 2);   // leaders only
  const double mma_per_cta = (double)iters * pairs * (KB / 32), mac = mma_per_cta * MROWS * n * 32.0;   // per CTA: its 128 rows
  printf("2-CTA M=256 N=%d pairs=%d iters=%d: %s, barrier timeouts %d; %.1f clk per MMA (ideal %.1f at 8192 MAC/clk/SM); kernel %.3f ms -> %.1f TOP/s int8 on %d SMs\n",
         n, pairs, iters, bad ? "MISMATCH" : "exact", fail, cmean / mma_per_cta, MROWS * n * 32.0 / 8192.0, ms, 2.0 * mac * nsm / (ms * 1e-3) / 1e12, nsm);
  return bad || fail ? 2 : 0;
}
