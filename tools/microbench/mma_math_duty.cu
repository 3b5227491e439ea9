// mma_math_duty.cu -- lab microbenchmark (not product code): how fast do CUDA-core instructions issue on an SM whose
// tensor core runs tcgen05.mma.kind::i8 part of the time?  k2_int8_pair.cuh's epilogue (int -> double bit trick, DADD,
// DFMA) runs ~3x slower beside the MMAs than alone; the question here is whether that throttle follows the tensor
// pipe's activity cycle by cycle (then a schedule that keeps the tensor pipe busy and lets the epilogue trail would win)
// or lingers after a burst of MMAs (then only the sums count).
//
// One CTA per SM: warps 0..7 run a math loop (two warps per sub-partition, like the pair kernel's epilogue), warp 8's
// lane 0 issues bursts of `burst` MMAs (128 x 224 x 32 each, ~112 clk), waits for their completion, then idles `idle`
// clocks; repeated until `total` clocks have passed.  Output: math instructions per clock per sub-partition, and the
// tensor duty cycle (time inside bursts / total).
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o mma_math_duty mma_math_duty.cu
// run:   timeout 120 ./mma_math_duty
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int MROWS = 128, KB = 224, N = 224;
constexpr int A_BYTES = MROWS * KB, B_BYTES = N * KB;
constexpr int SMEM_BYTES = A_BYTES + B_BYTES + 1024;
constexpr int MATH_WARPS = 8, THREADS = (MATH_WARPS + 1) * 32;

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
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
  long long* out;   // per CTA: [0] total clk, [1] clk inside bursts, [2..2+MATH_WARPS) math iterations per warp, [10] MMAs issued
  double* sink;
  int burst, idle, kind, skip0, mma_warp;   // skip0: no math warps on the MMA warp's sub-partition
  long long total;
};

// kind 0: the pair kernel's FP64 fold per element (LOP3, DADD, DFMA on 8 chains); 1: FFMA; 2: IMAD; 3: DFMA; 4: IMAD.WIDE;
// 5: I2FP + FFMA; 6: DADD; 7: I2F.F64.S32; 8: DMUL
constexpr int NKIND = 9;
constexpr int INSTR_PER_ITER[NKIND] = {3 * 8, 16, 16, 8, 8, 2 * 16, 8, 8, 8};

__global__ void __launch_bounds__(THREADS, 1) duty_kernel(Args g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (A_BYTES + B_BYTES) / 16; i += THREADS) ((int4*)smem)[i] = make_int4(0x01020304 * (i & 3), 0x01010101, 0x02020202, 0x03030303);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bar);
  if (tid == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
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
  if (warp == g.mma_warp) {
    if ((tid & 31) == 0) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(MROWS >> 4) << 24);
      const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem), sb = sa + A_BYTES;
      constexpr uint32_t A_LBO = (MROWS / 8) * 128, B_LBO = (N / 8) * 128, SBO = 128;
      long long in_burst = 0, nmma = 0;
      uint32_t phase = 0;
      const long long t0 = clock64();
      while (clock64() - t0 < g.total) {
        if (g.burst > 0) {
          const long long b0 = clock64();
          for (int i = 0; i < g.burst; ++i) {
            const int ks = i % (KB / 32);
            mma_i8(tmem + (uint32_t)((i & 1) * 256), smem_desc(sa + ks * 2 * A_LBO, A_LBO, SBO), smem_desc(sb + ks * 2 * B_LBO, B_LBO, SBO), idesc, i >= 2);
          }
          commit(bar0);
          wait_parity(bar0, phase);
          phase ^= 1;
          in_burst += clock64() - b0;
          nmma += g.burst;
        }
        const long long i0 = clock64();
        while (clock64() - i0 < g.idle) { }
      }
      const long long t1 = clock64();
      stop = 1;
      g.out[blockIdx.x * 16 + 0] = t1 - t0;
      g.out[blockIdx.x * 16 + 1] = in_burst;
      g.out[blockIdx.x * 16 + 10] = nmma;
    }
  } else if (!(g.skip0 && (warp & 3) == (g.mma_warp & 3))) {
    double d[8];
    float f[16];
    int n[16];
    long long w[8];
    uint32_t x = tid * 2654435761u;
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = 1.0 + tid * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) { f[i] = 1.0f + tid * 1e-3f + i; n[i] = tid + i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = tid * 77 + i;
    const double c1 = 1.0000001, c2 = 1e-9;
    long long iters = 0;
    while (!stop) {
#pragma unroll 1
      for (int rep = 0; rep < 16; ++rep) {
        if (g.kind == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            x ^= 0x80000000u + i;   // LOP3
            const double v = __hiloint2double(0x43300000, (int)x) - 4503601774854144.0;   // DADD
            d[i] = fma(v, c2, d[i]);   // DFMA
          }
        } else if (g.kind == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], 1.0000001f, 1e-9f);
        } else if (g.kind == 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) n[i] = n[i] * 3 + i;
        } else if (g.kind == 3) {
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = fma(d[i], c1, c2);
        } else if (g.kind == 4) {   // IMAD.WIDE: 32 x 32 + 64
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = (long long)n[i] * (long long)n[i + 8] + w[i];
        } else if (g.kind == 5) {   // I2FP.F32.S32 + FFMA (the FP32 fold)
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = fmaf((float)(n[i] ^ (int)x), 1e-9f, f[i]);
        } else if (g.kind == 6) {   // DADD
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = d[i] + c2;
        } else if (g.kind == 7) {   // I2F.F64.S32 (conversion pipe) feeding nothing FP64: result kept via integer xor of its words
#pragma unroll
          for (int i = 0; i < 8; ++i) { const double v = (double)n[i]; n[i] = __double2hiint(v) ^ __double2loint(v); }
        } else {                    // DMUL
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = d[i] * c1;
        }
      }
      iters += 16;
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i] + n[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (double)w[i];
    if (s == 1.2345) g.sink[0] = s + x;
    if ((tid & 31) == 0) g.out[blockIdx.x * 16 + 2 + (warp > g.mma_warp ? warp - 1 : warp)] = iters;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  long long* dout; double* dsink;
  CK(cudaMalloc(&dout, sizeof(long long) * 16 * nsm)); CK(cudaMalloc(&dsink, 8));
  CK(cudaFuncSetAttribute(duty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const char* kinds[NKIND] = {"LOP3+DADD+DFMA (the fold)", "FFMA", "IMAD", "DFMA", "IMAD.WIDE", "I2FP.F32.S32+FFMA", "DADD", "I2F.F64.S32 (+LOP3)", "DMUL"};
  struct Case { int burst, idle; };
  // one MMA ~ 112 clk: bursts of 4 / 18 / 90 MMAs ~ 0.45 k / 2 k / 10 k clk, each followed by as long an idle gap, plus always-on and never
  const Case cases[] = {{0, 1000}, {64, 0}, {18, 2000}, {18, 700}};
  for (int skip0 = 0; skip0 < 3; ++skip0)
  for (int kind = 0; kind < NKIND; ++kind) {
    if (skip0 && kind != 1 && kind != 4 && kind != 5) continue;
    printf("== math kind %d: %s%s\n", kind, kinds[kind], skip0 == 1 ? " -- no math warps on the MMA warp's sub-partition" : (skip0 == 2 ? " -- the MMA warp is warp 0 (the oldest)" : ""));
    for (const Case& cs : cases) {
      Args g{dout, dsink, cs.burst, cs.idle, kind, skip0 == 1, skip0 == 2 ? 0 : MATH_WARPS, 4000000ll};
      CK(cudaMemset(dout, 0, sizeof(long long) * 16 * nsm));
      duty_kernel<<<nsm, THREADS, SMEM_BYTES>>>(g);
      CK(cudaDeviceSynchronize());
      std::vector<long long> o(16 * nsm);
      CK(cudaMemcpy(o.data(), dout, sizeof(long long) * o.size(), cudaMemcpyDeviceToHost));
      double tot = 0, inb = 0, it = 0, nm = 0;
      for (int b = 0; b < nsm; ++b) {
        tot += (double)o[b * 16]; inb += (double)o[b * 16 + 1]; nm += (double)o[b * 16 + 10];
        for (int w = 0; w < MATH_WARPS; ++w) it += (double)o[b * 16 + 2 + w];
      }
      const double instr = it * INSTR_PER_ITER[kind];   // warp instructions over all math warps
      printf("burst %3d MMAs, idle %5d clk: tensor duty %.2f (%.0f clk per MMA in burst); math %.3f instr/clk/sub-partition\n", cs.burst, cs.idle,
             inb / tot, nm > 0 ? inb / nm : 0.0, instr / tot / (skip0 == 1 ? 3.0 : 4.0));
    }
  }
  return 0;
}
