// Bitwise check of k2::mul_f64_intpipe against the hardware DMUL on B200: random normals over the whole
// exponent range, values around the range ends, zeros, subnormals, infinities, NaNs, exact ties.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
// IEEE-754 binary64 multiplication (round to nearest even) on the integer pipe: tried for K2's X build
// (a hardware DMUL issued by the builder warps waits ~260 cycles for the FP64 pipe behind the DMMA
// stream).  Exact -- 0 mismatches below -- but ~80 integer instructions per product: the X build got
// slower (13 k -> 15 k cycles per tile) and the MMA warps' address arithmetic with it (K2 3.17 -> 4.83 ms
// on 2 M points), so the product kernel keeps the hardware multiply.  Kept here as the record.
namespace k2 {
__device__ __forceinline__ double mul_f64_intpipe(double a, double b, bool& ok) {
  const unsigned long long ua = (unsigned long long)__double_as_longlong(a), ub = (unsigned long long)__double_as_longlong(b);
  const int ea = (int)((ua >> 52) & 0x7ffull), eb = (int)((ub >> 52) & 0x7ffull);
  const int e = ea + eb - 1023;
  ok = ok && !(ea == 0 || eb == 0 || ea == 0x7ff || eb == 0x7ff || e < 2 || e > 0x7fc);
  const unsigned long long A = ((ua & 0x000fffffffffffffull) | 0x0010000000000000ull) << 11;
  const unsigned long long B = ((ub & 0x000fffffffffffffull) | 0x0010000000000000ull) << 11;
  const unsigned long long hi = __umul64hi(A, B), lo = A * B;  // A*B = (ma*mb) << 22, in [2^126, 2^128)
  const int top = (int)(hi >> 63);
  const int sh = 10 + top;                                     // top 53 bits of the product = hi >> sh
  unsigned long long m = hi >> sh;
  const unsigned long long rem = hi & ((1ull << sh) - 1ull), half = 1ull << (sh - 1);
  const bool up = rem > half || (rem == half && (lo != 0ull || (m & 1ull)));
  m += up ? 1ull : 0ull;
  const int carry = (int)(m >> 53);                            // mantissa rounded up to 2^53
  m >>= carry;
  const int ex = e + top + carry;
  const unsigned long long r = ((ua ^ ub) & 0x8000000000000000ull) | ((unsigned long long)(unsigned)ex << 52) | (m & 0x000fffffffffffffull);
  return __longlong_as_double((long long)r);
}
__device__ __noinline__ double hw_mul_f64(double a, double b) { return a * b; }
}  // namespace k2
__global__ void k(const double* a, const double* b, unsigned long long* bad, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool ok = true;
  double x = k2::mul_f64_intpipe(a[i], b[i], ok);
  if (!ok) x = k2::hw_mul_f64(a[i], b[i]);
  const double y = a[i] * b[i];
  const bool same = (__double_as_longlong(x) == __double_as_longlong(y)) || (x != x && y != y);
  if (!same) { unsigned long long k = atomicAdd(bad, 1ull); if (k < 8) printf("MISMATCH %a * %a: emu %a hw %a\n", a[i], b[i], x, y); }
}
static unsigned long long rng_state = 88172645463325252ull;
static unsigned long long rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
int main() {
  const size_t n = 1 << 24;
  std::vector<double> a(n), b(n);
  for (size_t i = 0; i < n; ++i) {
    unsigned long long u = rnd(), v = rnd();
    const int mode = (int)(i % 16);
    if (mode < 8) {            // random mantissas, exponents in a band where products stay normal
      u = (u & 0x800fffffffffffffull) | ((unsigned long long)(512 + rnd() % 1024) << 52);
      v = (v & 0x800fffffffffffffull) | ((unsigned long long)(512 + rnd() % 1024) << 52);
    } else if (mode < 11) {    // anything at all (subnormals, inf, nan, overflow, underflow)
    } else if (mode < 13) {    // few mantissa bits: exact products and ties
      u = (u & 0x8000000000000000ull) | ((unsigned long long)(900 + rnd() % 200) << 52) | ((rnd() & 0x3ffffffull) << 26);
      v = (v & 0x8000000000000000ull) | ((unsigned long long)(900 + rnd() % 200) << 52) | ((rnd() & 0x7ffffffull) << 25);
    } else if (mode == 13) {   // zeros and small integers
      double z[4] = {0.0, -0.0, 3.0, 0.5}; std::memcpy(&u, &z[rnd() % 4], 8); std::memcpy(&v, &z[rnd() % 4], 8);
    } else {                   // magnitudes like orbital values on a grid: 1e-14 .. 1
      double x = ((double)(rnd() % 2000001) / 1e6 - 1.0) * pow(10.0, -(double)(rnd() % 15));
      double y = ((double)(rnd() % 2000001) / 1e6 - 1.0) * pow(10.0, -(double)(rnd() % 15));
      std::memcpy(&u, &x, 8); std::memcpy(&v, &y, 8);
    }
    std::memcpy(&a[i], &u, 8); std::memcpy(&b[i], &v, 8);
  }
  double *da, *db; unsigned long long* dbad, bad = 0;
  cudaMalloc(&da, n * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dbad, 8); cudaMemset(dbad, 0, 8);
  cudaMemcpy(da, a.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
  k<<<(unsigned)((n + 255) / 256), 256>>>(da, db, dbad, n);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(&bad, dbad, 8, cudaMemcpyDeviceToHost);
  printf("mul_f64_intpipe vs DMUL: %zu products, %llu mismatches (%s)\n", n, bad, cudaGetErrorString(e));
  return bad != 0;
}
