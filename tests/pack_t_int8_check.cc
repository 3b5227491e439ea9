// Host-logic check of ordm::pack_t_int8 (csrc/rdm_pack.h): the byte image handed to the int8 tensor-core kernel
// (k2_int8.cuh) must decode, slice by slice, to T = block upper triangle of D~ (diagonal blocks halved) in 46-bit
// fixed point, with T + T^T = D~.  Compiled and run by tests/test_abi.py (no GPU).
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../openrdm_b200/csrc/rdm_pack.h"

int main(int argc, char** argv) {
  const int nact = argc > 1 ? atoi(argv[1]) : 20;
  const int M = ordm::pair_count(nact), kp = (M + 31) / 32 * 32;
  std::mt19937_64 rng(7 + nact);
  std::uniform_real_distribution<double> ud(-1.5, 1.5);
  std::vector<double> Dt((size_t)M * M);
  for (int i = 0; i < M; ++i)
    for (int j = i; j < M; ++j) Dt[(size_t)i * M + j] = Dt[(size_t)j * M + i] = ud(rng) * (((i * 31 + j) % 7) ? 1.0 : 1e-9);
  std::vector<uint8_t> img;
  int t_exp = 0;
  ordm::pack_t_int8(M, Dt, kp, img, t_exp);
  const size_t slice = (size_t)ordm::t_int8_slice_bytes(kp);
  if (img.size() != 6 * slice) { printf("image size %zu, expected %zu\n", img.size(), 6 * slice); return 1; }
  std::vector<double> T((size_t)kp * kp, 0.0);
  double worst = 0.0, tmax = 0.0;
  for (int kc = 0; kc < kp / 32; ++kc) {
    const int rows = kp - 32 * kc;
    for (int n = 0; n < rows; ++n)
      for (int k = 0; k < 32; ++k) {
        const size_t at = (size_t)ordm::t_int8_chunk_off(kp, kc) + (size_t)((k >> 4) * (rows >> 3) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 15);
        long long v = 0;
        for (int l = 0; l < 6; ++l) v += (long long)(int8_t)img[(size_t)l * slice + at] * (1ll << (8 * l));   // balanced signed digits
        if (v > (1ll << 46) || v < -(1ll << 46)) { printf("image out of range at chunk %d (%d, %d): %lld\n", kc, n, k, v); return 1; }
        T[(size_t)(32 * kc + k) * kp + 32 * kc + n] = std::ldexp((double)v, t_exp - 46);
      }
  }
  for (int i = 0; i < kp; ++i)
    for (int j = 0; j < kp; ++j) {
      const double want = (i < M && j < M) ? Dt[(size_t)i * M + j] : 0.0;
      const double got = T[(size_t)i * kp + j] + T[(size_t)j * kp + i];
      worst = std::max(worst, std::fabs(got - want));
      tmax = std::max(tmax, std::fabs(want));
      if (i / 32 > j / 32 && T[(size_t)i * kp + j] != 0.0) { printf("T not block upper triangular at (%d, %d)\n", i, j); return 1; }
    }
  const double bound = std::ldexp(1.0, t_exp - 46);   // two roundings of half an ulp of the image
  printf("nact %d: M %d, kp %d, t_exp %d, max |T + T^T - D~| = %.3e (bound %.3e, max |D~| %.3g)\n", nact, M, kp, t_exp, worst, bound, tmax);
  return worst <= bound && std::ldexp(1.0, t_exp) > tmax * 0.5 ? 0 : 1;
}
