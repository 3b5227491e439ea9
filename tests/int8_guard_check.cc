// Host-logic helper for tests/test_int8_scheme.py: ordm::int8_guard_constants (csrc/rdm_pack.h) on a folded pair matrix
// read from a raw file of M*M doubles.  Prints "t_exp g1 g2" for GMIN 4 and 5.  No GPU.
#include <cstdio>
#include <cstdlib>
#include "../openrdm_b200/csrc/rdm_pack.h"
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const int M = atoi(argv[1]), kp = (M + 31) / 32 * 32;
  std::vector<double> Dt((size_t)M * M);
  FILE* f = fopen(argv[2], "rb");
  if (!f || fread(Dt.data(), sizeof(double), Dt.size(), f) != Dt.size()) return 3;
  fclose(f);
  std::vector<uint8_t> img;
  int t_exp = 0;
  ordm::pack_t_int8(M, Dt, kp, img, t_exp);
  for (int gmin = 4; gmin <= 5; ++gmin) {
    double g1, g2;
    ordm::int8_guard_constants(M, Dt, t_exp, gmin, g1, g2);
    printf("%d %.17g %.17g\n", t_exp, g1, g2);
  }
  return 0;
}
