// ordm_synth_host.cc -- host-only build of the synthetic input generator (include/ordm_synth.h,
// csrc/synth_rdm.h) as its own tiny shared library, libordm_synth.so.
//
// "Data loading" row of BASELINE.json's north_star: synthetic RDMs and grids of the named shapes are
// generated on the host when no input file exists.  The same two entry points exist in
// libopenrdm_b200.so; this copy has no CUDA dependency, so that the CPU reference arm of bench.py and
// the golden-fixture generators can produce the benchmark inputs without mapping the GPU library.
// The grid fill is split over host threads by point ranges (every value is a pure function of
// (seed, stream, global index), so the split does not change a bit).
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/ordm_synth.h"
#include "../csrc/synth_rdm.h"

extern "C" {

int ordm_synth_fill_host(uint64_t seed, size_t npts_total, size_t p0, size_t m, int n, double* w, double* phi,
                         double* px, double* py, double* pz) {
  if (p0 + m > npts_total || n <= 0) return 1;
  const uint64_t env_key = ordm_stream_key(seed, ORDM_STREAM_ENV), w_key = ordm_stream_key(seed, ORDM_STREAM_W);
  const double inv = 1.0 / (double)npts_total;
  double* dst[4] = {phi, px, py, pz};
  uint64_t key[4];
  for (int k = 0; k < 4; ++k) key[k] = ordm_stream_key(seed, ORDM_STREAM_PHI + k);
  unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
  if (m < 65536) nt = 1;
  auto work = [&](size_t b, size_t e) {
    if (w) for (size_t p = b; p < e; ++p) w[p] = ordm_synth_w(w_key, inv, p0 + p);
    for (int k = 0; k < 4; ++k) {
      if (!dst[k]) continue;
      for (int i = 0; i < n; ++i) {
        double* col = dst[k] + (size_t)i * m;
        for (size_t p = b; p < e; ++p) col[p] = ordm_synth_phi(key[k], ordm_env_key(env_key, p0 + p), npts_total, p0 + p, (uint64_t)i);
      }
    }
  };
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, m * t / nt, m * (t + 1) / nt);
  for (auto& x : th) x.join();
  return 0;
}

int ordm_synth_rdms(uint64_t seed, int nact, int na, int nb, int ndet, double* A1a, double* A1b, double* D2act) {
  if (nact <= 0 || na < 0 || nb < 0 || na > nact || nb > nact || ndet <= 0 || !A1a || !A1b) return 1;
  ordm::synth_rdms(seed, nact, na, nb, ndet, A1a, A1b, D2act);
  return 0;
}

}  // extern "C"
