// synth_rdm.h -- host-side synthetic active-space RDMs for BASELINE.json's configs 3-5
// (SURVEY.md section 8d): an N-representable ensemble of `ndet` determinants
//     D1s  = sum_k c_k U_k diag(o_k^s) U_k^T                (s = alpha, beta)
//     D2ab = sum_k c_k kron(D1a^k, D1b^k)                    (index convention of build_tpdm,
//                                                             mcpdft.cc:640-652)
// so Pi = sum_k c_k rho_a^k rho_b^k >= 0 and R = 4 Pi / rho^2 falls on both sides of 1.
// U_k: Gram-Schmidt of a seeded uniform matrix; occupations: seeded Fisher-Yates subsets,
// alpha != beta pattern for odd k.  Only +,-,*,/ and sqrt are used (no libm transcendental),
// so the matrices are bit-reproducible across hosts.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>
#include "ordm_synth.h"

namespace ordm {

inline void synth_orthogonal(uint64_t key, int n, std::vector<double>& U) {  // column-major n x n
  U.assign((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) U[i + (size_t)j * n] = ordm_u01_key(key, (uint64_t)j * n + i) - 0.5 + (i == j ? 1.0 : 0.0);
  for (int pass = 0; pass < 2; ++pass)  // modified Gram-Schmidt, twice for orthogonality to ~1e-16
    for (int j = 0; j < n; ++j) {
      double* cj = &U[(size_t)j * n];
      for (int k = 0; k < j; ++k) {
        const double* ck = &U[(size_t)k * n];
        double d = 0.0;
        for (int i = 0; i < n; ++i) d += ck[i] * cj[i];
        for (int i = 0; i < n; ++i) cj[i] -= d * ck[i];
      }
      double nn = 0.0;
      for (int i = 0; i < n; ++i) nn += cj[i] * cj[i];
      nn = 1.0 / std::sqrt(nn);
      for (int i = 0; i < n; ++i) cj[i] *= nn;
    }
}

inline void synth_occupation(uint64_t key, int n, int nelec, std::vector<int>& occ) {
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int i = n - 1; i > 0; --i) {
    int j = (int)(ordm_u01_key(key, (uint64_t)i) * (i + 1));
    if (j > i) j = i;
    int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
  }
  occ.assign(n, 0);
  for (int i = 0; i < nelec && i < n; ++i) occ[perm[i]] = 1;
}

// A1a, A1b: n x n column-major.  D2: (n^2)x(n^2) column-major, may be null.
inline void synth_rdms(uint64_t seed, int n, int nelec_a, int nelec_b, int ndet, double* A1a, double* A1b, double* D2) {
  const size_t n2 = (size_t)n * n;
  for (size_t i = 0; i < n2; ++i) A1a[i] = A1b[i] = 0.0;
  if (D2) for (size_t i = 0; i < n2 * n2; ++i) D2[i] = 0.0;
  std::vector<double> c(ndet);
  double csum = 0.0;
  for (int k = 0; k < ndet; ++k) { c[k] = 0.25 + ordm_u01(seed, ORDM_STREAM_RDM, (uint64_t)k); csum += c[k]; }
  std::vector<double> U, Da(n2), Db(n2);
  std::vector<int> oa, ob;
  for (int k = 0; k < ndet; ++k) {
    const double ck = c[k] / csum;
    synth_orthogonal(ordm_stream_key(seed, ORDM_STREAM_RDM + 1 + 3 * k), n, U);
    synth_occupation(ordm_stream_key(seed, ORDM_STREAM_RDM + 2 + 3 * k), n, nelec_a, oa);
    if (k & 1) synth_occupation(ordm_stream_key(seed, ORDM_STREAM_RDM + 3 + 3 * k), n, nelec_b, ob);
    else { ob = oa; if (nelec_b != nelec_a) synth_occupation(ordm_stream_key(seed, ORDM_STREAM_RDM + 3 + 3 * k), n, nelec_b, ob); }
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        double sa = 0.0, sb = 0.0;
        for (int m = 0; m < n; ++m) {
          const double uu = U[i + (size_t)m * n] * U[j + (size_t)m * n];
          if (oa[m]) sa += uu;
          if (ob[m]) sb += uu;
        }
        Da[i + (size_t)j * n] = sa;
        Db[i + (size_t)j * n] = sb;
      }
    for (size_t i = 0; i < n2; ++i) { A1a[i] += ck * Da[i]; A1b[i] += ck * Db[i]; }
    if (D2)
      for (int j = 0; j < n; ++j)
        for (int l = 0; l < n; ++l) {
          double* col = D2 + ((size_t)j * n + l) * n2;
          for (int i = 0; i < n; ++i) {
            const double a = ck * Da[i + (size_t)j * n];
            for (int kk = 0; kk < n; ++kk) col[(size_t)i * n + kk] += a * Db[kk + (size_t)l * n];
          }
        }
  }
}

}  // namespace ordm
