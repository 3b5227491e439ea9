// rdm_pack.h -- host-side preparation of the reduced density matrices for the kernels.
//
//  * fold_tpdm:  the dense alpha-beta 2-RDM in the reference's compound-index convention
//                (mcpdft.cc:204: element (nu*n+mu, sigma*n+lambda) multiplies
//                phi_mu phi_nu phi_lambda phi_sigma) is folded onto unordered orbital pairs
//                I = {mu,nu}, J = {lambda,sigma} and symmetrised:  Dt (M x M), M = n(n+1)/2,
//                so that Pi = sum_IJ X_I Dt_IJ X_J  with X_I = phi_t phi_u.
//  * pack_dfrag: Dt cut into 32x32 blocks, upper block triangle only (off-diagonal blocks
//                doubled), laid out in DMMA.8x8x4 B-fragment order per task lane (k2_ontop.cuh).
//  * symmetrise_opdm: Ds = (D + D^T)/2 for K1.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace ordm {

inline int pair_count(int n) { return n * (n + 1) / 2; }

// packed index of the unordered pair {t,u}: rows t = 0..n-1, u = t..n-1
inline int pair_index(int n, int t, int u) {
  if (t > u) { int s = t; t = u; u = s; }
  return t * n - t * (t - 1) / 2 + (u - t);
}

inline void pair_table(int n, int Mpad, std::vector<uint16_t>& tab) {
  tab.assign((size_t)2 * Mpad, 0);
  int I = 0;
  for (int t = 0; t < n; ++t)
    for (int u = t; u < n; ++u, ++I) {
      tab[2 * I] = (uint16_t)t;
      tab[2 * I + 1] = (uint16_t)u;
    }
}

// D2: (n*n) x (n*n) column-major.  Dt: M x M row-major (symmetric).
inline void fold_tpdm(int n, const double* D2, std::vector<double>& Dt) {
  const int M = pair_count(n);
  const size_t n2 = (size_t)n * n;
  Dt.assign((size_t)M * M, 0.0);
  for (int si = 0; si < n; ++si)
    for (int la = 0; la < n; ++la) {
      const int J = pair_index(n, si, la);
      const double* col = D2 + ((size_t)si * n + la) * n2;
      for (int nu = 0; nu < n; ++nu)
        for (int mu = 0; mu < n; ++mu) Dt[(size_t)pair_index(n, nu, mu) * M + J] += col[(size_t)nu * n + mu];
    }
  for (int I = 0; I < M; ++I)
    for (int J = I + 1; J < M; ++J) {
      const double s = 0.5 * (Dt[(size_t)I * M + J] + Dt[(size_t)J * M + I]);
      Dt[(size_t)I * M + J] = Dt[(size_t)J * M + I] = s;
    }
}

struct DfragLayout {
  int M = 0, nblk = 0, ntask = 0, nlanes = 0;
  uint32_t lane_ofs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<double> data;
};

// blk = 32, 8 k-steps x (32 lanes x 4 n-tiles) doubles per task, `prefetch_steps` zero k-steps appended.
inline void pack_dfrag(int M, const std::vector<double>& Dt, int nlanes, int prefetch_steps, DfragLayout& out) {
  const int BLK = 32, KSTEPS = 8, NT = 4;
  out.M = M;
  out.nblk = (M + BLK - 1) / BLK;
  out.ntask = out.nblk * (out.nblk + 1) / 2;
  out.nlanes = nlanes;
  const size_t task_doubles = (size_t)KSTEPS * 32 * NT;
  out.data.assign((size_t)out.ntask * task_doubles + (size_t)prefetch_steps * 32 * NT, 0.0);
  uint32_t ofs = 0;
  for (int l = 0; l < 8; ++l) {
    out.lane_ofs[l] = ofs;
    if (l < nlanes && out.ntask > l) ofs += (uint32_t)((out.ntask - l + nlanes - 1) / nlanes);
  }
  // enumerate tasks tau = (jb outer, ib <= jb inner)
  int tau = 0;
  for (int jb = 0; jb < out.nblk; ++jb)
    for (int ib = 0; ib <= jb; ++ib, ++tau) {
      const int lane_q = tau % nlanes, k_in_lane = tau / nlanes;
      double* dst = out.data.data() + ((size_t)out.lane_ofs[lane_q] + k_in_lane) * task_doubles;
      const double scale = (ib < jb) ? 2.0 : 1.0;
      for (int ks = 0; ks < KSTEPS; ++ks)
        for (int lane = 0; lane < 32; ++lane)
          for (int nt = 0; nt < NT; ++nt) {
            const int I = ib * BLK + ks * 4 + (lane & 3);
            const int J = jb * BLK + nt * 8 + (lane >> 2);
            const double v = (I < M && J < M) ? scale * Dt[(size_t)I * M + J] : 0.0;
            dst[((size_t)ks * 32 + lane) * NT + nt] = v;
          }
    }
}

// Ds = (D + D^T)/2, written row-major with row stride `stride` (zero padded).
inline void symmetrise_opdm(int n, const double* D, int stride, std::vector<double>& Ds) {
  Ds.assign((size_t)stride * stride, 0.0);
  for (int t = 0; t < n; ++t)
    for (int u = 0; u < n; ++u) Ds[(size_t)t * stride + u] = 0.5 * (D[t + (size_t)u * n] + D[u + (size_t)t * n]);
}

}  // namespace ordm
