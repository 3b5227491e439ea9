// oracle/ref_harness.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A C-ABI wrapper around OpenRDM's *unmodified* hot-path classes, compiled by oracle/Makefile
// together with the reference's own sources where they lie under /root/reference:
//   src/libmcpdft/{mcpdft,utility,energy}.cc  src/libfunc/functional.cc  src/libmem/libMem.cc
// into oracle/_ref/libopenrdm_ref.so.  No reference source is copied into this repository.
//
// It feeds caller-provided arrays through the reference's public setters (mcpdft.h:112-164),
// runs mcpdft::mcpdft_energy (energy.cc:27-188) or its stages, and reads the per-point vectors
// back through the reference's getters (mcpdft.h:64-110).
//
// `#define private public` (this TU only): MCPDFT::is_gga_ is private with no setter and is
// only ever set from the test-case directory name (mcpdft.cc:24-29); the harness needs to
// select the GGA path for arbitrary inputs.  Access specifiers do not change the layout GCC
// gives the class, so this TU stays link-compatible with the unmodified mcpdft.cc.
#include <armadillo>
#define private public
#include "mcpdft.h"
#undef private
#include "energy.h"
#include "functional.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <unistd.h>

namespace {

arma::vec to_vec(const double* p, size_t n) {
  arma::vec v(n);
  if (p && n) std::memcpy(v.memptr(), p, n * sizeof(double));
  return v;
}
arma::mat to_mat(const double* p, size_t r, size_t c) {
  arma::mat m(r, c);
  if (p && r * c) std::memcpy(m.memptr(), p, r * c * sizeof(double));
  return m;
}
void out_vec(const arma::vec& v, double* dst, size_t n) {
  if (dst && v.n_elem == n && n) std::memcpy(dst, v.memptr(), n * sizeof(double));
}

// The reference prints a banner, RAM figures and tables on every call (mcpdft.cc:100-104,
// energy.cc:34-40,175-179).  Silence fd 1 while it runs when asked to.
struct StdoutMute {
  int saved = -1;
  explicit StdoutMute(bool on) {
    if (!on) return;
    std::fflush(stdout);
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    close(nul);
  }
  ~StdoutMute() {
    if (saved < 0) return;
    std::fflush(stdout);
    dup2(saved, 1);
    close(saved);
  }
};

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

// must match oracle/mcpdft_oracle.c
enum {
  ORC_RHOA, ORC_RHOB, ORC_RHO,
  ORC_RHOA_X, ORC_RHOA_Y, ORC_RHOA_Z, ORC_RHOB_X, ORC_RHOB_Y, ORC_RHOB_Z,
  ORC_SIGMA_AA, ORC_SIGMA_AB, ORC_SIGMA_BB,
  ORC_PI, ORC_R, ORC_TR_RHOA, ORC_TR_RHOB,
  ORC_TR_SIGMA_AA, ORC_TR_SIGMA_AB, ORC_TR_SIGMA_BB,
  ORC_PI_X, ORC_PI_Y, ORC_PI_Z,
  ORC_NVEC
};

typedef struct {
  double e_tot, ex, ec, eclass;
  double n_tot, n_a, n_b;
  double trn_tot, trn_a, trn_b;
} orc_result;

// per-stage wall-clock seconds of one staged run
typedef struct {
  double t_rho, t_grad_rho, t_pi, t_grad_pi, t_R_translate, t_functional, t_total;
} ref_timing;

static void fill_mc(mcpdft::MCPDFT& mc, size_t npts, int n, int is_gga, const double* w,
                    const double* phi, const double* phix, const double* phiy, const double* phiz,
                    const double* D1a, const double* D1b, double eclass) {
  mc.is_gga_ = is_gga != 0;
  mc.set_npts(npts);
  mc.set_nbfs(n);
  mc.set_w(to_vec(w, npts));
  mc.set_phi(to_mat(phi, npts, n));
  if (is_gga) {
    mc.set_phi_x(to_mat(phix, npts, n));
    mc.set_phi_y(to_mat(phiy, npts, n));
    mc.set_phi_z(to_mat(phiz, npts, n));
  }
  mc.set_D1a(to_mat(D1a, n, n));
  mc.set_D1b(to_mat(D1b, n, n));
  mc.set_eclass(eclass);
  mc.set_eref(0.0);
}

static void read_back(const mcpdft::MCPDFT& mc, size_t npts, int is_gga, double** vecs) {
  if (!vecs) return;
  out_vec(mc.get_rhoa(), vecs[ORC_RHOA], npts);
  out_vec(mc.get_rhob(), vecs[ORC_RHOB], npts);
  out_vec(mc.get_rho(), vecs[ORC_RHO], npts);
  out_vec(mc.get_pi(), vecs[ORC_PI], npts);
  out_vec(mc.get_R(), vecs[ORC_R], npts);
  out_vec(mc.get_tr_rhoa(), vecs[ORC_TR_RHOA], npts);
  out_vec(mc.get_tr_rhob(), vecs[ORC_TR_RHOB], npts);
  if (is_gga) {
    out_vec(mc.get_rhoa_x(), vecs[ORC_RHOA_X], npts);
    out_vec(mc.get_rhoa_y(), vecs[ORC_RHOA_Y], npts);
    out_vec(mc.get_rhoa_z(), vecs[ORC_RHOA_Z], npts);
    out_vec(mc.get_rhob_x(), vecs[ORC_RHOB_X], npts);
    out_vec(mc.get_rhob_y(), vecs[ORC_RHOB_Y], npts);
    out_vec(mc.get_rhob_z(), vecs[ORC_RHOB_Z], npts);
    out_vec(mc.get_sigma_aa(), vecs[ORC_SIGMA_AA], npts);
    out_vec(mc.get_sigma_ab(), vecs[ORC_SIGMA_AB], npts);
    out_vec(mc.get_sigma_bb(), vecs[ORC_SIGMA_BB], npts);
    out_vec(mc.get_tr_sigma_aa(), vecs[ORC_TR_SIGMA_AA], npts);
    out_vec(mc.get_tr_sigma_ab(), vecs[ORC_TR_SIGMA_AB], npts);
    out_vec(mc.get_tr_sigma_bb(), vecs[ORC_TR_SIGMA_BB], npts);
    // pi_x_/pi_z_ have setters but no getters (mcpdft.h:146-148); private members are
    // reachable here only because of the harness-only access hack above.
    out_vec(mc.pi_x_, vecs[ORC_PI_X], npts);
    out_vec(mc.pi_y_, vecs[ORC_PI_Y], npts);  // never set by the reference (mcpdft.cc:268)
    out_vec(mc.pi_z_, vecs[ORC_PI_Z], npts);
  }
}

// The integrated densities are only printed by the reference (mcpdft.cc:100-104,339-343);
// re-accumulate them exactly as those loops do (point order, same expressions).
static void integrate(const mcpdft::MCPDFT& mc, size_t npts, orc_result* out) {
  arma::vec w(mc.get_w()), a(mc.get_rhoa()), b(mc.get_rhob()), ta(mc.get_tr_rhoa()), tb(mc.get_tr_rhob());
  double s0 = 0, s1 = 0, s2 = 0, t0 = 0, t1 = 0, t2 = 0;
  for (size_t p = 0; p < npts; ++p) {
    s1 += a(p) * w(p);
    s2 += b(p) * w(p);
    s0 += (a(p) + b(p)) * w(p);
    t1 += ta(p) * w(p);
    t2 += tb(p) * w(p);
    t0 += (ta(p) + tb(p)) * w(p);
  }
  out->n_tot = s0; out->n_a = s1; out->n_b = s2;
  out->trn_tot = t0; out->trn_a = t1; out->trn_b = t2;
}

// Translation used by ref_energy_staged: 0 = MCPDFT::translate (what mcpdft_energy calls),
// 1 = MCPDFT::fully_translate (mcpdft.cc:418-606).  The fully-translated gradient routine reads
// pi_y_, which the reference's build_pi never stores (mcpdft.cc:268 writes pi_y into the pi_z
// slot); the harness obtains d(Pi)/dy from the reference's own routine by running it once more
// with the y-derivatives of the orbitals in the x slot, and installs it with set_pi_y().
static int g_fully_translated = 0;
void ref_set_translation(int fully) { g_fully_translated = fully; }

// functional: "SVWN" or anything else (= PBE), as energy.cc:114-120.
// Runs mcpdft::mcpdft_energy() exactly as tests/main.cc:38 does (for GGA that includes the
// grad-Pi stage whose output never reaches the energy).
int ref_energy_as_written(size_t npts, int n, int is_gga, const char* functional, const double* w,
                          const double* phi, const double* phix, const double* phiy,
                          const double* phiz, const double* D1a, const double* D1b,
                          const double* D2ab, double eclass, int quiet, orc_result* out,
                          double** vecs, double* seconds) {
  StdoutMute mute(quiet != 0);
  mcpdft::MCPDFT mc;
  fill_mc(mc, npts, n, is_gga, w, phi, phix, phiy, phiz, D1a, D1b, eclass);
  arma::mat d1a(mc.get_D1a()), d1b(mc.get_D1b());
  arma::mat d2 = to_mat(D2ab, (size_t)n * n, (size_t)n * n);
  std::string f(functional);
  double t0 = now_s();
  double e = mcpdft::mcpdft_energy(&mc, f, d1a, d1b, d2);
  if (seconds) *seconds = now_s() - t0;
  std::memset(out, 0, sizeof(*out));
  out->e_tot = e;
  out->eclass = eclass;
  integrate(mc, npts, out);
  // Ex and Ec are only printed by mcpdft_energy; recompute them with the reference's own
  // Functional class on the reference's own translated vectors.
  mcpdft::Functional fn;
  arma::vec ta(mc.get_tr_rhoa()), tb(mc.get_tr_rhob());
  arma::vec saa(npts, arma::fill::zeros), sab(npts, arma::fill::zeros), sbb(npts, arma::fill::zeros);
  if (is_gga) { saa = mc.get_tr_sigma_aa(); sab = mc.get_tr_sigma_ab(); sbb = mc.get_tr_sigma_bb(); }
  if (f == "SVWN") { out->ex = fn.EX_LSDA(&mc, ta, tb); out->ec = fn.EC_VWN3(&mc, ta, tb); }
  else { out->ex = fn.EX_PBE(&mc, ta, tb, saa, sbb); out->ec = fn.EC_PBE(&mc, ta, tb, saa, sab, sbb); }
  read_back(mc, npts, is_gga, vecs);
  return 0;
}

// Staged run with per-stage timing.  with_pi_grad=0 skips build_ontop_pair_density_gradients
// (mcpdft.cc:214-270), the "necessary work only" variant of BASELINE.md section 3.
//
// Core/active mode (ncore > 0): phi has ncore+nact columns, A1a/A1b are nact x nact, D2act is
// (nact^2)x(nact^2).  The reference's own kernels are run on what they can digest --
// build_rho on all ncore+nact orbitals with D1 = I (+) A, build_ontop_pair_density on the
// active columns with D2act -- and the closed-form core terms of SURVEY.md Appendix B.3
// (Pi = rc^2 + rc*rtb + rta*rc + Pi_act) are added by this harness before the reference's
// build_R / translate / Functional run.  ncore == 0 is the plain reference path.
int ref_energy_staged(size_t npts, int ncore, int nact, int is_gga, const char* functional,
                      const double* w, const double* phi, const double* phix, const double* phiy,
                      const double* phiz, const double* A1a, const double* A1b,
                      const double* D2act, double eclass, int with_pi_grad, int quiet,
                      orc_result* out, double** vecs, ref_timing* tm) {
  StdoutMute mute(quiet != 0);
  const int n = ncore + nact;
  std::vector<double> D1a((size_t)n * n, 0.0), D1b((size_t)n * n, 0.0);
  for (int i = 0; i < ncore; ++i) D1a[i + (size_t)i * n] = D1b[i + (size_t)i * n] = 1.0;
  for (int j = 0; j < nact; ++j)
    for (int i = 0; i < nact; ++i) {
      D1a[(ncore + i) + (size_t)(ncore + j) * n] = A1a[i + (size_t)j * nact];
      D1b[(ncore + i) + (size_t)(ncore + j) * n] = A1b[i + (size_t)j * nact];
    }
  mcpdft::MCPDFT mc;
  fill_mc(mc, npts, n, is_gga, w, phi, phix, phiy, phiz, D1a.data(), D1b.data(), eclass);
  ref_timing t;
  std::memset(&t, 0, sizeof(t));
  double T0 = now_s(), t0 = T0;
  mc.build_density_functions();
  t.t_rho = now_s() - t0;
  if (is_gga) { t0 = now_s(); mc.build_density_gradients(); t.t_grad_rho = now_s() - t0; }

  arma::mat d2 = to_mat(D2act, (size_t)nact * nact, (size_t)nact * nact);
  const bool ft = g_fully_translated != 0;
  if (ft && is_gga) with_pi_grad = 1;
  arma::vec piy_act(npts, arma::fill::zeros);
  if (ft && is_gga) {  // d(Pi_act)/dy through the reference's x-slot (see ref_set_translation)
    const size_t off = (size_t)ncore * npts;
    mcpdft::MCPDFT tmp;
    fill_mc(tmp, npts, nact, 1, w, phi + off, phiy + off, phiy + off, phiy + off, A1a, A1b, 0.0);
    tmp.build_ontop_pair_density_gradients(d2);
    piy_act = tmp.pi_x_;
  }
  if (ncore == 0) {
    t0 = now_s(); mc.build_ontop_pair_density(d2); t.t_pi = now_s() - t0;
    if (is_gga && with_pi_grad) { t0 = now_s(); mc.build_ontop_pair_density_gradients(d2); t.t_grad_pi = now_s() - t0; }
    if (ft && is_gga) mc.set_pi_y(piy_act);
  } else {
    const size_t off = (size_t)ncore * npts;
    mcpdft::MCPDFT act;  // the reference's kernels on the active columns only
    fill_mc(act, npts, nact, is_gga, w, phi + off, phix ? phix + off : 0, phiy ? phiy + off : 0,
            phiz ? phiz + off : 0, A1a, A1b, 0.0);
    t0 = now_s(); act.build_ontop_pair_density(d2); t.t_pi = now_s() - t0;
    if (is_gga && with_pi_grad) { t0 = now_s(); act.build_ontop_pair_density_gradients(d2); t.t_grad_pi = now_s() - t0; }
    t0 = now_s();
    act.build_density_functions();  // rho_t,a and rho_t,b
    arma::vec pa(act.get_pi()), rta(act.get_rhoa()), rtb(act.get_rhob());
    arma::vec pi(npts);
    for (size_t p = 0; p < npts; ++p) {
      double rc = 0.0;
      for (int i = 0; i < ncore; ++i) rc += phi[p + (size_t)i * npts] * phi[p + (size_t)i * npts];
      pi(p) = rc * rc + rc * rtb(p) + rta(p) * rc + pa(p);
    }
    mc.set_pi(pi);
    t.t_pi += now_s() - t0;
    if (is_gga && with_pi_grad) {  // core terms of grad Pi: 2 rc drc + drc (rta+rtb) + rc (drta + drtb) + dPi_act
      act.build_density_gradients();
      arma::vec gax[3] = {act.get_rhoa_x(), act.get_rhoa_y(), act.get_rhoa_z()};
      arma::vec gbx[3] = {act.get_rhob_x(), act.get_rhob_y(), act.get_rhob_z()};
      arma::vec pg[3] = {act.pi_x_, piy_act, act.pi_z_};
      const double* gphi[3] = {phix, phiy, phiz};
      for (size_t p = 0; p < npts; ++p) {
        double rc = 0.0, dc[3] = {0, 0, 0};
        for (int i = 0; i < ncore; ++i) {
          const size_t o = p + (size_t)i * npts;
          rc += phi[o] * phi[o];
          for (int c = 0; c < 3; ++c) dc[c] += gphi[c][o] * phi[o] + phi[o] * gphi[c][o];
        }
        for (int c = 0; c < 3; ++c) pg[c](p) += 2.0 * rc * dc[c] + dc[c] * (rta(p) + rtb(p)) + rc * (gax[c](p) + gbx[c](p));
      }
      mc.set_pi_x(pg[0]); mc.set_pi_y(pg[1]); mc.set_pi_z(pg[2]);
    }
  }
  t0 = now_s();
  mc.build_R();
  if (ft) mc.fully_translate(); else mc.translate();
  t.t_R_translate = now_s() - t0;

  mcpdft::Functional fn;
  arma::vec ta(mc.get_tr_rhoa()), tb(mc.get_tr_rhob());
  arma::vec saa(npts, arma::fill::zeros), sab(npts, arma::fill::zeros), sbb(npts, arma::fill::zeros);
  if (is_gga) { saa = mc.get_tr_sigma_aa(); sab = mc.get_tr_sigma_ab(); sbb = mc.get_tr_sigma_bb(); }
  std::memset(out, 0, sizeof(*out));
  t0 = now_s();
  if (std::string(functional) == "SVWN") { out->ex = fn.EX_LSDA(&mc, ta, tb); out->ec = fn.EC_VWN3(&mc, ta, tb); }
  else { out->ex = fn.EX_PBE(&mc, ta, tb, saa, sbb); out->ec = fn.EC_PBE(&mc, ta, tb, saa, sab, sbb); }
  t.t_functional = now_s() - t0;
  t.t_total = now_s() - T0;
  out->eclass = eclass;
  out->e_tot = 0.0; out->e_tot += eclass; out->e_tot += out->ex; out->e_tot += out->ec;  // energy.cc:181-183
  integrate(mc, npts, out);
  read_back(mc, npts, is_gga, vecs);
  if (tm) *tm = t;
  return 0;
}

// The four libfunc entry points on caller-provided vectors (functional.h:17-43).
double ref_functional(int which, size_t npts, const double* w, const double* ra, const double* rb,
                      const double* saa, const double* sab, const double* sbb) {
  mcpdft::MCPDFT mc;
  mc.set_npts(npts);
  mc.set_w(to_vec(w, npts));
  mcpdft::Functional fn;
  arma::vec a = to_vec(ra, npts), b = to_vec(rb, npts);
  switch (which) {
    case 0: return fn.EX_LSDA(&mc, a, b);
    case 1: return fn.EC_VWN3(&mc, a, b);
    case 2: return fn.EX_PBE(&mc, a, b, to_vec(saa, npts), to_vec(sbb, npts));
    default: return fn.EC_PBE(&mc, a, b, to_vec(saa, npts), to_vec(sab, npts), to_vec(sbb, npts));
  }
}

// D2ab = kron(D1a, D1b) through the reference's own build_tpdm (mcpdft.cc:640-652).
void ref_build_tpdm(int n, const double* D1a, const double* D1b, double* D2ab) {
  mcpdft::MCPDFT mc;
  mc.set_nbfs(n);
  mc.set_D1a(to_mat(D1a, n, n));
  mc.set_D1b(to_mat(D1b, n, n));
  mc.build_tpdm();
  arma::mat d2(mc.get_D2ab());
  std::memcpy(D2ab, d2.memptr(), d2.n_elem * sizeof(double));
}

// All-cores CPU baseline: T threads, each running the staged reference on its own contiguous
// slice of the points with its own MCPDFT object (variant B3; the
// reference's WITH_OPENMP mode is slower than serial).  Partial energies
// are added in slice order.  Returns wall-clock seconds of the parallel region.
double ref_energy_threads(int nthreads, size_t npts, int ncore, int nact, int is_gga,
                          const char* functional, const double* w, const double* phi,
                          const double* phix, const double* phiy, const double* phiz,
                          const double* A1a, const double* A1b, const double* D2act, double eclass,
                          int with_pi_grad, orc_result* out) {
  StdoutMute mute(true);
  const int n = ncore + nact;
  if (nthreads < 1) nthreads = 1;
  std::vector<orc_result> part(nthreads);
  std::vector<std::thread> th;
  // slices need contiguous per-slice column-major copies (the reference API is dense per object)
  std::vector<std::vector<double>> sl_phi(nthreads), sl_x(nthreads), sl_y(nthreads), sl_z(nthreads);
  std::vector<size_t> beg(nthreads + 1);
  for (int t = 0; t <= nthreads; ++t) beg[t] = npts * (size_t)t / nthreads;
  auto slice = [&](const double* src, std::vector<double>& dst, int t) {
    size_t m = beg[t + 1] - beg[t];
    dst.resize(m * n);
    for (int c = 0; c < n; ++c) std::memcpy(&dst[(size_t)c * m], src + (size_t)c * npts + beg[t], m * sizeof(double));
  };
  for (int t = 0; t < nthreads; ++t) {
    slice(phi, sl_phi[t], t);
    if (is_gga) { slice(phix, sl_x[t], t); slice(phiy, sl_y[t], t); slice(phiz, sl_z[t], t); }
  }
  double t0 = now_s();
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() {
      size_t m = beg[t + 1] - beg[t];
      ref_energy_staged(m, ncore, nact, is_gga, functional, w + beg[t], sl_phi[t].data(),
                        is_gga ? sl_x[t].data() : 0, is_gga ? sl_y[t].data() : 0,
                        is_gga ? sl_z[t].data() : 0, A1a, A1b, D2act, 0.0, with_pi_grad, 0, &part[t], 0, 0);
    });
  for (auto& x : th) x.join();
  double dt = now_s() - t0;
  std::memset(out, 0, sizeof(*out));
  for (int t = 0; t < nthreads; ++t) {
    out->ex += part[t].ex; out->ec += part[t].ec;
    out->n_tot += part[t].n_tot; out->n_a += part[t].n_a; out->n_b += part[t].n_b;
    out->trn_tot += part[t].trn_tot; out->trn_a += part[t].trn_a; out->trn_b += part[t].trn_b;
  }
  out->eclass = eclass;
  out->e_tot = eclass + out->ex + out->ec;
  return dt;
}

}  // extern "C"
