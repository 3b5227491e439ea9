// TEST INFRASTRUCTURE ONLY.
// C-ABI harness around the Psi4 plugin's OWN exchange-correlation routines
// (/root/reference/src/libpsi4/lsda_gga_functionals.cc, compiled unmodified against the stand-in
// Psi4 headers in shim_psi4/): the extended functional list of SURVEY.md section 8f rank 3 --
// revPBE (EX_PBE_I with kappa = 1.245, lsda_gga_functionals.cc:323), B88 exchange (EX_B88_I,
// :207-273), LYP correlation (EC_LYP_I, :764-829), B88-OP correlation (EC_B88_OP, :478-518) and the
// plugin's PBE pair (EX_PBE_I / EC_PBE_I) as dispatched in mcpdft_solver.cc:742-787; plus PW92 LSDA
// correlation (EC_PW92_I, :997-1093) and the range-separated omega-PBE exchange (EX_wPBE_I,
// rs_functionals.cc:86-300).
#define private public
#define protected public
#include "mcpdft_solver.h"
#undef private
#undef protected
#include <cstring>

namespace psi {
std::shared_ptr<PsiOutStream> outfile = std::make_shared<PsiOutStream>();
namespace mcpdft {
// the plugin's constructor/destructor live in mcpdft_solver.cc (needs Psi4 proper); the harness only
// needs an object whose phi_points_, grid_w_ and options_ are set
MCPDFTSolver::MCPDFTSolver(std::shared_ptr<psi::Wavefunction>, Options& options) : Wavefunction(options) {}
MCPDFTSolver::~MCPDFTSolver() {}
}  // namespace mcpdft
}  // namespace psi

using psi::Vector;
static std::shared_ptr<Vector> vec(size_t n, const double* src) {
  auto v = std::make_shared<Vector>((long int)n);
  if (src) std::memcpy(v->pointer(), src, n * sizeof(double));
  return v;
}

enum { PSI4_EX_LSDA = 0, PSI4_EC_VWN3_RPA_III, PSI4_EX_PBE_I, PSI4_EC_PBE_I, PSI4_EX_REVPBE_I, PSI4_EX_B88_I, PSI4_EC_LYP_I, PSI4_EC_B88_OP,
       PSI4_EX_WPBE_I, PSI4_EC_PW92_I };

static double g_omega = 0.0;  // MCPDFT_OMEGA (plugin.cc:50), read by EX_wPBE_I (rs_functionals.cc:89)
extern "C" void ref_psi4_set_omega(double omega) { g_omega = omega; }

extern "C" double ref_psi4_functional(int which, size_t npts, const double* w, const double* ra, const double* rb,
                                      const double* saa, const double* sab, const double* sbb) {
  psi::Options opt;
  opt.str["MCPDFT_FUNCTIONAL"] = (which == PSI4_EX_REVPBE_I) ? "REVPBE" : "PBE";
  opt.dbl["MCPDFT_OMEGA"] = g_omega;
  psi::mcpdft::MCPDFTSolver s(nullptr, opt);
  s.phi_points_ = (long int)npts;
  s.grid_w_ = vec(npts, w);
  auto A = vec(npts, ra), B = vec(npts, rb), AA = vec(npts, saa), AB = vec(npts, sab), BB = vec(npts, sbb);
  switch (which) {
    case PSI4_EX_LSDA: return s.EX_LSDA(A, B);
    case PSI4_EC_VWN3_RPA_III: return s.EC_VWN3_RPA_III(A, B);
    case PSI4_EX_PBE_I:
    case PSI4_EX_REVPBE_I: return s.EX_PBE_I(A, B, AA, BB);
    case PSI4_EC_PBE_I: return s.EC_PBE_I(A, B, AA, AB, BB);
    case PSI4_EX_B88_I: return s.EX_B88_I(A, B, AA, BB);
    case PSI4_EC_LYP_I: return s.EC_LYP_I(A, B, AA, AB, BB);
    case PSI4_EC_B88_OP: return s.EC_B88_OP(A, B, AA, BB);
    case PSI4_EX_WPBE_I: return s.EX_wPBE_I(A, B, AA, BB);
    case PSI4_EC_PW92_I: return s.EC_PW92_I(A, B);
  }
  return 0.0 / 0.0;
}
