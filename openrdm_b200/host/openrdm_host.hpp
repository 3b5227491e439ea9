// openrdm_host.hpp -- C++ host mirror of OpenRDM's libmcpdft / libfunc operator surface on top of
// the C ABI (include/openrdm_b200.h).  Same namespace, class and method names, argument meaning
// and call order as the reference, so code written against
//     mcpdft::MCPDFT          /root/reference/src/libmcpdft/mcpdft.h:8-155
//     mcpdft::mcpdft_energy   src/libmcpdft/energy.h:14-18
//     mcpdft::Functional      src/libfunc/functional.h:9-44
// re-links against this header + libopenrdm_host + libopenrdm_b200 unchanged (see
// INTEGRATION.md).  All numerics run on the GPU through the C ABI; this layer only keeps the
// host copies the reference's getters/setters expose and moves data.  Errors (which the
// reference mostly swallows, e.g. utility.cc:13-14) are thrown as std::runtime_error.
#pragma once
#include <string>
#include "arma_lite.hpp"

struct ordm_ctx;

namespace mcpdft {

class MCPDFT {
 public:
  MCPDFT();                               // mcpdft.cc:16
  explicit MCPDFT(std::string test_case);  // mcpdft.cc:15: reads ./<test_case>/*.txt (utility.cc)
  ~MCPDFT();
  MCPDFT(const MCPDFT&) = delete;
  MCPDFT& operator=(const MCPDFT&) = delete;

  void common_init(std::string test_case);  // mcpdft.cc:19-33 (GGA iff test_case == "h2_tpbe_sto3g")
  void build_opdm();                         // mcpdft.cc:608-638
  void build_tpdm();                         // mcpdft.cc:640-652: D2ab = kron(D1a, D1b)
  void build_rho();                          // mcpdft.cc:35-40
  void build_density_functions();            // mcpdft.cc:42-105
  void build_density_gradients();            // mcpdft.cc:107-176
  void build_pi(const arma::mat& D2ab);      // mcpdft.cc:178-183
  void build_ontop_pair_density(const arma::mat& D2ab);            // mcpdft.cc:185-212
  void build_ontop_pair_density_gradients(const arma::mat& D2ab);  // mcpdft.cc:214-270
  void build_R();                            // mcpdft.cc:272-285
  void translate();                          // mcpdft.cc:287-292
  void translate_density();                  // mcpdft.cc:294-344
  void translate_density_gradients();        // mcpdft.cc:346-416
  void fully_translate();                    // mcpdft.cc:418-423
  void fully_translate_density();            // mcpdft.cc:425-488
  void fully_translate_density_gradients();  // mcpdft.cc:490-606

  // ---- accessors (mcpdft.h:63-164); getters return copies, setters copy in, as the reference
  bool is_gga() const { return is_gga_; }
  std::size_t get_npts() const { return npts_; }
  int get_nbfs() const { return (int)nbfs_; }
  double get_eref() const { return eref_; }
  double get_eclass() const { return eclass_; }
  void set_npts(const std::size_t npts) { npts_ = npts; }
  void set_nbfs(const int nbfs) { nbfs_ = (std::size_t)nbfs; }
  void set_eref(const double e) { eref_ = e; }
  void set_eclass(const double e) { eclass_ = e; }
#define ORDM_HOST_INPUT(T, name, member) \
  T get_##name() const { return member; }  \
  void set_##name(const T& v) { member = v; dirty_inputs_ = true; }
  ORDM_HOST_INPUT(arma::vec, w, w_) ORDM_HOST_INPUT(arma::vec, x, x_) ORDM_HOST_INPUT(arma::vec, y, y_) ORDM_HOST_INPUT(arma::vec, z, z_)
  ORDM_HOST_INPUT(arma::mat, phi, phi_) ORDM_HOST_INPUT(arma::mat, phi_x, phi_x_) ORDM_HOST_INPUT(arma::mat, phi_y, phi_y_)
  ORDM_HOST_INPUT(arma::mat, phi_z, phi_z_) ORDM_HOST_INPUT(arma::mat, cmat, cmat_) ORDM_HOST_INPUT(arma::mat, D1a, D1a_)
  ORDM_HOST_INPUT(arma::mat, D1b, D1b_) ORDM_HOST_INPUT(arma::mat, D2ab, D2ab_)
#undef ORDM_HOST_INPUT
  // per-point results live on the device; each getter copies npts doubles back (mcpdft.cc:672-690), each setter
  // (mcpdft.cc:708-736) installs the caller's vector on the device (ordm_set_vector), where the following stages
  // read it exactly where the reference reads its members: rho, Pi and R by build_R / translate, the spin-density
  // gradients by translate_density_gradients, grad Pi by fully_translate
#define ORDM_HOST_OUTPUT(name) arma::vec get_##name() const; void set_##name(const arma::vec& v);
  ORDM_HOST_OUTPUT(rhoa) ORDM_HOST_OUTPUT(rhoa_x) ORDM_HOST_OUTPUT(rhoa_y) ORDM_HOST_OUTPUT(rhoa_z)
  ORDM_HOST_OUTPUT(rhob) ORDM_HOST_OUTPUT(rhob_x) ORDM_HOST_OUTPUT(rhob_y) ORDM_HOST_OUTPUT(rhob_z)
  ORDM_HOST_OUTPUT(rho) ORDM_HOST_OUTPUT(tr_rhoa) ORDM_HOST_OUTPUT(tr_rhob) ORDM_HOST_OUTPUT(pi) ORDM_HOST_OUTPUT(R)
  ORDM_HOST_OUTPUT(sigma_aa) ORDM_HOST_OUTPUT(sigma_ab) ORDM_HOST_OUTPUT(sigma_bb)
  ORDM_HOST_OUTPUT(tr_sigma_aa) ORDM_HOST_OUTPUT(tr_sigma_ab) ORDM_HOST_OUTPUT(tr_sigma_bb)
  // the reference has setters but no getters for grad Pi (mcpdft.h:139-141) and never stores pi_y
  // (mcpdft.cc:268); here all three components are retrievable
  ORDM_HOST_OUTPUT(pi_x) ORDM_HOST_OUTPUT(pi_y) ORDM_HOST_OUTPUT(pi_z)
#undef ORDM_HOST_OUTPUT
  // translated total density and translated spin-density gradients: the reference keeps members with setters
  // (mcpdft.cc:718-725) that none of its stages reads or writes (translate_density_gradients works on locals,
  // mcpdft.cc:350-372); get_tr_rho returns what set_tr_rho stored (mcpdft.cc:682), the gradient getters are
  // declared but never defined there (mcpdft.h:92-98) and are defined here
#define ORDM_HOST_PLAIN(name) arma::vec get_##name() const { return name##_; } void set_##name(const arma::vec& v) { name##_ = v; }
  ORDM_HOST_PLAIN(tr_rho) ORDM_HOST_PLAIN(tr_rhoa_x) ORDM_HOST_PLAIN(tr_rhoa_y) ORDM_HOST_PLAIN(tr_rhoa_z)
  ORDM_HOST_PLAIN(tr_rhob_x) ORDM_HOST_PLAIN(tr_rhob_y) ORDM_HOST_PLAIN(tr_rhob_z)
#undef ORDM_HOST_PLAIN
  void print_banner() const;

  // ---- extensions (no reference counterpart) ----
  void set_gga(bool gga) { is_gga_ = gga; dirty_inputs_ = true; }   // the reference keys GGA on the directory name
  // mcpdft_energy() uses fully_translate() instead of translate() (the reference hard-wires translate(), energy.cc:58)
  void set_fully_translated(bool ft);
  // build grad Pi in build_pi()/mcpdft_energy() for GGA cases.  The reference always runs that stage
  // (mcpdft.cc:181-183) but nothing can observe its result (no getters, unused by translate()), so
  // it is off by default here; fully_translate() switches it on by itself.
  void set_pi_gradient(bool on);
  void set_device(int device);                                     // before the first compute call; default 0
  void integrated_densities(double n[3], double trn[3]) const;     // what mcpdft.cc:100-104,339-343 print
  ordm_ctx* context() const { return ctx_; }
  void sync_inputs() const;  // upload host-side inputs if a setter changed them (used by mcpdft_energy)

 protected:  // text fixture readers, utility.cc:5-183
  void read_grids_from_file(std::string test_case);
  void read_orbitals_from_file(std::string test_case);
  void read_gradients_from_file(std::string test_case);
  void read_energies_from_file(std::string test_case);
  void read_opdm_from_file(std::string test_case);
  void read_cmat_from_file(std::string test_case);

 private:
  arma::vec fetch(int which) const;
  void install(int which, const arma::vec& v);
  void ensure_ctx() const;
  arma::vec tr_rho_, tr_rhoa_x_, tr_rhoa_y_, tr_rhoa_z_, tr_rhob_x_, tr_rhob_y_, tr_rhob_z_;
  bool is_gga_ = false;
  std::size_t npts_ = 0, nbfs_ = 0;
  double eref_ = 0.0, eclass_ = 0.0;
  arma::vec w_, x_, y_, z_;
  arma::mat phi_, phi_x_, phi_y_, phi_z_, cmat_, D1a_, D1b_, D2ab_;
  mutable ordm_ctx* ctx_ = nullptr;
  mutable bool dirty_inputs_ = true;
  int device_ = 0;
  unsigned opts_ = 1u;  // ORDM_OPT_DIAGNOSTICS
};

// energy.h:14-18.  As in the reference, the densities come from mc's own D1a/D1b members
// (energy.cc:49); the D1a/D1b arguments are accepted and ignored for the physics, D2ab is used.
double mcpdft_energy(MCPDFT* mc, std::string& functional, const arma::mat& D1a, const arma::mat& D1b,
                     const arma::mat& D2ab);

class Functional {  // functional.h:9-44: vectors in, quadrature-summed energy out (weights from mc)
 public:
  Functional();    // functional.cc:9
  ~Functional();   // functional.cc:10
  double EX_LSDA(const MCPDFT* mc, const arma::vec& rho_a, const arma::vec& rho_b);
  double EX_PBE(const MCPDFT* mc, const arma::vec& rho_a, const arma::vec& rho_b, const arma::vec& sigma_aa,
                const arma::vec& sigma_bb);
  double EC_VWN3(const MCPDFT* mc, const arma::vec& rho_a, const arma::vec& rho_b);
  double EC_PBE(const MCPDFT* mc, const arma::vec& rho_a, const arma::vec& rho_b, const arma::vec& sigma_aa,
                const arma::vec& sigma_ab, const arma::vec& sigma_bb);
};

}  // namespace mcpdft
