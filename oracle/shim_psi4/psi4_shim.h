// TEST INFRASTRUCTURE ONLY -- minimal stand-ins for the Psi4 headers that
// /root/reference/src/libpsi4/mcpdft_solver.h and lsda_gga_functionals.cc include, so that the
// reference's own functional routines (EX_B88_I, EX_PBE_I, EC_LYP_I, EC_B88_OP, EC_PBE_I, ...)
// compile UNMODIFIED without a Psi4 installation (Psi4 is not installed in this image).
// Only what those two files name is declared; nothing here computes anything.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

struct iwlbuf {};

namespace psi {

class PsiException : public std::exception {
 public:
  PsiException(const std::string& m, const char*, int) : msg_(m) {}
  const char* what() const noexcept override { return msg_.c_str(); }
 private:
  std::string msg_;
};

class PsiOutStream {
 public:
  void Printf(const char*, ...) {}
};
extern std::shared_ptr<PsiOutStream> outfile;

class Options {
 public:
  std::map<std::string, std::string> str;
  std::map<std::string, double> dbl;
  std::string get_str(const std::string& k) { return str[k]; }
  double get_double(const std::string& k) { return dbl[k]; }
  int get_int(const std::string& k) { return (int)dbl[k]; }
  bool get_bool(const std::string& k) { return dbl[k] != 0.0; }
};

class Vector {
 public:
  explicit Vector(long int n) : v_((size_t)n, 0.0) {}
  double* pointer() { return v_.data(); }
 private:
  std::vector<double> v_;
};
class Matrix;
class VBase;
class PointFunctions;
class BasisSet;
class Molecule;
class JK;
class SuperFunctional;
typedef std::shared_ptr<Matrix> SharedMatrix;
typedef std::shared_ptr<Vector> SharedVector;

class Wavefunction {
 public:
  explicit Wavefunction(Options& o) : options_(o) {}
  virtual ~Wavefunction() {}
 protected:
  Options& options_;
  std::shared_ptr<Wavefunction> reference_wavefunction_;
  bool same_a_b_orbs_ = false, same_a_b_dens_ = false;
};
typedef std::shared_ptr<Wavefunction> SharedWavefunction;

}  // namespace psi
