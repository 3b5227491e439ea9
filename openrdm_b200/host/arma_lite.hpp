// arma_lite.hpp -- the sliver of Armadillo's surface that OpenRDM's libmcpdft/libfunc API uses
// (column-major FP64 arma::mat / arma::vec with (i,j)/(i) access, n_rows/n_cols/n_elem,
// memptr(), fill::zeros, kron).  If the real Armadillo is installed, define
// ORDM_USE_ARMADILLO and it is used instead; the host mirror below only needs this much.
#pragma once
#if defined(ORDM_USE_ARMADILLO)
#include <armadillo>
#else
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
namespace arma {
namespace fill { struct zeros_t {}; static const zeros_t zeros{}; }
class vec;
class mat {
 public:
  std::size_t n_rows = 0, n_cols = 0, n_elem = 0;
  mat() = default;
  mat(std::size_t r, std::size_t c) : n_rows(r), n_cols(c), n_elem(r * c), v_(r * c, 0.0) {}
  mat(std::size_t r, std::size_t c, fill::zeros_t) : mat(r, c) {}
  inline mat(const vec& v);   // column vector -> n x 1 matrix, as Armadillo converts (utility.cc:31 relies on it)
  double& operator()(std::size_t i, std::size_t j) { return v_[i + j * n_rows]; }
  double operator()(std::size_t i, std::size_t j) const { return v_[i + j * n_rows]; }
  double* memptr() { return v_.data(); }
  const double* memptr() const { return v_.data(); }
  void print(const char* tag = "") const {
    std::printf("%s\n", tag);
    for (std::size_t i = 0; i < n_rows; ++i) { for (std::size_t j = 0; j < n_cols; ++j) std::printf(" %14.8f", (*this)(i, j)); std::printf("\n"); }
  }
 private:
  std::vector<double> v_;
};
class vec {
 public:
  std::size_t n_elem = 0, n_rows = 0, n_cols = 1;
  vec() = default;
  explicit vec(std::size_t n) : n_elem(n), n_rows(n), v_(n, 0.0) {}
  vec(std::size_t n, fill::zeros_t) : vec(n) {}
  double& operator()(std::size_t i) { return v_[i]; }
  double operator()(std::size_t i) const { return v_[i]; }
  double* memptr() { return v_.data(); }
  const double* memptr() const { return v_.data(); }
  void print(const char* tag = "") const {
    std::printf("%s\n", tag);
    for (std::size_t i = 0; i < n_elem; ++i) std::printf(" %14.8f\n", v_[i]);
  }
 private:
  std::vector<double> v_;
};
inline mat::mat(const vec& v) : n_rows(v.n_elem), n_cols(1), n_elem(v.n_elem), v_(v.memptr(), v.memptr() + v.n_elem) {}
inline vec operator+(const vec& a, const vec& b) {
  vec r(a.n_elem);
  for (std::size_t i = 0; i < a.n_elem; ++i) r(i) = a(i) + b(i);
  return r;
}
inline mat kron(const mat& A, const mat& B) {
  mat K(A.n_rows * B.n_rows, A.n_cols * B.n_cols);
  for (std::size_t j = 0; j < A.n_cols; ++j) for (std::size_t i = 0; i < A.n_rows; ++i)
    for (std::size_t l = 0; l < B.n_cols; ++l) for (std::size_t k = 0; k < B.n_rows; ++k)
      K(i * B.n_rows + k, j * B.n_cols + l) = A(i, j) * B(k, l);
  return K;
}
}  // namespace arma
#endif
