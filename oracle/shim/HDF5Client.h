// oracle/shim/HDF5Client.h -- TEST INFRASTRUCTURE ONLY.
// Shadows /root/reference/src/libio/HDF5Client.h (HDF5 is not installed here).
// energy.cc:132-140 only round-trips the RDMs through this class and discards the result,
// so a no-op stub leaves the returned energy untouched.
#pragma once
#include <armadillo>
enum H5D_layout_t { H5D_LAYOUT_ERROR = -1, H5D_COMPACT = 0, H5D_CONTIGUOUS = 1, H5D_CHUNKED = 2, H5D_VIRTUAL = 3, H5D_NLAYOUTS = 4 };
namespace mcpdft {
class HDF5Client {
 public:
  enum factory_mode { READ, WRITE };
  void factory_client(H5D_layout_t, factory_mode, const arma::mat&, const arma::mat&, const arma::mat&) {}
};
}  // namespace mcpdft
