// mcpdft_test.cc -- the reference's test driver flow (tests/main.cc:8-49) against the host mirror:
//   mcpdft_test <test_case> <functional>      (cwd must contain ./<test_case>/*.txt)
// Unlike the reference (which prints E - E_ref and returns 0 regardless, tests/main.cc:40-48) a
// third argument <tolerance> makes the exit status depend on |E - E_ref|.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <string>
#include "openrdm_host.hpp"

using namespace mcpdft;

int main(int argc, char* argv[]) {
  if (argc < 3) {
    std::printf("Usage: %s <test_case> <functional> [tolerance]\n", argv[0]);
    return 1;
  }
  try {
    MCPDFT* mc = new MCPDFT(argv[1]);
    double eref = mc->get_eref();
    arma::mat D1a(mc->get_D1a());
    arma::mat D1b(mc->get_D1b());
    mc->build_tpdm();
    arma::mat D2ab(mc->get_D2ab());
    std::string functional(argv[2]);
    double e = mcpdft_energy(mc, functional, D1a, D1b, D2ab);
    std::printf("=================================================\n");
    std::printf("   Reference energy      =  %-20.12lf\n", eref);
    std::printf("   MCPDFT energy         =  %-20.12lf\n", e);
    std::printf("   E(MCPDFT) - E(Ref)    =  %-20.2le\n", e - eref);
    std::printf("=================================================\n\n");
    delete mc;
    if (argc > 3 && !(std::fabs(e - eref) <= std::atof(argv[3]))) return 2;
    return 0;
  } catch (const std::exception& ex) {
    std::fprintf(stderr, "mcpdft_test: %s\n", ex.what());
    return 3;
  }
}
