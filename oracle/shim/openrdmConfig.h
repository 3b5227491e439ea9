// oracle/shim/openrdmConfig.h -- TEST INFRASTRUCTURE ONLY.
// Stands in for the CMake-generated config header (cmake/openrdmConfig.h.in:1-2).
// Serial build: WITH_OPENMP and WITH_LIBXC undefined = the reference's CMake defaults
// (CMakeLists.txt:86,89).  The Makefile adds -DORDM_REF_OPENMP to flip WITH_OPENMP on.
#pragma once
#ifdef ORDM_REF_OPENMP
#define WITH_OPENMP ON
#endif
