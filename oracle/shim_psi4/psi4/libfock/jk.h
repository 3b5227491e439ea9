// stand-in: see psi4_shim.h
#include "../../psi4_shim.h"
