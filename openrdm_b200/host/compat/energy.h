// compat/energy.h -- stands where OpenRDM's src/libmcpdft/energy.h stands (mcpdft::mcpdft_energy, energy.h:14-18;
// it also pulls in <iostream>, <fstream>, <string>, which the reference's callers rely on, energy.h:4-7).
#pragma once
#include <armadillo>
#include <fstream>
#include <iostream>
#include <string>
#include "mcpdft.h"
