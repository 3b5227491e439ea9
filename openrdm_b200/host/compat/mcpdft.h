// compat/mcpdft.h -- stands where OpenRDM's src/libmcpdft/mcpdft.h stands (class mcpdft::MCPDFT, mcpdft.h:8-155).
#pragma once
#include <armadillo>
#include "../openrdm_host.hpp"
