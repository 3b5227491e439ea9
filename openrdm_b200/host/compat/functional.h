// compat/functional.h -- stands where OpenRDM's src/libfunc/functional.h stands (class mcpdft::Functional, functional.h:9-44).
#pragma once
#include <armadillo>
#include "mcpdft.h"
