// compat forwarding header: the reference's `#include "gat.hpp"` resolves here (see benchmark_dist.hpp).
#pragma once
#include "hnh/gat.hpp"
#include "mpi_standins.h"
