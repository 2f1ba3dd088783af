// compat forwarding header: lets sources written against the reference (#include "25D_cannon_dense.hpp") build
// against the B200-native library.  See INTEGRATION.md.
#pragma once
#include "hnh/25D_cannon_dense.hpp"
#include "mpi_standins.h"
