// compat forwarding header: lets sources written against the reference (#include "15D_dense_shift.hpp") build
// against the B200-native library.  See INTEGRATION.md.
#pragma once
#include "hnh/15D_dense_shift.hpp"
#include "mpi_standins.h"
