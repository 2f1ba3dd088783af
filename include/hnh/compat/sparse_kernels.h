// compat forwarding header: lets sources written against the reference (#include "sparse_kernels.h") build
// against the B200-native library.  See INTEGRATION.md.
#pragma once
#include "hnh/sparse_kernels.h"
#include "mpi_standins.h"
