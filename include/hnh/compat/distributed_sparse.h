// compat forwarding header: lets sources written against the reference (#include "distributed_sparse.h") build
// against the B200-native library.  See INTEGRATION.md.
#pragma once
#include "hnh/distributed_sparse.h"
#include "mpi_standins.h"
