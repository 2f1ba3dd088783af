// compat forwarding header: lets sources written against the reference (#include "als_conjugate_gradients.h") build
// against the B200-native library.  See INTEGRATION.md.
#pragma once
#include "hnh/als_conjugate_gradients.h"
#include "mpi_standins.h"
