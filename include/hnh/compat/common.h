// compat forwarding header: lets sources written against the reference (#include "common.h") build
// against the B200-native library.  See INTEGRATION.md.
#pragma once
#include "hnh/common.h"
#include "mpi_standins.h"
