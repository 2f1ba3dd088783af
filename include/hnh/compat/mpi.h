// compat forwarding header: `#include <mpi.h>` of the reference's sources resolves here when include/hnh/compat is on
// the include path (the real MPI is not needed: see mpi_standins.h).
#pragma once
#include "mpi_standins.h"
