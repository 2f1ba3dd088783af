// compat forwarding header: lets sources written against the reference (#include "SpmatLocal.hpp") build
// against the B200-native library.  See INTEGRATION.md.
#pragma once
#include "hnh/SpmatLocal.hpp"
#include "mpi_standins.h"
