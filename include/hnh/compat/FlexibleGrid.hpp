// compat forwarding header: lets sources written against the reference (#include "FlexibleGrid.hpp") build
// against the B200-native library.  See INTEGRATION.md.
#pragma once
#include "hnh/FlexibleGrid.hpp"
#include "mpi_standins.h"
