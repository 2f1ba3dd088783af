// compat forwarding header for the reference's drivers (bench_erdos_renyi.cpp includes only
// "benchmark_dist.hpp" and calls MPI_Init / initialize_mpi_datatypes / MPI_Finalize around
// SpmatLocal::loadTuples and benchmark_algorithm): with `-I include/hnh/compat -I include` that file
// compiles UNCHANGED against libhnh_b200.so and runs on the GPUs (one process per GPU, torchrun-style
// RANK / WORLD_SIZE / LOCAL_RANK environment, HNH_NCCL_ID_FILE for the NCCL id).  See INTEGRATION.md.
#pragma once
#include <cassert>
#include <cstdlib>

#include "hnh/benchmark_dist.hpp"
#include "mpi_standins.h"

