// hnh/benchmark_dist.hpp -- the benchmark harness entry point of the reference
// (benchmark_dist.hpp:8-15, benchmark_dist.cpp:26-167) for the B200-native library.
#pragma once
#include <string>

#include "hnh/SpmatLocal.hpp"
#include "hnh/common.h"
#include "hnh/json.h"

#define MINIMUM_BENCH_TIME 10.0  // kept (and, as in the reference, unused)

// Benchmarks `algorithm_name` in {"15d_fusion1", "15d_fusion2", "15d_sparse",
// "25d_dense_replicate", "25d_sparse_replicate"} on the distributed matrix `spmat`:
// A = B = 0.001, S = 1.0 (benchmark_dist.cpp:102-106), `trials` timed calls of
// fusedSpMM(A, B, S, result, Amat) (fused) or sddmmA + spmmA (unfused), FLOP model
// 2 * nnz * 2 * R per call (:147).  Appends the JSON record + "," to output_file on rank 0.
// The reference hard-codes 5 trials and no warm-up; benchmark_algorithm() keeps that signature
// and behaviour, benchmark_algorithm_ex() adds warm-up calls and returns the record.
void benchmark_algorithm(SpmatLocal *spmat, string algorithm_name, string output_file, bool fused, int R, int c,
                         string app);
json benchmark_algorithm_ex(SpmatLocal *spmat, string algorithm_name, string output_file, bool fused, int R, int c,
                            string app, int trials, int warmup);

// Process-world bootstrap for stand-alone C++ drivers: reads RANK / WORLD_SIZE /
// LOCAL_RANK (torchrun conventions) and HNH_NCCL_ID_FILE (a path on a shared filesystem where
// rank 0 publishes the NCCL unique id).  World size 1 needs nothing.  These stand in for
// MPI_Init / MPI_Finalize of bench_erdos_renyi.cpp:20,120.
void hnh_world_init_from_env();
void hnh_world_finalize();
