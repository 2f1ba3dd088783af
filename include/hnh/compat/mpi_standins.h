// compat/mpi_standins.h -- the few MPI names the reference's DRIVER sources use directly (bench_*.cpp: MPI_Init,
// MPI_Finalize; scratch.cpp also MPI_Comm_rank / MPI_Comm_size / MPI_Allreduce of one double on MPI_COMM_WORLD),
// mapped onto hnh::Comm so that those files compile unchanged.  One process = one rank = one GPU; the world is
// built from the torchrun-style environment (RANK / WORLD_SIZE / LOCAL_RANK, HNH_NCCL_ID_FILE).  Not an MPI
// implementation: library code talks to hnh::Comm, never to these.
#pragma once
#include <cstddef>
#include <stdexcept>

#include "hnh/comm.h"
#include "hnh/runtime.h"

void hnh_world_init_from_env();
void hnh_world_finalize();

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_IN_PLACE ((void *)1)
#define MPI_DOUBLE 1
#define MPI_SUM 1
#define MPI_SUCCESS 0

inline int MPI_Init(int *, char ***) {
    hnh_world_init_from_env();
    return MPI_SUCCESS;
}
inline int MPI_Finalize() {
    hnh_world_finalize();
    return MPI_SUCCESS;
}
inline int MPI_Comm_rank(MPI_Comm, int *rank) {
    *rank = hnh::Comm::world()->rank();
    return MPI_SUCCESS;
}
inline int MPI_Comm_size(MPI_Comm, int *size) {
    *size = hnh::Comm::world()->size();
    return MPI_SUCCESS;
}
// A host barrier in the reference means "every rank has finished the work it issued"; with stream-ordered GPU work
// that includes draining this rank's streams (the reference's benchmark harness stops its wall clock right after
// MPI_Barrier, benchmark_dist.cpp:142-146).
inline int MPI_Barrier(MPI_Comm) {
    if (hnh::Runtime::get().has_device()) hnh::Runtime::get().sync_all();
    hnh::Comm::world()->barrier();
    return MPI_SUCCESS;
}
// in-place sum of host doubles over the world: the only reduction the reference's drivers issue themselves
inline int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
    if (sendbuf != MPI_IN_PLACE || type != MPI_DOUBLE || op != MPI_SUM || comm != MPI_COMM_WORLD)
        throw std::runtime_error("MPI_Allreduce stand-in: only MPI_IN_PLACE / MPI_DOUBLE / MPI_SUM on MPI_COMM_WORLD");
    hnh::Comm::world()->host_allreduce_sum_f64(static_cast<double *>(recvbuf), (size_t)count);
    return MPI_SUCCESS;
}
