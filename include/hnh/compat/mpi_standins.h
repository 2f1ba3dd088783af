// compat/mpi_standins.h -- the few MPI names the reference's DRIVER sources use directly (bench_*.cpp: MPI_Init,
// MPI_Finalize; scratch.cpp also MPI_Comm_rank / MPI_Comm_size / MPI_Allreduce of one double on MPI_COMM_WORLD),
// mapped onto hnh::Comm so that those files compile unchanged.  One process = one rank = one GPU; the world is
// built from the torchrun-style environment (RANK / WORLD_SIZE / LOCAL_RANK, HNH_NCCL_ID_FILE).  Not an MPI
// implementation: library code talks to hnh::Comm, never to these.
#pragma once
#include <chrono>
#include <cstddef>
#include <memory>
#include <stdexcept>

#include "hnh/comm.h"
#include "hnh/runtime.h"

void hnh_world_init_from_env();
void hnh_world_finalize();
// in-place sum over `comm` of `count` doubles at `buf` -- a host variable or device / managed memory (decided from the
// pointer's attributes); returns when the result is in place
void hnh_compat_allreduce_sum_f64(double *buf, size_t count, hnh::Comm &comm);

// A communicator handle IS the library's communicator: `MPI_Comm A_R_split_world = d_ops->A_R_split_world`
// (als_conjugate_gradients.cpp:151-152) then type-checks, and MPI_COMM_WORLD is the process world.
typedef std::shared_ptr<hnh::Comm> MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD (hnh::Comm::world())
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
inline int MPI_Comm_rank(const MPI_Comm &comm, int *rank) {
    *rank = comm->rank();
    return MPI_SUCCESS;
}
inline int MPI_Comm_size(const MPI_Comm &comm, int *size) {
    *size = comm->size();
    return MPI_SUCCESS;
}
inline double MPI_Wtime() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// A host barrier in the reference means "every rank has finished the work it issued"; with stream-ordered GPU work
// that includes draining this rank's streams (the reference's benchmark harness stops its wall clock right after
// MPI_Barrier, benchmark_dist.cpp:142-146).
inline int MPI_Barrier(const MPI_Comm &comm) {
    if (hnh::Runtime::get().has_device()) hnh::Runtime::get().sync_all();
    comm->barrier();
    return MPI_SUCCESS;
}
// In-place sum of doubles: what the reference's drivers and its ALS issue themselves (scratch.cpp,
// als_conjugate_gradients.cpp:33,234).  The buffer may be a host variable or the data() of a VectorXd (HBM / managed).
inline int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, const MPI_Comm &comm) {
    if (sendbuf != MPI_IN_PLACE || type != MPI_DOUBLE || op != MPI_SUM)
        throw std::runtime_error("MPI_Allreduce stand-in: only MPI_IN_PLACE / MPI_DOUBLE / MPI_SUM");
    hnh_compat_allreduce_sum_f64(static_cast<double *>(recvbuf), (size_t)count, *comm);
    return MPI_SUCCESS;
}
