// oracle/shims/mpi.h -- TEST INFRASTRUCTURE.  "MPI ranks as threads": the subset of the MPI C API
// that the reference's sources call, implemented over std::thread + shared memory
// (oracle/shims/hmpi.cpp), so that /root/reference compiles and runs UNMODIFIED in oracle/_ref on
// a box without any MPI.  Every rank is a thread of one process; hmpi_run(p, fn) starts p of
// them.  Semantics follow the MPI standard for the calls used (blocking collectives, eager
// buffered point-to-point with tag and MPI_ANY_SOURCE matching, rank-ordered reductions).
#pragma once
#include <cstddef>
#include <cstdint>

struct hmpi_comm;
typedef hmpi_comm *MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef long MPI_Aint;
struct MPI_Status {
    int MPI_SOURCE, MPI_TAG, MPI_ERROR;
};
struct hmpi_request;
typedef hmpi_request *MPI_Request;

#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_IN_PLACE ((void *)-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_SUM 1
// datatype = id; sizes are looked up in hmpi.cpp (struct types get ids >= 100)
#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_LONG 3
#define MPI_DOUBLE 4
#define MPI_UINT64_T 5
#define MPI_UNSIGNED_LONG 6
#define MPI_FLOAT 7

MPI_Comm hmpi_world();
#define MPI_COMM_WORLD (hmpi_world())

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize();
double MPI_Wtime();
int MPI_Comm_rank(MPI_Comm c, int *rank);
int MPI_Comm_size(MPI_Comm c, int *size);
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *out);
int MPI_Comm_dup(MPI_Comm c, MPI_Comm *out);
int MPI_Comm_free(MPI_Comm *c);
int MPI_Barrier(MPI_Comm c);
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Gather(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Allgather(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, MPI_Comm c);
int MPI_Allgatherv(const void *s, int sc, MPI_Datatype st, void *r, const int *rcounts, const int *displs,
                   MPI_Datatype rt, MPI_Comm c);
int MPI_Alltoall(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, MPI_Comm c);
int MPI_Alltoallv(const void *s, const int *scounts, const int *sdispls, MPI_Datatype st, void *r, const int *rcounts,
                  const int *rdispls, MPI_Datatype rt, MPI_Comm c);
int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Reduce_scatter(const void *s, void *r, const int *rcounts, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Send(const void *buf, int count, MPI_Datatype t, int dst, int tag, MPI_Comm c);
int MPI_Recv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st);
int MPI_Sendrecv(const void *s, int sc, MPI_Datatype st, int dst, int stag, void *r, int rc, MPI_Datatype rt, int src,
                 int rtag, MPI_Comm c, MPI_Status *status);
int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *st);
int MPI_Type_create_struct(int n, const int *blocklens, const MPI_Aint *offsets, const MPI_Datatype *types,
                           MPI_Datatype *out);
int MPI_Type_commit(MPI_Datatype *t);

// launcher: runs fn(rank, arg) on p rank-threads that share one MPI_COMM_WORLD; returns when all
// have finished.  threads_per_rank sets each rank's OpenMP team size (0 = leave alone).
void hmpi_run(int p, int threads_per_rank, void (*fn)(int rank, void *arg), void *arg);
