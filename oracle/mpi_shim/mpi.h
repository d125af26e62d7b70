/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Minimal "threads-as-ranks" MPI stand-in used solely to compile the UNMODIFIED reference
 * (icl-utk-edu/heffte, read in place from /root/reference) into oracle/_ref/ so that its stock
 * CPU backend can act as the parity oracle and as the CPU baseline.  The container has no MPI
 * (no mpi.h / mpicc / mpirun), and the reference does `#include <mpi.h>`
 * (reference: include/heffte_utils.h:25) and `find_package(MPI REQUIRED)` (CMakeLists.txt:110).
 *
 * Surface = exactly the symbols the reference's include/, src/, test/test_common.h and
 * benchmarks/ use.  Every rank is a std::thread of one process;
 * MPI_COMM_WORLD has SHIM_NP ranks (env var, default 1).
 */
#ifndef ORACLE_MPI_SHIM_H
#define ORACLE_MPI_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

struct shim_comm_s;
typedef struct shim_comm_s* MPI_Comm;
struct shim_group_s;
typedef struct shim_group_s* MPI_Group;
struct shim_request_s;
typedef struct shim_request_s* MPI_Request;

typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL ((MPI_Comm)0)
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_TAG (-1)

/* datatypes: the value is the size in bytes (all the shim needs) tagged in the high bits */
#define MPI_BYTE             ((MPI_Datatype)0x0101)
#define MPI_INT              ((MPI_Datatype)0x0204)
#define MPI_FLOAT            ((MPI_Datatype)0x0304)
#define MPI_DOUBLE           ((MPI_Datatype)0x0408)
#define MPI_C_COMPLEX        ((MPI_Datatype)0x0508)
#define MPI_C_DOUBLE_COMPLEX ((MPI_Datatype)0x0610)
#define MPI_LONG_LONG        ((MPI_Datatype)0x0708)

#define MPI_MAX ((MPI_Op)1)
#define MPI_SUM ((MPI_Op)2)
#define MPI_MIN ((MPI_Op)3)

MPI_Comm shim_comm_world(void);
#define MPI_COMM_WORLD (shim_comm_world())

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
double MPI_Wtime(void);

int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group);
int MPI_Group_incl(MPI_Group group, int n, const int ranks[], MPI_Group *newgroup);
int MPI_Group_free(MPI_Group *group);
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *newcomm);

int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                  void *recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Alltoall(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                 void *recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Alltoallv(const void *sendbuf, const int sendcounts[], const int sdispls[], MPI_Datatype sendtype,
                  void *recvbuf, const int recvcounts[], const int rdispls[], MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *request);
int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request *request);
int MPI_Waitany(int count, MPI_Request requests[], int *index, MPI_Status *status);
int MPI_Waitall(int count, MPI_Request requests[], MPI_Status statuses[]);

/* shim-only entry points (not MPI): run `fn(arg)` on `nranks` threads acting as the ranks of MPI_COMM_WORLD */
int shim_run(int nranks, int (*fn)(int argc, char **argv), int argc, char **argv);
int shim_world_size_from_env(void);

#ifdef __cplusplus
}
#endif

#endif
