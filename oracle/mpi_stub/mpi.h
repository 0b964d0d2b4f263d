/*
 * mpi.h -- an in-process / shared-memory stand-in for the MPI the reference suite is written against.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/): this image has no MPI, and without MPI the reference compiles its
 * Comm_HALO_EXCHANGE / HALO_EXCHANGE_FUSED / HALO_SENDRECV kernels out (RAJAPerfSuite.hpp:177-181).  Linking the
 * UNMODIFIED reference sources against this stub (oracle/build_ref_mpi.sh) gives a reference binary that runs those
 * kernels: on one rank, where every one of the 26 neighbours of the periodic rank grid is the rank itself, and on P ranks
 * -- P processes of the binary over a shared-memory arena (RPB_MPI_SIZE / RPB_MPI_RANK / RPB_MPI_SHM, see mpi_stub.c;
 * the launcher is mpirun() in tests/golden/make_golden.py).  make_golden.py --mpi mints the exchange goldens from it.
 *
 * Only what the suite calls (grep MPI_ src/): Init/Finalize, Comm_rank/size, Barrier, Allreduce, Gather, Bcast (and
 * Allgather, for the IPC-handle rendezvous of the Base_B200 exchange stub),
 * Isend/Irecv, Wait/Waitall/Waitany.  A message is matched to a receive by (source, tag), oldest first, as MPI does;
 * sends are buffered, so any call order the standard allows works.
 */
#ifndef RPB_ORACLE_MPI_STUB_H
#define RPB_ORACLE_MPI_STUB_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_UNDEFINED (-32766)
#define MPI_REQUEST_NULL (-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

#define MPI_DATATYPE_NULL 0
#define MPI_BYTE 1
#define MPI_CHAR 2
#define MPI_INT 3
#define MPI_LONG 4
#define MPI_LONG_LONG 5
#define MPI_UNSIGNED_LONG 6
#define MPI_FLOAT 7
#define MPI_DOUBLE 8
#define MPI_LONG_DOUBLE 9

#define MPI_OP_NULL 0
#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Barrier(MPI_Comm comm);
double MPI_Wtime(void);
int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Gather(const void* send, int scount, MPI_Datatype stype, void* recv, int rcount, MPI_Datatype rtype,
               int root, MPI_Comm comm);
int MPI_Allgather(const void* send, int scount, MPI_Datatype stype, void* recv, int rcount, MPI_Datatype rtype, MPI_Comm comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Wait(MPI_Request* req, MPI_Status* status);
int MPI_Waitall(int count, MPI_Request* reqs, MPI_Status* statuses);
int MPI_Waitany(int count, MPI_Request* reqs, int* index, MPI_Status* status);

#ifdef __cplusplus
}
#endif
#endif
