/* minimpi — a from-scratch, self-contained subset of MPI-1 sufficient for the
 * 24 symbols miniAMR uses.  MPI is not installed in the
 * build image, so the reference host code (driver/refine/rcb/... unmodified)
 * and the oracle build are compiled against this header instead.
 *
 * Two back-ends implement it (chosen at link time):
 *   minimpi_single.c  one rank, no transport (oracle / single-GPU runs)
 *   minimpi_shm.c     N processes on one node over POSIX shared memory
 *                     (one process per GPU; launched by minimpi/mpirun.py)
 *
 * Only host metadata and (for the pure-CPU reference) host payloads travel
 * through this channel; the GPU build moves ghost faces and migrated blocks
 * with NCCL (miniamr_b200/csrc).
 */
#ifndef MINIMPI_MPI_H
#define MINIMPI_MPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Errhandler;

typedef struct {
   int MPI_SOURCE;
   int MPI_TAG;
   int MPI_ERROR;
   int count_bytes;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_REQUEST_NULL (-1)
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_ERRORS_ARE_FATAL 0
#define MPI_ERRORS_RETURN 1
#define MPI_STATUS_IGNORE ((MPI_Status *) 0)

/* datatype handle == its size in bytes tagged with a kind in the high bits */
#define MINIMPI_KIND_INT 0x100
#define MINIMPI_KIND_DBL 0x200
#define MINIMPI_KIND_LL 0x300
#define MINIMPI_KIND_BYTE 0x400
#define MPI_INT (MINIMPI_KIND_INT | 4)
#define MPI_DOUBLE (MINIMPI_KIND_DBL | 8)
#define MPI_LONG_LONG_INT (MINIMPI_KIND_LL | 8)
#define MPI_LONG_LONG MPI_LONG_LONG_INT
#define MPI_BYTE (MINIMPI_KIND_BYTE | 1)
#define MPI_CHAR MPI_BYTE

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Comm_set_errhandler(MPI_Comm comm, MPI_Errhandler eh);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm);
int MPI_Barrier(MPI_Comm comm);
double MPI_Wtime(void);

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt,
                  MPI_Op op, MPI_Comm comm);
int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf,
                 int rcount, MPI_Datatype rdt, MPI_Comm comm);
int MPI_Scan(const void *sbuf, void *rbuf, int count, MPI_Datatype dt,
             MPI_Op op, MPI_Comm comm);

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag,
             MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int source, int tag,
             MPI_Comm comm, MPI_Status *status);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag,
              MPI_Comm comm, MPI_Request *request);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag,
              MPI_Comm comm, MPI_Request *request);
int MPI_Wait(MPI_Request *request, MPI_Status *status);
int MPI_Waitany(int count, MPI_Request requests[], int *index,
                MPI_Status *status);

#ifdef __cplusplus
}
#endif
#endif
