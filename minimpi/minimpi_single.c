/* minimpi single-rank back-end: collectives are copies, point-to-point is a
 * protocol error (a one-rank miniAMR never posts one). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "mpi.h"

static size_t dt_size(MPI_Datatype dt) { return (size_t)(dt & 0xff); }

static int p2p_error(const char *what)
{
   fprintf(stderr, "minimpi(single): %s called with one rank\n", what);
   exit(-1);
   return -1;
}

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm comm, int errorcode) { (void)comm; exit(errorcode ? errorcode : -1); return 0; }
int MPI_Comm_set_errhandler(MPI_Comm comm, MPI_Errhandler eh) { (void)comm; (void)eh; return 0; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = 0; return 0; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = 1; return 0; }
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm)
{ (void)color; (void)key; *newcomm = comm + 1; return 0; }
int MPI_Barrier(MPI_Comm comm) { (void)comm; return 0; }
double MPI_Wtime(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double)ts.tv_sec + 1.0e-9*(double)ts.tv_nsec;
}
int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm)
{ (void)buf; (void)count; (void)dt; (void)root; (void)comm; return 0; }
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt,
                  MPI_Op op, MPI_Comm comm)
{ (void)op; (void)comm; memmove(rbuf, sbuf, (size_t)count*dt_size(dt)); return 0; }
int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf,
                 int rcount, MPI_Datatype rdt, MPI_Comm comm)
{ (void)rcount; (void)rdt; (void)comm; memmove(rbuf, sbuf, (size_t)scount*dt_size(sdt)); return 0; }
int MPI_Scan(const void *sbuf, void *rbuf, int count, MPI_Datatype dt,
             MPI_Op op, MPI_Comm comm)
{ (void)op; (void)comm; memmove(rbuf, sbuf, (size_t)count*dt_size(dt)); return 0; }
int MPI_Send(const void *b, int c, MPI_Datatype d, int dest, int tag, MPI_Comm comm)
{ (void)b; (void)c; (void)d; (void)dest; (void)tag; (void)comm; return p2p_error("MPI_Send"); }
int MPI_Recv(void *b, int c, MPI_Datatype d, int s, int tag, MPI_Comm comm, MPI_Status *st)
{ (void)b; (void)c; (void)d; (void)s; (void)tag; (void)comm; (void)st; return p2p_error("MPI_Recv"); }
int MPI_Isend(const void *b, int c, MPI_Datatype d, int dest, int tag, MPI_Comm comm, MPI_Request *r)
{ (void)b; (void)c; (void)d; (void)dest; (void)tag; (void)comm; (void)r; return p2p_error("MPI_Isend"); }
int MPI_Irecv(void *b, int c, MPI_Datatype d, int s, int tag, MPI_Comm comm, MPI_Request *r)
{ (void)b; (void)c; (void)d; (void)s; (void)tag; (void)comm; (void)r; return p2p_error("MPI_Irecv"); }
int MPI_Wait(MPI_Request *r, MPI_Status *st) { (void)st; *r = MPI_REQUEST_NULL; return 0; }
int MPI_Waitany(int count, MPI_Request r[], int *index, MPI_Status *st)
{ (void)count; (void)r; (void)st; *index = MPI_UNDEFINED; return 0; }
