/* minimpi multi-process back-end: N ranks on one node over POSIX shared memory.
 *
 * Written from scratch for this repository (MPI is not installed in the image):
 * it implements exactly the subset of MPI-1 that miniAMR's host code uses
 *  so that the UNMODIFIED reference runs at N ranks -- as
 * the multi-rank CPU oracle / baseline, and as the host side of the GPU build
 * (one rank per GPU; ghost faces and block payloads of that build travel by NCCL,
 * only host metadata goes through here).
 *
 * Transport: one single-producer/single-consumer byte ring per ordered pair of
 * ranks in a shared segment created by minimpi/minimpirun.c.  A message is a
 * 16-byte frame header (tag, communicator, length) followed by the payload;
 * messages larger than the ring stream through it.  Every blocking call runs the
 * progress engine: push the pending sends of every destination as far as ring
 * space allows, and drain every incoming ring completely -- into the matching
 * posted receive, or into a private "unexpected message" buffer.  Draining never
 * blocks, so a rank inside any MPI call always frees ring space for its peers:
 * non-blocking sends complete without the receiver having posted anything, and
 * pairwise non-overtaking order follows from the rings being FIFO.
 *
 * Collectives are built on the same point-to-point layer in a private context
 * (communicator id | CTX_COLL), with reductions evaluated in rank order at the
 * root so that every rank sees the same bits.
 *
 * Environment (set by minimpirun): MINIMPI_SHM (segment name), MINIMPI_RANK,
 * MINIMPI_SIZE.  Without them the library runs as a single rank.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <signal.h>
#include <stdarg.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include "mpi.h"
#include "minimpi_shm.h"

/* ---- shared segment ------------------------------------------------------- */

typedef struct {
   _Atomic uint64_t head;            /* bytes ever written (producer) */
   char pad0[56];
   _Atomic uint64_t tail;            /* bytes ever consumed (consumer) */
   char pad1[56];
} ring_hdr;

static minimpi_seg *seg;
static size_t ring_bytes;            /* capacity of one ring, power of two */
static int world_size = 1, world_rank = 0;
static int initialised;

static ring_hdr *ring_of(int src, int dst)
{
   size_t stride = sizeof(ring_hdr) + ring_bytes;
   return (ring_hdr *) ((char *) seg + seg->rings_off + ((size_t) src*world_size + dst)*stride);
}
static char *ring_data(ring_hdr *r) { return (char *) r + sizeof(ring_hdr); }

static void fatal(const char *fmt, ...)
{
   va_list ap;
   fprintf(stderr, "minimpi[%d]: ", world_rank);
   va_start(ap, fmt);
   vfprintf(stderr, fmt, ap);
   va_end(ap);
   fprintf(stderr, "\n");
   fflush(stderr);
   if (seg) atomic_store(&seg->abort_flag, 1);
   _exit(86);
}

/* ---- communicators -------------------------------------------------------- */

#define MAX_COMMS 256
#define CTX_COLL 0x4000              /* private context of the collectives */
typedef struct { int size, rank, *world; } comm_t;
static comm_t comms_tab[MAX_COMMS];
static int next_comm = 1;

static comm_t *comm_of(MPI_Comm c)
{
   if (c < 0 || c >= MAX_COMMS || comms_tab[c].size == 0) fatal("invalid communicator %d", c);
   return &comms_tab[c];
}

/* ---- requests and messages ------------------------------------------------ */

typedef struct { int tag, comm; int64_t nbytes; } frame;     /* 16 bytes */

typedef struct msg {                 /* an incoming message */
   int src, tag, comm;               /* src = world rank */
   int64_t nbytes, got;
   char *data;                       /* user buffer (matched) or private buffer */
   int owned;                        /* data is a private buffer */
   int req;                          /* receive request it completes, or -1 */
   int complete;                     /* every byte has arrived */
   struct msg *next;                 /* unexpected list */
} msg;

enum { R_FREE = 0, R_SEND, R_RECV };
typedef struct {
   int kind, done;
   /* send */
   const char *sbuf; int64_t snbytes, spushed; frame fr; int dst; int next_send;
   /* recv */
   char *rbuf; int64_t rmax; int src, tag, comm; msg *m; int next_recv;
   MPI_Status st;
} req_t;

static req_t *reqs;
static int reqs_cap, req_free_head = -1;
static int *send_head, *send_tail;   /* per destination FIFO of send requests */
static int recv_head = -1, recv_tail = -1;   /* posted, unmatched receives, in post order */
static msg *unexp_head, *unexp_tail;

typedef struct { frame fr; int have; msg *cur; } parser;
static parser *parsers;              /* per source */

static int req_alloc(void)
{
   int i;
   if (req_free_head < 0) {
      int ncap = reqs_cap ? 2*reqs_cap : 1024;
      reqs = (req_t *) realloc(reqs, (size_t) ncap*sizeof(req_t));
      if (!reqs) fatal("out of memory (requests)");
      for (i = ncap - 1; i >= reqs_cap; i--) {
         reqs[i].kind = R_FREE;
         reqs[i].next_send = req_free_head;
         req_free_head = i;
      }
      reqs_cap = ncap;
   }
   i = req_free_head;
   req_free_head = reqs[i].next_send;
   memset(&reqs[i], 0, sizeof(req_t));
   reqs[i].next_send = reqs[i].next_recv = -1;
   return i;
}

static void req_release(int i)
{
   reqs[i].kind = R_FREE;
   reqs[i].next_send = req_free_head;
   req_free_head = i;
}

static int match(int want_src, int want_tag, int want_comm, int src, int tag, int comm)
{
   return want_comm == comm && (want_src == MPI_ANY_SOURCE || want_src == src) &&
          (want_tag == MPI_ANY_TAG || want_tag == tag);
}

static void complete_recv(int r, msg *m)
{
   req_t *q = &reqs[r];
   if (m->owned) {
      if (m->nbytes) memcpy(q->rbuf, m->data, (size_t) m->nbytes);
      free(m->data);
   }
   q->st.MPI_SOURCE = m->src;       /* translated to the communicator's rank on return */
   q->st.MPI_TAG = m->tag;
   q->st.MPI_ERROR = MPI_SUCCESS;
   q->st.count_bytes = (int) m->nbytes;
   q->done = 1;
   q->m = NULL;
   free(m);
}

/* a frame header has arrived from `src`: bind the message to a receive */
static msg *open_msg(int src, const frame *fr)
{
   int r, prev = -1;
   msg *m = (msg *) calloc(1, sizeof(msg));
   if (!m) fatal("out of memory (message)");
   m->src = src; m->tag = fr->tag; m->comm = fr->comm; m->nbytes = fr->nbytes; m->req = -1;
   for (r = recv_head; r >= 0; prev = r, r = reqs[r].next_recv)
      if (match(reqs[r].src, reqs[r].tag, reqs[r].comm, src, fr->tag, fr->comm)) break;
   if (r >= 0) {
      if (fr->nbytes > reqs[r].rmax)
         fatal("message truncated: %lld bytes from rank %d tag %d into a %lld-byte receive",
               (long long) fr->nbytes, src, fr->tag, (long long) reqs[r].rmax);
      if (prev >= 0) reqs[prev].next_recv = reqs[r].next_recv; else recv_head = reqs[r].next_recv;
      if (recv_tail == r) recv_tail = prev;
      reqs[r].next_recv = -1;
      reqs[r].m = m;
      m->req = r;
      m->data = reqs[r].rbuf;
   } else {
      m->owned = 1;
      m->data = (char *) malloc((size_t) (fr->nbytes > 0 ? fr->nbytes : 1));
      if (!m->data) fatal("out of memory (unexpected message of %lld bytes)", (long long) fr->nbytes);
      if (unexp_tail) unexp_tail->next = m; else unexp_head = m;
      unexp_tail = m;
   }
   return m;
}

static size_t ring_read(ring_hdr *r, uint64_t tail, char *dst, size_t n)
{
   size_t off = (size_t) (tail & (ring_bytes - 1)), first = ring_bytes - off;
   if (first > n) first = n;
   memcpy(dst, ring_data(r) + off, first);
   if (n > first) memcpy(dst + first, ring_data(r), n - first);
   return n;
}

static void ring_write(ring_hdr *r, uint64_t head, const char *src, size_t n)
{
   size_t off = (size_t) (head & (ring_bytes - 1)), first = ring_bytes - off;
   if (first > n) first = n;
   memcpy(ring_data(r) + off, src, first);
   if (n > first) memcpy(ring_data(r), src + first, n - first);
}

/* returns 1 if anything moved */
static int progress(void)
{
   int moved = 0, p;
   if (atomic_load_explicit(&seg->abort_flag, memory_order_relaxed)) _exit(87);
   /* outgoing */
   for (p = 0; p < world_size; p++) {
      ring_hdr *r = ring_of(world_rank, p);
      while (send_head[p] >= 0) {
         int s = send_head[p];
         req_t *q = &reqs[s];
         uint64_t head = atomic_load_explicit(&r->head, memory_order_relaxed);
         uint64_t tail = atomic_load_explicit(&r->tail, memory_order_acquire);
         size_t space = ring_bytes - (size_t) (head - tail);      /* a multiple of 8 */
         int finished = 0;
         if (q->spushed == 0 && space >= sizeof(frame)) {
            ring_write(r, head, (const char *) &q->fr, sizeof(frame));
            head += sizeof(frame); space -= sizeof(frame);
            q->spushed = sizeof(frame);
            moved = 1;
         }
         if (q->spushed > 0) {
            int64_t poff = q->spushed - (int64_t) sizeof(frame);   /* payload bytes pushed */
            int64_t left = q->snbytes - poff;
            if (left >= 8 && space > 0) {
               size_t n = (size_t) (left & ~(int64_t) 7);
               if (n > space) n = space;
               ring_write(r, head, q->sbuf + poff, n);
               head += n; space -= n; poff += (int64_t) n; left -= (int64_t) n;
               moved = 1;
            }
            if (left > 0 && left < 8 && space > 0) {                 /* tail, padded to 8 */
               char padbuf[8] = {0};
               memcpy(padbuf, q->sbuf + poff, (size_t) left);
               ring_write(r, head, padbuf, 8);
               head += 8; poff += left; left = 0;
               moved = 1;
            }
            q->spushed = poff + (int64_t) sizeof(frame);
            finished = left <= 0;
         }
         atomic_store_explicit(&r->head, head, memory_order_release);
         if (!finished) break;                                       /* ring full */
         send_head[p] = q->next_send;
         if (send_head[p] < 0) send_tail[p] = -1;
         q->done = 1;
      }
   }
   /* incoming */
   for (p = 0; p < world_size; p++) {
      ring_hdr *r = ring_of(p, world_rank);
      parser *ps = &parsers[p];
      uint64_t tail = atomic_load_explicit(&r->tail, memory_order_relaxed);
      uint64_t head = atomic_load_explicit(&r->head, memory_order_acquire);
      if (head == tail) continue;
      moved = 1;
      while (head != tail) {
         size_t avail = (size_t) (head - tail);
         if (!ps->cur) {
            if (avail < sizeof(frame)) break;          /* frames are written whole */
            tail += ring_read(r, tail, (char *) &ps->fr, sizeof(frame));
            ps->cur = open_msg(p, &ps->fr);
            avail -= sizeof(frame);
         }
         {
            msg *m = ps->cur;
            int64_t padded = (m->nbytes + 7) & ~(int64_t) 7;
            int64_t left = padded - m->got;            /* ring bytes of this message to go */
            size_t n = (size_t) left < avail ? (size_t) left : avail;
            if (n > 0) {
               int64_t real = m->nbytes - m->got;      /* payload bytes among them */
               if (real < 0) real = 0;
               if ((int64_t) n <= real)
                  ring_read(r, tail, m->data + m->got, n);
               else {
                  if (real > 0) ring_read(r, tail, m->data + m->got, (size_t) real);
               }
               tail += n;
               m->got += (int64_t) n;
            }
            if (m->got >= padded) {
               m->got = m->nbytes;
               m->complete = 1;
               ps->cur = NULL;
               if (m->req >= 0) complete_recv(m->req, m);   /* else it waits on the unexpected list */
            } else
               break;
         }
      }
      atomic_store_explicit(&r->tail, tail, memory_order_release);
   }
   return moved;
}

static void idle(int *spins)
{
   if (++*spins < 200) return;
   if (*spins < 2000) { sched_yield(); return; }
   { struct timespec ts = {0, 50000}; nanosleep(&ts, NULL); }
}

static void wait_req(int r)
{
   int spins = 0;
   while (!reqs[r].done)
      if (progress()) spins = 0; else idle(&spins);
}

/* ---- point to point (world ranks, explicit context) ------------------------ */

static int post_send(const void *buf, int64_t nbytes, int dst_world, int tag, int comm)
{
   int s = req_alloc();
   req_t *q = &reqs[s];
   q->kind = R_SEND;
   q->sbuf = (const char *) buf; q->snbytes = nbytes; q->dst = dst_world;
   q->fr.tag = tag; q->fr.comm = comm; q->fr.nbytes = nbytes;
   if (send_tail[dst_world] >= 0) reqs[send_tail[dst_world]].next_send = s; else send_head[dst_world] = s;
   send_tail[dst_world] = s;
   progress();
   return s;
}

static int post_recv(void *buf, int64_t maxbytes, int src_world, int tag, int comm)
{
   int r = req_alloc();
   req_t *q = &reqs[r];
   msg *m, *prev = NULL;
   q->kind = R_RECV;
   q->rbuf = (char *) buf; q->rmax = maxbytes; q->src = src_world; q->tag = tag; q->comm = comm;
   for (m = unexp_head; m; prev = m, m = m->next)
      if (match(src_world, tag, comm, m->src, m->tag, m->comm)) break;
   if (m) {
      if (m->nbytes > maxbytes)
         fatal("message truncated: %lld bytes from rank %d tag %d into a %lld-byte receive",
               (long long) m->nbytes, m->src, m->tag, (long long) maxbytes);
      if (prev) prev->next = m->next; else unexp_head = m->next;
      if (unexp_tail == m) unexp_tail = prev;
      m->next = NULL;
      if (m->complete) {
         complete_recv(r, m);
      } else {                          /* still streaming in: finish into the private buffer */
         m->req = r;
         q->m = m;
      }
   } else {
      if (recv_tail >= 0) reqs[recv_tail].next_recv = r; else recv_head = r;
      recv_tail = r;
   }
   return r;
}

static void finish(int r, comm_t *c, MPI_Status *st)
{
   if (st && reqs[r].kind == R_RECV) {
      int i;
      *st = reqs[r].st;
      if (c)
         for (i = 0; i < c->size; i++)
            if (c->world[i] == st->MPI_SOURCE) { st->MPI_SOURCE = i; break; }
   }
   req_release(r);
}

/* requests remember their communicator only through the status translation; the
 * reference always uses MPI_COMM_WORLD for point-to-point, sub-communicators only
 * for collectives -- still, keep a side table so that Wait translates correctly */
static MPI_Comm *req_comm;
static int req_comm_cap;
static void remember_comm(int r, MPI_Comm c)
{
   if (r >= req_comm_cap) {
      int n = req_comm_cap ? req_comm_cap : 1024;
      while (n <= r) n *= 2;
      req_comm = (MPI_Comm *) realloc(req_comm, (size_t) n*sizeof(MPI_Comm));
      req_comm_cap = n;
   }
   req_comm[r] = c;
}

static size_t dt_size(MPI_Datatype dt) { return (size_t) (dt & 0xff); }

int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm,
              MPI_Request *request)
{
   comm_t *c = comm_of(comm);
   int r;
   if (dest < 0 || dest >= c->size) fatal("MPI_Isend: bad destination %d", dest);
   r = post_send(buf, (int64_t) count*(int64_t) dt_size(dt), c->world[dest], tag, comm);
   remember_comm(r, comm);
   *request = r;
   return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm,
              MPI_Request *request)
{
   comm_t *c = comm_of(comm);
   int r;
   if (source != MPI_ANY_SOURCE && (source < 0 || source >= c->size))
      fatal("MPI_Irecv: bad source %d", source);
   r = post_recv(buf, (int64_t) count*(int64_t) dt_size(dt),
                 source == MPI_ANY_SOURCE ? MPI_ANY_SOURCE : c->world[source], tag, comm);
   remember_comm(r, comm);
   *request = r;
   return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *request, MPI_Status *status)
{
   int r = *request;
   if (r == MPI_REQUEST_NULL) return MPI_SUCCESS;
   wait_req(r);
   finish(r, comm_of(req_comm[r]), status);
   *request = MPI_REQUEST_NULL;
   return MPI_SUCCESS;
}

int MPI_Waitany(int count, MPI_Request requests[], int *index, MPI_Status *status)
{
   int spins = 0;
   for (;;) {
      int i, live = 0;
      for (i = 0; i < count; i++) {
         int r = requests[i];
         if (r == MPI_REQUEST_NULL) continue;
         live = 1;
         if (reqs[r].done) {
            finish(r, comm_of(req_comm[r]), status);
            requests[i] = MPI_REQUEST_NULL;
            *index = i;
            return MPI_SUCCESS;
         }
      }
      if (!live) { *index = MPI_UNDEFINED; return MPI_SUCCESS; }
      if (progress()) spins = 0; else idle(&spins);
   }
}

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm)
{
   MPI_Request r;
   MPI_Isend(buf, count, dt, dest, tag, comm, &r);
   return MPI_Wait(&r, MPI_STATUS_IGNORE);
}

int MPI_Recv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm,
             MPI_Status *status)
{
   MPI_Request r;
   MPI_Irecv(buf, count, dt, source, tag, comm, &r);
   return MPI_Wait(&r, status);
}

/* ---- collectives ---------------------------------------------------------- */

enum { T_BARRIER = 1, T_BCAST, T_REDUCE, T_A2A, T_SCAN, T_GATHER };

static void csend(comm_t *c, MPI_Comm id, const void *buf, int64_t n, int dest, int tag)
{
   int r = post_send(buf, n, c->world[dest], tag, id | CTX_COLL);
   wait_req(r);
   req_release(r);
}

static void crecv(comm_t *c, MPI_Comm id, void *buf, int64_t n, int src, int tag)
{
   int r = post_recv(buf, n, c->world[src], tag, id | CTX_COLL);
   wait_req(r);
   req_release(r);
}

static void bcast_bytes(comm_t *c, MPI_Comm id, void *buf, int64_t n, int root, int tag)
{
   /* binomial tree rooted at `root` */
   int rel = (c->rank - root + c->size)%c->size, mask;
   for (mask = 1; mask < c->size; mask <<= 1)
      if (rel & mask) {
         crecv(c, id, buf, n, (rel - mask + root)%c->size, tag);
         break;
      }
   for (mask >>= 1; mask > 0; mask >>= 1)
      if (rel + mask < c->size) csend(c, id, buf, n, (rel + mask + root)%c->size, tag);
}

int MPI_Barrier(MPI_Comm comm)
{
   comm_t *c = comm_of(comm);
   char b = 0;
   int i;
   if (c->size == 1) return MPI_SUCCESS;
   if (c->rank == 0) for (i = 1; i < c->size; i++) crecv(c, comm, &b, 1, i, T_BARRIER);
   else csend(c, comm, &b, 1, 0, T_BARRIER);
   bcast_bytes(c, comm, &b, 1, 0, T_BARRIER);
   return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm)
{
   comm_t *c = comm_of(comm);
   if (c->size > 1) bcast_bytes(c, comm, buf, (int64_t) count*(int64_t) dt_size(dt), root, T_BCAST);
   return MPI_SUCCESS;
}

static void combine(void *acc, const void *in, int count, MPI_Datatype dt, MPI_Op op)
{
   int i;
#define LOOP(T)                                                                 \
   for (i = 0; i < count; i++) {                                                \
      T a = ((T *) acc)[i], b = ((const T *) in)[i];                            \
      ((T *) acc)[i] = op == MPI_SUM ? a + b : (op == MPI_MAX ? (a > b ? a : b) \
                                                              : (a < b ? a : b)); \
   }
   switch (dt & 0xf00) {
   case MINIMPI_KIND_INT: LOOP(int) break;
   case MINIMPI_KIND_DBL: LOOP(double) break;
   case MINIMPI_KIND_LL: LOOP(long long) break;
   default: fatal("reduction on an unsupported datatype 0x%x", dt);
   }
#undef LOOP
   if (op != MPI_SUM && op != MPI_MAX && op != MPI_MIN) fatal("unsupported reduction op %d", op);
}

int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op,
                  MPI_Comm comm)
{
   comm_t *c = comm_of(comm);
   int64_t n = (int64_t) count*(int64_t) dt_size(dt);
   int i;
   if (sbuf != rbuf) memmove(rbuf, sbuf, (size_t) n);
   if (c->size == 1) return MPI_SUCCESS;
   if (c->rank == 0) {
      /* ranks are folded in ascending order: ((r0 + r1) + r2) + ... */
      void *tmp = malloc((size_t) (n > 0 ? n : 1));
      for (i = 1; i < c->size; i++) {
         crecv(c, comm, tmp, n, i, T_REDUCE);
         combine(rbuf, tmp, count, dt, op);
      }
      free(tmp);
   } else
      csend(c, comm, rbuf, n, 0, T_REDUCE);
   bcast_bytes(c, comm, rbuf, n, 0, T_REDUCE);
   return MPI_SUCCESS;
}

int MPI_Scan(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm)
{
   comm_t *c = comm_of(comm);
   int64_t n = (int64_t) count*(int64_t) dt_size(dt);
   if (sbuf != rbuf) memmove(rbuf, sbuf, (size_t) n);
   if (c->rank > 0) {
      void *tmp = malloc((size_t) (n > 0 ? n : 1));
      crecv(c, comm, tmp, n, c->rank - 1, T_SCAN);
      /* prefix (ranks 0..r-1) op mine, prefix first */
      {
         void *mine = malloc((size_t) (n > 0 ? n : 1));
         memcpy(mine, rbuf, (size_t) n);
         memcpy(rbuf, tmp, (size_t) n);
         combine(rbuf, mine, count, dt, op);
         free(mine);
      }
      free(tmp);
   }
   if (c->rank + 1 < c->size) csend(c, comm, rbuf, n, c->rank + 1, T_SCAN);
   return MPI_SUCCESS;
}

int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount,
                 MPI_Datatype rdt, MPI_Comm comm)
{
   comm_t *c = comm_of(comm);
   int64_t sn = (int64_t) scount*(int64_t) dt_size(sdt), rn = (int64_t) rcount*(int64_t) dt_size(rdt);
   int i, *rr = (int *) malloc((size_t) c->size*2*sizeof(int)), *sr = rr + c->size;
   for (i = 0; i < c->size; i++)
      rr[i] = post_recv((char *) rbuf + (size_t) i*rn, rn, c->world[i], T_A2A, comm | CTX_COLL);
   for (i = 0; i < c->size; i++)
      sr[i] = post_send((const char *) sbuf + (size_t) i*sn, sn, c->world[i], T_A2A, comm | CTX_COLL);
   for (i = 0; i < c->size; i++) {
      wait_req(rr[i]); req_release(rr[i]);
      wait_req(sr[i]); req_release(sr[i]);
   }
   free(rr);
   return MPI_SUCCESS;
}

int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm)
{
   comm_t *c = comm_of(comm);
   int i, j, n = 0, id, mine[3], *all = (int *) malloc((size_t) c->size*3*sizeof(int));
   int *members, *keys;
   /* allgather (color, key, next free id) over the parent */
   mine[0] = color; mine[1] = key; mine[2] = next_comm;
   if (c->rank == 0) {
      memcpy(all, mine, sizeof mine);
      for (i = 1; i < c->size; i++) crecv(c, comm, all + 3*i, sizeof mine, i, T_GATHER);
   } else
      csend(c, comm, mine, sizeof mine, 0, T_GATHER);
   bcast_bytes(c, comm, all, (int64_t) c->size*3*sizeof(int), 0, T_GATHER);
   id = 0;
   for (i = 0; i < c->size; i++) if (all[3*i + 2] > id) id = all[3*i + 2];
   if (id >= MAX_COMMS) fatal("too many communicators");
   next_comm = id + 1;
   if (color == MPI_UNDEFINED) { *newcomm = MPI_COMM_NULL; free(all); return MPI_SUCCESS; }
   members = (int *) malloc((size_t) c->size*sizeof(int));
   keys = (int *) malloc((size_t) c->size*sizeof(int));
   for (i = 0; i < c->size; i++)
      if (all[3*i] == color) {
         /* insertion sort by (key, parent rank) */
         for (j = n; j > 0 && keys[j - 1] > all[3*i + 1]; j--) {
            keys[j] = keys[j - 1];
            members[j] = members[j - 1];
         }
         keys[j] = all[3*i + 1];
         members[j] = i;
         n++;
      }
   comms_tab[id].size = n;
   comms_tab[id].world = (int *) malloc((size_t) n*sizeof(int));
   for (i = 0; i < n; i++) {
      comms_tab[id].world[i] = c->world[members[i]];
      if (members[i] == c->rank) comms_tab[id].rank = i;
   }
   free(members); free(keys); free(all);
   *newcomm = id;
   return MPI_SUCCESS;
}

/* ---- environment ---------------------------------------------------------- */

int MPI_Init(int *argc, char ***argv)
{
   const char *name = getenv("MINIMPI_SHM"), *rk = getenv("MINIMPI_RANK"), *sz = getenv("MINIMPI_SIZE");
   int i;
   (void) argc; (void) argv;
   if (initialised) return MPI_SUCCESS;
   if (name && rk && sz) {
      int fd;
      struct stat sb;
      world_rank = atoi(rk);
      world_size = atoi(sz);
      fd = shm_open(name, O_RDWR, 0600);
      if (fd < 0) fatal("shm_open(%s): %s", name, strerror(errno));
      if (fstat(fd, &sb) != 0) fatal("fstat: %s", strerror(errno));
      seg = (minimpi_seg *) mmap(NULL, (size_t) sb.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (seg == MAP_FAILED) { seg = NULL; fatal("mmap: %s", strerror(errno)); }
      close(fd);
      if (seg->magic != MINIMPI_MAGIC || seg->size != world_size)
         fatal("segment %s does not belong to this job", name);
      ring_bytes = (size_t) seg->ring_bytes;
   } else {
      /* single rank without a launcher: a private segment with one self ring */
      size_t total;
      world_rank = 0; world_size = 1;
      ring_bytes = 1 << 20;
      total = minimpi_seg_bytes(1, ring_bytes);
      seg = (minimpi_seg *) mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      if (seg == MAP_FAILED) { seg = NULL; fatal("mmap: %s", strerror(errno)); }
      minimpi_seg_init(seg, 1, ring_bytes);
   }
   send_head = (int *) malloc((size_t) world_size*2*sizeof(int));
   send_tail = send_head + world_size;
   for (i = 0; i < world_size; i++) send_head[i] = send_tail[i] = -1;
   parsers = (parser *) calloc((size_t) world_size, sizeof(parser));
   comms_tab[0].size = world_size;
   comms_tab[0].rank = world_rank;
   comms_tab[0].world = (int *) malloc((size_t) world_size*sizeof(int));
   for (i = 0; i < world_size; i++) comms_tab[0].world[i] = i;
   initialised = 1;
   MPI_Barrier(MPI_COMM_WORLD);
   return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
   if (!initialised) return MPI_SUCCESS;
   MPI_Barrier(MPI_COMM_WORLD);
   initialised = 0;
   return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int errorcode)
{
   (void) comm;
   fflush(stdout);
   if (seg) atomic_store(&seg->abort_flag, 1);
   _exit(errorcode ? (errorcode & 0xff ? errorcode & 0xff : 1) : 1);
   return 0;
}

int MPI_Comm_set_errhandler(MPI_Comm comm, MPI_Errhandler eh) { (void) comm; (void) eh; return 0; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = comm_of(comm)->rank; return 0; }
int MPI_Comm_size(MPI_Comm comm, int *size) { *size = comm_of(comm)->size; return 0; }

double MPI_Wtime(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double) ts.tv_sec + 1.0e-9*(double) ts.tv_nsec;
}
