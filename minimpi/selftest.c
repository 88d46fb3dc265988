/* Self-test of the minimpi multi-process back-end: run as
 *     minimpirun -n N [--ring-kb 4] selftest
 * Every rank checks what it receives; rank 0 prints "minimpi selftest OK N".
 * The small ring size forces streaming, wrap-around and back-pressure. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mpi.h"

static int me, np;
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "[%d] selftest failed: %s (line %d)\n", me, #c, __LINE__); MPI_Abort(MPI_COMM_WORLD, 3); } } while (0)

static double val(int src, int dst, int k, long i) { return src*1000003.0 + dst*10007.0 + k*101.0 + (double) i; }

int main(int argc, char **argv)
{
   int i, k, p;
   MPI_Init(&argc, &argv);
   MPI_Comm_rank(MPI_COMM_WORLD, &me);
   MPI_Comm_size(MPI_COMM_WORLD, &np);

   /* 1. all-pairs exchange of large messages, receives posted first (comm.c pattern) */
   {
      long len = 300007;                        /* doubles: 2.4 MB per message */
      double **rb = malloc(np*sizeof(double *)), **sb = malloc(np*sizeof(double *));
      MPI_Request *rq = malloc(2*np*sizeof(MPI_Request));
      MPI_Status st;
      for (p = 0; p < np; p++) {
         rb[p] = malloc(len*sizeof(double));
         sb[p] = malloc(len*sizeof(double));
         for (i = 0; i < len; i++) sb[p][i] = val(me, p, 0, i);
         MPI_Irecv(rb[p], (int) len, MPI_DOUBLE, p, 7, MPI_COMM_WORLD, &rq[p]);
      }
      for (p = 0; p < np; p++)
         MPI_Isend(sb[p], (int) len, MPI_DOUBLE, p, 7, MPI_COMM_WORLD, &rq[np + p]);
      for (k = 0; k < np; k++) {
         int which;
         MPI_Waitany(np, rq, &which, &st);
         CHECK(which >= 0 && which < np && st.MPI_SOURCE == which && st.MPI_TAG == 7);
         for (i = 0; i < len; i++) CHECK(rb[which][i] == val(which, me, 0, i));
      }
      MPI_Waitany(np, rq, &k, &st);
      CHECK(k == MPI_UNDEFINED);
      for (p = 0; p < np; p++) MPI_Wait(&rq[np + p], &st);
      for (p = 0; p < np; p++) { free(rb[p]); free(sb[p]); }
      free(rb); free(sb); free(rq);
   }
   MPI_Barrier(MPI_COMM_WORLD);

   /* 2. unexpected messages: sends complete before any receive is posted; same-tag
    *    messages keep their order, different tags are matched out of order; odd sizes */
   {
      int right = (me + 1)%np, left = (me + np - 1)%np, n;
      char a[13], b[5], c[1], ra[13], rb5[5], rc[1];
      int big[5000], rbig[5000];
      MPI_Request rq[4];
      MPI_Status st;
      for (n = 0; n < 13; n++) a[n] = (char) (me + n);
      for (n = 0; n < 5; n++) b[n] = (char) (me*3 + n);
      c[0] = (char) (me + 77);
      for (n = 0; n < 5000; n++) big[n] = me*7919 + n;
      MPI_Isend(a, 13, MPI_BYTE, right, 1, MPI_COMM_WORLD, &rq[0]);
      MPI_Isend(b, 5, MPI_BYTE, right, 1, MPI_COMM_WORLD, &rq[1]);
      MPI_Isend(big, 5000, MPI_INT, right, 2, MPI_COMM_WORLD, &rq[2]);
      MPI_Isend(c, 0, MPI_BYTE, right, 3, MPI_COMM_WORLD, &rq[3]);
      for (n = 0; n < 4; n++) MPI_Wait(&rq[n], MPI_STATUS_IGNORE);
      MPI_Barrier(MPI_COMM_WORLD);              /* everything is in flight or unexpected */
      MPI_Recv(rc, 1, MPI_BYTE, left, 3, MPI_COMM_WORLD, &st);
      CHECK(st.count_bytes == 0);
      MPI_Recv(rbig, 5000, MPI_INT, left, 2, MPI_COMM_WORLD, &st);
      for (n = 0; n < 5000; n++) CHECK(rbig[n] == left*7919 + n);
      MPI_Recv(ra, 13, MPI_BYTE, left, 1, MPI_COMM_WORLD, &st);
      CHECK(st.count_bytes == 13);
      for (n = 0; n < 13; n++) CHECK(ra[n] == (char) (left + n));
      MPI_Recv(rb5, 13 > 5 ? 5 : 5, MPI_BYTE, left, 1, MPI_COMM_WORLD, &st);
      for (n = 0; n < 5; n++) CHECK(rb5[n] == (char) (left*3 + n));
   }

   /* 3. blocking handshake of rcb.c:207-337 style: flag then data */
   if (np > 1) {
      int peer = me ^ 1, flag = me, rflag = -1;
      double d[64], rd[64];
      for (i = 0; i < 64; i++) d[i] = val(me, peer, 3, i);
      if (peer < np) {
         if (me < peer) {
            MPI_Send(&flag, 1, MPI_INT, peer, 41, MPI_COMM_WORLD);
            MPI_Recv(&rflag, 1, MPI_INT, peer, 41, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
            MPI_Send(d, 64, MPI_DOUBLE, peer, 40, MPI_COMM_WORLD);
            MPI_Recv(rd, 64, MPI_DOUBLE, peer, 40, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
         } else {
            MPI_Recv(&rflag, 1, MPI_INT, peer, 41, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
            MPI_Send(&flag, 1, MPI_INT, peer, 41, MPI_COMM_WORLD);
            MPI_Recv(rd, 64, MPI_DOUBLE, peer, 40, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
            MPI_Send(d, 64, MPI_DOUBLE, peer, 40, MPI_COMM_WORLD);
         }
         CHECK(rflag == peer);
         for (i = 0; i < 64; i++) CHECK(rd[i] == val(peer, me, 3, i));
      }
   }

   /* 4. collectives */
   {
      int iv[3] = { me, -me, 1 }, io[3];
      double dv[2] = { 0.5*me, 1.0/(me + 1) }, dd[2], ds = 0, dm;
      long long lv = 1LL << (me%40), lo;
      int *to = malloc(np*sizeof(int)), *from = malloc(np*sizeof(int));
      MPI_Allreduce(iv, io, 3, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
      CHECK(io[0] == np*(np - 1)/2 && io[1] == -np*(np - 1)/2 && io[2] == np);
      MPI_Allreduce(iv, io, 3, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
      CHECK(io[0] == np - 1 && io[1] == 0 && io[2] == 1);
      MPI_Allreduce(iv, io, 3, MPI_INT, MPI_MIN, MPI_COMM_WORLD);
      CHECK(io[0] == 0 && io[1] == -(np - 1));
      MPI_Allreduce(dv, dd, 2, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
      for (p = 0, ds = 0; p < np; p++) ds += 1.0/(p + 1);       /* rank order: same bits */
      CHECK(dd[0] == 0.5*(np*(np - 1)/2) && dd[1] == ds);
      MPI_Allreduce(&dv[1], &dm, 1, MPI_DOUBLE, MPI_MIN, MPI_COMM_WORLD);
      CHECK(dm == 1.0/np);
      MPI_Allreduce(&lv, &lo, 1, MPI_LONG_LONG_INT, MPI_MAX, MPI_COMM_WORLD);
      CHECK(lo == 1LL << ((np - 1 < 40 ? np - 1 : 39)));
      for (p = 0; p < np; p++) to[p] = me*100 + p;
      MPI_Alltoall(to, 1, MPI_INT, from, 1, MPI_INT, MPI_COMM_WORLD);
      for (p = 0; p < np; p++) CHECK(from[p] == p*100 + me);
      i = me + 1;
      MPI_Scan(&i, &k, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
      CHECK(k == (me + 1)*(me + 2)/2);
      for (p = 0; p < np; p++) {
         double x[3] = { me == p ? 3.25 : -1, me == p ? p : -1, 7 };
         MPI_Bcast(x, 3, MPI_DOUBLE, p, MPI_COMM_WORLD);
         CHECK(x[0] == 3.25 && x[1] == p && x[2] == 7);
      }
      free(to); free(from);
   }

   /* 5. Comm_split chains as init.c:147-170 builds them; collectives on the parts */
   {
      MPI_Comm half, quarter;
      int r, s, sum, pre;
      MPI_Comm_split(MPI_COMM_WORLD, me%2, np - me, &half);       /* reversed order */
      MPI_Comm_rank(half, &r);
      MPI_Comm_size(half, &s);
      CHECK(s == (np + 1 - me%2)/2);
      i = me;
      MPI_Allreduce(&i, &sum, 1, MPI_INT, MPI_SUM, half);
      for (p = me%2, k = 0; p < np; p += 2) k += p;
      CHECK(sum == k);
      i = 1;
      MPI_Scan(&i, &pre, 1, MPI_INT, MPI_SUM, half);
      CHECK(pre == r + 1);
      /* highest world rank of my parity is rank 0 of `half` */
      i = me;
      MPI_Bcast(&i, 1, MPI_INT, 0, half);
      CHECK(i == ((np - 1)%2 == me%2 ? np - 1 : np - 2) || np == 1);
      MPI_Comm_split(half, r/2, r, &quarter);
      MPI_Comm_size(quarter, &s);
      CHECK(s >= 1 && s <= 2);
      MPI_Barrier(quarter);
      MPI_Barrier(half);
   }

   /* 6. many small messages in flight to one rank (refine flags, comm_refine.c) */
   {
      int n = 2000, *got = calloc(n, sizeof(int)), x;
      MPI_Request *rq = malloc(n*sizeof(MPI_Request));
      int *vals = malloc(n*sizeof(int));
      if (me != 0) {
         for (i = 0; i < n; i++) { vals[i] = me*n + i; MPI_Isend(&vals[i], 1, MPI_INT, 0, 9, MPI_COMM_WORLD, &rq[i]); }
         for (i = 0; i < n; i++) MPI_Wait(&rq[i], MPI_STATUS_IGNORE);
      } else
         for (p = 1; p < np; p++)
            for (i = 0; i < n; i++) {
               MPI_Recv(&x, 1, MPI_INT, p, 9, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
               CHECK(x == p*n + i);                               /* pairwise order */
            }
      free(got); free(rq); free(vals);
   }

   MPI_Barrier(MPI_COMM_WORLD);
   if (me == 0) printf("minimpi selftest OK %d\n", np);
   MPI_Finalize();
   return 0;
}
