/* TEST INFRASTRUCTURE — CPU oracle for the miniAMR stage hot path.
 *
 * A plain-C restatement (written from the algorithm, direction-generic, not a
 * copy) of what the reference computes on the path
 *     comm() -> stencil_driver()/stencil_calc() -> check_sum()
 * plus the block data movement that shares its storage (split, consolidate,
 * pack/unpack for migration).  Every function cites the reference file:line
 * it follows.  PARITY IS PINNED: tests/test_oracle_vs_reference.py checks every
 * routine here bit-for-bit against the unmodified reference compiled into
 * oracle/_ref/libminiamr_ref.so, and tests/golden/ holds vectors generated
 * from that reference (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.  It is never linked into, or called from, the product
 * (miniamr_b200/), which has no CPU fallback.
 *
 * Storage convention (the oracle's own, chosen to be trivially comparable with
 * the reference's block.array[var][i][j][k], block.h:52):
 *     data[slot][var][i][j][k],  i in 0..nx+1, j in 0..ny+1, k in 0..nz+1,
 *     k fastest, ghosts at index 0 and n+1.
 * Topology is passed as slot-indexed int arrays mirroring block.h:36-53:
 *     level[slot], nei_level[slot][6], nei[slot][6][2][2].
 */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
   int n[3];            /* nx, ny, nz */
   int num_vars;
   int stencil;         /* 7 or 27 */
   size_t stride[3];    /* element strides of i, j, k inside a tile */
   size_t tile;         /* (nx+2)(ny+2)(nz+2) */
   double *data;
} orc_mesh;

static orc_mesh mk(double *data, int nx, int ny, int nz, int num_vars, int stencil)
{
   orc_mesh m;
   m.n[0] = nx; m.n[1] = ny; m.n[2] = nz;
   m.num_vars = num_vars;
   m.stencil = stencil;
   m.stride[2] = 1;
   m.stride[1] = (size_t)(nz + 2);
   m.stride[0] = (size_t)(ny + 2)*(nz + 2);
   m.tile = (size_t)(nx + 2)*(ny + 2)*(nz + 2);
   m.data = data;
   return m;
}

static double *tile_of(const orc_mesh *m, int slot, int var)
{
   return m->data + ((size_t)slot*m->num_vars + var)*m->tile;
}

/* in-face axes of direction d in buffer order (slow, fast): always the
 * lower-numbered then the higher-numbered remaining axis
 * (comm.c:266-270 X->(j,k), :311-320 Y->(i,k), :361-370 Z->(i,j)) */
static void face_axes(int d, int *slow, int *fast)
{
   *slow = (d == 0) ? 1 : 0;
   *fast = (d == 2) ? 1 : 2;
}

/* -------------------------------------------------------------------------
 * stencil_calc, stencil.c:76-145.  Jacobi within a block: every new interior
 * value is computed from the old tile, then written back.
 *   7-pt  (stencil.c:88-94): ((((((W+S)+D)+C)+U)+N)+E)/7.0 with
 *          W=[i-1], S=[j-1], D=[k-1], C, U=[k+1], N=[j+1], E=[i+1]
 *   27-pt (stencil.c:111-138): sb, sm, sf = 9-term sums of the i-1, i, i+1
 *          planes, each accumulated j-major / k-minor; ((sb+sm)+sf)/27.0
 * ---------------------------------------------------------------------- */
void orc_stencil_calc(double *data, int nx, int ny, int nz, int num_vars,
                      int num_active, const int *slots, int var, int stencil)
{
   orc_mesh m = mk(data, nx, ny, nz, num_vars, stencil);
   const size_t si = m.stride[0], sj = m.stride[1];
   double *work = (double *) malloc(m.tile*sizeof(double));
   int a, i, j, k, dj, dk;

   for (a = 0; a < num_active; a++) {
      double *t = tile_of(&m, slots[a], var);
      for (i = 1; i <= nx; i++)
         for (j = 1; j <= ny; j++)
            for (k = 1; k <= nz; k++) {
               const double *c = t + i*si + j*sj + k;
               if (stencil == 7) {
                  double s = c[-(ptrdiff_t)si] + c[-(ptrdiff_t)sj];
                  s += c[-1];
                  s += c[0];
                  s += c[1];
                  s += c[sj];
                  s += c[si];
                  work[i*si + j*sj + k] = s/7.0;
               } else {
                  double plane[3];
                  int p;
                  for (p = 0; p < 3; p++) {
                     const double *q = c + (ptrdiff_t)(p - 1)*(ptrdiff_t)si;
                     double s = q[-(ptrdiff_t)sj - 1];
                     for (dj = -1; dj <= 1; dj++)
                        for (dk = -1; dk <= 1; dk++)
                           if (!(dj == -1 && dk == -1))
                              s += q[(ptrdiff_t)dj*(ptrdiff_t)sj + dk];
                     plane[p] = s;
                  }
                  work[i*si + j*sj + k] = (plane[0] + plane[1] + plane[2])/27.0;
               }
            }
      for (i = 1; i <= nx; i++)
         for (j = 1; j <= ny; j++)
            for (k = 1; k <= nz; k++)
               t[i*si + j*sj + k] = work[i*si + j*sj + k];
   }
   free(work);
}

/* -------------------------------------------------------------------------
 * check_sum local part, check_sum.c:45-53: per block a sequential i,j,k sum of
 * the interior, blocks accumulated in sorted_list order.  (The MPI_Allreduce
 * at check_sum.c:57 is the caller's business.)
 * ---------------------------------------------------------------------- */
double orc_check_sum(const double *data, int nx, int ny, int nz, int num_vars,
                     int num_active, const int *slots, int var)
{
   orc_mesh m = mk((double *) data, nx, ny, nz, num_vars, 7);
   double sum = 0.0;
   int a, i, j, k;
   for (a = 0; a < num_active; a++) {
      const double *t = tile_of(&m, slots[a], var);
      double bs = 0.0;
      for (i = 1; i <= nx; i++)
         for (j = 1; j <= ny; j++)
            for (k = 1; k <= nz; k++)
               bs += t[i*m.stride[0] + j*m.stride[1] + k];
      sum += bs;
   }
   return sum;
}

/* per-block sums (for error localisation in tests) */
void orc_block_sums(const double *data, int nx, int ny, int nz, int num_vars,
                    int num_active, const int *slots, int var, double *out)
{
   int a;
   for (a = 0; a < num_active; a++)
      out[a] = orc_check_sum(data, nx, ny, nz, num_vars, 1, slots + a, var);
}

/* -------------------------------------------------------------------------
 * Ghost exchange, --code 0.
 * ---------------------------------------------------------------------- */

/* extent of in-face axis `ax` for a whole-face transfer in direction d:
 * 1..n for the 7-pt stencil; for the other stencils the axes of directions
 * already exchanged (ax < d) are widened to 0..n+1 so edges/corners travel
 * (on_proc_comm comm.c:1496-1527; pack_face cases 0/1 comm.c:306-320,356-370) */
static void whole_extent(const orc_mesh *m, int wide, int d, int ax, int *lo, int *hi)
{
   if (wide && ax < d) {
      *lo = 0; *hi = m->n[ax] + 1;
   } else {
      *lo = 1; *hi = m->n[ax];
   }
}

/* same level, both blocks on this rank: on_proc_comm, comm.c:1473-1534.
 * `lo` is the block on the minus side of the shared face, `hi` on the plus. */
static void exch_same(const orc_mesh *m, int lo, int hi, int d, int start, int num)
{
   int sa, fa, s0, s1, f0, f1, s, f, v;
   face_axes(d, &sa, &fa);
   whole_extent(m, m->stencil != 7, d, sa, &s0, &s1);
   whole_extent(m, m->stencil != 7, d, fa, &f0, &f1);
   for (v = start; v < start + num; v++) {
      double *tl = tile_of(m, lo, v), *th = tile_of(m, hi, v);
      for (s = s0; s <= s1; s++)
         for (f = f0; f <= f1; f++) {
            size_t o = s*m->stride[sa] + f*m->stride[fa];
            tl[o + (m->n[d] + 1)*m->stride[d]] = th[o + 1*m->stride[d]];
            th[o] = tl[o + m->n[d]*m->stride[d]];
         }
   }
}

/* different levels on this rank: on_proc_comm_diff, comm.c:1597-1688.
 * `c` coarse block, `f` fine block, l = face of the COARSE block, (iq,jq) the
 * quarter of that face: jq selects the half along the slow in-face axis, iq
 * along the fast one (comm.c:1616-1617, 1644-1645, 1672-1673).  Coarse->fine:
 * value/4 replicated 2x2; fine->coarse: 4-term sum in the order
 * [2a-1][2b-1] + [2a-1][2b] + [2a][2b-1] + [2a][2b]. */
static void exch_diff(const orc_mesh *m, int c, int f, int l, int iq, int jq,
                      int start, int num)
{
   int d = l/2, sa, fa, a, b, v;
   int hs, hf, os, of;
   int c_ghost, c_src, f_ghost, f_src;
   face_axes(d, &sa, &fa);
   hs = m->n[sa]/2; hf = m->n[fa]/2;
   os = jq*hs; of = iq*hf;
   if (l%2 == 0) {   /* fine block sits on the minus side of the coarse one */
      c_ghost = 0;            c_src = 1;
      f_ghost = m->n[d] + 1;  f_src = m->n[d];
   } else {
      c_ghost = m->n[d] + 1;  c_src = m->n[d];
      f_ghost = 0;            f_src = 1;
   }
   for (v = start; v < start + num; v++) {
      double *tc = tile_of(m, c, v), *tf = tile_of(m, f, v);
      size_t S = m->stride[sa], F = m->stride[fa], N = m->stride[d];
      for (a = 1; a <= hs; a++)
         for (b = 1; b <= hf; b++) {
            double q = tc[c_src*N + (a + os)*S + (b + of)*F]/4.0;
            double *g = tf + f_ghost*N;
            const double *p = tf + f_src*N;
            g[(2*a-1)*S + (2*b-1)*F] = q;
            g[(2*a-1)*S + (2*b  )*F] = q;
            g[(2*a  )*S + (2*b-1)*F] = q;
            g[(2*a  )*S + (2*b  )*F] = q;
            tc[c_ghost*N + (a + os)*S + (b + of)*F] =
               p[(2*a-1)*S + (2*b-1)*F] + p[(2*a-1)*S + (2*b)*F] +
               p[(2*a  )*S + (2*b-1)*F] + p[(2*a  )*S + (2*b)*F];
         }
   }
}

/* reflective boundary: apply_bc, comm.c:1911-1965.  7-pt code 0 copies the
 * 1..n x 1..n part of the adjacent interior plane; every other mode copies the
 * full 0..n+1 x 0..n+1 plane. */
static void exch_bc(const orc_mesh *m, int b, int l, int start, int num)
{
   int d = l/2, sa, fa, s, f, v, s0, s1, f0, f1;
   int to = (l%2) ? m->n[d] + 1 : 0, from = (l%2) ? m->n[d] : 1;
   face_axes(d, &sa, &fa);
   if (m->stencil == 7) {
      s0 = 1; s1 = m->n[sa]; f0 = 1; f1 = m->n[fa];
   } else {
      s0 = 0; s1 = m->n[sa] + 1; f0 = 0; f1 = m->n[fa] + 1;
   }
   for (v = start; v < start + num; v++) {
      double *t = tile_of(m, b, v);
      for (s = s0; s <= s1; s++)
         for (f = f0; f <= f1; f++) {
            size_t o = s*m->stride[sa] + f*m->stride[fa];
            t[o + to*m->stride[d]] = t[o + from*m->stride[d]];
         }
   }
}

/* On-rank part of ONE direction phase of comm(): the loop at comm.c:162-203.
 * Off-rank neighbours (nei < 0) are skipped here exactly as there (the pair
 * rule m > n can never hold for a negative m); their faces travel through
 * orc_pack_face / orc_unpack_face.  Returns -1 on a misconnected block
 * (comm.c:198-201), else 0.  counters[0..2] += same, diff, bc (comm.c:169,
 * 178,189,196). */
int orc_comm_dir_local(double *data, int nx, int ny, int nz, int num_vars,
                       int stencil, int num_active, const int *slots,
                       const int *level, const int *nei_level, const int *nei,
                       int dir, int start, int num_comm, int *counters)
{
   orc_mesh m = mk(data, nx, ny, nz, num_vars, stencil);
   int a, l, i, j;
   for (a = 0; a < num_active; a++) {
      int n = slots[a];
      for (l = 2*dir; l < 2*dir + 2; l++) {
         int nl = nei_level[n*6 + l];
         const int *ne = nei + ((size_t)n*6 + l)*4;
         if (nl == level[n]) {
            int mm = ne[0];
            if (mm > n) {
               if (l%2 == 0) exch_same(&m, mm, n, dir, start, num_comm);
               else          exch_same(&m, n, mm, dir, start, num_comm);
               if (counters) counters[0] += 2;
            }
         } else if (nl == level[n] + 1) {
            for (i = 0; i < 2; i++)
               for (j = 0; j < 2; j++) {
                  int mm = ne[i*2 + j];
                  if (mm > n) {
                     exch_diff(&m, n, mm, l, i, j, start, num_comm);
                     if (counters) counters[1] += 2;
                  }
               }
         } else if (nl == level[n] - 1) {
            int mm = ne[0];
            if (mm > n) {
               int k = 2*dir + 1 - l%2;
               const int *ne2 = nei + ((size_t)mm*6 + k)*4;
               for (i = 0; i < 2; i++)
                  for (j = 0; j < 2; j++)
                     if (ne2[i*2 + j] == n) {
                        exch_diff(&m, mm, n, k, i, j, start, num_comm);
                        if (counters) counters[1] += 2;
                     }
            }
         } else if (nl == -2) {
            exch_bc(&m, n, l, start, num_comm);
            if (counters) counters[2] += 1;
         } else
            return -1;
      }
   }
   return 0;
}

/* direction order of comm(): comm.c:45-55 */
void orc_phase_order(int permute, int stage, int *order)
{
   static const int perm[6][3] = { {0, 1, 2}, {1, 2, 0}, {2, 0, 1},
                                   {0, 2, 1}, {1, 0, 2}, {2, 1, 0} };
   int o;
   for (o = 0; o < 3; o++)
      order[o] = permute ? perm[stage%6][o] : o;
}

/* whole comm() on one rank with no off-rank partners: comm.c:42-242 */
int orc_comm_local(double *data, int nx, int ny, int nz, int num_vars,
                   int stencil, int num_active, const int *slots,
                   const int *level, const int *nei_level, const int *nei,
                   int start, int num_comm, int stage, int permute, int *counters)
{
   int order[3], o, rc;
   orc_phase_order(permute, stage, order);
   for (o = 0; o < 3; o++) {
      rc = orc_comm_dir_local(data, nx, ny, nz, num_vars, stencil, num_active,
                              slots, level, nei_level, nei, order[o], start,
                              num_comm, counters ? counters + 3*order[o] : NULL);
      if (rc) return rc;
   }
   return 0;
}

/* quarter selection shared by pack and unpack, cases 6-9:
 * slow half is the low one iff case%2==0, fast half is the low one iff
 * (case/2)%2==1 (comm.c:282-295 and 1029-1042). */
static void quarter_range(const orc_mesh *m, int fc, int sa, int fa,
                          int *s0, int *s1, int *f0, int *f1)
{
   int hs = m->n[sa]/2, hf = m->n[fa]/2;
   if (fc%2 == 0) { *s0 = 1; *s1 = hs; } else { *s0 = hs + 1; *s1 = m->n[sa]; }
   if ((fc/2)%2 == 1) { *f0 = 1; *f1 = hf; } else { *f0 = hf + 1; *f1 = m->n[fa]; }
}

/* pack_face, code 0: comm.c:254-401.  Returns the number of doubles written
 * (num_comm * face length); the caller places faces at comm_send_off. */
int orc_pack_face(const double *data, int nx, int ny, int nz, int num_vars,
                  int stencil, double *buf, int slot, int face_case, int dir,
                  int start, int num_comm)
{
   orc_mesh m = mk((double *) data, nx, ny, nz, num_vars, stencil);
   int sa, fa, s, f, v, n = 0, s0, s1, f0, f1, plane;
   face_axes(dir, &sa, &fa);
   if (face_case >= 10) { plane = m.n[dir]; face_case -= 10; } else plane = 1;
   for (v = start; v < start + num_comm; v++) {
      const double *t = tile_of(&m, slot, v) + plane*m.stride[dir];
      size_t S = m.stride[sa], F = m.stride[fa];
      if (face_case < 2) {
         /* whole -> whole.  X never widens (comm.c:266-270 treats cases 0 and
          * 1 alike); Y/Z widen only when the case says so (0 vs 1). */
         whole_extent(&m, face_case == 1, dir, sa, &s0, &s1);
         whole_extent(&m, face_case == 1, dir, fa, &f0, &f1);
         for (s = s0; s <= s1; s++)
            for (f = f0; f <= f1; f++)
               buf[n++] = t[s*S + f*F];
      } else if (face_case <= 5) {
         /* I am fine, neighbour coarse: restrict by 4-sum (comm.c:271-280) */
         for (s = 1; s < m.n[sa]; s += 2)
            for (f = 1; f < m.n[fa]; f += 2)
               buf[n++] = t[s*S + f*F] + t[s*S + (f+1)*F] +
                          t[(s+1)*S + f*F] + t[(s+1)*S + (f+1)*F];
      } else {
         /* I am coarse, neighbour fine: my quarter / 4 (comm.c:281-300) */
         quarter_range(&m, face_case, sa, fa, &s0, &s1, &f0, &f1);
         for (s = s0; s <= s1; s++)
            for (f = f0; f <= f1; f++)
               buf[n++] = t[s*S + f*F]/4.0;
      }
   }
   return n;
}

/* unpack_face, code 0: comm.c:1002-1150 (face_case is the receiver's own) */
int orc_unpack_face(double *data, int nx, int ny, int nz, int num_vars,
                    int stencil, const double *buf, int slot, int face_case,
                    int dir, int start, int num_comm)
{
   orc_mesh m = mk(data, nx, ny, nz, num_vars, stencil);
   int sa, fa, s, f, v, n = 0, s0, s1, f0, f1, plane;
   face_axes(dir, &sa, &fa);
   if (face_case >= 10) { plane = m.n[dir] + 1; face_case -= 10; } else plane = 0;
   for (v = start; v < start + num_comm; v++) {
      double *t = tile_of(&m, slot, v) + plane*m.stride[dir];
      size_t S = m.stride[sa], F = m.stride[fa];
      if (face_case < 2) {
         whole_extent(&m, face_case == 1, dir, sa, &s0, &s1);
         whole_extent(&m, face_case == 1, dir, fa, &f0, &f1);
         for (s = s0; s <= s1; s++)
            for (f = f0; f <= f1; f++)
               t[s*S + f*F] = buf[n++];
      } else if (face_case <= 5) {
         /* I am fine: each value fills a 2x2 patch (comm.c:1020-1028) */
         for (s = 1; s < m.n[sa]; s += 2)
            for (f = 1; f < m.n[fa]; f += 2) {
               double q = buf[n++];
               t[s*S + f*F] = q; t[s*S + (f+1)*F] = q;
               t[(s+1)*S + f*F] = q; t[(s+1)*S + (f+1)*F] = q;
            }
      } else {
         /* I am coarse: write my quarter (comm.c:1029-1048) */
         quarter_range(&m, face_case, sa, fa, &s0, &s1, &f0, &f1);
         for (s = s0; s <= s1; s++)
            for (f = f0; f <= f1; f++)
               t[s*S + f*F] = buf[n++];
      }
   }
   return n;
}

/* -------------------------------------------------------------------------
 * Refinement data movement.
 * ---------------------------------------------------------------------- */

/* split: block.c:143-173.  Octant o of the parent (i1=(o%2)nx/2,
 * j1=((o/2)%2)ny/2, k1=(o/4)nz/2) -> child o; every parent cell / 8.0 fills
 * the 2x2x2 child cells.  Child ghosts are not touched. */
void orc_split_block(double *data, int nx, int ny, int nz, int num_vars,
                     int parent_slot, const int *child_slots)
{
   orc_mesh m = mk(data, nx, ny, nz, num_vars, 7);
   int o, v, i, j, k, di, dj, dk;
   const size_t si = m.stride[0], sj = m.stride[1];
   for (o = 0; o < 8; o++) {
      int i1 = (o%2)*(nx/2), j1 = ((o/2)%2)*(ny/2), k1 = (o/4)*(nz/2);
      for (v = 0; v < num_vars; v++) {
         const double *p = tile_of(&m, parent_slot, v);
         double *c = tile_of(&m, child_slots[o], v);
         for (i = 1; i <= nx/2; i++)
            for (j = 1; j <= ny/2; j++)
               for (k = 1; k <= nz/2; k++) {
                  double q = p[(i+i1)*si + (j+j1)*sj + (k+k1)]/8.0;
                  for (di = 0; di < 2; di++)
                     for (dj = 0; dj < 2; dj++)
                        for (dk = 0; dk < 2; dk++)
                           c[(2*i-1+di)*si + (2*j-1+dj)*sj + (2*k-1+dk)] = q;
               }
      }
   }
}

/* consolidate: block.c:411-431.  Parent cell = 8-term sum in the order
 * (i2,j2,k2) (i2+1,j2,k2) (i2,j2+1,k2) (i2+1,j2+1,k2) then the same four at
 * k2+1, accumulated left to right. */
void orc_consolidate_block(double *data, int nx, int ny, int nz, int num_vars,
                           const int *child_slots, int parent_slot)
{
   orc_mesh m = mk(data, nx, ny, nz, num_vars, 7);
   int o, v, i, j, k;
   const size_t si = m.stride[0], sj = m.stride[1];
   for (o = 0; o < 8; o++) {
      int i1 = (o%2)*(nx/2), j1 = ((o/2)%2)*(ny/2), k1 = (o/4)*(nz/2);
      for (v = 0; v < num_vars; v++) {
         double *p = tile_of(&m, parent_slot, v);
         const double *c = tile_of(&m, child_slots[o], v);
         for (i = 1; i <= nx/2; i++)
            for (j = 1; j <= ny/2; j++)
               for (k = 1; k <= nz/2; k++) {
                  const double *q = c + (2*i-1)*si + (2*j-1)*sj + (2*k-1);
                  double s = q[0] + q[si];
                  s += q[sj];
                  s += q[si + sj];
                  s += q[1];
                  s += q[si + 1];
                  s += q[sj + 1];
                  s += q[si + sj + 1];
                  p[(i+i1)*si + (j+j1)*sj + (k+k1)] = s;
               }
      }
   }
}

/* migration payload: pack.c:66-70 / 103-107 — interiors only, var-major then
 * i, j, k.  (The 50-slot integer header of pack.c:40-64 is host metadata and is
 * not restated here.) */
void orc_pack_block(const double *data, int nx, int ny, int nz, int num_vars,
                    int slot, double *payload)
{
   orc_mesh m = mk((double *) data, nx, ny, nz, num_vars, 7);
   int v, i, j, k;
   size_t n = 0;
   for (v = 0; v < num_vars; v++) {
      const double *t = tile_of(&m, slot, v);
      for (i = 1; i <= nx; i++)
         for (j = 1; j <= ny; j++)
            for (k = 1; k <= nz; k++)
               payload[n++] = t[i*m.stride[0] + j*m.stride[1] + k];
   }
}

void orc_unpack_block(double *data, int nx, int ny, int nz, int num_vars,
                      int slot, const double *payload)
{
   orc_mesh m = mk(data, nx, ny, nz, num_vars, 7);
   int v, i, j, k;
   size_t n = 0;
   for (v = 0; v < num_vars; v++) {
      double *t = tile_of(&m, slot, v);
      for (i = 1; i <= nx; i++)
         for (j = 1; j <= ny; j++)
            for (k = 1; k <= nz; k++)
               t[i*m.stride[0] + j*m.stride[1] + k] = payload[n++];
   }
}

/* -------------------------------------------------------------------------
 * stencil_driver() with --stencil 0, stencil.c:43-74: the "variable work" mix.
 * mat = num_vars/4, a1, a0[mat]: init.c:418-423.  Variable 0 and variables
 * >= 4*mat take the 7-point average.  For the others stage%6 picks
 *   0 stencil_0 :147-226     1/2/3 stencil_x/y/z :228-659 (in place along one axis)
 *   4 stencil_7 :661-784     5 stencil_27 :786-957        (through work[])
 * and stencil_check :959-983 follows.  flops[0..2] += adds, muls, divs as the
 * reference books them (per cell for stencil_check).
 * Restated per cell with one accessor; evaluation order as C parses the
 * reference's expressions.
 * ---------------------------------------------------------------------- */
#include <math.h>

#define S0V(v, di, dj, dk) (base[((size_t)(v))*m.tile + (size_t)((i) + (di))*si + (size_t)((j) + (dj))*sj + (size_t)((k) + (dk))])
#define S0C(v) S0V(v, 0, 0, 0)

void orc_stencil0_driver(double *data, int nx, int ny, int nz, int num_vars, int num_active,
                         const int *slots, int var, int stage, int mat, double a1, const double *a0,
                         double *flops)
{
   orc_mesh m = mk(data, nx, ny, nz, num_vars, 0);
   const size_t si = m.stride[0], sj = m.stride[1];
   const int kind = stage%6;
   const double cells = (double)nx*ny*nz;
   int a, i, j, k, v, di, dj, dk;
   double *work;

   if (var == 0 || var >= 4*mat) {
      orc_stencil_calc(data, nx, ny, nz, num_vars, num_active, slots, var, 7);
      flops[0] += 6.0*cells*num_active;                           /* stencil.c:100-101 */
      flops[2] += cells*num_active;
      return;
   }
   work = (double *) malloc(m.tile*sizeof(double));
   for (a = 0; a < num_active; a++) {
      double *base = tile_of(&m, slots[a], 0);
      const int q = var/mat, col = var%mat;                       /* band and column */
      const int b1 = col + ((q + 1)%4)*mat, b2 = col + ((q + 2)%4)*mat, b3 = col + ((q + 3)%4)*mat;
      for (i = 1; i <= nx; i++)
         for (j = 1; j <= ny; j++)
            for (k = 1; k <= nz; k++) {
               double x = S0C(var), r;
               if (kind == 0) {
                  if (var == 1) {                                 /* :152-163 */
                     for (v = mat; v < 2*mat; v++) x += S0C(v)*S0C(0);
                     r = x;
                  } else if (q == 0)                              /* :164-176 */
                     r = x + x*(S0C(0) + S0C(1) - a1*x);
                  else if (q == 1)                                /* :177-193 */
                     r = x*(S0C(0) + x + a1*S0C(var + mat) + (1.0 - a1)*S0C(var + 2*mat))/S0C(1);
                  else if (q == 2)                                /* :194-209 */
                     r = x + S0C(var - mat)*(a1*S0C(0) + a0[var - 2*mat]*x + (1.0 - a1)*S0C(var + mat))/S0C(1);
                  else                                            /* :210-225 */
                     r = x + S0C(var - 2*mat)*(a1*S0C(0) + a0[var - 3*mat]*x +
                                               (1.0 - a0[var - 3*mat])*S0C(var - mat) +
                                               (1.0 - a1)*S0C(var - 2*mat))/(S0C(1)*S0C(1));
                  S0C(var) = r;
               } else if (kind <= 3) {
                  di = kind == 1; dj = kind == 2; dk = kind == 3;   /* the sweep axis */
                  if (var == 1) {                                 /* :234-248 */
                     for (v = 2; v < mat + 2; v++) x += S0C(v)*S0C(0);
                     r = x/(a1 + x);
                  } else if (q == 0)                              /* :249-262 */
                     r = x + x*(S0C(0) + S0C(1) - a1*x)/(a0[var] + S0C(1));
                  else {
                     /* slope variable: itself (q=1), var-mat (q=2), var-2mat (q=3) */
                     const int sv = q == 1 ? var : (q == 2 ? var - mat : var - 2*mat);
                     const double lo = S0V(var, -di, -dj, -dk), hi = S0V(var, di, dj, dk);
                     const double t1 = fabs(S0C(sv) - S0V(sv, -di, -dj, -dk));
                     const double t2 = fabs(S0C(sv) - S0V(sv, di, dj, dk));
                     double den;
                     if (q == 1)                                  /* :263-296 */
                        den = a1 + a0[var - mat] + lo + x + hi + S0C(0) + S0C(1);
                     else if (q == 2)                             /* :297-331 */
                        den = a1 + a0[var - 2*mat] + S0C(var - mat) + S0C(var + mat) + lo + x + hi;
                     else                                         /* :332-366 */
                        den = a1 + a0[var - 3*mat] + S0C(var - mat) + S0C(var - 2*mat) + lo + x + hi;
                     if (t1 > t2) r = (t1*lo + (t1 - t2)*(x + S0C(1)) + t2*hi)/den;
                     else r = (t1*lo + (t2 - t1)*(x + S0C(1)) + t2*hi)/den;
                  }
                  S0C(var) = r;                                   /* in place: [i-1] is new, [i+1] old */
               } else if (kind == 4) {                            /* :661-784 */
                  work[i*si + j*sj + k] = (S0V(var, -1, 0, 0)*S0V(b1, -1, 0, 0) + S0V(var, 0, -1, 0)*S0V(b2, 0, -1, 0) +
                                           S0V(var, 0, 0, -1)*S0V(b3, 0, 0, -1) + x*x +
                                           S0V(var, 0, 0, 1)*S0V(b3, 0, 0, 1) + S0V(var, 0, 1, 0)*S0V(b2, 0, 1, 0) +
                                           S0V(var, 1, 0, 0)*S0V(b1, 1, 0, 0))/7.0*(a1 + x);
               } else {                                           /* :786-957 */
                  double sum = 0.0;
                  int first = 1;
                  for (di = -1; di <= 1; di++)
                     for (dj = -1; dj <= 1; dj++)
                        for (dk = -1; dk <= 1; dk++) {
                           const int dist = (di != 0) + (dj != 0) + (dk != 0);
                           const int sv = dist == 0 ? var : (dist == 1 ? b1 : (dist == 2 ? b2 : b3));
                           const double t = S0V(sv, di, dj, dk);
                           sum = first ? t : sum + t;
                           first = 0;
                        }
                  work[i*si + j*sj + k] = sum/(a1 + 27.0);
               }
            }
      if (kind >= 4)
         for (i = 1; i <= nx; i++)
            for (j = 1; j <= ny; j++)
               for (k = 1; k <= nz; k++)
                  S0C(var) = work[i*si + j*sj + k];
      for (i = 1; i <= nx; i++)                                   /* stencil_check :959-983 */
         for (j = 1; j <= ny; j++)
            for (k = 1; k <= nz; k++) {
               double x = fabs(S0C(var));
               if (x >= 1.0) {
                  x /= (a1 + a0[0] + x);
                  flops[2] += 1.0; flops[0] += 2.0;
               } else if (x < 0.1) {
                  x *= 10.0 - a1;
                  flops[1] += 1.0; flops[0] += 1.0;
               }
               S0C(var) = x;
            }
   }
   free(work);
   {
      /* flops per cell of the update itself, as booked at the end of each branch */
      double ad, mu, dv;
      if (kind == 0) {
         if (var == 1) { ad = mat; mu = mat; dv = 0; }
         else if (var < mat) { ad = 3; mu = 2; dv = 0; }
         else if (var < 2*mat) { ad = 3; mu = 3; dv = 1; }
         else if (var < 3*mat) { ad = 4; mu = 3; dv = 1; }
         else { ad = 6; mu = 6; dv = 1; }
      } else if (kind <= 3) {
         if (var == 1) { ad = mat + 1; mu = mat; dv = 1; }
         else if (var < mat) { ad = 4; mu = 2; dv = 1; }
         else { ad = 12; mu = 3; dv = 1; }
      } else if (kind == 4) { ad = 7; mu = 8; dv = 1; }
      else { ad = 27; mu = 0; dv = 1; }
      flops[0] += ad*cells*num_active; flops[1] += mu*cells*num_active; flops[2] += dv*cells*num_active;
   }
}
#undef S0V
#undef S0C

/* one stage as driver.c:73-107 drives it on one rank (no checksum):
 * for each group of comm_vars variables: comm, then the stencil per variable */
int orc_stage_local(double *data, int nx, int ny, int nz, int num_vars,
                    int comm_vars, int stencil, int num_active, const int *slots,
                    const int *level, const int *nei_level, const int *nei,
                    int stage, int permute)
{
   int start, number, var, rc;
   for (start = 0; start < num_vars; start += comm_vars) {
      number = (start + comm_vars > num_vars) ? num_vars - start : comm_vars;
      rc = orc_comm_local(data, nx, ny, nz, num_vars, stencil, num_active, slots,
                          level, nei_level, nei, start, number, stage, permute,
                          NULL);
      if (rc) return rc;
      for (var = start; var < start + number; var++)
         orc_stencil_calc(data, nx, ny, nz, num_vars, num_active, slots, var,
                          stencil);
   }
   return 0;
}
