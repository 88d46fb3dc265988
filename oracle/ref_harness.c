/* TEST INFRASTRUCTURE — not part of the product path.
 *
 * Step-by-step harness around the UNMODIFIED miniAMR reference.  It is compiled
 * (oracle/Makefile) together with the reference sources where they lie under
 * /root/reference/{ref,openmp} into oracle/_ref/libminiamr_ref.so; nothing from
 * the reference is copied into this repository.  The reference's main() is
 * renamed (-Dmain=miniamr_ref_main) and its call to driver() (main.c:383) is
 * intercepted with -Wl,--wrap=driver, so that after the reference has parsed
 * its own command line and run allocate() (main.c:57-362) control comes back
 * here and the tests can drive init()/refine()/comm()/stencil_driver()/
 * check_sum() one call at a time and read the block arrays in between.
 *
 * Used to (1) pin oracle/oracle.c (the CPU restatement) against the real
 * reference, (2) generate tests/golden fixtures, (3) time the reference's own
 * CPU path for bench.py's cpu_baseline / --impl reference arm.
 */
#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mpi.h>

#include "block.h"
#include "comm.h"
#include "timer.h"
#include "proto.h"

int miniamr_ref_main(int argc, char **argv);
void __real_driver(void);

static jmp_buf refh_jb;
static int refh_mode; /* 0: stop before driver(); 1: run the real driver() */

void __wrap_driver(void)
{
   if (refh_mode == 1)
      __real_driver();
   longjmp(refh_jb, 1);
}

/* Parse argv with the reference's own parser and allocate (main.c:38-362).
 * run_driver=1 additionally runs the reference's whole driver() loop. */
int refh_start(int argc, char **argv, int run_driver)
{
   refh_mode = run_driver;
   if (!setjmp(refh_jb))
      miniamr_ref_main(argc, argv);
   return 0;
}

/* driver.c:44-57 up to (not including) the time-step loop */
void refh_init(void)
{
   init();
   init_profile();
}

void refh_refine(int ts)
{
   if (ts == 0) {
      first = 1;
      if (num_refine || uniform_refine)
         refine(0);
      first = 0;
   } else
      refine(ts);
}

void refh_move(double delta) { move(delta); }
double refh_calc_time_step(void) { return calc_time_step(); }
void refh_comm(int start, int num_comm, int stage) { comm(start, num_comm, stage); }
void refh_stencil_driver(int var, int stage) { stencil_driver(var, stage); }
double refh_check_sum(int var) { return check_sum(var); }
void refh_split_blocks(void) { split_blocks(); }
void refh_consolidate_blocks(void) { consolidate_blocks(); }

/* one full stage exactly as driver.c:73-107 (no checksum) */
void refh_stage(int stage)
{
   int start, number, var;
   total_blocks += global_active;
   for (start = 0; start < num_vars; start += comm_vars) {
      number = (start + comm_vars > num_vars) ? num_vars - start : comm_vars;
      comm(start, number, stage);
      for (var = start; var < start + number; var++)
         stencil_driver(var, stage);
   }
}

enum { P_NX, P_NY, P_NZ, P_NUM_VARS, P_COMM_VARS, P_MAX_BLOCKS, P_STENCIL,
       P_NUM_REFINE, P_NUM_ACTIVE, P_MAX_ACTIVE_BLOCK, P_CODE, P_PERMUTE,
       P_NUM_PES, P_MY_PE, P_UNIFORM, P_STAGES, P_TSTEPS, P_CHECKSUM_FREQ,
       P_REFINE_FREQ, P_NUM_PARENTS, P_MAX_ACTIVE_PARENT, P_ERROR_TOL, P_COUNT };

void refh_get_params(int *out)
{
   out[P_NX] = x_block_size; out[P_NY] = y_block_size; out[P_NZ] = z_block_size;
   out[P_NUM_VARS] = num_vars; out[P_COMM_VARS] = comm_vars;
   out[P_MAX_BLOCKS] = max_num_blocks; out[P_STENCIL] = stencil;
   out[P_NUM_REFINE] = num_refine; out[P_NUM_ACTIVE] = num_active;
   out[P_MAX_ACTIVE_BLOCK] = max_active_block; out[P_CODE] = code;
   out[P_PERMUTE] = permute; out[P_NUM_PES] = num_pes; out[P_MY_PE] = my_pe;
   out[P_UNIFORM] = uniform_refine; out[P_STAGES] = stages_per_ts;
   out[P_TSTEPS] = num_tsteps; out[P_CHECKSUM_FREQ] = checksum_freq;
   out[P_REFINE_FREQ] = refine_freq; out[P_NUM_PARENTS] = num_parents;
   out[P_MAX_ACTIVE_PARENT] = max_active_parent; out[P_ERROR_TOL] = error_tol;
}

/* active slots in sorted_list order (block.c:610-645); returns the count */
int refh_get_sorted(int *slots)
{
   int in, n = sorted_index[num_refine+1];
   if (slots)
      for (in = 0; in < n; in++)
         slots[in] = sorted_list[in].n;
   return n;
}

/* per-slot topology: out[0]=level, out[1]=refine, out[2..7]=nei_level,
 * out[8..31]=nei[6][2][2], out[32..34]=cen; returns block number (-1 free) */
long long refh_get_block(int slot, int *out)
{
   int c, i, j;
   block *bp = &blocks[slot];
   out[0] = bp->level;
   out[1] = bp->refine;
   for (c = 0; c < 6; c++) {
      out[2+c] = bp->nei_level[c];
      for (i = 0; i < 2; i++)
         for (j = 0; j < 2; j++)
            out[8 + c*4 + i*2 + j] = bp->nei[c][i][j];
   }
   out[32] = bp->cen[0]; out[33] = bp->cen[1]; out[34] = bp->cen[2];
   return (long long) bp->number;
}

void refh_set_block_refine(int slot, int refine_flag) { blocks[slot].refine = refine_flag; }

/* parents: out[0]=level, out[1]=refine, out[2..9]=child, out[10..17]=child_node */
long long refh_get_parent(int p, int *out)
{
   int o;
   parent *pp = &parents[p];
   out[0] = pp->level;
   out[1] = pp->refine;
   for (o = 0; o < 8; o++) {
      out[2+o] = (int) pp->child[o];
      out[10+o] = pp->child_node[o];
   }
   return (long long) pp->number;
}

/* integration build only (integration/glue.c): bring the device-resident block
 * data back into blocks[].array before the getters below read it */
void mamr_glue_sync_host(void) __attribute__((weak));
void refh_sync_host(void)
{
   if (mamr_glue_sync_host)
      mamr_glue_sync_host();
}

/* flatten one (slot, var) tile, k fastest, ghosts included */
void refh_get_tile(int slot, int var, double *out)
{
   int i, j, k, n = 0;
   block *bp = &blocks[slot];
   for (i = 0; i <= x_block_size+1; i++)
      for (j = 0; j <= y_block_size+1; j++)
         for (k = 0; k <= z_block_size+1; k++)
            out[n++] = bp->array[var][i][j][k];
}

void refh_set_tile(int slot, int var, const double *in)
{
   int i, j, k, n = 0;
   block *bp = &blocks[slot];
   for (i = 0; i <= x_block_size+1; i++)
      for (j = 0; j <= y_block_size+1; j++)
         for (k = 0; k <= z_block_size+1; k++)
            bp->array[var][i][j][k] = in[n++];
}

/* all vars of one slot: out[var][tile] */
void refh_get_slot(int slot, double *out)
{
   int v;
   size_t t = (size_t)(x_block_size+2)*(y_block_size+2)*(z_block_size+2);
   for (v = 0; v < num_vars; v++)
      refh_get_tile(slot, v, out + v*t);
}

void refh_set_slot(int slot, const double *in)
{
   int v;
   size_t t = (size_t)(x_block_size+2)*(y_block_size+2)*(z_block_size+2);
   for (v = 0; v < num_vars; v++)
      refh_set_tile(slot, v, in + v*t);
}

/* pack.c:34-72 / 74-108 through the reference's own send_buff/recv_buff */
int refh_pack_block(int slot, double *out, int max_doubles)
{
   int len = 50 + num_vars*num_cells;
   pack_block(slot);
   if (len > max_doubles) len = max_doubles;
   memcpy(out, send_buff, (size_t)len*sizeof(double));
   return 50 + num_vars*num_cells;
}

void refh_unpack_block(int slot, const double *in)
{
   memcpy(recv_buff, in, (size_t)(50 + num_vars*num_cells)*sizeof(double));
   unpack_block(slot);
}

/* timers (seconds) + counters after refh_start(..., run_driver=1) or manual
 * stepping: out[0]=calc, [1]=comm, [2]=checksum, [3]=refine, [4]=all,
 * [5]=total_blocks, [6]=num_tsteps, [7]=total_fp_adds, [8]=total_fp_divs */
void refh_get_timers(double *out)
{
   out[0] = timer_calc_all; out[1] = timer_comm_all; out[2] = timer_cs_all;
   out[3] = timer_refine_all; out[4] = timer_all;
   out[5] = (double) total_blocks; out[6] = (double) num_tsteps;
   out[7] = total_fp_adds; out[8] = total_fp_divs;
}

void refh_get_counters(int *out)
{
   int d;
   for (d = 0; d < 3; d++) {
      out[d] = counter_same[d];
      out[3+d] = counter_diff[d];
      out[6+d] = counter_bc[d];
   }
}

double refh_get_grid_sum(int var) { return grid_sum[var]; }
long long refh_global_active(void) { return (long long) global_active; }

/* comm.c:245 / 993 — public in proto.h:50-51, never reached at one rank.
 * (weak: the integration build replaces comm.c and has no such symbols) */
#pragma weak pack_face
#pragma weak unpack_face
void refh_pack_face(double *buf, int slot, int face_case, int dir, int start, int num_comm)
{ if (pack_face) pack_face(buf, slot, face_case, dir, start, num_comm); }
void refh_unpack_face(double *buf, int slot, int face_case, int dir, int start, int num_comm)
{ if (unpack_face) unpack_face(buf, slot, face_case, dir, start, num_comm); }

/* glibc rand() state is process-wide; reseed to the default (1) so that every
 * instance reproduces the reference's never-seeded sequence (init.c:490-494) */
void refh_reseed(void) { srand(1); }

/* ---- off-rank comm lists (comm.h:38-55) and the message exchange of one
 *      direction (comm.c:71-84,120-151), for the multi-rank planner tests ---- */

/* what: 0 partner, 1 index, 2 num, 3 send_size, 4 recv_size (per partner);
 *       5 block, 6 face_case, 7 send_off, 8 recv_off (per face).
 * Returns the element count; copies when out != NULL. */
int refh_get_comm_list(int dir, int what, int *out)
{
   int *src[9], n, i;
   src[0] = comm_partner[dir]; src[1] = comm_index[dir]; src[2] = comm_num[dir];
   src[3] = send_size[dir]; src[4] = recv_size[dir]; src[5] = comm_block[dir];
   src[6] = comm_face_case[dir]; src[7] = comm_send_off[dir]; src[8] = comm_recv_off[dir];
   n = what < 5 ? num_comm_partners[dir] : num_cases[dir];
   if (out)
      for (i = 0; i < n; i++)
         out[i] = src[what][i];
   return n;
}

/* one message per partner of direction `dir`, from `send` into `recv` (both laid
 * out like send_buff/recv_buff: partner message at the offset of its first face) */
void refh_exchange_dir(int dir, const double *send, double *recv)
{
   int i, np = num_comm_partners[dir];
   MPI_Request *rq = (MPI_Request *) malloc((size_t)(2*np + 1)*sizeof(MPI_Request));
   for (i = 0; i < np; i++)
      MPI_Irecv(recv + comm_recv_off[dir][comm_index[dir][i]], recv_size[dir][i], MPI_DOUBLE,
                comm_partner[dir][i], 900 + dir, MPI_COMM_WORLD, &rq[i]);
   for (i = 0; i < np; i++)
      MPI_Isend((void *)(send + comm_send_off[dir][comm_index[dir][i]]), send_size[dir][i], MPI_DOUBLE,
                comm_partner[dir][i], 900 + dir, MPI_COMM_WORLD, &rq[np + i]);
   for (i = 0; i < 2*np; i++)
      MPI_Wait(&rq[i], MPI_STATUS_IGNORE);
   free(rq);
}

void refh_barrier(void) { MPI_Barrier(MPI_COMM_WORLD); }

/* --stencil 0 coefficients (init.c:418-423) and the flop counters of stencil.c */
int refh_get_stencil0(double *a1_out, double *a0_out)
{
   int i;
   if (stencil) return 0;
   *a1_out = a1;
   if (a0_out)
      for (i = 0; i < mat; i++)
         a0_out[i] = a0[i];
   return mat;
}

void refh_get_flops(double *out)
{
   out[0] = total_fp_adds; out[1] = total_fp_muls; out[2] = total_fp_divs;
}
