/* miniamr_b200 — C ABI of the B200-native miniAMR stage hot path.
 *
 * The reference (Mantevo/miniAMR, ref/) has no plugin or FFI layer: its hot
 * path is a set of plain C functions that share state through globals
 * (SURVEY.md §8b).  This header is the boundary a maintainer binds instead:
 * plain pointers and sizes, no CUDA or torch types.  Each entry point cites the
 * reference routine (file:line under ref/) it replaces.  integration/glue.c
 * shows the reference-side binding: it exports the reference's own symbols
 * comm / stencil_driver / check_sum / pack_block / unpack_block on top of this
 * ABI by marshalling the reference's globals.
 *
 * All block data lives in a device-resident pool:
 *     pool[var][slot][tile],  tile = (nx+2)(ny+2)(nz+2) doubles, k fastest,
 *     ghosts at index 0 and n+1 exactly like block.array[var][i][j][k]
 *     (block.h:52; allocation main.c:429-450), tile stride padded to 128 B.
 * There are two such pools: a stage reads a variable from its current pool and
 * writes the other one (comm + stencil fused in one pass), then they swap.
 * `slot` is the reference's index into blocks[] (0 <= slot < max_blocks).
 *
 * There is NO CPU fallback: every call fails (non-zero return, message via
 * mamr_last_error) if CUDA is unavailable.
 *
 * Threading: like the reference, one host thread per context; calls are
 * executed in call order.  Everything except mamr_check_sum*, mamr_download_*,
 * mamr_pack_block and mamr_sync is asynchronous with respect to the host.
 */
#ifndef MINIAMR_B200_H
#define MINIAMR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MAMR_ABI_VERSION 1

typedef struct mamr_ctx mamr_ctx;

/* The command-line parameters the path depends on (main.c:57-174, param.h). */
typedef struct {
   int nx, ny, nz;      /* --nx --ny --nz: cells per block edge (even, > 0)      */
   int num_vars;        /* --num_vars                                            */
   int comm_vars;       /* --comm_vars; 0 or > num_vars means num_vars
                           (main.c:711-712)                                      */
   int max_blocks;      /* --max_blocks: pool slots on this rank                 */
   int stencil;         /* --stencil: 7, 27, or 0 (the variable-work mix; needs
                           mamr_set_stencil0)                                    */
   int code;            /* --code 0|1|2: all run the code-0 exchange (same result
                           on every cell the stencil reads); 1|2
                           with --permute and --stencil 27|0 differ in the
                           reference itself -> MAMR_EUNSUPPORTED                 */
   int permute;         /* --permute (comm.c:45-55)                              */
   int device;          /* CUDA device ordinal, -1 = current device              */
   int rank, num_ranks; /* my_pe, num_pes: one process per GPU                   */
} mamr_params;

/* What the path reads of one active block (block.h:36-53). */
typedef struct {
   int slot;            /* sorted_list[in].n                                     */
   int level;           /* blocks[slot].level                                    */
   int nei_level[6];    /* W,E,S,N,D,U; -2 = domain boundary                     */
   int nei[6][2][2];    /* neighbour slot(s); negative = off rank (-1 - rank)    */
} mamr_block;

/* One direction of the reference's off-rank comm lists (comm.h:38-55,
 * built by comm_util.c:36-229).  Offsets and sizes are in doubles. */
typedef struct {
   int num_partners;        /* num_comm_partners[dir]                            */
   const int *partner;      /* comm_partner[dir][i]: rank                        */
   const int *index;        /* comm_index[dir][i]: first face of partner i       */
   const int *num;          /* comm_num[dir][i]: faces of partner i              */
   const int *send_size;    /* send_size[dir][i]                                 */
   const int *recv_size;    /* recv_size[dir][i]                                 */
   int num_cases;           /* num_cases[dir]: total faces                       */
   const int *block;        /* comm_block[dir][f]: local slot                    */
   const int *face_case;    /* comm_face_case[dir][f]                            */
   const int *send_off;     /* comm_send_off[dir][f]                             */
   const int *recv_off;     /* comm_recv_off[dir][f]                             */
} mamr_comm_dir;

/* Counters the reference's profile.c consumes (timer.h:118-131,
 * block.h:144-146); the binding adds them to the reference globals. */
typedef struct {
   long long counter_same[3], counter_diff[3], counter_bc[3];   /* comm.c:169,178,189,196 */
   long long counter_halo_send[3], counter_halo_recv[3];        /* comm.c:81,152          */
   long long counter_face_send[3], counter_face_recv[3];        /* comm.c:143,226         */
   double size_mesg_send[3], size_mesg_recv[3];                 /* bytes, comm.c:82,153;
                                                                   = NVLink bytes of ghost traffic */
   double total_fp_adds, total_fp_divs;                         /* stencil.c:100-101,142-143 */
   long long total_red;                                         /* check_sum.c:62         */
   long long kernel_launches;                                   /* CUDA kernels launched  */
   double migrate_bytes;                                        /* block payload bytes moved */
   long long ghost_regens;                                      /* ghost layers an eliding stage
                                                                   left stale and that had to be
                                                                   regenerated on demand */
   double total_fp_muls;                                        /* --stencil 0, stencil.c:162 ... */
} mamr_counters;

enum { MAMR_OK = 0, MAMR_ECUDA = 1, MAMR_EINVAL = 2, MAMR_EUNSUPPORTED = 3,
       MAMR_ETOPOLOGY = 4, MAMR_ENCCL = 5, MAMR_EP2P = 6 };

/* ---- lifecycle: replaces the block-array part of allocate()/deallocate(),
 *      main.c:429-450, 613-624 ------------------------------------------- */
int  mamr_abi_version(void);
int  mamr_create(const mamr_params *params, mamr_ctx **out);
void mamr_destroy(mamr_ctx *ctx);
const char *mamr_last_error(void);
int  mamr_sync(mamr_ctx *ctx);                       /* run and wait for all queued work */
int  mamr_get_counters(mamr_ctx *ctx, mamr_counters *out);
int  mamr_reset_counters(mamr_ctx *ctx);
long long mamr_tile_doubles(mamr_ctx *ctx);          /* (nx+2)(ny+2)(nz+2)      */
long long mamr_pool_bytes(mamr_ctx *ctx);

/* ---- block data in / out: replaces the fill loops init.c:484-495 --------
 * host layout: tiles[var][i][j][k] with ghosts, (nx+2)(ny+2)(nz+2) doubles
 * per var, num_vars tiles per block.                                        */
int mamr_upload_block(mamr_ctx *ctx, int slot, const double *tiles);
int mamr_download_block(mamr_ctx *ctx, int slot, double *tiles);
int mamr_upload_tile(mamr_ctx *ctx, int slot, int var, const double *tile);
int mamr_download_tile(mamr_ctx *ctx, int slot, int var, double *tile);
int mamr_zero_block(mamr_ctx *ctx, int slot);
/* bulk form for slots [0, num_slots): host[var - var_start][slot][tile], the
 * layout of the pool without its tile padding.  Asynchronous when `host` is
 * pinned memory: the buffer must stay valid until mamr_sync(). */
int mamr_upload_vars(mamr_ctx *ctx, int var_start, int num, int num_slots, const double *host);
int mamr_download_vars(mamr_ctx *ctx, int var_start, int num, int num_slots, double *host);
/* The state exactly as init.c:484-495 defines it -- interiors, ghost layer zero
 * -- for slots [0, num_slots): host[slot][var - var_start][nx][ny][nz], the
 * block payloads of pack.c:66-70 back to back.  Fewer bytes over PCIe than
 * whole tiles ((n/(n+2))^3), copy and scatter pipelined.  Asynchronous when
 * `host` is pinned memory (valid until mamr_sync()). */
int mamr_upload_interiors(mamr_ctx *ctx, int var_start, int num, int num_slots, const double *host);

/* ---- topology: what comm()/stencil_calc()/check_sum() read through the
 *      globals blocks[], sorted_list, sorted_index (block.h:36-77) and the
 *      comm lists (comm.h:38-55).  Call after init()/refine()/load balance,
 *      i.e. whenever the host mutated them. -------------------------------- */
int mamr_set_topology(mamr_ctx *ctx, int num_active, const mamr_block *sorted_blocks);
int mamr_set_comm_lists(mamr_ctx *ctx, const mamr_comm_dir dirs[3]);

/* ---- the stage hot path ------------------------------------------------- */
/* comm(start, num_comm, stage): comm.c:42-242 (code 0: pack_face :254-401,
 * unpack_face :1002-1150, on_proc_comm :1473-1534, on_proc_comm_diff
 * :1597-1688, apply_bc :1911-1965).  With off-rank partners num_comm <= comm_vars (the
 * message buffers hold comm_vars variables per face) and, as in driver.c:75-89, start is a
 * multiple of comm_vars: start / comm_vars also selects the receive-buffer set. */
int mamr_comm(mamr_ctx *ctx, int start, int num_comm, int stage);
/* stencil_driver(var, calc_stage): stencil.c:43-74 -> stencil_calc :76-145 */
int mamr_stencil_driver(mamr_ctx *ctx, int var, int calc_stage);
/* --stencil 0 ("variable work", stencil.c:147-983): mat, a1 and a0[mat] as init()
 * drew them (init.c:418-423; host rand() stream).  Call once after init() and
 * before the first stencil_driver(); calc_stage % 6 then selects the update
 * kind and stencil_check() follows every update (stencil.c:49-70). */
int mamr_set_stencil0(mamr_ctx *ctx, int mat, double a1, const double *a0);
/* north_star alias: stencil_calc(var) == stencil_driver(var, 0) for 7/27 */
int mamr_stencil_calc(mamr_ctx *ctx, int var);
/* the same for a run of variables in one launch */
int mamr_stencil_vars(mamr_ctx *ctx, int var_start, int num);
/* check_sum(var): check_sum.c:36-65, including the global reduction
 * (num_ranks > 1: all-reduce through the peer-memory windows, or ncclAllReduce).
 * Returns the sum in *sum. */
int mamr_check_sum(mamr_ctx *ctx, int var, double *sum);
int mamr_check_sum_vars(mamr_ctx *ctx, int var_start, int num, double *sums);
/* one whole stage as driver.c:73-89 issues it (comm per group of comm_vars,
 * stencil per variable), without the host round trips */
int mamr_stage(mamr_ctx *ctx, int stage);

/* ---- refinement / migration data movement ------------------------------- */
/* split_blocks() data copy, block.c:143-173: octant o of parent -> child o */
int mamr_split_block(mamr_ctx *ctx, int parent_slot, const int child_slots[8]);
/* consolidate_blocks() data copy, block.c:411-431 */
int mamr_consolidate_block(mamr_ctx *ctx, const int child_slots[8], int parent_slot);
/* pack_block()/unpack_block() payload, pack.c:66-70 / 103-107: interiors
 * only, var-major then i,j,k: num_vars*nx*ny*nz doubles (host memory) */
int mamr_pack_block(mamr_ctx *ctx, int slot, double *payload);
int mamr_unpack_block(mamr_ctx *ctx, int slot, const double *payload);
/* the same payload GPU-to-GPU over NCCL (replaces MPI_Send/Irecv of
 * rcb.c:237,261): both ranks call, one as sender one as receiver */
int mamr_send_block(mamr_ctx *ctx, int slot, int dest_rank);
int mamr_recv_block(mamr_ctx *ctx, int slot, int src_rank);
/* The form the drop-in uses under the reference's blocking pairwise handshake
 * (exchange(), rcb.c:207-337): the sender packs the payload into a device
 * staging area at pack_block() time (its slot is reused right away,
 * rcb.c:259-266), the receiver records (slot, source) at unpack_block() time,
 * and every rank calls mamr_flush_block_moves() at the same point of the
 * program afterwards (end of move_blocks(), rcb.c:734-): every receiver fetches its
 * payloads out of the senders' windows (peer memory) -- or, with NCCL, ONE group with
 * all sends and receives, then the unpack kernels.  Between one pair of ranks
 * the k-th staged send matches the k-th staged receive. */
int mamr_stage_send_block(mamr_ctx *ctx, int slot, int dest_rank);
int mamr_stage_recv_block(mamr_ctx *ctx, int slot, int src_rank);
int mamr_flush_block_moves(mamr_ctx *ctx);
int mamr_pending_block_moves(mamr_ctx *ctx);      /* staged and not yet flushed */

/* ---- multi-GPU, alternative transport: NCCL send/recv over NVLink (the default is the
 *      peer-memory transport below) ---------------------------------------- */
#define MAMR_NCCL_ID_BYTES 128
int mamr_nccl_get_unique_id(char id[MAMR_NCCL_ID_BYTES]);   /* rank 0, then broadcast
                                                               over the host channel */
int mamr_nccl_init(mamr_ctx *ctx, const char id[MAMR_NCCL_ID_BYTES]);
/* --send_faces (comm.c:59-77, 96-128: one MPI message per face instead of one per partner):
 * the transfer stays one message per partner, the counters counter_halo_send/recv and
 * size_mesg_send/recv follow the reference's per-face accounting.  msg_len = the host's
 * msg_len[3][4] (init.c:77-117). */
int mamr_set_message_mode(mamr_ctx *ctx, int send_faces, const int *msg_len);
int mamr_device_count(void);        /* visible CUDA devices (0 without a driver): lets a
                                       host without CUDA headers map rank -> device */

/* ---- multi-GPU without a library in the data path: peer-memory transport ----
 * Every rank owns a window in device memory (flags, check_sum slots, the receive
 * buffers of all comm groups).  The ranks exchange opaque handles over the host
 * channel (MPI_Allgather in the reference's world; comm.c:71-84,120-151 posted
 * MPI_Irecv/MPI_Isend instead) and map each other's windows: CUDA IPC between
 * processes, plain pointers between ranks of one process (loopback tests: several
 * contexts on one GPU, one host thread each).  From then on comm() stores ghost
 * messages straight into the partner's receive buffer over NVLink and check_sum()
 * all-reduces through the windows and migrated blocks are fetched out of the sender's
 * window, all ordered by system-scope flags (MAMR_TRANSPORT=nccl with mamr_nccl_init done:
 * everything over NCCL instead).
 * A peer that never answers is reported after 20 s as MAMR_EP2P by the next
 * mamr_sync()/mamr_check_sum(), never as a hung GPU. */
#define MAMR_P2P_HANDLE_BYTES 128
int mamr_p2p_get_handle(mamr_ctx *ctx, char handle[MAMR_P2P_HANDLE_BYTES]);
/* handles[num_ranks][MAMR_P2P_HANDLE_BYTES] in rank order, this rank's included */
int mamr_p2p_connect(mamr_ctx *ctx, const char *handles);

/* ---- host-only view of the halo plan (no device needed; tests and tools) --
 * The plan is what the fused stage kernel executes for one comm() call: for
 * every ghost region of every active block the place its value comes from once
 * the three direction phases of comm.c:42-242 are done, and, per phase, the ops
 * that fill the send buffers (pack_face, comm.c:254-401).  An op is 19 values:
 *   dst_base src_base dst_vs src_vs ext[3] dst_str[3] src_str[3] S F first mode
 *   dst_mem src_mem
 * (csrc/common.cuh BoxOp; mode 0 copy, 1 /4, 2 prolong /4, 3 replicate, 4 4-term
 * sum; mem 0 = block pool / the block's own tile, 1+d = message buffer of
 * direction d).  which = 0: halo ops (CSR by active block via
 * mamr_plan_block_begin), 1..3: pack ops of phase which-1. */
#define MAMR_PLAN_OP_FIELDS 19
typedef struct mamr_plan mamr_plan;
int  mamr_plan_create(const mamr_params *params, int num_active, const mamr_block *sorted_blocks,
                      const mamr_comm_dir dirs[3], int stage, mamr_plan **out);
int  mamr_plan_phase_dir(mamr_plan *plan, int phase);          /* direction of phase 0..2 */
int  mamr_plan_num_ops(mamr_plan *plan, int which);
int  mamr_plan_get_ops(mamr_plan *plan, int which, long long *fields);
int  mamr_plan_block_begin(mamr_plan *plan, int *begin);       /* num_active + 1 entries */
void mamr_plan_destroy(mamr_plan *plan);

/* ---- measurement helpers (bench.py) ------------------------------------- */
/* CUDA-event timing on the library's own stream: mark begin/end around any
 * sequence of calls; elapsed in milliseconds. */
int mamr_timer_begin(mamr_ctx *ctx);
int mamr_timer_end(mamr_ctx *ctx, float *ms);
/* accumulated device time of the stencil kernel launches since the last
 * reset (events recorded around every stencil launch when enabled) */
int mamr_kernel_timing(mamr_ctx *ctx, int enable);
int mamr_kernel_time_ms(mamr_ctx *ctx, float *stencil_ms, float *ghost_ms,
                        float *checksum_ms, long long *stencil_launches,
                        long long *ghost_launches, long long *checksum_launches);

/* Device time by kernel kind, accumulated while kernel timing is enabled
 * (mamr_kernel_timing): what the reference books with timer() around comm(),
 * stencil_driver() and check_sum() (driver.c:80-106; comm.c:128-230 for the parts of
 * comm) cannot be measured on the host when the device runs asynchronously.  wait == 0
 * returns what has finished so far without synchronising. */
typedef struct {
   double fused_ms;        /* comm + stencil_calc in one kernel (fused stage kernels)        */
   double stencil_ms;      /* stencil_calc / --stencil 0 kernels on materialised ghosts     */
   double split_ghost_ms;  /* on_proc_comm, on_proc_comm_diff, apply_bc, pack_face kernels
                              of the split path (comm.c:162-203)                            */
   double pack_ms;         /* pack_face from resolved origins (comm.c:254-401)               */
   double exchange_ms;     /* message transfer + waiting for the partners (comm.c:120-157)   */
   double unpack_ms;       /* unpack_face (comm.c:1002-1150)                                 */
   double regen_ms;        /* ghost layers / Z-face exports made real on demand              */
   double checksum_ms;     /* check_sum kernels (check_sum.c:43-56)                          */
   double allreduce_ms;    /* check_sum.c:57                                                 */
   double halo_fraction;   /* part of fused_ms that is the exchange: halo bytes / all bytes  */
} mamr_device_times;
int mamr_get_device_times(mamr_ctx *ctx, int wait, mamr_device_times *out);

#ifdef __cplusplus
}
#endif
#endif
