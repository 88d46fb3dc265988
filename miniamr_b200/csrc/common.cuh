// miniamr_b200 — internal declarations shared by the kernels and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/miniamr_b200.h"

namespace mamr {

// ---------------------------------------------------------------------------
// Ghost-exchange descriptor.  Every producer of ghost values in comm.c --code 0
// (on_proc_comm :1473, on_proc_comm_diff :1597, apply_bc :1911, pack_face :254,
// unpack_face :1002) is "fill a destination rectangle of one face plane (or of
// a message buffer) from a source rectangle through one of five index/value
// transforms".  The host flattens the topology into an array of these; one
// kernel launch executes a whole direction phase.
// ---------------------------------------------------------------------------
enum FaceMode : int {
   FM_COPY = 0,      // dst(a,b) = src(a,b)
   FM_DIV4 = 1,      // dst(a,b) = src(a,b)/4.0           (coarse quarter -> message)
   FM_PROLONG = 2,   // dst(a,b) = src(a>>1,b>>1)/4.0     (coarse -> fine, on rank)
   FM_REPL = 3,      // dst(a,b) = src(a>>1,b>>1)         (message -> fine ghosts)
   FM_SUM4 = 4       // dst(a,b) = src(2a,2b)+src(2a,2b+1)+src(2a+1,2b)+src(2a+1,2b+1)
};

enum FaceMem : int { MEM_DST_SEND = 1, MEM_SRC_RECV = 2 };

struct FaceOp {
   long long dst_base, src_base;  // element offset of rectangle origin (var `start`)
   long long dst_vs, src_vs;      // element stride between consecutive variables
   int dst_S, dst_F;              // element strides of the slow / fast index
   int src_S, src_F;
   int Ns, Nf;                    // destination rectangle extents
   int mode;                      // FaceMode
   int mem;                       // FaceMem bits
};

struct Geometry {
   int n[3];               // nx, ny, nz
   int str[3];             // element strides of i, j, k inside a tile
   int tile;               // (nx+2)(ny+2)(nz+2)
   long long tile_stride;  // tile padded to a multiple of 16 doubles (128 B)
   long long var_stride;   // max_blocks * tile_stride
};

// ---------------------------------------------------------------------------
// Halo-gather op of the fused stage kernel (fused.cu) and of the pack kernel of
// the multi-GPU path: "fill a 3-D box of cells from a source through one of the
// FaceMode transforms".  plan.cu resolves, for every ghost region of every
// active block, where its value comes from after the three direction phases of
// comm() (a chain of same-level / boundary hops ends in an interior cell, a
// stored ghost cell or a receive buffer) and emits one BoxOp per region.
// ---------------------------------------------------------------------------
enum BoxMem : int { BM_POOL = 0, BM_BUF0 = 1 /* +dir: send (dst) or recv (src) buffer */ };

struct BoxOp {
   long long dst_base;   // element offset of the box origin: inside the tile (fused kernel)
                         // or inside the send buffer (pack kernel), variable `start`
   long long src_base;   // element offset in the pool (slot*tile_stride + cell) or recv buffer
   long long dst_vs, src_vs;   // element stride between consecutive variables (buffers only;
                               // the pool always uses Geometry::var_stride)
   int ext[3];           // box extents along i, j, k
   int dst_str[3];       // element strides of i, j, k on the destination side
   int src_str[3];       // ... and on the source side
   int S, F;             // FM_SUM4: source strides of the slow / fast in-face axis
   int first;            // index of the op's first cell in the block's flattened halo list
   short mode;           // FaceMode
   unsigned char dst_mem, src_mem;   // BoxMem
   int flags;            // BoxFlag bits (halo plan only)
};

enum BoxFlag : int {
   BF_IDENT = 1,       // ghost region no phase writes: the op copies the cell onto itself
   BF_GHOST_SRC = 2,   // the source is a stored ghost cell (of this or another tile)
   BF_FACE = 4         // destination is a pure face region (all the 7-point stencil reads)
};

// host copy of one direction of the reference's comm lists (comm.h:38-55)
struct DirLists {
   std::vector<int> partner, index, num, send_size, recv_size;
   std::vector<int> block, face_case, send_off, recv_off;
};

// what plan.cu reads: the topology as the reference's globals describe it
struct PlanInput {
   const Geometry *g;
   int stencil;
   int max_blocks;
   const std::vector<mamr_block> *blocks;   // sorted_list order
   const DirLists *cl;                      // [3]
   int order[3];                            // direction of phase 0, 1, 2 (comm.c:51-55)
};

// per launch: CSR of ops by active block
struct HaloPlan {
   std::vector<BoxOp> ops;
   std::vector<int> begin;   // num_active + 1
   int max_ops = 0;          // largest op count of one block
   // no op reads a stored ghost cell other than BF_IDENT ones: the stage result does
   // not depend on the ghost cells in memory, so the fused kernel may leave the
   // i-ghost planes and j-ghost rows of its output unwritten (api.cu regenerates
   // them from the previous pool on demand)
   bool elidable = false;
   bool ok = false;          // every ghost region resolved without an unsupported chain
   std::string why;          // reason when !ok
};

// resolve every ghost region of every active block (fused kernel input)
void build_halo_plan(const PlanInput &in, HaloPlan &out);
// ops that fill the send buffer of direction phase `phase` from resolved sources
// (fbegin: CSR of the ops by face of the direction's comm list, faces + 1 entries)
bool build_pack_plan(const PlanInput &in, int phase, std::vector<BoxOp> &ops, std::vector<int> &fbegin,
                     std::string &why);

// split / consolidate work item: parent slot + 8 child slots
struct RefineOp {
   int parent;
   int child[8];
};


// kernels (launchers) -------------------------------------------------------
void launch_ghost(const FaceOp *d_ops, int n_ops, double *pool, double *send_buf,
                  const double *recv_buf, long long pool_var_stride, int start,
                  int num, int buf_var0, cudaStream_t s);
void launch_stencil(double *pool, const Geometry &g, const int *d_slots,
                    int num_active, int var_start, int num_vars, int stencil,
                    cudaStream_t s);
void launch_fill_tiles(double *pool, const Geometry &g, const double *d_stage, int slot0, int nslots,
                       int stage_vars, int var_start, int v_first, int nv, cudaStream_t s);
// --stencil 0 updates of variables [v0, v1) (all in 1 .. 4*mat-1), stencil0.cu
void launch_stencil0(double *pool, const Geometry &g, const int *d_slots, int num_active, int v0, int v1,
                     int kind, int mat, double a1, const double *d_a0, double *d_work,
                     unsigned long long *d_chk, cudaStream_t s);
void launch_checksum(const double *pool, const Geometry &g, const int *d_slots,
                     int num_active, int var_start, int num_vars, double *d_partials,
                     double *d_sums, cudaStream_t s);
void launch_split(double *pool, const Geometry &g, const RefineOp *d_ops, int n_ops,
                  int var_start, int num_vars, cudaStream_t s);
void launch_consolidate(double *pool, const Geometry &g, const RefineOp *d_ops,
                        int n_ops, int var_start, int num_vars, cudaStream_t s);
void launch_pack_block(const double *pool, const Geometry &g, int slot, int var_start,
                       int num_vars, double *d_payload, cudaStream_t s);
void launch_unpack_block(double *pool, const Geometry &g, int slot, int var_start,
                         int num_vars, const double *d_payload, cudaStream_t s);
// fused halo-gather + stencil: reads pool_in (read-only), writes whole tiles of pool_out
bool fused_supported(const Geometry &g, std::string &why);
bool fused_configure(const Geometry &g, std::string &err);
void launch_fused(const double *pool_in, double *pool_out, const Geometry &g,
                  const int *d_slots, const int *d_order, int num_active, const BoxOp *d_ops,
                  const int *d_begin,
                  const double *const recv[3], int var_start, int num_vars, int buf_var0,
                  int stencil, cudaStream_t s);
// compile-time block size, trimmed tile traffic (fused2.cu)
bool fused2_supported(const Geometry &g);
bool fused2_configure(const Geometry &g, std::string &err);
void launch_fused2(const double *pool_in, double *pool_out, const Geometry &g, const int *d_slots,
                   const int *d_order, int num_active, const BoxOp *d_ops, const int *d_begin,
                   const double *const recv[3], int var_start, int num_vars, int buf_var0,
                   int stencil, bool elide, const double *zf_in, double *zf_out,
                   const long long *d_zsrc, double *d_cspart, long long cs_var_stride, cudaStream_t s);
// Z-face exports (the k=1 and k=nz interior planes of every tile, packed)
void launch_zface_extract(const double *pool, double *zf, const Geometry &g, const int *d_slots,
                          int num_active, int var_start, int num_vars, cudaStream_t s);
// 7-point stencil on tiles too big for shared memory: planes streamed (slab7.cu)
bool slab7_supported(const Geometry &g);
bool slab7_configure(const Geometry &g, std::string &err);
int slab7_max_cell_ops();
void launch_slab7(const double *pool_in, double *pool_out, const Geometry &g, const int *d_slots,
                  const int *d_order, int num_active, const long long *d_fsrc, const BoxOp *d_cops,
                  const int *d_cbegin, const double *const recv[3], int var_start, int num_vars,
                  int buf_var0, const double *zf_in, double *zf_out, double *d_cspart,
                  long long cs_var_stride, cudaStream_t s);
constexpr int CS_WARPS = 8;   // check_sum partial slots per tile-variable the fused kernels fill (<= 8 compute warps)
// fold `n` partials per variable (check_sum stage 2; also used on the fused kernels' partials)
void launch_checksum_final(const double *d_partials, int n, int num_vars, double *d_sums, cudaStream_t s);
// halo ops of every active block: pool_in (+ receive buffers) -> ghost cells of pool_out
void launch_halo_fill(const BoxOp *d_ops, const int *d_begin, const int *d_slots, int num_active,
                      const double *pool_in, double *pool_out, const Geometry &g,
                      const double *const recv[3], int var_start, int num_vars, int buf_var0,
                      bool only_ident, cudaStream_t s);
// pack_face of one direction phase: the BoxOps of face f are ops[fbegin[f] .. fbegin[f+1])
void launch_facepack(const BoxOp *d_ops, const int *d_fbegin, int n_faces, const double *pool_in,
                     long long var_stride, double *const send[3], const double *const recv[3],
                     int var_start, int num_vars, int buf_var0, cudaStream_t s);
int stencil_smem_bytes(const Geometry &g);
bool stencil_configure(const Geometry &g, std::string &err);

}  // namespace mamr
