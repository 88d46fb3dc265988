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

// split / consolidate work item: parent slot + 8 child slots
struct RefineOp {
   int parent;
   int child[8];
};

struct Geometry {
   int n[3];               // nx, ny, nz
   int str[3];             // element strides of i, j, k inside a tile
   int tile;               // (nx+2)(ny+2)(nz+2)
   long long tile_stride;  // tile padded to a multiple of 16 doubles (128 B)
   long long var_stride;   // max_blocks * tile_stride
};

// kernels (launchers) -------------------------------------------------------
void launch_ghost(const FaceOp *d_ops, int n_ops, double *pool, double *send_buf,
                  const double *recv_buf, long long pool_var_stride, int start,
                  int num, cudaStream_t s);
void launch_stencil(double *pool, const Geometry &g, const int *d_slots,
                    int num_active, int var_start, int num_vars, int stencil,
                    cudaStream_t s);
void launch_checksum(const double *pool, const Geometry &g, const int *d_slots,
                     int num_active, int var_start, int num_vars, double *d_partials,
                     double *d_sums, cudaStream_t s);
void launch_split(double *pool, const Geometry &g, const RefineOp *d_ops, int n_ops,
                  int num_vars, cudaStream_t s);
void launch_consolidate(double *pool, const Geometry &g, const RefineOp *d_ops,
                        int n_ops, int num_vars, cudaStream_t s);
void launch_pack_block(const double *pool, const Geometry &g, int slot, int num_vars,
                       double *d_payload, cudaStream_t s);
void launch_unpack_block(double *pool, const Geometry &g, int slot, int num_vars,
                         const double *d_payload, cudaStream_t s);
int stencil_smem_bytes(const Geometry &g);
bool stencil_configure(const Geometry &g, std::string &err);

}  // namespace mamr
