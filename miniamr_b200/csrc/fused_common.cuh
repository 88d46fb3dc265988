// Declarations shared by the fused stage kernels (fused.cu: any geometry;
// fused2.cu: compile-time block size, trimmed tile traffic).
#pragma once
#include "common.cuh"

namespace mamr {

constexpr int FUSED_THREADS = 256;
constexpr int MAX_OPS = 64;          // ops per block staged in shared memory

struct FusedArgs {
   const double *pool_in;
   double *pool_out;
   const int *slots;
   const int *order;      // processing order: CTA -> active block index
   const BoxOp *ops;
   const int *begin;
   const double *recv[3];
   long long tile_stride, var_stride;
   int num_active, var_start, var_end, vpc, buf_var0;
   int nx, ny, nz;
   int chunk;             // 27-point path: i-planes per thread
   // Z-face exports (fused2.cu, eliding launches): every tile's k=1 and k=nz interior
   // planes, packed [slot][side][i-1][j-1] per variable; zsrc[2a+side] = element
   // offset of the export that IS the Z halo face `side` of active block a, or -1
   const double *zf_in;
   double *zf_out;
   const long long *zsrc;
   long long zf_var_stride;
   int zf_slot;
   // check_sum() fused into the stage (fused2.cu, slab7.cu): every compute warp adds up
   // the interior values it has just produced and stores ONE partial per tile-variable,
   // cspart[var*cs_var_stride + a*CS_WARPS + warp]; check_sum then only folds them
   // (check_sum.c:36-65 without a second pass over the blocks).  nullptr = off.
   double *cspart;
   long long cs_var_stride;
};


__device__ __forceinline__ double cs_warp_sum(double v)
{
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
   return v;
}

// compact copy of a BoxOp in shared memory
struct SOp {
   long long src_base, src_vs;
   int first;                 // flattened index of its first element
   int dst_base;
   int e0, e1, e2;            // extents along i, j and k
   int ds0, ds1, ds2;         // destination strides (tile strides)
   int ss0, ss1, ss2;
   int S, F;
   int mode, src_mem;
};


// stage one block's op table in shared memory (one thread per op)
__device__ __forceinline__ void stage_ops(const BoxOp *__restrict__ ops, int nops, SOp *sops, int tid)
{
   if (tid < nops) {
      const BoxOp &g = ops[tid];
      SOp s;
      s.src_base = g.src_base; s.src_vs = g.src_vs;
      s.first = g.first; s.dst_base = (int)g.dst_base;
      s.e0 = g.ext[0]; s.e1 = g.ext[1]; s.e2 = g.ext[2];
      s.ds0 = g.dst_str[0]; s.ds1 = g.dst_str[1]; s.ds2 = g.dst_str[2];
      s.ss0 = g.src_str[0]; s.ss1 = g.src_str[1]; s.ss2 = g.src_str[2];
      s.S = g.S; s.F = g.F;
      s.mode = g.mode; s.src_mem = g.src_mem;
      sops[tid] = s;
   }
}

}  // namespace mamr
