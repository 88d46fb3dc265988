// check_sum reduction (check_sum.c:36-65) and the data movement of
// split_blocks (block.c:143-173), consolidate_blocks (block.c:411-431) and
// pack_block/unpack_block (pack.c:66-70, 103-107).
#include "common.cuh"

namespace mamr {

// ---------------------------------------------------------------------------
// check_sum: sum of the interior of one variable over all active blocks.
// Stage 1: one CTA per (tile, variable): threads own (j,k) columns, accumulate
// along i, then a warp-shuffle tree + shared-memory tree gives one partial per
// tile.  Stage 2: one CTA per variable folds the per-tile partials in a fixed
// order.  No floating-point atomics: the result is deterministic run to run.
// The association differs from the reference's sequential sum, so parity is
// within rounding (|rel| ~ 1e-15), far inside --error_tol (check_sum.c, driver.c:97).
// HBM-bound: 8 B per interior cell.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
   return v;
}

__device__ __forceinline__ double block_sum(double v, double *red)
{
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   v = warp_sum(v);
   if (lane == 0) red[w] = v;
   __syncthreads();
   const int nw = (blockDim.x + 31) >> 5;
   double r = 0.0;
   if (w == 0) {
      r = lane < nw ? red[lane] : 0.0;
      r = warp_sum(r);
   }
   return r;   // valid in thread 0
}

__global__ void __launch_bounds__(256)
checksum_tile_kernel(const double *__restrict__ pool, const int *__restrict__ slots,
                     int num_active, int var_start, int nx, int ny, int nz,
                     long long tile_stride, long long var_stride,
                     double *__restrict__ partials)
{
   __shared__ double red[8];
   const int a = blockIdx.x%num_active;
   const int vl = blockIdx.x/num_active;
   const double *tile = pool + (long long)(var_start + vl)*var_stride +
                        (long long)slots[a]*tile_stride;
   const int sj = nz + 2, plane = (ny + 2)*sj, cells = ny*nz;
   double acc = 0.0;
   for (int c = threadIdx.x; c < cells; c += blockDim.x) {
      const int j = c/nz;
      const double *col = tile + (j + 1)*sj + (c - j*nz) + 1;
      for (int i = 1; i <= nx; i++) acc += col[(size_t)i*plane];
   }
   const double s = block_sum(acc, red);
   if (threadIdx.x == 0) partials[(size_t)vl*num_active + a] = s;
}

__global__ void __launch_bounds__(256)
checksum_final_kernel(const double *__restrict__ partials, int num_active,
                      double *__restrict__ sums)
{
   __shared__ double red[8];
   const double *p = partials + (size_t)blockIdx.x*num_active;
   double acc = 0.0;
   for (int a = threadIdx.x; a < num_active; a += blockDim.x) acc += p[a];
   const double s = block_sum(acc, red);
   if (threadIdx.x == 0) sums[blockIdx.x] = s;
}

void launch_checksum(const double *pool, const Geometry &g, const int *d_slots,
                     int num_active, int var_start, int num_vars, double *d_partials,
                     double *d_sums, cudaStream_t s)
{
   if (num_vars <= 0) return;
   if (num_active > 0) {
      const int cells = g.n[1]*g.n[2];
      int threads = cells >= 256 ? 256 : ((cells + 31)/32)*32;
      const unsigned grid = (unsigned)((long long)num_active*num_vars);
      checksum_tile_kernel<<<grid, threads, 0, s>>>(pool, d_slots, num_active, var_start,
                                                    g.n[0], g.n[1], g.n[2], g.tile_stride,
                                                    g.var_stride, d_partials);
   }
   checksum_final_kernel<<<num_vars, 256, 0, s>>>(d_partials, num_active, d_sums);
}

void launch_checksum_final(const double *d_partials, int n, int num_vars, double *d_sums, cudaStream_t s)
{
   if (num_vars <= 0) return;
   checksum_final_kernel<<<num_vars, 256, 0, s>>>(d_partials, n, d_sums);
}

// ---------------------------------------------------------------------------
// split: child o gets octant o of the parent; each parent cell / 8.0 fills the
// 2x2x2 child cells (block.c:161-173).  grid = (ops*8, vars); one thread per
// child cell.  Child ghosts are left untouched, as in the reference.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_kernel(double *__restrict__ pool, const RefineOp *__restrict__ ops, int nx, int ny,
             int nz, long long tile_stride, long long var_stride, int var0)
{
   const int var = var0 + blockIdx.y;
   const RefineOp op = ops[blockIdx.x >> 3];
   const int o = blockIdx.x & 7;
   const int sj = nz + 2, si = (ny + 2)*sj;
   const int i1 = (o & 1)*(nx/2), j1 = ((o >> 1) & 1)*(ny/2), k1 = (o >> 2)*(nz/2);
   const double *par = pool + (long long)var*var_stride + (long long)op.parent*tile_stride;
   double *ch = pool + (long long)var*var_stride + (long long)op.child[o]*tile_stride;
   const int cells = nx*ny*nz;
   for (int c = threadIdx.x; c < cells; c += blockDim.x) {
      const int i = c/(ny*nz), r = c - i*ny*nz, j = r/nz, k = r - j*nz;   // 0-based child cell
      const double p = par[(size_t)((i >> 1) + 1 + i1)*si + ((j >> 1) + 1 + j1)*sj +
                           ((k >> 1) + 1 + k1)];
      ch[(size_t)(i + 1)*si + (j + 1)*sj + (k + 1)] = p/8.0;
   }
}

void launch_split(double *pool, const Geometry &g, const RefineOp *d_ops, int n_ops,
                  int var_start, int num_vars, cudaStream_t s)
{
   if (n_ops <= 0 || num_vars <= 0) return;
   dim3 grid((unsigned)n_ops*8u, (unsigned)num_vars);
   split_kernel<<<grid, 256, 0, s>>>(pool, d_ops, g.n[0], g.n[1], g.n[2], g.tile_stride,
                                     g.var_stride, var_start);
}

// ---------------------------------------------------------------------------
// consolidate: parent cell = 8-term sum of the child cells in the order
// (i2,j2,k2) (i2+1,j2,k2) (i2,j2+1,k2) (i2+1,j2+1,k2), then the same at k2+1
// (block.c:422-430), accumulated left to right.  One thread per parent cell.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
consolidate_kernel(double *__restrict__ pool, const RefineOp *__restrict__ ops, int nx,
                   int ny, int nz, long long tile_stride, long long var_stride, int var0)
{
   const int var = var0 + blockIdx.y;
   const RefineOp op = ops[blockIdx.x];
   const int sj = nz + 2, si = (ny + 2)*sj;
   const int hx = nx/2, hy = ny/2, hz = nz/2;
   double *par = pool + (long long)var*var_stride + (long long)op.parent*tile_stride;
   const int cells = nx*ny*nz;
   for (int c = threadIdx.x; c < cells; c += blockDim.x) {
      const int i = c/(ny*nz), r = c - i*ny*nz, j = r/nz, k = r - j*nz;   // 0-based parent cell
      const int o = (i >= hx ? 1 : 0) + (j >= hy ? 2 : 0) + (k >= hz ? 4 : 0);
      const int ci = 2*(i - (i >= hx ? hx : 0)) + 1, cj = 2*(j - (j >= hy ? hy : 0)) + 1,
                ck = 2*(k - (k >= hz ? hz : 0)) + 1;
      const double *q = pool + (long long)var*var_stride +
                        (long long)op.child[o]*tile_stride + (size_t)ci*si + cj*sj + ck;
      double s = q[0] + q[si];
      s += q[sj];
      s += q[si + sj];
      s += q[1];
      s += q[si + 1];
      s += q[sj + 1];
      s += q[si + sj + 1];
      par[(size_t)(i + 1)*si + (j + 1)*sj + (k + 1)] = s;
   }
}

void launch_consolidate(double *pool, const Geometry &g, const RefineOp *d_ops, int n_ops,
                        int var_start, int num_vars, cudaStream_t s)
{
   if (n_ops <= 0 || num_vars <= 0) return;
   dim3 grid((unsigned)n_ops, (unsigned)num_vars);
   consolidate_kernel<<<grid, 256, 0, s>>>(pool, d_ops, g.n[0], g.n[1], g.n[2],
                                           g.tile_stride, g.var_stride, var_start);
}

// ---------------------------------------------------------------------------
// migration payload: interiors only, var-major then i,j,k (pack.c:66-70).
// ---------------------------------------------------------------------------
template <bool PACK>
__global__ void __launch_bounds__(256)
block_payload_kernel(double *__restrict__ pool, int slot, int nx, int ny, int nz,
                     long long tile_stride, long long var_stride, double *__restrict__ payload,
                     int var0)
{
   const int var = var0 + blockIdx.y;
   const int sj = nz + 2, si = (ny + 2)*sj;
   const int cells = nx*ny*nz;
   double *tile = pool + (long long)var*var_stride + (long long)slot*tile_stride;
   double *pl = payload + (size_t)var*cells;
   for (int c = blockIdx.x*blockDim.x + threadIdx.x; c < cells; c += gridDim.x*blockDim.x) {
      const int i = c/(ny*nz), r = c - i*ny*nz, j = r/nz, k = r - j*nz;
      const size_t t = (size_t)(i + 1)*si + (j + 1)*sj + (k + 1);
      if (PACK) pl[c] = tile[t];
      else tile[t] = pl[c];
   }
}

void launch_pack_block(const double *pool, const Geometry &g, int slot, int var_start,
                       int num_vars, double *d_payload, cudaStream_t s)
{
   if (num_vars <= 0) return;
   const int cells = g.n[0]*g.n[1]*g.n[2];
   dim3 grid((unsigned)((cells + 1023)/1024), (unsigned)num_vars);
   block_payload_kernel<true><<<grid, 256, 0, s>>>(const_cast<double *>(pool), slot, g.n[0],
                                                   g.n[1], g.n[2], g.tile_stride,
                                                   g.var_stride, d_payload, var_start);
}

void launch_unpack_block(double *pool, const Geometry &g, int slot, int var_start,
                         int num_vars, const double *d_payload, cudaStream_t s)
{
   if (num_vars <= 0) return;
   const int cells = g.n[0]*g.n[1]*g.n[2];
   dim3 grid((unsigned)((cells + 1023)/1024), (unsigned)num_vars);
   block_payload_kernel<false><<<grid, 256, 0, s>>>(pool, slot, g.n[0], g.n[1], g.n[2],
                                                    g.tile_stride, g.var_stride,
                                                    const_cast<double *>(d_payload), var_start);
}

// mamr_upload_interiors: whole tiles (ghost layer zero, interior from the staged block
// payloads) for slots [slot0, slot0 + nslots) x variables [v_first, v_first + nv)
__global__ void __launch_bounds__(256)
fill_tiles_kernel(double *pool, const double *stage, int slot0, int stage_vars, int var_start, int v_first,
                  int nx, int ny, int nz, long long tile_stride, long long var_stride)
{
   const int v = v_first + blockIdx.y;
   const int cells = nx*ny*nz, SJ = nz + 2, PL = (ny + 2)*SJ, tile = (nx + 2)*PL;
   const double *src = stage + ((size_t)blockIdx.x*stage_vars + (v - var_start))*cells;
   double *dst = pool + (long long)v*var_stride + (long long)(slot0 + blockIdx.x)*tile_stride;
   for (int e = threadIdx.x; e < tile; e += blockDim.x) {
      const int i = e/PL, r = e - i*PL, j = r/SJ, k = r - j*SJ;
      const bool in = i >= 1 && i <= nx && j >= 1 && j <= ny && k >= 1 && k <= nz;
      dst[e] = in ? src[((i - 1)*ny + (j - 1))*nz + (k - 1)] : 0.0;
   }
}

void launch_fill_tiles(double *pool, const Geometry &g, const double *d_stage, int slot0, int nslots,
                       int stage_vars, int var_start, int v_first, int nv, cudaStream_t s)
{
   if (nslots <= 0 || nv <= 0) return;
   dim3 grid((unsigned)nslots, (unsigned)nv);
   fill_tiles_kernel<<<grid, 256, 0, s>>>(pool, d_stage, slot0, stage_vars, var_start, v_first, g.n[0],
                                           g.n[1], g.n[2], g.tile_stride, g.var_stride);
}

}  // namespace mamr
