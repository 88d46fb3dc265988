// --stencil 0 ("variable work", stencil.c:43-74,147-983) on the device: first correct
// path.  The per-cell arithmetic lives in stencil0.cuh (shared with the host test
// harness); this file only decides who computes which cell when.
//
// Everything the updates read lies inside one block's tiles (ghost layers included,
// filled by comm() beforehand), and the reference updates variables one after the other
// (a later variable reads the new values of an earlier one): one CTA owns one block and
// walks the variables of the launch in order, with a block barrier between them.
//   kind 0        cells are independent                          -> one thread per cell
//   kinds 1-3     in place along one axis: a cell reads the NEW value at -1 and the old
//                 one at +1 on that axis, nothing else of its own variable
//                                                                -> one thread per line,
//                                                                   marching along the axis
//   kinds 4, 5    through work[] (stencil.c:664,789)            -> all cells into a scratch
//                                                                   tile, barrier, copy back
//   stencil_check follows the update of a variable; folded into the update with the lag
//                 each kind allows (the sweeps read the un-checked new value at -1)
// Tiles are read and written in global memory (L1/L2-resident: one block's tiles are a
// few hundred KB); no shared-memory staging yet -- SURVEY.md §8f-1 is correctness first.
// This translation unit is compiled with -fmad=false (build.py).
#include "common.cuh"
#include "stencil0.cuh"

namespace mamr {

namespace {

struct S0Args {
   double *pool;                 // variable 0 of slot 0
   const int *slots;
   long long var_stride, tile_stride;
   int nx, ny, nz;
   int v0, v1;                   // variables [v0, v1), all in 1 .. 4*mat-1
   int kind;
   int mat;
   double a1;
   const double *a0;             // device, [mat]
   double *work;                 // [gridDim.x][tile_stride] scratch (kinds 4, 5)
   unsigned long long *chk;      // [2]: cells stencil_check divided / scaled
};

constexpr int S0_THREADS = 128;   // all blocks of a typical mesh resident at once (sweeps use one thread per line)

__global__ void __launch_bounds__(S0_THREADS) stencil0_kernel(const S0Args A)
{
   const int tid = threadIdx.x;
   const int nx = A.nx, ny = A.ny, nz = A.nz;
   const long long SJ = nz + 2, PL = (long long)(ny + 2)*SJ, VS = A.var_stride;
   double *t0 = A.pool + (long long)A.slots[blockIdx.x]*A.tile_stride;
   double *work = A.work + (long long)blockIdx.x*A.tile_stride;
   const S0Coef c = { A.mat, A.a1, A.a0 };
   const int cells = nx*ny*nz;
   unsigned long long n_div = 0, n_mul = 0;

   // stencil_check (stencil.c:959-983) runs on the variable AFTER its update.  It is folded
   // into the update wherever no other cell can still read the unchecked value:
   //   kind 0      nobody reads another cell of the variable          -> checked on the spot
   //   kinds 1-3   only the next cell of the same line reads it       -> checked one cell late
   //   kinds 4, 5  the update lands in work[]                          -> checked while copying back
   auto checked = [&](double x) {
      int what;
      x = s0_check(x, c, &what);
      n_div += what == 1;
      n_mul += what == 2;
      return x;
   };
   for (int var = A.v0; var < A.v1; var++) {
      double *tv = t0 + (long long)var*VS;
      if (A.kind == S0_POINT) {
         for (int e = tid; e < cells; e += S0_THREADS) {
            const int i = e/(ny*nz) + 1, r = e%(ny*nz), j = r/nz + 1, k = r%nz + 1;
            const long long cell = i*PL + j*SJ + k;
            tv[cell] = checked(s0_point(t0 + cell, VS, var, c));
         }
      } else if (A.kind <= S0_SWEEP_K) {
         // lines along the sweep axis; `d` = element offset of +1 on it, `len` = cells per line
         const int len = A.kind == S0_SWEEP_I ? nx : (A.kind == S0_SWEEP_J ? ny : nz);
         const long long d = A.kind == S0_SWEEP_I ? PL : (A.kind == S0_SWEEP_J ? SJ : 1);
         const int lines = cells/len;
         for (int e = tid; e < lines; e += S0_THREADS) {
            long long cell;          // first cell of the line
            if (A.kind == S0_SWEEP_I) cell = PL + (e/nz + 1)*SJ + (e%nz + 1);
            else if (A.kind == S0_SWEEP_J) cell = (e/nz + 1)*PL + SJ + (e%nz + 1);
            else cell = (e/ny + 1)*PL + (e%ny + 1)*SJ + 1;
            for (int q = 0; q < len; q++, cell += d) {
               tv[cell] = s0_sweep(t0 + cell, VS, var, c, d);
               if (q > 0) tv[cell - d] = checked(tv[cell - d]);
            }
            tv[cell - d] = checked(tv[cell - d]);
         }
      } else {
         for (int e = tid; e < cells; e += S0_THREADS) {
            const int i = e/(ny*nz) + 1, r = e%(ny*nz), j = r/nz + 1, k = r%nz + 1;
            const long long cell = i*PL + j*SJ + k;
            work[cell] = A.kind == S0_SEVEN ? s0_seven(t0 + cell, VS, var, c, PL, SJ)
                                            : s0_twenty7(t0 + cell, VS, var, c, PL, SJ);
         }
         __syncthreads();
         for (int e = tid; e < cells; e += S0_THREADS) {
            const int i = e/(ny*nz) + 1, r = e%(ny*nz), j = r/nz + 1, k = r%nz + 1;
            const long long cell = i*PL + j*SJ + k;
            tv[cell] = checked(work[cell]);
         }
      }
      __syncthreads();
   }
   // how often stencil_check took its two branches (the reference books flops per cell)
   for (int o = 16; o > 0; o >>= 1) {
      n_div += __shfl_down_sync(0xffffffffu, n_div, o);
      n_mul += __shfl_down_sync(0xffffffffu, n_mul, o);
   }
   if ((tid & 31) == 0) {
      if (n_div) atomicAdd(&A.chk[0], n_div);
      if (n_mul) atomicAdd(&A.chk[1], n_mul);
   }
}

}  // namespace

void launch_stencil0(double *pool, const Geometry &g, const int *d_slots, int num_active, int v0, int v1,
                     int kind, int mat, double a1, const double *d_a0, double *d_work,
                     unsigned long long *d_chk, cudaStream_t s)
{
   if (num_active <= 0 || v1 <= v0) return;
   S0Args A;
   A.pool = pool; A.slots = d_slots;
   A.var_stride = g.var_stride; A.tile_stride = g.tile_stride;
   A.nx = g.n[0]; A.ny = g.n[1]; A.nz = g.n[2];
   A.v0 = v0; A.v1 = v1; A.kind = kind; A.mat = mat; A.a1 = a1; A.a0 = d_a0;
   A.work = d_work; A.chk = d_chk;
   stencil0_kernel<<<(unsigned)num_active, S0_THREADS, 0, s>>>(A);
}

}  // namespace mamr
