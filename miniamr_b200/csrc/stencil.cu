// 7-point / 27-point per-variable averaging stencil over every active block:
// stencil_calc(), stencil.c:76-145.
//
// Reference semantics (Jacobi inside a block): every new interior value is
// computed from the OLD tile (interior + ghosts), then the interior is
// overwritten.  Summation order is preserved so results are bit-identical:
//   7-pt : ((((((W+S)+D)+C)+U)+N)+E)/7.0              stencil.c:88-94
//   27-pt: sb,sm,sf = 9-term sums of planes i-1,i,i+1 (j-major, k-minor);
//          ((sb+sm)+sf)/27.0                           stencil.c:111-138
// Division is a true IEEE division (no reciprocal multiply).
//
// sm_100a design.  One CTA per (tile, variable).  A tile is one contiguous
// (nx+2)(ny+2)(nz+2) run of doubles and every i-plane is a contiguous
// (ny+2)(nz+2) slab, so the tile is staged into shared memory plane by plane
// with 1-D bulk TMA copies (cp.async.bulk.shared::cluster.global, SASS UBLKCP)
// that complete on one mbarrier per plane.  Threads own (j,k) columns and march
// along i keeping the two previous plane contributions in registers, so a
// 7-pt update costs 5 shared loads and a 27-pt update 9.  The interior is
// written back in place to HBM as soon as a plane is done: the read set of the
// whole tile is either already in shared memory or (ring mode) still untouched
// in HBM, which is exactly the work[]-then-copy-back semantics of the
// reference.  Tiles that do not fit (nx = 32: 314 KB) use a ring of planes that
// is refilled behind the computation.
//
// Roofline: HBM.  Algorithmic bytes per cell-variable update:
// 16 + 8*H/n^3 (H = ghost cells the stencil needs), SURVEY.md §8(d).
#include "common.cuh"
#include "ptx.cuh"

namespace mamr {

// contribution of one staged plane to a (j,k) column
template <int STENCIL>
__device__ __forceinline__ double plane_sum9(const double *__restrict__ q, int sj)
{
   // j-major, k-minor, left to right (stencil.c:111-119)
   double s = q[-sj - 1] + q[-sj];
   s += q[-sj + 1];
   s += q[-1];
   s += q[0];
   s += q[1];
   s += q[sj - 1];
   s += q[sj];
   s += q[sj + 1];
   return s;
}

template <int STENCIL, int CPT>
__global__ void __launch_bounds__(256)
stencil_kernel(double *__restrict__ pool, const int *__restrict__ slots, int num_active,
               int var_start, int nx, int ny, int nz, long long tile_stride,
               long long var_stride, int ring)
{
   extern __shared__ __align__(128) unsigned char smem_raw[];
   const int sj = nz + 2;
   const int plane = (ny + 2)*sj;                 // doubles per i-plane
   const uint32_t plane_bytes = (uint32_t)plane*8u;
   const int nplanes = nx + 2;
   double *sm = reinterpret_cast<double *>(smem_raw);
   uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)ring*plane_bytes);

   const int tid = threadIdx.x;
   const int a = blockIdx.x%num_active;
   const int v = var_start + blockIdx.x/num_active;
   double *tile = pool + (long long)v*var_stride + (long long)slots[a]*tile_stride;

   if (tid == 0) {
      for (int r = 0; r < ring; r++) mbar_init(&full[r], 1);
      fence_barrier_init();
   }
   __syncthreads();
   if (tid == 0) {
      const int first = ring < nplanes ? ring : nplanes;
      for (int p = 0; p < first; p++) {
         mbar_arrive_expect_tx(&full[p], plane_bytes);
         bulk_g2s(sm + (size_t)p*plane, tile + (size_t)p*plane, plane_bytes, &full[p]);
      }
   }

   // the (j,k) columns this thread owns
   const int cells = ny*nz;
   int off[CPT];
   bool live[CPT];
#pragma unroll
   for (int q = 0; q < CPT; q++) {
      const int c = tid + q*blockDim.x;
      live[q] = c < cells;
      const int cc = live[q] ? c : 0;
      const int j = cc/nz;
      off[q] = (j + 1)*sj + (cc - j*nz) + 1;
   }

   mbar_wait(&full[0], 0);
   mbar_wait(&full[1%ring], 0);
   double prev[CPT], cur[CPT];
#pragma unroll
   for (int q = 0; q < CPT; q++) {
      if (STENCIL == 7) {
         prev[q] = sm[off[q]];                    // centre of plane 0  (W of i=1)
         cur[q] = sm[(size_t)(1%ring)*plane + off[q]];
      } else {
         prev[q] = plane_sum9<27>(sm + off[q], sj);
         cur[q] = plane_sum9<27>(sm + (size_t)(1%ring)*plane + off[q], sj);
      }
   }

   int slot_c = 1%ring;                            // ring slot of plane i
   for (int i = 1; i <= nx; i++) {
      const int pn = i + 1;                        // plane needed next
      const int slot_n = pn%ring;
      mbar_wait(&full[slot_n], (uint32_t)((pn/ring) & 1));
      const double *pc = sm + (size_t)slot_c*plane;
      const double *pnx = sm + (size_t)slot_n*plane;
      double *out = tile + (size_t)i*plane;
#pragma unroll
      for (int q = 0; q < CPT; q++) {
         double r;
         if (STENCIL == 7) {
            const double *c = pc + off[q];
            const double e = pnx[off[q]];
            double s = prev[q] + c[-sj];           // W + S
            s += c[-1];                            // + D
            s += cur[q];                           // + C
            s += c[1];                             // + U
            s += c[sj];                            // + N
            s += e;                                // + E
            r = s/7.0;
            prev[q] = cur[q];
            cur[q] = e;
         } else {
            const double nxt = plane_sum9<27>(pnx + off[q], sj);
            r = ((prev[q] + cur[q]) + nxt)/27.0;
            prev[q] = cur[q];
            cur[q] = nxt;
         }
         if (live[q]) out[off[q]] = r;
      }
      slot_c = slot_n;
      if (ring < nplanes) {
         // plane i-1 is dead: refill its slot with plane i-1+ring
         __syncthreads();
         const int p = i - 1 + ring;
         if (tid == 0 && p < nplanes) {
            const int sl = (i - 1)%ring;
            mbar_arrive_expect_tx(&full[sl], plane_bytes);
            bulk_g2s(sm + (size_t)sl*plane, tile + (size_t)p*plane, plane_bytes, &full[sl]);
         }
      }
   }
}

struct StencilPlan {
   int threads, cpt, ring, smem;
};

static StencilPlan make_plan(const Geometry &g)
{
   StencilPlan p;
   const int cells = g.n[1]*g.n[2];
   p.threads = cells >= 256 ? 256 : ((cells + 31)/32)*32;
   if (p.threads < 64) p.threads = 64;
   int cpt = (cells + p.threads - 1)/p.threads;
   p.cpt = 1;
   while (p.cpt < cpt) p.cpt *= 2;
   const int plane_bytes = (g.n[1] + 2)*(g.n[2] + 2)*8;
   const int nplanes = g.n[0] + 2;
   const int budget = 56*1024;                     // 4 CTAs per SM
   if (nplanes*plane_bytes <= budget)
      p.ring = nplanes;
   else {
      p.ring = budget/plane_bytes;
      if (p.ring < 4) p.ring = 4;
      if (p.ring > nplanes) p.ring = nplanes;
   }
   p.smem = p.ring*plane_bytes + p.ring*8 + 16;
   return p;
}

int stencil_smem_bytes(const Geometry &g) { return make_plan(g).smem; }

template <int STENCIL, int CPT>
static cudaError_t set_attr(int smem)
{
   return cudaFuncSetAttribute(stencil_kernel<STENCIL, CPT>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

bool stencil_configure(const Geometry &g, std::string &err)
{
   StencilPlan p = make_plan(g);
   if (p.cpt > 16) {
      err = "stencil: ny*nz too large for the column mapping (max 4096 cells per plane)";
      return false;
   }
   if (p.smem > 227*1024) {
      err = "stencil: a 4-plane ring does not fit in shared memory";
      return false;
   }
   cudaError_t e = cudaSuccess;
#define MAMR_SET(C)                                           \
   if (e == cudaSuccess && p.cpt == C) {                     \
      e = set_attr<7, C>(p.smem);                             \
      if (e == cudaSuccess) e = set_attr<27, C>(p.smem);      \
   }
   MAMR_SET(1) MAMR_SET(2) MAMR_SET(4) MAMR_SET(8) MAMR_SET(16)
#undef MAMR_SET
   if (e != cudaSuccess) {
      err = std::string("stencil: cudaFuncSetAttribute: ") + cudaGetErrorString(e);
      return false;
   }
   return true;
}

void launch_stencil(double *pool, const Geometry &g, const int *d_slots, int num_active,
                    int var_start, int num_vars, int stencil, cudaStream_t s)
{
   if (num_active <= 0 || num_vars <= 0) return;
   const StencilPlan p = make_plan(g);
   const long long total = (long long)num_active*num_vars;
   // gridDim.x < 2^31: split very large launches by variable
   const int max_vars = (int)(((1LL << 31) - 1)/num_active);
   for (int v0 = 0; v0 < num_vars; v0 += max_vars) {
      const int nv = (num_vars - v0 < max_vars) ? num_vars - v0 : max_vars;
      const unsigned grid = (unsigned)((long long)num_active*nv);
#define MAMR_GO(ST, C)                                                                       \
   stencil_kernel<ST, C><<<grid, p.threads, p.smem, s>>>(pool, d_slots, num_active,          \
                                                         var_start + v0, g.n[0], g.n[1],     \
                                                         g.n[2], g.tile_stride,              \
                                                         g.var_stride, p.ring)
#define MAMR_CASE(C)                         \
   case C:                                   \
      if (stencil == 7) MAMR_GO(7, C);       \
      else MAMR_GO(27, C);                   \
      break;
      switch (p.cpt) {
         MAMR_CASE(1) MAMR_CASE(2) MAMR_CASE(4) MAMR_CASE(8) MAMR_CASE(16)
      }
#undef MAMR_CASE
#undef MAMR_GO
   }
   (void)total;
}

}  // namespace mamr
