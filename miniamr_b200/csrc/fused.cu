// Fused stage kernel: comm() + stencil_calc() of one variable group in ONE pass
// over the block pool (comm.c:42-242 + stencil.c:76-145).
//
// The reference exchanges ghost faces in three direction phases and then runs
// the stencil in place.  On a GPU the exchange phases are the expensive part
// (the Z phase moves one 8-byte cell per 32-byte sector).  Here every CTA owns
// one (block, variable) tile and
//   1. bulk-copies (TMA, cp.async.bulk -> SASS UBLKCP) the i-planes 1..nx of its
//      tile from the CURRENT pool into shared memory,
//   2. meanwhile pulls every ghost cell of the tile straight from its origin
//      (neighbour interior, own interior at a reflective boundary, restriction /
//      prolongation at a level boundary, a stored ghost cell, a receive buffer)
//      as resolved by plan.cu — all loads of a thread are issued before the wait
//      on the bulk copy, so both latencies overlap,
//   3. writes the gathered ghost cells to shared memory AND to the tile in the
//      NEXT pool (so stored ghosts are what the reference would hold),
//   4. marches the (j,k) columns along i exactly like stencil.cu (same summation
//      order, bit-identical) and writes the new interior to the NEXT pool.
// The current pool is read-only during the launch, which is what makes the
// neighbour reads race-free (Jacobi across blocks, as in the reference where all
// of comm() precedes stencil_calc()); the C ABI flips the variable's current
// pool afterwards.
//
// Roofline: HBM.  Per tile: read n(n+2)^2 doubles (+ halo, mostly L2 hits: the
// neighbour planes are also some other CTA's own tile), write (n+2)^3 doubles.
// Algorithmic bytes per cell-variable update: 16 + 8 H/n^3 (the
// separate ghost-exchange traffic 16 H/n^3 is fused away).
#include "common.cuh"
#include "ptx.cuh"
#include "fused_common.cuh"

namespace mamr {

namespace {

// 27-point path: a thread owns a 2x2 patch of (j,k) columns.  One i-plane of the
// patch plus its ring is a 4x4 register tile (8 aligned 128-bit shared loads);
// the four 9-term plane sums are formed from registers in the reference's order.
struct Patch {
   double t[4][4];
};

__device__ __forceinline__ void load_patch(const double *__restrict__ p, int sj, Patch &P)
{
#pragma unroll
   for (int r = 0; r < 4; r++) {
      const double2 a = *reinterpret_cast<const double2 *>(p + r*sj);
      const double2 b = *reinterpret_cast<const double2 *>(p + r*sj + 2);
      P.t[r][0] = a.x; P.t[r][1] = a.y; P.t[r][2] = b.x; P.t[r][3] = b.y;
   }
}

// j-major, k-minor, left to right (stencil.c:111-119)
__device__ __forceinline__ double patch_sum9(const Patch &P, int jj, int kk)
{
   double s = P.t[jj][kk] + P.t[jj][kk + 1];
   s += P.t[jj][kk + 2];
   s += P.t[jj + 1][kk];
   s += P.t[jj + 1][kk + 1];
   s += P.t[jj + 1][kk + 2];
   s += P.t[jj + 2][kk];
   s += P.t[jj + 2][kk + 1];
   s += P.t[jj + 2][kk + 2];
   return s;
}

__device__ __forceinline__ void patch_sums(const double *__restrict__ p, int sj, double out[4])
{
   Patch P;
   load_patch(p, sj, P);
   out[0] = patch_sum9(P, 0, 0);
   out[1] = patch_sum9(P, 0, 1);
   out[2] = patch_sum9(P, 1, 0);
   out[3] = patch_sum9(P, 1, 1);
}

// mbarrier wait with a C++-visible parity (phase) argument
template <int STENCIL, int CPT, int Q>
__global__ void __launch_bounds__(FUSED_THREADS, 2)
fused_kernel(const FusedArgs A)
{
   extern __shared__ __align__(128) unsigned char smem_raw[];
   const int nx = A.nx, ny = A.ny, nz = A.nz;
   const int sj = nz + 2;
   const int plane = (ny + 2)*sj;
   const int nplanes = nx + 2;
   const int tile = nplanes*plane;
   // two tile buffers: while variable t is computed, the tile of t+1 is in flight
   // and the finished tile of t-1 drains to global memory
   double *buf0 = reinterpret_cast<double *>(smem_raw);
   SOp *sops = reinterpret_cast<SOp *>(smem_raw + (size_t)2*tile*8);
   uint64_t *full = reinterpret_cast<uint64_t *>(sops + MAX_OPS);   // [2]

   const int tid = threadIdx.x;
   const int a = A.order[blockIdx.x%A.num_active];
   const int grp = blockIdx.x/A.num_active;
   const int v0 = A.var_start + grp*A.vpc;
   const int nv = min(A.vpc, A.var_end - v0);
   const long long slot_off = (long long)A.slots[a]*A.tile_stride;

   const int ob = A.begin[a];
   const int nops = A.begin[a + 1] - ob;
   if (tid == 0) {
      mbar_init(&full[0], 1);
      mbar_init(&full[1], 1);
      fence_barrier_init();
   }
   // stage the op table (one thread per op)
   if (tid < nops) {
      const BoxOp &g = A.ops[ob + tid];
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
   __syncthreads();
   const uint32_t in_bytes = (uint32_t)nx*plane*8u;
   if (tid == 0) {
      const double *tin = A.pool_in + (long long)v0*A.var_stride + slot_off;
      mbar_arrive_expect_tx(&full[0], in_bytes);
      for (int p = 1; p <= nx; p++)   // planes 1..nx, one bulk copy per plane
         bulk_g2s(buf0 + (size_t)p*plane, tin + (size_t)p*plane, (uint32_t)plane*8u, &full[0]);
   }

   // ---- decode this thread's share of the halo once; it is the same for every
   // variable of the block (only the variable offset changes) ----
   const int last = nops - 1;
   const int E = nops > 0 ? sops[last].first + sops[last].e0*sops[last].e1*sops[last].e2 : 0;
   // per halo cell: source offset (elements, from the variable's base; the host
   // guarantees it fits 32 bits) and destination offset | op << 20 | mode << 26 |
   // src_mem << 29
   int soff[Q], dinfo[Q];
#pragma unroll
   for (int q = 0; q < Q; q++) {
      const int e = tid + q*FUSED_THREADS;
      soff[q] = 0;
      dinfo[q] = -1;
      if (e < E) {
         int lo = 0, hi = last;
         while (lo < hi) {                       // last op with first <= e
            const int mid = (lo + hi + 1) >> 1;
            if (sops[mid].first <= e) lo = mid; else hi = mid - 1;
         }
         const SOp &s = sops[lo];
         int r = e - s.first;
         const int c = r%s.e2; r /= s.e2;
         const int b = r%s.e1;
         const int aa = r/s.e1;
         const int dsto = s.dst_base + aa*s.ds0 + b*s.ds1 + c*s.ds2;
         const int mode = s.mode;
         long long o;
         if (mode == FM_COPY) o = (long long)aa*s.ss0 + b*s.ss1 + c*s.ss2;
         else if (mode == FM_SUM4) o = (long long)(2*aa)*s.ss0 + (2*b)*s.ss1 + (2*c)*s.ss2;
         else o = (long long)(aa >> 1)*s.ss0 + (b >> 1)*s.ss1 + (c >> 1)*s.ss2;
         soff[q] = (int)(s.src_base + o);
         dinfo[q] = dsto | (lo << 20) | (mode << 26) | (s.src_mem << 29);
      }
   }

   // the (j,k) columns this thread owns
   const int cells = ny*nz;
   int off[CPT];
   bool live[CPT];
#pragma unroll
   for (int q = 0; q < CPT; q++) {
      const int c = tid + q*FUSED_THREADS;
      live[q] = c < cells;
      const int cc = live[q] ? c : 0;
      const int j = cc/nz;
      off[q] = (j + 1)*sj + (cc - j*nz) + 1;
   }

   // 27-point path: thread -> (chunk of planes, 2x2 patch)
   const int hk = nz >> 1, groups = (ny >> 1)*hk;
   const int pchunk = tid/groups, pg = tid - pchunk*groups;
   const bool pact = STENCIL != 7 && pchunk*A.chunk < nx;
   const int poff = (2*(pg/hk))*sj + 2*(pg%hk);   // row 2jp, column 2kp: the ring's corner

   // pull this thread's halo cells of variable v into tile buffer `dst`
   auto gather = [&](int v, double *dst) {
      const double *pin = A.pool_in + (long long)v*A.var_stride;
#pragma unroll
      for (int q = 0; q < Q; q++) {
         if (dinfo[q] >= 0) {
            const int mode = (dinfo[q] >> 26) & 7, mem = dinfo[q] >> 29, op = (dinfo[q] >> 20) & 63;
            const double *p = pin;
            if (mem != BM_POOL)
               p = A.recv[mem - BM_BUF0] + (long long)(v - A.buf_var0)*sops[op].src_vs;
            p += soff[q];
            double *d = dst + (dinfo[q] & 0xfffff);
            if (mode == FM_COPY || mode == FM_REPL)
               cp_async8(d, p);
            else if (mode == FM_PROLONG)
               *d = __ldg(p)/4.0;
            else {   // FM_SUM4, left to right, slow index outer (comm.c:1626-1629)
               const int S = sops[op].S, F = sops[op].F;
               double x = __ldg(p) + __ldg(p + F);
               x += __ldg(p + S);
               x += __ldg(p + S + F);
               *d = x;
            }
         }
      }
   };
   bool pre = false;   // this thread's halo of the coming variable is already on its way
   // called once per variable from inside the stencil loop
   auto prefetch_halo = [&](int t) {
      if (t + 1 < nv && !pre && mbar_test(&full[(t + 1) & 1], (uint32_t)(((t + 1) >> 1) & 1))) {
         gather(v0 + t + 1, buf0 + (size_t)((t + 1) & 1)*tile);
         pre = true;
      }
   };

   for (int t = 0; t < nv; t++) {
      const int v = v0 + t;
      double *sm = buf0 + (size_t)(t & 1)*tile;
      // ---- halo gather for variable v.  Plain copies go global -> shared as
      // 8-byte cp.async (no registers, asynchronous); they are normally issued
      // during the PREVIOUS variable's stencil, as soon as this tile's bulk copy
      // has landed (it would otherwise overwrite the ghost cells afterwards).
      if (!pre) {
         mbar_wait(&full[t & 1], (uint32_t)((t >> 1) & 1));
         gather(v, sm);
      }
      pre = false;
      cp_async_wait_all();
      __syncthreads();
      // prefetch the next variable's tile into the other buffer once the bulk store
      // that last read that buffer (variable t-1) has drained
      if (tid == 0 && t + 1 < nv) {
         bulk_wait_read0();
         double *nb = buf0 + (size_t)((t + 1) & 1)*tile;
         const double *tin = A.pool_in + (long long)(v + 1)*A.var_stride + slot_off;
         mbar_arrive_expect_tx(&full[(t + 1) & 1], in_bytes);
         for (int p = 1; p <= nx; p++)
            bulk_g2s(nb + (size_t)p*plane, tin + (size_t)p*plane, (uint32_t)plane*8u,
                     &full[(t + 1) & 1]);
      }

      // ---- stencil.  The new value of a plane replaces the old one in shared
      // memory once every thread has finished reading that plane (one barrier per
      // step); the finished tile -- gathered ghosts + new interior -- then leaves
      // as ONE bulk copy.
      if (STENCIL == 7) {
         // march the (j,k) columns along i, W and C in registers
         double prev[CPT], cur[CPT];
#pragma unroll
         for (int q = 0; q < CPT; q++) {
            prev[q] = sm[off[q]];
            cur[q] = sm[plane + off[q]];
         }
         for (int i = 1; i <= nx; i++) {
            double *pc = sm + (size_t)i*plane;
            const double *pn = pc + plane;
            double r[CPT];
#pragma unroll
            for (int q = 0; q < CPT; q++) {
               const double *c = pc + off[q];
               const double e = pn[off[q]];
               double s = prev[q] + c[-sj];           // W + S
               s += c[-1];                            // + D
               s += cur[q];                           // + C
               s += c[1];                             // + U
               s += c[sj];                            // + N
               s += e;                                // + E
               r[q] = div_const<7>(s);
               prev[q] = cur[q];
               cur[q] = e;
            }
            __syncthreads();       // plane i has been read by everyone
#pragma unroll
            for (int q = 0; q < CPT; q++)
               if (live[q]) pc[off[q]] = r[q];
            if (i == (nx >> 1) || i == nx) prefetch_halo(t);
         }
      } else {
         // thread = (chunk of i-planes, 2x2 patch of columns); sb/sm/sf of
         // stencil.c:111-138 are the plane sums of planes i-1, i, i+1
         const int CH = A.chunk;
         const int i0 = pchunk*CH + 1;
         const int ilast = min(nx, i0 + CH - 1);
         double prev[4], cur[4], lst[4];
         if (pact) {
            const double *pp = sm + poff;
            patch_sums(pp + (size_t)(i0 - 1)*plane, sj, prev);
            patch_sums(pp + (size_t)i0*plane, sj, cur);
            // plane ilast+1 is the next chunk's first output plane: read it now
            patch_sums(pp + (size_t)(ilast + 1)*plane, sj, lst);
         }
         for (int s = 0; s < CH; s++) {
            const int i = i0 + s;
            double r[4];
            const bool on = pact && i <= ilast;
            if (on) {
               double nxt[4];
               if (i == ilast) {
#pragma unroll
                  for (int u = 0; u < 4; u++) nxt[u] = lst[u];
               } else
                  patch_sums(sm + poff + (size_t)(i + 1)*plane, sj, nxt);
#pragma unroll
               for (int u = 0; u < 4; u++) {
                  r[u] = div_const<27>((prev[u] + cur[u]) + nxt[u]);
                  prev[u] = cur[u];
                  cur[u] = nxt[u];
               }
            }
            __syncthreads();       // plane i has been read by everyone
            if (on) {
               double *o = sm + poff + (size_t)i*plane + sj + 1;
               o[0] = r[0]; o[1] = r[1];
               o[sj] = r[2]; o[sj + 1] = r[3];
            }
            if (s == (CH >> 1) || s == CH - 1) prefetch_halo(t);
         }
      }
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
         bulk_s2g(A.pool_out + (long long)v*A.var_stride + slot_off, sm, (uint32_t)tile*8u);
         bulk_commit();
      }
   }
   if (tid == 0) bulk_wait_read0();
}

struct FusedPlan {
   int cpt, q, smem, chunk;
   bool ok;
};

FusedPlan make_plan(const Geometry &g)
{
   FusedPlan p;
   const int cells = g.n[1]*g.n[2];
   const int cpt = (cells + FUSED_THREADS - 1)/FUSED_THREADS;
   p.cpt = 1;
   while (p.cpt < cpt) p.cpt *= 2;
   const int halo = g.tile - g.n[0]*g.n[1]*g.n[2];
   const int q = (halo + FUSED_THREADS - 1)/FUSED_THREADS;
   p.q = q <= 4 ? 4 : (q <= 8 ? 8 : (q <= 12 ? 12 : 16));
   p.smem = 2*g.tile*8 + MAX_OPS*(int)sizeof(SOp) + 32;
   // 27-point path: smallest chunk of planes such that (patches x chunks) fits the CTA
   const int groups = (g.n[1]/2)*(g.n[2]/2);
   p.chunk = 1;
   while (groups*((g.n[0] + p.chunk - 1)/p.chunk) > FUSED_THREADS && p.chunk < g.n[0]) p.chunk++;
   p.ok = p.cpt <= 4 && q <= 16 && p.smem <= 113*1024 &&                  // two CTAs per SM
          groups*((g.n[0] + p.chunk - 1)/p.chunk) <= FUSED_THREADS &&
          g.tile < (1 << 20) && g.var_stride < (1LL << 31);
   return p;
}

template <int ST, int C, int Q>
cudaError_t set_attr(int smem)
{
   return cudaFuncSetAttribute(fused_kernel<ST, C, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               smem);
}

#define MAMR_FUSED_VARIANTS(X) \
   X(1, 4) X(1, 8) X(1, 12) X(1, 16) X(2, 4) X(2, 8) X(2, 12) X(2, 16) X(4, 4) X(4, 8) X(4, 12) X(4, 16)

}  // namespace

bool fused_supported(const Geometry &g, std::string &why)
{
   const FusedPlan p = make_plan(g);
   if (!p.ok) why = "tile does not fit the whole-tile fused kernel (shared memory / column mapping)";
   return p.ok;
}

bool fused_configure(const Geometry &g, std::string &err)
{
   const FusedPlan p = make_plan(g);
   if (!p.ok) return true;   // the split path serves this geometry
   cudaError_t e = cudaSuccess;
#define X(C, QQ)                                                    \
   if (e == cudaSuccess && p.cpt == C && p.q == QQ) {               \
      e = set_attr<7, C, QQ>(p.smem);                               \
      if (e == cudaSuccess) e = set_attr<27, C, QQ>(p.smem);        \
   }
   MAMR_FUSED_VARIANTS(X)
#undef X
   if (e != cudaSuccess) {
      err = std::string("fused: cudaFuncSetAttribute: ") + cudaGetErrorString(e);
      return false;
   }
   return true;
}

void launch_fused(const double *pool_in, double *pool_out, const Geometry &g, const int *d_slots,
                  const int *d_order, int num_active, const BoxOp *d_ops, const int *d_begin,
                  const double *const recv[3], int var_start, int num_vars, int buf_var0,
                  int stencil, cudaStream_t s)
{
   if (num_active <= 0 || num_vars <= 0) return;
   const FusedPlan p = make_plan(g);
   FusedArgs A;
   A.cspart = nullptr; A.cs_var_stride = 0;
   A.pool_in = pool_in; A.pool_out = pool_out; A.slots = d_slots; A.order = d_order;
   A.ops = d_ops; A.begin = d_begin;
   for (int d = 0; d < 3; d++) A.recv[d] = recv ? recv[d] : nullptr;
   A.tile_stride = g.tile_stride; A.var_stride = g.var_stride;
   A.num_active = num_active; A.buf_var0 = buf_var0;
   A.nx = g.n[0]; A.ny = g.n[1]; A.nz = g.n[2];
   A.chunk = p.chunk;
   // one CTA per (block, group of `vpc` variables)
   static int vpc_env = -1;
   if (vpc_env < 0) {
      const char *e = getenv("MAMR_VPC");
      vpc_env = e ? atoi(e) : 0;
   }
   A.vpc = vpc_env > 0 ? vpc_env : 8;
   if (A.vpc > num_vars) A.vpc = num_vars;
   A.var_start = var_start;
   A.var_end = var_start + num_vars;
   const long long groups = (num_vars + A.vpc - 1)/A.vpc;
   const unsigned grid = (unsigned)((long long)num_active*groups);
#define X(C, QQ)                                                                       \
   if (p.cpt == C && p.q == QQ) {                                                      \
      if (stencil == 7) fused_kernel<7, C, QQ><<<grid, FUSED_THREADS, p.smem, s>>>(A); \
      else fused_kernel<27, C, QQ><<<grid, FUSED_THREADS, p.smem, s>>>(A);             \
   }
   MAMR_FUSED_VARIANTS(X)
#undef X
}

// ---------------------------------------------------------------------------
// pack_face (comm.c:254-401) of one direction phase from resolved origins: grid = (faces,
// groups of PACK_VPC variables).  A face's message is filled by 1..9 BoxOps (its interior
// run plus, for widened faces, the ghost rows / corners an earlier phase delivered --
// possibly out of an earlier receive buffer); the CTA walks them for its variables, so a
// one-cell corner op costs one loop trip instead of a CTA of its own.
// ---------------------------------------------------------------------------
// 128 threads and at most 50 registers: a CTA must fit beside the two resident CTAs of the
// interior blocks' stage kernel (~10 K registers of an SM stay free), or packing only advances as
// those retire.
constexpr int PACK_VPC = 4;

__global__ void __launch_bounds__(128, 10)
facepack_kernel(const BoxOp *__restrict__ ops, const int *__restrict__ fbegin,
                const double *__restrict__ pool_in, long long var_stride, double *send0, double *send1,
                double *send2, const double *recv0, const double *recv1, const double *recv2,
                int var_start, int var_end, int buf_var0)
{
   const int v0 = var_start + blockIdx.y*PACK_VPC;
   const int nv = min(PACK_VPC, var_end - v0);
   double *sends[3] = { send0, send1, send2 };
   const double *recvs[3] = { recv0, recv1, recv2 };
   for (int o = fbegin[blockIdx.x]; o < fbegin[blockIdx.x + 1]; o++) {
      const BoxOp op = ops[o];
      double *dst0 = sends[op.dst_mem - BM_BUF0] + op.dst_base;
      const bool pool = op.src_mem == BM_POOL;
      const double *src0 = (pool ? pool_in : recvs[op.src_mem - BM_BUF0]) + op.src_base;
      const long long svs = pool ? var_stride : op.src_vs;
      const int e1 = op.ext[1], e2 = op.ext[2];
      const int n = op.ext[0]*e1*e2;
      // PACK_U independent elements per thread and trip: the loads of a trip are all in flight
      // before the first store (the kernel runs beside the stage kernel of the interior blocks,
      // which saturates HBM; with one load at a time per thread it crawls)
      constexpr int PACK_U = 4;
      for (int w0 = threadIdx.x; w0 < n*nv; w0 += blockDim.x*PACK_U) {
         double x[PACK_U];
         long long doff[PACK_U];
#pragma unroll
         for (int u = 0; u < PACK_U; u++) {
            const int w = w0 + u*blockDim.x;
            x[u] = 0.0;
            doff[u] = -1;
            if (w >= n*nv) continue;
            const int t = w/n;
            int r = w - t*n;
            const int c = r%e2; r /= e2;
            const int b = r%e1;
            const int a = r/e1;
            const int v = v0 + t;
            const double *src = src0 + (pool ? (long long)v : (long long)(v - buf_var0))*svs;
            if (op.mode == FM_COPY || op.mode == FM_DIV4) {
               x[u] = __ldg(src + (long long)a*op.src_str[0] + b*op.src_str[1] + c*op.src_str[2]);
               if (op.mode == FM_DIV4) x[u] = x[u]/4.0;
            } else if (op.mode == FM_PROLONG || op.mode == FM_REPL) {
               x[u] = __ldg(src + (long long)(a >> 1)*op.src_str[0] + (b >> 1)*op.src_str[1] + (c >> 1)*op.src_str[2]);
               if (op.mode == FM_PROLONG) x[u] = x[u]/4.0;
            } else {
               const double *p = src + (long long)(2*a)*op.src_str[0] + (2*b)*op.src_str[1] +
                                 (2*c)*op.src_str[2];
               double y = __ldg(p) + __ldg(p + op.F);
               y += __ldg(p + op.S);
               y += __ldg(p + op.S + op.F);
               x[u] = y;
            }
            doff[u] = (long long)(v - buf_var0)*op.dst_vs + (long long)a*op.dst_str[0] + b*op.dst_str[1] +
                      c*op.dst_str[2];
         }
#pragma unroll
         for (int u = 0; u < PACK_U; u++)
            if (doff[u] >= 0) dst0[doff[u]] = x[u];
      }
   }
}

void launch_facepack(const BoxOp *d_ops, const int *d_fbegin, int n_faces, const double *pool_in,
                     long long var_stride, double *const send[3], const double *const recv[3],
                     int var_start, int num_vars, int buf_var0, cudaStream_t s)
{
   if (n_faces <= 0 || num_vars <= 0) return;
   dim3 grid((unsigned)n_faces, (unsigned)((num_vars + PACK_VPC - 1)/PACK_VPC));
   facepack_kernel<<<grid, 128, 0, s>>>(d_ops, d_fbegin, pool_in, var_stride, send[0], send[1], send[2],
                                        recv[0], recv[1], recv[2], var_start, var_start + num_vars, buf_var0);
}

}  // namespace mamr
