// Fused stage kernel for big tiles, 7-point stencil: comm() + stencil_calc() with
// the tile STREAMED through shared memory plane by plane (comm.c:42-242 +
// stencil.c:82-101).  A 32^3 tile with its ghost layer is 314 KB and does not fit
// an SM; the 7-point stencil only ever needs plane i complete (with its four
// in-plane halo lines) plus the own-column values of planes i-1 and i+1, so:
//
//  * a CTA owns one block and a group of variables; its copy warp streams planes
//    0..N+1 of every variable through a ring of R plane slots with bulk TMA copies.
//    Plane 0 / N+1 ARE the X halo: the copy simply reads the neighbour's plane N / 1
//    (its own plane 1 / N at a reflective boundary).  The Y halo of plane i is the
//    neighbour's padded row (i, N | 1, 0..N+1): one 16-byte aligned bulk copy into
//    row 0 / N+1 of the slot.  The Z halo of plane i is a row of the neighbour's
//    exported k=1 | k=N face (Z-face pool, fused2.cu): one bulk copy into a staging
//    row that the k=1 / k=N threads read instead of their k-1 / k+1 cell.
//  * a thread owns CPT (j,k) columns and marches along i with W and C in registers;
//    new values go to a ring of output plane slots (never in place), so there is no
//    block barrier at all: compute warps and copy warp meet on mbarriers only
//    (`full` plane landed, `odone` output plane written, `ofree` store drained).
//  * faces that are not a plain same-level / boundary copy (off-rank faces out of
//    the NCCL receive buffers) are filled cell by cell with cp.async by the compute
//    warps, three planes ahead.  Level boundaries are not handled here (the C ABI
//    falls back to the split path for such a plan).
//  * eliding form: only rows 1..N of planes 1..N are stored (ghost planes and ghost
//    rows stay stale and are regenerated on demand, api.cu: regen_ghosts); the
//    k-ghost cells inside the stored rows are don't-care as well.
// Tile traffic per update: 2 N^2 (N+2) + 4 N (N+2) + 2 N(N+2) [X planes] + 4 N^2
// [Z faces in and out] doubles per N^3 updates = 19.0 B at N = 32 (algorithmic
// 16 + 8*6/N = 17.5 B).  Same summation order as the reference, bit-identical.
#include "common.cuh"
#include "ptx.cuh"
#include "fused_common.cuh"

namespace mamr {

namespace {

__device__ __forceinline__ bool s7_in_range(double x)
{
   const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
   return (hi - 0x07b00000u) < (0x78400000u - 0x07b00000u);
}

__device__ __forceinline__ double s7_div7(double x)
{
   const double y = 1.0/7.0;
   const double q = x*y;
   const double r = fma(-7.0, q, x);
   return fma(r, y, q);
}

__device__ __forceinline__ void s7_arrive(uint64_t *bar)
{
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// wait until at most one bulk store group of this thread is still reading shared memory
__device__ __forceinline__ void s7_bulk_wait_read1()
{
   asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

__device__ __forceinline__ void s7_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void s7_cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

struct SlabArgs {
   FusedArgs f;
   const long long *fsrc;     // [6*num_active] plain faces: element offset, or -1
   const BoxOp *cops;         // cell ops of the faces that are not plain (FM_COPY only)
   const int *cbegin;         // CSR by active block
};

constexpr int S7_MAX_COPS = 24;

template <int N>
struct Slab {
   static constexpr int SJ = N + 2, PL = SJ*SJ;
   static constexpr int CT = 256, THREADS = CT + 32;
   static constexpr int CPT = (N*N + CT - 1)/CT;
   static constexpr int SLOT = PL + 2*N;          // plane + two Z staging rows
   static constexpr int OSLOT = N*SJ + 2*N;       // rows 1..N + two Z export rows
   static constexpr int O = 3;
   static constexpr int BUDGET = 113*1024 - 1024 - S7_MAX_COPS*(int)sizeof(SOp) - 256;
   static constexpr int R0 = (BUDGET - O*OSLOT*8)/(SLOT*8);
   static constexpr int R = R0 > 12 ? 12 : R0;
   static constexpr int SMEM = (R*SLOT + O*OSLOT)*8 + S7_MAX_COPS*(int)sizeof(SOp) + 256;
};

template <int N>
__global__ void __launch_bounds__(Slab<N>::THREADS, 2)
slab7_kernel(const SlabArgs B)
{
   using S = Slab<N>;
   constexpr int SJ = S::SJ, PL = S::PL, CT = S::CT, R = S::R, O = S::O, NP = N + 2;
   const FusedArgs &A = B.f;
   extern __shared__ __align__(128) unsigned char smem_raw[];
   double *ring = reinterpret_cast<double *>(smem_raw);
   double *oring = ring + (size_t)R*S::SLOT;
   SOp *sops = reinterpret_cast<SOp *>(oring + (size_t)O*S::OSLOT);
   uint64_t *full = reinterpret_cast<uint64_t *>(sops + S7_MAX_COPS);   // [R]
   uint64_t *odone = full + R;                                          // [O]
   uint64_t *ofree = odone + O;                                         // [O]

   const int tid = threadIdx.x;
   const int a = A.order[blockIdx.x%A.num_active];
   const int grp = blockIdx.x/A.num_active;
   const int v0 = A.var_start + grp*A.vpc;
   const int nv = min(A.vpc, A.var_end - v0);
   const int slot = A.slots[a];
   const long long slot_off = (long long)slot*A.tile_stride;
   const int cb = B.cbegin[a];
   const int ncops = B.cbegin[a + 1] - cb;

   if (tid == 0) {
      for (int r = 0; r < R; r++) mbar_init(&full[r], 1);
      for (int o = 0; o < O; o++) {
         mbar_init(&odone[o], CT/32);
         mbar_init(&ofree[o], 1);
      }
      fence_barrier_init();
   }
   stage_ops(B.cops + cb, ncops, sops, tid);
   __syncthreads();

   const int total_planes = nv*NP;     // global plane index g = t*(N+2) + p
   const int total_steps = nv*N;       // global step   index s = t*N + (i-1)

   if (tid >= CT) {
      // ------------------------------ copy warp ------------------------------
      if (tid != CT) return;
      // per face: offset (or -1) and where it points: 0 = the pool / Z-face pool,
      // 1..3 = receive buffer of direction 0..2 (a compact N x N plane per variable)
      long long fs[6];
      int fm[6];
#pragma unroll
      for (int f = 0; f < 6; f++) {
         const long long w = B.fsrc[6*a + f];
         fm[f] = w >= 0 ? (int)(w >> 56) : 0;
         fs[f] = w >= 0 ? (w & ((1LL << 56) - 1)) : w;
      }
      auto rbuf = [&](int f, int v) {
         const double *r = fm[f] == 1 ? A.recv[0] : (fm[f] == 2 ? A.recv[1] : A.recv[2]);
         return r + (long long)(v - A.buf_var0)*(N*N) + fs[f];
      };
      constexpr uint32_t ROWS_BYTES = (uint32_t)N*SJ*8u, ROW_BYTES = (uint32_t)SJ*8u,
                         ZROW_BYTES = (uint32_t)N*8u, FACE_BYTES = (uint32_t)N*N*8u;
      auto load_plane = [&](int g) {
         const int t = g/NP, p = g - t*NP;
         const int v = v0 + t;
         double *dst = ring + (size_t)(g%R)*S::SLOT;
         uint64_t *bar = &full[g%R];
         const double *pin = A.pool_in + (long long)v*A.var_stride;
         if (p == 0 || p == NP - 1) {
            const long long src = p ? fs[1] : fs[0];
            const int mem = p ? fm[1] : fm[0];
            if (src >= 0 && mem) {
               // off-rank X face: the compact plane lands behind row 0 (read through offX)
               mbar_arrive_expect_tx(bar, FACE_BYTES);
               bulk_g2s(dst + SJ, p ? rbuf(1, v) : rbuf(0, v), FACE_BYTES, bar);
            } else if (src >= 0) {
               mbar_arrive_expect_tx(bar, ROWS_BYTES);
               bulk_g2s(dst + SJ, pin + src, ROWS_BYTES, bar);
            } else
               mbar_arrive_expect_tx(bar, 0);      // filled by the compute warps
            return;
         }
         uint32_t bytes = ROWS_BYTES;
         if (fs[2] >= 0) bytes += fm[2] ? ZROW_BYTES : ROW_BYTES;
         if (fs[3] >= 0) bytes += fm[3] ? ZROW_BYTES : ROW_BYTES;
         if (fs[4] >= 0) bytes += ZROW_BYTES;
         if (fs[5] >= 0) bytes += ZROW_BYTES;
         mbar_arrive_expect_tx(bar, bytes);
         bulk_g2s(dst + SJ, pin + slot_off + (long long)p*PL + SJ, ROWS_BYTES, bar);
         // Y halo rows: the neighbour's padded row, or (off rank) row p of the compact
         // plane in the receive buffer, which lands one cell to the left (offS / offN)
         if (fs[2] >= 0) {
            if (fm[2]) bulk_g2s(dst, rbuf(2, v) + (long long)(p - 1)*N, ZROW_BYTES, bar);
            else bulk_g2s(dst, pin + fs[2] + (long long)p*PL, ROW_BYTES, bar);
         }
         if (fs[3] >= 0) {
            if (fm[3]) bulk_g2s(dst + (N + 1)*SJ, rbuf(3, v) + (long long)(p - 1)*N, ZROW_BYTES, bar);
            else bulk_g2s(dst + (N + 1)*SJ, pin + fs[3] + (long long)p*PL, ROW_BYTES, bar);
         }
         const double *zin = A.zf_in + (long long)v*A.zf_var_stride;
         if (fs[4] >= 0)
            bulk_g2s(dst + PL, (fm[4] ? rbuf(4, v) : zin + fs[4]) + (long long)(p - 1)*N, ZROW_BYTES, bar);
         if (fs[5] >= 0)
            bulk_g2s(dst + PL + N, (fm[5] ? rbuf(5, v) : zin + fs[5]) + (long long)(p - 1)*N, ZROW_BYTES,
                     bar);
      };
      int next = 0;                      // next plane to load
      for (; next < R && next < total_planes; next++) load_plane(next);
      for (int s = 0; s < total_steps; s++) {
         const int t = s/N, i = s - t*N + 1;
         const int v = v0 + t;
         mbar_wait(&odone[s%O], (uint32_t)((s/O) & 1));
         const double *src = oring + (size_t)(s%O)*S::OSLOT;
         bulk_s2g(A.pool_out + (long long)v*A.var_stride + slot_off + (long long)i*PL + SJ, src,
                  ROWS_BYTES);
         double *zout = A.zf_out + (long long)v*A.zf_var_stride + (long long)slot*A.zf_slot +
                        (long long)(i - 1)*N;
         bulk_s2g(zout, src + N*SJ, ZROW_BYTES);
         bulk_s2g(zout + N*N, src + N*SJ + N, ZROW_BYTES);
         bulk_commit();
         // every compute warp has finished step s: the planes it no longer needs
         // (centre plane i; plane 0 with i = 1; plane N+1 with i = N) are free
         const int freed = t*NP + i + (i == N ? 1 : 0);     // highest free global plane
         while (next < total_planes && next - R <= freed) load_plane(next++);
         if (s > 0) {
            s7_bulk_wait_read1();         // the store of step s-1 has left shared memory
            s7_arrive(&ofree[(s - 1)%O]);
         }
      }
      bulk_wait_read0();
      return;
   }

   // ------------------------------ compute warps ------------------------------
   // cells of the faces the copy warp does not bring (off-rank faces): plane g
   const bool has_cells = ncops > 0;
   auto issue_cells = [&](int g) {
      if (g >= total_planes) return;
      const int t = g/NP, p = g - t*NP;
      const int v = v0 + t;
      double *dst = ring + (size_t)(g%R)*S::SLOT;
      for (int o = 0; o < ncops; o++) {
         const SOp &q = sops[o];
         const int di0 = q.dst_base/PL, rem = q.dst_base - di0*PL;
         const int dj0 = rem/SJ, dk0 = rem - dj0*SJ;
         if (p < di0 || p >= di0 + q.e0) continue;
         const int aa = p - di0;
         const double *sp = (q.src_mem == BM_POOL)
                               ? A.pool_in + (long long)v*A.var_stride
                               : (q.src_mem == BM_BUF0 ? A.recv[0]
                                                       : (q.src_mem == BM_BUF0 + 1 ? A.recv[1] : A.recv[2])) +
                                    (long long)(v - A.buf_var0)*q.src_vs;
         sp += q.src_base + (long long)aa*q.ss0;
         const int cells = q.e1*q.e2;
         for (int c = tid; c < cells; c += CT) {
            const int b = c/q.e2, cc = c - b*q.e2;
            const int j = dj0 + b, k = dk0 + cc;
            // a Z-face cell lives in the staging row of its side
            double *d = (k == 0) ? dst + PL + (j - 1) : (k == N + 1 ? dst + PL + N + (j - 1) : dst + j*SJ + k);
            cp_async8(d, sp + (long long)b*q.ss1 + (long long)cc*q.ss2);
         }
      }
   };
   // Step s needs the cells of planes <= centre(s)+1.  They are requested two steps
   // ahead, one cp.async group per step, so that "all but the youngest group have
   // completed" is exactly what a step needs.  The furthest plane requested while
   // the slowest warp still reads plane centre(s)-1 is centre(s)+5 (a tile boundary
   // skips two planes): R >= 7 keeps their slots apart.
   auto centre = [&](int st) { const int tt = st/N; return tt*NP + (st - tt*N) + 1; };
   int nc = 0;                           // next plane whose cells have not been requested
   if (has_cells) {
      while (nc <= centre(0) + 1) issue_cells(nc++);
      s7_cp_async_commit();
      while (nc <= centre(1) + 1) issue_cells(nc++);
      s7_cp_async_commit();
   }

   // faces that arrive as compact planes out of a receive buffer (see the copy warp)
   bool fbuf[4];
#pragma unroll
   for (int f = 0; f < 4; f++) {
      const long long w = B.fsrc[6*a + f];
      fbuf[f] = w >= 0 && (w >> 56) != 0;
   }
   int offC[S::CPT], offD[S::CPT], offU[S::CPT], offO[S::CPT], offS[S::CPT], offN[S::CPT],
       offX[S::CPT];
   bool live[S::CPT];
#pragma unroll
   for (int q = 0; q < S::CPT; q++) {
      const int c = tid + q*CT;
      live[q] = c < N*N;
      const int cc = live[q] ? c : 0;
      const int j = cc/N + 1, k = cc%N + 1;
      offC[q] = j*SJ + k;
      offD[q] = k > 1 ? offC[q] - 1 : PL + (j - 1);
      offU[q] = k < N ? offC[q] + 1 : PL + N + (j - 1);
      offS[q] = (j == 1 && fbuf[2]) ? k - 1 : offC[q] - SJ;
      offN[q] = (j == N && fbuf[3]) ? (N + 1)*SJ + k - 1 : offC[q] + SJ;
      offX[q] = SJ + (j - 1)*N + (k - 1);
      offO[q] = (j - 1)*SJ + k;
   }

   double prev[S::CPT], cur[S::CPT];
   double cs = 0.0;                      // this thread's share of check_sum(v)
   for (int s = 0; s < total_steps; s++) {
      const int t = s/N, i = s - t*N + 1;
      const int g = t*NP + i;                               // centre plane
      if (has_cells) {
         // the cells of planes <= g+1 have landed (one younger group may be pending)
         s7_cp_async_wait1();
         named_bar_sync(1, CT);
      }
      if (i == 1) {
         mbar_wait(&full[(g - 1)%R], (uint32_t)(((g - 1)/R) & 1));
         mbar_wait(&full[g%R], (uint32_t)((g/R) & 1));
         const double *p0 = ring + (size_t)((g - 1)%R)*S::SLOT, *p1 = ring + (size_t)(g%R)*S::SLOT;
#pragma unroll
         for (int q = 0; q < S::CPT; q++) {
            prev[q] = p0[fbuf[0] ? offX[q] : offC[q]];
            cur[q] = p1[offC[q]];
         }
      }
      mbar_wait(&full[(g + 1)%R], (uint32_t)(((g + 1)/R) & 1));
      const double *pc = ring + (size_t)(g%R)*S::SLOT, *pn = ring + (size_t)((g + 1)%R)*S::SLOT;
      double r[S::CPT];
      const bool east_buf = fbuf[1] && i == N;
#pragma unroll
      for (int q = 0; q < S::CPT; q++) {
         const double e = pn[east_buf ? offX[q] : offC[q]];
         double x = prev[q] + pc[offS[q]];           // W + S
         x += pc[offD[q]];                           // + D
         x += cur[q];                                // + C
         x += pc[offU[q]];                           // + U
         x += pc[offN[q]];                           // + N
         x += e;                                     // + E
         r[q] = x;
         prev[q] = cur[q];
         cur[q] = e;
      }
      bool ok = true;
#pragma unroll
      for (int q = 0; q < S::CPT; q++) ok = ok && s7_in_range(r[q]);
      if (ok) {
#pragma unroll
         for (int q = 0; q < S::CPT; q++) r[q] = s7_div7(r[q]);
      } else {
#pragma unroll
         for (int q = 0; q < S::CPT; q++) r[q] = r[q]/7.0;
      }
      // the output slot of step s-O has been stored
      if (s >= O) mbar_wait(&ofree[s%O], (uint32_t)(((s - O)/O) & 1));
      double *po = oring + (size_t)(s%O)*S::OSLOT;
#pragma unroll
      for (int q = 0; q < S::CPT; q++) {
         if (!live[q]) continue;
         po[offO[q]] = r[q];
         cs += r[q];
         // columns k=1 and k=N are this tile's Z-face exports
         if (offD[q] >= PL) po[N*SJ + (offD[q] - PL)] = r[q];
         if (offU[q] >= PL) po[N*SJ + (offU[q] - PL)] = r[q];
      }
      if (i == N && A.cspart) {
         // last plane of variable v: one partial per warp (check_sum folds them)
         cs = cs_warp_sum(cs);
         if ((tid & 31) == 0)
            A.cspart[(long long)(v0 + t)*A.cs_var_stride + (long long)a*CS_WARPS + (tid >> 5)] = cs;
         cs = 0.0;
      }
      fence_proxy_async();
      __syncwarp();
      if ((tid & 31) == 0) s7_arrive(&odone[s%O]);
      if (has_cells) {
         while (nc <= centre(s + 2) + 1) issue_cells(nc++);
         s7_cp_async_commit();
      }
   }
}

#define MAMR_SLAB7_SIZES(X) X(32)

}  // namespace

bool slab7_supported(const Geometry &g)
{
   if (g.n[0] != g.n[1] || g.n[0] != g.n[2] || g.var_stride >= (1LL << 31)) return false;
#define X(NN) if (g.n[0] == NN) return true;
   MAMR_SLAB7_SIZES(X)
#undef X
   return false;
}

int slab7_max_cell_ops() { return S7_MAX_COPS; }

bool slab7_configure(const Geometry &g, std::string &err)
{
   if (!slab7_supported(g)) return true;
   cudaError_t e = cudaSuccess;
#define X(NN)                                                                             \
   if (g.n[0] == NN) {                                                                    \
      static_assert(Slab<NN>::R >= 7, "slab7: plane ring too short");                     \
      e = cudaFuncSetAttribute(slab7_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               Slab<NN>::SMEM);                                           \
   }
   MAMR_SLAB7_SIZES(X)
#undef X
   if (e != cudaSuccess) {
      err = std::string("slab7: cudaFuncSetAttribute: ") + cudaGetErrorString(e);
      return false;
   }
   return true;
}

void launch_slab7(const double *pool_in, double *pool_out, const Geometry &g, const int *d_slots,
                  const int *d_order, int num_active, const long long *d_fsrc, const BoxOp *d_cops,
                  const int *d_cbegin, const double *const recv[3], int var_start, int num_vars,
                  int buf_var0, const double *zf_in, double *zf_out, double *d_cspart,
                  long long cs_var_stride, cudaStream_t s)
{
   if (num_active <= 0 || num_vars <= 0) return;
   SlabArgs B;
   FusedArgs &A = B.f;
   A.cspart = d_cspart; A.cs_var_stride = cs_var_stride;
   A.pool_in = pool_in; A.pool_out = pool_out; A.slots = d_slots; A.order = d_order;
   A.ops = nullptr; A.begin = nullptr;
   for (int d = 0; d < 3; d++) A.recv[d] = recv ? recv[d] : nullptr;
   A.tile_stride = g.tile_stride; A.var_stride = g.var_stride;
   A.num_active = num_active; A.buf_var0 = buf_var0;
   A.nx = g.n[0]; A.ny = g.n[1]; A.nz = g.n[2];
   A.chunk = 0;
   A.zf_in = zf_in; A.zf_out = zf_out; A.zsrc = nullptr;
   A.zf_slot = 2*g.n[0]*g.n[1];
   A.zf_var_stride = (long long)A.zf_slot*(g.var_stride/g.tile_stride);
   B.fsrc = d_fsrc; B.cops = d_cops; B.cbegin = d_cbegin;
   static int vpc_env = -1;
   if (vpc_env < 0) {
      const char *e = getenv("MAMR_VPC");
      vpc_env = e ? atoi(e) : 0;
   }
   A.vpc = vpc_env > 0 ? vpc_env : 10;
   if (A.vpc > num_vars) A.vpc = num_vars;
   A.var_start = var_start;
   A.var_end = var_start + num_vars;
   const long long groups = (num_vars + A.vpc - 1)/A.vpc;
   const unsigned grid = (unsigned)((long long)num_active*groups);
#define X(NN) \
   if (g.n[0] == NN) slab7_kernel<NN><<<grid, Slab<NN>::THREADS, Slab<NN>::SMEM, s>>>(B);
   MAMR_SLAB7_SIZES(X)
#undef X
}

}  // namespace mamr
