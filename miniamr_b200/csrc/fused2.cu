// Fused stage kernel, compile-time block size: comm() + stencil_calc() of one
// variable group in one pass (comm.c:42-242 + stencil.c:76-145), the product hot
// path for the cubic block sizes it is instantiated for.  Same contract as
// fused.cu (read-only input pool, halo cells pulled from their resolved origins,
// bit-identical summation order); what differs:
//
//  * Trimmed tile traffic.  Of a tile's (N+2)^3 cells only rows 1..N of planes
//    1..N (k = 0..N+1, contiguous per plane) are bulk-loaded: every other cell of
//    the shared-memory tile is a ghost cell the halo gather writes anyway.  With
//    ELIDE the same region is all that is stored: the i-ghost planes and j-ghost
//    rows of the output tile stay unwritten and the C ABI regenerates them from
//    the previous pool when somebody asks for them (api.cu: regen_ghosts).  The
//    k-ghost cells inside the stored rows hold their gathered values.
//    Tile traffic per update: 2 N^2 (N+2) doubles instead of N(N+2)^2 + (N+2)^3.
//  * Read everything, one barrier, write everything.  A thread keeps the new
//    values of its cells in registers until every thread has finished reading the
//    old tile; the in-place update needs ONE block barrier per tile instead of one
//    per plane.
//  * All strides are immediates; the halo gather of a block whose ops are plain
//    pool copies (no level boundary, no receive buffer) is a table-driven loop of
//    8-byte cp.async.
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"
#include "fused_common.cuh"

#ifndef MAMR_CT128_FROM
#define MAMR_CT128_FROM 10
#endif
#ifndef MAMR_REGS_SMALL
#define MAMR_REGS_SMALL 96      // register budget per thread that sizes CTAs per SM for N < 14
#endif

namespace mamr {

namespace {

// exponent test on the high word: 2^-900 <= |x| < 2^901 (integer pipe, not FP64)
__device__ __forceinline__ bool div_in_range(double x)
{
   const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
   return (hi - 0x07b00000u) < (0x78400000u - 0x07b00000u);
}

// x / D, correctly rounded (proof: ptx.cuh div_const); `fast` says x is in range
template <int D>
__device__ __forceinline__ double div_fast(double x)
{
   const double y = 1.0/(double)D;
   const double q = x*y;
   const double r = fma(-(double)D, q, x);
   return fma(r, y, q);
}

// the four 9-term sums of a 2x2 patch of columns in one i-plane, j-major then k,
// left to right (stencil.c:111-119); p = corner of the 4x4 ring
template <int SJ>
__device__ __forceinline__ void patch_sums4(const double *__restrict__ p, double out[4])
{
   double t[4][4];
#pragma unroll
   for (int r = 0; r < 4; r++) {
      const double2 a = *reinterpret_cast<const double2 *>(p + r*SJ);
      const double2 b = *reinterpret_cast<const double2 *>(p + r*SJ + 2);
      t[r][0] = a.x; t[r][1] = a.y; t[r][2] = b.x; t[r][3] = b.y;
   }
#pragma unroll
   for (int jj = 0; jj < 2; jj++)
#pragma unroll
      for (int kk = 0; kk < 2; kk++) {
         double s = t[jj][kk] + t[jj][kk + 1];
         s += t[jj][kk + 2];
         s += t[jj + 1][kk];
         s += t[jj + 1][kk + 1];
         s += t[jj + 1][kk + 2];
         s += t[jj + 2][kk];
         s += t[jj + 2][kk + 1];
         s += t[jj + 2][kk + 2];
         out[jj*2 + kk] = s;
      }
}

template <int N>
struct Shape {
   static constexpr int SJ = N + 2, PL = SJ*SJ, TILE = (N + 2)*PL;
   static constexpr int HALO = TILE - N*N*N;
   // compute threads per CTA: small tiles take fewer threads and more CTAs per SM
   static constexpr int CT = N >= 14 ? 256 : (N >= MAMR_CT128_FROM ? 128 : 64);
   static constexpr int THREADS = CT + 32;        // + the copy warp
   static constexpr int Q = (HALO + CT - 1)/CT;
   // Z-face cells (k = 0 and N+1 of rows 1..N, planes 1..N) lie inside the rows the
   // bulk copy writes: they are gathered into a staging area behind the tile and
   // patched in once the tile has landed
   static constexpr int ZST = 2*N*N;
   static constexpr int ZPT = (ZST + CT - 1)/CT;
   // ... and the k=1 / k=N planes of the updated tile are packed into a second area
   // of the same size, from where they are exported (Z-face pool)
   static constexpr int TB = TILE + 2*ZST;        // doubles per buffer
   // 27-point: 2x2 patches x chunks of CH planes
   static constexpr int GROUPS = (N/2)*(N/2);
   static constexpr int NCH = CT/GROUPS > N ? N : CT/GROUPS;
   static constexpr int CH = (N + NCH - 1)/NCH;
   // 7-point: (j,k) columns
   static constexpr int CPT = (N*N + CT - 1)/CT;
   static constexpr int SMEM = 2*TB*8 + MAX_OPS*(int)sizeof(SOp) + 64;
   // CTAs per SM: shared memory (1 KB reserved per CTA), 96 registers per thread
   static constexpr int BY_SMEM = (228*1024)/(SMEM + 128 + 1024);
   static constexpr int BY_REGS = 65536/((N >= 14 ? 96 : MAMR_REGS_SMALL)*THREADS);
   static constexpr int CTAS = BY_SMEM < BY_REGS ? (BY_SMEM < 8 ? BY_SMEM : 8) : (BY_REGS < 8 ? BY_REGS : 8);
};


__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int ST, int N, bool ELIDE>
__global__ void __launch_bounds__(Shape<N>::THREADS, Shape<N>::CTAS)
fused2_kernel(const FusedArgs A)
{
   using S = Shape<N>;
   constexpr int SJ = S::SJ, PL = S::PL, TILE = S::TILE, Q = S::Q, TB = S::TB, CT = S::CT;
   extern __shared__ __align__(128) unsigned char smem_raw[];
   double *buf0 = reinterpret_cast<double *>(smem_raw);
   SOp *sops = reinterpret_cast<SOp *>(smem_raw + (size_t)2*TB*8);
   uint64_t *full = reinterpret_cast<uint64_t *>(sops + MAX_OPS);   // [2] tile landed
   uint64_t *done = full + 2;                                        // [2] tile updated in place

   const int tid = threadIdx.x;
   const int a = A.order[blockIdx.x%A.num_active];
   const int grp = blockIdx.x/A.num_active;
   const int v0 = A.var_start + grp*A.vpc;
   const int nv = min(A.vpc, A.var_end - v0);
   const int slot = A.slots[a];
   const long long slot_off = (long long)slot*A.tile_stride;

   const int ob = A.begin[a];
   const int nops = A.begin[a + 1] - ob;
   if (tid == 0) {
      mbar_init(&full[0], 1);
      mbar_init(&full[1], 1);
      mbar_init(&done[0], CT/32);
      mbar_init(&done[1], CT/32);
      fence_barrier_init();
   }
   stage_ops(A.ops + ob, nops, sops, tid);
   __syncthreads();

   constexpr uint32_t ROWS_BYTES = (uint32_t)N*SJ*8u;
   if (tid >= CT) {
      // ---- copy warp: every bulk copy of the CTA is issued here, so no compute
      // warp ever waits for a store to drain or spends issue slots on UBLKCP ----
      if (tid != CT) return;
      // rows 1..N of planes 1..N: one bulk copy per plane
      constexpr uint32_t ZF_BYTES = (uint32_t)N*N*8u;
      long long zs[2] = { -1, -1 };
      if (ELIDE) { zs[0] = A.zsrc[2*a]; zs[1] = A.zsrc[2*a + 1]; }
      auto load_tile = [&](int t) {
         const int b = t & 1;
         const double *tin = A.pool_in + (long long)(v0 + t)*A.var_stride + slot_off;
         double *dst = buf0 + (size_t)b*TB;
         mbar_arrive_expect_tx(&full[b], ROWS_BYTES*N + (zs[0] >= 0 ? ZF_BYTES : 0u) +
                                            (zs[1] >= 0 ? ZF_BYTES : 0u));
#pragma unroll 4
         for (int p = 1; p <= N; p++)
            bulk_g2s(dst + p*PL + SJ, tin + p*PL + SJ, ROWS_BYTES, &full[b]);
         // Z halo faces that are a neighbour's exported plane: straight into the staging area
         const double *zin = A.zf_in + (long long)(v0 + t)*A.zf_var_stride;
         if (zs[0] >= 0) bulk_g2s(dst + TILE, zin + zs[0], ZF_BYTES, &full[b]);
         if (zs[1] >= 0) bulk_g2s(dst + TILE + N*N, zin + zs[1], ZF_BYTES, &full[b]);
      };
      load_tile(0);
      if (nv > 1) load_tile(1);
      for (int t = 0; t < nv; t++) {
         double *sm = buf0 + (size_t)(t & 1)*TB;
         mbar_wait(&done[t & 1], (uint32_t)((t >> 1) & 1));
         double *tout = A.pool_out + (long long)(v0 + t)*A.var_stride + slot_off;
         if (ELIDE) {
#pragma unroll 4
            for (int p = 1; p <= N; p++)
               bulk_s2g(tout + p*PL + SJ, sm + p*PL + SJ, ROWS_BYTES);
            bulk_s2g(A.zf_out + (long long)(v0 + t)*A.zf_var_stride + (long long)slot*A.zf_slot,
                     sm + TILE + S::ZST, 2u*ZF_BYTES);
         } else
            bulk_s2g(tout, sm, (uint32_t)TILE*8u);
         bulk_commit();
         if (t + 2 < nv) {
            bulk_wait_read0();      // the store has finished reading the buffer
            load_tile(t + 2);
         }
      }
      bulk_wait_read0();
      return;
   }

   // ---- compute warps ----
   // decode this thread's share of the halo once per block
   const int last = nops - 1;
   const int E = nops > 0 ? sops[last].first + sops[last].e0*sops[last].e1*sops[last].e2 : 0;
   // dinfo: destination (20 bits) | op (6 bits) | mode (3 bits) | source memory (2 bits); a cell
   // that is an 8-byte copy out of the pool -- all of them on a uniform single-rank mesh -- has
   // nothing above bit 20 but its op index, which it never needs
   int soff[Q], dinfo[Q];
#pragma unroll
   for (int q = 0; q < Q; q++) {
      const int e = tid + q*CT;
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
         int dsto = s.dst_base + aa*s.ds0 + b*s.ds1 + c*s.ds2;
         if (ELIDE) {
            // Z-face cell -> its staging slot [side][i-1][j-1]
            const int di = dsto/PL, dj = (dsto - di*PL)/SJ, dk = dsto - di*PL - dj*SJ;
            if (di >= 1 && di <= N && dj >= 1 && dj <= N)
               dsto = TILE + (dk ? N*N : 0) + (di - 1)*N + (dj - 1);
         }
         const int mode = s.mode;
         long long o;
         if (mode == FM_COPY) o = (long long)aa*s.ss0 + b*s.ss1 + c*s.ss2;
         else if (mode == FM_SUM4) o = (long long)(2*aa)*s.ss0 + (2*b)*s.ss1 + (2*c)*s.ss2;
         else o = (long long)(aa >> 1)*s.ss0 + (b >> 1)*s.ss1 + (c >> 1)*s.ss2;
         soff[q] = (int)(s.src_base + o);
         const bool plain = mode == FM_COPY && s.src_mem == BM_POOL;
         dinfo[q] = dsto | (plain ? 0 : ((lo << 20) | (mode << 26) | (s.src_mem << 29)));
      }
   }

   // halo cell q of this thread for variable v -> buffer dst.  Per cell: the plain copy, or
   // (level boundary, receive buffer) the op's transform -- only those cells pay for it
   auto gather_one = [&](int q, int v, double *dst) {
      if (dinfo[q] < 0) return;
      const double *pin = A.pool_in + (long long)v*A.var_stride;
      if (!(dinfo[q] >> 20)) {
         cp_async8(dst + dinfo[q], pin + soff[q]);
         return;
      }
      const int mode = (dinfo[q] >> 26) & 7, mem = dinfo[q] >> 29, op = (dinfo[q] >> 20) & 63;
      const double *p = pin;
      if (mem != BM_POOL)
         p = (mem == BM_BUF0 ? A.recv[0] : (mem == BM_BUF0 + 1 ? A.recv[1] : A.recv[2])) +
             (long long)(v - A.buf_var0)*sops[op].src_vs;
      p += soff[q];
      double *d = dst + (dinfo[q] & 0xfffff);
      if (mode == FM_COPY || mode == FM_REPL)
         cp_async8(d, p);
      else if (mode == FM_PROLONG)
         *d = __ldg(p)/4.0;
      else {   // FM_SUM4, left to right, slow index outer (comm.c:1626-1629)
         const int SS = sops[op].S, FF = sops[op].F;
         double x = __ldg(p) + __ldg(p + FF);
         x += __ldg(p + SS);
         x += __ldg(p + SS + FF);
         *d = x;
      }
   };
   auto gather = [&](int v, double *dst) {
#pragma unroll
      for (int q = 0; q < Q; q++) gather_one(q, v, dst);
   };
   // Without ELIDE the whole tile is stored, so the halo of the next variable can
   // only enter its buffer once that tile has landed: probe during the stencil.
   bool pre = false;
   auto prefetch_halo = [&](int t) {
      if (ELIDE) return;
      if (t + 1 < nv && !pre && mbar_test(&full[(t + 1) & 1], (uint32_t)(((t + 1) >> 1) & 1))) {
         gather(v0 + t + 1, buf0 + (size_t)((t + 1) & 1)*TB);
         pre = true;
      }
   };
   // With ELIDE the ghost planes and ghost rows of a buffer are not part of any bulk
   // copy (and the Z-face cells go to the staging area): the halo of a variable is
   // requested a whole iteration ahead, a few cells per plane step of the stencil.
   if (ELIDE) gather(v0, buf0);
   auto early_halo = [&](int t, int step, int nsteps) {
      if (!ELIDE || t + 1 >= nv) return;
      double *dst = buf0 + (size_t)((t + 1) & 1)*TB;
#pragma unroll
      for (int q = 0; q < Q; q++)
         if (q*nsteps/Q == step) gather_one(q, v0 + t + 1, dst);
   };

   // ---- thread -> cells ----
   // 27-point: chunk of CH planes (warp-uniform) x 2x2 patch
   const int pg = tid%S::GROUPS, pch = tid/S::GROUPS;
   const int i0 = pch*S::CH;                       // ring planes i0 .. i0+CH+1
   const bool pact = pch < S::NCH && i0 < N;
   const int pjp = pg/(N/2), pkp = pg%(N/2);
   const int poff = i0*PL + (2*pjp)*SJ + 2*pkp;
   // 7-point: columns
   int coff[S::CPT];
   bool live[S::CPT];
#pragma unroll
   for (int q = 0; q < S::CPT; q++) {
      const int c = tid + q*CT;
      live[q] = c < N*N;
      const int cc = live[q] ? c : 0;
      coff[q] = (cc/N + 1)*SJ + cc%N + 1;
   }

   for (int t = 0; t < nv; t++) {
      const int v = v0 + t;
      double *sm = buf0 + (size_t)(t & 1)*TB;
      double cs = 0.0;
      if (ELIDE) {
         mbar_wait(&full[t & 1], (uint32_t)((t >> 1) & 1));
         cp_async_wait_all();
         named_bar_sync(1, CT);       // every thread's staged Z cells are visible
#pragma unroll
         for (int z = 0; z < S::ZPT; z++) {
            const int sl = tid + z*CT;
            if (sl < S::ZST) {
               const int side = sl/(N*N), rem = sl - side*(N*N);
               sm[(rem/N + 1)*PL + (rem%N + 1)*SJ + side*(N + 1)] = sm[TILE + sl];
            }
         }
      } else {
         if (!pre) {
            mbar_wait(&full[t & 1], (uint32_t)((t >> 1) & 1));
            gather(v, sm);
         }
         pre = false;
         cp_async_wait_all();
      }
      named_bar_sync(1, CT);

      if (ST == 7) {
         double r[S::CPT][N];
#pragma unroll
         for (int q = 0; q < S::CPT; q++) {
            const double *c = sm + coff[q];
            double prev = c[0], cur = c[PL];
#pragma unroll
            for (int i = 1; i <= N; i++) {
               const double *pc = c + i*PL;
               const double e = pc[PL];
               double s = prev + pc[-SJ];             // W + S
               s += pc[-1];                           // + D
               s += cur;                              // + C
               s += pc[1];                            // + U
               s += pc[SJ];                           // + N
               s += e;                                // + E
               r[q][i - 1] = s;
               prev = cur;
               cur = e;
               if (q == 0) early_halo(t, i - 1, N);
            }
            if (q == S::CPT/2) prefetch_halo(t);
         }
         named_bar_sync(1, CT);       // the old tile has been read by everyone
#pragma unroll
         for (int q = 0; q < S::CPT; q++) {
            if (!live[q]) continue;
            double *c = sm + coff[q];
            bool ok = true;
#pragma unroll
            for (int i = 0; i < N; i++) ok = ok && div_in_range(r[q][i]);
            if (ok) {
#pragma unroll
               for (int i = 0; i < N; i++) r[q][i] = div_fast<7>(r[q][i]);
            } else {
#pragma unroll
               for (int i = 0; i < N; i++) r[q][i] = r[q][i]/7.0;
            }
#pragma unroll
            for (int i = 0; i < N; i++) {
               c[(i + 1)*PL] = r[q][i];
               cs += r[q][i];
            }
            if (ELIDE) {
               // columns k=1 and k=N are the Z-face exports of this tile
               const int cc = tid + q*CT, cj = cc/N, ck = cc%N;
               if (ck == 0 || ck == N - 1) {
                  double *z = sm + TILE + S::ZST + (ck ? N*N : 0) + cj;
#pragma unroll
                  for (int i = 0; i < N; i++) z[i*N] = r[q][i];
               }
            }
         }
      } else {
         double r[S::CH][4];
         if (pact) {
            const double *pp = sm + poff;
            double sb[4], smid[4], sf[4];
            patch_sums4<SJ>(pp, sb);
            early_halo(t, 0, S::CH + 2);
            patch_sums4<SJ>(pp + PL, smid);
            early_halo(t, 1, S::CH + 2);
#pragma unroll
            for (int s = 0; s < S::CH; s++) {
               if (i0 + s < N) {
                  patch_sums4<SJ>(pp + (s + 2)*PL, sf);
#pragma unroll
                  for (int u = 0; u < 4; u++) {
                     r[s][u] = (sb[u] + smid[u]) + sf[u];
                     sb[u] = smid[u];
                     smid[u] = sf[u];
                  }
               }
               early_halo(t, s + 2, S::CH + 2);
               if (s == S::CH/2) prefetch_halo(t);
            }
         } else {
            if (ELIDE && t + 1 < nv) gather(v + 1, buf0 + (size_t)((t + 1) & 1)*TB);
            prefetch_halo(t);
         }
         named_bar_sync(1, CT);       // the old tile has been read by everyone
         if (pact) {
            // the two cells of a row pair are stored in an order that alternates with
            // the patch row: a half-warp then covers all 16 8-byte banks
            const int sw = pjp & 1;
            double *o = sm + poff + PL + SJ + 1;
#pragma unroll
            for (int s = 0; s < S::CH; s++) {
               if (i0 + s < N) {
                  const bool ok = div_in_range(r[s][0]) && div_in_range(r[s][1]) &&
                                  div_in_range(r[s][2]) && div_in_range(r[s][3]);
                  double w[4];
                  if (ok) {
#pragma unroll
                     for (int u = 0; u < 4; u++) w[u] = div_fast<27>(r[s][u]);
                  } else {
#pragma unroll
                     for (int u = 0; u < 4; u++) w[u] = r[s][u]/27.0;
                  }
                  double *q0 = o + s*PL + sw, *q1 = o + s*PL + (1 - sw);
                  cs += (w[0] + w[1]) + (w[2] + w[3]);
                  q0[0] = sw ? w[1] : w[0];
                  q1[0] = sw ? w[0] : w[1];
                  q0[SJ] = sw ? w[3] : w[2];
                  q1[SJ] = sw ? w[2] : w[3];
                  if (ELIDE && (pkp == 0 || pkp == N/2 - 1)) {
                     // columns k=1 and k=N are the Z-face exports of this tile
                     double *z = sm + TILE + S::ZST + (pkp ? N*N : 0) + (i0 + s)*N + 2*pjp;
                     z[0] = pkp ? w[1] : w[0];
                     z[1] = pkp ? w[3] : w[2];
                  }
               }
            }
         }
      }
      prefetch_halo(t);
      if (A.cspart) {
         // this warp's share of check_sum(v) over the tile it has just updated
         cs = cs_warp_sum(cs);
         if ((tid & 31) == 0)
            A.cspart[(long long)v*A.cs_var_stride + (long long)a*CS_WARPS + (tid >> 5)] = cs;
      }
      // hand the updated tile to the copy warp
      fence_proxy_async();
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&done[t & 1]);
   }
}

template <int N>
constexpr bool shape_ok()
{
   using S = Shape<N>;
   return (N%2) == 0 && S::CTAS >= 2 && S::TILE < (1 << 20) && S::NCH >= 1 &&
          S::GROUPS*S::NCH <= S::CT;
}

#define MAMR_FUSED2_SIZES(X) X(8) X(10) X(12) X(16)

template <int ST, int N, bool EL>
cudaError_t set_attr2()
{
   return cudaFuncSetAttribute(fused2_kernel<ST, N, EL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Shape<N>::SMEM);
}

}  // namespace

// is there an instantiation for this geometry?
bool fused2_supported(const Geometry &g)
{
   if (g.n[0] != g.n[1] || g.n[0] != g.n[2] || g.var_stride >= (1LL << 31)) return false;
#define X(NN) if (g.n[0] == NN) return true;
   MAMR_FUSED2_SIZES(X)
#undef X
   return false;
}

bool fused2_configure(const Geometry &g, std::string &err)
{
   if (!fused2_supported(g)) return true;
   cudaError_t e = cudaSuccess;
#define X(NN)                                                          \
   if (g.n[0] == NN) {                                                 \
      static_assert(shape_ok<NN>(), "fused2: block size does not fit"); \
      if (e == cudaSuccess) e = set_attr2<7, NN, false>();             \
      if (e == cudaSuccess) e = set_attr2<7, NN, true>();              \
      if (e == cudaSuccess) e = set_attr2<27, NN, false>();            \
      if (e == cudaSuccess) e = set_attr2<27, NN, true>();             \
   }
   MAMR_FUSED2_SIZES(X)
#undef X
   if (e != cudaSuccess) {
      err = std::string("fused2: cudaFuncSetAttribute: ") + cudaGetErrorString(e);
      return false;
   }
   return true;
}

void launch_fused2(const double *pool_in, double *pool_out, const Geometry &g, const int *d_slots,
                   const int *d_order, int num_active, const BoxOp *d_ops, const int *d_begin,
                   const double *const recv[3], int var_start, int num_vars, int buf_var0,
                   int stencil, bool elide, const double *zf_in, double *zf_out,
                   const long long *d_zsrc, double *d_cspart, long long cs_var_stride, cudaStream_t s)
{
   if (num_active <= 0 || num_vars <= 0) return;
   FusedArgs A;
   A.cspart = d_cspart; A.cs_var_stride = cs_var_stride;
   A.pool_in = pool_in; A.pool_out = pool_out; A.slots = d_slots; A.order = d_order;
   A.ops = d_ops; A.begin = d_begin;
   for (int d = 0; d < 3; d++) A.recv[d] = recv ? recv[d] : nullptr;
   A.tile_stride = g.tile_stride; A.var_stride = g.var_stride;
   A.num_active = num_active; A.buf_var0 = buf_var0;
   A.nx = g.n[0]; A.ny = g.n[1]; A.nz = g.n[2];
   A.zf_in = zf_in; A.zf_out = zf_out; A.zsrc = d_zsrc;
   A.zf_slot = 2*g.n[0]*g.n[1];
   A.zf_var_stride = (long long)A.zf_slot*(g.var_stride/g.tile_stride);
   static int vpc_env = -1;
   if (vpc_env < 0) {
      const char *e = getenv("MAMR_VPC");
      vpc_env = e ? atoi(e) : 0;
   }
   A.chunk = 0;
   // Variables per CTA: 20 amortises the per-block work (op table, halo decode) best, but a small
   // mesh must still give every SM several waves of CTAs or the last wave idles most of the GPU
   // (785 blocks x 2 groups = 2.7 waves): aim at six waves, never fewer than 4 variables per CTA.
   static int sms = 0;
   if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (sms <= 0) sms = 148;
   }
   int ctas_per_sm = 2;
#define X(NN) if (g.n[0] == NN) ctas_per_sm = Shape<NN>::CTAS;
   MAMR_FUSED2_SIZES(X)
#undef X
   const long long want = 6LL*sms*ctas_per_sm;
   const int groups_wanted = (int)std::min<long long>(num_vars, (want + num_active - 1)/num_active);
   A.vpc = vpc_env > 0 ? vpc_env : std::max(4, std::min(20, num_vars/std::max(1, groups_wanted)));
   if (A.vpc > num_vars) A.vpc = num_vars;
   A.var_start = var_start;
   A.var_end = var_start + num_vars;
   const long long groups = (num_vars + A.vpc - 1)/A.vpc;
   const unsigned grid = (unsigned)((long long)num_active*groups);
#define X(NN)                                                                                 \
   if (g.n[0] == NN) {                                                                        \
      const int sm = Shape<NN>::SMEM;                                                         \
      if (stencil == 7) {                                                                     \
         if (elide) fused2_kernel<7, NN, true><<<grid, Shape<NN>::THREADS, sm, s>>>(A);            \
         else fused2_kernel<7, NN, false><<<grid, Shape<NN>::THREADS, sm, s>>>(A);                 \
      } else {                                                                                \
         if (elide) fused2_kernel<27, NN, true><<<grid, Shape<NN>::THREADS, sm, s>>>(A);           \
         else fused2_kernel<27, NN, false><<<grid, Shape<NN>::THREADS, sm, s>>>(A);                \
      }                                                                                       \
   }
   MAMR_FUSED2_SIZES(X)
#undef X
}

// ---------------------------------------------------------------------------
// Ghost regeneration: execute the halo ops of every active block from pool_in
// (the state the last fused launch read) into the ghost cells of the tiles in
// pool_out.  grid = (active blocks, variables).  Ops flagged BF_IDENT are skipped
// (the cell already holds its value in both pools).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
halo_fill_kernel(const BoxOp *__restrict__ ops, const int *__restrict__ begin,
                 const int *__restrict__ slots, const double *__restrict__ pool_in,
                 double *__restrict__ pool_out, long long tile_stride, long long var_stride,
                 const double *recv0, const double *recv1, const double *recv2, int var_start,
                 int buf_var0, int only_ident)
{
   const int a = blockIdx.x;
   const int v = var_start + blockIdx.y;
   double *tile = pool_out + (long long)v*var_stride + (long long)slots[a]*tile_stride;
   for (int o = begin[a]; o < begin[a + 1]; o++) {
      const BoxOp op = ops[o];
      const bool ident = (op.flags & BF_IDENT) != 0;
      if (ident != (only_ident != 0)) continue;
      const double *src = (op.src_mem == BM_POOL)
                             ? pool_in + (long long)v*var_stride
                             : (op.src_mem == BM_BUF0 ? recv0 : (op.src_mem == BM_BUF0 + 1 ? recv1 : recv2)) +
                                  (long long)(v - buf_var0)*op.src_vs;
      src += op.src_base;
      double *dst = tile + op.dst_base;
      const int n = op.ext[0]*op.ext[1]*op.ext[2];
      for (int e = threadIdx.x; e < n; e += blockDim.x) {
         int r = e;
         const int c = r%op.ext[2]; r /= op.ext[2];
         const int b = r%op.ext[1];
         const int aa = r/op.ext[1];
         double x;
         if (op.mode == FM_COPY)
            x = src[(long long)aa*op.src_str[0] + b*op.src_str[1] + c*op.src_str[2]];
         else if (op.mode == FM_PROLONG || op.mode == FM_REPL) {
            x = src[(long long)(aa >> 1)*op.src_str[0] + (b >> 1)*op.src_str[1] +
                    (c >> 1)*op.src_str[2]];
            if (op.mode == FM_PROLONG) x = x/4.0;
         } else {
            const double *p = src + (long long)(2*aa)*op.src_str[0] + (2*b)*op.src_str[1] +
                              (2*c)*op.src_str[2];
            x = p[0] + p[op.F];
            x += p[op.S];
            x += p[op.S + op.F];
         }
         dst[(long long)aa*op.dst_str[0] + b*op.dst_str[1] + c*op.dst_str[2]] = x;
      }
   }
}

// ---------------------------------------------------------------------------
// Z-face exports of tiles that were not written by an eliding launch (upload,
// refinement, the generic kernels): zf[var][slot][side][i-1][j-1] = tile(i, j, 1 | nz)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
zface_extract_kernel(const double *__restrict__ pool, double *__restrict__ zf,
                     const int *__restrict__ slots, long long tile_stride, long long var_stride,
                     long long zf_var_stride, int nx, int ny, int nz, int var_start)
{
   const int slot = slots[blockIdx.x];
   const int v = var_start + blockIdx.y;
   const double *tile = pool + (long long)v*var_stride + (long long)slot*tile_stride;
   double *out = zf + (long long)v*zf_var_stride + (long long)slot*(2*nx*ny);
   const int sj = nz + 2, pl = (ny + 2)*sj;
   for (int e = threadIdx.x; e < 2*nx*ny; e += blockDim.x) {
      const int side = e/(nx*ny), r = e - side*nx*ny;
      const int i = r/ny + 1, j = r%ny + 1;
      out[e] = tile[(long long)i*pl + j*sj + (side ? nz : 1)];
   }
}

void launch_zface_extract(const double *pool, double *zf, const Geometry &g, const int *d_slots,
                          int num_active, int var_start, int num_vars, cudaStream_t s)
{
   if (num_active <= 0 || num_vars <= 0) return;
   dim3 grid((unsigned)num_active, (unsigned)num_vars);
   const long long zvs = (long long)2*g.n[0]*g.n[1]*(g.var_stride/g.tile_stride);
   zface_extract_kernel<<<grid, 256, 0, s>>>(pool, zf, d_slots, g.tile_stride, g.var_stride, zvs,
                                             g.n[0], g.n[1], g.n[2], var_start);
}

void launch_halo_fill(const BoxOp *d_ops, const int *d_begin, const int *d_slots, int num_active,
                      const double *pool_in, double *pool_out, const Geometry &g,
                      const double *const recv[3], int var_start, int num_vars, int buf_var0,
                      bool only_ident, cudaStream_t s)
{
   if (num_active <= 0 || num_vars <= 0) return;
   dim3 grid((unsigned)num_active, (unsigned)num_vars);
   halo_fill_kernel<<<grid, 256, 0, s>>>(d_ops, d_begin, d_slots, pool_in, pool_out, g.tile_stride,
                                         g.var_stride, recv[0], recv[1], recv[2], var_start,
                                         buf_var0, only_ident ? 1 : 0);
}

}  // namespace mamr
