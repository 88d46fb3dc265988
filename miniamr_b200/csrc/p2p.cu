// Peer-memory transport of the multi-GPU path: ghost-face messages and the check_sum
// all-reduce travel as plain stores into the RECEIVER's memory over NVLink (one process
// per GPU, windows mapped with CUDA IPC; ranks that live in one process -- the loopback
// tests -- use each other's pointers directly), ordered by flags with system-scope
// release / acquire.  Replaces MPI_Isend/Irecv/Waitany of comm.c:71-84,120-151 and
// MPI_Allreduce of check_sum.c:57 without a library collective in between:
//
//   credit   the receiver tells every partner "my receive buffer of (set, dir) is free
//            for exchange number seq" (its consumers of the previous contents have run:
//            stream order)
//   push     one CTA column per partner waits for that credit, copies its slice of the
//            packed send buffer into the partner's receive region and the last CTA of
//            the column raises the partner's arrival flag
//   wait     one thread per partner spins on the arrival flags; whatever is queued
//            behind it on the stream (a later phase's pack, the boundary blocks' stage
//            kernel) sees the messages
//
// A waiting thread gives up after P2P_TIMEOUT_NS and records the fact in the window
// header (mamr_sync / check_sum report it): a lost peer is an error, never a hung GPU.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "p2p.cuh"

namespace mamr {

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
   unsigned long long v;
   asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
   asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_ns()
{
   unsigned long long t;
   asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
   return t;
}

__device__ unsigned long long d_p2p_timeout_ns = P2P_TIMEOUT_NS;

// spin until *p >= want; false (and the error word set) after the timeout
__device__ bool spin_ge(const unsigned long long *p, unsigned long long want, P2PHeader *mine,
                        unsigned long long code)
{
   if (ld_acquire_sys(p) >= want) return true;
   const unsigned long long t0 = global_ns();
   unsigned ns = 32;
   for (;;) {
      if (ld_acquire_sys(p) >= want) return true;
      if (global_ns() - t0 > d_p2p_timeout_ns) {
         atomicCAS(&mine->error, 0ULL, code);
         return false;
      }
      __nanosleep(ns);
      if (ns < 1024) ns *= 2;
   }
}

__global__ void __launch_bounds__(128)
p2p_credit_kernel(const P2PTarget *targets, int n, char *const *peer, int me, int set,
                  unsigned long long seq)
{
   const int t = blockIdx.x*blockDim.x + threadIdx.x;
   if (t >= n) return;
   P2PHeader *H = reinterpret_cast<P2PHeader *>(peer[targets[t].rank]);
   st_release_sys(&H->credit[set][targets[t].dir][me], seq);
}

// grid = (chunks, partners of this direction).  128 threads and at most 40 registers: a CTA must
// fit beside the two resident CTAs of the interior blocks' stage kernel (which leave ~10 K
// registers of an SM free), or the exchange only advances as those retire.
__global__ void __launch_bounds__(128, 12)
p2p_push_kernel(const P2PTarget *parts, const double *send, char *const *peer, char *mine_raw,
                size_t data_off, unsigned *done, int me, int set, int dir, unsigned long long seq,
                unsigned long long epoch)
{
   const P2PTarget P = parts[blockIdx.y];
   P2PHeader *mine = reinterpret_cast<P2PHeader *>(mine_raw);
   P2PHeader *H = reinterpret_cast<P2PHeader *>(peer[P.rank]);
   __shared__ long long s_dst;
   if (threadIdx.x == 0) {
      // the receiver has released its buffer for this exchange, and its offset table
      // describes the current comm lists
      bool ok = spin_ge(&mine->credit[set][dir][P.rank], seq, mine, 0x100ULL | (unsigned)P.rank);
      ok = ok && spin_ge(&H->epoch, epoch, mine, 0x200ULL | (unsigned)P.rank);
      s_dst = ok ? (long long)ld_acquire_sys(reinterpret_cast<const unsigned long long *>(&H->rbase[set][dir][me]))
                 : -1;
   }
   __syncthreads();
   const long long dsto = s_dst;
   if (dsto >= 0) {
      // A thread keeps PUSH_U independent loads in flight before it stores: the copy runs beside
      // the interior blocks' stage kernel, which saturates HBM (load latency ~2 us), and one
      // load per thread at a time would move a 20 MB message at under 100 GB/s.
      double *__restrict__ dst = reinterpret_cast<double *>(peer[P.rank] + data_off) + dsto;
      const double *__restrict__ src = send + P.send_off;
      const long long per = (P.size + gridDim.x - 1)/gridDim.x;
      const long long lo = per*blockIdx.x, hi = min(P.size, lo + per);
      constexpr int PUSH_U = 8;
      for (long long e0 = lo + threadIdx.x; e0 < hi; e0 += (long long)blockDim.x*PUSH_U) {
         double v[PUSH_U];
#pragma unroll
         for (int u = 0; u < PUSH_U; u++) {
            const long long e = e0 + (long long)u*blockDim.x;
            v[u] = e < hi ? __ldg(src + e) : 0.0;
         }
#pragma unroll
         for (int u = 0; u < PUSH_U; u++) {
            const long long e = e0 + (long long)u*blockDim.x;
            if (e < hi) dst[e] = v[u];
         }
      }
   }
   __syncthreads();
   if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned old = atomicAdd(&done[blockIdx.y], 1u);
      if (old == gridDim.x - 1) {
         done[blockIdx.y] = 0;
         __threadfence_system();
         if (dsto >= 0) st_release_sys(&H->arrive[set][dir][me], seq);
      }
   }
}

__global__ void __launch_bounds__(64)
p2p_wait_kernel(const P2PTarget *parts, int n, char *mine_raw, int set, int dir,
                unsigned long long seq)
{
   P2PHeader *mine = reinterpret_cast<P2PHeader *>(mine_raw);
   for (int t = threadIdx.x; t < n; t += blockDim.x)
      spin_ge(&mine->arrive[set][dir][parts[t].rank], seq, mine, 0x300ULL | (unsigned)parts[t].rank);
}

// check_sum.c:57 as a one-shot all-reduce: every rank stores its `num` partial sums into
// every rank's window, raises a flag there, waits for everybody's flag in its own window
// and adds the contributions in rank order (the same order, hence the same bits, on
// every rank).  Two value buffers alternate: a rank can be at most one all-reduce ahead
// of the slowest reader (it needs everybody's contribution to finish its own).
__global__ void __launch_bounds__(256)
p2p_allreduce_kernel(double *sums, int num, char *const *peer, char *mine_raw, int me, int nranks,
                     int max_vars, unsigned long long seq)
{
   P2PHeader *mine = reinterpret_cast<P2PHeader *>(mine_raw);
   const int buf = (int)(seq & 1);
   for (int r = 0; r < nranks; r++) {
      double *dst = reinterpret_cast<double *>(peer[r] + P2P_CS_OFFSET) +
                    ((size_t)buf*P2P_MAX_RANKS + me)*max_vars;
      for (int v = threadIdx.x; v < num; v += blockDim.x) dst[v] = sums[v];
   }
   __syncthreads();
   if (threadIdx.x == 0) __threadfence_system();
   __syncthreads();
   for (int r = threadIdx.x; r < nranks; r += blockDim.x) {
      P2PHeader *H = reinterpret_cast<P2PHeader *>(peer[r]);
      st_release_sys(&H->cs_flag[buf][me], seq);
   }
   bool ok = true;
   for (int r = threadIdx.x; r < nranks; r += blockDim.x)
      ok = spin_ge(&mine->cs_flag[buf][r], seq, mine, 0x400ULL | (unsigned)r) && ok;
   __syncthreads();
   const volatile double *in = reinterpret_cast<const volatile double *>(mine_raw + P2P_CS_OFFSET) +
                               (size_t)buf*P2P_MAX_RANKS*max_vars;
   for (int v = threadIdx.x; v < num; v += blockDim.x) {
      double s = 0.0;
      for (int r = 0; r < nranks; r++) s += in[(size_t)r*max_vars + v];
      sums[v] = s;
   }
}

// ---- block migration: pull -------------------------------------------------------
// The sender has packed its outgoing blocks into its window and written, per staged block,
// the destination rank and the ordinal among its sends to that rank.  The receiver's k-th
// staged receive from rank S is S's k-th staged send to it (rcb.c:207-337 moves blocks of one
// pair in order).
__global__ void __launch_bounds__(256)
p2p_mv_resolve_kernel(const P2PMove *recvs, char *const *peer, char *mine_raw, size_t mv_off, int mv_cap,
                      int me, unsigned long long seq, int *k_out)
{
   const P2PMove R = recvs[blockIdx.x];
   P2PHeader *mine = reinterpret_cast<P2PHeader *>(mine_raw);
   P2PHeader *H = reinterpret_cast<P2PHeader *>(peer[R.src]);
   __shared__ int s_ok;
   if (threadIdx.x == 0) {
      k_out[blockIdx.x] = -1;
      s_ok = spin_ge(&H->mv_ready, seq, mine, 0x500ULL | (unsigned)R.src) ? 1 : 0;
   }
   __syncthreads();
   if (!s_ok) return;
   const int count = (int)ld_acquire_sys(&H->mv_count);
   const volatile int *dest = reinterpret_cast<const volatile int *>(peer[R.src] + mv_off);
   const volatile int *ord = dest + mv_cap;
   for (int k = threadIdx.x; k < count; k += blockDim.x)
      if (dest[k] == me && ord[k] == R.ordinal) k_out[blockIdx.x] = k;
}

// grid = (receives, variables)
__global__ void __launch_bounds__(256)
p2p_mv_unpack_kernel(const P2PMove *recvs, const int *k_in, char *const *peer, size_t pay_off,
                     double *pool0, double *pool1, const unsigned char *cur, int nx, int ny, int nz,
                     long long tile_stride, long long var_stride, int num_vars)
{
   const P2PMove R = recvs[blockIdx.x];
   const int k = k_in[blockIdx.x];
   if (k < 0) return;
   const int var = blockIdx.y;
   const int sj = nz + 2, si = (ny + 2)*sj, cells = nx*ny*nz;
   const double *pl = reinterpret_cast<const double *>(peer[R.src] + pay_off) +
                      ((size_t)k*num_vars + var)*cells;
   double *tile = (cur[var] ? pool1 : pool0) + (long long)var*var_stride + (long long)R.slot*tile_stride;
   for (int c = threadIdx.x; c < cells; c += blockDim.x) {
      const int i = c/(ny*nz), r = c - i*ny*nz, j = r/nz, kk = r - j*nz;
      tile[(size_t)(i + 1)*si + (j + 1)*sj + (kk + 1)] = pl[c];
   }
}

__global__ void p2p_mv_done_kernel(const int *ranks, int n, char *const *peer, int me, unsigned long long seq)
{
   for (int t = threadIdx.x; t < n; t += blockDim.x)
      st_release_sys(&reinterpret_cast<P2PHeader *>(peer[ranks[t]])->mv_done[me], seq);
}

__global__ void p2p_mv_wait_kernel(const int *ranks, int n, char *mine_raw, unsigned long long seq)
{
   P2PHeader *mine = reinterpret_cast<P2PHeader *>(mine_raw);
   for (int t = threadIdx.x; t < n; t += blockDim.x)
      spin_ge(&mine->mv_done[ranks[t]], seq, mine, 0x600ULL | (unsigned)ranks[t]);
}

}  // namespace

// MAMR_P2P_TIMEOUT_S: how long a kernel waits for a peer before it gives up (default 20 s)
double p2p_set_timeout_from_env()
{
   unsigned long long ns = P2P_TIMEOUT_NS;
   if (const char *e = getenv("MAMR_P2P_TIMEOUT_S")) {
      const double sec = atof(e);
      if (sec > 0.0) ns = (unsigned long long)(sec*1e9);
   }
   cudaMemcpyToSymbol(d_p2p_timeout_ns, &ns, sizeof ns);
   return (double)ns*1e-9;
}

void launch_p2p_mv_resolve(const P2PMove *d_recvs, int n, char *const *d_peer, char *mine, size_t mv_off,
                           int mv_cap, int me, unsigned long long seq, int *d_k, cudaStream_t s)
{
   if (n <= 0) return;
   p2p_mv_resolve_kernel<<<n, 256, 0, s>>>(d_recvs, d_peer, mine, mv_off, mv_cap, me, seq, d_k);
}

void launch_p2p_mv_unpack(const P2PMove *d_recvs, int n, const int *d_k, char *const *d_peer, size_t mv_off,
                          int mv_cap, double *pool0, double *pool1, const unsigned char *d_cur, int nx, int ny,
                          int nz, long long tile_stride, long long var_stride, int num_vars, cudaStream_t s)
{
   if (n <= 0) return;
   dim3 grid((unsigned)n, (unsigned)num_vars);
   p2p_mv_unpack_kernel<<<grid, 256, 0, s>>>(d_recvs, d_k, d_peer, mv_off + p2p_mv_payload_offset(mv_cap), pool0,
                                             pool1, d_cur, nx, ny, nz, tile_stride, var_stride, num_vars);
}

void launch_p2p_mv_done(const int *d_ranks, int n, char *const *d_peer, int me, unsigned long long seq,
                        cudaStream_t s)
{
   if (n <= 0) return;
   p2p_mv_done_kernel<<<1, 64, 0, s>>>(d_ranks, n, d_peer, me, seq);
}

void launch_p2p_mv_wait(const int *d_ranks, int n, char *mine, unsigned long long seq, cudaStream_t s)
{
   if (n <= 0) return;
   p2p_mv_wait_kernel<<<1, 64, 0, s>>>(d_ranks, n, mine, seq);
}

void launch_p2p_credit(const P2PTarget *d_targets, int n, char *const *d_peer, int me, int set,
                       unsigned long long seq, cudaStream_t s)
{
   if (n <= 0) return;
   p2p_credit_kernel<<<(n + 127)/128, 128, 0, s>>>(d_targets, n, d_peer, me, set, seq);
}

void launch_p2p_push(const P2PTarget *d_parts, int n, long long max_size, const double *send,
                     char *const *d_peer, char *mine, size_t data_off, unsigned *d_done, int me, int set,
                     int dir, unsigned long long seq, unsigned long long epoch, cudaStream_t s)
{
   if (n <= 0) return;
   // 16 KB per CTA, at most 256 CTAs per partner (x 128 threads x 8 loads in flight)
   int chunks = (int)std::min<long long>(256, std::max<long long>(1, (max_size + 2047)/2048));
   dim3 grid((unsigned)chunks, (unsigned)n);
   p2p_push_kernel<<<grid, 128, 0, s>>>(d_parts, send, d_peer, mine, data_off, d_done, me, set, dir, seq, epoch);
}

void launch_p2p_wait(const P2PTarget *d_parts, int n, char *mine, int set, int dir,
                     unsigned long long seq, cudaStream_t s)
{
   if (n <= 0) return;
   p2p_wait_kernel<<<1, 64, 0, s>>>(d_parts, n, mine, set, dir, seq);
}

void launch_p2p_allreduce(double *d_sums, int num, char *const *d_peer, char *mine, int me, int nranks,
                          int max_vars, unsigned long long seq, cudaStream_t s)
{
   p2p_allreduce_kernel<<<1, 256, 0, s>>>(d_sums, num, d_peer, mine, me, nranks, max_vars, seq);
}

}  // namespace mamr
