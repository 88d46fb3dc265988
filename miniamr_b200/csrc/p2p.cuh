// Peer-memory transport (p2p.cu): layout of a rank's window and the kernel launchers.
#pragma once
#include <cuda_runtime.h>

namespace mamr {

constexpr int P2P_MAX_RANKS = 64;      // one node: ranks whose windows can be mapped
constexpr int P2P_MAX_SETS = 16;       // = mamr_ctx::MAX_SETS (receive-buffer sets)
constexpr unsigned long long P2P_TIMEOUT_NS = 20ULL*1000*1000*1000;

// At the start of every rank's window.  Flags only ever grow.
struct P2PHeader {
   unsigned long long epoch;       // rbase[] describes the comm lists of this epoch (written by the owner)
   unsigned long long error;       // first wait of the owner that timed out (code | peer rank)
   // [set][dir][sender]: number of the last exchange whose message from `sender` has landed
   unsigned long long arrive[P2P_MAX_SETS][3][P2P_MAX_RANKS];
   // [set][dir][receiver]: `receiver` has released its receive buffer for this exchange number
   unsigned long long credit[P2P_MAX_SETS][3][P2P_MAX_RANKS];
   // [set][dir][sender]: where the sender's message starts, in doubles from the data area
   long long rbase[P2P_MAX_SETS][3][P2P_MAX_RANKS];
   // check_sum all-reduce: [buffer][contributor]
   unsigned long long cs_flag[2][P2P_MAX_RANKS];
   // block migration (pull): the owner's staged payloads and their table are complete for
   // move round mv_ready (mv_count entries); mv_done[puller]: `puller` has fetched its share
   unsigned long long mv_ready, mv_count;
   unsigned long long mv_done[P2P_MAX_RANKS];
};

// byte offsets inside a window: header | check_sum values [2][P2P_MAX_RANKS][max_vars] | data
constexpr size_t P2P_CS_OFFSET = (sizeof(P2PHeader) + 255)/256*256;
inline size_t p2p_data_offset(int max_vars)
{
   return (P2P_CS_OFFSET + (size_t)2*P2P_MAX_RANKS*max_vars*sizeof(double) + 255)/256*256;
}

// a (rank, direction) pair: a credit target, or one partner of a direction's push
struct P2PTarget {
   int rank, dir;
   long long send_off, size;       // push: the message inside my send buffer (doubles)
};

// one staged receive of a migrated block: the `ordinal`-th block rank `src` sends to me
struct P2PMove {
   int slot, src, ordinal, pad;
};

// Migration area of a window (byte offset mv_off): int dest[mv_cap], int ordinal[mv_cap], then
// the staged payloads (pack.c:66-70 layout) from p2p_mv_payload_offset() on.
inline size_t p2p_mv_payload_offset(int mv_cap) { return ((size_t)2*mv_cap*sizeof(int) + 255)/256*256; }

// find the staging index of every staged receive in its sender's table (waits for the sender)
void launch_p2p_mv_resolve(const P2PMove *d_recvs, int n, char *const *d_peer, char *mine, size_t mv_off,
                           int mv_cap, int me, unsigned long long seq, int *d_k, cudaStream_t s);
// fetch the payloads from the senders' windows straight into the tiles' interiors
void launch_p2p_mv_unpack(const P2PMove *d_recvs, int n, const int *d_k, char *const *d_peer, size_t mv_off,
                          int mv_cap, double *pool0, double *pool1, const unsigned char *d_cur, int nx, int ny,
                          int nz, long long tile_stride, long long var_stride, int num_vars, cudaStream_t s);
// tell the senders (ranks[0..n)) that their payloads have been fetched / wait for my pullers
void launch_p2p_mv_done(const int *d_ranks, int n, char *const *d_peer, int me, unsigned long long seq,
                        cudaStream_t s);
void launch_p2p_mv_wait(const int *d_ranks, int n, char *mine, unsigned long long seq, cudaStream_t s);

double p2p_set_timeout_from_env();      // seconds in force

void launch_p2p_credit(const P2PTarget *d_targets, int n, char *const *d_peer, int me, int set,
                       unsigned long long seq, cudaStream_t s);
void launch_p2p_push(const P2PTarget *d_parts, int n, long long max_size, const double *send,
                     char *const *d_peer, char *mine, size_t data_off, unsigned *d_done, int me, int set,
                     int dir, unsigned long long seq, unsigned long long epoch, cudaStream_t s);
void launch_p2p_wait(const P2PTarget *d_parts, int n, char *mine, int set, int dir,
                     unsigned long long seq, cudaStream_t s);
void launch_p2p_allreduce(double *d_sums, int num, char *const *d_peer, char *mine, int me, int nranks,
                          int max_vars, unsigned long long seq, cudaStream_t s);

}  // namespace mamr
