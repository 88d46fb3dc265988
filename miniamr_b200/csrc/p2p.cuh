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
