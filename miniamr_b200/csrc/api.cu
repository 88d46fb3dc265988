// C ABI of the B200-native miniAMR stage hot path (include/miniamr_b200.h).
// Host-side orchestration only: device pool, topology -> ghost descriptors,
// direction phases, NCCL send/recv of ghost faces and migrated blocks, deferred
// per-variable stencil launches.  All arithmetic is in the kernels
// (stencil.cu, ghost.cu, reduce_refine.cu).  There is no CPU fallback.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "p2p.cuh"
#include "stencil0.cuh"

using namespace mamr;

// ---------------------------------------------------------------------------
// NCCL, resolved at run time so that single-GPU users never need it.
// ---------------------------------------------------------------------------
namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_DOUBLE = 8, NCCL_SUM = 0 };
struct NcclApi {
   void *handle = nullptr;
   int (*GetUniqueId)(ncclUniqueId *) = nullptr;
   int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
   int (*CommDestroy)(ncclComm_t) = nullptr;
   int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
   int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
   int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
   int (*GroupStart)() = nullptr;
   int (*GroupEnd)() = nullptr;
   const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
   char buf[1024];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof buf, fmt, ap);
   va_end(ap);
   g_err = buf;
   return code;
}

bool load_nccl()
{
   if (g_nccl.handle) return true;
   const char *names[] = { getenv("MAMR_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
   void *h = nullptr;
   for (const char *n : names)
      if (n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
   if (!h) return false;
#define SYM(field, name)                                              \
   *(void **)(&g_nccl.field) = dlsym(h, name);                        \
   if (!g_nccl.field) return false;
   SYM(GetUniqueId, "ncclGetUniqueId")
   SYM(CommInitRank, "ncclCommInitRank")
   SYM(CommDestroy, "ncclCommDestroy")
   SYM(Send, "ncclSend")
   SYM(Recv, "ncclRecv")
   SYM(AllReduce, "ncclAllReduce")
   SYM(GroupStart, "ncclGroupStart")
   SYM(GroupEnd, "ncclGroupEnd")
   SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
   g_nccl.handle = h;
   return true;
}
}  // namespace

#define CU(call)                                                                          \
   do {                                                                                   \
      cudaError_t e_ = (call);                                                            \
      if (e_ != cudaSuccess)                                                              \
         return fail(MAMR_ECUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_),      \
                     __FILE__, __LINE__, cudaGetErrorString(e_));                         \
   } while (0)
#define NC(call)                                                                          \
   do {                                                                                   \
      int r_ = (call);                                                                    \
      if (r_ != 0)                                                                        \
         return fail(MAMR_ENCCL, "NCCL error at %s:%d: %s", __FILE__, __LINE__,           \
                     g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?");            \
   } while (0)
#define CK(call)                     \
   do {                              \
      int r_ = (call);               \
      if (r_ != MAMR_OK) return r_;  \
   } while (0)

struct EventPair { cudaEvent_t a, b; int cls, detail; };

struct mamr_ctx {
   mamr_params p;
   Geometry g;
   int comm_vars;
   // two pools: the fused stage kernel reads a variable's current pool and
   // writes the other one; cur[v] says which pool holds variable v
   double *pool[2] = {nullptr, nullptr};
   std::vector<unsigned char> cur;
   size_t pool_bytes = 0;       // of one pool
   // comm() deferred into the fused kernel: phase-order index (0..5) or -1, and
   // the first variable of that comm() call (the receive buffers' variable 0)
   std::vector<signed char> pc_ord;
   std::vector<int> pc_start;
   bool fused_geom = false;     // the tile fits the fused kernel
   bool use_fused = true;       // MAMR_NO_FUSED=1 forces the split path
   bool fused2_geom = false;    // fused2.cu has an instantiation for this block size
   bool use_fused2 = true;      // MAMR_NO_FUSED2=1: always the generic fused kernel
   bool use_elide = true;       // MAMR_NO_ELIDE=1: fused2 always stores whole tiles
   // Ghost elision.  After an eliding fused2 launch the i-ghost planes and j-ghost
   // rows of variable v's current tiles are stale; they equal the halo plan
   // `stale_ord[v]` applied to the OTHER pool (the state that launch read) and the
   // receive buffers of the comm() that started at variable stale_start[v].
   std::vector<char> stale;
   std::vector<signed char> stale_ord;
   std::vector<int> stale_start;
   // both pools agree on the ghost regions no phase ever writes (BF_IDENT)
   std::vector<char> shell_synced;
   // Z-face exports, one pool per tile pool (fused2.cu): zf[p][var][slot][side][nx*ny];
   // zf_ok[v]: the export pool of v's current pool matches its tiles
   double *zf[2] = {nullptr, nullptr};
   size_t zf_bytes = 0;
   std::vector<char> zf_ok;
   long long *d_zsrc[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   // streamed 7-point kernel (slab7.cu): per phase order, the plain faces of every
   // block and the cell ops of the others
   bool slab_geom = false;
   bool use_slab = true;        // MAMR_NO_SLAB=1
   bool slab_ok[6] = {false, false, false, false, false, false};
   long long *d_fsrc[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   BoxOp *d_cops[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   int *d_cbegin[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   BoxOp *d_lops[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // lean op lists
   int *d_lbegin[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   bool plan_has_ident[6] = {false, false, false, false, false, false};
   HaloPlan plan[6];
   std::vector<BoxOp> pack[6][3];       // multi-GPU: send-buffer fill per phase, CSR by face
   std::vector<int> pack_fb[6][3];
   BoxOp *d_pack[6][3] = {};
   int *d_pack_fb[6][3] = {};
   bool plan_built[6] = {false, false, false, false, false, false};
   BoxOp *d_hops[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   int *d_hbegin[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   cudaStream_t stream = nullptr;
   // Overlap of the off-rank exchange with the stencil of the blocks that do not read
   // it (MAMR_NO_OVERLAP=1 turns it off): pack kernels + NCCL run on `xstream`; the
   // fused kernel is launched for the interior blocks on the main stream right away
   // and for the boundary blocks on `bstream`, which waits for the exchange (ev_xchg).
   // order_ord[ord] = interior blocks, then boundary blocks (each in processing order).
   // mamr_upload_interiors: copy stream + two staging buffers (H2D of chunk k+1 runs
   // while chunk k is scattered into the tiles)
   cudaStream_t upstream = nullptr;
   double *d_up[2] = {nullptr, nullptr};
   size_t up_cap = 0;
   cudaEvent_t ev_up_copy[2] = {nullptr, nullptr}, ev_up_fill[2] = {nullptr, nullptr};
   bool up_fill_rec[2] = {false, false};   // ev_up_fill[b] has been recorded (by this or an earlier call)
   cudaStream_t xstream = nullptr, bstream = nullptr;
   cudaEvent_t ev_data = nullptr, ev_xchg = nullptr, ev_pre = nullptr, ev_bdone = nullptr;
   bool use_overlap = true;
   bool xchg_pending = false;
   int *d_order_ord[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   int n_interior[6] = {0, 0, 0, 0, 0, 0};
   std::vector<int> h_order;

   int num_active = 0;
   std::vector<mamr_block> blocks;
   int *d_slots = nullptr;
   int *d_order = nullptr;      // fused kernel: CTA index -> active block index
   size_t slots_cap = 0;

   // ghost descriptors, per direction: [local | pack] contiguous, then unpack
   std::vector<FaceOp> ops_main[3], ops_unpack[3];
   FaceOp *d_ops = nullptr;
   size_t ops_cap = 0;
   size_t off_main[3] = {0, 0, 0}, off_unpack[3] = {0, 0, 0};
   bool ops_dirty = true;
   long long n_same[3], n_diff[3], n_bc[3];

   DirLists cl[3];
   bool have_partners = false;
   // --send_faces: message accounting per face, with the host's msg_len[dir][0..3] (init.c:77-117)
   bool send_faces = false;
   int msg_len[3][4] = {};
   // One send buffer per direction.  Receive buffers come in `nsets` sets, one per
   // group of comm_vars variables (driver.c:75-89 calls comm() once per group and
   // stage): the messages of a group stay valid until the same group's next comm(),
   // so deferred exchanges and elided ghost layers of the OTHER groups need not be
   // made real when a buffer is reused.  pc_set / stale_set: the set a variable's
   // deferred comm() / stale ghost layer reads.
   static constexpr int MAX_SETS = 16;
   int nsets = 1;
   double *d_send[3] = {nullptr, nullptr, nullptr};
   double *d_recvs[MAX_SETS][3] = {};
   size_t send_cap[3] = {0, 0, 0}, recv_cap[3] = {0, 0, 0};
   std::vector<unsigned char> pc_set, stale_set;

   double *d_partials = nullptr;
   size_t partials_cap = 0;
   // partial sums the fused kernels leave behind (fused_common.cuh: cspart), CS_WARPS per
   // tile-variable; cs_fused[v]: they describe variable v's current interiors
   double *d_cspart = nullptr;
   size_t cspart_cap = 0;
   std::vector<char> cs_fused;
   bool use_cs_fused = true;    // MAMR_NO_FUSED_CS=1
   // ... and only while somebody asks for checksums: fused launches since the last
   // check_sum(); beyond about eight stages' worth the kernels stop producing partials
   int launches_since_cs = 0;
   double *d_sums = nullptr, *h_sums = nullptr;
   std::vector<char> cs_valid, cs_local_dirty;
   std::vector<double> cs_cache;
   bool modified_since_cs = true;

   int pend_start = 0, pend_num = 0;
   // Look-ahead over a comm group.  driver.c:85-103 interleaves stencil_driver(v) and
   // check_sum(v) on checksum stages, which flushes one variable at a time.  The fused
   // kernels write the OTHER pool and leave the input untouched, so the first flush also
   // computes the variables behind it that wait under the same deferred comm(); they are
   // committed (pool flip, flags) when the host asks for their stencil, and dropped if
   // anything else touches them first.  spec[v]: pool[cur[v]^1] holds that result;
   // spec_flags[v]: bit 0 eliding launch, bit 1 check_sum partials written.
   // spec_cs[v] (valid iff spec_cs_ok[v]): check_sum of that result, folded together with an
   // earlier variable's check_sum so that the host's next check_sum(v) is a cache hit
   std::vector<char> spec, spec_flags, spec_cs_ok;
   std::vector<double> spec_cs;
   bool use_lookahead = true;   // MAMR_NO_LOOKAHEAD=1
   int pend_stage = 0;          // calc_stage of the queued stencil calls (--stencil 0 only)
   // --stencil 0 (stencil.c:147-983): coefficients of init.c:418-423, scratch tiles for the
   // work[] kinds, and how often stencil_check took its two branches (flop counters)
   bool s0_set = false;
   int s0_mat = 0;
   double s0_a1 = 0.0;
   std::vector<double> s0_a0;
   double *d_s0_a0 = nullptr, *d_s0_work = nullptr;
   int s0_work_blocks = 0;
   unsigned long long *d_s0_chk = nullptr, *h_s0_chk = nullptr;

   RefineOp *d_rops = nullptr;
   int rops_cap = 4096, rops_pos = 0;
   double *d_payload = nullptr;
   double *h_stage = nullptr;   // pinned, one block
   // staged block migration (mamr_stage_send_block / _recv_block / mamr_flush_block_moves):
   // payloads packed at stage time, exchanged later in ONE NCCL group
   struct MoveRec { int slot, peer; };
   std::vector<MoveRec> mv_send, mv_recv;
   double *d_mv_send = nullptr, *d_mv_recv = nullptr;
   size_t mv_send_cap = 0, mv_recv_cap = 0;     // in blocks

   mamr_counters cnt;

   ncclComm_t nccl = nullptr;

   // Peer-memory transport (p2p.cu): every rank owns a window -- flags, check_sum slots and
   // the receive buffers of all sets -- that its partners store into directly.
   bool p2p = false;            // windows of all ranks are mapped: ghost messages and the
                                // check_sum all-reduce go through them
   bool p2p_inproc = false;     // some peer lives in this process (loopback): see dfree()
   double p2p_timeout_s = 20.0;
   char *win = nullptr;
   size_t win_bytes = 0, win_data_off = 0, win_data_cap = 0;   // capacity in doubles
   std::vector<char *> peer_win;          // mapped windows, by rank (mine included)
   std::vector<char> peer_ipc;            // opened with cudaIpcOpenMemHandle
   char **d_peer_win = nullptr;
   unsigned long long p2p_epoch = 0;      // mamr_set_comm_lists calls
   unsigned long long xseq[MAX_SETS] = {};   // ghost exchanges per receive-buffer set
   unsigned long long cs_seq = 0;         // check_sum all-reduces
   P2PTarget *d_push[3] = {nullptr, nullptr, nullptr}, *d_credit = nullptr;
   int n_credit = 0;
   long long push_max[3] = {0, 0, 0};
   unsigned *d_push_done = nullptr;
   unsigned long long *h_p2p_err = nullptr;   // pinned
   char *arena = nullptr;                 // dalloc(): ranks sharing a process
   size_t arena_cap = 0, arena_top = 0;
   std::map<void *, size_t> arena_size;
   std::multimap<size_t, void *> arena_free;
   std::vector<void *> garbage;           // release put off until mamr_destroy (dfree)
   // migrated blocks over the windows (pull): staged payloads + their (dest, ordinal) table live
   // in the sender's window from win_mv_off on; mv_seq counts mamr_flush_block_moves calls
   size_t win_mv_off = 0;
   int mv_cap_p2p = 0;
   unsigned long long mv_seq = 0;
   P2PMove *d_mv_moves = nullptr;
   int *d_mv_k = nullptr, *d_mv_ranks = nullptr;
   unsigned char *d_cur = nullptr;
   size_t mv_moves_cap = 0;

   cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
   bool ktiming = false;
   std::vector<EventPair> kev;
   std::vector<cudaEvent_t> ev_free;
   double k_ms[3] = {0, 0, 0};
   long long k_launches[3] = {0, 0, 0};
   double kd_ms[16] = {};      // the same by kernel kind (KD_*): mamr_get_device_times
   // MAMR_TRACE=1: where a stage's time goes when an off-rank exchange overlaps the interior
   // blocks' kernel.  Events: 0 comm() reached on the main stream, 1 exchange complete, 2 interior
   // launch complete, 3 boundary launch starts, 4 boundary launch complete; means printed by
   // mamr_destroy.
   bool trace = false;
   struct TraceRec { cudaEvent_t e[5]; int have; };
   TraceRec tr_cur = {};
   std::vector<TraceRec> tr_done;
   double tr_sum[4] = {0, 0, 0, 0};
   long long tr_n = 0;
};

namespace {

enum { KC_STENCIL = 0, KC_GHOST = 1, KC_CHECKSUM = 2 };
enum { KD_FUSED = 0, KD_STENCIL, KD_SPLIT, KD_PACK, KD_XCHG, KD_UNPACK, KD_REGEN, KD_CS, KD_ALLRED, KD_NUM };

int dfree(mamr_ctx *c, void *p);
template <typename T> int dalloc(mamr_ctx *c, T **p, size_t bytes);

const int kPerm[6][3] = { {0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0} };

inline int order_index(const mamr_ctx *c, int stage)   // comm.c:51-55
{
   return c->p.permute ? ((stage%6) + 6)%6 : 0;
}

inline double *vpool(mamr_ctx *c, int v) { return c->pool[c->cur[v]]; }

// [v0, v0+n) split into maximal runs that live in the same pool and have the same
// deferred-comm state
struct Run { int start, num; };
std::vector<Run> runs_of(const mamr_ctx *c, int v0, int n, bool split_on_comm)
{
   std::vector<Run> r;
   for (int v = v0; v < v0 + n; v++) {
      if (!r.empty()) {
         const int p = r.back().start;
         if (c->cur[p] == c->cur[v] &&
             (!split_on_comm || (c->pc_ord[p] == c->pc_ord[v] && c->pc_start[p] == c->pc_start[v] &&
                                 c->spec[p] == c->spec[v]))) {
            r.back().num++;
            continue;
         }
      }
      r.push_back({ v, 1 });
   }
   return r;
}

struct KTimer {
   mamr_ctx *c;
   EventPair ep;
   bool on;
   cudaStream_t st;
   KTimer(mamr_ctx *c_, int cls, cudaStream_t st_ = nullptr, int detail = -1)
      : c(c_), on(c_->ktiming), st(st_ ? st_ : c_->stream)
   {
      ep.detail = detail;
      if (!on) return;
      for (cudaEvent_t *e : { &ep.a, &ep.b }) {
         if (!c->ev_free.empty()) {
            *e = c->ev_free.back();
            c->ev_free.pop_back();
         } else
            cudaEventCreate(e);
      }
      ep.cls = cls;
      cudaEventRecord(ep.a, st);
   }
   ~KTimer()
   {
      if (!on) return;
      cudaEventRecord(ep.b, st);
      c->kev.push_back(ep);
   }
};

void trace_mark(mamr_ctx *c, int i, cudaStream_t st)
{
   if (!c->trace) return;
   if (i == 0) {
      c->tr_cur.have = 0;
      for (int k = 0; k < 5; k++) cudaEventCreate(&c->tr_cur.e[k]);
   }
   if (!c->tr_cur.e[i]) return;
   cudaEventRecord(c->tr_cur.e[i], st);
   c->tr_cur.have |= 1 << i;
   if (i == 4) {
      if (c->tr_cur.have == 31) c->tr_done.push_back(c->tr_cur);
      c->tr_cur = {};
   }
}

void trace_drain(mamr_ctx *c)
{
   for (mamr_ctx::TraceRec &r : c->tr_done) {
      cudaEventSynchronize(r.e[4]);
      cudaEventSynchronize(r.e[1]);
      for (int k = 1; k < 5; k++) {
         float ms = 0.f;
         cudaEventElapsedTime(&ms, r.e[0], r.e[k]);
         c->tr_sum[k - 1] += ms;
      }
      c->tr_n++;
      for (int k = 0; k < 5; k++) cudaEventDestroy(r.e[k]);
   }
   c->tr_done.clear();
}

// fold the timed intervals into the totals; wait == false: only those that have finished
void drain_ktimers(mamr_ctx *c, bool wait = true)
{
   size_t keep = 0;
   for (EventPair &ep : c->kev) {
      if (!wait && cudaEventQuery(ep.b) != cudaSuccess) {
         c->kev[keep++] = ep;
         continue;
      }
      float ms = 0.f;
      cudaEventSynchronize(ep.b);
      cudaEventElapsedTime(&ms, ep.a, ep.b);
      c->k_ms[ep.cls] += ms;
      c->k_launches[ep.cls]++;
      if (ep.detail >= 0) c->kd_ms[ep.detail] += ms;
      c->ev_free.push_back(ep.a);
      c->ev_free.push_back(ep.b);
   }
   c->kev.resize(keep);
   cudaGetLastError();      // cudaErrorNotReady of a query is not an error
}

// in-face axes in buffer order (slow, fast): comm.c:266-270, 311-320, 361-370
inline void face_axes(int d, int &sa, int &fa)
{
   sa = (d == 0) ? 1 : 0;
   fa = (d == 2) ? 1 : 2;
}

// whole-face extent of in-face axis ax: widened to 0..n+1 for the axes of the
// directions exchanged earlier when edges/corners must travel
// (comm.c:1496-1527, 306-320, 356-370)
inline void whole_extent(const Geometry &g, bool wide, int d, int ax, int &lo, int &hi)
{
   if (wide && ax < d) { lo = 0; hi = g.n[ax] + 1; }
   else { lo = 1; hi = g.n[ax]; }
}

inline long long tile_base(const Geometry &g, int slot) { return (long long)slot*g.tile_stride; }

FaceOp tile_copy(const Geometry &g, int d, int dst_slot, int dst_plane, int src_slot,
                 int src_plane, int s0, int s1, int f0, int f1)
{
   int sa, fa;
   face_axes(d, sa, fa);
   FaceOp op;
   const long long inplane = (long long)s0*g.str[sa] + (long long)f0*g.str[fa];
   op.dst_base = tile_base(g, dst_slot) + (long long)dst_plane*g.str[d] + inplane;
   op.src_base = tile_base(g, src_slot) + (long long)src_plane*g.str[d] + inplane;
   op.dst_vs = op.src_vs = g.var_stride;
   op.dst_S = op.src_S = g.str[sa];
   op.dst_F = op.src_F = g.str[fa];
   op.Ns = s1 - s0 + 1;
   op.Nf = f1 - f0 + 1;
   op.mode = FM_COPY;
   op.mem = 0;
   return op;
}

// on_proc_comm, comm.c:1473-1534: lo = block on the minus side of the face
void add_same(mamr_ctx *c, int d, int lo, int hi)
{
   const Geometry &g = c->g;
   int sa, fa, s0, s1, f0, f1;
   face_axes(d, sa, fa);
   const bool wide = c->p.stencil != 7;
   whole_extent(g, wide, d, sa, s0, s1);
   whole_extent(g, wide, d, fa, f0, f1);
   c->ops_main[d].push_back(tile_copy(g, d, lo, g.n[d] + 1, hi, 1, s0, s1, f0, f1));
   c->ops_main[d].push_back(tile_copy(g, d, hi, 0, lo, g.n[d], s0, s1, f0, f1));
}

// on_proc_comm_diff, comm.c:1597-1688: cs coarse slot, fs fine slot, l = face of
// the coarse block; jq picks the half along the slow axis, iq along the fast one
void add_diff(mamr_ctx *c, int cs, int fs, int l, int iq, int jq)
{
   const Geometry &g = c->g;
   const int d = l/2;
   int sa, fa;
   face_axes(d, sa, fa);
   const int hs = g.n[sa]/2, hf = g.n[fa]/2, os = jq*hs, of = iq*hf;
   int c_ghost, c_src, f_ghost, f_src;
   if (l%2 == 0) { c_ghost = 0; c_src = 1; f_ghost = g.n[d] + 1; f_src = g.n[d]; }
   else { c_ghost = g.n[d] + 1; c_src = g.n[d]; f_ghost = 0; f_src = 1; }
   const long long S = g.str[sa], F = g.str[fa], N = g.str[d];
   FaceOp pro;   // coarse -> fine ghosts: value/4 replicated 2x2
   pro.dst_base = tile_base(g, fs) + f_ghost*N + S + F;
   pro.src_base = tile_base(g, cs) + c_src*N + (1 + os)*S + (1 + of)*F;
   pro.dst_vs = pro.src_vs = g.var_stride;
   pro.dst_S = pro.src_S = (int)S;
   pro.dst_F = pro.src_F = (int)F;
   pro.Ns = g.n[sa];
   pro.Nf = g.n[fa];
   pro.mode = FM_PROLONG;
   pro.mem = 0;
   c->ops_main[d].push_back(pro);
   FaceOp res;   // fine -> coarse ghost quarter: 4-term sum
   res.dst_base = tile_base(g, cs) + c_ghost*N + (1 + os)*S + (1 + of)*F;
   res.src_base = tile_base(g, fs) + f_src*N + S + F;
   res.dst_vs = res.src_vs = g.var_stride;
   res.dst_S = res.src_S = (int)S;
   res.dst_F = res.src_F = (int)F;
   res.Ns = hs;
   res.Nf = hf;
   res.mode = FM_SUM4;
   res.mem = 0;
   c->ops_main[d].push_back(res);
}

// apply_bc, comm.c:1911-1965
void add_bc(mamr_ctx *c, int slot, int l)
{
   const Geometry &g = c->g;
   const int d = l/2;
   int sa, fa;
   face_axes(d, sa, fa);
   const int to = (l%2) ? g.n[d] + 1 : 0, from = (l%2) ? g.n[d] : 1;
   int s0 = 1, s1 = g.n[sa], f0 = 1, f1 = g.n[fa];
   if (c->p.stencil != 7) { s0 = 0; s1 = g.n[sa] + 1; f0 = 0; f1 = g.n[fa] + 1; }
   c->ops_main[d].push_back(tile_copy(g, d, slot, to, slot, from, s0, s1, f0, f1));
}

// quarter of a coarse face, cases 6-9 (comm.c:282-295, 1029-1042)
inline void quarter_range(const Geometry &g, int fc, int sa, int fa, int &s0, int &s1,
                          int &f0, int &f1)
{
   const int hs = g.n[sa]/2, hf = g.n[fa]/2;
   if (fc%2 == 0) { s0 = 1; s1 = hs; } else { s0 = hs + 1; s1 = g.n[sa]; }
   if ((fc/2)%2 == 1) { f0 = 1; f1 = hf; } else { f0 = hf + 1; f1 = g.n[fa]; }
}

// pack_face code 0, comm.c:254-401 -> device send buffer at comm_send_off
void add_pack(mamr_ctx *c, int d, int slot, int fc, int off)
{
   const Geometry &g = c->g;
   int sa, fa, s0, s1, f0, f1;
   face_axes(d, sa, fa);
   int plane = 1;
   if (fc >= 10) { plane = g.n[d]; fc -= 10; }
   const long long S = g.str[sa], F = g.str[fa], N = g.str[d];
   FaceOp op;
   op.mem = MEM_DST_SEND;
   op.src_vs = g.var_stride;
   op.src_S = (int)S;
   op.src_F = (int)F;
   if (fc < 2) {
      whole_extent(g, fc == 1, d, sa, s0, s1);
      whole_extent(g, fc == 1, d, fa, f0, f1);
      op.mode = FM_COPY;
      op.Ns = s1 - s0 + 1; op.Nf = f1 - f0 + 1;
      op.src_base = tile_base(g, slot) + plane*N + s0*S + f0*F;
   } else if (fc <= 5) {
      op.mode = FM_SUM4;
      op.Ns = g.n[sa]/2; op.Nf = g.n[fa]/2;
      op.src_base = tile_base(g, slot) + plane*N + S + F;
   } else {
      quarter_range(g, fc, sa, fa, s0, s1, f0, f1);
      op.mode = FM_DIV4;
      op.Ns = s1 - s0 + 1; op.Nf = f1 - f0 + 1;
      op.src_base = tile_base(g, slot) + plane*N + s0*S + f0*F;
   }
   op.dst_base = off;
   op.dst_S = op.Nf;
   op.dst_F = 1;
   op.dst_vs = (long long)op.Ns*op.Nf;
   c->ops_main[d].push_back(op);
}

// unpack_face code 0, comm.c:1002-1150 (the case is the receiver's own)
void add_unpack(mamr_ctx *c, int d, int slot, int fc, int off)
{
   const Geometry &g = c->g;
   int sa, fa, s0, s1, f0, f1;
   face_axes(d, sa, fa);
   int plane = 0;
   if (fc >= 10) { plane = g.n[d] + 1; fc -= 10; }
   const long long S = g.str[sa], F = g.str[fa], N = g.str[d];
   FaceOp op;
   op.mem = MEM_SRC_RECV;
   op.dst_vs = g.var_stride;
   op.dst_S = (int)S;
   op.dst_F = (int)F;
   op.src_base = off;
   op.src_F = 1;
   if (fc < 2) {
      whole_extent(g, fc == 1, d, sa, s0, s1);
      whole_extent(g, fc == 1, d, fa, f0, f1);
      op.mode = FM_COPY;
      op.Ns = s1 - s0 + 1; op.Nf = f1 - f0 + 1;
      op.dst_base = tile_base(g, slot) + plane*N + s0*S + f0*F;
      op.src_S = op.Nf;
      op.src_vs = (long long)op.Ns*op.Nf;
   } else if (fc <= 5) {
      const int hs = g.n[sa]/2, hf = g.n[fa]/2;
      op.mode = FM_REPL;
      op.Ns = g.n[sa]; op.Nf = g.n[fa];
      op.dst_base = tile_base(g, slot) + plane*N + S + F;
      op.src_S = hf;
      op.src_vs = (long long)hs*hf;
   } else {
      quarter_range(g, fc, sa, fa, s0, s1, f0, f1);
      op.mode = FM_COPY;
      op.Ns = s1 - s0 + 1; op.Nf = f1 - f0 + 1;
      op.dst_base = tile_base(g, slot) + plane*N + s0*S + f0*F;
      op.src_S = op.Nf;
      op.src_vs = (long long)op.Ns*op.Nf;
   }
   c->ops_unpack[d].push_back(op);
}

// Peer-memory transport: lay this rank's receive buffers out in its window, publish where
// every partner's message starts (the partner's push kernel reads the table through its
// mapping of the window) and rebuild the push / credit lists.  Runs on every rank after
// every mamr_set_comm_lists (build_ops).  The previous table is not read any more: every
// push that used it was awaited by this rank before the topology could change.
int p2p_publish(mamr_ctx *c, const size_t rneed[3])
{
   long long off = 0;
   long long region[mamr_ctx::MAX_SETS][3];
   for (int q = 0; q < c->nsets; q++)
      for (int d = 0; d < 3; d++) {
         region[q][d] = off;
         off += ((long long)rneed[d] + 15)/16*16;
      }
   if ((size_t)off > c->win_data_cap)
      return fail(MAMR_EP2P, "peer-memory window too small: the receive buffers of %d set(s) need %lld doubles, "
                  "the window holds %zu (raise MAMR_P2P_WINDOW_MB)", c->nsets, off, c->win_data_cap);
   double *data = reinterpret_cast<double *>(c->win + c->win_data_off);
   for (int q = 0; q < c->nsets; q++)
      for (int d = 0; d < 3; d++) c->d_recvs[q][d] = data + region[q][d];
   // rbase[set][dir][sender]
   std::vector<long long> rbase((size_t)P2P_MAX_SETS*3*P2P_MAX_RANKS, -1);
   std::vector<P2PTarget> credit;
   for (int d = 0; d < 3; d++) {
      const DirLists &L = c->cl[d];
      std::vector<P2PTarget> push;
      c->push_max[d] = 0;
      for (size_t i = 0; i < L.partner.size(); i++) {
         for (int q = 0; q < c->nsets; q++)
            rbase[((size_t)q*3 + d)*P2P_MAX_RANKS + L.partner[i]] = region[q][d] + L.recv_off[L.index[i]];
         P2PTarget t;
         t.rank = L.partner[i]; t.dir = d;
         t.send_off = L.send_off[L.index[i]]; t.size = L.send_size[i];
         push.push_back(t);
         credit.push_back(t);
         c->push_max[d] = std::max(c->push_max[d], t.size);
      }
      if (!push.empty())
         CU(cudaMemcpyAsync(c->d_push[d], push.data(), push.size()*sizeof(P2PTarget), cudaMemcpyHostToDevice,
                            c->stream));
   }
   c->n_credit = (int)credit.size();
   if (!credit.empty())
      CU(cudaMemcpyAsync(c->d_credit, credit.data(), credit.size()*sizeof(P2PTarget), cudaMemcpyHostToDevice,
                         c->stream));
   CU(cudaMemcpyAsync(c->win + offsetof(P2PHeader, rbase), rbase.data(), rbase.size()*sizeof(long long),
                      cudaMemcpyHostToDevice, c->stream));
   CU(cudaStreamSynchronize(c->stream));      // the table is in memory before the epoch says so
   CU(cudaMemcpyAsync(c->win + offsetof(P2PHeader, epoch), &c->p2p_epoch, sizeof(unsigned long long),
                      cudaMemcpyHostToDevice, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   return MAMR_OK;
}

// The on-rank loop of comm(), comm.c:162-203, flattened into descriptors.
int build_ops(mamr_ctx *c)
{
   const int nb = c->num_active;
   std::vector<int> slot2idx(c->p.max_blocks, -1);
   for (int a = 0; a < nb; a++) slot2idx[c->blocks[a].slot] = a;
   for (int d = 0; d < 3; d++) {
      c->ops_main[d].clear();
      c->ops_unpack[d].clear();
      c->n_same[d] = c->n_diff[d] = c->n_bc[d] = 0;
   }
   for (int a = 0; a < nb; a++) {
      const mamr_block &b = c->blocks[a];
      const int n = b.slot;
      for (int l = 0; l < 6; l++) {
         const int d = l/2, nl = b.nei_level[l];
         if (nl == b.level) {
            const int m = b.nei[l][0][0];
            if (m > n) {
               if (m >= c->p.max_blocks || slot2idx[m] < 0)
                  return fail(MAMR_ETOPOLOGY, "ERROR: misconnected block (slot %d face %d -> %d)", n, l, m);
               if (l%2 == 0) add_same(c, d, m, n);
               else add_same(c, d, n, m);
               c->n_same[d] += 2;
            }
         } else if (nl == b.level + 1) {
            for (int i = 0; i < 2; i++)
               for (int j = 0; j < 2; j++) {
                  const int m = b.nei[l][i][j];
                  if (m > n) {
                     if (m >= c->p.max_blocks || slot2idx[m] < 0)
                        return fail(MAMR_ETOPOLOGY, "ERROR: misconnected block (slot %d face %d -> %d)", n, l, m);
                     add_diff(c, n, m, l, i, j);
                     c->n_diff[d] += 2;
                  }
               }
         } else if (nl == b.level - 1) {
            const int m = b.nei[l][0][0];
            if (m > n) {
               if (m >= c->p.max_blocks || slot2idx[m] < 0)
                  return fail(MAMR_ETOPOLOGY, "ERROR: misconnected block (slot %d face %d -> %d)", n, l, m);
               const int k = 2*d + 1 - l%2;
               const mamr_block &cb = c->blocks[slot2idx[m]];
               for (int i = 0; i < 2; i++)
                  for (int j = 0; j < 2; j++)
                     if (cb.nei[k][i][j] == n) {
                        add_diff(c, m, n, k, i, j);
                        c->n_diff[d] += 2;
                     }
            }
         } else if (nl == -2) {
            add_bc(c, n, l);
            c->n_bc[d] += 1;
         } else
            return fail(MAMR_ETOPOLOGY, "ERROR: misconnected block (slot %d face %d level %d nei_level %d)",
                        n, l, b.level, nl);
      }
   }
   // off-rank faces: pack + unpack descriptors from the comm lists
   c->have_partners = false;
   size_t rneed[3] = {0, 0, 0};
   for (int d = 0; d < 3; d++) {
      const DirLists &L = c->cl[d];
      size_t smax = 0, rmax = 0;
      for (size_t i = 0; i < L.partner.size(); i++) {
         c->have_partners = true;
         for (int f = L.index[i]; f < L.index[i] + L.num[i]; f++) {
            add_pack(c, d, L.block[f], L.face_case[f], L.send_off[f]);
            add_unpack(c, d, L.block[f], L.face_case[f], L.recv_off[f]);
         }
         smax = std::max(smax, (size_t)L.send_off[L.index[i]] + (size_t)L.send_size[i]);
         rmax = std::max(rmax, (size_t)L.recv_off[L.index[i]] + (size_t)L.recv_size[i]);
      }
      rneed[d] = rmax;
      if (smax > c->send_cap[d]) {
         CK(dfree(c, c->d_send[d]));
         c->d_send[d] = nullptr;
         CK(dalloc(c, &c->d_send[d], smax*sizeof(double)));
         CU(cudaMemsetAsync(c->d_send[d], 0, smax*sizeof(double), c->stream));
         c->send_cap[d] = smax;
      }
      if (!c->p2p && rmax > c->recv_cap[d]) {
         for (int q = 0; q < c->nsets; q++) {
            CK(dfree(c, c->d_recvs[q][d]));
            c->d_recvs[q][d] = nullptr;
            CK(dalloc(c, &c->d_recvs[q][d], rmax*sizeof(double)));
            CU(cudaMemsetAsync(c->d_recvs[q][d], 0, rmax*sizeof(double), c->stream));
         }
         c->recv_cap[d] = rmax;
      }
   }
   if (c->p2p) CK(p2p_publish(c, rneed));
   // upload
   size_t total = 0;
   for (int d = 0; d < 3; d++) total += c->ops_main[d].size() + c->ops_unpack[d].size();
   if (total > c->ops_cap) {
      CU(cudaStreamSynchronize(c->stream));
      CK(dfree(c, c->d_ops));
      c->d_ops = nullptr;
      c->ops_cap = total + total/4 + 64;
      CK(dalloc(c, &c->d_ops, c->ops_cap*sizeof(FaceOp)));
   }
   std::vector<FaceOp> all;
   all.reserve(total);
   for (int d = 0; d < 3; d++) {
      c->off_main[d] = all.size();
      all.insert(all.end(), c->ops_main[d].begin(), c->ops_main[d].end());
   }
   for (int d = 0; d < 3; d++) {
      c->off_unpack[d] = all.size();
      all.insert(all.end(), c->ops_unpack[d].begin(), c->ops_unpack[d].end());
   }
   if (total) {
      // the previous descriptors may still be in use by queued launches
      CU(cudaStreamSynchronize(c->stream));
      CU(cudaMemcpyAsync(c->d_ops, all.data(), total*sizeof(FaceOp), cudaMemcpyHostToDevice,
                         c->stream));
      CU(cudaStreamSynchronize(c->stream));
   }
   c->ops_dirty = false;
   return MAMR_OK;
}

// the main stream is about to touch the message buffers: an exchange still in
// flight on the exchange stream comes first
int wait_xchg(mamr_ctx *c)
{
   if (!c->xchg_pending) return MAMR_OK;
   CU(cudaStreamWaitEvent(c->stream, c->ev_xchg, 0));
   c->xchg_pending = false;
   return MAMR_OK;
}


// Device memory for descriptors and staging areas that come and go with the topology.
// cudaMalloc() and cudaFree() (and a stream-ordered pool that has to grow) may wait for the
// whole device; when several ranks share this process (loopback over one GPU) another rank's
// kernel may be spinning on a flag this rank has yet to raise.  There the memory comes out of
// an arena set aside when the windows are connected: bump allocation plus a free list by size;
// a block is reused in stream order on the main stream, which has waited for the exchange and
// boundary streams whenever descriptors are replaced.
template <typename T> int dalloc(mamr_ctx *c, T **p, size_t bytes)
{
   if (c->arena) {
      const size_t sz = (std::max<size_t>(bytes, 1) + 255)/256*256;
      auto it = c->arena_free.lower_bound(sz);
      if (it != c->arena_free.end() && it->first <= 2*sz) {
         *p = (T *)it->second;
         c->arena_free.erase(it);
         return MAMR_OK;
      }
      if (c->arena_top + sz <= c->arena_cap) {
         *p = (T *)(c->arena + c->arena_top);
         c->arena_size[(void *)*p] = sz;
         c->arena_top += sz;
         return MAMR_OK;
      }
      static bool warned = false;
      if (!warned) fprintf(stderr, "miniamr_b200: in-process arena exhausted (MAMR_INPROC_ARENA_MB); falling back to cudaMalloc\n");
      warned = true;
   }
   CU(cudaMalloc((void **)p, bytes));
   return MAMR_OK;
}

int dfree(mamr_ctx *c, void *p)
{
   if (!p) return MAMR_OK;
   auto it = c->arena_size.find(p);
   if (it != c->arena_size.end()) c->arena_free.insert({ it->second, p });
   else if (c->p2p_inproc) c->garbage.push_back(p);     // cudaMalloc'ed: release put off until mamr_destroy
   else CU(cudaFree(p));
   return MAMR_OK;
}

// a wait of this rank's exchange kernels timed out (p2p.cu: spin_ge)?  Call after a
// synchronisation of the stream.
int p2p_check(mamr_ctx *c)
{
   if (!c->p2p) return MAMR_OK;
   CU(cudaMemcpyAsync(c->h_p2p_err, c->win + offsetof(P2PHeader, error), sizeof(unsigned long long),
                      cudaMemcpyDeviceToHost, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   if (*c->h_p2p_err) {
      const unsigned long long e = *c->h_p2p_err;
      const char *what = (e >> 8) == 1 ? "receive-buffer credit" : (e >> 8) == 2 ? "comm-list epoch"
                         : (e >> 8) == 3 ? "ghost message" : (e >> 8) == 4 ? "check_sum contribution"
                         : (e >> 8) == 5 ? "migrated-block table" : "migration acknowledgement";
      return fail(MAMR_EP2P, "peer-memory transport: rank %d waited %.0f s for the %s of rank %d",
                  c->p.rank, c->p2p_timeout_s, what, (int)(e & 0xff));
   }
   return MAMR_OK;
}

// A ghost exchange on receive-buffer set `set` begins (every rank, partners or not, once per
// comm() call: the exchange numbers of any two ranks agree).  The consumers of the set's
// previous contents are ahead of `st`: tell the partners that they may overwrite it.
int p2p_begin(mamr_ctx *c, cudaStream_t st, int set)
{
   if (!c->p2p) return MAMR_OK;
   ++c->xseq[set];
   if (c->n_credit > 0) {
      launch_p2p_credit(c->d_credit, c->n_credit, c->d_peer_win, c->p.rank, set, c->xseq[set], st);
      c->cnt.kernel_launches++;
   }
   return MAMR_OK;
}

// one message per (direction, partner), comm.c:71-84 / 120-151: stores into the partners'
// windows (p2p.cu), or NCCL send/recv
void count_messages(mamr_ctx *c, int d)
{
   const DirLists &L = c->cl[d];
   if (c->send_faces) {
      // --send_faces: the reference posts one message per face (comm.c:59-77, 96-128)
      for (size_t i = 0; i < L.partner.size(); i++)
         for (int f = L.index[i]; f < L.index[i] + L.num[i]; f++) {
            const int j = L.face_case[f]%10;
            const int st = j == 0 ? 0 : (j == 1 ? 1 : (j < 6 ? 2 : 3)), rt = j < 2 ? st : 5 - st;
            c->cnt.counter_halo_recv[d]++;
            c->cnt.counter_halo_send[d]++;
            c->cnt.size_mesg_recv[d] += (double)c->comm_vars*c->msg_len[d][rt]*sizeof(double);
            c->cnt.size_mesg_send[d] += (double)c->comm_vars*c->msg_len[d][st]*sizeof(double);
            c->cnt.counter_face_send[d]++;
            c->cnt.counter_face_recv[d]++;
         }
      return;
   }
   for (size_t i = 0; i < L.partner.size(); i++) {
      c->cnt.counter_halo_recv[d]++;
      c->cnt.counter_halo_send[d]++;
      c->cnt.size_mesg_recv[d] += (double)L.recv_size[i]*sizeof(double);
      c->cnt.size_mesg_send[d] += (double)L.send_size[i]*sizeof(double);
      c->cnt.counter_face_send[d] += L.num[i];
      c->cnt.counter_face_recv[d] += L.num[i];
   }
}

int p2p_push_dir(mamr_ctx *c, int d, cudaStream_t st, int set)
{
   const int np = (int)c->cl[d].partner.size();
   if (np == 0) return MAMR_OK;
   launch_p2p_push(c->d_push[d], np, c->push_max[d], c->d_send[d], c->d_peer_win, c->win, c->win_data_off,
                   c->d_push_done + d*P2P_MAX_RANKS, c->p.rank, set, d, c->xseq[set], c->p2p_epoch, st);
   c->cnt.kernel_launches++;
   count_messages(c, d);
   return MAMR_OK;
}

int p2p_wait_dir(mamr_ctx *c, int d, cudaStream_t st, int set)
{
   const int np = (int)c->cl[d].partner.size();
   if (np == 0) return MAMR_OK;
   launch_p2p_wait(c->d_push[d], np, c->win, set, d, c->xseq[set], st);
   c->cnt.kernel_launches++;
   return MAMR_OK;
}

int exchange_dir(mamr_ctx *c, int d, cudaStream_t st, int set)
{
   const DirLists &L = c->cl[d];
   if (c->p2p) {
      CK(p2p_push_dir(c, d, st, set));
      CK(p2p_wait_dir(c, d, st, set));
      CU(cudaGetLastError());
      return MAMR_OK;
   }
   NC(g_nccl.GroupStart());
   for (size_t i = 0; i < L.partner.size(); i++) {
      NC(g_nccl.Recv(c->d_recvs[set][d] + L.recv_off[L.index[i]], (size_t)L.recv_size[i], NCCL_DOUBLE,
                     L.partner[i], c->nccl, st));
      NC(g_nccl.Send(c->d_send[d] + L.send_off[L.index[i]], (size_t)L.send_size[i], NCCL_DOUBLE,
                     L.partner[i], c->nccl, st));
   }
   NC(g_nccl.GroupEnd());
   count_messages(c, d);
   return MAMR_OK;
}

// the three directions at once (their send buffers are packed): one round
int exchange_dirs(mamr_ctx *c, const int dirs[3], cudaStream_t st, int set)
{
   if (c->p2p) {
      for (int o = 0; o < 3; o++) CK(p2p_push_dir(c, dirs[o], st, set));
      for (int o = 0; o < 3; o++) CK(p2p_wait_dir(c, dirs[o], st, set));
      CU(cudaGetLastError());
      return MAMR_OK;
   }
   NC(g_nccl.GroupStart());
   for (int o = 0; o < 3; o++)
      if (!c->cl[dirs[o]].partner.empty()) CK(exchange_dir(c, dirs[o], st, set));
   NC(g_nccl.GroupEnd());
   return MAMR_OK;
}

// Execute one direction-phased comm() on variables [start, start+num) in place in their
// current pools: the split path.  buf_var0 is the first variable of the comm() call
// (variable 0 of the message buffers); with exchange == false the receive buffers already
// hold this call's messages.  Exactly one exchange per direction, however the variables
// are spread over the two pools (every rank issues the same number of exchanges).
int comm_split(mamr_ctx *c, int start, int num, int ord, int buf_var0, bool exchange, int set)
{
   if (c->ops_dirty) CK(build_ops(c));
   if (c->have_partners && !c->nccl && !c->p2p)
      return fail(MAMR_ENCCL, "comm: off-rank partners present but neither mamr_p2p_connect nor "
                  "mamr_nccl_init was called");
   CK(wait_xchg(c));
   if (exchange) CK(p2p_begin(c, c->stream, set));
   const std::vector<Run> runs = runs_of(c, start, num, false);
   for (int o = 0; o < 3; o++) {
      const int d = kPerm[ord][o];
      const DirLists &L = c->cl[d];
      if (!c->ops_main[d].empty())
         for (const Run &r : runs) {
            KTimer t(c, KC_GHOST, nullptr, KD_SPLIT);
            launch_ghost(c->d_ops + c->off_main[d], (int)c->ops_main[d].size(), vpool(c, r.start), c->d_send[d],
                         c->d_recvs[set][d], c->g.var_stride, r.start, r.num, buf_var0, c->stream);
            c->cnt.kernel_launches++;
         }
      if (!L.partner.empty()) {
         if (exchange) {
            KTimer t(c, KC_GHOST, nullptr, KD_XCHG);
            CK(exchange_dir(c, d, c->stream, set));
         }
         if (!c->ops_unpack[d].empty())
            for (const Run &r : runs) {
               KTimer t(c, KC_GHOST, nullptr, KD_UNPACK);
               launch_ghost(c->d_ops + c->off_unpack[d], (int)c->ops_unpack[d].size(), vpool(c, r.start),
                            c->d_send[d], c->d_recvs[set][d], c->g.var_stride, r.start, r.num, buf_var0,
                            c->stream);
               c->cnt.kernel_launches++;
            }
      }
   }
   CU(cudaGetLastError());
   return MAMR_OK;
}

// make the deferred comm() of variables [v0, v0+n) real (ghost cells in memory).
// Off-rank faces were already exchanged when the comm() was deferred.
int materialize_comm(mamr_ctx *c, int v0, int n)
{
   for (const Run &r : runs_of(c, v0, n, true)) {
      const int ord = c->pc_ord[r.start];
      if (ord < 0) continue;
      CK(comm_split(c, r.start, r.num, ord, c->pc_start[r.start], false, c->pc_set[r.start]));
      for (int v = r.start; v < r.start + r.num; v++) {
         c->pc_ord[v] = -1;
         c->spec[v] = 0;        // a look-ahead result for this exchange is void now
      }
   }
   return MAMR_OK;
}

int regen_ghosts(mamr_ctx *c, int v0, int n);

// receive-buffer set `set` is about to be overwritten: deferred exchanges and stale
// ghost layers that still read it become real (maximal runs of variables)
int release_recv_set(mamr_ctx *c, int set)
{
   const int V = c->p.num_vars;
   for (int v = 0; v < V;) {
      if (!(c->pc_ord[v] >= 0 && c->pc_set[v] == set)) { v++; continue; }
      int e = v;
      while (e < V && c->pc_ord[e] >= 0 && c->pc_set[e] == set) e++;
      CK(materialize_comm(c, v, e - v));
      v = e;
   }
   for (int v = 0; v < V;) {
      if (!(c->stale[v] && c->stale_set[v] == set)) { v++; continue; }
      int e = v;
      while (e < V && c->stale[e] && c->stale_set[e] == set) e++;
      CK(regen_ghosts(c, v, e - v));
      v = e;
   }
   return MAMR_OK;
}

// Streamed 7-point kernel: classify the six faces of every block.  A face that is
// a plain copy of a same-level neighbour's (or, at a reflective boundary, the
// block's own) interior plane is described by ONE source offset and fetched by bulk
// copies; any other face must consist of FM_COPY ops, executed cell by cell.
int build_slab_plan(mamr_ctx *c, int ord)
{
   const HaloPlan &P = c->plan[ord];
   for (void *p : { (void *)c->d_fsrc[ord], (void *)c->d_cops[ord], (void *)c->d_cbegin[ord] })
      CK(dfree(c, p));
   c->d_fsrc[ord] = nullptr; c->d_cops[ord] = nullptr; c->d_cbegin[ord] = nullptr;
   c->slab_ok[ord] = false;
   if (!c->slab_geom || !c->use_slab || !c->use_elide || c->p.stencil != 7 || !P.ok || !P.elidable)
      return MAMR_OK;
   const Geometry &g = c->g;
   const int N = g.n[0];
   const long long PL = g.str[0], SJ = g.str[1];
   const size_t nb = P.begin.size() - 1;
   std::vector<long long> fsrc(6*std::max<size_t>(1, nb), -1);
   std::vector<BoxOp> cops;
   std::vector<int> cbegin(nb + 1, 0);
   auto cell = [&](int i, int j, int k) { return (long long)i*PL + j*SJ + k; };
   for (size_t a = 0; a < nb; a++) {
      cbegin[a] = (int)cops.size();
      for (int o = P.begin[a]; o < P.begin[a + 1]; o++) {
         const BoxOp &op = P.ops[o];
         if (!(op.flags & BF_FACE)) continue;          // the 7-point stencil reads faces only
         if (op.mode != FM_COPY) return MAMR_OK;       // level boundary: not streamed
         int f = -1;
         if (op.ext[0] == 1 && op.ext[1] == N && op.ext[2] == N)
            f = op.dst_base == cell(0, 1, 1) ? 0 : (op.dst_base == cell(N + 1, 1, 1) ? 1 : -1);
         else if (op.ext[0] == N && op.ext[1] == 1 && op.ext[2] == N)
            f = op.dst_base == cell(1, 0, 1) ? 2 : (op.dst_base == cell(1, N + 1, 1) ? 3 : -1);
         else if (op.ext[0] == N && op.ext[1] == N && op.ext[2] == 1)
            f = op.dst_base == cell(1, 1, 0) ? 4 : (op.dst_base == cell(1, 1, N + 1) ? 5 : -1);
         const bool plain = f >= 0 && op.src_mem == BM_POOL && !(op.flags & BF_GHOST_SRC) &&
                            op.src_str[0] == PL && op.src_str[1] == SJ && op.src_str[2] == 1;
         if (plain) {
            const long long m = op.src_base/g.tile_stride, sc = op.src_base%g.tile_stride;
            if (f < 2) { fsrc[6*a + f] = op.src_base - 1; continue; }
            if (f < 4) { fsrc[6*a + f] = op.src_base - 1 - PL; continue; }
            if (sc == cell(1, 1, 1) || sc == cell(1, 1, N)) {
               fsrc[6*a + f] = m*2LL*N*N + (sc == cell(1, 1, 1) ? 0 : (long long)N*N);
               continue;
            }
         }
         // a whole same-level face out of a receive buffer is a compact N x N plane per
         // variable there (pack_face order, comm.c:266-270): bulk-copied as well.  The
         // buffer index travels in the top byte of the offset.
         if (f >= 0 && op.src_mem >= BM_BUF0 && op.src_mem < BM_BUF0 + 3 && op.src_vs == (long long)N*N &&
             op.src_base%2 == 0 && op.src_base < (1LL << 48)) {
            const int d = f/2, sa = d == 0 ? 1 : 0, fa = d == 2 ? 1 : 2;
            if (op.src_str[sa] == N && op.src_str[fa] == 1) {
               fsrc[6*a + f] = op.src_base | ((long long)(op.src_mem - BM_BUF0 + 1) << 56);
               continue;
            }
         }
         cops.push_back(op);
      }
      if ((int)cops.size() - cbegin[a] > slab7_max_cell_ops()) return MAMR_OK;
   }
   cbegin[nb] = (int)cops.size();
   CK(dalloc(c, &c->d_fsrc[ord], fsrc.size()*sizeof(long long)));
   CK(dalloc(c, &c->d_cops[ord], std::max<size_t>(1, cops.size())*sizeof(BoxOp)));
   CK(dalloc(c, &c->d_cbegin[ord], cbegin.size()*sizeof(int)));
   CU(cudaMemcpyAsync(c->d_fsrc[ord], fsrc.data(), fsrc.size()*sizeof(long long),
                      cudaMemcpyHostToDevice, c->stream));
   if (!cops.empty())
      CU(cudaMemcpyAsync(c->d_cops[ord], cops.data(), cops.size()*sizeof(BoxOp),
                         cudaMemcpyHostToDevice, c->stream));
   CU(cudaMemcpyAsync(c->d_cbegin[ord], cbegin.data(), cbegin.size()*sizeof(int),
                      cudaMemcpyHostToDevice, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   c->slab_ok[ord] = true;
   return MAMR_OK;
}

// build (once per topology and phase order) the halo plan of the fused kernel
int ensure_plan(mamr_ctx *c, int ord)
{
   if (c->plan_built[ord]) return MAMR_OK;
   PlanInput in;
   in.g = &c->g;
   in.stencil = c->p.stencil;
   in.max_blocks = c->p.max_blocks;
   in.blocks = &c->blocks;
   in.cl = c->cl;
   for (int o = 0; o < 3; o++) in.order[o] = kPerm[ord][o];
   HaloPlan &P = c->plan[ord];
   build_halo_plan(in, P);
   if (P.ok && P.max_ops > 64) {
      P.ok = false;
      P.why = "more than 64 halo ops on one block";
   }
   for (int o = 0; o < 3 && P.ok && c->have_partners; o++)
      if (!build_pack_plan(in, o, c->pack[ord][o], c->pack_fb[ord][o], P.why)) P.ok = false;
   c->plan_built[ord] = true;
   if (!P.ok) return MAMR_OK;
   CU(cudaStreamSynchronize(c->stream));
   CK(dfree(c, c->d_hops[ord]));
   CK(dfree(c, c->d_hbegin[ord]));
   c->d_hops[ord] = nullptr;
   c->d_hbegin[ord] = nullptr;
   CK(dalloc(c, &c->d_hops[ord], std::max<size_t>(1, P.ops.size())*sizeof(BoxOp)));
   CK(dalloc(c, &c->d_hbegin[ord], P.begin.size()*sizeof(int)));
   if (!P.ops.empty())
      CU(cudaMemcpyAsync(c->d_hops[ord], P.ops.data(), P.ops.size()*sizeof(BoxOp),
                         cudaMemcpyHostToDevice, c->stream));
   CU(cudaMemcpyAsync(c->d_hbegin[ord], P.begin.data(), P.begin.size()*sizeof(int),
                      cudaMemcpyHostToDevice, c->stream));
   // eliding launches: identity ops that the stencil does not read are dropped
   // (7-point: everything but the faces)
   CK(dfree(c, c->d_lops[ord]));
   CK(dfree(c, c->d_lbegin[ord]));
   c->d_lops[ord] = nullptr;
   c->d_lbegin[ord] = nullptr;
   c->plan_has_ident[ord] = false;
   for (const BoxOp &op : P.ops)
      if (op.flags & BF_IDENT) c->plan_has_ident[ord] = true;
   CK(dfree(c, c->d_zsrc[ord]));
   c->d_zsrc[ord] = nullptr;
   if (P.elidable && c->fused2_geom) {
      std::vector<BoxOp> lean;
      std::vector<int> lbegin(P.begin.size(), 0);
      const Geometry &g = c->g;
      const long long zslot = 2LL*g.n[0]*g.n[1];
      std::vector<long long> zsrc(2*std::max<size_t>(1, P.begin.size() - 1), -1);
      for (size_t a = 0; a + 1 < P.begin.size(); a++) {
         lbegin[a] = (int)lean.size();
         int first = 0;
         for (int o = P.begin[a]; o < P.begin[a + 1]; o++) {
            BoxOp op = P.ops[o];
            if (c->p.stencil == 7 && !(op.flags & BF_FACE)) continue;
            // a whole Z halo face that is a plain copy of a tile's k=1 / k=nz interior
            // plane (same-level neighbour or reflective boundary) arrives as one bulk
            // copy of that tile's export instead of nx*ny scattered 8-byte cells
            if (op.mode == FM_COPY && op.src_mem == BM_POOL && !(op.flags & BF_GHOST_SRC) &&
                op.ext[0] == g.n[0] && op.ext[1] == g.n[1] && op.ext[2] == 1 &&
                op.src_str[0] == g.str[0] && op.src_str[1] == g.str[1]) {
               const long long lo = g.str[0] + g.str[1], hi = lo + g.n[2] + 1;
               const long long m = op.src_base/g.tile_stride, cell = op.src_base%g.tile_stride;
               if ((op.dst_base == lo || op.dst_base == hi) &&
                   (cell == lo + 1 || cell == lo + g.n[2])) {
                  zsrc[2*a + (op.dst_base == hi ? 1 : 0)] =
                     m*zslot + (cell == lo + 1 ? 0 : (long long)g.n[0]*g.n[1]);
                  continue;
               }
            }
            op.first = first;
            first += op.ext[0]*op.ext[1]*op.ext[2];
            lean.push_back(op);
         }
      }
      lbegin[P.begin.size() - 1] = (int)lean.size();
      CK(dalloc(c, &c->d_lops[ord], std::max<size_t>(1, lean.size())*sizeof(BoxOp)));
      CK(dalloc(c, &c->d_lbegin[ord], lbegin.size()*sizeof(int)));
      if (!lean.empty())
         CU(cudaMemcpyAsync(c->d_lops[ord], lean.data(), lean.size()*sizeof(BoxOp),
                            cudaMemcpyHostToDevice, c->stream));
      CU(cudaMemcpyAsync(c->d_lbegin[ord], lbegin.data(), lbegin.size()*sizeof(int),
                         cudaMemcpyHostToDevice, c->stream));
      CK(dalloc(c, &c->d_zsrc[ord], zsrc.size()*sizeof(long long)));
      CU(cudaMemcpyAsync(c->d_zsrc[ord], zsrc.data(), zsrc.size()*sizeof(long long),
                         cudaMemcpyHostToDevice, c->stream));
      CU(cudaStreamSynchronize(c->stream));   // the host vectors go out of scope
   }
   CK(build_slab_plan(c, ord));
   // interior blocks (no halo cell out of a receive buffer) first, boundary blocks last
   CK(dfree(c, c->d_order_ord[ord]));
   c->d_order_ord[ord] = nullptr;
   c->n_interior[ord] = 0;
   if (c->have_partners && c->use_overlap && c->num_active > 0 &&
       (int)c->h_order.size() == c->num_active) {
      std::vector<char> bnd(c->num_active, 0);
      for (size_t a = 0; a + 1 < P.begin.size(); a++)
         for (int o = P.begin[a]; o < P.begin[a + 1]; o++)
            if (P.ops[o].src_mem != BM_POOL) { bnd[a] = 1; break; }
      std::vector<int> ordv;
      ordv.reserve(c->num_active);
      for (int a : c->h_order) if (!bnd[a]) ordv.push_back(a);
      c->n_interior[ord] = (int)ordv.size();
      for (int a : c->h_order) if (bnd[a]) ordv.push_back(a);
      CK(dalloc(c, &c->d_order_ord[ord], ordv.size()*sizeof(int)));
      CU(cudaMemcpyAsync(c->d_order_ord[ord], ordv.data(), ordv.size()*sizeof(int),
                         cudaMemcpyHostToDevice, c->stream));
      CU(cudaStreamSynchronize(c->stream));
   }
   for (int o = 0; o < 3; o++) {
      CK(dfree(c, c->d_pack[ord][o]));
      CK(dfree(c, c->d_pack_fb[ord][o]));
      c->d_pack[ord][o] = nullptr;
      c->d_pack_fb[ord][o] = nullptr;
      const std::vector<BoxOp> &K = c->pack[ord][o];
      const std::vector<int> &FB = c->pack_fb[ord][o];
      if (!c->have_partners || K.empty()) continue;
      CK(dalloc(c, &c->d_pack[ord][o], K.size()*sizeof(BoxOp)));
      CU(cudaMemcpyAsync(c->d_pack[ord][o], K.data(), K.size()*sizeof(BoxOp),
                         cudaMemcpyHostToDevice, c->stream));
      CK(dalloc(c, &c->d_pack_fb[ord][o], FB.size()*sizeof(int)));
      CU(cudaMemcpyAsync(c->d_pack_fb[ord][o], FB.data(), FB.size()*sizeof(int), cudaMemcpyHostToDevice,
                         c->stream));
   }
   CU(cudaStreamSynchronize(c->stream));
   return MAMR_OK;
}

// can a comm() with this phase order be deferred into the fused kernel?
int fused_ready(mamr_ctx *c, int ord, bool *yes)
{
   *yes = false;
   if (!c->use_fused || !(c->fused_geom || c->slab_geom) || c->num_active == 0) return MAMR_OK;
   if (c->have_partners && !c->nccl && !c->p2p) return MAMR_OK;   // comm_split reports the error
   CK(ensure_plan(c, ord));
   *yes = c->plan[ord].ok && (c->fused_geom || c->slab_ok[ord]);
   return MAMR_OK;
}

// Ghost cells an eliding fused launch left unwritten become real: the halo plan
// of that launch applied to the pool it read (fused2.cu: halo_fill_kernel).
int regen_ghosts(mamr_ctx *c, int v0, int n)
{
   int v = v0;
   for (int u = v0; u < v0 + n; u++)
      if (c->stale[u]) { CK(wait_xchg(c)); break; }
   while (v < v0 + n) {
      if (!c->stale[v]) { v++; continue; }
      int e = v + 1;
      while (e < v0 + n && c->stale[e] && c->cur[e] == c->cur[v] && c->stale_ord[e] == c->stale_ord[v] &&
             c->stale_start[e] == c->stale_start[v])
         e++;
      const int ord = c->stale_ord[v], in = c->cur[v] ^ 1;
      if (!c->plan_built[ord] || !c->plan[ord].ok || !c->d_hops[ord])
         return fail(MAMR_EINVAL, "internal: halo plan of an elided stage is gone");
      if (c->num_active > 0) {
         KTimer t(c, KC_GHOST, nullptr, KD_REGEN);
         double *const *rs = c->d_recvs[c->stale_set[v]];
         const double *recv[3] = { rs[0], rs[1], rs[2] };
         launch_halo_fill(c->d_hops[ord], c->d_hbegin[ord], c->d_slots, c->num_active, c->pool[in],
                          c->pool[in ^ 1], c->g, recv, v, e - v, c->stale_start[v], false, c->stream);
         c->cnt.kernel_launches++;
         c->cnt.ghost_regens++;
      }
      for (int u = v; u < e; u++) c->stale[u] = 0;
      v = e;
   }
   CU(cudaGetLastError());
   return MAMR_OK;
}

// stencil_check books its flops per cell (stencil.c:970-971, 975-976): the kernels count
// the two branches; the counts are folded into the counters wherever the stream is
// synchronised anyway (check_sum, mamr_sync), never by mamr_get_counters itself
int fold_s0_checks(mamr_ctx *c)
{
   if (c->p.stencil != 0 || !c->d_s0_chk) return MAMR_OK;
   CU(cudaMemcpyAsync(c->h_s0_chk, c->d_s0_chk, 2*sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                      c->stream));
   CU(cudaMemsetAsync(c->d_s0_chk, 0, 2*sizeof(unsigned long long), c->stream));
   CU(cudaStreamSynchronize(c->stream));
   const double nd = (double)c->h_s0_chk[0], nm = (double)c->h_s0_chk[1];
   c->cnt.total_fp_divs += nd;
   c->cnt.total_fp_adds += 2.0*nd + nm;
   c->cnt.total_fp_muls += nm;
   return MAMR_OK;
}

// stencil_driver() with --stencil 0 (stencil.c:43-74) for variables [v0, v0+n) of one
// pool: variable 0 and variables >= 4*mat take the 7-point average, the others the
// update kind pend_stage % 6 followed by stencil_check
int run_stencil0(mamr_ctx *c, int pool, int v0, int n)
{
   if (c->num_active <= 0) return MAMR_OK;
   if (!c->s0_set)
      return fail(MAMR_EINVAL, "--stencil 0: mamr_set_stencil0 (mat, a1, a0[] of init.c:418-423) was not called");
   const int mat = c->s0_mat, kind = c->pend_stage%6;
   if (c->s0_work_blocks < c->num_active && kind >= S0_SEVEN) {
      CU(cudaStreamSynchronize(c->stream));
      CK(dfree(c, c->d_s0_work));
      c->d_s0_work = nullptr;
      CK(dalloc(c, &c->d_s0_work, (size_t)c->p.max_blocks*c->g.tile_stride*sizeof(double)));
      c->s0_work_blocks = c->p.max_blocks;
   }
   int v = v0;
   while (v < v0 + n) {
      const bool plain = v == 0 || v >= 4*mat;
      int e = v + 1;
      while (e < v0 + n && (e == 0 || e >= 4*mat) == plain) e++;
      KTimer t(c, KC_STENCIL, nullptr, KD_STENCIL);
      if (plain)
         launch_stencil(c->pool[pool], c->g, c->d_slots, c->num_active, v, e - v, 7, c->stream);
      else
         launch_stencil0(c->pool[pool], c->g, c->d_slots, c->num_active, v, e, kind, mat, c->s0_a1,
                         c->d_s0_a0, c->d_s0_work, c->d_s0_chk, c->stream);
      c->cnt.kernel_launches++;
      v = e;
   }
   return MAMR_OK;
}

int flush_pending(mamr_ctx *c)
{
   if (c->pend_num == 0) return MAMR_OK;
   const int v0 = c->pend_start, n = c->pend_num;
   c->pend_num = 0;
   const std::vector<Run> runs = runs_of(c, v0, n, true);
   for (size_t ri = 0; ri < runs.size(); ri++) {
      Run r = runs[ri];
      const int nreq = r.num;         // variables the host has asked for
      const int ord = c->pc_ord[r.start];
      const int in = c->cur[r.start];
      auto commit = [&](int v, bool elide, bool cs) {
         c->cur[v] ^= 1;
         c->stale[v] = elide ? 1 : 0;
         c->zf_ok[v] = elide ? 1 : 0;
         c->stale_ord[v] = (signed char)ord;
         c->stale_start[v] = c->pc_start[v];
         c->stale_set[v] = c->pc_set[v];
         c->cs_fused[v] = cs ? 1 : 0;
         if (elide) c->shell_synced[v] = 1;
         c->spec[v] = 0;
      };
      if (ord >= 0 && c->spec[r.start]) {
         // computed by an earlier launch of this comm group (look-ahead): commit only
         for (int v = r.start; v < r.start + r.num; v++) {
            commit(v, c->spec_flags[v] & 1, c->spec_flags[v] & 2);
            c->pc_ord[v] = -1;
            if (c->spec_cs_ok[v]) {      // its check_sum is known already
               c->cs_cache[v] = c->spec_cs[v];
               c->cs_valid[v] = 1;
               c->spec_cs_ok[v] = 0;
            }
         }
         continue;
      }
      if (ord >= 0 && c->use_lookahead && ri + 1 == runs.size() && c->num_active > 0) {
         // look ahead: the variables right behind the queue that wait under the same comm()
         int e = r.start + r.num;
         while (e < c->p.num_vars && c->pc_ord[e] == ord && c->pc_start[e] == c->pc_start[r.start] &&
                c->pc_set[e] == c->pc_set[r.start] && c->cur[e] == in && !c->spec[e] && !c->stale[e])
            e++;
         r.num = e - r.start;
      }
      if (ord >= 0) {
         // comm() + stencil in one pass: current pool -> other pool
         double *const *rs = c->d_recvs[c->pc_set[r.start]];
         const double *recv[3] = { rs[0], rs[1], rs[2] };
         const bool slab = c->slab_ok[ord];
         const bool f2 = c->fused2_geom && c->use_fused2;
         const bool elide = slab || (f2 && c->use_elide && c->plan[ord].elidable && c->d_lops[ord]);
         if (elide && c->plan_has_ident[ord] && c->num_active > 0) {
            // never-written ghost regions: the output pool must already hold them
            bool synced = true;
            for (int v = r.start; v < r.start + r.num; v++) synced = synced && c->shell_synced[v];
            if (!synced) {
               CK(wait_xchg(c));
               KTimer t(c, KC_GHOST, nullptr, KD_REGEN);
               launch_halo_fill(c->d_hops[ord], c->d_hbegin[ord], c->d_slots, c->num_active,
                                c->pool[in], c->pool[in ^ 1], c->g, recv, r.start, r.num,
                                c->pc_start[r.start], true, c->stream);
               c->cnt.kernel_launches++;
            }
         }
         if (elide && c->num_active > 0) {
            // the input tiles' Z-face exports, where no eliding launch wrote them
            int v = r.start;
            while (v < r.start + r.num) {
               if (c->zf_ok[v]) { v++; continue; }
               int e = v;
               while (e < r.start + r.num && !c->zf_ok[e]) e++;
               KTimer t(c, KC_GHOST, nullptr, KD_REGEN);
               launch_zface_extract(c->pool[in], c->zf[in], c->g, c->d_slots, c->num_active, v,
                                    e - v, c->stream);
               c->cnt.kernel_launches++;
               v = e;
            }
         }
         // check_sum partials ride along (every active block is covered by the launch, or by
         // the interior + boundary pair)
         double *cspart = (c->use_cs_fused && (slab || f2) && c->launches_since_cs < 8*c->nsets)
                             ? c->d_cspart : nullptr;
         c->launches_since_cs++;
         const long long cs_stride = (long long)c->num_active*CS_WARPS;
         auto launch = [&](const int *order, int count, cudaStream_t st) {
            if (count <= 0) return;
            if (slab)
               launch_slab7(c->pool[in], c->pool[in ^ 1], c->g, c->d_slots, order,
                            count, c->d_fsrc[ord], c->d_cops[ord], c->d_cbegin[ord], recv,
                            r.start, r.num, c->pc_start[r.start], c->zf[in], c->zf[in ^ 1], cspart,
                            cs_stride, st);
            else if (f2)
               launch_fused2(c->pool[in], c->pool[in ^ 1], c->g, c->d_slots, order,
                             count, elide ? c->d_lops[ord] : c->d_hops[ord],
                             elide ? c->d_lbegin[ord] : c->d_hbegin[ord], recv, r.start, r.num,
                             c->pc_start[r.start], c->p.stencil, elide, c->zf[in], c->zf[in ^ 1],
                             c->d_zsrc[ord], cspart, cs_stride, st);
            else
               launch_fused(c->pool[in], c->pool[in ^ 1], c->g, c->d_slots, order,
                            count, c->d_hops[ord], c->d_hbegin[ord], recv, r.start, r.num,
                            c->pc_start[r.start], c->p.stencil, st);
            c->cnt.kernel_launches++;
         };
         if (c->xchg_pending && c->d_order_ord[ord] && c->n_interior[ord] > 0 &&
             c->n_interior[ord] < c->num_active) {
            // The exchange is still in flight.  Interior blocks do not need it: they
            // start now on the main stream.  The boundary blocks go to a second stream
            // that waits for the exchange only, so their CTAs fill the machine as the
            // interior launch drains instead of waiting for its tail (both launches
            // read the same pool and write disjoint tiles).  One timed interval.
            KTimer t(c, KC_STENCIL, nullptr, KD_FUSED);
            CU(cudaEventRecord(c->ev_pre, c->stream));
            launch(c->d_order_ord[ord], c->n_interior[ord], c->stream);
            trace_mark(c, 2, c->stream);
            CU(cudaStreamWaitEvent(c->bstream, c->ev_pre, 0));
            CU(cudaStreamWaitEvent(c->bstream, c->ev_xchg, 0));
            trace_mark(c, 3, c->bstream);
            launch(c->d_order_ord[ord] + c->n_interior[ord], c->num_active - c->n_interior[ord],
                   c->bstream);
            trace_mark(c, 4, c->bstream);
            CU(cudaEventRecord(c->ev_bdone, c->bstream));
            CU(cudaStreamWaitEvent(c->stream, c->ev_bdone, 0));
            CU(cudaStreamWaitEvent(c->stream, c->ev_xchg, 0));
            c->xchg_pending = false;
         } else {
            CK(wait_xchg(c));
            KTimer t(c, KC_STENCIL, nullptr, KD_FUSED);
            launch(c->d_order, c->num_active, c->stream);
         }
         if (c->num_active > 0)
            for (int v = r.start; v < r.start + nreq; v++) commit(v, elide, cspart != nullptr);
         for (int v = r.start; v < r.start + nreq; v++) c->pc_ord[v] = -1;
         for (int v = r.start + nreq; v < r.start + r.num; v++) {     // looked ahead: not committed
            if (cspart) c->cs_fused[v] = 0;   // its partials now describe the result, not the current data
            c->spec[v] = 1;
            c->spec_cs_ok[v] = 0;
            c->spec_flags[v] = (char)((elide ? 1 : 0) | (cspart ? 2 : 0));
         }
      } else {
         CK(regen_ghosts(c, r.start, r.num));   // the in-place stencil reads the stored ghosts
         for (int v = r.start; v < r.start + r.num; v++) c->zf_ok[v] = 0;
         if (c->p.stencil == 0)
            CK(run_stencil0(c, in, r.start, r.num));
         else {
            KTimer t(c, KC_STENCIL, nullptr, KD_STENCIL);
            launch_stencil(c->pool[in], c->g, c->d_slots, c->num_active, r.start, r.num,
                           c->p.stencil, c->stream);
            if (c->num_active > 0) c->cnt.kernel_launches++;
         }
      }
   }
   CU(cudaGetLastError());
   return MAMR_OK;
}

// everything queued for [v0, v0+n) is applied to the interiors; a deferred comm()
// becomes real.  Ghost cells an eliding launch skipped may still be stale.
int settle_data(mamr_ctx *c, int v0, int n)
{
   CK(flush_pending(c));
   return materialize_comm(c, v0, n);
}

// ... and every ghost cell in memory holds what the reference would hold
int settle(mamr_ctx *c, int v0, int n)
{
   CK(settle_data(c, v0, n));
   return regen_ghosts(c, v0, n);
}

int settle_all(mamr_ctx *c) { return settle(c, 0, c->p.num_vars); }

int check_slot(mamr_ctx *c, int slot)
{
   if (!c) return fail(MAMR_EINVAL, "null context");
   if (slot < 0 || slot >= c->p.max_blocks)
      return fail(MAMR_EINVAL, "slot %d out of range [0,%d)", slot, c->p.max_blocks);
   return MAMR_OK;
}

// Order in which the fused kernel visits the blocks: neighbouring blocks should be
// in flight at about the same time, so that a halo row pulled from a neighbour tile
// and that tile's own load share one trip from DRAM.  (MAMR_ORDER=zchain is the
// older order: +z chains started in sorted_list order.)
std::vector<int> processing_order(const mamr_ctx *c)
{
   const int nb = c->num_active;
   std::vector<int> slot2idx(c->p.max_blocks, -1), order;
   std::vector<char> done(nb, 0);
   for (int a = 0; a < nb; a++) slot2idx[c->blocks[a].slot] = a;
   auto link = [&](int a, int l) {   // same-level on-rank neighbour through face l, or -1
      const mamr_block &b = c->blocks[a];
      if (b.nei_level[l] != b.level) return -1;
      const int m = b.nei[l][0][0];
      return (m >= 0 && m < c->p.max_blocks) ? slot2idx[m] : -1;
   };
   order.reserve(nb);
   // Default: bricks of 4 x 4 x 4 blocks, z fastest inside a brick.  Integer coordinates
   // come from the same-level face links (one origin per connected same-level region).
   // x, y AND z neighbours are then visited within a few hundred CTAs of each other, i.e.
   // while their rows are still in L2: 2.60 -> 2.53 ms on cfg2 (z-chains only keep x and z
   // neighbours close; y halo rows had left the L2).  MAMR_ORDER=zchain | brick:bx,by,bz.
   int bx = 4, by = 4, bz = 4;
   if (const char *e = getenv("MAMR_ORDER")) {
      if (!strcmp(e, "zchain")) bx = 0;
      else if (sscanf(e, "brick:%d,%d,%d", &bx, &by, &bz) != 3 || bx < 1 || by < 1 || bz < 1)
         bx = by = bz = 4;
   }
   if (bx > 0 && nb > 0) {
      std::vector<int> cx(nb, 0), cy(nb, 0), cz(nb, 0), comp(nb, -1), stack;
      int ncomp = 0;
      for (int a0 = 0; a0 < nb; a0++) {
         if (comp[a0] >= 0) continue;
         comp[a0] = ncomp;
         stack.push_back(a0);
         while (!stack.empty()) {
            const int a = stack.back();
            stack.pop_back();
            for (int l = 0; l < 6; l++) {
               const int m = link(a, l);
               if (m < 0 || comp[m] >= 0) continue;
               comp[m] = ncomp;
               cx[m] = cx[a] + (l == 0 ? -1 : (l == 1 ? 1 : 0));
               cy[m] = cy[a] + (l == 2 ? -1 : (l == 3 ? 1 : 0));
               cz[m] = cz[a] + (l == 4 ? -1 : (l == 5 ? 1 : 0));
               stack.push_back(m);
            }
         }
         ncomp++;
      }
      std::vector<int> mnx(ncomp, 1 << 30), mny(ncomp, 1 << 30), mnz(ncomp, 1 << 30);
      for (int a = 0; a < nb; a++) {
         mnx[comp[a]] = std::min(mnx[comp[a]], cx[a]);
         mny[comp[a]] = std::min(mny[comp[a]], cy[a]);
         mnz[comp[a]] = std::min(mnz[comp[a]], cz[a]);
      }
      struct Key { int comp, Bz, By, Bx, y, x, z, a; };
      std::vector<Key> keys(nb);
      for (int a = 0; a < nb; a++) {
         const int x = cx[a] - mnx[comp[a]], y = cy[a] - mny[comp[a]], z = cz[a] - mnz[comp[a]];
         keys[a] = { comp[a], z/bz, y/by, x/bx, y%by, x%bx, z%bz, a };
      }
      std::sort(keys.begin(), keys.end(), [](const Key &p, const Key &q) {
         return std::tie(p.comp, p.Bz, p.By, p.Bx, p.y, p.x, p.z, p.a) <
                std::tie(q.comp, q.Bz, q.By, q.Bx, q.y, q.x, q.z, q.a);
      });
      for (const Key &k : keys) order.push_back(k.a);
      return order;
   }
   for (int a0 = 0; a0 < nb; a0++) {
      if (done[a0]) continue;
      int a = a0;
      for (int steps = 0; steps < nb; steps++) {   // walk to the -z end of the chain
         const int m = link(a, 4);
         if (m < 0 || done[m]) break;
         a = m;
      }
      while (a >= 0 && !done[a]) {
         done[a] = 1;
         order.push_back(a);
         a = link(a, 5);
      }
   }
   return order;
}

// Block data changed behind the stencil's back.  With more than one rank check_sum() is a
// collective: whether a rank joins it (or serves its cache) and how many variables it reduces
// may only depend on events every rank sees -- stencil calls, mamr_set_topology,
// mamr_flush_block_moves (`uniform`).  A rank-local change (upload, split, consolidate,
// unpack) therefore leaves the cache decision alone and marks the cached sums as not
// servable: check_sum() of such a variable without a uniform event in between is an error,
// not a hang.
void touch_all(mamr_ctx *c, bool uniform = false)
{
   if (c->p.num_ranks > 1 && !uniform)
      std::fill(c->cs_local_dirty.begin(), c->cs_local_dirty.end(), 1);
   else {
      std::fill(c->cs_valid.begin(), c->cs_valid.end(), 0);
      std::fill(c->cs_local_dirty.begin(), c->cs_local_dirty.end(), 0);
      c->modified_since_cs = true;
   }
   std::fill(c->cs_fused.begin(), c->cs_fused.end(), 0);
   std::fill(c->spec.begin(), c->spec.end(), 0);
   std::fill(c->shell_synced.begin(), c->shell_synced.end(), 0);
   std::fill(c->zf_ok.begin(), c->zf_ok.end(), 0);
}

// mamr_flush_block_moves over the windows.  Sender: table (destination, ordinal among the sends
// to that destination) next to the payloads it staged, then the ready flag.  Receiver: look
// its receives up in the senders' tables, fetch the payloads straight into the tiles, tell the
// senders.  The sender's staging area is free again when every destination has said so.
int p2p_flush_moves(mamr_ctx *c)
{
   ++c->mv_seq;
   if (c->mv_send.empty() && c->mv_recv.empty()) return MAMR_OK;
   CK(settle_all(c));
   const int R = c->p.num_ranks;
   std::vector<int> ranks;
   if (!c->mv_send.empty()) {
      const int ns = (int)c->mv_send.size();
      std::vector<int> tab((size_t)2*c->mv_cap_p2p, -1), per(R, 0);
      for (int k = 0; k < ns; k++) {
         tab[k] = c->mv_send[k].peer;
         tab[c->mv_cap_p2p + k] = per[c->mv_send[k].peer]++;
      }
      CU(cudaMemcpyAsync(c->win + c->win_mv_off, tab.data(), (size_t)ns*sizeof(int), cudaMemcpyHostToDevice,
                         c->stream));
      CU(cudaMemcpyAsync(c->win + c->win_mv_off + (size_t)c->mv_cap_p2p*sizeof(int), tab.data() + c->mv_cap_p2p,
                         (size_t)ns*sizeof(int), cudaMemcpyHostToDevice, c->stream));
      const unsigned long long cnt = (unsigned long long)ns;
      CU(cudaMemcpyAsync(c->win + offsetof(P2PHeader, mv_count), &cnt, sizeof cnt, cudaMemcpyHostToDevice,
                         c->stream));
      CU(cudaStreamSynchronize(c->stream));      // payloads, table and count are in memory ...
      CU(cudaMemcpyAsync(c->win + offsetof(P2PHeader, mv_ready), &c->mv_seq, sizeof c->mv_seq,
                         cudaMemcpyHostToDevice, c->stream));                      // ... before the flag says so
   }
   if (!c->mv_recv.empty()) {
      const int nr = (int)c->mv_recv.size();
      if ((size_t)nr > c->mv_moves_cap) {
         CK(dfree(c, c->d_mv_moves));
         CK(dfree(c, c->d_mv_k));
         c->d_mv_moves = nullptr; c->d_mv_k = nullptr;
         c->mv_moves_cap = (size_t)nr + nr/2 + 16;
         CK(dalloc(c, &c->d_mv_moves, c->mv_moves_cap*sizeof(P2PMove)));
         CK(dalloc(c, &c->d_mv_k, c->mv_moves_cap*sizeof(int)));
      }
      std::vector<P2PMove> mv(nr);
      std::vector<int> per(R, 0);
      std::vector<char> isrc(R, 0);
      for (int i = 0; i < nr; i++) {
         mv[i].slot = c->mv_recv[i].slot;
         mv[i].src = c->mv_recv[i].peer;
         mv[i].ordinal = per[mv[i].src]++;
         mv[i].pad = 0;
         isrc[mv[i].src] = 1;
      }
      for (int r = 0; r < R; r++) if (isrc[r]) ranks.push_back(r);
      CU(cudaMemcpyAsync(c->d_mv_moves, mv.data(), (size_t)nr*sizeof(P2PMove), cudaMemcpyHostToDevice, c->stream));
      CU(cudaMemcpyAsync(c->d_cur, c->cur.data(), (size_t)c->p.num_vars, cudaMemcpyHostToDevice, c->stream));
      CU(cudaMemcpyAsync(c->d_mv_ranks, ranks.data(), ranks.size()*sizeof(int), cudaMemcpyHostToDevice, c->stream));
      CU(cudaStreamSynchronize(c->stream));      // the host vectors go out of scope
      launch_p2p_mv_resolve(c->d_mv_moves, nr, c->d_peer_win, c->win, c->win_mv_off, c->mv_cap_p2p, c->p.rank,
                            c->mv_seq, c->d_mv_k, c->stream);
      launch_p2p_mv_unpack(c->d_mv_moves, nr, c->d_mv_k, c->d_peer_win, c->win_mv_off, c->mv_cap_p2p, c->pool[0],
                           c->pool[1], c->d_cur, c->p.nx, c->p.ny, c->p.nz, c->g.tile_stride, c->g.var_stride,
                           c->p.num_vars, c->stream);
      launch_p2p_mv_done(c->d_mv_ranks, (int)ranks.size(), c->d_peer_win, c->p.rank, c->mv_seq, c->stream);
      c->cnt.kernel_launches += 3;
   }
   if (!c->mv_send.empty()) {
      // my staging area is reused by the next load-balance step: every destination has fetched
      std::vector<char> isdst(R, 0);
      std::vector<int> dsts;
      for (const mamr_ctx::MoveRec &m : c->mv_send) isdst[m.peer] = 1;
      for (int r = 0; r < R; r++) if (isdst[r]) dsts.push_back(r);
      CU(cudaMemcpyAsync(c->d_mv_ranks + P2P_MAX_RANKS, dsts.data(), dsts.size()*sizeof(int), cudaMemcpyHostToDevice,
                         c->stream));
      CU(cudaStreamSynchronize(c->stream));
      launch_p2p_mv_wait(c->d_mv_ranks + P2P_MAX_RANKS, (int)dsts.size(), c->win, c->mv_seq, c->stream);
      c->cnt.kernel_launches++;
   }
   CU(cudaGetLastError());
   const size_t n = (size_t)c->p.num_vars*c->p.nx*c->p.ny*c->p.nz;
   c->cnt.migrate_bytes += (double)c->mv_send.size()*n*sizeof(double);
   if (!c->mv_recv.empty()) touch_all(c, true);
   c->mv_send.clear();
   c->mv_recv.clear();
   return MAMR_OK;
}

}  // namespace

extern "C" {

int mamr_abi_version(void) { return MAMR_ABI_VERSION; }

const char *mamr_last_error(void) { return g_err.c_str(); }

int mamr_create(const mamr_params *params, mamr_ctx **out)
{
   if (!params || !out) return fail(MAMR_EINVAL, "null argument");
   const mamr_params &p = *params;
   // main.c:657-723 check_input, the parts this path depends on
   if (p.nx <= 0 || p.ny <= 0 || p.nz <= 0 || (p.nx & 1) || (p.ny & 1) || (p.nz & 1))
      return fail(MAMR_EINVAL, "block size must be even and > 0 (main.c:673-684)");
   if (p.num_vars <= 0) return fail(MAMR_EINVAL, "num_vars must be > 0");
   if (p.max_blocks <= 0) return fail(MAMR_EINVAL, "max_blocks must be > 0");
   if (p.stencil != 0 && p.stencil != 7 && p.stencil != 27)
      return fail(MAMR_EINVAL, "--stencil %d: illegal value for stencil (main.c:701-704)", p.stencil);
   if (p.stencil == 0 && p.num_vars < 8)
      return fail(MAMR_EINVAL, "if stencil is 0, num_vars must be more than 8 (main.c:705-708)");
   // --code 1|2 ("send ghosts", "... and process on send", comm.c:403-989,1152-1461) change
   // what travels in a message and where the restriction runs.  In the reference itself
   // every cell the stencil reads ends up bit-identical to --code 0 -- EXCEPT with
   // --permute and a wide stencil (27 or 0), where the ghost edges and corners a later
   // phase forwards depend on the mode (checked at 1 and 4 ranks,
   // tests/test_plan_multi_rank.py).  The device path runs its code-0 exchange for the
   // equivalent combinations (the host's comm lists keep their larger code-1/2 offsets and
   // each face uses the front of its slot) and refuses the one that is not.
   if (p.code < 0 || p.code > 2)
      return fail(MAMR_EINVAL, "--code %d: must be 0, 1 or 2 (main.c:157)", p.code);
   if (p.code != 0 && p.permute && p.stencil != 7)
      return fail(MAMR_EUNSUPPORTED, "--code %d with --permute and --stencil %d: the reference's ghost edges "
                  "differ from --code 0 in this combination; only --code 0 is on the device path for it",
                  p.code, p.stencil);
   if (p.num_ranks < 1 || p.rank < 0 || p.rank >= p.num_ranks)
      return fail(MAMR_EINVAL, "bad rank %d of %d", p.rank, p.num_ranks);
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0)
      return fail(MAMR_ECUDA, "no CUDA device (%s): miniamr_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
   if (p.device >= 0) CU(cudaSetDevice(p.device));

   mamr_ctx *c = new mamr_ctx();
   c->p = p;
   c->comm_vars = (p.comm_vars <= 0 || p.comm_vars > p.num_vars) ? p.num_vars : p.comm_vars;
   Geometry &g = c->g;
   g.n[0] = p.nx; g.n[1] = p.ny; g.n[2] = p.nz;
   g.str[2] = 1; g.str[1] = p.nz + 2; g.str[0] = (p.ny + 2)*(p.nz + 2);
   g.tile = (p.nx + 2)*g.str[0];
   g.tile_stride = ((long long)g.tile + 15)/16*16;
   g.var_stride = g.tile_stride*p.max_blocks;
   memset(&c->cnt, 0, sizeof c->cnt);
   c->cs_valid.assign(p.num_vars, 0);
   c->cs_local_dirty.assign(p.num_vars, 0);
   c->cs_cache.assign(p.num_vars, 0.0);
   c->cur.assign(p.num_vars, 0);
   c->pc_ord.assign(p.num_vars, -1);
   c->pc_start.assign(p.num_vars, 0);
   c->stale.assign(p.num_vars, 0);
   c->stale_ord.assign(p.num_vars, 0);
   c->stale_start.assign(p.num_vars, 0);
   c->shell_synced.assign(p.num_vars, 0);
   c->zf_ok.assign(p.num_vars, 0);
   c->cs_fused.assign(p.num_vars, 0);
   c->spec.assign(p.num_vars, 0);
   c->spec_flags.assign(p.num_vars, 0);
   c->spec_cs_ok.assign(p.num_vars, 0);
   c->spec_cs.assign(p.num_vars, 0.0);
   { const char *e = getenv("MAMR_NO_LOOKAHEAD"); c->use_lookahead = !(e && e[0] == '1'); }
   { const char *e = getenv("MAMR_TRACE"); c->trace = e && e[0] == '1'; }
   { const char *e = getenv("MAMR_NO_FUSED_CS"); c->use_cs_fused = !(e && e[0] == '1'); }
   c->pc_set.assign(p.num_vars, 0);
   c->stale_set.assign(p.num_vars, 0);
   c->nsets = std::min<int>(mamr_ctx::MAX_SETS, (p.num_vars + c->comm_vars - 1)/c->comm_vars);
   std::string err, why;
   if (!stencil_configure(g, err) || !fused_configure(g, err) || !fused2_configure(g, err) ||
       !slab7_configure(g, err)) {
      delete c;
      return fail(MAMR_EUNSUPPORTED, "%s", err.c_str());
   }
   c->fused_geom = fused_supported(g, why);
   const char *nf = getenv("MAMR_NO_FUSED");
   c->use_fused = !(nf && nf[0] == '1') && p.stencil != 0;   // --stencil 0: split path + stencil0.cu
   c->fused2_geom = c->fused_geom && fused2_supported(g);
   const char *nf2 = getenv("MAMR_NO_FUSED2"), *ne = getenv("MAMR_NO_ELIDE");
   c->use_fused2 = !(nf2 && nf2[0] == '1');
   c->use_elide = !(ne && ne[0] == '1');
#define CUC(call)                                                                         \
   do {                                                                                   \
      cudaError_t e_ = (call);                                                            \
      if (e_ != cudaSuccess) {                                                            \
         int r_ = fail(MAMR_ECUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_),    \
                       __FILE__, __LINE__, cudaGetErrorString(e_));                       \
         mamr_destroy(c);                                                                 \
         return r_;                                                                       \
      }                                                                                   \
   } while (0)
   CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
   c->pool_bytes = (size_t)g.var_stride*p.num_vars*sizeof(double);
   for (int b = 0; b < 2; b++) {
      CUC(cudaMalloc(&c->pool[b], c->pool_bytes));
      CUC(cudaMemsetAsync(c->pool[b], 0, c->pool_bytes, c->stream));
   }
   const char *nsl = getenv("MAMR_NO_SLAB");
   c->use_slab = !(nsl && nsl[0] == '1');
   c->slab_geom = p.stencil == 7 && slab7_supported(g) && c->use_fused && c->use_slab && c->use_elide;
   if ((c->fused2_geom && c->use_fused2 && c->use_elide) || c->slab_geom) {
      c->zf_bytes = (size_t)2*p.nx*p.ny*p.max_blocks*p.num_vars*sizeof(double);
      for (int b = 0; b < 2; b++) CUC(cudaMalloc(&c->zf[b], c->zf_bytes));
   }
   if (p.stencil == 0) {
      CUC(cudaMalloc(&c->d_s0_a0, (size_t)p.num_vars*sizeof(double)));
      CUC(cudaMalloc(&c->d_s0_chk, 2*sizeof(unsigned long long)));
      CUC(cudaMemsetAsync(c->d_s0_chk, 0, 2*sizeof(unsigned long long), c->stream));
      CUC(cudaMallocHost(&c->h_s0_chk, 2*sizeof(unsigned long long)));
   }
   CUC(cudaMalloc(&c->d_sums, p.num_vars*sizeof(double)));
   CUC(cudaMallocHost(&c->h_sums, p.num_vars*sizeof(double)));
   CUC(cudaMalloc(&c->d_rops, c->rops_cap*sizeof(RefineOp)));
   CUC(cudaMalloc(&c->d_payload, (size_t)p.num_vars*p.nx*p.ny*p.nz*sizeof(double)));
   CUC(cudaMallocHost(&c->h_stage, (size_t)p.num_vars*g.tile*sizeof(double)));
   {
      // exchange stream: highest priority, so that its pack/NCCL CTAs are placed as
      // soon as CTAs of the interior stencil kernel retire
      int lo = 0, hi = 0;
      CUC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUC(cudaStreamCreateWithPriority(&c->xstream, cudaStreamNonBlocking, hi));
   }
   CUC(cudaEventCreateWithFlags(&c->ev_data, cudaEventDisableTiming));
   CUC(cudaEventCreateWithFlags(&c->ev_xchg, cudaEventDisableTiming));
   CUC(cudaStreamCreateWithFlags(&c->bstream, cudaStreamNonBlocking));
   CUC(cudaEventCreateWithFlags(&c->ev_pre, cudaEventDisableTiming));
   CUC(cudaEventCreateWithFlags(&c->ev_bdone, cudaEventDisableTiming));
   { const char *no = getenv("MAMR_NO_OVERLAP"); c->use_overlap = !(no && no[0] == '1'); }
   CUC(cudaEventCreate(&c->ev_begin));
   CUC(cudaEventCreate(&c->ev_end));
   CUC(cudaStreamSynchronize(c->stream));
#undef CUC
   *out = c;
   return MAMR_OK;
}

static void xfree(mamr_ctx *c, void *p)      // arena blocks go with the arena
{
   if (p && !c->arena_size.count(p)) cudaFree(p);
}

void mamr_destroy(mamr_ctx *c)
{
   if (!c) return;
   if (c->xstream) cudaStreamSynchronize(c->xstream);
   if (c->bstream) cudaStreamSynchronize(c->bstream);
   if (c->stream) cudaStreamSynchronize(c->stream);
   drain_ktimers(c);
   trace_drain(c);
   if (c->trace && c->tr_n > 0)
      fprintf(stderr, "miniamr_b200 trace rank %d: %lld overlapped stages; after comm(): exchange done +%.3f ms, "
              "interior blocks done +%.3f ms, boundary blocks start +%.3f ms, done +%.3f ms\n", c->p.rank, c->tr_n,
              c->tr_sum[0]/c->tr_n, c->tr_sum[1]/c->tr_n, c->tr_sum[2]/c->tr_n, c->tr_sum[3]/c->tr_n);
   for (cudaEvent_t e : c->ev_free) cudaEventDestroy(e);
   if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
   xfree(c, c->pool[0]);
   xfree(c, c->pool[1]);
   xfree(c, c->zf[0]);
   xfree(c, c->zf[1]);
   for (int o = 0; o < 6; o++) {
      xfree(c, c->d_hops[o]);
      xfree(c, c->d_hbegin[o]);
      xfree(c, c->d_lops[o]);
      xfree(c, c->d_lbegin[o]);
      xfree(c, c->d_zsrc[o]);
      xfree(c, c->d_fsrc[o]);
      xfree(c, c->d_cops[o]);
      xfree(c, c->d_cbegin[o]);
      for (int q = 0; q < 3; q++) {
         xfree(c, c->d_pack[o][q]);
         xfree(c, c->d_pack_fb[o][q]);
      }
   }
   xfree(c, c->d_slots);
   xfree(c, c->d_order);
   xfree(c, c->d_ops);
   for (int d = 0; d < 3; d++) {
      xfree(c, c->d_send[d]);
      if (!c->p2p)      // with the peer-memory transport they are part of the window
         for (int q = 0; q < mamr_ctx::MAX_SETS; q++) xfree(c, c->d_recvs[q][d]);
      xfree(c, c->d_push[d]);
   }
   for (size_t r = 0; r < c->peer_win.size(); r++)
      if (c->peer_ipc[r]) cudaIpcCloseMemHandle(c->peer_win[r]);
   xfree(c, c->win);
   xfree(c, c->d_peer_win);
   xfree(c, c->d_credit);
   xfree(c, c->d_push_done);
   xfree(c, c->d_mv_moves);
   xfree(c, c->d_mv_k);
   xfree(c, c->d_mv_ranks);
   xfree(c, c->d_cur);
   if (c->h_p2p_err) cudaFreeHost(c->h_p2p_err);
   for (void *g : c->garbage) cudaFree(g);
   cudaFree(c->arena);

   xfree(c, c->d_partials);
   xfree(c, c->d_cspart);
   xfree(c, c->d_sums);
   if (c->h_sums) cudaFreeHost(c->h_sums);
   xfree(c, c->d_rops);
   xfree(c, c->d_payload);
   xfree(c, c->d_s0_a0);
   xfree(c, c->d_s0_work);
   xfree(c, c->d_s0_chk);
   if (c->h_s0_chk) cudaFreeHost(c->h_s0_chk);
   xfree(c, c->d_mv_send);
   xfree(c, c->d_mv_recv);
   if (c->h_stage) cudaFreeHost(c->h_stage);
   for (int o = 0; o < 6; o++) xfree(c, c->d_order_ord[o]);
   if (c->ev_data) cudaEventDestroy(c->ev_data);
   if (c->ev_xchg) cudaEventDestroy(c->ev_xchg);
   for (int b = 0; b < 2; b++) {
      xfree(c, c->d_up[b]);
      if (c->ev_up_copy[b]) cudaEventDestroy(c->ev_up_copy[b]);
      if (c->ev_up_fill[b]) cudaEventDestroy(c->ev_up_fill[b]);
   }
   if (c->upstream) cudaStreamDestroy(c->upstream);
   if (c->ev_pre) cudaEventDestroy(c->ev_pre);
   if (c->ev_bdone) cudaEventDestroy(c->ev_bdone);
   if (c->xstream) cudaStreamDestroy(c->xstream);
   if (c->bstream) cudaStreamDestroy(c->bstream);
   if (c->ev_begin) cudaEventDestroy(c->ev_begin);
   if (c->ev_end) cudaEventDestroy(c->ev_end);
   if (c->stream) cudaStreamDestroy(c->stream);
   cudaGetLastError();
   delete c;
}

int mamr_sync(mamr_ctx *c)
{
   if (!c) return fail(MAMR_EINVAL, "null context");
   // queued stencils run; a deferred comm() and ghost layers an eliding stage left stale stay
   // lazy (whoever reads them makes them real) -- a sync must not cost a pass over the blocks
   CK(flush_pending(c));
   CK(wait_xchg(c));
   CU(cudaStreamSynchronize(c->stream));
   if (c->xstream) CU(cudaStreamSynchronize(c->xstream));
   CK(p2p_check(c));
   CK(fold_s0_checks(c));
   return MAMR_OK;
}

int mamr_get_counters(mamr_ctx *c, mamr_counters *out)
{
   if (!c || !out) return fail(MAMR_EINVAL, "null argument");
   *out = c->cnt;
   return MAMR_OK;
}

int mamr_reset_counters(mamr_ctx *c)
{
   if (!c) return fail(MAMR_EINVAL, "null context");
   memset(&c->cnt, 0, sizeof c->cnt);
   return MAMR_OK;
}

long long mamr_tile_doubles(mamr_ctx *c) { return c ? c->g.tile : 0; }
long long mamr_pool_bytes(mamr_ctx *c)
{
   return c ? 2*(long long)c->pool_bytes + (c->zf[0] ? 2*(long long)c->zf_bytes : 0) : 0;
}

// ---- block data in / out ---------------------------------------------------
int mamr_upload_block(mamr_ctx *c, int slot, const double *tiles)
{
   CK(check_slot(c, slot));
   if (!tiles) return fail(MAMR_EINVAL, "null tiles");
   CK(settle_all(c));
   const Geometry &g = c->g;
   for (const Run &r : runs_of(c, 0, c->p.num_vars, false))
      CU(cudaMemcpy2DAsync(vpool(c, r.start) + (long long)r.start*g.var_stride + tile_base(g, slot),
                           g.var_stride*sizeof(double), tiles + (size_t)r.start*g.tile,
                           g.tile*sizeof(double), g.tile*sizeof(double), r.num,
                           cudaMemcpyHostToDevice, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   touch_all(c);
   return MAMR_OK;
}

int mamr_download_block(mamr_ctx *c, int slot, double *tiles)
{
   CK(check_slot(c, slot));
   if (!tiles) return fail(MAMR_EINVAL, "null tiles");
   CK(settle_all(c));
   const Geometry &g = c->g;
   for (const Run &r : runs_of(c, 0, c->p.num_vars, false))
      CU(cudaMemcpy2DAsync(tiles + (size_t)r.start*g.tile, g.tile*sizeof(double),
                           vpool(c, r.start) + (long long)r.start*g.var_stride + tile_base(g, slot),
                           g.var_stride*sizeof(double), g.tile*sizeof(double), r.num,
                           cudaMemcpyDeviceToHost, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   CK(p2p_check(c));      // data behind a wait that gave up is not data
   return MAMR_OK;
}

int mamr_upload_tile(mamr_ctx *c, int slot, int var, const double *tile)
{
   CK(check_slot(c, slot));
   if (var < 0 || var >= c->p.num_vars || !tile) return fail(MAMR_EINVAL, "bad var %d", var);
   CK(settle(c, var, 1));
   const Geometry &g = c->g;
   CU(cudaMemcpyAsync(vpool(c, var) + (long long)var*g.var_stride + tile_base(g, slot), tile,
                      g.tile*sizeof(double), cudaMemcpyHostToDevice, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   if (c->p.num_ranks > 1) c->cs_local_dirty[var] = 1;      // see touch_all
   else {
      c->cs_valid[var] = 0;
      c->modified_since_cs = true;
   }
   c->cs_fused[var] = 0;
   c->shell_synced[var] = 0;
   c->zf_ok[var] = 0;
   return MAMR_OK;
}

int mamr_download_tile(mamr_ctx *c, int slot, int var, double *tile)
{
   CK(check_slot(c, slot));
   if (var < 0 || var >= c->p.num_vars || !tile) return fail(MAMR_EINVAL, "bad var %d", var);
   CK(settle(c, var, 1));
   const Geometry &g = c->g;
   CU(cudaMemcpyAsync(tile, vpool(c, var) + (long long)var*g.var_stride + tile_base(g, slot),
                      g.tile*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   return MAMR_OK;
}

int mamr_zero_block(mamr_ctx *c, int slot)
{
   CK(check_slot(c, slot));
   CK(settle_all(c));
   const Geometry &g = c->g;
   for (const Run &r : runs_of(c, 0, c->p.num_vars, false))
      CU(cudaMemset2DAsync(vpool(c, r.start) + (long long)r.start*g.var_stride + tile_base(g, slot),
                           g.var_stride*sizeof(double), 0, g.tile*sizeof(double), r.num, c->stream));
   touch_all(c);
   return MAMR_OK;
}

int mamr_upload_vars(mamr_ctx *c, int var_start, int num, int num_slots, const double *host)
{
   if (!c || !host) return fail(MAMR_EINVAL, "null argument");
   if (var_start < 0 || num <= 0 || var_start + num > c->p.num_vars || num_slots <= 0 ||
       num_slots > c->p.max_blocks)
      return fail(MAMR_EINVAL, "upload_vars: bad range vars [%d,%d) slots %d", var_start,
                  var_start + num, num_slots);
   CK(settle(c, var_start, num));
   const Geometry &g = c->g;
   for (int v = 0; v < num; v++)
      CU(cudaMemcpy2DAsync(vpool(c, var_start + v) + (long long)(var_start + v)*g.var_stride,
                           g.tile_stride*sizeof(double),
                           host + (size_t)v*num_slots*g.tile, g.tile*sizeof(double),
                           g.tile*sizeof(double), num_slots, cudaMemcpyHostToDevice, c->stream));
   touch_all(c);
   return MAMR_OK;
}

// The state as init.c:484-495 defines it -- interiors only, ghost layer zero -- for
// slots [0, num_slots): host[slot][var - var_start][nx][ny][nz], i.e. the block payloads
// of pack.c:66-70 back to back.  30 % fewer bytes over PCIe than whole tiles at 16^3;
// the copy of one chunk overlaps the scatter of the previous one.
int mamr_upload_interiors(mamr_ctx *c, int var_start, int num, int num_slots, const double *host)
{
   if (!c || !host) return fail(MAMR_EINVAL, "null argument");
   if (var_start < 0 || num <= 0 || var_start + num > c->p.num_vars || num_slots <= 0 ||
       num_slots > c->p.max_blocks)
      return fail(MAMR_EINVAL, "upload_interiors: bad range vars [%d,%d) slots %d", var_start,
                  var_start + num, num_slots);
   CK(settle(c, var_start, num));
   const Geometry &g = c->g;
   const size_t per_slot = (size_t)num*c->p.nx*c->p.ny*c->p.nz;         // doubles
   const size_t target = (size_t)16 << 20;                               // doubles per chunk (128 MB)
   const int S = (int)std::max<size_t>(1, std::min<size_t>((size_t)num_slots, target/per_slot));
   if (!c->upstream) {
      CU(cudaStreamCreateWithFlags(&c->upstream, cudaStreamNonBlocking));
      for (int b = 0; b < 2; b++) {
         CU(cudaEventCreateWithFlags(&c->ev_up_copy[b], cudaEventDisableTiming));
         CU(cudaEventCreateWithFlags(&c->ev_up_fill[b], cudaEventDisableTiming));
      }
   }
   if ((size_t)S*per_slot > c->up_cap) {
      CU(cudaStreamSynchronize(c->stream));
      CU(cudaStreamSynchronize(c->upstream));
      for (int b = 0; b < 2; b++) {
         CK(dfree(c, c->d_up[b]));
         c->d_up[b] = nullptr;
         CK(dalloc(c, &c->d_up[b], (size_t)S*per_slot*sizeof(double)));
         c->up_fill_rec[b] = false;
      }
      c->up_cap = (size_t)S*per_slot;
   }
   int k = 0;
   for (int s0 = 0; s0 < num_slots; s0 += S, k++) {
      const int b = k & 1, ns = std::min(S, num_slots - s0);
      // staging b is free again: the scatter that last read it (in this call or in an earlier,
      // still queued one) has finished
      if (c->up_fill_rec[b]) CU(cudaStreamWaitEvent(c->upstream, c->ev_up_fill[b], 0));
      CU(cudaMemcpyAsync(c->d_up[b], host + (size_t)s0*per_slot, (size_t)ns*per_slot*sizeof(double),
                         cudaMemcpyHostToDevice, c->upstream));
      CU(cudaEventRecord(c->ev_up_copy[b], c->upstream));
      CU(cudaStreamWaitEvent(c->stream, c->ev_up_copy[b], 0));
      for (const Run &r : runs_of(c, var_start, num, false)) {
         launch_fill_tiles(vpool(c, r.start), g, c->d_up[b], s0, ns, num, var_start, r.start, r.num,
                           c->stream);
         c->cnt.kernel_launches++;
      }
      CU(cudaEventRecord(c->ev_up_fill[b], c->stream));
      c->up_fill_rec[b] = true;
   }
   CU(cudaGetLastError());
   touch_all(c);
   return MAMR_OK;
}

int mamr_download_vars(mamr_ctx *c, int var_start, int num, int num_slots, double *host)
{
   if (!c || !host) return fail(MAMR_EINVAL, "null argument");
   if (var_start < 0 || num <= 0 || var_start + num > c->p.num_vars || num_slots <= 0 ||
       num_slots > c->p.max_blocks)
      return fail(MAMR_EINVAL, "download_vars: bad range vars [%d,%d) slots %d", var_start,
                  var_start + num, num_slots);
   CK(settle(c, var_start, num));
   const Geometry &g = c->g;
   for (int v = 0; v < num; v++)
      CU(cudaMemcpy2DAsync(host + (size_t)v*num_slots*g.tile, g.tile*sizeof(double),
                           vpool(c, var_start + v) + (long long)(var_start + v)*g.var_stride,
                           g.tile_stride*sizeof(double), g.tile*sizeof(double), num_slots,
                           cudaMemcpyDeviceToHost, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   return MAMR_OK;
}

// ---- topology --------------------------------------------------------------
int mamr_set_topology(mamr_ctx *c, int num_active, const mamr_block *sorted_blocks)
{
   if (!c || num_active < 0 || (num_active && !sorted_blocks))
      return fail(MAMR_EINVAL, "bad topology arguments");
   if (num_active > c->p.max_blocks)
      return fail(MAMR_EINVAL, "num_active %d > max_blocks %d", num_active, c->p.max_blocks);
   CK(settle_all(c));
   CK(wait_xchg(c));
   for (int a = 0; a < num_active; a++)
      if (sorted_blocks[a].slot < 0 || sorted_blocks[a].slot >= c->p.max_blocks)
         return fail(MAMR_EINVAL, "active block %d has slot %d out of range", a, sorted_blocks[a].slot);
   // A slot that just became inactive keeps its last content in the reference
   // (a later split_blocks()/unpack_block() only rewrites the interior, so the
   // ghost cells of the new block are whatever the slot held).  Make both pools
   // agree on it now: later pool flips then cannot change what a reused slot holds.
   {
      std::vector<char> now(c->p.max_blocks, 0);
      for (int a = 0; a < num_active; a++) now[sorted_blocks[a].slot] = 1;
      const Geometry &g = c->g;
      for (const mamr_block &b : c->blocks)
         if (!now[b.slot])
            for (const Run &r : runs_of(c, 0, c->p.num_vars, false))
               CU(cudaMemcpy2DAsync(c->pool[c->cur[r.start] ^ 1] + (long long)r.start*g.var_stride +
                                       tile_base(g, b.slot),
                                    g.var_stride*sizeof(double),
                                    vpool(c, r.start) + (long long)r.start*g.var_stride +
                                       tile_base(g, b.slot),
                                    g.var_stride*sizeof(double), g.tile*sizeof(double), r.num,
                                    cudaMemcpyDeviceToDevice, c->stream));
   }
   c->blocks.assign(sorted_blocks, sorted_blocks + num_active);
   c->num_active = num_active;
   if ((size_t)num_active > c->slots_cap) {
      CU(cudaStreamSynchronize(c->stream));
      CK(dfree(c, c->d_slots));
      CK(dfree(c, c->d_order));
      c->slots_cap = (size_t)num_active + num_active/4 + 64;
      CK(dalloc(c, &c->d_slots, c->slots_cap*sizeof(int)));
      CK(dalloc(c, &c->d_order, c->slots_cap*sizeof(int)));
   }
   if ((size_t)num_active*c->p.num_vars > c->partials_cap) {
      CU(cudaStreamSynchronize(c->stream));
      CK(dfree(c, c->d_partials));
      c->partials_cap = c->slots_cap*c->p.num_vars;
      CK(dalloc(c, &c->d_partials, c->partials_cap*sizeof(double)));
   }
   if (c->partials_cap != c->cspart_cap) {
      // the fused kernels' check_sum partials: CS_WARPS slots per tile-variable, zero where
      // a kernel has fewer compute warps
      CU(cudaStreamSynchronize(c->stream));
      CK(dfree(c, c->d_cspart));
      c->d_cspart = nullptr;
      c->cspart_cap = c->partials_cap;
      CK(dalloc(c, &c->d_cspart, c->cspart_cap*CS_WARPS*sizeof(double)));
      CU(cudaMemsetAsync(c->d_cspart, 0, c->cspart_cap*CS_WARPS*sizeof(double), c->stream));
   }
   std::vector<int> slots(num_active);
   for (int a = 0; a < num_active; a++) slots[a] = sorted_blocks[a].slot;
   CU(cudaStreamSynchronize(c->stream));
   const std::vector<int> order = processing_order(c);
   c->h_order = order;
   if (num_active) {
      CU(cudaMemcpyAsync(c->d_slots, slots.data(), num_active*sizeof(int),
                         cudaMemcpyHostToDevice, c->stream));
      CU(cudaMemcpyAsync(c->d_order, order.data(), num_active*sizeof(int),
                         cudaMemcpyHostToDevice, c->stream));
   }
   CU(cudaStreamSynchronize(c->stream));
   c->ops_dirty = true;
   for (int o = 0; o < 6; o++) c->plan_built[o] = false;
   touch_all(c, true);
   return MAMR_OK;
}

int mamr_set_comm_lists(mamr_ctx *c, const mamr_comm_dir dirs[3])
{
   if (!c || !dirs) return fail(MAMR_EINVAL, "null argument");
   CK(settle_all(c));
   CK(wait_xchg(c));
   c->p2p_epoch++;      // every rank calls this at the same program point (p2p.cu: rbase table)
   for (int o = 0; o < 6; o++) c->plan_built[o] = false;
   for (int d = 0; d < 3; d++) {
      const mamr_comm_dir &s = dirs[d];
      DirLists &L = c->cl[d];
      if (s.num_partners < 0 || s.num_cases < 0) return fail(MAMR_EINVAL, "negative list length");
      L.partner.assign(s.partner, s.partner + s.num_partners);
      L.index.assign(s.index, s.index + s.num_partners);
      L.num.assign(s.num, s.num + s.num_partners);
      L.send_size.assign(s.send_size, s.send_size + s.num_partners);
      L.recv_size.assign(s.recv_size, s.recv_size + s.num_partners);
      L.block.assign(s.block, s.block + s.num_cases);
      L.face_case.assign(s.face_case, s.face_case + s.num_cases);
      L.send_off.assign(s.send_off, s.send_off + s.num_cases);
      L.recv_off.assign(s.recv_off, s.recv_off + s.num_cases);
      for (int i = 0; i < s.num_partners; i++) {
         if (L.partner[i] < 0 || L.partner[i] >= c->p.num_ranks || L.partner[i] == c->p.rank)
            return fail(MAMR_EINVAL, "dir %d partner %d is not a valid peer rank", d, L.partner[i]);
         if (L.index[i] < 0 || L.index[i] + L.num[i] > s.num_cases)
            return fail(MAMR_EINVAL, "dir %d partner %d face range out of bounds", d, i);
      }
      for (int f = 0; f < s.num_cases; f++)
         if (L.block[f] < 0 || L.block[f] >= c->p.max_blocks)
            return fail(MAMR_EINVAL, "dir %d face %d slot out of range", d, f);
   }
   c->ops_dirty = true;
   c->have_partners = false;
   for (int d = 0; d < 3; d++)
      if (!c->cl[d].partner.empty()) c->have_partners = true;
   return MAMR_OK;
}

// ---- the stage hot path ----------------------------------------------------
int mamr_comm(mamr_ctx *c, int start, int num_comm, int stage)
{
   if (!c) return fail(MAMR_EINVAL, "null context");
   if (start < 0 || num_comm < 0 || start + num_comm > c->p.num_vars)
      return fail(MAMR_EINVAL, "comm: bad variable range [%d,%d)", start, start + num_comm);
   // the message buffers hold comm_vars variables per face (comm_util.c:50-75, driver.c:75-89)
   if (c->have_partners && num_comm > c->comm_vars)
      return fail(MAMR_EINVAL, "comm: %d variables in one call, but the message buffers are sized for "
                  "--comm_vars %d", num_comm, c->comm_vars);
   // a queued stencil or an earlier deferred comm() of these variables comes first
   CK(settle_data(c, start, num_comm));
   if (c->ops_dirty) CK(build_ops(c));   // also validates the topology (comm.c:198-201)
   const int ord = order_index(c, stage);
   const int set = (start/c->comm_vars)%c->nsets;
   bool defer = false;
   CK(fused_ready(c, ord, &defer));
   // Ghost cells an eliding launch left stale: this exchange overwrites every ghost
   // cell it does not leave alone, and when its plan reads no stored ghost cell the
   // old values are dead -- drop them.  Otherwise they become real first.
   if (defer && c->plan[ord].elidable) {
      for (int v = start; v < start + num_comm; v++) c->stale[v] = 0;
   } else
      CK(regen_ghosts(c, start, num_comm));
   if (defer) {
      // the fused kernel performs this exchange when the stencil of the variables
      // is launched; anything else that needs the ghost cells materialises it
      if (!c->have_partners && num_comm > 0) CK(p2p_begin(c, c->stream, set));   // exchange numbers stay aligned
      if (c->have_partners && num_comm > 0) {
         // this group's receive buffers are about to be reused: whatever else still
         // reads them (another group mapped to the same set) becomes real first
         CK(release_recv_set(c, set));
         // off-rank faces: per phase, fill the send buffers from resolved origins
         // (pack_face, comm.c:254-401) and exchange them (comm.c:71-84, 120-151)
         double *send[3] = { c->d_send[0], c->d_send[1], c->d_send[2] };
         const double *recv[3] = { c->d_recvs[set][0], c->d_recvs[set][1], c->d_recvs[set][2] };
         CK(wait_xchg(c));
         cudaStream_t xs = c->stream;
         if (c->use_overlap) {
            // the exchange reads the variables as the main stream leaves them now and
            // reuses buffers the main stream may still be reading
            xs = c->xstream;
            CU(cudaEventRecord(c->ev_data, c->stream));
            CU(cudaStreamWaitEvent(xs, c->ev_data, 0));
            trace_mark(c, 0, c->stream);
         }
         CK(p2p_begin(c, xs, set));
         auto pack_phase = [&](int o) {
            for (const Run &r : runs_of(c, start, num_comm, false)) {
               KTimer t(c, KC_GHOST, xs, KD_PACK);
               launch_facepack(c->d_pack[ord][o], c->d_pack_fb[ord][o], (int)c->pack_fb[ord][o].size() - 1,
                               vpool(c, r.start), c->g.var_stride, send, recv, r.start, r.num, start, xs);
               c->cnt.kernel_launches++;
            }
         };
         if (c->p.stencil == 7) {
            // The 7-point exchange never widens a face (comm.c:266-270): no phase forwards what an
            // earlier one delivered, so the three directions travel at once -- all packs, all
            // transfers, one wait -- instead of three dependent rounds.
            for (int o = 0; o < 3; o++)
               if (!c->cl[kPerm[ord][o]].partner.empty()) pack_phase(o);
            KTimer t(c, KC_GHOST, xs, KD_XCHG);
            CK(exchange_dirs(c, kPerm[ord], xs, set));
         } else
            for (int o = 0; o < 3; o++) {
               const int d = kPerm[ord][o];
               if (c->cl[d].partner.empty()) continue;
               pack_phase(o);
               KTimer t(c, KC_GHOST, xs, KD_XCHG);
               CK(exchange_dir(c, d, xs, set));
            }
         if (c->use_overlap) {
            CU(cudaEventRecord(c->ev_xchg, xs));
            c->xchg_pending = true;
            trace_mark(c, 1, xs);
         }
         CU(cudaGetLastError());
      }
      for (int v = start; v < start + num_comm; v++) {
         c->pc_ord[v] = (signed char)ord;
         c->pc_start[v] = start;
         c->pc_set[v] = (unsigned char)set;
      }
   } else {
      if (c->have_partners && num_comm > 0) CK(release_recv_set(c, set));
      if (num_comm > 0) CK(comm_split(c, start, num_comm, ord, start, true, set));
   }
   for (int d = 0; d < 3; d++) {
      c->cnt.counter_same[d] += c->n_same[d];
      c->cnt.counter_diff[d] += c->n_diff[d];
      c->cnt.counter_bc[d] += c->n_bc[d];
   }
   return MAMR_OK;
}

static int stencil_vars_stage(mamr_ctx *c, int var_start, int num, int calc_stage)
{
   if (!c) return fail(MAMR_EINVAL, "null context");
   if (var_start < 0 || num < 0 || var_start + num > c->p.num_vars)
      return fail(MAMR_EINVAL, "stencil: bad variable range [%d,%d)", var_start, var_start + num);
   if (num == 0) return MAMR_OK;
   // defer: consecutive variables are merged into one launch (driver.c:85-86
   // calls the stencil once per variable)
   if (c->pend_num > 0 && var_start == c->pend_start + c->pend_num &&
       (c->p.stencil != 0 || calc_stage == c->pend_stage))
      c->pend_num += num;
   else {
      CK(flush_pending(c));
      c->pend_start = var_start;
      c->pend_num = num;
      c->pend_stage = calc_stage;
   }
   for (int v = var_start; v < var_start + num; v++) c->cs_valid[v] = c->cs_fused[v] = c->cs_local_dirty[v] = 0;
   c->modified_since_cs = true;
   const double cells = (double)c->num_active*c->p.nx*c->p.ny*c->p.nz;
   if (c->p.stencil != 0) {
      // stencil.c:100-101, 142-143
      c->cnt.total_fp_divs += cells*num;
      c->cnt.total_fp_adds += (c->p.stencil == 7 ? 6.0 : 26.0)*cells*num;
   } else if (c->s0_set)
      for (int v = var_start; v < var_start + num; v++) {
         if (v == 0 || v >= 4*c->s0_mat) {
            c->cnt.total_fp_divs += cells;
            c->cnt.total_fp_adds += 6.0*cells;
         } else {
            const S0Flops f = s0_flops(calc_stage%6, v, c->s0_mat);
            c->cnt.total_fp_adds += f.adds*cells;
            c->cnt.total_fp_muls += f.muls*cells;
            c->cnt.total_fp_divs += f.divs*cells;
         }
      }
   return MAMR_OK;
}

int mamr_stencil_vars(mamr_ctx *c, int var_start, int num)
{
   if (c && c->p.stencil == 0)
      return fail(MAMR_EINVAL, "--stencil 0: the update depends on the stage, use mamr_stencil_driver / mamr_stage");
   return stencil_vars_stage(c, var_start, num, 0);
}

// stencil_driver(var, calc_stage), stencil.c:43-74: calc_stage only matters for
// --stencil 0, where calc_stage % 6 selects the update kind (:51-69)
int mamr_stencil_driver(mamr_ctx *c, int var, int calc_stage)
{
   return stencil_vars_stage(c, var, 1, calc_stage);
}

int mamr_set_stencil0(mamr_ctx *c, int mat, double a1, const double *a0)
{
   if (!c || !a0) return fail(MAMR_EINVAL, "null argument");
   if (c->p.stencil != 0) return fail(MAMR_EINVAL, "set_stencil0: the context was not created with --stencil 0");
   if (mat != c->p.num_vars/4) return fail(MAMR_EINVAL, "set_stencil0: mat must be num_vars/4 (init.c:419)");
   CK(flush_pending(c));
   c->s0_mat = mat;
   c->s0_a1 = a1;
   c->s0_a0.assign(a0, a0 + mat);
   CU(cudaMemcpyAsync(c->d_s0_a0, c->s0_a0.data(), (size_t)mat*sizeof(double), cudaMemcpyHostToDevice,
                      c->stream));
   CU(cudaStreamSynchronize(c->stream));
   c->s0_set = true;
   return MAMR_OK;
}

int mamr_stencil_calc(mamr_ctx *c, int var) { return stencil_vars_stage(c, var, 1, 0); }

int mamr_check_sum_vars(mamr_ctx *c, int var_start, int num, double *sums)
{
   if (!c || !sums) return fail(MAMR_EINVAL, "null argument");
   if (var_start < 0 || num <= 0 || var_start + num > c->p.num_vars)
      return fail(MAMR_EINVAL, "check_sum: bad variable range [%d,%d)", var_start, var_start + num);
   CK(flush_pending(c));
   c->launches_since_cs = 0;
   for (int v = var_start; v < var_start + num;) {
      // variables whose last writer was a fused stage kernel: their partial sums are
      // already in memory, only the fold is left
      const bool fz = c->cs_fused[v] && c->num_active > 0;
      int e = v + 1;
      while (e < var_start + num && (c->cs_fused[e] && c->num_active > 0) == fz) e++;
      KTimer t(c, KC_CHECKSUM, nullptr, KD_CS);
      if (fz) {
         launch_checksum_final(c->d_cspart + (size_t)v*c->num_active*CS_WARPS, c->num_active*CS_WARPS,
                               e - v, c->d_sums + (v - var_start), c->stream);
         c->cnt.kernel_launches++;
      } else
         for (const Run &r : runs_of(c, v, e - v, false)) {
            launch_checksum(vpool(c, r.start), c->g, c->d_slots, c->num_active, r.start, r.num,
                            c->d_partials, c->d_sums + (r.start - var_start), c->stream);
            c->cnt.kernel_launches += c->num_active > 0 ? 2 : 1;
         }
      v = e;
   }
   // look-ahead results right behind the range: fold their partials in the same round trip
   // (one rank only: the number of values in the all-reduce must not depend on a rank's state)
   int extra = 0;
   if (c->p.num_ranks == 1 && c->num_active > 0) {
      int u = var_start + num;
      while (u < c->p.num_vars && c->spec[u] && (c->spec_flags[u] & 2) && !c->spec_cs_ok[u]) u++;
      extra = u - (var_start + num);
      if (extra > 0) {
         KTimer t(c, KC_CHECKSUM, nullptr, KD_CS);
         launch_checksum_final(c->d_cspart + (size_t)(var_start + num)*c->num_active*CS_WARPS,
                               c->num_active*CS_WARPS, extra, c->d_sums + num, c->stream);
         c->cnt.kernel_launches++;
      }
   }
   CU(cudaGetLastError());
   if (c->p.num_ranks > 1) {      // check_sum.c:57
      if (c->p2p) {
         KTimer t(c, KC_CHECKSUM, nullptr, KD_ALLRED);
         launch_p2p_allreduce(c->d_sums, num, c->d_peer_win, c->win, c->p.rank, c->p.num_ranks,
                              c->p.num_vars, ++c->cs_seq, c->stream);
         c->cnt.kernel_launches++;
      } else {
         if (!c->nccl)
            return fail(MAMR_ENCCL, "check_sum: neither mamr_p2p_connect nor mamr_nccl_init was called");
         NC(g_nccl.AllReduce(c->d_sums, c->d_sums, (size_t)num, NCCL_DOUBLE, NCCL_SUM, c->nccl, c->stream));
      }
   }
   CU(cudaMemcpyAsync(c->h_sums, c->d_sums, (num + extra)*sizeof(double), cudaMemcpyDeviceToHost,
                      c->stream));
   CU(cudaStreamSynchronize(c->stream));
   if (c->p.num_ranks > 1) CK(p2p_check(c));
   CK(fold_s0_checks(c));
   for (int i = 0; i < extra; i++) {
      c->spec_cs[var_start + num + i] = c->h_sums[num + i];
      c->spec_cs_ok[var_start + num + i] = 1;
   }
   for (int i = 0; i < num; i++) {
      sums[i] = c->h_sums[i];
      c->cs_cache[var_start + i] = c->h_sums[i];
      c->cs_valid[var_start + i] = 1;
      c->cs_local_dirty[var_start + i] = 0;
   }
   c->modified_since_cs = false;
   return MAMR_OK;
}

int mamr_check_sum(mamr_ctx *c, int var, double *sum)
{
   if (!c || !sum) return fail(MAMR_EINVAL, "null argument");
   if (var < 0 || var >= c->p.num_vars) return fail(MAMR_EINVAL, "check_sum: bad var %d", var);
   c->cnt.total_red++;   // check_sum.c:62
   const bool had_pending = c->pend_num > 0;
   CK(flush_pending(c));    // may commit a look-ahead result whose check_sum is known already
   if (c->cs_valid[var]) {
      if (c->cs_local_dirty[var])
         return fail(MAMR_EINVAL, "check_sum(%d): this rank's blocks changed (upload / split / consolidate / "
                     "unpack) since the sum was reduced and no collective event (stencil, mamr_set_topology, "
                     "mamr_flush_block_moves) followed; the cached sum is stale and the all-reduce cannot be "
                     "joined by one rank alone", var);
      *sum = c->cs_cache[var];
      return MAMR_OK;
   }
   // init.c:681-682 asks for every variable back to back with no update in
   // between: after the first such call compute the rest in one launch
   int num = 1;
   if (!c->modified_since_cs && !had_pending)
      while (var + num < c->p.num_vars && !c->cs_valid[var + num]) num++;
   std::vector<double> tmp(num);
   CK(mamr_check_sum_vars(c, var, num, tmp.data()));
   *sum = tmp[0];
   return MAMR_OK;
}

int mamr_stage(mamr_ctx *c, int stage)
{
   if (!c) return fail(MAMR_EINVAL, "null context");
   for (int start = 0; start < c->p.num_vars; start += c->comm_vars) {   // driver.c:75-89
      const int number = std::min(c->comm_vars, c->p.num_vars - start);
      CK(mamr_comm(c, start, number, stage));
      CK(stencil_vars_stage(c, start, number, stage));
   }
   return MAMR_OK;
}

// ---- refinement / migration ------------------------------------------------
static int push_rop(mamr_ctx *c, const RefineOp &op, RefineOp **d_out)
{
   if (c->rops_pos == c->rops_cap) {
      CU(cudaStreamSynchronize(c->stream));
      c->rops_pos = 0;
   }
   *d_out = c->d_rops + c->rops_pos++;
   CU(cudaMemcpyAsync(*d_out, &op, sizeof op, cudaMemcpyHostToDevice, c->stream));
   return MAMR_OK;
}

int mamr_split_block(mamr_ctx *c, int parent_slot, const int child_slots[8])
{
   CK(check_slot(c, parent_slot));
   RefineOp op;
   op.parent = parent_slot;
   for (int o = 0; o < 8; o++) {
      CK(check_slot(c, child_slots[o]));
      if (child_slots[o] == parent_slot) return fail(MAMR_EINVAL, "split: child slot equals parent");
      op.child[o] = child_slots[o];
   }
   CK(settle_all(c));
   RefineOp *d;
   CK(push_rop(c, op, &d));
   for (const Run &r : runs_of(c, 0, c->p.num_vars, false)) {
      launch_split(vpool(c, r.start), c->g, d, 1, r.start, r.num, c->stream);
      c->cnt.kernel_launches++;
   }
   CU(cudaGetLastError());
   touch_all(c);
   return MAMR_OK;
}

int mamr_consolidate_block(mamr_ctx *c, const int child_slots[8], int parent_slot)
{
   CK(check_slot(c, parent_slot));
   RefineOp op;
   op.parent = parent_slot;
   for (int o = 0; o < 8; o++) {
      CK(check_slot(c, child_slots[o]));
      if (child_slots[o] == parent_slot) return fail(MAMR_EINVAL, "consolidate: child slot equals parent");
      op.child[o] = child_slots[o];
   }
   CK(settle_all(c));
   RefineOp *d;
   CK(push_rop(c, op, &d));
   for (const Run &r : runs_of(c, 0, c->p.num_vars, false)) {
      launch_consolidate(vpool(c, r.start), c->g, d, 1, r.start, r.num, c->stream);
      c->cnt.kernel_launches++;
   }
   CU(cudaGetLastError());
   touch_all(c);
   return MAMR_OK;
}

int mamr_pack_block(mamr_ctx *c, int slot, double *payload)
{
   CK(check_slot(c, slot));
   if (!payload) return fail(MAMR_EINVAL, "null payload");
   CK(flush_pending(c));
   const size_t n = (size_t)c->p.num_vars*c->p.nx*c->p.ny*c->p.nz;
   for (const Run &r : runs_of(c, 0, c->p.num_vars, false)) {
      launch_pack_block(vpool(c, r.start), c->g, slot, r.start, r.num, c->d_payload, c->stream);
      c->cnt.kernel_launches++;
   }
   CU(cudaMemcpyAsync(payload, c->d_payload, n*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   c->cnt.migrate_bytes += (double)n*sizeof(double);
   return MAMR_OK;
}

int mamr_unpack_block(mamr_ctx *c, int slot, const double *payload)
{
   CK(check_slot(c, slot));
   if (!payload) return fail(MAMR_EINVAL, "null payload");
   CK(settle_all(c));
   const size_t n = (size_t)c->p.num_vars*c->p.nx*c->p.ny*c->p.nz;
   CU(cudaMemcpyAsync(c->d_payload, payload, n*sizeof(double), cudaMemcpyHostToDevice, c->stream));
   for (const Run &r : runs_of(c, 0, c->p.num_vars, false)) {
      launch_unpack_block(vpool(c, r.start), c->g, slot, r.start, r.num, c->d_payload, c->stream);
      c->cnt.kernel_launches++;
   }
   CU(cudaStreamSynchronize(c->stream));
   touch_all(c);
   return MAMR_OK;
}

int mamr_send_block(mamr_ctx *c, int slot, int dest_rank)
{
   CK(check_slot(c, slot));
   if (!c->nccl) return fail(MAMR_ENCCL, "send_block: mamr_nccl_init was not called");
   if (dest_rank < 0 || dest_rank >= c->p.num_ranks || dest_rank == c->p.rank)
      return fail(MAMR_EINVAL, "send_block: bad destination rank %d", dest_rank);
   CK(flush_pending(c));
   const size_t n = (size_t)c->p.num_vars*c->p.nx*c->p.ny*c->p.nz;
   for (const Run &r : runs_of(c, 0, c->p.num_vars, false)) {
      launch_pack_block(vpool(c, r.start), c->g, slot, r.start, r.num, c->d_payload, c->stream);
      c->cnt.kernel_launches++;
   }
   NC(g_nccl.Send(c->d_payload, n, NCCL_DOUBLE, dest_rank, c->nccl, c->stream));
   c->cnt.migrate_bytes += (double)n*sizeof(double);
   CU(cudaGetLastError());
   return MAMR_OK;
}

int mamr_recv_block(mamr_ctx *c, int slot, int src_rank)
{
   CK(check_slot(c, slot));
   if (!c->nccl) return fail(MAMR_ENCCL, "recv_block: mamr_nccl_init was not called");
   if (src_rank < 0 || src_rank >= c->p.num_ranks || src_rank == c->p.rank)
      return fail(MAMR_EINVAL, "recv_block: bad source rank %d", src_rank);
   CK(settle_all(c));
   const size_t n = (size_t)c->p.num_vars*c->p.nx*c->p.ny*c->p.nz;
   NC(g_nccl.Recv(c->d_payload, n, NCCL_DOUBLE, src_rank, c->nccl, c->stream));
   for (const Run &r : runs_of(c, 0, c->p.num_vars, false)) {
      launch_unpack_block(vpool(c, r.start), c->g, slot, r.start, r.num, c->d_payload, c->stream);
      c->cnt.kernel_launches++;
   }
   CU(cudaGetLastError());
   touch_all(c);
   return MAMR_OK;
}

// ---- staged migration -------------------------------------------------------
// The reference moves blocks one at a time through a blocking pairwise handshake
// (rcb.c:207-337).  Issuing one ncclSend/ncclRecv per pack_block()/unpack_block()
// at those call sites can deadlock on the device (two ranks that each send before
// they receive), so the drop-in splits the move: the sender packs the payload into
// a staging area when the host says so (the slot is reused right afterwards,
// rcb.c:259-266), the receiver only records (slot, source), and every rank later
// calls mamr_flush_block_moves() at the same point of the program -- one NCCL
// group holding all sends and receives, then the unpack kernels.  Per pair of
// ranks the k-th staged send matches the k-th staged receive.
namespace {
int grow_stage(mamr_ctx *c, double **buf, size_t *cap, size_t need, size_t keep)
{
   if (need <= *cap) return MAMR_OK;
   const size_t n = (size_t)c->p.num_vars*c->p.nx*c->p.ny*c->p.nz;
   size_t ncap = *cap ? *cap : 16;
   while (ncap < need) ncap *= 2;
   double *nb = nullptr;
   CU(cudaMalloc(&nb, ncap*n*sizeof(double)));
   if (*buf && keep)
      CU(cudaMemcpyAsync(nb, *buf, keep*n*sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   CK(dfree(c, *buf));
   *buf = nb;
   *cap = ncap;
   return MAMR_OK;
}
}  // namespace

int mamr_stage_send_block(mamr_ctx *c, int slot, int dest_rank)
{
   CK(check_slot(c, slot));
   if (!c->nccl && !c->p2p)
      return fail(MAMR_ENCCL, "stage_send_block: neither mamr_p2p_connect nor mamr_nccl_init was called");
   if (dest_rank < 0 || dest_rank >= c->p.num_ranks || dest_rank == c->p.rank)
      return fail(MAMR_EINVAL, "stage_send_block: bad destination rank %d", dest_rank);
   for (const mamr_ctx::MoveRec &r : c->mv_recv)
      if (r.slot == slot)
         return fail(MAMR_EINVAL, "stage_send_block: slot %d still waits for its own payload", slot);
   CK(flush_pending(c));
   const size_t n = (size_t)c->p.num_vars*c->p.nx*c->p.ny*c->p.nz;
   double *dst;
   if (c->p2p) {
      // peer-memory transport: the payload waits in this rank's window for the receiver to fetch it
      if ((int)c->mv_send.size() >= c->mv_cap_p2p)
         return fail(MAMR_EP2P, "more than %d blocks leave this rank in one load-balance step: raise "
                     "MAMR_P2P_MOVE_BLOCKS", c->mv_cap_p2p);
      dst = reinterpret_cast<double *>(c->win + c->win_mv_off + p2p_mv_payload_offset(c->mv_cap_p2p)) +
            c->mv_send.size()*n;
   } else {
      CK(grow_stage(c, &c->d_mv_send, &c->mv_send_cap, c->mv_send.size() + 1, c->mv_send.size()));
      dst = c->d_mv_send + c->mv_send.size()*n;
   }
   for (const Run &r : runs_of(c, 0, c->p.num_vars, false)) {
      launch_pack_block(vpool(c, r.start), c->g, slot, r.start, r.num, dst, c->stream);
      c->cnt.kernel_launches++;
   }
   CU(cudaGetLastError());
   c->mv_send.push_back({ slot, dest_rank });
   return MAMR_OK;
}

int mamr_stage_recv_block(mamr_ctx *c, int slot, int src_rank)
{
   CK(check_slot(c, slot));
   if (!c->nccl && !c->p2p)
      return fail(MAMR_ENCCL, "stage_recv_block: neither mamr_p2p_connect nor mamr_nccl_init was called");
   if (src_rank < 0 || src_rank >= c->p.num_ranks || src_rank == c->p.rank)
      return fail(MAMR_EINVAL, "stage_recv_block: bad source rank %d", src_rank);
   for (const mamr_ctx::MoveRec &r : c->mv_recv)
      if (r.slot == slot)
         return fail(MAMR_EINVAL, "stage_recv_block: slot %d already waits for a payload", slot);
   c->mv_recv.push_back({ slot, src_rank });
   return MAMR_OK;
}

int mamr_pending_block_moves(mamr_ctx *c)
{
   return c ? (int)(c->mv_send.size() + c->mv_recv.size()) : 0;
}

int mamr_flush_block_moves(mamr_ctx *c)
{
   if (!c) return fail(MAMR_EINVAL, "null context");
   // every rank calls this at the same program point, also ranks that move nothing: a
   // collectively uniform invalidation of the check_sum cache (see touch_all)
   if (c->p.num_ranks > 1) {
      std::fill(c->cs_valid.begin(), c->cs_valid.end(), 0);
      std::fill(c->cs_local_dirty.begin(), c->cs_local_dirty.end(), 0);
      c->modified_since_cs = true;
   }
   if (c->p2p) return p2p_flush_moves(c);
   if (c->mv_send.empty() && c->mv_recv.empty()) return MAMR_OK;
   if (!c->nccl) return fail(MAMR_ENCCL, "flush_block_moves: mamr_nccl_init was not called");
   CK(settle_all(c));
   const size_t n = (size_t)c->p.num_vars*c->p.nx*c->p.ny*c->p.nz;
   CK(grow_stage(c, &c->d_mv_recv, &c->mv_recv_cap, c->mv_recv.size(), 0));
   NC(g_nccl.GroupStart());
   for (size_t i = 0; i < c->mv_send.size(); i++)
      NC(g_nccl.Send(c->d_mv_send + i*n, n, NCCL_DOUBLE, c->mv_send[i].peer, c->nccl, c->stream));
   for (size_t i = 0; i < c->mv_recv.size(); i++)
      NC(g_nccl.Recv(c->d_mv_recv + i*n, n, NCCL_DOUBLE, c->mv_recv[i].peer, c->nccl, c->stream));
   NC(g_nccl.GroupEnd());
   for (size_t i = 0; i < c->mv_recv.size(); i++)
      for (const Run &r : runs_of(c, 0, c->p.num_vars, false)) {
         launch_unpack_block(vpool(c, r.start), c->g, c->mv_recv[i].slot, r.start, r.num,
                             c->d_mv_recv + i*n, c->stream);
         c->cnt.kernel_launches++;
      }
   CU(cudaGetLastError());
   c->cnt.migrate_bytes += (double)c->mv_send.size()*n*sizeof(double);
   if (!c->mv_recv.empty()) touch_all(c, true);
   c->mv_send.clear();
   c->mv_recv.clear();
   return MAMR_OK;
}

int mamr_set_message_mode(mamr_ctx *c, int send_faces, const int *msg_len)
{
   if (!c || (send_faces && !msg_len)) return fail(MAMR_EINVAL, "null argument");
   c->send_faces = send_faces != 0;
   if (msg_len)
      for (int d = 0; d < 3; d++)
         for (int k = 0; k < 4; k++) c->msg_len[d][k] = msg_len[d*4 + k];
   return MAMR_OK;
}

int mamr_device_count(void)
{
   int n = 0;
   return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

// ---- multi-GPU -------------------------------------------------------------
int mamr_nccl_get_unique_id(char id[MAMR_NCCL_ID_BYTES])
{
   if (!id) return fail(MAMR_EINVAL, "null id");
   if (!load_nccl()) return fail(MAMR_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
   ncclUniqueId u;
   NC(g_nccl.GetUniqueId(&u));
   memcpy(id, u.internal, MAMR_NCCL_ID_BYTES);
   return MAMR_OK;
}

int mamr_nccl_init(mamr_ctx *c, const char id[MAMR_NCCL_ID_BYTES])
{
   if (!c || !id) return fail(MAMR_EINVAL, "null argument");
   if (!load_nccl()) return fail(MAMR_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
   ncclUniqueId u;
   memcpy(u.internal, id, MAMR_NCCL_ID_BYTES);
   NC(g_nccl.CommInitRank(&c->nccl, c->p.num_ranks, u, c->p.rank));
   return MAMR_OK;
}

// ---- peer-memory transport (p2p.cu) -------------------------------------------
namespace {
struct P2PBlob {            // what travels over the host channel, MAMR_P2P_HANDLE_BYTES
   unsigned long long magic;
   long long pid;
   unsigned long long ptr;          // the window's address in the owner's process
   unsigned long long bytes;
   int device, num_vars;
   cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(P2PBlob) <= MAMR_P2P_HANDLE_BYTES, "handle blob too big");
constexpr unsigned long long P2P_MAGIC = 0x6d616d7270327031ULL;
}  // namespace

int mamr_p2p_get_handle(mamr_ctx *c, char handle[MAMR_P2P_HANDLE_BYTES])
{
   if (!c || !handle) return fail(MAMR_EINVAL, "null argument");
   if (c->p.num_ranks > P2P_MAX_RANKS)
      return fail(MAMR_EUNSUPPORTED, "peer-memory transport: at most %d ranks (one node)", P2P_MAX_RANKS);
   if (!c->win) {
      // window = header + check_sum slots + receive buffers of every set.  Default capacity:
      // a sixteenth of one block pool (the ghost layer of a sub-cube of b^3 blocks of n^3 cells
      // is 6/(b n) of it per variable group), at least 64 MB; MAMR_P2P_WINDOW_MB overrides.
      size_t data = std::max<size_t>((size_t)64 << 20, c->pool_bytes/16);
      if (const char *e = getenv("MAMR_P2P_WINDOW_MB")) data = (size_t)atoll(e) << 20;
      c->win_data_off = p2p_data_offset(c->p.num_vars);
      c->win_data_cap = data/sizeof(double);
      // migration staging: room for half of the pool's slots, at most 2 GB (MAMR_P2P_MOVE_BLOCKS)
      const size_t pay = (size_t)c->p.num_vars*c->p.nx*c->p.ny*c->p.nz*sizeof(double);
      size_t mvb = std::max<size_t>(8, std::min<size_t>((size_t)c->p.max_blocks/2, ((size_t)2 << 30)/pay));
      if (const char *e = getenv("MAMR_P2P_MOVE_BLOCKS")) mvb = (size_t)std::max(1, atoi(e));
      c->mv_cap_p2p = (int)std::min<size_t>(mvb, (size_t)c->p.max_blocks);
      c->win_mv_off = (c->win_data_off + data + 255)/256*256;
      c->win_bytes = c->win_mv_off + p2p_mv_payload_offset(c->mv_cap_p2p) + (size_t)c->mv_cap_p2p*pay;
      CU(cudaMalloc(&c->win, c->win_bytes));
      CU(cudaMemsetAsync(c->win, 0, c->win_bytes, c->stream));
      CU(cudaStreamSynchronize(c->stream));
   }
   P2PBlob b;
   memset(&b, 0, sizeof b);
   b.magic = P2P_MAGIC;
   b.pid = (long long)getpid();
   b.ptr = (unsigned long long)(uintptr_t)c->win;
   b.bytes = c->win_bytes;
   b.num_vars = c->p.num_vars;
   CU(cudaGetDevice(&b.device));
   CU(cudaIpcGetMemHandle(&b.ipc, c->win));
   memset(handle, 0, MAMR_P2P_HANDLE_BYTES);
   memcpy(handle, &b, sizeof b);
   return MAMR_OK;
}

int mamr_p2p_connect(mamr_ctx *c, const char *handles)
{
   if (!c || !handles) return fail(MAMR_EINVAL, "null argument");
   if (!c->win) return fail(MAMR_EINVAL, "p2p_connect: call mamr_p2p_get_handle first");
   if (c->p2p) return fail(MAMR_EINVAL, "p2p_connect: already connected");
   const int R = c->p.num_ranks;
   CK(settle_all(c));
   c->peer_win.assign(R, nullptr);
   c->peer_ipc.assign(R, 0);
   for (int r = 0; r < R; r++) {
      P2PBlob b;
      memcpy(&b, handles + (size_t)r*MAMR_P2P_HANDLE_BYTES, sizeof b);
      if (b.magic != P2P_MAGIC || b.num_vars != c->p.num_vars)
         return fail(MAMR_EINVAL, "p2p_connect: handle of rank %d is not a window of this job", r);
      if (r == c->p.rank) {
         if ((char *)(uintptr_t)b.ptr != c->win) return fail(MAMR_EINVAL, "p2p_connect: handle %d is not mine", r);
         c->peer_win[r] = c->win;
      } else if (b.pid == (long long)getpid()) {
         // a rank of this process (loopback): same address space.  Another device needs
         // peer access; the same device needs nothing.
         int dev = 0;
         CU(cudaGetDevice(&dev));
         if (b.device != dev) {
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
               return fail(MAMR_ECUDA, "p2p_connect: no peer access from device %d to %d: %s", dev, b.device,
                           cudaGetErrorString(e));
            cudaGetLastError();
         }
         c->peer_win[r] = (char *)(uintptr_t)b.ptr;
         if (!c->arena) {
            size_t mb = 128;
            if (const char *e = getenv("MAMR_INPROC_ARENA_MB")) mb = (size_t)std::max(1, atoi(e));
            c->arena_cap = mb << 20;
            CU(cudaMalloc(&c->arena, c->arena_cap));
         }
         c->p2p_inproc = true;
      } else {
         void *q = nullptr;
         CU(cudaIpcOpenMemHandle(&q, b.ipc, cudaIpcMemLazyEnablePeerAccess));
         c->peer_win[r] = (char *)q;
         c->peer_ipc[r] = 1;
      }
   }
   CU(cudaMalloc(&c->d_peer_win, (size_t)R*sizeof(char *)));
   CU(cudaMemcpy(c->d_peer_win, c->peer_win.data(), (size_t)R*sizeof(char *), cudaMemcpyHostToDevice));
   for (int d = 0; d < 3; d++) CU(cudaMalloc(&c->d_push[d], (size_t)P2P_MAX_RANKS*sizeof(P2PTarget)));
   CU(cudaMalloc(&c->d_credit, (size_t)3*P2P_MAX_RANKS*sizeof(P2PTarget)));
   CU(cudaMalloc(&c->d_push_done, (size_t)3*P2P_MAX_RANKS*sizeof(unsigned)));
   CU(cudaMemset(c->d_push_done, 0, (size_t)3*P2P_MAX_RANKS*sizeof(unsigned)));
   CU(cudaMallocHost(&c->h_p2p_err, sizeof(unsigned long long)));
   *c->h_p2p_err = 0;
   CU(cudaMalloc(&c->d_mv_ranks, (size_t)2*P2P_MAX_RANKS*sizeof(int)));
   // (sized for every slot: no allocation while a peer may be spinning on this rank)
   c->mv_moves_cap = (size_t)c->p.max_blocks;
   CU(cudaMalloc(&c->d_mv_moves, c->mv_moves_cap*sizeof(P2PMove)));
   CU(cudaMalloc(&c->d_mv_k, c->mv_moves_cap*sizeof(int)));
   CU(cudaMalloc(&c->d_cur, (size_t)c->p.num_vars));
   // receive buffers allocated for another transport move into the window at the next comm()
   for (int d = 0; d < 3; d++) {
      for (int q = 0; q < mamr_ctx::MAX_SETS; q++) {
         if (c->d_recvs[q][d]) CU(cudaFree(c->d_recvs[q][d]));
         c->d_recvs[q][d] = nullptr;
      }
      c->recv_cap[d] = 0;
   }
   c->p2p = true;
   c->p2p_timeout_s = p2p_set_timeout_from_env();
   if (const char *e = getenv("MAMR_TRANSPORT"))
      if (!strcmp(e, "nccl") && c->nccl) c->p2p = false;     // keep NCCL for A/B measurements
   c->ops_dirty = true;
   for (int o = 0; o < 6; o++) c->plan_built[o] = false;
   return MAMR_OK;
}

// ---- host-only plan view ---------------------------------------------------
struct mamr_plan {
   HaloPlan halo;
   std::vector<BoxOp> pack[3];
   int order[3];
};

int mamr_plan_create(const mamr_params *params, int num_active, const mamr_block *sorted_blocks,
                     const mamr_comm_dir dirs[3], int stage, mamr_plan **out)
{
   if (!params || !out || num_active < 0 || (num_active && !sorted_blocks))
      return fail(MAMR_EINVAL, "null argument");
   const mamr_params &p = *params;
   if (p.nx <= 0 || p.ny <= 0 || p.nz <= 0 || ((p.nx | p.ny | p.nz) & 1) || p.max_blocks <= 0)
      return fail(MAMR_EINVAL, "bad geometry");
   Geometry g;
   g.n[0] = p.nx; g.n[1] = p.ny; g.n[2] = p.nz;
   g.str[2] = 1; g.str[1] = p.nz + 2; g.str[0] = (p.ny + 2)*(p.nz + 2);
   g.tile = (p.nx + 2)*g.str[0];
   g.tile_stride = ((long long)g.tile + 15)/16*16;
   g.var_stride = g.tile_stride*p.max_blocks;
   std::vector<mamr_block> blocks(sorted_blocks, sorted_blocks + num_active);
   for (const mamr_block &b : blocks)
      if (b.slot < 0 || b.slot >= p.max_blocks) return fail(MAMR_EINVAL, "slot out of range");
   DirLists cl[3];
   bool partners = false;
   for (int d = 0; d < 3 && dirs; d++) {
      const mamr_comm_dir &s = dirs[d];
      DirLists &L = cl[d];
      L.partner.assign(s.partner, s.partner + s.num_partners);
      L.index.assign(s.index, s.index + s.num_partners);
      L.num.assign(s.num, s.num + s.num_partners);
      L.send_size.assign(s.send_size, s.send_size + s.num_partners);
      L.recv_size.assign(s.recv_size, s.recv_size + s.num_partners);
      L.block.assign(s.block, s.block + s.num_cases);
      L.face_case.assign(s.face_case, s.face_case + s.num_cases);
      L.send_off.assign(s.send_off, s.send_off + s.num_cases);
      L.recv_off.assign(s.recv_off, s.recv_off + s.num_cases);
      if (s.num_partners) partners = true;
   }
   mamr_plan *P = new mamr_plan();
   const int ord = p.permute ? ((stage%6) + 6)%6 : 0;
   PlanInput in;
   in.g = &g; in.stencil = p.stencil; in.max_blocks = p.max_blocks; in.blocks = &blocks; in.cl = cl;
   for (int o = 0; o < 3; o++) in.order[o] = P->order[o] = kPerm[ord][o];
   build_halo_plan(in, P->halo);
   std::string why = P->halo.why;
   bool ok = P->halo.ok;
   std::vector<int> fb;
   for (int o = 0; o < 3 && ok && partners; o++) ok = build_pack_plan(in, o, P->pack[o], fb, why);
   if (!ok) {
      delete P;
      return fail(why.find("misconnected") != std::string::npos ? MAMR_ETOPOLOGY : MAMR_EUNSUPPORTED,
                  "%s", why.c_str());
   }
   *out = P;
   return MAMR_OK;
}

int mamr_plan_phase_dir(mamr_plan *P, int phase)
{
   return (P && phase >= 0 && phase < 3) ? P->order[phase] : -1;
}

static const std::vector<BoxOp> *plan_ops(mamr_plan *P, int which)
{
   if (!P || which < 0 || which > 3) return nullptr;
   return which == 0 ? &P->halo.ops : &P->pack[which - 1];
}

int mamr_plan_num_ops(mamr_plan *P, int which)
{
   const std::vector<BoxOp> *v = plan_ops(P, which);
   return v ? (int)v->size() : -1;
}

int mamr_plan_get_ops(mamr_plan *P, int which, long long *f)
{
   const std::vector<BoxOp> *v = plan_ops(P, which);
   if (!v || !f) return fail(MAMR_EINVAL, "bad plan query");
   for (const BoxOp &o : *v) {
      *f++ = o.dst_base; *f++ = o.src_base; *f++ = o.dst_vs; *f++ = o.src_vs;
      for (int a = 0; a < 3; a++) *f++ = o.ext[a];
      for (int a = 0; a < 3; a++) *f++ = o.dst_str[a];
      for (int a = 0; a < 3; a++) *f++ = o.src_str[a];
      *f++ = o.S; *f++ = o.F; *f++ = o.first; *f++ = o.mode; *f++ = o.dst_mem; *f++ = o.src_mem;
   }
   return MAMR_OK;
}

int mamr_plan_block_begin(mamr_plan *P, int *begin)
{
   if (!P || !begin) return fail(MAMR_EINVAL, "bad plan query");
   for (size_t i = 0; i < P->halo.begin.size(); i++) begin[i] = P->halo.begin[i];
   return MAMR_OK;
}

void mamr_plan_destroy(mamr_plan *P) { delete P; }

// ---- measurement -----------------------------------------------------------
int mamr_timer_begin(mamr_ctx *c)
{
   if (!c) return fail(MAMR_EINVAL, "null context");
   CK(flush_pending(c));
   CU(cudaEventRecord(c->ev_begin, c->stream));
   return MAMR_OK;
}

int mamr_timer_end(mamr_ctx *c, float *ms)
{
   if (!c || !ms) return fail(MAMR_EINVAL, "null argument");
   CK(flush_pending(c));
   CU(cudaEventRecord(c->ev_end, c->stream));
   CU(cudaEventSynchronize(c->ev_end));
   CU(cudaEventElapsedTime(ms, c->ev_begin, c->ev_end));
   return MAMR_OK;
}

int mamr_kernel_timing(mamr_ctx *c, int enable)
{
   if (!c) return fail(MAMR_EINVAL, "null context");
   CK(flush_pending(c));
   CU(cudaStreamSynchronize(c->stream));
   drain_ktimers(c);
   c->ktiming = enable != 0;
   for (int i = 0; i < 3; i++) { c->k_ms[i] = 0; c->k_launches[i] = 0; }
   for (int i = 0; i < KD_NUM; i++) c->kd_ms[i] = 0;
   return MAMR_OK;
}

int mamr_get_device_times(mamr_ctx *c, int wait, mamr_device_times *out)
{
   if (!c || !out) return fail(MAMR_EINVAL, "null argument");
   if (wait) {
      CK(flush_pending(c));
      CU(cudaStreamSynchronize(c->stream));
      if (c->xstream) CU(cudaStreamSynchronize(c->xstream));
   }
   drain_ktimers(c, wait != 0);
   out->fused_ms = c->kd_ms[KD_FUSED];
   out->stencil_ms = c->kd_ms[KD_STENCIL];
   out->split_ghost_ms = c->kd_ms[KD_SPLIT];
   out->pack_ms = c->kd_ms[KD_PACK];
   out->exchange_ms = c->kd_ms[KD_XCHG];
   out->unpack_ms = c->kd_ms[KD_UNPACK];
   out->regen_ms = c->kd_ms[KD_REGEN];
   out->checksum_ms = c->kd_ms[KD_CS];
   out->allreduce_ms = c->kd_ms[KD_ALLRED];
   // share of a fused launch that is the ghost exchange: halo bytes / all bytes
   const double n3 = (double)c->p.nx*c->p.ny*c->p.nz;
   const double H = c->p.stencil == 7 ? 2.0*(c->p.nx*c->p.ny + c->p.ny*c->p.nz + c->p.nx*c->p.nz)
                                      : (double)(c->p.nx + 2)*(c->p.ny + 2)*(c->p.nz + 2) - n3;
   out->halo_fraction = 8.0*H/n3/(16.0 + 8.0*H/n3);
   return MAMR_OK;
}

int mamr_kernel_time_ms(mamr_ctx *c, float *stencil_ms, float *ghost_ms, float *checksum_ms,
                        long long *stencil_launches, long long *ghost_launches,
                        long long *checksum_launches)
{
   if (!c) return fail(MAMR_EINVAL, "null context");
   CK(flush_pending(c));
   CU(cudaStreamSynchronize(c->stream));
   drain_ktimers(c);
   if (stencil_ms) *stencil_ms = (float)c->k_ms[KC_STENCIL];
   if (ghost_ms) *ghost_ms = (float)c->k_ms[KC_GHOST];
   if (checksum_ms) *checksum_ms = (float)c->k_ms[KC_CHECKSUM];
   if (stencil_launches) *stencil_launches = c->k_launches[KC_STENCIL];
   if (ghost_launches) *ghost_launches = c->k_launches[KC_GHOST];
   if (checksum_launches) *checksum_launches = c->k_launches[KC_CHECKSUM];
   return MAMR_OK;
}

}  // extern "C"
