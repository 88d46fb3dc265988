// Halo plan: where does every ghost cell of every active block get its value
// from in one comm() call?  (host code, integers only)
//
// comm() (comm.c:42-242, --code 0) runs three direction phases; phase o with
// direction d writes the two ghost planes normal to d of every block from the
// adjacent interior plane of the neighbour (same level: on_proc_comm
// :1473-1534 or pack/unpack_face cases 0/1 :254-401/:1002-1150), of the block
// itself (apply_bc :1911-1965), or through restriction/prolongation at a level
// boundary (on_proc_comm_diff :1597-1688, cases 2-9).  For stencils other than
// the 7-point one the in-face extent of a phase is widened to 0..n+1 along the
// axes numbered lower than d (boundary faces: along both axes), so a later
// phase forwards ghost cells an earlier phase wrote — that is how edge and
// corner values travel.
//
// Instead of replaying the phases on memory, the fused stage kernel pulls every
// ghost cell straight from its origin.  This file computes the origin: walk the
// phases backwards from the last one that writes the cell; every same-level or
// boundary hop replaces one ghost coordinate by an interior one (n or 1) in the
// neighbouring (or the same) block; the walk ends in an interior cell, in a
// ghost cell no earlier phase of this call wrote (its stored value is read, as
// the reference would), or in a receive buffer.  Level-boundary transfers are
// only supported as the last hop (which is all the 7-point stencil needs; the
// reference itself rejects other stencils on refined meshes, main.c:709-710).
#include <algorithm>
#include <map>

#include "common.cuh"

namespace mamr {
namespace {

enum FaceType { FT_BC, FT_SAME, FT_SAME_OFF, FT_COARSER, FT_FINER, FT_BAD };

struct Ctx {
   const PlanInput &in;
   const Geometry &g;
   std::vector<int> slot2idx;
   // (slot*6 + face) -> comm-list entries of that face, per direction
   std::map<long long, std::vector<int>> faces;
   bool wide;
   std::string err;
   explicit Ctx(const PlanInput &i) : in(i), g(*i.g), wide(i.stencil != 7)
   {
      slot2idx.assign(in.max_blocks, -1);
      const std::vector<mamr_block> &B = *in.blocks;
      for (size_t a = 0; a < B.size(); a++) slot2idx[B[a].slot] = (int)a;
      for (int d = 0; d < 3; d++) {
         const DirLists &L = in.cl[d];
         for (size_t f = 0; f < L.block.size(); f++) {
            const int l = 2*d + (L.face_case[f] >= 10 ? 1 : 0);
            faces[(long long)L.block[f]*6 + l].push_back((int)f);
         }
      }
   }
   const mamr_block &blk(int a) const { return (*in.blocks)[a]; }
};

inline void face_axes(int d, int &sa, int &fa)
{
   sa = (d == 0) ? 1 : 0;
   fa = (d == 2) ? 1 : 2;
}

FaceType face_type(const Ctx &c, int a, int l)
{
   const mamr_block &b = c.blk(a);
   const int nl = b.nei_level[l];
   if (nl == -2) return FT_BC;
   if (nl == b.level) return b.nei[l][0][0] >= 0 ? FT_SAME : FT_SAME_OFF;
   if (nl == b.level - 1) return FT_COARSER;
   if (nl == b.level + 1) return FT_FINER;
   return FT_BAD;
}

inline bool is_ghost(const Geometry &g, int ax, int v) { return v == 0 || v == g.n[ax] + 1; }

// does phase direction d write cell `c` (ghost along d) of block a?
bool covered(const Ctx &c, int a, int d, const int cc[3])
{
   const int l = 2*d + (cc[d] == 0 ? 0 : 1);
   const FaceType t = face_type(c, a, l);
   for (int ax = 0; ax < 3; ax++) {
      if (ax == d || cc[ax] < 0 || !is_ghost(c.g, ax, cc[ax])) continue;
      bool w;
      if (t == FT_BC) w = c.wide;                                   // comm.c:1945-1962
      else if (t == FT_SAME || t == FT_SAME_OFF) w = c.wide && ax < d;   // :1496-1527
      else w = false;                                               // :1618-1631
      if (!w) return false;
   }
   return true;
}

// whole-face extent of in-face axis ax in a message of case 0/1
inline void whole_extent(const Geometry &g, bool wide, int d, int ax, int &lo, int &hi)
{
   if (wide && ax < d) { lo = 0; hi = g.n[ax] + 1; }
   else { lo = 1; hi = g.n[ax]; }
}

inline void quarter_range(const Geometry &g, int fc, int sa, int fa, int &s0, int &s1, int &f0,
                          int &f1)
{
   const int hs = g.n[sa]/2, hf = g.n[fa]/2;
   if (fc%2 == 0) { s0 = 1; s1 = hs; } else { s0 = hs + 1; s1 = g.n[sa]; }
   if ((fc/2)%2 == 1) { f0 = 1; f1 = hf; } else { f0 = hf + 1; f1 = g.n[fa]; }
}

struct Origin {
   int kind;        // 0 pool cell(s), 1 receive buffer of direction rdir
   int slot;
   int cc[3];       // per axis: -1 = runs with the destination index, else fixed
   int rdir, rface;
};

// value of cell cc of block a just before phase `limit` (3 = after the whole call)
bool resolve(Ctx &c, int a, int cc[3], int limit, Origin &out)
{
   for (;;) {
      int o, d = -1;
      for (o = limit - 1; o >= 0; o--) {
         d = c.in.order[o];
         if (cc[d] >= 0 && is_ghost(c.g, d, cc[d]) && covered(c, a, d, cc)) break;
      }
      if (o < 0) {   // interior, or a ghost cell nobody wrote yet: its stored value
         out.kind = 0;
         out.slot = c.blk(a).slot;
         out.cc[0] = cc[0]; out.cc[1] = cc[1]; out.cc[2] = cc[2];
         return true;
      }
      const bool lo = cc[d] == 0;
      const int l = 2*d + (lo ? 0 : 1);
      switch (face_type(c, a, l)) {
      case FT_BC:
         cc[d] = lo ? 1 : c.g.n[d];
         break;
      case FT_SAME: {
         const int m = c.blk(a).nei[l][0][0];
         if (m >= c.in.max_blocks || c.slot2idx[m] < 0) {
            c.err = "ERROR: misconnected block";
            return false;
         }
         a = c.slot2idx[m];
         cc[d] = lo ? c.g.n[d] : 1;
         break;
      }
      case FT_SAME_OFF: {
         auto it = c.faces.find((long long)c.blk(a).slot*6 + l);
         if (it == c.faces.end() || it->second.size() != 1 ||
             c.in.cl[d].face_case[it->second[0]]%10 > 1) {
            c.err = "off-rank same-level face without a whole-face comm-list entry";
            return false;
         }
         out.kind = 1;
         out.slot = c.blk(a).slot;
         out.rdir = d;
         out.rface = it->second[0];
         out.cc[0] = cc[0]; out.cc[1] = cc[1]; out.cc[2] = cc[2];
         return true;
      }
      default:
         c.err = "a ghost value would have to travel through a level boundary and on "
                 "(only the 7-point stencil is supported on refined meshes)";
         return false;
      }
      limit = o;
   }
}

// region descriptor -> destination box inside the tile
struct Box {
   int lo[3], ext[3];
};

inline long long cell_off(const Geometry &g, int i, int j, int k)
{
   return (long long)i*g.str[0] + (long long)j*g.str[1] + k;
}

void init_op(const Geometry &g, BoxOp &op)
{
   op.dst_vs = op.src_vs = g.var_stride;
   for (int ax = 0; ax < 3; ax++) op.dst_str[ax] = op.src_str[ax] = g.str[ax];
   op.S = op.F = 0;
   op.first = 0;
   op.flags = 0;
   op.mode = FM_COPY;
   op.dst_mem = op.src_mem = BM_POOL;
}

// COPY op from a resolved origin; ext[] are the box extents (1 along fixed axes)
void origin_op(const Ctx &c, const Origin &og, const int ext[3], BoxOp &op)
{
   const Geometry &g = c.g;
   if (og.kind == 0) {
      op.src_base = (long long)og.slot*g.tile_stride +
                    cell_off(g, og.cc[0] < 0 ? 1 : og.cc[0], og.cc[1] < 0 ? 1 : og.cc[1],
                             og.cc[2] < 0 ? 1 : og.cc[2]);
      return;
   }
   const int d = og.rdir;
   const DirLists &L = c.in.cl[d];
   int sa, fa, s0, s1, f0, f1;
   face_axes(d, sa, fa);
   const bool w = L.face_case[og.rface]%10 == 1;
   whole_extent(g, w, d, sa, s0, s1);
   whole_extent(g, w, d, fa, f0, f1);
   const int Nf = f1 - f0 + 1, Ns = s1 - s0 + 1;
   const int cs = og.cc[sa] < 0 ? 1 : og.cc[sa], cf = og.cc[fa] < 0 ? 1 : og.cc[fa];
   op.src_mem = (unsigned char)(BM_BUF0 + d);
   op.src_base = (long long)L.recv_off[og.rface] + (long long)(cs - s0)*Nf + (cf - f0);
   op.src_vs = (long long)Ns*Nf;
   op.src_str[d] = 0;
   op.src_str[sa] = Nf;
   op.src_str[fa] = 1;
   (void)ext;
}

// ops of one ghost region r (per axis -1 / 0 / +1) of block a
bool region_ops(Ctx &c, int a, const int r[3], std::vector<BoxOp> &ops)
{
   const Geometry &g = c.g;
   const mamr_block &b = c.blk(a);
   int cc[3];
   BoxOp op;
   init_op(g, op);
   int lo[3];
   for (int ax = 0; ax < 3; ax++) {
      cc[ax] = r[ax] < 0 ? 0 : (r[ax] > 0 ? g.n[ax] + 1 : -1);
      lo[ax] = r[ax] < 0 ? 0 : (r[ax] > 0 ? g.n[ax] + 1 : 1);
      op.ext[ax] = r[ax] ? 1 : g.n[ax];
   }
   op.dst_base = cell_off(g, lo[0], lo[1], lo[2]);
   if ((r[0] != 0) + (r[1] != 0) + (r[2] != 0) == 1) op.flags |= BF_FACE;
   // last phase that writes the region
   int o, d = -1;
   for (o = 2; o >= 0; o--) {
      d = c.in.order[o];
      if (r[d] && covered(c, a, d, cc)) break;
   }
   FaceType t = FT_SAME;
   int l = 0;
   if (o >= 0) {
      l = 2*d + (r[d] < 0 ? 0 : 1);
      t = face_type(c, a, l);
   }
   if (o < 0 || t == FT_BC || t == FT_SAME || t == FT_SAME_OFF) {
      Origin og;
      if (!resolve(c, a, cc, 3, og)) return false;
      origin_op(c, og, op.ext, op);
      if (o < 0) op.flags |= BF_IDENT | BF_GHOST_SRC;
      else if (og.kind == 0)
         for (int ax = 0; ax < 3; ax++)
            if (og.cc[ax] >= 0 && is_ghost(g, ax, og.cc[ax])) op.flags |= BF_GHOST_SRC;
      ops.push_back(op);
      return true;
   }
   if (t == FT_BAD) {
      c.err = "ERROR: misconnected block";
      return false;
   }
   // level boundary: a pure face region (covered() guarantees it)
   int sa, fa;
   face_axes(d, sa, fa);
   const int hs = g.n[sa]/2, hf = g.n[fa]/2;
   const long long S = g.str[sa], F = g.str[fa], N = g.str[d];
   const bool minus = r[d] < 0;
   const int ghost = minus ? 0 : g.n[d] + 1;
   if (t == FT_COARSER) {
      const int m = b.nei[l][0][0];
      if (m >= 0) {
         // coarse neighbour on this rank: value/4 replicated 2x2 (comm.c:1621-1625)
         if (m >= c.in.max_blocks || c.slot2idx[m] < 0) {
            c.err = "ERROR: misconnected block";
            return false;
         }
         const mamr_block &cb = c.blk(c.slot2idx[m]);
         const int k = 2*d + (minus ? 1 : 0);     // the coarse block's face towards me
         int iq = -1, jq = -1;
         for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++)
               if (cb.nei[k][i][j] == b.slot) { iq = i; jq = j; }
         if (iq < 0) {
            c.err = "ERROR: misconnected block";
            return false;
         }
         const int c_src = minus ? g.n[d] : 1;
         op.mode = FM_PROLONG;
         op.src_base = (long long)m*g.tile_stride + c_src*N + (1 + jq*hs)*S + (1 + iq*hf)*F;
         ops.push_back(op);
         return true;
      }
      auto it = c.faces.find((long long)b.slot*6 + l);
      if (it == c.faces.end() || it->second.size() != 1) {
         c.err = "off-rank coarse neighbour without a comm-list entry";
         return false;
      }
      const int f = it->second[0], fc = c.in.cl[d].face_case[f]%10;
      if (fc < 2 || fc > 5) {
         c.err = "comm-list case does not match a coarser off-rank neighbour";
         return false;
      }
      // unpack cases 2-5: every received value replicated 2x2 (comm.c:1020-1028)
      op.mode = FM_REPL;
      op.src_mem = (unsigned char)(BM_BUF0 + d);
      op.src_base = c.in.cl[d].recv_off[f];
      op.src_vs = (long long)hs*hf;
      op.src_str[d] = 0;
      op.src_str[sa] = hf;
      op.src_str[fa] = 1;
      ops.push_back(op);
      return true;
   }
   // FT_FINER: four quarters, each from one fine neighbour (comm.c:1626-1629) or
   // from its own message (unpack cases 6-9, comm.c:1029-1048)
   const int f_src = minus ? g.n[d] : 1;
   for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) {
         const int m = b.nei[l][i][j];
         if (m < 0) continue;
         if (m >= c.in.max_blocks || c.slot2idx[m] < 0) {
            c.err = "ERROR: misconnected block";
            return false;
         }
         BoxOp q = op;
         q.mode = FM_SUM4;
         q.ext[d] = 1; q.ext[sa] = hs; q.ext[fa] = hf;
         q.dst_base = ghost*N + (1 + j*hs)*S + (1 + i*hf)*F;
         q.src_base = (long long)m*g.tile_stride + f_src*N + S + F;
         q.S = (int)S;
         q.F = (int)F;
         ops.push_back(q);
      }
   auto it = c.faces.find((long long)b.slot*6 + l);
   if (it != c.faces.end())
      for (int f : it->second) {
         const int fc = c.in.cl[d].face_case[f]%10;
         if (fc < 6) {
            c.err = "comm-list case does not match a finer off-rank neighbour";
            return false;
         }
         int s0, s1, f0, f1;
         quarter_range(g, fc, sa, fa, s0, s1, f0, f1);
         BoxOp q = op;
         q.ext[d] = 1; q.ext[sa] = s1 - s0 + 1; q.ext[fa] = f1 - f0 + 1;
         q.dst_base = ghost*N + s0*S + f0*F;
         q.src_mem = (unsigned char)(BM_BUF0 + d);
         q.src_base = c.in.cl[d].recv_off[f];
         q.src_vs = (long long)q.ext[sa]*q.ext[fa];
         q.src_str[d] = 0;
         q.src_str[sa] = q.ext[fa];
         q.src_str[fa] = 1;
         ops.push_back(q);
      }
   return true;
}

}  // namespace

void build_halo_plan(const PlanInput &in, HaloPlan &out)
{
   Ctx c(in);
   out.ops.clear();
   out.begin.assign(in.blocks->size() + 1, 0);
   out.ok = false;
   for (size_t a = 0; a < in.blocks->size(); a++) {
      out.begin[a] = (int)out.ops.size();
      // planes 0 and n+1 first, so the kernel can fill them while the bulk copy
      // of planes 1..n is still in flight
      for (int pass = 0; pass < 2; pass++)
         for (int ri = -1; ri <= 1; ri++)
            for (int rj = -1; rj <= 1; rj++)
               for (int rk = -1; rk <= 1; rk++) {
                  if (!ri && !rj && !rk) continue;
                  if ((pass == 0) != (ri != 0)) continue;
                  const int r[3] = { ri, rj, rk };
                  if (!region_ops(c, (int)a, r, out.ops)) {
                     out.why = c.err;
                     return;
                  }
               }
   }
   out.begin[in.blocks->size()] = (int)out.ops.size();
   out.max_ops = 0;
   for (size_t a = 0; a < in.blocks->size(); a++) {
      int first = 0;
      for (int o = out.begin[a]; o < out.begin[a + 1]; o++) {
         out.ops[o].first = first;
         first += out.ops[o].ext[0]*out.ops[o].ext[1]*out.ops[o].ext[2];
      }
      out.max_ops = std::max(out.max_ops, out.begin[a + 1] - out.begin[a]);
   }
   out.elidable = true;
   for (const BoxOp &op : out.ops)
      if ((op.flags & BF_GHOST_SRC) && !(op.flags & BF_IDENT)) out.elidable = false;
   out.ok = true;
}

// pack_face code 0 (comm.c:254-401) for the faces of direction phase `phase`,
// reading resolved origins instead of this rank's (unmaterialised) ghost cells
bool build_pack_plan(const PlanInput &in, int phase, std::vector<BoxOp> &ops, std::vector<int> &fbegin,
                     std::string &why)
{
   Ctx c(in);
   const Geometry &g = c.g;
   const int d = in.order[phase];
   const DirLists &L = in.cl[d];
   int sa, fa;
   face_axes(d, sa, fa);
   const long long S = g.str[sa], F = g.str[fa], N = g.str[d];
   ops.clear();
   fbegin.assign(L.block.size() + 1, 0);
   for (size_t f = 0; f < L.block.size(); f++) {
      fbegin[f] = (int)ops.size();
      const int slot = L.block[f];
      if (slot < 0 || slot >= in.max_blocks || c.slot2idx[slot] < 0) {
         why = "comm list names an inactive block";
         return false;
      }
      const int a = c.slot2idx[slot];
      int fc = L.face_case[f], plane = 1;
      if (fc >= 10) { plane = g.n[d]; fc -= 10; }
      BoxOp op;
      init_op(g, op);
      op.dst_mem = (unsigned char)(BM_BUF0 + d);
      op.dst_str[d] = 0;
      op.dst_str[fa] = 1;
      if (fc < 2) {
         int s0, s1, f0, f1;
         whole_extent(g, fc == 1, d, sa, s0, s1);
         whole_extent(g, fc == 1, d, fa, f0, f1);
         const int Nf = f1 - f0 + 1, Ns = s1 - s0 + 1;
         // split the (possibly widened) face into interior / ghost runs per axis
         for (int ps = -1; ps <= 1; ps++)
            for (int pf = -1; pf <= 1; pf++) {
               if (ps && s0 == 1) continue;
               if (pf && f0 == 1) continue;
               int cc[3];
               cc[d] = plane;
               cc[sa] = ps < 0 ? 0 : (ps > 0 ? g.n[sa] + 1 : -1);
               cc[fa] = pf < 0 ? 0 : (pf > 0 ? g.n[fa] + 1 : -1);
               BoxOp q = op;
               q.ext[d] = 1;
               q.ext[sa] = ps ? 1 : g.n[sa];
               q.ext[fa] = pf ? 1 : g.n[fa];
               const int cs = ps < 0 ? 0 : (ps > 0 ? g.n[sa] + 1 : 1);
               const int cf = pf < 0 ? 0 : (pf > 0 ? g.n[fa] + 1 : 1);
               q.dst_base = (long long)L.send_off[f] + (long long)(cs - s0)*Nf + (cf - f0);
               q.dst_vs = (long long)Ns*Nf;
               q.dst_str[sa] = Nf;
               Origin og;
               if (!resolve(c, a, cc, phase, og)) {
                  why = c.err;
                  return false;
               }
               origin_op(c, og, q.ext, q);
               ops.push_back(q);
            }
      } else if (fc <= 5) {
         // fine -> coarse: 4-term sums of my interior plane (comm.c:271-280)
         op.mode = FM_SUM4;
         op.ext[d] = 1; op.ext[sa] = g.n[sa]/2; op.ext[fa] = g.n[fa]/2;
         op.src_base = (long long)slot*g.tile_stride + plane*N + S + F;
         op.S = (int)S; op.F = (int)F;
         op.dst_base = L.send_off[f];
         op.dst_vs = (long long)op.ext[sa]*op.ext[fa];
         op.dst_str[sa] = op.ext[fa];
         ops.push_back(op);
      } else {
         // coarse -> fine: my quarter / 4 (comm.c:281-300)
         int s0, s1, f0, f1;
         quarter_range(g, fc, sa, fa, s0, s1, f0, f1);
         op.mode = FM_DIV4;
         op.ext[d] = 1; op.ext[sa] = s1 - s0 + 1; op.ext[fa] = f1 - f0 + 1;
         op.src_base = (long long)slot*g.tile_stride + plane*N + s0*S + f0*F;
         op.dst_base = L.send_off[f];
         op.dst_vs = (long long)op.ext[sa]*op.ext[fa];
         op.dst_str[sa] = op.ext[fa];
         ops.push_back(op);
      }
   }
   fbegin[L.block.size()] = (int)ops.size();
   return true;
}

}  // namespace mamr
