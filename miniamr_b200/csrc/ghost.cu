// Ghost-face exchange kernel: one launch executes one direction phase of
// comm() (comm.c:42-242) — every on-rank face copy (on_proc_comm :1473-1534),
// level-boundary restriction / prolongation (on_proc_comm_diff :1597-1688),
// reflective boundary (apply_bc :1911-1965) and, for off-rank partners, the
// packing of faces into the device send buffer (pack_face :254-401) or the
// unpacking of the receive buffer (unpack_face :1002-1150).
//
// Within a direction phase no location is both read and written and every ghost
// cell has exactly one producer, so the phase is one fully
// parallel kernel over (face op, variable, cell).
//
// HBM-bound copy kernel: grid = (ops, variables); the fast index of the face
// rectangle is mapped to consecutive threads, so X and Y faces move as
// contiguous rows of nz doubles.  Z faces are strided by the row pitch (one
// 32-byte sector per cell), which is inherent to a k-fastest tile.
#include "common.cuh"

namespace mamr {

template <int MODE>
__device__ __forceinline__ double face_value(const double *__restrict__ src, int a, int b,
                                             int S, int F)
{
   if (MODE == FM_COPY)
      return src[(long long)a*S + (long long)b*F];
   if (MODE == FM_DIV4)
      return src[(long long)a*S + (long long)b*F]/4.0;
   if (MODE == FM_PROLONG)
      return src[(long long)(a >> 1)*S + (long long)(b >> 1)*F]/4.0;
   if (MODE == FM_REPL)
      return src[(long long)(a >> 1)*S + (long long)(b >> 1)*F];
   // FM_SUM4: left-to-right, slow index outer (comm.c:1626-1629, 274-277)
   const double *p = src + (long long)(2*a)*S + (long long)(2*b)*F;
   double s = p[0] + p[F];
   s += p[S];
   s += p[S + F];
   return s;
}

__global__ void __launch_bounds__(128)
ghost_phase_kernel(const FaceOp *__restrict__ ops, double *__restrict__ pool,
                   double *__restrict__ send_buf, const double *__restrict__ recv_buf,
                   long long pool_var_stride, int start, int buf_off)
{
   const FaceOp op = ops[blockIdx.x];
   const int vloc = blockIdx.y;
   // message buffers are indexed from the first variable of the comm() call
   double *dst = (op.mem & MEM_DST_SEND)
                    ? send_buf + op.dst_base + (long long)(vloc + buf_off)*op.dst_vs
                    : pool + (long long)start*pool_var_stride + op.dst_base + (long long)vloc*op.dst_vs;
   const double *src = (op.mem & MEM_SRC_RECV)
                          ? recv_buf + op.src_base + (long long)(vloc + buf_off)*op.src_vs
                          : pool + (long long)start*pool_var_stride + op.src_base +
                               (long long)vloc*op.src_vs;
   const int cells = op.Ns*op.Nf;
   const int Nf = op.Nf;
   for (int c = threadIdx.x; c < cells; c += blockDim.x) {
      const int a = c/Nf, b = c - a*Nf;
      double v;
      switch (op.mode) {
      case FM_COPY: v = face_value<FM_COPY>(src, a, b, op.src_S, op.src_F); break;
      case FM_DIV4: v = face_value<FM_DIV4>(src, a, b, op.src_S, op.src_F); break;
      case FM_PROLONG: v = face_value<FM_PROLONG>(src, a, b, op.src_S, op.src_F); break;
      case FM_REPL: v = face_value<FM_REPL>(src, a, b, op.src_S, op.src_F); break;
      default: v = face_value<FM_SUM4>(src, a, b, op.src_S, op.src_F); break;
      }
      dst[(long long)a*op.dst_S + (long long)b*op.dst_F] = v;
   }
}

void launch_ghost(const FaceOp *d_ops, int n_ops, double *pool, double *send_buf,
                  const double *recv_buf, long long pool_var_stride, int start, int num,
                  int buf_var0, cudaStream_t s)
{
   if (n_ops <= 0 || num <= 0) return;
   // gridDim.x can hold 2^31-1 ops; gridDim.y (variables) is limited to 65535
   dim3 grid((unsigned)n_ops, (unsigned)num);
   ghost_phase_kernel<<<grid, 128, 0, s>>>(d_ops, pool, send_buf, recv_buf,
                                           pool_var_stride, start, start - buf_var0);
}

}  // namespace mamr
