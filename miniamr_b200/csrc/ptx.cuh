// PTX helpers shared by the stencil kernels: mbarrier + 1-D bulk TMA copies,
// and the exact division by the stencil constants.
#pragma once
#include <stdint.h>

namespace mamr {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
   return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_barrier_init()
{
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                "r"(bytes)
                : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
   asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// named barrier over `count` threads (a multiple of 32); id 0 is __syncthreads()
__device__ __forceinline__ void named_bar_sync(int id, int count)
{
   asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
   uint32_t ok;
   asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
   return ok != 0;
}

// 8-byte asynchronous copy global -> shared (SASS LDGSTS): no register staging
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc)
                : "memory");
}

__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// 1-D bulk TMA copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes,
                                         uint64_t *bar)
{
   asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
         "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// 1-D bulk TMA copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes)
{
   asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                "r"(smem_u32(smem_src)), "r"(bytes)
                : "memory");
}

__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

// wait until the bulk stores of this thread have finished READING shared memory
__device__ __forceinline__ void bulk_wait_read0()
{
   asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// make this thread's generic-proxy shared-memory writes visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async()
{
   asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// x / D for D = 7.0 or 27.0 with the result of an IEEE-754 division, in three
// FP64 pipe operations instead of the ~20-instruction generic division sequence.
// With y = RN(1/D): q = RN(x*y), r = x - D*q (exact in the FMA), q' = RN(q + r*y).
// q + r*y = Q + (Q - q)*delta with Q = x/D, |Q - q| <= 1.5 ulp and |delta| <= 2^-53,
// i.e. Q perturbed by < 2^-52 ulp; because D is an odd integer < 32, Q is at least
// 1/(2D) ulp away from every rounding midpoint, so RN of the perturbed value is
// RN(Q).  Outside the range where r is exact (zeros, subnormal-scale values,
// infinities, NaN) the true division is used.
template <int D>
__device__ __forceinline__ double div_const(double x)
{
   const double y = 1.0/(double)D;
   const double ax = fabs(x);
   if (!(ax >= 0x1p-900 && ax <= 0x1p+900)) return x/(double)D;
   const double q = x*y;
   const double r = fma(-(double)D, q, x);
   return fma(r, y, q);
}

}  // namespace mamr
