// Micro-benchmark: dependent-chain latency and per-SM throughput of DADD/DFMA on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void dadd_chain(double *out, long long *cyc, int iters, double seed)
{
   double a[CHAINS];
#pragma unroll
   for (int c = 0; c < CHAINS; c++) a[c] = seed + c + threadIdx.x;
   long long t0 = clock64();
   for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 8; u++)
#pragma unroll
         for (int c = 0; c < CHAINS; c++) a[c] = a[c] + seed;
   }
   long long t1 = clock64();
   double s = 0;
#pragma unroll
   for (int c = 0; c < CHAINS; c++) s += a[c];
   out[blockIdx.x*blockDim.x + threadIdx.x] = s;
   if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS>
void run(int warps, int blocks)
{
   double *out; long long *cyc, h;
   cudaMalloc(&out, sizeof(double)*blocks*warps*32); cudaMalloc(&cyc, 8);
   const int iters = 2000;
   dadd_chain<CHAINS><<<blocks, warps*32>>>(out, cyc, iters, 1.0000001);
   dadd_chain<CHAINS><<<blocks, warps*32>>>(out, cyc, iters, 1.0000001);
   cudaDeviceSynchronize();
   cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
   double per = (double)h/(iters*8.0*CHAINS);
   printf("chains=%d warps/CTA=%d CTAs=%d : %.2f cycles per DADD per warp (%.2f per chain step); warp-instr/clk/SM=%.2f\n",
          CHAINS, warps, blocks, per, per*CHAINS, warps/per);
   cudaFree(out); cudaFree(cyc);
}
int main()
{
   run<1>(1, 1); run<2>(1, 1); run<4>(1, 1); run<8>(1, 1);
   run<1>(4, 1); run<4>(4, 1); run<8>(4, 1); run<4>(8, 1); run<4>(16, 1); run<8>(16, 1); run<8>(32, 1);
   return 0;
}
