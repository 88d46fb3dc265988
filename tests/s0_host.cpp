// TEST INFRASTRUCTURE: the --stencil 0 per-cell functions of miniamr_b200/csrc/stencil0.cuh
// compiled for the host and applied to one block in the reference's own loop order
// (stencil.c:43-74,147-983), so that the CPU suite can pin them bit for bit against the
// unmodified reference (tests/test_stencil0_formulas.py) without a GPU.
//   g++ -O2 -ffp-contract=off -shared -fPIC -o tests/_build/libs0host.so tests/s0_host.cpp
#include <vector>
#include "../miniamr_b200/csrc/stencil0.cuh"

using namespace mamr;

extern "C" void s0_host_driver(double *data, int nx, int ny, int nz, int num_vars, int var, int stage,
                               int mat, double a1, const double *a0, double *flops)
{
   const long long SJ = nz + 2, PL = (long long)(ny + 2)*SJ, VS = (long long)(nx + 2)*PL;
   S0Coef c = { mat, a1, a0 };
   auto at = [&](int v, int i, int j, int k) -> double & { return data[v*VS + i*PL + j*SJ + k]; };
   const double cells = (double)nx*ny*nz;
   if (var == 0 || var >= 4*mat) {                       // stencil_calc(var, 7), stencil.c:82-101
      std::vector<double> work((size_t)VS);
      for (int i = 1; i <= nx; i++)
         for (int j = 1; j <= ny; j++)
            for (int k = 1; k <= nz; k++)
               work[i*PL + j*SJ + k] = (at(var, i - 1, j, k) + at(var, i, j - 1, k) + at(var, i, j, k - 1) +
                                        at(var, i, j, k) + at(var, i, j, k + 1) + at(var, i, j + 1, k) +
                                        at(var, i + 1, j, k))/7.0;
      for (int i = 1; i <= nx; i++)
         for (int j = 1; j <= ny; j++)
            for (int k = 1; k <= nz; k++) at(var, i, j, k) = work[i*PL + j*SJ + k];
      flops[0] += 6*cells; flops[2] += cells;
      return;
   }
   const int kind = stage%6;
   if (kind <= S0_SWEEP_K) {
      const long long d = kind == S0_SWEEP_I ? PL : (kind == S0_SWEEP_J ? SJ : 1);
      for (int i = 1; i <= nx; i++)
         for (int j = 1; j <= ny; j++)
            for (int k = 1; k <= nz; k++) {
               const double *p = &at(0, i, j, k);
               at(var, i, j, k) = kind == S0_POINT ? s0_point(p, VS, var, c) : s0_sweep(p, VS, var, c, d);
            }
   } else {
      std::vector<double> work((size_t)VS);
      for (int i = 1; i <= nx; i++)
         for (int j = 1; j <= ny; j++)
            for (int k = 1; k <= nz; k++) {
               const double *p = &at(0, i, j, k);
               work[i*PL + j*SJ + k] = kind == S0_SEVEN ? s0_seven(p, VS, var, c, PL, SJ)
                                                        : s0_twenty7(p, VS, var, c, PL, SJ);
            }
      for (int i = 1; i <= nx; i++)
         for (int j = 1; j <= ny; j++)
            for (int k = 1; k <= nz; k++) at(var, i, j, k) = work[i*PL + j*SJ + k];
   }
   const S0Flops f = s0_flops(kind, var, mat);
   flops[0] += f.adds*cells; flops[1] += f.muls*cells; flops[2] += f.divs*cells;
   for (int i = 1; i <= nx; i++)                          // stencil_check(var), stencil.c:959-983
      for (int j = 1; j <= ny; j++)
         for (int k = 1; k <= nz; k++) {
            int what;
            at(var, i, j, k) = s0_check(at(var, i, j, k), c, &what);
            if (what == 1) { flops[2] += 1; flops[0] += 2; }
            else if (what == 2) { flops[1] += 1; flops[0] += 1; }
         }
}
