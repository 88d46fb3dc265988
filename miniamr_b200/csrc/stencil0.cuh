// --stencil 0: the "variable work" updates of stencil_driver(), stencil.c:43-74 and
// :147-983, restated as per-cell functions that compile both for the device
// (stencil0.cu) and for the host (tests/s0_host.cpp, which checks them against the
// unmodified reference without a GPU).
//
// How the reference is organised (and how this file regroups it):
//   mat = num_vars/4 (init.c:419).  Variable 0 and variables >= 4*mat take the plain
//   7-point average (stencil.c:47-48,71-72).  Every other variable belongs to band
//   q = var/mat (0..3) and column b = var%mat, and stage%6 selects the update kind:
//     0  pointwise          stencil_0   :147-226
//     1  sweep along i      stencil_x   :228-370   } in place: the cell at -1 along the
//     2  sweep along j      stencil_y   :372-514   } sweep axis already holds its NEW
//     3  sweep along k      stencil_z   :516-659   } value, the cell at +1 its old one
//     4  7-point, weighted  stencil_7   :661-784   } through a work[] copy (Jacobi)
//     5  27-point, banded   stencil_27  :786-957   }
//   followed by stencil_check(var) :959-983 on the updated variable.
//   The three sweeps are the same code with the axis renamed; stencil_7 and
//   stencil_27 pick their coefficient variables from the other three bands of the
//   same column in rotating order: band (q+1)%4, (q+2)%4, (q+3)%4.
//   Variable 1 is special in kinds 0-3 (a sum over a range of other variables).
//
// Every expression below keeps the reference's association (C evaluates a+b+c as
// (a+b)+c and a*b/c as (a*b)/c), and the translation units that include this file
// are compiled without FMA contraction, so results are bit-identical.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define S0_HD __host__ __device__ __forceinline__
#else
#define S0_HD inline
#endif

namespace mamr {

struct S0Coef {
   int mat;            // num_vars/4
   double a1;          // init.c:420
   const double *a0;   // [mat], init.c:421-422
};

enum S0Kind { S0_POINT = 0, S0_SWEEP_I = 1, S0_SWEEP_J = 2, S0_SWEEP_K = 3, S0_SEVEN = 4, S0_TWENTY7 = 5 };

// p points at variable 0 of the current cell; VS = element stride between variables
#define S0_AT(v) p[(long long)(v)*VS]
#define S0_OFF(v, o) p[(long long)(v)*VS + (o)]

// kind 0, stencil.c:147-226
S0_HD double s0_point(const double *p, long long VS, int var, const S0Coef &c)
{
   const int mat = c.mat;
   const double a1 = c.a1;
   const double A0 = S0_AT(0);
   double A = S0_AT(var);
   if (var == 1) {                                      // :152-163
      for (int v = mat; v < 2*mat; v++) A += S0_AT(v)*A0;
      return A;
   }
   const double A1 = S0_AT(1);
   if (var < mat)                                       // :164-176
      return A + A*(A0 + A1 - a1*A);
   if (var < 2*mat)                                     // :177-193
      return A*(A0 + A + a1*S0_AT(var + mat) + (1.0 - a1)*S0_AT(var + 2*mat))/A1;
   if (var < 3*mat)                                     // :194-209
      return A + S0_AT(var - mat)*(a1*A0 + c.a0[var - 2*mat]*A + (1.0 - a1)*S0_AT(var + mat))/A1;
   {                                                    // :210-225
      const double a0v = c.a0[var - 3*mat];
      const double B2 = S0_AT(var - 2*mat);
      return A + B2*(a1*A0 + a0v*A + (1.0 - a0v)*S0_AT(var - mat) + (1.0 - a1)*B2)/(A1*A1);
   }
}

// kinds 1-3, stencil.c:228-370 (x), :372-514 (y), :516-659 (z); d = element offset of
// +1 along the sweep axis.  The caller visits cells in ascending order along that axis.
S0_HD double s0_sweep(const double *p, long long VS, int var, const S0Coef &c, long long d)
{
   const int mat = c.mat;
   const double a1 = c.a1;
   const double A0 = S0_AT(0);
   double A = S0_AT(var);
   if (var == 1) {                                      // :234-248
      for (int v = 2; v < mat + 2; v++) A += S0_AT(v)*A0;
      return A/(a1 + A);
   }
   const double A1 = S0_AT(1);
   if (var < mat)                                       // :249-262
      return A + A*(A0 + A1 - a1*A)/(c.a0[var] + A1);
   // the variable whose slope picks the branch, and the rest of the denominator
   const double TM = S0_OFF(var, -d), TP = S0_OFF(var, d);
   double tmp1, tmp2, den;
   if (var < 2*mat) {                                   // :263-296
      tmp1 = fabs(A - TM);
      tmp2 = fabs(A - TP);
      den = a1 + c.a0[var - mat] + TM + A + TP + A0 + A1;
   } else if (var < 3*mat) {                            // :297-331
      const double BC = S0_AT(var - mat);
      tmp1 = fabs(BC - S0_OFF(var - mat, -d));
      tmp2 = fabs(BC - S0_OFF(var - mat, d));
      den = a1 + c.a0[var - 2*mat] + BC + S0_AT(var + mat) + TM + A + TP;
   } else {                                             // :332-366
      const double BC = S0_AT(var - 2*mat);
      tmp1 = fabs(BC - S0_OFF(var - 2*mat, -d));
      tmp2 = fabs(BC - S0_OFF(var - 2*mat, d));
      den = a1 + c.a0[var - 3*mat] + S0_AT(var - mat) + BC + TM + A + TP;
   }
   const double diff = tmp1 > tmp2 ? tmp1 - tmp2 : tmp2 - tmp1;
   return (tmp1*TM + diff*(A + A1) + tmp2*TP)/den;
}

// kind 4, stencil.c:661-784; PL / SJ = element offsets of +1 along i / j (k is contiguous)
S0_HD double s0_seven(const double *p, long long VS, int var, const S0Coef &c, long long PL, long long SJ)
{
   const int mat = c.mat, q = var/mat, b = var - q*mat;
   const int cx = b + ((q + 1) & 3)*mat, cy = b + ((q + 2) & 3)*mat, cz = b + ((q + 3) & 3)*mat;
   const double C = S0_AT(var);
   return (S0_OFF(var, -PL)*S0_OFF(cx, -PL) +
           S0_OFF(var, -SJ)*S0_OFF(cy, -SJ) +
           S0_OFF(var, -1)*S0_OFF(cz, -1) +
           C*C +
           S0_OFF(var, 1)*S0_OFF(cz, 1) +
           S0_OFF(var, SJ)*S0_OFF(cy, SJ) +
           S0_OFF(var, PL)*S0_OFF(cx, PL))/7.0*(c.a1 + C);
}

// kind 5, stencil.c:786-957: the 27 neighbours in (i, j, k) lexicographic order, each
// taken from the band its distance class names (centre: the variable itself, face: band
// q+1, edge: q+2, corner: q+3)
S0_HD double s0_twenty7(const double *p, long long VS, int var, const S0Coef &c, long long PL, long long SJ)
{
   const int mat = c.mat, q = var/mat, b = var - q*mat;
   int band[4];
   band[0] = var;
   band[1] = b + ((q + 1) & 3)*mat;
   band[2] = b + ((q + 2) & 3)*mat;
   band[3] = b + ((q + 3) & 3)*mat;
   double s = 0.0;
   bool first = true;
   for (int di = -1; di <= 1; di++)
      for (int dj = -1; dj <= 1; dj++)
         for (int dk = -1; dk <= 1; dk++) {
            const int cls = (di != 0) + (dj != 0) + (dk != 0);
            const double x = S0_OFF(band[cls], di*PL + dj*SJ + dk);
            s = first ? x : s + x;
            first = false;
         }
   return s/(c.a1 + 27.0);
}

// stencil_check, stencil.c:959-983.  what: 0 untouched, 1 divided, 2 scaled (the
// reference counts its flops per cell, :970-971,975-976)
S0_HD double s0_check(double x, const S0Coef &c, int *what)
{
   x = fabs(x);
   *what = 0;
   if (x >= 1.0) {
      *what = 1;
      return x/(c.a1 + c.a0[0] + x);
   }
   if (x < 0.1) {
      *what = 2;
      return x*(10.0 - c.a1);
   }
   return x;
}

#undef S0_AT
#undef S0_OFF

// flops per cell the reference books for one update of `var` (without stencil_check)
struct S0Flops { double adds, muls, divs; };
inline S0Flops s0_flops(int kind, int var, int mat)
{
   if (kind == S0_POINT) {
      if (var == 1) return { (double)mat, (double)mat, 0.0 };        // :161-162
      if (var < mat) return { 3, 2, 0 };                              // :174-175
      if (var < 2*mat) return { 3, 3, 1 };                            // :190-192
      if (var < 3*mat) return { 4, 3, 1 };                            // :206-208
      return { 6, 6, 1 };                                             // :222-224
   }
   if (kind <= S0_SWEEP_K) {
      if (var == 1) return { (double)mat + 1, (double)mat, 1 };      // :245-247
      if (var < mat) return { 4, 2, 1 };                              // :259-261
      return { 12, 3, 1 };                                            // :294-296 ...
   }
   if (kind == S0_SEVEN) return { 7, 8, 1 };                          // :692-694 ...
   return { 27, 0, 1 };                                               // :831-832 ...
}

}  // namespace mamr
