// CPU check of the branch-free exp / exp10 of csrc/wt_simt.h (FMA + integer ops only, so the same bits as on
// the GPU) against libm: prints the largest ulp distance over a dense sample and the special cases.
#define WT_EMU
#include "wt_simt.h"
#include <stdio.h>
#include <stdlib.h>
static double ulps(double a, double b) {
  if (a == b || (isnan(a) && isnan(b))) return 0;
  if (!isfinite(a) || !isfinite(b)) return 1e30;
  int e; frexp(b, &e);
  return fabs(a - b) / ldexp(1.0, e - 53);
}
int main() {
  double we = 0, w10 = 0, wd = 0;
  unsigned long long st = 88172645463325252ull;
  for (int i = 0; i < 3000000; ++i) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    const double u = (st >> 11) * (1.0 / 9007199254740992.0);
    const double x = -700 + 1400 * u, y = -300 + 600 * u, z = -20 + 40 * u;
    double d = ulps(wt_exp_s(x), exp(x)); if (d > we) we = d;
    d = ulps(wt_exp10_s(y), pow(10.0, y)); if (d > w10) w10 = d;
    d = ulps(wt_exp10_s(z), pow(10.0, z)); if (d > w10) w10 = d;
    const double s = -745.1 + 37 * u;  /* gradual underflow: absolute error in units of the smallest denormal */
    d = fabs(wt_exp_s(s) - exp(s)) / 4.9406564584124654e-324 * (exp(s) < 2.3e-308 ? 1 : 0); if (d > wd) wd = d;
  }
  printf("max_ulp_exp %.3f\nmax_ulp_exp10 %.3f\nmax_denormal_units %.3f\n", we, w10, wd);
  const double xs[] = {0, -0.0, 1, -1, 709.7, 709.8, 745, -745, -745.2, -800, 1e10, -1e10, INFINITY, -INFINITY, NAN};
  for (double x : xs) printf("exp %a %a %a\n", x, wt_exp_s(x), exp(x));
  const double ys[] = {0, 7, -7, -14, 308, 308.3, 309, -307.7, -310, -323, -324, -400, 400, 1e300, -1e300, INFINITY, -INFINITY, NAN};
  for (double y : ys) printf("exp10 %a %a %a\n", y, wt_exp10_s(y), pow(10.0, y));
  return 0;
}
