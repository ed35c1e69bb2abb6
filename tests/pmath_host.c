/* Host build of the product's portable elementary functions (os-stab_b200/csrc/osb_pmath.h),
 * compiled with -ffp-contract=off.  TEST INFRASTRUCTURE: handed to the CPU oracle through
 * osr_set_math so that the oracle and the CUDA strict build evaluate EXP / SINCOS / HYPOT / ACOS with
 * the same IEEE-only code (tests/test_strict_parity.py). */
#include "../os-stab_b200/csrc/osb_pmath.h"

double pmh_exp(double x) { return pm_exp(x); }
void pmh_sincos(double x, double *s, double *c) { pm_sincos(x, s, c); }
double pmh_hypot(double x, double y) { return pm_hypot(x, y); }
double pmh_acos(double x) { return pm_acos(x); }

/* array forms for the accuracy test */
void pmh_exp_v(long n, const double *x, double *y) { for (long i = 0; i < n; ++i) y[i] = pm_exp(x[i]); }
void pmh_sincos_v(long n, const double *x, double *s, double *c) { for (long i = 0; i < n; ++i) pm_sincos(x[i], s + i, c + i); }
void pmh_hypot_v(long n, const double *x, const double *y, double *h) { for (long i = 0; i < n; ++i) h[i] = pm_hypot(x[i], y[i]); }
void pmh_acos_v(long n, const double *x, double *y) { for (long i = 0; i < n; ++i) y[i] = pm_acos(x[i]); }
