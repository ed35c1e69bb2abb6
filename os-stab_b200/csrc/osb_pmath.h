/*
 * osb_pmath.h -- portable elementary functions for the STRICT-parity build of the shooting kernel.
 *
 * The reference's arithmetic is IEEE double without FMA contraction (gcc.mak:73-79); everything in
 * its eigenvalue iteration is +,-,*,/ and sqrt -- which are bit-reproducible on any IEEE machine --
 * except a handful of libm calls per pass: EXP/SINCOS (free-stream initial vectors, contebl.f:269-279),
 * HYPOT (inside ABS and SQRT of a complex: the orthonormality test shoot.f:593-606, Gram-Schmidt,
 * Muller's update shoot.f:733-741) and ACOS (the test itself).  libm's results for those depend on the
 * libm (glibc version, CUDA's own implementation), so no two builds of the reference agree bitwise
 * either.  The strict build therefore evaluates them with the functions below: plain C, only IEEE
 * operations in a fixed order, accurate to ~1 ulp, compiled WITHOUT contraction on both sides
 * (nvcc -fmad=false for shoot_strict.cu; gcc -ffp-contract=off for the test-side copy that is handed
 * to the CPU checker by the tests).  With the same four functions on both sides the GPU
 * eigenvalue iteration is bit-identical to the oracle's (tests/test_strict_parity.py); how little the
 * choice of libm matters is measured separately by swapping them in the oracle alone.
 *
 * Algorithms: Cody-Waite argument reduction + the fdlibm kernels (exp by a Taylor/Horner polynomial).
 * No dependence on <math.h> beyond sqrt.
 */
#ifndef OSB_PMATH_H
#define OSB_PMATH_H
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PM_FN __host__ __device__ static inline
#else
#define PM_FN static inline
#endif

#if defined(__CUDA_ARCH__)
#define PM_SQRT(x) __dsqrt_rn(x)
#else
#include <math.h>
#define PM_SQRT(x) sqrt(x)
#endif

PM_FN double pm_from_bits(uint64_t u) {
  double d;
  memcpy(&d, &u, sizeof(d));
  return d;
}
PM_FN uint64_t pm_to_bits(double d) {
  uint64_t u;
  memcpy(&u, &d, sizeof(u));
  return u;
}
PM_FN double pm_fabs(double x) { return pm_from_bits(pm_to_bits(x) & 0x7fffffffffffffffull); }
/* 2^e for -1022 <= e <= 1023 */
PM_FN double pm_pow2(int e) { return pm_from_bits((uint64_t)(e + 1023) << 52); }

/* sqrt(x^2 + y^2); exact when one argument is zero */
PM_FN double pm_hypot(double x, double y) {
  double a = pm_fabs(x), b = pm_fabs(y);
  if (a < b) { const double t = a; a = b; b = t; }
  if (b == 0.0) return a;
  const double r = b / a;
  return a * PM_SQRT(1.0 + r * r);
}

/* exp(x) for |x| < 700 */
PM_FN double pm_exp(double x) {
  const double inv_ln2 = 1.44269504088896338700e+00;
  const double ln2_hi = 6.93147180369123816490e-01; /* 33 significant bits */
  const double ln2_lo = 1.90821492927058770002e-10;
  const double kd = x * inv_ln2;
  const int k = (int)(kd < 0.0 ? kd - 0.5 : kd + 0.5);
  const double fk = (double)k;
  const double r = (x - fk * ln2_hi) - fk * ln2_lo; /* |r| <= 0.3466 */
  /* Taylor polynomial of degree 14, Horner */
  double p = 1.1470745597729725e-11; /* 1/14! */
  p = p * r + 1.6059043836821613e-10; /* 1/13! */
  p = p * r + 2.08767569878681e-09;   /* 1/12! */
  p = p * r + 2.505210838544172e-08;  /* 1/11! */
  p = p * r + 2.755731922398589e-07;  /* 1/10! */
  p = p * r + 2.7557319223985893e-06; /* 1/9!  */
  p = p * r + 2.48015873015873e-05;   /* 1/8!  */
  p = p * r + 0.0001984126984126984;  /* 1/7!  */
  p = p * r + 0.001388888888888889;   /* 1/6!  */
  p = p * r + 0.008333333333333333;   /* 1/5!  */
  p = p * r + 0.041666666666666664;   /* 1/4!  */
  p = p * r + 0.16666666666666666;    /* 1/3!  */
  p = p * r + 0.5;
  /* exp(r) = 1 + r + r^2 p : the two leading terms are added last */
  const double e = 1.0 + (r + (r * r) * p);
  const int k1 = k / 2, k2 = k - k1;
  return (e * pm_pow2(k1)) * pm_pow2(k2);
}

/* fdlibm __kernel_sin / __kernel_cos on [-pi/4, pi/4] with a tail y */
PM_FN double pm_ksin(double x, double y) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double z = x * x, v = z * x;
  const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
PM_FN double pm_kcos(double x, double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double z = x * x;
  const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  const double ax = pm_fabs(x);
  if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
  /* qx ~ |x|/4 with the low word cleared, as fdlibm: keeps 1 - qx and 0.5 z - qx exact */
  const double qx = ax > 0.78125 ? 0.28125 : pm_from_bits((pm_to_bits(ax) - (2ull << 52)) & 0xffffffff00000000ull);
  const double hz = 0.5 * z - qx, a = 1.0 - qx;
  return a - (hz - (z * r - x * y));
}

/* sin and cos of y for |y| < ~1e5 (three-term Cody-Waite reduction by pi/2) */
PM_FN void pm_sincos(double y, double *s, double *c) {
  const double invpio2 = 6.36619772367581382433e-01;
  const double pio2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
  const double pio2_2 = 6.07710050630396597660e-11;  /* second 33 bits */
  const double pio2_3 = 2.02226624871116645580e-21;  /* third 33 bits */
  const double pio2_3t = 8.47842766036889956997e-32; /* pi/2 - (pio2_1 + pio2_2 + pio2_3) */
  const double kd = y * invpio2;
  const int n = (int)(kd < 0.0 ? kd - 0.5 : kd + 0.5);
  const double fn = (double)n;
  /* r = y - n pi/2 as head + tail; n * pio2_{1,2,3} are exact (33-bit constants, |n| < 2^20) */
  const double t1 = y - fn * pio2_1;
  const double t2 = t1 - fn * pio2_2;
  const double w = fn * pio2_3;
  const double hi = t2 - w;
  const double lo = ((t2 - hi) - w) - fn * pio2_3t;
  const double sn = pm_ksin(hi, lo), cs = pm_kcos(hi, lo);
  switch (n & 3) {
    case 0: *s = sn; *c = cs; break;
    case 1: *s = cs; *c = -sn; break;
    case 2: *s = -sn; *c = -cs; break;
    default: *s = -cs; *c = sn; break;
  }
}

/* acos(x), |x| <= 1 (fdlibm e_acos.c) */
PM_FN double pm_acos(double x) {
  const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
  const double pi = 3.14159265358979311600e+00;
  const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
               pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
               pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
               qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
               qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
  const double ax = pm_fabs(x);
  if (ax >= 1.0) return x > 0.0 ? 0.0 : pi + 2.0 * pio2_lo;
  if (ax < 0.5) {
    if (ax < 6.0e-18) return pio2_hi + pio2_lo;
    const double z = x * x;
    const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const double r = p / q;
    return pio2_hi - (x - (pio2_lo - x * r));
  }
  if (x < 0.0) {
    const double z = (1.0 + x) * 0.5;
    const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const double s = PM_SQRT(z);
    const double r = p / q;
    const double w = r * s - pio2_lo;
    return pi - 2.0 * (s + w);
  }
  {
    const double z = (1.0 - x) * 0.5;
    const double s = PM_SQRT(z);
    const double df = pm_from_bits(pm_to_bits(s) & 0xffffffff00000000ull);
    const double cc = (z - df * df) / (s + df);
    const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const double r = p / q;
    const double w = r * s + cc;
    return 2.0 * (df + w);
  }
}

#endif /* OSB_PMATH_H */
