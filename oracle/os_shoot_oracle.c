/*
 * os_shoot_oracle.c -- CPU restatement of the reference's shooting path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may include, link or
 * call this file; it is the checker for tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * The reference (sscollis/os-stab) is Fortran-77 and no Fortran compiler exists
 * in this image, so this is a restatement in C99 of:
 *   SOLVE_BL / BLASIUS            contebl.f:386-562, orrsom.f:687-804, orrspace.f:697-817
 *   SRK4 / SRKCK45                shoot.f:824-858, 1292-1364
 *   CRK4 / CRKCK45                shoot.f:869-903, 1367-1439
 *   SPLINE / SPEVAL / SPDER / BISECT   shoot.f:906-1101 (linear-search copies
 *                                 orrsom.f:885-908, orrspace.f:906-929)
 *   OSHOMO / FHOMO                contebl.f:300-340, conte.f:180-212,
 *                                 orrsom.f:603-642, orrspace.f:607-649
 *   INPROD                        contebl.f:219-244, orrspace.f:585-604
 *   BL_IC                         contebl.f:247-282
 *   CONTE_IC / CONTE              shoot.f:391-813, shoot.f:26-388,
 *                                 orrsom.f:165-506, orrspace.f:209-582
 *
 * Arithmetic: the reference is built with -freal-4-real-8 -fdefault-real-8 -O2
 * (gcc.mak:73-79): everything is IEEE double, no fast-math, no FMA contraction
 * on generic x86-64.  gfortran expands complex multiply/divide with the
 * "Fortran rules" (plain 4-multiply product, Smith's division, no NaN
 * recovery); compiling THIS file with
 *     gcc -O2 -ffp-contract=off -fcx-fortran-rules
 * makes gcc's C front end use the very same expansions for double _Complex.
 * ABS/SQRT/EXP of complex call libm cabs/csqrt/cexp in both.
 *
 * Parity pinning: README.md:24-26 (contebl default eigenvalue, 8 digits),
 * contebl.f:75-76, conte.f:65-66, sweep.inp:15-16 -- see tests/test_oracle_kat.py.
 *
 * Undefined behaviour in the reference is resolved as SURVEY.md Appendix D says
 * (each place is marked "D-n" below).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#undef I /* complex.h's imaginary unit macro; CMPLX() is used instead */

typedef double _Complex zc;

#define OSR_FLOW_CONTEBL 0
#define OSR_FLOW_CONTE 1
#define OSR_FLOW_ORRSOM 2
#define OSR_FLOW_ORRSPACE 3

#define OSR_INT_RK4 0
#define OSR_INT_CK45 1

/* ---------------------------------------------------------------------- */
/* Elementary functions.  By default the restatement calls libm's complex   */
/* cabs / csqrt / cexp and real acos exactly like the gfortran build of the */
/* reference (mode 0).  For the strict-parity tests (tests/                 */
/* test_strict_parity.py) the REAL functions underneath can be replaced:    */
/* mode 1 evaluates the complex functions with glibc's own formulas         */
/* (s_csqrt_template.c, s_cexp_template.c, cabs = hypot) over the four real */
/* functions in g_math.  Handing it libm's exp/sincos/hypot/acos must       */
/* reproduce mode 0 bit for bit (that pins the formulas); handing it the    */
/* portable functions the CUDA strict build uses makes the two sides        */
/* comparable bit for bit.                                                  */
/* ---------------------------------------------------------------------- */
typedef struct {
  double (*exp)(double);
  void (*sincos)(double, double *, double *);
  double (*hypot)(double, double);
  double (*acos)(double);
} osr_math;
static osr_math g_math;
static int g_math_mode = 0;

void osr_set_math(const osr_math *m) {
  if (m) { g_math = *m; g_math_mode = 1; }
  else g_math_mode = 0;
}
static void osr_libm_sincos(double x, double *s, double *c) { *s = sin(x); *c = cos(x); }
/* mode 1 over libm's own real functions (the self-check of the formulas below) */
void osr_set_math_libm(void) {
  g_math.exp = exp; g_math.sincos = osr_libm_sincos; g_math.hypot = hypot; g_math.acos = acos;
  g_math_mode = 1;
}

static double m_cabs(double _Complex z) {
  return g_math_mode ? g_math.hypot(creal(z), cimag(z)) : cabs(z);
}
static double m_acos(double x) { return g_math_mode ? g_math.acos(x) : acos(x); }
/* glibc s_cexp_template.c, finite arguments below the overflow threshold */
static double _Complex m_cexp(double _Complex z) {
  if (!g_math_mode) return cexp(z);
  double sn, cs;
  if (fabs(cimag(z)) > 2.2250738585072014e-308) g_math.sincos(cimag(z), &sn, &cs);
  else { sn = cimag(z); cs = 1.0; }
  const double e = g_math.exp(creal(z));
  return CMPLX(e * cs, e * sn);
}
/* glibc s_csqrt_template.c, finite arguments away from the overflow / underflow scaling branches */
static double _Complex m_csqrt(double _Complex z) {
  if (!g_math_mode) return csqrt(z);
  const double re = creal(z), im = cimag(z);
  if (im == 0.0) {
    if (re < 0.0) return CMPLX(0.0, copysign(sqrt(-re), im));
    return CMPLX(fabs(sqrt(re)), copysign(0.0, im));
  }
  if (re == 0.0) {
    const double r = sqrt(0.5 * fabs(im));
    return CMPLX(r, copysign(r, im));
  }
  const double d = g_math.hypot(re, im);
  double r, s2;
  if (re > 0.0) {
    r = sqrt(0.5 * (d + re));
    s2 = 0.5 * (im / r);
  } else {
    s2 = sqrt(0.5 * (d - re));
    r = fabs(0.5 * (im / s2));
  }
  return CMPLX(r, copysign(s2, im));
}

/* ---------------------------------------------------------------------- */
/* Cash-Karp tableau exactly as typed in shoot.f:1306-1319 / 1381-1394      */
/* (column-major DATA fill: b(m,k) below is b[k-1][m-1]).                   */
/* ---------------------------------------------------------------------- */
static const double ck_a[6] = {0.0, 0.2, 0.3, 0.6, 1.0, 0.875};
static const double ck_bcol[5][6] = {
    {0.0, 0.2, 0.075, 0.3, -0.2037037037037037, 0.029495804398148147},
    {0.0, 0.0, 0.225, -0.9, 2.5, 0.341796875},
    {0.0, 0.0, 0.0, 1.2, -2.5925925925925926, 0.041594328703703706},
    {0.0, 0.0, 0.0, 0.0, 1.2962962962962963, 0.40034541377314814},
    {0.0, 0.0, 0.0, 0.0, 0.0, 0.061767578125}};
static const double ck_c[6] = {0.09788359788359788, 0.0, 0.4025764895330113,
                               0.21043771043771045, 0.0, 0.2891022021456804};
#define CK_B(m, k) ck_bcol[(k) - 1][(m) - 1] /* 1-based b(m,k) */

/* ---------------------------------------------------------------------- */
/* Profile tables.  Arrays are allocated one element longer than nbl+1 and  */
/* zero filled, mirroring the zero-initialised COMMON storage that SPDER    */
/* touches when XX == X(N)  (Appendix D-1).                                 */
/* ---------------------------------------------------------------------- */
/* SPLINE, shoot.f:906-957.  0-based arrays, N points. */
static void osr_spline(int N, const double *X, const double *Y, double *FDP) {
  double *A = (double *)calloc((size_t)N + 2, sizeof(double));
  double *B = (double *)calloc((size_t)N + 2, sizeof(double));
  double *C = (double *)calloc((size_t)N + 2, sizeof(double));
  double *R = (double *)calloc((size_t)N + 2, sizeof(double));
  /* use 1-based indexing inside to stay readable against the Fortran */
  const double *x = X - 1, *y = Y - 1;
  double *fdp = FDP - 1;
  const double ALAMDA = 1.0;
  int NM2 = N - 2, NM1 = N - 1;
  C[1] = x[2] - x[1];
  for (int I = 2; I <= NM1; ++I) {
    C[I] = x[I + 1] - x[I];
    A[I] = C[I - 1];
    B[I] = 2.0 * (A[I] + C[I]);
    R[I] = 6.0 * ((y[I + 1] - y[I]) / C[I] - (y[I] - y[I - 1]) / C[I - 1]);
  }
  B[2] = B[2] + ALAMDA * C[1];
  B[NM1] = B[NM1] + ALAMDA * C[NM1];
  for (int I = 3; I <= NM1; ++I) {
    double T = A[I] / B[I - 1];
    B[I] = B[I] - T * C[I - 1];
    R[I] = R[I] - T * R[I - 1];
  }
  fdp[NM1] = R[NM1] / B[NM1];
  for (int I = 2; I <= NM2; ++I) {
    int NMI = N - I;
    fdp[NMI] = (R[NMI] - C[NMI] * fdp[NMI + 1]) / B[NMI];
  }
  fdp[1] = ALAMDA * fdp[2];
  fdp[N] = ALAMDA * fdp[NM1];
  free(A); free(B); free(C); free(R);
}

/* BISECT, shoot.f:1018-1034, with IOLD resolved to 1 (Appendix D-1): returns the
 * 1-based I with X(I) <= xx < X(I+1), clamped to [1, N-1]; the two exact-knot
 * shortcuts of SPEVAL/SPDER (shoot.f:990-996, 1070-1076) are applied first. */
static int osr_find_bisect(const double *X0, int N, double xx) {
  if (xx == X0[0]) return 1;
  if (xx == X0[N - 1]) return N;
  int il = 0, ir = N - 1;
  while (ir - il > 1) {
    int im = (ir + il) >> 1;
    if (X0[im] > xx) ir = im; else il = im; /* X(im+1) 1-based == X0[im] */
  }
  return il + 1;
}

/* Linear search of orrsom.f:885-908 / orrspace.f:906-929 ("DO 1 I=1,NM1; IF
 * (XX.LE.X(I+1)) GO TO 10"): first I with xx <= X(I+1); falls through with
 * I = N (DO-loop exit value NM1+1) when xx > X(N). */
static int osr_find_linear(const double *X0, int N, double xx) {
  int NM1 = N - 1;
  for (int I = 1; I <= NM1; ++I)
    if (xx <= X0[I]) return I; /* X(I+1) == X0[I] */
  return N;
}

/* SPDER evaluation part, shoot.f:1090-1099 (F and FPP only; FP is unused on
 * the hot path).  I is 1-based; arrays have a zero pad at index N. */
static void osr_spder_at(const double *X0, const double *Y0, const double *F0,
                         int I, double XX, double *F, double *FPP) {
  const double *X = X0 - 1, *Y = Y0 - 1, *FDP = F0 - 1;
  double DXM = XX - X[I];
  double DXP = X[I + 1] - XX;
  double DEL = X[I + 1] - X[I];
  *F = FDP[I] * DXP * (DXP * DXP / DEL - DEL) / 6.0 +
       FDP[I + 1] * DXM * (DXM * DXM / DEL - DEL) / 6.0 +
       Y[I] * DXP / DEL + Y[I + 1] * DXM / DEL;
  *FPP = FDP[I] * DXP / DEL + FDP[I + 1] * DXM / DEL;
}

/* SPEVAL evaluation part, shoot.f:1008-1013. */
static double osr_speval_at(const double *X0, const double *Y0, const double *F0,
                            int I, double XX) {
  const double *X = X0 - 1, *Y = Y0 - 1, *FDP = F0 - 1;
  double DXM = XX - X[I];
  double DXP = X[I + 1] - XX;
  double DEL = X[I + 1] - X[I];
  return FDP[I] * DXP * (DXP * DXP / DEL - DEL) / 6.0 +
         FDP[I + 1] * DXM * (DXM * DXM / DEL - DEL) / 6.0 +
         Y[I] * DXP / DEL + Y[I + 1] * DXM / DEL;
}

/* ---------------------------------------------------------------------- */
/* Blasius / Falkner-Skan profile.  BLASIUS contebl.f:536-562.              */
/* ---------------------------------------------------------------------- */
static void osr_blasius(double beta, const double *xi, double *f) {
  f[0] = xi[1];
  f[1] = xi[2];
  f[2] = -xi[0] * xi[2] - beta * (1.0 - xi[1] * xi[1]);
}

/* SRK4, shoot.f:824-858 */
static void osr_srk4(double beta, const double *yo, double *yf, double h) {
  double f[3], k1[3], k2[3], k3[3], k4[3], q[3];
  osr_blasius(beta, yo, f);
  for (int j = 0; j < 3; ++j) { k1[j] = h * f[j]; q[j] = yo[j] + 0.5 * k1[j]; }
  osr_blasius(beta, q, f);
  for (int j = 0; j < 3; ++j) { k2[j] = h * f[j]; q[j] = yo[j] + 0.5 * k2[j]; }
  osr_blasius(beta, q, f);
  for (int j = 0; j < 3; ++j) { k3[j] = h * f[j]; q[j] = yo[j] + k3[j]; }
  osr_blasius(beta, q, f);
  for (int j = 0; j < 3; ++j) {
    k4[j] = h * f[j];
    yf[j] = yo[j] + k1[j] / 6.0 + (k2[j] + k3[j]) / 3.0 + k4[j] / 6.0;
  }
}

/* SRKCK45, shoot.f:1292-1364 (fixed step; the embedded error is discarded) */
static void osr_srkck45(double beta, const double *yo, double *yf, double h) {
  double yk[6][3], yt[3];
  for (int m = 1; m <= 6; ++m) {
    for (int n = 0; n < 3; ++n) yt[n] = yo[n];
    for (int k = 1; k <= m - 1; ++k)
      for (int n = 0; n < 3; ++n) yt[n] = yt[n] + CK_B(m, k) * yk[k - 1][n];
    osr_blasius(beta, yt, yk[m - 1]);
    for (int n = 0; n < 3; ++n) yk[m - 1][n] = h * yk[m - 1][n];
  }
  for (int n = 0; n < 3; ++n) yf[n] = yo[n];
  for (int k = 1; k <= 6; ++k)
    for (int n = 0; n < 3; ++n) yf[n] = yf[n] + ck_c[k - 1] * yk[k - 1][n];
}

/*
 * SOLVE_BL.  variant: OSR_FLOW_CONTEBL (contebl.f:386-533: integrator as
 * built, d2U := Uspl, d2Uspl = SPLINE(d2U)); OSR_FLOW_ORRSOM (orrsom.f:687-804,
 * RK4 only, same spline-of-spline); OSR_FLOW_ORRSPACE (orrspace.f:697-817, RK4
 * only, d2U keeps the 2nd-order finite difference).
 * All output arrays must hold nbl+2 doubles (last one is set to 0).
 * Returns the number of secant passes; *f2p_last = f''(0) of the accepted pass.
 */
int osr_solve_bl(int variant, int integrator, int nbl, double ymin, double ymax,
                 double f2p, double beta, double *ydat, double *U, double *Uspl,
                 double *d2U, double *d2Uspl, double *f2p_last) {
  double h = (ymax - ymin) / (double)nbl;
  double(*xi)[3] = (double(*)[3])calloc((size_t)nbl + 2, sizeof(double[3]));
  xi[0][0] = 0.0; xi[0][1] = 0.0; xi[0][2] = f2p;
  double err = 1.0, xi3old = 0.0, xi2old = 0.0, used = f2p;
  int p = 1;
  if (variant != OSR_FLOW_CONTEBL) integrator = OSR_INT_RK4;
  while (fabs(err) > 1.e-10) {
    used = xi[0][2];
    for (int i = 1; i <= nbl; ++i) {
      if (integrator == OSR_INT_CK45) osr_srkck45(beta, xi[i - 1], xi[i], h);
      else osr_srk4(beta, xi[i - 1], xi[i], h);
    }
    if (p == 1) {
      xi3old = xi[0][2];
      xi[0][2] = xi[0][2] * .99;
    } else {
      double xi3temp = xi[0][2];
      xi[0][2] = xi[0][2] + ((xi3old - xi[0][2]) / (xi2old - xi[nbl][1])) *
                                (1.0 - xi[nbl][1]);
      xi3old = xi3temp;
    }
    p = p + 1;
    xi2old = xi[nbl][1];
    err = 1.0 - xi2old;
    if (p > 200) break; /* the reference would loop forever; give up */
  }
  if (f2p_last) *f2p_last = used;
  for (int i = 0; i <= nbl; ++i) U[i] = xi[i][1];
  for (int i = 1; i <= nbl - 1; ++i)
    d2U[i] = (U[i + 1] - 2 * U[i] + U[i - 1]) / (h * h);
  d2U[0] = (-U[4] + 4 * U[3] - 5 * U[2] + 2 * U[1]) / (h * h);
  d2U[nbl] = (2 * U[nbl] - 5 * U[nbl - 1] + 4 * U[nbl - 2] - U[nbl - 3]) / (h * h);
  for (int i = 0; i <= nbl; ++i) ydat[i] = ymin + i * h;
  osr_spline(nbl + 1, ydat, U, Uspl);
  if (variant != OSR_FLOW_ORRSPACE)
    for (int i = 0; i <= nbl; ++i) d2U[i] = Uspl[i];
  osr_spline(nbl + 1, ydat, d2U, d2Uspl);
  ydat[nbl + 1] = U[nbl + 1] = Uspl[nbl + 1] = d2U[nbl + 1] = d2Uspl[nbl + 1] = 0.0;
  free(xi);
  return p - 1;
}

/* ---------------------------------------------------------------------- */
/* Shooting                                                                 */
/* ---------------------------------------------------------------------- */
typedef struct {
  int flow;       /* OSR_FLOW_*            */
  int integrator; /* OSR_INT_*             */
  int nstep;
  double testalpha; /* contebl/orrsom: degrees; orrspace: cosine bound; conte: ignored (10 deg) */
  double to, tf;    /* ymin, ymax          */
  /* profile (BL flows); arrays nbl+2 long, zero padded */
  int nbl;
  const double *ydat, *U, *Uspl, *d2U, *d2Uspl;
  /* physics, in the solver's units (after the driver's rescale) */
  zc alpha; /* contebl/conte/orrsom: given; orrspace: the unknown */
  zc c;     /* temporal: the unknown */
  zc omega; /* orrspace: given */
  double Re;
} osr_state;

/* mean flow at t for the RHS */
static void osr_meanflow(const osr_state *s, double t, double *UU, double *d2UU) {
  switch (s->flow) {
    case OSR_FLOW_CONTE: /* CHANNEL, conte.f:311-323 */
      *UU = 1.0 - t * t;
      *d2UU = -2.0;
      break;
    case OSR_FLOW_CONTEBL: { /* SPDER on (U,Uspl), contebl.f:330 */
      int I = osr_find_bisect(s->ydat, s->nbl + 1, t);
      osr_spder_at(s->ydat, s->U, s->Uspl, I, t, UU, d2UU);
      break;
    }
    default: { /* orrsom.f:627-628, orrspace.f:634-635: two SPEVALs, linear search */
      int I = osr_find_linear(s->ydat, s->nbl + 1, t);
      *UU = osr_speval_at(s->ydat, s->U, s->Uspl, I, t);
      *d2UU = osr_speval_at(s->ydat, s->d2U, s->d2Uspl, I, t);
      break;
    }
  }
}

/* OSHOMO / FHOMO */
static void osr_rhs(const osr_state *s, const zc *yo, double t, zc *yf) {
  double UU, d2UU;
  osr_meanflow(s, t, &UU, &d2UU);
  const zc alpha = s->alpha;
  const double Re = s->Re;
  const zc I1 = CMPLX(0.0, 1.0);
  yf[0] = yo[1];
  yf[1] = yo[2];
  yf[2] = yo[3];
  zc a2 = alpha * alpha;
  zc a4 = a2 * a2;
  if (s->flow == OSR_FLOW_CONTEBL || s->flow == OSR_FLOW_CONTE) {
    /* contebl.f:335-337, conte.f:205-207 */
    zc c = s->c;
    yf[3] = 2.0 * a2 * yo[2] - a4 * yo[0] +
            I1 * alpha * Re * ((UU - c) * (yo[2] - a2 * yo[0]) - d2UU * yo[0]);
  } else if (s->flow == OSR_FLOW_ORRSOM) {
    /* orrsom.f:637-639, literally (Appendix D-14) */
    zc c = s->c;
    yf[3] = (1.0 / alpha / Re * (2.0 * a2 * yo[2] - a4 * yo[0]) +
             I1 * ((UU - c) * (yo[2] - a2 * yo[0]) - d2UU * yo[0])) *
            alpha * Re;
  } else {
    /* orrspace.f:645-647: c = omega/alpha */
    zc c = s->omega / alpha;
    yf[3] = 2.0 * a2 * yo[2] - a4 * yo[0] +
            I1 * alpha * Re * ((UU - c) * (yo[2] - a2 * yo[0]) - d2UU * yo[0]);
  }
}

/* CRK4, shoot.f:869-903 */
static void osr_crk4(const osr_state *s, const zc *yo, zc *yf, double to, double h) {
  zc f[4], k1[4], k2[4], k3[4], k4[4], q[4];
  osr_rhs(s, yo, to, f);
  for (int j = 0; j < 4; ++j) { k1[j] = h * f[j]; q[j] = yo[j] + 0.5 * k1[j]; }
  osr_rhs(s, q, to + 0.5 * h, f);
  for (int j = 0; j < 4; ++j) { k2[j] = h * f[j]; q[j] = yo[j] + 0.5 * k2[j]; }
  osr_rhs(s, q, to + 0.5 * h, f);
  for (int j = 0; j < 4; ++j) { k3[j] = h * f[j]; q[j] = yo[j] + k3[j]; }
  osr_rhs(s, q, to + h, f);
  for (int j = 0; j < 4; ++j) {
    k4[j] = h * f[j];
    yf[j] = yo[j] + k1[j] / 6.0 + (k2[j] + k3[j]) / 3.0 + k4[j] / 6.0;
  }
}

/* CRKCK45, shoot.f:1367-1439 (fixed step) */
static void osr_crkck45(const osr_state *s, const zc *yo, zc *yf, double to, double h) {
  zc yk[6][4], yt[4];
  for (int m = 1; m <= 6; ++m) {
    double t = to + ck_a[m - 1] * h;
    for (int n = 0; n < 4; ++n) yt[n] = yo[n];
    for (int k = 1; k <= m - 1; ++k)
      for (int n = 0; n < 4; ++n) yt[n] = yt[n] + CK_B(m, k) * yk[k - 1][n];
    osr_rhs(s, yt, t, yk[m - 1]);
    for (int n = 0; n < 4; ++n) yk[m - 1][n] = h * yk[m - 1][n];
  }
  for (int n = 0; n < 4; ++n) yf[n] = yo[n];
  for (int k = 1; k <= 6; ++k)
    for (int n = 0; n < 4; ++n) yf[n] = yf[n] + ck_c[k - 1] * yk[k - 1][n];
}

static void osr_step(const osr_state *s, const zc *yo, zc *yf, double to, double h) {
  if (s->integrator == OSR_INT_CK45) osr_crkck45(s, yo, yf, to, h);
  else osr_crk4(s, yo, yf, to, h);
}

/* INPROD: Hermitian contebl.f:238-241 (conte, orrsom same); analytic orrspace.f:597-601 */
static zc osr_inprod(const osr_state *s, const zc *v1, const zc *v2) {
  zc r = 0.0;
  if (s->flow == OSR_FLOW_ORRSPACE)
    for (int i = 0; i < 4; ++i) r = r + v1[i] * v2[i];
  else
    for (int i = 0; i < 4; ++i) r = r + v1[i] * conj(v2[i]);
  return r;
}

/* Gram-Schmidt of the two columns of Uk (shoot.f:490-515 / 621-639) and the P
 * matrix (shoot.f:645-657).  z overwrites Uk.  P is [2][2] as P(j,i). */
static void osr_gs(const osr_state *s, zc Uk[2][4], zc P[2][2]) {
  zc z[2][4], eta[4], w[2];
  w[0] = m_csqrt(osr_inprod(s, Uk[0], Uk[0]));
  for (int i = 0; i < 4; ++i) z[0][i] = Uk[0][i] / w[0];
  for (int i = 0; i < 4; ++i) eta[i] = Uk[1][i];
  for (int i = 0; i < 4; ++i) eta[i] = eta[i] - osr_inprod(s, Uk[1], z[0]) * z[0][i];
  w[1] = m_csqrt(osr_inprod(s, eta, eta));
  for (int i = 0; i < 4; ++i) z[1][i] = eta[i] / w[1];
  if (P) {
    P[0][0] = 1.0 / w[0];
    P[0][1] = 0.0; /* never referenced: P is lower triangular */
    P[1][1] = 1.0 / w[1];
    P[1][0] = 0.0;
    P[1][0] = P[1][0] - osr_inprod(s, Uk[1], z[0]) / w[1] * P[0][0];
  }
  for (int i = 0; i < 4; ++i) { Uk[0][i] = z[0][i]; Uk[1][i] = z[1][i]; }
}

typedef struct {
  int passes;     /* number of integrations performed                      */
  int qend;       /* normalisations in the last pass                       */
  int status;     /* 0 converged by tolerance, 1 maxcount, 2 non-finite    */
  double eig[2];  /* ctemp: the last value integrated (reported)           */
  double next[2]; /* the last update (c after the loop)                    */
  double abserr;  /* |err| of the last pass                                */
  double absdc;   /* |c - cm1| at exit                                     */
} osr_result;

/*
 * The eigenvalue iteration.  history (optional) receives 4 doubles per pass:
 * contebl/orrsom/orrspace: ctemp, err (shoot.f:745-746); conte: the UPDATED c,
 * err (shoot.f:322).  eigfun (optional): (nstep+1)*4 complex values y(i,k),
 * k = 0..nstep, UN-normalised, plus vmax (divide to get fort.11-14).
 */
int osr_shoot(osr_state *s, osr_result *res, double *history, int history_cap,
              double *eigfun_y, double *eigfun_vmax) {
  const int nstep = s->nstep;
  const int flow = s->flow;
  const int bl = (flow != OSR_FLOW_CONTE);
  const double to = s->to, tf = s->tf;
  double errtol, eigtol, pi;
  int maxcount;
  switch (flow) {
    case OSR_FLOW_CONTEBL: errtol = 1.0e-14; eigtol = 1.0e-14; maxcount = 20; break; /* shoot.f:461-463 */
    case OSR_FLOW_CONTE: errtol = 1.0e-15; eigtol = 1.0e-15; maxcount = 20; break;   /* shoot.f:53-55 */
    case OSR_FLOW_ORRSOM: errtol = 1.0e-8; eigtol = 1.0e-12; maxcount = 20; break;   /* orrsom.f:203-204 */
    default: errtol = 1.0e-8; eigtol = 1.0e-12; maxcount = 50; break;                /* orrspace.f:277-278 */
  }
  if (flow == OSR_FLOW_ORRSPACE) pi = 3.14159265358979323846; /* orrspace.f:267 */
  else pi = acos(-1.0);                                        /* shoot.f:467, orrsom.f:194 */
  const double h = (tf - to) / nstep;

  /* trajectory storage (the reference keeps it even when eigfun is false) */
  zc(*U)[2][4] = (zc(*)[2][4])calloc((size_t)nstep + 1, sizeof(zc[2][4]));
  zc(*P)[2][2] = (zc(*)[2][2])calloc((size_t)nstep + 1, sizeof(zc[2][2]));
  zc(*B)[2] = (zc(*)[2])calloc((size_t)nstep + 1, sizeof(zc[2]));
  double *tq = (double *)calloc((size_t)nstep + 1, sizeof(double));

  zc *ev = (flow == OSR_FLOW_ORRSPACE) ? &s->alpha : &s->c; /* the unknown */
  int icount = 1, qend = 0, q = 0;
  zc err = 1.0, ctemp = *ev, cm1, cm2 = 0, errm1 = 0, errm2 = 0;
  int cm1_defined;
  if (flow == OSR_FLOW_CONTEBL || flow == OSR_FLOW_CONTE) { cm1 = *ev + 1.0; cm1_defined = 1; }
  else { cm1 = 0; cm1_defined = 0; } /* D-3: uninitialised => first pass always runs */
  int status = 1;

  for (;;) {
    double dc = cm1_defined ? m_cabs(*ev - cm1) : 1.0;
    int go;
    if (flow == OSR_FLOW_CONTEBL)
      go = ((m_cabs(err) >= errtol) || (dc >= eigtol)) && (icount <= maxcount);
    else
      go = (m_cabs(err) >= errtol) && (icount <= maxcount) && (dc >= eigtol);
    if (!go) {
      if (icount <= maxcount) status = 0;
      break;
    }
    q = 0;
    tq[0] = bl ? tf : to;
    if (bl) {
      /* BL_IC contebl.f:269-279; orrsom.f:211-221; orrspace.f:285-295 */
      zc alpha = s->alpha;
      zc cc = (flow == OSR_FLOW_ORRSPACE) ? s->omega / alpha : s->c;
      zc ea = m_cexp(-alpha * tf);
      U[0][0][0] = ea;
      U[0][0][1] = (-alpha) * ea;
      U[0][0][2] = (alpha * alpha) * ea;
      U[0][0][3] = (-(alpha * alpha * alpha)) * ea;
      zc g = m_csqrt(alpha * alpha + CMPLX(0.0, 1.0) * alpha * s->Re * (1.0 - cc));
      zc eg = m_cexp(-g * tf);
      U[0][1][0] = eg;
      U[0][1][1] = (-g) * eg;
      U[0][1][2] = (g * g) * eg;
      U[0][1][3] = (-(g * g * g)) * eg;
      osr_gs(s, U[0], NULL);
    } else {
      /* unit vectors e3, e4 (shoot.f:88-96, r = 2); no Gram-Schmidt at k = 0 */
      memset(U[0], 0, sizeof(U[0]));
      U[0][0][2] = 1.0;
      U[0][1][3] = 1.0;
    }
    for (int k = 1; k <= nstep; ++k) {
      double t;
      if (bl) {
        t = tf - h * k;
        for (int m = 0; m < 2; ++m) osr_step(s, U[k - 1][m], U[k][m], t + h, -h);
      } else {
        t = to + h * k;
        for (int m = 0; m < 2; ++m) osr_step(s, U[k - 1][m], U[k][m], t - h, h);
      }
      /* orthonormality test.  (mi,mj) = (1,2) and (2,1) give the same three
       * numbers with aa/bb swapped; aa*bb commutes, so one evaluation. */
      double aa = m_cabs(osr_inprod(s, U[k][0], U[k][0]));
      double bb = m_cabs(osr_inprod(s, U[k][1], U[k][1]));
      double cc12 = m_cabs(osr_inprod(s, U[k][0], U[k][1]));
      double cc21 = m_cabs(osr_inprod(s, U[k][1], U[k][0]));
      int norm = 0;
      if (flow == OSR_FLOW_CONTEBL) { /* shoot.f:593-606 */
        double ta = pi * s->testalpha / 180.0;
        double x1 = cc12 / sqrt(aa * bb), x2 = cc21 / sqrt(bb * aa);
        if (x1 > 1.0) x1 = 1.0; /* D-5 */
        if (x2 > 1.0) x2 = 1.0;
        if (m_acos(x1) < ta) norm = 1;
        if (m_acos(x2) < ta) norm = 1;
      } else if (flow == OSR_FLOW_CONTE) { /* shoot.f:64-65, 132-151 */
        double ta = 10.0 * pi / 180.0;
        double x1 = cc12 / sqrt(aa * bb), x2 = cc21 / sqrt(bb * aa);
        if (x1 > 1.0) x1 = 1.0;
        if (x2 > 1.0) x2 = 1.0;
        if (m_acos(x1) <= ta) norm = 1;
        if (m_acos(x2) <= ta) norm = 1;
      } else if (flow == OSR_FLOW_ORRSOM) { /* orrsom.f:292-302 */
        double x1 = cc12 / sqrt(aa * bb), x2 = cc21 / sqrt(bb * aa);
        if (x1 > 1.0) x1 = 1.0;
        if (x2 > 1.0) x2 = 1.0;
        if (180.0 * m_acos(x1) / pi <= s->testalpha) norm = 1;
        if (180.0 * m_acos(x2) / pi <= s->testalpha) norm = 1;
      } else { /* orrspace.f:350-358 */
        if (cc12 / sqrt(aa * bb) > s->testalpha) norm = 1;
        if (cc21 / sqrt(bb * aa) > s->testalpha) norm = 1;
      }
      if (norm || k == nstep) {
        q = q + 1;
        tq[q] = t;
        if (k == nstep) qend = q;
        osr_gs(s, U[k], P[q]);
      }
    }
    B[qend][0] = -U[nstep][1][0] / U[nstep][0][0];
    B[qend][1] = 1.0;
    ctemp = *ev;
    err = U[nstep][0][1] * B[qend][0] + U[nstep][1][1] * B[qend][1];
    if (icount == 1) {
      cm2 = *ev;
      errm2 = err;
      if (flow == OSR_FLOW_ORRSPACE) /* orrspace.f:459 */
        *ev = CMPLX(creal(*ev) * 1.01, cimag(*ev) * 1.01);
      else /* shoot.f:718 */
        *ev = CMPLX(creal(*ev) * .9999, cimag(*ev));
    } else if (icount == 2) {
      cm1 = *ev; cm1_defined = 1;
      errm1 = err;
      double fdxr = creal(err - errm2) / creal(*ev - cm2);
      double fdxi = cimag(err - errm2) / creal(*ev - cm2);
      zc fd = CMPLX(fdxr, fdxi);
      *ev = *ev - err / fd;
    } else {
      zc c = *ev;
      zc qt = (c - cm1) / (cm1 - cm2);
      zc At = qt * err - qt * (1. + qt) * errm1 + qt * qt * errm2;
      zc Bt = (2. * qt + 1.) * err - (1. + qt) * (1. + qt) * errm1 + qt * qt * errm2;
      zc Ct = (1. + qt) * err;
      zc D = m_csqrt(Bt * Bt - 4. * At * Ct);
      if (m_cabs(Bt + D) > m_cabs(Bt - D))
        *ev = ctemp - (ctemp - cm1) * 2. * Ct / (Bt + D);
      else
        *ev = ctemp - (ctemp - cm1) * 2. * Ct / (Bt - D);
      cm2 = cm1;
      cm1 = ctemp;
      errm2 = errm1;
      errm1 = err;
    }
    if (history && icount <= history_cap) {
      zc shown = (flow == OSR_FLOW_CONTE) ? *ev : ctemp;
      history[4 * (icount - 1) + 0] = creal(shown);
      history[4 * (icount - 1) + 1] = cimag(shown);
      history[4 * (icount - 1) + 2] = creal(err);
      history[4 * (icount - 1) + 3] = cimag(err);
    }
    icount = icount + 1;
    if (!(isfinite(creal(*ev)) && isfinite(cimag(*ev)) && isfinite(creal(err)) &&
          isfinite(cimag(err)))) {
      status = 2; /* the reference would have trapped (-ffpe-trap) */
      break;
    }
  }
  res->passes = icount - 1;
  res->qend = qend;
  res->status = status;
  res->eig[0] = creal(ctemp); res->eig[1] = cimag(ctemp);
  res->next[0] = creal(*ev); res->next[1] = cimag(*ev);
  res->abserr = m_cabs(err);
  res->absdc = cm1_defined ? m_cabs(*ev - cm1) : 0.0;
  if (flow == OSR_FLOW_ORRSPACE) s->alpha = ctemp; /* orrspace.f:494 */

  /* ---- second pass: eigenfunction (shoot.f:762-810; conte shoot.f:338-385) */
  if (eigfun_y && res->passes > 0) {
    zc(*y)[4] = (zc(*)[4])calloc((size_t)nstep + 2, sizeof(zc[4]));
    zc vmax = 0.0;
    q = qend;
    int k = nstep;
    for (int i = 0; i < 4; ++i) {
      y[k][i] = 0.0; /* v(i,k): D-4 / D-6 */
      for (int m = 0; m < 2; ++m) y[k][i] = y[k][i] + U[k][m][i] * B[q][m];
    }
    for (int m = 0; m < 2; ++m) {
      B[q - 1][m] = 0.0;
      for (int j = 0; j < 2; ++j) B[q - 1][m] = B[q - 1][m] + P[q][j][m] * B[q][j];
    }
    for (k = nstep - 1; k >= 0; --k) {
      double t = bl ? tf - h * k : to + h * k;
      int back = bl ? (t > tq[q - 1]) : (t < tq[q - 1]);
      if (back) {
        q = q - 1;
        for (int m = 0; m < 2; ++m) {
          B[q - 1][m] = 0.0;
          for (int j = 0; j < 2; ++j) B[q - 1][m] = B[q - 1][m] + P[q][j][m] * B[q][j];
        }
      }
      for (int i = 0; i < 4; ++i) {
        y[k][i] = 0.0;
        for (int m = 0; m < 2; ++m) y[k][i] = y[k][i] + U[k][m][i] * B[q - 1][m];
      }
      /* D-2: the DO index is left at n+1, so the scan reads y(n+1,k) == y(1,k+1) */
      if (m_cabs(y[k + 1][0]) > m_cabs(vmax)) vmax = y[k + 1][0];
    }
    for (k = 0; k <= nstep; ++k)
      for (int i = 0; i < 4; ++i) {
        eigfun_y[2 * (4 * k + i) + 0] = creal(y[k][i]);
        eigfun_y[2 * (4 * k + i) + 1] = cimag(y[k][i]);
      }
    if (eigfun_vmax) { eigfun_vmax[0] = creal(vmax); eigfun_vmax[1] = cimag(vmax); }
    free(y);
  }
  free(U); free(P); free(B); free(tq);
  return status;
}

/* ---------------------------------------------------------------------- */
/* Flat entry points for ctypes / bench (plain pointers only).             */
/* ---------------------------------------------------------------------- */

/* One temporal eigenvalue (contebl / conte / orrsom flows).
 * alpha, Re, c are in the SOLVER's units (after the driver's sqrt(2) rescale).
 * prof: 5 arrays of nbl+2 doubles [ydat,U,Uspl,d2U,d2Uspl] or NULL for conte.
 * out: [c_re, c_im, next_re, next_im, |err|, |dc|]; iout: [passes, qend, status] */
int osr_shoot_temporal(int flow, int integrator, int nstep, double testalpha,
                       double ymin, double ymax, int nbl, const double *prof,
                       double alpha_re, double alpha_im, double Re, double c_re,
                       double c_im, double *out, int *iout, double *history,
                       int history_cap, double *eigfun_y, double *eigfun_vmax) {
  osr_state s;
  memset(&s, 0, sizeof(s));
  s.flow = flow; s.integrator = integrator; s.nstep = nstep; s.testalpha = testalpha;
  s.to = ymin; s.tf = ymax; s.nbl = nbl;
  if (prof) {
    size_t n = (size_t)nbl + 2;
    s.ydat = prof; s.U = prof + n; s.Uspl = prof + 2 * n; s.d2U = prof + 3 * n; s.d2Uspl = prof + 4 * n;
  }
  s.alpha = CMPLX(alpha_re, alpha_im); s.Re = Re; s.c = CMPLX(c_re, c_im);
  osr_result r;
  int st = osr_shoot(&s, &r, history, history_cap, eigfun_y, eigfun_vmax);
  out[0] = r.eig[0]; out[1] = r.eig[1]; out[2] = r.next[0]; out[3] = r.next[1];
  out[4] = r.abserr; out[5] = r.absdc;
  iout[0] = r.passes; iout[1] = r.qend; iout[2] = r.status;
  return st;
}

/* Batch of independent temporal points (used as the CPU baseline and as the
 * parity checker on grids).  alpha[2*npts], Re[npts], c[2*npts] in/out. */
int osr_shoot_temporal_batch(int flow, int integrator, int nstep, double testalpha,
                             double ymin, double ymax, int nbl, const double *prof,
                             int64_t npts, const double *alpha, const double *Re,
                             double *c, int *passes, int *qend, int *status, double *absdc) {
  for (int64_t p = 0; p < npts; ++p) {
    double out[6];
    int iout[3];
    osr_shoot_temporal(flow, integrator, nstep, testalpha, ymin, ymax, nbl, prof,
                       alpha[2 * p], alpha[2 * p + 1], Re[p], c[2 * p], c[2 * p + 1],
                       out, iout, NULL, 0, NULL, NULL);
    c[2 * p] = out[0]; c[2 * p + 1] = out[1];
    if (passes) passes[p] = iout[0];
    if (qend) qend[p] = iout[1];
    if (status) status[p] = iout[2];
    if (absdc) absdc[p] = out[5]; /* |c_next - ctemp|: the update the loop computed but did not integrate */
  }
  return 0;
}

/* One spatial continuation line (orrspace.f:186-195): niter points in omega,
 * each starting from the previous converged alpha.  All quantities in the
 * solver's (rescaled) units.  rows: niter x [omega_re, omega_im, alpha_re,
 * alpha_im]; ipass/iqend/istat: niter each. */
int osr_shoot_spatial_line(int integrator, int nstep, double testalpha, double ymin,
                           double ymax, int nbl, const double *prof, double Re,
                           double omega_re, double omega_im, double domega,
                           double alpha_re, double alpha_im, int niter, double *rows,
                           int *ipass, int *iqend, int *istat, double *absdc) {
  osr_state s;
  memset(&s, 0, sizeof(s));
  s.flow = OSR_FLOW_ORRSPACE; s.integrator = integrator; s.nstep = nstep;
  s.testalpha = testalpha; s.to = ymin; s.tf = ymax; s.nbl = nbl;
  size_t n = (size_t)nbl + 2;
  s.ydat = prof; s.U = prof + n; s.Uspl = prof + 2 * n; s.d2U = prof + 3 * n; s.d2Uspl = prof + 4 * n;
  s.Re = Re; s.alpha = CMPLX(alpha_re, alpha_im); s.omega = CMPLX(omega_re, omega_im);
  for (int i = 0; i < niter; ++i) {
    osr_result r;
    osr_shoot(&s, &r, NULL, 0, NULL, NULL);
    rows[4 * i + 0] = creal(s.omega); rows[4 * i + 1] = cimag(s.omega);
    rows[4 * i + 2] = creal(s.alpha); rows[4 * i + 3] = cimag(s.alpha);
    if (ipass) ipass[i] = r.passes;
    if (iqend) iqend[i] = r.qend;
    if (istat) istat[i] = r.status;
    if (absdc) absdc[i] = r.absdc;
    double omegar = creal(s.omega) + domega; /* orrspace.f:192-194 */
    s.omega = CMPLX(omegar, cimag(s.omega));
  }
  return 0;
}

/* One spatial point with the eigenfunction pass (orrspace with niter = 1, eigfun = 1,
 * e.g. space.inp).  out: [alpha_re, alpha_im]; iout: [passes, qend, status]. */
int osr_shoot_spatial_point(int integrator, int nstep, double testalpha, double ymin, double ymax,
                            int nbl, const double *prof, double Re, double omega_re, double omega_im,
                            double alpha_re, double alpha_im, double *out, int *iout,
                            double *eigfun_y, double *eigfun_vmax) {
  osr_state s;
  memset(&s, 0, sizeof(s));
  s.flow = OSR_FLOW_ORRSPACE; s.integrator = integrator; s.nstep = nstep;
  s.testalpha = testalpha; s.to = ymin; s.tf = ymax; s.nbl = nbl;
  size_t n = (size_t)nbl + 2;
  s.ydat = prof; s.U = prof + n; s.Uspl = prof + 2 * n; s.d2U = prof + 3 * n; s.d2Uspl = prof + 4 * n;
  s.Re = Re; s.alpha = CMPLX(alpha_re, alpha_im); s.omega = CMPLX(omega_re, omega_im);
  osr_result r;
  int st = osr_shoot(&s, &r, NULL, 0, eigfun_y, eigfun_vmax);
  out[0] = r.eig[0]; out[1] = r.eig[1];
  iout[0] = r.passes; iout[1] = r.qend; iout[2] = r.status;
  return st;
}

/* SHIFT_BL, conteucbl.f:924-968.  in/out: 5 arrays of nbl+2 doubles as in osr_solve_bl. */
void osr_profile_shift(int nbl, double uc, const double *in, double *out) {
  const size_t n2 = (size_t)nbl + 2;
  double *ydat = out, *U = out + n2, *Uspl = out + 2 * n2, *d2U = out + 3 * n2, *d2Uspl = out + 4 * n2;
  for (int i = 0; i <= nbl; ++i) { ydat[i] = in[i]; U[i] = in[n2 + i] - uc; }
  osr_spline(nbl + 1, ydat, U, Uspl);
  for (int i = 0; i <= nbl; ++i) d2U[i] = Uspl[i];
  osr_spline(nbl + 1, ydat, d2U, d2Uspl);
  ydat[nbl + 1] = U[nbl + 1] = Uspl[nbl + 1] = d2U[nbl + 1] = d2Uspl[nbl + 1] = 0.0;
}
