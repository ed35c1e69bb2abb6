/*
 * os_chan_oracle.c -- CPU restatement of the reference's MATRIX path
 * (channel flow: orrfdchan.f, orrcolchan.f, orrncbl.f).
 *
 * TEST INFRASTRUCTURE ONLY (see os_shoot_oracle.c for the rules).
 *
 * Restated here:
 *   FiniteD1                    orrfdchan.f:1183-1215
 *   CHEBYD                      orrcolchan.f:904-944, orrncbl.f:1339-1379
 *   MAKE_CHANNEL_METRICS        orrfdchan.f:144-200  (cosine grid, m1=1, m2..4=0)
 *   INIT_CHANNEL_PROFILE        orrfdchan.f:203-270 (analytic U', U''),
 *                               orrncbl.f:488-551 (U', U'' by collocation)
 *   MAKE_DERIVATIVES            orrfdchan.f:273-350, orrcolchan.f:77-154
 *   MAKE_MATRIX                 orrfdchan.f:553-659, orrncbl.f:553-695 (+ d/dalpha)
 *   SOLVE_ORR_SOM               orrfdchan.f:688-823 (inv(B)A + ZGEEV, SORT, selection)
 *                               orrcolchan.f:356-513 (BC rows, ZGEGV, PIKSR2)
 *                               orrncbl.f:758-981 (ZGETRF/ZGETRS/ZGEEV, Newton on alpha,
 *                                 bug-compatible: VL of inv(B)A used as the adjoint, D-7)
 *
 * Third-party arithmetic: the reference links whatever OpenBLAS is installed
 * (gcc.mak:85-99, unpinned).  This oracle calls the SAME LAPACK routines
 * (ZGETRF ZGETRI ZGETRS ZGEMM ZGEEV ZGEGV ZGEMV) from the OpenBLAS bundled with
 * scipy (0.3.31.dev, LP64, symbols prefixed scipy_), opened with dlopen at run
 * time by osr_lapack_open().  The Numerical-Recipes SORT / PIKSR2 the reference
 * needs are not shipped with it (README.md:110-118); both are restated as the
 * straight-insertion sort PIKSR2 is documented to be (ascending, stable).
 *
 * Parity pinning: the reference holds no golden output for this path; the only
 * anchor is chan.inp:4 (the TS mode is sorted index 62 for N=64).  Values are
 * additionally cross-checked against the surveyor's independent numpy
 * restatement and Orszag's 1971 eigenvalue.
 */
#include <complex.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#undef I
typedef double _Complex zc;

#define OSR_DISC_FD 0       /* orrfdchan: FiniteD1, analytic U', U''            */
#define OSR_DISC_CHEB 1     /* orrcolchan: CHEBYD, analytic U', U''             */
#define OSR_DISC_CHEB_NUM 2 /* orrncbl: CHEBYD, U', U'' by collocation          */

/* ---- LAPACK from scipy's OpenBLAS ---------------------------------------- */
typedef void (*zgetrf_t)(int *, int *, zc *, int *, int *, int *);
typedef void (*zgetri_t)(int *, zc *, int *, int *, zc *, int *, int *);
typedef void (*zgetrs_t)(char *, int *, int *, zc *, int *, int *, zc *, int *, int *);
typedef void (*zgemm_t)(char *, char *, int *, int *, int *, zc *, zc *, int *, zc *, int *, zc *, zc *, int *);
typedef void (*zgemv_t)(char *, int *, int *, zc *, zc *, int *, zc *, int *, zc *, zc *, int *);
typedef void (*zgeev_t)(char *, char *, int *, zc *, int *, zc *, zc *, int *, zc *, int *, zc *, int *, double *, int *);
typedef void (*zgegv_t)(char *, char *, int *, zc *, int *, zc *, int *, zc *, zc *, zc *, int *, zc *, int *, zc *, int *, double *, int *);

static struct {
  void *h;
  zgetrf_t zgetrf; zgetri_t zgetri; zgetrs_t zgetrs; zgemm_t zgemm; zgemv_t zgemv;
  zgeev_t zgeev; zgegv_t zgegv;
} L;

int osr_lapack_open(const char *path) {
  if (L.h) return 0;
  L.h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!L.h) { fprintf(stderr, "osr_lapack_open: %s\n", dlerror()); return -1; }
#define SYM(n) do { *(void **)(&L.n) = dlsym(L.h, "scipy_" #n "_"); if (!L.n) *(void **)(&L.n) = dlsym(L.h, #n "_"); if (!L.n) { fprintf(stderr, "missing " #n "\n"); return -2; } } while (0)
  SYM(zgetrf); SYM(zgetri); SYM(zgetrs); SYM(zgemm); SYM(zgemv); SYM(zgeev); SYM(zgegv);
#undef SYM
  return 0;
}

/* ---- grid, profile, derivative matrices ---------------------------------- */
typedef struct {
  int n;            /* N: grid is 0..N                                         */
  int disc;
  double *eta, *u, *d1u, *d2u;          /* N+1                                  */
  double *D1hat, *D1, *D2, *D3, *D4;    /* (N+1)^2, element (i,j) at [i*(N+1)+j] */
  double *m1, *m2, *m3, *m4;            /* N+1: mapping metrics (channel: 1,0,0,0) */
} osr_chan;

#define AT(M, i, j) (M)[(size_t)(i) * (size_t)(n + 1) + (size_t)(j)]

/* FiniteD1 orrfdchan.f:1183-1215 */
static void finite_d1(double *D, int n, const double *eta) {
  memset(D, 0, sizeof(double) * (size_t)(n + 1) * (n + 1));
  AT(D, 0, 0) = 1. / (eta[0] - eta[1]);
  AT(D, 0, 1) = -1. / (eta[0] - eta[1]);
  AT(D, n, n - 1) = 1. / (eta[n - 1] - eta[n]);
  AT(D, n, n) = -1. / (eta[n - 1] - eta[n]);
  for (int i = 1; i <= n - 1; ++i) {
    AT(D, i, i - 1) = 1. / (eta[i - 1] - eta[i + 1]);
    AT(D, i, i + 1) = -1. / (eta[i - 1] - eta[i + 1]);
  }
}

/* CHEBYD orrcolchan.f:904-944 */
static void chebyd(double *D, int n) {
  const double PI = acos(-1.0);
  double *C = (double *)malloc(sizeof(double) * (n + 1));
  for (int i = 1; i <= n - 1; ++i) C[i] = 1.0;
  C[0] = 2.0; C[n] = 2.0;
  const double fn = (double)n;
  for (int j = 0; j <= n; ++j)
    for (int k = 0; k <= n; ++k) {
      if (j == 0 && k == 0) AT(D, j, k) = (2.0 * (fn * fn) + 1.0) / 6.0;
      else if (j == n && k == n) AT(D, j, k) = -1.0 * (2. * (fn * fn) + 1.0) / 6.0;
      else if (j == k) {
        double cj = cos((double)j * PI / fn);
        AT(D, j, k) = -1.0 * cj / (2. * (1. - cj * cj));
      } else {
        double sgn = ((j + k) % 2 == 0) ? 1.0 : -1.0; /* (-1.0)**(J+K) */
        AT(D, j, k) = C[j] * sgn / (C[k] * (cos((double)j * PI / fn) - cos((double)k * PI / fn)));
      }
    }
  free(C);
}

/* C(i,j) = sum_k A(i,k) B(k,j), accumulated in k order (orrfdchan.f:318-342) */
static void matmul_ref(double *Cm, const double *A, const double *B, int n) {
  for (int i = 0; i <= n; ++i)
    for (int j = 0; j <= n; ++j) {
      double s = 0.0;
      for (int k = 0; k <= n; ++k) s = s + AT(A, i, k) * AT(B, k, j);
      AT(Cm, i, j) = s;
    }
}

osr_chan *osr_chan_create(int disc, int n) {
  osr_chan *c = (osr_chan *)calloc(1, sizeof(osr_chan));
  size_t n1 = (size_t)n + 1, nn = n1 * n1;
  c->n = n; c->disc = disc;
  c->eta = (double *)calloc(n1, sizeof(double)); c->u = (double *)calloc(n1, sizeof(double));
  c->d1u = (double *)calloc(n1, sizeof(double)); c->d2u = (double *)calloc(n1, sizeof(double));
  c->D1hat = (double *)calloc(nn, sizeof(double)); c->D1 = (double *)calloc(nn, sizeof(double));
  c->D2 = (double *)calloc(nn, sizeof(double)); c->D3 = (double *)calloc(nn, sizeof(double));
  c->D4 = (double *)calloc(nn, sizeof(double));
  c->m1 = (double *)calloc(n1, sizeof(double)); c->m2 = (double *)calloc(n1, sizeof(double));
  c->m3 = (double *)calloc(n1, sizeof(double)); c->m4 = (double *)calloc(n1, sizeof(double));
  for (int i = 0; i <= n; ++i) c->m1[i] = 1.0;   /* orrfdchan.f:185-190: m1 = 1, m2 = m3 = m4 = 0 */
  /* MAKE_CHANNEL_METRICS orrfdchan.f:177-183 */
  const double pi = acos(-1.0);
  const double dth = pi / (double)n;
  for (int i = 0; i <= n; ++i) c->eta[i] = cos((double)i * dth);
  /* MAKE_DERIVATIVES */
  if (disc == OSR_DISC_FD) { finite_d1(c->D1hat, n, c->eta); finite_d1(c->D1, n, c->eta); }
  else { chebyd(c->D1hat, n); chebyd(c->D1, n); }
  for (int i = 0; i <= n; ++i) { AT(c->D1, 0, i) = 0.0; AT(c->D1, n, i) = 0.0; }
  matmul_ref(c->D2, c->D1hat, c->D1, n);
  matmul_ref(c->D3, c->D1hat, c->D2, n);
  matmul_ref(c->D4, c->D1hat, c->D3, n);
  /* INIT_CHANNEL_PROFILE */
  for (int i = 0; i <= n; ++i) {
    c->u[i] = (1. - c->eta[i] * c->eta[i]);
    c->d1u[i] = -2.0 * c->eta[i];
    c->d2u[i] = -2.0;
  }
  if (disc == OSR_DISC_CHEB_NUM) { /* orrncbl.f:520-545 */
    double *D2hat = (double *)calloc(nn, sizeof(double));
    matmul_ref(D2hat, c->D1hat, c->D1hat, n);
    for (int i = 0; i <= n; ++i) {
      double a = 0.0, b = 0.0;
      for (int k = 0; k <= n; ++k) { a = a + AT(c->D1hat, i, k) * c->u[k]; b = b + AT(D2hat, i, k) * c->u[k]; }
      c->d1u[i] = a; c->d2u[i] = b;
    }
    free(D2hat);
  }
  return c;
}

void osr_chan_destroy(osr_chan *c) {
  if (!c) return;
  free(c->eta); free(c->u); free(c->d1u); free(c->d2u);
  free(c->D1hat); free(c->D1); free(c->D2); free(c->D3); free(c->D4);
  free(c->m1); free(c->m2); free(c->m3); free(c->m4);
  free(c);
}

/*
 * Mapped boundary layer (orrcolbl.f / orrfdbl.f): y = Lmap (1+eta)/(1-eta).
 *   MAKE_DERIVATIVES   orrcolbl.f:161-242 (CHEBYD on the cosine grid), orrfdbl.f:145-222 (FiniteD1 on the
 *                      UNIFORM grid eta_i = 1 - 2 i/N, orrfdbl.f:258-265)
 *   MAKE_BL_METRICS    orrcolbl.f:245-311 (m1..m4)
 *   INIT_BL_PROFILE    orrcolbl.f:334-441: the reference interpolates an external profile file onto the grid;
 *                      here the caller supplies U at the grid points (u[0..merge] = 1 exactly, the free
 *                      stream).  d1u = D1hat u, d2u = D1hat D1hat u; both zeroed for j <= merge-2 (:419-427).
 * disc: OSR_DISC_CHEB or OSR_DISC_FD.  The reference cannot be run (IMSL, missing profile file): parity of
 * this operator is unpinned; the physics check is the shooting eigenvalue of the same flow.
 */
osr_chan *osr_chan_create_bl(int disc, int n, double lmap, const double *U, int merge) {
  osr_chan *c = osr_chan_create(disc, n);
  const size_t n1 = (size_t)n + 1, nn = n1 * n1;
  if (disc == OSR_DISC_FD) {
    const double deta = 2. / (double)n;
    for (int i = 0; i <= n; ++i) c->eta[i] = 1. - (double)i * deta;
    finite_d1(c->D1hat, n, c->eta); finite_d1(c->D1, n, c->eta);
    for (int i = 0; i <= n; ++i) { AT(c->D1, 0, i) = 0.0; AT(c->D1, n, i) = 0.0; }
    matmul_ref(c->D2, c->D1hat, c->D1, n);
    matmul_ref(c->D3, c->D1hat, c->D2, n);
    matmul_ref(c->D4, c->D1hat, c->D3, n);
  }
  for (int i = 0; i <= n; ++i) {
    const double e1 = c->eta[i] - 1.;
    c->m1[i] = (e1 * e1) / (2. * lmap);
    c->m2[i] = (e1 * e1 * e1) / (2. * (lmap * lmap));
    c->m3[i] = 3. * (e1 * e1 * e1 * e1) / (4. * (lmap * lmap * lmap));
    c->m4[i] = 3. * (e1 * e1 * e1 * e1 * e1) / (2. * (lmap * lmap * lmap * lmap));
    c->u[i] = (i <= merge) ? 1.0 : U[i];
  }
  double *D2hat = (double *)calloc(nn, sizeof(double));
  matmul_ref(D2hat, c->D1hat, c->D1hat, n);
  for (int i = 0; i <= n; ++i) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k <= n; ++k) { a = a + AT(c->D1hat, i, k) * c->u[k]; b = b + AT(D2hat, i, k) * c->u[k]; }
    c->d1u[i] = a; c->d2u[i] = b;
  }
  for (int j = merge - 2; j >= 0; --j) { c->d1u[j] = 0.0; c->d2u[j] = 0.0; }
  free(D2hat);
  return c;
}

/* copy tables out for the tests: which = 0 eta, 1 u, 2 d1u, 3 d2u (N+1); 4 D2, 5 D4 ((N+1)^2) */
void osr_chan_get(const osr_chan *c, int which, double *out) {
  size_t n1 = (size_t)c->n + 1;
  const double *src[6] = {c->eta, c->u, c->d1u, c->d2u, c->D2, c->D4};
  memcpy(out, src[which], sizeof(double) * (which < 4 ? n1 : n1 * n1));
}

/*
 * MAKE_MATRIX + the interior/BC assembly of SOLVE_ORR_SOM.  With the channel
 * metrics (m1=1, m2=m3=m4=0) A3 and A1 are exact zeros and are added as such.
 * Output: column-major m x m matrices (m = N-1 interior, or N+1 with BC rows).
 */
static void assemble(const osr_chan *c, zc alpha, double Re, int full, zc *A, zc *B, zc *dA, zc *dB) {
  const int n = c->n;
  const int m = full ? n + 1 : n - 1;
  const int off = full ? 0 : 1;
  const zc iu = CMPLX(0., 1.);
  const zc a2 = alpha * alpha, a3 = a2 * alpha, a4 = a2 * a2;
  for (int i = 0; i < m; ++i) {
    const int gi = i + off;
    const double m1 = c->m1[gi], m2 = c->m2[gi], m3 = c->m3[gi], m4 = c->m4[gi];
    const zc w4 = m1 * m1 * m1 * m1;
    const zc w3 = 6. * (m1 * m1) * m2;
    const zc w2 = 3. * (m2 * m2) + 4. * m1 * m3 - 2. * a2 * (m1 * m1) - iu * alpha * Re * c->u[gi] * (m1 * m1);
    const zc w1 = m4 - 2. * a2 * m2 - iu * alpha * Re * c->u[gi] * m2;
    const zc w0 = a4 + iu * a3 * Re * c->u[gi] + iu * alpha * Re * (c->d2u[gi] * (m1 * m1) + c->d1u[gi] * m2);
    const zc b2 = -1.0 * iu * Re * (m1 * m1);
    const zc b1 = -1.0 * iu * Re * m2;
    const zc b0 = iu * a2 * Re;
    const zc d2 = -4. * alpha * (m1 * m1) - iu * Re * c->u[gi] * (m1 * m1);
    const zc d1 = -4. * alpha * m2 - iu * Re * c->u[gi] * m2;
    const zc d0 = 4. * a3 + 3. * iu * a2 * Re * c->u[gi] + iu * Re * (c->d2u[gi] * (m1 * m1) + c->d1u[gi] * m2);
    const zc db0 = iu * 2. * alpha * Re;
    for (int j = 0; j < m; ++j) {
      const int gj = j + off;
      const double id = (gi == gj) ? 1.0 : 0.0;
      zc A4 = w4 * AT(c->D4, gi, gj), A3 = w3 * AT(c->D3, gi, gj), A2 = w2 * AT(c->D2, gi, gj),
         A1 = w1 * AT(c->D1, gi, gj), A0 = w0 * id;
      zc B2 = b2 * AT(c->D2, gi, gj), B1 = b1 * AT(c->D1, gi, gj), B0 = b0 * id;
      A[(size_t)j * m + i] = A4 + A3 + A2 + A1 + A0;
      B[(size_t)j * m + i] = B2 + B1 + B0;
      if (dA) dA[(size_t)j * m + i] = d2 * AT(c->D2, gi, gj) + d1 * AT(c->D1, gi, gj) + d0 * id;
      if (dB) dB[(size_t)j * m + i] = db0 * id;
    }
  }
  if (full) { /* orrcolchan.f:424-435 */
    for (int j = 0; j <= n; ++j) {
      B[(size_t)j * m + 0] = 0; B[(size_t)0 * m + j] = 0; B[(size_t)j * m + n] = 0; B[(size_t)n * m + j] = 0;
      A[(size_t)j * m + 0] = 0; A[(size_t)0 * m + j] = 0; A[(size_t)j * m + n] = 0; A[(size_t)n * m + j] = 0;
    }
    A[0] = 1.0; A[(size_t)n * m + n] = 1.0;
  }
}

/* PIKSR2 / SORT: straight insertion, ascending in arr, brr carried along */
static void piksr2(int nn, double *arr, int *brr) {
  for (int j = 1; j < nn; ++j) {
    double a = arr[j]; int b = brr[j]; int i = j - 1;
    while (i >= 0 && arr[i] > a) { arr[i + 1] = arr[i]; brr[i + 1] = brr[i]; --i; }
    arr[i + 1] = a; brr[i + 1] = b;
  }
}

/* inv(B) A + ZGEEV as orrfdchan.f:792-797; returns eval, VL, VR (column-major m x m) */
static int eig_invBA(int m, zc *A, zc *B, zc *eval, zc *vl, zc *vr, int use_getrs) {
  int info = 0, lwork = 64 * m;
  int *ipiv = (int *)malloc(sizeof(int) * m);
  zc *work = (zc *)malloc(sizeof(zc) * lwork);
  double *rwork = (double *)malloc(sizeof(double) * 8 * m);
  zc *T1 = (zc *)malloc(sizeof(zc) * (size_t)m * m), *T4 = (zc *)malloc(sizeof(zc) * (size_t)m * m);
  memcpy(T1, B, sizeof(zc) * (size_t)m * m);
  L.zgetrf(&m, &m, T1, &m, ipiv, &info);
  if (use_getrs) { /* orrncbl.f:864-865: A <- inv(B) A in place */
    memcpy(T4, A, sizeof(zc) * (size_t)m * m);
    L.zgetrs("N", &m, &m, T1, &m, ipiv, T4, &m, &info);
  } else { /* orrfdchan.f:794-795 */
    L.zgetri(&m, T1, &m, ipiv, work, &lwork, &info);
    zc one = 1.0, zero = 0.0;
    L.zgemm("N", "N", &m, &m, &m, &one, T1, &m, A, &m, &zero, T4, &m);
  }
  L.zgeev("V", "V", &m, T4, &m, eval, vl, &m, vr, &m, work, &lwork, rwork, &info);
  free(ipiv); free(work); free(rwork); free(T1); free(T4);
  return info;
}

/*
 * orrfdchan SOLVE_ORR_SOM (also used for the Chebyshev discretisations so that
 * the GPU path has a like-for-like checker): eigenvalue = the largest Im w
 * with |Im w| < 10 and int(Re w) != 999 (orrfdchan.f:815-823).
 * spectrum (optional, 2*(N-1) doubles): all eigenvalues sorted by Im ascending.
 * evec (optional, 2*(N+1)): the eigenvector of the selected mode on the full
 * grid (zeros at 0 and N), normalised by its max-modulus component (orrfdchan.f:849-857).
 * avec (optional, 2*(N+1)): columns 4-5 of fort.11 as orrfdchan.f:865-871 writes them: row j holds
 * avec(j, ind(iloc)), ZGEEV's LEFT eigenvector of inv(B)A (unit 2-norm, largest component real) --
 * NOT shifted onto the grid like the right vector (the reference pairs eta(j) with interior entry j)
 * and read two entries past the N-1 that ZGEEV wrote, which are zeros of the static array.
 */
int osr_chan_solve_fdstyle(const osr_chan *c, double alpha_re, double alpha_im, double Re,
                           double *omega, double *spectrum, double *evec, double *avec) {
  const int n = c->n, m = n - 1;
  zc *A = (zc *)malloc(sizeof(zc) * (size_t)m * m), *B = (zc *)malloc(sizeof(zc) * (size_t)m * m);
  zc *eval = (zc *)malloc(sizeof(zc) * m), *vl = (zc *)malloc(sizeof(zc) * (size_t)m * m),
     *vr = (zc *)malloc(sizeof(zc) * (size_t)m * m);
  assemble(c, CMPLX(alpha_re, alpha_im), Re, 0, A, B, NULL, NULL);
  int info = eig_invBA(m, A, B, eval, vl, vr, 0);
  double *t1 = (double *)malloc(sizeof(double) * m), *t2 = (double *)malloc(sizeof(double) * m);
  int *ind = (int *)malloc(sizeof(int) * m);
  for (int i = 0; i < m; ++i) { t1[i] = creal(eval[i]); t2[i] = cimag(eval[i]); ind[i] = i; }
  piksr2(m, t2, ind);
  int first = 1, iloc = -1;
  zc ev = 0;
  for (int i = m - 1; i >= 0; --i)
    if (fabs(t2[i]) < 10.0 && first && ((int)t1[ind[i]]) != 999) {
      first = 0; ev = CMPLX(t1[ind[i]], t2[i]); iloc = i;
    }
  omega[0] = creal(ev); omega[1] = cimag(ev);
  if (spectrum) for (int i = 0; i < m; ++i) { spectrum[2 * i] = t1[ind[i]]; spectrum[2 * i + 1] = t2[i]; }
  if (evec && iloc >= 0) {
    const zc *v = vr + (size_t)ind[iloc] * m;
    zc escale = 0.0;
    for (int i = 0; i < m; ++i) if (cabs(v[i]) > cabs(escale)) escale = v[i];
    escale = 1.0 / escale;
    evec[0] = evec[1] = 0.0; evec[2 * n] = evec[2 * n + 1] = 0.0;
    for (int i = 0; i < m; ++i) { zc s = escale * v[i]; evec[2 * (i + 1)] = creal(s); evec[2 * (i + 1) + 1] = cimag(s); }
  }
  if (avec && iloc >= 0) {
    const zc *v = vl + (size_t)ind[iloc] * m;
    for (int j = 0; j <= n; ++j) {
      avec[2 * j] = j < m ? creal(v[j]) : 0.0;
      avec[2 * j + 1] = j < m ? cimag(v[j]) : 0.0;
    }
  }
  free(A); free(B); free(eval); free(vl); free(vr); free(t1); free(t2); free(ind);
  return info;
}

/*
 * orrcolchan SOLVE_ORR_SOM: (N+1)x(N+1) pencil with the BC rows, ZGEGV, eval =
 * alp/beta (0 when beta = 0), PIKSR2 by Im; spectrum: 2*(N+1) doubles sorted
 * ascending (index 0..N as the reference prints them, orrcolchan.f:505-513).
 * evec (optional): eigenvector `which` normalised by its max-modulus component.
 * avec (optional, 2*(N+1)): ZGEGV's left eigenvector of the same mode scaled by 1/conj(zdotc(avec, evec))
 * (orrcolchan.f:537-545), i.e. columns 4-5 of fort.11; dots (optional, 4 doubles): the two zdotc values the
 * reference prints (orrcolchan.f:539,546) -- the first depends on LAPACK's normalisation of avec, the second is 1.
 */
int osr_chan_solve_col(const osr_chan *c, double alpha_re, double alpha_im, double Re,
                       double *spectrum, int which, double *evec, double *avec, double *dots) {
  const int n = c->n;
  int m = n + 1, info = 0, lwork = 64 * m;
  zc *A = (zc *)malloc(sizeof(zc) * (size_t)m * m), *B = (zc *)malloc(sizeof(zc) * (size_t)m * m);
  zc *alp = (zc *)malloc(sizeof(zc) * m), *beta = (zc *)malloc(sizeof(zc) * m);
  zc *vl = (zc *)malloc(sizeof(zc) * (size_t)m * m), *vr = (zc *)malloc(sizeof(zc) * (size_t)m * m);
  zc *work = (zc *)malloc(sizeof(zc) * lwork);
  double *rwork = (double *)malloc(sizeof(double) * 8 * m);
  assemble(c, CMPLX(alpha_re, alpha_im), Re, 1, A, B, NULL, NULL);
  L.zgegv("V", "V", &m, A, &m, B, &m, alp, beta, vl, &m, vr, &m, work, &lwork, rwork, &info);
  double *t2 = (double *)malloc(sizeof(double) * m);
  int *ind = (int *)malloc(sizeof(int) * m);
  zc *eval = (zc *)malloc(sizeof(zc) * m);
  for (int i = 0; i < m; ++i) {
    eval[i] = (beta[i] != 0) ? alp[i] / beta[i] : 0.0;
    t2[i] = cimag(eval[i]); ind[i] = i;
  }
  piksr2(m, t2, ind);
  for (int i = 0; i < m; ++i) { spectrum[2 * i] = creal(eval[ind[i]]); spectrum[2 * i + 1] = t2[i]; }
  if (evec && which >= 0 && which < m) {
    const zc *v = vr + (size_t)ind[which] * m;
    zc escale = 0.0;
    for (int i = 0; i < m; ++i) if (cabs(v[i]) > cabs(escale)) escale = v[i];
    escale = 1.0 / escale;
    for (int i = 0; i < m; ++i) { zc s = escale * v[i]; evec[2 * i] = creal(s); evec[2 * i + 1] = cimag(s); }
    if (avec) {
      const zc *a = vl + (size_t)ind[which] * m;
      zc d1 = 0.0, d2 = 0.0;
      for (int i = 0; i < m; ++i) d1 += conj(a[i]) * (escale * v[i]);  /* zdotc(avec, evec) */
      const zc sc = 1.0 / conj(d1);
      for (int i = 0; i < m; ++i) {
        const zc t = sc * a[i];
        avec[2 * i] = creal(t); avec[2 * i + 1] = cimag(t);
        d2 += conj(t) * (escale * v[i]);
      }
      if (dots) { dots[0] = creal(d1); dots[1] = cimag(d1); dots[2] = creal(d2); dots[3] = cimag(d2); }
    }
  }
  free(A); free(B); free(alp); free(beta); free(vl); free(vr); free(work); free(rwork); free(t2); free(ind); free(eval);
  return info;
}

/*
 * orrncbl: Newton iteration on alpha to Im w = 0 (orrncbl.f:109-116 + 758-981),
 * bug-compatible.  trace: maxit rows of [nalpha, w_r, w_i, dwda_r, dwda_i, dalpha, nw]
 * (the reference prints nalpha, w_r, w_i per step, orrncbl.f:114-115).
 * Returns the number of Newton steps taken.
 */
int osr_chan_neutral(const osr_chan *c, double Re, double sfac, double alpha_r, double tol, int maxit,
                     double *trace) {
  const int n = c->n;
  int m = n - 1;
  zc *A = (zc *)malloc(sizeof(zc) * (size_t)m * m), *B = (zc *)malloc(sizeof(zc) * (size_t)m * m);
  zc *dA = (zc *)malloc(sizeof(zc) * (size_t)m * m), *dB = (zc *)malloc(sizeof(zc) * (size_t)m * m);
  zc *T1 = (zc *)malloc(sizeof(zc) * (size_t)m * m);
  zc *eval = (zc *)malloc(sizeof(zc) * m), *vl = (zc *)malloc(sizeof(zc) * (size_t)m * m),
     *vr = (zc *)malloc(sizeof(zc) * (size_t)m * m), *ctemp = (zc *)malloc(sizeof(zc) * m);
  double *t1 = (double *)malloc(sizeof(double) * m), *t2 = (double *)malloc(sizeof(double) * m);
  int *ind = (int *)malloc(sizeof(int) * m);
  double nw = -9999., nalpha = alpha_r;
  zc eigenvalue = CMPLX(1., 1.);
  int it = 0;
  while (fabs(cimag(eigenvalue)) > tol && it < maxit) {
    const zc alpha = CMPLX(nalpha, 0.0);
    assemble(c, alpha, Re, 0, A, B, dA, dB);
    eig_invBA(m, A, B, eval, vl, vr, 1);
    for (int i = 0; i < m; ++i) { t1[i] = creal(eval[i]); t2[i] = cimag(eval[i]); ind[i] = i; }
    piksr2(m, t2, ind);
    int first = 1, iloc = m - 1;
    double diff = .1;
    if (nw <= -999) {
      for (int i = m - 1; i >= 0; --i)
        if (fabs(t1[ind[i]]) < 1. && first) { eigenvalue = CMPLX(t1[ind[i]], t2[i]); iloc = i; first = 0; }
    } else {
      for (int i = m - 1; i >= 0; --i)
        if (fabs(t1[ind[i]] - nw) <= diff && first) {
          diff = fabs(t1[ind[i]] - nw); eigenvalue = CMPLX(t1[ind[i]], t2[i]); iloc = i; first = 0;
        }
    }
    const zc *tvec = vr + (size_t)ind[iloc] * m, *tavec = vl + (size_t)ind[iloc] * m;
    for (size_t k = 0; k < (size_t)m * m; ++k) T1[k] = dB[k] * eigenvalue - dA[k];
    zc one = 1.0, zero = 0.0;
    int inc = 1;
    L.zgemv("N", &m, &m, &one, T1, &m, (zc *)tvec, &inc, &zero, ctemp, &inc);
    zc dw = 0.0;
    for (int i = 0; i < m; ++i) dw += conj(tavec[i]) * ctemp[i];
    L.zgemv("N", &m, &m, &one, B, &m, (zc *)tvec, &inc, &zero, ctemp, &inc);
    zc dalp = 0.0;
    for (int i = 0; i < m; ++i) dalp += conj(tavec[i]) * ctemp[i];
    zc dwda = -1.0 * (dw / dalp);
    double dalpha = -cimag(eigenvalue) / cimag(dwda);
    nalpha = creal(alpha) + sfac * dalpha;
    nw = creal(eigenvalue) + sfac * dalpha * creal(dwda);
    if (trace) {
      double *r = trace + 7 * it;
      r[0] = nalpha; r[1] = creal(eigenvalue); r[2] = cimag(eigenvalue);
      r[3] = creal(dwda); r[4] = cimag(dwda); r[5] = dalpha; r[6] = nw;
    }
    ++it;
  }
  free(A); free(B); free(dA); free(dB); free(T1); free(eval); free(vl); free(vr); free(ctemp);
  free(t1); free(t2); free(ind);
  return it;
}
