/*
 * os_chan_arbiter.c -- quad-precision arbiter for the channel matrix path.
 *
 * TEST INFRASTRUCTURE ONLY.  For Chebyshev collocation the pencil is so ill-scaled
 * (||A|| ~ 5e18 at N = 512) that the reference's own LAPACK routes (ZGEEV on inv(B)A, ZGEGV)
 * disagree with each other far above 1e-10.  To judge the GPU path where the
 * fp64 oracle is itself inaccurate, this file solves the SAME discrete pencil -- the fp64
 * tables D2, D4, u, u'' promoted exactly to __float128 -- by shift-invert inverse iteration
 * in 113-bit arithmetic.  Its answer is the exact eigenvalue of the discrete operator to
 * ~1e-25, so  |w_gpu - w_quad|  and  |w_oracle - w_quad|  are the true errors of each.
 *
 * The assembly formulas are those of MAKE_MATRIX (orrfdchan.f:553-659) with the channel
 * metrics m1 = 1, m2 = m3 = m4 = 0; interior (N-1) x (N-1) pencil.
 */
#include <quadmath.h>
#include <stdlib.h>
#include <string.h>

typedef __float128 q;
typedef __complex128 qc;

/* D2, D4: (N+1)^2 row-major (i,j) fp64; u, d2u: N+1.  sigma: the shift (an fp64 estimate).
 * out[4]: Re w, Im w rounded to fp64, and the residual correction terms (w - fp64(w)). */
int osr_chan_arbiter(int N, const double *D2, const double *D4, const double *u, const double *d2u,
                     double alpha_re, double alpha_im, double Re, double sigma_re, double sigma_im,
                     int iters, double *out) {
  const int n = N - 1, n1 = N + 1;
  qc alpha = (q)alpha_re + (q)alpha_im * 1.0Qi;
  qc a2 = alpha * alpha, a3 = a2 * alpha, a4 = a2 * a2;
  qc iu = 1.0Qi;
  q R = (q)Re;
  qc b2 = -1.0Q * iu * R, b0 = iu * a2 * R;
  qc *M = (qc *)malloc(sizeof(qc) * (size_t)n * n);
  qc *x = (qc *)malloc(sizeof(qc) * n), *y = (qc *)malloc(sizeof(qc) * n), *bx = (qc *)malloc(sizeof(qc) * n);
  int *piv = (int *)malloc(sizeof(int) * n);
  qc sigma = (q)sigma_re + (q)sigma_im * 1.0Qi;
  qc w = sigma;
  for (int i = 0; i < n; ++i) x[i] = 1.0Q + 0.3Q * (q)(i % 7) + 0.1Qi * (q)(i % 3);
  for (int round = 0; round < 2; ++round) {
    for (int i = 0; i < n; ++i) {
      const int gi = i + 1;
      qc w2 = -2.0Q * a2 - iu * alpha * R * (q)u[gi];
      qc w0 = a4 + iu * a3 * R * (q)u[gi] + iu * alpha * R * (q)d2u[gi];
      for (int j = 0; j < n; ++j) {
        const int gj = j + 1;
        q d2 = (q)D2[(size_t)gi * n1 + gj], d4 = (q)D4[(size_t)gi * n1 + gj];
        qc A = d4 + w2 * d2 + (i == j ? w0 : 0.0Q);
        qc B = b2 * d2 + (i == j ? b0 : 0.0Q);
        M[(size_t)i * n + j] = A - sigma * B; /* row-major */
      }
    }
    /* LU with partial pivoting (row-major, in place) */
    for (int k = 0; k < n; ++k) {
      int p = k;
      q best = cabsq(M[(size_t)k * n + k]);
      for (int r = k + 1; r < n; ++r) { q v = cabsq(M[(size_t)r * n + k]); if (v > best) { best = v; p = r; } }
      piv[k] = p;
      if (p != k) for (int j = 0; j < n; ++j) { qc t = M[(size_t)k * n + j]; M[(size_t)k * n + j] = M[(size_t)p * n + j]; M[(size_t)p * n + j] = t; }
      qc rinv = 1.0Q / M[(size_t)k * n + k];
      for (int r = k + 1; r < n; ++r) {
        qc l = M[(size_t)r * n + k] * rinv;
        M[(size_t)r * n + k] = l;
        if (l != 0.0Q) for (int j = k + 1; j < n; ++j) M[(size_t)r * n + j] -= l * M[(size_t)k * n + j];
      }
    }
    for (int it = 0; it < iters; ++it) {
      /* bx = B x */
      for (int i = 0; i < n; ++i) {
        qc s = 0.0Q;
        for (int j = 0; j < n; ++j) s += (q)D2[(size_t)(i + 1) * n1 + (j + 1)] * x[j];
        bx[i] = b2 * s + b0 * x[i];
      }
      for (int k = 0; k < n; ++k) { int p = piv[k]; if (p != k) { qc t = bx[k]; bx[k] = bx[p]; bx[p] = t; } }
      for (int i = 0; i < n; ++i) { qc s = bx[i]; for (int j = 0; j < i; ++j) s -= M[(size_t)i * n + j] * y[j]; y[i] = s; }
      for (int i = n - 1; i >= 0; --i) { qc s = y[i]; for (int j = i + 1; j < n; ++j) s -= M[(size_t)i * n + j] * y[j]; y[i] = s / M[(size_t)i * n + i]; }
      qc yx = 0.0Q;
      q yy = 0.0Q;
      for (int i = 0; i < n; ++i) { yx += conjq(y[i]) * x[i]; yy += crealq(y[i] * conjq(y[i])); }
      w = sigma + yx / yy;
      q rn = 1.0Q / sqrtq(yy);
      for (int i = 0; i < n; ++i) x[i] = y[i] * rn;
    }
    sigma = w; /* re-shift at the converged estimate and polish */
  }
  out[0] = (double)crealq(w); out[1] = (double)cimagq(w);
  out[2] = (double)(crealq(w) - (q)out[0]); out[3] = (double)(cimagq(w) - (q)out[1]);
  free(M); free(x); free(y); free(bx); free(piv);
  return 0;
}
