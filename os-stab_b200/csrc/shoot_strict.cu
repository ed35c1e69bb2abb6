// shoot_strict.cu -- the STRICT-parity build of K2 (OSB_STRICT=1).
//
// Same iteration as shoot_kernels.cu (CONTE_IC shoot.f:391-813, CONTE shoot.f:26-388, the private
// copies orrsom.f:165-506 and orrspace.f:209-582), but every floating-point operation is performed in
// the REFERENCE's order and grouping, without FMA contraction (this file is compiled with
// -fmad=false; nvcc's division and sqrt are IEEE by default):
//   * the RHS as typed, per flow -- contebl.f:335-337 / conte.f:205-207, orrsom's scaled form
//     orrsom.f:637-639, orrspace's c = omega/alpha orrspace.f:645-647;
//   * CRKCK45 / CRK4 as coded (stage derivatives scaled by h first, then by the tableau; the two
//     zero weights included), one solution vector after the other;
//   * complex * and / expanded as gfortran does (4-multiply product, Smith's division);
//   * the orthonormality test with ABS, SQRT and ACOS as coded (shoot.f:593-606 and its three
//     variants), Gram-Schmidt with complex SQRT and complex division, Muller's update as coded.
// The few libm calls (EXP, SINCOS, HYPOT, ACOS) are evaluated with osb_pmath.h on both sides, so that
// the result -- eigenvalue, pass count, qend, the whole per-pass history -- is BIT-IDENTICAL to the CPU
// restatement of the reference run with the same four functions (tests/test_strict_parity.py).  This is
// the proof that the fast kernels differ from the reference by rounding only: the fast build regroups
// the RHS, folds h into the tableau, contracts to FMA and replaces acos by a squared comparison.
// About 3x slower than the fast kernel; never selected unless OSB_STRICT=1.
#include "osb_cplx.cuh"
#include "osb_internal.h"
#include "osb_pmath.h"

namespace osb {
namespace {

// Cash-Karp tableau as typed in the reference (shoot.f:1381-1394); b row-major lower triangle
__constant__ double kB[15] = {0.2,
                              0.075, 0.225,
                              0.3, -0.9, 1.2,
                              -0.2037037037037037, 2.5, -2.5925925925925926, 1.2962962962962963,
                              0.029495804398148147, 0.341796875, 0.041594328703703706,
                              0.40034541377314814, 0.061767578125};
__constant__ double kC[6] = {0.09788359788359788, 0.0, 0.4025764895330113,
                             0.21043771043771045, 0.0, 0.2891022021456804};

__device__ __forceinline__ zd rz(double r, zd z) { return zd{r * z.x, r * z.y}; }      // real * complex
__device__ __forceinline__ zd zr_div(zd z, double r) { return zd{z.x / r, z.y / r}; }  // complex / real
__device__ __forceinline__ double s_cabs(zd z) { return pm_hypot(z.x, z.y); }
// glibc s_csqrt_template.c (finite arguments away from the scaling branches)
__device__ __forceinline__ zd s_csqrt(zd z) {
  const double re = z.x, im = z.y;
  if (im == 0.0) {
    if (re < 0.0) return zd{0.0, copysign(__dsqrt_rn(-re), im)};
    return zd{fabs(__dsqrt_rn(re)), copysign(0.0, im)};
  }
  if (re == 0.0) {
    const double r = __dsqrt_rn(0.5 * fabs(im));
    return zd{r, copysign(r, im)};
  }
  const double d = pm_hypot(re, im);
  double r, s;
  if (re > 0.0) {
    r = __dsqrt_rn(0.5 * (d + re));
    s = 0.5 * (im / r);
  } else {
    s = __dsqrt_rn(0.5 * (d - re));
    r = fabs(0.5 * (im / s));
  }
  return zd{r, copysign(s, im)};
}
// glibc s_cexp_template.c (finite arguments below the overflow threshold)
__device__ __forceinline__ zd s_cexp(zd z) {
  double sn, cs;
  if (fabs(z.y) > 2.2250738585072014e-308) pm_sincos(z.y, &sn, &cs);
  else { sn = z.y; cs = 1.0; }
  const double e = pm_exp(z.x);
  return zd{e * cs, e * sn};
}

struct SPass {       // per-pass constants of the RHS; each is the value the reference recomputes per call
  zd alpha, c;
  zd a2, a4, ta2;    // alpha^2, alpha^4, 2 alpha^2
  zd k;              // (0,1) alpha Re
  zd q;              // orrsom: 1 / alpha / Re
  double Re;
  int flow;
};

__device__ __forceinline__ SPass s_pass(int flow, zd alpha, double Re, zd c) {
  SPass p;
  p.flow = flow; p.alpha = alpha; p.c = c; p.Re = Re;
  p.a2 = zmul(alpha, alpha);
  p.a4 = zmul(p.a2, p.a2);
  p.ta2 = rz(2.0, p.a2);
  p.k = rz(Re, zd{-alpha.y, alpha.x});
  p.q = zr_div(zdiv(zd{1.0, 0.0}, alpha), Re);
  return p;
}

// row 4 of OSHOMO / FHOMO as typed
__device__ __forceinline__ zd s_rhs4(const SPass &p, const zd *yo, double UU, double d2UU) {
  const zd P = zsub(zmul(p.ta2, yo[2]), zmul(p.a4, yo[0]));  // 2 a^2 y3 - a^4 y1
  const zd T1 = zd{UU - p.c.x, -p.c.y};                       // (U - c)
  const zd T2 = zsub(yo[2], zmul(p.a2, yo[0]));               // (y3 - a^2 y1)
  const zd T5 = zsub(zmul(T1, T2), rz(d2UU, yo[0]));          // (U-c)(...) - U'' y1
  if (p.flow == OSB_FLOW_ORRSOM) {                            // orrsom.f:637-639
    const zd S = zadd(zmul(p.q, P), zd{-T5.y, T5.x});
    return rz(p.Re, zmul(S, p.alpha));
  }
  return zadd(P, zmul(p.k, T5));                              // contebl.f:335-337, orrspace.f:645-647
}

__device__ __forceinline__ void s_rhs(const SPass &p, const zd *yo, double2 uu, zd *f) {
  f[0] = yo[1]; f[1] = yo[2]; f[2] = yo[3];
  f[3] = s_rhs4(p, yo, uu.x, uu.y);
}

// CRKCK45, shoot.f:1367-1439 (fixed step)
__device__ __forceinline__ void s_ck45(const SPass &p, zd *y, const double2 *__restrict__ trow, double h) {
  zd yk[6][4];
#pragma unroll
  for (int m = 0; m < 6; ++m) {
    zd yt[4], f[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) yt[n] = y[n];
#pragma unroll
    for (int k = 0; k < m; ++k)
#pragma unroll
      for (int n = 0; n < 4; ++n) yt[n] = zadd(yt[n], rz(kB[m * (m - 1) / 2 + k], yk[k][n]));
    s_rhs(p, yt, __ldg(&trow[m]), f);
#pragma unroll
    for (int n = 0; n < 4; ++n) yk[m][n] = rz(h, f[n]);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int n = 0; n < 4; ++n) y[n] = zadd(y[n], rz(kC[k], yk[k][n]));
}

// CRK4, shoot.f:869-903
__device__ __forceinline__ void s_rk4(const SPass &p, zd *y, const double2 *__restrict__ trow, double h) {
  zd f[4], k1[4], k2[4], k3[4], k4[4], q[4];
  s_rhs(p, y, __ldg(&trow[0]), f);
#pragma unroll
  for (int j = 0; j < 4; ++j) { k1[j] = rz(h, f[j]); q[j] = zadd(y[j], rz(0.5, k1[j])); }
  s_rhs(p, q, __ldg(&trow[1]), f);
#pragma unroll
  for (int j = 0; j < 4; ++j) { k2[j] = rz(h, f[j]); q[j] = zadd(y[j], rz(0.5, k2[j])); }
  s_rhs(p, q, __ldg(&trow[1]), f);
#pragma unroll
  for (int j = 0; j < 4; ++j) { k3[j] = rz(h, f[j]); q[j] = zadd(y[j], k3[j]); }
  s_rhs(p, q, __ldg(&trow[2]), f);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    k4[j] = rz(h, f[j]);
    y[j] = zadd(zadd(zadd(y[j], zr_div(k1[j], 6.0)), zr_div(zadd(k2[j], k3[j]), 3.0)), zr_div(k4[j], 6.0));
  }
}

// INPROD: sum v1 conj(v2) (contebl.f:238-241) or sum v1 v2 (orrspace.f:597-601), accumulated from 0
template <bool ANALYTIC>
__device__ __forceinline__ zd s_inprod(const zd *v1, const zd *v2) {
  zd r = zd{0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 4; ++i) r = zadd(r, zmul(v1[i], ANALYTIC ? v2[i] : zconj(v2[i])));
  return r;
}

// Gram-Schmidt, shoot.f:490-515 / 621-639
template <bool ANALYTIC>
__device__ __forceinline__ void s_gs(zd (&u)[2][4]) {
  const zd w0 = s_csqrt(s_inprod<ANALYTIC>(u[0], u[0]));
  zd z0[4], eta[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) z0[i] = zdiv(u[0][i], w0);
  const zd ip = s_inprod<ANALYTIC>(u[1], z0);
#pragma unroll
  for (int i = 0; i < 4; ++i) eta[i] = zsub(u[1][i], zmul(ip, z0[i]));
  const zd w1 = s_csqrt(s_inprod<ANALYTIC>(eta, eta));
#pragma unroll
  for (int i = 0; i < 4; ++i) { u[0][i] = z0[i]; u[1][i] = zdiv(eta[i], w1); }
}

// the orthonormality test as coded in each driver
template <bool ANALYTIC>
__device__ __forceinline__ bool s_need_norm(const ShootArgs &a, const zd (&u)[2][4]) {
  const double pi = 3.14159265358979323846;  // acos(-1.0) (shoot.f:467) and the literal of orrspace.f:267: same double
  const double aa = s_cabs(s_inprod<ANALYTIC>(u[0], u[0]));
  const double bb = s_cabs(s_inprod<ANALYTIC>(u[1], u[1]));
  const double cc = s_cabs(s_inprod<ANALYTIC>(u[0], u[1]));  // |<U2,U1>| is the same number: one evaluation
  double x1 = cc / __dsqrt_rn(aa * bb);
  switch (a.flow) {
    case OSB_FLOW_CONTEBL:  // shoot.f:593-606
      if (x1 > 1.0) x1 = 1.0;
      return pm_acos(x1) < pi * a.testalpha / 180.0;
    case OSB_FLOW_CONTE:  // shoot.f:64-65, 132-151
      if (x1 > 1.0) x1 = 1.0;
      return pm_acos(x1) <= 10.0 * pi / 180.0;
    case OSB_FLOW_ORRSOM:  // orrsom.f:292-302
      if (x1 > 1.0) x1 = 1.0;
      return 180.0 * pm_acos(x1) / pi <= a.testalpha;
    default:  // orrspace.f:350-358
      return x1 > a.testalpha;
  }
}

template <int INTEG, bool ANALYTIC>
__device__ void s_integrate(const ShootArgs &a, zd alpha, double Re, zd cw, zd &err, int &qend) {
  constexpr int NS = (INTEG == OSB_INT_CK45) ? 6 : 3;
  const SPass p = s_pass(a.flow, alpha, Re, cw);
  zd u[2][4];
  if (a.bl_ic) {  // BL_IC contebl.f:269-279; orrsom.f:211-221; orrspace.f:285-295
    const zd ea = s_cexp(rz(a.tf, zneg(alpha)));
    u[0][0] = ea;
    u[0][1] = zmul(zneg(alpha), ea);
    u[0][2] = zmul(p.a2, ea);
    u[0][3] = zmul(zneg(zmul(p.a2, alpha)), ea);
    const zd g = s_csqrt(zadd(p.a2, zmul(p.k, zd{1.0 - cw.x, -cw.y})));
    const zd eg = s_cexp(rz(a.tf, zneg(g)));
    const zd g2 = zmul(g, g);
    u[1][0] = eg;
    u[1][1] = zmul(zneg(g), eg);
    u[1][2] = zmul(g2, eg);
    u[1][3] = zmul(zneg(zmul(g2, g)), eg);
    s_gs<ANALYTIC>(u);
  } else {  // e3, e4 (shoot.f:88-96); no Gram-Schmidt at k = 0
#pragma unroll
    for (int i = 0; i < 4; ++i) { u[0][i] = zd{0.0, 0.0}; u[1][i] = zd{0.0, 0.0}; }
    u[0][2] = zd{1.0, 0.0};
    u[1][3] = zd{1.0, 0.0};
  }
  int q = 0;
  const double2 *__restrict__ trow = a.tab;
  for (int k = 1; k <= a.nstep; ++k) {
#pragma unroll 1
    for (int m = 0; m < 2; ++m) {
      if (INTEG == OSB_INT_CK45) s_ck45(p, u[m], trow, a.hs);
      else s_rk4(p, u[m], trow, a.hs);
    }
    trow += NS;
    if (s_need_norm<ANALYTIC>(a, u) || k == a.nstep) {
      q += 1;
      s_gs<ANALYTIC>(u);
    }
  }
  qend = q;
  // B(1) = -U(1,2)/U(1,1), B(2) = 1; err = U(2,1) B(1) + U(2,2) B(2)   (shoot.f:699-700, 716)
  const zd B1 = zdiv(zneg(u[1][0]), u[0][0]);
  err = zadd(zmul(u[0][1], B1), zmul(u[1][1], zd{1.0, 0.0}));
}

struct SIter {
  zd ev, cm1, cm2, errm1, errm2;
  int icount, cm1_defined;
};

// shoot.f:713-744 as coded
__device__ __forceinline__ void s_update(SIter &s, zd err, zd ctemp, int first_perturb) {
  if (s.icount == 1) {
    s.cm2 = s.ev;
    s.errm2 = err;
    if (first_perturb) s.ev = zd{s.ev.x * 1.01, s.ev.y * 1.01};
    else s.ev = zd{s.ev.x * .9999, s.ev.y};
  } else if (s.icount == 2) {
    s.cm1 = s.ev; s.cm1_defined = 1;
    s.errm1 = err;
    const double den = s.ev.x - s.cm2.x;
    const zd fd = zd{(err.x - s.errm2.x) / den, (err.y - s.errm2.y) / den};
    s.ev = zsub(s.ev, zdiv(err, fd));
  } else {
    const zd c = s.ev;
    const zd qt = zdiv(zsub(c, s.cm1), zsub(s.cm1, s.cm2));
    const zd one_qt = zd{1.0 + qt.x, qt.y};
    const zd qt2 = zmul(qt, qt);
    const zd At = zadd(zsub(zmul(qt, err), zmul(zmul(qt, one_qt), s.errm1)), zmul(qt2, s.errm2));
    const zd Bt = zadd(zsub(zmul(zd{2.0 * qt.x + 1.0, 2.0 * qt.y}, err), zmul(zmul(one_qt, one_qt), s.errm1)),
                       zmul(qt2, s.errm2));
    const zd Ct = zmul(one_qt, err);
    const zd D = s_csqrt(zsub(zmul(Bt, Bt), zmul(rz(4.0, At), Ct)));
    const zd dp = zadd(Bt, D), dm = zsub(Bt, D);
    const zd den = (s_cabs(dp) > s_cabs(dm)) ? dp : dm;
    s.ev = zsub(ctemp, zdiv(zmul(rz(2.0, zsub(ctemp, s.cm1)), Ct), den));
    s.cm2 = s.cm1;
    s.cm1 = ctemp;
    s.errm2 = s.errm1;
    s.errm1 = err;
  }
}

__device__ __forceinline__ bool s_keep_going(const ShootArgs &a, const SIter &s, zd err) {
  const double dc = s.cm1_defined ? s_cabs(zsub(s.ev, s.cm1)) : 1.0;
  const double ae = s_cabs(err);
  if (a.loop_rule == 0) return ((ae >= a.errtol) || (dc >= a.eigtol)) && (s.icount <= a.maxcount);
  return (ae >= a.errtol) && (s.icount <= a.maxcount) && (dc >= a.eigtol);
}

template <int INTEG, bool ANALYTIC>
__global__ void __launch_bounds__(64) shoot_strict_kernel(const __grid_constant__ ShootArgs a) {
  const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= a.nlines) return;
  const double Re = a.Re[line];
  zd alpha, omega = zd{0.0, 0.0};
  SIter s;
  {
    const double2 al = a.alpha[line];
    alpha = zd{al.x, al.y};
    if (a.spatial) {
      const double2 om = a.omega0[line];
      omega = zd{om.x, om.y};
      s.ev = alpha;
    } else {
      const double2 c0 = a.ev[line];
      s.ev = zd{c0.x, c0.y};
    }
  }
  for (int it = 0; it < a.niter; ++it) {
    const int64_t out = line * a.niter + it;
    s.icount = 1;
    s.cm1_defined = a.cm1_defined;
    s.cm1 = zd{s.ev.x + 1.0, s.ev.y};
    s.cm2 = zd{0.0, 0.0}; s.errm1 = zd{0.0, 0.0}; s.errm2 = zd{0.0, 0.0};
    zd err = zd{1.0, 0.0}, ctemp = s.ev;
    int qend = 0, status = OSB_ST_MAXCOUNT;
    while (true) {
      if (!s_keep_going(a, s, err)) {
        if (s.icount <= a.maxcount) status = OSB_ST_CONVERGED;
        break;
      }
      ctemp = s.ev;
      {
        const zd al = a.spatial ? s.ev : alpha;
        const zd cw = a.spatial ? zdiv(omega, s.ev) : s.ev;  // orrspace.f:290,645
        s_integrate<INTEG, ANALYTIC>(a, al, Re, cw, err, qend);
      }
      s_update(s, err, ctemp, a.first_perturb);
      if (a.history && s.icount <= a.hist_cap) {
        const zd shown = a.show_updated ? s.ev : ctemp;
        a.history[out * a.hist_cap + (s.icount - 1)] = make_double4(shown.x, shown.y, err.x, err.y);
      }
      s.icount += 1;
      if (!(zfinite(s.ev) && zfinite(err))) { status = OSB_ST_NONFINITE; break; }
    }
    a.ev[out] = make_double2(ctemp.x, ctemp.y);
    a.passes[out] = s.icount - 1;
    a.qend[out] = qend;
    a.status[out] = status;
    if (a.resid) a.resid[out] = make_double2(s_cabs(err), s.cm1_defined ? s_cabs(zsub(s.ev, s.cm1)) : 0.0);
    if (a.spatial) {
      s.ev = ctemp;                // orrspace.f:494
      omega.x += a.domega[line];   // orrspace.f:192-194
    }
  }
}

}  // namespace

cudaError_t launch_shoot_strict(const ShootArgs &a, int integrator, int analytic_inprod, cudaStream_t st) {
  if (a.nlines <= 0) return cudaSuccess;
  const int threads = 32;
  const int64_t blocks = (a.nlines + threads - 1) / threads;
  if (blocks > 0x7fffffff) return cudaErrorInvalidValue;
  const dim3 g((unsigned)blocks), b(threads);
  if (integrator == OSB_INT_CK45) {
    if (analytic_inprod) shoot_strict_kernel<OSB_INT_CK45, true><<<g, b, 0, st>>>(a);
    else shoot_strict_kernel<OSB_INT_CK45, false><<<g, b, 0, st>>>(a);
  } else {
    if (analytic_inprod) shoot_strict_kernel<OSB_INT_RK4, true><<<g, b, 0, st>>>(a);
    else shoot_strict_kernel<OSB_INT_RK4, false><<<g, b, 0, st>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace osb
