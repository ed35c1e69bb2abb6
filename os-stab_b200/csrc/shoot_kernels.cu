// shoot_kernels.cu -- K1b (stage table) and K2 (Godunov-Conte shooting) for sm_100a.
//
// K2 replaces the reference's CONTE_IC / CONTE eigenvalue iteration
// (shoot.f:391-813, shoot.f:26-388, orrsom.f:165-506, orrspace.f:209-582) for a
// whole batch of independent sweep points: one thread owns one point (or one
// continuation line), keeps both solution vectors (2 x 4 complex) and all
// Runge-Kutta stage derivatives in registers, and runs the secant/Muller update
// on the eigenvalue inside the kernel.  The work is bound by the FP64 pipe;
// HBM traffic is ~60 B per point.
//
// Design notes:
//  * The mean flow enters the RHS only through U(t), U''(t) at the integrator's
//    stage abscissae, and those abscissae are the same for every point of the
//    batch.  K1b evaluates the reference's spline (SPDER/SPEVAL, shoot.f:960-1101)
//    once per abscissa -- in the reference's own operation order, without FMA
//    contraction -- into a [nstep][stages] table.  K2 then reads one warp-uniform
//    row (96 B for Cash-Karp) per step instead of bisecting + evaluating a cubic
//    12 times per step per point as the reference does.
//  * RHS row 4 (contebl.f:335-337) is evaluated as G d + a^2 y3 - (ik U'') y1 with
//    d = y3 - a^2 y1 and G = a^2 + ik(U - c), k = alpha*Re; G and ik U'' are computed
//    once per stage and shared by both solution vectors (see rhs4 for why d must be
//    kept as its own rounded quantity).
//  * The step size is folded into the tableau on the host (bh = b*h, ch = c*h).
//  * The orthonormality test acos(cc/sqrt(aa*bb)) < ta (shoot.f:593-606) is
//    evaluated as cc^2 > cos^2(ta)*aa*bb (acos is monotone): no sqrt/div/acos on
//    the per-step path.
#include <stdlib.h>

#include <algorithm>

#include "osb_cplx.cuh"
#include "osb_internal.h"

namespace osb {

// ---------------------------------------------------------------------------
// K1b: stage table
// ---------------------------------------------------------------------------
#define OSB_MUL(a, b) __dmul_rn((a), (b))
#define OSB_ADD(a, b) __dadd_rn((a), (b))
#define OSB_SUB(a, b) __dsub_rn((a), (b))
#define OSB_DIV(a, b) __ddiv_rn((a), (b))

// BISECT (shoot.f:1018-1034) with the exact-knot shortcuts of SPDER
// (shoot.f:1070-1076); returns the 1-based interval index.
__device__ static int find_bisect(const double *X0, int N, double xx) {
  if (xx == X0[0]) return 1;
  if (xx == X0[N - 1]) return N;
  int il = 0, ir = N - 1;
  while (ir - il > 1) {
    int im = (ir + il) >> 1;
    if (X0[im] > xx) ir = im; else il = im;
  }
  return il + 1;
}

// linear search of orrsom.f:885-908 / orrspace.f:906-929: first I with xx <= X(I+1).
// X is ascending, so a binary search for the same predicate gives the same I.
__device__ static int find_linear(const double *X0, int N, double xx) {
  int lo = 1, hi = N;  // answer in [1, N]; N means "fell through"
  while (lo < hi) {
    int mid = (lo + hi) >> 1;  // candidate I = mid (<= N-1 here)
    if (xx <= X0[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// the cubic of SPEVAL/SPDER (shoot.f:1008-1013, 1090-1095), reference operation order
__device__ static double spline_value(const double *X0, const double *Y0, const double *F0,
                                      int I, double XX) {
  const double *X = X0 - 1, *Y = Y0 - 1, *FDP = F0 - 1;
  double DXM = OSB_SUB(XX, X[I]);
  double DXP = OSB_SUB(X[I + 1], XX);
  double DEL = OSB_SUB(X[I + 1], X[I]);
  double t1 = OSB_DIV(OSB_MUL(OSB_MUL(FDP[I], DXP), OSB_SUB(OSB_DIV(OSB_MUL(DXP, DXP), DEL), DEL)), 6.0);
  double t2 = OSB_DIV(OSB_MUL(OSB_MUL(FDP[I + 1], DXM), OSB_SUB(OSB_DIV(OSB_MUL(DXM, DXM), DEL), DEL)), 6.0);
  double t3 = OSB_DIV(OSB_MUL(Y[I], DXP), DEL);
  double t4 = OSB_DIV(OSB_MUL(Y[I + 1], DXM), DEL);
  return OSB_ADD(OSB_ADD(OSB_ADD(t1, t2), t3), t4);
}
// FPP of SPDER (shoot.f:1099)
__device__ static double spline_fpp(const double *X0, const double *F0, int I, double XX) {
  const double *X = X0 - 1, *FDP = F0 - 1;
  double DXM = OSB_SUB(XX, X[I]);
  double DXP = OSB_SUB(X[I + 1], XX);
  double DEL = OSB_SUB(X[I + 1], X[I]);
  return OSB_ADD(OSB_DIV(OSB_MUL(FDP[I], DXP), DEL), OSB_DIV(OSB_MUL(FDP[I + 1], DXM), DEL));
}

int stages_per_step(int integrator) { return integrator == OSB_INT_CK45 ? 6 : 3; }

__global__ void stage_table_kernel(const double *__restrict__ tables, int nbl, int flow,
                                   int integrator, int nstep, double to, double tf,
                                   double2 *__restrict__ tab) {
  const int NS = (integrator == OSB_INT_CK45) ? 6 : 3;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nstep * NS) return;
  int k = idx / NS + 1;
  int s = idx % NS;
  const double h = OSB_DIV(OSB_SUB(tf, to), (double)nstep);  // shoot.f:471
  double t0, hs;
  if (flow == OSB_FLOW_CONTE) {  // t = to + h*k; STEP(.., t-h, +h)   shoot.f:110,116
    double t = OSB_ADD(to, OSB_MUL(h, (double)k));
    t0 = OSB_SUB(t, h);
    hs = h;
  } else {  // t = tf - h*k; STEP(.., t+h, -h)   shoot.f:527,563
    double t = OSB_SUB(tf, OSB_MUL(h, (double)k));
    t0 = OSB_ADD(t, h);
    hs = -h;
  }
  double ts;
  if (integrator == OSB_INT_CK45) {  // t = to + a(m)*h   shoot.f:1410
    const double a[6] = {0.0, 0.2, 0.3, 0.6, 1.0, 0.875};
    ts = OSB_ADD(t0, OSB_MUL(a[s], hs));
  } else {  // to, to+0.5*h, to+h   shoot.f:881-896
    ts = (s == 0) ? t0 : (s == 1 ? OSB_ADD(t0, OSB_MUL(0.5, hs)) : OSB_ADD(t0, hs));
  }
  double UU, d2UU;
  if (flow == OSB_FLOW_CONTE) {  // CHANNEL conte.f:319-321
    UU = OSB_SUB(1.0, OSB_MUL(ts, ts));
    d2UU = -2.0;
  } else {
    const size_t n = (size_t)nbl + 2;
    const double *ydat = tables, *U = tables + n, *Uspl = tables + 2 * n,
                 *d2U = tables + 3 * n, *d2Uspl = tables + 4 * n;
    if (flow == OSB_FLOW_CONTEBL) {  // SPDER contebl.f:330
      int I = find_bisect(ydat, nbl + 1, ts);
      UU = spline_value(ydat, U, Uspl, I, ts);
      d2UU = spline_fpp(ydat, Uspl, I, ts);
    } else {  // two SPEVALs orrsom.f:627-628, orrspace.f:634-635
      int I = find_linear(ydat, nbl + 1, ts);
      UU = spline_value(ydat, U, Uspl, I, ts);
      d2UU = spline_value(ydat, d2U, d2Uspl, I, ts);
    }
  }
  tab[idx] = make_double2(UU, d2UU);
}

cudaError_t launch_stage_table(const ProfileDev &p, int flow, int integrator, int nstep,
                               double to, double tf, double2 *tab, cudaStream_t st) {
  int n = nstep * stages_per_step(integrator);
  int threads = 128, blocks = (n + threads - 1) / threads;
  stage_table_kernel<<<blocks, threads, 0, st>>>(p.tables, p.nbl, flow, integrator, nstep, to, tf, tab);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K2: shooting
// ---------------------------------------------------------------------------
struct PassConst {
  double a2r, a2i;  // alpha^2
  double kr, ki;    // i*alpha*Re
  double g0r, g0i;  // alpha^2 - i*alpha*Re*c
};

struct StageCoef {
  double Gr, Gi;  // G = a^2 + ik(U - c)
  double Kr, Ki;  // ik U''
};

// REAL_ALPHA: alpha is real, so a^2 is real and ik = i*alpha*Re is purely imaginary
// (Gr is then constant over the pass and Kr = 0).  The general path with those zeros
// produces the same bits, so mixing real and complex alpha in one batch is safe.
template <bool REAL_ALPHA>
__device__ __forceinline__ StageCoef stage_coef(const PassConst &pc, double2 uu) {
  StageCoef s;
  s.Gi = fma(pc.ki, uu.x, pc.g0i);
  s.Ki = pc.ki * uu.y;
  if (REAL_ALPHA) {
    s.Gr = pc.g0r;
    s.Kr = 0.0;
  } else {
    s.Gr = fma(pc.kr, uu.x, pc.g0r);
    s.Kr = pc.kr * uu.y;
  }
  return s;
}

// Row 4 of OSHOMO (contebl.f:335-337), regrouped but keeping the reference's
// d = y3 - a^2 y1 as a separately rounded quantity:
//     r = G d + a^2 y3 - (ik U'') y1,      G = a^2 + ik (U - c)
// (== 2a^2 y3 - a^4 y1 + ik[(U-c) d - U'' y1]).  This is not cosmetic.  In the
// free stream d cancels to a rounding residue, and that residue is what seeds the
// viscous content of the first solution vector -- which, after the Gram-Schmidt
// steps, fixes the PHASE of the wall residual err(c).  With d formed first, the
// eigenvalue guess multiplies a residue that does not itself depend on c, so
// err(c) stays smooth and the secant/Muller update converges as in the reference.
// Grouping the row as A(c) y1 + B(c) y3 makes the residue c-dependent noise and the
// iteration stalls at |err| ~ 1e-5 (measured; see DESIGN.md "err(c) is rounding-phased").
template <bool REAL_ALPHA>
__device__ __forceinline__ void rhs4(const StageCoef &s, const PassConst &pc, const double *y,
                                     double &rr, double &ri) {
  if (REAL_ALPHA) {
    const double dr = fma(-pc.a2r, y[0], y[4]);
    const double di = fma(-pc.a2r, y[1], y[5]);
    rr = fma(s.Gr, dr, fma(-s.Gi, di, fma(pc.a2r, y[4], s.Ki * y[1])));
    ri = fma(s.Gr, di, fma(s.Gi, dr, fma(pc.a2r, y[5], -s.Ki * y[0])));
  } else {
    const double dr = fma(-pc.a2r, y[0], fma(pc.a2i, y[1], y[4]));
    const double di = fma(-pc.a2r, y[1], fma(-pc.a2i, y[0], y[5]));
    rr = fma(s.Gr, dr, fma(-s.Gi, di, fma(pc.a2r, y[4], fma(-pc.a2i, y[5], fma(-s.Kr, y[0], s.Ki * y[1])))));
    ri = fma(s.Gr, di, fma(s.Gi, dr, fma(pc.a2r, y[5], fma(pc.a2i, y[4], fma(-s.Kr, y[1], -s.Ki * y[0])))));
  }
}

// One fixed Cash-Karp step (CRKCK45, shoot.f:1367-1439) of BOTH solution vectors.
// u[v][0..7] = (y1, y2, y3, y4) of vector v.  f[k] holds the stage derivative
// (y2, y3, y4, r) WITHOUT the step size; bh/ch carry it.
template <bool REAL_ALPHA>
__device__ __forceinline__ void step_ck45(double (&u)[2][8], const double2 *__restrict__ trow,
                                          const PassConst &pc, const ShootArgs &a) {
  double f[2][6][8];
#pragma unroll
  for (int m = 0; m < 6; ++m) {
    const StageCoef sc = stage_coef<REAL_ALPHA>(pc, __ldg(&trow[m]));
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      double yt[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        double acc = u[v][n];
#pragma unroll
        for (int k = 0; k < m; ++k) acc = fma(a.bh[m * (m - 1) / 2 + k], f[v][k][n], acc);
        yt[n] = acc;
      }
#pragma unroll
      for (int n = 0; n < 6; ++n) f[v][m][n] = yt[n + 2];
      rhs4<REAL_ALPHA>(sc, pc, yt, f[v][m][6], f[v][m][7]);
    }
  }
#pragma unroll
  for (int v = 0; v < 2; ++v)
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      // c(2) = c(5) = 0 (shoot.f:1390-1391)
      double acc = u[v][n];
      acc = fma(a.ch[0], f[v][0][n], acc);
      acc = fma(a.ch[2], f[v][2][n], acc);
      acc = fma(a.ch[3], f[v][3][n], acc);
      acc = fma(a.ch[5], f[v][5][n], acc);
      u[v][n] = acc;
    }
}

// Same step, one vector at a time with the stage coefficients recomputed for the second
// vector (+2 FP64 instructions per stage): ~100 fewer live registers, which lets three CTAs
// share an SM.  Bit-identical results (same operations per vector, same order).
template <bool REAL_ALPHA>
__device__ __forceinline__ void step_ck45_seq(double (&u)[2][8], const double2 *__restrict__ trow,
                                              const PassConst &pc, const ShootArgs &a) {
#pragma unroll 1
  for (int v = 0; v < 2; ++v) {
    double f[6][8];
    double y[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) y[n] = u[v][n];
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      const StageCoef sc = stage_coef<REAL_ALPHA>(pc, __ldg(&trow[m]));
      double yt[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        double acc = y[n];
#pragma unroll
        for (int k = 0; k < m; ++k) acc = fma(a.bh[m * (m - 1) / 2 + k], f[k][n], acc);
        yt[n] = acc;
      }
#pragma unroll
      for (int n = 0; n < 6; ++n) f[m][n] = yt[n + 2];
      rhs4<REAL_ALPHA>(sc, pc, yt, f[m][6], f[m][7]);
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      double acc = y[n];
      acc = fma(a.ch[0], f[0][n], acc);
      acc = fma(a.ch[2], f[2][n], acc);
      acc = fma(a.ch[3], f[3][n], acc);
      acc = fma(a.ch[5], f[5][n], acc);
      u[v][n] = acc;
    }
  }
}

// One classical RK4 step (CRK4, shoot.f:869-903) of both vectors.  Table row:
// (to, to+h/2, to+h).
template <bool REAL_ALPHA>
__device__ __forceinline__ void step_rk4(double (&u)[2][8], const double2 *__restrict__ trow,
                                         const PassConst &pc, const ShootArgs &a) {
  const double hs = a.hs, h2 = 0.5 * a.hs, h6 = a.hs / 6.0, h3 = a.hs / 3.0;
  const StageCoef s0 = stage_coef<REAL_ALPHA>(pc, __ldg(&trow[0]));
  const StageCoef s1 = stage_coef<REAL_ALPHA>(pc, __ldg(&trow[1]));
  const StageCoef s2 = stage_coef<REAL_ALPHA>(pc, __ldg(&trow[2]));
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    double f[8], q[8], acc[8];
    // k1
#pragma unroll
    for (int n = 0; n < 6; ++n) f[n] = u[v][n + 2];
    rhs4<REAL_ALPHA>(s0, pc, u[v], f[6], f[7]);
#pragma unroll
    for (int n = 0; n < 8; ++n) { acc[n] = fma(h6, f[n], u[v][n]); q[n] = fma(h2, f[n], u[v][n]); }
    // k2
#pragma unroll
    for (int n = 0; n < 6; ++n) f[n] = q[n + 2];
    rhs4<REAL_ALPHA>(s1, pc, q, f[6], f[7]);
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = fma(h3, f[n], acc[n]);
    {
      double q2[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) q2[n] = fma(h2, f[n], u[v][n]);
      // k3
#pragma unroll
      for (int n = 0; n < 6; ++n) f[n] = q2[n + 2];
      rhs4<REAL_ALPHA>(s1, pc, q2, f[6], f[7]);
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) { acc[n] = fma(h3, f[n], acc[n]); q[n] = fma(hs, f[n], u[v][n]); }
    // k4
#pragma unroll
    for (int n = 0; n < 6; ++n) f[n] = q[n + 2];
    rhs4<REAL_ALPHA>(s2, pc, q, f[6], f[7]);
#pragma unroll
    for (int n = 0; n < 8; ++n) u[v][n] = fma(h6, f[n], acc[n]);
  }
}

// <v1, v2>: Hermitian sum v1 conj(v2) (contebl.f:238-241) or analytic sum v1 v2
// (orrspace.f:597-601)
template <bool ANALYTIC>
__device__ __forceinline__ zd inprod(const double *v1, const double *v2) {
  double re = 0.0, im = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double ar = v1[2 * i], ai = v1[2 * i + 1], br = v2[2 * i], bi = v2[2 * i + 1];
    if (ANALYTIC) {
      re = fma(ar, br, fma(-ai, bi, re));
      im = fma(ar, bi, fma(ai, br, im));
    } else {
      re = fma(ar, br, fma(ai, bi, re));
      im = fma(ai, br, fma(-ar, bi, im));
    }
  }
  return zd{re, im};
}

// Gram-Schmidt (shoot.f:621-639).  Returns w1, w2 and p = <U2, z1> for the P matrix.
template <bool ANALYTIC>
__device__ __forceinline__ void gram_schmidt(double (&u)[2][8], zd &w1, zd &w2, zd &p) {
  if (ANALYTIC) {
    w1 = zsqrt(inprod<true>(u[0], u[0]));
    zd r1 = zdiv(zd{1.0, 0.0}, w1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      zd z = zmul(zd{u[0][2 * i], u[0][2 * i + 1]}, r1);
      u[0][2 * i] = z.x; u[0][2 * i + 1] = z.y;
    }
    p = inprod<true>(u[1], u[0]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      zd pz = zmul(p, zd{u[0][2 * i], u[0][2 * i + 1]});
      u[1][2 * i] -= pz.x; u[1][2 * i + 1] -= pz.y;
    }
    w2 = zsqrt(inprod<true>(u[1], u[1]));
    zd r2 = zdiv(zd{1.0, 0.0}, w2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      zd z = zmul(zd{u[1][2 * i], u[1][2 * i + 1]}, r2);
      u[1][2 * i] = z.x; u[1][2 * i + 1] = z.y;
    }
  } else {
    double aa = inprod<false>(u[0], u[0]).x;
    double s1 = sqrt(aa);
    w1 = zd{s1, 0.0};
    double r1 = 1.0 / s1;
#pragma unroll
    for (int n = 0; n < 8; ++n) u[0][n] *= r1;
    p = inprod<false>(u[1], u[0]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double zr = u[0][2 * i], zi = u[0][2 * i + 1];
      u[1][2 * i] -= p.x * zr - p.y * zi;
      u[1][2 * i + 1] -= p.x * zi + p.y * zr;
    }
    double ee = inprod<false>(u[1], u[1]).x;
    double s2 = sqrt(ee);
    w2 = zd{s2, 0.0};
    double r2 = 1.0 / s2;
#pragma unroll
    for (int n = 0; n < 8; ++n) u[1][n] *= r2;
  }
}

// orthonormality test (shoot.f:593-606 and its three variants), see header note
template <bool ANALYTIC>
__device__ __forceinline__ bool need_norm(const double (&u)[2][8], double thr, int strict) {
  double lhs, rhs;
  if (ANALYTIC) {
    zd a = inprod<true>(u[0], u[0]), b = inprod<true>(u[1], u[1]), c = inprod<true>(u[0], u[1]);
    lhs = fma(c.x, c.x, c.y * c.y);
    const double a2 = fma(a.x, a.x, a.y * a.y), b2 = fma(b.x, b.x, b.y * b.y);
    // The two square roots below were 17 % of the continuation-line kernel's stall samples (every step, on the path to
    // the branch).  The squared form of the same test decides all but a 1e-12-wide band around the threshold, where
    // the rounding of the two forms (a few 1e-16 each) could differ; only there is the original expression evaluated.
    // The quick form is only trusted while its right-hand side is far inside the normal range (relative rounding
    // errors); overflow, underflow and NaN take the original expression: identical decisions.
    {
      const double l2 = lhs * lhs, r2 = (thr * thr) * (a2 * b2);
      if (r2 > 1.0e-250 && r2 < 1.0e250) {
        if (l2 > r2 * 1.000000000001) return true;
        if (l2 < r2 * 0.999999999999) return false;
      }
    }
    rhs = thr * (sqrt(a2) * sqrt(b2));
  } else {
    double aa = 0.0, bb = 0.0;
#pragma unroll
    for (int n = 0; n < 8; ++n) { aa = fma(u[0][n], u[0][n], aa); bb = fma(u[1][n], u[1][n], bb); }
    zd c = inprod<false>(u[0], u[1]);
    lhs = fma(c.x, c.x, c.y * c.y);
    rhs = thr * (aa * bb);
  }
  return strict ? (lhs > rhs) : (lhs >= rhs);
}

__device__ __forceinline__ PassConst pass_const(zd alpha, double Re, zd cw) {
  PassConst pc;
  zd a2 = zmul(alpha, alpha);
  zd k = zd{-alpha.y * Re, alpha.x * Re};  // (0,1)*alpha*Re
  zd kc = zmul(k, cw);
  pc.a2r = a2.x; pc.a2i = a2.y; pc.kr = k.x; pc.ki = k.y;
  pc.g0r = a2.x - kc.x; pc.g0i = a2.y - kc.y;
  return pc;
}

// initial vectors: free-stream asymptotes (BL_IC contebl.f:269-279) or e3, e4 (shoot.f:88-96)
__device__ __forceinline__ void initial_vectors(double (&u)[2][8], const ShootArgs &a, zd alpha,
                                                double Re, zd cw) {
  if (a.bl_ic) {
    zd a2 = zmul(alpha, alpha), a3 = zmul(a2, alpha);
    zd ea = zexp(zd{-alpha.x * a.tf, -alpha.y * a.tf});
    zd v;
    u[0][0] = ea.x; u[0][1] = ea.y;
    v = zmul(zneg(alpha), ea); u[0][2] = v.x; u[0][3] = v.y;
    v = zmul(a2, ea); u[0][4] = v.x; u[0][5] = v.y;
    v = zmul(zneg(a3), ea); u[0][6] = v.x; u[0][7] = v.y;
    zd k = zd{-alpha.y * Re, alpha.x * Re};
    zd g = zsqrt(zadd(a2, zmul(k, zd{1.0 - cw.x, -cw.y})));
    zd g2 = zmul(g, g), g3 = zmul(g2, g);
    zd eg = zexp(zd{-g.x * a.tf, -g.y * a.tf});
    u[1][0] = eg.x; u[1][1] = eg.y;
    v = zmul(zneg(g), eg); u[1][2] = v.x; u[1][3] = v.y;
    v = zmul(g2, eg); u[1][4] = v.x; u[1][5] = v.y;
    v = zmul(zneg(g3), eg); u[1][6] = v.x; u[1][7] = v.y;
  } else {
#pragma unroll
    for (int n = 0; n < 8; ++n) { u[0][n] = 0.0; u[1][n] = 0.0; }
    u[0][4] = 1.0;  // e3
    u[1][6] = 1.0;  // e4
  }
}

// One integration across the layer at the current eigenvalue guess; returns the
// wall residual err (shoot.f:699-700,716) and the normalisation count qend.
template <int INTEG, bool ANALYTIC, bool REAL_ALPHA>
__device__ __forceinline__ void integrate_pass(const ShootArgs &a, zd alpha, double Re, zd cw,
                                               zd &err, int &qend) {
  constexpr int NS = (INTEG == OSB_INT_CK45) ? 6 : 3;
  const PassConst pc = pass_const(alpha, Re, cw);
  double u[2][8];
  initial_vectors(u, a, alpha, Re, cw);
  int q = 0;
  const double2 *__restrict__ trow = a.tab;
  // k = 0 is the Gram-Schmidt of the initial vectors (shoot.f:490-515); k >= 1
  // are the integration steps.  One Gram-Schmidt site keeps the code small.
  for (int k = a.bl_ic ? 0 : 1; k <= a.nstep; ++k) {
    bool norm;
    if (k > 0) {
#ifdef OSB_CK45_SEQUENTIAL
      if (INTEG == OSB_INT_CK45) step_ck45_seq<REAL_ALPHA>(u, trow, pc, a);
#else
      if (INTEG == OSB_INT_CK45) step_ck45<REAL_ALPHA>(u, trow, pc, a);
#endif
      else step_rk4<REAL_ALPHA>(u, trow, pc, a);
      trow += NS;
      norm = need_norm<ANALYTIC>(u, a.thr, a.strict) || (k == a.nstep);
      q += norm ? 1 : 0;
    } else {
      norm = true;
    }
    if (norm) {
      zd w1, w2, p;
      gram_schmidt<ANALYTIC>(u, w1, w2, p);
    }
  }
  qend = q;
  // B(1) = -U(1,2)/U(1,1), B(2) = 1; err = U(2,1) B(1) + U(2,2) B(2)
  zd B1 = zneg(zdiv(zd{u[1][0], u[1][1]}, zd{u[0][0], u[0][1]}));
  err = zadd(zmul(zd{u[0][2], u[0][3]}, B1), zd{u[1][2], u[1][3]});
}

// secant / Muller update of the eigenvalue, shoot.f:713-744
struct IterState {
  zd ev, cm1, cm2, errm1, errm2;
  int icount;
  int cm1_defined;
};

__device__ __forceinline__ void update_eigenvalue(IterState &s, zd err, zd ctemp, int first_perturb) {
  if (s.icount == 1) {
    s.cm2 = s.ev;
    s.errm2 = err;
    if (first_perturb) s.ev = zd{s.ev.x * 1.01, s.ev.y * 1.01};  // orrspace.f:459
    else s.ev = zd{s.ev.x * .9999, s.ev.y};                       // shoot.f:718
  } else if (s.icount == 2) {
    s.cm1 = s.ev; s.cm1_defined = 1;
    s.errm1 = err;
    double den = s.ev.x - s.cm2.x;
    zd fd = zd{(err.x - s.errm2.x) / den, (err.y - s.errm2.y) / den};
    s.ev = zsub(s.ev, zdiv(err, fd));
  } else {
    zd c = s.ev;
    zd qt = zdiv(zsub(c, s.cm1), zsub(s.cm1, s.cm2));
    zd one_qt = zd{1.0 + qt.x, qt.y};
    zd qt2 = zmul(qt, qt);
    zd At = zadd(zsub(zmul(qt, err), zmul(zmul(qt, one_qt), s.errm1)), zmul(qt2, s.errm2));
    zd Bt = zadd(zsub(zmul(zd{2.0 * qt.x + 1.0, 2.0 * qt.y}, err), zmul(zmul(one_qt, one_qt), s.errm1)),
                 zmul(qt2, s.errm2));
    zd Ct = zmul(one_qt, err);
    zd D = zsqrt(zsub(zmul(Bt, Bt), zmul(zscale(4.0, At), Ct)));
    zd dp = zadd(Bt, D), dm = zsub(Bt, D);
    zd den = (zabs(dp) > zabs(dm)) ? dp : dm;
    zd num = zmul(zscale(2.0, zsub(ctemp, s.cm1)), Ct);
    s.ev = zsub(ctemp, zdiv(num, den));
    s.cm2 = s.cm1;
    s.cm1 = ctemp;
    s.errm2 = s.errm1;
    s.errm1 = err;
  }
}

__device__ __forceinline__ bool keep_going(const ShootArgs &a, const IterState &s, zd err) {
  double dc = s.cm1_defined ? zabs(zsub(s.ev, s.cm1)) : 1.0;
  double ae = zabs(err);
  if (a.loop_rule == 0) return ((ae >= a.errtol) || (dc >= a.eigtol)) && (s.icount <= a.maxcount);
  return (ae >= a.errtol) && (s.icount <= a.maxcount) && (dc >= a.eigtol);
}

template <int INTEG, bool ANALYTIC>
__global__ void __launch_bounds__(128) shoot_kernel(const __grid_constant__ ShootArgs a) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // lanes past the end read the last line's inputs so that every warp is full for the
  // warp-wide vote below, then leave
  const bool active = gid < a.nlines;
  const int64_t line = active ? gid : a.nlines - 1;
  const double Re = a.Re[line];
  zd alpha, omega = zd{0.0, 0.0};
  IterState s;
  {
    double2 al = a.alpha[line];
    alpha = zd{al.x, al.y};
    if (a.spatial) {
      double2 om = a.omega0[line];
      omega = zd{om.x, om.y};
      s.ev = alpha;
    } else {
      double2 c0 = a.ev[line];
      s.ev = zd{c0.x, c0.y};
    }
  }
  // temporal sweeps normally have real alpha: take the cheaper RHS when the whole warp does
  const bool real_alpha = !a.spatial && __all_sync(0xffffffffu, alpha.y == 0.0);
  if (!active) return;
  // ONE loop, one pass per trip: a line that has finished a continuation point writes it out, seeds the next one
  // (orrspace.f:494) and joins the very next pass with the lines that are still iterating, so a warp runs as long as
  // the largest TOTAL pass count over its lines instead of the sum over the points of the per-point maxima (a loop
  // nest would give the latter; sweep.inp as 4096 lines: 1005 against 656 passes for the slowest warp).
  int it = 0;
  bool bad = false;
  zd err, ctemp;
  int qend;
  auto fresh = [&]() {
    s.icount = 1;
    s.cm1_defined = a.cm1_defined;
    s.cm1 = zd{s.ev.x + 1.0, s.ev.y};
    s.cm2 = zd{0.0, 0.0}; s.errm1 = zd{0.0, 0.0}; s.errm2 = zd{0.0, 0.0};
    err = zd{1.0, 0.0}; ctemp = s.ev;
    qend = 0;
    bad = false;
  };
  fresh();
  while (true) {
    while (bad || !keep_going(a, s, err)) {
      const int64_t out = line * a.niter + it;
      a.ev[out] = make_double2(ctemp.x, ctemp.y);
      a.passes[out] = s.icount - 1;
      a.qend[out] = qend;
      a.status[out] = bad ? OSB_ST_NONFINITE : (s.icount <= a.maxcount ? OSB_ST_CONVERGED : OSB_ST_MAXCOUNT);
      if (a.resid)
        a.resid[out] = make_double2(zabs(err), s.cm1_defined ? zabs(zsub(s.ev, s.cm1)) : 0.0);
      if (a.spatial) {
        s.ev = ctemp;  // orrspace.f:494: alpha = ctemp seeds the next omega
        omega.x += a.domega[line];  // orrspace.f:192-194
      }
      if (++it >= a.niter) return;
      fresh();
    }
    ctemp = s.ev;
    {
      // spatial: the unknown is alpha and c = omega/alpha (orrspace.f:290,645)
      const zd al = a.spatial ? s.ev : alpha;
      const zd cw = a.spatial ? zdiv(omega, s.ev) : s.ev;
      if (real_alpha) integrate_pass<INTEG, ANALYTIC, true>(a, al, Re, cw, err, qend);
      else integrate_pass<INTEG, ANALYTIC, false>(a, al, Re, cw, err, qend);
    }
    update_eigenvalue(s, err, ctemp, a.first_perturb);
    if (a.history && s.icount <= a.hist_cap) {
      const int64_t out = line * a.niter + it;
      zd shown = a.show_updated ? s.ev : ctemp;
      a.history[out * a.hist_cap + (s.icount - 1)] = make_double4(shown.x, shown.y, err.x, err.y);
    }
    s.icount += 1;
    if (!(zfinite(s.ev) && zfinite(err))) bad = true;
  }
}

// ---------------------------------------------------------------------------
// K2 with divergent convergence compacted through a shared-memory work queue
// (temporal sweeps).  Pass counts differ from point to point (5..20); with one
// point per thread for the whole solve a warp runs as long as its slowest lane
// (ncu round 1: 26.97 of 32 lanes active).  Here CTAs are persistent and the unit
// of work is one PASS: after every pass the CTA packs the points that need
// another pass into its lowest slots (order preserving, so neighbours in the
// sweep stay neighbours and keep normalising at the same steps) and refills the
// freed slots with fresh points taken from a global counter.  The per-point
// arithmetic is untouched, so results do not depend on the schedule.
// ---------------------------------------------------------------------------
constexpr int QT = 128;  // threads (= slots) per CTA

struct SlotState {  // what a point carries from one pass to the next
  zd ev, cm1, cm2, errm1, errm2, alpha;
  double Re;
  long long line;
  int icount, cm1_defined;
};

template <int INTEG, bool ANALYTIC>
#ifndef OSB_QUEUE_MIN_BLOCKS
#define OSB_QUEUE_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(QT, OSB_QUEUE_MIN_BLOCKS) shoot_queue_kernel(const __grid_constant__ ShootArgs a,
                                                         unsigned long long *__restrict__ counter) {
  __shared__ SlotState slots[QT];
  __shared__ int warp_keep[QT / 32];
  __shared__ long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  SlotState st;
  st.line = -1;
  bool first = true;
  for (;;) {
    // ---- compaction + refill
    bool keep = false;
    if (!first && st.line >= 0) keep = true;  // st holds a point that needs another pass
    first = false;
    const unsigned kmask = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_keep[warp] = __popc(kmask);
    __syncthreads();
    int before = 0, nkeep = 0;
#pragma unroll
    for (int w = 0; w < QT / 32; ++w) {
      if (w < warp) before += warp_keep[w];
      nkeep += warp_keep[w];
    }
    if (keep) slots[before + __popc(kmask & ((1u << lane) - 1u))] = st;
    if (tid == 0) s_base = (long long)atomicAdd(counter, (unsigned long long)(QT - nkeep));
    __syncthreads();
    if (tid < nkeep) {
      st = slots[tid];
    } else {
      const long long line = s_base + (tid - nkeep);
      if (line < a.nlines) {
        const double2 al = a.alpha[line], c0 = a.ev[line];
        st.line = line;
        st.alpha = zd{al.x, al.y};
        st.Re = a.Re[line];
        st.ev = zd{c0.x, c0.y};
        st.icount = 1;
        st.cm1_defined = a.cm1_defined;
        st.cm1 = zd{st.ev.x + 1.0, st.ev.y};
        st.cm2 = zd{0.0, 0.0}; st.errm1 = zd{0.0, 0.0}; st.errm2 = zd{0.0, 0.0};
      } else {
        st.line = -1;
      }
    }
    const bool active = st.line >= 0;
    if (__syncthreads_or(active) == 0) break;  // queue drained and nothing left to continue
    // ---- one pass
    const bool real_alpha = __all_sync(0xffffffffu, !active || st.alpha.y == 0.0);
    if (active) {
      zd err;
      int qend = 0;
      const zd ctemp = st.ev;
      if (real_alpha) integrate_pass<INTEG, ANALYTIC, true>(a, st.alpha, st.Re, st.ev, err, qend);
      else integrate_pass<INTEG, ANALYTIC, false>(a, st.alpha, st.Re, st.ev, err, qend);
      IterState it;
      it.ev = st.ev; it.cm1 = st.cm1; it.cm2 = st.cm2; it.errm1 = st.errm1; it.errm2 = st.errm2;
      it.icount = st.icount; it.cm1_defined = st.cm1_defined;
      update_eigenvalue(it, err, ctemp, a.first_perturb);
      if (a.history && it.icount <= a.hist_cap) {
        const zd shown = a.show_updated ? it.ev : ctemp;
        a.history[st.line * a.hist_cap + (it.icount - 1)] = make_double4(shown.x, shown.y, err.x, err.y);
      }
      it.icount += 1;
      int status = -1;
      if (!(zfinite(it.ev) && zfinite(err))) status = OSB_ST_NONFINITE;
      else if (!keep_going(a, it, err)) status = (it.icount <= a.maxcount) ? OSB_ST_CONVERGED : OSB_ST_MAXCOUNT;
      if (status >= 0) {  // the point is done: report ctemp, the last value integrated
        a.ev[st.line] = make_double2(ctemp.x, ctemp.y);
        a.passes[st.line] = it.icount - 1;
        a.qend[st.line] = qend;
        a.status[st.line] = status;
        if (a.resid)
          a.resid[st.line] = make_double2(zabs(err), it.cm1_defined ? zabs(zsub(it.ev, it.cm1)) : 0.0);
        st.line = -1;
      } else {
        st.ev = it.ev; st.cm1 = it.cm1; st.cm2 = it.cm2; st.errm1 = it.errm1; st.errm2 = it.errm2;
        st.icount = it.icount; st.cm1_defined = it.cm1_defined;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K2 on a LANE PAIR: the two solution vectors of a point live in two adjacent lanes.
// For sweeps that cannot fill the GPU with one thread per point -- the spatial continuation
// lines of orrspace (sequential inside a line), single-point driver runs, small
// batches -- this doubles the parallelism.  Each lane integrates its own vector; after every
// step the pair exchanges the vectors with shuffles, so that the orthonormality test, the
// Gram-Schmidt step and the wall residual are evaluated (redundantly, identically) on both
// lanes with the very same code as the one-thread kernel: results are bit-identical to it.
// ---------------------------------------------------------------------------
template <bool REAL_ALPHA>
__device__ __forceinline__ void step_ck45_one(double (&y)[8], const double2 *__restrict__ trow,
                                              const PassConst &pc, const ShootArgs &a) {
  double f[6][8];
#pragma unroll
  for (int m = 0; m < 6; ++m) {
    const StageCoef sc = stage_coef<REAL_ALPHA>(pc, __ldg(&trow[m]));
    double yt[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      double acc = y[n];
#pragma unroll
      for (int k = 0; k < m; ++k) acc = fma(a.bh[m * (m - 1) / 2 + k], f[k][n], acc);
      yt[n] = acc;
    }
#pragma unroll
    for (int n = 0; n < 6; ++n) f[m][n] = yt[n + 2];
    rhs4<REAL_ALPHA>(sc, pc, yt, f[m][6], f[m][7]);
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    double acc = y[n];
    acc = fma(a.ch[0], f[0][n], acc);
    acc = fma(a.ch[2], f[2][n], acc);
    acc = fma(a.ch[3], f[3][n], acc);
    acc = fma(a.ch[5], f[5][n], acc);
    y[n] = acc;
  }
}

template <bool REAL_ALPHA>
__device__ __forceinline__ void step_rk4_one(double (&y)[8], const double2 *__restrict__ trow,
                                             const PassConst &pc, const ShootArgs &a) {
  const double hs = a.hs, h2 = 0.5 * a.hs, h6 = a.hs / 6.0, h3 = a.hs / 3.0;
  const StageCoef s0 = stage_coef<REAL_ALPHA>(pc, __ldg(&trow[0]));
  const StageCoef s1 = stage_coef<REAL_ALPHA>(pc, __ldg(&trow[1]));
  const StageCoef s2 = stage_coef<REAL_ALPHA>(pc, __ldg(&trow[2]));
  double f[8], q[8], acc[8];
#pragma unroll
  for (int n = 0; n < 6; ++n) f[n] = y[n + 2];
  rhs4<REAL_ALPHA>(s0, pc, y, f[6], f[7]);
#pragma unroll
  for (int n = 0; n < 8; ++n) { acc[n] = fma(h6, f[n], y[n]); q[n] = fma(h2, f[n], y[n]); }
#pragma unroll
  for (int n = 0; n < 6; ++n) f[n] = q[n + 2];
  rhs4<REAL_ALPHA>(s1, pc, q, f[6], f[7]);
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n] = fma(h3, f[n], acc[n]);
  {
    double q2[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) q2[n] = fma(h2, f[n], y[n]);
#pragma unroll
    for (int n = 0; n < 6; ++n) f[n] = q2[n + 2];
    rhs4<REAL_ALPHA>(s1, pc, q2, f[6], f[7]);
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) { acc[n] = fma(h3, f[n], acc[n]); q[n] = fma(hs, f[n], y[n]); }
#pragma unroll
  for (int n = 0; n < 6; ++n) f[n] = q[n + 2];
  rhs4<REAL_ALPHA>(s2, pc, q, f[6], f[7]);
#pragma unroll
  for (int n = 0; n < 8; ++n) y[n] = fma(h6, f[n], acc[n]);
}

// ---- round-1/2 form of the pair kernel (OSB_PAIR_V1=1), kept for A/B measurements and bit-identity tests
template <int INTEG, bool ANALYTIC, bool REAL_ALPHA>
__device__ __forceinline__ void integrate_pass_pair_v1(const ShootArgs &a, int v, zd alpha, double Re, zd cw,
                                                    zd &err, int &qend) {
  constexpr int NS = (INTEG == OSB_INT_CK45) ? 6 : 3;
  // pairs of one warp converge after different numbers of passes, so the exchange is
  // synchronised over the two lanes of the pair only
  const unsigned pmask = 3u << (threadIdx.x & 30u);
  const PassConst pc = pass_const(alpha, Re, cw);
  double u[2][8];
  initial_vectors(u, a, alpha, Re, cw);
  int q = 0;
  const double2 *__restrict__ trow = a.tab;
  for (int k = a.bl_ic ? 0 : 1; k <= a.nstep; ++k) {
    bool norm;
    if (k > 0) {
      // integrate my vector, then swap in the partner's
      double y[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) y[n] = v ? u[1][n] : u[0][n];
      if (INTEG == OSB_INT_CK45) step_ck45_one<REAL_ALPHA>(y, trow, pc, a);
      else step_rk4_one<REAL_ALPHA>(y, trow, pc, a);
      trow += NS;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const double o = __shfl_xor_sync(pmask, y[n], 1);
        u[0][n] = v ? o : y[n];
        u[1][n] = v ? y[n] : o;
      }
      norm = need_norm<ANALYTIC>(u, a.thr, a.strict) || (k == a.nstep);
      q += norm ? 1 : 0;
    } else {
      norm = true;
    }
    if (norm) {
      zd w1, w2, p;
      gram_schmidt<ANALYTIC>(u, w1, w2, p);
    }
  }
  qend = q;
  zd B1 = zneg(zdiv(zd{u[1][0], u[1][1]}, zd{u[0][0], u[0][1]}));
  err = zadd(zmul(zd{u[0][2], u[0][3]}, B1), zd{u[1][2], u[1][3]});
}

template <int INTEG, bool ANALYTIC>
__global__ void __launch_bounds__(64) shoot_pair_kernel_v1(const __grid_constant__ ShootArgs a) {
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = (int)(gtid & 1);
  const int64_t gline = gtid >> 1;
  const bool active = gline < a.nlines;
  const int64_t line = active ? gline : a.nlines - 1;
  const double Re = a.Re[line];
  zd alpha, omega = zd{0.0, 0.0};
  IterState s;
  {
    double2 al = a.alpha[line];
    alpha = zd{al.x, al.y};
    if (a.spatial) {
      double2 om = a.omega0[line];
      omega = zd{om.x, om.y};
      s.ev = alpha;
    } else {
      double2 c0 = a.ev[line];
      s.ev = zd{c0.x, c0.y};
    }
  }
  const bool real_alpha = !a.spatial && __all_sync(0xffffffffu, alpha.y == 0.0);
  for (int it = 0; it < a.niter; ++it) {
    const int64_t out = line * a.niter + it;
    s.icount = 1;
    s.cm1_defined = a.cm1_defined;
    s.cm1 = zd{s.ev.x + 1.0, s.ev.y};
    s.cm2 = zd{0.0, 0.0}; s.errm1 = zd{0.0, 0.0}; s.errm2 = zd{0.0, 0.0};
    zd err = zd{1.0, 0.0}, ctemp = s.ev;
    int qend = 0, status = OSB_ST_MAXCOUNT;
    while (true) {
      // both lanes of a pair hold identical iteration state, so they leave the loop together
      if (!keep_going(a, s, err)) {
        if (s.icount <= a.maxcount) status = OSB_ST_CONVERGED;
        break;
      }
      ctemp = s.ev;
      {
        const zd al = a.spatial ? s.ev : alpha;
        const zd cw = a.spatial ? zdiv(omega, s.ev) : s.ev;
        if (real_alpha) integrate_pass_pair_v1<INTEG, ANALYTIC, true>(a, v, al, Re, cw, err, qend);
        else integrate_pass_pair_v1<INTEG, ANALYTIC, false>(a, v, al, Re, cw, err, qend);
      }
      update_eigenvalue(s, err, ctemp, a.first_perturb);
      if (active && v == 0 && a.history && s.icount <= a.hist_cap) {
        zd shown = a.show_updated ? s.ev : ctemp;
        a.history[out * a.hist_cap + (s.icount - 1)] = make_double4(shown.x, shown.y, err.x, err.y);
      }
      s.icount += 1;
      if (!(zfinite(s.ev) && zfinite(err))) { status = OSB_ST_NONFINITE; break; }
    }
    if (active && v == 0) {
      a.ev[out] = make_double2(ctemp.x, ctemp.y);
      a.passes[out] = s.icount - 1;
      a.qend[out] = qend;
      a.status[out] = status;
      if (a.resid)
        a.resid[out] = make_double2(zabs(err), s.cm1_defined ? zabs(zsub(s.ev, s.cm1)) : 0.0);
    }
    if (a.spatial) {
      s.ev = ctemp;
      omega.x += a.domega[line];
    }
  }
}

// The pass as the pair kernel runs it.  A lane's own vector needs nothing from its partner until a Gram-Schmidt
// step happens, so step k+1 is started from the un-normalised result of step k while the exchange and the
// orthonormality test of step k are still in flight (one basic block: the scheduler interleaves the two dependent
// chains); only when the test asks for a normalisation (2-4 % of the steps) is step k+1 redone from the normalised
// vector.  Same operations on the same values as integrate_pass: bit-identical results, ~40 % less latency per step
// in the regime this kernel serves (one warp per scheduler, nothing else to overlap with).
template <int INTEG, bool ANALYTIC, bool REAL_ALPHA>
__device__ __forceinline__ void integrate_pass_pair(const ShootArgs &a, int v, zd alpha, double Re, zd cw,
                                                    zd &err, int &qend) {
  constexpr int NS = (INTEG == OSB_INT_CK45) ? 6 : 3;
  // pairs of one warp are at different passes of different continuation points, so the exchange is
  // synchronised over the two lanes of the pair only
  const unsigned pmask = 3u << (threadIdx.x & 30u);
  const PassConst pc = pass_const(alpha, Re, cw);
  double u[2][8];
  initial_vectors(u, a, alpha, Re, cw);
  if (a.bl_ic) {  // k = 0: Gram-Schmidt of the initial vectors (shoot.f:490-515)
    zd w1, w2, p;
    gram_schmidt<ANALYTIC>(u, w1, w2, p);
  }
  int q = 0;
  const double2 *__restrict__ trow = a.tab;
  double y[8];  // my vector after step k, not yet tested
#pragma unroll
  for (int n = 0; n < 8; ++n) y[n] = v ? u[1][n] : u[0][n];
  if (INTEG == OSB_INT_CK45) step_ck45_one<REAL_ALPHA>(y, trow, pc, a);
  else step_rk4_one<REAL_ALPHA>(y, trow, pc, a);
  trow += NS;
  for (int k = 1; k < a.nstep; ++k) {
    double ys[8];  // step k+1, valid unless step k ends in a normalisation
#pragma unroll
    for (int n = 0; n < 8; ++n) ys[n] = y[n];
    if (INTEG == OSB_INT_CK45) step_ck45_one<REAL_ALPHA>(ys, trow, pc, a);
    else step_rk4_one<REAL_ALPHA>(ys, trow, pc, a);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const double o = __shfl_xor_sync(pmask, y[n], 1);
      u[0][n] = v ? o : y[n];
      u[1][n] = v ? y[n] : o;
    }
    if (need_norm<ANALYTIC>(u, a.thr, a.strict)) {
      q += 1;
      zd w1, w2, p;
      gram_schmidt<ANALYTIC>(u, w1, w2, p);
#pragma unroll
      for (int n = 0; n < 8; ++n) ys[n] = v ? u[1][n] : u[0][n];
      if (INTEG == OSB_INT_CK45) step_ck45_one<REAL_ALPHA>(ys, trow, pc, a);
      else step_rk4_one<REAL_ALPHA>(ys, trow, pc, a);
    }
    trow += NS;
#pragma unroll
    for (int n = 0; n < 8; ++n) y[n] = ys[n];
  }
  // k = nstep always ends in a normalisation (shoot.f:607)
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const double o = __shfl_xor_sync(pmask, y[n], 1);
    u[0][n] = v ? o : y[n];
    u[1][n] = v ? y[n] : o;
  }
  {
    zd w1, w2, p;
    gram_schmidt<ANALYTIC>(u, w1, w2, p);
  }
  qend = q + 1;
  zd B1 = zneg(zdiv(zd{u[1][0], u[1][1]}, zd{u[0][0], u[0][1]}));
  err = zadd(zmul(zd{u[0][2], u[0][3]}, B1), zd{u[1][2], u[1][3]});
}

// One pass per trip of ONE loop: a pair that has finished a continuation point writes it out, seeds the next point
// (orrspace.f:494) and joins the very next pass together with the pairs that are still iterating.  With a loop nest
// (points outside, passes inside) a warp spent, at every point, as many passes as its slowest line; now a warp runs
// as long as the largest TOTAL over its lines.  a.pair_lpw lines share a warp (the lanes above them exit): fewer
// lines per warp shorten that maximum further and spread a small sweep over more schedulers.
template <int INTEG, bool ANALYTIC>
__global__ void __launch_bounds__(64) shoot_pair_kernel(const __grid_constant__ ShootArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int v = lane & 1;
  const int64_t gline = gwarp * a.pair_lpw + (lane >> 1);
  const bool active = gline < a.nlines && (lane >> 1) < a.pair_lpw;
  const int64_t line = gline < a.nlines ? gline : a.nlines - 1;
  const double Re = a.Re[line];
  zd alpha, omega = zd{0.0, 0.0};
  IterState s;
  {
    double2 al = a.alpha[line];
    alpha = zd{al.x, al.y};
    if (a.spatial) {
      double2 om = a.omega0[line];
      omega = zd{om.x, om.y};
      s.ev = alpha;
    } else {
      double2 c0 = a.ev[line];
      s.ev = zd{c0.x, c0.y};
    }
  }
  const bool real_alpha = !a.spatial && __all_sync(0xffffffffu, alpha.y == 0.0);
  if (!active) return;  // both lanes of a pair together; the exchange is pair-masked
  int it = 0;
  bool bad = false;
  zd err, ctemp;
  int qend;
  auto fresh = [&]() {
    s.icount = 1;
    s.cm1_defined = a.cm1_defined;
    s.cm1 = zd{s.ev.x + 1.0, s.ev.y};
    s.cm2 = zd{0.0, 0.0}; s.errm1 = zd{0.0, 0.0}; s.errm2 = zd{0.0, 0.0};
    err = zd{1.0, 0.0}; ctemp = s.ev;
    qend = 0;
    bad = false;
  };
  fresh();
  while (true) {
    // both lanes of a pair hold identical iteration state, so they take the same way through here
    while (bad || !keep_going(a, s, err)) {
      const int64_t out = line * a.niter + it;
      if (v == 0) {
        a.ev[out] = make_double2(ctemp.x, ctemp.y);
        a.passes[out] = s.icount - 1;
        a.qend[out] = qend;
        a.status[out] = bad ? OSB_ST_NONFINITE : (s.icount <= a.maxcount ? OSB_ST_CONVERGED : OSB_ST_MAXCOUNT);
        if (a.resid)
          a.resid[out] = make_double2(zabs(err), s.cm1_defined ? zabs(zsub(s.ev, s.cm1)) : 0.0);
      }
      if (a.spatial) {
        s.ev = ctemp;
        omega.x += a.domega[line];
      }
      if (++it >= a.niter) return;
      fresh();
    }
    ctemp = s.ev;
    {
      const zd al = a.spatial ? s.ev : alpha;
      const zd cw = a.spatial ? zdiv(omega, s.ev) : s.ev;
      if (real_alpha) integrate_pass_pair<INTEG, ANALYTIC, true>(a, v, al, Re, cw, err, qend);
      else integrate_pass_pair<INTEG, ANALYTIC, false>(a, v, al, Re, cw, err, qend);
    }
    update_eigenvalue(s, err, ctemp, a.first_perturb);
    if (v == 0 && a.history && s.icount <= a.hist_cap) {
      const int64_t out = line * a.niter + it;
      zd shown = a.show_updated ? s.ev : ctemp;
      a.history[out * a.hist_cap + (s.icount - 1)] = make_double4(shown.x, shown.y, err.x, err.y);
    }
    s.icount += 1;
    if (!(zfinite(s.ev) && zfinite(err))) bad = true;
  }
}

// ---------------------------------------------------------------------------
// K2b: eigenfunction pass (shoot.f:762-810; conte shoot.f:338-385).  One thread
// per requested point: re-integrates at the converged eigenvalue keeping the
// (orthonormalised) trajectory, the P matrices and the normalisation abscissae,
// then back-substitutes B and applies the reference's vmax scan (oracle D-2).
// ---------------------------------------------------------------------------
template <int INTEG, bool ANALYTIC>
__global__ void __launch_bounds__(64) eigfun_kernel(const __grid_constant__ EigfunArgs e, int64_t npts) {
  constexpr int NS = (INTEG == OSB_INT_CK45) ? 6 : 3;
  const ShootArgs &a = e.s;
  const int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= npts) return;
  const int nstep = a.nstep;
  const size_t n1 = (size_t)nstep + 1;
  double2 *Ut = e.U + pt * n1 * 8;
  double2 *Pq = e.P + pt * n1 * 4;
  double *tq = e.tq + pt * n1;
  double2 *yo = e.y + pt * n1 * 4;
  const double Re = a.Re[pt];
  zd alpha, cw;
  {
    double2 ev = e.eig[pt];
    if (a.spatial) {
      double2 om = e.omega[pt];
      alpha = zd{ev.x, ev.y};
      cw = zdiv(zd{om.x, om.y}, alpha);
    } else {
      double2 al = a.alpha[pt];
      alpha = zd{al.x, al.y};
      cw = zd{ev.x, ev.y};
    }
  }
  const PassConst pc = pass_const(alpha, Re, cw);
  const bool bl = a.hs < 0.0;
  double u[2][8];
  initial_vectors(u, a, alpha, Re, cw);
  int q = 0;
  tq[0] = bl ? a.tf : e.t0;
  const double2 *__restrict__ trow = a.tab;
  for (int k = a.bl_ic ? 0 : 1; k <= nstep; ++k) {
    bool norm;
    double t = bl ? __dsub_rn(a.tf, __dmul_rn(e.habs, (double)k)) : __dadd_rn(e.t0, __dmul_rn(e.habs, (double)k));
    if (k > 0) {
      if (INTEG == OSB_INT_CK45) step_ck45<false>(u, trow, pc, a);  // same bits as the real-alpha path
      else step_rk4<false>(u, trow, pc, a);
      trow += NS;
      norm = need_norm<ANALYTIC>(u, a.thr, a.strict) || (k == nstep);
    } else {
      norm = true;
    }
    if (norm) {
      zd w1, w2, p;
      gram_schmidt<ANALYTIC>(u, w1, w2, p);
      if (k > 0) {
        q += 1;
        tq[q] = t;
        zd P11 = zdiv(zd{1.0, 0.0}, w1), P22 = zdiv(zd{1.0, 0.0}, w2);
        zd P21 = zneg(zmul(zdiv(p, w2), P11));  // shoot.f:652-653
        Pq[(size_t)q * 4 + 0] = make_double2(P11.x, P11.y);
        Pq[(size_t)q * 4 + 1] = make_double2(P21.x, P21.y);
        Pq[(size_t)q * 4 + 2] = make_double2(P22.x, P22.y);
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      Ut[(size_t)k * 8 + n] = make_double2(u[0][2 * n], u[0][2 * n + 1]);
      Ut[(size_t)k * 8 + 4 + n] = make_double2(u[1][2 * n], u[1][2 * n + 1]);
    }
  }
  if (!a.bl_ic) {  // k = 0 of the channel flow: the unit vectors themselves
    double v[2][8];
    initial_vectors(v, a, alpha, Re, cw);
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      Ut[n] = make_double2(v[0][2 * n], v[0][2 * n + 1]);
      Ut[4 + n] = make_double2(v[1][2 * n], v[1][2 * n + 1]);
    }
  }
  const int qend = q;
  if (e.qend) e.qend[pt] = qend;
  // B(:,qend): phi(wall) = 0 enforced (shoot.f:699-700)
  zd Bq1 = zneg(zdiv(zd{u[1][0], u[1][1]}, zd{u[0][0], u[0][1]})), Bq2 = zd{1.0, 0.0};
  auto apply_PT = [&](int qq, zd b1, zd b2, zd &o1, zd &o2) {
    // B(m,q-1) = sum_j P(j,m,q) B(j,q)   (shoot.f:772-777)
    double2 p11 = Pq[(size_t)qq * 4 + 0], p21 = Pq[(size_t)qq * 4 + 1], p22 = Pq[(size_t)qq * 4 + 2];
    o1 = zadd(zmul(zd{p11.x, p11.y}, b1), zmul(zd{p21.x, p21.y}, b2));
    o2 = zmul(zd{p22.x, p22.y}, b2);
  };
  auto combine = [&](int k, zd b1, zd b2) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      double2 u1 = Ut[(size_t)k * 8 + n], u2 = Ut[(size_t)k * 8 + 4 + n];
      zd v = zadd(zmul(zd{u1.x, u1.y}, b1), zmul(zd{u2.x, u2.y}, b2));
      yo[(size_t)k * 4 + n] = make_double2(v.x, v.y);
    }
  };
  combine(nstep, Bq1, Bq2);
  zd Bp1, Bp2;
  apply_PT(q, Bq1, Bq2, Bp1, Bp2);
  zd vmax = zd{0.0, 0.0};
  for (int k = nstep - 1; k >= 0; --k) {
    double t = bl ? __dsub_rn(a.tf, __dmul_rn(e.habs, (double)k)) : __dadd_rn(e.t0, __dmul_rn(e.habs, (double)k));
    bool back = bl ? (t > tq[q - 1]) : (t < tq[q - 1]);
    if (back) {
      q -= 1;
      Bq1 = Bp1; Bq2 = Bp2;
      apply_PT(q, Bq1, Bq2, Bp1, Bp2);
    }
    combine(k, Bp1, Bp2);
    // the reference's scan reads y(n+1,k) == y(1,k+1)  (shoot.f:795-797)
    double2 y1 = yo[(size_t)(k + 1) * 4];
    if (zabs(zd{y1.x, y1.y}) > zabs(vmax)) vmax = zd{y1.x, y1.y};
  }
  e.vmax[pt] = make_double2(vmax.x, vmax.y);
}

cudaError_t launch_eigfun(const EigfunArgs &e, int integrator, int analytic_inprod, int64_t npts,
                          cudaStream_t st) {
  if (npts <= 0) return cudaSuccess;
  int threads = 32;
  unsigned blocks = (unsigned)((npts + threads - 1) / threads);
  if (integrator == OSB_INT_CK45) {
    if (analytic_inprod) eigfun_kernel<OSB_INT_CK45, true><<<blocks, threads, 0, st>>>(e, npts);
    else eigfun_kernel<OSB_INT_CK45, false><<<blocks, threads, 0, st>>>(e, npts);
  } else {
    if (analytic_inprod) eigfun_kernel<OSB_INT_RK4, true><<<blocks, threads, 0, st>>>(e, npts);
    else eigfun_kernel<OSB_INT_RK4, false><<<blocks, threads, 0, st>>>(e, npts);
  }
  return cudaGetLastError();
}

// persistent-CTA launch of the queue kernel; `counter` is a zeroed device word
static cudaError_t launch_shoot_queue(const ShootArgs &a, int integrator, int analytic_inprod,
                                      unsigned long long *counter, cudaStream_t st) {
  int dev = 0, sms = 0, per = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const void *fn = integrator == OSB_INT_CK45
                       ? (analytic_inprod ? (const void *)shoot_queue_kernel<OSB_INT_CK45, true>
                                          : (const void *)shoot_queue_kernel<OSB_INT_CK45, false>)
                       : (analytic_inprod ? (const void *)shoot_queue_kernel<OSB_INT_RK4, true>
                                          : (const void *)shoot_queue_kernel<OSB_INT_RK4, false>);
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fn, QT, 0);
  if (e != cudaSuccess) return e;
  if (per < 1) per = 1;
  const int64_t want = (a.nlines + QT - 1) / QT;
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * per);
  if (integrator == OSB_INT_CK45) {
    if (analytic_inprod) shoot_queue_kernel<OSB_INT_CK45, true><<<grid, QT, 0, st>>>(a, counter);
    else shoot_queue_kernel<OSB_INT_CK45, false><<<grid, QT, 0, st>>>(a, counter);
  } else {
    if (analytic_inprod) shoot_queue_kernel<OSB_INT_RK4, true><<<grid, QT, 0, st>>>(a, counter);
    else shoot_queue_kernel<OSB_INT_RK4, false><<<grid, QT, 0, st>>>(a, counter);
  }
  return cudaGetLastError();
}

cudaError_t launch_shoot(const ShootArgs &a, int integrator, int analytic_inprod, cudaStream_t st,
                         int *nlaunch, unsigned long long *queue_counter) {
  if (a.nlines <= 0) return cudaSuccess;
  // OSB_STRICT=1: the strict-parity build (reference operation order, no FMA; shoot_strict.cu) -- bit-identical
  // to the CPU restatement of the reference, ~3x slower; a verification switch, never a default
  if (getenv("OSB_STRICT")) {
    if (nlaunch) *nlaunch += 1;
    return launch_shoot_strict(a, integrator, analytic_inprod, st);
  }
  // Which kernel (measured on B200, tools/bench_small_batches.py, config-5 points):
  //   <= 9472 lines      two lanes per line: 1.2-1.25x over one thread per line
  //   up to ~65k points  one thread per point: the GPU is not yet oversubscribed, compaction only costs
  //   beyond             the compacting persistent-CTA kernel: +9 % at 131k points, +11 % at 524k, more at 16.7M
  // OSB_QUEUE_MIN overrides the last threshold (tests run the queue kernel on small batches with it).
  int64_t queue_min = 65536;
  if (const char *e = getenv("OSB_QUEUE_MIN")) queue_min = std::max<int64_t>(4 * QT, atoll(e));
  if (queue_counter && !a.spatial && a.niter == 1 && a.nlines >= queue_min) {
    if (nlaunch) *nlaunch += 1;
    return launch_shoot_queue(a, integrator, analytic_inprod, queue_counter, st);
  }
  int64_t pair_max = 148 * 64;
  if (const char *e = getenv("OSB_PAIR_MAX")) pair_max = atoll(e);  // measurement knob
  if (a.nlines <= pair_max && !getenv("OSB_NO_PAIR")) {
    const int threads = 64;
    // Lines per warp (measured on B200, tools/bench_pair2.py, tools/bench_pair_lpw.py).  A warp instruction costs the FP64
    // pipe the same 2.25 cycles whether 2 or 32 lanes are active and a lone warp keeps its scheduler's pipe ~50 % busy,
    // so the count that matters is warps per scheduler, and there is a cliff at one: with 148 SMs x 4 schedulers the
    // lines are spread over at most 592 warps (4096 contebl points: 8 lines per warp = 512 warps 367 k eig/s, 6 = 683
    // warps 241 k, 16 = 256 warps 312 k; 6144 points: 12 per warp 498 k, 10 per warp 416 k; continuation lines, 4096 x
    // 50 omega: 8 per warp 367 k points/s, 16: 313 k, 4: 339 k).  Fewer lines per warp also means fewer divergent
    // Gram-Schmidt events.  Single solves at 16 lines per warp (> 8880 points) gain nothing from the speculative step
    // (9472 points: 708 k eig/s in the plain form against 681 k).  OSB_PAIR_LPW / OSB_PAIR_V1=1|0 override.
    int lpw = (int)((a.nlines + 591) / 592);
    if (const char *e = getenv("OSB_PAIR_LPW")) lpw = atoi(e);
    lpw = std::min(16, std::max(1, lpw));
    bool v1 = a.niter == 1 && lpw >= 16;
    if (const char *e = getenv("OSB_PAIR_V1")) v1 = atoi(e) != 0;
    if (v1) {
      const unsigned blocks = (unsigned)((2 * a.nlines + threads - 1) / threads);
      if (integrator == OSB_INT_CK45) {
        if (analytic_inprod) shoot_pair_kernel_v1<OSB_INT_CK45, true><<<blocks, threads, 0, st>>>(a);
        else shoot_pair_kernel_v1<OSB_INT_CK45, false><<<blocks, threads, 0, st>>>(a);
      } else {
        if (analytic_inprod) shoot_pair_kernel_v1<OSB_INT_RK4, true><<<blocks, threads, 0, st>>>(a);
        else shoot_pair_kernel_v1<OSB_INT_RK4, false><<<blocks, threads, 0, st>>>(a);
      }
      if (nlaunch) *nlaunch += 1;
      return cudaGetLastError();
    }
    ShootArgs b = a;
    b.pair_lpw = lpw;
    const int64_t warps = (a.nlines + b.pair_lpw - 1) / b.pair_lpw;
    const unsigned blocks = (unsigned)((warps * 32 + threads - 1) / threads);
    if (integrator == OSB_INT_CK45) {
      if (analytic_inprod) shoot_pair_kernel<OSB_INT_CK45, true><<<blocks, threads, 0, st>>>(b);
      else shoot_pair_kernel<OSB_INT_CK45, false><<<blocks, threads, 0, st>>>(b);
    } else {
      if (analytic_inprod) shoot_pair_kernel<OSB_INT_RK4, true><<<blocks, threads, 0, st>>>(b);
      else shoot_pair_kernel<OSB_INT_RK4, false><<<blocks, threads, 0, st>>>(b);
    }
    if (nlaunch) *nlaunch += 1;
    return cudaGetLastError();
  }
  // small batches: narrow CTAs so the points spread over all 148 SMs
  int threads = (a.nlines >= 148 * 4 * 128) ? 128 : (a.nlines >= 148 * 4 * 64 ? 64 : 32);
  int64_t blocks = (a.nlines + threads - 1) / threads;
  if (blocks > 0x7fffffff) return cudaErrorInvalidValue;
  dim3 g((unsigned)blocks), b(threads);
  if (integrator == OSB_INT_CK45) {
    if (analytic_inprod) shoot_kernel<OSB_INT_CK45, true><<<g, b, 0, st>>>(a);
    else shoot_kernel<OSB_INT_CK45, false><<<g, b, 0, st>>>(a);
  } else {
    if (analytic_inprod) shoot_kernel<OSB_INT_RK4, true><<<g, b, 0, st>>>(a);
    else shoot_kernel<OSB_INT_RK4, false><<<g, b, 0, st>>>(a);
  }
  if (nlaunch) *nlaunch += 1;
  return cudaGetLastError();
}

}  // namespace osb
