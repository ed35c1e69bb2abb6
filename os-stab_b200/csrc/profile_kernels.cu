// profile_kernels.cu -- K1: batched Blasius / Falkner-Skan mean-profile generator
// (replaces SOLVE_BL: contebl.f:386-533, orrsom.f:687-804, orrspace.f:697-817),
// plus the FP64 peak micro-benchmark used as the roofline denominator.
//
// One thread integrates one similarity profile: the march in y is strictly
// sequential (nbl steps x ~6 secant passes on f''(0)), so parallelism is over
// profiles only.  Tables are stored profile-fastest ([array][i][p]) so that a
// warp's 32 profiles read/write 256 contiguous bytes per access.  All
// arithmetic uses the non-contracting __d*_rn intrinsics in the reference's
// operation order: the tables are bit-identical to the CPU restatement (and to
// a no-FMA gfortran build) and everything downstream starts from the same data.
#include "osb_internal.h"

namespace osb {

#define M(a, b) __dmul_rn((a), (b))
#define A(a, b) __dadd_rn((a), (b))
#define S(a, b) __dsub_rn((a), (b))
#define D(a, b) __ddiv_rn((a), (b))

// BLASIUS contebl.f:556-559:  f1'=f2, f2'=f3, f3' = -f1 f3 - beta (1 - f2^2)
__device__ __forceinline__ void blasius_rhs(double beta, const double *xi, double *f) {
  f[0] = xi[1];
  f[1] = xi[2];
  f[2] = S(M(-xi[0], xi[2]), M(beta, S(1.0, M(xi[1], xi[1]))));
}

// SRK4 shoot.f:836-855
__device__ __forceinline__ void srk4(double beta, const double *yo, double *yf, double h) {
  double f[3], k1[3], k2[3], k3[3], k4[3], q[3];
  blasius_rhs(beta, yo, f);
#pragma unroll
  for (int j = 0; j < 3; ++j) { k1[j] = M(h, f[j]); q[j] = A(yo[j], M(0.5, k1[j])); }
  blasius_rhs(beta, q, f);
#pragma unroll
  for (int j = 0; j < 3; ++j) { k2[j] = M(h, f[j]); q[j] = A(yo[j], M(0.5, k2[j])); }
  blasius_rhs(beta, q, f);
#pragma unroll
  for (int j = 0; j < 3; ++j) { k3[j] = M(h, f[j]); q[j] = A(yo[j], k3[j]); }
  blasius_rhs(beta, q, f);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    k4[j] = M(h, f[j]);
    yf[j] = A(A(A(yo[j], D(k1[j], 6.0)), D(A(k2[j], k3[j]), 3.0)), D(k4[j], 6.0));
  }
}

// SRKCK45 shoot.f:1334-1361 (fixed step, error estimate dropped)
__device__ __forceinline__ void srkck45(double beta, const double *yo, double *yf, double h) {
  // b(m,k), m = 2..6 rows, as typed at shoot.f:1307-1314
  const double b[6][5] = {{0, 0, 0, 0, 0},
                          {0.2, 0, 0, 0, 0},
                          {0.075, 0.225, 0, 0, 0},
                          {0.3, -0.9, 1.2, 0, 0},
                          {-0.2037037037037037, 2.5, -2.5925925925925926, 1.2962962962962963, 0},
                          {0.029495804398148147, 0.341796875, 0.041594328703703706,
                           0.40034541377314814, 0.061767578125}};
  const double c[6] = {0.09788359788359788, 0.0, 0.4025764895330113,
                       0.21043771043771045, 0.0, 0.2891022021456804};
  double yk[6][3], yt[3];
#pragma unroll
  for (int m = 0; m < 6; ++m) {
#pragma unroll
    for (int n = 0; n < 3; ++n) yt[n] = yo[n];
#pragma unroll
    for (int k = 0; k < m; ++k)
#pragma unroll
      for (int n = 0; n < 3; ++n) yt[n] = A(yt[n], M(b[m][k], yk[k][n]));
    blasius_rhs(beta, yt, yk[m]);
#pragma unroll
    for (int n = 0; n < 3; ++n) yk[m][n] = M(h, yk[m][n]);
  }
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    double acc = yo[n];
#pragma unroll
    for (int k = 0; k < 6; ++k) acc = A(acc, M(c[k], yk[k][n]));
    yf[n] = acc;
  }
}

// SPLINE shoot.f:906-957 on the strided tables.  X, Y, FDP: element i at [i*stride].
// Bw, Rw: scratch (same stride).  1-based indices as in the Fortran.
__device__ static void spline_strided(int N, const double *X, const double *Y, double *FDP,
                                      double *Bw, double *Rw, size_t stride) {
#define X_(I) X[(size_t)((I)-1) * stride]
#define Y_(I) Y[(size_t)((I)-1) * stride]
#define F_(I) FDP[(size_t)((I)-1) * stride]
#define B_(I) Bw[(size_t)((I)-1) * stride]
#define R_(I) Rw[(size_t)((I)-1) * stride]
#define C_(I) S(X_((I) + 1), X_(I))
  const int NM2 = N - 2, NM1 = N - 1;
  for (int I = 2; I <= NM1; ++I) {
    double Ci = C_(I), Ai = C_(I - 1);
    B_(I) = M(2.0, A(Ai, Ci));
    R_(I) = M(6.0, S(D(S(Y_(I + 1), Y_(I)), Ci), D(S(Y_(I), Y_(I - 1)), Ai)));
  }
  B_(2) = A(B_(2), C_(1));      // ALAMDA = 1
  B_(NM1) = A(B_(NM1), C_(NM1));
  for (int I = 3; I <= NM1; ++I) {
    double T = D(C_(I - 1), B_(I - 1));  // A(I) = C(I-1)
    B_(I) = S(B_(I), M(T, C_(I - 1)));
    R_(I) = S(R_(I), M(T, R_(I - 1)));
  }
  F_(NM1) = D(R_(NM1), B_(NM1));
  for (int I = 2; I <= NM2; ++I) {
    int NMI = N - I;
    F_(NMI) = D(S(R_(NMI), M(C_(NMI), F_(NMI + 1))), B_(NMI));
  }
  F_(1) = F_(2);
  F_(N) = F_(NM1);
#undef X_
#undef Y_
#undef F_
#undef B_
#undef R_
#undef C_
}

// ---------------------------------------------------------------------------
// The spline (SPLINE, shoot.f:906-957) of a batch.  Its tridiagonal system has the SAME matrix for every
// profile and for both splines of a profile -- it depends on the grid only -- so the part of the Thomas
// algorithm that does not involve the data is done once per launch by spline_grid_kernel, with the reference's
// own operations in the reference's order:
//     Cg[I] = X(I+1) - X(I)                                   I = 1 .. N-1
//     Bp[I] = the eliminated diagonal B(I) after shoot.f:934-939  (incl. the ALAMDA = 1 end corrections)
//     Tp[I] = A(I) / B(I-1), the multiplier of row I
// A profile's spline is then two sweeps over its own data: forward R(I) = 6 (dy/C - dy/C) - Tp(I) R(I-1),
// stored in the output array itself, and backward FDP(I) = (R(I) - Cg(I) FDP(I+1)) / Bp(I) in place.
// Bit-identical to the per-profile elimination (same operations on the same numbers); per profile it moves
// 32 KB per spline instead of ~110 KB.  grid: [3][nbl+2] doubles (Cg, Bp, Tp), 1-based index I at [I].
// ---------------------------------------------------------------------------
__global__ void spline_grid_kernel(int nbl, double ymin, double ymax, double *__restrict__ grid) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const int N = nbl + 1, NM1 = N - 1;
  const size_t n2 = (size_t)nbl + 2;
  double *Cg = grid, *Bp = grid + n2, *Tp = grid + 2 * n2;
  const double h = D(S(ymax, ymin), (double)nbl);
#define X_(I) A(ymin, M((double)((I)-1), h))  /* ydat(I-1), contebl.f:499 */
  for (int I = 1; I <= NM1; ++I) Cg[I] = S(X_(I + 1), X_(I));
#undef X_
  for (int I = 2; I <= NM1; ++I) Bp[I] = M(2.0, A(Cg[I - 1], Cg[I]));
  Bp[2] = A(Bp[2], Cg[1]);
  Bp[NM1] = A(Bp[NM1], Cg[NM1]);
  Tp[2] = 0.0;
  for (int I = 3; I <= NM1; ++I) {
    const double T = D(Cg[I - 1], Bp[I - 1]);
    Tp[I] = T;
    Bp[I] = S(Bp[I], M(T, Cg[I - 1]));
  }
}

// Y -> FDP for one profile of the batch (element i of an array at [i * st]); FDP may not alias Y.
// copy (optional): receives FDP as well (contebl.f:505-507: d2U := Uspl).
__device__ __forceinline__ void spline_batched(int N, const double *__restrict__ grid, size_t n2, const double *Y,
                                               double *FDP, double *copy, size_t st) {
  const double *Cg = grid, *Bp = grid + n2, *Tp = grid + 2 * n2;
  const int NM1 = N - 1, NM2 = N - 2;
#define Y_(I) Y[(size_t)((I)-1) * st]
#define F_(I) FDP[(size_t)((I)-1) * st]
  double ym = Y_(1), y0 = Y_(2), rprev = 0.0;
  for (int I = 2; I <= NM1; ++I) {
    const double yp = Y_(I + 1);
    double r = M(6.0, S(D(S(yp, y0), __ldg(Cg + I)), D(S(y0, ym), __ldg(Cg + I - 1))));
    if (I >= 3) r = S(r, M(__ldg(Tp + I), rprev));
    F_(I) = r;
    rprev = r;
    ym = y0; y0 = yp;
  }
  double f = D(F_(NM1), __ldg(Bp + NM1));
  F_(NM1) = f;
  F_(N) = f;
  if (copy) { copy[(size_t)(NM1 - 1) * st] = f; copy[(size_t)(N - 1) * st] = f; }
  for (int I = 2; I <= NM2; ++I) {
    const int NMI = N - I;
    f = D(S(F_(NMI), M(__ldg(Cg + NMI), f)), __ldg(Bp + NMI));
    F_(NMI) = f;
    if (copy) copy[(size_t)(NMI - 1) * st] = f;
  }
  F_(1) = f;  // FDP(1) = FDP(2)
  if (copy) copy[0] = f;
#undef Y_
#undef F_
}

// tables: [5][nbl+2][nprof] (ydat, U, Uspl, d2U, d2Uspl); grid: the tables of spline_grid_kernel
__global__ void blasius_kernel(int variant, int integrator, int nbl, double ymin, double ymax,
                               int64_t nprof, const double *__restrict__ f2p_in,
                               const double *__restrict__ beta_in, double *__restrict__ tables,
                               const double *__restrict__ grid, double *__restrict__ f2p_out,
                               int32_t *__restrict__ npass_out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nprof) return;
  const size_t st = (size_t)nprof;
  const size_t n2 = (size_t)nbl + 2;
  double *ydat = tables + p, *U = tables + n2 * st + p, *Uspl = tables + 2 * n2 * st + p,
         *d2U = tables + 3 * n2 * st + p, *d2Uspl = tables + 4 * n2 * st + p;
  const double beta = beta_in[p];
  const double h = D(S(ymax, ymin), (double)nbl);  // contebl.f:177
  const bool ck45 = (variant == OSB_FLOW_CONTEBL) && (integrator == OSB_INT_CK45);
  // secant iteration on f''(0), contebl.f:422-478.  The reference stores the profile on every pass and keeps the
  // last one; here the passes only carry f'(ymax) and the accepted pass is integrated once more with the stores
  // (same starting value, same operations: same bits) -- the table is written once instead of ~6 times.
  double x30 = f2p_in[p], used = x30, xi3old = 0.0, xi2old = 0.0, err = 1.0;
  int pass = 1;
  while (fabs(err) > 1.e-10) {
    used = x30;
    double xi[3] = {0.0, 0.0, x30}, xn[3];
    for (int i = 1; i <= nbl; ++i) {
      if (ck45) srkck45(beta, xi, xn, h); else srk4(beta, xi, xn, h);
      xi[0] = xn[0]; xi[1] = xn[1]; xi[2] = xn[2];
    }
    const double x2n = xi[1];
    if (pass == 1) {
      xi3old = x30;
      x30 = M(x30, .99);
    } else {
      double tmp = x30;
      x30 = A(x30, M(D(S(xi3old, x30), S(xi2old, x2n)), S(1.0, x2n)));
      xi3old = tmp;
    }
    pass += 1;
    xi2old = x2n;
    err = S(1.0, xi2old);
    if (pass > 200) break;  // the reference would spin forever
  }
  if (f2p_out) f2p_out[p] = used;
  if (npass_out) npass_out[p] = pass - 1;
  {
    double xi[3] = {0.0, 0.0, used}, xn[3];
    U[0] = xi[1];
    ydat[0] = A(ymin, M(0.0, h));
    for (int i = 1; i <= nbl; ++i) {
      if (ck45) srkck45(beta, xi, xn, h); else srk4(beta, xi, xn, h);
      xi[0] = xn[0]; xi[1] = xn[1]; xi[2] = xn[2];
      U[(size_t)i * st] = xi[1];
      ydat[(size_t)i * st] = A(ymin, M((double)i, h));
    }
  }
  spline_batched(nbl + 1, grid, n2, U, Uspl, variant != OSB_FLOW_ORRSPACE ? d2U : nullptr, st);
  if (variant == OSB_FLOW_ORRSPACE) {
#define U_(i) U[(size_t)(i) * st]
    // 2nd-order finite-difference U'' (orrspace.f:784-788); contebl / orrsom overwrite it with Uspl (contebl.f:505-507)
    const double hh = M(h, h);
    for (int i = 1; i <= nbl - 1; ++i)
      d2U[(size_t)i * st] = D(A(S(U_(i + 1), M(2.0, U_(i))), U_(i - 1)), hh);
    d2U[0] = D(A(S(A(-U_(4), M(4.0, U_(3))), M(5.0, U_(2))), M(2.0, U_(1))), hh);
    d2U[(size_t)nbl * st] =
        D(S(A(S(M(2.0, U_(nbl)), M(5.0, U_(nbl - 1))), M(4.0, U_(nbl - 2))), U_(nbl - 3)), hh);
#undef U_
  }
  spline_batched(nbl + 1, grid, n2, d2U, d2Uspl, nullptr, st);
  const size_t last = (size_t)(nbl + 1) * st;  // zero pad (oracle D-1)
  ydat[last] = 0.0; U[last] = 0.0; Uspl[last] = 0.0; d2U[last] = 0.0; d2Uspl[last] = 0.0;
}

// [nprof-fastest] -> [profile][i]: out[a][p][i] = in[a][i][p] for the four tables the batched ABI returns, so that
// the host receives each table with ONE contiguous copy.  32 x 32 tiles through shared memory.
__global__ void transpose_tables_kernel(const double *__restrict__ in, double *__restrict__ out, int n1 /* nbl+1 */,
                                        size_t n2 /* nbl+2 */, int64_t nprof) {
  __shared__ double tile[32][33];
  const int a = blockIdx.z;  // table 0..3 = U, Uspl, d2U, d2Uspl
  const double *src = in + (size_t)(a + 1) * n2 * nprof;
  double *dst = out + (size_t)a * n1 * nprof;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int i0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r;
    const int64_t pp = p0 + threadIdx.x;
    if (i < n1 && pp < nprof) tile[r][threadIdx.x] = src[(size_t)i * nprof + pp];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t pp = p0 + r;
    const int i = i0 + threadIdx.x;
    if (i < n1 && pp < nprof) dst[(size_t)pp * n1 + i] = tile[threadIdx.x][r];
  }
}

cudaError_t launch_transpose_tables(const double *d_tables, double *d_out, int nbl, int64_t nprof, cudaStream_t st) {
  const dim3 grid((unsigned)((nprof + 31) / 32), (unsigned)((nbl + 1 + 31) / 32), 4), block(32, 8);
  transpose_tables_kernel<<<grid, block, 0, st>>>(d_tables, d_out, nbl + 1, (size_t)nbl + 2, nprof);
  return cudaGetLastError();
}

// SHIFT_BL of conteucbl.f:924-968: move the profile to a frame travelling with u_c and
// rebuild both splines:  U = Uorg - uc;  Uspl = SPLINE(U);  d2U := Uspl;  d2Uspl = SPLINE(d2U).
// in/out tables: [5][nbl+2] (ydat, U, Uspl, d2U, d2Uspl); scratch: 2*(nbl+2).  One thread.
__global__ void shift_profile_kernel(int nbl, double uc, const double *__restrict__ in,
                                     double *__restrict__ out, double *__restrict__ scratch) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const size_t n2 = (size_t)nbl + 2;
  double *ydat = out, *U = out + n2, *Uspl = out + 2 * n2, *d2U = out + 3 * n2, *d2Uspl = out + 4 * n2;
  for (int i = 0; i <= nbl; ++i) { ydat[i] = in[i]; U[i] = S(in[n2 + i], uc); }
  spline_strided(nbl + 1, ydat, U, Uspl, scratch, scratch + n2, 1);
  for (int i = 0; i <= nbl; ++i) d2U[i] = Uspl[i];
  spline_strided(nbl + 1, ydat, d2U, d2Uspl, scratch, scratch + n2, 1);
  ydat[nbl + 1] = 0.0; U[nbl + 1] = 0.0; Uspl[nbl + 1] = 0.0; d2U[nbl + 1] = 0.0; d2Uspl[nbl + 1] = 0.0;
}

cudaError_t launch_shift_profile(int nbl, double uc, const double *d_in, double *d_out, double *d_scratch,
                                 cudaStream_t st) {
  shift_profile_kernel<<<1, 32, 0, st>>>(nbl, uc, d_in, d_out, d_scratch);
  return cudaGetLastError();
}

// the grid tables of spline_grid_kernel (independent of the number of profiles)
size_t blasius_scratch_doubles(int nbl, int64_t nprof) {
  (void)nprof;
  return 3 * ((size_t)nbl + 2);
}

cudaError_t launch_blasius(int variant, int integrator, int nbl, double ymin, double ymax,
                           int64_t nprof, const double *d_f2p, const double *d_beta,
                           double *d_tables, double *d_scratch, double *d_f2p_out,
                           int32_t *d_npass, cudaStream_t st) {
  int threads = nprof >= 148 * 64 ? 64 : 32;
  int64_t blocks = (nprof + threads - 1) / threads;
  spline_grid_kernel<<<1, 32, 0, st>>>(nbl, ymin, ymax, d_scratch);
  blasius_kernel<<<(unsigned)blocks, threads, 0, st>>>(variant, integrator, nbl, ymin, ymax, nprof,
                                                       d_f2p, d_beta, d_tables, d_scratch,
                                                       d_f2p_out, d_npass);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// FP64 peak: 8 independent DFMA chains per thread, no memory traffic.
// ---------------------------------------------------------------------------
__global__ void fp64_peak_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

cudaError_t launch_fp64_peak(double *d_out, int iters, int blocks, int threads, cudaStream_t st) {
  fp64_peak_kernel<<<blocks, threads, 0, st>>>(d_out, iters);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// FP64 TENSOR peak: 8 independent chains of mma.sync.m8n8k4.f64 (DMMA) per warp, no memory traffic.
// One instruction = 8 x 8 x 4 multiply-adds = 512 flops per warp.  This is the roofline of the dense LU's
// trailing update (chan_kernels.cu); tcgen05 has no fp64 kind.
// ---------------------------------------------------------------------------
__global__ void dmma_peak_kernel(double *out, int iters) {
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x * 1e-9 + j; c[j][1] = 0.5 * j; }
  const double a = 0.999999999 + 1e-12 * threadIdx.x, b = 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[j][0]), "+d"(c[j][1])
                     : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  if (s == 123.456) out[0] = s;
}

cudaError_t launch_dmma_peak(double *d_out, int iters, int blocks, int threads, cudaStream_t st) {
  dmma_peak_kernel<<<blocks, threads, 0, st>>>(d_out, iters);
  return cudaGetLastError();
}

}  // namespace osb
