// chan_kernels.cu -- K3/K4: the channel MATRIX path (orrfdchan / orrcolchan / orrncbl).
//
// Replaces MAKE_MATRIX + SOLVE_ORR_SOM (orrfdchan.f:553-659,688-823; orrcolchan.f:272-513;
// orrncbl.f:553-695,758-981) for a batch of independent (alpha, Re) points.  The
// reference forms inv(B) A and calls ZGEEV for the FULL spectrum (25 n^3 flops) only
// to keep one eigenvalue; here every point is a shift-invert inverse iteration
//     (A - sigma B) y = B x,   omega = sigma + (y^H x)/(y^H y)
// on the pencil itself:
//   K3 (fused): M = A - sigma B is assembled straight into the LU workspace from the
//       shared real tables D2, D4 (A3 = A1 = B1 = 0 for the channel metrics); A and B
//       are never materialised -- B x and the d/dalpha products are applied from D2.
//   K4: one CTA per instance: blocked right-looking complex LU with partial pivoting
//       (IZAMAX's |re|+|im| rule, like ZGETRF), the 16-column panel factorised in
//       shared memory, the trailing update A22 -= L21 U12 on the FP64 tensor path
//       (mma.sync.m8n8k4.f64 -> DMMA; complex product as four real ones), then
//       triangular solves + the D2 mat-vec per inverse iteration.  tcgen05 has no
//       fp64 kind, so DMMA via mma.sync is the sm_100a FP64 tensor path.
//   Newton on alpha (orrncbl) runs inside the kernel, one CTA per Re, including the
//       reference's use of the LEFT eigenvector of inv(B)A (oracle D-7), obtained as
//       vl = B^H u with u from inverse iteration on (A - sigma B)^H.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "osb_cplx.cuh"
#include "osb_internal.h"

namespace osb {

constexpr int CH_THREADS = 256;      // CTAs of the 2-per-SM variant and of the Newton kernel
constexpr int CH_THREADS_BIG = 512;  // the 1-per-SM variant (n > ~350): more warps to hide L2/HBM latency
constexpr int CH_THREADS_MAX = 512;
constexpr int CH_THREADS_SMALL = 128;
constexpr int CH_NB = 16;

// optional phase timers (cycles of block 0, thread 0); read back when OSB_CHAN_PROFILE is set
__device__ long long g_chan_prof[16];
#define PROF_T0() long long prof_t0 = clock64()
#define PROF_ADD(slot)                                                        \
  do {                                                                        \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                \
      long long t1 = clock64();                                               \
      g_chan_prof[slot] += t1 - prof_t0;                                      \
      prof_t0 = t1;                                                           \
    } else {                                                                  \
      prof_t0 = 0;                                                            \
    }                                                                         \
  } while (0)

__device__ __forceinline__ double2 cmul2(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc2(double2 a, double2 b) {  // conj(a) * b
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double2 csub2(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// a - l*b with four DFMAs (instead of DMUL + DFMA + DADD per component)
__device__ __forceinline__ double2 cfnma2(double2 a, double2 l, double2 b) {
  return make_double2(fma(-l.x, b.x, fma(l.y, b.y, a.x)), fma(-l.x, b.y, fma(-l.y, b.x, a.y)));
}
__device__ __forceinline__ double2 cadd2(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 cdiv2(double2 a, double2 b) {
  zd r = zdiv(zd{a.x, a.y}, zd{b.x, b.y});
  return make_double2(r.x, r.y);
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// block-wide sum of a complex value; result valid in every thread.  red: >= 2*(CH_THREADS/32) doubles
__device__ double2 block_sum(double2 v, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) { red[2 * w] = v.x; red[2 * w + 1] = v.y; }
  __syncthreads();
  double2 s = make_double2(0.0, 0.0);
  for (int i = 0; i < nw; ++i) { s.x += red[2 * i]; s.y += red[2 * i + 1]; }
  return s;
}

// ---------------------------------------------------------------------------
// per-instance coefficients of the pencil (MAKE_MATRIX with m1=1, m2=m3=m4=0)
// ---------------------------------------------------------------------------
struct PencilCoef {
  double2 alpha, a2, a3, a4;
  double Re;
  double2 b2, b0;  // B = b2 D2 + b0 I          (orrfdchan.f:633-652)
  double2 db0;     // dB/dalpha = db0 I          (orrncbl.f:686-691)
};

__device__ __forceinline__ PencilCoef pencil_coef(double2 alpha, double Re) {
  PencilCoef p;
  p.alpha = alpha; p.Re = Re;
  p.a2 = cmul2(alpha, alpha); p.a3 = cmul2(p.a2, alpha); p.a4 = cmul2(p.a2, p.a2);
  p.b2 = make_double2(0.0, -Re);                                  // -1.0*(0,1)*Re
  p.b0 = make_double2(-p.a2.y * Re, p.a2.x * Re);                 // (0,1)*alpha^2*Re
  p.db0 = make_double2(-2.0 * alpha.y * Re, 2.0 * alpha.x * Re);  // (0,1)*2*alpha*Re
  return p;
}
// row coefficients: A = D4 + w2_i D2 + w0_i I ; dA = d2_i D2 + d0_i I
__device__ __forceinline__ double2 coef_w2(const PencilCoef &p, double u) {
  // -2 alpha^2 - (0,1) alpha Re U
  return make_double2(-2.0 * p.a2.x + p.alpha.y * p.Re * u, -2.0 * p.a2.y - p.alpha.x * p.Re * u);
}
__device__ __forceinline__ double2 coef_w0(const PencilCoef &p, double u, double d2u) {
  // alpha^4 + (0,1) alpha^3 Re U + (0,1) alpha Re U''
  return make_double2(p.a4.x - p.a3.y * p.Re * u - p.alpha.y * p.Re * d2u,
                      p.a4.y + p.a3.x * p.Re * u + p.alpha.x * p.Re * d2u);
}
__device__ __forceinline__ double2 coef_d2(const PencilCoef &p, double u) {
  // -4 alpha - (0,1) Re U
  return make_double2(-4.0 * p.alpha.x, -4.0 * p.alpha.y - p.Re * u);
}
__device__ __forceinline__ double2 coef_d0(const PencilCoef &p, double u, double d2u) {
  // 4 alpha^3 + 3 (0,1) alpha^2 Re U + (0,1) Re U''
  return make_double2(4.0 * p.a3.x - 3.0 * p.a2.y * p.Re * u, 4.0 * p.a3.y + 3.0 * p.a2.x * p.Re * u + p.Re * d2u);
}

// K3 (fused): M = A - sigma B, column-major n x n, ld = n
__device__ void assemble_shifted(double2 *__restrict__ M, int n, const double *__restrict__ D2,
                                 const double *__restrict__ D4, const double *__restrict__ u,
                                 const double *__restrict__ d2u, const PencilCoef &p, double2 sigma) {
  const double2 sb2 = cmul2(sigma, p.b2), sb0 = cmul2(sigma, p.b0);
  // a thread keeps its row(s) i and walks the columns: the row coefficient is computed once
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double2 w2 = csub2(coef_w2(p, u[i]), sb2);
    const double2 w0 = csub2(coef_w0(p, u[i], d2u[i]), sb0);
    // eight columns per trip, the sixteen table loads issued before the first store: the loop waits for the L2 once
    // per trip (the tables are shared by every CTA and stay L2-resident, the matrix is written through to DRAM)
    for (int j0 = 0; j0 < n; j0 += 8) {
      double d2[8], d4[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const size_t e = (size_t)min(j0 + q, n - 1) * n + i;
        d2[q] = D2[e]; d4[q] = D4[e];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int j = j0 + q;
        if (j < n) {
          double2 v = make_double2(fma(w2.x, d2[q], d4[q]), w2.y * d2[q]);
          if (i == j) v = cadd2(v, w0);
          M[(size_t)j * n + i] = v;
        }
      }
    }
  }
  __syncthreads();
}

// The composed row permutation of a factorisation lives right behind its pivot list: perm[k] = the original row that
// ends at position k.  The solves gather / scatter through it in parallel instead of replaying the n interchanges on
// one thread (n dependent shared-memory round trips per solve).
__host__ __device__ inline int lu_perm_offset(int n) { return (n + 8 + 3) & ~3; }
__device__ __forceinline__ int *lu_perm(const int *ipiv, int n) { return const_cast<int *>(ipiv) + lu_perm_offset(n); }

constexpr int LU_RPT = 4;  // panel rows per thread: n <= LU_RPT * blockDim.x (every launcher satisfies it)

// The composed permutation follows the rows of a factorised panel to their positions (all reads before the barrier, all
// writes after it).  Out of line: its four values held across the barrier would otherwise weigh on the register
// allocation of the LU around it.
__device__ __noinline__ void lu_perm_follow(int *perm, int mp, int p0, int p1, int p2, int p3) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int r0 = tid, r1 = tid + nt, r2 = tid + 2 * nt, r3 = tid + 3 * nt;
  const int v0 = r0 < mp ? perm[r0] : 0, v1 = r1 < mp ? perm[r1] : 0, v2 = r2 < mp ? perm[r2] : 0, v3 = r3 < mp ? perm[r3] : 0;
  __syncthreads();
  if (r0 < mp) perm[p0] = v0;
  if (r1 < mp) perm[p1] = v1;
  if (r2 < mp) perm[p2] = v2;
  if (r3 < mp) perm[p3] = v3;
}
// x <- P x (forward) or x <- P^T x through the composed permutation; ends with a barrier
__device__ __noinline__ void cta_permute(double2 *x, const int *perm, int n, bool forward) {
  const int tid = threadIdx.x, nt = blockDim.x;
  double2 xv[LU_RPT];
#pragma unroll
  for (int q = 0; q < LU_RPT; ++q) { const int r = tid + q * nt; if (r < n) xv[q] = forward ? x[perm[r]] : x[r]; }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < LU_RPT; ++q) { const int r = tid + q * nt; if (r < n) x[forward ? r : perm[r]] = xv[q]; }
  __syncthreads();
}

// The nb columns of a staged panel, ONE barrier each.  Rows stay where they are in shared memory while the panel is factorised
// (a thread owns the physical rows tid, tid + nt, ... and keeps, in registers, the POSITION each would have after
// LAPACK's interchanges): the pivot row is only read once chosen, every other row is updated by its owner alone,
// so the one thing the threads exchange per column is the pivot search.  Same pivots (ties go to the smaller
// position, as IZAMAX), same arithmetic per row, same ipiv as the swapping form -- 4 barriers per column there.
__device__ __forceinline__ void lu_panel_columns(double2 *panel, int ldp, int mp, int nb, int k0, int *ipiv, int *perm,
                                                 double *red, int (&pos)[LU_RPT]) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  int *red_i = (int *)(red + 2 * (CH_THREADS_MAX / 32));
  unsigned alive = 0;
#pragma unroll
  for (int q = 0; q < LU_RPT; ++q) {
    pos[q] = tid + q * nt;
    if (pos[q] < mp) alive |= 1u << q;
  }
  for (int j = 0; j < nb; ++j) {
    // pivot search in INTEGER arithmetic (the bit pattern of a non-negative double orders like the double, a NaN
    // above everything): the FP64 pipe is shared with the other CTA of the SM, which keeps it busy with DMMAs, so
    // every FP64 instruction taken off this serial chain shortens the column
    long long best = -1;
    int key = 0x7fffffff;  // position << 16 | physical row
#pragma unroll
    for (int q = 0; q < LU_RPT; ++q)
      if (alive >> q & 1) {
        const int r = tid + q * nt;
        const double2 v = panel[j * ldp + r];
        const long long a = __double_as_longlong(fabs(v.x) + fabs(v.y)) & 0x7fffffffffffffffll;
        const int k = pos[q] << 16 | r;
        if (a > best || (a == best && k < key)) { best = a; key = k; }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const long long ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ok = __shfl_xor_sync(0xffffffffu, key, o);
      if (ob > best || (ob == best && ok < key)) { best = ob; key = ok; }
    }
    long long *redj = (long long *)red + (j & 1) * nwarp;  // double-buffered: column j + 1 writes the other half
    int *redij = red_i + (j & 1) * nwarp;
    if (lane == 0) { redj[warp] = best; redij[warp] = key; }
    __syncthreads();
    {
      long long b[CH_THREADS_MAX / 32];
      int k[CH_THREADS_MAX / 32];
#pragma unroll
      for (int w = 0; w < CH_THREADS_MAX / 32; ++w) {
        b[w] = w < nwarp ? redj[w] : -1;
        k[w] = w < nwarp ? redij[w] : 0x7fffffff;
      }
#pragma unroll
      for (int o = CH_THREADS_MAX / 64; o > 0; o >>= 1)
#pragma unroll
        for (int w = 0; w < o; ++w)
          if (b[w + o] > b[w] || (b[w + o] == b[w] && k[w + o] < k[w])) { b[w] = b[w + o]; k[w] = k[w + o]; }
      key = k[0];
    }
    const int P = key & 0xffff, posP = key >> 16;
    if (tid == 0) ipiv[k0 + j] = k0 + posP;
    // 1 / pivot with one division (the entries of these matrices are far from the overflow range of |pivot|^2)
    double2 piv = panel[j * ldp + P];
    if (piv.x == 0.0 && piv.y == 0.0) piv = make_double2(1.0e-290, 0.0);
    const double pinv = 1.0 / fma(piv.x, piv.x, piv.y * piv.y);
    const double2 rinv = make_double2(piv.x * pinv, -piv.y * pinv);
#pragma unroll
    for (int q = 0; q < LU_RPT; ++q)
      if (alive >> q & 1) {
        const int r = tid + q * nt;
        if (r == P) { alive &= ~(1u << q); pos[q] = j; continue; }
        if (pos[q] == j) pos[q] = posP;
        const double2 l = cmul2(panel[j * ldp + r], rinv);
        panel[j * ldp + r] = l;
        for (int c0 = j + 1; c0 < nb; c0 += 4) {
          double2 uu[4], m[4];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const int c = min(c0 + qq, nb - 1);
            uu[qq] = panel[c * ldp + P];
            m[qq] = panel[c * ldp + r];
          }
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            if (c0 + qq < nb) panel[(c0 + qq) * ldp + r] = cfnma2(m[qq], l, uu[qq]);
        }
      }
  }
  lu_perm_follow(perm + k0, mp, pos[0], pos[1], pos[2], pos[3]);  // holds the barrier the callers rely on
}

// The factorised panel goes back to the matrix with every row at its position ...
__device__ __forceinline__ void lu_panel_store(double2 *__restrict__ M, int n, int k0, int nb, int mp, const double2 *panel,
                                               int ldp, const int (&pos)[LU_RPT]) {
  const int tid = threadIdx.x, nt = blockDim.x;
#pragma unroll
  for (int q = 0; q < LU_RPT; ++q) {
    const int r = tid + q * nt;
    if (r < mp)
      for (int c = 0; c < nb; ++c) M[(size_t)(k0 + c) * n + k0 + pos[q]] = panel[c * ldp + r];
  }
}
// ... and, after a barrier, the (at most 2 nb) rows whose position differs from where they lie are put in place in
// shared memory too, from the copy just written (U12 and the trailing update read the panel by position)
__device__ __forceinline__ void lu_panel_settle(const double2 *__restrict__ M, int n, int k0, int nb, int mp, double2 *panel,
                                                int ldp, const int (&pos)[LU_RPT]) {
  const int tid = threadIdx.x, nt = blockDim.x;
#pragma unroll
  for (int q = 0; q < LU_RPT; ++q) {
    const int r = tid + q * nt;
    if (r < mp && pos[q] != r)
      for (int c = 0; c < nb; ++c) panel[c * ldp + pos[q]] = M[(size_t)(k0 + c) * n + k0 + pos[q]];
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// K4: blocked LU, P M = L U in place.  panel: shared, (n + 16) x CH_NB complex.
// ---------------------------------------------------------------------------
__device__ void cta_lu(double2 *__restrict__ M, int n, int *ipiv, double2 *panel, double *red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  int *perm = lu_perm(ipiv, n);
  for (int i = tid; i < n; i += nt) perm[i] = i;  // (visible to thread 0 after the barrier that follows the first staging)
  const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const int ldp = n + 16;
  for (int k0 = 0; k0 < n; k0 += CH_NB) {
    const int nb = min(CH_NB, n - k0), mp = n - k0;
    PROF_T0();
    // ---- stage the panel
    for (int e = tid; e < mp * nb; e += nt) {
      const int r = e % mp, c = e / mp;
      panel[c * ldp + r] = M[(size_t)(k0 + c) * n + k0 + r];
    }
    __syncthreads();
    // ---- unblocked factorisation inside the panel (ZGETF2's pivots and arithmetic, rows left in place)
    int pos[LU_RPT];
    lu_panel_columns(panel, ldp, mp, nb, k0, ipiv, perm, red, pos);
    PROF_ADD(1);
    // ---- row interchanges on the columns outside the panel; write the panel back
    // two columns per thread, interleaved: two independent chains of dependent swaps in flight
    for (int c = tid; c < n; c += 2 * nt) {
      const int cb = c + nt;
      const bool doa = !(c >= k0 && c < k0 + nb), dob = cb < n && !(cb >= k0 && cb < k0 + nb);
      double2 *cola = M + (size_t)c * n, *colb = M + (size_t)(dob ? cb : c) * n;
      for (int j = 0; j < nb; ++j) {
        const int p = ipiv[k0 + j];
        if (p != k0 + j) {
          double2 ta, pa, tb, pb;
          if (doa) { ta = cola[k0 + j]; pa = cola[p]; }
          if (dob) { tb = colb[k0 + j]; pb = colb[p]; }
          if (doa) { cola[k0 + j] = pa; cola[p] = ta; }
          if (dob) { colb[k0 + j] = pb; colb[p] = tb; }
        }
      }
    }
    lu_panel_store(M, n, k0, nb, mp, panel, ldp, pos);
    __syncthreads();
    lu_panel_settle(M, n, k0, nb, mp, panel, ldp, pos);
    PROF_ADD(2);
    const int mt = n - k0 - nb;  // trailing size
    if (mt <= 0) break;
    // ---- U12 = L11^{-1} A12  (unit lower L11 from the panel), one thread per column
    for (int c = tid; c < mt; c += nt) {
      double2 *col = M + (size_t)(k0 + nb + c) * n + k0;
      double2 a[CH_NB];
#pragma unroll
      for (int i = 0; i < CH_NB; ++i) a[i] = (i < nb) ? col[i] : make_double2(0.0, 0.0);
#pragma unroll
      for (int i = 1; i < CH_NB; ++i) {
        if (i < nb) {
          double2 s = a[i];
#pragma unroll
          for (int t = 0; t < i; ++t) s = cfnma2(s, panel[t * ldp + i], a[t]);
          a[i] = s;
        }
      }
#pragma unroll
      for (int i = 0; i < CH_NB; ++i) if (i < nb) col[i] = a[i];
    }
    __syncthreads();
    PROF_ADD(3);
    // ---- trailing update A22 -= L21 U12 on the FP64 tensor path; a warp owns 16x16 blocks
    if (nb == CH_NB) {
      const int g = lane >> 2, t4 = lane & 3;
      const int nblk = (mt + 15) >> 4;
      for (int blk = warp; blk < nblk * nblk; blk += nwarp) {
        const int R0 = (blk % nblk) << 4, C0 = (blk / nblk) << 4;
        double cr[2][2][2], ci[2][2][2];
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
          for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = R0 + 8 * rt + g, j = C0 + 8 * ct + 2 * t4 + e;
              double2 v = make_double2(0.0, 0.0);
              if (i < mt && j < mt) v = M[(size_t)(k0 + nb + j) * n + k0 + nb + i];
              cr[rt][ct][e] = v.x; ci[rt][ct][e] = v.y;
            }
#pragma unroll
        for (int ks = 0; ks < CH_NB / 4; ++ks) {
          const int k = 4 * ks + t4;
          double lr[2], li[2], ur[2], ui[2];
#pragma unroll
          for (int rt = 0; rt < 2; ++rt) {  // rows past the end read the zero/garbage pad; their outputs are dropped
            const double2 l = panel[k * ldp + nb + R0 + 8 * rt + g];
            lr[rt] = l.x; li[rt] = l.y;
          }
#pragma unroll
          for (int ct = 0; ct < 2; ++ct) {
            const int j = min(C0 + 8 * ct + g, mt - 1);
            const double2 uu = M[(size_t)(k0 + nb + j) * n + k0 + k];
            ur[ct] = uu.x; ui[ct] = uu.y;
          }
#pragma unroll
          for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
              // C -= L U :  Cr += (-Lr) Ur + Li Ui ;  Ci += (-Lr) Ui + (-Li) Ur.  The two products of an accumulator are issued
              // eight instructions apart (first halves of all eight accumulators, then the second halves): a dependent DMMA
              // back to back would stall the warp for the pipe's latency.  Same order per accumulator, same bits.
              dmma884(cr[rt][ct][0], cr[rt][ct][1], -lr[rt], ur[ct]);
              dmma884(ci[rt][ct][0], ci[rt][ct][1], -lr[rt], ui[ct]);
            }
#pragma unroll
          for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
              dmma884(cr[rt][ct][0], cr[rt][ct][1], li[rt], ui[ct]);
              dmma884(ci[rt][ct][0], ci[rt][ct][1], -li[rt], ur[ct]);
            }
        }
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
          for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = R0 + 8 * rt + g, j = C0 + 8 * ct + 2 * t4 + e;
              if (i < mt && j < mt)
                M[(size_t)(k0 + nb + j) * n + k0 + nb + i] = make_double2(cr[rt][ct][e], ci[rt][ct][e]);
            }
      }
    }
    __syncthreads();
    PROF_ADD(4);
  }
}

// 16-byte / 4-byte asynchronous global -> shared copies; !valid zero-fills the destination
__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


// ---------------------------------------------------------------------------
// K4 for large n (the one-CTA-per-SM kernel, n > ~350): TWO-LEVEL blocked LU.
// With a single 16-column blocking the right-looking update sweeps the whole trailing matrix n/16 times; at
// n = 511 that is 253 MB of DRAM traffic for a 4.2 MB matrix (148 concurrent matrices do not fit the L2) and
// the kernel is bound by it (ncu round 1: long_scoreboard 49 %, DMMA pipe 15 %).  Here the matrix is processed
// in OUTER panels of 64 columns:
//   * inside an outer panel the 16-column panels are factorised as before (shared-memory panel, pivot search,
//     rank-1 updates), but their row interchanges, U12 solves and DMMA updates touch only the other columns
//     of the outer panel (<= 48 columns, L2-resident);
//   * once per outer panel: the 64 interchanges are applied to the rest of the matrix, U12 = L11^{-1} A12 is
//     solved with the 64 x 64 unit-lower L11 staged in shared memory, and the trailing matrix is updated with
//     k = 64: A22 -= L21 U12 in 64 x 64 super-tiles whose operands are staged in shared memory once per CTA
//     (the 64 x 64 L21 block and the 64 x 64 U12 strip, reused by all warps) and read from there by the
//     mma.sync.m8n8k4.f64 (DMMA) tiles.  The trailing matrix is read and written n/64 times: 4x less traffic.
// Same pivoting rule and the same elimination order inside a column as cta_lu: identical pivots; the updates
// of a trailing entry are accumulated in the same order k = 0, 1, ... (results agree to rounding).
// ---------------------------------------------------------------------------
// Two shapes: <16, 64> for the 512-thread one-CTA-per-SM kernel, and <8, 32> whose panel (8 columns) and
// staging area (32-deep) leave room for TWO 256-thread CTAs per SM, so that one CTA's latency-bound phases
// (pivot search, substitutions) overlap the other's DMMA update.
constexpr int CH_ST = 64;  // columns of a super-tile of the big update
// Super-tile rows SR and staging of the L21 blocks, per shape:
//   <16, 64> (512 threads): SR = 64, 16 warp tiles, one L buffer (a second one does not fit next to the U strip)
//   <8, 32>  (256 threads): SR = 32,  8 warp tiles, TWO L buffers: the next block arrives by cp.async while this one is used
__host__ __device__ constexpr int ch_sr(int nbo) { return nbo == 64 ? 64 : 32; }
__host__ __device__ constexpr bool ch_db(int nbo) { return nbo != 64; }
__host__ __device__ constexpr int ch_ldl(int nbo) { return ch_sr(nbo) + 2; }  // staged L21 block: [k][row], padded against bank conflicts
__host__ __device__ constexpr int ch_ldu(int nbo) { return nbo + 4; }         // staged U12 strip: [col][k]
__host__ __device__ constexpr size_t chan_stage_elems(int nbo) {
  return (size_t)(ch_db(nbo) ? 2 : 1) * nbo * ch_ldl(nbo) + (size_t)CH_ST * ch_ldu(nbo);
}

// A22 -= L21 U12 with k = nbo (<= NBO): SR x 64 super-tiles, operands staged in shared memory, a warp per 16 x 16 tile
template <int NBO>
__device__ __forceinline__ void lu2_big_update(double2 *__restrict__ M, int n, int K0, int KE, int nbo, int MT, double2 *stage) {
  constexpr int SR = ch_sr(NBO), LDL = ch_ldl(NBO), LDU = ch_ldu(NBO), LBUF = NBO * LDL;
  constexpr bool DB = ch_db(NBO);
  constexpr int TR = SR / 16;  // tile rows per super-tile; 4 tile columns
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  double2 *sU = stage + (DB ? 2 : 1) * LBUF;
  const int ncs = (MT + CH_ST - 1) / CH_ST, nrs = (MT + SR - 1) / SR;
  auto stage_L = [&](int I, double2 *dst) {
    for (int e = tid; e < NBO * SR; e += nt) {
      const int r = e % SR, k = e / SR;
      const bool ok = k < nbo && I * SR + r < MT;
      const double2 *src = M + (size_t)(K0 + (ok ? k : 0)) * n + KE + (ok ? I * SR + r : 0);
      if (DB) cp_async16(dst + k * LDL + r, src, ok);
      else dst[k * LDL + r] = ok ? *src : make_double2(0.0, 0.0);
    }
  };
  for (int J = 0; J < ncs; ++J) {
    const int C0 = J * CH_ST;
    __syncthreads();  // every warp is done with the previous strip's buffers
    for (int e = tid; e < NBO * CH_ST; e += nt) {  // U12 strip: sU[col][k]
      const int k = e % NBO, c = e / NBO;
      double2 v = make_double2(0.0, 0.0);
      if (k < nbo && C0 + c < MT) v = M[(size_t)(KE + C0 + c) * n + K0 + k];
      sU[c * LDU + k] = v;
    }
    if (DB) { stage_L(0, stage); cp_async_commit(); }
    for (int I = 0; I < nrs; ++I) {
      const int R0 = I * SR;
      double2 *sL = stage + (DB ? (I & 1) * LBUF : 0);
      // one tile per warp (TR * 4 == warps per CTA for both shapes).  The C block of the NEXT row block is requested
      // into the L2 now (one 128-byte line per lane: 16 columns x 2 halves), a whole super-tile ahead of its use;
      // holding it in registers across the barrier instead spills at the 128-register budget (ncu: the spill
      // stores waiting for the loads were 11 % of all stall samples).
      const int tile0 = warp;
      const int tr0 = (tile0 % TR) << 4, tc0 = (tile0 / TR) << 4;
      const bool have0 = tile0 < TR * 4 && R0 + tr0 < MT && C0 + tc0 < MT;
      if (tile0 < TR * 4 && I + 1 < nrs) {
        const int i = R0 + SR + tr0 + 8 * (lane & 1), j = C0 + tc0 + (lane >> 1);
        if (i < MT && j < MT) prefetch_l2(M + (size_t)(KE + j) * n + KE + i);
      }
      double cr[2][2][2], ci[2][2][2];
      if (DB) {
        cp_async_wait<0>();
        __syncthreads();  // block I has landed for everybody; everybody is done with block I-1
        if (I + 1 < nrs) { stage_L(I + 1, stage + ((I + 1) & 1) * LBUF); cp_async_commit(); }
      } else {
        __syncthreads();
        stage_L(I, sL);
        __syncthreads();
      }
      if (have0) {
        const int tr = tr0, tc = tc0;
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
          for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = R0 + tr + 8 * rt + g, j = C0 + tc + 8 * ct + 2 * t4 + e;
              double2 v = make_double2(0.0, 0.0);
              if (i < MT && j < MT) v = M[(size_t)(KE + j) * n + KE + i];
              cr[rt][ct][e] = v.x; ci[rt][ct][e] = v.y;
            }
#pragma unroll 4
        for (int ks = 0; ks < NBO / 4; ++ks) {
          const int k = 4 * ks + t4;
          double lr[2], li[2], ur[2], ui[2];
#pragma unroll
          for (int rt = 0; rt < 2; ++rt) {
            const double2 l = sL[k * LDL + tr + 8 * rt + g];
            lr[rt] = l.x; li[rt] = l.y;
          }
#pragma unroll
          for (int ct = 0; ct < 2; ++ct) {
            const double2 uu = sU[(tc + 8 * ct + g) * LDU + k];
            ur[ct] = uu.x; ui[ct] = uu.y;
          }
#pragma unroll
          for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
              // C -= L U :  Cr += (-Lr) Ur + Li Ui ;  Ci += (-Lr) Ui + (-Li) Ur.  The two products of an accumulator are issued
              // eight instructions apart (first halves of all eight accumulators, then the second halves): a dependent DMMA
              // back to back would stall the warp for the pipe's latency.  Same order per accumulator, same bits.
              dmma884(cr[rt][ct][0], cr[rt][ct][1], -lr[rt], ur[ct]);
              dmma884(ci[rt][ct][0], ci[rt][ct][1], -lr[rt], ui[ct]);
            }
#pragma unroll
          for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
              dmma884(cr[rt][ct][0], cr[rt][ct][1], li[rt], ui[ct]);
              dmma884(ci[rt][ct][0], ci[rt][ct][1], -li[rt], ur[ct]);
            }
        }
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
          for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = R0 + tr + 8 * rt + g, j = C0 + tc + 8 * ct + 2 * t4 + e;
              if (i < MT && j < MT) M[(size_t)(KE + j) * n + KE + i] = make_double2(cr[rt][ct][e], ci[rt][ct][e]);
            }
      }
    }
  }
}

template <int NBI, int NBO>
__device__ __noinline__ void cta_lu2(double2 *__restrict__ M, int n, int *ipiv, double2 *panel, double *red) {
  constexpr int CH_NBO = NBO;
  const int tid = threadIdx.x, nt = blockDim.x;
  int *perm = lu_perm(ipiv, n);
  for (int i = tid; i < n; i += nt) perm[i] = i;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const int ldp = n + 16;
  const int g = lane >> 2, t4 = lane & 3;
  for (int K0 = 0; K0 < n; K0 += CH_NBO) {
    const int nbo = min(CH_NBO, n - K0);
    const int KE = K0 + nbo;  // end of the outer panel
    // ================= factorise the outer panel, 16 columns at a time =================
    for (int k0 = K0; k0 < KE; k0 += NBI) {
      const int nb = min(NBI, KE - k0), mp = n - k0;
      PROF_T0();
      for (int e = tid; e < mp * nb; e += nt) {
        const int r = e % mp, c = e / mp;
        panel[c * ldp + r] = M[(size_t)(k0 + c) * n + k0 + r];
      }
      __syncthreads();
      PROF_ADD(7);
      int pos[LU_RPT];
      lu_panel_columns(panel, ldp, mp, nb, k0, ipiv, perm, red, pos);
      PROF_ADD(1);
      // ---- write the panel back (rows by position); the interchanges on the OTHER columns of the outer panel
      lu_panel_store(M, n, k0, nb, mp, panel, ldp, pos);
      for (int c = K0 + tid; c < KE; c += nt) {
        if (c >= k0 && c < k0 + nb) continue;
        double2 *col = M + (size_t)c * n;
        for (int j = 0; j < nb; ++j) {
          const int p = ipiv[k0 + j];
          if (p != k0 + j) { const double2 t = col[k0 + j]; col[k0 + j] = col[p]; col[p] = t; }
        }
      }
      __syncthreads();
      lu_panel_settle(M, n, k0, nb, mp, panel, ldp, pos);
      PROF_ADD(2);
      const int nc = KE - k0 - nb;     // columns of the outer panel to the right of this panel
      const int mt = n - k0 - nb;      // rows below it
      if (nc <= 0) break;
      // ---- U12 = L11^{-1} A12 for those columns, one thread per column
      for (int c = tid; c < nc; c += nt) {
        double2 *col = M + (size_t)(k0 + nb + c) * n + k0;
        double2 a[NBI];
#pragma unroll
        for (int i = 0; i < NBI; ++i) a[i] = (i < nb) ? col[i] : make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 1; i < NBI; ++i) {
          if (i < nb) {
            double2 sacc = a[i];
#pragma unroll
            for (int t = 0; t < i; ++t) sacc = cfnma2(sacc, panel[t * ldp + i], a[t]);
            a[i] = sacc;
          }
        }
#pragma unroll
        for (int i = 0; i < NBI; ++i) if (i < nb) col[i] = a[i];
      }
      __syncthreads();
      PROF_ADD(3);
      // ---- update of the outer panel's remaining columns (k = 16, DMMA); a warp owns 16x16 blocks
      if (nb == NBI && mt > 0) {
        const int nbr = (mt + 15) >> 4, nbc = (nc + 15) >> 4;
        for (int blk = warp; blk < nbr * nbc; blk += nwarp) {
          const int R0 = (blk % nbr) << 4, C0 = (blk / nbr) << 4;
          double cr[2][2][2], ci[2][2][2];
#pragma unroll
          for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int i = R0 + 8 * rt + g, j = C0 + 8 * ct + 2 * t4 + e;
                double2 v = make_double2(0.0, 0.0);
                if (i < mt && j < nc) v = M[(size_t)(k0 + nb + j) * n + k0 + nb + i];
                cr[rt][ct][e] = v.x; ci[rt][ct][e] = v.y;
              }
#pragma unroll
          for (int ks = 0; ks < NBI / 4; ++ks) {
            const int k = 4 * ks + t4;
            double lr[2], li[2], ur[2], ui[2];
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
              const double2 l = panel[k * ldp + nb + R0 + 8 * rt + g];
              lr[rt] = l.x; li[rt] = l.y;
            }
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
              const int j = min(C0 + 8 * ct + g, nc - 1);
              const double2 uu = M[(size_t)(k0 + nb + j) * n + k0 + k];
              ur[ct] = uu.x; ui[ct] = uu.y;
            }
#pragma unroll
            for (int rt = 0; rt < 2; ++rt)
#pragma unroll
              for (int ct = 0; ct < 2; ++ct) {
                // C -= L U :  Cr += (-Lr) Ur + Li Ui ;  Ci += (-Lr) Ui + (-Li) Ur.  The two products of an accumulator are issued
                // eight instructions apart (first halves of all eight accumulators, then the second halves): a dependent DMMA
                // back to back would stall the warp for the pipe's latency.  Same order per accumulator, same bits.
                dmma884(cr[rt][ct][0], cr[rt][ct][1], -lr[rt], ur[ct]);
                dmma884(ci[rt][ct][0], ci[rt][ct][1], -lr[rt], ui[ct]);
              }
#pragma unroll
            for (int rt = 0; rt < 2; ++rt)
#pragma unroll
              for (int ct = 0; ct < 2; ++ct) {
                dmma884(cr[rt][ct][0], cr[rt][ct][1], li[rt], ui[ct]);
                dmma884(ci[rt][ct][0], ci[rt][ct][1], -li[rt], ur[ct]);
              }
          }
#pragma unroll
          for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int i = R0 + 8 * rt + g, j = C0 + 8 * ct + 2 * t4 + e;
                if (i < mt && j < nc)
                  M[(size_t)(k0 + nb + j) * n + k0 + nb + i] = make_double2(cr[rt][ct][e], ci[rt][ct][e]);
              }
        }
      }
      __syncthreads();
      PROF_ADD(4);
    }
    // ================= the rest of the matrix, once per outer panel =================
    PROF_T0();
    // ---- the nbo interchanges on the columns outside the outer panel (two interleaved columns per thread)
    {
      const int nout = n - nbo;  // columns 0..K0-1 and KE..n-1
      for (int e = tid; e < nout; e += 2 * nt) {
        const int eb = e + nt;
        const bool dob = eb < nout;
        const int ca = e < K0 ? e : e + nbo, cb = dob ? (eb < K0 ? eb : eb + nbo) : ca;
        double2 *cola = M + (size_t)ca * n, *colb = M + (size_t)cb * n;
        for (int j = 0; j < nbo; ++j) {
          const int p = ipiv[K0 + j];
          if (p != K0 + j) {
            const double2 ta = cola[K0 + j], pa = cola[p];
            double2 tb, pb;
            if (dob) { tb = colb[K0 + j]; pb = colb[p]; }
            cola[K0 + j] = pa; cola[p] = ta;
            if (dob) { colb[K0 + j] = pb; colb[p] = tb; }
          }
        }
      }
    }
    __syncthreads();
    PROF_ADD(2);
    const int MT = n - KE;  // trailing size
    if (MT <= 0) break;
    // ---- U12 = L11^{-1} A12 (nbo x MT): L11 (unit lower, nbo x nbo) staged in shared memory [col][row]
    double2 *l11 = panel;  // the panel buffer is free between outer-panel factorisations
    for (int e = tid; e < nbo * nbo; e += nt) {
      const int r = e % nbo, c = e / nbo;
      l11[c * CH_NBO + r] = M[(size_t)(K0 + c) * n + K0 + r];
    }
    __syncthreads();
    for (int c = tid; c < MT; c += nt) {
      double2 *col = M + (size_t)(KE + c) * n + K0;
      for (int i0 = 0; i0 < nbo; i0 += CH_NB) {   // 16 rows at a time: registers hold one chunk
        const int ni = min(CH_NB, nbo - i0);
        double2 a[CH_NB];
#pragma unroll
        for (int i = 0; i < CH_NB; ++i) a[i] = (i < ni) ? col[i0 + i] : make_double2(0.0, 0.0);
        for (int t = 0; t < i0; ++t) {            // rows solved in earlier chunks
          const double2 ut = col[t];
#pragma unroll
          for (int i = 0; i < CH_NB; ++i)
            if (i < ni) a[i] = cfnma2(a[i], l11[t * CH_NBO + i0 + i], ut);
        }
#pragma unroll
        for (int i = 1; i < CH_NB; ++i) {
          if (i < ni) {
            double2 sacc = a[i];
#pragma unroll
            for (int t = 0; t < i; ++t) sacc = cfnma2(sacc, l11[(i0 + t) * CH_NBO + i0 + i], a[t]);
            a[i] = sacc;
          }
        }
#pragma unroll
        for (int i = 0; i < CH_NB; ++i) if (i < ni) col[i0 + i] = a[i];
      }
    }
    __syncthreads();
    PROF_ADD(3);
    // ---- A22 -= L21 U12 with k = nbo, operands staged in shared memory (the panel buffer is free here)
    lu2_big_update<NBO>(M, n, K0, KE, nbo, MT, panel);
    __syncthreads();
    PROF_ADD(4);
  }
}

// x_r - sum_t m[t] xs[t] over the nb <= 16 solved values of a block step, summed in column order.  (Four independent
// partial sums would shorten the chain of dependent DFMAs, but measured no faster and the rounding noise of the solve
// rose enough to stall 5 % of the N = 512 inverse iterations above the 1e-12 criterion.)
__device__ __forceinline__ double2 solve_row16(double2 xr, const double2 (&m)[16], const double2 *xs, int nb) {
  double2 acc = xr;
#pragma unroll
  for (int t = 0; t < 16; ++t) if (t < nb) acc = cfnma2(acc, m[t], xs[t]);
  return acc;
}

// x <- M^{-1} x using P M = L U.  x: shared, n complex.  Blocked by 16 columns with a one-block look-ahead: while
// warps 1.. apply the 16 values solved last to the rows further down (coalesced column reads), warp 0 applies them
// to the 16 rows of the NEXT diagonal block and solves that block with shuffles -- the serial chain of the diagonal
// solve runs beside the row update instead of in front of it, one barrier per block step.  The diagonal blocks are
// staged in shared memory two steps ahead by the updating warps.
__device__ void cta_solve(const double2 *__restrict__ M, int n, const int *ipiv, double2 *x) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int ut = tid - 32, nut = nt - 32;  // the updating threads (warps 1..)
  __shared__ double2 dblk[2][16][17];
  // Large matrices are streamed from DRAM by every solve (hundreds of concurrent 4 MB factorisations do not fit the
  // L2): the NEXT step's columns are requested into the L2 one step ahead (one prefetch per 128-byte line).
  const bool pf = n > 128;
  // n >= 300 (the factorisation streams from DRAM): all 16 column entries of a row are requested before the first
  // one is used -- one memory latency per row instead of one per unrolled group (+4.5 % at N = 384 / 512; below,
  // with the matrix L2-resident, the extra registers cost more than the latency: -4 % at N = 128 / 256)
  const bool deep = n >= 300;
  cta_permute(x, lu_perm(ipiv, n), n, true);  // x <- P x
  // ---------------- L y = P b (unit lower) ----------------
  auto stage_l = [&](int b, int j0, int e) {  // entry e of the diagonal block at j0 -> dblk[b]
    const int r = e & 15, c = e >> 4;
    if (e < 256 && j0 + r < n && j0 + c < n) dblk[b][r][c] = M[(size_t)(j0 + c) * n + j0 + r];
  };
  auto diag_l = [&](int b, int j0, double2 xi) {  // warp 0: forward substitution inside the block at j0
    const int nb = min(16, n - j0);
    for (int t = 0; t < nb; ++t) {
      const double xr = __shfl_sync(0xffffffffu, xi.x, t), xim = __shfl_sync(0xffffffffu, xi.y, t);
      if (lane > t && lane < nb) xi = cfnma2(xi, dblk[b][lane][t], make_double2(xr, xim));
    }
    if (lane < nb) x[j0 + lane] = xi;
  };
  for (int e = tid; e < 512; e += nt) stage_l(e >> 8, (e >> 8) * 16, e & 255);
  __syncthreads();
  if (warp == 0) diag_l(0, 0, lane < min(16, n) ? x[lane] : make_double2(0.0, 0.0));
  __syncthreads();
  int cur = 1;  // dblk[cur] holds the block after the one just solved
  for (int j0 = 0; j0 + 16 < n; j0 += 16, cur ^= 1) {
    const int jn = j0 + 16;
    if (warp == 0) {
      const int nbn = min(16, n - jn);
      double2 xi = make_double2(0.0, 0.0);
      if (lane < nbn) {
        double2 m[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) m[t] = M[(size_t)(j0 + t) * n + jn + lane];
        xi = solve_row16(x[jn + lane], m, x + j0, 16);
      }
      diag_l(cur, jn, xi);
    } else {
      if (pf && jn + 16 < n)
        for (int r = jn + 16 + 8 * ut; r < n; r += 8 * nut)
#pragma unroll
          for (int t = 0; t < 16; ++t) prefetch_l2(M + (size_t)(jn + t) * n + r);
      if (deep) {
        for (int r = jn + 16 + ut; r < n; r += nut) {
          double2 m[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) m[t] = M[(size_t)(j0 + t) * n + r];
          x[r] = solve_row16(x[r], m, x + j0, 16);
        }
      } else {
        for (int r = jn + 16 + ut; r < n; r += nut) {
          double2 acc = x[r];
#pragma unroll 4
          for (int t = 0; t < 16; ++t) acc = cfnma2(acc, M[(size_t)(j0 + t) * n + r], x[j0 + t]);
          x[r] = acc;
        }
      }
      for (int e = ut; e < 256; e += nut) stage_l(cur ^ 1, jn + 16, e);
    }
    __syncthreads();
  }
  // ---------------- U x = y ----------------
  // The diagonal blocks are staged as D^{-1} U (unit upper, the reciprocal of the pivot on the diagonal): the serial
  // chain of the 16 back-substitution steps then holds no division
  auto stage_u = [&](int b, int j0, int e) {
    const int r = e & 15, c = e >> 4;
    if (e < 256 && j0 >= 0 && j0 + r < n && j0 + c < n && c >= r) {
      const double2 a = M[(size_t)(j0 + c) * n + j0 + r];
      double2 d = M[(size_t)(j0 + r) * n + j0 + r];
      if (d.x == 0.0 && d.y == 0.0) d = make_double2(1.0e-290, 0.0);
      const double sc = 1.0 / fma(d.x, d.x, d.y * d.y);
      const double2 dinv = make_double2(d.x * sc, -d.y * sc);
      dblk[b][r][c] = (c == r) ? dinv : cmul2(a, dinv);
    }
  };
  auto diag_u = [&](int b, int j0, double2 xi) {  // warp 0: back substitution inside the block at j0
    const int nb = min(16, n - j0);
    if (lane < nb) xi = cmul2(xi, dblk[b][lane][lane]);
    for (int t = nb - 1; t >= 0; --t) {
      const double xr = __shfl_sync(0xffffffffu, xi.x, t), xim = __shfl_sync(0xffffffffu, xi.y, t);
      if (lane < t) xi = cfnma2(xi, dblk[b][lane][t], make_double2(xr, xim));
    }
    if (lane < nb) x[j0 + lane] = xi;
  };
  const int jlast = ((n - 1) / 16) * 16;
  for (int e = tid; e < 512; e += nt) stage_u(e >> 8, jlast - (e >> 8) * 16, e & 255);
  __syncthreads();
  if (warp == 0) diag_u(0, jlast, lane < n - jlast ? x[jlast + lane] : make_double2(0.0, 0.0));
  __syncthreads();
  cur = 1;
  for (int j0 = jlast; j0 >= 16; j0 -= 16, cur ^= 1) {
    const int nb = min(16, n - j0), jp = j0 - 16;  // the block above is always full
    if (warp == 0) {
      double2 xi = make_double2(0.0, 0.0);
      if (lane < 16) {
        double2 m[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) m[t] = M[(size_t)(j0 + min(t, nb - 1)) * n + jp + lane];
        xi = solve_row16(x[jp + lane], m, x + j0, nb);
      }
      diag_u(cur, jp, xi);
    } else {
      if (pf && jp >= 16)
        for (int r = 8 * ut; r < jp; r += 8 * nut)
#pragma unroll
          for (int t = 0; t < 16; ++t) prefetch_l2(M + (size_t)(jp + t) * n + r);
      if (deep) {
        for (int r = ut; r < jp; r += nut) {
          double2 m[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) m[t] = M[(size_t)(j0 + min(t, nb - 1)) * n + r];
          x[r] = solve_row16(x[r], m, x + j0, nb);
        }
      } else {
        for (int r = ut; r < jp; r += nut) {
          double2 acc = x[r];
#pragma unroll 4
          for (int t = 0; t < nb; ++t) acc = cfnma2(acc, M[(size_t)(j0 + t) * n + r], x[j0 + t]);
          x[r] = acc;
        }
      }
      for (int e = ut; e < 256; e += nut) stage_u(cur ^ 1, jp - 16, e);
    }
    __syncthreads();
  }
}

// x <- M^{-H} x :  M^H = U^H L^H P
__device__ void cta_solve_h(const double2 *__restrict__ M, int n, const int *ipiv, double2 *x) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int j = 0; j < n; ++j) {  // U^H a = r : lower triangular, row j of U^H = conj(column j of U)
    const double2 *col = M + (size_t)j * n;
    // a_j = (r_j - sum_{i<j} conj(U(i,j)) a_i) / conj(U(j,j)) : dot product over the column
    double2 s = make_double2(0.0, 0.0);
    for (int i = tid; i < j; i += nt) s = cadd2(s, cmulc2(col[i], x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
      s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    }
    __shared__ double2 part[CH_THREADS_MAX / 32];
    if ((tid & 31) == 0) part[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
      double2 tot = make_double2(0.0, 0.0);
      for (int w = 0; w < (nt >> 5); ++w) tot = cadd2(tot, part[w]);
      const double2 d = make_double2(col[j].x, -col[j].y);
      x[j] = cdiv2(csub2(x[j], tot), d);
    }
    __syncthreads();
  }
  for (int j = n - 1; j >= 0; --j) {  // L^H b = a : unit upper, row j of L^H = conj(column j of L)
    const double2 *col = M + (size_t)j * n;
    double2 s = make_double2(0.0, 0.0);
    for (int i = j + 1 + tid; i < n; i += nt) s = cadd2(s, cmulc2(col[i], x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
      s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    }
    __shared__ double2 part2[CH_THREADS_MAX / 32];
    if ((tid & 31) == 0) part2[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
      double2 tot = make_double2(0.0, 0.0);
      for (int w = 0; w < (nt >> 5); ++w) tot = cadd2(tot, part2[w]);
      x[j] = csub2(x[j], tot);
    }
    __syncthreads();
  }
  cta_permute(x, lu_perm(ipiv, n), n, false);  // z = P^T b
}

// y = D2 x (or D2^T x); D2 real column-major n x n; x, y shared
__device__ void cta_d2_matvec(const double *__restrict__ D2, int n, const double2 *x, double2 *y, bool transpose) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (!transpose) {
    for (int i = tid; i < n; i += nt) {
      double sr = 0.0, si = 0.0;
      for (int j = 0; j < n; ++j) {
        const double d = D2[(size_t)j * n + i];
        sr = fma(d, x[j].x, sr); si = fma(d, x[j].y, si);
      }
      y[i] = make_double2(sr, si);
    }
  } else {
    for (int j = tid; j < n; j += nt) {
      const double *col = D2 + (size_t)j * n;
      double sr = 0.0, si = 0.0;
      for (int i = 0; i < n; ++i) { sr = fma(col[i], x[i].x, sr); si = fma(col[i], x[i].y, si); }
      y[j] = make_double2(sr, si);
    }
  }
  __syncthreads();
}

struct ChanShared {
  double2 *panel, *x, *y, *t;
  int *ipiv;
  double *red;
};

// n >= CH_BIG_MIN_N run the kernels with the two-level LU (two 256-thread CTAs per SM while they fit).  The threshold was
// 351 (where two CTAs of the single-level kernel stop fitting an SM) until the two-level LU was measured below it
// (4096 alpha, eig/s, single-level / two-level): N = 184: 227 k / 176 k; 200: 174 k / 183 k; 256: 135.5 k / 136.9 k;
// 288: 93 k / 108 k; 320: 74 k / 86 k; 350: 36 k / 67 k.  OSB_CHAN_BIG_MIN overrides.
constexpr int CH_BIG_MIN_N = 185;
// LU variants: 0 = single-level, 16-column panel (cta_lu); 1 = two-level <16, 64>; 2 = two-level <8, 32>
__host__ __device__ inline size_t chan_panel_elems(int n, int variant) {
  if (variant == 0) return (size_t)(n + 16) * CH_NB;
  const size_t p = (size_t)(n + 16) * (variant == 1 ? 16 : 8), st = chan_stage_elems(variant == 1 ? 64 : 32);  // (NBO)
  return p > st ? p : st;
}
__device__ ChanShared carve(unsigned char *smem, int n, int variant = 0) {
  ChanShared s;
  size_t off = 0;
  s.panel = (double2 *)(smem + off); off += sizeof(double2) * chan_panel_elems(n, variant);
  s.x = (double2 *)(smem + off); off += sizeof(double2) * (size_t)n;
  s.y = (double2 *)(smem + off); off += sizeof(double2) * (size_t)n;
  s.t = (double2 *)(smem + off); off += sizeof(double2) * (size_t)n;
  s.red = (double *)(smem + off); off += sizeof(double) * 64;
  s.ipiv = (int *)(smem + off);
  return s;
}

// Small operators (n <= 96: chan.inp, nc.inp, fd-64.inp) keep the whole matrix in shared
// memory next to the panel, so every phase runs at shared-memory latency.
constexpr int CH_NSMEM = 64;
__host__ __device__ inline bool chan_matrix_in_smem(int n) { return n <= CH_NSMEM; }
__host__ __device__ inline size_t chan_fixed_smem(int n, int variant = 0) {
  size_t b = sizeof(double2) * (chan_panel_elems(n, variant) + 3 * (size_t)n) + sizeof(double) * 64 + sizeof(int) * (size_t)(lu_perm_offset(n) + n);
  return (b + 15) & ~(size_t)15;
}
size_t chan_smem_bytes(int n) {
  return chan_fixed_smem(n) + (chan_matrix_in_smem(n) ? sizeof(double2) * (size_t)n * n : 0);
}

// Convergence of the (linear) inverse iteration at a fixed shift.  Successive changes of omega contract by
// rho = d0 / d1, so after a change d0 the remaining error of the current estimate is ~ d0 rho / (1 - rho).
// The iteration stops when the measured change is below the goal, or -- one solve earlier -- when that
// geometric-series estimate is (with a factor 2 of margin).  The extrapolation is only trusted in the
// asymptotic regime: three measured changes between successive estimates (the first "change", measured
// against the shift, does not count), two contraction ratios that are both fast (< 0.05) and do not
// deteriorate (rho <= 2 rho_prev).  `est` is what the kernels report as delta: the last MEASURED change, or,
// after an early stop, the remaining-error estimate.  OSB_NO_EARLY_STOP=1 (a.early = 0) disables the
// extrapolation (tests compare the two).
struct ConvTrack {
  double d0, d1, d2;  // the last three measured changes of omega, d0 newest
  double dmax;        // the largest change between successive estimates seen so far
};
__device__ __forceinline__ void conv_reset(ConvTrack &t) { t.d0 = t.d1 = t.d2 = 1.0e300; t.dmax = 0.0; }
// returns 0: keep iterating; 1: converged (measured change below the goal, or the early stop above);
// 2: stalled at the rounding floor -- the changes have dropped by 1e6 from their maximum and stopped
// contracting although the goal is not met (ill-conditioned pencils: Chebyshev collocation at N >= 256, SURVEY
// H7).  More solves, or a re-shift with a new factorisation, cannot improve such a point; *est stays the
// measured change and the host reports the point as not converged if that is above its 1e-12 criterion.
__device__ __forceinline__ int conv_push(ConvTrack &t, int it, double change, double goal, int early, double *est) {
  t.d2 = t.d1; t.d1 = t.d0; t.d0 = change;
  *est = change;
  if (it > 0 && change <= goal) return 1;
  if (it >= 1) t.dmax = fmax(t.dmax, change);
  if (early && it >= 3 && t.d2 < 1.0e299) {
    const double rho = t.d0 / t.d1, rho1 = t.d1 / t.d2;
    if (rho < 0.05 && rho1 < 0.05 && rho <= 2.0 * rho1) {
      const double rem = t.d0 * rho / (1.0 - rho);
      if (rem <= 0.5 * goal) { *est = rem; return 1; }
    }
  }
  if (it >= 3 && t.d0 >= 0.5 * t.d1 && t.d0 <= 1.0e-6 * t.dmax) return 2;
  return 0;
}

// Right inverse iteration with the current LU: x (unit 2-norm, shared) in/out.
// Returns the eigenvalue estimate after up to maxit iterations; *delta = last change.
__device__ double2 inverse_iterate(const double2 *M, int n, const double *D2, const PencilCoef &p,
                                   double2 sigma, ChanShared &s, int maxit, double tol, int early, double *delta,
                                   int *iters, int *stalled = nullptr) {
  double2 om = sigma;
  ConvTrack ct;
  conv_reset(ct);
  double est = 1.0e300;
  int it = 0;
  // maxit is the nominal budget of one shift.  A dense refactorisation costs 6-9 solves, so when
  // the budget is used up the iteration goes on (up to 3 budgets) as long as the observed
  // contraction says that at most 6 more solves reach the tolerance.
  for (; it < 3 * maxit; ++it) {
    PROF_T0();
    // y = B x = b2 D2 x + b0 x
    cta_d2_matvec(D2, n, s.x, s.y, false);
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      s.y[i] = cadd2(cmul2(p.b2, s.y[i]), cmul2(p.b0, s.x[i]));
    __syncthreads();
    PROF_ADD(5);
    cta_solve(M, n, s.ipiv, s.y);
    PROF_ADD(6);
    double2 yx = make_double2(0.0, 0.0), yy = make_double2(0.0, 0.0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      yx = cadd2(yx, cmulc2(s.y[i], s.x[i]));
      yy.x += s.y[i].x * s.y[i].x + s.y[i].y * s.y[i].y;
    }
    yx = block_sum(yx, s.red);
    yy = block_sum(yy, s.red);
    const double2 om_new = cadd2(sigma, make_double2(yx.x / yy.x, yx.y / yy.x));
    const double rn = rsqrt(yy.x);
    for (int i = threadIdx.x; i < n; i += blockDim.x) s.x[i] = make_double2(s.y[i].x * rn, s.y[i].y * rn);
    __syncthreads();
    const double change = hypot(om_new.x - om.x, om_new.y - om.y);
    om = om_new;
    const double goal = tol * hypot(om.x, om.y);
    const int cv = conv_push(ct, it, change, goal, early, &est);
    if (cv) { ++it; if (stalled) *stalled = (cv == 2); break; }
    if (it + 1 >= maxit) {
      const double rho = ct.d0 / ct.d1;
      if (!(rho < 0.5) || !(ct.d0 > 0.0) || log(goal / ct.d0) / log(rho) > 6.0) { ++it; break; }
    }
  }
  *delta = est;
  *iters = it;
  return om;
}

struct ChanArgs {
  int n;                 // interior size N-1
  const double *D2, *D4; // n x n column-major
  const double *u, *d2u; // interior
  double2 *work;         // gridDim.x * n*n
  int64_t ninst;
  const double2 *alpha;  // [ninst]
  const double *Re;      // [ninst]
  const double2 *sigma;  // [ninst] shifts
  int outer, inner;      // re-shift rounds, iterations per round
  int early;             // allow the one-solve-early stop of conv_push
  double tol;
  double2 *omega;        // [ninst]
  double *delta;         // [ninst] error estimate of omega: last measured change, or conv_push's remaining-error estimate
  int32_t *iters;        // [ninst] total inverse iterations
  double2 *evec;         // optional [ninst][n+2] (grid 0..N, zeros at the walls), max-modulus normalised
  double2 *adj;          // optional [ninst][n+2] left (adjoint) eigenvector, see adj_form
  int adj_form;          // OSB_ADJ_PENCIL: u with u^H (A - w B) = 0 scaled to u^H evec = 1 (orrcolchan.f:537-545);
                         // OSB_ADJ_INVBA: B^H u = ZGEEV's left vector of inv(B)A, unit 2-norm, largest component
                         // real positive (orrfdchan.f:796)
};

// index of the component of largest modulus, lowest index on ties (the scan of orrfdchan.f:849-857)
__device__ int block_argmax_abs(const double2 *v, int n) {
  __shared__ double sb[CH_THREADS_MAX / 32];
  __shared__ int si[CH_THREADS_MAX / 32];
  double best = -1.0;
  int bi = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double m2 = v[i].x * v[i].x + v[i].y * v[i].y;
    if (m2 > best) { best = m2; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  __syncthreads();  // a previous call's readers are done with sb / si
  if ((threadIdx.x & 31) == 0) { sb[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  best = sb[0]; bi = si[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
    if (sb[w] > best || (sb[w] == best && si[w] < bi)) { best = sb[w]; bi = si[w]; }
  return bi;
}

__device__ void start_vector(double2 *x, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double s = sin(3.0 * (i + 1.0) / (n + 1.0)) + 0.37 * cos(11.0 * (i + 1.0) / (n + 1.0));
    x[i] = make_double2(s, 0.25 * s * s - 0.1);
  }
  __syncthreads();
}

// Left (adjoint) eigenvector of the converged mode; kept out of line so that its second LU call site does not
// weigh on the register allocation of the sweep path.
template <int LUV>
__device__ __noinline__ void cta_adjoint(const ChanArgs &a, ChanShared &s, double2 *M, const PencilCoef &p, double2 om,
                                         double2 esc, int64_t inst) {
  const int n = a.n;
  // Left eigenvector u, u^H (A - w B) = 0, by inverse iteration on the adjoint pencil
  //     u <- (A - s B)^{-H} B^H u
  // with a fresh factorisation at s = w (1 + 1e-8): every solve then contracts the other modes by ~1e-8 |w| / gap.
  const double2 sg = make_double2(om.x + 1.0e-8 * hypot(om.x, om.y), om.y);
  assemble_shifted(M, n, a.D2, a.D4, a.u, a.d2u, p, sg);
  if (LUV == 1) cta_lu2<16, 64>(M, n, s.ipiv, s.panel, s.red);
  else if (LUV == 2) cta_lu2<8, 32>(M, n, s.ipiv, s.panel, s.red);
  else cta_lu(M, n, s.ipiv, s.panel, s.red);
  for (int i = threadIdx.x; i < n; i += blockDim.x) s.t[i] = make_double2(s.x[i].x, 0.3 - s.x[i].y);
  __syncthreads();
  for (int k = 0; k < 4; ++k) {
    cta_d2_matvec(a.D2, n, s.t, s.y, true);  // D2^T u
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      s.y[i] = cadd2(cmulc2(p.b2, s.y[i]), cmulc2(p.b0, s.t[i]));  // B^H u
    __syncthreads();
    cta_solve_h(M, n, s.ipiv, s.y);
    double nn = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) nn += s.y[i].x * s.y[i].x + s.y[i].y * s.y[i].y;
    const double rn = rsqrt(block_sum(make_double2(nn, 0.0), s.red).x);
    for (int i = threadIdx.x; i < n; i += blockDim.x) s.t[i] = make_double2(s.y[i].x * rn, s.y[i].y * rn);
    __syncthreads();
  }
  double2 *av = a.adj + (size_t)inst * (n + 2);
  if (a.adj_form == OSB_ADJ_PENCIL) {
    // escale = zdotc(avec, evec); avec <- avec / conj(escale)   (orrcolchan.f:537-545)
    double2 d = make_double2(0.0, 0.0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) d = cadd2(d, cmulc2(s.t[i], cmul2(esc, s.x[i])));
    d = block_sum(d, s.red);
    const double2 sc = cdiv2(make_double2(1.0, 0.0), make_double2(d.x, -d.y));
    for (int i = threadIdx.x; i < n; i += blockDim.x) av[i + 1] = cmul2(sc, s.t[i]);
  } else {
    // vl = B^H u, then ZGEEV's normalisation: unit 2-norm, largest component real (positive)
    cta_d2_matvec(a.D2, n, s.t, s.y, true);
    double nn = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double2 v = cadd2(cmulc2(p.b2, s.y[i]), cmulc2(p.b0, s.t[i]));
      s.y[i] = v;
      nn += v.x * v.x + v.y * v.y;
    }
    const double rn = rsqrt(block_sum(make_double2(nn, 0.0), s.red).x);
    const int k = block_argmax_abs(s.y, n);
    const double2 vk = s.y[k];
    const double ak = hypot(vk.x, vk.y);
    const double2 ph = make_double2(rn * vk.x / ak, -rn * vk.y / ak);  // rn * conj(v_k) / |v_k|
    for (int i = threadIdx.x; i < n; i += blockDim.x) av[i + 1] = cmul2(ph, s.y[i]);
  }
  if (threadIdx.x == 0) { av[0] = make_double2(0.0, 0.0); av[n + 1] = make_double2(0.0, 0.0); }
}

// Two register budgets for the same body: up to n ~ 350 two CTAs fit an SM by shared
// memory and 128 registers/thread keeps them both resident; beyond that the 135 KB panel
// allows one CTA only and the body may use the full register file.
template <int LUV>
__device__ __forceinline__ void chan_eig_body(const ChanArgs &a);
__global__ void __launch_bounds__(CH_THREADS, 2) chan_eig_kernel(const __grid_constant__ ChanArgs a) { chan_eig_body<0>(a); }
// n >= CH_BIG_MIN_N, three builds (OSB_CHAN_BIG = 2 | 1 | 0 selects; default 2 while two CTAs fit an SM):
//   _big2    256 threads, two CTAs per SM, two-level LU <8, 32>
//   _big     512 threads, one CTA per SM, two-level LU <16, 64>
//   _big_lu1 512 threads, one CTA per SM, the single-level LU (the round-1 kernel, kept for A/B measurements)
__global__ void __launch_bounds__(CH_THREADS, 2) chan_eig_kernel_big2(const __grid_constant__ ChanArgs a) { chan_eig_body<2>(a); }
__global__ void __launch_bounds__(CH_THREADS_BIG, 1) chan_eig_kernel_big(const __grid_constant__ ChanArgs a) { chan_eig_body<1>(a); }
__global__ void __launch_bounds__(CH_THREADS_BIG, 1) chan_eig_kernel_big_lu1(const __grid_constant__ ChanArgs a) { chan_eig_body<0>(a); }
// 64 < n <= 200: the serial phases (panel, triangular solves) dominate and a row per thread is all the
// panel can use; four small CTAs per SM overlap each other's barriers
__global__ void __launch_bounds__(CH_THREADS_SMALL, 4) chan_eig_kernel_small(const __grid_constant__ ChanArgs a) { chan_eig_body<0>(a); }
template <int LUV>
__device__ __forceinline__ void chan_eig_body(const ChanArgs &a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int n = a.n;
  ChanShared s = carve(smem, n, LUV);
  double2 *M = chan_matrix_in_smem(n) ? (double2 *)(smem + chan_fixed_smem(n)) : a.work + (size_t)blockIdx.x * n * n;
  for (int64_t inst = blockIdx.x; inst < a.ninst; inst += gridDim.x) {
    const PencilCoef p = pencil_coef(a.alpha[inst], a.Re[inst]);
    double2 sigma = a.sigma[inst];
    start_vector(s.x, n);
    double2 om = sigma;
    double dl = 1.0e300;
    int total = 0;
    for (int o = 0; o < a.outer; ++o) {
      PROF_T0();
      assemble_shifted(M, n, a.D2, a.D4, a.u, a.d2u, p, sigma);
      PROF_ADD(0);
      if (LUV == 1) cta_lu2<16, 64>(M, n, s.ipiv, s.panel, s.red);
      else if (LUV == 2) cta_lu2<8, 32>(M, n, s.ipiv, s.panel, s.red);
      else cta_lu(M, n, s.ipiv, s.panel, s.red);
      int it = 0, stalled = 0;
      om = inverse_iterate(M, n, a.D2, p, sigma, s, a.inner, a.tol, a.early, &dl, &it, &stalled);
      total += it;
      if (dl <= a.tol * hypot(om.x, om.y) || stalled) break;
      sigma = om;  // re-shift (Rayleigh-quotient style) and refactorise
    }
    if (threadIdx.x == 0) {
      a.omega[inst] = om;
      a.delta[inst] = dl;
      a.iters[inst] = total;
    }
    if (a.evec || a.adj) {
      // normalise by the component of largest modulus (orrfdchan.f:849-857)
      const int imax = block_argmax_abs(s.x, n);
      const double2 esc = cdiv2(make_double2(1.0, 0.0), s.x[imax]);
      if (a.evec) {
        double2 *ev = a.evec + (size_t)inst * (n + 2);
        for (int i = threadIdx.x; i < n; i += blockDim.x) ev[i + 1] = cmul2(esc, s.x[i]);
        if (threadIdx.x == 0) { ev[0] = make_double2(0.0, 0.0); ev[n + 1] = make_double2(0.0, 0.0); }
      }
      if (a.adj) cta_adjoint<LUV>(a, s, M, p, om, esc, inst);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// K4c: the FULL spectrum of one pencil -- what SOLVE_ORR_SOM computes before it keeps one mode
// (orrfdchan.f:792-797: C = inv(B) A by ZGETRF/ZGETRI/ZGEMM, then ZGEEV; orrcolchan.f:505-513 lists it).
// Not on the sweep path (the sweep isolates its mode by shift-invert); it serves the interactive
// listings of the single-alpha programs and lets the tests apply the reference's selection rule to a
// spectrum computed on the device.  One CTA per pencil:
//   C = inv(B) A        with the blocked LU above (n triangular solves),
//   H = Q^H C Q         Householder reduction to Hessenberg form, CTA-parallel rank-1 updates,
//   eigenvalues of H    single-shift complex QR with Wilkinson shifts and LAPACK's (ZLAHQR) deflation
//                       test; rotations are applied to the active block only (eigenvalues only).
// ---------------------------------------------------------------------------
struct SpectrumArgs {
  int n;
  const double *D2, *D4, *u, *d2u;
  double2 *work;          // gridDim.x * 2 * n*n
  int64_t ninst;
  const double2 *alpha;   // [ninst]
  const double *Re;       // [ninst]
  double2 *eig;           // [ninst][n]  unsorted
  int32_t *status;        // [ninst]  0 ok, 1 QR iteration limit hit
};

__device__ __forceinline__ double cabs1d(double2 z) { return fabs(z.x) + fabs(z.y); }
__device__ __forceinline__ double2 csqrt2(double2 z) {
  const zd r = zsqrt(zd{z.x, z.y});
  return make_double2(r.x, r.y);
}

// B = b2 D2 + b0 I, column-major n x n
__device__ void assemble_B(double2 *__restrict__ M, int n, const double *__restrict__ D2, const PencilCoef &p) {
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    for (int j = 0; j < n; ++j) {
      const size_t e = (size_t)j * n + i;
      double2 v = make_double2(p.b2.x * D2[e], p.b2.y * D2[e]);
      if (i == j) v = cadd2(v, p.b0);
      M[e] = v;
    }
  __syncthreads();
}

// H <- Q^H H Q (upper Hessenberg), column-major, ld = n.  v: shared, n complex
__device__ void cta_hessenberg(double2 *H, int n, double2 *v, double *red) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  for (int k = 0; k < n - 2; ++k) {
    const int m = n - k - 1;                 // length of the column below the diagonal
    double2 *col = H + (size_t)k * n + k + 1;
    double2 part = make_double2(0.0, 0.0);
    for (int i = tid; i < m; i += nt) { const double2 x = col[i]; part.x += x.x * x.x + x.y * x.y; v[i] = x; }
    const double xn = sqrt(block_sum(part, red).x);
    __syncthreads();
    if (xn == 0.0) continue;
    const double2 a0 = v[0];
    const double a0n = hypot(a0.x, a0.y);
    const double2 ph = a0n > 0.0 ? make_double2(a0.x / a0n, a0.y / a0n) : make_double2(1.0, 0.0);
    const double2 beta = make_double2(-ph.x * xn, -ph.y * xn);
    __syncthreads();
    if (tid == 0) v[0] = csub2(a0, beta);
    __syncthreads();
    // |v|^2 = |x|^2 - |a0|^2 + |a0 - beta|^2
    const double2 v0 = v[0];
    const double vn2 = xn * xn - a0n * a0n + (v0.x * v0.x + v0.y * v0.y);
    if (!(vn2 > 0.0)) continue;
    const double tau = 2.0 / vn2;
    // left: H[k+1:, j] -= tau v (v^H H[k+1:, j]) for j = k..n-1, a warp per column
    for (int j = k + warp; j < n; j += nwarp) {
      double2 *cj = H + (size_t)j * n + k + 1;
      double2 w = make_double2(0.0, 0.0);
      for (int i = lane; i < m; i += 32) w = cadd2(w, cmulc2(v[i], cj[i]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        w.x += __shfl_xor_sync(0xffffffffu, w.x, o);
        w.y += __shfl_xor_sync(0xffffffffu, w.y, o);
      }
      w = make_double2(tau * w.x, tau * w.y);
      for (int i = lane; i < m; i += 32) cj[i] = csub2(cj[i], cmul2(v[i], w));
    }
    __syncthreads();
    // right: H[i, k+1:] -= tau (H[i, k+1:] v) v^H, a thread per row
    for (int i = tid; i < n; i += nt) {
      double2 r = make_double2(0.0, 0.0);
      for (int j = 0; j < m; ++j) r = cadd2(r, cmul2(H[(size_t)(k + 1 + j) * n + i], v[j]));
      r = make_double2(tau * r.x, tau * r.y);
      for (int j = 0; j < m; ++j) {
        const double2 vc = make_double2(v[j].x, -v[j].y);
        double2 *e = H + (size_t)(k + 1 + j) * n + i;
        *e = csub2(*e, cmul2(r, vc));
      }
    }
    __syncthreads();
    for (int i = tid; i < m; i += nt) col[i] = (i == 0) ? beta : make_double2(0.0, 0.0);
    __syncthreads();
  }
}

// eigenvalues of the upper Hessenberg H (destroyed); returns 0 or 1 (iteration limit)
__device__ int cta_hqr(double2 *H, int n, double2 *w_out, double *red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ int sh_l, sh_fail;
  __shared__ double sh_c;
  __shared__ double2 sh_s, sh_t;
  const double eps = 2.220446049250313e-16, safmin = 2.2250738585072014e-308;
  const double smlnum = safmin * ((double)n / eps);
  int *red_i = (int *)red;
#define HQ(i, j) H[(size_t)(j) * n + (i)]
  if (tid == 0) sh_fail = 0;
  __syncthreads();
  for (int hi = n - 1; hi >= 0; --hi) {
    int its = 0;
    while (true) {
      // ---- largest l in (0, hi] with a negligible subdiagonal H(l, l-1)   (ZLAHQR's test)
      int best = 0;
      for (int l = hi - tid; l > 0; l -= nt) {
        const double2 sub = HQ(l, l - 1);
        bool small = cabs1d(sub) <= smlnum;
        if (!small) {
          double tst = cabs1d(HQ(l - 1, l - 1)) + cabs1d(HQ(l, l));
          if (tst == 0.0) {
            if (l - 2 >= 0) tst += fabs(HQ(l - 1, l - 2).x);
            if (l + 1 <= n - 1) tst += fabs(HQ(l + 1, l).x);
          }
          if (fabs(sub.x) <= eps * tst) {
            const double h12 = cabs1d(HQ(l - 1, l)), h21 = cabs1d(sub);
            const double ab = fmax(h21, h12), ba = fmin(h21, h12);
            const double d11 = cabs1d(HQ(l, l)), d = cabs1d(csub2(HQ(l - 1, l - 1), HQ(l, l)));
            const double aa = fmax(d11, d), bb = fmin(d11, d);
            const double sc = aa + ab;
            small = ba * (ab / sc) <= fmax(smlnum, eps * (bb * (aa / sc)));
          }
        }
        if (small) { best = l; break; }   // this thread's candidates are visited in descending order
      }
      // block-wide maximum
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
      __syncthreads();
      if ((tid & 31) == 0) red_i[tid >> 5] = best;
      __syncthreads();
      if (tid == 0) {
        int b = 0;
        for (int wv = 0; wv < (nt >> 5); ++wv) b = max(b, red_i[wv]);
        sh_l = b;
        if (b > 0) HQ(b, b - 1) = make_double2(0.0, 0.0);
      }
      __syncthreads();
      const int l = sh_l;
      if (l >= hi) break;          // H(hi,hi) has converged
      ++its;
      if (its > 300) { if (tid == 0) sh_fail = 1; break; }
      // ---- shift (thread 0)
      if (tid == 0) {
        double2 t;
        if (its % 10 == 0 && its % 20 != 0) {
          t = HQ(l, l); t.x += 0.75 * fabs(HQ(l + 1, l).x);
        } else if (its % 20 == 0) {
          t = HQ(hi, hi); t.x += 0.75 * fabs(HQ(hi, hi - 1).x);
        } else {
          t = HQ(hi, hi);
          const double2 uu = cmul2(csqrt2(HQ(hi - 1, hi)), csqrt2(HQ(hi, hi - 1)));
          double sc = cabs1d(uu);
          if (sc != 0.0) {
            const double2 x = make_double2(0.5 * (HQ(hi - 1, hi - 1).x - t.x), 0.5 * (HQ(hi - 1, hi - 1).y - t.y));
            const double sx = cabs1d(x);
            sc = fmax(sc, sx);
            const double2 xs = make_double2(x.x / sc, x.y / sc), us = make_double2(uu.x / sc, uu.y / sc);
            double2 y = csqrt2(cadd2(cmul2(xs, xs), cmul2(us, us)));
            y = make_double2(sc * y.x, sc * y.y);
            if (sx > 0.0 && (x.x / sx) * y.x + (x.y / sx) * y.y < 0.0) y = make_double2(-y.x, -y.y);
            t = csub2(t, cmul2(uu, cdiv2(uu, cadd2(x, y))));
          }
        }
        sh_t = t;
      }
      __syncthreads();
      // ---- one implicit single-shift QR sweep over the active block [l, hi]
      for (int k = l; k < hi; ++k) {
        if (tid == 0) {
          double2 a, b;
          if (k == l) { a = csub2(HQ(k, k), sh_t); b = HQ(k + 1, k); }
          else { a = HQ(k, k - 1); b = HQ(k + 1, k - 1); }
          double c;
          double2 sn, r;
          const double nb = hypot(b.x, b.y), na = hypot(a.x, a.y);
          if (nb == 0.0) { c = 1.0; sn = make_double2(0.0, 0.0); r = a; }
          else if (na == 0.0) { c = 0.0; sn = make_double2(b.x / nb, -b.y / nb); r = make_double2(nb, 0.0); }
          else {
            const double nrm = hypot(na, nb);
            const double2 ph = make_double2(a.x / na, a.y / na);
            c = na / nrm;
            const double2 cb = make_double2(b.x / nrm, -b.y / nrm);
            sn = cmul2(ph, cb);
            r = make_double2(ph.x * nrm, ph.y * nrm);
          }
          if (k > l) { HQ(k, k - 1) = r; HQ(k + 1, k - 1) = make_double2(0.0, 0.0); }
          sh_c = c; sh_s = sn;
        }
        __syncthreads();
        const double c = sh_c;
        const double2 sn = sh_s, snc = make_double2(sn.x, -sn.y);
        // rows k, k+1 over the columns k..hi
        for (int j = k + tid; j <= hi; j += nt) {
          const double2 x = HQ(k, j), y = HQ(k + 1, j);
          HQ(k, j) = cadd2(make_double2(c * x.x, c * x.y), cmul2(sn, y));
          HQ(k + 1, j) = csub2(make_double2(c * y.x, c * y.y), cmul2(snc, x));
        }
        __syncthreads();
        // columns k, k+1 over the rows l..min(k+2, hi)
        const int top = min(k + 2, hi);
        for (int i = l + tid; i <= top; i += nt) {
          const double2 x = HQ(i, k), y = HQ(i, k + 1);
          HQ(i, k) = cadd2(make_double2(c * x.x, c * x.y), cmul2(snc, y));
          HQ(i, k + 1) = csub2(make_double2(c * y.x, c * y.y), cmul2(sn, x));
        }
        __syncthreads();
      }
    }
    if (tid == 0) w_out[hi] = HQ(hi, hi);
    __syncthreads();
    if (sh_fail) {
      // give the remaining diagonal as it stands so that the caller sees finite numbers
      for (int i = tid; i < hi; i += nt) w_out[i] = HQ(i, i);
      __syncthreads();
      return 1;
    }
  }
#undef HQ
  return 0;
}

__global__ void __launch_bounds__(CH_THREADS) chan_spectrum_kernel(const __grid_constant__ SpectrumArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int n = a.n;
  ChanShared s = carve(smem, n);
  double2 *M1 = a.work + (size_t)blockIdx.x * 2 * n * n, *M2 = M1 + (size_t)n * n;
  for (int64_t inst = blockIdx.x; inst < a.ninst; inst += gridDim.x) {
    const PencilCoef p = pencil_coef(a.alpha[inst], a.Re[inst]);
    assemble_B(M1, n, a.D2, p);
    cta_lu(M1, n, s.ipiv, s.panel, s.red);
    assemble_shifted(M2, n, a.D2, a.D4, a.u, a.d2u, p, make_double2(0.0, 0.0));   // A
    for (int j = 0; j < n; ++j) {   // C(:, j) = inv(B) A(:, j)
      double2 *cj = M2 + (size_t)j * n;
      for (int i = threadIdx.x; i < n; i += blockDim.x) s.x[i] = cj[i];
      __syncthreads();
      cta_solve(M1, n, s.ipiv, s.x);
      for (int i = threadIdx.x; i < n; i += blockDim.x) cj[i] = s.x[i];
      __syncthreads();
    }
    cta_hessenberg(M2, n, s.x, s.red);
    const int st = cta_hqr(M2, n, a.eig + (size_t)inst * n, s.red);
    if (threadIdx.x == 0) a.status[inst] = st;
    __syncthreads();
  }
}

cudaError_t launch_chan_spectrum(int n, const double *D2, const double *D4, const double *u, const double *d2u,
                                 double2 *work, int grid, size_t smem, int64_t ninst, const double2 *alpha,
                                 const double *Re, double2 *eig, int32_t *status, cudaStream_t st) {
  SpectrumArgs a;
  a.n = n; a.D2 = D2; a.D4 = D4; a.u = u; a.d2u = d2u; a.work = work; a.ninst = ninst; a.alpha = alpha; a.Re = Re;
  a.eig = eig; a.status = status;
  chan_spectrum_kernel<<<grid, CH_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t chan_spectrum_grid(int n, int64_t ninst, int *grid, size_t *smem) {
  *smem = chan_fixed_smem(n);
  cudaError_t e = cudaFuncSetAttribute(chan_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0, per = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, chan_spectrum_kernel, CH_THREADS, *smem);
  if (e != cudaSuccess) return e;
  if (per < 1) return cudaErrorInvalidConfiguration;
  *grid = (int)std::min<int64_t>(ninst, (int64_t)sms * per);
  return cudaSuccess;
}

// ---------------------------------------------------------------------------
// K4b: the FD operator is banded.  D1hat is tridiagonal (FiniteD1, orrfdchan.f:1183-1215),
// so D2 = D1hat D1 has half bandwidth 2 and D4 half bandwidth 4: A - sigma B is a 9-diagonal
// matrix.  Gaussian elimination with partial pivoting on it is the SAME elimination as the
// dense LU (a pivot can only come from the 4 rows below the diagonal, fill-in stays within
// 8 super-diagonals) at n*4*8 complex FMAs instead of (8/3) n^3 flops.  One THREAD per
// instance: a 5-row x 9-column window of the active rows lives in registers, L (4 per
// column), U (9 per row), the pivots and the two iteration vectors live in global memory
// interleaved by instance so that a warp's accesses coalesce.
// ---------------------------------------------------------------------------
constexpr int BKL = 4, BKU = 4, BW = BKL + BKU + 1;  // 9 stored entries per U row

struct BandArgs {
  int n;
  const double *D2, *D4, *u, *d2u;  // dense tables (only the band is read)
  const double2 *band24;            // [n][BW]: (D2, D4)(r, r-4+j), zero outside the matrix (warp kernel)
  int64_t ninst;
  const double2 *alpha;
  const double *Re;
  const double2 *sigma;
  int outer, inner;
  int early;       // allow the one-solve-early stop of conv_push
  double tol;
  double2 *omega;
  double *delta;
  int32_t *iters;
  double2 *evec;   // optional [ninst][n+2]
  double2 *adj;    // optional [ninst][n+2] left eigenvector (one-thread kernel only), see ChanArgs
  int adj_form;
  double2 *U;      // [n][BW][ninst]
  double2 *L;      // [n][BKL][ninst]
  int8_t *piv;     // [n][ninst]  pivot offset 0..4 (one-thread kernel)
  int32_t *piv32;  // [ninst][n]  the same for the nine-lane kernel (4-byte cp.async granularity)
  double2 *x, *y;  // [n][ninst]
};

__device__ __forceinline__ double2 band_entry(const BandArgs &a, const PencilCoef &p, double2 sb2, double2 sb0,
                                              int r, int c) {
  // M(r,c) = D4 + (w2_r - sigma b2) D2 + delta_rc (w0_r - sigma b0); zero outside the matrix / band
  if (c < 0 || c >= a.n || r >= a.n || c < r - BKL || c > r + BKU) return make_double2(0.0, 0.0);
  const size_t e = (size_t)c * a.n + r;
  const double2 w2 = csub2(coef_w2(p, a.u[r]), sb2);
  const double d2 = a.D2[e], d4 = a.D4[e];
  double2 v = make_double2(fma(w2.x, d2, d4), w2.y * d2);
  if (r == c) v = cadd2(v, csub2(coef_w0(p, a.u[r], a.d2u[r]), sb0));
  return v;
}

// banded LU with partial pivoting of A - sigma B, one thread per instance (window of rows k..k+4, columns k..k+8)
__device__ __noinline__ void band_lu_thread(const BandArgs &a, const PencilCoef &p, double2 sigma, double2 *U,
                                               double2 *Lm, int8_t *piv, size_t S) {
  const int n = a.n;
  const double2 sb2 = cmul2(sigma, p.b2), sb0 = cmul2(sigma, p.b0);
  double2 w[BKL + 1][BW];
#pragma unroll
  for (int sidx = 0; sidx <= BKL; ++sidx)
#pragma unroll
    for (int j = 0; j < BW; ++j) w[sidx][j] = band_entry(a, p, sb2, sb0, sidx, j);
  for (int k = 0; k < n; ++k) {
    int pv = 0;
    double best = fabs(w[0][0].x) + fabs(w[0][0].y);
#pragma unroll
    for (int sidx = 1; sidx <= BKL; ++sidx) {
      const double v = fabs(w[sidx][0].x) + fabs(w[sidx][0].y);
      if (k + sidx < n && v > best) { best = v; pv = sidx; }
    }
    piv[(size_t)k * S] = (int8_t)pv;
#pragma unroll
    for (int sidx = 1; sidx <= BKL; ++sidx)
      if (pv == sidx) {
#pragma unroll
        for (int j = 0; j < BW; ++j) { const double2 t = w[0][j]; w[0][j] = w[sidx][j]; w[sidx][j] = t; }
      }
    double2 pvt = w[0][0];
    if (pvt.x == 0.0 && pvt.y == 0.0) pvt = make_double2(1.0e-290, 0.0);
    const double2 rinv = cdiv2(make_double2(1.0, 0.0), pvt);
#pragma unroll
    for (int sidx = 1; sidx <= BKL; ++sidx) {
      const double2 l = cmul2(w[sidx][0], rinv);
      Lm[((size_t)k * BKL + (sidx - 1)) * S] = l;
#pragma unroll
      for (int j = 1; j < BW; ++j) w[sidx][j] = cfnma2(w[sidx][j], l, w[0][j]);
    }
#pragma unroll
    for (int j = 0; j < BW; ++j) U[((size_t)k * BW + j) * S] = (j == 0) ? pvt : w[0][j];
    // slide the window: row k leaves, row k+5 enters
#pragma unroll
    for (int sidx = 0; sidx < BKL; ++sidx) {
#pragma unroll
      for (int j = 0; j < BW - 1; ++j) w[sidx][j] = w[sidx + 1][j + 1];
      w[sidx][BW - 1] = make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int j = 0; j < BW; ++j) w[BKL][j] = band_entry(a, p, sb2, sb0, k + 1 + BKL, k + 1 + j);
  }
}

__global__ void __launch_bounds__(64) chan_fd_banded_kernel(const __grid_constant__ BandArgs a) {
  const int64_t inst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= a.ninst) return;
  const int n = a.n;
  const size_t S = (size_t)a.ninst;  // interleave stride
  const PencilCoef p = pencil_coef(a.alpha[inst], a.Re[inst]);
  double2 sigma = a.sigma[inst];
  double2 *U = a.U + inst, *Lm = a.L + inst, *x = a.x + inst, *y = a.y + inst;
  int8_t *piv = a.piv + inst;
  // start vector (same as the dense path)
  for (int i = 0; i < n; ++i) {
    const double sv = sin(3.0 * (i + 1.0) / (n + 1.0)) + 0.37 * cos(11.0 * (i + 1.0) / (n + 1.0));
    x[(size_t)i * S] = make_double2(sv, 0.25 * sv * sv - 0.1);
  }
  double2 om = sigma;
  double dl = 1.0e300;
  int total = 0;
  for (int o = 0; o < a.outer; ++o) {
    band_lu_thread(a, p, sigma, U, Lm, piv, S);
    // ---- inverse iteration
    int it = 0, stalled = 0;
    ConvTrack ct;
    conv_reset(ct);
    for (; it < a.inner; ++it) {
      // y = B x = b2 D2 x + b0 x   (D2 has half bandwidth 2)
      for (int r = 0; r < n; ++r) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int dc = -2; dc <= 2; ++dc) {
          const int c = r + dc;
          if (c >= 0 && c < n) {
            const double d = a.D2[(size_t)c * n + r];
            const double2 xc = x[(size_t)c * S];
            sr = fma(d, xc.x, sr); si = fma(d, xc.y, si);
          }
        }
        y[(size_t)r * S] = cadd2(cmul2(p.b2, make_double2(sr, si)), cmul2(p.b0, x[(size_t)r * S]));
      }
      // forward: L y' = P y, window of 5 right-hand-side values
      double2 bw[BKL + 1];
#pragma unroll
      for (int sidx = 0; sidx <= BKL; ++sidx) bw[sidx] = sidx < n ? y[(size_t)sidx * S] : make_double2(0.0, 0.0);
      for (int k = 0; k < n; ++k) {
        const int pv = piv[(size_t)k * S];
#pragma unroll
        for (int sidx = 1; sidx <= BKL; ++sidx)
          if (pv == sidx) { const double2 t = bw[0]; bw[0] = bw[sidx]; bw[sidx] = t; }
#pragma unroll
        for (int sidx = 1; sidx <= BKL; ++sidx) bw[sidx] = cfnma2(bw[sidx], Lm[((size_t)k * BKL + (sidx - 1)) * S], bw[0]);
        y[(size_t)k * S] = bw[0];
#pragma unroll
        for (int sidx = 0; sidx < BKL; ++sidx) bw[sidx] = bw[sidx + 1];
        bw[BKL] = (k + 1 + BKL < n) ? y[(size_t)(k + 1 + BKL) * S] : make_double2(0.0, 0.0);
      }
      // backward: U y'' = y', window of 8 solved values; accumulate y^H x and y^H y on the way
      double2 xw[BW - 1];
#pragma unroll
      for (int j = 0; j < BW - 1; ++j) xw[j] = make_double2(0.0, 0.0);
      double2 yx = make_double2(0.0, 0.0);
      double yy = 0.0;
      for (int k = n - 1; k >= 0; --k) {
        double2 acc = y[(size_t)k * S];
#pragma unroll
        for (int j = 1; j < BW; ++j) acc = cfnma2(acc, U[((size_t)k * BW + j) * S], xw[j - 1]);
        const double2 v = cdiv2(acc, U[((size_t)k * BW) * S]);
        y[(size_t)k * S] = v;
        const double2 xo = x[(size_t)k * S];
        yx = cadd2(yx, cmulc2(v, xo));
        yy += v.x * v.x + v.y * v.y;
#pragma unroll
        for (int j = BW - 2; j > 0; --j) xw[j] = xw[j - 1];
        xw[0] = v;
      }
      const double2 om_new = cadd2(sigma, make_double2(yx.x / yy, yx.y / yy));
      const double rn = rsqrt(yy);
      for (int i = 0; i < n; ++i) {
        const double2 v = y[(size_t)i * S];
        x[(size_t)i * S] = make_double2(v.x * rn, v.y * rn);
      }
      const double change = hypot(om_new.x - om.x, om_new.y - om.y);
      om = om_new;
      const int cv = conv_push(ct, it, change, a.tol * hypot(om.x, om.y), a.early, &dl);
      if (cv) { ++it; stalled = (cv == 2); break; }
    }
    total += it;
    if (dl <= a.tol * hypot(om.x, om.y) || stalled) break;
    sigma = om;
  }
  a.omega[inst] = om;
  a.delta[inst] = dl;
  a.iters[inst] = total;
  if (a.evec || a.adj) {
    double best = -1.0;
    int bi = 0;
    for (int i = 0; i < n; ++i) {
      const double2 v = x[(size_t)i * S];
      const double m2 = v.x * v.x + v.y * v.y;
      if (m2 > best) { best = m2; bi = i; }
    }
    const double2 esc = cdiv2(make_double2(1.0, 0.0), x[(size_t)bi * S]);
    if (a.evec) {
      double2 *ev = a.evec + (size_t)inst * (n + 2);
      ev[0] = make_double2(0.0, 0.0);
      ev[n + 1] = make_double2(0.0, 0.0);
      for (int i = 0; i < n; ++i) ev[i + 1] = cmul2(esc, x[(size_t)i * S]);
    }
    if (a.adj) {
      // Left eigenvector by inverse iteration on the adjoint pencil, u <- (A - s B)^{-H} B^H u, after a fresh
      // banded factorisation at s = w (1 + 1e-8) (see chan_eig_body).  With E_k = L_k P_k the elimination steps,
      // M = E_0^{-1} ... E_{n-1}^{-1} U, so M^{-H} r = E_0^H ... E_{n-1}^H U^{-H} r.  The iterate lives in y and
      // is written to the output array (used as the second vector) between the phases.
      double2 *av = a.adj + (size_t)inst * (n + 2);
      const double2 sg = make_double2(om.x + 1.0e-8 * hypot(om.x, om.y), om.y);
      band_lu_thread(a, p, sg, U, Lm, piv, S);
      for (int i = 0; i < n; ++i) { const double2 v = x[(size_t)i * S]; av[i + 1] = make_double2(v.x, 0.3 - v.y); }
      for (int it2 = 0; it2 < 4; ++it2) {
        // y = B^H u = conj(b2) D2^T u + conj(b0) u
        for (int r = 0; r < n; ++r) {
          double sr = 0.0, si = 0.0;
#pragma unroll
          for (int dc = -2; dc <= 2; ++dc) {
            const int c = r + dc;
            if (c >= 0 && c < n) {
              const double d = a.D2[(size_t)r * n + c];  // D2(c, r)
              const double2 uc = av[c + 1];
              sr = fma(d, uc.x, sr); si = fma(d, uc.y, si);
            }
          }
          y[(size_t)r * S] = cadd2(cmulc2(p.b2, make_double2(sr, si)), cmulc2(p.b0, av[r + 1]));
        }
        // U^H a = y: a_k = (y_k - sum_j conj(U(k-j, k)) a_{k-j}) / conj(U(k,k))
        for (int k = 0; k < n; ++k) {
          double2 acc = y[(size_t)k * S];
#pragma unroll
          for (int j = 1; j < BW; ++j)
            if (k - j >= 0) acc = csub2(acc, cmulc2(U[((size_t)(k - j) * BW + j) * S], y[(size_t)(k - j) * S]));
          const double2 d = U[((size_t)k * BW) * S];
          y[(size_t)k * S] = cdiv2(acc, make_double2(d.x, -d.y));
        }
        // z = E_0^H ... E_{n-1}^H a: for k = n-1 .. 0: a_k -= sum_s conj(l_{k,s}) a_{k+s}; swap(a_k, a_{k+piv_k})
        for (int k = n - 1; k >= 0; --k) {
          double2 acc = y[(size_t)k * S];
#pragma unroll
          for (int sidx = 1; sidx <= BKL; ++sidx)
            if (k + sidx < n) acc = csub2(acc, cmulc2(Lm[((size_t)k * BKL + (sidx - 1)) * S], y[(size_t)(k + sidx) * S]));
          const int pv = piv[(size_t)k * S];
          if (pv) {
            y[(size_t)k * S] = y[(size_t)(k + pv) * S];
            y[(size_t)(k + pv) * S] = acc;
          } else {
            y[(size_t)k * S] = acc;
          }
        }
        double nn = 0.0;
        for (int i = 0; i < n; ++i) { const double2 v = y[(size_t)i * S]; nn += v.x * v.x + v.y * v.y; }
        const double rn = rsqrt(nn);
        for (int i = 0; i < n; ++i) { const double2 v = y[(size_t)i * S]; av[i + 1] = make_double2(v.x * rn, v.y * rn); }
      }
      if (a.adj_form == OSB_ADJ_PENCIL) {
        double2 d = make_double2(0.0, 0.0);
        for (int i = 0; i < n; ++i) d = cadd2(d, cmulc2(av[i + 1], cmul2(esc, x[(size_t)i * S])));
        const double2 sc = cdiv2(make_double2(1.0, 0.0), make_double2(d.x, -d.y));
        for (int i = 0; i < n; ++i) av[i + 1] = cmul2(sc, av[i + 1]);
      } else {
        double nn = 0.0, vb = -1.0;
        int vi = 0;
        for (int r = 0; r < n; ++r) {
          double sr = 0.0, si = 0.0;
#pragma unroll
          for (int dc = -2; dc <= 2; ++dc) {
            const int c = r + dc;
            if (c >= 0 && c < n) {
              const double d = a.D2[(size_t)r * n + c];
              const double2 uc = av[c + 1];
              sr = fma(d, uc.x, sr); si = fma(d, uc.y, si);
            }
          }
          const double2 v = cadd2(cmulc2(p.b2, make_double2(sr, si)), cmulc2(p.b0, av[r + 1]));
          y[(size_t)r * S] = v;
          const double m2 = v.x * v.x + v.y * v.y;
          nn += m2;
          if (m2 > vb) { vb = m2; vi = r; }
        }
        const double rn = rsqrt(nn);
        const double2 vk = y[(size_t)vi * S];
        const double ak = hypot(vk.x, vk.y);
        const double2 ph = make_double2(rn * vk.x / ak, -rn * vk.y / ak);
        for (int i = 0; i < n; ++i) av[i + 1] = cmul2(ph, y[(size_t)i * S]);
      }
      av[0] = make_double2(0.0, 0.0);
      av[n + 1] = make_double2(0.0, 0.0);
    }
  }
}

// ---------------------------------------------------------------------------
// K4b-w: the same banded elimination on NINE LANES per instance, three instances per warp,
// for sweeps that cannot fill the GPU with a thread per instance (the reference's decks hold
// 21 / 41 alpha values; the throughput variant of config 4 holds 4096).  What counts there is
// the latency of one instance, and at this occupancy that is the number of instructions on
// its serial path:
//   * LU: lane l of a group owns window column j = l (mod 9) -- columns never move between
//     lanes.  The 5 candidates of column k are broadcast, every lane derives pivot and
//     multipliers redundantly and updates its own 4 entries (4 complex FMAs per step instead
//     of 36 in the one-thread kernel); row interchange and window shift are one select.
//     The entering rows are evaluated 32 rows at a time, one diagonal per lane, from the
//     compact band tables into shared memory.
//   * U is stored by COLUMNS (Uc[k][j] = U(k-j, k)) with the reciprocal of the pivot in
//     Uc[k][0], so that the back substitution is column-oriented: its chain per row is one
//     complex multiply + one complex FMA instead of 8 dependent FMAs and a division.
//   * the substitutions are sequential: the leader lanes of the three groups run them side
//     by side out of shared memory, 32 rows per staged chunk; B x, the inner products and
//     the normalisation use all 32 lanes.
// Same algorithm as the one-thread kernel (same pivoting rule, same iteration); the pivots'
// reciprocals and the order of the substitution sums differ in rounding.
// ---------------------------------------------------------------------------
constexpr int BWW = 4;    // warps per CTA
constexpr int BG = 3;     // instances per warp
constexpr int64_t BAND_WARP_MAX_INST = 131072;  // beyond: the one-thread kernel fills the GPU better

template <int BCH>             // rows per staged chunk
struct BandStage {             // per instance, shared memory; [2]: the next chunk arrives (cp.async) while this one is used
  double2 ub[2][BCH * BW];     // LU: entering band rows ([0]); back substitution: chunk of Uc
  double2 lb[2][BCH * BKL];    // forward substitution: chunk of L
  double2 yi[2][BCH], yo[BCH]; // substitution: entering right-hand sides, solved rows
  int pv[2][BCH];
};

__device__ __forceinline__ double2 shfl2(double2 v, int src) {
  return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double2 sel2(bool c, double2 a, double2 b) { return c ? a : b; }

// entry (r, r-4+j) of A - sigma B from the compact band table (x: D2, y: D4); same arithmetic as band_entry
__device__ __forceinline__ double2 band_entry_c(const BandArgs &a, const PencilCoef &p, double2 sb2, double2 sb0, int r, int j) {
  if (r >= a.n || j < 0 || j >= BW) return make_double2(0.0, 0.0);
  const int c = r - BKL + j;
  if (c < 0 || c >= a.n) return make_double2(0.0, 0.0);
  const double2 b = a.band24[(size_t)r * BW + j];
  const double2 w2 = csub2(coef_w2(p, a.u[r]), sb2);
  double2 v = make_double2(fma(w2.x, b.x, b.y), w2.y * b.x);
  if (j == BKL) v = cadd2(v, csub2(coef_w0(p, a.u[r], a.d2u[r]), sb0));
  return v;
}

// Three builds by sweep size: <32 rows per chunk, 1 CTA/SM> has the shortest latency per instance (fewest
// chunk turn-arounds; deck-sized sweeps), <16, 2> and <8, 3> trade some of it for 8 / 12 resident warps per
// SM, which pays when the sweep needs that many CTAs per SM to fit in one wave.
template <int BCH, int MINB>
__global__ void __launch_bounds__(BWW * 32, MINB) chan_fd_banded_warp_kernel(const __grid_constant__ BandArgs a) {
  typedef BandStage<BCH> Stage;
  constexpr int RT = MINB >= 3 ? 2 : 4;  // rows per lane and trip of the vector loops (B x, inner products)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t inst0 = ((int64_t)blockIdx.x * BWW + warp) * BG;
  if (inst0 >= a.ninst) return;  // whole warps leave; no block-wide barrier is used below
  const int n = a.n;
  const int g = min(lane / BW, BG - 1);   // lanes 27..31 shadow group 2 and never store
  const int gl = lane - g * BW;
  const int lead = g * BW;
  const bool lane_ok = gl < BW;
  const bool alive = lane_ok && inst0 + g < a.ninst;
  const int64_t inst = min(inst0 + g, a.ninst - 1);
  Stage *stg = (Stage *)smem_raw + (warp * BG + g);
  Stage *stg0 = (Stage *)smem_raw + warp * BG;
  const size_t nU = (size_t)n * BW, nL = (size_t)n * BKL;
  double2 *Uc = a.U + (size_t)inst * nU, *Lm = a.L + (size_t)inst * nL;
  int32_t *piv = a.piv32 + (size_t)inst * n;
  const PencilCoef p = pencil_coef(a.alpha[inst], a.Re[inst]);
  double2 sigma = a.sigma[inst];
  const int ngrp = (int)min((int64_t)BG, a.ninst - inst0);
  // start vector (same as the dense path); entries of Uc above the first row are never produced
  for (int gp = 0; gp < ngrp; ++gp) {
    double2 *xg = a.x + (size_t)(inst0 + gp) * n;
    for (int i = lane; i < n; i += 32) {
      const double sv = sin(3.0 * (i + 1.0) / (n + 1.0)) + 0.37 * cos(11.0 * (i + 1.0) / (n + 1.0));
      xg[i] = make_double2(sv, 0.25 * sv * sv - 0.1);
    }
  }
  if (alive)
    for (int e = gl; e < 8 * BW; e += BW) {
      const int k = e / BW, j = e % BW;
      if (k < n && j > k) Uc[(size_t)k * BW + j] = make_double2(0.0, 0.0);
    }
  double2 om = sigma;
  double dl = 1.0e300;
  int total = 0;
  bool done = !alive;
  for (int o = 0; o < a.outer; ++o) {
    if (!__any_sync(0xffffffffu, !done)) break;
    const bool act = !done;  // group-uniform
    PROF_T0();
    // ================= banded LU =================
    {
      const double2 sb2 = cmul2(sigma, p.b2), sb0 = cmul2(sigma, p.b0);
      double2 w[BKL + 1];
#pragma unroll
      for (int sidx = 0; sidx <= BKL; ++sidx) w[sidx] = band_entry_c(a, p, sb2, sb0, sidx, lane_ok ? gl - sidx + BKL : -1);
      int src = 0;      // group lane that holds column k
      int mycol = gl;   // the column this lane holds
      for (int c0 = 0; c0 < n; c0 += BCH) {
        // entering rows c0+5 .. c0+36, diagonal gl on lane gl
        __syncwarp();
        if (lane_ok)
          for (int q = 0; q < BCH; ++q) stg->ub[0][q * BW + gl] = band_entry_c(a, p, sb2, sb0, c0 + q + 1 + BKL, gl);
        __syncwarp();
        const int hi = min(BCH, n - c0);
        for (int kk = 0; kk < hi; ++kk) {
          const int k = c0 + kk;
          double2 cc[BKL + 1];
#pragma unroll
          for (int sidx = 0; sidx <= BKL; ++sidx) cc[sidx] = shfl2(w[sidx], lead + src);
          int pv = 0;
          double best = fabs(cc[0].x) + fabs(cc[0].y);
#pragma unroll
          for (int sidx = 1; sidx <= BKL; ++sidx) {
            const double v = fabs(cc[sidx].x) + fabs(cc[sidx].y);
            if (k + sidx < n && v > best) { best = v; pv = sidx; }
          }
          double2 prow = w[0], pc = cc[0];
#pragma unroll
          for (int sidx = 1; sidx <= BKL; ++sidx) {
            prow = sel2(pv == sidx, w[sidx], prow);
            pc = sel2(pv == sidx, cc[sidx], pc);
          }
          // 1 / pivot = conj(p) / |p|^2 ; an exactly singular pivot only happens when sigma hits an
          // eigenvalue to the last bit: inverse iteration wants the huge solve, not a NaN
          double d = fma(pc.x, pc.x, pc.y * pc.y);
          double2 rinv;
          if (d > 1.0e-280 && d < 1.0e280) {
            const double rd = 1.0 / d;
            rinv = make_double2(pc.x * rd, -pc.y * rd);
          } else {
            if (pc.x == 0.0 && pc.y == 0.0) pc = make_double2(1.0e-290, 0.0);
            rinv = cdiv2(make_double2(1.0, 0.0), pc);
          }
          const bool is_src = (gl == src);
          double2 l[BKL + 1];
          const double2 w0 = w[0];
#pragma unroll
          for (int sidx = 1; sidx <= BKL; ++sidx) {
            const double2 tc = sel2(pv == sidx, cc[0], cc[sidx]);   // interchange rows 0 and pv ...
            const double2 t = sel2(pv == sidx, w0, w[sidx]);
            l[sidx] = cmul2(tc, rinv);
            const double2 upd = cfnma2(t, l[sidx], prow);
            w[sidx - 1] = is_src ? make_double2(0.0, 0.0) : upd;      // ... and shift the window up
          }
          if (act && lane_ok) {
            if (is_src) {
              Uc[(size_t)k * BW] = rinv;
              piv[k] = pv;
#pragma unroll
              for (int sidx = 1; sidx <= BKL; ++sidx) Lm[(size_t)k * BKL + (sidx - 1)] = l[sidx];
            } else if (mycol < n) {
              Uc[(size_t)mycol * BW + (mycol - k)] = prow;   // U(k, mycol), stored by column
            }
          }
          int idx = is_src ? BW - 1 : mycol - k - 1;
          idx = max(0, min(idx, BW - 1));
          w[BKL] = stg->ub[0][kk * BW + idx];
          if (is_src) mycol = k + BW;
          src = (src + 1 == BW) ? 0 : src + 1;
        }
      }
    }
    __syncwarp();
    PROF_ADD(1);
    // ================= inverse iteration =================
    bool inner_on = act, stalled = false;
    int it = 0;
    ConvTrack ct;
    conv_reset(ct);
    for (int step = 0; step < a.inner; ++step) {
      if (!__any_sync(0xffffffffu, inner_on)) break;
      // ---- y = B x = b2 D2 x + b0 x (D2: half bandwidth 2), all lanes, one group after the other
      for (int gp = 0; gp < ngrp; ++gp) {
        const bool on = __shfl_sync(0xffffffffu, (int)inner_on, gp * BW) != 0;
        const double2 b2 = shfl2(p.b2, gp * BW), b0 = shfl2(p.b0, gp * BW);
        if (!on) continue;
        const double2 *xg = a.x + (size_t)(inst0 + gp) * n;
        double2 *yg = a.y + (size_t)(inst0 + gp) * n;
        // RT rows per lane and trip (four, or two in the register-limited builds), all loads first: x, y and the band table come from the L2 / DRAM and a trip
        // exposes one round trip (the SASS-level stall attribution had 13 % of the deck-sized kernel waiting here)
        for (int r0 = lane; r0 < n; r0 += 32 * RT) {
          double dd[RT][5];
          double2 xc[RT][5];
#pragma unroll
          for (int t = 0; t < RT; ++t) {
            const int r = min(r0 + 32 * t, n - 1);
#pragma unroll
            for (int dc = -2; dc <= 2; ++dc) {
              dd[t][dc + 2] = a.band24[(size_t)r * BW + BKL + dc].x;
              xc[t][dc + 2] = xg[min(max(r + dc, 0), n - 1)];
            }
          }
#pragma unroll
          for (int t = 0; t < RT; ++t) {
            const int r = r0 + 32 * t;
            if (r < n) {
              double sr = 0.0, si = 0.0;
#pragma unroll
              for (int dc = -2; dc <= 2; ++dc) {
                const int c = r + dc;
                if (c >= 0 && c < n) { sr = fma(dd[t][dc + 2], xc[t][dc + 2].x, sr); si = fma(dd[t][dc + 2], xc[t][dc + 2].y, si); }
              }
              yg[r] = cadd2(cmul2(b2, make_double2(sr, si)), cmul2(b0, xc[t][2]));
            }
          }
        }
      }
      __syncwarp();
      PROF_ADD(5);
      const bool leader = lane_ok && gl == 0 && inner_on;
      const double2 *yme = a.y + (size_t)inst * n;
      // ---- forward: L y' = P y.  Lane s of a group holds the right-hand side of row k+s (s = 0..4).
      // Chunk i+1 of L / pivots / entering rows travels by cp.async while chunk i is eliminated.
      {
        double2 bw = (lane_ok && gl <= BKL && gl < n) ? yme[gl] : make_double2(0.0, 0.0);
        auto stage = [&](int c0, int buf) {
          for (int gp = 0; gp < ngrp; ++gp) {
            const size_t ib = (size_t)(inst0 + gp);
            Stage *sg = stg0 + gp;
#pragma unroll
            for (int q = 0; q < (BCH * BKL + 31) / 32; ++q) {
              const int el = q * 32 + lane, e = c0 * BKL + el;
              if (el < BCH * BKL) cp_async16(&sg->lb[buf][el], a.L + ib * nL + min(e, n * BKL - 1), e < n * BKL);
            }
            if (lane < BCH) {
              const int k = c0 + lane, r = k + 1 + BKL;
              cp_async4(&sg->pv[buf][lane], a.piv32 + ib * n + min(k, n - 1), k < n);
              cp_async16(&sg->yi[buf][lane], a.y + ib * n + min(r, n - 1), r < n);
            }
          }
          cp_async_commit();
        };
        stage(0, 0);
        int buf = 0;
        for (int c0 = 0; c0 < n; c0 += BCH, buf ^= 1) {
          if (c0 + BCH < n) { stage(c0 + BCH, buf ^ 1); cp_async_wait<1>(); }
          else cp_async_wait<0>();
          __syncwarp();
          const int hi = min(BCH, n - c0);
          const int lrow = min(max(gl - 1, 0), BKL - 1);
          for (int kk = 0; kk < hi; ++kk) {
            const int pv = stg->pv[buf][kk];
            const double2 pvv = shfl2(bw, lead + pv), b0 = shfl2(bw, lead);
            const double2 t = sel2(gl == pv, b0, bw);                           // row pv receives the old row 0
            const double2 upd = cfnma2(t, stg->lb[buf][kk * BKL + lrow], pvv);  // rows k+1..k+4 (lanes 1..4)
            if (leader) stg->yo[kk] = pvv;
            // shift: lane s takes the updated row s+1; lane 4 takes the entering row k+5
            const double2 nxt = make_double2(__shfl_down_sync(0xffffffffu, upd.x, 1), __shfl_down_sync(0xffffffffu, upd.y, 1));
            bw = sel2(gl == BKL, stg->yi[buf][kk], nxt);
          }
          __syncwarp();
          for (int gp = 0; gp < ngrp; ++gp)
            if (lane < BCH && c0 + lane < n) a.y[(size_t)(inst0 + gp) * n + c0 + lane] = stg0[gp].yo[lane];
          __syncwarp();
        }
      }
      PROF_ADD(2);
      // ---- backward, column-oriented: x_k = acc_k / U(k,k); acc_{k-j} -= U(k-j,k) x_k.
      // Lane j of a group holds the accumulator of row k-j (j = 0..8)
      {
        double2 acc = (lane_ok && n - 1 - gl >= 0) ? yme[n - 1 - gl] : make_double2(0.0, 0.0);
        const int ucol = min(gl, BW - 1);
        auto stage = [&](int c0, int buf) {
          for (int gp = 0; gp < ngrp; ++gp) {
            const size_t ib = (size_t)(inst0 + gp);
            Stage *sg = stg0 + gp;
#pragma unroll
            for (int q = 0; q < (BCH * BW + 31) / 32; ++q) {
              const int el = q * 32 + lane, e = c0 * BW + el;
              if (el < BCH * BW) cp_async16(&sg->ub[buf][el], a.U + ib * nU + min(e, n * BW - 1), e < n * BW);
            }
            if (lane < BCH) {
              const int r = c0 + lane - BW;
              cp_async16(&sg->yi[buf][lane], a.y + ib * n + max(r, 0), r >= 0);
            }
          }
          cp_async_commit();
        };
        const int ctop = ((n - 1) / BCH) * BCH;
        stage(ctop, 0);
        int buf = 0;
        for (int c0 = ctop; c0 >= 0; c0 -= BCH, buf ^= 1) {
          if (c0 - BCH >= 0) { stage(c0 - BCH, buf ^ 1); cp_async_wait<1>(); }
          else cp_async_wait<0>();
          __syncwarp();
          const int hi = min(BCH, n - c0);
          for (int kk = hi - 1; kk >= 0; --kk) {
            const double2 um = stg->ub[buf][kk * BW + ucol];     // lane 0: 1/U(k,k); lane j: U(k-j, k)
            const double2 v = shfl2(cmul2(acc, um), lead);       // x_k
            if (leader) stg->yo[kk] = v;
            const double2 upd = cfnma2(acc, um, v);
            const double2 nxt = make_double2(__shfl_down_sync(0xffffffffu, upd.x, 1), __shfl_down_sync(0xffffffffu, upd.y, 1));
            acc = sel2(gl == BW - 1, stg->yi[buf][kk], nxt);     // lane 8 takes the entering row k-9
          }
          __syncwarp();
          for (int gp = 0; gp < ngrp; ++gp)
            if (lane < BCH && c0 + lane < n) a.y[(size_t)(inst0 + gp) * n + c0 + lane] = stg0[gp].yo[lane];
          __syncwarp();
        }
      }
      PROF_ADD(6);
      // ---- omega = sigma + y^H x / y^H y ; x = y / |y|
      double myyx = 0.0, myyxi = 0.0, myyy = 1.0;
      for (int gp = 0; gp < ngrp; ++gp) {
        const bool on = __shfl_sync(0xffffffffu, (int)inner_on, gp * BW) != 0;
        if (!on) continue;
        double2 *xg = a.x + (size_t)(inst0 + gp) * n;
        const double2 *yg = a.y + (size_t)(inst0 + gp) * n;
        double yxr = 0.0, yxi = 0.0, yy = 0.0;
        // RT rows per lane and trip, loads first; the sums keep their order i = lane, lane + 32, ...
        for (int i0 = lane; i0 < n; i0 += 32 * RT) {
          double2 v[RT], xo[RT];
#pragma unroll
          for (int t = 0; t < RT; ++t) { const int i = min(i0 + 32 * t, n - 1); v[t] = yg[i]; xo[t] = xg[i]; }
#pragma unroll
          for (int t = 0; t < RT; ++t)
            if (i0 + 32 * t < n) {
              const double2 tt = cmulc2(v[t], xo[t]);
              yxr += tt.x; yxi += tt.y;
              yy += v[t].x * v[t].x + v[t].y * v[t].y;
            }
        }
        yxr = warp_sum(yxr); yxi = warp_sum(yxi); yy = warp_sum(yy);
        const double rn = rsqrt(yy);
        for (int i0 = lane; i0 < n; i0 += 32 * RT) {
          double2 v[RT];
#pragma unroll
          for (int t = 0; t < RT; ++t) v[t] = yg[min(i0 + 32 * t, n - 1)];
#pragma unroll
          for (int t = 0; t < RT; ++t)
            if (i0 + 32 * t < n) xg[i0 + 32 * t] = make_double2(v[t].x * rn, v[t].y * rn);
        }
        if (g == gp) { myyx = yxr; myyxi = yxi; myyy = yy; }
      }
      __syncwarp();
      if (inner_on) {
        const double2 om_new = cadd2(sigma, make_double2(myyx / myyy, myyxi / myyy));
        const double change = hypot(om_new.x - om.x, om_new.y - om.y);
        om = om_new;
        it = step + 1;
        const int cv = conv_push(ct, step, change, a.tol * hypot(om.x, om.y), a.early, &dl);
        if (cv) inner_on = false;
        if (cv == 2) stalled = true;
      }
      PROF_ADD(0);
    }
    if (act) {
      total += it;
      if (dl <= a.tol * hypot(om.x, om.y) || stalled) done = true;
      else sigma = om;  // re-shift (Rayleigh-quotient style) and refactorise
    }
  }
  if (alive && gl == 0) {
    a.omega[inst] = om;
    a.delta[inst] = dl;
    a.iters[inst] = total;
  }
  if (a.evec) {
    // normalise by the component of largest modulus, lowest index on ties (orrfdchan.f:849-857)
    for (int gp = 0; gp < ngrp; ++gp) {
      const double2 *xg = a.x + (size_t)(inst0 + gp) * n;
      double best = -1.0;
      int bi = 0;
      for (int i = lane; i < n; i += 32) {
        const double2 v = xg[i];
        const double m2 = v.x * v.x + v.y * v.y;
        if (m2 > best) { best = m2; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      const double2 esc = cdiv2(make_double2(1.0, 0.0), xg[bi]);
      double2 *ev = a.evec + (size_t)(inst0 + gp) * (n + 2);
      if (lane == 0) { ev[0] = make_double2(0.0, 0.0); ev[n + 1] = make_double2(0.0, 0.0); }
      for (int i = lane; i < n; i += 32) ev[i + 1] = cmul2(esc, xg[i]);
    }
  }
}

size_t chan_band_work_bytes(int n, int64_t ninst) {
  return (size_t)ninst * (size_t)n * (sizeof(double2) * (BW + BKL + 2) + 1 + 4) + 512;
}

cudaError_t launch_chan_fd_banded(int n, const double *D2, const double *D4, const double *u, const double *d2u,
                                  const double2 *band24, void *work, int64_t ninst, const double2 *alpha,
                                  const double *Re, const double2 *sigma, int outer, int inner, double tol,
                                  double2 *omega, double *delta, int32_t *iters, double2 *evec, double2 *adj, int adj_form,
                                  cudaStream_t st) {
  BandArgs a;
  a.n = n; a.D2 = D2; a.D4 = D4; a.u = u; a.d2u = d2u; a.band24 = band24; a.ninst = ninst; a.alpha = alpha; a.Re = Re;
  a.sigma = sigma; a.outer = outer; a.inner = inner; a.tol = tol; a.omega = omega; a.delta = delta; a.iters = iters;
  a.evec = evec; a.adj = adj; a.adj_form = adj_form;
  a.early = getenv("OSB_NO_EARLY_STOP") ? 0 : 1;
  double2 *w = (double2 *)work;
  const size_t N1 = (size_t)ninst * n;
  a.U = w; a.L = w + N1 * BW; a.x = w + N1 * (BW + BKL); a.y = w + N1 * (BW + BKL + 1);
  a.piv = (int8_t *)(w + N1 * (BW + BKL + 2));
  a.piv32 = (int32_t *)(((uintptr_t)(a.piv + N1) + 255) & ~(uintptr_t)255);
  // nine lanes per instance while that still leaves the GPU room (the one-thread kernel needs
  // ~10^5 instances to fill it); OSB_FD_THREAD=1 forces the one-thread kernel
  // (the adjoint eigenvector is implemented in the one-thread kernel only: it is asked for by the single-alpha
  // listings, never by sweeps)
  if (band24 && ninst <= BAND_WARP_MAX_INST && !adj && !getenv("OSB_FD_THREAD")) {
    const int64_t per_cta = (int64_t)BG * BWW;
    const unsigned grid = (unsigned)((ninst + per_cta - 1) / per_cta);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e;
    if ((int64_t)grid <= (int64_t)sms) {
      const size_t smem = sizeof(BandStage<32>) * BG * BWW;
      e = cudaFuncSetAttribute(chan_fd_banded_warp_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      chan_fd_banded_warp_kernel<32, 1><<<grid, BWW * 32, smem, st>>>(a);
    } else if ((int64_t)grid <= 2 * (int64_t)sms) {
      const size_t smem = sizeof(BandStage<16>) * BG * BWW;
      e = cudaFuncSetAttribute(chan_fd_banded_warp_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      chan_fd_banded_warp_kernel<16, 2><<<grid, BWW * 32, smem, st>>>(a);
    } else if ((int64_t)grid <= 3 * (int64_t)sms || getenv("OSB_FD_NO4") ||
               // beyond one wave of the 3-per-SM build: the 4-per-SM build (128 registers, a few spills, ~12 % slower
               // per warp) wins when it saves a wave
               ((grid + 3 * sms - 1) / (3 * sms)) * 100 <= ((grid + 4 * sms - 1) / (4 * sms)) * 112) {
      const size_t smem = sizeof(BandStage<8>) * BG * BWW;
      e = cudaFuncSetAttribute(chan_fd_banded_warp_kernel<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      chan_fd_banded_warp_kernel<8, 3><<<grid, BWW * 32, smem, st>>>(a);
    } else {
      const size_t smem = sizeof(BandStage<8>) * BG * BWW;
      e = cudaFuncSetAttribute(chan_fd_banded_warp_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      chan_fd_banded_warp_kernel<8, 4><<<grid, BWW * 32, smem, st>>>(a);
    }
    return cudaGetLastError();
  }
  const int threads = 32;  // few, long-running threads: spread them over all SMs
  const unsigned grid = (unsigned)((ninst + threads - 1) / threads);
  chan_fd_banded_kernel<<<grid, threads, 0, st>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Newton on alpha to Im(omega) = 0   (orrncbl.f:109-116, 962-981), one CTA per Re
// ---------------------------------------------------------------------------
struct NeutralArgs {
  int n;
  const double *D2, *D4, *u, *d2u;
  double2 *work;
  int64_t ninst;
  const double *Re;       // [ninst]
  const double *alpha0;   // [ninst] starting alpha_r
  const double2 *sigma0;  // [ninst] eigenvalue at alpha0 (from the scan)
  double sfac, tol;
  int maxit;
  int early;              // allow the one-solve-early stop of conv_push
  double *alpha_out;      // [ninst] nalpha after the last step
  double2 *omega_out;     // [ninst]
  int32_t *steps;         // [ninst]
  int32_t *status;        // [ninst]
  double *trace;          // optional [ninst][maxit][7]
};

__global__ void __launch_bounds__(CH_THREADS) chan_neutral_kernel(const __grid_constant__ NeutralArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int n = a.n;
  ChanShared s = carve(smem, n);
  double2 *M = chan_matrix_in_smem(n) ? (double2 *)(smem + chan_fixed_smem(n)) : a.work + (size_t)blockIdx.x * n * n;
  for (int64_t inst = blockIdx.x; inst < a.ninst; inst += gridDim.x) {
    const double Re = a.Re[inst];
    double nalpha = a.alpha0[inst];
    double2 sigma = a.sigma0[inst];
    double2 ev = make_double2(1.0, 1.0);
    int it = 0, st = OSB_ST_MAXCOUNT;
    while (fabs(ev.y) > a.tol && it < a.maxit) {
      const PencilCoef p = pencil_coef(make_double2(nalpha, 0.0), Re);
      // right eigenpair, up to 3 factorisations
      start_vector(s.x, n);
      double dl = 1.0e300;
      for (int o = 0; o < 3; ++o) {
        assemble_shifted(M, n, a.D2, a.D4, a.u, a.d2u, p, sigma);
        cta_lu(M, n, s.ipiv, s.panel, s.red);
        int k = 0;
        ev = inverse_iterate(M, n, a.D2, p, sigma, s, 8, 1.0e-14, a.early, &dl, &k);
        if (dl <= 1.0e-14 * hypot(ev.x, ev.y)) break;
        sigma = ev;
      }
      // left vector of inv(B)A: vl = B^H u, (A - sigma B)^H u' = vl  (same LU)
      for (int i = threadIdx.x; i < n; i += blockDim.x) s.t[i] = make_double2(s.x[i].x, -s.x[i].y + 0.3);
      __syncthreads();
      for (int k = 0; k < 6; ++k) {
        cta_solve_h(M, n, s.ipiv, s.t);                     // u = M^{-H} vl
        cta_d2_matvec(a.D2, n, s.t, s.y, true);             // D2^T u
        double nn = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          // vl = conj(b2) D2^T u + conj(b0) u
          const double2 v = cadd2(cmulc2(p.b2, s.y[i]), cmulc2(p.b0, s.t[i]));
          s.t[i] = v;
          nn += v.x * v.x + v.y * v.y;
        }
        const double2 tot = block_sum(make_double2(nn, 0.0), s.red);
        const double rn = rsqrt(tot.x);
        for (int i = threadIdx.x; i < n; i += blockDim.x) s.t[i] = make_double2(s.t[i].x * rn, s.t[i].y * rn);
        __syncthreads();
      }
      // dw = vl^H (dB w - dA) x ; dalp = vl^H B x        (orrncbl.f:962-976)
      cta_d2_matvec(a.D2, n, s.x, s.y, false);              // D2 x
      double2 dw = make_double2(0.0, 0.0), dalp = make_double2(0.0, 0.0);
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double2 d2x = s.y[i], xi = s.x[i];
        const double2 tdw = csub2(csub2(cmul2(cmul2(p.db0, ev), xi), cmul2(coef_d2(p, a.u[i]), d2x)),
                                  cmul2(coef_d0(p, a.u[i], a.d2u[i]), xi));
        const double2 tbx = cadd2(cmul2(p.b2, d2x), cmul2(p.b0, xi));
        dw = cadd2(dw, cmulc2(s.t[i], tdw));
        dalp = cadd2(dalp, cmulc2(s.t[i], tbx));
      }
      dw = block_sum(dw, s.red);
      dalp = block_sum(dalp, s.red);
      const double2 q = cdiv2(dw, dalp);
      const double2 dwda = make_double2(-q.x, -q.y);
      const double dalpha = -ev.y / dwda.y;
      const double alpha_now = nalpha;
      nalpha = alpha_now + a.sfac * dalpha;
      const double nw = ev.x + a.sfac * dalpha * dwda.x;
      if (a.trace && threadIdx.x == 0) {
        double *r = a.trace + ((size_t)inst * a.maxit + it) * 7;
        r[0] = nalpha; r[1] = ev.x; r[2] = ev.y; r[3] = dwda.x; r[4] = dwda.y; r[5] = dalpha; r[6] = nw;
      }
      sigma = make_double2(nw, ev.y);  // next shift: predicted real part (orrncbl.f:907-914 selects near nw)
      ++it;
      if (!(isfinite(nalpha) && isfinite(ev.x) && isfinite(ev.y))) { st = OSB_ST_NONFINITE; break; }
      __syncthreads();
    }
    if (fabs(ev.y) <= a.tol) st = OSB_ST_CONVERGED;
    if (threadIdx.x == 0) {
      a.alpha_out[inst] = nalpha;
      a.omega_out[inst] = ev;
      a.steps[inst] = it;
      a.status[inst] = st;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// K4s: small operators (n <= 64: chan.inp, nc.inp), ONE WARP per instance.  The CTA-wide
// kernels above spend most of their time in barriers when the matrix is only 63 x 63;
// here a warp keeps its matrix in shared memory (lane l owns rows l and l+32), pivots with
// shuffles and needs only __syncwarp.  Same algorithm (partial-pivoting LU, shift-invert
// iteration with re-shift, Newton on alpha), same results to rounding.
// ---------------------------------------------------------------------------
constexpr int WN_MAX = 64;

struct WarpShared {
  double2 *M, *x, *y, *t;
  int *ipiv;
};
__host__ __device__ inline size_t warp_smem_bytes(int n) {
  return sizeof(double2) * ((size_t)n * n + 3 * (size_t)n) + sizeof(int) * (size_t)(lu_perm_offset(n) + n);
}
__device__ WarpShared wcarve(unsigned char *smem, int n) {
  WarpShared s;
  s.M = (double2 *)smem;
  s.x = s.M + (size_t)n * n;
  s.y = s.x + n;
  s.t = s.y + n;
  s.ipiv = (int *)(s.t + n);
  return s;
}
__device__ __forceinline__ double2 wsum(double2 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

template <bool TWO>  // TWO: n > 32, a lane owns rows lane and lane + 32; otherwise one row and none of the second row's work
__device__ void w_assemble(double2 *M, int n, const double *__restrict__ D2, const double *__restrict__ D4,
                           const double *__restrict__ u, const double *__restrict__ d2u, const PencilCoef &p,
                           double2 sigma) {
  const int lane = threadIdx.x & 31;
  const double2 sb2 = cmul2(sigma, p.b2), sb0 = cmul2(sigma, p.b0);
  // rows lane and lane + 32 together: twice the table loads in flight per trip (the kernel runs one warp per
  // scheduler, so nothing else hides their latency)
  const int i0 = min(lane, n - 1), i1 = min(lane + 32, n - 1);
  const bool h0 = lane < n, h1 = TWO && lane + 32 < n;
  const double2 w2a = csub2(coef_w2(p, u[i0]), sb2), w2b = csub2(coef_w2(p, u[i1]), sb2);
  const double2 w0a = csub2(coef_w0(p, u[i0], d2u[i0]), sb0), w0b = csub2(coef_w0(p, u[i1], d2u[i1]), sb0);
  for (int j = 0; j < n; j += 8) {  // eight columns per trip, table loads first; the last trip is masked
    double d2a[8], d4a[8], d2b[8], d4b[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const size_t c = (size_t)min(j + t, n - 1) * n;
      d2a[t] = __ldg(D2 + c + i0); d4a[t] = __ldg(D4 + c + i0);
      if (TWO) { d2b[t] = __ldg(D2 + c + i1); d4b[t] = __ldg(D4 + c + i1); }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int jt = j + t;
      if (jt < n) {
        double2 va = make_double2(fma(w2a.x, d2a[t], d4a[t]), w2a.y * d2a[t]);
        if (i0 == jt) va = cadd2(va, w0a);
        if (h0) M[(size_t)jt * n + i0] = va;
        if (TWO) {
          double2 vb = make_double2(fma(w2b.x, d2b[t], d4b[t]), w2b.y * d2b[t]);
          if (i1 == jt) vb = cadd2(vb, w0b);
          if (h1) M[(size_t)jt * n + i1] = vb;
        }
      }
    }
  }
  __syncwarp();
}

// Partial-pivoting LU of the shared-memory matrix by one warp; lane l owns rows l and l + 32.  The warp is alone on its
// scheduler and three warps share the SM's shared-memory pipe, so two things count (SASS-level stall attribution of
// round 2, tools/capture_source.sh): the latency of the warp's own instruction stream, and shared-memory traffic --
// a rank-1 update reads and writes the trailing matrix once per pivot (16 LDS.128 + 16 STS.128 against 64 DFMAs per
// trip: the three warps' 384 cycles of shared-memory pipe per trip against 144 cycles of FP64 pipe each).  Hence
// FOUR pivots per visit of the trailing matrix:
//   * the four panel columns are factorised column by column as before (pivot search by three redux.sync reductions,
//     branch-free over the lane's two entries; interchange over ALL columns; multipliers; rank-1 update of the panel's
//     remaining <= 3 columns only);
//   * the pivot rows' entries in the trailing columns (U12 = L11^{-1} A12) are finished with a lane per column;
//   * every row below the panel then receives its four updates in registers: one load and one store per entry and
//     four pivots, eight complex FMAs per loaded pair, all loads of a trip first (the compiler cannot move loads above
//     the stores of the previous columns: it cannot prove that pivot-row entries and updated rows do not alias).
// Every entry receives the same updates with the same operands in the same order k = 0, 1, ... as in the unblocked
// right-looking elimination: same pivots, same bits.
template <bool TWO>
__device__ void w_lu(double2 *M, int n, int *ipiv) {
  constexpr int NB = 4;
  const int lane = threadIdx.x & 31;
  int *perm = lu_perm(ipiv, n);
  for (int i = lane; i < n; i += 32) perm[i] = i;
  __syncwarp();
  const int r0 = lane, r1 = lane + 32;
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int ke = min(k0 + NB, n);  // panel columns k0 .. ke - 1
    for (int k = k0; k < ke; ++k) {
      double2 *ck = M + (size_t)k * n;
      // pivot search on the bit patterns (non-negative doubles order like their bits; ties to the lower row as IZAMAX)
      long long best = -1;
      int bi = k;
      {
        const bool h0 = r0 >= k && r0 < n, h1 = TWO && r1 >= k && r1 < n;
        const double2 v0 = ck[h0 ? r0 : k], v1 = TWO ? ck[h1 ? r1 : k] : v0;
        const long long a0 = h0 ? (__double_as_longlong(fabs(v0.x) + fabs(v0.y)) & 0x7fffffffffffffffll) : -1ll;
        const long long a1 = h1 ? (__double_as_longlong(fabs(v1.x) + fabs(v1.y)) & 0x7fffffffffffffffll) : -1ll;
        if (a0 > best) { best = a0; bi = r0; }
        if (a1 > best) { best = a1; bi = r1; }
      }
      {
        // warp maximum of the 64-bit key, lowest row on ties, by three integer reductions (redux.sync) instead of five
        // rounds of three shuffles: high words, then low words among the lanes that hold the maximal high word, then
        // the smallest row among the lanes that hold the maximum.  key + 1 maps "no candidate" (-1) to 0.
        const unsigned long long key = (unsigned long long)(best + 1);
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
        bi = (int)__reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? (unsigned)bi : 0xffffffffu);
      }
      if (lane == 0) {
        ipiv[k] = bi;
        const int ta = perm[k];
        perm[k] = perm[bi];
        perm[bi] = ta;
      }
      if (bi != k) {  // interchange rows k and bi: columns lane and lane + 32
        const int j0 = min(r0, n - 1), j1 = min(r1, n - 1);
        double2 *c0 = M + (size_t)j0 * n, *c1 = M + (size_t)j1 * n;
        const double2 t0 = c0[k], b0 = c0[bi];
        double2 t1 = t0, b1 = b0;
        if (TWO) { t1 = c1[k]; b1 = c1[bi]; }
        if (r0 < n) { c0[k] = b0; c0[bi] = t0; }
        if (TWO && r1 < n) { c1[k] = b1; c1[bi] = t1; }
      }
      __syncwarp();
      double2 piv = ck[k];
      if (piv.x == 0.0 && piv.y == 0.0) piv = make_double2(1.0e-290, 0.0);
      const double pinv = 1.0 / fma(piv.x, piv.x, piv.y * piv.y);  // one division in the column's serial chain
      const double2 rinv = make_double2(piv.x * pinv, -piv.y * pinv);
      const bool a0 = r0 > k && r0 < n, a1 = TWO && r1 > k && r1 < n;
      const int q0 = a0 ? r0 : k, q1 = a1 ? r1 : k;  // rows to read where a lane has nothing to update (never stored)
      double2 l0 = make_double2(0.0, 0.0), l1 = l0;
      if (a0) { l0 = cmul2(ck[r0], rinv); ck[r0] = l0; }
      if (a1) { l1 = cmul2(ck[r1], rinv); ck[r1] = l1; }
      if (k + 1 < ke) {  // the panel's remaining columns (<= 3): loads first
        double2 uu[NB - 1], ma[NB - 1], mb[NB - 1];
#pragma unroll
        for (int t = 0; t < NB - 1; ++t) {
          const double2 *c = M + (size_t)min(k + 1 + t, ke - 1) * n;
          uu[t] = c[k]; ma[t] = c[q0]; mb[t] = TWO ? c[q1] : ma[t];
        }
#pragma unroll
        for (int t = 0; t < NB - 1; ++t)
          if (k + 1 + t < ke) {
            double2 *c = M + (size_t)(k + 1 + t) * n;
            if (a0) c[r0] = cfnma2(ma[t], l0, uu[t]);
            if (a1) c[r1] = cfnma2(mb[t], l1, uu[t]);
          }
      }
      __syncwarp();
    }
    if (ke >= n) break;  // (a ragged last panel has no trailing columns)
    // ---- pivot rows k0+1 .. k0+3 in the trailing columns: U12 = L11^{-1} A12, a lane per column
    for (int j = ke + lane; j < n; j += 32) {
      double2 *c = M + (size_t)j * n;
      double2 u[NB];
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) u[cc] = c[k0 + cc];
#pragma unroll
      for (int cc = 1; cc < NB; ++cc) {
#pragma unroll
        for (int d = 0; d < cc; ++d) u[cc] = cfnma2(u[cc], M[(size_t)(k0 + d) * n + k0 + cc], u[d]);
        c[k0 + cc] = u[cc];
      }
    }
    __syncwarp();
    // ---- rows below the panel: four updates per visit
    const bool b0 = r0 >= ke && r0 < n, b1 = TWO && r1 >= ke && r1 < n;
    const int q0 = b0 ? r0 : k0, q1 = b1 ? r1 : k0;
    double2 l0[NB], l1[NB];
#pragma unroll
    for (int cc = 0; cc < NB; ++cc) {
      l0[cc] = M[(size_t)(k0 + cc) * n + q0];
      l1[cc] = TWO ? M[(size_t)(k0 + cc) * n + q1] : l0[cc];
    }
    if (!TWO || ke > 31) {  // one row per lane left: r0 (n <= 32) or r1
      const bool bb = TWO ? b1 : b0;
      const int qq = TWO ? q1 : q0, rr = TWO ? r1 : r0;
      for (int j = ke; j < n; j += 4) {
        double2 uu[4][NB], mm[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const double2 *c = M + (size_t)min(j + t, n - 1) * n;
#pragma unroll
          for (int cc = 0; cc < NB; ++cc) uu[t][cc] = c[k0 + cc];
          mm[t] = c[qq];
        }
#pragma unroll
        for (int cc = 0; cc < NB; ++cc)
#pragma unroll
          for (int t = 0; t < 4; ++t) mm[t] = cfnma2(mm[t], TWO ? l1[cc] : l0[cc], uu[t][cc]);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (bb && j + t < n) M[(size_t)(j + t) * n + rr] = mm[t];
      }
    } else {
      for (int j = ke; j < n; j += 4) {
        double2 uu[4][NB], ma[4], mb[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const double2 *c = M + (size_t)min(j + t, n - 1) * n;
#pragma unroll
          for (int cc = 0; cc < NB; ++cc) uu[t][cc] = c[k0 + cc];
          ma[t] = c[q0]; mb[t] = c[q1];
        }
#pragma unroll
        for (int cc = 0; cc < NB; ++cc)
#pragma unroll
          for (int t = 0; t < 4; ++t) { ma[t] = cfnma2(ma[t], l0[cc], uu[t][cc]); mb[t] = cfnma2(mb[t], l1[cc], uu[t][cc]); }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (b0 && j + t < n) M[(size_t)(j + t) * n + r0] = ma[t];
          if (b1 && j + t < n) M[(size_t)(j + t) * n + r1] = mb[t];
        }
      }
    }
    __syncwarp();
  }
}

// x <- M^{-1} x, n <= 64: a lane keeps rows lane and lane + 32 of x in registers and the solved value of a column
// reaches the others by shuffle -- no shared-memory round trip and no warp barrier per column; the reciprocals of the
// pivots are formed up front, two per lane, so the back substitution's chain holds no division
template <bool TWO>
__device__ void w_solve(const double2 *M, int n, const int *ipiv, double2 *x) {
  const int lane = threadIdx.x & 31;
  const int *perm = lu_perm(ipiv, n);
  const int r0 = lane, r1 = lane + 32;
  double2 x0 = r0 < n ? x[perm[r0]] : make_double2(0.0, 0.0), x1 = r1 < n ? x[perm[r1]] : make_double2(0.0, 0.0);  // P x
  __syncwarp();
  double2 dinv0 = make_double2(0.0, 0.0), dinv1 = dinv0;
  {
    double2 d = r0 < n ? M[(size_t)r0 * n + r0] : make_double2(1.0, 0.0);
    if (d.x == 0.0 && d.y == 0.0) d = make_double2(1.0e-290, 0.0);
    double sc = 1.0 / fma(d.x, d.x, d.y * d.y);
    dinv0 = make_double2(d.x * sc, -d.y * sc);
    d = r1 < n ? M[(size_t)r1 * n + r1] : make_double2(1.0, 0.0);
    if (d.x == 0.0 && d.y == 0.0) d = make_double2(1.0e-290, 0.0);
    sc = 1.0 / fma(d.x, d.x, d.y * d.y);
    dinv1 = make_double2(d.x * sc, -d.y * sc);
  }
  // The factor entries of four columns are loaded before the four dependent steps that use them (addresses clamped
  // where a lane has no row): the chain then holds only the shuffle and the complex FMA.
  const int c0 = min(r0, n - 1), c1 = min(r1, n - 1);
  {  // L y = P b
    int k = 0;
    for (; k + 3 < n - 1; k += 4) {
      double2 m0[4], m1[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) { m0[t] = M[(size_t)(k + t) * n + c0]; m1[t] = TWO ? M[(size_t)(k + t) * n + c1] : m0[t]; }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int kk = k + t;
        const double2 src = kk < 32 ? x0 : x1;
        const double2 xk = make_double2(__shfl_sync(0xffffffffu, src.x, kk & 31), __shfl_sync(0xffffffffu, src.y, kk & 31));
        if (r0 > kk && r0 < n) x0 = cfnma2(x0, m0[t], xk);
        if (TWO && r1 > kk && r1 < n) x1 = cfnma2(x1, m1[t], xk);
      }
    }
    for (; k < n - 1; ++k) {
      const double2 src = k < 32 ? x0 : x1;
      const double2 xk = make_double2(__shfl_sync(0xffffffffu, src.x, k & 31), __shfl_sync(0xffffffffu, src.y, k & 31));
      if (r0 > k && r0 < n) x0 = cfnma2(x0, M[(size_t)k * n + r0], xk);
      if (TWO && r1 > k && r1 < n) x1 = cfnma2(x1, M[(size_t)k * n + r1], xk);
    }
  }
  {  // U x = y
    int k = n - 1;
    for (; k - 3 >= 0; k -= 4) {
      double2 m0[4], m1[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) { m0[t] = M[(size_t)(k - t) * n + c0]; m1[t] = TWO ? M[(size_t)(k - t) * n + c1] : m0[t]; }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int kk = k - t;
        if (kk == r0) x0 = cmul2(x0, dinv0);
        if (TWO && kk == r1) x1 = cmul2(x1, dinv1);
        const double2 src = kk < 32 ? x0 : x1;
        const double2 xk = make_double2(__shfl_sync(0xffffffffu, src.x, kk & 31), __shfl_sync(0xffffffffu, src.y, kk & 31));
        if (r0 < kk) x0 = cfnma2(x0, m0[t], xk);
        if (TWO && r1 < kk) x1 = cfnma2(x1, m1[t], xk);
      }
    }
    for (; k >= 0; --k) {
      if (k == r0) x0 = cmul2(x0, dinv0);
      if (TWO && k == r1) x1 = cmul2(x1, dinv1);
      const double2 src = k < 32 ? x0 : x1;
      const double2 xk = make_double2(__shfl_sync(0xffffffffu, src.x, k & 31), __shfl_sync(0xffffffffu, src.y, k & 31));
      if (r0 < k) x0 = cfnma2(x0, M[(size_t)k * n + r0], xk);
      if (TWO && r1 < k) x1 = cfnma2(x1, M[(size_t)k * n + r1], xk);
    }
  }
  if (r0 < n) x[r0] = x0;
  if (r1 < n) x[r1] = x1;
  __syncwarp();
}

template <bool TWO>
__device__ void w_solve_h(const double2 *M, int n, const int *ipiv, double2 *x) {
  // x <- M^{-H} x,  M^H = U^H L^H P.  Column-oriented like w_solve: x lives in registers (rows lane, lane + 32), the
  // solved component is broadcast by shuffle, the pivots' reciprocals are formed up front.
  const int lane = threadIdx.x & 31;
  const int r0 = lane, r1 = lane + 32;
  double2 x0 = r0 < n ? x[r0] : make_double2(0.0, 0.0), x1 = r1 < n ? x[r1] : make_double2(0.0, 0.0);
  double2 dinv0, dinv1;  // 1 / conj(U(r, r))
  {
    double2 d = r0 < n ? M[(size_t)r0 * n + r0] : make_double2(1.0, 0.0);
    if (d.x == 0.0 && d.y == 0.0) d = make_double2(1.0e-290, 0.0);
    double sc = 1.0 / fma(d.x, d.x, d.y * d.y);
    dinv0 = make_double2(d.x * sc, d.y * sc);
    d = r1 < n ? M[(size_t)r1 * n + r1] : make_double2(1.0, 0.0);
    if (d.x == 0.0 && d.y == 0.0) d = make_double2(1.0e-290, 0.0);
    sc = 1.0 / fma(d.x, d.x, d.y * d.y);
    dinv1 = make_double2(d.x * sc, d.y * sc);
  }
  // as in w_solve: the factor entries of four steps are loaded ahead of the dependent chain.  Row r of U / L is read
  // along the row here (stride n), clamped where a lane has no row.
  const int c0 = min(r0, n - 1), c1 = min(r1, n - 1);
  {  // U^H a = r : lower triangular, (U^H)(j, k) = conj(U(k, j)) = conj(M[j n + k])
    int k = 0;
    for (; k + 3 < n; k += 4) {
      double2 m0[4], m1[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) { m0[t] = M[(size_t)c0 * n + k + t]; m1[t] = TWO ? M[(size_t)c1 * n + k + t] : m0[t]; }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int kk = k + t;
        if (kk == r0) x0 = cmul2(x0, dinv0);
        if (TWO && kk == r1) x1 = cmul2(x1, dinv1);
        const double2 src = kk < 32 ? x0 : x1;
        const double2 ak = make_double2(__shfl_sync(0xffffffffu, src.x, kk & 31), __shfl_sync(0xffffffffu, src.y, kk & 31));
        if (r0 > kk && r0 < n) x0 = cfnma2(x0, make_double2(m0[t].x, -m0[t].y), ak);
        if (TWO && r1 > kk && r1 < n) x1 = cfnma2(x1, make_double2(m1[t].x, -m1[t].y), ak);
      }
    }
    for (; k < n; ++k) {
      if (k == r0) x0 = cmul2(x0, dinv0);
      if (TWO && k == r1) x1 = cmul2(x1, dinv1);
      const double2 src = k < 32 ? x0 : x1;
      const double2 ak = make_double2(__shfl_sync(0xffffffffu, src.x, k & 31), __shfl_sync(0xffffffffu, src.y, k & 31));
      if (r0 > k && r0 < n) { const double2 u = M[(size_t)r0 * n + k]; x0 = cfnma2(x0, make_double2(u.x, -u.y), ak); }
      if (TWO && r1 > k && r1 < n) { const double2 u = M[(size_t)r1 * n + k]; x1 = cfnma2(x1, make_double2(u.x, -u.y), ak); }
    }
  }
  {  // L^H b = a : unit upper, (L^H)(j, k) = conj(L(k, j)) = conj(M[j n + k])
    int k = n - 1;
    for (; k - 3 > 0; k -= 4) {
      double2 m0[4], m1[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) { m0[t] = M[(size_t)c0 * n + k - t]; m1[t] = TWO ? M[(size_t)c1 * n + k - t] : m0[t]; }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int kk = k - t;
        const double2 src = kk < 32 ? x0 : x1;
        const double2 bk = make_double2(__shfl_sync(0xffffffffu, src.x, kk & 31), __shfl_sync(0xffffffffu, src.y, kk & 31));
        if (r0 < kk) x0 = cfnma2(x0, make_double2(m0[t].x, -m0[t].y), bk);
        if (TWO && r1 < kk) x1 = cfnma2(x1, make_double2(m1[t].x, -m1[t].y), bk);
      }
    }
    for (; k > 0; --k) {
      const double2 src = k < 32 ? x0 : x1;
      const double2 bk = make_double2(__shfl_sync(0xffffffffu, src.x, k & 31), __shfl_sync(0xffffffffu, src.y, k & 31));
      if (r0 < k) { const double2 l = M[(size_t)r0 * n + k]; x0 = cfnma2(x0, make_double2(l.x, -l.y), bk); }
      if (TWO && r1 < k) { const double2 l = M[(size_t)r1 * n + k]; x1 = cfnma2(x1, make_double2(l.x, -l.y), bk); }
    }
  }
  const int *perm = lu_perm(ipiv, n);  // P^T: the value at position r goes back to row perm[r]
  if (r0 < n) x[perm[r0]] = x0;
  if (r1 < n) x[perm[r1]] = x1;
  __syncwarp();
}

template <bool TWO>
__device__ void w_d2_matvec(const double *__restrict__ D2, int n, const double2 *x, double2 *y, bool transpose) {
  // Rows lane and lane + 32 in one sweep; each row's sum keeps its order j = 0, 1, ...  The table comes from the L2 (the
  // matrices leave the L1 no room), so the loads are issued sixteen columns at a time ahead of the sums, and the last
  // block is masked (index clamped, sums predicated) instead of left to a one-column-per-round-trip remainder loop.
  const int lane = threadIdx.x & 31;
  const int i0 = min(lane, n - 1), i1 = min(lane + 32, n - 1);
  double sr0 = 0.0, si0 = 0.0, sr1 = 0.0, si1 = 0.0;
  const size_t s0 = transpose ? (size_t)i0 * n : (size_t)i0, s1 = transpose ? (size_t)i1 * n : (size_t)i1;
  const size_t sj = transpose ? 1 : (size_t)n;   // D2[j n + i] or, transposed, D2[i n + j]
  for (int j = 0; j < n; j += 16) {
    double d0[16], d1[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const size_t jj = (size_t)min(j + t, n - 1) * sj;
      d0[t] = __ldg(D2 + s0 + jj);
      if (TWO) d1[t] = __ldg(D2 + s1 + jj);
    }
#pragma unroll
    for (int t = 0; t < 16; ++t)
      if (j + t < n) {
        const double2 xj = x[j + t];
        sr0 = fma(d0[t], xj.x, sr0); si0 = fma(d0[t], xj.y, si0);
        if (TWO) { sr1 = fma(d1[t], xj.x, sr1); si1 = fma(d1[t], xj.y, si1); }
      }
  }
  if (lane < n) y[lane] = make_double2(sr0, si0);
  if (TWO && lane + 32 < n) y[lane + 32] = make_double2(sr1, si1);
  __syncwarp();
}

__device__ void w_start_vector(double2 *x, int n) {
  for (int i = threadIdx.x & 31; i < n; i += 32) {
    const double s = sin(3.0 * (i + 1.0) / (n + 1.0)) + 0.37 * cos(11.0 * (i + 1.0) / (n + 1.0));
    x[i] = make_double2(s, 0.25 * s * s - 0.1);
  }
  __syncwarp();
}

template <bool TWO>
__device__ double2 w_inverse_iterate(const WarpShared &s, int n, const double *D2, const PencilCoef &p, double2 sigma,
                                     int maxit, double tol, int early, double *delta, int *iters, int *stalled = nullptr) {
  const int lane = threadIdx.x & 31;
  double2 om = sigma;
  double dl = 1.0e300;
  ConvTrack ct;
  conv_reset(ct);
  int it = 0;
  for (; it < maxit; ++it) {
    w_d2_matvec<TWO>(D2, n, s.x, s.y, false);
    for (int i = lane; i < n; i += 32) s.y[i] = cadd2(cmul2(p.b2, s.y[i]), cmul2(p.b0, s.x[i]));
    __syncwarp();
    w_solve<TWO>(s.M, n, s.ipiv, s.y);
    double2 yx = make_double2(0.0, 0.0), yy = make_double2(0.0, 0.0);
    for (int i = lane; i < n; i += 32) {
      yx = cadd2(yx, cmulc2(s.y[i], s.x[i]));
      yy.x += s.y[i].x * s.y[i].x + s.y[i].y * s.y[i].y;
    }
    yx = wsum(yx);
    yy = wsum(yy);
    const double2 om_new = cadd2(sigma, make_double2(yx.x / yy.x, yx.y / yy.x));
    const double rn = rsqrt(yy.x);
    for (int i = lane; i < n; i += 32) s.x[i] = make_double2(s.y[i].x * rn, s.y[i].y * rn);
    __syncwarp();
    const double change = hypot(om_new.x - om.x, om_new.y - om.y);
    om = om_new;
    const int cv = conv_push(ct, it, change, tol * hypot(om.x, om.y), early, &dl);
    if (cv) { ++it; if (stalled) *stalled = (cv == 2); break; }
  }
  *delta = dl;
  *iters = it;
  return om;
}

template <bool TWO>
__global__ void __launch_bounds__(32) chan_eig_warp_kernel(const __grid_constant__ ChanArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int n = a.n, lane = threadIdx.x & 31;
  WarpShared s = wcarve(smem, n);
  for (int64_t inst = blockIdx.x; inst < a.ninst; inst += gridDim.x) {
    const PencilCoef p = pencil_coef(a.alpha[inst], a.Re[inst]);
    double2 sigma = a.sigma[inst];
    w_start_vector(s.x, n);
    double2 om = sigma;
    double dl = 1.0e300;
    int total = 0;
    for (int o = 0; o < a.outer; ++o) {
      w_assemble<TWO>(s.M, n, a.D2, a.D4, a.u, a.d2u, p, sigma);
      w_lu<TWO>(s.M, n, s.ipiv);
      int it = 0, stalled = 0;
      om = w_inverse_iterate<TWO>(s, n, a.D2, p, sigma, a.inner, a.tol, a.early, &dl, &it, &stalled);
      total += it;
      if (dl <= a.tol * hypot(om.x, om.y) || stalled) break;
      sigma = om;
    }
    if (lane == 0) { a.omega[inst] = om; a.delta[inst] = dl; a.iters[inst] = total; }
    if (a.evec || a.adj) {
      double best = -1.0;
      int bi = 0;
      for (int i = lane; i < n; i += 32) {
        const double m2 = s.x[i].x * s.x[i].x + s.x[i].y * s.x[i].y;
        if (m2 > best) { best = m2; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      const double2 esc = cdiv2(make_double2(1.0, 0.0), s.x[bi]);
      if (a.evec) {
        double2 *ev = a.evec + (size_t)inst * (n + 2);
        for (int i = lane; i < n; i += 32) ev[i + 1] = cmul2(esc, s.x[i]);
        if (lane == 0) { ev[0] = make_double2(0.0, 0.0); ev[n + 1] = make_double2(0.0, 0.0); }
      }
      if (a.adj) {  // see chan_eig_body
        const double2 sg = make_double2(om.x + 1.0e-8 * hypot(om.x, om.y), om.y);
        w_assemble<TWO>(s.M, n, a.D2, a.D4, a.u, a.d2u, p, sg);
        w_lu<TWO>(s.M, n, s.ipiv);
        for (int i = lane; i < n; i += 32) s.t[i] = make_double2(s.x[i].x, 0.3 - s.x[i].y);
        __syncwarp();
        for (int k = 0; k < 4; ++k) {
          w_d2_matvec<TWO>(a.D2, n, s.t, s.y, true);
          for (int i = lane; i < n; i += 32) s.y[i] = cadd2(cmulc2(p.b2, s.y[i]), cmulc2(p.b0, s.t[i]));
          __syncwarp();
          w_solve_h<TWO>(s.M, n, s.ipiv, s.y);
          double nn = 0.0;
          for (int i = lane; i < n; i += 32) nn += s.y[i].x * s.y[i].x + s.y[i].y * s.y[i].y;
          const double rn = rsqrt(wsum(make_double2(nn, 0.0)).x);
          for (int i = lane; i < n; i += 32) s.t[i] = make_double2(s.y[i].x * rn, s.y[i].y * rn);
          __syncwarp();
        }
        double2 *av = a.adj + (size_t)inst * (n + 2);
        if (a.adj_form == OSB_ADJ_PENCIL) {
          double2 d = make_double2(0.0, 0.0);
          for (int i = lane; i < n; i += 32) d = cadd2(d, cmulc2(s.t[i], cmul2(esc, s.x[i])));
          d = wsum(d);
          const double2 sc = cdiv2(make_double2(1.0, 0.0), make_double2(d.x, -d.y));
          for (int i = lane; i < n; i += 32) av[i + 1] = cmul2(sc, s.t[i]);
        } else {
          w_d2_matvec<TWO>(a.D2, n, s.t, s.y, true);
          double nn = 0.0, vb = -1.0;
          int vi = 0;
          for (int i = lane; i < n; i += 32) {
            const double2 v = cadd2(cmulc2(p.b2, s.y[i]), cmulc2(p.b0, s.t[i]));
            s.y[i] = v;
            const double m2 = v.x * v.x + v.y * v.y;
            nn += m2;
            if (m2 > vb) { vb = m2; vi = i; }
          }
          const double rn = rsqrt(wsum(make_double2(nn, 0.0)).x);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, vb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
            if (ob > vb || (ob == vb && oi < vi)) { vb = ob; vi = oi; }
          }
          __syncwarp();
          const double2 vk = s.y[vi];
          const double ak = hypot(vk.x, vk.y);
          const double2 ph = make_double2(rn * vk.x / ak, -rn * vk.y / ak);
          for (int i = lane; i < n; i += 32) av[i + 1] = cmul2(ph, s.y[i]);
        }
        if (lane == 0) { av[0] = make_double2(0.0, 0.0); av[n + 1] = make_double2(0.0, 0.0); }
      }
    }
    __syncwarp();
  }
}

template <bool TWO>
__global__ void __launch_bounds__(32) chan_neutral_warp_kernel(const __grid_constant__ NeutralArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int n = a.n, lane = threadIdx.x & 31;
  WarpShared s = wcarve(smem, n);
  for (int64_t inst = blockIdx.x; inst < a.ninst; inst += gridDim.x) {
    const double Re = a.Re[inst];
    double nalpha = a.alpha0[inst];
    double2 sigma = a.sigma0[inst];
    double2 ev = make_double2(1.0, 1.0);
    int it = 0, st = OSB_ST_MAXCOUNT;
    while (fabs(ev.y) > a.tol && it < a.maxit) {
      const PencilCoef p = pencil_coef(make_double2(nalpha, 0.0), Re);
      w_start_vector(s.x, n);
      double dl = 1.0e300;
      for (int o = 0; o < 3; ++o) {
        w_assemble<TWO>(s.M, n, a.D2, a.D4, a.u, a.d2u, p, sigma);
        w_lu<TWO>(s.M, n, s.ipiv);
        int k = 0;
        ev = w_inverse_iterate<TWO>(s, n, a.D2, p, sigma, 8, 1.0e-14, a.early, &dl, &k);
        if (dl <= 1.0e-14 * hypot(ev.x, ev.y)) break;
        sigma = ev;
      }
      for (int i = lane; i < n; i += 32) s.t[i] = make_double2(s.x[i].x, -s.x[i].y + 0.3);
      __syncwarp();
      for (int k = 0; k < 6; ++k) {
        w_solve_h<TWO>(s.M, n, s.ipiv, s.t);
        w_d2_matvec<TWO>(a.D2, n, s.t, s.y, true);
        double nn = 0.0;
        for (int i = lane; i < n; i += 32) {
          const double2 v = cadd2(cmulc2(p.b2, s.y[i]), cmulc2(p.b0, s.t[i]));
          s.t[i] = v;
          nn += v.x * v.x + v.y * v.y;
        }
        const double rn = rsqrt(wsum(make_double2(nn, 0.0)).x);
        for (int i = lane; i < n; i += 32) s.t[i] = make_double2(s.t[i].x * rn, s.t[i].y * rn);
        __syncwarp();
      }
      w_d2_matvec<TWO>(a.D2, n, s.x, s.y, false);
      double2 dw = make_double2(0.0, 0.0), dalp = make_double2(0.0, 0.0);
      for (int i = lane; i < n; i += 32) {
        const double2 d2x = s.y[i], xi = s.x[i];
        const double2 tdw = csub2(csub2(cmul2(cmul2(p.db0, ev), xi), cmul2(coef_d2(p, a.u[i]), d2x)),
                                  cmul2(coef_d0(p, a.u[i], a.d2u[i]), xi));
        const double2 tbx = cadd2(cmul2(p.b2, d2x), cmul2(p.b0, xi));
        dw = cadd2(dw, cmulc2(s.t[i], tdw));
        dalp = cadd2(dalp, cmulc2(s.t[i], tbx));
      }
      dw = wsum(dw);
      dalp = wsum(dalp);
      const double2 q = cdiv2(dw, dalp);
      const double2 dwda = make_double2(-q.x, -q.y);
      const double dalpha = -ev.y / dwda.y;
      nalpha = nalpha + a.sfac * dalpha;
      const double nw = ev.x + a.sfac * dalpha * dwda.x;
      if (a.trace && lane == 0) {
        double *r = a.trace + ((size_t)inst * a.maxit + it) * 7;
        r[0] = nalpha; r[1] = ev.x; r[2] = ev.y; r[3] = dwda.x; r[4] = dwda.y; r[5] = dalpha; r[6] = nw;
      }
      sigma = make_double2(nw, ev.y);
      ++it;
      if (!(isfinite(nalpha) && isfinite(ev.x) && isfinite(ev.y))) { st = OSB_ST_NONFINITE; break; }
      __syncwarp();
    }
    if (fabs(ev.y) <= a.tol) st = OSB_ST_CONVERGED;
    if (lane == 0) { a.alpha_out[inst] = nalpha; a.omega_out[inst] = ev; a.steps[inst] = it; a.status[inst] = st; }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// MAKE_DERIVATIVES' products D2 = D1hat D1, D3 = D1hat D2, D4 = D1hat D3 (orrfdchan.f:318-342: three (N+1)^3 triple
// loops on one core, ~3 s at N = 1024).  One thread per entry, the k loop ascending with separately rounded
// multiply and add: the same operations in the same order as the reference's loop nest, so the tables keep
// the bits of a no-FMA CPU build.  Row-major (i, j) at [i * ld + j].
// ---------------------------------------------------------------------------
__global__ void table_matmul_kernel(double *__restrict__ C, const double *__restrict__ A, const double *__restrict__ B,
                                    int ld) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= ld) return;
  const double *a = A + (size_t)i * ld;
  double s = 0.0;
  for (int k = 0; k < ld; ++k) s = __dadd_rn(s, __dmul_rn(__ldg(a + k), B[(size_t)k * ld + j]));
  C[(size_t)i * ld + j] = s;
}

cudaError_t launch_table_matmul(double *C, const double *A, const double *B, int ld, cudaStream_t st) {
  const dim3 block(128), grid((ld + 127) / 128, ld);
  table_matmul_kernel<<<grid, block, 0, st>>>(C, A, B, ld);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
struct ChanLaunch {
  int n;
  const double *D2, *D4, *u, *d2u;
};

static cudaError_t chan_grid(int n, int64_t ninst, int *grid, size_t *smem, const void *func, int threads = CH_THREADS,
                             int variant = 0) {
  *smem = variant ? chan_fixed_smem(n, variant) : chan_smem_bytes(n);
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0, per = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, func, threads, *smem);
  if (e != cudaSuccess) return e;
  if (per < 1) return cudaErrorInvalidConfiguration;
  *grid = (int)std::min<int64_t>(ninst, (int64_t)sms * per);
  return cudaSuccess;
}

void chan_profile_dump(const char *tag) {
  long long h[16];
  if (cudaMemcpyFromSymbol(h, g_chan_prof, sizeof(h)) != cudaSuccess) return;
  const char *names[8] = {"assemble", "panel", "swap+writeback", "U12", "trailing(DMMA)", "B x", "solve", "stage"};
  double tot = 0;
  for (int i = 0; i < 8; ++i) tot += (double)h[i];
  fprintf(stderr, "[osb chan profile %s] block 0 cycles:", tag);
  for (int i = 0; i < 8; ++i) fprintf(stderr, " %s=%.1f%%", names[i], tot > 0 ? 100.0 * h[i] / tot : 0.0);
  fprintf(stderr, " total=%.3g\n", tot);
  memset(h, 0, sizeof(h));
  cudaMemcpyToSymbol(g_chan_prof, h, sizeof(h));
}

static bool use_warp_kernels(int n) { return n <= WN_MAX && !getenv("OSB_CHAN_NO_WARP"); }

static cudaError_t warp_grid(int n, int64_t ninst, int *grid, size_t *smem, const void *func) {
  *smem = warp_smem_bytes(n);
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0, per = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, func, 32, *smem);
  if (e != cudaSuccess) return e;
  if (per < 1) return cudaErrorInvalidConfiguration;
  *grid = (int)std::min<int64_t>(ninst, (int64_t)sms * per);
  return cudaSuccess;
}

static bool use_big_kernel(int n) {
  const char *e = getenv("OSB_CHAN_BIG_MIN");  // tuning knob; default from measurements on B200
  return n >= (e ? atoi(e) : CH_BIG_MIN_N);
}
// which build of the big kernel: 2 = two CTAs per SM (while 2 x its shared memory fits), 1 = one CTA per SM two-level
// LU, 0 = the single-level LU; OSB_CHAN_BIG overrides (A/B measurements)
static int big_variant(int n) {
  // two CTAs: 2 x (dynamic + static + 1 KB reserved) must fit the 228 KB of an SM
  static size_t stat = 0;
  if (!stat) {
    cudaFuncAttributes fa;
    stat = cudaFuncGetAttributes(&fa, (const void *)chan_eig_kernel_big2) == cudaSuccess ? fa.sharedSizeBytes : 12 * 1024;
  }
  const bool two_fit = 2 * (chan_fixed_smem(n, 2) + stat + 1024) <= 228 * 1024;
  // measured on B200 (tools/bench_chan.py, 4096 alpha): N = 384: 37.7 k eig/s (build 2) / 27.2 k (1) / 21.7 k (0);
  // N = 512: 22.8 k / 17.6 k / 19.5 k; N = 600 (two CTAs no longer fit): 11.1 k (1) / 10.8 k (0)
  int v = two_fit ? 2 : 0;
  if (const char *e = getenv("OSB_CHAN_BIG")) v = atoi(e);
  if (getenv("OSB_CHAN_LU1")) v = 0;
  if (v == 2 && !two_fit) v = 1;
  return v;
}
static bool use_small_kernel(int n) {
  const char *e = getenv("OSB_CHAN_SMALL_MAX");  // tuning knob; default from measurements on B200
  const int nmax = e ? atoi(e) : 184;  // N = 160: 279 k (small) / 217 k eig/s; N = 200: 140 k / 162 k
  return n <= nmax && !chan_matrix_in_smem(n) && !getenv("OSB_CHAN_NO_SMALL");
}

// largest interior size n the dense CTA kernels can take: the 16-column panel, three vectors and the pivots
// must fit the opt-in shared memory next to the kernels' static arrays
int chan_dense_max_n() {
  static int cached = -1;
  if (cached >= 0) return cached;
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaFuncAttributes fa;
  size_t stat = 0;
  if (cudaFuncGetAttributes(&fa, (const void *)chan_eig_kernel_big) == cudaSuccess) stat = fa.sharedSizeBytes;
  if (cudaFuncGetAttributes(&fa, (const void *)chan_spectrum_kernel) == cudaSuccess) stat = std::max(stat, fa.sharedSizeBytes);
  int n = 64;
  while (n < 4096 && chan_fixed_smem(n + 1) + stat <= (size_t)optin) ++n;
  cached = n;
  return n;
}

cudaError_t chan_eig_grid(int n, int64_t ninst, int *grid, size_t *smem) {
  if (use_warp_kernels(n)) return warp_grid(n, ninst, grid, smem, n > 32 ? (const void *)chan_eig_warp_kernel<true> : (const void *)chan_eig_warp_kernel<false>);
  if (use_big_kernel(n)) {
    const int v = big_variant(n);
    if (v == 2) return chan_grid(n, ninst, grid, smem, (const void *)chan_eig_kernel_big2, CH_THREADS, 2);
    if (v == 1) return chan_grid(n, ninst, grid, smem, (const void *)chan_eig_kernel_big, CH_THREADS_BIG, 1);
    return chan_grid(n, ninst, grid, smem, (const void *)chan_eig_kernel_big_lu1, CH_THREADS_BIG);
  }
  if (use_small_kernel(n)) return chan_grid(n, ninst, grid, smem, (const void *)chan_eig_kernel_small, CH_THREADS_SMALL);
  return chan_grid(n, ninst, grid, smem, (const void *)chan_eig_kernel);
}
cudaError_t chan_neutral_grid(int n, int64_t ninst, int *grid, size_t *smem) {
  if (use_warp_kernels(n)) return warp_grid(n, ninst, grid, smem, n > 32 ? (const void *)chan_neutral_warp_kernel<true> : (const void *)chan_neutral_warp_kernel<false>);
  return chan_grid(n, ninst, grid, smem, (const void *)chan_neutral_kernel);
}

cudaError_t launch_chan_eig(int n, const double *D2, const double *D4, const double *u, const double *d2u,
                            double2 *work, int grid, size_t smem, int64_t ninst, const double2 *alpha,
                            const double *Re, const double2 *sigma, int outer, int inner, double tol,
                            double2 *omega, double *delta, int32_t *iters, double2 *evec, double2 *adj, int adj_form,
                            cudaStream_t st) {
  ChanArgs a;
  a.n = n; a.D2 = D2; a.D4 = D4; a.u = u; a.d2u = d2u; a.work = work; a.ninst = ninst;
  a.alpha = alpha; a.Re = Re; a.sigma = sigma; a.outer = outer; a.inner = inner; a.tol = tol;
  a.omega = omega; a.delta = delta; a.iters = iters; a.evec = evec; a.adj = adj; a.adj_form = adj_form;
  a.early = getenv("OSB_NO_EARLY_STOP") ? 0 : 1;
  if (use_warp_kernels(n)) {
    if (n > 32) chan_eig_warp_kernel<true><<<grid, 32, smem, st>>>(a);
    else chan_eig_warp_kernel<false><<<grid, 32, smem, st>>>(a);
  }
  else if (use_big_kernel(n)) {
    const int v = big_variant(n);
    if (v == 2) chan_eig_kernel_big2<<<grid, CH_THREADS, smem, st>>>(a);
    else if (v == 1) chan_eig_kernel_big<<<grid, CH_THREADS_BIG, smem, st>>>(a);
    else chan_eig_kernel_big_lu1<<<grid, CH_THREADS_BIG, smem, st>>>(a);
  }
  else if (use_small_kernel(n)) chan_eig_kernel_small<<<grid, CH_THREADS_SMALL, smem, st>>>(a);
  else chan_eig_kernel<<<grid, CH_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_chan_neutral(int n, const double *D2, const double *D4, const double *u, const double *d2u,
                                double2 *work, int grid, size_t smem, int64_t ninst, const double *Re,
                                const double *alpha0, const double2 *sigma0, double sfac, double tol, int maxit,
                                double *alpha_out, double2 *omega_out, int32_t *steps, int32_t *status,
                                double *trace, cudaStream_t st) {
  NeutralArgs a;
  a.n = n; a.D2 = D2; a.D4 = D4; a.u = u; a.d2u = d2u; a.work = work; a.ninst = ninst;
  a.Re = Re; a.alpha0 = alpha0; a.sigma0 = sigma0; a.sfac = sfac; a.tol = tol; a.maxit = maxit;
  a.alpha_out = alpha_out; a.omega_out = omega_out; a.steps = steps; a.status = status; a.trace = trace;
  a.early = getenv("OSB_NO_EARLY_STOP") ? 0 : 1;
  if (use_warp_kernels(n)) {
    if (n > 32) chan_neutral_warp_kernel<true><<<grid, 32, smem, st>>>(a);
    else chan_neutral_warp_kernel<false><<<grid, 32, smem, st>>>(a);
  }
  else chan_neutral_kernel<<<grid, CH_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace osb
