/*
 * osb.h -- C ABI of the B200-native Orr-Sommerfeld engine (libosb.so).
 *
 * This is the drop-in boundary for the one data-parallel hot path of
 * sscollis/os-stab: the per-point eigenvalue solves behind the reference's
 * Fortran-77 operator interfaces.  Each entry point below names the reference
 * interface it replaces (file:line into the reference tree).  The reference has
 * no FFI of its own; INTEGRATION.md shows the ISO_C_BINDING interface block and
 * the one-line change in each driver that binds these symbols.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; complex data is interleaved
 *     (re, im) double pairs == Fortran complex(c_double_complex) == CUDA double2.
 *   - All physical quantities are in the SOLVER's units, i.e. what the reference
 *     passes to CONTE/CONTE_IC after its driver-side rescale (contebl.f:181-182,
 *     orrspace.f:160-168).
 *   - Every function returns 0 on success, a negative OSB_E_* code otherwise;
 *     osb_last_error() returns the message.  Per-point outcomes travel in the
 *     `status` arrays (OSB_ST_*).  Nothing here falls back to a CPU path: with
 *     no usable CUDA device every call fails with OSB_E_CUDA.
 *   - `*_dev` variants take DEVICE pointers (inputs already resident in HBM);
 *     the plain variants take HOST pointers and do the copies themselves.
 *   - A context owns one CUDA device (one process per GPU is the intended
 *     deployment; osb_create_multi() drives several devices from one host
 *     thread for the single-process Fortran hosts).
 */
#ifndef OSB_H
#define OSB_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSB_VERSION 2

/* error codes */
#define OSB_OK 0
#define OSB_E_ARG (-1)   /* bad argument                                     */
#define OSB_E_CUDA (-2)  /* CUDA runtime error / no device                   */
#define OSB_E_ALLOC (-3) /* out of memory                                    */
#define OSB_E_STATE (-4) /* object used before it is ready                   */

/* per-point status */
#define OSB_ST_CONVERGED 0 /* eigenvalue loop left by tolerance               */
#define OSB_ST_MAXCOUNT 1  /* left by the pass limit (shoot.f:478-479)        */
#define OSB_ST_NONFINITE 2 /* NaN/Inf met (the reference would SIGFPE)        */
#define OSB_ST_UNVERIFIED 3 /* osb_chan_eig only: converged, but the choice of mode could not be
                              cross-checked against the full spectrum (see osb_chan_eig)       */

/* flows: which reference driver's iteration is reproduced */
#define OSB_FLOW_CONTEBL 0  /* CONTE_IC via contebl.f:205   (shoot.f:391-813) */
#define OSB_FLOW_CONTE 1    /* CONTE via conte.f:131        (shoot.f:26-388)  */
#define OSB_FLOW_ORRSOM 2   /* private CONTE orrsom.f:165-506                 */
#define OSB_FLOW_ORRSPACE 3 /* private CONTE orrspace.f:209-582 (spatial)     */

/* integrators (gcc.mak:46-48: USE_RKCK45 or the default RK4) */
#define OSB_INT_RK4 0
#define OSB_INT_CK45 1

typedef struct osb_ctx osb_ctx;
typedef struct osb_profile osb_profile;

/* ---- context ---------------------------------------------------------- */
int osb_version(void);
/* one context bound to CUDA device `device` */
int osb_create(osb_ctx **ctx, int device);
/* one context driving `ndev` devices (ndev = 0: all visible) from one thread */
int osb_create_multi(osb_ctx **ctx, int ndev, const int *devices);
int osb_destroy(osb_ctx *ctx);
int osb_device_count(const osb_ctx *ctx);
const char *osb_last_error(const osb_ctx *ctx);
/* the stream the context launches on (as a cudaStream_t cast to void*) for
 * device 0 of the context; lets a caller time the kernels with CUDA events */
void *osb_stream(osb_ctx *ctx);
/* number of kernels launched by this context so far */
int64_t osb_launch_count(const osb_ctx *ctx);

/* ---- mean profile (replaces SOLVE_BL: contebl.f:386-533, orrsom.f:687-804,
 *      orrspace.f:697-817; and the analytic CHANNEL conte.f:311-323) -------- */
/* Solve the Blasius / Falkner-Skan similarity equation on the device and keep
 * the tables resident.  variant = OSB_FLOW_CONTEBL | _ORRSOM | _ORRSPACE picks
 * which copy of SOLVE_BL is reproduced (integrator is honoured by CONTEBL only,
 * the other two are RK4 as in the reference). */
int osb_profile_blasius(osb_ctx *ctx, int variant, int integrator, int nbl,
                        double ymin, double ymax, double f2p, double beta,
                        osb_profile **prof);
/* Batched form: nprof independent (beta, f2p) pairs, same grid; tables land in
 * caller (host) arrays of shape [nprof][nbl+1] each; any pointer may be NULL. */
int osb_profile_blasius_batch(osb_ctx *ctx, int variant, int integrator, int nbl,
                              double ymin, double ymax, int64_t nprof,
                              const double *f2p, const double *beta, double *U,
                              double *Uspl, double *d2U, double *d2Uspl,
                              double *f2p_out, int32_t *npass);
/* Upload tables computed elsewhere (e.g. by an unmodified host SOLVE_BL). */
int osb_profile_upload(osb_ctx *ctx, int variant, int nbl, double ymin, double ymax,
                       const double *U, const double *Uspl, const double *d2U,
                       const double *d2Uspl, osb_profile **prof);
/* A copy of `prof` in the frame moving with u_c, splines rebuilt (replaces SHIFT_BL,
 * conteucbl.f:924-968): U = U - uc; Uspl = SPLINE(U); d2U := Uspl; d2Uspl = SPLINE(d2U). */
int osb_profile_shift(osb_ctx *ctx, const osb_profile *prof, double uc, osb_profile **shifted);
/* The channel profile U = 1 - y^2 (no tables). */
int osb_profile_channel(osb_ctx *ctx, osb_profile **prof);
/* Copy the tables back ([nbl+1] each; any pointer may be NULL). */
int osb_profile_get(osb_ctx *ctx, const osb_profile *prof, double *ydat, double *U,
                    double *Uspl, double *d2U, double *d2Uspl, int32_t *npass,
                    double *f2p_last);
int osb_profile_nbl(const osb_profile *prof);
int osb_profile_free(osb_ctx *ctx, osb_profile *prof);

/* ---- shooting (replaces CONTE_IC shoot.f:391-392, CONTE shoot.f:26-27,
 *      and the private CONTEs orrsom.f:165, orrspace.f:209) ----------------- */
typedef struct {
  int32_t flow;       /* OSB_FLOW_*                                          */
  int32_t integrator; /* OSB_INT_*                                           */
  int32_t nstep;      /* integration steps (CONTE's nstep)                   */
  int32_t maxcount;   /* 0 = the flow's own limit (20; orrspace 50)          */
  double testalpha;   /* contebl/orrsom: degrees; orrspace: cosine bound;
                         conte: ignored (hard-wired 10 deg, shoot.f:64-65)   */
  double ymin, ymax;  /* CONTE's to, tf                                      */
  double errtol;      /* 0 = the flow's own (1e-14 / 1e-15 / 1e-8 / 1e-8)    */
  double eigtol;      /* 0 = the flow's own (1e-14 / 1e-15 / 1e-12 / 1e-12)  */
} osb_shoot_opts;

/* Fill opts with what the named reference driver uses. */
int osb_shoot_defaults(int flow, osb_shoot_opts *opts);

/*
 * Batched temporal eigenvalue solve: npts independent points, each exactly the
 * iteration of CONTE_IC / CONTE (secant then Muller on c).
 *   alpha  [2*npts]  complex wavenumber           (in)
 *   Re     [npts]                                  (in)
 *   c      [2*npts]  guess (in) -> ctemp, the last value integrated (out)
 *   passes [npts]    integrations performed ("Eigenvalue iterations")
 *   qend   [npts]    normalisations in the last pass
 *   status [npts]    OSB_ST_*
 *   resid  [2*npts]  optional: |err|, |c-cm1| at exit (shoot.f:756-758)
 *   history[npts*hist_cap*4] optional: per pass ctemp (conte: updated c), err
 *                    -- the rows of the reference's stdout table
 */
int osb_shoot_temporal(osb_ctx *ctx, const osb_shoot_opts *opts,
                       const osb_profile *prof, int64_t npts, const double *alpha,
                       const double *Re, double *c, int32_t *passes, int32_t *qend,
                       int32_t *status, double *resid, double *history,
                       int32_t hist_cap);
int osb_shoot_temporal_dev(osb_ctx *ctx, const osb_shoot_opts *opts,
                           const osb_profile *prof, int64_t npts,
                           const double *d_alpha, const double *d_Re, double *d_c,
                           int32_t *d_passes, int32_t *d_qend, int32_t *d_status,
                           double *d_resid, double *d_history, int32_t hist_cap);

/*
 * Neutral curve by shooting (the reference has no program for it -- orrncbl's
 * boundary-layer branch is dead code -- so there is no parity target, only physics checks):
 * for every Re_i the real alpha_i with Im c(alpha_i, Re_i) = 0, by a secant iteration on alpha
 * whose every step is ONE batched temporal solve (osb_shoot_temporal's kernels) over the points
 * still iterating.  The eigenvalue guess of a step is continued linearly from the two previous ones.
 *   Re [npts], alpha0 [npts] (real, solver units), c0 [2*npts] eigenvalue guess at alpha0
 *   dalpha_rel   the second secant point is alpha0 (1 + dalpha_rel)
 *   tol          stop when |Im c| <= tol ;  maxit secant steps at most
 *   alpha_out [npts], c_out [2*npts], steps [npts],
 *   status [npts]: OSB_ST_CONVERGED, OSB_ST_MAXCOUNT (maxit reached or an inner solve did not
 *   converge), OSB_ST_NONFINITE
 * Temporal flows only (CONTEBL, CONTE, ORRSOM).
 */
int osb_shoot_neutral(osb_ctx *ctx, const osb_shoot_opts *opts, const osb_profile *prof,
                      int64_t npts, const double *Re, const double *alpha0, const double *c0,
                      double dalpha_rel, double tol, int32_t maxit,
                      double *alpha_out, double *c_out, int32_t *steps, int32_t *status);

/*
 * Self-seeding temporal sweep over a tensor grid (alpha_j, Re_i), j < na, i < nre (both ascending, solver units).
 * The reference's only seeding scheme is continuation from a guess typed into a deck (contebl.f:75-76; the omega
 * loop orrspace.f:186-195 hands each converged alpha to the next point), and CONTE_IC only converges from within
 * ~3e-3 of the root, so a grid sweep needs a seed per point.  This driver needs ONE: the guess anchor_c at
 * (anchor_alpha, anchor_Re) (e.g. the contebl defaults).  It solves the anchor, then a coarse grid of at most
 * `coarse` x `coarse` nodes (0 = 65) by batched continuation on the device (linear extrapolation along each
 * walker's path; failed steps bisected in (alpha, log Re)), interpolates the coarse solutions bilinearly in
 * (alpha, log Re) to every grid point and solves the whole grid in one batched call (osb_shoot_temporal's kernels).
 *   c, passes, qend, status [nre*na]   point (i, j) at i*na + j (c interleaved re, im)
 *   seed  [2*nre*na] optional          the guess each point was started from
 *   info  [4] optional                 coarse solves, coarse rounds (batched calls), bisections, coarse nodes given up
 * Temporal flows only.  Points whose iteration does not converge are reported through `status` as always.
 */
int osb_shoot_grid_temporal(osb_ctx *ctx, const osb_shoot_opts *opts, const osb_profile *prof, int32_t na,
                            const double *alpha, int32_t nre, const double *Re, double anchor_alpha,
                            double anchor_Re, const double *anchor_c, int32_t coarse, double *c,
                            int32_t *passes, int32_t *qend, int32_t *status, double *seed, int32_t *info);

/*
 * Batched spatial continuation lines (the omega sweep of orrspace.f:186-195):
 * nlines independent lines, each niter points; point i of a line starts from
 * the converged alpha of point i-1 (sequential inside a line, parallel across).
 *   Re     [nlines]
 *   omega0 [2*nlines], domega [nlines]   omega_i = omega0 + i*domega (real part)
 *   alpha0 [2*nlines]  guess for the first point
 *   alpha_out [nlines*niter*2], passes/qend/status [nlines*niter]
 *   history [nlines*niter*hist_cap*4] optional: per pass alpha (ctemp), err -- the rows of
 *   the per-pass table orrspace prints (orrspace.f:485-488)
 */
int osb_shoot_spatial_lines(osb_ctx *ctx, const osb_shoot_opts *opts,
                            const osb_profile *prof, int64_t nlines, int32_t niter,
                            const double *Re, const double *omega0,
                            const double *domega, const double *alpha0,
                            double *alpha_out, int32_t *passes, int32_t *qend,
                            int32_t *status, double *history, int32_t hist_cap);

/*
 * Eigenfunction pass (the "second pass" shoot.f:762-810 / 338-385): for each
 * point re-integrates at eig (the ctemp returned above), back-substitutes and
 * returns y(i,k), i=1..4, k=0..nstep UN-normalised plus the reference's vmax
 * (divide to obtain the fort.11-14 columns).
 *   y    [npts*(nstep+1)*4*2], vmax [npts*2]
 */
int osb_eigfun(osb_ctx *ctx, const osb_shoot_opts *opts, const osb_profile *prof,
               int64_t npts, const double *alpha, const double *Re,
               const double *eig, const double *omega, double *y, double *vmax,
               int32_t *qend);

/* ---- channel matrix path (replaces MAKE_DERIVATIVES + MAKE_MATRIX +
 *      SOLVE_ORR_SOM of orrfdchan.f:273-350,553-659,688-823; orrcolchan.f:77-154,
 *      272-513; orrncbl.f:122-202,553-695,758-981) ------------------------- */
#define OSB_DISC_FD 0       /* orrfdchan: FiniteD1 (orrfdchan.f:1183), analytic U''   */
#define OSB_DISC_CHEB 1     /* orrcolchan: CHEBYD (orrcolchan.f:904), analytic U''    */
#define OSB_DISC_CHEB_NUM 2 /* orrncbl: CHEBYD, U'' by collocation (orrncbl.f:520-545) */

typedef struct osb_chan osb_chan;

/* Grid, profile and derivative tables for N mesh intervals (the reference's n):
 * cosine grid, D1hat, D1 with rows 0,N zeroed, D2 = D1hat D1, D3, D4 formed with the
 * reference's own triple loops on the host (once per N), interior blocks uploaded. */
int osb_chan_create(osb_ctx *ctx, int disc, int N, osb_chan **chan);
/*
 * Mapped boundary-layer operator (orrcolbl.f / orrfdbl.f): the semi-infinite domain
 * y = Lmap (1+eta)/(1-eta), metrics m1..m4 of MAKE_BL_METRICS (orrcolbl.f:245-311), derivative matrices of
 * MAKE_DERIVATIVES (orrcolbl.f:161-242: CHEBYD on the cosine grid; orrfdbl.f:145-222: FiniteD1 on the uniform
 * grid eta_i = 1 - 2i/N), operator of MAKE_MATRIX (orrcolbl.f:470-600).  The reference reads the mean profile
 * from an external file that it does not ship; here the caller passes U at the N+1 grid points
 * (osb_chan_bl_points gives y_i; U[0..merge] is the free stream, set to 1 exactly) -- e.g. the Blasius
 * profile of osb_profile_blasius.  d1u = D1hat u, d2u = D1hat^2 u, zeroed for j <= merge-2 (orrcolbl.f:404-427).
 * The metrics are folded into row-scaled real tables on the host,
 *     A = D4' - (2 alpha^2 + i alpha Re U) D2' + w0 I,   B = -i Re D2' + i alpha^2 Re I,
 *     D2' = m1^2 D2 + m2 D1,   D4' = m1^4 D4 + 6 m1^2 m2 D3 + (3 m2^2 + 4 m1 m3) D2 + m4 D1,
 * which is the channel operator's form: every osb_chan_* entry point (and every kernel) works on the
 * handle unchanged.  osb_chan_eig without shifts keeps the most unstable mode found by the scan of
 * 0 < c_r < 1 (orrcolbl.f:771-779 keeps the largest Im w of the full spectrum).
 * Parity of this operator is unpinned (the reference programs need IMSL and the missing profile file);
 * tests check it against the oracle's restatement and against the shooting eigenvalue of the same flow.
 */
int osb_chan_create_bl(osb_ctx *ctx, int disc /* OSB_DISC_CHEB | OSB_DISC_FD */, int N, double lmap,
                       const double *U /* [N+1] */, int merge, osb_chan **out);
/* y_i = Lmap (1+eta_i)/(1-eta_i) of the grid osb_chan_create_bl uses (y_0 = +inf); host-only helper */
int osb_chan_bl_points(int disc, int N, double lmap, double *y /* [N+1] */);

int osb_chan_free(osb_ctx *ctx, osb_chan *chan);
/* eta, u, d2u: [N+1]; D2, D4: [(N+1)*(N+1)] row-major (i,j); any may be NULL */
int osb_chan_tables(osb_ctx *ctx, const osb_chan *chan, double *eta, double *u, double *d2u,
                    double *D2, double *D4);

/*
 * Batched SOLVE_ORR_SOM: for each alpha the eigenvalue the reference's driver keeps
 * (the least stable mode: largest Im w, orrfdchan.f:815-823 / orrncbl.f:897-905) and,
 * optionally, its eigenvector on the full grid normalised by the component of largest
 * modulus (orrfdchan.f:849-857).  The interior (N-1)x(N-1) pencil is used (the BC rows
 * of orrcolchan.f:424-435 only add infinite eigenvalues).
 *   alpha  [2*npts]; sigma0 [2*npts] optional shift guesses; omega [2*npts] out.
 *   sigma0 = NULL: the mode is chosen by the reference's rule.  A scan of real shifts (16 over 0 < c_r < 1,
 *   plus 32 over -4 < w_r < 0 for the FD rule / 8 over -1 < w_r < 0 for the |w_r| < 1 rule) proposes it; the
 *   proposal is cross-checked against the device spectrum (osb_chan_spectrum) under the same rule wherever no
 *   shift settled, at both ends of the sweep and across every jump of the chosen branch, and the whole sweep
 *   (up to 512 points) is re-checked if any of those disagrees.  Points left unchecked after a disagreement
 *   come back as OSB_ST_UNVERIFIED.  For N beyond the dense kernels' limit (FD, N > ~700) no spectrum is
 *   available and the scan's choice stands (a mode with w_r outside the scanned windows is then missed).
 *   evec   [npts*(N+1)*2] optional; iters [npts] inverse iterations; status [npts].
 *   evec_adj [npts*(N+1)*2] optional: the LEFT (adjoint) eigenvector of the same mode -- the reference solves
 *   with JOBVL = 'V' (orrfdchan.f:796, orrcolchan.f:461) and writes it as columns 4-5 of fort.11.  By inverse
 *   iteration on the adjoint pencil, u <- (A - s B)^{-H} B^H u, zeros at the walls, in the form adj_form:
 *     OSB_ADJ_PENCIL  u with u^H (A - w B) = 0 scaled so that u^H evec = 1: ZGEGV's left vector after the
 *                     bi-orthonormalisation of orrcolchan.f:537-545
 *     OSB_ADJ_INVBA   B^H u, the left eigenvector of inv(B) A with ZGEEV's normalisation (unit 2-norm, largest
 *                     component real and positive): what orrfdchan.f:796 returns in `avec`
 */
#define OSB_ADJ_PENCIL 0
#define OSB_ADJ_INVBA 1
int osb_chan_eig(osb_ctx *ctx, const osb_chan *chan, double Re, int64_t npts, const double *alpha,
                 const double *sigma0, double *omega, double *evec, double *evec_adj, int adj_form,
                 int32_t *iters, int32_t *status);

/*
 * The FULL spectrum of the interior pencil at each alpha: what SOLVE_ORR_SOM computes before it keeps one
 * mode (orrfdchan.f:792-797 inv(B)A + ZGEEV; the table orrcolchan.f:505-513 / orrfdchan.f:834-838 prints).
 * C = inv(B) A, Hessenberg reduction and single-shift complex QR run on the device, one CTA per alpha;
 * each value is then polished by one shift-invert iteration on the pencil itself (kept when it settles
 * next to its seed), which removes the conditioning of inv(B)A from the well-separated modes.
 * Not a sweep path: osb_chan_eig isolates its mode without the spectrum and is orders of magnitude faster;
 * this entry point serves the interactive listings and lets a caller apply the reference's selection rule
 * to the whole spectrum.
 *   spectrum [npts * (N-1) * 2]  eigenvalues omega of each alpha, ascending in Im(omega) like the
 *                                reference's SORT / PIKSR2 pass (orrfdchan.f:803-811)
 *   status   [npts]              OSB_ST_CONVERGED, or OSB_ST_MAXCOUNT if a QR iteration limit was hit
 */
int osb_chan_spectrum(osb_ctx *ctx, const osb_chan *chan, double Re, int64_t npts, const double *alpha,
                      double *spectrum, int32_t *status);

/*
 * Batched neutral-point search, the Newton iteration on alpha of orrncbl.f:109-116
 * (bug-compatible step: the adjoint is the left eigenvector of inv(B)A, D-7), one
 * independent search per Re:  while |Im w| > tol: alpha += sfac * (-Im w / Im dw/da).
 *   Re, alpha0 [nRe]; alpha_out, omega [nRe] / [2*nRe]; steps, status [nRe];
 *   trace optional [nRe*maxit*7]: nalpha, w_r, w_i, dwda_r, dwda_i, dalpha, nw per step
 */
int osb_chan_neutral(osb_ctx *ctx, const osb_chan *chan, int64_t nRe, const double *Re,
                     const double *alpha0, double sfac, double tol, int32_t maxit, double *alpha_out,
                     double *omega, int32_t *steps, int32_t *status, double *trace);

/* ---- measurement helpers ---------------------------------------------- */
/* Device FP64 peak micro-benchmark: runs a dependent-chain-free DFMA kernel for
 * ~ms_target milliseconds on the context's device; returns TFLOP/s (2 flops
 * per DFMA).  Used as the roofline denominator of the shooting kernel. */
int osb_measure_fp64_peak(osb_ctx *ctx, double ms_target, double *tflops,
                          double *sm_clock_mhz_est);
/* The same for the FP64 tensor path (mma.sync.m8n8k4.f64 = DMMA, 512 flops per warp instruction): the
 * roofline denominator of the dense LU's trailing update (BASELINE.md section 3 asks for both). */
int osb_measure_dmma_peak(osb_ctx *ctx, double ms_target, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* OSB_H */
