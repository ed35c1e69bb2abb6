// osb_internal.h -- shared between the C-ABI layer (osb_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/osb.h"

namespace osb {

// Device-resident mean profile.  The five tables mirror the reference's COMMON
// arrays (contebl.f:44-48): nbl+2 doubles each, the last one zero (SPDER reads
// one element past the table when XX == X(N), see oracle D-1).
struct ProfileDev {
  int variant = 0;      // OSB_FLOW_CONTEBL/_ORRSOM/_ORRSPACE; OSB_FLOW_CONTE = analytic channel
  int nbl = 0;
  double ymin = 0, ymax = 0;
  double *tables = nullptr;  // [5][nbl+2]: ydat, U, Uspl, d2U, d2Uspl  (device)
  int npass = 0;
  double f2p_last = 0;
  int device = 0;
};

// Stream-ordered scratch allocations of one call on one device; released on every exit
// path (the error returns of the C ABI included) when the guard leaves scope.  The memory
// comes from the CONTEXT's private pool (osb_create_multi), not the device's default pool:
// what the library caches between calls is returned to the driver by osb_destroy and no
// process-wide pool attribute is touched.
class Scratch {
 public:
  Scratch(int device, cudaStream_t st, cudaMemPool_t pool) : device_(device), st_(st), pool_(pool) {}
  Scratch(const Scratch &) = delete;
  Scratch &operator=(const Scratch &) = delete;
  ~Scratch() {
    if (ptrs_.empty()) return;
    cudaSetDevice(device_);
    for (void *q : ptrs_) cudaFreeAsync(q, st_);
  }
  template <class T>
  cudaError_t alloc(T **out, size_t bytes) {
    void *q = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync(&q, bytes ? bytes : 1, pool_, st_);
    if (e == cudaSuccess) ptrs_.push_back(q);
    *out = static_cast<T *>(q);
    return e;
  }

 private:
  int device_;
  cudaStream_t st_;
  cudaMemPool_t pool_;
  std::vector<void *> ptrs_;
};

// Everything the shooting kernel needs that is uniform over the launch.
struct ShootArgs {
  // stage table: for step k (1-based) and stage s: tab[(k-1)*NS + s] = (U, U'') at
  // the abscissa the reference's integrator evaluates the RHS at.
  const double2 *tab;
  int nstep;
  int maxcount;
  int niter;         // continuation points per line (1 for temporal solves)
  int spatial;       // unknown is alpha (orrspace) instead of c
  int bl_ic;         // 1: free-stream asymptotes + GS at k=0; 0: unit vectors e3,e4
  int loop_rule;     // 0: (|err|>=errtol || dc>=eigtol) && icount<=max   (shoot.f:478)
                     // 1:  |err|>=errtol && icount<=max && dc>=eigtol     (shoot.f:76, orrsom, orrspace)
  int cm1_defined;   // contebl/conte initialise cm1 = c+1; orrsom/orrspace leave it undefined
  int strict;        // normalise when lhs > thr*rhs (1) or >= (0)
  int first_perturb; // 0: (Re c * .9999, Im c)   1: alpha*1.01 (both parts)
  int show_updated;  // history shows the updated c (conte) instead of ctemp
  int flow;          // OSB_FLOW_* (the strict build evaluates RHS and orthonormality test as typed per driver)
  double testalpha;  // as given (degrees / cosine bound); the fast kernels use thr
  double thr;        // Hermitian: cos^2(testalpha); analytic: testalpha^4
  double hs;         // signed step (-h for the BL flows, +h for the channel)
  double tf;         // where the initial condition is evaluated (ymax)
  double errtol, eigtol;
  double bh[15];     // Cash-Karp b(m,k)*hs, row-major lower triangle; RK4: unused
  double ch[6];      // Cash-Karp c(k)*hs
  int64_t nlines;
  const double2 *alpha;   // [nlines] temporal: given; spatial: first guess
  const double *Re;       // [nlines]
  const double2 *omega0;  // spatial only
  const double *domega;   // spatial only
  double2 *ev;            // temporal: [nlines] c in/out; spatial: [nlines*niter] alpha out
  int32_t *passes, *qend, *status;  // [nlines*niter]
  double2 *resid;         // optional [nlines*niter]: |err|, |ev-cm1|
  double4 *history;       // optional [nlines*niter*hist_cap]
  int hist_cap;
  int pair_lpw;           // lane-pair kernel: lines per warp (1..16), set by launch_shoot
};

// Eigenfunction kernel args (K2b)
struct EigfunArgs {
  ShootArgs s;            // tab, steps, flow switches (iteration fields unused)
  const double2 *eig;     // [npts] converged eigenvalue (c or alpha)
  const double2 *omega;   // spatial only
  double2 *U;             // scratch trajectory [npts][(nstep+1)][8]
  double2 *P;             // scratch [npts][(nstep+1)][4]   (P11, P21, P22, unused)
  double *tq;             // scratch [npts][(nstep+1)]
  double2 *y;             // out [npts][(nstep+1)][4]
  double2 *vmax;          // out [npts]
  int32_t *qend;          // out [npts]
  double t0;              // ymin (to)
  double habs;            // |h|
};

cudaError_t launch_stage_table(const ProfileDev &p, int flow, int integrator, int nstep,
                               double to, double tf, double2 *tab, cudaStream_t st);
int stages_per_step(int integrator);
// queue_counter: zeroed device word enabling the compacting persistent-CTA kernel
// for temporal sweeps (nullptr: one thread per point for the whole solve)
cudaError_t launch_shoot(const ShootArgs &a, int integrator, int analytic_inprod,
                         cudaStream_t st, int *nlaunch, unsigned long long *queue_counter = nullptr);
// strict-parity build of K2 (shoot_strict.cu, -fmad=false); chosen by launch_shoot when OSB_STRICT is set
cudaError_t launch_shoot_strict(const ShootArgs &a, int integrator, int analytic_inprod, cudaStream_t st);
cudaError_t launch_eigfun(const EigfunArgs &a, int integrator, int analytic_inprod,
                          int64_t npts, cudaStream_t st);
cudaError_t launch_blasius(int variant, int integrator, int nbl, double ymin, double ymax,
                           int64_t nprof, const double *d_f2p, const double *d_beta,
                           double *d_tables /*[nprof][5][nbl+2]*/, double *d_scratch,
                           double *d_f2p_out, int32_t *d_npass, cudaStream_t st);
size_t blasius_scratch_doubles(int nbl, int64_t nprof);
// d_out[4][nprof][nbl+1] <- U, Uspl, d2U, d2Uspl of the profile-fastest batch tables
cudaError_t launch_transpose_tables(const double *d_tables, double *d_out, int nbl, int64_t nprof, cudaStream_t st);
cudaError_t launch_shift_profile(int nbl, double uc, const double *d_in, double *d_out, double *d_scratch,
                                 cudaStream_t st);
// channel matrix path (chan_kernels.cu)
size_t chan_smem_bytes(int n);
int chan_dense_max_n();
void chan_profile_dump(const char *tag);
cudaError_t chan_eig_grid(int n, int64_t ninst, int *grid, size_t *smem);
cudaError_t chan_neutral_grid(int n, int64_t ninst, int *grid, size_t *smem);
cudaError_t launch_chan_eig(int n, const double *D2, const double *D4, const double *u, const double *d2u,
                            double2 *work, int grid, size_t smem, int64_t ninst, const double2 *alpha,
                            const double *Re, const double2 *sigma, int outer, int inner, double tol,
                            double2 *omega, double *delta, int32_t *iters, double2 *evec, double2 *adj, int adj_form,
                            cudaStream_t st);
cudaError_t chan_spectrum_grid(int n, int64_t ninst, int *grid, size_t *smem);
cudaError_t launch_chan_spectrum(int n, const double *D2, const double *D4, const double *u, const double *d2u,
                                 double2 *work, int grid, size_t smem, int64_t ninst, const double2 *alpha,
                                 const double *Re, double2 *eig, int32_t *status, cudaStream_t st);
size_t chan_band_work_bytes(int n, int64_t ninst);
cudaError_t launch_chan_fd_banded(int n, const double *D2, const double *D4, const double *u, const double *d2u,
                                  const double2 *band24, void *work, int64_t ninst, const double2 *alpha, const double *Re,
                                  const double2 *sigma, int outer, int inner, double tol, double2 *omega,
                                  double *delta, int32_t *iters, double2 *evec, double2 *adj, int adj_form,
                                  cudaStream_t st);
cudaError_t launch_chan_neutral(int n, const double *D2, const double *D4, const double *u, const double *d2u,
                                double2 *work, int grid, size_t smem, int64_t ninst, const double *Re,
                                const double *alpha0, const double2 *sigma0, double sfac, double tol, int maxit,
                                double *alpha_out, double2 *omega_out, int32_t *steps, int32_t *status,
                                double *trace, cudaStream_t st);
// C = A B for the (N+1) x (N+1) derivative tables, k ascending, no FMA (row-major, ld = N+1); device pointers
cudaError_t launch_table_matmul(double *C, const double *A, const double *B, int ld, cudaStream_t st);
cudaError_t launch_fp64_peak(double *d_out, int iters, int blocks, int threads, cudaStream_t st);
cudaError_t launch_dmma_peak(double *d_out, int iters, int blocks, int threads, cudaStream_t st);

}  // namespace osb
