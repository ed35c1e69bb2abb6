// osb_chan_api.cu -- C ABI of the channel matrix path (include/osb.h, "channel matrix path").
//
// Host side of MAKE_CHANNEL_METRICS / INIT_CHANNEL_PROFILE / MAKE_DERIVATIVES: the grid,
// the profile and D1hat..D4 are formed once per (discretisation, N) with the reference's
// own loops and summation order (orrfdchan.f:177-183, 235-239, 305-342, 1183-1215;
// orrcolchan.f:904-944; orrncbl.f:520-545), then the interior blocks are uploaded.
// Everything per alpha / per Re runs in chan_kernels.cu.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "osb_internal.h"

using namespace osb;

// the context is defined in osb_api.cu; only these accessors are needed here
extern "C" {
int osb_device_count(const osb_ctx *ctx);
void *osb_stream(osb_ctx *ctx);
}
namespace osb {
int ctx_fail(osb_ctx *ctx, int code, const std::string &msg);
int ctx_device(const osb_ctx *ctx, int slot);
void ctx_count_launch(osb_ctx *ctx, int n);
cudaMemPool_t ctx_pool(const osb_ctx *ctx, int slot);
cudaStream_t ctx_stream(const osb_ctx *ctx, int slot);
}  // namespace osb

struct osb_chan {
  int disc = 0, N = 0, n = 0;
  int bl = 0;  // mapped boundary-layer operator (osb_chan_create_bl): tables are the metric-folded D2', D4'
  std::vector<double> eta, u, d1u, d2u, D2full, D4full;  // host copies, full (N+1) grid, row-major (i,j)
  struct Dev {              // the tables on one device of the context (every device holds its own copy)
    int device = 0;
    double *dD2 = nullptr, *dD4 = nullptr, *du = nullptr, *dd2u = nullptr;  // interior, column-major
    double2 *dBand = nullptr;  // FD only: [n][9] (D2, D4)(r, r-4+j), the band of the 9-diagonal operator
  };
  std::vector<Dev> dev;
  osb_chan() = default;
  osb_chan(const osb_chan &) = delete;
  osb_chan &operator=(const osb_chan &) = delete;
  ~osb_chan() {
    for (Dev &d : dev) {
      if (!d.dD2 && !d.dD4 && !d.du && !d.dd2u && !d.dBand) continue;
      cudaSetDevice(d.device);
      cudaFree(d.dD2); cudaFree(d.dD4); cudaFree(d.du); cudaFree(d.dd2u); cudaFree(d.dBand);
    }
  }
};

namespace {

#define CH_CUDA(ctx, call)                                                                \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      (void)cudaGetLastError();                                                           \
      return ctx_fail((ctx), OSB_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + \
                                             " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    }                                                                                     \
  } while (0)

inline double &at(std::vector<double> &M, int N, int i, int j) { return M[(size_t)i * (N + 1) + j]; }

// first-derivative matrix of orrfdchan (FiniteD1, orrfdchan.f:1183-1215): central differences on the given grid,
// one-sided in the first and last row
void fd_first_derivative(std::vector<double> &D, int N, const std::vector<double> &eta) {
  std::fill(D.begin(), D.end(), 0.0);
  at(D, N, 0, 0) = 1. / (eta[0] - eta[1]);
  at(D, N, 0, 1) = -1. / (eta[0] - eta[1]);
  at(D, N, N, N - 1) = 1. / (eta[N - 1] - eta[N]);
  at(D, N, N, N) = -1. / (eta[N - 1] - eta[N]);
  for (int i = 1; i <= N - 1; ++i) {
    at(D, N, i, i - 1) = 1. / (eta[i - 1] - eta[i + 1]);
    at(D, N, i, i + 1) = -1. / (eta[i - 1] - eta[i + 1]);
  }
}

// Chebyshev collocation first-derivative matrix on eta_j = cos(j pi / N) as CHEBYD forms it (orrcolchan.f:904-944)
void cheb_first_derivative(std::vector<double> &D, int N) {
  const double PI = acos(-1.0), fn = (double)N;
  std::vector<double> C(N + 1, 1.0);
  C[0] = 2.0; C[N] = 2.0;
  for (int j = 0; j <= N; ++j)
    for (int k = 0; k <= N; ++k) {
      double v;
      if (j == 0 && k == 0) v = (2.0 * (fn * fn) + 1.0) / 6.0;
      else if (j == N && k == N) v = -1.0 * (2. * (fn * fn) + 1.0) / 6.0;
      else if (j == k) {
        const double cj = cos((double)j * PI / fn);
        v = -1.0 * cj / (2. * (1. - cj * cj));
      } else {
        const double sgn = ((j + k) % 2 == 0) ? 1.0 : -1.0;
        v = C[j] * sgn / (C[k] * (cos((double)j * PI / fn) - cos((double)k * PI / fn)));
      }
      at(D, N, j, k) = v;
    }
}

// The products of MAKE_DERIVATIVES on the device (table_matmul_kernel: k ascending, no FMA = the reference's
// triple loops, orrfdchan.f:318-342).  Operands are uploaded once; results stay on the device for the next product.
class TableProducts {
 public:
  TableProducts(osb_ctx *ctx, int device, int N) : ctx_(ctx), device_(device), ld_(N + 1), st_(ctx_stream(ctx, 0)) {}
  ~TableProducts() {
    cudaSetDevice(device_);
    for (double *p : bufs_) cudaFree(p);
  }
  // device copy of a host table
  int upload(const std::vector<double> &h, double **d) {
    CH_CUDA(ctx_, cudaSetDevice(device_));
    CH_CUDA(ctx_, cudaMalloc(d, bytes()));
    bufs_.push_back(*d);
    CH_CUDA(ctx_, cudaMemcpyAsync(*d, h.data(), bytes(), cudaMemcpyHostToDevice, st_));
    return OSB_OK;
  }
  // *dC = dA dB on the device; host copy in hC when given
  int product(const double *dA, const double *dB, double **dC, std::vector<double> *hC) {
    CH_CUDA(ctx_, cudaSetDevice(device_));
    CH_CUDA(ctx_, cudaMalloc(dC, bytes()));
    bufs_.push_back(*dC);
    CH_CUDA(ctx_, launch_table_matmul(*dC, dA, dB, ld_, st_));
    ctx_count_launch(ctx_, 1);
    if (hC) {
      hC->resize((size_t)ld_ * ld_);
      CH_CUDA(ctx_, cudaMemcpyAsync(hC->data(), *dC, bytes(), cudaMemcpyDeviceToHost, st_));
    }
    return OSB_OK;
  }
  int sync() {
    CH_CUDA(ctx_, cudaStreamSynchronize(st_));
    return OSB_OK;
  }

 private:
  size_t bytes() const { return (size_t)ld_ * ld_ * sizeof(double); }
  osb_ctx *ctx_;
  int device_, ld_;
  cudaStream_t st_;
  std::vector<double *> bufs_;
};

// One worker thread per device of the context, each with its own contiguous share of a batch (the instances of a
// sweep cost the same to first order; every device holds its own copy of the tables).  fn(slot, lo, m) -> rc.
template <class F>
static int for_each_device_share(osb_ctx *ctx, int64_t nitems, F fn) {
  const int ndev = osb_device_count(ctx);
  if (ndev == 1 || nitems < 2 * (int64_t)ndev) return fn(0, (int64_t)0, nitems);
  const int64_t per = (nitems + ndev - 1) / ndev;
  std::vector<int> rcs(ndev, OSB_OK);
  std::vector<std::thread> th;
  for (int d = 0; d < ndev; ++d) {
    const int64_t lo = std::min<int64_t>(nitems, per * d), m = std::min<int64_t>(nitems, per * (d + 1)) - lo;
    if (m <= 0) continue;
    th.emplace_back([&, d, lo, m] { rcs[d] = fn(d, lo, m); });
  }
  for (auto &t : th) t.join();
  for (int d = 0; d < ndev; ++d) if (rcs[d]) return rcs[d];
  return OSB_OK;
}

}  // namespace

// interior blocks of D2, D4 (column-major n x n), profile rows and -- for the banded FD path -- the
// compact band table go to the device
static int upload_tables(osb_ctx *ctx, osb_chan *c) {
  const int N = c->N, disc = c->disc;
  const int n = c->n;
  std::vector<double> h2((size_t)n * n), h4((size_t)n * n), hu(n), hd(n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      h2[(size_t)j * n + i] = at(c->D2full, N, i + 1, j + 1);
      h4[(size_t)j * n + i] = at(c->D4full, N, i + 1, j + 1);
    }
  for (int i = 0; i < n; ++i) { hu[i] = c->u[i + 1]; hd[i] = c->d2u[i + 1]; }
  std::vector<double> hb;
  if (disc == OSB_DISC_FD) {
    hb.assign((size_t)n * 9 * 2, 0.0);
    for (int r = 0; r < n; ++r)
      for (int j = 0; j < 9; ++j) {
        const int cc = r - 4 + j;
        if (cc < 0 || cc >= n) continue;
        hb[((size_t)r * 9 + j) * 2] = at(c->D2full, N, r + 1, cc + 1);
        hb[((size_t)r * 9 + j) * 2 + 1] = at(c->D4full, N, r + 1, cc + 1);
      }
  }
  const int ndev = osb_device_count(ctx);
  c->dev.resize(ndev);
  for (int s = 0; s < ndev; ++s) {
    osb_chan::Dev &d = c->dev[s];
    d.device = ctx_device(ctx, s);
    CH_CUDA(ctx, cudaSetDevice(d.device));
    CH_CUDA(ctx, cudaMalloc(&d.dD2, h2.size() * sizeof(double)));
    CH_CUDA(ctx, cudaMalloc(&d.dD4, h4.size() * sizeof(double)));
    CH_CUDA(ctx, cudaMalloc(&d.du, n * sizeof(double)));
    CH_CUDA(ctx, cudaMalloc(&d.dd2u, n * sizeof(double)));
    CH_CUDA(ctx, cudaMemcpy(d.dD2, h2.data(), h2.size() * sizeof(double), cudaMemcpyHostToDevice));
    CH_CUDA(ctx, cudaMemcpy(d.dD4, h4.data(), h4.size() * sizeof(double), cudaMemcpyHostToDevice));
    CH_CUDA(ctx, cudaMemcpy(d.du, hu.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    CH_CUDA(ctx, cudaMemcpy(d.dd2u, hd.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    if (disc == OSB_DISC_FD) {
      CH_CUDA(ctx, cudaMalloc(&d.dBand, hb.size() * sizeof(double)));
      CH_CUDA(ctx, cudaMemcpy(d.dBand, hb.data(), hb.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
  }
  return OSB_OK;
}

extern "C" {

int osb_chan_create(osb_ctx *ctx, int disc, int N, osb_chan **out) {
  if (!ctx || !out) return OSB_E_ARG;
  if (disc != OSB_DISC_FD && disc != OSB_DISC_CHEB && disc != OSB_DISC_CHEB_NUM)
    return ctx_fail(ctx, OSB_E_ARG, "unknown discretisation");
  if (N < 8 || N > 1024) return ctx_fail(ctx, OSB_E_ARG, "N out of range [8,1024]");
  std::unique_ptr<osb_chan> c(new osb_chan);
  c->disc = disc; c->N = N; c->n = N - 1;
  const size_t n1 = (size_t)N + 1, nn = n1 * n1;
  c->eta.resize(n1); c->u.resize(n1); c->d1u.resize(n1); c->d2u.resize(n1);
  std::vector<double> D1hat(nn), D1(nn), D3(nn);
  c->D2full.resize(nn); c->D4full.resize(nn);
  const double pi = acos(-1.0), dth = pi / (double)N;  // orrfdchan.f:177-183
  for (int i = 0; i <= N; ++i) c->eta[i] = cos((double)i * dth);
  if (disc == OSB_DISC_FD) { fd_first_derivative(D1hat, N, c->eta); fd_first_derivative(D1, N, c->eta); }
  else { cheb_first_derivative(D1hat, N); cheb_first_derivative(D1, N); }
  for (int i = 0; i <= N; ++i) { at(D1, N, 0, i) = 0.0; at(D1, N, N, i) = 0.0; }  // phi' = 0
  TableProducts tp(ctx, ctx_device(ctx, 0), N);
  double *dD1hat = nullptr, *dD1 = nullptr, *dD2 = nullptr, *dD3 = nullptr, *dD4 = nullptr, *dD2hat = nullptr;
  std::vector<double> D2hat;
  int rc = tp.upload(D1hat, &dD1hat);
  if (!rc) rc = tp.upload(D1, &dD1);
  if (!rc) rc = tp.product(dD1hat, dD1, &dD2, &c->D2full);
  if (!rc) rc = tp.product(dD1hat, dD2, &dD3, nullptr);
  if (!rc) rc = tp.product(dD1hat, dD3, &dD4, &c->D4full);
  if (!rc && disc == OSB_DISC_CHEB_NUM) rc = tp.product(dD1hat, dD1hat, &dD2hat, &D2hat);
  if (!rc) rc = tp.sync();
  if (rc) return rc;
  for (int i = 0; i <= N; ++i) {  // INIT_CHANNEL_PROFILE
    c->u[i] = (1. - c->eta[i] * c->eta[i]);
    c->d1u[i] = -2.0 * c->eta[i];
    c->d2u[i] = -2.0;
  }
  if (disc == OSB_DISC_CHEB_NUM) {  // orrncbl.f:520-545
    for (int i = 0; i <= N; ++i) {
      double a = 0.0, b = 0.0;
      for (int k = 0; k <= N; ++k) { a = a + at(D1hat, N, i, k) * c->u[k]; b = b + at(D2hat, N, i, k) * c->u[k]; }
      c->d1u[i] = a; c->d2u[i] = b;
    }
  }
  rc = upload_tables(ctx, c.get());
  if (rc) return rc;
  *out = c.release();
  return OSB_OK;
}

static void bl_grid(int disc, int N, std::vector<double> &eta) {
  eta.resize((size_t)N + 1);
  if (disc == OSB_DISC_FD) {  // orrfdbl.f:258-265
    const double deta = 2. / (double)N;
    for (int i = 0; i <= N; ++i) eta[i] = 1. - (double)i * deta;
  } else {  // orrcolbl.f:290-296
    const double pi = acos(-1.0), dth = pi / (double)N;
    for (int i = 0; i <= N; ++i) eta[i] = cos((double)i * dth);
  }
}

int osb_chan_bl_points(int disc, int N, double lmap, double *y) {
  if ((disc != OSB_DISC_FD && disc != OSB_DISC_CHEB) || N < 8 || !y || !(lmap > 0.0)) return OSB_E_ARG;
  std::vector<double> eta;
  bl_grid(disc, N, eta);
  y[0] = INFINITY;
  for (int i = 1; i <= N; ++i) y[i] = lmap * (1. + eta[i]) / (1. - eta[i]);  // orrcolbl.f:391
  return OSB_OK;
}

int osb_chan_create_bl(osb_ctx *ctx, int disc, int N, double lmap, const double *U, int merge, osb_chan **out) {
  if (!ctx || !out || !U) return OSB_E_ARG;
  if (disc != OSB_DISC_FD && disc != OSB_DISC_CHEB) return ctx_fail(ctx, OSB_E_ARG, "disc must be OSB_DISC_FD or OSB_DISC_CHEB");
  if (N < 8 || N > 1024) return ctx_fail(ctx, OSB_E_ARG, "N out of range [8,1024]");
  if (!(lmap > 0.0)) return ctx_fail(ctx, OSB_E_ARG, "lmap <= 0");
  if (merge < 0 || merge >= N) return ctx_fail(ctx, OSB_E_ARG, "merge out of range [0,N)");
  std::unique_ptr<osb_chan> c(new osb_chan);
  c->disc = disc; c->N = N; c->n = N - 1; c->bl = 1;
  const size_t n1 = (size_t)N + 1, nn = n1 * n1;
  bl_grid(disc, N, c->eta);
  c->u.resize(n1); c->d1u.resize(n1); c->d2u.resize(n1);
  std::vector<double> D1hat(nn), D1(nn), D2(nn), D3(nn), D4(nn), D2hat(nn);
  if (disc == OSB_DISC_FD) { fd_first_derivative(D1hat, N, c->eta); fd_first_derivative(D1, N, c->eta); }
  else { cheb_first_derivative(D1hat, N); cheb_first_derivative(D1, N); }
  for (int i = 0; i <= N; ++i) { at(D1, N, 0, i) = 0.0; at(D1, N, N, i) = 0.0; }  // v' = 0 at both ends
  {
    TableProducts tp(ctx, ctx_device(ctx, 0), N);
    double *dD1hat = nullptr, *dD1 = nullptr, *dD2 = nullptr, *dD3 = nullptr, *dD4 = nullptr, *dD2hat = nullptr;
    int rcp = tp.upload(D1hat, &dD1hat);
    if (!rcp) rcp = tp.upload(D1, &dD1);
    if (!rcp) rcp = tp.product(dD1hat, dD1, &dD2, &D2);
    if (!rcp) rcp = tp.product(dD1hat, dD2, &dD3, &D3);
    if (!rcp) rcp = tp.product(dD1hat, dD3, &dD4, &D4);
    if (!rcp) rcp = tp.product(dD1hat, dD1hat, &dD2hat, &D2hat);
    if (!rcp) rcp = tp.sync();
    if (rcp) return rcp;
  }
  // INIT_BL_PROFILE with the caller's profile values (orrcolbl.f:398-427)
  for (int i = 0; i <= N; ++i) c->u[i] = (i <= merge) ? 1.0 : U[i];
  std::vector<double> d1u(n1), d2u(n1);
  for (int i = 0; i <= N; ++i) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k <= N; ++k) { a = a + at(D1hat, N, i, k) * c->u[k]; b = b + at(D2hat, N, i, k) * c->u[k]; }
    d1u[i] = a; d2u[i] = b;
  }
  for (int j = merge - 2; j >= 0; --j) { d1u[j] = 0.0; d2u[j] = 0.0; }
  // MAKE_BL_METRICS (orrcolbl.f:300-306) folded into the tables: the rows of the derivative matrices are
  // scaled once here instead of once per alpha in MAKE_MATRIX (orrcolbl.f:495-560)
  c->D2full.assign(nn, 0.0); c->D4full.assign(nn, 0.0);
  for (int i = 0; i <= N; ++i) {
    const double e1 = c->eta[i] - 1.;
    const double m1 = (e1 * e1) / (2. * lmap), m2 = (e1 * e1 * e1) / (2. * (lmap * lmap));
    const double m3 = 3. * (e1 * e1 * e1 * e1) / (4. * (lmap * lmap * lmap));
    const double m4 = 3. * (e1 * e1 * e1 * e1 * e1) / (2. * (lmap * lmap * lmap * lmap));
    const double g4 = m1 * m1 * m1 * m1, g3 = 6. * (m1 * m1) * m2, g2 = 3. * (m2 * m2) + 4. * m1 * m3, m1s = m1 * m1;
    for (int j = 0; j <= N; ++j) {
      at(c->D2full, N, i, j) = m1s * at(D2, N, i, j) + m2 * at(D1, N, i, j);
      at(c->D4full, N, i, j) = g4 * at(D4, N, i, j) + g3 * at(D3, N, i, j) + g2 * at(D2, N, i, j) + m4 * at(D1, N, i, j);
    }
    c->d1u[i] = d1u[i];
    c->d2u[i] = d2u[i] * m1s + d1u[i] * m2;  // the U'' of the mapped operator (orrcolbl.f:529-530)
  }
  int rc = upload_tables(ctx, c.get());
  if (rc) return rc;
  *out = c.release();
  return OSB_OK;
}

int osb_chan_free(osb_ctx *ctx, osb_chan *c) {
  if (!c) return OSB_E_ARG;
  delete c;
  (void)ctx;
  return OSB_OK;
}

int osb_chan_tables(osb_ctx *ctx, const osb_chan *c, double *eta, double *u, double *d2u, double *D2, double *D4) {
  if (!c) return OSB_E_ARG;
  const size_t n1 = (size_t)c->N + 1;
  if (eta) memcpy(eta, c->eta.data(), n1 * sizeof(double));
  if (u) memcpy(u, c->u.data(), n1 * sizeof(double));
  if (d2u) memcpy(d2u, c->d2u.data(), n1 * sizeof(double));
  if (D2) memcpy(D2, c->D2full.data(), n1 * n1 * sizeof(double));
  if (D4) memcpy(D4, c->D4full.data(), n1 * n1 * sizeof(double));
  (void)ctx;
  return OSB_OK;
}

// run one batch of instances (alpha, Re, sigma) -> (omega, delta, iters[, evec])
static int run_eig_chunk(osb_ctx *ctx, const osb_chan *c, int slot, int64_t ninst, const std::vector<double> &alpha,
                         const std::vector<double> &Re, const std::vector<double> &sigma, int outer, int inner,
                         double tol, std::vector<double> &omega, std::vector<double> &delta,
                         std::vector<int32_t> &iters, double *evec_host, double *adj_host = nullptr, int adj_form = 0) {
  cudaStream_t st = ctx_stream(ctx, slot);
  const osb_chan::Dev &cd = c->dev[slot];
  const int n = c->n;
  int grid = 0;
  size_t smem = 0;
  CH_CUDA(ctx, cudaSetDevice(cd.device));
  // the FD operator is banded (half bandwidth 4): banded LU, one thread per instance
  const bool banded = (c->disc == OSB_DISC_FD) && !getenv("OSB_FD_DENSE");
  if (!banded) {
    if (n > chan_dense_max_n())
      return ctx_fail(ctx, OSB_E_ARG, "N too large for the dense kernels (their 16-column LU panel lives in shared memory): "
                                      "limit N = " + std::to_string(chan_dense_max_n() + 1) +
                                      "; the FD operator has no such limit (banded kernels)");
  }
  if (!banded) CH_CUDA(ctx, chan_eig_grid(n, ninst, &grid, &smem));
  Scratch scr(cd.device, st, ctx_pool(ctx, slot));
  double *d_in = nullptr, *d_out = nullptr;
  double2 *work = nullptr, *d_evec = nullptr, *d_adj = nullptr;
  int32_t *d_it = nullptr;
  CH_CUDA(ctx, scr.alloc(&d_in, ninst * 5 * sizeof(double)));   // alpha(2), sigma(2), Re
  CH_CUDA(ctx, scr.alloc(&d_out, ninst * 3 * sizeof(double)));  // omega(2), delta
  CH_CUDA(ctx, scr.alloc(&d_it, ninst * sizeof(int32_t)));
  CH_CUDA(ctx, scr.alloc(&work, banded ? chan_band_work_bytes(n, ninst) : (size_t)grid * n * n * sizeof(double2)));
  if (evec_host) CH_CUDA(ctx, scr.alloc(&d_evec, (size_t)ninst * (n + 2) * sizeof(double2)));
  if (adj_host) CH_CUDA(ctx, scr.alloc(&d_adj, (size_t)ninst * (n + 2) * sizeof(double2)));
  CH_CUDA(ctx, cudaMemcpyAsync(d_in, alpha.data(), ninst * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  CH_CUDA(ctx, cudaMemcpyAsync(d_in + 2 * ninst, sigma.data(), ninst * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  CH_CUDA(ctx, cudaMemcpyAsync(d_in + 4 * ninst, Re.data(), ninst * sizeof(double), cudaMemcpyHostToDevice, st));
  if (banded)
    CH_CUDA(ctx, launch_chan_fd_banded(n, cd.dD2, cd.dD4, cd.du, cd.dd2u, cd.dBand, work, ninst, (const double2 *)d_in,
                                       d_in + 4 * ninst, (const double2 *)(d_in + 2 * ninst), outer, inner, tol,
                                       (double2 *)d_out, d_out + 2 * ninst, d_it, d_evec, d_adj, adj_form, st));
  else
    CH_CUDA(ctx, launch_chan_eig(n, cd.dD2, cd.dD4, cd.du, cd.dd2u, work, grid, smem, ninst,
                                 (const double2 *)d_in, d_in + 4 * ninst, (const double2 *)(d_in + 2 * ninst), outer,
                                 inner, tol, (double2 *)d_out, d_out + 2 * ninst, d_it, d_evec, d_adj, adj_form, st));
  ctx_count_launch(ctx, 1);
  omega.resize(2 * ninst); delta.resize(ninst); iters.resize(ninst);
  CH_CUDA(ctx, cudaMemcpyAsync(omega.data(), d_out, ninst * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CH_CUDA(ctx, cudaMemcpyAsync(delta.data(), d_out + 2 * ninst, ninst * sizeof(double), cudaMemcpyDeviceToHost, st));
  CH_CUDA(ctx, cudaMemcpyAsync(iters.data(), d_it, ninst * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (evec_host)
    CH_CUDA(ctx, cudaMemcpyAsync(evec_host, d_evec, (size_t)ninst * (n + 2) * sizeof(double2), cudaMemcpyDeviceToHost, st));
  if (adj_host)
    CH_CUDA(ctx, cudaMemcpyAsync(adj_host, d_adj, (size_t)ninst * (n + 2) * sizeof(double2), cudaMemcpyDeviceToHost, st));
  CH_CUDA(ctx, cudaStreamSynchronize(st));
  if (getenv("OSB_CHAN_PROFILE")) chan_profile_dump("eig");
  return OSB_OK;
}

// The banded path keeps 13 n complex factors per instance in HBM (123 KB at N = 512): batches are walked in
// chunks of at most ~8 GB of device scratch so that any batch size fits.
static int run_eig_slot(osb_ctx *ctx, const osb_chan *c, int slot, int64_t ninst, const std::vector<double> &alpha,
                        const std::vector<double> &Re, const std::vector<double> &sigma, int outer, int inner,
                        double tol, std::vector<double> &omega, std::vector<double> &delta,
                        std::vector<int32_t> &iters, double *evec_host, double *adj_host = nullptr, int adj_form = 0) {
  const bool banded = (c->disc == OSB_DISC_FD) && !getenv("OSB_FD_DENSE");
  int64_t chunk = ninst;
  if (banded) {
    const size_t per = chan_band_work_bytes(c->n, 1) + ((evec_host ? 1 : 0) + (adj_host ? 1 : 0)) * (size_t)(c->n + 2) * 16 + 64;
    chunk = std::max<int64_t>(1, (int64_t)((8ull << 30) / per));
    if (const char *e = getenv("OSB_CHAN_CHUNK")) chunk = std::max<int64_t>(1, atoll(e));  // test hook
  }
  if (ninst <= chunk)
    return run_eig_chunk(ctx, c, slot, ninst, alpha, Re, sigma, outer, inner, tol, omega, delta, iters, evec_host, adj_host, adj_form);
  omega.resize(2 * ninst); delta.resize(ninst); iters.resize(ninst);
  std::vector<double> a, r, sg, om, dl;
  std::vector<int32_t> it;
  for (int64_t lo = 0; lo < ninst; lo += chunk) {
    const int64_t m = std::min(chunk, ninst - lo);
    a.assign(alpha.begin() + 2 * lo, alpha.begin() + 2 * (lo + m));
    r.assign(Re.begin() + lo, Re.begin() + lo + m);
    sg.assign(sigma.begin() + 2 * lo, sigma.begin() + 2 * (lo + m));
    int rc = run_eig_chunk(ctx, c, slot, m, a, r, sg, outer, inner, tol, om, dl, it,
                           evec_host ? evec_host + (size_t)lo * (c->n + 2) * 2 : nullptr,
                           adj_host ? adj_host + (size_t)lo * (c->n + 2) * 2 : nullptr, adj_form);
    if (rc) return rc;
    std::copy(om.begin(), om.end(), omega.begin() + 2 * lo);
    std::copy(dl.begin(), dl.end(), delta.begin() + lo);
    std::copy(it.begin(), it.end(), iters.begin() + lo);
  }
  return OSB_OK;
}

// a batch of instances over all devices of the context
static int run_eig(osb_ctx *ctx, const osb_chan *c, int64_t ninst, const std::vector<double> &alpha,
                   const std::vector<double> &Re, const std::vector<double> &sigma, int outer, int inner,
                   double tol, std::vector<double> &omega, std::vector<double> &delta,
                   std::vector<int32_t> &iters, double *evec_host, double *adj_host = nullptr, int adj_form = 0) {
  if (osb_device_count(ctx) == 1)
    return run_eig_slot(ctx, c, 0, ninst, alpha, Re, sigma, outer, inner, tol, omega, delta, iters, evec_host, adj_host, adj_form);
  omega.resize(2 * ninst); delta.resize(ninst); iters.resize(ninst);
  const size_t ev_stride = (size_t)(c->n + 2) * 2;
  return for_each_device_share(ctx, ninst, [&](int slot, int64_t lo, int64_t m) -> int {
    std::vector<double> a(alpha.begin() + 2 * lo, alpha.begin() + 2 * (lo + m)), r(Re.begin() + lo, Re.begin() + lo + m),
        sg(sigma.begin() + 2 * lo, sigma.begin() + 2 * (lo + m)), om, dl;
    std::vector<int32_t> it;
    int rc = run_eig_slot(ctx, c, slot, m, a, r, sg, outer, inner, tol, om, dl, it, evec_host ? evec_host + (size_t)lo * ev_stride : nullptr,
                          adj_host ? adj_host + (size_t)lo * ev_stride : nullptr, adj_form);
    if (rc) return rc;
    std::copy(om.begin(), om.end(), omega.begin() + 2 * lo);
    std::copy(dl.begin(), dl.end(), delta.begin() + lo);
    std::copy(it.begin(), it.end(), iters.begin() + lo);
    return OSB_OK;
  });
}

// The reference keeps ONE eigenvalue out of the full spectrum:
//   orrfdchan.f:815-823   the largest Im(w) among |Im w| < 10 -- whatever its real part.
//                         For the FD operator this is sometimes a spurious mode with
//                         w_r < 0 (e.g. N=64, alpha=0.76: -2.144-0.0027i beats the TS
//                         mode 0.156-0.0034i), and the drop-in must return the same one.
//   orrncbl.f:897-905     the topmost mode with |w_r| < 1 (also what chan.inp's index 62
//                         denotes for orrcolchan: the two Im~3 modes have |w_r| ~ 21).
// Without the spectrum: a scan of shifts on the real w axis, each refined by re-shifted
// inverse iteration to its nearest eigenvalue; the settled candidate that the
// reference's rule would keep is the result.  16 shifts cover 0 < c_r < 1; the FD rule
// adds 32 shifts over -4 < w_r < 0, the |w_r| < 1 rule adds 8 over -1 < w_r < 0.
static std::vector<double> scan_list(int disc, double alpha_r, int bl = 0) {
  std::vector<double> s;
  for (int k = 0; k < 16; ++k) s.push_back(alpha_r * ((k + 0.5) / 16.0));
  if (bl) return s;  // boundary layer: discrete modes and the continuous branch all have 0 < c_r <= 1
  if (disc == OSB_DISC_FD) for (int k = 0; k < 32; ++k) s.push_back(-(k + 0.5) * 0.125);
  else for (int k = 0; k < 8; ++k) s.push_back(-(k + 0.5) * 0.125);
  return s;
}

static int scan_shifts(osb_ctx *ctx, const osb_chan *c, int64_t npts, const double *alpha, const double *Re,
                       std::vector<double> &sigma_out, std::vector<int32_t> &ok) {
  const int ns = (int)scan_list(c->disc, 1.0, c->bl).size();
  const int64_t ninst = npts * ns;
  std::vector<double> a(2 * ninst), r(ninst), s(2 * ninst), om, dl;
  std::vector<int32_t> it;
  for (int64_t p = 0; p < npts; ++p) {
    const std::vector<double> sh = scan_list(c->disc, alpha[2 * p], c->bl);
    for (int k = 0; k < ns; ++k) {
      const int64_t e = p * ns + k;
      a[2 * e] = alpha[2 * p]; a[2 * e + 1] = alpha[2 * p + 1];
      r[e] = Re[p];
      s[2 * e] = sh[k];
      s[2 * e + 1] = 0.0;
    }
  }
  int rc = run_eig(ctx, c, ninst, a, r, s, 3, 6, 1.0e-10, om, dl, it, nullptr);
  if (rc) return rc;
  sigma_out.assign(2 * npts, 0.0);
  ok.assign(npts, 0);
  for (int64_t p = 0; p < npts; ++p) {
    double best = -1.0e300;
    for (int k = 0; k < ns; ++k) {
      const int64_t e = p * ns + k;
      const double wr = om[2 * e], wi = om[2 * e + 1];
      if (!(std::isfinite(wr) && std::isfinite(wi))) continue;
      if (!(dl[e] <= 1.0e-8 * hypot(wr, wi))) continue;  // not settled on one mode
      if (c->disc == OSB_DISC_FD || c->bl) {
        if (!(fabs(wi) < 10.0) || (int)wr == 999) continue;  // orrfdchan.f:818-819 (orrcolbl.f:771-779: no filter)
      } else {
        if (!(fabs(wr) < 1.0)) continue;                      // orrncbl.f:898
      }
      if (wi > best) { best = wi; sigma_out[2 * p] = wr; sigma_out[2 * p + 1] = wi; ok[p] = 1; }
    }
  }
  return OSB_OK;
}

int osb_chan_eig(osb_ctx *ctx, const osb_chan *c, double Re, int64_t npts, const double *alpha,
                 const double *sigma0, double *omega, double *evec, double *evec_adj, int adj_form,
                 int32_t *iters, int32_t *status) {
  if (!ctx || !c || npts < 0 || !alpha || !omega || !status) return ctx_fail(ctx, OSB_E_ARG, "null argument");
  if (evec_adj && adj_form != OSB_ADJ_PENCIL && adj_form != OSB_ADJ_INVBA) return ctx_fail(ctx, OSB_E_ARG, "unknown adj_form");
  if (npts == 0) return OSB_OK;
  std::vector<double> Rev(npts, Re), sig(2 * npts);
  std::vector<int32_t> ok(npts, 1);
  std::vector<char> verified(npts, 0), unverified(npts, 0);
  if (sigma0) memcpy(sig.data(), sigma0, 2 * npts * sizeof(double));
  else {
    int rc = scan_shifts(ctx, c, npts, alpha, Rev.data(), sig, ok);
    if (rc) return rc;
    if (getenv("OSB_CHAN_SCAN_MISS")) std::fill(ok.begin(), ok.end(), 0);  // test hook: pretend the scan found nothing
    // The reference's rule runs over the FULL spectrum; the scan only sees the modes nearest to its shifts
    // (0 < c_r < 1 and -4 < w_r < 0), and e.g. the FD operator's spurious branch sits at w_r ~ -2 / alpha:
    // outside the windows for alpha < 0.5.  So the scan's choice is cross-checked against the device spectrum
    // (K4c) under the same rule: at every point where no shift settled, at the first and the last point of the
    // sweep, and on both sides of every jump of the chosen branch between neighbours.  If a checked point
    // disagrees, the spectrum's mode is taken there and the check is extended to the whole sweep (up to 512
    // points; beyond that the unchecked points are reported as OSB_ST_UNVERIFIED).
    const int n = c->n;
    const bool can_verify = n <= chan_dense_max_n() && !getenv("OSB_CHAN_NO_VERIFY");
    auto rule_ok = [&](double wr, double wi) {
      if (!(std::isfinite(wr) && std::isfinite(wi))) return false;
      if (c->disc == OSB_DISC_FD || c->bl) return fabs(wi) < 10.0 && (int)wr != 999;  // orrfdchan.f:818-819
      return fabs(wr) < 1.0;                                                          // orrncbl.f:898
    };
    // Cross-check of the points `pts`.  The device spectrum proposes the values the rule would prefer to the
    // scan's choice (larger Im w); a proposal only counts once a shift-invert solve started from it has settled
    // next to it -- the QR values of inv(B)A are not trustworthy by themselves for ill-conditioned pencils
    // (collocation at N >= 256; the spectrum kernel's own polish leaves those unsettled).
    auto verify = [&](const std::vector<int64_t> &pts, int *nfixed) -> int {
      *nfixed = 0;
      if (pts.empty()) return OSB_OK;
      std::vector<double> am(2 * pts.size()), sp((size_t)pts.size() * n * 2);
      std::vector<int32_t> sst(pts.size());
      for (size_t k = 0; k < pts.size(); ++k) { am[2 * k] = alpha[2 * pts[k]]; am[2 * k + 1] = alpha[2 * pts[k] + 1]; }
      int rcv = osb_chan_spectrum(ctx, c, Re, (int64_t)pts.size(), am.data(), sp.data(), sst.data());
      if (rcv) return rcv;
      const int kMaxCand = 6;
      std::vector<int64_t> cpt;           // candidate -> index into pts
      std::vector<double> ca, cr, cs;     // alpha, Re, shift of every candidate
      for (size_t k = 0; k < pts.size(); ++k) {
        const int64_t p = pts[k];
        const double *w = sp.data() + k * n * 2;
        const double floor_i = ok[p] ? sig[2 * p + 1] + 1.0e-7 * std::max(1.0, hypot(sig[2 * p], sig[2 * p + 1])) : -1.0e300;
        int taken = 0;
        for (int e = n - 1; e >= 0 && taken < kMaxCand; --e) {  // ascending in Im(w): walk from the top
          if (!rule_ok(w[2 * e], w[2 * e + 1]) || !(w[2 * e + 1] > floor_i)) continue;
          cpt.push_back((int64_t)k);
          ca.push_back(alpha[2 * p]); ca.push_back(alpha[2 * p + 1]);
          cr.push_back(Re);
          cs.push_back(w[2 * e]); cs.push_back(w[2 * e + 1]);
          ++taken;
        }
      }
      std::vector<double> om, dl;
      std::vector<int32_t> it;
      if (!cpt.empty()) {
        rcv = run_eig(ctx, c, (int64_t)cpt.size(), ca, cr, cs, 2, 6, 1.0e-10, om, dl, it, nullptr);
        if (rcv) return rcv;
      }
      std::vector<char> fixed(pts.size(), 0);
      for (size_t e = 0; e < cpt.size(); ++e) {  // candidates of a point are in descending Im: the first confirmed one wins
        const size_t k = (size_t)cpt[e];
        if (fixed[k]) continue;
        const int64_t p = pts[k];
        const double wr = om[2 * e], wi = om[2 * e + 1], mag = hypot(wr, wi);
        if (!rule_ok(wr, wi) || !(dl[e] <= 1.0e-8 * mag)) continue;
        if (hypot(wr - cs[2 * e], wi - cs[2 * e + 1]) > 1.0e-3 * mag) continue;  // wandered off to another mode
        if (ok[p] && !(wi > sig[2 * p + 1] + 1.0e-9 * std::max(1.0, mag))) continue;  // the scan's own mode again
        sig[2 * p] = wr; sig[2 * p + 1] = wi; ok[p] = 1;
        fixed[k] = 1;
        *nfixed += 1;
      }
      for (int64_t p : pts) verified[p] = 1;
      return OSB_OK;
    };
    if (can_verify) {
      std::vector<int64_t> pts;
      std::vector<char> want(npts, 0);
      for (int64_t p = 0; p < npts; ++p) if (!ok[p]) want[p] = 1;
      want[0] = 1; want[npts - 1] = 1;
      int njump = 0;
      for (int64_t p = 1; p < npts && njump < 8; ++p) {
        if (!ok[p] || !ok[p - 1]) continue;
        const double d = hypot(sig[2 * p] - sig[2 * p - 2], sig[2 * p + 1] - sig[2 * p - 1]);
        const double m = 0.5 * (hypot(sig[2 * p], sig[2 * p + 1]) + hypot(sig[2 * p - 2], sig[2 * p - 1]));
        if (d > 0.25 * m) { want[p] = 1; want[p - 1] = 1; ++njump; }
      }
      for (int64_t p = 0; p < npts; ++p) if (want[p]) pts.push_back(p);
      int nfixed = 0;
      rc = verify(pts, &nfixed);
      if (rc) return rc;
      if (nfixed > 0) {  // the scan missed the reference's mode somewhere: trust it nowhere
        std::vector<int64_t> rest;
        for (int64_t p = 0; p < npts; ++p) if (!verified[p]) rest.push_back(p);
        if (rest.size() <= 512) {
          rc = verify(rest, &nfixed);
          if (rc) return rc;
        } else {
          for (int64_t p : rest) unverified[p] = 1;
        }
      }
    }
  }
  std::vector<double> a(alpha, alpha + 2 * npts), om, dl;
  std::vector<int32_t> it;
  std::vector<double> ev, av;
  if (evec) ev.resize((size_t)npts * (c->n + 2) * 2);
  if (evec_adj) av.resize((size_t)npts * (c->n + 2) * 2);
  int rc = run_eig(ctx, c, npts, a, Rev, sig, 4, 8, 1.0e-14, om, dl, it, evec ? ev.data() : nullptr,
                   evec_adj ? av.data() : nullptr, adj_form);
  if (rc) return rc;
  for (int64_t p = 0; p < npts; ++p) {
    omega[2 * p] = om[2 * p]; omega[2 * p + 1] = om[2 * p + 1];
    if (iters) iters[p] = it[p];
    const bool fin = std::isfinite(om[2 * p]) && std::isfinite(om[2 * p + 1]);
    const bool conv = fin && dl[p] <= 1.0e-12 * hypot(om[2 * p], om[2 * p + 1]);
    status[p] = !fin ? OSB_ST_NONFINITE : ((conv && ok[p]) ? (unverified[p] ? OSB_ST_UNVERIFIED : OSB_ST_CONVERGED) : OSB_ST_MAXCOUNT);
  }
  if (evec) memcpy(evec, ev.data(), ev.size() * sizeof(double));
  if (evec_adj) memcpy(evec_adj, av.data(), av.size() * sizeof(double));
  return OSB_OK;
}

int osb_chan_spectrum(osb_ctx *ctx, const osb_chan *c, double Re, int64_t npts, const double *alpha,
                      double *spectrum, int32_t *status) {
  if (!ctx || !c || npts < 0 || !alpha || !spectrum || !status) return ctx_fail(ctx, OSB_E_ARG, "null argument");
  if (npts == 0) return OSB_OK;
  const int n = c->n;
  CH_CUDA(ctx, cudaSetDevice(c->dev[0].device));
  if (n > chan_dense_max_n())
    return ctx_fail(ctx, OSB_E_ARG, "N too large for the dense kernels (their 16-column LU panel lives in shared memory): "
                                    "limit N = " + std::to_string(chan_dense_max_n() + 1));
  // inv(B)A, Hessenberg, QR: every device of the context takes a contiguous share of the alphas
  {
    int rc = for_each_device_share(ctx, npts, [&](int slot, int64_t lo, int64_t m) -> int {
      cudaStream_t st = ctx_stream(ctx, slot);
      const osb_chan::Dev &cd = c->dev[slot];
      int grid = 0;
      size_t smem = 0;
      CH_CUDA(ctx, cudaSetDevice(cd.device));
      CH_CUDA(ctx, chan_spectrum_grid(n, m, &grid, &smem));
      Scratch scr(cd.device, st, ctx_pool(ctx, slot));
      double *d_in = nullptr;
      double2 *work = nullptr, *d_eig = nullptr;
      int32_t *d_st = nullptr;
      std::vector<double> Rev(m, Re);
      CH_CUDA(ctx, scr.alloc(&d_in, m * 3 * sizeof(double)));  // alpha(2), Re
      CH_CUDA(ctx, scr.alloc(&work, (size_t)grid * 2 * n * n * sizeof(double2)));
      CH_CUDA(ctx, scr.alloc(&d_eig, (size_t)m * n * sizeof(double2)));
      CH_CUDA(ctx, scr.alloc(&d_st, m * sizeof(int32_t)));
      CH_CUDA(ctx, cudaMemcpyAsync(d_in, alpha + 2 * lo, m * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
      CH_CUDA(ctx, cudaMemcpyAsync(d_in + 2 * m, Rev.data(), m * sizeof(double), cudaMemcpyHostToDevice, st));
      CH_CUDA(ctx, launch_chan_spectrum(n, cd.dD2, cd.dD4, cd.du, cd.dd2u, work, grid, smem, m, (const double2 *)d_in,
                                        d_in + 2 * m, d_eig, d_st, st));
      ctx_count_launch(ctx, 1);
      CH_CUDA(ctx, cudaMemcpyAsync(spectrum + (size_t)lo * n * 2, d_eig, (size_t)m * n * sizeof(double2), cudaMemcpyDeviceToHost, st));
      CH_CUDA(ctx, cudaMemcpyAsync(status + lo, d_st, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      CH_CUDA(ctx, cudaStreamSynchronize(st));
      return OSB_OK;
    });
    if (rc) return rc;
  }
  // Polish: eigenvalues of inv(B)A inherit its conditioning (the collocation pencil: ~1e-7 on the least stable
  // mode), shift-invert on the pencil itself does not.  Every QR value seeds one inverse iteration; the refined
  // value replaces it when the iteration settled next to its seed.
  {
    const int64_t ninst = npts * n;
    std::vector<double> a2(2 * ninst), r2(ninst, Re), s2(spectrum, spectrum + 2 * ninst), om, dl;
    std::vector<int32_t> it;
    for (int64_t p = 0; p < npts; ++p)
      for (int k = 0; k < n; ++k) { a2[2 * (p * n + k)] = alpha[2 * p]; a2[2 * (p * n + k) + 1] = alpha[2 * p + 1]; }
    int rc = run_eig(ctx, c, ninst, a2, r2, s2, 2, 6, 1.0e-13, om, dl, it, nullptr);
    if (rc) return rc;
    for (int64_t e = 0; e < ninst; ++e) {
      const double wr = om[2 * e], wi = om[2 * e + 1], qr = spectrum[2 * e], qi = spectrum[2 * e + 1];
      const double mag = hypot(qr, qi);
      if (std::isfinite(wr) && std::isfinite(wi) && dl[e] <= 1.0e-10 * hypot(wr, wi) && hypot(wr - qr, wi - qi) <= 1.0e-5 * mag) {
        spectrum[2 * e] = wr; spectrum[2 * e + 1] = wi;
      }
    }
  }
  // the reference sorts by Im(w) ascending, carrying the index (straight insertion = stable)
  std::vector<std::pair<double, double>> tmp(n);
  for (int64_t p = 0; p < npts; ++p) {
    double *w = spectrum + (size_t)p * n * 2;
    for (int k = 0; k < n; ++k) tmp[k] = {w[2 * k], w[2 * k + 1]};
    std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<double, double> &x, const std::pair<double, double> &y) {
      return x.second < y.second;
    });
    for (int k = 0; k < n; ++k) { w[2 * k] = tmp[k].first; w[2 * k + 1] = tmp[k].second; }
    status[p] = status[p] ? OSB_ST_MAXCOUNT : OSB_ST_CONVERGED;
  }
  return OSB_OK;
}

int osb_chan_neutral(osb_ctx *ctx, const osb_chan *c, int64_t nRe, const double *Re, const double *alpha0,
                     double sfac, double tol, int32_t maxit, double *alpha_out, double *omega, int32_t *steps,
                     int32_t *status, double *trace) {
  if (!ctx || !c || nRe < 0 || !Re || !alpha0 || !alpha_out || !omega || !steps || !status)
    return ctx_fail(ctx, OSB_E_ARG, "null argument");
  if (nRe == 0) return OSB_OK;
  if (maxit < 1) return ctx_fail(ctx, OSB_E_ARG, "maxit < 1");
  // first call of the reference: the topmost mode with |w_r| < 1 (orrncbl.f:897-905)
  std::vector<double> a(2 * nRe), sig;
  std::vector<int32_t> ok;
  for (int64_t p = 0; p < nRe; ++p) { a[2 * p] = alpha0[p]; a[2 * p + 1] = 0.0; }
  int rc = scan_shifts(ctx, c, nRe, a.data(), Re, sig, ok);
  if (rc) return rc;
  const int n = c->n;
  // the Newton searches are independent: every device of the context takes a contiguous share of the Reynolds numbers
  rc = for_each_device_share(ctx, nRe, [&](int slot, int64_t lo, int64_t m) -> int {
    cudaStream_t st = ctx_stream(ctx, slot);
    const osb_chan::Dev &cd = c->dev[slot];
    int grid = 0;
    size_t smem = 0;
    CH_CUDA(ctx, cudaSetDevice(cd.device));
    CH_CUDA(ctx, chan_neutral_grid(n, m, &grid, &smem));
    Scratch scr(cd.device, st, ctx_pool(ctx, slot));
    double *d_in = nullptr, *d_out = nullptr, *d_trace = nullptr;
    double2 *work = nullptr;
    int32_t *d_i = nullptr;
    CH_CUDA(ctx, scr.alloc(&d_in, m * 4 * sizeof(double)));   // sigma0(2), Re, alpha0
    CH_CUDA(ctx, scr.alloc(&d_out, m * 3 * sizeof(double)));  // omega(2), alpha
    CH_CUDA(ctx, scr.alloc(&d_i, m * 2 * sizeof(int32_t)));
    CH_CUDA(ctx, scr.alloc(&work, (size_t)grid * n * n * sizeof(double2)));
    if (trace) {
      CH_CUDA(ctx, scr.alloc(&d_trace, (size_t)m * maxit * 7 * sizeof(double)));
      CH_CUDA(ctx, cudaMemsetAsync(d_trace, 0, (size_t)m * maxit * 7 * sizeof(double), st));
    }
    CH_CUDA(ctx, cudaMemcpyAsync(d_in, sig.data() + 2 * lo, m * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CH_CUDA(ctx, cudaMemcpyAsync(d_in + 2 * m, Re + lo, m * sizeof(double), cudaMemcpyHostToDevice, st));
    CH_CUDA(ctx, cudaMemcpyAsync(d_in + 3 * m, alpha0 + lo, m * sizeof(double), cudaMemcpyHostToDevice, st));
    CH_CUDA(ctx, launch_chan_neutral(n, cd.dD2, cd.dD4, cd.du, cd.dd2u, work, grid, smem, m, d_in + 2 * m,
                                     d_in + 3 * m, (const double2 *)d_in, sfac, tol, maxit, d_out + 2 * m,
                                     (double2 *)d_out, d_i, d_i + m, d_trace, st));
    ctx_count_launch(ctx, 1);
    CH_CUDA(ctx, cudaMemcpyAsync(omega + 2 * lo, d_out, m * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CH_CUDA(ctx, cudaMemcpyAsync(alpha_out + lo, d_out + 2 * m, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    CH_CUDA(ctx, cudaMemcpyAsync(steps + lo, d_i, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CH_CUDA(ctx, cudaMemcpyAsync(status + lo, d_i + m, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (trace)
      CH_CUDA(ctx, cudaMemcpyAsync(trace + (size_t)lo * maxit * 7, d_trace, (size_t)m * maxit * 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CH_CUDA(ctx, cudaStreamSynchronize(st));
    return OSB_OK;
  });
  if (rc) return rc;
  for (int64_t p = 0; p < nRe; ++p)
    if (!ok[p] && status[p] == OSB_ST_CONVERGED) status[p] = OSB_ST_MAXCOUNT;
  return OSB_OK;
}

}  // extern "C"
