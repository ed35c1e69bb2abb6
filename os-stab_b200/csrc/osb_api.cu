// osb_api.cu -- the C ABI of libosb.so (declared in include/osb.h).
//
// Host-side glue only: argument checking, device buffers, stream-ordered copies
// and kernel launches.  There is deliberately no CPU implementation of any
// solver here: with no CUDA device every entry point fails with OSB_E_CUDA.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <thread>

#include "osb_internal.h"

using namespace osb;

struct osb_ctx {
  std::vector<int> devices;
  std::vector<cudaStream_t> streams;
  std::vector<cudaMemPool_t> pools;  // private stream-ordered pools, one per device (Scratch)
  std::string err;
  std::mutex mu;                     // guards err: multi-device calls run one worker thread per device
  std::atomic<int64_t> launches{0};
};

struct osb_profile {
  std::vector<ProfileDev> dev;  // one per context device
  osb_profile() = default;
  osb_profile(const osb_profile &) = delete;
  osb_profile &operator=(const osb_profile &) = delete;
  ~osb_profile() {
    for (auto &d : dev)
      if (d.tables) {
        cudaSetDevice(d.device);
        cudaFree(d.tables);
      }
  }
};

namespace {

int fail(osb_ctx *ctx, int code, const std::string &msg) {
  if (ctx) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err = msg;
  }
  return code;
}

#define OSB_CUDA(ctx, call)                                                              \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      (void)cudaGetLastError(); /* do not let it surface again at the next launch check */ \
      return fail((ctx), OSB_E_CUDA,                                                     \
                  std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + \
                      std::to_string(__LINE__) + ")");                                   \
    }                                                                                    \
  } while (0)

// Cash-Karp tableau as typed in the reference (shoot.f:1381-1391)
const double kCKb[15] = {0.2,
                         0.075, 0.225,
                         0.3, -0.9, 1.2,
                         -0.2037037037037037, 2.5, -2.5925925925925926, 1.2962962962962963,
                         0.029495804398148147, 0.341796875, 0.041594328703703706,
                         0.40034541377314814, 0.061767578125};
const double kCKc[6] = {0.09788359788359788, 0.0, 0.4025764895330113,
                        0.21043771043771045, 0.0, 0.2891022021456804};

struct FlowRules {
  int maxcount;
  double errtol, eigtol;
  int loop_rule, cm1_defined, strict, first_perturb, show_updated, bl_ic, spatial, analytic;
};

bool flow_rules(int flow, FlowRules *r) {
  switch (flow) {
    case OSB_FLOW_CONTEBL:  // shoot.f:461-463,478-479,602,718
      *r = FlowRules{20, 1.0e-14, 1.0e-14, 0, 1, 1, 0, 0, 1, 0, 0};
      return true;
    case OSB_FLOW_CONTE:  // shoot.f:53-55,76-77,144,294,322
      *r = FlowRules{20, 1.0e-15, 1.0e-15, 1, 1, 0, 0, 1, 0, 0, 0};
      return true;
    case OSB_FLOW_ORRSOM:  // orrsom.f:203-204,302,422
      *r = FlowRules{20, 1.0e-8, 1.0e-12, 1, 0, 0, 0, 0, 1, 0, 0};
      return true;
    case OSB_FLOW_ORRSPACE:  // orrspace.f:277-278,357-358,459
      *r = FlowRules{50, 1.0e-8, 1.0e-12, 1, 0, 1, 1, 0, 1, 1, 1};
      return true;
  }
  return false;
}

// threshold of the orthonormality test in the form the kernel evaluates it
double norm_threshold(int flow, double testalpha) {
  const double pi = acos(-1.0);
  switch (flow) {
    case OSB_FLOW_CONTEBL: { double c = cos(pi * testalpha / 180.0); return c * c; }  // shoot.f:594
    case OSB_FLOW_CONTE: { double c = cos(10.0 * pi / 180.0); return c * c; }          // shoot.f:64-65
    case OSB_FLOW_ORRSOM: { double c = cos(testalpha * pi / 180.0); return c * c; }    // orrsom.f:292
    default: return testalpha * testalpha;                                             // orrspace.f:357
  }
}

int fill_args(osb_ctx *ctx, const osb_shoot_opts *o, ShootArgs *a, int *analytic) {
  FlowRules r;
  if (!o || !flow_rules(o->flow, &r)) return fail(ctx, OSB_E_ARG, "unknown flow");
  if (o->integrator != OSB_INT_RK4 && o->integrator != OSB_INT_CK45)
    return fail(ctx, OSB_E_ARG, "unknown integrator");
  if (o->nstep < 1) return fail(ctx, OSB_E_ARG, "nstep < 1");
  if (!(o->ymax > o->ymin)) return fail(ctx, OSB_E_ARG, "ymax <= ymin");
  memset(a, 0, sizeof(*a));
  a->nstep = o->nstep;
  a->maxcount = o->maxcount > 0 ? o->maxcount : r.maxcount;
  a->errtol = o->errtol > 0 ? o->errtol : r.errtol;
  a->eigtol = o->eigtol > 0 ? o->eigtol : r.eigtol;
  a->niter = 1;
  a->spatial = r.spatial;
  a->bl_ic = r.bl_ic;
  a->loop_rule = r.loop_rule;
  a->cm1_defined = r.cm1_defined;
  a->strict = r.strict;
  a->first_perturb = r.first_perturb;
  a->show_updated = r.show_updated;
  a->flow = o->flow;
  a->testalpha = o->testalpha;
  a->thr = norm_threshold(o->flow, o->testalpha);
  const double h = (o->ymax - o->ymin) / (double)o->nstep;  // shoot.f:471
  a->hs = (o->flow == OSB_FLOW_CONTE) ? h : -h;
  a->tf = o->ymax;
  for (int i = 0; i < 15; ++i) a->bh[i] = kCKb[i] * a->hs;
  for (int i = 0; i < 6; ++i) a->ch[i] = kCKc[i] * a->hs;
  *analytic = r.analytic;
  return OSB_OK;
}

int check_profile(osb_ctx *ctx, const osb_shoot_opts *o, const osb_profile *prof) {
  if (!prof || prof->dev.size() != ctx->devices.size())
    return fail(ctx, OSB_E_ARG, "profile does not belong to this context");
  const ProfileDev &p = prof->dev[0];
  if (o->flow == OSB_FLOW_CONTE) {
    if (p.variant != OSB_FLOW_CONTE) return fail(ctx, OSB_E_ARG, "conte flow needs the channel profile");
  } else {
    if (p.variant == OSB_FLOW_CONTE || !p.tables) return fail(ctx, OSB_E_ARG, "BL flow needs Blasius tables");
  }
  return OSB_OK;
}

// split [0,n) into ndev contiguous chunks
void chunk(int64_t n, int ndev, int d, int64_t *lo, int64_t *hi) {
  int64_t per = (n + ndev - 1) / ndev;
  *lo = std::min<int64_t>(n, per * d);
  *hi = std::min<int64_t>(n, per * (d + 1));
}

// pass counts vary by region of a sweep, so contiguous blocks would imbalance the devices: tiles of the
// flattened unit index (points, continuation lines) are dealt cyclically (tile t -> device t mod ndev) and packed
// contiguously in each device's buffers.  Per-unit results do not depend on the partition.
struct Seg { int64_t host_lo, n, dev_off; };
struct Deal { std::vector<std::vector<Seg>> segs; std::vector<int64_t> count; };
Deal deal_tiles(int64_t n, int ndev, int64_t tile) {
  Deal d;
  d.segs.resize(ndev);
  d.count.assign(ndev, 0);
  if (ndev == 1) tile = n;
  int64_t t = 0;
  for (int64_t lo = 0; lo < n; lo += tile, ++t) {
    const int dev = (int)(t % ndev);
    const int64_t m = std::min(tile, n - lo);
    d.segs[dev].push_back(Seg{lo, m, d.count[dev]});
    d.count[dev] += m;
  }
  return d;
}

}  // namespace

// accessors for the other translation units of the library
namespace osb {
int ctx_fail(osb_ctx *ctx, int code, const std::string &msg) { return fail(ctx, code, msg); }
int ctx_device(const osb_ctx *ctx, int slot) { return ctx->devices[slot]; }
void ctx_count_launch(osb_ctx *ctx, int n) { ctx->launches += n; }
cudaMemPool_t ctx_pool(const osb_ctx *ctx, int slot) { return ctx->pools[slot]; }
cudaStream_t ctx_stream(const osb_ctx *ctx, int slot) { return ctx->streams[slot]; }
}  // namespace osb

extern "C" {

int osb_version(void) { return OSB_VERSION; }

int osb_create_multi(osb_ctx **out, int ndev, const int *devices) {
  if (!out) return OSB_E_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return OSB_E_CUDA;
  std::unique_ptr<osb_ctx> c(new osb_ctx);
  if (ndev <= 0) {
    for (int d = 0; d < count; ++d) c->devices.push_back(d);
  } else {
    for (int i = 0; i < ndev; ++i) {
      int d = devices ? devices[i] : i;
      if (d < 0 || d >= count) return OSB_E_ARG;
      c->devices.push_back(d);
    }
  }
  for (int d : c->devices) {
    if (cudaSetDevice(d) != cudaSuccess) return OSB_E_CUDA;
    cudaStream_t s;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return OSB_E_CUDA;
    c->streams.push_back(s);
    // A private pool per context and device keeps stream-ordered scratch cached between calls (a pool
    // hands memory back to the driver at every synchronisation unless its release threshold says
    // otherwise) without changing the device's DEFAULT pool, which other allocators of the process share.
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = d;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
      for (size_t i = 0; i < c->streams.size(); ++i) {
        cudaSetDevice(c->devices[i]);
        cudaStreamDestroy(c->streams[i]);
        if (i < c->pools.size()) cudaMemPoolDestroy(c->pools[i]);
      }
      return OSB_E_CUDA;
    }
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    c->pools.push_back(pool);
  }
  *out = c.release();
  return OSB_OK;
}

int osb_create(osb_ctx **out, int device) { return osb_create_multi(out, 1, &device); }

int osb_destroy(osb_ctx *ctx) {
  if (!ctx) return OSB_E_ARG;
  for (size_t i = 0; i < ctx->devices.size(); ++i) {
    cudaSetDevice(ctx->devices[i]);
    cudaStreamSynchronize(ctx->streams[i]);
    cudaStreamDestroy(ctx->streams[i]);
    cudaMemPoolTrimTo(ctx->pools[i], 0);
    cudaMemPoolDestroy(ctx->pools[i]);
  }
  delete ctx;
  return OSB_OK;
}

int osb_device_count(const osb_ctx *ctx) { return ctx ? (int)ctx->devices.size() : 0; }
const char *osb_last_error(const osb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void *osb_stream(osb_ctx *ctx) { return ctx && !ctx->streams.empty() ? (void *)ctx->streams[0] : nullptr; }
int64_t osb_launch_count(const osb_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

int osb_shoot_defaults(int flow, osb_shoot_opts *o) {
  FlowRules r;
  if (!o || !flow_rules(flow, &r)) return OSB_E_ARG;
  memset(o, 0, sizeof(*o));
  o->flow = flow;
  o->maxcount = r.maxcount;
  o->errtol = r.errtol;
  o->eigtol = r.eigtol;
  switch (flow) {
    case OSB_FLOW_CONTEBL:  // contebl.f:64-83, built with USE_RKCK45 (README.md:60-62)
      o->integrator = OSB_INT_CK45; o->nstep = 1000; o->testalpha = 10.0; o->ymin = 0.0; o->ymax = 17.0;
      break;
    case OSB_FLOW_CONTE:  // conte.f:61-68
      o->integrator = OSB_INT_CK45; o->nstep = 2000; o->testalpha = 10.0; o->ymin = -1.0; o->ymax = 1.0;
      break;
    case OSB_FLOW_ORRSOM:  // test.inp
      o->integrator = OSB_INT_RK4; o->nstep = 2000; o->testalpha = 10.0; o->ymin = 0.0; o->ymax = 17.0;
      break;
    default:  // sweep.inp (ymax is rescaled by the driver, orrspace.f:160-167)
      o->integrator = OSB_INT_RK4; o->nstep = 1000; o->testalpha = 0.9; o->ymin = 0.0; o->ymax = 17.0;
      break;
  }
  return OSB_OK;
}

// ---------------------------------------------------------------- profiles
static int profile_alloc(osb_ctx *ctx, int variant, int nbl, double ymin, double ymax,
                         std::unique_ptr<osb_profile> *out) {
  std::unique_ptr<osb_profile> p(new osb_profile);
  p->dev.resize(ctx->devices.size());
  for (size_t i = 0; i < ctx->devices.size(); ++i) {
    ProfileDev &d = p->dev[i];
    d.variant = variant; d.nbl = nbl; d.ymin = ymin; d.ymax = ymax; d.device = ctx->devices[i];
    if (variant != OSB_FLOW_CONTE) {
      OSB_CUDA(ctx, cudaSetDevice(d.device));
      OSB_CUDA(ctx, cudaMalloc(&d.tables, 5 * ((size_t)nbl + 2) * sizeof(double)));
    }
  }
  *out = std::move(p);
  return OSB_OK;
}

int osb_profile_blasius(osb_ctx *ctx, int variant, int integrator, int nbl, double ymin,
                        double ymax, double f2p, double beta, osb_profile **out) {
  if (!ctx || !out) return OSB_E_ARG;
  if (variant != OSB_FLOW_CONTEBL && variant != OSB_FLOW_ORRSOM && variant != OSB_FLOW_ORRSPACE)
    return fail(ctx, OSB_E_ARG, "variant must be CONTEBL, ORRSOM or ORRSPACE");
  if (nbl < 8) return fail(ctx, OSB_E_ARG, "nbl < 8");
  std::unique_ptr<osb_profile> p;
  int rc = profile_alloc(ctx, variant, nbl, ymin, ymax, &p);
  if (rc) return rc;
  for (size_t i = 0; i < ctx->devices.size(); ++i) {
    ProfileDev &d = p->dev[i];
    cudaStream_t st = ctx->streams[i];
    OSB_CUDA(ctx, cudaSetDevice(d.device));
    Scratch scr(d.device, st, ctx->pools[i]);
    double *work = nullptr, *par = nullptr;
    int32_t *np = nullptr;
    OSB_CUDA(ctx, scr.alloc(&work, blasius_scratch_doubles(nbl, 1) * sizeof(double)));
    OSB_CUDA(ctx, scr.alloc(&par, 3 * sizeof(double)));
    OSB_CUDA(ctx, scr.alloc(&np, sizeof(int32_t)));
    double hp[2] = {f2p, beta};
    OSB_CUDA(ctx, cudaMemcpyAsync(par, hp, sizeof(hp), cudaMemcpyHostToDevice, st));
    OSB_CUDA(ctx, launch_blasius(variant, integrator, nbl, ymin, ymax, 1, par, par + 1, d.tables, work,
                                 par + 2, np, st));
    ctx->launches++;
    int32_t hnp = 0;
    double hf2 = 0;
    OSB_CUDA(ctx, cudaMemcpyAsync(&hnp, np, sizeof(hnp), cudaMemcpyDeviceToHost, st));
    OSB_CUDA(ctx, cudaMemcpyAsync(&hf2, par + 2, sizeof(hf2), cudaMemcpyDeviceToHost, st));
    OSB_CUDA(ctx, cudaStreamSynchronize(st));
    d.npass = hnp; d.f2p_last = hf2;
  }
  *out = p.release();
  return OSB_OK;
}

// one device's share [lo, lo+m) of a profile batch: enqueue everything on that device's stream (no synchronisation)
static int blasius_batch_slice(osb_ctx *ctx, int slot, Scratch &scr, int variant, int integrator, int nbl, double ymin,
                               double ymax, int64_t lo, int64_t m, const double *f2p, const double *beta, double *U,
                               double *Uspl, double *d2U, double *d2Uspl, double *f2p_out, int32_t *npass) {
  cudaStream_t st = ctx->streams[slot];
  OSB_CUDA(ctx, cudaSetDevice(ctx->devices[slot]));
  const size_t n2 = (size_t)nbl + 2, n1 = (size_t)nbl + 1;
  double *tab = nullptr, *work = nullptr, *par = nullptr;
  int32_t *np = nullptr;
  OSB_CUDA(ctx, scr.alloc(&tab, 5 * n2 * m * sizeof(double)));
  OSB_CUDA(ctx, scr.alloc(&work, blasius_scratch_doubles(nbl, m) * sizeof(double)));
  OSB_CUDA(ctx, scr.alloc(&par, 3 * m * sizeof(double)));
  OSB_CUDA(ctx, scr.alloc(&np, m * sizeof(int32_t)));
  OSB_CUDA(ctx, cudaMemcpyAsync(par, f2p + lo, m * sizeof(double), cudaMemcpyHostToDevice, st));
  OSB_CUDA(ctx, cudaMemcpyAsync(par + m, beta + lo, m * sizeof(double), cudaMemcpyHostToDevice, st));
  OSB_CUDA(ctx, launch_blasius(variant, integrator, nbl, ymin, ymax, m, par, par + m, tab, work, par + 2 * m, np, st));
  ctx->launches += 2;  // spline_grid_kernel + blasius_kernel
  // device layout is [array][i][p] (profile-fastest, coalesced for the march); the caller wants [p][i]:
  // transposed on the device, then one contiguous copy per requested table
  double *outs[4] = {U, Uspl, d2U, d2Uspl};
  if (U || Uspl || d2U || d2Uspl) {
    double *tr = nullptr;
    OSB_CUDA(ctx, scr.alloc(&tr, 4 * n1 * m * sizeof(double)));
    OSB_CUDA(ctx, launch_transpose_tables(tab, tr, nbl, m, st));
    ctx->launches++;
    for (int a = 0; a < 4; ++a)
      if (outs[a])
        OSB_CUDA(ctx, cudaMemcpyAsync(outs[a] + (size_t)lo * n1, tr + (size_t)a * n1 * m, n1 * m * sizeof(double),
                                      cudaMemcpyDeviceToHost, st));
  }
  if (f2p_out) OSB_CUDA(ctx, cudaMemcpyAsync(f2p_out + lo, par + 2 * m, m * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (npass) OSB_CUDA(ctx, cudaMemcpyAsync(npass + lo, np, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  return OSB_OK;
}

int osb_profile_blasius_batch(osb_ctx *ctx, int variant, int integrator, int nbl, double ymin,
                              double ymax, int64_t nprof, const double *f2p, const double *beta,
                              double *U, double *Uspl, double *d2U, double *d2Uspl,
                              double *f2p_out, int32_t *npass) {
  if (!ctx || nprof < 1 || !f2p || !beta) return OSB_E_ARG;
  if (variant != OSB_FLOW_CONTEBL && variant != OSB_FLOW_ORRSOM && variant != OSB_FLOW_ORRSPACE)
    return fail(ctx, OSB_E_ARG, "variant must be CONTEBL, ORRSOM or ORRSPACE");
  if (nbl < 8) return fail(ctx, OSB_E_ARG, "nbl < 8");
  // profiles cost the same to first order: contiguous shares, one per device, all enqueued before the first wait
  const int ndev = (int)ctx->devices.size();
  std::vector<std::unique_ptr<Scratch>> guards(ndev);
  for (int d = 0; d < ndev; ++d) {
    int64_t lo, hi;
    chunk(nprof, ndev, d, &lo, &hi);
    if (hi <= lo) continue;
    guards[d].reset(new Scratch(ctx->devices[d], ctx->streams[d], ctx->pools[d]));
    int rc = blasius_batch_slice(ctx, d, *guards[d], variant, integrator, nbl, ymin, ymax, lo, hi - lo, f2p, beta, U, Uspl, d2U,
                                 d2Uspl, f2p_out, npass);
    if (rc) return rc;
  }
  for (int d = 0; d < ndev; ++d) {
    OSB_CUDA(ctx, cudaSetDevice(ctx->devices[d]));
    OSB_CUDA(ctx, cudaStreamSynchronize(ctx->streams[d]));
  }
  return OSB_OK;
}

int osb_profile_upload(osb_ctx *ctx, int variant, int nbl, double ymin, double ymax,
                       const double *U, const double *Uspl, const double *d2U,
                       const double *d2Uspl, osb_profile **out) {
  if (!ctx || !out || !U || !Uspl) return OSB_E_ARG;
  if (variant != OSB_FLOW_CONTEBL && variant != OSB_FLOW_ORRSOM && variant != OSB_FLOW_ORRSPACE)
    return fail(ctx, OSB_E_ARG, "variant must be CONTEBL, ORRSOM or ORRSPACE");
  if (variant != OSB_FLOW_CONTEBL && (!d2U || !d2Uspl))
    return fail(ctx, OSB_E_ARG, "orrsom/orrspace need d2U and d2Uspl");
  const size_t n2 = (size_t)nbl + 2;
  std::vector<double> h(5 * n2, 0.0);
  const double hh = (ymax - ymin) / (double)nbl;
  for (int i = 0; i <= nbl; ++i) {
    h[i] = ymin + i * hh;  // ydat(i) = ymin + i*h  contebl.f:499
    h[n2 + i] = U[i];
    h[2 * n2 + i] = Uspl[i];
    h[3 * n2 + i] = d2U ? d2U[i] : Uspl[i];
    h[4 * n2 + i] = d2Uspl ? d2Uspl[i] : 0.0;
  }
  std::unique_ptr<osb_profile> p;
  int rc = profile_alloc(ctx, variant, nbl, ymin, ymax, &p);
  if (rc) return rc;
  for (size_t i = 0; i < ctx->devices.size(); ++i) {
    OSB_CUDA(ctx, cudaSetDevice(p->dev[i].device));
    OSB_CUDA(ctx, cudaMemcpy(p->dev[i].tables, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  *out = p.release();
  return OSB_OK;
}

int osb_profile_shift(osb_ctx *ctx, const osb_profile *prof, double uc, osb_profile **out) {
  if (!ctx || !prof || !out || prof->dev.size() != ctx->devices.size()) return OSB_E_ARG;
  const ProfileDev &p0 = prof->dev[0];
  if (!p0.tables) return fail(ctx, OSB_E_ARG, "profile has no tables");
  std::unique_ptr<osb_profile> q;
  int rc = profile_alloc(ctx, p0.variant, p0.nbl, p0.ymin, p0.ymax, &q);
  if (rc) return rc;
  for (size_t i = 0; i < ctx->devices.size(); ++i) {
    cudaStream_t st = ctx->streams[i];
    OSB_CUDA(ctx, cudaSetDevice(q->dev[i].device));
    Scratch scr(q->dev[i].device, st, ctx->pools[i]);
    double *work = nullptr;
    OSB_CUDA(ctx, scr.alloc(&work, 2 * ((size_t)p0.nbl + 2) * sizeof(double)));
    OSB_CUDA(ctx, launch_shift_profile(p0.nbl, uc, prof->dev[i].tables, q->dev[i].tables, work, st));
    ctx->launches++;
    OSB_CUDA(ctx, cudaStreamSynchronize(st));
    q->dev[i].npass = p0.npass; q->dev[i].f2p_last = p0.f2p_last;
  }
  *out = q.release();
  return OSB_OK;
}

int osb_profile_channel(osb_ctx *ctx, osb_profile **out) {
  if (!ctx || !out) return OSB_E_ARG;
  std::unique_ptr<osb_profile> p;
  int rc = profile_alloc(ctx, OSB_FLOW_CONTE, 0, -1.0, 1.0, &p);
  if (rc) return rc;
  *out = p.release();
  return OSB_OK;
}

int osb_profile_get(osb_ctx *ctx, const osb_profile *prof, double *ydat, double *U, double *Uspl,
                    double *d2U, double *d2Uspl, int32_t *npass, double *f2p_last) {
  if (!ctx || !prof || prof->dev.empty()) return OSB_E_ARG;
  const ProfileDev &d = prof->dev[0];
  if (!d.tables) return fail(ctx, OSB_E_STATE, "profile has no tables");
  const size_t n2 = (size_t)d.nbl + 2;
  double *outs[5] = {ydat, U, Uspl, d2U, d2Uspl};
  OSB_CUDA(ctx, cudaSetDevice(d.device));
  for (int a = 0; a < 5; ++a)
    if (outs[a])
      OSB_CUDA(ctx, cudaMemcpy(outs[a], d.tables + a * n2, ((size_t)d.nbl + 1) * sizeof(double), cudaMemcpyDeviceToHost));
  if (npass) *npass = d.npass;
  if (f2p_last) *f2p_last = d.f2p_last;
  return OSB_OK;
}

int osb_profile_nbl(const osb_profile *prof) { return prof && !prof->dev.empty() ? prof->dev[0].nbl : -1; }

int osb_profile_free(osb_ctx *ctx, osb_profile *prof) {
  if (!prof) return OSB_E_ARG;
  delete prof;
  (void)ctx;
  return OSB_OK;
}

// ---------------------------------------------------------------- shooting
// build the stage table for device slot i in the call's scratch
static int make_stage_table(osb_ctx *ctx, size_t i, Scratch &scr, const osb_shoot_opts *o,
                            const osb_profile *prof, double2 **tab) {
  cudaStream_t st = ctx->streams[i];
  size_t n = (size_t)o->nstep * stages_per_step(o->integrator);
  OSB_CUDA(ctx, scr.alloc(tab, n * sizeof(double2)));  // released with the call's other scratch
  OSB_CUDA(ctx, launch_stage_table(prof->dev[i], o->flow, o->integrator, o->nstep, o->ymin, o->ymax, *tab, st));
  ctx->launches++;
  return OSB_OK;
}

int osb_shoot_temporal_dev(osb_ctx *ctx, const osb_shoot_opts *opts, const osb_profile *prof,
                           int64_t npts, const double *d_alpha, const double *d_Re, double *d_c,
                           int32_t *d_passes, int32_t *d_qend, int32_t *d_status, double *d_resid,
                           double *d_history, int32_t hist_cap) {
  if (!ctx || npts < 0 || !d_alpha || !d_Re || !d_c || !d_passes || !d_qend || !d_status)
    return fail(ctx, OSB_E_ARG, "null argument");
  if (opts && opts->flow == OSB_FLOW_ORRSPACE) return fail(ctx, OSB_E_ARG, "orrspace is spatial: use osb_shoot_spatial_lines");
  ShootArgs a;
  int analytic = 0;
  int rc = fill_args(ctx, opts, &a, &analytic);
  if (rc) return rc;
  rc = check_profile(ctx, opts, prof);
  if (rc) return rc;
  if (npts == 0) return OSB_OK;
  OSB_CUDA(ctx, cudaSetDevice(ctx->devices[0]));
  cudaStream_t st = ctx->streams[0];
  Scratch scr(ctx->devices[0], st, ctx->pools[0]);
  double2 *tab = nullptr;
  rc = make_stage_table(ctx, 0, scr, opts, prof, &tab);
  if (rc) return rc;
  a.tab = tab;
  a.nlines = npts;
  a.alpha = (const double2 *)d_alpha;
  a.Re = d_Re;
  a.ev = (double2 *)d_c;
  a.passes = d_passes; a.qend = d_qend; a.status = d_status;
  a.resid = (double2 *)d_resid;
  a.history = (double4 *)d_history;
  a.hist_cap = d_history ? hist_cap : 0;
  int nl = 0;
  unsigned long long *qc = nullptr;
  if (!getenv("OSB_NO_QUEUE")) {
    OSB_CUDA(ctx, scr.alloc(&qc, sizeof(unsigned long long)));
    OSB_CUDA(ctx, cudaMemsetAsync(qc, 0, sizeof(unsigned long long), st));
  }
  OSB_CUDA(ctx, launch_shoot(a, opts->integrator, analytic, st, &nl, qc));
  ctx->launches += nl;
  return OSB_OK;
}

int osb_shoot_temporal(osb_ctx *ctx, const osb_shoot_opts *opts, const osb_profile *prof,
                       int64_t npts, const double *alpha, const double *Re, double *c,
                       int32_t *passes, int32_t *qend, int32_t *status, double *resid,
                       double *history, int32_t hist_cap) {
  if (!ctx || npts < 0 || !alpha || !Re || !c || !passes || !qend || !status)
    return fail(ctx, OSB_E_ARG, "null argument");
  if (opts && opts->flow == OSB_FLOW_ORRSPACE) return fail(ctx, OSB_E_ARG, "orrspace is spatial: use osb_shoot_spatial_lines");
  ShootArgs base;
  int analytic = 0;
  int rc = fill_args(ctx, opts, &base, &analytic);
  if (rc) return rc;
  rc = check_profile(ctx, opts, prof);
  if (rc) return rc;
  if (npts == 0) return OSB_OK;
  if (history && hist_cap < 1) return fail(ctx, OSB_E_ARG, "hist_cap < 1");
  const int ndev = (int)ctx->devices.size();
  struct Buf { double *alpha = 0, *Re = 0, *c = 0, *resid = 0, *hist = 0; int32_t *ip = 0; double2 *tab = 0; int64_t n = 0; std::vector<Seg> segs; };
  std::vector<Buf> bufs(ndev);
  {
    int64_t tile = ((npts / ((int64_t)ndev * 32) + 4095) / 4096) * 4096;
    tile = std::min<int64_t>(std::max<int64_t>(tile, 4096), (int64_t)1 << 20);
    Deal dl = deal_tiles(npts, ndev, tile);
    for (int d = 0; d < ndev; ++d) { bufs[d].segs = dl.segs[d]; bufs[d].n = dl.count[d]; }
  }
  std::vector<std::unique_ptr<Scratch>> guards(ndev);
  // phase 1: per device H2D + launch (kernels of different devices overlap)
  for (int d = 0; d < ndev; ++d) {
    Buf &b = bufs[d];
    const int64_t n = b.n;
    if (n <= 0) continue;
    cudaStream_t st = ctx->streams[d];
    OSB_CUDA(ctx, cudaSetDevice(ctx->devices[d]));
    guards[d].reset(new Scratch(ctx->devices[d], st, ctx->pools[d]));
    Scratch &scr = *guards[d];
    OSB_CUDA(ctx, scr.alloc(&b.alpha, n * 2 * sizeof(double)));
    OSB_CUDA(ctx, scr.alloc(&b.Re, n * sizeof(double)));
    OSB_CUDA(ctx, scr.alloc(&b.c, n * 2 * sizeof(double)));
    OSB_CUDA(ctx, scr.alloc(&b.ip, n * 3 * sizeof(int32_t)));
    if (resid) OSB_CUDA(ctx, scr.alloc(&b.resid, n * 2 * sizeof(double)));
    if (history) OSB_CUDA(ctx, scr.alloc(&b.hist, n * hist_cap * 4 * sizeof(double)));
    for (const Seg &g : b.segs) {
      OSB_CUDA(ctx, cudaMemcpyAsync(b.alpha + 2 * g.dev_off, alpha + 2 * g.host_lo, g.n * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
      OSB_CUDA(ctx, cudaMemcpyAsync(b.Re + g.dev_off, Re + g.host_lo, g.n * sizeof(double), cudaMemcpyHostToDevice, st));
      OSB_CUDA(ctx, cudaMemcpyAsync(b.c + 2 * g.dev_off, c + 2 * g.host_lo, g.n * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    if (history) OSB_CUDA(ctx, cudaMemsetAsync(b.hist, 0, n * hist_cap * 4 * sizeof(double), st));
    rc = make_stage_table(ctx, d, scr, opts, prof, &b.tab);
    if (rc) return rc;
    ShootArgs a = base;
    a.tab = b.tab;
    a.nlines = n;
    a.alpha = (const double2 *)b.alpha; a.Re = b.Re; a.ev = (double2 *)b.c;
    a.passes = b.ip; a.qend = b.ip + n; a.status = b.ip + 2 * n;
    a.resid = (double2 *)b.resid;
    a.history = (double4 *)b.hist;
    a.hist_cap = history ? hist_cap : 0;
    int nl = 0;
    unsigned long long *qc = nullptr;
    if (!getenv("OSB_NO_QUEUE")) {
      OSB_CUDA(ctx, scr.alloc(&qc, sizeof(unsigned long long)));
      OSB_CUDA(ctx, cudaMemsetAsync(qc, 0, sizeof(unsigned long long), st));
    }
    OSB_CUDA(ctx, launch_shoot(a, opts->integrator, analytic, st, &nl, qc));
    ctx->launches += nl;
  }
  // phase 2: D2H
  for (int d = 0; d < ndev; ++d) {
    Buf &b = bufs[d];
    const int64_t n = b.n;
    if (n <= 0) continue;
    cudaStream_t st = ctx->streams[d];
    OSB_CUDA(ctx, cudaSetDevice(ctx->devices[d]));
    for (const Seg &g : b.segs) {
      OSB_CUDA(ctx, cudaMemcpyAsync(c + 2 * g.host_lo, b.c + 2 * g.dev_off, g.n * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
      OSB_CUDA(ctx, cudaMemcpyAsync(passes + g.host_lo, b.ip + g.dev_off, g.n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      OSB_CUDA(ctx, cudaMemcpyAsync(qend + g.host_lo, b.ip + n + g.dev_off, g.n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      OSB_CUDA(ctx, cudaMemcpyAsync(status + g.host_lo, b.ip + 2 * n + g.dev_off, g.n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      if (resid) OSB_CUDA(ctx, cudaMemcpyAsync(resid + 2 * g.host_lo, b.resid + 2 * g.dev_off, g.n * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (history)
        OSB_CUDA(ctx, cudaMemcpyAsync(history + g.host_lo * hist_cap * 4, b.hist + g.dev_off * hist_cap * 4,
                                      g.n * hist_cap * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
  }
  for (int d = 0; d < ndev; ++d) {
    OSB_CUDA(ctx, cudaSetDevice(ctx->devices[d]));
    OSB_CUDA(ctx, cudaStreamSynchronize(ctx->streams[d]));
  }
  return OSB_OK;
}

int osb_shoot_spatial_lines(osb_ctx *ctx, const osb_shoot_opts *opts, const osb_profile *prof,
                            int64_t nlines, int32_t niter, const double *Re, const double *omega0,
                            const double *domega, const double *alpha0, double *alpha_out,
                            int32_t *passes, int32_t *qend, int32_t *status, double *history,
                            int32_t hist_cap) {
  if (!ctx || nlines < 0 || niter < 1 || !Re || !omega0 || !domega || !alpha0 || !alpha_out ||
      !passes || !qend || !status)
    return fail(ctx, OSB_E_ARG, "null argument");
  if (!opts || opts->flow != OSB_FLOW_ORRSPACE) return fail(ctx, OSB_E_ARG, "spatial lines need the ORRSPACE flow");
  ShootArgs base;
  int analytic = 0;
  int rc = fill_args(ctx, opts, &base, &analytic);
  if (rc) return rc;
  rc = check_profile(ctx, opts, prof);
  if (rc) return rc;
  if (nlines == 0) return OSB_OK;
  const int ndev = (int)ctx->devices.size();
  if (history && hist_cap < 1) return fail(ctx, OSB_E_ARG, "hist_cap < 1");
  // continuation lines are the indivisible unit; tiles of lines are dealt cyclically to the devices
  struct Buf { double *in = 0, *out = 0, *hist = 0; int32_t *ip = 0; double2 *tab = 0; int64_t n = 0; std::vector<Seg> segs; };
  std::vector<Buf> bufs(ndev);
  {
    const int64_t tile = std::max<int64_t>(32, (nlines + (int64_t)ndev * 8 - 1) / ((int64_t)ndev * 8));
    Deal dl = deal_tiles(nlines, ndev, tile);
    for (int d = 0; d < ndev; ++d) { bufs[d].segs = dl.segs[d]; bufs[d].n = dl.count[d]; }
  }
  std::vector<std::unique_ptr<Scratch>> guards(ndev);
  for (int d = 0; d < ndev; ++d) {
    Buf &b = bufs[d];
    const int64_t n = b.n;
    if (n <= 0) continue;
    cudaStream_t st = ctx->streams[d];
    OSB_CUDA(ctx, cudaSetDevice(ctx->devices[d]));
    guards[d].reset(new Scratch(ctx->devices[d], st, ctx->pools[d]));
    Scratch &scr = *guards[d];
    // in: Re[n], domega[n], omega0[2n], alpha0[2n]
    OSB_CUDA(ctx, scr.alloc(&b.in, n * 6 * sizeof(double)));
    OSB_CUDA(ctx, scr.alloc(&b.out, n * niter * 2 * sizeof(double)));
    OSB_CUDA(ctx, scr.alloc(&b.ip, n * niter * 3 * sizeof(int32_t)));
    if (history) {
      OSB_CUDA(ctx, scr.alloc(&b.hist, (size_t)n * niter * hist_cap * 4 * sizeof(double)));
      OSB_CUDA(ctx, cudaMemsetAsync(b.hist, 0, (size_t)n * niter * hist_cap * 4 * sizeof(double), st));
    }
    for (const Seg &g : b.segs) {
      OSB_CUDA(ctx, cudaMemcpyAsync(b.in + g.dev_off, Re + g.host_lo, g.n * sizeof(double), cudaMemcpyHostToDevice, st));
      OSB_CUDA(ctx, cudaMemcpyAsync(b.in + n + g.dev_off, domega + g.host_lo, g.n * sizeof(double), cudaMemcpyHostToDevice, st));
      OSB_CUDA(ctx, cudaMemcpyAsync(b.in + 2 * n + 2 * g.dev_off, omega0 + 2 * g.host_lo, g.n * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
      OSB_CUDA(ctx, cudaMemcpyAsync(b.in + 4 * n + 2 * g.dev_off, alpha0 + 2 * g.host_lo, g.n * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    rc = make_stage_table(ctx, d, scr, opts, prof, &b.tab);
    if (rc) return rc;
    ShootArgs a = base;
    a.tab = b.tab;
    a.nlines = n;
    a.niter = niter;
    a.Re = b.in; a.domega = b.in + n;
    a.omega0 = (const double2 *)(b.in + 2 * n);
    a.alpha = (const double2 *)(b.in + 4 * n);
    a.ev = (double2 *)b.out;
    a.passes = b.ip; a.qend = b.ip + n * niter; a.status = b.ip + 2 * n * niter;
    a.history = (double4 *)b.hist;
    a.hist_cap = history ? hist_cap : 0;
    int nl = 0;
    OSB_CUDA(ctx, launch_shoot(a, opts->integrator, analytic, st, &nl));
    ctx->launches += nl;
  }
  for (int d = 0; d < ndev; ++d) {
    Buf &b = bufs[d];
    const int64_t n = b.n;
    if (n <= 0) continue;
    cudaStream_t st = ctx->streams[d];
    const int64_t m = n * niter;
    OSB_CUDA(ctx, cudaSetDevice(ctx->devices[d]));
    for (const Seg &g : b.segs) {
      const int64_t ho = g.host_lo * niter, dof = g.dev_off * niter, cnt = g.n * niter;
      OSB_CUDA(ctx, cudaMemcpyAsync(alpha_out + 2 * ho, b.out + 2 * dof, cnt * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
      OSB_CUDA(ctx, cudaMemcpyAsync(passes + ho, b.ip + dof, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      OSB_CUDA(ctx, cudaMemcpyAsync(qend + ho, b.ip + m + dof, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      OSB_CUDA(ctx, cudaMemcpyAsync(status + ho, b.ip + 2 * m + dof, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      if (history)
        OSB_CUDA(ctx, cudaMemcpyAsync(history + (size_t)ho * hist_cap * 4, b.hist + (size_t)dof * hist_cap * 4,
                                      (size_t)cnt * hist_cap * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
  }
  for (int d = 0; d < ndev; ++d) {
    OSB_CUDA(ctx, cudaSetDevice(ctx->devices[d]));
    OSB_CUDA(ctx, cudaStreamSynchronize(ctx->streams[d]));
  }
  return OSB_OK;
}

int osb_shoot_neutral(osb_ctx *ctx, const osb_shoot_opts *opts, const osb_profile *prof,
                      int64_t npts, const double *Re, const double *alpha0, const double *c0,
                      double dalpha_rel, double tol, int32_t maxit,
                      double *alpha_out, double *c_out, int32_t *steps, int32_t *status) {
  if (!ctx || npts < 0 || !Re || !alpha0 || !c0 || !alpha_out || !c_out || !steps || !status)
    return fail(ctx, OSB_E_ARG, "null argument");
  if (!opts || opts->flow == OSB_FLOW_ORRSPACE) return fail(ctx, OSB_E_ARG, "neutral search needs a temporal flow");
  if (maxit < 1 || !(tol > 0.0) || dalpha_rel == 0.0) return fail(ctx, OSB_E_ARG, "maxit < 1, tol <= 0 or dalpha_rel == 0");
  if (npts == 0) return OSB_OK;
  // secant state per point: (a0, c at a0) and (a1, c at a1)
  std::vector<double> a_prev(alpha0, alpha0 + npts), a_cur(npts), c_prev(2 * npts), c_cur(2 * npts);
  std::vector<int64_t> active(npts);
  for (int64_t i = 0; i < npts; ++i) {
    active[i] = i;
    steps[i] = 0;
    status[i] = OSB_ST_MAXCOUNT;
    alpha_out[i] = alpha0[i];
    c_out[2 * i] = c0[2 * i]; c_out[2 * i + 1] = c0[2 * i + 1];
  }
  std::vector<double> al, rr, cc, res;
  std::vector<int32_t> ps, qe, st;
  auto solve = [&](const std::vector<int64_t> &idx, const std::vector<double> &alpha, std::vector<double> &cio) -> int {
    const int64_t m = (int64_t)idx.size();
    al.resize(2 * m); rr.resize(m); cc.resize(2 * m); ps.resize(m); qe.resize(m); st.resize(m);
    for (int64_t k = 0; k < m; ++k) {
      const int64_t i = idx[k];
      al[2 * k] = alpha[i]; al[2 * k + 1] = 0.0;
      rr[k] = Re[i];
      cc[2 * k] = cio[2 * i]; cc[2 * k + 1] = cio[2 * i + 1];
    }
    int rc = osb_shoot_temporal(ctx, opts, prof, m, al.data(), rr.data(), cc.data(), ps.data(), qe.data(), st.data(),
                                nullptr, nullptr, 0);
    if (rc) return rc;
    for (int64_t k = 0; k < m; ++k) {
      const int64_t i = idx[k];
      cio[2 * i] = cc[2 * k]; cio[2 * i + 1] = cc[2 * k + 1];
    }
    return OSB_OK;
  };
  // first point of the secant: c at alpha0
  for (int64_t i = 0; i < npts; ++i) { c_prev[2 * i] = c0[2 * i]; c_prev[2 * i + 1] = c0[2 * i + 1]; }
  int rc = solve(active, a_prev, c_prev);
  if (rc) return rc;
  {
    std::vector<int64_t> keep;
    for (int64_t k = 0; k < (int64_t)active.size(); ++k) {
      const int64_t i = active[k];
      const bool fin = std::isfinite(c_prev[2 * i]) && std::isfinite(c_prev[2 * i + 1]);
      alpha_out[i] = a_prev[i]; c_out[2 * i] = c_prev[2 * i]; c_out[2 * i + 1] = c_prev[2 * i + 1];
      if (!fin) { status[i] = OSB_ST_NONFINITE; continue; }
      if (st[k] != OSB_ST_CONVERGED) continue;                      // inner solve failed: stays MAXCOUNT
      if (fabs(c_prev[2 * i + 1]) <= tol) { status[i] = OSB_ST_CONVERGED; continue; }
      a_cur[i] = a_prev[i] * (1.0 + dalpha_rel);
      c_cur[2 * i] = c_prev[2 * i]; c_cur[2 * i + 1] = c_prev[2 * i + 1];
      keep.push_back(i);
    }
    active.swap(keep);
  }
  for (int it = 1; it <= maxit && !active.empty(); ++it) {
    rc = solve(active, a_cur, c_cur);
    if (rc) return rc;
    std::vector<int64_t> keep;
    for (int64_t k = 0; k < (int64_t)active.size(); ++k) {
      const int64_t i = active[k];
      steps[i] = it;
      const double cr = c_cur[2 * i], ci = c_cur[2 * i + 1];
      if (!(std::isfinite(cr) && std::isfinite(ci))) { status[i] = OSB_ST_NONFINITE; continue; }
      if (st[k] != OSB_ST_CONVERGED) continue;
      alpha_out[i] = a_cur[i]; c_out[2 * i] = cr; c_out[2 * i + 1] = ci;
      if (fabs(ci) <= tol) { status[i] = OSB_ST_CONVERGED; continue; }
      const double da = a_cur[i] - a_prev[i], dci = ci - c_prev[2 * i + 1];
      if (da == 0.0 || dci == 0.0) continue;                        // stalled
      const double step = -ci * da / dci;                            // secant on g(alpha) = Im c
      const double a_new = a_cur[i] + step;
      if (!(a_new > 0.0) || !std::isfinite(a_new)) continue;
      // continue the eigenvalue guess along the secant
      const double dcr = (cr - c_prev[2 * i]) / da, dcim = dci / da;
      a_prev[i] = a_cur[i]; c_prev[2 * i] = cr; c_prev[2 * i + 1] = ci;
      a_cur[i] = a_new;
      c_cur[2 * i] = cr + dcr * step; c_cur[2 * i + 1] = ci + dcim * step;
      keep.push_back(i);
    }
    active.swap(keep);
  }
  return OSB_OK;
}

// one chunk of osb_eigfun: all scratch of the chunk lives in `scr`
static int eigfun_chunk(osb_ctx *ctx, int slot, const osb_shoot_opts *opts, const osb_profile *prof, const ShootArgs &base,
                        int analytic, int64_t npts, const double *alpha_or_omega, const double *Re,
                        const double *eig, double *y, double *vmax, int32_t *qend) {
  cudaStream_t st = ctx->streams[slot];
  OSB_CUDA(ctx, cudaSetDevice(ctx->devices[slot]));
  const size_t n1 = (size_t)opts->nstep + 1;
  Scratch scr(ctx->devices[slot], st, ctx->pools[slot]);
  double *d_in = nullptr, *d_tq = nullptr;
  double2 *d_U = nullptr, *d_P = nullptr, *d_y = nullptr, *d_vmax = nullptr, *tab = nullptr;
  int32_t *d_q = nullptr;
  OSB_CUDA(ctx, scr.alloc(&d_in, npts * 5 * sizeof(double)));  // eig(2), alpha|omega(2), Re
  OSB_CUDA(ctx, scr.alloc(&d_U, npts * n1 * 8 * sizeof(double2)));
  OSB_CUDA(ctx, scr.alloc(&d_P, npts * n1 * 4 * sizeof(double2)));
  OSB_CUDA(ctx, scr.alloc(&d_tq, npts * n1 * sizeof(double)));
  OSB_CUDA(ctx, scr.alloc(&d_y, npts * n1 * 4 * sizeof(double2)));
  OSB_CUDA(ctx, scr.alloc(&d_vmax, npts * sizeof(double2)));
  OSB_CUDA(ctx, scr.alloc(&d_q, npts * sizeof(int32_t)));
  // layout: eig(2n), alpha|omega(2n), Re(n)  -- the double2 arrays first (16-byte alignment)
  OSB_CUDA(ctx, cudaMemcpyAsync(d_in, eig, npts * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  OSB_CUDA(ctx, cudaMemcpyAsync(d_in + 2 * npts, alpha_or_omega, npts * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  OSB_CUDA(ctx, cudaMemcpyAsync(d_in + 4 * npts, Re, npts * sizeof(double), cudaMemcpyHostToDevice, st));
  int rc = make_stage_table(ctx, slot, scr, opts, prof, &tab);
  if (rc) return rc;
  ShootArgs a = base;
  EigfunArgs e;
  memset(&e, 0, sizeof(e));
  a.tab = tab;
  a.Re = d_in + 4 * npts;
  a.alpha = (const double2 *)(d_in + 2 * npts);
  e.s = a;
  e.eig = (const double2 *)d_in;
  e.omega = (const double2 *)(d_in + 2 * npts);
  e.U = d_U; e.P = d_P; e.tq = d_tq; e.y = d_y; e.vmax = d_vmax; e.qend = d_q;
  e.t0 = opts->ymin;
  e.habs = fabs(a.hs);
  OSB_CUDA(ctx, launch_eigfun(e, opts->integrator, analytic, npts, st));
  ctx->launches++;
  OSB_CUDA(ctx, cudaMemcpyAsync(y, d_y, npts * n1 * 4 * sizeof(double2), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(ctx, cudaMemcpyAsync(vmax, d_vmax, npts * sizeof(double2), cudaMemcpyDeviceToHost, st));
  if (qend) OSB_CUDA(ctx, cudaMemcpyAsync(qend, d_q, npts * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(ctx, cudaStreamSynchronize(st));
  return OSB_OK;
}

int osb_eigfun(osb_ctx *ctx, const osb_shoot_opts *opts, const osb_profile *prof, int64_t npts,
               const double *alpha, const double *Re, const double *eig, const double *omega,
               double *y, double *vmax, int32_t *qend) {
  if (!ctx || npts < 0 || !Re || !eig || !y || !vmax) return fail(ctx, OSB_E_ARG, "null argument");
  ShootArgs a;
  int analytic = 0;
  int rc = fill_args(ctx, opts, &a, &analytic);
  if (rc) return rc;
  rc = check_profile(ctx, opts, prof);
  if (rc) return rc;
  if (a.spatial ? !omega : !alpha) return fail(ctx, OSB_E_ARG, "alpha (temporal) / omega (spatial) required");
  if (npts == 0) return OSB_OK;
  // the second pass keeps the whole orthonormalised trajectory (shoot.f:762-810): 272 bytes per step per
  // point of device scratch.  Bound it to ~8 GB per chunk so that any npts fits the GPU.
  const size_t n1 = (size_t)opts->nstep + 1;
  const size_t per_point = n1 * (8 + 4 + 4) * sizeof(double2) + n1 * sizeof(double) + 64;
  int64_t chunk = std::max<int64_t>(1, (int64_t)((8ull << 30) / per_point));
  if (const char *e = getenv("OSB_EIGFUN_CHUNK")) chunk = std::max<int64_t>(1, atoll(e));  // test hook
  const double *ao = a.spatial ? omega : alpha;
  const int ndev = (int)ctx->devices.size();
  if (ndev > 1) chunk = std::min<int64_t>(chunk, std::max<int64_t>(1, (npts + ndev - 1) / ndev));
  const int64_t nchunk = (npts + chunk - 1) / chunk;
  // chunk k runs on device k mod ndev; one worker thread per device walks its chunks (points are independent)
  auto worker = [&](int slot) -> int {
    for (int64_t k = slot; k < nchunk; k += ndev) {
      const int64_t lo = k * chunk, m = std::min(chunk, npts - lo);
      int rcw = eigfun_chunk(ctx, slot, opts, prof, a, analytic, m, ao + 2 * lo, Re + lo, eig + 2 * lo, y + (size_t)lo * n1 * 8,
                             vmax + 2 * lo, qend ? qend + lo : nullptr);
      if (rcw) return rcw;
    }
    return OSB_OK;
  };
  if (ndev == 1 || nchunk == 1) return worker(0);
  std::vector<int> rcs(ndev, OSB_OK);
  std::vector<std::thread> th;
  for (int d = 0; d < ndev; ++d) th.emplace_back([&, d] { rcs[d] = worker(d); });
  for (auto &t : th) t.join();
  for (int d = 0; d < ndev; ++d) if (rcs[d]) return rcs[d];
  return OSB_OK;
}

// ---------------------------------------------------------------- measurement
// shared body of the two peak micro-benchmarks: `flops_per_thread_iter` = flops one thread performs per loop trip
static int measure_peak(osb_ctx *ctx, double ms_target, bool dmma, double *tflops) {
  OSB_CUDA(ctx, cudaSetDevice(ctx->devices[0]));
  cudaStream_t st = ctx->streams[0];
  cudaDeviceProp prop;
  OSB_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->devices[0]));
  const int threads = 256, blocks = prop.multiProcessorCount * 8;
  Scratch scr(ctx->devices[0], st, ctx->pools[0]);
  double *d_out = nullptr;
  OSB_CUDA(ctx, scr.alloc(&d_out, sizeof(double)));
  struct Events {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Events() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
  } ev;
  OSB_CUDA(ctx, cudaEventCreate(&ev.e0));
  OSB_CUDA(ctx, cudaEventCreate(&ev.e1));
  cudaEvent_t e0 = ev.e0, e1 = ev.e1;
  // DFMA kernel: 8 chains x 16 unrolled FMAs x 2 flops per thread and trip;
  // DMMA kernel: 4 x 8 instructions of 512 flops per WARP and trip = 512 flops per thread and trip
  const double per_thread_trip = dmma ? 4.0 * 8.0 * 512.0 / 32.0 : 2.0 * 8.0 * 16.0;
  int iters = 2000;
  double best = 0.0;
  float ms = 0.f;
  for (int rep = 0; rep < 6; ++rep) {
    OSB_CUDA(ctx, cudaEventRecord(e0, st));
    OSB_CUDA(ctx, dmma ? launch_dmma_peak(d_out, iters, blocks, threads, st) : launch_fp64_peak(d_out, iters, blocks, threads, st));
    ctx->launches++;
    OSB_CUDA(ctx, cudaEventRecord(e1, st));
    OSB_CUDA(ctx, cudaEventSynchronize(e1));
    OSB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    const double flops = per_thread_trip * (double)iters * threads * (double)blocks;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0) best = std::max(best, tf);
    if (ms < ms_target && iters < (1 << 24)) iters = (int)std::min<double>(1 << 24, iters * std::max(1.5, ms_target / std::max(ms, 1e-3f)));
  }
  *tflops = best;
  OSB_CUDA(ctx, cudaStreamSynchronize(st));
  return OSB_OK;
}

int osb_measure_fp64_peak(osb_ctx *ctx, double ms_target, double *tflops, double *sm_clock_mhz_est) {
  if (!ctx || !tflops) return OSB_E_ARG;
  int rc = measure_peak(ctx, ms_target, false, tflops);
  if (rc) return rc;
  if (sm_clock_mhz_est) {  // 64 DFMA / clk / SM on B200
    cudaDeviceProp prop;
    OSB_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->devices[0]));
    *sm_clock_mhz_est = *tflops * 1e12 / (2.0 * 64.0 * prop.multiProcessorCount) / 1e6;
  }
  return OSB_OK;
}

int osb_measure_dmma_peak(osb_ctx *ctx, double ms_target, double *tflops) {
  if (!ctx || !tflops) return OSB_E_ARG;
  return measure_peak(ctx, ms_target, true, tflops);
}

}  // extern "C"
