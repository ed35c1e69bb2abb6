// osb_sweep_api.cu -- self-seeding temporal sweep over a tensor grid (include/osb.h, osb_shoot_grid_temporal).
//
// The reference's only seeding scheme is continuation: every program starts from a guess typed into its deck and the
// spatial sweep hands each converged alpha to the next omega (orrspace.f:186-195, 494).  CONTE_IC converges from a
// guess within ~3e-3 of the root (1e-5 / 1e-4 / 1e-3 / 3e-3 -> 5 / 6 / 7-8 / 8-9 passes, 2e-2
// fails), so a sweep over an (alpha, Re) grid needs a seed per point.  This driver builds them on the device from ONE
// anchor guess (SURVEY 7.3 P10 / 8d "multilevel"):
//   1. the anchor is solved;
//   2. a coarse grid (<= 65 x 65 nodes of the user's grid) is solved by batched continuation: one walker runs from
//      the anchor along the coarse row through it, then two walkers per coarse column run up and down -- all walkers
//      of a phase advance together, one batched osb_shoot_temporal call per round.  A walker's guess is the linear
//      extrapolation of its last two solutions; a step that fails (not converged, > 12 passes, or landed > 1e-2
//      from its guess: another branch) is bisected in (alpha, log Re), up to 7 levels;
//   3. the coarse solutions are interpolated bilinearly in (alpha, log Re) to every point of the grid and the whole
//      grid is solved in one batched call.
// Host logic only: every eigenvalue solve goes through osb_shoot_temporal (the CUDA kernels).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <complex>
#include <string>
#include <vector>

#include "osb_internal.h"

namespace osb {
int ctx_fail(osb_ctx *ctx, int code, const std::string &msg);
}
using namespace osb;

namespace {

typedef std::complex<double> cd;

struct Pt { double a, r; };

struct Walker {
  std::vector<Pt> targets;      // coarse nodes to visit, in order
  std::vector<int64_t> slots;   // where each target's solution goes in the coarse table
  size_t next = 0;
  Pt cur{0, 0}, prev{0, 0};
  cd ccur{0, 0}, cprev{0, 0};
  bool have_prev = false, dead = false;
  std::vector<Pt> stack;        // pending sub-targets; back() is the one being tried
  cd guess{0, 0};
};

struct Metric {                 // distances in the normalised (alpha, log Re) plane
  double a_span, lr_span;
  double dist(const Pt &p, const Pt &q) const {
    return hypot((p.a - q.a) / a_span, log(p.r / q.r) / lr_span);
  }
};

// advance every live walker until all targets are visited or given up; one batched solve per round
int run_walkers(osb_ctx *ctx, const osb_shoot_opts *opts, const osb_profile *prof, const Metric &mt,
                std::vector<Walker> &ws, std::vector<cd> &table, std::vector<char> &table_ok, int32_t *info) {
  std::vector<double> al, re, cc;
  std::vector<int32_t> ps, qe, st;
  std::vector<size_t> who;
  for (;;) {
    al.clear(); re.clear(); cc.clear(); who.clear();
    for (size_t w = 0; w < ws.size(); ++w) {
      Walker &k = ws[w];
      if (k.dead || k.next >= k.targets.size()) continue;
      if (k.stack.empty()) k.stack.push_back(k.targets[k.next]);
      const Pt t = k.stack.back();
      cd g = k.ccur;
      if (k.have_prev) {
        const double d0 = mt.dist(k.cur, k.prev), d1 = mt.dist(t, k.cur);
        if (d0 > 0.0) g = k.ccur + (k.ccur - k.cprev) * (d1 / d0);
      }
      k.guess = g;
      al.push_back(t.a); al.push_back(0.0);
      re.push_back(t.r);
      cc.push_back(g.real()); cc.push_back(g.imag());
      who.push_back(w);
    }
    if (who.empty()) return OSB_OK;
    const int64_t m = (int64_t)who.size();
    ps.resize(m); qe.resize(m); st.resize(m);
    int rc = osb_shoot_temporal(ctx, opts, prof, m, al.data(), re.data(), cc.data(), ps.data(), qe.data(), st.data(),
                                nullptr, nullptr, 0);
    if (rc) return rc;
    if (info) { info[0] += (int32_t)m; info[1] += 1; }
    for (int64_t e = 0; e < m; ++e) {
      Walker &k = ws[who[e]];
      const cd c(cc[2 * e], cc[2 * e + 1]);
      const bool ok = st[e] == OSB_ST_CONVERGED && ps[e] <= 12 && std::abs(c - k.guess) < 1.0e-2;
      if (ok) {
        k.prev = k.cur; k.cprev = k.ccur; k.have_prev = true;
        k.cur = k.stack.back(); k.ccur = c;
        k.stack.pop_back();
        if (k.stack.empty()) {  // that was the coarse node itself
          table[k.slots[k.next]] = c;
          table_ok[k.slots[k.next]] = 1;
          k.next += 1;
        }
      } else if (k.stack.size() > 7) {
        k.dead = true;  // the remaining nodes of this walker keep the value of the last solved one
        if (info) info[3] += (int32_t)(k.targets.size() - k.next);
        for (size_t q = k.next; q < k.targets.size(); ++q) table[k.slots[q]] = k.ccur;
      } else {
        const Pt t = k.stack.back();
        k.stack.push_back(Pt{0.5 * (k.cur.a + t.a), sqrt(k.cur.r * t.r)});
        if (info) info[2] += 1;
      }
    }
  }
}

// coarse node indices along an axis of n points: first, last and evenly spaced between (<= maxn nodes)
std::vector<int> coarse_axis(int n, int maxn) {
  std::vector<int> idx;
  const int m = std::max(2, std::min(n, maxn));
  for (int k = 0; k < m; ++k) {
    const int i = (int)llround((double)k * (n - 1) / (double)(m - 1));
    if (idx.empty() || i != idx.back()) idx.push_back(i);
  }
  return idx;
}

}  // namespace

extern "C" int osb_shoot_grid_temporal(osb_ctx *ctx, const osb_shoot_opts *opts, const osb_profile *prof, int32_t na,
                                       const double *alpha, int32_t nre, const double *Re, double anchor_alpha,
                                       double anchor_Re, const double *anchor_c, int32_t coarse, double *c,
                                       int32_t *passes, int32_t *qend, int32_t *status, double *seed, int32_t *info) {
  if (!ctx || !opts || !prof || na < 1 || nre < 1 || !alpha || !Re || !anchor_c || !c || !passes || !qend || !status)
    return ctx_fail(ctx, OSB_E_ARG, "null argument");
  if (opts->flow == OSB_FLOW_ORRSPACE) return ctx_fail(ctx, OSB_E_ARG, "the grid driver is for the temporal flows");
  for (int j = 1; j < na; ++j) if (!(alpha[j] > alpha[j - 1])) return ctx_fail(ctx, OSB_E_ARG, "alpha must be ascending");
  for (int i = 1; i < nre; ++i) if (!(Re[i] > Re[i - 1])) return ctx_fail(ctx, OSB_E_ARG, "Re must be ascending");
  if (!(Re[0] > 0.0) || !(anchor_Re > 0.0)) return ctx_fail(ctx, OSB_E_ARG, "Re must be positive");
  if (info) memset(info, 0, 4 * sizeof(int32_t));
  const int maxn = coarse > 1 ? coarse : 65;
  const std::vector<int> ja = coarse_axis(na, maxn), ia = coarse_axis(nre, maxn);
  const int nca = (int)ja.size(), ncr = (int)ia.size();
  Metric mt;
  mt.a_span = std::max(alpha[na - 1] - alpha[0], 1.0e-300);
  mt.lr_span = std::max(log(Re[nre - 1] / Re[0]), 1.0e-300);
  if (na == 1) mt.a_span = std::max(fabs(alpha[0]), 1.0);
  if (nre == 1) mt.lr_span = 1.0;
  std::vector<cd> table((size_t)nca * ncr, cd(anchor_c[0], anchor_c[1]));  // [ir][ja]
  std::vector<char> table_ok(table.size(), 0);
  // ---- phase A: anchor -> nearest coarse node -> along its coarse row (both directions)
  int j0 = 0, i0 = 0;
  {
    double best = 1.0e300;
    for (int k = 0; k < nca; ++k) { const double d = fabs(alpha[ja[k]] - anchor_alpha); if (d < best) { best = d; j0 = k; } }
    best = 1.0e300;
    for (int k = 0; k < ncr; ++k) { const double d = fabs(log(Re[ia[k]] / anchor_Re)); if (d < best) { best = d; i0 = k; } }
  }
  {
    // the anchor itself, from the caller's guess
    double al[2] = {anchor_alpha, 0.0}, cc[2] = {anchor_c[0], anchor_c[1]};
    int32_t ps, qe, st;
    int rc = osb_shoot_temporal(ctx, opts, prof, 1, al, &anchor_Re, cc, &ps, &qe, &st, nullptr, nullptr, 0);
    if (rc) return rc;
    if (info) { info[0] += 1; info[1] += 1; }
    if (st != OSB_ST_CONVERGED) return ctx_fail(ctx, OSB_E_STATE, "the anchor guess did not converge");
    std::vector<Walker> ws(1);
    ws[0].cur = Pt{anchor_alpha, anchor_Re};
    ws[0].ccur = cd(cc[0], cc[1]);
    ws[0].targets.push_back(Pt{alpha[ja[j0]], Re[ia[i0]]});
    ws[0].slots.push_back((int64_t)i0 * nca + j0);
    rc = run_walkers(ctx, opts, prof, mt, ws, table, table_ok, info);
    if (rc) return rc;
    if (!table_ok[(size_t)i0 * nca + j0]) return ctx_fail(ctx, OSB_E_STATE, "continuation from the anchor to the grid failed");
    std::vector<Walker> row(2);
    for (int dir = 0; dir < 2; ++dir) {
      Walker &k = row[dir];
      k.cur = Pt{alpha[ja[j0]], Re[ia[i0]]};
      k.ccur = table[(size_t)i0 * nca + j0];
      for (int j = dir ? j0 - 1 : j0 + 1; dir ? j >= 0 : j < nca; j += dir ? -1 : 1) {
        k.targets.push_back(Pt{alpha[ja[j]], Re[ia[i0]]});
        k.slots.push_back((int64_t)i0 * nca + j);
      }
    }
    rc = run_walkers(ctx, opts, prof, mt, row, table, table_ok, info);
    if (rc) return rc;
  }
  // ---- phase B: two walkers per coarse column, all advancing together
  {
    std::vector<Walker> cols;
    for (int j = 0; j < nca; ++j)
      for (int dir = 0; dir < 2; ++dir) {
        Walker k;
        k.cur = Pt{alpha[ja[j]], Re[ia[i0]]};
        k.ccur = table[(size_t)i0 * nca + j];
        for (int i = dir ? i0 - 1 : i0 + 1; dir ? i >= 0 : i < ncr; i += dir ? -1 : 1) {
          k.targets.push_back(Pt{alpha[ja[j]], Re[ia[i]]});
          k.slots.push_back((int64_t)i * nca + j);
        }
        if (!k.targets.empty()) cols.push_back(std::move(k));
      }
    int rc = run_walkers(ctx, opts, prof, mt, cols, table, table_ok, info);
    if (rc) return rc;
  }
  // ---- bilinear interpolation in (alpha, log Re) between coarse nodes -> the guess of every grid point
  const int64_t npts = (int64_t)na * nre;
  std::vector<double> al(2 * npts), rr(npts);
  {
    std::vector<int> cja(na), cia(nre);  // coarse cell of every fine index
    std::vector<double> ta(na), tr(nre);
    for (int j = 0, k = 0; j < na; ++j) {
      while (k + 2 < nca && j >= ja[k + 1]) ++k;
      cja[j] = k;
      ta[j] = nca > 1 ? (alpha[j] - alpha[ja[k]]) / (alpha[ja[std::min(k + 1, nca - 1)]] - alpha[ja[k]]) : 0.0;
    }
    for (int i = 0, k = 0; i < nre; ++i) {
      while (k + 2 < ncr && i >= ia[k + 1]) ++k;
      cia[i] = k;
      tr[i] = ncr > 1 ? log(Re[i] / Re[ia[k]]) / log(Re[ia[std::min(k + 1, ncr - 1)]] / Re[ia[k]]) : 0.0;
    }
    for (int i = 0; i < nre; ++i) {
      const int ki = cia[i], ki1 = std::min(ki + 1, ncr - 1);
      for (int j = 0; j < na; ++j) {
        const int kj = cja[j], kj1 = std::min(kj + 1, nca - 1);
        const double fx = ta[j], fy = tr[i];
        const cd g = (1 - fy) * ((1 - fx) * table[(size_t)ki * nca + kj] + fx * table[(size_t)ki * nca + kj1]) +
                     fy * ((1 - fx) * table[(size_t)ki1 * nca + kj] + fx * table[(size_t)ki1 * nca + kj1]);
        const int64_t p = (int64_t)i * na + j;
        al[2 * p] = alpha[j]; al[2 * p + 1] = 0.0;
        rr[p] = Re[i];
        c[2 * p] = g.real(); c[2 * p + 1] = g.imag();
      }
    }
  }
  if (seed) memcpy(seed, c, 2 * npts * sizeof(double));
  // ---- the whole grid in one batched call
  return osb_shoot_temporal(ctx, opts, prof, npts, al.data(), rr.data(), c, passes, qend, status, nullptr, nullptr, 0);
}
