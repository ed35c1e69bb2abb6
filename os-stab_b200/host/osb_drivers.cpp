// osb_drivers.cpp -- C++ twin drivers of the reference's programs over the C ABI.
//
//   osb_drivers contebl | conte | orrsom | orrspace | orrfdchan | orrcolchan | orrncbl | conteucbl | orrfdbl | orrcolbl
//
// Each twin reads the same stdin (prompts, "key = value" decks, list-directed numbers),
// echoes the same text and writes the same fort.N / <root>.g/.dat/.q files as the
// Fortran program of that name; the only thing that changes is where the work is done:
// `call SOLVE_BL`, `call CONTE_IC/CONTE`, `call MAKE_MATRIX + SOLVE_ORR_SOM` become the
// libosb.so entry points that INTEGRATION.md binds for the Fortran hosts.  (No Fortran
// compiler exists in this image; these twins are how the host layer is exercised.)
//
// Known, documented deviations from the reference's text:
//   * columns 4-5 of the per-pass table (err) agree in modulus only; their phase is a
//     rounding artefact of the reference's arithmetic;
//   * orrcolchan prints the table row of the mode the GPU path isolates (the TS mode,
//     sorted index 62 of chan.inp) instead of the whole QZ spectrum;
//   * gfortran's list-directed `write(*,*)` spacing (orrncbl.f:902, orrcolchan.f:539,546)
//     is not reproduced.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/osb.h"
#include "fortfmt.hpp"

using fortfmt::E;
using fortfmt::ES;
using fortfmt::I;
using fortfmt::F;
typedef std::complex<double> cd;

static void die(osb_ctx *ctx, const char *what, int rc) {
  fprintf(stderr, "osb_drivers: %s failed (%d): %s\n", what, rc, ctx ? osb_last_error(ctx) : "");
  exit(2);
}
#define CK(ctx, call)                 \
  do {                                \
    int rc__ = (call);                \
    if (rc__) die((ctx), #call, rc__); \
  } while (0)

static std::string getline80() {
  std::string s;
  if (!std::getline(std::cin, s)) s.clear();
  return s;
}

// ---- deck (contebl.f:122-149): skip leading '#' lines, then positional "key = value" lines
struct Deck {
  std::ifstream f;
  bool first = true;
  bool open(const std::string &name) { f.open(name); return (bool)f; }
  std::string next_line() {
    std::string s;
    if (first) {
      first = false;
      do { if (!std::getline(f, s)) return ""; } while (!s.empty() && s[0] == '#');
      return s;
    }
    std::getline(f, s);
    return s;
  }
  static void input_error() { printf("\n\n ERROR in input file...\n\n\n"); exit(0); }  // contebl.f:586-588
  long long geti() {
    std::string fld;
    long long v = 0;
    if (!fortfmt::deck_field(next_line(), fld) || !fortfmt::read_I10(fld, v)) input_error();
    return v;
  }
  double getr() {
    std::string fld;
    double v = 0;
    if (!fortfmt::deck_field(next_line(), fld) || !fortfmt::read_G20_10(fld, v)) input_error();
    return v;
  }
};

// SPDER (shoot.f:1037-1101) on the host, only to print fort.10 from the tables
static int interval(const std::vector<double> &X, double xx) {
  const int N = (int)X.size();
  if (xx == X[0]) return 1;
  if (xx == X[N - 1]) return N;
  int il = 0, ir = N - 1;
  while (ir - il > 1) { int im = (ir + il) >> 1; if (X[im] > xx) ir = im; else il = im; }
  return il + 1;
}
static void spder(const std::vector<double> &X, const std::vector<double> &Y, const std::vector<double> &F, double xx,
                  double &f, double &fpp) {
  const int N = (int)X.size();
  const int Ii = interval(X, xx);
  auto x = [&](int i) { return i <= N ? X[i - 1] : 0.0; };  // one element past the table reads COMMON zeros
  auto y = [&](int i) { return i <= N ? Y[i - 1] : 0.0; };
  auto d = [&](int i) { return i <= N ? F[i - 1] : 0.0; };
  const double DXM = xx - x(Ii), DXP = x(Ii + 1) - xx, DEL = x(Ii + 1) - x(Ii);
  f = d(Ii) * DXP * (DXP * DXP / DEL - DEL) / 6.0 + d(Ii + 1) * DXM * (DXM * DXM / DEL - DEL) / 6.0 + y(Ii) * DXP / DEL +
      y(Ii + 1) * DXM / DEL;
  fpp = d(Ii) * DXP / DEL + d(Ii + 1) * DXM / DEL;
}
static double speval_linear(const std::vector<double> &X, const std::vector<double> &Y, const std::vector<double> &F, double xx) {
  const int N = (int)X.size();
  int Ii = N;
  for (int i = 1; i <= N - 1; ++i) if (xx <= X[i]) { Ii = i; break; }
  auto x = [&](int i) { return i <= N ? X[i - 1] : 0.0; };
  auto y = [&](int i) { return i <= N ? Y[i - 1] : 0.0; };
  auto d = [&](int i) { return i <= N ? F[i - 1] : 0.0; };
  const double DXM = xx - x(Ii), DXP = x(Ii + 1) - xx, DEL = x(Ii + 1) - x(Ii);
  return d(Ii) * DXP * (DXP * DXP / DEL - DEL) / 6.0 + d(Ii + 1) * DXM * (DXM * DXM / DEL - DEL) / 6.0 + y(Ii) * DXP / DEL +
         y(Ii + 1) * DXM / DEL;
}

struct Tables {
  std::vector<double> ydat, U, Uspl, d2U, d2Uspl;
};
static Tables fetch_tables(osb_ctx *ctx, osb_profile *prof, int nbl) {
  Tables t;
  t.ydat.resize(nbl + 1); t.U.resize(nbl + 1); t.Uspl.resize(nbl + 1); t.d2U.resize(nbl + 1); t.d2Uspl.resize(nbl + 1);
  int32_t np = 0;
  double f2 = 0;
  CK(ctx, osb_profile_get(ctx, prof, t.ydat.data(), t.U.data(), t.Uspl.data(), t.d2U.data(), t.d2Uspl.data(), &np, &f2));
  return t;
}

// fort.11-14 (shoot.f:802-809): t, Re(y_i/vmax), Im(y_i/vmax)
static void write_eigfun(const std::vector<double> &y, cd vmax, int nstep, double tf, double to, double h, bool bl,
                         int wES, bool lead1x, double tscale = 1.0) {
  for (int comp = 0; comp < 4; ++comp) {
    char name[32];
    snprintf(name, sizeof name, "fort.%d", 11 + comp);
    FILE *f = fopen(name, "w");
    for (int k = 0; k <= nstep; ++k) {
      const double t = (bl ? tf - h * k : to + h * k) * tscale;
      const cd v = cd(y[2 * (4 * k + comp)], y[2 * (4 * k + comp) + 1]) / vmax;
      fprintf(f, "%s%s %s %s \n", lead1x ? " " : "", ES(t, wES, 8, 3).c_str(), ES(v.real(), wES, 8, 3).c_str(),
              ES(v.imag(), wES, 8, 3).c_str());
    }
    fclose(f);
  }
}

// ---------------------------------------------------------------- contebl
static int run_contebl(osb_ctx *ctx) {
  int nbl = 1000, nstep = 1000;  // contebl.f:64-83
  double testalpha = 10.0, Re = 580., beta = 0.0, alphar = 0.179, alphai = 0.0;
  double cr = 0.36413124E+00, ci = 0.79527387E-02, ymin = 0.0, ymax = 17.0, f2p = 0.5;
  const bool eigfun = true;
  printf("\n          Solve Orr-Sommerfeld for Boundary Layer using shooting\n\n");
  printf("\n Read from keyboard, file, default (k,f,d) ==> ");
  fflush(stdout);
  const std::string input = getline80();
  const char ch = input.empty() ? ' ' : input[0];
  if (ch == 'k' || ch == 'K') {
    printf("\n Enter number of steps for BL solve ==> "); std::cin >> nbl;
    printf("\n Enter number of steps for shooting ==> "); std::cin >> nstep;
    printf("\n Enter Reynolds number ==> "); std::cin >> Re;
    printf("\n Enter beta ==> "); std::cin >> beta;
    printf("\n Enter alpha_r ==> "); std::cin >> alphar;
    printf("\n Enter alpha_i ==> "); std::cin >> alphai;
    printf("\n Enter guess for (c_r, c_i) ==> "); std::cin >> cr >> ci;
    printf("\n Enter Ymin and Ymax ==> "); std::cin >> ymin >> ymax;
  } else if (ch == 'f' || ch == 'F') {
    printf("\n Enter filename ==> ");
    fflush(stdout);
    std::string infile = getline80();
    while (!infile.empty() && infile.back() == ' ') infile.pop_back();
    Deck d;
    if (!d.open(infile)) { printf("\n\n Error in input file...\n\n\n"); return 0; }
    nbl = (int)d.geti(); nstep = (int)d.geti(); testalpha = d.getr(); Re = d.getr(); beta = d.getr(); f2p = d.getr();
    alphar = d.getr(); alphai = d.getr(); cr = d.getr(); ci = d.getr(); ymin = d.getr(); ymax = d.getr();
  }
  printf("\n nbl = %s\n", I(nbl, 5).c_str());
  printf(" nstep = %s\n", I(nstep, 5).c_str());
  printf(" Test Alpha = %s\n", E(testalpha, 20, 10).c_str());
  printf(" Re = %s\n", E(Re, 20, 10).c_str());
  printf(" Beta = %s\n", E(beta, 20, 10).c_str());
  printf(" f2p = %s\n", E(f2p, 20, 10).c_str());
  printf(" Ymin = %s  Ymax = %s\n", E(ymin, 20, 10).c_str(), E(ymax, 20, 10).c_str());
  printf(" alpha = (%s, %s)\n", E(alphar, 20, 10).c_str(), E(alphai, 20, 10).c_str());
  printf(" c = (%s, %s)\n\n", E(cr, 20, 10).c_str(), E(ci, 20, 10).c_str());
  cd alpha(alphar, alphai), c(cr, ci);
  const double hbl = (ymax - ymin) / (double)nbl;
  Re = Re * std::sqrt(2.);  // contebl.f:181-182
  alpha = alpha * std::sqrt(2.);
  if (nbl > 50000) { printf("\n\n N > Idim...Stopping\n\n\n"); return 0; }
  // call SOLVE_BL  (contebl.f:196)
  osb_profile *prof = nullptr;
  CK(ctx, osb_profile_blasius(ctx, OSB_FLOW_CONTEBL, OSB_INT_CK45, nbl, ymin, ymax, f2p, beta, &prof));
  const Tables t = fetch_tables(ctx, prof, nbl);
  {
    FILE *f = fopen("fort.10", "w");  // contebl.f:511-526
    for (int i = 0; i <= nbl; ++i) {
      const double y = ymin + i * hbl;
      double us, d2us;
      spder(t.ydat, t.U, t.Uspl, y, us, d2us);
      fprintf(f, " %s %s %s %s %s \n", ES(y, 16, 8, 3).c_str(), ES(us, 16, 8, 3).c_str(), ES(d2us, 16, 8, 3).c_str(),
              ES(t.U[i], 16, 8, 3).c_str(), ES(t.d2U[i], 16, 8, 3).c_str());
    }
    fclose(f);
  }
  printf(" Blasius velocity profile completed...\n\n");
  // call CONTE_IC(...)  (contebl.f:205)
  osb_shoot_opts o;
  CK(ctx, osb_shoot_defaults(OSB_FLOW_CONTEBL, &o));
  o.nstep = nstep; o.testalpha = testalpha; o.ymin = ymin; o.ymax = ymax;
  double al[2] = {alpha.real(), alpha.imag()}, cc[2] = {c.real(), c.imag()}, resid[2], hist[4 * 20];
  int32_t passes, qend, status;
  CK(ctx, osb_shoot_temporal(ctx, &o, prof, 1, al, &Re, cc, &passes, &qend, &status, resid, hist, 20));
  for (int i = 0; i < passes && i < 20; ++i)  // shoot.f:745-747
    printf(" %s%s%s   %s%s   \n", I(i + 1, 4).c_str(), E(hist[4 * i], 17, 8).c_str(), E(hist[4 * i + 1], 17, 8).c_str(),
           E(hist[4 * i + 2], 17, 8).c_str(), E(hist[4 * i + 3], 17, 8).c_str());
  printf("\nEigenvalue = %s %s  %s\n\n", E(cc[0], 17, 8).c_str(), E(cc[1], 17, 8).c_str(), I(qend, 5).c_str());
  printf("  Eigenvalue iterations = %s\n  |error| = %s\n  |c-cm1| = %s\n\n", I(passes, 5).c_str(),
         ES(resid[0], 17, 8).c_str(), ES(resid[1], 17, 8).c_str());
  if (eigfun) {
    std::vector<double> y((size_t)(nstep + 1) * 8);
    double vmax[2];
    int32_t q2;
    CK(ctx, osb_eigfun(ctx, &o, prof, 1, al, &Re, cc, nullptr, y.data(), vmax, &q2));
    write_eigfun(y, cd(vmax[0], vmax[1]), nstep, ymax, ymin, (ymax - ymin) / nstep, true, 16, true);
  }
  osb_profile_free(ctx, prof);
  return status;
}

// ---------------------------------------------------------------- conteucbl
// Saddle-point tracking in a frame moving with u_c (conteucbl.f:5-300).  Its private CONTE is
// CONTE_IC's iteration (tolerances 1e-14, shoot.f:461-479 rules) with CRK4 and the SPDER right-hand
// side -- i.e. flow CONTEBL with the RK4 integrator -- on the profile shifted by SHIFT_BL.
static void conteucbl_solve(osb_ctx *ctx, const osb_shoot_opts &o, osb_profile *prof, cd alpha, double Re, cd &c, int &qend) {
  double al[2] = {alpha.real(), alpha.imag()}, cc[2] = {c.real(), c.imag()}, resid[2], hist[4 * 20];
  int32_t passes, status;
  int32_t q = 0;
  CK(ctx, osb_shoot_temporal(ctx, &o, prof, 1, al, &Re, cc, &passes, &q, &status, resid, hist, 20));
  for (int i = 0; i < passes && i < 20; ++i)  // conteucbl.f:577-579
    printf(" %s%s%s   %s%s   \n", I(i + 1, 4).c_str(), E(hist[4 * i], 17, 8).c_str(), E(hist[4 * i + 1], 17, 8).c_str(),
           E(hist[4 * i + 2], 17, 8).c_str(), E(hist[4 * i + 3], 17, 8).c_str());
  // eigfun is forced .true. inside the reference's CONTE (conteucbl.f:333): the block always prints
  printf("\nEigenvalue = %s %s  %s\n\n", E(cc[0], 17, 8).c_str(), E(cc[1], 17, 8).c_str(), I(q, 5).c_str());
  printf("  Eigenvalue iterations = %s\n  |error| = %s\n  |c-cm1| = %s\n\n", I(passes, 5).c_str(),
         ES(resid[0], 17, 8).c_str(), ES(resid[1], 17, 8).c_str());
  c = cd(cc[0], cc[1]);
  qend = q;
}

static int run_conteucbl(osb_ctx *ctx) {
  int nbl = 1000, nstep = 1000, num = 1, omega = 0, saddle = 0;  // conteucbl.f:56-76
  double testalpha = 10.0, Re = 580., beta = 0.0, alphar1 = 0.179, alphai1 = 0.0;
  double cr = 0.36413124E+00, ci = 0.79527387E-02, ymin = 0.0, ymax = 17.0, f2p = 0.5, delta = 0.0, deltauc = 0.0, uc = 0.0;
  printf("\n          Solve Orr-Sommerfeld (Shooting)\n\n");
  printf("\n Read from keyboard, file, default (k,f,d) ==> ");
  fflush(stdout);
  const std::string input = getline80();
  const char ch = input.empty() ? ' ' : input[0];
  if (ch == 'k' || ch == 'K') {
    printf("\n Enter number of profile points ==> "); std::cin >> nbl;
    printf("\n Enter number of integration steps ==> "); std::cin >> nstep;
    printf("\n Enter Test Alpha ==> "); std::cin >> testalpha;
    printf("\n Enter Reynolds number ==> "); std::cin >> Re;
    printf("\n Enter beta ==> "); std::cin >> beta;
    printf("\n Enter guess for BL second deriv. at wall ==> "); std::cin >> f2p;
    printf("\n Enter alpha_r ==> "); std::cin >> alphar1;
    printf("\n Enter alpha_i ==> "); std::cin >> alphai1;
    printf("\n Enter guess for (c_r, c_i) ==> "); std::cin >> cr >> ci;
    printf("\n Enter Ymin and Ymax ==> "); std::cin >> ymin >> ymax;
    printf("\n Enter reference frame velocity u_c ==> "); std::cin >> uc;
    printf("\n Enter delta ==> "); std::cin >> delta;
    printf("\n Enter deltauc ==> "); std::cin >> deltauc;
  } else if (ch == 'f' || ch == 'F') {
    printf("\n Enter filename ==> ");
    fflush(stdout);
    std::string infile = getline80();
    while (!infile.empty() && infile.back() == ' ') infile.pop_back();
    Deck d;
    if (!d.open(infile)) { printf("\n\n Error in input file...\n\n\n"); return 0; }
    nbl = (int)d.geti(); nstep = (int)d.geti(); testalpha = d.getr(); Re = d.getr(); beta = d.getr(); f2p = d.getr();
    alphar1 = d.getr(); alphai1 = d.getr(); cr = d.getr(); ci = d.getr(); ymin = d.getr(); ymax = d.getr();
    uc = d.getr(); delta = d.getr(); deltauc = d.getr(); num = (int)d.geti(); omega = (int)d.geti(); saddle = (int)d.geti();
  }
  if (nbl > 20000) { printf("\n\n N > Idim...Stopping\n\n\n"); return 0; }
  cd alpha(alphar1, alphai1), c(cr, ci);
  if (omega == 1) c = c / alpha;  // conteucbl.f:186
  printf("\n nbl = %s\n", I(nbl, 5).c_str());
  printf(" nstep = %s\n", I(nstep, 5).c_str());
  printf(" Test Alpha = %s\n", E(testalpha, 20, 10).c_str());
  printf(" Re = %s\n", E(Re, 20, 10).c_str());
  printf(" Beta = %s\n", E(beta, 20, 10).c_str());
  printf(" f2p = %s\n", E(f2p, 20, 10).c_str());
  printf(" Ymin = %s  Ymax = %s\n", E(ymin, 20, 10).c_str(), E(ymax, 20, 10).c_str());
  printf(" alpha = (%s, %s)\n", E(alphar1, 20, 10).c_str(), E(alphai1, 20, 10).c_str());
  printf(" c = (%s, %s)\n", E(c.real(), 20, 10).c_str(), E(c.imag(), 20, 10).c_str());
  printf(" Uc = %s\n", E(uc, 20, 10).c_str());
  printf(" delta = %s\n", E(delta, 20, 10).c_str());
  printf(" deltauc = %s\n", E(deltauc, 20, 10).c_str());
  printf(" num = %s\n\n", I(num, 5).c_str());
  Re = Re * std::sqrt(2.);
  double ar[3] = {alphar1, alphar1 + delta, alphar1 - delta}, ai[3] = {alphai1, alphai1 - delta, alphai1 - delta};
  // SOLVE_BL (SRK4) then SHIFT_BL; the un-shifted tables play the role of Uorg
  osb_profile *porg = nullptr, *prof = nullptr;
  CK(ctx, osb_profile_blasius(ctx, OSB_FLOW_CONTEBL, OSB_INT_RK4, nbl, ymin, ymax, f2p, beta, &porg));
  CK(ctx, osb_profile_shift(ctx, porg, uc, &prof));
  printf(" Velocity Profile completed...\n\n");
  osb_shoot_opts o;
  CK(ctx, osb_shoot_defaults(OSB_FLOW_CONTEBL, &o));
  o.integrator = OSB_INT_RK4; o.nstep = nstep; o.testalpha = testalpha; o.ymin = ymin; o.ymax = ymax;
  int qend = 0;
  if (saddle == 1) {
    cd alphas(10., 10.), eigenvalue[3], alp[3];
    for (int k = 1; k <= num; ++k) {
      osb_profile_free(ctx, prof);
      CK(ctx, osb_profile_shift(ctx, porg, uc, &prof));  // call SHIFT_BL
      cd resid(10., 0.);
      int icount = 0;
      while (std::abs(resid) > 1.E-5) {
        icount = icount + 1;
        cd alphasold = alphas;
        for (int i = 0; i < 3; ++i) {
          alphasold = alphas;
          const cd a = cd(ar[i], ai[i]) * std::sqrt(2.);
          alp[i] = a / std::sqrt(2.);
          conteucbl_solve(ctx, o, prof, a, Re, c, qend);  // c is carried from call to call (COMMON /eig/)
          eigenvalue[i] = c * alp[i];
        }
        const cd A = cd(0., -1.) * (eigenvalue[1] - eigenvalue[0]), B = cd(0., -1.) * (eigenvalue[2] - eigenvalue[0]);
        alphas = (A * (alp[2] * alp[2] - alp[0] * alp[0]) - B * (alp[1] * alp[1] - alp[0] * alp[0])) /
                 (2. * B * (alp[0] - alp[1]) - 2. * A * (alp[0] - alp[2]));
        resid = alphas - alphasold;
        if (k == 1 || icount > 4)
          printf(" ==> %s    %s    %s    \n", E(alphas.real(), 17, 10).c_str(), E(alphas.imag(), 17, 10).c_str(),
                 E(std::abs(resid), 17, 10).c_str());
        ar[0] = alphas.real(); ai[0] = alphas.imag();
        ar[1] = ar[0] + delta; ar[2] = ar[0] - delta;
        ai[1] = ai[0] - delta; ai[2] = ai[0] - delta;
        if (icount > 200) { fprintf(stderr, "saddle iteration did not converge\n"); return 1; }
      }
      printf(" %s  %s    %s    %s    %s    \n", fortfmt::F(uc, 6, 4).c_str(), E(eigenvalue[0].real(), 17, 10).c_str(),
             E(eigenvalue[0].imag(), 17, 10).c_str(), E(alphas.real(), 17, 10).c_str(), E(alphas.imag(), 17, 10).c_str());
      uc = uc - deltauc;
    }
  } else {
    const cd a = cd(ar[0], ai[0]) * std::sqrt(2.);
    conteucbl_solve(ctx, o, prof, a, Re, c, qend);
    printf(" %s %s %s %s \n", ES(a.real(), 17, 9, 3).c_str(), ES(a.imag(), 17, 9, 3).c_str(), ES(c.real(), 17, 9, 3).c_str(),
           ES(c.imag(), 17, 9, 3).c_str());
  }
  osb_profile_free(ctx, prof);
  osb_profile_free(ctx, porg);
  return 0;
}

// ---------------------------------------------------------------- conte
static int run_conte(osb_ctx *ctx) {
  int n = 2000;  // conte.f:61-68
  double Re = 10000.0, alphar = 1.0, alphai = 0.0, cr = 0.23753, ci = 0.00374, ymin = -1.0, ymax = 1.0;
  printf("\n          Solve Orr-Sommerfeld for Channel using shooting\n\n");
  printf("\n Read from keyboard, file, default (k,d) ==> ");
  fflush(stdout);
  const std::string input = getline80();
  const char ch = input.empty() ? ' ' : input[0];
  if (ch == 'k' || ch == 'K') {
    printf("\n Enter number of steps ==> "); std::cin >> n;
    printf("\n Enter Reynolds number ==> "); std::cin >> Re;
    printf("\n Enter alpha_r ==> "); std::cin >> alphar;
    printf("\n Enter alpha_i ==> "); std::cin >> alphai;
    printf("\n Enter guess for (c_r, c_i) ==> "); std::cin >> cr >> ci;
  }
  printf("\n n = %s\n", I(n, 5).c_str());
  printf(" Re = %s\n", E(Re, 20, 10).c_str());
  printf(" alpha = (%s, %s)\n", E(alphar, 20, 10).c_str(), E(alphai, 20, 10).c_str());
  printf(" c = (%s, %s)\n\n", E(cr, 20, 10).c_str(), E(ci, 20, 10).c_str());
  const double h = (ymax - ymin) / (double)n;
  if (n > 10000) { printf("\n\n N > Idim...Stopping\n\n\n"); return 0; }
  {
    FILE *f = fopen("fort.10", "w");  // conte.f:288-294: 2x resolved mesh
    for (int i = 0; i <= 2 * n; ++i) {
      const double y = ymin + i * h / 2.0;
      fprintf(f, " %s %s %s \n", ES(y, 16, 8, 3).c_str(), ES(1.0 - y * y, 16, 8, 3).c_str(), ES(-2.0, 16, 8, 3).c_str());
    }
    fclose(f);
  }
  printf(" Channel velocity profile completed...\n\n");
  osb_profile *prof = nullptr;
  CK(ctx, osb_profile_channel(ctx, &prof));
  osb_shoot_opts o;
  CK(ctx, osb_shoot_defaults(OSB_FLOW_CONTE, &o));
  o.nstep = n; o.ymin = ymin; o.ymax = ymax;
  double al[2] = {alphar, alphai}, cc[2] = {cr, ci}, resid[2], hist[4 * 21];
  int32_t passes, qend, status;
  CK(ctx, osb_shoot_temporal(ctx, &o, prof, 1, al, &Re, cc, &passes, &qend, &status, resid, hist, 21));
  for (int i = 0; i < passes && i < 21; ++i)  // shoot.f:322-323 (shows the UPDATED c)
    printf(" %s %s %s    %s %s   \n", I(i + 1, 4).c_str(), ES(hist[4 * i], 16, 8, 3).c_str(), ES(hist[4 * i + 1], 16, 8, 3).c_str(),
           ES(hist[4 * i + 2], 16, 8, 3).c_str(), ES(hist[4 * i + 3], 16, 8, 3).c_str());
  printf("\nEigenvalue = %s %s  %s\n\n", E(cc[0], 17, 8, 3).c_str(), E(cc[1], 17, 8, 3).c_str(), I(qend, 5).c_str());
  // conte does not decrement icount before printing (shoot.f:332)
  printf("  Eigenvalue iterations = %s\n  |error| = %s\n  |c-cm1| = %s\n\n", I(passes + 1, 5).c_str(),
         ES(resid[0], 17, 8, 3).c_str(), ES(resid[1], 17, 8, 3).c_str());
  std::vector<double> y((size_t)(n + 1) * 8);
  double vmax[2];
  int32_t q2;
  CK(ctx, osb_eigfun(ctx, &o, prof, 1, al, &Re, cc, nullptr, y.data(), vmax, &q2));
  write_eigfun(y, cd(vmax[0], vmax[1]), n, ymax, ymin, h, false, 16, true);
  osb_profile_free(ctx, prof);
  return status;
}

// ---------------------------------------------------------------- orrsom (deck on stdin, orrsom.f:63-91)
struct StdinDeck {
  bool first = true;
  std::string next_line() {
    std::string s;
    if (first) {
      first = false;
      do { if (!std::getline(std::cin, s)) return ""; } while (!s.empty() && s[0] == '#');
      return s;
    }
    std::getline(std::cin, s);
    return s;
  }
  long long geti() {
    std::string fld; long long v = 0;
    if (!fortfmt::deck_field(next_line(), fld) || !fortfmt::read_I10(fld, v)) Deck::input_error();
    return v;
  }
  double getr() {
    std::string fld; double v = 0;
    if (!fortfmt::deck_field(next_line(), fld) || !fortfmt::read_G20_10(fld, v)) Deck::input_error();
    return v;
  }
};

static int run_orrsom(osb_ctx *ctx) {
  printf("\n          Solve Orr-Sommerfeld (Shooting) Temporal Problem\n\n");
  StdinDeck d;
  const int n = (int)d.geti(), nstep = (int)d.geti();
  const double testalpha = d.getr();
  double Re = d.getr();
  const double beta = d.getr(), f2p = d.getr(), alphar = d.getr(), alphai = d.getr(), cr = d.getr(), ci = d.getr(),
               ymin = d.getr(), ymax = d.getr();
  printf("\n n = %s\n", I(n, 5).c_str());
  printf(" nstep = %s\n", I(nstep, 5).c_str());
  printf(" Test Alpha = %s\n", E(testalpha, 20, 10).c_str());
  printf(" Re = %s\n", E(Re, 20, 10).c_str());
  printf(" Beta = %s\n", E(beta, 20, 10).c_str());
  printf(" f2p = %s\n", E(f2p, 20, 10).c_str());
  printf(" Ymin = %s  Ymax = %s\n", E(ymin, 20, 10).c_str(), E(ymax, 20, 10).c_str());
  printf(" alpha = (%s, %s)\n", E(alphar, 20, 10).c_str(), E(alphai, 20, 10).c_str());
  printf(" c = (%s, %s)\n\n", E(cr, 20, 10).c_str(), E(ci, 20, 10).c_str());
  cd alpha(alphar, alphai);
  const double h = (ymax - ymin) / (double)n;
  Re = Re * std::sqrt(2.);  // orrsom.f:129-130
  alpha = alpha * std::sqrt(2.);
  if (n > 20000) { printf("\n\n N > Idim...Stopping\n\n\n"); return 0; }
  osb_profile *prof = nullptr;
  CK(ctx, osb_profile_blasius(ctx, OSB_FLOW_ORRSOM, OSB_INT_RK4, n, ymin, ymax, f2p, beta, &prof));
  const Tables t = fetch_tables(ctx, prof, n);
  {
    FILE *f = fopen("fort.10", "w");  // orrsom.f:790-797
    for (int i = 0; i <= n; ++i) {
      const double y = ymin + i * h;
      const double us = speval_linear(t.ydat, t.U, t.Uspl, y), d2us = speval_linear(t.ydat, t.d2U, t.d2Uspl, y);
      fprintf(f, " %s %s %s %s %s \n", ES(y, 20, 12, 3).c_str(), ES(us, 20, 12, 3).c_str(), ES(d2us, 20, 12, 3).c_str(),
              ES(t.U[i], 20, 12, 3).c_str(), ES(t.d2U[i], 20, 12, 3).c_str());
    }
    fclose(f);
  }
  printf(" Velocity Profile completed...\n\n");
  osb_shoot_opts o;
  CK(ctx, osb_shoot_defaults(OSB_FLOW_ORRSOM, &o));
  o.nstep = nstep; o.testalpha = testalpha; o.ymin = ymin; o.ymax = ymax;
  double al[2] = {alpha.real(), alpha.imag()}, cc[2] = {cr, ci}, resid[2], hist[4 * 20];
  int32_t passes, qend, status;
  CK(ctx, osb_shoot_temporal(ctx, &o, prof, 1, al, &Re, cc, &passes, &qend, &status, resid, hist, 20));
  for (int i = 0; i < passes && i < 20; ++i)  // orrsom.f:442-444
    printf(" %s%s%s   %s%s   \n", I(i + 1, 4).c_str(), E(hist[4 * i], 17, 8).c_str(), E(hist[4 * i + 1], 17, 8).c_str(),
           E(hist[4 * i + 2], 17, 8).c_str(), E(hist[4 * i + 3], 17, 8).c_str());
  printf("\nEigenvalue = %s %s  %s\n\n", E(cc[0], 17, 8).c_str(), E(cc[1], 17, 8).c_str(), I(qend, 5).c_str());
  printf("Second pass: computing and writing eigenfunction\n");
  std::vector<double> y((size_t)(nstep + 1) * 8);
  double vmax[2];
  int32_t q2;
  CK(ctx, osb_eigfun(ctx, &o, prof, 1, al, &Re, cc, nullptr, y.data(), vmax, &q2));
  const cd vm(vmax[0], vmax[1]);
  write_eigfun(y, vm, nstep, ymax, ymin, (ymax - ymin) / nstep, true, 17, true);
  {
    FILE *f = fopen("fort.15", "w");  // orrsom.f:500: t, Re(phi'/max), Re(-i alpha phi/max)
    const double hh = (ymax - ymin) / nstep;
    for (int k = 0; k <= nstep; ++k) {
      const cd y1(y[2 * (4 * k)], y[2 * (4 * k) + 1]), y2(y[2 * (4 * k + 1)], y[2 * (4 * k + 1) + 1]);
      fprintf(f, " %s %s %s \n", ES(ymax - hh * k, 17, 8, 3).c_str(), ES((y2 / vm).real(), 17, 8, 3).c_str(),
              ES((cd(0., -1.) * alpha * y1 / vm).real(), 17, 8, 3).c_str());
    }
    fclose(f);
  }
  osb_profile_free(ctx, prof);
  return status;
}

// ---------------------------------------------------------------- orrspace
static int run_orrspace(osb_ctx *ctx) {
  printf("\n          Solve Orr-Sommerfeld (Shooting)\n\n");
  printf("     This version solves the spatial problem\n\n");
  printf("\n Enter filename ==> ");
  fflush(stdout);
  std::string infile = getline80();
  while (!infile.empty() && infile.back() == ' ') infile.pop_back();
  Deck d;
  if (!d.open(infile)) { printf("\n\n Error in input file...\n\n\n"); return 0; }
  const int n = (int)d.geti(), nstep = (int)d.geti();
  const double testalpha = d.getr();
  double Re = d.getr();
  const double beta = d.getr(), f2p = d.getr(), alphar = d.getr(), alphai = d.getr(), omegar = d.getr(), omegai = d.getr();
  double ymin = d.getr(), ymax = d.getr();
  const int niter = (int)d.geti();
  double domega = d.getr();
  const int eigfun = (int)d.geti();
  printf("\n n = %s\n", I(n, 5).c_str());
  printf(" nstep = %s\n", I(nstep, 5).c_str());
  printf(" Test Alpha = %s\n", E(testalpha, 20, 10).c_str());
  printf(" Re = %s\n", E(Re, 20, 10).c_str());
  printf(" Beta = %s\n", E(beta, 20, 10).c_str());
  printf(" f2p = %s\n", E(f2p, 20, 10).c_str());
  printf(" Ymin = %s  Ymax = %s\n", E(ymin, 20, 10).c_str(), E(ymax, 20, 10).c_str());
  printf(" alpha = (%s, %s)\n", E(alphar, 20, 10).c_str(), E(alphai, 20, 10).c_str());
  printf(" omega = (%s, %s)\n\n", E(omegar, 20, 10).c_str(), E(omegai, 20, 10).c_str());
  cd alpha(alphar, alphai), omega(omegar, omegai);
  printf("\n Note:  changing nondimsensionalization to displacement thickness to match Mack...\n\n");
  const double fact = std::sqrt(2.0) / 1.7207876, facti = 1.0 / fact;  // orrspace.f:160-168
  Re = Re * fact; alpha = alpha * fact; omega = omega * fact; domega = domega * fact;
  ymin = ymin * fact; ymax = ymax * fact;
  const double hbl = (ymax - ymin) / (double)n;
  if (n > 20000) { printf("\n\n N > Idim...Stopping\n\n\n"); return 0; }
  osb_profile *prof = nullptr;
  CK(ctx, osb_profile_blasius(ctx, OSB_FLOW_ORRSPACE, OSB_INT_RK4, n, ymin, ymax, f2p, beta, &prof));
  const Tables t = fetch_tables(ctx, prof, n);
  {
    FILE *f = fopen("fort.10", "w");  // orrspace.f:804-810, format 5(1PE20.13E2,1x)
    for (int i = 0; i <= n; ++i) {
      const double y = ymin + i * hbl;
      const double us = speval_linear(t.ydat, t.U, t.Uspl, y), d2us = speval_linear(t.ydat, t.d2U, t.d2Uspl, y);
      fprintf(f, " %s %s %s %s %s \n", ES(y / std::sqrt(2.) * 1.7207876, 20, 13, 2).c_str(), ES(us, 20, 13, 2).c_str(),
              ES(d2us, 20, 13, 2).c_str(), ES(t.U[i], 20, 13, 2).c_str(), ES(t.d2U[i], 20, 13, 2).c_str());
    }
    fclose(f);
  }
  printf(" Velocity Profile completed...\n\n");
  osb_shoot_opts o;
  CK(ctx, osb_shoot_defaults(OSB_FLOW_ORRSPACE, &o));
  o.nstep = nstep; o.testalpha = testalpha; o.ymin = ymin; o.ymax = ymax;
  // the whole omega loop of orrspace.f:186-195 is one call: continuation runs inside the kernel
  std::vector<double> aout((size_t)niter * 2);
  std::vector<int32_t> passes(niter), qend(niter), status(niter);
  double om0[2] = {omega.real(), omega.imag()}, a0[2] = {alpha.real(), alpha.imag()};
  const int hcap = 50;  // orrspace's pass limit (orrspace.f:277)
  std::vector<double> hist((size_t)niter * hcap * 4);
  CK(ctx, osb_shoot_spatial_lines(ctx, &o, prof, 1, niter, &Re, om0, &domega, a0, aout.data(), passes.data(), qend.data(),
                                  status.data(), hist.data(), hcap));
  FILE *f17 = fopen("fort.17", "w");
  const double pfac = 1.7207877 / std::sqrt(2.);  // orrspace.f:485-486,495-498 (note the ...77)
  int bad = 0;
  for (int i = 0; i < niter; ++i) {
    const cd om(omega.real() + i * domega, omega.imag()), al(aout[2 * i], aout[2 * i + 1]);
    for (int k = 0; k < passes[i] && k < hcap; ++k) {  // per-pass table, orrspace.f:485-488 (format 30)
      const double *hr = &hist[((size_t)i * hcap + k) * 4];
      printf(" %s    %s %s    %s %s    %s\n", I(k + 1, 2).c_str(), E(hr[0] * pfac, 15, 8).c_str(), E(hr[1] * pfac, 15, 8).c_str(),
             E(hr[2], 15, 8).c_str(), E(hr[3], 15, 8).c_str(), E(std::hypot(hr[2], hr[3]), 15, 8).c_str());
    }
    printf("\nOmega = (%s %s)    Alpha = (%s %s)\n\n", E(om.real() * pfac, 15, 8).c_str(), E(om.imag() * pfac, 15, 8).c_str(),
           E(al.real() * pfac, 15, 8).c_str(), E(al.imag() * pfac, 15, 8).c_str());
    fprintf(f17, " %s  %s  %s  \n", E((om * facti).real(), 17, 8).c_str(), E((al * facti).real(), 17, 8).c_str(),
            E((al * facti).imag(), 17, 8).c_str());
    bad |= status[i];
  }
  fclose(f17);
  if (eigfun == 1) {  // eigenfunction of the LAST point (each CONTE call overwrites fort.11-14)
    const cd om(omega.real() + (niter - 1) * domega, omega.imag());
    double omr[2] = {om.real(), om.imag()}, ev[2] = {aout[2 * (niter - 1)], aout[2 * (niter - 1) + 1]};
    std::vector<double> y((size_t)(nstep + 1) * 8);
    double vmax[2];
    int32_t q2;
    CK(ctx, osb_eigfun(ctx, &o, prof, 1, nullptr, &Re, ev, omr, y.data(), vmax, &q2));
    write_eigfun(y, cd(vmax[0], vmax[1]), nstep, ymax, ymin, (ymax - ymin) / nstep, true, 16, false, std::sqrt(2.) / 1.7207876);
    // u and v disturbance velocities (orrspace.f:547-578): fort.15 / fort.16 hold u = phi' and v = -i alpha phi
    // at x = 0 and nine phases of one half period, fort.18 both normalised by the u of largest modulus
    {
      const double pi = 3.14159265358979323846, h = (ymax - ymin) / nstep, ts = std::sqrt(2.) / 1.7207876;  // orrspace.f:267
      const cd alpha(ev[0], ev[1]), mx(vmax[0], vmax[1]), iu(0.0, 1.0);
      const int ntime = 8;
      const double x = 0.0, dtime = pi / om.real() / ntime;
      auto Y = [&](int comp, int k) { return cd(y[2 * (4 * k + comp)], y[2 * (4 * k + comp) + 1]); };
      FILE *f15 = fopen("fort.15", "w"), *f16 = fopen("fort.16", "w"), *f18 = fopen("fort.18", "w");
      for (int k = 0; k <= nstep; ++k) {
        const double t = (ymax - h * k) * ts;
        fprintf(f15, "%s ", ES(t, 16, 8, 3).c_str());
        fprintf(f16, "%s ", ES(t, 16, 8, 3).c_str());
        for (int i = 0; i <= ntime; ++i) {
          const cd ph = std::exp(iu * (alpha * x - om * (double)i * dtime));
          fprintf(f15, "%s ", ES((Y(1, k) / mx * ph).real(), 16, 8, 3).c_str());
          fprintf(f16, "%s ", ES((cd(0.0, -1.0) * alpha * Y(0, k) / mx * ph).real(), 16, 8, 3).c_str());
        }
        fprintf(f15, "\n");
        fprintf(f16, "\n");
      }
      cd umax(0.0, 0.0);
      for (int k = 0; k <= nstep; ++k)
        if (std::abs(Y(1, k)) > std::abs(umax)) umax = Y(1, k);
      for (int k = 0; k <= nstep; ++k) {
        const double t = (ymax - h * k) * ts;
        const cd uu = Y(1, k) / umax, vv = cd(0.0, -1.0) * alpha * Y(0, k) / umax;
        fprintf(f18, "%s %s %s %s %s \n", ES(t, 16, 8, 3).c_str(), ES(uu.real(), 16, 8, 3).c_str(), ES(uu.imag(), 16, 8, 3).c_str(),
                ES(vv.real(), 16, 8, 3).c_str(), ES(vv.imag(), 16, 8, 3).c_str());
      }
      fclose(f15); fclose(f16); fclose(f18);
    }
  }
  osb_profile_free(ctx, prof);
  return bad;
}

// ---------------------------------------------------------------- orrfdchan
static int run_orrfdchan(osb_ctx *ctx) {
  int n, ians;
  double Re, minar, maxar, incar, minai, maxai, incai;
  std::string filename;
  printf("\n\n          Solve Orr-Sommerfeld (Finite-difference)\n");
  printf("\n Enter the number of mesh points ==> "); std::cin >> n;
  printf("\n Enter Reynolds number ==> "); std::cin >> Re;
  printf("\n Enter alpha_r (min,max,inc) ==> "); std::cin >> minar >> maxar >> incar;
  printf("\n Enter alpha_i (min,max,inc) ==> "); std::cin >> minai >> maxai >> incai;
  printf("\n Enter root filename ==> "); std::cin >> filename;
  printf("\n Print eigenvectors (1,0) ==> "); std::cin >> ians;
  const int nar = (int)std::trunc((maxar - minar) / incar) + 1, nai = (int)std::trunc((maxai - minai) / incai) + 1;  // :89-90
  printf("\n  nar = %s  nai = %s\n", I(nar, 5).c_str(), I(nai, 5).c_str());
  if (nar > 100 || nai > 100 || nar < 1 || nai < 1) { fprintf(stderr, "sweep exceeds the reference's 100x100 arrays\n"); return 1; }
  std::vector<double> alphar(nar), alphai(nai);
  for (int i = 1; i <= nar; ++i) alphar[i - 1] = minar + (i - 1.) * incar;
  for (int i = 1; i <= nai; ++i) alphai[i - 1] = minai + (i - 1.) * incai;
  auto write5 = [](FILE *f, const std::vector<double> &v) {  // format 110: 5(ES16.8E3,1x) per record
    for (size_t k = 0; k < v.size(); ++k) {
      fprintf(f, "%s ", ES(v[k], 16, 8, 3).c_str());
      if (k % 5 == 4 || k + 1 == v.size()) fprintf(f, "\n");
    }
  };
  {
    FILE *g = fopen((filename + ".g").c_str(), "w");
    fprintf(g, "%s %s \n", I(nar, 5).c_str(), I(nai, 5).c_str());
    std::vector<double> v;
    for (int j = 0; j < nai; ++j) for (int i = 0; i < nar; ++i) v.push_back(alphar[i]);
    for (int j = 0; j < nai; ++j) for (int i = 0; i < nar; ++i) v.push_back(alphai[j]);
    write5(g, v);
    fclose(g);
  }
  osb_chan *chan = nullptr;
  CK(ctx, osb_chan_create(ctx, OSB_DISC_FD, n, &chan));
  {
    std::vector<double> eta(n + 1), u(n + 1), d2u(n + 1);
    CK(ctx, osb_chan_tables(ctx, chan, eta.data(), u.data(), d2u.data(), nullptr, nullptr));
    FILE *f = fopen("fort.10", "w");  // orrfdchan.f:262-266
    for (int i = 0; i <= n; ++i)
      fprintf(f, " %s    %s    %s    %s    \n", ES(eta[i], 16, 8, 3).c_str(), ES(u[i], 16, 8, 3).c_str(),
              ES(-2.0 * eta[i], 16, 8, 3).c_str(), ES(d2u[i], 16, 8, 3).c_str());
    fclose(f);
  }
  // the (j, i) loop nest of orrfdchan.f:116-128 hoisted into one batched call
  const int64_t npts = (int64_t)nar * nai;
  std::vector<double> al(2 * npts), om(2 * npts);
  std::vector<int32_t> it(npts), st(npts);
  for (int j = 0; j < nai; ++j) for (int i = 0; i < nar; ++i) { al[2 * (j * nar + i)] = alphar[i]; al[2 * (j * nar + i) + 1] = alphai[j]; }
  const bool print = (ians == 1);
  std::vector<double> ev, av;
  if (print) { ev.resize((size_t)npts * (n + 1) * 2); av.resize((size_t)npts * (n + 1) * 2); }
  CK(ctx, osb_chan_eig(ctx, chan, Re, npts, al.data(), nullptr, om.data(), print ? ev.data() : nullptr,
                       print ? av.data() : nullptr, OSB_ADJ_INVBA, it.data(), st.data()));
  std::vector<double> eta;
  if (print) {
    eta.resize(n + 1);
    CK(ctx, osb_chan_tables(ctx, chan, eta.data(), nullptr, nullptr, nullptr, nullptr));
  }
  FILE *dat = fopen((filename + ".dat").c_str(), "w");
  FILE *f20 = print ? fopen("fort.20", "w") : nullptr;
  int bad = 0;
  for (int64_t p = 0; p < npts; ++p) {
    if (print) {
      // the `if (print)` block of SOLVE_ORR_SOM (orrfdchan.f:832-873): the spectrum listing, then -- whatever
      // index is typed -- the eigenvector of the KEPT mode (iloc), normalised by its largest component, in fort.11.
      const cd alpha(al[2 * p], al[2 * p + 1]), w(om[2 * p], om[2 * p + 1]);
      printf("\n Residual = %s\n", F(0.0, 17, 10).c_str());
      printf("\n %s    %s    %s    \n", E(Re, 17, 10).c_str(), E(alpha.real(), 17, 10).c_str(), E(alpha.imag(), 17, 10).c_str());
      printf("\n        w_r                w_i              count\n\n");
      {
        // the reference lists ZGEEV's raw output order (and one entry past it, orrfdchan.f:834-838); here
        // the device spectrum in the sorted order of its SORT pass
        std::vector<double> sp((size_t)(n - 1) * 2);
        int32_t sst = 0;
        CK(ctx, osb_chan_spectrum(ctx, chan, Re, 1, al.data() + 2 * p, sp.data(), &sst));
        for (int i = 0; i < n - 1; ++i) {
          printf(" %s    %s    %s\n", E(sp[2 * i], 17, 10).c_str(), E(sp[2 * i + 1], 17, 10).c_str(), I(i + 1, 5).c_str());
          const cd cph = cd(sp[2 * i], sp[2 * i + 1]) / alpha;
          fprintf(f20, " %s    %s    \n", E(cph.real(), 17, 10).c_str(), E(cph.imag(), 17, 10).c_str());
        }
        bad |= sst;
        (void)w;
      }
      int which = 0;
      printf("\n Eigenvector for which eigenvalue (0 quits) ==> ");
      while (std::cin >> which && which != 0) {
        const double *v = ev.data() + (size_t)p * (n + 1) * 2, *a = av.data() + (size_t)p * (n + 1) * 2;
        // escale (orrfdchan.f:849-856): the largest component of ZGEEV's unit-norm right vector, which ZGEEV makes
        // real and positive: 1 / |v|_2 of the max-normalised vector
        double nrm = 0.0;
        for (int j = 0; j <= n; ++j) nrm += v[2 * j] * v[2 * j] + v[2 * j + 1] * v[2 * j + 1];
        printf(" escale =  (%.16g,%.16g)\n", 1.0 / std::sqrt(nrm), 0.0);
        FILE *f = fopen("fort.11", "w");  // format 60
        for (int j = 0; j <= n; ++j) {
          // columns 4-5: avec(j, ind(iloc)), j = 0..N (orrfdchan.f:865-871): ZGEEV's left vector of inv(B)A, NOT shifted
          // onto the grid like the right vector, and two entries past the N-1 that ZGEEV wrote (zeros of the static array)
          const double ar = j < n - 1 ? a[2 * (j + 1)] : 0.0, ai = j < n - 1 ? a[2 * (j + 1) + 1] : 0.0;
          fprintf(f, " %s %s %s %s %s \n", ES(eta[j], 16, 8, 3).c_str(), ES(v[2 * j], 16, 8, 3).c_str(),
                  ES(v[2 * j + 1], 16, 8, 3).c_str(), ES(ar, 16, 8, 3).c_str(), ES(ai, 16, 8, 3).c_str());
        }
        fclose(f);
        printf("\n Eigenvector for which eigenvalue (0 quits) ==> ");
      }
    }
    const std::string row = " " + ES(al[2 * p], 16, 8, 3) + " " + ES(al[2 * p + 1], 16, 8, 3) + " " + ES(om[2 * p], 16, 8, 3) + " " +
                            ES(om[2 * p + 1], 16, 8, 3) + " ";
    fprintf(dat, "%s\n", row.c_str());
    printf("%s\n", row.c_str());
    bad |= st[p];
  }
  fclose(dat);
  if (f20) fclose(f20);
  {
    FILE *q = fopen((filename + ".q").c_str(), "w");  // orrfdchan.f:130-138
    fprintf(q, "%s %s \n", I(nar, 5).c_str(), I(nai, 5).c_str());
    write5(q, {0.0, 0.0, 0.0, 0.0});
    std::vector<double> v;
    for (int64_t p = 0; p < npts; ++p) v.push_back(1.0);
    for (int64_t p = 0; p < npts; ++p) v.push_back(om[2 * p]);
    for (int64_t p = 0; p < npts; ++p) v.push_back(om[2 * p + 1]);
    for (int64_t p = 0; p < npts; ++p) v.push_back(1.0);
    write5(q, v);
    fclose(q);
  }
  osb_chan_free(ctx, chan);
  return bad;
}

// ---------------------------------------------------------------- orrfdbl / orrcolbl
// Fortran sequential unformatted record (gfortran: 4-byte length before and after the payload)
static void f_record(FILE *f, const void *data, size_t bytes) {
  const int32_t len = (int32_t)bytes;
  fwrite(&len, 4, 1, f);
  fwrite(data, 1, bytes, f);
  fwrite(&len, 4, 1, f);
}

// orrfdbl.f:58-141 / orrcolbl.f:60-157.  The reference programs read the mean profile from a file they do
// not ship (INIT_BL_PROFILE, orrcolbl.f:366-388) and link against IMSL; the twin keeps their prompts and
// files and takes the Blasius profile of this engine's SOLVE_BL kernel as the mean flow.  orrcolbl's
// dw/dalpha column (adjoint formula, orrcolbl.f:815-824) is a central difference of two extra batched solves.
static int run_orrbl(osb_ctx *ctx, bool colloc) {
  int n, ians;
  double Re, lmap, minar, maxar, incar, minai, maxai, incai;
  std::string filename, profname;
  printf("\n\n          Solve Orr-Sommerfeld (%s)\n", colloc ? "Collocation" : "Finite-difference");
  printf("\n Enter the number of %s ==> ", colloc ? "modes" : "mesh points"); std::cin >> n;
  printf("\n Enter Reynolds number ==> "); std::cin >> Re;
  printf("\n Enter mapping metric ==> "); std::cin >> lmap;
  printf("\n Enter alpha_r (min,max,inc) ==> "); std::cin >> minar >> maxar >> incar;
  printf("\n Enter alpha_i (min,max,inc) ==> "); std::cin >> minai >> maxai >> incai;
  printf("\n Enter root filename ==> "); std::cin >> filename;
  printf("\n Print eigenvectors (1,0) ==> "); std::cin >> ians;
  const int nar = (int)std::trunc((maxar - minar) / incar) + 1, nai = (int)std::trunc((maxai - minai) / incai) + 1;
  printf("\n  nar = %s  nai = %s\n", I(nar, 5).c_str(), I(nai, 5).c_str());
  if (nar > 100 || nai > 100 || nar < 1 || nai < 1) { fprintf(stderr, "sweep exceeds the reference's 100x100 arrays\n"); return 1; }
  if (n < 8 || n > (colloc ? 256 : 512)) { fprintf(stderr, "n exceeds the reference's idim\n"); return 1; }
  std::vector<double> alphar(nar), alphai(nai);
  for (int i = 1; i <= nar; ++i) alphar[i - 1] = minar + (i - 1.) * incar;
  for (int i = 1; i <= nai; ++i) alphai[i - 1] = minai + (i - 1.) * incai;
  const int64_t npts = (int64_t)nar * nai;
  const int64_t nxy[2] = {nar, nai};  // -fdefault-integer-8 (gcc.mak:79)
  {
    FILE *g = fopen((filename + ".g").c_str(), "wb");
    f_record(g, nxy, sizeof(nxy));
    std::vector<double> v;
    for (int j = 0; j < nai; ++j) for (int i = 0; i < nar; ++i) v.push_back(alphar[i]);
    for (int j = 0; j < nai; ++j) for (int i = 0; i < nar; ++i) v.push_back(alphai[j]);
    f_record(g, v.data(), v.size() * sizeof(double));
    fclose(g);
  }
  // INIT_BL_PROFILE: prompt kept; the name is read and ignored (see above)
  printf("\n Read Mean Profile\n\n Enter filename ==> ");
  std::cin >> profname;
  const int disc = colloc ? OSB_DISC_CHEB : OSB_DISC_FD;
  const int nbl = 2000;
  const double ymax = 20.0;
  osb_profile *prof = nullptr;
  CK(ctx, osb_profile_blasius(ctx, OSB_FLOW_CONTEBL, OSB_INT_CK45, nbl, 0.0, ymax, 0.5, 0.0, &prof));
  std::vector<double> yd(nbl + 1), Ut(nbl + 1), Us(nbl + 1), y(n + 1), U(n + 1, 1.0);
  int32_t npass = 0;
  double f2p = 0.0;
  CK(ctx, osb_profile_get(ctx, prof, yd.data(), Ut.data(), Us.data(), nullptr, nullptr, &npass, &f2p));
  osb_profile_free(ctx, prof);
  if (osb_chan_bl_points(disc, n, lmap, y.data())) { fprintf(stderr, "bad grid parameters\n"); return 1; }
  int merge = 0;
  const double h = yd[1] - yd[0];
  for (int i = n; i >= 0; --i) {  // from the wall outwards, as orrcolbl.f:390-403
    if (!(y[i] <= ymax)) { merge = i; break; }
    int k = std::min(std::max((int)(y[i] / h), 0), nbl - 1);
    const double A = (yd[k + 1] - y[i]) / h, B = (y[i] - yd[k]) / h;  // SPEVAL, shoot.f:996-1013
    U[i] = A * Ut[k] + B * Ut[k + 1] + ((A * A * A - A) * Us[k] + (B * B * B - B) * Us[k + 1]) * (h * h) / 6.0;
  }
  printf("\n Merging at i = %s\n", I(merge, 4).c_str());
  osb_chan *chan = nullptr;
  CK(ctx, osb_chan_create_bl(ctx, disc, n, lmap, U.data(), merge, &chan));
  if (colloc) printf("\n");
  std::vector<double> al(2 * npts), om(2 * npts);
  std::vector<int32_t> it(npts), st(npts);
  for (int j = 0; j < nai; ++j) for (int i = 0; i < nar; ++i) { al[2 * (j * nar + i)] = alphar[i]; al[2 * (j * nar + i) + 1] = alphai[j]; }
  CK(ctx, osb_chan_eig(ctx, chan, Re, npts, al.data(), nullptr, om.data(), nullptr, nullptr, 0, it.data(), st.data()));
  std::vector<double> dw(2 * npts, 0.0);
  if (colloc) {
    std::vector<double> a2(4 * npts), s2(4 * npts), o2(4 * npts);
    std::vector<int32_t> i2(2 * npts), t2(2 * npts);
    std::vector<double> hh(npts);
    for (int64_t p = 0; p < npts; ++p) {
      hh[p] = 1.0e-6 * std::max(std::hypot(al[2 * p], al[2 * p + 1]), 1.0e-3);
      for (int sgn = 0; sgn < 2; ++sgn) {
        const int64_t e = 2 * p + sgn;
        a2[2 * e] = al[2 * p] + (sgn ? hh[p] : -hh[p]); a2[2 * e + 1] = al[2 * p + 1];
        s2[2 * e] = om[2 * p]; s2[2 * e + 1] = om[2 * p + 1];
      }
    }
    CK(ctx, osb_chan_eig(ctx, chan, Re, 2 * npts, a2.data(), s2.data(), o2.data(), nullptr, nullptr, 0, i2.data(), t2.data()));
    for (int64_t p = 0; p < npts; ++p) {
      dw[2 * p] = (o2[2 * (2 * p + 1)] - o2[2 * (2 * p)]) / (2.0 * hh[p]);
      dw[2 * p + 1] = (o2[2 * (2 * p + 1) + 1] - o2[2 * (2 * p) + 1]) / (2.0 * hh[p]);
    }
  }
  FILE *dat = fopen((filename + ".dat").c_str(), "w");
  int bad = 0;
  for (int64_t p = 0; p < npts; ++p) {
    std::string row = " ";
    const char *sep = colloc ? " " : "    ";  // format 60: 6(e17.10,1x) / 4(e17.10,4x)
    row += E(al[2 * p], 17, 10) + sep + E(al[2 * p + 1], 17, 10) + sep + E(om[2 * p], 17, 10) + sep + E(om[2 * p + 1], 17, 10) + sep;
    if (colloc) row += E(dw[2 * p], 17, 10) + sep + E(dw[2 * p + 1], 17, 10) + sep;
    fprintf(dat, "%s\n", row.c_str());
    printf("%s\n", row.c_str());
    bad |= st[p];
  }
  fclose(dat);
  {
    FILE *q = fopen((filename + ".q").c_str(), "wb");
    f_record(q, nxy, sizeof(nxy));
    const double z4[4] = {0.0, 0.0, 0.0, 0.0};
    f_record(q, z4, sizeof(z4));
    std::vector<double> v;
    for (int64_t p = 0; p < npts; ++p) v.push_back(1.0);
    for (int64_t p = 0; p < npts; ++p) v.push_back(om[2 * p]);
    for (int64_t p = 0; p < npts; ++p) v.push_back(om[2 * p + 1]);
    for (int64_t p = 0; p < npts; ++p) v.push_back(1.0);
    f_record(q, v.data(), v.size() * sizeof(double));
    fclose(q);
  }
  if (ians == 1) fprintf(stderr, "note: the eigenvector listing of SOLVE's print branch is not reproduced by this twin\n");
  osb_chan_free(ctx, chan);
  return bad;
}

// ---------------------------------------------------------------- orrcolchan
static int run_orrcolchan(osb_ctx *ctx) {
  int n;
  double Re, alphar, alphai;
  printf("\n\n          Solve Orr-Sommerfeld (Collocation)\n");
  printf("\n Enter the number of modes ==> "); std::cin >> n;
  printf("\n Enter Reynolds number ==> "); std::cin >> Re;
  printf("\n Enter alpha (R,I) ==> "); std::cin >> alphar >> alphai;
  osb_chan *chan = nullptr;
  CK(ctx, osb_chan_create(ctx, OSB_DISC_CHEB, n, &chan));
  std::vector<double> eta(n + 1), u(n + 1), d2u(n + 1);
  CK(ctx, osb_chan_tables(ctx, chan, eta.data(), u.data(), d2u.data(), nullptr, nullptr));
  {
    FILE *f = fopen("fort.10", "w");  // orrcolchan.f:235-240
    for (int i = 0; i <= n; ++i)
      fprintf(f, " %s %s %s %s \n", ES(eta[i], 16, 8, 3).c_str(), ES(u[i], 16, 8, 3).c_str(), ES(-2. * eta[i], 16, 8, 3).c_str(),
              ES(d2u[i], 16, 8, 3).c_str());
    fclose(f);
  }
  double al[2] = {alphar, alphai};
  const cd alpha(alphar, alphai);
  // the whole spectrum on the device (inv(B)A, Hessenberg, shifted QR).  The reference's (N+1) x (N+1) pencil
  // carries two more eigenvalues from its boundary rows -- infinite, which it maps to 0 (orrcolchan.f:466-472) --
  // and sorts everything by Im(w) with PIKSR2 (straight insertion: stable)
  std::vector<double> sp((size_t)(n - 1) * 2);
  int32_t sst = 0;
  CK(ctx, osb_chan_spectrum(ctx, chan, Re, 1, al, sp.data(), &sst));
  std::vector<cd> omega;
  omega.push_back(cd(0.0, 0.0));
  omega.push_back(cd(0.0, 0.0));
  for (int k = 0; k < n - 1; ++k) omega.push_back(cd(sp[2 * k], sp[2 * k + 1]));
  std::stable_sort(omega.begin(), omega.end(), [](const cd &x, const cd &y) { return x.imag() < y.imag(); });
  printf("\n Eigenvalues for N+1 = %s\n\n", I(n + 1, 4).c_str());
  for (int i = 0; i <= n; ++i) {
    const cd c = omega[i] / alpha;
    printf(" %s  %s %s %s %s \n", I(i, 5).c_str(), ES(omega[i].real(), 17, 10).c_str(), ES(omega[i].imag(), 17, 10).c_str(),
           ES(c.real(), 17, 10).c_str(), ES(c.imag(), 17, 10).c_str());
  }
  int which = -1, bad = sst;
  printf("\n Eigenvector for which eigenvalue (-1 quits) ==> ");
  while (std::cin >> which && which != -1) {
    if (which < 0 || which > n) { fprintf(stderr, "index out of range\n"); bad = 1; break; }
    // the eigenvector of the chosen mode: shift-invert with its eigenvalue as the shift
    double sg[2] = {omega[which].real(), omega[which].imag()}, om[2];
    std::vector<double> ev((size_t)(n + 1) * 2), av((size_t)(n + 1) * 2);
    int32_t it, st;
    CK(ctx, osb_chan_eig(ctx, chan, Re, 1, al, sg, om, ev.data(), av.data(), OSB_ADJ_PENCIL, &it, &st));
    bad |= st;
    {
      // the two zdotc(avec, evec) checks (orrcolchan.f:537-546).  The first is taken with avec as LAPACK leaves it
      // (largest |re|+|im| component scaled to modulus 1, phase arbitrary -- here: that component = 1); the second,
      // after the scaling by 1/conj(escale), is 1.
      int k = 0;
      double best = -1.0;
      for (int i = 0; i <= n; ++i) {
        const double m = std::fabs(av[2 * i]) + std::fabs(av[2 * i + 1]);
        if (m > best) { best = m; k = i; }
      }
      const cd e1 = std::conj(cd(1.0, 0.0) / cd(av[2 * k], av[2 * k + 1]));
      cd e2(0.0, 0.0);
      for (int i = 0; i <= n; ++i) e2 += std::conj(cd(av[2 * i], av[2 * i + 1])) * cd(ev[2 * i], ev[2 * i + 1]);
      printf(" (%.16g,%.16g)\n (%.16g,%.16g)\n", e1.real(), e1.imag(), e2.real(), e2.imag());
    }
    FILE *f = fopen("fort.11", "w");  // orrcolchan.f:551-555
    for (int i = 0; i <= n; ++i)
      fprintf(f, " %s %s %s %s %s \n", ES(eta[i], 16, 8, 3).c_str(), ES(ev[2 * i], 16, 8, 3).c_str(), ES(ev[2 * i + 1], 16, 8, 3).c_str(),
              ES(av[2 * i], 16, 8, 3).c_str(), ES(av[2 * i + 1], 16, 8, 3).c_str());
    fclose(f);
    printf("\n Eigenvector for which eigenvalue (-1 quits) ==> ");
  }
  osb_chan_free(ctx, chan);
  return bad;
}

// ---------------------------------------------------------------- orrncbl
static int run_orrncbl(osb_ctx *ctx) {
  int n;
  double Re, sfac, alpha_r;
  printf("\n\n          Find Neutral Curve (Collocation)\n");
  printf("\n Enter the number of modes ==> "); std::cin >> n;
  printf("\n Enter Reynolds number ==> "); std::cin >> Re;
  printf("\n Enter step size ==> "); std::cin >> sfac;
  printf("\n Enter alpha_r ==> "); std::cin >> alpha_r;
  osb_chan *chan = nullptr;
  CK(ctx, osb_chan_create(ctx, OSB_DISC_CHEB_NUM, n, &chan));
  const int maxit = 500;
  std::vector<double> trace((size_t)maxit * 7);
  double aout, om[2];
  int32_t steps, st;
  CK(ctx, osb_chan_neutral(ctx, chan, 1, &Re, &alpha_r, sfac, 1.0e-8, maxit, &aout, om, &steps, &st, trace.data()));
  for (int k = 0; k < steps; ++k) {
    const double *r = &trace[(size_t)k * 7];
    if (k == 0) printf(" Eigenvalue =  (%.16g,%.16g)\n", r[1], r[2]);  // list-directed in the reference (orrncbl.f:902)
    printf(" dw/da = %s %si, dalpha = %s, nw = %s\n", E(r[3], 12, 5).c_str(), E(r[4], 12, 5).c_str(), E(r[5], 12, 5).c_str(),
           E(r[6], 12, 5).c_str());  // orrncbl.f:982,990-991
    printf(" %s %s %s \n", E(r[0], 15, 8).c_str(), E(r[1], 15, 8).c_str(), E(r[2], 15, 8).c_str());  // :114-115
  }
  osb_chan_free(ctx, chan);
  return st;
}

// ---------------------------------------------------------------- self tests (no GPU)
static int selftest_format() {
  printf("E17.8|%s|%s|%s|\n", E(0.36412287e0, 17, 8).c_str(), E(-9.4619873e-5, 17, 8).c_str(), E(0.0, 17, 8).c_str());
  printf("E20.10|%s|%s|\n", E(580.0, 20, 10).c_str(), E(0.179, 20, 10).c_str());
  printf("E17.8E3|%s|\n", E(0.23752649, 17, 8, 3).c_str());
  printf("ES16.8E3|%s|%s|%s|\n", ES(17.0, 16, 8, 3).c_str(), ES(-1.5e-7, 16, 8, 3).c_str(), ES(0.0, 16, 8, 3).c_str());
  printf("1PE20.12E3|%s|\n", ES(0.5, 20, 12, 3).c_str());
  printf("ES17.8|%s|%s|\n", ES(2.88657986e-15, 17, 8).c_str(), ES(1.0e-101, 17, 8).c_str());
  printf("E15.8|%s|%s|\n", E(0.999999996, 15, 8).c_str(), E(1.0e100, 15, 8).c_str());
  printf("I5|%s|%s|\n", I(21, 5).c_str(), I(123456, 5).c_str());
  return 0;
}
static int selftest_deck(const char *path) {
  Deck d;
  if (!d.open(path)) return 1;
  const long long nbl = d.geti(), nstep = d.geti();
  printf("%lld %lld", nbl, nstep);
  for (int k = 0; k < 10; ++k) printf(" %.17g", d.getr());
  printf("\n");
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: osb_drivers contebl|conte|orrsom|orrspace|orrfdchan|orrcolchan|orrncbl|conteucbl|orrfdbl|orrcolbl  (input on stdin)\n");
    return 64;
  }
  const std::string prog = argv[1];
  if (prog == "--selftest-format") return selftest_format();
  if (prog == "--selftest-deck" && argc > 2) return selftest_deck(argv[2]);
  if (prog == "--selftest-readr" && argc > 2) {
    double v = 0;
    std::string fld;
    if (!fortfmt::deck_field(argv[2], fld) || !fortfmt::read_G20_10(fld, v)) return 1;
    printf("%.17g\n", v);
    return 0;
  }
  osb_ctx *ctx = nullptr;
  const int rc = osb_create(&ctx, 0);
  if (rc) {
    fprintf(stderr, "osb_drivers: no usable CUDA device (osb_create = %d); there is no CPU fallback\n", rc);
    return 3;
  }
  int st = 0;
  if (prog == "contebl") st = run_contebl(ctx);
  else if (prog == "conteucbl") st = run_conteucbl(ctx);
  else if (prog == "conte") st = run_conte(ctx);
  else if (prog == "orrsom") st = run_orrsom(ctx);
  else if (prog == "orrspace") st = run_orrspace(ctx);
  else if (prog == "orrfdchan") st = run_orrfdchan(ctx);
  else if (prog == "orrcolchan") st = run_orrcolchan(ctx);
  else if (prog == "orrncbl") st = run_orrncbl(ctx);
  else if (prog == "orrfdbl") st = run_orrbl(ctx, false);
  else if (prog == "orrcolbl") st = run_orrbl(ctx, true);
  else { fprintf(stderr, "unknown program %s\n", prog.c_str()); st = 64; }
  osb_destroy(ctx);
  return st == 0 ? 0 : 1;
}
