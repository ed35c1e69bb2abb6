// fortfmt.hpp -- the small part of Fortran formatted I/O the reference's drivers use,
// so that the C++ twin drivers write byte-compatible stdout / fort.N records.
//
//   Ew.d      0.ddddddddE+ee        (contebl.f:159 e20.10, shoot.f:747 e17.8, orrspace.f:191)
//   Ew.dEe    0.ddddddddE+eee       (shoot.f:331 e17.8e3)
//   ESw.dEe   d.ddddddddE+eee       (shoot.f:808 ES16.8E3, contebl.f:525)
//   1PEw.dEe  d.ddddddddE+eee       (orrsom.f:797 1PE20.12E3, orrcolchan.f:512 1pe17.10)
//   Iw, Fw.d
//   READI / READR: "key = value # comment" with (I10) / (G20.10)  (contebl.f:565-636)
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace fortfmt {

inline std::string pad(const std::string &s, int w) {
  if ((int)s.size() > w) return std::string(w, '*');  // Fortran fills an overflowing field with '*'
  return std::string(w - s.size(), ' ') + s;
}

// significant decimal digits of |x| and its decimal exponent: x = 0.d1d2...dn * 10^exp10
inline void digits(double x, int n, std::string &dig, int &exp10) {
  if (x == 0.0 || !std::isfinite(x)) { dig.assign(n, '0'); exp10 = 0; return; }
  char buf[64];
  snprintf(buf, sizeof buf, "%.*e", n - 1, std::fabs(x));  // d.ddd...e+XX, correctly rounded
  dig.clear();
  const char *p = buf;
  for (; *p && *p != 'e'; ++p) if (*p >= '0' && *p <= '9') dig.push_back(*p);
  exp10 = atoi(p + 1) + 1;
}

inline std::string expo(int e, int edigits, bool allow_drop_E) {
  char buf[16];
  const int ae = std::abs(e);
  if (edigits == 0) {  // default exponent width: E+ee, or +eee without the letter when |e| > 99
    if (ae <= 99) snprintf(buf, sizeof buf, "E%c%02d", e < 0 ? '-' : '+', ae);
    else if (allow_drop_E) snprintf(buf, sizeof buf, "%c%03d", e < 0 ? '-' : '+', ae);
    else snprintf(buf, sizeof buf, "E%c%d", e < 0 ? '-' : '+', ae);
  } else {
    snprintf(buf, sizeof buf, "E%c%0*d", e < 0 ? '-' : '+', edigits, ae);
  }
  return buf;
}

// Ew.d[Ee]
inline std::string E(double x, int w, int d, int e = 0) {
  if (std::isnan(x)) return pad("NaN", w);
  if (std::isinf(x)) return pad(x < 0 ? "-Infinity" : "Infinity", w);
  std::string dig;
  int ex;
  digits(x, d, dig, ex);
  std::string s = (std::signbit(x) && x != 0.0 ? "-0." : "0.") + dig + expo(x == 0.0 ? 0 : ex, e, true);
  if ((int)s.size() > w && s.compare(0, 2, "0.") == 0) s = s.substr(1);        // drop the optional leading zero
  if ((int)s.size() > w && s.compare(0, 3, "-0.") == 0) s = "-" + s.substr(2);
  return pad(s, w);
}

// ESw.d[Ee] and 1PEw.d[Ee]: one non-zero digit before the point, d after
inline std::string ES(double x, int w, int d, int e = 0) {
  if (std::isnan(x)) return pad("NaN", w);
  if (std::isinf(x)) return pad(x < 0 ? "-Infinity" : "Infinity", w);
  std::string dig;
  int ex;
  digits(x, d + 1, dig, ex);
  std::string s = std::string(std::signbit(x) && x != 0.0 ? "-" : "") + dig.substr(0, 1) + "." + dig.substr(1) +
                  expo(x == 0.0 ? 0 : ex - 1, e, true);
  return pad(s, w);
}

inline std::string I(long long v, int w) { return pad(std::to_string(v), w); }

inline std::string F(double x, int w, int d) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.*f", d, x);
  return pad(buf, w);
}

// ---- deck reading ----------------------------------------------------------
// text between '=' and '#' (or column 80), as READI/READR slice it
inline bool deck_field(const std::string &line80, std::string &field) {
  std::string s = line80;
  s.resize(80, ' ');
  const size_t iloc = s.find('=');
  if (iloc == std::string::npos) return false;
  size_t floc = s.find('#');
  if (floc == std::string::npos) floc = 80;  // "floc = 80"
  field = s.substr(iloc + 1, floc > iloc + 1 ? floc - (iloc + 1) : 0);
  return true;
}

// (I10): the first 10 characters, blanks ignored
inline bool read_I10(const std::string &field, long long &v) {
  std::string t;
  for (size_t i = 0; i < field.size() && i < 10; ++i) if (field[i] != ' ') t.push_back(field[i]);
  if (t.empty()) { v = 0; return true; }
  char *end = nullptr;
  v = strtoll(t.c_str(), &end, 10);
  return end && *end == 0;
}

// (G20.10): the first 20 characters; WITHOUT a decimal point the digits are scaled by
// 10**-10 (the "d" of the edit descriptor) -- the reference's decks all carry a point
inline bool read_G20_10(const std::string &field, double &v) {
  std::string t;
  for (size_t i = 0; i < field.size() && i < 20; ++i) if (field[i] != ' ') t.push_back(field[i]);
  if (t.empty()) { v = 0.0; return true; }
  for (auto &ch : t) if (ch == 'd' || ch == 'D') ch = 'E';
  const bool has_point = t.find('.') != std::string::npos;
  // Fortran also accepts an exponent without the letter: 1.5+3
  for (size_t i = 1; i < t.size(); ++i)
    if ((t[i] == '+' || t[i] == '-') && t[i - 1] != 'E' && t[i - 1] != 'e') { t.insert(i, "E"); break; }
  char *end = nullptr;
  double x = strtod(t.c_str(), &end);
  if (!end || *end != 0) return false;
  if (!has_point) {
    // digits before any exponent are an integer to be scaled by 1e-10
    x *= 1.0e-10;
  }
  v = x;
  return true;
}

}  // namespace fortfmt
