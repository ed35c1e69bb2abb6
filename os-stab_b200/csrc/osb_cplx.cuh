// osb_cplx.cuh -- fp64 complex helpers for the device code.
//
// The reference is Fortran `complex` promoted to complex*16 (gcc.mak:78); gfortran
// expands * as the plain 4-multiply product and / as Smith's algorithm; ABS, SQRT
// and EXP of a complex go to libm (cabs, csqrt, cexp).  The helpers below follow
// the same algorithms so that the per-pass scalar work (initial condition,
// Gram-Schmidt, secant/Muller update) differs from the reference only by
// rounding (FMA contraction, CUDA libm's 1-2 ulp transcendentals).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace osb {

struct zd {
  double x, y;
};

__host__ __device__ __forceinline__ zd zmk(double x, double y) { return zd{x, y}; }
__host__ __device__ __forceinline__ zd zadd(zd a, zd b) { return zd{a.x + b.x, a.y + b.y}; }
__host__ __device__ __forceinline__ zd zsub(zd a, zd b) { return zd{a.x - b.x, a.y - b.y}; }
__host__ __device__ __forceinline__ zd zneg(zd a) { return zd{-a.x, -a.y}; }
__host__ __device__ __forceinline__ zd zconj(zd a) { return zd{a.x, -a.y}; }
__host__ __device__ __forceinline__ zd zscale(double s, zd a) { return zd{s * a.x, s * a.y}; }
__host__ __device__ __forceinline__ zd zmul(zd a, zd b) {
  return zd{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
// Smith's division as GCC's expand_complex_div_wide (what gfortran emits).
__host__ __device__ __forceinline__ zd zdiv(zd a, zd b) {
  if (fabs(b.x) < fabs(b.y)) {
    double ratio = b.x / b.y;
    double div = b.x * ratio + b.y;
    return zd{(a.x * ratio + a.y) / div, (a.y * ratio - a.x) / div};
  } else {
    double ratio = b.y / b.x;
    double div = b.y * ratio + b.x;
    return zd{(a.y * ratio + a.x) / div, (a.y - a.x * ratio) / div};
  }
}
__host__ __device__ __forceinline__ double zabs(zd a) { return hypot(a.x, a.y); }
// principal square root, the classic stable form used by libm's csqrt
__host__ __device__ __forceinline__ zd zsqrt(zd a) {
  if (a.x == 0.0 && a.y == 0.0) return zd{0.0, a.y};
  double m = hypot(a.x, a.y);
  if (a.x >= 0.0) {
    double t = sqrt(0.5 * (m + a.x));
    return zd{t, a.y / (2.0 * t)};
  } else {
    double t = sqrt(0.5 * (m - a.x));
    return zd{fabs(a.y) / (2.0 * t), copysign(t, a.y)};
  }
}
__device__ __forceinline__ zd zexp(zd a) {
  double e = exp(a.x), s, c;
  sincos(a.y, &s, &c);
  return zd{e * c, e * s};
}
__host__ __device__ __forceinline__ bool zfinite(zd a) { return isfinite(a.x) && isfinite(a.y); }

}  // namespace osb
