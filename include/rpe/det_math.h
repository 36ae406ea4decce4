// rpe/det_math.h — bit-reproducible elementary functions for host and device.
//
// Why this exists: the reference calls libm / libstdc++ transcendentals inside its
// minimal solvers and its adaptive stopping rule:
//   * acos, AngleAxis->Quaternion (cos/sin of half angle)   AbsoluteOrientationNormal.hpp:89,94,103,107,119,122
//   * std::complex pow / sqrt in the Ferrari quartic         P3P.hpp:34-57
//   * std::log / std::pow in RANSACUpdateNumIters            P3P.hpp:305-317
// glibc and the CUDA math library disagree in the last ulp for these, so a GPU
// generator could never be compared bit-for-bit with a CPU oracle that calls libm.
// Every function below is built from + - * / sqrt on IEEE binary64 only (each op
// individually rounded: explicit __d*_rn intrinsics on the device, plain operators on the
// host where x86-64 baseline has no FMA), so the SAME source yields the SAME bits on
// both sides. Results are rounded to float by the *_f wrappers; the double kernels are
// accurate to ~1e-15, i.e. the float results are correctly rounded except in
// astronomically rare double-rounding cases, hence equal to a correctly rounded libm.
//
// Header-only, no dependencies beyond <stdint.h>/<string.h>.
#ifndef RPE_DET_MATH_H_
#define RPE_DET_MATH_H_

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RPE_HD __host__ __device__ __forceinline__
#else
#define RPE_HD inline
#endif

namespace rpe {
namespace det {

// ---- individually rounded binary64 primitives -------------------------------------
RPE_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
RPE_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
RPE_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
RPE_HD double ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
RPE_HD double dsqrt(double a) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(a);
#else
  return __builtin_sqrt(a);
#endif
}
RPE_HD uint64_t dbits(double d) {
  uint64_t u;
  memcpy(&u, &d, sizeof(u));
  return u;
}
RPE_HD double dfrombits(uint64_t u) {
  double d;
  memcpy(&d, &u, sizeof(d));
  return d;
}
RPE_HD double dabs(double a) { return dfrombits(dbits(a) & 0x7fffffffffffffffULL); }
RPE_HD bool disnan(double a) { return a != a; }
RPE_HD double dnan() { return dfrombits(0x7ff8000000000000ULL); }
RPE_HD double dinf() { return dfrombits(0x7ff0000000000000ULL); }

// ---- constants ---------------------------------------------------------------------
#define RPE_DET_PI 3.14159265358979323846
#define RPE_DET_PIO2 1.57079632679489661923
#define RPE_DET_PIO2_HI 1.57079632673412561417e+00 /* 0x3FF921FB54400000 */
#define RPE_DET_PIO2_LO 6.07710050650619224932e-11 /* 0x3DD0B4611A626331 */
#define RPE_DET_LN2_HI 6.93147180369123816490e-01  /* 0x3FE62E42FEE00000 */
#define RPE_DET_LN2_LO 1.90821492927058770002e-10  /* 0x3DEA39EF35793C76 */
#define RPE_DET_SQRT2 1.41421356237309504880

// ---- natural logarithm ---------------------------------------------------------------
// x = m * 2^e, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716;
// atanh series to s^27 (truncation < 1e-20).
RPE_HD double log_d(double x) {
  if (disnan(x) || x < 0.0) return dnan();
  if (x == 0.0) return -dinf();
  if (x == dinf()) return x;
  uint64_t u = dbits(x);
  int e = (int)((u >> 52) & 0x7ff);
  if (e == 0) {  // subnormal: renormalise exactly
    x = dmul(x, 18014398509481984.0);  // 2^54
    u = dbits(x);
    e = (int)((u >> 52) & 0x7ff) - 54;
  }
  e -= 1023;
  double m = dfrombits((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
  if (m > RPE_DET_SQRT2) {
    m = dmul(m, 0.5);
    e += 1;
  }
  const double s = ddiv(dsub(m, 1.0), dadd(m, 1.0));
  const double z = dmul(s, s);
  double p = 1.0 / 27.0;
  p = dadd(dmul(p, z), 1.0 / 25.0);
  p = dadd(dmul(p, z), 1.0 / 23.0);
  p = dadd(dmul(p, z), 1.0 / 21.0);
  p = dadd(dmul(p, z), 1.0 / 19.0);
  p = dadd(dmul(p, z), 1.0 / 17.0);
  p = dadd(dmul(p, z), 1.0 / 15.0);
  p = dadd(dmul(p, z), 1.0 / 13.0);
  p = dadd(dmul(p, z), 1.0 / 11.0);
  p = dadd(dmul(p, z), 1.0 / 9.0);
  p = dadd(dmul(p, z), 1.0 / 7.0);
  p = dadd(dmul(p, z), 1.0 / 5.0);
  p = dadd(dmul(p, z), 1.0 / 3.0);
  p = dadd(dmul(p, z), 1.0);
  const double lm = dmul(dmul(2.0, s), p);
  const double de = (double)e;
  return dadd(dmul(de, RPE_DET_LN2_HI), dadd(lm, dmul(de, RPE_DET_LN2_LO)));
}

// ---- arctangent ------------------------------------------------------------------------
// |x|>1 -> pi/2 - atan(1/x); two half-angle reductions a <- a/(1+sqrt(1+a^2)) bring the
// argument below tan(pi/16) = 0.19892; alternating series to a^27 (truncation < 1e-20).
RPE_HD double atan_d(double x) {
  if (disnan(x)) return x;
  const bool neg = x < 0.0;
  double a = dabs(x);
  const bool inv = a > 1.0;
  if (inv) a = ddiv(1.0, a);  // a = 0 for x = inf
  a = ddiv(a, dadd(1.0, dsqrt(dadd(1.0, dmul(a, a)))));
  a = ddiv(a, dadd(1.0, dsqrt(dadd(1.0, dmul(a, a)))));
  const double z = dmul(a, a);
  double p = 1.0 / 27.0;
  p = dsub(1.0 / 25.0, dmul(p, z));
  p = dsub(1.0 / 23.0, dmul(p, z));
  p = dsub(1.0 / 21.0, dmul(p, z));
  p = dsub(1.0 / 19.0, dmul(p, z));
  p = dsub(1.0 / 17.0, dmul(p, z));
  p = dsub(1.0 / 15.0, dmul(p, z));
  p = dsub(1.0 / 13.0, dmul(p, z));
  p = dsub(1.0 / 11.0, dmul(p, z));
  p = dsub(1.0 / 9.0, dmul(p, z));
  p = dsub(1.0 / 7.0, dmul(p, z));
  p = dsub(1.0 / 5.0, dmul(p, z));
  p = dsub(1.0 / 3.0, dmul(p, z));
  p = dsub(1.0, dmul(p, z));
  double r = dmul(4.0, dmul(a, p));
  if (inv) r = dsub(RPE_DET_PIO2, r);
  return neg ? -r : r;
}

RPE_HD double atan2_d(double y, double x) {
  if (disnan(x) || disnan(y)) return dnan();
  if (x > 0.0) return atan_d(ddiv(y, x));
  if (x < 0.0) {
    const double r = atan_d(ddiv(y, x));
    // y/x <= 0 when y >= 0, so r <= 0: add pi; y < 0: subtract pi. (-0.0 counts as y >= 0 here.)
    return (y < 0.0) ? dsub(r, RPE_DET_PI) : dadd(r, RPE_DET_PI);
  }
  if (y > 0.0) return RPE_DET_PIO2;
  if (y < 0.0) return -RPE_DET_PIO2;
  return 0.0;
}

// acos(x) = 2 atan2(sqrt(1-x), sqrt(1+x)); NaN outside [-1,1] like libm.
RPE_HD double acos_d(double x) {
  if (disnan(x) || x > 1.0 || x < -1.0) return dnan();
  return dmul(2.0, atan2_d(dsqrt(dsub(1.0, x)), dsqrt(dadd(1.0, x))));
}

// ---- sine / cosine ---------------------------------------------------------------------
// Cody-Waite reduction by pi/2 (two constants, good for |a| < ~1e5), Taylor on |r| <= pi/4.
RPE_HD void sincos_d(double a, double* sn, double* cs) {
  if (disnan(a) || dabs(a) == dinf()) {
    *sn = dnan();
    *cs = dnan();
    return;
  }
  const double kf = dmul(a, 0.63661977236758134308);  // 2/pi
  const long long k = (long long)(kf < 0.0 ? dsub(kf, 0.5) : dadd(kf, 0.5));
  const double kd = (double)k;
  const double r = dsub(dsub(a, dmul(kd, RPE_DET_PIO2_HI)), dmul(kd, RPE_DET_PIO2_LO));
  const double z = dmul(r, r);
  // sin r = r * (1 - z/3! + z^2/5! - ... - z^9/19!)
  double ps = -1.0 / 121645100408832000.0;           // 1/19!
  ps = dadd(dmul(ps, z), 1.0 / 355687428096000.0);   // 1/17!
  ps = dadd(dmul(ps, z), -1.0 / 1307674368000.0);    // 1/15!
  ps = dadd(dmul(ps, z), 1.0 / 6227020800.0);        // 1/13!
  ps = dadd(dmul(ps, z), -1.0 / 39916800.0);         // 1/11!
  ps = dadd(dmul(ps, z), 1.0 / 362880.0);            // 1/9!
  ps = dadd(dmul(ps, z), -1.0 / 5040.0);             // 1/7!
  ps = dadd(dmul(ps, z), 1.0 / 120.0);               // 1/5!
  ps = dadd(dmul(ps, z), -1.0 / 6.0);                // 1/3!
  ps = dadd(dmul(ps, z), 1.0);
  const double s = dmul(r, ps);
  // cos r = 1 - z/2! + z^2/4! - ... + z^10/20!
  double pc = 1.0 / 2432902008176640000.0;            // 1/20!
  pc = dadd(dmul(pc, z), -1.0 / 6402373705728000.0);  // 1/18!
  pc = dadd(dmul(pc, z), 1.0 / 20922789888000.0);     // 1/16!
  pc = dadd(dmul(pc, z), -1.0 / 87178291200.0);       // 1/14!
  pc = dadd(dmul(pc, z), 1.0 / 479001600.0);          // 1/12!
  pc = dadd(dmul(pc, z), -1.0 / 3628800.0);           // 1/10!
  pc = dadd(dmul(pc, z), 1.0 / 40320.0);              // 1/8!
  pc = dadd(dmul(pc, z), -1.0 / 720.0);               // 1/6!
  pc = dadd(dmul(pc, z), 1.0 / 24.0);                 // 1/4!
  pc = dadd(dmul(pc, z), -0.5);                       // 1/2!
  const double c = dadd(dmul(pc, z), 1.0);
  switch ((int)(k & 3)) {
    case 0: *sn = s;  *cs = c;  break;
    case 1: *sn = c;  *cs = -s; break;
    case 2: *sn = -s; *cs = -c; break;
    default: *sn = -c; *cs = s; break;
  }
}

// ---- real cube root ----------------------------------------------------------------------
// x = m * 8^q with m in [1,8); linear seed + 6 Newton steps y <- y - (y^3-m)/(3y^2).
RPE_HD double cbrt_d(double x) {
  if (disnan(x) || x == 0.0 || dabs(x) == dinf()) return x;
  const bool neg = x < 0.0;
  double a = dabs(x);
  uint64_t u = dbits(a);
  int e = (int)((u >> 52) & 0x7ff);
  if (e == 0) {
    a = dmul(a, 18014398509481984.0);  // 2^54
    u = dbits(a);
    e = (int)((u >> 52) & 0x7ff) - 54;
  }
  e -= 1023;
  // floor division of e by 3
  int q = e / 3;
  int rem = e - 3 * q;
  if (rem < 0) {
    rem += 3;
    q -= 1;
  }
  double m = dfrombits((u & 0x000fffffffffffffULL) | ((uint64_t)(1023 + rem) << 52));  // [1,8)
  double y = dadd(0.8, dmul(0.15, m));                                                  // crude seed in [0.95,2]
  for (int it = 0; it < 7; ++it) {
    const double y2 = dmul(y, y);
    y = dsub(y, ddiv(dsub(dmul(y2, y), m), dmul(3.0, y2)));
  }
  const double scale = dfrombits((uint64_t)(1023 + q) << 52);
  const double r = dmul(y, scale);
  return neg ? -r : r;
}

// ---- float front-ends (compute in binary64, round once) -----------------------------------
RPE_HD float log_f(float x) { return (float)log_d((double)x); }
RPE_HD float acos_f(float x) { return (float)acos_d((double)x); }
RPE_HD float atan2_f(float y, float x) { return (float)atan2_d((double)y, (double)x); }
RPE_HD float cbrt_f(float x) { return (float)cbrt_d((double)x); }
RPE_HD void sincos_f(float a, float* s, float* c) {
  double sd, cd;
  sincos_d((double)a, &sd, &cd);
  *s = (float)sd;
  *c = (float)cd;
}

// Overloads so templated code can call rpe::det::log_t<Tp>() etc.
RPE_HD float log_t(float x) { return log_f(x); }
RPE_HD double log_t(double x) { return log_d(x); }
RPE_HD float acos_t(float x) { return acos_f(x); }
RPE_HD double acos_t(double x) { return acos_d(x); }
RPE_HD float atan2_t(float y, float x) { return atan2_f(y, x); }
RPE_HD double atan2_t(double y, double x) { return atan2_d(y, x); }
RPE_HD float cbrt_t(float x) { return cbrt_f(x); }
RPE_HD double cbrt_t(double x) { return cbrt_d(x); }
RPE_HD void sincos_t(float a, float* s, float* c) { sincos_f(a, s, c); }
RPE_HD void sincos_t(double a, double* s, double* c) { sincos_d(a, s, c); }

}  // namespace det
}  // namespace rpe

#endif  // RPE_DET_MATH_H_
