// rpe/ransac_rule.h — the reference's adaptive stopping rule, host + device, bit-reproducible.
//
// Restates RANSACUpdateNumIters<float> (/root/reference/pose/P3P.hpp:296-318) including the mixed
// precision its C++ typing produces ("1. - p" and std::pow(float,int) are evaluated in double; the
// final rounding adds a float literal 0.5f), with std::log replaced by rpe::det::log_f so that the
// host adapters, the device replay kernel and the DET-mode oracle agree bit for bit.
// Also the outlier ratio expressions handed to it:
//   (Tp)(N - votes) / N                AbsoluteOrientation.hpp:150,207,265  P3P.hpp:383,460
//   (Tp)(N*2 - votes) / N / 2          AbsoluteOrientation.hpp:429,506  AbsoluteOrientationNormal.hpp:276,346
//   (Tp)(N*3 - votes) / N / 3          AbsoluteOrientationNormal.hpp:435
#ifndef RPE_RANSAC_RULE_H_
#define RPE_RANSAC_RULE_H_

#include "det_math.h"

namespace rpe {

RPE_HD float rule_fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
RPE_HD float rule_fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
RPE_HD float rule_fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}

// The rule split into its three data dependencies, so that a device replay can evaluate the expensive
// part (the logarithm of the denominator, a function of the vote count only) for many candidates in
// parallel and keep only the cheap clamp/divide sequential. update_num_iters() composes them.
struct RuleDenominator {
  int state;        // 0: "1 - (1-ep)^K < eps" -> the rule returns 0;  1: log_denom valid
  float log_denom;  // log(1 - (1-ep)^K)
};

RPE_HD float rule_log_numerator(float p) {
  const float feps = 1.1920928955078125e-07f;  // std::numeric_limits<float>::epsilon()
  p = p > 0.f ? p : 0.f;
  p = p < 1.f ? p : 1.f;
  float num = (float)det::dsub(1.0, (double)p);
  num = num > feps ? num : feps;
  return det::log_f(num);
}

RPE_HD RuleDenominator rule_denominator(float ep, const int model_points) {
  const float feps = 1.1920928955078125e-07f;
  ep = ep > 0.f ? ep : 0.f;
  ep = ep < 1.f ? ep : 1.f;
  const double base = (double)((float)det::dsub(1.0, (double)ep));
  double pw = 1.0;
  for (int i = 0; i < model_points; ++i) pw = det::dmul(pw, base);
  const float denom = (float)det::dsub(1.0, pw);
  RuleDenominator r;
  if (denom < feps) {
    r.state = 0;
    r.log_denom = 0.f;
  } else {
    r.state = 1;
    r.log_denom = det::log_f(denom);
  }
  return r;
}

RPE_HD int rule_finish(float log_num, RuleDenominator d, const int max_iters) {
  if (d.state == 0) return 0;
  const float num = log_num, denom = d.log_denom;
  if (denom >= 0.f || -num >= rule_fmul((float)max_iters, -denom)) return max_iters;
  return (int)rule_fadd(rule_fdiv(num, denom), 0.5f);
}

RPE_HD int update_num_iters(float p, float ep, const int model_points, const int max_iters) {
  return rule_finish(rule_log_numerator(p), rule_denominator(ep, model_points), max_iters);
}

// ---- Tp = double: RANSACUpdateNumIters<double> (every operation in binary64; `+ 0.5f` promotes to double) --------
RPE_HD double rule_ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
struct RuleDenominatorD {
  int state;
  double log_denom;
};
RPE_HD double rule_log_numerator_d(double p) {
  const double deps = 2.220446049250313e-16;  // std::numeric_limits<double>::epsilon()
  p = p > 0.0 ? p : 0.0;
  p = p < 1.0 ? p : 1.0;
  double num = det::dsub(1.0, p);
  num = num > deps ? num : deps;
  return det::log_d(num);
}
RPE_HD RuleDenominatorD rule_denominator_d(double ep, const int model_points) {
  const double deps = 2.220446049250313e-16;
  ep = ep > 0.0 ? ep : 0.0;
  ep = ep < 1.0 ? ep : 1.0;
  const double base = det::dsub(1.0, ep);
  double pw = 1.0;
  for (int i = 0; i < model_points; ++i) pw = det::dmul(pw, base);
  const double denom = det::dsub(1.0, pw);
  RuleDenominatorD r;
  if (denom < deps) {
    r.state = 0;
    r.log_denom = 0.0;
  } else {
    r.state = 1;
    r.log_denom = det::log_d(denom);
  }
  return r;
}
RPE_HD int rule_finish_d(double log_num, RuleDenominatorD d, const int max_iters) {
  if (d.state == 0) return 0;
  const double num = log_num, denom = d.log_denom;
  if (denom >= 0.0 || -num >= det::dmul((double)max_iters, -denom)) return max_iters;
  return (int)det::dadd(rule_ddiv(num, denom), 0.5);
}
RPE_HD int update_num_iters_d(double p, double ep, const int model_points, const int max_iters) {
  return rule_finish_d(rule_log_numerator_d(p), rule_denominator_d(ep, model_points), max_iters);
}
RPE_HD double outlier_ratio_d(int modalities, int n, int votes) {
  if (modalities == 1) return rule_ddiv((double)(n - votes), (double)n);
  return rule_ddiv(rule_ddiv((double)(n * modalities - votes), (double)n), (double)modalities);
}

RPE_HD float outlier_ratio(int modalities, int n, int votes) {
  if (modalities == 1) return rule_fdiv((float)(n - votes), (float)n);
  return rule_fdiv(rule_fdiv((float)(n * modalities - votes), (float)n), (float)modalities);
}

}  // namespace rpe

#endif  // RPE_RANSAC_RULE_H_
