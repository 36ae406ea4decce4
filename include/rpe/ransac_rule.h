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

RPE_HD int update_num_iters(float p, float ep, const int model_points, const int max_iters) {
  const float feps = 1.1920928955078125e-07f;  // std::numeric_limits<float>::epsilon()
  p = p > 0.f ? p : 0.f;
  p = p < 1.f ? p : 1.f;
  ep = ep > 0.f ? ep : 0.f;
  ep = ep < 1.f ? ep : 1.f;
  float num = (float)det::dsub(1.0, (double)p);
  num = num > feps ? num : feps;
  const double base = (double)((float)det::dsub(1.0, (double)ep));
  double pw = 1.0;
  for (int i = 0; i < model_points; ++i) pw = det::dmul(pw, base);
  float denom = (float)det::dsub(1.0, pw);
  if (denom < feps) return 0;
  num = det::log_f(num);
  denom = det::log_f(denom);
  if (denom >= 0.f || -num >= rule_fmul((float)max_iters, -denom)) return max_iters;
  return (int)rule_fadd(rule_fdiv(num, denom), 0.5f);
}

RPE_HD float outlier_ratio(int modalities, int n, int votes) {
  if (modalities == 1) return rule_fdiv((float)(n - votes), (float)n);
  return rule_fdiv(rule_fdiv((float)(n * modalities - votes), (float)n), (float)modalities);
}

}  // namespace rpe

#endif  // RPE_RANSAC_RULE_H_
