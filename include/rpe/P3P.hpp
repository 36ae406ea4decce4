// rpe/P3P.hpp — mirrors /root/reference/pose/P3P.hpp.
//
//   o4_roots                 :11-60     Ferrari quartic (host; same template as the device generator)
//   kneip_main / kneip       :63-294    Kneip P3P, up to 4 poses, 4th-point disambiguation (host)
//   RANSACUpdateNumIters     :296-318   (rpe/Estimators.hpp)
//   kneip_ransac / _prosac   :320-469   -> rpe_ransac(RPE_KNEIP / RPE_KNEIP_QUAT)
//   lsq_pnp                  :472-502   sums and prints the reprojection error (host)
#ifndef RPE_P3P_HPP_
#define RPE_P3P_HPP_

#include <limits>
#include <iostream>
#include <vector>

#include "Estimators.hpp"
#include "PnPPoseAdapter.hpp"
#include "solvers_p3p.h"

template <typename Tp, class M>
std::vector<Tp> o4_roots(const M& p_) {
  Tp f[5], r[4];
  for (int i = 0; i < 5; ++i) f[i] = p_(i, 0);
  rpe::o4_roots_dev<Tp>(f, r);
  return std::vector<Tp>(r, r + 4);
}

// X_w, bv: 3 x (>=3) column-major
template <typename Tp, class M>
void kneip_main(const M& X_w, const M& bv, std::vector<rpe::SE3<Tp> >* p_solutions_) {
  p_solutions_->clear();
  Tp qs[4][4], ts[4][3];
  const int ns = rpe::kneip_main_dev<Tp>(X_w.data(), bv.data(), qs, ts);
  for (int i = 0; i < ns; ++i)
    p_solutions_->push_back(rpe::SE3<Tp>(rpe::SO3<Tp>::fromRawQuaternion(qs[i]), rpe::Vec3<Tp>(ts[i][0], ts[i][1], ts[i][2])));
}

template <typename Tp>
std::vector<rpe::SE3<Tp> > kneip(PnPPoseAdapter<Tp>& adapter, int i0 = 0, int i1 = 1, int i2 = 2) {
  rpe::MatrixX<Tp> bv(3, 3), X_w(3, 3);
  const int idx[3] = {i0, i1, i2};
  for (int k = 0; k < 3; ++k) {
    bv.setCol(k, adapter.getBearingVector(idx[k]));
    X_w.setCol(k, adapter.getPointGlob(idx[k]));
  }
  std::vector<rpe::SE3<Tp> > solutions;
  kneip_main<Tp>(X_w, bv, &solutions);
  return solutions;
}

// 3 x 4 inputs; the 4th column disambiguates [reference :250-294]
template <typename Tp, class M>
bool kneip(const M& X_w_, const M& bv_, rpe::SE3<Tp>* p_sol_) {
  Tp q[4], t[3];
  if (!rpe::kneip_select<Tp>(X_w_.data(), bv_.data(), std::numeric_limits<Tp>::max(), q, t)) return false;
  *p_sol_ = rpe::SE3<Tp>(rpe::SO3<Tp>::fromRawQuaternion(q), rpe::Vec3<Tp>(t[0], t[1], t[2]));
  return true;
}

template <typename Tp>
void kneip_ransac(PnPPoseAdapter<Tp>& adapter, const Tp thre_2d_, int& Iter, Tp confidence = 0.99) {
  const Tp cos_thr = std::cos(std::atan(thre_2d_ / adapter.getFocal()));  // [reference :323]
  rpe::detail::RansacRows rows(adapter.getNumberCorrespondences(), 4);
  rpe::detail::run_ransac<Tp>(adapter, RPE_KNEIP, rows, nullptr, Tp(0), cos_thr, Tp(0), Iter, confidence);
  adapter.cvtInlier();  // [reference :389]
}

template <typename Tp>
void kneip_prosac(PnPPoseAdapter<Tp>& adapter, const Tp thre_2d_, int& Iter, Tp confidence = 0.99) {
  const Tp cos_thr = std::cos(std::atan(thre_2d_ / adapter.getFocal()));  // [reference :398]
  rpe::detail::ProsacRows<Tp, PnPPoseAdapter<Tp> > rows(adapter, 4);
  rpe::detail::run_ransac<Tp>(adapter, RPE_KNEIP_QUAT, rows, nullptr, Tp(0), cos_thr, Tp(0), Iter, confidence);
  adapter.cvtInlier();  // [reference :466]
}

template <typename Tp>
void lsq_pnp(PnPPoseAdapter<Tp>& adapter) {  // [reference :472-502]
  Tp total_err = 0.;
  for (int i = 0; i < adapter.getNumberCorrespondences(); i++) total_err += adapter.getError(i);
  std::cout << total_err << std::endl;
}

#endif  // RPE_P3P_HPP_
