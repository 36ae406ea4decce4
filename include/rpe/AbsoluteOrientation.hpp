// rpe/AbsoluteOrientation.hpp — mirrors /root/reference/pose/AbsoluteOrientation.hpp.
//
//   calc_percentage_err / calc_err        :11-43    error metrics (host)
//   shinji                                :47-99    closed-form absolute orientation (host; the same solver
//                                                   template the device generator instantiates)
//   shinji_ransac / _ransac2 / _prosac    :101-271  -> rpe_ransac(RPE_SHINJI)
//   shinji_ls / _ls1 / _ls2               :273-342  -> rpe_refit(KABSCH_INLIERS / KABSCH_ALL)
//   shinji_kneip_ransac / _prosac         :367-515  -> rpe_ransac(RPE_SHINJI_KNEIP)
// Same function names, argument order and in/out `Iter` semantics.
#ifndef RPE_ABSOLUTE_ORIENTATION_HPP_
#define RPE_ABSOLUTE_ORIENTATION_HPP_

#include "AOOnlyPoseAdapter.hpp"
#include "AOPoseAdapter.hpp"
#include "Estimators.hpp"
#include "P3P.hpp"

template <typename Tp>
rpe::Vec3<Tp> calc_percentage_err_impl(const rpe::SO3<Tp>& R_cw_, const rpe::Vec3<Tp>& t_w_, const rpe::SO3<Tp>& R_est,
                                       const rpe::Vec3<Tp>& t_est) {
  const rpe::Vec3<Tp> te = R_cw_ * t_w_ - R_est * t_est;  // [reference :13]
  const Tp t_e = te.norm() / t_est.norm() * 100;
  const rpe::Quaternion<Tp> a = R_cw_.unit_quaternion(), b = R_est.unit_quaternion();
  const Tp d[4] = {a.w() - b.w(), a.x() - b.x(), a.y() - b.y(), a.z() - b.z()};
  const Tp r_e = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]) / b.norm() * 100;
  return rpe::Vec3<Tp>(t_e, r_e, Tp(0));
}
// returns (translation error %, rotation error %) as the first two entries
template <typename Tp>
rpe::Vec3<Tp> calc_percentage_err(const rpe::SO3<Tp>& R_cw_, const rpe::Vec3<Tp>& t_w_, const PoseAdapterBase<Tp>* p_ad) {
  return calc_percentage_err_impl<Tp>(R_cw_, t_w_, p_ad->getRcw(), p_ad->gettw());
}
// (|translation|, rotation angle) of a relative transform given as rotation + translation [reference :29-37]
template <typename Tp>
rpe::Vec3<Tp> calc_err(const rpe::Mat3<Tp>& R_diff, const rpe::Vec3<Tp>& t_diff) {
  Tp q[4];
  rpe::so3_from_matrix<Tp>(R_diff.m, q);
  const Tp n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  const Tp angle = Tp(2) * std::atan2(n, std::fabs(q[3]));  // Eigen AngleAxis(Matrix3)
  return rpe::Vec3<Tp>(t_diff.norm(), angle, Tp(0));
}

// shinji: X_c = R_cw * X_w + t_w, least squares over the first K columns [reference :47-99]
template <typename Tp, class M>
rpe::SE3<Tp> shinji(const M& X_w_, const M& X_c_, int K) {
  assert(3 <= K && K <= (int)X_w_.cols() && X_w_.cols() == X_c_.cols() && X_w_.rows() == 3);
  const Tp* xw = X_w_.data();
  const Tp* xc = X_c_.data();
  Tp Cw[3] = {0, 0, 0}, Cc[3] = {0, 0, 0};
  for (int n = 0; n < K; ++n)
    for (int r = 0; r < 3; ++r) {
      Cw[r] = Cw[r] + xw[3 * n + r];
      Cc[r] = Cc[r] + xc[3 * n + r];
    }
  for (int r = 0; r < 3; ++r) {
    Cw[r] = Cw[r] / (Tp)K;
    Cc[r] = Cc[r] / (Tp)K;
  }
  Tp Mm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int n = 0; n < K; ++n) {
    Tp Aw[3], Ac[3];
    for (int r = 0; r < 3; ++r) {
      Aw[r] = xw[3 * n + r] - Cw[r];
      Ac[r] = xc[3 * n + r] - Cc[r];
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Mm[3 * i + j] = Mm[3 * i + j] + Ac[i] * Aw[j];
  }
  for (int i = 0; i < 9; ++i) Mm[i] = Mm[i] / (Tp)X_w_.cols();
  Tp q[4];
  rpe::rotation_from_covariance<Tp>(Mm, q);
  const rpe::SO3<Tp> R = rpe::SO3<Tp>::fromRawQuaternion(q);
  const rpe::Vec3<Tp> t = rpe::Vec3<Tp>(Cc[0], Cc[1], Cc[2]) - R * rpe::Vec3<Tp>(Cw[0], Cw[1], Cw[2]);
  return rpe::SE3<Tp>(R, t);
}

template <typename Tp>
void shinji_ransac(AOPoseAdapter<Tp>& adapter, const Tp dist_thre_3d_, int& Iter, Tp confidence = 0.99) {
  rpe::detail::RansacRows rows(adapter.getNumberCorrespondences(), 3);
  rpe::detail::run_ransac<Tp>(adapter, RPE_SHINJI, rows, nullptr, dist_thre_3d_, Tp(0), Tp(0), Iter, confidence);
  adapter.cvtInlier();  // [reference :153]
}

template <typename Tp>
void shinji_ransac2(AOOnlyPoseAdapter<Tp>& adapter, const Tp dist_thre_3d_, int& Iter, Tp confidence = 0.99) {
  rpe::detail::RansacRows rows(adapter.getNumberCorrespondences(), 3);
  rpe::detail::run_ransac<Tp>(adapter, RPE_SHINJI, rows, nullptr, dist_thre_3d_, Tp(0), Tp(0), Iter, confidence);
  adapter.cvtInlier();  // [reference :210]
}

template <typename Tp>
void shinji_prosac(AOOnlyPoseAdapter<Tp>& adapter, const Tp dist_thre_3d_, int& Iter, Tp confidence = 0.99) {
  rpe::detail::ProsacRows<Tp, AOOnlyPoseAdapter<Tp> > rows(adapter, 3);
  rpe::detail::run_ransac<Tp>(adapter, RPE_SHINJI, rows, nullptr, dist_thre_3d_, Tp(0), Tp(0), Iter, confidence);
  adapter.cvtInlier();  // [reference :268]
}

template <typename Tp>
void shinji_ls(AOPoseAdapter<Tp>& adapter) {  // Kabsch over the 3-D inliers [reference :273-296]
  rpe::detail::run_refit<Tp>(adapter, RPE_REFIT_KABSCH_INLIERS, nullptr, 0);
}
template <typename Tp>
void shinji_ls1(AOOnlyPoseAdapter<Tp>& adapter) {  // [reference :298-320]
  rpe::detail::run_refit<Tp>(adapter, RPE_REFIT_KABSCH_INLIERS, nullptr, 0);
}
template <typename Tp>
void shinji_ls2(AOOnlyPoseAdapter<Tp>& adapter) {  // all correspondences [reference :322-342]
  rpe::detail::run_refit<Tp>(adapter, RPE_REFIT_KABSCH_ALL, nullptr, 0);
}

// assign_sample [reference :344-365]: gather the K = size-1 sample columns (+ the 4th point used to pick the P3P root)
// out of the adapter; true iff all K camera points are valid, i.e. the 3-D--3-D solver can run for this sample.
// (On the device the same gather happens inside the generator kernel; this host form serves callers of the header.)
template <typename Tp>
bool assign_sample(const AOPoseAdapter<Tp>& adapter, const std::vector<int>& selected_cols_, rpe::MatrixX<Tp>* p_X_w_,
                   rpe::MatrixX<Tp>* p_X_c_, rpe::MatrixX<Tp>* p_bv_) {
  const int K = (int)selected_cols_.size() - 1;
  int n_valid = 0;
  for (int k = 0; k < K; ++k) {
    const int c = selected_cols_[k];
    p_X_w_->setCol(k, adapter.getPointGlob(c));
    p_bv_->setCol(k, adapter.getBearingVector(c));
    if (adapter.isValid(c)) {
      p_X_c_->setCol(k, adapter.getPointCurr(c));
      ++n_valid;
    }
  }
  p_X_w_->setCol(3, adapter.getPointGlob(selected_cols_[3]));
  p_bv_->setCol(3, adapter.getBearingVector(selected_cols_[3]));
  return n_valid == K;
}

template <typename Tp>
void shinji_kneip_ransac(AOPoseAdapter<Tp>& adapter, const Tp dist_thre_3d_, const Tp thre_2d_, int& Iter,
                         Tp confidence = 0.99) {
  const Tp cos_thr = std::cos(std::atan(thre_2d_ / adapter.getFocal()));  // [reference :373]
  rpe::detail::RansacRows rows(adapter.getNumberCorrespondences(), 4);
  rpe::detail::run_ransac<Tp>(adapter, RPE_SHINJI_KNEIP, rows, nullptr, dist_thre_3d_, cos_thr, Tp(0), Iter, confidence);
  PnPPoseAdapter<Tp>* pAdapter = &adapter;  // [reference :433-435]
  pAdapter->cvtInlier();
  adapter.cvtInlier();
}

template <typename Tp>
void shinji_kneip_prosac(AOPoseAdapter<Tp>& adapter, const Tp dist_thre_3d_, const Tp thre_2d_, int& Iter,
                         Tp confidence = 0.99) {
  const Tp cos_thr = std::cos(std::atan(thre_2d_ / adapter.getFocal()));  // [reference :446]
  rpe::detail::ProsacRows<Tp, AOPoseAdapter<Tp> > rows(adapter, 4);
  rpe::detail::run_ransac<Tp>(adapter, RPE_SHINJI_KNEIP, rows, nullptr, dist_thre_3d_, cos_thr, Tp(0), Iter, confidence);
  PnPPoseAdapter<Tp>* pAdapter = &adapter;
  pAdapter->cvtInlier();
  adapter.cvtInlier();
}

// North-star addition (no reference counterpart): Levenberg-Marquardt on SE3 over the inliers of every modality the
// adapter carries, starting from the adapter's current pose. weights = {w2d, w3d, wN}.
template <typename Tp, class Adapter>
rpe_result refine_lm(Adapter& adapter, const float* modality_weights = nullptr, int max_iters = 6) {
  return rpe::detail::run_refit<Tp>(adapter, RPE_REFIT_GN, modality_weights, max_iters);
}

#endif  // RPE_ABSOLUTE_ORIENTATION_HPP_
