// rpe/AbsoluteOrientationNormal.hpp — mirrors /root/reference/pose/AbsoluteOrientationNormal.hpp.
//
//   nl_2p                       :77-142    one oriented point + one point (host; same template as the device generator)
//   nl_kneip_ransac             :215-284   -> rpe_ransac(RPE_NL_KNEIP)
//   nl_shinji_ransac            :286-354   -> rpe_ransac(RPE_NL_SHINJI)
//   nl_shinji_kneip_ransac      :356-445   -> rpe_ransac(RPE_NL_SHINJI_KNEIP)
//   nl_shinji_kneip_ls          :447-552   -> rpe_refit(RPE_REFIT_NL_SK_LS)   (find_opt_cc :13-46 runs inside it)
#ifndef RPE_ABSOLUTE_ORIENTATION_NORMAL_HPP_
#define RPE_ABSOLUTE_ORIENTATION_NORMAL_HPP_

#include "AbsoluteOrientation.hpp"
#include "NormalAOPoseAdapter.hpp"

template <typename Tp>
void nl_2p(const rpe::Vec3<Tp>& pt1_c, const rpe::Vec3<Tp>& nl1_c, const rpe::Vec3<Tp>& pt2_c, const rpe::Vec3<Tp>& pt1_w,
           const rpe::Vec3<Tp>& nl1_w, const rpe::Vec3<Tp>& pt2_w, rpe::SE3<Tp>* p_solution) {
  Tp q[4], t[3];
  rpe::nl_2p<Tp>(pt1_c.v, nl1_c.v, pt2_c.v, pt1_w.v, nl1_w.v, pt2_w.v, q, t);
  *p_solution = rpe::SE3<Tp>(rpe::SO3<Tp>::fromRawQuaternion(q), rpe::Vec3<Tp>(t[0], t[1], t[2]));
}

namespace rpe {
namespace detail {
template <class Tp>
inline void nl_cvt_all(NormalAOPoseAdapter<Tp>& adapter, bool pnp, bool ao) {
  if (pnp) {
    PnPPoseAdapter<Tp>* p = &adapter;
    p->cvtInlier();
  }
  if (ao) {
    AOPoseAdapter<Tp>* p = &adapter;
    p->cvtInlier();
  }
  adapter.cvtInlier();
}
}  // namespace detail
}  // namespace rpe

// assign_sample [reference :48-75]: as the AOPoseAdapter form, with the normals of both frames.
template <typename Tp>
bool assign_sample(const NormalAOPoseAdapter<Tp>& adapter, const std::vector<int>& selected_cols_, rpe::MatrixX<Tp>* p_X_w_,
                   rpe::MatrixX<Tp>* p_N_w_, rpe::MatrixX<Tp>* p_X_c_, rpe::MatrixX<Tp>* p_N_c_, rpe::MatrixX<Tp>* p_bv_) {
  const int K = (int)selected_cols_.size() - 1;
  int n_valid = 0;
  for (int k = 0; k < K; ++k) {
    const int c = selected_cols_[k];
    p_X_w_->setCol(k, adapter.getPointGlob(c));
    p_N_w_->setCol(k, adapter.getNormalGlob(c));
    p_bv_->setCol(k, adapter.getBearingVector(c));
    if (adapter.isValid(c)) {
      p_X_c_->setCol(k, adapter.getPointCurr(c));
      p_N_c_->setCol(k, adapter.getNormalCurr(c));
      ++n_valid;
    }
  }
  const int c3 = selected_cols_[3];
  p_X_w_->setCol(3, adapter.getPointGlob(c3));
  p_N_w_->setCol(3, adapter.getNormalGlob(c3));
  p_bv_->setCol(3, adapter.getBearingVector(c3));
  return n_valid == K;
}

template <typename Tp>
void nl_kneip_ransac(NormalAOPoseAdapter<Tp>& adapter, const Tp thre_2d_, const Tp nl_thre, int& Iter, Tp confidence = 0.99) {
  const Tp cos_thr = std::cos(std::atan(thre_2d_ / adapter.getFocal()));  // [reference :222]
  const Tp cos_nl_thre = std::cos(nl_thre);                          // [reference :223]
  rpe::detail::RansacRows rows(adapter.getNumberCorrespondences(), 4);
  rpe::detail::run_ransac<Tp>(adapter, RPE_NL_KNEIP, rows, nullptr, Tp(0), cos_thr, cos_nl_thre, Iter, confidence);
  rpe::detail::nl_cvt_all(adapter, true, false);  // [reference :279-281]
}

template <typename Tp>
void nl_shinji_ransac(NormalAOPoseAdapter<Tp>& adapter, const Tp thre_3d_, const Tp nl_thre, int& Iter, Tp confidence = 0.99) {
  const Tp cos_nl_thre = std::cos(nl_thre);  // [reference :294]
  rpe::detail::RansacRows rows(adapter.getNumberCorrespondences(), 4);
  rpe::detail::run_ransac<Tp>(adapter, RPE_NL_SHINJI, rows, nullptr, thre_3d_, Tp(0), cos_nl_thre, Iter, confidence);
  rpe::detail::nl_cvt_all(adapter, false, true);  // [reference :350-352]
}

template <typename Tp>
void nl_shinji_kneip_ransac(NormalAOPoseAdapter<Tp>& adapter, const Tp thre_3d_, const Tp thre_2d_, const Tp nl_thre,
                            int& Iter, Tp confidence = 0.99) {
  const Tp cos_thr = std::cos(std::atan(thre_2d_ / adapter.getFocal()));  // [reference :363]
  const Tp cos_nl_thre = std::cos(nl_thre);                          // [reference :364]
  rpe::detail::RansacRows rows(adapter.getNumberCorrespondences(), 4);
  rpe::detail::run_ransac<Tp>(adapter, RPE_NL_SHINJI_KNEIP, rows, nullptr, thre_3d_, cos_thr, cos_nl_thre, Iter, confidence);
  rpe::detail::nl_cvt_all(adapter, true, true);  // [reference :439-443]
}

// The reference's iterative weighted refinement over all three modalities. Per-correspondence weights are the
// ones given to adapter.setWeights (n x 3), or the adapters' default of 1.
template <typename Tp>
void nl_shinji_kneip_ls(NormalAOPoseAdapter<Tp>& adapter) {
  if (adapter.getMaxVotes() == 0) return;  // [reference :454]
  const int n = adapter.getNumberCorrespondences();
  std::vector<float> w;
  const std::vector<Tp>&w2 = adapter.rpeWeights23(), &w3 = adapter.rpeWeights33(), &wn = adapter.rpeWeightsNN();
  if (!w2.empty() || !w3.empty() || !wn.empty()) {
    // the adapters fall back to 1 for a column that was never set; 3-3 and N-N weights are divided by 32 767 on use
    w.assign((size_t)3 * n, 1.f);
    for (int i = 0; i < n; ++i) {
      if (!w2.empty()) w[i] = (float)w2[i];
      w[n + i] = w3.empty() ? 32767.f : (float)w3[i];
      w[2 * n + i] = wn.empty() ? 32767.f : (float)wn[i];
    }
  }
  rpe::detail::run_refit<Tp>(adapter, RPE_REFIT_NL_SK_LS, w.empty() ? nullptr : w.data(), 0);
}

#endif  // RPE_ABSOLUTE_ORIENTATION_NORMAL_HPP_
