// rpe/AOPoseAdapter.hpp — mirrors /root/reference/pose/AOPoseAdapter.hpp:26-217 (2-D + 3-D correspondences).
//
// Constructor order (bearingVectors, points_c, points_g [, t], [R]) as :47-69. isValid keeps the reference's
// "any coordinate is not NaN" test (:147-152); weight33 keeps the division by 32 767 (:161-169); setWeights
// keeps the rows()==1 test of :186-198.
#ifndef RPE_AO_POSE_ADAPTER_HPP_
#define RPE_AO_POSE_ADAPTER_HPP_

#include <limits>

#include "PnPPoseAdapter.hpp"

template <typename Tp>
class AOPoseAdapter : public PnPPoseAdapter<Tp> {
 protected:
  using PoseAdapterBase<Tp>::_t_w;
  using PoseAdapterBase<Tp>::_R_cw;
  using PnPPoseAdapter<Tp>::_bearingVectors;
  using PnPPoseAdapter<Tp>::_points_g;

 public:
  typedef typename PoseAdapterBase<Tp>::Vector3 Vector3;
  typedef typename PoseAdapterBase<Tp>::SO3_T SO3_T;
  typedef typename PoseAdapterBase<Tp>::Point3 Point3;
  typedef typename PnPPoseAdapter<Tp>::MatrixX MatrixX;

  template <class M>
  AOPoseAdapter(const M& bearingVectors, const M& points_c, const M& points_g)
      : PnPPoseAdapter<Tp>(bearingVectors, points_g), _points_c(rpe::View3<Tp>::of(points_c)) {
    _inliers_3d.assign(_bearingVectors.n, 1);
  }
  template <class M>
  AOPoseAdapter(const M& bearingVectors, const M& points_c, const M& points_g, const SO3_T& R)
      : PnPPoseAdapter<Tp>(bearingVectors, points_g, R), _points_c(rpe::View3<Tp>::of(points_c)) {
    _inliers_3d.assign(_bearingVectors.n, 1);
  }
  template <class M>
  AOPoseAdapter(const M& bearingVectors, const M& points_c, const M& points_g, const Vector3& t, const SO3_T& R)
      : PnPPoseAdapter<Tp>(bearingVectors, points_g, t, R), _points_c(rpe::View3<Tp>::of(points_c)) {
    _inliers_3d.assign(_bearingVectors.n, 1);
  }
  virtual ~AOPoseAdapter() {}

  bool isInlier33(int index) const { return _inliers_3d[index] == 1; }
  Tp weight33(int index) const {
    return _weights_3d.empty() ? Tp(1.0) : Tp(_weights_3d[index]) / std::numeric_limits<short>::max();
  }
  virtual Point3 getPointCurr(int index) const { return _points_c.col(index); }
  virtual bool isValid(int index) const {
    const Point3 p = _points_c.col(index);
    return p[0] == p[0] || p[1] == p[1] || p[2] == p[2];
  }
  virtual void setInlier(const rpe::MaskX& inliers) {  // [reference :171-184]
    PnPPoseAdapter<Tp>::setInlier(inliers);
    if (inliers.cols() != 1) _inliers_3d.assign(inliers.colPtr(1), inliers.colPtr(1) + inliers.rows());
  }
  virtual void setWeights(const MatrixX& weights) {
    PnPPoseAdapter<Tp>::setWeights(weights);
    if (weights.rows() != 1) _weights_3d.assign(weights.colPtr(1), weights.colPtr(1) + weights.rows());
  }
  virtual void printInlier() const {
    PnPPoseAdapter<Tp>::printInlier();
    for (size_t i = 0; i < _inliers_3d.size(); ++i) std::cout << _inliers_3d[i] << " ";
    std::cout << std::endl;
  }
  const std::vector<int>& getInlierIdx() const { return _vInliersAO; }
  void cvtInlier() {  // [reference :207-217]
    _vInliersAO.clear();
    for (int r = 0; r < (int)_inliers_3d.size(); r++)
      if (1 == _inliers_3d[r]) _vInliersAO.push_back(r);
  }
  const std::vector<Tp>& rpeWeights33() const { return _weights_3d; }

  virtual void rpeArrays(const Tp** bv, const Tp** xc, const Tp** nc, const Tp** xw, const Tp** nw) const {
    PnPPoseAdapter<Tp>::rpeArrays(bv, xc, nc, xw, nw);
    *xc = _points_c.p;
  }
  virtual int rpeMask(std::vector<short>* flags) const {
    PnPPoseAdapter<Tp>::rpeMask(flags);
    flags->insert(flags->end(), _inliers_3d.begin(), _inliers_3d.end());
    return 2;
  }

 protected:
  rpe::View3<Tp> _points_c;  // points, camera frame
  std::vector<short> _inliers_3d;
  std::vector<Tp> _weights_3d;
  std::vector<int> _vInliersAO;
};

#endif  // RPE_AO_POSE_ADAPTER_HPP_
