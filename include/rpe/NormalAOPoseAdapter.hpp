// rpe/NormalAOPoseAdapter.hpp — mirrors /root/reference/pose/NormalAOPoseAdapter.hpp:16-231
// (2-D + 3-D + normal correspondences). Constructor order (bearingVectors, points_c, normal_c, points_g,
// normal_g [, t], [R]) as :37-63.
#ifndef RPE_NORMAL_AO_POSE_ADAPTER_HPP_
#define RPE_NORMAL_AO_POSE_ADAPTER_HPP_

#include "AOPoseAdapter.hpp"

template <typename Tp>
class NormalAOPoseAdapter : public AOPoseAdapter<Tp> {
 protected:
  using PoseAdapterBase<Tp>::_t_w;
  using PoseAdapterBase<Tp>::_R_cw;
  using PnPPoseAdapter<Tp>::_bearingVectors;

 public:
  typedef typename PoseAdapterBase<Tp>::Vector3 Vector3;
  typedef typename PoseAdapterBase<Tp>::SO3_T SO3_T;
  typedef typename PoseAdapterBase<Tp>::Point3 Point3;
  typedef typename PnPPoseAdapter<Tp>::MatrixX MatrixX;

  template <class M>
  NormalAOPoseAdapter(const M& bearingVectors, const M& points_c, const M& normal_c, const M& points_g, const M& normal_g)
      : AOPoseAdapter<Tp>(bearingVectors, points_c, points_g), _normal_c(rpe::View3<Tp>::of(normal_c)),
        _normal_g(rpe::View3<Tp>::of(normal_g)) {
    _inliers_nl.assign(_bearingVectors.n, 1);
  }
  template <class M>
  NormalAOPoseAdapter(const M& bearingVectors, const M& points_c, const M& normal_c, const M& points_g, const M& normal_g,
                      const SO3_T& R)
      : AOPoseAdapter<Tp>(bearingVectors, points_c, points_g, R), _normal_c(rpe::View3<Tp>::of(normal_c)),
        _normal_g(rpe::View3<Tp>::of(normal_g)) {
    _inliers_nl.assign(_bearingVectors.n, 1);
  }
  template <class M>
  NormalAOPoseAdapter(const M& bearingVectors, const M& points_c, const M& normal_c, const M& points_g, const M& normal_g,
                      const Vector3& t, const SO3_T& R)
      : AOPoseAdapter<Tp>(bearingVectors, points_c, points_g, t, R), _normal_c(rpe::View3<Tp>::of(normal_c)),
        _normal_g(rpe::View3<Tp>::of(normal_g)) {
    _inliers_nl.assign(_bearingVectors.n, 1);
  }
  virtual ~NormalAOPoseAdapter() {}

  bool isInlierNN(int index) const { return _inliers_nl[index] == 1; }
  Tp weightNN(int index) const {
    return _weights_nl.empty() ? Tp(1.0) : Tp(_weights_nl[index]) / std::numeric_limits<short>::max();
  }
  virtual Point3 getNormalCurr(int index) const { return _normal_c.col(index); }
  virtual Point3 getNormalGlob(int index) const { return _normal_g.col(index); }
  virtual void setInlier(const rpe::MaskX& inliers) {  // [reference :179-195]
    if (inliers.cols() == 1) PnPPoseAdapter<Tp>::setInlier(inliers);
    if (inliers.cols() == 2) AOPoseAdapter<Tp>::setInlier(inliers);
    if (inliers.cols() == 3) {
      AOPoseAdapter<Tp>::setInlier(inliers);
      _inliers_nl.assign(inliers.colPtr(2), inliers.colPtr(2) + inliers.rows());
    }
  }
  virtual void setWeights(const MatrixX& weights) {  // [reference :197-212]
    if (weights.cols() == 1) PnPPoseAdapter<Tp>::setWeights(weights);
    if (weights.cols() == 2) AOPoseAdapter<Tp>::setWeights(weights);
    if (weights.cols() == 3) {
      AOPoseAdapter<Tp>::setWeights(weights);
      _weights_nl.assign(weights.colPtr(2), weights.colPtr(2) + weights.rows());
    }
  }
  virtual void printInlier() const {
    AOPoseAdapter<Tp>::printInlier();
    for (size_t i = 0; i < _inliers_nl.size(); ++i) std::cout << _inliers_nl[i] << " ";
    std::cout << std::endl;
  }
  const std::vector<int>& getInlierIdx() const { return _vInliersNN; }
  void cvtInlier() {
    _vInliersNN.clear();
    for (int r = 0; r < (int)_inliers_nl.size(); r++)
      if (1 == _inliers_nl[r]) _vInliersNN.push_back(r);
  }
  const std::vector<Tp>& rpeWeightsNN() const { return _weights_nl; }

  virtual void rpeArrays(const Tp** bv, const Tp** xc, const Tp** nc, const Tp** xw, const Tp** nw) const {
    AOPoseAdapter<Tp>::rpeArrays(bv, xc, nc, xw, nw);
    *nc = _normal_c.p;
    *nw = _normal_g.p;
  }
  virtual int rpeMask(std::vector<short>* flags) const {
    AOPoseAdapter<Tp>::rpeMask(flags);
    flags->insert(flags->end(), _inliers_nl.begin(), _inliers_nl.end());
    return 3;
  }

 protected:
  rpe::View3<Tp> _normal_c;  // normals, camera frame
  rpe::View3<Tp> _normal_g;  // normals, world frame
  std::vector<short> _inliers_nl;
  std::vector<Tp> _weights_nl;
  std::vector<int> _vInliersNN;
};

#endif  // RPE_NORMAL_AO_POSE_ADAPTER_HPP_
