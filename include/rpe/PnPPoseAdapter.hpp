// rpe/PnPPoseAdapter.hpp — mirrors /root/reference/pose/PnPPoseAdapter.hpp:27-255 (2-D/3-D correspondences).
//
// Constructor order (bearingVectors, points_g [, t], [R]) as :45-62. Inlier flags are shorts (one per
// correspondence) like the reference's Matrix<short,Dynamic,1>; the compacted index list is int32 — the
// reference's `for (short r = 0; r < (short)rows; r++)` (:232) silently yields nothing above 32 767 rows.
#ifndef RPE_PNP_POSE_ADAPTER_HPP_
#define RPE_PNP_POSE_ADAPTER_HPP_

#include <string.h>

#include <iostream>
#include <vector>

#include "PoseAdapterBase.hpp"
#include "Utility.hpp"

template <typename Tp>
class PnPPoseAdapter : public PoseAdapterBase<Tp> {
 protected:
  using PoseAdapterBase<Tp>::_t_w;
  using PoseAdapterBase<Tp>::_R_cw;

 public:
  typedef typename PoseAdapterBase<Tp>::Vector3 Vector3;
  typedef typename PoseAdapterBase<Tp>::SO3_T SO3_T;
  typedef typename PoseAdapterBase<Tp>::Point3 Point3;
  typedef rpe::MatrixX<Tp> MatrixX;

  template <class M>
  PnPPoseAdapter(const M& bearingVectors, const M& points)
      : PoseAdapterBase<Tp>(), _bearingVectors(rpe::View3<Tp>::of(bearingVectors)), _points_g(rpe::View3<Tp>::of(points)) {
    init();
  }
  template <class M>
  PnPPoseAdapter(const M& bearingVectors, const M& points, const SO3_T& R)
      : PoseAdapterBase<Tp>(R), _bearingVectors(rpe::View3<Tp>::of(bearingVectors)), _points_g(rpe::View3<Tp>::of(points)) {
    init();
  }
  template <class M>
  PnPPoseAdapter(const M& bearingVectors, const M& points, const Vector3& t, const SO3_T& R)
      : PoseAdapterBase<Tp>(t, R), _bearingVectors(rpe::View3<Tp>::of(bearingVectors)), _points_g(rpe::View3<Tp>::of(points)) {
    init();
  }
  virtual ~PnPPoseAdapter() {}

  virtual Point3 getBearingVector(int index) const { return _bearingVectors.col(index); }
  virtual Tp getWeight(int) const { return Tp(1.); }
  virtual Point3 getPointGlob(int index) const { return _points_g.col(index); }
  virtual int getNumberCorrespondences() const { return _bearingVectors.n; }

  // `inliers` is n x m column-major; this class keeps column 0 (2-D flags)   [reference :196-202]
  virtual void setInlier(const rpe::MaskX& inliers) {
    memcpy(_inliers.data(), inliers.data(), sizeof(short) * _inliers.size());
    this->_rpe_state_token = 0;
  }
  virtual void setWeights(const MatrixX& weights) {  // column 0   [reference :212-219]
    _weights.assign(weights.colPtr(0), weights.colPtr(0) + weights.rows());
  }
  virtual void printInlier() const {
    for (size_t i = 0; i < _inliers.size(); ++i) std::cout << _inliers[i] << " ";
    std::cout << std::endl;
  }
  const std::vector<int>& getInlierIdx() const { return _vInliersPnP; }
  void cvtInlier() {  // [reference :227-237]
    _vInliersPnP.clear();
    for (int r = 0; r < (int)_inliers.size(); r++)
      if (1 == _inliers[r]) _vInliersPnP.push_back(r);
  }
  Tp getError(int index) const {  // sine of the angle between the reprojected ray and the bearing vector [:204-210]
    Point3 Xc = _R_cw * getPointGlob(index) + _t_w;
    Xc.normalize();
    return Xc.cross(getBearingVector(index)).norm();
  }
  void setMaxVotes(int votes) { _max_votes = votes; }
  int getMaxVotes() { return _max_votes; }
  bool isInlier23(int index) const { return _inliers[index] == 1; }
  Tp weight23(int index) const { return _weights.empty() ? Tp(1.0) : _weights[index]; }
  void sortIdx() { _idx = sortIndexes<Tp>(_weights); }  // [reference :239-244]
  void getSortedIdx(std::vector<int>& select_) const {  // [reference :246-255]
    for (int i = 0; i < (int)select_.size(); ++i) {
      const int j = select_[i];
      if (j < (int)_idx.size()) select_[i] = _idx[j];
    }
  }
  const std::vector<Tp>& rpeWeights23() const { return _weights; }

  virtual void rpeArrays(const Tp** bv, const Tp** xc, const Tp** nc, const Tp** xw, const Tp** nw) const {
    *bv = _bearingVectors.p;
    *xc = nullptr;
    *nc = nullptr;
    *xw = _points_g.p;
    *nw = nullptr;
  }
  virtual int rpeMask(std::vector<short>* flags) const {
    *flags = _inliers;
    return 1;
  }

 protected:
  void init() {
    _inliers.assign(_bearingVectors.n, 1);  // setOnes() [reference :118-119]
    _max_votes = 0;
  }
  rpe::View3<Tp> _bearingVectors;  // unit bearing vectors, camera frame
  rpe::View3<Tp> _points_g;        // points, world frame
  std::vector<short> _inliers;
  std::vector<Tp> _weights;
  std::vector<int> _idx;  // weight-sorted order for PROSAC
  std::vector<int> _vInliersPnP;
  int _max_votes;
};

#endif  // RPE_PNP_POSE_ADAPTER_HPP_
