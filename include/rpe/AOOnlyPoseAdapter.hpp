// rpe/AOOnlyPoseAdapter.hpp — mirrors /root/reference/pose/AOOnlyPoseAdapter.hpp:26-254 (3-D/3-D only).
// Constructor order (points_c, points_g [, t], [R]) as :46-64. getBearingVector returns a default point
// (:73), weight33 is NOT divided by 32 767 here (:176-183), setInlier keeps column 1 (:185-198).
#ifndef RPE_AO_ONLY_POSE_ADAPTER_HPP_
#define RPE_AO_ONLY_POSE_ADAPTER_HPP_

#include <iostream>
#include <vector>

#include "PoseAdapterBase.hpp"
#include "Utility.hpp"

template <typename Tp>
class AOOnlyPoseAdapter : public PoseAdapterBase<Tp> {
 protected:
  using PoseAdapterBase<Tp>::_t_w;
  using PoseAdapterBase<Tp>::_R_cw;

 public:
  typedef typename PoseAdapterBase<Tp>::Vector3 Vector3;
  typedef typename PoseAdapterBase<Tp>::SO3_T SO3_T;
  typedef typename PoseAdapterBase<Tp>::Point3 Point3;
  typedef rpe::MatrixX<Tp> MatrixX;

  template <class M>
  AOOnlyPoseAdapter(const M& points_c, const M& points_g)
      : PoseAdapterBase<Tp>(), _points_c(rpe::View3<Tp>::of(points_c)), _points_g(rpe::View3<Tp>::of(points_g)) {
    init();
  }
  template <class M>
  AOOnlyPoseAdapter(const M& points_c, const M& points_g, const SO3_T& R)
      : PoseAdapterBase<Tp>(R), _points_c(rpe::View3<Tp>::of(points_c)), _points_g(rpe::View3<Tp>::of(points_g)) {
    init();
  }
  template <class M>
  AOOnlyPoseAdapter(const M& points_c, const M& points_g, const Vector3& t, const SO3_T& R)
      : PoseAdapterBase<Tp>(t, R), _points_c(rpe::View3<Tp>::of(points_c)), _points_g(rpe::View3<Tp>::of(points_g)) {
    init();
  }
  virtual ~AOOnlyPoseAdapter() {}

  bool isInlier33(int index) const { return _inliers_3d[index] == 1; }
  Tp weight33(int index) const { return _weights_3d.empty() ? Tp(1.0) : _weights_3d[index]; }
  virtual Point3 getBearingVector(int) const { return Point3(); }
  virtual Point3 getPointCurr(int index) const { return _points_c.col(index); }
  virtual Point3 getPointGlob(int index) const { return _points_g.col(index); }
  virtual Tp getWeight(int) const { return Tp(1.); }
  virtual int getNumberCorrespondences() const { return _points_g.n; }
  void setMaxVotes(int votes) { _max_votes = votes; }
  int getMaxVotes() { return _max_votes; }
  virtual bool isValid(int index) const {
    const Point3 p = _points_c.col(index);
    return p[0] == p[0] || p[1] == p[1] || p[2] == p[2];
  }
  virtual void setInlier(const rpe::MaskX& inliers) {
    if (inliers.cols() != 1) _inliers_3d.assign(inliers.colPtr(1), inliers.colPtr(1) + inliers.rows());
    this->_rpe_state_token = 0;
  }
  virtual void setWeights(const MatrixX& weights) {
    if (weights.rows() != 1) _weights_3d.assign(weights.colPtr(1), weights.colPtr(1) + weights.rows());
  }
  virtual void printInlier() const {
    for (size_t i = 0; i < _inliers_3d.size(); ++i) std::cout << _inliers_3d[i] << " ";
    std::cout << std::endl;
  }
  const std::vector<int>& getInlierIdx() const { return _vInliersAO; }
  void cvtInlier() {
    _vInliersAO.clear();
    for (int r = 0; r < (int)_inliers_3d.size(); r++)
      if (1 == _inliers_3d[r]) _vInliersAO.push_back(r);
  }
  void sortIdx() { _idx = sortIndexes<Tp>(_weights_3d); }
  void getSortedIdx(std::vector<int>& select_) const {
    for (int i = 0; i < (int)select_.size(); ++i) {
      const int j = select_[i];
      if (j < (int)_idx.size()) select_[i] = _idx[j];
    }
  }

  virtual void rpeArrays(const Tp** bv, const Tp** xc, const Tp** nc, const Tp** xw, const Tp** nw) const {
    *bv = nullptr;
    *xc = _points_c.p;
    *nc = nullptr;
    *xw = _points_g.p;
    *nw = nullptr;
  }
  virtual int rpeMask(std::vector<short>* flags) const {
    flags->assign(_inliers_3d.size(), 0);  // column 0 (2-D) is unused by this adapter
    flags->insert(flags->end(), _inliers_3d.begin(), _inliers_3d.end());
    return 2;
  }

 protected:
  void init() {
    _inliers_3d.assign(_points_c.n, 1);
    _max_votes = 0;
  }
  rpe::View3<Tp> _points_c;
  rpe::View3<Tp> _points_g;
  std::vector<short> _inliers_3d;
  std::vector<Tp> _weights_3d;
  std::vector<int> _idx;
  std::vector<int> _vInliersAO;
  int _max_votes;
};

#endif  // RPE_AO_ONLY_POSE_ADAPTER_HPP_
