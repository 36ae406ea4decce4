// rpe/PoseAdapterBase.hpp — mirrors /root/reference/pose/PoseAdapterBase.hpp:28-146.
//
// Same class name, same accessor names and meanings (getBearingVector/getWeight/getPointGlob/
// getNumberCorrespondences :81-101, gettw/sett/getRcw/setRcw/setFocal/getFocal/getTcw :109-130, state
// _t_w/_R_cw/_fx,_fy,_cx,_cy :133-144). Added, for the GPU back end only: rpeArrays() hands the raw
// column-major arrays to the C-ABI, and a revision counter lets the estimators skip redundant transfers.
#ifndef RPE_POSE_ADAPTERBASE_HPP_
#define RPE_POSE_ADAPTERBASE_HPP_

#include <stdlib.h>

#include <vector>

#include "so3.hpp"
#include "types.hpp"

template <typename Tp>
class PoseAdapterBase {
 public:
  typedef rpe::Vec3<Tp> Vector3;
  typedef rpe::SO3<Tp> SO3_T;
  typedef rpe::SE3<Tp> SE3_T;
  typedef rpe::Vec3<Tp> Point3;

  PoseAdapterBase() : _t_w(Vector3::Zero()), _fx(0), _fy(0), _cx(0), _cy(0), _rpe_state_token(0) {}
  explicit PoseAdapterBase(const SO3_T& R) : _t_w(Vector3::Zero()), _R_cw(R), _fx(0), _fy(0), _cx(0), _cy(0), _rpe_state_token(0) {}
  PoseAdapterBase(const Vector3& t, const SO3_T& R) : _t_w(t), _R_cw(R), _fx(0), _fy(0), _cx(0), _cy(0), _rpe_state_token(0) {}
  virtual ~PoseAdapterBase() {}

  // access of correspondences
  virtual Point3 getBearingVector(int index) const = 0;
  virtual Tp getWeight(int index) const = 0;
  virtual Point3 getPointGlob(int index) const = 0;
  virtual int getNumberCorrespondences() const = 0;

  // access of priors or known values
  Vector3 gettw() const { return _t_w; }
  void sett(const Vector3& t) {
    _t_w = t;
    _rpe_state_token = 0;
  }
  SO3_T getRcw() const { return _R_cw; }
  void setRcw(const SO3_T& R) {
    _R_cw = R;
    _rpe_state_token = 0;
  }
  void setFocal(const Tp fx, const Tp fy) {
    _fx = fx;
    _fy = fy;
  }
  Tp getFocal() const { return (_fx + _fy) / 2; }
  SE3_T getTcw() { return SE3_T(_R_cw, _t_w); }

  // ---- GPU back end plumbing (not part of the reference interface) ----
  // raw column-major 3 x n arrays in the C-ABI order; nullptr where the adapter has no such modality
  virtual void rpeArrays(const Tp** bv, const Tp** xc, const Tp** nc, const Tp** xw, const Tp** nw) const = 0;
  // n x cols flags in the layout setInlier() takes; cols = 0 when the adapter has no flags yet
  virtual int rpeMask(std::vector<short>* flags) const = 0;
  unsigned long long rpeStateToken() const { return _rpe_state_token; }
  void rpeSetStateToken(unsigned long long t) { _rpe_state_token = t; }

 protected:
  Vector3 _t_w;  // translation of x_c = R_cw x_w + t_w
  SO3_T _R_cw;   // rotation world -> camera
  Tp _fx, _fy, _cx, _cy;
  unsigned long long _rpe_state_token;  // == Session::token() while the device still mirrors pose + flags
};

#endif  // RPE_POSE_ADAPTERBASE_HPP_
