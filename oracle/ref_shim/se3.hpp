// oracle/ref_shim/se3.hpp — TEST INFRASTRUCTURE (see oracle/README.md). PARITY UNPINNED.
//
// Stand-in for the Sophus SO3 / SE3 classes as far as /root/reference/pose/*.hpp uses them:
// the operations are the ones oracle/sophus_model.hpp restates from /root/reference/sophus/so3.hpp and se3.hpp, behind
// Sophus' own class and member names, so that the pose headers compile unmodified. Where the real Sophus would
// std::abort() in SOPHUS_ENSURE (a non-orthogonal matrix, a zero quaternion) this stand-in counts the event in
// Sophus::shim_ensure_failures() and carries on; the harness reports the counter.
#ifndef ORACLE_REF_SHIM_SE3_HPP_
#define ORACLE_REF_SHIM_SE3_HPP_

#include "eigen_shim.hpp"
#include "../sophus_model.hpp"

namespace Sophus {

inline long long& shim_ensure_failures() {
  static long long n = 0;
  return n;
}

template <class T>
class SO3 {
  orc::SO3<T> r_;
  explicit SO3(const orc::SO3<T>& r) : r_(r) {
    if (!r.ok) ++shim_ensure_failures();
  }

 public:
  typedef Eigen::Matrix<T, 3, 1> Point;
  typedef Eigen::Matrix<T, 3, 3> Transformation;
  SO3() {}
  // so3.hpp:561-566: Quaternion(R) without renormalisation + ENSURE(orthogonal, det > 0)
  template <class O>
  explicit SO3(const Eigen::MatrixBase<O>& R) : r_(orc::SO3<T>::from_matrix(R.to_m3())) {
    if (!r_.ok) ++shim_ensure_failures();
  }
  // so3.hpp:578-585: explicit quaternion, normalised
  explicit SO3(const Eigen::Quaternion<T>& q) : r_(orc::SO3<T>::from_quat(q.q)) {
    if (!r_.ok) ++shim_ensure_failures();
  }
  SO3 inverse() const { return SO3(r_.inverse()); }
  Transformation matrix() const {
    const orc::M3<T> m = r_.matrix();
    Transformation out;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) out(i, j) = m(i, j);
    return out;
  }
  Eigen::Quaternion<T> unit_quaternion() const { return Eigen::Quaternion<T>(r_.q); }
  SO3 operator*(const SO3& o) const { return SO3(r_ * o.r_); }
  template <class O>
  Point operator*(const Eigen::MatrixBase<O>& p) const {
    const orc::V3<T> v = r_ * p.to_v3();
    return Point(v[0], v[1], v[2]);
  }
  const orc::SO3<T>& model() const { return r_; }
};

template <class T>
class SE3 {
  SO3<T> so3_;
  Eigen::Matrix<T, 3, 1> t_;

 public:
  SE3() {}
  template <class O>
  SE3(const SO3<T>& r, const Eigen::MatrixBase<O>& t) : so3_(r), t_(t) {}
  SO3<T>& so3() { return so3_; }
  const SO3<T>& so3() const { return so3_; }
  Eigen::Matrix<T, 3, 1>& translation() { return t_; }
  const Eigen::Matrix<T, 3, 1>& translation() const { return t_; }
};

}  // namespace Sophus

#endif  // ORACLE_REF_SHIM_SE3_HPP_
