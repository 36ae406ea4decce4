// rpe/so3.hpp — the slice of Sophus::SO3 / Sophus::SE3 the pose headers use.
//
// The reference vendors Sophus (sophus/so3.hpp, se3.hpp) on top of Eigen. The adapters only need:
//   SO3()  identity (so3.hpp:548-549)            SO3(Matrix3)  Shoemake quaternion, NOT renormalised, with the
//   SO3(quaternion)  normalising (so3.hpp:578)   orthogonality / det>0 checks of SOPHUS_ENSURE (so3.hpp:561-566)
//   matrix() (so3.hpp:204)   inverse() (so3.hpp:176)   operator*(point) (so3.hpp:238)   operator*(SO3) (so3.hpp:218,255)
//   unit_quaternion()        SE3(SO3, t), so3(), translation() (se3.hpp:552-560, 673-687)
// They are provided here in namespace Sophus unless a real Sophus has been included first.
// Where the reference would std::abort() (common.hpp:115-132) this class records `ok() == false` instead.
#ifndef RPE_SO3_HPP_
#define RPE_SO3_HPP_

#include "solvers_p3p.h"
#include "types.hpp"

namespace rpe {

template <class Tp>
struct Quaternion {
  Tp c[4];  // x, y, z, w (Eigen coefficient order)
  Quaternion() : c{Tp(0), Tp(0), Tp(0), Tp(1)} {}
  Quaternion(Tp w, Tp x, Tp y, Tp z) : c{x, y, z, w} {}
  Tp& x() { return c[0]; }
  Tp& y() { return c[1]; }
  Tp& z() { return c[2]; }
  Tp& w() { return c[3]; }
  Tp x() const { return c[0]; }
  Tp y() const { return c[1]; }
  Tp z() const { return c[2]; }
  Tp w() const { return c[3]; }
  const Tp* coeffs() const { return c; }
  Tp* coeffs() { return c; }
  Tp norm() const { return std::sqrt((c[0] * c[0] + c[1] * c[1]) + (c[2] * c[2] + c[3] * c[3])); }
  Quaternion conjugate() const { return Quaternion(c[3], -c[0], -c[1], -c[2]); }
};

template <class Tp>
class SO3 {
 public:
  typedef Vec3<Tp> Point;
  typedef Mat3<Tp> Transformation;
  SO3() : ok_(true) {}
  SO3(const Transformation& R) { ok_ = so3_from_matrix<Tp>(R.m, q_.c); }  // no renormalisation, like the reference
  explicit SO3(const Quaternion<Tp>& q) {
    const Tp len = q.norm();
    ok_ = len >= Lim<Tp>::sophus_eps();
    for (int k = 0; k < 4; ++k) q_.c[k] = q.c[k] / len;
  }
  static SO3 fromRawQuaternion(const Tp xyzw[4]) {  // adopt (x,y,z,w) bits as they are (results coming back from the GPU)
    SO3 s;
    for (int k = 0; k < 4; ++k) s.q_.c[k] = xyzw[k];
    return s;
  }
  bool ok() const { return ok_; }
  const Quaternion<Tp>& unit_quaternion() const { return q_; }
  Transformation matrix() const {
    Transformation R;
    quat_to_matrix_t<Tp>(q_.c, R.m);
    return R;
  }
  SO3 inverse() const { return SO3(q_.conjugate()); }
  Point operator*(const Point& p) const {
    Point r;
    quat_rotate<Tp>(q_.c, p.v, r.v);
    return r;
  }
  SO3 operator*(const SO3& o) const {
    SO3 r;
    so3_mul<Tp>(q_.c, o.q_.c, r.q_.c);
    r.ok_ = ok_ && o.ok_;
    return r;
  }

 private:
  Quaternion<Tp> q_;
  bool ok_;
};

template <class Tp>
class SE3 {
 public:
  SE3() {}
  SE3(const SO3<Tp>& R, const Vec3<Tp>& t) : R_(R), t_(t) {}
  SO3<Tp>& so3() { return R_; }
  const SO3<Tp>& so3() const { return R_; }
  Vec3<Tp>& translation() { return t_; }
  const Vec3<Tp>& translation() const { return t_; }
  Vec3<Tp> operator*(const Vec3<Tp>& p) const { return R_ * p + t_; }

 private:
  SO3<Tp> R_;
  Vec3<Tp> t_;
};

}  // namespace rpe

#if !defined(SOPHUS_SO3_HPP) && !defined(RPE_NO_SOPHUS_ALIAS)
namespace Sophus {
template <class Tp>
using SO3 = rpe::SO3<Tp>;
template <class Tp>
using SE3 = rpe::SE3<Tp>;
}  // namespace Sophus
#endif

#endif  // RPE_SO3_HPP_
