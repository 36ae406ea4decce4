// oracle/sophus_model.hpp — TEST INFRASTRUCTURE (see oracle/README.md). PARITY UNPINNED.
//
// The five Sophus operations the pose headers actually execute,
// restated on top of eig_model.hpp. Citations are into /root/reference/sophus/.
#ifndef ORACLE_SOPHUS_MODEL_HPP_
#define ORACLE_SOPHUS_MODEL_HPP_

#include "eig_model.hpp"

namespace orc {

template <class T>
struct SophusEps {
  static T value() { return T(1e-10); }  // common.hpp:137-141
};
template <>
struct SophusEps<float> {
  static float value() { return static_cast<float>(1e-5); }  // common.hpp:143-151
};

template <class T>
struct SO3 {
  Quat<T> q;  // "unit_quaternion_", identity by default (so3.hpp:548-549)
  bool ok;    // false where the reference would std::abort() in SOPHUS_ENSURE (common.hpp:115-132)
  SO3() : q(), ok(true) {}

  // so3.hpp:561-566 — Quaternion(R) WITHOUT renormalisation, then ENSURE(isOrthogonal) and
  // ENSURE(det > 0). rotation_matrix.hpp:13-24: ||R R^T - I||_F < epsilon.
  static SO3 from_matrix(const M3<T>& R) {
    SO3 s;
    s.q = quat_from_matrix(R);
    const M3<T> rrt = R * transpose(R);
    const T dev = frob_norm(rrt - M3<T>::identity());
    s.ok = (dev < SophusEps<T>::value()) && (det3(R) > T(0));
    return s;
  }
  // so3.hpp:578-585 + :190-196 — explicit quaternion ctor normalises: coeffs /= norm().
  static SO3 from_quat(const Quat<T>& qq) {
    SO3 s;
    const T len = std::sqrt(quat_squared_norm(qq));
    s.ok = len >= SophusEps<T>::value();
    s.q.x = qq.x / len;
    s.q.y = qq.y / len;
    s.q.z = qq.z / len;
    s.q.w = qq.w / len;
    return s;
  }
  // so3.hpp:176-178 — inverse() = SO3(conjugate) which goes through the normalising ctor.
  SO3 inverse() const {
    SO3 r = from_quat(Quat<T>(q.w, -q.x, -q.y, -q.z));
    r.ok = r.ok && ok;
    return r;
  }
  // so3.hpp:204-206
  M3<T> matrix() const { return quat_to_matrix(q); }
  // so3.hpp:238-240
  V3<T> operator*(const V3<T>& p) const { return quat_rotate(q, p); }
  // so3.hpp:218-222, 255-272 — Hamilton product, then the cheap first-order renormalisation.
  SO3 operator*(const SO3& o) const {
    SO3 r;
    r.q = quat_mul(q, o.q);
    const T sn = quat_squared_norm(r.q);
    if (sn != T(1.0)) {
      const T k = T(2.0) / (T(1.0) + sn);
      r.q.x *= k;
      r.q.y *= k;
      r.q.z *= k;
      r.q.w *= k;
    }
    r.ok = ok && o.ok;
    return r;
  }
};

// se3.hpp:552-560 (ctor from SO3 + translation), :673-687 (accessors)
template <class T>
struct SE3 {
  SO3<T> so3;
  V3<T> t;
  SE3() {}
  SE3(const SO3<T>& r, const V3<T>& tt) : so3(r), t(tt) {}
};

}  // namespace orc

#endif  // ORACLE_SOPHUS_MODEL_HPP_
