// rpe/MinimalSolvers.hpp — mirrors /root/reference/pose/MinimalSolvers.hpp.
//
// The reference file is never called and half-finished: `ms` stops after step 4 of a 2-point+normal solver and
// mutates const references (:22-46); the eigenvector half of `ev` does not compile (:92). Built here is their
// intent: `ev` = closed-form eigenvalues of a symmetric 3x3 (:49-83), `ms` = the working point+normal minimal
// solver (nl_2p, AbsoluteOrientationNormal.hpp:77-142) applied to the first correspondence pair.
#ifndef RPE_MINIMAL_SOLVERS_HPP_
#define RPE_MINIMAL_SOLVERS_HPP_

#include <cmath>

#include "so3.hpp"
#include "solvers_p3p.h"

template <class T>
void ev(const rpe::Mat3<T>& M_, rpe::Vec3<T>* pE_) {
  const T p1 = M_(0, 1) * M_(0, 1) + M_(0, 2) * M_(0, 2) + M_(1, 2) * M_(1, 2);
  if (std::fabs(p1) < 0.00001) {  // diagonal
    (*pE_)(0) = M_(0, 0);
    (*pE_)(1) = M_(1, 1);
    (*pE_)(2) = M_(2, 2);
    return;
  }
  T q = M_(0, 0) + M_(1, 1) + M_(2, 2);
  q /= 3;
  const T t1 = M_(0, 0) - q, t2 = M_(1, 1) - q, t3 = M_(2, 2) - q;
  const T p2 = t1 * t1 + t2 * t2 + t3 * t3 + 2 * p1;
  const T p = std::sqrt(p2 / 6);
  rpe::Mat3<T> B;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B(i, j) = (1 / p) * (M_(i, j) - q * (i == j ? T(1) : T(0)));
  const T r = B.determinant() / 2;
  T phi;
  if (r <= -1)
    phi = T(3.141592653589793238) / 3;
  else if (r >= 1)
    phi = 0;
  else
    phi = std::acos(r) / 3;
  (*pE_)(0) = q + 2 * p * std::cos(phi);
  (*pE_)(2) = q + 2 * p * std::cos(phi + (2 * T(3.141592653589793238) / 3));
  (*pE_)(1) = 3 * q - (*pE_)(0) - (*pE_)(2);
}

// Two correspondences with positions (A, B) and normals (N, M) in both frames -> R_cw, t_w (uses A, N, B).
template <typename T>
void ms(const rpe::Vec3<T>& Aw_, const rpe::Vec3<T>& Bw_, const rpe::Vec3<T>& Nw_, const rpe::Vec3<T>& /*Mw_*/,
        const rpe::Vec3<T>& Ac_, const rpe::Vec3<T>& Bc_, const rpe::Vec3<T>& Nc_, const rpe::Vec3<T>& /*Mc_*/,
        rpe::SO3<T>* pR_cw_, rpe::Vec3<T>* pTw_) {
  T q[4], t[3];
  rpe::nl_2p<T>(Ac_.v, Nc_.v, Bc_.v, Aw_.v, Nw_.v, Bw_.v, q, t);
  *pR_cw_ = rpe::SO3<T>::fromRawQuaternion(q);
  *pTw_ = rpe::Vec3<T>(t[0], t[1], t[2]);
}

#endif  // RPE_MINIMAL_SOLVERS_HPP_
