// rpe/MinimalSolvers.hpp — mirrors /root/reference/pose/MinimalSolvers.hpp.
//
// The reference file is never called and half-finished: `ms` stops after step 4 of a 2-point+normal solver and
// mutates const references (:22-46); the eigenvector half of `ev` does not compile (:92). Built here is their
// intent: `ev` = closed-form eigenvalues of a symmetric 3x3 (:49-83), `ms` = the working point+normal minimal
// solver (nl_2p, AbsoluteOrientationNormal.hpp:77-142) applied to the first correspondence pair. Both are host +
// device templates (rpe/solvers_min.h); batches run on the GPU one problem per thread through rpe_min_ev / rpe_min_ms.
#ifndef RPE_MINIMAL_SOLVERS_HPP_
#define RPE_MINIMAL_SOLVERS_HPP_

#include <cmath>

#include "so3.hpp"
#include "solvers_min.h"

// eigenvalues, eig(0) >= eig(1) >= eig(2); the arithmetic lives in solvers_min.h and is the one the GPU runs
// (rpe_min_ev), so host and device agree bit for bit
template <class T>
void ev(const rpe::Mat3<T>& M_, rpe::Vec3<T>* pE_) {
  T m[9], e[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) m[3 * i + j] = M_(i, j);
  rpe::sym3_eigenvalues<T>(m, e);
  for (int i = 0; i < 3; ++i) (*pE_)(i) = e[i];
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
