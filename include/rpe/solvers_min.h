// rpe/solvers_min.h — the two routines of /root/reference/pose/MinimalSolvers.hpp as host + device templates
// (one problem per thread on the GPU: minsolv_*_kernel in csrc/pipeline.cu, C-ABI rpe_min_ev / rpe_min_ms).
//
//   sym3_eigenvalues   ev()  MinimalSolvers.hpp:49-83   closed-form eigenvalues of a symmetric 3x3 (trigonometric
//                                                       method), eig[0] >= eig[1] >= eig[2]; operation order of the
//                                                       reference incl. Eigen's 3x3 determinant (mat_det); acos / cos
//                                                       through the bit-reproducible helpers of det_math.h, so host and
//                                                       device return identical bits
//   min_solver_2pn     ms()  MinimalSolvers.hpp:10-46   the reference stops after step 4 of a 2-point + normal solver and
//                                                       does not compile as written (it assigns to const references);
//                                                       its finished form in the same repository is nl_2p
//                                                       (AbsoluteOrientationNormal.hpp:77-142), applied here to (A, N_A, B)
// Same compile-flag contract as solvers.h (-fmad=false / -ffp-contract=off).
#ifndef RPE_SOLVERS_MIN_H_
#define RPE_SOLVERS_MIN_H_

#include "solvers.h"
#include "solvers_p3p.h"

namespace rpe {

// M row-major 3x3 (only its symmetric part is read the way the reference reads it: (0,1), (0,2), (1,2) + diagonal for
// p1/q, the full matrix for B)
template <class T>
RPE_FN void sym3_eigenvalues(const T* M, T* eig) {
  const T p1 = M[1] * M[1] + M[2] * M[2] + M[5] * M[5];  // :53
  if (t_abs(p1) < T(0.00001)) {                           // :54 A is diagonal
    eig[0] = M[0];
    eig[1] = M[4];
    eig[2] = M[8];
    return;
  }
  T q = M[0] + M[4] + M[8];  // :61
  q /= T(3);
  const T t1 = M[0] - q, t2 = M[4] - q, t3 = M[8] - q;
  const T p2 = t1 * t1 + t2 * t2 + t3 * t3 + T(2) * p1;  // :65
  const T p = t_sqrt(p2 / T(6));
  T B[9];  // :67 (1/p) * (M - q I)
  const T ip = T(1) / p;
RPE_UNROLL
  for (int i = 0; i < 9; ++i) B[i] = ip * (M[i] - q * ((i == 0 || i == 4 || i == 8) ? T(1) : T(0)));
  const T r = mat_det(B) / T(2);  // :68
  T phi;
  if (r <= T(-1))
    phi = T(3.14159265358979323846) / T(3);
  else if (r >= T(1))
    phi = T(0);
  else
    phi = det::acos_t(r) / T(3);
  T s, c;
  det::sincos_t(phi, &s, &c);
  eig[0] = q + T(2) * p * c;  // :80
  T s2, c2;
  det::sincos_t(phi + T(2.0 * 3.14159265358979323846 / 3.0), &s2, &c2);
  eig[2] = q + T(2) * p * c2;
  eig[1] = T(3) * q - eig[0] - eig[2];  // trace
}

// in: Aw Bw Nw Mw Ac Bc Nc Mc (3 values each, the argument order of ms()); out: R_cw as quaternion x,y,z,w and t_w
template <class T>
RPE_FN void min_solver_2pn(const T* in24, T* q, T* t) {
  const T *Aw = in24, *Bw = in24 + 3, *Nw = in24 + 6, *Ac = in24 + 12, *Bc = in24 + 15, *Nc = in24 + 18;
  nl_2p<T>(Ac, Nc, Bc, Aw, Nw, Bw, q, t);
}

}  // namespace rpe

#endif  // RPE_SOLVERS_MIN_H_
