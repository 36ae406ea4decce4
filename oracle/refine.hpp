// oracle/refine.hpp — TEST INFRASTRUCTURE (see oracle/README.md). PARITY UNPINNED.
//
// (1) nl_shinji_kneip_ls + find_opt_cc: the reference's own multi-modal refinement
//     (AbsoluteOrientationNormal.hpp:13-46, 447-552), restated with its quirks.
// (2) refine_gn: the north-star Levenberg-Marquardt refinement on SE(3). The reference has NO
//     such routine (its "refinement" is closed form); this twin DEFINES the algorithm the CUDA
//     kernel implements. Residual definitions follow the reference's own error terms:
//       3-D    r = (R x_w + t) - x_c                       AbsoluteOrientation.hpp:137
//       2-D    r = b x normalize(R x_w + t)  (|r| = sin of the angle)   PnPPoseAdapter.hpp:204-210, P3P.hpp:482-484
//       normal r = R n_w - n_c                             AbsoluteOrientationNormal.hpp:248
//     update T <- exp(delta) * T with Sophus' SE3 exponential (se3.hpp:321-342).
#ifndef ORACLE_REFINE_HPP_
#define ORACLE_REFINE_HPP_

#include <cmath>
#include <vector>

#include "ransac.hpp"

namespace orc {

// ---- find_opt_cc — AbsoluteOrientationNormal.hpp:13-46 --------------------------------------
// flags23: 2-D inlier flags; Rcw: current adapter rotation. Returns false when the reference
// returns the NaN vector (|det(AA)| < 1e-4).
template <class T>
inline bool find_opt_cc(const Corr<T>& d, const short* flags23, const SO3<T>& Rcw, V3<T>* c_w) {
  const M3<T> Rwc = Rcw.inverse().matrix();  // :21
  M3<T> AA;
  V3<T> bb;
  for (int i = 0; i < d.n; i++) {
    if (flags23[i] != 1) continue;  // :26
    const V3<T> vr = Rwc * col3(d.bv, i);
    M3<T> A;
    A(0, 0) = 1 - vr[0] * vr[0];
    A(1, 0) = A(0, 1) = -vr[0] * vr[1];
    A(2, 0) = A(0, 2) = -vr[0] * vr[2];
    A(1, 1) = 1 - vr[1] * vr[1];
    A(2, 1) = A(1, 2) = -vr[1] * vr[2];
    A(2, 2) = 1 - vr[2] * vr[2];
    const V3<T> b = A * col3(d.xw, i);  // :35
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) AA(r, c) = AA(r, c) + A(r, c);
      bb[r] = bb[r] + b[r];
    }
  }
  if (std::fabs(det3(AA)) < T(0.0001)) return false;  // :41-42
  // :44 — JacobiSVD(...).solve(bb): x = V * diag(1/s_i, i < rank) * U^T * bb,
  // rank = #{ s_i > max(s_0 * diagSize * eps, min) } (Eigen SVDBase::rank / _solve_impl)
  const SVD3<T> svd = jacobi_svd3(AA);
  const T thr0 = svd.s[0] * (T(3) * std::numeric_limits<T>::epsilon());
  const T thr = thr0 > (std::numeric_limits<T>::min)() ? thr0 : (std::numeric_limits<T>::min)();
  int rank = 0;
  while (rank < 3 && svd.s[rank] > thr) ++rank;
  V3<T> tmp;
  for (int k = 0; k < rank; ++k) {
    T acc = sum3(svd.U(0, k) * bb[0], svd.U(1, k) * bb[1], svd.U(2, k) * bb[2]);
    tmp[k] = (T(1) / svd.s[k]) * acc;
  }
  V3<T> x;
  for (int r = 0; r < 3; ++r) {
    if (rank == 3)
      x[r] = sum3(svd.V(r, 0) * tmp[0], svd.V(r, 1) * tmp[1], svd.V(r, 2) * tmp[2]);
    else {
      T acc = T(0);
      for (int k = 0; k < rank; ++k) acc += svd.V(r, k) * tmp[k];
      x[r] = acc;
    }
  }
  *c_w = x;
  return true;
}

// ---- nl_shinji_kneip_ls — AbsoluteOrientationNormal.hpp:447-552 -----------------------------
// mask3: n x 3 column-major flags (col 0 = 2-3, col 1 = 3-3, col 2 = N-N). weights3: n x 3
// column-major per-correspondence weights or null (adapters return 1 when no weights are set:
// PnPPoseAdapter.hpp:163, AOPoseAdapter.hpp:163, NormalAOPoseAdapter.hpp:155). Note the
// reference divides the 3-3 and N-N weights by 32767 (AOPoseAdapter.hpp:167, NormalAOPoseAdapter.hpp:159)
// and never resets M23/M33/MNN/K/TW/M/TL between its three passes (:473-477 are outside the loop).
template <class T>
inline void nl_shinji_kneip_ls(const Corr<T>& d, const short* mask3, const T* weights3, int max_votes, SE3<T>* pose) {
  if (max_votes == 0) return;  // :454
  const int n = d.n;
  const short* f23 = mask3;
  const short* f33 = mask3 + n;
  const short* fnn = mask3 + 2 * n;
  auto w23 = [&](int i) { return weights3 ? weights3[i] : T(1.0); };
  auto w33 = [&](int i) { return weights3 ? T(weights3[n + i]) / std::numeric_limits<short>::max() : T(1.0); };
  auto wnn = [&](int i) { return weights3 ? T(weights3[2 * n + i]) / std::numeric_limits<short>::max() : T(1.0); };
  V3<T> Cw, Cc;
  int N = 0;
  T TV = 0;
  for (int i = 0; i < n; i++) {  // :458-465
    if (f33[i] == 1) {
      const T v = w33(i);
      Cw = Cw + v * col3(d.xw, i);
      Cc = Cc + v * col3(d.xc, i);
      TV += v;
      N++;
    }
  }
  if (N > 2) {  // :466-469
    Cw = Cw / TV;
    Cc = Cc / TV;
  }
  M3<T> M33, MNN, M23;
  int M = 0;
  T TL = 0;
  int K = 0;
  T TW = 0;
  const SO3<T> Rcw0 = pose->so3;
  V3<T> c_opt = Rcw0.inverse() * (-pose->t);  // :479
  SO3<T> R_opt;
  auto add_scaled_outer = [](M3<T>& acc, T w, const V3<T>& a, const V3<T>& b) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) acc(r, c) = acc(r, c) + (w * a[r]) * b[c];
  };
  for (int ii = 0; ii < 3; ii++) {  // :481
    T sigma_w_sqr = 0.;
    for (int i = 0; i < n; i++) {
      if (f23[i] == 1) {  // :485-491
        const T w = w23(i);
        V3<T> Aw = col3(d.xw, i) - c_opt;
        normalize(Aw);
        add_scaled_outer(M23, w, col3(d.bv, i), Aw);
        TW += w;
        K++;
      }
      if (f33[i] == 1) {  // :492-497
        const T v = w33(i);
        const V3<T> Aw = col3(d.xw, i) - Cw;
        const V3<T> Ac = col3(d.xc, i) - Cc;
        sigma_w_sqr += (v * squared_norm(Ac));
        add_scaled_outer(M33, v, Ac, Aw);
      }
      if (fnn[i] == 1) {  // :498-504
        const T lambda = wnn(i);
        add_scaled_outer(MNN, lambda, col3(d.nc, i), col3(d.nw, i));
        TL += lambda;
        M++;
      }
    }
    auto scale = [](M3<T>& a, T s) {
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) a(r, c) = a(r, c) / s;
    };
    if (N > 2) {  // :506
      scale(M33, TV);
      sigma_w_sqr /= TV;
    } else {
      M33 = M3<T>();
      sigma_w_sqr = 1.;
    }
    if (M > 0) scale(MNN, TL); else MNN = M3<T>();  // :507
    if (K > 0) scale(M23, TW); else M23 = M3<T>();  // :508
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) M33(r, c) = M33(r, c) + sigma_w_sqr * (M23(r, c) + MNN(r, c));  // :510
    const SVD3<T> svd = jacobi_svd3(M33);  // :512
    const M3<T> Vt = transpose(svd.V);
    const M3<T> TMP = svd.U * Vt;
    if (det3(TMP) < 0) {  // :522-525
      M3<T> I = M3<T>::identity();
      I(2, 2) = T(-1);
      R_opt = SO3<T>::from_matrix((svd.U * I) * Vt);
    } else {
      R_opt = SO3<T>::from_matrix(TMP);
    }
    const V3<T> c = Cw - R_opt.inverse() * Cc;  // :530
    V3<T> cp;
    const bool cp_ok = find_opt_cc(d, f23, Rcw0, &cp);  // :531 (adapter rotation is not updated inside the loop)
    if (N > 2) {                                          // :532-537
      if (cp_ok)
        c_opt = (T(K) / (K + N)) * cp + (T(N) / (K + N)) * c;
      else
        c_opt = c;
    } else {  // :538-543
      if (cp_ok)
        c_opt = cp;
      else
        break;
    }
  }
  pose->so3 = R_opt;           // :545
  pose->t = R_opt * (-c_opt);  // :546
}

// ---- refine_gn: LM on SE(3), binary64 throughout (the twin of the CUDA refine kernel) ----------
struct GnAccum {
  double H[21];  // upper triangle of J^T J, row-major (00,01,..,05,11,..)
  double g[6];   // J^T r
  double cost;
  long long rows;
};

inline void gn_add_row(GnAccum& a, const double J[6], double r, double w) {
  int k = 0;
  for (int i = 0; i < 6; ++i) {
    for (int j = i; j < 6; ++j) a.H[k++] += w * J[i] * J[j];
    a.g[i] += w * J[i] * r;
  }
  a.cost += w * r * r;
}

// Jacobian rows of y = R x + t under T <- exp([v,w]) T:  dy = v + w x y  = [ I | -[y]x ] (v,w)
inline void gn_point_rows(const double y[3], double J[3][6]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 6; ++c) J[r][c] = 0.0;
  J[0][0] = J[1][1] = J[2][2] = 1.0;
  J[0][4] = y[2];  J[0][5] = -y[1];
  J[1][3] = -y[2]; J[1][5] = y[0];
  J[2][3] = y[1];  J[2][4] = -y[0];
}

template <class T>
inline void gn_evaluate(const Corr<T>& d, const short* mask, int mask_cols, double w2d, double w3d, double wnl,
                        const double R[9], const double t[3], GnAccum* acc) {
  GnAccum a;
  for (int i = 0; i < 21; ++i) a.H[i] = 0;
  for (int i = 0; i < 6; ++i) a.g[i] = 0;
  a.cost = 0;
  a.rows = 0;
  const int n = d.n;
  const short* f2 = (mask_cols >= 1 && d.bv) ? mask : 0;
  const short* f3 = (mask_cols >= 2 && d.xc) ? mask + n : 0;
  const short* fn = (mask_cols >= 3 && d.nc) ? mask + 2 * n : 0;
  for (int i = 0; i < n; ++i) {
    const bool u2 = f2 && f2[i] == 1, u3 = f3 && f3[i] == 1, un = fn && fn[i] == 1;
    if (!(u2 || u3 || un)) continue;
    double x[3] = {(double)d.xw[3 * i], (double)d.xw[3 * i + 1], (double)d.xw[3 * i + 2]};
    double y[3];
    for (int r = 0; r < 3; ++r) y[r] = R[3 * r] * x[0] + R[3 * r + 1] * x[1] + R[3 * r + 2] * x[2] + t[r];
    double Jy[3][6];
    gn_point_rows(y, Jy);
    if (u3 && w3d > 0) {
      for (int r = 0; r < 3; ++r) {
        gn_add_row(a, Jy[r], y[r] - (double)d.xc[3 * i + r], w3d);
        a.rows++;
      }
    }
    if (u2 && w2d > 0) {
      const double b[3] = {(double)d.bv[3 * i], (double)d.bv[3 * i + 1], (double)d.bv[3 * i + 2]};
      const double ny = std::sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
      const double u[3] = {y[0] / ny, y[1] / ny, y[2] / ny};
      // du = (I - u u^T)/|y| dy ; r = b x u ; dr = [b]x du
      double P[3][3];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) P[r][c] = ((r == c ? 1.0 : 0.0) - u[r] * u[c]) / ny;
      double Ju[3][6];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 6; ++c) Ju[r][c] = P[r][0] * Jy[0][c] + P[r][1] * Jy[1][c] + P[r][2] * Jy[2][c];
      const double res[3] = {b[1] * u[2] - b[2] * u[1], b[2] * u[0] - b[0] * u[2], b[0] * u[1] - b[1] * u[0]};
      double Jr[3][6];
      for (int c = 0; c < 6; ++c) {
        Jr[0][c] = b[1] * Ju[2][c] - b[2] * Ju[1][c];
        Jr[1][c] = b[2] * Ju[0][c] - b[0] * Ju[2][c];
        Jr[2][c] = b[0] * Ju[1][c] - b[1] * Ju[0][c];
      }
      for (int r = 0; r < 3; ++r) {
        gn_add_row(a, Jr[r], res[r], w2d);
        a.rows++;
      }
    }
    if (un && wnl > 0) {
      const double nw[3] = {(double)d.nw[3 * i], (double)d.nw[3 * i + 1], (double)d.nw[3 * i + 2]};
      double m[3];
      for (int r = 0; r < 3; ++r) m[r] = R[3 * r] * nw[0] + R[3 * r + 1] * nw[1] + R[3 * r + 2] * nw[2];
      double Jn[3][6];
      gn_point_rows(m, Jn);
      for (int r = 0; r < 3; ++r) {
        Jn[r][0] = Jn[r][1] = Jn[r][2] = 0.0;  // normals do not translate
        gn_add_row(a, Jn[r], m[r] - (double)d.nc[3 * i + r], wnl);
        a.rows++;
      }
    }
  }
  *acc = a;
}

// Solve (H + mu diag(H)) delta = -g by Cholesky. Returns false if not positive definite.
inline bool gn_solve(const double Hu[21], const double g[6], double mu, double delta[6]) {
  double A[6][6];
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) {
      A[i][j] = A[j][i] = Hu[k++];
    }
  for (int i = 0; i < 6; ++i) A[i][i] += mu * A[i][i];
  double L[6][6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) L[i][j] = 0.0;
  for (int j = 0; j < 6; ++j) {
    double s = A[j][j];
    for (int p = 0; p < j; ++p) s -= L[j][p] * L[j][p];
    if (!(s > 0.0)) return false;
    L[j][j] = std::sqrt(s);
    for (int i = j + 1; i < 6; ++i) {
      double v = A[i][j];
      for (int p = 0; p < j; ++p) v -= L[i][p] * L[j][p];
      L[i][j] = v / L[j][j];
    }
  }
  double z[6];
  for (int i = 0; i < 6; ++i) {
    double v = -g[i];
    for (int p = 0; p < i; ++p) v -= L[i][p] * z[p];
    z[i] = v / L[i][i];
  }
  for (int i = 5; i >= 0; --i) {
    double v = z[i];
    for (int p = i + 1; p < 6; ++p) v -= L[p][i] * delta[p];
    delta[i] = v / L[i][i];
  }
  return true;
}

// T <- exp(delta) T, delta = (v, w). SO3 exp by Rodrigues, V-matrix as in se3.hpp:321-342.
inline void se3_exp_left(const double delta[6], double R[9], double t[3]) {
  const double v[3] = {delta[0], delta[1], delta[2]};
  const double w[3] = {delta[3], delta[4], delta[5]};
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = std::sqrt(th2);
  double A, B, C;  // sin(th)/th, (1-cos th)/th^2, (th - sin th)/th^3
  if (th < 1e-6) {
    A = 1.0 - th2 / 6.0;
    B = 0.5 - th2 / 24.0;
    C = 1.0 / 6.0 - th2 / 120.0;
  } else {
    A = std::sin(th) / th;
    B = (1.0 - std::cos(th)) / th2;
    C = (th - std::sin(th)) / (th2 * th);
  }
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double W2[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) W2[3 * r + c] = W[3 * r] * W[c] + W[3 * r + 1] * W[3 + c] + W[3 * r + 2] * W[6 + c];
  double Rd[9], V[9];
  for (int i = 0; i < 9; ++i) {
    const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    Rd[i] = I + A * W[i] + B * W2[i];
    V[i] = I + B * W[i] + C * W2[i];
  }
  double Rn[9], tn[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Rn[3 * r + c] = Rd[3 * r] * R[c] + Rd[3 * r + 1] * R[3 + c] + Rd[3 * r + 2] * R[6 + c];
    tn[r] = Rd[3 * r] * t[0] + Rd[3 * r + 1] * t[1] + Rd[3 * r + 2] * t[2] + V[3 * r] * v[0] + V[3 * r + 1] * v[1] +
            V[3 * r + 2] * v[2];
  }
  for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  for (int i = 0; i < 3; ++i) t[i] = tn[i];
}

// LM schedule (shared verbatim with the device tail): mu0 = 1e-4; accept if cost decreased
// (first evaluation always accepted) -> mu *= 0.1 (floor 1e-12); reject -> mu *= 10 and re-solve
// from the last accepted normal equations; stop when max|delta| < 1e-10 or max_iters evaluations.
// info[0] = final cost, info[1] = evaluations, info[2] = accepted steps, info[3] = final mu.
template <class T>
inline int refine_gn(const Corr<T>& d, const short* mask, int mask_cols, T w2d, T w3d, T wnl, int max_iters, SE3<T>* pose,
                     double* info) {
  // start: normalised quaternion -> matrix, in binary64
  double q[4] = {(double)pose->so3.q.x, (double)pose->so3.q.y, (double)pose->so3.q.z, (double)pose->so3.q.w};
  const double qn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= qn;
  Quat<double> qd(q[3], q[0], q[1], q[2]);
  const M3<double> R0 = quat_to_matrix(qd);
  double Rp[9], tp[3];  // proposal
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Rp[3 * r + c] = R0(r, c);
  for (int r = 0; r < 3; ++r) tp[r] = (double)pose->t[r];
  double Ra[9] = {0}, ta[3] = {0};  // accepted (read only once `have` is set)
  GnAccum acc_a;
  double mu = 1e-4;
  int evals = 0, accepted = 0;
  bool have = false;
  for (int it = 0; it < max_iters; ++it) {
    GnAccum acc;
    gn_evaluate(d, mask, mask_cols, (double)w2d, (double)w3d, (double)wnl, Rp, tp, &acc);
    ++evals;
    if (!have || acc.cost < acc_a.cost) {
      for (int i = 0; i < 9; ++i) Ra[i] = Rp[i];
      for (int i = 0; i < 3; ++i) ta[i] = tp[i];
      acc_a = acc;
      if (have) {
        mu = mu * 0.1;
        if (mu < 1e-12) mu = 1e-12;
        ++accepted;
      }
      have = true;
    } else {
      mu = mu * 10.0;
    }
    if (acc_a.rows < 6) break;
    double delta[6];
    int tries = 0;
    while (!gn_solve(acc_a.H, acc_a.g, mu, delta) && tries < 8) {
      mu = mu * 10.0;
      ++tries;
    }
    if (tries == 8) break;
    double mx = 0;
    for (int i = 0; i < 6; ++i) mx = std::fabs(delta[i]) > mx ? std::fabs(delta[i]) : mx;
    if (mx < 1e-10) break;
    for (int i = 0; i < 9; ++i) Rp[i] = Ra[i];
    for (int i = 0; i < 3; ++i) tp[i] = ta[i];
    se3_exp_left(delta, Rp, tp);
  }
  if (have) {
    M3<double> Rm;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rm(r, c) = Ra[3 * r + c];
    const Quat<double> qo = quat_from_matrix(Rm);
    const double nn = std::sqrt(quat_squared_norm(qo));
    pose->so3.q.x = (T)(qo.x / nn);
    pose->so3.q.y = (T)(qo.y / nn);
    pose->so3.q.z = (T)(qo.z / nn);
    pose->so3.q.w = (T)(qo.w / nn);
    pose->t = V3<T>((T)ta[0], (T)ta[1], (T)ta[2]);
    if (info) {
      info[0] = acc_a.cost;
      info[1] = evals;
      info[2] = accepted;
      info[3] = mu;
    }
  }
  return evals;
}

}  // namespace orc

#endif  // ORACLE_REFINE_HPP_
