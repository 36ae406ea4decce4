// oracle/solvers.hpp — TEST INFRASTRUCTURE (see oracle/README.md). PARITY UNPINNED.
//
// CPU restatement of the reference's minimal solvers and of its adaptive stopping rule.
// Every routine cites the /root/reference lines it follows. `Tp` is the reference's scalar
// template parameter (float in SimpleMain.cpp:19 / Library.cpp, double in TestMain.cpp:163).
//
// Math mode (orc::g_math_mode): 0 = libm/libstdc++ exactly where the reference calls them;
// 1 = rpe::det:: helpers (include/rpe/det_math.h) at the same call sites, which is what the
// CUDA product computes with, so DET-mode results are bit-comparable with the GPU.
#ifndef ORACLE_SOLVERS_HPP_
#define ORACLE_SOLVERS_HPP_

#include <complex>
#include <limits>
#include <vector>

#include "../include/rpe/det_math.h"
#include "sophus_model.hpp"

namespace orc {

extern int g_math_mode;

template <class T>
inline T m_acos(T x) {
  return g_math_mode ? rpe::det::acos_t(x) : std::acos(x);
}
template <class T>
inline T m_log(T x) {
  return g_math_mode ? rpe::det::log_t(x) : std::log(x);
}
template <class T>
inline void m_sincos(T a, T* s, T* c) {
  if (g_math_mode) {
    rpe::det::sincos_t(a, s, c);
  } else {
    *s = std::sin(a);
    *c = std::cos(a);
  }
}

// Column accessor for a column-major 3 x n array (Eigen::Matrix<Tp,Dynamic,Dynamic>::col).
template <class T>
inline V3<T> col3(const T* a, int i) {
  return V3<T>(a[3 * i + 0], a[3 * i + 1], a[3 * i + 2]);
}

// ------------------------------------------------------------------------------------------
// shinji — AbsoluteOrientation.hpp:47-99. Xw, Xc: column-major 3 x cols; first K columns used;
// `cols` is X_w_.cols() (the divisor at :75 — a pure scale, kept for faithfulness).
// ------------------------------------------------------------------------------------------
template <class T>
inline SE3<T> shinji(const T* Xw, const T* Xc, int K, int cols) {
  V3<T> Cw, Cc;  // :55
  for (int n = 0; n < K; ++n) {
    Cw = Cw + col3(Xw, n);  // :57
    Cc = Cc + col3(Xc, n);  // :58
  }
  Cw = Cw / (T)K;  // :60
  Cc = Cc / (T)K;  // :61
  M3<T> M;         // :65
  for (int n = 0; n < K; ++n) {
    const V3<T> Aw = col3(Xw, n) - Cw;  // :69
    const V3<T> Ac = col3(Xc, n) - Cc;  // :70
    const M3<T> N = outer(Ac, Aw);      // :71
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) M(i, j) = M(i, j) + N(i, j);  // :72
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M(i, j) = M(i, j) / (T)cols;  // :75
  const SVD3<T> svd = jacobi_svd3(M);                          // :79
  const M3<T> Vt = transpose(svd.V);
  const M3<T> Tmp = svd.U * Vt;  // :85
  const T d = det3(Tmp);         // :86
  SO3<T> R;
  if (d < T(0)) {  // :88-91  U*I*V^T, I = diag(1,1,-1)
    M3<T> I = M3<T>::identity();
    I(2, 2) = T(-1);
    R = SO3<T>::from_matrix((svd.U * I) * Vt);
  } else {
    R = SO3<T>::from_matrix(Tmp);  // :93 (U*V^T evaluated again; same arithmetic)
  }
  const V3<T> t = Cc - R * Cw;  // :95
  return SE3<T>(R, t);
}

// ------------------------------------------------------------------------------------------
// RANSACUpdateNumIters — P3P.hpp:296-318. Note the mixed precision the C++ typing produces:
// "1. - p" and std::pow(T, int) are evaluated in double, and the final rounding uses a
// float literal 0.5f.
// ------------------------------------------------------------------------------------------
template <class T>
inline int ransac_update_num_iters(T p, T ep, const int modelPoints, const int maxIters) {
  p = std::max(p, T(0.));   // :299
  p = std::min(p, T(1.));   // :300
  ep = std::max(ep, T(0.));  // :301
  ep = std::min(ep, T(1.));  // :302
  T num = std::max(T(1. - p), std::numeric_limits<T>::epsilon());  // :305
  double pw;
  if (g_math_mode) {
    const double base = (double)T(1. - ep);
    pw = 1.0;
    for (int i = 0; i < modelPoints; ++i) pw = pw * base;
  } else {
    pw = std::pow((double)T(1. - ep), (double)modelPoints);  // std::pow(T,int) promotes to double
  }
  T denom = (T)(1.0 - pw);                                   // :306
  if (denom < std::numeric_limits<T>::epsilon()) return 0;  // :307-308
  num = m_log(num);                                          // :310
  denom = m_log(denom);                                      // :311
  return denom >= 0 || -num >= maxIters * (-denom) ? maxIters : int(num / denom + 0.5f);  // :317
}

// ------------------------------------------------------------------------------------------
// o4_roots — P3P.hpp:11-60 (Ferrari closed form through std::complex<Tp>).
// ------------------------------------------------------------------------------------------
template <class T>
struct Cx {
  T re, im;
};
template <class T>
inline Cx<T> cx_mul(Cx<T> a, Cx<T> b) {
  return Cx<T>{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <class T>
inline Cx<T> cx_div(Cx<T> a, Cx<T> b) {  // Smith's algorithm
  if (std::abs(b.re) < std::abs(b.im)) {
    const T ratio = b.re / b.im;
    const T den = (b.re * ratio) + b.im;
    return Cx<T>{((a.re * ratio) + a.im) / den, ((a.im * ratio) - a.re) / den};
  }
  const T ratio = b.im / b.re;
  const T den = (b.im * ratio) + b.re;
  return Cx<T>{((a.im * ratio) + a.re) / den, (a.im - (a.re * ratio)) / den};
}
template <class T>
inline Cx<T> cx_sqrt(Cx<T> z) {  // principal branch; imag -0 treated as +0
  if (z.re == T(0) && z.im == T(0)) return Cx<T>{T(0), T(0)};
  const T h = std::sqrt(z.re * z.re + z.im * z.im);
  const T t = std::sqrt((std::abs(z.re) + h) * T(0.5));
  if (z.re >= T(0)) return Cx<T>{t, z.im / (T(2) * t)};
  return Cx<T>{std::abs(z.im) / (T(2) * t), z.im < T(0) ? -t : t};
}
template <class T>
inline Cx<T> cx_cbrt(Cx<T> z) {  // principal cube root via polar form, det helpers
  if (z.re == T(0) && z.im == T(0)) return Cx<T>{T(0), T(0)};
  const T h = std::sqrt(z.re * z.re + z.im * z.im);
  const T mag = rpe::det::cbrt_t(h);
  const T th = rpe::det::atan2_t(z.im, z.re) / T(3);
  T s, c;
  rpe::det::sincos_t(th, &s, &c);
  return Cx<T>{mag * c, mag * s};
}

template <class T>
inline void o4_roots(const T f[5], T roots[4]) {
  const T A = f[0], B = f[1], C = f[2], D = f[3], E = f[4];  // :14-18
  const T A_pw2 = A * A, B_pw2 = B * B;
  const T A_pw3 = A_pw2 * A, B_pw3 = B_pw2 * B;
  const T A_pw4 = A_pw3 * A, B_pw4 = B_pw3 * B;
  const T alpha = -3 * B_pw2 / (8 * A_pw2) + C / A;                                                     // :27
  const T beta = B_pw3 / (8 * A_pw3) - B * C / (2 * A_pw2) + D / A;                                     // :28
  const T gamma = -3 * B_pw4 / (256 * A_pw4) + B_pw2 * C / (16 * A_pw3) - B * D / (4 * A_pw2) + E / A;  // :29
  const T alpha_pw2 = alpha * alpha;
  const T alpha_pw3 = alpha_pw2 * alpha;
  const T Pre = -alpha_pw2 / 12 - gamma;  // :34
  // :35 — pow(beta, 2) is std::pow(T,int) -> double, so the last subtraction is done in double
  const T Qre = (T)((-alpha_pw3 / 108 + alpha * gamma / 3) - ((double)beta * (double)beta) / 8);
  const T b4a = -B / (T(4.) * A);
  if (!g_math_mode) {
    typedef std::complex<T> C_;
    const C_ P(Pre, 0), Q(Qre, 0);
    const C_ R = -Q / T(2.0) + std::sqrt(std::pow(Q, T(2.)) / T(4.) + std::pow(P, T(3.)) / T(27.));  // :36
    const C_ U = std::pow(R, T(1.0 / 3.0));                                                             // :38
    C_ y;
    if (U.real() == 0)
      y = -T(5.0) * alpha / T(6.) - std::pow(Q, T(1.0 / 3.0));  // :42
    else
      y = -T(5.0) * alpha / T(6.) - P / (T(3.) * U) + U;  // :44
    const C_ w = std::sqrt(alpha + T(2.) * y);             // :46
    C_ temp;
    temp = b4a + T(0.5) * (w + std::sqrt(-(T(3.) * alpha + T(2.) * y + T(2.) * beta / w)));  // :50
    roots[0] = temp.real();
    temp = b4a + T(0.5) * (w - std::sqrt(-(T(3.) * alpha + T(2.) * y + T(2.) * beta / w)));  // :52
    roots[1] = temp.real();
    temp = b4a + T(0.5) * (-w + std::sqrt(-(T(3.) * alpha + T(2.) * y - T(2.) * beta / w)));  // :54
    roots[2] = temp.real();
    temp = b4a + T(0.5) * (-w - std::sqrt(-(T(3.) * alpha + T(2.) * y - T(2.) * beta / w)));  // :56
    roots[3] = temp.real();
    return;
  }
  // DET mode: the same formula with explicit complex arithmetic (exact real powers of the
  // real P, Q instead of polar-form pow, principal branches, imaginary -0 == +0).
  const Cx<T> P{Pre, T(0)}, Q{Qre, T(0)};
  const Cx<T> Q2 = cx_mul(Q, Q);
  const Cx<T> P3 = cx_mul(cx_mul(P, P), P);
  const Cx<T> rad{Q2.re / T(4.) + P3.re / T(27.), Q2.im / T(4.) + P3.im / T(27.)};
  const Cx<T> sq = cx_sqrt(rad);
  const Cx<T> R{-Q.re / T(2.0) + sq.re, -Q.im / T(2.0) + sq.im};
  const Cx<T> U = cx_cbrt(R);
  const T m56a = -T(5.0) * alpha / T(6.);
  Cx<T> y;
  if (U.re == 0) {
    const Cx<T> cq = cx_cbrt(Q);
    y = Cx<T>{m56a - cq.re, -cq.im};
  } else {
    const Cx<T> pu = cx_div(P, Cx<T>{T(3.) * U.re, T(3.) * U.im});
    y = Cx<T>{(m56a - pu.re) + U.re, (-pu.im) + U.im};
  }
  const Cx<T> w = cx_sqrt(Cx<T>{alpha + T(2.) * y.re, T(2.) * y.im});
  const Cx<T> bw = cx_div(Cx<T>{T(2.) * beta, T(0)}, w);
  const T a3 = T(3.) * alpha;
  const Cx<T> base{a3 + T(2.) * y.re, T(2.) * y.im};
  const Cx<T> s1 = cx_sqrt(Cx<T>{-(base.re + bw.re), -(base.im + bw.im)});
  const Cx<T> s2 = cx_sqrt(Cx<T>{-(base.re - bw.re), -(base.im - bw.im)});
  roots[0] = b4a + T(0.5) * (w.re + s1.re);
  roots[1] = b4a + T(0.5) * (w.re - s1.re);
  roots[2] = b4a + T(0.5) * (-w.re + s2.re);
  roots[3] = b4a + T(0.5) * (-w.re - s2.re);
}

// ------------------------------------------------------------------------------------------
// kneip_main — P3P.hpp:63-232. Xw, bv: column-major 3 x (>=3). Up to 4 solutions, in root order.
// A solution whose SO3(R) constructor would abort (so3.hpp:561-566) is DROPPED (the reference
// process would die there); that is the one deliberate deviation, mirrored by the product.
// ------------------------------------------------------------------------------------------
template <class T>
inline int kneip_main(const T* Xw, const T* bv, SE3<T> sol[4]) {
  V3<T> P1 = col3(Xw, 0), P2 = col3(Xw, 1), P3 = col3(Xw, 2);  // :68-70
  const V3<T> temp1 = P2 - P1, temp2 = P3 - P1;                // :72-73
  if (norm(cross(temp1, temp2)) == 0) return 0;                // :75
  V3<T> f1 = col3(bv, 0), f2 = col3(bv, 1), f3 = col3(bv, 2);  // :78-80
  V3<T> e1 = f1;
  V3<T> e3 = cross(f1, f2);
  e3 = e3 / norm(e3);  // :84
  V3<T> e2 = cross(e3, e1);
  M3<T> RR;
  auto set_rows = [](M3<T>& m, const V3<T>& a, const V3<T>& b, const V3<T>& c) {
    for (int j = 0; j < 3; ++j) {
      m(0, j) = a[j];
      m(1, j) = b[j];
      m(2, j) = c[j];
    }
  };
  set_rows(RR, e1, e2, e3);  // :88-90
  f3 = RR * f3;              // :92
  if (f3[2] > 0) {           // :94-114
    f1 = col3(bv, 1);
    f2 = col3(bv, 0);
    f3 = col3(bv, 2);
    e1 = f1;
    e3 = cross(f1, f2);
    e3 = e3 / norm(e3);
    e2 = cross(e3, e1);
    set_rows(RR, e1, e2, e3);
    f3 = RR * f3;
    P1 = col3(Xw, 1);
    P2 = col3(Xw, 0);
    P3 = col3(Xw, 2);
  }
  V3<T> n1 = P2 - P1;  // :116
  n1 = n1 / norm(n1);
  V3<T> n3 = cross(n1, P3 - P1);
  n3 = n3 / norm(n3);
  const V3<T> n2 = cross(n3, n1);
  M3<T> N;
  set_rows(N, n1, n2, n3);  // :123-125
  P3 = N * (P3 - P1);       // :127
  const T d_12 = norm(temp1);  // :129 (temp1 is NOT recomputed after the swap; |P2-P1| is symmetric)
  const T f_1 = f3[0] / f3[2];
  const T f_2 = f3[1] / f3[2];
  const T p_1 = P3[0];
  const T p_2 = P3[1];
  const T cos_beta = dot(f1, f2);  // :135
  // :136 — pow(cos_beta, 2) is double, so the whole right-hand side is double, narrowed on assignment
  T b = (T)(1 / (1 - (double)cos_beta * (double)cos_beta) - 1);
  if (cos_beta < 0)
    b = -std::sqrt(b);
  else
    b = std::sqrt(b);
  const T f_1_pw2 = f_1 * f_1;  // pow(x,2) in double narrowed to T == correctly rounded x*x
  const T f_2_pw2 = f_2 * f_2;
  const T p_1_pw2 = p_1 * p_1;
  const T p_1_pw3 = p_1_pw2 * p_1;
  const T p_1_pw4 = p_1_pw3 * p_1;
  const T p_2_pw2 = p_2 * p_2;
  const T p_2_pw3 = p_2_pw2 * p_2;
  const T p_2_pw4 = p_2_pw3 * p_2;
  const T d_12_pw2 = d_12 * d_12;
  const T b_pw2 = b * b;
  T factors[5];
  factors[0] = -f_2_pw2 * p_2_pw4 - p_2_pw4 * f_1_pw2 - p_2_pw4;  // :156-158
  factors[1] = 2 * p_2_pw3 * d_12 * b + 2 * f_2_pw2 * p_2_pw3 * d_12 * b - 2 * f_2 * p_2_pw3 * f_1 * d_12;  // :160-162
  factors[2] = -f_2_pw2 * p_2_pw2 * p_1_pw2 - f_2_pw2 * p_2_pw2 * d_12_pw2 * b_pw2 - f_2_pw2 * p_2_pw2 * d_12_pw2 +
               f_2_pw2 * p_2_pw4 + p_2_pw4 * f_1_pw2 + 2 * p_1 * p_2_pw2 * d_12 +
               2 * f_1 * f_2 * p_1 * p_2_pw2 * d_12 * b - p_2_pw2 * p_1_pw2 * f_1_pw2 +
               2 * p_1 * p_2_pw2 * f_2_pw2 * d_12 - p_2_pw2 * d_12_pw2 * b_pw2 - 2 * p_1_pw2 * p_2_pw2;  // :164-174
  factors[3] = 2 * p_1_pw2 * p_2 * d_12 * b + 2 * f_2 * p_2_pw3 * f_1 * d_12 - 2 * f_2_pw2 * p_2_pw3 * d_12 * b -
               2 * p_1 * p_2 * d_12_pw2 * b;  // :176-179
  factors[4] = -2 * f_2 * p_2_pw2 * f_1 * p_1 * d_12 * b + f_2_pw2 * p_2_pw2 * d_12_pw2 + 2 * p_1_pw3 * d_12 -
               p_1_pw2 * d_12_pw2 + f_2_pw2 * p_2_pw2 * p_1_pw2 - p_1_pw4 - 2 * f_2_pw2 * p_2_pw2 * p_1 * d_12 +
               p_2_pw2 * f_1_pw2 * p_1_pw2 + f_2_pw2 * p_2_pw2 * d_12_pw2 * b_pw2;  // :181-189
  T realRoots[4];
  o4_roots(factors, realRoots);  // :191
  int ns = 0;
  const M3<T> RRt = transpose(RR), Nt = transpose(N);
  for (int i = 0; i < 4; i++) {  // :193
    const T root = realRoots[i];
    if (root != root) continue;  // :195
    const T cot_alpha = (-f_1 * p_1 / f_2 - root * p_2 + d_12 * b) / (-f_1 * root * p_2 / f_2 + p_1 - d_12);  // :196-198
    const T cos_theta = root;
    if (cos_theta > T(1) || cos_theta < T(-1)) continue;  // :200
    const T sin_theta = std::sqrt(1 - root * root);       // :201
    const T sin_alpha = std::sqrt(1 / (cot_alpha * cot_alpha + 1));
    T cos_alpha = std::sqrt(1 - sin_alpha * sin_alpha);
    if (cot_alpha < 0) cos_alpha = -cos_alpha;  // :205-206
    V3<T> C;
    C[0] = d_12 * cos_alpha * (sin_alpha * b + cos_alpha);              // :209
    C[1] = cos_theta * d_12 * sin_alpha * (sin_alpha * b + cos_alpha);  // :210
    C[2] = sin_theta * d_12 * sin_alpha * (sin_alpha * b + cos_alpha);  // :211
    C = P1 + Nt * C;                                                    // :213
    M3<T> R;
    R(0, 0) = -cos_alpha;
    R(0, 1) = -sin_alpha * cos_theta;
    R(0, 2) = -sin_alpha * sin_theta;
    R(1, 0) = sin_alpha;
    R(1, 1) = -cos_alpha * cos_theta;
    R(1, 2) = -cos_alpha * sin_theta;
    R(2, 0) = T(0.0);
    R(2, 1) = -sin_theta;
    R(2, 2) = cos_theta;
    R = (RRt * R) * N;               // :227
    if (R(0, 0) != R(0, 0)) continue;  // :228
    const SO3<T> so3 = SO3<T>::from_matrix(R);  // :230
    if (!so3.ok) continue;                      // reference: std::abort()
    M3<T> negR;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) negR(r, c) = -R(r, c);
    sol[ns] = SE3<T>(so3, negR * C);  // :230  (-R*C)
    ++ns;
  }
  return ns;
}

// kneip(X_w, bv, *sol) — P3P.hpp:250-294: disambiguate with the 4th column (matrix form of R).
template <class T>
inline bool kneip4(const T* Xw, const T* bv, SE3<T>* out) {
  SE3<T> sols[4];
  const int ns = kneip_main(Xw, bv, sols);
  T minScore = std::numeric_limits<T>::max();  // :258
  int minIndex = -1;
  const V3<T> x3 = col3(Xw, 3), b3 = col3(bv, 3);
  for (int i = 0; i < ns; i++) {
    V3<T> pc = sols[i].so3.matrix() * x3 + sols[i].t;  // :262
    pc = pc / norm(pc);                                 // :273
    const T score = (T)(1.0 - (double)dot(pc, b3));     // :276 (1.0 is double)
    if (score < minScore) {
      minScore = score;
      minIndex = i;
    }
  }
  if (minIndex != -1) {
    *out = sols[minIndex];
    return true;
  }
  return false;
}

// ------------------------------------------------------------------------------------------
// nl_2p — AbsoluteOrientationNormal.hpp:77-142 (one oriented point + one point).
// ------------------------------------------------------------------------------------------
template <class T>
inline SO3<T> so3_from_angle_axis(T angle, const V3<T>& axis) {
  // Eigen Quaternion = AngleAxis: ha = 0.5*angle; w = cos(ha); vec = sin(ha)*axis; then the
  // explicit-quaternion SO3 ctor normalises (so3.hpp:578-585).
  const T ha = T(0.5) * angle;
  T s, c;
  m_sincos(ha, &s, &c);
  return SO3<T>::from_quat(Quat<T>(c, s * axis[0], s * axis[1], s * axis[2]));
}

template <class T>
inline SE3<T> nl_2p(const V3<T>& pt1_c, const V3<T>& nl1_c, const V3<T>& pt2_c, const V3<T>& pt1_w,
                    const V3<T>& nl1_w, const V3<T>& pt2_w) {
  const V3<T> c_w = pt1_w;                  // :87
  const T alpha = m_acos(nl1_w[0]);         // :89
  V3<T> axis(T(0), nl1_w[2], -nl1_w[1]);    // :90
  normalize(axis);                          // :91 (:100 normalises again: idempotent up to rounding, result unused)
  const SO3<T> R_g_f_w = so3_from_angle_axis(alpha, axis);  // :94-96
  const V3<T> c_c = pt1_c;                  // :102
  const T beta = m_acos(nl1_c[0]);          // :103
  V3<T> axis2(T(0), nl1_c[2], -nl1_c[1]);   // :104
  normalize(axis2);                         // :105
  const SO3<T> R_gp_f_c = so3_from_angle_axis(beta, axis2);  // :107-109
  V3<T> pt2_g = R_g_f_w * (pt2_w - c_w);    // :116
  pt2_g[0] = T(0);
  normalize(pt2_g);
  V3<T> pt2_gp = R_gp_f_c * (pt2_c - c_c);  // :117
  pt2_gp[0] = T(0);
  normalize(pt2_gp);
  const T gamma = m_acos(dot(pt2_g, pt2_gp));  // :119 (unsigned angle: reference quirk)
  const SO3<T> R_gp_f_g = so3_from_angle_axis(gamma, V3<T>(T(1), T(0), T(0)));  // :120-124
  const SO3<T> R_c_f_gp = R_gp_f_c.inverse();                                   // :127
  SE3<T> sol;
  sol.so3 = (R_c_f_gp * R_gp_f_g) * R_g_f_w;  // :128
  sol.t = c_c - sol.so3 * c_w;                 // :139
  return sol;
}

// ------------------------------------------------------------------------------------------
// MinimalSolvers.hpp:49-104 `ev` — closed-form eigenvalues of a symmetric 3x3 (trigonometric
// method). Never called by the reference and its eigenvector half does not compile
// (MinimalSolvers.hpp:92); the eigenvalue half is restated for completeness of row #9.
// ------------------------------------------------------------------------------------------
template <class T>
inline void sym3_eigenvalues(const M3<T>& A, T eig[3]) {
  const T p1 = A(0, 1) * A(0, 1) + A(0, 2) * A(0, 2) + A(1, 2) * A(1, 2);
  if (std::fabs(p1) < 0.00001) {  // MinimalSolvers.hpp:54
    eig[0] = A(0, 0);
    eig[1] = A(1, 1);
    eig[2] = A(2, 2);
    return;
  }
  T q = A(0, 0) + A(1, 1) + A(2, 2);  // :61
  q /= 3;
  const T t1 = A(0, 0) - q, t2 = A(1, 1) - q, t3 = A(2, 2) - q;
  const T p2 = t1 * t1 + t2 * t2 + t3 * t3 + 2 * p1;  // :65
  const T p = std::sqrt(p2 / 6);
  M3<T> B;  // :67  (1/p) * (M - q I)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B(i, j) = (1 / p) * (A(i, j) - q * (i == j ? T(1) : T(0)));
  const T r = det3(B) / T(2);
  T phi;
  if (r <= T(-1))
    phi = T(3.14159265358979323846) / T(3);
  else if (r >= T(1))
    phi = T(0);
  else
    phi = m_acos(r) / T(3);
  T s, c;
  m_sincos(phi, &s, &c);
  eig[0] = q + T(2) * p * c;
  T s2, c2;
  m_sincos(phi + T(2.0 * 3.14159265358979323846 / 3.0), &s2, &c2);
  eig[2] = q + T(2) * p * c2;
  eig[1] = T(3) * q - eig[0] - eig[2];
}

}  // namespace orc

#endif  // ORACLE_SOLVERS_HPP_
