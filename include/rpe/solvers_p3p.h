// rpe/solvers_p3p.h — P3P (Kneip) with the Ferrari quartic, and the point+normal solver; host + device.
//
// One hypothesis per thread on the GPU; -fmad=false / -ffp-contract=off TUs (see solvers.h). Transcendentals go through
// include/rpe/det_math.h so that the CPU oracle in DET mode produces the same bits.
//
// Reference routines (paths into /root/reference/pose):
//   o4_roots      P3P.hpp:11-60      (std::complex<Tp> pow/sqrt -> explicit complex arithmetic,
//                                     exact real powers of the real P and Q, principal branches)
//   kneip_main    P3P.hpp:63-232
//   kneip (4th-point disambiguation)  P3P.hpp:250-294 and the inline copy in kneip_ransac :338-354
//   nl_2p         AbsoluteOrientationNormal.hpp:77-142
#ifndef RPE_SOLVERS_P3P_H_
#define RPE_SOLVERS_P3P_H_

#include "solvers.h"

namespace rpe {

template <class T>
struct Cplx {
  T re, im;
};
template <class T>
RPE_FN Cplx<T> c_mul(Cplx<T> a, Cplx<T> b) {
  Cplx<T> r;
  r.re = a.re * b.re - a.im * b.im;
  r.im = a.re * b.im + a.im * b.re;
  return r;
}
template <class T>
RPE_FN Cplx<T> c_div(Cplx<T> a, Cplx<T> b) {  // Smith
  Cplx<T> r;
  if (t_abs(b.re) < t_abs(b.im)) {
    const T ratio = b.re / b.im;
    const T den = (b.re * ratio) + b.im;
    r.re = ((a.re * ratio) + a.im) / den;
    r.im = ((a.im * ratio) - a.re) / den;
  } else {
    const T ratio = b.im / b.re;
    const T den = (b.im * ratio) + b.re;
    r.re = ((a.im * ratio) + a.re) / den;
    r.im = (a.im - (a.re * ratio)) / den;
  }
  return r;
}
template <class T>
RPE_FN Cplx<T> c_sqrt(Cplx<T> z) {
  Cplx<T> r;
  if (z.re == T(0) && z.im == T(0)) {
    r.re = T(0);
    r.im = T(0);
    return r;
  }
  const T h = t_sqrt(z.re * z.re + z.im * z.im);
  const T t = t_sqrt((t_abs(z.re) + h) * T(0.5));
  if (z.re >= T(0)) {
    r.re = t;
    r.im = z.im / (T(2) * t);
  } else {
    r.re = t_abs(z.im) / (T(2) * t);
    r.im = z.im < T(0) ? -t : t;
  }
  return r;
}
template <class T>
RPE_FN Cplx<T> c_cbrt(Cplx<T> z) {
  Cplx<T> r;
  if (z.re == T(0) && z.im == T(0)) {
    r.re = T(0);
    r.im = T(0);
    return r;
  }
  const T h = t_sqrt(z.re * z.re + z.im * z.im);
  const T mag = det::cbrt_t(h);
  const T th = det::atan2_t(z.im, z.re) / T(3);
  T s, c;
  det::sincos_t(th, &s, &c);
  r.re = mag * c;
  r.im = mag * s;
  return r;
}

template <class T>
RPE_FN void o4_roots_dev(const T* f, T* roots) {
  const T A = f[0], B = f[1], C = f[2], D = f[3], E = f[4];
  const T A_pw2 = A * A, B_pw2 = B * B;
  const T A_pw3 = A_pw2 * A, B_pw3 = B_pw2 * B;
  const T A_pw4 = A_pw3 * A, B_pw4 = B_pw3 * B;
  const T alpha = -3 * B_pw2 / (8 * A_pw2) + C / A;
  const T beta = B_pw3 / (8 * A_pw3) - B * C / (2 * A_pw2) + D / A;
  const T gamma = -3 * B_pw4 / (256 * A_pw4) + B_pw2 * C / (16 * A_pw3) - B * D / (4 * A_pw2) + E / A;
  const T alpha_pw2 = alpha * alpha;
  const T alpha_pw3 = alpha_pw2 * alpha;
  const T Pre = -alpha_pw2 / 12 - gamma;
  const T Qre = (T)((-alpha_pw3 / 108 + alpha * gamma / 3) - ((double)beta * (double)beta) / 8);
  const T b4a = -B / (T(4.) * A);
  Cplx<T> P, Q;
  P.re = Pre;
  P.im = T(0);
  Q.re = Qre;
  Q.im = T(0);
  const Cplx<T> Q2 = c_mul(Q, Q);
  const Cplx<T> P3 = c_mul(c_mul(P, P), P);
  Cplx<T> rad;
  rad.re = Q2.re / T(4.) + P3.re / T(27.);
  rad.im = Q2.im / T(4.) + P3.im / T(27.);
  const Cplx<T> sq = c_sqrt(rad);
  Cplx<T> R;
  R.re = -Q.re / T(2.0) + sq.re;
  R.im = -Q.im / T(2.0) + sq.im;
  const Cplx<T> U = c_cbrt(R);
  const T m56a = -T(5.0) * alpha / T(6.);
  Cplx<T> y;
  if (U.re == 0) {
    const Cplx<T> cq = c_cbrt(Q);
    y.re = m56a - cq.re;
    y.im = -cq.im;
  } else {
    Cplx<T> U3;
    U3.re = T(3.) * U.re;
    U3.im = T(3.) * U.im;
    const Cplx<T> pu = c_div(P, U3);
    y.re = (m56a - pu.re) + U.re;
    y.im = (-pu.im) + U.im;
  }
  Cplx<T> wa;
  wa.re = alpha + T(2.) * y.re;
  wa.im = T(2.) * y.im;
  const Cplx<T> w = c_sqrt(wa);
  Cplx<T> b2;
  b2.re = T(2.) * beta;
  b2.im = T(0);
  const Cplx<T> bw = c_div(b2, w);
  const T a3 = T(3.) * alpha;
  Cplx<T> base;
  base.re = a3 + T(2.) * y.re;
  base.im = T(2.) * y.im;
  Cplx<T> n1, n2;
  n1.re = -(base.re + bw.re);
  n1.im = -(base.im + bw.im);
  n2.re = -(base.re - bw.re);
  n2.im = -(base.im - bw.im);
  const Cplx<T> s1 = c_sqrt(n1);
  const Cplx<T> s2 = c_sqrt(n2);
  roots[0] = b4a + T(0.5) * (w.re + s1.re);
  roots[1] = b4a + T(0.5) * (w.re - s1.re);
  roots[2] = b4a + T(0.5) * (-w.re + s2.re);
  roots[3] = b4a + T(0.5) * (-w.re - s2.re);
}

template <class T>
RPE_FN void v_cross(const T* a, const T* b, T* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
template <class T>
RPE_FN T v_dot(const T* a, const T* b) {
  return sum3(a[0] * b[0], a[1] * b[1], a[2] * b[2]);
}
template <class T>
RPE_FN T v_norm(const T* a) {
  return t_sqrt(v_dot(a, a));
}
template <class T>
RPE_FN void mat_vec(const T* M, const T* x, T* y) {
  y[0] = sum3(M[0] * x[0], M[1] * x[1], M[2] * x[2]);
  y[1] = sum3(M[3] * x[0], M[4] * x[1], M[5] * x[2]);
  y[2] = sum3(M[6] * x[0], M[7] * x[1], M[8] * x[2]);
}
template <class T>
RPE_FN void quat_to_matrix_t(const T* q, T* R) {
  const T tx = T(2) * q[0], ty = T(2) * q[1], tz = T(2) * q[2];
  const T twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const T txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const T tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = T(1) - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = T(1) - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = T(1) - (txx + tyy);
}

// Up to 4 solutions (quaternion x,y,z,w + translation) in root order; returns their number.
template <class T>
RPE_FN int kneip_main_dev(const T* Xw, const T* bv, T (*qs)[4], T (*ts)[3]) {
  T P1[3], P2[3], P3[3];
  for (int r = 0; r < 3; ++r) {
    P1[r] = Xw[r];
    P2[r] = Xw[3 + r];
    P3[r] = Xw[6 + r];
  }
  T temp1[3], temp2[3], cr[3];
  for (int r = 0; r < 3; ++r) {
    temp1[r] = P2[r] - P1[r];
    temp2[r] = P3[r] - P1[r];
  }
  v_cross(temp1, temp2, cr);
  if (v_norm(cr) == 0) return 0;
  T f1[3], f2[3], f3[3];
  for (int r = 0; r < 3; ++r) {
    f1[r] = bv[r];
    f2[r] = bv[3 + r];
    f3[r] = bv[6 + r];
  }
  T RR[9];
  {
    T e3[3], e2[3];
    v_cross(f1, f2, e3);
    const T n = v_norm(e3);
    for (int r = 0; r < 3; ++r) e3[r] = e3[r] / n;
    v_cross(e3, f1, e2);
    for (int r = 0; r < 3; ++r) {
      RR[r] = f1[r];
      RR[3 + r] = e2[r];
      RR[6 + r] = e3[r];
    }
  }
  T f3r[3];
  mat_vec(RR, f3, f3r);
  if (f3r[2] > 0) {
    for (int r = 0; r < 3; ++r) {
      f1[r] = bv[3 + r];
      f2[r] = bv[r];
      f3[r] = bv[6 + r];
    }
    T e3[3], e2[3];
    v_cross(f1, f2, e3);
    const T n = v_norm(e3);
    for (int r = 0; r < 3; ++r) e3[r] = e3[r] / n;
    v_cross(e3, f1, e2);
    for (int r = 0; r < 3; ++r) {
      RR[r] = f1[r];
      RR[3 + r] = e2[r];
      RR[6 + r] = e3[r];
    }
    mat_vec(RR, f3, f3r);
    for (int r = 0; r < 3; ++r) {
      P1[r] = Xw[3 + r];
      P2[r] = Xw[r];
      P3[r] = Xw[6 + r];
    }
  }
  T N[9];
  T P3n[3];
  {
    T n1[3], n3[3], n2[3], d31[3];
    for (int r = 0; r < 3; ++r) n1[r] = P2[r] - P1[r];
    const T nn1 = v_norm(n1);
    for (int r = 0; r < 3; ++r) n1[r] = n1[r] / nn1;
    for (int r = 0; r < 3; ++r) d31[r] = P3[r] - P1[r];
    v_cross(n1, d31, n3);
    const T nn3 = v_norm(n3);
    for (int r = 0; r < 3; ++r) n3[r] = n3[r] / nn3;
    v_cross(n3, n1, n2);
    for (int r = 0; r < 3; ++r) {
      N[r] = n1[r];
      N[3 + r] = n2[r];
      N[6 + r] = n3[r];
    }
    mat_vec(N, d31, P3n);
  }
  const T d_12 = v_norm(temp1);
  const T f_1 = f3r[0] / f3r[2];
  const T f_2 = f3r[1] / f3r[2];
  const T p_1 = P3n[0];
  const T p_2 = P3n[1];
  const T cos_beta = v_dot(f1, f2);
  T b = (T)(1 / (1 - (double)cos_beta * (double)cos_beta) - 1);
  if (cos_beta < 0)
    b = -t_sqrt(b);
  else
    b = t_sqrt(b);
  const T f_1_pw2 = f_1 * f_1;
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
  factors[0] = -f_2_pw2 * p_2_pw4 - p_2_pw4 * f_1_pw2 - p_2_pw4;
  factors[1] = 2 * p_2_pw3 * d_12 * b + 2 * f_2_pw2 * p_2_pw3 * d_12 * b - 2 * f_2 * p_2_pw3 * f_1 * d_12;
  factors[2] = -f_2_pw2 * p_2_pw2 * p_1_pw2 - f_2_pw2 * p_2_pw2 * d_12_pw2 * b_pw2 - f_2_pw2 * p_2_pw2 * d_12_pw2 +
               f_2_pw2 * p_2_pw4 + p_2_pw4 * f_1_pw2 + 2 * p_1 * p_2_pw2 * d_12 +
               2 * f_1 * f_2 * p_1 * p_2_pw2 * d_12 * b - p_2_pw2 * p_1_pw2 * f_1_pw2 +
               2 * p_1 * p_2_pw2 * f_2_pw2 * d_12 - p_2_pw2 * d_12_pw2 * b_pw2 - 2 * p_1_pw2 * p_2_pw2;
  factors[3] = 2 * p_1_pw2 * p_2 * d_12 * b + 2 * f_2 * p_2_pw3 * f_1 * d_12 - 2 * f_2_pw2 * p_2_pw3 * d_12 * b -
               2 * p_1 * p_2 * d_12_pw2 * b;
  factors[4] = -2 * f_2 * p_2_pw2 * f_1 * p_1 * d_12 * b + f_2_pw2 * p_2_pw2 * d_12_pw2 + 2 * p_1_pw3 * d_12 -
               p_1_pw2 * d_12_pw2 + f_2_pw2 * p_2_pw2 * p_1_pw2 - p_1_pw4 - 2 * f_2_pw2 * p_2_pw2 * p_1 * d_12 +
               p_2_pw2 * f_1_pw2 * p_1_pw2 + f_2_pw2 * p_2_pw2 * d_12_pw2 * b_pw2;
  T realRoots[4];
  o4_roots_dev(factors, realRoots);
  T RRt[9], Nt[9];
  mat_transpose(RR, RRt);
  mat_transpose(N, Nt);
  int ns = 0;
  for (int i = 0; i < 4; i++) {
    const T root = realRoots[i];
    if (root != root) continue;
    const T cot_alpha = (-f_1 * p_1 / f_2 - root * p_2 + d_12 * b) / (-f_1 * root * p_2 / f_2 + p_1 - d_12);
    const T cos_theta = root;
    if (cos_theta > T(1) || cos_theta < T(-1)) continue;
    const T sin_theta = t_sqrt(1 - root * root);
    const T sin_alpha = t_sqrt(1 / (cot_alpha * cot_alpha + 1));
    T cos_alpha = t_sqrt(1 - sin_alpha * sin_alpha);
    if (cot_alpha < 0) cos_alpha = -cos_alpha;
    T C[3];
    C[0] = d_12 * cos_alpha * (sin_alpha * b + cos_alpha);
    C[1] = cos_theta * d_12 * sin_alpha * (sin_alpha * b + cos_alpha);
    C[2] = sin_theta * d_12 * sin_alpha * (sin_alpha * b + cos_alpha);
    T NtC[3];
    mat_vec(Nt, C, NtC);
    for (int r = 0; r < 3; ++r) C[r] = P1[r] + NtC[r];
    T R0[9];
    R0[0] = -cos_alpha;
    R0[1] = -sin_alpha * cos_theta;
    R0[2] = -sin_alpha * sin_theta;
    R0[3] = sin_alpha;
    R0[4] = -cos_alpha * cos_theta;
    R0[5] = -cos_alpha * sin_theta;
    R0[6] = T(0.0);
    R0[7] = -sin_theta;
    R0[8] = cos_theta;
    T T1[9], R[9];
    mat_mul(RRt, R0, T1);
    mat_mul(T1, N, R);
    if (R[0] != R[0]) continue;
    T q[4];
    if (!so3_from_matrix(R, q)) continue;  // the reference would abort() here
    T negR[9];
    for (int k = 0; k < 9; ++k) negR[k] = -R[k];
    T tt[3];
    mat_vec(negR, C, tt);
    for (int k = 0; k < 4; ++k) qs[ns][k] = q[k];
    for (int k = 0; k < 3; ++k) ts[ns][k] = tt[k];
    ++ns;
  }
  return ns;
}

// P3P + 4th-point disambiguation. Xw, bv: 3 x 4 column-major. `start` is the initial minScore.
template <class T>
RPE_FN bool kneip_select(const T* Xw, const T* bv, T start, T* q_out, T* t_out) {
  T qs[4][4], ts[4][3];
  const int ns = kneip_main_dev<T>(Xw, bv, qs, ts);
  T minScore = start;
  int minIndex = -1;
  for (int i = 0; i < ns; ++i) {
    T Rm[9], pc[3];
    quat_to_matrix_t(qs[i], Rm);
    mat_vec(Rm, Xw + 9, pc);
    for (int r = 0; r < 3; ++r) pc[r] = pc[r] + ts[i][r];
    const T n = v_norm(pc);
    for (int r = 0; r < 3; ++r) pc[r] = pc[r] / n;
    const T score = (T)(1.0 - (double)v_dot(pc, bv + 9));
    if (score < minScore) {
      minScore = score;
      minIndex = i;
    }
  }
  if (minIndex < 0) return false;
  for (int k = 0; k < 4; ++k) q_out[k] = qs[minIndex][k];
  for (int k = 0; k < 3; ++k) t_out[k] = ts[minIndex][k];
  return true;
}

// ---- nl_2p ---------------------------------------------------------------------------------------
template <class T>
RPE_FN void quat_normalized(const T* qin, T* q) {  // Sophus explicit-quaternion ctor
  const T len = t_sqrt((qin[0] * qin[0] + qin[1] * qin[1]) + (qin[2] * qin[2] + qin[3] * qin[3]));
  for (int k = 0; k < 4; ++k) q[k] = qin[k] / len;
}
template <class T>
RPE_FN void quat_from_angle_axis(T angle, const T* axis, T* q) {
  const T ha = T(0.5) * angle;
  T s, c;
  det::sincos_t(ha, &s, &c);
  const T raw[4] = {s * axis[0], s * axis[1], s * axis[2], c};
  quat_normalized(raw, q);
}
template <class T>
RPE_FN void v_normalize(T* a) {
  const T z = v_dot(a, a);
  if (z > T(0)) {
    const T n = t_sqrt(z);
    a[0] = a[0] / n;
    a[1] = a[1] / n;
    a[2] = a[2] / n;
  }
}
// Hamilton product + Sophus' first-order renormalisation (so3.hpp:255-272); coefficient order x,y,z,w
template <class T>
RPE_FN void so3_mul(const T* a, const T* b, T* r) {
  const T w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  const T x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const T y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const T z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  r[0] = x;
  r[1] = y;
  r[2] = z;
  r[3] = w;
  const T sn = (x * x + y * y) + (z * z + w * w);
  if (sn != T(1.0)) {
    const T k = T(2.0) / (T(1.0) + sn);
    r[0] *= k;
    r[1] *= k;
    r[2] *= k;
    r[3] *= k;
  }
}

template <class T>
RPE_FN void nl_2p(const T* pt1_c, const T* nl1_c, const T* pt2_c, const T* pt1_w, const T* nl1_w, const T* pt2_w,
                      T* q_out, T* t_out) {
  const T alpha = det::acos_t(nl1_w[0]);
  T axis[3] = {T(0), nl1_w[2], -nl1_w[1]};
  v_normalize(axis);
  T q_g_w[4];
  quat_from_angle_axis(alpha, axis, q_g_w);
  const T beta = det::acos_t(nl1_c[0]);
  T axis2[3] = {T(0), nl1_c[2], -nl1_c[1]};
  v_normalize(axis2);
  T q_gp_c[4];
  quat_from_angle_axis(beta, axis2, q_gp_c);
  T dw[3], dc[3], pt2_g[3], pt2_gp[3];
  for (int r = 0; r < 3; ++r) {
    dw[r] = pt2_w[r] - pt1_w[r];
    dc[r] = pt2_c[r] - pt1_c[r];
  }
  quat_rotate(q_g_w, dw, pt2_g);
  pt2_g[0] = T(0);
  v_normalize(pt2_g);
  quat_rotate(q_gp_c, dc, pt2_gp);
  pt2_gp[0] = T(0);
  v_normalize(pt2_gp);
  const T gamma = det::acos_t(v_dot(pt2_g, pt2_gp));
  const T ax3[3] = {T(1), T(0), T(0)};
  T q_gp_g[4];
  quat_from_angle_axis(gamma, ax3, q_gp_g);
  // inverse(): conjugate through the normalising constructor
  const T conj[4] = {-q_gp_c[0], -q_gp_c[1], -q_gp_c[2], q_gp_c[3]};
  T q_c_gp[4];
  quat_normalized(conj, q_c_gp);
  T tmp[4];
  so3_mul(q_c_gp, q_gp_g, tmp);
  so3_mul(tmp, q_g_w, q_out);
  T rc[3];
  quat_rotate(q_out, pt1_w, rc);
  for (int r = 0; r < 3; ++r) t_out[r] = pt1_c[r] - rc[r];
}

}  // namespace rpe

#endif  // RPE_SOLVERS_P3P_H_
