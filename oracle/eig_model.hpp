// oracle/eig_model.hpp — TEST INFRASTRUCTURE (see oracle/README.md). PARITY UNPINNED.
//
// Model of the Eigen 3.3 arithmetic the reference reaches (Eigen itself is an external,
// un-vendored, un-pinned dependency: /root/reference/CMakeLists.txt:17). Each routine
// states the Eigen rule it follows; all of it is recalled, none could be checked here.
//
// Reduction rule used throughout ("redux tree"): Eigen's non-vectorised, fully unrolled
// redux of a fixed-size expression of length L splits at L/2 recursively, so
//   sum3(a,b,c)   = a + (b + c)
//   sum4(a,b,c,d) = (a + b) + (c + d)
//   sum9          = ((c0+c1)+(c2+c3)) + ((c4+c5)+(c6+(c7+c8)))
// Vector3/Matrix3 of float/double are not packet-aligned sizes, so dot(), squaredNorm(),
// trace() and the coefficient of a small lazy matrix product
// ((lhs.row(i).transpose().cwiseProduct(rhs.col(j))).sum()) all follow that tree.
// The reference is built with "-Wall -std=c++11" only (CMakeLists.txt:13-15): SSE2, no FMA.
#ifndef ORACLE_EIG_MODEL_HPP_
#define ORACLE_EIG_MODEL_HPP_

#include <cmath>
#include <limits>

namespace orc {

template <class T>
struct V3 {
  T v[3];
  V3() : v{T(0), T(0), T(0)} {}
  V3(T a, T b, T c) : v{a, b, c} {}
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
};

template <class T>
struct M3 {
  // m[r][c]
  T m[3][3];
  M3() {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m[r][c] = T(0);
  }
  static M3 identity() {
    M3 a;
    a.m[0][0] = a.m[1][1] = a.m[2][2] = T(1);
    return a;
  }
  T& operator()(int r, int c) { return m[r][c]; }
  const T& operator()(int r, int c) const { return m[r][c]; }
};

// Eigen::Quaternion stores coeffs as (x, y, z, w).
template <class T>
struct Quat {
  T x, y, z, w;
  Quat() : x(T(0)), y(T(0)), z(T(0)), w(T(1)) {}
  Quat(T w_, T x_, T y_, T z_) : x(x_), y(y_), z(z_), w(w_) {}
};

template <class T>
inline T sum3(T a, T b, T c) {
  return a + (b + c);
}
template <class T>
inline T sum4(T a, T b, T c, T d) {
  return (a + b) + (c + d);
}
// Sensitivity builds (oracle/README.md, "version-dependent forks"; never used by the parity tests): -DORC_EIG_VARIANT=
//   bit 0: dot / squaredNorm of a 3-vector as (a0 + a1) + a2 — what Eigen >= 3.3 does for double with one Packet2d;
//   bit 1: the coefficient of a small matrix product accumulated in index order — Eigen 3.2.
#ifndef ORC_EIG_VARIANT
#define ORC_EIG_VARIANT 0
#endif
template <class T>
inline T vsum3(T a, T b, T c) {  // 3-vector redux
  return (ORC_EIG_VARIANT & 1) ? (a + b) + c : sum3(a, b, c);
}
template <class T>
inline T psum3(T a, T b, T c) {  // product coefficient
  return (ORC_EIG_VARIANT & 2) ? (a + b) + c : sum3(a, b, c);
}

template <class T>
inline V3<T> operator+(const V3<T>& a, const V3<T>& b) {
  return V3<T>(a[0] + b[0], a[1] + b[1], a[2] + b[2]);
}
template <class T>
inline V3<T> operator-(const V3<T>& a, const V3<T>& b) {
  return V3<T>(a[0] - b[0], a[1] - b[1], a[2] - b[2]);
}
template <class T>
inline V3<T> operator-(const V3<T>& a) {
  return V3<T>(-a[0], -a[1], -a[2]);
}
template <class T>
inline V3<T> operator*(T s, const V3<T>& a) {
  return V3<T>(s * a[0], s * a[1], s * a[2]);
}
template <class T>
inline V3<T> operator*(const V3<T>& a, T s) {
  return V3<T>(a[0] * s, a[1] * s, a[2] * s);
}
// Eigen: vector / scalar is a true per-coefficient division (CwiseBinaryOp quotient).
template <class T>
inline V3<T> operator/(const V3<T>& a, T s) {
  return V3<T>(a[0] / s, a[1] / s, a[2] / s);
}
template <class T>
inline T dot(const V3<T>& a, const V3<T>& b) {
  return vsum3(a[0] * b[0], a[1] * b[1], a[2] * b[2]);
}
template <class T>
inline T squared_norm(const V3<T>& a) {
  return vsum3(a[0] * a[0], a[1] * a[1], a[2] * a[2]);
}
template <class T>
inline T norm(const V3<T>& a) {
  return std::sqrt(squared_norm(a));
}
// Eigen 3.3 MatrixBase::normalize(): z = squaredNorm(); if (z > 0) *this /= sqrt(z).
template <class T>
inline void normalize(V3<T>& a) {
  T z = squared_norm(a);
  if (z > T(0)) a = a / std::sqrt(z);
}
// Eigen cross3 (non-vectorised): (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0).
template <class T>
inline V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return V3<T>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}

template <class T>
inline M3<T> transpose(const M3<T>& a) {
  M3<T> r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r(i, j) = a(j, i);
  return r;
}
// Small fixed-size lazy product: each coefficient is a redux-tree dot of row and column.
template <class T>
inline M3<T> operator*(const M3<T>& a, const M3<T>& b) {
  M3<T> r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r(i, j) = psum3(a(i, 0) * b(0, j), a(i, 1) * b(1, j), a(i, 2) * b(2, j));
  return r;
}
template <class T>
inline V3<T> operator*(const M3<T>& a, const V3<T>& x) {
  return V3<T>(psum3(a(0, 0) * x[0], a(0, 1) * x[1], a(0, 2) * x[2]),
               psum3(a(1, 0) * x[0], a(1, 1) * x[1], a(1, 2) * x[2]),
               psum3(a(2, 0) * x[0], a(2, 1) * x[1], a(2, 2) * x[2]));
}
template <class T>
inline M3<T> operator-(const M3<T>& a, const M3<T>& b) {
  M3<T> r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r(i, j) = a(i, j) - b(i, j);
  return r;
}
template <class T>
inline M3<T> outer(const V3<T>& a, const V3<T>& b) {
  M3<T> r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r(i, j) = a[i] * b[j];
  return r;
}
// Eigen determinant_impl<Derived,3>: bruteforce_det3_helper(m,0,1,2) - helper(m,1,0,2) + helper(m,2,0,1)
// with helper(m,a,b,c) = m(0,a) * (m(1,b)*m(2,c) - m(1,c)*m(2,b)).
template <class T>
inline T det3(const M3<T>& m) {
  auto h = [&](int a, int b, int c) { return m(0, a) * (m(1, b) * m(2, c) - m(1, c) * m(2, b)); };
  return h(0, 1, 2) - h(1, 0, 2) + h(2, 0, 1);
}
// Frobenius norm of a 3x3 (column-major linear index), redux tree of 9 squares.
template <class T>
inline T frob_norm(const M3<T>& a) {
  T c[9];
  int k = 0;
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      c[k] = a(i, j) * a(i, j);
      ++k;
    }
  return std::sqrt(((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + (c[7] + c[8]))));
}

// Eigen QuaternionBase::_transformVector (reached from so3.hpp:238-240):
//   uv = vec().cross(v); uv += uv; return v + w()*uv + vec().cross(uv);
template <class T>
inline V3<T> quat_rotate(const Quat<T>& q, const V3<T>& v) {
  const V3<T> qv(q.x, q.y, q.z);
  V3<T> uv = cross(qv, v);
  uv = uv + uv;
  const V3<T> c2 = cross(qv, uv);
  return V3<T>((v[0] + q.w * uv[0]) + c2[0], (v[1] + q.w * uv[1]) + c2[1], (v[2] + q.w * uv[2]) + c2[2]);
}

// Eigen QuaternionBase::toRotationMatrix (reached from so3.hpp:204-206).
template <class T>
inline M3<T> quat_to_matrix(const Quat<T>& q) {
  const T tx = T(2) * q.x, ty = T(2) * q.y, tz = T(2) * q.z;
  const T twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const T txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const T tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3<T> r;
  r(0, 0) = T(1) - (tyy + tzz);
  r(0, 1) = txy - twz;
  r(0, 2) = txz + twy;
  r(1, 0) = txy + twz;
  r(1, 1) = T(1) - (txx + tzz);
  r(1, 2) = tyz - twx;
  r(2, 0) = txz - twy;
  r(2, 1) = tyz + twx;
  r(2, 2) = T(1) - (txx + tyy);
  return r;
}

// Eigen quaternionbase_assign_impl<Other,3,3> (Shoemake), reached from SO3(Matrix3) so3.hpp:561.
template <class T>
inline Quat<T> quat_from_matrix(const M3<T>& mat) {
  Quat<T> q;
  T t = sum3(mat(0, 0), mat(1, 1), mat(2, 2));  // trace(): diagonal().sum()
  if (t > T(0)) {
    t = std::sqrt(t + T(1.0));
    q.w = T(0.5) * t;
    t = T(0.5) / t;
    q.x = (mat(2, 1) - mat(1, 2)) * t;
    q.y = (mat(0, 2) - mat(2, 0)) * t;
    q.z = (mat(1, 0) - mat(0, 1)) * t;
  } else {
    int i = 0;
    if (mat(1, 1) > mat(0, 0)) i = 1;
    if (mat(2, 2) > mat(i, i)) i = 2;
    const int j = (i + 1) % 3;
    const int k = (j + 1) % 3;
    t = std::sqrt(mat(i, i) - mat(j, j) - mat(k, k) + T(1.0));
    T c[3];
    c[i] = T(0.5) * t;
    t = T(0.5) / t;
    q.w = (mat(k, j) - mat(j, k)) * t;
    c[j] = (mat(j, i) + mat(i, j)) * t;
    c[k] = (mat(k, i) + mat(i, k)) * t;
    q.x = c[0];
    q.y = c[1];
    q.z = c[2];
  }
  return q;
}

// Generic (non-SSE) Eigen quat_product.
template <class T>
inline Quat<T> quat_mul(const Quat<T>& a, const Quat<T>& b) {
  return Quat<T>(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
                 a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                 a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
                 a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x);
}
template <class T>
inline T quat_squared_norm(const Quat<T>& q) {
  return sum4(q.x * q.x, q.y * q.y, q.z * q.z, q.w * q.w);
}

// ---- JacobiSVD of a square 3x3 (two-sided Jacobi, no QR preconditioner for square input) ----
// Follows Eigen 3.3 JacobiSVD::compute + real_2x2_jacobi_svd + JacobiRotation::makeJacobi.
// Reached from AbsoluteOrientation.hpp:79, AbsoluteOrientationNormal.hpp:44,188,512.
template <class T>
struct JRot {
  T c, s;
};
template <class T>
inline JRot<T> jrot_mul(const JRot<T>& a, const JRot<T>& b) {
  // JacobiRotation::operator*: (c*oc - s*os, c*os + s*oc) for real scalars
  return JRot<T>{a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c};
}
template <class T>
inline JRot<T> jrot_transpose(const JRot<T>& a) {
  return JRot<T>{a.c, -a.s};
}
// apply_rotation_in_the_plane: x_i' = c*x_i + s*y_i ; y_i' = -s*x_i + c*y_i
template <class T>
inline void rot_rows(M3<T>& w, int p, int q, const JRot<T>& j) {  // applyOnTheLeft(p,q,j)
  if (j.c == T(1) && j.s == T(0)) return;
  for (int i = 0; i < 3; ++i) {
    const T xi = w(p, i), yi = w(q, i);
    w(p, i) = j.c * xi + j.s * yi;
    w(q, i) = -j.s * xi + j.c * yi;
  }
}
template <class T>
inline void rot_cols(M3<T>& w, int p, int q, const JRot<T>& j) {  // applyOnTheRight(p,q,j): uses j.transpose()
  const JRot<T> jt = jrot_transpose(j);
  if (jt.c == T(1) && jt.s == T(0)) return;
  for (int i = 0; i < 3; ++i) {
    const T xi = w(i, p), yi = w(i, q);
    w(i, p) = jt.c * xi + jt.s * yi;
    w(i, q) = -jt.s * xi + jt.c * yi;
  }
}

template <class T>
struct SVD3 {
  M3<T> U, V;
  T s[3];
};

template <class T>
inline SVD3<T> jacobi_svd3(const M3<T>& a) {
  using std::abs;
  using std::sqrt;
  const T precision = T(2) * std::numeric_limits<T>::epsilon();
  const T consider_as_zero = (std::numeric_limits<T>::min)();
  T scale = T(0);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const T v = abs(a(i, j));
      if (v > scale) scale = v;  // maxCoeff; NaN never wins
    }
  if (scale == T(0)) scale = T(1);
  M3<T> W;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) W(i, j) = a(i, j) / scale;
  SVD3<T> out;
  out.U = M3<T>::identity();
  out.V = M3<T>::identity();
  T max_diag = abs(W(0, 0));
  if (abs(W(1, 1)) > max_diag) max_diag = abs(W(1, 1));
  if (abs(W(2, 2)) > max_diag) max_diag = abs(W(2, 2));
  bool finished = false;
  int sweeps = 0;
  while (!finished && sweeps < 64) {  // Eigen has no cap; 64 is never reached on finite input
    finished = true;
    ++sweeps;
    for (int p = 1; p < 3; ++p) {
      for (int q = 0; q < p; ++q) {
        const T pm = precision * max_diag;
        const T threshold = consider_as_zero > pm ? consider_as_zero : pm;  // numext::maxi
        if (abs(W(p, q)) > threshold || abs(W(q, p)) > threshold) {
          finished = false;
          // real_2x2_jacobi_svd
          T m00 = W(p, p), m01 = W(p, q), m10 = W(q, p), m11 = W(q, q);
          JRot<T> rot1;
          const T t = m00 + m11;
          const T d = m10 - m01;
          if (abs(d) < consider_as_zero) {
            rot1.s = T(0);
            rot1.c = T(1);
          } else {
            const T u = t / d;
            const T tmp = sqrt(T(1) + u * u);
            rot1.s = T(1) / tmp;
            rot1.c = u / tmp;
          }
          // m.applyOnTheLeft(0,1,rot1)
          if (!(rot1.c == T(1) && rot1.s == T(0))) {
            const T a0 = m00, a1 = m01, b0 = m10, b1 = m11;
            m00 = rot1.c * a0 + rot1.s * b0;
            m01 = rot1.c * a1 + rot1.s * b1;
            m10 = -rot1.s * a0 + rot1.c * b0;
            m11 = -rot1.s * a1 + rot1.c * b1;
          }
          // j_right.makeJacobi(m00, m01, m11)
          JRot<T> jr;
          const T deno = T(2) * abs(m01);
          if (deno < consider_as_zero) {
            jr.c = T(1);
            jr.s = T(0);
          } else {
            const T tau = (m00 - m11) / deno;
            const T w = sqrt(tau * tau + T(1));
            T tt;
            if (tau > T(0))
              tt = T(1) / (tau + w);
            else
              tt = T(1) / (tau - w);
            const T sign_t = tt > T(0) ? T(1) : T(-1);
            const T n = T(1) / sqrt(tt * tt + T(1));
            jr.s = -sign_t * (m01 / abs(m01)) * abs(tt) * n;
            jr.c = n;
          }
          const JRot<T> jl = jrot_mul(rot1, jrot_transpose(jr));
          rot_rows(W, p, q, jl);
          rot_cols(out.U, p, q, jrot_transpose(jl));
          rot_cols(W, p, q, jr);
          rot_cols(out.V, p, q, jr);
          const T app = abs(W(p, p)), aqq = abs(W(q, q));
          const T mx = app > aqq ? app : aqq;  // numext::maxi(a,b) = a < b ? b : a
          if (max_diag < mx) max_diag = mx;
        }
      }
    }
  }
  for (int i = 0; i < 3; ++i) {
    const T aii = W(i, i);
    out.s[i] = abs(aii);
    if (aii < T(0))
      for (int r = 0; r < 3; ++r) out.U(r, i) = -out.U(r, i);
  }
  for (int i = 0; i < 3; ++i) out.s[i] *= scale;
  // selection sort, descending; maxCoeff returns the FIRST maximum
  for (int i = 0; i < 3; ++i) {
    int pos = i;
    T best = out.s[i];
    for (int k = i + 1; k < 3; ++k)
      if (out.s[k] > best) {
        best = out.s[k];
        pos = k;
      }
    if (best == T(0)) break;
    if (pos != i) {
      std::swap(out.s[i], out.s[pos]);
      for (int r = 0; r < 3; ++r) {
        std::swap(out.U(r, i), out.U(r, pos));
        std::swap(out.V(r, i), out.V(r, pos));
      }
    }
  }
  return out;
}

}  // namespace orc

#endif  // ORACLE_EIG_MODEL_HPP_
