// rpe/solvers.h — minimal solvers, host + device, one hypothesis per thread on the GPU.
//
// Device TUs MUST be compiled with -fmad=false (csrc/Makefile does) and host TUs with -ffp-contract=off
// (or for a target without FMA): every `a*b + c` below has to stay two IEEE roundings so that the results
// are bit-identical to the reference's non-FMA x86-64 build (CMakeLists.txt:13-15 sets only -Wall
// -std=c++11). Division and sqrt are IEEE (nvcc defaults -prec-div=true -prec-sqrt=true, -ftz=false).
//
// Reference routines implemented here (paths into /root/reference/pose):
//   svd3_jacobi      Eigen::JacobiSVD as used at AbsoluteOrientation.hpp:79 (two-sided Jacobi, square input)
//   so3_from_matrix  Sophus::SO3(Matrix3) sophus/so3.hpp:561-566 (Eigen Quaternion(Matrix3) + ENSURE checks)
//   shinji3          shinji<Tp>(X_w, X_c, K=3) AbsoluteOrientation.hpp:47-99
//   kabsch_from_moments  the same closed form from accumulated moments (shinji_ls*, :273-342)
#ifndef RPE_SOLVERS_H_
#define RPE_SOLVERS_H_

#include <float.h>
#include <math.h>

#include "det_math.h"

#if defined(__CUDACC__)
#define RPE_FN __host__ __device__ __forceinline__
#define RPE_UNROLL _Pragma("unroll")
#else
#define RPE_FN inline
#define RPE_UNROLL
#endif

namespace rpe {

template <class T>
struct Lim;
template <>
struct Lim<float> {
  RPE_FN static float eps() { return FLT_EPSILON; }
  RPE_FN static float tiny() { return FLT_MIN; }
  RPE_FN static float sophus_eps() { return 1e-5f; }  // sophus/common.hpp:143-151
};
template <>
struct Lim<double> {
  RPE_FN static double eps() { return DBL_EPSILON; }
  RPE_FN static double tiny() { return DBL_MIN; }
  RPE_FN static double sophus_eps() { return 1e-10; }  // sophus/common.hpp:137-141
};

template <class T>
RPE_FN T t_abs(T a) {
  return a < T(0) ? -a : (a == T(0) ? T(0) : a);  // clears -0 like fabs
}
RPE_FN float t_sqrt(float a) { return sqrtf(a); }
RPE_FN double t_sqrt(double a) { return sqrt(a); }
template <class T>
RPE_FN T sum3(T a, T b, T c) {
  return a + (b + c);
}

// Plane rotation (c,s) applied to rows p,q: x' = c x + s y ; y' = -s x + c y
template <class T>
RPE_FN void rot_rows(T* W, int p, int q, T c, T s) {
  if (c == T(1) && s == T(0)) return;
RPE_UNROLL
  for (int i = 0; i < 3; ++i) {
    const T xi = W[3 * p + i], yi = W[3 * q + i];
    W[3 * p + i] = c * xi + s * yi;
    W[3 * q + i] = -s * xi + c * yi;
  }
}
// applyOnTheRight(p,q,j): columns p,q rotated by j.transpose() = (c,-s)
template <class T>
RPE_FN void rot_cols(T* W, int p, int q, T c, T s) {
  const T sc = -s;
  if (c == T(1) && sc == T(0)) return;
RPE_UNROLL
  for (int i = 0; i < 3; ++i) {
    const T xi = W[3 * i + p], yi = W[3 * i + q];
    W[3 * i + p] = c * xi + sc * yi;
    W[3 * i + q] = -sc * xi + c * yi;
  }
}

// Two-sided Jacobi SVD of a 3x3 (row-major). U, V row-major, s descending.
template <class T>
RPE_FN void svd3_jacobi(const T* A, T* U, T* V, T* s) {
  const T precision = T(2) * Lim<T>::eps();
  const T tiny = Lim<T>::tiny();
  T scale = T(0);
RPE_UNROLL
  for (int i = 0; i < 9; ++i) {
    const T v = t_abs(A[i]);
    if (v > scale) scale = v;
  }
  if (scale == T(0)) scale = T(1);
  T W[9];
RPE_UNROLL
  for (int i = 0; i < 9; ++i) {
    W[i] = A[i] / scale;
    U[i] = (i == 0 || i == 4 || i == 8) ? T(1) : T(0);
    V[i] = U[i];
  }
  T max_diag = t_abs(W[0]);
  if (t_abs(W[4]) > max_diag) max_diag = t_abs(W[4]);
  if (t_abs(W[8]) > max_diag) max_diag = t_abs(W[8]);
  bool finished = false;
  int sweeps = 0;
  while (!finished && sweeps < 64) {
    finished = true;
    ++sweeps;
    // (both loops unrolled on the device: with constant p, q the three matrices stay in registers instead of local memory)
RPE_UNROLL
    for (int p = 1; p < 3; ++p) {
RPE_UNROLL
      for (int q = 0; q < p; ++q) {
        const T pm = precision * max_diag;
        const T threshold = tiny > pm ? tiny : pm;
        if (t_abs(W[3 * p + q]) > threshold || t_abs(W[3 * q + p]) > threshold) {
          finished = false;
          T m00 = W[3 * p + p], m01 = W[3 * p + q], m10 = W[3 * q + p], m11 = W[3 * q + q];
          T r1c, r1s;
          const T tt = m00 + m11;
          const T d = m10 - m01;
          if (t_abs(d) < tiny) {
            r1s = T(0);
            r1c = T(1);
          } else {
            const T u = tt / d;
            const T tmp = t_sqrt(T(1) + u * u);
            r1s = T(1) / tmp;
            r1c = u / tmp;
          }
          if (!(r1c == T(1) && r1s == T(0))) {
            const T a0 = m00, a1 = m01, b0 = m10, b1 = m11;
            m00 = r1c * a0 + r1s * b0;
            m01 = r1c * a1 + r1s * b1;
            m10 = -r1s * a0 + r1c * b0;
            m11 = -r1s * a1 + r1c * b1;
          }
          T jrc, jrs;
          const T deno = T(2) * t_abs(m01);
          if (deno < tiny) {
            jrc = T(1);
            jrs = T(0);
          } else {
            const T tau = (m00 - m11) / deno;
            const T w = t_sqrt(tau * tau + T(1));
            T t2;
            if (tau > T(0))
              t2 = T(1) / (tau + w);
            else
              t2 = T(1) / (tau - w);
            const T sign_t = t2 > T(0) ? T(1) : T(-1);
            const T n = T(1) / t_sqrt(t2 * t2 + T(1));
            jrs = -sign_t * (m01 / t_abs(m01)) * t_abs(t2) * n;
            jrc = n;
          }
          // j_left = rot1 * j_right^T
          const T jtc = jrc, jts = -jrs;
          const T jlc = r1c * jtc - r1s * jts;
          const T jls = r1c * jts + r1s * jtc;
          rot_rows(W, p, q, jlc, jls);
          rot_cols(U, p, q, jlc, -jls);  // applyOnTheRight(p,q,j_left.transpose())
          rot_cols(W, p, q, jrc, jrs);
          rot_cols(V, p, q, jrc, jrs);
          const T app = t_abs(W[3 * p + p]), aqq = t_abs(W[3 * q + q]);
          const T mx = app > aqq ? app : aqq;
          if (max_diag < mx) max_diag = mx;
        }
      }
    }
  }
RPE_UNROLL
  for (int i = 0; i < 3; ++i) {
    const T aii = W[4 * i];
    s[i] = t_abs(aii);
    if (aii < T(0)) {
      U[i] = -U[i];
      U[3 + i] = -U[3 + i];
      U[6 + i] = -U[6 + i];
    }
  }
RPE_UNROLL
  for (int i = 0; i < 3; ++i) s[i] *= scale;
  // selection sort, largest first (the swap target is tested against constants so that no index is dynamic)
  bool sorting = true;
RPE_UNROLL
  for (int i = 0; i < 3; ++i) {
    int pos = i;
    T best = s[i];
RPE_UNROLL
    for (int k = i + 1; k < 3; ++k)
      if (sorting && s[k] > best) {
        best = s[k];
        pos = k;
      }
    if (best == T(0)) sorting = false;
RPE_UNROLL
    for (int k = i + 1; k < 3; ++k)
      if (sorting && pos == k) {
        T tmp = s[i];
        s[i] = s[k];
        s[k] = tmp;
RPE_UNROLL
        for (int r = 0; r < 3; ++r) {
          tmp = U[3 * r + i];
          U[3 * r + i] = U[3 * r + k];
          U[3 * r + k] = tmp;
          tmp = V[3 * r + i];
          V[3 * r + i] = V[3 * r + k];
          V[3 * r + k] = tmp;
        }
      }
  }
}

// C = A * B^T? No: plain C = A*B with redux-tree coefficients (Eigen lazy product, small fixed size).
template <class T>
RPE_FN void mat_mul(const T* A, const T* B, T* C) {
RPE_UNROLL
  for (int i = 0; i < 3; ++i)
RPE_UNROLL
    for (int j = 0; j < 3; ++j) C[3 * i + j] = sum3(A[3 * i] * B[j], A[3 * i + 1] * B[3 + j], A[3 * i + 2] * B[6 + j]);
}
template <class T>
RPE_FN void mat_transpose(const T* A, T* At) {
RPE_UNROLL
  for (int i = 0; i < 3; ++i)
RPE_UNROLL
    for (int j = 0; j < 3; ++j) At[3 * i + j] = A[3 * j + i];
}
// Eigen determinant_impl<3>: h(0,1,2) - h(1,0,2) + h(2,0,1), h(a,b,c) = m(0,a)*(m(1,b)*m(2,c) - m(1,c)*m(2,b))
template <class T>
RPE_FN T mat_det(const T* m) {
  const T h0 = m[0] * (m[4] * m[8] - m[5] * m[7]);
  const T h1 = m[1] * (m[3] * m[8] - m[5] * m[6]);
  const T h2 = m[2] * (m[3] * m[7] - m[4] * m[6]);
  return h0 - h1 + h2;
}

// Sophus::SO3(Matrix3): Shoemake quaternion WITHOUT renormalisation; returns false where the
// reference would abort (||R R^T - I||_F >= eps or det <= 0).
template <class T>
RPE_FN bool so3_from_matrix(const T* R, T* q /*x,y,z,w*/) {
  T t = sum3(R[0], R[4], R[8]);
  if (t > T(0)) {
    t = t_sqrt(t + T(1.0));
    q[3] = T(0.5) * t;
    t = T(0.5) / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3;
    const int k = (j + 1) % 3;
    t = t_sqrt(R[4 * i] - R[4 * j] - R[4 * k] + T(1.0));
    T c[3];
    c[i] = T(0.5) * t;
    t = T(0.5) / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    c[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    c[k] = (R[3 * k + i] + R[3 * i + k]) * t;
    q[0] = c[0];
    q[1] = c[1];
    q[2] = c[2];
  }
  // isOrthogonal: Frobenius norm of R R^T - I, squares summed in column-major order with the
  // 9-element redux tree ((c0+c1)+(c2+c3)) + ((c4+c5)+(c6+(c7+c8)))
  T Rt[9], P[9];
  mat_transpose(R, Rt);
  mat_mul(R, Rt, P);
  T c[9];
  int n = 0;
  for (int col = 0; col < 3; ++col)
    for (int row = 0; row < 3; ++row) {
      const T dlt = P[3 * row + col] - (row == col ? T(1) : T(0));
      c[n++] = dlt * dlt;
    }
  const T fro = t_sqrt(((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + (c[7] + c[8]))));
  return (fro < Lim<T>::sophus_eps()) && (mat_det(R) > T(0));
}

// v + w*uv + qv x uv (plain operators; this TU is compiled with -fmad=false)
template <class T>
RPE_FN void quat_rotate(const T* q, const T* v, T* out) {
  T uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  uv[0] = uv[0] + uv[0];
  uv[1] = uv[1] + uv[1];
  uv[2] = uv[2] + uv[2];
  const T c2[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  out[0] = (v[0] + q[3] * uv[0]) + c2[0];
  out[1] = (v[1] + q[3] * uv[1]) + c2[1];
  out[2] = (v[2] + q[3] * uv[2]) + c2[2];
}

// Rotation from the cross-covariance M (row-major): R = U diag(1,1,sign det(U V^T)) V^T.
template <class T>
RPE_FN bool rotation_from_covariance(const T* M, T* q) {
  T U[9], V[9], s[3];
  svd3_jacobi(M, U, V, s);
  T Vt[9], Tmp[9];
  mat_transpose(V, Vt);
  mat_mul(U, Vt, Tmp);
  const T d = mat_det(Tmp);
  if (d < T(0)) {
    // (U * I) * V^T with I = diag(1,1,-1): the product U*I is evaluated coefficient-wise by Eigen:
    // sum3(U(i,0)*I(0,j), U(i,1)*I(1,j), U(i,2)*I(2,j))
    T UI[9], R[9];
    const T I[9] = {T(1), T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(-1)};
    mat_mul(U, I, UI);
    mat_mul(UI, Vt, R);
    return so3_from_matrix(R, q);
  }
  return so3_from_matrix(Tmp, q);
}

// shinji with K = 3 sample columns; `cols` is the divisor X_w_.cols() (3 in shinji_ransac*, 4 in the hybrids).
template <class T>
RPE_FN bool shinji3(const T Xw[9], const T Xc[9], int cols, T* q, T* t) {
  T Cw[3] = {T(0), T(0), T(0)}, Cc[3] = {T(0), T(0), T(0)};
  for (int n = 0; n < 3; ++n)
    for (int r = 0; r < 3; ++r) {
      Cw[r] = Cw[r] + Xw[3 * n + r];
      Cc[r] = Cc[r] + Xc[3 * n + r];
    }
  for (int r = 0; r < 3; ++r) {
    Cw[r] = Cw[r] / T(3);
    Cc[r] = Cc[r] / T(3);
  }
  T M[9] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
  for (int n = 0; n < 3; ++n) {
    T Aw[3], Ac[3];
    for (int r = 0; r < 3; ++r) {
      Aw[r] = Xw[3 * n + r] - Cw[r];
      Ac[r] = Xc[3 * n + r] - Cc[r];
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) M[3 * i + j] = M[3 * i + j] + Ac[i] * Aw[j];
  }
  const T dv = (T)cols;
  for (int i = 0; i < 9; ++i) M[i] = M[i] / dv;
  const bool ok = rotation_from_covariance(M, q);
  T rc[3];
  quat_rotate(q, Cw, rc);
  t[0] = Cc[0] - rc[0];
  t[1] = Cc[1] - rc[1];
  t[2] = Cc[2] - rc[2];
  return ok;
}

}  // namespace rpe

#endif  // RPE_SOLVERS_H_
