// oracle/ref_shim/eigen_shim.hpp — TEST INFRASTRUCTURE (see oracle/README.md). PARITY UNPINNED.
//
// A small stand-in for the part of the Eigen 3 API that /root/reference/pose/*.hpp instantiates, so that the
// reference's OWN pose headers (samplers, minimal solvers, the seven RANSAC / PROSAC loops, adapters, refits) can be
// compiled UNMODIFIED from where they lie and run next to the oracle restatement (oracle/Makefile target `ref`).
// Eigen itself is absent from this image; every arithmetic rule below is the one oracle/eig_model.hpp documents
// (redux trees, determinant, JacobiSVD, quaternion conversions) and is taken from there, so what this build pins is the
// reference's control flow, formulas and operation order as written in its sources — not Eigen's internals.
//
// Evaluation is eager (every operator returns a concrete Matrix): for the small fixed-size expressions of the pose
// headers that gives the same coefficient arithmetic as Eigen's lazy evaluation (no reassociation is involved).
#ifndef ORACLE_REF_SHIM_EIGEN_SHIM_HPP_
#define ORACLE_REF_SHIM_EIGEN_SHIM_HPP_

#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <vector>

#include "../eig_model.hpp"

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

const int Dynamic = -1;
enum { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };

constexpr int shim_pick(int a, int b) { return a != Dynamic ? a : b; }

template <class T, int R, int C, int Opt = 0, int MR = R, int MC = C>
class Matrix;
template <class T, int R, int C>
class Block;
template <class M>
class JacobiSVD;

template <class D>
struct shim_traits;
template <class T, int R, int C, int O, int MR, int MC>
struct shim_traits<Matrix<T, R, C, O, MR, MC> > {
  typedef T Scalar;
  enum { Rows = R, Cols = C };
};
template <class T, int R, int C>
struct shim_traits<Block<T, R, C> > {
  typedef T Scalar;
  enum { Rows = R, Cols = C };
};

// 1 x 1 results (row vector times column vector) convert to their scalar, like Eigen's inner products.
template <class D, int R, int C>
struct ShimScalarConv {};
template <class D>
struct ShimScalarConv<D, 1, 1> {
  operator typename shim_traits<D>::Scalar() const { return static_cast<const D*>(this)->coeff(0, 0); }
};

template <class D>
class MatrixBase : public ShimScalarConv<D, shim_traits<D>::Rows, shim_traits<D>::Cols> {
 public:
  typedef typename shim_traits<D>::Scalar Scalar;
  enum { RowsAtCompileTime = shim_traits<D>::Rows, ColsAtCompileTime = shim_traits<D>::Cols };
  const D& derived() const { return *static_cast<const D*>(this); }
  D& derived() { return *static_cast<D*>(this); }
  int rows() const { return derived().rows_(); }
  int cols() const { return derived().cols_(); }
  int size() const { return rows() * cols(); }

  // ---- coefficient access (column-major linear index for the one-argument forms)
  Scalar operator()(int i, int j) const { return derived().coeff(i, j); }
  Scalar& operator()(int i, int j) { return derived().ref(i, j); }
  Scalar operator()(int i) const { return lin(i); }
  Scalar& operator()(int i) { return lin_ref(i); }
  Scalar operator[](int i) const { return lin(i); }
  Scalar& operator[](int i) { return lin_ref(i); }
  Scalar x() const { return lin(0); }
  Scalar y() const { return lin(1); }
  Scalar z() const { return lin(2); }

  // ---- reductions (oracle/eig_model.hpp: redux tree of a fixed-size expression of length 3)
  Scalar squaredNorm() const {
    if (size() == 3) return orc::sum3(lin(0) * lin(0), lin(1) * lin(1), lin(2) * lin(2));
    Scalar acc = Scalar(0);
    for (int i = 0; i < size(); ++i) acc += lin(i) * lin(i);
    return acc;
  }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  template <class O>
  Scalar dot(const MatrixBase<O>& o) const {
    assert(size() == o.size());
    if (size() == 3) return orc::sum3(lin(0) * o.lin(0), lin(1) * o.lin(1), lin(2) * o.lin(2));
    Scalar acc = Scalar(0);
    for (int i = 0; i < size(); ++i) acc += lin(i) * o.lin(i);
    return acc;
  }
  Scalar sum() const {
    Scalar acc = Scalar(0);
    for (int j = 0; j < cols(); ++j)
      for (int i = 0; i < rows(); ++i) acc += derived().coeff(i, j);
    return acc;
  }
  // Eigen 3.3 MatrixBase::normalize(): z = squaredNorm(); if (z > 0) *this /= sqrt(z)
  void normalize() {
    const Scalar z = squaredNorm();
    if (z > Scalar(0)) *this /= std::sqrt(z);
  }
  template <class O>
  Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {
    assert(size() == 3 && o.size() == 3);
    Matrix<Scalar, 3, 1> r;
    r(0) = lin(1) * o.lin(2) - lin(2) * o.lin(1);
    r(1) = lin(2) * o.lin(0) - lin(0) * o.lin(2);
    r(2) = lin(0) * o.lin(1) - lin(1) * o.lin(0);
    return r;
  }
  Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> transpose() const {
    Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> r;
    r.resize(cols(), rows());
    for (int i = 0; i < rows(); ++i)
      for (int j = 0; j < cols(); ++j) r.ref(j, i) = derived().coeff(i, j);
    return r;
  }
  Scalar determinant() const {
    assert(rows() == 3 && cols() == 3);
    return orc::det3(to_m3());
  }
  JacobiSVD<Matrix<Scalar, 3, 3> > jacobiSvd(unsigned int flags = 0) const;

  // ---- in-place
  D& setZero() {
    for (int j = 0; j < cols(); ++j)
      for (int i = 0; i < rows(); ++i) derived().ref(i, j) = Scalar(0);
    return derived();
  }
  D& setOnes() {
    for (int j = 0; j < cols(); ++j)
      for (int i = 0; i < rows(); ++i) derived().ref(i, j) = Scalar(1);
    return derived();
  }
  template <class O>
  D& operator+=(const MatrixBase<O>& o) {
    assert(rows() == o.rows() && cols() == o.cols());
    for (int j = 0; j < cols(); ++j)
      for (int i = 0; i < rows(); ++i) derived().ref(i, j) = derived().coeff(i, j) + o.derived().coeff(i, j);
    return derived();
  }
  template <class O>
  D& operator-=(const MatrixBase<O>& o) {
    assert(rows() == o.rows() && cols() == o.cols());
    for (int j = 0; j < cols(); ++j)
      for (int i = 0; i < rows(); ++i) derived().ref(i, j) = derived().coeff(i, j) - o.derived().coeff(i, j);
    return derived();
  }
  D& operator/=(Scalar s) {  // a true per-coefficient division (CwiseBinaryOp quotient)
    for (int j = 0; j < cols(); ++j)
      for (int i = 0; i < rows(); ++i) derived().ref(i, j) = derived().coeff(i, j) / s;
    return derived();
  }
  D& operator*=(Scalar s) {
    for (int j = 0; j < cols(); ++j)
      for (int i = 0; i < rows(); ++i) derived().ref(i, j) = derived().coeff(i, j) * s;
    return derived();
  }

  // ---- helpers shared with the Sophus stand-in
  Scalar lin(int i) const { return derived().coeff(i % rows(), i / rows()); }
  Scalar& lin_ref(int i) { return derived().ref(i % rows(), i / rows()); }
  orc::M3<Scalar> to_m3() const {
    assert(rows() == 3 && cols() == 3);
    orc::M3<Scalar> m;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) m(i, j) = derived().coeff(i, j);
    return m;
  }
  orc::V3<Scalar> to_v3() const {
    assert(size() == 3);
    return orc::V3<Scalar>(lin(0), lin(1), lin(2));
  }
};

// ------------------------------------------------------------------------------------------------------------
// storage: inline for fixed sizes, heap otherwise; column-major like Eigen's default
// ------------------------------------------------------------------------------------------------------------
template <class T, int R, int C, bool Fixed = (R != Dynamic && C != Dynamic)>
struct ShimStore;
template <class T, int R, int C>
struct ShimStore<T, R, C, true> {
  T d[R * C];
  ShimStore() {
    for (int i = 0; i < R * C; ++i) d[i] = T(0);
  }
  int rows_() const { return R; }
  int cols_() const { return C; }
  void resize(int r, int c) {
    assert(r == R && c == C);
    (void)r;
    (void)c;
  }
  T* data() { return d; }
  const T* data() const { return d; }
};
template <class T, int R, int C>
struct ShimStore<T, R, C, false> {
  std::vector<T> d;
  int r, c;
  ShimStore() : r(R == Dynamic ? 0 : R), c(C == Dynamic ? 0 : C) {}
  int rows_() const { return r; }
  int cols_() const { return c; }
  void resize(int rr, int cc) {
    assert((R == Dynamic || rr == R) && (C == Dynamic || cc == C));
    if (rr != r || cc != c) {
      r = rr;
      c = cc;
      d.assign((size_t)rr * cc, T(0));
    }
  }
  T* data() { return d.data(); }
  const T* data() const { return d.data(); }
};

template <class T, int R, int C, int Opt, int MR, int MC>
class Matrix : public MatrixBase<Matrix<T, R, C, Opt, MR, MC> >, public ShimStore<T, R, C> {
  typedef ShimStore<T, R, C> S;
  typedef MatrixBase<Matrix> B;

 public:
  using S::cols_;
  using S::data;
  using S::rows_;
  Matrix() {}
  Matrix(const Matrix&) = default;
  Matrix& operator=(const Matrix&) = default;
  // (rows, cols) for dynamic sizes, two coefficients for a fixed 2-vector
  template <class A, class Bb>
  Matrix(A a, Bb b) {
    init2(a, b, std::integral_constant<bool, (R != Dynamic && C != Dynamic)>());
  }
  template <class A, class Bb, class Cc>
  Matrix(A a, Bb b, Cc c) {
    static_assert(R * C == 3, "three-coefficient constructor of a 3-vector");
    S::d[0] = T(a);
    S::d[1] = T(b);
    S::d[2] = T(c);
  }
  template <class O>
  Matrix(const MatrixBase<O>& o) {
    assign(o);
  }
  template <class O>
  Matrix& operator=(const MatrixBase<O>& o) {
    assign(o);
    return *this;
  }
  void resize(int r, int c) { S::resize(r, c); }
  void resize(int n) {  // vectors
    if (C == 1)
      S::resize(n, 1);
    else
      S::resize(1, n);
  }
  T coeff(int i, int j) const {
    assert(i >= 0 && i < rows_() && j >= 0 && j < cols_());
    return data()[i + (size_t)j * rows_()];
  }
  T& ref(int i, int j) {
    assert(i >= 0 && i < rows_() && j >= 0 && j < cols_());
    return data()[i + (size_t)j * rows_()];
  }
  Block<T, R, 1> col(int j) { return Block<T, R, 1>(data() + (size_t)j * rows_(), rows_(), 1, rows_()); }
  const Block<T, R, 1> col(int j) const {
    return Block<T, R, 1>(const_cast<T*>(data()) + (size_t)j * rows_(), rows_(), 1, rows_());
  }
  Block<T, 1, C> row(int i) { return Block<T, 1, C>(data() + i, 1, cols_(), rows_()); }
  const Block<T, 1, C> row(int i) const { return Block<T, 1, C>(const_cast<T*>(data()) + i, 1, cols_(), rows_()); }
  static Matrix Zero() { return Matrix(); }
  static Matrix Ones(int r, int c) {
    Matrix m;
    m.resize(r, c);
    m.setOnes();
    return m;
  }
  // DenseBase::Random: every coefficient is internal::random<Scalar>() = x + (y - x) * Scalar(std::rand()) / Scalar(RAND_MAX)
  // with (x, y) = (-1, 1) for a signed scalar (Eigen 3.3 MathFunctions.h, random_default_impl), assigned in storage
  // (column-major) order. Used by the reference's Simulator only (Simulator.hpp:19,33,140).
  static Matrix Random(int r, int c) {
    Matrix m;
    m.resize(r, c);
    for (int j = 0; j < c; ++j)
      for (int i = 0; i < r; ++i) m.ref(i, j) = T(-1) + (T(1) - T(-1)) * T(std::rand()) / T(RAND_MAX);
    return m;
  }
  static Matrix Random() { return Random(R, C); }
  static Matrix Identity() {
    Matrix m;
    for (int i = 0; i < m.rows_() && i < m.cols_(); ++i) m.ref(i, i) = T(1);
    return m;
  }

 private:
  template <class A, class Bb>
  void init2(A a, Bb b, std::true_type) {
    static_assert(R * C == 2 || R == Dynamic, "two-coefficient constructor of a 2-vector");
    S::d[0] = T(a);
    S::d[1] = T(b);
  }
  template <class A, class Bb>
  void init2(A a, Bb b, std::false_type) {
    S::resize((int)a, (int)b);
  }
  template <class O>
  void assign(const MatrixBase<O>& o) {
    S::resize(o.rows(), o.cols());
    for (int j = 0; j < o.cols(); ++j)
      for (int i = 0; i < o.rows(); ++i) ref(i, j) = T(o.derived().coeff(i, j));
  }
};

// A column / row of a matrix: a view, assignable.
template <class T, int R, int C>
class Block : public MatrixBase<Block<T, R, C> > {
  T* p_;
  int r_, c_, ld_;  // ld_ = rows of the parent (column-major)

 public:
  Block(T* p, int r, int c, int ld) : p_(p), r_(r), c_(c), ld_(ld) {}
  Block(const Block&) = default;
  int rows_() const { return r_; }
  int cols_() const { return c_; }
  T coeff(int i, int j) const {
    assert(i >= 0 && i < r_ && j >= 0 && j < c_);
    return p_[i + (size_t)j * ld_];
  }
  T& ref(int i, int j) {
    assert(i >= 0 && i < r_ && j >= 0 && j < c_);
    return p_[i + (size_t)j * ld_];
  }
  T* data() { return p_; }
  const T* data() const { return p_; }
  Block& operator=(const Block& o) { return assign(o); }
  template <class O>
  Block& operator=(const MatrixBase<O>& o) {
    return assign(o);
  }

 private:
  template <class O>
  Block& assign(const MatrixBase<O>& o) {
    assert(o.rows() == r_ && o.cols() == c_);
    for (int j = 0; j < c_; ++j)
      for (int i = 0; i < r_; ++i) ref(i, j) = T(o.derived().coeff(i, j));
    return *this;
  }
};

// ------------------------------------------------------------------------------------------------------------
// operators
// ------------------------------------------------------------------------------------------------------------
template <class A, class B>
Matrix<typename A::Scalar, shim_pick(shim_traits<A>::Rows, shim_traits<B>::Rows), shim_pick(shim_traits<A>::Cols, shim_traits<B>::Cols)>
operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  Matrix<typename A::Scalar, shim_pick(shim_traits<A>::Rows, shim_traits<B>::Rows), shim_pick(shim_traits<A>::Cols, shim_traits<B>::Cols)> r;
  r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); ++j)
    for (int i = 0; i < a.rows(); ++i) r.ref(i, j) = a.derived().coeff(i, j) + b.derived().coeff(i, j);
  return r;
}
template <class A, class B>
Matrix<typename A::Scalar, shim_pick(shim_traits<A>::Rows, shim_traits<B>::Rows), shim_pick(shim_traits<A>::Cols, shim_traits<B>::Cols)>
operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  Matrix<typename A::Scalar, shim_pick(shim_traits<A>::Rows, shim_traits<B>::Rows), shim_pick(shim_traits<A>::Cols, shim_traits<B>::Cols)> r;
  r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); ++j)
    for (int i = 0; i < a.rows(); ++i) r.ref(i, j) = a.derived().coeff(i, j) - b.derived().coeff(i, j);
  return r;
}
template <class A>
Matrix<typename A::Scalar, shim_traits<A>::Rows, shim_traits<A>::Cols> operator-(const MatrixBase<A>& a) {
  Matrix<typename A::Scalar, shim_traits<A>::Rows, shim_traits<A>::Cols> r;
  r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); ++j)
    for (int i = 0; i < a.rows(); ++i) r.ref(i, j) = -a.derived().coeff(i, j);
  return r;
}
// Small product: every coefficient is the redux-tree dot of a row and a column (inner length 3: a + (b + c)).
template <class A, class B>
Matrix<typename A::Scalar, shim_traits<A>::Rows, shim_traits<B>::Cols> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  typedef typename A::Scalar T;
  assert(a.cols() == b.rows());
  Matrix<T, shim_traits<A>::Rows, shim_traits<B>::Cols> r;
  r.resize(a.rows(), b.cols());
  const int L = a.cols();
  for (int j = 0; j < b.cols(); ++j)
    for (int i = 0; i < a.rows(); ++i) {
      T v;
      if (L == 3)
        v = orc::sum3(a.derived().coeff(i, 0) * b.derived().coeff(0, j), a.derived().coeff(i, 1) * b.derived().coeff(1, j),
                      a.derived().coeff(i, 2) * b.derived().coeff(2, j));
      else if (L == 1)
        v = a.derived().coeff(i, 0) * b.derived().coeff(0, j);
      else {
        v = T(0);
        for (int k = 0; k < L; ++k) v += a.derived().coeff(i, k) * b.derived().coeff(k, j);
      }
      r.ref(i, j) = v;
    }
  return r;
}
template <class A>
Matrix<typename A::Scalar, shim_traits<A>::Rows, shim_traits<A>::Cols> operator*(typename A::Scalar s, const MatrixBase<A>& a) {
  Matrix<typename A::Scalar, shim_traits<A>::Rows, shim_traits<A>::Cols> r;
  r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); ++j)
    for (int i = 0; i < a.rows(); ++i) r.ref(i, j) = s * a.derived().coeff(i, j);
  return r;
}
template <class A>
Matrix<typename A::Scalar, shim_traits<A>::Rows, shim_traits<A>::Cols> operator*(const MatrixBase<A>& a, typename A::Scalar s) {
  Matrix<typename A::Scalar, shim_traits<A>::Rows, shim_traits<A>::Cols> r;
  r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); ++j)
    for (int i = 0; i < a.rows(); ++i) r.ref(i, j) = a.derived().coeff(i, j) * s;
  return r;
}
template <class A>
Matrix<typename A::Scalar, shim_traits<A>::Rows, shim_traits<A>::Cols> operator/(const MatrixBase<A>& a, typename A::Scalar s) {
  Matrix<typename A::Scalar, shim_traits<A>::Rows, shim_traits<A>::Cols> r;
  r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); ++j)
    for (int i = 0; i < a.rows(); ++i) r.ref(i, j) = a.derived().coeff(i, j) / s;
  return r;
}
template <class A>
std::ostream& operator<<(std::ostream& os, const MatrixBase<A>& a) {
  for (int i = 0; i < a.rows(); ++i) {
    for (int j = 0; j < a.cols(); ++j) os << (j ? " " : "") << a.derived().coeff(i, j);
    if (i + 1 < a.rows()) os << "\n";
  }
  return os;
}

// ------------------------------------------------------------------------------------------------------------
// JacobiSVD of a 3 x 3 (oracle/eig_model.hpp jacobi_svd3) and its solve() (oracle/refine.hpp find_opt_cc)
// ------------------------------------------------------------------------------------------------------------
template <class M>
class JacobiSVD {
  typedef typename M::Scalar T;
  orc::SVD3<T> s_;

 public:
  template <class O>
  JacobiSVD(const MatrixBase<O>& m, unsigned int = 0) : s_(orc::jacobi_svd3(m.to_m3())) {}
  Matrix<T, 3, 3> matrixU() const { return from_m3(s_.U); }
  Matrix<T, 3, 3> matrixV() const { return from_m3(s_.V); }
  Matrix<T, 3, 1> singularValues() const { return Matrix<T, 3, 1>(s_.s[0], s_.s[1], s_.s[2]); }
  // x = V diag(1 / s_i, i < rank) U^T b, rank = #{ s_i > max(s_0 diagSize eps, min) } (SVDBase::rank / _solve_impl)
  template <class O>
  Matrix<T, 3, 1> solve(const MatrixBase<O>& b) const {
    const T thr0 = s_.s[0] * (T(3) * std::numeric_limits<T>::epsilon());
    const T thr = thr0 > (std::numeric_limits<T>::min)() ? thr0 : (std::numeric_limits<T>::min)();
    int rank = 0;
    while (rank < 3 && s_.s[rank] > thr) ++rank;
    T tmp[3] = {T(0), T(0), T(0)};
    for (int k = 0; k < rank; ++k) {
      const T acc = orc::sum3(s_.U(0, k) * b.lin(0), s_.U(1, k) * b.lin(1), s_.U(2, k) * b.lin(2));
      tmp[k] = (T(1) / s_.s[k]) * acc;
    }
    Matrix<T, 3, 1> x;
    for (int r = 0; r < 3; ++r) {
      if (rank == 3)
        x(r) = orc::sum3(s_.V(r, 0) * tmp[0], s_.V(r, 1) * tmp[1], s_.V(r, 2) * tmp[2]);
      else {
        T acc = T(0);
        for (int k = 0; k < rank; ++k) acc += s_.V(r, k) * tmp[k];
        x(r) = acc;
      }
    }
    return x;
  }

 private:
  static Matrix<T, 3, 3> from_m3(const orc::M3<T>& m) {
    Matrix<T, 3, 3> r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r(i, j) = m(i, j);
    return r;
  }
};
template <class D>
JacobiSVD<Matrix<typename MatrixBase<D>::Scalar, 3, 3> > MatrixBase<D>::jacobiSvd(unsigned int flags) const {
  return JacobiSVD<Matrix<Scalar, 3, 3> >(*this, flags);
}

// ------------------------------------------------------------------------------------------------------------
// AngleAxis / Quaternion: what nl_2p and the Sophus stand-in need
// ------------------------------------------------------------------------------------------------------------
template <class T>
class AngleAxis {
  T angle_;
  Matrix<T, 3, 1> axis_;

 public:
  template <class O>
  AngleAxis(T angle, const MatrixBase<O>& axis) : angle_(angle), axis_(axis) {}
  T angle() const { return angle_; }
  const Matrix<T, 3, 1>& axis() const { return axis_; }
};

template <class T>
class Quaternion {
 public:
  orc::Quat<T> q;  // (x, y, z, w) like Eigen's coeffs()
  Quaternion() {}
  Quaternion(T w, T x, T y, T z) : q(w, x, y, z) {}
  explicit Quaternion(const orc::Quat<T>& o) : q(o) {}
  // Quaternion = AngleAxis: ha = 0.5 angle; w = cos(ha); vec = sin(ha) axis
  Quaternion(const AngleAxis<T>& aa) {
    const T ha = T(0.5) * aa.angle();
    q.w = std::cos(ha);
    const T s = std::sin(ha);
    q.x = s * aa.axis()(0);
    q.y = s * aa.axis()(1);
    q.z = s * aa.axis()(2);
  }
  T w() const { return q.w; }
  T x() const { return q.x; }
  T y() const { return q.y; }
  T z() const { return q.z; }
  T& w() { return q.w; }
  T& x() { return q.x; }
  T& y() { return q.y; }
  T& z() { return q.z; }
};

// Map<MatrixXf>(ptr, rows, cols) as Library.cpp:20-23 uses it: read once into a matrix (the reference copies it into
// a MatrixXf right away).
template <class M>
class Map : public M {
 public:
  Map(const typename M::Scalar* p, int r, int c) {
    this->resize(r, c);
    memcpy(this->data(), p, sizeof(typename M::Scalar) * (size_t)r * c);
  }
};

typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 3, 1> Vector3d;

}  // namespace Eigen

#endif  // ORACLE_REF_SHIM_EIGEN_SHIM_HPP_
