// rpe/types.hpp — the small dense-algebra vocabulary of the host API.
//
// The reference passes Eigen::Matrix<Tp,Dynamic,Dynamic> (3 x N, column-major) and returns
// Eigen::Matrix<Tp,3,1> / Sophus::SO3<Tp>. Eigen is an external dependency of the reference and is not
// part of this project; these few POD-like types carry the same data with the same memory layout
// (column-major, so `data()` of a 3 x N matrix is N contiguous xyz triples — exactly what the C-ABI
// takes). Every adapter constructor is a template over "anything with data()/rows()/cols()", so an
// Eigen matrix can be handed over unchanged when Eigen is available.
#ifndef RPE_TYPES_HPP_
#define RPE_TYPES_HPP_

#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <utility>
#include <vector>

namespace rpe {

const int Dynamic = -1;

template <class Tp>
struct Vec3 {
  Tp v[3];
  Vec3() : v{Tp(0), Tp(0), Tp(0)} {}
  Vec3(Tp x, Tp y, Tp z) : v{x, y, z} {}
  template <class Other>
  explicit Vec3(const Other& o) : v{(Tp)o[0], (Tp)o[1], (Tp)o[2]} {}
  static Vec3 Zero() { return Vec3(); }
  Tp& operator[](int i) { return v[i]; }
  const Tp& operator[](int i) const { return v[i]; }
  Tp& operator()(int i) { return v[i]; }
  const Tp& operator()(int i) const { return v[i]; }
  Tp* data() { return v; }
  const Tp* data() const { return v; }
  Tp dot(const Vec3& o) const { return v[0] * o.v[0] + (v[1] * o.v[1] + v[2] * o.v[2]); }
  Vec3 cross(const Vec3& o) const {
    return Vec3(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
  }
  Tp squaredNorm() const { return dot(*this); }
  Tp norm() const { return std::sqrt(squaredNorm()); }
  void normalize() {
    const Tp z = squaredNorm();
    if (z > Tp(0)) {
      const Tp n = std::sqrt(z);
      v[0] /= n;
      v[1] /= n;
      v[2] /= n;
    }
  }
  Vec3 normalized() const {
    Vec3 r = *this;
    r.normalize();
    return r;
  }
};
template <class Tp>
inline Vec3<Tp> operator+(const Vec3<Tp>& a, const Vec3<Tp>& b) {
  return Vec3<Tp>(a[0] + b[0], a[1] + b[1], a[2] + b[2]);
}
template <class Tp>
inline Vec3<Tp> operator-(const Vec3<Tp>& a, const Vec3<Tp>& b) {
  return Vec3<Tp>(a[0] - b[0], a[1] - b[1], a[2] - b[2]);
}
template <class Tp>
inline Vec3<Tp> operator-(const Vec3<Tp>& a) {
  return Vec3<Tp>(-a[0], -a[1], -a[2]);
}
template <class Tp>
inline Vec3<Tp> operator*(Tp s, const Vec3<Tp>& a) {
  return Vec3<Tp>(s * a[0], s * a[1], s * a[2]);
}
template <class Tp>
inline Vec3<Tp> operator*(const Vec3<Tp>& a, Tp s) {
  return Vec3<Tp>(a[0] * s, a[1] * s, a[2] * s);
}
template <class Tp>
inline Vec3<Tp> operator/(const Vec3<Tp>& a, Tp s) {
  return Vec3<Tp>(a[0] / s, a[1] / s, a[2] / s);
}
template <class Tp>
inline std::ostream& operator<<(std::ostream& os, const Vec3<Tp>& a) {
  return os << a[0] << " " << a[1] << " " << a[2];
}

// 3 x 3, stored row-major (m[r][c]); data() is only used internally.
template <class Tp>
struct Mat3 {
  Tp m[9];
  Mat3() {
    for (int i = 0; i < 9; ++i) m[i] = Tp(0);
  }
  static Mat3 Identity() {
    Mat3 a;
    a.m[0] = a.m[4] = a.m[8] = Tp(1);
    return a;
  }
  Tp& operator()(int r, int c) { return m[3 * r + c]; }
  const Tp& operator()(int r, int c) const { return m[3 * r + c]; }
  Mat3 transpose() const {
    Mat3 t;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) t(r, c) = (*this)(c, r);
    return t;
  }
  Tp determinant() const {
    const Tp* a = m;
    return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
  }
  Vec3<Tp> operator*(const Vec3<Tp>& x) const {
    return Vec3<Tp>(m[0] * x[0] + (m[1] * x[1] + m[2] * x[2]), m[3] * x[0] + (m[4] * x[1] + m[5] * x[2]),
                    m[6] * x[0] + (m[7] * x[1] + m[8] * x[2]));
  }
  Mat3 operator*(const Mat3& b) const {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r(i, j) = m[3 * i] * b(0, j) + (m[3 * i + 1] * b(1, j) + m[3 * i + 2] * b(2, j));
    return r;
  }
};

// Dynamic column-major matrix (the stand-in for Eigen::Matrix<Tp,Dynamic,Dynamic>).
template <class Tp>
class MatrixX {
 public:
  MatrixX() : rows_(0), cols_(0) {}
  MatrixX(int rows, int cols) : rows_(rows), cols_(cols), d_((size_t)rows * cols) {}
  // From any dense column-major matrix with data() / rows() / cols() (an Eigen::Matrix<Tp, Dynamic, Dynamic> in a
  // program written for the reference): lets adapter.setWeights(all_weights) / setInlier(inliers) take it as is.
  template <class M, class = decltype(static_cast<const Tp*>(std::declval<const M&>().data()), std::declval<const M&>().rows(),
                                      std::declval<const M&>().cols(), void())>
  MatrixX(const M& m) : rows_((int)m.rows()), cols_((int)m.cols()), d_(m.data(), m.data() + (size_t)m.rows() * (size_t)m.cols()) {}
  void resize(int rows, int cols) {
    rows_ = rows;
    cols_ = cols;
    d_.assign((size_t)rows * cols, Tp(0));
  }
  void setZero() { d_.assign(d_.size(), Tp(0)); }
  void setOnes() { d_.assign(d_.size(), Tp(1)); }
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  Tp* data() { return d_.data(); }
  const Tp* data() const { return d_.data(); }
  Tp& operator()(int r, int c) { return d_[(size_t)c * rows_ + r]; }
  const Tp& operator()(int r, int c) const { return d_[(size_t)c * rows_ + r]; }
  Tp& operator()(int i) { return d_[i]; }
  const Tp& operator()(int i) const { return d_[i]; }
  // column of a 3-row matrix as a value
  Vec3<Tp> col(int c) const {
    assert(rows_ == 3);
    return Vec3<Tp>(d_[(size_t)3 * c], d_[(size_t)3 * c + 1], d_[(size_t)3 * c + 2]);
  }
  void setCol(int c, const Vec3<Tp>& v) {
    assert(rows_ == 3);
    for (int r = 0; r < 3; ++r) d_[(size_t)3 * c + r] = v[r];
  }
  const Tp* colPtr(int c) const { return d_.data() + (size_t)c * rows_; }
  Tp* colPtr(int c) { return d_.data() + (size_t)c * rows_; }

 private:
  int rows_, cols_;
  std::vector<Tp> d_;
};

typedef MatrixX<short> MaskX;  // N x m inlier flags, column-major: col 0 = 2-D, col 1 = 3-D, col 2 = normal

// non-owning view of a column-major 3 x n array
template <class Tp>
struct View3 {
  const Tp* p;
  int n;
  View3() : p(nullptr), n(0) {}
  View3(const Tp* p_, int n_) : p(p_), n(n_) {}
  template <class M>
  static View3 of(const M& m) {
    assert(m.rows() == 3);
    return View3(m.data(), (int)m.cols());
  }
  Vec3<Tp> col(int i) const { return Vec3<Tp>(p[3 * (size_t)i], p[3 * (size_t)i + 1], p[3 * (size_t)i + 2]); }
};

}  // namespace rpe

#endif  // RPE_TYPES_HPP_
