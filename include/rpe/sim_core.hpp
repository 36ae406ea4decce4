// rpe/sim_core.hpp — synthetic correspondence generators on raw column-major arrays.
//
// Restates the generators of /root/reference/pose/Simulator.hpp with an explicit 64-bit seed.
// The reference draws from three hidden global sources (::rand() through Eigen::Random,
// std::default_random_engine + std::normal_distribution, and RandomElements -> ::rand()); only
// the DISTRIBUTIONS are part of its contract (the same arrays feed every estimator it compares),
// so this file keeps the distributions and the order of operations but draws from one seeded
// xoshiro256** stream:
//   generate_random_translation_uniform   Simulator.hpp:16-21
//   generate_random_rotation              :23-83   (R = Rz * Ry * Rx, y-angle halved, clamped)
//   simulate_rand_point_cloud_in_frustum  :158-173 (640x480, principal point centred)
//   simulate_3d_3d_correspondences        :268-314
//   simulate_2d_3d_correspondences        :175-233
//   simulate_nl_nl_correspondences        :85-130  (incl. the quirk that normal outliers overwrite
//                                                   columns 0..out-1 instead of the drawn indices, :114-120)
//   simulate_2d_3d_nl_correspondences     :316-367
// All matrices are 3 x n column-major (n contiguous xyz triples), weights are n x 3 column-major.
#ifndef RPE_SIM_CORE_HPP_
#define RPE_SIM_CORE_HPP_

#include <stdint.h>

#include <cmath>
#include <vector>

namespace rpe {
namespace sim {

class Rng {
 public:
  explicit Rng(uint64_t seed) {
    uint64_t z = seed;
    for (int i = 0; i < 4; ++i) {  // splitmix64 expansion of the seed
      z += 0x9e3779b97f4a7c15ULL;
      uint64_t x = z;
      x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
      x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
      s_[i] = x ^ (x >> 31);
    }
    have_spare_ = false;
    spare_ = 0.0;
  }
  uint64_t next_u64() {
    const uint64_t result = rotl(s_[1] * 5, 7) * 9;
    const uint64_t t = s_[1] << 17;
    s_[2] ^= s_[0];
    s_[3] ^= s_[1];
    s_[1] ^= s_[2];
    s_[0] ^= s_[3];
    s_[2] ^= t;
    s_[3] = rotl(s_[3], 45);
    return result;
  }
  double unit() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double uniform_pm1() { return 2.0 * unit() - 1.0; }                               // Eigen::Random range
  double normal() {                                                                  // N(0,1), polar Box-Muller
    if (have_spare_) {
      have_spare_ = false;
      return spare_;
    }
    double u, v, s;
    do {
      u = uniform_pm1();
      v = uniform_pm1();
      s = u * u + v * v;
    } while (s >= 1.0 || s == 0.0);
    const double k = std::sqrt(-2.0 * std::log(s) / s);
    spare_ = v * k;
    have_spare_ = true;
    return u * k;
  }
  int below(int bound) { return (int)(next_u64() % (uint64_t)bound); }

 private:
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t s_[4];
  bool have_spare_;
  double spare_;
};

// m distinct indices of [0,n): the RandomElements draw (Utility.hpp:138-155) on this stream.
inline void pick_distinct(Rng& rng, int n, int m, std::vector<int>* out) {
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  out->clear();
  for (int j = n - 1; j > n - m - 1 && j >= 0; --j) {
    const int r = rng.below(j + 1);
    const int tmp = perm[r];
    perm[r] = perm[j];
    perm[j] = tmp;
    out->push_back(tmp);
  }
}

template <class T>
struct Pose {
  T q[4];  // x, y, z, w
  T t[3];
};

template <class T>
inline void quat_to_R(const T* q, T* R /*row-major*/) {
  const T x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z);
  R[1] = 2 * (x * y - z * w);
  R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);
  R[4] = 1 - 2 * (x * x + z * z);
  R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);
  R[7] = 2 * (y * z + x * w);
  R[8] = 1 - 2 * (x * x + y * y);
}
template <class T>
inline void R_to_quat(const T* R, T* q) {
  const T tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    T s = std::sqrt(tr + T(1));
    q[3] = T(0.5) * s;
    s = T(0.5) / s;
    q[0] = (R[7] - R[5]) * s;
    q[1] = (R[2] - R[6]) * s;
    q[2] = (R[3] - R[1]) * s;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    T s = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + T(1));
    T v[3];
    v[i] = T(0.5) * s;
    s = T(0.5) / s;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * s;
    v[j] = (R[3 * j + i] + R[3 * i + j]) * s;
    v[k] = (R[3 * k + i] + R[3 * i + k]) * s;
    q[0] = v[0];
    q[1] = v[1];
    q[2] = v[2];
  }
  const T n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int c = 0; c < 4; ++c) q[c] /= n;
}

// generate_random_rotation (Simulator.hpp:23-83)
template <class T>
inline void random_rotation(Rng& rng, T max_angle, bool gaussian, T* R /*row-major*/) {
  double rv[3];
  for (int i = 0; i < 3; ++i) rv[i] = gaussian ? rng.normal() : rng.uniform_pm1();
  const double pi = 3.14159265358979323846;
  rv[0] = max_angle * rv[0];
  rv[1] = max_angle * rv[1] * 0.5;
  rv[2] = max_angle * rv[2];
  rv[0] = rv[0] > pi ? pi : (rv[0] < -pi ? -pi : rv[0]);
  rv[1] = rv[1] > pi / 2 ? pi / 2 : (rv[1] < -pi / 2 ? -pi / 2 : rv[1]);
  rv[2] = rv[2] > pi ? pi : (rv[2] < -pi ? -pi : rv[2]);
  const double cx = std::cos(rv[0]), sx = std::sin(rv[0]);
  const double cy = std::cos(rv[1]), sy = std::sin(rv[1]);
  const double cz = std::cos(rv[2]), sz = std::sin(rv[2]);
  // Rz * Ry * Rx
  const double M[9] = {cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx,
                       sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx,
                       -sy,     cy * sx,                cy * cx};
  for (int i = 0; i < 9; ++i) R[i] = (T)M[i];
}

template <class T>
inline Pose<T> random_pose(Rng& rng, T max_angle, T t_size) {
  Pose<T> p;
  for (int i = 0; i < 3; ++i) p.t[i] = t_size * (T)rng.uniform_pm1();  // :16-21
  T R[9];
  random_rotation<T>(rng, max_angle, false, R);
  R_to_quat(R, p.q);
  return p;
}

// one point inside the 640x480 viewing frustum (Simulator.hpp:135-145, 158-173)
template <class T>
inline void frustum_point(Rng& rng, T f, T min_depth, T max_depth, T* P) {
  const T tan_x = T(320.) / f, tan_y = T(240.) / f;
  for (;;) {
    const T x = (T)rng.uniform_pm1() * tan_x * max_depth;
    const T y = (T)rng.uniform_pm1() * tan_y * max_depth;
    const T z = ((T)rng.uniform_pm1() + T(1)) / T(2) * (max_depth - min_depth) + min_depth;
    if (std::fabs(x / z) < tan_x && std::fabs(y / z) < tan_y) {
      P[0] = x;
      P[1] = y;
      P[2] = z;
      return;
    }
  }
}
template <class T>
inline void frustum_cloud(Rng& rng, int n, T f, T min_depth, T max_depth, T* P) {
  for (int i = 0; i < n; ++i) frustum_point(rng, f, min_depth, max_depth, P + 3 * i);
}

// world point of a camera point: R_cw^-1 (P - t)
template <class T>
inline void cam_to_world(const T* R, const T* t, const T* P, T* Q) {
  const T d[3] = {P[0] - t[0], P[1] - t[1], P[2] - t[2]};
  for (int r = 0; r < 3; ++r) Q[r] = R[r] * d[0] + R[3 + r] * d[1] + R[6 + r] * d[2];  // R^T d
}

template <class T>
inline void noise_vec(Rng& rng, bool gaussian, int dim, T* rv) {
  for (int i = 0; i < dim; ++i) rv[i] = (T)(gaussian ? rng.normal() : rng.uniform_pm1());
}

// simulate_3d_3d_correspondences (:268-314). Q: noisy world points (outliers are raw frustum points),
// P_gt: clean camera points, weights3 (optional, n x 3 col-major): column 1 = 1/|noise draw|.
template <class T>
inline void simulate_3d_3d(Rng& rng, const Pose<T>& pose, int n, T noise, T outlier_ratio, T min_depth, T max_depth, T f,
                           bool gaussian, T* Q, T* P_gt, T* weights3) {
  T R[9];
  quat_to_R(pose.q, R);
  frustum_cloud(rng, n, f, min_depth, max_depth, P_gt);
  for (int i = 0; i < n; ++i) cam_to_world(R, pose.t, P_gt + 3 * i, Q + 3 * i);
  for (int i = 0; i < n; ++i) {
    T rv[3];
    noise_vec(rng, gaussian, 3, rv);
    if (weights3) weights3[n + i] = T(1) / std::sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
    for (int r = 0; r < 3; ++r) Q[3 * i + r] += noise * rv[r];
  }
  const int out = (int)(outlier_ratio * n + .5);
  std::vector<T> outp((size_t)3 * (out > 0 ? out : 1));
  frustum_cloud(rng, out, f, min_depth, max_depth, outp.data());
  std::vector<int> idx;
  pick_distinct(rng, n, out, &idx);
  for (int i = 0; i < out; ++i)
    for (int r = 0; r < 3; ++r) Q[3 * idx[i] + r] = outp[3 * i + r];
}

// simulate_2d_3d_correspondences (:175-233). U: unit bearing vectors; weights column 0.
template <class T>
inline void simulate_2d_3d(Rng& rng, const Pose<T>& pose, int n, T noise_px, T outlier_ratio, T min_depth, T max_depth,
                           T f, bool gaussian, T* Q, T* U, T* P_gt, T* weights3) {
  T R[9];
  quat_to_R(pose.q, R);
  std::vector<T> own;
  if (!P_gt) {
    own.resize((size_t)3 * n);
    P_gt = own.data();
  }
  frustum_cloud(rng, n, f, min_depth, max_depth, P_gt);
  std::vector<T> kp((size_t)2 * n);
  for (int i = 0; i < n; ++i) {
    kp[2 * i] = f * P_gt[3 * i] / P_gt[3 * i + 2];
    kp[2 * i + 1] = f * P_gt[3 * i + 1] / P_gt[3 * i + 2];
  }
  for (int i = 0; i < n; ++i) cam_to_world(R, pose.t, P_gt + 3 * i, Q + 3 * i);
  for (int i = 0; i < n; ++i) {
    T rv[2];
    noise_vec(rng, gaussian, 2, rv);
    if (weights3) weights3[i] = T(1) / std::sqrt(rv[0] * rv[0] + rv[1] * rv[1]);
    kp[2 * i] += noise_px * rv[0];
    kp[2 * i + 1] += noise_px * rv[1];
  }
  const int out = (int)(outlier_ratio * n + .5);
  std::vector<T> outp((size_t)3 * (out > 0 ? out : 1));
  frustum_cloud(rng, out, f, min_depth, max_depth, outp.data());
  std::vector<int> idx;
  pick_distinct(rng, n, out, &idx);
  for (int i = 0; i < out; ++i) {
    kp[2 * idx[i]] = f * outp[3 * i] / outp[3 * i + 2];
    kp[2 * idx[i] + 1] = f * outp[3 * i + 1] / outp[3 * i + 2];
  }
  for (int i = 0; i < n; ++i) {
    const T x = kp[2 * i], y = kp[2 * i + 1];
    const T nn = std::sqrt(x * x + y * y + f * f);
    U[3 * i] = x / nn;
    U[3 * i + 1] = y / nn;
    U[3 * i + 2] = f / nn;
  }
}

// simulate_nl_nl_correspondences (:85-130). M: world normals, N: noisy camera normals; weights column 2.
template <class T>
inline void simulate_nl_nl(Rng& rng, const Pose<T>& pose, int n, T noise_nl, T outlier_ratio, bool gaussian, T* M, T* N,
                           T* N_gt_out, T* weights3) {
  T R[9];
  quat_to_R(pose.q, R);
  auto unit = [](T* v) {
    const T nn = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] /= nn;
    v[1] /= nn;
    v[2] /= nn;
  };
  auto facing_draw = [&](T max_angle, bool g, const T* src, T* dst) {
    T Rr[9];
    random_rotation<T>(rng, max_angle, g, Rr);
    for (int r = 0; r < 3; ++r) dst[r] = Rr[3 * r] * src[0] + Rr[3 * r + 1] * src[1] + Rr[3 * r + 2] * src[2];
    unit(dst);
  };
  const T back[3] = {0, 0, -1};
  const T half_pi = T(3.14159265358979323846 / 2.);
  for (int i = 0; i < n; ++i) {
    T ngt[3];
    do {
      facing_draw(half_pi, false, back, ngt);
      for (int r = 0; r < 3; ++r) M[3 * i + r] = R[r] * ngt[0] + R[3 + r] * ngt[1] + R[6 + r] * ngt[2];  // R^T n
      unit(M + 3 * i);
      facing_draw(noise_nl, gaussian, ngt, N + 3 * i);
    } while (N[3 * i + 2] > 0);  // acos(n_z) < pi/2 <=> n_z > 0: keep normals that face the camera
    if (N_gt_out)
      for (int r = 0; r < 3; ++r) N_gt_out[3 * i + r] = ngt[r];
    if (weights3) weights3[2 * n + i] = N[3 * i] * ngt[0] + N[3 * i + 1] * ngt[1] + N[3 * i + 2] * ngt[2];
  }
  const int out = (int)(outlier_ratio * n + T(.5));
  std::vector<int> idx;
  pick_distinct(rng, n, out, &idx);  // drawn but unused, as in the reference (:112-114)
  for (int i = 0; i < out; ++i) {
    do {
      facing_draw(half_pi, false, back, N + 3 * i);  // column i, not idx[i] (:117)
    } while (N[3 * i + 2] > 0);
  }
}

// simulate_2d_3d_nl_correspondences (:316-367)
template <class T>
inline void simulate_2d_3d_nl(Rng& rng, const Pose<T>& pose, int n, T n2d, T or2d, T n3d, T or3d, T nnl, T ornl,
                              T min_depth, T max_depth, T f, bool gaussian, T* Q, T* M, T* P, T* N, T* U, T* weights3) {
  std::vector<T> P_gt((size_t)3 * n);
  simulate_2d_3d(rng, pose, n, n2d, or2d, min_depth, max_depth, f, gaussian, Q, U, P_gt.data(), weights3);
  simulate_nl_nl(rng, pose, n, nnl, ornl, true, M, N, (T*)0, weights3);
  for (int i = 0; i < n; ++i) {
    T rv[3];
    noise_vec(rng, gaussian, 3, rv);
    if (weights3) weights3[n + i] = T(1) / std::sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
    for (int r = 0; r < 3; ++r) P[3 * i + r] = P_gt[3 * i + r] + n3d * rv[r];
  }
  const int out = (int)(or3d * n + .5);
  std::vector<int> idx;
  pick_distinct(rng, n, out, &idx);
  std::vector<T> outp((size_t)3 * (out > 0 ? out : 1));
  frustum_cloud(rng, out, f, min_depth, max_depth, outp.data());
  for (int i = 0; i < out; ++i)
    for (int r = 0; r < 3; ++r) P[3 * idx[i] + r] = outp[3 * i + r];
}

// Kinect depth-noise model of Nguyen, Izadi & Lovell (3DIMPVT 2012), the functions the reference keeps at
// Simulator.hpp:368-386: lateral sigma in metres for a surface seen under angle theta at depth z with focal length f,
// and axial sigma (quadratic in depth, with the grazing-angle term beyond 60 degrees).
template <class T>
inline T kinect_lateral_sigma(T theta, T z, T f) {
  const T half_pi = T(3.14159265358979323846 / 2.);
  const T px = T(.8) + T(.035) * theta / (half_pi - theta);  // pixels
  return px * z / f;
}
template <class T>
inline T kinect_axial_sigma(T theta, T z) {
  const T half_pi = T(3.14159265358979323846 / 2.);
  const T dz = z - T(0.4);
  T s = T(.0012) + T(.0019) * dz * dz;
  if (std::fabs(theta) > T(3.14159265358979323846 / 3.)) {
    const T g = half_pi - theta;
    s += T(.0001) * theta * theta / std::sqrt(z) / g / g;
  }
  return s;
}

// simulate_kinect_2d_3d_nl_correspondences (:389-436): 2-D and normal channels as in simulate_2d_3d_nl, camera points
// perturbed by the Kinect model (lateral on x, y; axial on z; theta = angle between the true normal and the optical
// axis towards the camera), weight column 1 = sigma_axial(0, min_depth) / sigma_axial, 3-D outliers from the frustum.
template <class T>
inline void simulate_kinect_2d_3d_nl(Rng& rng, const Pose<T>& pose, int n, T n2d, T or2d, T or3d, T nnl, T ornl, T min_depth,
                                     T max_depth, T f, T* Q, T* M, T* P, T* N, T* U, T* weights3) {
  std::vector<T> P_gt((size_t)3 * n), N_gt((size_t)3 * n);
  simulate_2d_3d(rng, pose, n, n2d, or2d, min_depth, max_depth, f, true, Q, U, P_gt.data(), weights3);
  simulate_nl_nl(rng, pose, n, nnl, ornl, true, M, N, N_gt.data(), weights3);
  const T sigma_min = kinect_axial_sigma<T>(T(0), min_depth);
  for (int i = 0; i < n; ++i) {
    T c = -N_gt[3 * i + 2];  // n . (0, 0, -1)
    c = c > T(1) ? T(1) : (c < T(-1) ? T(-1) : c);
    const T theta = std::acos(c);
    const T z = P_gt[3 * i + 2];
    const T sl = kinect_lateral_sigma<T>(theta, z, f), sa = kinect_axial_sigma<T>(theta, z);
    T rv[3];
    noise_vec(rng, true, 3, rv);
    P[3 * i] = P_gt[3 * i] + sl * rv[0];
    P[3 * i + 1] = P_gt[3 * i + 1] + sl * rv[1];
    P[3 * i + 2] = P_gt[3 * i + 2] + sa * rv[2];
    if (weights3) weights3[n + i] = sigma_min / sa;
  }
  const int out = (int)(or3d * n + .5);
  std::vector<int> idx;
  pick_distinct(rng, n, out, &idx);
  std::vector<T> outp((size_t)3 * (out > 0 ? out : 1));
  frustum_cloud(rng, out, f, min_depth, max_depth, outp.data());
  for (int i = 0; i < out; ++i)
    for (int r = 0; r < 3; ++r) P[3 * idx[i] + r] = outp[3 * i + r];
}

}  // namespace sim
}  // namespace rpe

#endif  // RPE_SIM_CORE_HPP_
