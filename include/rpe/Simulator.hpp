// rpe/Simulator.hpp — mirrors /root/reference/pose/Simulator.hpp (same function names and argument order) on top
// of rpe/sim_core.hpp. The reference draws from hidden global state (::rand(), a global std::default_random_engine,
// :13-14); here the single global stream is rpe::sim::global_rng(), reseedable with rpe::sim::seed(s).
#ifndef RPE_SIMULATOR_HPP_
#define RPE_SIMULATOR_HPP_

#include "sim_core.hpp"
#include "so3.hpp"
#include "types.hpp"

namespace rpe {
namespace sim {
inline Rng& global_rng() {
  static Rng g(1);
  return g;
}
inline void seed(uint64_t s) { global_rng() = Rng(s); }
template <class T>
inline Pose<T> make_pose(const rpe::SO3<T>& R, const rpe::Vec3<T>& t) {
  Pose<T> p;
  for (int k = 0; k < 4; ++k) p.q[k] = R.unit_quaternion().c[k];
  for (int k = 0; k < 3; ++k) p.t[k] = t[k];
  return p;
}
}  // namespace sim
}  // namespace rpe

template <typename T>
rpe::Vec3<T> generate_random_translation_uniform(T size) {  // [reference :16-21]
  rpe::sim::Rng& g = rpe::sim::global_rng();
  return rpe::Vec3<T>(size * (T)g.uniform_pm1(), size * (T)g.uniform_pm1(), size * (T)g.uniform_pm1());
}

template <typename T>
rpe::SO3<T> generate_random_rotation(T max_angle_radian_, bool use_guassian_ = true) {  // [reference :23-83]
  rpe::Mat3<T> R;
  rpe::sim::random_rotation<T>(rpe::sim::global_rng(), max_angle_radian_, use_guassian_, R.m);
  T q[4];
  rpe::sim::R_to_quat<T>(R.m, q);
  return rpe::SO3<T>::fromRawQuaternion(q);
}

// a candidate point of the box around the frustum: x, y uniform in +-tan_fov * max_depth, z uniform in [min, max]
template <typename T>
rpe::Vec3<T> generate_a_random_point(T min_depth_, T max_depth_, T tan_fov_x, T tan_fov_y) {  // [reference :135-145]
  rpe::sim::Rng& rng = rpe::sim::global_rng();
  const T x = (T)rng.uniform_pm1() * tan_fov_x * max_depth_;
  const T y = (T)rng.uniform_pm1() * tan_fov_y * max_depth_;
  const T z = ((T)rng.uniform_pm1() + T(1)) / T(2) * (max_depth_ - min_depth_) + min_depth_;
  return rpe::Vec3<T>(x, y, z);
}

// pinhole projection with focal length f_ and the principal point at the origin: 2 x n pixel coordinates
template <typename T>
rpe::MatrixX<T> project_point_cloud(const rpe::MatrixX<T>& pt_c, T f_) {  // [reference :148-156]
  rpe::MatrixX<T> uv(2, pt_c.cols());
  for (int i = 0; i < pt_c.cols(); ++i) {
    uv(2 * i) = f_ * pt_c(3 * i) / pt_c(3 * i + 2);
    uv(2 * i + 1) = f_ * pt_c(3 * i + 1) / pt_c(3 * i + 2);
  }
  return uv;
}

template <typename T>
rpe::MatrixX<T> simulate_rand_point_cloud_in_frustum(int number_, T f_, T min_depth_, T max_depth_) {  // [:158-173]
  rpe::MatrixX<T> P(3, number_);
  rpe::sim::frustum_cloud<T>(rpe::sim::global_rng(), number_, f_, min_depth_, max_depth_, P.data());
  return P;
}

template <typename T>
void simulate_3d_3d_correspondences(const rpe::SO3<T>& R_cw_, const rpe::Vec3<T>& t_w_, int number_, T noise_,
                                    T outlier_ratio_, T min_depth_, T max_depth_, T f_, bool use_guassian_,
                                    rpe::MatrixX<T>* pQ_, rpe::MatrixX<T>* pP_gt = NULL,
                                    rpe::MatrixX<T>* p_all_weights_ = NULL) {  // [reference :268-314]
  pQ_->resize(3, number_);
  rpe::MatrixX<T> P(3, number_);
  rpe::sim::simulate_3d_3d<T>(rpe::sim::global_rng(), rpe::sim::make_pose(R_cw_, t_w_), number_, noise_, outlier_ratio_,
                              min_depth_, max_depth_, f_, use_guassian_, pQ_->data(), P.data(),
                              p_all_weights_ ? p_all_weights_->data() : (T*)0);
  if (pP_gt) *pP_gt = P;
}

template <typename T>
void simulate_2d_3d_correspondences(const rpe::SO3<T>& R_cw_, const rpe::Vec3<T>& t_w_, int number_, T noise_,
                                    T outlier_ratio_, T min_depth_, T max_depth_, T f_, bool use_guassian_,
                                    rpe::MatrixX<T>* pQ_, rpe::MatrixX<T>* pU_, rpe::MatrixX<T>* pP_gt = NULL,
                                    rpe::MatrixX<T>* p_all_weights_ = NULL) {  // [reference :175-233]
  pQ_->resize(3, number_);
  pU_->resize(3, number_);
  rpe::MatrixX<T> P(3, number_);
  rpe::sim::simulate_2d_3d<T>(rpe::sim::global_rng(), rpe::sim::make_pose(R_cw_, t_w_), number_, noise_, outlier_ratio_,
                              min_depth_, max_depth_, f_, use_guassian_, pQ_->data(), pU_->data(), P.data(),
                              p_all_weights_ ? p_all_weights_->data() : (T*)0);
  if (pP_gt) *pP_gt = P;
}

template <typename T>
void simulate_2d_3d_3d_correspondences(const rpe::SO3<T>& R_cw_, const rpe::Vec3<T>& t_w_, int number_, T noise_2d_,
                                       T noise_3d_, T outlier_ratio_, T min_depth_, T max_depth_, T f_, bool use_guassian_,
                                       rpe::MatrixX<T>* pQ_, rpe::MatrixX<T>* pU_, rpe::MatrixX<T>* pP_gt = NULL,
                                       rpe::MatrixX<T>* p_all_weights_ = NULL) {  // [reference :235-265]
  simulate_2d_3d_correspondences<T>(R_cw_, t_w_, number_, noise_2d_, outlier_ratio_, min_depth_, max_depth_, f_,
                                    use_guassian_, pQ_, pU_, pP_gt, p_all_weights_);
  rpe::sim::Rng& g = rpe::sim::global_rng();
  for (int i = 0; i < number_; i++) {
    T rv[3];
    rpe::sim::noise_vec<T>(g, use_guassian_, 3, rv);
    if (p_all_weights_) (*p_all_weights_)(i, 1) = T(1.) / std::sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
    for (int r = 0; r < 3; ++r) (*pQ_)(r, i) += noise_3d_ * rv[r];
  }
}

template <typename T>
void simulate_nl_nl_correspondences(const rpe::SO3<T>& R_cw_, int number_, T noise_nl_, T outlier_ratio_nl_,
                                    bool use_guassian_, rpe::MatrixX<T>* pM_, rpe::MatrixX<T>* pN_,
                                    rpe::MatrixX<T>* pN_gt = NULL, rpe::MatrixX<T>* p_all_weights_ = NULL) {  // [:85-130]
  pM_->resize(3, number_);
  pN_->resize(3, number_);
  rpe::MatrixX<T> Ngt(3, number_);
  rpe::sim::simulate_nl_nl<T>(rpe::sim::global_rng(), rpe::sim::make_pose(R_cw_, rpe::Vec3<T>()), number_, noise_nl_,
                              outlier_ratio_nl_, use_guassian_, pM_->data(), pN_->data(), Ngt.data(),
                              p_all_weights_ ? p_all_weights_->data() : (T*)0);
  if (pN_gt) *pN_gt = Ngt;
}

template <typename T>
void simulate_2d_3d_nl_correspondences(const rpe::SO3<T>& R_cw_, const rpe::Vec3<T>& t_w_, int number_, T n2D_, T or_2D_,
                                       T n3D_, T or_3D_, T nNl_, T or_Nl_, T min_depth_, T max_depth_, T f_,
                                       bool use_guassian_, rpe::MatrixX<T>* pQ_, rpe::MatrixX<T>* pM_, rpe::MatrixX<T>* pP_,
                                       rpe::MatrixX<T>* pN_, rpe::MatrixX<T>* pU_,
                                       rpe::MatrixX<T>* p_all_weights_ = NULL) {  // [reference :316-367]
  pQ_->resize(3, number_);
  pM_->resize(3, number_);
  pP_->resize(3, number_);
  pN_->resize(3, number_);
  pU_->resize(3, number_);
  rpe::MatrixX<T> w(number_, 3);
  rpe::sim::simulate_2d_3d_nl<T>(rpe::sim::global_rng(), rpe::sim::make_pose(R_cw_, t_w_), number_, n2D_, or_2D_, n3D_,
                                 or_3D_, nNl_, or_Nl_, min_depth_, max_depth_, f_, use_guassian_, pQ_->data(), pM_->data(),
                                 pP_->data(), pN_->data(), pU_->data(), w.data());
  if (p_all_weights_) *p_all_weights_ = w;
}

template <typename T>
T lateral_noise_kinect(T theta_, T z_, T f_) {  // [reference :368-377]
  return rpe::sim::kinect_lateral_sigma<T>(theta_, z_, f_);
}
template <typename T>
T axial_noise_kinect(T theta_, T z_) {  // [reference :379-387]
  return rpe::sim::kinect_axial_sigma<T>(theta_, z_);
}

template <typename T>
void simulate_kinect_2d_3d_nl_correspondences(const rpe::SO3<T>& R_cw_, const rpe::Vec3<T>& t_w_, int number_, T noise_2d_,
                                              T outlier_ratio_2d_, T outlier_ratio_3d_, T noise_nl_, T outlier_ratio_nl_,
                                              T min_depth_, T max_depth_, T f_, rpe::MatrixX<T>* p_pt_w_,
                                              rpe::MatrixX<T>* p_nl_w_, rpe::MatrixX<T>* p_pt_c_, rpe::MatrixX<T>* p_nl_c_,
                                              rpe::MatrixX<T>* p_bv_, rpe::MatrixX<T>* p_weights_ = NULL) {  // [reference :389-436]
  p_pt_w_->resize(3, number_);
  p_nl_w_->resize(3, number_);
  p_pt_c_->resize(3, number_);
  p_nl_c_->resize(3, number_);
  p_bv_->resize(3, number_);
  rpe::MatrixX<T> w(number_, 3);
  rpe::sim::simulate_kinect_2d_3d_nl<T>(rpe::sim::global_rng(), rpe::sim::make_pose(R_cw_, t_w_), number_, noise_2d_,
                                        outlier_ratio_2d_, outlier_ratio_3d_, noise_nl_, outlier_ratio_nl_, min_depth_,
                                        max_depth_, f_, p_pt_w_->data(), p_nl_w_->data(), p_pt_c_->data(), p_nl_c_->data(),
                                        p_bv_->data(), w.data());
  if (p_weights_) *p_weights_ = w;
}

#endif  // RPE_SIMULATOR_HPP_
