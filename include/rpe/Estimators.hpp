// rpe/Estimators.hpp — the marshalling layer shared by every estimator template: adapter -> C-ABI -> adapter.
//
// What the reference does inside each *_ransac loop on the CPU (sample, minimal solve, score all N, keep
// best, shrink Iter: e.g. AbsoluteOrientation.hpp:101-156) happens on the GPU behind rpe_ransac(); this file
// only (1) draws the same sample stream the reference would draw (RandomElements / ProsacSampler on ::rand()),
// (2) moves the adapter's arrays to the device, (3) writes the result back through the adapter's own setters
// (setMaxVotes / setRcw / sett / setInlier / cvtInlier), exactly the observable state the reference leaves.
#ifndef RPE_ESTIMATORS_HPP_
#define RPE_ESTIMATORS_HPP_

#include <limits>
#include <stdint.h>

#include <cmath>
#include <vector>

#include "../rpe_c_api.h"
#include "PoseAdapterBase.hpp"
#include "Utility.hpp"
#include "session.hpp"

namespace rpe {
namespace detail {

inline int mask_cols(int method) {
  return (method == RPE_KNEIP || method == RPE_KNEIP_QUAT) ? 1 : ((method == RPE_SHINJI || method == RPE_SHINJI_KNEIP) ? 2 : 3);
}

// H rows of `m` distinct indices, the draws of `RandomElements<int> re(n); re.run(m, &sel)` per iteration.
inline void draw_ransac_table(int n, int m, int H, std::vector<int32_t>* table, RandSource* src = nullptr) {
  table->assign((size_t)H * 4, -1);
  RandomElements<int> re(n, src);
  std::vector<int> sel;
  for (int h = 0; h < H; ++h) {
    re.run(m, &sel);
    for (int k = 0; k < m; ++k) (*table)[(size_t)4 * h + k] = sel[k];
  }
}

// H rows of ProsacSampler draws mapped through adapter.getSortedIdx (e.g. AbsoluteOrientation.hpp:221-229).
// The reference's sampler can emit the index n == N (Utility.hpp:238), one past the last correspondence; it is
// clamped to N-1 here instead of being read out of bounds.
template <class Tp, class Adapter>
inline void draw_prosac_table(Adapter& adapter, int m, int H, std::vector<int32_t>* table, RandSource* src = nullptr) {
  const int n = adapter.getNumberCorrespondences();
  table->assign((size_t)H * 4, -1);
  adapter.sortIdx();
  ProsacSampler<Tp> ps(m, n, src);
  for (int h = 0; h < H; ++h) {
    std::vector<int> sel;
    ps.sample(&sel);
    adapter.getSortedIdx(sel);
    for (int k = 0; k < m; ++k) (*table)[(size_t)4 * h + k] = sel[k] < n ? sel[k] : n - 1;
  }
}

// ---- row producers: one sample row per RANSAC iteration, drawn in order on demand -------------------------------
// (the GPU asks for the rows of one pass at a time — rpe_ransac_stream — so a caller whose Iter is 100 000 does not pay
// for 100 000 draws when the adaptive bound stops the loop after a few hundred iterations)
class RansacRows {  // `RandomElements<int> re(n); re.run(m, &sel)` per iteration (e.g. AbsoluteOrientation.hpp:111,124)
 public:
  RansacRows(int n, int m, RandSource* src = nullptr) : m_(m), re_(n, src) {}
  void restart() {}  // RandomElements leaves the identity permutation behind after every run
  void row(int32_t* out) {
    re_.run(m_, &sel_);
    for (int k = 0; k < 4; ++k) out[k] = k < (int)sel_.size() ? sel_[k] : -1;
  }

 private:
  int m_;
  RandomElements<int> re_;
  std::vector<int> sel_;
};
// ProsacSampler draws mapped through adapter.getSortedIdx (e.g. AbsoluteOrientation.hpp:221-229); index n == N is clamped
template <class Tp, class Adapter>
class ProsacRows {
 public:
  ProsacRows(Adapter& adapter, int m, RandSource* src = nullptr)
      : adapter_(adapter), m_(m), n_(adapter.getNumberCorrespondences()), src_(src), ps_(m, n_, src) {
    adapter_.sortIdx();
  }
  void restart() { ps_ = ProsacSampler<Tp>(m_, n_, src_); }
  void row(int32_t* out) {
    std::vector<int> sel;
    ps_.sample(&sel);
    adapter_.getSortedIdx(sel);
    for (int k = 0; k < 4; ++k) out[k] = k < m_ ? (sel[k] < n_ ? sel[k] : n_ - 1) : -1;
  }

 private:
  Adapter& adapter_;
  int m_, n_;
  RandSource* src_;
  ProsacSampler<Tp> ps_;
};
template <class Rows>
inline int rows_callback(void* user, int /*first_iteration*/, int count, int32_t* out) {
  Rows* rows = static_cast<Rows*>(user);
  for (int i = 0; i < count; ++i) rows->row(out + 4 * (size_t)i);
  return 0;
}

template <class Tp>
inline void upload(Session& s, const PoseAdapterBase<Tp>& adapter) {
  const Tp *bv, *xc, *nc, *xw, *nw;
  adapter.rpeArrays(&bv, &xc, &nc, &xw, &nw);
  s.upload(bv, xc, nc, xw, nw, adapter.getNumberCorrespondences());  // double: binary64 path, see Session::upload
}

// qd / td carry the accepted hypothesis in binary64 on the binary64 path and the widened float pose otherwise
template <class Tp>
inline void pose_from_result(const rpe_result& r, SO3<Tp>* R, Vec3<Tp>* t) {
  const Tp q[4] = {(Tp)r.qd[0], (Tp)r.qd[1], (Tp)r.qd[2], (Tp)r.qd[3]};
  *R = SO3<Tp>::fromRawQuaternion(q);
  *t = Vec3<Tp>((Tp)r.td[0], (Tp)r.td[1], (Tp)r.td[2]);
}

// the RANSAC call itself: binary64 thresholds for Tp = double, the float entry point otherwise
template <class Tp>
struct RansacCall {
  static int run(rpe_ctx* ctx, int method, rpe_sample_fn fn, void* user, int H, Tp thr3d, Tp cos2, Tp cosn, Tp conf,
                 rpe_result* out, int16_t* mask) {
    return rpe_ransac_stream(ctx, method, fn, user, H, (float)thr3d, (float)cos2, (float)cosn, (float)conf, out, mask);
  }
};
template <>
struct RansacCall<double> {
  static int run(rpe_ctx* ctx, int method, rpe_sample_fn fn, void* user, int H, double thr3d, double cos2, double cosn,
                 double conf, rpe_result* out, int16_t* mask) {
    return rpe_ransac_f64(ctx, method, nullptr, fn, user, H, thr3d, cos2, cosn, conf, out, mask);
  }
};

inline int method_slots(int method) {
  return (method == RPE_SHINJI_KNEIP || method == RPE_NL_SHINJI) ? 2 : (method == RPE_NL_SHINJI_KNEIP ? 3 : 1);
}

// The common body of every *_ransac / *_prosac template.
// Random stream: the reference draws inside its loop and stops drawing when the loop ends, i.e. after
// max(Iter_final, i_winner + 1) iterations (all of them when nothing was accepted). Here whole passes are drawn ahead,
// so the generator is snapshotted first and, once the result is known, rewound and advanced by exactly the rows the
// reference would have drawn. (A custom RandSource without save/load keeps the extra draws.)
template <class Tp, class Adapter, class Rows>
inline rpe_result run_ransac(Adapter& adapter, int method, Rows& rows, RandSource* src, Tp thr3d, Tp cos_thr2d, Tp cos_thrN,
                             int& Iter, Tp confidence) {
  Session& s = Session::local();
  upload<Tp>(s, adapter);
  const int n = adapter.getNumberCorrespondences();
  const int cols = mask_cols(method);
  std::vector<int16_t> mask((size_t)n * cols);
  rpe_result res;
  adapter.setMaxVotes(-1);
  const int iter0 = Iter;
  RandState snapshot;
  const bool saved = rand_save(src, &snapshot);
  s.check(RansacCall<Tp>::run(s.ctx(), method, &rows_callback<Rows>, &rows, Iter, thr3d, cos_thr2d, cos_thrN, confidence, &res,
                              mask.data()),
          "rpe_ransac");
  if (saved && rand_load(src, snapshot)) {
    int executed = iter0;
    if (res.winner >= 0) {
      executed = res.winner / method_slots(method) + 1;
      if (res.iter_final > executed) executed = res.iter_final;
      if (executed > iter0) executed = iter0;
    }
    rows.restart();
    int32_t scratch[4];
    for (int i = 0; i < executed; ++i) rows.row(scratch);
  }
  if (res.winner >= 0) {
    adapter.setMaxVotes(res.max_votes);
    SO3<Tp> R;
    Vec3<Tp> t;
    pose_from_result<Tp>(res, &R, &t);
    adapter.setRcw(R);
    adapter.sett(t);
    MaskX m(n, cols);
    for (size_t i = 0; i < mask.size(); ++i) m(i) = mask[i];
    adapter.setInlier(m);
  }
  Iter = res.iter_final;
  adapter.rpeSetStateToken(s.bump());  // the device now mirrors this adapter's pose and flags
  return res;
}

// Make sure the device holds the adapter's arrays, pose and inlier flags (after user-side setInlier/setRcw the
// token no longer matches and everything is sent again).
template <class Tp, class Adapter>
inline void sync_state(Session& s, Adapter& adapter) {
  if (adapter.rpeStateToken() != 0 && adapter.rpeStateToken() == s.token()) return;
  upload<Tp>(s, adapter);
  std::vector<short> flags;
  const int cols = adapter.rpeMask(&flags);
  if (cols > 0) s.check(rpe_set_mask(s.ctx(), flags.data(), cols), "rpe_set_mask");
  const Quaternion<Tp>& q = adapter.getRcw().unit_quaternion();
  const float qf[4] = {(float)q.x(), (float)q.y(), (float)q.z(), (float)q.w()};
  const Vec3<Tp> t = adapter.gettw();
  const float tf[3] = {(float)t[0], (float)t[1], (float)t[2]};
  s.check(rpe_set_pose(s.ctx(), qf, tf, 0), "rpe_set_pose");
}

template <class Tp, class Adapter>
inline rpe_result run_refit(Adapter& adapter, int kind, const float* weights, int max_iters) {
  Session& s = Session::local();
  sync_state<Tp>(s, adapter);
  rpe_result res;
  s.check(rpe_refit(s.ctx(), kind, weights, max_iters, &res), "rpe_refit");
  if (res.refit_ok) {
    SO3<Tp> R;
    Vec3<Tp> t;
    pose_from_result<Tp>(res, &R, &t);
    adapter.setRcw(R);
    adapter.sett(t);
  }
  adapter.rpeSetStateToken(s.bump());
  return res;
}

}  // namespace detail
}  // namespace rpe

// RANSACUpdateNumIters — /root/reference/pose/P3P.hpp:296-318 (same name, same arguments). For float and double it
// is the bit-reproducible rule the device replays (rpe/ransac_rule.h); other scalar types use libm like the reference.
#include "ransac_rule.h"
template <typename T>
int RANSACUpdateNumIters(T p, T ep, const int modelPoints, const int maxIters) {
  p = std::max(p, T(0.));
  p = std::min(p, T(1.));
  ep = std::max(ep, T(0.));
  ep = std::min(ep, T(1.));
  T num = std::max(T(1. - p), std::numeric_limits<T>::epsilon());
  T denom = T(1.) - std::pow(T(1. - ep), modelPoints);
  if (denom < std::numeric_limits<T>::epsilon()) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= maxIters * (-denom) ? maxIters : int(num / denom + 0.5f);
}
template <>
inline int RANSACUpdateNumIters<float>(float p, float ep, const int modelPoints, const int maxIters) {
  return rpe::update_num_iters(p, ep, modelPoints, maxIters);
}
template <>
inline int RANSACUpdateNumIters<double>(double p, double ep, const int modelPoints, const int maxIters) {
  return rpe::update_num_iters_d(p, ep, modelPoints, maxIters);
}

#endif  // RPE_ESTIMATORS_HPP_
