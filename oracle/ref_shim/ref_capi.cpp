// oracle/ref_shim/ref_capi.cpp — TEST INFRASTRUCTURE (see oracle/README.md).
//
// extern "C" harness around the reference's OWN pose headers, included unmodified from /root/reference/pose and
// compiled against the Eigen / Sophus API stand-ins of this directory (Eigen is not installed in this image). It
// exists so that tests/test_ref_shim.py can run the reference's functions next to the oracle restatement on the same
// inputs and the same ::rand() seed and compare every output bit for bit. Built only where /root/reference exists
// (oracle/Makefile target `ref`, output oracle/_ref/libref_shim.so); nothing in the product links or loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <AbsoluteOrientationNormal.hpp>  // /root/reference/pose: pulls in every adapter, solver and estimator

namespace {

template <class T>
Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> load3(const T* p, int n) {
  Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> m(3, n);
  if (p) memcpy(m.data(), p, sizeof(T) * 3 * (size_t)n);
  return m;
}
template <class T>
void put_pose(const Sophus::SO3<T>& R, const Eigen::Matrix<T, 3, 1>& t, T* q4, T* t3) {
  const Eigen::Quaternion<T> q = R.unit_quaternion();
  q4[0] = q.x();
  q4[1] = q.y();
  q4[2] = q.z();
  q4[3] = q.w();
  for (int i = 0; i < 3; ++i) t3[i] = t(i);
}
struct RefOut {
  int max_votes;
  int iter_final;
  int n_idx[3];  // lengths of the 2-D / 3-D / normal inlier index lists after cvtInlier
  long long ensure_failures;
};

// method ids follow oracle/ransac.hpp (orc::Method); sampler: 0 = RandomElements, 1 = ProsacSampler
template <class T>
int run_ransac(int method, int sampler, const T* bv, const T* xc, const T* nc, const T* xw, const T* nw, int n,
               const T* weights3, unsigned seed, int iter_in, T thr3d, T thr2d, T focal, T thrN, T confidence, int refit,
               RefOut* out, T* q4, T* t3, T* q4r, T* t3r, short* mask3) {
  typedef Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> MX;
  const MX BV = load3(bv, n), XC = load3(xc, n), NC = load3(nc, n), XW = load3(xw, n), NW = load3(nw, n);
  MX W;
  if (weights3) {
    W.resize(n, 3);
    memcpy(W.data(), weights3, sizeof(T) * 3 * (size_t)n);
  }
  Sophus::shim_ensure_failures() = 0;
  int Iter = iter_in;
  memset(mask3, 0, sizeof(short) * 3 * (size_t)n);
  out->n_idx[0] = out->n_idx[1] = out->n_idx[2] = -1;
  srand(seed);
  if (method == 0) {  // orc::M_SHINJI
    AOOnlyPoseAdapter<T> ad(XC, XW);
    if (weights3) ad.setWeights(W);
    if (sampler)
      shinji_prosac<T>(ad, thr3d, Iter, confidence);
    else
      shinji_ransac2<T>(ad, thr3d, Iter, confidence);
    out->max_votes = ad.getMaxVotes();
    put_pose<T>(ad.getRcw(), ad.gettw(), q4, t3);
    for (int i = 0; i < n; ++i) mask3[n + i] = ad.isInlier33(i) ? 1 : 0;
    out->n_idx[1] = (int)ad.getInlierIdx().size();
    if (refit == 1) shinji_ls1<T>(ad);
    if (refit == 3) shinji_ls2<T>(ad);
    put_pose<T>(ad.getRcw(), ad.gettw(), q4r, t3r);
  } else if (method == 1 || method == 6) {  // orc::M_KNEIP (kneip_ransac), orc::M_KNEIP_QUAT (kneip_prosac's scoring)
    PnPPoseAdapter<T> ad(BV, XW);
    ad.setFocal(focal, focal);
    if (weights3) ad.setWeights(W);
    if (method == 6)
      kneip_prosac<T>(ad, thr2d, Iter, confidence);
    else
      kneip_ransac<T>(ad, thr2d, Iter, confidence);
    out->max_votes = ad.getMaxVotes();
    put_pose<T>(ad.getRcw(), ad.gettw(), q4, t3);
    for (int i = 0; i < n; ++i) mask3[i] = ad.isInlier23(i) ? 1 : 0;
    out->n_idx[0] = (int)ad.getInlierIdx().size();
    put_pose<T>(ad.getRcw(), ad.gettw(), q4r, t3r);
  } else if (method == 2) {  // orc::M_SHINJI_KNEIP
    AOPoseAdapter<T> ad(BV, XC, XW);
    ad.setFocal(focal, focal);
    if (weights3) ad.setWeights(W);
    if (sampler)
      shinji_kneip_prosac<T>(ad, thr3d, thr2d, Iter, confidence);
    else
      shinji_kneip_ransac<T>(ad, thr3d, thr2d, Iter, confidence);
    out->max_votes = ad.getMaxVotes();
    put_pose<T>(ad.getRcw(), ad.gettw(), q4, t3);
    for (int i = 0; i < n; ++i) {
      mask3[i] = ad.isInlier23(i) ? 1 : 0;
      mask3[n + i] = ad.isInlier33(i) ? 1 : 0;
    }
    out->n_idx[0] = (int)static_cast<PnPPoseAdapter<T>&>(ad).getInlierIdx().size();
    out->n_idx[1] = (int)ad.getInlierIdx().size();
    if (refit == 1) shinji_ls<T>(ad);
    put_pose<T>(ad.getRcw(), ad.gettw(), q4r, t3r);
  } else if (method >= 3 && method <= 5) {  // orc::M_NL_KNEIP, M_NL_SHINJI, M_NL_SHINJI_KNEIP
    NormalAOPoseAdapter<T> ad(BV, XC, NC, XW, NW);
    ad.setFocal(focal, focal);
    if (weights3) ad.setWeights(W);
    if (method == 3)
      nl_kneip_ransac<T>(ad, thr2d, thrN, Iter, confidence);
    else if (method == 4)
      nl_shinji_ransac<T>(ad, thr3d, thrN, Iter, confidence);
    else
      nl_shinji_kneip_ransac<T>(ad, thr3d, thr2d, thrN, Iter, confidence);
    out->max_votes = ad.getMaxVotes();
    put_pose<T>(ad.getRcw(), ad.gettw(), q4, t3);
    for (int i = 0; i < n; ++i) {
      mask3[i] = ad.isInlier23(i) ? 1 : 0;
      mask3[n + i] = ad.isInlier33(i) ? 1 : 0;
      mask3[2 * n + i] = ad.isInlierNN(i) ? 1 : 0;
    }
    out->n_idx[0] = (int)static_cast<PnPPoseAdapter<T>&>(ad).getInlierIdx().size();
    out->n_idx[1] = (int)static_cast<AOPoseAdapter<T>&>(ad).getInlierIdx().size();
    out->n_idx[2] = (int)ad.getInlierIdx().size();
    if (refit == 2) nl_shinji_kneip_ls<T>(ad);
    put_pose<T>(ad.getRcw(), ad.gettw(), q4r, t3r);
  } else {
    return -1;
  }
  out->iter_final = Iter;
  out->ensure_failures = Sophus::shim_ensure_failures();
  return 0;
}

template <class T>
int run_shinji(const T* Xw, const T* Xc, int K, int cols, T* q4, T* t3) {
  Sophus::shim_ensure_failures() = 0;
  const Sophus::SE3<T> s = shinji<T>(load3(Xw, cols), load3(Xc, cols), K);
  put_pose<T>(s.so3(), s.translation(), q4, t3);
  return Sophus::shim_ensure_failures() == 0;
}
template <class T>
int run_kneip_main(const T* Xw3, const T* bv3, T* q44, T* t43) {
  std::vector<Sophus::SE3<T> > sol;
  kneip_main<T>(load3(Xw3, 3), load3(bv3, 3), &sol);
  for (size_t i = 0; i < sol.size() && i < 4; ++i) put_pose<T>(sol[i].so3(), sol[i].translation(), q44 + 4 * i, t43 + 3 * i);
  return (int)sol.size();
}
template <class T>
int run_kneip4(const T* Xw4, const T* bv4, T* q4, T* t3) {
  Sophus::SE3<T> s;
  const bool ok = kneip<T>(load3(Xw4, 4), load3(bv4, 4), &s);
  put_pose<T>(s.so3(), s.translation(), q4, t3);
  return ok ? 1 : 0;
}
template <class T>
void run_nl_2p(const T* a, const T* b, const T* c, const T* d, const T* e, const T* f, T* q4, T* t3) {
  typedef Eigen::Matrix<T, 3, 1> V;
  Sophus::SE3<T> s;
  nl_2p<T>(V(a[0], a[1], a[2]), V(b[0], b[1], b[2]), V(c[0], c[1], c[2]), V(d[0], d[1], d[2]), V(e[0], e[1], e[2]),
           V(f[0], f[1], f[2]), &s);
  put_pose<T>(s.so3(), s.translation(), q4, t3);
}
template <class T>
void run_o4_roots(const T* f5, T* r4) {
  Eigen::Matrix<T, 5, 1> f;
  for (int i = 0; i < 5; ++i) f(i, 0) = f5[i];
  const std::vector<T> r = o4_roots<T>(f);
  for (int i = 0; i < 4; ++i) r4[i] = r[i];
}

}  // namespace

extern "C" {

#define REF_DEFINE(SUF, T)                                                                                              \
  int ref_ransac_##SUF(int method, int sampler, const T* bv, const T* xc, const T* nc, const T* xw, const T* nw, int n, \
                       const T* weights3, unsigned seed, int iter_in, T thr3d, T thr2d, T focal, T thrN, T confidence,  \
                       int refit, RefOut* out, T* q4, T* t3, T* q4r, T* t3r, short* mask3) {                            \
    return run_ransac<T>(method, sampler, bv, xc, nc, xw, nw, n, weights3, seed, iter_in, thr3d, thr2d, focal, thrN,    \
                         confidence, refit, out, q4, t3, q4r, t3r, mask3);                                              \
  }                                                                                                                     \
  int ref_shinji_##SUF(const T* Xw, const T* Xc, int K, int cols, T* q4, T* t3) {                                       \
    return run_shinji<T>(Xw, Xc, K, cols, q4, t3);                                                                      \
  }                                                                                                                     \
  int ref_kneip_main_##SUF(const T* Xw3, const T* bv3, T* q44, T* t43) { return run_kneip_main<T>(Xw3, bv3, q44, t43); } \
  int ref_kneip4_##SUF(const T* Xw4, const T* bv4, T* q4, T* t3) { return run_kneip4<T>(Xw4, bv4, q4, t3); }            \
  void ref_nl_2p_##SUF(const T* a, const T* b, const T* c, const T* d, const T* e, const T* f, T* q4, T* t3) {          \
    run_nl_2p<T>(a, b, c, d, e, f, q4, t3);                                                                             \
  }                                                                                                                     \
  void ref_o4_roots_##SUF(const T* f5, T* r4) { run_o4_roots<T>(f5, r4); }                                              \
  int ref_update_num_iters_##SUF(T p, T ep, int model_points, int max_iters) {                                          \
    return RANSACUpdateNumIters<T>(p, ep, model_points, max_iters);                                                     \
  }                                                                                                                     \
  /* the thresholds exactly as the estimators form them (P3P.hpp:323, AbsoluteOrientationNormal.hpp:223) */            \
  T ref_cos_thr_##SUF(T thr2d, T focal) { return cos(atan(thr2d / ((focal + focal) / 2))); }                            \
  T ref_cos_nl_##SUF(T thrN) { return cos(thrN); }

REF_DEFINE(f, float)
REF_DEFINE(d, double)

}  // extern "C"

// ---- the reference's own Simulator (Simulator.hpp), for inputs made exactly the way SimpleMain / TestMain make them ----
#include <Simulator.hpp>

namespace {
template <class T>
void copy3(const Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic>& m, T* out) {
  if (out) memcpy(out, m.data(), sizeof(T) * (size_t)m.rows() * m.cols());
}
// kind 0: simulate_3d_3d_correspondences, 1: simulate_2d_3d_correspondences, 2: simulate_2d_3d_nl_correspondences,
// 3: simulate_kinect_2d_3d_nl_correspondences.
// Pose drawn as SimpleMain.cpp:22-23 does; ::rand() seeded with `seed`, the global normal generator re-seeded with it too.
template <class T>
int run_sim(int kind, unsigned seed, int n, T n2d, T or2d, T n3d, T or3d, T nnl, T ornl, T min_depth, T max_depth, T f,
            int gaussian, T* q4, T* t3, T* xw, T* nw, T* xc, T* nc, T* bv, T* weights3) {
  typedef Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> MX;
  srand(seed);
  generator.seed(seed);
  distribution.reset();
  Sophus::shim_ensure_failures() = 0;
  const Eigen::Matrix<T, 3, 1> t = generate_random_translation_uniform<T>(T(5.0));
  const Sophus::SO3<T> R = generate_random_rotation<T>(T(M_PI / 2), false);
  put_pose<T>(R, t, q4, t3);
  MX Q, M, P, N, U, W(n, 3);
  if (kind == 0) {
    simulate_3d_3d_correspondences<T>(R, t, n, n3d, or3d, min_depth, max_depth, f, gaussian != 0, &Q, &P, &W);
  } else if (kind == 1) {
    simulate_2d_3d_correspondences<T>(R, t, n, n2d, or2d, min_depth, max_depth, f, gaussian != 0, &Q, &U, &P, &W);
  } else if (kind == 2) {
    simulate_2d_3d_nl_correspondences<T>(R, t, n, n2d, or2d, n3d, or3d, nnl, ornl, min_depth, max_depth, f, gaussian != 0, &Q,
                                         &M, &P, &N, &U, &W);
  } else if (kind == 3) {  // Kinect lateral / axial noise model (Simulator.hpp:368-436); n3d is not used
    simulate_kinect_2d_3d_nl_correspondences<T>(R, t, n, n2d, or2d, or3d, nnl, ornl, min_depth, max_depth, f, &Q, &M, &P, &N, &U,
                                                &W);
  } else {
    return -1;
  }
  copy3(Q, xw);
  copy3(P, xc);
  if (kind != 0) copy3(U, bv);
  if (kind >= 2) {
    copy3(M, nw);
    copy3(N, nc);
  }
  copy3(W, weights3);
  return Sophus::shim_ensure_failures() == 0 ? 0 : 1;
}
}  // namespace

extern "C" {
int ref_sim_f(int kind, unsigned seed, int n, float n2d, float or2d, float n3d, float or3d, float nnl, float ornl,
              float min_depth, float max_depth, float f, int gaussian, float* q4, float* t3, float* xw, float* nw, float* xc,
              float* nc, float* bv, float* weights3) {
  return run_sim<float>(kind, seed, n, n2d, or2d, n3d, or3d, nnl, ornl, min_depth, max_depth, f, gaussian, q4, t3, xw, nw, xc,
                        nc, bv, weights3);
}
int ref_sim_d(int kind, unsigned seed, int n, double n2d, double or2d, double n3d, double or3d, double nnl, double ornl,
              double min_depth, double max_depth, double f, int gaussian, double* q4, double* t3, double* xw, double* nw,
              double* xc, double* nc, double* bv, double* weights3) {
  return run_sim<double>(kind, seed, n, n2d, or2d, n3d, or3d, nnl, ornl, min_depth, max_depth, f, gaussian, q4, t3, xw, nw,
                         xc, nc, bv, weights3);
}
}
