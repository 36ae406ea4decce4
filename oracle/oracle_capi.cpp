// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE (see oracle/README.md). PARITY UNPINNED.
//
// extern "C" surface of the CPU oracle so that tests/ (ctypes), __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py can call it. Nothing in the product links this.
#include <stdint.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

#include "ransac.hpp"
#include "refine.hpp"
#include "sampler_model.hpp"

namespace orc {
int g_math_mode = 0;
int g_stale_sample_buffers = 0;
}

using namespace orc;

namespace {

template <class T>
void put_pose(const SE3<T>& s, T* q4, T* t3) {
  q4[0] = s.so3.q.x;
  q4[1] = s.so3.q.y;
  q4[2] = s.so3.q.z;
  q4[3] = s.so3.q.w;
  t3[0] = s.t[0];
  t3[1] = s.t[1];
  t3[2] = s.t[2];
}
template <class T>
SE3<T> get_pose(const T* q4, const T* t3) {
  SE3<T> s;
  s.so3.q.x = q4[0];
  s.so3.q.y = q4[1];
  s.so3.q.z = q4[2];
  s.so3.q.w = q4[3];
  s.t = V3<T>(t3[0], t3[1], t3[2]);
  return s;
}

struct RansacOutC {
  int max_votes;
  int iter_final;
  int winner;
  int iters_run;
  long long evals;
  double seconds;
};

template <class T>
void ransac_c(int method, const T* bv, const T* xc, const T* nc, const T* xw, const T* nw, int n, const int32_t* samples,
              int iter_in, T thr3d, T cos_thr, T cos_nl, T confidence, int full, int nthreads, RansacOutC* out, T* q4,
              T* t3, int* votes_out, T* hyps_out, short* mask_out) {
  Corr<T> d{bv, xc, nc, xw, nw, n};
  Thresholds<T> th{thr3d, cos_thr, cos_nl};
  const auto t0 = std::chrono::steady_clock::now();
  RansacResult<T> r;
  if (full && nthreads > 1) {
    // Hypotheses sharded over host threads (every iteration is scored), replay afterwards.
    const int S = method_slots(method);
    std::vector<int> votes((size_t)iter_in * S, -1);
    std::vector<T> hyps((size_t)iter_in * S * 7, T(0));
    std::vector<long long> evals(nthreads, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
      pool.emplace_back([&, t]() {
        for (int ii = t; ii < iter_in; ii += nthreads) {
          SE3<T> hyp[3];
          bool has[3];
          generate_iteration(method, d, samples + 4 * ii, hyp, has);
          for (int s = 0; s < S; ++s) {
            if (!has[s]) continue;
            votes[(size_t)ii * S + s] = score_hypothesis(method, d, hyp[s], th, (short*)0);
            put_pose(hyp[s], &hyps[((size_t)ii * S + s) * 7], &hyps[((size_t)ii * S + s) * 7 + 4]);
            evals[t] += n;
          }
        }
      });
    }
    for (auto& th_ : pool) th_.join();
    r.max_votes = -1;
    r.winner = -1;
    r.evals = 0;
    for (long long e : evals) r.evals += e;
    int Iter = iter_in;
    const int K = method_model_points(method);
    for (int ii = 0; ii < Iter; ++ii)
      for (int s = 0; s < S; ++s) {
        const int v = votes[(size_t)ii * S + s];
        if (v < 0) continue;
        if (v > r.max_votes) {
          r.max_votes = v;
          r.winner = ii * S + s;
          Iter = ransac_update_num_iters<T>(confidence, outlier_ratio<T>(method, n, v), K, Iter);
        }
      }
    r.iter_final = Iter;
    r.iters_run = iter_in;
    if (r.winner >= 0) r.best = get_pose(&hyps[(size_t)r.winner * 7], &hyps[(size_t)r.winner * 7 + 4]);
    if (votes_out) memcpy(votes_out, votes.data(), votes.size() * sizeof(int));
    if (hyps_out) memcpy(hyps_out, hyps.data(), hyps.size() * sizeof(T));
    if (mask_out) {
      if (r.winner >= 0)
        score_hypothesis(method, d, r.best, th, mask_out);
      else
        for (int i = 0; i < n * method_mask_cols(method); ++i) mask_out[i] = 1;
    }
  } else {
    r = ransac<T>(method, d, samples, iter_in, th, confidence, full != 0, votes_out, hyps_out, mask_out);
  }
  const auto t1 = std::chrono::steady_clock::now();
  out->max_votes = r.max_votes;
  out->iter_final = r.iter_final;
  out->winner = r.winner;
  out->iters_run = r.iters_run;
  out->evals = r.evals;
  out->seconds = std::chrono::duration<double>(t1 - t0).count();
  put_pose(r.best, q4, t3);
}

}  // namespace

extern "C" {

void orc_set_math_mode(int m) { g_math_mode = m ? 1 : 0; }
int orc_get_math_mode() { return g_math_mode; }
// the reference's stale sample columns (ransac.hpp, StaleCols): only for the comparison with the reference's own sources
void orc_set_stale_sample_buffers(int on) { g_stale_sample_buffers = on ? 1 : 0; }

// ---- samplers ----------------------------------------------------------------------------
void orc_rand_seq(unsigned seed, int n, int* out) {
  GlibcRand g(seed);
  for (int i = 0; i < n; ++i) out[i] = g.next();
}
// H consecutive RandomElements::run(m) draws from one generator (Utility.hpp:138-155); rows padded to 4.
void orc_sample_table(unsigned seed, int n_corr, int m, int H, int32_t* out) {
  GlibcRand g(seed);
  RandomElementsModel re(n_corr, &g);
  std::vector<int> sel;
  for (int h = 0; h < H; ++h) {
    re.run(m, &sel);
    for (int k = 0; k < 4; ++k) out[4 * h + k] = k < m ? sel[k] : -1;
  }
}
// Same, after discarding `skip` rand() draws (a program that already ran other samplers on the same stream).
void orc_sample_table_skip(unsigned seed, long long skip, int n_corr, int m, int H, int32_t* out) {
  GlibcRand g(seed);
  for (long long i = 0; i < skip; ++i) (void)g.next();
  RandomElementsModel re(n_corr, &g);
  std::vector<int> sel;
  for (int h = 0; h < H; ++h) {
    re.run(m, &sel);
    for (int k = 0; k < 4; ++k) out[4 * h + k] = k < m ? sel[k] : -1;
  }
}
// H consecutive ProsacSampler::sample draws mapped through the weight-sorted index
// (Utility.hpp:183-243, PnPPoseAdapter.hpp:239-255); weights may be null (identity order).
void orc_prosac_table_f(unsigned seed, int n_corr, int m, int H, const float* weights, int32_t* out) {
  GlibcRand g(seed);
  ProsacSamplerModel<float> ps(m, n_corr, &g);
  std::vector<int> idx(n_corr);
  for (int i = 0; i < n_corr; ++i) idx[i] = i;
  if (weights) std::sort(idx.begin(), idx.end(), [&](int a, int b) { return weights[a] > weights[b]; });  // Utility.hpp:115
  for (int h = 0; h < H; ++h) {
    std::vector<int> sel;
    ps.sample(&sel);
    for (int k = 0; k < 4; ++k) {
      int j = k < m ? sel[k] : -1;
      if (j >= 0 && j < n_corr) j = idx[j];  // getSortedIdx leaves out-of-range j untouched
      out[4 * h + k] = j;
    }
  }
}

// ---- Eigen / Sophus model probes ------------------------------------------------------------
#define ORC_DEFINE(SUF, T)                                                                                            \
  void orc_jacobi_svd3_##SUF(const T* A_rowmajor, T* U_rowmajor, T* S, T* V_rowmajor) {                               \
    M3<T> a;                                                                                                          \
    for (int i = 0; i < 3; ++i)                                                                                       \
      for (int j = 0; j < 3; ++j) a(i, j) = A_rowmajor[3 * i + j];                                                    \
    SVD3<T> s = jacobi_svd3(a);                                                                                       \
    for (int i = 0; i < 3; ++i) {                                                                                     \
      S[i] = s.s[i];                                                                                                  \
      for (int j = 0; j < 3; ++j) {                                                                                   \
        U_rowmajor[3 * i + j] = s.U(i, j);                                                                            \
        V_rowmajor[3 * i + j] = s.V(i, j);                                                                            \
      }                                                                                                               \
    }                                                                                                                 \
  }                                                                                                                   \
  void orc_quat_to_matrix_##SUF(const T* q4, T* R_rowmajor) {                                                         \
    Quat<T> q(q4[3], q4[0], q4[1], q4[2]);                                                                            \
    M3<T> r = quat_to_matrix(q);                                                                                      \
    for (int i = 0; i < 3; ++i)                                                                                       \
      for (int j = 0; j < 3; ++j) R_rowmajor[3 * i + j] = r(i, j);                                                    \
  }                                                                                                                   \
  int orc_quat_from_matrix_##SUF(const T* R_rowmajor, T* q4) {                                                        \
    M3<T> r;                                                                                                          \
    for (int i = 0; i < 3; ++i)                                                                                       \
      for (int j = 0; j < 3; ++j) r(i, j) = R_rowmajor[3 * i + j];                                                    \
    SO3<T> s = SO3<T>::from_matrix(r);                                                                                \
    q4[0] = s.q.x;                                                                                                    \
    q4[1] = s.q.y;                                                                                                    \
    q4[2] = s.q.z;                                                                                                    \
    q4[3] = s.q.w;                                                                                                    \
    return s.ok ? 1 : 0;                                                                                              \
  }                                                                                                                   \
  void orc_quat_rotate_##SUF(const T* q4, const T* v3, T* out3) {                                                     \
    Quat<T> q(q4[3], q4[0], q4[1], q4[2]);                                                                            \
    V3<T> r = quat_rotate(q, V3<T>(v3[0], v3[1], v3[2]));                                                             \
    out3[0] = r[0];                                                                                                   \
    out3[1] = r[1];                                                                                                   \
    out3[2] = r[2];                                                                                                   \
  }                                                                                                                   \
  /* ---- solvers ---- */                                                                                             \
  int orc_shinji_##SUF(const T* Xw, const T* Xc, int K, int cols, T* q4, T* t3) {                                     \
    SE3<T> s = shinji(Xw, Xc, K, cols);                                                                               \
    put_pose(s, q4, t3);                                                                                              \
    return s.so3.ok ? 1 : 0;                                                                                          \
  }                                                                                                                   \
  int orc_update_num_iters_##SUF(T p, T ep, int model_points, int max_iters) {                                        \
    return ransac_update_num_iters<T>(p, ep, model_points, max_iters);                                                \
  }                                                                                                                   \
  void orc_o4_roots_##SUF(const T* f5, T* r4) { o4_roots(f5, r4); }                                                   \
  int orc_kneip_main_##SUF(const T* Xw, const T* bv, T* q16, T* t12) {                                                \
    SE3<T> sols[4];                                                                                                   \
    int k = kneip_main(Xw, bv, sols);                                                                                 \
    for (int i = 0; i < k; ++i) put_pose(sols[i], q16 + 4 * i, t12 + 3 * i);                                          \
    return k;                                                                                                         \
  }                                                                                                                   \
  int orc_kneip4_##SUF(const T* Xw, const T* bv, T* q4, T* t3) {                                                      \
    SE3<T> s;                                                                                                         \
    if (!kneip4(Xw, bv, &s)) return 0;                                                                                \
    put_pose(s, q4, t3);                                                                                              \
    return 1;                                                                                                         \
  }                                                                                                                   \
  void orc_nl_2p_##SUF(const T* pt1_c, const T* nl1_c, const T* pt2_c, const T* pt1_w, const T* nl1_w, const T* pt2_w, \
                       T* q4, T* t3) {                                                                                \
    SE3<T> s = nl_2p(col3(pt1_c, 0), col3(nl1_c, 0), col3(pt2_c, 0), col3(pt1_w, 0), col3(nl1_w, 0), col3(pt2_w, 0)); \
    put_pose(s, q4, t3);                                                                                              \
  }                                                                                                                   \
  void orc_sym3_eigenvalues_##SUF(const T* A_rowmajor, T* e3) {                                                       \
    M3<T> a;                                                                                                          \
    for (int i = 0; i < 3; ++i)                                                                                       \
      for (int j = 0; j < 3; ++j) a(i, j) = A_rowmajor[3 * i + j];                                                    \
    sym3_eigenvalues(a, e3);                                                                                          \
  }                                                                                                                   \
  /* ---- scoring / RANSAC / refits ---- */                                                                           \
  int orc_score_##SUF(int method, const T* bv, const T* xc, const T* nc, const T* xw, const T* nw, int n, const T* q4, \
                      const T* t3, T thr3d, T cos_thr, T cos_nl, short* mask) {                                       \
    Corr<T> d{bv, xc, nc, xw, nw, n};                                                                                 \
    Thresholds<T> th{thr3d, cos_thr, cos_nl};                                                                         \
    return score_hypothesis(method, d, get_pose(q4, t3), th, mask);                                                   \
  }                                                                                                                   \
  void orc_ransac_##SUF(int method, const T* bv, const T* xc, const T* nc, const T* xw, const T* nw, int n,           \
                        const int32_t* samples, int iter_in, T thr3d, T cos_thr, T cos_nl, T confidence, int full,    \
                        int nthreads, RansacOutC* out, T* q4, T* t3, int* votes_out, T* hyps_out, short* mask_out) {  \
    ransac_c<T>(method, bv, xc, nc, xw, nw, n, samples, iter_in, thr3d, cos_thr, cos_nl, confidence, full, nthreads,  \
                out, q4, t3, votes_out, hyps_out, mask_out);                                                          \
  }                                                                                                                   \
  int orc_shinji_ls_##SUF(const T* xc, const T* xw, int n, const short* flags3d, T* q4, T* t3) {                      \
    Corr<T> d{(const T*)0, xc, (const T*)0, xw, (const T*)0, n};                                                      \
    SE3<T> s = shinji_ls(d, flags3d);                                                                                 \
    put_pose(s, q4, t3);                                                                                              \
    return s.so3.ok ? 1 : 0;                                                                                          \
  }                                                                                                                   \
  int orc_refine_gn_##SUF(const T* bv, const T* xc, const T* nc, const T* xw, const T* nw, int n, const short* mask,  \
                          int mask_cols, T w2d, T w3d, T wnl, int max_iters, T* q4, T* t3, double* info) {            \
    Corr<T> d{bv, xc, nc, xw, nw, n};                                                                                 \
    SE3<T> s = get_pose(q4, t3);                                                                                      \
    int it = refine_gn(d, mask, mask_cols, w2d, w3d, wnl, max_iters, &s, info);                                       \
    put_pose(s, q4, t3);                                                                                              \
    return it;                                                                                                        \
  }                                                                                                                   \
  void orc_nl_shinji_kneip_ls_##SUF(const T* bv, const T* xc, const T* nc, const T* xw, const T* nw, int n,           \
                                    const short* mask3, const T* weights3, int max_votes, T* q4, T* t3) {             \
    Corr<T> d{bv, xc, nc, xw, nw, n};                                                                                 \
    SE3<T> s = get_pose(q4, t3);                                                                                      \
    nl_shinji_kneip_ls(d, mask3, weights3, max_votes, &s);                                                            \
    put_pose(s, q4, t3);                                                                                              \
  }

ORC_DEFINE(f, float)
ORC_DEFINE(d, double)

// det_math probes (so tests can measure them against libm)
double orc_det_log(double x) { return rpe::det::log_d(x); }
double orc_det_acos(double x) { return rpe::det::acos_d(x); }
double orc_det_atan2(double y, double x) { return rpe::det::atan2_d(y, x); }
double orc_det_cbrt(double x) { return rpe::det::cbrt_d(x); }
void orc_det_sincos(double a, double* s, double* c) { rpe::det::sincos_d(a, s, c); }

}  // extern "C"
