// host_utils.cpp — the C-ABI entry points that the reference also computes on the CPU:
// sample tables (Utility.hpp:125-250), the adaptive stopping rule (P3P.hpp:296-318) and the
// synthetic generators (Simulator.hpp). No device code here.
#include <string.h>

#include <vector>

#include "../../include/rpe/Utility.hpp"
#include "../../include/rpe/ransac_rule.h"
#include "../../include/rpe/sim_core.hpp"
#include "../../include/rpe/solvers_min.h"
#include "../../include/rpe_c_api.h"

struct rpe_sampler {
  rpe::GlibcRandom src;
  RandomElements<int> re;
  std::vector<int> sel;
  int n;
  rpe_sampler(uint32_t seed, int n_) : src(seed), re(n_, &src), n(n_) {}
};

extern "C" {

int rpe_update_num_iters(float p, float ep, int model_points, int max_iters) {
  return rpe::update_num_iters(p, ep, model_points, max_iters);
}

// MinimalSolvers.hpp on the host: the same templates the device kernels instantiate (identical bits)
int rpe_min_ev_host(const float* M9, int count, float* E3) {
  if (!M9 || !E3 || count < 0) return RPE_ERR_ARG;
  for (int i = 0; i < count; ++i) rpe::sym3_eigenvalues<float>(M9 + 9 * (size_t)i, E3 + 3 * (size_t)i);
  return RPE_OK;
}
int rpe_min_ev_host_f64(const double* M9, int count, double* E3) {
  if (!M9 || !E3 || count < 0) return RPE_ERR_ARG;
  for (int i = 0; i < count; ++i) rpe::sym3_eigenvalues<double>(M9 + 9 * (size_t)i, E3 + 3 * (size_t)i);
  return RPE_OK;
}
int rpe_min_ms_host(const float* in24, int count, float* q4, float* t3) {
  if (!in24 || !q4 || !t3 || count < 0) return RPE_ERR_ARG;
  for (int i = 0; i < count; ++i) rpe::min_solver_2pn<float>(in24 + 24 * (size_t)i, q4 + 4 * (size_t)i, t3 + 3 * (size_t)i);
  return RPE_OK;
}

int rpe_sample_table(uint32_t seed, int n, int m, int H, int32_t* samples) {
  if (!samples || n <= 0 || m <= 0 || m > 4 || m > n || H < 0) return RPE_ERR_ARG;
  rpe::GlibcRandom src(seed);
  RandomElements<int> re(n, &src);
  std::vector<int> sel;
  for (int h = 0; h < H; ++h) {
    re.run(m, &sel);
    for (int k = 0; k < 4; ++k) samples[4 * h + k] = k < m ? sel[k] : -1;
  }
  return RPE_OK;
}

int rpe_sampler_create(uint32_t seed, int n, rpe_sampler** out) {
  if (!out || n <= 0) return RPE_ERR_ARG;
  *out = new rpe_sampler(seed, n);
  return RPE_OK;
}
int rpe_sampler_rows(rpe_sampler* s, int m, int H, int32_t* samples) {
  if (!s || !samples || m <= 0 || m > 4 || m > s->n || H < 0) return RPE_ERR_ARG;
  for (int h = 0; h < H; ++h) {
    s->re.run(m, &s->sel);
    for (int k = 0; k < 4; ++k) samples[4 * h + k] = k < m ? s->sel[k] : -1;
  }
  return RPE_OK;
}
void rpe_sampler_destroy(rpe_sampler* s) { delete s; }
int rpe_sampler_reseed(rpe_sampler* s, uint32_t seed) {
  if (!s) return RPE_ERR_ARG;
  s->src.seed_with(seed);  // RandomElements::run leaves the identity permutation behind: nothing else to reset
  return RPE_OK;
}

int rpe_prosac_table(uint32_t seed, int n, int m, int H, const float* weights, int32_t* samples) {
  if (!samples || n <= 0 || m <= 1 || m > 4 || m > n || H < 0) return RPE_ERR_ARG;
  rpe::GlibcRandom src(seed);
  ProsacSampler<float> ps(m, n, &src);
  std::vector<int> order;
  if (weights) {
    std::vector<float> w(weights, weights + n);
    order = sortIndexes<float>(w);  // adapter.sortIdx() (PnPPoseAdapter.hpp:239-244)
  }
  for (int h = 0; h < H; ++h) {
    std::vector<int> sel;
    ps.sample(&sel);
    for (int k = 0; k < 4; ++k) {
      int j = k < m ? sel[k] : -1;
      // getSortedIdx (PnPPoseAdapter.hpp:246-255) leaves indices beyond the table untouched; the sampler's out-of-range
      // index n == N (Utility.hpp:238) is clamped to N-1, the same policy as the header path (rpe/Estimators.hpp)
      if (j >= 0 && weights && j < (int)order.size()) j = order[j];
      if (j >= n) j = n - 1;
      samples[4 * h + k] = j;
    }
  }
  return RPE_OK;
}

int rpe_sim_pose(uint64_t seed, float max_angle_rad, float t_size, float q_xyzw[4], float t[3]) {
  if (!q_xyzw || !t) return RPE_ERR_ARG;
  rpe::sim::Rng rng(seed);
  const rpe::sim::Pose<float> p = rpe::sim::random_pose<float>(rng, max_angle_rad, t_size);
  memcpy(q_xyzw, p.q, sizeof(p.q));
  memcpy(t, p.t, sizeof(p.t));
  return RPE_OK;
}

static rpe::sim::Pose<float> make_pose(const float q[4], const float t[3]) {
  rpe::sim::Pose<float> p;
  memcpy(p.q, q, sizeof(p.q));
  memcpy(p.t, t, sizeof(p.t));
  return p;
}

int rpe_sim_3d_3d(uint64_t seed, const float q_xyzw[4], const float t[3], int n, float noise, float outlier_ratio,
                  float min_depth, float max_depth, float f, int use_gaussian, float* Q_xw, float* P_xc, float* weights3) {
  if (!q_xyzw || !t || n <= 0 || !Q_xw || !P_xc) return RPE_ERR_ARG;
  rpe::sim::Rng rng(seed);
  rpe::sim::simulate_3d_3d<float>(rng, make_pose(q_xyzw, t), n, noise, outlier_ratio, min_depth, max_depth, f,
                                  use_gaussian != 0, Q_xw, P_xc, weights3);
  return RPE_OK;
}

int rpe_sim_2d_3d(uint64_t seed, const float q_xyzw[4], const float t[3], int n, float noise_px, float outlier_ratio,
                  float min_depth, float max_depth, float f, int use_gaussian, float* Q_xw, float* U_bv, float* P_gt,
                  float* weights3) {
  if (!q_xyzw || !t || n <= 0 || !Q_xw || !U_bv) return RPE_ERR_ARG;
  rpe::sim::Rng rng(seed);
  rpe::sim::simulate_2d_3d<float>(rng, make_pose(q_xyzw, t), n, noise_px, outlier_ratio, min_depth, max_depth, f,
                                  use_gaussian != 0, Q_xw, U_bv, P_gt, weights3);
  return RPE_OK;
}

int rpe_sim_2d_3d_nl(uint64_t seed, const float q_xyzw[4], const float t[3], int n, float n2d, float or2d, float n3d,
                     float or3d, float nnl, float ornl, float min_depth, float max_depth, float f, int use_gaussian,
                     float* Q_xw, float* M_nw, float* P_xc, float* N_nc, float* U_bv, float* weights3) {
  if (!q_xyzw || !t || n <= 0 || !Q_xw || !M_nw || !P_xc || !N_nc || !U_bv) return RPE_ERR_ARG;
  rpe::sim::Rng rng(seed);
  rpe::sim::simulate_2d_3d_nl<float>(rng, make_pose(q_xyzw, t), n, n2d, or2d, n3d, or3d, nnl, ornl, min_depth, max_depth,
                                     f, use_gaussian != 0, Q_xw, M_nw, P_xc, N_nc, U_bv, weights3);
  return RPE_OK;
}

int rpe_sim_kinect_2d_3d_nl(uint64_t seed, const float q_xyzw[4], const float t[3], int n, float n2d, float or2d, float or3d,
                            float nnl, float ornl, float min_depth, float max_depth, float f, float* Q_xw, float* M_nw,
                            float* P_xc, float* N_nc, float* U_bv, float* weights3) {
  if (!q_xyzw || !t || n <= 0 || !Q_xw || !M_nw || !P_xc || !N_nc || !U_bv) return RPE_ERR_ARG;
  rpe::sim::Rng rng(seed);
  rpe::sim::simulate_kinect_2d_3d_nl<float>(rng, make_pose(q_xyzw, t), n, n2d, or2d, or3d, nnl, ornl, min_depth, max_depth, f,
                                            Q_xw, M_nw, P_xc, N_nc, U_bv, weights3);
  return RPE_OK;
}

}  // extern "C"
