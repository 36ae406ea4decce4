// oracle/ransac.hpp — TEST INFRASTRUCTURE (see oracle/README.md). PARITY UNPINNED.
//
// CPU restatement of the reference's robust estimators: the per-iteration
// sample -> minimal solve -> score all N -> keep best -> shrink Iter loops of
//   shinji_ransac / shinji_ransac2      AbsoluteOrientation.hpp:101-213
//   shinji_kneip_ransac                 AbsoluteOrientation.hpp:367-438
//   kneip_ransac / kneip_prosac         P3P.hpp:320-469
//   nl_kneip_ransac                     AbsoluteOrientationNormal.hpp:215-284
//   nl_shinji_ransac                    AbsoluteOrientationNormal.hpp:286-354
//   nl_shinji_kneip_ransac              AbsoluteOrientationNormal.hpp:356-445
// and the closed-form refits shinji_ls / shinji_ls1 / shinji_ls2 (AbsoluteOrientation.hpp:273-342).
// The sample table (H x 4 int32, produced by RandomElements / ProsacSampler on the host) is an
// INPUT, so the same draws can be fed to the CUDA product.
//
// Deliberate deviations from the reference, all mirrored by the product:
//  * inlier index lists are int32 (the reference's `short` loops overflow for N > 32767:
//    PnPPoseAdapter.hpp:232, AOPoseAdapter.hpp:212, NormalAOPoseAdapter.hpp:226, AOOnlyPoseAdapter.hpp:226);
//  * a hypothesis whose SO3(R) constructor would abort() is dropped (slot votes = -1);
//  * sample buffers are not "stale" across iterations (AbsoluteOrientation.hpp:377,353-359): an
//    invalid (all-NaN) camera point simply propagates NaN into nl_2p, which then scores 0 votes.
#ifndef ORACLE_RANSAC_HPP_
#define ORACLE_RANSAC_HPP_

#include <stdint.h>

#include <vector>

#include "solvers.hpp"

namespace orc {

enum Method {
  M_SHINJI = 0,            // 3-D only, 1 slot
  M_KNEIP = 1,             // 2-D only, matrix-form rotation (P3P.hpp:365), model points 4
  M_SHINJI_KNEIP = 2,      // 3-D + 2-D, slots {shinji, kneip}
  M_NL_KNEIP = 3,          // normal + 2-D, slot {kneip}
  M_NL_SHINJI = 4,         // normal + 3-D, slots {shinji, nl_2p}
  M_NL_SHINJI_KNEIP = 5,   // normal + 3-D + 2-D, slots {shinji, kneip, nl_2p}
  M_KNEIP_QUAT = 6         // kneip_prosac's scoring: 2-D only, quaternion-form rotation (P3P.hpp:442)
};

inline int method_slots(int m) {
  switch (m) {
    case M_SHINJI_KNEIP: return 2;
    case M_NL_SHINJI: return 2;
    case M_NL_SHINJI_KNEIP: return 3;
    default: return 1;
  }
}
inline int method_mask_cols(int m) {
  switch (m) {
    case M_KNEIP:
    case M_KNEIP_QUAT: return 1;
    case M_SHINJI:
    case M_SHINJI_KNEIP: return 2;
    default: return 3;
  }
}
inline int method_sample_size(int m) { return m == M_SHINJI ? 3 : 4; }
inline int method_model_points(int m) { return (m == M_KNEIP || m == M_KNEIP_QUAT) ? 4 : 3; }
// number of vote-casting modalities (the divisor of the outlier ratio)
inline int method_modalities(int m) {
  switch (m) {
    case M_SHINJI:
    case M_KNEIP:
    case M_KNEIP_QUAT: return 1;
    case M_NL_SHINJI_KNEIP: return 3;
    default: return 2;
  }
}

template <class T>
struct Corr {
  const T* bv;  // 3 x n bearing vectors (camera frame)      PnPPoseAdapter.hpp:96
  const T* xc;  // 3 x n points, camera frame                 AOPoseAdapter.hpp:88
  const T* nc;  // 3 x n normals, camera frame                NormalAOPoseAdapter.hpp:83
  const T* xw;  // 3 x n points, world frame                  PnPPoseAdapter.hpp:98
  const T* nw;  // 3 x n normals, world frame                 NormalAOPoseAdapter.hpp:84
  int n;
  // AOPoseAdapter.hpp:147-152 / AOOnlyPoseAdapter.hpp:161-166: valid iff ANY coordinate is not NaN
  bool valid(int i) const {
    return xc[3 * i] == xc[3 * i] || xc[3 * i + 1] == xc[3 * i + 1] || xc[3 * i + 2] == xc[3 * i + 2];
  }
};

template <class T>
struct Thresholds {
  T thr3d;    // metres, compared with ||e||
  T cos_thr;  // cos(atan(thre_2d / focal))   P3P.hpp:323
  T cos_nl;   // cos(nl_thre)                 AbsoluteOrientationNormal.hpp:223
};

// Score ONE hypothesis against all correspondences. `mask` (may be null) is column-major
// n x cols shorts: col 0 = 2-D flags, col 1 = 3-D flags, col 2 = normal flags, exactly the
// layout the reference hands to setInlier (e.g. AbsoluteOrientation.hpp:134,139; P3P.hpp:359,373).
template <class T>
inline int score_hypothesis(int method, const Corr<T>& d, const SE3<T>& s, const Thresholds<T>& th, short* mask) {
  const int n = d.n;
  const int cols = method_mask_cols(method);
  if (mask)
    for (int i = 0; i < n * cols; ++i) mask[i] = 0;
  int votes = 0;
  const bool use_n = method == M_NL_KNEIP || method == M_NL_SHINJI || method == M_NL_SHINJI_KNEIP;
  const bool use_3d = method == M_SHINJI || method == M_SHINJI_KNEIP || method == M_NL_SHINJI || method == M_NL_SHINJI_KNEIP;
  const bool use_2d = method != M_SHINJI && method != M_NL_SHINJI;
  M3<T> Rm;
  if (method == M_KNEIP) Rm = s.so3.matrix();  // P3P.hpp:365 evaluates .matrix() per point: same value
  for (int c = 0; c < n; ++c) {
    if ((use_n || use_3d) && d.valid(c)) {
      if (use_n) {
        // AbsoluteOrientationNormal.hpp:248-252 / :325-329 / :400-404
        const T cos_alpha = dot(col3(d.nc, c), s.so3 * col3(d.nw, c));
        if (cos_alpha > th.cos_nl) {
          if (mask) mask[2 * n + c] = 1;
          votes++;
        }
      }
      if (use_3d) {
        // AbsoluteOrientation.hpp:137-141 / :194-198 / :406-410, AbsoluteOrientationNormal.hpp:331-335 / :407-411
        const V3<T> e = col3(d.xc, c) - (s.so3 * col3(d.xw, c) + s.t);
        if (norm(e) < th.thr3d) {
          if (mask) mask[1 * n + c] = 1;
          votes++;
        }
      }
    }
    if (use_2d) {
      V3<T> pc;
      if (method == M_KNEIP)
        pc = Rm * col3(d.xw, c) + s.t;  // P3P.hpp:365
      else
        pc = s.so3 * col3(d.xw, c) + s.t;  // P3P.hpp:442, AbsoluteOrientation.hpp:413, AbsoluteOrientationNormal.hpp:255,414
      pc = pc / norm(pc);                   // P3P.hpp:366
      const T cos_a = dot(pc, col3(d.bv, c));  // P3P.hpp:369
      if (cos_a > th.cos_thr) {
        if (mask) mask[c] = 1;
        votes++;
      }
    }
  }
  return votes;
}

// Generate the (up to 3) hypotheses of one RANSAC iteration from its 4 (or 3) sampled columns.
// Order of slots = order of v_solutions.push_back in the reference. has[s] = false when the
// reference pushes nothing for that slot (invalid sample, P3P without solution, aborted SO3).
// The reference hoists its sample buffers out of the loop and assign_sample leaves the camera-side columns of an
// INVALID (all-NaN) sample point untouched (AbsoluteOrientationNormal.hpp:48-75, :299-300): nl_2p, which is called in
// every iteration (:315, :389), then pairs the current world point / normal with the camera point / normal of an
// EARLIER sample. The product and this oracle deliberately do not (header of this file): NaN goes into nl_2p and the
// hypothesis scores nothing. `StaleCols` is the reference's behaviour, for the comparison with its own sources only
// (orc_set_stale_sample_buffers; columns start at zero, the reference's are uninitialised memory).
template <class T>
struct StaleCols {
  V3<T> xc[3], nc[3];
};
extern int g_stale_sample_buffers;

template <class T>
inline void generate_iteration(int method, const Corr<T>& d, const int* sel, SE3<T> hyp[3], bool has[3],
                               StaleCols<T>* stale = nullptr) {
  has[0] = has[1] = has[2] = false;
  T Xw[12], Xc[12], bv[12], Nw[12], Nc[12];
  for (int i = 0; i < 12; ++i) Xw[i] = Xc[i] = bv[i] = Nw[i] = Nc[i] = std::numeric_limits<T>::quiet_NaN();
  const int ns = method_sample_size(method);
  bool all_valid = true;
  for (int k = 0; k < ns; ++k) {
    const int c = sel[k];
    for (int r = 0; r < 3; ++r) {
      Xw[3 * k + r] = d.xw[3 * c + r];
      if (d.bv) bv[3 * k + r] = d.bv[3 * c + r];
      if (d.nw) Nw[3 * k + r] = d.nw[3 * c + r];
      if (d.xc) Xc[3 * k + r] = d.xc[3 * c + r];
      if (d.nc) Nc[3 * k + r] = d.nc[3 * c + r];
    }
    if (k < 3 && d.xc && !d.valid(c)) all_valid = false;
  }
  if (method == M_SHINJI) {
    // AbsoluteOrientation.hpp:118-130 — `continue` on an invalid sample
    if (all_valid) {
      hyp[0] = shinji(Xw, Xc, 3, 3);
      has[0] = hyp[0].so3.ok;
    }
    return;
  }
  if (method == M_KNEIP || method == M_KNEIP_QUAT) {
    // P3P.hpp:336-356. kneip(adapter,i0,i1,i2) + 4th-point test with minScore = 1000000.0
    SE3<T> sols[4];
    const int k = kneip_main(Xw, bv, sols);
    T minScore = T(1000000.0);
    int minIndex = -1;
    const V3<T> pw = col3(Xw, 3), b3 = col3(bv, 3);
    for (int i = 0; i < k; ++i) {
      V3<T> pc = sols[i].so3.matrix() * pw + sols[i].t;
      pc = pc / norm(pc);
      const T score = (T)(1.0 - (double)dot(pc, b3));
      if (score < minScore) {
        minScore = score;
        minIndex = i;
      }
    }
    if (minIndex != -1) {
      hyp[0] = sols[minIndex];
      has[0] = true;
    }
    return;
  }
  int slot = 0;
  const bool want_shinji = method == M_SHINJI_KNEIP || method == M_NL_SHINJI || method == M_NL_SHINJI_KNEIP;
  const bool want_kneip = method == M_SHINJI_KNEIP || method == M_NL_KNEIP || method == M_NL_SHINJI_KNEIP;
  const bool want_nl2p = method == M_NL_SHINJI || method == M_NL_SHINJI_KNEIP;
  if (want_shinji) {
    // assign_sample -> use_shinji iff the 3 camera points are valid (AbsoluteOrientation.hpp:344-365,
    // AbsoluteOrientationNormal.hpp:48-75); shinji(X_w, X_c, K=3) on 3 x 4 buffers => cols = 4
    if (all_valid) {
      hyp[slot] = shinji(Xw, Xc, 3, 4);
      has[slot] = hyp[slot].so3.ok;
    }
    ++slot;
  }
  if (want_kneip) {
    SE3<T> sk;
    if (kneip4(Xw, bv, &sk)) {
      hyp[slot] = sk;
      has[slot] = true;
    }
    ++slot;
  }
  if (want_nl2p) {
    // AbsoluteOrientationNormal.hpp:315 / :389
    if (stale) {
      for (int k = 0; k < 3; ++k)
        if (d.valid(sel[k])) {
          stale->xc[k] = col3(Xc, k);
          stale->nc[k] = col3(Nc, k);
        }
      hyp[slot] = nl_2p(stale->xc[0], stale->nc[0], stale->xc[1], col3(Xw, 0), col3(Nw, 0), col3(Xw, 1));
    } else
    hyp[slot] = nl_2p(col3(Xc, 0), col3(Nc, 0), col3(Xc, 1), col3(Xw, 0), col3(Nw, 0), col3(Xw, 1));
    has[slot] = true;
    ++slot;
  }
}

// Outlier ratio handed to RANSACUpdateNumIters, with the reference's exact typing:
//   (Tp)(N - votes) / N                      AbsoluteOrientation.hpp:150, P3P.hpp:383
//   (Tp)(N*2 - votes) / N / 2                AbsoluteOrientation.hpp:429, AbsoluteOrientationNormal.hpp:276,346
//   (Tp)(N*3 - votes) / N / 3                AbsoluteOrientationNormal.hpp:435
template <class T>
inline T outlier_ratio(int method, int n, int votes) {
  const int m = method_modalities(method);
  if (m == 1) return (T)(n - votes) / n;
  return (T)(n * m - votes) / n / m;
}

template <class T>
struct RansacResult {
  SE3<T> best;
  int max_votes;    // adapter.getMaxVotes()
  int iter_final;   // the in/out `Iter`
  int winner;       // iteration * slots + slot of the accepted hypothesis, -1 if none
  int iters_run;    // outer iterations actually executed
  long long evals;  // hypothesis x correspondence evaluations actually performed
};

// The RANSAC loop. `samples` = H x 4 int32 (3 used by M_SHINJI). If `full` every one of the
// `iter_in` iterations is generated and scored (votes_out/hyps_out filled for all slots) and
// the adaptive rule is applied afterwards as a replay — which visits the same hypotheses in the
// same order with the same strict `>` and therefore returns what the early-stopping loop returns.
// votes_out: iter_in*slots ints (-1 = slot empty, -2 = not evaluated). hyps_out: iter_in*slots*7
// (qx,qy,qz,qw,tx,ty,tz). mask_out: n*cols shorts of the winner.
template <class T>
inline RansacResult<T> ransac(int method, const Corr<T>& d, const int32_t* samples, int iter_in, const Thresholds<T>& th,
                              T confidence, bool full, int* votes_out, T* hyps_out, short* mask_out) {
  const int S = method_slots(method);
  const int K = method_model_points(method);
  RansacResult<T> res;
  res.max_votes = -1;  // setMaxVotes(-1)
  res.winner = -1;
  res.iters_run = 0;
  res.evals = 0;
  int Iter = iter_in;
  if (votes_out)
    for (int i = 0; i < iter_in * S; ++i) votes_out[i] = -2;
  std::vector<int> votes_all;
  std::vector<SE3<T> > hyps_all;
  if (full) {
    votes_all.assign((size_t)iter_in * S, -1);
    hyps_all.resize((size_t)iter_in * S);
  }
  const int upper = full ? iter_in : 0;
  StaleCols<T> stale_cols;
  for (int ii = 0; ii < (full ? upper : Iter); ++ii) {
    SE3<T> hyp[3];
    bool has[3];
    generate_iteration(method, d, samples + 4 * ii, hyp, has, g_stale_sample_buffers ? &stale_cols : (StaleCols<T>*)0);
    for (int s = 0; s < S; ++s) {
      int v = -1;
      if (has[s]) {
        v = score_hypothesis(method, d, hyp[s], th, (short*)0);
        res.evals += d.n;
      }
      if (votes_out) votes_out[ii * S + s] = v;
      if (hyps_out && has[s]) {
        T* h = hyps_out + (size_t)(ii * S + s) * 7;
        h[0] = hyp[s].so3.q.x;
        h[1] = hyp[s].so3.q.y;
        h[2] = hyp[s].so3.q.z;
        h[3] = hyp[s].so3.q.w;
        h[4] = hyp[s].t[0];
        h[5] = hyp[s].t[1];
        h[6] = hyp[s].t[2];
      }
      if (full) {
        votes_all[(size_t)ii * S + s] = v;
        hyps_all[(size_t)ii * S + s] = hyp[s];
      } else if (has[s] && v > res.max_votes) {
        res.max_votes = v;
        res.best = hyp[s];
        res.winner = ii * S + s;
        Iter = ransac_update_num_iters<T>(confidence, outlier_ratio<T>(method, d.n, v), K, Iter);
      }
    }
    res.iters_run = ii + 1;
  }
  if (full) {
    for (int ii = 0; ii < Iter; ++ii) {
      for (int s = 0; s < S; ++s) {
        const int v = votes_all[(size_t)ii * S + s];
        if (v < 0) continue;
        if (v > res.max_votes) {
          res.max_votes = v;
          res.best = hyps_all[(size_t)ii * S + s];
          res.winner = ii * S + s;
          Iter = ransac_update_num_iters<T>(confidence, outlier_ratio<T>(method, d.n, v), K, Iter);
        }
      }
    }
  }
  res.iter_final = Iter;
  if (mask_out) {
    const int cols = method_mask_cols(method);
    if (res.winner >= 0)
      score_hypothesis(method, d, res.best, th, mask_out);
    else
      for (int i = 0; i < d.n * cols; ++i) mask_out[i] = 1;  // adapters start with setOnes() (PnPPoseAdapter.hpp:118-119)
  }
  return res;
}

// shinji_ls / shinji_ls1: Kabsch over the 3-D inliers (mask column 1), in index order
// (AbsoluteOrientation.hpp:273-320). flags == null -> shinji_ls2 (all points, :322-342).
template <class T>
inline SE3<T> shinji_ls(const Corr<T>& d, const short* flags3d) {
  std::vector<T> Xw, Xc;
  Xw.reserve((size_t)3 * d.n);
  Xc.reserve((size_t)3 * d.n);
  int K = 0;
  for (int i = 0; i < d.n; ++i) {
    if (flags3d && flags3d[i] != 1) continue;
    for (int r = 0; r < 3; ++r) {
      Xw.push_back(d.xw[3 * i + r]);
      Xc.push_back(d.xc[3 * i + r]);
    }
    ++K;
  }
  return shinji(Xw.data(), Xc.data(), K, K);
}

}  // namespace orc

#endif  // ORACLE_RANSAC_HPP_
