// pipeline.cu — everything around the scorer: hypothesis generation, the replay of the reference's
// sequential keep-best / adaptive-stop rule, the winner's inlier mask, and the refits.
//
// Reference code replaced (paths into /root/reference/pose):
//   generation   shinji (AbsoluteOrientation.hpp:47-99) on RandomElements samples (:113-130, :170-187)
//   replay       `if (votes > max) {...; Iter = RANSACUpdateNumIters(...)}` AbsoluteOrientation.hpp:145-151,
//                P3P.hpp:378-385, AbsoluteOrientationNormal.hpp:267-277,339-347,426-436
//   mask         setInlier(inliers) of the accepted hypothesis, PnPPoseAdapter.hpp:196-202 etc.
//   refit        shinji_ls / _ls1 / _ls2 (AbsoluteOrientation.hpp:273-342); LM on SE3 (north-star
//                addition, twin in oracle/refine.hpp)
//
// Compiled with -fmad=false like every TU of this library.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/rpe/ransac_rule.h"
#include "kernels.cuh"
#include "../../include/rpe/solvers.h"
#include "../../include/rpe/solvers_p3p.h"
#include "../../include/rpe/solvers_min.h"

namespace rpe {

// ================================================================================================
// generation
// ================================================================================================
// -R and -t for the fast scorer. R is the quaternion polynomial evaluated in binary64 and rounded
// once, so each entry is within one float rounding of the exact polynomial value.
__device__ __forceinline__ void derive_fast(const HypGen& g, HypFast* f) {
  const double x = g.q[0], y = g.q[1], z = g.q[2], w = g.q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  f->nR[0] = -(float)(1.0 - (tyy + tzz));
  f->nR[1] = -(float)(txy - twz);
  f->nR[2] = -(float)(txz + twy);
  f->nR[3] = -(float)(txy + twz);
  f->nR[4] = -(float)(1.0 - (txx + tzz));
  f->nR[5] = -(float)(tyz - twx);
  f->nR[6] = -(float)(txz - twy);
  f->nR[7] = -(float)(tyz + twx);
  f->nR[8] = -(float)(1.0 - (txx + tyy));
  f->nt[0] = -g.t[0];
  f->nt[1] = -g.t[1];
  f->nt[2] = -g.t[2];
}

__device__ __forceinline__ void publish_hypothesis(const HypGen& g, int slot, HypGen* gen, HypFast* fast, int32_t* votes,
                                                   FrameStats* st) {
  gen[slot] = g;
  HypFast f;
  derive_fast(g, &f);
  fast[slot] = f;
  votes[slot] = g.valid ? 0 : -1;
  if (g.valid) {
    const float tn = sqrtf(g.t[0] * g.t[0] + g.t[1] * g.t[1] + g.t[2] * g.t[2]) * 1.000001f;
    atomic_max_nonneg(&st->t_max_bits, tn);
  }
}

// grid.x over iterations, grid.y = slot within the iteration (so a CTA runs one solver only).
__global__ void __launch_bounds__(128)
hypgen_kernel(int method, FrameView f, const int32_t* __restrict__ samples, int H, HypGen* __restrict__ gen,
              HypFast* __restrict__ fast, int32_t* __restrict__ votes, FrameStats* __restrict__ st,
              const int32_t* __restrict__ stale_eff) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= H) return;
  const int S = method_slots(method);
  const int s = blockIdx.y;
  const int solver = method_slot_solver(method, s);
  const int32_t* sel = samples + 4 * ii;
  HypGen g;
  g.q[0] = g.q[1] = g.q[2] = 0.f;
  g.q[3] = 1.f;
  g.t[0] = g.t[1] = g.t[2] = 0.f;
  g.valid = 0;
  // a row that points outside the frame (caller's table; the reference's samplers cannot produce one) yields an empty slot
  bool in_range = true;
  for (int k = 0; k < method_sample_size(method); ++k) in_range = in_range && sel[k] >= 0 && sel[k] < f.n;
  if (!in_range) {
    publish_hypothesis(g, ii * S + s, gen, fast, votes, st);
    return;
  }
  if (solver == SOLVER_AO) {
    float Xw[9], Xc[9];
    bool all_valid = true;
    for (int k = 0; k < 3; ++k) {
      const int c = sel[k];
      const F3 pw = load_col(f.xw, c), pc = load_col(f.xc, c);
      Xw[3 * k] = pw.x;
      Xw[3 * k + 1] = pw.y;
      Xw[3 * k + 2] = pw.z;
      Xc[3 * k] = pc.x;
      Xc[3 * k + 1] = pc.y;
      Xc[3 * k + 2] = pc.z;
      all_valid = all_valid && ex_is_valid(pc);
    }
    if (all_valid) {
      const bool ok = shinji3<float>(Xw, Xc, method == RPE_SHINJI ? 3 : 4, g.q, g.t);
      g.valid = ok ? 1 : 0;
    }
  } else if (solver == SOLVER_P3P) {
    float Xw[12], bv[12];
    for (int k = 0; k < 4; ++k) {
      const int c = sel[k];
      const F3 pw = load_col(f.xw, c), b = load_col(f.bv, c);
      Xw[3 * k] = pw.x;
      Xw[3 * k + 1] = pw.y;
      Xw[3 * k + 2] = pw.z;
      bv[3 * k] = b.x;
      bv[3 * k + 1] = b.y;
      bv[3 * k + 2] = b.z;
    }
    // kneip_ransac/prosac start the 4th-point search from 1000000.0 (P3P.hpp:338,415); the
    // matrix-argument overload used by the hybrids starts from numeric_limits::max() (P3P.hpp:258)
    const float start = (method == RPE_KNEIP || method == RPE_KNEIP_QUAT) ? 1000000.0f : 3.402823466e+38f;
    g.valid = kneip_select<float>(Xw, bv, start, g.q, g.t) ? 1 : 0;
  } else {
    float pc0[3], nc0[3], pc1[3], pw0[3], nw0[3], pw1[3];
    const int c0 = sel[0], c1 = sel[1];
    // Camera-side columns: by default the sample's own (an invalid, all-NaN camera point then simply propagates NaN into
    // the translation). With rpe_set_stale_sample_columns the reference's hoisted buffers are reproduced
    // (AbsoluteOrientationNormal.hpp:48-75, 299-315): an invalid sample point leaves column k of X_c / N_c as an EARLIER
    // iteration wrote it; stale_eff[2 ii + k] is that earlier correspondence (-1: never written; the reference reads
    // uninitialised memory there, zeros here and in the oracle's StaleCols model).
    const int e0 = stale_eff ? stale_eff[2 * ii] : c0, e1 = stale_eff ? stale_eff[2 * ii + 1] : c1;
    for (int r = 0; r < 3; ++r) {
      pc0[r] = e0 >= 0 ? f.xc[3 * e0 + r] : 0.f;
      nc0[r] = e0 >= 0 ? f.nc[3 * e0 + r] : 0.f;
      pc1[r] = e1 >= 0 ? f.xc[3 * e1 + r] : 0.f;
      pw0[r] = f.xw[3 * c0 + r];
      nw0[r] = f.nw[3 * c0 + r];
      pw1[r] = f.xw[3 * c1 + r];
    }
    nl_2p<float>(pc0, nc0, pc1, pw0, nw0, pw1, g.q, g.t);
    g.valid = 1;
  }
  publish_hypothesis(g, ii * S + s, gen, fast, votes, st);
}

// MinimalSolvers.hpp as batch kernels, one problem per thread (rpe_min_ev / rpe_min_ms)
__global__ void __launch_bounds__(128) minsolv_ev_kernel(const float* __restrict__ M9, int count, float* __restrict__ E3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float m[9], e[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) m[k] = M9[9 * (size_t)i + k];
  sym3_eigenvalues<float>(m, e);
#pragma unroll
  for (int k = 0; k < 3; ++k) E3[3 * (size_t)i + k] = e[k];
}
__global__ void __launch_bounds__(128) minsolv_ms_kernel(const float* __restrict__ in24, int count, float* __restrict__ q4,
                                                         float* __restrict__ t3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float in[24], q[4], t[3];
#pragma unroll
  for (int k = 0; k < 24; ++k) in[k] = in24[24 * (size_t)i + k];
  min_solver_2pn<float>(in, q, t);
#pragma unroll
  for (int k = 0; k < 4; ++k) q4[4 * (size_t)i + k] = q[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) t3[3 * (size_t)i + k] = t[k];
}
void launch_minsolv_ev(const float* M9, int count, float* E3, cudaStream_t s) {
  if (count > 0) minsolv_ev_kernel<<<(count + 127) / 128, 128, 0, s>>>(M9, count, E3);
}
void launch_minsolv_ms(const float* in24, int count, float* q4, float* t3, cudaStream_t s) {
  if (count > 0) minsolv_ms_kernel<<<(count + 127) / 128, 128, 0, s>>>(in24, count, q4, t3);
}

void launch_hypgen(int method, const FrameView& f, const int32_t* samples_dev, int H, HypGen* gen, HypFast* fast,
                   int32_t* votes, FrameStats* st, cudaStream_t s, const int32_t* stale_eff) {
  if (H <= 0) return;
  dim3 grid((H + 127) / 128, method_slots(method));
  hypgen_kernel<<<grid, 128, 0, s>>>(method, f, samples_dev, H, gen, fast, votes, st, stale_eff);
}

// The reference's stale sample columns as a prefix scan: eff[2 ii + k] = sel[j][k] of the latest iteration j <= ii whose
// k-th sample point has a valid camera point (k = 0, 1: the columns nl_2p reads), carried across device passes in
// carry[2] (reset_carry on the frame's first pass). One warp per column; 32 iterations per step.
__device__ __forceinline__ bool corr_valid_f(const float* xc, int c) {
  return xc[3 * c] == xc[3 * c] || xc[3 * c + 1] == xc[3 * c + 1] || xc[3 * c + 2] == xc[3 * c + 2];
}
__device__ __forceinline__ bool corr_valid_f(const double* xc, int c) {
  return xc[3 * c] == xc[3 * c] || xc[3 * c + 1] == xc[3 * c + 1] || xc[3 * c + 2] == xc[3 * c + 2];
}
template <class T>
__global__ void __launch_bounds__(64) stale_cols_kernel(const int32_t* __restrict__ samples, int H, const T* __restrict__ xc, int n,
                                                        int32_t* __restrict__ carry, int reset_carry, int32_t* __restrict__ eff) {
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int last = reset_carry ? -1 : carry[k];
  for (int base = 0; base < H; base += 32) {
    const int ii = base + lane;
    int v = -1;
    if (ii < H) {
      const int c = samples[4 * ii + k];
      if (c >= 0 && c < n && corr_valid_f(xc, c)) v = c;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {  // inclusive scan with "the later valid one wins"
      const int up = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o && v < 0) v = up;
    }
    if (v < 0) v = last;
    if (ii < H) eff[2 * ii + k] = v;
    last = __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) carry[k] = last;
}
void launch_stale_cols(const int32_t* samples_dev, int H, const float* xc, const double* xc64, int n, int32_t* carry,
                       bool reset_carry, int32_t* eff, cudaStream_t s) {
  if (H <= 0) return;
  if (xc64)
    stale_cols_kernel<double><<<1, 64, 0, s>>>(samples_dev, H, xc64, n, carry, reset_carry ? 1 : 0, eff);
  else
    stale_cols_kernel<float><<<1, 64, 0, s>>>(samples_dev, H, xc, n, carry, reset_carry ? 1 : 0, eff);
}

__global__ void derive_fast_kernel(const HypGen* __restrict__ gen, HypFast* __restrict__ fast, int32_t* __restrict__ votes,
                                   int n_slots, FrameStats* __restrict__ st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  HypGen g = gen[i];
  HypFast f;
  derive_fast(g, &f);
  fast[i] = f;
  votes[i] = g.valid ? 0 : -1;
  if (g.valid) {
    const float tn = sqrtf(g.t[0] * g.t[0] + g.t[1] * g.t[1] + g.t[2] * g.t[2]) * 1.000001f;
    atomic_max_nonneg(&st->t_max_bits, tn);
  }
}
void launch_derive_fast(const HypGen* gen, HypFast* fast, int32_t* votes, int n_slots, FrameStats* st, cudaStream_t s) {
  if (n_slots <= 0) return;
  derive_fast_kernel<<<(n_slots + 127) / 128, 128, 0, s>>>(gen, fast, votes, n_slots, st);
}

// ================================================================================================
// replay of the sequential rule (one warp; all lanes run the scalar part redundantly)
// ================================================================================================
constexpr int kReplayChunk = 4096;  // vote-table entries staged in shared memory per pass
// votes/gen hold the H iterations [iter_base, iter_base + H) of this device pass; `rs` carries the sequential
// state between passes. With `finalize` the accepted hypothesis is published to `out`.
__global__ void __launch_bounds__(256)
replay_kernel(int method, const HypGen* __restrict__ gen, const int32_t* __restrict__ votes, int H, int iter_base, int n,
              float confidence, FrameStats* __restrict__ st, ReplayState* __restrict__ rs, ReplayOut* __restrict__ out,
              int finalize, int begin_iter_max) {
  __shared__ int32_t sv[kReplayChunk];
  __shared__ int s_state[5];  // best, Iter, win, cur_iter, stop
  const int lane = threadIdx.x & 31;
  const int S = method_slots(method);
  const int K = method_model_points(method);
  const int mod = method_modalities(method);
  const int E = H * S;
  const int slot_base = iter_base * S;
  if (threadIdx.x == 0) {
    if (begin_iter_max >= 0) {  // first pass of a frame: the state the reference's loop starts from
      rs->best = -1;  // setMaxVotes(-1)
      rs->iter = begin_iter_max;
      rs->win = -1;
      rs->cur_iter = -1;
      rs->stop = 0;
      rs->slots_done = 0;
      rs->borderline = 0;
      rs->overflow = 0;
      rs->q[0] = rs->q[1] = rs->q[2] = 0.f;
      rs->q[3] = 1.f;
      rs->t[0] = rs->t[1] = rs->t[2] = 0.f;
    }
    s_state[0] = rs->best;
    s_state[1] = rs->iter;
    s_state[2] = rs->win;
    s_state[3] = rs->cur_iter;
    s_state[4] = rs->stop;
  }
  __syncthreads();
  for (int cbase = 0; cbase < E; cbase += kReplayChunk) {
    if (s_state[4]) break;
    const int cn = min(kReplayChunk, E - cbase);
    for (int i = threadIdx.x; i < cn; i += blockDim.x) sv[i] = votes[cbase + i];  // all loads in flight at once
    __syncthreads();
    if (threadIdx.x < 32) {
      const float log_num = rule_log_numerator(confidence);
      int best = s_state[0], Iter = s_state[1], win = s_state[2], cur_iter = s_state[3];
      bool stop = false;
      for (int base = 0; base < cn && !stop; base += 32) {
        const int i = base + lane;
        const int v = i < cn ? sv[i] : -1;
        int pm = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int tv = __shfl_up_sync(0xffffffffu, pm, o);
          if (lane >= o) pm = max(pm, tv);
        }
        int excl = __shfl_up_sync(0xffffffffu, pm, 1);
        if (lane == 0) excl = -2147483647;
        const bool rec = v >= 0 && v > max(excl, best);
        unsigned int m = __ballot_sync(0xffffffffu, rec);
        // every candidate's log-denominator depends on its own vote count only: all lanes evaluate theirs at once
        RuleDenominator rd;
        rd.state = 0;
        rd.log_denom = 0.f;
        if (rec) rd = rule_denominator(outlier_ratio(mod, n, v), K);
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const int idx = slot_base + cbase + base + b;  // global slot index
          const int it = idx / S;
          const int vb = __shfl_sync(0xffffffffu, v, b);
          if (it != cur_iter) {
            if (it >= Iter) {  // `for (ii = 0; ii < Iter; ii++)` would not have reached this iteration
              stop = true;
              break;
            }
            cur_iter = it;
          }
          best = vb;
          win = idx;
          RuleDenominator rb_;
          rb_.state = __shfl_sync(0xffffffffu, rd.state, b);
          rb_.log_denom = __shfl_sync(0xffffffffu, rd.log_denom, b);
          Iter = rule_finish(log_num, rb_, Iter);
        }
        // nothing below the loop bound is left in (or after) this pass; a pass may end inside this 32-slot group
        if ((long long)(slot_base + cbase + min(base + 32, cn)) >= (long long)Iter * S) stop = true;
      }
      if (lane == 0) {
        s_state[0] = best;
        s_state[1] = Iter;
        s_state[2] = win;
        s_state[3] = cur_iter;
        s_state[4] = stop ? 1 : 0;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int win = s_state[2];
    if (win >= slot_base && win != rs->win) {  // accepted in this pass: keep its pose
      const HypGen g = gen[win - slot_base];
      for (int k = 0; k < 4; ++k) rs->q[k] = g.q[k];
      for (int k = 0; k < 3; ++k) rs->t[k] = g.t[k];
    }
    rs->best = s_state[0];
    rs->iter = s_state[1];
    rs->win = win;
    rs->cur_iter = s_state[3];
    rs->stop = s_state[4] || ((long long)(slot_base + E) >= (long long)s_state[1] * S) ? 1 : 0;
    rs->slots_done += E;
    rs->borderline += (int)(st->wl_count + st->wl_consumed);
    rs->overflow |= st->wl_overflow ? 1 : 0;
    // per-pass counters are consumed: leave them clean for the next pass / frame
    st->t_max_bits = 0;
    st->wl_count = 0;
    st->wl_consumed = 0;
    st->wl_overflow = 0;
    st->ticket = 0;
    st->ticket2 = 0;
    if (finalize) {
      if (win >= 0) {
        for (int k = 0; k < 4; ++k) out->q[k] = rs->q[k];
        for (int k = 0; k < 3; ++k) out->t[k] = rs->t[k];
      } else {
        out->q[0] = out->q[1] = out->q[2] = 0.f;
        out->q[3] = 1.f;
        out->t[0] = out->t[1] = out->t[2] = 0.f;
      }
      out->max_votes = rs->best;
      out->iter_final = rs->iter;
      out->winner = win;
      out->n_slots = rs->slots_done;
      out->n_borderline = rs->borderline;
      out->flags = rs->overflow ? 1 : 0;
      out->n_inliers[0] = out->n_inliers[1] = out->n_inliers[2] = 0;
      out->refit_ok = 0;
    }
  }
}

void launch_replay(int method, const HypGen* gen, const int32_t* votes, int H, int iter_base, int n, float confidence,
                   FrameStats* st, ReplayState* rs, ReplayOut* out, bool finalize, int begin_iter_max, cudaStream_t s) {
  replay_kernel<<<1, 256, 0, s>>>(method, gen, votes, H, iter_base, n, confidence, st, rs, out, finalize ? 1 : 0,
                                  begin_iter_max);
}

// ================================================================================================
// block-level deterministic reduction of NV doubles -> partials[blockIdx][NV]
// ================================================================================================
template <int NV>
__device__ __forceinline__ void block_reduce_store(double* v, double* __restrict__ partials, double* smem /*[8*NV]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) smem[warp * NV + k] = x;
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  if (threadIdx.x < kMomentCount) {
    double x = 0.0;
    if (threadIdx.x < NV)
      for (int w = 0; w < nw; ++w) x += smem[w * NV + threadIdx.x];
    partials[(size_t)blockIdx.x * kMomentCount + threadIdx.x] = x;
  }
}
// Final reduction by the LAST CTA (256 threads): thread (g = tid/32, k = tid%32) sums component k over the CTAs
// b = g, g+8, ... (independent loads, all in flight together), then the 8 groups are added in a fixed order.
// Deterministic for a given grid size. Result in smem_out[0..kMomentCount).
__device__ __forceinline__ void final_reduce_partials(const double* __restrict__ partials, int nblocks,
                                                      double* smem_scratch /*[8*32]*/, double* smem_out /*[kMomentCount]*/) {
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  for (int k0 = 0; k0 < kMomentCount; k0 += 32) {
    const int k = k0 + lane;
    double x = 0.0;
    if (k < kMomentCount)
      for (int b = g; b < nblocks; b += 8) x += partials[(size_t)b * kMomentCount + k];
    smem_scratch[g * 32 + lane] = x;
    __syncthreads();
    if (threadIdx.x < 32 && k < kMomentCount) {
      double t = 0.0;
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) t += smem_scratch[gg * 32 + threadIdx.x];
      smem_out[k] = t;
    }
    __syncthreads();
  }
}

// The rotation of a cross-covariance with three healthy singular values and a positive determinant is its orthogonal
// polar factor U V^T. Scaled Newton iteration X <- (g X + X^-T / g) / 2, g^2 = |X^-1|_F / |X|_F (Higham): quadratic
// convergence, ~6 steps of one 3x3 inverse each — about a sixth of the instructions of the two-sided Jacobi SVD a single
// thread otherwise walks at the end of the mask kernel (profiles/r02m_small_kernels.md). Accepted only if the result is
// orthogonal to 1e-13 with determinant +1; reflections (det M <= 0), rank-deficient and badly conditioned covariances
// (the K = 3 samples of the generator never come here) return false and take the SVD path. Agreement with the SVD route:
// < 1e-12 in R (tests/test_gpu_ao.py refits, tools/fuzz_refit.py), against a 1e-6 rad bar.
__device__ int g_kabsch_polar = 1;  // 0: always the Jacobi SVD (rpe_debug_set_kabsch_polar, A/B in the tests)
void set_kabsch_polar(int on) {
  const int v = on ? 1 : 0;
  cudaMemcpyToSymbol(g_kabsch_polar, &v, sizeof(int));
}
__device__ bool polar_rotation_newton(const double* M, double* R) {
  double X[9], fro2 = 0.0;
  for (int i = 0; i < 9; ++i) fro2 += M[i] * M[i];
  if (!(fro2 > 0.0) || !(fro2 < 1e300)) return false;
  const double inv_fro = 1.0 / sqrt(fro2);
  for (int i = 0; i < 9; ++i) X[i] = M[i] * inv_fro;  // singular values now in (0, 1]
  for (int it = 0; it < 16; ++it) {
    double C[9];  // cofactors: X^-T = C / det
    C[0] = X[4] * X[8] - X[5] * X[7];
    C[1] = X[5] * X[6] - X[3] * X[8];
    C[2] = X[3] * X[7] - X[4] * X[6];
    C[3] = X[2] * X[7] - X[1] * X[8];
    C[4] = X[0] * X[8] - X[2] * X[6];
    C[5] = X[1] * X[6] - X[0] * X[7];
    C[6] = X[1] * X[5] - X[2] * X[4];
    C[7] = X[2] * X[3] - X[0] * X[5];
    C[8] = X[0] * X[4] - X[1] * X[3];
    const double det = X[0] * C[0] + X[1] * C[1] + X[2] * C[2];
    if (it == 0 && !(det > 1e-7)) return false;  // sigma_1 sigma_2 sigma_3 / |M|_F^3: reflection, rank loss or cond >~ 1e6
    if (!(det > 0.0)) return false;
    double nx = 0.0, nc = 0.0;
    for (int i = 0; i < 9; ++i) {
      nx += X[i] * X[i];
      nc += C[i] * C[i];
    }
    const double g2 = sqrt(nc / nx) / det;  // |X^-1|_F / |X|_F
    const double g = sqrt(g2);
    const double a = 0.5 * g, b = 0.5 / (g * det);
    double change = 0.0;
    for (int i = 0; i < 9; ++i) {
      const double xn = a * X[i] + b * C[i];
      const double d = xn - X[i];
      change += d * d;
      X[i] = xn;
    }
    if (change < 1e-30) break;
  }
  // orthogonality and orientation of what came out
  double worst = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) {
      const double d = X[i] * X[j] + X[3 + i] * X[3 + j] + X[6 + i] * X[6 + j] - (i == j ? 1.0 : 0.0);
      worst = fmax(worst, fabs(d));
    }
  const double det = X[0] * (X[4] * X[8] - X[5] * X[7]) - X[1] * (X[3] * X[8] - X[5] * X[6]) + X[2] * (X[3] * X[7] - X[4] * X[6]);
  if (!(worst < 1e-13) || !(fabs(det - 1.0) < 1e-12)) return false;
  for (int i = 0; i < 9; ++i) R[i] = X[i];
  return true;
}

// Kabsch from moments m[0]=K, m[1..3]=sum x_w, m[4..6]=sum x_c, m[7..15]=sum x_c x_w^T (row-major).
// Same closed form as shinji (centroids, cross-covariance / K, SVD, det fix, t = c_c - R c_w), evaluated in binary64.
__device__ bool kabsch_from_moments(const double* m, float* q_out, float* t_out) {
  const double K = m[0];
  if (!(K >= 3.0)) return false;  // reference asserts 3 <= K (AbsoluteOrientation.hpp:53)
  double cw[3], cc[3];
  for (int r = 0; r < 3; ++r) {
    cw[r] = m[1 + r] / K;
    cc[r] = m[4 + r] / K;
  }
  double M[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[3 * i + j] = (m[7 + 3 * i + j] - K * cc[i] * cw[j]) / K;
  double q[4], Rp[9];
  const bool ok = (g_kabsch_polar && polar_rotation_newton(M, Rp)) ? so3_from_matrix<double>(Rp, q) : rotation_from_covariance<double>(M, q);
  double rc[3];
  quat_rotate<double>(q, cw, rc);
  for (int k = 0; k < 4; ++k) q_out[k] = (float)q[k];
  for (int r = 0; r < 3; ++r) t_out[r] = (float)(cc[r] - rc[r]);
  return ok;
}

// Sufficient statistics of the 3-D and normal rows (binary64 sums of binary32 inputs). Every quantity the Kabsch refit
// and the LM normal equations of those rows need is a polynomial in (R, t) with these coefficients:
//   [0] n3  [1..3] sum x_w  [4..6] sum x_c  [7..15] sum x_c x_w^T (row-major)  [16..21] sum x_w x_w^T (xx,xy,xz,yy,yz,zz)
//   [22] sum |x_c|^2        [23] nn  [24..29] sum n_w n_w^T  [30..38] sum n_c n_w^T (row-major)  [39] sum |n_c|^2
constexpr int kSuff3 = 23, kSuffAll = 40;
__device__ __forceinline__ void suff_add_3d(double* m, const F3& xw, const F3& xc) {
  const double w[3] = {xw.x, xw.y, xw.z}, c[3] = {xc.x, xc.y, xc.z};
  m[0] += 1.0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    m[1 + r] += w[r];
    m[4 + r] += c[r];
#pragma unroll
    for (int q = 0; q < 3; ++q) m[7 + 3 * r + q] += c[r] * w[q];
  }
  m[16] += w[0] * w[0];
  m[17] += w[0] * w[1];
  m[18] += w[0] * w[2];
  m[19] += w[1] * w[1];
  m[20] += w[1] * w[2];
  m[21] += w[2] * w[2];
  m[22] += c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
}
__device__ __forceinline__ void suff_add_nl(double* m, const F3& nw, const F3& nc) {
  const double w[3] = {nw.x, nw.y, nw.z}, c[3] = {nc.x, nc.y, nc.z};
  m[23] += 1.0;
  m[24] += w[0] * w[0];
  m[25] += w[0] * w[1];
  m[26] += w[0] * w[2];
  m[27] += w[1] * w[1];
  m[28] += w[1] * w[2];
  m[29] += w[2] * w[2];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) m[30 + 3 * r + q] += c[r] * w[q];
  m[39] += c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
}

// ================================================================================================
// winner's mask (+ Kabsch moments of the 3-D inliers, + the Kabsch solve in the last CTA)
// ================================================================================================
template <bool UN>
__global__ void __launch_bounds__(256, 2)
mask_kernel(int method, FrameView f, ReplayOut* pose_rw, Thresh th, int16_t* __restrict__ mask, RefitBuffers rb,
            ReplayOut* kabsch_out, FrameStats* st, uint32_t* __restrict__ bits) {
  // bits (optional): the same flags as one bit per correspondence, cols x ceil(n / 32) words, column after column —
  // what rpe_set_mask_transfer(1) sends to the host instead of the 16-bit matrix
  const ReplayOut* pose = pose_rw;
  constexpr int NV = UN ? kSuffAll : kSuff3;
  __shared__ double red[8 * NV];
  __shared__ int cnts[3];
  __shared__ bool is_last;
  const int n = f.n;
  const int cols = method_mask_cols(method);
  if (threadIdx.x < 3) cnts[threadIdx.x] = 0;
  __syncthreads();
  HypGen h;
  for (int k = 0; k < 4; ++k) h.q[k] = pose->q[k];
  for (int k = 0; k < 3; ++k) h.t[k] = pose->t[k];
  const bool have = pose->winner >= 0;
  float Rm[9];
  ex_quat_to_matrix(h.q, Rm);
  const bool u2 = method_uses_2d(method), u3 = method_uses_3d(method), un = UN;
  double mom[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) mom[k] = 0.0;
  int c2 = 0, c3 = 0, cn = 0;
  const int wpc = (n + 31) >> 5;  // words per column of the bit form
  // (the loop condition is uniform over a warp: its 32 lanes hold 32 consecutive correspondences starting at a multiple of 32)
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c - (int)(threadIdx.x & 31) < n; c += gridDim.x * blockDim.x) {
    const bool inb = c < n;
    bool f2 = false, f3d = false, fn = false;
    if (!inb) {
      // beyond the frame: contributes nothing
    } else if (have) {
      const F3 xw = load_col(f.xw, c);
      bool valid = false;
      F3 xc = f3(0.f, 0.f, 0.f);
      if (u3 || un) {
        xc = load_col(f.xc, c);
        valid = ex_is_valid(xc);
      }
      if (UN && valid) {
        const F3 nw = load_col(f.nw, c), nc = load_col(f.nc, c);
        fn = ex_test_nl(h.q, nw, nc, th.cos_nl);
        if (fn) suff_add_nl(mom, nw, nc);
      }
      if (u3 && valid) f3d = ex_test_3d(h.q, h.t, xw, xc, th.thr3d);
      if (u2) f2 = ex_test_2d(h.q, h.t, method == RPE_KNEIP ? Rm : nullptr, xw, load_col(f.bv, c), th.cos_thr);
      if (f3d) suff_add_3d(mom, xw, xc);
      c2 += f2 ? 1 : 0;
      c3 += f3d ? 1 : 0;
      cn += fn ? 1 : 0;
    } else {
      f2 = f3d = fn = true;  // adapters start with setOnes() and setInlier is never called
      // keep the statistics consistent with the all-ones columns a later refinement would read
      if (cols >= 2 && f.xc) suff_add_3d(mom, load_col(f.xw, c), load_col(f.xc, c));
      if (UN && f.nw && f.nc) suff_add_nl(mom, load_col(f.nw, c), load_col(f.nc, c));
    }
    // column 0 of a family without the 2-D test is never set by the reference's loop: 0 once a hypothesis has been
    // accepted (the per-iteration matrix starts from zero), still the initial 1 when nothing was accepted
    const bool v0 = !have ? true : ((cols == 1 || u2) ? f2 : false);
    const bool v1 = u3 ? f3d : !have;
    const bool v2 = fn;
    if (inb) {
      mask[c] = (int16_t)(v0 ? 1 : 0);
      if (cols >= 2) mask[n + c] = (int16_t)(v1 ? 1 : 0);
      if (cols >= 3) mask[2 * n + c] = (int16_t)(v2 ? 1 : 0);
    }
    if (bits) {
      const unsigned int b0 = __ballot_sync(0xffffffffu, inb && v0);
      const unsigned int b1 = __ballot_sync(0xffffffffu, inb && v1);
      const unsigned int b2 = __ballot_sync(0xffffffffu, inb && v2);
      if ((threadIdx.x & 31) == 0) {
        const int w = c >> 5;
        bits[w] = b0;
        if (cols >= 2) bits[wpc + w] = b1;
        if (cols >= 3) bits[2 * wpc + w] = b2;
      }
    }
  }
  // per-column counts
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    c3 += __shfl_xor_sync(0xffffffffu, c3, o);
    cn += __shfl_xor_sync(0xffffffffu, cn, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c2) atomicAdd(&cnts[0], c2);
    if (c3) atomicAdd(&cnts[1], c3);
    if (cn) atomicAdd(&cnts[2], cn);
  }
  block_reduce_store<NV>(mom, rb.partials, red);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; ++k)
      if (cnts[k]) atomicAdd(&pose_rw->n_inliers[k], cnts[k]);
    __threadfence();
    const unsigned int t = atomicAdd(&st->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    __shared__ double fin_scratch[8 * 32];
    __shared__ double fin[kMomentCount];
    final_reduce_partials(rb.partials, gridDim.x, fin_scratch, fin);
    if (threadIdx.x < kSuffAll) rb.suff[threadIdx.x] = fin[threadIdx.x];
    if (threadIdx.x == 0) {
      double m[16];
      for (int k = 0; k < 16; ++k) m[k] = fin[k];
      for (int k = 0; k < 16; ++k) rb.moments[k] = m[k];
      ReplayOut o = *pose_rw;
      float q[4], t[3];
      bool ok = false;
      if (u3 && have) ok = kabsch_from_moments(m, q, t);
      if (ok) {
        for (int k = 0; k < 4; ++k) o.q[k] = q[k];
        for (int k = 0; k < 3; ++k) o.t[k] = t[k];
      }
      o.refit_ok = ok ? 1 : 0;
      *kabsch_out = o;
      st->ticket = 0;
    }
  }
}

static int refit_grid(int n, int num_sms_hint) {
  const int full = (n + 255) / 256;
  const int cap = 2 * (num_sms_hint > 0 ? num_sms_hint : 148);
  return full < cap ? full : cap;
}
void launch_mask(int method, const FrameView& f, ReplayOut* pose_rw, Thresh th, int16_t* mask, ReplayOut* kabsch_out,
                 RefitBuffers rb, FrameStats* st, cudaStream_t s, uint32_t* bits) {
  if (method_uses_nl(method))
    mask_kernel<true><<<refit_grid(f.n, rb.num_sms), 256, 0, s>>>(method, f, pose_rw, th, mask, rb, kabsch_out, st, bits);
  else
    mask_kernel<false><<<refit_grid(f.n, rb.num_sms), 256, 0, s>>>(method, f, pose_rw, th, mask, rb, kabsch_out, st, bits);
}

// Stand-alone statistics over explicit flag columns (after rpe_set_mask) or over all points (shinji_ls2).
__global__ void __launch_bounds__(256)
suffstat_kernel(FrameView f, const int16_t* __restrict__ flags3d, int all3d, const int16_t* __restrict__ flagsN,
                RefitBuffers rb) {
  __shared__ double red[8 * kSuffAll];
  double mom[kSuffAll];
#pragma unroll
  for (int k = 0; k < kSuffAll; ++k) mom[k] = 0.0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < f.n; c += gridDim.x * blockDim.x) {
    if (all3d || (flags3d && flags3d[c] == 1)) suff_add_3d(mom, load_col(f.xw, c), load_col(f.xc, c));
    if (flagsN && flagsN[c] == 1) suff_add_nl(mom, load_col(f.nw, c), load_col(f.nc, c));
  }
  block_reduce_store<kSuffAll>(mom, rb.partials, red);
}
int launch_suffstats(const FrameView& f, const int16_t* flags3d, bool all3d, const int16_t* flagsN, RefitBuffers rb,
                     cudaStream_t s) {
  const int blocks = refit_grid(f.n, rb.num_sms);
  suffstat_kernel<<<blocks, 256, 0, s>>>(f, flags3d, all3d ? 1 : 0, flagsN, rb);
  return blocks;
}
int launch_kabsch_moments(const FrameView& f, const int16_t* flags3d, RefitBuffers rb, FrameStats* st, cudaStream_t s) {
  (void)st;
  return launch_suffstats(f, flags3d, flags3d == nullptr, nullptr, rb, s);
}
__global__ void __launch_bounds__(256)
kabsch_solve_kernel(RefitBuffers rb, int blocks_used, ReplayOut* __restrict__ pose, int32_t* refit_ok) {
  __shared__ double fin_scratch[8 * 32];
  __shared__ double fin[kMomentCount];
  final_reduce_partials(rb.partials, blocks_used, fin_scratch, fin);
  if (threadIdx.x == 0) {
    double m[16];
    for (int k = 0; k < 16; ++k) m[k] = fin[k];
    for (int k = 0; k < 16; ++k) rb.moments[k] = m[k];
    float q[4], t[3];
    const bool ok = kabsch_from_moments(m, q, t);
    if (ok) {
      for (int k = 0; k < 4; ++k) pose->q[k] = q[k];
      for (int k = 0; k < 3; ++k) pose->t[k] = t[k];
    }
    pose->refit_ok = ok ? 1 : 0;
    if (refit_ok) *refit_ok = ok ? 1 : 0;
  }
}
void launch_kabsch_solve(RefitBuffers rb, int blocks_used, ReplayOut* pose_inout, int32_t* refit_ok, cudaStream_t s) {
  kabsch_solve_kernel<<<1, 256, 0, s>>>(rb, blocks_used, pose_inout, refit_ok);
}

// ================================================================================================
// LM / Gauss-Newton on SE(3): fused residual + Jacobian + normal equations, FP64 reductions,
// 6x6 Cholesky and the SE3 exponential in the last CTA. Twin of oracle/refine.hpp::refine_gn.
// ================================================================================================
__device__ void gn_state_init(const ReplayOut* __restrict__ pose, GnState* __restrict__ gs) {
  double q[4] = {pose->q[0], pose->q[1], pose->q[2], pose->q[3]};
  const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; ++k) q[k] /= qn;
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  double* R = gs->Rp;
  R[0] = 1.0 - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1.0 - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1.0 - (txx + tyy);
  for (int k = 0; k < 3; ++k) gs->tp[k] = pose->t[k];
  for (int k = 0; k < 9; ++k) gs->Ra[k] = R[k];
  for (int k = 0; k < 3; ++k) gs->ta[k] = gs->tp[k];
  gs->cost_acc = 0.0;
  gs->mu = 1e-4;
  gs->rows = 0;
  gs->have = 0;
  gs->done = 0;
  gs->evals = 0;
  gs->accepted = 0;
}
__global__ void gn_init_kernel(const ReplayOut* __restrict__ pose, GnState* __restrict__ gs) {
  if (threadIdx.x == 0) gn_state_init(pose, gs);
}
void launch_gn_init(const ReplayOut* pose, GnState* st, cudaStream_t s) { gn_init_kernel<<<1, 32, 0, s>>>(pose, st); }

// acc layout: [0..20] upper triangle of J^T J, [21..26] J^T r, [27] cost, [28] rows
__device__ __forceinline__ void gn_add_row(double* acc, const double J[6], double r, double w) {
  int k = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = i; j < 6; ++j) acc[k++] += w * J[i] * J[j];
    acc[21 + i] += w * J[i] * r;
  }
  acc[27] += w * r * r;
  acc[28] += 1.0;
}
__device__ __forceinline__ void gn_point_rows(const double y[3], double J[3][6]) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) J[r][c] = 0.0;
  J[0][0] = J[1][1] = J[2][2] = 1.0;
  J[0][4] = y[2];
  J[0][5] = -y[1];
  J[1][3] = -y[2];
  J[1][5] = y[0];
  J[2][3] = y[1];
  J[2][4] = -y[0];
}

__device__ bool gn_solve(const double* Hu, const double* g, double mu, double* delta) {
  // fully unrolled so that A and L stay in registers (the single-thread tail of the refinement is latency-bound)
  double A[6][6], L[6][6];
  {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = i; j < 6; ++j) {
        A[i][j] = A[j][i] = Hu[k++];
      }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) A[i][i] += mu * A[i][i];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) L[i][j] = 0.0;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = A[j][j];
#pragma unroll
    for (int p = 0; p < j; ++p) s -= L[j][p] * L[j][p];
    if (!(s > 0.0)) ok = false;
    L[j][j] = sqrt(s);
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = A[i][j];
#pragma unroll
      for (int p = 0; p < j; ++p) v -= L[i][p] * L[j][p];
      L[i][j] = v / L[j][j];
    }
  }
  if (!ok) return false;
  double z[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double v = -g[i];
#pragma unroll
    for (int p = 0; p < i; ++p) v -= L[i][p] * z[p];
    z[i] = v / L[i][i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double v = z[i];
#pragma unroll
    for (int p = i + 1; p < 6; ++p) v -= L[p][i] * delta[p];
    delta[i] = v / L[i][i];
  }
  return true;
}

// T <- exp(delta) T (sophus/se3.hpp:321-342)
__device__ void se3_exp_left(const double* delta, double* R, double* t) {
  const double v[3] = {delta[0], delta[1], delta[2]};
  const double w[3] = {delta[3], delta[4], delta[5]};
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double A, B, C;
  if (th < 1e-6) {
    A = 1.0 - th2 / 6.0;
    B = 0.5 - th2 / 24.0;
    C = 1.0 / 6.0 - th2 / 120.0;
  } else {
    A = sin(th) / th;
    B = (1.0 - cos(th)) / th2;
    C = (th - sin(th)) / (th2 * th);
  }
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double W2[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) W2[3 * r + c] = W[3 * r] * W[c] + W[3 * r + 1] * W[3 + c] + W[3 * r + 2] * W[6 + c];
  double Rd[9], V[9];
  for (int i = 0; i < 9; ++i) {
    const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    Rd[i] = I + A * W[i] + B * W2[i];
    V[i] = I + B * W[i] + C * W2[i];
  }
  double Rn[9], tn[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Rn[3 * r + c] = Rd[3 * r] * R[c] + Rd[3 * r + 1] * R[3 + c] + Rd[3 * r + 2] * R[6 + c];
    tn[r] = Rd[3 * r] * t[0] + Rd[3 * r + 1] * t[1] + Rd[3 * r + 2] * t[2] + V[3 * r] * v[0] + V[3 * r + 1] * v[1] +
            V[3 * r + 2] * v[2];
  }
  for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  for (int i = 0; i < 3; ++i) t[i] = tn[i];
}

// Accumulator layouts (per thread, FP64):
//   GENERIC (any modality, needed for the 2-D rows): [0..20] upper triangle of J^T J, [21..26] J^T r,
//            [27] cost, [28] rows
//   MOMENTS (3-D and normal rows only; their normal equations are polynomial in y = R x + t, m = R n):
//            [0] sum w3, [1..3] sum w3 y, [4..9] sum w3 y y^T (xx,xy,xz,yy,yz,zz), [10..12] sum w3 r,
//            [13..15] sum w3 (y x r), [16..21] sum wn m m^T, [22..24] sum wn (m x rn), [25] cost, [26] rows
//   J^T J = [[S I, -[Sy]x], [., (tr Syy) I - Syy + (tr Smm) I - Smm]],  J^T r = [Sr ; Syxr + Smxr]
constexpr int kGnAcc = 29;

__device__ __forceinline__ void gn_moments_to_normal_eq(const double* a, double* H21, double* g6, double* cost,
                                                        double* rows) {
  double H[6][6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) H[i][j] = 0.0;
  H[0][0] = H[1][1] = H[2][2] = a[0];
  const double sx = a[1], sy = a[2], sz = a[3];
  H[0][4] = sz;
  H[0][5] = -sy;
  H[1][3] = -sz;
  H[1][5] = sx;
  H[2][3] = sy;
  H[2][4] = -sx;
  const double yy[6] = {a[4] + a[16], a[5] + a[17], a[6] + a[18], a[7] + a[19], a[8] + a[20], a[9] + a[21]};
  const double tr = yy[0] + yy[3] + yy[5];
  H[3][3] = tr - yy[0];
  H[3][4] = -yy[1];
  H[3][5] = -yy[2];
  H[4][4] = tr - yy[3];
  H[4][5] = -yy[4];
  H[5][5] = tr - yy[5];
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) H21[k++] = H[i][j];
  g6[0] = a[10];
  g6[1] = a[11];
  g6[2] = a[12];
  g6[3] = a[13] + a[22];
  g6[4] = a[14] + a[23];
  g6[5] = a[15] + a[24];
  *cost = a[25];
  *rows = a[26];
}

// One LM evaluation's bookkeeping (oracle/refine.hpp::refine_gn): accept / reject the proposal whose normal
// equations are (Hn, gn, cost, rows), then solve for the next proposal.
__device__ void gn_lm_update(GnState* gs, const double* Hn, const double* gn, double cost, double rows) {
  gs->evals += 1;
  bool finished = false;
  if (!gs->have || cost < gs->cost_acc) {
    for (int k = 0; k < 9; ++k) gs->Ra[k] = gs->Rp[k];
    for (int k = 0; k < 3; ++k) gs->ta[k] = gs->tp[k];
    for (int k = 0; k < 21; ++k) gs->H[k] = Hn[k];
    for (int k = 0; k < 6; ++k) gs->g[k] = gn[k];
    gs->cost_acc = cost;
    gs->rows = (long long)rows;
    if (gs->have) {
      const double mu = gs->mu * 0.1;
      gs->mu = mu < 1e-12 ? 1e-12 : mu;
      gs->accepted += 1;
    }
    gs->have = 1;
  } else {
    gs->mu = gs->mu * 10.0;
  }
  if (gs->rows < 6) finished = true;
  if (!finished) {
    double delta[6];
    int tries = 0;
    double mu = gs->mu;
    while (!gn_solve(gs->H, gs->g, mu, delta) && tries < 8) {
      mu *= 10.0;
      ++tries;
    }
    gs->mu = mu;
    if (tries == 8) {
      finished = true;
    } else {
      double mx = 0.0;
      for (int k = 0; k < 6; ++k) mx = fabs(delta[k]) > mx ? fabs(delta[k]) : mx;
      if (mx < 1e-10) {
        finished = true;
      } else {
        for (int k = 0; k < 9; ++k) gs->Rp[k] = gs->Ra[k];
        for (int k = 0; k < 3; ++k) gs->tp[k] = gs->ta[k];
        se3_exp_left(delta, gs->Rp, gs->tp);
      }
    }
  }
  if (finished) gs->done = 1;
}
__device__ void gn_publish(const GnState* gs, ReplayOut* pose_out, double* cost_out, int32_t* evals_out) {
  // publish the best pose accepted so far (the final answer if no later evaluation improves on it)
  {
    double q[4];
    so3_from_matrix<double>(gs->Ra, q);
    const double nn = sqrt((q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3]));
    for (int k = 0; k < 4; ++k) pose_out->q[k] = (float)(q[k] / nn);
    for (int k = 0; k < 3; ++k) pose_out->t[k] = (float)gs->ta[k];
    pose_out->refit_ok = 1;
    *cost_out = gs->cost_acc;
    *evals_out = gs->evals;
  }
}

// One LM evaluation over the correspondences (needed when 2-D rows take part: they are not polynomial in the pose).
// Rows are formed in binary64 from the binary32 inputs with the twin's expressions (oracle/refine.hpp::gn_evaluate),
// so a row has the same bits on both sides and only the summation order differs.
template <bool GENERIC>
__global__ void __launch_bounds__(256)
gn_iteration_kernel(FrameView f, const int16_t* __restrict__ mask, int mask_cols, float w2d, float w3d, float wnl,
                    RefitBuffers rb, GnState* __restrict__ gs, FrameStats* __restrict__ st, ReplayOut* __restrict__ pose_out,
                    double* __restrict__ cost_out, int32_t* __restrict__ evals_out) {
  static_assert(GENERIC, "3-D / normal rows alone are handled by gn_from_stats_kernel");
  if (gs->done) return;
  __shared__ double red[8 * kGnAcc];
  __shared__ bool is_last;
  const int n = f.n;
  double acc[kGnAcc];
#pragma unroll
  for (int k = 0; k < kGnAcc; ++k) acc[k] = 0.0;
  double R[9], t[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = gs->Rp[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = gs->tp[k];
  const bool m2 = mask_cols >= 1 && f.bv && w2d > 0.f;
  const bool m3 = mask_cols >= 2 && f.xc && w3d > 0.f;
  const bool mn = mask_cols >= 3 && f.nc && wnl > 0.f;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    const bool u2 = m2 && mask[c] == 1;
    const bool u3 = m3 && mask[n + c] == 1;
    const bool un = mn && mask[2 * n + c] == 1;
    if (!(u2 || u3 || un)) continue;
    const F3 xf = load_col(f.xw, c);
    const double x[3] = {(double)xf.x, (double)xf.y, (double)xf.z};
    double y[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) y[r] = R[3 * r] * x[0] + R[3 * r + 1] * x[1] + R[3 * r + 2] * x[2] + t[r];
    double Jy[3][6];
    gn_point_rows(y, Jy);
    if (u3) {
      const F3 p = load_col(f.xc, c);
      const double pd[3] = {(double)p.x, (double)p.y, (double)p.z};
#pragma unroll
      for (int r = 0; r < 3; ++r) gn_add_row(acc, Jy[r], y[r] - pd[r], (double)w3d);
    }
    if (u2) {
      const F3 bf = load_col(f.bv, c);
      const double b[3] = {(double)bf.x, (double)bf.y, (double)bf.z};
      const double ny = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
      const double u[3] = {y[0] / ny, y[1] / ny, y[2] / ny};
      double P[3][3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) P[r][k] = ((r == k ? 1.0 : 0.0) - u[r] * u[k]) / ny;
      double Ju[3][6];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) Ju[r][cc] = P[r][0] * Jy[0][cc] + P[r][1] * Jy[1][cc] + P[r][2] * Jy[2][cc];
      const double res[3] = {b[1] * u[2] - b[2] * u[1], b[2] * u[0] - b[0] * u[2], b[0] * u[1] - b[1] * u[0]};
      double Jr[3][6];
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        Jr[0][cc] = b[1] * Ju[2][cc] - b[2] * Ju[1][cc];
        Jr[1][cc] = b[2] * Ju[0][cc] - b[0] * Ju[2][cc];
        Jr[2][cc] = b[0] * Ju[1][cc] - b[1] * Ju[0][cc];
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) gn_add_row(acc, Jr[r], res[r], (double)w2d);
    }
    if (un) {
      const F3 nwf = load_col(f.nw, c), ncf = load_col(f.nc, c);
      const double nw[3] = {(double)nwf.x, (double)nwf.y, (double)nwf.z}, nc[3] = {(double)ncf.x, (double)ncf.y, (double)ncf.z};
      double m[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) m[r] = R[3 * r] * nw[0] + R[3 * r + 1] * nw[1] + R[3 * r + 2] * nw[2];
      double Jn[3][6];
      gn_point_rows(m, Jn);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        Jn[r][0] = Jn[r][1] = Jn[r][2] = 0.0;  // normals do not translate
        gn_add_row(acc, Jn[r], m[r] - nc[r], (double)wnl);
      }
    }
  }
  block_reduce_store<kGnAcc>(acc, rb.partials, red);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int tk = atomicAdd(&st->ticket2, 1u);
    is_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  __shared__ double fin_scratch[8 * 32];
  __shared__ double fin[kMomentCount];
  final_reduce_partials(rb.partials, gridDim.x, fin_scratch, fin);
  if (threadIdx.x != 0) return;
  double tot[kGnAcc];
  for (int k = 0; k < kGnAcc; ++k) tot[k] = fin[k];
  st->ticket2 = 0;
  double Hn[21], gn[6], cost, rows;
  if (GENERIC) {
    for (int k = 0; k < 21; ++k) Hn[k] = tot[k];
    for (int k = 0; k < 6; ++k) gn[k] = tot[21 + k];
    cost = tot[27];
    rows = tot[28];
  } else {
    gn_moments_to_normal_eq(tot, Hn, gn, &cost, &rows);
  }
  gn_lm_update(gs, Hn, gn, cost, rows);
  gn_publish(gs, pose_out, cost_out, evals_out);
}

// ---- 3-D / normal rows only: the whole LM loop from the sufficient statistics, no further pass over the data ----
// With y = R x + t, r = y - p (3-D) and m = R n_w, r = m - n_c (normals) every entry of the MOMENTS layout above is a
// polynomial in (R, t) whose coefficients are the statistics S (suff_add_3d / suff_add_nl):
//   sum y = R sx + n t              sum y y^T = R Sxx R^T + (R sx) t^T + t (R sx)^T + n t t^T
//   sum r = sum y - sp              sum y x r = -sum y x p = axial(C - C^T), C = Spx R^T + sp t^T (C_jk = sum p_j y_k)
//   cost  = tr(sum y y^T) - 2 tr C + spp          (normals: the same with t = 0 and Snn, Scn, scc)
// The residuals are evaluated in binary64 here (the per-row kernel rounds them to binary32 first); the two
// agree to the rounding of the rows, far inside the 1e-6 refinement tolerance.
__device__ void gn_eval_from_stats(const double* S, double w3, double wn, const double* R, const double* t, double* a) {
  for (int k = 0; k < kGnAcc; ++k) a[k] = 0.0;
  auto sym = [](const double* u, int i, int j) -> double {  // (xx,xy,xz,yy,yz,zz) -> entry (i,j)
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    return u[lo == 0 ? hi : (lo == 1 ? 2 + hi : 5)];
  };
  if (w3 > 0.0) {
    const double n = S[0];
    const double* sx = S + 1;
    const double* sp = S + 4;
    const double* Spx = S + 7;
    const double* Sxx = S + 16;
    double Rsx[3], Sy[3];
    for (int i = 0; i < 3; ++i) {
      Rsx[i] = R[3 * i] * sx[0] + R[3 * i + 1] * sx[1] + R[3 * i + 2] * sx[2];
      Sy[i] = Rsx[i] + n * t[i];
    }
    double RS[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        RS[3 * i + j] = R[3 * i] * sym(Sxx, 0, j) + R[3 * i + 1] * sym(Sxx, 1, j) + R[3 * i + 2] * sym(Sxx, 2, j);
    double Syy[9], C[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        Syy[3 * i + j] = RS[3 * i] * R[3 * j] + RS[3 * i + 1] * R[3 * j + 1] + RS[3 * i + 2] * R[3 * j + 2] +
                         Rsx[i] * t[j] + t[i] * Rsx[j] + n * t[i] * t[j];
        C[3 * i + j] = Spx[3 * i] * R[3 * j] + Spx[3 * i + 1] * R[3 * j + 1] + Spx[3 * i + 2] * R[3 * j + 2] + sp[i] * t[j];
      }
    a[0] = w3 * n;
    for (int i = 0; i < 3; ++i) a[1 + i] = w3 * Sy[i];
    a[4] = w3 * Syy[0];
    a[5] = w3 * Syy[1];
    a[6] = w3 * Syy[2];
    a[7] = w3 * Syy[4];
    a[8] = w3 * Syy[5];
    a[9] = w3 * Syy[8];
    for (int i = 0; i < 3; ++i) a[10 + i] = w3 * (Sy[i] - sp[i]);
    a[13] = w3 * (C[3 * 1 + 2] - C[3 * 2 + 1]);
    a[14] = w3 * (C[3 * 2 + 0] - C[3 * 0 + 2]);
    a[15] = w3 * (C[3 * 0 + 1] - C[3 * 1 + 0]);
    a[25] += w3 * ((Syy[0] + Syy[4] + Syy[8]) - 2.0 * (C[0] + C[4] + C[8]) + S[22]);
    a[26] += 3.0 * n;
  }
  if (wn > 0.0) {
    const double nn = S[23];
    const double* Snn = S + 24;
    const double* Scn = S + 30;
    double RS[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        RS[3 * i + j] = R[3 * i] * sym(Snn, 0, j) + R[3 * i + 1] * sym(Snn, 1, j) + R[3 * i + 2] * sym(Snn, 2, j);
    double Smm[9], D[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        Smm[3 * i + j] = RS[3 * i] * R[3 * j] + RS[3 * i + 1] * R[3 * j + 1] + RS[3 * i + 2] * R[3 * j + 2];
        D[3 * i + j] = Scn[3 * i] * R[3 * j] + Scn[3 * i + 1] * R[3 * j + 1] + Scn[3 * i + 2] * R[3 * j + 2];
      }
    a[16] = wn * Smm[0];
    a[17] = wn * Smm[1];
    a[18] = wn * Smm[2];
    a[19] = wn * Smm[4];
    a[20] = wn * Smm[5];
    a[21] = wn * Smm[8];
    a[22] = wn * (D[3 * 1 + 2] - D[3 * 2 + 1]);
    a[23] = wn * (D[3 * 2 + 0] - D[3 * 0 + 2]);
    a[24] = wn * (D[3 * 0 + 1] - D[3 * 1 + 0]);
    a[25] += wn * ((Smm[0] + Smm[4] + Smm[8]) - 2.0 * (D[0] + D[4] + D[8]) + S[39]);
    a[26] += 3.0 * nn;
  }
}

__global__ void __launch_bounds__(256)
gn_from_stats_kernel(RefitBuffers rb, int blocks_used, float w3d, float wnl, int max_iters, ReplayOut* __restrict__ pose_io,
                     GnState* __restrict__ gs_out, double* __restrict__ cost_out, int32_t* __restrict__ evals_out) {
  __shared__ double fin_scratch[8 * 32];
  __shared__ double fin[kMomentCount];
  if (blocks_used > 0) {
    final_reduce_partials(rb.partials, blocks_used, fin_scratch, fin);
    if (threadIdx.x < kSuffAll) rb.suff[threadIdx.x] = fin[threadIdx.x];
  } else {
    if (threadIdx.x < kSuffAll) fin[threadIdx.x] = rb.suff[threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  GnState gs;
  gn_state_init(pose_io, &gs);
  for (int it = 0; it < max_iters && !gs.done; ++it) {
    double a[kGnAcc], Hn[21], gn[6], cost, rows;
    gn_eval_from_stats(fin, (double)w3d, (double)wnl, gs.Rp, gs.tp, a);
    gn_moments_to_normal_eq(a, Hn, gn, &cost, &rows);
    gn_lm_update(&gs, Hn, gn, cost, rows);
  }
  if (gs.evals > 0) gn_publish(&gs, pose_io, cost_out, evals_out);
  *gs_out = gs;
}
void launch_gn_from_stats(RefitBuffers rb, int blocks_used, float w3d, float wnl, int max_iters, ReplayOut* pose_io,
                          GnState* gs, double* cost_out, int32_t* evals_out, cudaStream_t s) {
  gn_from_stats_kernel<<<1, 256, 0, s>>>(rb, blocks_used, w3d, wnl, max_iters, pose_io, gs, cost_out, evals_out);
}

void launch_gn_iteration(const FrameView& f, const int16_t* mask, int mask_cols, float w2d, float w3d, float wnl,
                         RefitBuffers rb, GnState* gs, FrameStats* st, ReplayOut* pose_out, double* cost_out,
                         int32_t* evals_out, cudaStream_t s) {
  const int blocks = refit_grid(f.n, rb.num_sms);
  const bool generic = mask_cols >= 1 && f.bv != nullptr && w2d > 0.f;
  (void)generic;  // 3-D / normal rows alone never come here (launch_gn_from_stats)
  gn_iteration_kernel<true><<<blocks, 256, 0, s>>>(f, mask, mask_cols, w2d, w3d, wnl, rb, gs, st, pose_out, cost_out,
                                                   evals_out);
}

// ================================================================================================
// nl_shinji_kneip_ls — the reference's multi-modal refinement (AbsoluteOrientationNormal.hpp:447-552 with
// find_opt_cc :13-46), quirks included: the 3-3 / N-N weights are divided by 32767 (AOPoseAdapter.hpp:167,
// NormalAOPoseAdapter.hpp:159) and M23 / M33 / MNN / K / TW / M / TL are NOT reset between the three passes.
// ================================================================================================
constexpr int kNlskAcc = 40;
// prepass accumulator layout: [0] TV [1] N [2..4] sum v xw [5..7] sum v xc [8..16] sum v xc xw^T [17] sum v |xc|^2
// [18..26] sum l nc nw^T [27] tl [28] mnn [29] tw [30] k23 [31..36] AA (00,01,02,11,12,22) [37..39] bb
__global__ void __launch_bounds__(256)
nlsk_prepass_kernel(FrameView f, const int16_t* __restrict__ mask3, const float* __restrict__ w3, const ReplayOut* pose,
                    RefitBuffers rb, NlskState* ns, FrameStats* st) {
  __shared__ double red[8 * kNlskAcc];
  __shared__ bool is_last;
  const int n = f.n;
  double acc[kNlskAcc];
#pragma unroll
  for (int k = 0; k < kNlskAcc; ++k) acc[k] = 0.0;
  // R_wc = R_cw^-1 as a matrix, from the adapter's current rotation (find_opt_cc :21)
  double q[4] = {pose->q[0], pose->q[1], pose->q[2], pose->q[3]};
  {
    const double len = sqrt((q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3]));
    for (int k = 0; k < 4; ++k) q[k] /= len;  // inverse() goes through the normalising constructor
  }
  const double qi[4] = {-q[0], -q[1], -q[2], q[3]};
  double Rwc[9];
  quat_to_matrix_t<double>(qi, Rwc);
  const double inv_short = 1.0 / 32767.0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    const bool i23 = mask3[c] == 1, i33 = mask3[n + c] == 1, inn = mask3[2 * n + c] == 1;
    if (!(i23 || i33 || inn)) continue;
    const F3 xwf = load_col(f.xw, c);
    const double xw[3] = {xwf.x, xwf.y, xwf.z};
    if (i33) {
      const double v = w3 ? (double)w3[n + c] * inv_short : 1.0;
      const F3 xcf = load_col(f.xc, c);
      const double xc[3] = {xcf.x, xcf.y, xcf.z};
      acc[0] += v;
      acc[1] += 1.0;
      for (int r = 0; r < 3; ++r) {
        acc[2 + r] += v * xw[r];
        acc[5 + r] += v * xc[r];
        for (int k = 0; k < 3; ++k) acc[8 + 3 * r + k] += v * xc[r] * xw[k];
      }
      acc[17] += v * (xc[0] * xc[0] + xc[1] * xc[1] + xc[2] * xc[2]);
    }
    if (inn) {
      const double l = w3 ? (double)w3[2 * n + c] * inv_short : 1.0;
      const F3 ncf = load_col(f.nc, c), nwf = load_col(f.nw, c);
      const double nc[3] = {ncf.x, ncf.y, ncf.z}, nw[3] = {nwf.x, nwf.y, nwf.z};
      for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k) acc[18 + 3 * r + k] += l * nc[r] * nw[k];
      acc[27] += l;
      acc[28] += 1.0;
    }
    if (i23) {
      const double w = w3 ? (double)w3[c] : 1.0;
      acc[29] += w;
      acc[30] += 1.0;
      const F3 bf = load_col(f.bv, c);
      const double b[3] = {bf.x, bf.y, bf.z};
      double vr[3];
      for (int r = 0; r < 3; ++r) vr[r] = Rwc[3 * r] * b[0] + Rwc[3 * r + 1] * b[1] + Rwc[3 * r + 2] * b[2];
      const double A00 = 1 - vr[0] * vr[0], A01 = -vr[0] * vr[1], A02 = -vr[0] * vr[2], A11 = 1 - vr[1] * vr[1],
                   A12 = -vr[1] * vr[2], A22 = 1 - vr[2] * vr[2];
      acc[31] += A00;
      acc[32] += A01;
      acc[33] += A02;
      acc[34] += A11;
      acc[35] += A12;
      acc[36] += A22;
      acc[37] += A00 * xw[0] + A01 * xw[1] + A02 * xw[2];
      acc[38] += A01 * xw[0] + A11 * xw[1] + A12 * xw[2];
      acc[39] += A02 * xw[0] + A12 * xw[1] + A22 * xw[2];
    }
  }
  block_reduce_store<kNlskAcc>(acc, rb.partials, red);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int tk = atomicAdd(&st->ticket2, 1u);
    is_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  __shared__ double fin_scratch[8 * 32];
  __shared__ double fin[kMomentCount];
  final_reduce_partials(rb.partials, gridDim.x, fin_scratch, fin);
  if (threadIdx.x != 0) return;
  st->ticket2 = 0;
  const double* a = fin;
  ns->TV = a[0];
  ns->N = (int)a[1];
  for (int r = 0; r < 3; ++r) {
    ns->Cw[r] = a[2 + r];
    ns->Cc[r] = a[5 + r];
  }
  if (ns->N > 2)
    for (int r = 0; r < 3; ++r) {  // :466-469 (with N <= 2 the raw weighted sums are kept, as in the reference)
      ns->Cw[r] /= ns->TV;
      ns->Cc[r] /= ns->TV;
    }
  // centred sums from the raw moments: sum v (xc-Cc)(xw-Cw)^T = sum v xc xw^T - Cc (sum v xw)^T - (sum v xc) Cw^T + TV Cc Cw^T
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k)
      ns->S33[3 * r + k] = a[8 + 3 * r + k] - ns->Cc[r] * a[2 + k] - a[5 + r] * ns->Cw[k] + a[0] * ns->Cc[r] * ns->Cw[k];
  ns->sig = a[17] - 2.0 * (ns->Cc[0] * a[5] + ns->Cc[1] * a[6] + ns->Cc[2] * a[7]) +
            a[0] * (ns->Cc[0] * ns->Cc[0] + ns->Cc[1] * ns->Cc[1] + ns->Cc[2] * ns->Cc[2]);
  for (int k = 0; k < 9; ++k) ns->SNN[k] = a[18 + k];
  ns->tl = a[27];
  ns->mnn = (int)a[28];
  ns->tw = a[29];
  ns->k23 = (int)a[30];
  // find_opt_cc: solve AA x = bb through the SVD with Eigen's rank threshold, NaN when |det AA| < 1e-4 (:41-44)
  const double AA[9] = {a[31], a[32], a[33], a[32], a[34], a[35], a[33], a[35], a[36]};
  ns->cp_ok = 0;
  if (fabs(mat_det<double>(AA)) >= 0.0001) {
    double U[9], V[9], sv[3];
    svd3_jacobi<double>(AA, U, V, sv);
    const double thr0 = sv[0] * (3.0 * DBL_EPSILON);
    const double thr = thr0 > DBL_MIN ? thr0 : DBL_MIN;
    int rank = 0;
    while (rank < 3 && sv[rank] > thr) ++rank;
    double tmp[3] = {0, 0, 0};
    for (int k = 0; k < rank; ++k) tmp[k] = (U[k] * a[37] + U[3 + k] * a[38] + U[6 + k] * a[39]) / sv[k];
    for (int r = 0; r < 3; ++r) {
      double x = 0.0;
      for (int k = 0; k < rank; ++k) x += V[3 * r + k] * tmp[k];
      ns->cp[r] = x;
    }
    ns->cp_ok = 1;
  }
  for (int k = 0; k < 9; ++k) ns->M23[k] = ns->M33[k] = ns->MNN[k] = 0.0;
  ns->TW = ns->TL = 0.0;
  ns->K = ns->M = 0;
  ns->stopped = 0;
  ns->iter = 0;
  for (int k = 0; k < 4; ++k) {
    ns->q0[k] = q[k];
    ns->q_opt[k] = k == 3 ? 1.0 : 0.0;  // Sophus::SO3 default = identity (:480)
  }
  for (int k = 0; k < 3; ++k) ns->t0[k] = pose->t[k];
  // c_opt = R_cw^-1 * (-t_w)  (:479)
  const double nt[3] = {-ns->t0[0], -ns->t0[1], -ns->t0[2]};
  quat_rotate<double>(qi, nt, ns->c_opt);
}

__global__ void __launch_bounds__(256)
nlsk_iteration_kernel(FrameView f, const int16_t* __restrict__ mask3, const float* __restrict__ w3, RefitBuffers rb,
                      NlskState* ns, FrameStats* st, ReplayOut* pose_out) {
  if (ns->stopped) return;
  __shared__ double red[8 * 9];
  __shared__ bool is_last;
  const int n = f.n;
  double acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.0;
  const double co[3] = {ns->c_opt[0], ns->c_opt[1], ns->c_opt[2]};
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    if (mask3[c] != 1) continue;
    const double w = w3 ? (double)w3[c] : 1.0;
    const F3 xwf = load_col(f.xw, c), bf = load_col(f.bv, c);
    double Aw[3] = {xwf.x - co[0], xwf.y - co[1], xwf.z - co[2]};
    const double z = Aw[0] * Aw[0] + Aw[1] * Aw[1] + Aw[2] * Aw[2];
    if (z > 0.0) {
      const double nn = sqrt(z);
      Aw[0] /= nn;
      Aw[1] /= nn;
      Aw[2] /= nn;
    }
    const double b[3] = {bf.x, bf.y, bf.z};
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 3; ++k) acc[3 * r + k] += (w * b[r]) * Aw[k];
  }
  block_reduce_store<9>(acc, rb.partials, red);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int tk = atomicAdd(&st->ticket2, 1u);
    is_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  __shared__ double fin_scratch[8 * 32];
  __shared__ double fin[kMomentCount];
  final_reduce_partials(rb.partials, gridDim.x, fin_scratch, fin);
  if (threadIdx.x != 0) return;
  st->ticket2 = 0;
  // ---- one pass of the reference's loop body (:483-543), accumulators carried over as in the reference ----
  for (int k = 0; k < 9; ++k) {
    ns->M23[k] += fin[k];
    ns->M33[k] += ns->S33[k];
    ns->MNN[k] += ns->SNN[k];
  }
  ns->TW += ns->tw;
  ns->K += ns->k23;
  ns->TL += ns->tl;
  ns->M += ns->mnn;
  double sigma = ns->sig;
  if (ns->N > 2) {
    for (int k = 0; k < 9; ++k) ns->M33[k] /= ns->TV;
    sigma /= ns->TV;
  } else {
    for (int k = 0; k < 9; ++k) ns->M33[k] = 0.0;
    sigma = 1.0;
  }
  if (ns->M > 0) {
    for (int k = 0; k < 9; ++k) ns->MNN[k] /= ns->TL;
  } else {
    for (int k = 0; k < 9; ++k) ns->MNN[k] = 0.0;
  }
  if (ns->K > 0) {
    for (int k = 0; k < 9; ++k) ns->M23[k] /= ns->TW;
  } else {
    for (int k = 0; k < 9; ++k) ns->M23[k] = 0.0;
  }
  for (int k = 0; k < 9; ++k) ns->M33[k] += sigma * (ns->M23[k] + ns->MNN[k]);
  double qo[4];
  rotation_from_covariance<double>(ns->M33, qo);
  for (int k = 0; k < 4; ++k) ns->q_opt[k] = qo[k];
  // c = Cw - R_opt^-1 * Cc
  double qn[4];
  {
    const double len = sqrt((qo[0] * qo[0] + qo[1] * qo[1]) + (qo[2] * qo[2] + qo[3] * qo[3]));
    qn[0] = -qo[0] / len;
    qn[1] = -qo[1] / len;
    qn[2] = -qo[2] / len;
    qn[3] = qo[3] / len;
  }
  double rc[3];
  quat_rotate<double>(qn, ns->Cc, rc);
  const double cvec[3] = {ns->Cw[0] - rc[0], ns->Cw[1] - rc[1], ns->Cw[2] - rc[2]};
  bool brk = false;
  if (ns->N > 2) {
    if (ns->cp_ok) {
      const double fk = (double)ns->K / (double)(ns->K + ns->N), fn = (double)ns->N / (double)(ns->K + ns->N);
      for (int r = 0; r < 3; ++r) ns->c_opt[r] = fk * ns->cp[r] + fn * cvec[r];
    } else {
      for (int r = 0; r < 3; ++r) ns->c_opt[r] = cvec[r];
    }
  } else {
    if (ns->cp_ok) {
      for (int r = 0; r < 3; ++r) ns->c_opt[r] = ns->cp[r];
    } else {
      brk = true;
    }
  }
  ns->iter += 1;
  if (brk || ns->iter >= 3) {
    ns->stopped = 1;
    // setRcw(R_opt); sett(R_opt * (-c_opt))  (:545-546)
    const double nc[3] = {-ns->c_opt[0], -ns->c_opt[1], -ns->c_opt[2]};
    double tt[3];
    quat_rotate<double>(qo, nc, tt);
    for (int k = 0; k < 4; ++k) pose_out->q[k] = (float)qo[k];
    for (int k = 0; k < 3; ++k) pose_out->t[k] = (float)tt[k];
    pose_out->refit_ok = 1;
  }
}

void launch_nlsk_prepass(const FrameView& f, const int16_t* mask3, const float* weights3, const ReplayOut* pose,
                         RefitBuffers rb, NlskState* ns, FrameStats* st, cudaStream_t s) {
  nlsk_prepass_kernel<<<refit_grid(f.n, rb.num_sms), 256, 0, s>>>(f, mask3, weights3, pose, rb, ns, st);
}
void launch_nlsk_iteration(const FrameView& f, const int16_t* mask3, const float* weights3, RefitBuffers rb, NlskState* ns,
                           FrameStats* st, ReplayOut* pose_out, cudaStream_t s) {
  nlsk_iteration_kernel<<<refit_grid(f.n, rb.num_sms), 256, 0, s>>>(f, mask3, weights3, rb, ns, st, pose_out);
}

}  // namespace rpe
