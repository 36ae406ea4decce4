// f64.cu — the binary64 RANSAC path for adapters instantiated with Tp = double.
//
// The reference's TestMain.cpp runs its estimators as <double> (e.g. nl_shinji_kneip_ransac<double>, TestMain.cpp:210).
// A double adapter must see the inlier masks of that double CPU path, which a binary32 scorer cannot guarantee, so the
// whole decision chain is evaluated in binary64 here, in the reference's operation order:
//   generation   the same solver templates as the float path (include/rpe/solvers*.h) instantiated for double
//   scoring      binary32 tiled prefilter (the float path's kernels on float copies of the arrays, same guard bands —
//                they budget the extra input roundings) + binary64 exact-order evaluation of every
//                borderline evaluation: quaternion sandwich / matrix form, unfused, IEEE sqrt and division
//   replay       strict `votes > max`, Iter = RANSACUpdateNumIters<double>(..) (include/rpe/ransac_rule.h)
//   mask         the winner's flags, same exact tests
// Refits keep using the float copies of the arrays with binary64 accumulation (pose tolerance 1e-6).
// Compiled with -fmad=false like every TU of this library; the __d*_rn intrinsics pin one rounding per operation.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/rpe/ransac_rule.h"
#include "kernels.cuh"
#include "../../include/rpe/solvers.h"
#include "../../include/rpe/solvers_p3p.h"

namespace rpe {

// ---- exact-order primitives, binary64 (twins of the ex_* functions of rpe_device.cuh) ---------------------------
struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) {
  D3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double dx_sum3(double a, double b, double c) { return __dadd_rn(a, __dadd_rn(b, c)); }
__device__ __forceinline__ D3 dx_cross(D3 a, D3 b) {
  return d3(dsub(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)), dsub(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
            dsub(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__device__ __forceinline__ double dx_dot(D3 a, D3 b) {
  return dx_sum3(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y), __dmul_rn(a.z, b.z));
}
__device__ __forceinline__ double dx_norm(D3 a) { return __dsqrt_rn(dx_dot(a, a)); }
__device__ __forceinline__ D3 dx_quat_rotate(const double q[4], D3 v) {
  const D3 qv = d3(q[0], q[1], q[2]);
  D3 uv = dx_cross(qv, v);
  uv = d3(__dadd_rn(uv.x, uv.x), __dadd_rn(uv.y, uv.y), __dadd_rn(uv.z, uv.z));
  const D3 c2 = dx_cross(qv, uv);
  const double w = q[3];
  return d3(__dadd_rn(__dadd_rn(v.x, __dmul_rn(w, uv.x)), c2.x), __dadd_rn(__dadd_rn(v.y, __dmul_rn(w, uv.y)), c2.y),
            __dadd_rn(__dadd_rn(v.z, __dmul_rn(w, uv.z)), c2.z));
}
__device__ __forceinline__ void dx_quat_to_matrix(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = __dmul_rn(2.0, x), ty = __dmul_rn(2.0, y), tz = __dmul_rn(2.0, z);
  const double twx = __dmul_rn(tx, w), twy = __dmul_rn(ty, w), twz = __dmul_rn(tz, w);
  const double txx = __dmul_rn(tx, x), txy = __dmul_rn(ty, x), txz = __dmul_rn(tz, x);
  const double tyy = __dmul_rn(ty, y), tyz = __dmul_rn(tz, y), tzz = __dmul_rn(tz, z);
  R[0] = dsub(1.0, __dadd_rn(tyy, tzz));
  R[1] = dsub(txy, twz);
  R[2] = __dadd_rn(txz, twy);
  R[3] = __dadd_rn(txy, twz);
  R[4] = dsub(1.0, __dadd_rn(txx, tzz));
  R[5] = dsub(tyz, twx);
  R[6] = dsub(txz, twy);
  R[7] = __dadd_rn(tyz, twx);
  R[8] = dsub(1.0, __dadd_rn(txx, tyy));
}
__device__ __forceinline__ D3 dx_mat_vec(const double R[9], D3 v) {
  return d3(dx_sum3(__dmul_rn(R[0], v.x), __dmul_rn(R[1], v.y), __dmul_rn(R[2], v.z)),
            dx_sum3(__dmul_rn(R[3], v.x), __dmul_rn(R[4], v.y), __dmul_rn(R[5], v.z)),
            dx_sum3(__dmul_rn(R[6], v.x), __dmul_rn(R[7], v.y), __dmul_rn(R[8], v.z)));
}
__device__ __forceinline__ bool dx_is_valid(D3 p) { return p.x == p.x || p.y == p.y || p.z == p.z; }
__device__ __forceinline__ D3 load_col64(const double* __restrict__ a, int i) { return d3(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }

// AbsoluteOrientation.hpp:137-138
__device__ __forceinline__ bool dx_test_3d(const double q[4], const double t[3], D3 xw, D3 xc, double thr3d) {
  const D3 r = dx_quat_rotate(q, xw);
  const D3 y = d3(__dadd_rn(r.x, t[0]), __dadd_rn(r.y, t[1]), __dadd_rn(r.z, t[2]));
  const D3 e = d3(dsub(xc.x, y.x), dsub(xc.y, y.y), dsub(xc.z, y.z));
  return dx_norm(e) < thr3d;
}
// P3P.hpp:365-372 (matrix form, Rm != null) / :442-449 (quaternion form)
__device__ __forceinline__ bool dx_test_2d(const double q[4], const double t[3], const double* Rm, D3 xw, D3 bv,
                                           double cos_thr) {
  const D3 r = Rm ? dx_mat_vec(Rm, xw) : dx_quat_rotate(q, xw);
  D3 pc = d3(__dadd_rn(r.x, t[0]), __dadd_rn(r.y, t[1]), __dadd_rn(r.z, t[2]));
  const double nrm = dx_norm(pc);
  pc = d3(__ddiv_rn(pc.x, nrm), __ddiv_rn(pc.y, nrm), __ddiv_rn(pc.z, nrm));
  return dx_dot(pc, bv) > cos_thr;
}
// AbsoluteOrientationNormal.hpp:248-249
__device__ __forceinline__ bool dx_test_nl(const double q[4], D3 nw, D3 nc, double cos_nl) {
  return dx_dot(nc, dx_quat_rotate(q, nw)) > cos_nl;
}

// ================================================================================================
// binary64 -> binary32 copies for the refit kernels
// ================================================================================================
__global__ void f64_to_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, size_t count) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}
void launch_f64_to_f32(const double* src, float* dst, size_t count, cudaStream_t s) {
  if (!count) return;
  const int blocks = (int)((count + 255) / 256 < 1184 ? (count + 255) / 256 : 1184);
  f64_to_f32_kernel<<<blocks, 256, 0, s>>>(src, dst, count);
}

// ================================================================================================
// generation (twin of hypgen_kernel in pipeline.cu)
// ================================================================================================
__global__ void __launch_bounds__(128)
hypgen64_kernel(int method, FrameView64 f, const int32_t* __restrict__ samples, int H, HypGen64* __restrict__ gen,
                int32_t* __restrict__ votes,
                const int32_t* __restrict__ stale_eff) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= H) return;
  const int S = method_slots(method);
  const int s = blockIdx.y;
  const int solver = method_slot_solver(method, s);
  const int32_t* sel = samples + 4 * ii;
  HypGen64 g;
  g.q[0] = g.q[1] = g.q[2] = 0.0;
  g.q[3] = 1.0;
  g.t[0] = g.t[1] = g.t[2] = 0.0;
  g.valid = 0;
  g.pad = 0;
  bool in_range = true;  // see hypgen_kernel
  for (int k = 0; k < method_sample_size(method); ++k) in_range = in_range && sel[k] >= 0 && sel[k] < f.n;
  if (!in_range) {
    gen[ii * S + s] = g;
    votes[ii * S + s] = -1;
    return;
  }
  if (solver == SOLVER_AO) {
    double Xw[9], Xc[9];
    bool all_valid = true;
    for (int k = 0; k < 3; ++k) {
      const int c = sel[k];
      const D3 pw = load_col64(f.xw, c), pc = load_col64(f.xc, c);
      Xw[3 * k] = pw.x;
      Xw[3 * k + 1] = pw.y;
      Xw[3 * k + 2] = pw.z;
      Xc[3 * k] = pc.x;
      Xc[3 * k + 1] = pc.y;
      Xc[3 * k + 2] = pc.z;
      all_valid = all_valid && dx_is_valid(pc);
    }
    if (all_valid) g.valid = shinji3<double>(Xw, Xc, method == RPE_SHINJI ? 3 : 4, g.q, g.t) ? 1 : 0;
  } else if (solver == SOLVER_P3P) {
    double Xw[12], bv[12];
    for (int k = 0; k < 4; ++k) {
      const int c = sel[k];
      const D3 pw = load_col64(f.xw, c), b = load_col64(f.bv, c);
      Xw[3 * k] = pw.x;
      Xw[3 * k + 1] = pw.y;
      Xw[3 * k + 2] = pw.z;
      bv[3 * k] = b.x;
      bv[3 * k + 1] = b.y;
      bv[3 * k + 2] = b.z;
    }
    // P3P.hpp:338,415 start from 1000000.0; the matrix overload (P3P.hpp:258) from numeric_limits<Tp>::max()
    const double start = (method == RPE_KNEIP || method == RPE_KNEIP_QUAT) ? 1000000.0 : 1.7976931348623157e308;
    g.valid = kneip_select<double>(Xw, bv, start, g.q, g.t) ? 1 : 0;
  } else {
    double pc0[3], nc0[3], pc1[3], pw0[3], nw0[3], pw1[3];
    const int c0 = sel[0], c1 = sel[1];
    const int e0 = stale_eff ? stale_eff[2 * ii] : c0, e1 = stale_eff ? stale_eff[2 * ii + 1] : c1;  // see hypgen_kernel
    for (int r = 0; r < 3; ++r) {
      pc0[r] = e0 >= 0 ? f.xc[3 * e0 + r] : 0.0;
      nc0[r] = e0 >= 0 ? f.nc[3 * e0 + r] : 0.0;
      pc1[r] = e1 >= 0 ? f.xc[3 * e1 + r] : 0.0;
      pw0[r] = f.xw[3 * c0 + r];
      nw0[r] = f.nw[3 * c0 + r];
      pw1[r] = f.xw[3 * c1 + r];
    }
    nl_2p<double>(pc0, nc0, pc1, pw0, nw0, pw1, g.q, g.t);
    g.valid = 1;
  }
  gen[ii * S + s] = g;
  votes[ii * S + s] = g.valid ? 0 : -1;
}
void launch_hypgen64(int method, const FrameView64& f, const int32_t* samples_dev, int H, HypGen64* gen, int32_t* votes,
                     cudaStream_t s, const int32_t* stale_eff) {
  if (H <= 0) return;
  dim3 grid((H + 127) / 128, method_slots(method));
  hypgen64_kernel<<<grid, 128, 0, s>>>(method, f, samples_dev, H, gen, votes, stale_eff);
}

// ================================================================================================
// operands of the binary32 prefilter, derived from the binary64 hypotheses
// ================================================================================================
__global__ void derive_fast64_kernel(const HypGen64* __restrict__ g64, HypGen* __restrict__ gen, HypFast* __restrict__ fast,
                                     int n_slots) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  const HypGen64 g = g64[i];
  HypGen gf;
  for (int k = 0; k < 4; ++k) gf.q[k] = (float)g.q[k];
  for (int k = 0; k < 3; ++k) gf.t[k] = (float)g.t[k];
  gf.valid = g.valid;
  gen[i] = gf;
  const double x = g.q[0], y = g.q[1], z = g.q[2], w = g.q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  HypFast f;
  f.nR[0] = -(float)(1.0 - (tyy + tzz));
  f.nR[1] = -(float)(txy - twz);
  f.nR[2] = -(float)(txz + twy);
  f.nR[3] = -(float)(txy + twz);
  f.nR[4] = -(float)(1.0 - (txx + tzz));
  f.nR[5] = -(float)(tyz - twx);
  f.nR[6] = -(float)(txz - twy);
  f.nR[7] = -(float)(tyz + twx);
  f.nR[8] = -(float)(1.0 - (txx + tyy));
  for (int k = 0; k < 3; ++k) f.nt[k] = -(float)g.t[k];
  fast[i] = f;
}
void launch_derive_fast64(const HypGen64* g64, HypGen* gen, HypFast* fast, int n_slots, cudaStream_t s) {
  if (n_slots <= 0) return;
  derive_fast64_kernel<<<(n_slots + 127) / 128, 128, 0, s>>>(g64, gen, fast, n_slots);
}

// binary64 re-evaluation of the prefilter's borderline evaluations (twin of fixup_kernel in score.cu)
__global__ void __launch_bounds__(256)
fixup64_kernel(int method, FrameView64 f, const HypGen64* __restrict__ gen, Thresh64 th, int32_t* __restrict__ votes,
               const FrameStats* __restrict__ st, Worklist wl, int nseg) {
  if (st->wl_overflow) return;  // the whole frame is rescored in binary64 instead
  const unsigned int cap = wl.capacity / (unsigned int)nseg;
  const unsigned int count = min(wl.counts[blockIdx.x], cap);
  const uint2* seg = wl.entries + (size_t)blockIdx.x * cap;
  for (unsigned int i = blockIdx.y * blockDim.x + threadIdx.x; i < count; i += gridDim.y * blockDim.x) {
    const uint2 e = seg[i];
    const int slot = (int)e.x;
    double q[4], t[3];
    for (int k = 0; k < 4; ++k) q[k] = gen[slot].q[k];
    for (int k = 0; k < 3; ++k) t[k] = gen[slot].t[k];
    auto one = [&](int modality, int c) -> bool {
      if (modality == 1) {
        const D3 xc = load_col64(f.xc, c);
        return dx_is_valid(xc) && dx_test_3d(q, t, load_col64(f.xw, c), xc, th.thr3d);
      }
      if (modality == 2) return dx_is_valid(load_col64(f.xc, c)) && dx_test_nl(q, load_col64(f.nw, c), load_col64(f.nc, c), th.cos_nl);
      double Rm[9];
      if (method == RPE_KNEIP) dx_quat_to_matrix(q, Rm);
      return dx_test_2d(q, t, method == RPE_KNEIP ? Rm : nullptr, load_col64(f.xw, c), load_col64(f.bv, c), th.cos_thr);
    };
    int add = 0;
    if ((e.y >> 30) == 3u) {  // unit entry of the multi-modality scorer: pair index, six borderline bits (see fixup_kernel)
      const int c0 = 2 * (int)(e.y & 0xffffffu);
      unsigned int bits = (e.y >> 24) & 0x3fu;
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1u;
        add += one(b >> 1, c0 + (b & 1)) ? 1 : 0;
      }
    } else {
      add = one((int)(e.y >> 30), (int)(e.y & 0x3fffffffu)) ? 1 : 0;
    }
    if (add) atomicAdd(&votes[slot], add);
  }
}
void launch_fixup64(int method, const FrameView64& f, const HypGen64* gen, Thresh64 th, int32_t* votes, FrameStats* st,
                    Worklist wl, int nseg, cudaStream_t s) {
  if (nseg <= 0) return;
  fixup64_kernel<<<dim3(nseg, nseg > 1024 ? 1 : 4), 256, 0, s>>>(method, f, gen, th, votes, st, wl, nseg);
}
__global__ void zero_votes64_kernel(const HypGen64* __restrict__ gen, int n_slots, int32_t* __restrict__ votes,
                                    const FrameStats* __restrict__ st) {
  if (!st->wl_overflow) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x)
    votes[i] = gen[i].valid ? 0 : -1;
}

// ================================================================================================
// scoring: thread <-> slot, CTA column <-> correspondence slice (every lane reads the same correspondence)
// ================================================================================================
__global__ void __launch_bounds__(128)
score64_kernel(int method, FrameView64 f, const HypGen64* __restrict__ gen, int n_slots, Thresh64 th,
               int32_t* __restrict__ votes, int corr_per_cta, const FrameStats* __restrict__ st /*null: unconditional*/) {
  if (st && !st->wl_overflow) return;
  const int slot = blockIdx.y * blockDim.x + threadIdx.x;
  const bool live = slot < n_slots && gen[slot].valid != 0;
  double q[4], t[3];
  for (int k = 0; k < 4; ++k) q[k] = live ? gen[slot].q[k] : CUDART_NAN;
  for (int k = 0; k < 3; ++k) t[k] = live ? gen[slot].t[k] : CUDART_NAN;
  double Rm[9];
  dx_quat_to_matrix(q, Rm);
  const bool u2 = method_uses_2d(method), u3 = method_uses_3d(method), un = method_uses_nl(method);
  const double* Rp = method == RPE_KNEIP ? Rm : nullptr;
  const int c0 = blockIdx.x * corr_per_cta;
  const int c1 = min(c0 + corr_per_cta, f.n);
  int cnt = 0;
  for (int c = c0; c < c1; ++c) {
    const D3 xw = load_col64(f.xw, c);
    bool valid = false;
    D3 xc = d3(0.0, 0.0, 0.0);
    if (u3 || un) {
      xc = load_col64(f.xc, c);
      valid = dx_is_valid(xc);
    }
    if (un && valid) cnt += dx_test_nl(q, load_col64(f.nw, c), load_col64(f.nc, c), th.cos_nl) ? 1 : 0;
    if (u3 && valid) cnt += dx_test_3d(q, t, xw, xc, th.thr3d) ? 1 : 0;
    if (u2) cnt += dx_test_2d(q, t, Rp, xw, load_col64(f.bv, c), th.cos_thr) ? 1 : 0;
  }
  if (live && cnt) atomicAdd(&votes[slot], cnt);
}
// only_if_overflow != null: the fallback after a prefilter whose worklist overflowed (votes are reset first)
void launch_score64(int method, const FrameView64& f, const HypGen64* gen, int n_slots, Thresh64 th, int32_t* votes,
                    int num_sms, const FrameStats* only_if_overflow, cudaStream_t s) {
  if (n_slots <= 0 || f.n <= 0) return;
  if (only_if_overflow) zero_votes64_kernel<<<4, 256, 0, s>>>(gen, n_slots, votes, only_if_overflow);
  const int threads = 128;
  const int gy = (n_slots + threads - 1) / threads;
  int gx = (8 * num_sms) / gy;
  if (gx < 1) gx = 1;
  if (gx > f.n) gx = f.n;
  const int corr_per_cta = (f.n + gx - 1) / gx;
  gx = (f.n + corr_per_cta - 1) / corr_per_cta;
  score64_kernel<<<dim3(gx, gy), threads, 0, s>>>(method, f, gen, n_slots, th, votes, corr_per_cta, only_if_overflow);
}

// ================================================================================================
// replay of the sequential rule (twin of replay_kernel in pipeline.cu, binary64 rule)
// ================================================================================================
__global__ void replay64_begin_kernel(ReplayState64* rs, int iter_max) {
  if (threadIdx.x != 0) return;
  rs->best = -1;
  rs->iter = iter_max;
  rs->win = -1;
  rs->cur_iter = -1;
  rs->stop = 0;
  rs->slots_done = 0;
  rs->borderline = 0;
  rs->overflow = 0;
  for (int k = 0; k < 4; ++k) rs->q[k] = k == 3 ? 1.0 : 0.0;
  for (int k = 0; k < 3; ++k) rs->t[k] = 0.0;
}
void launch_replay64_begin(ReplayState64* rs, int iter_max, cudaStream_t s) { replay64_begin_kernel<<<1, 32, 0, s>>>(rs, iter_max); }

constexpr int kReplay64Chunk = 4096;
__global__ void __launch_bounds__(256)
replay64_kernel(int method, const HypGen64* __restrict__ gen, const int32_t* __restrict__ votes, int H, int iter_base, int n,
                double confidence, FrameStats* __restrict__ st, ReplayState64* __restrict__ rs, ReplayOut* __restrict__ out,
                Pose64* __restrict__ out64, int finalize) {
  __shared__ int32_t sv[kReplay64Chunk];
  __shared__ int s_state[5];  // best, Iter, win, cur_iter, stop
  const int lane = threadIdx.x & 31;
  const int S = method_slots(method);
  const int K = method_model_points(method);
  const int mod = method_modalities(method);
  const int E = H * S;
  const int slot_base = iter_base * S;
  if (threadIdx.x == 0) {
    s_state[0] = rs->best;
    s_state[1] = rs->iter;
    s_state[2] = rs->win;
    s_state[3] = rs->cur_iter;
    s_state[4] = rs->stop;
  }
  __syncthreads();
  for (int cbase = 0; cbase < E; cbase += kReplay64Chunk) {
    if (s_state[4]) break;
    const int cn = min(kReplay64Chunk, E - cbase);
    for (int i = threadIdx.x; i < cn; i += blockDim.x) sv[i] = votes[cbase + i];
    __syncthreads();
    if (threadIdx.x < 32) {
      const double log_num = rule_log_numerator_d(confidence);
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
        RuleDenominatorD rd;
        rd.state = 0;
        rd.log_denom = 0.0;
        if (rec) rd = rule_denominator_d(outlier_ratio_d(mod, n, v), K);
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const int idx = slot_base + cbase + base + b;
          const int it = idx / S;
          const int vb = __shfl_sync(0xffffffffu, v, b);
          if (it != cur_iter) {
            if (it >= Iter) {
              stop = true;
              break;
            }
            cur_iter = it;
          }
          best = vb;
          win = idx;
          RuleDenominatorD rb_;
          rb_.state = __shfl_sync(0xffffffffu, rd.state, b);
          rb_.log_denom = __shfl_sync(0xffffffffu, rd.log_denom, b);
          Iter = rule_finish_d(log_num, rb_, Iter);
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
    if (win >= slot_base && win != rs->win) {
      const HypGen64 g = gen[win - slot_base];
      for (int k = 0; k < 4; ++k) rs->q[k] = g.q[k];
      for (int k = 0; k < 3; ++k) rs->t[k] = g.t[k];
    }
    rs->best = s_state[0];
    rs->iter = s_state[1];
    rs->win = win;
    rs->cur_iter = s_state[3];
    rs->stop = s_state[4] || ((long long)(slot_base + E) >= (long long)s_state[1] * S) ? 1 : 0;
    rs->slots_done += E;
    rs->borderline += (int)st->wl_count;
    rs->overflow |= st->wl_overflow ? 1 : 0;
    st->t_max_bits = 0;  // per-pass counters are consumed (as in replay_kernel)
    st->wl_count = 0;
    st->wl_consumed = 0;
    st->wl_overflow = 0;
    st->ticket = 0;
    st->ticket2 = 0;
    if (finalize) {
      ReplayOut o;
      for (int k = 0; k < 4; ++k) o.q[k] = (float)rs->q[k];
      for (int k = 0; k < 3; ++k) o.t[k] = (float)rs->t[k];
      o.max_votes = rs->best;
      o.iter_final = rs->iter;
      o.winner = rs->win;
      o.n_slots = rs->slots_done;
      o.n_borderline = rs->borderline;
      o.flags = 2 | (rs->overflow ? 1 : 0);  // bit 1: binary64 path
      o.n_inliers[0] = o.n_inliers[1] = o.n_inliers[2] = 0;
      o.refit_ok = 0;
      *out = o;
      Pose64 p;
      for (int k = 0; k < 4; ++k) p.q[k] = rs->q[k];
      for (int k = 0; k < 3; ++k) p.t[k] = rs->t[k];
      *out64 = p;
    }
  }
}
void launch_replay64(int method, const HypGen64* gen, const int32_t* votes, int H, int iter_base, int n, double confidence,
                     FrameStats* st, ReplayState64* rs, ReplayOut* out, Pose64* out64, bool finalize, cudaStream_t s) {
  replay64_kernel<<<1, 256, 0, s>>>(method, gen, votes, H, iter_base, n, confidence, st, rs, out, out64, finalize ? 1 : 0);
}

// ================================================================================================
// the winner's mask (twin of mask_kernel's tests; the refit statistics are gathered by the float kernels later)
// ================================================================================================
__global__ void __launch_bounds__(256)
mask64_kernel(int method, FrameView64 f, ReplayOut* pose_rw, const Pose64* __restrict__ pose64, Thresh64 th,
              int16_t* __restrict__ mask) {
  __shared__ int cnts[3];
  const int n = f.n;
  const int cols = method_mask_cols(method);
  if (threadIdx.x < 3) cnts[threadIdx.x] = 0;
  __syncthreads();
  double q[4], t[3];
  for (int k = 0; k < 4; ++k) q[k] = pose64->q[k];
  for (int k = 0; k < 3; ++k) t[k] = pose64->t[k];
  const bool have = pose_rw->winner >= 0;
  double Rm[9];
  dx_quat_to_matrix(q, Rm);
  const bool u2 = method_uses_2d(method), u3 = method_uses_3d(method), un = method_uses_nl(method);
  int c2 = 0, c3 = 0, cn = 0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    bool f2 = false, f3d = false, fn = false;
    if (have) {
      const D3 xw = load_col64(f.xw, c);
      bool valid = false;
      D3 xc = d3(0.0, 0.0, 0.0);
      if (u3 || un) {
        xc = load_col64(f.xc, c);
        valid = dx_is_valid(xc);
      }
      if (un && valid) fn = dx_test_nl(q, load_col64(f.nw, c), load_col64(f.nc, c), th.cos_nl);
      if (u3 && valid) f3d = dx_test_3d(q, t, xw, xc, th.thr3d);
      if (u2) f2 = dx_test_2d(q, t, method == RPE_KNEIP ? Rm : nullptr, xw, load_col64(f.bv, c), th.cos_thr);
      c2 += f2 ? 1 : 0;
      c3 += f3d ? 1 : 0;
      cn += fn ? 1 : 0;
    } else {
      f2 = f3d = fn = true;  // adapters start with setOnes() and setInlier is never called
    }
    // column 0 of a family without the 2-D test is never set by the reference's loop: 0 once a hypothesis has been
    // accepted (the per-iteration matrix starts from zero), still the initial 1 when nothing was accepted
    mask[c] = (int16_t)(!have ? 1 : ((cols == 1 || u2) ? (f2 ? 1 : 0) : 0));
    if (cols >= 2) mask[n + c] = (int16_t)(u3 ? (f3d ? 1 : 0) : (have ? 0 : 1));
    if (cols >= 3) mask[2 * n + c] = (int16_t)(fn ? 1 : 0);
  }
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
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < 3; ++k)
      if (cnts[k]) atomicAdd(&pose_rw->n_inliers[k], cnts[k]);
}
void launch_mask64(int method, const FrameView64& f, ReplayOut* pose_rw, const Pose64* pose64, Thresh64 th, int16_t* mask,
                   int num_sms, cudaStream_t s) {
  const int full = (f.n + 255) / 256;
  const int cap = 4 * (num_sms > 0 ? num_sms : 148);
  mask64_kernel<<<full < cap ? full : cap, 256, 0, s>>>(method, f, pose_rw, pose64, th, mask);
}

}  // namespace rpe
