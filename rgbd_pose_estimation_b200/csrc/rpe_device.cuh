// rpe_device.cuh — device-side types and the EXACT-ORDER arithmetic of the reference.
//
// Two arithmetic regimes live in this library:
//   * exact order  : the operation sequence of the reference's Eigen/Sophus CPU path, one IEEE
//                    rounding per operation, spelled with __f*_rn intrinsics so that no compiler
//                    flag can fuse or reorder it. Used by the generators' final stages, the
//                    borderline fix-up, the winner's mask and the replay. Bit-identical to the
//                    CPU path by construction.
//   * fast order   : FFMA / packed FFMA2 matrix form in the tiled scorer (score.cu). Its decisions
//                    are only trusted outside a rigorously sized guard band; inside
//                    the band the evaluation is redone in exact order.
//
// Reference arithmetic restated here (paths into /root/reference):
//   quaternion sandwich   sophus/so3.hpp:238-240 -> Eigen QuaternionBase::_transformVector
//   toRotationMatrix      sophus/so3.hpp:204-206 -> Eigen QuaternionBase::toRotationMatrix
//   3-D test              pose/AbsoluteOrientation.hpp:137-138
//   2-D test              pose/P3P.hpp:365-372 (matrix form), :442-449 (quaternion form)
//   normal test           pose/AbsoluteOrientationNormal.hpp:248-249
//   isValid               pose/AOPoseAdapter.hpp:147-152
// Fixed-size Eigen reductions (dot, squaredNorm, 3x3*3x1 coefficient) follow the unrolled redux
// tree a + (b + c).
#ifndef RPE_DEVICE_CUH_
#define RPE_DEVICE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rpe/det_math.h"
#include "../../include/rpe_c_api.h"

namespace rpe {

// ---- method traits (host + device) ---------------------------------------------------------
__host__ __device__ inline int method_slots(int m) {
  return (m == RPE_SHINJI_KNEIP || m == RPE_NL_SHINJI) ? 2 : (m == RPE_NL_SHINJI_KNEIP ? 3 : 1);
}
__host__ __device__ inline int method_mask_cols(int m) {
  return (m == RPE_KNEIP || m == RPE_KNEIP_QUAT) ? 1 : ((m == RPE_SHINJI || m == RPE_SHINJI_KNEIP) ? 2 : 3);
}
__host__ __device__ inline int method_sample_size(int m) { return m == RPE_SHINJI ? 3 : 4; }
__host__ __device__ inline int method_model_points(int m) { return (m == RPE_KNEIP || m == RPE_KNEIP_QUAT) ? 4 : 3; }
__host__ __device__ inline int method_modalities(int m) {
  return (m == RPE_SHINJI || m == RPE_KNEIP || m == RPE_KNEIP_QUAT) ? 1 : (m == RPE_NL_SHINJI_KNEIP ? 3 : 2);
}
__host__ __device__ inline bool method_uses_2d(int m) { return m != RPE_SHINJI && m != RPE_NL_SHINJI; }
__host__ __device__ inline bool method_uses_3d(int m) {
  return m == RPE_SHINJI || m == RPE_SHINJI_KNEIP || m == RPE_NL_SHINJI || m == RPE_NL_SHINJI_KNEIP;
}
__host__ __device__ inline bool method_uses_nl(int m) {
  return m == RPE_NL_KNEIP || m == RPE_NL_SHINJI || m == RPE_NL_SHINJI_KNEIP;
}
// slot kinds in v_solutions push order
enum { SOLVER_AO = 0, SOLVER_P3P = 1, SOLVER_NL2P = 2 };
__host__ __device__ inline int method_slot_solver(int m, int slot) {
  switch (m) {
    case RPE_SHINJI: return SOLVER_AO;
    case RPE_KNEIP:
    case RPE_KNEIP_QUAT:
    case RPE_NL_KNEIP: return SOLVER_P3P;
    case RPE_SHINJI_KNEIP: return slot == 0 ? SOLVER_AO : SOLVER_P3P;
    case RPE_NL_SHINJI: return slot == 0 ? SOLVER_AO : SOLVER_NL2P;
    default: return slot == 0 ? SOLVER_AO : (slot == 1 ? SOLVER_P3P : SOLVER_NL2P);
  }
}

// ---- device-resident records -----------------------------------------------------------------
// Generator output: exactly what the CPU path holds in a Sophus::SE3 (unit_quaternion + translation).
struct __align__(16) HypGen {
  float q[4];  // x, y, z, w
  float t[3];
  int32_t valid;  // 1 scored, 0 empty slot (invalid sample / no P3P root / SOPHUS_ENSURE would abort)
};
// Derived operands of the fast scorer: -R (rounded once from a binary64 evaluation of the
// quaternion polynomial) and -t.
struct __align__(16) HypFast {
  float nR[9];  // row-major
  float nt[3];
};

struct FrameStats {
  unsigned int m_corr_bits;  // max_i (|x_w,i| + |x_c,i|) over finite points, float bits
  unsigned int m_bv_bits;    // max_i max(|n_w,i|, |n_c,i|) over finite normals, float bits (normal-test guard band)
  unsigned int t_max_bits;   // max_h |t_h| over valid hypotheses, float bits
  unsigned int wl_count;     // borderline worklist length
  unsigned int wl_overflow;  // 1 if the worklist overflowed -> exact rescoring of the frame
  unsigned int ticket;       // last-block-done counter (refit kernels)
  unsigned int ticket2;
  unsigned int wl_consumed;  // borderline evaluations already resolved by earlier rpe_score calls of this frame
};

struct ReplayOut {  // written by the replay kernel, read by mask/refit kernels and copied to rpe_result
  float q[4];
  float t[3];
  int32_t max_votes;
  int32_t iter_final;
  int32_t winner;
  int32_t n_slots;
  int32_t n_borderline;
  int32_t flags;
  int32_t n_inliers[3];  // per mask column, filled by the mask kernel
  int32_t refit_ok;
};

// Sequential state of the keep-best / adaptive-stop rule, carried across chunks of iterations when the
// caller's Iter is larger than one device pass (e.g. SimpleMain.cpp's 100 000).
struct ReplayState {
  int32_t best;      // adapter.getMaxVotes()
  int32_t iter;      // current `Iter`
  int32_t win;       // global slot index of the accepted hypothesis
  int32_t cur_iter;  // iteration of the last accepted record
  int32_t stop;      // the `ii < Iter` loop has ended
  int32_t slots_done;
  int32_t borderline;
  int32_t overflow;
  float q[4];
  float t[3];
  int32_t pad;
};

// ---- exact-order primitives ------------------------------------------------------------------
struct F3 {
  float x, y, z;
};
__device__ __forceinline__ F3 f3(float x, float y, float z) {
  F3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
__device__ __forceinline__ float ex_sum3(float a, float b, float c) { return __fadd_rn(a, __fadd_rn(b, c)); }
__device__ __forceinline__ F3 ex_cross(F3 a, F3 b) {
  return f3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
            __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float ex_dot(F3 a, F3 b) {
  return ex_sum3(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ float ex_norm(F3 a) { return __fsqrt_rn(ex_dot(a, a)); }
// v + w*uv + qv x uv with uv = 2 (qv x v)
__device__ __forceinline__ F3 ex_quat_rotate(const float q[4], F3 v) {
  const F3 qv = f3(q[0], q[1], q[2]);
  F3 uv = ex_cross(qv, v);
  uv = f3(__fadd_rn(uv.x, uv.x), __fadd_rn(uv.y, uv.y), __fadd_rn(uv.z, uv.z));
  const F3 c2 = ex_cross(qv, uv);
  const float w = q[3];
  return f3(__fadd_rn(__fadd_rn(v.x, __fmul_rn(w, uv.x)), c2.x), __fadd_rn(__fadd_rn(v.y, __fmul_rn(w, uv.y)), c2.y),
            __fadd_rn(__fadd_rn(v.z, __fmul_rn(w, uv.z)), c2.z));
}
__device__ __forceinline__ void ex_quat_to_matrix(const float q[4], float R[9]) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  const float tx = __fmul_rn(2.f, x), ty = __fmul_rn(2.f, y), tz = __fmul_rn(2.f, z);
  const float twx = __fmul_rn(tx, w), twy = __fmul_rn(ty, w), twz = __fmul_rn(tz, w);
  const float txx = __fmul_rn(tx, x), txy = __fmul_rn(ty, x), txz = __fmul_rn(tz, x);
  const float tyy = __fmul_rn(ty, y), tyz = __fmul_rn(tz, y), tzz = __fmul_rn(tz, z);
  R[0] = __fsub_rn(1.f, __fadd_rn(tyy, tzz));
  R[1] = __fsub_rn(txy, twz);
  R[2] = __fadd_rn(txz, twy);
  R[3] = __fadd_rn(txy, twz);
  R[4] = __fsub_rn(1.f, __fadd_rn(txx, tzz));
  R[5] = __fsub_rn(tyz, twx);
  R[6] = __fsub_rn(txz, twy);
  R[7] = __fadd_rn(tyz, twx);
  R[8] = __fsub_rn(1.f, __fadd_rn(txx, tyy));
}
__device__ __forceinline__ F3 ex_mat_vec(const float R[9], F3 v) {
  return f3(ex_sum3(__fmul_rn(R[0], v.x), __fmul_rn(R[1], v.y), __fmul_rn(R[2], v.z)),
            ex_sum3(__fmul_rn(R[3], v.x), __fmul_rn(R[4], v.y), __fmul_rn(R[5], v.z)),
            ex_sum3(__fmul_rn(R[6], v.x), __fmul_rn(R[7], v.y), __fmul_rn(R[8], v.z)));
}
__device__ __forceinline__ bool ex_is_valid(F3 p) { return p.x == p.x || p.y == p.y || p.z == p.z; }

// The three inlier tests, in the reference's operation order.
__device__ __forceinline__ bool ex_test_3d(const float q[4], const float t[3], F3 xw, F3 xc, float thr3d) {
  const F3 r = ex_quat_rotate(q, xw);
  const F3 y = f3(__fadd_rn(r.x, t[0]), __fadd_rn(r.y, t[1]), __fadd_rn(r.z, t[2]));
  const F3 e = f3(__fsub_rn(xc.x, y.x), __fsub_rn(xc.y, y.y), __fsub_rn(xc.z, y.z));
  return ex_norm(e) < thr3d;
}
__device__ __forceinline__ bool ex_test_2d(const float q[4], const float t[3], const float* Rm /*null: quaternion form*/,
                                           F3 xw, F3 bv, float cos_thr) {
  const F3 r = Rm ? ex_mat_vec(Rm, xw) : ex_quat_rotate(q, xw);
  F3 pc = f3(__fadd_rn(r.x, t[0]), __fadd_rn(r.y, t[1]), __fadd_rn(r.z, t[2]));
  const float nrm = ex_norm(pc);
  pc = f3(__fdiv_rn(pc.x, nrm), __fdiv_rn(pc.y, nrm), __fdiv_rn(pc.z, nrm));
  return ex_dot(pc, bv) > cos_thr;
}
__device__ __forceinline__ bool ex_test_nl(const float q[4], F3 nw, F3 nc, float cos_nl) {
  return ex_dot(nc, ex_quat_rotate(q, nw)) > cos_nl;
}

__device__ __forceinline__ F3 load_col(const float* __restrict__ a, int i) {
  return f3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
}

// atomicMax on non-negative floats through their bit patterns
__device__ __forceinline__ void atomic_max_nonneg(unsigned int* addr, float v) {
  if (v == v && v >= 0.f) atomicMax(addr, __float_as_uint(v));
}

}  // namespace rpe

#endif  // RPE_DEVICE_CUH_
