// simulate.cu — device-side synthetic correspondences (SURVEY.md §8f "next" row 3).
//
// Same generators as include/rpe/sim_core.hpp (restating /root/reference/pose/Simulator.hpp:85-367) with a
// counter-based random stream, so that a 640x480 frame is produced directly in HBM instead of crossing PCIe
// (config #5: 4096 frames x 7.4 MB). Only the DISTRIBUTIONS are part of the reference's contract; differences
// from the host generator, by construction of a parallel generator:
//   * every random number is hash(seed, stream, index, draw) (splitmix64 finaliser) instead of a sequential stream;
//   * the `out` outlier positions are the image of [0,out) under a seeded affine permutation of [0,n)
//     (exactly `out` distinct positions, like RandomElements::run, Simulator.hpp:299-303) instead of a
//     partial Fisher-Yates shuffle.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "kernels.cuh"

namespace rpe {

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
struct CRng {  // counter-based: (seed, stream, index) fixed, `draw` advances
  unsigned long long key;
  unsigned int draw;
  __device__ CRng(unsigned long long seed, unsigned int stream, unsigned int index)
      : key(mix64(seed ^ ((unsigned long long)stream << 40) ^ ((unsigned long long)index * 0x9E3779B1ULL))), draw(0) {}
  __device__ float unit() {  // [0,1)
    const unsigned long long r = mix64(key + (unsigned long long)(draw++) * 0xD1B54A32D192ED03ULL);
    return (float)(r >> 40) * (1.0f / 16777216.0f);
  }
  __device__ float pm1() { return 2.f * unit() - 1.f; }
  __device__ float normal() {  // Box-Muller (one value per call; distribution only)
    float u1 = unit();
    const float u2 = unit();
    u1 = fmaxf(u1, 1e-12f);
    return sqrtf(-2.f * logf(u1)) * cosf(6.28318530718f * u2);
  }
};


__device__ __forceinline__ void frustum_point(CRng& g, float f, float dmin, float dmax, float* P) {
  const float tx = 320.f / f, ty = 240.f / f;
  for (int k = 0; k < 64; ++k) {
    const float x = g.pm1() * tx * dmax, y = g.pm1() * ty * dmax;
    const float z = (g.pm1() + 1.f) * 0.5f * (dmax - dmin) + dmin;
    P[0] = x;
    P[1] = y;
    P[2] = z;
    if (fabsf(x / z) < tx && fabsf(y / z) < ty) return;
  }
}
__device__ __forceinline__ bool is_outlier(const SimParams& p, int which, unsigned int i) {
  const unsigned int n = (unsigned int)p.n;
  const unsigned long long d = ((unsigned long long)i + n - p.b[which] % n) % n;
  const unsigned int j = (unsigned int)((d * (unsigned long long)p.ainv[which]) % n);
  const int out = which == 0 ? p.out2d : (which == 1 ? p.out3d : p.outnl);
  return (int)j < out;
}
// random rotation R = Rz Ry Rx from three draws (Simulator.hpp:23-83)
__device__ __forceinline__ void rand_rot(CRng& g, float max_angle, bool gaussian, float* M) {
  float rv[3];
  for (int k = 0; k < 3; ++k) rv[k] = gaussian ? g.normal() : g.pm1();
  const float pi = 3.14159265358979f;
  rv[0] = fminf(fmaxf(max_angle * rv[0], -pi), pi);
  rv[1] = fminf(fmaxf(max_angle * rv[1] * 0.5f, -pi / 2), pi / 2);
  rv[2] = fminf(fmaxf(max_angle * rv[2], -pi), pi);
  float sx, cx, sy, cy, sz, cz;
  sincosf(rv[0], &sx, &cx);
  sincosf(rv[1], &sy, &cy);
  sincosf(rv[2], &sz, &cz);
  M[0] = cz * cy;  M[1] = cz * sy * sx - sz * cx;  M[2] = cz * sy * cx + sz * sx;
  M[3] = sz * cy;  M[4] = sz * sy * sx + cz * cx;  M[5] = sz * sy * cx - cz * sx;
  M[6] = -sy;      M[7] = cy * sx;                 M[8] = cy * cx;
}
__device__ __forceinline__ void facing_normal(CRng& g, float max_angle, bool gaussian, const float* src, float* dst) {
  float M[9];
  rand_rot(g, max_angle, gaussian, M);
  for (int r = 0; r < 3; ++r) dst[r] = M[3 * r] * src[0] + M[3 * r + 1] * src[1] + M[3 * r + 2] * src[2];
  const float nn = sqrtf(dst[0] * dst[0] + dst[1] * dst[1] + dst[2] * dst[2]);
  dst[0] /= nn;
  dst[1] /= nn;
  dst[2] /= nn;
}

// One thread per correspondence. Any output pointer may be null.
//   xw: world points (3-D noise / outliers as in simulate_3d_3d when bv == null, clean otherwise: Simulator.hpp:193)
//   xc: camera points (clean for 3d_3d; noisy + outliers for the nl generator, :339-359)
//   bv: unit bearing vectors with pixel noise + outliers (:195-223); nw, nc: normals (:85-130)
__global__ void sim_kernel(SimParams p, float* __restrict__ xw, float* __restrict__ xc, float* __restrict__ bv,
                           float* __restrict__ nw, float* __restrict__ nc, int mode_3d3d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  CRng g(p.seed, 0, (unsigned int)i);
  float P[3];
  frustum_point(g, p.f, p.min_depth, p.max_depth, P);
  float Q[3];
  {
    const float d[3] = {P[0] - p.t[0], P[1] - p.t[1], P[2] - p.t[2]};
    for (int r = 0; r < 3; ++r) Q[r] = p.R[r] * d[0] + p.R[3 + r] * d[1] + p.R[6 + r] * d[2];  // R^T d
  }
  if (mode_3d3d) {
    // simulate_3d_3d_correspondences: noise and outliers live in the WORLD points, camera points stay clean
    for (int r = 0; r < 3; ++r) Q[r] += p.noise3d * (p.gaussian ? g.normal() : g.pm1());
    if (is_outlier(p, 1, (unsigned int)i)) {
      CRng go(p.seed, 1, (unsigned int)i);
      frustum_point(go, p.f, p.min_depth, p.max_depth, Q);  // outliers are raw frustum points (:298-303)
    }
    for (int r = 0; r < 3; ++r) {
      xw[3 * (size_t)i + r] = Q[r];
      xc[3 * (size_t)i + r] = P[r];
    }
    return;
  }
  for (int r = 0; r < 3; ++r) xw[3 * (size_t)i + r] = Q[r];
  if (bv) {
    float kx = p.f * P[0] / P[2], ky = p.f * P[1] / P[2];
    kx += p.noise2d * (p.gaussian ? g.normal() : g.pm1());
    ky += p.noise2d * (p.gaussian ? g.normal() : g.pm1());
    if (is_outlier(p, 0, (unsigned int)i)) {
      CRng go(p.seed, 2, (unsigned int)i);
      float O[3];
      frustum_point(go, p.f, p.min_depth, p.max_depth, O);
      kx = p.f * O[0] / O[2];
      ky = p.f * O[1] / O[2];
    }
    const float nn = sqrtf(kx * kx + ky * ky + p.f * p.f);
    bv[3 * (size_t)i] = kx / nn;
    bv[3 * (size_t)i + 1] = ky / nn;
    bv[3 * (size_t)i + 2] = p.f / nn;
  }
  float ngt[3] = {0.f, 0.f, -1.f};  // true camera-frame normal (frontal when the normal channel is absent)
  if (nw && nc) {
    CRng gn(p.seed, 3, (unsigned int)i);
    const float back[3] = {0.f, 0.f, -1.f};
    float m[3], nn[3];
    for (int k = 0; k < 64; ++k) {
      facing_normal(gn, 1.57079632679f, false, back, ngt);
      for (int r = 0; r < 3; ++r) m[r] = p.R[r] * ngt[0] + p.R[3 + r] * ngt[1] + p.R[6 + r] * ngt[2];
      facing_normal(gn, p.noise_nl, true, ngt, nn);
      if (!(nn[2] > 0.f)) break;  // keep normals that face the camera (:106)
    }
    if (is_outlier(p, 2, (unsigned int)i)) {
      CRng go(p.seed, 4, (unsigned int)i);
      for (int k = 0; k < 64; ++k) {
        facing_normal(go, 1.57079632679f, false, back, nn);
        if (!(nn[2] > 0.f)) break;
      }
    }
    const float mn = sqrtf(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    for (int r = 0; r < 3; ++r) {
      nw[3 * (size_t)i + r] = m[r] / mn;
      nc[3 * (size_t)i + r] = nn[r];
    }
  }
  if (xc) {
    float C[3] = {P[0], P[1], P[2]};
    CRng gc(p.seed, 5, (unsigned int)i);
    if (p.kinect) {
      // Nguyen, Izadi & Lovell (3DIMPVT 2012), as used at Simulator.hpp:368-423: lateral sigma (pixels -> metres) on x, y,
      // axial sigma on z; theta = angle between the true normal and the optical axis towards the camera
      const float half_pi = 1.57079632679f;
      const float theta = acosf(fminf(fmaxf(-ngt[2], -1.f), 1.f));
      const float z = P[2];
      const float sl = (0.8f + 0.035f * theta / (half_pi - theta)) * z / p.f;
      float sa = 0.0012f + 0.0019f * (z - 0.4f) * (z - 0.4f);
      if (fabsf(theta) > 1.0471975512f) sa += 0.0001f * theta * theta / sqrtf(z) / ((half_pi - theta) * (half_pi - theta));
      C[0] += sl * gc.normal();
      C[1] += sl * gc.normal();
      C[2] += sa * gc.normal();
    } else {
      for (int r = 0; r < 3; ++r) C[r] += p.noise3d * (p.gaussian ? gc.normal() : gc.pm1());
    }
    if (is_outlier(p, 1, (unsigned int)i)) {
      CRng go(p.seed, 6, (unsigned int)i);
      frustum_point(go, p.f, p.min_depth, p.max_depth, C);
    }
    for (int r = 0; r < 3; ++r) xc[3 * (size_t)i + r] = C[r];
  }
}

void launch_simulate(const SimParams& p, float* xw, float* xc, float* bv, float* nw, float* nc, int mode_3d3d, cudaStream_t s) {
  sim_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p, xw, xc, bv, nw, nc, mode_3d3d);
}

}  // namespace rpe
