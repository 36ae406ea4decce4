// score.cu — hypothesis x correspondence scoring for sm_100a.
//
// Replaces the reference's inner `for c` loops (the ~100 % hot spot of the CPU path):
//   3-D test   pose/AbsoluteOrientation.hpp:135-143 (=:192-200, :250-258, :403-411)
//   2-D test   pose/P3P.hpp:362-376 / :439-453, AbsoluteOrientation.hpp:413-421
//   normal     pose/AbsoluteOrientationNormal.hpp:245-253, :322-329, :397-404
//
// Design: one thread owns kHypPerThread hypotheses in registers (-R, -t, each value
// duplicated into an f32x2 register pair); a CTA streams its slice of the correspondences through a
// double-buffered shared-memory ring filled by 1-D bulk TMA (cp.async.bulk + mbarrier); every lane
// reads the same pair record (LDS.128 broadcast) so one record feeds 32 x kHypPerThread x 2
// evaluations. Votes accumulate in per-thread registers — no cross-thread reduction per evaluation —
// and are added to the global vote table once per CTA.
//
// Exactness: the fast path evaluates  s = |x_c - t - R x_w|^2 - thr^2  with 12 (packed) FMA + 3 ADD.
// sign(s) decides, unless |s| <= band, where `band` rigorously covers the rounding error of BOTH
// this path and the reference's quaternion-sandwich/sqrt path. Borderline
// evaluations are removed from the fast count and appended to a worklist which fixup_kernel
// re-evaluates in the reference's exact operation order. Inlier counts are therefore identical to the
// CPU path's, evaluation by evaluation.
#include <atomic>
#include <cstdlib>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "kernels.cuh"
#include "score_common.cuh"

namespace rpe {

// ================================================================================================
// reset / pack
// ================================================================================================
__global__ void reset_stats_kernel(FrameStats* st) {
  if (threadIdx.x == 0) {  // m_corr_bits belongs to the packed frame and is reset by the pack only
    st->t_max_bits = 0;
    st->wl_count = 0;
    st->wl_overflow = 0;
    st->ticket = 0;
    st->ticket2 = 0;
    st->wl_consumed = 0;
  }
}
void launch_reset_stats(FrameStats* st, cudaStream_t s) { reset_stats_kernel<<<1, 32, 0, s>>>(st); }
__global__ void reset_corr_bound_kernel(FrameStats* st) {
  if (threadIdx.x == 0) {
    st->m_corr_bits = 0;
    st->m_bv_bits = 0;
  }
}
void launch_reset_corr_bound(FrameStats* st, cudaStream_t s) { reset_corr_bound_kernel<<<1, 32, 0, s>>>(st); }

struct PackSrc {
  const float* a[5];
  int count;
  int idx_nw, idx_nc;     // positions of the normal arrays among the sources, -1 if absent
  const float* valid_xc;  // raw camera points: an all-NaN point makes its packed camera normal NaN (isValid gate)
};

// A CTA packs kPackPairs pairs of correspondences (2j, 2j+1). The raw xyz triples of the CTA's slice are
// first staged in shared memory with fully coalesced loads, then every thread emits whole float4 records at
// consecutive addresses. Record layout per pair, for every source array in order:
//   x(2j) x(2j+1) y(2j) y(2j+1) z(2j) z(2j+1)   — each 64-bit half is a ready-made f32x2 operand.
// Also reduces max_i(|x_w| + |x_c|) over finite points for the guard band.
constexpr int kPackPairs = 256;
__global__ void __launch_bounds__(256)
pack_kernel(PackSrc src, int n, int npairs_pad, int f4_per_pair, float4* __restrict__ out, int has_xc,
            FrameStats* st) {
  extern __shared__ float sm_raw[];  // [count][kPackPairs * 6]
  const int pair0 = blockIdx.x * kPackPairs;
  const int c_base = 2 * pair0;
  const int floats = kPackPairs * 6;
  for (int k = 0; k < src.count; ++k) {
    const float* a = src.a[k];
    for (int i = threadIdx.x; i < floats; i += blockDim.x) {
      const long long g = (long long)c_base * 3 + i;
      sm_raw[k * floats + i] = (g < (long long)n * 3) ? a[g] : CUDART_NAN_F;
    }
  }
  __syncthreads();
  const int pairs_here = min(kPackPairs, npairs_pad - pair0);
  // The normal test sits inside `if (adapter.isValid(c))` (AbsoluteOrientationNormal.hpp:246,323,398): a correspondence
  // whose camera point is all-NaN must never cast a normal vote -> poison its packed camera normal.
  float nloc = 0.f;
  if (src.idx_nc >= 0) {
    for (int i = threadIdx.x; i < 2 * pairs_here; i += blockDim.x) {
      const int c = c_base + i;
      float* pn = &sm_raw[src.idx_nc * floats + i * 3];
      if (c < n) {
        const float* pw = &sm_raw[src.idx_nw * floats + i * 3];
        const float a = sqrtf(pn[0] * pn[0] + pn[1] * pn[1] + pn[2] * pn[2]);
        const float b = sqrtf(pw[0] * pw[0] + pw[1] * pw[1] + pw[2] * pw[2]);
        if (a == a && a < CUDART_INF_F) nloc = fmaxf(nloc, a);
        if (b == b && b < CUDART_INF_F) nloc = fmaxf(nloc, b);
        const float v0 = src.valid_xc[3 * (size_t)c], v1 = src.valid_xc[3 * (size_t)c + 1], v2 = src.valid_xc[3 * (size_t)c + 2];
        if (!(v0 == v0 || v1 == v1 || v2 == v2)) pn[0] = pn[1] = pn[2] = CUDART_NAN_F;
      }
    }
    __syncthreads();
  }
  const int total_f4 = pairs_here * f4_per_pair;
  const int fpp = f4_per_pair * 4;
  for (int o = threadIdx.x; o < total_f4; o += blockDim.x) {
    const int j = o / f4_per_pair;
    const int part = o - j * f4_per_pair;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int w = part * 4 + e;  // float index inside the pair record
      const int k = w / 6;
      const int r = (w - k * 6) >> 1;
      const int u = w & 1;
      v[e] = (k < src.count) ? sm_raw[k * floats + (2 * j + u) * 3 + r] : 0.f;
    }
    (void)fpp;
    out[(size_t)pair0 * f4_per_pair + o] = make_float4(v[0], v[1], v[2], v[3]);
  }
  // magnitude bound: thread <-> pair, source 0 is x_w, source 1 is x_c when present
  float mloc = 0.f;
  if (threadIdx.x < pairs_here) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int c = c_base + 2 * threadIdx.x + u;
      if (c < n) {
        const float* pw = &sm_raw[(2 * threadIdx.x + u) * 3];
        float m = sqrtf(pw[0] * pw[0] + pw[1] * pw[1] + pw[2] * pw[2]);
        if (has_xc) {
          const float* pc = &sm_raw[floats + (2 * threadIdx.x + u) * 3];
          const float mc = sqrtf(pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]);
          if (mc == mc && mc < CUDART_INF_F) m += mc;
        }
        if (m == m && m < CUDART_INF_F) mloc = fmaxf(mloc, m);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, o));
  if ((threadIdx.x & 31) == 0 && mloc > 0.f) atomic_max_nonneg(&st->m_corr_bits, mloc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nloc = fmaxf(nloc, __shfl_xor_sync(0xffffffffu, nloc, o));
  if ((threadIdx.x & 31) == 0 && nloc > 0.f) atomic_max_nonneg(&st->m_bv_bits, nloc);
}

void launch_pack(const FrameView& f, int kind, float4* pk_out, FrameStats* st, cudaStream_t s) {
  PackSrc src;
  src.count = 0;
  src.idx_nw = src.idx_nc = -1;
  src.valid_xc = f.xc;
  src.a[src.count++] = f.xw;
  if (kind & 2) src.a[src.count++] = f.xc;
  if (kind & 1) src.a[src.count++] = f.bv;
  if (kind & 4) {
    src.idx_nw = src.count;
    src.a[src.count++] = f.nw;
    src.idx_nc = src.count;
    src.a[src.count++] = f.nc;
  }
  const int blocks = (f.npairs_pad + kPackPairs - 1) / kPackPairs;
  const size_t smem = (size_t)src.count * kPackPairs * 6 * sizeof(float);
  pack_kernel<<<blocks, 256, smem, s>>>(src, f.n, f.npairs_pad, f.pk_f4_per_pair, pk_out, (kind & 2) ? 1 : 0, st);
}

// ================================================================================================
// fast tiled scorer — 3-D modality (RPE_SHINJI)
// ================================================================================================
// Guard band on s = r^2 - thr^2:  band = thr * u * (64 M + 16 thr),  u = 2^-24,
// M >= |x_w| + |x_c| + |t| for every correspondence of the frame and THIS hypothesis (per-hypothesis |t|: one wild
// hypothesis must not widen the band of the others).
__device__ __forceinline__ float hyp_magnitude(const FrameStats* st, float nt0, float nt1, float nt2) {
  const float tn = sqrtf(nt0 * nt0 + nt1 * nt1 + nt2 * nt2);
  return (__uint_as_float(st->m_corr_bits) + tn) * 1.0001f;
}
template <bool PACKED>
struct HypRegs;

template <>
struct HypRegs<true> {
  float2 nR[9];
  float2 nt[3];
  __device__ __forceinline__ void load(const HypFast* h, bool live) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float v = live ? h->nR[i] : CUDART_NAN_F;
      nR[i] = make_float2(v, v);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float v = live ? h->nt[i] : CUDART_NAN_F;
      nt[i] = make_float2(v, v);
    }
  }
  __device__ __forceinline__ float magnitude(const FrameStats* st) const { return hyp_magnitude(st, nt[0].x, nt[1].x, nt[2].x); }
  // s for both correspondences of the pair
  __device__ __forceinline__ float2 eval(const float4& a, const float4& b, const float4& c, float2 nlo) const {
    const float2 X0 = make_float2(a.x, a.y), X1 = make_float2(a.z, a.w), X2 = make_float2(b.x, b.y);
    const float2 P0 = make_float2(b.z, b.w), P1 = make_float2(c.x, c.y), P2 = make_float2(c.z, c.w);
    float2 e0 = __fadd2_rn(P0, nt[0]);
    float2 e1 = __fadd2_rn(P1, nt[1]);
    float2 e2 = __fadd2_rn(P2, nt[2]);
    e0 = __ffma2_rn(nR[0], X0, e0);
    e1 = __ffma2_rn(nR[3], X0, e1);
    e2 = __ffma2_rn(nR[6], X0, e2);
    e0 = __ffma2_rn(nR[1], X1, e0);
    e1 = __ffma2_rn(nR[4], X1, e1);
    e2 = __ffma2_rn(nR[7], X1, e2);
    e0 = __ffma2_rn(nR[2], X2, e0);
    e1 = __ffma2_rn(nR[5], X2, e1);
    e2 = __ffma2_rn(nR[8], X2, e2);
    float2 s = __ffma2_rn(e0, e0, nlo);
    s = __ffma2_rn(e1, e1, s);
    s = __ffma2_rn(e2, e2, s);
    return s;
  }
};

template <>
struct HypRegs<false> {
  float nR[9];
  float nt[3];
  __device__ __forceinline__ void load(const HypFast* h, bool live) {
#pragma unroll
    for (int i = 0; i < 9; ++i) nR[i] = live ? h->nR[i] : CUDART_NAN_F;
#pragma unroll
    for (int i = 0; i < 3; ++i) nt[i] = live ? h->nt[i] : CUDART_NAN_F;
  }
  __device__ __forceinline__ float magnitude(const FrameStats* st) const { return hyp_magnitude(st, nt[0], nt[1], nt[2]); }
  __device__ __forceinline__ float one(float x0, float x1, float x2, float p0, float p1, float p2, float nlo) const {
    float e0 = p0 + nt[0], e1 = p1 + nt[1], e2 = p2 + nt[2];
    e0 = fmaf(nR[0], x0, e0);
    e1 = fmaf(nR[3], x0, e1);
    e2 = fmaf(nR[6], x0, e2);
    e0 = fmaf(nR[1], x1, e0);
    e1 = fmaf(nR[4], x1, e1);
    e2 = fmaf(nR[7], x1, e2);
    e0 = fmaf(nR[2], x2, e0);
    e1 = fmaf(nR[5], x2, e1);
    e2 = fmaf(nR[8], x2, e2);
    float s = fmaf(e0, e0, nlo);
    s = fmaf(e1, e1, s);
    s = fmaf(e2, e2, s);
    return s;
  }
  __device__ __forceinline__ float2 eval(const float4& a, const float4& b, const float4& c, float2 nlo) const {
    return make_float2(one(a.x, a.z, b.x, b.z, c.x, c.z, nlo.x), one(a.y, a.w, b.y, b.w, c.y, c.w, nlo.y));
  }
};

template <bool PACKED, int HPT, int TILE, int THREADS, int MINB, int SUB>
__global__ void __launch_bounds__(THREADS, MINB)
score3d_fast_kernel(const float4* __restrict__ pk, int npairs_pad, int pairs_per_cta, const HypFast* __restrict__ fast,
                    const HypGen* __restrict__ gen, int slot_begin, int slot_end, float thr, int32_t* __restrict__ votes,
                    FrameStats* __restrict__ st, Worklist wl, int nosync) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4(*tile)[TILE * 3] = reinterpret_cast<float4(*)[TILE * 3]>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * TILE * 3 * sizeof(float4));
  constexpr int kTilePairs = TILE;
  constexpr int kHypPerThread = HPT;
  constexpr int kScoreThreads = THREADS;
  constexpr int kSubPairs = SUB;

  const int tid = threadIdx.x;
  const int p_begin = blockIdx.x * pairs_per_cta;
  const int p_end = min(p_begin + pairs_per_cta, npairs_pad);
  const int npairs = p_end - p_begin;
  const int ntiles = (npairs + kTilePairs - 1) / kTilePairs;
  WlSegment seg(wl);

  // hypotheses of this thread
  HypRegs<PACKED> hyp[kHypPerThread];
  int slot[kHypPerThread];
  int cnt[kHypPerThread];
#pragma unroll
  for (int k = 0; k < kHypPerThread; ++k) {
    slot[k] = slot_begin + (blockIdx.y * kHypPerThread + k) * kScoreThreads + tid;
    const bool live = slot[k] < slot_end && gen[slot[k]].valid != 0;
    hyp[k].load(&fast[live ? slot[k] : slot_begin], live);
    if (!live) slot[k] = -1;
    cnt[k] = 0;
  }
  float band[kHypPerThread];
#pragma unroll
  for (int k = 0; k < kHypPerThread; ++k) band[k] = guard_band_3d(hyp[k].magnitude(st), thr);  // NaN for a dead slot
  const float thr2 = __fmul_rn(thr, thr);
  const float2 nlo = make_float2(-thr2, -thr2);

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int t = 0; t < (nosync ? 1 : 2) && t < ntiles; ++t) {
      const int tp = min(kTilePairs, npairs - t * kTilePairs);
      const uint32_t bytes = (uint32_t)tp * 48u;
      mbar_expect_tx(&bars[t], bytes);
      tma_load_1d(&tile[t][0], pk + (size_t)(p_begin + t * kTilePairs) * 3, bytes, &bars[t]);
    }
  }

  for (int t = 0; t < ntiles; ++t) {
    const int buf = nosync ? 0 : (t & 1);
    if (!nosync || t == 0) mbar_wait(&bars[buf], (uint32_t)((t >> 1) & 1));
    const int tp = min(kTilePairs, npairs - t * kTilePairs);
    const float4* sp = &tile[buf][0];
    for (int sub = 0; sub < tp; sub += kSubPairs) {
      // smallest |s| of the group: one 3-input FMNMX3 per pair and hypothesis (NaN operands are ignored, and a NaN
      // evaluation is never an inlier, so it never needs the exact path)
      float smin[kHypPerThread];
#pragma unroll
      for (int k = 0; k < kHypPerThread; ++k) smin[k] = CUDART_INF_F;
#pragma unroll
      for (int pp = 0; pp < kSubPairs; ++pp) {
        const float4 a = sp[(sub + pp) * 3 + 0];
        const float4 b = sp[(sub + pp) * 3 + 1];
        const float4 c = sp[(sub + pp) * 3 + 2];
#pragma unroll
        for (int k = 0; k < kHypPerThread; ++k) {
          const float2 s = hyp[k].eval(a, b, c, nlo);
          cnt[k] += (int)(__float_as_uint(s.x) >> 31) + (int)(__float_as_uint(s.y) >> 31);
          smin[k] = fminf(fminf(smin[k], fabsf(s.x)), fabsf(s.y));
        }
      }
      bool any = false;
#pragma unroll
      for (int k = 0; k < kHypPerThread; ++k) any = any || (smin[k] <= band[k]);
      if (any) {
        // Rare: some evaluation of this 16-correspondence group sits inside the guard band.
        // Re-walk the group, take the borderline evaluations OUT of the fast count and queue them.
        for (int pp = 0; pp < kSubPairs; ++pp) {
          const float4 a = sp[(sub + pp) * 3 + 0];
          const float4 b = sp[(sub + pp) * 3 + 1];
          const float4 c = sp[(sub + pp) * 3 + 2];
#pragma unroll
          for (int k = 0; k < kHypPerThread; ++k) {
            const float2 s = hyp[k].eval(a, b, c, nlo);
            const float sv[2] = {s.x, s.y};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (fabsf(sv[u]) <= band[k]) {
                cnt[k] -= (int)(__float_as_uint(sv[u]) >> 31);
                const unsigned int corr = (unsigned int)(2 * (p_begin + t * kTilePairs + sub + pp) + u);
                seg.push(make_uint2((unsigned int)slot[k], corr | (1u << 30)), st);
              }
            }
          }
        }
      }
    }
    if (nosync) continue;
    __syncthreads();
    if (tid == 0 && t + 2 < ntiles) {
      const int tn = t + 2;
      const int tpn = min(kTilePairs, npairs - tn * kTilePairs);
      const uint32_t bytes = (uint32_t)tpn * 48u;
      mbar_expect_tx(&bars[buf], bytes);
      tma_load_1d(&tile[buf][0], pk + (size_t)(p_begin + tn * kTilePairs) * 3, bytes, &bars[buf]);
    }
  }
#pragma unroll
  for (int k = 0; k < kHypPerThread; ++k)
    if (slot[k] >= 0 && cnt[k] != 0) atomicAdd(&votes[slot[k]], cnt[k]);
  seg.publish(wl, st);
}

// ================================================================================================
// the same scorer fed from the caller's arrays (no packed copy of the frame)
// ================================================================================================
// The 3-D / 3-D frame is two column-major 3 x n arrays, i.e. n contiguous xyz triples each. A stage is TILE pairs =
// 2 TILE correspondences: two 1-D bulk TMA copies (24 TILE bytes each) land the raw triples in shared memory, the CTA
// transposes them once into the pair-interleaved records the FFMA2 loop wants (x0 x1 y0 y1 | z0 z1 px0 px1 | py0 py1
// pz0 pz1; 6 LDS.128 + 6 STS.128 per thread against ~40 000 instructions of scoring per stage) and takes the stage's
// own magnitude bound max(|x_w| + |x_c|) on the way — the guard band then uses the bound of the correspondences it is
// applied to instead of the frame maximum. No pack kernel, no second copy of the frame in HBM.
// TMA needs 16-byte aligned addresses and sizes: stages start at multiples of 4 correspondences and copy a multiple
// of 4; the last 0..3 correspondences of the frame are read from global memory by the transposing threads.
template <int HPT, int TILE, int THREADS, int MINB, int SUB, int PCOUNT>
__global__ void __launch_bounds__(THREADS, MINB)
score3d_raw_kernel(const float* __restrict__ xw, const float* __restrict__ xc, int n, int npairs_pad, int pairs_per_cta,
                   const HypFast* __restrict__ fast, const HypGen* __restrict__ gen, int slot_begin, int slot_end, float thr,
                   int32_t* __restrict__ votes, FrameStats* __restrict__ st, Worklist wl, int corr_base) {
  // corr_base: index, in the whole frame, of the first correspondence of xw / xc (a launch may score one uploaded chunk)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kStageFloats = TILE * 2 * 3;  // floats of one array in one stage
  float4* packed = reinterpret_cast<float4*>(smem_raw);                                  // [TILE * 3]
  float* raw = reinterpret_cast<float*>(smem_raw + (size_t)TILE * 3 * sizeof(float4));    // [2 stages][xw | xc][kStageFloats]
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw + 4 * kStageFloats);
  unsigned int* mtile = reinterpret_cast<unsigned int*>(bars + 2);

  const int tid = threadIdx.x;
  const int p_begin = blockIdx.x * pairs_per_cta;
  const int p_end = min(p_begin + pairs_per_cta, npairs_pad);
  const int npairs = p_end - p_begin;
  const int ntiles = (npairs + TILE - 1) / TILE;
  WlSegment seg(wl);

  HypRegs<true> hyp[HPT];
  int slot[HPT];
  int cnt[HPT];
  unsigned int pacc[HPT];
  float tnorm[HPT];
#pragma unroll
  for (int k = 0; k < HPT; ++k) {
    slot[k] = slot_begin + (blockIdx.y * HPT + k) * THREADS + tid;
    const bool live = slot[k] < slot_end && gen[slot[k]].valid != 0;
    hyp[k].load(&fast[live ? slot[k] : slot_begin], live);
    if (!live) slot[k] = -1;
    cnt[k] = 0;
    pacc[k] = 0u;
    tnorm[k] = sqrtf(hyp[k].nt[0].x * hyp[k].nt[0].x + hyp[k].nt[1].x * hyp[k].nt[1].x + hyp[k].nt[2].x * hyp[k].nt[2].x);
  }
  const float thr2 = __fmul_rn(thr, thr);
  const float2 nlo = make_float2(-thr2, -thr2);

  // correspondences [c0, c0 + cnt4) of stage t go through TMA (cnt4 a multiple of 4, possibly 0)
  auto stage_range = [&](int t, int& c0, int& cnt4) {
    c0 = 2 * (p_begin + t * TILE);
    const int tp = min(TILE, npairs - t * TILE);
    int c = min(2 * tp, n - c0);
    if (c < 0) c = 0;
    cnt4 = c & ~3;
  };
  auto issue = [&](int t, int buf) {
    int c0, cnt4;
    stage_range(t, c0, cnt4);
    const uint32_t bytes = (uint32_t)cnt4 * 12u;
    mbar_expect_tx(&bars[buf], 2u * bytes);  // 0 bytes: the phase completes on this arrival alone
    if (bytes) {
      tma_load_1d(raw + (size_t)(2 * buf) * kStageFloats, xw + (size_t)c0 * 3, bytes, &bars[buf]);
      tma_load_1d(raw + (size_t)(2 * buf + 1) * kStageFloats, xc + (size_t)c0 * 3, bytes, &bars[buf]);
    }
  };
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
    mtile[0] = mtile[1] = 0u;
  }
  __syncthreads();
  if (tid == 0)
    for (int t = 0; t < 2 && t < ntiles; ++t) issue(t, t);

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    mbar_wait(&bars[buf], (uint32_t)((t >> 1) & 1));
    const int tp = min(TILE, npairs - t * TILE);
    int c0, cnt4;
    stage_range(t, c0, cnt4);
    // ---- transpose the stage: thread <-> 4 correspondences = 2 pair records; stage magnitude bound
    const float* rw = raw + (size_t)(2 * buf) * kStageFloats;
    const float* rc = raw + (size_t)(2 * buf + 1) * kStageFloats;
    float mloc = 0.f;
    for (int qd = tid; qd < (tp + 1) / 2; qd += THREADS) {
      float w[12], c[12];
      if (4 * qd + 4 <= cnt4) {
        const float4* w4 = reinterpret_cast<const float4*>(rw + 12 * qd);
        const float4* c4 = reinterpret_cast<const float4*>(rc + 12 * qd);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float4 a = w4[i], b = c4[i];
          w[4 * i] = a.x; w[4 * i + 1] = a.y; w[4 * i + 2] = a.z; w[4 * i + 3] = a.w;
          c[4 * i] = b.x; c[4 * i + 1] = b.y; c[4 * i + 2] = b.z; c[4 * i + 3] = b.w;
        }
      } else {  // frame tail (or padding): element-wise, from global memory, NaN beyond the last correspondence
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ci = c0 + 4 * qd + j;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            w[3 * j + r] = ci < n ? xw[(size_t)ci * 3 + r] : CUDART_NAN_F;
            c[3 * j + r] = ci < n ? xc[(size_t)ci * 3 + r] : CUDART_NAN_F;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float m = sqrtf(w[3 * j] * w[3 * j] + w[3 * j + 1] * w[3 * j + 1] + w[3 * j + 2] * w[3 * j + 2]);
        const float mc = sqrtf(c[3 * j] * c[3 * j] + c[3 * j + 1] * c[3 * j + 1] + c[3 * j + 2] * c[3 * j + 2]);
        if (mc == mc && mc < CUDART_INF_F) m += mc;
        if (m == m && m < CUDART_INF_F) mloc = fmaxf(mloc, m);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {  // pair 2 qd + h = correspondences 4 qd + 2h, 4 qd + 2h + 1
        const int pr = 2 * qd + h;
        if (pr < tp) {
          const float* a = w + 6 * h;
          const float* b = c + 6 * h;
          packed[pr * 3 + 0] = make_float4(a[0], a[3], a[1], a[4]);
          packed[pr * 3 + 1] = make_float4(a[2], a[5], b[0], b[3]);
          packed[pr * 3 + 2] = make_float4(b[1], b[4], b[2], b[5]);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, o));
    if ((tid & 31) == 0 && mloc > 0.f) atomicMax(&mtile[buf], __float_as_uint(mloc));
    __syncthreads();  // records + bound complete; raw[buf] is free
    if (tid == 0) {
      if (t + 2 < ntiles) issue(t + 2, buf);
      mtile[buf ^ 1] = 0u;  // the next stage's slot: untouched until everybody has passed the barrier below
    }
    const float mcorr = __uint_as_float(mtile[buf]);
    float band[HPT];
#pragma unroll
    for (int k = 0; k < HPT; ++k) band[k] = guard_band_3d((mcorr + tnorm[k]) * 1.0001f, thr);  // NaN for a dead slot

    const float4* sp = packed;
    for (int sub = 0; sub < tp; sub += SUB) {
      float smin[HPT];
#pragma unroll
      for (int k = 0; k < HPT; ++k) smin[k] = CUDART_INF_F;
      if (PCOUNT) {
        // Packed sign count: one PRMT with sign replication turns the two decision values of a pair into
        // (s.y < 0 ? 0xffff0000 : 0) | (s.x < 0 ? 0x0000ffff : 0), and one IADD3 subtracts the words of TWO pairs from a
        // packed accumulator: 3 ALU-pipe instructions per 4 evaluations instead of 4 LEA.HI. acc = (ny - nx) 2^16 + nx.
#pragma unroll
        for (int pp = 0; pp < SUB; pp += 2) {
          unsigned int tw[2][HPT];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 a = sp[(sub + pp + h) * 3 + 0];
            const float4 b = sp[(sub + pp + h) * 3 + 1];
            const float4 c = sp[(sub + pp + h) * 3 + 2];
#pragma unroll
            for (int k = 0; k < HPT; ++k) {
              const float2 s = hyp[k].eval(a, b, c, nlo);
              tw[h][k] = sign_words(s);
              smin[k] = fminf(fminf(smin[k], fabsf(s.x)), fabsf(s.y));
            }
          }
#pragma unroll
          for (int k = 0; k < HPT; ++k) pacc[k] = pacc[k] - tw[0][k] - tw[1][k];
        }
      } else {
#pragma unroll
      for (int pp = 0; pp < SUB; ++pp) {
        const float4 a = sp[(sub + pp) * 3 + 0];
        const float4 b = sp[(sub + pp) * 3 + 1];
        const float4 c = sp[(sub + pp) * 3 + 2];
#pragma unroll
        for (int k = 0; k < HPT; ++k) {
          const float2 s = hyp[k].eval(a, b, c, nlo);
          cnt[k] += (int)(__float_as_uint(s.x) >> 31) + (int)(__float_as_uint(s.y) >> 31);
          smin[k] = fminf(fminf(smin[k], fabsf(s.x)), fabsf(s.y));
        }
      }
      }
      bool any = false;
#pragma unroll
      for (int k = 0; k < HPT; ++k) any = any || (smin[k] <= band[k]);
      if (any) {
        for (int pp = 0; pp < SUB; ++pp) {
          const float4 a = sp[(sub + pp) * 3 + 0];
          const float4 b = sp[(sub + pp) * 3 + 1];
          const float4 c = sp[(sub + pp) * 3 + 2];
#pragma unroll
          for (int k = 0; k < HPT; ++k) {
            const float2 s = hyp[k].eval(a, b, c, nlo);
            const float sv[2] = {s.x, s.y};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (fabsf(sv[u]) <= band[k]) {
                cnt[k] -= (int)(__float_as_uint(sv[u]) >> 31);  // (the packed accumulator is folded into cnt per stage)
                const unsigned int corr = (unsigned int)(corr_base + 2 * (p_begin + t * TILE + sub + pp) + u);
                seg.push(make_uint2((unsigned int)slot[k], corr | (1u << 30)), st);
              }
            }
          }
        }
      }
    }
    if (PCOUNT) {  // fold the packed fields (at most TILE <= 1024 pairs since the last fold: no field overflows)
#pragma unroll
      for (int k = 0; k < HPT; ++k) {
        const unsigned int lo = pacc[k] & 0xffffu, hi = pacc[k] >> 16;
        cnt[k] += (int)(lo + ((hi + lo) & 0xffffu));
        pacc[k] = 0u;
      }
    }
    __syncthreads();  // everybody is done with the records (and has read the bound) before the next transpose
  }
#pragma unroll
  for (int k = 0; k < HPT; ++k)
    if (slot[k] >= 0 && cnt[k] != 0) atomicAdd(&votes[slot[k]], cnt[k]);
  seg.publish(wl, st);
}

// ================================================================================================
// 3-D / 3-D scorer that STREAMS the frame from page-locked host memory and leaves a device copy behind
// ================================================================================================
// Bulk TMA reads mapped host memory like any other global memory (measured: 53 GB/s, more than the copy engine gets
// for the same arrays), so a frame that lives in page-locked host memory needs no upload before it is scored: the CTA
// asks for its whole slice at once, in sub-stages of SUBT pairs with one mbarrier each, and transposes + scores a
// sub-stage as soon as it has landed while the rest is still on the bus. Each landed sub-stage is also written to the
// context's device arrays with a bulk store (shared -> global), so that the fix-up, mask and refit kernels find the
// frame in HBM afterwards. Arithmetic, guard bands (per sub-stage bound) and vote counts are those of score3d_raw_kernel.
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int TILE, int THREADS, int SUB, int SUBT>
__global__ void __launch_bounds__(THREADS, 1)
score3d_stream_kernel(const float* __restrict__ xw, const float* __restrict__ xc, float* __restrict__ dxw, float* __restrict__ dxc,
                      int n, int npairs_pad, int pairs_per_cta, const HypFast* __restrict__ fast,
                      const HypGen* __restrict__ gen, int slot_begin, int slot_end, float thr, int32_t* __restrict__ votes,
                      FrameStats* __restrict__ st, Worklist wl) {
  constexpr int HPT = 2;
  constexpr int NSUB = TILE / SUBT;
  static_assert(TILE % SUBT == 0 && SUBT % SUB == 0 && SUBT % 2 == 0, "sub-stages are whole groups");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kStageFloats = TILE * 2 * 3;
  float4* packed = reinterpret_cast<float4*>(smem_raw);                                  // [TILE * 3]
  float* raw = reinterpret_cast<float*>(smem_raw + (size_t)TILE * 3 * sizeof(float4));    // [2 stages][xw | xc][kStageFloats]
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw + 4 * kStageFloats);                  // [2 stages][NSUB]
  unsigned int* mtile = reinterpret_cast<unsigned int*>(bars + 2 * NSUB);                // [2 stages][NSUB]

  const int tid = threadIdx.x;
  const int p_begin = blockIdx.x * pairs_per_cta;
  const int p_end = min(p_begin + pairs_per_cta, npairs_pad);
  const int npairs = p_end - p_begin;
  const int ntiles = (npairs + TILE - 1) / TILE;
  WlSegment seg(wl);

  HypRegs<true> hyp[HPT];
  int slot[HPT];
  int cnt[HPT];
  unsigned int pacc[HPT];
  float tnorm[HPT];
#pragma unroll
  for (int k = 0; k < HPT; ++k) {
    slot[k] = slot_begin + (blockIdx.y * HPT + k) * THREADS + tid;
    const bool live = slot[k] < slot_end && gen[slot[k]].valid != 0;
    hyp[k].load(&fast[live ? slot[k] : slot_begin], live);
    if (!live) slot[k] = -1;
    cnt[k] = 0;
    pacc[k] = 0u;
    tnorm[k] = sqrtf(hyp[k].nt[0].x * hyp[k].nt[0].x + hyp[k].nt[1].x * hyp[k].nt[1].x + hyp[k].nt[2].x * hyp[k].nt[2].x);
  }
  const float thr2 = __fmul_rn(thr, thr);
  const float2 nlo = make_float2(-thr2, -thr2);

  auto stage_range = [&](int t, int& c0, int& cnt4) {
    c0 = 2 * (p_begin + t * TILE);
    const int tp = min(TILE, npairs - t * TILE);
    int c = min(2 * tp, n - c0);
    if (c < 0) c = 0;
    cnt4 = c & ~3;
  };
  // correspondences [c0 + 2 SUBT sub, ...) of sub-stage `sub`: how many go through TMA
  auto sub_count = [&](int cnt4, int sub) {
    int c = cnt4 - 2 * SUBT * sub;
    return c < 0 ? 0 : (c > 2 * SUBT ? 2 * SUBT : c);
  };
  auto issue = [&](int t, int buf) {
    int c0, cnt4;
    stage_range(t, c0, cnt4);
#pragma unroll 1
    for (int sub = 0; sub < NSUB; ++sub) {
      const uint32_t bytes = (uint32_t)sub_count(cnt4, sub) * 12u;
      uint64_t* bar = &bars[buf * NSUB + sub];
      mbar_expect_tx(bar, 2u * bytes);  // 0 bytes: the phase completes on this arrival alone
      if (bytes) {
        const size_t off = (size_t)sub * 2 * SUBT * 3;
        tma_load_1d(raw + (size_t)(2 * buf) * kStageFloats + off, xw + (size_t)c0 * 3 + off, bytes, bar);
        tma_load_1d(raw + (size_t)(2 * buf + 1) * kStageFloats + off, xc + (size_t)c0 * 3 + off, bytes, bar);
      }
    }
  };
  if (tid == 0) {
    for (int i = 0; i < 2 * NSUB; ++i) {
      mbar_init(&bars[i], 1);
      mtile[i] = 0u;
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int t = 0; t < 2 && t < ntiles; ++t) issue(t, t);

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    const uint32_t parity = (uint32_t)((t >> 1) & 1);
    const int tp = min(TILE, npairs - t * TILE);
    int c0, cnt4;
    stage_range(t, c0, cnt4);
    const float* rw = raw + (size_t)(2 * buf) * kStageFloats;
    const float* rc = raw + (size_t)(2 * buf + 1) * kStageFloats;
    const int nsub = (tp + SUBT - 1) / SUBT;
    for (int sub = 0; sub < nsub; ++sub) {
      mbar_wait(&bars[buf * NSUB + sub], parity);
      if (tid == 0 && dxw) {  // the device copy of what has just landed
        const uint32_t bytes = (uint32_t)sub_count(cnt4, sub) * 12u;
        if (bytes) {
          const size_t off = (size_t)sub * 2 * SUBT * 3;
          tma_store_1d(dxw + (size_t)c0 * 3 + off, rw + off, bytes);
          tma_store_1d(dxc + (size_t)c0 * 3 + off, rc + off, bytes);
          tma_store_commit();
        }
      }
      const int pr_lo = sub * SUBT, pr_hi = min(tp, pr_lo + SUBT);
      // ---- transpose the sub-stage: thread <-> 4 correspondences = 2 pair records; its magnitude bound
      float mloc = 0.f;
      for (int qd = pr_lo / 2 + tid; qd < (pr_hi + 1) / 2; qd += THREADS) {
        float w[12], c[12];
        if (4 * qd + 4 <= cnt4) {
          const float4* w4 = reinterpret_cast<const float4*>(rw + 12 * qd);
          const float4* c4 = reinterpret_cast<const float4*>(rc + 12 * qd);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float4 a = w4[i], b = c4[i];
            w[4 * i] = a.x; w[4 * i + 1] = a.y; w[4 * i + 2] = a.z; w[4 * i + 3] = a.w;
            c[4 * i] = b.x; c[4 * i + 1] = b.y; c[4 * i + 2] = b.z; c[4 * i + 3] = b.w;
          }
        } else {  // frame tail (or padding): element-wise from the source arrays, NaN beyond the last correspondence
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ci = c0 + 4 * qd + j;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              w[3 * j + r] = ci < n ? xw[(size_t)ci * 3 + r] : CUDART_NAN_F;
              c[3 * j + r] = ci < n ? xc[(size_t)ci * 3 + r] : CUDART_NAN_F;
              if (dxw && ci < n) {
                dxw[(size_t)ci * 3 + r] = w[3 * j + r];
                dxc[(size_t)ci * 3 + r] = c[3 * j + r];
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float m = sqrtf(w[3 * j] * w[3 * j] + w[3 * j + 1] * w[3 * j + 1] + w[3 * j + 2] * w[3 * j + 2]);
          const float mc = sqrtf(c[3 * j] * c[3 * j] + c[3 * j + 1] * c[3 * j + 1] + c[3 * j + 2] * c[3 * j + 2]);
          if (mc == mc && mc < CUDART_INF_F) m += mc;
          if (m == m && m < CUDART_INF_F) mloc = fmaxf(mloc, m);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int pr = 2 * qd + h;
          if (pr < tp) {
            const float* a = w + 6 * h;
            const float* b = c + 6 * h;
            packed[pr * 3 + 0] = make_float4(a[0], a[3], a[1], a[4]);
            packed[pr * 3 + 1] = make_float4(a[2], a[5], b[0], b[3]);
            packed[pr * 3 + 2] = make_float4(b[1], b[4], b[2], b[5]);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, o));
      if ((tid & 31) == 0 && mloc > 0.f) atomicMax(&mtile[buf * NSUB + sub], __float_as_uint(mloc));
      __syncthreads();  // records + bound of this sub-stage complete
      const float mcorr = __uint_as_float(mtile[buf * NSUB + sub]);
      float band[HPT];
#pragma unroll
      for (int k = 0; k < HPT; ++k) band[k] = guard_band_3d((mcorr + tnorm[k]) * 1.0001f, thr);  // NaN for a dead slot

      const float4* sp = packed;
      for (int g0 = pr_lo; g0 < pr_hi; g0 += SUB) {
        float smin[HPT];
#pragma unroll
        for (int k = 0; k < HPT; ++k) smin[k] = CUDART_INF_F;
#pragma unroll
        for (int pp = 0; pp < SUB; pp += 2) {
          unsigned int tw[2][HPT];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 a = sp[(g0 + pp + h) * 3 + 0];
            const float4 b = sp[(g0 + pp + h) * 3 + 1];
            const float4 c = sp[(g0 + pp + h) * 3 + 2];
#pragma unroll
            for (int k = 0; k < HPT; ++k) {
              const float2 s = hyp[k].eval(a, b, c, nlo);
              tw[h][k] = sign_words(s);
              smin[k] = fminf(fminf(smin[k], fabsf(s.x)), fabsf(s.y));
            }
          }
#pragma unroll
          for (int k = 0; k < HPT; ++k) pacc[k] = pacc[k] - tw[0][k] - tw[1][k];
        }
        bool any = false;
#pragma unroll
        for (int k = 0; k < HPT; ++k) any = any || (smin[k] <= band[k]);
        if (any) {
          for (int pp = 0; pp < SUB; ++pp) {
            const float4 a = sp[(g0 + pp) * 3 + 0];
            const float4 b = sp[(g0 + pp) * 3 + 1];
            const float4 c = sp[(g0 + pp) * 3 + 2];
#pragma unroll
            for (int k = 0; k < HPT; ++k) {
              const float2 s = hyp[k].eval(a, b, c, nlo);
              const float sv[2] = {s.x, s.y};
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                if (fabsf(sv[u]) <= band[k]) {
                  cnt[k] -= (int)(__float_as_uint(sv[u]) >> 31);
                  const unsigned int corr = (unsigned int)(2 * (p_begin + t * TILE + g0 + pp) + u);
                  seg.push(make_uint2((unsigned int)slot[k], corr | (1u << 30)), st);
                }
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < HPT; ++k) {  // fold the packed fields (at most TILE <= 1024 pairs since the last fold)
      const unsigned int lo = pacc[k] & 0xffffu, hi = pacc[k] >> 16;
      cnt[k] += (int)(lo + ((hi + lo) & 0xffffu));
      pacc[k] = 0u;
    }
    __syncthreads();  // everybody is done with the records and the bounds of this stage
    if (tid == 0) {
      for (int i = 0; i < NSUB; ++i) mtile[buf * NSUB + i] = 0u;
      if (t + 2 < ntiles) {
        tma_store_wait_read();  // the bulk stores have read raw[buf]
        issue(t + 2, buf);
      }
    }
  }
  if (tid == 0) tma_store_wait_all();
#pragma unroll
  for (int k = 0; k < HPT; ++k)
    if (slot[k] >= 0 && cnt[k] != 0) atomicAdd(&votes[slot[k]], cnt[k]);
  seg.publish(wl, st);
}

// host side of the streaming scorer: the full 1024-hypothesis column shape of the default variant
int launch_score3d_stream(const FrameView& fsrc, float* dxw, float* dxc, const HypGen* gen, const HypFast* fast, int slot_begin,
                          int slot_end, float thr3d, int32_t* votes, FrameStats* st, Worklist wl, int num_sms, cudaStream_t s) {
  constexpr int TILE = 1024, THREADS = 512, SUB = 4, SUBT = 128, HPT = 2;
  const int nslots = slot_end - slot_begin;
  if (nslots <= 0 || fsrc.n <= 0) return 0;
  const int gy = (nslots + THREADS * HPT - 1) / (THREADS * HPT);
  const int groups = fsrc.npairs_pad / SUB;
  int gx = num_sms / gy;
  if (gx < 1) gx = 1;
  if (gx > groups) gx = groups;
  const int groups_per_cta = (groups + gx - 1) / gx;
  const int pairs_per_cta = groups_per_cta * SUB;
  gx = (fsrc.npairs_pad + pairs_per_cta - 1) / pairs_per_cta;
  const size_t smem = (size_t)TILE * 3 * sizeof(float4) + 4 * (size_t)TILE * 6 * sizeof(float) + 2 * (TILE / SUBT) * (sizeof(uint64_t) + sizeof(unsigned int)) + 16;
  auto k = score3d_stream_kernel<TILE, THREADS, SUB, SUBT>;
  static std::atomic<bool> attr[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr[dev].load(std::memory_order_acquire)) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr[dev].store(true, std::memory_order_release);
  }
  // (gy > 1 would stream the frame once per hypothesis column: the caller only uses this for H * S <= 1024)
  k<<<dim3(gx, gy), THREADS, smem, s>>>(fsrc.xw, fsrc.xc, gy == 1 ? dxw : nullptr, gy == 1 ? dxc : nullptr, fsrc.n, fsrc.npairs_pad,
                                       pairs_per_cta, fast, gen, slot_begin, slot_end, thr3d, votes, st, wl);
  return gx * gy;
}

// ================================================================================================
// fast tiled scorer — 2-D, 3-D and normal modalities in any combination (all other estimator families)
// ================================================================================================
// Shared per (pair, hypothesis): ny = -(R x_w + t) = nR x_w + nt (9 FFMA2). Then
//   3-D   : e = x_c + ny, s3 = |e|^2 - thr^2                      (3 FADD2 + 3 FFMA2)    inlier <=> s3 < 0
//   2-D   : d' = ny.b, n2 = |ny|^2, F' = d'|d'| + cos_thr^2 n2     (6 FP2 + 2x(FMUL,FFMA))  inlier <=> F' < 0
//           (F' < 0 <=> y.b > 0 and (y.b)^2 > cos_thr^2 |y|^2 <=> normalize(y).b > cos_thr, cos_thr > 0 always: P3P.hpp:323)
//   normal: g' = cos_nl - n_c.(R n_w) = n_c.(nR n_w) + cos_nl     (3 FMUL2 + 6 FFMA2 + 3 FFMA2)  inlier <=> g' < 0
// Guard bands (same rigorous scheme as §4.2 of DESIGN.md; u = 2^-24, M2 >= |x_w| + |t|, Nn >= |n_c||n_w|):
//   2-D   : |F'| <= cos_thr * u * 1.1 * (32 M2 (1 + n2) + 26 n2)   (reference error 19.1u M2/|y| + 7u, fast 12.5u M2/|y| + 5.5u in
//           cosine units, times 2 cos_thr |y|^2, |y| <= (1 + n2)/2)
//   normal: |g'| <= u * 1.1 * (29 Nn + 2)                          (reference 19.1u Nn, fast 9.2u Nn + u)
// Packed layout per pair: x_w, [x_c], [b], [n_w, n_c], 6 floats each (see pack_kernel), padded to whole float4.
template <int KIND>
struct KindTraits {
  static constexpr bool k2 = (KIND & 1) != 0, k3 = (KIND & 2) != 0, kn = (KIND & 4) != 0;
  static constexpr int arrays = 1 + (k2 ? 1 : 0) + (k3 ? 1 : 0) + (kn ? 2 : 0);
  static constexpr int f4pp = (arrays * 6 + 3) / 4;
  // RAW records carry, instead of n_w and n_c, the nine products n_c,i n_w,j of each correspondence (18 floats per pair,
  // formed once per stage while transposing): the normal test is then 9 FFMA2 per (pair, hypothesis) instead of 12
  static constexpr int f4raw = kn ? ((arrays - 2) * 6 + 18 + 3) / 4 : f4pp;
  static constexpr int off_xc = 6;                       // floats
  static constexpr int off_bv = 6 + (k3 ? 6 : 0);
  static constexpr int off_nw = 6 + (k3 ? 6 : 0) + (k2 ? 6 : 0);
  static constexpr int off_nc = off_nw + 6;
};

struct BandConsts {   // per hypothesis
  float band3;            // on s3
  float k1_2d, k2_2d, c0_2d;  // band2 = -k1 d' + k2 n2 + c0 (k1_2d holds -k1)
  float band_n;           // on g'
};

// Where a scoring CTA gets its correspondences from: the packed pair records (pack_kernel) or the caller's own
// arrays (RAW: bulk TMA of the raw xyz triples of a stage + transpose in shared memory, no packed copy in HBM).
struct MultiSrc {
  const float4* pk;   // packed records (RAW == false)
  const float* a[5];  // RAW: the record's source arrays in record order: x_w, [x_c], [b], [n_w, n_c]
  const float* xc;    // RAW: camera points for the isValid gate of the normal test (may equal a[1])
  int n;
  int unit_entries;   // DIRECT: one worklist entry per borderline (pair, hypothesis) unit (needs n <= 2^25), else one per value
};

template <int KIND, int TILE, int THREADS, bool DIRECT, bool RAW>
__global__ void __launch_bounds__(THREADS, 1)
score_multi_fast_kernel(MultiSrc src, int npairs_pad, int pairs_per_cta, const HypFast* __restrict__ fast,
                        const HypGen* __restrict__ gen, int slot_begin, int slot_end, Thresh th,
                        int32_t* __restrict__ votes, FrameStats* __restrict__ st, Worklist wl) {
  typedef KindTraits<KIND> KT;
  constexpr int HPT = 2;
  constexpr int SUB = kSubPairs;
  constexpr int F4 = RAW ? KT::f4raw : KT::f4pp;
  constexpr bool kBandAbs = KIND == 1;  // 2-D band from |d'| (2-D-only kind) or from -d' (hybrids), see set_bands
  // RAW: arrays staged per stage = the record's arrays, plus x_c when the kind has the normal test but not the 3-D one
  constexpr int NREC = KT::arrays;
  constexpr bool XC_EXTRA = KT::kn && !KT::k3;
  constexpr int NRAW = NREC + (XC_EXTRA ? 1 : 0);
  constexpr int kStageFloats = TILE * 2 * 3;  // floats of one array in one stage
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // packed: [2 stages][TILE * F4] float4 | bars.   RAW: [TILE * F4] float4 records | [2 stages][NRAW][kStageFloats] | bars | bounds
  float4(*tile)[TILE * F4] = reinterpret_cast<float4(*)[TILE * F4]>(smem_raw);
  float* rawbuf = reinterpret_cast<float*>(smem_raw + (size_t)TILE * F4 * sizeof(float4));
  uint64_t* bars = RAW ? reinterpret_cast<uint64_t*>(rawbuf + (size_t)2 * NRAW * kStageFloats)
                       : reinterpret_cast<uint64_t*>(smem_raw + 2 * TILE * F4 * sizeof(float4));
  unsigned int* mstage = reinterpret_cast<unsigned int*>(bars + 2);  // RAW: [2 stages][3] = max|x_w|, max(|x_w|+|x_c|), max normal
  const float4* pk = src.pk;

  const int tid = threadIdx.x;
  const bool unit_entries = src.unit_entries != 0;
  const int p_begin = blockIdx.x * pairs_per_cta;
  const int p_end = min(p_begin + pairs_per_cta, npairs_pad);
  const int npairs = p_end - p_begin;
  const int ntiles = (npairs + TILE - 1) / TILE;
  WlSegment seg(wl);

  float nR[HPT][9], nt[HPT][3];
  int slot[HPT], cnt[HPT];
#pragma unroll
  for (int k = 0; k < HPT; ++k) {
    slot[k] = slot_begin + (blockIdx.y * HPT + k) * THREADS + tid;
    const bool live = slot[k] < slot_end && gen[slot[k]].valid != 0;
    const HypFast* h = &fast[live ? slot[k] : slot_begin];
#pragma unroll
    for (int i = 0; i < 9; ++i) nR[k][i] = live ? h->nR[i] : CUDART_NAN_F;
#pragma unroll
    for (int i = 0; i < 3; ++i) nt[k][i] = live ? h->nt[i] : CUDART_NAN_F;
    if (!live) slot[k] = -1;
    cnt[k] = 0;
  }
  // Guard bands from the frame's magnitude bound and THIS hypothesis' |t| (M >= |x_w| + |x_c| + |t|).
  // 2-D: the rigorous band is beta(|y|) = c u 1.1 (64 M |y| + 26 |y|^2). |y| needs no square root:
  //   * if |F'| <= beta <= 0.19 c^2 n2 then d'^2 >= 0.81 c^2 n2, i.e. |y| <= 1.112 |d'| / c, hence
  //     beta <= u 1.1 (71.2 M |d'| + 26 c n2);
  //   * otherwise |y| < y0 = 375 u M / c (a point within ~1e-5 M of the camera centre) and beta <= beta(y0) =: c0.
  // band2 = k1 |d'| + k2 n2 + c0 therefore covers beta in both regimes.
  BandConsts bc[HPT];
  float tnorm[HPT];
#pragma unroll
  for (int k = 0; k < HPT; ++k) tnorm[k] = sqrtf(nt[k][0] * nt[k][0] + nt[k][1] * nt[k][1] + nt[k][2] * nt[k][2]);  // NaN for a dead slot
  // mw >= |x_w|, mwc >= |x_w| + |x_c|, nmax >= max(|n_w|, |n_c|) over the correspondences the bands are applied to:
  // the frame (packed records, bounds taken by the pack kernel) or the current stage (RAW, taken while transposing)
  auto set_bands = [&](float mw, float mwc, float nmax) {
#pragma unroll
    for (int k = 0; k < HPT; ++k) {
      const float u = 5.9604644775390625e-08f;
      const float M3 = (mwc + tnorm[k]) * 1.0001f;  // NaN for a dead slot: never borderline
      const float M2 = (mw + tnorm[k]) * 1.0001f;
      bc[k].band3 = guard_band_3d(M3, th.thr3d);
      // stored NEGATED: the band uses -d' instead of |d'| (one FFMA2, no absolute values). For d' < 0 that is the same
      // number. For d' >= 0, F' = d'^2 + c^2 n2 >= c^2 n2, so |F'| <= beta(|y|) forces |y| <= y1 = 70.5 u M2 / c (< y0), and
      // the band  c0 + k2 n2 - k1 d' >= c0 - k1 y1  must still cover beta(y1): c0 >= beta(y1) + k1 y1 =
      // 1.1 u M2 y1 (64 c + 71.2) <= 1.1 * 9532 u^2 M2^2 / c, while beta(y0) = 1.1 * 24000 u^2 M2^2: (2 + 1/c) beta(y0) covers it
      // for every cos_thr > 0.
      // (the 2-D-only kind keeps |d'|: measured 3 % faster there, 4-7 % slower for the hybrids — round 2, r02d)
      bc[k].k1_2d = kBandAbs ? u * 1.1f * 71.2f * M2 : -(u * 1.1f * 71.2f * M2);
      bc[k].k2_2d = u * 1.1f * 26.f * th.cos_thr;
      const float y0 = 375.f * u * M2 / th.cos_thr;
      bc[k].c0_2d = (kBandAbs ? 1.f : 2.f + 1.f / th.cos_thr) * th.cos_thr * u * 1.1f * (64.f * M2 * y0 + 26.f * y0 * y0);
      const float nm = nmax * 1.0001f;
      // packed records: reference 19.1 u N^2, fast (3 FMUL2 + 9 FFMA2) 9.2 u N^2 + u.
      // RAW records (products p_ij = fl(n_c,i n_w,j), then one 9-term FFMA chain from cos_nl): product and nR roundings
      // 2 x 1.74 u N^2 (|| |R| ||_2 <= sqrt 3), chain 9 u (1 + 1.74 N^2)  ->  fast <= u (19.1 N^2 + 9)
      bc[k].band_n = RAW ? u * 1.1f * (40.f * nm * nm + 12.f) : u * 1.1f * (29.f * nm * nm + 2.f);
    }
  };
  if (!RAW) {
    const float mc = __uint_as_float(st->m_corr_bits);  // the pack kernel's frame bound is |x_w| + |x_c| (or |x_w| alone)
    set_bands(mc, mc, __uint_as_float(st->m_bv_bits));
  }
  const float thr2 = __fmul_rn(th.thr3d, th.thr3d);
  const float2 nlo = make_float2(-thr2, -thr2);
  const float c2 = (float)((double)th.cos_thr * (double)th.cos_thr);
  const float2 cnl2 = make_float2(th.cos_nl, th.cos_nl);
  // RAW: correspondences [c0, c0 + cnt4) of stage t go through TMA (cnt4 a multiple of 4, possibly 0); the frame's last
  // 0..3 correspondences are read from global memory by the transposing threads
  const int n = src.n;
  auto stage_range = [&](int t, int& c0, int& cnt4) {
    c0 = 2 * (p_begin + t * TILE);
    const int tp = min(TILE, npairs - t * TILE);
    int c = min(2 * tp, n - c0);
    if (c < 0) c = 0;
    cnt4 = c & ~3;
  };
  auto issue_raw = [&](int t, int buf) {
    int c0, cnt4;
    stage_range(t, c0, cnt4);
    const uint32_t bytes = (uint32_t)cnt4 * 12u;
    mbar_expect_tx(&bars[buf], (uint32_t)NRAW * bytes);  // 0 bytes: the phase completes on this arrival alone
    if (bytes) {
#pragma unroll
      for (int a = 0; a < NREC; ++a)
        tma_load_1d(rawbuf + (size_t)(buf * NRAW + a) * kStageFloats, src.a[a] + (size_t)c0 * 3, bytes, &bars[buf]);
      if (XC_EXTRA) tma_load_1d(rawbuf + (size_t)(buf * NRAW + NREC) * kStageFloats, src.xc + (size_t)c0 * 3, bytes, &bars[buf]);
    }
  };
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
    if (RAW)
      for (int i = 0; i < 6; ++i) mstage[i] = 0u;
  }
  __syncthreads();
  if (tid == 0) {
    for (int t = 0; t < 2 && t < ntiles; ++t) {
      if (RAW) {
        issue_raw(t, t);
      } else {
        const int tp = min(TILE, npairs - t * TILE);
        const uint32_t bytes = (uint32_t)tp * F4 * 16u;
        mbar_expect_tx(&bars[t], bytes);
        tma_load_1d(&tile[t][0], pk + (size_t)(p_begin + t * TILE) * F4, bytes, &bars[t]);
      }
    }
  }

  int xtra = 0;  // borderline evaluations beyond one per queued entry (unit entries)
  // One (pair, hypothesis) unit.
  // DIRECT (default): a borderline evaluation is queued for the exact fix-up on the spot instead of being counted —
  //   frames with many near-threshold evaluations (dense frames, low outlier ratios, pixel-level 2-D thresholds) would
  //   otherwise re-walk almost every group;
  // otherwise: rescan == false counts sign bits and ORs the band flags of a 16-correspondence group, rescan == true
  //   re-walks a flagged group, takes the borderline evaluations out of the count and queues them.
  auto unit = [&](const float* rec, int k, bool rescan, bool& flag, int corr0) {
    const float2 X0 = make_float2(rec[0], rec[1]), X1 = make_float2(rec[2], rec[3]), X2 = make_float2(rec[4], rec[5]);
    float2 ny0 = make_float2(nt[k][0], nt[k][0]), ny1 = make_float2(nt[k][1], nt[k][1]), ny2 = make_float2(nt[k][2], nt[k][2]);
    ny0 = __ffma2_rn(make_float2(nR[k][2], nR[k][2]), X2, ny0);
    ny1 = __ffma2_rn(make_float2(nR[k][5], nR[k][5]), X2, ny1);
    ny2 = __ffma2_rn(make_float2(nR[k][8], nR[k][8]), X2, ny2);
    ny0 = __ffma2_rn(make_float2(nR[k][1], nR[k][1]), X1, ny0);
    ny1 = __ffma2_rn(make_float2(nR[k][4], nR[k][4]), X1, ny1);
    ny2 = __ffma2_rn(make_float2(nR[k][7], nR[k][7]), X1, ny2);
    ny0 = __ffma2_rn(make_float2(nR[k][0], nR[k][0]), X0, ny0);
    ny1 = __ffma2_rn(make_float2(nR[k][3], nR[k][3]), X0, ny1);
    ny2 = __ffma2_rn(make_float2(nR[k][6], nR[k][6]), X0, ny2);
    float val[3][2];   // per modality (2-D, 3-D, normal) the decision value of both correspondences
    float bnd[3][2];
    if (KT::k3) {
      const float* p = rec + KT::off_xc;
      const float2 e0 = __fadd2_rn(make_float2(p[0], p[1]), ny0);
      const float2 e1 = __fadd2_rn(make_float2(p[2], p[3]), ny1);
      const float2 e2 = __fadd2_rn(make_float2(p[4], p[5]), ny2);
      float2 s = __ffma2_rn(e0, e0, nlo);
      s = __ffma2_rn(e1, e1, s);
      s = __ffma2_rn(e2, e2, s);
      val[1][0] = s.x;
      val[1][1] = s.y;
      bnd[1][0] = bnd[1][1] = bc[k].band3;
    }
    if (KT::k2) {
      const float* b = rec + KT::off_bv;
      float2 d = __fmul2_rn(ny0, make_float2(b[0], b[1]));
      d = __ffma2_rn(ny1, make_float2(b[2], b[3]), d);
      d = __ffma2_rn(ny2, make_float2(b[4], b[5]), d);
      float2 n2 = __fmul2_rn(ny0, ny0);
      n2 = __ffma2_rn(ny1, ny1, n2);
      n2 = __ffma2_rn(ny2, ny2, n2);
      const float2 dabs = make_float2(fabsf(d.x), fabsf(d.y));  // two LOP3 on the ALU pipe, which has room
      float2 band2 = __ffma2_rn(make_float2(bc[k].k2_2d, bc[k].k2_2d), n2, make_float2(bc[k].c0_2d, bc[k].c0_2d));
      band2 = __ffma2_rn(make_float2(bc[k].k1_2d, bc[k].k1_2d), kBandAbs ? dabs : d,
                         band2);  // hybrids: k1 is stored negated, -k1 d' (see set_bands)
      // F' = d'|d'| + cos^2 |y|^2 of both correspondences as two packed instructions (same products, same single rounding
      // of the sum as the scalar fmaf(c2, n2, d * |d|) they replace: four FMA-pipe cycles instead of eight)
      // (measured: 1-2 % for shinji_kneip, nothing for the three-modality kind, 1 % slower for the 2-D-only kind, which keeps the scalar form)
      if (KIND == 1) {
        val[0][0] = fmaf(c2, n2.x, d.x * fabsf(d.x));
        val[0][1] = fmaf(c2, n2.y, d.y * fabsf(d.y));
      } else {
        const float2 fv = __ffma2_rn(make_float2(c2, c2), n2, __fmul2_rn(d, dabs));
        val[0][0] = fv.x;
        val[0][1] = fv.y;
      }
      bnd[0][0] = band2.x;
      bnd[0][1] = band2.y;
    }
    if (KT::kn && RAW) {
      const float* pr = rec + KT::off_nw;  // p_ij of both correspondences, (i, j) row-major, interleaved
      float2 g = cnl2;
#pragma unroll
      for (int e = 0; e < 9; ++e) g = __ffma2_rn(make_float2(nR[k][e], nR[k][e]), make_float2(pr[2 * e], pr[2 * e + 1]), g);
      val[2][0] = g.x;
      val[2][1] = g.y;
      bnd[2][0] = bnd[2][1] = bc[k].band_n;
    }
    if (KT::kn && !RAW) {
      const float* w = rec + KT::off_nw;
      const float* c = rec + KT::off_nc;
      const float2 W0 = make_float2(w[0], w[1]), W1 = make_float2(w[2], w[3]), W2 = make_float2(w[4], w[5]);
      float2 m0 = __fmul2_rn(make_float2(nR[k][0], nR[k][0]), W0);
      float2 m1 = __fmul2_rn(make_float2(nR[k][3], nR[k][3]), W0);
      float2 m2 = __fmul2_rn(make_float2(nR[k][6], nR[k][6]), W0);
      m0 = __ffma2_rn(make_float2(nR[k][1], nR[k][1]), W1, m0);
      m1 = __ffma2_rn(make_float2(nR[k][4], nR[k][4]), W1, m1);
      m2 = __ffma2_rn(make_float2(nR[k][7], nR[k][7]), W1, m2);
      m0 = __ffma2_rn(make_float2(nR[k][2], nR[k][2]), W2, m0);
      m1 = __ffma2_rn(make_float2(nR[k][5], nR[k][5]), W2, m1);
      m2 = __ffma2_rn(make_float2(nR[k][8], nR[k][8]), W2, m2);
      float2 g = __ffma2_rn(make_float2(c[0], c[1]), m0, cnl2);
      g = __ffma2_rn(make_float2(c[2], c[3]), m1, g);
      g = __ffma2_rn(make_float2(c[4], c[5]), m2, g);
      val[2][0] = g.x;
      val[2][1] = g.y;
      bnd[2][0] = bnd[2][1] = bc[k].band_n;
    }
    if (DIRECT) {
      // Straight-line common case: every sign bit is counted and the band tests are folded into one predicate;
      // only a unit that holds a borderline value branches, takes that value out of the count and queues it.
      bool any = false;
#pragma unroll
      for (int mod = 0; mod < 3; ++mod) {
        if ((mod == 0 && !KT::k2) || (mod == 1 && !KT::k3) || (mod == 2 && !KT::kn)) continue;
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
          cnt[k] += (int)(__float_as_uint(val[mod][uu]) >> 31);
          any = any || (fabsf(val[mod][uu]) <= bnd[mod][uu]);
        }
      }
      // (Two warp-cooperative variants of this slow path — a warp-uniform vote + ballots into per-warp sub-segments, and
      // ballot aggregation among the branching lanes with a shared-memory warp counter — were measured in round 2 and
      // were slower for the 2-D kinds: the vote sits in the fast path, the counter adds two shared-memory round trips.)
      if (KIND != 1 && any && unit_entries) {  // (compiled out for the 2-D-only kind: at most two values per unit, no gain)
        // ONE entry for the unit: (slot, pair index | six borderline bits << 24 | 3 << 30), bit 2 mod + uu. The fix-up kernel
        // expands it. Half the instructions of the per-value form below, and the branch is taken by a quarter of the
        // warp-iterations of a dense frame at 30 % outliers (a percent of a good hypothesis' 2-D evaluations is borderline).
        unsigned int bits = 0u;
#pragma unroll
        for (int mod = 0; mod < 3; ++mod) {
          if ((mod == 0 && !KT::k2) || (mod == 1 && !KT::k3) || (mod == 2 && !KT::kn)) continue;
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {
            const float v = val[mod][uu];
            if (fabsf(v) <= bnd[mod][uu]) {
              bits |= 1u << (2 * mod + uu);
              cnt[k] -= (int)(__float_as_uint(v) >> 31);
            }
          }
        }
        xtra += __popc(bits) - 1;
        seg.push(make_uint2((unsigned int)slot[k], ((unsigned int)corr0 >> 1) | (bits << 24) | (3u << 30)), st);
        return;
      }
      if (any) {  // one reservation for all borderline values of this thread's unit
        unsigned int nb = 0;
#pragma unroll
        for (int mod = 0; mod < 3; ++mod) {
          if ((mod == 0 && !KT::k2) || (mod == 1 && !KT::k3) || (mod == 2 && !KT::kn)) continue;
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) nb += (fabsf(val[mod][uu]) <= bnd[mod][uu]) ? 1u : 0u;
        }
        unsigned int wi = seg.reserve(nb);
#pragma unroll
        for (int mod = 0; mod < 3; ++mod) {
          if ((mod == 0 && !KT::k2) || (mod == 1 && !KT::k3) || (mod == 2 && !KT::kn)) continue;
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {
            const float v = val[mod][uu];
            if (fabsf(v) <= bnd[mod][uu]) {
              cnt[k] -= (int)(__float_as_uint(v) >> 31);
              seg.put(wi++, make_uint2((unsigned int)slot[k], (unsigned int)(corr0 + uu) | ((unsigned int)mod << 30)), st);
            }
          }
        }
      }
      return;
    }
#pragma unroll
    for (int mod = 0; mod < 3; ++mod) {
      if ((mod == 0 && !KT::k2) || (mod == 1 && !KT::k3) || (mod == 2 && !KT::kn)) continue;
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) {
        const float v = val[mod][uu];
        if (!rescan) {
          cnt[k] += (int)(__float_as_uint(v) >> 31);
          flag = flag || (fabsf(v) <= bnd[mod][uu]);
        } else if (fabsf(v) <= bnd[mod][uu]) {
          cnt[k] -= (int)(__float_as_uint(v) >> 31);
          seg.push(make_uint2((unsigned int)slot[k], (unsigned int)(corr0 + uu) | ((unsigned int)mod << 30)), st);
        }
      }
    }
  };

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    mbar_wait(&bars[buf], (uint32_t)((t >> 1) & 1));
    const int tp = min(TILE, npairs - t * TILE);
    if (RAW) {
      // ---- transpose the stage into pair records: thread <-> 4 correspondences = 2 records; stage magnitude bounds
      int c0, cnt4;
      stage_range(t, c0, cnt4);
      float* recs = reinterpret_cast<float*>(&tile[0][0]);
      float m_w = 0.f, m_wc = 0.f, m_n = 0.f;
      for (int qd = tid; qd < (tp + 1) / 2; qd += THREADS) {
        const bool in_smem = 4 * qd + 4 <= cnt4;
        float vxc[12];  // camera points of the 4 correspondences (validity gate / 3-D bound), when the kind needs them
        float wn[4] = {0.f, 0.f, 0.f, 0.f};  // |x_w|
        float vnw[12];                       // world normals of the 4 correspondences (kinds with the normal test)
#pragma unroll
        for (int a = 0; a < NRAW; ++a) {
          const bool is_extra = a >= NREC;
          const float* g = is_extra ? src.xc : src.a[a];
          float v[12];
          if (in_smem) {
            const float4* p4 = reinterpret_cast<const float4*>(rawbuf + (size_t)(buf * NRAW + a) * kStageFloats + 12 * qd);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const float4 x = p4[i];
              v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
            }
          } else {  // frame tail (or padding): element-wise from global memory, NaN beyond the last correspondence
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ci = c0 + 4 * qd + j;
#pragma unroll
              for (int r = 0; r < 3; ++r) v[3 * j + r] = ci < n ? g[(size_t)ci * 3 + r] : CUDART_NAN_F;
            }
          }
          // what this array is: record order x_w, [x_c], [b], [n_w, n_c] (+ x_c staged on the side)
          const bool is_xw = a == 0;
          const bool is_xc = (KT::k3 && a == 1) || is_extra;
          const bool is_nw = KT::kn && a == NREC - 2, is_nc = KT::kn && a == NREC - 1;
          if (is_xc) {
#pragma unroll
            for (int i = 0; i < 12; ++i) vxc[i] = v[i];
          }
          if (is_xw || (is_xc && KT::k3) || is_nw || is_nc) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float m = sqrtf(v[3 * j] * v[3 * j] + v[3 * j + 1] * v[3 * j + 1] + v[3 * j + 2] * v[3 * j + 2]);
              const bool fin = m == m && m < CUDART_INF_F;
              if (is_xw) {
                wn[j] = fin ? m : CUDART_NAN_F;
                if (fin) m_w = fmaxf(m_w, m);
                if (!KT::k3 && fin) m_wc = fmaxf(m_wc, m);
              } else if (is_xc) {  // k3: bound of |x_w| + |x_c| like the pack kernel (a non-finite x_c contributes nothing)
                const float wv = wn[j];
                if (wv == wv) m_wc = fmaxf(m_wc, fin ? wv + m : wv);
              } else if (fin) {
                m_n = fmaxf(m_n, m);
              }
            }
          }
          if (is_nw) {
#pragma unroll
            for (int i = 0; i < 12; ++i) vnw[i] = v[i];
          }
          if (is_nc) {
            // the normal test sits inside `if (adapter.isValid(c))` (AbsoluteOrientationNormal.hpp:246,323,398): an
            // all-NaN camera point must never cast a normal vote -> poison its camera normal (x_c precedes n_c in NRAW order
            // only for k3; for the side-staged x_c the poison is applied below, after the loop)
            if (KT::k3) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (!(vxc[3 * j] == vxc[3 * j] || vxc[3 * j + 1] == vxc[3 * j + 1] || vxc[3 * j + 2] == vxc[3 * j + 2]))
                  v[3 * j] = v[3 * j + 1] = v[3 * j + 2] = CUDART_NAN_F;
            }
            // the nine products n_c,i n_w,j of every correspondence replace n_w and n_c in the record
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int pr = 2 * qd + h;
              if (pr < tp) {
                float2* dst = reinterpret_cast<float2*>(recs + (size_t)pr * (F4 * 4) + KT::off_nw);
                const float* c0p = v + 6 * h;      // n_c of correspondences 4 qd + 2h, + 1
                const float* w0p = vnw + 6 * h;    // n_w of the same two
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                  for (int j = 0; j < 3; ++j) dst[3 * i + j] = make_float2(__fmul_rn(c0p[i], w0p[j]), __fmul_rn(c0p[3 + i], w0p[3 + j]));
              }
            }
          }
          if (!is_extra && !is_nw && !is_nc) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // record 2 qd + h = correspondences 4 qd + 2h, 4 qd + 2h + 1
              const int pr = 2 * qd + h;
              if (pr < tp) {
                const float* e = v + 6 * h;
                float2* dst = reinterpret_cast<float2*>(recs + (size_t)pr * (F4 * 4) + 6 * a);
                dst[0] = make_float2(e[0], e[3]);
                dst[1] = make_float2(e[1], e[4]);
                dst[2] = make_float2(e[2], e[5]);
              }
            }
          }
        }
        if (XC_EXTRA) {  // x_c arrived after n_c: poison the records just written
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (!(vxc[3 * j] == vxc[3 * j] || vxc[3 * j + 1] == vxc[3 * j + 1] || vxc[3 * j + 2] == vxc[3 * j + 2])) {
              const int pr = 2 * qd + (j >> 1);
              if (pr < tp) {
                float* e = recs + (size_t)pr * (F4 * 4) + KT::off_nw + (j & 1);
#pragma unroll
                for (int q9 = 0; q9 < 9; ++q9) e[2 * q9] = CUDART_NAN_F;
              }
            }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        m_w = fmaxf(m_w, __shfl_xor_sync(0xffffffffu, m_w, o));
        m_wc = fmaxf(m_wc, __shfl_xor_sync(0xffffffffu, m_wc, o));
        m_n = fmaxf(m_n, __shfl_xor_sync(0xffffffffu, m_n, o));
      }
      if ((tid & 31) == 0) {
        if (m_w > 0.f) atomicMax(&mstage[3 * buf + 0], __float_as_uint(m_w));
        if (m_wc > 0.f) atomicMax(&mstage[3 * buf + 1], __float_as_uint(m_wc));
        if (m_n > 0.f) atomicMax(&mstage[3 * buf + 2], __float_as_uint(m_n));
      }
      __syncthreads();  // records + bounds complete; rawbuf[buf] is free
      if (tid == 0) {
        if (t + 2 < ntiles) issue_raw(t + 2, buf);
        mstage[3 * (buf ^ 1) + 0] = mstage[3 * (buf ^ 1) + 1] = mstage[3 * (buf ^ 1) + 2] = 0u;  // next stage's slots
      }
      set_bands(__uint_as_float(mstage[3 * buf + 0]), __uint_as_float(mstage[3 * buf + 1]), __uint_as_float(mstage[3 * buf + 2]));
    }
    const float* sp = reinterpret_cast<const float*>(&tile[RAW ? 0 : buf][0]);
    for (int sub = 0; sub < tp; sub += SUB) {
      bool flag = false;
#pragma unroll 2
      for (int pp = 0; pp < SUB; ++pp) {
        float rec[F4 * 4];
#pragma unroll
        for (int i = 0; i < F4; ++i) {
          const float4 v = reinterpret_cast<const float4*>(sp)[(sub + pp) * F4 + i];
          rec[4 * i] = v.x;
          rec[4 * i + 1] = v.y;
          rec[4 * i + 2] = v.z;
          rec[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < HPT; ++k) unit(rec, k, false, flag, 2 * (p_begin + t * TILE + sub + pp));
      }
      if (!DIRECT && flag) {
        for (int pp = 0; pp < SUB; ++pp) {
          float rec[F4 * 4];
#pragma unroll
          for (int i = 0; i < F4; ++i) {
            const float4 v = reinterpret_cast<const float4*>(sp)[(sub + pp) * F4 + i];
            rec[4 * i] = v.x;
            rec[4 * i + 1] = v.y;
            rec[4 * i + 2] = v.z;
            rec[4 * i + 3] = v.w;
          }
          bool dummy = false;
#pragma unroll
          for (int k = 0; k < HPT; ++k) unit(rec, k, true, dummy, 2 * (p_begin + t * TILE + sub + pp));
        }
      }
    }
    __syncthreads();  // RAW: everybody is done with the records (and has read the bounds) before the next transpose
    if (!RAW && tid == 0 && t + 2 < ntiles) {
      const int tn = t + 2;
      const int tpn = min(TILE, npairs - tn * TILE);
      const uint32_t bytes = (uint32_t)tpn * F4 * 16u;
      mbar_expect_tx(&bars[buf], bytes);
      tma_load_1d(&tile[buf][0], pk + (size_t)(p_begin + tn * TILE) * F4, bytes, &bars[buf]);
    }
  }
#pragma unroll
  for (int k = 0; k < HPT; ++k)
    if (slot[k] >= 0 && cnt[k] != 0) atomicAdd(&votes[slot[k]], cnt[k]);
  __shared__ unsigned int xtra_cta;
  if (DIRECT) {  // (uniform) evaluations behind the unit entries, for the n_borderline diagnostic
    if (tid == 0) xtra_cta = 0u;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xtra += __shfl_xor_sync(0xffffffffu, xtra, o);
    if ((tid & 31) == 0 && xtra) atomicAdd(&xtra_cta, (unsigned int)xtra);
    __syncthreads();
  }
  seg.publish(wl, st, DIRECT ? xtra_cta : 0u);
}

template <int KIND, int THREADS>
static int launch_multi_t(const FrameView& f, const HypGen* gen, const HypFast* fast, int slot_begin, int slot_end, Thresh th,
                           int32_t* votes, FrameStats* st, Worklist wl, int num_sms, cudaStream_t s) {
  typedef KindTraits<KIND> KT;
  constexpr int HPT = 2;
  constexpr int CTAS_PER_SM = 512 / THREADS;  // 16 warps per SM either way
  constexpr int F4 = KT::f4pp;
  // two stages of TILE pairs; sized so that CTAS_PER_SM CTAs fit on an SM for every kind (<= 98 KB per 256 threads)
  constexpr int TILE = (F4 <= 3 ? 1024 : (F4 <= 6 ? 512 : 256)) * (THREADS == 512 && F4 > 3 ? 2 : 1);
  // RAW: records (single) + two stages of raw triples of every staged array: (16 F4 + 48 NRAW) bytes per pair
  constexpr int NRAW = KT::arrays + ((KT::kn && !KT::k3) ? 1 : 0);
  constexpr int RTILE = (F4 <= 3 ? 1024 : 512) / CTAS_PER_SM;
  const int nslots = slot_end - slot_begin;
  const int gy = (nslots + THREADS * HPT - 1) / (THREADS * HPT);
  const int groups = f.npairs_pad / kSubPairs;
  int gx = (CTAS_PER_SM * num_sms) / gy;  // never more CTAs than the resident slots: a partial second wave doubles the time
  if (gx < 1) gx = 1;
  if (gx > groups) gx = groups;
  const int groups_per_cta = (groups + gx - 1) / gx;
  const int pairs_per_cta = groups_per_cta * kSubPairs;
  gx = (f.npairs_pad + pairs_per_cta - 1) / pairs_per_cta;
  // kinds with the 2-D test queue borderline evaluations on the spot (pixel-level thresholds put a percent of a good
  // hypothesis' evaluations inside the band); without it the group flag + rare re-walk is cheaper
  static const bool direct = getenv("RPE_MULTI_DIRECT") ? getenv("RPE_MULTI_DIRECT")[0] != '0' : KT::k2;
  int dev = 0;
  cudaGetDevice(&dev);
  MultiSrc src;
  src.pk = f.pk;
  src.n = f.n;
  src.xc = f.xc;
  static const bool unit_off = getenv("RPE_UNIT_ENTRIES") && getenv("RPE_UNIT_ENTRIES")[0] == '0';  // measurement aid
  src.unit_entries = (!unit_off && KIND != 1 && f.n <= (1 << 25)) ? 1 : 0;  // (2-D only: at most two values per unit, no gain)
  {
    int c = 0;
    src.a[c++] = f.xw;
    if (KT::k3) src.a[c++] = f.xc;
    if (KT::k2) src.a[c++] = f.bv;
    if (KT::kn) {
      src.a[c++] = f.nw;
      src.a[c++] = f.nc;
    }
    for (; c < 5; ++c) src.a[c] = nullptr;
  }
  if (frame_raw_ok(f, KIND)) {  // stream the caller's arrays: no packed copy
    const size_t rsmem = (size_t)RTILE * KT::f4raw * sizeof(float4) + 2 * (size_t)NRAW * RTILE * 6 * sizeof(float) + 2 * sizeof(uint64_t) + 32;
    static std::atomic<bool> rattr_set[64];
    if (dev >= 0 && dev < 64 && !rattr_set[dev].load(std::memory_order_acquire)) {
      cudaFuncSetAttribute(score_multi_fast_kernel<KIND, RTILE, THREADS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
      cudaFuncSetAttribute(score_multi_fast_kernel<KIND, RTILE, THREADS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
      rattr_set[dev].store(true, std::memory_order_release);
    }
    if (direct)
      score_multi_fast_kernel<KIND, RTILE, THREADS, true, true><<<dim3(gx, gy), THREADS, rsmem, s>>>(
          src, f.npairs_pad, pairs_per_cta, fast, gen, slot_begin, slot_end, th, votes, st, wl);
    else
      score_multi_fast_kernel<KIND, RTILE, THREADS, false, true><<<dim3(gx, gy), THREADS, rsmem, s>>>(
          src, f.npairs_pad, pairs_per_cta, fast, gen, slot_begin, slot_end, th, votes, st, wl);
    return gx * gy;
  }
  const size_t smem = 2 * (size_t)TILE * F4 * sizeof(float4) + 2 * sizeof(uint64_t) + 32;
  static std::atomic<bool> attr_set[64];  // the attribute is per device; setting it twice from two host threads is harmless
  if (dev >= 0 && dev < 64 && !attr_set[dev].load(std::memory_order_acquire)) {
    cudaFuncSetAttribute(score_multi_fast_kernel<KIND, TILE, THREADS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(score_multi_fast_kernel<KIND, TILE, THREADS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set[dev].store(true, std::memory_order_release);
  }
  if (direct)
    score_multi_fast_kernel<KIND, TILE, THREADS, true, false><<<dim3(gx, gy), THREADS, smem, s>>>(
        src, f.npairs_pad, pairs_per_cta, fast, gen, slot_begin, slot_end, th, votes, st, wl);
  else
    score_multi_fast_kernel<KIND, TILE, THREADS, false, false><<<dim3(gx, gy), THREADS, smem, s>>>(
        src, f.npairs_pad, pairs_per_cta, fast, gen, slot_begin, slot_end, th, votes, st, wl);
  return gx * gy;
}

// CTA width: 512 threads (one CTA per SM, a column of 1024 slots) for wide slot ranges, 256 threads (two CTAs per SM,
// 512 slots) for narrow ones (short passes, hypothesis-sharded frames). RPE_MULTI_THREADS=256|512 forces one.
template <int KIND>
static int launch_multi(const FrameView& f, const HypGen* gen, const HypFast* fast, int slot_begin, int slot_end, Thresh th,
                         int32_t* votes, FrameStats* st, Worklist wl, int num_sms, cudaStream_t s) {
  static const int forced = getenv("RPE_MULTI_THREADS") ? atoi(getenv("RPE_MULTI_THREADS")) : 0;
  const bool wide = forced ? forced == 512 : (slot_end - slot_begin) > 512;
  if (wide) return launch_multi_t<KIND, 512>(f, gen, fast, slot_begin, slot_end, th, votes, st, wl, num_sms, s);
  return launch_multi_t<KIND, 256>(f, gen, fast, slot_begin, slot_end, th, votes, st, wl, num_sms, s);
}

static bool g_use_packed = true;
static int g_nosync = 0;
void set_nosync(int v) { g_nosync = v; }
static int g_variant = 14;  // HPT=2, 1024-pair stages, 512 threads, 1 CTA per SM, 4-pair groups (profiles/r01_variant_sweep.md, r02_variant_sweep.md)
void set_use_packed(bool v) { g_use_packed = v; }
bool use_packed() { return g_use_packed; }
void set_score_variant(int v) { g_variant = v; }
int score_variant() { return g_variant; }

// RPE_SCORER_SHARED_SM=1 in the environment lets two scorer CTAs share an SM (measurement aid)
static int g_exclusive_sm = getenv("RPE_SCORER_SHARED_SM") ? 0 : 1;

template <bool PACKED, int HPT, int TILE, int THREADS, int MINB, int SUB>
static int launch_variant(const FrameView& f, const HypGen* gen, const HypFast* fast, int slot_begin, int slot_end,
                           Thresh th, int32_t* votes, FrameStats* st, Worklist wl, int num_sms, cudaStream_t s,
                           int corr_base, unsigned int seg_cap) {
  const int nslots = slot_end - slot_begin;
  const int hyp_per_cta = THREADS * HPT;
  const int gy = (nslots + hyp_per_cta - 1) / hyp_per_cta;
  // correspondences are handed out in groups of kSubPairs pairs (the pack pads to that), SUB divides into it
  const int groups = f.npairs_pad / SUB;
  int gx = (MINB * num_sms) / gy;  // at most MINB resident CTAs per SM in total (no partial second wave)
  if (gx < 1) gx = 1;
  if (gx > groups) gx = groups;
  const int groups_per_cta = (groups + gx - 1) / gx;
  const int pairs_per_cta = groups_per_cta * SUB;
  gx = (f.npairs_pad + pairs_per_cta - 1) / pairs_per_cta;
  size_t smem = 2 * (size_t)TILE * 3 * sizeof(float4) + 2 * sizeof(uint64_t);
  // One scorer CTA per SM is the design point (MINB = 1): ask for more than half of the SM's shared memory so that a
  // second context's scorer queues behind this one instead of time-slicing the same FMA pipe (same throughput,
  // twice the latency per launch); the small kernels of other frames still co-run in the remaining space.
  if (MINB == 1 && g_exclusive_sm && smem < (size_t)116 * 1024) smem = (size_t)116 * 1024;
  int dev = 0;
  cudaGetDevice(&dev);
  // measurement aid: RPE_FORCE_STREAM_SCORER=1 runs the sub-staged streaming scorer (no device copy) in place of the default one
  static const bool force_stream = getenv("RPE_FORCE_STREAM_SCORER") != nullptr;
  if (force_stream && PACKED && frame_raw_ok(f, 2) && MINB == 1 && THREADS == 512 && corr_base == 0 && seg_cap == 0 && gy == 1)
    return launch_score3d_stream(f, nullptr, nullptr, gen, fast, slot_begin, slot_end, th.thr3d, votes, st, wl, num_sms, s);
  if (PACKED && frame_raw_ok(f, 2)) {  // stream the caller's arrays: no packed copy
    constexpr int RT = (TILE > 1024 / MINB ? 1024 / MINB : TILE);  // 144 bytes of shared memory per pair and CTA
    static const int pcount = getenv("RPE_PCOUNT") ? getenv("RPE_PCOUNT")[0] - '0' : 1;  // 1: packed sign count (1 % faster, r02d)
    auto rk = pcount ? score3d_raw_kernel<HPT, RT, THREADS, MINB, SUB, 1> : score3d_raw_kernel<HPT, RT, THREADS, MINB, SUB, 0>;
    size_t rsmem = (size_t)RT * 3 * sizeof(float4) + 4 * (size_t)RT * 6 * sizeof(float) + 2 * sizeof(uint64_t) + 16;
    if (MINB == 1 && g_exclusive_sm && rsmem < (size_t)116 * 1024) rsmem = (size_t)116 * 1024;
    static std::atomic<bool> rattr_set[64];
    if (dev >= 0 && dev < 64 && !rattr_set[dev].load(std::memory_order_acquire)) {
      cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
      rattr_set[dev].store(true, std::memory_order_release);
    }
    if (seg_cap) wl.capacity = seg_cap * (unsigned int)(gx * gy);  // fixed segment size (chunked frames share one list;
                                                                   // the caller provides room for num_sms segments per chunk: gx * gy <= num_sms with MINB = 1)
    rk<<<dim3(gx, gy), THREADS, rsmem, s>>>(f.xw, f.xc, f.n, f.npairs_pad, pairs_per_cta, fast, gen, slot_begin, slot_end,
                                            th.thr3d, votes, st, wl, corr_base);
    return gx * gy;
  }
  auto kern = score3d_fast_kernel<PACKED, HPT, TILE, THREADS, MINB, SUB>;
  static std::atomic<bool> attr_set[64];  // the attribute is per device
  if (dev >= 0 && dev < 64 && !attr_set[dev].load(std::memory_order_acquire)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > (size_t)116 * 1024 ? smem : (size_t)116 * 1024));
    attr_set[dev].store(true, std::memory_order_release);
  }
  kern<<<dim3(gx, gy), THREADS, smem, s>>>(f.pk, f.npairs_pad, pairs_per_cta, fast, gen, slot_begin, slot_end, th.thr3d,
                                            votes, st, wl, g_nosync);
  return gx * gy;
}

int launch_score_fast(int method, const FrameView& f, const HypGen* gen, const HypFast* fast, int slot_begin,
                      int slot_end, Thresh th, int32_t* votes, FrameStats* st, Worklist wl, int num_sms,
                      cudaStream_t s, int corr_base, unsigned int seg_cap, int ur_lane) {
  if (slot_end - slot_begin <= 0 || f.n <= 0) return 0;
  if (method != RPE_SHINJI) {
    const int kind = (method_uses_2d(method) ? 1 : 0) | (method_uses_3d(method) ? 2 : 0) | (method_uses_nl(method) ? 4 : 0);
    if (!frame_raw_ok(f, kind) && f.pk_kind != kind) return 0;  // nothing packed for this family (caller bug)
    switch (kind) {
      case 1: return launch_multi<1>(f, gen, fast, slot_begin, slot_end, th, votes, st, wl, num_sms, s);
      case 3: return launch_multi<3>(f, gen, fast, slot_begin, slot_end, th, votes, st, wl, num_sms, s);
      case 5: return launch_multi<5>(f, gen, fast, slot_begin, slot_end, th, votes, st, wl, num_sms, s);
      case 6: return launch_multi<6>(f, gen, fast, slot_begin, slot_end, th, votes, st, wl, num_sms, s);
      case 7: return launch_multi<7>(f, gen, fast, slot_begin, slot_end, th, votes, st, wl, num_sms, s);
      default: return 0;
    }
  }
#define RPE_V(P, H, T, TH, MB, SB) return launch_variant<P, H, T, TH, MB, SB>(f, gen, fast, slot_begin, slot_end, th, votes, st, wl, num_sms, s, corr_base, seg_cap)
  if (!g_use_packed) {
    RPE_V(false, 2, 256, 256, 2, 8);
  }
  // A slot range narrower than one 1024-hypothesis CTA column (hypothesis-sharded frames, short passes): use narrower
  // CTAs and more of them per SM, so that the work shrinks with the range instead of idling half-empty threads.
  int variant = g_variant;
  // The uniform-register scorer (score_ur.cu) is an opt-in: measured on config #4 it reaches 76 % of the FMA pipe against
  // this file's 78 % (profiles/r02k_ur_kernel_ncu.md) — RPE_UR=1 lets it take whole dense frames against a full hypothesis
  // column when the launch goes to a scorer lane, variant 30 forces it for every frame it can take (tests).
  static const bool ur_on = getenv("RPE_UR") && getenv("RPE_UR")[0] == '1';
  if ((variant == 30 || (variant == 14 && ur_on && slot_end - slot_begin > 512 && f.n >= 65536)) && (ur_lane == 0 || ur_lane == 1) &&
      corr_base == 0 && seg_cap == 0 && slot_end - slot_begin <= 1024 && frame_raw_ok(f, 2)) {
    const int nseg = ur_lane == 0 ? launch_score3d_ur_lane0(f, gen, fast, slot_begin, slot_end, th.thr3d, votes, st, wl, num_sms, s)
                                  : launch_score3d_ur_lane1(f, gen, fast, slot_begin, slot_end, th.thr3d, votes, st, wl, num_sms, s);
    if (nseg > 0) return nseg;
  }
  if (variant == 30) variant = 14;
  if (variant == 14) {
    const int nslots = slot_end - slot_begin;
    if (nslots <= 128) variant = 16;
    else if (nslots <= 256) variant = 7;
    else if (nslots <= 512) variant = 1;
  }
  switch (variant) {  // (packed, hypotheses per thread, pairs per stage, threads, CTAs per SM, rescan group)
    case 1: RPE_V(true, 2, 512, 256, 2, 8);   // <= 512 slots
    case 7: RPE_V(true, 2, 256, 128, 4, 8);   // <= 256 slots
    case 16: RPE_V(true, 2, 256, 64, 8, 8);   // <= 128 slots
    // round-2 sweep of the raw-array scorer (profiles/r02_variant_sweep.md)
    case 24: RPE_V(true, 2, 1024, 512, 1, 8);   // the round-1 default (8-pair groups)
    case 27: RPE_V(true, 4, 1024, 256, 1, 4);
    default:
    case 14: RPE_V(true, 2, 1024, 512, 1, 4);  // the full 1024-hypothesis column; 4-pair groups keep the X0 / X1 / X2 runs of FFMA2
               // on the operand-reuse cache more often than 8-pair groups do (3 % faster, profiles/r02_variant_sweep.md)
  }
#undef RPE_V
}

// ================================================================================================
// exact-order evaluation: worklist fix-up and whole-frame fallback
// ================================================================================================
// modality codes carried in bits 30..31 of a worklist entry: 0 = 2-D, 1 = 3-D, 2 = normal, 3 = a unit entry (see fixup_kernel)
__device__ __forceinline__ bool exact_eval(int method, int modality, const FrameView& f, const HypGen& h, const float* Rm,
                                           int c, const Thresh& th) {
  if (modality == 1) {
    const F3 xc = load_col(f.xc, c);
    if (!ex_is_valid(xc)) return false;
    return ex_test_3d(h.q, h.t, load_col(f.xw, c), xc, th.thr3d);
  }
  if (modality == 2) {
    if (!ex_is_valid(load_col(f.xc, c))) return false;
    return ex_test_nl(h.q, load_col(f.nw, c), load_col(f.nc, c), th.cos_nl);
  }
  return ex_test_2d(h.q, h.t, method == RPE_KNEIP ? Rm : nullptr, load_col(f.xw, c), load_col(f.bv, c), th.cos_thr);
}

constexpr int kFixupChunks = 4;  // CTAs per worklist segment
constexpr unsigned int kFixupWindow = 1024;  // slots per histogram: the widest scorer column (512 threads x 2 hypotheses)
__global__ void __launch_bounds__(256)
fixup_kernel(int method, FrameView f, const HypGen* __restrict__ gen, Thresh th, int32_t* __restrict__ votes,
             const FrameStats* __restrict__ st, Worklist wl, int nseg, int slot_begin, int slot_end) {
  if (st->wl_overflow) {
    // the slot range is rescored exactly by the kernel that follows: hand it a clean vote table
    const int nthreads = gridDim.x * gridDim.y * blockDim.x;
    const int me = (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    for (int sidx = slot_begin + me; sidx < slot_end; sidx += nthreads) votes[sidx] = gen[sidx].valid ? 0 : -1;
    return;
  }
  const unsigned int cap = wl.capacity / (unsigned int)nseg;
  const unsigned int count = min(wl.counts[blockIdx.x], cap);
  if (count <= blockIdx.y * blockDim.x) return;  // nothing for this CTA (uniform)
  const uint2* seg = wl.entries + (size_t)blockIdx.x * cap;
  // A segment was filled by one scorer CTA, i.e. by one column of at most 1024 consecutive slots, and the good
  // hypotheses collect most of the borderline evaluations: votes are gathered in a shared-memory histogram of that
  // window and flushed with one global atomic per touched slot (a slot outside the window goes to memory directly).
  __shared__ int hist[kFixupWindow];
  const unsigned int window = ((unsigned int)seg[0].x - (unsigned int)slot_begin) / kFixupWindow;
  for (int i = threadIdx.x; i < kFixupWindow; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (unsigned int i = blockIdx.y * blockDim.x + threadIdx.x; i < count; i += gridDim.y * blockDim.x) {
    const uint2 e = seg[i];
    const int slot = (int)e.x;
    const int modality = (int)(e.y >> 30);
    const HypGen h = gen[slot];
    float Rm[9];
    if (method == RPE_KNEIP) ex_quat_to_matrix(h.q, Rm);
    int add = 0;
    if (modality == 3) {  // a unit entry of the multi-modality scorer: pair index, six borderline bits (2 modality + which of the pair)
      const int c0 = 2 * (int)(e.y & 0xffffffu);
      unsigned int bits = (e.y >> 24) & 0x3fu;
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1u;
        add += exact_eval(method, b >> 1, f, h, Rm, c0 + (b & 1), th) ? 1 : 0;
      }
    } else {
      add = exact_eval(method, modality, f, h, Rm, (int)(e.y & 0x3fffffffu), th) ? 1 : 0;
    }
    if (add) {
      const unsigned int rel = (unsigned int)slot - (unsigned int)slot_begin;
      if (rel / kFixupWindow == window)
        atomicAdd(&hist[rel % kFixupWindow], add);
      else
        atomicAdd(&votes[slot], add);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kFixupWindow; i += blockDim.x)
    if (hist[i]) atomicAdd(&votes[slot_begin + (int)(window * kFixupWindow) + i], hist[i]);
}

// Stage API (rpe_score called several times per frame): mark the queued evaluations as resolved so that the next
// call's fix-up does not add them again.
__global__ void consume_worklist_kernel(FrameStats* st) {
  if (threadIdx.x == 0) {
    st->wl_consumed += st->wl_count < 0xffffffffu ? st->wl_count : 0u;
    st->wl_count = 0;
  }
}
void launch_consume_worklist(FrameStats* st, cudaStream_t s) { consume_worklist_kernel<<<1, 32, 0, s>>>(st); }

void launch_fixup(int method, const FrameView& f, const HypGen* gen, Thresh th, int32_t* votes, FrameStats* st,
                  Worklist wl, int nseg, int slot_begin, int slot_end, cudaStream_t s) {
  if (nseg <= 0) return;
  // per-warp segments (2-D kinds) are 8-16 times smaller than per-CTA ones: one CTA each is enough
  fixup_kernel<<<dim3(nseg, nseg > 1024 ? 1 : kFixupChunks), 256, 0, s>>>(method, f, gen, th, votes, st, wl, nseg, slot_begin,
                                                                          slot_end);
}

// Whole-frame exact scoring. Thread <-> slot, CTA column <-> correspondence slice; every lane reads the
// same correspondence (uniform, L1-broadcast loads).
__global__ void __launch_bounds__(128)
score_exact_kernel(int method, FrameView f, const HypGen* __restrict__ gen, int slot_begin, int slot_end, Thresh th,
                   int32_t* __restrict__ votes, const FrameStats* __restrict__ st, int only_if_overflow, int corr_per_cta) {
  if (only_if_overflow && !st->wl_overflow) return;
  const int slot = slot_begin + blockIdx.y * blockDim.x + threadIdx.x;
  const bool live = slot < slot_end && gen[slot].valid != 0;
  HypGen h;
  if (live) h = gen[slot];
  else {
    h.q[0] = h.q[1] = h.q[2] = h.q[3] = CUDART_NAN_F;
    h.t[0] = h.t[1] = h.t[2] = CUDART_NAN_F;
  }
  float Rm[9];
  ex_quat_to_matrix(h.q, Rm);
  const bool u2 = method_uses_2d(method), u3 = method_uses_3d(method), un = method_uses_nl(method);
  const int c0 = blockIdx.x * corr_per_cta;
  const int c1 = min(c0 + corr_per_cta, f.n);
  int cnt = 0;
  for (int c = c0; c < c1; ++c) {
    if (un) cnt += exact_eval(method, 2, f, h, Rm, c, th) ? 1 : 0;
    if (u3) cnt += exact_eval(method, 1, f, h, Rm, c, th) ? 1 : 0;
    if (u2) cnt += exact_eval(method, 0, f, h, Rm, c, th) ? 1 : 0;
  }
  if (live && cnt) atomicAdd(&votes[slot], cnt);
}

__global__ void zero_votes_if_overflow_kernel(const HypGen* __restrict__ gen, int slot_begin, int slot_end,
                                              int32_t* __restrict__ votes, const FrameStats* __restrict__ st,
                                              int only_if_overflow) {
  if (only_if_overflow && !st->wl_overflow) return;
  for (int sidx = slot_begin + blockIdx.x * blockDim.x + threadIdx.x; sidx < slot_end; sidx += gridDim.x * blockDim.x)
    votes[sidx] = gen[sidx].valid ? 0 : -1;
}

void launch_score_exact(int method, const FrameView& f, const HypGen* gen, int slot_begin, int slot_end, Thresh th,
                        int32_t* votes, FrameStats* st, bool only_if_overflow, int num_sms, cudaStream_t s) {
  const int nslots = slot_end - slot_begin;
  if (nslots <= 0 || f.n <= 0) return;
  // after a fast pass the fix-up kernel has already cleaned the vote table when the worklist overflowed
  if (!only_if_overflow) zero_votes_if_overflow_kernel<<<4, 256, 0, s>>>(gen, slot_begin, slot_end, votes, st, 0);
  const int threads = 128;
  const int gy = (nslots + threads - 1) / threads;
  int gx = (4 * num_sms + gy - 1) / gy;
  if (gx < 1) gx = 1;
  if (gx > f.n) gx = f.n;
  const int corr_per_cta = (f.n + gx - 1) / gx;
  gx = (f.n + corr_per_cta - 1) / corr_per_cta;
  score_exact_kernel<<<dim3(gx, gy), threads, 0, s>>>(method, f, gen, slot_begin, slot_end, th, votes, st,
                                                      only_if_overflow ? 1 : 0, corr_per_cta);
}

// ================================================================================================
// FP32 FFMA peak microbenchmark (roofline denominator measured on the same device and run)
// ================================================================================================
template <bool PACKED>
__global__ void __launch_bounds__(256) ffma_bench_kernel(float* sink, int iters) {
  float2 acc[8];
  const float seed = 1.0f + 1e-7f * (float)(threadIdx.x + blockIdx.x);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(seed + i, seed - i);
  const float2 m = make_float2(0.999999f, 1.000001f);
  const float2 a = make_float2(1e-6f, -1e-6f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (PACKED) {
          acc[i] = __ffma2_rn(acc[i], m, a);
        } else {
          acc[i].x = fmaf(acc[i].x, m.x, a.x);
          acc[i].y = fmaf(acc[i].y, m.y, a.y);
        }
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y;
  if (r == 123.456f) sink[0] = r;
}
void launch_ffma_bench(float* sink, int iters, bool packed, int blocks, cudaStream_t s) {
  if (packed)
    ffma_bench_kernel<true><<<blocks, 256, 0, s>>>(sink, iters);
  else
    ffma_bench_kernel<false><<<blocks, 256, 0, s>>>(sink, iters);
}

int grid_blocks_for(int n, int threads) { return (n + threads - 1) / threads; }

}  // namespace rpe
