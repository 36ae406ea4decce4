// score_ur.cu — the 3-D / 3-D scorer, correspondence-stationary: hypotheses in UNIFORM registers.
//
// Same evaluations, bit for bit, as score3d_raw_kernel (score.cu; reference loop pose/AbsoluteOrientation.hpp:135-143
// = :192-200): s = |x_c - t - R x_w|^2 - thr^2 with 3 FADD2 + 12 FFMA2 per pair of correspondences, sign(s) decides unless
// |s| <= band (then the evaluation goes to the exact fix-up). What changes is who holds what:
//
//   score3d_raw_kernel   thread <-> 2 hypotheses (24 scalars in registers), every lane reads the same pair record:
//                        FFMA2 acc, nR.F32, X, acc reads five distinct 32-bit registers -> 76 % of the lane rate
//   this kernel          thread <-> P pairs of correspondences (12 P registers, loaded ONCE), the CTA walks its
//                        hypotheses in a loop whose 12 scalars come from the constant bank into uniform registers:
//                        FFMA2 acc, X, UR.F32, acc reads four registers -> 92-98 % of the lane rate without any help
//                        from the operand-reuse cache (profiles/r02_ubench_ur.log)
//
// The price is the vote count: a hypothesis' votes are spread over the threads. Each thread packs the sign words of its
// P pairs (PRMT + IADD3, as before), one REDUX adds the packed words of the warp and one lane parks the sum, together with
// the ballot of the lanes that saw a borderline value, in the warp's own row of a shared-memory table (no atomics); flagged
// hypotheses are re-walked per warp after the loop, and the rows are added, decoded and flushed to the vote table when the
// CTA ends. Measured: the loop reaches 82.8 % of the FMA pipe, the launch 75.8 % (default kernel: 78.0 %) — an opt-in
// (RPE_UR=1), see profiles/r02_ur_scorer.md.
//
// A slice rarely is a whole number of T x P pairs (307 200 correspondences over 148 SMs: 2 076 pairs per CTA column against
// 512 x 4 = 2 048), and the FMA pipe wants the same number of warps on all four schedulers of the SM (704 threads = 6 6 5 5
// warps run as slowly as 768 would). So the CTA takes T x P pairs through the loop above and the remaining few ("tail", at
// most kUrMaxTail pairs) the other way round once the loop is done: thread <-> hypothesis, the tail's pair records read as
// shared-memory broadcasts.
//
// The hypotheses (HypFast, 48 B each, <= 1024 per launch) are copied into a __constant__ buffer in front of the launch,
// stream-ordered. A constant bank has 64 KB, so every scorer lane (capi.cu: consecutive scorers alternate between two lane
// streams and overlap head to tail) owns a buffer in a translation unit of its own: this file is compiled once per
// lane (-DRPE_UR_LANE=0 / 1).
#include <atomic>
#include <cstdlib>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "kernels.cuh"
#include "score_common.cuh"

#ifndef RPE_UR_LANE
#define RPE_UR_LANE 0
#endif

namespace rpe {

constexpr int kUrMaxHyp = 1024;
static __constant__ float4 c_hyp4[3 * kUrMaxHyp];  // HypFast[kUrMaxHyp]: nR[9] row-major, nt[3]
static_assert(sizeof(HypFast) == 48, "HypFast is three float4");

// a hypothesis as three 16-byte words: nR[0..3] | nR[4..7] | nR[8], nt[0..2]
struct UrHyp {
  float4 a, b, c;
  __device__ __forceinline__ explicit UrHyp(const float4* p) : a(p[0]), b(p[1]), c(p[2]) {}
};
// s of both correspondences of a pair (X0 X1 X2 = x_w, X3 X4 X5 = x_c, each .x / .y = first / second correspondence) under
// hypothesis H — the operation order of HypRegs<true>::eval in score.cu
__device__ __forceinline__ float2 ur_eval(const UrHyp& H, const float2 (&X)[6], float2 nlo) {
  const float nR[9] = {H.a.x, H.a.y, H.a.z, H.a.w, H.b.x, H.b.y, H.b.z, H.b.w, H.c.x};
  float2 e0 = __fadd2_rn(X[3], make_float2(H.c.y, H.c.y));
  float2 e1 = __fadd2_rn(X[4], make_float2(H.c.z, H.c.z));
  float2 e2 = __fadd2_rn(X[5], make_float2(H.c.w, H.c.w));
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    e0 = __ffma2_rn(make_float2(nR[j], nR[j]), X[j], e0);
    e1 = __ffma2_rn(make_float2(nR[3 + j], nR[3 + j]), X[j], e1);
    e2 = __ffma2_rn(make_float2(nR[6 + j], nR[6 + j]), X[j], e2);
  }
  float2 s = __ffma2_rn(e0, e0, nlo);
  s = __ffma2_rn(e1, e1, s);
  s = __ffma2_rn(e2, e2, s);
  return s;
}

constexpr int kUrMaxTail = 256;  // pairs of a slice beyond T x P

template <int LANE, int P, int TMAX>
__global__ void __launch_bounds__(TMAX, 1)
score3d_ur_kernel(const float* __restrict__ xw, const float* __restrict__ xc, int n, int npairs_pad, int pairs_per_cta,
                  const HypFast* __restrict__ fast, const HypGen* __restrict__ gen, int slot_begin, int nslots, int hyp_per_cta,
                  float thr, int32_t* __restrict__ votes, FrameStats* __restrict__ st, Worklist wl) {
  // shared memory: [band: hyp_per_cta floats][barrier, bound][tail pair records][raw x_w | raw x_c triples of the slice, LATER
  // the per-warp table [warps][hyp_per_cta] of (sign-word sum, borderline ballot)]
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int T = blockDim.x;
  const int cap = T * P;  // pairs the hot loop takes
  float* band_s = reinterpret_cast<float*>(smem_raw);  // NaN for an empty slot
  uint64_t* bar = reinterpret_cast<uint64_t*>(band_s + ((hyp_per_cta + 3) & ~3));
  unsigned int* mtile = reinterpret_cast<unsigned int*>(bar + 1);
  float4* tailrec = reinterpret_cast<float4*>(bar + 2);  // [kUrMaxTail][3]
  float* rw = reinterpret_cast<float*>(tailrec + 3 * kUrMaxTail);
  float* rc = rw + (size_t)pairs_per_cta * 6;
  uint2* tbl = reinterpret_cast<uint2*>(rw);

  const int tid = threadIdx.x;
  const int p_begin = blockIdx.x * pairs_per_cta;
  const int npairs = min(pairs_per_cta, npairs_pad - p_begin);
  const int ntail = max(npairs - cap, 0);
  const int h_begin = blockIdx.y * hyp_per_cta;
  const int nh = min(hyp_per_cta, nslots - h_begin);
  WlSegment seg(wl);

  // correspondences [c0, c0 + cnt4) of the slice go through TMA (c0 and cnt4 multiples of 4)
  const int c0 = 2 * p_begin;
  int cnt4 = min(2 * npairs, n - c0);
  cnt4 = cnt4 < 0 ? 0 : (cnt4 & ~3);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    *mtile = 0u;
    const uint32_t bytes = (uint32_t)cnt4 * 12u;
    mbar_expect_tx(bar, 2u * bytes);
    if (bytes) {
      tma_load_1d(rw, xw + (size_t)c0 * 3, bytes, bar);
      tma_load_1d(rc, xc + (size_t)c0 * 3, bytes, bar);
    }
  }
  __syncthreads();  // barrier initialised
  mbar_wait(bar, 0u);

  // pair lp of the slice as FFMA2 operands, and the largest |x_w| + |x_c| of its finite correspondences
  float mloc = 0.f;
  auto load_pair = [&](int lp, float2 (&Xp)[6]) {
    float w[6], c[6];
    if (2 * lp + 2 <= cnt4) {
      const float2* w2 = reinterpret_cast<const float2*>(rw + 6 * lp);
      const float2* c2 = reinterpret_cast<const float2*>(rc + 6 * lp);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float2 a = w2[i], b = c2[i];
        w[2 * i] = a.x; w[2 * i + 1] = a.y;
        c[2 * i] = b.x; c[2 * i + 1] = b.y;
      }
    } else {  // frame tail, padding, or beyond the slice: element-wise from the source arrays, NaN where there is nothing
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ci = c0 + 2 * lp + j;
        const bool have = lp < npairs && ci < n;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          w[3 * j + r] = have ? xw[(size_t)ci * 3 + r] : CUDART_NAN_F;
          c[3 * j + r] = have ? xc[(size_t)ci * 3 + r] : CUDART_NAN_F;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float m = sqrtf(w[3 * j] * w[3 * j] + w[3 * j + 1] * w[3 * j + 1] + w[3 * j + 2] * w[3 * j + 2]);
      const float mc = sqrtf(c[3 * j] * c[3 * j] + c[3 * j + 1] * c[3 * j + 1] + c[3 * j + 2] * c[3 * j + 2]);
      if (mc == mc && mc < CUDART_INF_F) m += mc;
      if (m == m && m < CUDART_INF_F) mloc = fmaxf(mloc, m);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      Xp[r] = make_float2(w[r], w[3 + r]);
      Xp[3 + r] = make_float2(c[r], c[3 + r]);
    }
  };
  // ---- this thread's P pairs (loaded ONCE), the tail's records
  float2 X[P][6];
#pragma unroll
  for (int p = 0; p < P; ++p) load_pair(p * T + tid, X[p]);
  for (int i = tid; i < ntail; i += T) {
    float2 Xt[6];
    load_pair(cap + i, Xt);
    tailrec[3 * i + 0] = make_float4(Xt[0].x, Xt[0].y, Xt[1].x, Xt[1].y);
    tailrec[3 * i + 1] = make_float4(Xt[2].x, Xt[2].y, Xt[3].x, Xt[3].y);
    tailrec[3 * i + 2] = make_float4(Xt[4].x, Xt[4].y, Xt[5].x, Xt[5].y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, o));
  if ((tid & 31) == 0 && mloc > 0.f) atomicMax(mtile, __float_as_uint(mloc));
  __syncthreads();  // bound complete; nobody reads the raw triples any more: the space becomes the per-warp table
  const float mcorr = __uint_as_float(*mtile);  // max (|x_w| + |x_c|) over the finite correspondences of the slice
  for (int h = tid; h < nh; h += T) {
    const int slot = slot_begin + h_begin + h;
    const HypFast& hf = fast[slot];
    const float tn = sqrtf(hf.nt[0] * hf.nt[0] + hf.nt[1] * hf.nt[1] + hf.nt[2] * hf.nt[2]);
    band_s[h] = gen[slot].valid != 0 ? guard_band_3d((mcorr + tn) * 1.0001f, thr) : CUDART_NAN_F;
  }
  __syncthreads();

  const float thr2 = __fmul_rn(thr, thr);
  const float2 nlo = make_float2(-thr2, -thr2);
  const bool lane0 = (tid & 31) == 0;
  // this warp's row of the table: per hypothesis (sum of the warp's sign words, ballot of the lanes that saw a borderline value)
  uint2* my_tbl = tbl + (size_t)(tid >> 5) * hyp_per_cta;

  // ---- the hot loop: nothing thread-divergent in here (ptxas gives up the uniform registers for the whole loop otherwise).
  // The index into the constant bank is an induction variable of its own, hidden from the optimiser: merged with h, which
  // also addresses the per-warp table, it ends up in a vector register and the scalars are fetched with LDC, not LDCU.
  int hu = h_begin;
#pragma unroll 2
  for (int h = 0; h < nh; ++h) {
    asm volatile("" : "+r"(hu));
    const UrHyp H(c_hyp4 + 3 * hu);
    ++hu;
    const float band = band_s[h];  // NaN for an empty slot: its evaluations are never borderline and never flushed
    unsigned int pacc = 0u;
    float smin = CUDART_INF_F;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float2 s = ur_eval(H, X[p], nlo);
      pacc += sign_words(s);
      smin = fminf(fminf(smin, fabsf(s.x)), fabsf(s.y));  // NaN operands are ignored: a NaN evaluation is never an inlier
    }
    const unsigned int hot = __ballot_sync(0xffffffffu, smin <= band);
    const unsigned int wsum = __reduce_add_sync(0xffffffffu, pacc);
    if (lane0) my_tbl[h] = make_uint2(wsum, hot);
  }
  __syncwarp();
  // ---- second pass, per warp: the hypotheses some lane flagged, thread-divergent. A flagged lane takes its borderline
  // evaluations out of the sum (a negative s.x entered it as 0xffff, a negative s.y as 0xffff0000) and queues them for the
  // exact fix-up.
  for (int hb = 0; hb < nh; hb += 32) {
    const int hl = hb + (tid & 31);
    unsigned int pending = __ballot_sync(0xffffffffu, hl < nh && my_tbl[hl].y != 0u);
    while (pending) {
      const int h = hb + (__ffs(pending) - 1);
      pending &= pending - 1u;
      const unsigned int hot = my_tbl[h].y;
      unsigned int back = 0u;
      if ((hot >> (tid & 31)) & 1u) {
        const UrHyp H(c_hyp4 + 3 * (h_begin + h));
        const float band = band_s[h];
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const float2 s = ur_eval(H, X[p], nlo);
          const float sv[2] = {s.x, s.y};
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (fabsf(sv[u]) <= band) {
              if (__float_as_uint(sv[u]) >> 31) back += u == 0 ? 0xffffu : 0xffff0000u;
              const unsigned int corr = (unsigned int)(2 * (p_begin + p * T + tid) + u);
              seg.push(make_uint2((unsigned int)(slot_begin + h_begin + h), corr | (1u << 30)), st);
            }
          }
        }
      }
      back = __reduce_add_sync(0xffffffffu, back);
      __syncwarp();
      if (lane0) my_tbl[h].x -= back;
      __syncwarp();
    }
  }
  __syncthreads();
  // ---- flush. A warp's word is W = 0xffff nx + 0xffff0000 ny mod 2^32 (nx / ny: negative decision values in the first /
  // second correspondence of its pairs); the words of the warps add up, and -W = nx + 2^16 (ny - nx) (sums stay below
  // 2^16: T P <= 4096 pairs). The thread that flushes a hypothesis first scores the slice's tail pairs against it.
  const int nwarps = T >> 5;
  for (int h = tid; h < nh; h += T) {
    unsigned int v = 0u;
    for (int w = 0; w < nwarps; ++w) v -= tbl[(size_t)w * hyp_per_cta + h].x;
    const unsigned int lo = v & 0xffffu, hi = v >> 16;
    int cnt = (int)(lo + ((hi + lo) & 0xffffu));
    const float band = band_s[h];
    if (band == band) {
      if (ntail > 0) {
        const UrHyp H(reinterpret_cast<const float4*>(fast + slot_begin + h_begin + h));
        for (int i = 0; i < ntail; ++i) {
          const float4 ra = tailrec[3 * i], rb = tailrec[3 * i + 1], rcd = tailrec[3 * i + 2];
          const float2 Xt[6] = {make_float2(ra.x, ra.y), make_float2(ra.z, ra.w), make_float2(rb.x, rb.y),
                                make_float2(rb.z, rb.w), make_float2(rcd.x, rcd.y), make_float2(rcd.z, rcd.w)};
          const float2 s = ur_eval(H, Xt, nlo);
          const float sv[2] = {s.x, s.y};
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (fabsf(sv[u]) <= band)
              seg.push(make_uint2((unsigned int)(slot_begin + h_begin + h), (unsigned int)(2 * (p_begin + cap + i) + u) | (1u << 30)), st);
            else
              cnt += (int)(__float_as_uint(sv[u]) >> 31);
          }
        }
      }
      if (cnt != 0) atomicAdd(&votes[slot_begin + h_begin + h], cnt);
    }
  }
  seg.publish(wl, st);
}

// ---- host side ---------------------------------------------------------------------------------------------------
struct UrShape {
  int P, T, hsplit, gx, pairs_per_cta;
  double cost;
};

// The CTA shape for a frame: `hsplit` CTA rows share the hypotheses, num_sms / hsplit CTA columns share the pairs, every
// thread holds P pairs, pairs beyond T P go to the tail. Rates: share of the FMA pipe the hot loop reaches with that shape
// (tools/ubench_ur.cu, profiles/r02_ubench_ur2.log): the warps must spread evenly over the four schedulers.
static UrShape ur_pick_shape(int npairs_pad, int nslots, int num_sms) {
  static const int forceP = getenv("RPE_UR_P") ? atoi(getenv("RPE_UR_P")) : 0;
  static const int forceT = getenv("RPE_UR_T") ? atoi(getenv("RPE_UR_T")) : 0;
  static const int forceH = getenv("RPE_UR_HSPLIT") ? atoi(getenv("RPE_UR_HSPLIT")) : 0;
  struct Cand { int P, T; double rate; };
  static const Cand cands[] = {{4, 512, 0.82}, {3, 640, 0.81}, {2, 1024, 0.78}, {3, 704, 0.75}, {4, 384, 0.72}, {3, 512, 0.78},
                               {2, 512, 0.73}, {4, 256, 0.62}, {2, 256, 0.55}, {4, 128, 0.45}, {2, 128, 0.4}, {2, 64, 0.25}};
  UrShape best{0, 0, 0, 0, 0, 1e300};
  for (int hs = 1; hs <= 8; hs *= 2) {
    if (forceH && hs != forceH) continue;
    if (hs > nslots) break;
    int gx = num_sms / hs;
    if (gx < 1) continue;
    int ppc = (npairs_pad + gx - 1) / gx;
    ppc = (ppc + 1) & ~1;  // even: a slice starts at a multiple of 4 correspondences (bulk-copy alignment)
    if (ppc < 2) ppc = 2;
    gx = (npairs_pad + ppc - 1) / ppc;
    const int nh = (nslots + hs - 1) / hs;
    for (const Cand& c : cands) {
      if (forceP && c.P != forceP) continue;
      if (forceT && c.T != forceT) continue;
      const int cap = c.T * c.P;
      const int tail = ppc > cap ? ppc - cap : 0;
      if (tail > kUrMaxTail) continue;
      if ((size_t)ppc * 48 + (size_t)nh * 4 + kUrMaxTail * 48 + 64 > (size_t)200 * 1024) continue;
      if ((size_t)(c.T / 32) * nh * 8 + (size_t)nh * 4 + kUrMaxTail * 48 + 64 > (size_t)200 * 1024) continue;
      // time ~ per CTA: hot loop nh x cap thread-pairs at `rate`, tail nh x tail pairs one thread each at a fifth of it
      const double cost = (double)nh * cap / c.rate + (double)((nh + c.T - 1) / c.T) * c.T * tail / 0.2 + 0.02 * cap * 1024;
      if (cost < best.cost) best = UrShape{c.P, c.T, hs, gx, ppc, cost};
    }
  }
  return best;
}

template <int P, int TMAX>
static int ur_launch(const UrShape& sh, const FrameView& f, const HypGen* gen, const HypFast* fast, int slot_begin, int nslots,
                     float thr3d, int32_t* votes, FrameStats* st, Worklist wl, cudaStream_t s) {
  auto k = score3d_ur_kernel<RPE_UR_LANE, P, TMAX>;
  const int hyp_per_cta = (nslots + sh.hsplit - 1) / sh.hsplit;
  const size_t raw_bytes = (size_t)sh.pairs_per_cta * 48, tbl_bytes = (size_t)(sh.T / 32) * hyp_per_cta * 8;
  size_t smem = (size_t)((hyp_per_cta + 3) & ~3) * 4 + 16 + (size_t)kUrMaxTail * 48 + (raw_bytes > tbl_bytes ? raw_bytes : tbl_bytes);
  if (smem < (size_t)116 * 1024) smem = (size_t)116 * 1024;  // one scorer CTA per SM (see launch_variant in score.cu)
  static std::atomic<int> attr[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && attr[dev].load(std::memory_order_acquire) < (int)smem) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr[dev].store((int)smem, std::memory_order_release);
  }
  k<<<dim3(sh.gx, sh.hsplit), sh.T, smem, s>>>(f.xw, f.xc, f.n, f.npairs_pad, sh.pairs_per_cta, fast, gen, slot_begin, nslots,
                                               hyp_per_cta, thr3d, votes, st, wl);
  return sh.gx * sh.hsplit;
}

#if RPE_UR_LANE == 0
// host logic only (no device needed): the shape the launcher would pick; out = {P, T, hsplit, columns, pairs per column, tail}
int ur_shape_for(int npairs_pad, int nslots, int num_sms, int* out6) {
  if (npairs_pad <= 0 || nslots <= 0 || nslots > kUrMaxHyp || num_sms <= 0) return 0;
  const UrShape sh = ur_pick_shape(npairs_pad, nslots, num_sms);
  if (sh.P == 0) return 0;
  out6[0] = sh.P;
  out6[1] = sh.T;
  out6[2] = sh.hsplit;
  out6[3] = sh.gx;
  out6[4] = sh.pairs_per_cta;
  out6[5] = sh.pairs_per_cta > sh.T * sh.P ? sh.pairs_per_cta - sh.T * sh.P : 0;
  return 1;
}
#endif

#define RPE_UR_CAT2(a, b) a##b
#define RPE_UR_CAT(a, b) RPE_UR_CAT2(a, b)
// Returns the number of worklist segments (= CTAs), 0 if the frame / slot range is not for this kernel.
int RPE_UR_CAT(launch_score3d_ur_lane, RPE_UR_LANE)(const FrameView& f, const HypGen* gen, const HypFast* fast, int slot_begin,
                                                    int slot_end, float thr3d, int32_t* votes, FrameStats* st, Worklist wl,
                                                    int num_sms, cudaStream_t s) {
  const int nslots = slot_end - slot_begin;
  if (nslots <= 0 || nslots > kUrMaxHyp || f.n <= 0 || !frame_raw_ok(f, 2)) return 0;
  const UrShape sh = ur_pick_shape(f.npairs_pad, nslots, num_sms);
  if (sh.P == 0) return 0;
  void* sym = nullptr;
  if (cudaGetSymbolAddress(&sym, c_hyp4) != cudaSuccess) return 0;
  if (cudaMemcpyAsync(sym, fast + slot_begin, (size_t)nslots * sizeof(HypFast), cudaMemcpyDeviceToDevice, s) != cudaSuccess) return 0;
  switch (sh.P) {
    case 2: return ur_launch<2, 1024>(sh, f, gen, fast, slot_begin, nslots, thr3d, votes, st, wl, s);
    case 3: return ur_launch<3, 704>(sh, f, gen, fast, slot_begin, nslots, thr3d, votes, st, wl, s);
    default: return ur_launch<4, 576>(sh, f, gen, fast, slot_begin, nslots, thr3d, votes, st, wl, s);
  }
}

}  // namespace rpe
