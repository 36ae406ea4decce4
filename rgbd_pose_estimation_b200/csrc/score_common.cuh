// score_common.cuh — device helpers shared by the scorer TUs (score.cu, score_ur.cu): mbarrier + 1-D bulk TMA,
// the CTA's segment of the borderline worklist, the 3-D guard band, the packed sign words.
#ifndef RPE_SCORE_COMMON_CUH_
#define RPE_SCORE_COMMON_CUH_

#include <cuda_runtime.h>
#include <math_constants.h>

#include "kernels.cuh"

namespace rpe {

// ================================================================================================
// small PTX helpers: mbarrier + 1-D bulk TMA
// ================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

// ================================================================================================
// this CTA's segment of the borderline worklist (see struct Worklist)
// ================================================================================================
struct WlSegment {
  uint2* base;
  unsigned int cap;
  unsigned int* n;  // shared-memory counter
  __device__ __forceinline__ explicit WlSegment(const Worklist& wl) {
    __shared__ unsigned int counter;
    n = &counter;
    const unsigned int nseg = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    cap = wl.capacity / nseg;
    base = wl.entries + (size_t)cta * cap;
    if (threadIdx.x == 0) counter = 0;  // ordered before the first push by the __syncthreads after the mbarrier init
  }
  __device__ __forceinline__ void push(uint2 e, FrameStats* st) {
    const unsigned int i = atomicAdd(n, 1u);
    if (i < cap)
      base[i] = e;
    else
      st->wl_overflow = 1u;
  }
  // several entries of one thread with a single shared-memory atomic: reserve(k), then put(i), put(i + 1), ...
  __device__ __forceinline__ unsigned int reserve(unsigned int k) { return atomicAdd(n, k); }
  __device__ __forceinline__ void put(unsigned int i, uint2 e, FrameStats* st) {
    if (i < cap)
      base[i] = e;
    else
      st->wl_overflow = 1u;
  }
  // every thread of the CTA calls this once, after its last push. `extra`: evaluations the entries stand for beyond one
  // each (a unit entry carries up to six), summed over the CTA by the caller — only the n_borderline diagnostic sees it.
  __device__ __forceinline__ void publish(const Worklist& wl, FrameStats* st, unsigned int extra = 0u) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int cnt = *n;
      wl.counts[blockIdx.y * gridDim.x + blockIdx.x] = cnt < cap ? cnt : cap;
      if (cnt + extra) atomicAdd(&st->wl_count, cnt + extra);
    }
  }
};

// The error of s is 2|e| de + de^2 with |e| ~ thr and de <= 26.8 u M: the second-order term only matters for
// thresholds down at the rounding level of the coordinates (thr <~ 30 u M), where it is covered by widening thr.
__device__ __forceinline__ float guard_band_3d(float M, float thr) {
  const float u = 5.9604644775390625e-08f;
  return (thr + 32.f * u * M) * u * (64.f * M + 16.f * thr);
}

// (s.y < 0 ? 0xffff0000 : 0) | (s.x < 0 ? 0x0000ffff : 0): PRMT in its generic mode replicates the sign of the selected byte
// when bit 3 of the selector nibble is set (0xB = sign of byte 3 of a, 0xF = sign of byte 3 of b)
__device__ __forceinline__ unsigned int sign_words(float2 s) {
  unsigned int d;
  asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(d) : "r"(__float_as_uint(s.x)), "r"(__float_as_uint(s.y)));
  return d;
}

}  // namespace rpe

#endif  // RPE_SCORE_COMMON_CUH_
