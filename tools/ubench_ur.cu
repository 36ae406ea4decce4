// Ablation of the uniform-register scorer's hot loop (score_ur.cu) on sm_100a: which of the per-hypothesis extras keeps the
// FMA pipe at ~76 % when a pure FFMA2 stream with a uniform operand reaches 97 %?  Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_ur tools/ubench_ur.cu
// FLAGS bit 0: packed sign count (PRMT + IADD3 + REDUX + STS), bit 1: guard-band test (FMNMX3 + FSETP + VOTE),
// bit 2: band from shared memory (LDS) instead of a constant.
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>

__constant__ float4 chyp[3 * 1024];

__device__ __forceinline__ unsigned int sign_words(float2 s) {
  unsigned int d;
  asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(d) : "r"(__float_as_uint(s.x)), "r"(__float_as_uint(s.y)));
  return d;
}

template <int P, int FLAGS, int UNROLL, int ADDMODE = 0>
__global__ void __launch_bounds__(P == 2 ? 1024 : P == 3 ? 704 : 576, 1) k(const float2* __restrict__ rec, int nh, uint2* __restrict__ out, float thr2) {
  extern __shared__ float smem[];
  float* band_s = smem;
  uint2* tbl = reinterpret_cast<uint2*>(smem + 1024);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) band_s[i] = 1e-6f * (float)(i & 7);
  __syncthreads();
  float2 X[P][6];
#pragma unroll
  for (int p = 0; p < P; ++p)
#pragma unroll
    for (int i = 0; i < 6; ++i) X[p][i] = rec[((blockIdx.x * blockDim.x + threadIdx.x) * P + p) * 6 + i];
  const float2 nlo = make_float2(-thr2, -thr2);
  float2 one2 = make_float2(1.f, 1.f);
  asm volatile("" : "+f"(one2.x), "+f"(one2.y));  // keep it a register pair
  const bool lane0 = (threadIdx.x & 31) == 0;
  uint2* my_tbl = tbl + (threadIdx.x >> 5) * nh;
  unsigned int keep = 0u;
  int hu = 0;
#pragma unroll UNROLL
  for (int h = 0; h < nh; ++h) {
    asm volatile("" : "+r"(hu));
    const float4 a = chyp[3 * hu], b = chyp[3 * hu + 1], c = chyp[3 * hu + 2];
    ++hu;
    const float nR[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x};
    const float band = (FLAGS & 4) ? band_s[h] : 1e-6f;
    unsigned int pacc = 0u;
    float smin = CUDART_INF_F;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      float2 e0, e1, e2;
      if (ADDMODE == 0) {
        e0 = __fadd2_rn(X[p][3], make_float2(c.y, c.y));
        e1 = __fadd2_rn(X[p][4], make_float2(c.z, c.z));
        e2 = __fadd2_rn(X[p][5], make_float2(c.w, c.w));
      } else {  // the same sum as a packed FMA: 1 * nt + x_c (one rounding, bit-identical)
        e0 = __ffma2_rn(one2, make_float2(c.y, c.y), X[p][3]);
        e1 = __ffma2_rn(one2, make_float2(c.z, c.z), X[p][4]);
        e2 = __ffma2_rn(one2, make_float2(c.w, c.w), X[p][5]);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        e0 = __ffma2_rn(make_float2(nR[j], nR[j]), X[p][j], e0);
        e1 = __ffma2_rn(make_float2(nR[3 + j], nR[3 + j]), X[p][j], e1);
        e2 = __ffma2_rn(make_float2(nR[6 + j], nR[6 + j]), X[p][j], e2);
      }
      float2 s = __ffma2_rn(e0, e0, nlo);
      s = __ffma2_rn(e1, e1, s);
      s = __ffma2_rn(e2, e2, s);
      if (FLAGS & 1) pacc += sign_words(s);
      else keep ^= __float_as_uint(s.x) ^ __float_as_uint(s.y);
      if (FLAGS & 2) smin = fminf(fminf(smin, fabsf(s.x)), fabsf(s.y));
    }
    unsigned int hot = 0u, wsum = 0u;
    if (FLAGS & 2) hot = __ballot_sync(0xffffffffu, smin <= band);
    if (FLAGS & 1) wsum = __reduce_add_sync(0xffffffffu, pacc);
    if (FLAGS & 3) {
      if (lane0) my_tbl[h] = make_uint2(wsum, hot);
    }
  }
  if (keep == 0x12345u || !(FLAGS & 3)) out[blockIdx.x * blockDim.x + threadIdx.x] = make_uint2(keep, 0);
  __syncthreads();
  if (FLAGS & 3) out[blockIdx.x * blockDim.x + threadIdx.x] = tbl[threadIdx.x];
}

template <int P, int FLAGS, int UNROLL, int ADDMODE = 0>
void run(const char* name, int T, int nh, const float2* rec, uint2* out) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  size_t smem = 1024 * 4 + (size_t)(T / 32) * nh * 8;
  if (smem < 120 * 1024) smem = 120 * 1024;
  cudaFuncSetAttribute(k<P, FLAGS, UNROLL, ADDMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<P, FLAGS, UNROLL, ADDMODE><<<148, T, smem>>>(rec, nh, out, 0.0625f);
  float best = 1e9f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a);
    k<P, FLAGS, UNROLL, ADDMODE><<<148, T, smem>>>(rec, nh, out, 0.0625f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  cudaError_t e = cudaGetLastError();
  const double lane_fma = 148.0 * T * P * nh * 15 * 2;
  printf("%-58s add=%d P=%d T=%4d nh=%4d unroll=%d  %7.4f ms  %5.1f%% of the FMA pipe%s\n", name, ADDMODE, P, T, nh, UNROLL, best,
         lane_fma / best / 1e9 / 37.22 * 100, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  float2* rec;
  uint2* out;
  cudaMalloc(&rec, (size_t)148 * 1024 * 4 * 6 * sizeof(float2));
  cudaMalloc(&out, (size_t)148 * 1024 * sizeof(uint2));
  cudaMemset(rec, 0x3c, (size_t)148 * 1024 * 4 * 6 * sizeof(float2));
  float h[12 * 1024];
  for (int i = 0; i < 12 * 1024; ++i) h[i] = 0.001f * (float)((i * 7919) % 1000) - 0.5f;
  cudaMemcpyToSymbol(chyp, h, sizeof(h));
  run<3, 0, 2>("evaluations only (+ 1 LOP3 per pair)", 704, 512, rec, out);
  run<3, 0, 2, 1>("evaluations only (+ 1 LOP3 per pair)", 704, 512, rec, out);
  run<3, 7, 2>("all (the kernel's loop)", 704, 512, rec, out);
  run<3, 7, 2, 1>("all (the kernel's loop)", 704, 512, rec, out);
  run<4, 0, 2>("evaluations only", 512, 512, rec, out);
  run<4, 0, 2, 1>("evaluations only", 512, 512, rec, out);
  run<4, 7, 2>("all", 512, 512, rec, out);
  run<4, 7, 2, 1>("all", 512, 512, rec, out);
  run<3, 7, 2>("all", 768, 512, rec, out);
  run<3, 7, 2, 1>("all", 768, 512, rec, out);
  run<3, 7, 2, 1>("all", 640, 512, rec, out);
  run<2, 7, 2>("all", 1024, 512, rec, out);
  run<2, 7, 2, 1>("all", 1024, 512, rec, out);
  run<2, 7, 2, 1>("all", 512, 1024, rec, out);
  run<2, 7, 4, 1>("all", 1024, 512, rec, out);
  run<4, 7, 4, 1>("all", 512, 512, rec, out);
  return 0;
}
