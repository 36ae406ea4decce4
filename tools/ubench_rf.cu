// Register-file bandwidth of FFMA2 on sm_100a: does a packed FMA whose five 32-bit source registers are all distinct
// (scalar multiplicand + 64-bit pair + 64-bit accumulator) issue every 2 cycles or every 3?  Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_rf tools/ubench_rf.cu
// MODE 0: runs of 6 FFMA2 share the pair operand (operand-reuse cache can serve it)
// MODE 1: every FFMA2 of a run has its own pair operand (nothing to reuse)
// MODE 2: like 1 with scalar FFMA (3 distinct registers)
// MODE 3: like 0 with scalar FFMA (multiplicand shared by the run)
// MODE 4: like 0, the scalar multiplicand in a UNIFORM register (FFMA2 R, R.F32x2, UR.F32, R.F32x2): pair shared by runs
// MODE 5: like 1, the scalar multiplicand in a uniform register: four distinct registers per FFMA2, nothing to reuse
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

__constant__ float cscal[32];

template <int MODE>
__global__ void __launch_bounds__(256) k(float* sink, const float* src, int iters) {
  float r[18];
  for (int i = 0; i < 18; ++i) r[i] = MODE >= 4 ? cscal[i] : src[i] + 1e-6f * threadIdx.x;
  float2 X[6];
  for (int i = 0; i < 6; ++i) X[i] = make_float2(src[20 + 2 * i] + 1e-7f * threadIdx.x, src[21 + 2 * i]);
  float2 acc[6];
  for (int i = 0; i < 6; ++i) acc[i] = make_float2(src[i], src[i + 1]);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (MODE == 4) acc[i] = __ffma2_rn(make_float2(r[6 * j + i], r[6 * j + i]), X[j], acc[i]);
        if (MODE == 5) acc[i] = __ffma2_rn(make_float2(r[6 * j + i], r[6 * j + i]), X[i], acc[i]);
        if (MODE == 0) acc[i] = __ffma2_rn(make_float2(r[6 * j + i], r[6 * j + i]), X[j], acc[i]);
        if (MODE == 1) acc[i] = __ffma2_rn(make_float2(r[6 * j + i], r[6 * j + i]), X[i], acc[i]);
        if (MODE == 2) {
          acc[i].x = fmaf(r[6 * j + i], X[i].x, acc[i].x);
          acc[i].y = fmaf(r[(6 * j + i + 7) % 18], X[i].y, acc[i].y);
        }
        if (MODE == 3) {
          acc[i].x = fmaf(r[6 * j + i], X[j].x, acc[i].x);
          acc[i].y = fmaf(r[(6 * j + i + 7) % 18], X[j].x, acc[i].y);
        }
      }
    }
  }
  float o = 0.f;
  for (int i = 0; i < 6; ++i) o += acc[i].x + acc[i].y;
  if (o == 123.456f) sink[0] = o;
}

template <int MODE>
void run(const char* name, float* sink, float* src, int ctas_per_sm) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const size_t smem = (size_t)(220 * 1024 / ctas_per_sm) & ~(size_t)1023;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int blocks = 148 * ctas_per_sm;
  k<MODE><<<blocks, 256, smem>>>(sink, src, ITERS);
  cudaEventRecord(a);
  k<MODE><<<blocks, 256, smem>>>(sink, src, ITERS);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double lane_fma = (double)blocks * 256 * ITERS * 18 * 2;
  printf("%-52s %d CTA/SM (%2d warps/SM) %8.3f ms = %6.2f T lane-FMA/s (%.0f%% of 37.2)\n", name, ctas_per_sm, 8 * ctas_per_sm, ms,
         lane_fma / ms / 1e9, lane_fma / ms / 1e9 / 37.22 * 100);
}

int main() {
  float *sink, *src;
  cudaMalloc(&sink, 64);
  cudaMalloc(&src, 256);
  float h[64];
  for (int i = 0; i < 64; ++i) h[i] = 0.001f * (i + 1);
  cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(cscal, h, 32 * sizeof(float));
  for (int c : {2, 3, 4, 8}) {
    run<4>("FFMA2 UNIFORM scalar x pair + acc, pair shared by runs", sink, src, c);
    run<5>("FFMA2 UNIFORM scalar x pair + acc, 4 distinct registers", sink, src, c);
    run<0>("FFMA2 scalar x pair + acc, pair shared by runs of 6", sink, src, c);
    run<1>("FFMA2 scalar x pair + acc, 5 distinct registers", sink, src, c);
    run<2>("FFMA 3 distinct registers", sink, src, c);
    run<3>("FFMA multiplicand shared by runs", sink, src, c);
  }
  return 0;
}
