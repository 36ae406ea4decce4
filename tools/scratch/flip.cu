#include <cuda_runtime.h>
#include <cstdint>
struct Hyp { float nR[9]; float nt[3]; };
__constant__ Hyp chyp[1024];
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c){ return __ffma2_rn(a,b,c);} 
template<int P>
__global__ void __launch_bounds__(512,1) flipk(const float2* __restrict__ rec, int nh, unsigned* __restrict__ out, float thr2){
  float2 X[P][6];
  #pragma unroll
  for(int p=0;p<P;++p)
  #pragma unroll
  for(int i=0;i<6;++i) X[p][i]=rec[(threadIdx.x*P+p)*6+i];
  const float2 nlo=make_float2(-thr2,-thr2);
  unsigned acc=0;
  #pragma unroll 2
  for(int h=0;h<nh;++h){
    const Hyp& H=chyp[h];
    #pragma unroll
    for(int p=0;p<P;++p){
      float2 e0=__fadd2_rn(X[p][3],make_float2(H.nt[0],H.nt[0]));
      float2 e1=__fadd2_rn(X[p][4],make_float2(H.nt[1],H.nt[1]));
      float2 e2=__fadd2_rn(X[p][5],make_float2(H.nt[2],H.nt[2]));
      #pragma unroll
      for(int j=0;j<3;++j){
        e0=ffma2(make_float2(H.nR[j],H.nR[j]),X[p][j],e0);
        e1=ffma2(make_float2(H.nR[3+j],H.nR[3+j]),X[p][j],e1);
        e2=ffma2(make_float2(H.nR[6+j],H.nR[6+j]),X[p][j],e2);
      }
      float2 s=ffma2(e0,e0,nlo); s=ffma2(e1,e1,s); s=ffma2(e2,e2,s);
      acc += (__float_as_uint(s.x)>>31) + (__float_as_uint(s.y)>>31);
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}
template __global__ void flipk<4>(const float2*,int,unsigned*,float);
