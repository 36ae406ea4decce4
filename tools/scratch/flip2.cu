#include <cuda_runtime.h>
#include <cstdint>
#include <math_constants.h>
struct Hyp { float nR[9]; float nt[3]; };
__constant__ Hyp chyp[1024];
__device__ __forceinline__ unsigned int sign_words(float2 s) {
  unsigned int d;
  asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(d) : "r"(__float_as_uint(s.x)), "r"(__float_as_uint(s.y)));
  return d;
}
template<int P>
__device__ __forceinline__ float2 eval(const Hyp& H, const float2 (&X)[P][6], int p, float2 nlo){
      float2 e0=__fadd2_rn(X[p][3],make_float2(H.nt[0],H.nt[0]));
      float2 e1=__fadd2_rn(X[p][4],make_float2(H.nt[1],H.nt[1]));
      float2 e2=__fadd2_rn(X[p][5],make_float2(H.nt[2],H.nt[2]));
      #pragma unroll
      for(int j=0;j<3;++j){
        e0=__ffma2_rn(make_float2(H.nR[j],H.nR[j]),X[p][j],e0);
        e1=__ffma2_rn(make_float2(H.nR[3+j],H.nR[3+j]),X[p][j],e1);
        e2=__ffma2_rn(make_float2(H.nR[6+j],H.nR[6+j]),X[p][j],e2);
      }
      float2 s=__ffma2_rn(e0,e0,nlo); s=__ffma2_rn(e1,e1,s); s=__ffma2_rn(e2,e2,s);
      return s;
}
template<int P>
__device__ __noinline__ unsigned cold(const Hyp& H, const float2 (&X)[P][6], float2 nlo, float band, unsigned* wl, unsigned* wn, int h){
  unsigned adj=0;
  #pragma unroll
  for(int p=0;p<P;++p){ float2 s=eval<P>(H,X,p,nlo); float sv[2]={s.x,s.y};
    #pragma unroll
    for(int u=0;u<2;++u) if(fabsf(sv[u])<=band){ if(__float_as_uint(sv[u])>>31) adj += u?0xffff0000u:0xffffu; unsigned i=atomicAdd(wn,1u); wl[i]=h*64+p*2+u; } }
  return adj;
}
template<int P, int MODE>
__global__ void __launch_bounds__(704,1) flipk(const float2* __restrict__ rec, int nh, unsigned* __restrict__ out, float thr2, unsigned* wl, unsigned* wn){
  __shared__ float band_s[1024];
  __shared__ unsigned cnt_s[1024];
  for(int i=threadIdx.x;i<1024;i+=blockDim.x){band_s[i]=rec[i].x; cnt_s[i]=0;}
  __syncthreads();
  float2 X[P][6];
  #pragma unroll
  for(int p=0;p<P;++p)
  #pragma unroll
  for(int i=0;i<6;++i) X[p][i]=rec[(threadIdx.x*P+p)*6+i];
  const float2 nlo=make_float2(-thr2,-thr2);
  const bool lane0=(threadIdx.x&31)==0;
  #pragma unroll 2
  for(int h=0;h<nh;++h){
    const Hyp& H=chyp[h];
    const float band=band_s[h];
    unsigned pacc=0; float smin=CUDART_INF_F;
    #pragma unroll
    for(int p=0;p<P;++p){
      float2 s=eval<P>(H,X,p,nlo);
      pacc-=sign_words(s);
      smin=fminf(fminf(smin,fabsf(s.x)),fabsf(s.y));
    }
    if(MODE==0){ if(smin<=band){ pacc+=cold<P>(H,X,nlo,band,wl,wn,h);} }
    if(MODE==1){ if(__any_sync(0xffffffffu,smin<=band)){ if(smin<=band) pacc+=cold<P>(H,X,nlo,band,wl,wn,h);} }
    if(MODE==2){ if(__any_sync(0xffffffffu,smin<=band)){ pacc+=cold<P>(H,X,nlo,band,wl,wn,h);} }
    unsigned w=__reduce_add_sync(0xffffffffu,pacc);
    if(lane0) atomicAdd(&cnt_s[h],w);
  }
  __syncthreads();
  for(int i=threadIdx.x;i<1024;i+=blockDim.x) out[blockIdx.x*1024+i]=cnt_s[i];
}
template __global__ void flipk<3,0>(const float2*,int,unsigned*,float,unsigned*,unsigned*);
template __global__ void flipk<3,1>(const float2*,int,unsigned*,float,unsigned*,unsigned*);
template __global__ void flipk<3,2>(const float2*,int,unsigned*,float,unsigned*,unsigned*);
