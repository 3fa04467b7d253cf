// FFMA2 (fma.rn.f32x2) throughput and its mix with MUFU.EX2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ unsigned long long pk(float a, float b){unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(unsigned long long v, float&a, float&b){asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v));}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
template<int MODE> __global__ void k(float* out, int iters, float seed){
  unsigned long long v[8]; float f[16];
  #pragma unroll
  for(int i=0;i<8;i++){ v[i]=pk(seed*(i+threadIdx.x)*1e-3f, seed*i); }
  #pragma unroll
  for(int i=0;i<16;i++) f[i]=seed*(i+threadIdx.x)*1e-3f;
  unsigned long long m=pk(1.0001f,0.9999f), c=pk(0.5f,0.25f);
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<8;i++){
      if(MODE==0){ v[i]=fma2(v[i],m,c); }                       // 1 FFMA2 (=2 lane-FMAs)
      if(MODE==1){ v[i]=fma2(v[i],m,c); v[i]=fma2(v[i],m,c); v[i]=fma2(v[i],m,c); v[i]=fma2(v[i],m,c);
                   float a,b; upk(v[i],a,b); a=ex2(a); b=ex2(b); v[i]=pk(a,b);}   // 2 MUFU + 4 FFMA2 per 2 elements
      if(MODE==2){ f[2*i]=fmaf(f[2*i],1.0001f,0.5f); f[2*i+1]=fmaf(f[2*i+1],1.0001f,0.5f);} // 2 FFMA
    }
  }
  float s=0;
  #pragma unroll
  for(int i=0;i<8;i++){float a,b; upk(v[i],a,b); s+=a+b+f[2*i]+f[2*i+1];}
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name, int threads, int bps){
  int iters=4096; float* out; int nb=148*bps; cudaMalloc(&out, nb*threads*4);
  k<MODE><<<nb,threads>>>(out,16,0.5f); cudaDeviceSynchronize();
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); k<MODE><<<nb,threads>>>(out,iters,0.5f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  double ops=(double)nb*threads*iters*16;   // 16 scalar "elements" per iter
  printf("%-34s thr %4d x%d: %.3f ms  %.2f T elem/s  (per SM per clk @1.965GHz: %.2f)\n",name,threads,bps,ms,ops/ms/1e9, ops/(ms*1e-3)/148/1.965e9);
  cudaFree(out);
}
int main(){
  run<2>("ffma (1 per elem)",1024,2);
  run<0>("ffma2 (0.5 instr per elem)",1024,2);
  run<1>("2 mufu + 4 ffma2 per 2 elem",1024,2);
  run<1>("2 mufu + 4 ffma2 per 2 elem",512,1);
  run<1>("2 mufu + 4 ffma2 per 2 elem",256,1);
  run<1>("2 mufu + 4 ffma2 per 2 elem",128,1);
  return 0;
}
