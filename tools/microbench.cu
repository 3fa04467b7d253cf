// Pipe-throughput microbenchmarks on the B200 SM (ex2 on the MUFU pipe, FFMA, mixes) — design inputs for the scan kernels.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ unsigned ex2h2(unsigned x){unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;":"=r"(y):"r"(x)); return y;}
__device__ __forceinline__ unsigned ex2b2(unsigned x){unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;":"=r"(y):"r"(x)); return y;}

template<int MODE> __global__ void k(float* out, int iters, float seed){
  float v[16]; unsigned w[16];
  #pragma unroll
  for(int i=0;i<16;i++){ v[i]=seed*(i+threadIdx.x)*1e-3f; w[i]=__float_as_uint(v[i]); }
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<16;i++){
      if(MODE==0) v[i]=ex2(v[i]);                     // MUFU only (dependent chain per i, 16 independent)
      if(MODE==1) v[i]=fmaf(v[i],1.0001f,0.5f);       // FFMA only
      if(MODE==2){ v[i]=ex2(v[i]); v[i]=fmaf(v[i],0.999f,-1.f); v[i]=fmaf(v[i],0.5f,0.1f); v[i]=fmaf(v[i],0.5f,0.1f); v[i]=fmaf(v[i],0.5f,0.1f);} // 1 MUFU + 4 FMA
      if(MODE==3) w[i]=ex2h2(w[i]);
      if(MODE==4) w[i]=ex2b2(w[i]);
      if(MODE==5){ v[i]=ex2(v[i]); 
        #pragma unroll
        for(int q=0;q<7;q++) v[i]=fmaf(v[i],0.5f,0.1f);} // 1 MUFU + 7 FMA
    }
  }
  float s=0; 
  #pragma unroll
  for(int i=0;i<16;i++) s+=v[i]+__uint_as_float(w[i]);
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name, int ops_per_iter_per_thread, int blocks_per_sm, int threads){
  int iters=4096; float* out; int nb=148*blocks_per_sm; cudaMalloc(&out, nb*threads*4);
  k<MODE><<<nb,threads>>>(out,16,0.5f); cudaDeviceSynchronize();
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); k<MODE><<<nb,threads>>>(out,iters,0.5f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  double ops=(double)nb*threads*iters*16; 
  printf("%-28s blocks/SM %d thr %4d: %.3f ms  %.2f Tlane-iter/s  (per SM per clk @1.965GHz: %.2f)\n",name,blocks_per_sm,threads,ms,ops/ms/1e9, ops/(ms*1e-3)/148/1.965e9);
  cudaFree(out);
}
int main(){
  for(int thr: {128,256,512,1024}){
    run<0>("ex2.f32", 1, 1, thr);
  }
  run<0>("ex2.f32", 1, 2, 1024);
  run<1>("ffma", 1, 2, 1024);
  run<2>("ex2+4ffma", 1, 2, 1024);
  run<5>("ex2+7ffma", 1, 2, 1024);
  run<3>("ex2.f16x2 (2 exps/lane)", 1, 2, 1024);
  run<4>("ex2.bf16x2 (2 exps/lane)", 1, 2, 1024);
  run<2>("ex2+4ffma", 1, 1, 128);
  run<2>("ex2+4ffma", 1, 1, 256);
  run<2>("ex2+4ffma", 1, 1, 512);
  return 0;
}
