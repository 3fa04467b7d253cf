// FMA-pipe throughput with REGISTER operands (the earlier microbenchmarks reused one multiplier / addend register pair,
// which the operand-reuse cache serves): scalar FFMA and packed FFMA2 / FMUL2 with 3 (2) distinct register operands.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v, float&a, float&b){asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v));}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
__device__ __forceinline__ u64 mul2(u64 a, u64 b){u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ float fma1(float a, float b, float c){float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;":"=f"(d):"f"(a),"f"(b),"f"(c)); return d;}
__device__ __forceinline__ float ex2(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
// MODE 0: FFMA2 d_i = a_i * b_i + d_i  (3 distinct register pairs per op, 8 independent chains)
// MODE 1: FFMA   d_i = a_i * b_i + d_i  (16 chains)
// MODE 2: FMUL2  d_i = d_i * b_i
// MODE 3: FFMA2  d_i = d_i * m + c (shared operands: the old benchmark)
// MODE 4: scan step mix per state pair: FMUL2, 2 MUFU, FMUL2, FFMA2, FFMA2  (distinct registers)
// MODE 5: as 4 without the MUFU
template<int MODE> __global__ void k(float* out, int iters, float seed){
  u64 a[8], b[8], d[8]; float fa[16], fb[16], fd[16];
  #pragma unroll
  for(int i=0;i<8;i++){ a[i]=pk(1.f+seed*(i+threadIdx.x)*1e-6f, 1.f-seed*i*1e-6f); b[i]=pk(1.f-seed*i*1e-6f, 1.f+seed*i*1e-7f); d[i]=pk(seed*i, seed); }
  #pragma unroll
  for(int i=0;i<16;i++){ fa[i]=1.f+seed*(i+threadIdx.x)*1e-6f; fb[i]=seed*i*1e-6f; fd[i]=seed*i; }
  u64 m=pk(1.0001f,0.9999f), c=pk(0.5f,0.25f);
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<8;i++){
      if(MODE==0) d[i]=fma2(a[i],b[i],d[i]);
      if(MODE==1){ fd[2*i]=fma1(fa[2*i],fb[2*i],fd[2*i]); fd[2*i+1]=fma1(fa[2*i+1],fb[2*i+1],fd[2*i+1]); }
      if(MODE==2) d[i]=mul2(d[i],b[i]);
      if(MODE==3) d[i]=fma2(d[i],m,c);
      if(MODE==4 || MODE==5){
        u64 g=mul2(a[i],m); float g0,g1; upk(g,g0,g1);
        if(MODE==4){ g0=ex2(g0); g1=ex2(g1);} 
        u64 x=mul2(b[i],c);
        d[i]=fma2(pk(g0,g1),d[i],x);
        a[i]=fma2(d[i],b[i],a[i]);
      }
    }
  }
  float s=0;
  #pragma unroll
  for(int i=0;i<8;i++){float x,y; upk(d[i],x,y); s+=x+y+fd[2*i]+fd[2*i+1]; upk(a[i],x,y); s+=x+y;}
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name, int threads, int bps, double lane_ops_per_iter){
  int iters=4096; float* out; int nb=148*bps; cudaMalloc(&out, nb*threads*4);
  k<MODE><<<nb,threads>>>(out,16,0.5f); cudaDeviceSynchronize();
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); k<MODE><<<nb,threads>>>(out,iters,0.5f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  double ops=(double)nb*threads*iters*lane_ops_per_iter;
  printf("%-44s thr %4d x%d: %.3f ms  fma-pipe lane-ops per SM per clk @1.965GHz: %.2f\n",name,threads,bps,ms, ops/(ms*1e-3)/148/1.965e9);
  cudaFree(out);
}
int main(){
  for (int thr : {256, 512, 1024}) {
    run<0>("FFMA2 3 distinct reg pairs",thr,1,16);
    run<1>("FFMA  3 distinct regs",thr,1,16);
    run<2>("FMUL2 2 distinct reg pairs",thr,1,16);
    run<3>("FFMA2 shared multiplier/addend",thr,1,16);
    run<4>("scan mix (4 packed + 2 MUFU per pair)",thr,1,64);
    run<5>("scan mix without MUFU (4 packed per pair)",thr,1,64);
  }
  return 0;
}
