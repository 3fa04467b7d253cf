// Where does the scan loop lose its MUFU rate?  One CTA of 8 warps per SM, 8 states per thread (the forward's tiling at
// d_state 64); features of the real loop are added one at a time.  Prints cycles per timestep per SM (MUFU floor: 128).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x){float y; asm("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ void sts(float* p, float v){ asm volatile("st.shared.f32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v)); }
// FEAT bit 0: (delta, delta*u) from shared memory (LDS.64 per step, 2 steps ahead); 1: B/C by broadcast LDS.128 (1 step
// ahead); 2: partial <h,C> stored to shared memory; 3: __syncthreads every 16 steps; 4: scalar FFMA for the 3-operand ops
template<int FEAT, int NPER> __global__ void __launch_bounds__(512,1) k(float* out, int steps, float seed){
  extern __shared__ __align__(16) float sm[];
  float2* dlu = reinterpret_cast<float2*>(sm);            // [64][32]
  float* Bs = sm + 64*32*2;                                // [64][64]
  float* Cs = Bs + 64*64;                                  // [64][64]
  float* yp = Cs + 64*64;                                  // [16 warps][16][32]
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  for(int i=threadIdx.x;i<64*32;i+=blockDim.x) dlu[i]=make_float2(0.01f+1e-5f*i, 1e-3f);
  for(int i=threadIdx.x;i<64*64;i+=blockDim.x){ Bs[i]=0.5f+1e-4f*i; Cs[i]=0.25f; }
  __syncthreads();
  float2 A2[NPER/2], h[NPER/2];
  #pragma unroll
  for(int j=0;j<NPER/2;j++){ A2[j]=make_float2(-1.f-j-w*seed, -1.5f-j); h[j]=make_float2(0.f,0.f); }
  float ysum=0.f;
  const float2* dl = dlu + lane;
  const float* Bf = Bs + (w*NPER)%64; const float* Cf = Cs + (w*NPER)%64;
  float2 dd_cur=dl[0], dd_nxt=dl[32], dd_n2;
  float2 Bc[NPER/2], Cc[NPER/2], Bn[NPER/2], Cn[NPER/2], a_cur[NPER/2], a_nxt[NPER/2];
  auto fetch=[&](int t, float2(&Bv)[NPER/2], float2(&Cv)[NPER/2]){
    #pragma unroll
    for(int q=0;q<NPER/4;q++){ float4 b=*reinterpret_cast<const float4*>(Bf+t*64+4*q), c=*reinterpret_cast<const float4*>(Cf+t*64+4*q);
      Bv[2*q]=make_float2(b.x,b.y); Bv[2*q+1]=make_float2(b.z,b.w); Cv[2*q]=make_float2(c.x,c.y); Cv[2*q+1]=make_float2(c.z,c.w);} };
  auto decay=[&](float2 dd, float2(&a)[NPER/2]){ float2 d2=make_float2(dd.x,dd.x);
    #pragma unroll
    for(int q=0;q<NPER/2;q++){ float2 g=__fmul2_rn(d2,A2[q]); a[q]=make_float2(ex2(g.x),ex2(g.y)); } };
  fetch(0,Bc,Cc); decay(dd_cur,a_cur);
  for(int s0=0;s0<steps;s0+=16){
    #pragma unroll
    for(int t=0;t<16;t++){
      if(FEAT&1){ dd_n2 = dl[((t+2)&63)*32]; } else { dd_n2 = make_float2(dd_cur.x+1e-7f, dd_cur.y); }
      if(FEAT&2){ fetch((t+1)&63,Bn,Cn); } else {
        #pragma unroll
        for(int q=0;q<NPER/2;q++){ Bn[q]=Cc[q]; Cn[q]=Bc[q]; } }
      decay(dd_nxt,a_nxt);
      float2 du2=make_float2(dd_cur.y,dd_cur.y), acc=make_float2(0.f,0.f);
      #pragma unroll
      for(int q=0;q<NPER/2;q++){
        float2 x=__fmul2_rn(du2,Bc[q]);
        if(FEAT&16){ h[q].x=fmaf(a_cur[q].x,h[q].x,x.x); h[q].y=fmaf(a_cur[q].y,h[q].y,x.y); acc.x=fmaf(h[q].x,Cc[q].x,acc.x); acc.y=fmaf(h[q].y,Cc[q].y,acc.y);} 
        else { h[q]=__ffma2_rn(a_cur[q],h[q],x); acc=__ffma2_rn(h[q],Cc[q],acc); }
      }
      if(FEAT&4) sts(yp+(w*16+t)*32+lane, acc.x+acc.y); else ysum+=acc.x+acc.y;
      dd_cur=dd_nxt; dd_nxt=dd_n2;
      #pragma unroll
      for(int q=0;q<NPER/2;q++){ Bc[q]=Bn[q]; Cc[q]=Cn[q]; a_cur[q]=a_nxt[q]; }
    }
    if(FEAT&8) __syncthreads();
  }
  float s=ysum;
  #pragma unroll
  for(int q=0;q<NPER/2;q++) s+=h[q].x+h[q].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s+yp[threadIdx.x];
}
template<int FEAT,int NPER> void run(const char* name){
  const int threads=64/NPER*32, steps=4096; float* out; cudaMalloc(&out,148*threads*4);
  size_t smem=(64*32*2+64*64*2+16*16*32)*4;
  cudaFuncSetAttribute(k<FEAT,NPER>, cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem);
  k<FEAT,NPER><<<148,threads,smem>>>(out,64,0.f); cudaDeviceSynchronize();
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); k<FEAT,NPER><<<148,threads,smem>>>(out,steps,0.f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  printf("NPER %d feat %2d %-52s %.3f ms  %.1f clk/step @1.965GHz  err=%s\n",NPER,FEAT,name,ms,ms*1e-3*1.965e9/steps,cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main(){
  run<0,8>("math only (registers)");
  run<1,8>("+ LDS.64 (delta, delta*u)");
  run<2,8>("+ LDS.128 B/C broadcast");
  run<3,8>("+ both loads");
  run<7,8>("+ both loads + STS partial");
  run<15,8>("+ loads + STS + barrier/16 steps");
  run<16,8>("math only, scalar FFMA");
  run<31,8>("everything, scalar FFMA");
  run<0,4>("math only (registers)");
  run<3,4>("+ both loads");
  run<15,4>("+ loads + STS + barrier/16 steps");
  run<31,4>("everything, scalar FFMA");
  return 0;
}
