// common.cuh — shared device helpers for the sm_100a Mamba kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mamba_b200.h"

namespace mb {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- error plumbing (abi.cu) ------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
void count_launch(int n = 1);

// ---- dtype traits ---------------------------------------------------------------------------
template <typename T>
struct IO;
template <>
struct IO<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ float cvt(float v) { return v; }
};
template <>
struct IO<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float cvt(__nv_bfloat16 v) { return __bfloat162float(v); }
};

// ---- math ---------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// F.softplus with beta=1, threshold=20 (torch default): x if x > 20 else log1p(exp(x)).
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
// d softplus / dx = sigmoid(x) (1 above the threshold, as torch's softplus_backward).
__device__ __forceinline__ float softplus_grad_f(float x) {
  return x > 20.f ? 1.f : 1.f / (1.f + expf(-x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

// ---- async copies (LDGSTS) ----------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Streaming 128-bit global store (outputs are written once and not re-read by this kernel).
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Load a [rows x cols] tile (element type T, row stride `ls` elements in global memory) into a
// dense shared tile [rows][cols_s].  Rows >= rows_valid and columns >= cols_valid are zero-filled.
// `vec_ok` (block-uniform) says 16-byte cp.async is legal for every fully in-range vector (base
// pointer, batch stride and row stride are multiples of 16 bytes); vectors that straddle
// cols_valid, and everything when !vec_ok, take a guarded scalar path.
template <typename T>
__device__ __forceinline__ void load_tile_async(T* __restrict__ s, int cols_s, const T* __restrict__ g, int64_t ls,
                                                int rows, int rows_valid, int cols_valid, bool vec_ok, int tid,
                                                int nthreads) {
  constexpr int VEC = 16 / sizeof(T);
  const int vpr = cols_s / VEC;  // vectors per row (cols_s is a multiple of VEC)
  const int nvec = rows * vpr;
  for (int i = tid; i < nvec; i += nthreads) {
    const int r = i / vpr;
    const int c = (i - r * vpr) * VEC;
    T* dst = s + r * cols_s + c;
    if (r < rows_valid && vec_ok && c + VEC <= cols_valid) {
      cp_async16(dst, g + (int64_t)r * ls + c);
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        T v = T(0.f);
        if (r < rows_valid && c + k < cols_valid) v = g[(int64_t)r * ls + c + k];
        dst[k] = v;
      }
    }
  }
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace mb
