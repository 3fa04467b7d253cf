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

__device__ __forceinline__ void st_cs_f4(float* p, float4 v);

// A[d, n] as the scans use it: the stored value, or -exp(stored) when the caller passes A_log (MAMBA_FLAG_A_IS_LOG;
// reference: A = -torch.exp(self.A_log.float()), simple_mamba.pyc @L270)
__device__ __forceinline__ float load_A(const float* A, int64_t i, int flags) {
  const float v = A[i];
  return (flags & MAMBA_FLAG_A_IS_LOG) ? -expf(v) : v;
}

// ---- fast transcendental helpers (MUFU ex2 / lg2 / rcp; relative error ~1e-7, far inside rtol 1e-4) ----------
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// softplus with torch's threshold (x > 20 -> x).  For x < -4 the series of log1p(e) avoids the cancellation
// of forming 1 + e in fp32.
__device__ __forceinline__ float softplus_fast(float x) {
  const float e = ex2_approx(x * kLog2e);
  const float series = e * fmaf(e, fmaf(e, fmaf(e, -0.25f, 0.33333334f), -0.5f), 1.f);
  const float full = lg2_approx(1.f + e) * kLn2;
  const float r = x < -4.f ? series : full;
  return x > 20.f ? x : r;
}
__device__ __forceinline__ float silu_fast(float x) { return x * rcp_approx(1.f + ex2_approx(-x * kLog2e)); }

// Shared-memory store issued through asm so that nvvm does not order later shared loads behind it.
__device__ __forceinline__ void sts_f32(float* p, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "f"(v));
}
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// 4 consecutive elements of T <-> 4 floats (128-bit / 64-bit accesses)
template <typename T>
struct V4;
template <>
struct V4<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  static __device__ __forceinline__ void st_global(float* p, const float (&v)[4]) { st_cs_f4(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <>
struct V4<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    v[0] = __low2float(a), v[1] = __high2float(a), v[2] = __low2float(b), v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void st_global(__nv_bfloat16* p, const float (&v)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<const uint32_t*>(&a), t.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};



// 8 consecutive bf16 -> 8 floats through one 16-byte shared load and two 16-byte shared stores
__device__ __forceinline__ void cvt8_bf16_f32(const __nv_bfloat16* src, float* dst) {
  const uint4 r = *reinterpret_cast<const uint4*>(src);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  float o[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[2 * k] = __uint_as_float(w[k] << 16);
    o[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
  }
  *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
}
__device__ __forceinline__ void cvt8_bf16_f32(const float*, float*) {}  // never called (fp32 tiles are read in place)

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

// Fixed-order column sums of a [nrows][ncols] fp32 partial buffer, for blocks of (32, 8) threads: thread (x, y) adds
// rows y, y+8, ... of column blockIdx.x*32 + x (4 independent loads in flight), the 8 row groups are combined in
// shared memory in a fixed order.  Returns the total on the threads with y == 0 (others: undefined).
__device__ __forceinline__ float colsum_32x8(const float* __restrict__ ws, int nrows, int64_t row_stride, int col, bool ok) {
  __shared__ float part[8][33];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (ok) {
    int r = threadIdx.y;
    for (; r + 24 < nrows; r += 32) {
      a0 += ws[(int64_t)r * row_stride + col];
      a1 += ws[(int64_t)(r + 8) * row_stride + col];
      a2 += ws[(int64_t)(r + 16) * row_stride + col];
      a3 += ws[(int64_t)(r + 24) * row_stride + col];
    }
    for (; r < nrows; r += 8) a0 += ws[(int64_t)r * row_stride + col];
  }
  part[threadIdx.y][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  float s = 0.f;
  if (threadIdx.y == 0) {
#pragma unroll
    for (int y = 0; y < 8; ++y) s += part[y][threadIdx.x];
  }
  return s;
}

// ---- programmatic dependent launch (decode chain) -----------------------------------------------------------------
// The decode step is a chain of ~40 short kernels, each of which streams weights that do not depend on its
// predecessor.  Launched with the programmatic-stream-serialization attribute, a kernel starts while the previous one
// is still running: it signals its own dependents at once, prefetches its weights / state into L2, and only then
// waits (griddepcontrol.wait) for the predecessor's results.  Everything before the wait must be independent of
// earlier kernels of the chain.  In a kernel launched without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
bool pdl_enabled();  // abi.cu: MAMBA_B200_PDL == "1" (off by default)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute is per (function, DEVICE): the cache is keyed by
// the current device so that one host thread driving several GPUs configures each of them (one `static thread_local
// SmemConfig` per kernel instantiation at the call site).
constexpr int kMaxDevices = 64;
struct SmemConfig {
  size_t bytes[kMaxDevices] = {};
};
template <typename K>
static inline int ensure_dynamic_smem(K kern, size_t smem, SmemConfig& cfg, const char* what) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = -1;
  if (dev >= 0 && smem <= cfg.bytes[dev]) return MAMBA_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(MAMBA_ELAUNCH, "%s: cudaFuncSetAttribute(%zu B): %s", what, smem, cudaGetErrorString(e));
  if (dev >= 0) cfg.bytes[dev] = smem;
  return MAMBA_OK;
}

}  // namespace mb
