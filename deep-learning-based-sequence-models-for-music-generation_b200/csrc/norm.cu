// norm.cu — RMSNorm with fused residual add, forward and backward, sm_100a.
//
// Replaces RMSNorm.forward (reference simple_mamba.pyc @L346:
//     x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight)
// and the `+ x` of ResidualBlock.forward (@L179).  One warp per row; a lane owns groups of 4 consecutive
// columns (128-bit loads/stores, a warp instruction covers 512 contiguous bytes in fp32) and the row
// stays in registers between the reduction and the scaling, so every tensor is read exactly once.
// Two element types: T for the mixer side (x, y, dy, dx), TR for the residual stream.
#include "common.cuh"

namespace mb {

struct NormParams {
  int64_t rows;
  int dim;
  float eps;
  const void *x, *residual, *dy, *dres;
  void *y, *resid_out, *dx, *dres_out;
  const float* w;
  float *rstd, *dw, *ws;
  int nblocks;
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- 4-wide row access (V = 4: vector path, V = 1: scalar path for odd dims / unaligned rows) ------
template <typename T, int V>
struct Row;
template <>
struct Row<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Row<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    v[0] = __low2float(a), v[1] = __high2float(a), v[2] = __low2float(b), v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<const uint32_t*>(&a), t.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};
template <typename T>
struct Row<T, 1> {
  static __device__ __forceinline__ void ld(const T* p, float (&v)[1]) { v[0] = IO<T>::ld(p); }
  static __device__ __forceinline__ void st(T* p, const float (&v)[1]) { IO<T>::st(p, v[0]); }
};

// Column of group k of this lane: c = (lane + 32 * k) * V.
template <typename T, typename TR, int V, int G>
__global__ void __launch_bounds__(256) rmsnorm_fwd_kernel(const NormParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const T* x = p.x ? static_cast<const T*>(p.x) + row * p.dim : nullptr;
  const TR* res = p.residual ? static_cast<const TR*>(p.residual) + row * p.dim : nullptr;
  TR* ro = p.resid_out ? static_cast<TR*>(p.resid_out) + row * p.dim : nullptr;
  T* y = static_cast<T*>(p.y) + row * p.dim;
  float v[G][V];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < G; ++k) {
    const int c = (lane + 32 * k) * V;
#pragma unroll
    for (int e = 0; e < V; ++e) v[k][e] = 0.f;
    if (c < p.dim) {
      if (x) Row<T, V>::ld(x + c, v[k]);
      if (res) {
        float r[V];
        Row<TR, V>::ld(res + c, r);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          v[k][e] += r[e];
          if (x) v[k][e] = IO<TR>::cvt(TR(v[k][e]));  // normalise exactly what the stream stores
        }
      }
      if (ro) Row<TR, V>::st(ro + c, v[k]);
    }
#pragma unroll
    for (int e = 0; e < V; ++e) ss = fmaf(v[k][e], v[k][e], ss);
  }
  ss = wsum(ss);
  const float rstd = rsqrtf(ss / (float)p.dim + p.eps);
  if (lane == 0 && p.rstd) p.rstd[row] = rstd;
#pragma unroll
  for (int k = 0; k < G; ++k) {
    const int c = (lane + 32 * k) * V;
    if (c < p.dim) {
      float w[V], o[V];
      Row<float, V>::ld(p.w + c, w);
#pragma unroll
      for (int e = 0; e < V; ++e) o[e] = v[k][e] * rstd * w[e];
      Row<T, V>::st(y + c, o);
    }
  }
}

// Backward.  Blocks stride over rows; dweight is accumulated per owned column in registers, reduced across
// the block's warps in shared memory and written as one partial row per block (fixed-order finalize).
template <typename T, typename TR, int V, int G>
__global__ void __launch_bounds__(256) rmsnorm_bwd_kernel(const NormParams p) {
  extern __shared__ float sm[];  // [nwarps][dim]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float dwacc[G][V], wv[G][V];
#pragma unroll
  for (int k = 0; k < G; ++k) {
    const int c = (lane + 32 * k) * V;
#pragma unroll
    for (int e = 0; e < V; ++e) dwacc[k][e] = 0.f, wv[k][e] = 0.f;
    if (c < p.dim) Row<float, V>::ld(p.w + c, wv[k]);
  }
  for (int64_t row = (int64_t)blockIdx.x * nw + warp; row < p.rows; row += (int64_t)gridDim.x * nw) {
    const TR* r = static_cast<const TR*>(p.residual) + row * p.dim;
    const T* dy = static_cast<const T*>(p.dy) + row * p.dim;
    const TR* dres = p.dres ? static_cast<const TR*>(p.dres) + row * p.dim : nullptr;
    const float rstd = p.rstd[row];
    // Rows of up to 1024 columns: the incoming residual gradient is fetched with the other two operands (one DRAM
    // round trip per row, not two); wider rows load it late to stay inside the register budget.
    constexpr bool kEarly = G * V <= 32;
    float rh[G][V], g[G][V], dr[kEarly ? G : 1][V];
    float dot = 0.f;
    if constexpr (kEarly) {
#pragma unroll
      for (int k = 0; k < G; ++k) {
        const int c = (lane + 32 * k) * V;
#pragma unroll
        for (int e = 0; e < V; ++e) dr[k][e] = 0.f;
        if (dres && c < p.dim) Row<TR, V>::ld(dres + c, dr[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const int c = (lane + 32 * k) * V;
      float rv[V], dv[V];
#pragma unroll
      for (int e = 0; e < V; ++e) rv[e] = 0.f, dv[e] = 0.f;
      if (c < p.dim) {
        Row<TR, V>::ld(r + c, rv);
        Row<T, V>::ld(dy + c, dv);
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        rh[k][e] = rv[e] * rstd;
        g[k][e] = dv[e] * wv[k][e];
        dwacc[k][e] = fmaf(dv[e], rh[k][e], dwacc[k][e]);
        dot = fmaf(g[k][e], rh[k][e], dot);
      }
    }
    dot = wsum(dot) / (float)p.dim;
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const int c = (lane + 32 * k) * V;
      if (c < p.dim) {
        float o[V];
#pragma unroll
        for (int e = 0; e < V; ++e) o[e] = rstd * (g[k][e] - rh[k][e] * dot);
        if constexpr (kEarly) {
#pragma unroll
          for (int e = 0; e < V; ++e) o[e] += dr[k][e];
        } else if (dres) {
          float dl[V];
          Row<TR, V>::ld(dres + c, dl);
#pragma unroll
          for (int e = 0; e < V; ++e) o[e] += dl[e];
        }
        if (p.dx) Row<T, V>::st(static_cast<T*>(p.dx) + row * p.dim + c, o);
        if (p.dres_out) Row<TR, V>::st(static_cast<TR*>(p.dres_out) + row * p.dim + c, o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < G; ++k) {
    const int c = (lane + 32 * k) * V;
    if (c < p.dim) Row<float, V>::st(sm + warp * p.dim + c, dwacc[k]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.dim; c += blockDim.x) {
    float s = 0.f;
    for (int ww = 0; ww < nw; ++ww) s += sm[ww * p.dim + c];
    p.ws[(int64_t)blockIdx.x * p.dim + c] = s;
  }
}

// grid: ceil(dim / 32) blocks of (32, 8) threads
__global__ void rmsnorm_bwd_finalize_kernel(const NormParams p) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < p.dim;
  const float s = colsum_32x8(p.ws, p.nblocks, p.dim, c, ok);
  if (ok && threadIdx.y == 0) p.dw[c] = s;
}

static int norm_blocks(int64_t rows) {
  const int64_t want = (rows + 7) / 8;
  return (int)(want < 2 * kNumSMs ? want : 2 * kNumSMs);
}

template <typename T, typename TR, int V, int G>
static int norm_launch_g(const NormParams& p, bool bwd, cudaStream_t st) {
  if (!bwd) {
    rmsnorm_fwd_kernel<T, TR, V, G><<<(unsigned)((p.rows + 7) / 8), 256, 0, st>>>(p);
    count_launch();
    return check_launch("rmsnorm_fwd");
  }
  const size_t smem = (size_t)8 * p.dim * 4;
  auto kern = rmsnorm_bwd_kernel<T, TR, V, G>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(MAMBA_ELAUNCH, "rmsnorm_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  kern<<<p.nblocks, 256, smem, st>>>(p);
  count_launch();
  int rc = check_launch("rmsnorm_bwd");
  if (rc) return rc;
  rmsnorm_bwd_finalize_kernel<<<ceil_div(p.dim, 32), dim3(32, 8), 0, st>>>(p);
  count_launch();
  return check_launch("rmsnorm_bwd_finalize");
}

template <typename T, typename TR>
static int norm_launch(const NormParams& p, bool bwd, bool vec, cudaStream_t st) {
  if (vec) {
    const int groups = ceil_div(p.dim, 128);  // 4-column groups per lane
    if (groups <= 2) return norm_launch_g<T, TR, 4, 2>(p, bwd, st);
    if (groups <= 4) return norm_launch_g<T, TR, 4, 4>(p, bwd, st);
    if (groups <= 8) return norm_launch_g<T, TR, 4, 8>(p, bwd, st);
    if (groups <= 16) return norm_launch_g<T, TR, 4, 16>(p, bwd, st);
    return set_error(MAMBA_ESIZE, "rmsnorm: dim %d above 2048", p.dim);
  }
  const int groups = ceil_div(p.dim, 32);
  if (groups <= 8) return norm_launch_g<T, TR, 1, 8>(p, bwd, st);
  if (groups <= 32) return norm_launch_g<T, TR, 1, 32>(p, bwd, st);
  if (groups <= 64) return norm_launch_g<T, TR, 1, 64>(p, bwd, st);
  return set_error(MAMBA_ESIZE, "rmsnorm: dim %d above 2048", p.dim);
}

static int norm_common(const MambaNormArgs* a, bool bwd, void* stream) {
  if (!a || a->struct_size != (int32_t)sizeof(MambaNormArgs))
    return set_error(MAMBA_EINVAL, "rmsnorm: bad args pointer or struct_size");
  if (a->rows <= 0 || a->dim <= 0) return set_error(MAMBA_EINVAL, "rmsnorm: rows/dim must be positive");
  const bool f32 = a->dtype == MAMBA_F32 && a->resid_dtype == MAMBA_F32;
  const bool mixed = a->dtype == MAMBA_BF16 && a->resid_dtype == MAMBA_F32;
  const bool bf16 = a->dtype == MAMBA_BF16 && a->resid_dtype == MAMBA_BF16;
  if (!f32 && !mixed && !bf16)
    return set_error(MAMBA_EDTYPE, "rmsnorm: unsupported (dtype, resid_dtype) = (%d, %d)", a->dtype, a->resid_dtype);
  if (!a->weight) return set_error(MAMBA_EINVAL, "rmsnorm: null weight");
  NormParams p{};
  p.rows = a->rows, p.dim = a->dim, p.eps = a->eps;
  p.x = a->x, p.residual = a->residual, p.dy = a->dy, p.dres = a->dresid_in;
  p.y = a->y, p.resid_out = a->resid_out, p.dx = a->dx, p.dres_out = a->dresid_out;
  p.w = a->weight, p.rstd = a->rstd, p.dw = a->dweight;
  p.nblocks = norm_blocks(a->rows);
  bool vec = (a->dim % 4 == 0) && aligned16(a->weight);
  auto al = [&](const void* q) { return q == nullptr || aligned16(q); };
  if (!bwd) {
    if (!a->y) return set_error(MAMBA_EINVAL, "rmsnorm_fwd: null y");
    if (!a->x && !a->residual) return set_error(MAMBA_EINVAL, "rmsnorm_fwd: x and residual both NULL");
    vec = vec && al(a->x) && al(a->residual) && al(a->y) && al(a->resid_out);
  } else {
    if (!a->dy || !a->residual || !a->dweight || !a->rstd)
      return set_error(MAMBA_EINVAL, "rmsnorm_bwd: null dy/residual/dweight/rstd");
    if (!a->dx && !a->dresid_out) return set_error(MAMBA_EINVAL, "rmsnorm_bwd: dx and dresid_out both NULL");
    const size_t need = mamba_rmsnorm_bwd_workspace_bytes(a->rows, a->dim);
    if (!a->workspace || a->workspace_bytes < need)
      return set_error(MAMBA_ESIZE, "rmsnorm_bwd: workspace %zu B < required %zu B", a->workspace_bytes, need);
    p.ws = static_cast<float*>(a->workspace);
    vec = vec && al(a->residual) && al(a->dy) && al(a->dresid_in) && al(a->dx) && al(a->dresid_out);
  }
  // rows are dim elements apart: 8-byte (bf16) / 16-byte (fp32) vector alignment of every row needs dim % 4 == 0
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32) return norm_launch<float, float>(p, bwd, vec, st);
  if (mixed) return norm_launch<__nv_bfloat16, float>(p, bwd, vec, st);
  return norm_launch<__nv_bfloat16, __nv_bfloat16>(p, bwd, vec, st);
}

}  // namespace mb

extern "C" size_t mamba_rmsnorm_bwd_workspace_bytes(int64_t rows, int dim) {
  if (rows <= 0 || dim <= 0) return 0;
  return (size_t)4 * mb::norm_blocks(rows) * dim;
}
extern "C" int mamba_rmsnorm_fwd(const MambaNormArgs* a, void* stream) { return mb::norm_common(a, false, stream); }
extern "C" int mamba_rmsnorm_bwd(const MambaNormArgs* a, void* stream) { return mb::norm_common(a, true, stream); }
