// step.cu — single-token recurrent decode step, sm_100a.
//
// The reference has no recurrent step: scripts/generate.py:26-31 re-runs the full model on a sliding
// window for every new token.  These kernels evaluate MambaBlock's recurrence (simple_mamba.pyc
// @L233-241 conv+SiLU, @L276 dt_proj+softplus, @L310-333 scan, @L241 gate) for ONE new position with
// carried state, so a generated token costs O(1) in the context length.
//   mamba_conv_step : thread per (b, d); K-deep shift register in conv_state[b, d, :].
//   mamba_ssm_step  : warp per channel d, lanes over the state axis (coalesced on ssm_state[b, d, :])
//                     and over dt_rank for the fused dt_proj row-dot; loops over the (small) batch so
//                     dt_proj's row is read once.
#include "common.cuh"

namespace mb {

struct StepParams {
  int B, D, N, K, R, flags;
  const void *x, *dt_in, *Bv, *Cv, *z;
  void *conv_state, *xc, *y;
  int64_t x_bs, xc_bs, dt_in_bs, Bv_bs, Cv_bs, z_bs, y_bs;
  const float *cw, *cb, *dtw, *dtb, *A, *Dv;
  float* h;
};

template <typename T>
__global__ void conv_step_kernel(const StepParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)p.B * p.D) return;
  const int b = (int)(i / p.D), d = (int)(i % p.D);
  T* st = static_cast<T*>(p.conv_state) + i * p.K;
  float acc = p.cb ? p.cb[d] : 0.f;
  const float xn = IO<T>::ld(static_cast<const T*>(p.x) + (int64_t)b * p.x_bs + d);
  for (int k = 0; k < p.K; ++k) {
    const float v = (k + 1 < p.K) ? IO<T>::ld(st + k + 1) : xn;
    IO<T>::st(st + k, v);
    // round-trip through T so that the value used equals the value stored (bf16 state)
    acc = fmaf(p.cw[(int64_t)d * p.K + k], IO<T>::ld(st + k), acc);
  }
  IO<T>::st(static_cast<T*>(p.xc) + (int64_t)b * p.xc_bs + d, silu_f(acc));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256) ssm_step_kernel(const StepParams p) {
  const int lane = threadIdx.x & 31;
  const int d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (d >= p.D) return;
  const bool has_z = p.flags & MAMBA_FLAG_HAS_Z;
  const float bias = p.dtb ? p.dtb[d] : 0.f;
  const float Dd = (p.flags & MAMBA_FLAG_HAS_D) ? p.Dv[d] : 0.f;
  for (int b = 0; b < p.B; ++b) {
    // delta = softplus(dt_proj.weight[d, :] . dt_in[b, :] + bias)
    float dot = 0.f;
    const T* dti = static_cast<const T*>(p.dt_in) + (int64_t)b * p.dt_in_bs;
    for (int r = lane; r < p.R; r += 32) dot = fmaf(p.dtw[(int64_t)d * p.R + r], IO<T>::ld(dti + r), dot);
    dot = warp_sum(dot) + bias;
    const float delta = (p.flags & MAMBA_FLAG_DELTA_SOFTPLUS) ? softplus_f(dot) : dot;
    const float xc = IO<T>::ld(static_cast<const T*>(p.xc) + (int64_t)b * p.xc_bs + d);
    const float du = delta * xc;
    float* h = p.h + ((int64_t)b * p.D + d) * p.N;
    const T* Bv = static_cast<const T*>(p.Bv) + (int64_t)b * p.Bv_bs;
    const T* Cv = static_cast<const T*>(p.Cv) + (int64_t)b * p.Cv_bs;
    float y = 0.f;
    for (int n = lane; n < p.N; n += 32) {
      const float a = expf(delta * p.A[(int64_t)d * p.N + n]);
      const float hn = fmaf(a, h[n], du * IO<T>::ld(Bv + n));
      h[n] = hn;
      y = fmaf(hn, IO<T>::ld(Cv + n), y);
    }
    y = warp_sum(y);
    if (lane == 0) {
      y = fmaf(Dd, xc, y);
      if (has_z) y *= silu_f(IO<T>::ld(static_cast<const T*>(p.z) + (int64_t)b * p.z_bs + d));
      IO<T>::st(static_cast<T*>(p.y) + (int64_t)b * p.y_bs + d, y);
    }
  }
}

// Bandwidth-shaped variant (N % 4 == 0, R % 4 == 0, 16-byte aligned rows): a HALF-warp owns one (b, d) pair, so a
// warp instruction touches two full 256-byte state rows with 128-bit accesses; the fused dt_proj row-dot and the
// <h, C> reduction are 4 butterfly steps inside the half-warp.  All (b, d) pairs are independent, so the grid
// covers them all at once instead of looping over the batch.
template <typename T>
__global__ void __launch_bounds__(256) ssm_step_fast_kernel(const StepParams p) {
  const int hl = threadIdx.x & 15;
  const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4);
  pdl_launch_dependents();  // decode chain (common.cuh)
  if (item < (int64_t)p.B * p.D && hl < 2) prefetch_l2(p.h + item * p.N + hl * 32);  // this pair's state row (<= 256 B)
  pdl_wait();
  if (item >= (int64_t)p.B * p.D) return;  // whole half-warps leave together; shuffles below use per-half masks
  const unsigned mask = 0xffffu << (threadIdx.x & 16);
  const int b = (int)(item / p.D), d = (int)(item - (int64_t)b * p.D);
  float dot = 0.f;
  {
    const float* w = p.dtw + (int64_t)d * p.R;
    const T* x = static_cast<const T*>(p.dt_in) + (int64_t)b * p.dt_in_bs;
    for (int r4 = hl; r4 < (p.R >> 2); r4 += 16) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + r4);
      float xv[4];
      V4<T>::ld(x + 4 * r4, xv);
      dot = fmaf(wv.x, xv[0], fmaf(wv.y, xv[1], fmaf(wv.z, xv[2], fmaf(wv.w, xv[3], dot))));
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(mask, dot, o);
  }
  dot += p.dtb ? p.dtb[d] : 0.f;
  const float delta = (p.flags & MAMBA_FLAG_DELTA_SOFTPLUS) ? softplus_fast(dot) : dot;
  const float xc = IO<T>::ld(static_cast<const T*>(p.xc) + (int64_t)b * p.xc_bs + d);
  const float du = delta * xc, dl2 = delta * kLog2e;
  float* h = p.h + ((int64_t)b * p.D + d) * p.N;
  const float* A = p.A + (int64_t)d * p.N;
  const T* Bv = static_cast<const T*>(p.Bv) + (int64_t)b * p.Bv_bs;
  const T* Cv = static_cast<const T*>(p.Cv) + (int64_t)b * p.Cv_bs;
  float y = 0.f;
  for (int n4 = hl; n4 < (p.N >> 2); n4 += 16) {
    const float4 a4 = __ldg(reinterpret_cast<const float4*>(A) + n4);
    float4 h4 = reinterpret_cast<float4*>(h)[n4];
    float bb[4], cc[4];
    V4<T>::ld(Bv + 4 * n4, bb);
    V4<T>::ld(Cv + 4 * n4, cc);
    h4.x = fmaf(ex2_approx(dl2 * a4.x), h4.x, du * bb[0]);
    h4.y = fmaf(ex2_approx(dl2 * a4.y), h4.y, du * bb[1]);
    h4.z = fmaf(ex2_approx(dl2 * a4.z), h4.z, du * bb[2]);
    h4.w = fmaf(ex2_approx(dl2 * a4.w), h4.w, du * bb[3]);
    reinterpret_cast<float4*>(h)[n4] = h4;
    y = fmaf(h4.x, cc[0], fmaf(h4.y, cc[1], fmaf(h4.z, cc[2], fmaf(h4.w, cc[3], y))));
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) y += __shfl_xor_sync(mask, y, o);
  if (hl == 0) {
    y = fmaf((p.flags & MAMBA_FLAG_HAS_D) ? p.Dv[d] : 0.f, xc, y);
    if (p.flags & MAMBA_FLAG_HAS_Z) y *= silu_fast(IO<T>::ld(static_cast<const T*>(p.z) + (int64_t)b * p.z_bs + d));
    IO<T>::st(static_cast<T*>(p.y) + (int64_t)b * p.y_bs + d, y);
  }
}

static int step_common(const MambaStepArgs* a, StepParams& p, const char* what) {
  if (!a || a->struct_size != (int32_t)sizeof(MambaStepArgs))
    return set_error(MAMBA_EINVAL, "%s: bad args pointer or struct_size", what);
  if (a->batch <= 0 || a->dim <= 0) return set_error(MAMBA_EINVAL, "%s: batch/dim must be positive", what);
  if (a->dtype != MAMBA_F32 && a->dtype != MAMBA_BF16) return set_error(MAMBA_EDTYPE, "%s: dtype %d", what, a->dtype);
  p.B = a->batch, p.D = a->dim, p.N = a->dstate, p.K = a->width, p.R = a->dt_rank, p.flags = a->flags;
  p.x = a->x, p.dt_in = a->dt_in, p.Bv = a->Bv, p.Cv = a->Cv, p.z = a->z;
  p.conv_state = a->conv_state, p.xc = a->xc, p.y = a->y;
  p.x_bs = a->x_bs, p.xc_bs = a->xc_bs, p.dt_in_bs = a->dt_in_bs, p.Bv_bs = a->Bv_bs, p.Cv_bs = a->Cv_bs;
  p.z_bs = a->z_bs, p.y_bs = a->y_bs;
  p.cw = a->conv_weight, p.cb = a->conv_bias, p.dtw = a->dt_weight, p.dtb = a->dt_bias, p.A = a->A, p.Dv = a->D;
  p.h = a->ssm_state;
  return MAMBA_OK;
}

}  // namespace mb

extern "C" int mamba_conv_step(const MambaStepArgs* a, void* stream) {
  using namespace mb;
  StepParams p{};
  int rc = step_common(a, p, "conv_step");
  if (rc) return rc;
  if (!a->x || !a->conv_state || !a->conv_weight || !a->xc) return set_error(MAMBA_EINVAL, "conv_step: null pointer");
  if (a->width < 1 || a->width > 16) return set_error(MAMBA_EINVAL, "conv_step: width %d out of range", a->width);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n = (int64_t)p.B * p.D;
  const int blocks = (int)((n + 255) / 256);
  if (a->dtype == MAMBA_F32)
    conv_step_kernel<float><<<blocks, 256, 0, st>>>(p);
  else
    conv_step_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p);
  count_launch();
  return check_launch("conv_step");
}

extern "C" int mamba_ssm_step(const MambaStepArgs* a, void* stream) {
  using namespace mb;
  StepParams p{};
  int rc = step_common(a, p, "ssm_step");
  if (rc) return rc;
  if (!a->xc || !a->dt_in || !a->Bv || !a->Cv || !a->dt_weight || !a->A || !a->ssm_state || !a->y)
    return set_error(MAMBA_EINVAL, "ssm_step: null pointer");
  if (a->dstate <= 0 || a->dt_rank <= 0) return set_error(MAMBA_EINVAL, "ssm_step: dstate/dt_rank must be positive");
  if ((a->flags & MAMBA_FLAG_HAS_Z) && !a->z) return set_error(MAMBA_EINVAL, "ssm_step: HAS_Z but z == NULL");
  if ((a->flags & MAMBA_FLAG_HAS_D) && !a->D) return set_error(MAMBA_EINVAL, "ssm_step: HAS_D but D == NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t elt = a->dtype == MAMBA_F32 ? 4 : 2;
  auto al = [&](const void* q, int64_t stride, size_t e) {
    return reinterpret_cast<uintptr_t>(q) % (4 * e) == 0 && (stride * e) % (4 * e) == 0;
  };
  const bool fast = p.N % 4 == 0 && p.R % 4 == 0 && aligned16(a->dt_weight) && aligned16(a->A) && aligned16(a->ssm_state) &&
                    al(a->dt_in, a->dt_in_bs, elt) && al(a->Bv, a->Bv_bs, elt) && al(a->Cv, a->Cv_bs, elt);
  if (fast) {
    const int64_t items = (int64_t)p.B * p.D;
    const int blocks = (int)((items + 15) / 16);
    if (a->dtype == MAMBA_F32)
      launch_chain(ssm_step_fast_kernel<float>, dim3(blocks), dim3(256), 0, st, p);
    else
      launch_chain(ssm_step_fast_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, p);
  } else {
    const int warps = 8;
    const int blocks = ceil_div(p.D, warps);
    if (a->dtype == MAMBA_F32)
      ssm_step_kernel<float><<<blocks, warps * 32, 0, st>>>(p);
    else
      ssm_step_kernel<__nv_bfloat16><<<blocks, warps * 32, 0, st>>>(p);
  }
  count_launch();
  return check_launch("ssm_step");
}
