// conv1d.cu — causal depthwise conv1d + SiLU, forward and backward, channels-last, sm_100a.
//
// Replaces (reference simple_mamba.pyc @L233-237, module built at @L193-199)
//     x = rearrange(x, 'b l d -> b d l'); x = conv1d(x)[:, :, :l]; x = rearrange(x, 'b d l -> b l d'); F.silu(x)
// The reference pays two transposes, a padded grouped conv, a slice and an activation pass.  Here x
// stays [B, L, D]: a thread owns VEC consecutive channels and slides a K-wide register window along
// a segment of the sequence, so every global access is a coalesced vector and each input element is
// read once per segment (plus a K-1 halo).
#include "common.cuh"

namespace mb {

constexpr int kConvTS = 8;       // timesteps per thread segment: short segments whose rows are ALL loaded up front
                                 // (clamped addresses, no control dependence), so that every thread has its
                                 // whole working set in flight and there are enough threads per SM to cover HBM
                                 // latency; the K-1 halo rows are L1/L2 hits (neighbouring segments)
constexpr int kConvWarps = 8;    // segments per block

struct ConvParams {
  int B, L, D, K, nseg;
  const void *x, *dout;
  void *out, *dx, *final_state;
  int64_t x_bs, x_ls, out_bs, out_ls, dout_bs, dout_ls, dx_bs, dx_ls;
  const float *w, *bias;
  float* ws;  // [B * nsegblk][K + 1][D] partials of dweight / dbias
  float *dw, *dbias;
  int nsegblk;
};

template <typename T, int VEC>
struct Vec;
template <>
struct Vec<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a), v[1] = __high2float(a), v[2] = __low2float(b), v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a), t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};
template <typename T>
struct Vec<T, 1> {
  static __device__ __forceinline__ void ld(const T* p, float (&v)[1]) { v[0] = IO<T>::ld(p); }
  static __device__ __forceinline__ void st(T* p, const float (&v)[1]) { IO<T>::st(p, v[0]); }
};

// The thread's VEC x K filter taps.  [D][K] row-major: with K == 4 a channel's taps are one 16-byte vector (the scalar
// form cost 16 load instructions and 512 L1 sectors per warp — more L1 traffic than the activations, ncu r02)
template <int VEC, int K>
__device__ __forceinline__ void load_taps(const float* __restrict__ w, int dv, int D, float (&out)[K][VEC]) {
  if constexpr (K == 4) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(w + (int64_t)min(dv + v, D - 1) * 4));
      out[0][v] = t.x, out[1][v] = t.y, out[2][v] = t.z, out[3][v] = t.w;
    }
  } else {
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
      for (int k = 0; k < K; ++k) out[k][v] = w[(int64_t)min(dv + v, D - 1) * K + k];
  }
}

// One row of a thread's tile: loaded unconditionally from a clamped timestep (so that all of a thread's loads
// are independent of each other and of any branch) and zeroed afterwards if the timestep is outside [0, L).
template <typename T, int VEC>
__device__ __forceinline__ void ld_row_clamped(const T* base, int64_t ls, int t, int L, float (&v)[VEC]) {
  const int tc = min(max(t, 0), L - 1);
  Vec<T, VEC>::ld(base + (int64_t)tc * ls, v);
  if (t < 0 || t >= L) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = 0.f;
  }
}

// grid: (ceil(D/VEC/32), ceil(nseg/kConvWarps), B); block: (32, kConvWarps)
template <typename T, int VEC, int K>
__global__ void __launch_bounds__(32 * kConvWarps) conv_fwd_kernel(const ConvParams p) {
  const int dv = (blockIdx.x * 32 + threadIdx.x) * VEC;
  const int seg = blockIdx.y * kConvWarps + threadIdx.y;
  const int b = blockIdx.z;
  if (dv >= p.D || seg >= p.nseg) return;
  const int t0 = seg * kConvTS, t1 = min(t0 + kConvTS, p.L);
  const T* x = static_cast<const T*>(p.x) + (int64_t)b * p.x_bs + dv;
  T* out = static_cast<T*>(p.out) + (int64_t)b * p.out_bs + dv;

  float w[K][VEC], bias[VEC];
  load_taps<VEC, K>(p.w, dv, p.D, w);
#pragma unroll
  for (int v = 0; v < VEC; ++v) bias[v] = p.bias ? p.bias[dv + v] : 0.f;
  float xr[kConvTS + K - 1][VEC];  // xr[i] = x[t0 - (K-1) + i]
#pragma unroll
  for (int i = 0; i < kConvTS + K - 1; ++i) ld_row_clamped<T, VEC>(x, p.x_ls, t0 - (K - 1) + i, p.L, xr[i]);
#pragma unroll
  for (int i = 0; i < kConvTS; ++i) {
    const int t = t0 + i;
    if (t < t1) {
      float o[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float acc = bias[v];
#pragma unroll
        for (int k = 0; k < K; ++k) acc = fmaf(w[k][v], xr[i + k][v], acc);
        o[v] = silu_fast(acc);
      }
      Vec<T, VEC>::st(out + (int64_t)t * p.out_ls, o);
    }
  }
  if (p.final_state != nullptr && t1 == p.L) {
    // conv_state[b, d, k] = x[b, L-K+k, d] (zero where L-K+k < 0); win holds x[L-K .. L-1] unless the
    // segment is shorter than K, so re-read instead of relying on the window.
    T* fs = static_cast<T*>(p.final_state) + ((int64_t)b * p.D + dv) * K;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int t = p.L - K + k;
      float xv[VEC];
      if (t >= 0) {
        Vec<T, VEC>::ld(x + (int64_t)t * p.x_ls, xv);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) xv[v] = 0.f;
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) IO<T>::st(fs + v * K + k, xv[v]);
    }
  }
}

// Backward.  dpre[t] = dout[t] * silu'(pre[t]);  dx[t] = sum_k w[k] * dpre[t + (K-1) - k];
// dw[k] = sum_{b,t} dpre[t] * x[t-(K-1)+k];  dbias = sum dpre.
template <typename T, int VEC, int K>
__global__ void __launch_bounds__(32 * kConvWarps, 2) conv_bwd_kernel(const ConvParams p) {   // <= 128 registers: two blocks per SM
  __shared__ float red[kConvWarps][K + 1][32 * VEC];
  const int dv = (blockIdx.x * 32 + threadIdx.x) * VEC;
  const int seg = blockIdx.y * kConvWarps + threadIdx.y;
  const int b = blockIdx.z;
  const bool active = dv < p.D && seg < p.nseg;
  float dwacc[K][VEC], dbacc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    dbacc[v] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) dwacc[k][v] = 0.f;
  }
  if (active) {
    const int t0 = seg * kConvTS, t1 = min(t0 + kConvTS, p.L);
    const T* x = static_cast<const T*>(p.x) + (int64_t)b * p.x_bs + dv;
    const T* dout = static_cast<const T*>(p.dout) + (int64_t)b * p.dout_bs + dv;
    T* dx = static_cast<T*>(p.dx) + (int64_t)b * p.dx_bs + dv;
    float w[K][VEC], bias[VEC];
    load_taps<VEC, K>(p.w, dv, p.D, w);
#pragma unroll
    for (int v = 0; v < VEC; ++v) bias[v] = p.bias ? p.bias[dv + v] : 0.f;
    // rows needed: x[t0-(K-1) .. t1+K-2] and dout[t0 .. t1+K-2] (dx[t] needs dpre[t .. t+K-1])
    constexpr int NX = kConvTS + 2 * (K - 1), ND = kConvTS + K - 1;
    float xr[NX][VEC], dp[ND][VEC];
#pragma unroll
    for (int i = 0; i < NX; ++i) ld_row_clamped<T, VEC>(x, p.x_ls, t0 - (K - 1) + i, p.L, xr[i]);
#pragma unroll
    for (int i = 0; i < ND; ++i) ld_row_clamped<T, VEC>(dout, p.dout_ls, t0 + i, p.L, dp[i]);
    // dpre[tt] = dout[tt] * silu'(pre[tt]) in place (rows past L are zero: dout was zeroed)
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      const bool own = t0 + i < t1;  // own timestep: contributes to the parameter gradients exactly once
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float pre = bias[v];
#pragma unroll
        for (int k = 0; k < K; ++k) pre = fmaf(w[k][v], xr[i + k][v], pre);
        const float sg = rcp_approx(1.f + ex2_approx(-pre * kLog2e));
        const float d = dp[i][v] * sg * fmaf(pre, 1.f - sg, 1.f);
        dp[i][v] = d;
        if (own) {
          dbacc[v] += d;
#pragma unroll
          for (int k = 0; k < K; ++k) dwacc[k][v] = fmaf(d, xr[i + k][v], dwacc[k][v]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kConvTS; ++i) {
      const int to = t0 + i;  // dx[to] = sum_k w[k] * dpre[to + (K-1) - k]
      if (to < t1) {
        float o[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < K; ++k) acc = fmaf(w[k][v], dp[i + (K - 1) - k][v], acc);
          o[v] = acc;
        }
        Vec<T, VEC>::st(dx + (int64_t)to * p.dx_ls, o);
      }
    }
  }
  // block-level reduction over the kConvWarps segments, then one partial row per block
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
#pragma unroll
    for (int k = 0; k < K; ++k) red[threadIdx.y][k][threadIdx.x * VEC + v] = dwacc[k][v];
    red[threadIdx.y][K][threadIdx.x * VEC + v] = dbacc[v];
  }
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;
  float* wsrow = p.ws + ((int64_t)b * p.nsegblk + blockIdx.y) * (K + 1) * p.D;
  for (int i = tid; i < (K + 1) * 32 * VEC; i += 32 * kConvWarps) {
    const int k = i / (32 * VEC), c = i % (32 * VEC);
    const int d = blockIdx.x * 32 * VEC + c;
    if (d < p.D) {
      float s = 0.f;
#pragma unroll
      for (int ww = 0; ww < kConvWarps; ++ww) s += red[ww][k][c];
      wsrow[(int64_t)k * p.D + d] = s;
    }
  }
}

// grid: ceil((K+1)*D / 32) blocks of (32, 8) threads
template <int K>
__global__ void conv_bwd_finalize_kernel(const ConvParams p) {
  const int nrow = p.B * p.nsegblk;
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;  // column of the [nrow][(K+1)*D] partial buffer
  const bool ok = i < (int64_t)(K + 1) * p.D;
  const float s = colsum_32x8(p.ws, nrow, (int64_t)(K + 1) * p.D, (int)i, ok);
  if (ok && threadIdx.y == 0) {
    const int k = (int)(i / p.D), d = (int)(i % p.D);
    if (k < K)
      p.dw[(int64_t)d * K + k] = s;
    else if (p.dbias)
      p.dbias[d] = s;
  }
}

template <typename T, int VEC, int K>
static int launch_conv(const ConvParams& p, bool bwd, cudaStream_t st) {
  dim3 block(32, kConvWarps);
  dim3 grid(ceil_div(ceil_div(p.D, VEC), 32), p.nsegblk, p.B);
  if (!bwd) {
    conv_fwd_kernel<T, VEC, K><<<grid, block, 0, st>>>(p);
    count_launch();
    return check_launch("conv1d_silu_fwd");
  }
  conv_bwd_kernel<T, VEC, K><<<grid, block, 0, st>>>(p);
  count_launch();
  int rc = check_launch("conv1d_silu_bwd");
  if (rc) return rc;
  conv_bwd_finalize_kernel<K><<<ceil_div((K + 1) * p.D, 32), dim3(32, 8), 0, st>>>(p);
  count_launch();
  return check_launch("conv1d_silu_bwd_finalize");
}

template <typename T, int VEC>
static int conv_dispatch_k(const ConvParams& p, bool bwd, cudaStream_t st) {
  switch (p.K) {
    case 2: return launch_conv<T, VEC, 2>(p, bwd, st);
    case 3: return launch_conv<T, VEC, 3>(p, bwd, st);
    case 4: return launch_conv<T, VEC, 4>(p, bwd, st);
  }
  return set_error(MAMBA_EINVAL, "conv1d: width must be 2, 3 or 4 (got %d)", p.K);
}

static int conv_common(const MambaConvArgs* a, bool bwd, void* stream) {
  if (!a || a->struct_size != (int32_t)sizeof(MambaConvArgs))
    return set_error(MAMBA_EINVAL, "conv1d: bad args pointer or struct_size");
  if (a->batch <= 0 || a->seqlen <= 0 || a->dim <= 0)
    return set_error(MAMBA_EINVAL, "conv1d: batch/seqlen/dim must be positive (got %d/%d/%d)", a->batch, a->seqlen,
                     a->dim);
  if (a->batch > 65535) return set_error(MAMBA_ESIZE, "conv1d: batch %d above 65535", a->batch);
  if (!a->x || !a->weight) return set_error(MAMBA_EINVAL, "conv1d: null x/weight");
  if (a->width == 4 && !mb::aligned16(a->weight))
    return set_error(MAMBA_EALIGN, "conv1d: weight [D, 4] must be 16-byte aligned (one vector load per channel)");
  if (!bwd && !a->out) return set_error(MAMBA_EINVAL, "conv1d_fwd: null out");
  if (bwd && (!a->dout || !a->dx || !a->dweight)) return set_error(MAMBA_EINVAL, "conv1d_bwd: null dout/dx/dweight");
  if (a->dtype != MAMBA_F32 && a->dtype != MAMBA_BF16) return set_error(MAMBA_EDTYPE, "conv1d: dtype %d", a->dtype);
  ConvParams p{};
  p.B = a->batch, p.L = a->seqlen, p.D = a->dim, p.K = a->width;
  p.nseg = ceil_div(p.L, kConvTS);
  p.nsegblk = ceil_div(p.nseg, kConvWarps);
  p.x = a->x, p.dout = a->dout, p.out = a->out, p.dx = a->dx, p.final_state = a->final_state;
  p.x_bs = a->x_bs, p.x_ls = a->x_ls, p.out_bs = a->out_bs, p.out_ls = a->out_ls;
  p.dout_bs = a->dout_bs, p.dout_ls = a->dout_ls, p.dx_bs = a->dx_bs, p.dx_ls = a->dx_ls;
  p.w = a->weight, p.bias = a->bias, p.dw = a->dweight, p.dbias = a->dbias;
  const size_t elt = a->dtype == MAMBA_F32 ? 4 : 2;
  const size_t va = 4 * elt;  // bytes of one 4-channel vector
  auto ok = [&](const void* ptr, int64_t bs, int64_t ls) {
    return (reinterpret_cast<uintptr_t>(ptr) % va) == 0 && (bs * elt) % va == 0 && (ls * elt) % va == 0;
  };
  bool vec = (p.D % 4 == 0) && ok(a->x, a->x_bs, a->x_ls);
  if (!bwd) vec = vec && ok(a->out, a->out_bs, a->out_ls);
  if (bwd) vec = vec && ok(a->dout, a->dout_bs, a->dout_ls) && ok(a->dx, a->dx_bs, a->dx_ls);
  if (bwd) {
    const size_t need = mamba_conv1d_bwd_workspace_bytes(p.B, p.L, p.D, p.K);
    if (!a->workspace || a->workspace_bytes < need)
      return set_error(MAMBA_ESIZE, "conv1d_bwd: workspace %zu B < required %zu B", a->workspace_bytes, need);
    p.ws = static_cast<float*>(a->workspace);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->dtype == MAMBA_F32)
    return vec ? conv_dispatch_k<float, 4>(p, bwd, st) : conv_dispatch_k<float, 1>(p, bwd, st);
  return vec ? conv_dispatch_k<__nv_bfloat16, 4>(p, bwd, st) : conv_dispatch_k<__nv_bfloat16, 1>(p, bwd, st);
}

}  // namespace mb

extern "C" size_t mamba_conv1d_bwd_workspace_bytes(int batch, int seqlen, int dim, int width) {
  if (batch <= 0 || seqlen <= 0 || dim <= 0 || width <= 0) return 0;
  const int nseg = mb::ceil_div(seqlen, mb::kConvTS);
  const int nsegblk = mb::ceil_div(nseg, mb::kConvWarps);
  return (size_t)4 * batch * nsegblk * (width + 1) * dim;
}
extern "C" int mamba_conv1d_silu_fwd(const MambaConvArgs* a, void* stream) { return mb::conv_common(a, false, stream); }
extern "C" int mamba_conv1d_silu_bwd(const MambaConvArgs* a, void* stream) { return mb::conv_common(a, true, stream); }
