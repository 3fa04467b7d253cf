// gemv.cu — y[b, :] = W x[b, :] (+ bias) for a handful of rows (one new token per sequence), sm_100a.
//
// Decode-time replacement of the four nn.Linear calls of MambaBlock / Mamba for ONE position per sequence
// (reference simple_mamba.pyc: in_proj @L230, x_proj @L273, out_proj @L243, lm_head @L94).  With <= 16 rows
// the product is a weight stream, not a GEMM: cuBLAS falls onto SIMT sgemm kernels that reach a few percent of
// HBM bandwidth.  Here every warp streams whole weight rows with 128-bit loads (each weight byte is read exactly
// once), the activation rows sit in shared memory as fp32, products accumulate in packed fp32x2, and one
// warp-level reduction per output row finishes the dot products of all batch rows at once.
#include "common.cuh"

namespace mb {

constexpr int kGemvThreads = 256;

struct GemvParams {
  int B, K, N;
  const void *x, *w, *bias;
  void* y;
  int64_t x_bs, y_bs;
  // fused prologue: x := rmsnorm(x + resid_in) * norm_w (fp32 residual stream; x may be NULL = zeros)
  const float *norm_w, *resid_in;
  float* resid_out;
  int64_t resid_in_bs, resid_out_bs;
  float eps;
  // fused epilogue: output columns [0, conv_dim) feed the depthwise conv step of their channel
  int conv_dim, conv_k;
  void *conv_state, *conv_out;
  int64_t conv_out_bs;
  const float *conv_w, *conv_b;
};

// one new input of channel n for sequence b: shift the K-deep register, y = silu(bias + sum_k w[k] * state[k])
// (the arithmetic of conv_step_kernel in step.cu)
template <typename TX>
__device__ __forceinline__ void conv_epilogue(const GemvParams& p, int b, int n, float v) {
  TX* st = static_cast<TX*>(p.conv_state) + ((int64_t)b * p.conv_dim + n) * p.conv_k;
  float acc = p.conv_b ? p.conv_b[n] : 0.f;
  for (int k = 0; k < p.conv_k; ++k) {
    const float sv = (k + 1 < p.conv_k) ? IO<TX>::ld(st + k + 1) : IO<TX>::cvt(TX(v));
    IO<TX>::st(st + k, sv);
    acc = fmaf(p.conv_w[(int64_t)n * p.conv_k + k], sv, acc);
  }
  IO<TX>::st(static_cast<TX*>(p.conv_out) + (int64_t)b * p.conv_out_bs + n, silu_f(acc));
}

template <typename TW>
__device__ __forceinline__ void ld_w4(const TW* p, float (&w)[4]);
template <>
__device__ __forceinline__ void ld_w4<float>(const float* p, float (&w)[4]) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
}
template <>
__device__ __forceinline__ void ld_w4<__nv_bfloat16>(const __nv_bfloat16* p, float (&w)[4]) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  w[0] = __uint_as_float(v.x << 16), w[1] = __uint_as_float(v.x & 0xffff0000u);
  w[2] = __uint_as_float(v.y << 16), w[3] = __uint_as_float(v.y & 0xffff0000u);
}

// Work split: a group of 8 lanes owns ROWS consecutive output rows and walks their K axis with 16-byte loads
// (8 lanes x 16 B = one 128-byte line per row and step); a warp is 4 such groups.  Every activation vector
// fetched from shared memory is reused for ROWS weight rows, which keeps the shared-memory pipe (12 LDS.128 per
// step at batch 12) below the weight stream.  Reduction over the 8 lanes: 3 butterfly steps per (row, batch).
template <typename TX, typename TW, int BT, int ROWS, int GW>
__global__ void __launch_bounds__(kGemvThreads) gemv_kernel(const GemvParams p) {
  static_assert(GW == 8 || GW == 32, "group width");
  extern __shared__ __align__(16) float xs[];  // [BT][K]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int K4 = p.K >> 2;
  {  // decode chain (common.cuh): let the next kernel start, pull this group's first weight rows into L2, then wait
     // for the kernel that produces x
    pdl_launch_dependents();
    const int gl0 = tid & (GW - 1);
    const int group0 = (blockIdx.x * nthr + tid) / GW;
    const size_t row_bytes = (size_t)p.K * sizeof(TW);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int n = group0 * ROWS + r;
      if (n < p.N) {
        const char* wr = static_cast<const char*>(p.w) + (size_t)n * row_bytes;
        for (size_t off = (size_t)gl0 * 128; off < row_bytes; off += (size_t)GW * 128) prefetch_l2(wr + off);
      }
    }
    pdl_wait();
  }
  {  // stage the activation rows as fp32 (128-bit where the rows allow it); rows >= B are zero.  All of a thread's
     // loads are issued before its first store: the copy is a handful of dependent L2 round trips otherwise, and
     // at 10-16 rows that latency, not the weight stream, was the whole kernel.
    const bool vec = (reinterpret_cast<uintptr_t>(p.x) % (4 * sizeof(TX)) == 0) && (p.x_bs % 4 == 0);
    if (vec) {
      constexpr int U = 16;
      const int total4 = BT * K4;
      for (int base = tid; base < total4; base += nthr * U) {
        float v[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * nthr;
          const int b = i / K4, k4 = i - b * K4;
          v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
          if (i < total4 && b < p.B && p.x != nullptr) V4<TX>::ld(static_cast<const TX*>(p.x) + (int64_t)b * p.x_bs + 4 * k4, v[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * nthr;
          if (i < total4) *reinterpret_cast<float4*>(xs + 4 * i) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
        }
      }
    } else {
      for (int b = 0; b < BT; ++b) {
        const TX* xr = static_cast<const TX*>(p.x) + (int64_t)b * p.x_bs;
        float* dst = xs + b * p.K;
        for (int k = tid; k < p.K; k += nthr) dst[k] = (b < p.B && p.x != nullptr) ? IO<TX>::ld(xr + k) : 0.f;
      }
    }
  }
  __syncthreads();
  if (p.norm_w != nullptr) {
    // fused `normed, resid = norm(hidden, resid)` of the residual block (simple_mamba.pyc @L179 / @L346): every block
    // normalises its own copy of the <= 16 rows (40 KB of L2 reads); block 0 writes the new residual stream
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    for (int b = warp; b < p.B; b += nw) {
      float* row = xs + b * p.K;
      const float* rin = p.resid_in ? p.resid_in + (int64_t)b * p.resid_in_bs : nullptr;
      float ss = 0.f;
      for (int k = lane; k < p.K; k += 32) {
        const float v = row[k] + (rin ? rin[k] : 0.f);
        row[k] = v;
        ss = fmaf(v, v, ss);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = rsqrtf(ss / (float)p.K + p.eps);
      for (int k = lane; k < p.K; k += 32) {
        const float v = row[k];
        if (blockIdx.x == 0 && p.resid_out) p.resid_out[(int64_t)b * p.resid_out_bs + k] = v;
        row[k] = v * rstd * p.norm_w[k];
      }
    }
    __syncthreads();
  }
  const int gl = tid & (GW - 1);                           // lane within the group
  const int group = (blockIdx.x * nthr + tid) / GW;        // global group id
  const int ngroups = (gridDim.x * nthr) / GW;
  const unsigned gmask = GW == 32 ? 0xffffffffu : (0xffu << (tid & 24));
  for (int n0 = group * ROWS; n0 < p.N; n0 += ngroups * ROWS) {
    // scalar FFMA: packed FFMA2 with three distinct register pairs runs at half the lane rate (tools/microbench3.cu),
    // and at 10+ rows per weight vector the multiply-adds, not the weight stream, set this kernel's time
    float acc[ROWS][BT];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[r][b] = 0.f;
    const TW* wrow[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) wrow[r] = static_cast<const TW*>(p.w) + (int64_t)min(n0 + r, p.N - 1) * p.K;
    // weight loads of 16 (one row per warp) or 4 (four rows) steps in flight per lane
    constexpr int kUnrollK = ROWS == 1 ? 16 : 4;
#pragma unroll kUnrollK
    for (int k4 = gl; k4 < K4; k4 += GW) {
      float w[ROWS][4];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) ld_w4<TW>(wrow[r] + 4 * k4, w[r]);
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + b * p.K + 4 * k4);
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
          acc[r][b] = fmaf(w[r][3], xv.w, fmaf(w[r][2], xv.z, fmaf(w[r][1], xv.y, fmaf(w[r][0], xv.x, acc[r][b]))));
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      float mine = 0.f, mine_hi = 0.f;  // group lane b (and, for 8-lane groups, b - 8) keeps batch row b
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        float v = acc[r][b];
#pragma unroll
        for (int o = GW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
        if (GW == 32) {
          if (b == gl) mine = v;
        } else {
          if ((b & 7) == gl) (b < 8 ? mine : mine_hi) = v;
        }
      }
      const int n = n0 + r;
      if (n < p.N) {
        const float bias = p.bias ? IO<TW>::ld(static_cast<const TW*>(p.bias) + n) : 0.f;
        if (gl < p.B && gl < (GW == 32 ? BT : 8)) {
          if (n < p.conv_dim) conv_epilogue<TX>(p, gl, n, mine + bias);
          else IO<TX>::st(static_cast<TX*>(p.y) + (int64_t)gl * p.y_bs + n, mine + bias);
        }
        if (GW == 8 && BT > 8 && gl + 8 < p.B) {
          if (n < p.conv_dim) conv_epilogue<TX>(p, gl + 8, n, mine_hi + bias);
          else IO<TX>::st(static_cast<TX*>(p.y) + (int64_t)(gl + 8) * p.y_bs + n, mine_hi + bias);
        }
      }
    }
  }
}

template <typename TX, typename TW, int BT, int ROWS, int GW>
static int gemv_launch_cfg(const GemvParams& p, cudaStream_t st) {
  const size_t smem = (size_t)BT * p.K * 4;
  auto kern = gemv_kernel<TX, TW, BT, ROWS, GW>;
  static thread_local SmemConfig cfg;
  if (smem > 48 * 1024)
    if (int rc = ensure_dynamic_smem(kern, smem, cfg, "linear_step")) return rc;
  const int rows_per_block = (kGemvThreads / GW) * ROWS;
  int blocks = ceil_div(p.N, rows_per_block);
  const int per_sm = smem > 100 * 1024 ? 1 : (smem > 50 * 1024 ? 2 : 4);
  if (blocks > kNumSMs * per_sm) blocks = kNumSMs * per_sm;
  launch_chain(kern, dim3(blocks), dim3(kGemvThreads), smem, st, p);
  count_launch();
  return check_launch("linear_step");
}

template <typename TX, typename TW, int BT>
static int gemv_launch(const GemvParams& p, cudaStream_t st) {
  if ((size_t)BT * p.K * 4 > 200 * 1024)
    return set_error(MAMBA_ESIZE, "linear_step: %d x %d activations exceed shared memory", BT, p.K);
  // rows per warp ~ N / (148 SMs x 8 warps): spread the weight rows over the whole chip, reuse each activation
  // fetch for 4 rows when there are enough of them
  const int rows_per_warp = p.N / (kNumSMs * 8);
  if (rows_per_warp >= 8) return gemv_launch_cfg<TX, TW, BT, 4, 8>(p, st);    // lm_head
  if (rows_per_warp >= 2) return gemv_launch_cfg<TX, TW, BT, 4, 32>(p, st);   // in_proj
  return gemv_launch_cfg<TX, TW, BT, 1, 32>(p, st);                           // out_proj, x_proj
}

template <typename TX, typename TW>
static int gemv_dispatch_b(const GemvParams& p, cudaStream_t st) {
  if (p.B <= 1) return gemv_launch<TX, TW, 1>(p, st);
  if (p.B <= 2) return gemv_launch<TX, TW, 2>(p, st);
  if (p.B <= 4) return gemv_launch<TX, TW, 4>(p, st);
  if (p.B <= 8) return gemv_launch<TX, TW, 8>(p, st);
  if (p.B <= 10) return gemv_launch<TX, TW, 10>(p, st);   // BASELINE config 4: 5 composer bands x 2 samples
  if (p.B <= 12) return gemv_launch<TX, TW, 12>(p, st);
  if (p.B <= 16) return gemv_launch<TX, TW, 16>(p, st);
  return set_error(MAMBA_ESIZE, "linear_step: batch %d above 16 (use a GEMM)", p.B);
}

static int gemv_dispatch_types(const GemvParams& p, int xd, int wd, cudaStream_t st) {
  if (xd == MAMBA_F32 && wd == MAMBA_F32) return gemv_dispatch_b<float, float>(p, st);
  if (xd == MAMBA_BF16 && wd == MAMBA_BF16) return gemv_dispatch_b<__nv_bfloat16, __nv_bfloat16>(p, st);
  if (xd == MAMBA_F32 && wd == MAMBA_BF16) return gemv_dispatch_b<float, __nv_bfloat16>(p, st);
  if (xd == MAMBA_BF16 && wd == MAMBA_F32) return gemv_dispatch_b<__nv_bfloat16, float>(p, st);
  return set_error(MAMBA_EDTYPE, "linear_step: dtype %d / w_dtype %d", xd, wd);
}

}  // namespace mb

extern "C" int mamba_linear_step(const MambaLinearStepArgs* a, void* stream) {
  using namespace mb;
  if (!a || a->struct_size != (int32_t)sizeof(MambaLinearStepArgs))
    return set_error(MAMBA_EINVAL, "linear_step: bad args pointer or struct_size");
  if (a->batch <= 0 || a->in_features <= 0 || a->out_features <= 0)
    return set_error(MAMBA_EINVAL, "linear_step: batch/in_features/out_features must be positive");
  if (!a->x || !a->weight || !a->y) return set_error(MAMBA_EINVAL, "linear_step: null x/weight/y");
  if (a->in_features % 4 != 0) return set_error(MAMBA_EALIGN, "linear_step: in_features %d must be a multiple of 4", a->in_features);
  const size_t welt = a->w_dtype == MAMBA_F32 ? 4 : 2;
  if (reinterpret_cast<uintptr_t>(a->weight) % (4 * welt) != 0)
    return set_error(MAMBA_EALIGN, "linear_step: weight must be %zu-byte aligned", 4 * welt);
  GemvParams p{};
  p.B = a->batch, p.K = a->in_features, p.N = a->out_features;
  p.x = a->x, p.w = a->weight, p.bias = a->bias, p.y = a->y, p.x_bs = a->x_bs, p.y_bs = a->y_bs;
  return mb::gemv_dispatch_types(p, a->dtype, a->w_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int mamba_fused_linear_step(const MambaFusedLinearStepArgs* a, void* stream) {
  using namespace mb;
  if (!a || a->struct_size != (int32_t)sizeof(MambaFusedLinearStepArgs))
    return set_error(MAMBA_EINVAL, "fused_linear_step: bad args pointer or struct_size");
  if (a->batch <= 0 || a->in_features <= 0 || a->out_features <= 0)
    return set_error(MAMBA_EINVAL, "fused_linear_step: batch/in_features/out_features must be positive");
  if (!a->weight || !a->y) return set_error(MAMBA_EINVAL, "fused_linear_step: null weight/y");
  if (!a->x && !(a->norm_weight && a->residual_in))
    return set_error(MAMBA_EINVAL, "fused_linear_step: x may be NULL only with a fused norm over residual_in");
  if (a->in_features % 4 != 0)
    return set_error(MAMBA_EALIGN, "fused_linear_step: in_features %d must be a multiple of 4", a->in_features);
  const size_t welt = a->w_dtype == MAMBA_F32 ? 4 : 2;
  if (reinterpret_cast<uintptr_t>(a->weight) % (4 * welt) != 0)
    return set_error(MAMBA_EALIGN, "fused_linear_step: weight must be %zu-byte aligned", 4 * welt);
  if (a->conv_dim < 0 || a->conv_dim > a->out_features || (a->conv_dim > 0 && (!a->conv_state || !a->conv_out || !a->conv_weight || a->conv_width <= 0)))
    return set_error(MAMBA_EINVAL, "fused_linear_step: conv epilogue needs conv_state/conv_out/conv_weight and 0 < conv_dim <= out_features");
  if (a->residual_out && a->residual_out == a->residual_in)
    return set_error(MAMBA_EINVAL, "fused_linear_step: residual_out must not alias residual_in (other blocks still read it)");
  GemvParams p{};
  p.B = a->batch, p.K = a->in_features, p.N = a->out_features;
  p.x = a->x, p.w = a->weight, p.bias = a->bias, p.y = a->y, p.x_bs = a->x_bs, p.y_bs = a->y_bs;
  p.norm_w = a->norm_weight, p.resid_in = a->residual_in, p.resid_out = a->residual_out;
  p.resid_in_bs = a->residual_in_bs, p.resid_out_bs = a->residual_out_bs, p.eps = a->eps;
  p.conv_dim = a->conv_dim, p.conv_k = a->conv_width, p.conv_state = a->conv_state, p.conv_out = a->conv_out;
  p.conv_out_bs = a->conv_out_bs, p.conv_w = a->conv_weight, p.conv_b = a->conv_bias;
  return gemv_dispatch_types(p, a->dtype, a->w_dtype, static_cast<cudaStream_t>(stream));
}
