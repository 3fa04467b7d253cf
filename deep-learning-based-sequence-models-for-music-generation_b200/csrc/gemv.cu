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
};

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
          if (i < total4 && b < p.B) V4<TX>::ld(static_cast<const TX*>(p.x) + (int64_t)b * p.x_bs + 4 * k4, v[u]);
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
        for (int k = tid; k < p.K; k += nthr) dst[k] = b < p.B ? IO<TX>::ld(xr + k) : 0.f;
      }
    }
  }
  __syncthreads();
  const int gl = tid & (GW - 1);                           // lane within the group
  const int group = (blockIdx.x * nthr + tid) / GW;        // global group id
  const int ngroups = (gridDim.x * nthr) / GW;
  const unsigned gmask = GW == 32 ? 0xffffffffu : (0xffu << (tid & 24));
  for (int n0 = group * ROWS; n0 < p.N; n0 += ngroups * ROWS) {
    float2 acc[ROWS][BT];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[r][b] = make_float2(0.f, 0.f);
    const TW* wrow[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) wrow[r] = static_cast<const TW*>(p.w) + (int64_t)min(n0 + r, p.N - 1) * p.K;
    // weight loads of 16 (one row per warp) or 4 (four rows) steps in flight per lane
    constexpr int kUnrollK = ROWS == 1 ? 16 : 4;
#pragma unroll kUnrollK
    for (int k4 = gl; k4 < K4; k4 += GW) {
      float2 w01[ROWS], w23[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        float w[4];
        ld_w4<TW>(wrow[r] + 4 * k4, w);
        w01[r] = make_float2(w[0], w[1]), w23[r] = make_float2(w[2], w[3]);
      }
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + b * p.K + 4 * k4);
        const float2 x01 = make_float2(xv.x, xv.y), x23 = make_float2(xv.z, xv.w);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          acc[r][b] = __ffma2_rn(w01[r], x01, acc[r][b]);
          acc[r][b] = __ffma2_rn(w23[r], x23, acc[r][b]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      float mine = 0.f, mine_hi = 0.f;  // group lane b (and, for 8-lane groups, b - 8) keeps batch row b
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        float v = acc[r][b].x + acc[r][b].y;
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
        if (gl < p.B && gl < (GW == 32 ? BT : 8)) IO<TX>::st(static_cast<TX*>(p.y) + (int64_t)gl * p.y_bs + n, mine + bias);
        if (GW == 8 && BT > 8 && gl + 8 < p.B)
          IO<TX>::st(static_cast<TX*>(p.y) + (int64_t)(gl + 8) * p.y_bs + n, mine_hi + bias);
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
  kern<<<blocks, kGemvThreads, smem, st>>>(p);
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
  if (p.B <= 12) return gemv_launch<TX, TW, 12>(p, st);
  if (p.B <= 16) return gemv_launch<TX, TW, 16>(p, st);
  return set_error(MAMBA_ESIZE, "linear_step: batch %d above 16 (use a GEMM)", p.B);
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int xd = a->dtype, wd = a->w_dtype;
  if (xd == MAMBA_F32 && wd == MAMBA_F32) return gemv_dispatch_b<float, float>(p, st);
  if (xd == MAMBA_BF16 && wd == MAMBA_BF16) return gemv_dispatch_b<__nv_bfloat16, __nv_bfloat16>(p, st);
  if (xd == MAMBA_F32 && wd == MAMBA_BF16) return gemv_dispatch_b<float, __nv_bfloat16>(p, st);
  if (xd == MAMBA_BF16 && wd == MAMBA_F32) return gemv_dispatch_b<__nv_bfloat16, float>(p, st);
  return set_error(MAMBA_EDTYPE, "linear_step: dtype %d / w_dtype %d", xd, wd);
}
