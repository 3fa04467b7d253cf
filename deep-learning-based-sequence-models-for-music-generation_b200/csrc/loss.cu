// loss.cu — the reference's grammar-masked loss as four streaming passes over the logits, sm_100a.
//
// Replaces train.py:133-138 (filtered_logit) + :161-165 (reshape, CrossEntropyLoss) of the reference:
//     weights   = distributions[bucketize(src)]                       # [B, T, V], row picked by the PREVIOUS token
//     log_probs = F.log_softmax(output, dim=1)                        # over the SEQUENCE axis
//     f         = -log_probs * weights
//     loss      = CrossEntropyLoss()(f.reshape(-1, V), trg.view(-1))  # softmax over the vocab axis, mean over B*T
// In torch this is ~10 full passes over a [B, T, V] fp32 tensor (293 MB at the configured sizes) plus a strided
// softmax kernel; here it is
//   forward : (1) column pass  — per (b, v): logsumexp over t of the logits          -> col_lse[B, V]
//             (2) row pass     — per (b, t): f on the fly, logsumexp over v, f[trg]   -> row_lse[B, T], loss
//   backward: (3) column pass  — per (b, v): sum over t of d loss / d log_probs       -> colsum[B, V]
//             (4) element pass — d logits = dlp - softmax_t(logits) * colsum
// Every pass reads the logits in their storage dtype (bf16 or fp32) exactly once with V-contiguous coalesced
// accesses; col_lse / colsum / the 5-row weight table stay L2-resident.  All reductions are fixed-order
// (deterministic).  Column passes split T into segments for parallelism and are combined by a small second kernel.
#include "common.cuh"

namespace mb {

constexpr int kLossSeg = 64;       // rows per column-pass segment
constexpr int kLossColThreads = 128;
constexpr int kLossRowThreads = 256;

struct LossParams {
  int B, T, V, S;  // S = number of T-segments
  const void* logits;
  int64_t l_bs, l_ts;
  const int64_t *src, *trg;
  const float* table;  // [5][V]
  int bnd[4];
  float *col_part_m, *col_part_s;  // [B][S][V]
  float* col_lse;                  // [B][V]
  float *row_lse, *row_loss;       // [B][T]
  float* loss;                     // [1]
  const float* grad_out;           // [1] device scalar (or NULL = 1)
  float* colsum_part;              // [B][S][V]
  float* colsum;                   // [B][V]
  void* dlogits;
  int64_t d_bs, d_ts;
};

__device__ __forceinline__ int bucket_of(int64_t tok, const int (&bnd)[4]) {
  // torch.bucketize(tok, boundaries, right=False): number of boundaries strictly below tok
  return (tok > bnd[0]) + (tok > bnd[1]) + (tok > bnd[2]) + (tok > bnd[3]);
}
__device__ __forceinline__ float exp_fast(float x) { return ex2_approx(x * kLog2e); }
__device__ __forceinline__ float log_fast(float x) { return lg2_approx(x) * kLn2; }

// online logsumexp update with one exponential: (m, s) <- (m, s) (+) x
__device__ __forceinline__ void lse_push(float& m, float& s, float x) {
  const float d = x - m;
  const float e = exp_fast(-fabsf(d));
  s = d > 0.f ? fmaf(s, e, 1.f) : s + e;
  m = fmaxf(m, x);
}
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  s = s * exp_fast(m - mm) + s2 * exp_fast(m2 - mm);
  m = mm;
}

// (1) grid (ceil(V / threads), S, B): one column per thread, kLossSeg rows
template <typename T>
__global__ void __launch_bounds__(kLossColThreads) loss_col_lse_kernel(const LossParams p) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int seg = blockIdx.y, b = blockIdx.z;
  if (v >= p.V) return;
  const int t0 = seg * kLossSeg, t1 = min(t0 + kLossSeg, p.T);
  const T* x = static_cast<const T*>(p.logits) + (int64_t)b * p.l_bs + v;
  float m = -INFINITY, s = 0.f;
#pragma unroll 8
  for (int t = t0; t < t1; ++t) lse_push(m, s, IO<T>::ld(x + (int64_t)t * p.l_ts));
  const int64_t o = ((int64_t)b * p.S + seg) * p.V + v;
  p.col_part_m[o] = m;
  p.col_part_s[o] = s;
}
__global__ void loss_col_lse_combine_kernel(const LossParams p) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (v >= p.V) return;
  float m = -INFINITY, s = 0.f;
  for (int seg = 0; seg < p.S; ++seg) {
    const int64_t o = ((int64_t)b * p.S + seg) * p.V + v;
    lse_merge(m, s, p.col_part_m[o], p.col_part_s[o]);
  }
  p.col_lse[(int64_t)b * p.V + v] = m + log_fast(s);
}

// (2) one CTA per (b, t) row
template <typename T>
__global__ void __launch_bounds__(kLossRowThreads) loss_row_kernel(const LossParams p) {
  __shared__ float sm_m[kLossRowThreads / 32], sm_s[kLossRowThreads / 32];
  const int row = blockIdx.x;  // b * T + t
  const int b = row / p.T, t = row - b * p.T;
  const T* x = static_cast<const T*>(p.logits) + (int64_t)b * p.l_bs + (int64_t)t * p.l_ts;
  const float* cl = p.col_lse + (int64_t)b * p.V;
  const float* w = p.table + (int64_t)bucket_of(p.src[row], p.bnd) * p.V;
  float m = -INFINITY, s = 0.f;
  // 8 elements per thread and trip, all loads issued before the first use (the loop is latency-bound otherwise)
  constexpr int U = 8;
  for (int v0 = threadIdx.x; v0 < p.V; v0 += U * kLossRowThreads) {
    float xv[U], cv[U], wv[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int v = v0 + k * kLossRowThreads;
      const bool ok = v < p.V;
      xv[k] = ok ? IO<T>::ld(x + v) : 0.f;
      cv[k] = ok ? cl[v] : 0.f;
      wv[k] = ok ? w[v] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < U; ++k)
      if (v0 + k * kLossRowThreads < p.V) lse_push(m, s, -(xv[k] - cv[k]) * wv[k]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_merge(m, s, m2, s2);
  }
  if ((threadIdx.x & 31) == 0) sm_m[threadIdx.x >> 5] = m, sm_s[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    m = sm_m[0], s = sm_s[0];
    for (int k = 1; k < kLossRowThreads / 32; ++k) lse_merge(m, s, sm_m[k], sm_s[k]);
    const float lse = m + log_fast(s);
    const int64_t tg = p.trg[row];
    const float ft = -(IO<T>::ld(x + tg) - cl[tg]) * w[tg];
    p.row_lse[row] = lse;
    p.row_loss[row] = lse - ft;
  }
}
// mean of row_loss in a fixed order (one CTA)
__global__ void loss_mean_kernel(const LossParams p) {
  __shared__ float sm[256];
  const int n = p.B * p.T;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += p.row_loss[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) p.loss[0] = sm[0] / (float)n;
}

// d loss / d log_probs[b, t, v] = -w * (softmax_v(f) - onehot(trg)) * grad_out / (B*T)
template <typename T>
__device__ __forceinline__ float dlp_of(const LossParams& p, float x, float cl, float w, float row_lse, bool is_trg,
                                        float scale) {
  const float f = -(x - cl) * w;
  const float sm = exp_fast(f - row_lse) - (is_trg ? 1.f : 0.f);
  return -w * sm * scale;
}

// (3) same grid as (1).  The per-row scalars (weight-table row, row logsumexp, target) are the same for every
//     column of the block: staged once in shared memory instead of being re-derived per element.
template <typename T>
__global__ void __launch_bounds__(kLossColThreads) loss_colsum_kernel(const LossParams p) {
  __shared__ int s_woff[kLossSeg];   // bucket * V
  __shared__ float s_rl[kLossSeg];   // row logsumexp
  __shared__ int s_tg[kLossSeg];
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int seg = blockIdx.y, b = blockIdx.z;
  const int t0 = seg * kLossSeg, t1 = min(t0 + kLossSeg, p.T);
  for (int i = threadIdx.x; i < t1 - t0; i += blockDim.x) {
    const int64_t row = (int64_t)b * p.T + t0 + i;
    s_woff[i] = bucket_of(p.src[row], p.bnd) * p.V;
    s_rl[i] = p.row_lse[row];
    s_tg[i] = (int)p.trg[row];
  }
  __syncthreads();
  if (v >= p.V) return;
  const T* x = static_cast<const T*>(p.logits) + (int64_t)b * p.l_bs + (int64_t)t0 * p.l_ts + v;
  const float* tab = p.table + v;
  const float cl = p.col_lse[(int64_t)b * p.V + v];
  const float scale = (p.grad_out ? p.grad_out[0] : 1.f) / (float)(p.B * p.T);
  float acc = 0.f;
  constexpr int U = 8;
  const int n = t1 - t0;
  for (int i0 = 0; i0 < n; i0 += U) {
    float xv[U], wv[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int i = min(i0 + k, n - 1);
      xv[k] = IO<T>::ld(x + (int64_t)i * p.l_ts);
      wv[k] = tab[s_woff[i]];
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int i = i0 + k;
      if (i < n) acc += dlp_of<T>(p, xv[k], cl, wv[k], s_rl[i], s_tg[i] == v, scale);
    }
  }
  p.colsum_part[((int64_t)b * p.S + seg) * p.V + v] = acc;
}
__global__ void loss_colsum_combine_kernel(const LossParams p) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (v >= p.V) return;
  float acc = 0.f;
  for (int seg = 0; seg < p.S; ++seg) acc += p.colsum_part[((int64_t)b * p.S + seg) * p.V + v];
  p.colsum[(int64_t)b * p.V + v] = acc;
}

// (4) grid (ceil(V / (threads * VEC)), ceil(B*T / kDlRows)): VEC consecutive vocab entries per thread (128-bit /
//     64-bit accesses when the logits rows are 4-element aligned, which the padded LM-head output is) for kDlRows
//     consecutive rows: the per-column operands (col_lse, colsum) are loaded once, and the rows' loads are all in
//     flight before the first one is used.
constexpr int kDlRows = 1;   // (4 rows per thread measured no faster: 156 vs 148 us at the repo shape)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) loss_dlogits_kernel(const LossParams p) {
  const int row0 = blockIdx.y * kDlRows;
  const int nrows = p.B * p.T;
  const int v0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (v0 >= p.V) return;
  const float scale = (p.grad_out ? p.grad_out[0] : 1.f) / (float)(p.B * p.T);
  const bool vec = VEC == 4 && v0 + 4 <= p.V;
  float x[kDlRows][VEC], w[kDlRows][VEC], rl[kDlRows];
  int tg[kDlRows], bb[kDlRows];
#pragma unroll
  for (int r = 0; r < kDlRows; ++r) {
    const int row = min(row0 + r, nrows - 1);
    const int b = row / p.T, t = row - b * p.T;
    bb[r] = b;
    const T* xr = static_cast<const T*>(p.logits) + (int64_t)b * p.l_bs + (int64_t)t * p.l_ts + v0;
    const float* wr = p.table + (int64_t)bucket_of(p.src[row], p.bnd) * p.V + v0;
    rl[r] = p.row_lse[row];
    tg[r] = (int)p.trg[row] - v0;
    if (vec) {
      float xv[4];
      V4<T>::ld(xr, xv);
#pragma unroll
      for (int e = 0; e < VEC; ++e) x[r][e] = xv[e < 4 ? e : 0];
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) x[r][e] = v0 + e < p.V ? IO<T>::ld(xr + e) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) w[r][e] = v0 + e < p.V ? wr[e] : 0.f;
  }
  // per-column operands: the kDlRows rows share a batch element unless they straddle a boundary
  float cl[VEC], cs[VEC];
  int cur_b = -1;
#pragma unroll
  for (int r = 0; r < kDlRows; ++r) {
    if (row0 + r >= nrows) break;
    if (bb[r] != cur_b) {
      cur_b = bb[r];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const bool ok = v0 + e < p.V;
        cl[e] = ok ? p.col_lse[(int64_t)cur_b * p.V + v0 + e] : 0.f;
        cs[e] = ok ? p.colsum[(int64_t)cur_b * p.V + v0 + e] : 0.f;
      }
    }
    const int row = row0 + r;
    const int t = row - cur_b * p.T;
    T* gr = static_cast<T*>(p.dlogits) + (int64_t)cur_b * p.d_bs + (int64_t)t * p.d_ts + v0;
    float g[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (v0 + e < p.V) {
        const float dlp = dlp_of<T>(p, x[r][e], cl[e], w[r][e], rl[r], tg[r] == e, scale);
        g[e] = dlp - exp_fast(x[r][e] - cl[e]) * cs[e];
      } else {
        g[e] = 0.f;
      }
    }
    if (vec) {
      float gv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) gv[e] = g[e < VEC ? e : 0];
      V4<T>::st_global(gr, gv);
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        if (v0 + e < p.V) IO<T>::st(gr + e, g[e]);
    }
  }
}

static int loss_fill(const MambaLossArgs* a, LossParams& p, bool bwd) {
  if (!a || a->struct_size != (int32_t)sizeof(MambaLossArgs))
    return set_error(MAMBA_EINVAL, "filtered_ce: bad args pointer or struct_size");
  if (a->batch <= 0 || a->seqlen <= 0 || a->vocab <= 0)
    return set_error(MAMBA_EINVAL, "filtered_ce: batch/seqlen/vocab must be positive");
  if (a->dtype != MAMBA_F32 && a->dtype != MAMBA_BF16) return set_error(MAMBA_EDTYPE, "filtered_ce: dtype %d", a->dtype);
  if (!a->logits || !a->src || !a->trg || !a->table || !a->col_lse || !a->row_lse || !a->workspace)
    return set_error(MAMBA_EINVAL, "filtered_ce: null pointer");
  if ((int64_t)a->batch * a->seqlen > 65535) return set_error(MAMBA_ESIZE, "filtered_ce: more than 65535 rows");
  p.B = a->batch, p.T = a->seqlen, p.V = a->vocab;
  p.S = ceil_div(p.T, kLossSeg);
  const size_t need = mamba_filtered_ce_workspace_bytes(a->batch, a->seqlen, a->vocab);
  if (a->workspace_bytes < need)
    return set_error(MAMBA_ESIZE, "filtered_ce: workspace %zu B < required %zu B", a->workspace_bytes, need);
  p.logits = a->logits, p.l_bs = a->logits_bs, p.l_ts = a->logits_ts;
  p.src = a->src, p.trg = a->trg, p.table = a->table;
  for (int i = 0; i < 4; ++i) p.bnd[i] = a->boundaries[i];
  float* ws = static_cast<float*>(a->workspace);
  const size_t part = (size_t)p.B * p.S * p.V;
  p.col_part_m = ws, p.col_part_s = ws + part;
  p.colsum_part = ws;               // backward reuses the same region
  p.colsum = ws + part;             // [B][V] fits in the second half
  p.row_loss = ws + 2 * part;       // [B][T]
  p.col_lse = a->col_lse, p.row_lse = a->row_lse, p.loss = a->loss;
  p.grad_out = a->grad_out, p.dlogits = a->dlogits, p.d_bs = a->dlogits_bs, p.d_ts = a->dlogits_ts;
  if (!bwd && !a->loss) return set_error(MAMBA_EINVAL, "filtered_ce_fwd: null loss");
  if (bwd && !a->dlogits) return set_error(MAMBA_EINVAL, "filtered_ce_bwd: null dlogits");
  return MAMBA_OK;
}

template <typename T>
static int loss_fwd_launch(const LossParams& p, cudaStream_t st) {
  dim3 gcol(ceil_div(p.V, kLossColThreads), p.S, p.B);
  loss_col_lse_kernel<T><<<gcol, kLossColThreads, 0, st>>>(p);
  loss_col_lse_combine_kernel<<<dim3(ceil_div(p.V, 256), p.B), 256, 0, st>>>(p);
  loss_row_kernel<T><<<p.B * p.T, kLossRowThreads, 0, st>>>(p);
  loss_mean_kernel<<<1, 256, 0, st>>>(p);
  count_launch(4);
  return check_launch("filtered_ce_fwd");
}
template <typename T>
static int loss_bwd_launch(const LossParams& p, cudaStream_t st) {
  dim3 gcol(ceil_div(p.V, kLossColThreads), p.S, p.B);
  loss_colsum_kernel<T><<<gcol, kLossColThreads, 0, st>>>(p);
  loss_colsum_combine_kernel<<<dim3(ceil_div(p.V, 256), p.B), 256, 0, st>>>(p);
  const size_t elt = sizeof(T);
  const bool vec = reinterpret_cast<uintptr_t>(p.logits) % (4 * elt) == 0 && reinterpret_cast<uintptr_t>(p.dlogits) % (4 * elt) == 0 &&
                   p.l_bs % 4 == 0 && p.l_ts % 4 == 0 && p.d_bs % 4 == 0 && p.d_ts % 4 == 0;
  if (vec)
    loss_dlogits_kernel<T, 4><<<dim3(ceil_div(p.V, 256 * 4), ceil_div(p.B * p.T, kDlRows)), 256, 0, st>>>(p);
  else
    loss_dlogits_kernel<T, 1><<<dim3(ceil_div(p.V, 256), ceil_div(p.B * p.T, kDlRows)), 256, 0, st>>>(p);
  count_launch(3);
  return check_launch("filtered_ce_bwd");
}

}  // namespace mb

extern "C" size_t mamba_filtered_ce_workspace_bytes(int batch, int seqlen, int vocab) {
  if (batch <= 0 || seqlen <= 0 || vocab <= 0) return 0;
  const size_t S = (size_t)mb::ceil_div(seqlen, mb::kLossSeg);
  const size_t part = (size_t)batch * S * vocab;
  return 4 * (2 * part + (size_t)batch * seqlen);
}
extern "C" int mamba_filtered_ce_fwd(const MambaLossArgs* a, void* stream) {
  mb::LossParams p{};
  int rc = mb::loss_fill(a, p, false);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return a->dtype == MAMBA_F32 ? mb::loss_fwd_launch<float>(p, st) : mb::loss_fwd_launch<__nv_bfloat16>(p, st);
}
extern "C" int mamba_filtered_ce_bwd(const MambaLossArgs* a, void* stream) {
  mb::LossParams p{};
  int rc = mb::loss_fill(a, p, true);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return a->dtype == MAMBA_F32 ? mb::loss_bwd_launch<float>(p, st) : mb::loss_bwd_launch<__nv_bfloat16>(p, st);
}
