// decode.cu — one new token through the WHOLE model in one persistent kernel, sm_100a.
//
// The per-token chain of the recurrent decode (no counterpart in the reference, which re-runs the full model per
// token: scripts/generate.py:26-31) is, per layer,  RMSNorm(+residual) -> in_proj -> conv step -> x_proj -> dt_proj +
// softplus + SSM step + D skip + gate -> out_proj  (simple_mamba.pyc @L179, @L228-245 for one position), then the
// final norm and the LM head (@L94).  Launched as separate kernels that is 42 launches of 5-25 us whose time is
// launch latency, prologues and tails, not the 450 MB of weights and state they stream (profiles/r02_decode_*).
// Here ONE cooperative grid (one CTA per SM, all co-resident) walks the phases and meets at a grid barrier between
// them (4 per layer); while a CTA waits at a barrier its warps have already asked L2 for the first weight rows of
// the next phase.  Activations between phases live in a small global scratch (read with ld.global.cg: they were
// written by other SMs in the same launch); the residual stream is fp32; weights are fp32 or bf16; batch <= 16.
//
// Work split of a linear phase y[b, n] = sum_k W[n, k] x[b, k]: the activation rows sit in shared memory as fp32;
// a warp owns RW consecutive weight rows per task, lanes stride the K axis with 16-byte loads (every weight byte is
// read once, all of a task's loads are issued before the first multiply-add), one butterfly per (row, sequence).
#include <cooperative_groups.h>

#include "common.cuh"

namespace mb {

namespace {

constexpr int kDecThreads = 512;
constexpr int kDecWarps = kDecThreads / 32;
constexpr int kMaxB = 16;

// L2 loads (no L1 allocation): the scratch was written by other SMs earlier in this launch.  Plain intrinsics, so that
// the compiler can batch them; the grid barrier's fences and "memory" clobbers keep every load inside its phase.
__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// ---- grid barrier: monotone generation counter, one arrival per CTA ------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks, unsigned int& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++gen;
    __threadfence();
    const unsigned int arrived = atomicAdd(bar, 1u) + 1u;
    if (arrived == gen * nblocks) {
      asm volatile("st.global.release.gpu.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen) : "memory");
    } else {
      unsigned int g;
      do {
        asm volatile("ld.global.acquire.gpu.u32 %0, [%1];" : "=r"(g) : "l"(bar + 1) : "memory");
      } while (g < gen);
    }
    __threadfence();
  }
  __syncthreads();
}

template <typename TW>
__device__ __forceinline__ void ldw4(const TW* p, float (&w)[4]);
template <>
__device__ __forceinline__ void ldw4<float>(const float* p, float (&w)[4]) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
}
template <>
__device__ __forceinline__ void ldw4<__nv_bfloat16>(const __nv_bfloat16* p, float (&w)[4]) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  w[0] = __uint_as_float(v.x << 16), w[1] = __uint_as_float(v.x & 0xffff0000u);
  w[2] = __uint_as_float(v.y << 16), w[3] = __uint_as_float(v.y & 0xffff0000u);
}

// pull the first weight rows this warp will own in the NEXT phase into L2 (issued before the grid barrier)
template <typename TW>
__device__ __forceinline__ void prefetch_rows(const TW* W, int N, int K, int RW, int gw, int nwarps_total, int lane) {
  const size_t row_bytes = (size_t)K * sizeof(TW);
  for (int task = gw; task * RW < N; task += nwarps_total) {
    for (int r = 0; r < RW; ++r) {
      const int n = task * RW + r;
      if (n < N) {
        const char* p = reinterpret_cast<const char*>(W) + (size_t)n * row_bytes;
        for (size_t off = (size_t)lane * 128; off < row_bytes; off += 32 * 128) prefetch_l2(p + off);
      }
    }
    break;  // first task only: the rest streams behind it
  }
}

// One linear phase.  xs: [B][K] fp32 in shared memory.  epi(n, b, value) is called by lane b of the owning warp.
// The weight rows of a warp's FIRST task are loaded by load_task() before the activations are staged (they do not
// depend on the previous phase), so the DRAM round trip of the weights overlaps the L2 round trip of the staging.
template <typename TW, int RW, int KI>
struct WTile {
  float w[RW][KI][4];
};
template <typename TW, int RW, int KI>
__device__ __forceinline__ void load_task(WTile<TW, RW, KI>& t, const TW* __restrict__ W, int N, int K, int task, int lane) {
  const int n0 = task * RW;
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const TW* wr = W + (size_t)min(n0 + r, N - 1) * K;
#pragma unroll
    for (int i = 0; i < KI; ++i) ldw4<TW>(wr + 4 * (lane + 32 * i), t.w[r][i]);
  }
}
template <typename TW, int RW, int KI, typename Epi>
__device__ __forceinline__ void linear_phase(WTile<TW, RW, KI>& t, const TW* __restrict__ W, const TW* __restrict__ bias, int N,
                                             int K, int B, const float* xs, int gw, int nwarps_total, int lane, Epi epi) {
  // KI = K / 128: 16-byte steps per lane (compile time so that all weight loads of a task are in flight together)
  for (int task = gw; task * RW < N; task += nwarps_total) {
    const int n0 = task * RW;
    if (task != gw) load_task<TW, RW, KI>(t, W, N, K, task, lane);
    float acc[RW][kMaxB];
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int b = 0; b < kMaxB; ++b) acc[r][b] = 0.f;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
#pragma unroll
      for (int b = 0; b < kMaxB; ++b) {
        if (b < B) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + b * K + 4 * (lane + 32 * i));
#pragma unroll
          for (int r = 0; r < RW; ++r)
            acc[r][b] = fmaf(t.w[r][i][3], xv.w, fmaf(t.w[r][i][2], xv.z, fmaf(t.w[r][i][1], xv.y, fmaf(t.w[r][i][0], xv.x, acc[r][b]))));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      float mine = 0.f;
#pragma unroll
      for (int b = 0; b < kMaxB; ++b) {
        if (b < B) {
          float v = acc[r][b];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == b) mine = v;
        }
      }
      const int n = n0 + r;
      if (n < N && lane < B) epi(n, lane, mine + (bias ? IO<TW>::ld(bias + n) : 0.f));
    }
  }
}

// stage s = hidden + resid (or the embedding rows for the first layer), write s as the new residual (CTA 0), and leave
// rmsnorm(s) * w in shared memory:  `normed, resid = norm(hidden, resid)`  (simple_mamba.pyc @L179 / @L346)
template <typename TW, int KI>
__device__ __forceinline__ void stage_norm(float* xs, const float* hidden, const float* resid_in, float* resid_out,
                                           const TW* emb, const int64_t* tok, const float* norm_w, float eps, int B, int K) {
  // one warp per sequence; the lane's KI 16-byte pieces of the row are all loaded before the first use (one L2 round
  // trip per row instead of one per piece)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < B; b += kDecWarps) {
    float* row = xs + b * K;
    float4 v[KI], g[KI];
    if (emb != nullptr) {
      const TW* er = emb + (size_t)tok[b] * K;
#pragma unroll
      for (int i = 0; i < KI; ++i) {
        float e[4];
        ldw4<TW>(er + 4 * (lane + 32 * i), e);
        v[i] = make_float4(e[0], e[1], e[2], e[3]);
      }
    } else {
      float4 h[KI], r[KI];
#pragma unroll
      for (int i = 0; i < KI; ++i) h[i] = ldcg4(hidden + (size_t)b * K + 4 * (lane + 32 * i)), r[i] = ldcg4(resid_in + (size_t)b * K + 4 * (lane + 32 * i));
#pragma unroll
      for (int i = 0; i < KI; ++i) v[i] = make_float4(h[i].x + r[i].x, h[i].y + r[i].y, h[i].z + r[i].z, h[i].w + r[i].w);
    }
#pragma unroll
    for (int i = 0; i < KI; ++i) g[i] = __ldg(reinterpret_cast<const float4*>(norm_w) + lane + 32 * i);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      if (blockIdx.x == 0 && resid_out) *reinterpret_cast<float4*>(resid_out + (size_t)b * K + 4 * (lane + 32 * i)) = v[i];
      ss = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, ss))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / (float)K + eps);
#pragma unroll
    for (int i = 0; i < KI; ++i)
      *reinterpret_cast<float4*>(row + 4 * (lane + 32 * i)) =
          make_float4(v[i].x * rstd * g[i].x, v[i].y * rstd * g[i].y, v[i].z * rstd * g[i].z, v[i].w * rstd * g[i].w);
  }
}

__device__ __forceinline__ void stage_rows(float* xs, const float* src, int B, int K) {
  // up to 8 independent 16-byte loads per thread in flight per round (16 x 2048 floats = 16 per thread at most)
  const int total = B * K;
  for (int base = threadIdx.x * 4; base < total; base += kDecThreads * 4 * 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * kDecThreads * 4;
      v[u] = i < total ? ldcg4(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * kDecThreads * 4;
      if (i < total) *reinterpret_cast<float4*>(xs + i) = v[u];
    }
  }
}

template <typename TW>
__global__ void __launch_bounds__(kDecThreads, 1) decode_token_kernel(const MambaDecodeTokenArgs a) {
  extern __shared__ __align__(16) float xs[];  // [B][max(d_model, d_inner)]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nblk = gridDim.x;
  const int gw = warp * nblk + blockIdx.x;     // warp w of every CTA before warp w+1 of any: tasks spread over the SMs
  const int nwt = nblk * kDecWarps;
  const int B = a.batch, dm = a.d_model, di = a.d_inner, N = a.d_state, R = a.dt_rank, KC = a.d_conv;
  const int XD = R + 2 * N;
#ifdef MB_DEC_PROFILE
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(a.barrier + 4);
  int nstamp = 0;
  auto stamp = [&]() {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      stamps[nstamp] = t;
    }
    ++nstamp;
  };
#else
  auto stamp = []() {};
#endif
  stamp();
  unsigned int gen = 0;                         // the host zeroes both barrier words before every launch
  unsigned int* bar = a.barrier;
  // scratch carve-up (fp32)
  float* resid[2] = {a.scratch, a.scratch + (size_t)kMaxB * dm};
  float* hidden = resid[1] + (size_t)kMaxB * dm;
  float* zbuf = hidden + (size_t)kMaxB * dm;
  float* xc = zbuf + (size_t)kMaxB * di;
  float* xdbl = xc + (size_t)kMaxB * di;
  float* ybuf = xdbl + (size_t)kMaxB * XD;
  const TW* emb = static_cast<const TW*>(a.embedding);

  for (int l = 0; l < a.n_layers; ++l) {
    const MambaDecodeLayer L = a.layers[l];
    const TW* w_in = static_cast<const TW*>(L.in_proj_weight);
    // ---- phase 1: norm + in_proj (+ conv step on the x half) ---------------------------------------------------------
    auto epi1 = [&](int n, int b, float v) {
      if (n < di) {  // conv branch: shift register + SiLU (simple_mamba.pyc @L233-237 for one position)
        float* st = L.conv_state + ((size_t)b * di + n) * KC;
        float acc = L.conv_bias ? L.conv_bias[n] : 0.f;
        for (int k = 0; k < KC; ++k) {
          const float sv = (k + 1 < KC) ? st[k + 1] : v;
          st[k] = sv;
          acc = fmaf(L.conv_weight[(size_t)n * KC + k], sv, acc);
        }
        xc[(size_t)b * di + n] = silu_f(acc);
      } else {
        zbuf[(size_t)b * di + (n - di)] = v;
      }
    };
    if (dm == 1024) {
      WTile<TW, 2, 8> t;
      if (gw * 2 < 2 * di) load_task<TW, 2, 8>(t, w_in, 2 * di, dm, gw, lane);
      stage_norm<TW, 8>(xs, hidden, resid[(l + 1) & 1], resid[l & 1], l == 0 ? emb : nullptr, a.token, L.norm_weight, a.eps, B, dm);
      __syncthreads();
      linear_phase<TW, 2, 8>(t, w_in, static_cast<const TW*>(L.in_proj_bias), 2 * di, dm, B, xs, gw, nwt, lane, epi1);
    } else {
      WTile<TW, 1, 1> t;
      if (gw < 2 * di) load_task<TW, 1, 1>(t, w_in, 2 * di, dm, gw, lane);
      stage_norm<TW, 1>(xs, hidden, resid[(l + 1) & 1], resid[l & 1], l == 0 ? emb : nullptr, a.token, L.norm_weight, a.eps, B, dm);
      __syncthreads();
      linear_phase<TW, 1, 1>(t, w_in, static_cast<const TW*>(L.in_proj_bias), 2 * di, dm, B, xs, gw, nwt, lane, epi1);
    }
    prefetch_rows<TW>(static_cast<const TW*>(L.x_proj_weight), XD, di, 1, gw, nwt, lane);
    stamp();
    grid_barrier(bar, nblk, gen);
    stamp();
    // ---- phase 2: x_proj ------------------------------------------------------------------------------------------------
    auto epi2 = [&](int n, int b, float v) { xdbl[(size_t)b * XD + n] = v; };
    const TW* w_x = static_cast<const TW*>(L.x_proj_weight);
    if (di == 2048) {
      WTile<TW, 1, 16> t;
      if (gw < XD) load_task<TW, 1, 16>(t, w_x, XD, di, gw, lane);
      stage_rows(xs, xc, B, di);                // (every CTA owns at least one of the R + 2N rows: warp 0)
      __syncthreads();
      linear_phase<TW, 1, 16>(t, w_x, (const TW*)nullptr, XD, di, B, xs, gw, nwt, lane, epi2);
    } else {
      WTile<TW, 1, 2> t;
      if (gw < XD) load_task<TW, 1, 2>(t, w_x, XD, di, gw, lane);
      stage_rows(xs, xc, B, di);
      __syncthreads();
      linear_phase<TW, 1, 2>(t, w_x, (const TW*)nullptr, XD, di, B, xs, gw, nwt, lane, epi2);
    }
    prefetch_rows<TW>(static_cast<const TW*>(L.out_proj_weight), dm, di, 1, gw, nwt, lane);
    stamp();
    grid_barrier(bar, nblk, gen);
    stamp();
    // ---- phase 3: dt_proj + softplus + SSM step + D skip + gate; a half-warp per (b, d) ------------------------------------
    {
      const int hl = threadIdx.x & 15;
      const unsigned mask = 0xffffu << (threadIdx.x & 16);
      const int64_t items = (int64_t)B * di;
      const int64_t hw0 = ((int64_t)(threadIdx.x >> 4)) * nblk + blockIdx.x, nhw = (int64_t)nblk * (kDecThreads >> 4);
      for (int64_t item = hw0; item < items; item += nhw) {
        const int b = (int)(item / di), d = (int)(item - (int64_t)b * di);
        float dot = 0.f;
        const float* wdt = L.dt_weight + (size_t)d * R;
        const float* xd = xdbl + (size_t)b * XD;
        for (int r4 = hl; r4 < (R >> 2); r4 += 16) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wdt) + r4);
          const float4 xv = ldcg4(xd + 4 * r4);
          dot = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, dot))));
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(mask, dot, o);
        dot += L.dt_bias ? L.dt_bias[d] : 0.f;
        const float delta = softplus_fast(dot);
        const float xcv = ldcg(xc + (size_t)b * di + d);
        const float du = delta * xcv, dl2 = delta * kLog2e;
        float* h = L.ssm_state + ((size_t)b * di + d) * N;
        const float* A = L.A + (size_t)d * N;
        float y = 0.f;
        for (int n4 = hl; n4 < (N >> 2); n4 += 16) {
          const float4 a4 = __ldg(reinterpret_cast<const float4*>(A) + n4);
          float4 h4 = reinterpret_cast<float4*>(h)[n4];
          const float4 bb = ldcg4(xd + R + 4 * n4), cc = ldcg4(xd + R + N + 4 * n4);
          h4.x = fmaf(ex2_approx(dl2 * a4.x), h4.x, du * bb.x);
          h4.y = fmaf(ex2_approx(dl2 * a4.y), h4.y, du * bb.y);
          h4.z = fmaf(ex2_approx(dl2 * a4.z), h4.z, du * bb.z);
          h4.w = fmaf(ex2_approx(dl2 * a4.w), h4.w, du * bb.w);
          reinterpret_cast<float4*>(h)[n4] = h4;
          y = fmaf(h4.x, cc.x, fmaf(h4.y, cc.y, fmaf(h4.z, cc.z, fmaf(h4.w, cc.w, y))));
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) y += __shfl_xor_sync(mask, y, o);
        if (hl == 0) {
          y = fmaf(L.D ? L.D[d] : 0.f, xcv, y);
          y *= silu_fast(ldcg(zbuf + (size_t)b * di + d));
          ybuf[(size_t)b * di + d] = y;
        }
      }
    }
    stamp();
    grid_barrier(bar, nblk, gen);
    stamp();
    // ---- phase 4: out_proj -----------------------------------------------------------------------------------------------
    auto epi4 = [&](int n, int b, float v) { hidden[(size_t)b * dm + n] = v; };
    const TW* w_o = static_cast<const TW*>(L.out_proj_weight);
    if (di == 2048) {
      WTile<TW, 1, 16> t;
      if (gw < dm) load_task<TW, 1, 16>(t, w_o, dm, di, gw, lane);
      stage_rows(xs, ybuf, B, di);
      __syncthreads();
      linear_phase<TW, 1, 16>(t, w_o, static_cast<const TW*>(L.out_proj_bias), dm, di, B, xs, gw, nwt, lane, epi4);
    } else {
      WTile<TW, 1, 2> t;
      if (gw < dm) load_task<TW, 1, 2>(t, w_o, dm, di, gw, lane);
      stage_rows(xs, ybuf, B, di);
      __syncthreads();
      linear_phase<TW, 1, 2>(t, w_o, static_cast<const TW*>(L.out_proj_bias), dm, di, B, xs, gw, nwt, lane, epi4);
    }
    if (l + 1 < a.n_layers)
      prefetch_rows<TW>(static_cast<const TW*>(a.layers[l + 1].in_proj_weight), 2 * di, dm, dm == 1024 ? 2 : 1, gw, nwt, lane);
    else
      prefetch_rows<TW>(static_cast<const TW*>(a.head_weight), a.vocab, dm, dm == 1024 ? 2 : 1, gw, nwt, lane);
    stamp();
    grid_barrier(bar, nblk, gen);
    stamp();
  }
  // ---- final norm + LM head ----------------------------------------------------------------------------------------------
  auto epih = [&](int n, int b, float v) { a.logits[(size_t)b * a.logits_bs + n] = v; };
  const TW* w_h = static_cast<const TW*>(a.head_weight);
  if (dm == 1024) {
    WTile<TW, 2, 8> t;
    if (gw * 2 < a.vocab) load_task<TW, 2, 8>(t, w_h, a.vocab, dm, gw, lane);
    stage_norm<TW, 8>(xs, hidden, resid[(a.n_layers + 1) & 1], nullptr, (const TW*)nullptr, a.token, a.norm_f_weight, a.eps, B, dm);
    __syncthreads();
    linear_phase<TW, 2, 8>(t, w_h, static_cast<const TW*>(a.head_bias), a.vocab, dm, B, xs, gw, nwt, lane, epih);
  } else {
    WTile<TW, 1, 1> t;
    if (gw < a.vocab) load_task<TW, 1, 1>(t, w_h, a.vocab, dm, gw, lane);
    stage_norm<TW, 1>(xs, hidden, resid[(a.n_layers + 1) & 1], nullptr, (const TW*)nullptr, a.token, a.norm_f_weight, a.eps, B, dm);
    __syncthreads();
    linear_phase<TW, 1, 1>(t, w_h, static_cast<const TW*>(a.head_bias), a.vocab, dm, B, xs, gw, nwt, lane, epih);
  }
  stamp();
}

}  // namespace

}  // namespace mb

extern "C" size_t mamba_decode_token_scratch_bytes(int d_model, int d_inner, int d_state, int dt_rank) {
  if (d_model <= 0 || d_inner <= 0 || d_state <= 0 || dt_rank <= 0) return 0;
  return sizeof(float) * (size_t)mb::kMaxB * (3 * (size_t)d_model + 3 * (size_t)d_inner + dt_rank + 2 * d_state) + 256;
}

extern "C" int mamba_decode_token(const MambaDecodeTokenArgs* a, void* stream) {
  using namespace mb;
  if (!a || a->struct_size != (int32_t)sizeof(MambaDecodeTokenArgs))
    return set_error(MAMBA_EINVAL, "decode_token: bad args pointer or struct_size");
  if (a->batch <= 0 || a->batch > kMaxB) return set_error(MAMBA_ESIZE, "decode_token: batch %d outside 1..%d", a->batch, kMaxB);
  if (a->n_layers <= 0 || !a->layers || !a->token || !a->embedding || !a->head_weight || !a->norm_f_weight || !a->logits ||
      !a->scratch || !a->barrier)
    return set_error(MAMBA_EINVAL, "decode_token: null pointer");
  // supported shapes: (d_model, d_inner) = (1024, 2048) [the repo's model] or (128, 256) [the tests' small model];
  // d_state and dt_rank multiples of 4
  const bool big = a->d_model == 1024 && a->d_inner == 2048, small = a->d_model == 128 && a->d_inner == 256;
  if (!(big || small) || a->d_state % 4 || a->dt_rank % 4 || a->d_conv < 1 || a->d_conv > 8)
    return set_error(MAMBA_ESIZE, "decode_token: unsupported shape d_model %d d_inner %d d_state %d dt_rank %d", a->d_model,
                     a->d_inner, a->d_state, a->dt_rank);
  if (a->scratch_bytes < mamba_decode_token_scratch_bytes(a->d_model, a->d_inner, a->d_state, a->dt_rank))
    return set_error(MAMBA_ESIZE, "decode_token: scratch too small");
  if (a->w_dtype != MAMBA_F32 && a->w_dtype != MAMBA_BF16) return set_error(MAMBA_EDTYPE, "decode_token: w_dtype %d", a->w_dtype);
  const size_t smem = sizeof(float) * (size_t)a->batch * (size_t)(a->d_inner > a->d_model ? a->d_inner : a->d_model);
  int dev = 0, nsm = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nsm), cfg.blockDim = dim3(kDecThreads), cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // every CTA resident at once: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  cudaError_t e = cudaMemsetAsync(a->barrier, 0, 2 * sizeof(unsigned int), cfg.stream);   // grid-barrier state
  if (e != cudaSuccess) return set_error(MAMBA_ELAUNCH, "decode_token: memset: %s", cudaGetErrorString(e));
  if (a->w_dtype == MAMBA_F32) {
    static thread_local SmemConfig c32;
    if (int rc = ensure_dynamic_smem(decode_token_kernel<float>, smem, c32, "decode_token")) return rc;
    e = cudaLaunchKernelEx(&cfg, decode_token_kernel<float>, *a);
  } else {
    static thread_local SmemConfig c16;
    if (int rc = ensure_dynamic_smem(decode_token_kernel<__nv_bfloat16>, smem, c16, "decode_token")) return rc;
    e = cudaLaunchKernelEx(&cfg, decode_token_kernel<__nv_bfloat16>, *a);
  }
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return set_error(MAMBA_ELAUNCH, "decode_token: %s", cudaGetErrorString(e));
  }
  count_launch();
  return check_launch("decode_token");
}
