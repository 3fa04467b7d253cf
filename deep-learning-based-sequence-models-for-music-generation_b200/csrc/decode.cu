// decode.cu — one new token through the WHOLE model in one persistent kernel, sm_100a.
//
// The per-token chain of the recurrent decode (no counterpart in the reference, which re-runs the full model per
// token: scripts/generate.py:26-31) is, per layer,  RMSNorm(+residual) -> in_proj -> conv step -> x_proj -> dt_proj +
// softplus + SSM step + D skip + gate -> out_proj  (simple_mamba.pyc @L179, @L228-245 for one position), then the
// final norm and the LM head (@L94).  Launched as separate kernels that is 42 launches of 5-25 us whose time is
// launch latency, prologues and tails, not the 450 MB of weights and state they stream (profiles/r02_decode_*).
// Here ONE cooperative grid (one CTA per SM, all co-resident) walks the phases and meets at a grid barrier between
// them (4 per layer).  What bounds a phase is latency (DRAM round trip of the weights, L2 round trip of the
// activations other SMs wrote, the barrier itself), so the kernel is organised around taking those off the critical
// path:
//   * the barrier is split into arrive (one release-RED per CTA) and wait (acquire-poll of the same counter by thread
//     0); between the two every warp issues what the NEXT phases need and does not depend on this one: L2 prefetches
//     of weight rows (one or two phases ahead), of the SSM state rows, A and dt_proj rows, and register loads of the
//     first weight tiles / conv state / state rows.  The polling thread must not have DRAM loads outstanding (its
//     acquire waits for them): whatever is loaded into registers before the poll was prefetched into L2 a phase earlier;
//   * phases with many rows per CTA (in_proj, head; "rows" mapping) give a warp four weight rows and the whole K axis
//     (lanes stride K with 16-byte loads; activations normalised once into shared memory by all threads).  K is cut
//     in four chunks through a ring of two register tiles, so that the next chunk — of this task or of the warp's next
//     one — is in flight while this one is multiplied; packed fp32x2 accumulation; one warp reduce-scatter per task;
//   * phases with few rows per CTA (x_proj: 192 rows, out_proj: 1024 rows over 148 CTAs; "K-split" mapping) split K
//     over the CTA's 8 warps instead: a warp reads ITS slice of the activations straight from L2 into registers (no
//     staging of the full 80 KB row block, no __syncthreads before the math), lanes = 2 batch halves x 16 K-lanes,
//     reduce-scatter over the K-lanes, a fixed-order sum over the warps through shared memory.  out_proj's epilogue
//     adds into the residual stream in place;
//   * the SSM phase: one thread per (channel, sequence) forms delta = softplus(<dt_proj row, x_dbl> + bias) once, then 16
//     lanes per state row (4 states each) do the update; the state rows are in registers before the barrier opens.
// Activations between phases live in a small global scratch (read with ld.global.cg: they were written by other SMs in
// the same launch); the residual stream is fp32; weights are fp32 or bf16; batch <= 16 (template BMAX = batch rounded
// up to even; padded rows are zeros and their outputs are dropped).
#include "common.cuh"

namespace mb {

namespace {

constexpr int kDecThreads = 256;   // 8 warps x up to 255 registers: the phases keep whole weight tiles and state rows in registers
constexpr int kDecWarps = kDecThreads / 32;
constexpr int kMaxB = 16;
constexpr int kStampWords = 4;  // barrier words before the time stamps (MAMBA_DECODE_FLAG_STAMPS)

// L2 loads (no L1 allocation): the scratch was written by other SMs earlier in this launch.
__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// ---- grid barrier: monotone arrival counter (zero at launch: the previous launch's last CTA resets it) -----------------------------------
__device__ __forceinline__ void grid_arrive(unsigned int* bar) {
  __syncthreads();   // every write of this CTA happens-before thread 0's release
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
}
__device__ __forceinline__ void grid_wait(unsigned int* bar, unsigned int target) {
  if (threadIdx.x == 0) {
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
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

__device__ __forceinline__ float dot4(const float (&w)[4], const float4 x, float acc) {
  return fmaf(w[3], x.w, fmaf(w[2], x.z, fmaf(w[1], x.y, fmaf(w[0], x.x, acc))));
}

// ---- "rows" mapping: a warp owns RW weight rows and the whole K axis ---------------------------------------------------------
// K = NQ chunks x KH x 128 floats; a tile holds rows x KH 16-byte pieces per lane of ONE chunk.  Two tiles form a ring:
// while chunk q is multiplied, chunk q+1 is already in registers or in flight, and chunk q+2 (of this task or of the
// warp's next one) is requested as soon as q's tile is free.
template <int RW, int KH>
struct WTile {
  float w[RW][KH][4];
};
template <typename TW, int RW, int KH>
__device__ __forceinline__ void load_chunk(WTile<RW, KH>& t, const TW* __restrict__ W, int N, int K, int task, int q, int lane) {
  const int n0 = task * RW;
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const TW* wr = W + (size_t)min(n0 + r, N - 1) * K + q * (KH * 128);
#pragma unroll
    for (int i = 0; i < KH; ++i) ldw4<TW>(wr + 4 * (lane + 32 * i), t.w[r][i]);
  }
}
template <int RW, int KH, int BMAX>
__device__ __forceinline__ void mul_chunk(const WTile<RW, KH>& t, const float* xs, int K, int q, int lane, float2 (&acc)[RW][BMAX]) {
  // packed fp32x2 (FFMA2): .x accumulates the even k of the lane's 16-byte pieces, .y the odd ones
#pragma unroll
  for (int i = 0; i < KH; ++i) {
#pragma unroll
    for (int b = 0; b < BMAX; ++b) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + b * K + q * (KH * 128) + 4 * (lane + 32 * i));
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        acc[r][b] = __ffma2_rn(make_float2(t.w[r][i][0], t.w[r][i][1]), make_float2(xv.x, xv.y), acc[r][b]);
        acc[r][b] = __ffma2_rn(make_float2(t.w[r][i][2], t.w[r][i][3]), make_float2(xv.z, xv.w), acc[r][b]);
      }
    }
  }
}
// Sum NV values over the 32 lanes with NV - 1 + ... shuffles instead of 5 per value: at every step a lane keeps one half
// of its values and sends the other half to its partner (NV -> ceil(NV/2) -> ...).  Afterwards lane L holds, in
// v[0 .. NF), the totals of the values rs_index<NV>(L, j) (or nothing, if that index is >= NV).
template <int NV>
struct RsPlan {
  static constexpr int n1 = (NV + 1) / 2, n2 = (n1 + 1) / 2, n3 = (n2 + 1) / 2, n4 = (n3 + 1) / 2, n5 = (n4 + 1) / 2;
  static constexpr int NF = n5;
};
template <int N, int O>
__device__ __forceinline__ void rs_step(float* v, int lane) {
  constexpr int H = (N + 1) / 2;
  const bool up = lane & O;
#pragma unroll
  for (int j = 0; j < H; ++j) {
    const float lo = v[j], hi = (j + H < N) ? v[j + H] : 0.f;
    const float send = up ? lo : hi, keep = up ? hi : lo;
    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, O);
  }
}
template <int NV>
__device__ __forceinline__ void reduce_scatter(float (&v)[NV], int lane) {
  using P = RsPlan<NV>;
  rs_step<NV, 16>(v, lane);
  rs_step<P::n1, 8>(v, lane);
  rs_step<P::n2, 4>(v, lane);
  rs_step<P::n3, 2>(v, lane);
  rs_step<P::n4, 1>(v, lane);
}
// index of the value whose total lane `lane` holds in slot j (>= NV: none)
template <int NV>
__device__ __forceinline__ int rs_index(int lane, int j) {
  using P = RsPlan<NV>;
  int i = j;                                   // position among the n5 values after the last step
  bool ok = true;
  if (lane & 1) i += P::n5;
  ok = ok && i < P::n4;
  if (lane & 2) i += P::n4;
  ok = ok && i < P::n3;
  if (lane & 4) i += P::n3;
  ok = ok && i < P::n2;
  if (lane & 8) i += P::n2;
  ok = ok && i < P::n1;
  if (lane & 16) i += P::n1;
  return (ok && i < NV) ? i : NV;
}

// the same over the 16 lanes of a half-warp (offsets 8, 4, 2, 1): NV -> ... -> RsPlan16<NV>::NF values per lane
template <int NV>
struct RsPlan16 {
  static constexpr int n1 = (NV + 1) / 2, n2 = (n1 + 1) / 2, n3 = (n2 + 1) / 2, n4 = (n3 + 1) / 2;
  static constexpr int NF = n4;
};
template <int NV>
__device__ __forceinline__ void reduce_scatter16(float (&v)[NV], int lane) {
  using P = RsPlan16<NV>;
  rs_step<NV, 8>(v, lane);
  rs_step<P::n1, 4>(v, lane);
  rs_step<P::n2, 2>(v, lane);
  rs_step<P::n3, 1>(v, lane);
}
template <int NV>
__device__ __forceinline__ int rs_index16(int lane, int j) {
  using P = RsPlan16<NV>;
  int i = j;
  bool ok = true;
  if (lane & 1) i += P::n4;
  ok = ok && i < P::n3;
  if (lane & 2) i += P::n3;
  ok = ok && i < P::n2;
  if (lane & 4) i += P::n2;
  ok = ok && i < P::n1;
  if (lane & 8) i += P::n1;
  return (ok && i < NV) ? i : NV;
}

// One "rows" phase.  The first NB chunks of the warp's first task are loaded by the caller (before the grid barrier
// opens); pre(task) issues a task's other independent loads, epi(slot, r, n, b, value) is called by the lane that holds
// the total of (row r, sequence b) after the reduce-scatter (slot = rs slot of that lane).
template <typename TW, int RW, int KH, int NQ, int NB, int BMAX, typename Pre, typename Epi, typename Mark>
__device__ __forceinline__ void rows_phase(WTile<RW, KH> (&t)[NB], const TW* __restrict__ W, const TW* __restrict__ bias, int N, int K,
                                           int B, const float* xs, int gw, int nwt, int lane, Pre pre, Epi epi, Mark mark) {
  static_assert(NQ % NB == 0, "the ring position of a chunk must not depend on the task");
  constexpr int NV = RW * BMAX;
  for (int task = gw; task * RW < N; task += nwt) {
    if (task != gw) pre(task);
    const int nxt = task + nwt;
    const bool more = nxt * RW < N;
    float2 acc[RW][BMAX];
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int b = 0; b < BMAX; ++b) acc[r][b] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int q0 = 0; q0 < NQ; q0 += NB) {   // (rolled: the body is long and the instruction cache is small)
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int q = q0 + j;
        mul_chunk<RW, KH, BMAX>(t[j], xs, K, q, lane, acc);
        if (q + NB < NQ)
          load_chunk<TW, RW, KH>(t[j], W, N, K, task, q + NB, lane);
        else if (more)
          load_chunk<TW, RW, KH>(t[j], W, N, K, nxt, q + NB - NQ, lane);
      }
    }
    mark();
    float v[NV];
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int b = 0; b < BMAX; ++b) v[r * BMAX + b] = acc[r][b].x + acc[r][b].y;
    reduce_scatter<NV>(v, lane);
#pragma unroll
    for (int j = 0; j < RsPlan<NV>::NF; ++j) {
      const int i = rs_index<NV>(lane, j);
      const int r = i / BMAX, b = i - r * BMAX, n = task * RW + r;
      if (i < NV && n < N && b < B) epi(j, r, n, b, v[j] + (bias ? IO<TW>::ld(bias + n) : 0.f));
    }
  }
}

// ---- "K-split" mapping: the CTA owns rows blockIdx + j*gridDim (j < RG); warp w owns K-slice w ----------------------------------
// lane = bq*16 + kq: bq = batch parity (b = bq, bq+2, ...), kq = K-lane.  A warp's slice is K/warps floats, walked in NP
// pieces of KQ K-lanes x KP 16-byte loads; per piece the RG weight rows and the BH activation rows of the lane are
// loaded together (one L2 round trip: the weights were prefetched into L2 before the barrier opened), multiplied
// into acc[RG][BH]; four shuffle steps per output reduce over the K-lanes and the warp's partial sums go to
// red[warp][j][b].
template <typename TW, int KQ>
__device__ __forceinline__ void prefetch_rows_l2(const TW* __restrict__ W, int N, int K, int nrows, int warp, int lane) {
  // the warp's slice of rows j < nrows: K/warps elements each, one prefetch per 128 bytes
  const int bytes = (K / kDecWarps) * (int)sizeof(TW);
  const int pieces = (bytes + 127) >> 7;
  for (int i = lane; i < nrows * pieces; i += 32) {
    const int j = i / pieces, piece = i - j * pieces;
    const int n = blockIdx.x + j * gridDim.x;
    if (n < N) prefetch_l2(reinterpret_cast<const char*>(W + (size_t)n * K + warp * (K / kDecWarps)) + piece * 128);
  }
}
template <typename TW, int RG, int KQ, int KP, int NP, int BH, int BMAX, int RMAX>
__device__ __forceinline__ void ksplit_rows(const TW* __restrict__ W, int N, int K, const float* src, int B, float* red, int warp, int kq,
                                            int bq) {
  float acc[RG][BH];
#pragma unroll
  for (int r = 0; r < RG; ++r)
#pragma unroll
    for (int i = 0; i < BH; ++i) acc[r][i] = 0.f;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    float w[RG][KP][4];
    float4 x[BH][KP];
    const int k0 = warp * (K / kDecWarps) + 4 * (p * KP * KQ + kq);   // + 4 * KQ * j
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      const int n = blockIdx.x + r * gridDim.x;
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        if (n < N && kq < KQ)
          ldw4<TW>(W + (size_t)n * K + k0 + 4 * KQ * j, w[r][j]);
        else
          w[r][j][0] = w[r][j][1] = w[r][j][2] = w[r][j][3] = 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < BH; ++i) {
      const int b = bq + 2 * i;
#pragma unroll
      for (int j = 0; j < KP; ++j)
        x[i][j] = (b < B && kq < KQ) ? ldcg4(src + (size_t)b * K + k0 + 4 * KQ * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < RG; ++r)
#pragma unroll
      for (int i = 0; i < BH; ++i)
#pragma unroll
        for (int j = 0; j < KP; ++j) acc[r][i] = dot4(w[r][j], x[i][j], acc[r][i]);
  }
  // reduce over the 16 K-lanes of each batch-parity half: RG*BH values, reduce-scatter (one shuffle per value and step
  // on half as many values each step)
  constexpr int NV = RG * BH;
  float v[NV];
#pragma unroll
  for (int r = 0; r < RG; ++r)
#pragma unroll
    for (int i = 0; i < BH; ++i) v[r * BH + i] = acc[r][i];
  reduce_scatter16<NV>(v, kq);
#pragma unroll
  for (int j = 0; j < RsPlan16<NV>::NF; ++j) {
    const int idx = rs_index16<NV>(kq, j);
    const int r = idx / BH, i = idx - r * BH;
    if (idx < NV) red[(warp * RMAX + r) * BMAX + bq + 2 * i] = v[j];
  }
}
// fixed-order sum over the warps; thread (j, b) owns output row blockIdx + j*gridDim of sequence b
template <typename TW, int BMAX, int RMAX, typename Epi>
__device__ __forceinline__ void finish_rows(const float* red, const TW* __restrict__ bias, int N, int B, int nrows, Epi epi) {
  __syncthreads();
  const int j = threadIdx.x / BMAX, b = threadIdx.x - j * BMAX;
  const int n = blockIdx.x + j * gridDim.x;
  if (j < nrows && n < N && b < B) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) s += red[(w * RMAX + j) * BMAX + b];
    epi(n, b, s + (bias ? IO<TW>::ld(bias + n) : 0.f));
  }
}

// RMSNorm of the residual stream into shared memory:  `normed, resid = norm(hidden, resid)`  (simple_mamba.pyc @L179 /
// @L346).  The stream S = hidden + resid is kept in global memory by out_proj's epilogue (which adds its result in
// place); the first layer reads the embedding rows instead and CTA 0 writes them out as S.
// Thread t owns the 16-byte column pieces t, t + 256, ... of EVERY sequence: all of the CTA's loads are in flight
// together (one L2 round trip), the per-sequence sums of squares meet in shared memory (fixed order).
template <typename TW, int K, int BMAX>
__device__ __forceinline__ void stage_norm(float* xs, float* ssq, float* S, const TW* emb, const int64_t* tok, const float* norm_w,
                                           float eps, int B) {
  constexpr int C4 = K / 4;                                              // 16-byte pieces per row
  constexpr int NC = (C4 + kDecThreads - 1) / kDecThreads;              // ... per thread and row (1)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 g[NC];   // the norm weights of this thread's columns: requested first, used last
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = threadIdx.x + c * kDecThreads;
    g[c] = col < C4 ? __ldg(reinterpret_cast<const float4*>(norm_w) + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = threadIdx.x + c * kDecThreads;
    const bool on = col < C4;
    float4 v[BMAX];
#pragma unroll
    for (int b = 0; b < BMAX; ++b) {
      v[b] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (on && b < B) {
        if (emb != nullptr) {
          float e[4];
          ldw4<TW>(emb + (size_t)tok[b] * K + 4 * col, e);
          v[b] = make_float4(e[0], e[1], e[2], e[3]);
        } else {
          v[b] = ldcg4(S + (size_t)b * K + 4 * col);
        }
      }
    }
    float ss[BMAX];
#pragma unroll
    for (int b = 0; b < BMAX; ++b) {
      if (on && b < B) {
        if (emb != nullptr && blockIdx.x == 0) *reinterpret_cast<float4*>(S + (size_t)b * K + 4 * col) = v[b];
        *reinterpret_cast<float4*>(xs + b * K + 4 * col) = v[b];
      }
      ss[b] = fmaf(v[b].x, v[b].x, fmaf(v[b].y, v[b].y, fmaf(v[b].z, v[b].z, v[b].w * v[b].w)));
    }
    reduce_scatter<BMAX>(ss, lane);
#pragma unroll
    for (int j = 0; j < RsPlan<BMAX>::NF; ++j) {
      const int i = rs_index<BMAX>(lane, j);
      if (i < BMAX) ssq[(c * kDecWarps + warp) * BMAX + i] = ss[j];
    }
  }
  __syncthreads();
  // lane b of every warp sums sequence b's partials (fixed order) and forms rstd; the other lanes fetch it by shuffle
  float rs = 0.f;
  if (lane < BMAX) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < NC * kDecWarps; ++w) tot += ssq[w * BMAX + lane];
    rs = rsqrtf(tot / (float)K + eps);
  }
#pragma unroll
  for (int b = 0; b < BMAX; ++b) {
    const float rstd = __shfl_sync(0xffffffffu, rs, b);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = threadIdx.x + c * kDecThreads;
      if (col < C4 && b < B) {
        float4 v = *reinterpret_cast<float4*>(xs + b * K + 4 * col);
        v = make_float4(v.x * rstd * g[c].x, v.y * rstd * g[c].y, v.z * rstd * g[c].z, v.w * rstd * g[c].w);
        *reinterpret_cast<float4*>(xs + b * K + 4 * col) = v;
      }
    }
  }
}

constexpr int kRW = 4;   // weight rows per task of the "rows" phases

template <typename TW, int BMAX, bool BIG>
__global__ void __launch_bounds__(kDecThreads, 1) decode_token_kernel(const MambaDecodeTokenArgs a) {
  constexpr int K1 = BIG ? 1024 : 128;          // d_model
  constexpr int K2 = BIG ? 2048 : 256;          // d_inner
  constexpr int NQ1 = BIG ? 8 : 1;              // "rows" phases: chunks along K, tiles in the ring (half of K in flight)
  constexpr int NB1 = BIG ? 4 : 1;
  constexpr int KH1 = K1 / 128 / NQ1;           // 16-byte pieces per lane, row and chunk
  constexpr int NV1 = kRW * BMAX, NF1 = RsPlan<NV1>::NF;   // outputs per task / per lane after the reduce-scatter
  constexpr int SL4 = K2 / kDecWarps / 4;       // K-split phases: 16-byte pieces per warp slice (64 : 8)
  constexpr int KQ = SL4 < 16 ? SL4 : 16;       // active K-lanes
  constexpr int KV = SL4 / KQ;                  // 16-byte pieces per K-lane ...
  constexpr int KP = KV < 2 ? KV : 2;           // ... loaded KP at a time
  constexpr int NP = KV / KP;
  constexpr int RMAX4 = BIG ? 7 : 1;            // out_proj rows per CTA (x_proj: 2; both checked against gridDim by the host)
  constexpr int DJ = BIG ? 14 : 2;              // SSM phase: channels per CTA (d = blockIdx + j*gridDim; checked by the host)
  constexpr int NPAIR = DJ * BMAX;              // (channel, sequence) pairs per CTA
  constexpr int NRND = (NPAIR + 15) / 16;       // state rows per half-warp
  constexpr int XDP = 3 * 64 + 4;               // row stride of the x_proj outputs in shared memory (max XD, + 4: banks)
  static_assert(NPAIR <= kDecThreads, "one thread per (channel, sequence) pair in the dt step");
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                             // [BMAX][K1] normalised activations of the "rows" phases
  float* red = xs + BMAX * K1;                  // [warps][8 rows][BMAX] partial sums of the K-split phases
  float* xd_s = red + kDecWarps * 8 * BMAX;     // [BMAX][XDP] x_proj output of every sequence (SSM phase)
  float* A_s = xd_s + BMAX * XDP;               // [DJ][64] A rows of the CTA's channels
  float* pr_s = A_s + DJ * 64;                  // [4][NPAIR]: delta*log2e, delta*u, D*u, silu(z) per pair
  MambaDecodeLayer* Ls = reinterpret_cast<MambaDecodeLayer*>(pr_s + 4 * NPAIR);   // the layer descriptors
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kq = lane & 15, bq = lane >> 4;
  const int nblk = gridDim.x;
  const int gw = warp * nblk + blockIdx.x;      // warp w of every CTA before warp w+1 of any: tasks spread over the SMs
  const int nwt = nblk * kDecWarps;
  const int B = a.batch, N = a.d_state, R = a.dt_rank;
  const int XD = R + 2 * N;
  int nstamp = 0;
  auto stamp = [&](int id) {   // (id << 56 | ns) per event, CTA 0 thread 0
    if ((a.flags & MAMBA_DECODE_FLAG_STAMPS) && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      reinterpret_cast<unsigned long long*>(a.barrier + kStampWords)[nstamp++] = ((unsigned long long)id << 56) | (t & 0xffffffffffffffull);
    }
  };
  stamp(0);
  unsigned int target = 0;
  // MAMBA_DECODE_FLAG_BARRIER_STAMPS: every CTA records when it arrives at and when it leaves each grid barrier
  unsigned long long* const bst = reinterpret_cast<unsigned long long*>(a.barrier + kStampWords) + (16 * a.n_layers + 8) +
                                  (size_t)blockIdx.x * 2 * (4 * a.n_layers);
  int nbar = 0;
  auto bstamp = [&](int leave) {
    if ((a.flags & MAMBA_DECODE_FLAG_BARRIER_STAMPS) && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      bst[2 * nbar + leave] = t;
    }
    nbar += leave;
  };
  // scratch carve-up (fp32): S (residual stream) | z | xc | x_dbl | y
  float* const S = a.scratch;
  float* const zbuf = a.scratch + (size_t)kMaxB * K1;
  float* const xc = zbuf + (size_t)kMaxB * K2;
  float* const xdbl = xc + (size_t)kMaxB * K2;
  float* const ybuf = xdbl + (size_t)kMaxB * XD;
  for (int i = threadIdx.x; i < (BMAX - B) * K1; i += kDecThreads) xs[B * K1 + i] = 0.f;   // padded sequences
  for (int i = threadIdx.x; i < a.n_layers * (int)(sizeof(MambaDecodeLayer) / 8); i += kDecThreads)
    reinterpret_cast<unsigned long long*>(Ls)[i] = reinterpret_cast<const unsigned long long*>(a.layers)[i];
  __syncthreads();

  // ---- registers a warp loads ahead of the phase that uses them -------------------------------------------------------------
  WTile<kRW, KH1> t1[NB1];        // "rows" phases: the ring, holding the head of the warp's first task
  float4 cst[NF1], cw[NF1];       // conv state / taps of the (row, sequence) outputs this lane will hold
  float cb[NF1];
  int l = 0;
  auto conv_pre = [&](int task) {
    const MambaDecodeLayer& L = Ls[l];
#pragma unroll
    for (int j = 0; j < NF1; ++j) {
      const int i = rs_index<NV1>(lane, j);
      const int r = i / BMAX, b = i - r * BMAX, n = task * kRW + r;
      if (i < NV1 && n < K2 && b < B) {
        cst[j] = *reinterpret_cast<const float4*>(L.conv_state + ((size_t)b * K2 + n) * 4);
        cw[j] = __ldg(reinterpret_cast<const float4*>(L.conv_weight + (size_t)n * 4));
        cb[j] = L.conv_bias ? __ldg(L.conv_bias + n) : 0.f;
      }
    }
  };
  auto load_first = [&](const TW* W, int nrows) {
    if (gw * kRW < nrows) {
#pragma unroll
      for (int q = 0; q < NB1; ++q) load_chunk<TW, kRW, KH1>(t1[q], W, nrows, K1, gw, q, lane);
    }
  };
  // the warp's first task of a "rows" phase, pulled into L2 two phases ahead (kRW rows x K1 elements, 128 bytes per prefetch)
  auto prefetch_first = [&](const TW* W, int nrows) {
    if (gw * kRW < nrows) {
      const char* p0 = reinterpret_cast<const char*>(W + (size_t)(gw * kRW) * K1);
      const int bytes = min(kRW, nrows - gw * kRW) * K1 * (int)sizeof(TW);
      for (int off = lane * 128; off < bytes; off += 32 * 128) prefetch_l2(p0 + off);
    }
  };
  load_first(static_cast<const TW*>(Ls[0].in_proj_weight), 2 * K2);
  conv_pre(gw);

  for (; l < a.n_layers; ++l) {
    // ---- phase 1: norm + in_proj (+ conv step on the x half); "rows" mapping ----------------------------------------------
    stage_norm<TW, K1, BMAX>(xs, red, S, l == 0 ? static_cast<const TW*>(a.embedding) : nullptr, a.token, Ls[l].norm_weight, a.eps, B);
    __syncthreads();
    stamp(10);
    auto epi1 = [&](int j, int, int n, int b, float v) {
      if (n < K2) {  // conv branch: shift register + SiLU (simple_mamba.pyc @L233-237 for one position)
        const float4 s = cst[j], w = cw[j];
        *reinterpret_cast<float4*>(Ls[l].conv_state + ((size_t)b * K2 + n) * 4) = make_float4(s.y, s.z, s.w, v);
        const float acc = fmaf(w.w, v, fmaf(w.z, s.w, fmaf(w.y, s.z, fmaf(w.x, s.y, cb[j]))));
        xc[(size_t)b * K2 + n] = silu_f(acc);
      } else {
        zbuf[(size_t)b * K2 + (n - K2)] = v;
      }
    };
    rows_phase<TW, kRW, KH1, NQ1, NB1, BMAX>(t1, static_cast<const TW*>(Ls[l].in_proj_weight), static_cast<const TW*>(Ls[l].in_proj_bias),
                                             2 * K2, K1, B, xs, gw, nwt, lane, conv_pre, epi1, [&]() { stamp(11); });
    stamp(1);
    bstamp(0);
    grid_arrive(a.barrier);
    target += nblk;
    // ---- phase 2: x_proj; K-split mapping ------------------------------------------------------------------------------------
    prefetch_rows_l2<TW, KQ>(static_cast<const TW*>(Ls[l].x_proj_weight), XD, K2, 2, warp, lane);
    prefetch_rows_l2<TW, KQ>(static_cast<const TW*>(Ls[l].out_proj_weight), K1, K2, RMAX4, warp, lane);
    {  // what the SSM phase will load while ITS barrier opens: state rows, A and dt_proj rows of the CTA's channels
      const MambaDecodeLayer& L = Ls[l];
      const int nl = (N * 4 + 127) >> 7, rl = (R * 4 + 127) >> 7;   // 128-byte lines per row
      for (int i = threadIdx.x; i < NPAIR * nl; i += kDecThreads) {
        const int pi = i / nl, c = i - pi * nl, j = pi / BMAX, b = pi - j * BMAX, d = blockIdx.x + j * nblk;
        if (d < K2 && b < B) prefetch_l2(reinterpret_cast<const char*>(L.ssm_state + ((size_t)b * K2 + d) * N) + c * 128);
      }
      for (int i = threadIdx.x; i < DJ * (nl + rl); i += kDecThreads) {
        const int j = i / (nl + rl), c = i - j * (nl + rl), d = blockIdx.x + j * nblk;
        if (d < K2)
          prefetch_l2(c < nl ? reinterpret_cast<const char*>(L.A + (size_t)d * N) + c * 128
                             : reinterpret_cast<const char*>(L.dt_weight + (size_t)d * R) + (c - nl) * 128);
      }
      if (threadIdx.x < 2 * DJ) {   // dt_proj bias and D of those channels
        const int d = blockIdx.x + (threadIdx.x >> 1) * nblk;
        const float* v = (threadIdx.x & 1) ? L.D : L.dt_bias;
        if (d < K2 && v) prefetch_l2(v + d);
      }
    }
    grid_wait(a.barrier, target);
    bstamp(1);
    stamp(2);
    ksplit_rows<TW, 2, KQ, KP, NP, BMAX / 2, BMAX, 8>(static_cast<const TW*>(Ls[l].x_proj_weight), XD, K2, xc, B, red, warp, kq, bq);
    finish_rows<TW, BMAX, 8>(red, (const TW*)nullptr, XD, B, 2, [&](int n, int b, float v) { xdbl[(size_t)b * XD + n] = v; });
    stamp(3);
    bstamp(0);
    grid_arrive(a.barrier);
    target += nblk;
    // ---- phase 3: dt_proj + softplus, then SSM step + D skip + gate --------------------------------------------------------------
    // The CTA owns channels d = blockIdx + j*gridDim.  Step A: one thread per (channel, sequence) pair forms
    // delta = softplus(<dt_proj row, x_dbl[:R]> + bias) once (not once per lane of a reduction).  Step B: 16 lanes per
    // pair, 4 states each; a half-warp walks pairs hp, hp + 16, ... whose state rows it loaded before the barrier opened.
    {
      const MambaDecodeLayer& L = Ls[l];
      const int hl = lane & 15, hp = threadIdx.x >> 4;
      const bool nact = hl < (N >> 2);
      float4 h[NRND];
#pragma unroll
      for (int k = 0; k < NRND; ++k) {
        const int pi = hp + 16 * k, j = pi / BMAX, b = pi - j * BMAX, d = blockIdx.x + j * nblk;
        h[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pi < NPAIR && d < K2 && b < B && nact) h[k] = *(reinterpret_cast<const float4*>(L.ssm_state + ((size_t)b * K2 + d) * N) + hl);
      }
      // A rows of the CTA's channels: one 16-byte piece per thread, requested now, parked in shared memory after the barrier
      static_assert(DJ * 16 <= kDecThreads, "one A piece per thread");
      float4 a_pre = make_float4(0.f, 0.f, 0.f, 0.f);
      {
        const int j = threadIdx.x >> 4, d = blockIdx.x + j * nblk;
        if (j < DJ && d < K2 && nact) a_pre = __ldg(reinterpret_cast<const float4*>(L.A + (size_t)d * N) + hl);
      }
      {  // the next RMSNorm's weights (K1 floats)
        const float* nw = l + 1 < a.n_layers ? Ls[l + 1].norm_weight : a.norm_f_weight;
        if (threadIdx.x * 32 < K1) prefetch_l2(nw + threadIdx.x * 32);
      }
      if (l + 1 < a.n_layers) {
        prefetch_first(static_cast<const TW*>(Ls[l + 1].in_proj_weight), 2 * K2);
        const int n0 = gw * kRW;   // conv state of the task's rows: kRW x 16 bytes per sequence
        if (n0 < K2 && lane < B) prefetch_l2(Ls[l + 1].conv_state + ((size_t)lane * K2 + n0) * 4);
        if (n0 < K2 && lane == 31) prefetch_l2(Ls[l + 1].conv_weight + (size_t)n0 * 4);
      } else {
        prefetch_first(static_cast<const TW*>(a.head_weight), a.vocab);
      }
      grid_wait(a.barrier, target);
      bstamp(1);
      stamp(4);
      // x_proj's output of every sequence -> shared memory (one L2 round trip for the CTA)
      for (int i = threadIdx.x; i < B * (XD >> 2); i += kDecThreads) {
        const int b = i / (XD >> 2), c = i - b * (XD >> 2);
        *reinterpret_cast<float4*>(xd_s + b * XDP + 4 * c) = ldcg4(xdbl + (size_t)b * XD + 4 * c);
      }
      // step A operands that do not need the staged rows
      const int pj = threadIdx.x / BMAX, pb = threadIdx.x - pj * BMAX, pd = blockIdx.x + pj * nblk;
      const bool pon = threadIdx.x < NPAIR && pd < K2 && pb < B;
      float xcv = 0.f, zv = 0.f, dtb = 0.f, Dd = 0.f;
      if (pon) {
        xcv = ldcg(xc + (size_t)pb * K2 + pd), zv = ldcg(zbuf + (size_t)pb * K2 + pd);
        dtb = L.dt_bias ? __ldg(L.dt_bias + pd) : 0.f, Dd = L.D ? __ldg(L.D + pd) : 0.f;
      }
      if ((threadIdx.x >> 4) < DJ) *reinterpret_cast<float4*>(A_s + (threadIdx.x >> 4) * 64 + 4 * hl) = a_pre;
      __syncthreads();
      stamp(12);
      if (pon) {
        const float4* wr = reinterpret_cast<const float4*>(L.dt_weight + (size_t)pd * R);
        const float4* xr = reinterpret_cast<const float4*>(xd_s + pb * XDP);
        float4 wdt[16];   // the pair's dt_proj row (R <= 64), all loads in flight at once
#pragma unroll
        for (int r4 = 0; r4 < 16; ++r4) wdt[r4] = r4 < (R >> 2) ? __ldg(wr + r4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int r4 = 0; r4 < 16; ++r4) {
          if (r4 < (R >> 2)) {
            const float4 w = wdt[r4], x = xr[r4];
            d0 = fmaf(w.x, x.x, d0), d1 = fmaf(w.y, x.y, d1), d2 = fmaf(w.z, x.z, d2), d3 = fmaf(w.w, x.w, d3);
          }
        }
        const float delta = softplus_fast((d0 + d1) + (d2 + d3) + dtb);
        pr_s[threadIdx.x] = delta * kLog2e;
        pr_s[NPAIR + threadIdx.x] = delta * xcv;
        pr_s[2 * NPAIR + threadIdx.x] = Dd * xcv;
        pr_s[3 * NPAIR + threadIdx.x] = silu_fast(zv);
      }
      __syncthreads();
      stamp(13);
      const unsigned mask = 0xffffu << (lane & 16);
#pragma unroll
      for (int k = 0; k < NRND; ++k) {
        const int pi = hp + 16 * k, j = pi / BMAX, b = pi - j * BMAX, d = blockIdx.x + j * nblk;
        const bool on = pi < NPAIR && d < K2 && b < B;     // (uniform over the half-warp)
        const int pc = on ? pi : 0, bc = on ? b : 0, jc = on ? j : 0;
        const float dl2 = pr_s[pc], du = pr_s[NPAIR + pc];
        const float4 A4 = *reinterpret_cast<const float4*>(A_s + jc * 64 + 4 * hl);
        const float4 bb = *reinterpret_cast<const float4*>(xd_s + bc * XDP + R + (nact ? 4 * hl : 0));
        const float4 cc = *reinterpret_cast<const float4*>(xd_s + bc * XDP + R + N + (nact ? 4 * hl : 0));
        float4 h4 = h[k];
        h4.x = fmaf(ex2_approx(dl2 * A4.x), h4.x, du * bb.x);
        h4.y = fmaf(ex2_approx(dl2 * A4.y), h4.y, du * bb.y);
        h4.z = fmaf(ex2_approx(dl2 * A4.z), h4.z, du * bb.z);
        h4.w = fmaf(ex2_approx(dl2 * A4.w), h4.w, du * bb.w);
        if (on && nact) *(reinterpret_cast<float4*>(L.ssm_state + ((size_t)b * K2 + d) * N) + hl) = h4;
        float y = nact ? fmaf(h4.x, cc.x, fmaf(h4.y, cc.y, fmaf(h4.z, cc.z, h4.w * cc.w))) : 0.f;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) y += __shfl_xor_sync(mask, y, o);
        if (on && hl == 0) ybuf[(size_t)b * K2 + d] = (y + pr_s[2 * NPAIR + pc]) * pr_s[3 * NPAIR + pc];
      }
    }
    stamp(5);
    bstamp(0);
    grid_arrive(a.barrier);
    target += nblk;
    // ---- phase 4: out_proj, added into the residual stream; K-split mapping -------------------------------------------------------
    grid_wait(a.barrier, target);
    bstamp(1);
    stamp(6);
    ksplit_rows<TW, RMAX4, KQ, KP, NP, BMAX / 2, BMAX, 8>(static_cast<const TW*>(Ls[l].out_proj_weight), K1, K2, ybuf, B, red, warp, kq, bq);
    stamp(14);
    finish_rows<TW, BMAX, 8>(red, static_cast<const TW*>(Ls[l].out_proj_bias), K1, B, RMAX4, [&](int n, int b, float v) {
      float* sp = S + (size_t)b * K1 + n;   // hidden + residual: this thread is the only one that touches the element
      *sp = v + ldcg(sp);
    });
    stamp(7);
    bstamp(0);
    grid_arrive(a.barrier);
    target += nblk;
    if (l + 1 < a.n_layers) {
      ++l;   // conv_pre reads the NEXT layer's descriptor
      load_first(static_cast<const TW*>(Ls[l].in_proj_weight), 2 * K2);
      conv_pre(gw);
      --l;
    } else {
      load_first(static_cast<const TW*>(a.head_weight), a.vocab);
    }
    grid_wait(a.barrier, target);
    bstamp(1);
    stamp(8);
  }
  // ---- final norm + LM head; "rows" mapping ------------------------------------------------------------------------------------
  stage_norm<TW, K1, BMAX>(xs, red, S, (const TW*)nullptr, a.token, a.norm_f_weight, a.eps, B);
  __syncthreads();
  rows_phase<TW, kRW, KH1, NQ1, NB1, BMAX>(t1, static_cast<const TW*>(a.head_weight), static_cast<const TW*>(a.head_bias), a.vocab, K1, B,
                                           xs, gw, nwt, lane, [](int) {}, [&](int, int, int n, int b, float v) { a.logits[(size_t)b * a.logits_bs + n] = v; },
                                           []() {});
  stamp(9);
  // the last CTA to finish zeroes the barrier words for the next launch (no CTA reads them after the final barrier)
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(a.barrier + 1, 1u);
    if (done == (unsigned int)nblk - 1u) {
      a.barrier[0] = 0u;
      a.barrier[1] = 0u;
    }
  }
}

template <typename TW, int BMAX, bool BIG>
int launch_decode(const MambaDecodeTokenArgs& a, cudaLaunchConfig_t& cfg) {
  static thread_local SmemConfig sc;
  auto kern = decode_token_kernel<TW, BMAX, BIG>;
  if (int rc = ensure_dynamic_smem(kern, cfg.dynamicSmemBytes, sc, "decode_token")) return rc;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return set_error(MAMBA_ELAUNCH, "decode_token: %s", cudaGetErrorString(e));
  }
  return MAMBA_OK;
}
template <typename TW, bool BIG>
int launch_decode_b(const MambaDecodeTokenArgs& a, cudaLaunchConfig_t& cfg, int bmax) {
#ifdef MB_DEC_DEV   // development builds: one batch size per shape
  (void)bmax;
  return launch_decode<TW, BIG ? 10 : 6, BIG>(a, cfg);
#else
  switch (bmax) {
    case 2: return launch_decode<TW, 2, BIG>(a, cfg);
    case 4: return launch_decode<TW, 4, BIG>(a, cfg);
    case 6: return launch_decode<TW, 6, BIG>(a, cfg);
    case 8: return launch_decode<TW, 8, BIG>(a, cfg);
    case 10: return launch_decode<TW, 10, BIG>(a, cfg);
    case 12: return launch_decode<TW, 12, BIG>(a, cfg);
    case 14: return launch_decode<TW, 14, BIG>(a, cfg);
    default: return launch_decode<TW, 16, BIG>(a, cfg);
  }
#endif
}

}  // namespace

}  // namespace mb

extern "C" size_t mamba_decode_token_scratch_bytes(int d_model, int d_inner, int d_state, int dt_rank) {
  if (d_model <= 0 || d_inner <= 0 || d_state <= 0 || dt_rank <= 0) return 0;
  return sizeof(float) * (size_t)mb::kMaxB * (3 * (size_t)d_model + 3 * (size_t)d_inner + dt_rank + 2 * d_state) + 256;
}

extern "C" size_t mamba_decode_token_barrier_bytes(int n_layers) {
  if (n_layers <= 0) return 0;
  return sizeof(unsigned int) * mb::kStampWords + sizeof(unsigned long long) * ((size_t)(16 * n_layers + 8) + (size_t)256 * 2 * 4 * n_layers);
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
  // d_state and dt_rank multiples of 4 up to 64, d_conv 4
  const bool big = a->d_model == 1024 && a->d_inner == 2048, small = a->d_model == 128 && a->d_inner == 256;
  if (!(big || small) || a->d_state % 4 || a->dt_rank % 4 || a->d_state > 64 || a->dt_rank > 64 || a->d_conv != 4)
    return set_error(MAMBA_ESIZE, "decode_token: unsupported shape d_model %d d_inner %d d_state %d dt_rank %d d_conv %d", a->d_model,
                     a->d_inner, a->d_state, a->dt_rank, a->d_conv);
  if (a->scratch_bytes < mamba_decode_token_scratch_bytes(a->d_model, a->d_inner, a->d_state, a->dt_rank))
    return set_error(MAMBA_ESIZE, "decode_token: scratch too small");
  if (a->w_dtype != MAMBA_F32 && a->w_dtype != MAMBA_BF16) return set_error(MAMBA_EDTYPE, "decode_token: w_dtype %d", a->w_dtype);
  int dev = 0, nsm = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  // rows per CTA of the K-split phases are compile-time: x_proj <= 2, out_proj <= 7 (1)
  if (a->dt_rank + 2 * a->d_state > 2 * nsm || a->d_model > (big ? 7 : 1) * nsm || a->d_inner > (big ? 14 : 2) * nsm)
    return set_error(MAMBA_ESIZE, "decode_token: %d SMs are too few for this shape", nsm);
  const int bmax = (a->batch + 1) & ~1;
  const int dj = big ? 14 : 2;
  const size_t smem = sizeof(float) * ((size_t)bmax * a->d_model + (size_t)kDecWarps * 8 * bmax + (size_t)bmax * 196 + (size_t)dj * 64 +
                                       (size_t)4 * dj * bmax) +
                      sizeof(MambaDecodeLayer) * (size_t)a->n_layers;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nsm), cfg.blockDim = dim3(kDecThreads), cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // every CTA resident at once: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  int rc;
  if (a->w_dtype == MAMBA_F32)
    rc = big ? launch_decode_b<float, true>(*a, cfg, bmax) : launch_decode_b<float, false>(*a, cfg, bmax);
  else
    rc = big ? launch_decode_b<__nv_bfloat16, true>(*a, cfg, bmax) : launch_decode_b<__nv_bfloat16, false>(*a, cfg, bmax);
  if (rc != MAMBA_OK) return rc;
  count_launch();
  return check_launch("decode_token");
}
