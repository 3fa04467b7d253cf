// sample.cu — next-token choice of the reference's two generation loops as ONE kernel per decode step, sm_100a.
//
// Replaces, for the last position of every sequence (one CTA per sequence):
//   * train.filtered_logit (train.py:133-138): -log_softmax(logits, dim = SEQUENCE) * weights[bucket(prev token)],
//     with the sequence-axis logsumexp carried as a running per-(b, v) value (lse' = logaddexp(lse, logit));
//   * the repetition penalties over the look-back window:
//       mode 0, scripts/generate_midi_many.py:20-46 — window = the last 100 tokens, argmax;
//       mode 1, scripts/generate.py:33-85 — window = the tokens after the one at which the summed time shifts,
//       walked backwards, reach 64*16; k in {1,2,3} drawn by the class of the last token, top-k, one draw from
//       the k values normalised by their sum;
//   * the bookkeeping the host loops do with python lists: the chosen token is appended to `generated`, the window
//     is advanced and `counts` (occurrences inside the window) is kept in step.
// Randomness enters as two uniforms per (step, sequence) supplied by the caller (`uniforms[step][b][0..1]`, e.g. a
// Philox stream of torch.rand): k = choices[floor(u0 * len)], the draw is the inverse CDF at u1.  Parity is defined as
// identical tokens for identical uniforms (oracle/train_ref.py restates the loop in that form).
#include "common.cuh"

namespace mb {

constexpr int kSampleThreads = 1024;

struct Cand {
  float v;
  int i;
};
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) { return a.v > b.v || (a.v == b.v && a.i < b.i); }

__device__ __forceinline__ int token_class(int64_t tok, const int32_t* bnd) {  // 0 pitch, 1 dyn, 2 length, 3 time, 4 tempo
  return (tok >= bnd[0]) + (tok >= bnd[1]) + (tok >= bnd[2]) + (tok >= bnd[3]);
}

__global__ void __launch_bounds__(kSampleThreads) sample_step_kernel(const MambaSampleStepArgs a) {
  __shared__ Cand warp_best[kSampleThreads / 32];
  __shared__ Cand top[3];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = a.vocab;
  pdl_launch_dependents();  // decode chain (common.cuh): the next kernel (embedding / in_proj) may start its prologue
  pdl_wait();               // logits come from the head kernel
  const float* lg = a.logits + (int64_t)b * a.logits_bs;
  float* lse = a.lse + (int64_t)b * V;
  int32_t* counts = a.counts + (int64_t)b * V;
  int64_t* gen = a.generated + (int64_t)b * a.generated_bs;
  const int n = a.gen_len[b];
  const int64_t prev = gen[n - 1];
  // weights row: torch.bucketize(prev, [dyn-1, length-1, time-1, tempo-1], right=False) (train.py:114-131)
  const int bucket = (prev > a.bucket_bounds[0]) + (prev > a.bucket_bounds[1]) + (prev > a.bucket_bounds[2]) + (prev > a.bucket_bounds[3]);
  const float* w = a.dist + (int64_t)bucket * V;

  Cand c0{-INFINITY, 0x7fffffff}, c1 = c0, c2 = c0;  // this thread's three best, in order
  // kU vocabulary entries per thread and round: their 4 x kU loads are all issued before the first use (the rows are
  // cold — the decode kernel has just streamed 450 MB through L2 — so a round is one DRAM round trip, not kU of them)
  constexpr int kU = 6;
#pragma unroll 1
  for (int v0 = tid; v0 < V; v0 += kSampleThreads * kU) {
    float xs[kU], ls[kU], ws[kU];
    int cs[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int v = v0 + u * kSampleThreads;
      const bool ok = v < V;
      xs[u] = ok ? lg[v] : 0.f, ls[u] = ok ? lse[v] : 0.f, ws[u] = ok ? w[v] : 0.f, cs[u] = ok ? counts[v] : 0;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int v = v0 + u * kSampleThreads;
      if (v >= V) break;
      const float x = xs[u], l0 = ls[u];
      // logaddexp as torch computes it: max + log1p(exp(-|d|))
      const float m = fmaxf(x, l0);
      const float l1 = (x == l0 && isinf(x)) ? x : m + log1pf(expf(-fabsf(x - l0)));
      lse[v] = l1;
      float f = -(x - l1) * ws[u];
      const int c = cs[u];
      if (c > 0) {
        const int cls = (v >= a.class_bounds[0]) + (v >= a.class_bounds[1]) + (v >= a.class_bounds[2]) + (v >= a.class_bounds[3]);
        const int rule = a.pen_rule[cls];
        float pen = 1.f;
        if (rule == 1) pen = a.pen_table[cls * 128 + min(c, 127)];                           // python: min(base ** count, cap)
        else if (rule == 2) pen = c >= 10 ? (float)(1.1 * (double)c) : 1.f;                  // generate_midi_many.py:33-35
        f = f / pen;
      }
      const Cand x3{f, v};
      if (better(x3, c0)) c2 = c1, c1 = c0, c0 = x3;
      else if (better(x3, c1)) c2 = c1, c1 = x3;
      else if (better(x3, c2)) c2 = x3;
    }
  }
  // three rounds of block-wide argmax; the winner's thread moves its next candidate up
  const int rounds = a.mode == 0 ? 1 : 3;
  for (int r = 0; r < rounds; ++r) {
    Cand m = c0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Cand t{__shfl_xor_sync(0xffffffffu, m.v, o), __shfl_xor_sync(0xffffffffu, m.i, o)};
      if (better(t, m)) m = t;
    }
    if (lane == 0) warp_best[warp] = m;
    __syncthreads();
    if (warp == 0) {
      Cand t = warp_best[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Cand u{__shfl_xor_sync(0xffffffffu, t.v, o), __shfl_xor_sync(0xffffffffu, t.i, o)};
        if (better(u, t)) t = u;
      }
      if (lane == 0) top[r] = t;
    }
    __syncthreads();
    if (c0.i == top[r].i) c0 = c1, c1 = c2, c2 = Cand{-INFINITY, 0x7fffffff};
  }
  if (tid != 0) return;

  // ---- the choice ------------------------------------------------------------------------------------------------
  int64_t tok = top[0].i;
  if (a.mode == 1) {
    const int step = n - a.prompt_len;
    const float u0 = a.uniforms[((int64_t)step * gridDim.x + b) * 2], u1 = a.uniforms[((int64_t)step * gridDim.x + b) * 2 + 1];
    // scripts/generate.py:48-58: k = random.choice(...) by the class of the last token
    const int cls = token_class(prev, a.class_bounds);
    int k = 1;
    if (cls == 4) k = (min((int)(u0 * 5.f), 4) >= 3) ? 2 : 1;       // [1,1,1,2,2]
    else if (cls == 1) k = (min((int)(u0 * 2.f), 1) == 1) ? 3 : 1;  // [1,3]
    else if (cls == 0) k = (min((int)(u0 * 2.f), 1) == 1) ? 2 : 1;  // [1,2]
    // :75-79 top-k values normalised by their sum, one draw (inverse CDF at u1)
    float sum = 0.f;
    for (int j = 0; j < k; ++j) sum += top[j].v;
    float acc = 0.f;
    int pick = k - 1;
    for (int j = 0; j < k; ++j) {
      acc += top[j].v / sum;
      if (u1 < acc) {
        pick = j;
        break;
      }
    }
    tok = top[pick].i;
  }
  a.next_token[b] = tok;

  // ---- append, advance the window, keep counts in step ---------------------------------------------------------------
  gen[n] = tok;
  a.gen_len[b] = n + 1;
  counts[tok] += 1;
  if (a.mode == 0) {
    if (n + 1 > 100) counts[gen[n - 100]] -= 1;  // window = generated[-100:]
  } else {
    // window = (q, n]: q is the last position whose suffix of time shifts sums to >= 64*16 (position 0 if none does)
    const int t0 = a.class_bounds[2], t1 = a.class_bounds[3];
    auto tv = [&](int64_t t) { return (t >= t0 && t < t1) ? (int)(t - t0) : 0; };
    int q = a.win_q[b];
    int s = a.win_sum[b] + tv(tok);
    while (q < n && s - tv(gen[q]) >= a.time_budget) {
      s -= tv(gen[q]);
      ++q;
      counts[gen[q]] -= 1;  // position q was inside the window and is now its (excluded) left edge
    }
    a.win_q[b] = q, a.win_sum[b] = s;
  }
}

}  // namespace mb

extern "C" int mamba_sample_step(const MambaSampleStepArgs* a, void* stream) {
  using namespace mb;
  if (!a || a->struct_size != (int32_t)sizeof(MambaSampleStepArgs))
    return set_error(MAMBA_EINVAL, "sample_step: bad args pointer or struct_size");
  if (a->batch <= 0 || a->vocab <= 0) return set_error(MAMBA_EINVAL, "sample_step: batch/vocab must be positive");
  if (a->mode != 0 && a->mode != 1) return set_error(MAMBA_EINVAL, "sample_step: mode must be 0 (greedy) or 1 (top-k draw)");
  if (!a->logits || !a->lse || !a->dist || !a->counts || !a->generated || !a->gen_len || !a->next_token || !a->pen_table)
    return set_error(MAMBA_EINVAL, "sample_step: null logits/lse/dist/counts/generated/gen_len/next_token/pen_table");
  if (a->mode == 1 && (!a->uniforms || !a->win_q || !a->win_sum))
    return set_error(MAMBA_EINVAL, "sample_step: mode 1 needs uniforms, win_q and win_sum");
  launch_chain(sample_step_kernel, dim3(a->batch), dim3(kSampleThreads), 0, static_cast<cudaStream_t>(stream), *a);
  count_launch();
  return check_launch("sample_step");
}
