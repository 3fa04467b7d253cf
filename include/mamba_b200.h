/*
 * mamba_b200.h — C-ABI of the B200-native (sm_100a) Mamba hot path.
 *
 * The reference (thorGabe123/Deep-Learning-Based-Sequence-Models-for-Music-Generation) has
 * no FFI of its own: its hot path is Python (`nn.Module` duck typing).  Every entry point
 * below therefore cites the *Python* statement(s) of the reference it replaces.  Paths are
 * relative to the reference root; `simple_mamba.pyc @Lnnn` means original source line nnn of
 * models/mamba/__pycache__/simple_mamba.cpython-311.pyc (the pure-PyTorch Mamba-1 whose
 * source file was deleted upstream; transcription in SURVEY.md Appendix A).
 *
 * Conventions
 *  - plain pointers and sizes only; no torch types.  All pointers are DEVICE pointers unless
 *    the name ends in `_host`.
 *  - the caller owns and allocates every input, output and workspace buffer; the library
 *    never allocates, never synchronises, and enqueues all work on the `stream` argument (a
 *    `cudaStream_t` passed as `void*`) of the CURRENT device.  Process-wide state is limited to
 *    the monotonic launch counter behind mamba_launch_count(); per-thread state to the last
 *    error message and a per-device record of which kernels have had their dynamic
 *    shared-memory limit raised (so one host thread may drive several devices in turn).
 *  - activations are row-major `[batch, seqlen, channels]` with the channel axis contiguous;
 *    `*_bs` / `*_ls` are the batch and sequence strides IN ELEMENTS, so the split views the
 *    reference takes of `in_proj(x)` and `x_proj(x)` are consumed without a copy.
 *  - `dtype` selects the activation I/O element type (MAMBA_F32 or MAMBA_BF16).  Parameters
 *    A, D, delta_bias, their gradients, the SSM state and all accumulation are always fp32.
 *  - every function returns MAMBA_OK (0) or a negative MAMBA_E* code; `mamba_last_error()`
 *    returns a thread-local human-readable message for the last failure on this thread.
 */
#ifndef MAMBA_B200_H_
#define MAMBA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAMBA_ABI_VERSION 3

enum { MAMBA_F32 = 0, MAMBA_BF16 = 1 };

enum {
  MAMBA_OK = 0,
  MAMBA_EINVAL = -1,   /* bad shape / null pointer / bad struct_size            */
  MAMBA_EDTYPE = -2,   /* unsupported dtype                                     */
  MAMBA_EALIGN = -3,   /* pointer or stride breaks the documented alignment     */
  MAMBA_ELAUNCH = -4,  /* CUDA launch failure (cudaPeekAtLastError != success)  */
  MAMBA_ESIZE = -5     /* workspace too small / dimension above supported limit */
};

enum {
  MAMBA_FLAG_HAS_Z = 1,          /* out = y * silu(z)             (simple_mamba.pyc @L241)  */
  MAMBA_FLAG_DELTA_SOFTPLUS = 2, /* delta = softplus(delta_raw)   (simple_mamba.pyc @L276)  */
  MAMBA_FLAG_HAS_DELTA_BIAS = 4, /* delta_raw += delta_bias[d]    (dt_proj bias, @L276)     */
  MAMBA_FLAG_HAS_D = 8,          /* y += u * D                    (simple_mamba.pyc @L331)  */
  MAMBA_FLAG_A_IS_LOG = 16       /* `A` holds A_log: the kernels use A = -exp(A_log) (@L270) and the backward
                                    returns dA w.r.t. A_log (= dA * A)                                      */
};

/* ------------------------------------------------------------------------------------------
 * Selective scan (forward).  Replaces MambaBlock.selective_scan (simple_mamba.pyc @L310-333)
 * fused with softplus (@L276), the D skip (@L331) and the z gate (@L241):
 *     delta = softplus(delta_raw + delta_bias)                       (threshold 20, as F.softplus)
 *     h_t   = exp(delta_t * A) * h_{t-1} + delta_t * B_t * u_t        (h_{-1} = 0, per (b, d, n))
 *     y_t   = <h_t, C_t> + D * u_t
 *     out_t = y_t * silu(z_t)
 * The [B, L, D, N] tensors deltaA / deltaB_u of the reference are never materialised.
 * When `ckpt` is non-null the state at the start of every `chunk`-timestep block is written to
 * it (fp32, layout [batch, nchunks, ceil(dstate/4), dim, 4], nchunks = ceil(seqlen/chunk); slot 0 is
 * not written; size from mamba_scan_ckpt_elems) for use by mamba_scan_bwd; with a z gate the backward also needs `y_pre`.  When `h_last` is non-null the final state h_{L-1} is
 * written to it ([batch, dim, dstate] fp32 — the decode step's `ssm_state` layout).
 * ------------------------------------------------------------------------------------------ */
typedef struct MambaScanFwdArgs {
  int32_t struct_size; /* = sizeof(MambaScanFwdArgs) */
  int32_t dtype;
  int32_t batch, seqlen, dim, dstate;
  int32_t chunk; /* checkpoint interval in timesteps: 8 or 16 (ignored if ckpt == NULL)     */
  int32_t flags;
  int32_t variant; /* 0 = auto; 4, 8, 16 = LDGSTS-staged kernel with that many states per thread; 100 + 10*shape + split =
                      TMA-staged kernel (shape 0 auto / 1 / 2 = tiling, split 0 / 1 = share of exp2 on the FMA pipe);
                      tuning knob, results are the same within rounding */
  int32_t reserved;
  const void* u;     int64_t u_bs, u_ls;         /* [B, L, D]                 */
  const void* delta; int64_t delta_bs, delta_ls; /* [B, L, D] raw dt_proj out */
  const float* A;                                /* [D, N]  (= -exp(A_log))   */
  const void* B;     int64_t B_bs, B_ls;         /* [B, L, N]                 */
  const void* C;     int64_t C_bs, C_ls;         /* [B, L, N]                 */
  const float* D;                                /* [D] or NULL               */
  const void* z;     int64_t z_bs, z_ls;         /* [B, L, D] or NULL         */
  const float* delta_bias;                       /* [D] or NULL               */
  void* out;         int64_t out_bs, out_ls;     /* [B, L, D]                 */
  float* ckpt;                                   /* see above, or NULL        */
  float* h_last;                                 /* [B, D, N] or NULL         */
  const float* h_init;                           /* [B, D, N] or NULL (zeros) */
  void* y_pre;       int64_t y_pre_bs, y_pre_ls; /* [B, L, D] or NULL: the output BEFORE the z gate
                                                    (scan + D*u), saved for mamba_scan_bwd's dz */
} MambaScanFwdArgs;

int mamba_scan_fwd(const MambaScanFwdArgs* args, void* stream);

/* Number of floats in the `ckpt` buffer for the given problem. */
size_t mamba_scan_ckpt_elems(int batch, int seqlen, int dim, int dstate, int chunk);

/* ------------------------------------------------------------------------------------------
 * Selective scan (backward).  Replaces what torch autograd derives from the Python loop of
 * simple_mamba.pyc @L310-333 (+ softplus/D/z as above).  The forward is recomputed inside each
 * `chunk` from the checkpoints written by mamba_scan_fwd; nothing of size B*L*D*N touches HBM.
 * Gradients du/ddelta/dz/dB/dC have the activation dtype; ddelta is w.r.t. delta_raw (softplus
 * derivative applied).  dA [D, N], dD [D], ddelta_bias [D] are fp32 and are OVERWRITTEN.
 * `workspace` must hold mamba_scan_bwd_workspace_bytes(...) bytes (any contents).
 * ------------------------------------------------------------------------------------------ */
typedef struct MambaScanBwdArgs {
  int32_t struct_size;
  int32_t dtype;
  int32_t batch, seqlen, dim, dstate;
  int32_t chunk;
  int32_t flags;
  int32_t variant;   /* 0 = auto; 1 = fused recompute/reverse kernel with tensor-pipe channel sums (d_state 32 or 64;
                        auto for bf16 I/O); 2 / 4 / 8 / 16 = lane<->channel kernel with that many states per thread
                        (2: d_state <= 16); 104 = 4 states per thread with two helper-warp teams (d_state 9..16) */
  int32_t reserved;
  const void* u;     int64_t u_bs, u_ls;
  const void* delta; int64_t delta_bs, delta_ls;
  const float* A;
  const void* B;     int64_t B_bs, B_ls;
  const void* C;     int64_t C_bs, C_ls;
  const float* D;
  const void* z;     int64_t z_bs, z_ls;
  const float* delta_bias;
  const void* dout;  int64_t dout_bs, dout_ls;
  const float* ckpt;
  void* du;          int64_t du_bs, du_ls;
  void* ddelta;      int64_t ddelta_bs, ddelta_ls;
  void* dz;          int64_t dz_bs, dz_ls;         /* NULL iff !HAS_Z */
  void* dB;          int64_t dB_bs, dB_ls;
  void* dC;          int64_t dC_bs, dC_ls;
  float* dA;
  float* dD;          /* NULL iff !HAS_D          */
  float* ddelta_bias; /* NULL iff !HAS_DELTA_BIAS */
  void* workspace;
  size_t workspace_bytes;
  const void* y_pre; int64_t y_pre_bs, y_pre_ls; /* required iff HAS_Z: y_pre written by mamba_scan_fwd */
} MambaScanBwdArgs;

int mamba_scan_bwd(const MambaScanBwdArgs* args, void* stream);
size_t mamba_scan_bwd_workspace_bytes(int batch, int seqlen, int dim, int dstate);

/* ------------------------------------------------------------------------------------------
 * Causal depthwise conv1d + SiLU.  Replaces
 *     x = rearrange(x,'b l d -> b d l'); x = self.conv1d(x)[:, :, :l]; rearrange back; F.silu(x)
 * (simple_mamba.pyc @L233-237; nn.Conv1d(groups=d_inner, kernel_size=d_conv, padding=d_conv-1)
 * built at @L193-199):   out[b,t,d] = silu(bias[d] + sum_k w[d,k] * x[b, t-(K-1)+k, d]).
 * weight is [D, K] contiguous fp32 (the reference's [D,1,K] Conv1d weight viewed 2-D), bias [D]
 * fp32 or NULL, 2 <= K <= 4.  No transposes: x stays [B, L, D].
 * `final_state` (optional, [B, D, K] activation dtype) receives the last K inputs of each
 * channel (zero-padded on the left) — the decode step's `conv_state`.
 * ------------------------------------------------------------------------------------------ */
typedef struct MambaConvArgs {
  int32_t struct_size;
  int32_t dtype;
  int32_t batch, seqlen, dim, width;
  const void* x;   int64_t x_bs, x_ls;
  const float* weight;
  const float* bias;
  void* out;       int64_t out_bs, out_ls;
  void* final_state;
  /* backward only */
  const void* dout; int64_t dout_bs, dout_ls;
  void* dx;         int64_t dx_bs, dx_ls;
  float* dweight;  /* [D, K] fp32, overwritten */
  float* dbias;    /* [D]    fp32, overwritten (NULL iff bias == NULL) */
  void* workspace;
  size_t workspace_bytes;
} MambaConvArgs;

int mamba_conv1d_silu_fwd(const MambaConvArgs* args, void* stream);
int mamba_conv1d_silu_bwd(const MambaConvArgs* args, void* stream);
size_t mamba_conv1d_bwd_workspace_bytes(int batch, int seqlen, int dim, int width);

/* ------------------------------------------------------------------------------------------
 * Single-token recurrent step (decode).  The reference has no such step (its generate() re-runs
 * the whole model on a sliding window, scripts/generate.py:26-31); this is the recurrence of
 * simple_mamba.pyc @L233-241 and @L310-333 evaluated for one new position with carried state.
 *
 * mamba_conv_step:  conv_state[b,d,:] is shifted left by one, x_t appended, and
 *     xc[b,d] = silu(bias[d] + sum_k w[d,k] * conv_state[b,d,k])
 * mamba_ssm_step:   delta = softplus(dt_w[d,:] . dt_in[b,:] + dt_b[d])     (dt_proj fused, @L276)
 *     h[b,d,:] = exp(delta*A[d,:]) * h[b,d,:] + delta * Bv[b,:] * xc[b,d]
 *     y[b,d]   = (<h[b,d,:], Cv[b,:]> + D[d]*xc[b,d]) * silu(z[b,d])
 * conv_state is [B, D, K] and ssm_state [B, D, N] fp32, both updated in place.
 * ------------------------------------------------------------------------------------------ */
typedef struct MambaStepArgs {
  int32_t struct_size;
  int32_t dtype;
  int32_t batch, dim, dstate, width, dt_rank;
  int32_t flags;
  /* conv step */
  const void* x;  int64_t x_bs;           /* [B, D] new input column (in_proj x half) */
  void* conv_state;                       /* [B, D, K] activation dtype               */
  const float* conv_weight;               /* [D, K] */
  const float* conv_bias;                 /* [D] or NULL */
  void* xc;       int64_t xc_bs;          /* [B, D] out of conv step / in of ssm step */
  /* ssm step */
  const void* dt_in; int64_t dt_in_bs;    /* [B, R] low-rank delta (x_proj split)     */
  const void* Bv;    int64_t Bv_bs;       /* [B, N] */
  const void* Cv;    int64_t Cv_bs;       /* [B, N] */
  const float* dt_weight;                 /* [D, R] fp32 */
  const float* dt_bias;                   /* [D]    fp32 or NULL */
  const float* A;                         /* [D, N] */
  const float* D;                         /* [D] or NULL */
  const void* z;     int64_t z_bs;        /* [B, D] or NULL */
  float* ssm_state;                       /* [B, D, N] fp32 */
  void* y;           int64_t y_bs;        /* [B, D] */
} MambaStepArgs;

int mamba_conv_step(const MambaStepArgs* args, void* stream);
int mamba_ssm_step(const MambaStepArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * One-position linear layer (decode).  Replaces the nn.Linear calls of the block for ONE new position per
 * sequence: in_proj (simple_mamba.pyc @L230), x_proj (@L273), out_proj (@L243), lm_head (@L94):
 *     y[b, :] = weight @ x[b, :] (+ bias),   weight [out_features, in_features] row-major (nn.Linear.weight)
 * batch <= 16, in_features % 4 == 0.  `dtype` is the element type of x and y, `w_dtype` of weight and bias.
 * ------------------------------------------------------------------------------------------ */
typedef struct MambaLinearStepArgs {
  int32_t struct_size;
  int32_t dtype, w_dtype;
  int32_t batch, in_features, out_features;
  const void* x;      int64_t x_bs;   /* [B, in_features]  */
  const void* weight;                 /* [out_features, in_features] contiguous */
  const void* bias;                   /* [out_features] or NULL */
  void* y;            int64_t y_bs;   /* [B, out_features] */
} MambaLinearStepArgs;

int mamba_linear_step(const MambaLinearStepArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * mamba_linear_step with the neighbouring per-token work of the residual block folded in (decode only):
 *   prologue (norm_weight != NULL):  s = x + residual_in  (x NULL = zeros; fp32 residual stream);
 *                                    residual_out = s;  x' = s * rsqrt(mean(s^2) + eps) * norm_weight
 *                                    — `normed, resid = norm(hidden, resid)` of ResidualBlock.forward
 *                                    (simple_mamba.pyc @L179 with RMSNorm @L346) in front of in_proj @L230 / lm_head @L94;
 *   epilogue (conv_dim > 0):         output columns [0, conv_dim) are the conv branch of in_proj: each value is pushed
 *                                    into conv_state[b, n, :] (shift register of `conv_width` taps) and
 *                                    conv_out[b, n] = silu(conv_bias[n] + sum_k conv_weight[n, k] * conv_state[b, n, k])
 *                                    (@L233-237 for one new position: mamba_conv_step); the remaining columns
 *                                    (the gate z) go to y as usual.
 * residual_out must not alias residual_in.
 * ------------------------------------------------------------------------------------------ */
typedef struct MambaFusedLinearStepArgs {
  int32_t struct_size;
  int32_t dtype, w_dtype;
  int32_t batch, in_features, out_features;
  const void* x;      int64_t x_bs;   /* [B, in_features] or NULL (with a fused norm)  */
  const void* weight;                 /* [out_features, in_features] contiguous */
  const void* bias;                   /* [out_features] or NULL */
  void* y;            int64_t y_bs;   /* [B, out_features] */
  const float* norm_weight;           /* [in_features] or NULL (no prologue) */
  float eps;
  int32_t conv_dim;                   /* 0 = no epilogue */
  const float* residual_in;  int64_t residual_in_bs;   /* [B, in_features] fp32 or NULL */
  float* residual_out;       int64_t residual_out_bs;  /* [B, in_features] fp32 or NULL */
  int32_t conv_width, reserved;
  void* conv_state;                   /* [B, conv_dim, conv_width], activation dtype, updated in place */
  const float* conv_weight;           /* [conv_dim, conv_width] */
  const float* conv_bias;             /* [conv_dim] or NULL */
  void* conv_out;     int64_t conv_out_bs;  /* [B, conv_dim] */
} MambaFusedLinearStepArgs;

int mamba_fused_linear_step(const MambaFusedLinearStepArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Next-token choice for one decode step, one CTA per sequence (csrc/sample.cu).  Replaces the per-token host code of
 *   mode 0: scripts/generate_midi_many.py:20-46  (filtered_logit at the last position, penalties over the last 100
 *           tokens, argmax), and
 *   mode 1: scripts/generate.py:33-85  (filtered_logit, penalties over the time-bounded look-back window, k drawn by
 *           the class of the last token, top-k, one draw from the k values normalised by their sum),
 * including the bookkeeping: the chosen token is appended to `generated`, `gen_len` is advanced and `counts` (the
 * occurrences of every token inside the look-back window) is kept in step.  train.filtered_logit's SEQUENCE-axis
 * log_softmax (train.py:133-138) is carried as a running logsumexp per (sequence, token): lse <- logaddexp(lse, logits).
 * Randomness (mode 1) is supplied by the caller: uniforms[step][b][0] picks k, uniforms[step][b][1] is the draw
 * (inverse CDF), step = gen_len[b] - prompt_len.
 * ------------------------------------------------------------------------------------------ */
typedef struct MambaSampleStepArgs {
  int32_t struct_size;
  int32_t mode;
  int32_t batch, vocab;
  int32_t bucket_bounds[4]; /* {dyn-1, length-1, time-1, tempo-1}: torch.bucketize boundaries of train.py:114-131 */
  int32_t class_bounds[4];  /* {dyn, length, time, tempo}: first token id of each class after pitch             */
  int32_t pen_rule[5];      /* per token class: 0 none, 1 pen_table[class][min(count, 127)], 2 (1.1*count if count >= 10) */
  int32_t prompt_len;
  int32_t time_budget;      /* mode 1: 64*16 (scripts/generate.py:43) */
  int32_t reserved;
  const float* pen_table;   /* [5, 128] fp32: min(base ** count, cap) per token class, tabulated by the caller in double
                               precision exactly as the python loops evaluate it (every cap is reached below count 20) */
  const float* logits;  int64_t logits_bs;   /* [B, V] fp32: the model's output for the last position */
  float* lse;               /* [B, V] running logsumexp over the positions seen so far (updated) */
  const float* dist;        /* [5, V] weights of train.make_distributions (train.py:79-111)     */
  int32_t* counts;          /* [B, V] occurrences inside the look-back window (updated)         */
  int64_t* generated;   int64_t generated_bs; /* [B, capacity] token ids (appended)             */
  int32_t* gen_len;         /* [B] number of valid entries of `generated` (advanced)            */
  int64_t* next_token;      /* [B] the chosen token                                             */
  const float* uniforms;    /* mode 1: [steps, B, 2] uniforms in [0, 1)                          */
  int32_t* win_q;           /* mode 1: [B] left edge of the look-back window (updated)           */
  int32_t* win_sum;         /* mode 1: [B] sum of time shifts over generated[win_q:] (updated)   */
} MambaSampleStepArgs;

int mamba_sample_step(const MambaSampleStepArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * One new token through the whole model in ONE persistent kernel (csrc/decode.cu; Layout P, fp32 activations and
 * states, fp32 or bf16 weights, batch <= 16).  Per layer: RMSNorm(hidden + resid) -> in_proj -> conv step + SiLU ->
 * x_proj -> dt_proj + softplus + SSM step + D skip + gate -> out_proj  (simple_mamba.pyc @L179, @L228-245 for one
 * position), then the final norm and the LM head (@L94): the 42 launches of the step-by-step path (mamba_fused_linear_step,
 * mamba_linear_step, mamba_ssm_step) become one cooperative grid (one CTA per SM) with a grid barrier between phases.
 * `layers` is a DEVICE array of n_layers descriptors; `token` [B] int64 is read on the device (it is the sampler's
 * output of the previous step); `logits` [B, vocab] fp32 feeds mamba_sample_step.  conv_state / ssm_state are updated
 * in place.  `scratch` (mamba_decode_token_scratch_bytes) and `barrier` (mamba_decode_token_barrier_bytes: the two
 * counters of the grid barrier followed by room for time stamps) are caller-owned; the caller zeroes the first 16 bytes
 * of `barrier` once (and again after a failed launch), every completed call leaves them zero.
 * With MAMBA_DECODE_FLAG_STAMPS thread 0 of CTA 0 appends (event id << 56 | %globaltimer ns) as uint64 from byte 16 of
 * `barrier`: id 0 at kernel start; per layer 1..8 (end of phase / barrier open, for the 4 phases) and 10..14 inside the
 * phases; 9 at the end of the head (at most 16*n_layers + 8 values).
 * Supported shapes: (d_model, d_inner) = (1024, 2048) or (128, 256); d_state, dt_rank multiples of 4 up to 64; d_conv 4.
 * ------------------------------------------------------------------------------------------ */
#define MAMBA_DECODE_FLAG_STAMPS 1
#define MAMBA_DECODE_FLAG_BARRIER_STAMPS 2 /* every CTA c: uint64 ns [4*n_layers][arrive, leave] after the event stamps */
typedef struct MambaDecodeLayer {
  const float* norm_weight;       /* [d_model] RMSNorm of the residual block                 */
  const void* in_proj_weight;     /* [2*d_inner, d_model]  (w_dtype)                         */
  const void* in_proj_bias;       /* [2*d_inner] or NULL                                      */
  const float* conv_weight;       /* [d_inner, d_conv]                                        */
  const float* conv_bias;         /* [d_inner] or NULL                                        */
  float* conv_state;              /* [B, d_inner, d_conv] fp32, updated                       */
  const void* x_proj_weight;      /* [dt_rank + 2*d_state, d_inner]  (w_dtype)                */
  const float* dt_weight;         /* [d_inner, dt_rank] fp32                                  */
  const float* dt_bias;           /* [d_inner] or NULL                                        */
  const float* A;                 /* [d_inner, d_state] (= -exp(A_log))                       */
  const float* D;                 /* [d_inner] or NULL                                        */
  float* ssm_state;               /* [B, d_inner, d_state] fp32, updated                      */
  const void* out_proj_weight;    /* [d_model, d_inner]  (w_dtype)                            */
  const void* out_proj_bias;      /* [d_model] or NULL                                        */
} MambaDecodeLayer;

typedef struct MambaDecodeTokenArgs {
  int32_t struct_size;
  int32_t w_dtype;
  int32_t batch, n_layers, vocab;
  int32_t d_model, d_inner, d_state, dt_rank, d_conv;
  float eps;
  int32_t flags;                  /* MAMBA_DECODE_FLAG_*                                      */
  const int64_t* token;           /* [B] the token each sequence consumes                    */
  const void* embedding;          /* [vocab, d_model]  (w_dtype)                              */
  const MambaDecodeLayer* layers; /* DEVICE array [n_layers]                                  */
  const float* norm_f_weight;     /* [d_model]                                                */
  const void* head_weight;        /* [vocab, d_model]  (w_dtype; the tied embedding)          */
  const void* head_bias;          /* [vocab] or NULL                                          */
  float* logits;  int64_t logits_bs;  /* [B, vocab] fp32                                      */
  float* scratch; size_t scratch_bytes;
  unsigned int* barrier;          /* mamba_decode_token_barrier_bytes(n_layers); first 16 bytes zero */
} MambaDecodeTokenArgs;

size_t mamba_decode_token_scratch_bytes(int d_model, int d_inner, int d_state, int dt_rank);
size_t mamba_decode_token_barrier_bytes(int n_layers);
int mamba_decode_token(const MambaDecodeTokenArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * RMSNorm with fused residual add.  Replaces RMSNorm.forward (simple_mamba.pyc @L346) and the
 * `+ x` of ResidualBlock.forward (@L179):
 *     r = x + residual   (either operand may be NULL, not both; the sum is rounded to resid_dtype
 *                         when residual != NULL so that what is normalised is what is stored)
 *     y = r * rsqrt(mean(r^2) + eps) * weight
 * Two element types: `dtype` for the mixer-side activations (x, y, dy, dx) and `resid_dtype` for the
 * residual stream (residual, resid_out, dresid_in, dresid_out) — the stream can stay fp32 while the
 * mixer runs in bf16.  Supported (dtype, resid_dtype): (F32,F32), (BF16,F32), (BF16,BF16).
 * Backward: `residual` must point to the tensor that was normalised (resid_out of the forward, or x
 * itself — then pass resid_dtype = dtype); with g = dy*w, rh = r*rstd:
 *     dr = rstd * (g - rh * mean(g*rh)) + dresid_in;   dweight[c] = sum_rows dy*rh   (overwritten)
 * dr is written to dx (dtype) and/or dresid_out (resid_dtype); at least one must be non-NULL.
 * ------------------------------------------------------------------------------------------ */
typedef struct MambaNormArgs {
  int32_t struct_size;
  int32_t dtype;
  int32_t resid_dtype;
  int32_t dim;
  int64_t rows;
  float eps;
  int32_t reserved;
  const void* x;        /* [rows, dim] dtype, contiguous, or NULL      */
  const void* residual; /* [rows, dim] resid_dtype or NULL             */
  const float* weight;  /* [dim] fp32                                  */
  void* y;              /* [rows, dim] dtype                           */
  void* resid_out;      /* [rows, dim] resid_dtype or NULL             */
  float* rstd;          /* [rows] fp32 or NULL (saved for backward)    */
  /* backward only */
  const void* dy;        /* [rows, dim] dtype                          */
  const void* dresid_in; /* [rows, dim] resid_dtype or NULL: gradient arriving on resid_out */
  void* dx;              /* [rows, dim] dtype or NULL                  */
  void* dresid_out;      /* [rows, dim] resid_dtype or NULL            */
  float* dweight;        /* [dim] fp32 overwritten                     */
  void* workspace;
  size_t workspace_bytes;
} MambaNormArgs;

int mamba_rmsnorm_fwd(const MambaNormArgs* args, void* stream);
int mamba_rmsnorm_bwd(const MambaNormArgs* args, void* stream);
size_t mamba_rmsnorm_bwd_workspace_bytes(int64_t rows, int dim);

/* ------------------------------------------------------------------------------------------
 * Grammar-masked loss.  Replaces train.py:133-138 (filtered_logit: weights picked by the bucket of the
 * PREVIOUS token from a [5, V] table, log_softmax over dim=1 — the SEQUENCE axis —, f = -log_probs *
 * weights) and train.py:161-165 (CrossEntropyLoss over the vocab axis, mean over B*T):
 *     loss = mean_{b,t} [ logsumexp_v f[b,t,:] - f[b,t,trg[b,t]] ]
 * `boundaries` are the 4 bucket edges of train.py:117-121 (bucketize, right=False); `table` is the
 * [5, vocab] fp32 tensor of train.make_distributions (train.py:79-111).  logits: [B, T, V] in `dtype`,
 * vocab axis contiguous, batch / time strides in elements (the padded LM-head output is consumed in
 * place).  The forward writes loss[1], col_lse[B, V] and row_lse[B, T] (fp32; both are inputs of the
 * backward).  The backward writes dlogits = d(loss * grad_out[0]) / d logits in `dtype`.
 * ------------------------------------------------------------------------------------------ */
typedef struct MambaLossArgs {
  int32_t struct_size;
  int32_t dtype;
  int32_t batch, seqlen, vocab;
  int32_t boundaries[4];
  int32_t reserved;
  const void* logits;  int64_t logits_bs, logits_ts;
  const int64_t* src;   /* [B, T] int64, contiguous */
  const int64_t* trg;   /* [B, T] int64, contiguous */
  const float* table;   /* [5, V] fp32              */
  float* col_lse;       /* [B, V] fp32              */
  float* row_lse;       /* [B, T] fp32              */
  float* loss;          /* [1] fp32 (forward)       */
  const float* grad_out; /* [1] fp32 device scalar or NULL (= 1)  (backward) */
  void* dlogits;       int64_t dlogits_bs, dlogits_ts;  /* [B, T, V] dtype (backward) */
  void* workspace;
  size_t workspace_bytes; /* >= mamba_filtered_ce_workspace_bytes(); the forward's contents are not needed by the backward */
} MambaLossArgs;

int mamba_filtered_ce_fwd(const MambaLossArgs* args, void* stream);
int mamba_filtered_ce_bwd(const MambaLossArgs* args, void* stream);
size_t mamba_filtered_ce_workspace_bytes(int batch, int seqlen, int vocab);

/* ------------------------------------------------------------------------------------------ */
int mamba_abi_version(void);
const char* mamba_last_error(void);
/* Number of kernels this library has launched from the calling process (monotonic counter;
 * bench.py reports the difference across the timed region as "gpu_launches"). */
uint64_t mamba_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MAMBA_B200_H_ */
