"""Host-side mirror of the reference's generation loops for the Mamba path.

  * `generate_literal`   — scripts/generate_midi_many.py:13-56 as written: every new token re-runs the FULL
    model over the (sliding) window, applies train.filtered_logit, the repetition penalties and argmax.
    Runs on the fused kernels; token-for-token comparable with the oracle loop.
  * `generate_recurrent` — prefill once, then one `Mamba.step` per token (state carried in HBM), O(1) per
    token.  filtered_logit's sequence-axis log_softmax is reproduced with a running
    per-(batch, vocab) logsumexp over all positions seen; exact w.r.t. the literal loop while the window
    has not started to slide (T_prompt + generated <= context_len), and the only mode that scales.
    Batched: B independent sequences ("batch of 5 composer conditions"), shardable across GPUs with no
    collective.
"""
from __future__ import annotations

import torch

from . import ops, train
from .configs import common as cc


def _penalty_table_many(device):
    """Per-token (base, cap, rule) of scripts/generate_midi_many.py:24-43 as dense vectors over the vocab:
    rule 0 = no penalty, 1 = min(base**count, cap), 2 = 1.1*count if count >= 10."""
    V, s = cc.vocab_size, cc.start_idx
    rule = torch.zeros(V, dtype=torch.long, device=device)
    base = torch.ones(V, device=device)
    cap = torch.ones(V, device=device)
    rule[s["time"]:s["tempo"]] = 2
    rule[s["length"]:s["time"]] = 1
    base[s["length"]:s["time"]], cap[s["length"]:s["time"]] = 1.015, 1.08
    rule[:s["dyn"]] = 1
    base[:s["dyn"]], cap[:s["dyn"]] = 1.04, 1.25
    return rule, base, cap


def _apply_penalty_many(logits_last, counts, table):
    """logits_last [B, V] /= penalty(count) with count = occurrences in the last 100 tokens (on device)."""
    rule, base, cap = table
    c = counts.to(logits_last.dtype)
    # python's `1.04 ** count` is a double power; the float64 detour keeps argmax ties identical
    p1 = torch.minimum(torch.pow(base.double(), c.double()), cap.double()).to(logits_last.dtype)
    p2 = torch.where(c >= 10, 1.1 * c, torch.ones_like(c))
    pen = torch.where(rule == 1, p1, torch.where(rule == 2, p2, torch.ones_like(c)))
    pen = torch.where(c > 0, pen, torch.ones_like(pen))
    return logits_last / pen


@torch.no_grad()
def generate_literal(model, context_len, token_ids, meta_ids, num_tokens=1000):
    """scripts/generate_midi_many.py:13-56 (batch generalised: each row is an independent sequence)."""
    model.eval()
    dev = token_ids.device
    B = token_ids.shape[0]
    table = _penalty_table_many(dev)
    generated = token_ids.clone()
    for _ in range(num_tokens):
        logits = model(token_ids, meta_ids)
        filtered = train.filtered_logit(token_ids, logits)
        logits_last = filtered[:, -1, :]
        recent = generated[:, -100:]
        counts = torch.zeros(B, cc.vocab_size, device=dev).scatter_add_(
            1, recent, torch.ones_like(recent, dtype=torch.float32))
        logits_last = _apply_penalty_many(logits_last, counts, table)
        next_token = logits_last.argmax(-1, keepdim=True)
        generated = torch.cat([generated, next_token], dim=1)
        token_ids = torch.cat([token_ids, next_token], dim=1)[:, -context_len:]
    return generated


class RecurrentDecoder:
    """Prefill + CUDA-graphed single-token step for a fixed batch of independent sequences.

    One step = ONE launch of `mamba_decode_token` (csrc/decode.cu: the whole model as one persistent cooperative
    kernel; Layout P, fp32 states) — or, for models it does not cover, the embedding lookup + 4 launches per layer of
    `Mamba.step`'s fused path + the head — and ONE launch of `mamba_sample_step`, which applies filtered_logit and the
    repetition penalties, picks the token, appends it to `generated` and advances the look-back window.
    mode "many":   greedy, scripts/generate_midi_many.py:13-56.
    mode "sample": top-k draw of scripts/generate.py:14-95; `uniforms` [max_new, B, 2] supplies the randomness
                   (default: torch.rand from `generator`, i.e. a Philox stream with a fixed seed)."""

    def __init__(self, model, batch_size, use_graph=True, dtype=None, mode="many", max_new_tokens=4096, uniforms=None,
                 generator=None, persistent=None, weight_dtype=None):
        if mode not in ("many", "sample"):
            raise ValueError("mode must be 'many' or 'sample'")
        self.model = model.eval()
        self.B = batch_size
        self.mode = 0 if mode == "many" else 1
        self.dev = next(model.parameters()).device
        self.cache = model.allocate_inference_cache(batch_size, dtype=dtype)
        V = cc.vocab_size
        self.max_new = int(max_new_tokens)
        self.lse = torch.zeros(batch_size, V, device=self.dev)                      # running logsumexp over positions
        self.counts = torch.zeros(batch_size, V, dtype=torch.int32, device=self.dev)  # occurrences in the window
        self.logits = torch.zeros(batch_size, V, device=self.dev)
        self.nxt = torch.zeros(batch_size, dtype=torch.long, device=self.dev)
        self.gen_len = torch.zeros(batch_size, dtype=torch.int32, device=self.dev)
        self.win_q = torch.zeros(batch_size, dtype=torch.int32, device=self.dev)
        self.win_sum = torch.zeros(batch_size, dtype=torch.int32, device=self.dev)
        self.generated = None
        self.uniforms = uniforms
        self.generator = generator
        self.dist = train.make_distributions(self.dev).contiguous()
        self.use_graph = use_graph
        self.graph = None
        self.sargs = None
        # one-launch-per-token path (csrc/decode.cu): default whenever the model qualifies; `weight_dtype=torch.bfloat16`
        # streams decode-only bf16 copies of the weight matrices (activations, states and accumulation stay fp32)
        self.plan = None
        want = ops.DecodeTokenPlan.eligible(model, self.cache, batch_size) if persistent is None else bool(persistent)
        if want:
            if not ops.DecodeTokenPlan.eligible(model, self.cache, batch_size):
                raise ValueError("RecurrentDecoder(persistent=True): model / cache not supported by mamba_decode_token")
            self.plan = ops.DecodeTokenPlan(model, self.cache, self.nxt, self.logits, weight_dtype)
        elif weight_dtype is not None:
            raise ValueError("weight_dtype needs the persistent decode kernel")

    @torch.no_grad()
    def prefill(self, token_ids, meta_ids):
        B, T = token_ids.shape
        s = cc.start_idx
        logits = self.model.prefill(token_ids, meta_ids, self.cache).float()      # [B, T, V]
        # sequence-axis logsumexp over the positions BEFORE the last one; the sample kernel adds the last
        if T > 1:
            self.lse.copy_(torch.logsumexp(logits[:, :-1], dim=1))
        else:
            self.lse.fill_(float("-inf"))
        self.logits.copy_(logits[:, -1])
        self.generated = torch.zeros(B, T + self.max_new + 1, dtype=torch.long, device=self.dev)
        self.generated[:, :T] = token_ids
        self.gen_len.fill_(T)
        # look-back window over the prompt (host side, once): counts of the tokens inside it
        toks = token_ids.cpu()
        counts = torch.zeros(B, cc.vocab_size, dtype=torch.int32)
        q0 = torch.zeros(B, dtype=torch.int32)
        sum0 = torch.zeros(B, dtype=torch.int32)
        for b in range(B):
            row = toks[b].tolist()
            if self.mode == 0:
                win = row[-100:]
            else:
                tv = [t - s["time"] if s["time"] <= t < s["tempo"] else 0 for t in row]
                q, tot = 0, sum(tv)
                while q < T - 1 and tot - tv[q] >= 64 * 16:
                    tot -= tv[q]
                    q += 1
                q0[b], sum0[b] = q, tot
                win = row[q + 1:] if T > 1 else row     # scripts/generate.py:38-46 (`cur_gen[-0:]` is the whole list)
            for t in win:
                counts[b, t] += 1
        self.counts.copy_(counts)
        self.win_q.copy_(q0)
        self.win_sum.copy_(sum0)
        if self.mode == 1 and self.uniforms is None:
            self.uniforms = torch.rand(self.max_new + 1, B, 2, device=self.dev, generator=self.generator)
        self.sargs = ops.sample_step_args(
            self.mode, self.logits, self.lse, self.dist, self.counts, self.generated, self.gen_len, self.nxt,
            (s["dyn"], s["length"], s["time"], s["tempo"]), T, uniforms=self.uniforms, win_q=self.win_q, win_sum=self.win_sum)
        ops.sample_step(self.sargs, self.dev)
        return self.nxt.clone()

    def _step_body(self):
        if self.plan is not None:
            self.plan.run()                                   # the whole model, one cooperative launch
        else:
            self.model.step(self.nxt, self.cache, logits_out=self.logits)
        ops.sample_step(self.sargs, self.dev)

    @torch.no_grad()
    def step(self):
        """Consume the previously chosen token, produce the next one (device tensor [B])."""
        if self.use_graph:
            if self.graph is None:
                s = torch.cuda.Stream(device=self.dev)
                s.wait_stream(torch.cuda.current_stream(self.dev))
                saved = self._snapshot()
                with torch.cuda.stream(s):
                    self._step_body()
                torch.cuda.current_stream(self.dev).wait_stream(s)
                torch.cuda.synchronize(self.dev)
                self._restore(saved)
                saved = self._snapshot()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._step_body()
                self._restore(saved)
            self.graph.replay()
        else:
            self._step_body()
        return self.nxt

    def tokens(self):
        """[B, prompt + generated so far] (all rows have the same length)."""
        n = int(self.gen_len[0])
        return self.generated[:, :n].clone()

    def _state_tensors(self):
        ts = [self.lse, self.counts, self.logits, self.nxt, self.gen_len, self.win_q, self.win_sum, self.generated]
        for cs, hs in self.cache:
            ts += [cs, hs]
        return ts

    def _snapshot(self):
        return [t.clone() for t in self._state_tensors()]

    def _restore(self, saved):
        for t, s in zip(self._state_tensors(), saved):
            t.copy_(s)


@torch.no_grad()
def generate_recurrent(model, token_ids, meta_ids, num_tokens=1000, use_graph=True, dtype=None, mode="many",
                       uniforms=None, generator=None, persistent=None, weight_dtype=None):
    """`num_tokens` new tokens for each row of token_ids; returns [B, T + num_tokens].  mode "many": greedy
    (scripts/generate_midi_many.py); mode "sample": the top-k draw of scripts/generate.py with the given uniforms."""
    dec = RecurrentDecoder(model, token_ids.shape[0], use_graph=use_graph, dtype=dtype, mode=mode,
                           max_new_tokens=num_tokens, uniforms=uniforms, generator=generator, persistent=persistent,
                           weight_dtype=weight_dtype)
    dec.prefill(token_ids, meta_ids)
    for _ in range(1, num_tokens):
        dec.step()
    return dec.tokens()


def generate(model, context_len, token_ids, meta_ids, num_tokens=1000, device="cuda", seed=None):
    """Drop-in for scripts/generate.py:14-95 (same arguments, returns a list of token lists): recurrent decode with
    the on-device sampler.  The window never slides (state is carried), so `context_len` only bounds the prompt."""
    token_ids, meta_ids = token_ids.to(device), meta_ids.to(device)
    gen = None
    if seed is not None:
        gen = torch.Generator(device=token_ids.device).manual_seed(int(seed))
    out = generate_recurrent(model, token_ids[:, -context_len:], meta_ids, num_tokens, mode="sample", generator=gen)
    return [row.tolist() for row in out.cpu()]
