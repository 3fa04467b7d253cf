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

from . import train
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
    """Prefill + CUDA-graphed single-token step for a fixed batch size."""

    def __init__(self, model, batch_size, use_graph=True, dtype=None):
        self.model = model.eval()
        self.B = batch_size
        self.dev = next(model.parameters()).device
        self.cache = model.allocate_inference_cache(batch_size, dtype=dtype)
        self.table = _penalty_table_many(self.dev)
        V = cc.vocab_size
        self.lse = torch.zeros(batch_size, V, device=self.dev)            # running logsumexp over positions
        self.counts = torch.zeros(batch_size, V, device=self.dev)         # occurrences in the last 100 tokens
        self.window = torch.zeros(batch_size, 100, dtype=torch.long, device=self.dev)
        self.wfill = 0
        self.cur = torch.zeros(batch_size, dtype=torch.long, device=self.dev)   # last token (input of the step)
        self.nxt = torch.zeros(batch_size, dtype=torch.long, device=self.dev)
        self.use_graph = use_graph
        self.graph = None

    @torch.no_grad()
    def prefill(self, token_ids, meta_ids):
        logits = self.model.prefill(token_ids, meta_ids, self.cache).float()      # [B, T, V]
        self.lse.copy_(torch.logsumexp(logits, dim=1))
        recent = token_ids[:, -100:]
        self.counts.zero_().scatter_add_(1, recent, torch.ones_like(recent, dtype=torch.float32))
        n = recent.shape[1]
        self.window.zero_()
        self.window[:, 100 - n:] = recent
        self.wfill = n
        self._choose(logits[:, -1, :], token_ids[:, -1])
        return self.nxt.clone()

    def _choose(self, logits_last, prev_token):
        """filtered_logit at the last position + penalties + argmax (generate_midi_many.py:20-46)."""
        weights = train.pick_distributions_by_prev_token(prev_token)             # [B, V]
        f = -(logits_last - self.lse) * weights
        f = _apply_penalty_many(f, self.counts, self.table)
        self.nxt.copy_(f.argmax(-1))

    def _slide(self, tok):
        """Push `tok` into the 100-token look-back window and keep `counts` in step."""
        full = self.wfill >= 100
        if full:
            old = self.window[:, 0:1]
            self.counts.scatter_add_(1, old, -torch.ones_like(old, dtype=torch.float32))
        self.window.copy_(torch.cat((self.window[:, 1:], tok[:, None]), dim=1))
        self.counts.scatter_add_(1, tok[:, None], torch.ones(self.B, 1, device=self.dev))
        if not full:
            self.wfill += 1

    def _step_body(self):
        self.cur.copy_(self.nxt)
        self._slide(self.cur)
        logits = self.model.step(self.cur, self.cache).float()
        self.lse.copy_(torch.logaddexp(self.lse, logits))
        self._choose(logits, self.cur)

    @torch.no_grad()
    def step(self):
        """Consume the previously chosen token, produce the next one (device tensor [B])."""
        if self.use_graph and self.wfill >= 100:
            if self.graph is None:
                s = torch.cuda.Stream(device=self.dev)
                s.wait_stream(torch.cuda.current_stream(self.dev))
                saved = self._snapshot()
                with torch.cuda.stream(s):
                    self._step_body()
                torch.cuda.current_stream(self.dev).wait_stream(s)
                torch.cuda.synchronize(self.dev)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._step_body()
                self._restore(saved)
            self.graph.replay()
        else:
            self._step_body()
        return self.nxt

    def _state_tensors(self):
        ts = [self.lse, self.counts, self.window, self.cur, self.nxt]
        for cs, hs in self.cache:
            ts += [cs, hs]
        return ts

    def _snapshot(self):
        return [t.clone() for t in self._state_tensors()]

    def _restore(self, saved):
        for t, s in zip(self._state_tensors(), saved):
            t.copy_(s)


@torch.no_grad()
def generate_recurrent(model, token_ids, meta_ids, num_tokens=1000, use_graph=True, dtype=None):
    """Greedy decode of `num_tokens` new tokens for each row of token_ids; returns [B, T + num_tokens]."""
    dec = RecurrentDecoder(model, token_ids.shape[0], use_graph=use_graph, dtype=dtype)
    out = torch.empty(token_ids.shape[0], num_tokens, dtype=torch.long, device=token_ids.device)
    out[:, 0] = dec.prefill(token_ids, meta_ids)
    for i in range(1, num_tokens):
        out[:, i] = dec.step()
    return torch.cat((token_ids, out), dim=1)
