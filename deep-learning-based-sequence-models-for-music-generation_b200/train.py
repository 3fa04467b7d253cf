"""Host-side mirror of the reference's train.py / train_parallel.py entry points for the Mamba path.

Same function names and behaviour as the reference: `get_actual_vocab_size`, `get_mamba_dict`
(train.py:21-36), `new_model` (:52-61), `make_distributions` (:79-111),
`pick_distributions_by_prev_token` (:114-131), `filtered_logit` (:133-138).  The reference's epoch loop
(:140-217) is I/O around one repeated step; `train_step` is that step (:160-169) and `Trainer` runs it
as one CUDA graph (forward, loss, backward, gradient all-reduce, Adam) — the B200-native replacement for
the python-launched loop.  Data loading, checkpoint naming and JSON logging stay the caller's.

The loss is outside the scan/conv hot path but sits on the fwd+bwd step, so it is restated verbatim —
including its quirk that `log_softmax` runs over dim=1, the SEQUENCE axis.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import models
from .configs import common as cc
from .configs import mamba as cm


def get_actual_vocab_size(type):  # train.py:21-26
    config = cm.config.model_values
    new_vocab_size = cc.vocab_size
    if type == "mamba":
        if new_vocab_size % config.pad_vocab_size_multiple != 0:
            new_vocab_size += config.pad_vocab_size_multiple - new_vocab_size % config.pad_vocab_size_multiple
    return new_vocab_size


def get_mamba_dict(pad_vocab: bool = False):
    """train.py:28-36.  The reference pads vocab_size to a multiple of 8 here (17914 -> 17920) and then
    reshapes logits to cc.vocab_size columns in its loop (:162), which cannot both hold; like the shipped
    wrapper (models/mamba/mamba.py:12-14) the default here is the unpadded vocabulary."""
    mv = cm.config.model_values
    config = SimpleNamespace(**vars(mv))
    config.d_inner = int(config.expand * config.d_model)
    config.dt_rank = math.ceil(config.d_model / 16)
    config.vocab_size = get_actual_vocab_size("mamba") if pad_vocab else cc.vocab_size
    config.metadata_vocab_size = cc.metadata_vocab_size
    return SimpleNamespace(**vars(config), **vars(cc.config.values))


def new_model(type="mamba", layout="P"):  # train.py:52-61
    if type != "mamba":
        raise ValueError("mamba_b200 provides the Mamba path only (xlstm / transformer are out of scope)")
    if layout == "S":
        return models.mamba.Mamba()              # the shipped call, train.py:54
    return models.mamba.Mamba(get_mamba_dict())  # the pure-PyTorch model's call (train.cpython-313.pyc @L49)


_dist_cache = {}


def make_distributions(device=None):  # train.py:79-111
    device = torch.device(cc.config.values.device if device is None else device)
    key = str(device)
    if key in _dist_cache:  # constant table: built once per device instead of once per call
        return _dist_cache[key]
    vocab_size = cc.vocab_size
    distributions = torch.zeros(5, vocab_size, device=device)
    s = cc.start_idx
    start = [s["pitch"], s["dyn"], s["length"], s["time"], s["tempo"]]
    end = [s["dyn"] - 1, s["length"] - 1, s["time"] - 1, s["tempo"] - 1, vocab_size]
    for token in range(5):
        distributions[token - 1, start[token]:end[token]] = 1
    distributions[2, start[4]:end[4]] = 1
    length_tensor = torch.linspace(1, 3, steps=cc.config.discretization.length - 1).to(device)  # train.py:20
    distributions[1, s["length"]:s["time"] - 1] *= length_tensor
    distributions[4, s["pitch"]:s["dyn"] - 1] *= 10
    _dist_cache[key] = distributions
    return distributions


def pick_distributions_by_prev_token(input_tokens):  # train.py:114-131
    s = cc.start_idx
    boundaries = [s["dyn"] - 1, s["length"] - 1, s["time"] - 1, s["tempo"] - 1]
    key = ("bins", str(input_tokens.device))
    if key not in _dist_cache:
        _dist_cache[key] = torch.tensor(boundaries, device=input_tokens.device)
    buckets = torch.bucketize(input_tokens, _dist_cache[key], right=False)
    return make_distributions(input_tokens.device)[buckets]


def filtered_logit(input, output):  # train.py:133-138
    weights = pick_distributions_by_prev_token(input)
    log_probs = F.log_softmax(output.float(), dim=1)
    return -log_probs * weights


def loss_fn_torch(src, trg, output):
    """train.py:161-165 spelled with torch ops, as the reference does (about ten passes over [B, T, V])."""
    filtered_output = filtered_logit(src, output).reshape(-1, cc.vocab_size)
    return F.cross_entropy(filtered_output, trg.reshape(-1))


def loss_fn(src, trg, output):
    """train.py:161-165 (filtered_logit + CrossEntropyLoss) as one fused CUDA op: four streaming passes over the
    logits in their storage dtype, deterministic.  CUDA only."""
    from . import ops
    s = cc.start_idx
    boundaries = (s["dyn"] - 1, s["length"] - 1, s["time"] - 1, s["tempo"] - 1)
    return ops.filtered_ce_fn(output, src, trg, make_distributions(output.device), boundaries)


def train_step(model, optimizer, src, trg, meta, autocast_dtype=None):
    """One iteration of the reference's hot loop (train.py:160-169): returns the loss tensor (no .item())."""
    with torch.autocast("cuda", dtype=autocast_dtype, enabled=autocast_dtype is not None):
        output = model(src, meta)
    loss = loss_fn(src, trg, output)
    optimizer.zero_grad(set_to_none=False)
    loss.backward()
    optimizer.step()
    return loss


class FlatGrads:
    """All gradients of a model in ONE flat fp32 buffer (every p.grad is a view into it), exchanged as a few large
    buckets.  Data-parallel semantics of train_parallel.py:151 (DDP: gradient mean over ranks): each bucket is
    pre-scaled by 1/world and sum-all-reduced, which is what DDP's default hook does and works on NCCL and gloo.

    `overlap_with_backward(world, group)` arms per-parameter hooks that launch a bucket's all-reduce on a side
    stream as soon as the last gradient of that bucket has been accumulated, so the exchange of the late layers
    runs under the backward of the early ones (buckets complete in reverse layer order); `finish()` joins the side
    stream.  Both work under CUDA-graph capture (the side stream forks from / joins the capturing stream)."""

    def __init__(self, params, bucket_mb=64, groups=None):
        """`groups` (optional): list of parameter lists — one bucket per group, laid out in that order (the staged
        backward of Trainer reduces bucket i as soon as stage i's gradients exist); otherwise buckets of about
        `bucket_mb` MiB in parameter order."""
        if groups is not None:
            params = [p for g in groups for p in g]
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        per = max(1, int(bucket_mb * (1 << 20) // 4))
        # buckets end on parameter boundaries so that "bucket complete" is a count of parameters
        self.buckets, self._bucket_of = [], {}
        off, start, members = 0, 0, []
        ends = None
        if groups is not None:
            ends, acc = set(), 0
            for g in groups:
                acc += sum(1 for p in g if p.requires_grad)
                ends.add(acc)
        for k, p in enumerate(self.params):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            members.append(p)
            if (k + 1 in ends) if ends is not None else (off - start >= per):
                self._close_bucket(start, off, members)
                start, members = off, []
        if members:
            self._close_bucket(start, off, members)
        self._pending = [0] * len(self.buckets)
        self._hooks, self._side, self._world, self._group = [], None, 1, None
        self._done = {}

    def _close_bucket(self, start, end, members):
        idx = len(self.buckets)
        self.buckets.append(self.flat[start:end])
        for p in members:
            self._bucket_of[id(p)] = idx

    def zero(self):
        self.flat.zero_()
        self._pending = [sum(1 for p in self.params if self._bucket_of[id(p)] == i) for i in range(len(self.buckets))]

    def _reduce_bucket(self, i):
        import torch.distributed as dist
        b = self.buckets[i]
        if os.environ.get("MAMBA_B200_DEBUG_NO_COMM") == "1":   # measurement aid: everything but the collective
            return
        if b.is_cuda and dist.get_backend(self._group) == "nccl" and os.environ.get("MAMBA_B200_NCCL_AVG", "0") == "1":
            # the mean inside the collective saves the pre-scale pass, but ncclAvg measured SLOWER than pre-scale + sum
            # at 8 GPUs (14.68 vs 14.31 ms/step: it leaves the in-switch NVLS reduction) — opt-in only
            dist.all_reduce(b, op=dist.ReduceOp.AVG, group=self._group)
        else:
            b.mul_(1.0 / self._world)
            dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self._group)

    def allreduce_mean(self, world_size, group=None):
        """Exchange every bucket now (no overlap)."""
        if world_size <= 1:
            return
        self._world, self._group = world_size, group
        for i in range(len(self.buckets)):
            self._reduce_bucket(i)

    def overlap_with_backward(self, world_size, group=None):
        if world_size <= 1 or self._hooks:
            return
        self._world, self._group = world_size, group
        if self.flat.is_cuda:
            self._side = torch.cuda.Stream(device=self.flat.device)

        def hook(p):
            i = self._bucket_of[id(p)]
            self._pending[i] -= 1
            if self._pending[i] != 0:
                return
            if self._side is None:
                self._reduce_bucket(i)
                return
            self._side.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self._side):
                self._reduce_bucket(i)

        self._hooks = [p.register_post_accumulate_grad_hook(hook) for p in self.params]

    def launch_bucket(self, i, world_size, group=None):
        """All-reduce bucket i NOW, from the calling thread: on CUDA on a side stream forked from the current one
        (so that it runs under whatever the current stream does next), on CPU inline.  `finish()` joins."""
        self._world, self._group = world_size, group
        if world_size <= 1:
            return
        if not self.flat.is_cuda:
            self._reduce_bucket(i)
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.flat.device)
        self._side.wait_stream(torch.cuda.current_stream(self.flat.device))
        with torch.cuda.stream(self._side):
            self._reduce_bucket(i)
            ev = torch.cuda.Event()
            ev.record(self._side)
            self._done[i] = ev

    def wait_bucket(self, i):
        """Make the current stream wait for bucket i's exchange only (the optimizer of that bucket's parameters can
        then run while later buckets are still in flight)."""
        ev = self._done.pop(i, None)
        if ev is not None:
            torch.cuda.current_stream(self.flat.device).wait_event(ev)

    def finish(self):
        """Join the side stream (after backward, before the optimizer reads the gradients)."""
        if self._side is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._side)


def shard_rows(n_rows, rank, world_size):
    """Contiguous shard [lo, hi) of `n_rows` independent sequences for `rank` (batch-sharded generation: the
    sequences never exchange anything, so there is no collective on this path)."""
    base, rem = divmod(n_rows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Trainer:
    """The reference's training step as ONE CUDA graph per rank.

    `train_parallel.py:143-185` wraps the model in DDP (gradient all-reduce(mean) over NCCL, overlapped with
    backward) and launches every kernel from python.  At the reference's batch (2 x 2048 tokens per GPU) the
    step is a few milliseconds of GPU work, so launch latency and the exposed all-reduce dominate.  Here
    the whole step — H2D-staged batch -> forward -> loss -> backward -> all-reduce of the flat gradient
    buffer -> Adam — is captured once and replayed; the gradients of all parameters live in one flat buffer
    (bucket views) so the data-parallel exchange is `n_buckets` NCCL all-reduces issued inside the graph
    on a side stream as soon as each bucket's last gradient is produced.
    """

    # stages: the default overlap — staged backward, per-stage all-reduce + Adam on a side stream, issued from the
    # capturing thread (see __init__).  overlap_allreduce: the older hook-driven variant (each bucket's all-reduce
    # launched from autograd's post-accumulate hooks); verified on CPU/gloo (tests/test_dist_cpu.py), but under
    # NCCL + CUDA-graph capture it hung on the 2-GPU box, so it stays off.
    def __init__(self, model, lr=None, autocast_dtype=torch.bfloat16, world_size=1, process_group=None,
                 batch_size=None, block_len=None, use_graph=True, bucket_mb=64, overlap_allreduce=False,
                 async_wgrad=None, stages=None):
        self.model = model
        self.device = next(model.parameters()).device
        self.autocast_dtype = autocast_dtype
        self.world_size = world_size
        self.pg = process_group
        B = cc.config.values.batch_size if batch_size is None else batch_size
        T = cc.config.values.block_len if block_len is None else block_len
        self.src = torch.zeros(B, T, dtype=torch.long, device=self.device)
        self.trg = torch.zeros(B, T, dtype=torch.long, device=self.device)
        self.meta = torch.zeros(B, cc.N_META, dtype=torch.long, device=self.device)
        self.loss = torch.zeros((), device=self.device)
        # Staged backward (world_size > 1, layout P): the layer stack is cut into `stages` groups; the backward runs
        # group by group from the top and each group's gradient bucket is all-reduced on a side stream while the
        # groups below it are still being differentiated — DDP's overlap (train_parallel.py:151), issued from the
        # capturing thread so that it can live inside the CUDA graph.  stages = 1: one exchange after the backward.
        if stages is None:
            stages = int(os.environ.get("MAMBA_B200_STAGES", "5"))
        self._stage_groups = self._stage_params = None
        staged_ok = world_size > 1 or self.device.type == "cuda"   # one rank on CPU: nothing to overlap
        if staged_ok and stages > 1 and getattr(model, "layout", None) == "P" and len(model.layers) >= stages:
            layers = list(model.layers)
            per = -(-len(layers) // stages)
            self._stage_groups = [layers[i:i + per] for i in range(0, len(layers), per)]
            # parameters of stage i.  The final norm joins the last stage (its gradient is the first one to exist).
            # Everything else outside the layer stack (embeddings, the tied head) receives its last gradient
            # contribution at the very end of the backward and forms a stage of its own, AFTER stage 0: the bottom
            # layers' bucket is then exchanged under the embedding backward instead of waiting for it, and what stays
            # exposed at the end of the step is the exchange + update of the embedding matrix alone.
            groups = [[p for layer in g for p in layer.parameters()] for g in self._stage_groups]
            seen = {id(p) for g in groups for p in g}
            groups[-1] = groups[-1] + [p for p in model.norm_f.parameters() if id(p) not in seen]
            seen |= {id(p) for p in model.norm_f.parameters()}
            rest = [p for p in model.parameters() if id(p) not in seen]
            if any(p.requires_grad for p in rest):
                groups.append(rest)
            self._stage_params = [[p for p in g if p.requires_grad] for g in groups]
        # DDP broadcasts rank 0's parameters and buffers at construction (train_parallel.py:151); do the same, so that
        # ranks that seeded differently (or loaded different checkpoints) cannot apply averaged gradients to
        # different weights
        if world_size > 1:
            self._broadcast_parameters()
        self._flatten_grads(bucket_mb)
        # weight-gradient GEMMs of the mixer's linear layers on a side stream (models.mamba.AsyncWgrad): needs
        # autocast (the fp32-output GEMM path) and gradient buffers that exist before backward
        # (measured: +1 % on one GPU — cuBLAS's 2-CTA-cluster GEMMs find few free TPCs next to the scans — and +2 % at
        # two GPUs, where it also saves the accumulate pass into the flat gradient buffer)
        if async_wgrad is None:
            async_wgrad = os.environ.get("MAMBA_B200_ASYNC_WGRAD", "1") == "1"
        self.async_wgrad = bool(async_wgrad) and self.device.type == "cuda" and autocast_dtype is not None
        self._wgrad_params, self._wgrad_stream = [], None
        if self.async_wgrad:
            from .models.mamba.mamba import MambaBlock
            self._wgrad_stream = torch.cuda.Stream(device=self.device)
            for m in model.modules():
                if isinstance(m, MambaBlock):
                    for lin in (m.in_proj, m.x_proj, m.dt_proj, m.out_proj):
                        if lin.bias is None or lin is m.dt_proj:
                            self._wgrad_params.append(lin.weight)
            if self.grads is None:
                for p in self._wgrad_params:
                    p.grad = torch.zeros_like(p)
        self._wgrad_ids = frozenset(id(p) for p in self._wgrad_params)
        # the hook-driven overlap and the staged step both exchange every bucket: together they would reduce twice
        # (gradients scaled by 1/world^2), so the staged step wins and the hooks are not armed
        if overlap_allreduce and self._stage_groups is not None:
            import warnings
            warnings.warn("Trainer: overlap_allreduce is ignored when the staged step is active (stages > 1); "
                          "pass stages=1 to use the hook-driven overlap")
            overlap_allreduce = False
        self.overlap = overlap_allreduce and world_size > 1
        if self.overlap:
            self.grads.overlap_with_backward(world_size, process_group)
        lr = cc.config.values.learning_rate if lr is None else lr
        self.optimizer = torch.optim.Adam(model.parameters(), lr=lr, capturable=use_graph, fused=True)  # train.py:146
        # staged exchange: the same Adam, one instance per gradient bucket, so that bucket k's parameters are updated
        # as soon as ITS all-reduce is done — the update of the upper layers hides the exchange of the last bucket
        # On one GPU the same structure overlaps the (memory-bound) parameter update of the upper stages with the
        # (compute-bound) backward of the lower ones.
        self._stage_optimizers, self._side = None, None
        if self._stage_groups is not None and self.device.type == "cuda":
            self._stage_optimizers = [torch.optim.Adam(ps, lr=lr, capturable=use_graph, fused=True)
                                      for ps in self._stage_params]
            self._side = torch.cuda.Stream(device=self.device)
            if self.grads is not None:
                self.grads._side = self._side
        # autocast-dtype shadows of the mixer's linear weights, refreshed after each Adam update (off the critical
        # path) instead of being re-cast in every forward (models.mamba.WeightShadows)
        self._shadow_params = []
        if self.device.type == "cuda" and autocast_dtype is not None and os.environ.get("MAMBA_B200_SHADOWS", "1") == "1":
            from .models.mamba.mamba import MambaBlock, WeightShadows
            for m in model.modules():
                if isinstance(m, MambaBlock):
                    for lin in (m.in_proj, m.x_proj, m.dt_proj, m.out_proj):
                        if lin.bias is None or lin is m.dt_proj:
                            self._shadow_params.append(lin.weight)
            self._shadow_refs = WeightShadows.register(self._shadow_params, autocast_dtype)
        self.use_graph = use_graph
        self.graph = None

    @property
    def optimizers(self):
        """The Adam instances that actually step: one per stage in the staged step (their union covers every
        parameter exactly once), otherwise the single `self.optimizer`.  Use this for state_dict()/load_state_dict()."""
        return list(self._stage_optimizers) if self._stage_optimizers is not None else [self.optimizer]

    def _refresh_shadows(self, params=None):
        if self._shadow_params:
            from .models.mamba.mamba import WeightShadows
            WeightShadows.refresh(self._shadow_params if params is None else params)

    def _flatten_grads(self, bucket_mb):
        # one rank: autograd writes each gradient straight into a fresh buffer (no accumulate-add per parameter);
        # several ranks: gradients live in one flat buffer so that the exchange is a few large all-reduces
        if self.world_size <= 1:
            self.grads = None
        elif self._stage_groups is None:
            self.grads = FlatGrads(self.model.parameters(), bucket_mb)
        else:
            self.grads = FlatGrads(None, bucket_mb, groups=self._stage_params)   # bucket i = parameters of stage i

    def _allreduce(self):
        if self.grads is None:
            return
        if self.overlap:
            self.grads.finish()          # the buckets were launched from the backward hooks
        else:
            self.grads.allreduce_mean(self.world_size, self.pg)

    def _join_wgrad(self):
        if self.async_wgrad:
            torch.cuda.current_stream(self.device).wait_stream(self._wgrad_stream)

    def _wgrad_scope(self):
        """Context in which this Trainer's backward may write ITS weight gradients on the side stream."""
        if not self.async_wgrad:
            import contextlib
            return contextlib.nullcontext()
        from .models.mamba.mamba import AsyncWgrad
        return AsyncWgrad.scope(self._wgrad_stream, self._wgrad_ids)

    def _broadcast_parameters(self):
        """Rank 0's parameters and buffers to every rank, one flat collective per dtype."""
        import torch.distributed as dist
        with torch.no_grad():
            by_dtype = {}
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                by_dtype.setdefault(t.dtype, []).append(t)
            for ts in by_dtype.values():
                seen, uniq = set(), []
                for t in ts:   # tied parameters appear once
                    if t.data_ptr() not in seen:
                        seen.add(t.data_ptr())
                        uniq.append(t)
                flat = torch.cat([t.detach().reshape(-1) for t in uniq])
                dist.broadcast(flat, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
                off = 0
                for t in uniq:
                    t.copy_(flat[off:off + t.numel()].view_as(t))
                    off += t.numel()

    def _step_body_staged(self):
        m, groups = self.model, self._stage_groups
        cuts = []
        with torch.autocast(self.device.type, dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            emb = m._embed(self.src, self.meta)
            # cut below the layer stack too: the embeddings (and the tied head) are the stage after stage 0
            resid, hidden = (emb.detach().requires_grad_(True) if emb.requires_grad else emb), None
            emb_in = resid
            for gi, group in enumerate(groups):
                if gi > 0:  # cut the autograd graph below this group
                    rd, hd = resid.detach().requires_grad_(True), hidden.detach().requires_grad_(True)
                    cuts.append((resid, hidden, rd, hd))
                    resid, hidden = rd, hd
                for layer in group:
                    normed, resid = layer.norm(hidden, resid)
                    hidden = layer.mixer(normed)
            normed, _ = m.norm_f(hidden, resid)
            output = m._head(normed[:, self.meta.shape[-1]:])
        loss = loss_fn(self.src, self.trg, output)
        self._zero_grads()
        with self._wgrad_scope():
            loss.backward()
        self._finish_stage(len(groups) - 1)
        for gi in range(len(cuts) - 1, -1, -1):
            r, h, rd, hd = cuts[gi]
            with self._wgrad_scope():
                torch.autograd.backward([r, h], [rd.grad, hd.grad])
            self._finish_stage(gi)
        if emb_in is not emb:
            torch.autograd.backward([emb], [emb_in.grad])
        if len(self._stage_params) > len(groups):   # the embedding / tied-head stage
            self._finish_stage(len(groups))
        if self._stage_optimizers is None:
            self.grads.finish()
            self.optimizer.step()
            self._refresh_shadows()
        else:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
        self.loss.copy_(loss.detach())

    def _finish_stage(self, gi):
        """Stage gi's gradients are complete on the current stream: exchange them (world_size > 1) and update the
        stage's parameters, both on the side stream, while the current stream goes on with the stages below."""
        if self._stage_optimizers is None:   # CPU (gloo tests): exchange inline, one optimizer step at the end
            self.grads.launch_bucket(gi, self.world_size, self.pg)
            return
        self._side.wait_stream(torch.cuda.current_stream(self.device))
        if self.async_wgrad:   # the stage's side-stream weight gradients must be in place too; the main stream need
            self._side.wait_stream(self._wgrad_stream)   # not wait for them
        with torch.cuda.stream(self._side):
            if self.grads is not None and self.world_size > 1:
                self.grads._world, self.grads._group = self.world_size, self.pg
                self.grads._reduce_bucket(gi)
            self._stage_optimizers[gi].step()
            self._refresh_shadows(self._stage_params[gi])

    def _zero_grads(self):
        if self.grads is None:
            keep = [p.grad for p in self._wgrad_params]       # overwritten (not accumulated) by the side-stream GEMMs
            self.optimizer.zero_grad(set_to_none=True)
            for p, g in zip(self._wgrad_params, keep):
                p.grad = g
        else:
            self.grads.zero()

    def _step_body(self):
        if self._stage_groups is not None:
            return self._step_body_staged()
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            output = self.model(self.src, self.meta)
        loss = loss_fn(self.src, self.trg, output)
        self._zero_grads()
        with self._wgrad_scope():
            loss.backward()
        self._join_wgrad()
        self._allreduce()
        self.optimizer.step()
        self._refresh_shadows()
        self.loss.copy_(loss.detach())

    def capture(self, warmup=3):
        """Warm up on a side stream, capture one step, then put parameters and Adam state back to where they
        were: capturing must not train."""
        saved = [p.detach().clone() for p in self.model.parameters()]
        opts = [self.optimizer] + (self._stage_optimizers or [])
        # Adam state that exists BEFORE the warm-up (a resume: Trainer.optimizers[i].load_state_dict(...)) is put back
        # afterwards; state the warm-up itself creates is zeroed.  Keyed by (optimizer, parameter, state name).
        saved_state = {(oi, id(p), k): v.detach().clone()
                       for oi, opt in enumerate(opts) for p, st in opt.state.items()
                       for k, v in st.items() if torch.is_tensor(v)}
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step_body()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        mode = os.environ.get("MAMBA_B200_CAPTURE_MODE", "global")
        with torch.cuda.graph(self.graph, capture_error_mode=mode):
            self._step_body()
        with torch.no_grad():
            for p, q in zip(self.model.parameters(), saved):
                p.copy_(q)
            for oi, opt in enumerate(opts):
                for p, st in opt.state.items():   # exp_avg, exp_avg_sq and the device-side step counter
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            old = saved_state.get((oi, id(p), k))
                            if old is None:
                                v.zero_()
                            else:
                                v.copy_(old)
            self._refresh_shadows()   # the restore above rewrote the parameters
        torch.cuda.synchronize(self.device)

    def step(self, src, trg, meta):
        """src/trg/meta: host (pinned) or device tensors of the configured shape.  Returns the device-side
        loss scalar (read it with .item() only when you need it: train.py:171 syncs every step)."""
        self.src.copy_(src, non_blocking=True)
        self.trg.copy_(trg, non_blocking=True)
        self.meta.copy_(meta, non_blocking=True)
        if self._shadow_params:   # someone else wrote the parameters (load_state_dict, capture's restore): re-cast
            from .models.mamba.mamba import WeightShadows
            if WeightShadows.stale(self._shadow_params):
                WeightShadows.refresh(self._shadow_params)
        if self.use_graph:
            if self.graph is None:
                self.capture()
            self.graph.replay()
        else:
            self._step_body()
        return self.loss
