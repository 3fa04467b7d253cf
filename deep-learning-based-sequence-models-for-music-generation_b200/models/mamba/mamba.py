"""Drop-in mirror of the reference's Mamba modules, running on the sm_100a kernels of libmamba_b200.so.

Same class names, constructor arguments, call signatures and parameter names/shapes (hence state_dict layout) as the
reference, for both model layouts:
  * `ModelArgs`, `Mamba(params)`, `ResidualBlock`, `MambaBlock`, `RMSNorm` — the pure-PyTorch Mamba-1 of
    models/mamba/__pycache__/simple_mamba.cpython-311.pyc (source deleted upstream; SURVEY.md Appendix A;
    `@Lnnn` = original source line).  "Layout P": keys embedding / metadata_embedding /
    layers.{i}.mixer.* / layers.{i}.norm.weight / norm_f.weight / lm_head.weight (tied).
  * `Mamba(d_model=1024, n_layers=10)` — the shipped wrapper (models/mamba/mamba.py:8-35), "Layout S": keys
    token_embedding / metadata_embedding / output_layer / layers.{i}.* / norm; no residuals, final LayerNorm,
    untied head with bias.  Its layers are `mamba2.Mamba2` (mamba_ssm.Mamba2's parameters and arithmetic on this
    repo's kernels), so a checkpoint of the reference's shipped model loads; parity for that layer is UNPINNED
    (mamba_ssm is outside the reference tree — oracle/mamba2_ref.py restates the published recurrence).
Both forms take `forward(tokens[B,T] long, meta[B,6] long)` and return logits `[B, T, V]` (first 6
positions dropped, mamba.py:35 / simple_mamba @L96).

What runs where: conv1d+SiLU, softplus(dt)+scan+D-skip+z-gate, RMSNorm(+residual) and the decode step are
this repo's CUDA kernels (ops.py -> C-ABI).  in_proj / x_proj / dt_proj / out_proj / lm_head are plain
`F.linear` (cuBLAS; bf16 under `torch.autocast`), embeddings and `cat` are torch.  There is NO CPU path:
calling a module on CPU tensors raises.

Not in the reference (F3): `MambaBlock.step`, `Mamba.allocate_inference_cache / prefill / step` — the
recurrent decode that replaces the full re-forward per token of scripts/generate.py:26-31.
"""
from __future__ import annotations

import contextlib
import math
import os
import warnings
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...configs import common as cc
from ...configs import mamba as cm

__all__ = ["ModelArgs", "Mamba", "ResidualBlock", "MambaBlock", "RMSNorm"]


class _Nvtx:
    """NVTX ranges per layer and phase (tracing), opt-in with MAMBA_B200_NVTX=1: off by default
    because a range is two python calls per phase and the step is launch-sensitive outside a CUDA graph.  Forward
    ranges only (the backward's kernels carry their own names in nsys / ncu)."""
    on = os.environ.get("MAMBA_B200_NVTX", "0") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Nvtx.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _Nvtx.on:
            torch.cuda.nvtx.range_pop()


@dataclass
class ModelArgs:  # simple_mamba @L33-54
    d_model: int
    n_layer: int
    vocab_size: int
    d_state: int = 16
    expand: int = 2
    dt_rank: Union[int, str] = "auto"
    d_conv: int = 4
    pad_vocab_size_multiple: int = 8
    conv_bias: bool = True
    bias: bool = False
    metadata_vocab_size: int = cc.metadata_vocab_size

    def __post_init__(self):
        self.d_inner = int(self.expand * self.d_model)
        if self.dt_rank == "auto":
            self.dt_rank = math.ceil(self.d_model / 16)
        if self.vocab_size % self.pad_vocab_size_multiple != 0:
            self.vocab_size += self.pad_vocab_size_multiple - self.vocab_size % self.pad_vocab_size_multiple


def _act_dtype(x: torch.Tensor) -> torch.dtype:
    """dtype the mixer runs in: the autocast dtype when autocast is on (fp32 residual stream, bf16 mixer),
    else the tensor's own dtype."""
    if torch.is_autocast_enabled("cuda"):
        return torch.get_autocast_dtype("cuda")
    return x.dtype


class AsyncWgrad:
    """Weight-gradient GEMMs on a side stream.  The scan kernels put one CTA on 128 of the 148 SMs for most of the
    backward; a weight gradient is needed only by the optimizer, so `_LinearFn.backward` can write it straight into
    `param.grad` on a second stream, where it runs on the SMs the scans leave idle.

    The side path OVERWRITES param.grad and returns no gradient to autograd, so it is safe only for an owner that
    (i) pre-allocates the gradient buffers, (ii) joins the stream before anything reads them and (iii) never
    accumulates over micro-batches.  It is therefore scoped: active only inside `with AsyncWgrad.scope(stream,
    ids)` — which a Trainer opens around ITS OWN backward calls, naming ITS parameters — and off everywhere else
    (train.train_step, a second model, gradient accumulation all take the ordinary autograd path).  Works eagerly
    and under CUDA-graph capture (the fork/join becomes graph edges).  The flag is process-wide rather than
    thread-local because autograd runs CUDA backward nodes on its own device threads."""
    stream = None
    owned = frozenset()   # ids of the parameters of the scope's owner

    @classmethod
    @contextlib.contextmanager
    def scope(cls, stream, param_ids):
        prev = (cls.stream, cls.owned)
        cls.stream, cls.owned = stream, param_ids
        try:
            yield
        finally:
            cls.stream, cls.owned = prev

    @classmethod
    def disable(cls):
        cls.stream, cls.owned = None, frozenset()


class WeightShadows:
    """Autocast-dtype copies of the mixer's linear weights, kept by a Trainer.  Under autocast every forward casts
    each fp32 weight matrix (41 launches on the critical path of the step); a Trainer that owns the optimizer can
    instead refresh a persistent copy right after it has updated the parameter — on its side stream, under the
    backward.  A copy is used only while the parameter's version counter still is the one it was made from, so any
    other writer (load_state_dict, a user's in-place edit) silently falls back to the cast."""
    table = {}   # id(param) -> [shadow, version, weakref(param)]

    @classmethod
    def register(cls, params, dtype):
        """Returns the shadow tensors (the caller keeps them alive for as long as its CUDA graph may use them)."""
        import weakref
        cls.table = {k: e for k, e in cls.table.items() if e[2]() is not None}   # drop entries of dead parameters
        for p in params:
            cls.table[id(p)] = [torch.empty_like(p, dtype=dtype), -1, weakref.ref(p)]
        cls.refresh(params)
        return [cls.table[id(p)][0] for p in params]

    @classmethod
    def refresh(cls, params):
        ps = [p for p in params if id(p) in cls.table]
        if not ps:
            return
        with torch.no_grad():
            torch._foreach_copy_([cls.table[id(p)][0] for p in ps], [p.detach() for p in ps])
        for p in ps:
            cls.table[id(p)][1] = p._version

    @classmethod
    def stale(cls, params):
        return any(id(p) in cls.table and cls.table[id(p)][1] != p._version for p in params)

    @classmethod
    def get(cls, w, dtype):
        e = cls.table.get(id(w))
        if e is not None and e[2]() is w and e[1] == w._version and e[0].dtype == dtype:
            return e[0]
        return None


class _LinearFn(torch.autograd.Function):
    """y = x @ W^T in the autocast dtype (cuBLAS plumbing, no arithmetic of its own).  Unlike nn.Linear under
    autocast, the weight gradient leaves the GEMM already in the parameter's dtype (bf16 x bf16 -> fp32 output)
    instead of as a bf16 product followed by a separate widening pass over every weight matrix."""

    @staticmethod
    def forward(ctx, x, w, dt, plan=None):
        ctx.plan = plan
        wc = WeightShadows.get(w, dt) if WeightShadows.table else None
        if wc is None:
            wc = w.to(dt)
        x2 = x.reshape(-1, x.shape[-1])
        if x2.dtype != dt:
            x2 = x2.to(dt)
        out = torch.mm(x2, wc.t())
        ctx.save_for_backward(x2, wc)
        ctx.meta = (x.shape, x.dtype, w.dtype)
        ctx.param = w if isinstance(w, nn.Parameter) else None
        return out.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, g):
        x2, wc = ctx.saved_tensors
        xshape, xdt, wdt = ctx.meta
        g2 = g.reshape(-1, g.shape[-1])
        if g2.dtype != wc.dtype:
            g2 = g2.to(wc.dtype)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            plan = ctx.plan
            if plan is not None and plan.xdbl_shape is not None and xdt == wc.dtype:
                # dt_proj: d(dt_r) is the first slice of d(x_dbl); the GEMM writes it in place (ops.MixerGradPlan)
                buf = plan.buffer("xdbl", plan.xdbl_shape, xdt, g2.device)
                R = xshape[-1]
                torch.mm(g2, wc, out=buf.view(-1, buf.shape[-1])[:, :R])
                dx = buf[..., :R]
            else:
                dx = torch.mm(g2, wc).view(xshape)
                if dx.dtype != xdt:
                    dx = dx.to(xdt)
        if ctx.needs_input_grad[1]:
            side, p = AsyncWgrad.stream, ctx.param
            if (side is not None and p is not None and id(p) in AsyncWgrad.owned and p.grad is not None
                    and p.grad.dtype == wdt and p.grad.is_contiguous() and wdt != g2.dtype):
                # straight into param.grad on the side stream (this layer is the parameter's only user)
                cur = torch.cuda.current_stream(g2.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    torch.mm(g2.t(), x2, out_dtype=wdt, out=p.grad)
                g2.record_stream(side), x2.record_stream(side)
            else:
                dw = torch.mm(g2.t(), x2) if wdt == g2.dtype else torch.mm(g2.t(), x2, out_dtype=wdt)
        return dx, dw, None, None


def _linear(x, weight, bias=None, plan=None):
    """F.linear; on CUDA under autocast (and without a bias) through _LinearFn."""
    if bias is None and x.is_cuda and torch.is_autocast_enabled("cuda"):
        return _LinearFn.apply(x, weight, torch.get_autocast_dtype("cuda"), plan)
    return F.linear(x, weight, bias)


def _linear_step(x, lin):
    """nn.Linear for one position per sequence: the weight-streaming kernel for <= 16 rows, cuBLAS otherwise."""
    if x.shape[0] <= 16 and x.shape[1] % 4 == 0 and lin.weight.is_contiguous():
        return ops.linear_step(x.contiguous() if x.stride(-1) != 1 else x, lin.weight.detach(),
                               None if lin.bias is None else lin.bias.detach())
    return lin(x)


class RMSNorm(nn.Module):  # simple_mamba @L336-348
    def __init__(self, d_model: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d_model))

    def forward(self, x, residual=None):
        """`forward(x)` is the reference call (@L346).  `forward(x, residual)` is the fused form used inside
        Mamba.forward: returns (rmsnorm(x + residual), x + residual)."""
        if residual is None:
            return ops.rmsnorm_fn(x, self.weight, None, self.eps)[0]
        return ops.rmsnorm_fn(x, self.weight, residual, self.eps, _act_dtype(residual))


class MambaBlock(nn.Module):  # simple_mamba @L184
    def __init__(self, params, layer_idx=None):  # @L185-211
        super().__init__()
        self.params = params
        self.layer_idx = layer_idx
        self.in_proj = nn.Linear(params.d_model, params.d_inner * 2, bias=params.bias)
        self.conv1d = nn.Conv1d(in_channels=params.d_inner, out_channels=params.d_inner, bias=params.conv_bias,
                                kernel_size=params.d_conv, groups=params.d_inner, padding=params.d_conv - 1)
        self.x_proj = nn.Linear(params.d_inner, params.dt_rank + params.d_state * 2, bias=False)
        self.dt_proj = nn.Linear(params.dt_rank, params.d_inner, bias=True)
        A = torch.arange(1, params.d_state + 1, dtype=torch.float32).repeat(params.d_inner, 1)
        self.A_log = nn.Parameter(torch.log(A))
        self.D = nn.Parameter(torch.ones(params.d_inner))
        self.out_proj = nn.Linear(params.d_inner, params.d_model, bias=params.bias)

    # ---- training / full-sequence forward (@L228-245 with ssm @L263-280 and selective_scan @L310-333) --
    def forward(self, x):
        p = self.params
        xz = _linear(x, self.in_proj.weight, self.in_proj.bias)           # @L230  [B, L, 2*d_inner]
        # training on CUDA under autocast: the gradients of the split() views are written in place into one buffer
        # per split instead of being concatenated (ops.MixerGradPlan)
        plan = None
        if torch.is_grad_enabled() and xz.requires_grad and xz.is_cuda and torch.is_autocast_enabled("cuda"):
            plan = ops.MixerGradPlan(xz.shape, xz.shape[:-1] + (p.dt_rank + 2 * p.d_state,))
        xs, res = ops.split_fn(xz, [p.d_inner, p.d_inner], plan, "xz")    # @L231  views, no copy
        with _Nvtx("conv1d_silu"):
            xc = ops.causal_conv1d_silu_fn(xs, self.conv1d.weight, self.conv1d.bias, plan=plan)   # @L233-237 (one kernel)
        with _Nvtx("x_proj_dt_proj"):
            x_dbl = _linear(xc, self.x_proj.weight)                           # @L273
            dt_r, Bm, Cm = ops.split_fn(x_dbl, [p.dt_rank, p.d_state, p.d_state], plan, "xdbl")  # @L275 views
            dt_raw = _linear(dt_r, self.dt_proj.weight, plan=plan)            # @L276: bias + softplus fused below
        # A = -exp(A_log) (@L270) is formed inside the kernels; the gradient comes back w.r.t. A_log
        with _Nvtx("selective_scan"):
            y = ops.selective_scan_fn(xc, dt_raw, self.A_log, Bm, Cm, self.D.float(), z=res,
                                      delta_bias=self.dt_proj.bias.float(), delta_softplus=True,
                                      A_is_log=True, plan=plan)               # @L270, @L276-278, @L241
        with _Nvtx("out_proj"):
            return _linear(y, self.out_proj.weight, self.out_proj.bias)       # @L243

    # ---- inference: full-sequence forward that also leaves the recurrent state behind ---------------
    @torch.no_grad()
    def prefill(self, x, conv_state, ssm_state):
        p = self.params
        xz = self.in_proj(x)
        xs, res = xz.split([p.d_inner, p.d_inner], dim=-1)
        xc, cs = ops.causal_conv1d_silu_prefill(xs, self.conv1d.weight, self.conv1d.bias)
        conv_state.copy_(cs)
        A = -torch.exp(self.A_log.float())
        x_dbl = self.x_proj(xc)
        dt_r, Bm, Cm = x_dbl.split([p.dt_rank, p.d_state, p.d_state], dim=-1)
        dt_raw = F.linear(dt_r, self.dt_proj.weight)
        y, h_last = ops.selective_scan_prefill(xc, dt_raw, A, Bm, Cm, self.D.float(), z=res,
                                               delta_bias=self.dt_proj.bias.float(), delta_softplus=True)
        ssm_state.copy_(h_last)
        return self.out_proj(y)

    def allocate_inference_cache(self, batch_size, max_seqlen=None, dtype=None, device=None):
        """(conv_state [B, d_inner, d_conv], ssm_state [B, d_inner, d_state] fp32), zero-initialised —
        the two tensors mamba_ssm's `allocate_inference_cache` hands out (used in the reference's
        scripts/test_inference.ipynb:137)."""
        p = self.params
        device = device or self.in_proj.weight.device
        dtype = dtype or self.in_proj.weight.dtype
        conv_state = torch.zeros(batch_size, p.d_inner, p.d_conv, device=device, dtype=dtype)
        ssm_state = torch.zeros(batch_size, p.d_inner, p.d_state, device=device, dtype=torch.float32)
        return conv_state, ssm_state

    @torch.no_grad()
    def step_constants(self):
        """Per-layer tensors the decode step needs in kernel form (fp32, contiguous): built once per decode run
        instead of once per token — A = -exp(A_log) alone is three launches."""
        p = self.params
        return dict(
            conv_w=self.conv1d.weight.detach().float().reshape(p.d_inner, p.d_conv).contiguous(),
            conv_b=None if self.conv1d.bias is None else self.conv1d.bias.detach().float().contiguous(),
            dt_w=self.dt_proj.weight.detach().float().contiguous(),
            dt_b=self.dt_proj.bias.detach().float().contiguous(),
            A=(-torch.exp(self.A_log.detach().float())).contiguous(),
            D=self.D.detach().float().contiguous())

    @torch.no_grad()
    def step(self, x_t, conv_state, ssm_state, consts=None):
        """One new position: x_t [B, d_model] -> [B, d_model]; both states are updated in place."""
        p = self.params
        k = consts if consts is not None else self.step_constants()
        T = conv_state.dtype                                              # the step's activation dtype
        if x_t.dtype != T:
            x_t = x_t.to(T)
        xz = _linear_step(x_t, self.in_proj)                              # [B, 2*d_inner]
        xs, res = xz.split([p.d_inner, p.d_inner], dim=-1)
        xc = ops.conv_step(xs, conv_state, k["conv_w"], k["conv_b"])
        x_dbl = _linear_step(xc, self.x_proj)
        dt_r, Bv, Cv = x_dbl.split([p.dt_rank, p.d_state, p.d_state], dim=-1)
        y = ops.ssm_step(xc, dt_r, Bv, Cv, k["dt_w"], k["dt_b"], k["A"], k["D"], res, ssm_state)
        return _linear_step(y, self.out_proj)


    @torch.no_grad()
    def step_fused(self, hidden, resid_in, resid_out, norm, conv_state, ssm_state, k):
        """One new position through `norm -> mixer` of the residual block in FOUR launches (decode, <= 16 sequences):
        in_proj with the block's RMSNorm + residual add as its prologue and the conv step as its epilogue, x_proj,
        the SSM step (dt_proj + softplus + recurrence + D skip + gate), out_proj.
        hidden: previous mixer output [B, d_model] or None (first layer); resid_in / resid_out: fp32 residual stream
        (ping-pong buffers).  Returns the mixer output [B, d_model]."""
        p = self.params
        xz, xc = ops.fused_linear_step(hidden, self.in_proj.weight.detach(), None if self.in_proj.bias is None else self.in_proj.bias.detach(),
                                       norm_weight=k["norm_w"], eps=norm.eps, residual_in=resid_in, residual_out=resid_out,
                                       conv_state=conv_state, conv_weight=k["conv_w"], conv_bias=k["conv_b"])
        res = xz[:, p.d_inner:]
        x_dbl = ops.linear_step(xc, self.x_proj.weight.detach(), None)
        dt_r, Bv, Cv = x_dbl.split([p.dt_rank, p.d_state, p.d_state], dim=-1)
        y = ops.ssm_step(xc, dt_r, Bv, Cv, k["dt_w"], k["dt_b"], k["A"], k["D"], res, ssm_state)
        return ops.linear_step(y, self.out_proj.weight.detach(), None if self.out_proj.bias is None else self.out_proj.bias.detach())


class ResidualBlock(nn.Module):  # simple_mamba @L151-181
    def __init__(self, params, layer_idx=None):
        super().__init__()
        self.params = params
        self.mixer = MambaBlock(params, layer_idx)
        self.norm = RMSNorm(params.d_model)

    def forward(self, x):  # @L179
        return self.mixer(self.norm(x)) + x


def _params_from_configs(d_model=None, n_layers=None, vocab_size=None, pad=False):
    mv = cm.config.model_values
    d_model = mv.d_model if d_model is None else d_model
    return ModelArgs(d_model=d_model, n_layer=mv.n_layer if n_layers is None else n_layers,
                     vocab_size=cc.vocab_size if vocab_size is None else vocab_size, d_state=mv.d_state,
                     expand=mv.expand, d_conv=mv.d_conv, conv_bias=mv.conv_bias, bias=mv.bias,
                     pad_vocab_size_multiple=mv.pad_vocab_size_multiple if pad else 1,
                     metadata_vocab_size=cc.metadata_vocab_size)


class InferenceCache(list):
    """Per-layer (conv_state, ssm_state) pairs; also carries the per-layer step constants once they are built."""
    consts = None


class _PaddedHeadFn(torch.autograd.Function):
    """logits = x @ W^T computed as x @ pad8(W)^T (cuBLAS plumbing, no arithmetic of its own).  The gradient that
    comes back from ops.filtered_ce_fn already lives in a zero-padded buffer of the same layout and is used in
    place; any other gradient is padded first."""

    @staticmethod
    def forward(ctx, x, w, dt):
        V, d = w.shape
        Vp = V + (-V) % 8
        wp = torch.zeros((Vp, d), dtype=dt, device=w.device)
        wp[:V].copy_(w)
        x2 = x.to(dt).reshape(-1, d)
        out = torch.mm(x2, wp.t())                       # [rows, Vp]
        ctx.save_for_backward(x2, wp)
        ctx.meta = (x.shape, x.dtype, w.dtype, V, Vp)
        return out.view(*x.shape[:-1], Vp)[..., :V]

    @staticmethod
    def backward(ctx, g):
        x2, wp = ctx.saved_tensors
        xshape, xdt, wdt, V, Vp = ctx.meta
        rows = x2.shape[0]
        g3 = g.reshape(-1, g.shape[-2], V) if g.dim() > 2 else g.reshape(1, -1, V)
        if (g.dtype == wp.dtype and g3.stride(2) == 1 and g3.stride(1) == Vp and g3.stride(0) == g3.shape[1] * Vp
                and g.storage_offset() == 0 and g.untyped_storage().nbytes() >= rows * Vp * g.element_size()):
            gp = g.as_strided((rows, Vp), (Vp, 1))       # filtered_ce_fn's padded gradient, pad columns are zero
        else:
            gp = torch.zeros((rows, Vp), dtype=wp.dtype, device=g.device)
            gp[:, :V].copy_(g.reshape(rows, V))
        dx = torch.mm(gp, wp).view(xshape).to(xdt)
        dw = (torch.mm(gp.t(), x2) if wdt == gp.dtype else torch.mm(gp.t(), x2, out_dtype=wdt))[:V]
        return dx, dw, None


class Mamba(nn.Module):
    """`Mamba(params)` -> Layout P (simple_mamba @L57-96);  `Mamba()` / `Mamba(d_model=1024, n_layers=10)` ->
    the shipped wrapper's layout (models/mamba/mamba.py:8-35)."""

    def __init__(self, d_model: Union[int, SimpleNamespace, ModelArgs] = 1024, n_layers: int = 10):
        super().__init__()
        if isinstance(d_model, int):
            self.layout = "S"
            params = _params_from_configs(d_model, n_layers)
            self.params = params
            self.token_embedding = nn.Embedding(cc.vocab_size, d_model)            # mamba.py:12
            self.metadata_embedding = nn.Embedding(cc.metadata_vocab_size, d_model)  # :13
            self.output_layer = nn.Linear(d_model, cc.vocab_size)                  # :14
            from .mamba2 import Mamba2
            mv = cm.config.model_values
            self.layers = nn.ModuleList([Mamba2(d_model=d_model, d_state=mv.d_state, d_conv=mv.d_conv, expand=mv.expand,
                                                layer_idx=i) for i in range(n_layers)])            # :16-24
            self.norm = nn.LayerNorm(d_model)                                      # :25
        else:
            self.layout = "P"
            params = d_model
            self.params = params
            self.vocab_size = params.vocab_size
            self.metadata_vocab_size = params.metadata_vocab_size
            self.embedding = nn.Embedding(params.vocab_size, params.d_model)       # @L62
            self.metadata_embedding = nn.Embedding(params.metadata_vocab_size, params.d_model)
            self.layers = nn.ModuleList([ResidualBlock(params, layer_idx=i) for i in range(params.n_layer)])
            self.norm_f = RMSNorm(params.d_model)
            self.lm_head = nn.Linear(params.d_model, params.vocab_size, bias=False)
            self.lm_head.weight = self.embedding.weight                            # tied, @L70

    def get_name(self):  # @L147
        return "Mamba"

    # ---- embeddings (both layouts): cat((meta_emb, token_emb), dim=-2) --------------------------------
    def _embed(self, tokens, meta):
        tok = self.embedding if self.layout == "P" else self.token_embedding
        return torch.cat((self.metadata_embedding(meta), tok(tokens)), dim=-2)

    def forward(self, tokens, meta):
        x = self._embed(tokens, meta)
        n_meta = meta.shape[-1]
        if self.layout == "S":  # mamba.py:32-35
            for layer in self.layers:
                x = layer(x)
            x = self.norm(x)
            return self.output_layer(x[:, n_meta:])
        # Layout P, @L90-96.  `mixer(norm(x)) + x` per layer, with each `+ x` fused into the next RMSNorm:
        # the residual stream is read and written once per layer.
        resid, hidden = x, None
        for i, layer in enumerate(self.layers):
            with _Nvtx(f"layer{i}"):
                with _Nvtx("rmsnorm_residual"):
                    normed, resid = layer.norm(hidden, resid)
                hidden = layer.mixer(normed)
        with _Nvtx("norm_f_head"):
            normed, _ = self.norm_f(hidden, resid)
            return self._head(normed[:, n_meta:])  # rows are independent: slicing before the GEMM == logits[:, 6:]

    def _head(self, x):
        """lm_head (layout P) with the vocabulary padded to a multiple of 8 INSIDE the GEMM: 17914 columns make
        every row of the logits 4-byte aligned only, which pushes cuBLAS onto its slow align2 kernels.  The
        returned tensor is the [.., :vocab] view of the padded product, so values and shape are unchanged."""
        w = self.lm_head.weight
        if w.shape[0] % 8 == 0 or not x.is_cuda:
            return self.lm_head(x)
        dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else w.dtype
        return _PaddedHeadFn.apply(x, w, dt)

    # ---- recurrent decode (no reference counterpart) -------------------------
    def allocate_inference_cache(self, batch_size, max_seqlen=None, dtype=None):
        mixers = [l.mixer if self.layout == "P" else l for l in self.layers]
        return InferenceCache(m.allocate_inference_cache(batch_size, max_seqlen, dtype) for m in mixers)

    @torch.no_grad()
    def prefill(self, tokens, meta, cache):
        """Full forward over the prompt that leaves (conv_state, ssm_state) of every layer in `cache`.
        Returns logits [B, T, V] exactly as forward()."""
        x = self._embed(tokens, meta)
        n_meta = meta.shape[-1]
        if self.layout == "S":
            for layer, (cs, hs) in zip(self.layers, cache):
                x = layer.prefill(x, cs, hs)
            return self.output_layer(self.norm(x)[:, n_meta:])
        resid, hidden = x, None
        for layer, (cs, hs) in zip(self.layers, cache):
            normed, resid = layer.norm(hidden, resid)
            hidden = layer.mixer.prefill(normed, cs, hs)
        normed, _ = self.norm_f(hidden, resid)
        return self._head(normed[:, n_meta:])

    @torch.no_grad()
    def step(self, token, cache, logits_out=None):
        """token [B] long -> logits [B, V]; advances every layer's state by one position.  Layout P on CUDA with
        <= 16 sequences takes the fused path (4 launches per layer, norm and conv folded into in_proj; the final
        norm folded into the head); `logits_out` (fp32 [B, V]) receives the logits when given."""
        tok = self.embedding if self.layout == "P" else self.token_embedding
        mixers = [l.mixer if self.layout == "P" else l for l in self.layers]
        consts = getattr(cache, "consts", None)
        if consts is None:
            consts = [m.step_constants() for m in mixers]
            try:
                cache.consts = consts
            except AttributeError:  # a plain list was passed: rebuilt per call
                pass
        if (self.layout == "P" and token.is_cuda and token.shape[0] <= 16 and self.params.d_model % 4 == 0
                and getattr(cache, "fused", True)):
            return self._step_fused(tok, token, mixers, cache, consts, logits_out)
        x = tok(token)
        if self.layout == "S":
            for layer, (cs, hs), k in zip(self.layers, cache, consts):
                x = layer.step(x, cs, hs, k)
            return _linear_step(self.norm(x), self.output_layer)
        resid, hidden = x, None
        for layer, (cs, hs), k in zip(self.layers, cache, consts):
            normed, resid = layer.norm(hidden, resid)
            hidden = layer.mixer.step(normed, cs, hs, k)
        normed, _ = self.norm_f(hidden, resid)
        out = _linear_step(normed, self.lm_head)
        if logits_out is not None:
            logits_out.copy_(out)
            return logits_out
        return out

    def _step_fused(self, tok, token, mixers, cache, consts, logits_out):
        B = token.shape[0]
        dev = token.device
        for layer, k in zip(self.layers, consts):
            if "norm_w" not in k:
                k["norm_w"] = layer.norm.weight.detach().float().contiguous()
        resid = tok(token).float()                       # residual stream, fp32 (as Mamba.forward keeps it)
        spare = torch.empty_like(resid)
        hidden = None
        for layer, (cs, hs), k in zip(self.layers, cache, consts):
            hidden = layer.mixer.step_fused(hidden, resid, spare, layer.norm, cs, hs, k)
            resid, spare = spare, resid
        w = self.lm_head.weight.detach()
        out = ops.fused_linear_step(hidden, w, None if self.lm_head.bias is None else self.lm_head.bias.detach(),
                                    norm_weight=self.norm_f.weight.detach().float(), eps=self.norm_f.eps, residual_in=resid,
                                    out=logits_out if (logits_out is not None and logits_out.dtype == hidden.dtype) else None)
        if logits_out is not None and out is not logits_out:
            logits_out.copy_(out)
            return logits_out
        return out
