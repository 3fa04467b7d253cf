"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's pure-PyTorch Mamba-1.

What it restates
----------------
The reference's own pure-PyTorch Mamba lives only as orphaned CPython-3.11 bytecode,
`/root/reference/models/mamba/__pycache__/simple_mamba.cpython-311.pyc` (its .py was deleted upstream;
SURVEY.md F2, transcription in Appendix A).  `@Lnnn` below is the ORIGINAL source line recorded in that
code object's line table.  The outer wrapper of the shipped model is `models/mamba/mamba.py:8-35`.

Pinning
-------
The reference holds no golden vectors, tests or fixtures and cannot be imported under
Python 3.12.  `oracle/pyc311.py` therefore EXECUTES the reference's own bytecode (selective_scan, ssm,
MambaBlock.forward, RMSNorm.forward, ResidualBlock.forward, Mamba.forward) with a small CPython-3.11
bytecode interpreter, `tests/golden/make_golden.py` stores its outputs as fixtures, and
`tests/test_oracle.py` checks this restatement against those fixtures bit for bit.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package never does.

Two scan forms:
  * scan_impl="literal": the loop exactly as written at @L310-333 (indexing deltaA[:, i]); its autograd
    backward is O(L^2) on CPU, so it is used only at small L.
  * scan_impl="unbind" : same arithmetic, but the per-step slices come from `unbind(1)`; bit-identical in
    the forward and in every gradient (tests/test_oracle.py proves it) and linear in L.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import einsum, rearrange, repeat


@dataclass
class ModelArgs:  # @L33-54
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
    metadata_vocab_size: int = 568  # supplied by train.get_mamba_dict (train.py:35)

    def __post_init__(self):  # @L46-54
        self.d_inner = int(self.expand * self.d_model)
        if self.dt_rank == "auto":
            self.dt_rank = math.ceil(self.d_model / 16)
        if self.vocab_size % self.pad_vocab_size_multiple != 0:
            self.vocab_size += self.pad_vocab_size_multiple - self.vocab_size % self.pad_vocab_size_multiple


class RMSNorm(nn.Module):  # @L336-348
    def __init__(self, d_model: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d_model))

    def forward(self, x):  # @L346
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.eps) * self.weight


class MambaBlock(nn.Module):  # @L184
    def __init__(self, params, scan_impl: str = "unbind"):  # @L185-211
        super().__init__()
        self.params = params
        self.scan_impl = scan_impl
        self.in_proj = nn.Linear(params.d_model, params.d_inner * 2, bias=params.bias)
        self.conv1d = nn.Conv1d(in_channels=params.d_inner, out_channels=params.d_inner, bias=params.conv_bias,
                                kernel_size=params.d_conv, groups=params.d_inner, padding=params.d_conv - 1)
        self.x_proj = nn.Linear(params.d_inner, params.dt_rank + params.d_state * 2, bias=False)
        self.dt_proj = nn.Linear(params.dt_rank, params.d_inner, bias=True)
        A = repeat(torch.arange(1, params.d_state + 1), "n -> d n", d=params.d_inner)
        self.A_log = nn.Parameter(torch.log(A))
        self.D = nn.Parameter(torch.ones(params.d_inner))
        self.out_proj = nn.Linear(params.d_inner, params.d_model, bias=params.bias)

    def forward(self, x):  # @L228-245
        (b, l, d) = x.shape
        x_and_res = self.in_proj(x)
        (x, res) = x_and_res.split(split_size=[self.params.d_inner, self.params.d_inner], dim=-1)
        x = rearrange(x, "b l d_in -> b d_in l")
        x = self.conv1d(x)[:, :, :l]
        x = rearrange(x, "b d_in l -> b l d_in")
        x = F.silu(x)
        y = self.ssm(x)
        y = y * F.silu(res)
        return self.out_proj(y)

    def ssm(self, x):  # @L263-280
        (d_in, n) = self.A_log.shape
        A = -torch.exp(self.A_log.float())
        D = self.D.float()
        x_dbl = self.x_proj(x)
        (delta, B, C) = x_dbl.split(split_size=[self.params.dt_rank, n, n], dim=-1)
        delta = F.softplus(self.dt_proj(delta))
        return self.selective_scan(x, delta, A, B, C, D)

    def selective_scan(self, u, delta, A, B, C, D):  # @L310-333
        return selective_scan(u, delta, A, B, C, D, impl=self.scan_impl)

    # -- not in the reference: single-token recurrence derived from forward()/ssm()/selective_scan() above
    #    (SURVEY.md F3 / §8 row A9).  conv_state [b, d_in, d_conv] holds the last d_conv inputs, ssm_state
    #    [b, d_in, n] the state the loop at @L325-328 calls `x`.
    def step(self, x_t, conv_state, ssm_state):
        xz = self.in_proj(x_t)  # [b, 2*d_in]
        (x, res) = xz.split(split_size=[self.params.d_inner, self.params.d_inner], dim=-1)
        conv_state = torch.cat((conv_state[:, :, 1:], x.unsqueeze(-1)), dim=-1)
        x = (conv_state * self.conv1d.weight[:, 0, :]).sum(-1)
        if self.conv1d.bias is not None:
            x = x + self.conv1d.bias
        x = F.silu(x)
        (d_in, n) = self.A_log.shape
        A = -torch.exp(self.A_log.float())
        x_dbl = self.x_proj(x)
        (delta, B, C) = x_dbl.split(split_size=[self.params.dt_rank, n, n], dim=-1)
        delta = F.softplus(self.dt_proj(delta))
        deltaA = torch.exp(einsum(delta, A, "b d_in, d_in n -> b d_in n"))
        deltaB_u = einsum(delta, B, x, "b d_in, b n, b d_in -> b d_in n")
        ssm_state = deltaA * ssm_state + deltaB_u
        y = einsum(ssm_state, C, "b d_in n, b n -> b d_in")
        y = y + x * self.D.float()
        y = y * F.silu(res)
        return self.out_proj(y), conv_state, ssm_state


def selective_scan(u, delta, A, B, C, D, impl: str = "unbind", return_last_state: bool = False):
    """MambaBlock.selective_scan @L310-333.  u, delta [b, l, d_in]; A [d_in, n]; B, C [b, l, n]; D [d_in]."""
    (b, l, d_in) = u.shape
    n = A.shape[1]
    deltaA = torch.exp(einsum(delta, A, "b l d_in, d_in n -> b l d_in n"))  # @L320
    deltaB_u = einsum(delta, B, u, "b l d_in, b l n, b l d_in -> b l d_in n")  # @L321
    x = torch.zeros((b, d_in, n), device=deltaA.device)  # @L324
    ys = []
    if impl == "literal":
        for i in range(l):  # @L325-328
            x = deltaA[:, i] * x + deltaB_u[:, i]
            y = einsum(x, C[:, i, :], "b d_in n, b n -> b d_in")
            ys.append(y)
    elif impl == "unbind":
        for (dA_i, dBu_i, C_i) in zip(deltaA.unbind(1), deltaB_u.unbind(1), C.unbind(1)):
            x = dA_i * x + dBu_i
            y = einsum(x, C_i, "b d_in n, b n -> b d_in")
            ys.append(y)
    else:
        raise ValueError(impl)
    y = torch.stack(ys, dim=1)  # @L329
    y = y + u * D  # @L331
    if return_last_state:
        return y, x
    return y


class ResidualBlock(nn.Module):  # @L151-181
    def __init__(self, params, scan_impl: str = "unbind"):
        super().__init__()
        self.params = params
        self.mixer = MambaBlock(params, scan_impl)
        self.norm = RMSNorm(params.d_model)

    def forward(self, x):  # @L179
        return self.mixer(self.norm(x)) + x


class Mamba(nn.Module):  # @L57-96  ("Layout P")
    def __init__(self, params, scan_impl: str = "unbind"):  # @L58-70
        super().__init__()
        self.params = params
        self.vocab_size = params.vocab_size
        self.metadata_vocab_size = params.metadata_vocab_size
        self.embedding = nn.Embedding(params.vocab_size, params.d_model)
        self.metadata_embedding = nn.Embedding(params.metadata_vocab_size, params.d_model)
        self.layers = nn.ModuleList([ResidualBlock(params, scan_impl) for _ in range(params.n_layer)])
        self.norm_f = RMSNorm(params.d_model)
        self.lm_head = nn.Linear(params.d_model, params.vocab_size, bias=False)
        self.lm_head.weight = self.embedding.weight  # tied, @L70

    def forward(self, input_ids, metadata_ids, checkpoint_layers: bool = False):  # @L74-96
        token_emb = self.embedding(input_ids)
        meta_emb = self.metadata_embedding(metadata_ids)
        x = torch.cat((meta_emb, token_emb), dim=-2)
        for layer in self.layers:
            if checkpoint_layers:  # memory only: same arithmetic, recomputed in backward
                from torch.utils.checkpoint import checkpoint
                x = checkpoint(layer, x, use_reentrant=False)
            else:
                x = layer(x)
        x = self.norm_f(x)
        logits = self.lm_head(x)
        return logits[:, 6:]


class ShippedWrapper(nn.Module):
    """Outer wrapper of the SHIPPED model, models/mamba/mamba.py:8-35 ("Layout S" outer keys), with the
    pure-PyTorch MambaBlock above standing in for the external mamba_ssm.Mamba2 layer (:17-23), which is
    not in the reference tree (parity for Mamba2 itself is unpinned)."""

    def __init__(self, params, d_model: int = 1024, n_layers: int = 10, vocab_size: int = 17914,
                 metadata_vocab_size: int = 568, scan_impl: str = "unbind"):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab_size, d_model)  # :12
        self.metadata_embedding = nn.Embedding(metadata_vocab_size, d_model)  # :13
        self.output_layer = nn.Linear(d_model, vocab_size)  # :14
        self.layers = nn.ModuleList([MambaBlock(params, scan_impl) for _ in range(n_layers)])  # :16-24
        self.norm = nn.LayerNorm(d_model)  # :25

    def forward(self, tokens, meta):  # :27-35
        x = self.token_embedding(tokens)
        x = torch.cat((self.metadata_embedding(meta), x), dim=-2)
        for layer in self.layers:
            x = layer(x)
        x = self.norm(x)
        return self.output_layer(x)[:, 6:]
