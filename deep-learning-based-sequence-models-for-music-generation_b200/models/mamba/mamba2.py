"""`Mamba2` — the layer the reference's SHIPPED model stacks (models/mamba/mamba.py:16-24: mamba_ssm.Mamba2 with
d_state 64, d_conv 4, expand 2; library defaults headdim 64, ngroups 1, gated RMSNorm, bias False, conv_bias True),
on this repo's sm_100a kernels.  Same constructor arguments, parameter names and shapes as mamba_ssm.Mamba2, so a
state_dict of the reference's shipped model loads (in_proj.weight [2*d_inner + 2*d_state + nheads, d_model],
conv1d.weight [d_inner + 2*d_state, 1, d_conv], conv1d.bias, dt_bias / A_log / D [nheads], norm.weight [d_inner],
out_proj.weight [d_model, d_inner]; parameter count of the wrapper 101,972,666 as the
reference prints at scripts/Test Accuracy.ipynb:52).

How it runs.  Mamba-2's recurrence  H_t = exp(dt_t A_h) H_{t-1} + dt_t x_t (x) B_t,  y_t = H_t C_t + D_h x_t  (one
scalar decay per head h, state [headdim, d_state] per head) is the Mamba-1 selective scan with the head's dt, A, D
broadcast over its headdim channels: channel d of head h has delta[t, d] = dt[t, h], A[d, :] = A_h, D[d] = D_h, and
B_t / C_t shared by all channels (ngroups == 1).  So the layer is: in_proj (cuBLAS) -> split -> causal conv + SiLU over
x|B|C (mamba_conv1d_silu) -> mamba_scan_fwd / mamba_scan_bwd with the broadcast operands and the z gate fused ->
mamba_rmsnorm (the gated norm: gate first, norm_before_gate = False) -> out_proj.  This is the hot path's own kernels
reused, not the SSD block decomposition: it evaluates one exp per (t, channel, state) where SSD needs one per
(t, head), i.e. it is correct and checkpoint-compatible but not the fast formulation of this layer.
Parity is UNPINNED (mamba_ssm is outside the reference tree): the oracle is oracle/mamba2_ref.py, a restatement of the
published recurrence.  ngroups > 1 is not supported (the reference uses the default 1).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops

__all__ = ["Mamba2"]


class _GatedNormWeight(nn.Module):
    """Holds `norm.weight` under the key mamba_ssm's RMSNormGated uses."""

    def __init__(self, d, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d))


class Mamba2(nn.Module):
    def __init__(self, d_model, d_state=64, d_conv=4, expand=2, headdim=64, ngroups=1, dt_min=0.001, dt_max=0.1,
                 dt_init_floor=1e-4, A_init_range=(1, 16), bias=False, conv_bias=True, layer_idx=None, **unused):
        super().__init__()
        if ngroups != 1:
            raise NotImplementedError("Mamba2: ngroups must be 1 (the reference's configuration)")
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = expand * d_model
        if self.d_inner % headdim:
            raise ValueError("Mamba2: expand * d_model must be a multiple of headdim")
        self.headdim, self.ngroups, self.nheads = headdim, ngroups, self.d_inner // headdim
        self.layer_idx = layer_idx
        self.in_proj = nn.Linear(d_model, 2 * self.d_inner + 2 * d_state + self.nheads, bias=bias)
        conv_dim = self.d_inner + 2 * d_state
        self.conv1d = nn.Conv1d(conv_dim, conv_dim, bias=conv_bias, kernel_size=d_conv, groups=conv_dim, padding=d_conv - 1)
        dt = torch.exp(torch.rand(self.nheads) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
        dt = torch.clamp(dt, min=dt_init_floor)
        self.dt_bias = nn.Parameter(dt + torch.log(-torch.expm1(-dt)))
        self.dt_bias._no_weight_decay = True
        self.A_log = nn.Parameter(torch.log(torch.empty(self.nheads).uniform_(*A_init_range)))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(self.nheads))
        self.D._no_weight_decay = True
        self.norm = _GatedNormWeight(self.d_inner)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias)

    # per-head parameters broadcast over the head's channels (autograd sums the gradients back)
    def _per_channel(self):
        P = self.headdim
        A = (-torch.exp(self.A_log.float())).repeat_interleave(P)[:, None].expand(self.d_inner, self.d_state).contiguous()
        return A, self.D.float().repeat_interleave(P), self.dt_bias.float().repeat_interleave(P)

    def forward(self, u):
        from .mamba import _linear
        N = self.d_state
        zxbcdt = _linear(u, self.in_proj.weight, self.in_proj.bias)
        z, xBC, dt = torch.split(zxbcdt, [self.d_inner, self.d_inner + 2 * N, self.nheads], dim=-1)
        xBC = ops.causal_conv1d_silu_fn(xBC, self.conv1d.weight, self.conv1d.bias)
        x, Bm, Cm = torch.split(xBC, [self.d_inner, N, N], dim=-1)
        A, Dc, bias = self._per_channel()
        delta = dt.repeat_interleave(self.headdim, dim=-1)                      # [B, L, d_inner]: the head's dt per channel
        y = ops.selective_scan_fn(x, delta, A, Bm, Cm, Dc, z=z, delta_bias=bias, delta_softplus=True)  # (y + D x) * silu(z)
        y = ops.rmsnorm_fn(y, self.norm.weight, None, self.norm.eps)[0]           # gated norm: gate first, then RMSNorm
        return _linear(y, self.out_proj.weight, self.out_proj.bias)

    # ---- inference: state-carrying prefill and one-token step (no counterpart in the reference) --------
    def allocate_inference_cache(self, batch_size, max_seqlen=None, dtype=None, device=None):
        """(conv_state [B, d_inner + 2*d_state, d_conv], ssm_state [B, d_inner, d_state] fp32 — mamba_ssm's
        [B, nheads, headdim, d_state] flattened over (nheads, headdim))."""
        device = device or self.in_proj.weight.device
        dtype = dtype or self.in_proj.weight.dtype
        conv_state = torch.zeros(batch_size, self.d_inner + 2 * self.d_state, self.d_conv, device=device, dtype=dtype)
        ssm_state = torch.zeros(batch_size, self.d_inner, self.d_state, device=device, dtype=torch.float32)
        return conv_state, ssm_state

    @torch.no_grad()
    def prefill(self, u, conv_state, ssm_state):
        N = self.d_state
        zxbcdt = self.in_proj(u)
        z, xBC, dt = torch.split(zxbcdt, [self.d_inner, self.d_inner + 2 * N, self.nheads], dim=-1)
        xBC, cs = ops.causal_conv1d_silu_prefill(xBC, self.conv1d.weight, self.conv1d.bias)
        conv_state.copy_(cs)
        x, Bm, Cm = torch.split(xBC, [self.d_inner, N, N], dim=-1)
        A, Dc, bias = self._per_channel()
        y, h_last = ops.selective_scan_prefill(x, dt.repeat_interleave(self.headdim, dim=-1), A, Bm, Cm, Dc, z=z,
                                               delta_bias=bias, delta_softplus=True)
        ssm_state.copy_(h_last)
        y = ops.rmsnorm_fn(y, self.norm.weight, None, self.norm.eps)[0]
        return self.out_proj(y)

    @torch.no_grad()
    def step_constants(self):
        """Kernel-form constants of the decode step.  The SSM step kernel fuses a dt projection
        (delta = softplus(W_dt . dt_in + b)): here W_dt is the 0/1 matrix that hands every channel its head's dt."""
        P, dev = self.headdim, self.A_log.device
        A, Dc, bias = self._per_channel()
        sel = torch.zeros(self.d_inner, self.nheads, device=dev)
        sel[torch.arange(self.d_inner, device=dev), torch.arange(self.d_inner, device=dev) // P] = 1.0
        cd = self.d_inner + 2 * self.d_state
        return dict(conv_w=self.conv1d.weight.detach().float().reshape(cd, self.d_conv).contiguous(),
                    conv_b=None if self.conv1d.bias is None else self.conv1d.bias.detach().float().contiguous(),
                    dt_w=sel, dt_b=bias.contiguous(), A=A, D=Dc.contiguous())

    @torch.no_grad()
    def step(self, u_t, conv_state, ssm_state, consts=None):
        """One new position: u_t [B, d_model] -> [B, d_model]; both states are updated in place."""
        from .mamba import _linear_step
        k = consts if consts is not None else self.step_constants()
        N, T = self.d_state, conv_state.dtype
        if u_t.dtype != T:
            u_t = u_t.to(T)
        zxbcdt = _linear_step(u_t, self.in_proj)
        z, xBC, dt = torch.split(zxbcdt, [self.d_inner, self.d_inner + 2 * N, self.nheads], dim=-1)
        xBC = ops.conv_step(xBC, conv_state, k["conv_w"], k["conv_b"])
        x, Bv, Cv = torch.split(xBC, [self.d_inner, N, N], dim=-1)
        y = ops.ssm_step(x, dt.contiguous(), Bv, Cv, k["dt_w"], k["dt_b"], k["A"], k["D"], z, ssm_state)
        y = ops.rmsnorm_fn(y, self.norm.weight, None, self.norm.eps)[0]
        return _linear_step(y, self.out_proj)
