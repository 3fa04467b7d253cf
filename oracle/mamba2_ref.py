"""CPU restatement of the layer the reference's SHIPPED model stacks: `mamba_ssm.Mamba2` as configured at
models/mamba/mamba.py:17-23 (d_model 1024, d_state 64, d_conv 4, expand 2; library defaults headdim 64, ngroups 1,
rmsnorm gated, bias False, conv_bias True).

PARITY UNPINNED.  mamba_ssm is a third-party dependency that is not under /root/reference (requirements.txt:59 points
at a local checkout, no version) and cannot be installed here (GPU-only build, no network).  This file restates the
published Mamba-2 / SSD recurrence (Dao & Gu 2024, "Transformers are SSMs", the `Mamba2` module's reference path:
split -> causal conv + SiLU -> per-head scalar-decay state update -> gated RMSNorm -> out_proj) as a plain loop over
time.  What IS pinned by the reference: the parameter names and shapes (state_dict layout) through the parameter count
the reference prints for the shipped model, 101,972,666 (scripts/Test Accuracy.ipynb:52; tests/test_host.py).

Test infrastructure only: imported by tests/, never by the product package.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class Mamba2Ref(nn.Module):
    def __init__(self, d_model, d_state=64, d_conv=4, expand=2, headdim=64, ngroups=1, dt_min=0.001, dt_max=0.1,
                 dt_init_floor=1e-4, A_init_range=(1, 16), layer_idx=None):
        super().__init__()
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = expand * d_model
        self.headdim, self.ngroups = headdim, ngroups
        self.nheads = self.d_inner // headdim
        self.layer_idx = layer_idx
        d_in_proj = 2 * self.d_inner + 2 * ngroups * d_state + self.nheads
        self.in_proj = nn.Linear(d_model, d_in_proj, bias=False)
        conv_dim = self.d_inner + 2 * ngroups * d_state
        self.conv1d = nn.Conv1d(conv_dim, conv_dim, bias=True, kernel_size=d_conv, groups=conv_dim, padding=d_conv - 1)
        dt = torch.exp(torch.rand(self.nheads) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
        dt = torch.clamp(dt, min=dt_init_floor)
        self.dt_bias = nn.Parameter(dt + torch.log(-torch.expm1(-dt)))           # inverse softplus
        A = torch.empty(self.nheads).uniform_(*A_init_range)
        self.A_log = nn.Parameter(torch.log(A))
        self.D = nn.Parameter(torch.ones(self.nheads))
        self.norm = nn.Module()
        self.norm.weight = nn.Parameter(torch.ones(self.d_inner))                # RMSNormGated weight
        self.norm_eps = 1e-5
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=False)

    def forward(self, u):
        Bsz, L, _ = u.shape
        H, P, N, G = self.nheads, self.headdim, self.d_state, self.ngroups
        zxbcdt = self.in_proj(u)
        z, xBC, dt = torch.split(zxbcdt, [self.d_inner, self.d_inner + 2 * G * N, H], dim=-1)
        dt = F.softplus(dt + self.dt_bias)                                       # [B, L, H]
        xBC = F.silu(self.conv1d(xBC.transpose(1, 2))[..., :L].transpose(1, 2))
        x, Bm, Cm = torch.split(xBC, [self.d_inner, G * N, G * N], dim=-1)
        A = -torch.exp(self.A_log.float())                                       # [H]
        x = x.reshape(Bsz, L, H, P)
        Bm, Cm = Bm.reshape(Bsz, L, G, N), Cm.reshape(Bsz, L, G, N)
        hpg = H // G                                                             # heads per group
        state = torch.zeros(Bsz, H, P, N, dtype=torch.float32, device=u.device)
        ys = []
        for t in range(L):
            a = torch.exp(dt[:, t].float() * A)                                  # [B, H] one decay per head
            Bt = Bm[:, t].float().repeat_interleave(hpg, dim=1)                  # [B, H, N]
            Ct = Cm[:, t].float().repeat_interleave(hpg, dim=1)
            dx = (dt[:, t].float()[..., None] * x[:, t].float())                 # [B, H, P]
            state = a[..., None, None] * state + dx[..., None] * Bt[:, :, None, :]
            y = torch.einsum("bhpn,bhn->bhp", state, Ct) + self.D.float()[None, :, None] * x[:, t].float()
            ys.append(y)
        y = torch.stack(ys, dim=1).reshape(Bsz, L, self.d_inner)
        y = y * F.silu(z.float())                                                # gate first (norm_before_gate=False)
        y = y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + self.norm_eps) * self.norm.weight.float()
        return self.out_proj(y.to(u.dtype))


class ShippedMambaRef(nn.Module):
    """models/mamba/mamba.py:8-35 with Mamba2Ref layers (no residuals, final LayerNorm, untied head with bias)."""

    def __init__(self, d_model=1024, n_layers=10, vocab_size=17914, metadata_vocab_size=568, d_state=64):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab_size, d_model)
        self.metadata_embedding = nn.Embedding(metadata_vocab_size, d_model)
        self.output_layer = nn.Linear(d_model, vocab_size)
        self.layers = nn.ModuleList([Mamba2Ref(d_model, d_state=d_state, d_conv=4, expand=2, layer_idx=i) for i in range(n_layers)])
        self.norm = nn.LayerNorm(d_model)

    def forward(self, tokens, meta):
        x = self.token_embedding(tokens)
        x = torch.cat((self.metadata_embedding(meta), x), dim=-2)
        for layer in self.layers:
            x = layer(x)
        x = self.norm(x)
        return self.output_layer(x)[:, meta.shape[-1]:]
