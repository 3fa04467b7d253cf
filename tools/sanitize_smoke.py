#!/usr/bin/env python
"""One small call of every kernel at awkward shapes (ragged channel tile, odd d_state, short sequences, strided views)
— meant to run under `compute-sanitizer --tool memcheck` (see profiles/)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mamba_b200 import ops, train  # noqa: E402
from mamba_b200.models.mamba import Mamba, ModelArgs  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
for (B, L, D, N, dt) in [(2, 37, 40, 5, torch.float32), (1, 100, 96, 64, torch.float32), (2, 70, 64, 32, torch.bfloat16),
                         (1, 17, 36, 48, torch.float32), (2, 33, 64, 16, torch.bfloat16),
                         (2, 50, 72, 64, torch.bfloat16), (1, 21, 40, 64, torch.bfloat16)]:
    xz = torch.randn(B, L, 2 * D, device=dev, generator=g).to(dt).requires_grad_(True)
    xdbl = torch.randn(B, L, 4 + 2 * N, device=dev, generator=g).to(dt).requires_grad_(True)
    u, z = xz.split([D, D], dim=-1)
    _, Bm, Cm = xdbl.split([4, N, N], dim=-1)
    dl = (torch.randn(B, L, D, device=dev, generator=g) - 3).to(dt).requires_grad_(True)
    A = (-torch.rand(D, N, device=dev, generator=g) - 0.5).requires_grad_(True)
    Dv = torch.randn(D, device=dev, generator=g).requires_grad_(True)
    bias = torch.randn(D, device=dev, generator=g).requires_grad_(True)
    for chunk in (8, 16):
        for variant in ((0, 1) if N in (32, 64) else (0,)):   # 1: the fused backward, also with fp32 I/O
            ops.SCAN_BWD_VARIANT = variant
            y = ops.selective_scan_fn(u, dl, A, Bm, Cm, Dv, z=z, delta_bias=bias, delta_softplus=True, chunk=chunk)
            y.float().square().sum().backward()
    ops.SCAN_BWD_VARIANT = 0
    w = torch.randn(D, 1, 4, device=dev, generator=g).requires_grad_(True)
    cb = torch.randn(D, device=dev, generator=g).requires_grad_(True)
    c = ops.causal_conv1d_silu_fn(u, w, cb)
    c.float().sum().backward()
    ops.causal_conv1d_silu_prefill(u.detach(), w.detach(), cb.detach())
    ops.selective_scan_prefill(u.detach(), dl.detach(), A.detach(), Bm.detach(), Cm.detach(), Dv.detach(), z=z.detach(),
                               delta_bias=bias.detach(), delta_softplus=True)
    nw = torch.ones(2 * D, device=dev, requires_grad=True)
    yn, st = ops.rmsnorm_fn(xz, nw, xz.detach().float(), 1e-5, dt)
    (yn.float().sum() + st.float().sum()).backward()
torch.cuda.synchronize()
model = Mamba(ModelArgs(d_model=64, n_layer=2, vocab_size=17914, d_state=16, pad_vocab_size_multiple=1)).to(dev)
from mamba_b200 import synthetic, generate
src, trg, meta = (t.to(dev) for t in synthetic.batch(2, 45, seed=1))
with torch.autocast("cuda", dtype=torch.bfloat16):
    out = model(src, meta)
loss = train.loss_fn(src, trg, out)
loss.backward()
out = generate.generate_recurrent(model.eval(), src[:, :30], meta, 6, use_graph=False)
torch.cuda.synchronize()
print("sanitize_smoke: ok", float(loss), out.shape)
