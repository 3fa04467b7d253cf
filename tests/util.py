"""Shared helpers for the parity tests."""
import torch


def assert_close(got, ref, rtol, atol_frac=1e-5, what="", atol_abs=0.0):
    """|got - ref| <= atol + rtol*|ref| with atol = atol_frac * max|ref| (magnitude-aware floor)."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} != {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    atol = atol_frac * float(ref.abs().max()) + atol_abs + 1e-30
    err = (got - ref).abs()
    bound = atol + rtol * ref.abs()
    bad = err > bound
    if bad.any():
        i = int((err - bound).argmax())
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance (rtol={rtol}, atol={atol:.3e}); "
            f"worst at flat index {i}: got {got.flatten()[i].item():.8g} ref {ref.flatten()[i].item():.8g} "
            f"max_abs_err {err.max().item():.3e} max|ref| {ref.abs().max().item():.3e}")


def scan_inputs(B, L, D, N, seed=0, dtype=torch.float32, device="cpu"):
    """Kernel-only sweep inputs of SURVEY.md §8(d): u,z ~ N(0,1); delta_raw = N(0,1) - 4 (softplus ~ 0.02);
    A = -(1..N) per channel (+ jitter so that rows differ); B,C ~ N(0,1); D = 1 (+ jitter)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(B, L, D, generator=g)
    z = torch.randn(B, L, D, generator=g)
    delta_raw = torch.randn(B, L, D, generator=g) - 4.0
    A = -(torch.arange(1, N + 1, dtype=torch.float32).repeat(D, 1) * (1 + 0.1 * torch.rand(D, N, generator=g)))
    Bm = torch.randn(B, L, N, generator=g)
    Cm = torch.randn(B, L, N, generator=g)
    Dv = 1 + 0.1 * torch.randn(D, generator=g)
    bias = 0.5 * torch.randn(D, generator=g)
    t = dict(u=u, z=z, delta_raw=delta_raw, A=A, B=Bm, C=Cm, D=Dv, bias=bias)
    out = {}
    for k, v in t.items():
        if k in ("A", "D", "bias"):
            out[k] = v.to(device)
        else:
            out[k] = v.to(dtype).to(device)
    return out
