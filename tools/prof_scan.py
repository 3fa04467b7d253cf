#!/usr/bin/env python
"""One selective-scan launch (forward and optionally backward) of a named shape for ncu captures.

    python tools/prof_scan.py --shape long --dtype f32 --fwd-variant 100 [--bwd] [--bwd-variant 0] [--train]
"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mamba_b200 import ops  # noqa: E402

SHAPES = {"repo": (2, 2054, 2048, 64), "long": (2, 8192, 2048, 16), "long8": (8, 8192, 2048, 16)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="repo")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--fwd-variant", type=int, default=0)
    ap.add_argument("--bwd-variant", type=int, default=0)
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--chunk", type=int, default=16)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    B, L, D, N = SHAPES[a.shape]
    dt = torch.float32 if a.dtype == "f32" else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    u, z, dl = r(B, L, D).to(dt), r(B, L, D).to(dt), (r(B, L, D) - 4).to(dt)
    A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(D, 1)
    Bm, Cm, Dv, bias, dout = r(B, L, N).to(dt), r(B, L, N).to(dt), torch.ones(D, device="cuda"), torch.zeros(D, device="cuda"), r(B, L, D).to(dt)
    ops.SCAN_FWD_VARIANT, ops.SCAN_BWD_VARIANT = a.fwd_variant, a.bwd_variant
    for _ in range(a.iters):
        if a.bwd:
            leaves = [t.clone().requires_grad_(True) for t in (u, dl, A, Bm, Cm, Dv, z, bias)]
            out = ops.selective_scan_fn(*leaves[:6], z=leaves[6], delta_bias=leaves[7], delta_softplus=True, chunk=a.chunk)
            torch.autograd.grad(out, leaves, dout)
        else:
            with torch.no_grad():
                ops.selective_scan_fn(u, dl, A, Bm, Cm, Dv, z=z, delta_bias=bias, delta_softplus=True)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
