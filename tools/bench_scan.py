#!/usr/bin/env python
"""Kernel-only sweep of the selective-scan / conv kernels (BASELINE configs[4] and the repo's training shape).

    python tools/bench_scan.py [--shapes repo,long] [--dtype f32,bf16] [--fwd-variants 0,4,8,16] [--bwd-variants 0,1,2,4]

Times each C-ABI call with CUDA events on the launching stream (L2 flushed between iterations by writing a
256 MB buffer) and prints one JSON line per (op, shape, dtype, variant) with the achieved algorithmic GB/s.
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mamba_b200 import ops  # noqa: E402

SHAPES = {"repo": (2, 2054, 2048, 64), "long": (2, 8192, 2048, 16), "long1": (1, 8192, 2048, 16),
          "long8": (8, 8192, 2048, 16), "repo16": (2, 2054, 2048, 16)}


def algo_bytes(B, L, D, N, s):
    BLD, BLN = B * L * D, B * L * N
    return {"fwd": s * (4 * BLD + 2 * BLN) + 4 * (D * N + 2 * D), "bwd": s * (8 * BLD + 4 * BLN) + 8 * (D * N + 2 * D),
            "conv_fwd": s * 2 * BLD + 4 * 5 * D, "conv_bwd": s * 3 * BLD + 8 * 5 * D}


def timeit(fn, iters, flush):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="repo,long")
    ap.add_argument("--dtype", default="f32,bf16")
    ap.add_argument("--fwd-variants", default="0")
    ap.add_argument("--bwd-variants", default="0")
    ap.add_argument("--chunks", default="16")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ops", default="fwd,bwd,conv")
    args = ap.parse_args()
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    for sname in args.shapes.split(","):
        B, L, D, N = SHAPES[sname]
        for dname in args.dtype.split(","):
            dt = torch.float32 if dname == "f32" else torch.bfloat16
            s = 4 if dname == "f32" else 2
            g = torch.Generator(device=dev).manual_seed(0)
            u = torch.randn(B, L, D, device=dev, generator=g).to(dt)
            z = torch.randn(B, L, D, device=dev, generator=g).to(dt)
            dl = (torch.randn(B, L, D, device=dev, generator=g) - 4).to(dt)
            A = -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(D, 1)
            Bm = torch.randn(B, L, N, device=dev, generator=g).to(dt)
            Cm = torch.randn(B, L, N, device=dev, generator=g).to(dt)
            Dv = torch.ones(D, device=dev)
            bias = torch.zeros(D, device=dev)
            dout = torch.randn(B, L, D, device=dev, generator=g).to(dt)
            ab = algo_bytes(B, L, D, N, s)

            def report(op, variant, chunk, med, best):
                print(json.dumps({"op": op, "shape": sname, "BLDN": [B, L, D, N], "dtype": dname, "variant": variant,
                                  "chunk": chunk, "median_us": round(med, 1), "min_us": round(best, 1),
                                  "algo_GBs": round(ab[op] / med / 1e3, 1), "frac_hbm": round(ab[op] / med / 1e3 / peak, 4),
                                  "Gelem_s": round(B * L * D * N / med / 1e3, 1) if op in ("fwd", "bwd") else None}), flush=True)

            if "fwd" in args.ops:
                for v in map(int, args.fwd_variants.split(",")):
                    ops.SCAN_FWD_VARIANT = v
                    with torch.no_grad():
                        fn = lambda: ops.selective_scan_fn(u, dl, A, Bm, Cm, Dv, z=z, delta_bias=bias, delta_softplus=True)
                        try:
                            med, best = timeit(fn, args.iters, flush)
                            report("fwd", v, None, med, best)
                        except Exception as e:
                            print(json.dumps({"op": "fwd", "shape": sname, "variant": v, "error": str(e)[:200]}), flush=True)
                ops.SCAN_FWD_VARIANT = 0
            if "bwd" in args.ops:
                for chunk in map(int, args.chunks.split(",")):
                    for v in map(int, args.bwd_variants.split(",")):
                        ops.SCAN_BWD_VARIANT = v
                        leaves = [t.clone().requires_grad_(True) for t in (u, dl, A, Bm, Cm, Dv, z, bias)]
                        try:
                            out = ops.selective_scan_fn(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[5],
                                                        z=leaves[6], delta_bias=leaves[7], delta_softplus=True, chunk=chunk)
                            fn = lambda: torch.autograd.grad(out, leaves, dout, retain_graph=True)
                            ops.KERNEL_TIMES = None
                            med, best = timeit(fn, args.iters, flush)
                            report("bwd", v, chunk, med, best)
                        except Exception as e:
                            print(json.dumps({"op": "bwd", "shape": sname, "variant": v, "chunk": chunk, "error": str(e)[:200]}), flush=True)
                        del leaves
                ops.SCAN_BWD_VARIANT = 0
            if "conv" in args.ops:
                w = torch.randn(D, 1, 4, device=dev)
                cb = torch.zeros(D, device=dev)
                with torch.no_grad():
                    med, best = timeit(lambda: ops.causal_conv1d_silu_fn(u, w, cb), args.iters, flush)
                report("conv_fwd", 0, None, med, best)
                ul, wl, bl = u.clone().requires_grad_(True), w.clone().requires_grad_(True), cb.clone().requires_grad_(True)
                o = ops.causal_conv1d_silu_fn(ul, wl, bl)
                med, best = timeit(lambda: torch.autograd.grad(o, [ul, wl, bl], dout, retain_graph=True), args.iters, flush)
                report("conv_bwd", 0, None, med, best)


if __name__ == "__main__":
    main()
