#!/usr/bin/env python
"""Development check: the TMA-staged forward (variants 100+) against the LDGSTS forward (variant 0) — outputs, y_pre,
checkpoints, h_last — on ragged and full shapes, then a timing sweep.  (The parity tests proper, against the oracle,
are in tests/.)"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mamba_b200 import _lib, ops  # noqa: E402


def inputs(B, L, D, N, dt, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    u, z, dl = r(B, L, D).to(dt), r(B, L, D).to(dt), (r(B, L, D) - 3).to(dt)
    A = -(torch.rand(D, N, device="cuda", generator=g) * 8 + 0.05)
    return u, dl, A, r(B, L, N).to(dt), r(B, L, N).to(dt), torch.randn(D, device="cuda", generator=g), z, 0.3 * r(D)


def run(variant, t, chunk, softplus=True, use_z=True):
    u, dl, A, Bm, Cm, Dv, z, bias = t
    B, L, D = u.shape
    N = A.shape[1]
    n = _lib.lib().mamba_scan_ckpt_elems(B, L, D, N, chunk)
    ck = torch.zeros(n, dtype=torch.float32, device="cuda")
    yp = torch.zeros_like(u)
    hl = torch.zeros(B, D, N, device="cuda")
    out = ops._scan_fwd_raw(u, dl, A, Bm, Cm, Dv, z if use_z else None, bias, softplus, ck, chunk, h_last=hl, variant=variant,
                            y_pre=yp if use_z else None)
    torch.cuda.synchronize()
    return out.float(), yp.float(), ck, hl


def main():
    ok = True
    for (B, L, D, N) in [(2, 70, 64, 16), (1, 333, 100, 64), (3, 257, 32 * 3 + 8, 16), (2, 130, 96, 32), (2, 64, 64, 8),
                         (1, 50, 40, 128), (2, 2054, 256, 64)]:
        for dt in (torch.float32, torch.bfloat16):
            for chunk in (8, 16):
                t = inputs(B, L, D, N, dt)
                for sp, uz in ((True, True), (False, False)):
                    ref = run(0, t, chunk, sp, uz)
                    for v in (100, 101, 102, 110, 111, 112) + ((113, 114) if N > 32 else ()):
                        try:
                            got = run(v, t, chunk, sp, uz)
                        except Exception as e:
                            print("ERR", (B, L, D, N), dt, v, str(e)[:150])
                            ok = False
                            continue
                        errs = []
                        for a, b in zip(got, ref):
                            errs.append(float((a - b).abs().max() / b.abs().max().clamp_min(1e-20)))
                        tol = 3e-5 if dt == torch.float32 else 2e-2
                        bad = max(errs) > tol or any(e != e for e in errs)
                        if bad or v == 112:
                            print(("BAD " if bad else "ok  "), (B, L, D, N), str(dt)[6:], "chunk", chunk, "sp/z", sp, uz, "variant", v,
                                  " ".join(f"{e:.1e}" for e in errs), flush=True)
                        ok &= not bad
    print("ALL OK" if ok else "FAILURES")
    # timing
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, (B, L, D, N), dt in [("repo_bf16", (2, 2054, 2048, 64), torch.bfloat16), ("repo_f32", (2, 2054, 2048, 64), torch.float32),
                                   ("long_f32", (2, 8192, 2048, 16), torch.float32), ("long8_f32", (8, 8192, 2048, 16), torch.float32),
                                   ("long_bf16", (2, 8192, 2048, 16), torch.bfloat16)]:
        t = inputs(B, L, D, N, dt)
        u, dl, A, Bm, Cm, Dv, z, bias = t
        n = _lib.lib().mamba_scan_ckpt_elems(B, L, D, N, 16)
        ck = torch.zeros(n, dtype=torch.float32, device="cuda")
        yp = torch.zeros_like(u)
        out = torch.empty_like(u)
        for train in (True,):
            for v in (0, 100, 101, 102, 110, 111, 112) + ((113,) if N > 32 else ()):
                fn = lambda: ops._scan_fwd_raw(u, dl, A, Bm, Cm, Dv, z, bias, True, ck if train else None, 16, variant=v,
                                               y_pre=yp if train else None, out=out)
                for _ in range(2):
                    fn()
                ts = []
                for _ in range(7):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); fn(); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                print(json.dumps({"shape": name, "train": train, "variant": v, "median_us": round(ts[3], 1), "min_us": round(ts[0], 1)}), flush=True)


if __name__ == "__main__":
    main()
